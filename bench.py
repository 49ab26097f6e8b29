#!/usr/bin/env python
"""bench.py -- read-bases classified per second on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (config.workload = "c2"): BASELINE.json configs[1] -- 10 000 synthetic long
reads (gamma lengths, mean 20 kb, ~0.2 Gbases) against 10 M + 10 M distinct random
21-mers, per GPU (weak scaling: every rank classifies its own 10 000 reads against
its own replica of both sets; no collective on the data path).

A "step" is one pass of the hot path over that batch.
  value  : bases/s with the batch already resident in HBM (pack + index + classify
           + fix-up + decide kernels; CUDA events on the launching stream)
  e2e    : bases/s through the C-ABI classifier with ASCII reads in PINNED HOST buffers --
           H2D copies and the D2H copy of counts/scores/bins inside the timed region,
           one slot (stream) per batch in flight.  Three routes are timed and the best
           is reported (e2e.route; all three in e2e_variants): ASCII over the link and
           packed on the device, packed to 2 bits by the host cores first, and a mix of
           the two that keeps the PCIe link and the host cores busy together.
  roofline: classify kernel only, SURVEY 8(d)'s algorithmic bytes (32 B per window of the merged
           table + 0.25 B per base) over the kernel time against the measured HBM copy peak, with
           three explanatory views beside it: the bytes the kernel really needs from HBM, its
           sector-request rate against the measured L2 random-sector ceiling, and the busiest
           unit per ncu.
Inputs and tables exceed the 126 MB L2 (config.l2); value_l2_flushed is the same
measurement with L2 overwritten between the timed steps.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_reads, mean_len, k, n_a, n_b, fixed_len)
    "c2": (10_000, 20_000, 21, 10_000_000, 10_000_000, False),
    "c2_small": (1_000, 20_000, 21, 1_000_000, 1_000_000, False),   # quick local checks
    "c5_slice": (1_000_000, 150, 21, 10_000_000, 10_000_000, True),  # short-read shape
}
METRIC = "read_bases_classified_per_sec"
UNIT = "bases/s"


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class Dist:
    """torch.distributed as plumbing only: barrier + max/sum over ranks."""

    def __init__(self, world, backend):
        self.world = world
        self.torch = None
        if world > 1:
            import torch
            import torch.distributed as dist

            self.torch, self.dist = torch, dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            self.dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if backend == "nccl" else torch.device("cpu")
            if backend == "nccl":
                torch.cuda.set_device(self.dev)
            # NCCL prints its version banner on stdout when the first communicator is created;
            # stdout carries exactly one JSON line, so send fd 1 to stderr until that is done
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group(backend=backend)
                dist.barrier()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def reduce(self, value, op):
        if self.world == 1:
            return value
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled
    in-process through NVML every 10 ms (a looping `nvidia-smi -lms 20` child is the fallback;
    `value_sampler_off` repeats the HBM-resident measurement after the sampler has stopped, as a
    check that sampling does not disturb the kernels)."""

    REASONS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    Q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.thread = None
        self.sm, self.smax, self.reasons = [], [], set()
        self.path = os.path.join(tempfile.gettempdir(), f"tb_clocks_{os.getpid()}.csv")
        self.mode = os.environ.get("TB_CLOCK_SAMPLER", "nvml")

    def _nvml_loop(self, nv, handle):
        try:
            smax = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        except Exception:
            smax = None
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                if smax:
                    self.smax.append(smax)
                mask = int(get_reasons(handle)) if get_reasons else 0
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.010)

    def start(self):
        if self.mode == "nvml":
            try:
                import pynvml as nv

                nv.nvmlInit()
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu)
                self._stop = threading.Event()
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.thread = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_min_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "sampler": None}
        if self.thread:
            self._stop.set()
            self.thread.join(timeout=2)
            out["sampler"] = "nvml thread, 10 ms"
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.f.close()
            out["sampler"] = "nvidia-smi -lms 20"
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    self.sm.append(float(parts[1]))
                    self.smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                     parts[3:7]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
            os.unlink(self.path)
        else:
            return out
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), sm_min_mhz=float(min(self.sm)),
                       sm_max_mhz=float(max(self.smax)) if self.smax else None,
                       reasons=sorted(self.reasons), samples=len(self.sm))
        return out


def bind_to_gpu_numa_node(gpu_index):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity) BEFORE any pinned
    buffer is allocated, so that the staging memory is first-touched on the GPU's NUMA node:
    with 8 ranks streaming 55 GB/s each, copies that cross the socket interconnect cap the
    aggregate host-to-device rate."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = {i for i in range(n_cpus) if (int(words[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:  # noqa: BLE001  (no NVML, no permission: stay where we are)
        pass
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def build_workload(name, rank, tmp_dir):
    from trio_binning_b200 import synth

    n_reads, mean_len, k, n_a, n_b, fixed = WORKLOADS[name]
    keys_a, keys_b, path_a, path_b = synth.materialise_sets(tmp_dir, n_a, n_b, k, seed=42)
    bases, offsets, _ = synth.random_reads(n_reads, mean_len, k, keys_a, keys_b, seed=43 + rank,
                                           plant_every=1000, fixed_len=fixed)
    return dict(k=k, keys_a=keys_a, keys_b=keys_b, path_a=path_a, path_b=path_b,
                bases=bases, offsets=offsets, n_reads=n_reads, mean_len=mean_len, n_a=n_a, n_b=n_b)


# --------------------------------------------------------------------------- CPU arms
def _ref_worker_init():
    pass


_REF = {}


def _ref_count_range(args):
    lo, hi = args
    lib, A, B, bases, offsets = _REF["lib"], _REF["A"], _REF["B"], _REF["bases"], _REF["offsets"]
    ca = cb = 0
    n = 0
    for r in range(lo, hi):
        a, b = lib.count(bytes(bases[offsets[r]:offsets[r + 1]]), A, B)
        ca += a
        cb += b
        n += int(offsets[r + 1] - offsets[r])
    return n, ca, cb


def cpu_reference_setup(wl):
    """Load the reference's own CPU implementation (compiled from /root/reference/c/kmers.c
    into oracle/_ref by oracle/build_oracle.py) and build both sets with it; fall back to
    the oracle port if the compiled reference did not travel."""
    from oracle.oracle import CompiledReference, Oracle

    if CompiledReference.available():
        lib = CompiledReference()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)  # the reference prints progress to stderr
        try:
            A, B = lib.set_from_file(wl["path_a"]), lib.set_from_file(wl["path_b"])
        finally:
            os.dup2(saved, 2)
            os.close(devnull)
            os.close(saved)
        kind = "reference"
    else:
        lib = Oracle()
        A, B = lib.set_from_ints(wl["keys_a"], wl["k"]), lib.set_from_ints(wl["keys_b"], wl["k"])
        kind = "port"
    _REF.update(lib=lib, A=A, B=B, bases=wl["bases"], offsets=wl["offsets"])
    return kind


def cpu_time_reads(read_lo, read_hi, procs):
    """Classify reads [read_lo, read_hi) on `procs` host processes (fork: the sets built
    in the parent are shared copy-on-write).  Returns (bases, seconds)."""
    import multiprocessing as mp

    t0 = time.perf_counter()
    if procs <= 1:
        n, _, _ = _ref_count_range((read_lo, read_hi))
    else:
        bounds = np.linspace(read_lo, read_hi, procs * 4 + 1).astype(int)
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_ref_count_range, list(zip(bounds[:-1], bounds[1:])))
        n = sum(r[0] for r in res)
    return n, time.perf_counter() - t0


def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    tmp_dir = os.path.join(tempfile.gettempdir(), "tb_bench_data")
    wl = build_workload(args.workload, 0, tmp_dir)
    kind = cpu_reference_setup(wl)
    procs = os.cpu_count() or 1
    # bounded sample per step: ~2.5 s of work at ~0.8 Mbases/s/core
    mean_len = wl["mean_len"]
    reads_per_step = max(procs, int(2.5 * 0.8e6 * procs / mean_len))
    reads_per_step = min(reads_per_step, wl["n_reads"])
    times, bases_done = [], []
    cursor = 0
    for step in range(args.warmup + args.steps):
        lo = cursor % max(1, wl["n_reads"] - reads_per_step + 1)
        n, dt = cpu_time_reads(lo, lo + reads_per_step, procs)
        cursor += reads_per_step
        if step >= args.warmup:
            times.append(dt)
            bases_done.append(n)
    total_t, total_b = sum(times), sum(bases_done)
    value = total_b / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args, wl),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{reads_per_step} reads (~{total_b // args.steps} bases) per step, "
                                   f"{procs} forked processes over reads, sets built once (untimed)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, wl):
    return {"workload": f"{args.workload}: {wl['n_reads']} reads, mean {wl['mean_len']} bases, "
                        f"k={wl['k']}, {wl['n_a']}+{wl['n_b']} distinct random {wl['k']}-mers "
                        f"(raw orientation, ~50% findable), per GPU",
            "reads_per_gpu": wl["n_reads"], "bases_per_gpu": int(wl["offsets"][-1]),
            "k": wl["k"], "set_a": wl["n_a"], "set_b": wl["n_b"],
            "l2": "per step the kernels read 250 MB of fresh input (ASCII + packed) and probe 2 x 160 MB of "
                  "tables at random, both larger than the 126 MB L2, so steps are timed back to back; "
                  "value_l2_flushed repeats the measurement with a 512 MB buffer overwritten between steps",
            "parallelism": f"reads sharded over {args.gpus} GPU(s), sets replicated, no collective"}


# --------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args, rank, local_rank, world):
    from trio_binning_b200 import kmers, synth

    if kmers.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    numa = bind_to_gpu_numa_node(local_rank)
    dist = Dist(world, "nccl")
    kmers.set_device(local_rank)
    tmp_dir = os.path.join(tempfile.gettempdir(), f"tb_bench_data_r{rank}" if world > 1 else "tb_bench_data")
    t_setup = time.perf_counter()
    wl = build_workload(args.workload, rank, tmp_dir)
    k = wl["k"]
    A = kmers.create_kmer_hash_set(wl["path_a"])
    B = kmers.create_kmer_hash_set(wl["path_b"])
    build_ms = A.contents.build_ms + B.contents.build_ms
    bases, offsets = wl["bases"], wl["offsets"]
    total = int(offsets[-1])
    n_reads = wl["n_reads"]
    windows = synth.total_windows(offsets, k)

    n_slots = 5  # two for the double-buffered routes, up to five for the hybrid cycle below
    clf = kmers.Classifier(A, B, max_bases=total, max_reads=n_reads, n_slots=n_slots)
    for slot in range(2):  # fill the pinned staging buffers once: these ARE the host inputs
        hb, ho = clf.staging(slot)
        hb[:total] = bases
        ho[: n_reads + 1] = offsets
    setup_s = time.perf_counter() - t_setup

    # ---- end to end: pinned host -> H2D -> kernels -> D2H, two slots in flight
    def e2e_pass(n_steps):
        clf.mark(0, 0)
        for i in range(n_steps):
            slot = i & 1
            if i >= 2:
                clf.wait(slot)  # results of the batch that used this slot two steps ago
            clf.submit(slot, None, None, n_reads=n_reads)
        clf.mark(0, 1)
        clf.mark(1, 1)
        clf.wait(0)
        clf.wait(1)
        return max(clf.elapsed_ms(0, 0, 0, 1), clf.elapsed_ms(0, 0, 1, 1))

    e2e_pass(max(3, args.warmup))
    ca, cb, sa, sb, bins = (x.copy() for x in clf.result(0))
    hits_a = int(ca.sum())
    probes = 2 * windows - hits_a

    sampler = ClockSampler(local_rank)
    dist.barrier()
    sampler.start()
    e2e_ms = e2e_pass(args.steps)
    dist.barrier()
    e2e_ms = dist.reduce(e2e_ms, "MAX")

    # ---- same batch through the host-packed entry (north_star's "pinned, 2-bit-packed
    # ---- batches"): (a) packing timed on its own, (b) packed batches already in pinned
    # ---- memory -> H2D of 0.25 B/base -> kernels -> D2H, (c) ASCII in, host pack INSIDE
    # ---- the timed region, then (b)
    host_threads = max(1, (len(os.sched_getaffinity(0)) or 1) // world)  # ranks share the host cores
    pbs = [kmers.PackedBatch(total, n_reads) for _ in range(n_slots)]
    for pb in pbs:
        pb.pack(bases, offsets, k, threads=host_threads)
    t0 = time.perf_counter()
    for _ in range(10):
        pbs[0].pack(bases, offsets, k, threads=host_threads)
    pack_s = (time.perf_counter() - t0) / 10

    def packed_pass(n_steps, pack_inside):
        clf.mark(0, 0)
        for i in range(n_steps):
            slot = i & 1
            if i >= 2:
                clf.wait(slot)
            if pack_inside:
                pbs[slot].pack(bases, offsets, k, threads=host_threads)
            clf.submit_packed(slot, pbs[slot])
        clf.mark(0, 1)
        clf.mark(1, 1)
        clf.wait(0)
        clf.wait(1)
        return max(clf.elapsed_ms(0, 0, 0, 1), clf.elapsed_ms(0, 0, 1, 1))

    def hybrid_pass(n_cycles, n_packed):
        """ASCII in all cases.  One cycle = one batch on the ASCII route (slot 0: the PCIe link does
        the work, the device packs) followed by `n_packed` batches on the host-pack route (slots
        1..n_packed: host cores pack, 0.25 B/base cross the link) -- the two resources work in
        parallel, every batch in its own slot.  Returns ms for n_cycles * (1 + n_packed) batches."""
        cyc = 1 + n_packed
        clf.mark(0, 0)
        for i in range(n_cycles * cyc):
            slot = i % cyc
            if i >= cyc:
                clf.wait(slot)
            if slot == 0:
                clf.submit(0, None, None, n_reads=n_reads)
            else:
                pbs[slot].pack(bases, offsets, k, threads=host_threads)
                clf.submit_packed(slot, pbs[slot])
        for s_ in range(cyc):
            clf.mark(s_, 1)
        for s_ in range(cyc):
            clf.wait(s_)
        return max(clf.elapsed_ms(0, 0, s_, 1) for s_ in range(cyc))

    packed_pass(3, False)
    pca, pcb, *_ = clf.result(0)
    assert (pca == ca).all() and (pcb == cb).all(), "host-packed path differs from the ASCII path"
    dist.barrier()
    e2e_packed_ms = dist.reduce(packed_pass(args.steps, False), "MAX")
    dist.barrier()
    e2e_hostpack_ms = dist.reduce(packed_pass(max(3, args.steps // 5), True), "MAX") * args.steps / max(3, args.steps // 5)
    # the mix of the two routes that balances link and host cores depends on the box: time
    # 1 ASCII : 1..4 host-packed and keep the best (all are the same public calls)
    hybrid = {}
    for n_packed in range(1, n_slots):
        cyc = 1 + n_packed
        hybrid_pass(2, n_packed)
        dist.barrier()
        n_cycles = max(3, args.steps // (2 * cyc))
        ms = dist.reduce(hybrid_pass(n_cycles, n_packed), "MAX")
        hybrid[n_packed] = ms / (n_cycles * cyc)  # ms per batch
    hyb_n = min(hybrid, key=hybrid.get)
    e2e_hybrid_ms = hybrid[hyb_n] * args.steps
    # leave an ASCII batch resident in slot 0 for the device-only timing below
    clf.submit(0, None, None, n_reads=n_reads)
    clf.wait(0)

    # ---- HBM-resident: device part only, CUDA events on the slot's stream
    clf.rerun_resident(0, max(3, args.warmup))
    dist.barrier()
    res_ms, cls_ms = clf.rerun_resident(0, args.steps)
    dist.barrier()
    clocks = sampler.stop()
    quiet_ms, _ = clf.rerun_resident(0, args.steps)  # same measurement with the clock sampler off
    quiet_ms_max = dist.reduce(quiet_ms, "MAX")
    # same, with a 512 MB buffer overwritten between the timed steps (cold L2 every step)
    flush_ms, flush_cls_ms = clf.rerun_resident(0, args.steps, flush_bytes=512 << 20)
    res_ms_max = dist.reduce(res_ms, "MAX")
    flush_ms_max = dist.reduce(flush_ms, "MAX")
    total_all = dist.reduce(total, "SUM")
    probes_all = dist.reduce(probes, "SUM")

    # results unchanged by the re-runs (idempotence of the device pipeline)
    clf.submit(0, None, None, n_reads=n_reads)
    ca2, cb2, *_ = clf.result(0)
    assert (ca2 == ca).all() and (cb2 == cb).all(), "re-run changed the results"
    finfo = clf.filter_info()
    passes = clf.filter_passes(0)

    line = None
    if rank == 0:
        hbm_gbs, peak_src = measured_peaks()
        cls_ms_per = cls_ms / args.steps
        cls_s = cls_ms_per * 1e-3
        ceil_dram = ceil_l2 = None
        if not args.no_gather:
            # random-32-B-sector request ceilings of THIS GPU, same run: a table far larger
            # than L2 (HBM random-access bound) and one that stays L2-resident
            ceil_dram = max(kmers.gather_bench(4 << 30, mlp, tpb, bps, 48, 0)[0]
                            for mlp, tpb, bps in ((8, 256, 8), (16, 256, 4), (4, 512, 4)))
            ceil_l2 = max(kmers.gather_bench(32 << 20, mlp, tpb, bps, 256, 0)[0]
                          for mlp, tpb, bps in ((8, 256, 8), (16, 256, 4), (8, 128, 16)))
        if finfo["log2_bits"]:
            # filtered kernels: one L2 sector request per window (bitmap) + one (pair table) or
            # two (A and B) HBM sector requests per window that passed the bitmap
            pair = finfo["pair_buckets"] > 0
            table_probes = passes if pair else 2 * passes
            sector_requests = windows + table_probes
            # 32 B per probe + 0.25 B per base; with the merged A+B table one lookup per
            # window answers both sets, so 32 B per WINDOW (no phantom second probes)
            model_units = windows if pair else probes
            algo_bytes = 32.0 * model_units + 0.25 * total
            achieved = algo_bytes / cls_s / 1e9
            l2_bytes = 32.0 * sector_requests + 0.25 * total
            hbm_actual_bytes = 32.0 * table_probes + 0.25 * total + (1 << finfo["log2_bits"]) / 8
            roofline = {
                "bound": "hbm",
                "kernel": (f"k_classify_pair_mini<{k},8,{finfo['bits_per_key']}>" if finfo.get("minimizer") else
                           f"k_classify_pair<{k},16,{finfo['bits_per_key']}>") if pair else f"k_classify_filtered<{k},16>",
                "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "peak_source": peak_src,
                "traffic": None,
                "algorithmic_bytes_per_launch": algo_bytes,
                "algorithmic_bytes_definition": "SURVEY 8(d): 32 B x %s + 0.25 B x bases" % (
                    "windows (merged A+B table: one lookup per window)" if pair else "probes"),
                "note": "frac > 1 means the kernel no longer moves the bytes the 8(d) model charges: the bitmap "
                        "pre-filter (2^%d bits, %d bit(s) per key, %s, read evict_last) answers %.2f%% of the "
                        "windows from L1/L2, and only the survivors read %s from HBM. What the kernel does fetch "
                        "from HBM is in hbm_actual{} (and, measured, in traffic); its cache-side request rate "
                        "against the measured L2 random-sector ceiling is in l2_requests{}"
                        % (finfo["log2_bits"], finfo["bits_per_key"],
                           "sectors addressed by the k-mers' minimizers so that consecutive windows re-use them from L1"
                           if finfo.get("minimizer") else "L2-resident",
                           100.0 * (1 - passes / max(1, windows)),
                           "one bucket of the tagged A+B pair table" if pair else "one bucket of A and one of B"),
                "windows_per_launch": windows, "filter_passes_per_launch": passes,
                "kernel_ms": cls_ms_per, "kernel_share_of_step": cls_ms / res_ms,
                "reference_probes_per_launch": probes, "reference_probes_per_s": probes / cls_s,
                "hbm_actual": {
                    "what": "bytes this kernel must fetch from HBM per launch: 32 B per bitmap survivor, the "
                            "2-bit read stream, the bitmap once",
                    "bytes_per_launch": hbm_actual_bytes,
                    "achieved": hbm_actual_bytes / cls_s / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                    "frac": hbm_actual_bytes / cls_s / 1e9 / hbm_gbs,
                    "table_probes_per_s": table_probes / cls_s,
                    "random_sector_ceiling_per_s": ceil_dram,
                    "frac_of_random_sector_ceiling": (table_probes / cls_s / ceil_dram) if ceil_dram else None},
                "l2_requests": {
                    "what": "one 32-B bitmap sector request per window + one per table probe, against the "
                            "random-32-B-sector request ceiling of an L2-resident table measured in this run "
                            "(tb_gather_bench, 32 MB table, requests that bypass L1)",
                    "sector_requests_per_launch": sector_requests,
                    "achieved": l2_bytes / cls_s / 1e9,
                    "peak": (ceil_l2 * 32 / 1e9) if ceil_l2 else None, "unit": "GB/s",
                    "frac": (l2_bytes / cls_s / 1e9 / (ceil_l2 * 32 / 1e9)) if ceil_l2 else None},
            }
        else:
            algo_bytes = 32.0 * probes + 0.25 * total
            achieved = algo_bytes / cls_s / 1e9
            roofline = {"bound": "hbm", "kernel": f"k_classify<{k},4>", "achieved": achieved, "peak": hbm_gbs,
                        "unit": "GB/s", "frac": achieved / hbm_gbs, "peak_source": peak_src, "traffic": None,
                        "algorithmic_bytes_per_launch": algo_bytes, "probes_per_launch": probes,
                        "kernel_ms": cls_ms_per, "probes_per_s": probes / cls_s,
                        "kernel_share_of_step": cls_ms / res_ms,
                        "random_sector_ceiling_per_s": ceil_dram,
                        "frac_of_random_sector_ceiling": (probes / cls_s / ceil_dram) if ceil_dram else None}
        traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(traffic_file):
            with open(traffic_file) as f:
                t = json.load(f).get(roofline["kernel"].split("<")[0])
            if t:
                roofline["traffic"] = t["dram_bytes_per_launch"]
                roofline["traffic_source"] = t["source"]
                if "alu_pipe_pct" in t:
                    roofline["binding_resource"] = {
                        "unit": "integer ALU pipe", "pct_of_peak": t["alu_pipe_pct"],
                        "l1_sector_hit_pct": t["l1_sector_hit_pct"], "source": "ncu (traffic_source)",
                        "what": "the busiest unit per ncu; the minimizer-addressed sectors let L1 serve most "
                                "bitmap requests, which is why l2_requests.frac can exceed 1 as well"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            kind = cpu_reference_setup(wl)
            sample_reads = max(1, int(15e6 / wl["mean_len"]))  # ~15 Mbases: 15-20 s on one core
            sample_reads = min(sample_reads, n_reads)
            nb, dt = cpu_time_reads(0, sample_reads, 1)
            cpu = {"value": nb / dt, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"first {sample_reads} reads ({nb} bases) of the same batch, same full-size "
                             f"sets, 1 process / 1 thread (the reference as shipped), {dt:.1f} s"}
            # parity spot check while we are here: the sample's counts equal the GPU's
            lib = _REF["lib"]
            for r in range(0, sample_reads, max(1, sample_reads // 25)):
                got = lib.count(bytes(bases[offsets[r]:offsets[r + 1]]), _REF["A"], _REF["B"])
                assert got == (ca[r], cb[r]), f"parity failure on read {r}: cpu {got} gpu {(ca[r], cb[r])}"
        d2h = n_reads * (4 + 4 + 8 + 8 + 1)
        ascii_h2d = total + 8 * (n_reads + 1)
        e2e_variants = {
            "ascii_device_pack": {
                "value": total_all / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": ascii_h2d, "d2h_bytes_per_step": d2h,
                "api": "tb_classifier_submit/wait (C ABI): ASCII reads in pinned host memory -> H2D 1 B/base -> "
                       "k_pack on the device -> classify -> D2H; 2 slots in flight"},
            "ascii_host_pack": {
                "value": total_all / (e2e_hostpack_ms / args.steps * 1e-3), "unit": UNIT,
                "ms_per_step": e2e_hostpack_ms / args.steps,
                "h2d_bytes_per_step": pbs[0].h2d_bytes, "d2h_bytes_per_step": d2h,
                "api": f"ASCII reads in pinned host memory -> tb_pack_reads on {host_threads} host threads INSIDE the "
                       "timed region -> tb_classifier_submit_packed (H2D 0.25 B/base) -> classify -> D2H"},
            "ascii_hybrid": {
                "value": total_all / (e2e_hybrid_ms / args.steps * 1e-3), "unit": UNIT,
                "ms_per_step": e2e_hybrid_ms / args.steps,
                "h2d_bytes_per_step": (ascii_h2d + hyb_n * pbs[0].h2d_bytes) // (1 + hyb_n),
                "d2h_bytes_per_step": d2h,
                "cycle": f"1 ASCII : {hyb_n} host-packed",
                "ms_per_batch_by_cycle": {f"1:{n}": round(v, 4) for n, v in hybrid.items()},
                "api": "ASCII reads in pinned host memory; in every cycle one batch takes the ASCII route "
                       f"(tb_classifier_submit) and {hyb_n} take the host-pack route (tb_pack_reads on {host_threads} "
                       "threads + tb_classifier_submit_packed), each in its own slot, all inside the timed region: "
                       "PCIe link and host cores work in parallel"},
        }
        best = max(e2e_variants, key=lambda name: e2e_variants[name]["value"])
        e2e_best = dict(e2e_variants[best], route=best)
        line = {
            "metric": METRIC, "value": total_all / (res_ms_max / args.steps * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": res_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, wl),
            "e2e": e2e_best,
            "e2e_variants": e2e_variants,
            "e2e_packed": {"value": total_all / (e2e_packed_ms / args.steps * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": pbs[0].h2d_bytes, "d2h_bytes_per_step": n_reads * (4 + 4 + 8 + 8 + 1),
                           "ms_per_step": e2e_packed_ms / args.steps,
                           "api": "tb_classifier_submit_packed: 2-bit batches packed by tb_pack_reads, already in "
                                  "pinned host memory when the timed region starts (packing is reported in host_pack)"},
            "host_pack": {"value": total / pack_s, "unit": UNIT, "threads": host_threads,
                          "what": "tb_pack_reads (AVX-512/AVX2, multi-threaded): ASCII -> 2-bit pinned batch, this rank, mean of 10 calls"},
            "gpu_launches": 5 * args.steps,
            "value_l2_flushed": total_all / (flush_ms_max / args.steps * 1e-3),
            "value_sampler_off": total_all / (quiet_ms_max / args.steps * 1e-3),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "probes_per_s": probes_all / (res_ms_max / args.steps * 1e-3),
            "set_build_ms": build_ms,
            "setup_s": setup_s,
            "cpu_affinity": f"{len(numa)} cpus next to the GPU (NVML)" if numa else "unchanged",
            "bins": {"A": int((bins == 65).sum()), "B": int((bins == 66).sum()), "U": int((bins == 85).sum())},
        }
        print(json.dumps(line), flush=True)
    clf.close()
    dist.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    return run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
