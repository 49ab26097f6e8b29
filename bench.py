#!/usr/bin/env python
"""bench.py -- read-bases classified per second on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (config.workload = "c3"): BASELINE.json configs[2] at its full SET sizes -- 150 M + 150 M
distinct 21-mers (4.8 GB tagged table + 0.5 GB bitmap per GPU) -- and, per step and per GPU, a slice
of its reads: 10 000 synthetic long reads (gamma lengths, mean 25 kb, ~0.25 Gbases).  Weak
scaling: every rank classifies its own reads against its own replica of both sets; no collective
on the data path.  configs[1], [3] and [4] (c2, c4, c5) are measured beside it at N = 1
(`other_workloads`).

A "step" is one pass of the hot path over that batch.
  value    bases/s with the batch already resident in HBM (pack + index + classify + fix-up +
           decide kernels; CUDA events on the launching stream), steps back to back
  value_sustained  the same over >= 2 s of back-to-back steps, clocks sampled
  e2e      bases/s through the C ABI from HOST buffers in the form the product's parser emits
           them -- pinned 2-bit-packed batches (tb_pack_reads) -- with the H2D copy of every
           batch and the D2H copy of counts/scores/bins inside the timed region, two slots in
           flight.  e2e_variants also times ASCII reads in pinned memory (packed on the device,
           or packed by the host cores inside the timed region).
  roofline the classify kernel against the resource that binds it on this workload: HBM for
           c3/c4/c5 (every run of windows that share a minimizer costs one random bitmap fetch),
           the integer ALU pipe for c2 (bitmap L2-resident).  See DESIGN.md section 4.
  pipeline file -> files through the native pipeline (parse, pack, classify, compress, write),
           measured here on a bounded synthetic FASTQ, stage times from tb_stream_stats.
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
    # name: k, set sizes, reads per step and GPU, mean read length, fixed length?, one planted k-mer per N bases
    "c3": dict(k=21, n_a=150_000_000, n_b=150_000_000, n_reads=10_000, mean_len=25_000, fixed=False, plant=1000,
               baseline="configs[2]: 150M+150M 21-mers at full size; reads: a 10 000-read slice of the 1 M x 25 kb"),
    "c2": dict(k=21, n_a=10_000_000, n_b=10_000_000, n_reads=10_000, mean_len=20_000, fixed=False, plant=1000,
               baseline="configs[1]: 10 k long reads (mean 20 kb) vs 10M+10M 21-mers, complete"),
    "c4": dict(k=31, n_a=400_000_000, n_b=400_000_000, n_reads=10_000, mean_len=25_000, fixed=False, plant=1000,
               baseline="configs[3]: k=31, 400M+400M at full size; reads: a 10 000-read slice"),
    "c5": dict(k=21, n_a=50_000_000, n_b=50_000_000, n_reads=2_000_000, mean_len=150, fixed=True, plant=100,
               baseline="configs[4]: 50M+50M 21-mers at full size; reads: 2 M of the 100 M 150-bp reads"),
    "c2_small": dict(k=21, n_a=1_000_000, n_b=1_000_000, n_reads=1_000, mean_len=20_000, fixed=False, plant=1000,
                     baseline="quick local check"),
}
METRIC = "read_bases_classified_per_sec"
UNIT = "bases/s"
SM_COUNT, ALU_LANES_PER_SM = 148, 64  # B200: 4 sub-partitions x 16 INT32 lanes


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
    in-process through NVML every 10 ms (a looping `nvidia-smi -lms 20` child is the fallback)."""

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
        self.sm, self.smax, self.reasons = [], [], set()
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
            self.thread = None
            out["sampler"] = "nvml thread, 10 ms"
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.proc = None
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
    buffer is allocated, so that the staging memory is first-touched on the GPU's NUMA node."""
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


def ncu_constants(workload):
    """Per-launch counters of the classify kernel from the committed ncu captures
    (profiles/ncu_kernels.json, written by tools/ncu_to_json.py from `ncu --set full` reports)."""
    path = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload)
    return None


def build_workload(name, rank):
    from trio_binning_b200 import synth

    w = WORKLOADS[name]
    keys_a, keys_b = synth.sorted_disjoint_sets(w["n_a"], w["n_b"], w["k"], seed=1000 + w["k"])
    bases, offsets, _ = synth.random_reads(w["n_reads"], w["mean_len"], w["k"], keys_a, keys_b, seed=43 + rank,
                                           plant_every=w["plant"], fixed_len=w["fixed"])
    return dict(w, name=name, keys_a=keys_a, keys_b=keys_b, bases=bases, offsets=offsets)


def workload_config(args, wl, world):
    return {"workload": f"{wl['name']}: {wl['n_a']}+{wl['n_b']} distinct {wl['k']}-mers (raw orientation, about half "
                        f"findable; sorted-gap generator, seed {1000 + wl['k']}), {wl['n_reads']} reads per step and GPU, "
                        f"mean {wl['mean_len']} bases, one planted k-mer per {wl['plant']} bases; {wl['baseline']}",
            "reads_per_gpu": wl["n_reads"], "bases_per_gpu": int(wl["offsets"][-1]),
            "k": wl["k"], "set_a": wl["n_a"], "set_b": wl["n_b"],
            "l2": "every step reads a fresh 0.25-0.3 GB batch and probes tables (and, from c3 up, a bitmap) far larger "
                  "than the 126 MB L2, so steps are timed back to back; value_l2_flushed repeats the measurement with "
                  "a 512 MB buffer overwritten between steps",
            "parallelism": f"reads sharded over {world} GPU(s), sets replicated, no collective"}


# --------------------------------------------------------------------------- CPU arms
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
    """The reference's own CPU implementation (compiled unmodified from /root/reference/c/kmers.c
    into oracle/_ref by oracle/build_oracle.py; the oracle port if that file did not travel) with
    both sets at the workload's full size.  The sets' tables are filled by the oracle's insert
    loop, which reproduces the reference's layout slot for slot (tests/test_oracle.py) without
    writing and re-parsing 6.6 GB of text: set construction is not what is timed."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle.oracle import CompiledReference, Oracle

    orc = Oracle()
    with ThreadPoolExecutor(2) as pool:  # ctypes drops the GIL: both tables fill side by side
        fa = pool.submit(orc.set_from_ints, wl["keys_a"], wl["k"])
        fb = pool.submit(orc.set_from_ints, wl["keys_b"], wl["k"])
        A, B = fa.result(), fb.result()
    if CompiledReference.available():
        lib, kind = CompiledReference(), "reference"
    else:
        lib, kind = orc, "port"
    _REF.update(lib=lib, A=A, B=B, bases=wl["bases"], offsets=wl["offsets"], orc=orc)
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
    wl = build_workload(args.workload, 0)
    t0 = time.perf_counter()
    kind = cpu_reference_setup(wl)
    setup_s = time.perf_counter() - t0
    procs = os.cpu_count() or 1
    # bounded sample per step: ~2.5 s of work at ~0.8 Mbases/s/core
    reads_per_step = max(procs, int(2.5 * 0.8e6 * procs / wl["mean_len"]))
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
        "data": "synthetic", "config": workload_config(args, wl, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{reads_per_step} reads (~{total_b // args.steps} bases) per step, "
                                   f"{procs} forked processes over reads, both full-size sets built once "
                                   f"(untimed, {setup_s:.0f} s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------- GPU arm
def classify_rooflines(wl, fi, windows, passes, runs, total, cls_ms, clocks, gather):
    """The classify kernel against the two resources that can bind it.  Algorithmic work per launch
    (restating SURVEY 8(d) for the merged, pre-filtered lookup):
      HBM:  32 B per bitmap sector fetch (one per run of consecutive windows of a lane that share a
            minimizer: `runs`, counted on the device) + 32 B per window that passes the bitmap (one
            bucket of the tagged A+B table) + 0.25 B per base of the 2-bit read stream
      ALU:  integer-pipe thread instructions per window, counted by ncu on this workload
            (sm__inst_executed_pipe_alu x 32 / windows; profiles/ncu_kernels.json)"""
    hbm_gbs, peak_src = measured_peaks()
    cls_s = cls_ms * 1e-3
    ncu = ncu_constants(wl["name"]) or {}
    algo_bytes = 32.0 * runs + 32.0 * passes + 0.25 * total
    hbm = {"bound": "hbm", "achieved": algo_bytes / cls_s / 1e9, "peak": hbm_gbs, "unit": "GB/s",
           "frac": algo_bytes / cls_s / 1e9 / hbm_gbs, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": algo_bytes,
           "algorithmic_bytes_definition": "32 B x bitmap sector fetches (runs of windows sharing a minimizer, "
                                           "counted on the device) + 32 B x windows passing the bitmap + 0.25 B x bases",
           "bitmap_sector_fetches_per_launch": runs, "fetches_per_window": runs / max(1, windows),
           "traffic": ncu.get("dram_bytes_per_launch"), "traffic_source": ncu.get("source"),
           "random_sector": {
               "what": "random 32-B sector fetches the kernel needs (bitmap runs + table probes) per second against "
                       "the measured ceiling for uniformly random sectors of a table of the bitmap's size; above 1 "
                       "where neighbouring misses share a 128-B line that L2 fetches whole",
               "achieved_per_s": (runs + passes) / cls_s, "ceiling_per_s": gather.get("dram"),
               "ceiling_table_bytes": gather.get("dram_table_bytes"),
               "frac": ((runs + passes) / cls_s / gather["dram"]) if gather.get("dram") else None}}
    if ncu.get("dram_bytes_per_launch"):
        # the captured launch's own traffic over ITS duration: how busy HBM really is
        hbm["traffic_gbs_at_capture"] = ncu["dram_bytes_per_launch"] / (ncu["duration_ms"] * 1e-3) / 1e9
        hbm["traffic_frac_of_peak_at_capture"] = hbm["traffic_gbs_at_capture"] / hbm_gbs
        hbm["traffic_over_algorithmic"] = ncu["dram_bytes_per_launch"] / algo_bytes
    alu = None
    if ncu.get("alu_thread_insts_per_window") and clocks.get("sm_mhz"):
        ops = ncu["alu_thread_insts_per_window"] * windows
        peak = SM_COUNT * ALU_LANES_PER_SM * clocks["sm_mhz"] * 1e6
        alu = {"bound": "int-alu", "achieved": ops / cls_s / 1e12, "peak": peak / 1e12, "unit": "Tinst/s",
               "frac": ops / cls_s / peak,
               "alu_thread_insts_per_window": ncu["alu_thread_insts_per_window"],
               "peak_definition": f"{SM_COUNT} SMs x {ALU_LANES_PER_SM} integer lanes x {clocks['sm_mhz']:.0f} MHz "
                                  "(median SM clock sampled during this run)",
               "source": ncu.get("source"), "alu_pipe_pct_at_capture": ncu.get("alu_pipe_pct")}
    # The binding resource is the one ncu found closest to its peak on this workload (ALU pipe
    # utilisation against measured DRAM traffic over the measured copy peak).  When the bitmap is
    # L2-resident (<= 64 MB, c2) its sector fetches never reach HBM and do not belong in the HBM model.
    if fi["log2_bits"] <= 29:
        hbm_bytes = 32.0 * passes + 0.25 * total + (1 << fi["log2_bits"]) / 8
        hbm.update(achieved=hbm_bytes / cls_s / 1e9, frac=hbm_bytes / cls_s / 1e9 / hbm_gbs,
                   algorithmic_bytes_per_launch=hbm_bytes,
                   algorithmic_bytes_definition="32 B x windows passing the bitmap + 0.25 B x bases + the bitmap once "
                                                "(it stays in L2: its sector fetches are not HBM traffic)")
        if ncu.get("dram_bytes_per_launch"):
            hbm["traffic_over_algorithmic"] = ncu["dram_bytes_per_launch"] / hbm_bytes
        rs = hbm["random_sector"]  # only the table probes reach HBM here
        rs["achieved_per_s"] = passes / cls_s
        rs["frac"] = (passes / cls_s / gather["dram"]) if gather.get("dram") else None
    main = hbm
    if alu and alu["frac"] > (hbm.get("traffic_frac_of_peak_at_capture") or hbm["frac"]):
        main = alu
    out = dict(main)
    out["kernel"] = f"k_classify_pair_mini<{wl['k']},8,{fi['bits_per_key']},2>" if fi.get("minimizer") else "k_classify*"
    out["kernel_ms"] = cls_ms
    out["windows_per_launch"] = windows
    out["bitmap_log2_bits"] = fi["log2_bits"]
    out["bitmap_pass_fraction"] = passes / max(1, windows)
    out["views"] = {"hbm": hbm, "int_alu": alu}
    return out


def measure_resident(clf, kmers, wl, steps, warmup, want_flush=True):
    """HBM-resident figures of one workload on an existing classifier whose slot 0 holds the batch."""
    clf.rerun_resident(0, max(3, warmup))
    res_ms, cls_ms = clf.rerun_resident(0, steps)
    out = {"res_ms": res_ms, "cls_ms": cls_ms}
    if want_flush:
        out["flush_ms"], _ = clf.rerun_resident(0, steps, flush_bytes=512 << 20)
    return out


def run_other_workload(name, kmers, synth, args, gather):
    """c2 / c4 / c5 beside the headline (N = 1 only): kernel-only and HBM-resident step rates."""
    wl = build_workload(name, 0)
    k = wl["k"]
    A = kmers.create_kmer_hash_set_from_ints(wl["keys_a"], k)
    B = kmers.create_kmer_hash_set_from_ints(wl["keys_b"], k)
    total, n_reads = int(wl["offsets"][-1]), wl["n_reads"]
    windows = synth.total_windows(wl["offsets"], k)
    clf = kmers.Classifier(A, B, max_bases=total, max_reads=n_reads, n_slots=2)
    clf.submit(0, wl["bases"], wl["offsets"])
    ca, cb, *_ = (x.copy() for x in clf.result(0))
    pb = kmers.PackedBatch(total, n_reads)
    pb.pack(wl["bases"], wl["offsets"], k)

    def packed_pass(n_steps):
        clf.mark(0, 0)
        for i in range(n_steps):
            slot = i & 1
            if i >= 2:
                clf.wait(slot)
            clf.submit_packed(slot, pb)
        clf.mark(0, 1)
        clf.mark(1, 1)
        clf.wait(0)
        clf.wait(1)
        return max(clf.elapsed_ms(0, 0, 0, 1), clf.elapsed_ms(0, 0, 1, 1))

    packed_pass(3)
    pa, pbb, *_ = clf.result(0)
    assert (pa == ca).all() and (pbb == cb).all(), "host-packed path differs from the ASCII path"
    steps = max(5, args.steps // 2)
    e2e_ms = packed_pass(steps)
    clf.submit_packed(0, pb)
    clf.wait(0)
    m = measure_resident(clf, kmers, wl, steps, args.warmup, want_flush=False)
    sampler = ClockSampler(0)
    sampler.start()
    sus_iters = int(min(4096, max(steps, 400.0 / (m["res_ms"] / steps))))
    sus_ms, _ = clf.rerun_resident(0, sus_iters)  # ~0.4 s: long enough for the clock sampler
    clocks = sampler.stop()
    fi, passes, runs = clf.filter_info(), clf.filter_passes(0), clf.sector_runs(0)
    roof = classify_rooflines(wl, fi, windows, passes, runs, total, m["cls_ms"] / steps, clocks, gather)
    out = {"workload": wl["baseline"], "k": k, "set_a": wl["n_a"], "set_b": wl["n_b"], "reads": n_reads, "bases": total,
           "value": total / (m["res_ms"] / steps * 1e-3), "unit": UNIT, "ms_per_step": m["res_ms"] / steps,
           "kernel_only": total / (m["cls_ms"] / steps * 1e-3),
           "value_sustained": total / (sus_ms / sus_iters * 1e-3),
           "e2e": {"value": total / (e2e_ms / steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": pb.h2d_bytes,
                   "d2h_bytes_per_step": n_reads * 25, "route": "packed"},
           "roofline": roof, "clocks": clocks, "steps": steps,
           "hits": [int(ca.sum()), int(cb.sum())]}
    clf.close()
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)
    return out


def run_counter_leg(kmers):
    """find-unique-kmers' compute beside the headline (N = 1 only): two read
    libraries (150-base reads, 30 x of a 5-Mbase genome and of a haplotype 0.5 % away from it, 0.5 %
    errors) counted at k = 21 from page-locked memory, both histograms, the reference's cut-off
    search, both subtractions.  tools/count_bench.py is the full-size version with a CPU port beside
    it.  Never fatal for the bench line: a failure is reported in place of the numbers."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import count_bench as cbench
        from trio_binning_b200 import find_unique_kmers as fuk

        k, genome_size, coverage, read_len, err = 21, 5_000_000, 30.0, 150, 0.005
        rng = np.random.default_rng(11)
        genome_a = cbench.ACGT[rng.integers(0, 4, genome_size)]
        genome_b = genome_a.copy()
        snp = rng.integers(0, genome_size, genome_size // 200)
        genome_b[snp] = cbench.ACGT[(rng.integers(1, 4, len(snp)) + cbench.CODE[genome_b[snp]]) % 4]
        libs = []
        for genome in (genome_a, genome_b):
            bases, offsets = cbench.library(genome, coverage, read_len, err, rng)
            pin = kmers.PinnedBuffer(len(bases), np.uint8)
            pin.array[:] = bases
            libs.append((pin, offsets))
        n_bases = sum(len(pin.array) for pin, _ in libs)
        n_slots = int((genome_size + n_bases / 2 * err * k) * 1.6)

        def one_pass():
            t = {}
            counters = [kmers.KmerCounter(k, n_slots) for _ in libs]
            t0 = time.perf_counter()
            for c, (pin, offsets) in zip(counters, libs):
                c.add_packed(pin.array, offsets)
            stats = [c.stats() for c in counters]
            t["count"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            hists = [c.histogram(256) for c in counters]
            t["histograms"] = time.perf_counter() - t0
            cuts = [fuk.analyze_histogram(fuk.histogram_lines(h)) for h in hists]
            t0 = time.perf_counter()
            unique = [len(counters[0].select_unique(counters[1], *cuts[0])), len(counters[1].select_unique(counters[0], *cuts[1]))]
            t["subtract_select_sort"] = time.perf_counter() - t0
            for c in counters:
                c.close()
            return t, stats, cuts, unique

        one_pass()
        t, stats, cuts, unique = one_pass()
        for pin, _ in libs:
            pin.free()
        return {"metric": "find_unique_kmers_read_bases_counted_per_second", "value": n_bases / t["count"], "unit": UNIT, "k": k,
                "workload": f"2 libraries x {coverage:g}x of a {genome_size / 1e6:g}-Mbase genome, {read_len}-base reads",
                "bases": n_bases, "windows_counted": sum(s["windows"] for s in stats),
                "distinct": [s["distinct"] for s in stats], "slot_bytes": kmers.lib.tb_counter_slot_bytes(k),
                "seconds": t, "cutoffs": [list(c) for c in cuts], "unique_kmers": unique,
                "input": "ASCII in page-locked host memory, H2D inside the timed region"}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def run_pipeline_leg(kmers, synth, A, B, k, keys_a, keys_b, gbases):
    """file -> files through the native pipeline on a bounded synthetic FASTQ (HiFi-like: 15 kb
    reads, run-length qualities), one GPU: plain, gzip and BGZF input; FASTQ-aware encoder and
    zlib -6 output; plain-text output.  Times are the pipeline's own (tb_stream_stats): start-up
    (interpreter, CUDA context, set build) is not in them."""
    tmp = tempfile.mkdtemp(prefix="tb_pipe_")
    import shutil

    free = shutil.disk_usage(tmp).free
    need = gbases * 1e9 * 2 * 3.2  # FASTQ + its two compressed copies + one set of outputs
    if need > 0.8 * free:
        gbases = max(0.05, gbases * 0.8 * free / need)
    n_reads = max(100, int(gbases * 1e9 / 15_000))
    t0 = time.perf_counter()
    bases, offsets, _ = synth.random_reads(n_reads, 15_000, k, keys_a, keys_b, seed=99, plant_every=1000)
    rng = np.random.default_rng(5)
    plain = os.path.join(tmp, "reads.fastq")
    with open(plain, "wb") as f:
        qual_pool = np.repeat(rng.integers(40, 94, size=1 << 16, dtype=np.uint8), rng.integers(1, 60, size=1 << 16))
        for r in range(n_reads):
            s = bases[offsets[r]:offsets[r + 1]]
            start = int(rng.integers(0, len(qual_pool) - len(s) - 1))
            f.write(b"@m64_%d/ccs np=12\n" % r + s.tobytes() + b"\n+\n" + qual_pool[start:start + len(s)].tobytes() + b"\n")
    size = os.path.getsize(plain)
    gz = os.path.join(tmp, "reads_gz.fastq.gz")
    bgzf = os.path.join(tmp, "reads_bgzf.fastq.gz")
    kmers.gzip_copy(plain, gz, level=6, threads=os.cpu_count() or 4, piece=1 << 20)
    kmers.gzip_copy(plain, bgzf, level=6, threads=os.cpu_count() or 4, piece=1 << 20, bgzf=True)
    gen_s = time.perf_counter() - t0
    total = int(offsets[-1])
    runs = []
    for tag, path, kw in (("plain in, gzip out (FASTQ-aware encoder)", plain, {}),
                          ("gzip -6 in, gzip out (FASTQ-aware encoder)", gz, {}),
                          ("BGZF in, gzip out (FASTQ-aware encoder)", bgzf, {}),
                          ("plain in, gzip out (zlib -6)", plain, dict(fastq_encoder=False, gzip_level=6)),
                          ("plain in, plain out (--no-gzip-output)", plain, dict(gzip_output=False))):
        d = tempfile.mkdtemp(dir=tmp)
        with open(os.path.join(d, "tsv"), "wb") as t:
            st = kmers.classify_fastx_file(path, A, B, os.path.join(d, "A"), os.path.join(d, "B"), os.path.join(d, "U"),
                                           tsv_fd=t.fileno(), batch_bases=64 << 20, **kw)
        runs.append({"run": tag, "wall_s": st["seconds_total"], "mbases_per_s": total / st["seconds_total"] / 1e6,
                     "input_mb_per_s": os.path.getsize(path) / st["seconds_total"] / 1e6,
                     "seconds": {key[8:]: round(st[key], 3) for key in st if key.startswith("seconds_") and key != "seconds_total"},
                     "output_bytes": st["output_gz_bytes"], "output_raw_bytes": st["output_raw_bytes"],
                     "reads_abu": [st["reads_a"], st["reads_b"], st["reads_u"]]})
        for name in os.listdir(d):
            os.unlink(os.path.join(d, name))
    for p_ in (plain, gz, bgzf):
        os.unlink(p_)
    return {"what": "native pipeline, file -> three files + TSV, one GPU, inside one process (start-up and set build "
                    "excluded); seconds.parse is the reader thread's wall time, inflate / deflate_cpu are summed over "
                    "their threads, gpu_wait over the device workers",
            "reads": n_reads, "bases": total, "fastq_bytes": size, "gzip_input": "4 MiB members, zlib -6, no index (read as one gzip stream: cut into 1 MB spans and decoded on several cores, csrc/gz_stream.cpp SpanInflate)", "host_threads": os.cpu_count(),
            "generate_s": round(gen_s, 1), "runs": runs}


def run_gpu_arm(args, rank, local_rank, world):
    from trio_binning_b200 import kmers, synth

    if kmers.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    numa = bind_to_gpu_numa_node(local_rank)
    dist = Dist(world, "nccl")
    kmers.set_device(local_rank)
    t_setup = time.perf_counter()
    wl = build_workload(args.workload, rank)
    k = wl["k"]
    A = kmers.create_kmer_hash_set_from_ints(wl["keys_a"], k)
    B = kmers.create_kmer_hash_set_from_ints(wl["keys_b"], k)
    build_ms = A.contents.build_ms + B.contents.build_ms
    bases, offsets = wl["bases"], wl["offsets"]
    total = int(offsets[-1])
    n_reads = wl["n_reads"]
    windows = synth.total_windows(offsets, k)

    n_slots = 3  # batches in flight on the end-to-end routes
    clf = kmers.Classifier(A, B, max_bases=total, max_reads=n_reads, n_slots=n_slots)
    for slot in range(n_slots):  # fill the pinned staging buffers once: these ARE the host inputs of the ASCII routes
        hb, ho = clf.staging(slot)
        hb[:total] = bases
        ho[: n_reads + 1] = offsets
    host_threads = max(1, (len(os.sched_getaffinity(0)) or 1) // world)  # ranks share the host cores
    pbs = [kmers.PackedBatch(total, n_reads) for _ in range(n_slots)]
    for pb in pbs:
        pb.pack(bases, offsets, k, threads=host_threads)
    t0 = time.perf_counter()
    for _ in range(5):
        pbs[0].pack(bases, offsets, k, threads=host_threads)
    pack_s = (time.perf_counter() - t0) / 5
    setup_s = time.perf_counter() - t_setup

    def in_flight_pass(n_steps, submit):
        """n_steps batches through the slots round-robin, each slot re-used once its previous batch's
        results are back; device time from the first copy to the last result (CUDA events)."""
        clf.mark(0, 0)
        for i in range(n_steps):
            slot = i % n_slots
            if i >= n_slots:
                clf.wait(slot)  # results of the batch that used this slot n_slots steps ago
            submit(slot)
        for s_ in range(n_slots):
            clf.mark(s_, 1)
        for s_ in range(n_slots):
            clf.wait(s_)
        return max(clf.elapsed_ms(0, 0, s_, 1) for s_ in range(n_slots))

    def ascii_pass(n_steps):
        return in_flight_pass(n_steps, lambda slot: clf.submit(slot, None, None, n_reads=n_reads))

    def packed_pass(n_steps, pack_inside):
        def submit(slot):
            if pack_inside:
                pbs[slot].pack(bases, offsets, k, threads=host_threads)
            clf.submit_packed(slot, pbs[slot])
        return in_flight_pass(n_steps, submit)

    warm = max(3, args.warmup)
    ascii_pass(warm)
    ca, cb, sa, sb, bins = (x.copy() for x in clf.result(0))
    hits_a = int(ca.sum())
    probes = 2 * windows - hits_a
    packed_pass(warm, False)
    pca, pcb, *_ = clf.result(0)
    assert (pca == ca).all() and (pcb == cb).all(), "host-packed path differs from the ASCII path"

    sampler = ClockSampler(local_rank)
    # ---- end to end, the product's route: pinned 2-bit batches -> H2D -> kernels -> D2H
    dist.barrier()
    sampler.start()
    e2e_packed_ms = dist.reduce(packed_pass(args.steps, False), "MAX")
    dist.barrier()
    clocks_e2e = sampler.stop()
    # ---- the same from ASCII in pinned memory: packed on the device, or by the host cores in the timed region
    e2e_ascii_ms = dist.reduce(ascii_pass(args.steps), "MAX")
    dist.barrier()
    few = max(3, args.steps // 5)
    e2e_hostpack_ms = dist.reduce(packed_pass(few, True), "MAX") * args.steps / few
    # ---- what the box can feed its GPUs with: plain pinned -> device copies, all ranks at once
    dist.barrier()
    h2d_gbs = kmers.h2d_bench(256 << 20, 60)
    h2d_sum, h2d_min = dist.reduce(h2d_gbs, "SUM"), dist.reduce(h2d_gbs, "MIN")
    # ---- HBM-resident, ASCII form (k_pack on the device is part of the step)
    clf.submit(0, None, None, n_reads=n_reads)
    clf.wait(0)
    clf.rerun_resident(0, warm)
    ascii_res_ms, _ = clf.rerun_resident(0, args.steps)
    ascii_res_ms = dist.reduce(ascii_res_ms, "MAX")
    # ---- HBM-resident, the product's form: a host-packed batch in HBM; device part only
    # ---- (index, classify, fix-up, decide), CUDA events on the slot's stream
    clf.submit_packed(0, pbs[0])
    clf.wait(0)
    clf.rerun_resident(0, warm)
    dist.barrier()
    sampler.start()
    res_ms, cls_ms = clf.rerun_resident(0, args.steps)
    dist.barrier()
    clocks = sampler.stop()
    flush_ms, _ = clf.rerun_resident(0, args.steps, flush_bytes=512 << 20)
    # ---- sustained: >= 2 s of back-to-back steps (the burst above lasts milliseconds)
    per_step = res_ms / args.steps
    sus_iters = int(min(4096, max(args.steps, 2200.0 / per_step)))
    sampler.start()
    sus_ms, sus_cls_ms = clf.rerun_resident(0, sus_iters)
    clocks_sus = sampler.stop()
    if not clocks.get("sm_mhz"):  # the burst is milliseconds long: the sampler may have missed it
        clocks = dict(clocks_sus, note="sampled during the sustained run that follows the timed steps")
    res_ms_max = dist.reduce(res_ms, "MAX")
    flush_ms_max = dist.reduce(flush_ms, "MAX")
    sus_ms_max = dist.reduce(sus_ms, "MAX")
    total_all = dist.reduce(total, "SUM")
    probes_all = dist.reduce(probes, "SUM")

    passes = clf.filter_passes(0)
    runs = clf.sector_runs(0)
    # results unchanged by the re-runs (idempotence of the device pipeline)
    clf.submit(0, None, None, n_reads=n_reads)
    ca2, cb2, *_ = clf.result(0)
    assert (ca2 == ca).all() and (cb2 == cb).all(), "re-run changed the results"
    finfo = clf.filter_info()

    line = None
    if rank == 0:
        gather = {}
        if not args.no_gather:
            # random-32-B-sector request ceilings of THIS GPU, same run (the denominator north_star
            # names): uniformly random sectors of a table as large as this workload's bitmap (4 GB if
            # the bitmap is L2-resident), read with the kernel's own bitmap load (32-bit, allocating in
            # L1, evict_last in L2) and with the 256-bit bucket load; the rate does not depend on the
            # load flavour, the loads in flight or the launch shape (profiles/r2_gather_ceiling.jsonl)
            bitmap_bytes = (1 << finfo["log2_bits"]) // 8 if finfo["log2_bits"] else 0
            table = bitmap_bytes if bitmap_bytes >= (256 << 20) else (4 << 30)
            gather["dram"] = max(kmers.gather_bench(table, mlp, tpb, bps, 48, variant)[0]
                                 for variant in (13, 0) for mlp, tpb, bps in ((8, 256, 8), (16, 256, 4)))
            gather["dram_table_bytes"] = table
            gather["l2"] = max(kmers.gather_bench(32 << 20, mlp, tpb, bps, 256, 0)[0]
                               for mlp, tpb, bps in ((8, 256, 8), (16, 256, 4), (8, 128, 16)))
        roofline = classify_rooflines(wl, finfo, windows, passes, runs, total, cls_ms / args.steps, clocks, gather)
        roofline["kernel_share_of_step"] = cls_ms / res_ms
        roofline["reference_probes_per_launch"] = probes
        roofline["reference_probes_per_s"] = probes / (cls_ms / args.steps * 1e-3)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.perf_counter()
            kind = cpu_reference_setup(wl)
            cpu_setup_s = time.perf_counter() - t0
            sample_reads = max(1, int(15e6 / wl["mean_len"]))  # ~15 Mbases: 15-20 s on one core
            sample_reads = min(sample_reads, n_reads)
            nb, dt = cpu_time_reads(0, sample_reads, 1)
            cpu = {"value": nb / dt, "unit": UNIT, "cores": 1, "kind": kind,
                   "sample": f"first {sample_reads} reads ({nb} bases) of the same batch, same full-size "
                             f"sets (built untimed in {cpu_setup_s:.0f} s), 1 process / 1 thread (the reference as "
                             f"shipped), {dt:.1f} s"}
            # parity spot check while we are here: the sample's counts equal the GPU's
            lib = _REF["lib"]
            for r in range(0, sample_reads, max(1, sample_reads // 25)):
                got = lib.count(bytes(bases[offsets[r]:offsets[r + 1]]), _REF["A"], _REF["B"])
                assert got == (ca[r], cb[r]), f"parity failure on read {r}: cpu {got} gpu {(ca[r], cb[r])}"
            orc = _REF.pop("orc")
            orc.set_free(_REF.pop("A"))
            orc.set_free(_REF.pop("B"))
        d2h = n_reads * (4 + 4 + 8 + 8 + 1)
        ascii_h2d = total + 8 * (n_reads + 1)

        def rate(ms):
            return total_all / (ms / args.steps * 1e-3)

        e2e_variants = {
            "packed": {
                "value": rate(e2e_packed_ms), "unit": UNIT, "ms_per_step": e2e_packed_ms / args.steps,
                "h2d_bytes_per_step": pbs[0].h2d_bytes, "d2h_bytes_per_step": d2h,
                "api": "tb_classifier_submit_packed/wait (C ABI): pinned 2-bit batches as the native pipeline's packer "
                       "emits them (tb_pack_reads) -> H2D 0.25 B/base -> index + classify + fix-up + decide -> D2H; "
                       f"{n_slots} slots in flight"},
            "ascii_device_pack": {
                "value": rate(e2e_ascii_ms), "unit": UNIT, "ms_per_step": e2e_ascii_ms / args.steps,
                "h2d_bytes_per_step": ascii_h2d, "d2h_bytes_per_step": d2h,
                "api": "tb_classifier_submit/wait: ASCII reads in pinned host memory -> H2D 1 B/base -> k_pack on the "
                       "device -> classify -> D2H; " + f"{n_slots} slots in flight"},
            "ascii_host_pack": {
                "value": rate(e2e_hostpack_ms), "unit": UNIT, "ms_per_step": e2e_hostpack_ms / args.steps,
                "h2d_bytes_per_step": pbs[0].h2d_bytes, "d2h_bytes_per_step": d2h,
                "api": f"ASCII reads in host memory -> tb_pack_reads on {host_threads} host threads INSIDE the "
                       "timed region -> tb_classifier_submit_packed -> classify -> D2H (what the native pipeline "
                       "does per batch)"},
        }
        line = {
            "metric": METRIC, "value": rate(res_ms_max), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": res_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, wl, world),
            "e2e": dict(e2e_variants["packed"], route="packed", clocks=clocks_e2e),
            "e2e_variants": e2e_variants,
            "h2d_ceiling": {
                "aggregate_gb_per_s": h2d_sum, "slowest_rank_gb_per_s": h2d_min,
                "e2e_h2d_gb_per_s": rate(e2e_packed_ms) * pbs[0].h2d_bytes / total / 1e9,
                "what": "plain cudaMemcpyAsync of a 256 MB page-locked buffer, 60 times, by every rank at the same time "
                        "(tb_h2d_bench), summed over ranks; beside it the host-to-device bytes per second the e2e "
                        "figure moves -- at N = 8 the two meet: the host side of the box, not the GPUs, bounds e2e"},
            "host_pack": {"value": total / pack_s, "unit": UNIT, "threads": host_threads,
                          "what": "tb_pack_reads (AVX-512/AVX2, multi-threaded): ASCII -> 2-bit pinned batch, this rank, mean of 5 calls"},
            "gpu_launches": 5 * args.steps,
            "value_ascii_resident": rate(ascii_res_ms),
            "value_definition": "a host-packed batch (the form the product's pipeline ships) resident in HBM: index + "
                                "classify + fix-up + decide kernels; value_ascii_resident starts from ASCII in HBM and "
                                "adds the device-side k_pack",
            "value_l2_flushed": rate(flush_ms_max),
            "value_sustained": {"value": total_all / (sus_ms_max / sus_iters * 1e-3), "unit": UNIT, "steps": sus_iters,
                                "seconds": sus_ms_max * 1e-3, "kernel_only": total / (sus_cls_ms / sus_iters * 1e-3),
                                "clocks": clocks_sus},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "probes_per_s": probes_all / (res_ms_max / args.steps * 1e-3),
            "set_build_ms": build_ms,
            "pair_and_bitmap_build_ms": finfo["build_ms"],
            "setup_s": setup_s,
            "cpu_affinity": f"{len(numa)} cpus next to the GPU (NVML)" if numa else "unchanged",
            "bins": {"A": int((bins == 65).sum()), "B": int((bins == 66).sum()), "U": int((bins == 85).sum())},
        }
        if world == 1 and not args.no_pipeline:
            line["pipeline"] = run_pipeline_leg(kmers, synth, A, B, k, wl["keys_a"], wl["keys_b"], args.pipeline_gbases)
    clf.close()
    kmers.destroy_kmer_hash_set(A)
    kmers.destroy_kmer_hash_set(B)
    if rank == 0:
        if world == 1 and not args.no_other_workloads:
            del wl, bases, offsets
            others = {}
            for name in ("c2", "c4", "c5"):
                if name != args.workload:
                    others[name] = run_other_workload(name, kmers, synth, args, gather)
            line["other_workloads"] = others
            line["find_unique_kmers"] = run_counter_leg(kmers)
        print(json.dumps(line), flush=True)
    dist.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    ap.add_argument("--pipeline-gbases", type=float, default=1.0)
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    return run_gpu_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
