"""Host-side logic that needs neither a GPU nor the reference: batch packing helpers,
synthetic workload generators, bench.py's reference arm and its JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def test_pack_reads_utf8_and_nul_semantics():
    from trio_binning_b200 import kmers

    bases, offsets = kmers.pack_reads(["ACGT", "", "GG"])
    assert bytes(bases) == b"ACGTGG" and offsets.tolist() == [0, 4, 4, 6]
    # the reference encodes UTF-8 and the C side stops at the first NUL (kmers.py:147, c/kmers.c:285)
    bases, offsets = kmers.pack_reads(["AC\x00GT", "é", "T"])
    assert bytes(bases) == b"AC" + "é".encode() + b"T" and offsets.tolist() == [0, 2, 4, 5]
    bases, offsets = kmers.pack_reads([])
    assert len(bases) == 0 and offsets.tolist() == [0]


def test_synth_is_deterministic_and_well_formed(tmp_path):
    from trio_binning_b200 import synth

    a1, b1 = synth.random_kmer_sets(5000, 4000, 21, seed=3)
    a2, b2 = synth.random_kmer_sets(5000, 4000, 21, seed=3)
    assert (a1 == a2).all() and (b1 == b2).all()
    assert len(np.unique(np.concatenate([a1, b1]))) == 9000           # distinct and disjoint
    ca, _ = synth.random_kmer_sets(1000, 10, 21, seed=3, canonical=True)
    assert (ca <= synth.revcomp_ints(ca, 21)).all()                    # all findable (quirk Q1)
    bases, offsets, hap = synth.random_reads(50, 500, 21, a1, b1, seed=4, plant_every=100, invalid_rate=0.01)
    assert offsets[0] == 0 and offsets[-1] == len(bases) and (np.diff(offsets) >= 22).all()
    assert set(hap.tolist()) <= {0, 1, 2}
    for r in range(50):                                                # junk never inside the first k bases
        head = bytes(bases[offsets[r]:offsets[r] + 21])
        assert set(head) <= set(b"ACGT")
    p = tmp_path / "k.txt"
    synth.write_kmer_file(str(p), a1[:10], 21)
    lines = p.read_text().split("\n")
    assert len(lines) == 11 and lines[-1] == "" and all(len(x) == 21 for x in lines[:10])
    assert synth.total_windows(np.array([0, 10, 40, 61]), 21) == 0 + 10 + 1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` on the small workload: one JSON line with the contract's keys."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "c2_small", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "bases/s" and line["value"] > 1e5
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in line["config"]


def test_bench_reference_arm_under_torchrun_prints_one_line():
    """N > 1: the driver launches the reference arm under torchrun; rank 0 alone works and prints
    the line, the other ranks exit 0 silently."""
    port = 29650 + os.getpid() % 200
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                        "c2_small", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [x for x in p.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2


def _dist_stdout_worker(rank, world, port, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench

    r, w = os.pipe()
    saved = os.dup(1)
    os.dup2(w, 1)  # capture fd 1 of this process
    try:
        d = bench.Dist(world, "gloo")
        total = d.reduce(rank + 1, "SUM")
        biggest = d.reduce(rank + 1, "MAX")
        d.barrier()
        if rank == 0:
            print(json.dumps({"sum": total, "max": biggest}), flush=True)
        d.close()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(w)
    out = os.read(r, 1 << 16).decode() if rank == 0 else ""
    queue.put((rank, out))


def test_bench_dist_plumbing_gloo_world2():
    """bench.py's torch.distributed plumbing (barrier, max/sum over ranks) on a world_size-2 gloo
    group, and that bringing the group up leaves stdout to the one JSON line."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_dist_stdout_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    lines = [x for x in got[0].splitlines() if x.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"sum": 3.0, "max": 2.0}
