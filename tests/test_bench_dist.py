"""bench.py's multi-rank plumbing on CPU: world_size-2 `gloo` process group, barrier, MAX / SUM
over ranks -- what turns per-rank device times into the one whole-job figure (the data path itself
has no collective: reads shard by batch, sets are replicated)."""
import multiprocessing as mp
import os
import socket
import sys

from conftest import ROOT


def _rank_main(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import bench

    assert bench.dist_env() == (rank, rank, world)
    d = bench.Dist(world, "gloo")
    d.barrier()
    ms = 10.0 + rank          # rank 1 is the slow one
    bases = 1000 * (rank + 1)
    out = (d.reduce(ms, "MAX"), d.reduce(bases, "SUM"))
    d.barrier()
    d.close()
    q.put((rank, out))


def test_two_rank_gloo_reduce_and_barrier():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # whole-job throughput = all ranks' bases over the slowest rank's time
    assert got[0] == got[1] == (11.0, 3000.0)


def test_optional_bench_legs_never_break_the_line():
    """bench.py's find-unique-kmers leg is an extra: whatever goes wrong inside it (here: no GPU)
    must come back as an `error` field, not as an exception that would cost the whole JSON line."""
    import sys

    from conftest import ROOT

    sys.path.insert(0, ROOT)
    import bench
    from trio_binning_b200 import kmers

    if kmers.device_count() > 0:
        pytest.skip("a GPU is present")
    out = bench.run_counter_leg(kmers)
    assert set(out) == {"error"} and "no CPU fallback" in out["error"]
