"""Multi-GPU host logic on CPU: world_size-2 `gloo` process group, batches dealt
round-robin, results gathered to rank 0 and emitted in input order.  The per-rank
"GPU" is replaced by the oracle here (tests may use it; the product default is the
CUDA path) -- what is under test is the sharding / merge logic of
trio_binning_b200/sharding.py, which must reproduce the single-process output."""
import importlib.util
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _load(name):
    spec = importlib.util.spec_from_file_location(f"tb_{name}", os.path.join(ROOT, "trio_binning_b200", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


sharding = _load("sharding")
seq = _load("seq")


def _reads(case):
    return [seq.Read(f"r{i}", s, None) for i, s in enumerate(case["reads"])]


def _oracle_classifier(case):
    from oracle.oracle import Oracle

    orc = Oracle()
    a = orc.set_from_file(os.path.join(GOLDEN, case["set_a"]))
    b = orc.set_from_file(os.path.join(GOLDEN, case["set_b"]))
    sa, sb = orc.scaling_factors(a.contents.num_kmers, b.contents.num_kmers)

    def classify(seqs):
        counts = [orc.count(s, a, b) for s in seqs]
        ca = [c[0] for c in counts]
        cb = [c[1] for c in counts]
        xa, xb, bins = orc.decide(ca, cb, sa, sb)
        return [(ca[i], cb[i], float(xa[i]), float(xb[i]), chr(bins[i])) for i in range(len(seqs))]
    return classify


def _worker(rank, world, port, case, batch_bases, queue):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sharding.Communicator(rank, world, dist)
    emitted = []
    seen, mine = sharding.run_sharded(_reads(case), _oracle_classifier(case),
                                      lambda read, res: emitted.append((read.name, res)),
                                      batch_bases, comm)
    queue.put((rank, seen, mine, emitted))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_batching_is_deterministic_and_complete(golden_counts):
    reads = _reads(golden_counts[0])
    batches = list(sharding.iter_batches(reads, 2000))
    assert [r.name for b in batches for r in b] == [r.name for r in reads]
    assert all(sum(len(r.seq) for r in b[:-1]) < 2000 for b in batches)
    rounds = list(sharding.iter_rounds(batches, 3))
    assert sum(len(r) for r in rounds) == len(batches) and all(len(r) == 3 for r in rounds[:-1])
    assert [sharding.batch_owner(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
    assert list(sharding.iter_batches([], 10)) == []


def test_single_process_matches_golden(golden_counts):
    case = golden_counts[0]
    emitted = []
    sharding.run_sharded(_reads(case), _oracle_classifier(case), lambda r, res: emitted.append((r.name, res)), 5000)
    assert [list(e[1][:2]) for e in emitted] == case["counts"]
    assert [e[0] for e in emitted] == [f"r{i}" for i in range(len(case["reads"]))]


@pytest.mark.parametrize("batch_bases", [1, 3000, 10**9])
def test_world2_gloo_equals_single_process(golden_counts, batch_bases):
    case = golden_counts[2]
    single = []
    sharding.run_sharded(_reads(case), _oracle_classifier(case), lambda r, res: single.append((r.name, res)), batch_bases)
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, batch_bases, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, seen, mine, emitted = queue.get(timeout=120)
        got[rank] = (seen, mine, emitted)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = len(case["reads"])
    assert got[0][0] == got[1][0] == n            # every rank parses everything
    assert got[0][1] + got[1][1] == n             # each read classified exactly once
    if batch_bases == 1:
        assert abs(got[0][1] - got[1][1]) <= 6    # round-robin deal is balanced (empty reads ride along)
    assert got[1][2] == []                        # only rank 0 emits
    assert got[0][2] == single                    # identical output, input order
