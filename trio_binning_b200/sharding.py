"""Read-sharded multi-GPU execution: one process per GPU, sets replicated, no
collective on the data path.

The reference is single-process (classify_by_kmers.py:99-117 is one loop over
reads); this module is the host logic that lets N processes share that loop:

* every rank parses the same input and cuts it into the same batches
  (deterministic: batches close when they hold >= batch_bases bases);
* batch i is classified by rank i % world on its own GPU, against its own
  replica of both k-mer sets;
* after every round of `world` batches the per-read results (counts, scores,
  bins: ~25 bytes per read) are gathered to rank 0, which writes the three
  output files and the TSV in input order -- byte-identical to a 1-GPU run.

torch.distributed is used for plumbing only (gather of small result arrays;
`nccl` or `gloo` process groups both work because only objects are gathered).
"""
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple


def batch_owner(batch_index: int, world: int) -> int:
    """Rank that classifies batch `batch_index` (round-robin deal)."""
    return batch_index % world


def iter_batches(reads: Iterable, batch_bases: int) -> Iterator[list]:
    """Cut a stream of records (objects with .seq) into consecutive batches; a batch
    closes as soon as it holds at least `batch_bases` bases."""
    batch, held = [], 0
    for read in reads:
        batch.append(read)
        held += len(read.seq)
        if held >= batch_bases:
            yield batch
            batch, held = [], 0
    if batch:
        yield batch


def iter_rounds(batches: Iterable[list], world: int) -> Iterator[List[list]]:
    """Group consecutive batches into rounds of `world` (the last may be short)."""
    current: List[list] = []
    for batch in batches:
        current.append(batch)
        if len(current) == world:
            yield current
            current = []
    if current:
        yield current


class Communicator:
    """Minimal wrapper over an initialised torch.distributed process group."""

    def __init__(self, rank: int = 0, world: int = 1, dist=None):
        self.rank, self.world, self.dist = rank, world, dist

    @classmethod
    def from_env(cls, backend: Optional[str] = None) -> "Communicator":
        """RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT as set by torchrun."""
        import os

        world = int(os.environ.get("WORLD_SIZE", "1"))
        rank = int(os.environ.get("RANK", "0"))
        if world == 1:
            return cls()
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend=backend or "gloo")
        return cls(rank, world, dist)

    def gather_to_root(self, obj):
        """List of every rank's object on rank 0, None elsewhere."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(obj, out, dst=0)
        return out

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()


def run_sharded(reads: Iterable, classify_fn: Callable[[Sequence[str]], list],
                emit_fn: Callable[[object, tuple], None], batch_bases: int,
                comm: Optional[Communicator] = None) -> Tuple[int, int]:
    """Drive the sharded loop.

    classify_fn(list of sequences) -> list of per-read result tuples (this rank's GPU);
    emit_fn(read, result) is called on rank 0 only, in input order.
    Returns (reads seen, reads classified by this rank).
    """
    comm = comm or Communicator()
    seen = mine = 0
    batch_index = 0
    for round_batches in iter_rounds(iter_batches(reads, batch_bases), comm.world):
        local = None
        for offset, batch in enumerate(round_batches):
            seen += len(batch)
            if batch_owner(batch_index + offset, comm.world) == comm.rank:
                local = classify_fn([r.seq for r in batch])
                mine += len(batch)
        gathered = comm.gather_to_root(local)
        if comm.rank == 0:
            for offset, batch in enumerate(round_batches):
                results = gathered[batch_owner(batch_index + offset, comm.world)]
                if results is None or len(results) != len(batch):
                    raise RuntimeError("sharded run: a rank returned a result of the wrong size")
                for read, result in zip(batch, results):
                    emit_fn(read, result)
        batch_index += len(round_batches)
    comm.barrier()
    return seen, mine
