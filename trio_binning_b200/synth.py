"""Seeded synthetic workloads for tests and bench.py (not on the product path).

Shapes follow SURVEY 8(d) / BASELINE.json `configs`:
  * k-mer sets: N distinct uniform-random k-mers per set, A and B disjoint,
    as RAW LSB-first integer encodings (c/kmers.c:50-72) and, on request, as
    text files (one k-mer per line) to exercise the real loader;
  * reads: gamma-distributed lengths around a stated mean (or fixed length),
    uniform ACGT, with k-mers of one haplotype planted about once per
    `plant_every` bases in either strand so that A, B and U reads all occur;
    optional non-ACGT bytes at positions >= k.
Everything is numpy-vectorised so the 0.2-Gbase config builds in seconds.
"""
import os

import numpy as np

_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)


def revcomp_ints(keys: np.ndarray, k: int) -> np.ndarray:
    """Integer encoding of the reverse complement of each encoded k-mer."""
    keys = keys.astype(np.uint64)
    out = np.zeros_like(keys)
    for j in range(k):
        digit = (keys >> np.uint64(2 * j)) & np.uint64(3)
        out |= (np.uint64(3) - digit) << np.uint64(2 * (k - 1 - j))
    return out


def random_kmer_sets(n_a: int, n_b: int, k: int, seed: int, canonical: bool = False):
    """Two disjoint sets of distinct random k-mers (uint64 raw encodings).

    canonical=False: stored as drawn, so about half are unfindable (quirk Q1:
    the reference inserts raw and looks up min(fwd, revcomp)).
    canonical=True: every k-mer is replaced by min(enc, enc(revcomp)) -- all findable.
    """
    rng = np.random.default_rng(seed)
    space_bits = 2 * k
    need = n_a + n_b
    if space_bits < 63 and need > (1 << space_bits) // 2:
        raise ValueError("k too small for that many distinct k-mers")
    keys = np.zeros(0, dtype=np.uint64)
    while len(keys) < need:
        draw = rng.integers(0, 1 << min(space_bits, 63), size=int((need - len(keys)) * 1.05) + 16,
                            dtype=np.uint64)
        if space_bits == 64:
            draw |= rng.integers(0, 2, size=len(draw), dtype=np.uint64) << np.uint64(63)
        if canonical:
            draw = np.minimum(draw, revcomp_ints(draw, k))
        keys = np.unique(np.concatenate([keys, draw]))
    keys = rng.permutation(keys)[:need]
    return keys[:n_a].copy(), keys[n_a:].copy()


def sorted_disjoint_sets(n_a: int, n_b: int, k: int, seed: int):
    """Two disjoint sets of distinct k-mers as ASCENDING uint64 arrays, in seconds even for the
    400 M + 400 M config: the keys are the running sums of random gaps (so they spread over the
    whole 4^k space, every low-order digit uniform) dealt alternately to A and B.  Sorted arrays
    let a checker test membership by binary search without building a table.  Raw orientation:
    about half of the keys are unfindable (quirk Q1)."""
    rng = np.random.default_rng(seed)
    n = n_a + n_b
    space = 1 << (2 * k)
    if n > space // 4:
        raise ValueError("k too small for that many distinct k-mers")
    mean_gap = space // (n + 1) // 16 * 15  # 6 % head-room: the sum's spread is 1/sqrt(3n) of it
    keys = rng.integers(1, max(2, 2 * mean_gap - 1), size=n, dtype=np.uint64)
    np.cumsum(keys, out=keys)
    while int(keys[-1]) >= space:  # only for a handful of keys
        keys = np.unique(keys >> np.uint64(1))
        n = len(keys)
    # deal: pairs (A, B) while both remain, the rest to the larger set
    m = min(n_a, n_b, n // 2)
    a = keys[0:2 * m:2]
    b = keys[1:2 * m:2]
    rest = keys[2 * m:]
    if n_a > n_b:
        a = np.concatenate([a, rest])
    elif n_b > n_a:
        b = np.concatenate([b, rest])
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def ints_to_ascii(keys: np.ndarray, k: int) -> np.ndarray:
    """(N, k) uint8 matrix of the k-mer letters."""
    keys = keys.astype(np.uint64)
    out = np.empty((len(keys), k), dtype=np.uint8)
    for j in range(k):
        out[:, j] = _LETTERS[((keys >> np.uint64(2 * j)) & np.uint64(3)).astype(np.intp)]
    return out


def write_kmer_file(path: str, keys: np.ndarray, k: int, chunk: int = 1 << 22) -> None:
    """One k-mer per line, '\\n'-terminated: the format the reference's loader reads."""
    with open(path, "wb") as f:
        for i in range(0, len(keys), chunk):
            mat = np.empty((min(chunk, len(keys) - i), k + 1), dtype=np.uint8)
            mat[:, :k] = ints_to_ascii(keys[i:i + chunk], k)
            mat[:, k] = 10
            f.write(mat.tobytes())


def read_lengths(n_reads: int, mean_len: int, rng, fixed: bool = False, min_len: int = 64) -> np.ndarray:
    if fixed:
        return np.full(n_reads, mean_len, dtype=np.int64)
    # gamma(shape 2) has the long right tail of PacBio/ONT length distributions
    lens = rng.gamma(2.0, mean_len / 2.0, size=n_reads).astype(np.int64)
    return np.maximum(lens, min_len)


def random_reads(n_reads: int, mean_len: int, k: int, keys_a: np.ndarray, keys_b: np.ndarray,
                 seed: int, plant_every: int = 1000, fixed_len: bool = False,
                 invalid_rate: float = 0.0):
    """Returns (bases uint8[total], offsets int64[n+1], hap int8[n]) where hap is the
    haplotype whose k-mers were planted in the read (0 = A, 1 = B, 2 = none)."""
    rng = np.random.default_rng(seed)
    lens = read_lengths(n_reads, mean_len, rng, fixed=fixed_len, min_len=max(k + 1, 32))
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    bases = _LETTERS[rng.integers(0, 4, size=total, dtype=np.uint8)]
    hap = rng.integers(0, 3, size=n_reads).astype(np.int8)
    if plant_every > 0:
        for h, keys in ((0, keys_a), (1, keys_b)):
            reads_h = np.nonzero(hap == h)[0]
            if len(reads_h) == 0 or len(keys) == 0:
                continue
            n_plants = np.maximum((lens[reads_h] - k) // plant_every, 1)
            read_of_plant = np.repeat(reads_h, n_plants)
            pos_in_read = (rng.random(len(read_of_plant)) * (lens[read_of_plant] - k + 1)).astype(np.int64)
            chosen = keys[rng.integers(0, len(keys), size=len(read_of_plant))]
            flip = rng.random(len(chosen)) < 0.5
            chosen = np.where(flip, revcomp_ints(chosen, k), chosen)
            letters = ints_to_ascii(chosen, k)
            start = offsets[read_of_plant] + pos_in_read
            idx = start[:, None] + np.arange(k)[None, :]
            bases[idx.ravel()] = letters.ravel()
    if invalid_rate > 0:
        n_bad = int(total * invalid_rate)
        pos = rng.integers(0, total, size=n_bad)
        read_idx = np.searchsorted(offsets, pos, side="right") - 1
        keep = (pos - offsets[read_idx]) >= k  # never inside the first k bases (Q3-UB)
        junk = np.frombuffer(b"NnacgtRY", dtype=np.uint8)
        bases[pos[keep]] = junk[rng.integers(0, len(junk), size=int(keep.sum()))]
    return bases, offsets, hap


def genome_kmer_set(genome: np.ndarray, k: int) -> np.ndarray:
    """All k-mers of a genome (uint8 letters) as the classifier finds them: the LSB-first
    encoding of each window or of its reverse complement, whichever is smaller (so every stored
    k-mer is findable, quirk Q1), distinct and sorted."""
    code = np.zeros(256, dtype=np.uint64)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    c = code[genome]
    n = len(c) - k + 1
    fwd = np.zeros(n, dtype=np.uint64)
    rc = np.zeros(n, dtype=np.uint64)
    for j in range(k):
        fwd |= c[j:j + n] << np.uint64(2 * j)
        rc |= (np.uint64(3) - c[j:j + n]) << np.uint64(2 * (k - 1 - j))
    return np.unique(np.minimum(fwd, rc))


def mosaic_reads(n_reads: int, mean_len: int, k: int, genome_a: np.ndarray, genome_b: np.ndarray, seed: int,
                 fraction: float, segment: int = 200):
    """Reads whose windows hit a set at a chosen density: random sequence with `fraction` of each
    read overwritten by `segment`-base pieces of its haplotype's genome (either strand).  With sets
    from genome_kmer_set about fraction * (segment - k + 1) / segment of all windows are hits.
    Returns (bases, offsets, hap) like random_reads."""
    rng = np.random.default_rng(seed)
    lens = read_lengths(n_reads, mean_len, rng, min_len=max(k + 1, segment + 1))
    offsets = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    bases = _LETTERS[rng.integers(0, 4, size=total, dtype=np.uint8)]
    hap = rng.integers(0, 2, size=n_reads).astype(np.int8)
    comp = np.zeros(256, dtype=np.uint8)
    comp[_LETTERS] = _LETTERS[::-1]
    for h, genome in ((0, genome_a), (1, genome_b)):
        reads_h = np.nonzero(hap == h)[0]
        n_seg = rng.poisson(lens[reads_h] * fraction / segment)
        read_of = np.repeat(reads_h, n_seg)
        if len(read_of) == 0:
            continue
        pos = (rng.random(len(read_of)) * (lens[read_of] - segment)).astype(np.int64) + offsets[read_of]
        src = rng.integers(0, len(genome) - segment, size=len(read_of))
        idx_dst = pos[:, None] + np.arange(segment)[None, :]
        piece = genome[src[:, None] + np.arange(segment)[None, :]]
        flip = rng.random(len(read_of)) < 0.5
        piece[flip] = comp[piece[flip][:, ::-1]]
        bases[idx_dst.ravel()] = piece.ravel()
    return bases, offsets, hap


def total_windows(offsets: np.ndarray, k: int) -> int:
    lens = np.diff(offsets)
    return int(np.maximum(lens - k + 1, 0).sum())


def materialise_sets(tmp_dir: str, n_a: int, n_b: int, k: int, seed: int, canonical: bool = False):
    """Generate the two sets and write them as text files; returns (keys_a, keys_b, path_a, path_b)."""
    os.makedirs(tmp_dir, exist_ok=True)
    keys_a, keys_b = random_kmer_sets(n_a, n_b, k, seed, canonical=canonical)
    tag = f"k{k}_{n_a}_{n_b}_s{seed}{'_canon' if canonical else ''}"
    path_a = os.path.join(tmp_dir, f"hapA_{tag}.txt")
    path_b = os.path.join(tmp_dir, f"hapB_{tag}.txt")
    for path, keys in ((path_a, keys_a), (path_b, keys_b)):
        if not (os.path.exists(path) and os.path.getsize(path) == len(keys) * (k + 1)):
            write_kmer_file(path, keys, k)
    return keys_a, keys_b, path_a, path_b
