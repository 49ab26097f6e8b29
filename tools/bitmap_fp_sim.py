"""False-positive rate of the minimizer-addressed bitmap pre-filter, simulated in numpy with the
kernel's own functions (classify.cu: mmer_mix, coord_of_key, mini_word_index, filter_bits), so
that changes to those hashes can be judged without a GPU.  Random keys, random probe reads.

usage: bitmap_fp_sim.py [k] [n_keys] [log2_bits]      (defaults: 21 20000000 29 = BASELINE C2)
Prints the measured pass rate of non-member windows, the bitmap fill, the rate a perfectly uniform
addressing would give at that fill (fill^2), and how often consecutive windows change sector."""
import sys
import time

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def mul32(a, c):
    return (a * np.uint64(c)) & M32


def umulhi(a, b):
    return (a * np.uint64(b)) >> np.uint64(32)


def mmer_mix(c):
    c = mul32(c, 0x9E3779B1)
    c ^= c >> np.uint64(15)
    c = mul32(c, 0x85EBCA77)
    c ^= c >> np.uint64(13)
    return c


def revcomp_m(f, m):
    x = (~f) & ((np.uint64(1) << np.uint64(2 * m)) - np.uint64(1))
    out = np.zeros_like(x)
    for i in range(m):
        out |= ((x >> np.uint64(2 * i)) & np.uint64(3)) << np.uint64(2 * (m - 1 - i))
    return out


def coords(keys, k, m):
    """min_hash (minimizer) and fh (first + last canonical m-mer hash) of each key"""
    mmask = np.uint64((1 << (2 * m)) - 1)
    mn = first = last = None
    for i in range(k - m + 1):
        f = (keys >> np.uint64(2 * i)) & mmask
        h = mmer_mix(np.minimum(f, revcomp_m(f, m)))
        mn = h if mn is None else np.minimum(mn, h)
        if i == 0:
            first = h
        last = h
    return mn, (first + last) & M32


def bit_addresses(mn, fh, n_sectors):
    sector = umulhi(mul32(mn, 0xC2B2AE35), n_sectors)
    word = sector * np.uint64(8) + umulhi(mul32(fh, 0x9E3779B1), 8)
    p1, p2 = fh & np.uint64(31), (fh >> np.uint64(5)) & np.uint64(31)
    return word * np.uint64(32) + p1, word * np.uint64(32) + p2


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 21
    n_keys = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
    log2_bits = int(sys.argv[3]) if len(sys.argv) > 3 else 29
    m = 16 if k >= 22 else max(k - 6, (k + 1) // 2)
    t0 = time.time()
    rng = np.random.default_rng(1)
    n_sectors = (1 << log2_bits) // 256
    bits = np.zeros(1 << log2_bits, dtype=bool)
    for s0 in range(0, n_keys, 5_000_000):
        keys = rng.integers(0, 1 << (2 * k), size=min(5_000_000, n_keys - s0), dtype=np.uint64)
        i1, i2 = bit_addresses(*coords(keys, k, m), n_sectors)
        bits[i1] = True
        bits[i2] = True
    n_probe = 4_000_000
    seq = rng.integers(0, 4, size=n_probe + k - 1, dtype=np.uint64)
    fwd = np.zeros(n_probe, dtype=np.uint64)
    rc = np.zeros(n_probe, dtype=np.uint64)
    for i in range(k):
        fwd |= seq[i:i + n_probe] << np.uint64(2 * i)
        rc |= (np.uint64(3) - seq[i:i + n_probe]) << np.uint64(2 * (k - 1 - i))
    i1, i2 = bit_addresses(*coords(np.minimum(fwd, rc), k, m), n_sectors)
    fp, fill = float((bits[i1] & bits[i2]).mean()), float(bits.mean())
    sector = i1 >> np.uint64(8)
    print(f"k={k} m={m} keys={n_keys} bitmap=2^{log2_bits}: pass rate of non-members {fp * 100:.3f}%, fill {fill * 100:.2f}%, "
          f"uniform addressing would give {fill * fill * 100:.3f}%, new sector every "
          f"{1 / float((sector[1:] != sector[:-1]).mean()):.2f} windows ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
