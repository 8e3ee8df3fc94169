"""CPU: the two pieces of arithmetic the hand-written surfel depth sort (csrc/gsl_sort.cu) rests on, restated in numpy with
the kernels' own indexing -- TEST INFRASTRUCTURE, not a product path (the product has no CPU fallback).

 * warp_sort_bucket<E>: one warp sorts up to 32 E elements held E per lane (element i = e * 32 + lane, padding = +inf)
   with a bitonic network whose exchanges at partner distance j < 32 are lane shuffles (lane ^ j) and at j >= 32
   register-local (e ^ (j / 32)); the directions fold into ((i & j) == 0) == ((i & k) == 0).  The emulation below runs
   exactly these steps on a (E, 32) array and must produce the sorted sequence for every E and every fill.
 * sort_domain: bucket = (key - kmin) >> shift with the smallest shift that maps [kmin, kmax] into [0, nb); buckets must be
   monotone in the key, so that concatenating the sorted buckets is the sorted sequence (which is what makes the order of
   the reference's stable 64-bit radix sort, rasterizer_impl.cu:338-344, reproducible bucket by bucket)."""
import numpy as np


def warp_sort_bucket(vals, E):
    n, m = len(vals), 32 * E
    v = np.full((E, 32), np.iinfo(np.uint64).max, dtype=np.uint64)
    for i, x in enumerate(vals):
        v[i // 32, i % 32] = x
    lane = np.arange(32)
    k = 2
    while k <= m:
        j = k >> 1
        while j > 0:
            if j >= 32:
                jj = j >> 5
                for e in range(E):
                    if e & jj == 0:
                        a, b = v[e].copy(), v[e | jj].copy()
                        up = ((e * 32) & k) == 0
                        swap = (a > b) == up
                        v[e], v[e | jj] = np.where(swap, b, a), np.where(swap, a, b)
            else:
                for e in range(E):
                    o = v[e][lane ^ j]  # __shfl_xor_sync(v[e], j)
                    i = e * 32 + lane
                    take_min = ((i & j) == 0) == ((i & k) == 0)
                    v[e] = np.where((v[e] < o) == take_min, v[e], o)
            j >>= 1
        k <<= 1
    return v.reshape(-1)[:n]


def test_register_bitonic_network_sorts_every_fill_of_every_width():
    rng = np.random.default_rng(0)
    for E in (1, 2, 4, 8):
        for n in sorted({2, 3, 31, 32, 16 * E + 1, 32 * E - 1, 32 * E}):
            if n > 32 * E:
                continue
            for _ in range(3):
                keys = rng.integers(0, 1 << 20, size=n).astype(np.uint64)          # duplicate depth keys happen
                ids = rng.permutation(1 << 20)[:n].astype(np.uint64)               # ids are unique
                vals = (keys << np.uint64(32)) | ids
                assert np.array_equal(warp_sort_bucket(vals, E), np.sort(vals)), (E, n)


def sort_domain(kmin, kmax, nb):
    span = kmax - kmin
    bits_span = int(span | 1).bit_length()
    bits_nb = nb.bit_length() - 1
    return max(0, bits_span - bits_nb)


def sort_num_buckets(P):
    nb = 256
    while nb < 65536 and nb * 64 < max(P, 1):
        nb <<= 1
    return nb


def test_bucket_domain_is_monotone_and_in_range():
    rng = np.random.default_rng(1)
    for P in (1, 300, 40000, 1000000, 4000000):
        nb = sort_num_buckets(P)
        assert nb & (nb - 1) == 0 and 256 <= nb <= 65536
        for lo, hi in ((0.02, 30.0), (1.0, 1.0), (0.3, 8.0), (1e-3, 1e4)):
            r = np.exp(rng.uniform(np.log(lo), np.log(hi), size=min(P, 20000))).astype(np.float32)
            keys = r.view(np.uint32).astype(np.int64)  # bit patterns of positive floats are monotone in the value
            kmin, kmax = int(keys.min()), int(keys.max())
            shift = sort_domain(kmin, kmax, nb)
            b = (keys - kmin) >> shift
            assert b.min() >= 0 and b.max() < nb
            order = np.argsort(keys, kind="stable")
            assert np.all(np.diff(b[order]) >= 0)
            # a stale, wider range (the key range is reset by the previous sort's last kernel; see launch_depth_keys)
            # only widens the domain: still in range and monotone
            shift2 = sort_domain(max(0, kmin - 12345), kmax + 999, nb)
            b2 = (keys - max(0, kmin - 12345)) >> shift2
            assert b2.min() >= 0 and b2.max() < nb and np.all(np.diff(b2[order]) >= 0)
