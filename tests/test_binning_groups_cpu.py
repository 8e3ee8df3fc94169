"""CPU: the binning scheme of csrc/gsl_binning.cu restated in numpy -- TEST INFRASTRUCTURE.  The reference duplicates one
(tile << 32 | depth bits) key per (surfel, tile) instance, emitted y-major / x-minor per surfel in id order, and sorts them
with a stable radix sort (rasterizer_impl.cu:68-111,338-344).  The product instead sorts the SURFELS once by (depth bits,
id) and distributes their instances with a stable counting pass over chunks of 256 depth-ranked surfels, one group of at
most 1024 consecutive tile ids at a time (gsl_bin_groups).  Both must give the same point list and tile ranges -- including
images of more than 1024 tiles (several groups) and tile rows wider than one group."""
import ctypes as C

import numpy as np


def reference_binning(rects, keys, gx):
    """rasterizer_impl.cu:68-111 (duplicateWithKeys) + stable sort + identifyTileRanges (:116-142)."""
    k64, ids = [], []
    for i, (x0, y0, x1, y1) in enumerate(rects):
        for y in range(y0, y1):
            for x in range(x0, x1):
                k64.append(((y * gx + x) << 32) | int(keys[i]))
                ids.append(i)
    k64, ids = np.array(k64, dtype=np.uint64), np.array(ids, dtype=np.int64)
    order = np.argsort(k64, kind="stable")
    return k64[order], ids[order]


def grouped_counting_pass(rects, keys, gx, gy, groups):
    """launch_binning of gsl_binning.cu: k_bin_count / k_bin_scan (+ its last CTA) / k_bin_scatter per tile group."""
    P = len(rects)
    order = np.lexsort((np.arange(P), keys))  # launch_surfel_sort: by (depth bits, id)
    nchunks = (P + 255) // 256
    tiles = gx * gy
    ranges = np.zeros((tiles, 2), dtype=np.int64)
    point_list = []
    R = 0
    for t0, nt, y0, y1, x0, x1 in groups:
        def clipped(rc):  # for_each_chunk_tile: the part of the rect inside the group
            rx, ry = max(rc[0], x0), max(rc[1], y0)
            w, h = min(rc[2], x1) - rx, min(rc[3], y1) - ry
            return [(y * gx + x - t0) for y in range(ry, ry + h) for x in range(rx, rx + w)] if (w > 0 and h > 0) else []
        hist = np.zeros((nt, nchunks), dtype=np.int64)
        for c in range(nchunks):
            for j in order[c * 256:(c + 1) * 256]:
                for lt in clipped(rects[j]):
                    hist[lt, c] += 1
        excl = np.cumsum(hist, axis=1) - hist            # k_bin_scan: exclusive scan over the chunks of every tile
        total = hist.sum(axis=1)
        base = R + np.cumsum(total) - total              # its last CTA: scan of the totals, continued at R
        for lt in range(nt):
            ranges[t0 + lt] = (base[lt], base[lt] + total[lt]) if total[lt] else (0, 0)
        out = np.full(int(total.sum()), -1, dtype=np.int64)
        for c in range(nchunks):                         # k_bin_scatter: rank inside the chunk = earlier surfels of the tile
            seen = np.zeros(nt, dtype=np.int64)
            for j in order[c * 256:(c + 1) * 256]:
                for lt in clipped(rects[j]):
                    out[base[lt] - R + excl[lt, c] + seen[lt]] = j
                    seen[lt] += 1
        assert (out >= 0).all()
        point_list.append(out)
        R += int(total.sum())
    return np.concatenate(point_list) if point_list else np.zeros(0, dtype=np.int64), ranges, R


def _groups(lib, W, H):
    n = lib.gsl_bin_groups(W, H, None, 0)
    buf = (C.c_int32 * (6 * n))()
    assert lib.gsl_bin_groups(W, H, buf, n) == n
    return [tuple(buf[6 * i:6 * i + 6]) for i in range(n)]


def test_grouped_counting_pass_reproduces_the_reference_key_sort():
    from gs_lidar_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(5)
    # (W, H): one group; 1105 tiles = two groups of tile rows; a tile row of 1032 tiles = two pieces per row
    for W, H, P in ((1030, 66, 700), (1040, 272, 900), (16500, 40, 900)):
        gx, gy = (W + 15) // 16, (H + 15) // 16
        groups = _groups(lib, W, H)
        assert (len(groups) > 1) == (gx * gy > 1024)
        rects = []
        for _ in range(P):
            kind = rng.integers(0, 10)
            if kind == 0:      # culled: empty rect
                rects.append((0, 0, 0, 0))
                continue
            w = gx if kind == 1 else int(rng.integers(1, 5))          # kind 1: a seam surfel covering whole tile rows
            h = int(rng.integers(1, 4))
            x0 = 0 if w == gx else int(rng.integers(0, gx - w + 1))
            y0 = int(rng.integers(0, max(gy - h, 0) + 1))
            rects.append((x0, y0, x0 + w, min(y0 + h, gy)))
        keys = rng.integers(1, 200, size=P).astype(np.uint32)          # few distinct depths: ties are broken by the id
        ref_keys, ref_list = reference_binning(rects, keys, gx)
        got_list, ranges, R = grouped_counting_pass(rects, keys, gx, gy, groups)
        assert R == len(ref_list)
        assert np.array_equal(got_list, ref_list)
        ref_tiles = (ref_keys >> np.uint64(32)).astype(np.int64)
        for t in range(gx * gy):
            idx = np.nonzero(ref_tiles == t)[0]
            want = (idx[0], idx[-1] + 1) if len(idx) else (0, 0)
            assert tuple(ranges[t]) == want, t
