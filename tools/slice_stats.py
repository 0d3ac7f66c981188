#!/usr/bin/env python
"""Shape of the gather kernel's work on the bench workload: distribution of range-slice sizes, how the
walked elements spread over them, and how often one suffix-array range is walked by several queries of
a batch. Uses the debug hook fm_debug_last_slices. Usage: python tools/slice_stats.py [--queries N]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fuzzy_match_b200 as fmb  # noqa: E402
from fuzzy_match_b200 import capi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sentences", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--fuzzy", type=float, default=0.7)
    args = ap.parse_args()
    tm, off, V = synth.make_tm(args.sentences, seed=1234)
    q, qo = synth.make_queries(tm, off, args.queries, seed=5678)
    index = fmb.Index(tm, off, V)
    index.set_profiling(True)
    index.match_batch(q, qo, cap=1, fuzzy=args.fuzzy, n=1, ml=3)
    prof = index.profile()
    lib = capi.load_library()
    cap = 4 * int(prof["n_slices"]) + (1 << 21)
    rec = np.zeros((cap, 4), dtype=np.int32)
    start = np.zeros(cap, dtype=np.int64)
    lib.fm_debug_last_slices.restype = C.c_int64
    n = lib.fm_debug_last_slices(index.h, C.c_void_p(rec.ctypes.data), C.c_void_p(start.ctypes.data), C.c_int64(cap))
    rec = rec[:n]
    size = rec[:, 3].astype(np.int64)
    lm = rec[:, 2] & 1023
    total = size.sum()
    print("slices %d elements %d (profile: %d / %d) survivors %d" % (n, total, prof["n_slices"], prof["n_elements"], prof["n_survivors"]))
    edges = [1, 2, 3, 5, 9, 17, 33, 65, 129, 257, 513, 1025, 4097, 16385, 1 << 30]
    print("%-14s %10s %8s %12s %8s" % ("slice size", "slices", "share", "elements", "share"))
    for a, b in zip(edges[:-1], edges[1:]):
        m = (size >= a) & (size < b)
        print("%-14s %10d %7.1f%% %12d %7.1f%%" % ("[%d,%d)" % (a, b), m.sum(), 100.0 * m.sum() / n, size[m].sum(), 100.0 * size[m].sum() / total))
    print("match length of the elements:", {int(k): int(size[lm == k].sum()) for k in np.unique(lm)[:8]})
    # same (begin, size) walked by several queries
    key = rec[:, 1].astype(np.int64) << 32 | size
    u, cnt = np.unique(key, return_counts=True)
    usize = u & 0xffffffff
    print("distinct ranges %d; elements in ranges shared by >= 2 / 4 / 8 / 32 slices: %.1f%% %.1f%% %.1f%% %.1f%%"
          % (len(u), *[100.0 * (usize * cnt)[cnt >= k].sum() / total for k in (2, 4, 8, 32)]))
    print("distinct-range elements (each range once): %d (%.1f%% of walked)" % (usize.sum(), 100.0 * usize.sum() / total))
    # consecutive slices of one query (chains of one query are emitted by neighbouring threads)
    same_q = (rec[1:, 0] == rec[:-1, 0]).mean()
    print("P(next slice has the same query) = %.2f" % same_q)


if __name__ == "__main__":
    main()
