#!/usr/bin/env python
"""Runs the other BASELINE.json configs (3: f=0.5 ml=3 on a larger TM, 4: 200-300 token patterns,
5: contrastive n=10 + idf) through the C ABI: parity against the oracle on a sample of the queries and
device timings. Usage: python tools/config_sweep.py [--sentences N] [--queries Q]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fuzzy_match_b200 as fmb  # noqa: E402
from fuzzy_match_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402


def run(name, index, oracle, q, qo, cap, sample, **params):
    index.set_profiling(True)
    index.match_batch(q, qo, cap=cap, **params)  # warm-up (workspace growth)
    t = time.perf_counter()
    out, cnt = index.match_batch(q, qo, cap=cap, **params)
    dt = time.perf_counter() - t
    prof = index.profile()
    n_q = len(qo) - 1
    t = time.perf_counter()
    ro, oc = oracle.match_batch(q[:qo[sample]], qo[:sample + 1], cap=cap, nthreads=os.cpu_count(), **params)
    cpu = time.perf_counter() - t
    same = (cnt[:sample] == oc).all() and all(out[i, :min(cnt[i], cap)].tobytes() == ro[i].tobytes() for i in range(sample))
    stages = {k[3:]: round(prof[k], 3) for k in prof if k.startswith("ms_")}
    print("%s: %d queries in %.1f ms e2e (%.0f q/s), found %d; oracle sample %d in %.2f s (%.0f q/s, %d threads); identical=%s"
          % (name, n_q, dt * 1e3, n_q / dt, int((cnt > 0).sum()), sample, cpu, sample / cpu, os.cpu_count(), same))
    print("   stages(ms, last chunk):", stages, "elements", prof["n_elements"], "stage2", prof["n_stage2"], "verified", prof["n_verified"], "survivors", prof["n_survivors"], "retries", prof["retries"])
    return same


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sentences", type=int, default=1000000)
    ap.add_argument("--queries", type=int, default=100000)
    ap.add_argument("--only-config4", action="store_true", help="only the long-pattern configuration (for profiling)")
    args = ap.parse_args()
    ok = True
    if not args.only_config4:
        ok &= short_configs(args)
    ok &= long_config(args)
    print("ALL IDENTICAL" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


def short_configs(args):
    ok = True
    tm, off, V = synth.make_tm(args.sentences, seed=1234)
    q, qo = synth.make_queries(tm, off, args.queries, seed=5678)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    ok &= run("config2 f=0.7 n=1 ml=3", index, oracle, q, qo, 1, 3000, fuzzy=0.7, n=1, ml=3)
    ok &= run("config2' CLI defaults f=0.8 n=5 ml=3 mr=0.3", index, oracle, q, qo, 5, 3000, fuzzy=0.8, n=5, ml=3, mr=0.3)
    ok &= run("config3-shape f=0.5 n=1 ml=3", index, oracle, q, qo, 1, 2000, fuzzy=0.5, n=1, ml=3)
    ok &= run("config5 contrastive n=10 c=0.5 idf=1", index, oracle, q, qo, 10, 2000, fuzzy=0.7, n=10, ml=3, idf=1.0, contrast=0.5)
    return ok


def long_config(args):
    # config 4: 980k short + 20k long sentences, queries = perturbed long sentences
    n_long = max(200, args.sentences // 50)
    tm, off, V = synth.make_tm(args.sentences, seed=1234, n_long=n_long)
    src = np.arange(args.sentences - n_long, args.sentences)
    nq4 = max(100, args.queries // 50)
    q, qo = synth.make_queries(tm, off, nq4, seed=5678, source_ids=src, frac_random=0.2, len_lo=200, len_hi=300)
    index, oracle = fmb.Index(tm, off, V), ob.OracleIndex(tm, off, V)
    return run("config4 long patterns f=0.7 n=1 ml=3", index, oracle, q, qo, 1, min(nq4, 300), fuzzy=0.7, n=1, ml=3)


if __name__ == "__main__":
    main()
