import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import fuzzy_match_b200 as fmb
from fuzzy_match_b200 import synth
tm, off, V = synth.make_tm(60000, vocab=8, len_lo=4, len_hi=24, seed=601)
q, qo = synth.make_queries(tm, off, 12, vocab=8, seed=602, len_lo=6, len_hi=20)
index = fmb.Index(tm, off, V)
index.set_profiling(True)
out, cnt = index.match_batch(q, qo, cap=8, fuzzy=0.2, n=5, ml=1)
print(index.profile())
out, cnt = index.match_batch(q, qo, cap=64, fuzzy=0.3, n=0, ml=2, costs=(1, 0, 1))
print(index.profile())
