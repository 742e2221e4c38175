"""Seed-search kernel timing on one GPU: hash (3) vs bucket (2) vs sweep (1) kernel on a config-3 chunk, with a
candidate-for-candidate comparison between the two."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghostm_b200 import capi, workloads
n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ctx = capi.Context(0)
ctx.set_options(0xF, workloads.blosum62())
ctx.set_candidate_capacity(1 << 26)
seq, starts = workloads.synth_chunk(1, 0, 120 << 20)
ctx.db_build_index(0, seq, starts, 0xF)
q = workloads.synth_queries(2, seq[:4 << 20].copy(), n_q, 75)
ctx.query_upload(q)
ref = None
for variant in (4, 2, 1):
    ctx.set_search_variant(variant)
    best = 1e9
    for rep in range(4):
        st = capi.GmStats()
        counts, total = ctx.search(0, st)
        best = min(best, st.ms_search)
    ids, cand = ctx.candidates(0, n_q, total)
    if ref is None:
        ref = (counts, ids, cand)
        same = True
    else:
        same = bool(np.array_equal(ref[0], counts) and np.array_equal(ref[1], ids) and np.array_equal(ref[2], cand))
    gbs = (st.seed_positions * 4 + n_q * 36 * 12 + total * 4) / (best * 1e-3) / 1e9
    print(json.dumps({"variant": variant, "ms_search": round(best, 3), "cands": total,
                      "positions": int(st.seed_positions), "launches": st.kernel_launches,
                      "algorithmic_GBps": round(gbs, 1), "identical_to_first": same}), flush=True)
