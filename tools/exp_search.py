"""Ablation timing of the seed-search kernel (GM_SEARCH_DEBUG bits); results are NOT valid."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghostm_b200 import capi, workloads
ctx = capi.Context(0)
ctx.set_options(0xF, workloads.blosum62())
ctx.set_candidate_capacity(1 << 27)
seq, starts = workloads.synth_chunk(1, 0, 120 << 20)
ctx.db_build_index(0, seq, starts, 0xF)
q = workloads.synth_queries(2, seq[:4 << 20].copy(), 8192, 75)
ctx.query_upload(q)
for variant in (1, 0):
    ctx.set_search_variant(bool(variant))
    for d in ([0, 1, 2, 4, 8, 16, 32, 3, 7, 15, 31, 63] if variant else [0]):
        os.environ["GM_SEARCH_DEBUG"] = str(d)
        best = 1e9
        for rep in range(3):
            st = capi.GmStats()
            try:
                counts, total = ctx.search(0, st)
            except Exception as e:
                total = -1
            best = min(best, st.ms_search)
        print(json.dumps({"fast": variant, "debug": d, "ms_search": round(best, 3), "cands": total}), flush=True)
