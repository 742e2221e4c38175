"""Tile seed-search kernel: tuning sweep (GM_TILE_CFG = warps per CTA,CTAs per SM)
on one config-3 chunk, every setting compared candidate for candidate with the bucket kernel."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghostm_b200 import capi, workloads
n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfgs = sys.argv[2:] or ["16,2", "12,2", "20,2", "10,3", "8,4", "24,1", "32,1"]
ctx = capi.Context(0)
ctx.set_options(0xF, workloads.blosum62())
ctx.set_candidate_capacity(1 << 26)
seq, starts = workloads.synth_chunk(1, 0, 120 << 20)
ctx.db_build_index(0, seq, starts, 0xF)
q = workloads.synth_queries(2, seq[:4 << 20].copy(), n_q, 75)
ctx.query_upload(q)

def run(variant, reps=4):
    ctx.set_search_variant(variant)
    best = 1e9
    for rep in range(reps):
        st = capi.GmStats()
        counts, total = ctx.search(0, st)
        best = min(best, st.ms_search)
    ids, cand = ctx.candidates(0, n_q, total)
    return best, st, counts, total, ids, cand

best, st, rc, rt, ri, rcand = run(2)
print(json.dumps({"variant": 2, "ms_search": round(best, 3), "cands": int(rt), "positions": int(st.seed_positions)}), flush=True)
for cfg in cfgs:
    os.environ["GM_TILE_CFG"] = cfg
    try:
        best, st, counts, total, ids, cand = run(4, reps=5)   # the first repetition also builds the split table
    except Exception as e:
        print(json.dumps({"cfg": cfg, "error": str(e)}), flush=True)
        continue
    same = bool(np.array_equal(rc, counts) and np.array_equal(ri, ids) and np.array_equal(rcand, cand))
    gbs = (st.seed_positions * 4 + n_q * 36 * 12 + total * 4) / (best * 1e-3) / 1e9
    print(json.dumps({"cfg": cfg, "ms_search": round(best, 3), "cands": int(total), "launches": st.kernel_launches,
                      "algorithmic_GBps": round(gbs, 1), "identical": same}), flush=True)
