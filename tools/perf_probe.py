"""Stage timing probe on one GPU: synthetic C3-like chunk(s), prints gm_stats per stage."""
import argparse, json, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ghostm_b200 import capi, synth, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--chunk-mib", type=float, default=120)
ap.add_argument("--chunks", type=int, default=1)
ap.add_argument("--queries", type=int, default=16384)
ap.add_argument("--length", type=int, default=75)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--repeat-db", action="store_true")
ap.add_argument("--capacity", type=int, default=1 << 28)
a = ap.parse_args()

ctx = capi.Context(0)
mat = workloads.blosum62()
ctx.set_options(0xF, mat)
ctx.set_candidate_capacity(a.capacity)
t = time.time()
src = None
for c in range(a.chunks):
    seq, starts = workloads.synth_chunk(1, c, int(a.chunk_mib * (1 << 20)), repeats=a.repeat_db)
    if c == 0: src = seq[: 4 << 20].copy()
    ctx.db_build_index(c, seq, starts, 0xF)
print("db gen+index %.1fs" % (time.time() - t), file=sys.stderr)
q = workloads.synth_queries(2, src, a.queries, a.length)
print("dpx peak G lane-instr/s", ctx.measure_dpx_peak() / 1e9, file=sys.stderr)
for rep in range(a.reps):
    ctx.query_upload(q)
    st = capi.GmStats()
    t = time.time()
    for c in range(a.chunks):
        ctx.align_chunk(c, st)
    ctx.traceback_pending(st)
    wall = time.time() - t
    d = st.as_dict(); d["wall_s"] = wall
    d["gcups_sw"] = d["cells"] / (d["ms_score"] * 1e-3) / 1e9
    d["gcups_path"] = d["cells"] / wall / 1e9
    print(json.dumps(d))
