#!/usr/bin/env python
"""bench.py - GCUPS / queries per second of the `ghostm aln` hot path on B200.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference ...                    (the reference CPU aligner, host cores)

Workload (BASELINE.json config 3): synthetic metagenomic reads, 75-aa queries against a synthetic
1 G-residue protein db formatted like `ghostm db -l 120` (8 chunks), BLOSUM62, default options.
One STEP = one batch of --queries queries through the whole path (seed search, candidate
chunking, SW extension, Merge, TraceBack) against the WHOLE db.  The db chunks are resident in
HBM and sharded by chunk over the N ranks (strong scaling: total work per step is fixed): every
rank searches and extends its own chunks for ALL queries, the scored candidates are exchanged by
query slice (NCCL all-to-all) and every rank runs Merge + TraceBack for its slice of the queries
over all chunks in ascending order (ghostm_b200/shard.py).

value  = SW cells of the step / device time, inputs (db, index, queries) resident in HBM.
e2e    = same through the C ABI with HOST buffers: queries H2D from pinned memory and hit lists
         D2H inside the timed region, every step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gcups"
UNIT = "GCUPS"
INT_OPS_PER_CELL = 10          # BASELINE.md "Roofline units": aligner.cpp:614-654
DPX_OPS_PER_LANE_INSTR = 4     # VIADDMNMX.S16x2 = 2 halves x (add + max)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=65536, help="queries per step (batch)")
    ap.add_argument("--length", type=int, default=75)
    ap.add_argument("--db-residues", type=float, default=1e9)
    ap.add_argument("--chunk-mib", type=float, default=120.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-mib", type=float, default=32.0)
    ap.add_argument("--cpu-sample-queries", type=int, default=2048)
    return ap.parse_args()


def n_db_chunks(a) -> int:
    # residues + one separator per ~351-residue sequence
    total_bytes = a.db_residues * (1 + 1 / 351.0)
    return max(1, int(np.ceil(total_bytes / (a.chunk_mib * (1 << 20)))))


def chunk_bytes_of(a, c: int, n: int) -> int:
    full = int(a.chunk_mib * (1 << 20))
    if c < n - 1:
        return full
    rest = int(a.db_residues * (1 + 1 / 351.0)) - full * (n - 1)
    return max(min(rest, full), 1 << 16)


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = None
        if shutil.which("nvidia-smi"):
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.out = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.out,
                stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.out.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- our arm

class _DevArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False),
                                         "version": 3, "strides": None}


def run_ours(a):
    import torch
    from ghostm_b200 import capi, ring, shard, workloads

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = capi.Context(local)
    matrix = workloads.blosum62()
    ctx.set_options(0xF, matrix)                      # -k 4, aligner.cpp:227-245 defaults
    ctx.set_candidate_capacity(min(max(a.queries * 1200, 1 << 22), (1 << 32) - 1))
    ctx.set_deferred_traceback(True)                  # TraceBack once, for the survivors only
    sharded = world > 1
    back_ctx = None
    if sharded:    # Merge + TraceBack of this rank's query slice: residues + .pos of every chunk
        back_ctx = capi.Context(local)
        back_ctx.set_options(0xF, matrix)
        back_ctx.set_candidate_capacity(min(max(a.queries * 1200, 1 << 22), (1 << 32) - 1))

    n_chunks = n_db_chunks(a)
    mine = shard.chunks_of_rank(n_chunks, rank, world)
    t_setup = time.time()
    source = None
    sample = None
    for c in range(n_chunks):
        if c not in mine and not sharded and c != 0:
            continue
        seq, starts = workloads.synth_chunk(1, c, chunk_bytes_of(a, c, n_chunks))
        if c in mine:
            ctx.db_build_index(c, seq, starts, 0xF)
        if sharded:
            back_ctx.db_upload_seq(c, seq, starts)
        if c == 0:
            source = seq[: 4 << 20].copy()
            cut = int(np.searchsorted(starts, int(a.cpu_sample_mib * (1 << 20))))
            cut = max(cut, 2)
            sample = (seq[: starts[cut]].copy(), starts[:cut].copy())
        del seq
    # query batches: generated on rank 0 from db chunk 0, broadcast to every rank
    n_batches = 4
    q_all = torch.empty((n_batches, a.queries, a.length), dtype=torch.uint8)
    if rank == 0:
        for b in range(n_batches):
            q_all[b] = torch.from_numpy(workloads.synth_queries(2 + b, source, a.queries, a.length))
    if world > 1:
        qd = q_all.cuda()
        dist.broadcast(qd, 0)
        q_all = qd.cpu()
    q_pinned = q_all.pin_memory()
    cap = 10
    bounds = shard.slice_bounds(None, a.queries, world)
    base, stop = int(bounds[rank]), int(bounds[rank + 1])
    hits_pinned = torch.empty(((stop - base) * cap * 9,), dtype=torch.int32).pin_memory()
    counts_pinned = torch.empty((stop - base,), dtype=torch.int32).pin_memory()
    setup_s = time.time() - t_setup

    dpx_rate = ctx.measure_dpx_peak()
    stats = capi.GmStats()
    stats_back = capi.GmStats()
    front = back = None
    if sharded:
        front = shard.GpuFront(ctx, a.queries, min(max(a.queries * 1200, 1 << 22), (1 << 32) - 1),
                               f"cuda:{local}", stats)
        back = shard.GpuBack(back_ctx, stats_back)

    class GpuEngine(ring.Engine):
        def prepare(self, c):
            ctx.align_prepare(c, stats)

        def merge(self):
            ctx.align_merge(stats)

        def list_tensors(self):
            hp, cp = ctx.results_device()
            dev = f"cuda:{local}"
            return (torch.as_tensor(_DevArray(hp, a.queries * cap * 9), device=dev),
                    torch.as_tensor(_DevArray(cp, a.queries), device=dev))

        def lists_received(self):
            torch.cuda.synchronize()

    engine = GpuEngine()
    stream = torch.cuda.ExternalStream(ctx.stream())
    stream_end = torch.cuda.ExternalStream(back_ctx.stream()) if sharded else stream

    host_t = {}

    def step(s: int, e2e: bool):
        b = s % n_batches
        t_0 = time.perf_counter()
        if not sharded:
            if e2e:
                ctx.query_upload_ptr(q_pinned[b].data_ptr(), a.queries, a.length)
            else:
                ctx.results_clear()
            ring.ring_step(engine, dist, rank, world, n_chunks)
            ctx.traceback_pending(stats)
            if e2e:
                ctx.results_download_ptr(hits_pinned.data_ptr(), counts_pinned.data_ptr())
            return
        if e2e:   # every rank needs all queries (front) and its own slice again (back)
            ctx.query_upload_ptr(q_pinned[b].data_ptr(), a.queries, a.length)
            back_ctx.query_upload_ptr(q_pinned[b][base:stop].data_ptr(), stop - base, a.length)
        else:
            back_ctx.results_clear()
        host_t["upload"] = host_t.get("upload", 0.0) + time.perf_counter() - t_0
        shard.shard_step(front, back, dist, rank, world, n_chunks, bounds,
                         before_back=torch.cuda.synchronize, timers=host_t)
        t_1 = time.perf_counter()
        if e2e:
            back_ctx.results_download_ptr(hits_pinned.data_ptr(), counts_pinned.data_ptr())
        host_t["download"] = host_t.get("download", 0.0) + time.perf_counter() - t_1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(e2e: bool):
        if not e2e:
            ctx.query_upload_ptr(q_pinned[0].data_ptr(), a.queries, a.length)
            if sharded:
                back_ctx.query_upload_ptr(q_pinned[0][base:stop].data_ptr(), stop - base, a.length)
        for s in range(a.warmup):
            step(s, e2e)
        barrier()
        for st_ in (stats, stats_back):      # in place: the engines hold references
            C.memset(C.byref(st_), 0, C.sizeof(st_))
        if sharded:
            front.launches = back.launches = 0
        host_t.clear()
        sampler = ClockSampler(local) if rank == 0 else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for s in range(a.steps):
            step(a.warmup + s, e2e)
        ev1.record(stream_end)
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=f"cuda:{local}")
        launches = stats.kernel_launches + stats_back.kernel_launches
        if sharded:
            launches += front.launches + back.launches
        cells = torch.tensor([float(stats.cells), float(stats.candidates), float(launches),
                              float(stats.seed_positions)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cells, op=dist.ReduceOp.SUM)
        st = stats.as_dict()
        for k in ("ms_merge", "ms_traceback", "tracebacks", "candidate_chunks"):
            st[k] += getattr(stats_back, k)
        st["host_ms_per_step"] = {k: round(v * 1e3 / a.steps, 3) for k, v in host_t.items()}
        return t.tolist(), cells.tolist(), st, clocks

    (dev_ms, wall_ms), (cells, cands, launches, positions), st, clocks = timed(False)
    (e_dev_ms, e_wall_ms), (e_cells, _, _, _), e_st, _ = timed(True)

    out = None
    if rank == 0:
        ms_per_step = dev_ms / a.steps
        value = cells / (dev_ms * 1e-3) / 1e9
        e2e_value = e_cells / (e_wall_ms * 1e-3) / 1e9
        sw_s = st["ms_score"] * 1e-3
        achieved = st["cells"] * INT_OPS_PER_CELL / sw_s / 1e12 if sw_s > 0 else 0.0
        peak = dpx_rate * DPX_OPS_PER_LANE_INSTR / 1e12
        search_s = st["ms_search"] * 1e-3
        search_bytes = (st["seed_positions"] * 4 + a.queries * a.steps * len(mine) * 36 * 12
                        + st["candidates"] * 4)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "s16", "data": "synthetic",
            "queries_per_s": a.queries / (ms_per_step * 1e-3),
            "config": {
                "workload": "config3: synthetic 75-aa reads vs synthetic 1 G-residue protein db, "
                            "BLOSUM62, ghostm aln defaults",
                "queries_per_step": a.queries, "query_len": a.length,
                "db_residues": a.db_residues, "db_chunks": n_chunks, "chunk_mib": a.chunk_mib,
                "parallelism": (f"db chunks (index) sharded over {world} rank(s) for search + SW; "
                                "candidates all-to-all by query slice over NCCL; Merge + TraceBack "
                                "per query slice" if sharded else "1 rank, all db chunks resident"),
                "cache": "inputs larger than L2: every step streams the index positions and "
                         "windows of all db chunks (>=0.6 GB per chunk) from HBM",
                "traceback": "deferred to survivors",
            },
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e_wall_ms / a.steps,
                    "queries_per_s": a.queries / (e_wall_ms / a.steps * 1e-3),
                    "h2d_bytes_per_step": int(a.queries * a.length) * (world + (1 if sharded else 0)),
                    "d2h_bytes_per_step": int(a.queries * cap * 36 + a.queries * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "int_dpx", "kernel": "sw_extend_dpx_kernel<75>",
                         "achieved": achieved, "peak": peak, "unit": "Tint-op/s",
                         "frac": achieved / peak if peak else None,
                         # dram__bytes_read+write of one `ncu --set full` capture of this kernel
                         # (profiles/r1_sw_extend_ncu.md: 611.0 MB for 5,595,674 candidates),
                         # scaled to the candidates of one launch of this run
                         "traffic": 611.0e6 / 5595674 * (st["candidates"] / max(a.steps * len(mine), 1)),
                         "algorithmic_bytes": (a.length + 36 + 12) * (st["candidates"] / max(a.steps * len(mine), 1)),
                         "note": "achieved = SW cells x 10 int ops / SW kernel time (CUDA events, "
                                 "rank 0); peak = measured VIADDMNMX.S16x2 issue rate x 4 ops "
                                 "(gm_measure_dpx_peak, same process)",
                         "sw_gcups": st["cells"] / sw_s / 1e9 if sw_s > 0 else None},
            "roofline_seed_search": {"bound": "hbm", "achieved": search_bytes / search_s / 1e9
                                     if search_s > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": search_bytes / search_s / 1e9 / hbm_peak
                                     if search_s > 0 else None,
                                     "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "stage_ms_per_step_rank0": {k: st[k] / a.steps for k in
                                        ("ms_search", "ms_score", "ms_merge", "ms_traceback")},
            "host_ms_per_step_rank0": st.get("host_ms_per_step"),
            "wall_ms_per_step": wall_ms / a.steps,
            "e2e_host_ms_per_step_rank0": e_st.get("host_ms_per_step"),
            "setup_s": setup_s,
        }
        if world == 1 and not a.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(a, ctx, sample, q_all[0].numpy(), matrix)
            except Exception as e:  # the baseline is reported, never allowed to hide the GPU number
                out["cpu_baseline"] = {"error": repr(e)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


# ----------------------------------------------------------------------------- CPU baseline

def _write_sample(tmp, seq, starts, queries, n_qchunks):
    from ghostm_b200 import formats
    kc, pos = formats.build_index(seq, starts, 0xF)
    names = [f"s{i}" for i in range(starts.shape[0])]
    chunk = formats.DbChunk(seq, starts.astype(np.uint32), names, 0xF, kc, pos)
    db = formats.Db(seed=0xF, max_chunk_len=1 << 27, sum_residues=int(seq.shape[0] - starts.shape[0]),
                    chunks=[chunk])
    formats.write_db(os.path.join(tmp, "db"), db)
    per = (queries.shape[0] + n_qchunks - 1) // n_qchunks
    qcs = []
    for i in range(n_qchunks):
        part = queries[i * per:(i + 1) * per]
        if part.shape[0]:
            qcs.append(formats.QueryChunk(np.ascontiguousarray(part),
                                          [f"q{i * per + j}" for j in range(part.shape[0])]))
    formats.write_queries(os.path.join(tmp, "q"), qcs)
    return db, qcs


def _ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "ghostm")
    return p if os.path.exists(p) else None


def cpu_baseline(a, ctx, sample, queries, matrix):
    """The reference CPU aligner (oracle/_ref/ghostm, else the oracle port) on a bounded sample of
    the same workload, one host thread, with a parity check of its output against the GPU path."""
    from ghostm_b200 import capi
    from oracle import oracle as O
    seq, starts = sample
    qs = np.ascontiguousarray(queries[: a.cpu_sample_queries])
    tmp = tempfile.mkdtemp(prefix="gm_cpu_")
    try:
        db, qcs = _write_sample(tmp, seq, starts, qs, 1)
        # GPU on the sample: cells and hit lists
        g = capi.Context(int(os.environ.get("LOCAL_RANK", 0)))
        g.set_options(0xF, matrix)
        g.db_upload(0, db.chunks[0])
        g.query_upload(qs)
        st = capi.GmStats()
        g.align_chunk(0, st)
        hits, counts = g.results()
        g.close()
        res = O.ResultLists(qs.shape[0], 10)
        res.hits[:] = hits
        res.counts[:] = counts
        gpu_text = O.format_output(res, qcs[0], db, O.Options())
        ref = _ref_binary()
        t0 = time.perf_counter()
        if ref:
            subprocess.check_call([ref, "aln", "-i", os.path.join(tmp, "q"), "-d",
                                   os.path.join(tmp, "db"), "-o", os.path.join(tmp, "out.txt")],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            wall = time.perf_counter() - t0
            cpu_text = open(os.path.join(tmp, "out.txt"), encoding="latin-1").read()
            kind = "reference"
        else:
            r = O.align_chunk(qcs[0], db, O.Options())
            wall = time.perf_counter() - t0
            cpu_text = O.format_output(r, qcs[0], db, O.Options())
            kind = "port"
        return {"value": st.cells / wall / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                "sample": f"{qs.shape[0]} queries x {a.length} aa vs the first "
                          f"{seq.shape[0] / (1 << 20):.1f} MiB of db chunk 0, whole `ghostm aln` run "
                          f"(search+SW+merge+traceback+output), {wall:.1f} s",
                "queries_per_s": qs.shape[0] / wall,
                "hit_list_identical_to_gpu": cpu_text == gpu_text,
                "hit_rows": cpu_text.count("\n")}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ----------------------------------------------------------------------------- reference arm

def run_reference(a):
    """--impl reference: the reference's own CPU aligner on the box's host cores, one process per
    query chunk (-S i -L i, aligner.cpp:282-288), each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from ghostm_b200 import workloads
    cores = os.cpu_count() or 1
    ref = _ref_binary()
    probe = os.path.join(ROOT, "oracle", "_ref", "ref_probe")
    tmp = tempfile.mkdtemp(prefix="gm_ref_")
    try:
        seq, starts = workloads.synth_chunk(1, 0, 16 << 20)
        per_proc = 128
        qs = workloads.synth_queries(2, seq[: 4 << 20].copy(), per_proc * cores, a.length)
        db, qcs = _write_sample(tmp, seq, starts, qs, cores)
        # cells of the sample (untimed): candidates from the reference stage probe or the oracle
        cells = 0
        base_len = a.length + 2 * 2 + 2 * 16
        if ref and os.path.exists(probe):
            from oracle import oracle as O
            procs = []
            for i in range(len(qcs)):
                procs.append(subprocess.Popen([probe, os.path.join(tmp, f"dump{i}.bin"), "-i",
                                               os.path.join(tmp, "q"), "-d", os.path.join(tmp, "db"),
                                               "-S", str(i), "-L", str(i)],
                                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            for p in procs:
                p.wait()
            for i in range(len(qcs)):
                stages, _ = O.read_probe_dump(os.path.join(tmp, f"dump{i}.bin"))
                for s in stages:
                    st = s[4].astype(np.int64)
                    off = np.maximum(st - 2, 0)
                    w = np.minimum(base_len, seq.shape[0] - off)
                    cells += int(w.sum()) * a.length
            kind = "reference"
        else:
            from oracle import oracle as O
            kind = "port"
            opt = O.Options()
            for qc in qcs:
                for ids, st in O.search_chunks(qc.seqs, db.chunks[0], opt):
                    off = np.maximum(st.astype(np.int64) - 2, 0)
                    cells += int(np.minimum(base_len, seq.shape[0] - off).sum()) * a.length

        def one_step():
            if kind == "reference":
                procs = [subprocess.Popen([ref, "aln", "-i", os.path.join(tmp, "q"), "-d",
                                           os.path.join(tmp, "db"), "-o",
                                           os.path.join(tmp, f"out{i}.txt"), "-S", str(i), "-L", str(i)],
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                         for i in range(len(qcs))]
                for p in procs:
                    if p.wait() != 0:
                        raise RuntimeError("reference aligner failed")
            else:
                import multiprocessing as mp
                with mp.Pool(len(qcs)) as pool:
                    pool.map(_port_worker, [(os.path.join(tmp, "q"), os.path.join(tmp, "db"), i)
                                            for i in range(len(qcs))])

        for _ in range(min(a.warmup, 1)):
            one_step()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            one_step()
        wall = time.perf_counter() - t0
        value = cells * a.steps / wall / 1e9
        sample = (f"{qs.shape[0]} queries x {a.length} aa ({len(qcs)} query chunks, one process each) "
                  f"vs a {seq.shape[0] / (1 << 20):.0f} MiB chunk of the config-3 db, whole "
                  f"`ghostm aln` per step")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": wall / a.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s32",
            "data": "synthetic", "queries_per_s": qs.shape[0] * a.steps / wall,
            "config": {"workload": "config3 (bounded sample): " + sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(qcs), "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def _port_worker(args):
    from ghostm_b200 import formats
    from oracle import oracle as O
    qprefix, dbprefix, i = args
    db = formats.read_db(dbprefix)
    qc = formats.read_queries(qprefix)[i]
    O.align_chunk(qc, db, O.Options())


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
