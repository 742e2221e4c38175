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
    ap.add_argument("--ref-sample-mib", type=float, default=32.0, help="--impl reference: db sample per step")
    ap.add_argument("--ref-queries-per-core", type=int, default=256)
    ap.add_argument("--no-extra", action="store_true", help="skip the config 4 / config 5 blocks")
    ap.add_argument("--extra-steps", type=int, default=2)
    ap.add_argument("--extra-max-gpus", type=int, default=2)
    ap.add_argument("--c4-queries", type=int, default=64, help="config 4: queries per step (sample of the 10 k)")
    ap.add_argument("--c5-queries", type=int, default=10000)
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
              "clocks_event_reasons.sw_power_cap,utilization.gpu")

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
        sm, mx, busy, reasons = [], [], [], set()
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
            try:
                busy.append(float(f[7]))
            except (ValueError, IndexError):
                pass
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm),
                "reasons": sorted(reasons),
                "gpu_busy_pct": round(statistics.mean(busy), 1) if busy else None}


# ----------------------------------------------------------------------------- our arm

class Spec:
    """One workload: db chunk sizes, query batches and the aligner options that differ from the
    `ghostm aln` defaults (aligner.cpp:227-245)."""

    def __init__(self, name, workload, queries, length, chunk_bytes, db_seed, q_seed, repeats=False,
                 varlen=None, frac_db=0.5, max_list_length=1 << 27, capacity=None, n_batches=4,
                 cpu_sample_mib=32.0, cpu_sample_queries=2048, ref_binary="ghostm"):
        self.name, self.workload, self.queries, self.length = name, workload, queries, length
        self.chunk_bytes, self.db_seed, self.q_seed, self.repeats = chunk_bytes, db_seed, q_seed, repeats
        self.varlen, self.frac_db, self.max_list_length = varlen, frac_db, max_list_length
        self.capacity = capacity or min(max(queries * 1200, 1 << 22), (1 << 32) - 1)
        self.n_batches, self.ref_binary = n_batches, ref_binary
        self.cpu_sample_mib, self.cpu_sample_queries = cpu_sample_mib, cpu_sample_queries
        self.list_len = (length - 4) // 2 + 1        # (L - k) / shift + 1, aligner.cpp:399

    def make_queries(self, batch, source):
        from ghostm_b200 import workloads
        q = workloads.synth_queries(self.q_seed + batch, source, self.queries, self.length,
                                    frac_db=self.frac_db)
        if self.varlen:      # `qry -l L`: shorter queries are right-padded with X = 23 (query_creator.cpp:410-412)
            lo, hi = self.varlen
            lens = np.random.default_rng([self.q_seed, batch, 99]).integers(lo, hi + 1, size=self.queries)
            q[np.arange(self.length)[None, :] >= lens[:, None]] = 23
        return q


def _new_context(capi, local, matrix, spec):
    ctx = capi.Context(local)
    ctx.set_options(0xF, matrix, max_list_length=spec.max_list_length)   # -k 4, otherwise defaults
    ctx.set_candidate_capacity(spec.capacity)
    ctx.set_deferred_traceback(True)                  # TraceBack once, for the survivors only
    return ctx


class Arm:
    """This rank's share of one workload: front context (index of the chunks c % world == rank),
    back context (residues + .pos of every chunk; Merge + TraceBack of the rank's query slice),
    pinned host buffers, and the pipeline (ghostm_b200.shard.GpuPipeline) - the same code path at
    every N."""

    def __init__(self, spec, torch, dist, rank, world, local, matrix, chunks=None, queries=None):
        from ghostm_b200 import capi, shard, workloads
        self.spec, self.torch, self.dist, self.rank, self.world, self.local = spec, torch, dist, rank, world, local
        self.front_ctx = _new_context(capi, local, matrix, spec)
        self.back_ctx = _new_context(capi, local, matrix, spec)
        self.n_chunks = len(chunks) if chunks is not None else len(spec.chunk_bytes)
        self.mine = shard.chunks_of_rank(self.n_chunks, rank, world)
        self.sample = None
        source = None
        t0 = time.time()
        for c in range(self.n_chunks):
            if chunks is not None:
                seq, starts = chunks[c]
            else:
                seq, starts = workloads.synth_chunk(spec.db_seed, c, spec.chunk_bytes[c], repeats=spec.repeats)
            if c in self.mine:
                self.front_ctx.db_build_index(c, seq, starts, 0xF)
            self.back_ctx.db_upload_seq(c, seq, starts)
            if c == 0:
                source = seq[: 4 << 20].copy()
                cut = max(int(np.searchsorted(starts, int(spec.cpu_sample_mib * (1 << 20)))), 2)
                cut = min(cut, starts.shape[0] - 1)
                self.sample = (seq[: starts[cut]].copy(), starts[:cut].copy())
            del seq
        if queries is not None:
            q_all = torch.from_numpy(np.ascontiguousarray(queries))[None]
        else:   # generated on rank 0 from db chunk 0, broadcast to every rank
            q_all = torch.empty((spec.n_batches, spec.queries, spec.length), dtype=torch.uint8)
            if rank == 0:
                for b in range(spec.n_batches):
                    q_all[b] = torch.from_numpy(spec.make_queries(b, source))
            if world > 1:
                qd = q_all.cuda()
                dist.broadcast(qd, 0)
                q_all = qd.cpu()
        self.q_all = q_all
        self.q_pinned = q_all.pin_memory()
        self.cap = 10
        self.bounds = shard.slice_bounds(None, spec.queries, world)
        self.base, self.stop = int(self.bounds[rank]), int(self.bounds[rank + 1])
        n_slice = max(self.stop - self.base, 1)
        self.hits_pinned = torch.empty((n_slice * self.cap * 9,), dtype=torch.int32).pin_memory()
        self.counts_pinned = torch.empty((n_slice,), dtype=torch.int32).pin_memory()
        self.stats, self.stats_back = capi.GmStats(), capi.GmStats()
        self.pipe = shard.GpuPipeline(self.front_ctx, self.back_ctx, spec.queries, spec.length, spec.capacity,
                                      f"cuda:{local}", dist, rank, world, self.n_chunks, self.bounds,
                                      self.stats, self.stats_back, threaded=True)
        self.front_stream = torch.cuda.ExternalStream(self.front_ctx.stream())
        self.back_stream = torch.cuda.ExternalStream(self.back_ctx.stream())
        self.setup_s = time.time() - t0

    def submit(self, batch: int, e2e: bool):
        if not e2e:
            self.pipe.submit()
            return
        q = self.q_pinned[batch % self.q_pinned.shape[0]]
        self.pipe.submit(queries_ptr=q.data_ptr(), slice_ptr=q[self.base:self.stop].data_ptr()
                         if self.stop > self.base else 0, hits_ptr=self.hits_pinned.data_ptr(),
                         counts_ptr=self.counts_pinned.data_ptr())

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps: int, warmup: int, e2e: bool):
        """-> ([device ms, wall ms] max over ranks, [cells, candidates, launches, positions] summed over
        ranks, rank-local stage stats, clocks (rank 0))."""
        torch = self.torch
        if not e2e:     # resident inputs: batch 0 uploaded once, outside the timed region
            q = self.q_pinned[0]
            self.front_ctx.query_upload_ptr(q.data_ptr(), self.spec.queries, self.spec.length)
            if self.stop > self.base:
                self.back_ctx.query_upload_ptr(q[self.base:self.stop].data_ptr(), self.stop - self.base,
                                               self.spec.length)
        for s in range(warmup):
            self.submit(s, e2e)
        self.pipe.drain()
        self.barrier()
        for st_ in (self.stats, self.stats_back):      # in place: the engines hold references
            C.memset(C.byref(st_), 0, C.sizeof(st_))
        self.pipe.front.launches = self.pipe.back.launches = 0
        self.pipe.timers.clear()
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(self.front_stream)
        for s in range(steps):
            self.submit(warmup + s, e2e)
        self.pipe.drain()
        ev1.record(self.back_stream)
        self.barrier()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if sampler else None
        dev = f"cuda:{self.local}"
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        launches = (self.stats.kernel_launches + self.stats_back.kernel_launches
                    + self.pipe.front.launches + self.pipe.back.launches)
        tot = torch.tensor([float(self.stats.cells), float(self.stats.candidates), float(launches),
                            float(self.stats.seed_positions)], dtype=torch.float64, device=dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(tot, op=self.dist.ReduceOp.SUM)
        st = self.stats.as_dict()
        for k in ("ms_merge", "ms_traceback", "tracebacks"):
            st[k] += getattr(self.stats_back, k)
        st["host_ms_per_step"] = {k: round(v * 1e3 / steps, 3) for k, v in self.pipe.timers.items()}
        return t.tolist(), tot.tolist(), st, clocks

    def checksum(self, batch: int = 0):
        """One untimed batch through the e2e path; 64-bit checksum of the downloaded hit lists, summed
        over the ranks' slices (identical for every N: the proof that a sharded run returns the
        single-GPU lists)."""
        from ghostm_b200 import shard
        self.submit(batch, True)
        self.pipe.drain()
        n_slice = self.stop - self.base
        cs, rows = 0, 0
        if n_slice:
            hits = self.hits_pinned.numpy().view(np.uint32)[: n_slice * self.cap * 9].reshape(n_slice, self.cap, 9)
            counts = self.counts_pinned.numpy().view(np.uint32)[:n_slice]
            cs, rows = shard.hit_checksum(hits, counts, self.base), int(counts.sum())
        if self.world > 1:
            t = self.torch.tensor([cs >> 32, cs & 0xFFFFFFFF, rows], dtype=self.torch.int64, device=f"cuda:{self.local}")
            g = [self.torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(g, t)
            cs = sum((int(x[0]) << 32) | int(x[1]) for x in g) & 0xFFFFFFFFFFFFFFFF
            rows = sum(int(x[2]) for x in g)
        return f"{cs:016x}", rows

    def close(self):
        self.pipe.close()
        self.front_ctx.close()
        self.back_ctx.close()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def _measured_traffic():
    """dram__bytes per launch of the dominant kernels from the committed ncu summary of this round
    (profiles/traffic.json, written from `ncu --set full` captures of this bench's launches)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def rooflines(spec, st, n_launch_front, dpx_rate):
    """Both roofline objects from the rank-0 stage stats of a timed region."""
    peaks = _peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sw_s, search_s = st["ms_score"] * 1e-3, st["ms_search"] * 1e-3
    achieved = st["cells"] * INT_OPS_PER_CELL / sw_s / 1e12 if sw_s > 0 else 0.0
    peak = dpx_rate * DPX_OPS_PER_LANE_INSTR / 1e12
    # 8 B of keys_count + 4 B of key per query k-mer, 4 B per index position, 4 B per candidate
    search_bytes = st["seed_positions"] * 4 + spec.queries * n_launch_front * spec.list_len * 12 + st["candidates"] * 4
    cand_per_launch = st["candidates"] / max(n_launch_front, 1)     # per db chunk pass (one SW launch per candidate chunk)
    traffic = _measured_traffic()
    sw_t, se_t = traffic.get(spec.name + ":sw_extend"), traffic.get(spec.name + ":seed_search")
    window = spec.length + 2 * 2 + 2 * 16
    return ({"bound": "int_dpx", "kernel": "sw_extend_pair_kernel (41 <= L), sw_extend_dpx_kernel (L <= 40)", "achieved": achieved, "peak": peak,
             "unit": "Tint-op/s", "frac": achieved / peak if peak else None,
             "traffic": sw_t["dram_bytes_per_candidate"] * cand_per_launch if sw_t else None,
             "traffic_source": sw_t["source"] if sw_t else None,
             "algorithmic_bytes": (window + 12) * cand_per_launch,
             "note": "achieved = SW cells x 10 int ops / SW kernel time (CUDA events on the kernel's stream, "
                     "rank 0); peak = measured VIADDMNMX.S16x2 issue rate x 4 ops (gm_measure_dpx_peak, "
                     "same process)",
             "sw_gcups": st["cells"] / sw_s / 1e9 if sw_s > 0 else None},
            {"bound": "hbm", "kernel": "seed_search_tile_kernel",
             "achieved": search_bytes / search_s / 1e9 if search_s > 0 else None, "peak": hbm_peak,
             "unit": "GB/s", "frac": search_bytes / search_s / 1e9 / hbm_peak if search_s > 0 else None,
             "traffic": se_t["dram_bytes_per_position"] * st["seed_positions"] / max(n_launch_front, 1) if se_t else None,
             "traffic_source": se_t["source"] if se_t else None,
             "algorithmic_bytes": search_bytes / max(n_launch_front, 1),
             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"})


def config_dict(spec, world):
    """The workload description both arms print (the reference arm times a bounded sample of it)."""
    sharded = world > 1
    return {
        "workload": spec.workload, "queries_per_step": spec.queries, "query_len": spec.length,
        "db_chunks": len(spec.chunk_bytes), "db_bytes": int(sum(spec.chunk_bytes)),
        "max_list_length": spec.max_list_length,
        "parallelism": (f"db chunks (index) sharded over {world} rank(s) for search + SW; candidates "
                        "all-to-all by query slice over NCCL; Merge + TraceBack per query slice"
                        if sharded else "1 rank: front context (search + SW) and back context (Merge + "
                        "TraceBack) on one GPU, all db chunks resident"),
        "pipeline": "back stage of batch s overlaps the front stage of batch s+1 (own stream, own host thread)",
        "cache": "inputs larger than L2: every step streams the index positions and windows of all db "
                 "chunks from HBM",
        "traceback": "deferred to survivors",
    }


def result_block(spec, arm, steps, warmup, dpx_rate, with_clocks=True):
    """Device-resident and e2e timing of one workload + checksum -> dict (valid on rank 0)."""
    (dev_ms, wall_ms), (cells, cands, launches, positions), st, clocks = arm.timed(steps, warmup, False)
    (e_dev_ms, e_wall_ms), (e_cells, _, _, _), e_st, e_clocks = arm.timed(steps, warmup, True)
    checksum, rows = arm.checksum(0)
    world, sharded = arm.world, arm.world > 1
    ms_per_step = dev_ms / steps
    roof_sw, roof_search = rooflines(spec, st, steps * len(arm.mine), dpx_rate)
    out = {
        "value": cells / (dev_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_per_step,
        "queries_per_s": spec.queries / (ms_per_step * 1e-3),
        "config": config_dict(spec, world),
        "e2e": {"value": e_cells / (e_wall_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e_wall_ms / steps,
                "device_ms_per_step": e_dev_ms / steps,
                "queries_per_s": spec.queries / (e_wall_ms / steps * 1e-3),
                "h2d_bytes_per_step": int(spec.queries * spec.length) * world + int(spec.queries * spec.length),
                "d2h_bytes_per_step": int(spec.queries * arm.cap * 36 + spec.queries * 4),
                "gpu_busy_pct": e_clocks.get("gpu_busy_pct") if e_clocks else None},
        "gpu_launches": int(launches),
        "hit_list_checksum": {"value": checksum, "rows": rows,
                              "what": "64-bit checksum of the hit lists of batch 0 downloaded through the e2e "
                                      "path, summed over the ranks' query slices: equal at every N"},
        "roofline": roof_sw, "roofline_seed_search": roof_search,
        "stage_ms_per_step_rank0": {k: st[k] / steps for k in ("ms_search", "ms_score", "ms_merge", "ms_traceback")},
        "host_ms_per_step_rank0": st.get("host_ms_per_step"),
        "wall_ms_per_step": wall_ms / steps,
        "e2e_host_ms_per_step_rank0": e_st.get("host_ms_per_step"),
        "e2e_stage_ms_per_step_rank0": {k: e_st[k] / steps for k in ("ms_search", "ms_score", "ms_merge", "ms_traceback")},
        "candidates_per_step": cands / steps, "cells_per_step": cells / steps,
        "seed_positions_per_step": positions / steps,
        "setup_s": arm.setup_s,
    }
    if with_clocks:
        out["clocks"] = clocks
    return out


def specs(a):
    n = n_db_chunks(a)
    c3 = Spec("config3", "config3: synthetic 75-aa reads vs synthetic 1 G-residue protein db, BLOSUM62, "
              "ghostm aln defaults", a.queries, a.length, [chunk_bytes_of(a, c, n) for c in range(n)], 1, 2,
              cpu_sample_mib=a.cpu_sample_mib, cpu_sample_queries=a.cpu_sample_queries)
    c4 = Spec("config4", "config4 (bounded sample of the 10 k queries): queries of U[300,1000] aa, 15 %-mutated "
              "db substrings, X-padded to 1000 (`qry -l 1000`) vs synthetic 256 M-residue db (2 chunks); "
              "reference semantics with the column cap raised (common.h:38 -> 1024)", a.c4_queries, 1000,
              [128 << 20, 128 << 20], 3, 4, varlen=(300, 1000), frac_db=1.0, capacity=1 << 25, n_batches=2,
              cpu_sample_mib=4.0, cpu_sample_queries=2, ref_binary="ghostm_l1024")
    c5 = Spec("config5", "config5: 10 k x 75 aa queries with tandem repeats vs 64 M-residue db with 30 % "
              "tandem repeats (period 1-4 over {A,G,S,L}), ghostm aln defaults", a.c5_queries, 75, [64 << 20], 5, 6,
              repeats=True, capacity=1 << 29, n_batches=2, cpu_sample_mib=2.0, cpu_sample_queries=128)
    c5l = Spec("config5_l16", c5.workload + ", -l 16 (candidate chunks of 16 Mi)", a.c5_queries, 75, [64 << 20], 5, 6,
               repeats=True, max_list_length=16 << 20, capacity=1 << 29, n_batches=2, cpu_sample_mib=2.0,
               cpu_sample_queries=128)
    return c3, c4, c5, c5l


def run_ours(a):
    import torch
    from ghostm_b200 import capi, workloads

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
    matrix = workloads.blosum62()
    c3, c4, c5, c5l = specs(a)

    arm = Arm(c3, torch, dist, rank, world, local, matrix)
    dpx_rate = arm.front_ctx.measure_dpx_peak()
    block = result_block(c3, arm, a.steps, a.warmup, dpx_rate)
    out = None
    if rank == 0:
        out = {"metric": METRIC, "value": block.pop("value"), "unit": block.pop("unit"), "n_gpus": world,
               "steps": a.steps, "warmup": a.warmup, "ms_per_step": block.pop("ms_per_step"),
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s16",
               "data": "synthetic"}
        out.update(block)
        if world == 1 and not a.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(c3, arm.sample, arm.q_all[0].numpy(), matrix, torch, local)
            except Exception as e:  # the baseline is reported, never allowed to hide the GPU number
                out["cpu_baseline"] = {"error": repr(e)}
    arm.close()
    del arm

    extra = {}
    # config 4 has two db chunks and config 5 one: beyond two ranks there is nothing left to shard
    if not a.no_extra and world > a.extra_max_gpus:
        extra = {"skipped": f"config 4 / config 5 blocks run at N <= {a.extra_max_gpus} (2 and 1 db chunks)"}
    elif not a.no_extra:
        for spec in (c4, c5, c5l):
            try:
                arm = Arm(spec, torch, dist, rank, world, local, matrix)
                # warm-up = one pass over every query batch: device buffers that grow with a batch's
                # candidate count (Merge scratch) reach their steady size before the timed region
                blk = result_block(spec, arm, a.extra_steps, spec.n_batches, dpx_rate, with_clocks=False)
                blk["steps"], blk["warmup"] = a.extra_steps, spec.n_batches
                if rank == 0 and world == 1 and not a.no_cpu_baseline and spec is not c5l:
                    try:
                        blk["cpu_baseline"] = cpu_baseline(spec, arm.sample, arm.q_all[0].numpy(), matrix, torch, local)
                    except Exception as e:
                        blk["cpu_baseline"] = {"error": repr(e)}
                arm.close()
                del arm
                extra[spec.name] = blk
            except Exception as e:
                if world > 1:
                    raise          # a rank that skips a collective would hang the others
                extra[spec.name] = {"error": repr(e)}
    if out is not None:
        out["extra"] = extra
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


def _ref_binary(name="ghostm"):
    p = os.path.join(ROOT, "oracle", "_ref", name)
    return p if os.path.exists(p) else None


def cpu_baseline(spec, sample, queries, matrix, torch, local):
    """The reference CPU aligner (oracle/_ref/<binary>, else the oracle port) on a bounded sample of
    the same workload, one host thread, with a byte comparison of its output against the GPU path
    (the same front / back pipeline as the timed run)."""
    from oracle import oracle as O
    seq, starts = sample
    qs = np.ascontiguousarray(queries[: spec.cpu_sample_queries])
    tmp = tempfile.mkdtemp(prefix="gm_cpu_")
    try:
        db, qcs = _write_sample(tmp, seq, starts, qs, 1)
        # GPU on the sample: cells and hit lists
        sub = Spec(spec.name + "_sample", "", qs.shape[0], spec.length, [seq.shape[0]], 0, 0,
                   max_list_length=spec.max_list_length, capacity=1 << 26)
        arm = Arm(sub, torch, None, 0, 1, local, matrix, chunks=[(seq, starts)], queries=qs)
        arm.submit(0, True)
        arm.pipe.drain()
        cells = int(arm.stats.cells)
        hits = arm.hits_pinned.numpy().view(np.uint32).reshape(qs.shape[0], arm.cap, 9).copy()
        counts = arm.counts_pinned.numpy().view(np.uint32).copy()
        arm.close()
        opt = O.Options(max_list_length=spec.max_list_length)
        res = O.ResultLists(qs.shape[0], 10)
        res.hits[:] = hits.view(res.hits.dtype).reshape(res.hits.shape)
        res.counts[:] = counts
        gpu_text = O.format_output(res, qcs[0], db, opt)
        ref = _ref_binary(spec.ref_binary)
        t0 = time.perf_counter()
        if ref:
            cmd = [ref, "aln", "-i", os.path.join(tmp, "q"), "-d", os.path.join(tmp, "db"), "-o",
                   os.path.join(tmp, "out.txt")]
            if spec.max_list_length != 1 << 27:
                cmd += ["-l", str(spec.max_list_length >> 20)]
            subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
            wall = time.perf_counter() - t0
            cpu_text = open(os.path.join(tmp, "out.txt"), encoding="latin-1").read()
            kind = "reference"
        else:
            r = O.align_chunk(qcs[0], db, opt)
            wall = time.perf_counter() - t0
            cpu_text = O.format_output(r, qcs[0], db, opt)
            kind = "port"
        return {"value": cells / wall / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                "binary": spec.ref_binary if ref else None,
                "sample": f"{qs.shape[0]} queries x {spec.length} aa vs the first "
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
        seq, starts = workloads.synth_chunk(1, 0, int(a.ref_sample_mib * (1 << 20)))
        per_proc = a.ref_queries_per_core
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
            "config": config_dict(specs(a)[0], int(os.environ.get("WORLD_SIZE", 1))),
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
