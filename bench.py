#!/usr/bin/env python
"""bench.py -- GCUPS of the all-vs-all Gotoh distance-matrix path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4|c5|c5s]
                    [--inprocess] [--no-cpu] [--no-c3] [--no-plugin]

A step = one all-vs-all pass over one batch of synthetic sequences.

Headline line (what the driver records):
  N = 1 : BASELINE.json configs[1] (1,000 protein seqs x 300 aa, 499,500 pairs, 4.4955e10 cells).
  N > 1 : (torchrun, one rank per GPU, NCCL process group) the same per-GPU work, weak scaling --
          round(1000*sqrt(N)) sequences x 300 aa, the sorted rows cut into N slabs by the library's planner.
Every line ALSO carries a "c3" block: BASELINE.json configs[2] (10,000 x 400 aa, 49,995,000 pairs,
8.0e12 cells), the multi-GPU configuration, STRONG-scaled over the N ranks -- device time, end to end,
per-rank kernel ms -- so that its 1/2/4/8-GPU efficiency can be computed from the driver's own runs.
Every block carries a "parity" object computed OUTSIDE the timed regions: >= 10^4 seeded sampled pairs
plus the complete first and last rows of the result the end-to-end step delivered to the host,
against the CPU oracle (mismatches must be 0).  --workload c4 / c5 make BASELINE configs[3] / [4] the
headline (>= 500 / >= 10^6 sampled pairs).

value      = cells of the whole job / device time of a step, inputs resident in HBM
             (CUDA events on the stream the kernels run on; max over ranks).
e2e        = same metric through the public calls with HOST buffers: encode + sort/pack + H2D from
             pinned staging + kernels + un-sort/distances + D2H of scores and distances.  N > 1 with
             fixed-length input: every rank finalizes its slab and copies it over its OWN PCIe link into
             one shared host result (tweakseq_b200/distributed.py); ragged input: NCCL gather to rank 0.
e2e_plugin = (N = 1) the call the editor would make: tsq_run_fasta, FASTA file in -> aligned FASTA out,
             with the stage split the library logs.
roofline   = the dominant kernel against (a) SURVEY.md 8d's integer/DPX roofline (5 lane-ops per cell,
             two cells per 16x2 instruction, at the DPX rate measured live) -- "frac" -- and (b) the
             two-pipe bound "frac_dpx_issue": min(DPX pipe: 3 DPX per packed cell; issue: 6 slots per
             packed cell) from rates measured live on this GPU (tsq_measure_pipe_rates), plus the
             isolated instruction mix's own ceiling (frac_mix).
cpu_baseline / --impl reference = the CPU oracle (oracle/gotoh_oracle.c, "port": the reference
             has no in-process implementation; ClustalO --full is run when a clustalo binary is on
             PATH, else reported absent) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS, all-vs-all Gotoh distance matrix"
UNIT = "GCUPS"
ORACLE_NOTE = "oracle/gotoh_oracle.c (CPU restatement; parity unpinned by the reference, which holds no Gotoh code or vectors)"


def workload(name: str, n_gpus: int):
    """(sequences, label, alphabet, strong-scaled?, seed index)"""
    from tweakseq_b200 import synth
    if name == "c2":
        n = 1000 if n_gpus == 1 else int(round(1000 * math.sqrt(n_gpus)))
        seqs = synth.protein(n, 300, 2)
        label = (f"configs[1]: {n} protein seqs x 300 aa all-vs-all"
                 + ("" if n_gpus == 1 else f" (weak scaling of configs[1]: 1000*sqrt({n_gpus}) seqs, same cells per GPU)"))
        return seqs, label, 0, False, 2
    if name == "c3":
        return (synth.protein(10000, 400, 3),
                "configs[2]: 10,000 protein seqs x 400 aa all-vs-all (strong scaling across ranks)", 0, True, 3)
    if name == "c5":
        return (synth.protein(100000, 150, 5),
                "configs[4]: 100,000 protein seqs x 150 aa all-vs-all (4,999,950,000 pairs; strong scaling across ranks; scores only)",
                0, True, 5)
    if name == "c4":
        return (synth.nucleotide(500, 10000, 30000, 4),
                "configs[3]: 500 nucleotide seqs x 10-30 kb all-vs-all (124,750 pairs; packed wavefront kernel)", 1, True, 4)
    if name == "c5s":
        return synth.protein(20000, 150, 5), "configs[4] scaled twin: 20,000 protein seqs x 150 aa all-vs-all", 0, True, 5
    if name == "c1":
        return synth.config(1)[1], "configs[0]: 100 protein seqs ~300 aa", 0, True, 1
    raise SystemExit(f"unknown workload {name}")


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), 0
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        mhz = sorted(m for m, _ in self.samples)
        # the region is short: every sample inside it counts as "under load"
        med = mhz[len(mhz) // 2] if mhz else 0
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(mhz), "sm_mhz_max_seen": mhz[-1] if mhz else 0}


# ---- CPU side: oracle timing, parity check, ClustalO leg --------------------------------------------------
def cpu_oracle_gcups(seqs, alphabet: int, budget_s: float, threads: int):
    """Times the oracle on a bounded prefix of the packed pair list.  Returns (gcups, sample)."""
    from oracle import pyoracle as o
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    go = 10 if alphabet else 11
    n = len(enc)
    total = n * (n - 1) // 2
    probe = min(total, 4000 * threads)
    t0 = time.perf_counter()
    _, cells = o.all_pairs(enc, mat, go, 1, nthreads=threads, pair_begin=0, pair_end=probe)
    dt = time.perf_counter() - t0
    rate = cells / dt
    per_pair = cells / probe
    npairs = int(min(total, max(probe, budget_s * rate / per_pair)))
    t0 = time.perf_counter()
    _, cells = o.all_pairs(enc, mat, go, 1, nthreads=threads, pair_begin=0, pair_end=npairs)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, f"first {npairs} of {total} packed pairs ({cells:.3e} cells, {dt:.1f} s)"


def cpu_simd_gcups(seqs, alphabet: int, budget_s: float, threads: int):
    """Times the SIMD CPU kernel (oracle/gotoh_simd.c: one subject per int16 lane, 32 lanes; AVX-512BW / AVX2 by
    target_clones; bit-identical to the scalar port) on a bounded prefix of the rows.  Returns (gcups, sample)."""
    from oracle import pyoracle as o
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    go = 10 if alphabet else 11
    n = len(enc)
    probe_rows = min(n - 1, max(1, 2 * threads))
    t0 = time.perf_counter()
    _, cells = o.rows_simd(enc, mat, go, 1, nthreads=threads, row_begin=0, row_end=probe_rows)
    dt = max(time.perf_counter() - t0, 1e-6)
    rows = int(min(n - 1, max(probe_rows, budget_s * (cells / dt) / max(cells / probe_rows, 1))))
    t0 = time.perf_counter()
    _, cells = o.rows_simd(enc, mat, go, 1, nthreads=threads, row_begin=0, row_end=rows)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, f"rows 0..{rows} of {n} against all later sequences ({cells:.3e} cells, {dt:.1f} s)"


def parity_block(seqs, alphabet: int, scores, dist, n_sample: int, seed: int, threads: int):
    """Sampled pairs + complete first and last rows of the DELIVERED host result against the oracle."""
    import numpy as np
    from oracle import pyoracle as o
    t0 = time.perf_counter()
    n = len(seqs)
    total = n * (n - 1) // 2
    if total == 0:
        return {"pairs_checked": 0, "mismatches": 0}
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    go = 10 if alphabet else 11
    rng = np.random.default_rng(seed)
    if n_sample >= total:
        idx = np.arange(total, dtype=np.int64)                                 # small job: every pair
    else:
        idx = np.unique(rng.integers(0, total, int(n_sample * 1.02) + 16, dtype=np.int64))   # a few collisions at most
    last_rows = min(3, n - 1)
    first_row = np.arange(0, n - 1, dtype=np.int64)                            # pairs (0, j)
    tail = np.arange(total - last_rows * (last_rows + 1) // 2, total, dtype=np.int64)   # rows n-1-last_rows .. n-2
    idx = np.unique(np.concatenate([idx, first_row, tail]))
    ii = np.arange(n, dtype=np.int64)
    rowstart = ii * n - ii * (ii + 1) // 2
    pi = np.searchsorted(rowstart[: n - 1], idx, side="right") - 1
    pj = idx - rowstart[pi] + pi + 1
    ref, cells = o.pair_list(enc, pi.astype(np.uint32), pj.astype(np.uint32), mat, go, 1, nthreads=threads)
    got = np.asarray(scores)[idx]
    mism = int((got != ref).sum())
    out = {"pairs_checked": int(len(idx)), "sampled": int(min(n_sample, total)), "first_row_pairs": int(n - 1),
           "last_rows": int(last_rows), "mismatches": mism, "cells_checked": int(cells), "seed": seed, "oracle": ORACLE_NOTE}
    if dist is not None:
        selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int64)
        mn = np.minimum(selfs[pi], selfs[pj])
        want = np.ones(len(idx), dtype=np.float64)
        ok = mn > 0
        want[ok] = 1.0 - ref[ok].astype(np.float64) / mn[ok].astype(np.float64)   # the oracle's two IEEE operations
        out["distance_mismatches"] = int((np.asarray(dist)[idx].view(np.int64) != want.view(np.int64)).sum())
    out["seconds"] = round(time.perf_counter() - t0, 2)
    return out


def clustalo_leg(seqs, our_dist):
    """when a clustalo binary resolves on PATH, run the reference wrapper's argv
    (Core/ClustalO.cpp:51) + --full --distmat-out on configs[0] and report its wall time and the rank
    correlation of its distances with ours.  Never estimated, never faked."""
    exe = shutil.which("clustalo")
    if not exe:
        return {"available": False, "note": "ClustalO not available in image"}
    import numpy as np
    from tweakseq_b200.fasta import write_fasta, read_distmat
    with tempfile.TemporaryDirectory() as td:
        fin, fout, fmat = os.path.join(td, "in.fa"), os.path.join(td, "out.fa"), os.path.join(td, "dist.mat")
        labels = [f"s{k}" for k in range(len(seqs))]
        write_fasta(fin, labels, seqs, [f">{l}" for l in labels])
        argv = [exe, "--force", "-v", "--outfmt=fa", "--output-order=tree-order", "-i", fin, "-o", fout,
                "--full", f"--distmat-out={fmat}"]
        t0 = time.perf_counter()
        try:
            pr = subprocess.run(argv, capture_output=True, text=True, timeout=600)
        except Exception as e:
            return {"available": True, "path": exe, "error": str(e)}
        dt = time.perf_counter() - t0
        res = {"available": True, "path": exe, "argv": " ".join(argv[1:]), "seconds": dt, "returncode": pr.returncode,
               "threads": "clustalo default", "n": len(seqs)}
        try:
            lab, rows = read_distmat(fmat)
            n = len(seqs)
            order = [lab.index(l) for l in labels]
            theirs = np.array([rows[order[i]][order[j]] for i in range(n) for j in range(i + 1, n)])
            from scipy.stats import spearmanr
            res["spearman_vs_ours"] = float(spearmanr(theirs, np.asarray(our_dist)).correlation)
            res["note"] = "clustalo --full reports k-tuple distances for unaligned input: rank correlation, not bit-equality"
        except Exception as e:
            res["distmat_error"] = str(e)
        return res


def synth_cells(seqs) -> int:
    from tweakseq_b200 import synth
    return synth.total_cells(seqs)


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  groundstate/tweakseq has
    none in-process (it execs clustalo, absent from this image), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    seqs, label, alphabet, strong, _ = workload(args.workload, args.gpus)
    from oracle import pyoracle as o
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    go = 10 if alphabet else 11
    n = len(enc)
    total = n * (n - 1) // 2
    # The CPU arm is the better of the oracle's two kernels: the inter-sequence SIMD one (gotoh_simd.c; int16 lanes,
    # AVX-512BW / AVX2) where the scores fit 16 bits, on all host threads.  A step = a block of rows against all
    # later sequences, sized for ~0.5 s.
    steps, warm = args.steps, args.warmup
    rows_per_step = max(1, min(n - 1, 4 * threads))
    t0 = time.perf_counter()
    _, c0 = o.rows_simd(enc, mat, go, 1, nthreads=threads, row_begin=0, row_end=rows_per_step)
    dt0 = max(time.perf_counter() - t0, 1e-6)
    rows_per_step = int(max(1, min(n - 1, rows_per_step * 0.5 / dt0)))
    cells_total, t_total = 0, 0.0
    for it in range(warm + steps):
        b = (it * rows_per_step) % max(n - 1 - rows_per_step, 1)
        t0 = time.perf_counter()
        _, cells = o.rows_simd(enc, mat, go, 1, nthreads=threads, row_begin=b, row_end=b + rows_per_step)
        dt = time.perf_counter() - t0
        if it >= warm:
            cells_total += cells
            t_total += dt
    gcups = cells_total / t_total / 1e9
    # the scalar port beside it (what round 1 reported as the CPU arm)
    per_step = min(total, 1500 * threads)
    t0 = time.perf_counter()
    _, sc = o.all_pairs(enc, mat, go, 1, nthreads=threads, pair_begin=0, pair_end=per_step)
    scalar_gcups = sc / (time.perf_counter() - t0) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "int16 SIMD lanes (exact; int32 scalar beyond the 16-bit range)", "data": "synthetic",
            "config": {"workload": label, "n_sequences": n, "pairs": total, "cells": synth_cells(seqs),
                       "gap_open": go, "gap_extend": 1,
                       "matrix": "ACGTN +5/-4 (SURVEY 8c)" if alphabet else "BLOSUM62 (Consensus.cpp:34-59)",
                       "seed": 20261017 + (4 if alphabet else 2),
                       "sample_rows_per_step": rows_per_step},
            "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{rows_per_step} rows against all later sequences per step, {steps} steps",
                             "kernel": "oracle/gotoh_simd.c: inter-sequence SIMD, one subject per int16 lane x 32 (AVX-512BW / AVX2 by "
                                       "target_clones), bit-identical to the scalar port",
                             "scalar_port_gcups": scalar_gcups,
                             "note": "the reference holds no Gotoh code and no clustalo binary exists in the image: this is the "
                                     "oracle port's SIMD kernel on all host threads, the competent CPU path; the scalar int32 port "
                                     "(round 1's CPU arm) is quoted beside it"},
            "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "clustalo": shutil.which("clustalo") or "ClustalO not available in image"}
    print(json.dumps(line), flush=True)


# ---- one workload, timed on the device and end to end ------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.inprocess = args.inprocess and args.gpus > 1
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if self.inprocess and self.world > 1:
            raise SystemExit("--inprocess drives all GPUs from ONE process: run it without torchrun")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.n_gpus = args.gpus if self.inprocess else self.world
        self.threads = os.cpu_count() or 1
        # L2 flush buffer: inputs (0.3 MB) are far smaller than the 126 MB L2, so flush between steps
        self.flush = [torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}")
                      for d in (range(args.gpus) if self.inprocess else [self.local])]

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        if self.inprocess:
            for d in range(self.n_gpus):
                self.torch.cuda.synchronize(d)
        else:
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        tt = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def all_ranks(self, x: float):
        if self.world == 1:
            return [x]
        tt = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        out = [self.torch.zeros_like(tt) for _ in range(self.world)]
        self.dist.all_gather(out, tt)
        return [float(v.item()) for v in out]


class InProcessRun:
    """ShardedRun's interface over ONE multi-device context (tsq_params.n_devices): the C-ABI path a
    plugin takes on a multi-GPU box, no torchrun, no process group."""

    def __init__(self, seqs, alphabet, flags, n_devices):
        import tweakseq_b200 as t
        self.ctx = t.Context(alphabet=alphabet, flags=flags, device=0, n_devices=n_devices)
        self.ctx.set_sequences(seqs)
        self.stream = None
        self.sharded = False

    def upload(self):
        self.ctx.upload()
        self.sharded = self.ctx.results_sharded()

    def compute(self):
        self.ctx.compute()

    def finish(self):
        self.ctx.download()

    def scores(self):
        return self.ctx.scores(copy=False)

    def distances(self):
        return self.ctx.distances(copy=False)

    def close(self):
        self.ctx.close()


def time_workload(env: Env, name: str, steps: int, warmup: int, e2e_steps: int, e2e_warm: int, n_sample: int,
                  sample_clocks: bool):
    """Returns (block dict on rank 0 else None, extras for the caller)."""
    import numpy as np
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    from tweakseq_b200.capi import flatten
    from tweakseq_b200.distributed import ShardedRun
    torch = env.torch
    seqs, label, alphabet, strong, seed_ix = workload(name, env.n_gpus)
    flags = t.FLAG_NO_DISTANCES if name == "c5" else 0     # 40 GB of fp64 distances: scores only
    cells_total = synth.total_cells(seqs)
    run = InProcessRun(seqs, alphabet, flags, env.n_gpus) if env.inprocess else ShardedRun(seqs, alphabet=alphabet, flags=flags, device=env.local)
    run.upload()
    streams = [run.stream] if run.stream is not None else []

    def flush_l2(k):
        if env.inprocess:
            for d, f in enumerate(env.flush):
                f.fill_(k & 0xff)
            for d in range(env.n_gpus):
                torch.cuda.synchronize(d)
        else:
            with torch.cuda.stream(run.stream):
                env.flush[0].fill_(k & 0xff)

    # ---- device-timed steps (inputs resident in HBM) -----------------------------------------
    for _ in range(warmup):
        run.compute()
    env.barrier()
    sampler = ClockSampler(env.local) if sample_clocks else None
    if sampler:
        sampler.start()
    step_ms = []
    env.barrier()
    t_wall0 = time.perf_counter()
    if env.inprocess:
        # one process, N devices: the library times each device with CUDA events on that device's own stream;
        # a step = the slowest device (they run side by side)
        for k in range(steps):
            flush_l2(k)
            run.compute()
            run.ctx.synchronize()
            step_ms.append(run.ctx.stats()["kernel_ms"])
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush_l2(k)                           # evict L2 (not timed)
            evs[k][0].record(run.stream)
            run.compute()                         # this rank's kernels (+ NCCL gather when the input is ragged)
            evs[k][1].record(run.stream)
        env.barrier()
        step_ms = [a.elapsed_time(b) for a, b in evs]
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    dev_ms = env.max_over_ranks(sum(step_ms))
    ms_per_step = dev_ms / steps
    value = cells_total / (ms_per_step * 1e6)
    st = run.ctx.stats()
    launches_per_step = st["launches"]

    # this rank's kernels alone, CUDA events inside the library on the same stream
    run.ctx.compute(); run.ctx.synchronize()
    kst = run.ctx.stats()
    kernel_ms, kernel_cells = kst["kernel_ms"], kst["cells"]
    if env.inprocess:
        per_rank_ms = [run.ctx.device_stats(d)["kernel_ms"] for d in range(env.n_gpus)]
        d0 = run.ctx.device_stats(0)
        kernel_ms, kernel_cells = d0["kernel_ms"], d0["cells"]
    else:
        per_rank_ms = env.all_ranks(kernel_ms)

    # ---- end to end through the public calls, host buffers ------------------------------------
    host_buf, host_offs = flatten(seqs)     # the job's input as it sits in host memory: ASCII residues
    run.ctx.stream_results(True)            # finished row ranges leave for the host behind their launch (as tsq_run does)
    h2d = d2h = 0
    e2e_t = 0.0
    e2e_launches = 0
    e2e_stages = {}
    for k in range(e2e_warm + e2e_steps):
        env.barrier()
        t0 = time.perf_counter()
        run.ctx.set_sequences_flat(host_buf, host_offs)   # host ASCII residues -> encode
        run.upload()                          # sort/pack + H2D (pinned staging)
        run.compute()                         # kernels (+ gather)
        run.finish()                          # finalize + D2H: every rank its slab (sharded) or rank 0 everything
        if not env.inprocess:
            torch.cuda.synchronize()
        dt = env.max_over_ranks(time.perf_counter() - t0)
        if k >= e2e_warm:
            e2e_t += dt
            s2 = run.ctx.stats()
            h2d, d2h = s2["h2d_bytes"], s2["d2h_bytes"]
            e2e_launches = s2["launches"] + s2["upload_launches"]
            e2e_stages = {"encode": round(s2["encode_ms"], 3), "sort_pack_h2d": round(s2["upload_ms"], 3),
                          "kernels": round(s2["kernel_ms"], 3), "wait_for_results": round(s2["download_ms"], 3)}
    if env.world > 1:     # whole-job bytes: every rank copies its own share
        tt = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        env.dist.all_reduce(tt)
        h2d, d2h = int(tt[0].item()), int(tt[1].item())
    e2e_value = cells_total / (e2e_t / e2e_steps) / 1e9

    block, extra = None, {"run": run, "seqs": seqs, "alphabet": alphabet, "host": (host_buf, host_offs), "clocks": clocks,
                          "kernel_ms": kernel_ms, "kernel_cells": kernel_cells, "stats": kst, "t_wall": t_wall,
                          "launches_per_step": launches_per_step}
    if env.rank == 0:
        # ---- parity of what the end-to-end step delivered, outside every timed region -------------
        scores = run.scores()
        dist = None if flags & t.FLAG_NO_DISTANCES else run.distances()
        par = parity_block(seqs, alphabet, scores, dist, n_sample, 20261017 + 100 * seed_ix + env.n_gpus, env.threads)
        block = {
            "workload": label, "n_sequences": len(seqs), "pairs": len(seqs) * (len(seqs) - 1) // 2, "cells": cells_total,
            "scaling": "strong" if strong else "weak", "n_gpus": env.n_gpus,
            "value": value, "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
            "kernel_ms_per_rank": [round(x, 4) for x in per_rank_ms],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_t / e2e_steps, "steps": e2e_steps, "launches_per_step": int(e2e_launches),
                    "stages_ms_rank0": e2e_stages,
                    "results": ("one multi-device context (tsq_params.n_devices): " if env.inprocess else "") +
                               ("every rank finalizes its slab and copies it over its own PCIe link into one shared host result"
                                if run.sharded else ("one device" if env.n_gpus == 1 else
                                                     "slabs gathered on the first device (NCCL send/recv; peer stores in-process), one download"))},
            "parity": par,
        }
    return block, extra


# ---------------------------------------------------------------------------------------------
def plugin_e2e(device: int):
    """The call the editor makes (SeqEditMainWin.cpp:1654-1660 replaced): tsq_run_fasta, FASTA in, aligned
    FASTA out, on configs[0] and configs[1], with the library's own stage split."""
    import re
    from tweakseq_b200 import capi, synth
    from tweakseq_b200.fasta import write_fasta
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for key, seqs in (("c1", synth.config(1)[1]), ("c1_family", synth.protein(100, (200, 400, 300, 30), 1, family=True)),
                          ("c2", synth.config(2)[1])):
            fin, fout = os.path.join(td, key + ".fa"), os.path.join(td, key + ".aln.fa")
            labels = [f"s{k}" for k in range(len(seqs))]
            write_fasta(fin, labels, seqs, [f">{l}" for l in labels])
            best = None
            for rep in range(3):        # first call pays module load and allocations; report the best of the rest
                log = []
                t0 = time.perf_counter()
                rc = capi.run_fasta(fin, fout, log=log.append, flags=capi.FLAG_MSA_OUT, device=device, alphabet=capi.ALPHABET_AUTO)
                dt = 1e3 * (time.perf_counter() - t0)
                if rc != 0:
                    best = {"error": rc}
                    break
                line = next((m for m in log if "timing ms" in m), "")
                stages = {k: float(v) for k, v in re.findall(r"(\w+)=([0-9.]+)", line.split("timing ms:")[-1])}
                if rep > 0 and (best is None or dt < best["wall_ms"]):
                    best = {"wall_ms": dt, "n": len(seqs), "stages_ms": stages}
            out[key] = best
    out["call"] = "tsq_run_fasta(fin, fout, TSQ_FLAG_MSA_OUT, alphabet auto): FASTA file in, aligned FASTA (tree order) out"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--inprocess", action="store_true", help="N > 1 from ONE process through tsq_params.n_devices (no torchrun)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c3", action="store_true", help="skip the configs[2] block")
    ap.add_argument("--no-plugin", action="store_true", help="skip the tsq_run_fasta block")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import tweakseq_b200 as t
    env = Env(args)
    wl = args.workload
    big = wl in ("c3", "c4", "c5", "c5s")
    steps, warm = (args.steps, args.warmup) if not big else (min(args.steps, 3), 3)
    if wl == "c5":
        e2e_steps, e2e_warm = 1, 1            # every step moves 20 GB of scores to the host
    elif big:
        e2e_steps, e2e_warm = 2, 1
    else:
        e2e_steps, e2e_warm = max(3, min(args.steps, 10)), 2
    n_sample = {"c4": 600, "c5": 1_000_000}.get(wl, 10_000)
    head, ex = time_workload(env, wl, steps, warm, e2e_steps, e2e_warm, n_sample, sample_clocks=True)
    run = ex["run"]

    line = None
    if env.rank == 0:
        clocks = ex["clocks"] or {}
        kst, kernel_ms, kernel_cells = ex["stats"], ex["kernel_ms"], ex["kernel_cells"]
        rates = run.ctx.measure_pipe_rates()
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or rates["sm_mhz"] or 1965
        sms = kst["sm_count"] // (env.n_gpus if env.inprocess else 1)
        dpx, issue, mix = rates["dpx_per_clk_sm"], rates["issue_per_clk_sm"], rates["mix_packed_cells_per_clk_sm"]
        wave = kst["cells_s32"] > kst["cells_s16"]
        hz = sm_mhz * 1e6
        peak = sms * dpx * hz / 2.5 / 1e9                                  # 5 lane-ops/cell, 16x2 packing
        issue_nominal = 128.0                                              # 4 schedulers x 32 lanes x 1 instruction / clk
        two_pipe_cells = min(dpx / 3.0, max(issue, issue_nominal) / 6.0)   # packed cells / clk / SM
        peak2 = sms * two_pipe_cells * 2 * hz / 1e9
        peak_mix = sms * mix * 2 * hz / 1e9
        achieved = kernel_cells / (kernel_ms * 1e6) if kernel_ms > 0 else 0.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        kname = "wave16_kernel" if wave else "gotoh16_kernel"
        try:   # dram bytes of one launch of this kernel on this workload, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kname]
            if tr.get("workload") == wl and env.n_gpus == 1:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        seqs = ex["seqs"]
        lens_bytes = sum(len(s) for s in seqs)
        alg_bytes = lens_bytes + 4 * kst["n_pairs"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": env.n_gpus, "steps": steps,
            "warmup": warm, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": head["scaling"], "vs_baseline": None, "dtype": "u16x2", "data": "synthetic",
            "config": {"workload": head["workload"], "n_sequences": head["n_sequences"], "pairs": head["pairs"],
                       "cells": head["cells"], "gap_open": 10 if ex["alphabet"] else 11, "gap_extend": 1,
                       "matrix": "ACGTN +5/-4 (SURVEY 8c)" if ex["alphabet"] else "BLOSUM62 (Consensus.cpp:34-59)",
                       "strip_width": kst["strip_width"], "l2": "flushed between timed steps (256 MiB fill)",
                       "seed": 20261017 + (4 if ex["alphabet"] else 2),
                       "launch": "one process, tsq_params.n_devices" if env.inprocess else ("torchrun, one rank per GPU, NCCL" if env.world > 1 else "one process, one GPU")},
            "clocks": clocks,
            "e2e": head["e2e"],
            "parity": head["parity"],
            "kernel_ms_per_rank": head["kernel_ms_per_rank"],
            "gpu_launches": int(ex["launches_per_step"] * steps),
            "roofline": {
                "bound": "dpx-alu", "kernel": kname, "achieved": achieved, "peak": peak, "unit": UNIT,
                "frac": achieved / peak if peak else None,
                "frac_dpx_issue": achieved / peak2 if peak2 else None, "peak_dpx_issue": peak2,
                "frac_mix": achieved / peak_mix if peak_mix else None, "peak_mix": peak_mix,
                "traffic": traffic,
                "traffic_note": (f"dram read+write bytes of one launch (ncu --set full, profiles/ncu_traffic.json: {kname}): the "
                                 "strip-boundary scratch column, not input re-reads") if traffic else
                                "no ncu capture of this workload (profiles/ncu_traffic.json)",
                "algorithmic_bytes": alg_bytes,
                "rates_measured_live": {"dpx_lane_results_per_clk_sm": dpx, "issue_thread_instr_per_clk_sm": issue,
                                        "mix_packed_cells_per_clk_sm": mix, "sm_mhz_probe": rates["sm_mhz"], "sm_mhz_used": sm_mhz,
                                        "sms": sms},
                "peak_how": f"SURVEY 8d: {sms} SMs x {dpx:.1f} DPX lane-results/clk/SM x {sm_mhz} MHz / 2.5 instr per cell "
                            "(5 integer lane-ops, 16x2 packing); it charges the two adds of a cell to the DPX pipe although ptxas "
                            "places them on the FMA pipe, so frac can pass 1 -- frac_dpx_issue is the honest figure",
                "peak_dpx_issue_how": f"two-pipe bound: min(DPX pipe {dpx:.1f}/3 DPX per packed cell, issue max(measured {issue:.1f}, "
                                      f"nominal {issue_nominal:.0f})/6 slots per packed cell: 3 DPX + 2 adds + 1 LDS) = "
                                      f"{two_pipe_cells:.2f} packed cells/clk/SM x 2 cells x {sms} SMs x {sm_mhz} MHz",
                "peak_mix_how": "packed cells/clk/SM the inner loop's own instruction mix issues in isolation, dependency-free, "
                                "measured live (tsq_measure_pipe_rates), x 2 cells x SMs x MHz",
                "kernel_ms": kernel_ms, "kernel_cells": kernel_cells,
                "hbm_algorithmic_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0,
                "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
            "wall_s_timed_region": ex["t_wall"],
        }

    # ---- the multi-GPU configuration, strong-scaled, at every N (its own block) -----------------------------
    c3 = None
    if wl == "c2" and not args.no_c3:
        run.close()
        ex["run"] = run = None
        c3, ex3 = time_workload(env, "c3", 3, 3, 2, 1, 10_000, sample_clocks=False)
        ex3["run"].close()
        if env.rank == 0:
            line["c3"] = c3

    if env.rank == 0:
        ctx = None
        if wl == "c2" and env.n_gpus == 1:
            host_buf, host_offs = ex["host"]
            ctx = t.Context(device=env.local)
            ctx.set_sequences_flat(host_buf, host_offs)
            ctx.run()
            # the "next" row: UPGMA guide tree of the same matrix, device time
            try:
                ctx.guide_tree()
                line["guide_tree"] = {"algorithm": "UPGMA", "n": ctx.n, "gpu_ms": ctx.stats()["tree_ms"]}
            except Exception as e:   # never let the extra row break the contract line
                line["guide_tree"] = {"error": str(e)}
            # the step after the tree: progressive alignment along it (tsq_msa), wall time of the call
            try:
                rows, _ = ctx.msa()
                mst = ctx.stats()
                line["msa"] = {"algorithm": "progressive, sum-of-pairs profile Gotoh along the UPGMA tree", "n": ctx.n,
                               "columns": len(rows[0]) if rows else 0, "gpu_ms": mst["msa_ms"],
                               "note": "plan + kernels + copies of one tsq_msa call; one CTA per merge, one launch per tree level"}
            except Exception as e:
                line["msa"] = {"error": str(e)}
            # the other "next" row: identity-aware scoring, 32-bit inter-task kernel
            try:
                with t.Context(flags=t.FLAG_IDENTITY | t.FLAG_NO_DISTANCES, device=env.local) as ictx:
                    ictx.set_sequences_flat(host_buf, host_offs)
                    ictx.upload()
                    for _ in range(3):
                        ictx.compute(); ictx.synchronize()
                    ist = ictx.stats()
                line["identity_mode"] = {"kernel": "gotoh32_kernel", "kernel_ms": ist["kernel_ms"], "gcups": ist["gcups_kernel"],
                                         "dtype": "int32 keys = score * 2^k + identities"}
            except Exception as e:
                line["identity_mode"] = {"error": str(e)}
            if not args.no_plugin:
                try:
                    line["e2e_plugin"] = plugin_e2e(env.local)
                except Exception as e:
                    line["e2e_plugin"] = {"error": str(e)}
        if env.n_gpus == 1 and not args.no_cpu:
            g, sample = cpu_oracle_gcups(ex["seqs"], ex["alphabet"], 8.0, env.threads)
            gs, ssample = cpu_simd_gcups(ex["seqs"], ex["alphabet"], 8.0, env.threads)
            line["cpu_baseline"] = {"value": gs, "unit": UNIT, "cores": env.threads, "kind": "port", "sample": ssample,
                                    "kernel": "oracle/gotoh_simd.c: inter-sequence SIMD, one subject per int16 lane x 32 (AVX-512BW / "
                                              "AVX2 by target_clones); bit-identical to the scalar port (tests/test_oracle.py)",
                                    "scalar_port": {"value": g, "unit": UNIT, "sample": sample,
                                                    "kernel": "oracle/gotoh_oracle.c: scalar int32, two rolling rows"},
                                    "note": "a reported baseline, not the target: kernel quality is the roofline fraction"}
            if ctx is not None and "gpu_ms" in line.get("guide_tree", {}):
                from oracle import pyoracle as o
                d = ctx.distances()
                t0 = time.perf_counter()
                o.upgma(d, ctx.n)
                line["guide_tree"]["cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
                line["guide_tree"]["cpu_oracle"] = "naive O(n^3) restatement, 1 thread"
            if wl == "c2":
                try:   # the ClustalO --full leg on configs[0], only when a binary exists
                    from tweakseq_b200 import synth
                    c1 = synth.config(1)[1]
                    with t.Context(device=env.local) as cctx:
                        cctx.set_sequences(c1)
                        cctx.run()
                        line["cpu_baseline"]["clustalo"] = clustalo_leg(c1, cctx.distances())
                except Exception as e:
                    line["cpu_baseline"]["clustalo"] = {"error": str(e)}
        if ctx is not None:
            ctx.close()
        print(json.dumps(line), flush=True)
    if run is not None:
        run.close()
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
