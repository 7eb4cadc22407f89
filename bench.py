#!/usr/bin/env python
"""bench.py -- GCUPS of the all-vs-all Gotoh distance-matrix path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4|c5|c5s]

A step = one all-vs-all pass over one batch of synthetic sequences.
N = 1: BASELINE.json configs[1] (1,000 protein seqs x 300 aa, 499,500 pairs, 4.4955e10 cells).
N > 1 (torchrun, one rank per GPU): the same per-GPU work, weak scaling -- round(1000*sqrt(N))
sequences x 300 aa, the sorted rows cut into N contiguous slabs by the library's planner, the
slabs gathered to rank 0 over NCCL inside the timed step.

value      = cells of the whole job / device time of a step, inputs resident in HBM
             (CUDA events on the stream the kernels run on; max over ranks).
e2e        = same metric through the public call with HOST buffers: encode + sort/pack + H2D from
             pinned staging + kernels + gather + un-sort/distances + D2H of scores and distances.
roofline   = the dominant kernel (packed 16-bit Gotoh) against the integer/DPX issue roofline of
             SURVEY.md 8d: 5 integer lane-ops per cell, two cells per 16x2 instruction, at the DPX
             rate measured live on this GPU (tsq_measure_dpx_rate) and the SM clock seen under load.
cpu_baseline / --impl reference = the CPU oracle (oracle/gotoh_oracle.c, "port": the reference
             has no in-process implementation and no clustalo binary exists in this image) on all
             host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCUPS, all-vs-all Gotoh distance matrix"
UNIT = "GCUPS"


def workload(name: str, n_gpus: int):
    from tweakseq_b200 import synth
    if name == "c2":
        n = 1000 if n_gpus == 1 else int(round(1000 * math.sqrt(n_gpus)))
        seqs = synth.protein(n, 300, 2)
        label = (f"configs[1]: {n} protein seqs x 300 aa all-vs-all"
                 + ("" if n_gpus == 1 else f" (weak scaling of configs[1]: 1000*sqrt({n_gpus}) seqs, same cells per GPU)"))
    elif name == "c3":
        seqs = synth.protein(10000, 400, 3)
        label = "configs[2]: 10,000 protein seqs x 400 aa all-vs-all (strong scaling across ranks)"
    elif name == "c5":
        seqs = synth.protein(100000, 150, 5)
        label = "configs[4]: 100,000 protein seqs x 150 aa all-vs-all (4,999,950,000 pairs; strong scaling across ranks; scores only)"
    elif name == "c4":
        seqs = synth.nucleotide(500, 10000, 30000, 4)
        label = "configs[3]: 500 nucleotide seqs x 10-30 kb all-vs-all (124,750 pairs; packed wavefront kernel)"
    elif name == "c5s":
        seqs = synth.protein(20000, 150, 5)
        label = "configs[4] scaled twin: 20,000 protein seqs x 150 aa all-vs-all"
    else:
        raise SystemExit(f"unknown workload {name}")
    return seqs, label


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


def cpu_oracle_gcups(seqs, budget_s: float, threads: int):
    """Times the oracle on a bounded prefix of the packed pair list.  Returns (gcups, sample)."""
    from oracle import pyoracle as o
    enc = [o.encode(s, 0) for s in seqs]
    mat = o.matrix(0)
    n = len(enc)
    total = n * (n - 1) // 2
    probe = min(total, 4000 * threads)
    t0 = time.perf_counter()
    _, cells = o.all_pairs(enc, mat, 11, 1, nthreads=threads, pair_begin=0, pair_end=probe)
    dt = time.perf_counter() - t0
    rate = cells / dt
    per_pair = cells / probe
    npairs = int(min(total, max(probe, budget_s * rate / per_pair)))
    t0 = time.perf_counter()
    _, cells = o.all_pairs(enc, mat, 11, 1, nthreads=threads, pair_begin=0, pair_end=npairs)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, f"first {npairs} of {total} packed pairs ({cells:.3e} cells, {dt:.1f} s)"


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path.  groundstate/tweakseq has
    none in-process (it execs clustalo, absent from this image), so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    seqs, label = workload(args.workload, args.gpus)
    from oracle import pyoracle as o
    enc = [o.encode(s, 0) for s in seqs]
    mat = o.matrix(0)
    n = len(enc)
    total = n * (n - 1) // 2
    per_step = min(total, 3000 * threads)       # ~0.3-0.5 s of CPU work per step
    steps, warm = args.steps, args.warmup
    cells_total, t_total = 0, 0.0
    for it in range(warm + steps):
        b = (it * per_step) % max(total - per_step, 1)
        t0 = time.perf_counter()
        _, cells = o.all_pairs(enc, mat, 11, 1, nthreads=threads, pair_begin=b, pair_end=b + per_step)
        dt = time.perf_counter() - t0
        if it >= warm:
            cells_total += cells
            t_total += dt
    gcups = cells_total / t_total / 1e9
    clustalo = shutil.which("clustalo")
    line = {"impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": label, "sample_pairs_per_step": per_step},
            "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{per_step} consecutive packed pairs per step, {steps} steps"},
            "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "clustalo": clustalo or "ClustalO not available in image"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import tweakseq_b200 as t
    from tweakseq_b200.distributed import ShardedRun

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    strong = args.workload in ("c3", "c4", "c5")
    seqs, label = workload(args.workload, world)
    flags = t.FLAG_NO_DISTANCES if args.workload == "c5" else 0   # 40 GB of fp64 distances: scores only
    alphabet = 1 if args.workload == "c4" else 0
    run = ShardedRun(seqs, alphabet=alphabet, flags=flags, device=local)
    run.upload()
    from tweakseq_b200 import synth
    cells_total = synth.total_cells(seqs)

    # L2 flush buffer: inputs (0.3 MB) are far smaller than the 126 MB L2, so flush between steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = run.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed steps (inputs resident in HBM) -----------------------------------------
    for _ in range(args.warmup):
        run.compute()
    barrier()
    dpx_ops, _ = run.ctx.measure_dpx_rate()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kern_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(k & 0xff)             # evict L2 (not timed)
        evs[k][0].record(stream)
        run.compute()                         # kernels (+ NCCL gather for N > 1) on this stream
        evs[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    dev_ms = sum(step_ms)
    if world > 1:
        tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())
    ms_per_step = dev_ms / args.steps
    value = cells_total / (ms_per_step * 1e6)
    st = run.ctx.stats()
    launches_per_step = st["launches"]

    # dominant kernel alone (this rank's share), CUDA events inside the library on the same stream
    run.ctx.compute(); run.ctx.synchronize()
    kst = run.ctx.stats()
    kernel_ms, kernel_cells = kst["kernel_ms"], kst["cells"]

    # ---- end to end through the public call, host buffers ------------------------------------
    e2e_steps = max(3, min(args.steps, 10))
    e2e_warm = 2
    if args.workload == "c5":
        e2e_steps, e2e_warm = 1, 1        # every step moves 20 GB of scores to the host
    from tweakseq_b200.capi import flatten
    host_buf, host_offs = flatten(seqs)     # the job's input as it sits in host memory: ASCII residues
    h2d = d2h = 0
    e2e_t = 0.0
    e2e_launches = 0
    for k in range(e2e_warm + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        run.ctx.set_sequences_flat(host_buf, host_offs)   # host ASCII residues -> encode
        run.upload()                          # sort/pack + H2D (pinned staging)
        run.compute()                         # kernels + gather
        run.finish()                          # rank 0: un-sort + distances + D2H
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        if k >= e2e_warm:
            e2e_t += dt
            s2 = run.ctx.stats()
            h2d, d2h = s2["h2d_bytes"], s2["d2h_bytes"]
            e2e_launches = s2["launches"] + s2["upload_launches"]
    e2e_value = cells_total / (e2e_t / e2e_steps) / 1e9

    if rank == 0:
        sm_mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965
        sms = st["sm_count"]
        peak = sms * dpx_ops * sm_mhz * 1e6 / 2.5 / 1e9          # 5 lane-ops/cell, 16x2 packing
        achieved = kernel_cells / (kernel_ms * 1e6) if kernel_ms > 0 else 0.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:   # dram bytes of one launch of this kernel on this workload, from the committed ncu capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["gotoh16_kernel"]
            if args.workload == "c2" and world == 1:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        lens_bytes = sum(len(s) for s in seqs)
        hbm_alg = (lens_bytes + 4 * st["n_pairs"]) / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "u16x2", "data": "synthetic",
            "config": {"workload": label, "n_sequences": len(seqs), "pairs": len(seqs) * (len(seqs) - 1) // 2,
                       "cells": cells_total, "gap_open": 10 if alphabet else 11, "gap_extend": 1,
                       "matrix": "ACGTN +5/-4 (SURVEY 8c)" if alphabet else "BLOSUM62 (Consensus.cpp:34-59)",
                       "strip_width": st["strip_width"], "l2": "flushed between timed steps (256 MiB fill)",
                       "seed": 20261017 + 2},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_t / e2e_steps, "steps": e2e_steps, "launches_per_step": int(e2e_launches)},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "dpx-alu", "kernel": "wave16_kernel" if kst["cells_s32"] > kst["cells_s16"] else "gotoh16_kernel", "achieved": achieved, "peak": peak, "unit": UNIT,
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "traffic_note": ("dram read+write bytes of one launch (ncu --set full, profiles/ncu_gotoh16_c2_r01.txt): the "
                                          "strip-boundary scratch column, not input re-reads; algorithmic bytes are len_i+len_j in, "
                                          "4 B out per pair") if traffic else "no ncu capture of this workload (profiles/ncu_traffic.json)",
                         "algorithmic_bytes": lens_bytes + 4 * st["n_pairs"],
                         "peak_how": f"{sms} SMs x {dpx_ops:.1f} DPX lane-results/clk/SM (measured live) x {sm_mhz} MHz "
                                     "(median under load) / 2.5 instr per cell (5 integer lane-ops, 16x2 packing: SURVEY 8d)",
                         "frac_note": "the SURVEY 8d denominator puts all 5 lane-ops per cell on the DPX (ALU-pipe) rate; ptxas places "
                                      "the two adds on the FMA pipe (IMAD.IADD / VIADD), so frac can pass 1; frac_mix is against the "
                                      "measured issue ceiling of the kernel's own instruction mix (3 DPX + 2 adds + 1 LDS)",
                         "peak_mix": sms * 16.6 * 2 * sm_mhz * 1e6 / 1e9,
                         "frac_mix": achieved / (sms * 16.6 * 2 * sm_mhz * 1e6 / 1e9) if sm_mhz else None,
                         "peak_mix_how": "16.4-16.8 packed cells/clk/SM issued by the inner loop's instruction mix in isolation "
                                         "(tools/pipe_probe2.cu, profiles/pipe_probe2_r01.jsonl), x2 cells, x SMs x MHz",
                         "kernel_ms": kernel_ms, "kernel_cells": kernel_cells,
                         "hbm_algorithmic_gbs": hbm_alg, "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and args.workload == "c2":
            # the "next" row: UPGMA guide tree of the same matrix, device time
            try:
                run.ctx.guide_tree()
                line["guide_tree"] = {"algorithm": "UPGMA", "n": len(seqs), "gpu_ms": run.ctx.stats()["tree_ms"]}
            except Exception as e:   # never let the extra row break the contract line
                line["guide_tree"] = {"error": str(e)}
            # the step after the tree: progressive alignment along it (tsq_msa), wall time of the call
            try:
                rows, _ = run.ctx.msa()
                mst = run.ctx.stats()
                line["msa"] = {"algorithm": "progressive, sum-of-pairs profile Gotoh along the UPGMA tree", "n": len(seqs),
                               "columns": len(rows[0]) if rows else 0, "gpu_ms": mst["msa_ms"],
                               "note": "plan + kernels + copies of one tsq_msa call; one CTA per merge, one launch per tree level"}
            except Exception as e:
                line["msa"] = {"error": str(e)}
        if world == 1 and args.workload == "c2":
            # the other "next" row: identity-aware scoring, 32-bit inter-task kernel
            try:
                with t.Context(flags=t.FLAG_IDENTITY | t.FLAG_NO_DISTANCES, device=local) as ictx:
                    ictx.set_sequences_flat(host_buf, host_offs)
                    ictx.upload()
                    for _ in range(3):
                        ictx.compute(); ictx.synchronize()
                    ist = ictx.stats()
                line["identity_mode"] = {"kernel": "gotoh32_kernel", "kernel_ms": ist["kernel_ms"], "gcups": ist["gcups_kernel"],
                                         "dtype": "int32 keys = score * 2^k + identities"}
            except Exception as e:
                line["identity_mode"] = {"error": str(e)}
        if world == 1 and not args.no_cpu and alphabet == 0:
            threads = os.cpu_count() or 1
            g, sample = cpu_oracle_gcups(seqs, 10.0, threads)
            if "gpu_ms" in line.get("guide_tree", {}):
                from oracle import pyoracle as o
                d = run.ctx.distances()
                t0 = time.perf_counter()
                o.upgma(d, len(seqs))
                line["guide_tree"]["cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t0)
                line["guide_tree"]["cpu_oracle"] = "naive O(n^3) restatement, 1 thread"
            line["cpu_baseline"] = {"value": g, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                    "clustalo": shutil.which("clustalo") or "ClustalO not available in image"}
        print(json.dumps(line), flush=True)
    run.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
