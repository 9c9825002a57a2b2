#!/usr/bin/env python
"""Benchmark of the counts -> normalise -> HVG -> scale -> PCA -> kNN -> graph path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3] [--impl reference]

Default workload: C3 = BASELINE.json configs[2], the 1M-cell x 33k-gene matrix the metric is quoted on (it fits one
B200).  One "step" = one pass of the whole hot path over one synthetic count matrix.
  value   cells/s with the sharded CSC counts already resident in HBM when the timed region starts
          (CUDA events on the library's stream, max over ranks).
  e2e     the same metric through the reference-facing host API with HOST buffers: H2D of the Int64/Float64
          CSC arrays and D2H of the embedding, the kNN table and the adjacency inside the timed region.
  roofline  the dominant kernel (the all-pairs kNN scan): algorithmic flops 2*N_q*N_keys*n_pcs per launch
          over its CUDA-event time, against the measured tensor peak.
  cpu_baseline  the oracle (restated reference algorithm, exact kNN) on a bounded cell sample of the same
          workload on the host cores.
`--impl reference` times that CPU restatement alone (Julia is not installed: see DESIGN.md) on ALL host cores.
Rendezvous for N > 1: launched by torchrun (one rank per GPU), NCCL for the data-path collectives.
"""
from __future__ import annotations

import os
import sys

# The CPU legs (the reference arm, the cpu_baseline of the 1-GPU run) use every host core.  torchrun exports
# OMP_NUM_THREADS=1 to its workers, and BLAS reads the variable when numpy is imported: set it first.
_WORLD_ENV = int(os.environ.get("WORLD_SIZE", "1"))
if "reference" in sys.argv or _WORLD_ENV == 1:
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse  # noqa: E402
import json  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_genes, n_cells, density, n_features, n_pcs, k)  -- BASELINE.json configs
    "C1": (33000, 10000, 0.05, 2000, 50, 20),
    "C2": (5000, 500000, 0.05, 2000, 50, 20),
    "C3": (33000, 1000000, 0.05, 2000, 50, 20),
    "C4": (18000, 2000000, 0.005, 2000, 50, 20),
    "tiny": (2000, 20000, 0.05, 500, 20, 10),
}
WORKLOAD_DESC = {
    "C1": "scRNA-seq tutorial scale 10k cells x 33k genes d=5% (configs[0])",
    "C2": "Xenium-scale 500k cells x 5k-gene panel d=5% (configs[1])",
    "C3": "1M cells x 33k genes d=5% atlas (configs[2], the configuration the metric is quoted on)",
    "C4": "VisiumHD 2M spots x 18k genes d=0.5% (configs[3])",
    "tiny": "20k cells x 2k genes (debug)",
}
# cells of the CPU samples (reference arm / cpu_baseline): sized so that one oracle pass takes a few seconds
REF_SAMPLE = {"C1": 10000, "C2": 8000, "C3": 6000, "C4": 12000, "tiny": 4000}


def workload_config(name, world, graph):
    """The `config` object of the JSON line: identical for the b200 arm and the reference arm."""
    n_genes, n_cells, density, n_features, n_pcs, k = WORKLOADS[name]
    return {"workload": WORKLOAD_DESC[name], "n_genes": n_genes, "n_cells": n_cells, "density": density,
            "n_features": n_features, "n_pcs": n_pcs, "k": k, "graph": graph,
            "parallelism": f"cells sharded over {world} GPU(s)",
            "l2": "inputs larger than L2 (CSC shard >= 1 GB) and a 256 MiB flush between iterations",
            "ref_sample": f"the CPU reference arm / cpu_baseline time the first {REF_SAMPLE[name]} cells of this workload "
                          f"(its cells/s is measured on that sample, NOT on {n_cells} cells: the exact SVD is O(N H^2) and the "
                          f"brute-force kNN O(N^2), so the figure does not extrapolate neutrally)"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (pynvml, 100 ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_threads():
    """What the BLAS behind numpy actually runs with (threadpoolctl), next to the core count."""
    info = {"cores": os.cpu_count(), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}
    try:
        from threadpoolctl import threadpool_info
        info["blas_threads"] = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        pass
    return info


def cpu_pipeline(counts, n_features, n_pcs, k):
    """The oracle restatement, end to end, on the host cores (checker / baseline only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cellscopes_oracle as oracle
    # the reference's object constructor drops all-zero genes/cells first (utils.jl:313-320)
    counts, _, _ = oracle.subset_matrix(counts)
    t0 = time.perf_counter()
    oracle.run_pipeline(counts, nFeatures=min(n_features, counts.shape[0]), n_pcs=n_pcs, k=k, pca_method="svd")
    return time.perf_counter() - t0


def cpu_sample_note(sample, n_cells, dt=None):
    th = cpu_threads()
    s = (f"first {sample} of the workload's {n_cells} cells through the whole path (exact LAPACK SVD + exact brute-force "
         f"kNN); cells/s measured on that sample, not extrapolated to {n_cells}; restated reference algorithm "
         f"(NumPy/SciPy/OpenBLAS, {th.get('blas_threads', '?')} BLAS threads on {th['cores']} cores), "
         f"JULIA_NUM_THREADS n/a (Julia absent)")
    if dt is not None:
        s += f"; {dt:.1f} s"
    return s


def run_reference(args, wl):
    """--impl reference: the reference's CPU algorithm (restated; Julia absent) on a bounded sample per step."""
    import csb200 as cs
    n_genes, n_cells, density, n_features, n_pcs, k = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(n_cells, args.ref_sample or REF_SAMPLE[args.workload])
    p = cs.synth.make_params(n_genes, n_cells, density=density, seed=1234)
    counts = cs.synth.sample_counts(p, 0, sample)
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_pipeline(counts, n_features, n_pcs, k)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    v = sample / t
    th = cpu_threads()
    line = {
        "impl": "reference", "metric": "cells/sec counts->PCA->kNN/SNN", "value": v, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, args.graph),
        "cpu_baseline": {"value": v, "unit": "cells/s", "cores": th.get("blas_threads", th["cores"]), "kind": "port",
                         "sample": cpu_sample_note(sample, n_cells)},
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "threads": th,
    }
    emit_json(line)


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    stdout when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the JSON line
    goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=-1, help="cells in the cpu_baseline sample (0 = skip, -1 = per workload)")
    ap.add_argument("--ref-sample", type=int, default=0, help="cells per step of the reference arm (0 = per workload)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the 1-GPU re-run that checks the sharded results")
    ap.add_argument("--graph", default="sum", choices=["sum", "binary", "snn"])
    ap.add_argument("--k", type=int, default=0, help="override the neighbour count of the workload")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.k:
        wl = wl[:5] + (args.k,)
        WORKLOADS[args.workload] = wl
    if args.impl == "reference":
        run_reference(args, wl)
        return

    import csb200 as cs
    n_genes, n_cells, density, n_features, n_pcs, k = wl
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = cs.Context(local_rank)
    if world > 1:
        # the data-path collectives are NCCL calls inside libcellscopes_cuda.so on the context's stream
        # (CSC_BENCH_COMM=torch: the round-1 route through torch.distributed, for comparison)
        comm = cs.TorchComm(device=torch.device("cuda", local_rank)) if os.environ.get("CSC_BENCH_COMM") == "torch" \
            else cs.NcclComm.from_torch(ctx)
    else:
        comm = cs.LocalComm()
    offsets = cs.partition_cells(n_cells, world)
    lo, hi = int(offsets[rank]), int(offsets[rank + 1])
    params = cs.PathParams(n_features=n_features, n_pcs=n_pcs, k=k, graph_mode=args.graph)

    # ---- synthetic input, generated on the device --------------------------------
    sp_params = cs.synth.make_params(n_genes, n_cells, density=density, seed=1234)
    raw = ctx.generate_counts(sp_params, lo, hi)
    nnz_local = raw.nnz

    def barrier():
        ctx.synchronize()
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def one_step():
        return cs.run_path(ctx, comm, raw, n_cells, lo, params, offsets)

    # ---- device-resident metric ------------------------------------------------------------
    for _ in range(args.warmup):
        out = one_step()
        del out
    ctx.stage_times(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms = []
    out = last = None
    for _ in range(args.steps):
        out = last = None          # the previous step's device results are released before the next one starts
        ctx.flush_l2()
        barrier()
        ctx.timer_start()
        out = one_step()
        ms = ctx.timer_stop()
        barrier()
        step_ms.append(ms)
        last = out
    sampler.stop_flag.set()
    sampler.join()
    stage_ms, launches = ctx.stage_times()
    total_ms = float(np.sum(step_ms))
    if rank == 0:
        print("[bench] step ms", [round(x, 2) for x in step_ms], "stage ms/step",
              {kk: round(v / args.steps, 3) for kk, v in stage_ms.items() if v > 0}, file=sys.stderr, flush=True)
    if dist is not None:
        import torch
        t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n_cells / (ms_per_step * 1e-3)

    # solver / data diagnostics of the last step
    info = last["pca"].info()
    pv = info["principalvars"]
    diag = {"pca_iters": last["pca"].iters, "pca_resid": last["pca"].resid,
            "lambda_1": float(pv[0]), "lambda_k": float(pv[-1]), "tvar": float(info["tvar"]),
            "mean_rest_eigenvalue": float((info["tvar"] - pv.sum()) / max(1, info["n_feat"] - pv.size)),
            "min_rel_gap": float(np.min(-np.diff(pv) / pv[:-1])),
            "nnz_local": int(nnz_local), "nnz_selected_local": int(last.get("nnz_selected", nnz_local)),
            "knn_fallback_rows": int(ctx.knn_last_fallback_rows())}
    # margin at the HVG rank boundary (published with every benchmark): relative distance between the last
    # selected and the first rejected variance_standardized
    hv = last.get("hvg") if isinstance(last, dict) else None
    if hv is not None and "order" in hv and hv["order"].size > n_features:
        vs = hv["variance_standardized"][hv["order"]]
        diag["hvg_boundary_rel_margin"] = float((vs[n_features - 1] - vs[n_features]) / max(abs(vs[n_features - 1]), 1e-300))
    # ---- N > 1: the sharded results against a 1-GPU run of the whole matrix on rank 0 ---------
    if dist is not None and not args.no_verify:
        diag["one_vs_n"] = verify_against_one_gpu(cs, ctx, dist, rank, sp_params, n_cells, params, last)
    # ---- roofline of the dominant kernel (kNN scan) -------------------------------------------
    peaks, peak_kind = measured_peaks()
    scan_ms = stage_ms["knn_scan"] / args.steps
    flops = 2.0 * (hi - lo) * n_cells * n_pcs
    achieved = flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else 0.0
    # the scan runs on tcgen05 kind::f16 (fp16 operands, fp32 accumulate): same rate as the measured bf16 GEMM; it
    # runs inside a long step -> the sustained figure
    peak_f16 = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    traffic = None
    for tname in ("r2_traffic.json", "r1_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if traffic is None and os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(f"k_knn_scan_f16@{args.workload}x{world}")
    roofline = {"bound": "tensor", "kernel": "k_knn_scan_f16", "achieved": achieved, "peak": peak_f16, "unit": "TFLOP/s",
                "frac": achieved / peak_f16, "traffic": traffic,
                "peak_kind": f"{peak_kind} bf16 sustained (kind::f16 rate)",
                "ms_per_launch": scan_ms, "share_of_step": scan_ms / ms_per_step,
                "candidates_per_s": (hi - lo) * n_cells / (scan_ms * 1e-3) if scan_ms > 0 else 0.0,
                "note": "algorithmic flops 2*Nq*Nkeys*n_pcs (K padded 50->64 on the MMA).  Every fp32 score has to come out "
                        "of TMEM and through a max tree once; the scan is bound by the instruction streams that do that "
                        "(drain warps and the single MMA-issuing warp), not by the tensor pipe or TMEM bandwidth: see "
                        "DESIGN.md"}
    # the PCA Gram kernel against the 16-bit tensor peak: three fp16 passes (hi'hi, hi'lo, lo'hi) over the upper-triangle
    # tiles (0.5625 of the full square for H = 2000)
    tensor_stages = {}
    gm = stage_ms.get("gram_mma", 0.0) / args.steps
    if gm > 0:
        H = int(info["n_feat"])
        gf = 2.0 * (hi - lo) * H * H
        issued = 3 * 0.5625 * gf / (gm * 1e-3) / 1e12
        tensor_stages["k_gram_tc"] = {"algorithmic_TFLOP/s": gf / (gm * 1e-3) / 1e12, "ms": gm,
                                      "issued_TFLOP/s_3_fp16_passes_upper_tiles": issued, "f16_peak": peak_f16,
                                      "frac_issued": issued / peak_f16}
    nnz_sel = diag["nnz_selected_local"]
    del last, out
    # sparse stages against the HBM roofline (context for the dominant-kernel figure); algorithmic bytes as in
    # SURVEY 8d, with the passes after subset_count charged for the SELECTED entries they read
    hbm = {}
    per_step = {kk: v / args.steps for kk, v in stage_ms.items()}
    n_loc = hi - lo

    def hbm_entry(name, nbytes, ms, **extra):
        if ms > 0:
            gbs = nbytes / (ms * 1e-3) / 1e9
            hbm[name] = {"bytes": nbytes, "ms": ms, "GB/s": gbs, "frac": gbs / peaks["hbm_gbs"], **extra}
    hbm_entry("normalize_fused_raw_moments", 12.0 * nnz_local + 8.0 * n_loc, per_step["normalize"], bytes_per_nnz=12)
    selected = nnz_sel != nnz_local
    if selected:
        # subset_count (count pass: 4 B/nnz; fill pass: 8 B/nnz read + 8 B per kept entry written) + the moments of the
        # kept rows (8 B per kept entry)
        hbm_entry("select_rows_and_moments", 12.0 * nnz_local + 16.0 * nnz_sel + 24.0 * n_loc, per_step["scale_stats"])
    else:
        hbm_entry("scale_moments", 8.0 * nnz_local + 8.0 * n_loc, per_step["scale_stats"])
    # scale -> fp16 hi/lo operands of the Gram: 8 B per entry of the matrix it reads, 4 B per (cell, selected gene) written
    hbm_entry("scale_split", 8.0 * nnz_sel + 4.0 * params.n_features * n_loc, per_step["scale_fill"])
    # embedding = tensor-core GEMM over the fp16 hi/lo operands kept from the Gram pass: 2 x 2 B per (cell, padded gene)
    # read once, 4 B per (cell, component) written
    pitch = ((params.n_features + 1 + 63) // 64) * 64
    hbm_entry("pca_transform_split_gemm", 4.0 * pitch * n_loc + 4.0 * params.n_pcs * n_loc, per_step["pca_transform"],
              note="k_transform_tc (tcgen05 kind::f16, three passes): HBM-bound by design")

    # ---- e2e: host buffers through the reference-facing API -----------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(cs, ctx, comm, dist, raw, n_genes, n_cells, lo, hi, params, offsets, args, world, local_rank)

    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only) -------------------------------
    cpu = None
    cpu_sample = REF_SAMPLE[args.workload] if args.cpu_sample < 0 else args.cpu_sample
    if rank == 0 and world == 1 and cpu_sample > 0:
        sample = min(hi - lo, cpu_sample)
        counts = raw.to_scipy(0, sample)
        dt = cpu_pipeline(counts, n_features, n_pcs, k)
        th = cpu_threads()
        cpu = {"value": sample / dt, "unit": "cells/s", "cores": th.get("blas_threads", th["cores"]), "kind": "port",
               "sample": cpu_sample_note(sample, n_cells, dt)}

    if rank == 0:
        line = {
            "metric": "cells/sec counts->PCA->kNN/SNN", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world, args.graph),
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "stage_ms": {kk: round(v, 3) for kk, v in per_step.items() if v > 0}, "step_ms": [round(x, 2) for x in step_ms],
            "hbm_stages": hbm, "tensor_stages": tensor_stages, "diag": diag,
        }
        emit_json(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def verify_against_one_gpu(cs, ctx, dist, rank, sp_params, n_cells, params, last):
    """N > 1: rank 0 re-runs the whole matrix alone and compares with the sharded run (the all-gathered embedding and
    neighbour table every rank already holds): HVG set, explained variance, embedding, neighbour lists, graph size."""
    import torch
    nnz_local = torch.tensor([last["graph"].info()[2]], dtype=torch.int64, device=f"cuda:{ctx.device}")
    dist.all_reduce(nnz_local)
    res = None
    if rank == 0:
        try:
            emb_n = last["keys"].download(np.float64)
            idx_n = last["knn_idx_all"].download().reshape(n_cells, params.k)
            pv_n = last["pca"].info()["principalvars"]
            raw1 = ctx.generate_counts(sp_params, 0, n_cells)
            one = cs.run_path(ctx, cs.LocalComm(), raw1, n_cells, 0, params)
            e1 = one["embedding"].download(np.float64)
            i1 = one["knn_idx"].download().reshape(n_cells, params.k)
            pv1 = one["pca"].info()["principalvars"]
            sign = np.sign(np.sum(emb_n * e1, axis=0))
            sign[sign == 0] = 1
            scale = np.sqrt((e1 ** 2).mean(axis=0))
            res = {"hvg_identical": bool(np.array_equal(last["hvg"]["features"], one["hvg"]["features"])),
                   "principalvars_max_rel_diff": float(np.max(np.abs(pv_n - pv1) / pv1)),
                   "embedding_max_diff_rel_to_component_rms": float(np.max(np.abs(emb_n * sign - e1) / scale)),
                   "rows_with_identical_neighbour_lists": float((idx_n == i1).all(axis=1).mean()),
                   "neighbour_ids_identical": float((idx_n == i1).mean()),
                   "graph_nnz_n": int(nnz_local.item()), "graph_nnz_1": int(one["graph"].info()[2])}
            del one, raw1
        except Exception as ex:      # the check must never take the benchmark line down
            res = {"error": repr(ex)}
    dist.barrier()
    return res


def run_e2e(cs, ctx, comm, dist, raw, n_genes, n_cells, lo, hi, params, offsets, args, world, local_rank):
    """Host CSC (Int64 indices, Float64 values, page-locked) -> results on the host, every copy inside the timed region."""
    import torch
    n_loc = hi - lo
    # the host copy of this rank's shard lands in page-locked blocks of the context (csc_host_alloc): 16 B per stored entry
    colptr, rowval, nzval = raw.download(0, n_loc, 0)
    h2d = colptr.nbytes + rowval.nbytes + nzval.nbytes
    times = []
    d2h = 0
    call_ms = None
    n_warm = 2      # untimed passes: the first allocates the pinned result buffers, the second re-uses them
    for it in range(max(1, min(args.steps, 3)) + n_warm):
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        if world == 1:
            # the reference-facing calls, in the atlas order of docs/scRNA_tutorial/README.md:260-280
            # the host matrix wraps the pinned Int64/Float64 arrays as they are (scipy's constructor would copy and
            # down-cast the indices, i.e. un-pin them)
            m = cs.csc_from_arrays(nzval, rowval, colptr, (n_genes, n_loc))
            obj = cs.scRNAObject.__new__(cs.scRNAObject)
            obj.rawCount = cs.RawCountObject(m, np.arange(n_loc), np.arange(n_genes))
            obj.normCount = obj.scaleCount = obj.varGene = obj.dimReduction = obj.clustData = None
            obj.metaData, obj.imaging, obj._session = None, False, None
            tc = [time.perf_counter()]
            obj = cs.normalize_object(obj, ctx=ctx)
            tc.append(time.perf_counter())
            obj = cs.find_variable_genes(obj, nFeatures=params.n_features, span=params.span, ctx=ctx)
            tc.append(time.perf_counter())
            obj = cs.scale_object(obj, features=obj.varGene.var_gene, ctx=ctx)
            tc.append(time.perf_counter())
            obj = cs.run_pca(obj, maxoutdim=params.n_pcs, ctx=ctx)
            tc.append(time.perf_counter())
            # the metric ends at the graph (counts -> PCA -> kNN/SNN): the community detection that follows it in
            # run_clustering (csc_leiden) is timed separately (tools/leiden_bench.py)
            obj = cs.run_clustering(obj, n_neighbors=params.k, graph=params.graph_mode, leiden=False, ctx=ctx)
            tc.append(time.perf_counter())
            call_ms = [round((b_ - a_) * 1e3, 1) for a_, b_ in zip(tc[:-1], tc[1:])]
            emb = obj.dimReduction.pca.cell_embedding
            adj = obj.clustData.adj_mat
            d2h = emb.nbytes + obj.clustData.knn_indices.nbytes + obj.clustData.knn_distances.nbytes + \
                adj.indptr.nbytes + adj.indices.nbytes + adj.data.nbytes
            # drop every reference to the results: they live in pinned host buffers of the context's cache, and a
            # buffer still referenced from the previous pass forces a fresh cudaHostAlloc (~300 ms) in this one
            del obj, emb, adj, m
        else:
            shard = ctx.sparse_upload(n_genes, n_loc, colptr, rowval, nzval, 0)
            out = cs.run_path(ctx, comm, shard, n_cells, lo, params, offsets)
            emb = out["embedding"].download(np.float64)
            idx = out["knn_idx"].download()
            dd = out["knn_dist"].download()
            gcp, grv, gv = out["graph"].download(0)
            d2h = emb.nbytes + idx.nbytes + dd.nbytes + gcp.nbytes + grv.nbytes + gv.nbytes
            del out, shard, emb, idx, dd, gcp, grv, gv
        ctx.synchronize()
        dt = time.perf_counter() - t0
        if it >= n_warm:
            times.append(dt)
            if world == 1:
                print(f"[bench] e2e pass {dt * 1e3:.1f} ms; calls [normalize, hvg, scale, pca, clustering] {call_ms}",
                      file=sys.stderr, flush=True)
    t = float(np.mean(times))
    if dist is not None:
        tt = torch.tensor([t], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt.item())
    return {"value": n_cells / t, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": t * 1e3, "api": "normalize_object/find_variable_genes/scale_object/run_pca/run_clustering"
            if world == 1 else "csc_sparse_upload + run_path + downloads per rank",
            "call_ms": call_ms,
            "host_outputs": "copied to the host inside the timed region: pca.cell_embedding (Float64), knn indices/distances, "
                            "the adjacency (Int64 CSC).  normCount.count_mtx and scaleCount.count_mtx stay on the device "
                            "as lazy matrices (downloaded only when read): the reference's consumers on this path never "
                            "read them",
            "h2d_note": "h2d_bytes_per_step counts the host arrays handed to the API (Int64 colptr / rowval, Float64 values: 16 B per "
                        "stored entry).  csc_sparse_upload moves them through two lanes: host threads narrow pieces from the front "
                        "(to uint16 where lossless, else int32 / fp32) and copy 2-4 B per entry, the copy engine carries wide pieces "
                        "from the back; fewer bytes than counted cross the bus, all of them inside the timed region",
            "timer": "host wall clock around synchronised calls (includes H2D, D2H and host marshalling)"}


if __name__ == "__main__":
    main()
