#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native interpolation hot path.

Metric (BASELINE.json): FP64 Kriging predictions/sec, local OrdinaryKriging, k = 32.
Workload (config C3): 200 000 random 2-D samples -> 2048 x 2048 CartesianGrid per GPU,
SphericalVariogram(range=50, sill=1, nugget=0.1), maxneighbors=32, mean + variance.
A "step" is one full fitpredict over one batch: grid-bin build + kNN + assemble/factor/solve for every
target of this rank's slab.  With N GPUs the domain grows to 2048 x (2048 N) with 200 000 N samples
(replicated on every GPU by one NCCL broadcast) and rank r predicts slab r: weak scaling.

  value : whole-job predictions/s, samples already resident in HBM, outputs left in HBM
  e2e   : same metric through the public `fitpredict` call with pinned HOST buffers (H2D of the samples
          and D2H of mean/variance/status inside the timed region)
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference algorithm's CPU restatement (oracle/gsm_oracle.c, all host
threads) on a bounded sample of the same workload; Julia is not installed, so that is the only form
in which the reference path can run on this box (cpu_baseline.kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

# ------------------------------------------------------------------------------------------------
# workload C3
# ------------------------------------------------------------------------------------------------
NX = NY = 2048
NS_PER_GPU = 200_000
K = 32
VARIO = dict(kind="spherical", sill=1.0, nugget=0.1, range=50.0)
SEED = 3


def make_samples(n_gpus: int):
    rng = np.random.default_rng(SEED)
    n = NS_PER_GPU * n_gpus
    X = np.empty((n, 2))
    X[:, 0] = rng.uniform(0, NX, size=n)
    X[:, 1] = rng.uniform(0, NY * n_gpus, size=n)
    z = np.sin(X[:, 0] / 20.0) + np.cos(X[:, 1] / 30.0) + 0.1 * rng.standard_normal(n)
    return X, z


def flops_per_prediction(k=K, ncon=1, dim=2, c_gamma=8, ndrift=0):
    """Algorithmic FP64 flops of one local prediction:
    [k(k+1)/2 + k] E + m^3/3 + 2 m^2 + 4 m + 3 k ndrift,  E = 3 dim + 1 + c_gamma,  m = k + ncon."""
    E = 3 * dim + 1 + c_gamma
    m = k + ncon
    return (k * (k + 1) / 2 + k) * E + m ** 3 / 3 + 2 * m * m + 4 * m + 3 * k * ndrift


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The
    sampler starts before the warm-up (nvidia-smi takes a moment to come up) and only the samples whose
    timestamp falls inside [mark_begin, mark_end] are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
                if self.t0 and self.t1 and not (self.t0 <= ts <= self.t1 + datetime.timedelta(milliseconds=30)):
                    continue
                sm.append(float(f[2]))
                mx.append(float(f[3]))
                power.append(float(f[4]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in C (oracle/), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(o, oc, X, z, n_gpus, rows, threads):
    """Predict the first `rows` grid rows of the C3 domain with the C oracle (KD-tree kNN + Bunch-Kaufman)."""
    model = o.ordinary_kriging(o.Variogram(o.SPHERICAL, VARIO["sill"], VARIO["nugget"], VARIO["range"]))
    grid = o.Grid((0.0, 0.0), (1.0, 1.0), (NX, NY * n_gpus))
    t0 = time.perf_counter()
    mu, var, st, _, _ = oc.fitpredict(model, X, z, grid, prob=True, maxneighbors=K, first=0, count=rows * NX,
                                      use_tree=True, nthreads=threads)
    dt = time.perf_counter() - t0
    assert np.isfinite(mu).all()
    return rows * NX, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    o = entry.load_oracle()
    import gsm_oracle_c as oc
    threads = host_threads()
    X, z = make_samples(args.gpus)
    rows = args.cpu_rows
    for _ in range(max(args.warmup, 0)):
        cpu_reference_step(o, oc, X, z, args.gpus, max(rows // 8, 1), threads)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_reference_step(o, oc, X, z, args.gpus, rows, threads)
        tot_n += n
        tot_t += dt
    val = tot_n / tot_t
    line = {
        "impl": "reference", "metric": "FP64 kriging predictions/sec (local OK, k=32)", "value": val,
        "unit": "predictions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "predictions/s", "cores": threads, "kind": "port",
                         "sample": f"first {rows} of {NY * args.gpus} grid rows ({rows * NX} targets) per step, "
                                   f"KD-tree build included; C restatement of the reference algorithm "
                                   f"(Julia is not installed on this image)"},
        "e2e": {"value": val, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n_gpus):
    return {"workload": "C3: local OrdinaryKriging SphericalVariogram(range=50,sill=1,nugget=0.1) maxneighbors=32, "
                        f"{NS_PER_GPU} 2-D samples -> {NX}x{NY} grid per GPU, mean+variance",
            "samples": NS_PER_GPU * n_gpus, "targets": NX * NY * n_gpus, "k": K,
            "parallelism": f"target slabs x{n_gpus}, samples replicated (1 NCCL broadcast)",
            "l2": "256 MiB L2 flush between timed steps"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    gsm = entry.load_package()
    L = gsm._lib
    ctx = gsm.Context(local_rank)
    M = NX * NY                      # targets of this rank's slab
    n = NS_PER_GPU * world

    # ---- samples: generated on rank 0, replicated with one NCCL broadcast --------------------------
    alpha = 1.0 / max(VARIO["range"], float(NX), float(NY * world))    # _scale, models.jl:272-284
    Xs = zs = None
    if rank == 0:
        X, z = make_samples(world)
        assert alpha == 1.0 / max(VARIO["range"], np.abs(X).max(), float(max(NX, NY * world)))
        Xs, zs = X * alpha, z
    if world > 1:
        dist.barrier()            # communicator set-up is not part of the broadcast time
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    coords_d, values_d = gsm.distributed.replicate_samples(Xs, zs, src=0, device=dev)
    e1.record()
    torch.cuda.synchronize()
    bcast_ms = e0.elapsed_time(e1) if world > 1 else 0.0
    first, count = gsm.distributed.shard(M * world, rank, world)
    assert (first, count) == (rank * M, M)

    vario = L.make_variogram(L.GSM_VARIO_SPHERICAL, VARIO["sill"], VARIO["nugget"], alpha * VARIO["range"])
    model = L.make_model(L.GSM_KRIG_ORDINARY)
    tgt = gsm.Context.grid_targets((0.0, 0.0), (alpha * 1.0, alpha * 1.0), (NX, NY * world), first=rank * M, count=M)
    out_mean = torch.empty(M, dtype=torch.float64, device=dev)
    out_var = torch.empty(M, dtype=torch.float64, device=dev)
    out_stat = torch.empty(M, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        """inputs resident in HBM -> outputs in HBM; returns per-phase device ms of this step"""
        ctx.set_samples(coords_d, values_d, dim=2, n=n)
        t_bins = ctx.timings()
        ctx.local_krige(vario, model, tgt, K, 1, 0.0, out_mean=out_mean, out_var=out_var, out_status=out_stat)
        t = ctx.timings()
        return {"bin_ms": t_bins["total_ms"], "knn_ms": t["knn_ms"], "solve_ms": t["solve_ms"], "index_ms": t["index_ms"],
                "total_ms": t_bins["total_ms"] + t["total_ms"], "launches": t_bins["launches"] + t["launches"]}

    fp64_dfma, fp64_dmma = ctx.measure_fp64_peak()
    # third opinion on the FP64 peak: the library DGEMM (cuBLAS through torch; plumbing, not on the timed path)
    fp64_cublas = 0.0
    try:
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        b = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        fp64_cublas = 5 * 2 * 4096.0 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b
    except Exception:
        pass

    # ---- value: device-resident --------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler.mark_begin()
    wall0 = time.perf_counter()
    phases = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        phases.append(device_step())
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark_end()
    clocks = sampler.stop()
    dev_ms = sum(p["total_ms"] for p in phases)
    launches = int(sum(p["launches"] for p in phases))

    # ---- e2e: public API, pinned host buffers, H2D + D2H inside the timed region -------------------
    Xh = gsm.PinnedArray((n, 2), np.float64)
    zh = gsm.PinnedArray((n,), np.float64)
    Xh.array[:] = (coords_d / alpha).cpu().numpy()
    zh.array[:] = values_d.cpu().numpy()
    pm, pv, ps = gsm.PinnedArray((M,), np.float64), gsm.PinnedArray((M,), np.float64), gsm.PinnedArray((M,), np.uint8)
    table = gsm.georef({"z": zh.array}, gsm.PointSet(Xh.array))   # contiguous float64: wrapped, not copied
    grid = gsm.CartesianGrid(NX, NY * world)
    okm = gsm.OrdinaryKriging(gsm.SphericalVariogram(sill=VARIO["sill"], range=VARIO["range"], nugget=VARIO["nugget"]))

    def e2e_step():
        return gsm.fitpredict(okm, table, grid, maxneighbors=K, prob=True, ctx=ctx, first=rank * M, count=M,
                              out={"mean": pm.array, "var": pv.array, "status": ps.array})

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pred = e2e_step()
        chk = float(pm.array[0] + pv.array[-1])   # device -> host result is read on the host
    barrier()
    e2e_wall = time.perf_counter() - t0
    assert np.isfinite(chk)

    # ---- reduce over ranks (max time) --------------------------------------------------------------
    stats = torch.tensor([dev_ms, wall * 1e3, e2e_wall * 1e3, sum(p["solve_ms"] for p in phases),
                          sum(p["knn_ms"] for p in phases), sum(p["bin_ms"] for p in phases),
                          sum(p["index_ms"] for p in phases)],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, e2e_ms, solve_ms, knn_ms, bin_ms, index_ms = [float(x) for x in stats.cpu()]

    if rank == 0:
        steps = args.steps
        value = M * world * steps / (dev_ms * 1e-3)
        e2e_val = M * world * steps / (e2e_ms * 1e-3)
        fpp = flops_per_prediction()
        # dominant kernel = the shared-factor solve kernel; the job-index kernel that feeds it is timed apart
        kern_ms = solve_ms - index_ms
        solve_launch_ms = kern_ms / (steps * max(1, -(-M // (1 << 20))))
        solve_tflops = fpp * M * steps / (kern_ms * 1e-3) / 1e12
        peak = max(fp64_dfma, fp64_dmma, fp64_cublas)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = 16 + 1 + K * 4 * 2   # mean+var+status out, neighbour list written by kNN and read by solve
        line = {
            "metric": "FP64 kriging predictions/sec (local OK, k=32)", "value": value, "unit": "predictions/s",
            "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": dev_ms / steps,
            "wall_ms_per_step": wall_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "predictions/s", "ms_per_step": e2e_ms / steps,
                    "h2d_bytes_per_step": int(n * 3 * 8), "d2h_bytes_per_step": int(M * 17)},
            "gpu_launches": launches,
            "phases_ms_per_step": {"bins": bin_ms / steps, "knn": knn_ms / steps, "job_index": index_ms / steps,
                                   "solve": (solve_ms - index_ms) / steps,
                                   "nccl_broadcast_once": bcast_ms},
            "roofline": {"bound": "fp64", "kernel": "local_krige_shr_kernel<1,false,4> (gather + shared-factor blocked Cholesky + solve)",
                         "achieved": solve_tflops, "peak": peak, "unit": "TFLOP/s", "frac": solve_tflops / peak,
                         "peak_source": "measured in this run (gsm_measure_fp64_peak: DFMA %.1f, DMMA m8n8k4 %.1f TFLOP/s; "
                                        "cuBLAS DGEMM 4096^3 %.1f TFLOP/s); MEASURED_PEAKS.json has no FP64 entry"
                                        % (fp64_dfma, fp64_dmma, fp64_cublas),
                         "flops_per_prediction": fpp, "launch_ms": solve_launch_ms,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch (2^20 systems) from the
                         # `ncu --set full` capture summarised in profiles/r01_knn_index_local_krige_shr_ncu_summary.txt;
                         # algorithmic: 2^17 job descriptors x 768 B + 2^20 x 17 B of outputs = 118.5 MB
                         "traffic": 115.9e6, "traffic_unit": "bytes per launch (ncu)",
                         "algorithmic_flops_per_launch": fpp * min(M, 1 << 20),
                         "bound_note": "FP64 pipe; on B200 the DMMA (tensor) and DFMA paths are one pipe "
                                       "(tools/pipebench.cu), so the peak is the larger of the two measured rates",
                         "whole_step_frac": fpp * M * world * steps / (dev_ms * 1e-3) / 1e12 / (peak * world),
                         "hbm": {"algorithmic_bytes_per_prediction": alg_bytes,
                                 "achieved_gbs": alg_bytes * M * steps / (dev_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        }
        if world == 1 and not args.no_cpu_baseline:
            o = entry.load_oracle()
            import gsm_oracle_c as oc
            threads = host_threads()
            Xc, zc = make_samples(1)
            cpu_reference_step(o, oc, Xc, zc, 1, 8, threads)
            ncpu, dt, passes = 0, 0.0, 0
            while dt < args.cpu_seconds and passes < 50:
                n1, d1 = cpu_reference_step(o, oc, Xc, zc, 1, args.cpu_rows, threads)
                ncpu += n1
                dt += d1
                passes += 1
            line["cpu_baseline"] = {"value": ncpu / dt, "unit": "predictions/s", "cores": threads, "kind": "port",
                                    "sample": f"{passes} passes over the first {args.cpu_rows} of {NY} grid rows "
                                              f"({ncpu} targets), KD-tree build included in every pass, {dt:.1f} s; "
                                              f"C restatement of the reference algorithm (oracle/gsm_oracle.c)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rows", type=int, default=2048, help="grid rows of the bounded CPU sample (per pass)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="repeat the CPU sample until this much time is spent")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
