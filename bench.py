#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native interpolation hot path.

Metric (BASELINE.json): FP64 Kriging predictions/sec, local OrdinaryKriging, k = 32, at 1/2/4/8 B200.
Workload = the north-star configuration of BASELINE.json: 1 000 000 random 3-D samples -> 256 x 256 x 256
CartesianGrid (16.8 M targets), SphericalVariogram(range=20, sill=1, nugget=0.1), maxneighbors=32, mean + variance.
The domain is FIXED: with N GPUs rank r predicts slab r (256 / N planes) of the same grid -- strong scaling -- and the
samples are replicated from rank 0 by an NCCL broadcast.
A "step" is one full fitpredict over the domain: grid-bin build + kNN + assemble / factor / solve for every target.

  value : whole-job predictions/s, samples already resident in HBM on every rank, outputs left in HBM
          (per-phase CUDA events on the library's stream; max over ranks)
  e2e   : the same metric through the public API with HOST buffers: every step uploads the samples from pinned host
          memory on rank 0, replicates them (NCCL broadcast, N > 1), scales + bins them on every rank, predicts the
          slab and downloads mean / variance / status into pinned host memory (wall clock, barrier on both sides)
  configs : (N = 1) the other BASELINE configs, each timed through the public API: C1, C2, C3, C4, k = 64
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference algorithm's CPU restatement (oracle/gsm_oracle.c, all host threads) on a
bounded sample of the same workload; Julia is not installed, so that is the only form in which the reference path
can run on this box (cpu_baseline.kind = "port").
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
# workload: the north-star configuration (BASELINE.json north_star)
# ------------------------------------------------------------------------------------------------
G3 = 256
NS_SAMPLES = 1_000_000
K = 32
VARIO = dict(kind="spherical", sill=1.0, nugget=0.1, range=20.0)
SEED = 6
METRIC = "FP64 kriging predictions/sec (local OK, k=32)"


def make_samples():
    rng = np.random.default_rng(SEED)
    X = rng.uniform(0, G3, size=(NS_SAMPLES, 3))
    z = np.sin(X[:, 0] / 20.0) + np.cos(X[:, 2] / 30.0) + 0.1 * rng.standard_normal(NS_SAMPLES)
    return X, z


def flops_per_prediction(k=K, ncon=1, dim=3, c_gamma=8, ndrift=0):
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


def workload_config(n_gpus):
    return {"workload": "north-star: local OrdinaryKriging SphericalVariogram(range=20,sill=1,nugget=0.1) maxneighbors=32, "
                        f"{NS_SAMPLES} 3-D samples -> {G3}^3 grid, mean+variance",
            "samples": NS_SAMPLES, "targets": G3 ** 3, "k": K,
            "parallelism": f"fixed domain, {n_gpus} target slab(s) of {G3 // max(n_gpus, 1)} planes, samples replicated "
                           "(NCCL broadcast from rank 0, inside the e2e timed region)",
            "l2": "256 MiB L2 flush between timed steps (and the 402 MB of outputs + neighbour scratch per step exceed L2)"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in C (oracle/), bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(o, oc, X, z, planes, threads):
    """Predict `planes` z-planes from the middle of the north-star domain with the C oracle (KD-tree kNN +
    Bunch-Kaufman), KD-tree build included (the reference builds its search tree in every fitpredict)."""
    model = o.ordinary_kriging(o.Variogram(o.SPHERICAL, VARIO["sill"], VARIO["nugget"], VARIO["range"]))
    grid = o.Grid((0.0,) * 3, (1.0,) * 3, (G3, G3, G3))
    first = (G3 // 2) * G3 * G3
    t0 = time.perf_counter()
    mu, var, st, _, _ = oc.fitpredict(model, X, z, grid, prob=True, maxneighbors=K, first=first, count=planes * G3 * G3,
                                      use_tree=True, nthreads=threads)
    dt = time.perf_counter() - t0
    assert np.isfinite(mu).all()
    return planes * G3 * G3, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    o = entry.load_oracle()
    import gsm_oracle_c as oc
    threads = host_threads()
    X, z = make_samples()
    planes = args.cpu_planes
    for _ in range(min(max(args.warmup, 0), 1)):
        cpu_reference_step(o, oc, X, z, 1, threads)
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = cpu_reference_step(o, oc, X, z, planes, threads)
        tot_n += n
        tot_t += dt
    val = tot_n / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": val,
        "unit": "predictions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "predictions/s", "cores": threads, "kind": "port",
                         "sample": f"{planes} of {G3} grid planes ({planes * G3 * G3} targets) per step, "
                                   f"KD-tree build over the 1 M samples included; C restatement of the reference algorithm "
                                   f"(Julia is not installed on this image)"},
        "e2e": {"value": val, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs (N = 1 sub-records), each through the public API with pinned host buffers
# ------------------------------------------------------------------------------------------------
def other_configs(gsm, ctx, peak_tflops, reps=3):
    out = {}

    def timed(fn):
        fn()
        ts, tim = [], None
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
            tim = ctx.timings()
        return min(ts), float(np.mean(ts)), tim

    def pinned(X, z):
        Xp, zp = gsm.PinnedArray(X.shape, np.float64), gsm.PinnedArray(z.shape, np.float64)
        Xp.array[:] = X
        zp.array[:] = z
        return Xp, zp

    def record(name, M, best, mean, tim, fpp=None, **extra):
        r = {"targets": M, "e2e_ms": best * 1e3, "e2e_ms_mean": mean * 1e3, "predictions_per_s": M / best,
             "device_ms": {k: tim[k] for k in ("knn_ms", "index_ms", "solve_ms", "total_ms")}, **extra}
        if fpp:
            kern_ms = tim["solve_ms"] - tim["index_ms"]
            r["solve_kernel_tflops"] = fpp * M / (kern_ms * 1e-3) / 1e12
            r["solve_kernel_frac_of_fp64_peak"] = r["solve_kernel_tflops"] / peak_tflops
            r["flops_per_prediction"] = fpp
        out[name] = r

    # the reference's own (unreported) micro-benchmarks, benchmark/benchmarks.jl:25-26: OK GaussianVariogram(range=35),
    # 3 samples -> 100 x 100 grid, neighbors=true, without and with a missing value
    coords = np.array([[25.0, 25.0], [50.0, 75.0], [75.0, 50.0]])
    okb = gsm.OrdinaryKriging(gsm.GaussianVariogram(range=35.0))
    gridb = gsm.CartesianGrid(100, 100)
    for nm, zz in (("kriging/full (benchmarks.jl:25)", [1.0, 2.0, 3.0]), ("kriging/miss (benchmarks.jl:26)", [np.nan, 2.0, 3.0])):
        tabb = gsm.georef({"z": np.array(zz)}, gsm.PointSet(coords))
        b, m, t = timed(lambda: gsm.fitpredict(okb, tabb, gridb, neighbors=True, ctx=ctx))
        record(nm, 100 * 100, b, m, t)
    # C1: IDW (exponent 2), 10 k 2-D samples -> 256 x 256, all samples
    rng = np.random.default_rng(1)
    X = rng.uniform(0, 256, size=(10_000, 2))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(10_000)
    Xp, zp = pinned(X, z)
    tab, grid = gsm.georef({"z": zp.array}, gsm.PointSet(Xp.array)), gsm.CartesianGrid(256, 256)
    b, m, t = timed(lambda: gsm.fitpredict(gsm.IDW(2), tab, grid, neighbors=False, ctx=ctx))
    record("C1 IDW e=2 global 10k -> 256^2", 256 * 256, b, m, t, fpp=10_000 * 13)
    # C2: OrdinaryKriging GaussianVariogram, global fit, 4 k samples -> 1024 x 1024, mean + variance
    rng = np.random.default_rng(2)
    X = rng.uniform(0, 1024, size=(4000, 2))
    z = np.sin(X[:, 0] / 80) + np.cos(X[:, 1] / 120) + 0.1 * rng.standard_normal(4000)
    Xp, zp = pinned(X, z)
    tab, grid = gsm.georef({"z": zp.array}, gsm.PointSet(Xp.array)), gsm.CartesianGrid(1024, 1024)
    okg = gsm.OrdinaryKriging(gsm.GaussianVariogram(range=100.0, sill=1.0, nugget=0.01))
    b, m, t = timed(lambda: gsm.fitpredict(okg, tab, grid, neighbors=False, prob=True, ctx=ctx))
    record("C2 OK Gaussian global 4k -> 1024^2 (fit + predict)", 1024 * 1024, b, m, t, fpp=4000 * 13 + 4001.0 ** 2,
           note="e2e includes the one-off fit; solve_ms is the prediction kernels (triangular form: N^2 flops per target)")
    # C3: local OK k = 32, 200 k 2-D samples -> 2048 x 2048 (the round-1 headline workload)
    rng = np.random.default_rng(3)
    X = rng.uniform(0, 2048, size=(200_000, 2))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 1] / 30) + 0.1 * rng.standard_normal(200_000)
    Xp, zp = pinned(X, z)
    tab, grid = gsm.georef({"z": zp.array}, gsm.PointSet(Xp.array)), gsm.CartesianGrid(2048, 2048)
    ok50 = gsm.OrdinaryKriging(gsm.SphericalVariogram(range=50.0, sill=1.0, nugget=0.1))
    b, m, t = timed(lambda: gsm.fitpredict(ok50, tab, grid, maxneighbors=32, prob=True, ctx=ctx))
    record("C3 local OK k=32 200k -> 2048^2", 2048 * 2048, b, m, t, fpp=flops_per_prediction(dim=2))
    # the same with volume support (point=false): right-hand sides averaged over 3 x 3 points per cell
    b, m, t = timed(lambda: gsm.fitpredict(ok50, tab, grid, maxneighbors=32, prob=True, point=False, ctx=ctx))
    record("C3 with change of support (point=false, 9 points per cell)", 2048 * 2048, b, m, t,
           note="general per-system kernel; 9 x 32 covariance evaluations per right-hand side")
    # k = 64 (C5 column): SimpleKriging, 1 M 2-D samples -> 2048 x 2048
    rng = np.random.default_rng(5)
    X = rng.uniform(0, 2048, size=(1_000_000, 2))
    z = rng.standard_normal(1_000_000)
    Xp, zp = pinned(X, z)
    tab = gsm.georef({"z": zp.array}, gsm.PointSet(Xp.array))
    sk = gsm.SimpleKriging(gsm.SphericalVariogram(range=2048 / 20, nugget=0.1), float(z.mean()))
    b, m, t = timed(lambda: gsm.fitpredict(sk, tab, grid, maxneighbors=64, prob=True, ctx=ctx))
    record("C5 SK k=64 1M 2-D -> 2048^2", 2048 * 2048, b, m, t, fpp=flops_per_prediction(k=64, ncon=0, dim=2))
    # C4: UniversalKriging, quadratic drift (10 monomials), 1 M 3-D samples -> 256^3
    rng = np.random.default_rng(4)
    X = rng.uniform(0, 256, size=(1_000_000, 3))
    z = np.sin(X[:, 0] / 20) + np.cos(X[:, 2] / 30) + 0.1 * rng.standard_normal(1_000_000)
    Xp, zp = pinned(X, z)
    tab, grid = gsm.georef({"z": zp.array}, gsm.PointSet(Xp.array)), gsm.CartesianGrid(256, 256, 256)
    uk = gsm.UniversalKriging(gsm.SphericalVariogram(range=20.0, sill=1.0, nugget=0.1), 2, 3)
    b, m, t = timed(lambda: gsm.fitpredict(uk, tab, grid, maxneighbors=32, prob=True, ctx=ctx))
    record("C4 UK quadratic k=32 1M 3-D -> 256^3", 256 ** 3, b, m, t, fpp=flops_per_prediction(ncon=10, dim=3, ndrift=10))
    return out


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
    gsm.set_default_device(local_rank)
    ctx = gsm.default_context(local_rank)
    Mtot = G3 ** 3
    n = NS_SAMPLES
    grid = gsm.CartesianGrid(G3, G3, G3)
    first, count = gsm.distributed.shard(Mtot, rank, world, gsm.distributed.grid_unit(grid.dims))

    # ---- samples: generated on rank 0 into pinned host memory -------------------------------------------------
    alpha = 1.0 / max(VARIO["range"], float(G3))    # _scale, models.jl:272-284 (bounding box of the domain: [0, 256]^3)
    table = None
    if rank == 0:
        X, z = make_samples()
        assert alpha == 1.0 / max(VARIO["range"], np.abs(X).max(), float(G3))
        Xh, zh = gsm.PinnedArray((n, 3), np.float64), gsm.PinnedArray((n,), np.float64)
        Xh.array[:] = X
        zh.array[:] = z
        table = gsm.georef({"z": zh.array}, gsm.PointSet(Xh.array))   # contiguous float64: wrapped, not copied
    # replicate once for the device-resident measurement; the e2e steps replicate again, every step
    if world > 1:
        dist.barrier()
    coords_d, values_d = gsm.distributed.replicate_samples(Xh.array if rank == 0 else None, zh.array if rank == 0 else None,
                                                           src=0, device=dev)
    coords_d.mul_(alpha)
    torch.cuda.synchronize()
    # one more, timed: the steady-state cost of the broadcast alone (communicator warm, buffers allocated)
    bcast_ms = 0.0
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        dist.broadcast(coords_d, src=0)
        dist.broadcast(values_d, src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)

    vario = L.make_variogram(L.GSM_VARIO_SPHERICAL, VARIO["sill"], VARIO["nugget"], alpha * VARIO["range"])
    model = L.make_model(L.GSM_KRIG_ORDINARY)
    tgt = gsm.Context.grid_targets((0.0,) * 3, (alpha,) * 3, (G3, G3, G3), first=first, count=count)
    out_mean = torch.empty(count, dtype=torch.float64, device=dev)
    out_var = torch.empty(count, dtype=torch.float64, device=dev)
    out_stat = torch.empty(count, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        """inputs resident in HBM -> outputs in HBM; returns per-phase device ms of this step"""
        ctx.set_samples(coords_d, values_d, dim=3, n=n)
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

    # ---- e2e: public API, pinned host buffers; upload + broadcast + bins + predict + download per step -----
    pm, pv, ps = gsm.PinnedArray((count,), np.float64), gsm.PinnedArray((count,), np.float64), gsm.PinnedArray((count,), np.uint8)
    okm = gsm.OrdinaryKriging(gsm.SphericalVariogram(sill=VARIO["sill"], range=VARIO["range"], nugget=VARIO["nugget"]))
    outs = {"mean": pm.array, "var": pv.array, "status": ps.array}

    def e2e_step():
        if world == 1:
            return gsm.fitpredict(okm, table, grid, maxneighbors=K, prob=True, ctx=ctx, out=outs)
        return gsm.distributed.fitpredict_sharded(okm, table, grid, src=0, ctx=ctx, out=outs, maxneighbors=K, prob=True)[0]

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for _ in range(args.steps):
        e2e_step()
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
        value = Mtot * steps / (dev_ms * 1e-3)
        e2e_val = Mtot * steps / (e2e_ms * 1e-3)
        fpp = flops_per_prediction()
        # dominant kernel = the shared-factor solve kernel; the job-index kernel that feeds it is timed apart
        kern_ms = solve_ms - index_ms
        chunks = max(1, -(-count // (1 << 20)))
        solve_launch_ms = kern_ms / (steps * chunks)
        solve_tflops = fpp * count * steps / (kern_ms * 1e-3) / 1e12
        peak = max(fp64_dfma, fp64_dmma, fp64_cublas)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # DRAM traffic of one launch of the dominant kernel, from the committed ncu capture of this command
        # (profiles/r02_solve_kernel_traffic.json, written by tools/ncu_traffic.py); null when that file is absent
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_solve_kernel_traffic.json")))
        except Exception:
            pass
        alg_bytes = 16 + 1 + K * 4 * 2   # mean+var+status out, neighbour list written by kNN and read by the index kernel
        line = {
            "metric": METRIC, "value": value, "unit": "predictions/s",
            "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": dev_ms / steps,
            "wall_ms_per_step": wall_ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "predictions/s", "ms_per_step": e2e_ms / steps,
                    "h2d_bytes_per_step": int(n * 4 * 8), "d2h_bytes_per_step": int(Mtot * 17),
                    "note": "rank 0 uploads the samples from pinned host memory every step; N > 1: + NCCL broadcast to "
                            "every rank; every rank downloads its slab into pinned host memory"},
            "gpu_launches": launches,
            "phases_ms_per_step": {"bins": bin_ms / steps, "knn": knn_ms / steps, "job_index": index_ms / steps,
                                   "solve": (solve_ms - index_ms) / steps,
                                   "nccl_broadcast_steady_state": bcast_ms},
            "roofline": {"bound": "fp64", "kernel": "local_krige_shr_kernel<1,true,4> (gather + shared-factor blocked Cholesky + solve)",
                         "achieved": solve_tflops, "peak": peak, "unit": "TFLOP/s", "frac": solve_tflops / peak,
                         "peak_source": "measured in this run (gsm_measure_fp64_peak: DFMA %.1f, DMMA m8n8k4 %.1f TFLOP/s; "
                                        "cuBLAS DGEMM 4096^3 %.1f TFLOP/s); MEASURED_PEAKS.json has no FP64 entry"
                                        % (fp64_dfma, fp64_dmma, fp64_cublas),
                         "flops_per_prediction": fpp, "launch_ms": solve_launch_ms,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic.get("source") if traffic else None,
                         "algorithmic_flops_per_launch": fpp * min(count, 1 << 20),
                         "bound_note": "FP64 pipe; on B200 the DMMA (tensor) and DFMA paths are one pipe "
                                       "(tools/pipebench.cu), so the peak is the larger of the two measured rates",
                         "whole_step_frac": fpp * Mtot * steps / (dev_ms * 1e-3) / 1e12 / (peak * world),
                         "hbm": {"algorithmic_bytes_per_prediction": alg_bytes,
                                 "achieved_gbs": alg_bytes * Mtot * steps / (dev_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        }
        if world == 1 and not args.no_configs:
            try:
                line["configs"] = other_configs(gsm, ctx, peak)
            except Exception as ex:   # a sub-record must never take the headline down
                line["configs"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            o = entry.load_oracle()
            import gsm_oracle_c as oc
            threads = host_threads()
            Xc, zc = make_samples()
            ncpu, dt, passes = 0, 0.0, 0
            while dt < args.cpu_seconds and passes < 50:
                n1, d1 = cpu_reference_step(o, oc, Xc, zc, args.cpu_planes, threads)
                ncpu += n1
                dt += d1
                passes += 1
            line["cpu_baseline"] = {"value": ncpu / dt, "unit": "predictions/s", "cores": threads, "kind": "port",
                                    "sample": f"{passes} passes over {args.cpu_planes} of {G3} grid planes "
                                              f"({ncpu} targets), KD-tree build over the 1 M samples included in every pass, {dt:.1f} s; "
                                              f"C restatement of the reference algorithm (oracle/gsm_oracle.c)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-planes", type=int, default=2, help="grid planes of the bounded CPU sample (per pass / step)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="repeat the CPU sample until this much time is spent")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-records of the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
