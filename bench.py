#!/usr/bin/env python
"""Benchmark of the BAO-reconstruction hot path (BASELINE.json metric:
"ms per reconstruction (1024^3 mesh, 1e8 particles)").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference ...                     # BAOrec.jl's CPU path, restated (oracle)

One "step" = one reconstruction = run!(IterativeRecon, 1024^3, data) + read_shifts(:sum) on the
data catalog (BASELINE configs[3]; periodic box L=2500 Mpc/h, CIC, n_iter=3, R=15 Mpc/h,
los=(0,0,1), 1e8 synthetic uniform particles, seed 42).  `value` is timed with the catalog
already resident in HBM; `e2e` goes through the host-buffer C-ABI pipeline (pinned host
catalogs in, shifts out, copies inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ms per reconstruction (1024^3 mesh, 1e8 particles)"
UNIT = "ms"
PARAMS = dict(bias=2.2, f=0.757, smoothing_radius=15.0, n_iter=3, los=(0.0, 0.0, 1.0))
BOX_L = 2500.0


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_catalog(n, L, seed, pinned=False, share=None):
    """Uniform periodic catalog, float32 SoA, weights 1.  share = (lo, hi): only particles
    lo .. hi-1 of the SAME catalog are kept (every rank draws the whole sequence, so the catalog is identical at every
    rank count: rank r of P holds the r-th contiguous part of it, NOT the particles of its slab)."""
    import torch
    rng = np.random.default_rng(seed)
    lo, hi = share if share is not None else (0, n)
    arrs = []
    for _ in range(3):
        t = torch.empty(hi - lo, dtype=torch.float32, pin_memory=pinned)
        a = t.numpy()
        chunk = 1 << 24
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            c = rng.random(e - s, dtype=np.float32) * np.float32(L)
            a0, a1 = max(s, lo), min(e, hi)
            if a1 > a0:
                a[a0 - lo:a1 - lo] = c[a0 - s:a1 - s]
        np.minimum(a, np.nextafter(np.float32(L), np.float32(0)), out=a)
        arrs.append(t)
    w = torch.ones(hi - lo, dtype=torch.float32, pin_memory=pinned)
    return arrs, w


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes(name, M, Mc, N, chunks=1):
    """Compulsory HBM traffic per launch ('kernels'): every operand read once, every
    result written once.  M, Mc, N are this rank's cells / complex modes / particles; entries launched once per plane
    chunk of the slab transforms (2-D transforms, pack / transpose kernels) move 1/chunks of the slab per launch."""
    if name.startswith("cufft_2d") or name.startswith("rows_kernel") or name.startswith("transpose_kernel"):
        return (4 * M + 8 * Mc if name.startswith("cufft_2d") else 16 * Mc) // max(chunks, 1)
    if name.startswith("peer_") or name.startswith("nccl_"):
        return None
    if "usort_count" in name or "hash_positions" in name:
        return 12 * N                       # positions in
    if "usort_reorder" in name:
        return 16 * N + 16 * N + 4 * N      # x, y, z, w in; float4 records + inverse permutation out
    if "shard_count" in name:
        return 4 * N
    if "shard_reorder" in name:
        return 16 * N + 16 * N + 4 * N      # columns in, columns grouped by owner + map out
    if "unshard_gather" in name:
        return 4 * N + 12 * N + 12 * N
    if "scatter_sorted" in name or "scatter_direct" in name or "scatter_records" in name:
        return 16 * N + 4 * M              # records/SoA in, every mesh cell written once
    if "gather_sorted_kernel<3" in name or "gather_direct_kernel<3" in name:
        return 16 * N + 12 * N + 12 * M    # positions(+index) in, 3 outputs, 3 displacement meshes once
    if "gather_sorted_kernel<1" in name or "gather_direct_kernel<1" in name:
        return 16 * N + 4 * N + 4 * M
    if "bin_count" in name or "tile_count" in name:
        return 12 * N
    if "bin_reorder" in name:
        return 16 * N + 16 * N
    if "tile_reorder" in name:
        return 12 * N + 16 * N + 4 * N
    if "gather_tile_kernel<3" in name or "gather_tile_tma_kernel<3" in name:
        return 16 * N + 16 * N + 12 * M    # records in, float4 results out (sorted order), 3 meshes once
    if "unsort" in name:
        return 4 * N + 16 * N + 12 * N
    if "DispOp" in name:
        return 8 * Mc + 24 * Mc
    if "FusedLosOp" in name:
        return 24 * Mc                      # rho_k in, delta_k/M out for the C2R, delta_k kept for the read-back
    if name.startswith("kspace_kernel"):
        return 16 * Mc
    if name.startswith("axpy") or name.startswith("radial_update") or name.startswith("randoms_combine"):
        return 12 * M
    if name.startswith("cufft_1d_x"):
        return 4 * M + 8 * Mc
    if name.startswith("cufft_1d"):
        return 16 * Mc
    if name.startswith("fft_cols_kernel"):
        return 16 * Mc
    if name.startswith("fft_z_solve"):
        return 24 * Mc                      # column in, column out, delta_k kept for the read-back
    if name.startswith("fft_z_disp"):
        return 8 * Mc + 24 * Mc
    if name.startswith("cufft"):
        return 4 * M + 8 * Mc              # single-pass lower bound (a 3-axis-pass FFT moves ~3x this)
    if name.startswith("rows_kernel") or name.startswith("transpose_kernel"):
        return 16 * Mc
    if "DispCompOp" in name:
        return 16 * Mc
    return None


# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of THIS workload
# (1024^3, 1e8 particles, one GPU; profiles/r1_ncu_full_final_own_kernels.csv, r1_ncu_full_prof_r1_a.csv, r2_ncu_full_*.csv)
NCU_TRAFFIC_BYTES_C4 = {
    "gather_tile_kernel<3>": 14.502e9 + 1.598e9,
    "gather_tile_tma_kernel<3>": 14.502e9 + 1.598e9,  # profiles/r2_ncu_full_gather_tma.csv
    "fft_z_disp_kernel": 4.805e9 + 13.877e9,         # profiles/r2_ncu_full_own_kernels.csv
    "fft_z_solve_kernel": 4.330e9 + 8.576e9,
    "fft_cols_kernel": 4.303e9 + 4.251e9,
    "scatter_sorted_kernel": 5.851e9 + 4.100e9,      # z-slab order (option unified_sort=0)
    "scatter_records_kernel": 5.849e9 + 4.096e9,     # profiles/r1_ncu_full_unified_sort.csv
    "usort_reorder_kernel": 3.509e9 + 3.308e9,
    "tile_reorder_kernel": 3.090e9 + 3.300e9,
    "bin_reorder_kernel": 1.645e9 + 1.587e9,
    "unsort_kernel": 7.531e9 + 1.195e9,
}


def ncu_traffic(name, n, N, world):
    if n != 1024 or N != 100_000_000 or world != 1:
        return None
    for key, val in NCU_TRAFFIC_BYTES_C4.items():
        if key in name:
            return val
    return None


def cpu_oracle():
    """The CPU stand-in for BAOrec.jl's threaded CPU path: the oracle with compiled loops
    (oracle/baorec_oracle_c.c: serial scatter like the reference's, OpenMP gather / k-space loops,
    scipy.fft on all host threads) when oracle/_c/ was built, else the pure-numpy port."""
    # torchrun exports OMP_NUM_THREADS=1 to its workers, and scipy's ducc FFT sizes its thread pool from it (measured
    # here: the transforms 2.5x slower, the whole CPU arm 3.5x -- round 1's 100 s instead of 20 s under torchrun);
    # the CPU arm runs on rank 0 alone and is meant to use the host's cores
    ncpu = str(os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = ncpu
    os.environ["DUCC0_NUM_THREADS"] = ncpu
    assert "scipy.fft" not in sys.modules or os.environ.get("BAOREC_BENCH_ALLOW_SCIPY_PRELOAD"), "scipy.fft was imported before the thread count was set"
    sys.path.insert(0, str(ROOT / "oracle"))
    import baorec_oracle_fast as fast
    if fast.available():
        O = fast.load(threads=os.cpu_count() or 1)      # torchrun sets OMP_NUM_THREADS=1; this leg runs on rank 0 alone
        return O, (f"C/OpenMP port of the reference's CPU methods ({O.threads} threads; scatter serial like the "
                   "reference, src/mas.jl:5) + scipy.fft on all host threads")
    import baorec_oracle as O
    return O, ("numpy port of the reference's CPU methods: scipy.fft on all host threads, scatter serial like the "
               "reference (src/mas.jl:5), gather/k-space loops single-threaded numpy")


def cpu_reference(mesh_n, n_part, reps, warm):
    """BAOrec.jl's CPU path, restated (oracle/; NOT the Julia binary): run! + read_shifts(:sum) on a
    bounded sample: a (mesh_n)^3 sub-volume with the same cell size and particle density as the
    1024^3 workload.  Returns (seconds per sample reconstruction, description of the port)."""
    O, what = cpu_oracle()
    L = BOX_L * mesh_n / 1024.0
    rng = np.random.default_rng(42)
    pos = [(rng.random(n_part, dtype=np.float32) * np.float32(L)) for _ in range(3)]
    for p in pos:
        np.minimum(p, np.nextafter(np.float32(L), np.float32(0)), out=p)
    w = np.ones(n_part, np.float32)
    times = []
    for i in range(warm + reps):
        rec = O.IterativeRecon(box_size=np.full(3, L, np.float32), box_min=np.zeros(3, np.float32), **PARAMS)
        t0 = time.perf_counter()
        mesh = O.run(rec, (mesh_n,) * 3, *[p.copy() for p in pos], w)
        O.read_shifts(rec, *pos, mesh, "sum")
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times, what


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    mesh_n = args.ref_mesh
    scale = (1024 // mesh_n) ** 3
    n_part = int(args.particles) // scale
    times, what = cpu_reference(mesh_n, n_part, args.steps, args.warmup)
    ms_sample = 1e3 * float(np.mean(times))
    value = ms_sample * scale
    cores = os.cpu_count() or 1
    sample = (f"{mesh_n}^3 mesh / {n_part} particles sub-volume (same cell size and density), "
              f"{ms_sample:.0f} ms per sample reconstruction, scaled x{scale} (linear) to 1024^3 / 1e8; {what}")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_sample, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "IterativeRecon periodic box, 1024^3 mesh, 1e8 particles, CIC, n_iter=3, "
                               "R=15, los=(0,0,1) + read_shifts(:sum) -- CPU port on a bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh", type=int, default=1024)
    ap.add_argument("--particles", type=float, default=1e8)
    ap.add_argument("--ref-mesh", type=int, default=512, help="mesh size of the CPU sample (sub-volume of the workload)")
    ap.add_argument("--catalog", default="uniform", choices=["uniform", "lognormal"],
                    help="synthetic catalog (SURVEY.md 8d C4): uniform (the default workload) or lognormal "
                         "(sigma = 1 on a 512^3 generation mesh + linear RSD shift, generated on the device)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-batch", type=int, default=4,
                    help="K > 0 (one GPU): additionally time the batched host pipeline (baorec_batch_host_f32) over K catalogs "
                         "and report it as e2e['batched'] (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as G
    B = G.load_package()            # raises if libbaorec_b200.so is missing: no fallback

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = args.mesh
    N = int(args.particles)
    M = n ** 3
    Mc = (n // 2 + 1) * n * n
    L = BOX_L * n / 1024.0   # same 2.44 Mpc/h cells at every mesh size
    grid = (n, n, n)
    kw = dict(box_size=np.full(3, L, np.float32), box_min=np.zeros(3, np.float32), **PARAMS)

    ctx = B.Context.get(local_rank)
    rec = B.IterativeRecon(**kw)
    if world == 1 and args.catalog == "lognormal":
        sys.path.insert(0, str(ROOT / "benchmarks"))
        import catalogs
        dpos, dwt = catalogs.lognormal_box(N, L, seed=42, device=dev, n_gen=min(n, 512), sigma=1.0, f_rsd=PARAMS["f"])
        hx, hy, hz, hw = (torch.empty(N, dtype=torch.float32, pin_memory=True).copy_(t) for t in (*dpos, dwt))
        del dpos, dwt
        torch.cuda.empty_cache()
        n_loc = N
    elif world == 1:
        (hx, hy, hz), hw = make_catalog(N, L, seed=42, pinned=True)
        n_loc = N
    else:
        # strong scaling of the SAME 1e8-particle catalog: rank r holds the r-th contiguous part of it (not the
        # particles of its slab); sharding by slab, the reconstruction and the way back of the shifts are one library
        # call (baorec_reconstruct_dist_f32) inside the timed region
        B.dist.init_comm(ctx)
        if os.environ.get("BAOREC_A2A_CHUNKS"):      # A/B knobs
            ctx.set_option("a2a_chunks", int(os.environ["BAOREC_A2A_CHUNKS"]))
        if os.environ.get("BAOREC_PEER_HALO"):
            ctx.set_option("peer_halo", int(os.environ["BAOREC_PEER_HALO"]))
        if os.environ.get("BAOREC_PUSH_SM"):
            ctx.set_option("push_sm", int(os.environ["BAOREC_PUSH_SM"]))
        if os.environ.get("BAOREC_COMM_SPLIT"):
            ctx.set_option("comm_split", int(os.environ["BAOREC_COMM_SPLIT"]))
        B.dist.plan(ctx, grid, kw["box_size"], kw["box_min"], exchange=os.environ.get("BAOREC_EXCHANGE"))
        lo, hi = rank * N // world, (rank + 1) * N // world
        if args.catalog == "lognormal":
            # every rank draws the whole catalog on its own device (same seed, same generator: the same catalog) and
            # keeps its contiguous share
            sys.path.insert(0, str(ROOT / "benchmarks"))
            import catalogs
            dpos, dwt = catalogs.lognormal_box(N, L, seed=42, device=dev, n_gen=min(n, 512), sigma=1.0, f_rsd=PARAMS["f"])
            hx, hy, hz, hw = (torch.empty(hi - lo, dtype=torch.float32, pin_memory=True).copy_(t[lo:hi]) for t in (*dpos, dwt))
            del dpos, dwt
            torch.cuda.empty_cache()
        else:
            (hx, hy, hz), hw = make_catalog(N, L, seed=42, pinned=True, share=(lo, hi))
        n_loc = hi - lo
    dx, dy, dz, dw = (t.to(dev) for t in (hx, hy, hz, hw))

    # result buffers are allocated once (run! zero-fills the mesh on every call, as the reference
    # does); keeps the torch caching allocator out of the timed region
    out_buf = tuple(torch.empty_like(dx) for _ in range(3))
    if world == 1:
        mesh_buf = torch.empty((n, n, n), dtype=torch.float32, device=dev)

        def step_device():
            mesh = B.run(rec, grid, dx, dy, dz, dw, mesh_out=mesh_buf)
            return B.read_shifts(rec, dx, dy, dz, mesh, field="sum", out=out_buf)
    else:
        def step_device():
            return B.dist.reconstruct_dist(rec, grid, dx, dy, dz, dw, field="sum", ctx=ctx, out=out_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # two untimed settle steps (plans, scratch and the torch caching allocator reach steady state:
    # run! allocates its 4 GiB result mesh on every call, like the reference), then W warm-ups
    for _ in range(2 + args.warmup):
        step_device()
    barrier()
    k0, f0 = ctx.launch_counts()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.stop()
    k1, f1 = ctx.launch_counts()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    # mean |shift_z| over ALL particles of the catalog: comparable across rank counts
    cs = out[2].double().abs().sum().reshape(1)
    if world > 1:
        dist.all_reduce(cs)
    checksum = float(cs.item()) / N

    # ---- end-to-end through the host-buffer C-ABI pipeline (pinned host catalogs) ------------
    e2e = None
    if not args.no_e2e:
        rec_h = B.IterativeRecon(**kw)
        ax, ay, az, aw = (t.numpy() for t in (hx, hy, hz, hw))
        outs_t = [torch.empty(n_loc, dtype=torch.float32, pin_memory=True) for _ in range(3)]
        outs = [t.numpy() for t in outs_t]

        def step_host_dist():
            # baorec_reconstruct_dist_host_f32: pinned host share of the catalog -> device, sharding by slab,
            # distributed solve + read-back, shifts back in the caller's order -> pinned host
            B.dist.reconstruct_dist(rec_h, grid, ax, ay, az, aw, field="sum", ctx=ctx, out=outs)

        def step_host():
            if world > 1:
                return step_host_dist()
            B.run(rec_h, grid, ax, ay, az, aw)
            c = rec_h._ctx()
            p = rec_h._params()
            import ctypes as C
            B.lib_loader.check(c.lib.baorec_read_host_f32(
                c.handle, C.byref(p), rec_h.algorithm, None, ax.ctypes.data, ay.ctypes.data, az.ctypes.data, n_loc,
                B.lib_loader.FIELD_SUM, 1, outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data))

        del out
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3 / args.steps
        stage = ctx.stage_ms()
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cs = torch.tensor([float(np.abs(outs[2]).sum(dtype=np.float64))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(cs)
        e2e = {"value": float(t.item()), "unit": UNIT,
               "api": ("baorec_run_host_f32 + baorec_read_host_f32" if world == 1 else
                       "baorec_reconstruct_dist_host_f32 on every rank (its N/P share of the unsharded catalog)"),
               "h2d_bytes_per_step": (16 * N + 12 * N) if world == 1 else 16 * N,
               "d2h_bytes_per_step": 12 * N,
               "stage_ms": {"h2d_catalog": stage[0], "solve": stage[1], "d2h_run": stage[2],
                            "h2d_pos+disp_meshes": stage[3], "gather": stage[4], "d2h_shifts": stage[5]}
               if (len(stage) >= 6 and world == 1) else None,
               "checksum_abs_mean_shift_z": float(cs.item()) / N}

    # ---- optional: the batched host pipeline (transfers of neighbouring catalogs overlap the solve) -----------
    if e2e is not None and args.e2e_batch > 0 and world == 1:
        K = args.e2e_batch
        (gx, gy, gz), gw = make_catalog(N, L, seed=43, pinned=True)          # a second catalog: the batch alternates
        two = [(ax, ay, az, aw), tuple(t.numpy() for t in (gx, gy, gz, gw))]
        outs2_t = [torch.empty(n_loc, dtype=torch.float32, pin_memory=True) for _ in range(3)]
        two_out = [tuple(outs), tuple(t.numpy() for t in outs2_t)]
        cats = [two[i & 1] for i in range(K)]
        bouts = [two_out[i & 1] for i in range(K)]
        rec_b = B.IterativeRecon(**kw)
        B.run_batch(rec_b, grid, cats[:2], field="sum", positions=False, out=bouts[:2])     # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        B.run_batch(rec_b, grid, cats, field="sum", positions=False, out=bouts)
        torch.cuda.synchronize()
        dtb = (time.perf_counter() - t0) * 1e3 / K
        e2e["batched"] = {"value": dtb, "unit": UNIT, "catalogs": K,
                          "h2d_bytes_per_catalog": 16 * N, "d2h_bytes_per_catalog": 12 * N,
                          "what": "baorec_batch_host_f32: run! + read_shifts(:sum) per catalog, upload of catalog i+1 and "
                                  "download of catalog i-1 overlapping the reconstruction of catalog i",
                          "checksum_abs_mean_shift_z": float(np.abs(bouts[-1][2][: 1 << 20]).mean())}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant hand-written kernel + per-kernel table ---------------------
    peak, peak_src = measured_peak_gbs()
    kernels = {}
    chunks = 1
    if world > 1:      # launches of the 2-D transforms per slab transform = plane chunks of the pipelined exchange
        c2 = sum(c for k, (_, c) in prof.items() if k.startswith("cufft_2d"))
        # z passes of the slab transforms: cuFFT's strided plan, or the column kernel (at N > 1 it runs nowhere else)
        c1 = sum(c for k, (_, c) in prof.items() if k.startswith("cufft_1d_z") or k.startswith("fft_cols_kernel"))
        chunks = max(1, round(c2 / max(c1, 1)))
    for name, (ms, cnt) in prof.items():
        ab = algorithmic_bytes(name, M // world, Mc // world, N // world, chunks)   # per-rank (slab) bytes
        per = ms / max(cnt, 1)
        kernels[name] = {"ms_per_launch": round(per, 4), "launches_per_step": cnt / args.steps,
                         "ms_per_step": round(ms / args.steps, 3),
                         "alg_GBs": round(ab / per / 1e6, 1) if ab else None,
                         "frac_of_peak": round(ab / per / 1e6 / peak, 3) if ab else None}
    own = {k: v for k, v in kernels.items() if not k.startswith(("cufft", "peer_", "nccl_"))}
    top = max(own, key=lambda k: own[k]["ms_per_step"]) if own else None
    roofline = None
    if top and kernels[top]["alg_GBs"]:
        roofline = {"kernel": top, "bound": "hbm", "achieved": kernels[top]["alg_GBs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[top]["frac_of_peak"], "traffic": ncu_traffic(top, n, N, world),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algorithmic_bytes(top, M // world, Mc // world, N // world, chunks)}
    exchange = None
    if world > 1:
        # The exchange of the slab transforms: every rank sends (P-1)/P of its complex slab per transform.  At N > 1
        # this, not an HBM kernel, is what bounds the step, so it is the line's roofline (rank 0's CUDA-event time of
        # the P-1 remote copies per chunk on the communication stream; the local 1/P block is an HBM copy on a second
        # stream and is not in the span).
        name = "peer_copies" if "peer_copies" in prof else "nccl_all_to_all"
        if name in prof:
            ms, cnt = prof[name]
            transforms = sum(c for k, (_, c) in prof.items() if k.startswith("cufft_1d_z") or k.startswith("fft_cols_kernel"))
            sent = 8 * (Mc // world) * (world - 1) / world * transforms          # bytes over NVLink, this rank, timed region
            nv_peak, nv_src = 770.0, "measured peer copy per direction (B200_PROFILING.md; nominal 900)"
            exchange = {"scheme": ("peer copies + sequence flags (remote blocks: %s)" % ("one SM kernel per chunk" if world >= 5 and not os.environ.get("BAOREC_PUSH_SM") == "0" else "copy engines"))
                        if name == "peer_copies" else "grouped ncclSend/ncclRecv",
                        "launches_per_step": cnt / args.steps, "ms_per_step": round(ms / args.steps, 3),
                        "nvlink_bytes_per_rank_per_step": int(sent / args.steps), "GBs_per_direction": round(sent / ms / 1e6, 1)}
            hbm_roofline = roofline
            roofline = {"kernel": name + " (slab-transform exchange, %d per step)" % round(cnt / args.steps), "bound": "nvlink",
                        "achieved": exchange["GBs_per_direction"], "peak": nv_peak, "unit": "GB/s",
                        "frac": round(exchange["GBs_per_direction"] / nv_peak, 3), "traffic": None, "peak_source": nv_src,
                        "algorithmic_bytes_per_launch": int(sent / max(cnt, 1)), "top_hbm_kernel": hbm_roofline}
    fft_ms = sum(v["ms_per_step"] for k, v in kernels.items() if k.startswith("cufft"))
    own_ms = sum(v["ms_per_step"] for k, v in own.items())

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        mesh_n = min(args.ref_mesh, n)
        scale = (n // mesh_n) ** 3
        times, what = cpu_reference(mesh_n, max(1, N // scale), reps=2, warm=1)
        ms_sample = 1e3 * float(np.mean(times))
        cpu_baseline = {"value": ms_sample * scale, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"{mesh_n}^3 mesh / {N // scale} particles sub-volume (same cell size and density): "
                                  f"{ms_sample:.0f} ms per sample reconstruction, scaled x{scale}; {what}"}

    out = {
        "metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
        "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"IterativeRecon periodic box, {n}^3 mesh, {N:.0e} particles ({args.catalog}, seed 42), CIC, "
                               f"n_iter=3, R=15 Mpc/h, L={L:g} Mpc/h, los=(0,0,1): run! + read_shifts(:sum)"
                               + ("" if world == 1 else f"; the same catalog at every rank count, rank r holds its r-th 1/{world} "
                                  "(unsharded): sharding by slab and the way back of the shifts are inside the step"),
                   "l2_policy": "inputs larger than L2 (4 GiB meshes, 1.6 GB catalog vs 126 MB L2)",
                   "ffts_per_step": (f1 - f0) / args.steps},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(k1 - k0),
        "cufft_execs": int(f1 - f0),
        "roofline": roofline, "exchange": exchange, "cpu_baseline": cpu_baseline,
        "breakdown_ms_per_step": {"own_kernels": round(own_ms, 3), "cufft": round(fft_ms, 3)},
        "kernels": kernels, "checksum_abs_mean_shift_z": checksum,
        "scratch_GiB": round(ctx.scratch_bytes() / 2 ** 30, 2),
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
