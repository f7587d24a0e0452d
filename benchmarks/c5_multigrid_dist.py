#!/usr/bin/env python
"""BASELINE.json configs[4]: MultigridRecon periodic box on a slab-decomposed mesh over all ranks
(default 2048^3 mesh, 1e9 particles, 8 x B200): halo-plane exchange per sweep, particles sharded by
slab with the ghost-plane exchange of the CIC stencil, slab FFT for the smoothing and the read-back.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 benchmarks/c5_multigrid_dist.py [--mesh 2048] [--particles 1e9]

Every rank draws its N/P particles inside its own slab on the device (per-rank seed); one step =
run! (scatter, set-up, full multigrid) + read_shifts(:sum).  Timed with CUDA events, max over
ranks; rank 0 prints one JSON line with the per-kernel breakdown of its own slab."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=2048)
    ap.add_argument("--particles", type=float, default=1e9)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--algorithm", default="multigrid", choices=["multigrid", "iterative"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = G.load_package()
    ctx = B.Context.get(local)
    B.dist.init_comm(ctx)
    n, N = args.mesh, int(args.particles)
    L = 2500.0 * n / 1024.0            # 2.44 Mpc/h cells as in the 1024^3 workload
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    B.dist.plan(ctx, (n, n, n), bs, bm)
    z_lo, nzl = B.dist.slab_range(ctx)
    cell = L / n
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_loc = N // world
    x = torch.rand(n_loc, device=dev, generator=gen) * L
    y = torch.rand(n_loc, device=dev, generator=gen) * L
    z = torch.rand(n_loc, device=dev, generator=gen) * (nzl * cell) + z_lo * cell
    x.clamp_(0, float(np.nextafter(np.float32(L), np.float32(0))))
    y.clamp_(0, float(np.nextafter(np.float32(L), np.float32(0))))
    own = B.dist.slab_owner(ctx, z)
    z[own != rank] = (z_lo + 0.5 * nzl) * cell          # particles within an ulp of the slab faces
    w = torch.ones(n_loc, device=dev)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=bs, box_min=bm, los=(0.0, 0.0, 1.0))
    rec = B.MultigridRecon(**kw) if args.algorithm == "multigrid" else B.IterativeRecon(n_iter=3, **kw)

    def step():
        B.dist.run_dist(rec, (n, n, n), x, y, z, w, ctx=ctx)
        return B.dist.read_shifts_dist(rec, x, y, z, field="sum")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    prof = ctx.profile_read()
    ctx.profile(False)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    chk = out[2].double().abs().mean()
    if world > 1:
        dist.all_reduce(chk)
    if rank == 0:
        kern = {k: {"ms_per_step": round(ms / args.steps, 3), "launches_per_step": c / args.steps}
                for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        cells_loc = n * n * nzl
        sweeps = {k: v for k, v in kern.items() if "mg_stencil_smem" in k}
        for k, v in sweeps.items():
            v["note"] = "all levels; the finest level moves 12 B/cell"
        print(json.dumps({"workload": f"{type(rec).__name__} periodic box, {n}^3 mesh, {N:.0e} particles, "
                                      f"{world} ranks (z slabs of {nzl} planes), CIC, los=(0,0,1): run! + read_shifts(:sum)",
                          "ms_per_reconstruction": float(t.item()), "n_gpus": world, "steps": args.steps,
                          "cells_per_rank": cells_loc, "scratch_GiB_rank0": round(ctx.scratch_bytes() / 2 ** 30, 2),
                          "checksum_mean_abs_shift_z": float(chk.item()) / world, "kernels_rank0": kern}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
