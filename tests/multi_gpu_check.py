"""Multi-rank parity check of the slab-decomposed path (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank builds the same seeded catalog, keeps the particles of its own slab, runs the
distributed reconstruction and compares its mesh slab and its particles' shifts with the CPU
oracle (tolerances of BASELINE.json: rel. rms <= 1e-4, max |ds| <= 1e-3 Mpc/h)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))

import __graft_entry__ as G  # noqa: E402
import baorec_oracle as O  # noqa: E402
from util import clustered_box, rel_rms, maxabs  # noqa: E402


def more_modes(B, ctx, rank, world, quick=False):
    """MultigridRecon on slabs (halo exchange per sweep, slab restriction / prolongation, all-gathered
    coarse levels), radial line of sight and the randoms set-up, all against the oracle."""
    ok = True
    n, L = 64, 1000.0
    if n % (2 * world):
        return ok

    def put(a, mine):
        return torch.from_numpy(np.ascontiguousarray(a[mine])).cuda()

    def report(tag, e_mesh, e_rms, e_max, tol_max=1e-3):
        good = e_mesh < 1e-4 and e_rms < 1e-4 and e_max < tol_max
        print(f"[rank {rank}/{world}] {tag}: mesh rel.rms={e_mesh:.2e} shift rel.rms={e_rms:.2e} "
              f"max={e_max:.2e} {'OK' if good else 'FAIL'}", flush=True)
        return good

    # 1. MultigridRecon, periodic box, fixed and radial LOS; all levels on slabs / coarse levels replicated
    for los, lo, min_cells in (((0.0, 0.0, 1.0), 0.0, 0), (None, 700.0, 16 ** 3), ((0.0, 1.0, 0.0), 0.0, 1 << 22)):
        pos, w = clustered_box(200_000, L, seed=21, lo=lo)
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
                  box_min=np.full(3, lo, np.float32), los=los)
        orec = O.MultigridRecon(**kw)
        ophi = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
        oshift = O.read_shifts(orec, *pos, ophi, "sum")
        mine = B.dist.owner_of_z(pos[2], lo, L, n, world) == rank
        d = [put(p, mine) for p in pos]
        ctx.set_option("mg_slab_min_cells", min_cells)
        rec = B.MultigridRecon(**kw)
        phi = B.dist.run_dist(rec, (n, n, n), *d, put(w, mine), ctx=ctx)
        z_lo, nzl = B.dist.slab_range(ctx)
        s = B.dist.read_shifts_dist(rec, *d, field="sum")
        ok &= report(f"multigrid box los={los} min_cells={min_cells}",
                     rel_rms(phi.cpu().numpy(), ophi[z_lo:z_lo + nzl]),
                     max(rel_rms(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3)),
                     max(maxabs(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3)))
    ctx.set_option("mg_slab_min_cells", 16 ** 3)
    if quick:
        ctx.set_option("mg_slab_min_cells", 1 << 22)
        return ok

    # 2. randoms set-up (box filled by the randoms: no cell on the ran > threshold discontinuity)
    lo = 700.0
    span = L * (1.0 - 2.0 / n)
    dd, wd = clustered_box(60_000, span, seed=31, lo=lo)
    rr, wr = clustered_box(600_000, span, seed=32, lo=lo, nclump=1, sigma=10.0)
    md = B.dist.owner_of_z(dd[2], lo, L, n, world) == rank
    mr = B.dist.owner_of_z(rr[2], lo, L, n, world) == rank
    for algo in ("iterative", "multigrid"):
        for los in (None, (0.0, 0.0, 1.0)):
            kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
                      box_min=np.full(3, lo, np.float32), los=los)
            if algo == "iterative":
                orec, rec = O.IterativeRecon(**kw), B.IterativeRecon(**kw)
                omesh = O.reconstructed_overdensity(np.zeros((n, n, n), np.float32), orec, *dd, wd, *rr, wr)
            else:
                orec, rec = O.MultigridRecon(**kw), B.MultigridRecon(**kw)
                omesh = O.reconstructed_potential(np.zeros((n, n, n), np.float32), orec, *dd, wd, *rr, wr)
            oshift = O.read_shifts(orec, *dd, omesh, "sum")
            # yardstick for max |ds| (as in tests/test_gpu_multigrid.py): 1 / (alpha ran) amplifies Float32 rounding in the
            # sparse cells and the multigrid solve spreads it, so the Float32 oracle's own distance to the Float64
            # oracle bounds what any Float32 summation order can promise (first 2-rank run: 0.6 - 1.1e-3 Mpc/h here)
            d64, r64 = [q.astype(np.float64) for q in dd], [q.astype(np.float64) for q in rr]
            o64 = (O.IterativeRecon if algo == "iterative" else O.MultigridRecon)(**kw)
            run64 = O.reconstructed_overdensity if algo == "iterative" else O.reconstructed_potential
            m64 = run64(np.zeros((n, n, n), np.float64), o64, *d64, wd.astype(np.float64), *r64, wr.astype(np.float64))
            s64 = O.read_shifts(o64, *d64, m64, "sum")
            tol_max = max(1e-3, 3 * max(maxabs(oshift[a], s64[a]) for a in range(3)))
            gd = [put(p, md) for p in dd]
            mesh = B.dist.run_dist(rec, (n, n, n), *gd, put(wd, md), *[put(p, mr) for p in rr], put(wr, mr), ctx=ctx)
            z_lo, nzl = B.dist.slab_range(ctx)
            g, o = mesh.cpu().numpy().astype(np.float64), omesh[z_lo:z_lo + nzl].astype(np.float64)
            if algo == "multigrid":   # potential: defined up to a constant (global mean)
                gm = torch.tensor([g.sum()], dtype=torch.float64, device="cuda")
                dist.all_reduce(gm)
                g, o = g - gm.item() / n ** 3, o - omesh.astype(np.float64).mean()
                e_mesh = float(np.sqrt(np.mean((g - o) ** 2)) / (omesh.astype(np.float64) - omesh.astype(np.float64).mean()).std())
            else:
                e_mesh = rel_rms(g, o)
            s = B.dist.read_shifts_dist(rec, *gd, field="sum")
            ok &= report(f"{algo} randoms los={los}", e_mesh,
                         max(rel_rms(s[a].cpu().numpy(), oshift[a][md]) for a in range(3)),
                         max(maxabs(s[a].cpu().numpy(), oshift[a][md]) for a in range(3)), tol_max)
    ctx.set_option("mg_slab_min_cells", 1 << 22)
    return ok


def tsc_modes(B, ctx, rank, world):
    """IterativeRecon with TSC and PCS on slabs (both reach one plane below the slab and two above): periodic box (wrap
    across the last / first slab) and a radial line of sight."""
    ok = True
    n, L, N = 64, 1000.0, 400_000
    if n % (2 * world):
        return ok
    for los, lo, mas in (((0.0, 0.0, 1.0), 0.0, "tsc"), (None, 700.0, "tsc"), ((0.0, 0.0, 1.0), 0.0, "pcs"), (None, 700.0, "pcs")):
        pos, w = clustered_box(N, L, seed=13, lo=lo)
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
                  box_min=np.full(3, lo, np.float32), los=los, n_iter=3)
        orec = O.IterativeRecon(mas=mas, **kw)
        omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
        oshift = O.read_shifts(orec, *pos, omesh, "sum")
        mine = B.dist.owner_of_z(pos[2], lo, L, n, world) == rank
        d = [torch.from_numpy(np.ascontiguousarray(p[mine])).cuda() for p in pos]
        rec = B.IterativeRecon(mas=mas, **kw)
        mesh = B.dist.run_dist(rec, (n, n, n), *d, torch.from_numpy(np.ascontiguousarray(w[mine])).cuda(), ctx=ctx)
        z_lo, nzl = B.dist.slab_range(ctx)
        s = B.dist.read_shifts_dist(rec, *d, field="sum")
        e_mesh = rel_rms(mesh.cpu().numpy(), omesh[z_lo:z_lo + nzl])
        e_rms = max(rel_rms(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        e_max = max(maxabs(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        good = e_mesh < 1e-4 and e_rms < 1e-4 and e_max < 1e-3
        ok &= good
        print(f"[rank {rank}/{world}] {mas.upper()} los={los}: mesh rel.rms={e_mesh:.2e} shift rel.rms={e_rms:.2e} max={e_max:.2e} "
              f"{'OK' if good else 'FAIL'}", flush=True)
    return ok


def fft_and_sharding(B, ctx, rank, world):
    """(1) the slab transform alone against numpy's rfftn, both exchange schemes, repeated (flag sequence);
    (2) baorec_reconstruct_dist_f32 from an INTERLEAVED split of the catalog (rank r holds particles r, r + P, ...):
    sharding by slab inside the library, reconstruction, results back in the caller's order -- against the oracle."""
    ok = True
    n = 64
    if n % world:
        return ok
    bs, bm = np.full(3, 1000.0, np.float32), np.zeros(3, np.float32)
    rng = np.random.default_rng(3)
    for exchange in ("peer", "peer+sm", "nccl"):       # peer+sm: the remote blocks pushed by the SM kernel (default with 5+ ranks)
        ctx.plan_key = None
        ctx.set_option("push_sm", 1 if exchange == "peer+sm" else 0)
        B.dist.plan(ctx, (n, n, n), bs, bm, exchange=exchange.split("+")[0])
        peer = B.dist.peer_exchange(ctx)
        z_lo, nzl = B.dist.slab_range(ctx)
        nyl = n // world
        for rep in range(3):
            a = rng.standard_normal((n, n, n)).astype(np.float32)
            ref = np.fft.rfftn(a.astype(np.float64))[:, rank * nyl:(rank + 1) * nyl, :]      # [z][yl][x]
            K = B.dist.dist_r2c(ctx, torch.from_numpy(a[z_lo:z_lo + nzl].copy()).cuda())
            got = K.cpu().numpy() if peer else K.cpu().numpy().transpose(2, 0, 1)
            e = max(rel_rms(got.real, ref.real), rel_rms(got.imag, ref.imag))
            back = torch.empty((nzl, n, n), dtype=torch.float32, device="cuda")
            B.dist.dist_c2r(ctx, K, back)
            e2 = rel_rms(back.cpu().numpy() / a.size, a[z_lo:z_lo + nzl])
            good = e < 1e-5 and e2 < 1e-5 and peer == exchange.startswith("peer")
            ok &= good
            if rep == 2 or not good:
                print(f"[rank {rank}/{world}] slab FFT exchange={exchange} (peer copies: {peer}): forward {e:.2e} round trip {e2:.2e} "
                      f"{'OK' if good else 'FAIL'}", flush=True)
    ctx.plan_key = None
    ctx.set_option("dist_exchange", 1)
    ctx.set_option("push_sm", 1 if os.environ.get("MGC_PUSH_SM") == "1" else -1)
    L, N = 1000.0, 400_000
    pos, w = clustered_box(N, L, seed=17)
    pos[2][:64] += np.float32(L)
    for algo, los in (("iterative", (0.0, 0.0, 1.0)), ("multigrid", None)):
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=bs, box_min=bm, los=los)
        orec = O.IterativeRecon(n_iter=3, **kw) if algo == "iterative" else O.MultigridRecon(**kw)
        opos = [p.copy() for p in pos]
        omesh = O.run(orec, (n, n, n), *opos, w)
        oshift = O.read_shifts(orec, *opos, omesh, "sum")
        rec = B.IterativeRecon(n_iter=3, **kw) if algo == "iterative" else B.MultigridRecon(**kw)
        mine = slice(rank, None, world)
        d = [torch.from_numpy(np.ascontiguousarray(p[mine])).cuda() for p in pos]
        got = B.dist.reconstruct_dist(rec, (n, n, n), *d, torch.from_numpy(np.ascontiguousarray(w[mine])).cuda(), field="sum", ctx=ctx)
        e_rms = max(rel_rms(got[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        e_max = max(maxabs(got[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        good = e_rms < 1e-4 and e_max < 1e-3
        ok &= good
        print(f"[rank {rank}/{world}] reconstruct_dist {algo} from an interleaved split: shift rel.rms={e_rms:.2e} max={e_max:.2e} "
              f"{'OK' if good else 'FAIL'}", flush=True)
    # an out-of-box particle on one rank raises on every rank
    bad = [q.clone() for q in d]
    if rank == world - 1:
        bad[2][3] = -7.0
    try:
        B.dist.exchange_catalog(*bad, torch.ones_like(bad[0]), ctx=ctx)
        ok = False
        print(f"[rank {rank}/{world}] out-of-box particle on rank {world - 1} was NOT reported here: FAIL", flush=True)
    except B.OutOfBoxError:
        pass
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = G.load_package()
    ctx = B.Context.get(local)
    B.dist.init_comm(ctx)
    ok = True
    quick = os.environ.get("MGC_QUICK") == "1"       # multigrid cases only (8-rank runs are charged 8x)
    for n, los in ((64, (0.0, 0.0, 1.0)), (96, (0.0, 1.0, 0.0))):
        if n % world or quick:
            continue
        L, N = 1000.0, 400_000
        pos, w = clustered_box(N, L, seed=11)
        pos[2][:64] += np.float32(L)                      # wrap across the last/first slab
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
                  box_min=np.zeros(3, np.float32), los=los, n_iter=3)
        orec = O.IterativeRecon(**kw)
        opos = [p.copy() for p in pos]
        omesh = O.run(orec, (n, n, n), *opos, w)
        oshift = O.read_shifts(orec, *opos, omesh, "sum")
        mine = B.dist.owner_of_z(pos[2], 0.0, L, n, world) == rank
        d = [torch.from_numpy(p[mine]).cuda() for p in pos]
        rec = B.IterativeRecon(**kw)
        mesh = B.dist.run_dist(rec, (n, n, n), *d, torch.from_numpy(w[mine]).cuda(), ctx=ctx)
        z_lo, nzl = B.dist.slab_range(ctx)
        e_mesh = rel_rms(mesh.cpu().numpy(), omesh[z_lo:z_lo + nzl])
        s = B.dist.read_shifts_dist(rec, *d, field="sum")
        e_rms = max(rel_rms(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        e_max = max(maxabs(s[a].cpu().numpy(), oshift[a][mine]) for a in range(3))
        good = e_mesh < 1e-4 and e_rms < 1e-4 and e_max < 1e-3
        ok &= good
        print(f"[rank {rank}/{world}] n={n} los={los} particles={int(mine.sum())} slab=[{z_lo},{z_lo + nzl}) "
              f"mesh rel.rms={e_mesh:.2e} shift rel.rms={e_rms:.2e} max={e_max:.2e} {'OK' if good else 'FAIL'}",
              flush=True)
    ok &= fft_and_sharding(B, ctx, rank, world)
    ok &= more_modes(B, ctx, rank, world, quick)
    if os.environ.get("MGC_TSC") == "1":                 # TSC on slabs (boundary-cell exchange: 1 ghost plane below, 2 above);
        ok &= tsc_modes(B, ctx, rank, world)             # opt-in until it has run on hardware once
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
