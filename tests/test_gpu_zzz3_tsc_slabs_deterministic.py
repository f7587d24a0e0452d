"""GPU: TSC on the slab-decomposed (multi-GPU) path -- the boundary-cell exchange for the 27-point stencil
(one ghost plane below the slab, two above; csrc/dist.cu: dist_scatter_tsc, csrc/mas_math.cuh: local_plane1).
With one GPU the whole distributed code path runs (slab-layout scatter and gather kernels, ghost / halo planes,
split FFT) with the exchanges degenerating to copies; tests/multi_gpu_check.py holds the same comparison under
torchrun with 2+ ranks.  The index mapping and the exchange pattern are validated rank by rank on the CPU
(tests/test_mas_hostcheck.py, P = 1, 2, 4, 8).  PCS (64-point stencil) reaches exactly the planes TSC reaches and
shares its slab layout and exchange."""
import numpy as np
import pytest

from util import clustered_box, rel_rms, maxabs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture()
def dctx(B):
    ctx = B.Context.get(0)
    B.dist.init_comm(ctx)
    yield ctx
    ctx.plan_key = None       # force a fresh single-GPU plan for whoever comes next


@pytest.mark.parametrize("N", [20_000, 300_000])          # catalog-order kernels / binned kernels
@pytest.mark.parametrize("los,mas", [((0.0, 0.0, 1.0), "tsc"), (None, "tsc"), ((0.0, 0.0, 1.0), "pcs"), (None, "pcs")])   # PCS reaches the same planes as TSC
def test_iterative_tsc_on_slabs_matches_the_oracle(B, O, dctx, N, los, mas):
    n, L = 64, 1000.0
    lo = 0.0 if los is not None else 700.0
    pos, w = clustered_box(N, L, seed=5, lo=lo)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32), box_min=np.full(3, lo, np.float32),
              los=los, n_iter=3)
    orec = O.IterativeRecon(**kw)
    orec.mas = mas
    omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
    rec = B.IterativeRecon(mas=mas, **kw)
    d = [dev(p) for p in pos]
    mesh = B.dist.run_dist(rec, (n, n, n), *d, dev(w), ctx=dctx)
    hmesh = mesh.cpu().numpy()
    assert rel_rms(hmesh, omesh) < 1e-4
    for f in ("disp", "sum"):
        s = B.dist.read_shifts_dist(rec, *d, field=f)
        ref = O.read_shifts(orec, *pos, omesh, f)
        for a in range(3):
            assert rel_rms(s[a].cpu().numpy(), ref[a]) < 1e-4
            assert maxabs(s[a].cpu().numpy(), ref[a]) < 1e-3
    # the single-GPU TSC path (validated on hardware) gives the same mesh (last: it re-plans the context)
    rec1 = B.IterativeRecon(mas=mas, **kw)
    mesh1 = B.run(rec1, (n, n, n), *[dev(p) for p in pos], dev(w))
    assert rel_rms(hmesh, mesh1.cpu().numpy()) < 1e-5


def test_multigrid_tsc_on_slabs_matches_the_oracle(B, O, dctx):
    n, L, N = 64, 1000.0, 200_000
    pos, w = clustered_box(N, L, seed=21)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32), box_min=np.zeros(3, np.float32),
              los=(0.0, 0.0, 1.0))
    orec = O.MultigridRecon(**kw)
    orec.mas = "tsc"
    ophi = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
    rec = B.MultigridRecon(mas="tsc", **kw)
    d = [dev(p) for p in pos]
    phi = B.dist.run_dist(rec, (n, n, n), *d, dev(w), ctx=dctx)
    assert rel_rms(phi.cpu().numpy(), ophi) < 1e-4
    s = B.dist.read_shifts_dist(rec, *d, field="sum")
    ref = O.read_shifts(orec, *pos, ophi, "sum")
    for a in range(3):
        assert rel_rms(s[a].cpu().numpy(), ref[a]) < 1e-4
        assert maxabs(s[a].cpu().numpy(), ref[a]) < 1e-3


# ---- option "deterministic_scatter" (64-bit fixed-point integer reductions; csrc/mas_math.cuh: deposit_fixed) --------
# Arithmetic validated on the CPU (tests/test_mas_hostcheck.py); green on a B200 since round 2 (profiles/r2_gpu_suite_final2.txt).
@pytest.mark.parametrize("N", [50_000, 400_000])          # catalog-order kernel / unified sort + records kernel
def test_deterministic_scatter_is_bit_reproducible(B, O, N):
    n, L = 64, 1000.0
    pos, w = clustered_box(N, L, seed=17)
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    ctx = B.Context.get(0)
    ref = O.cic_scatter(np.zeros((n, n, n), np.float32), *[p.copy() for p in pos], w, bs, bm, True)
    perm = np.random.default_rng(1).permutation(N)
    try:
        ctx.set_option("deterministic_scatter", 1)
        meshes = []
        for order in (np.arange(N), np.arange(N), perm):
            rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
            B.cic(rho, *(dev(p[order]) for p in pos), dev(w[order]), bs, bm, wrap=True)
            meshes.append(rho)
        assert torch.equal(meshes[0], meshes[1]) and torch.equal(meshes[0], meshes[2])      # any order, same bits
        assert maxabs(meshes[0].cpu().numpy(), ref) <= 2e-6 * float(ref.max())
        kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=bs, box_min=bm, los=(0.0, 0.0, 1.0), n_iter=3)
        runs = [B.run(B.IterativeRecon(**kw), (n, n, n), *(dev(p) for p in pos), dev(w)) for _ in range(2)]
        assert torch.equal(runs[0], runs[1])                                                 # the whole reconstruction
        with pytest.raises(B.BaorecError):                                                   # TSC / PCS / slabs: refused, not silently ignored
            B.cic(torch.zeros((n, n, n), dtype=torch.float32, device="cuda"), *(dev(p) for p in pos), dev(w), bs, bm, wrap=True, mas="tsc")
    finally:
        ctx.set_option("deterministic_scatter", 0)
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.cic(rho, *(dev(p) for p in pos), dev(w), bs, bm, wrap=True)
    assert maxabs(rho.cpu().numpy(), meshes[0].cpu().numpy()) <= 2e-6 * float(ref.max())     # default path: same to rounding
