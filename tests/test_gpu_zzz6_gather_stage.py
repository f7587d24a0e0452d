"""GPU: option "gather_stage" -- how the tile gather stages its window in shared memory: 0 = the round-1 cp.async loop,
1 = the strength-reduced cp.async staging (gather_tile_kernel<3, 1>, csrc/mas.cu: each warp owns rows py = warp,
warp + 4 (, 8) of every (field, plane), shift decode of the tile index, first record prefetched), 2 = tensor-map TMA
(gather_tile_tma_kernel: six cp.async.bulk.tensor.3d copies per tile, the default since it was measured).  Same
shared-memory window, same arithmetic: the shifts must be bit-identical across the three.  Also the options of the
scatter ("scatter_pairs")."""
import numpy as np
import pytest

from util import clustered_box, maxabs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
f32 = np.float32
DEFAULT_GATHER_STAGE, DEFAULT_SCATTER_PAIRS = 2, 2      # csrc/internal.cuh (both measured on hardware in round 2, then made the default)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("n,los", [(128, (0.0, 0.0, 1.0)), (256, None)])     # one / two full tiles along x; fixed and radial epilogue
def test_staged_variant_is_bit_identical(B, O, n, los):
    L, N = 1000.0, 400_000                                                    # >= 2^18 particles: the tile-sorted gather
    lo = 0.0 if los is not None else 600.0
    pos, w = clustered_box(N, L, seed=31, lo=lo)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, f32), box_min=np.full(3, lo, f32), los=los, n_iter=3)
    ctx = B.Context.get(0)
    d = [dev(p) for p in pos]
    rec = B.IterativeRecon(**kw)
    mesh = B.run(rec, (n, n, n), *d, dev(w))
    other = [dev(p[::-1].copy()) for p in pos]                               # a second catalog: no reuse of run!'s sort
    try:
        ctx.set_option("gather_stage", 0)
        ref = [B.read_shifts(rec, *cat, mesh, field="sum") for cat in (d, other)]
        ctx.set_option("gather_stage", 1)
        got = [B.read_shifts(rec, *cat, mesh, field="sum") for cat in (d, other)]
    finally:
        ctx.set_option("gather_stage", DEFAULT_GATHER_STAGE)
    for r, g in zip(ref, got):
        for a in range(3):
            assert torch.equal(r[a], g[a])
    if n == 128 and los is not None:                                          # and the oracle agrees
        orec = O.IterativeRecon(**kw)
        omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
        oref = O.read_shifts(orec, *pos, omesh, "sum")
        for a in range(3):
            assert maxabs(got[0][a].cpu().numpy(), oref[a]) < 1e-3


@pytest.mark.parametrize("grid,los", [((128, 128, 128), (0.0, 0.0, 1.0)),    # one tile along x: every tile takes column 128 from x = 0
                                      ((256, 256, 64), None),                 # two tiles along x, radial epilogue
                                      ((136, 68, 40), (0.0, 0.0, 1.0)),       # meshes off the tile grid: partial tiles stage with plain loads
                                      ((40, 68, 136), None)])
def test_tma_staged_variant_is_bit_identical(B, grid, los):
    """Option "gather_stage" = 2: the window of every full tile is fetched by the TMA engine through one 3-D tensor map
    per field (gather_tile_tma_kernel, csrc/mas.cu: 2 x 3 cp.async.bulk.tensor.3d copies issued by one thread, one
    mbarrier; the periodic column / row the engine cannot fetch come from ordinary loads).  Same window, same
    arithmetic: bit-identical shifts."""
    L, N = 1000.0, 400_000
    lo = 0.0 if los is not None else 600.0
    pos, w = clustered_box(N, L, seed=41, lo=lo)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, f32), box_min=np.full(3, lo, f32), los=los, n_iter=2)
    ctx = B.Context.get(0)
    d = [dev(p) for p in pos]
    rec = B.IterativeRecon(**kw)
    mesh = B.run(rec, grid, *d, dev(w))
    other = [dev(p[::-1].copy()) for p in pos]
    try:
        ctx.set_option("gather_stage", 1)
        ref = [B.read_shifts(rec, *cat, mesh, field="sum") for cat in (d, other)]
        ctx.set_option("gather_stage", 2)
        ctx.profile(True)
        got = [B.read_shifts(rec, *cat, mesh, field="sum") for cat in (d, other)]
        names = set(ctx.profile_read())
        ctx.profile(False)
    finally:
        ctx.set_option("gather_stage", DEFAULT_GATHER_STAGE)
    assert any("gather_tile_tma_kernel" in k for k in names), names
    for r, g in zip(ref, got):
        for a in range(3):
            assert torch.equal(r[a], g[a])


def test_paired_scatter_matches_the_oracle(B, O):
    """Option "scatter_pairs" = 1 / 2: aligned x pairs (and, 2, aligned quads {0, a, b, 0}) of cic! through
    red.global.add.v2.f32 / .v4.f32 (csrc/mas_math.cuh:
    deposit_pairs; arithmetic checked on the CPU in tests/test_mas_hostcheck.py).  Same cells and values as the default
    kernel: the mesh differs from the oracle's serial sum only by the order of the Float32 additions."""
    n, L, N = 128, 1000.0, 400_000
    pos, w = clustered_box(N, L, seed=33)
    pos[0][:50] += f32(L)                                                     # wrapped by cic!, written back
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)
    opos = [p.copy() for p in pos]
    ref = O.cic_scatter(np.zeros((n, n, n), f32), *opos, w, bs, bm, True)
    ctx = B.Context.get(0)
    meshes = []
    for on in (0, 1, 2):                                                      # scalar; aligned pairs; pairs + aligned quads
        try:
            ctx.set_option("scatter_pairs", on)
            rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
            d = [dev(p) for p in pos]
            B.cic(rho, *d, dev(w), bs, bm, wrap=True)
        finally:
            ctx.set_option("scatter_pairs", DEFAULT_SCATTER_PAIRS)
        for g, o in zip(d, opos):
            assert np.array_equal(g.cpu().numpy().view(np.uint32), o.view(np.uint32))
        meshes.append(rho.cpu().numpy())
        assert maxabs(meshes[-1], ref) <= 2e-6 * float(ref.max())
    for m in meshes[1:]:
        assert abs(float(m.sum(dtype=np.float64)) / float(w.sum(dtype=np.float64)) - 1) < 1e-6


def test_vector_tsc_scatter_matches_the_oracle(B, O):
    """Option "scatter_pairs" for the binned TSC scatter: one aligned quad or two aligned pairs per stencil row
    (csrc/mas_math.cuh: deposit_tsc_vec; against the scalar deposit bit for bit on the CPU, tests/test_mas_hostcheck.py)."""
    n, L, N = 128, 1000.0, 400_000
    pos, w = clustered_box(N, L, seed=35)
    bs, bm = np.full(3, L, f32), np.zeros(3, f32)
    ref = O.tsc_scatter(np.zeros((n, n, n), f32), *pos, w, bs, bm, True)
    ctx = B.Context.get(0)
    for on in (0, 1, 2):
        try:
            ctx.set_option("scatter_pairs", on)
            rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
            B.cic(rho, *(dev(p) for p in pos), dev(w), bs, bm, wrap=True, mas="tsc")
        finally:
            ctx.set_option("scatter_pairs", DEFAULT_SCATTER_PAIRS)
        h = rho.cpu().numpy()
        assert maxabs(h, ref) <= 3e-6 * float(ref.max())
        assert abs(float(h.sum(dtype=np.float64)) / float(w.sum(dtype=np.float64)) - 1) < 1e-6


@pytest.mark.parametrize("mas", ["cic", "tsc"])
def test_vector_reductions_on_the_slab_path(B, O, mas):
    """Option "scatter_pairs" on the z-binned scatter of the slab-decomposed path (scatter_sorted_cic_vec_kernel /
    scatter_sorted_tsc_vec_kernel writing into the ghost-plane layouts): the distributed reconstruction still matches
    the oracle."""
    n, L, N = 64, 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=37)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, f32), box_min=np.zeros(3, f32),
              los=(0.0, 0.0, 1.0), n_iter=3)
    orec = O.IterativeRecon(**kw)
    orec.mas = mas
    omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
    ctx = B.Context.get(0)
    B.dist.init_comm(ctx)
    try:
        ctx.set_option("scatter_pairs", 1)
        rec = B.IterativeRecon(mas=mas, **kw)
        mesh = B.dist.run_dist(rec, (n, n, n), *(dev(p) for p in pos), dev(w), ctx=ctx)
        h = mesh.cpu().numpy()
    finally:
        ctx.set_option("scatter_pairs", DEFAULT_SCATTER_PAIRS)
        ctx.plan_key = None
    assert float(np.sqrt(np.mean((h.astype(np.float64) - omesh) ** 2)) / np.sqrt(np.mean(omesh.astype(np.float64) ** 2))) < 1e-4


def test_tma_staging_on_the_slab_path(B):
    """Option "gather_stage" = 2 on the slab-decomposed read-back: the tensor maps describe the nzp planes of the slab
    buffers (ghost plane included), z + 1 never wraps there.  Bit-identical to the cp.async staging."""
    n, L, N = 128, 1000.0, 400_000
    pos, w = clustered_box(N, L, seed=43)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, f32), box_min=np.zeros(3, f32),
              los=(0.0, 0.0, 1.0), n_iter=2)
    ctx = B.Context.get(0)
    B.dist.init_comm(ctx)
    d = [dev(p) for p in pos]
    out = {}
    try:
        rec = B.IterativeRecon(**kw)
        B.dist.run_dist(rec, (n, n, n), *d, dev(w), ctx=ctx)                  # ONE solve (its scatter's atomics are not reproducible), two read-backs
        for stage in (1, 2):
            ctx.set_option("gather_stage", stage)
            ctx.profile(True)
            out[stage] = B.dist.read_shifts_dist(rec, *d, field="sum")
            names = set(ctx.profile_read())
            ctx.profile(False)
            tiles = [k for k in names if "gather_tile" in k]
            assert tiles and all(("gather_tile_tma_kernel" in k) == (stage == 2) for k in tiles), names
    finally:
        ctx.set_option("gather_stage", DEFAULT_GATHER_STAGE)
        ctx.plan_key = None
    for a in range(3):
        assert torch.equal(out[1][a], out[2][a])
