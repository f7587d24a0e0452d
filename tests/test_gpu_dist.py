"""Slab-decomposed (multi-GPU) path.  With one GPU the whole distributed code path still runs
(2-D/1-D split FFT, pack + tile transposes, ghost/halo planes, slab-mode scatter/gather) with the
all-to-all degenerating to a copy; tests/multi_gpu_check.py is the same comparison under
torchrun with 2+ ranks (run with gpurun --gpus 2)."""
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


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("chunks", [4, 1, 3])
@pytest.mark.parametrize("shape", [(32, 32, 32), (48, 32, 16), (64, 64, 64), (64, 32, 48)])
def test_dist_fft_roundtrip_and_matches_rfftn(B, dctx, shape, chunks, exchange):
    """Slab transforms, pipelined in `chunks` plane chunks (1 = unpipelined; a chunk count that does
    not divide the local planes falls back to the next smaller one), with both exchange schemes: peer copies +
    sequence flags (k space K[z][yl][x]) and pack / transpose kernels + NCCL (T[yl][x][z]).  Repeated transforms
    exercise the buffer-free / arrived flag sequence."""
    nx, ny, nz = shape
    dctx.set_option("a2a_chunks", chunks)
    try:
        B.dist.plan(dctx, shape, np.full(3, 100.0, np.float32), np.zeros(3, np.float32), exchange=exchange)
        assert B.dist.peer_exchange(dctx) == (exchange == "peer")
        rng = np.random.default_rng(0)
        for rep in range(3):
            a = rng.standard_normal((nz, ny, nx)).astype(np.float32)
            T = B.dist.dist_r2c(dctx, dev(a))
            ref = np.fft.rfftn(a.astype(np.float64), axes=(0, 1, 2))  # [z][y][x]
            got = T.cpu().numpy() if exchange == "peer" else T.cpu().numpy().transpose(2, 0, 1)
            assert rel_rms(got.real, ref.real) < 1e-5 and rel_rms(got.imag, ref.imag) < 1e-5
            back = torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda")
            B.dist.dist_c2r(dctx, T, back)
            assert rel_rms(back.cpu().numpy() / a.size, a) < 1e-5
    finally:
        dctx.set_option("a2a_chunks", 4)
        dctx.set_option("dist_exchange", 1)


def test_slab_owner_matches_host_logic(B, dctx):
    n, L = 64, 1000.0
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    B.dist.plan(dctx, (n, n, n), bs, bm)
    rng = np.random.default_rng(1)
    z = np.concatenate([(rng.random(5000) * L).astype(np.float32), np.float32([0.0, L, L + 3.0, -1.0])])
    got = B.dist.slab_owner(dctx, dev(z)).cpu().numpy()
    for world in (1,):
        assert np.array_equal(got, B.dist.owner_of_z(z, 0.0, L, n, world))
    assert got[-1] == -1 and got[-2] == 0 and got[-3] == 0


@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0)])
def test_run_dist_matches_single_gpu_and_oracle(B, O, dctx, los):
    n, L, N = 64, 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=5)
    pos[2][:40] += np.float32(L)          # exercises the wrap + write-back in slab mode
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=los, n_iter=3)
    orec = O.IterativeRecon(**kw)
    opos = [p.copy() for p in pos]
    omesh = O.run(orec, (n, n, n), *opos, w)
    oshift = O.read_shifts(orec, *opos, omesh, "sum")
    d = [dev(p) for p in pos]
    rec = B.IterativeRecon(**kw)
    mesh = B.dist.run_dist(rec, (n, n, n), *d, dev(w), ctx=dctx)
    assert mesh.shape == (n, n, n)
    assert rel_rms(mesh.cpu().numpy(), omesh) < 1e-4
    for g, o in zip(d, opos):                                   # wrapped positions written back
        assert np.array_equal(g.cpu().numpy().view(np.uint32), o.view(np.uint32))
    for f in ("disp", "sum"):
        s = B.dist.read_shifts_dist(rec, *d, field=f)
        ref = O.read_shifts(orec, *opos, omesh, f)
        for a in range(3):
            assert rel_rms(s[a].cpu().numpy(), ref[a]) < 1e-4
            assert maxabs(s[a].cpu().numpy(), ref[a]) < 1e-3
    newpos = B.dist.read_shifts_dist(rec, *d, field="sum", positions=True)
    for a in range(3):
        assert maxabs(newpos[a].cpu().numpy(), opos[a] - oshift[a]) < 1e-3


def test_shard_unshard_and_reconstruct_dist(B, O, dctx):
    """baorec_shard_catalog_f32 / baorec_unshard_f32 / baorec_reconstruct_dist_f32 with one rank (every particle is
    this rank's; the counting sort, the routing table and the way back still run).  tests/multi_gpu_check.py is the
    same under torchrun with an interleaved split of the catalog."""
    n, L, N = 64, 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=6)
    pos[2][:40] += np.float32(L)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=(0.0, 0.0, 1.0), n_iter=3)
    B.dist.plan(dctx, (n, n, n), kw["box_size"], kw["box_min"])
    d = [dev(p) for p in pos]
    tag = torch.arange(N, dtype=torch.float32, device="cuda")
    sx, sy, sz, st_ = B.dist.exchange_catalog(*d, tag, ctx=dctx)
    assert len(sx) == N
    idx = st_.long()
    assert torch.equal(torch.sort(idx).values, torch.arange(N, device="cuda"))          # a permutation
    assert torch.equal(sx, d[0][idx]) and torch.equal(sy, d[1][idx]) and torch.equal(sz, d[2][idx])
    bx, bt = B.dist.unshard(dctx, (2.0 * sx + sz, st_), d[0])
    assert torch.equal(bt, tag) and torch.equal(bx, 2.0 * d[0] + d[2])
    with pytest.raises(B.OutOfBoxError):
        bad = d[2].clone()
        bad[5] = -3.0
        B.dist.exchange_catalog(d[0], d[1], bad, tag, ctx=dctx)
    # the composite against the oracle, device and host arrays
    orec = O.IterativeRecon(**kw)
    opos = [p.copy() for p in pos]
    omesh = O.run(orec, (n, n, n), *opos, w)
    oshift = O.read_shifts(orec, *opos, omesh, "sum")
    rec = B.IterativeRecon(**kw)
    before = [q.clone() for q in d]
    got = B.dist.reconstruct_dist(rec, (n, n, n), *d, dev(w), field="sum", ctx=dctx)
    hgot = B.dist.reconstruct_dist(B.IterativeRecon(**kw), (n, n, n), *[p.copy() for p in pos], w, field="sum", ctx=dctx)
    for a in range(3):
        assert torch.equal(d[a], before[a])                                             # the caller's arrays are inputs
        assert rel_rms(got[a].cpu().numpy(), oshift[a]) < 1e-4 and maxabs(got[a].cpu().numpy(), oshift[a]) < 1e-3
        assert maxabs(hgot[a], oshift[a]) < 1e-3


def test_read_shifts_dist_needs_a_run_first(B):
    """(TSC on slabs used to be rejected here; it is supported now: tests/test_gpu_zzz3_tsc_slabs_deterministic.py.)"""
    lib = B.lib_loader.load()
    assert lib.baorec_read_shifts_dist_f32(None, None, None, None, None, 0, 0, 0, None, None, None, None) == B.lib_loader.ERR_INVALID


def filled_box(n_data, n_rand, L, lo, n, seed):
    """Data (clumpy) and randoms (uniform) covering the whole box except its last two cells per axis
    (the scatter without wrap rejects the last cell, src/mas.jl:33-35): the smoothed randoms density
    stays far above the ran_min threshold everywhere, so no cell sits on the `ran > threshold`
    discontinuity and the comparison with the oracle is clean."""
    span = L * (1.0 - 2.0 / n)
    d, wd = clustered_box(n_data, span, seed=seed, lo=lo)
    r, wr = clustered_box(n_rand, span, seed=seed + 1, lo=lo, nclump=1, sigma=10.0)   # ~uniform
    return d, wd, r, wr


@pytest.mark.parametrize("min_cells", [0, 16 ** 3])
@pytest.mark.parametrize("los,lo", [((0.0, 0.0, 1.0), 0.0), (None, 700.0)])
def test_multigrid_dist_box(B, O, dctx, min_cells, los, lo):
    """MultigridRecon on slabs (halo planes, slab restriction / prolongation, replicated coarse
    levels below `mg_slab_min_cells`) against the oracle."""
    n, L, N = 64, 1000.0, 200_000
    pos, w = clustered_box(N, L, seed=21, lo=lo)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.full(3, lo, np.float32), los=los)
    orec = O.MultigridRecon(**kw)
    opos = [p.copy() for p in pos]
    ophi = O.run(orec, (n, n, n), *opos, w)
    dctx.set_option("mg_slab_min_cells", min_cells)
    try:
        rec = B.MultigridRecon(**kw)
        d = [dev(p) for p in pos]
        phi = B.dist.run_dist(rec, (n, n, n), *d, dev(w), ctx=dctx)
        assert rel_rms(phi.cpu().numpy(), ophi) < 1e-4
        for f in ("disp", "sum"):
            s = B.dist.read_shifts_dist(rec, *d, field=f)
            ref = O.read_shifts(orec, *opos, ophi, f)
            for a in range(3):
                assert rel_rms(s[a].cpu().numpy(), ref[a]) < 1e-4
                assert maxabs(s[a].cpu().numpy(), ref[a]) < 1e-3
    finally:
        dctx.set_option("mg_slab_min_cells", 1 << 22)


@pytest.mark.parametrize("algo", ["iterative", "multigrid"])
@pytest.mark.parametrize("los,lo", [(None, 700.0), ((0.0, 0.0, 1.0), 700.0)])
def test_run_dist_with_randoms(B, O, dctx, algo, los, lo):
    """Randoms set-up (two smoothed meshes, DC sums, global randoms count, threshold) on slabs;
    IterativeRecon radial LOS = iterate! on slabs, fixed LOS = fused pass on delta_s."""
    n, L = 64, 1000.0
    d, wd, r, wr = filled_box(60_000, 600_000, L, lo, n, seed=31)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.full(3, lo, np.float32), los=los)
    if algo == "iterative":
        orec, rec = O.IterativeRecon(**kw), B.IterativeRecon(**kw)
        omesh = O.reconstructed_overdensity(np.zeros((n, n, n), np.float32), orec, *d, wd, *r, wr)
    else:
        orec, rec = O.MultigridRecon(**kw), B.MultigridRecon(**kw)
        omesh = O.reconstructed_potential(np.zeros((n, n, n), np.float32), orec, *d, wd, *r, wr)
    gd, gr = [dev(p) for p in d], [dev(p) for p in r]
    dctx.set_option("mg_slab_min_cells", 16 ** 3)
    try:
        mesh = B.dist.run_dist(rec, (n, n, n), *gd, dev(wd), *gr, dev(wr), ctx=dctx)
    finally:
        dctx.set_option("mg_slab_min_cells", 1 << 22)
    g = mesh.cpu().numpy()
    if algo == "multigrid":   # the potential is defined up to a constant
        assert rel_rms(g - g.mean(), omesh - omesh.mean()) < 1e-4
    else:
        assert rel_rms(g, omesh) < 1e-4
    s = B.dist.read_shifts_dist(rec, *gd, field="sum")
    ref = O.read_shifts(orec, *d, omesh, "sum")
    for a in range(3):
        assert rel_rms(s[a].cpu().numpy(), ref[a]) < 1e-4
        assert maxabs(s[a].cpu().numpy(), ref[a]) < 1e-3


@pytest.mark.parametrize("own", [-1, 0])
def test_slab_z_pass_of_a_1024_point_axis(B, dctx, own):
    """nz = 1024 on slabs: the z pass of the K[z][yl][x] layout is the column kernel of csrc/fft.cu (16 columns per tile,
    own_slab_z) instead of cuFFT's strided plan (option own_fft = 0): both match rfftn and invert back."""
    shape = (40, 16, 1024)
    nx, ny, nz = shape
    dctx.set_option("own_fft", own)
    try:
        B.dist.plan(dctx, shape, np.full(3, 100.0, np.float32), np.zeros(3, np.float32), exchange="peer")
        a = np.random.default_rng(3).standard_normal((nz, ny, nx)).astype(np.float32)
        dctx.profile(True)
        T = B.dist.dist_r2c(dctx, dev(a))
        back = torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda")
        B.dist.dist_c2r(dctx, T.clone(), back)
        names = set(dctx.profile_read())
        dctx.profile(False)
        assert any(k.startswith("fft_cols_kernel") for k in names) == (own != 0), names
        assert any(k.startswith("cufft_1d_z") for k in names) == (own == 0), names
        ref = np.fft.rfftn(a.astype(np.float64), axes=(0, 1, 2))
        got = T.cpu().numpy()
        assert rel_rms(got.real, ref.real) < 1e-5 and rel_rms(got.imag, ref.imag) < 1e-5
        assert rel_rms(back.cpu().numpy() / a.size, a) < 1e-5
    finally:
        dctx.set_option("own_fft", -1)
