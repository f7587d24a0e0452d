"""GPU parity of the fast paths against the oracle and against the reference-sequence path:
  * z-slab binned scatter / gather (forced on for small catalogs with bin_min_particles = 0);
  * fused fixed-LOS k-space solve (fuse_kspace = 1) versus the sequence of iterate! calls."""
import contextlib

import numpy as np
import pytest

from util import uniform_box, clustered_box, lightcone, rel_rms, maxabs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@contextlib.contextmanager
def options(B, **kw):
    ctx = B.Context.get(0)
    defaults = {"bin_min_particles": 1 << 18, "fuse_kspace": 1, "own_fft": -1, "gather_tiles": 1,
                "fft_split_planes": 0, "scatter_tiles": 0, "fft_tile_cols": -1,
                "unified_sort": 1}
    try:
        for k, v in kw.items():
            ctx.set_option(k, v)
        yield ctx
    finally:
        for k in kw:
            ctx.set_option(k, defaults[k])


def test_unknown_option(B):
    with pytest.raises(B.BaorecError):
        B.Context.get(0).set_option("no_such_option", 1)


@pytest.mark.parametrize("tiles", [0, 1, 2])      # z-slab bins, scatter tiles, unified sort (default)
@pytest.mark.parametrize("wrap", [True, False])
@pytest.mark.parametrize("n", [64, 20])
def test_binned_scatter(B, O, wrap, n, tiles):
    L, N = 1000.0, 200_000
    pos, w = clustered_box(N, L, seed=17)
    if wrap:
        pos[0][:50] += np.float32(L)
        pos[2][50:80] += np.float32(L)
    else:
        for p in pos:
            np.clip(p, 0, np.float32(L - L / n - 1e-3), out=p)
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    ox, oy, oz = (p.copy() for p in pos)
    orho = O.cic_scatter(np.zeros((n, n, n), np.float32), ox, oy, oz, w, bs, bm, wrap)
    gx, gy, gz = (dev(p) for p in pos)
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    with options(B, bin_min_particles=0, scatter_tiles=tiles % 2, unified_sort=int(tiles == 2)):
        B.cic(rho, gx, gy, gz, dev(w), bs, bm, wrap=wrap)
    assert maxabs(rho.cpu().numpy(), orho) <= 2e-5 * max(1.0, float(orho.max()))
    for g, o in zip((gx, gy, gz), (ox, oy, oz)):       # wrapped positions written back
        assert np.array_equal(g.cpu().numpy().view(np.uint32), o.view(np.uint32))


def test_binned_scatter_out_of_box(B):
    n, L = 32, 100.0
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    rng = np.random.default_rng(0)
    x = (rng.random(1000) * L).astype(np.float32)
    y = (rng.random(1000) * L).astype(np.float32)
    z = (rng.random(1000) * L).astype(np.float32)
    x[7], z[11], y[500] = -3.0, 420.0, np.nan
    for tiles in (0, 1, 2):
        rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
        with options(B, bin_min_particles=0, scatter_tiles=tiles % 2, unified_sort=int(tiles == 2)):
            with pytest.raises(B.OutOfBoxError) as ei:
                B.cic(rho, dev(x), dev(y), dev(z), dev(np.ones(1000, np.float32)), bs, bm, wrap=True)
        assert "3 particle(s)" in str(ei.value)
        assert abs(float(rho.sum()) - 997.0) < 1e-2       # the others were deposited


@pytest.mark.parametrize("mas", ["cic", "tsc"])
def test_binned_gather_bit_exact(B, O, mas):
    n, L, lo, N = 48, 777.0, -123.0, 150_000
    rng = np.random.default_rng(5)
    fld = rng.standard_normal((n, n, n)).astype(np.float32)
    pos, _ = uniform_box(N, L, seed=6, lo=lo)
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    ref = (O.read_cic if mas == "cic" else O.read_tsc)(fld, *pos, bs, bm)
    out = torch.empty(N, dtype=torch.float32, device="cuda")
    with options(B, bin_min_particles=0):
        B.read_cic(out, dev(fld), *(dev(p) for p in pos), bs, bm, mas=mas)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_binned_tsc_scatter(B, O):
    n, L, N = 40, 500.0, 100_000
    pos, w = clustered_box(N, L, seed=3)
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    orho = O.tsc_scatter(np.zeros((n, n, n), np.float32), *pos, w, bs, bm, True)
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    with options(B, bin_min_particles=0):
        B.cic(rho, *(dev(p) for p in pos), dev(w), bs, bm, wrap=True, mas="tsc")
    assert maxabs(rho.cpu().numpy(), orho) <= 2e-5 * float(orho.max())


@pytest.mark.parametrize("opts", [dict(bin_min_particles=0, fuse_kspace=1), dict(bin_min_particles=0, fuse_kspace=0),
                                  dict(bin_min_particles=0, fuse_kspace=1, own_fft=1),
                                  dict(bin_min_particles=0, fuse_kspace=1, fft_split_planes=8),
                                  dict(bin_min_particles=0, fuse_kspace=1, scatter_tiles=1, unified_sort=0),
                                  dict(bin_min_particles=0, fuse_kspace=1, unified_sort=0),
                                  dict(bin_min_particles=0, fuse_kspace=0, fft_split_planes=24),
                                  dict(bin_min_particles=0, fuse_kspace=0, own_fft=1, gather_tiles=0),
                                  dict(bin_min_particles=1 << 40, fuse_kspace=0),
                                  dict(bin_min_particles=1 << 40, fuse_kspace=1)])
@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (0.0, 1.0, 0.0)])
def test_box_recon_all_paths(B, O, opts, los):
    """Every combination of the fast paths reproduces the oracle's run! + read_shifts."""
    n, L, N = 64, 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=33)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=los, n_iter=3)
    orec = O.IterativeRecon(**kw)
    omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
    oshift = O.read_shifts(orec, *pos, omesh, "sum")
    d = [dev(p) for p in pos]
    with options(B, **opts):
        rec = B.IterativeRecon(**kw)
        mesh = B.run(rec, (n, n, n), *d, dev(w))
        s = B.read_shifts(rec, *d, mesh, field="sum")
        # host pipeline (uses the cached delta_k when fused)
        rec_h = B.IterativeRecon(**kw)
        hmesh = np.empty((n, n, n), np.float32)
        B.run(rec_h, (n, n, n), *[p.copy() for p in pos], w, mesh_out=hmesh)
        sh = B.read_shifts(rec_h, *pos, None, field="sum")
    assert rel_rms(mesh.cpu().numpy(), omesh) < 1e-4
    assert rel_rms(hmesh, omesh) < 1e-4
    for a in range(3):
        for got in (s[a].cpu().numpy(), sh[a]):
            assert rel_rms(got, oshift[a]) < 1e-4
            assert maxabs(got, oshift[a]) < 1e-3


@pytest.mark.parametrize("n_iter", [0, 1, 5])
def test_fused_n_iter(B, O, n_iter):
    n, L, N = 48, 800.0, 100_000
    pos, w = clustered_box(N, L, seed=2)
    kw = dict(bias=1.7, f=0.9, smoothing_radius=12.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=(1.0, 0.0, 0.0), n_iter=n_iter)
    omesh = O.run(O.IterativeRecon(**kw), (n, n, n), *[p.copy() for p in pos], w)
    mesh = B.run(B.IterativeRecon(**kw), (n, n, n), *(dev(p) for p in pos), dev(w))
    assert rel_rms(mesh.cpu().numpy(), omesh) < 1e-4


def test_fused_with_randoms_fixed_los(B, O):
    """Fixed LOS + randoms (survey geometry with a plane-parallel LOS): fused MODE 1."""
    from test_gpu_iterative import check_flips_explained, LC, NLC
    n = NLC
    d, wd, r, wr = lightcone(60_000, 600_000, seed=21, **LC)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, los=(1.0, 0.0, 0.0), n_iter=3)
    gd, gr = [dev(p) for p in d], [dev(p) for p in r]
    outs = {}
    for fuse in (0, 1):
        with options(B, fuse_kspace=fuse):
            rec = B.IterativeRecon(**kw)
            rec.box_size, rec.box_min = B.setup_box(*gr, 500.0)
            ds = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
            B.setup_fft(rec, ds)
            B.setup_overdensity(ds, rec, *gd, dev(wd), *gr, dev(wr))
            mesh = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
            B.reconstructed_overdensity(mesh, rec, *gd, dev(wd), *gr, dev(wr))
            outs[fuse] = (ds.cpu().numpy() != 0, mesh.cpu().numpy())
    orec = O.IterativeRecon(**kw)
    orec.box_size, orec.box_min = O.setup_box(*r, np.float32(500))
    info = {}
    O.setup_overdensity(np.zeros((n, n, n), np.float32), orec, *d, wd, *r, wr, info=info)
    for fuse in (0, 1):
        nfl = check_flips_explained(outs[fuse][0], info)
        if nfl == 0:   # same threshold decisions as the oracle: compare the reconstructed mesh
            omesh = O.run(O.IterativeRecon(**kw), (n, n, n), *d, wd, *r, wr)
            assert rel_rms(outs[fuse][1], omesh) < 3e-4


# ---- own FFT path (cuFFT 1-D along x + column kernels along y/z with fused k-space operators) -----
@pytest.mark.parametrize("shape", [(64, 256, 256), (128, 512, 256), (40, 256, 1024), (96, 1024, 256), (24, 256, 2048),
                                   (16, 2048, 512)])
def test_own_fft_smooth_and_displacements_match_cufft_and_oracle(B, O, shape):
    """Column kernels of csrc/fft.cu for every supported length (256 ... 2048 along y and z; shorter axes fall back to
    cuFFT), partial column tiles included (nx/2 + 1 is never a multiple of 8)."""
    nx, ny, nz = shape
    L = 1000.0
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    rng = np.random.default_rng(3)
    fld = rng.standard_normal((nz, ny, nx)).astype(np.float32)
    ref = O.smooth(fld.copy(), np.float32(9.0), bs)
    out = {}
    for own in (1, 0):
        with options(B, own_fft=own):
            g = dev(fld)
            B.smooth(g, 9.0, bs)
            rec = B.IterativeRecon(bias=1.0, f=0.5, smoothing_radius=9.0, box_size=bs, box_min=bm, los=(0, 0, 1))
            psi = B.displacement_meshes(dev(fld), rec)
            out[own] = (g.cpu().numpy(), [p.cpu().numpy() for p in psi])
    assert rel_rms(out[1][0], ref) < 2e-6 and rel_rms(out[0][0], ref) < 2e-6
    opsi = O.displacement_meshes(fld, O.IterativeRecon(bias=1.0, f=0.5, smoothing_radius=9.0, box_size=bs, box_min=bm))
    for a in range(3):
        # white noise has full power at the Nyquist modes, where i*k breaks Hermitian symmetry: the own
        # path drops the offending imaginary parts like FFTW/pocketfft (= the oracle); cuFFT's C2R
        # does not, so the library path is only compared on the smoothed field above
        assert rel_rms(out[1][1][a], opsi[a]) < 5e-6


@pytest.mark.parametrize("los,tile_cols", [((0.0, 0.0, 1.0), 8), ((0.0, 1.0, 0.0), 8), ((0.0, 0.0, 1.0), 4), ((0.0, 0.0, 1.0), 16)])
def test_own_fft_fused_z_pass_matches_the_oracle(B, O, los, tile_cols):
    """run! + read_shifts with own_fft = 1 on a mesh with nz = 1024: the z pass is ONE kernel (forward z transform,
    smoothing + normalisation + all iterations in registers, delta_k kept, inverse z transform) and the read-back emits
    the three displacement fields from one read of delta_k."""
    grid, L, N = (32, 256, 1024), 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=41)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=los, n_iter=3)
    orec = O.IterativeRecon(**kw)
    omesh = O.run(orec, grid, *[p.copy() for p in pos], w)
    oshift = O.read_shifts(orec, *pos, omesh, "sum")
    d = [dev(p) for p in pos]
    with options(B, own_fft=1, fft_tile_cols=tile_cols):      # 4 / 16 columns per tile: the experimental widths of the N = 1024 kernels
        rec = B.IterativeRecon(**kw)
        mesh = B.run(rec, grid, *d, dev(w))
        s = B.read_shifts(rec, *d, mesh, field="sum")
        names = set()
        ctx = B.Context.get(0)
        ctx.profile(True)
        B.run(rec, grid, *d, dev(w))
        names = set(ctx.profile_read())
        ctx.profile(False)
    assert "fft_z_solve_kernel" in names and not any(k.startswith("cufft_r2c") for k in names), names
    assert rel_rms(mesh.cpu().numpy(), omesh) < 1e-4
    for a in range(3):
        assert rel_rms(s[a].cpu().numpy(), oshift[a]) < 1e-4 and maxabs(s[a].cpu().numpy(), oshift[a]) < 1e-3


def test_result_cache_read_skips_forward_transform_and_matches(B, O):
    """read_shifts(recon, x,y,z, recon.result_cache) reuses the delta_k kept by run!; a copy of the mesh
    (or a mesh modified in place) goes through the forward transform like the reference."""
    n, L, N = 64, 1000.0, 200_000
    pos, w = clustered_box(N, L, seed=8)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=(0.0, 0.0, 1.0), n_iter=3)
    d = [dev(p) for p in pos]
    rec = B.IterativeRecon(**kw)
    mesh = B.run(rec, (n, n, n), *d, dev(w))
    ctx = rec.fft_plan.ctx
    _, f0 = ctx.launch_counts()
    s_cached = B.read_shifts(rec, *d, mesh, field="sum")
    _, f1 = ctx.launch_counts()
    s_plain = B.read_shifts(rec, *d, mesh.clone(), field="sum")
    _, f2 = ctx.launch_counts()
    assert f1 - f0 == 3 and f2 - f1 == 4            # 3 C2R vs R2C + 3 C2R
    for a in range(3):
        assert maxabs(s_cached[a].cpu().numpy(), s_plain[a].cpu().numpy()) < 1e-4   # one C2R/R2C round trip of rounding
    orec = O.IterativeRecon(**kw)
    omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
    oshift = O.read_shifts(orec, *pos, omesh, "sum")
    for a in range(3):
        assert maxabs(s_cached[a].cpu().numpy(), oshift[a]) < 1e-3
    mesh.mul_(2.0)                                     # in-place edit: the cache must not be used
    s_mod = B.read_shifts(rec, *d, mesh, field="disp")
    s_ref = B.read_shifts(rec, *d, mesh.clone(), field="disp")
    for a in range(3):
        assert maxabs(s_mod[a].cpu().numpy(), s_ref[a].cpu().numpy()) < 1e-6


@pytest.mark.parametrize("algo", ["iterative", "multigrid"])
def test_displacement_meshes_are_reused_across_catalogs(B, O, algo):
    """reconstructed_positions for data, then for two other catalogs against the same
    recon.result_cache (examples/simulation.jl:32-35): only the first call pays for the transforms;
    a new run! or an edited mesh invalidates the kept displacement meshes."""
    n, L = 64, 1000.0
    pos, w = clustered_box(200_000, L, seed=9)
    rnd, _ = clustered_box(300_000, L, seed=10, nclump=1, sigma=10.0)
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=(0.0, 0.0, 1.0))
    Rec, ORec = (B.IterativeRecon, O.IterativeRecon) if algo == "iterative" else (B.MultigridRecon, O.MultigridRecon)
    d, r = [dev(p) for p in pos], [dev(p) for p in rnd]
    rec = Rec(**kw)
    mesh = B.run(rec, (n, n, n), *d, dev(w))
    ctx = rec.fft_plan.ctx
    _, f0 = ctx.launch_counts()
    pd = B.reconstructed_positions(rec, *d, field="sum")
    _, f1 = ctx.launch_counts()
    pr_sym = B.reconstructed_positions(rec, *r, field="sum")
    pr_iso = B.reconstructed_positions(rec, *r, field="disp")
    _, f2 = ctx.launch_counts()
    assert f1 - f0 == (3 if algo == "iterative" else 4) and f2 == f1
    orec = ORec(**kw)
    omesh = O.run(orec, (n, n, n), *[p.copy() for p in pos], w)
    for got, cat, f in ((pd, pos, "sum"), (pr_sym, rnd, "sum"), (pr_iso, rnd, "disp")):
        sh = O.read_shifts(orec, *cat, omesh, f)
        for a in range(3):
            assert maxabs(got[a].cpu().numpy(), cat[a] - sh[a]) < 1e-3
    # host pipeline: same reuse against the library-owned cached mesh
    rec_h = Rec(**kw)
    B.run(rec_h, (n, n, n), *[p.copy() for p in pos], w)
    _, f3 = ctx.launch_counts()
    hd = B.read_shifts(rec_h, *pos, None, field="sum")
    _, f4 = ctx.launch_counts()
    hr = B.read_shifts(rec_h, *rnd, None, field="sum")
    _, f5 = ctx.launch_counts()
    assert f4 > f3 and f5 == f4
    sh = O.read_shifts(orec, *rnd, omesh, "sum")
    for a in range(3):
        assert maxabs(hr[a], sh[a]) < 1e-3
    # a new reconstruction on the context invalidates the kept meshes
    mesh2 = B.run(rec, (n, n, n), *r, dev(np.ones(len(rnd[0]), np.float32)))
    _, f6 = ctx.launch_counts()
    B.reconstructed_positions(rec, *d, field="sum")
    _, f7 = ctx.launch_counts()
    assert f7 - f6 >= 3


def test_unified_sort_is_reused_only_for_the_same_unmodified_catalog(B, O):
    """run! sorts the catalog once into the gather's tile order; read_shifts on the same position
    arrays (same pointers, same content by 64-bit hash, wrapped positions included) reuses that sort.
    Anything else -- other arrays, arrays edited in place, another catalog scattered in between --
    sorts again.  Results agree with the oracle either way."""
    n, L, N = 64, 1000.0, 300_000
    pos, w = clustered_box(N, L, seed=41)
    pos[0][:60] += np.float32(L)                 # exercised: wrap + write-back, hash of the wrapped positions
    kw = dict(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=np.full(3, L, np.float32),
              box_min=np.zeros(3, np.float32), los=(0.0, 0.0, 1.0), n_iter=3)
    orec = O.IterativeRecon(**kw)
    opos = [p.copy() for p in pos]
    omesh = O.run(orec, (n, n, n), *opos, w)
    oshift = O.read_shifts(orec, *opos, omesh, "sum")

    def check(s, ref):
        for a in range(3):
            assert rel_rms(s[a].cpu().numpy() if hasattr(s[a], "cpu") else s[a], ref[a]) < 1e-4
            assert maxabs(s[a].cpu().numpy() if hasattr(s[a], "cpu") else s[a], ref[a]) < 1e-3

    with options(B, bin_min_particles=0):
        ctx = B.Context.get(0)
        d = [dev(p) for p in pos]
        rec = B.IterativeRecon(**kw)
        mesh = B.run(rec, (n, n, n), *d, dev(w))
        c0 = ctx.sort_reuse_count()
        s = B.read_shifts(rec, *d, mesh, field="sum")
        assert ctx.sort_reuse_count() == c0 + 1
        check(s, oshift)
        s = B.read_shifts(rec, *d, mesh, field="sum")                  # the kept sort survives a reuse
        assert ctx.sort_reuse_count() == c0 + 2
        check(s, oshift)
        clones = [t.clone() for t in d]                                 # same content, other arrays: sorted again
        s = B.read_shifts(rec, *clones, mesh, field="sum")
        assert ctx.sort_reuse_count() == c0 + 2
        check(s, oshift)
        # edited in place after run!: the hash no longer matches
        mesh = B.run(rec, (n, n, n), *d, dev(w))
        c1 = ctx.sort_reuse_count()
        d[1].add_(3.0).remainder_(L)
        moved = [t.cpu().numpy() for t in d]
        s = B.read_shifts(rec, *d, mesh, field="sum")
        assert ctx.sort_reuse_count() == c1
        check(s, O.read_shifts(orec, *moved, omesh, "sum"))
        # host pipeline: the read-back catalog is uploaded to the arrays run! used
        rec_h = B.IterativeRecon(**kw)
        hp = [p.copy() for p in pos]
        B.run(rec_h, (n, n, n), *hp, w)
        c2 = ctx.sort_reuse_count()
        sh = B.read_shifts(rec_h, *hp, None, field="sum")
        assert ctx.sort_reuse_count() == c2 + 1
        check(sh, oshift)
        other, _ = clustered_box(N, L, seed=43)
        sh = B.read_shifts(rec_h, *other, None, field="sum")           # same size, different catalog
        assert ctx.sort_reuse_count() == c2 + 1
        check(sh, O.read_shifts(orec, *other, omesh, "sum"))
