"""GPU: size-independent properties at BASELINE.json's full sizes, where the oracle cannot follow
(configs[3]: IterativeRecon periodic box 1024^3 / 1e8 particles; configs[1]/[2]: 512^3 lightcone-sized meshes with
TSC and with the multigrid solver).  What is checked needs no reference run: conservation, invariances and
linearity of the operators, agreement between independent code paths of the engine (tile-sorted gather vs the
catalog-order gather; cached vs recomputed read-back), exact identities between the entry points.
(Written at the end of round 1 without GPU time left, hence the file name that sorts late; green on a B200 since round 2.)"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

L_BOX, N_MESH, N_PART = 2500.0, 1024, 100_000_000
KW = dict(bias=2.2, f=0.757, smoothing_radius=15.0, n_iter=3, los=(0.0, 0.0, 1.0))


def need_memory(gib):
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip(f"needs {gib} GiB of free device memory")


def uniform(N, L, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    top = float(np.nextafter(np.float32(L), np.float32(0)))
    return [(torch.rand(N, device="cuda", generator=g) * L).clamp_(max=top) for _ in range(3)]


def rel_rms(a, b):
    return float(((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt().clamp_min(1e-300)).item())


@pytest.fixture(scope="module")
def big(B):
    """One reconstruction of the bench workload, shared by the property tests below."""
    need_memory(90)
    pos = uniform(N_PART, L_BOX, 42)
    w = torch.ones(N_PART, device="cuda")
    kw = dict(KW, box_size=np.full(3, L_BOX, np.float32), box_min=np.zeros(3, np.float32))
    rec = B.IterativeRecon(**kw)
    mesh = B.run(rec, (N_MESH,) * 3, *pos, w)
    shifts = B.read_shifts(rec, *pos, mesh, field="sum")
    yield dict(B=B, pos=pos, w=w, kw=kw, rec=rec, mesh=mesh, shifts=shifts)
    torch.cuda.empty_cache()


def test_scatter_conserves_mass_at_full_size(B):
    need_memory(30)
    pos = uniform(N_PART, L_BOX, 7)
    g = torch.Generator(device="cuda").manual_seed(8)
    w = 0.5 + torch.rand(N_PART, device="cuda", generator=g)
    bs, bm = np.full(3, L_BOX, np.float32), np.zeros(3, np.float32)
    before = [p.clone() for p in pos]
    rho = torch.zeros((N_MESH,) * 3, dtype=torch.float32, device="cuda")
    B.cic(rho, *pos, w, bs, bm, wrap=True)
    total, want = float(rho.sum(dtype=torch.float64)), float(w.sum(dtype=torch.float64))
    assert abs(total / want - 1) < 1e-6                      # sum of 8e8 rounded deposits vs sum of weights
    assert float(rho.min()) >= 0 and float(rho.max()) < 40    # 0.09 particles per cell: nothing piles up
    for a, b in zip(pos, before):                             # in-box positions are written back unchanged (src/mas.jl:8-10)
        assert torch.equal(a, b)
    del rho, pos, before, w
    torch.cuda.empty_cache()


def test_overdensity_has_zero_mean_and_is_reproducible(big):
    B, mesh = big["B"], big["mesh"]
    assert abs(float(mesh.mean(dtype=torch.float64))) < 1e-6          # (rho/mean - 1)/bias: the DC mode is zeroed in k space
    assert 1e-3 < float(mesh.double().pow(2).mean().sqrt()) < 1.0      # shot noise smoothed on 15 Mpc/h: ~0.04
    rec2 = B.IterativeRecon(**big["kw"])
    mesh2 = B.run(rec2, (N_MESH,) * 3, *big["pos"], big["w"])
    assert rel_rms(mesh2, mesh) < 1e-5                                 # float reductions commute only up to rounding
    del mesh2
    torch.cuda.empty_cache()


def test_overdensity_does_not_depend_on_the_weight_normalisation(big):
    """delta = (rho / mean(rho) - 1) / bias: doubling every weight doubles every deposit exactly and cancels."""
    B = big["B"]
    rec2 = B.IterativeRecon(**big["kw"])
    mesh2 = B.run(rec2, (N_MESH,) * 3, *big["pos"], big["w"] * 2.0)
    assert rel_rms(mesh2, big["mesh"]) < 1e-5
    del mesh2
    torch.cuda.empty_cache()


def test_positions_are_positions_minus_shifts_bit_for_bit(big):
    B, rec, pos = big["B"], big["rec"], big["pos"]
    # both read-backs in the SAME cache state: the fixture's shifts came from the delta_k run! kept, but the runs of
    # the tests above have since replaced that cache, and R2C(mesh) differs from the kept delta_k by rounding --
    # so the shifts are taken again here, through the same path as the positions (forward transform of the mesh)
    shifts = B.read_shifts(rec, *pos, big["mesh"], field="sum")
    new = B.reconstructed_positions(rec, *pos, field="sum")
    for a in range(3):
        assert torch.equal(new[a], pos[a] - shifts[a])                 # src/recon.jl:377, same Float32 subtraction
        assert float((shifts[a] - big["shifts"][a]).abs().max()) < 1e-3   # kept delta_k vs R2C(mesh): rounding only
    del shifts
    assert float(big["shifts"][0].abs().max()) < 50 and float(big["shifts"][2].abs().max()) < 100
    s = torch.stack([t.double().pow(2).mean().sqrt() for t in big["shifts"]])
    # fixed line of sight (0,0,1): x and y are statistically equivalent; with field = :sum the z component is
    # (1 + f) Psi_z, and Psi_z itself is damped by the RSD iteration (delta_r ~ delta_s / (1 + beta mu^2)), so the
    # ratio lies between 1 and 1 + f
    assert abs(float(s[0] / s[1]) - 1) < 0.05 and 1.15 < float(s[2] / s[0]) < (1 + KW["f"]) * 1.02
    del new


def test_tile_sorted_gather_agrees_with_the_catalog_order_gather(big):
    """1e8 particles go through the unified sort + shared-memory tile gather; a 1e5 subset of the same particles,
    read back on its own, takes the catalog-order kernel.  Same displacement meshes, same arithmetic: same shifts."""
    B, rec, pos, mesh = big["B"], big["rec"], big["pos"], big["mesh"]
    g = torch.Generator(device="cuda").manual_seed(9)
    idx = torch.randint(0, N_PART, (100_000,), device="cuda", generator=g)
    sub = [p[idx].contiguous() for p in pos]
    for field in ("disp", "sum"):
        full = B.read_shifts(rec, *pos, mesh, field=field)
        part = B.read_shifts(rec, *sub, mesh, field=field)
        for a in range(3):
            assert float((full[a][idx] - part[a]).abs().max()) < 1e-4   # Mpc/h; shifts are O(1-10)
        del full, part


def test_cached_read_back_equals_a_recomputed_one(big):
    """read_shifts against recon.result_cache reuses delta_k and the displacement meshes; against a copy of the
    mesh it redoes the forward transform and the three inverse ones (src/iterative.jl:236-246)."""
    B, rec, pos = big["B"], big["rec"], big["pos"]
    sub = [p[:1_000_000].contiguous() for p in pos]
    cached = B.read_shifts(rec, *sub, big["mesh"], field="sum")
    copy = big["mesh"].clone()
    redo = B.read_shifts(rec, *sub, copy, field="sum")
    for a in range(3):
        assert float((cached[a] - redo[a]).abs().max()) < 1e-3 and rel_rms(redo[a], cached[a]) < 1e-4
    del copy


def test_smoothing_is_linear_at_full_size(B):
    need_memory(40)
    g = torch.Generator(device="cuda").manual_seed(11)
    f = torch.rand((N_MESH,) * 3, device="cuda", generator=g)
    bs = np.full(3, L_BOX, np.float32)
    a = f.clone()
    B.smooth(a, 15.0, bs)
    b = f * 2.0
    B.smooth(b, 15.0, bs)
    assert rel_rms(b, a * 2.0) < 1e-6                               # scaling by 2 is exact in every Float32 operation
    assert abs(float(a.mean(dtype=torch.float64)) / float(f.mean(dtype=torch.float64)) - 1) < 1e-5   # the kernel integrates to 1
    assert float(a.var()) < 0.05 * float(f.var())                    # 15 Mpc/h on 2.44 Mpc/h cells averages ~1000 cells
    del a, b, f
    torch.cuda.empty_cache()


def test_tsc_scatter_conserves_mass_on_the_lightcone_mesh(B):
    """configs[1] size: 512^3, 5e6 data + 5e7 randoms, TSC (27-point stencil, weights sum to 1)."""
    need_memory(20)
    n, L = 512, 3000.0
    bs, bm = np.full(3, L, np.float32), np.full(3, -200.0, np.float32)
    g = torch.Generator(device="cuda").manual_seed(12)
    for N in (5_000_000, 50_000_000):
        pos = [(-200.0 + L * (0.05 + 0.9 * torch.rand(N, device="cuda", generator=g))) for _ in range(3)]
        w = 1.0 / (1.0 + 0.2 * torch.rand(N, device="cuda", generator=g))
        rho = torch.zeros((n,) * 3, dtype=torch.float32, device="cuda")
        B.cic(rho, *pos, w, bs, bm, wrap=False, mas="tsc")
        assert abs(float(rho.sum(dtype=torch.float64)) / float(w.sum(dtype=torch.float64)) - 1) < 1e-6
        del rho, pos, w
    torch.cuda.empty_cache()


@pytest.mark.parametrize("los,lo", [((0.0, 0.0, 1.0), 0.0), (None, 900.0)])
def test_multigrid_is_linear_and_converges_at_512(B, los, lo):
    """configs[2] size.  Every multigrid operator is linear in (v, f): fmg(2 f) = 2 fmg(f); a V-cycle from zero
    reduces the residual, the full multigrid solve reduces it by orders of magnitude (the oracle shows 0.1-0.25 and
    < 1e-3 at 32^3 ... 128^3: the ratios do not depend on the mesh size, that is the point of multigrid)."""
    need_memory(20)
    n, L = 512, 2500.0
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    g = torch.Generator(device="cuda").manual_seed(13)
    f = torch.randn((n,) * 3, device="cuda", generator=g)
    B.smooth(f, 15.0, bs)
    f -= f.mean()
    beta = 0.757 / 2.2

    def residual_norm(v):
        r = torch.empty_like(v)
        B.residual(r, v, f, None, bs, bm, beta, los=los)
        return float(r.double().pow(2).sum().sqrt())

    nf = float(f.double().pow(2).sum().sqrt())
    v = torch.zeros_like(f)
    B.vcycle(v, f, bs, bm, beta, 0.4, 5, los=los)
    assert residual_norm(v) < 0.6 * nf
    phi = torch.zeros_like(f)
    B.fmg(f, phi, bs, bm, beta, 0.4, 5, 6, los=los)
    assert residual_norm(phi) < 0.02 * nf
    phi2 = torch.zeros_like(f)
    B.fmg(f * 2.0, phi2, bs, bm, beta, 0.4, 5, 6, los=los)
    assert rel_rms(phi2, phi * 2.0) < 1e-5
    del v, phi, phi2, f
    torch.cuda.empty_cache()
