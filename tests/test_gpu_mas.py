"""GPU parity: mass assignment (cic!/read_cic!) through the C ABI against the oracle.
Bit-exact: cell indices and interpolation weights, wrapped positions, the gather given
the same field.  Tolerance: the scattered mesh (float atomics -> summation order)."""
import numpy as np
import pytest

from util import uniform_box, clustered_box, maxabs

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def edge_positions(L, lo, n, seed=1):
    """Random positions plus the edge cases of src/mas.jl:7-35: exactly on grid points, on both
    box faces, one ulp inside, and (for wrap) just beyond the upper face."""
    rng = np.random.default_rng(seed)
    cell = L / n
    base = (lo + L * rng.random(4000)).astype(np.float32)
    grid = (lo + cell * rng.integers(0, n, 500)).astype(np.float32)
    special = np.array([lo, np.nextafter(np.float32(lo + L), np.float32(lo)), lo + L - cell, lo + L - 0.5 * cell,
                        lo + cell * 0.5, lo + cell * (n - 1)], dtype=np.float32)
    return np.concatenate([base, grid, special]).astype(np.float32)


@pytest.mark.parametrize("n,L,lo", [(64, 1000.0, 0.0), (96, 2500.0, 0.0), (128, 1373.5, -412.25)])
@pytest.mark.parametrize("wrap", [True, False])
def test_cic_cells_bit_exact(B, O, n, L, lo, wrap):
    x = edge_positions(L, lo, n, 1)
    y = edge_positions(L, lo, n, 2)
    z = edge_positions(L, lo, n, 3)
    if wrap:  # particles beyond the upper face get wrapped by cic! (lower face uses axis-1 min: quirk)
        x = np.concatenate([x, np.float32([lo + L + 0.25, lo + L + 3.0])])
        y = np.concatenate([y, np.float32([lo + 1.0, lo + L + 7.5])])
        z = np.concatenate([z, np.float32([lo + L + 11.0, lo + 2.0])])
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    pos, i0, i1, w0, w1 = O.cic_cells(x, y, z, (n, n, n), bs, bm, wrap)
    g = B.cic_cells((n, n, n), dev(x), dev(y), dev(z), bs, bm, wrap)
    gi0, gi1, gw0, gw1 = (t.cpu().numpy() for t in g)
    for a in range(3):
        valid = (i0[a] >= 0) & (i0[a] < n) & (i1[a] < n)
        assert valid.sum() > 4000
        assert np.array_equal(gi0[a][valid], i0[a][valid].astype(np.int32))
        assert np.array_equal(gi1[a][valid], i1[a][valid].astype(np.int32))
        assert np.array_equal(gw0[a][valid].view(np.uint32), w0[a][valid].view(np.uint32))
        assert np.array_equal(gw1[a][valid].view(np.uint32), w1[a][valid].view(np.uint32))
        assert (gi0[a][~valid] == -1).all()      # out-of-box flagged, not silently clamped


@pytest.mark.parametrize("n,L,lo", [(64, 1000.0, 0.0), (128, 1373.5, -412.25)])
@pytest.mark.parametrize("formula", ["cpu", "gpu"])
def test_gather_cells_bit_exact(B, O, n, L, lo, formula):
    x, y, z = (edge_positions(L, lo, n, s) for s in (4, 5, 6))
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    idn, iup, wd, wu = O.gather_cells(x, y, z, (n, n, n), bs, bm, True, formula)
    g = B.gather_cells((n, n, n), dev(x), dev(y), dev(z), bs, bm, gpu_formula=(formula == "gpu"))
    gid, giu, gwd, gwu = (t.cpu().numpy() for t in g)
    for a in range(3):
        assert np.array_equal(gid[a], idn[a].astype(np.int32))
        assert np.array_equal(giu[a], iup[a].astype(np.int32))
        assert np.array_equal(gwd[a].view(np.uint32), wd[a].view(np.uint32))
        assert np.array_equal(gwu[a].view(np.uint32), wu[a].view(np.uint32))


@pytest.mark.parametrize("maker", [uniform_box, clustered_box])
@pytest.mark.parametrize("wrap", [True, False])
def test_cic_scatter_mesh(B, O, maker, wrap):
    n, L, N = 64, 1000.0, 200_000
    pos, w = maker(N, L, seed=11)
    if wrap:   # push a few particles beyond the upper face: cic! wraps them and writes them back
        pos[0][:50] += np.float32(L)
        pos[2][50:80] += np.float32(L)
    else:      # wrap=false needs x0 <= n-1
        for p in pos:
            np.clip(p, 0, np.float32(L - L / n - 1e-3), out=p)
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    ox, oy, oz = (p.copy() for p in pos)
    orho = O.cic_scatter(np.zeros((n, n, n), np.float32), ox, oy, oz, w, bs, bm, wrap)
    gx, gy, gz = (dev(p) for p in pos)
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.cic(rho, gx, gy, gz, dev(w), bs, bm, wrap=wrap)
    rho = rho.cpu().numpy()
    # serial (oracle) vs atomic (GPU) summation order: a few ulp of the cell value
    assert maxabs(rho, orho) <= 2e-5 * max(1.0, float(orho.max()))
    assert abs(float(rho.sum(dtype=np.float64)) - float(w.sum(dtype=np.float64))) < 1e-3 * N ** 0.5
    # positions are mutated exactly like the reference (src/mas.jl:8-10)
    for g, o in zip((gx, gy, gz), (ox, oy, oz)):
        assert np.array_equal(g.cpu().numpy().view(np.uint32), o.view(np.uint32))


def test_scatter_out_of_box_is_an_error(B):
    n, L = 32, 100.0
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    x = dev(np.float32([10.0, -5.0, 50.0]))
    y = dev(np.float32([10.0, 20.0, 50.0]))
    z = dev(np.float32([10.0, 20.0, 350.0]))
    w = dev(np.ones(3, np.float32))
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    with pytest.raises(B.OutOfBoxError):
        B.cic(rho, x, y, z, w, bs, bm, wrap=True)
    # the in-box particle was still deposited, the others skipped
    assert abs(float(rho.sum()) - 1.0) < 1e-6
    with pytest.raises(B.OutOfBoxError):   # wrap=false and base cell == n (reference: BoundsError)
        B.cic(rho, dev(np.float32([99.9])), dev(np.float32([1.0])), dev(np.float32([1.0])), dev(np.ones(1, np.float32)),
              bs, bm, wrap=False)


def test_empty_catalog(B):
    n, L = 32, 100.0
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    e = torch.empty(0, dtype=torch.float32, device="cuda")
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.cic(rho, e, e, e, e, bs, bm)
    assert float(rho.abs().max()) == 0.0
    out = torch.empty(0, dtype=torch.float32, device="cuda")
    B.read_cic(out, rho, e, e, e, bs, bm)


@pytest.mark.parametrize("n,L,lo", [(64, 1000.0, 0.0), (96, 1373.5, -412.25)])
def test_read_cic_bit_exact(B, O, n, L, lo):
    rng = np.random.default_rng(5)
    fld = rng.standard_normal((n, n, n)).astype(np.float32)
    x, y, z = (edge_positions(L, lo, n, s) for s in (7, 8, 9))
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    ref = O.read_cic(fld, x, y, z, bs, bm)
    out = torch.empty(len(x), dtype=torch.float32, device="cuda")
    B.read_cic(out, dev(fld), dev(x), dev(y), dev(z), bs, bm)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("mas", ["tsc", "pcs"])
@pytest.mark.parametrize("N", [100_000, 400_000])                             # catalog-order kernels / z-binned kernels
def test_stencil_scatter_gather(B, O, mas, N):
    """TSC (27 cells) and PCS (64 cells, the cubic B-spline): the mesh agrees with the oracle's to the order
    of the Float32 additions, the gather is bit-identical (same cells, same weights, same order)."""
    n, L = 48, 500.0
    pos, w = clustered_box(N, L, seed=3)
    bs, bm = np.full(3, L, np.float32), np.zeros(3, np.float32)
    scatter, read = (O.tsc_scatter, O.read_tsc) if mas == "tsc" else (O.pcs_scatter, O.read_pcs)
    orho = scatter(np.zeros((n, n, n), np.float32), *pos, w, bs, bm, True)
    rho = torch.zeros((n, n, n), dtype=torch.float32, device="cuda")
    B.cic(rho, *(dev(p) for p in pos), dev(w), bs, bm, wrap=True, mas=mas)
    assert maxabs(rho.cpu().numpy(), orho) <= 2e-5 * float(orho.max())
    assert abs(float(rho.sum(dtype=torch.float64)) / float(w.sum(dtype=np.float64)) - 1) < 2e-6
    fld = np.random.default_rng(1).standard_normal((n, n, n)).astype(np.float32)
    ref = read(fld, *pos, bs, bm)
    out = torch.empty(N, dtype=torch.float32, device="cuda")
    B.read_cic(out, dev(fld), *(dev(p) for p in pos), bs, bm, mas=mas)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_pcs_cells_at_the_edges(B, O):
    """PCS gather at the faces and cell boundaries of the box (edge_positions): bit-identical to the oracle."""
    n, L, lo = 48, 500.0, -120.5
    x, y, z = (edge_positions(L, lo, n, s) for s in (16, 17, 18))
    bs, bm = np.full(3, L, np.float32), np.full(3, lo, np.float32)
    fld = np.random.default_rng(2).standard_normal((n, n, n)).astype(np.float32)
    ref = O.read_pcs(fld, x, y, z, bs, bm)
    out = torch.empty(len(x), dtype=torch.float32, device="cuda")
    B.read_cic(out, dev(fld), dev(x), dev(y), dev(z), bs, bm, mas="pcs")
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))


def test_setup_box(B, O):
    rng = np.random.default_rng(2)
    pos = [(rng.standard_normal(100_001) * s + c).astype(np.float32) for s, c in ((300, 2000), (500, -100), (200, 50))]
    obs, obm = O.setup_box(*pos, np.float32(500))
    bs, bm = B.setup_box(*(dev(p) for p in pos), 500.0)
    assert np.array_equal(bs, obs) and np.array_equal(bm, obm)
