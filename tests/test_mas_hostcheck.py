"""CPU: the cell / weight arithmetic the CUDA scatter and gather kernels execute
(baorec.jl_b200/csrc/mas_math.cuh), compiled as plain C++ by tests/hostcheck/ and run on the CPU, against the
oracle -- the bit-exact half of the parity claim (cic! cells / weights / wrapped positions, read_cic! cells /
weights for the reference's CPU and GPU coordinate formulas, TSC stencil), checked where no GPU is available.
The same comparisons run on the device in tests/test_gpu_mas.py; the edge cases are the same."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import baorec_oracle as O
from test_gpu_mas import edge_positions

ROOT = Path(__file__).resolve().parent.parent
f32 = np.float32
_F, _I = C.POINTER(C.c_float), C.POINTER(C.c_int32)


def fp(a):
    return a.ctypes.data_as(_F)


def ip(a):
    return a.ctypes.data_as(_I)


@pytest.fixture(scope="module")
def HC():
    out = ROOT / "tests" / "_build" / "libmas_hostcheck.so"
    src = ROOT / "tests" / "hostcheck" / "mas_hostcheck.cpp"
    hdrs = [ROOT / "baorec.jl_b200" / "csrc" / h for h in ("mas_math.cuh", "host_shim.cuh")]
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in [src] + hdrs):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.run([gxx, "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I", str(ROOT / "include"), "-shared", "-fPIC",
                        "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    i64, i = C.c_int64, C.c_int
    lib.hc_cic_cells.restype = None
    lib.hc_cic_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, i, _I, _I, _F, _F, _F]
    lib.hc_gather_cells.restype = None
    lib.hc_gather_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, _F, i, _I, _I, _F, _F]
    lib.hc_tsc_cells.restype = None
    lib.hc_tsc_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, i, _I, _F, _I]
    lib.hc_deposit.restype = i64
    lib.hc_deposit.argtypes = [i, _F, _F, _F, _F, _F, i64, _I, _F, _F, i, i, i, i, i]
    lib.hc_tsc_gather.restype = i64
    lib.hc_tsc_gather.argtypes = [_F, _F, _F, _F, i64, _I, _F, _F, i, i, i, i, _F]
    lib.hc_pcs_cells.restype = None
    lib.hc_pcs_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, i, _I, _F, _I]
    lib.hc_pcs_gather.restype = i64
    lib.hc_pcs_gather.argtypes = [_F, _F, _F, _F, i64, _I, _F, _F, i, i, i, i, _F]
    lib.hc_deposit_pairs.restype = i64
    lib.hc_deposit_pairs.argtypes = [i, _F, _F, _F, _F, _F, i64, _I, _F, _F, i, C.POINTER(i64)]
    lib.hc_deposit_tsc_vec.restype = i64
    lib.hc_deposit_tsc_vec.argtypes = [_F, _F, _F, _F, _F, i64, _I, _F, _F, i, i, i, i, i]
    lib.hc_deposit_pcs_vec.restype = i64
    lib.hc_deposit_pcs_vec.argtypes = [_F, _F, _F, _F, _F, i64, _I, _F, _F, i, i, i, i, i]
    lib.hc_deposit_fixed.restype = i64
    lib.hc_deposit_fixed.argtypes = [_F, C.POINTER(C.c_uint64), _F, _F, _F, _F, i64, _I, _F, _F, i]
    lib.hc_shifts_epilogue.restype = None
    lib.hc_shifts_epilogue.argtypes = [_F, _F, _F, _F, _F, _F, i64, i, i, i, _F, C.c_float, _F, _F, _F]
    lib.hc_cic_gather.restype = i64
    lib.hc_cic_gather.argtypes = [_F, _F, _F, _F, i64, _I, _F, _F, i, i, i, i, _F]
    return lib


def u32(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("n,L,lo", [(64, 1000.0, 0.0), (96, 2500.0, 0.0), (128, 1373.5, -412.25), (1024, 2500.0, 0.0), (24, 431.7, 7.0)])
@pytest.mark.parametrize("wrap", [True, False])
def test_cic_cells_bit_exact(HC, n, L, lo, wrap):
    x, y, z = (edge_positions(L, lo, n, s) for s in (1, 2, 3))
    if wrap:  # particles beyond the upper face get wrapped by cic! (the test uses axis-1 min/size for every axis: quirk)
        x = np.concatenate([x, f32([lo + L + 0.25, lo + L + 3.0])])
        y = np.concatenate([y, f32([lo + 1.0, lo + L + 7.5])])
        z = np.concatenate([z, f32([lo + L + 11.0, lo + 2.0])])
    # a few particles outside the mesh: flagged (-1), never clamped
    x, y, z = (np.concatenate([a, f32([lo - 5.0, lo + 0.5 * L])]) for a in (x, y, z))
    bs, bm = np.full(3, L, f32), np.full(3, lo, f32)
    N = len(x)
    pos, i0, i1, w0, w1 = O.cic_cells(x, y, z, (n, n, n), bs, bm, wrap)
    gi0, gi1 = np.empty((3, N), np.int32), np.empty((3, N), np.int32)
    gw0, gw1, gwr = np.empty((3, N), f32), np.empty((3, N), f32), np.empty((3, N), f32)
    ng = np.full(3, n, np.int32)
    HC.hc_cic_cells(fp(x), fp(y), fp(z), N, ip(ng), fp(bs), fp(bm), int(wrap), ip(gi0), ip(gi1), fp(gw0), fp(gw1), fp(gwr))
    for a in range(3):
        valid = (i0[a] >= 0) & (i0[a] < n) & (i1[a] < n)
        assert valid.sum() > 4000 and (~valid).sum() >= 1
        assert np.array_equal(gi0[a][valid], i0[a][valid].astype(np.int32))
        assert np.array_equal(gi1[a][valid], i1[a][valid].astype(np.int32))
        assert np.array_equal(u32(gw0[a][valid]), u32(w0[a][valid])) and np.array_equal(u32(gw1[a][valid]), u32(w1[a][valid]))
        assert (gi0[a][~valid] == -1).all()
        if wrap:
            assert np.array_equal(u32(gwr[a]), u32(pos[a]))          # the positions cic! writes back (src/mas.jl:8-10)


@pytest.mark.parametrize("n,L,lo", [(64, 1000.0, 0.0), (128, 1373.5, -412.25), (96, 2500.0, 0.0), (24, 431.7, 7.0)])
@pytest.mark.parametrize("formula", ["cpu", "gpu"])
def test_gather_cells_bit_exact(HC, n, L, lo, formula):
    x, y, z = (edge_positions(L, lo, n, s) for s in (4, 5, 6))
    bs, bm = np.full(3, L, f32), np.full(3, lo, f32)
    cell = np.array([f32(bs[a] / f32(n)) for a in range(3)], f32)             # T(L/n), src/mas.jl:221
    N = len(x)
    idn, iup, wd, wu = O.gather_cells(x, y, z, (n, n, n), bs, bm, True, formula)
    gid, giu, gwd, gwu = np.empty((3, N), np.int32), np.empty((3, N), np.int32), np.empty((3, N), f32), np.empty((3, N), f32)
    HC.hc_gather_cells(fp(x), fp(y), fp(z), N, ip(np.full(3, n, np.int32)), fp(bs), fp(bm), fp(cell), int(formula == "gpu"),
                       ip(gid), ip(giu), fp(gwd), fp(gwu))
    for a in range(3):
        assert np.array_equal(gid[a], idn[a].astype(np.int32)) and np.array_equal(giu[a], iup[a].astype(np.int32))
        assert np.array_equal(u32(gwd[a]), u32(wd[a])) and np.array_equal(u32(gwu[a]), u32(wu[a]))


def test_the_two_gather_formulas_differ_only_off_powers_of_two(HC):
    """src/mas.jl:224 (CPU: (p - min)/cell) vs :274 (GPU: (p - min) n / L): identical for n = 2^k, last-bit different otherwise."""
    for n, differ in ((64, False), (96, True)):
        L, lo = 1373.5, -412.25
        x, y, z = (edge_positions(L, lo, n, s) for s in (7, 8, 9))
        bs, bm = np.full(3, L, f32), np.full(3, lo, f32)
        cell = np.array([f32(bs[a] / f32(n)) for a in range(3)], f32)
        N = len(x)
        out = {}
        for g in (0, 1):
            o = [np.empty((3, N), np.int32), np.empty((3, N), np.int32), np.empty((3, N), f32), np.empty((3, N), f32)]
            HC.hc_gather_cells(fp(x), fp(y), fp(z), N, ip(np.full(3, n, np.int32)), fp(bs), fp(bm), fp(cell), g, ip(o[0]), ip(o[1]), fp(o[2]), fp(o[3]))
            out[g] = o
        assert np.array_equal(u32(out[0][3]), u32(out[1][3])) != differ


@pytest.mark.parametrize("n,L,lo", [(48, 500.0, 0.0), (96, 1373.5, -412.25)])
def test_tsc_cells_bit_exact(HC, n, L, lo):
    x, y, z = (edge_positions(L, lo, n, s) for s in (10, 11, 12))
    bs, bm = np.full(3, L, f32), np.full(3, lo, f32)
    N = len(x)
    ic, wm, wc, wp = O.tsc_cells(x, y, z, (n, n, n), bs, bm, True)
    idx, w, ok = np.empty((3, 3, N), np.int32), np.empty((3, 3, N), f32), np.empty(N, np.int32)
    HC.hc_tsc_cells(fp(x), fp(y), fp(z), N, ip(np.full(3, n, np.int32)), fp(bs), fp(bm), 1, ip(idx), fp(w), ip(ok))
    assert ok.all()
    for a in range(3):
        for o, wref in enumerate((wm[a], wc[a], wp[a])):
            assert np.array_equal(idx[a, o], np.mod(ic[a] + o - 1, n).astype(np.int32))
            assert np.array_equal(u32(w[a, o]), u32(wref))
        assert np.abs(w[a].sum(axis=0) - 1).max() < 3e-7                      # partition of unity


@pytest.mark.parametrize("n,L,lo", [(48, 500.0, 0.0), (96, 1373.5, -412.25)])
def test_pcs_cells_bit_exact(HC, n, L, lo):
    """pcs_axis (csrc/mas_math.cuh) against the oracle's pcs_cells: the four indices and the four cubic B-spline weights
    of every axis, bit for bit, at the faces, the cell boundaries and inside the cells."""
    x, y, z = (edge_positions(L, lo, n, s) for s in (13, 14, 15))
    bs, bm = np.full(3, L, f32), np.full(3, lo, f32)
    N = len(x)
    ic, W = O.pcs_cells(x, y, z, (n, n, n), bs, bm, True)
    idx, w, ok = np.empty((3, 4, N), np.int32), np.empty((3, 4, N), f32), np.empty(N, np.int32)
    HC.hc_pcs_cells(fp(x), fp(y), fp(z), N, ip(np.full(3, n, np.int32)), fp(bs), fp(bm), 1, ip(idx), fp(w), ip(ok))
    assert ok.all()
    for a in range(3):
        for o in range(4):
            assert np.array_equal(idx[a, o], np.mod(ic[a] + o - 1, n).astype(np.int32))
            assert np.array_equal(u32(w[a, o]), u32(W[a][o]))
        assert np.abs(w[a].sum(axis=0) - 1).max() < 4e-7                      # partition of unity
        assert (w[a] >= 0).all()


# ---- the slab-decomposed (multi-GPU) scatter / gather schemes, emulated rank by rank on the CPU -------------------
# Each "rank" runs the product's own deposit<MAS> / tsc_axis / local_plane(s) on its particles into a local buffer
# with ghost planes; numpy then performs the boundary-cell exchange exactly as dist.cu's dist_scatter /
# dist_scatter_tsc do (ring_exchange + add_plane), and the assembled mesh must be the oracle's global one.
CIC, TSC, PCS = 0, 1, 2


def shard(B, pos, w, lo, L, nz, P):
    own = B.dist.owner_of_z(pos[2], lo, L, nz, P)
    assert (own >= 0).all()
    return [([np.ascontiguousarray(p[own == r]) for p in pos], np.ascontiguousarray(w[own == r])) for r in range(P)]


def slab_catalog(n, L, lo, seed, wrap):
    rng = np.random.default_rng(seed)
    pos, w = [], (0.5 + rng.random(30000)).astype(f32)
    for a in range(3):
        frac = rng.random(30000) if wrap else 0.15 + 0.65 * rng.random(30000)     # wrap = false: stencils stay inside the mesh
        p = (lo + L[a] * frac).astype(f32)
        if wrap:                                                                    # include both faces and cell boundaries
            p[:4] = f32([lo, np.nextafter(f32(lo + L[a]), f32(lo)), lo + L[a] / n[a] * 3, lo + L[a] * 0.5])
        np.clip(p, f32(lo), np.nextafter(f32(lo + L[a]), f32(lo)), out=p)
        pos.append(p)
    return pos, w


@pytest.mark.parametrize("P", [1, 2, 4, 8])
@pytest.mark.parametrize("mas,wrap", [(CIC, True), (TSC, True), (TSC, False), (CIC, False), (PCS, True), (PCS, False)])
def test_slab_scatter_scheme_reassembles_the_global_mesh(HC, B, P, mas, wrap):
    n = (12, 10, 16)                                   # (nx, ny, nz); nz divisible by every P with >= 2 planes per rank
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, w = slab_catalog(n, L, lo, 21 + P, wrap)
    if wrap and mas == CIC:                            # cic! wraps with the axis-1 box (quirk): keep the catalog inside it
        pos = [np.minimum(p, np.nextafter(f32(lo + L[0] * 0.83), f32(lo))) if a == 2 else p for a, p in enumerate(pos)]
        pos[1] = np.minimum(pos[1], np.nextafter(f32(lo + L[1]), f32(lo)))
    ref = np.zeros((n[2], n[1], n[0]), f32)
    if mas == CIC:
        O.cic_scatter(ref, *[p.copy() for p in pos], w, bs, bm, wrap)
    elif mas == PCS:
        O.pcs_scatter(ref, *[p.copy() for p in pos], w, bs, bm, wrap)                 # PCS reaches the same planes as TSC: same layout, same exchange
    else:
        O.tsc_scatter(ref, *[p.copy() for p in pos], w, bs, bm, wrap)
    nzl, plane = n[2] // P, n[1] * n[0]
    ghosts_below, ghosts_above = (0, 1) if mas == CIC else (1, 2)
    nzp = nzl + ghosts_below + ghosts_above
    ng = np.asarray(n, np.int32)
    bufs = []
    for r, ((px, py, pz), pw) in enumerate(shard(B, pos, w, lo, float(L[2]), n[2], P)):
        buf = np.zeros((nzp, n[1], n[0]), f32)
        bad = HC.hc_deposit(mas, fp(buf), fp(px), fp(py), fp(pz), fp(pw), len(px), ip(ng), fp(bs), fp(bm), int(wrap), 1, r * nzl,
                            ghosts_below, nzp)
        assert bad == 0                                  # every owned particle finds its whole stencil in the local buffer
        bufs.append(buf)
    out = [b.copy() for b in bufs]
    for r in range(P):                                   # the exchange of dist_scatter (CIC) / dist_scatter_tsc
        nxt, prv = (r + 1) % P, (r - 1) % P
        if mas == CIC:
            out[r][0] += bufs[prv][nzl]                  # ghost plane above -> next rank's first plane
        else:
            out[r][nzl] += bufs[nxt][0]                  # plane below the slab -> previous rank's last real plane
            out[r][1:3] += bufs[prv][nzl + 1:nzl + 3]    # two planes above -> next rank's first two real planes
    got = np.concatenate([o[ghosts_below:ghosts_below + nzl] for o in out])
    assert np.abs(got - ref).max() <= 5e-6 * float(ref.max())                       # summation order only
    assert abs(float(got.sum(dtype=np.float64)) - float(w.sum(dtype=np.float64))) < (5e-2 if mas == PCS else 1e-2)   # 64 Float32 products per particle (PCS), 27, 8


@pytest.mark.parametrize("P", [1, 2, 4, 8])
@pytest.mark.parametrize("n,Lz", [((12, 10, 16), 400.0), ((12, 10, 48), 431.7)])
def test_cic_slab_gather_scheme(HC, B, P, n, Lz):
    """The halo layout of read_shifts_dist (one plane below the slab, two above) is enough for every owned particle --
    including those whose gather coordinate (p - min)/cell rounds into the cell below or above the scatter's
    (p - min) n / L (src/mas.jl:13 vs :224) -- and the result is the oracle's read_cic! bit for bit."""
    L, lo = f32([300.0, 250.0, Lz]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, _ = slab_catalog(n, L, lo, 51 + P, True)
    cellz = f32(L[2] / f32(n[2]))
    pos[2][4:4 + n[2]] = (lo + cellz * np.arange(n[2])).astype(f32)           # exactly on the planes: the last-bit cases
    fld = np.random.default_rng(6).standard_normal((n[2], n[1], n[0])).astype(f32)
    ref = O.read_cic(fld, *pos, bs, bm, True, "cpu")
    nzl = n[2] // P
    own = B.dist.owner_of_z(pos[2], lo, float(L[2]), n[2], P)
    ng = np.asarray(n, np.int32)
    for r in range(P):
        buf = np.ascontiguousarray(fld[[(r * nzl - 1 + k) % n[2] for k in range(nzl + 3)]])
        sel = own == r
        px, py, pz = (np.ascontiguousarray(p[sel]) for p in pos)
        out = np.empty(len(px), f32)
        assert HC.hc_cic_gather(fp(buf), fp(px), fp(py), fp(pz), len(px), ip(ng), fp(bs), fp(bm), 1, r * nzl, 1, nzl + 3, fp(out)) == 0
        assert np.array_equal(u32(out), u32(ref[sel]))


@pytest.mark.parametrize("P", [1, 2, 4, 8])
@pytest.mark.parametrize("mas", [TSC, PCS])
def test_tsc_slab_gather_scheme(HC, B, P, mas):
    n = (12, 10, 16)
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, _ = slab_catalog(n, L, lo, 31 + P, True)
    fld = np.random.default_rng(5).standard_normal((n[2], n[1], n[0])).astype(f32)
    ref = (O.read_tsc if mas == TSC else O.read_pcs)(fld, *pos, bs, bm, True)
    hc_gather = HC.hc_tsc_gather if mas == TSC else HC.hc_pcs_gather
    nzl = n[2] // P
    own = B.dist.owner_of_z(pos[2], lo, float(L[2]), n[2], P)
    ng = np.asarray(n, np.int32)
    for r in range(P):
        # the halo layout read_shifts_dist builds: plane 0 <- previous rank's last, planes nzl+1, nzl+2 <- next rank's first two
        planes = [(r * nzl - 1 + k) % n[2] for k in range(nzl + 3)]
        buf = np.ascontiguousarray(fld[planes])
        sel = own == r
        px, py, pz = (np.ascontiguousarray(p[sel]) for p in pos)
        out = np.empty(len(px), f32)
        bad = hc_gather(fp(buf), fp(px), fp(py), fp(pz), len(px), ip(ng), fp(bs), fp(bm), 1, r * nzl, 1, nzl + 3, fp(out))
        assert bad == 0
        assert np.array_equal(u32(out), u32(ref[sel]))                              # bit-exact, like the single-GPU gather
    # a particle of another slab is rejected (counted out-of-box), never read from the wrong plane
    if P > 2:
        other = own == 2
        out = np.empty(int(other.sum()), f32)
        buf = np.ascontiguousarray(fld[[(0 * nzl - 1 + k) % n[2] for k in range(nzl + 3)]])
        px, py, pz = (np.ascontiguousarray(p[other]) for p in pos)
        assert hc_gather(fp(buf), fp(px), fp(py), fp(pz), len(px), ip(ng), fp(bs), fp(bm), 1, 0, 1, nzl + 3, fp(out)) > 0


def test_single_gpu_deposit_matches_the_oracle(HC):
    """slab = 0: the whole-mesh deposit<CIC> in catalog order IS the reference's serial loop (bit-identical); the
    oracle's TSC scatter sums stencil offset by stencil offset, so deposit<TSC> agrees to summation order."""
    n = (12, 10, 16)
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, w = slab_catalog(n, L, lo, 41, False)
    ng = np.asarray(n, np.int32)
    for mas, fn in ((CIC, O.cic_scatter), (TSC, O.tsc_scatter), (PCS, O.pcs_scatter)):
        ref = fn(np.zeros((n[2], n[1], n[0]), f32), *[p.copy() for p in pos], w, bs, bm, False)
        buf = np.zeros_like(ref)
        assert HC.hc_deposit(mas, fp(buf), fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(w), len(w), ip(ng), fp(bs), fp(bm), 0, 0, 0, 0, n[2]) == 0
        if mas == CIC:
            assert np.array_equal(u32(buf), u32(ref))
        else:
            assert np.abs(buf - ref).max() <= 5e-6 * float(ref.max())


@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (0.6, 0.0, 0.8), None])
def test_read_shifts_epilogue_bit_exact(HC, los):
    """read_shifts' :disp / :rsd / :sum arithmetic (src/recon.jl:277-304) and reconstructed_positions' pos - shift
    (:377), as the gather kernels apply it per particle, against the oracle's -- same operations, same order."""
    rng = np.random.default_rng(7)
    N = 20000
    pos = [(900.0 + 600.0 * rng.random(N)).astype(f32) for _ in range(3)]
    disp = [(5.0 * rng.standard_normal(N)).astype(f32) for _ in range(3)]

    class Rec:                       # what oracle.read_shifts needs besides the displacements
        f = 0.757
        los = None

    Rec.los = los
    fgrowth = f32(0.757)
    for code, field in enumerate(("disp", "rsd", "sum")):
        # the oracle's epilogue, fed with the same displacement values
        if field == "disp":
            ref = disp
        else:
            if los is None:
                dist = np.sqrt(((pos[0] * pos[0]) + (pos[1] * pos[1])).astype(f32) + (pos[2] * pos[2])).astype(f32)
                l = [(p / dist).astype(f32) for p in pos]
            else:
                l = [f32(v) for v in los]
            dot = (((disp[0] * l[0]) + (disp[1] * l[1])).astype(f32) + (disp[2] * l[2])).astype(f32)
            rsd = [((fgrowth * dot).astype(f32) * li).astype(f32) for li in l]
            ref = rsd if field == "rsd" else [(d + r).astype(f32) for d, r in zip(disp, rsd)]
        for positions in (0, 1):
            want = [(p - s).astype(f32) for p, s in zip(pos, ref)] if positions else ref
            out = [np.empty(N, f32) for _ in range(3)]
            HC.hc_shifts_epilogue(fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(disp[0]), fp(disp[1]), fp(disp[2]), N, code, positions,
                                  int(los is not None), fp(np.asarray(los if los is not None else (0, 0, 0), f32)), fgrowth,
                                  fp(out[0]), fp(out[1]), fp(out[2]))
            for a in range(3):
                assert np.array_equal(u32(out[a]), u32(np.broadcast_to(want[a], (N,)).astype(f32)))


@pytest.mark.parametrize("wrap", [True, False])
def test_deterministic_scatter_is_order_independent_and_correctly_rounded(HC, wrap):
    """Option "deterministic_scatter": 2^-40 fixed-point integer accumulation + ONE rounding to Float32.  Any particle
    order gives the same bits; the result is the correctly rounded sum of the reference's Float32 deposit values
    (all but never off by the bits dropped below 2^-40), hence within summation-order distance of the serial loop."""
    n = (12, 10, 16)
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, w = slab_catalog(n, L, lo, 61, False)
    w = (w * f32(3.7)).astype(f32)
    ng = np.asarray(n, np.int32)
    M = n[0] * n[1] * n[2]

    def run(order):
        rho, acc = np.zeros(M, f32), np.zeros(M, np.uint64)
        p = [np.ascontiguousarray(a[order]) for a in pos]
        bad = HC.hc_deposit_fixed(fp(rho), acc.ctypes.data_as(C.POINTER(C.c_uint64)), fp(p[0]), fp(p[1]), fp(p[2]),
                                  fp(np.ascontiguousarray(w[order])), len(w), ip(ng), fp(bs), fp(bm), int(wrap))
        assert bad == 0
        return rho.reshape(n[2], n[1], n[0])

    N = len(w)
    a = run(np.arange(N))
    rng = np.random.default_rng(0)
    for _ in range(3):
        assert np.array_equal(u32(run(rng.permutation(N))), u32(a))          # bit-reproducible in any order
    ref = O.cic_scatter(np.zeros((n[2], n[1], n[0]), f32), *[p.copy() for p in pos], w, bs, bm, wrap)
    assert np.abs(a - ref).max() <= 2e-6 * float(ref.max())
    # exact sum of the very same Float32 deposit values, in Float64, rounded once
    _, i0, i1, w0, w1 = O.cic_cells(*[p.copy() for p in pos], n, bs, bm, wrap)
    exact = np.zeros(M, np.float64)
    wx = ((w0[0] * w).astype(f32), (w1[0] * w).astype(f32))
    ix, iy, iz = (i0[0], i1[0]), (i0[1], i1[1]), (i0[2], i1[2])
    wy, wz = (w0[1], w1[1]), (w0[2], w1[2])
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                v = ((wx[cx] * wy[cy]).astype(f32) * wz[cz]).astype(f32)
                np.add.at(exact, (iz[cz] * n[1] + iy[cy]) * n[0] + ix[cx], v.astype(np.float64))
    exact32 = exact.astype(f32).reshape(a.shape)
    assert (a == exact32).mean() > 0.999 and np.abs(a - exact32).max() <= np.spacing(exact32.max())
    assert abs(float(a.sum(dtype=np.float64)) - float(w.sum(dtype=np.float64))) < 2e-3


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,wrap", [((12, 10, 16), True), ((12, 10, 16), False), ((13, 8, 8), True), ((14, 8, 8), True)])
def test_paired_deposit_equals_the_scalar_one(HC, n, wrap, mode):
    """Option "scatter_pairs": the aligned x pairs of cic! through one vector reduction (red.global.add.v2.f32 on the
    device, two additions on the host).  Same cells, same values, same particle order -> the serial result is the
    reference's loop bit for bit; about half of the particles take the paired path on an even mesh, none on an odd one."""
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, w = slab_catalog(n, L, lo, 67, wrap)       # wrap: both faces and the periodic pair (nx - 1, 0) included
    ng = np.asarray(n, np.int32)
    M = n[0] * n[1] * n[2]
    got, paired = np.zeros(M, f32), C.c_int64(0)
    bad = HC.hc_deposit_pairs(mode, fp(got), fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(w), len(w), ip(ng), fp(bs), fp(bm), int(wrap), C.byref(paired))
    ref = O.cic_scatter(np.zeros((n[2], n[1], n[0]), f32), *[p.copy() for p in pos], w, bs, bm, wrap)
    assert bad == 0 and np.array_equal(u32(got.reshape(ref.shape)), u32(ref))
    # share of the particles that take a vector reduction: pairs need an even row, quads (mode 2) a row of 4 k cells
    want = 0.0 if n[0] % 2 else (1.0 if (mode == 2 and n[0] % 4 == 0) else 0.5)
    assert abs(paired.value / len(w) - want) < 0.1


@pytest.mark.parametrize("n,wrap", [((16, 10, 12), True), ((16, 10, 12), False), ((12, 8, 8), True), ((14, 8, 8), True)])
def test_vector_tsc_deposit_equals_the_scalar_one(HC, n, wrap):
    """Option "scatter_pairs" for TSC: one aligned quad or two aligned pairs per stencil row (13.5 reductions per
    particle instead of 27).  Against the product's own scalar deposit<TSC> in the same particle order: identical bits
    (rows that wrap in x, and the 14-cell mesh whose rows are not a multiple of 4, take the scalar path)."""
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, w = slab_catalog(n, L, lo, 71, wrap)
    ng = np.asarray(n, np.int32)
    M = n[0] * n[1] * n[2]
    a, b = np.zeros(M, f32), np.zeros(M, f32)
    TSC = 1
    bad_a = HC.hc_deposit(TSC, fp(a), fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(w), len(w), ip(ng), fp(bs), fp(bm), int(wrap), 0, 0, 0, n[2])
    bad_b = HC.hc_deposit_tsc_vec(fp(b), fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(w), len(w), ip(ng), fp(bs), fp(bm), int(wrap), 0, 0, 0, n[2])
    assert bad_a == bad_b and np.array_equal(u32(a), u32(b))
    assert abs(float(b.sum(dtype=np.float64)) / float(w.sum(dtype=np.float64)) - 1) < 1e-5 or bad_b > 0


@pytest.mark.parametrize("n,wrap", [((16, 10, 12), True), ((16, 10, 12), False), ((12, 8, 8), True), ((14, 8, 8), True)])
def test_vector_pcs_deposit_equals_the_scalar_one(HC, n, wrap):
    """Option "scatter_pairs" for PCS: one or two aligned quads per stencil row (16 - 32 reductions per particle instead
    of 64).  Against the product's own scalar deposit<PCS> in the same particle order: identical bits (rows that wrap in
    x, and the 14-cell mesh whose rows are not a multiple of 4, take the scalar path)."""
    L, lo = f32([300.0, 250.0, 400.0]), -50.0
    bs, bm = L, np.full(3, lo, f32)
    pos, w = slab_catalog(n, L, lo, 73, wrap)
    ng = np.asarray(n, np.int32)
    M = n[0] * n[1] * n[2]
    a, b = np.zeros(M, f32), np.zeros(M, f32)
    bad_a = HC.hc_deposit(PCS, fp(a), fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(w), len(w), ip(ng), fp(bs), fp(bm), int(wrap), 0, 0, 0, n[2])
    bad_b = HC.hc_deposit_pcs_vec(fp(b), fp(pos[0]), fp(pos[1]), fp(pos[2]), fp(w), len(w), ip(ng), fp(bs), fp(bm), int(wrap), 0, 0, 0, n[2])
    assert bad_a == bad_b and np.array_equal(u32(a), u32(b))
    assert abs(float(b.sum(dtype=np.float64)) / float(w.sum(dtype=np.float64)) - 1) < 1e-5 or bad_b > 0
