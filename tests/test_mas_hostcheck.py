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
    hdr = ROOT / "baorec.jl_b200" / "csrc" / "mas_math.cuh"
    if not out.exists() or out.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.run([gxx, "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    i64, i = C.c_int64, C.c_int
    lib.hc_cic_cells.restype = None
    lib.hc_cic_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, i, _I, _I, _F, _F, _F]
    lib.hc_gather_cells.restype = None
    lib.hc_gather_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, _F, i, _I, _I, _F, _F]
    lib.hc_tsc_cells.restype = None
    lib.hc_tsc_cells.argtypes = [_F, _F, _F, i64, _I, _F, _F, i, _I, _F, _I]
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
