"""CPU: the per-mode k-space operators the CUDA kernels apply (baorec.jl_b200/csrc/kspace_ops.cuh), compiled as
plain C++ by tests/hostcheck/ and applied to whole k-space meshes on the CPU (numpy does the transforms), against
the oracle.  The point of it: FusedLosOp folds the smoothing, the (rho/mean - 1)/bias normalisation and ALL n_iter
fixed-line-of-sight iterations into one recurrence per mode -- here the product's own
functor is held to the reference's sequence setup_overdensity! -> n_iter x iterate! without a GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import baorec_oracle as O
from util import clustered_box, rel_rms

ROOT = Path(__file__).resolve().parent.parent
f32 = np.float32
_F, _D = C.POINTER(C.c_float), C.POINTER(C.c_double)


def fp(a):
    return a.ctypes.data_as(_F)


def dp(a):
    return a.ctypes.data_as(_D)


@pytest.fixture(scope="module")
def HC():
    out = ROOT / "tests" / "_build" / "libkspace_hostcheck.so"
    src = ROOT / "tests" / "hostcheck" / "kspace_hostcheck.cpp"
    hdrs = [ROOT / "baorec.jl_b200" / "csrc" / h for h in ("kspace_ops.cuh", "host_shim.cuh")]
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in [src] + hdrs):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.run([gxx, "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    i, f, d = C.c_int, C.c_float, C.c_double
    mesh = [_F, _F, _F, i, i, i]
    for name, args in (("hc_gauss", [_F, _F] + mesh + [_D, _D, _D, d]), ("hc_setup_box", [_F, _F] + mesh + [_D, _D, _D, f, _D]),
                       ("hc_iter_los", [_F, _F] + mesh + [_F, f]), ("hc_iter_pair", [_F, _F] + mesh + [i, i, f]),
                       ("hc_fused_los", [i, _F, _F, _F] + mesh + [_D, _D, _D, f, _D, _F, f, i, f]),
                       ("hc_disp", [i, _F, _F, _F, _F] + mesh + [f])):
        getattr(lib, name).restype = None
        getattr(lib, name).argtypes = args
    return lib


class K:
    """A mesh shape with its k tables and (for radius R) the per-axis Gaussian tables the plan builds (ctx.cu: gauss_tables)."""

    def __init__(self, shape_xyz, L):
        self.nx, self.ny, self.nz = shape_xyz
        self.xh = self.nx // 2 + 1
        self.M = self.nx * self.ny * self.nz
        self.bs = np.asarray(L, f32) if np.ndim(L) else np.full(3, L, f32)
        self.kv = [np.ascontiguousarray(k, f32) for k in O.k_vec(shape_xyz, self.bs, f32)]
        self.mesh = (fp(self.kv[0]), fp(self.kv[1]), fp(self.kv[2]), self.xh, self.ny, self.nz)

    def gauss(self, R):
        R2 = f32(f32(R) * f32(R))
        return [np.exp(-0.5 * np.float64(R2) * (k * k).astype(f32).astype(np.float64)) for k in self.kv]

    def rfft(self, a):
        return np.ascontiguousarray(O.rfft(a).astype(np.complex64))

    def c2r(self, ak):
        """cuFFT's unnormalised C2R."""
        return (O.irfft(ak.astype(np.complex128), (self.nz, self.ny, self.nx)) * self.M).astype(f32)

    def empty(self):
        return np.empty((self.nz, self.ny, self.xh), np.complex64)


def cv(a):
    return fp(a.view(f32))


SHAPES = [((16, 16, 16), 500.0), ((24, 12, 20), (600.0, 330.0, 410.0))]


@pytest.mark.parametrize("shape,L", SHAPES)
def test_gauss_and_setup_box(HC, shape, L):
    k = K(shape, L)
    rng = np.random.default_rng(1)
    rho = rng.random((k.nz, k.ny, k.nx)).astype(f32) + f32(0.5)
    g = k.gauss(15.0)
    out = k.empty()
    HC.hc_gauss(cv(k.rfft(rho)), cv(out), *k.mesh, dp(g[0]), dp(g[1]), dp(g[2]), 1.0 / k.M)
    assert rel_rms(k.c2r(out), O.smooth(rho.copy(), f32(15.0), k.bs)) < 2e-6
    # smooth + (rho/mean - 1)/bias, with mean(rho) read from the DC mode (stash_dc: dc[8] = 1 / (A0 bias))
    rk = k.rfft(rho)
    dc = np.zeros(16)
    dc[0], dc[8] = float(rk[0, 0, 0].real), (1.0 / 2.2) / float(rk[0, 0, 0].real)
    HC.hc_setup_box(cv(rk), cv(out), *k.mesh, dp(g[0]), dp(g[1]), dp(g[2]), 2.2, dp(dc))
    ref = O.smooth(rho.copy(), f32(15.0), k.bs)
    ref = ((ref / f32(ref.mean(dtype=np.float64)) - f32(1)) / f32(2.2)).astype(f32)
    got = k.c2r(out)
    assert rel_rms(got, ref) < 1e-5 and abs(float(got.mean(dtype=np.float64))) < 1e-7


@pytest.mark.parametrize("shape,L", SHAPES)
@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (0.6, 0.0, 0.8)])
@pytest.mark.parametrize("it", [1, 2])
def test_one_iterate_step_fixed_los(HC, shape, L, los, it):
    k = K(shape, L)
    rng = np.random.default_rng(2)
    ds = (0.3 * rng.standard_normal((k.nz, k.ny, k.nx))).astype(f32)
    dr = (ds * f32(0.8)).astype(f32)
    beta = f32(0.344)
    ref = O.iterate(dr.copy(), ds, k.kv, it, beta, los, None)
    out = k.empty()
    HC.hc_iter_los(cv(k.rfft(dr)), cv(out), *k.mesh, fp(np.asarray(los, f32)), f32(1.0 / k.M))
    fac = f32(beta / (f32(1) + beta)) if it == 1 else beta
    got = (ds - fac * k.c2r(out)).astype(f32)              # src/iterative.jl:56-59
    assert rel_rms(got, ref) < 2e-6


@pytest.mark.parametrize("i,j", [(0, 0), (0, 1), (1, 2), (2, 2)])
def test_radial_pair_term(HC, i, j):
    k = K((24, 12, 20), (600.0, 330.0, 410.0))
    rng = np.random.default_rng(3)
    dr = rng.standard_normal((k.nz, k.ny, k.nx)).astype(f32)
    out = k.empty()
    HC.hc_iter_pair(cv(k.rfft(dr)), cv(out), *k.mesh, i, j, f32(1.0 / k.M))
    kb = (k.kv[0][None, None, :], k.kv[1][None, :, None], k.kv[2][:, None, None])
    k2 = O._k2(k.kv, f32)
    dk = O.rfft(dr) / np.where(k2 == 0, 1, k2)
    dk[0, 0, 0] = 0
    ref = O.irfft((kb[i] * kb[j]) * dk, dr.shape).astype(f32)          # src/iterative.jl:27-30
    assert rel_rms(k.c2r(out), ref) < 2e-6


@pytest.mark.parametrize("shape,L", SHAPES)
@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (0.6, 0.0, 0.8)])
@pytest.mark.parametrize("n_iter", [1, 3, 5])
def test_fused_recurrence_equals_the_sequence_of_iterations(HC, shape, L, los, n_iter):
    """MODE 0 (input rho_k): scatter -> [R2C] -> FusedLosOp -> [C2R]  ==  setup_overdensity! + n_iter x iterate!"""
    k = K(shape, L)
    N = 20000
    pos, w = clustered_box(N, 1.0, seed=4)
    pos = [(p * k.bs[a] * f32(0.999)).astype(f32) for a, p in enumerate(pos)]
    rec = O.IterativeRecon(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=k.bs, box_min=np.zeros(3, f32), los=los, n_iter=n_iter)
    rho = O.cic_scatter(np.zeros((k.nz, k.ny, k.nx), f32), *[p.copy() for p in pos], w, k.bs, np.zeros(3, f32), True)
    ref = O.reconstructed_overdensity(np.zeros_like(rho), rec, *[p.copy() for p in pos], w)
    rk = k.rfft(rho)
    g = k.gauss(15.0)
    dc = np.zeros(16)
    dc[0], dc[8] = float(rk[0, 0, 0].real), (k.M / 2.2) / float(rk[0, 0, 0].real)      # stash_dc(mul = M / bias)
    out, keep = k.empty(), k.empty()
    HC.hc_fused_los(0, cv(rk), cv(out), cv(keep), *k.mesh, dp(g[0]), dp(g[1]), dp(g[2]), 2.2, dp(dc), fp(np.asarray(los, f32)),
                    f32(rec.beta), n_iter, f32(1.0 / k.M))
    got = k.c2r(out)
    assert rel_rms(got, ref) < 1e-5
    # the kept delta_k is the unnormalised transform of that mesh: read_shifts starts from it instead of an R2C
    assert rel_rms(O.irfft(keep.astype(np.complex128), rho.shape), ref) < 1e-5
    assert np.allclose(keep * f32(1.0 / k.M), out, rtol=1e-6, atol=0)


@pytest.mark.parametrize("los", [(0.0, 0.0, 1.0), (0.6, 0.0, 0.8)])
def test_fused_recurrence_from_delta_s(HC, los):
    """MODE 1 (input the R2C of delta_s): the k = 0 mode keeps delta_s's (the reference zeroes it in the Hessian term only)."""
    k = K((16, 16, 16), 500.0)
    rng = np.random.default_rng(5)
    ds = (0.3 * rng.standard_normal((k.nz, k.ny, k.nx))).astype(f32) + f32(0.05)
    beta, n_iter = f32(0.344), 3
    ref = ds.copy()
    for it in range(1, n_iter + 1):
        O.iterate(ref, ds, k.kv, it, beta, los, None)
    out = k.empty()
    g = k.gauss(15.0)
    HC.hc_fused_los(1, cv(k.rfft(ds)), cv(out), None, *k.mesh, dp(g[0]), dp(g[1]), dp(g[2]), 2.2, dp(np.zeros(16)),
                    fp(np.asarray(los, f32)), beta, n_iter, f32(1.0 / k.M))
    got = k.c2r(out)
    assert rel_rms(got, ref) < 1e-5 and abs(float(got.mean(dtype=np.float64)) - float(ds.mean(dtype=np.float64))) < 1e-6


@pytest.mark.parametrize("potential", [0, 1])
def test_displacement_components(HC, potential):
    k = K((24, 12, 20), (600.0, 330.0, 410.0))
    rng = np.random.default_rng(6)
    mesh = rng.standard_normal((k.nz, k.ny, k.nx)).astype(f32)
    cls = O.MultigridRecon if potential else O.IterativeRecon
    rec = cls(bias=2.2, f=0.757, smoothing_radius=15.0, box_size=k.bs, box_min=np.zeros(3, f32), los=(0.0, 0.0, 1.0))
    ref = O.displacement_meshes(mesh, rec)                 # src/iterative.jl:260-271 / src/multigrid.jl:761-769
    outs = [k.empty() for _ in range(3)]
    HC.hc_disp(potential, cv(k.rfft(mesh)), cv(outs[0]), cv(outs[1]), cv(outs[2]), *k.mesh, f32(1.0 / k.M))
    for a in range(3):
        assert rel_rms(k.c2r(outs[a]), ref[a]) < 2e-6
