"""CPU: the per-mode arithmetic of the device multipole estimator (baorec.jl_b200/csrc/pk_ops.cuh, what pk_kernel
calls), compiled as plain C++ by tests/hostcheck/ and walked over whole half meshes on the CPU, against
oracle/pk_oracle.py: identical mode counts per bin (the bin index is Float64 arithmetic on the same Float32 k
tables on both sides), multipoles to Float64 rounding.  The host-side finish of csrc/pk.cu (window tables,
V / rho_0^2, 2l+1, shot noise) is restated in `finish` below."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import scipy.fft

import baorec_oracle as O
import pk_oracle as PK
from util import clustered_box

ROOT = Path(__file__).resolve().parent.parent
f32 = np.float32
_F, _D = C.POINTER(C.c_float), C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def HC():
    out = ROOT / "tests" / "_build" / "libpk_hostcheck.so"
    src = ROOT / "tests" / "hostcheck" / "pk_hostcheck.cpp"
    hdrs = [ROOT / "baorec.jl_b200" / "csrc" / h for h in ("pk_ops.cuh", "host_shim.cuh")]
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in [src] + hdrs):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.run([gxx, "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    lib.hc_pk.restype = C.c_int64
    lib.hc_pk.argtypes = [_F, _F, C.c_double, C.c_double, _F, _F, _F, C.c_int, C.c_int, C.c_int, _D, _D, _D, _D, C.c_double, C.c_double, C.c_int, _D]
    return lib


def device_estimate(HC, rho, bs, los, kmin, dk, nbins, power, shot, randoms=None):
    """csrc/pk.cu with the kernel's mode loop on the CPU: R2C (complex64, like cuFFT), window tables from the
    context's k tables, pk_mode per mode, then the host-side finish."""
    nz, ny, nx = rho.shape
    # pk_sum_kernel + pk_contrast_kernel: Float64 sums, the contrast formed in real space and rounded once to Float32,
    # so that the complex64 transform never sees the mean (a Float32 transform of the raw density is off by 5e-6 of
    # the largest bin here and by 4e-4 on the 32 x 32 x 33 mesh of the first hardware run)
    M = float(rho.size)
    d = rho.astype(np.float64) * (M / float(rho.sum(dtype=np.float64)))
    d = d - 1.0 if randoms is None else d - randoms.astype(np.float64) * (M / float(randoms.sum(dtype=np.float64)))
    rk = np.ascontiguousarray(scipy.fft.rfftn(d.astype(f32)).astype(np.complex64))
    sk, sa, sb = None, 1.0 / M, 0.0
    kv = [np.ascontiguousarray(k, f32) for k in O.k_vec((nx, ny, nz), bs, f32)]
    h = np.asarray(bs, f32).astype(np.float64) / np.array([nx, ny, nz], np.float64)

    def window(k, h1):
        x = k.astype(np.float64) * h1 / 2.0
        s = np.where(x == 0, 1.0, np.sin(x) / np.where(x == 0, 1.0, x))
        r = np.ones_like(s)
        for _ in range(power):
            r = r * s
        return np.ascontiguousarray(1.0 / (r * r))               # csrc/pk.cu: inv_window2

    wt = [window(kv[a], h[a]) for a in range(3)]
    lv = np.asarray(los, f32).astype(np.float64)
    lv = np.ascontiguousarray(lv / np.sqrt((lv * lv).sum()))
    acc = np.zeros(5 * nbins, np.float64)
    fp, dp = (lambda a: a.ctypes.data_as(_F)), (lambda a: a.ctypes.data_as(_D))
    HC.hc_pk(fp(rk.view(f32)), None if sk is None else fp(sk.view(f32)), sa, sb, fp(kv[0]), fp(kv[1]), fp(kv[2]), nx, ny, nz, dp(wt[0]), dp(wt[1]), dp(wt[2]), dp(lv),
             float(kmin), float(dk), nbins, dp(acc))
    acc = acc.reshape(5, nbins)
    norm = float(np.prod(np.asarray(bs, f32).astype(np.float64)))
    with np.errstate(invalid="ignore", divide="ignore"):
        cnt = acc[0]
        return dict(nmodes=cnt, k=acc[1] / cnt, p0=acc[2] / cnt * norm - shot, p2=5 * acc[3] / cnt * norm, p4=9 * acc[4] / cnt * norm)


@pytest.mark.parametrize("shape,L,los,power", [((32, 32, 32), 500.0, (0.0, 0.0, 1.0), 2), ((24, 20, 28), (300.0, 250.0, 350.0), (0.3, -0.5, 0.8), 3),
                                               ((16, 16, 17), 200.0, (1.0, 0.0, 0.0), 0)])
def test_device_mode_arithmetic_matches_the_oracle(HC, shape, L, los, power):
    nx, ny, nz = shape
    bs = np.broadcast_to(np.asarray(L, f32), 3).copy()
    rng = np.random.default_rng(2)
    pos = [(bs[a] * rng.random(60_000)).astype(f32) for a in range(3)]
    w = (0.5 + rng.random(60_000)).astype(f32)
    scatter = O.cic_scatter if power != 3 else O.tsc_scatter
    rho = scatter(np.zeros((nz, ny, nx), f32), *pos, w, bs, np.zeros(3, f32), True)
    kf = 2 * np.pi / float(bs.max())
    for kmin, dk, nbins in ((0.0, kf, 12), (0.5 * kf, kf, 40), (0.013, 0.0071, 25)):
        shot = float(np.prod(bs.astype(np.float64))) * float((w.astype(np.float64) ** 2).sum()) / float(w.sum(dtype=np.float64)) ** 2
        ref = PK.power_multipoles(rho, bs, los=los, kmin=kmin, dk=dk, nbins=nbins, mas_power=power, shot=shot)
        got = device_estimate(HC, rho, bs, los, kmin, dk, nbins, power, shot)
        assert np.array_equal(got["nmodes"], ref["nmodes"])              # same bin for every mode, edges included
        ok = ref["nmodes"] > 0
        assert np.allclose(got["k"][ok], ref["k"][ok], rtol=1e-13, atol=0)
        scale = np.abs(ref["p0"][ok] + shot)
        for key in ("p0", "p2", "p4"):
            assert np.abs(got[key][ok] - ref[key][ok]).max() <= 2e-6 * scale.max()     # complex64 vs the oracle's transform
        assert np.all(np.isnan(got["p0"][~ok]))


def test_plane_wave_through_the_device_arithmetic(HC):
    n, L, A, m = 32, 100.0, 0.1, 3
    x = np.arange(n) * L / n
    rho = (np.ones((n, n, n)) * (1 + A * np.cos(2 * np.pi * m * x / L))[None, None, :]).astype(f32)
    kf = 2 * np.pi / L
    r = device_estimate(HC, rho, np.full(3, L, f32), (0.0, 0.0, 1.0), 0.5 * kf, kf, 8, 0, 0.0)
    assert np.isclose(r["p0"][m - 1], L ** 3 * A * A / 4 * 2 / r["nmodes"][m - 1], rtol=1e-5)
    assert np.isclose(r["p2"][m - 1] / r["p0"][m - 1], -2.5, rtol=1e-6)


def test_device_arithmetic_against_the_golden_fixture(HC):
    """No oracle run: the committed fixture tests/golden/pk_40.npz (anisotropic 40 x 36 x 44 mesh, oblique line of sight)."""
    with np.load(ROOT / "tests" / "golden" / "pk_40.npz") as zf:
        g = {k: zf[k] for k in zf.files}
    for power, tag in ((2, "cic"), (0, "raw")):
        got = device_estimate(HC, g["rho"], g["box_size"], g["los"], float(g["kmin"]), float(g["dk"]), int(g["nbins"]), power, float(g["shot"]))
        assert np.array_equal(got["nmodes"], g[f"{tag}_nmodes"])
        ok = got["nmodes"] > 0
        scale = np.abs(g[f"{tag}_p0"][ok] + float(g["shot"])).max()
        for key in ("p0", "p2", "p4"):
            assert np.abs(got[key][ok] - g[f"{tag}_{key}"][ok]).max() < 2e-6 * scale


def test_data_minus_shifted_randoms(HC):
    """The second mesh of the estimator (compute_auto_box_rand of the reference's helpers)."""
    n, L = 32, 400.0
    bs = np.full(3, L, f32)
    pos, w = clustered_box(80_000, L, seed=4)
    rng = np.random.default_rng(5)
    ran = [(L * rng.random(200_000)).astype(f32) for _ in range(3)]
    rho = O.cic_scatter(np.zeros((n, n, n), f32), *pos, w, bs, np.zeros(3, f32), True)
    rmesh = O.cic_scatter(np.zeros((n, n, n), f32), *ran, np.ones(200_000, f32), bs, np.zeros(3, f32), True)
    ref = PK.power_multipoles(rho, bs, los=(0.0, 0.6, 0.8), kmin=0.0, dk=0.02, nbins=12, mas_power=2, shot=3.0, randoms=rmesh)
    got = device_estimate(HC, rho, bs, (0.0, 0.6, 0.8), 0.0, 0.02, 12, 2, 3.0, randoms=rmesh)
    assert np.array_equal(got["nmodes"], ref["nmodes"])
    ok = ref["nmodes"] > 0
    for key in ("p0", "p2", "p4"):
        assert np.abs(got[key][ok] - ref[key][ok]).max() <= 2e-6 * np.abs(ref["p0"][ok]).max()
    plain = PK.power_multipoles(rho, bs, los=(0.0, 0.6, 0.8), kmin=0.0, dk=0.02, nbins=12, mas_power=2, shot=3.0)
    assert np.abs(ref["p0"][ok] - plain["p0"][ok]).max() > 1e-3 * np.abs(plain["p0"][ok]).max()      # the randoms do enter


def test_segmented_warp_reduction_logic():
    """pk_kernel's accumulation, restated lane by lane: run id = popcount of the head-flag ballot up to the lane, five
    shuffle-down steps that add only within a run, run heads hold the run totals (csrc/pk.cu).  Checks the algorithm
    (the compiled kernel is checked on the GPU against the oracle): arbitrary runs, including bins that reappear."""
    rng = np.random.default_rng(0)
    for _ in range(500):
        bins = []
        while len(bins) < 32:
            bins += [int(rng.integers(-1, 6))] * int(rng.integers(1, 12))
        bins, val = np.array(bins[:32]), rng.random(32)
        head = np.array([1] + [int(bins[i] != bins[i - 1]) for i in range(1, 32)])
        ballot = sum(int(h) << i for i, h in enumerate(head))
        run = np.array([bin(ballot & (0xFFFFFFFF >> (31 - lane))).count("1") for lane in range(32)])
        s = val.copy()
        d = 1
        while d < 32:
            t = np.array([s[i + d] if i + d < 32 else s[i] for i in range(32)])         # __shfl_down: out-of-range lanes read themselves
            r = np.array([run[i + d] if i + d < 32 else run[i] for i in range(32)])
            s = np.where((np.arange(32) + d < 32) & (r == run), s + t, s)
            d *= 2
        acc = {}
        for lane in range(32):
            if head[lane] and bins[lane] >= 0:
                acc[bins[lane]] = acc.get(bins[lane], 0.0) + s[lane]                    # the atomics of the run heads
        for b in set(bins[bins >= 0]):
            assert abs(acc[b] - val[bins == b].sum()) < 1e-12
