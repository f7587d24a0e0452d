"""CPU: the arithmetic of the own column FFT kernels (baorec.jl_b200/csrc/fft_column.cuh, fft_radix.cuh -- what
fft_cols_kernel, fft_z_solve_kernel and fft_z_disp_kernel call) compiled as plain C++ by tests/hostcheck/ and walked
thread by thread on the CPU, against numpy's FFT in Float64: the generated in-register butterflies (8 ... 64 points) and
the whole four-step transform N = 32 x M with its twiddle indices, exchange layout and output order, forward and
inverse, for every length and tile width the kernels are instantiated with."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "baorec.jl_b200" / "csrc"
_F = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def HC():
    out = ROOT / "tests" / "_build" / "libfft_hostcheck.so"
    src = ROOT / "tests" / "hostcheck" / "fft_hostcheck.cpp"
    deps = [src] + [CSRC / h for h in ("fft_column.cuh", "fft_radix.cuh", "host_shim.cuh")]
    if not out.exists() or out.stat().st_mtime < max(p.stat().st_mtime for p in deps):
        out.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        subprocess.run([gxx, "-std=c++17", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-I", str(CSRC),
                        "-o", str(out), str(src)], check=True)
    lib = C.CDLL(str(out))
    lib.hc_fft_reg.argtypes = [C.c_int, C.c_int, _F, _F]
    lib.hc_fft_tile.argtypes = [C.c_int, C.c_int, C.c_int, _F, _F, _F]
    lib.hc_fft_exchange_floats.argtypes = [C.c_int, C.c_int]
    return lib


def ptr(a):
    return a.ctypes.data_as(_F)


def reference(x, direction):
    x = x.astype(np.complex128)
    return np.fft.fft(x, axis=-1) if direction > 0 else np.fft.ifft(x, axis=-1) * x.shape[-1]     # unnormalised, like cuFFT


@pytest.mark.parametrize("R", [8, 16, 32, 64])
@pytest.mark.parametrize("direction", [1, -1])
def test_generated_butterflies(HC, R, direction):
    rng = np.random.default_rng(R)
    for _ in range(20):
        x = (rng.standard_normal(R) + 1j * rng.standard_normal(R)).astype(np.complex64)
        out = np.empty(R, np.complex64)
        assert HC.hc_fft_reg(R, direction, ptr(x.view(np.float32)), ptr(out.view(np.float32))) == 0
        want = reference(x, direction)
        assert np.abs(out - want).max() < 4e-7 * np.log2(R) * np.abs(want).max()
    # a delta at position p: the pure twiddle column, every output of modulus 1
    for p in (0, 1, R // 2 + 1):
        x = np.zeros(R, np.complex64)
        x[p] = 1
        out = np.empty(R, np.complex64)
        HC.hc_fft_reg(R, direction, ptr(x.view(np.float32)), ptr(out.view(np.float32)))
        assert np.abs(out - np.exp(-2j * np.pi * direction * p * np.arange(R) / R)).max() < 3e-7


@pytest.mark.parametrize("N,TX", [(256, 8), (512, 8), (1024, 8), (2048, 8), (1024, 4), (1024, 16)])
@pytest.mark.parametrize("direction", [1, -1])
def test_four_step_column_transform(HC, N, TX, direction):
    rng = np.random.default_rng(N + TX)
    k = np.arange(N)
    tw = np.exp(-2j * np.pi * k / N)                          # own_fft_setup: cos / sin in Float64, rounded to Float32
    tw = (tw.real.astype(np.float32) + 1j * tw.imag.astype(np.float32)).astype(np.complex64)
    x = (rng.standard_normal((TX, N)) + 1j * rng.standard_normal((TX, N))).astype(np.complex64)
    x[0] = 0
    x[0, 3] = 1                                               # one column is a delta: exposes any index slip exactly
    x[1] = np.exp(2j * np.pi * 5 * k / N)                     # and one a single mode: a spike at +5 (forward) / -5 (inverse)
    out = np.full((TX, N), np.nan + 0j, np.complex64)         # every output must be written
    assert HC.hc_fft_tile(N, direction, TX, ptr(x.view(np.float32)), ptr(tw.view(np.float32)), ptr(out.view(np.float32))) == 0
    want = reference(x, direction)
    assert np.isfinite(out.view(np.float32)).all()
    err = np.abs(out - want).max(axis=1) / np.abs(want).max(axis=1)
    assert err.max() < 1.5e-6, err
    spike = 5 if direction > 0 else N - 5
    assert abs(out[1, spike] - N) < 1e-3 * N and np.abs(np.delete(out[1], spike)).max() < 2e-3 * np.sqrt(N)
    # forward then inverse returns N x the input (what the fused z pass relies on: the forward output order is the inverse's input order)
    back = np.empty_like(out)
    HC.hc_fft_tile(N, -direction, TX, ptr(out.view(np.float32)), ptr(tw.view(np.float32)), ptr(back.view(np.float32)))
    assert np.abs(back / N - x).max() < 3e-6 * max(1.0, np.abs(x).max())


def test_exchange_buffer_layout_is_conflict_free(HC):
    """Xch<N, TX>: per column M rows of 33 slots, column stride 128 / TX bytes off a multiple of 128 -- a half-warp of
    TX columns x 16 / TX threads then covers the 32 banks once, for the writes (slot t of row k2) and the reads (row t)."""
    for N, TX in [(256, 8), (512, 8), (1024, 8), (2048, 8), (1024, 4), (1024, 16)]:
        M = N // 32
        col = HC.hc_fft_exchange_floats(N, TX) // 2 // TX           # float2 per column
        assert col >= M * 33 and (col * 8) % 128 == 128 // TX
        per = 16 // TX if TX <= 16 else 1
        for k2 in (0, 1, M - 1):
            banks = {((c * col + k2 * 33 + t) * 2 + h) % 32 for c in range(TX) for t in range(per) for h in (0, 1)}
            assert len(banks) == 32, (N, TX, k2)
