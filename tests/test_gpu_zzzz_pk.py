"""GPU parity: the device power-spectrum multipole estimator (baorec.jl_b200/csrc/pk.cu, C ABI
baorec_power_multipoles_f32) against oracle/pk_oracle.py.  Mode counts per bin are exact (Float64 bin arithmetic on
the same Float32 k tables); multipoles agree to the Float32 transform's rounding (cuFFT vs pocketfft).
The per-mode arithmetic is validated on the CPU (tests/test_pk_hostcheck.py); the kernel was written after this
round's GPU budget was spent and has NOT YET RUN ON HARDWARE -- hence the file name that sorts last."""
import numpy as np
import pytest

import pk_oracle as PK

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
f32 = np.float32


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape,L,los,mas", [((64, 64, 64), 500.0, (0.0, 0.0, 1.0), "cic"), ((48, 40, 56), (300.0, 250.0, 350.0), (0.3, -0.5, 0.8), "tsc"),
                                             ((32, 32, 33), 200.0, (1.0, 0.0, 0.0), None)])
def test_power_multipoles_match_the_oracle(B, O, shape, L, los, mas):
    nx, ny, nz = shape
    bs, bm = np.broadcast_to(np.asarray(L, f32), 3).copy(), np.zeros(3, f32)
    rng = np.random.default_rng(2)
    N = 300_000
    pos = [(bs[a] * rng.random(N)).astype(f32) for a in range(3)]
    w = (0.5 + rng.random(N)).astype(f32)
    rho = torch.zeros((nz, ny, nx), dtype=torch.float32, device="cuda")
    B.cic(rho, *(dev(p) for p in pos), dev(w), bs, bm, wrap=True, mas=mas or "cic")
    before = rho.clone()
    hrho = rho.cpu().numpy()
    power = {None: 0, "cic": 2, "tsc": 3}[mas]
    kf = 2 * np.pi / float(bs.max())
    shot = float(np.prod(bs.astype(np.float64))) * float((w.astype(np.float64) ** 2).sum()) / float(w.sum(dtype=np.float64)) ** 2
    for kmin, dk, nbins in ((0.0, kf, 20), (0.5 * kf, kf, 64), (0.013, 0.0071, 25)):
        ref = PK.power_multipoles(hrho, bs, los=los, kmin=kmin, dk=dk, nbins=nbins, mas_power=power, shot=shot)
        got = B.power_multipoles(rho, bs, los=los, kmin=kmin, dk=dk, nbins=nbins, mas=mas, shot=shot)
        assert np.array_equal(got["nmodes"], ref["nmodes"])
        ok = ref["nmodes"] > 0
        assert np.allclose(got["k"][ok], ref["k"][ok], rtol=1e-12, atol=0)
        scale = float(np.abs(ref["p0"][ok] + shot).max())
        for key in ("p0", "p2", "p4"):
            assert np.abs(got[key][ok] - ref[key][ok]).max() <= 1e-5 * scale
        assert np.all(np.isnan(got["p0"][~ok]))
    assert torch.equal(rho, before)                                       # the mesh is an input


def test_plane_wave_and_errors(B):
    n, L, A, m = 64, 100.0, 0.1, 5
    x = np.arange(n) * L / n
    rho = dev((np.ones((n, n, n)) * (1 + A * np.cos(2 * np.pi * m * x / L))[None, None, :]).astype(f32))
    kf = 2 * np.pi / L
    r = B.power_multipoles(rho, np.full(3, L, f32), los=(0.0, 0.0, 1.0), kmin=0.5 * kf, dk=kf, nbins=16, mas=None)
    assert np.isclose(r["p0"][m - 1], L ** 3 * A * A / 4 * 2 / r["nmodes"][m - 1], rtol=1e-5)
    assert np.isclose(r["p2"][m - 1] / r["p0"][m - 1], -2.5, rtol=1e-5)
    others = np.delete(np.arange(16), m - 1)
    assert np.abs(r["p0"][others]).max() < 1e-8 * r["p0"][m - 1]
    with pytest.raises(B.BaorecError):
        B.power_multipoles(torch.zeros((n, n, n), dtype=torch.float32, device="cuda"), np.full(3, L, f32))
    with pytest.raises(B.BaorecError):
        B.power_multipoles(rho, np.full(3, L, f32), los=(0.0, 0.0, 0.0))
