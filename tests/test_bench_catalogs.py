"""CPU: the synthetic catalog generators of the benchmarks (benchmarks/catalogs.py) at small sizes."""
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "benchmarks"))
import catalogs  # noqa: E402


def counts_in_cells(pos, L, n):
    idx = [np.minimum((p.numpy() / L * n).astype(np.int64), n - 1) for p in pos]
    return np.bincount((idx[2] * n + idx[1]) * n + idx[0], minlength=n ** 3)


def test_uniform_box():
    pos, w = catalogs.uniform_box(200_000, 2500.0, seed=1)
    assert all(p.dtype == torch.float32 and float(p.min()) >= 0 and float(p.max()) < 2500.0 for p in pos)
    c = counts_in_cells(pos, 2500.0, 16)
    assert abs(c.var() / c.mean() - 1) < 0.1 and float(w.sum()) == 200_000          # Poisson


@pytest.mark.parametrize("N", [150_000, 400_000])
def test_lognormal_box_is_clustered_exact_and_deterministic(N):
    L = 1000.0
    pos, w = catalogs.lognormal_box(N, L, seed=3, n_gen=32, sigma=1.0)
    assert all(len(p) == N and p.dtype == torch.float32 and float(p.min()) >= 0 and float(p.max()) < L for p in pos)
    c = counts_in_cells(pos, L, 32)
    assert c.var() / c.mean() > 3                                   # far above shot noise: exp(sigma^2) - 1 = 1.7 of the mean^2
    assert abs(np.corrcoef(np.diff(pos[0].numpy()[:5000]), np.diff(pos[1].numpy()[:5000]))[0, 1]) < 0.1
    # not sorted by cell: neighbours in catalog order are far apart
    assert np.median(np.abs(np.diff(pos[2].numpy()))) > L / 10
    pos2, _ = catalogs.lognormal_box(N, L, seed=3, n_gen=32, sigma=1.0)
    assert all(torch.equal(a, b) for a, b in zip(pos, pos2))
    pos3, _ = catalogs.lognormal_box(N, L, seed=4, n_gen=32, sigma=1.0)
    assert not torch.equal(pos[0], pos3[0])


def test_lognormal_rsd_shift_moves_only_z():
    a, _ = catalogs.lognormal_box(50_000, 800.0, seed=5, n_gen=32, f_rsd=0.0)
    b, _ = catalogs.lognormal_box(50_000, 800.0, seed=5, n_gen=32, f_rsd=0.757)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and not torch.equal(a[2], b[2])
    d = (b[2] - a[2] + 400.0) % 800.0 - 400.0
    assert 0.1 < float(d.abs().mean()) < 30.0 and float(b[2].min()) >= 0 and float(b[2].max()) < 800.0
