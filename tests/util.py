"""Shared helpers for the parity tests: seeded synthetic catalogs."""
import numpy as np


def uniform_box(n, L, seed=42, lo=0.0):
    rng = np.random.default_rng(seed)
    pos = [(lo + L * rng.random(n)).astype(np.float32) for _ in range(3)]
    for p in pos:  # keep strictly inside [lo, lo+L)
        np.clip(p, np.float32(lo), np.nextafter(np.float32(lo + L), np.float32(lo)), out=p)
    w = np.ones(n, np.float32)
    return pos, w


def clustered_box(n, L, seed=42, lo=0.0, nclump=200, sigma=0.02):
    """Clumpy periodic catalog: Gaussian blobs around random centres (wrapped), plus 30% uniform."""
    rng = np.random.default_rng(seed)
    nu = int(0.3 * n)
    nc = n - nu
    centres = rng.random((nclump, 3))
    which = rng.integers(0, nclump, nc)
    p = centres[which] + sigma * rng.standard_normal((nc, 3))
    p = np.concatenate([p, rng.random((nu, 3))]) % 1.0
    pos = [(lo + L * p[:, a]).astype(np.float32) for a in range(3)]
    for q in pos:
        np.clip(q, np.float32(lo), np.nextafter(np.float32(lo + L), np.float32(lo)), out=q)
    w = (0.5 + rng.random(n)).astype(np.float32)
    return pos, w


def lightcone(n_data, n_rand, seed=42, rmin=1900.0, rmax=2300.0, half_angle_deg=30.0):
    """Shell sector around +x seen from the origin; data are clumpy, randoms uniform in volume."""
    rng = np.random.default_rng(seed)

    def draw(n, clumpy):
        u = rng.random(n)
        r = (rmin ** 3 + u * (rmax ** 3 - rmin ** 3)) ** (1.0 / 3.0)
        ca = np.cos(np.deg2rad(half_angle_deg))
        mu = ca + (1 - ca) * rng.random(n)          # cos(angle from +x axis)
        ph = 2 * np.pi * rng.random(n)
        s = np.sqrt(1 - mu * mu)
        p = np.stack([r * mu, r * s * np.cos(ph), r * s * np.sin(ph)], axis=1)
        if clumpy:
            p += 8.0 * np.sin(p / 55.0)             # smooth, deterministic clustering signal
        return [p[:, a].astype(np.float32) for a in range(3)]

    d = draw(n_data, True)
    r = draw(n_rand, False)
    wd = (1.0 / (1.0 + 0.2 * rng.random(n_data))).astype(np.float32)   # FKP-like weights
    wr = (1.0 / (1.0 + 0.2 * rng.random(n_rand))).astype(np.float32)
    return d, wd, r, wr


def rel_rms(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-300))


def maxabs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def lognormal_radial(N, L, ng, sigma, f, observer, seed):
    """Periodic lognormal box [0, L)^3 (numpy only) with the linear redshift-space shift f (Psi . rhat) rhat seen
    from `observer` (a point outside the box).  Returns (real-space positions, redshift-space positions), (n, 3) Float64."""
    rng = np.random.default_rng(seed)
    wk = np.fft.rfftn(rng.standard_normal((ng, ng, ng)))
    kf = 2 * np.pi / L
    kx, ky = np.fft.rfftfreq(ng, 1.0 / ng) * kf, np.fft.fftfreq(ng, 1.0 / ng) * kf
    KX, KY, KZ = kx[None, None, :], ky[None, :, None], ky[:, None, None]
    k2 = KX ** 2 + KY ** 2 + KZ ** 2
    k = np.sqrt(k2)
    pk = k / (1 + (k / 0.05) ** 2) ** 1.25
    pk[0, 0, 0] = 0
    dk = wk * np.sqrt(pk)
    axes = (0, 1, 2)
    delta = np.fft.irfftn(dk, s=(ng,) * 3, axes=axes)
    scale = sigma / delta.std()
    delta, dk = delta * scale, dk * scale
    k2s = np.where(k2 > 0, k2, 1.0)
    psi = [np.fft.irfftn(1j * K * dk / k2s, s=(ng,) * 3, axes=axes) for K in (KX, KY, KZ)]      # -div Psi = delta
    rho = np.exp(delta - 0.5 * delta.var())
    cells = np.repeat(np.arange(ng ** 3), rng.poisson(rho * (N / rho.sum())).ravel())
    rng.shuffle(cells)
    idx = (cells % ng, (cells // ng) % ng, cells // (ng * ng))
    pos = np.stack([(i + rng.random(len(cells))) * (L / ng) for i in idx], 1)
    P = np.stack([p.ravel()[cells] for p in psi], 1)
    r = pos - np.asarray(observer, np.float64)[None, :]
    rhat = r / np.linalg.norm(r, axis=1)[:, None]
    return pos, pos + f * (P * rhat).sum(1)[:, None] * rhat
