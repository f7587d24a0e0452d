"""Synthetic catalogs for the benchmarks: uniform and lognormal periodic boxes.
Plumbing only (torch tensor ops on whatever device is asked for); nothing here is on the measured path.

lognormal_box: Gaussian field with a smooth P(k) on a generation mesh -> rho = exp(delta_G - sigma^2/2) ->
Poisson counts per cell (trimmed / topped up to exactly N) -> uniform jitter inside the cell -> optional
linear redshift-space shift f * Psi_z -> random permutation (a catalog sorted by cell would flatter the
sorting kernels).  Deterministic for a given (seed, device type)."""
from __future__ import annotations

import math

import numpy as np
import torch


def _below(L):
    """Largest float32 below L (positions are kept strictly inside [0, L))."""
    return float(np.nextafter(np.float32(L), np.float32(0)))


def uniform_box(N, L, seed=42, device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    top = _below(L)
    pos = [(torch.rand(N, generator=g, device=device, dtype=torch.float32) * L).clamp_(max=top) for _ in range(3)]
    return pos, torch.ones(N, dtype=torch.float32, device=device)


def gaussian_field(n, L, seed, device, sigma_target=1.0, k0=0.05, slope=-1.5):
    """delta_G on an n^3 mesh with P(k) ~ k / (1 + (k/k0)^2)^((1 - slope)/2) (turn-over at k0, ~k^slope beyond),
    rescaled to a mesh-scale standard deviation sigma_target.  Also returns delta_k for the displacement."""
    g = torch.Generator(device=device).manual_seed(seed)
    white = torch.randn((n, n, n), generator=g, device=device, dtype=torch.float32)
    wk = torch.fft.rfftn(white)
    kf = 2 * math.pi / L
    kx = torch.fft.rfftfreq(n, 1.0 / n).to(device) * kf
    ky = torch.fft.fftfreq(n, 1.0 / n).to(device) * kf
    k2 = kx[None, None, :] ** 2 + ky[None, :, None] ** 2 + ky[:, None, None] ** 2
    k = k2.sqrt()
    pk = k / (1 + (k / k0) ** 2) ** ((1 - slope) / 2)
    pk[0, 0, 0] = 0
    dk = wk * pk.sqrt()
    delta = torch.fft.irfftn(dk, s=(n, n, n))
    scale = sigma_target / float(delta.std())
    return delta * scale, dk * scale, (kx, ky, k2)


def lognormal_box(N, L, seed=42, device="cpu", n_gen=256, sigma=1.0, f_rsd=0.0):
    """N particles in [0, L)^3 following a lognormal density; returns ([x, y, z], w) float32 on `device`."""
    delta, dk, (kx, ky, k2) = gaussian_field(n_gen, L, seed, device, sigma)
    var = float(delta.var())
    rho = torch.exp(delta - 0.5 * var).double().flatten()
    lam = rho * (N / float(rho.sum()))
    g = torch.Generator(device=device).manual_seed(seed + 1)
    counts = torch.poisson(lam.float(), generator=g).long()
    total = int(counts.sum())
    cells = torch.repeat_interleave(torch.arange(counts.numel(), device=device), counts)
    perm = torch.randperm(total, generator=g, device=device)
    cells = cells[perm]                                        # random order, then fix the count
    if total >= N:
        cells = cells[:N]
    else:                                                      # top up from the same density: the cells of randomly chosen
        pick = torch.randint(0, total, (N - total,), generator=g, device=device)   # particles already drawn (torch.multinomial
        cells = torch.cat([cells, cells[pick]])[torch.randperm(N, generator=g, device=device)]  # stops at 2^24 categories)
    iz = cells // (n_gen * n_gen)
    iy = (cells // n_gen) % n_gen
    ix = cells % n_gen
    cell = L / n_gen
    top = _below(L)
    pos = []
    for idx in (ix, iy, iz):
        jitter = torch.rand(N, generator=g, device=device, dtype=torch.float32)
        pos.append(((idx.float() + jitter) * cell).clamp_(0.0, top))
    if f_rsd:
        k2s = torch.where(k2 > 0, k2, torch.ones_like(k2))
        kz = ky[:, None, None]
        psi_z = torch.fft.irfftn(1j * kz * dk / k2s, s=(n_gen,) * 3)        # -div Psi = delta  (src/recon.jl:377 convention)
        shift = f_rsd * psi_z.flatten()[cells].float()
        pos[2] = torch.remainder(pos[2] + shift, L).clamp_(0.0, top)
    return pos, torch.ones(N, dtype=torch.float32, device=device)
