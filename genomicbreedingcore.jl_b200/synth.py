"""Synthetic allele-frequency matrices with the value distribution of the reference's simulator.

`simulategenomes` (src/simulations/simulate_genomes.jl:622-752) draws per-locus means mu ~ Beta(0.5, 0.5) (:372-373),
a per-locus spread sigma = min(mu, 1-mu) (:376-379), one normal draw per entry (:489,499) clipped to [0,1]
(:515-516); with `ld_corr_50perc_kb = 0` it is plain `rand(n, p)` (:687-689).  LD between loci is ignored here (it
does not change the arithmetic of the GRM path).  Julia's RNG stream cannot be replayed without Julia, so these are
distribution-equivalent, not bit-identical, inputs.

The reference never emits dosage-quantised data; `dosage` is the generator for the int8 path:
c ~ Binomial(ploidy, mu_k), x = c / ploidy.

All host generators return Fortran-ordered (n, p) float64 arrays (Julia's layout).
"""
from __future__ import annotations

import numpy as np


def _mu(rng: np.random.Generator, p: int) -> np.ndarray:
    return rng.beta(0.5, 0.5, size=p)


def continuous(n: int, p: int, seed: int = 42) -> np.ndarray:
    rng = np.random.default_rng(seed)
    mu = _mu(rng, p)
    sigma = np.minimum(mu, 1.0 - mu)
    out = np.empty((n, p), dtype=np.float64, order="F")
    step = max(1, (1 << 22) // max(n, 1))
    for k0 in range(0, p, step):
        k1 = min(p, k0 + step)
        z = rng.standard_normal((k1 - k0, n)).T  # column blocks, F-friendly
        out[:, k0:k1] = np.clip(mu[k0:k1] + sigma[k0:k1] * z, 0.0, 1.0)
    return out


def uniform(n: int, p: int, seed: int = 42) -> np.ndarray:
    """simulategenomes(ld_corr_50perc_kb=0): rand(n, p)."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(rng.random((p, n)).T)


def dosage(n: int, p: int, ploidy: int = 2, seed: int = 42) -> np.ndarray:
    rng = np.random.default_rng(seed)
    mu = _mu(rng, p)
    out = np.empty((n, p), dtype=np.float64, order="F")
    step = max(1, (1 << 22) // max(n, 1))
    for k0 in range(0, p, step):
        k1 = min(p, k0 + step)
        c = rng.binomial(ploidy, mu[k0:k1], size=(n, k1 - k0))
        out[:, k0:k1] = c / float(ploidy)
    return out


def add_missing(A: np.ndarray, sparsity: float, seed: int = 42) -> np.ndarray:
    """round(sparsity*n*p) cells chosen without replacement become NaN (simulate_genomes.jl:737-742)."""
    rng = np.random.default_rng(seed)
    n, p = A.shape
    k = int(round(sparsity * n * p))
    out = np.array(A, order="F", copy=True)
    if k > 0:
        idx = rng.choice(n * p, size=k, replace=False)
        out[idx % n, idx // n] = np.nan
    return out


# ---- device-side generators (bench: never materialise tens of GB on the host) ----------------------------------
def dosage_device(n: int, p: int, ploidy: int, seed: int, device, chunk: int = 8192):
    """(p, n) contiguous CUDA tensor == n x p column-major.  mu_k = sin^2(pi U / 2) is the Beta(1/2,1/2) law."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((p, n), dtype=torch.float64, device=device)
    for k0 in range(0, p, chunk):
        k1 = min(p, k0 + chunk)
        mu = torch.sin(torch.rand((k1 - k0, 1), generator=g, device=device, dtype=torch.float32) * (np.pi / 2)) ** 2
        c = torch.zeros((k1 - k0, n), dtype=torch.float32, device=device)
        for _ in range(ploidy):
            c += (torch.rand((k1 - k0, n), generator=g, device=device, dtype=torch.float32) < mu).float()
        out[k0:k1] = (c / ploidy).double()
    return out


def continuous_device(n: int, p: int, seed: int, device, chunk: int = 8192):
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((p, n), dtype=torch.float64, device=device)
    for k0 in range(0, p, chunk):
        k1 = min(p, k0 + chunk)
        mu = torch.sin(torch.rand((k1 - k0, 1), generator=g, device=device, dtype=torch.float64) * (np.pi / 2)) ** 2
        sigma = torch.minimum(mu, 1.0 - mu)
        z = torch.randn((k1 - k0, n), generator=g, device=device, dtype=torch.float64)
        out[k0:k1] = torch.clamp(mu + sigma * z, 0.0, 1.0)
    return out
