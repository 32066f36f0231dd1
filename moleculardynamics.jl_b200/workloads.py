"""Synthetic, overlap-free input configurations (C1..C5).  Host-side setup only (numpy); the reference
does this step with Packmol (src/initialization.jl:20-30), which is outside the hot path."""
import math

import numpy as np

BASE_SEED = 20261018
PHI_README = 0.47
KT_README = 1.4737


def rho_from_phi(phi, dim=3):
    return 6.0 * phi / math.pi if dim == 3 else 4.0 * phi / math.pi


def _lattice(kind, m, dim):
    g = np.stack(np.meshgrid(*[np.arange(m, dtype=np.float64)] * dim, indexing="ij"), axis=-1).reshape(-1, dim)
    if kind == "sc":
        basis = np.zeros((1, dim))
    elif kind == "bcc":
        basis = np.array([[0.0] * dim, [0.5] * dim])
    elif kind == "fcc":
        basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]], dtype=np.float64)
    else:
        raise ValueError(kind)
    return (g[:, None, :] + basis[None, :, :] + 0.25).reshape(-1, dim) / m  # fractional coordinates


def phs_fluid(n, phi=PHI_README, dim=3, seed=BASE_SEED, jitter=None):
    """Monodisperse sigma = 1 start at packing fraction phi: lattice (+ random vacancies) + uniform jitter.
    Returns dict(x [n][dim], diam [n], box [dim], rho)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rho = rho_from_phi(phi, dim)
    L = (n / rho) ** (1.0 / dim)
    best = None
    for kind, per in (("sc", 1), ("bcc", 2), ("fcc", 4)):
        if dim == 2 and kind != "sc":
            continue
        m = int(math.ceil((n / per) ** (1.0 / dim) - 1e-9))
        sites = per * m ** dim
        a = L / m
        nn = {"sc": a, "bcc": a * math.sqrt(3) / 2, "fcc": a / math.sqrt(2)}[kind]
        score = (sites == n, nn)
        if best is None or score > best[0]:
            best = (score, kind, m, nn)
    _, kind, m, nn = best
    frac = _lattice(kind, m, dim)
    if frac.shape[0] > n:
        keep = np.sort(rng.permutation(frac.shape[0])[:n])
        frac = frac[keep]
    if jitter is None:
        jitter = max(0.0, min(0.02, 0.45 * (nn - 1.0) / math.sqrt(dim)))
    x = frac * L + rng.uniform(-jitter, jitter, size=frac.shape)
    x -= L * np.floor(x / L)
    return dict(x=np.ascontiguousarray(x), diam=np.ones(n), box=np.full(dim, L), rho=rho, lattice=kind, nn=nn)


def poly2d(n=1200, rho=1.0, seed=BASE_SEED, smin=0.73, smax=1.62):
    """C2: 2-D polydisperse mixture, P(sigma) ~ sigma^-3 on [smin, smax] by inverse CDF; square lattice minus vacancies."""
    rng = np.random.Generator(np.random.PCG64(seed + 2))
    L = math.sqrt(n / rho)
    m = int(math.ceil(math.sqrt(n)))
    frac = _lattice("sc", m, 2)
    keep = np.sort(rng.permutation(frac.shape[0])[:n])
    x = frac[keep] * L + rng.uniform(-0.01, 0.01, size=(n, 2))
    u = rng.uniform(size=n)
    a, b = smin ** -2, smax ** -2
    diam = (a - u * (a - b)) ** -0.5
    x -= L * np.floor(x / L)
    return dict(x=np.ascontiguousarray(x), diam=diam, box=np.full(2, L), rho=rho)


def lj_fluid(n, rho=0.8, dim=3, seed=BASE_SEED):
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    L = (n / rho) ** (1.0 / dim)
    m = int(math.ceil(n ** (1.0 / dim) - 1e-9))
    frac = _lattice("sc", m, dim)
    keep = np.sort(rng.permutation(frac.shape[0])[:n])
    x = frac[keep] * L + rng.uniform(-0.05, 0.05, size=(n, dim))
    x -= L * np.floor(x / L)
    return dict(x=np.ascontiguousarray(x), diam=np.ones(n), box=np.full(dim, L), rho=rho)


def velocities(n, dim, ktemp, seed=BASE_SEED):
    """initialize_velocities (src/initialization.jl:32-47) with a NumPy PCG64 stream"""
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    V = rng.standard_normal((dim, n))
    V -= V.mean(axis=1, keepdims=True)
    fs = math.sqrt(ktemp / (float(np.sum(V * V)) / ((n - 1) * dim)))
    return np.ascontiguousarray((V * fs).T)
