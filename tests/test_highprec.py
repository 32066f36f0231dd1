"""An arbiter that owes nothing to libm, the oracle or the GPU: the pair sums of the committed snapshots evaluated with
60-digit arithmetic (mpmath) straight from the reference formulas (src/pairwise.jl:26-39, src/potentials.jl:11-29, 66-77,
README.md:89-145).  Double-precision inputs are exact in mpmath, so these are the mathematically exact values of the
functions the reference computes in Float64; the oracle (CPU) and the CUDA path (-m gpu) must both sit within the
tolerance north_star states (1e-12 relative).  What the distance measures is mostly the conditioning of the reference's
own Float64 formula, not an implementation: f = a (50 s^51 - 49 s^50) amplifies the rounding of s = sigma/r by ~50 and then
cancels two terms ~50x larger than their difference, ~5e-13 in the worst case (measured on C1: forces 9e-14, E 1e-15).
The bounds below sit just above the measured values so that a regression shows up long before 1e-12."""
import os

import numpy as np
import pytest

from conftest import force_error, relerr

mp = pytest.importorskip("mpmath")
mp.mp.dps = 60
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _pairs(x, box, rmax):
    """candidate pairs (i < j) within rmax under the minimum image, found in float64 with a margin"""
    n = x.shape[0]
    out = []
    for i in range(n - 1):
        d = x[i] - x[i + 1:]
        d -= box * np.rint(d / box)
        r2 = np.einsum("ij,ij->i", d, d)
        for j in np.nonzero(r2 < (rmax * (1 + 1e-9)) ** 2)[0]:
            out.append((i, i + 1 + int(j)))
    return out


def _exact(x, diam, box, pot, params, cutoff):
    """(E, W, F[n][dim], n_interacting) in 60-digit arithmetic"""
    n, dim = x.shape
    mpf = mp.mpf
    L = [mpf(float(b)) for b in box]
    F = [[mpf(0)] * dim for _ in range(n)]
    E = W = mpf(0)
    n_int = 0
    if pot == "phs":
        rng_max = 1.0204081632653061
    elif pot == "lj":
        rng_max = params[1]
    else:
        smax, smin = float(diam.max()), float(diam.min())
        rng_max = params[0] * smax * (1 + abs(params[1]) * (smax - smin))
    margin = []
    for i, j in _pairs(x, box, min(cutoff, rng_max)):
        r = []
        for k in range(dim):
            dk = mpf(float(x[i, k])) - mpf(float(x[j, k]))
            dk -= L[k] * mp.nint(dk / L[k])
            r.append(dk)
        d2 = sum(c * c for c in r)
        d = mp.sqrt(d2)
        if d2 > mpf(cutoff) ** 2:
            continue
        s1, s2 = mpf(float(diam[i])), mpf(float(diam[j]))
        if pot == "phs":
            b, a = mpf(1.0204081632653061), mpf(134.5526623421209)      # the reference's Float64 constants (src/potentials.jl:2-3)
            margin.append(abs(d - b))
            if not d < b:
                continue
            s = (s1 + s2) / 2 / d
            u = a * (s ** 50 - s ** 49) + 1
            f = a * (50 * s ** 51 - 49 * s ** 50)
        elif pot == "lj":
            eps, rc = mpf(params[0]), mpf(params[1])
            margin.append(abs(d - rc))
            if d >= rc:
                continue
            sr6 = ((s1 + s2) / 2 / d) ** 6
            u = 4 * eps * (sr6 * sr6 - sr6)
            f = 24 * eps * (2 * sr6 * sr6 - sr6) / d
        else:
            rc, na = mpf(params[0]), mpf(params[1])
            s = (s1 + s2) / 2 * (1 - na * abs(s1 - s2))
            margin.append(abs(d - rc * s))
            if not d < rc * s:
                continue
            c0, c2, c4 = -28 / rc ** 12, 48 / rc ** 14, -21 / rc ** 16
            u = (s / d) ** 12 + c0 + c2 * (d / s) ** 2 + c4 * (d / s) ** 4
            f = 12 * s ** 12 / d ** 13 - 2 * c2 * d / s ** 2 - 4 * c4 * d ** 3 / s ** 4
        n_int += 1
        E += u
        for k in range(dim):
            sk = f * r[k] / d
            W += sk * r[k]
            F[i][k] += sk
            F[j][k] -= sk
    assert min(margin) > 1e-10, "a pair sits on the potential's range boundary: the exact pair set is ambiguous"
    return float(E), float(W), np.array([[float(c) for c in row] for row in F]), n_int


_cache = {}


def _case(name):
    if name in _cache:
        return _cache[name]
    if name == "c1_phs":
        g = np.load(os.path.join(GOLD, "c1_phs_n1024.npz"))
        spec = (g["x"], g["diam"], np.asarray(g["box"], float).ravel()[:3], "phs", (), 1.5)
    elif name == "c1_lj":
        g = np.load(os.path.join(GOLD, "c1_phs_n1024.npz"))
        spec = (g["x"], g["diam"], np.asarray(g["box"], float).ravel()[:3], "lj", (1.0, 1.7), 1.7)
    else:
        g = np.load(os.path.join(GOLD, "c2_poly_n1200_cut2.03.npz"))
        spec = (g["x"], g["diam"], np.asarray(g["box"], float).ravel()[:2], "poly", (1.25, 0.2), 2.03)
    _cache[name] = spec + (_exact(*spec),)
    return _cache[name]


CASES = ["c1_phs", "c1_lj", "c2_poly"]


def _check(E, W, n_int, F, exact, what):
    Ex, Wx, Fx, nx = exact
    assert n_int == nx, what
    ferr = force_error(F, Fx)
    # measured (oracle): c1_phs forces 9.3e-14, E 1.2e-15, W 1.8e-16; c1_lj forces 2.9e-14; c2_poly below 2e-14.  Contract: 1e-12.
    fmax = 4e-13 if "phs" in what else 1e-13
    assert ferr <= fmax and relerr(E, Ex) <= 2e-14 and relerr(W, Wx) <= 2e-14, (what, ferr, relerr(E, Ex), relerr(W, Wx))


@pytest.mark.parametrize("name", CASES)
def test_oracle_against_60_digit_pair_sums(orc, name):
    x, diam, box, pot, params, cutoff, exact = _case(name)
    tag = {"phs": orc.POT_PHS, "lj": orc.POT_LJ, "poly": orc.POT_POLY}[pot]
    o = orc.forces(x, diam, box, cutoff, tag, params)
    _check(o["E"], o["W"], o["n_int"], o["F"], exact, "oracle " + name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["list", "cells"])
def test_gpu_against_60_digit_pair_sums(md, name, mode):
    x, diam, box, pot, params, cutoff, exact = _case(name)
    n, dim = x.shape
    tag = {"phs": md._capi.POT_PSEUDOHS, "lj": md._capi.POT_LJ, "poly": md._capi.POT_POLY}[pot]
    e = md.Engine(dim, n, box, cutoff, tag, params, seed=1, mode={"list": md._capi.MODE_LIST, "cells": md._capi.MODE_CELLS}[mode])
    e.upload(x, diam)
    E, W, npairs = e.compute_forces()
    _check(E, W, npairs, e.download()[2], exact, "gpu %s %s" % (name, mode))
    e.close()
