"""ctypes binding of the CPU oracle (oracle/md_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of md_oracle.c.  Imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs; never by the
product package.  Arrays are AoS [n][dim] float64 (C order) / int32 images.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmdoracle.so")

POT_PHS, POT_LJ, POT_XPLOR, POT_POLY, POT_SOFT = 0, 1, 2, 3, 4
NVE, NVT, BROWNIAN = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "md_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmdoracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_evaluate.restype = C.c_int
        _lib.orc_evaluate.argtypes = [C.c_int, _dp, C.c_double, C.c_double, C.c_double, _dp, _dp]
        for name in ("orc_forces", "orc_forces_brute"):
            fn = getattr(_lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _lp, _lp, _ip]
        _lib.orc_wrap.argtypes = [C.c_int, _dp, _ip, _dp]
        _lib.orc_integrate_half.argtypes = [C.c_int, C.c_int64, _dp, _ip, _dp, _dp, C.c_double, _dp]
        _lib.orc_integrate_second_half.argtypes = [C.c_int, C.c_int64, _dp, _dp, C.c_double]
        _lib.orc_kinetic.restype = C.c_double
        _lib.orc_kinetic.argtypes = [C.c_int, C.c_int64, _dp]
        _lib.orc_bussi_scale.restype = C.c_double
        _lib.orc_bussi_scale.argtypes = [C.c_double] * 7
        _lib.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        _lib.orc_bussi_noises.argtypes = [C.c_uint64, C.c_uint64, C.c_double, _dp, _dp]
        _lib.orc_thermo_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, _dp]
        _lib.orc_chi2.restype = C.c_double
        _lib.orc_chi2.argtypes = [C.c_uint64, C.c_uint64, C.c_double]
        _lib.orc_brownian_noise.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, _dp]
        _lib.orc_integrate_brownian.argtypes = [C.c_int, C.c_int64, _dp, _ip, _dp, C.c_double, C.c_double, _dp,
                                                C.c_uint64, C.c_uint64, _ip]
        _lib.orc_run.restype = C.c_int
        _lib.orc_run.argtypes = [C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _ip, _dp, _dp, C.c_double, C.c_int, _dp,
                                 C.c_double, C.c_int64, _dp, C.c_double, C.c_double, C.c_uint64, C.c_uint64, _dp]
        _lib.orc_fire.restype = C.c_int64
        _lib.orc_fire.argtypes = [C.c_int, C.c_int64, _dp, _ip, _dp, _dp, C.c_double, C.c_int, _dp, C.c_int64] + [C.c_double] * 6 + \
            [C.c_int, _dp, C.POINTER(C.c_int), _dp]
        _lib.orc_init_velocities.argtypes = [C.c_int, C.c_int64, C.c_double, C.c_uint64, C.c_uint64, _dp]
        _lib.orc_random_positions.argtypes = [C.c_int, C.c_int64, _dp, C.c_uint64, C.c_uint64, _dp]
        _lib.orc_cell_inverse.restype = C.c_double
        _lib.orc_cell_inverse.argtypes = [_dp, _dp]
        _lib.orc_wrap_tri.argtypes = [C.c_int, _dp, _ip, _dp, _dp]
        _lib.orc_forces_tri.restype = C.c_int
        _lib.orc_forces_tri.argtypes = [C.c_int, C.c_int64, _dp, _dp, _dp, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _lp, _lp]
        _lib.orc_run_tri.restype = C.c_int
        _lib.orc_run_tri.argtypes = [C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _ip, _dp, _dp, C.c_double, C.c_int, _dp,
                                     C.c_double, C.c_int64, _dp, C.c_double, C.c_double, C.c_uint64, C.c_uint64, _dp]
        _lib.orc_threads.restype = C.c_int
        _lib.orc_run_timing.restype = C.c_int
        _lib.orc_run_timing.argtypes = [C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _ip, _dp, _dp, C.c_double, C.c_int,
                                        _dp, C.c_double, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_uint64, _dp]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _params(p):
    q = np.zeros(8)
    p = np.asarray(p, dtype=np.float64).ravel()
    q[: p.size] = p
    return q


def evaluate(tag, params, r, s1=1.0, s2=1.0):
    u, f = C.c_double(), C.c_double()
    q = _params(params)
    inr = lib().orc_evaluate(tag, _d(q), float(r), float(s1), float(s2), C.byref(u), C.byref(f))
    return u.value, f.value, inr


def forces(x, diam, box, cutoff, tag, params=(), brute=False, counts=False):
    """-> dict(F [n][dim], E, W, n_cut, n_int[, nbr])"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, dim = x.shape
    diam = np.ascontiguousarray(diam, dtype=np.float64)
    box = np.ascontiguousarray(box, dtype=np.float64)
    q = _params(params)
    F = np.zeros_like(x)
    E, W = C.c_double(), C.c_double()
    ncut, nint = C.c_int64(), C.c_int64()
    nbr = np.zeros(n, dtype=np.int32) if counts else None
    fn = lib().orc_forces_brute if brute else lib().orc_forces
    rc = fn(dim, n, _d(x), _d(diam), _d(box), float(cutoff), tag, _d(q), _d(F), C.byref(E), C.byref(W),
            C.byref(ncut), C.byref(nint), _i(nbr) if counts else None)
    if rc != 0:
        raise RuntimeError("oracle forces failed rc=%d" % rc)
    out = dict(F=F, E=E.value, W=W.value, n_cut=ncut.value, n_int=nint.value)
    if counts:
        out["nbr"] = nbr
    return out


def wrap(x, img, box):
    x = np.array(x, dtype=np.float64)
    img = np.array(img, dtype=np.int32)
    box = np.ascontiguousarray(box, dtype=np.float64)
    lib().orc_wrap(x.size, _d(x), _i(img), _d(box))
    return x, img


def kinetic(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return lib().orc_kinetic(v.shape[1], v.shape[0], _d(v))


def bussi_scale(ke, ktemp, nf, dt, tau, r1, r2):
    return lib().orc_bussi_scale(ke, ktemp, nf, dt, tau, r1, r2)


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)


def bussi_noises(seed, step, nf):
    r1, r2 = C.c_double(), C.c_double()
    lib().orc_bussi_noises(seed, step, float(nf), C.byref(r1), C.byref(r2))
    return r1.value, r2.value


def thermo_normals(seed, step, count):
    out = np.zeros(count)
    lib().orc_thermo_normals(seed, step, count, _d(out))
    return out


def chi2(seed, step, nf):
    return lib().orc_chi2(seed, step, float(nf))


def brownian_noise(seed, step, pid, dim):
    out = np.zeros(3)
    lib().orc_brownian_noise(seed, step, pid, dim, _d(out))
    return out[:dim]


def cell3(cell, dim):
    """dim x dim cell matrix (lattice vectors in the columns) embedded row-major in 3x3"""
    U = np.eye(3)
    U[:dim, :dim] = np.asarray(cell, dtype=np.float64)[:dim, :dim]
    return np.ascontiguousarray(U.ravel())


def cell_inverse(cell, dim):
    U = cell3(cell, dim)
    Ui = np.zeros(9)
    det = lib().orc_cell_inverse(_d(U), _d(Ui))
    return Ui.reshape(3, 3), det


def wrap_tri(x, img, cell):
    """wrap_to_box with a full matrix (src/boundary.jl:7-17), row by row"""
    x = np.array(x, dtype=np.float64)
    img = np.array(img, dtype=np.int32)
    n, dim = x.shape
    U = cell3(cell, dim)
    Ui = np.ascontiguousarray(cell_inverse(cell, dim)[0].ravel())
    for i in range(n):
        xi, ii = np.ascontiguousarray(x[i]), np.ascontiguousarray(img[i])
        lib().orc_wrap_tri(dim, _d(xi), _i(ii), _d(U), _d(Ui))
        x[i], img[i] = xi, ii
    return x, img


def forces_tri(x, diam, cell, cutoff, tag, params=()):
    """all-pairs enumeration under a general unit cell -> dict(F, E, W, n_cut, n_int)"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, dim = x.shape
    diam = np.ascontiguousarray(diam, dtype=np.float64)
    U, q = cell3(cell, dim), _params(params)
    F = np.zeros_like(x)
    E, W = C.c_double(), C.c_double()
    ncut, nint = C.c_int64(), C.c_int64()
    rc = lib().orc_forces_tri(dim, n, _d(x), _d(diam), _d(U), float(cutoff), tag, _d(q), _d(F), C.byref(E), C.byref(W),
                              C.byref(ncut), C.byref(nint))
    if rc != 0:
        raise RuntimeError("oracle forces_tri failed rc=%d" % rc)
    return dict(F=F, E=E.value, W=W.value, n_cut=ncut.value, n_int=nint.value)


def run_tri(ensemble, x, v, f, img, diam, cell, cutoff, tag, params, dt, nsteps, ktemp=None, tau=1.0, nf=None, seed=0,
            rng_step0=0):
    """orc_run with a general unit cell; returns (x, v, f, img, thermo)"""
    x, v, f = (np.array(a, dtype=np.float64) for a in (x, v, f))
    img = np.array(img, dtype=np.int32)
    n, dim = x.shape
    diam = np.ascontiguousarray(diam, dtype=np.float64)
    U, q = cell3(cell, dim), _params(params)
    if nf is None:
        nf = dim * (n - 1.0)
    kt = np.ascontiguousarray(np.broadcast_to(np.asarray(1.0 if ktemp is None else ktemp, dtype=np.float64), (max(nsteps, 1),)))
    thermo = np.zeros((nsteps, 4))
    rc = lib().orc_run_tri(ensemble, dim, n, _d(x), _d(v), _d(f), _i(img), _d(diam), _d(U), float(cutoff), tag, _d(q), dt, nsteps,
                           _d(kt), tau, nf, seed, rng_step0, _d(thermo))
    if rc != 0:
        raise RuntimeError("oracle run_tri failed rc=%d" % rc)
    return x, v, f, img, thermo


def random_positions(dim, n, box, seed, stream=0):
    """src/initialization.jl:22-27: uniform points of the cell from the counter-based generator; returns (n, dim)"""
    x = np.zeros((n, dim))
    b = np.ascontiguousarray(np.asarray(box, dtype=np.float64) * np.ones(3))
    lib().orc_random_positions(dim, n, _d(b), seed, stream, _d(x))
    return x


def init_velocities(dim, n, ktemp, seed, stream=0):
    """src/initialization.jl:32-47 with the counter-based normals; returns (n, dim)"""
    v = np.zeros((n, dim))
    lib().orc_init_velocities(dim, n, float(ktemp), seed, stream, _d(v))
    return v


def run(ensemble, x, v, f, img, diam, box, cutoff, tag, params, dt, nsteps, ktemp=None, tau=1.0, nf=None, seed=0,
        rng_step0=0):
    """Advance copies of (x, v, f, img) by nsteps; returns (x, v, f, img, thermo[nsteps][4] = U, W, KE, n_int)."""
    x = np.array(x, dtype=np.float64, order="C")
    n, dim = x.shape
    v = np.array(v, dtype=np.float64, order="C") if v is not None else np.zeros_like(x)
    f = np.array(f, dtype=np.float64, order="C")
    img = np.array(img, dtype=np.int32, order="C")
    diam = np.ascontiguousarray(diam, dtype=np.float64)
    box = np.ascontiguousarray(box, dtype=np.float64)
    q = _params(params)
    if nf is None:
        nf = dim * (n - 1.0)
    if ktemp is None:
        kt = np.zeros(max(nsteps, 1))
    else:
        kt = np.ascontiguousarray(np.broadcast_to(np.asarray(ktemp, dtype=np.float64), (max(nsteps, 1),)))
    thermo = np.zeros((nsteps, 4))
    rc = lib().orc_run(ensemble, dim, n, _d(x), _d(v), _d(f), _i(img), _d(diam), _d(box), float(cutoff), tag, _d(q),
                       float(dt), nsteps, _d(kt), float(tau), float(nf), seed, rng_step0, _d(thermo))
    if rc != 0:
        raise RuntimeError("oracle run failed rc=%d" % rc)
    return x, v, f, img, thermo


def threads():
    return lib().orc_threads()


def set_threads(n):
    """OpenMP team size of the timing baseline (launchers like torchrun export OMP_NUM_THREADS=1)"""
    lib().orc_set_threads(int(n))


def run_timing(ensemble, x, v, f, img, diam, box, cutoff, tag, params, dt, nsteps, ktemp=1.0, tau=1.0, seed=0):
    """Reference-shaped OpenMP loop, in place on the given arrays; returns (E, W, KE) of the last step."""
    n, dim = x.shape
    q = _params(params)
    out = np.zeros(3)
    nf = dim * (n - 1.0)
    rc = lib().orc_run_timing(ensemble, dim, n, _d(x), _d(v), _d(f), _i(img), _d(diam), _d(box), float(cutoff), tag,
                              _d(q), float(dt), nsteps, float(ktemp), float(tau), nf, seed, _d(out))
    if rc < 0:
        raise RuntimeError("oracle timing run failed rc=%d" % rc)
    return out


def fire(x, img, diam, box, cutoff, tag, params, max_steps=10000, tol=1e-6, dt_initial=0.01, dt_max=0.1, alpha0=0.1, f_inc=1.2,
         f_dec=0.2, n_min=5):
    """fire_minimize! (src/minimize.jl:31-135) on copies; -> (x, img, energy, steps, converged, trace[steps][3])"""
    x = np.array(x, dtype=np.float64, order="C")
    n, dim = x.shape
    img = np.array(img, dtype=np.int32, order="C")
    diam = np.ascontiguousarray(diam, dtype=np.float64)
    box = np.ascontiguousarray(box, dtype=np.float64)
    q = _params(params)
    e, conv = C.c_double(), C.c_int()
    trace = np.zeros((max_steps, 3))
    steps = lib().orc_fire(dim, n, _d(x), _i(img), _d(diam), _d(box), float(cutoff), tag, _d(q), max_steps, tol, dt_initial, dt_max,
                           alpha0, f_inc, f_dec, n_min, C.byref(e), C.byref(conv), _d(trace))
    return x, img, e.value, int(steps), bool(conv.value), trace[:steps]
