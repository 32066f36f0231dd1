"""ctypes binding of the C ABI in include/mdb200.h (csrc/libmdb200.so).

Fails loudly when the CUDA library is missing or no device is present: there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

from . import _build

# enum mdb_status
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED_POTENTIAL, ERR_UNSUPPORTED_CELL, ERR_BOX_TOO_SMALL = 0, 1, 2, 3, 4, 5
ERR_NO_DEVICE, ERR_NCCL, ERR_STATE, ERR_NONFINITE, ERR_NVRTC, ERR_IO = 6, 7, 8, 9, 10, 11
STATUS_NAMES = {0: "MDB_OK", 1: "MDB_ERR_INVALID_ARG", 2: "MDB_ERR_CUDA", 3: "MDB_ERR_UNSUPPORTED_POTENTIAL",
                4: "MDB_ERR_UNSUPPORTED_CELL", 5: "MDB_ERR_BOX_TOO_SMALL", 6: "MDB_ERR_NO_DEVICE", 7: "MDB_ERR_NCCL",
                8: "MDB_ERR_STATE", 9: "MDB_ERR_NONFINITE", 10: "MDB_ERR_NVRTC", 11: "MDB_ERR_IO"}
POT_PSEUDOHS, POT_LJ, POT_LJ_XPLOR, POT_POLY, POT_SOFT, POT_USER = 0, 1, 2, 3, 4, 100
NVE, NVT, BROWNIAN = 0, 1, 2
MODE_AUTO, MODE_CELLS, MODE_LIST, MODE_SMALL = 0, 1, 2, 3

# every symbol include/mdb200.h declares (tests/test_capi_symbols.py checks the header against this list and the .so)
SYMBOLS = [
    "mdb_version", "mdb_last_error", "mdb_create", "mdb_destroy", "mdb_upload", "mdb_set_velocities", "mdb_download",
    "mdb_download_owned", "mdb_upload_owned", "mdb_compute_forces", "mdb_count_pairs", "mdb_run_nve", "mdb_run_nvt", "mdb_run_brownian",
    "mdb_thermo", "mdb_fire_minimize", "mdb_bussi_scale_from", "mdb_bussi_noises", "mdb_get_rng_step",
    "mdb_set_rng_step", "mdb_set_user_potential", "mdb_comm_unique_id", "mdb_comm_init", "mdb_comm_init_local",
    "mdb_get_stats", "mdb_device_ptr", "mdb_stream", "mdb_synchronize",
    "mdb_frame_capture", "mdb_frame_wait", "mdb_frame_write_lammps", "mdb_frame_flush",
    "mdb_init_velocities", "mdb_random_positions", "mdb_checkpoint_save", "mdb_checkpoint_load", "mdb_measure_fp64_peak",
    "mdb_force_kernel_info",
]
FRAME_SLOTS = 2


class MdbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s: %s" % (STATUS_NAMES.get(code, code), msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [("dim", C.c_int32), ("potential", C.c_int32), ("n_particles", C.c_int64), ("unitcell", C.c_double * 9),
                ("cutoff", C.c_double), ("pot_params", C.c_double * 8), ("seed", C.c_uint64), ("device", C.c_int32),
                ("mode", C.c_int32), ("skin", C.c_double), ("use_graph", C.c_int32), ("rank", C.c_int32),
                ("nranks", C.c_int32), ("no_fuse", C.c_int32), ("skin_inner", C.c_double), ("slab_transport", C.c_int32),
                ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("rebuilds", C.c_int64), ("kernel_launches", C.c_int64), ("n_owned", C.c_int64),
                ("n_ghost", C.c_int64), ("list_capacity", C.c_int64), ("max_neighbors", C.c_int64),
                ("r_search", C.c_double), ("cell_len", C.c_double * 3), ("ncell", C.c_int32 * 3), ("mode", C.c_int32),
                ("last_run_ms", C.c_double), ("last_force_ms", C.c_double),
                ("prof_kick_ms", C.c_double), ("prof_force_ms", C.c_double), ("prof_rebuild_ms", C.c_double),
                ("prof_steps", C.c_int64), ("slab_transport", C.c_int64), ("slab_graph", C.c_int64)]


class FireParams(C.Structure):
    _fields_ = [("max_steps", C.c_int64), ("tol", C.c_double), ("dt_initial", C.c_double), ("dt_max", C.c_double),
                ("alpha0", C.c_double), ("f_inc", C.c_double), ("f_dec", C.c_double), ("n_min", C.c_int32),
                ("reserved", C.c_int32)]


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_H = C.c_void_p


def lib_path():
    return _build.LIB


def load():
    """Load csrc/libmdb200.so; raises if it has not been built (build with __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError("%s is missing: build it with `python __graft_entry__.py build` "
                           "(nvcc, sm_100a). There is no CPU fallback." % path)
    L = C.CDLL(path)
    L.mdb_version.restype = C.c_int
    L.mdb_last_error.restype = C.c_char_p
    L.mdb_last_error.argtypes = [_H]
    L.mdb_create.argtypes = [C.POINTER(Config), C.POINTER(_H)]
    L.mdb_destroy.argtypes = [_H]
    L.mdb_upload.argtypes = [_H, _dp, _dp, _dp, _dp, _ip]
    L.mdb_set_velocities.argtypes = [_H, _dp]
    L.mdb_download.argtypes = [_H, _dp, _dp, _dp, _ip]
    L.mdb_download_owned.argtypes = [_H, C.c_int64, _ip, _dp, _dp, _dp, _ip, C.POINTER(C.c_int64)]
    L.mdb_upload_owned.argtypes = [_H, C.c_int64, _ip, _dp, _dp, _dp, _dp, _ip]
    L.mdb_compute_forces.argtypes = [_H, _dp, _dp, C.POINTER(C.c_int64)]
    L.mdb_count_pairs.argtypes = [_H, C.c_double, C.POINTER(C.c_int64), _ip]
    L.mdb_run_nve.argtypes = [_H, C.c_int64, C.c_double, _dp]
    L.mdb_run_nvt.argtypes = [_H, C.c_int64, C.c_double, _dp, C.c_double, _dp]
    L.mdb_run_brownian.argtypes = [_H, C.c_int64, C.c_double, C.c_double, _dp]
    L.mdb_thermo.argtypes = [_H, _dp]
    L.mdb_fire_minimize.argtypes = [_H, C.POINTER(FireParams), _dp, _ip]
    L.mdb_bussi_scale_from.argtypes = [_H] + [C.c_double] * 7 + [_dp]
    L.mdb_bussi_noises.argtypes = [_H, C.c_uint64, C.c_double, _dp, _dp]
    L.mdb_get_rng_step.argtypes = [_H, C.POINTER(C.c_uint64)]
    L.mdb_set_rng_step.argtypes = [_H, C.c_uint64]
    L.mdb_set_user_potential.argtypes = [_H, C.c_char_p, _dp, C.c_int32, C.c_double]
    L.mdb_comm_unique_id.argtypes = [C.c_char_p]
    L.mdb_comm_init.argtypes = [_H, C.c_char_p]
    L.mdb_comm_init_local.argtypes = [C.POINTER(_H), C.c_int32]
    L.mdb_get_stats.argtypes = [_H, C.POINTER(Stats)]
    L.mdb_device_ptr.argtypes = [_H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.mdb_stream.argtypes = [_H, C.POINTER(C.c_void_p)]
    L.mdb_synchronize.argtypes = [_H]
    L.mdb_frame_capture.argtypes = [_H, C.c_int32]
    L.mdb_frame_wait.argtypes = [_H, C.c_int32, C.POINTER(_dp), C.POINTER(C.c_int32)]
    L.mdb_frame_write_lammps.argtypes = [_H, C.c_int32, C.c_char_p, C.c_int64, C.c_int32]
    L.mdb_frame_flush.argtypes = [_H]
    L.mdb_init_velocities.argtypes = [_H, C.c_double, C.c_uint64]
    L.mdb_random_positions.argtypes = [_H, C.c_uint64]
    L.mdb_checkpoint_save.argtypes = [_H, C.c_char_p]
    L.mdb_checkpoint_load.argtypes = [_H, C.c_char_p]
    L.mdb_measure_fp64_peak.argtypes = [_H, _dp]
    L.mdb_force_kernel_info.argtypes = [_H, _ip]
    for name in SYMBOLS:
        if name not in ("mdb_last_error",):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError("expected array of shape %s, got %s" % (shape, a.shape))
    return a


def unique_id():
    """128-byte NCCL unique id (rank 0 creates it and ships it to the other ranks, e.g. torch.distributed.broadcast)"""
    buf = C.create_string_buffer(128)
    rc = load().mdb_comm_unique_id(buf)
    if rc != OK:
        raise MdbError(rc, (load().mdb_last_error(None) or b"").decode())
    return buf.raw


class Engine:
    """Thin object wrapper over one mdb_handle (one GPU, one host thread)."""

    def __init__(self, dim, n_particles, box, cutoff, potential, pot_params=(), seed=0, device=0, mode=MODE_AUTO,
                 skin=0.0, use_graph=True, rank=0, nranks=1, skin_inner=0.0, no_fuse=False, slab_transport=0):
        L = load()
        cfg = Config()
        cfg.dim = dim
        cfg.potential = potential
        cfg.n_particles = n_particles
        box = np.asarray(box, dtype=np.float64)
        cell = np.zeros((3, 3))
        if box.ndim == 0:
            cell[:dim, :dim] = np.eye(dim) * float(box)
        elif box.ndim == 1:
            cell[:dim, :dim] = np.diag(box[:dim])
        else:
            cell[:dim, :dim] = box[:dim, :dim]
        cfg.unitcell = (C.c_double * 9)(*cell.ravel())
        cfg.cutoff = cutoff
        pp = np.zeros(8)
        pot_params = np.asarray(pot_params, dtype=np.float64).ravel()
        pp[: pot_params.size] = pot_params
        cfg.pot_params = (C.c_double * 8)(*pp)
        cfg.seed = seed
        cfg.device = device
        cfg.mode = mode
        cfg.skin = skin or 0.0
        cfg.skin_inner = skin_inner or 0.0
        cfg.use_graph = 1 if use_graph else 0
        cfg.no_fuse = 1 if no_fuse else 0
        cfg.rank = rank
        cfg.nranks = nranks
        cfg.slab_transport = slab_transport
        self._lib = L
        self._nranks = max(1, nranks)
        self.dim = dim
        self.n = n_particles
        self.box = np.array([cell[k, k] for k in range(dim)])
        self._h = _H()
        rc = L.mdb_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            raise MdbError(rc, (L.mdb_last_error(None) or b"").decode())

    def _check(self, rc):
        if rc != OK:
            raise MdbError(rc, (self._lib.mdb_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mdb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # state -------------------------------------------------------------------------------------------
    def upload(self, positions, diameters, velocities=None, forces=None, images=None):
        shape = (self.n, self.dim)
        x = _f64(positions, shape)
        d = _f64(diameters, (self.n,))
        v = _f64(velocities, shape)
        f = _f64(forces, shape)
        im = None if images is None else np.ascontiguousarray(images, dtype=np.int32)
        self._check(self._lib.mdb_upload(self._h, _d(x), _d(v), _d(f), _d(d), _i(im)))

    def upload_owned(self, ids, positions, diameters, velocities=None, forces=None, images=None):
        """slab handle: re-upload the rows this rank owns (as returned by download_owned; `diameters` per row)"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        m = ids.size
        x = _f64(positions, (m, self.dim))
        v = _f64(velocities, (m, self.dim))
        f = _f64(forces, (m, self.dim))
        dm = _f64(diameters, (m,))
        im = None if images is None else np.ascontiguousarray(images, dtype=np.int32)
        self._check(self._lib.mdb_upload_owned(self._h, m, _i(ids), _d(x), _d(v), _d(f), _d(dm), _i(im)))

    def set_velocities(self, velocities):
        v = _f64(velocities, (self.n, self.dim))
        self._check(self._lib.mdb_set_velocities(self._h, _d(v)))

    def download(self, positions=True, velocities=True, forces=True, images=True):
        shape = (self.n, self.dim)
        x = np.empty(shape) if positions else None
        v = np.empty(shape) if velocities else None
        f = np.empty(shape) if forces else None
        im = np.empty(shape, dtype=np.int32) if images else None
        self._check(self._lib.mdb_download(self._h, _d(x), _d(v), _d(f), _i(im)))
        return x, v, f, im

    def download_into(self, x=None, v=None, f=None, im=None):
        self._check(self._lib.mdb_download(self._h, _d(x), _d(v), _d(f), _i(im)))

    # physics -----------------------------------------------------------------------------------------
    def compute_forces(self):
        e, w, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self._lib.mdb_compute_forces(self._h, C.byref(e), C.byref(w), C.byref(n)))
        return e.value, w.value, n.value

    def count_pairs(self, cutoff, per_particle=False):
        n = C.c_int64()
        per = np.zeros(self.n, dtype=np.int32) if per_particle else None
        self._check(self._lib.mdb_count_pairs(self._h, cutoff, C.byref(n), _i(per)))
        return (n.value, per) if per_particle else n.value

    def run_nve(self, nsteps, dt, thermo=True):
        t = np.zeros((nsteps, 4)) if thermo else None
        self._check(self._lib.mdb_run_nve(self._h, nsteps, dt, _d(t)))
        return t

    def run_nvt(self, nsteps, dt, ktemp, tau, thermo=True):
        kt = np.ascontiguousarray(np.broadcast_to(np.asarray(ktemp, dtype=np.float64), (max(nsteps, 1),)))
        t = np.zeros((nsteps, 4)) if thermo else None
        self._check(self._lib.mdb_run_nvt(self._h, nsteps, dt, _d(kt), tau, _d(t)))
        return t

    def run_brownian(self, nsteps, dt, ktemp, thermo=True):
        t = np.zeros((nsteps, 4)) if thermo else None
        self._check(self._lib.mdb_run_brownian(self._h, nsteps, dt, ktemp, _d(t)))
        return t

    def thermo(self):
        out = np.zeros(4)
        self._check(self._lib.mdb_thermo(self._h, _d(out)))
        return out

    def fire_minimize(self, max_steps=10000, tol=1e-6, dt_initial=0.01, dt_max=0.1, alpha0=0.1, f_inc=1.2, f_dec=0.2,
                      n_min=5):
        p = FireParams(max_steps, tol, dt_initial, dt_max, alpha0, f_inc, f_dec, n_min, 0)
        out = np.zeros(3)
        conv = C.c_int32()
        self._check(self._lib.mdb_fire_minimize(self._h, C.byref(p), _d(out), C.byref(conv)))
        return out[0], out[1], int(out[2]), bool(conv.value)

    # ---- trajectory frames ----
    def frame_capture(self, slot=0):
        """pack {radius, x, unwrapped x} per particle on the device (original order) and start the copy to pinned host
        memory on the copy stream; returns without waiting"""
        self._check(self._lib.mdb_frame_capture(self._h, slot))

    def frame_wait(self, slot=0):
        """(n, 2*dim+1) view of the slot's pinned host frame once its copy has finished (valid until the next capture)"""
        ptr, w = _dp(), C.c_int32()
        self._check(self._lib.mdb_frame_wait(self._h, slot, C.byref(ptr), C.byref(w)))
        rows = self.n if self._nranks == 1 else int(self.stats()["n_owned"])   # slab handle: the rows it owned at capture time
        return np.ctypeslib.as_array(ptr, shape=(rows, w.value))

    def frame_write_lammps(self, slot, path, step, append=True):
        """queue the frame for the library's writer thread (write_to_file_lammps layout, src/io.jl:78-170)"""
        self._check(self._lib.mdb_frame_write_lammps(self._h, slot, os.fsencode(path), int(step), 1 if append else 0))

    def frame_flush(self):
        self._check(self._lib.mdb_frame_flush(self._h))

    # ---- device-side set-up and exact restart ----
    def init_velocities(self, ktemp, stream=0):
        """initialize_velocities (src/initialization.jl:32-47) on the device"""
        self._check(self._lib.mdb_init_velocities(self._h, float(ktemp), int(stream)))

    def random_positions(self, stream=0):
        """uniform random positions in the cell, drawn on the device (src/initialization.jl:20-27)"""
        self._check(self._lib.mdb_random_positions(self._h, int(stream)))

    def checkpoint_save(self, path):
        self._check(self._lib.mdb_checkpoint_save(self._h, os.fsencode(path)))

    def checkpoint_load(self, path):
        self._check(self._lib.mdb_checkpoint_load(self._h, os.fsencode(path)))

    def measure_fp64_peak(self):
        """measured DFMA throughput of the device, TFLOP/s"""
        t = C.c_double()
        self._check(self._lib.mdb_measure_fp64_peak(self._h, C.byref(t)))
        return t.value

    def force_kernel_info(self):
        """identity of the dominant (fused NVE pair-force) kernel as built into the loaded library"""
        a = np.zeros(6, dtype=np.int32)
        self._check(self._lib.mdb_force_kernel_info(self._h, _i(a)))
        return dict(zip(("registers", "static_smem_bytes", "ctas_per_sm", "variant", "threads_per_cta", "local_bytes"), (int(v) for v in a)))

    def bussi_scale_from(self, ke, ktemp, nf, dt, tau, r1, r2):
        s = C.c_double()
        self._check(self._lib.mdb_bussi_scale_from(self._h, ke, ktemp, nf, dt, tau, r1, r2, C.byref(s)))
        return s.value

    def bussi_noises(self, step, nf):
        r1, r2 = C.c_double(), C.c_double()
        self._check(self._lib.mdb_bussi_noises(self._h, step, nf, C.byref(r1), C.byref(r2)))
        return r1.value, r2.value

    @property
    def rng_step(self):
        s = C.c_uint64()
        self._check(self._lib.mdb_get_rng_step(self._h, C.byref(s)))
        return s.value

    @rng_step.setter
    def rng_step(self, value):
        self._check(self._lib.mdb_set_rng_step(self._h, value))

    def set_user_potential(self, body, params=(), rng=1.0):
        p = np.ascontiguousarray(params, dtype=np.float64)
        self._check(self._lib.mdb_set_user_potential(self._h, body.encode(), _d(p), p.size, rng))

    def download_owned(self):
        """-> (ids, x, v, f, img) of the particles this slab currently owns (device slot order)"""
        # the owned count can change at rebuilds (migration moves a thin layer per rebuild): room for 25 % more
        cap = max(int(int(self.stats()["n_owned"]) * 1.25) + 4096, 4096)
        ids = np.empty(cap, dtype=np.int32)
        x = np.empty((cap, self.dim))
        v = np.empty((cap, self.dim))
        f = np.empty((cap, self.dim))
        im = np.empty((cap, self.dim), dtype=np.int32)
        cnt = C.c_int64()
        self._check(self._lib.mdb_download_owned(self._h, cap, _i(ids), _d(x), _d(v), _d(f), _i(im), C.byref(cnt)))
        k = cnt.value
        return ids[:k], x[:k], v[:k], f[:k], im[:k]

    def download_owned_into(self, ids, x=None, v=None, f=None, im=None):
        """like download_owned, into caller-owned buffers (e.g. pinned host memory) of at least n_owned rows; returns the
        number of rows written.  The buffers are reused call after call: nothing is allocated and the copies run at
        pinned-memory speed."""
        cnt = C.c_int64()
        self._check(self._lib.mdb_download_owned(self._h, ids.shape[0], _i(ids), _d(x), _d(v), _d(f), _i(im), C.byref(cnt)))
        return cnt.value

    def comm_init(self, unique_id):
        """join the NCCL slab ring (one process per GPU); unique_id = bytes from unique_id() on rank 0"""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.mdb_comm_init(self._h, buf))

    def stats(self):
        s = Stats()
        self._check(self._lib.mdb_get_stats(self._h, C.byref(s)))
        return {k: (list(getattr(s, k)) if hasattr(getattr(s, k), "__len__") else getattr(s, k)) for k, _ in s._fields_}

    def device_ptr(self, which):
        p, st = C.c_void_p(), C.c_int64()
        self._check(self._lib.mdb_device_ptr(self._h, which, C.byref(p), C.byref(st)))
        return p.value, st.value

    def stream(self):
        p = C.c_void_p()
        self._check(self._lib.mdb_stream(self._h, C.byref(p)))
        return p.value

    def synchronize(self):
        self._check(self._lib.mdb_synchronize(self._h))


class SlabRing:
    """x-slab decomposition driver.

    * in-process ring (`SlabRing.local(P, ...)`): P slab engines on ONE device sharing a stream, exchanging through
      device copies -- the single-GPU emulation of the multi-rank protocol used by the tests;
    * NCCL ring (`SlabRing.nccl(rank, world, unique_id, ...)`): one process per GPU; every call is collective.
    Run calls return GLOBAL thermo rows (identical on every rank)."""

    def __init__(self, engines, lead):
        self.engines = engines
        self.lead = lead

    @classmethod
    def local(cls, nranks, dim, n_particles, box, cutoff, potential, pot_params=(), **kw):
        kw.setdefault("use_graph", True)   # only the peer-memory transport (slab_transport=2) replays its step as a graph
        engines = [Engine(dim, n_particles, box, cutoff, potential, pot_params, rank=r, nranks=nranks, **kw)
                   for r in range(nranks)]
        arr = (_H * nranks)(*[e._h for e in engines])
        rc = load().mdb_comm_init_local(arr, nranks)
        if rc != OK:
            raise MdbError(rc, (load().mdb_last_error(engines[0]._h) or b"").decode())
        return cls(engines, engines[0])

    @classmethod
    def nccl(cls, rank, nranks, uid, dim, n_particles, box, cutoff, potential, pot_params=(), **kw):
        kw.setdefault("use_graph", True)
        e = Engine(dim, n_particles, box, cutoff, potential, pot_params, rank=rank, nranks=nranks, **kw)
        e.comm_init(uid)
        return cls([e], e)

    def upload(self, positions, diameters, velocities=None, forces=None, images=None):
        for e in self.engines:   # every rank is shown the global arrays and keeps its slab
            e.upload(positions, diameters, velocities, forces, images)

    def set_velocities(self, v):
        for e in self.engines:
            e.set_velocities(v)

    def compute_forces(self):
        return self.lead.compute_forces()

    def run_nve(self, nsteps, dt, thermo=True):
        return self.lead.run_nve(nsteps, dt, thermo)

    def run_nvt(self, nsteps, dt, ktemp, tau, thermo=True):
        return self.lead.run_nvt(nsteps, dt, ktemp, tau, thermo)

    def run_brownian(self, nsteps, dt, ktemp, thermo=True):
        return self.lead.run_brownian(nsteps, dt, ktemp, thermo)

    def upload_owned(self, parts):
        """re-upload owned rows without the global arrays: `parts` = one (ids, x, diameters, v, f, img) tuple per slab this
        process holds (a single tuple is accepted for the one-slab-per-process NCCL case)"""
        if isinstance(parts, tuple):
            parts = [parts]
        for e, (ids, x, diam, v, f, im) in zip(self.engines, parts):
            e.upload_owned(ids, x, diam, velocities=v, forces=f, images=im)

    def download_local(self):
        """owned particles of the slabs this process holds, concatenated: (ids, x, v, f, img)"""
        parts = [e.download_owned() for e in self.engines]
        if len(parts) == 1:
            return parts[0]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(5))

    def download_local_into(self, ids, x=None, v=None, f=None, im=None):
        """one-slab-per-process (NCCL / peer ring) form of download_local into caller-owned (pinned) buffers -> row count"""
        if len(self.engines) != 1:
            raise RuntimeError("download_local_into serves one slab per process; use download_local for an in-process ring")
        return self.engines[0].download_owned_into(ids, x, v, f, im)

    def download(self):
        """global arrays in original particle order (in-process ring only; with NCCL gather download_local across ranks)"""
        ids, x, v, f, im = self.download_local()
        n = self.lead.n
        if ids.size != n or np.unique(ids).size != n:
            raise RuntimeError("slab ownership is not a partition: %d ids for %d particles" % (ids.size, n))
        order = np.argsort(ids)
        return x[order], v[order], f[order], im[order]

    def stats(self):
        return [e.stats() for e in self.engines]

    def init_velocities(self, ktemp, stream=0):
        """initialize_velocities on the device, collective over the ring"""
        self.lead.init_velocities(ktemp, stream)

    def checkpoint_save(self, path):
        """every slab writes <path>.<rank> (collective by convention)"""
        for e in self.engines:
            e.checkpoint_save(path)

    def checkpoint_load(self, path):
        for e in self.engines:
            e.checkpoint_load(path)

    def frame_capture(self, slot=0):
        for e in self.engines:
            e.frame_capture(slot)

    def frame_write_lammps(self, slot, path, step, append=True):
        """one LAMMPS dump per slab: <path>.<rank>"""
        for e in self.engines:
            e.frame_write_lammps(slot, path, step, append)

    def frame_flush(self):
        for e in self.engines:
            e.frame_flush()

    def close(self):
        for e in self.engines:
            e.close()
