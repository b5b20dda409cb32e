"""admm-elastic_b200: B200-native ADMM-elastic time step behind the reference's plugin surface.

The product is native code: CUDA kernels + C-ABI (``libadmm_b200.so``, include/admm_b200.h) and
the C++ host mirror of ``admm::Solver`` (``libadmm_b200_host.so``, host/admm_b200.hpp).  This
Python module is only a ctypes view of those two libraries for tests and bench.py:

* ``Solver``        -> admm_b200::Solver through host/c_api.cpp (the call a user makes)
* ``DeviceSolver``  -> the raw C-ABI handle (kernel-level entry points used by parity tests)

There is no CPU fallback anywhere: if the native libraries are missing, import fails loudly; if no
CUDA device is usable, ``Solver.initialize`` / ``DeviceSolver()`` raise.
"""
import ctypes
import os

import numpy as np

from . import meshes  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_CUDA_PATH = os.path.join(_HERE, "libadmm_b200.so")
LIB_HOST_PATH = os.path.join(_HERE, "libadmm_b200_host.so")

TET_LINEAR, TET_NEOHOOKEAN, TET_STVK, TET_SPLINE_NH, TET_SPLINE_STVK, TET_SPLINE_COROT = range(6)
LDLT, MCGS, UZAWA = 0, 1, 2
FP32, FP64 = 0, 1
COLOR_GREEDY, COLOR_RANDOM, COLOR_USER = 0, 1, 2
IPC_BYTES = 256

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int)


def _dp(a):
    return a.ctypes.data_as(_c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_c_int_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class NativeLibraryMissing(ImportError):
    pass


def _load():
    for p in (LIB_CUDA_PATH, LIB_HOST_PATH):
        if not os.path.exists(p):
            raise NativeLibraryMissing(
                "%s is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no Python/CPU fallback for the CUDA path)" % p)
    cuda = ctypes.CDLL(LIB_CUDA_PATH, mode=ctypes.RTLD_GLOBAL)
    host = ctypes.CDLL(LIB_HOST_PATH)
    cuda.admm_b200_last_error.restype = ctypes.c_char_p
    cuda.admm_b200_last_error.argtypes = [ctypes.c_void_p]
    cuda.admm_b200_launch_count.restype = ctypes.c_longlong
    cuda.admm_b200_launch_count.argtypes = [ctypes.c_void_p]
    cuda.admm_b200_solver_info.restype = ctypes.c_char_p
    cuda.admm_b200_solver_info.argtypes = [ctypes.c_void_p]
    host.admmhost_create.restype = ctypes.c_void_p
    host.admmhost_last_error.restype = ctypes.c_char_p
    host.admmhost_last_error.argtypes = [ctypes.c_void_p]
    host.admmhost_device_handle.restype = ctypes.c_void_p
    host.admmhost_device_handle.argtypes = [ctypes.c_void_p]
    host.admmhost_x_ptr.restype = ctypes.c_void_p
    return cuda, host


_cuda, _host = _load()
cuda_lib, host_lib = _cuda, _host


class AdmmError(RuntimeError):
    pass


class DeviceSolver(object):
    """Raw C-ABI handle (include/admm_b200.h)."""

    def __init__(self, device=0, handle=None):
        self._own = handle is None
        if handle is None:
            h = ctypes.c_void_p()
            rc = _cuda.admm_b200_create(int(device), ctypes.byref(h))
            if rc:
                raise AdmmError(_cuda.admm_b200_last_error(None).decode())
            handle = h
        self.h = handle if isinstance(handle, ctypes.c_void_p) else ctypes.c_void_p(handle)

    def close(self):
        if self._own and self.h:
            _cuda.admm_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise AdmmError(_cuda.admm_b200_last_error(self.h).decode())

    def set_stream(self, stream_ptr):
        self._ck(_cuda.admm_b200_set_stream(self.h, ctypes.c_void_p(stream_ptr)))

    def synchronize(self):
        self._ck(_cuda.admm_b200_synchronize(self.h))

    def set_nodes(self, x, m, v=None):
        x, m = _f64(x).ravel(), _f64(m).ravel()
        v = _f64(v).ravel() if v is not None else None
        self._ck(_cuda.admm_b200_set_nodes(self.h, x.size // 3, _dp(x), _dp(v), _dp(m)))

    def add_tets(self, idx, dminv, weight, model, mu, lam, kappa=0.0, row_offset=None, bulk_modulus=0.0):
        idx, dminv, weight = _i32(idx).ravel(), _f64(dminv).ravel(), _f64(weight).ravel()
        ro = _i32(row_offset).ravel() if row_offset is not None else None
        self._ck(_cuda.admm_b200_add_tets(self.h, weight.size, _ip(idx), _dp(dminv), _dp(weight), int(model),
                                          ctypes.c_double(mu), ctypes.c_double(lam), ctypes.c_double(kappa), ctypes.c_double(bulk_modulus), _ip(ro)))

    def set_system(self, rowptr, cols, vals):
        rowptr, cols, vals = _i32(rowptr), _i32(cols), _f64(vals)
        self._ck(_cuda.admm_b200_set_system(self.h, rowptr.size - 1, _ip(rowptr), _ip(cols), _dp(vals)))

    def set_colors(self, offsets, nodes):
        offsets, nodes = _i32(offsets), _i32(nodes)
        self._ck(_cuda.admm_b200_set_colors(self.h, offsets.size - 1, _ip(offsets), _ip(nodes)))

    def set_gs_pins(self, idx, pos):
        idx, pos = _i32(idx).ravel(), _f64(pos).ravel()
        self._ck(_cuda.admm_b200_set_gs_pins(self.h, idx.size, _ip(idx), _dp(pos)))

    def add_obstacle(self, kind, params):
        p = _f64(list(params) + [0.0] * (4 - len(params)))
        self._ck(_cuda.admm_b200_add_obstacle(self.h, int(kind), _dp(p)))

    def finalize(self, dt, linsolver, gs_iters=30, gs_omega=1.9, gs_tol=1e-10, precision=FP32):
        self._ck(_cuda.admm_b200_finalize(self.h, ctypes.c_double(dt), int(linsolver), int(gs_iters),
                                          ctypes.c_double(gs_omega), ctypes.c_double(gs_tol), int(precision)))

    def step(self, admm_iters, gravity):
        self._ck(_cuda.admm_b200_step(self.h, int(admm_iters), ctypes.c_double(gravity), None))

    def step_host(self, admm_iters, gravity, x, v):
        """x, v: C-contiguous float64 arrays (ideally pinned), updated in place."""
        self._ck(_cuda.admm_b200_step_host(self.h, int(admm_iters), ctypes.c_double(gravity), _dp(x), _dp(v), None))

    def step_host_ptr(self, admm_iters, gravity, x_ptr, v_ptr):
        self._ck(_cuda.admm_b200_step_host(self.h, int(admm_iters), ctypes.c_double(gravity),
                                           ctypes.cast(x_ptr, _c_double_p), ctypes.cast(v_ptr, _c_double_p), None))

    def upload_state(self, x, v):
        x, v = _f64(x).ravel(), _f64(v).ravel()
        self._ck(_cuda.admm_b200_upload_state(self.h, _dp(x), _dp(v)))

    def download_state(self, n_nodes):
        x, v = np.empty(3 * n_nodes), np.empty(3 * n_nodes)
        self._ck(_cuda.admm_b200_download_state(self.h, _dp(x), _dp(v)))
        return x, v

    def prox_tets(self, model, mu, lam, z, kappa=0.0, precision=FP32, bulk_modulus=0.0):
        """bulk_modulus: K of the prox penalty (the element's Lame); 0 = lambda + 2/3 mu of the model constants."""
        z = _f64(z).reshape(-1, 9)
        out = np.empty_like(z)
        self._ck(_cuda.admm_b200_prox_tets(self.h, int(model), ctypes.c_double(mu), ctypes.c_double(lam),
                                           ctypes.c_double(kappa), ctypes.c_double(bulk_modulus), int(precision), z.shape[0], _dp(z), _dp(out)))
        return out

    def prox_tris(self, z, limit_min=-100.0, limit_max=100.0, precision=FP32):
        z = _f64(z).reshape(-1, 6)
        out = np.empty_like(z)
        self._ck(_cuda.admm_b200_prox_tris(self.h, ctypes.c_double(limit_min), ctypes.c_double(limit_max),
                                           int(precision), z.shape[0], _dp(z), _dp(out)))
        return out

    def linsolve(self, x, b):
        x = _f64(x).ravel().copy()
        b = _f64(b).ravel()
        it = ctypes.c_int(0)
        self._ck(_cuda.admm_b200_linsolve(self.h, _dp(x), _dp(b), ctypes.byref(it)))
        return x, it.value

    def debug_get(self, name, n):
        out = np.zeros(int(n))
        self._ck(_cuda.admm_b200_debug_get(self.h, name.encode(), _dp(out), ctypes.c_longlong(out.size)))
        return out

    def time_kernels(self, reps=10):
        out = np.zeros(3)
        self._ck(_cuda.admm_b200_time_kernels(self.h, int(reps), _dp(out)))
        return {"local_ms": out[0], "assemble_ms": out[1], "global_ms": out[2]}

    def set_deferred_timers(self, on=True, stride=1):
        """Timed steps without a synchronise per step (admm_b200_set_deferred_timers); read with collect_timers().
        stride = n > 1: only every n-th step records its events (and is counted by collect_timers)."""
        self._ck(_cuda.admm_b200_set_deferred_timers(self.h, int(max(1, stride)) if on else 0))

    def collect_timers(self):
        """Sums over all steps since the last collection: RuntimeData fields + 'steps' (admm_b200_collect_timers)."""
        class _RT(ctypes.Structure):
            _fields_ = [("global_ms", ctypes.c_double), ("local_ms", ctypes.c_double), ("collision_ms", ctypes.c_double),
                        ("inner_iters", ctypes.c_int), ("assemble_ms", ctypes.c_double), ("step_ms", ctypes.c_double)]
        rt, steps = _RT(), ctypes.c_int(0)
        self._ck(_cuda.admm_b200_collect_timers(self.h, ctypes.byref(rt), ctypes.byref(steps)))
        return {"global_ms": rt.global_ms, "local_ms": rt.local_ms, "collision_ms": rt.collision_ms, "inner_iters": rt.inner_iters,
                "assemble_ms": rt.assemble_ms, "step_ms": rt.step_ms, "steps": steps.value}

    def kernel_times(self):
        """Kernel-only times of the last timed step: {name: (summed ms, launches)} (admm_b200_kernel_times)."""
        ms, n = np.zeros(3), np.zeros(3, dtype=np.int64)
        self._ck(_cuda.admm_b200_kernel_times(self.h, _dp(ms), n.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong))))
        return {"tet_local_kernel": (ms[0], int(n[0])), "assemble_kernel": (ms[1], int(n[1])), "solve_kernel": (ms[2], int(n[2]))}

    def info(self):
        return _cuda.admm_b200_solver_info(self.h).decode()

    def launch_count(self):
        return int(_cuda.admm_b200_launch_count(self.h))


class Solver(object):
    """admm_b200::Solver (host/admm_b200.hpp), the mirror of admm::Solver (src/Solver.hpp:33-141)."""

    def __init__(self):
        self.h = ctypes.c_void_p(_host.admmhost_create())
        self._opts = dict(device=0, precision=FP32, gs_max_iters=30, gs_tol=1e-10, gs_omega=1.9,
                          coloring=COLOR_GREEDY, keep_z=False, timers=True, stream=None, gs_parts=0)
        self._user_colors = None

    def close(self):
        if self.h:
            _host.admmhost_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc == 2:
            return False
        if rc:
            raise AdmmError(_host.admmhost_last_error(self.h).decode())
        return True

    # --- scene -------------------------------------------------------------------------------
    def add_nodes(self, x, m):
        x, m = _f64(x).ravel(), _f64(m).ravel()
        if m.size * 3 == x.size:
            m = np.repeat(m, 3)
        return _host.admmhost_add_nodes(self.h, _dp(x), _dp(m), x.size // 3)

    def add_tets(self, verts, inds, model, mu, lam, kappa=0.0, vertex_offset=0):
        verts, inds = _f64(verts).ravel(), _i32(inds).ravel()
        self._ck(_host.admmhost_add_tets(self.h, _dp(verts), _ip(inds), inds.size // 4, int(model),
                                         ctypes.c_double(mu), ctypes.c_double(lam), ctypes.c_double(kappa), int(vertex_offset)))

    def add_spline_tets(self, verts, inds, spline_type, mu, lam, spline, vertex_offset=0):
        """SplineTet(tet, verts, Lame(mu, lam), spline) with spline = (mu, lambda, kappa) of its own
        (src/TetEnergyTerm.hpp:200-205); spline_type 0 NeoHookean, 1 StVK, 2 CoRotated."""
        verts, inds = _f64(verts).ravel(), _i32(inds).ravel()
        self._ck(_host.admmhost_add_spline_tets(self.h, _dp(verts), _ip(inds), inds.size // 4, int(spline_type), ctypes.c_double(mu), ctypes.c_double(lam),
                                                ctypes.c_double(spline[0]), ctypes.c_double(spline[1]), ctypes.c_double(spline[2]), int(vertex_offset)))

    def add_tris(self, verts, inds, mu, lam, limit_min=-100.0, limit_max=100.0, vertex_offset=0):
        verts, inds = _f64(verts).ravel(), _i32(inds).ravel()
        self._ck(_host.admmhost_add_tris(self.h, _dp(verts), _ip(inds), inds.size // 3, ctypes.c_double(mu),
                                         ctypes.c_double(lam), ctypes.c_double(limit_min), ctypes.c_double(limit_max), int(vertex_offset)))

    def set_pins(self, inds, points=None):
        inds = _i32(inds).ravel()
        pts = _f64(points).ravel() if points is not None else None
        self._ck(_host.admmhost_set_pins(self.h, _ip(inds), _dp(pts), inds.size))

    def add_floor(self, y):
        self._ck(_host.admmhost_add_floor(self.h, ctypes.c_double(y)))

    def add_sphere(self, center, radius):
        c = _f64(center)
        self._ck(_host.admmhost_add_sphere(self.h, _dp(c), ctypes.c_double(radius)))

    def add_wind(self, tris, direction):
        """Solver::ext_forces.push_back(WindForce(tris)) with WindForce::direction (src/ExplicitForce.hpp:40-48); returns the
        index to pass to set_wind_direction.  Call before initialize."""
        tris, d = _i32(tris).ravel(), _f64(direction).ravel()
        return int(_host.admmhost_add_wind(self.h, _ip(tris), tris.size // 3, _dp(d)))

    def set_wind_direction(self, index, direction):
        d = _f64(direction).ravel()
        self._ck(_host.admmhost_set_wind_direction(self.h, int(index), _dp(d)))

    def set_surface_inds(self, inds):
        """Solver::surface_inds: the vertices UzawaCG's collision detection tests, in that order (empty: all nodes)."""
        inds = _i32(inds).ravel()
        _host.admmhost_set_surface_inds(self.h, _ip(inds), inds.size)

    def set_options(self, **kw):
        for k in kw:
            if k not in self._opts:
                raise KeyError(k)
        self._opts.update(kw)

    # --- multi-GPU (one process per GPU) -----------------------------------------------------------
    def set_rank(self, rank, world):
        """Call before initialize(): this process is `rank` of `world` (<= 8) on one NVLink box."""
        _host.admmhost_set_rank(self.h, int(rank), int(world))
        self._rank, self._world = int(rank), int(world)

    def mgpu_connect(self, all_gather_bytes):
        """After initialize() on every rank: swaps the CUDA-IPC blobs.  `all_gather_bytes(b) -> [b_0..b_{w-1}]`
        is the caller's collective (e.g. built on torch.distributed.all_gather_object)."""
        blob = ctypes.create_string_buffer(IPC_BYTES)
        self._ck(_host.admmhost_mgpu_export(self.h, blob))
        blobs = all_gather_bytes(bytes(blob.raw))
        for r, b in enumerate(blobs):
            if r != self._rank:
                self._ck(_host.admmhost_mgpu_import(self.h, r, ctypes.create_string_buffer(b, IPC_BYTES)))
        self._ck(_host.admmhost_mgpu_ready(self.h))

    def mgpu_nodes(self):
        """(owned, ghost) node counts of this rank: step() moves owned + ghost nodes up and down."""
        out = (ctypes.c_int * 2)()
        self._ck(_host.admmhost_mgpu_nodes(self.h, out))
        return int(out[0]), int(out[1])

    def node_owner(self):
        out = np.zeros(self.dof // 3, dtype=np.int32)
        n = _host.admmhost_get_node_owner(self.h, _ip(out))
        return out[:n]

    def set_colors(self, colors):
        """colors: list of node lists (colour -> nodes), e.g. read from the reference."""
        self._user_colors = [np.asarray(c, dtype=np.int32) for c in colors]

    def _apply_options(self):
        o = self._opts
        _host.admmhost_set_options(self.h, int(o["device"]), int(o["precision"]), int(o["gs_max_iters"]),
                                   ctypes.c_double(o["gs_tol"]), ctypes.c_double(o["gs_omega"]), int(o["coloring"]),
                                   int(bool(o["keep_z"])), int(bool(o["timers"])), ctypes.c_void_p(o["stream"] or 0))
        _host.admmhost_set_gs_parts(self.h, int(o["gs_parts"]))

    def set_timers(self, on):
        """DeviceOptions::timers after initialize(): whether step() / step_device() fill RuntimeData (and synchronise)."""
        self._opts["timers"] = bool(on)
        self._apply_options()

    def initialize(self, dt=1.0 / 24.0, admm_iters=10, gravity=-9.8, linsolver=0, constraint_w=-1.0):
        self._apply_options()
        if self._user_colors is not None:
            off = np.zeros(len(self._user_colors) + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(c) for c in self._user_colors])
            nodes = np.concatenate(self._user_colors).astype(np.int32) if len(self._user_colors) else np.zeros(0, np.int32)
            _host.admmhost_set_colors(self.h, len(self._user_colors), _ip(off), _ip(nodes))
        return self._ck(_host.admmhost_initialize(self.h, ctypes.c_double(dt), int(admm_iters), ctypes.c_double(gravity),
                                                  int(linsolver), ctypes.c_double(constraint_w)))

    def set_admm_iters(self, it):
        _host.admmhost_set_admm_iters(self.h, int(it))

    # --- stepping ----------------------------------------------------------------------------
    def step(self):
        self._ck(_host.admmhost_step(self.h))

    def step_device(self):
        self._ck(_host.admmhost_step_device(self.h))

    def upload_state(self):
        """Host m_x / m_v -> device (step_device() does not look at the host copies)."""
        self._ck(_host.admmhost_upload_state(self.h))

    def sync_state(self):
        self._ck(_host.admmhost_sync_state(self.h))

    # --- state -------------------------------------------------------------------------------
    @property
    def dof(self):
        return _host.admmhost_dof(self.h)

    def get_x(self):
        out = np.empty(self.dof)
        _host.admmhost_get_x(self.h, _dp(out))
        return out

    def get_v(self):
        out = np.empty(self.dof)
        _host.admmhost_get_v(self.h, _dp(out))
        return out

    def set_x(self, x):
        x = _f64(x).ravel()
        assert x.size == self.dof
        _host.admmhost_set_x(self.h, _dp(x))

    def set_v(self, v):
        v = _f64(v).ravel()
        assert v.size == self.dof
        _host.admmhost_set_v(self.h, _dp(v))

    def runtime_data(self):
        out = np.zeros(6)
        _host.admmhost_runtime(self.h, _dp(out))
        return {"global_ms": out[0], "local_ms": out[1], "collision_ms": out[2], "inner_iters": int(out[3]),
                "assemble_ms": out[4], "step_ms": out[5]}

    def n_rows(self):
        return _host.admmhost_n_rows(self.h)

    def row_offsets(self):
        out = np.zeros(_host.admmhost_n_terms(self.h), dtype=np.int32)
        _host.admmhost_get_row_offsets(self.h, _ip(out))
        return out

    def tet_rest_data(self):
        """(idx [n,4], Dm^-1 [n,9], weight [n], g_index [n]) of the tet terms as handed to the device."""
        n = _host.admmhost_get_tet_rest(self.h, None, None, None, None)
        idx, dminv, w, row = np.zeros((n, 4), np.int32), np.zeros((n, 9)), np.zeros(n), np.zeros(n, np.int32)
        _host.admmhost_get_tet_rest(self.h, _ip(idx), _dp(dminv), _dp(w), _ip(row))
        return idx, dminv, w, row

    def tri_rest_data(self):
        n = _host.admmhost_get_tri_rest(self.h, None, None, None, None)
        idx, rest, w, row = np.zeros((n, 3), np.int32), np.zeros((n, 4)), np.zeros(n), np.zeros(n, np.int32)
        _host.admmhost_get_tri_rest(self.h, _ip(idx), _dp(rest), _dp(w), _ip(row))
        return idx, rest, w, row

    def system_matrix(self):
        shape = (ctypes.c_longlong * 2)()
        _host.admmhost_system_shape(self.h, shape)
        n, nnz = int(shape[0]), int(shape[1])
        rowptr, cols, vals = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
        _host.admmhost_system_get(self.h, _ip(rowptr), _ip(cols), _dp(vals))
        return rowptr, cols, vals

    def colors(self):
        nc = _host.admmhost_n_colors(self.h)
        off, nodes = np.zeros(nc + 1, np.int32), np.zeros(self.dof // 3, np.int32)
        _host.admmhost_get_colors(self.h, _ip(off), _ip(nodes))
        return [nodes[off[i]:off[i + 1]].copy() for i in range(nc)]

    def device(self):
        """The C-ABI handle behind this solver (valid after initialize)."""
        h = _host.admmhost_device_handle(self.h)
        if not h:
            raise AdmmError("solver is not initialized")
        return DeviceSolver(handle=h)


def color_matrix(rowptr, cols, vals, method=COLOR_GREEDY):
    """Host colouring alone (no device)."""
    rowptr, cols, vals = _i32(rowptr), _i32(cols), _f64(vals)
    n = rowptr.size - 1
    nc = ctypes.c_int(0)
    off, nodes = np.zeros(n + 2, np.int32), np.zeros(n, np.int32)
    rc = _host.admmhost_color_matrix(n, _ip(rowptr), _ip(cols), _dp(vals), int(method), ctypes.byref(nc), _ip(off), _ip(nodes))
    if rc:
        raise AdmmError("colouring failed (%d)" % rc)
    return [nodes[off[i]:off[i + 1]].copy() for i in range(nc.value)]


def ldlt_solve_host(rowptr, cols, vals, pos, b):
    """Host factorisation check (no device): solves A x = b with the nested-dissection LDL^T."""
    rowptr, cols, vals, pos, b = _i32(rowptr), _i32(cols), _f64(vals), _f64(pos).ravel(), _f64(b)
    x = np.zeros_like(b)
    stats = (ctypes.c_longlong * 4)()
    rc = _host.admmhost_ldlt_check(rowptr.size - 1, _ip(rowptr), _ip(cols), _dp(vals), _dp(pos), _dp(b), _dp(x), stats)
    if rc:
        raise AdmmError("ldlt failed")
    return x, int(stats[0])


def ldlt_blocks_solve_host(rowptr, cols, vals, pos, b):
    """Host check of the device's block (supernodal) solve plan (no device): factor with nested dissection, then walk
    the block plan.  Returns (x, stats) with stats = {blocks, largest, levels_f, levels_b, nnz_out, nnz_inv, cut, segments}."""
    rowptr, cols, vals, pos, b = _i32(rowptr), _i32(cols), _f64(vals), _f64(pos).ravel(), _f64(b)
    x = np.zeros_like(b)
    stats = (ctypes.c_longlong * 8)()
    rc = _host.admmhost_ldlt_blocks_check(rowptr.size - 1, _ip(rowptr), _ip(cols), _dp(vals), _dp(pos), _dp(b), _dp(x), stats)
    if rc:
        raise AdmmError("ldlt block plan failed")
    return x, dict(zip(("blocks", "largest", "levels_f", "levels_b", "nnz_out", "nnz_inv", "cut", "segments"), [int(v) for v in stats]))


def wind_project(tris, direction, dt, x, v):
    """WindForce::project alone (host form of the device kernels): returns the new velocities."""
    tris, d, x = _i32(tris).ravel(), _f64(direction).ravel(), _f64(x).ravel()
    v = _f64(v).ravel().copy()
    _host.admmhost_wind_project(_ip(tris), tris.size // 3, _dp(d), ctypes.c_double(dt), x.size // 3, _dp(x), _dp(v))
    return v
