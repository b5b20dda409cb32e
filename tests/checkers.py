"""ctypes views of the two CPU checkers (TEST INFRASTRUCTURE ONLY):

  oracle/liboracle.so          plain-C restatement of the reference algorithm
  oracle/_ref/libadmm_ref.so   the unmodified reference compiled here (may be absent)

Both are wrapped in the same Python class so a test can run either through identical calls.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libadmm_ref.so")
BINDING_PATH = os.path.join(ROOT, "oracle", "_ref", "libadmm_gpubinding.so")

_dpt = ctypes.POINTER(ctypes.c_double)
_ipt = ctypes.POINTER(ctypes.c_int)
D = ctypes.c_double


def dp(a):
    return a.ctypes.data_as(_dpt) if a is not None else None


def ip(a):
    return a.ctypes.data_as(_ipt) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        L = ctypes.CDLL(ORACLE_PATH)
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_last_error.restype = ctypes.c_char_p
        L.oracle_last_error.argtypes = [ctypes.c_void_p]
        L.oracle_term_energy.restype = ctypes.c_double
        for fn in ("oracle_destroy", "oracle_add_nodes", "oracle_add_tets", "oracle_add_tris", "oracle_set_pins", "oracle_add_obstacle",
                   "oracle_set_colors", "oracle_gs_params", "oracle_initialize", "oracle_step", "oracle_step_traced", "oracle_dof",
                   "oracle_n_rows", "oracle_n_terms", "oracle_get_x", "oracle_get_v", "oracle_set_x", "oracle_set_v", "oracle_set_admm_iters",
                   "oracle_runtime", "oracle_get_row_offsets", "oracle_get_weights", "oracle_A_shape", "oracle_A_get", "oracle_linsolve",
                   "oracle_apply_D", "oracle_term_energy", "oracle_set_uzawa", "oracle_add_spline_tets", "oracle_prox_tets_k", "oracle_add_wind", "oracle_wind_mode", "oracle_wind_project"):
            getattr(L, fn).argtypes = None
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_PATH)


def ref_lib():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(REF_PATH)
        L.ref_create.restype = ctypes.c_void_p
        L.ref_last_error.restype = ctypes.c_char_p
        L.ref_last_error.argtypes = [ctypes.c_void_p]
        L.ref_tet_energy.restype = ctypes.c_double
        _ref = L
    return _ref


class CheckerError(RuntimeError):
    pass


_binding = None


def have_binding():
    return os.path.exists(BINDING_PATH)


def binding_lib():
    """oracle/_ref/libadmm_gpubinding.so: the reference-side binding (integration/GpuSolver.hpp, admm::GpuSolver :
    admm::Solver) compiled against the reference's own headers (oracle/gpu_binding.cpp)."""
    global _binding
    if _binding is None:
        L = ctypes.CDLL(BINDING_PATH)
        L.gpub_create.restype = ctypes.c_void_p
        L.gpub_last_error.restype = ctypes.c_char_p
        L.gpub_last_error.argtypes = [ctypes.c_void_p]
        L.gpub_solver_info.restype = ctypes.c_char_p
        L.gpub_solver_info.argtypes = [ctypes.c_void_p]
        _binding = L
    return _binding


class GpuBinding(object):
    """admm::GpuSolver through oracle/gpu_binding.cpp: step() runs on the GPU, cpu_step() is the reference's own
    Solver::step() on the same object (same D, W, A, colours, pins)."""

    def __init__(self, precision=0, gs_parts=0, keep_z=False):
        self.L = binding_lib()
        self.h = ctypes.c_void_p(self.L.gpub_create())
        self.L.gpub_set_options(self.h, int(precision), int(gs_parts), int(bool(keep_z)))

    def close(self):
        if self.h:
            self.L.gpub_destroy(self.h)
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
            raise CheckerError(self.L.gpub_last_error(self.h).decode())
        return True

    def add_nodes(self, x, m):
        x, m = f64(x).ravel(), f64(m).ravel()
        if m.size * 3 == x.size:
            m = np.repeat(m, 3)
        return self.L.gpub_add_nodes(self.h, dp(x), dp(m), x.size // 3)

    def add_tets(self, verts, inds, model, mu, lam, kappa=0.0, vertex_offset=0, spline=None):
        verts, inds = f64(verts).ravel(), i32(inds).ravel()
        sp = spline if spline is not None else (mu, lam, kappa)
        self._ck(self.L.gpub_add_tets(self.h, dp(verts), ip(inds), inds.size // 4, int(model), D(mu), D(lam), D(sp[0]), D(sp[1]), D(sp[2]), int(vertex_offset)))

    def add_tris(self, verts, inds, mu, lam, limit_min=-100.0, limit_max=100.0, vertex_offset=0):
        verts, inds = f64(verts).ravel(), i32(inds).ravel()
        self._ck(self.L.gpub_add_tris(self.h, dp(verts), ip(inds), inds.size // 3, D(mu), D(lam), D(limit_min), D(limit_max), int(vertex_offset)))

    def set_pins(self, inds, points=None):
        inds = i32(inds).ravel()
        pts = f64(points).ravel() if points is not None else None
        self._ck(self.L.gpub_set_pins(self.h, ip(inds), dp(pts), inds.size))

    def add_floor(self, y):
        self._ck(self.L.gpub_add_floor(self.h, D(y)))

    def add_wind(self, tris, direction):
        tris, d = i32(tris).ravel(), f64(direction).ravel()
        self._ck(self.L.gpub_add_wind(self.h, ip(tris), tris.size // 3, dp(d)))

    def add_sphere(self, c, r):
        cc = f64(c)
        self._ck(self.L.gpub_add_sphere(self.h, dp(cc), D(r)))

    def initialize(self, dt=1.0 / 24.0, admm_iters=10, gravity=-9.8, linsolver=0):
        return self._ck(self.L.gpub_initialize(self.h, D(dt), int(admm_iters), D(gravity), int(linsolver)))

    def step(self):
        self._ck(self.L.gpub_step(self.h))

    def cpu_step(self):
        self._ck(self.L.gpub_cpu_step(self.h))

    @property
    def dof(self):
        return self.L.gpub_dof(self.h)

    def get_x(self):
        out = np.empty(self.dof)
        self.L.gpub_get_x(self.h, dp(out))
        return out

    def get_v(self):
        out = np.empty(self.dof)
        self.L.gpub_get_v(self.h, dp(out))
        return out

    def set_x(self, x):
        x = f64(x).ravel()
        self.L.gpub_set_x(self.h, dp(x))

    def set_v(self, v):
        v = f64(v).ravel()
        self.L.gpub_set_v(self.h, dp(v))

    def runtime_data(self):
        out = np.zeros(4)
        self.L.gpub_runtime(self.h, dp(out))
        return {"global_ms": out[0], "local_ms": out[1], "collision_ms": out[2], "inner_iters": int(out[3])}

    def info(self):
        return self.L.gpub_solver_info(self.h).decode()

    def tets(self):
        n = self.L.gpub_n_tets(self.h)
        idx, dminv, w, row, model = np.zeros((n, 4), np.int32), np.zeros((n, 9)), np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.gpub_get_tets(self.h, ip(idx), dp(dminv), dp(w), ip(row), ip(model))
        return idx, dminv, w, row, model

    def tris(self):
        n = self.L.gpub_n_tris(self.h)
        idx, rest, w, row = np.zeros((n, 3), np.int32), np.zeros((n, 4)), np.zeros(n), np.zeros(n, np.int32)
        self.L.gpub_get_tris(self.h, ip(idx), dp(rest), dp(w), ip(row))
        return idx, rest, w, row

    def pins(self):
        n = self.L.gpub_n_pins(self.h)
        idx, row, w = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        self.L.gpub_get_pins(self.h, ip(idx), ip(row), dp(w))
        return idx, row, w

    def colors(self):
        nc = self.L.gpub_n_colors(self.h)
        off, nodes = np.zeros(nc + 1, np.int32), np.zeros(self.dof // 3, np.int32)
        self.L.gpub_get_colors(self.h, ip(off), ip(nodes))
        return [nodes[off[i]:off[i + 1]].copy() for i in range(nc)]

    def debug_get(self, name, n):
        out = np.zeros(int(n))
        if self.L.gpub_debug_get(self.h, name.encode(), dp(out), ctypes.c_longlong(out.size)):
            raise CheckerError(self.L.gpub_last_error(self.h).decode())
        return out


class CpuSolver(object):
    """kind = 'oracle' | 'ref'"""

    def __init__(self, kind):
        self.kind = kind
        if kind == "oracle":
            self.L, self.p = oracle_lib(), "oracle_"
        else:
            self.L, self.p = ref_lib(), "ref_"
        self.h = ctypes.c_void_p(self._f("create")())
        self.linsolver = 0
        self._initialized = False
        self._colors = None

    def _f(self, name):
        return getattr(self.L, self.p + name)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
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
            raise CheckerError(self._f("last_error")(self.h).decode())
        return True

    def add_nodes(self, x, m):
        x, m = f64(x).ravel(), f64(m).ravel()
        if m.size * 3 == x.size:
            m = np.repeat(m, 3)
        return self._f("add_nodes")(self.h, dp(x), dp(m), x.size // 3)

    def add_tets(self, verts, inds, model, mu, lam, kappa=0.0, vertex_offset=0):
        verts, inds = f64(verts).ravel(), i32(inds).ravel()
        self._ck(self._f("add_tets")(self.h, dp(verts), ip(inds), inds.size // 4, int(model), D(mu), D(lam), D(kappa), int(vertex_offset)))

    def add_spline_tets(self, verts, inds, spline_type, mu, lam, spline, vertex_offset=0):
        """SplineTet(tet, verts, Lame(mu, lam), spline = (mu, lambda, kappa) of its own); type 0 NeoHookean, 1 StVK, 2 CoRotated."""
        verts, inds = f64(verts).ravel(), i32(inds).ravel()
        if self.kind == "oracle":
            self._ck(self.L.oracle_add_spline_tets(self.h, dp(verts), ip(inds), inds.size // 4, 3 + int(spline_type), D(mu), D(lam), D(spline[0]), D(spline[1]), D(spline[2]), int(vertex_offset)))
        else:
            self._ck(self.L.ref_add_spline_tets(self.h, dp(verts), ip(inds), inds.size // 4, int(spline_type), D(mu), D(lam), D(spline[0]), D(spline[1]), D(spline[2]), int(vertex_offset)))

    def add_tris(self, verts, inds, mu, lam, limit_min=-100.0, limit_max=100.0, vertex_offset=0):
        verts, inds = f64(verts).ravel(), i32(inds).ravel()
        self._ck(self._f("add_tris")(self.h, dp(verts), ip(inds), inds.size // 3, D(mu), D(lam), D(limit_min), D(limit_max), int(vertex_offset)))

    def set_pins(self, inds, points=None):
        inds = i32(inds).ravel()
        pts = f64(points).ravel() if points is not None else None
        self._ck(self._f("set_pins")(self.h, ip(inds), dp(pts), inds.size))

    def add_floor(self, y):
        if self.kind == "oracle":
            p = f64([y, 0, 0, 0])
            self._ck(self.L.oracle_add_obstacle(self.h, 0, dp(p)))
        else:
            self._ck(self.L.ref_add_floor(self.h, D(y)))

    def add_sphere(self, c, r):
        if self.kind == "oracle":
            p = f64([c[0], c[1], c[2], r])
            self._ck(self.L.oracle_add_obstacle(self.h, 1, dp(p)))
        else:
            cc = f64(c)
            self._ck(self.L.ref_add_sphere(self.h, dp(cc), D(r)))

    def add_wind(self, tris, direction, sequential=False):
        """Solver::ext_forces.push_back(WindForce(tris)) with WindForce::direction.  The oracle forms every force from the
        velocities before the call unless sequential=True (the reference's one-thread order, see oracle_wind_project)."""
        tris, d = i32(tris).ravel(), f64(direction).ravel()
        if self.kind == "oracle":
            self.L.oracle_wind_mode(self.h, int(bool(sequential)))
            self._ck(self.L.oracle_add_wind(self.h, ip(tris), tris.size // 3, dp(d)))
        else:
            self._ck(self.L.ref_add_wind(self.h, ip(tris), tris.size // 3, dp(d)))

    def set_surface_inds(self, inds, constraint_w=-1.0):
        """Solver::surface_inds (+ Settings::constraint_w for the oracle, which takes both through one call; the reference
        gets constraint_w through initialize).  Call before initialize."""
        inds = i32(inds).ravel()
        self._surf = inds
        if self.kind == "oracle":
            self.L.oracle_set_uzawa(self.h, inds.size, ip(inds), D(constraint_w))
        else:
            self.L.ref_set_surface_inds(self.h, ip(inds), inds.size)

    def set_colors(self, colors):
        off = np.zeros(len(colors) + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(c) for c in colors])
        nodes = np.concatenate([np.asarray(c, dtype=np.int32) for c in colors]).astype(np.int32)
        self._colors = (off, nodes)
        if self.kind == "ref" and not self._initialized:
            return  # NodalMultiColorGS exists only after initialize(): applied there
        self._f("set_colors")(self.h, len(colors), ip(off), ip(nodes))

    def get_colors(self):
        assert self.kind == "ref"
        nc = self.L.ref_n_colors(self.h)
        off, nodes = np.zeros(nc + 1, np.int32), np.zeros(self.dof // 3, np.int32)
        self.L.ref_get_colors(self.h, ip(off), ip(nodes))
        return [nodes[off[i]:off[i + 1]].copy() for i in range(nc)]

    def gs_params(self, max_iters=30, tol=1e-10, omega=1.9):
        self._f("gs_params")(self.h, int(max_iters), D(tol), D(omega))

    def initialize(self, dt=1.0 / 24.0, admm_iters=10, gravity=-9.8, linsolver=0, constraint_w=-1.0):
        self.linsolver = linsolver
        if self.kind == "oracle":
            return self._ck(self.L.oracle_initialize(self.h, D(dt), int(admm_iters), D(gravity), int(linsolver)))
        ok = self._ck(self.L.ref_initialize(self.h, D(dt), int(admm_iters), D(gravity), int(linsolver), D(constraint_w)))
        self._initialized = bool(ok)
        if ok and linsolver == 1 and self._colors is not None:
            off, nodes = self._colors   # replace the reference's own (randomised) colour lists by the caller's
            self.L.ref_set_colors(self.h, len(off) - 1, ip(off), ip(nodes))
        return ok

    def step(self):
        self._ck(self._f("step")(self.h))

    def traced_step(self, admm_iters):
        R, dof = self.n_rows, self.dof
        z, u = np.zeros((admm_iters, R)), np.zeros((admm_iters, R))
        b, x = np.zeros((admm_iters, dof)), np.zeros((admm_iters, dof))
        name = "step_traced" if self.kind == "oracle" else "traced_step"
        self._ck(self._f(name)(self.h, dp(z), dp(u), dp(b), dp(x)))
        return z, u, b, x

    @property
    def dof(self):
        return self._f("dof")(self.h)

    @property
    def n_rows(self):
        return self.L.oracle_n_rows(self.h) if self.kind == "oracle" else self.L.ref_n_weights(self.h)

    def get_x(self):
        out = np.empty(self.dof)
        self._f("get_x")(self.h, dp(out))
        return out

    def get_v(self):
        out = np.empty(self.dof)
        self._f("get_v")(self.h, dp(out))
        return out

    def set_x(self, x):
        x = f64(x).ravel()
        self._f("set_x")(self.h, dp(x))

    def set_v(self, v):
        v = f64(v).ravel()
        self._f("set_v")(self.h, dp(v))

    def runtime_data(self):
        out = np.zeros(4)
        self._f("runtime")(self.h, dp(out))
        return {"global_ms": out[0], "local_ms": out[1], "collision_ms": out[2], "inner_iters": int(out[3])}

    def matrix_A(self):
        """3n x 3n system matrix as scipy CSR."""
        import scipy.sparse as sp
        shape = (ctypes.c_longlong * 3)()
        if self.kind == "oracle":
            self.L.oracle_A_shape(self.h, shape)
            n, nnz = int(shape[0]), int(shape[1])
            rp, ci, va = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
            self.L.oracle_A_get(self.h, ip(rp), ip(ci), dp(va))
        else:
            self.L.ref_sparse_shape(self.h, 1, shape)
            n, nnz = int(shape[0]), int(shape[2])
            rp, ci, va = np.zeros(n + 1, np.int32), np.zeros(nnz, np.int32), np.zeros(nnz)
            self.L.ref_sparse_get(self.h, 1, ip(rp), ip(ci), dp(va))
        return sp.csr_matrix((va, ci, rp), shape=(n, n))

    def linsolve(self, x, b):
        x = f64(x).ravel().copy()
        b = f64(b).ravel()
        it = self._f("linsolve")(self.h, dp(x), dp(b))
        return x, it


def prox_tets(kind, model, mu, lam, z, kappa=0.0):
    z = f64(z).reshape(-1, 9)
    out = np.empty_like(z)
    if kind == "oracle":
        rc = oracle_lib().oracle_prox_tets(int(model), D(mu), D(lam), D(kappa), z.shape[0], dp(z), dp(out))
    else:
        rc = ref_lib().ref_prox_tets(int(model), D(mu), D(lam), D(kappa), z.shape[0], dp(z), dp(out))
    return out, rc


def prox_spline_tets(kind, spline_type, mu, lam, spline, z):
    """SplineTet::prox of an element with Lame (mu, lam) and a spline (mu, lambda, kappa) of its own."""
    z = f64(z).reshape(-1, 9)
    out = np.empty_like(z)
    if kind == "oracle":
        K = lam + (2.0 / 3.0) * mu
        rc = oracle_lib().oracle_prox_tets_k(3 + int(spline_type), D(spline[0]), D(spline[1]), D(spline[2]), D(K), z.shape[0], dp(z), dp(out))
    else:
        rc = ref_lib().ref_prox_spline_tets(int(spline_type), D(mu), D(lam), D(spline[0]), D(spline[1]), D(spline[2]), z.shape[0], dp(z), dp(out))
    return out, rc


def prox_tris(kind, mu, lam, z, limit_min=-100.0, limit_max=100.0):
    z = f64(z).reshape(-1, 6)
    out = np.empty_like(z)
    if kind == "oracle":
        oracle_lib().oracle_prox_tris(D(limit_min), D(limit_max), z.shape[0], dp(z), dp(out))
    else:
        ref_lib().ref_prox_tris(D(mu), D(lam), D(limit_min), D(limit_max), z.shape[0], dp(z), dp(out))
    return out


def random_F(n, sigma, seed=1234, rotate=True):
    """F = R (I + N(0, sigma^2)) with random rotations R, column-major 9-vectors."""
    rng = np.random.RandomState(seed)
    F = np.eye(3)[None, :, :] + sigma * rng.randn(n, 3, 3)
    if rotate:
        Q, _ = np.linalg.qr(rng.randn(n, 3, 3))
        det = np.linalg.det(Q)
        Q[:, :, 2] *= det[:, None]
        F = Q @ F
    return np.ascontiguousarray(F.transpose(0, 2, 1).reshape(n, 9))


def wind_project(kind, tris, direction, dt, x, v, sequential=False):
    """WindForce::project (src/ExplicitForce.cpp:47-104) alone: returns the new velocities.  kind = "oracle" | "ref"."""
    tris, d, x = i32(tris).ravel(), f64(direction).ravel(), f64(x).ravel()
    v = f64(v).ravel().copy()
    if kind == "oracle":
        oracle_lib().oracle_wind_project(tris.size // 3, ip(tris), dp(d), D(dt), x.size // 3, dp(x), dp(v), int(bool(sequential)))
    else:
        ref_lib().ref_wind_project(ip(tris), tris.size // 3, dp(d), D(dt), x.size // 3, dp(x), dp(v))
    return v
