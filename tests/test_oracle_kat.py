"""Pins the C oracle (and the compiled reference, when present) on the reference's own known-answer
tests: samples/tests/test_lineartet.cpp.  No GPU needed."""
import numpy as np
import pytest

import checkers
from checkers import CpuSolver

KINDS = ["oracle"] + (["ref"] if checkers.have_ref() else [])

# SingleTet::init (test_lineartet.cpp:343-396)
VERTS = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
TET = np.array([[0, 1, 2, 3]], dtype=np.int32)


def lame(E, nu):
    return E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))


def single_tet(kind, mu, lam, model=0, masses=None):
    s = CpuSolver(kind)
    s.add_nodes(VERTS, np.ones(12) if masses is None else masses)
    s.add_tets(VERTS, TET, model, mu, lam)
    return s


def volume(x):
    x = x.reshape(4, 3)
    return np.linalg.det(np.stack([x[1] - x[0], x[2] - x[0], x[3] - x[0]], axis=1)) / 6.0


@pytest.mark.parametrize("kind", KINDS)
def test_reduction_sizes_and_weight(cpu, kind):
    # weights.size()==9, w^2 == K*vol with mu=0, lambda=1 (test_lineartet.cpp:74-78, 376-392)
    s = single_tet(kind, 0.0, 1.0)
    assert s.initialize(dt=1.0 / 24, admm_iters=1, gravity=0.0, linsolver=0)
    assert s.n_rows == 9
    if kind == "oracle":
        w = np.zeros(1)
        s.L.oracle_get_weights(s.h, checkers.dp(w))
        assert abs(w[0] ** 2 - 1.0 * (1.0 / 6.0)) < 1e-12


def test_deformation_gradient_layout(cpu):
    # F of scale(3.1,4.2,5.3) is diag(3.1,4.2,5.3), zero off-diagonal to 1e-12 (test_lineartet.cpp:136-156)
    s = single_tet("oracle", 0.0, 1.0)
    assert s.initialize(dt=1.0 / 24, admm_iters=1, gravity=0.0, linsolver=0)
    x = (VERTS * np.array([3.1, 4.2, 5.3])).ravel()
    out = np.zeros(9)
    s.L.oracle_apply_D(s.h, checkers.dp(x), checkers.dp(out))
    F = out.reshape(3, 3).T  # column-major
    assert np.abs(F - np.diag([3.1, 4.2, 5.3])).max() < 1e-12


def test_energy_known_answers(cpu):
    # energy 0 at rest and under rotation, 0.25 after uniform scale 2, linear in lambda (test_lineartet.cpp:81-118)
    s = single_tet("oracle", 0.0, 1.0)
    assert s.initialize(dt=1.0 / 24, admm_iters=1, gravity=0.0, linsolver=0)
    L = s.L

    def energy(x):
        x = np.ascontiguousarray(x.ravel())
        return L.oracle_term_energy(s.h, 0, checkers.dp(x))
    assert abs(energy(VERTS)) < 1e-12
    ax = np.array([1.0, 1.0, 1.0]) / np.sqrt(3.0)
    a = np.pi / 4
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx
    assert abs(energy(VERTS @ R.T)) < 1e-12
    assert abs(energy(2.0 * VERTS) - 0.25) < 1e-12
    s2 = single_tet("oracle", 0.0, 3.0)
    assert s2.initialize(dt=1.0 / 24, admm_iters=1, gravity=0.0, linsolver=0)
    x2 = np.ascontiguousarray((2.0 * VERTS).ravel())
    assert abs(s2.L.oracle_term_energy(s2.h, 0, checkers.dp(x2)) - 0.75) < 1e-12
    if checkers.have_ref():
        R_ = checkers.ref_lib()
        v = np.ascontiguousarray(VERTS.ravel())
        assert abs(R_.ref_tet_energy(0, checkers.D(0.0), checkers.D(1.0), checkers.dp(v), checkers.dp(x2)) - 0.25) < 1e-12


@pytest.mark.parametrize("kind", KINDS)
def test_solver_iters_known_answer(cpu, kind):
    # x of vertex 3 -> 52.2321 +- 1e-4 for every admm_iters in 21..99, error non-increasing for
    # 5..20 (test_lineartet.cpp:165-230); a subset of iteration counts keeps the suite fast.
    mu, lam = lame(500000, 0.25)
    dt = float(np.float32(1.0) / np.float32(24.0))
    true_x, last_err = 52.2321, -1.0
    for it in list(range(5, 24)) + [30, 50, 99]:
        s = single_tet(kind, mu, lam)
        assert s.initialize(dt=dt, admm_iters=it, gravity=0.0, linsolver=0)
        x = VERTS.ravel().copy()
        x[9:12] = [200, 0, 0]
        s.set_x(x)
        s.step()
        new_x = s.get_x()[9]
        if it > 20:
            assert abs(new_x - true_x) < 1e-4, (it, new_x)
        elif last_err >= 1e-8:
            assert (true_x - new_x) ** 2 <= last_err
        last_err = (true_x - new_x) ** 2
        s.close()


@pytest.mark.parametrize("kind", KINDS)
def test_inversion_known_answer(cpu, kind):
    # mu=lambda=100, dt=0.7, vertex 0 moved to (1,1,1): after 10 steps the volume is back to rest
    # +-1e-6 and the position is independent of admm_iters to 1e-6 (test_lineartet.cpp:236-323)
    target_v = volume(VERTS.ravel())
    last = None
    for it in (10, 11, 25, 60, 99):
        s = single_tet(kind, 100.0, 100.0)
        assert s.initialize(dt=0.7, admm_iters=it, gravity=0.0, linsolver=0)
        x = VERTS.ravel().copy()
        x[0:3] = [1, 1, 1]
        assert volume(x) < 0
        s.set_x(x)
        for _ in range(10):
            s.step()
        xs = s.get_x()
        nv = volume(xs)
        assert nv > 0
        assert abs(nv - target_v) < 1e-6
        if last is not None:
            assert np.linalg.norm(last - xs[0:3]) < 1e-6
        last = xs[0:3].copy()
        s.close()
