"""Solver::ext_forces / WindForce (src/Solver.cpp:53-54, src/ExplicitForce.cpp:47-104): the oracle's two readings pinned
against the compiled reference on the CPU, the host mirror, the device kernels and the reference-side binding."""
import json
import os

import numpy as np
import pytest

import checkers
import scenes
from checkers import CpuSolver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def record(name, **kw):
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **{k: float(v) for k, v in kw.items()})) + "\n")
    except OSError:
        pass


MU, LAM = scenes.lame(*scenes.LAME_SOFT)
WIND = [1.5, -0.3, 2.0]


def _fine_cloth(pkg, n=24):
    """A cloth fine enough for the explicit drag not to be stiff (alpha_n area |v| dt << 1), moving."""
    v, t, m, pins = scenes.cloth(pkg.meshes, n)
    rng = np.random.RandomState(5)
    x = v + 0.01 * rng.randn(*v.shape)
    vel = 0.5 * rng.randn(*v.shape)
    return v, t, m, pins, x, vel


need_ref = pytest.mark.skipif(not checkers.have_ref(), reason="oracle/_ref not built (no /root/reference)")


@need_ref
def test_wind_project_sequential_reading_is_the_reference_with_one_thread(pkg, cpu):
    v, t, m, pins, x, vel = _fine_cloth(pkg, 8)   # coarse: kicks of several m/s, the two readings differ visibly
    vel = 4.0 * vel
    checkers.ref_lib().ref_set_omp_threads(1)
    try:
        vr = checkers.wind_project("ref", t, WIND, 1.0 / 24, x, vel)
    finally:
        checkers.ref_lib().ref_set_omp_threads(os.cpu_count() or 1)
    vs = checkers.wind_project("oracle", t, WIND, 1.0 / 24, x, vel, sequential=True)
    assert np.abs(vr - vel.ravel()).max() > 0.05        # the force did something
    assert np.array_equal(vr, vs)                        # same arithmetic in the same order: bit for bit


def test_wind_project_mirror_equals_order_independent_oracle(pkg, cpu):
    v, t, m, pins, x, vel = _fine_cloth(pkg)
    vo = checkers.wind_project("oracle", t, WIND, 1.0 / 24, x, vel, sequential=False)
    vm = pkg.wind_project(t, WIND, 1.0 / 24, x, vel)
    assert np.abs(vo - vel.ravel()).max() > 1e-3
    assert np.abs(vo - vm).max() < 1e-14
    # the two readings differ at second order in the kick (here 13 % of it; bit-equal only for non-adjacent triangles)
    vs = checkers.wind_project("oracle", t, WIND, 1.0 / 24, x, vel, sequential=True)
    kick = np.abs(vo - vel.ravel()).max()
    assert np.abs(vs - vo).max() < 0.25 * kick


@need_ref
def test_step_with_wind_oracle_vs_reference(pkg, cpu):
    v, t, m, pins, x, vel = _fine_cloth(pkg, 12)
    checkers.ref_lib().ref_set_omp_threads(1)   # the reference's wind force is only defined up to its thread order
    try:
        out = []
        for kind in ("oracle", "ref"):
            s = CpuSolver(kind)
            s.add_nodes(v, m)
            s.add_tris(v, t, MU, LAM)
            s.set_pins(pins)
            s.add_wind(t, WIND, sequential=True)
            assert s.initialize(dt=1.0 / 24, admm_iters=6, gravity=-9.8, linsolver=0)
            s.set_x(x.ravel()); s.set_v(vel.ravel())
            for _ in range(3):
                s.step()
            out.append((s.get_x(), s.get_v()))
    finally:
        checkers.ref_lib().ref_set_omp_threads(os.cpu_count() or 1)
    assert np.abs(out[0][0] - out[1][0]).max() < 2e-7
    assert np.abs(out[0][1] - out[1][1]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("linsolver", [0, 1])
def test_device_wind_vs_oracle(pkg, cpu, linsolver):
    """The device kernels (order-independent reading) inside whole steps, against the oracle in the same reading; the wind
    turns between steps (WindForce::direction is a public member of the reference class)."""
    v, t, m, pins, x, vel = _fine_cloth(pkg, 16)
    for precision, tol in ((1, 1e-7), (0, 1e-4 * float(np.linalg.norm(v.max(0) - v.min(0))))):
        g = pkg.Solver()
        g.set_options(precision=precision)
        g.add_nodes(v, m); g.add_tris(v, t, MU, LAM); g.set_pins(pins)
        wid = g.add_wind(t, WIND)
        assert g.initialize(dt=1.0 / 24, admm_iters=6, gravity=-9.8, linsolver=linsolver)
        o = CpuSolver("oracle")
        o.add_nodes(v, m); o.add_tris(v, t, MU, LAM); o.set_pins(pins)
        o.add_wind(t, WIND, sequential=False)
        if linsolver == 1:
            o.set_colors(g.colors())
        assert o.initialize(dt=1.0 / 24, admm_iters=6, gravity=-9.8, linsolver=linsolver)
        g.set_x(x.ravel()); g.set_v(vel.ravel())
        o.set_x(x.ravel()); o.set_v(vel.ravel())
        for _ in range(3):
            g.step(); o.step()
        err_x, err_v = np.abs(g.get_x() - o.get_x()).max(), np.abs(g.get_v() - o.get_v()).max()
        record("wind_device", linsolver=linsolver, precision=precision, err=err_x, err_v=err_v)
        assert err_x < tol, (precision, err_x)
        # without the wind the cloth ends up somewhere else: the comparison is not vacuous
        g2 = pkg.Solver()
        g2.set_options(precision=precision)
        g2.add_nodes(v, m); g2.add_tris(v, t, MU, LAM); g2.set_pins(pins)
        assert g2.initialize(dt=1.0 / 24, admm_iters=6, gravity=-9.8, linsolver=linsolver)
        g2.set_x(x.ravel()); g2.set_v(vel.ravel())
        for _ in range(3):
            g2.step()
        assert np.abs(g2.get_x() - g.get_x()).max() > 100 * tol
        # resident stepping applies the force too, and the direction can be changed
        g.set_wind_direction(wid, [0.0, 0.0, -3.0])
        g.upload_state()
        g.step_device()
        g.sync_state()
        assert np.isfinite(g.get_x()).all()


@pytest.mark.gpu
@need_ref
def test_binding_applies_ext_forces_like_the_reference(pkg, cpu):
    """admm::GpuSolver::step runs the reference's own ExplicitForce::project on the host arrays before the device step:
    with one OpenMP thread (the only setting in which the reference's wind force is reproducible) the GPU step and the
    reference's Solver::step() of the same object agree."""
    if not checkers.have_binding():
        pytest.skip("oracle/_ref/libadmm_gpubinding.so not built")
    v, t, m, pins, x, vel = _fine_cloth(pkg, 12)
    checkers.ref_lib().ref_set_omp_threads(1)
    try:
        res = []
        for use_gpu in (True, False):
            b = checkers.GpuBinding(precision=1)
            b.add_nodes(v, m); b.add_tris(v, t, MU, LAM); b.set_pins(pins)
            b.add_wind(t, WIND)
            assert b.initialize(admm_iters=6, linsolver=0)
            b.set_x(x.ravel()); b.set_v(vel.ravel())
            for _ in range(3):
                b.step() if use_gpu else b.cpu_step()
            res.append(b.get_x())
            b.close()
    finally:
        checkers.ref_lib().ref_set_omp_threads(os.cpu_count() or 1)
    err = np.abs(res[0] - res[1]).max()
    record("wind_binding", err=err)
    assert err < 1e-6, err
