"""Pins the C oracle against the UNMODIFIED reference compiled into oracle/_ref (skipped when the
reference could not be built).  No GPU needed."""
import numpy as np
import pytest

import checkers
import scenes
from checkers import CpuSolver

pytestmark = pytest.mark.skipif(not checkers.have_ref(), reason="oracle/_ref not built (no /root/reference)")

MU, LAM = scenes.lame(*scenes.LAME_SOFT)


@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("sigma", [0.01, 0.1, 0.3])
def test_prox_tets_match_reference(cpu, model, sigma):
    z = checkers.random_F(400, sigma, seed=100 + model)
    kappa = 0.0 if model < 3 else 1000.0
    a, _ = checkers.prox_tets("oracle", model, MU, LAM, z, kappa)
    b, _ = checkers.prox_tets("ref", model, MU, LAM, z, kappa)
    # both run the same L-BFGS in fp64; differences come from the SVD (products agree to rounding)
    # amplified by the 1e-6 stopping rule of the reference (src/TetEnergyTerm.hpp:93-95)
    # (relative for the rare elements where the reference's own line search runs away: the spline
    # compression term -kappa x^3 is unbounded below, src/XuSpline.hpp:44)
    err = (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()
    assert err < 5e-6, err
    if model == 0:
        assert np.abs(a - b).max() < 1e-12


def test_prox_tets_inverted_and_collapsed(cpu):
    z = checkers.random_F(64, 0.2, seed=7)
    z[:32, 6:9] *= -1.0           # inverted (det < 0)
    z[32:40] *= 1e-9              # collapsed to a point
    for model in (0, 1, 2):
        a, _ = checkers.prox_tets("oracle", model, MU, LAM, z)
        b, _ = checkers.prox_tets("ref", model, MU, LAM, z)
        ok = np.abs(a - b).max(axis=1) < 5e-6
        # collapsed elements have no defined rotation: compare singular values of the result instead
        sa = np.linalg.svd(a.reshape(-1, 3, 3), compute_uv=False)
        sb = np.linalg.svd(b.reshape(-1, 3, 3), compute_uv=False)
        assert np.abs(sa - sb).max() < 5e-6
        assert ok[:32].all()


@pytest.mark.parametrize("limits", [(-100.0, 100.0), (0.95, 1.05)])
def test_prox_tris_match_reference(cpu, limits):
    rng = np.random.RandomState(5)
    z = np.zeros((300, 6))
    z[:, 0] = 1.0
    z[:, 4] = 1.0
    z += 0.2 * rng.randn(300, 6)
    a = checkers.prox_tris("oracle", 100.0, 100.0, z, *limits)
    b = checkers.prox_tris("ref", 100.0, 100.0, z, *limits)
    assert np.abs(a - b).max() < 1e-12


def _pair(pkg, scene, model, **kw):
    ref = scenes.build_tet_scene(CpuSolver("ref"), scene, model, **kw)
    if kw.get("linsolver", 0) == 1 and kw.get("colors") is None:
        kw = dict(kw, colors=ref.get_colors())
    orc = scenes.build_tet_scene(CpuSolver("oracle"), scene, model, **kw)
    return ref, orc


@pytest.mark.parametrize("model", [0, 1, 2])
def test_system_matrix_matches_reference(pkg, cpu, model):
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    ref, orc = _pair(pkg, scene, model, linsolver=0)
    A, B = ref.matrix_A(), orc.matrix_A()
    assert A.shape == B.shape
    d = abs(A - B)
    assert d.max() <= 1e-12 * abs(A).max()


@pytest.mark.parametrize("model,linsolver", [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (1, 2)])
def test_steps_match_reference(pkg, cpu, model, linsolver):
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    ref, orc = _pair(pkg, scene, model, linsolver=linsolver, iters=10)
    x0 = scenes.bend(scene[0]).ravel()
    for s in (ref, orc):
        s.set_x(x0)
    for step in range(3):
        ref.step()
        orc.step()
    err = np.abs(ref.get_x() - orc.get_x()).max()
    # fp64 on both sides; the only slack is the 1e-6 prox stopping rule hit from different SVD roundings
    assert err < 2e-7, err
    assert ref.runtime_data()["inner_iters"] == orc.runtime_data()["inner_iters"]


def test_traced_step_matches_reference(pkg, cpu):
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    ref, orc = _pair(pkg, scene, 1, linsolver=1, iters=4)
    x0 = scenes.bend(scene[0]).ravel()
    for s in (ref, orc):
        s.set_x(x0)
    zr, ur, br, xr = ref.traced_step(4)
    zo, uo, bo, xo = orc.traced_step(4)
    assert np.abs(zr - zo).max() < 5e-6
    assert np.abs(ur - uo).max() < 5e-6
    assert np.abs(br - bo).max() < 1e-6 * np.abs(br).max()
    assert np.abs(xr - xo).max() < 1e-7


def test_floor_in_gauss_seidel_matches_reference(pkg, cpu):
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    floor_y = scene[0][:, 1].min() - 0.02
    ref, orc = _pair(pkg, scene, 2, linsolver=1, iters=8, floor=floor_y, pin=False)
    for step in range(6):
        ref.step()
        orc.step()
    xr, xo = ref.get_x(), orc.get_x()
    assert xr.reshape(-1, 3)[:, 1].min() >= floor_y - 1e-9
    assert (np.abs(xr.reshape(-1, 3)[:, 1] - floor_y) < 1e-12).any()  # the floor was actually hit
    assert np.abs(xr - xo).max() < 2e-7


def test_sphere_in_gauss_seidel_matches_reference(pkg, cpu):
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    c = np.array([0.0, scene[0][:, 1].min() - 0.45, 0.0])
    ref, orc = _pair(pkg, scene, 1, linsolver=1, iters=8, sphere=(c, 0.5), pin=False)
    for step in range(6):
        ref.step()
        orc.step()
    assert np.abs(ref.get_x() - orc.get_x()).max() < 2e-7


def test_uzawa_with_floor_matches_reference(pkg, cpu):
    # Vertices that a constrained solve leaves exactly ON the floor are re-tested with dx < 0 in the next
    # ADMM iteration (src/PassiveObject.hpp:38-39), so whole trajectories flip on 1e-16 noise.  The
    # comparison is therefore solve by solve: the oracle gets the reference's own (x_in, b) of every
    # ADMM iteration, in order (the multipliers y are warm-started across solves, src/UzawaCG.hpp:69-74).
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    floor_y = scene[0][:, 1].min() - 0.02
    ref, orc = _pair(pkg, scene, 1, linsolver=2, iters=8, floor=floor_y, pin=False)
    n_constrained = 0
    for step in range(4):
        x_prev = ref.get_x() + (1.0 / 24) * (ref.get_v() + np.tile([0, (1.0 / 24) * -9.8, 0], ref.dof // 3))
        z, u, b, x = ref.traced_step(8)
        for it in range(8):
            x_in = x_prev if it == 0 else x[it - 1]
            xo, iters = orc.linsolve(x_in, b[it])
            assert np.abs(xo - x[it]).max() < 1e-10
            n_constrained += int((x_in.reshape(-1, 3)[:, 1] < floor_y).any())
    assert n_constrained > 5  # the constrained branch was exercised


@pytest.mark.parametrize("model,linsolver", [(1, 0), (2, 0), (1, 1), (0, 2)])
def test_unstructured_mesh_steps_match_reference(pkg, cpu, model, linsolver):
    """A Delaunay mesh instead of the block beams: irregular valence, poorly shaped elements near the hull, more
    colours.  Whole steps (prox of every element, assembly, LDLT / multi-colour Gauss-Seidel with the reference's own
    colours / Uzawa)."""
    scene = scenes.blob(pkg.meshes)
    assert len(scene[1]) > 400
    ref, orc = _pair(pkg, scene, model, linsolver=linsolver, iters=8)
    if linsolver == 1:
        assert len(ref.get_colors()) >= 5
    x0 = scenes.bend(scene[0], 0.08).ravel()
    ref.set_x(x0)
    orc.set_x(x0)
    for step in range(3):
        ref.step()
        orc.step()
    assert np.abs(ref.get_x() - orc.get_x()).max() < 5e-7


def test_uzawa_with_sphere_matches_reference(pkg, cpu):
    """As above with a Sphere (normals differ per hit, src/PassiveObject.hpp:47-64): solve by solve on the
    reference's own (x_in, b)."""
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    c = np.array([scene[0][:, 0].mean(), scene[0][:, 1].min() - 0.45, scene[0][:, 2].mean()])
    ref, orc = _pair(pkg, scene, 1, linsolver=2, iters=8, sphere=(c, 0.5), pin=False)
    n_constrained = 0
    for step in range(5):
        x_prev = ref.get_x() + (1.0 / 24) * (ref.get_v() + np.tile([0, (1.0 / 24) * -9.8, 0], ref.dof // 3))
        z, u, b, x = ref.traced_step(8)
        for it in range(8):
            x_in = x_prev if it == 0 else x[it - 1]
            xo, iters = orc.linsolve(x_in, b[it])
            assert np.abs(xo - x[it]).max() < 1e-10
            n_constrained += int((np.linalg.norm(x_in.reshape(-1, 3) - c, axis=1) < 0.5).any())
    assert n_constrained > 3  # the constrained branch was exercised


@pytest.mark.parametrize("linsolver", [0, 2])
@pytest.mark.parametrize("limits", [(-100.0, 100.0), (0.95, 1.05)])
def test_cloth_steps_match_reference(pkg, cpu, linsolver, limits):
    v64, tris, masses, pins = scenes.cloth(pkg.meshes, 8)
    mu, lam = scenes.lame(100.0, 0.1)  # samples/sca2016/trianglestrain.cpp
    sol = []
    for kind in ("ref", "oracle"):
        s = CpuSolver(kind)
        s.add_nodes(v64, masses)
        s.add_tris(v64, tris, mu, lam, *limits)
        s.set_pins(pins)
        assert s.initialize(dt=1.0 / 24, admm_iters=10, gravity=-9.8, linsolver=linsolver)
        sol.append(s)
    for step in range(4):
        for s in sol:
            s.step()
    assert np.abs(sol[0].get_x() - sol[1].get_x()).max() < 1e-9


def test_moving_pins_match_reference(pkg, cpu):
    # stretch_beams (samples/sca2016/beams.cpp:107-133): pins move every frame through set_pins
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    v64, tets, masses, pins = scene
    right = np.nonzero(v64[:, 0] > v64[:, 0].max() - 1e-2)[0].astype(np.int32)
    allp = np.concatenate([pins, right])
    pts = v64[allp].copy()
    sol = []
    for kind in ("ref", "oracle"):
        s = CpuSolver(kind)
        s.add_nodes(v64, masses)
        s.add_tets(v64, tets, 1, MU, LAM)
        s.set_pins(allp, pts)
        assert s.initialize(dt=1.0 / 24, admm_iters=10, gravity=-9.8, linsolver=0)
        sol.append(s)
    for step in range(4):
        pts[:len(pins), 0] -= 1.0 / 24
        pts[len(pins):, 0] += 1.0 / 24
        for s in sol:
            s.set_pins(allp, pts)
            s.step()
    assert np.abs(sol[0].get_x() - sol[1].get_x()).max() < 2e-7


def test_uzawa_surface_inds_and_constraint_weight_match_reference(pkg, cpu):
    """Solver::surface_inds (only those vertices are tested for hits, and in that order, src/Solver.cpp:93) and
    Settings::constraint_w (-ck: C and c scaled by sqrt(w), src/ConstraintSet.hpp:66,84-88, which moves the r^2 < tol^2
    exit of the conjugate gradients): the oracle solve by solve on the reference's own (x_in, b)."""
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    v = scene[0]
    floor_y = v[:, 1].min() - 0.02
    # "surface": every vertex of the outer faces, deliberately NOT in node order
    surf = np.nonzero((np.abs(v - v.min(0)) < 1e-9).any(axis=1) | (np.abs(v - v.max(0)) < 1e-9).any(axis=1))[0].astype(np.int32)[::-1].copy()
    assert 0 < len(surf) < len(v)
    cw = 9.0
    ref, orc = CpuSolver("ref"), CpuSolver("oracle")
    for s in (ref, orc):
        s.set_surface_inds(surf, cw)
    scenes.build_tet_scene(ref, scene, 1, linsolver=2, iters=8, floor=floor_y, pin=False, constraint_w=cw)
    scenes.build_tet_scene(orc, scene, 1, linsolver=2, iters=8, floor=floor_y, pin=False)
    n_constrained = 0
    for step in range(4):
        x_prev = ref.get_x() + (1.0 / 24) * (ref.get_v() + np.tile([0, (1.0 / 24) * -9.8, 0], ref.dof // 3))
        z, u, b, x = ref.traced_step(8)
        for it in range(8):
            x_in = x_prev if it == 0 else x[it - 1]
            xo, iters = orc.linsolve(x_in, b[it])
            assert np.abs(xo - x[it]).max() < 1e-10
            n_constrained += int((x_in.reshape(-1, 3)[surf, 1] < floor_y).any())
    assert n_constrained > 5


@pytest.mark.parametrize("spline_type", [0, 1, 2])
def test_spline_tet_with_its_own_constants_matches_reference(pkg, cpu, spline_type):
    """SplineTet(tet, verts, lame, spline) with spline constants different from the element's Lame
    (src/TetEnergyTerm.hpp:200-205): the prox penalty K and the weight come from the Lame, the energy from the spline."""
    mu, lam = scenes.lame(*scenes.LAME_SOFT)
    spline = (0.6 * mu, 1.7 * lam, 0.0)
    z = checkers.random_F(600, 0.15, seed=40 + spline_type)
    zr, rc = checkers.prox_spline_tets("ref", spline_type, mu, lam, spline, z)
    zo, _ = checkers.prox_spline_tets("oracle", spline_type, mu, lam, spline, z)
    assert rc == 0 and np.abs(zr - zo).max() < 5e-6
    # and it is NOT what the same spline gives with K taken from its own constants (the bug this test pins)
    zw, _ = checkers.prox_tets("oracle", 3 + spline_type, spline[0], spline[1], z)
    assert np.abs(zw - zr).max() > 1e-3
    # whole steps
    scene = scenes.beam(pkg.meshes, 5, 2, 2)
    sol = []
    for kind in ("ref", "oracle"):
        s = CpuSolver(kind)
        s.add_nodes(scene[0], scene[2])
        s.add_spline_tets(scene[0], scene[1], spline_type, mu, lam, spline)
        s.set_pins(scene[3])
        assert s.initialize(dt=1.0 / 24, admm_iters=8, gravity=-9.8, linsolver=0)
        s.set_x(scenes.bend(scene[0]).ravel())
        sol.append(s)
    for _ in range(3):
        for s in sol:
            s.step()
    assert np.abs(sol[0].get_x() - sol[1].get_x()).max() < 5e-7
