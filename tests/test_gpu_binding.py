"""The reference-side binding: integration/GpuSolver.hpp (admm::GpuSolver : admm::Solver), compiled against the
reference's own headers by oracle/Makefile (oracle/gpu_binding.cpp -> oracle/_ref/libadmm_gpubinding.so).

  * the reference's UNMODIFIED test program samples/tests/test_lineartet.cpp runs through the binding
    (x = 52.2321 +- 1e-4 for 21..99 ADMM iterations, inversion recovery; fp64 element mode);
  * what the binding harvests from the reference's get_reduction() triplets -- vertex ids, Dm^-1, weights, g_index --
    equals what this repository's host mirror computes (indices bit-exact; values to the last bits);
  * GPU step() against the reference's own Solver::step() on the SAME object (cpu_step) for every global solver:
    NodalMultiColorGS with the reference's colour lists, LDLT and UzawaCG with the factor of the reference's own
    Eigen::SimplicialLDLT, moving SpringPins, Floor, cloth with strain limits.
"""
import os

import numpy as np
import pytest

import checkers
import scenes

MU, LAM = scenes.lame(*scenes.LAME_SOFT)


def need_binding():
    if not checkers.have_binding():
        pytest.skip("oracle/_ref/libadmm_gpubinding.so is not built (needs /root/reference at build time)")


def test_binding_library_exports(cpu):
    """No GPU needed: the binding library loads (it links libadmm_b200.so through the C-ABI) and exports its entry points."""
    need_binding()
    L = checkers.binding_lib()
    for fn in ("gpub_run_reference_lineartet_test", "gpub_create", "gpub_initialize", "gpub_step", "gpub_cpu_step", "gpub_get_tets", "gpub_get_tris",
               "gpub_get_pins", "gpub_get_colors"):
        assert hasattr(L, fn), fn


def test_binding_harvest_equals_mirror_cpu(pkg, cpu):
    """No GPU needed for the mirror side (terms are host descriptors): the harvest needs initialize() and therefore a
    device, so here only the mirror's own rest data is checked against the oracle's F-layout known answer."""
    scene = scenes.beam(pkg.meshes, 3, 2, 2)
    s = pkg.Solver()
    s.add_nodes(scene[0], scene[2])
    s.add_tets(scene[0], scene[1], 1, MU, LAM)
    idx, dminv, w, row = s.tet_rest_data()
    assert (idx == scene[1]).all()
    # F = Ds * Dm^-1 = I at rest
    p = scene[0][idx]
    Ds = np.stack([p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], p[:, 3] - p[:, 0]], axis=2)
    F = Ds @ dminv.reshape(-1, 3, 3)
    assert np.abs(F - np.eye(3)).max() < 1e-12


@pytest.mark.gpu
def test_reference_lineartet_program_through_the_binding(cpu):
    """samples/tests/test_lineartet.cpp, unmodified, with admm::Solver spelled admm::GpuSolver: prints SUCCESS, returns 0."""
    need_binding()
    old = os.environ.get("ADMM_B200_PRECISION")
    os.environ["ADMM_B200_PRECISION"] = "64"   # |x| = 200 m in that test: fp32 deformation gradients resolve 1.5e-5
    try:
        assert checkers.binding_lib().gpub_run_reference_lineartet_test() == 0
    finally:
        if old is None:
            del os.environ["ADMM_B200_PRECISION"]
        else:
            os.environ["ADMM_B200_PRECISION"] = old


@pytest.mark.gpu
def test_harvest_equals_mirror(pkg, cpu):
    need_binding()
    scene = scenes.beam(pkg.meshes, 6, 3, 2)
    b = checkers.GpuBinding(precision=1)
    b.add_nodes(scene[0], scene[2])
    b.add_tets(scene[0], scene[1], 1, MU, LAM)
    b.set_pins(scene[3])
    assert b.initialize(admm_iters=2, linsolver=0)
    m = pkg.Solver()
    m.set_options(precision=1)
    scenes.build_tet_scene(m, scene, 1, linsolver=0, iters=2)
    bi, bd, bw, br, bm = b.tets()
    mi, md, mw, mr = m.tet_rest_data()
    assert (bi == mi).all() and (br == mr).all() and (bm == 1).all()      # vertex ids, g_index: bit-exact
    assert np.abs(bw - mw).max() <= 4e-16 * np.abs(mw).max()
    assert np.abs(bd - md).max() <= 1e-14 * np.abs(md).max()
    # SpringPins: 6 rows each after the tets' 9; the reference walks an unordered_map, so compare as sets
    pi, pr, pw = b.pins()
    assert sorted(pi) == sorted(scene[3]) and sorted(pr) == list(9 * len(scene[1]) + 6 * np.arange(len(pi)))
    # cloth
    cl = scenes.cloth(pkg.meshes, 6)
    b2 = checkers.GpuBinding(precision=1)
    b2.add_nodes(cl[0], cl[2])
    b2.add_tris(cl[0], cl[1], 100.0, 50.0, 0.95, 1.05)
    b2.set_pins(cl[3])
    assert b2.initialize(admm_iters=2, linsolver=2)
    m2 = pkg.Solver()
    m2.add_nodes(cl[0], cl[2])
    m2.add_tris(cl[0], cl[1], 100.0, 50.0, 0.95, 1.05)
    ti, tr_, tw, trow = b2.tris()
    qi, qr, qw, qrow = m2.tri_rest_data()
    assert (ti == qi).all() and np.abs(tr_ - qr).max() <= 1e-13 * np.abs(qr).max() and np.abs(tw - qw).max() <= 4e-16 * np.abs(qw).max()
    assert (trow == 6 * np.arange(len(ti))).all()


def _ab(b, x0, steps, mover=None):
    """GPU steps, then the reference's CPU steps from the same start on the same object."""
    out = []
    for fn in (b.step, b.cpu_step):
        b.set_x(x0)
        b.set_v(np.zeros_like(x0))
        for k in range(steps):
            if mover is not None:
                mover(k)
            fn()
        out.append(b.get_x())
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("linsolver", [0, 1, 2])
@pytest.mark.parametrize("model", [0, 1, 2])
def test_gpu_step_vs_reference_step_on_the_same_object(pkg, cpu, linsolver, model):
    need_binding()
    scene = scenes.beam(pkg.meshes, 10, 3, 3)
    bbox = float(np.linalg.norm(scene[0].max(0) - scene[0].min(0)))
    for precision, tol in ((1, 1e-6), (0, 1e-4 * bbox)):
        b = checkers.GpuBinding(precision=precision)
        b.add_nodes(scene[0], scene[2])
        b.add_tets(scene[0], scene[1], model, MU, LAM)
        b.set_pins(scene[3])
        assert b.initialize(admm_iters=8, linsolver=linsolver)
        xg, xc = _ab(b, scenes.bend(scene[0]).ravel(), 3)
        err = np.abs(xg - xc).max()
        assert err < tol, (precision, err)
        if linsolver == 1:
            assert len(b.colors()) >= 4 and b.info().startswith("resident")
        else:
            assert b.info().startswith("ldlt")
        b.close()


@pytest.mark.gpu
def test_binding_moving_pins_floor_and_cloth(pkg, cpu):
    need_binding()
    # stretch_beams (samples/sca2016/beams.cpp:107-133): energy-based pins moved every frame through set_pins
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    v64 = scene[0]
    right = np.nonzero(v64[:, 0] > v64[:, 0].max() - 1e-2)[0].astype(np.int32)
    allp = np.concatenate([scene[3], right])
    b = checkers.GpuBinding(precision=1)
    b.add_nodes(v64, scene[2])
    b.add_tets(v64, scene[1], 1, MU, LAM)
    b.set_pins(allp, v64[allp])
    assert b.initialize(admm_iters=10, linsolver=0)

    def mover(k):
        pts = v64[allp].copy()
        pts[:len(scene[3]), 0] -= (k + 1) / 24.0
        pts[len(scene[3]):, 0] += (k + 1) / 24.0
        b.set_pins(allp, pts)
    xg, xc = _ab(b, v64.ravel(), 3, mover)
    assert np.abs(xg - xc).max() < 1e-6
    # Floor inside the Gauss-Seidel sweep (no pins)
    b = checkers.GpuBinding(precision=1)
    b.add_nodes(v64, scene[2])
    b.add_tets(v64, scene[1], 2, MU, LAM)
    b.add_floor(float(v64[:, 1].min() - 0.02))
    assert b.initialize(admm_iters=8, linsolver=1)
    xg, xc = _ab(b, v64.ravel(), 6)
    assert np.abs(xg - xc).max() < 5e-6
    assert np.abs(xg.reshape(-1, 3)[:, 1] - (v64[:, 1].min() - 0.02)).min() < 1e-12   # it landed
    # cloth with strain limits, UzawaCG (empty constraint matrix: the prefactored solve), two corner pins
    cl = scenes.cloth(pkg.meshes, 8)
    b = checkers.GpuBinding(precision=1)
    b.add_nodes(cl[0], cl[2])
    b.add_tris(cl[0], cl[1], 100.0 / 2.2, 100.0 * 0.1 / (1.1 * 0.8), 0.95, 1.05)
    b.set_pins(cl[3])
    assert b.initialize(admm_iters=10, linsolver=2)
    xg, xc = _ab(b, cl[0].ravel(), 4)
    assert np.abs(xg - xc).max() < 1e-8
