"""Parity of the CUDA path (through the C-ABI, include/admm_b200.h) against the CPU oracle and the
committed golden fixtures of the compiled reference.  Needs a B200: run with `-m gpu`.

Tolerances (stated once, used below):
  TOL_Z64   5e-6   prox output, fp64 element path.  The reference stops its L-BFGS at |grad|<1e-6 or
                   |dx|<1e-6 (src/TetEnergyTerm.hpp:93-95), so its own z is only defined to ~1e-6;
                   the GPU Newton iteration converges to the exact minimiser.
  TOL_Z32   5e-5 * max(1,|z|)   prox output, fp32 element path (hard part 1).
  TOL_X64   1e-6   node positions after a few steps, fp64 element path (scene size ~1 m).
  TOL_X32   1e-4 * bbox diagonal   node positions, fp32 element path (SURVEY.md 8d parity gate).
Index work (row offsets, colour lists, incidence) is compared bit-exact.
"""
import json
import os

import numpy as np
import pytest

import checkers
import scenes
from checkers import CpuSolver

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_Z64, TOL_Z32, TOL_X64, TOL_X32_REL = 5e-6, 5e-5, 1e-6, 1e-4
MU, LAM = scenes.lame(*scenes.LAME_SOFT)


def record(name, **kw):
    """Appends measured errors to gpurun_out/parity_report.jsonl (evidence for DESIGN.md)."""
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **{k: float(v) for k, v in kw.items()})) + "\n")
    except OSError:
        pass


@pytest.fixture(scope="module")
def dev(pkg):
    d = pkg.DeviceSolver(0)
    yield d
    d.close()


def colors_from(g, key):
    off, nodes = g[key + "_color_off"], g[key + "_color_nodes"]
    return [nodes[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def bbox_diag(v):
    v = np.asarray(v).reshape(-1, 3)
    return float(np.linalg.norm(v.max(0) - v.min(0)))


def gpu_solver(pkg, precision, colors=None, **opts):
    s = pkg.Solver()
    s.set_options(precision=precision, **opts)
    if colors is not None:
        s.set_colors(colors)
        s.set_options(coloring=pkg.COLOR_USER)
    return s


# ---------------------------------------------------------------------------------------------
# local step: prox kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model", range(6))
@pytest.mark.parametrize("precision", [0, 1])
def test_prox_tets_golden(pkg, dev, model, precision):
    g = np.load(os.path.join(G, "prox_vectors.npz"))
    mu, lam = g["mu_lambda"]
    z_in, z_ref = g["tet%d_in" % model], g["tet%d_out" % model]
    out = dev.prox_tets(model, mu, lam, z_in, precision=precision)
    err = (np.abs(out - z_ref) / np.maximum(1.0, np.abs(z_ref))).max()
    record("prox_tets_golden", model=model, precision=precision, err=err)
    assert err < (TOL_Z64 if precision else TOL_Z32), err


@pytest.mark.parametrize("model", range(6))
@pytest.mark.parametrize("sigma", [1e-4, 0.01, 0.1, 0.3, 0.6])
def test_prox_tets_vs_oracle(pkg, dev, cpu, model, sigma):
    n = 4096 + 37  # ragged: not a multiple of the warp or block size
    z = checkers.random_F(n, sigma, seed=300 + model)
    ref, _ = checkers.prox_tets("oracle", model, MU, LAM, z)
    for precision, tol in ((1, TOL_Z64), (0, TOL_Z32)):
        out = dev.prox_tets(model, MU, LAM, z, precision=precision)
        e = np.abs(out - ref) / np.maximum(1.0, np.abs(ref))
        record("prox_tets_vs_oracle", model=model, sigma=sigma, precision=precision, err=e.max(), p999=np.quantile(e.max(axis=1), 0.999))
        assert e.max() < tol, (precision, e.max())


@pytest.mark.parametrize("model", [3, 4, 5])
def test_prox_spline_compression_term(pkg, dev, model):
    """kappa > 0 (xu::Spline::compress_term, src/XuSpline.hpp:43-45) makes the reference's objective
    unbounded below (-kappa J^3): its first steepest-descent step (alpha = 1 against a gradient of order
    K*strain) throws about a third of all elements to stretches of 1e5 and the line search accepts
    them -- there is nothing to be in parity WITH.  Required here: finite output, and the local
    minimiser next to the start, i.e. the kappa = 0 answer up to the (tiny) kappa/K perturbation."""
    z = checkers.random_F(3000, 0.2, seed=21)
    base = dev.prox_tets(model, MU, LAM, z, kappa=0.0, precision=1)
    out = dev.prox_tets(model, MU, LAM, z, kappa=1000.0, precision=1)
    assert np.isfinite(out).all()
    assert np.abs(out - base).max() < 1e-4


@pytest.mark.parametrize("spline_type", [0, 1, 2])
def test_spline_tet_with_its_own_constants(pkg, dev, cpu, spline_type):
    """SplineTet whose spline constants differ from the element's Lame: the prox penalty K is the Lame's bulk modulus
    (src/TetEnergyTerm.hpp:193-205), handed over separately (admm_b200_add_tets: bulk_modulus)."""
    spline = (0.6 * MU, 1.7 * LAM, 0.0)
    K = LAM + (2.0 / 3.0) * MU
    z = checkers.random_F(2000 + 5, 0.15, seed=60 + spline_type)
    ref, _ = checkers.prox_spline_tets("oracle", spline_type, MU, LAM, spline, z)
    for precision, tol in ((1, TOL_Z64), (0, TOL_Z32)):
        out = dev.prox_tets(3 + spline_type, spline[0], spline[1], z, precision=precision, bulk_modulus=K)
        e = (np.abs(out - ref) / np.maximum(1.0, np.abs(ref))).max()
        record("prox_spline_own_constants", spline=spline_type, precision=precision, err=e)
        assert e < tol, (precision, e)
    # whole steps through the host mirror (SplineTet descriptor carries the Lame and the spline)
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    gpu = gpu_solver(pkg, 1)
    orc = CpuSolver("oracle")
    for s in (gpu, orc):
        s.add_nodes(scene[0], scene[2])
        s.add_spline_tets(scene[0], scene[1], spline_type, MU, LAM, spline)
        s.set_pins(scene[3])
        assert s.initialize(dt=1.0 / 24, admm_iters=8, gravity=-9.8, linsolver=0)
        s.set_x(scenes.bend(scene[0]).ravel())
    for _ in range(3):
        gpu.step()
        orc.step()
    assert np.abs(gpu.get_x() - orc.get_x()).max() < TOL_X64


def test_prox_tets_inverted_collapsed_and_rest(pkg, dev, cpu):
    z = checkers.random_F(96, 0.2, seed=7)
    z[:32, 6:9] *= -1.0           # inverted (det F < 0), src/TetEnergyTerm.cpp:122,131
    z[32:40] *= 1e-9              # collapsed to a point, :126-129
    z[40:48] = np.eye(3).T.ravel()  # exactly at rest (the reference's line search spins here)
    for model in (0, 1, 2):
        ref, _ = checkers.prox_tets("oracle", model, MU, LAM, z)
        for precision, tol in ((1, TOL_Z64), (0, TOL_Z32)):
            out = dev.prox_tets(model, MU, LAM, z, precision=precision)
            assert np.isfinite(out).all()
            ok = np.abs(out - ref).max(axis=1)
            assert ok[:32].max() < tol and ok[40:].max() < tol, (model, precision, ok.max())
            # collapsed elements have no defined rotation: compare the singular values of the result
            sa = np.linalg.svd(out[32:40].reshape(-1, 3, 3), compute_uv=False)
            sb = np.linalg.svd(ref[32:40].reshape(-1, 3, 3), compute_uv=False)
            assert np.abs(sa - sb).max() < tol
            np.testing.assert_allclose(out[40:48], z[40:48], atol=tol)


@pytest.mark.parametrize("precision", [0, 1])
def test_prox_tris_golden_and_oracle(pkg, dev, cpu, precision):
    g = np.load(os.path.join(G, "prox_vectors.npz"))
    tol = 1e-10 if precision else 2e-6
    assert np.abs(dev.prox_tris(g["tri_in"], precision=precision) - g["tri_out"]).max() < tol
    assert np.abs(dev.prox_tris(g["tri_in"], 0.95, 1.05, precision=precision) - g["tri_lim_out"]).max() < tol
    rng = np.random.RandomState(5)
    z = np.zeros((1000 + 13, 6))
    z[:, 0] = 1.0
    z[:, 4] = 1.0
    z += 0.3 * rng.randn(*z.shape)
    for lim in ((-100.0, 100.0), (0.9, 1.1)):
        ref = checkers.prox_tris("oracle", 100.0, 100.0, z, *lim)
        err = np.abs(dev.prox_tris(z, *lim, precision=precision) - ref).max()
        record("prox_tris_vs_oracle", precision=precision, err=err)
        assert err < tol, err


def test_prox_empty_and_single(pkg, dev):
    assert dev.prox_tets(1, MU, LAM, np.zeros((0, 9))).shape == (0, 9)
    assert dev.prox_tris(np.zeros((0, 6))).shape == (0, 6)
    one = dev.prox_tets(1, MU, LAM, np.eye(3).ravel()[None, :])
    np.testing.assert_allclose(one[0], np.eye(3).ravel(), atol=1e-6)


def test_prox_rotation_equivariance_full_size(pkg, dev):
    """Size-independent property at the bench size (1M tets): prox(R F Q) = R prox(F) Q for rotations
    R, Q (the energies are isotropic), and prox is the identity on rotations."""
    n = 1000000
    rng = np.random.RandomState(11)
    F = checkers.random_F(n, 0.2, seed=12).reshape(n, 3, 3).transpose(0, 2, 1)  # row-major matrices
    R, _ = np.linalg.qr(rng.randn(3, 3))
    R *= np.sign(np.linalg.det(R))
    Q, _ = np.linalg.qr(rng.randn(3, 3))
    Q *= np.sign(np.linalg.det(Q))

    def cm(M):
        return np.ascontiguousarray(M.transpose(0, 2, 1).reshape(-1, 9))

    def rm(z):
        return z.reshape(-1, 3, 3).transpose(0, 2, 1)

    for model in (1, 2):
        a = rm(dev.prox_tets(model, MU, LAM, cm(F), precision=0))
        b = rm(dev.prox_tets(model, MU, LAM, cm(R @ F @ Q), precision=0))
        err = np.abs(R @ a @ Q - b).max()
        record("prox_equivariance_1M", model=model, err=err)
        assert err < TOL_Z32, err
    rot = np.broadcast_to(R, (1024, 3, 3))
    out = rm(dev.prox_tets(1, MU, LAM, cm(rot), precision=0))
    assert np.abs(out - rot).max() < 1e-5


# ---------------------------------------------------------------------------------------------
# whole steps against the golden fixtures of the compiled reference
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("model,linsolver", [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (1, 2)])
@pytest.mark.parametrize("precision", [0, 1])
def test_beam_steps_golden(pkg, model, linsolver, precision):
    g = np.load(os.path.join(G, "beam_steps.npz"))
    scene = (g["verts"], g["tets"], g["masses"], g["pins"])
    key = "m%d_ls%d" % (model, linsolver)
    colors = colors_from(g, key) if linsolver == 1 else None
    tol_x = TOL_X64 if precision else TOL_X32_REL * bbox_diag(g["verts"])
    tol_z = TOL_Z64 if precision else TOL_Z32
    R = 9 * len(g["tets"])

    # (a) the first ADMM iteration in detail: z, u (reference row layout), b, x
    s = gpu_solver(pkg, precision, colors, keep_z=True)
    scenes.build_tet_scene(s, scene, model, linsolver=linsolver, iters=1, colors=None)
    ro = s.row_offsets()
    assert (ro[:len(g["tets"])] == 9 * np.arange(len(g["tets"]))).all()  # g_index, bit-exact
    s.set_x(g["x0"].ravel())
    s.step()
    d = s.device()
    z, u, b = d.debug_get("z", s.n_rows()), d.debug_get("u", s.n_rows()), d.debug_get("b", s.dof)
    ez = np.abs(z[:R] - g[key + "_z_it0"][:R]).max()
    eu = np.abs(u[:R] - g[key + "_u_it0"][:R]).max()
    eb = np.abs(b - g[key + "_b_it0"]).max() / np.abs(g[key + "_b_it0"]).max()
    ex = np.abs(s.get_x() - g[key + "_x_it"][0]).max()
    record("beam_iter0", model=model, linsolver=linsolver, precision=precision, ez=ez, eu=eu, eb=eb, ex=ex)
    assert ez < tol_z and eu < tol_z, (ez, eu)
    assert eb < (1e-6 if precision else 1e-5), eb
    assert ex < tol_x, ex

    # (b) x after k ADMM iterations (u restarts at 0 every step, so a k-iteration step from the same
    # state reproduces the k-th iterate of the reference's 10-iteration step)
    for k in (3, 10):
        s = gpu_solver(pkg, precision, colors)
        scenes.build_tet_scene(s, scene, model, linsolver=linsolver, iters=k, colors=None)
        s.set_x(g["x0"].ravel())
        s.step()
        ex = np.abs(s.get_x() - g[key + "_x_it"][k - 1]).max()
        record("beam_iter_k", model=model, linsolver=linsolver, precision=precision, k=k, ex=ex)
        assert ex < tol_x, (k, ex)
    # (c) two more full steps: positions and velocities
    s.step()
    s.step()
    ex = np.abs(s.get_x() - g[key + "_x3"]).max()
    ev = np.abs(s.get_v() - g[key + "_v3"]).max()
    record("beam_3steps", model=model, linsolver=linsolver, precision=precision, ex=ex, ev=ev)
    assert ex < tol_x, ex
    assert ev < 24 * 2 * tol_x, ev
    if linsolver == 1:
        assert s.runtime_data()["inner_iters"] == 30 * 10  # never converges early


@pytest.mark.parametrize("precision", [0, 1])
def test_floor_golden(pkg, precision):
    g = np.load(os.path.join(G, "beam_steps.npz"))
    scene = (g["verts"], g["tets"], g["masses"], g["pins"])
    floor_y = float(g["floor_y"][0])
    s = gpu_solver(pkg, precision, colors_from(g, "floor"))
    scenes.build_tet_scene(s, scene, 2, linsolver=1, iters=8, floor=floor_y, pin=False)
    for _ in range(6):
        s.step()
    x = s.get_x()
    err = np.abs(x - g["floor_x6"]).max()
    record("floor_golden", precision=precision, err=err)
    assert x.reshape(-1, 3)[:, 1].min() >= floor_y - 1e-12
    assert (np.abs(x.reshape(-1, 3)[:, 1] - floor_y) < 1e-12).any()  # the floor was actually hit
    assert err < (5e-6 if precision else TOL_X32_REL * bbox_diag(g["verts"])), err


@pytest.mark.parametrize("linsolver", [0, 2])
@pytest.mark.parametrize("name,limits", [("nolim", (-100.0, 100.0)), ("lim", (0.95, 1.05))])
@pytest.mark.parametrize("precision", [0, 1])
def test_cloth_golden(pkg, linsolver, name, limits, precision):
    g = np.load(os.path.join(G, "cloth_steps.npz"))
    mu, lam = g["mu_lambda"]
    s = gpu_solver(pkg, precision)
    s.add_nodes(g["verts"], g["masses"])
    s.add_tris(g["verts"], g["tris"], mu, lam, *limits)
    s.set_pins(g["pins"])
    assert s.initialize(dt=1.0 / 24, admm_iters=10, gravity=-9.8, linsolver=linsolver)
    for _ in range(4):
        s.step()
    err = np.abs(s.get_x() - g["ls%d_%s_x4" % (linsolver, name)]).max()
    record("cloth_golden", linsolver=linsolver, limits=limits[1], precision=precision, err=err)
    assert err < (1e-8 if precision else TOL_X32_REL * bbox_diag(g["verts"])), err


def test_single_tet_known_answers(pkg):
    """samples/tests/test_lineartet.cpp:165-230 (x = 52.2321 +- 1e-4 for admm_iters 21..99) and the
    golden iterates of the compiled reference."""
    g = np.load(os.path.join(G, "single_tet.npz"))
    V = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    T = np.array([[0, 1, 2, 3]], dtype=np.int32)
    mu, lam = scenes.lame(500000, 0.25)
    for precision in (1, 0):
        for it, xg in zip(g["iters"], g["x"]):
            s = gpu_solver(pkg, precision)
            s.add_nodes(V, np.ones(12))
            s.add_tets(V, T, 0, mu, lam)
            assert s.initialize(dt=float(np.float32(1) / np.float32(24)), admm_iters=int(it), gravity=0.0, linsolver=0)
            x = V.ravel().copy()
            x[9:12] = [200, 0, 0]
            s.set_x(x)
            s.step()
            err = np.abs(s.get_x() - xg).max()
            record("single_tet", precision=precision, iters=it, err=err)
            assert err < (1e-9 if precision else 2e-3), err  # |x| = 200: fp32 F has ulp 1.5e-5
            if it > 20:
                assert abs(s.get_x()[9] - 52.2321) < (1e-4 if precision else 1e-3)


def test_inversion_known_answer(pkg):
    """samples/tests/test_lineartet.cpp:236-323: an inverted tet recovers its rest volume and the answer
    does not depend on the ADMM iteration count."""
    V = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    T = np.array([[0, 1, 2, 3]], dtype=np.int32)

    def volume(x):
        p = x.reshape(4, 3)
        return np.linalg.det(np.stack([p[1] - p[0], p[2] - p[0], p[3] - p[0]], axis=1)) / 6.0

    first = None
    for it in (10, 25, 60, 99):
        s = gpu_solver(pkg, 1)
        s.add_nodes(V, np.ones(12))
        s.add_tets(V, T, 0, 100.0, 100.0)
        assert s.initialize(dt=0.7, admm_iters=it, gravity=0.0, linsolver=0)
        x = V.ravel().copy()
        x[0:3] = [1, 1, 1]
        s.set_x(x)
        for _ in range(10):
            s.step()
        x = s.get_x()
        assert volume(x) > 0 and abs(volume(x) - 1.0 / 6.0) < 1e-6
        if first is None:
            first = x[0:3].copy()
        assert np.abs(x[0:3] - first).max() < 1e-6


# ---------------------------------------------------------------------------------------------
# GPU vs the C oracle on freshly built scenes (same arrays, same colours)
# ---------------------------------------------------------------------------------------------
def _pair(pkg, scene, model, precision, linsolver, iters, **kw):
    gpu = gpu_solver(pkg, precision)
    scenes.build_tet_scene(gpu, scene, model, linsolver=linsolver, iters=iters, **kw)
    cpu = CpuSolver("oracle")
    colors = gpu.colors() if linsolver == 1 else None
    scenes.build_tet_scene(cpu, scene, model, linsolver=linsolver, iters=iters, colors=colors, **kw)
    return gpu, cpu


@pytest.mark.parametrize("model", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("linsolver", [0, 1])
def test_steps_match_oracle(pkg, cpu, model, linsolver):
    scene = scenes.beam(pkg.meshes, 10, 3, 3)
    for precision in (1, 0):
        gpu, orc = _pair(pkg, scene, model, precision, linsolver, 10)
        x0 = scenes.bend(scene[0]).ravel()
        gpu.set_x(x0)
        orc.set_x(x0)
        for _ in range(5):
            gpu.step()
            orc.step()
        err = np.abs(gpu.get_x() - orc.get_x()).max()
        record("steps_vs_oracle", model=model, linsolver=linsolver, precision=precision, err=err)
        assert err < (TOL_X64 if precision else TOL_X32_REL * bbox_diag(scene[0])), err


def test_system_matrix_and_indices_bit_exact(pkg, cpu):
    scene = scenes.beam(pkg.meshes, 6, 3, 2)
    gpu, orc = _pair(pkg, scene, 1, 1, 1, 2)
    n = len(scene[0])
    rp, ci, va = gpu.system_matrix()
    import scipy.sparse as sp
    Ls = sp.csr_matrix((va, ci, rp), shape=(n, n))
    A3 = sp.kron(Ls, sp.identity(3), format="csr") + sp.diags(np.repeat(scene[2], 3))
    Ao = orc.matrix_A()
    assert abs(A3 - Ao).max() < 1e-9 * abs(Ao).max()
    # every node appears in exactly one colour and no two neighbours share one
    cols = gpu.colors()
    flat = np.sort(np.concatenate(cols))
    assert (flat == np.arange(n)).all()
    color_of = np.empty(n, np.int32)
    for c, nodes in enumerate(cols):
        color_of[nodes] = c
    coo = Ls.tocoo()
    off = (coo.row != coo.col) & (coo.data != 0)
    assert (color_of[coo.row[off]] != color_of[coo.col[off]]).all()


@pytest.mark.parametrize("linsolver", [0, 1])
def test_linsolve_matches_oracle(pkg, cpu, linsolver):
    scene = scenes.beam(pkg.meshes, 8, 3, 3)
    gpu, orc = _pair(pkg, scene, 1, 1, linsolver, 2)
    rng = np.random.RandomState(3)
    n3 = gpu.dof
    b = rng.randn(n3) * 10
    x0 = scenes.bend(scene[0]).ravel()
    xg, itg = gpu.device().linsolve(x0, b)
    xo, ito = orc.linsolve(x0, b)
    err = np.abs(xg - xo).max() / np.abs(xo).max()
    record("linsolve", linsolver=linsolver, err=err)
    assert err < 1e-10, err
    if linsolver == 1:
        assert itg == ito == 30


def test_moving_pins_match_oracle(pkg, cpu):
    # stretch_beams (samples/sca2016/beams.cpp:107-133): pins move every frame through set_pins
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    v64, tets, masses, pins = scene
    right = np.nonzero(v64[:, 0] > v64[:, 0].max() - 1e-2)[0].astype(np.int32)
    allp = np.concatenate([pins, right])
    for linsolver in (0, 1):
        pts = v64[allp].copy()
        gpu = gpu_solver(pkg, 1)
        orc = CpuSolver("oracle")
        for s in (gpu, orc):
            s.add_nodes(v64, masses)
            s.add_tets(v64, tets, 1, MU, LAM)
            s.set_pins(allp, pts)
            if s is orc and linsolver == 1:
                s.set_colors(gpu.colors())
            assert s.initialize(dt=1.0 / 24, admm_iters=10, gravity=-9.8, linsolver=linsolver)
        for step in range(4):
            pts[:len(pins), 0] -= 1.0 / 24
            pts[len(pins):, 0] += 1.0 / 24
            for s in (gpu, orc):
                s.set_pins(allp, pts)
                s.step()
        err = np.abs(gpu.get_x() - orc.get_x()).max()
        record("moving_pins", linsolver=linsolver, err=err)
        assert err < TOL_X64, err


def test_floor_and_sphere_match_oracle(pkg, cpu):
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    floor_y = scene[0][:, 1].min() - 0.02
    c = np.array([0.0, scene[0][:, 1].min() - 0.45, 0.0])
    for kw in (dict(floor=floor_y), dict(sphere=(c, 0.5))):
        gpu, orc = _pair(pkg, scene, 1, 1, 1, 8, pin=False, **kw)
        for _ in range(6):
            gpu.step()
            orc.step()
        err = np.abs(gpu.get_x() - orc.get_x()).max()
        record("obstacle_vs_oracle", floor=float("floor" in kw), err=err)
        assert err < 5e-6, err


def test_uzawa_with_collisions_matches_oracle(pkg, cpu):
    """UzawaCG::solve with passive hits (src/UzawaCG.hpp:57-125, csrc/uzawa.cuh): conjugate gradients on the
    Schur complement, the constraint rows detected on the device.  Compared solve by solve on identical
    (x_in, b) -- whole trajectories are chaotic here (a vertex left exactly ON the floor is re-tested with
    dx < 0 next time, src/PassiveObject.hpp:38-39) -- over a sequence of solves, so that both the reset and
    the warm start of the multipliers (:69-74) are exercised."""
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    v = scene[0]
    floor_y = v[:, 1].min() + 0.15   # cuts through the beam: many hits
    c = np.array([v[:, 0].mean(), v[:, 1].min() - 0.3, v[:, 2].mean()])
    rng = np.random.RandomState(5)
    for kw in (dict(floor=floor_y), dict(sphere=(c, 0.75))):
        gpu, orc = _pair(pkg, scene, 1, 1, 2, 4, pin=False, **kw)
        x0 = scenes.bend(v).ravel()
        rp, ci, va = gpu.system_matrix()
        import scipy.sparse as sp
        A = sp.csr_matrix((va, ci, rp), shape=(len(v), len(v)))
        counts = []
        for k in range(6):
            x_in = x0 + (0.0 if k % 2 else 0.02) * rng.randn(x0.size)   # odd solves repeat the hit set: y is warm-started
            b = (A @ (x_in.reshape(-1, 3) + 0.01 * rng.randn(len(v), 3))).ravel()
            xg, itg = gpu.device().linsolve(x_in, b)
            xo, ito = orc.linsolve(x_in, b)
            err = np.abs(xg - xo).max() / np.abs(xo).max()
            record("uzawa_collisions", floor=float("floor" in kw), solve=k, err=err, iters=itg)
            assert err < 1e-8, (k, err)
            assert abs(itg - ito) <= 1, (itg, ito)
            counts.append(itg)
        assert any(c != 1 for c in counts)  # the constrained branch ran (an unconstrained solve returns exactly 1)
        if "floor" in kw:
            assert min(counts) > 3              # many rows: conjugate gradients really iterate
    # whole steps: the solved positions respect the floor up to the CG tolerance
    gpu, _ = _pair(pkg, scene, 1, 1, 2, 8, pin=False, floor=v[:, 1].min() - 0.02)
    for _ in range(8):
        gpu.step()
    rd = gpu.runtime_data()
    assert np.isfinite(gpu.get_x()).all() and rd["inner_iters"] >= 8


def test_uzawa_surface_inds_and_constraint_weight(pkg, cpu):
    """Solver::surface_inds (candidate vertices of the hit detection, in that order) and Settings::constraint_w (-ck) on
    the device's UzawaCG (csrc/uzawa_blocks.cuh) against the oracle, solve by solve with warm-started multipliers."""
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    v = scene[0]
    floor_y = v[:, 1].min() + 0.15
    surf = np.nonzero((np.abs(v - v.min(0)) < 1e-9).any(axis=1) | (np.abs(v - v.max(0)) < 1e-9).any(axis=1))[0].astype(np.int32)[::-1].copy()
    cw = 9.0
    gpu = gpu_solver(pkg, 1)
    gpu.set_surface_inds(surf)
    scenes.build_tet_scene(gpu, scene, 1, linsolver=2, iters=4, floor=floor_y, pin=False, constraint_w=cw)
    orc = CpuSolver("oracle")
    orc.set_surface_inds(surf, cw)
    scenes.build_tet_scene(orc, scene, 1, linsolver=2, iters=4, floor=floor_y, pin=False)
    rp, ci, va = gpu.system_matrix()
    import scipy.sparse as sp
    A = sp.csr_matrix((va, ci, rp), shape=(len(v), len(v)))
    rng = np.random.RandomState(7)
    x0 = scenes.bend(v).ravel()
    counts = []
    for k in range(6):
        x_in = x0 + (0.0 if k % 2 else 0.02) * rng.randn(x0.size)
        b = (A @ (x_in.reshape(-1, 3) + 0.01 * rng.randn(len(v), 3))).ravel()
        xg, itg = gpu.device().linsolve(x_in, b)
        xo, ito = orc.linsolve(x_in, b)
        err = np.abs(xg - xo).max() / np.abs(xo).max()
        record("uzawa_surface_ck", solve=k, err=err, iters=itg)
        assert err < 1e-8, (k, err)
        assert abs(itg - ito) <= 1, (itg, ito)
        counts.append(itg)
    assert min(counts) > 2
    # an interior vertex below the floor is NOT a candidate: it stays below
    inner = np.setdiff1d(np.arange(len(v)), surf)
    if len(inner):
        x_in = x0.copy().reshape(-1, 3)
        x_in[inner[0], 1] = floor_y - 0.5
        b = (A @ x_in).ravel()
        xg, _ = gpu.device().linsolve(x_in.ravel(), b)
        xo, _ = orc.linsolve(x_in.ravel(), b)
        assert np.abs(xg - xo).max() / np.abs(xo).max() < 1e-8


def test_uzawa_with_floor_golden(pkg):
    """The device UzawaCG against the reference's own solves (tests/golden/uzawa_floor.npz): same (x_in, b)
    sequence, multipliers warm-started across solves exactly as in the reference run."""
    g = np.load(os.path.join(G, "uzawa_floor.npz"))
    scene = (g["verts"], g["tets"], g["masses"], np.zeros(0, np.int32))
    s = gpu_solver(pkg, 1)
    scenes.build_tet_scene(s, scene, 1, linsolver=2, iters=int(g["iters"][0]), floor=float(g["floor_y"][0]), pin=False)
    worst = 0.0
    for k in range(len(g["x_in"])):
        x, it = s.device().linsolve(g["x_in"][k], g["b"][k])
        worst = max(worst, np.abs(x - g["x_out"][k]).max())
        assert (it == 1) == (g["hits"][k] == 0) or it == 1  # an unconstrained solve returns exactly 1
    record("uzawa_floor_golden", err=worst)
    assert worst < 1e-9, worst


def test_device_resident_steps_equal_host_steps(pkg):
    """step_device()+sync_state() (state stays in HBM) gives bit-identical results to step() (host
    buffers every step): the e2e path and the resident path are the same arithmetic."""
    scene = scenes.beam(pkg.meshes, 8, 3, 3)
    res = []
    for resident in (False, True):
        s = gpu_solver(pkg, 0)
        scenes.build_tet_scene(s, scene, 1, linsolver=1, iters=5)
        s.set_x(scenes.bend(scene[0]).ravel())
        if resident:
            s.upload_state()
        for _ in range(3):
            s.step_device() if resident else s.step()
        if resident:
            s.sync_state()
        res.append((s.get_x(), s.get_v()))
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all()


def test_deferred_timers(pkg):
    """Deferred timers (admm_b200_set_deferred_timers / _collect_timers): the steps run without a host synchronise
    each, give bit-identical positions, and the collected sums cover every step, iteration and kernel launch."""
    scene = scenes.beam(pkg.meshes, 8, 3, 3)
    res = []
    for deferred in (False, True):
        s = gpu_solver(pkg, 0)
        scenes.build_tet_scene(s, scene, 1, linsolver=1, iters=5)
        s.set_x(scenes.bend(scene[0]).ravel())
        s.upload_state()
        if deferred:
            s.set_timers(False)
            s.device().set_deferred_timers(True)
        for _ in range(3):
            s.step_device()
        if deferred:
            acc = s.device().collect_timers()
            kt = s.device().kernel_times()
            assert acc["steps"] == 3 and acc["inner_iters"] == 3 * 5 * 30
            assert all(n == 15 for _, n in kt.values()) and all(ms > 0 for ms, _ in kt.values())
            assert acc["step_ms"] > 0 and acc["local_ms"] > 0 and acc["global_ms"] > acc["assemble_ms"] > 0
            assert s.device().collect_timers()["steps"] == 0   # collected once
        s.sync_state()
        res.append(s.get_x())
    assert (res[0] == res[1]).all()


def test_no_silent_fallback(pkg):
    """The product path must be the CUDA one: kernels were launched by this handle."""
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    s = gpu_solver(pkg, 0)
    scenes.build_tet_scene(s, scene, 1, linsolver=1, iters=2)
    n0 = s.device().launch_count()
    s.step()
    assert s.device().launch_count() >= n0 + 2 + 2 * 3
