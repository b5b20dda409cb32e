"""The C oracle against the committed golden fixtures (generated from the compiled reference by
tests/golden/gen_golden.py).  Runs anywhere, no GPU, no /root/reference."""
import os

import numpy as np
import pytest

import checkers
import scenes
from checkers import CpuSolver

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def colors_from(g, key):
    off, nodes = g[key + "_color_off"], g[key + "_color_nodes"]
    return [nodes[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.mark.parametrize("model", range(6))
def test_prox_tets_golden(cpu, model):
    g = np.load(os.path.join(G, "prox_vectors.npz"))
    mu, lam = g["mu_lambda"]
    out, rc = checkers.prox_tets("oracle", model, mu, lam, g["tet%d_in" % model])
    assert np.abs(out - g["tet%d_out" % model]).max() < 5e-6


def test_prox_tris_golden(cpu):
    g = np.load(os.path.join(G, "prox_vectors.npz"))
    assert np.abs(checkers.prox_tris("oracle", 100.0, 100.0, g["tri_in"]) - g["tri_out"]).max() < 1e-12
    assert np.abs(checkers.prox_tris("oracle", 100.0, 100.0, g["tri_in"], 0.95, 1.05) - g["tri_lim_out"]).max() < 1e-12


@pytest.mark.parametrize("model,linsolver", [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (1, 2)])
def test_beam_steps_golden(cpu, model, linsolver):
    g = np.load(os.path.join(G, "beam_steps.npz"))
    scene = (g["verts"], g["tets"], g["masses"], g["pins"])
    key = "m%d_ls%d" % (model, linsolver)
    colors = colors_from(g, key) if linsolver == 1 else None
    s = scenes.build_tet_scene(CpuSolver("oracle"), scene, model, linsolver=linsolver, iters=10, colors=colors)
    s.set_x(g["x0"].ravel())
    z, u, b, x = s.traced_step(10)
    # only the tet rows: SpringPin rows follow in unordered_map order in the reference and their 3 dead
    # rows hold whatever the reference's out-of-bounds read found
    R = 9 * len(g["tets"])
    assert np.abs(z[0][:R] - g[key + "_z_it0"][:R]).max() < 5e-6
    assert np.abs(u[0][:R] - g[key + "_u_it0"][:R]).max() < 5e-6
    assert np.abs(b[0] - g[key + "_b_it0"]).max() < 1e-6 * np.abs(b[0]).max()
    assert np.abs(x - g[key + "_x_it"]).max() < 2e-7
    s.step()
    s.step()
    assert np.abs(s.get_x() - g[key + "_x3"]).max() < 2e-7
    assert np.abs(s.get_v() - g[key + "_v3"]).max() < 1e-5


def test_floor_golden(cpu):
    g = np.load(os.path.join(G, "beam_steps.npz"))
    scene = (g["verts"], g["tets"], g["masses"], g["pins"])
    s = scenes.build_tet_scene(CpuSolver("oracle"), scene, 2, linsolver=1, iters=8, floor=float(g["floor_y"][0]), pin=False, colors=colors_from(g, "floor"))
    for _ in range(6):
        s.step()
    assert np.abs(s.get_x() - g["floor_x6"]).max() < 2e-7


@pytest.mark.parametrize("linsolver", [0, 2])
@pytest.mark.parametrize("name,limits", [("nolim", (-100.0, 100.0)), ("lim", (0.95, 1.05))])
def test_cloth_golden(cpu, linsolver, name, limits):
    g = np.load(os.path.join(G, "cloth_steps.npz"))
    mu, lam = g["mu_lambda"]
    s = CpuSolver("oracle")
    s.add_nodes(g["verts"], g["masses"])
    s.add_tris(g["verts"], g["tris"], mu, lam, *limits)
    s.set_pins(g["pins"])
    assert s.initialize(dt=1.0 / 24, admm_iters=10, gravity=-9.8, linsolver=linsolver)
    for _ in range(4):
        s.step()
    assert np.abs(s.get_x() - g["ls%d_%s_x4" % (linsolver, name)]).max() < 1e-9


def test_single_tet_golden(cpu):
    g = np.load(os.path.join(G, "single_tet.npz"))
    V = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    T = np.array([[0, 1, 2, 3]], dtype=np.int32)
    mu, lam = scenes.lame(500000, 0.25)
    for it, xg in zip(g["iters"], g["x"]):
        s = CpuSolver("oracle")
        s.add_nodes(V, np.ones(12))
        s.add_tets(V, T, 0, mu, lam)
        assert s.initialize(dt=float(np.float32(1) / np.float32(24)), admm_iters=int(it), gravity=0.0, linsolver=0)
        x = V.ravel().copy()
        x[9:12] = [200, 0, 0]
        s.set_x(x)
        s.step()
        assert np.abs(s.get_x() - xg).max() < 1e-9
        if it > 20:
            assert abs(s.get_x()[9] - 52.2321) < 1e-4


def test_uzawa_with_floor_golden(cpu):
    """UzawaCG with passive hits, solve by solve against the reference's own (x_in, b) -> x_out of 4 steps x 8
    ADMM iterations (tests/golden/uzawa_floor.npz; the multipliers are warm-started across solves)."""
    g = np.load(os.path.join(G, "uzawa_floor.npz"))
    scene = (g["verts"], g["tets"], g["masses"], np.zeros(0, np.int32))
    s = scenes.build_tet_scene(CpuSolver("oracle"), scene, 1, linsolver=2, iters=int(g["iters"][0]), floor=float(g["floor_y"][0]), pin=False)
    assert (g["hits"] > 0).sum() > 5
    for k in range(len(g["x_in"])):
        x, _ = s.linsolve(g["x_in"][k], g["b"][k])
        assert np.abs(x - g["x_out"][k]).max() < 1e-10, k


@pytest.mark.parametrize("linsolver", [0, 1])
def test_unstructured_mesh_golden(cpu, linsolver):
    """Delaunay blob (irregular valence, >= 5 colours): the oracle against the reference's positions after 3 steps."""
    g = np.load(os.path.join(G, "unstructured_steps.npz"))
    scene = (g["verts"], g["tets"], g["masses"], g["pins"])
    colors = colors_from(g, "ls1") if linsolver == 1 else None
    s = scenes.build_tet_scene(CpuSolver("oracle"), scene, 1, linsolver=linsolver, iters=8, colors=colors)
    s.set_x(g["x0"].ravel())
    for _ in range(3):
        s.step()
    assert np.abs(s.get_x() - g["ls%d_x3" % linsolver]).max() < 5e-7


def test_bunny_golden(cpu):
    """A real irregular mesh (the reference's samples/data/bunny_2250, 9 752 tets): StVK dropped on a Floor handled inside
    the Gauss-Seidel sweep with the reference's 16 colour lists, and Neo-Hookean with pins + LDLT."""
    g = np.load(os.path.join(G, "bunny_steps.npz"))
    scene = (g["verts"], g["tets"], g["masses"], g["pins"])
    s = scenes.build_tet_scene(CpuSolver("oracle"), scene, 2, linsolver=1, iters=8, floor=float(g["floor_y"][0]), pin=False, colors=colors_from(g, "floor"))
    for _ in range(5):
        s.step()
    assert np.abs(s.get_x() - g["floor_x5"]).max() < 5e-7
    s = scenes.build_tet_scene(CpuSolver("oracle"), scene, 1, linsolver=0, iters=8)
    for _ in range(3):
        s.step()
    assert np.abs(s.get_x() - g["ldlt_x3"]).max() < 5e-7
