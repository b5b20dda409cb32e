"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile).  Run in the build container (needs /root/reference):

    OMP_NUM_THREADS=1 python tests/golden/gen_golden.py

OMP_NUM_THREADS=1 makes the reference's graph colouring deterministic (seed 0*time,
GraphColor.hpp:156).  The fixtures travel with the repo; the reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import checkers  # noqa: E402
import conftest  # noqa: E402
import scenes  # noqa: E402
from checkers import CpuSolver  # noqa: E402

pkg = conftest.load_package()
MU, LAM = scenes.lame(*scenes.LAME_SOFT)


def prox_vectors():
    out = {}
    for model in range(6):
        zs = np.concatenate([checkers.random_F(200, s, seed=1234 + 10 * model + i) for i, s in enumerate((0.01, 0.1, 0.3))])
        if model in (1, 2):
            inv = checkers.random_F(40, 0.2, seed=99 + model)
            inv[:, 6:9] *= -1.0
            zs = np.concatenate([zs, inv])
        kappa = 0.0
        res, rc = checkers.prox_tets("ref", model, MU, LAM, zs, kappa)
        assert rc == 0
        out["tet%d_in" % model] = zs
        out["tet%d_out" % model] = res
    rng = np.random.RandomState(5)
    z = np.zeros((300, 6))
    z[:, 0] = 1.0
    z[:, 4] = 1.0
    z += 0.2 * rng.randn(300, 6)
    out["tri_in"] = z
    out["tri_out"] = checkers.prox_tris("ref", 100.0, 100.0, z)
    out["tri_lim_out"] = checkers.prox_tris("ref", 100.0, 100.0, z, 0.95, 1.05)
    out["mu_lambda"] = np.array([MU, LAM])
    np.savez_compressed(os.path.join(HERE, "prox_vectors.npz"), **out)


def beam_steps():
    """6x2x2 beam (120 tets), left face pinned, bent start, 3 steps x 10 ADMM iterations."""
    scene = scenes.beam(pkg.meshes, 6, 2, 2)
    out = {"verts": scene[0], "tets": scene[1], "masses": scene[2], "pins": scene[3], "x0": scenes.bend(scene[0])}
    for model, linsolver in ((0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (1, 2)):
        s = scenes.build_tet_scene(CpuSolver("ref"), scene, model, linsolver=linsolver, iters=10)
        s.set_x(out["x0"].ravel())
        key = "m%d_ls%d" % (model, linsolver)
        if linsolver == 1:
            colors = s.get_colors()
            out[key + "_color_off"] = np.cumsum([0] + [len(c) for c in colors]).astype(np.int32)
            out[key + "_color_nodes"] = np.concatenate(colors).astype(np.int32)
        z, u, b, x = s.traced_step(10)
        out[key + "_z_it0"] = z[0]
        out[key + "_u_it0"] = u[0]
        out[key + "_b_it0"] = b[0]
        out[key + "_x_it"] = x
        for _ in range(2):
            s.step()
        out[key + "_x3"] = s.get_x()
        out[key + "_v3"] = s.get_v()
    # floor inside Gauss-Seidel: free-falling StVK block
    floor_y = scene[0][:, 1].min() - 0.02
    s = scenes.build_tet_scene(CpuSolver("ref"), scene, 2, linsolver=1, iters=8, floor=floor_y, pin=False)
    colors = s.get_colors()
    out["floor_color_off"] = np.cumsum([0] + [len(c) for c in colors]).astype(np.int32)
    out["floor_color_nodes"] = np.concatenate(colors).astype(np.int32)
    out["floor_y"] = np.array([floor_y])
    for _ in range(6):
        s.step()
    out["floor_x6"] = s.get_x()
    np.savez_compressed(os.path.join(HERE, "beam_steps.npz"), **out)


def cloth_steps():
    v64, tris, masses, pins = scenes.cloth(pkg.meshes, 8)
    mu, lam = scenes.lame(100.0, 0.1)
    out = {"verts": v64, "tris": tris, "masses": masses, "pins": pins, "mu_lambda": np.array([mu, lam])}
    for linsolver in (0, 2):
        for name, limits in (("nolim", (-100.0, 100.0)), ("lim", (0.95, 1.05))):
            s = CpuSolver("ref")
            s.add_nodes(v64, masses)
            s.add_tris(v64, tris, mu, lam, *limits)
            s.set_pins(pins)
            assert s.initialize(dt=1.0 / 24, admm_iters=10, gravity=-9.8, linsolver=linsolver)
            for _ in range(4):
                s.step()
            out["ls%d_%s_x4" % (linsolver, name)] = s.get_x()
    np.savez_compressed(os.path.join(HERE, "cloth_steps.npz"), **out)


def uzawa_floor():
    """UzawaCG with passive hits (src/UzawaCG.hpp:57-125): the reference's own (x_in, b) -> x_out of every ADMM
    iteration of 4 steps of a free-falling 4x2x2 Neo-Hookean beam landing on a Floor.  Solve-level data: whole
    trajectories are chaotic in this scene (see tests/test_oracle_vs_ref.py)."""
    scene = scenes.beam(pkg.meshes, 4, 2, 2)
    floor_y = scene[0][:, 1].min() - 0.02
    iters = 8
    s = scenes.build_tet_scene(CpuSolver("ref"), scene, 1, linsolver=2, iters=iters, floor=floor_y, pin=False)
    x_in, bs, x_out = [], [], []
    for step in range(4):
        x_prev = s.get_x() + (1.0 / 24) * (s.get_v() + np.tile([0, (1.0 / 24) * -9.8, 0], s.dof // 3))
        z, u, b, x = s.traced_step(iters)
        for it in range(iters):
            x_in.append(x_prev if it == 0 else x[it - 1])
            bs.append(b[it])
            x_out.append(x[it])
    x_in, bs, x_out = np.array(x_in), np.array(bs), np.array(x_out)
    hits = (x_in.reshape(len(x_in), -1, 3)[:, :, 1] < floor_y).sum(axis=1)
    assert (hits > 0).sum() > 5
    np.savez_compressed(os.path.join(HERE, "uzawa_floor.npz"), verts=scene[0], tets=scene[1], masses=scene[2], floor_y=np.array([floor_y]),
                        iters=np.array([iters]), x_in=x_in, b=bs, x_out=x_out, hits=hits)


def unstructured_steps():
    """Delaunay blob (tests/scenes.py: blob), Neo-Hookean, 3 steps x 8 ADMM iterations with LDLT and with the
    multi-colour Gauss-Seidel (the reference's colours are stored)."""
    scene = scenes.blob(pkg.meshes)
    out = {"verts": scene[0], "tets": scene[1], "masses": scene[2], "pins": scene[3], "x0": scenes.bend(scene[0], 0.08)}
    for linsolver in (0, 1):
        s = scenes.build_tet_scene(CpuSolver("ref"), scene, 1, linsolver=linsolver, iters=8)
        s.set_x(out["x0"].ravel())
        if linsolver == 1:
            colors = s.get_colors()
            out["ls1_color_off"] = np.cumsum([0] + [len(c) for c in colors]).astype(np.int32)
            out["ls1_color_nodes"] = np.concatenate(colors).astype(np.int32)
        for _ in range(3):
            s.step()
        out["ls%d_x3" % linsolver] = s.get_x()
    np.savez_compressed(os.path.join(HERE, "unstructured_steps.npz"), **out)


def single_tet():
    """test_lineartet.cpp known answers as produced by the reference here."""
    V = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    T = np.array([[0, 1, 2, 3]], dtype=np.int32)
    mu, lam = scenes.lame(500000, 0.25)
    its = [5, 10, 21, 50, 99]
    xs = []
    for it in its:
        s = CpuSolver("ref")
        s.add_nodes(V, np.ones(12))
        s.add_tets(V, T, 0, mu, lam)
        assert s.initialize(dt=float(np.float32(1) / np.float32(24)), admm_iters=it, gravity=0.0, linsolver=0)
        x = V.ravel().copy()
        x[9:12] = [200, 0, 0]
        s.set_x(x)
        s.step()
        xs.append(s.get_x())
    np.savez_compressed(os.path.join(HERE, "single_tet.npz"), iters=np.array(its), x=np.array(xs))


def bunny_steps():
    """A REAL irregular mesh from the reference's sample data (samples/data/bunny_2250: 9 752 tets), read by the package's
    own .ele/.node loader (checked against the reference's MeshIO in tests/test_meshes_cpu.py):
      * StVK, dropped on a Floor handled inside the multi-colour Gauss-Seidel (BASELINE config 3 style, the reference's own
        colour lists are stored), 5 steps x 8 ADMM iterations;
      * Neo-Hookean, the topmost vertices pinned, LDLT, 3 steps x 8 ADMM iterations."""
    prefix = "/root/reference/samples/data/bunny_2250"
    verts, tets = pkg.meshes.load_elenode(prefix)
    masses = pkg.meshes.lumped_masses_tets(verts, tets, 1522.0).astype(np.float64)
    v64 = verts.astype(np.float64)
    pins = np.nonzero(v64[:, 1] > v64[:, 1].max() - 0.08 * (v64[:, 1].max() - v64[:, 1].min()))[0].astype(np.int32)
    floor_y = float(v64[:, 1].min() - 0.01)
    out = {"verts": v64, "tets": tets, "masses": masses, "pins": pins, "floor_y": np.array([floor_y])}
    scene = (v64, tets, masses, pins)
    s = scenes.build_tet_scene(CpuSolver("ref"), scene, 2, linsolver=1, iters=8, floor=floor_y, pin=False)
    colors = s.get_colors()
    out["floor_color_off"] = np.cumsum([0] + [len(c) for c in colors]).astype(np.int32)
    out["floor_color_nodes"] = np.concatenate(colors).astype(np.int32)
    for _ in range(5):   # it bounces: in contact after steps 1, 3 and 5
        s.step()
    out["floor_x5"] = s.get_x()
    assert (np.abs(out["floor_x5"].reshape(-1, 3)[:, 1] - floor_y) < 1e-12).any()
    s = scenes.build_tet_scene(CpuSolver("ref"), scene, 1, linsolver=0, iters=8)
    for _ in range(3):
        s.step()
    out["ldlt_x3"] = s.get_x()
    np.savez_compressed(os.path.join(HERE, "bunny_steps.npz"), **out)


if __name__ == "__main__":
    assert checkers.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    prox_vectors()
    beam_steps()
    cloth_steps()
    single_tet()
    uzawa_floor()
    unstructured_steps()
    bunny_steps()
    print("golden fixtures written to", HERE)
