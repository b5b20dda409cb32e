"""Shared scene builders for the parity tests: every solver kind (oracle / ref / gpu) gets the SAME
arrays through the same calls."""
import numpy as np

LAME_SOFT = (1e7, 0.399)  # admm::Lame softRubber, samples/sca2016/beams.cpp:87


def lame(E, nu):
    return E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))


def beam(meshes, nx=6, ny=2, nz=2):
    verts, tets = meshes.make_tet_blocks(nx, ny, nz)
    masses = meshes.lumped_masses_tets(verts, tets)
    v64 = verts.astype(np.float64)
    pins = np.nonzero(v64[:, 0] < v64[:, 0].min() + 1e-2)[0].astype(np.int32)
    return v64, tets, masses.astype(np.float64), pins


def cloth(meshes, n=8):
    verts, tris = meshes.make_plane(n, n)
    masses = meshes.lumped_masses_tris(verts, tris)
    v64 = verts.astype(np.float64)
    pins = np.array([0, n], dtype=np.int32)  # two corners of one edge
    return v64, tris, masses.astype(np.float64), pins


def blob(meshes, n_points=160, seed=3):
    """An UNSTRUCTURED tet mesh: Delaunay tetrahedralisation of random points in a 1.6 x 1 x 1 box (irregular valence,
    badly shaped elements near the hull, more colours than a block beam), slivers below 2 % of the mean volume
    dropped, all elements oriented positively.  Nodes with x below the 12 % quantile are pinned."""
    from scipy.spatial import Delaunay
    rng = np.random.RandomState(seed)
    pts = rng.rand(n_points, 3) * np.array([1.6, 1.0, 1.0])
    tets = Delaunay(pts).simplices.astype(np.int32)
    p = pts[tets]
    vol = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0]) / 6.0
    flip = vol < 0
    tets[flip] = tets[flip][:, [0, 2, 1, 3]]
    vol = np.abs(vol)
    tets = np.ascontiguousarray(tets[vol > 0.02 * vol.mean()])
    used = np.unique(tets)
    remap = -np.ones(n_points, dtype=np.int64)
    remap[used] = np.arange(len(used))
    verts = np.ascontiguousarray(pts[used].astype(np.float32))   # the reference's meshes are float (AddMeshes.hpp:120)
    tets = remap[tets].astype(np.int32)
    masses = meshes.lumped_masses_tets(verts, tets)
    v64 = verts.astype(np.float64)
    pins = np.nonzero(v64[:, 0] < np.quantile(v64[:, 0], 0.12))[0].astype(np.int32)
    return v64, tets, masses.astype(np.float64), pins


def bend(x, amount=0.05):
    """A smooth non-rigid deformation so that the prox runs away from F = I."""
    x = x.copy()
    x[:, 1] += amount * (x[:, 0] - x[:, 0].min()) ** 2
    x[:, 2] += 0.5 * amount * np.sin(2.0 * x[:, 0])
    return x


def build_tet_scene(solver, scene, model, lame_pair=None, linsolver=0, dt=1.0 / 24, iters=10, gravity=-9.8,
                    floor=None, sphere=None, colors=None, pin=True, constraint_w=-1.0):
    v64, tets, masses, pins = scene
    mu, lam = lame_pair if lame_pair is not None else lame(*LAME_SOFT)
    solver.add_nodes(v64, masses)
    solver.add_tets(v64, tets, model, mu, lam)
    if pin:
        solver.set_pins(pins)
    if floor is not None:
        solver.add_floor(floor)
    if sphere is not None:
        solver.add_sphere(sphere[0], sphere[1])
    if colors is not None:
        solver.set_colors(colors)
    ok = solver.initialize(dt=dt, admm_iters=iters, gravity=gravity, linsolver=linsolver, constraint_w=constraint_w) \
        if "constraint_w" in solver.initialize.__code__.co_varnames else solver.initialize(dt=dt, admm_iters=iters, gravity=gravity, linsolver=linsolver)
    assert ok
    return solver
