"""Synthetic meshes for the ADMM-elastic benchmarks and tests (the caller side of the boundary).

make_tet_blocks reproduces the element pattern of the reference's mcl::factory::make_tet_blocks
(deps/mclscene/include/MCL/ShapeFactory.hpp:296-357: 5 tets per unit cube, corners a..h, tets
(0,5,7,4) (5,7,2,0) (5,0,2,1) (7,2,0,3) (5,2,7,6)) but numbers the shared vertices directly on the
grid instead of calling the reference's O(n_tets * n_verts) TetMesh::refine, so
1M / 8M element beams build in seconds.  lumped_masses follows TetMesh::weighted_masses
(deps/mclscene/include/MCL/TetMesh.hpp:297-315) and TriangleMesh::weighted_masses
(TriangleMesh.hpp:281-296) in float32 like the reference.
"""
import numpy as np


def make_tet_blocks(nx, ny, nz, height=1.0):
    """nx*ny*nz unit cubes split into 5 tets each, scaled so the y extent is `height` metres and
    centred at the origin (samples/sca2016/beams.cpp:61-68).  Returns float32 verts [n,3] (the
    reference meshes are float) and int32 tets [m,4]."""
    nx, ny, nz = max(1, nx), max(1, ny), max(1, nz)
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    verts = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1).astype(np.float32)

    def vid(x, y, z):
        return (x * (ny + 1) + y) * (nz + 1) + z

    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cx, cy, cz = cx.ravel(), cy.ravel(), cz.ravel()
    # corners a..h as in ShapeFactory.hpp:313-323 (min = (x,y,z), max = min + 1)
    corner = [
        vid(cx + 1, cy + 1, cz + 1),  # a = max
        vid(cx, cy + 1, cz + 1),      # b
        vid(cx, cy + 1, cz),          # c
        vid(cx + 1, cy + 1, cz),      # d
        vid(cx + 1, cy, cz + 1),      # e
        vid(cx, cy, cz + 1),          # f
        vid(cx, cy, cz),              # g
        vid(cx + 1, cy, cz),          # h
    ]
    pattern = [(0, 5, 7, 4), (5, 7, 2, 0), (5, 0, 2, 1), (7, 2, 0, 3), (5, 2, 7, 6)]
    tets = np.empty((cx.size, 5, 4), dtype=np.int32)
    for t, p in enumerate(pattern):
        for k in range(4):
            tets[:, t, k] = corner[p[k]]
    tets = tets.reshape(-1, 4)
    lo, hi = verts.min(0), verts.max(0)
    centre = np.float32(0.5) * (lo + hi)
    scale = np.float32(height) / np.float32(hi[1] - lo[1])
    verts = ((verts - centre) * scale).astype(np.float32)
    return verts, tets


def make_plane(nx, ny, width=2.0):
    """Regular triangle sheet in the xz-plane, (nx+1)*(ny+1) verts, 2*nx*ny triangles."""
    gx, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    verts = np.stack([gx.ravel(), np.zeros(gx.size), gz.ravel()], axis=1).astype(np.float32)
    verts[:, 0] = (verts[:, 0] / nx - 0.5) * width
    verts[:, 2] = (verts[:, 2] / ny - 0.5) * width

    def vid(i, j):
        return i * (ny + 1) + j

    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    ci, cj = ci.ravel(), cj.ravel()
    t0 = np.stack([vid(ci, cj), vid(ci, cj + 1), vid(ci + 1, cj)], axis=1)
    t1 = np.stack([vid(ci + 1, cj), vid(ci, cj + 1), vid(ci + 1, cj + 1)], axis=1)
    tris = np.concatenate([t0, t1], axis=0).astype(np.int32)
    return verts.astype(np.float32), tris


def lumped_masses_tets(verts, tets, density=1522.0):
    """float32 lumped vertex masses, a quarter of each tet's mass per corner."""
    v = verts.astype(np.float32)
    e1, e2, e3 = v[tets[:, 1]] - v[tets[:, 0]], v[tets[:, 2]] - v[tets[:, 0]], v[tets[:, 3]] - v[tets[:, 0]]
    vol = np.abs(np.einsum("ij,ij->i", e1, np.cross(e2, e3)).astype(np.float32) / np.float32(6.0))
    tm = (np.float32(density) * vol / np.float32(4.0)).astype(np.float32)
    m = np.zeros(len(v), dtype=np.float32)
    for c in range(4):
        np.add.at(m, tets[:, c], tm)
    return m


def lumped_masses_tris(verts, tris, density=1.0):
    v = verts.astype(np.float32)
    e1, e2 = v[tris[:, 1]] - v[tris[:, 0]], v[tris[:, 2]] - v[tris[:, 0]]
    area = (np.float32(0.5) * np.linalg.norm(np.cross(e1, e2), axis=1)).astype(np.float32)
    tm = (np.float32(density) * area / np.float32(3.0)).astype(np.float32)
    m = np.zeros(len(v), dtype=np.float32)
    for c in range(3):
        np.add.at(m, tris[:, c], tm)
    return m


def lame(youngs, poisson):
    """admm::Lame(k, v) (src/EnergyTerm.hpp:50-54) -> (mu, lambda)."""
    mu = youngs / (2.0 * (1.0 + poisson))
    lam = youngs * poisson / ((1.0 + poisson) * (1.0 - 2.0 * poisson))
    return mu, lam
