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


def make_plane_sym(tess_x, tess_y):
    """mcl::factory::make_plane (deps/mclscene/include/MCL/ShapeFactory.hpp:424-484): a [-1,1]^2 sheet in the xy-plane,
    (tess_x+1)(tess_y+1) grid vertices followed by one centre vertex per cell, 4 triangles per cell in mkquad_sym's
    order (ll,lr,c) (lr,ur,c) (c,ur,ul) (ll,c,ul).  512 x 512 gives BASELINE config 4's 525 313 vertices / 1 048 576
    triangles.  float32 arithmetic like the reference."""
    tx, ty = max(1, tess_x), max(1, tess_y)
    gx, gy = np.meshgrid(np.arange(tx + 1), np.arange(ty + 1), indexing="ij")
    f = np.float32
    grid = np.stack([f(-1.0) + f(2.0) * gx.ravel().astype(f) / f(tx), f(-1.0) + f(2.0) * gy.ravel().astype(f) / f(ty),
                     np.zeros(gx.size, f)], axis=1)
    cx, cy = np.meshgrid(np.arange(tx), np.arange(ty), indexing="ij")
    cx, cy = cx.ravel(), cy.ravel()
    cen = np.stack([f(-1.0) + f(2.0) * cx.astype(f) / f(tx) + f(1.0) / f(tx), f(-1.0) + f(2.0) * cy.astype(f) / f(ty) + f(1.0) / f(ty),
                    np.zeros(cx.size, f)], axis=1)
    verts = np.concatenate([grid, cen], axis=0).astype(np.float32)
    ll = cy + cx * (ty + 1)
    lr = cy + (cx + 1) * (ty + 1)
    ul, ur = ll + 1, lr + 1
    c = (tx + 1) * (ty + 1) + cx * ty + cy
    quads = np.stack([np.stack([ll, lr, c], 1), np.stack([lr, ur, c], 1), np.stack([c, ur, ul], 1), np.stack([ll, c, ul], 1)], axis=1)
    return verts, quads.reshape(-1, 3).astype(np.int32)


def cloth_corner_pins(verts):
    """get_pins of samples/sca2016/trianglestrain.cpp:103-135: among the vertices of the top edge (y >= max_y - 1e-3) the
    one with the smallest and the one with the largest x, scanned in index order with the sample's if / else-if."""
    v = np.asarray(verts, dtype=np.float32)
    top = np.nonzero(~(v[:, 1] < v[:, 1].max() - np.float32(1e-3)))[0]
    left = right = -1
    mn, mx = np.float32(99999.0), np.float32(-99999.0)
    for i in top:
        x = v[i, 0]
        if x < mn:
            left, mn = int(i), x
        elif x > mx:
            right, mx = int(i), x
    if left < 0 or right < 0:
        raise RuntimeError("Failed to find pin locations")
    return np.array([left, right], dtype=np.int32)


def lumped_masses_tets(verts, tets, density=1522.0):
    """float32 lumped vertex masses, a quarter of each tet's mass per corner, accumulated tet by tet and corner by corner
    like TetMesh::weighted_masses (deps/mclscene/include/MCL/TetMesh.hpp:297-315)."""
    v = verts.astype(np.float32)
    e1, e2, e3 = v[tets[:, 1]] - v[tets[:, 0]], v[tets[:, 2]] - v[tets[:, 0]], v[tets[:, 3]] - v[tets[:, 0]]
    vol = np.abs(np.einsum("ij,ij->i", e1, np.cross(e2, e3)).astype(np.float32) / np.float32(6.0))
    tm = ((np.float32(density) * vol).astype(np.float32) / np.float32(4.0)).astype(np.float32)
    m = np.zeros(len(v), dtype=np.float32)
    np.add.at(m, np.asarray(tets).ravel(), np.repeat(tm, 4))   # unbuffered, in (tet, corner) order
    return m


def load_elenode(prefix):
    """TetGen .ele / .node pair -> (float32 verts [n,3], int32 tets [m,4]) with the semantics of mcl::meshio::load_elenode
    (deps/mclscene/include/MCL/MeshIO.hpp:180-311): the first line holds the count; every further line is `id a b c d` /
    `id x y z`; ids are 1-based when the first one is 1; a record is stored AT its id; every id must occur; vertices are
    rounded to float; tets with negative float32 volume get their corners 1 and 2 swapped."""
    def read(path, n_cols, dtype):
        with open(path) as f:
            header = f.readline().split()
            n = int(header[0]) if header else 0
            rows = []
            for _ in range(n):
                rows.append(f.readline().split()[:1 + n_cols])
        ids = np.array([int(r[0]) for r in rows], dtype=np.int64)
        vals = np.array([[float(t) for t in r[1:1 + n_cols]] for r in rows], dtype=np.float64).reshape(n, n_cols)
        one = n > 0 and ids[0] == 1
        if one:
            ids = ids - 1
        if n == 0 or ids.min() < 0 or ids.max() >= n or len(np.unique(ids)) != n:
            raise RuntimeError("**TetMesh Error: Your indices are bad for file %s" % path)
        out = np.zeros((n, n_cols), dtype=np.float64)
        out[ids] = vals
        return out.astype(dtype), one
    tets, one = read(prefix + ".ele", 4, np.int64)
    if one:
        tets = tets - 1
    verts, _ = read(prefix + ".node", 3, np.float32)
    if len(verts) == 0 or len(tets) == 0:
        raise RuntimeError("**TetMesh Error: Problem loading files")
    tets = fix_inverted_tets(verts, tets.astype(np.int32))
    return verts, tets


def fix_inverted_tets(verts, tets):
    """Swaps corners 1 and 2 of every tet whose float32 signed volume is negative (MeshIO.hpp:291-303)."""
    v = np.asarray(verts, dtype=np.float32)
    t = np.array(tets, dtype=np.int32, copy=True)
    a = v[t[:, 0]]
    vol = (np.einsum("ij,ij->i", v[t[:, 1]] - a, np.cross(v[t[:, 2]] - a, v[t[:, 3]] - a)).astype(np.float32) / np.float32(6.0))
    flip = vol < 0
    t[flip, 1], t[flip, 2] = tets[flip, 2], tets[flip, 1]
    return t


def tile_mesh(verts, tets, counts, gap=0.05):
    """counts = (cx, cy, cz) translated copies of a mesh on a grid (spacing = bounding box + gap x its largest side), vertex
    ids offset copy by copy -- how a 50k-tet bunny becomes a 1M- or 8M-tet scene (20 / 146 tiled copies)."""
    v = np.asarray(verts, dtype=np.float32)
    t = np.asarray(tets, dtype=np.int32)
    ext = v.max(0) - v.min(0)
    pitch = (ext + np.float32(gap) * ext.max()).astype(np.float32)
    vs, ts = [], []
    k = 0
    for ix in range(counts[0]):
        for iy in range(counts[1]):
            for iz in range(counts[2]):
                vs.append((v + pitch * np.array([ix, iy, iz], dtype=np.float32)).astype(np.float32))
                ts.append(t + np.int32(k * len(v)))
                k += 1
    return np.concatenate(vs, axis=0), np.concatenate(ts, axis=0).astype(np.int32)


def surface_vertices(tets):
    """Vertices of the boundary faces (faces that belong to exactly one tet), ascending: the set TetMesh::surface_inds
    returns (deps/mclscene/include/MCL/TetMesh.hpp:317-342; the reference's ORDER is that of an unordered_map)."""
    t = np.asarray(tets, dtype=np.int64)
    faces = np.concatenate([t[:, [0, 1, 2]], t[:, [0, 1, 3]], t[:, [0, 2, 3]], t[:, [1, 2, 3]]], axis=0)
    key = np.sort(faces, axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    boundary = faces[cnt[inv.ravel()] == 1]
    return np.unique(boundary).astype(np.int32)


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
