"""Mesh ingestion: the package's .ele/.node loader, inverted-tet fix, lumped masses, surface vertices
and tiling against the reference's own mcl::meshio::load_elenode / TetMesh (through oracle/_ref), on the reference's sample
meshes.  The data files live in /root/reference (absent on the GPU box): these tests run in the build container only."""
import ctypes
import os

import numpy as np
import pytest

import checkers

DATA = ["/root/reference/samples/data/bunny_1124", "/root/reference/samples/data/bunny_closed", "/root/reference/samples/data/box768",
        "/root/reference/samples/data/torus", "/root/reference/deps/mclscene/src/data/armadillo_10k"]


def ref_mesh(prefix, density=1522.0):
    L = checkers.ref_lib()
    L.ref_mesh_load_elenode.restype = ctypes.c_void_p
    m = ctypes.c_void_p(L.ref_mesh_load_elenode(prefix.encode()))
    assert m
    nv, nt = L.ref_mesh_n_verts(m), L.ref_mesh_n_tets(m)
    verts, tets, masses = np.zeros((nv, 3), np.float32), np.zeros((nt, 4), np.int32), np.zeros(nv, np.float32)
    L.ref_mesh_get(m, verts.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), tets.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                   masses.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_float(density))
    ns = L.ref_mesh_surface_inds(m, None)
    surf = np.zeros(ns, np.int32)
    L.ref_mesh_surface_inds(m, surf.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    L.ref_mesh_free(m)
    return verts, tets, masses, surf


@pytest.mark.parametrize("prefix", DATA)
def test_loader_matches_reference_meshio(pkg, cpu, prefix):
    if not (os.path.exists(prefix + ".ele") and checkers.have_ref()):
        pytest.skip("reference sample data / oracle/_ref not present")
    rv, rt, rm, rs = ref_mesh(prefix)
    v, t = pkg.meshes.load_elenode(prefix)
    assert v.dtype == np.float32 and t.dtype == np.int32
    assert v.shape == rv.shape and (v == rv).all()                 # vertices: bit-exact floats
    assert t.shape == rt.shape and (t == rt).all()                 # indices incl. the inverted-tet reordering: bit-exact
    m = pkg.meshes.lumped_masses_tets(v, t, 1522.0)
    assert np.abs(m - rm).max() <= 2e-7 * np.abs(rm).max()         # float32 accumulation in the same (tet, corner) order
    assert (np.sort(rs) == pkg.meshes.surface_vertices(t)).all()   # same set; the reference's order is an unordered_map's
    # every element is positively oriented afterwards
    a = v[t[:, 0]].astype(np.float64)
    vol = np.einsum("ij,ij->i", v[t[:, 1]] - a, np.cross(v[t[:, 2]] - a, v[t[:, 3]] - a)) / 6.0
    assert (vol > -1e-12).all()


def test_inverted_tet_fix_and_tiling(pkg):
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
    good = np.array([[0, 1, 2, 3]], dtype=np.int32)
    bad = np.array([[0, 2, 1, 3]], dtype=np.int32)
    assert (pkg.meshes.fix_inverted_tets(v, good) == good).all()
    assert (pkg.meshes.fix_inverted_tets(v, bad) == good).all()
    vv, tt = pkg.meshes.tile_mesh(v, good, (2, 1, 3))
    assert vv.shape == (24, 3) and tt.shape == (6, 4)
    assert (tt[4] == good[0] + 16).all()
    # copies do not touch each other
    assert len(np.unique(np.round(vv, 5), axis=0)) == 24
    masses = pkg.meshes.lumped_masses_tets(vv, tt, 1000.0)
    assert np.allclose(masses, 1000.0 / 6.0 / 4.0)


def test_loader_rejects_bad_indices(pkg, tmp_path):
    p = str(tmp_path / "m")
    open(p + ".node", "w").write("4 3 0 0\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 0 0 1\n")
    open(p + ".ele", "w").write("1 4 0\n1 1 3 2 4\n")              # 1-based, inverted
    v, t = pkg.meshes.load_elenode(p)
    assert (t == [[0, 1, 2, 3]]).all() and v.shape == (4, 3)
    open(p + ".ele", "w").write("2 4 0\n1 1 2 3 4\n1 1 2 3 4\n")    # an id twice, one missing
    with pytest.raises(RuntimeError):
        pkg.meshes.load_elenode(p)
