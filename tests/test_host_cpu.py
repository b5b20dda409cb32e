"""Host-side logic and the C-ABI surface, without a GPU: the libraries load, export every symbol
include/admm_b200.h declares, the host colouring / ordering / factorisation are correct, and the
compute entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "admm_b200.h")).read()
    names = sorted(set(re.findall(r"\b(admm_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = ctypes.CDLL(pkg.LIB_CUDA_PATH)
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_no_cpu_fallback_without_device(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.AdmmError) as e:
        pkg.DeviceSolver(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)
    s = pkg.Solver()
    V = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    s.add_nodes(V, np.ones(4))
    s.add_tets(V, np.array([[0, 1, 2, 3]], np.int32), pkg.TET_LINEAR, 1.0, 1.0)
    with pytest.raises(pkg.AdmmError):
        s.initialize(linsolver=0)


def test_initialize_rejects_bad_node_data(pkg):
    # "**Solver Error: Problem with node data!" -> false (src/Solver.cpp:180-183)
    s = pkg.Solver()
    assert s.initialize(linsolver=0) is False


def test_inverted_rest_tet_throws(pkg):
    s = pkg.Solver()
    V = np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    s.add_nodes(V, np.ones(4))
    with pytest.raises(pkg.AdmmError) as e:
        s.add_tets(V, np.array([[0, 2, 1, 3]], np.int32), pkg.TET_LINEAR, 1.0, 1.0)
    assert "Inverted initial tet" in str(e.value)


def _beam_matrix(pkg, nx=6, ny=3, nz=3):
    verts, tets = pkg.meshes.make_tet_blocks(nx, ny, nz)
    n = len(verts)
    rows = np.repeat(tets, 4, axis=1).ravel()
    cols = np.tile(tets, (1, 4)).ravel()
    A = sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    L = sp.csr_matrix(-A)
    L.setdiag(0)
    L = L + sp.diags(np.asarray(abs(L).sum(axis=1)).ravel() + 1.0)
    L = sp.csr_matrix(L)
    L.sort_indices()
    return verts.astype(np.float64), tets, L


@pytest.mark.parametrize("method", [0, 1])
def test_coloring_is_valid(pkg, method):
    # validity checks of deps/mclscene/src/tests/test_graphcolor.cpp:75-130: no two neighbours share
    # a colour, no two vertices of a tet share a colour, every node coloured exactly once
    verts, tets, L = _beam_matrix(pkg)
    colors = pkg.color_matrix(L.indptr, L.indices, L.data, method)
    color_of = np.full(L.shape[0], -1)
    for c, nodes in enumerate(colors):
        assert len(nodes) > 0
        assert (color_of[nodes] == -1).all()
        color_of[nodes] = c
    assert (color_of >= 0).all()
    coo = L.tocoo()
    off = coo.row != coo.col
    assert (color_of[coo.row[off]] != color_of[coo.col[off]]).all()
    for t in tets:
        assert len(set(color_of[t])) == 4
    assert len(colors) <= 24


def test_mesh_generator_matches_reference_pattern(pkg):
    verts, tets = pkg.meshes.make_tet_blocks(3, 2, 2)
    assert len(tets) == 3 * 2 * 2 * 5 and len(verts) == 4 * 3 * 3
    v = verts.astype(np.float64)
    e = np.stack([v[tets[:, 1]] - v[tets[:, 0]], v[tets[:, 2]] - v[tets[:, 0]], v[tets[:, 3]] - v[tets[:, 0]]], axis=2)
    vol = np.linalg.det(e) / 6.0
    assert (vol > 0).all()  # no inverted rest tets
    assert abs(vol.sum() - np.prod(v.max(0) - v.min(0))) < 1e-5  # the 5 tets tile each cube
    assert abs((v[:, 1].max() - v[:, 1].min()) - 1.0) < 1e-6  # 1 m tall (beams.cpp:61-68)
    m = pkg.meshes.lumped_masses_tets(verts, tets)
    assert abs(m.sum() - 1522.0 * vol.sum()) < 1e-2


def test_nested_dissection_ldlt_solves(pkg):
    verts, tets, L = _beam_matrix(pkg, 8, 3, 3)
    rng = np.random.RandomState(0)
    b = rng.randn(L.shape[0])
    x, nnz = pkg.ldlt_solve_host(L.indptr, L.indices, L.data, verts, b)
    assert np.abs(L @ x - b).max() < 1e-10
    assert nnz > 0


def test_reference_harness_has_no_product_dependency(pkg):
    # the product libraries must not link against the checkers
    import subprocess
    for lib in (pkg.LIB_CUDA_PATH, pkg.LIB_HOST_PATH):
        out = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
        assert "liboracle" not in out and "libadmm_ref" not in out


@pytest.mark.parametrize("dims,parts", [((6, 2, 2), 148), ((20, 6, 5), 148), ((3, 1, 1), 148), ((12, 4, 4), 7)])
def test_resident_gauss_seidel_plan(pkg, dims, parts):
    """csrc/partition.hpp: the shared-memory-resident MCGS plan covers every node exactly once,
    colour by colour, and reproduces L_offdiag * x (walked on the host exactly like the kernel's gather)."""
    import ctypes
    import scipy.sparse as sp
    verts, tets = pkg.meshes.make_tet_blocks(*dims)
    n = len(verts)
    rows, cols = np.repeat(tets, 4, axis=1).ravel(), np.tile(tets, (1, 4)).ravel()
    A = sp.csr_matrix((np.random.RandomState(0).rand(rows.size) + 0.1, (rows, cols)), shape=(n, n))
    A = (A + A.T).tocsr()
    A.sort_indices()
    colors = pkg.color_matrix(A.indptr, A.indices, A.data, 0)
    off = np.zeros(len(colors) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in colors])
    nodes = np.concatenate(colors).astype(np.int32)
    x = np.random.RandomState(1).randn(n)
    pos = np.ascontiguousarray(verts.astype(np.float64))
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    for val_bytes, lanes, tol in ((8, 1, 1e-12), (4, 1, 1e-12), (8, 2, 1e-12), (4, 4, 1e-12)):
        err, stats, part = ctypes.c_double(0), (ctypes.c_longlong * 6)(), np.zeros(n, np.int32)
        rc = pkg.cuda_lib.admm_b200_plan_check(n, ip(rp), ip(ci), dp(va), len(colors), ip(off), ip(nodes), dp(pos), parts, val_bytes, lanes,
                                               dp(x), ctypes.byref(err), stats, ip(part))
        assert rc == 0, pkg.cuda_lib.admm_b200_last_error(None)
        assert err.value < tol
        assert stats[4] == A.nnz - n          # every off-diagonal entry placed exactly once
        assert 0 <= part.min() and part.max() < parts
        sizes = np.bincount(part, minlength=parts)
        if n >= 4 * parts:
            assert sizes.max() <= 1.5 * n / parts + 8   # balanced


@pytest.mark.parametrize("dims,parts,warps", [((6, 2, 2), 5, 4), ((10, 4, 3), 12, 16), ((8, 3, 3), 148, 16)])
def test_barrier_free_schedule_model(pkg, dims, parts, warps):
    """csrc/dataflow_plan.hpp: per-slice dependencies + tagged mailboxes are enough -- random legal schedules of the
    barrier-free Gauss-Seidel reproduce colour-by-colour sweeps bit for bit and never deadlock (host model)."""
    import ctypes
    import scipy.sparse as sp
    verts, tets = pkg.meshes.make_tet_blocks(*dims)
    n = len(verts)
    rows, cols = np.repeat(tets, 4, axis=1).ravel(), np.tile(tets, (1, 4)).ravel()
    A = sp.csr_matrix((np.random.RandomState(0).rand(rows.size) + 0.1, (rows, cols)), shape=(n, n))
    A = (A + A.T).tolil()
    A.setdiag(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)   # diagonally dominant: the sweeps stay bounded
    A = A.tocsr()
    A.sort_indices()
    colors = pkg.color_matrix(A.indptr, A.indices, A.data, 0)
    off = np.zeros(len(colors) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in colors])
    nodes = np.concatenate(colors).astype(np.int32)
    pos = np.ascontiguousarray(verts.astype(np.float64))
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    stats = (ctypes.c_longlong * 4)()
    rc = pkg.cuda_lib.admm_b200_dataflow_check(n, ip(rp), ip(ci), dp(va), len(colors), ip(off), ip(nodes), dp(pos), parts, warps, 4, 3, stats)
    assert rc == 0, pkg.cuda_lib.admm_b200_last_error(None)
    assert stats[2] > 0 and stats[0] >= 1


def test_host_code_on_an_unstructured_mesh(pkg):
    """The host pieces that only ever saw block beams and cloth -- colouring, nested-dissection LDL^T, resident plan,
    mailboxes -- on a Delaunay mesh with irregular valence (tests/scenes.py: blob)."""
    import ctypes
    import scipy.sparse as sp
    import scenes
    verts, tets, masses, pins = scenes.blob(pkg.meshes)
    n = len(verts)
    rows, cols = np.repeat(tets, 4, axis=1).ravel(), np.tile(tets, (1, 4)).ravel()
    A = sp.csr_matrix((np.random.RandomState(0).rand(rows.size) + 0.1, (rows, cols)), shape=(n, n))
    A = (A + A.T).tolil()
    A.setdiag(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)
    A = A.tocsr()
    A.sort_indices()
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), np.ascontiguousarray(A.data)
    # colouring: valid, every node once
    for method in (0, 1):
        colors = pkg.color_matrix(rp, ci, va, method)
        color_of = -np.ones(n, int)
        for c, l in enumerate(colors):
            assert (color_of[l] == -1).all()
            color_of[l] = c
        coo = A.tocoo()
        off = coo.row != coo.col
        assert (color_of >= 0).all() and (color_of[coo.row[off]] != color_of[coo.col[off]]).all()
    # LDL^T
    b = np.random.RandomState(1).randn(n)
    x, stats = pkg.ldlt_solve_host(rp, ci, va, np.ascontiguousarray(verts), b)
    assert np.abs(A @ x - b).max() < 1e-9 * np.abs(b).max()
    # resident plan + mailboxes + split-free walk, several part counts and lane widths
    off = np.zeros(len(colors) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in colors])
    nodes = np.concatenate(colors).astype(np.int32)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    xx = np.random.RandomState(2).randn(n)
    pos = np.ascontiguousarray(verts)
    for parts, lanes in ((148, 1), (9, 1), (148, 4)):
        err, st, part = ctypes.c_double(0), (ctypes.c_longlong * 6)(), np.zeros(n, np.int32)
        rc = pkg.cuda_lib.admm_b200_plan_check(n, ip(rp), ip(ci), dp(va), len(colors), ip(off), ip(nodes), dp(pos), parts, 4, lanes,
                                               dp(xx), ctypes.byref(err), st, ip(part))
        assert rc == 0, pkg.cuda_lib.admm_b200_last_error(None)
        assert err.value < 1e-12 and st[4] == A.nnz - n
    stats4 = (ctypes.c_longlong * 4)()
    assert pkg.cuda_lib.admm_b200_dataflow_check(n, ip(rp), ip(ci), dp(va), len(colors), ip(off), ip(nodes), dp(pos), 9, 4, 3, 2, stats4) == 0


def test_ldlt_block_plan_host(pkg):
    """The block (supernodal) plan of the device's L D L^T solve (csrc/ldlt_blocks.hpp), walked on the host exactly as
    sptrsv_blocks.cuh walks it: solves A x = b to rounding, and the block tree is shallow (a level-scheduled solve of
    the same factors needs one grid barrier per separator COLUMN: hundreds to thousands of levels)."""
    import scipy.sparse as sp

    def laplacian(n, elems):
        k = elems.shape[1]
        rows = np.repeat(elems, k, axis=1).ravel()
        cols = np.tile(elems, (1, k)).ravel()
        A = sp.coo_matrix((-np.ones(rows.size), (rows, cols)), shape=(n, n)).tocsr()
        A.setdiag(0)
        A.eliminate_zeros()
        A = (A + sp.diags(-np.asarray(A.sum(1)).ravel() + 0.5)).tocsr()
        A.sort_indices()
        return A

    for verts, elems, max_levels in ((pkg.meshes.make_tet_blocks(24, 8, 6) + (16,)), (pkg.meshes.make_plane_sym(48, 40) + (24,))):
        A = laplacian(len(verts), elems)
        b = np.random.RandomState(1).randn(len(verts))
        x, st = pkg.ldlt_blocks_solve_host(A.indptr, A.indices, A.data, verts.astype(np.float64), b)
        assert np.abs(A @ x - b).max() < 1e-11 * np.abs(b).max()
        x_ref, nnz = pkg.ldlt_solve_host(A.indptr, A.indices, A.data, verts.astype(np.float64), b)
        assert np.abs(x - x_ref).max() < 1e-11 * np.abs(x_ref).max()
        assert st["levels_f"] <= max_levels and st["levels_b"] <= max_levels, st
        assert st["nnz_out"] + st["nnz_inv"] >= nnz          # explicit zeros inside the diagonal blocks only ever add entries
    # a path graph: one long chain of the elimination tree -- blocks are capped, the plan stays correct
    n = 6000
    A = sp.diags([-np.ones(n - 1), 2.5 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).tocsr()
    pos = np.zeros((n, 3))
    pos[:, 0] = np.arange(n)
    b = np.random.RandomState(2).randn(n)
    x, st = pkg.ldlt_blocks_solve_host(A.indptr, A.indices, A.data, pos, b)
    assert np.abs(A @ x - b).max() < 1e-11
