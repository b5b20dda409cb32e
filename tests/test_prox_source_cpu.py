"""The device prox SOURCE (admm-elastic_b200/csrc/prox.cuh) recompiled for the host by
tests/tools/prox_host.cu, against the golden vectors and the C oracle.  No GPU needed.

This is a development aid: it catches algorithmic parity problems (basins of the Newton iteration,
the hand-over to the reference-faithful L-BFGS path for degenerate elements) in the container.  It is
not a parity claim for the CUDA path -- that is tests/test_gpu_parity.py on the B200 -- and nothing
in the package loads the shim library.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import checkers
import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tools", "prox_host.cu")
OUT = os.path.join(HERE, "tools", "_build", "libprox_host.so")
MU, LAM = scenes.lame(*scenes.LAME_SOFT)
D = ctypes.c_double


@pytest.fixture(scope="module")
def shim(cpu):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    dep = os.path.join(os.path.dirname(HERE), "admm-elastic_b200", "csrc", "prox.cuh")
    stale = not os.path.exists(OUT) or any(os.path.getmtime(p) > os.path.getmtime(OUT) for p in (SRC, dep))
    if stale:
        if not os.path.exists(nvcc):
            pytest.skip("nvcc not available to build the host shim")
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC])
    L = ctypes.CDLL(OUT)

    def prox_tets(model, mu, lam, z, kappa=0.0, precision=0):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 9)
        out = np.empty_like(z)
        L.shim_prox_tets(int(model), D(mu), D(lam), D(kappa), int(precision), z.shape[0], checkers.dp(z), checkers.dp(out))
        return out

    def prox_tris(z, lmin=-100.0, lmax=100.0, precision=0):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 6)
        out = np.empty_like(z)
        L.shim_prox_tris(D(lmin), D(lmax), int(precision), z.shape[0], checkers.dp(z), checkers.dp(out))
        return out

    def prox_tets_warm(model, mu, lam, z, q, precision=0):
        z = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 9)
        out = np.empty_like(z)
        assert L.shim_prox_tets_warm(int(model), D(mu), D(lam), D(0.0), int(precision), z.shape[0], checkers.dp(z), checkers.dp(out), checkers.dp(q)) == 0
        return out

    prox_tets.warm = prox_tets_warm
    return prox_tets, prox_tris


@pytest.mark.parametrize("model", range(6))
def test_device_prox_source_vs_golden(shim, model):
    g = np.load(os.path.join(HERE, "golden", "prox_vectors.npz"))
    mu, lam = g["mu_lambda"]
    for precision, tol in ((1, 5e-6), (0, 5e-5)):
        out = shim[0](model, mu, lam, g["tet%d_in" % model], precision=precision)
        ref = g["tet%d_out" % model]
        assert (np.abs(out - ref) / np.maximum(1.0, np.abs(ref))).max() < tol


@pytest.mark.parametrize("model", [1, 2, 4])
@pytest.mark.parametrize("sigma", [0.01, 0.3, 0.6])
def test_device_prox_source_vs_oracle(shim, model, sigma):
    z = checkers.random_F(1500, sigma, seed=300 + model)
    ref, _ = checkers.prox_tets("oracle", model, MU, LAM, z)
    for precision, tol in ((1, 5e-6), (0, 5e-5)):
        out = shim[0](model, MU, LAM, z, precision=precision)
        assert (np.abs(out - ref) / np.maximum(1.0, np.abs(ref))).max() < tol


def test_device_prox_source_degenerate_elements(shim):
    """Inverted, collapsed and resting elements: the cases that take the reference-faithful path."""
    z = checkers.random_F(96, 0.2, seed=7)
    z[:32, 6:9] *= -1.0
    z[32:40] *= 1e-9
    z[40:48] = np.eye(3).ravel()
    for model in (0, 1, 2):
        ref, _ = checkers.prox_tets("oracle", model, MU, LAM, z)
        for precision, tol in ((1, 5e-6), (0, 5e-5)):
            out = shim[0](model, MU, LAM, z, precision=precision)
            e = np.abs(out - ref).max(axis=1)
            assert e[:32].max() < tol and e[40:].max() < tol
            sa = np.linalg.svd(out[32:40].reshape(-1, 3, 3), compute_uv=False)
            sb = np.linalg.svd(ref[32:40].reshape(-1, 3, 3), compute_uv=False)
            assert np.abs(sa - sb).max() < tol


def test_device_tri_prox_source_vs_golden(shim):
    g = np.load(os.path.join(HERE, "golden", "prox_vectors.npz"))
    for precision, tol in ((1, 1e-10), (0, 2e-6)):
        assert np.abs(shim[1](g["tri_in"], precision=precision) - g["tri_out"]).max() < tol
        assert np.abs(shim[1](g["tri_in"], 0.95, 1.05, precision=precision) - g["tri_lim_out"]).max() < tol


@pytest.mark.parametrize("model", [0, 1, 2])
def test_device_prox_source_warm_started_svd(shim, model):
    """The SVD warm start (quaternion of V carried between calls, csrc/prox.cuh svd3_signed): same answers as
    the cold start, call after call on a slowly changing F -- including inverted and resting elements -- and
    the carried quaternion never leaves the rotations."""
    rng = np.random.RandomState(11)
    z = checkers.random_F(2000, 0.3, seed=500 + model)
    z[:200, 6:9] *= -1.0           # inverted
    z[200:300] = np.eye(3).ravel()  # at rest: V is arbitrary
    for precision, tol in ((1, 5e-6), (0, 5e-5)):
        q = np.zeros((len(z), 4))
        zz = z.copy()
        for it in range(6):
            ref, _ = checkers.prox_tets("oracle", model, MU, LAM, zz)
            out = shim[0].warm(model, MU, LAM, zz, q, precision=precision)
            e = np.abs(out - ref).max(axis=1) / np.maximum(1.0, np.abs(ref).max(axis=1))
            # inverted StVK elements take the reference-faithful path: same tolerance, checked by the degenerate test
            assert e.max() < tol, (precision, it, e.argmax(), e.max())
            assert np.isfinite(q).all() and (np.abs(q).max(axis=1) > 0.2).all()
            zz = zz + 0.02 * rng.randn(*zz.shape)  # the next ADMM iteration's F + u is close to this one's
