"""Reprojection refinement on the CPU: the oracle restatement of cv2.solvePnP(SOLVEPNP_ITERATIVE) and the
kernel source (egonet_b200/csrc/pnp_math.h compiled for the host) against goldens produced by the
reference's own pnp_refine."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pnp_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tests', 'native', 'libpnp_host.so')
CASES = {'p9': dict(n=48, points=9), 'p33': dict(n=16, points=33),
         'p9_noisy': dict(n=32, points=9, noise_3d=0.25, noise_px=1.5)}
TOL_CONVERGED = 1e-6      # metres, on boxes 6-60 m away (relative ~1e-8)


def cases(g, tag):
    preds, obs = pnp_ref.synth_cases(seed=int(g[tag + '_seed']), **CASES[tag])
    np.testing.assert_allclose([preds.sum(), obs.sum()], g[tag + '_digest'], rtol=1e-12)
    return preds, obs


@pytest.fixture(scope='module')
def host():
    src = os.path.join(ROOT, 'tests', 'native', 'pnp_host.cpp')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'egonet_b200', 'csrc'),
                           src, '-o', SO])
    L = ctypes.CDLL(SO)
    L.host_pnp_refine.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] +
                                  [ctypes.c_double] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 4)
    return L


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_refine(L, preds, obs, K):
    n, p = preds.shape[:2]
    out, pose, info, st = np.zeros((n, p, 3)), np.zeros((n, 6)), np.zeros((n, 2)), np.zeros(n, np.int32)
    L.host_pnp_refine(vp(np.ascontiguousarray(preds)), vp(np.ascontiguousarray(obs)), n, p, K[0, 0], K[1, 1],
                      K[0, 2], K[1, 2], 0, vp(out), vp(pose), vp(info), vp(st))
    return out, pose, info, st


@pytest.mark.parametrize('tag', list(CASES))
def test_oracle_pnp_vs_reference_golden(golden, tag):
    g = golden('pnp.npz')
    preds, obs = cases(g, tag)
    conv = g[tag + '_converged']
    assert conv.sum() >= len(conv) // 2
    for i in np.where(conv)[0]:
        ref = pnp_ref.pnp_refine(preds[i], obs[i], g['K'], np.zeros((4, 1)))
        np.testing.assert_allclose(ref.T, g[tag + '_refined'][i], rtol=0, atol=TOL_CONVERGED)
        r, t, _, _ = pnp_ref.solve_pnp_iterative(preds[i], obs[i], g['K'])
        np.testing.assert_allclose(np.concatenate([r, t]), g[tag + '_rt'][i], rtol=0, atol=TOL_CONVERGED)


@pytest.mark.parametrize('tag', list(CASES))
def test_kernel_source_on_host_vs_reference_golden(golden, host, tag):
    g = golden('pnp.npz')
    preds, obs = cases(g, tag)
    out, pose, info, st = host_refine(host, preds, obs, g['K'])
    conv = g[tag + '_converged']
    assert not st.any()
    np.testing.assert_allclose(out[conv], g[tag + '_refined'][conv], rtol=0, atol=TOL_CONVERGED)
    np.testing.assert_allclose(pose[conv], g[tag + '_rt'][conv], rtol=0, atol=TOL_CONVERGED)
    assert (info[conv, 0] < 20).all() and (info[~conv, 0] == 20).all()
    # 20-step cap reached (cv2 not converged): same iteration, chaotic amplification of rounding only
    rel = np.abs(out[~conv] - g[tag + '_refined'][~conv]).max(axis=(1, 2)) / np.abs(g[tag + '_refined'][~conv]).max(axis=(1, 2))
    assert np.median(rel) < 1e-6


def test_kernel_source_on_host_properties(host):
    """Exact data: the refinement recovers a known rigid motion; planar / degenerate inputs are flagged and
    returned unchanged; SO(3) helpers round-trip."""
    from oracle.egonet_ref import KITTI_K as K
    preds, _ = pnp_ref.synth_cases(12, 40, noise_3d=0.0, noise_px=0.0, offset=0.0)
    rng = np.random.Generator(np.random.PCG64(41))
    moved = np.zeros_like(preds)
    obs = np.zeros(preds.shape[:2] + (2,))
    for i, X in enumerate(preds):
        R = pnp_ref.rodrigues(rng.uniform(-0.2, 0.2, 3))
        t = rng.uniform(-1, 1, 3)
        moved[i] = X @ R.T + t
        uv = moved[i] @ K.T
        obs[i] = uv[:, :2] / uv[:, 2:3]
    out, pose, info, st = host_refine(host, preds, obs, K)
    assert not st.any()
    np.testing.assert_allclose(out, moved, rtol=0, atol=1e-7)
    assert (info[:, 1] < 1e-6).all()
    planar = preds.copy()
    planar[:, :, 1] = 1.5
    out, _, _, st = host_refine(host, planar, obs, K)
    assert (st == 1).all()
    np.testing.assert_array_equal(out, planar)
    same = np.repeat(preds[:, :1], preds.shape[1], axis=1)
    assert (host_refine(host, same, obs, K)[3] == 2).all()
    R9, back, Jl = np.zeros(9), np.zeros(3), np.zeros(9)
    for r in (np.array([0.3, -1.2, 0.5]), np.array([1e-10, 0, 0]), np.array([0, 3.1, 0.2])):
        host.host_so3(vp(r), vp(R9), vp(back), vp(Jl))
        np.testing.assert_allclose(R9.reshape(3, 3), pnp_ref.rodrigues(r), atol=1e-14)
        np.testing.assert_allclose(back, r, atol=1e-9)
        np.testing.assert_allclose(Jl.reshape(3, 3), pnp_ref.left_jacobian(r), atol=1e-12)
    A = rng.standard_normal((12, 12))
    A = A @ A.T
    w, V = np.zeros(12), np.zeros((12, 12))
    host.host_eigh12(vp(np.ascontiguousarray(A)), vp(w), vp(V))
    np.testing.assert_allclose(np.sort(w), np.linalg.eigvalsh(A), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(V @ np.diag(w) @ V.T, A, atol=1e-9)
