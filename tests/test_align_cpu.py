"""Point-set alignment (compute_rigid_transform / procrustes_transform / compute_similarity_transform,
libs/common/transformation.py:48-141): the oracle and the host-compiled kernel math (pose_math.h) against
goldens produced by the reference's own functions (tests/golden/make_golden.py::golden_align)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pnp_ref, pose_ref
from oracle.egonet_ref import KITTI_K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tests', 'native', 'libpose_host.so')


@pytest.fixture(scope='module')
def host():
    src = os.path.join(ROOT, 'tests', 'native', 'pose_host.cpp')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'egonet_b200', 'csrc'),
                           src, '-o', SO])
    return ctypes.CDLL(SO)


def dp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


@pytest.mark.parametrize('P', [8, 9, 32])
def test_oracle_alignment_vs_reference_golden(golden, P):
    g = golden('align.npz')
    t = 'p%d_' % P
    X, Y = g[t + 'X'], g[t + 'Y']
    for name, Ws in (('plain', [None] * len(X)), ('diag', g[t + 'Wd']), ('full', g[t + 'Wf'])):
        for i in range(len(X)):
            R, tt = pose_ref.compute_rigid_transform(X[i], Y[i], Ws[i])
            np.testing.assert_allclose(R, g[t + name + '_R'][i], atol=1e-12)
            np.testing.assert_allclose(tt, g[t + name + '_t'][i], atol=1e-12)
    for i in range(len(X)):
        np.testing.assert_allclose(pose_ref.procrustes_transform(X[i], Y[i]), g[t + 'procrustes'][i], atol=1e-12)
        for scale, k in ((False, 'sim_'), (True, 'sim_scale_')):
            d, Z, T, b, c = pose_ref.compute_similarity_transform(X[i].T.copy(), Y[i].T.copy(), scale)
            assert d == pytest.approx(g[t + k + 'd'][i], abs=1e-12)
            np.testing.assert_allclose(Z, g[t + k + 'Z'][i], atol=1e-11)
            np.testing.assert_allclose(T, g[t + k + 'T'][i], atol=1e-12)
            assert b == pytest.approx(g[t + k + 'b'][i], abs=1e-12)
            np.testing.assert_allclose(c, g[t + k + 'c'][i], atol=1e-11)


@pytest.mark.parametrize('P', [8, 9, 32])
def test_kernel_math_alignment_vs_reference_golden(host, golden, P):
    """The exact __host__ __device__ source the GPU kernels run, compiled with g++."""
    g = golden('align.npz')
    t = 'p%d_' % P
    X = np.ascontiguousarray(g[t + 'X'].transpose(0, 2, 1))     # [N,P,3]
    Y = np.ascontiguousarray(g[t + 'Y'].transpose(0, 2, 1))
    n = len(X)
    for name, mode, W in (('plain', 0, None), ('diag', 1, np.ascontiguousarray(g[t + 'Wd'])),
                          ('full', 2, np.ascontiguousarray(g[t + 'Wf']))):
        R, tt, al = np.zeros((n, 3, 3)), np.zeros((n, 3)), np.zeros((n, P, 3))
        host.host_rigid_transform(dp(X), dp(Y), dp(W), mode, n, P, dp(R), dp(tt), dp(al))
        np.testing.assert_allclose(R, g[t + name + '_R'], atol=1e-10)
        np.testing.assert_allclose(tt, g[t + name + '_t'][:, :, 0], atol=1e-9)
        if mode == 0:
            np.testing.assert_allclose(al.transpose(0, 2, 1), g[t + 'procrustes'], atol=1e-9)
        assert np.allclose(np.linalg.det(R), 1.0, atol=1e-12)   # reflected clouds included
    for scale, k in ((0, 'sim_'), (1, 'sim_scale_')):
        d, b, Z, T, c = np.zeros(n), np.zeros(n), np.zeros((n, P, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3))
        host.host_similarity_transform(dp(X), dp(Y), n, P, scale, dp(d), dp(b), dp(Z), dp(T), dp(c))
        np.testing.assert_allclose(d, g[t + k + 'd'], atol=1e-10)
        np.testing.assert_allclose(b, g[t + k + 'b'], atol=1e-10)
        np.testing.assert_allclose(Z, g[t + k + 'Z'], atol=1e-9)
        np.testing.assert_allclose(T, g[t + k + 'T'], atol=1e-10)
        np.testing.assert_allclose(c, g[t + k + 'c'], atol=1e-9)


def test_oracle_refine_with_predicted_bbox_vs_reference_golden(golden):
    g = golden('align.npz')
    preds, obs = pnp_ref.synth_cases(n=32, seed=34, points=9, offset=1.5)
    rel = preds.copy()
    rel[:, 1:] -= rel[:, :1]
    np.testing.assert_allclose([rel.sum(), obs.sum()], g['bbox_digest'], rtol=1e-12)
    conv = g['bbox_converged']
    for thr in (5.0, 1.5):
        ok_ref, out_ref = g['bbox_thr%g_ok' % thr], g['bbox_thr%g_refined' % thr]
        assert 0 < ok_ref.sum()
        for i in range(len(rel)):
            if not conv[i]:
                continue
            ok, r = pnp_ref.refine_with_predicted_bbox(rel[i], obs[i], KITTI_K, thr)
            assert ok == ok_ref[i]
            if ok:
                np.testing.assert_allclose(r.T, out_ref[i], atol=1e-6)
