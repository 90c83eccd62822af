"""The fp64 geometry source (egonet_b200/csrc/pose_math.h) compiled for the HOST
with g++ and checked against the golden vectors / the oracle -- validates the
exact kernel source on a machine without a GPU.  The harness
(tests/native/pose_host.cpp) is test-only and never loaded by the product."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pose_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tests', 'native', 'libpose_host.so')


@pytest.fixture(scope='module')
def host():
    src = os.path.join(ROOT, 'tests', 'native', 'pose_host.cpp')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'egonet_b200', 'csrc'),
                           src, '-o', SO])
    L = ctypes.CDLL(SO)
    L.host_pose_solve.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                  ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


def dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_pose_math_matches_reference_golden(host, golden):
    g = golden('pose.npz')
    pr = np.ascontiguousarray(g['preds'])
    k2 = np.ascontiguousarray(g['kpts'])
    n = len(pr)
    out, rot = np.zeros((n, 7)), np.zeros((n, 9))
    for mode, key in ((0, 'alpha_trans'), (1, 'alpha_proj')):
        host.host_pose_solve(dp(pr), n, 32, dp(k2), 66, g['K'][0, 0], g['K'][0, 2], mode, dp(out), dp(rot))
        np.testing.assert_allclose(out[:, :3], g['angles'], atol=1e-9)
        np.testing.assert_allclose(rot.reshape(n, 3, 3), g['R'], atol=1e-10)
        np.testing.assert_array_equal(out[:, 3:6], g['translation'])
        np.testing.assert_allclose(out[:, 6], g[key], atol=1e-9)


def test_affine_math_matches_reference_golden(host, golden):
    g = golden('affine.npz')
    for ar, res in ((1.0, (256, 256)), (256 / 192, (192, 256))):
        sel = np.where(g['ars'] == ar)[0]
        co = np.ascontiguousarray(g['coords'][sel])
        ce = np.ascontiguousarray(g['centers'][sel])
        sc = np.ascontiguousarray(g['scales'][sel])
        scr = np.zeros((len(sel), 33, 2))
        host.host_local_to_screen(dp(co), dp(ce), dp(sc), None, len(sel), 33, res[0], res[1], dp(scr))
        np.testing.assert_allclose(scr, g['screen'][sel], rtol=0, atol=1e-9)


def test_kabsch_properties(host):
    """Random, reflected and rank-deficient cross-covariances against numpy's SVD route."""
    rng = np.random.Generator(np.random.PCG64(42))
    R = np.zeros(9)
    for trial in range(200):
        H = rng.standard_normal((3, 3))
        if trial % 5 == 0:
            H[:, 2] = H[:, 0] * 0.5 - H[:, 1]           # rank 2
        if trial % 7 == 0:
            H *= 1e-6
        host.host_kabsch(dp(np.ascontiguousarray(H)), dp(R))
        Rm = R.reshape(3, 3)
        np.testing.assert_allclose(Rm @ Rm.T, np.eye(3), atol=1e-12)
        assert np.linalg.det(Rm) == pytest.approx(1.0, abs=1e-12)
        U, S, Vt = np.linalg.svd(H)
        Rn = Vt.T @ U.T
        if np.linalg.det(Rn) < 0:
            Vt[-1] *= -1
            Rn = Vt.T @ U.T
        if S[1] > 1e-9 * S[0] and (S[2] > 1e-9 * S[0] or trial % 5 == 0):
            np.testing.assert_allclose(Rm, Rn, atol=1e-7 if trial % 5 == 0 else 1e-9)


def test_euler_roundtrip_property(host):
    """Rotating a canonical cuboid by R(ry, rx, rz) returns those angles (away from gimbal lock)."""
    from scipy.spatial.transform import Rotation
    rng = np.random.Generator(np.random.PCG64(43))
    n = 128
    ang = np.stack([rng.uniform(-1.2, 1.2, n), rng.uniform(-np.pi, np.pi, n), rng.uniform(-np.pi, np.pi, n)], 1)
    cub = []
    for i in range(n):
        l, h, w = rng.uniform(3, 5), rng.uniform(1.2, 2), rng.uniform(1.4, 2)
        x = np.array([l, l, l, l, 0, 0, 0, 0]) - l / 2
        y = np.array([0, h, 0, h, 0, h, 0, h]) - h
        z = np.array([w, w, 0, 0, w, w, 0, 0]) - w / 2
        c = np.array([x, y, z])
        par, chi = pose_ref.BBOX12_PARENTS - 1, pose_ref.BBOX12_CHILDREN - 1
        c = np.hstack([c, c[:, par] + 0.332 * (c[:, chi] - c[:, par]), c[:, par] + 0.667 * (c[:, chi] - c[:, par])])
        Rm = Rotation.from_euler('yxz', [ang[i, 1], ang[i, 0], ang[i, 2]]).as_matrix()
        cub.append((Rm @ c).T + rng.uniform(-10, 10, 3))
    pr = np.ascontiguousarray(np.array(cub))
    out = np.zeros((n, 7))
    host.host_pose_solve(dp(pr), n, 32, None, 0, 0.0, 0.0, 0, dp(out), None)
    np.testing.assert_allclose(out[:, :3], ang, atol=1e-6)
    assert np.all(np.abs(out[:, 6]) <= np.pi + 1e-12)
