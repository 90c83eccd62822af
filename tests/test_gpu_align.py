"""GPU parity of the batched point-set alignment entries (``egn_rigid_transform``, ``egn_similarity_transform``,
``egn_refine_with_bbox``) through the mirror of ``libs/common/transformation.py`` against goldens produced by the
reference's own functions (transformation.py:48-141, tools/inference_legacy.py:518-547)."""
import numpy as np
import pytest
import torch

from oracle import pnp_ref
from oracle.egonet_ref import KITTI_K

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from egonet_b200.libs.common import transformation as T


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


@pytest.mark.parametrize('P', [8, 9, 32])
def test_rigid_and_similarity_vs_reference_golden(golden, P):
    g = golden('align.npz')
    t = 'p%d_' % P
    X, Y = dev(g[t + 'X'].transpose(0, 2, 1)), dev(g[t + 'Y'].transpose(0, 2, 1))
    for name, W in (('plain', None), ('diag', dev(g[t + 'Wd'])), ('full', dev(g[t + 'Wf']))):
        R, tt, al = T.rigid_transform_batch(X, Y, W, want_aligned=True)
        np.testing.assert_allclose(R.cpu().numpy(), g[t + name + '_R'], atol=1e-10)            # 1e-10 on R
        np.testing.assert_allclose(tt.cpu().numpy(), g[t + name + '_t'][:, :, 0], atol=1e-9)
        if W is None:
            np.testing.assert_allclose(al.cpu().numpy().transpose(0, 2, 1), g[t + 'procrustes'], atol=1e-9)
    for scale, k in ((False, 'sim_'), (True, 'sim_scale_')):
        d, Z, Tm, b, c = T.similarity_transform_batch(X, Y, scale)
        np.testing.assert_allclose(d.cpu().numpy(), g[t + k + 'd'], atol=1e-10)
        np.testing.assert_allclose(Z.cpu().numpy(), g[t + k + 'Z'], atol=1e-9)
        np.testing.assert_allclose(Tm.cpu().numpy(), g[t + k + 'T'], atol=1e-10)
        np.testing.assert_allclose(b.cpu().numpy(), g[t + k + 'b'], atol=1e-10)
        np.testing.assert_allclose(c.cpu().numpy(), g[t + k + 'c'], atol=1e-9)
    # reference signatures ([3,P] numpy in, numpy out)
    i = 3
    R1, t1 = T.compute_rigid_transform(g[t + 'X'][i], g[t + 'Y'][i], g[t + 'Wd'][i])
    np.testing.assert_allclose(R1, g[t + 'diag_R'][i], atol=1e-10)
    np.testing.assert_allclose(t1, g[t + 'diag_t'][i], atol=1e-9)
    assert t1.shape == (3, 1)
    np.testing.assert_allclose(T.procrustes_transform(g[t + 'X'][i], g[t + 'Y'][i]), g[t + 'procrustes'][i], atol=1e-9)
    d1, Z1, T1, b1, c1 = T.compute_similarity_transform(g[t + 'X'][i].T, g[t + 'Y'][i].T, True)
    assert d1 == pytest.approx(g[t + 'sim_scale_d'][i], abs=1e-10) and b1 == pytest.approx(g[t + 'sim_scale_b'][i], abs=1e-10)
    np.testing.assert_allclose(Z1, g[t + 'sim_scale_Z'][i], atol=1e-9)


def test_rigid_transform_properties():
    """Exact rigid motions are recovered; N = 0; a 4096-instance batch repeats its rows exactly."""
    from scipy.spatial.transform import Rotation
    rng = np.random.Generator(np.random.PCG64(17))
    X = rng.standard_normal((64, 12, 3))
    Rm = Rotation.from_rotvec(rng.uniform(-3, 3, (64, 3))).as_matrix()
    tv = rng.uniform(-10, 10, (64, 3))
    Y = np.einsum('nab,npb->npa', Rm, X) + tv[:, None]
    R, t = T.rigid_transform_batch(dev(X), dev(Y))
    np.testing.assert_allclose(R.cpu().numpy(), Rm, atol=1e-12)
    np.testing.assert_allclose(t.cpu().numpy(), tv, atol=1e-11)
    R0, t0 = T.rigid_transform_batch(dev(np.zeros((0, 12, 3))), dev(np.zeros((0, 12, 3))))
    assert R0.shape == (0, 3, 3) and t0.shape == (0, 3)
    big = T.rigid_transform_batch(dev(np.tile(X, (64, 1, 1))), dev(np.tile(Y, (64, 1, 1))))[0]
    assert torch.equal(big[:64], big[-64:]) and torch.equal(big[:64], R)
    with pytest.raises(ValueError):
        T.rigid_transform_batch(dev(X), dev(Y[:, :5]))


def test_refine_with_predicted_bbox_vs_reference_golden(golden):
    g = golden('align.npz')
    preds, obs = pnp_ref.synth_cases(n=32, seed=34, points=9, offset=1.5)
    rel = preds.copy()
    rel[:, 1:] -= rel[:, :1]
    conv = g['bbox_converged']
    for thr in (5.0, 1.5):
        ok, out = T.refine_with_predicted_bbox_batch(rel, obs, KITTI_K, thr)
        ok, out = ok.cpu().numpy(), out.cpu().numpy()
        ok_ref, out_ref = g['bbox_thr%g_ok' % thr], g['bbox_thr%g_refined' % thr]
        np.testing.assert_array_equal(ok[conv], ok_ref[conv])
        sel = conv & ok_ref
        np.testing.assert_allclose(out[sel], out_ref[sel], rtol=0, atol=1e-6)       # metres, as test_gpu_pnp
        rej = conv & ~ok_ref
        np.testing.assert_allclose(out[rej], preds[rej], rtol=0, atol=1e-12)        # discarded: absolute unrefined box
    i = int(np.where(conv & g['bbox_thr5_ok'])[0][0])
    ok1, r1 = T.refine_with_predicted_bbox(rel[i], obs[i], KITTI_K, np.zeros((4, 1)))
    assert ok1
    np.testing.assert_allclose(r1, g['bbox_thr5_refined'][i].T, atol=1e-6)
    j = np.where(conv & ~g['bbox_thr1.5_ok'])[0]
    if len(j):
        assert T.refine_with_predicted_bbox(rel[j[0]], obs[j[0]], KITTI_K, np.zeros((4, 1)), threshold=1.5) == (False, None)


def test_kpts_to_euler_uses_the_callers_template(golden):
    """EgoNet.kpts_to_euler(template, prediction) (egonet.py:265-277) with the reference's templates, and with a
    template that is NOT get_template(prediction): R, T must follow the template that was passed."""
    from egonet_b200.libs.model.egonet import EgoNet
    from oracle import pose_ref
    g = golden('pose.npz')
    ego = EgoNet.__new__(EgoNet)
    for i in (0, 5, 20, 63):
        ang, t = ego.kpts_to_euler(g['templates'][i], g['preds'][i].T)
        np.testing.assert_allclose(ang, g['angles'][i], atol=1e-9)
        R, t_ref = pose_ref.compute_rigid_transform(g['templates'][i], g['preds'][i].T)
        np.testing.assert_allclose(t, t_ref, atol=1e-9)
    other = g['templates'][7] * np.array([[1.3], [0.8], [1.1]])
    ang, t = ego.kpts_to_euler(other, g['preds'][9].T)
    R, t_ref = pose_ref.compute_rigid_transform(other, g['preds'][9].T)
    np.testing.assert_allclose(ang, pose_ref.euler_yxz(R), atol=1e-9)
    np.testing.assert_allclose(t, t_ref, atol=1e-9)
