"""GPU parity of the batched reprojection refinement (``egn_pnp_refine`` through the mirror of
``libs/common/transformation.py``) against goldens produced by the reference's own ``pnp_refine``
(cv2.solvePnP ITERATIVE executed in the build container) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import pnp_ref
from oracle.egonet_ref import KITTI_K

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from egonet_b200.libs.common import transformation

CASES = {'p9': dict(n=48, points=9), 'p33': dict(n=16, points=33),
         'p9_noisy': dict(n=32, points=9, noise_3d=0.25, noise_px=1.5)}
TOL_CONVERGED = 1e-6      # metres (boxes 6-60 m from the camera), where cv2 met its own stop criterion


@pytest.mark.parametrize('tag', list(CASES))
def test_pnp_refine_vs_reference_golden(golden, tag):
    g = golden('pnp.npz')
    preds, obs = pnp_ref.synth_cases(seed=int(g[tag + '_seed']), **CASES[tag])
    np.testing.assert_allclose([preds.sum(), obs.sum()], g[tag + '_digest'], rtol=1e-12)
    out, pose, info, status = transformation.pnp_refine_batch(preds, obs, g['K'], return_info=True)
    out, pose, info, status = out.cpu().numpy(), pose.cpu().numpy(), info.cpu().numpy(), status.cpu().numpy()
    conv = g[tag + '_converged']
    assert not status.any()
    np.testing.assert_allclose(out[conv], g[tag + '_refined'][conv], rtol=0, atol=TOL_CONVERGED)
    np.testing.assert_allclose(pose[conv], g[tag + '_rt'][conv], rtol=0, atol=TOL_CONVERGED)
    assert (info[conv, 0] < 20).all()
    # cv2 cut off by its 20-step cap: the same iteration, but chaotic -- rounding differences are amplified
    ref = g[tag + '_refined'][~conv]
    rel = np.abs(out[~conv] - ref).max(axis=(1, 2)) / np.abs(ref).max(axis=(1, 2))
    assert np.median(rel) < 1e-4
    # single-instance signature of the reference: [P,3], [P,2] -> [3,P]
    i = int(np.where(conv)[0][0])
    one = transformation.pnp_refine(preds[i], obs[i], g['K'], np.zeros((4, 1)))
    np.testing.assert_allclose(one, g[tag + '_refined'][i].T, rtol=0, atol=TOL_CONVERGED)


def test_pnp_refine_properties_and_edge_cases():
    """Exact data recovers the rigid motion; planar / degenerate instances are flagged and unchanged;
    N = 0; a batch of 4096 gives the same answer for duplicated instances."""
    K = KITTI_K
    preds, _ = pnp_ref.synth_cases(16, 40, noise_3d=0.0, noise_px=0.0, offset=0.0)
    rng = np.random.Generator(np.random.PCG64(41))
    moved, obs = np.zeros_like(preds), np.zeros(preds.shape[:2] + (2,))
    for i, X in enumerate(preds):
        moved[i] = X @ pnp_ref.rodrigues(rng.uniform(-0.2, 0.2, 3)).T + rng.uniform(-1, 1, 3)
        uv = moved[i] @ K.T
        obs[i] = uv[:, :2] / uv[:, 2:3]
    out, pose, info, status = transformation.pnp_refine_batch(preds, obs, K, return_info=True)
    assert not status.any().item()
    np.testing.assert_allclose(out.cpu().numpy(), moved, rtol=0, atol=1e-7)
    assert (info[:, 1] < 1e-6).all().item()
    planar = preds.copy()
    planar[:, :, 1] = 1.5
    o2, _, _, s2 = transformation.pnp_refine_batch(planar, obs, K, return_info=True)
    assert (s2 == 1).all().item()
    np.testing.assert_array_equal(o2.cpu().numpy(), planar)
    same = np.repeat(preds[:, :1], preds.shape[1], axis=1)
    o3, _, _, s3 = transformation.pnp_refine_batch(same, obs, K, return_info=True)
    assert (s3 != 0).all().item()           # all points identical: degenerate (or, by rounding, "planar"): not refined
    np.testing.assert_array_equal(o3.cpu().numpy(), same)
    empty = transformation.pnp_refine_batch(np.zeros((0, 9, 3)), np.zeros((0, 9, 2)), K)
    assert empty.shape == (0, 9, 3)
    big_p, big_o = np.tile(preds, (256, 1, 1)), np.tile(obs, (256, 1, 1))
    big = transformation.pnp_refine_batch(big_p, big_o, K)
    assert torch.equal(big[:16], big[-16:]) and torch.equal(big[:16], out)
    with pytest.raises(RuntimeError):
        transformation.pnp_refine_batch(np.zeros((1, 4, 3)), np.zeros((1, 4, 2)), K)
    with pytest.raises(NotImplementedError):
        transformation.pnp_refine(preds[0], obs[0], K, np.array([0.1, 0, 0, 0]))
