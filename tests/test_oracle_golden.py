"""The oracle (oracle/*) replayed against outputs of the UPSTREAM reference code.

Fixtures under tests/golden/*.npz were produced by tests/golden/make_golden.py,
which imports and executes the reference from /root/reference.  These tests run
anywhere (no GPU, no /root/reference).
"""
import numpy as np
import pytest
import torch

from oracle import affine_ref, configs, decode_ref, egonet_ref, hrnet_ref, lifter_ref, loss_ref, pose_ref

HRNET_CASES = [('tiny', configs.tiny_cfgs()), ('tiny_heatmap', configs.tiny_cfgs('heatmap')),
               ('ped', configs.ped_cfgs()), ('demo', configs.demo_cfgs())]


@pytest.mark.parametrize('tag,cfgs', HRNET_CASES, ids=[c[0] for c in HRNET_CASES])
def test_hrnet_oracle_matches_reference(golden, tag, cfgs):
    g = golden('hrnet_%s.npz' % tag)
    sd = hrnet_ref.make_weights(cfgs, int(g['seed_w']))
    assert hrnet_ref.weights_digest(sd) == pytest.approx(float(g['weights_digest']), rel=1e-12)
    x = egonet_ref.synth_crops(int(g['batch']), cfgs, int(g['seed_x']))
    out = hrnet_ref.hrnet_forward(sd, cfgs, x)
    maps = (out[0] if isinstance(out, tuple) else out).numpy()
    rs = int(g['map_row_stride'])
    # same torch-CPU kernels as the reference module: expect (near) bit equality
    np.testing.assert_allclose(maps[:, :, ::rs, :], g['maps_sub'], rtol=0, atol=1e-5)
    flat = maps.reshape(maps.shape[0], maps.shape[1], -1)
    np.testing.assert_array_equal(flat.argmax(2), g['maps_argmax'])
    np.testing.assert_allclose(flat.max(2), g['maps_max'], atol=1e-5)
    if isinstance(out, tuple):
        np.testing.assert_allclose(out[1].numpy(), g['coords'], rtol=0, atol=1e-6)


def test_hrnet_param_and_mac_counts():
    # SURVEY.md 8c known answers: 63,978,471 params, 21,017,256,960 MAC/crop
    cfgs = configs.demo_cfgs()
    spec = hrnet_ref.state_dict_spec(cfgs)
    n = sum(int(np.prod(s)) for k, s in spec.items()
            if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked')))
    assert n == 63978471
    assert len(spec) == 1828
    assert hrnet_ref.macs_per_crop(cfgs) == 21017256960
    assert hrnet_ref.macs_per_crop(configs.demo_cfgs('heatmap')) == 20925579264
    lspec = lifter_ref.state_dict_spec(cfgs)
    assert sum(int(np.prod(s)) for k, s in lspec.items()
               if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))) == 4375648


def test_quantized_model_is_close_to_exact():
    cfgs = configs.tiny_cfgs()
    sd = hrnet_ref.make_weights(cfgs, 1)
    x = egonet_ref.synth_crops(2, cfgs, 0)
    m0, c0 = hrnet_ref.hrnet_forward(sd, cfgs, x)
    m1, c1 = hrnet_ref.hrnet_forward(sd, cfgs, x, ctx=hrnet_ref.Quantized(torch.float16))
    assert (c1 - c0).abs().max() < 2e-3
    assert (m1 - m0).abs().max() < 0.05


def test_decode_oracle_matches_reference(golden):
    g = golden('decode.npz')
    for tag in ('hm', 'pos'):
        arr = g[tag]
        p, m, idx = decode_ref.get_max_preds(arr)
        np.testing.assert_array_equal(p, g[tag + '_max_preds'])
        np.testing.assert_array_equal(m, g[tag + '_max_vals'])
        p, m = decode_ref.soft_arg_max(arr)
        np.testing.assert_allclose(p, g[tag + '_soft_preds'], rtol=0, atol=1e-4)
        np.testing.assert_array_equal(m, g[tag + '_soft_vals'])
        with np.errstate(all='ignore'):
            p, m = decode_ref.soft_arg_max_np(arr)
        ref = g[tag + '_softnp_preds']
        ok = np.isfinite(ref)
        assert np.array_equal(np.isfinite(p), ok)
        # sum-normalisation of zero-mean maps is ill-conditioned (division by a
        # near-zero sum): compare relative to magnitude
        np.testing.assert_allclose(p[ok], ref[ok], rtol=2e-3, atol=1e-3)
        np.testing.assert_array_equal(m, g[tag + '_softnp_vals'])


def test_affine_oracle_matches_reference(golden):
    g = golden('affine.npz')
    n = len(g['boxes'])
    for i in range(n):
        ret = affine_ref.modify_bbox(g['boxes'][i], g['ars'][i])
        np.testing.assert_allclose(ret['c'], g['centers'][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(ret['s'], g['scales'][i], rtol=0, atol=1e-12)
        np.testing.assert_allclose(ret['bbox'], g['bbox_resize'][i], rtol=0, atol=1e-10)
        res = (256, 256) if g['ars'][i] == 1.0 else (192, 256)
        ti = affine_ref.get_affine_transform(ret['c'], ret['s'], 0., (res[1], res[0]), inv=1)
        tf = affine_ref.get_affine_transform(ret['c'], ret['s'], 0., (res[1], res[0]), inv=0)
        np.testing.assert_allclose(ti, g['trans_inv'][i], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(tf, g['trans_fwd'][i], rtol=1e-12, atol=1e-9)
        scr = affine_ref.local_to_screen(g['coords'][i:i + 1], [ret['c']], [ret['s']], [0.], res)[0]
        np.testing.assert_allclose(scr, g['screen'][i], rtol=0, atol=1e-8)
    # SURVEY.md 8c known answer (5)
    t = affine_ref.get_affine_transform([600, 180], [0.5, 0.5], 0., (256, 256), inv=1)
    np.testing.assert_allclose(t, [[0.390625, 0, 550], [0, 0.390625, 130]], atol=1e-9)


@pytest.mark.parametrize('tag', ['demo', 'tiny'])
def test_lifter_oracle_matches_reference(golden, tag):
    g = golden('lifter_%s.npz' % tag)
    cfgs = configs.demo_cfgs() if tag == 'demo' else configs.tiny_cfgs()
    sd = lifter_ref.make_weights(cfgs, 11)
    assert hrnet_ref.weights_digest(sd) == pytest.approx(float(g['digest']), rel=1e-12)
    stats = lifter_ref.make_stats(cfgs, 12)
    out = lifter_ref.lift_2d_to_3d(sd, cfgs, stats, g['kpts'])
    np.testing.assert_allclose(out, g['kpts_3d'], rtol=0, atol=1e-5)


def test_pose_oracle_matches_reference(golden):
    g = golden('pose.npz')
    preds = g['preds']
    for i in range(len(preds)):
        tpl = pose_ref.get_template(preds[i])
        np.testing.assert_allclose(tpl, g['templates'][i], rtol=0, atol=1e-12)
        R, _ = pose_ref.compute_rigid_transform(tpl, preds[i].T)
        np.testing.assert_allclose(R, g['R'][i], rtol=0, atol=1e-10)
    angles, trans = pose_ref.get_6d_rep(preds)
    np.testing.assert_allclose(angles, g['angles'], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(trans, g['translation'])
    # noise-free cuboids recover the generating angles (SURVEY.md 8c known answer 2)
    np.testing.assert_allclose(angles[:8], g['true_angles'][:8], atol=1e-6)
    np.testing.assert_allclose(pose_ref.observation_angle_trans(angles, trans), g['alpha_trans'], atol=1e-12)
    kp = [k.reshape(1, -1) for k in g['kpts']]
    np.testing.assert_allclose(pose_ref.observation_angle_proj(angles, kp, g['K']), g['alpha_proj'], atol=1e-12)
    # SURVEY.md 8c known answers (4)
    assert float(g['ka_trans'][0]) == pytest.approx(0.60033135, abs=1e-7)
    assert float(g['ka_proj'][0]) == pytest.approx(0.5650404, abs=1e-6)
    assert np.all(np.abs(g['alpha_trans']) <= np.pi) and np.all(np.abs(g['alpha_proj']) <= np.pi)


def test_pipeline_oracle_matches_reference(golden):
    g = golden('pipeline_tiny.npz')
    cfgs = configs.tiny_cfgs()
    n = len(g['centers'])
    crops = egonet_ref.synth_crops(n, cfgs, 0)
    recs = egonet_ref.synth_boxes(n, cfgs, 2)
    np.testing.assert_allclose(np.array([r['center'] for r in recs]), g['centers'], atol=1e-12)
    np.testing.assert_allclose(np.array([r['scale'] for r in recs]), g['scales'], atol=1e-12)
    out = egonet_ref.run_pipeline(hrnet_ref.make_weights(cfgs, 1), lifter_ref.make_weights(cfgs, 11),
                                  lifter_ref.make_stats(cfgs, 12), cfgs, crops, recs)
    np.testing.assert_allclose(out['kpts_2d'], g['kpts_2d'], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out['kpts_3d'], g['kpts_3d'], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out['euler'], g['euler'], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out['translation'], g['translation'], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out['alpha_trans'], g['alpha_trans'], rtol=0, atol=1e-4)
    np.testing.assert_allclose(out['alpha_proj'], g['alpha_proj'], rtol=0, atol=1e-4)


def test_loss_oracle_matches_reference(golden):
    g = golden('loss.npz')
    for use_w in (0, 1):
        loss, grad = loss_ref.joints_mse_loss(g['pred'], g['gt'], g['w'] if use_w else None)
        assert loss == pytest.approx(float(g['loss_w%d' % use_w]), rel=1e-6)
        np.testing.assert_allclose(grad, g['grad_w%d' % use_w], rtol=1e-5, atol=1e-10)
    assert float(g['calc_hm_loss']) == pytest.approx(float(g['loss_w0']), rel=1e-7)


COMPOSITE_CASES = {'shipped': dict(coor_kind='l1', coor_weight=0.1, hm_w=1.0), 'all': dict(coor_kind='l1', coor_weight=0.1, cr_kind='sl1', cr_weight=0.01, hm_w=1.0),
                   'sl1_mse': dict(coor_kind='sl1', coor_weight=2.0, cr_kind='mse', cr_weight=0.05, hm_w=0.5),
                   'coor_only': dict(coor_kind='mse', coor_weight=1.0, hm_w=0.0)}


@pytest.mark.parametrize('tag', list(COMPOSITE_CASES))
def test_composite_loss_oracle_matches_reference(golden, tag):
    """Coordinate + cross-ratio terms (function.py:113-202) and the heat-map term against the reference's
    JointsCompositeLoss and its autograd gradients (make_golden.golden_composite)."""
    g = golden('loss_composite.npz')
    kw = dict(COMPOSITE_CASES[tag])
    hm_w = kw.pop('hm_w')
    total, coor, cr, grad = loss_ref.composite_coord_terms(g['coords'], g['joints'], [256, 256], cr_indices=g['cr_indices'], **kw)
    hm_loss, hm_grad = loss_ref.joints_mse_loss(g['hm_pred'], g['hm_gt'])
    assert total + hm_w * hm_loss == pytest.approx(float(g[tag + '_loss']), rel=2e-6)
    np.testing.assert_allclose(grad, g[tag + '_dcoords'], rtol=0, atol=2e-6 * max(1e-3, np.abs(g[tag + '_dcoords']).max()))
    np.testing.assert_allclose(hm_w * hm_grad, g[tag + '_dhm'], rtol=1e-5, atol=1e-10)
    if tag + '_mask' in g:
        m = loss_ref.cr_mask(g['coords'], g['cr_indices'], 0.15)
        np.testing.assert_array_equal(m, g[tag + '_mask'][:, :, 0])
        assert 0 < m.sum() < m.size                     # both masked and unmasked lines are present
