"""Generate the golden fixtures by EXECUTING THE UPSTREAM REFERENCE CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every array below is an output of the reference's own functions/modules
(imported unmodified from /root/reference; matplotlib stubbed because it is
only used for plotting) on seeded inputs.  Inputs and weights are NOT stored:
they are regenerated from PCG64 seeds by ``oracle.*.make_weights`` /
``oracle.egonet_ref.synth_*`` (a digest of the regenerated weights is stored to
detect drift).  Library versions used are recorded in ``versions.json``.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from _refimport import import_reference  # noqa: E402
from oracle import configs, crop_ref, egonet_ref, hrnet_ref, lifter_ref, pnp_ref, train_ref  # noqa: E402


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print('%-28s %7.1f KB' % (name, os.path.getsize(path) / 1024))


def golden_hrnet(ref, tag, cfgs, batch, seed_w, seed_x, map_row_stride):
    model = ref['hrnet'].get_pose_net(cfgs, is_train=False).eval()
    sd = hrnet_ref.make_weights(cfgs, seed_w)
    model.load_state_dict(sd)                       # strict: key set must match upstream
    x = egonet_ref.synth_crops(batch, cfgs, seed_x)
    taps = {}
    hooks = []
    # per-stage activation statistics from the upstream module (forward hooks)
    def stat_hook(name):
        def fn(_m, _i, out):
            o = out[0] if isinstance(out, (list, tuple)) else out
            taps[name] = np.array([o.double().mean().item(), o.double().std().item(),
                                   o.double().abs().max().item()])
        return fn
    for name in ('layer1', 'stage2', 'stage3', 'stage4'):
        hooks.append(getattr(model, name).register_forward_hook(stat_hook(name)))
    with torch.no_grad():
        out = model(x)
    for h in hooks:
        h.remove()
    arrays = {'seed_w': seed_w, 'seed_x': seed_x, 'batch': batch,
              'weights_digest': hrnet_ref.weights_digest(sd),
              'map_row_stride': map_row_stride}
    for k, v in taps.items():
        arrays['stat_' + k] = v
    if isinstance(out, tuple):
        maps, coords = out
        arrays['coords'] = coords.numpy()
    else:
        maps = out
    maps = maps.numpy()
    arrays['maps_sub'] = maps[:, :, ::map_row_stride, :]
    flat = maps.reshape(maps.shape[0], maps.shape[1], -1)
    arrays['maps_argmax'] = flat.argmax(2).astype(np.int32)
    arrays['maps_max'] = flat.max(2)
    arrays['maps_sum'] = flat.astype(np.float64).sum(2)
    save('hrnet_%s.npz' % tag, **arrays)


def golden_decode(ref):
    ip = ref['img_proc']
    g = rng(3)
    hm = g.standard_normal((3, 33, 64, 64), dtype=np.float32)
    # adversarial maps: duplicated maxima, all-negative, all-zero, plateau, single spike, non-square
    hm[0, 0, 10, 20] = hm[0, 0, 40, 5] = 9.0
    hm[0, 1] = -np.abs(hm[0, 1]) - 0.1
    hm[0, 2] = 0.0
    hm[0, 3, :, :] = 1.5
    hm[0, 4] = 0.0
    hm[0, 4, 63, 63] = 2.0
    hm[0, 5, 0, 0] = hm[0, 5, 0, 1] = 7.0
    pos = np.abs(g.standard_normal((2, 5, 12, 20), dtype=np.float32)) + 0.01   # H != W, positive
    arrays = {'hm': hm, 'pos': pos}
    for tag, arr in (('hm', hm), ('pos', pos)):
        p, m = ip.get_max_preds(arr.copy())
        arrays[tag + '_max_preds'], arrays[tag + '_max_vals'] = p, m
        p, m = ip.soft_arg_max_np(arr.copy())       # divides its argument in place upstream
        arrays[tag + '_softnp_preds'], arrays[tag + '_softnp_vals'] = p, m
        # upstream soft_arg_max is CUDA-only (img_proc.py:696-700); run it unmodified on CPU by
        # pointing the two CUDA-only names it touches at their CPU equivalents
        import torch.cuda.comm  # noqa: F401  (lazy sub-module in current torch)
        saved = (torch.cuda.FloatTensor, torch.cuda.comm.broadcast)
        torch.cuda.FloatTensor = torch.FloatTensor
        torch.cuda.comm.broadcast = lambda t, devices: [t]
        try:
            p, m = ip.soft_arg_max(torch.from_numpy(arr.copy()))
        finally:
            torch.cuda.FloatTensor, torch.cuda.comm.broadcast = saved
        arrays[tag + '_soft_preds'], arrays[tag + '_soft_vals'] = p.numpy(), m.numpy()
    save('decode.npz', **arrays)


def golden_affine(ref):
    ip = ref['img_proc']
    g = rng(5)
    n = 24
    boxes = np.stack([g.uniform(0, 1100, n), g.uniform(0, 300, n)], 1)
    boxes = np.concatenate([boxes, boxes + np.stack([g.uniform(20, 420, n), g.uniform(15, 260, n)], 1)], 1)
    ars = np.where(np.arange(n) % 3 == 0, 256 / 192, 1.0)
    centers, scales, bbs, tinv, tfwd, screens = [], [], [], [], [], []
    coords = g.uniform(0, 1, (n, 33, 2)).astype(np.float32)
    for i in range(n):
        ret = ip.modify_bbox(boxes[i], ars[i])
        res = (256, 256) if ars[i] == 1.0 else (192, 256)   # (width, height)
        centers.append(ret['c']); scales.append(ret['s']); bbs.append(ret['bbox'])
        ti = ip.get_affine_transform(ret['c'], ret['s'], 0., (res[1], res[0]), inv=1)
        tf = ip.get_affine_transform(ret['c'], ret['s'], 0., (res[1], res[0]), inv=0)
        local = coords[i].copy()
        local *= np.array(res).reshape(1, 2)
        screens.append(ip.affine_transform_modified(local, ti))
        tinv.append(ti); tfwd.append(tf)
    save('affine.npz', boxes=boxes, ars=ars, centers=np.array(centers), scales=np.array(scales),
         bbox_resize=np.array(bbs), trans_inv=np.array(tinv), trans_fwd=np.array(tfwd),
         coords=coords, screen=np.array(screens))


def golden_lifter(ref):
    for tag, cfgs in (('demo', configs.demo_cfgs()), ('tiny', configs.tiny_cfgs())):
        model = ref['fcmodel'].get_fc_model(1, cfgs, cfgs['FCModel']['input_size'],
                                            cfgs['FCModel']['output_size']).eval()
        sd = lifter_ref.make_weights(cfgs, 11)
        model.load_state_dict(sd)
        stats = lifter_ref.make_stats(cfgs, 12)
        g = rng(13)
        n = 19
        kpts = stats['mean_in'] + stats['std_in'] * g.standard_normal((n, cfgs['FCModel']['input_size']))
        # the upstream dtype chain of EgoNet.lift_2d_to_3d (egonet.py:473-485)
        nop = ref['operations']
        data = nop.normalize_1d(kpts, stats['mean_in'], stats['std_in']).astype(np.float32)
        with torch.no_grad():
            raw = model(torch.from_numpy(data)).numpy()
        pred = nop.unnormalize_1d(raw, stats['mean_out'], stats['std_out'])
        save('lifter_%s.npz' % tag, kpts=kpts, raw=raw, kpts_3d=pred.reshape(n, -1, 3),
             digest=hrnet_ref.weights_digest(sd))


def _cuboid(l, h, w):
    x = np.array([l, l, l, l, 0, 0, 0, 0]) - l / 2
    y = np.array([0, h, 0, h, 0, h, 0, h]) - h
    z = np.array([w, w, 0, 0, w, w, 0, 0]) - w / 2
    c = np.array([x, y, z])
    par = np.array([1, 3, 5, 7, 1, 2, 3, 4, 1, 2, 5, 6]) - 1
    chi = np.array([2, 4, 6, 8, 5, 6, 7, 8, 3, 4, 7, 8]) - 1
    seg = c[:, chi] - c[:, par]
    return np.hstack([c, c[:, par] + 0.332 * seg, c[:, par] + 0.667 * seg]).T   # [32,3]


def golden_pose(ref):
    from scipy.spatial.transform import Rotation
    ego = ref['egonet'].EgoNet.__new__(ref['egonet'].EgoNet)   # methods only use their arguments
    g = rng(7)
    n = 64
    preds = np.zeros((n, 32, 3))
    true_angles = np.zeros((n, 3))
    for i in range(n):
        l, h, w = g.uniform(3, 5), g.uniform(1.2, 2), g.uniform(1.4, 2)
        ry, rx, rz = g.uniform(-np.pi, np.pi), g.uniform(-0.3, 0.3), g.uniform(-0.3, 0.3)
        if i < 4:
            ry, rx, rz = [(0.7, 0, 0), (-2.5, 0.1, 0.05), (3.0, 0, 0), (0, 0, 0)][i]
        R = Rotation.from_euler('yxz', [ry, rx, rz]).as_matrix()
        pts = (R @ _cuboid(l, h, w).T).T + np.array([g.uniform(-20, 20), g.uniform(0, 3), g.uniform(5, 60)])
        noise = 0.0 if i < 8 else (0.05 if i < 40 else 0.4)
        preds[i] = pts + noise * g.standard_normal((32, 3))
        true_angles[i] = [rx, ry, rz]
    # two reflected (mirrored) clouds force the det(R) < 0 branch of the Kabsch solve
    preds[-1, :, 0] *= -1
    preds[-2, :, 2] *= -1
    angles, trans = ego.get_6d_rep(preds.reshape(n, -1))
    templates = np.array([ego.get_template(p) for p in preds])
    Rs = np.array([ref['transformation'].compute_rigid_transform(templates[i], preds[i].T)[0] for i in range(n)])
    a_trans = ego.get_observation_angle_trans(angles, trans)
    kx = g.uniform(0, 1242, n)
    kpts = [np.concatenate([[kx[i]], g.uniform(0, 375, 65)]).reshape(1, 66) for i in range(n)]
    a_proj = ego.get_observation_angle_proj(angles, kpts, egonet_ref.KITTI_K)
    save('pose.npz', preds=preds, true_angles=true_angles, angles=angles, translation=trans,
         templates=templates, R=Rs, alpha_trans=a_trans, alpha_proj=a_proj,
         kpts=np.concatenate(kpts, 0), K=egonet_ref.KITTI_K,
         ka_trans=ego.get_observation_angle_trans(np.array([[0, 0.7, 0.]]), np.array([[2., 1, 20]])),
         ka_proj=ego.get_observation_angle_proj(np.array([[0, 0.7, 0.]]), [np.array([[700.0]])],
                                                np.array([[707., 0, 604], [0, 707, 180], [0, 0, 1]])))


def golden_pipeline(ref, tag='tiny', cfgs=None, paths=None):
    """EgoNet.get_keypoints -> lift_2d_to_3d -> gather_lifting_results of the upstream class
    (tiny config: 6 crops in 3 images; demo config = the benchmarked HRNet-W48: 8 crops in 3 images)."""
    cfgs = cfgs or configs.tiny_cfgs()
    paths = paths or ['img_a.png'] * 2 + ['img_b.png'] * 3 + ['img_c.png']
    ego = ref['egonet'].EgoNet(cfgs, pre_trained=False).eval()
    hc_sd = hrnet_ref.make_weights(cfgs, 1)
    l_sd = lifter_ref.make_weights(cfgs, 11)
    ego.HC.load_state_dict(hc_sd)
    ego.L.load_state_dict(l_sd)
    ego.LS = lifter_ref.make_stats(cfgs, 12)
    n = len(paths)
    crops = egonet_ref.synth_crops(n, cfgs, 0)
    recs = egonet_ref.synth_boxes(n, cfgs, 2)
    for r, p in zip(recs, paths):
        r.update(path=p, label=-1, score=-1.0)
    with torch.no_grad():
        records = ego.get_keypoints(crops, [dict(r) for r in recs], is_cuda=False)
        records = ego.lift_2d_to_3d(records, cuda=False)
    out = {'kpts_2d': [], 'kpts_3d': [], 'euler': [], 'translation': [], 'alpha_trans': [], 'alpha_proj': []}
    for p in sorted(set(paths)):
        rec = records[p]
        rec['K'] = egonet_ref.KITTI_K
        for mode in ('trans', 'proj'):
            rec = ego.gather_lifting_results(rec, None, None, alpha_mode=mode)
            out['alpha_' + mode].append(rec['alphas'].copy())
        out['kpts_2d'].append(np.concatenate(rec['kpts_2d_pred'], 0))
        out['kpts_3d'].append(rec['kpts_3d_pred'])
        out['euler'].append(rec['euler_angles'])
        out['translation'].append(rec['translation'])
    save('pipeline_%s.npz' % tag, **{k: np.concatenate(v, 0) for k, v in out.items()},
         centers=np.array([r['center'] for r in recs]), scales=np.array([r['scale'] for r in recs]),
         paths=np.array(paths))


def golden_loss(ref):
    """JointsMSELoss (with / without target weights, incl. autograd gradient) and calc_hm_loss."""
    import libs.loss.function as F
    g = rng(21)
    pred = g.standard_normal((4, 33, 16, 12), dtype=np.float32)
    gt = np.exp(-g.uniform(0, 6, (4, 33, 16, 12))).astype(np.float32)
    w = (g.uniform(0, 1, (4, 33, 1)) > 0.3).astype(np.float32)
    out = {'pred': pred, 'gt': gt, 'w': w}
    for use_w in (False, True):
        p = torch.tensor(pred, requires_grad=True)
        loss = F.JointsMSELoss(use_w)(p, torch.tensor(gt), torch.tensor(w))
        loss.backward()
        out['loss_w%d' % use_w] = loss.detach().numpy()
        out['grad_w%d' % use_w] = p.grad.numpy()
    comp = F.JointsCompositeLoss.__new__(F.JointsCompositeLoss)
    comp.comp_dict = {'hm': (torch.nn.MSELoss(reduction='mean'), 1.0)}
    out['calc_hm_loss'] = comp.calc_hm_loss(torch.tensor(pred), torch.tensor(gt)).numpy()
    save('loss.npz', **out)

# mean / std of cfgs['dataset']['pth_transform'] (configs/KITTI_inference:demo.yml:49-53)
CROP_MEAN = [0.485, 0.456, 0.406]
CROP_STD = [0.229, 0.224, 0.225]
# x1, y1, x2, y2 on a 375 x 1242 image: ordinary boxes, up-/down-sampling extremes, boxes cut by each
# border, one entirely outside (all-zero crop), one with integer-aligned geometry
CROP_BOXES = np.array([
    [600.3, 150.2, 760.9, 260.7], [35.5, 170.0, 130.25, 240.5], [1100.0, 120.0, 1240.0, 330.0],
    [400.0, 175.0, 424.0, 193.0], [100.0, 10.0, 1100.0, 370.0], [-60.0, 200.0, 90.0, 300.0],
    [500.0, -40.0, 700.0, 80.0], [1180.0, 300.0, 1300.0, 420.0], [1400.0, 100.0, 1500.0, 200.0],
    [512.0, 128.0, 768.0, 256.0], [611.7, 180.3, 640.2, 201.9], [300.123, 90.456, 555.789, 310.012]])


def golden_crop(ref):
    """crop_single_instance of the reference (cv2.warpAffine + torchvision ToTensor/Normalize) on a seeded
    synthetic KITTI-sized image; the image itself is regenerated from its seed by the tests."""
    import hashlib
    from torchvision import transforms
    ego_cls = ref['egonet'].EgoNet
    pth_trans = transforms.Compose([transforms.ToTensor(),
                                    transforms.Normalize(mean=CROP_MEAN, std=CROP_STD)])   # car_instance.py:522-531
    img = crop_ref.synth_image(375, 1242, 21)
    arrays = {'image_seed': 21, 'image_shape': np.array(img.shape), 'mean': np.array(CROP_MEAN), 'std': np.array(CROP_STD),
              'image_sha1': np.frombuffer(hashlib.sha1(img.tobytes()).digest(), np.uint8), 'boxes': CROP_BOXES}
    # full uint8 crops are stored for a few boxes only (they do not compress); every box has its SHA-1
    for tag, res, full in (('sq', (256, 256), (0, 3, 5, 9)), ('ped', (192, 256), (0, 7))):   # res = (width, height)
        u8, sha, sub, sums, cs, ss = [], [], [], [], [], []
        for i, b in enumerate(CROP_BOXES):
            raw = ego_cls.crop_single_instance(None, img, b, res, pth_trans=None, xy_dict=None)
            ten = ego_cls.crop_single_instance(None, img, b, res, pth_trans=pth_trans, xy_dict=None)
            assert raw.dtype == np.uint8 and raw.shape == (res[1], res[0], 3) and ten.dtype == torch.float32
            ret = ref['img_proc'].modify_bbox(b, res[1] / res[0])
            if i in full:
                u8.append(raw)
            sha.append(np.frombuffer(hashlib.sha1(np.ascontiguousarray(raw).tobytes()).digest(), np.uint8))
            sub.append(ten.numpy()[:, ::4, ::4])
            sums.append(ten.double().sum().item())
            cs.append(ret['c']); ss.append(ret['s'])
        arrays.update({tag + '_u8': np.array(u8), tag + '_u8_index': np.array(full), tag + '_u8_sha1': np.array(sha),
                       tag + '_norm_sub': np.array(sub), tag + '_norm_sum': np.array(sums),
                       tag + '_centers': np.array(cs), tag + '_scales': np.array(ss)})
    save('crop.npz', **arrays)

def golden_pnp(ref):
    """pnp_refine of the reference (cv2.solvePnP ITERATIVE + Rodrigues) on seeded cuboids; inputs are
    regenerated from the seeds by ``oracle.pnp_ref.synth_cases``.  ``converged`` marks the instances where
    cv2 stopped on its relative-step criterion (fewer than 20 LM iterations in the restatement) -- on the
    others the 20-step cap cuts a still-moving, chaotic iteration and only loose agreement is meaningful."""
    import cv2
    fn = ref['transformation'].pnp_refine
    K = egonet_ref.KITTI_K
    arrays = {'K': K}
    for tag, kw in (('p9', dict(n=48, seed=31, points=9)), ('p33', dict(n=16, seed=32, points=33)),
                    ('p9_noisy', dict(n=32, seed=33, points=9, noise_3d=0.25, noise_px=1.5))):
        preds, obs = pnp_ref.synth_cases(**kw)
        refined, rts, conv = [], [], []
        for X, uv in zip(preds, obs):
            refined.append(fn(X, uv, K, np.zeros((4, 1))).T)
            ok, rv, tv = cv2.solvePnP(X, uv, K, np.zeros((4, 1)), flags=cv2.SOLVEPNP_ITERATIVE)
            rts.append(np.concatenate([rv.ravel(), tv.ravel()]))
            conv.append(pnp_ref.solve_pnp_iterative(X, uv, K)[2] < 20)
        arrays.update({tag + '_seed': kw['seed'], tag + '_refined': np.array(refined), tag + '_rt': np.array(rts),
                       tag + '_converged': np.array(conv), tag + '_digest': np.array([preds.sum(), obs.sum()])})
    save('pnp.npz', **arrays)

def golden_align(ref):
    """compute_rigid_transform (plain / diagonal W / full W / reflected), procrustes_transform,
    compute_similarity_transform (both scale modes) of the reference on seeded point sets, and
    refine_with_predicted_bbox (tools/inference_legacy.py:518-547) on the pnp cases."""
    tr = ref['transformation']
    g = rng(91)
    arrays = {}
    for P in (8, 9, 32):
        n = 24
        X = g.standard_normal((n, 3, P)) * g.uniform(0.5, 3, (n, 1, 1))
        Y = np.zeros_like(X)
        from scipy.spatial.transform import Rotation
        for i in range(n):
            Rm = Rotation.from_rotvec(g.uniform(-3, 3, 3)).as_matrix()
            Y[i] = g.uniform(0.5, 2) * (Rm @ X[i]) + g.uniform(-5, 5, (3, 1)) + 0.1 * g.standard_normal((3, P))
        Y[-1, 0] *= -1                       # mirrored clouds: det(R) < 0 branch
        Y[-2, 2] *= -1
        Wd = g.uniform(0.1, 2, (n, P))
        Wf = g.uniform(0, 1, (n, P, P))
        tag = 'p%d_' % P
        arrays[tag + 'X'], arrays[tag + 'Y'], arrays[tag + 'Wd'], arrays[tag + 'Wf'] = X, Y, Wd, Wf
        for name, Ws in (('plain', [None] * n), ('diag', Wd), ('full', Wf)):
            Rt = [tr.compute_rigid_transform(X[i], Y[i], Ws[i]) for i in range(n)]
            arrays[tag + name + '_R'] = np.array([r for r, _ in Rt])
            arrays[tag + name + '_t'] = np.array([t for _, t in Rt])
        arrays[tag + 'procrustes'] = np.array([tr.procrustes_transform(X[i], Y[i]) for i in range(n)])
        for scale in (False, True):
            outs = [tr.compute_similarity_transform(X[i].T.copy(), Y[i].T.copy(), scale) for i in range(n)]
            k = tag + ('sim_scale_' if scale else 'sim_')
            arrays[k + 'd'] = np.array([o[0] for o in outs])
            arrays[k + 'Z'] = np.array([o[1] for o in outs])
            arrays[k + 'T'] = np.array([o[2] for o in outs])
            arrays[k + 'b'] = np.array([float(o[3]) for o in outs])
            arrays[k + 'c'] = np.array([o[4] for o in outs])
    # refine_with_predicted_bbox is defined in tools/inference_legacy.py, whose module-level imports need a
    # display stack; the function only uses ltr.pnp_refine + numpy, so execute its source with those bound
    import ast
    src = open('/root/reference/tools/inference_legacy.py').read()
    fn_src = next(ast.get_source_segment(src, n) for n in ast.parse(src).body
                  if isinstance(n, ast.FunctionDef) and n.name == 'refine_with_predicted_bbox')
    ns = {'np': np, 'ltr': tr}
    exec(compile(fn_src, 'inference_legacy.refine_with_predicted_bbox', 'exec'), ns)
    K = egonet_ref.KITTI_K
    preds, obs = pnp_ref.synth_cases(n=32, seed=34, points=9, offset=1.5)
    rel = preds.copy()
    rel[:, 1:] -= rel[:, :1]
    for thr in (5.0, 1.5):
        oks, outs = [], []
        for X, uv in zip(rel, obs):
            ok, r = ns['refine_with_predicted_bbox'](X, uv, K, np.zeros((4, 1)), threshold=thr)
            oks.append(ok)
            outs.append(r.T if ok else np.full((9, 3), np.nan))
        arrays['bbox_thr%g_ok' % thr] = np.array(oks)
        arrays['bbox_thr%g_refined' % thr] = np.array(outs)
    arrays['bbox_converged'] = np.array([pnp_ref.solve_pnp_iterative(X, uv, K)[2] < 20 for X, uv in zip(preds, obs)])
    arrays['bbox_digest'] = np.array([rel.sum(), obs.sum()])
    # EgoNet.add_orientation_arrow (egonet.py:157-179) on seeded records (short and long arrows)
    ego = ref['egonet'].EgoNet.__new__(ref['egonet'].EgoNet)
    ga = rng(92)
    gt = ga.standard_normal((12, 32, 3)) + np.array([0., 1., 20.])
    gt[:, :, 2] = np.abs(gt[:, :, 2]) + 4
    pk = gt + ga.standard_normal((12, 32, 3)) * np.linspace(0.05, 4, 12)[:, None, None]
    arrays['arrow_pred'], arrays['arrow_gt'] = pk, gt
    arrays['arrow'] = ego.add_orientation_arrow({'kpts_3d_pred': pk, 'kpts_3d_gt': gt, 'K': K})
    save('align.npz', **arrays)


def golden_format(ref):
    """get_pred_str / get_instance_str of the reference (libs/common/format.py:25-61) on seeded detector rows."""
    import importlib
    fmt = importlib.import_module('libs.common.format')
    g = rng(61)
    n = 7
    rows = []
    for i in range(n):
        row = {'class': ['Car', 'Van', 'Pedestrian'][i % 3], 'truncation': float(g.uniform(0, 1)),
               'occlusion': float(g.integers(0, 3)), 'alpha': float(g.uniform(-3, 3)),
               'bbox': [float(v) for v in g.uniform(0, 1200, 4)],
               'dimensions': [float(v) for v in g.uniform(1, 5, 3)],
               'locations': [float(v) for v in g.uniform(-30, 60, 3)], 'rot_y': float(g.uniform(-3, 3))}
        if i % 2 == 0:
            row['score'] = float(g.uniform(0, 1))
        rows.append(row)
    record = {'raw_txt_format': rows, 'euler_angles': g.uniform(-3.2, 3.2, (n, 3)), 'alphas': g.uniform(-3.2, 3.2, n)}
    out = {'rows': rows, 'euler_angles': record['euler_angles'].tolist(), 'alphas': record['alphas'].tolist(),
           'pred_str': fmt.get_pred_str(record), 'instance_strs': [fmt.get_instance_str(r) for r in rows]}
    with open(os.path.join(HERE, 'format.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('format.json')

TRAIN_FULL_GRADS = ('conv1.weight', 'stage3.0.branches.1.1.bn1.weight', 'final_layer.bias',
                    'stage2.0.fuse_layers.1.0.0.0.weight')
TRAIN_STATS = ('bn1', 'stage4.0.branches.3.1.bn2')


def golden_train(ref):
    """One training-mode forward + backward of the reference HC module (heat-map head, JointsMSELoss with target
    weights) on Gaussian targets made by the reference's generate_target: loss, a norm of EVERY parameter
    gradient, a few gradients in full, updated BN running statistics."""
    import libs.loss.function as LF
    cfgs = configs.tiny_cfgs('heatmap')
    hm = cfgs['heatmapModel']
    model = ref['hrnet'].get_pose_net(cfgs, is_train=True)
    sd = hrnet_ref.make_weights(cfgs, 5)
    model.load_state_dict(sd)
    model.train()
    B, K = 3, hm['num_joints']
    x = egonet_ref.synth_crops(B, cfgs, 7)
    g = rng(71)
    joints = np.concatenate([g.uniform(-60, hm['input_size'][0] + 60, (B, K, 2)), np.ones((B, K, 1))], 2)
    vis = (g.uniform(0, 1, (B, K)) > 0.2).astype(np.float32)
    params = {'num_joints': K, 'target_type': 'gaussian', 'input_size': np.array(hm['input_size']),
              'heatmap_size': np.array(hm['heatmap_size']), 'sigma': 2, 'use_different_joints_weight': False}
    tgts, wts = zip(*[ref['img_proc'].generate_target(joints[b], vis[b], params) for b in range(B)])
    target, weight = torch.from_numpy(np.stack(tgts)), torch.from_numpy(np.stack(wts))
    out = model(x)
    loss = LF.JointsMSELoss(True)(out, target, weight)
    loss.backward()
    named = dict(model.named_parameters())
    new_sd = model.state_dict()
    arrays = {'joints': joints, 'vis': vis, 'target': target.numpy(), 'target_weight': weight.numpy(),
              'loss': loss.detach().numpy(), 'seed_w': 5, 'seed_x': 7, 'sigma': 2,
              'grad_names': np.array(list(named.keys())),
              'grad_norms': np.array([named[k].grad.double().norm().item() for k in named]),
              'grad_sums': np.array([named[k].grad.double().sum().item() for k in named]),
              'out_sum': np.array(out.detach().double().sum().item())}
    for k in TRAIN_FULL_GRADS:
        arrays['grad__' + k] = named[k].grad.numpy()
    for k in TRAIN_STATS:
        arrays['stat__' + k + '.running_mean'] = new_sd[k + '.running_mean'].numpy()
        arrays['stat__' + k + '.running_var'] = new_sd[k + '.running_var'].numpy()
        arrays['stat__' + k + '.num_batches_tracked'] = new_sd[k + '.num_batches_tracked'].numpy()
    save('train_tiny.npz', **arrays)


CR_BBOX12 = np.array([[1, 9, 21, 2], [3, 10, 22, 4], [5, 11, 23, 6], [7, 12, 24, 8], [1, 13, 25, 5], [2, 14, 26, 6],
                      [3, 15, 27, 7], [4, 16, 28, 8], [1, 17, 29, 3], [2, 18, 30, 4], [5, 19, 31, 7], [6, 20, 32, 8]])


class _cpu_cuda:
    """JointsCompositeLoss.forward moves its targets with ``.cuda()`` (function.py:187); point it at the CPU while
    the goldens are produced (the arithmetic is device independent)."""

    def __enter__(self):
        self.saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t

    def __exit__(self, *exc):
        torch.Tensor.cuda = self.saved


def golden_composite(ref):
    """JointsCompositeLoss of the reference (function.py:61-202) with the shipped specification
    (['mse', 'l1', 'sl1'], weights [1.0, 0.1, 'None'], KITTI_train_IGRs.yml:88-90) and with every term on
    (cross-ratio weight 0.01, apply_cr_loss): loss and autograd gradients w.r.t. heat-maps and coordinates."""
    import libs.loss.function as LF
    import libs.dataset.KITTI.car_instance as CI
    assert np.array_equal(CI.cr_indices_dict['bbox12'], CR_BBOX12)
    g = rng(101)
    B, K = 5, 33
    arrays = {'cr_indices': CR_BBOX12}
    hm_pred = g.standard_normal((B, K, 16, 16)).astype(np.float32)
    hm_gt = g.uniform(0, 1, (B, K, 16, 16)).astype(np.float32)
    # a plausible projected cuboid per sample + noise, in (0, 1)
    base = g.uniform(0.2, 0.8, (B, 1, 2)) + 0.15 * g.standard_normal((B, K, 2))
    coords = np.clip(base, 0.02, 0.98).astype(np.float32)
    coords[0, 9] = coords[0, 1] + 0.01          # a fore-shortened edge: masked out of the cross-ratio term
    joints = np.concatenate([(coords + 0.03 * g.standard_normal((B, K, 2))) * 256, np.ones((B, K, 1))], 2)
    arrays.update(hm_pred=hm_pred, hm_gt=hm_gt, coords=coords, joints=joints)
    for tag, specs, weights, apply_cr in (('shipped', ['mse', 'l1', 'sl1'], [1.0, 0.1, 'None'], True),
                                          ('all', ['mse', 'l1', 'sl1'], [1.0, 0.1, 0.01], True),
                                          ('sl1_mse', ['mse', 'sl1', 'mse'], [0.5, 2.0, 0.05], True),
                                          ('coor_only', ['None', 'mse', 'None'], [1, 1, 1], False)):
        f = LF.JointsCompositeLoss(spec_list=specs, img_size=[256, 256], hm_size=[16, 16], loss_weights=weights,
                                   cr_loss_thres=0.15)
        f.cr_indices, f.target_cr, f.apply_cr_loss = CR_BBOX12, 4 / 3, apply_cr
        hp = torch.from_numpy(hm_pred).requires_grad_(True)
        cp = torch.from_numpy(coords).requires_grad_(True)
        with _cpu_cuda():
            loss = f((hp, cp), torch.from_numpy(hm_gt), None, {'transformed_joints': joints.copy()})
        loss.backward()
        arrays[tag + '_loss'] = loss.detach().numpy()
        arrays[tag + '_dcoords'] = cp.grad.numpy()
        arrays[tag + '_dhm'] = hp.grad.numpy() if hp.grad is not None else np.zeros_like(hm_pred)
        if 'cr' in f.comp_dict and f.comp_dict['cr'][1] != 'None' and apply_cr:
            arrays[tag + '_mask'] = f.get_cr_mask(coords, 0.15).numpy()
    save('loss_composite.npz', **arrays)


def golden_train_coord(ref):
    """One training-mode forward + backward of the reference HC module with the COORDINATE head and the shipped
    composite loss (heat-map MSE + 0.1 * L1 on coordinates): loss, every gradient norm, a few gradients in full."""
    import libs.loss.function as LF
    cfgs = configs.tiny_cfgs()
    hm = cfgs['heatmapModel']
    model = ref['hrnet'].get_pose_net(cfgs, is_train=True)
    model.load_state_dict(hrnet_ref.make_weights(cfgs, 6))
    model.train()
    B, K = 3, hm['num_joints']
    x = egonet_ref.synth_crops(B, cfgs, 8)
    g = rng(72)
    joints = np.concatenate([g.uniform(5, hm['input_size'][0] - 5, (B, K, 2)), np.ones((B, K, 1))], 2)
    vis = np.ones((B, K), dtype=np.float32)
    params = {'num_joints': K, 'target_type': 'gaussian', 'input_size': np.array(hm['input_size']),
              'heatmap_size': np.array(hm['heatmap_size']), 'sigma': 2, 'use_different_joints_weight': False}
    tgts, wts = zip(*[ref['img_proc'].generate_target(joints[b], vis[b], params) for b in range(B)])
    target = torch.from_numpy(np.stack(tgts))
    f = LF.JointsCompositeLoss(spec_list=['mse', 'l1', 'sl1'], img_size=hm['input_size'], hm_size=hm['heatmap_size'],
                               loss_weights=[1.0, 0.1, 'None'], cr_loss_thres=0.15)
    f.cr_indices, f.target_cr = CR_BBOX12, 4 / 3
    out = model(x)
    with _cpu_cuda():
        loss = f(out, target, None, {'transformed_joints': joints.copy()})
    loss.backward()
    named = dict(model.named_parameters())
    arrays = {'joints': joints, 'target': target.numpy(), 'loss': loss.detach().numpy(), 'seed_w': 6, 'seed_x': 8,
              'coords': out[1].detach().numpy(), 'maps_sum': np.array(out[0].detach().double().sum().item()),
              'grad_names': np.array(list(named.keys())),
              'grad_norms': np.array([named[k].grad.double().norm().item() for k in named])}
    for k in ('head2.4.weight', 'head2.4.bias', 'head1.0.weight', 'head2.0.conv1.weight', 'head2.3.bn2.weight'):
        arrays['grad__' + k] = named[k].grad.numpy()
    save('train_tiny_coord.npz', **arrays)


KITTI_LABEL = """Car 0.00 0 -1.58 587.01 173.33 614.12 200.12 1.65 1.67 3.64 -0.65 1.71 46.70 -1.59 0.91
Pedestrian 0.00 1 0.21 423.17 173.67 433.17 224.03 1.60 0.38 0.30 -5.87 1.63 23.11 -0.03 0.55
Car 0.88 3 1.88 0.00 192.37 402.31 374.00 1.48 1.60 3.69 -2.84 1.65 4.11 1.33 0.99
Cyclist 0.00 0 -1.40 676.60 163.95 688.98 193.93 1.86 0.60 2.02 4.59 1.32 45.84 -1.30 0.30
Car 0.00 2 -1.69 657.39 190.13 700.07 223.39 1.41 1.58 4.36 3.18 2.27 34.38 -1.60 0.42
DontCare -1 -1 -10 503.89 169.71 590.61 190.13 -1 -1 -1 -1000 -1000 -1000 -10 0.10
"""
KITTI_CALIB = """P0: 7.215377e+02 0.0 6.095593e+02 0.0 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P1: 7.215377e+02 0.0 6.095593e+02 -3.875744e+02 0.0 7.215377e+02 1.728540e+02 0.0 0.0 0.0 1.0 0.0
P2: 7.215377e+02 0.000000e+00 6.095593e+02 4.485728e+01 0.000000e+00 7.215377e+02 1.728540e+02 2.163791e-01 0.000000e+00 0.000000e+00 1.000000e+00 2.745884e-03
P3: 7.215377e+02 0.0 6.095593e+02 -3.395242e+02 0.0 7.215377e+02 1.728540e+02 2.199936e+00 0.0 0.0 1.0 2.729905e-03
R0_rect: 9.999239e-01 9.837760e-03 -7.445048e-03 -9.869795e-03 9.999421e-01 -4.278459e-03 7.402527e-03 4.351614e-03 9.999631e-01
"""


def golden_kitti_io(ref):
    """csv_read_annot / csv_read_calib of the reference's KITTI class (car_instance.py:792-842) on a label file with
    scores and a calibration file written here (KITTI text format)."""
    import tempfile
    import types
    import libs.dataset.KITTI.car_instance as CI
    with tempfile.TemporaryDirectory() as d:
        lp, cp = os.path.join(d, '000001.txt'), os.path.join(d, 'calib.txt')
        open(lp, 'w').write(KITTI_LABEL)
        open(cp, 'w').write(KITTI_CALIB)
        out = {}
        for tag, classes in (('car', ['Car']), ('all', ['Car', 'Pedestrian', 'Cyclist'])):
            fake = types.SimpleNamespace(_classes=classes)
            out[tag] = CI.KITTI.csv_read_annot(fake, lp, CI.FIELDNAMES_P)
        out['car_no_score'] = CI.KITTI.csv_read_annot(types.SimpleNamespace(_classes=['Car']), lp, CI.FIELDNAMES)
        P = CI.KITTI.csv_read_calib(types.SimpleNamespace(), cp)
    with open(os.path.join(HERE, 'kitti_io.json'), 'w') as f:
        json.dump({'label': KITTI_LABEL, 'calib': KITTI_CALIB, 'annots': out, 'P': P.tolist(),
                   'fieldnames': CI.FIELDNAMES, 'fieldnames_p': CI.FIELDNAMES_P, 'type_id': CI.TYPE_ID_CONVERSION}, f, indent=1)
    print('kitti_io.json')


def main():
    ref = import_reference()
    torch.set_num_threads(os.cpu_count())
    golden_hrnet(ref, 'tiny', configs.tiny_cfgs(), 3, 1, 0, 1)
    golden_hrnet(ref, 'tiny_heatmap', configs.tiny_cfgs('heatmap'), 2, 1, 0, 1)
    golden_hrnet(ref, 'ped', configs.ped_cfgs(), 1, 1, 0, 8)
    golden_hrnet(ref, 'demo', configs.demo_cfgs(), 2, 1, 0, 4)
    golden_decode(ref)
    golden_affine(ref)
    golden_lifter(ref)
    golden_pose(ref)
    golden_pipeline(ref)
    golden_pipeline(ref, 'demo', configs.demo_cfgs(), ['img_a.png'] * 3 + ['img_b.png'] * 4 + ['img_c.png'])
    golden_loss(ref)
    golden_crop(ref)
    golden_pnp(ref)
    golden_format(ref)
    golden_train(ref)
    golden_align(ref)
    golden_composite(ref)
    golden_train_coord(ref)
    golden_kitti_io(ref)
    import cv2, scipy
    with open(os.path.join(HERE, 'versions.json'), 'w') as f:
        json.dump({'torch': torch.__version__, 'numpy': np.__version__, 'scipy': scipy.__version__,
                   'cv2': cv2.__version__, 'python': sys.version.split()[0],
                   'reference': 'Nicholasli1995/EgoNet @ 13e3758 (/root/reference)'}, f, indent=1)


if __name__ == '__main__':
    main()
