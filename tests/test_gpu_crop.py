"""GPU parity of the crop front-end (``egn_crop_instances`` through the reference-shaped mirror):
bit-exact uint8 crops and bit-exact fp32 normalised tensors against the goldens made by the reference's
own ``crop_single_instance`` (cv2.warpAffine + torchvision) and against the CPU oracle."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import affine_ref, configs, crop_ref, egonet_ref, hrnet_ref, lifter_ref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from egonet_b200.libs.common import img_proc
    from egonet_b200.libs.model.egonet import EgoNet

DEV = 'cuda'
RES = {'sq': (256, 256), 'ped': (192, 256)}
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def sha1(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


class _Compose:
    """Duck-typed stand-in for torchvision's Compose([ToTensor(), Normalize(mean, std)]) so that the test
    does not need torchvision on the GPU box."""
    class ToTensor:
        pass

    class Normalize:
        def __init__(self, mean, std):
            self.mean, self.std = mean, std

    def __init__(self, mean, std):
        self.transforms = [self.ToTensor(), self.Normalize(mean, std)]


@pytest.mark.parametrize('tag', ['sq', 'ped'])
def test_crop_kernel_bit_exact_vs_reference_golden(golden, tag):
    g = golden('crop.npz')
    res = RES[tag]
    img = crop_ref.synth_image(int(g['image_shape'][0]), int(g['image_shape'][1]), int(g['image_seed']))
    np.testing.assert_array_equal(sha1(img), g['image_sha1'])
    n = len(g['boxes'])
    out, u8 = img_proc.crop_instances_device([torch.from_numpy(img).to(DEV)], [0] * n, g[tag + '_centers'],
                                             g[tag + '_scales'], res, g['mean'], g['std'], return_u8=True)
    out, u8 = out.cpu().numpy(), u8.cpu().numpy()
    for k, i in enumerate(g[tag + '_u8_index']):
        np.testing.assert_array_equal(u8[i], g[tag + '_u8'][k])                     # uint8: bit-exact
    for i in range(n):
        np.testing.assert_array_equal(sha1(u8[i]), g[tag + '_u8_sha1'][i])
        np.testing.assert_array_equal(out[i][:, ::4, ::4], g[tag + '_norm_sub'][i])  # fp32: bit-exact
        np.testing.assert_array_equal(out[i], crop_ref.to_tensor_normalize(u8[i], g['mean'], g['std']))
    assert not u8[8].any()                                                           # box outside the image


def test_crop_kernel_vs_oracle_multi_image_ragged():
    """Three images of different sizes (one a strided view), ragged boxes per image, odd output sizes
    (scalar store path), N = 0."""
    rng = np.random.Generator(np.random.PCG64(17))
    imgs = [crop_ref.synth_image(120, 400, 5), crop_ref.synth_image(75, 90, 6), crop_ref.synth_image(200, 333, 7)]
    wide = torch.zeros((75, 128, 3), dtype=torch.uint8, device=DEV)
    wide[:, :90] = torch.from_numpy(imgs[1]).to(DEV)
    dev_imgs = [torch.from_numpy(imgs[0]).to(DEV), wide[:, :90], torch.from_numpy(imgs[2]).to(DEV)]
    counts = [5, 1, 9]
    for res in ((64, 64), (50, 38), (48, 64)):
        which, boxes = [], []
        for k, c in enumerate(counts):
            h, w = imgs[k].shape[:2]
            cx, cy = rng.uniform(-10, w + 10, c), rng.uniform(-10, h + 10, c)
            bw, bh = rng.uniform(5, w, c), rng.uniform(5, h, c)
            boxes += list(np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], 1))
            which += [k] * c
        rets = [affine_ref.modify_bbox(b, res[1] / res[0]) for b in boxes]
        ce, sc = np.array([r['c'] for r in rets]), np.array([r['s'] for r in rets])
        out, u8 = img_proc.crop_instances_device(dev_imgs, which, ce, sc, res, MEAN, STD, return_u8=True)
        out, u8 = out.cpu().numpy(), u8.cpu().numpy()
        for i, b in enumerate(boxes):
            crop, norm, _, _ = crop_ref.crop_single_instance(imgs[which[i]], b, res, MEAN, STD)
            np.testing.assert_array_equal(u8[i], crop)
            np.testing.assert_array_equal(out[i], norm)
    empty = img_proc.crop_instances_device(dev_imgs, [], np.zeros((0, 2)), np.zeros((0, 2)), (64, 64), MEAN, STD)
    assert empty.shape == (0, 3, 64, 64)
    with pytest.raises(ValueError):
        img_proc.crop_instances_device(dev_imgs, [3], np.zeros((1, 2)), np.ones((1, 2)), (64, 64))


def test_crop_kernel_full_size_properties():
    """BASELINE batch size (256 crops of 256x256 from one KITTI-sized image): an axis-aligned integer
    box whose side equals the output size is a plain copy of the source window; duplicate boxes give
    identical crops; the crops equal the oracle on a sample."""
    img = crop_ref.synth_image(375, 1242, 33)
    cfgs = configs.demo_cfgs()
    recs = egonet_ref.synth_boxes(256, cfgs, 2)
    ce = np.array([r['center'] for r in recs])
    sc = np.array([r['scale'] for r in recs])
    ce[0], sc[0] = (600.0, 180.0), (256 / 200.0, 256 / 200.0)          # identity-scale window [472,728) x [52,308)
    ce[1], sc[1] = ce[100], sc[100]
    out, u8 = img_proc.crop_instances_device([torch.from_numpy(img).to(DEV)], [0] * 256, ce, sc, (256, 256),
                                             MEAN, STD, return_u8=True)
    u8 = u8.cpu().numpy()
    np.testing.assert_array_equal(u8[0], img[52:308, 472:728])
    np.testing.assert_array_equal(u8[1], u8[100])
    assert torch.equal(out[1], out[100])
    for i in (2, 77, 255):
        M = affine_ref.get_affine_transform(ce[i], sc[i], 0., (256, 256))
        np.testing.assert_array_equal(u8[i], crop_ref.warp_affine(img, M, (256, 256)))


def test_egonet_forward_from_images_matches_forward_crops():
    """EgoNet.forward (crop_instances on the device -> HC -> lifter) produces the records that
    forward_crops gives for the oracle's crops of the same boxes."""
    cfgs = configs.tiny_cfgs()
    ego = EgoNet(cfgs, pre_trained=False).eval()
    ego.HC.load_state_dict(hrnet_ref.make_weights(cfgs, 1))
    ego.L.load_state_dict(lifter_ref.make_weights(cfgs, 11))
    ego.LS = lifter_ref.make_stats(cfgs, 12)
    ego = ego.cuda()
    ego.pth_trans = _Compose(MEAN, STD)
    res = ego.resolution
    imgs = [crop_ref.synth_image(120, 400, 5), crop_ref.synth_image(100, 300, 6)]
    boxes = [np.array([[30., 20., 150., 90.], [200., 10., 390., 110.]]), np.array([[50., 30., 120., 80.]])]
    annot = {'path': ['a.png', 'b.png'], 'images': imgs, 'boxes': boxes}
    crops, recs = ego.crop_instances(annot, res, pth_trans=ego.pth_trans)
    assert crops.is_cuda and crops.shape == (3, 3, res[1], res[0]) and len(recs) == 3
    k = 0
    for im, bs in zip(imgs, boxes):
        for b in bs:
            _, norm, c, s = crop_ref.crop_single_instance(im, b, res, MEAN, STD)
            np.testing.assert_array_equal(crops[k].cpu().numpy(), norm)
            np.testing.assert_array_equal(recs[k]['center'], c)
            np.testing.assert_array_equal(recs[k]['scale'], s)
            k += 1
    records = ego(annot)
    pose = ego.forward_crops(crops, np.array([r['center'] for r in recs]), np.array([r['scale'] for r in recs]),
                             return_all=True)
    k2 = pose['kpts_2d'].cpu().numpy()
    np.testing.assert_array_equal(np.concatenate(records['a.png']['kpts_2d_pred']), k2[:2])
    np.testing.assert_array_equal(np.concatenate(records['b.png']['kpts_2d_pred']), k2[2:])
    # single-instance entry: uint8 without pth_trans (as upstream), tensor with it
    raw = ego.crop_single_instance(imgs[0], boxes[0][0], res)
    np.testing.assert_array_equal(raw, crop_ref.crop_single_instance(imgs[0], boxes[0][0], res)[0])
    one = ego.crop_single_instance(imgs[0], boxes[0][0], res, pth_trans=ego.pth_trans)
    assert torch.equal(one, crops[0])
