"""Crop front-end on the CPU: (1) the oracle restatement of cv2.warpAffine + ToTensor/Normalize against
the golden crops produced by the reference's own crop_single_instance, (2) the kernel's arithmetic
(egonet_b200/csrc/crop_math.h compiled for the host) bit for bit against the same goldens."""
import ctypes
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import affine_ref, crop_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'tests', 'native', 'libcrop_host.so')
RES = {'sq': (256, 256), 'ped': (192, 256)}


def sha1(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.fixture(scope='module')
def image(golden):
    g = golden('crop.npz')
    img = crop_ref.synth_image(int(g['image_shape'][0]), int(g['image_shape'][1]), int(g['image_seed']))
    np.testing.assert_array_equal(sha1(img), g['image_sha1'])
    return img


@pytest.fixture(scope='module')
def host():
    src = os.path.join(ROOT, 'tests', 'native', 'crop_host.cpp')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'egonet_b200', 'csrc'),
                           src, '-o', SO])
    return ctypes.CDLL(SO)


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def host_crops(L, img, centers, scales, res, mean, std):
    n = len(centers)
    out = np.zeros((n, 3, res[1], res[0]), np.float32)
    u8 = np.zeros((n, res[1], res[0], 3), np.uint8)
    ce, sc = np.ascontiguousarray(centers, np.float64), np.ascontiguousarray(scales, np.float64)
    m, s = np.asarray(mean, np.float32), np.asarray(std, np.float32)
    L.host_crop_instances(vp(img), img.shape[0], img.shape[1], img.strides[0], vp(ce), vp(sc), n,
                          res[0], res[1], vp(m), vp(s), vp(out), vp(u8))
    return out, u8


def test_bilinear_table_closed_form(host):
    """OpenCV's BilinearTab_i restated with its float arithmetic == the kernel's integer closed form."""
    tab = crop_ref.bilinear_tab()
    w = (ctypes.c_int * 4)()
    for fy in range(32):
        for fx in range(32):
            host.host_bilinear_weights(fx, fy, w)
            assert list(w) == tab[fy, fx].reshape(-1).tolist()
            assert sum(w) == 32768
    assert tab[0, 0].reshape(-1).tolist() == [32767, 0, 0, 1]


@pytest.mark.parametrize('tag', ['sq', 'ped'])
def test_oracle_crop_vs_reference_golden(golden, image, tag):
    g = golden('crop.npz')
    res = RES[tag]
    full = {int(i): k for k, i in enumerate(g[tag + '_u8_index'])}
    for i, box in enumerate(g['boxes']):
        crop, norm, c, s = crop_ref.crop_single_instance(image, box, res, g['mean'], g['std'])
        np.testing.assert_array_equal(c, g[tag + '_centers'][i])
        np.testing.assert_array_equal(s, g[tag + '_scales'][i])
        if i in full:
            np.testing.assert_array_equal(crop, g[tag + '_u8'][full[i]])
        np.testing.assert_array_equal(sha1(crop), g[tag + '_u8_sha1'][i])
        np.testing.assert_array_equal(norm[:, ::4, ::4], g[tag + '_norm_sub'][i])       # bit-exact fp32
        assert norm.astype(np.float64).sum() == pytest.approx(g[tag + '_norm_sum'][i], abs=1e-6)
    assert not crop_ref.crop_single_instance(image, g['boxes'][8], res)[0].any()          # box outside the image


@pytest.mark.parametrize('tag', ['sq', 'ped'])
def test_kernel_source_on_host_vs_reference_golden(golden, image, host, tag):
    g = golden('crop.npz')
    res = RES[tag]
    out, u8 = host_crops(host, image, g[tag + '_centers'], g[tag + '_scales'], res, g['mean'], g['std'])
    for k, i in enumerate(g[tag + '_u8_index']):
        np.testing.assert_array_equal(u8[i], g[tag + '_u8'][k])
    for i in range(len(g['boxes'])):
        np.testing.assert_array_equal(sha1(u8[i]), g[tag + '_u8_sha1'][i])
        np.testing.assert_array_equal(out[i][:, ::4, ::4], g[tag + '_norm_sub'][i])
        np.testing.assert_array_equal(out[i], crop_ref.to_tensor_normalize(u8[i], g['mean'], g['std']))


def test_kernel_source_on_host_vs_oracle_random_boxes(host):
    """Seeded KITTI-shaped boxes (the bench's distribution) on a second image, odd output sizes included."""
    img = crop_ref.synth_image(120, 400, 5)
    rng = np.random.Generator(np.random.PCG64(9))
    for res in ((64, 64), (48, 64), (50, 38)):
        n = 24
        cx, cy = rng.uniform(-20, 420, n), rng.uniform(-10, 130, n)
        w, h = rng.uniform(6, 300, n), rng.uniform(5, 120, n)
        boxes = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
        rets = [affine_ref.modify_bbox(b, res[1] / res[0]) for b in boxes]
        ce, sc = np.array([r['c'] for r in rets]), np.array([r['s'] for r in rets])
        out, u8 = host_crops(host, img, ce, sc, res, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
        for i, b in enumerate(boxes):
            crop, norm, _, _ = crop_ref.crop_single_instance(img, b, res, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
            np.testing.assert_array_equal(u8[i], crop)
            np.testing.assert_array_equal(out[i], norm)


def test_forward_affine_closed_form_vs_oracle(host, golden):
    """The kernel's closed-form 3-point affine against the oracle's 6x6 solve (what cv2.getAffineTransform
    does) and the reference's own matrices in affine.npz."""
    g = golden('affine.npz')
    M = (ctypes.c_double * 6)()
    host.host_forward_crop_affine.argtypes = [ctypes.c_double] * 3 + [ctypes.c_int] * 2 + [ctypes.c_void_p]
    for i in range(len(g['centers'])):
        res = (256, 256) if g['ars'][i] == 1.0 else (192, 256)
        host.host_forward_crop_affine(g['centers'][i][0], g['centers'][i][1], g['scales'][i][0], res[0], res[1], M)
        np.testing.assert_allclose(np.array(M).reshape(2, 3), g['trans_fwd'][i], rtol=0, atol=1e-11)
