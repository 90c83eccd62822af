"""KITTI prediction strings (host text I/O right after the hot path): the mirror of
libs/common/format.py against strings produced by the reference's own functions."""
import json
import os

import numpy as np

from egonet_b200.libs.common import format as fmt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'format.json')


def test_pred_str_matches_reference_golden(tmp_path):
    g = json.load(open(GOLDEN))
    assert [fmt.get_instance_str(r) for r in g['rows']] == g['instance_strs']
    record = {'raw_txt_format': g['rows'], 'euler_angles': np.array(g['euler_angles']), 'alphas': np.array(g['alphas'])}
    before = json.dumps(g['rows'])
    pred = fmt.get_pred_str(record)
    assert pred == g['pred_str']
    assert json.dumps(record['raw_txt_format']) == before          # the detector rows are not modified in place
    # only rot_y (= Euler y) and alpha are replaced, everything else passes through
    first = pred.split('\n')[0].split(' ')
    assert float(first[3]) == float('%.6f' % g['alphas'][0]) and float(first[14]) == float('%.6f' % g['euler_angles'][0][1])
    fmt.save_txt_file('/data/kitti/image_2/000123.png', {'pred_str': pred}, {'flag': True, 'save_dir': str(tmp_path)})
    assert open(tmp_path / '000123.txt').read() == g['pred_str']
    fmt.save_txt_file('/x/000124.png', {'pred_str': pred}, {'flag': False, 'save_dir': str(tmp_path)})
    assert not (tmp_path / '000124.txt').exists()


def test_mirror_crop_geometry_matches_reference_golden(golden):
    """Host-side crop geometry of the mirror (modify_bbox, get_affine_transform) against the reference's own
    outputs in affine.npz, and the pth_trans parser the device crop front-end relies on."""
    import pytest
    from egonet_b200.libs.common import img_proc
    g = golden('affine.npz')
    for i in range(len(g['boxes'])):
        ret = img_proc.modify_bbox(g['boxes'][i], g['ars'][i])
        np.testing.assert_array_equal(ret['c'], g['centers'][i])
        np.testing.assert_array_equal(ret['s'], g['scales'][i])
        np.testing.assert_allclose(ret['bbox'], g['bbox_resize'][i], rtol=0, atol=1e-12)
        res = (256, 256) if g['ars'][i] == 1.0 else (192, 256)
        np.testing.assert_allclose(img_proc.get_affine_transform(ret['c'], ret['s'], 0., (res[1], res[0])),
                                   g['trans_fwd'][i], rtol=0, atol=1e-10)
        np.testing.assert_allclose(img_proc.get_affine_transform(ret['c'], ret['s'], 0., (res[1], res[0]), inv=1),
                                   g['trans_inv'][i], rtol=0, atol=1e-9)

    class ToTensor:
        pass

    class Normalize:
        mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]

    class RandomHorizontalFlip:
        pass

    class Compose:
        def __init__(self, ts):
            self.transforms = ts

    assert img_proc.normalize_params(Compose([ToTensor(), Normalize()])) == (Normalize.mean, Normalize.std)
    assert img_proc.normalize_params(Compose([ToTensor()])) == ([0., 0., 0.], [1., 1., 1.])
    for bad in (None, Compose([Normalize()]), Compose([ToTensor(), RandomHorizontalFlip()])):
        with pytest.raises(NotImplementedError):
            img_proc.normalize_params(bad)


def test_add_orientation_arrow_vs_reference_golden(golden):
    """EgoNet.add_orientation_arrow (egonet.py:157-179), host-side record field, vectorised here."""
    from egonet_b200.libs.model.egonet import EgoNet
    from oracle.egonet_ref import KITTI_K
    g = golden('align.npz')
    ego = EgoNet.__new__(EgoNet)
    got = ego.add_orientation_arrow({'kpts_3d_pred': g['arrow_pred'], 'kpts_3d_gt': g['arrow_gt'], 'K': KITTI_K})
    lengths = np.linalg.norm(g['arrow'][:, :, 1] - g['arrow'][:, :, 0], axis=1)
    assert (lengths < 50).any() and np.isclose(lengths, 60).any()          # both branches present
    np.testing.assert_allclose(got, g['arrow'], rtol=0, atol=1e-9)


def test_modify_bbox_batch_rows_equal_per_box_function():
    """The vectorised crop geometry used by EgoNet.crop_instances equals modify_bbox (img_proc.py:411-459, itself
    pinned to the reference golden above) row by row, bit for bit, for float64 / float32 / integer boxes."""
    from egonet_b200.libs.common import img_proc as lip
    rng = np.random.Generator(np.random.PCG64(3))
    x0, y0 = rng.uniform(0, 1100, 200), rng.uniform(0, 300, 200)
    boxes = np.stack([x0, y0, x0 + rng.uniform(5, 400, 200), y0 + rng.uniform(5, 250, 200)], 1)
    for arr in (boxes, boxes.astype(np.float32), boxes.astype(np.int64)):
        for ar in (1.0, 256 / 192):
            box, c, s = lip.modify_bbox_batch(arr, ar)
            for i in range(len(arr)):
                ref = lip.modify_bbox(arr[i], ar)
                assert np.array_equal(np.asarray(ref['bbox'], dtype=box.dtype), box[i])
                assert np.array_equal(ref['c'], c[i]) and np.array_equal(ref['s'], s[i])
