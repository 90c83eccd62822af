"""SURVEY.md 8f row 4 on the CPU: KITTI label / calibration readers against the reference class's output
(tests/golden/kitti_io.json), and the rotated-box overlap math -- the oracle (oracle/eval_ref.py) against closed-form
cases, then the host-compiled kernel source (csrc/eval_math.h) against the oracle on seeded boxes."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import eval_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kitti_label_and_calib_readers_vs_reference_golden(tmp_path):
    from egonet_b200.libs.dataset.KITTI import car_instance as ci
    g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'kitti_io.json')))
    assert ci.FIELDNAMES == g['fieldnames'] and ci.FIELDNAMES_P == g['fieldnames_p'] and ci.TYPE_ID_CONVERSION == g['type_id']
    lp, cp = tmp_path / '000001.txt', tmp_path / 'calib.txt'
    lp.write_text(g['label'])
    cp.write_text(g['calib'])
    assert ci.csv_read_annot(str(lp), ci.FIELDNAMES_P, ('Car',)) == g['annots']['car']
    assert ci.csv_read_annot(str(lp), ci.FIELDNAMES_P, ('Car', 'Pedestrian', 'Cyclist')) == g['annots']['all']
    assert ci.csv_read_annot(str(lp), ci.FIELDNAMES, ('Car',)) == g['annots']['car_no_score']
    P = ci.csv_read_calib(str(cp))
    assert P.dtype == np.float32 and np.array_equal(P, np.array(g['P'], dtype=np.float32))
    rows, P2 = ci.load_annotations(str(lp), str(cp), ci.FIELDNAMES_P)
    assert len(rows) == 3 and np.array_equal(P2, P)
    meta = ci.annot_dict_for_inference(['a/000001.png'], [str(lp)], [str(cp)])
    assert meta['boxes'][0].shape == (3, 4) and meta['K'][0].shape == (3, 3) and meta['raw_txt_format'][0] == rows
    from egonet_b200.libs.common.img_proc import modify_bbox
    np.testing.assert_array_equal(meta['boxes'][0][1], np.array(modify_bbox(np.array(rows[1]['bbox']), 1.0, 1.2)['bbox']))
    from egonet_b200.libs.metric.kitti_eval import boxes_from_annotations
    b = boxes_from_annotations(rows)
    assert b.shape == (3, 7) and b[0].tolist() == [-1.59, 1.65, 1.67, 3.64, -0.65, 1.71, 46.70]


def _box(ry=0.0, h=1.5, w=2.0, l=4.0, x=0.0, y=1.5, z=10.0):
    return np.array([ry, h, w, l, x, y, z])


def test_oracle_overlaps_closed_form_cases():
    a = _box()
    assert eval_ref.ground_box_overlap(a, a) == pytest.approx(1.0, abs=1e-12)
    assert eval_ref.box3d_overlap(a, a) == pytest.approx(1.0, abs=1e-12)
    # axis aligned, shifted by 1 along x (length axis): inter 3 x 2, union 8 + 8 - 6
    b = _box(x=1.0)
    assert eval_ref.ground_box_overlap(a, b) == pytest.approx(6 / 10, abs=1e-12)
    assert eval_ref.ground_box_overlap(a, b, 0) == pytest.approx(6 / 8, abs=1e-12)
    # vertical offset 0.5: height overlap 1.0 of 1.5
    c = _box(x=1.0, y=2.0)
    assert eval_ref.box3d_overlap(a, c) == pytest.approx(6 * 1.0 / (12 + 12 - 6), abs=1e-12)
    # ry = pi/2 swaps length and width: 4x2 against 2x4 about the same centre: inter 2 x 2
    d = _box(ry=np.pi / 2)
    assert eval_ref.ground_box_overlap(a, d) == pytest.approx(4 / 12, abs=1e-12)
    # a square against its 45-degree rotation: regular octagon, area 2 (sqrt 2 - 1) s^2
    s1, s2 = _box(w=2.0, l=2.0), _box(w=2.0, l=2.0, ry=np.pi / 4)
    inter = 2 * (np.sqrt(2) - 1) * 4
    assert eval_ref.ground_box_overlap(s1, s2) == pytest.approx(inter / (8 - inter), abs=1e-12)
    # containment and disjoint boxes; criterion 1 = / ground-truth area
    small = _box(w=1.0, l=1.0, ry=0.3)
    assert eval_ref.ground_box_overlap(small, a, 0) == pytest.approx(1.0, abs=1e-12)
    assert eval_ref.ground_box_overlap(small, a, 1) == pytest.approx(1.0 / 8.0, abs=1e-12)
    assert eval_ref.ground_box_overlap(a, _box(x=50.0)) == 0.0
    assert eval_ref.box3d_overlap(a, _box(y=5.0)) == 0.0            # same footprint, no vertical overlap
    assert eval_ref.image_box_overlap([0, 0, 10, 10], [5, 5, 15, 15]) == pytest.approx(25 / 175)
    assert eval_ref.image_box_overlap([0, 0, 10, 10], [10, 0, 20, 10]) == 0.0


def test_oracle_intersection_area_vs_monte_carlo():
    rng = np.random.Generator(np.random.PCG64(3))
    boxes = eval_ref.synth_boxes(6, 11)
    pts = rng.uniform([-8, 6], [8, 22], (400000, 2))
    cell = 16 * 16 / len(pts)

    def inside(p, poly):
        d = np.roll(poly, -1, axis=0) - poly
        cr = d[None, :, 0] * (p[:, None, 1] - poly[None, :, 1]) - d[None, :, 1] * (p[:, None, 0] - poly[None, :, 0])
        return np.all(cr >= 0, axis=1) | np.all(cr <= 0, axis=1)
    for i in range(3):
        pa, pb = eval_ref.ground_polygon(boxes[2 * i]), eval_ref.ground_polygon(boxes[2 * i + 1])
        mc = (inside(pts, pa) & inside(pts, pb)).sum() * cell
        assert eval_ref.convex_intersection_area(pa, pb) == pytest.approx(mc, abs=0.08)


@pytest.fixture(scope='module')
def host():
    so = os.path.join(ROOT, 'tests', 'native', 'libeval_host.so')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(ROOT, 'egonet_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'native', 'eval_host.cpp'), '-o', so])
    return ctypes.CDLL(so)


def dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('criterion', [-1, 0, 1])
def test_kernel_math_vs_oracle(host, criterion):
    det, gt = eval_ref.synth_boxes(24, 5), eval_ref.synth_boxes(17, 6)
    det[3], det[4] = gt[2], gt[5]                                    # identical boxes
    det[5, 0] = 0.0
    gt[7] = det[5] + np.array([0, 0, 0, 0, 1.0, 0.2, 0])            # parallel edges (degenerate crossings)
    ground, box3d = np.zeros((24, 17)), np.zeros((24, 17))
    host.host_box_overlaps(dp(det), dp(gt), 24, 17, criterion, dp(ground), dp(box3d))
    ref_g = np.array([[eval_ref.ground_box_overlap(d, g, criterion) for g in gt] for d in det])
    ref_b = np.array([[eval_ref.box3d_overlap(d, g, criterion) for g in gt] for d in det])
    assert (ref_g > 0).mean() > 0.3 and (ref_g == 0).any()
    np.testing.assert_allclose(ground, ref_g, rtol=0, atol=1e-10)
    np.testing.assert_allclose(box3d, ref_b, rtol=0, atol=1e-10)
    rng = np.random.Generator(np.random.PCG64(9))
    a = np.sort(rng.uniform(0, 100, (9, 2, 2)), axis=1).transpose(0, 2, 1).reshape(9, 4)[:, [0, 2, 1, 3]]
    b = np.sort(rng.uniform(0, 100, (7, 2, 2)), axis=1).transpose(0, 2, 1).reshape(7, 4)[:, [0, 2, 1, 3]]
    out = np.zeros((9, 7))
    host.host_image_overlaps(dp(np.ascontiguousarray(a)), dp(np.ascontiguousarray(b)), 9, 7, criterion, dp(out))
    np.testing.assert_allclose(out, [[eval_ref.image_box_overlap(x, y, criterion) for y in b] for x in a], atol=1e-12)
