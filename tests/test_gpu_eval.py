"""GPU parity of the batched KITTI evaluation overlaps (``egn_box_overlaps`` / ``egn_image_box_overlaps`` through
``libs/metric/kitti_eval.py``) against the CPU oracle (closed-form-anchored restatement of
tools/kitti-eval/evaluate_object_3d_offline.cpp:224-344; the reference binary needs Boost and cannot be built here)."""
import numpy as np
import pytest
import torch

from oracle import eval_ref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from egonet_b200.libs.metric import kitti_eval


@pytest.mark.parametrize('criterion', [-1, 0, 1])
def test_box_overlaps_vs_oracle(criterion):
    det, gt = eval_ref.synth_boxes(40, 5), eval_ref.synth_boxes(23, 6)
    det[3], det[4] = gt[2], gt[5]
    out = kitti_eval.box_overlaps(det, gt, criterion)
    ref_g = np.array([[eval_ref.ground_box_overlap(d, g, criterion) for g in gt] for d in det])
    ref_b = np.array([[eval_ref.box3d_overlap(d, g, criterion) for g in gt] for d in det])
    np.testing.assert_allclose(out['ground'].cpu().numpy(), ref_g, rtol=0, atol=1e-10)      # fp64, 1e-10
    np.testing.assert_allclose(out['box3d'].cpu().numpy(), ref_b, rtol=0, atol=1e-10)
    assert out['ground'][3, 2].item() == pytest.approx(1.0, abs=1e-12)


def test_overlap_properties_and_edge_cases():
    det = eval_ref.synth_boxes(300, 7)
    o = kitti_eval.box_overlaps(det, det)
    g, b = o['ground'], o['box3d']
    assert torch.allclose(torch.diagonal(g), torch.ones(300, dtype=torch.float64, device=g.device), atol=1e-12)
    assert torch.allclose(g, g.T, atol=1e-12) and torch.allclose(b, b.T, atol=1e-12)      # union criterion is symmetric
    assert float(g.min()) >= 0 and float(g.max()) <= 1 + 1e-12 and bool((b <= g + 1e-12).all())
    e = kitti_eval.box_overlaps(np.zeros((0, 7)), det)
    assert e['ground'].shape == (0, 300)
    only = kitti_eval.box_overlaps(det[:5], det[:7], want=('box3d',))
    assert set(only) == {'box3d'} and torch.equal(only['box3d'], b[:5, :7])
    a = np.array([[0., 0, 10, 10]])
    c = np.array([[5., 5, 15, 15], [10, 0, 20, 10]])
    np.testing.assert_allclose(kitti_eval.image_box_overlaps(a, c).cpu().numpy(), [[25 / 175, 0.0]], atol=1e-12)
