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
