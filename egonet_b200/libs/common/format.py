"""KITTI-style prediction strings (host text I/O just after the hot path).

Mirrors the behaviour of the reference's ``libs/common/format.py``
(``get_instance_str`` :25-42, ``get_pred_str`` :44-61, ``save_txt_file`` :63-74):
the detector's label fields pass through unchanged, only ``rot_y`` (= Euler y)
and ``alpha`` are replaced by Ego-Net's estimates.
"""
import copy
import os

_ORDER = (('truncation', 1), ('occlusion', 1), ('alpha', 6))


def get_instance_str(dic):
    fields = [dic['class']]
    fields += ['{:.{p}f}'.format(dic[k], p=p) for k, p in _ORDER]
    fields += ['{:.6f}'.format(v) for v in dic['bbox'][:4]]
    d = dic['dimensions']
    fields += ['{:.6f}'.format(v) for v in (d[1], d[2], d[0])]
    fields += ['{:.6f}'.format(v) for v in dic['locations'][:3]]
    fields.append('{:.6f}'.format(dic['rot_y']))
    fields.append('{:.8f}'.format(dic.get('score', 1.0)))
    return ' '.join(fields) + ' '


def get_pred_str(record):
    rows = copy.deepcopy(record['raw_txt_format'])
    n = len(record['euler_angles'])
    for i in range(n):
        rows[i]['rot_y'] = record['euler_angles'][i, 1]
        rows[i]['alpha'] = record['alphas'][i]
    return '\n'.join(get_instance_str(rows[i]) for i in range(n))


def save_txt_file(img_path, prediction, params):
    if not params['flag']:
        return
    save_path = os.path.join(params['save_dir'], img_path.split('/')[-1][:-3] + 'txt')
    with open(save_path, 'w') as f:
        f.write(prediction['pred_str'])
    print('Wrote prediction file at {:s}'.format(save_path))
