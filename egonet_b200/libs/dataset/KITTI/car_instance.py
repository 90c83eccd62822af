"""KITTI label / calibration text I/O of the reference's ``libs/dataset/KITTI/car_instance.py``:
the readers that sit right before the hot path (detector rows -> boxes) and right after it (prediction files).
Host text parsing, as upstream; the rest of that module (dataset class, augmentation, plotting) is out of scope.
"""
import csv

import numpy as np

# [car_instance.py:35-59]
TYPE_ID_CONVERSION = {'Car': 0, 'Cyclist': 1, 'Pedestrian': 2}
FIELDNAMES = ['type', 'truncated', 'occluded', 'alpha', 'xmin', 'ymin', 'xmax', 'ymax', 'dh', 'dw', 'dl', 'lx', 'ly',
              'lz', 'ry']
FIELDNAMES_P = FIELDNAMES.copy() + ['score']


def csv_read_annot(file_path, fieldnames=FIELDNAMES, classes=('Car',)):
    """[car_instance.py:792-828] one dictionary per instance of the selected classes, KITTI label format."""
    annotations = []
    with open(file_path, 'r') as csv_file:
        for row in csv.DictReader(csv_file, delimiter=' ', fieldnames=fieldnames):
            if row['type'] not in classes:
                continue
            annot = {'class': row['type'], 'label': TYPE_ID_CONVERSION[row['type']],
                     'truncation': float(row['truncated']), 'occlusion': float(row['occluded']),
                     'alpha': float(row['alpha']),
                     'dimensions': [float(row['dl']), float(row['dh']), float(row['dw'])],
                     'locations': [float(row['lx']), float(row['ly']), float(row['lz'])],
                     'rot_y': float(row['ry']),
                     'bbox': [float(row['xmin']), float(row['ymin']), float(row['xmax']), float(row['ymax'])]}
            if 'score' in fieldnames:
                annot['score'] = float(row['score'])
            annotations.append(annot)
    return annotations


def csv_read_calib(file_path):
    """[car_instance.py:830-842] the P2 projection matrix, float32 [3,4]."""
    with open(file_path, 'r') as csv_file:
        for row in csv.reader(csv_file, delimiter=' '):
            if row and row[0] == 'P2:':
                return np.array([float(v) for v in row[1:]], dtype=np.float32).reshape(3, 4)
    raise ValueError('no P2 row in {}'.format(file_path))


def load_annotations(label_path, calib_path, fieldnames=FIELDNAMES, classes=('Car',)):
    """[car_instance.py:844-853] (instances, P)."""
    return csv_read_annot(label_path, fieldnames, classes), csv_read_calib(calib_path)


def annot_dict_for_inference(image_paths, label_paths, calib_paths, fieldnames=FIELDNAMES_P, classes=('Car',),
                             larger=True, target_ar=1.0, enlarge=1.2):
    """The dictionary ``tools/inference.py::gather_dict`` [inference.py:86-127] hands to ``model(meta)``, built
    straight from detector label files + calibration files (upstream goes through ``read_single_file``
    [car_instance.py:383-449] with ``use_raw_bbox``): per image the boxes (enlarged with ``modify_bbox`` as
    :108-115 does), the detector rows (``raw_txt_format``), the scores and the intrinsics ``K = P2[:, :3]``."""
    from ...common.img_proc import modify_bbox_batch
    meta = {'path': [], 'boxes': [], 'raw_txt_format': [], 'K': [], 'scores': []}
    for img, lab, cal in zip(image_paths, label_paths, calib_paths):
        rows, P = load_annotations(lab, cal, fieldnames, classes)
        if not rows:
            continue
        meta['path'].append(img)
        boxes = np.array([r['bbox'] for r in rows], dtype=np.float64).reshape(-1, 4)
        meta['boxes'].append(modify_bbox_batch(boxes, target_ar, enlarge)[0] if larger else boxes)
        meta['raw_txt_format'].append(rows)
        meta['K'].append(P[:, :3])
        meta['scores'].append(np.array([r.get('score', -1.0) for r in rows]))
    return meta
