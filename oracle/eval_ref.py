"""CPU restatement of the KITTI evaluation overlaps (TEST INFRASTRUCTURE).

``ground_box_overlap`` / ``box3d_overlap`` / ``image_box_overlap`` follow ``groundBoxOverlap`` / ``box3DOverlap`` /
``imageBoxOverlap`` of ``tools/kitti-eval/evaluate_object_3d_offline.cpp:224-344`` (``toPolygon`` ``:266-290``).

PARITY UNPINNED against the reference binary: the evaluator is C++ on Boost.Geometry / Boost.uBLAS, Boost is absent
from this image and the file cannot be built from its own sources alone.  Boost's ``intersection`` of two convex
polygons is restated as half-plane clipping (numpy, independent of the kernel's Sutherland-Hodgman code: here the
intersection polygon is built from the vertices of each rectangle inside the other plus all edge-edge crossings, sorted
by angle), and anchored on closed-form cases (tests/test_eval_cpu.py): axis-aligned rectangles, a square against its
45-degree rotation (area 2 (sqrt 2 - 1) s^2), containment, disjoint boxes, criterion 0 / 1 normalisation.
"""
import numpy as np


def ground_polygon(box):
    """box = (ry, h, w, l, t1, t2, t3) -> [4,2] corners (x, z) in the evaluator's order."""
    ry, _, w, l, t1, _, t3 = box
    c, s = np.cos(ry), np.sin(ry)
    loc = np.array([[l / 2, l / 2, -l / 2, -l / 2], [w / 2, -w / 2, -w / 2, w / 2]])
    R = np.array([[c, s], [-s, c]])
    return (R @ loc).T + np.array([t1, t3])


def _area(p):
    x, y = p[:, 0], p[:, 1]
    return 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def _inside(pt, poly):
    """pt inside (or on) the convex polygon, either orientation."""
    d = np.roll(poly, -1, axis=0) - poly
    cr = d[:, 0] * (pt[1] - poly[:, 1]) - d[:, 1] * (pt[0] - poly[:, 0])
    return np.all(cr >= -1e-12) or np.all(cr <= 1e-12)


def convex_intersection_area(a, b):
    pts = [p for p in a if _inside(p, b)] + [p for p in b if _inside(p, a)]
    for i in range(len(a)):
        p, r = a[i], a[(i + 1) % len(a)] - a[i]
        for j in range(len(b)):
            q, s = b[j], b[(j + 1) % len(b)] - b[j]
            den = r[0] * s[1] - r[1] * s[0]
            if abs(den) < 1e-15:
                continue
            t = ((q[0] - p[0]) * s[1] - (q[1] - p[1]) * s[0]) / den
            u = ((q[0] - p[0]) * r[1] - (q[1] - p[1]) * r[0]) / den
            if 0 <= t <= 1 and 0 <= u <= 1:
                pts.append(p + t * r)
    if len(pts) < 3:
        return 0.0
    pts = np.array(pts)
    c = pts.mean(0)
    order = np.argsort(np.arctan2(pts[:, 1] - c[1], pts[:, 0] - c[0]))
    return _area(pts[order])


def _ratio(inter, a, b, criterion):
    den = a + b - inter if criterion == -1 else (a if criterion == 0 else b)
    return inter / den if den > 0 else 0.0


def ground_box_overlap(d, g, criterion=-1):
    dp, gp = ground_polygon(d), ground_polygon(g)
    return _ratio(convex_intersection_area(gp, dp), _area(dp), _area(gp), criterion)


def box3d_overlap(d, g, criterion=-1):
    dp, gp = ground_polygon(d), ground_polygon(g)
    inter = convex_intersection_area(gp, dp)
    ymax, ymin = min(d[5], g[5]), max(d[5] - d[1], g[5] - g[1])
    vol = inter * max(0.0, ymax - ymin)
    return _ratio(vol, d[1] * d[3] * d[2], g[1] * g[3] * g[2], criterion)


def image_box_overlap(a, b, criterion=-1):
    w, h = min(a[2], b[2]) - max(a[0], b[0]), min(a[3], b[3]) - max(a[1], b[1])
    if w <= 0 or h <= 0:
        return 0.0
    return _ratio(w * h, (a[2] - a[0]) * (a[3] - a[1]), (b[2] - b[0]) * (b[3] - b[1]), criterion)


def synth_boxes(n, seed):
    """KITTI-like cars: (ry, h, w, l, x, y, z) clustered so that many pairs overlap."""
    g = np.random.Generator(np.random.PCG64(seed))
    return np.stack([g.uniform(-np.pi, np.pi, n), g.uniform(1.3, 1.9, n), g.uniform(1.5, 2.0, n), g.uniform(3.2, 4.8, n),
                     g.uniform(-4, 4, n), g.uniform(1.2, 2.0, n), g.uniform(10, 18, n)], 1)
