"""B200-native drop-in for the reference's ``libs/model/egonet.py`` (class ``EgoNet``).

Same attributes (``HC``, ``L``, ``LS``, ``resolution``, ``xy_dict``, ``pth_trans``)
and methods as upstream [egonet.py:28-507]; the per-instance / per-image Python
loops of the reference are replaced by batched device calls:

  get_keypoints      HC forward -> coordinate head -> inverse crop affine, all on
                     the device; ONE D2H copy of the [N,33,2] screen key-points
  lift_2d_to_3d      one fused lifter chain over every instance of the batch
                     (upstream: one launch train + 2 PCIe round trips per image)
  gather_lifting_results / get_6d_rep
                     one fp64 pose kernel (template, Kabsch SVD, Euler, alpha)
  forward_crops      the whole path for already-cropped tensors with a single
                     D2H copy of the [N,7] pose records (what bench.py times)

  crop_instances     decoded uint8 images are uploaded once; one launch produces every
                     normalised crop (cv2.warpAffine's fixed-point bilinear + ToTensor +
                     Normalize, bit-identical).  Image decoding (cv2.imread) stays host code.
"""
import math
from os.path import join as pjoin

import numpy as np
import torch
import torch.nn as nn

from . import FCmodel
from . import heatmapModel
from .. import common  # noqa: F401
from ..common import img_proc as lip
from ..common import transformation as ltr
from ..common.format import get_pred_str, save_txt_file
from ... import _native as N

# 1-based edge list of the cuboid (reference: interp_dict['bbox12'], car_instance.py:63-70)
_EDGE_PARENTS = np.array([1, 3, 5, 7, 1, 2, 3, 4, 1, 2, 5, 6])
_EDGE_CHILDREN = np.array([2, 4, 6, 8, 5, 6, 7, 8, 3, 4, 7, 8])


class EgoNet(nn.Module):
    def __init__(self, cfgs, pre_trained=False):
        super().__init__()
        hm_cfg = cfgs['heatmapModel']
        backbone = getattr(heatmapModel, hm_cfg['name'], None)
        if backbone is None:
            raise NotImplementedError('heatmapModel %r has no native implementation' % hm_cfg['name'])
        self.HC = backbone.get_pose_net(cfgs, is_train=False)
        self.resolution = hm_cfg['input_size']
        self.xy_dict = {'flag': hm_cfg['add_xy']} if 'add_xy' in hm_cfg else None
        self.L = FCmodel.get_fc_model(stage_id=1, cfgs=cfgs, input_size=cfgs['FCModel']['input_size'],
                                      output_size=cfgs['FCModel']['output_size'])
        self.pth_trans = None
        self.LS = None
        if pre_trained:
            ckpt = cfgs['dirs']['ckpt']
            self.HC.load_state_dict(torch.load(pjoin(ckpt, 'HC.pth'), map_location='cpu'))
            self.LS = np.load(pjoin(ckpt, 'LS.npy'), allow_pickle=True).item()
            self.L.load_state_dict(torch.load(pjoin(ckpt, 'L.pth'), map_location='cpu'))

    # LS is a plain attribute upstream; forward it to the fused lifter whenever it is set
    def __setattr__(self, name, value):
        if name == 'LS':
            object.__setattr__(self, name, value)
            if 'L' in self._modules:
                self.L.set_stats(value)
            return
        super().__setattr__(name, value)

    def _device(self):
        return torch.device('cuda', torch.cuda.current_device())

    # ------------------------------------------------------------------ crops (device front-end)
    def crop_single_instance(self, img, bbox, resolution, pth_trans=None, xy_dict=None):
        """[egonet.py:68-95] one box of one decoded image through the device crop kernel.
        Returns what upstream returns: the uint8 [h,w,3] warp result (numpy) without ``pth_trans``,
        else the normalised fp32 [3,h,w] tensor (left on the device)."""
        if xy_dict is not None and xy_dict['flag']:
            raise NotImplementedError('add_xy input channels need generate_xy_map (out of the hot path)')
        bbox = lip.to_npy(bbox)
        width, height = resolution
        ret = lip.modify_bbox(bbox, height / width)
        image = img if torch.is_tensor(img) else torch.from_numpy(np.ascontiguousarray(img))
        image = image.to(self._device())
        mean, std = (None, None) if pth_trans is None else lip.normalize_params(pth_trans)
        out, u8 = lip.crop_instances_device([image], [0], ret['c'][None], ret['s'][None], resolution,
                                            mean, std, return_u8=True)
        return u8[0].cpu().numpy() if pth_trans is None else out[0]

    def load_cv2(self, path, rgb=True):
        """[egonet.py:97-103] image decoding stays host code (cv2.imread), as upstream."""
        import cv2
        data = cv2.imread(path, 1 | 128)
        if data is None:
            raise ValueError('Fail to read {}'.format(path))
        return cv2.cvtColor(data, cv2.COLOR_BGR2RGB) if rgb else data

    def crop_instances(self, annot_dict, resolution, pth_trans=None, rgb=True, xy_dict=None):
        """[egonet.py:105-155] every box of every image of the batch: each decoded image is uploaded
        once (uint8) and ONE launch writes all normalised crops, instead of a cv2.warpAffine +
        ToTensor + Normalize per box on the host.  ``annot_dict['images']`` (decoded uint8 RGB arrays
        or CUDA tensors, one per path) skips the file read.  Returns (CUDA fp32 [N,3,h,w], records)."""
        if xy_dict is not None and xy_dict['flag']:
            raise NotImplementedError('add_xy input channels need generate_xy_map (out of the hot path)')
        mean, std = lip.normalize_params(pth_trans)
        dev = self._device()
        images, image_of_crop, centers, scales, records = [], [], [], [], []
        target_ar = resolution[1] / resolution[0]
        for img_idx, path in enumerate(annot_dict['path']):
            boxes = annot_dict['boxes'][img_idx]
            n = len(boxes)
            if n == 0:
                continue
            image = annot_dict['images'][img_idx] if 'images' in annot_dict else self.load_cv2(path, rgb)
            if not torch.is_tensor(image):
                image = torch.from_numpy(np.ascontiguousarray(image))
            images.append(image.to(dev, non_blocking=True))
            labels = annot_dict['labels'][img_idx] if 'labels' in annot_dict else -np.ones(n, dtype=np.int64)
            scores = annot_dict['scores'][img_idx] if 'scores' in annot_dict else -np.ones(n)
            # all boxes of the image at once (row-wise identical to modify_bbox per box)
            boxes_np = boxes if isinstance(boxes, np.ndarray) else np.stack([lip.to_npy(b) for b in boxes])
            resized, cs, ss = lip.modify_bbox_batch(boxes_np, target_ar)
            image_of_crop.extend([len(images) - 1] * n)
            centers.append(cs)
            scales.append(ss)
            for k in range(n):
                records.append({'path': path, 'center': cs[k], 'scale': ss[k], 'bbox': boxes_np[k],
                                'bbox_resize': list(resized[k]), 'rotation': 0., 'label': labels[k],
                                'score': scores[k]})
        if not records:
            return torch.empty((0, 3, int(resolution[1]), int(resolution[0])), device=dev), records
        crops = lip.crop_instances_device(images, image_of_crop, np.concatenate(centers), np.concatenate(scales),
                                          resolution, mean, std)
        return crops, records

    # ------------------------------------------------------------------ HC + affine
    def new_img_dict(self):
        """[egonet.py:410-422]"""
        return {'center': [], 'scale': [], 'rotation': [], 'bbox_resize': [], 'kpts_2d_pred': [],
                'label': [], 'score': []}

    def keypoints_device(self, instances, centers, scales, rots=None):
        """HC forward + inverse crop affine, device in / device out.
        instances [N,C,H,W] CUDA fp32 -> (screen key-points fp64 [N,J,2], coords fp32 [N,J,2])."""
        out = self.HC(instances)
        coords = out[1]
        return lip.local_to_screen(coords, centers, scales, self.resolution, rots), coords

    def get_keypoints(self, instances, records, is_cuda=True):
        """[egonet.py:424-467]"""
        if not is_cuda:
            raise RuntimeError('egonet_b200 has no CPU path (is_cuda=False is not supported)')
        dev = self._device()
        instances = instances.to(dev, non_blocking=True)
        n = len(records)
        centers = np.array([r['center'] for r in records], dtype=np.float64).reshape(n, 2)
        scales = np.array([r['scale'] for r in records], dtype=np.float64).reshape(n, 2)
        rots = np.array([r['rotation'] for r in records], dtype=np.float64).reshape(n)
        screen, _ = self.keypoints_device(instances, centers, scales, rots)
        screen = screen.cpu().numpy()                      # the single D2H copy of this stage
        ret = {}
        for i, record in enumerate(records):
            record['kpts'] = screen[i]
            img = ret.setdefault(record['path'], self.new_img_dict())
            img['kpts_2d_pred'].append(screen[i].reshape(1, -1))
            for key in ('center', 'scale', 'bbox_resize', 'label', 'score', 'rotation'):
                img[key].append(record[key])
        return ret

    # ------------------------------------------------------------------ lifter
    def lift_2d_to_3d(self, records, cuda=True):
        """[egonet.py:469-486] -- one device call for all images of the batch."""
        if not cuda:
            raise RuntimeError('egonet_b200 has no CPU path (cuda=False is not supported)')
        if self.LS is None:
            raise ValueError('EgoNet.LS (lifter statistics) is not set')
        paths = list(records.keys())
        rows = [np.concatenate(records[p]['kpts_2d_pred'], axis=0) for p in paths]
        counts = [len(r) for r in rows]
        data = torch.from_numpy(np.concatenate(rows, axis=0)).to(self._device())
        pred = self.L.lift(data).cpu().numpy()
        start = 0
        for p, c in zip(paths, counts):
            records[p]['kpts_3d_pred'] = pred[start:start + c].reshape(c, -1, 3)
            start += c
        return records

    # ------------------------------------------------------------------ pose
    def get_template(self, prediction, interp_coef=[0.332, 0.667]):
        """[egonet.py:238-263] host helper (the kernel rebuilds the template itself)."""
        prediction = np.asarray(prediction)
        seg = prediction[_EDGE_PARENTS - 1] - prediction[_EDGE_CHILDREN - 1]
        seg = np.sqrt(np.sum(seg ** 2, axis=1))
        h, l, w = np.sum(seg[:4]) / 4, np.sum(seg[4:8]) / 4, np.sum(seg[8:]) / 4
        box = np.array([np.array([l, l, l, l, 0, 0, 0, 0]) - np.float32(l) / 2,
                        np.array([0, h, 0, h, 0, h, 0, h]) - np.float32(h),
                        np.array([w, w, 0, 0, w, w, 0, 0]) - np.float32(w) / 2])
        if len(prediction) == 32:
            a, b = box[:, _EDGE_PARENTS - 1], box[:, _EDGE_CHILDREN - 1]
            box = np.hstack([box] + [a + c * (b - a) for c in interp_coef])
        return box

    def _solve(self, predictions, kpts_2d=None, K=None, alpha_mode='trans', want_rotation=False):
        pred = torch.from_numpy(np.ascontiguousarray(predictions, dtype=np.float64)).to(self._device())
        k2 = None
        if kpts_2d is not None:
            k2 = torch.from_numpy(np.ascontiguousarray(kpts_2d, dtype=np.float64)).to(self._device())
        return ltr.pose_solve(pred.view(len(pred), -1, 3), k2, K, alpha_mode, want_rotation)

    def kpts_to_euler(self, template, prediction):
        """[egonet.py:265-277] single instance, template and prediction [3,P] (any template, as upstream):
        Kabsch on the device (``egn_rigid_transform``), Euler angles of the extrinsic 'yxz' sequence
        reordered to [x, y, z]; returns (angles [3], T [3,1])."""
        R, t = ltr.compute_rigid_transform(np.asarray(template, dtype=np.float64),
                                           np.asarray(prediction, dtype=np.float64))
        angles = np.array([math.asin(max(-1.0, min(1.0, R[2, 1]))), math.atan2(-R[2, 0], R[2, 2]),
                           math.atan2(-R[0, 1], R[1, 1])])
        return angles, t

    def get_6d_rep(self, predictions, ax=None, color="black"):
        """[egonet.py:279-295] -> (angles [N,3], translation [N,3]) fp64 numpy."""
        predictions = np.asarray(predictions).reshape(len(predictions), -1, 3)
        pose = self._solve(predictions).cpu().numpy()
        return pose[:, :3], predictions[:, 0, :]

    def _angles(self, ry, x, z, x_offset):
        dev = self._device()
        ry = torch.from_numpy(np.ascontiguousarray(ry, dtype=np.float64)).to(dev)
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        z = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float64)).to(dev)
        out = torch.empty_like(ry)
        with torch.cuda.device(dev):
            N.check(N.lib().egn_observation_angle(N.ptr(ry), N.ptr(x), 1, N.ptr(z), 0 if z.numel() == 1 else 1,
                                                  float(x_offset), ry.numel(), N.ptr(out), N.current_stream()))
        return out.cpu().numpy()

    def get_observation_angle_trans(self, euler_angles, translations):
        """[egonet.py:203-217]"""
        translations = np.asarray(translations)
        return self._angles(euler_angles[:, 1], translations[:, 0], translations[:, 2], 0.0)

    def get_observation_angle_proj(self, euler_angles, kpts, K):
        """[egonet.py:219-236]"""
        kx = np.array([np.asarray(kpts[i])[0, 0] for i in range(len(kpts))], dtype=np.float64)
        return self._angles(euler_angles[:, 1], kx, np.array([K[0, 0]], dtype=np.float64), K[0, 2])

    def gather_lifting_results(self, record, data, prediction, target=None, pose_vecs_gt=None,
                               intrinsics=None, refine=False, visualize=False, template=None,
                               dist_coeffs=np.zeros((4, 1)), color='r', get_str=False, alpha_mode='trans'):
        """[egonet.py:297-339] angles, translation and alpha from ONE pose-kernel launch."""
        if visualize:
            raise NotImplementedError('plotting is outside the native hot path')
        if alpha_mode not in ('trans', 'proj'):
            raise NotImplementedError
        kpts_3d = np.asarray(record['kpts_3d_pred'])
        k2 = np.concatenate(record['kpts_2d_pred'], axis=0) if alpha_mode == 'proj' else None
        pose = self._solve(kpts_3d, k2, record.get('K'), alpha_mode).cpu().numpy()
        record['euler_angles'] = pose[:, :3]
        record['translation'] = kpts_3d.reshape(len(kpts_3d), -1, 3)[:, 0, :]
        record['alphas'] = pose[:, 6]
        if get_str:
            record['pred_str'] = get_pred_str(record)
        return record

    def plot_one_image(self, img_path, record, visualize=False, color_dict=None,
                       save_dict={'flag': False, 'save_dir': None}, alpha_mode='trans'):
        """[egonet.py:341-383] (plotting itself is not provided)"""
        record = self.gather_lifting_results(record, None, None, visualize=visualize,
                                             get_str=save_dict['flag'], alpha_mode=alpha_mode)
        save_txt_file(img_path, record, save_dict)
        return record

    def post_process(self, records, visualize=False, color_dict=None,
                     save_dict={'flag': False, 'save_dir': None}, alpha_mode='trans'):
        """[egonet.py:385-408]"""
        for img_path in records.keys():
            print("Processing {:s}".format(img_path))
            records[img_path] = self.plot_one_image(img_path, records[img_path], visualize=visualize,
                                                    save_dict=save_dict, alpha_mode=alpha_mode)
        return records

    def add_orientation_arrow(self, record):
        """[egonet.py:157-179] screen-space arrow per instance (a visualisation aid carried in the record):
        the predicted heading (point 1 - point 5) drawn from the ground-truth centre, projected by K and
        clipped to 60 px when longer than 50 px.  Vectorised over the instances of the image."""
        pred = np.asarray(record['kpts_3d_pred'])
        gt = np.asarray(record['kpts_3d_gt'])
        K = np.asarray(record['K'])
        n = len(pred)
        heading = pred[:, 1] - pred[:, 5]
        ends = np.stack([gt[:n, 0], gt[:n, 0] + heading], axis=2)          # [n,3,2]
        proj = np.einsum('ab,nbk->nak', K, ends)
        arrow = proj[:, :2, :] / proj[:, 2:3, :]                           # [n,2(xy),2(start,end)]
        vec = arrow[:, :, 1] - arrow[:, :, 0]
        length = np.linalg.norm(vec, axis=1, keepdims=True)
        long = length[:, 0] > 50
        vec[long] = vec[long] / length[long] * 60
        arrow[:, :, 1] = arrow[:, :, 0] + vec
        return arrow

    def write_annot_dict(self, annot_dict, records):
        """[egonet.py:181-201] pass-through of the caller's annotations."""
        for idx, path in enumerate(annot_dict['path']):
            rec = records[path]
            for src, dst in (('boxes', 'boxes'), ('kpts', 'kpts_2d_gt'), ('kpts_3d_gt', 'kpts_3d_gt'),
                             ('pose_vecs_gt', 'pose_vecs_gt'), ('kpts_3d_before', 'kpts_3d_before')):
                if src in annot_dict:
                    rec[dst] = lip.to_npy(annot_dict[src][idx])
            for key in ('raw_txt_format', 'K'):
                if key in annot_dict:
                    rec[key] = annot_dict[key][idx]
            if 'kpts_3d_gt' in annot_dict and 'K' in annot_dict:
                rec['arrow'] = self.add_orientation_arrow(rec)
        return records

    def forward(self, annot_dict):
        """[egonet.py:488-507]"""
        instances, inst_records = self.crop_instances(annot_dict, resolution=self.resolution,
                                                      pth_trans=self.pth_trans, xy_dict=self.xy_dict)
        records = self.get_keypoints(instances, inst_records)
        records = self.lift_2d_to_3d(records)
        return self.write_annot_dict(annot_dict, records)

    # ------------------------------------------------------------------ fused fast path
    @torch.no_grad()
    def forward_crops_graphed(self, instances, centers, scales, K=None, alpha_mode='trans'):
        """``forward_crops`` replayed from a CUDA graph: the ~340 launches of the path (HC, decode-free affine,
        lifter, pose) are captured once per (batch size, K, alpha_mode) into static buffers and replayed with ONE
        graph launch per call, which removes the per-launch host cost that dominates small batches.
        Inputs are copied into the graph's static buffers (device-to-device); the returned [N,7] tensor is the
        graph's static output -- valid until the next call with the same batch size."""
        instances = instances.contiguous()
        n = instances.shape[0]
        dev = instances.device
        ce = torch.as_tensor(np.asarray(centers, dtype=np.float64) if not torch.is_tensor(centers) else centers,
                             dtype=torch.float64).to(dev).contiguous()
        sc = torch.as_tensor(np.asarray(scales, dtype=np.float64) if not torch.is_tensor(scales) else scales,
                             dtype=torch.float64).to(dev).contiguous()
        key = (n, tuple(instances.shape[1:]), dev.index, alpha_mode, None if K is None else tuple(np.asarray(K).ravel()))
        cache = self.__dict__.setdefault('_graphs', {})
        entry = cache.get(key)
        if entry is None:
            sx, sce, ssc = instances.clone(), ce.clone(), sc.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                 # warm-up outside capture: weight upload, workspaces, tensor maps
                for _ in range(2):
                    self.forward_crops(sx, sce, ssc, K=K, alpha_mode=alpha_mode)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward_crops(sx, sce, ssc, K=K, alpha_mode=alpha_mode)
            entry = cache[key] = (graph, sx, sce, ssc, out)
        graph, sx, sce, ssc, out = entry
        sx.copy_(instances, non_blocking=True)
        sce.copy_(ce, non_blocking=True)
        ssc.copy_(sc, non_blocking=True)
        graph.replay()
        return out

    @torch.no_grad()
    def forward_crops(self, instances, centers, scales, K=None, alpha_mode='trans', return_all=False):
        """Whole per-crop path on the device for already-cropped tensors.

        instances: CUDA fp32 [N,C,H,W]; centers/scales: [N,2] fp64 (CUDA tensors or
        arrays).  Returns the CUDA fp64 [N,7] pose records (Euler x,y,z | translation |
        alpha); with ``return_all`` also screen key-points, 3D key-points, coords."""
        if self.LS is None:
            raise ValueError('EgoNet.LS (lifter statistics) is not set')
        screen, coords = self.keypoints_device(instances, centers, scales)
        kpts_2d = screen.view(screen.shape[0], -1)
        kpts_3d = self.L.lift(kpts_2d)
        pose = ltr.pose_solve(kpts_3d.view(len(kpts_3d), -1, 3), kpts_2d, K, alpha_mode)
        if return_all:
            return {'pose': pose, 'kpts_2d': kpts_2d, 'kpts_3d': kpts_3d, 'coords': coords}
        return pose
