#!/usr/bin/env python
"""Benchmark of the EgoNet per-crop inference hot path (BASELINE.json metric: crops/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

One "step" = one pass of the whole per-crop path over one batch of synthetic 256x256 crops:
HC (HRNet-W48 heat-maps + coordinate head, tcgen05 convs) -> heat-map decode (hard arg-max and
soft-argmax of the 33 maps) -> inverse crop affine -> lifter -> pose solve.  The headline mode is
`fp16x2`: error-compensated split-fp16 tensor-core convs, the mode whose results meet the 1e-4 parity
bound against the reference (tests/test_gpu_parity.py); the plain fp16 mode (faster, ~1e-3 on
coordinates) is timed beside it in `config.fast_mode_fp16` with its measured error.  Workload = BASELINE.json
configs[2] ("full inference: heatmap + lifter + pose, batch=256, 1xB200" -- the configuration whose
stages are exactly the metric's "heatmap+lift+pose"), per GPU; the batch-64 figure of configs[1] is
reported alongside in `config.batch64`.

N > 1 is launched by torchrun (one process per GPU): crops are sharded with no data-path collective
(weak scaling: every rank processes its own batch) and the [B,7] pose records are all-gathered
over NCCL at the end of every step, as BASELINE.json configs[4] asks.

Prints ONE JSON line (rank 0).  `value` = whole-job crops/s with the inputs resident in HBM;
`e2e` = the same path through the public API (EgoNet.forward_crops) from pinned HOST memory
including the H2D copy of the crops and the D2H copy of the pose records; `roofline` = the
dominant kernel class timed live with CUDA events; `cpu_baseline` = the CPU oracle (a port of the
reference's algorithm on torch-CPU, the arithmetic the reference itself would run on CPU) on a
bounded sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'crops/sec (heatmap+lift+pose) on 256x256 synthetic batches'
L2_BYTES = 126 * 1024 * 1024


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'bf16_tflops': p['bf16_tflops'],
                'bf16_tflops_sustained': p.get('bf16_tflops_sustained', p['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        reasons = []
        for i, name in enumerate(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')):
            if any(s[3 + i].lower().startswith('active') for s in self.samples):
                reasons.append(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.samples)}


# ----------------------------------------------------------------------------- workload
def build_model(cfgs, device):
    from egonet_b200 import synth
    from egonet_b200.libs.model.egonet import EgoNet
    ego = EgoNet(cfgs, pre_trained=False).eval()
    ego.HC.load_state_dict(synth.hc_weights(ego.HC.state_dict(), 1))
    ego.L.load_state_dict(synth.lifter_weights(ego.L.state_dict(), 11))
    ego.LS = synth.lifter_stats(cfgs, 12)
    return ego.to(device)


def launches_per_step(ego):
    st = ego.HC.stats()
    lifter = 2 * ego.L.num_blocks + 2
    return st['launches'] + 2 + 1 + lifter + 1   # HC + (argmax, soft-argmax) + affine + lifter + pose


def one_step(ego, x, centers, scales, K):
    """The device-resident step: everything after the crops are in HBM."""
    from egonet_b200.libs.common import img_proc, transformation
    maps, coords = ego.HC(x)
    img_proc.get_max_preds(maps)
    img_proc.soft_arg_max(maps)
    screen = img_proc.local_to_screen(coords, centers, scales, ego.resolution)
    k2 = screen.view(screen.shape[0], -1)
    k3 = ego.L.lift(k2)
    return transformation.pose_solve(k3.view(len(k3), -1, 3), k2, K, 'proj')


def profile_hc(ego, x, passes=3):
    """Per-op device times of the HC engine (CUDA events between launches), grouped by kernel class."""
    from egonet_b200 import _native as N
    L = N.lib()
    h = ego.HC._handle
    n = L.egn_hrnet_num_launches(h)
    B = x.shape[0]
    need = L.egn_hrnet_workspace_bytes(h, B)
    ws = torch.empty(need, device=x.device, dtype=torch.uint8)
    acc = np.zeros(n)
    buf = (ctypes.c_float * n)()
    for _ in range(passes):
        N.check(L.egn_hrnet_profile(h, N.ptr(x), B, N.ptr(ws), need, N.current_stream(), buf))
        acc += np.frombuffer(buf, dtype=np.float32)
    acc /= passes
    classes = {}
    info = N.OpInfo()
    kinds = {0: 'stem_conv', 1: 'conv', 2: 'fuse', 3: 'head_tail'}
    for i in range(n):
        N.check(L.egn_hrnet_op_info(h, i, ctypes.byref(info)))
        if info.kind == 1:
            name = '%s %dx%d s%d %d->%d @%dx%d%s' % ('conv_tc' if info.use_tc else 'conv_simt', info.ksize, info.ksize,
                                                     info.stride, info.Cin, info.Cout, info.OH, info.OW,
                                                     '+res' if info.has_res else '')
        else:
            name = kinds[info.kind]
        c = classes.setdefault(name, {'ms': 0.0, 'launches': 0, 'macs': 0, 'act_bytes': 0, 'weight_bytes': 0})
        c['ms'] += float(acc[i])
        c['launches'] += 1
        c['macs'] += info.macs * B
        c['act_bytes'] += info.act_bytes * B
        c['weight_bytes'] += info.weight_bytes
    return classes, float(acc.sum())


def roofline_block(classes, total_ms, pk, batch, precision='fp16'):
    name, c = max(classes.items(), key=lambda kv: kv[1]['ms'])
    dur = c['ms'] / c['launches'] * 1e-3                      # seconds per launch
    bytes_per_launch = (c['act_bytes'] + c['weight_bytes']) / c['launches']
    flops_per_launch = 2.0 * c['macs'] / c['launches']
    hbm = bytes_per_launch / dur / 1e9
    tf = flops_per_launch / dur / 1e12
    f_h, f_t = hbm / pk['hbm_gbs'], tf / pk['bf16_tflops']
    bound = 'hbm' if bytes_per_launch / (pk['hbm_gbs'] * 1e9) >= flops_per_launch / (pk['bf16_tflops'] * 1e12) else 'tensor'
    blk = {'kernel': name, 'bound': bound,
           'achieved': round(hbm if bound == 'hbm' else tf, 2), 'peak': pk['hbm_gbs'] if bound == 'hbm' else pk['bf16_tflops'],
           'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s', 'frac': round(f_h if bound == 'hbm' else f_t, 4),
           'peak_source': pk['source'] + (' (burst)' if bound == 'tensor' else ''),
           'traffic': None, 'other_roof_frac': round(f_t if bound == 'hbm' else f_h, 4),
           'share_of_hc_time': round(c['ms'] / total_ms, 4), 'launches': c['launches'],
           'avg_launch_us': round(dur * 1e6, 2), 'algorithmic_bytes_per_launch': int(bytes_per_launch),
           'flops_per_launch': int(flops_per_launch)}
    try:
        table = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        tr = table.get(name + ' | ' + precision) or (table.get(name) if precision == 'fp16' else None)
        if tr and tr['batch'] == batch:
            blk['traffic'] = tr['bytes']
            blk['traffic_source'] = 'ncu --set full capture of this kernel class at this batch size (profiles/)'
    except (OSError, ValueError):
        pass
    top = sorted(classes.items(), key=lambda kv: -kv[1]['ms'])
    blk['top_classes'] = [{'kernel': k, 'ms': round(v['ms'], 3), 'launches': v['launches']} for k, v in top[:6]]
    # every class: ms per batch, launches, algorithmic TFLOP/s and GB/s (activation + weight bytes)
    blk['all_classes'] = [[k, round(v['ms'], 3), v['launches'], round(2.0 * v['macs'] / max(v['ms'], 1e-9) / 1e9, 1),
                           round((v['act_bytes'] + v['weight_bytes']) / max(v['ms'], 1e-9) / 1e6, 1)] for k, v in top]
    return blk


def side_stages(ego, recs, B, K, dev, pk_hbm):
    """Device time of the crop front-end (B crops cut from one KITTI-sized uint8 image) and of the
    optional reprojection refinement (B instances of 33 points), each timed alone with CUDA events."""
    from egonet_b200.libs.common import img_proc, transformation
    g = torch.Generator().manual_seed(5)
    image = torch.randint(0, 256, (375, 1242, 3), generator=g, dtype=torch.uint8).to(dev)
    centers = np.array([r['center'] for r in recs])
    scales = np.array([r['scale'] for r in recs])
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    ms_crop = timed(lambda: img_proc.crop_instances_device([image], [0] * B, centers, scales, ego.resolution, mean, std))
    out_bytes = B * 3 * ego.resolution[0] * ego.resolution[1] * 4
    k3 = torch.randn((B, 33, 3), generator=g, dtype=torch.float64) * 0.8 + torch.tensor([0., 1., 25.], dtype=torch.float64)
    Kt = torch.as_tensor(np.asarray(K, dtype=np.float64))
    uv = k3 @ Kt.T
    uv = uv[..., :2] / uv[..., 2:3]
    k3d, uvd = (k3 + 0.02 * torch.randn(k3.shape, generator=g, dtype=torch.float64)).to(dev), uv.to(dev)
    ms_pnp = timed(lambda: transformation.pnp_refine_batch(k3d, uvd, K))
    return {'crop_frontend': {'ms_per_batch': round(ms_crop, 4), 'crops_per_s': round(B / ms_crop * 1e3, 1),
                              'write_gbs': round(out_bytes / ms_crop / 1e6, 1),
                              'hbm_frac': round(out_bytes / ms_crop / 1e6 / pk_hbm, 4),
                              'note': 'uint8 image in HBM -> fp32 NCHW crops incl. host-side table upload; '
                                      'algorithmic bytes = the fp32 crops written'},
            'pnp_refine': {'ms_per_batch': round(ms_pnp, 4), 'instances_per_s': round(B / ms_pnp * 1e3, 1),
                           'points': 33, 'note': 'optional stage (disabled in the reference\'s maintained path)'}}


def fast_mode_block(cfgs, dev, x_sets, centers, scales, K, exact_model, B):
    """The plain fp16 tensor-core mode on the same batches: crops/s and its measured error against the
    headline (fp16x2) mode's coordinates and poses on the same crops."""
    fast_cfgs = json.loads(json.dumps(cfgs))
    fast_cfgs['heatmapModel']['b200_precision'] = 'fp16'
    fast = build_model(fast_cfgs, dev)
    for i in range(3):
        one_step(fast, x_sets[i % len(x_sets)], centers, scales, K)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(10):
        one_step(fast, x_sets[i % len(x_sets)], centers, scales, K)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    of = fast.forward_crops(x_sets[0], centers, scales, K=K, alpha_mode='proj', return_all=True)
    oe = exact_model.forward_crops(x_sets[0], centers, scales, K=K, alpha_mode='proj', return_all=True)
    return {'crops_per_s': round(B / ms * 1e3, 1), 'ms_per_step': round(ms, 4),
            'coords_err_vs_headline_mode': float((of['coords'] - oe['coords']).abs().max()),
            'euler_err_vs_headline_mode_rad': float((of['pose'][:, :3] - oe['pose'][:, :3]).abs().max()),
            'note': 'b200_precision=fp16: single fp16 operands (the reference is fp32: this mode misses the 1e-4 '
                    'parity bound and is opt-in)'}


def e2e_images_block(ego, cfgs, dev, K, steps):
    """End to end through the reference's own entry, EgoNet.forward(annot_dict) + post_process, from decoded
    uint8 images in HOST memory: per step 16 KITTI-sized images x 16 boxes = 256 crops.  Inside the timed
    region: H2D of the images, device crop front-end, HC, affine, lifter, pose, D2H of the records."""
    from egonet_b200 import synth
    g = np.random.Generator(np.random.PCG64(77))
    n_img, per = 16, 16
    images = [g.integers(0, 256, (375, 1242, 3), dtype=np.uint8) for _ in range(n_img)]
    boxes = []
    for i in range(n_img):
        recs = synth.boxes(per, cfgs, 300 + i)
        boxes.append(np.array([r['bbox'] for r in recs]) if 'bbox' in recs[0] else None)
    if boxes[0] is None:                      # synth.boxes carries centre / scale only: rebuild corner boxes
        boxes = []
        for i in range(n_img):
            cx, cy = g.uniform(50, 1190, per), g.uniform(120, 330, per)
            w, h = g.uniform(30, 400, per), g.uniform(25, 250, per)
            boxes.append(np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1))
    annot = {'path': ['img_%02d.png' % i for i in range(n_img)], 'boxes': boxes, 'images': images,
             'K': [K] * n_img}
    if ego.pth_trans is None:                 # what tools/inference.py takes from the dataset (car_instance.py:522-531)
        import torchvision.transforms as tvt
        ego.pth_trans = tvt.Compose([tvt.ToTensor(), tvt.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    import contextlib
    import io

    def run():
        with contextlib.redirect_stdout(io.StringIO()):        # post_process prints one line per image, as upstream
            rec = ego(annot)
            return ego.post_process(rec, alpha_mode='proj')
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {'crops_per_s': round(n_img * per / dt, 1), 'ms_per_step': round(dt * 1e3, 3), 'images_per_step': n_img,
            'crops_per_step': n_img * per, 'h2d_bytes_per_step': int(sum(im.nbytes for im in images)),
            'note': 'EgoNet.forward(annot_dict) + post_process: uint8 images from host memory, device crop '
                    'front-end, record dicts back on the host (wall clock around synchronised steps)'}


def train_block(dev, batch, steps, pk):
    """BASELINE configs[3]: heat-map network forward + MSE loss + backward + Adam step at batch 128 on one GPU,
    through the reference-shaped API (model(data) -> loss_func -> loss.backward() -> optim.step(), trainer.py:183-198)
    on the native training engine.  fp32 arithmetic (the reference trains in fp32; a bf16 tensor-core path is not
    built), synthetic crops and Gaussian-like targets resident in HBM."""
    from egonet_b200 import synth
    from egonet_b200.libs.loss.function import JointsMSELoss
    from egonet_b200.libs.model.heatmapModel.hrnet import get_pose_net
    from egonet_b200.libs.optimizer.optimizer import prepare_optim
    from egonet_b200.libs.trainer.trainer import train_step
    cfgs = synth.demo_cfgs('heatmap')
    hm = cfgs['heatmapModel']
    model = get_pose_net(cfgs, is_train=False)
    model.load_state_dict(synth.hc_weights(model.state_dict(), 1))
    model = model.to(dev).train()
    cfgs['optimizer'] = dict(optim_type='adam', lr=1e-4, weight_decay=0.0, momentum=0.0, milestones=[1000], gamma=0.1)
    optim, _ = prepare_optim(model, cfgs)
    crit = JointsMSELoss(True)
    xs = [synth.crops(batch, cfgs, 50 + i).to(dev) for i in range(2)]
    g = torch.Generator().manual_seed(9)
    tgt = torch.rand((batch, hm['num_joints'], hm['heatmap_size'][1], hm['heatmap_size'][0]), generator=g).to(dev)
    w = torch.ones((batch, hm['num_joints'], 1), device=dev)
    losses = []
    for i in range(2):
        losses.append(float(train_step(model, crit, optim, xs[i % 2], tgt, w)))
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = train_step(model, crit, optim, xs[i % 2], tgt, w)
    b.record()
    torch.cuda.synchronize()
    losses.append(float(loss))
    ms = a.elapsed_time(b) / steps
    flops = model.train_flops_per_sample()
    rate = batch / ms * 1e3
    ws_gb = model._train['workspace'].numel() / 1e9
    del model, optim
    torch.cuda.empty_cache()
    return {'metric': 'samples/sec (heat-map forward + MSE loss + backward + Adam step)', 'value': round(rate, 2),
            'unit': 'samples/s', 'batch': batch, 'steps': steps, 'ms_per_step': round(ms, 2), 'dtype': 'f32',
            'flops_per_sample': int(flops), 'achieved_tflops': round(rate * flops / 1e12, 2),
            'frac_of_bf16_tensor_peak': round(rate * flops / 1e12 / pk['bf16_tflops_sustained'], 4),
            'workspace_gb': round(ws_gb, 1), 'loss_first_last': [round(losses[0], 6), round(losses[-1], 6)],
            'note': 'configs[3] asks for bf16: this engine is fp32 on CUDA cores (reference precision, parity-tested '
                    'against the reference module and an fp64 oracle); convs are 64x64-tile FFMA kernels'}


def stream4096_block(ego, dev_sets, centers, scales, K, world, rank, dist, B):
    """BASELINE configs[4]: a 4096-crop stream block-partitioned over the ranks (4096 / world each), processed
    in micro-batches of B, ONE all_gather of the [4096/world, 7] pose records at the end."""
    per_rank = 4096 // world
    n_micro = max(1, per_rank // B)
    poses = torch.empty((n_micro * B, 7), device=centers.device, dtype=torch.float64)
    gathered = torch.empty((world * n_micro * B, 7), device=centers.device, dtype=torch.float64) if dist else None

    def run():
        for i in range(n_micro):
            poses[i * B:(i + 1) * B] = one_step(ego, dev_sets[i % len(dev_sets)], centers, scales, K)
        if dist:
            dist.all_gather_into_tensor(gathered, poses)
    run()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=centers.device, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    return {'crops': world * n_micro * B, 'ms': round(ms, 3), 'crops_per_s': round(world * n_micro * B / ms * 1e3, 1),
            'micro_batch': B, 'micro_batches_per_rank': n_micro, 'gathers': 1 if dist else 0}


# ----------------------------------------------------------------------------- CPU oracle legs
def cpu_pipeline_rate(cfgs, batch, iters, threads):
    """crops/s of the oracle port (reference algorithm on torch-CPU / numpy) for the full path."""
    from oracle import decode_ref, egonet_ref, hrnet_ref, lifter_ref
    torch.set_num_threads(threads)
    hc = hrnet_ref.make_weights(cfgs, 1)
    lw = lifter_ref.make_weights(cfgs, 11)
    stats = lifter_ref.make_stats(cfgs, 12)
    x = egonet_ref.synth_crops(batch, cfgs, 0)
    recs = egonet_ref.synth_boxes(batch, cfgs, 2)

    def step():
        out = egonet_ref.run_pipeline(hc, lw, stats, cfgs, x, recs)
        decode_ref.get_max_preds(out['maps'])
        decode_ref.soft_arg_max(out['maps'])

    step()  # warm-up
    times = []
    for _ in range(iters):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return batch / float(np.median(times)), times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    Python/PyTorch, its algorithm restated on torch-CPU + numpy), all host threads, bounded sample."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from egonet_b200 import synth
    cfgs = synth.demo_cfgs()
    ncpu = os.cpu_count() or 1
    threads, best = ncpu, 0.0
    for t in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):      # torch-CPU convs do not scale to all cores
        r, _ = cpu_pipeline_rate(cfgs, 4, 1, t)
        if r > best:
            threads, best = t, r
    batch = args.ref_batch
    from oracle import decode_ref, egonet_ref, hrnet_ref, lifter_ref
    torch.set_num_threads(threads)
    hc = hrnet_ref.make_weights(cfgs, 1)
    lw = lifter_ref.make_weights(cfgs, 11)
    stats = lifter_ref.make_stats(cfgs, 12)
    x = egonet_ref.synth_crops(batch, cfgs, 0)
    recs = egonet_ref.synth_boxes(batch, cfgs, 2)

    def step():
        out = egonet_ref.run_pipeline(hc, lw, stats, cfgs, x, recs)
        decode_ref.get_max_preds(out['maps'])
        decode_ref.soft_arg_max(out['maps'])

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': round(value, 3), 'unit': 'crops/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(dt / args.steps * 1e3, 2),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[2]: full per-crop inference (HRNet-W48 heatmap+coords, decode, '
                                   'affine, lifter, pose), 256x256 crops', 'batch_per_step': batch,
                       'note': 'bounded sample of the GPU arm\'s workload (same model, same stages)'},
            'cpu_baseline': {'value': round(value, 3), 'unit': 'crops/s', 'cores': threads, 'kind': 'port',
                             'sample': '%d steps x %d crops, torch-CPU fp32 + numpy, best of {8,16,32,64,%d} threads' % (args.steps, batch, ncpu)},
            'e2e': {'value': round(value, 3), 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=256, help='crops per GPU per step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--ref-batch', type=int, default=16)
    ap.add_argument('--cpu-baseline-crops', type=int, default=16, help='crops per CPU pass (same batch as --ref-batch)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='inference', choices=['inference', 'train'],
                    help="'train': BASELINE configs[3] (batch 128 training step) as the headline line")
    ap.add_argument('--train-batch', type=int, default=128)
    ap.add_argument('--no-train', action='store_true', help='skip the configs[3] block of the default line')
    ap.add_argument('--precision', default='fp16x2', choices=['fp16', 'fp32', 'fp16x2'])
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a GPU (the product has no CPU path); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group('nccl', device_id=dev)

    if args.workload == 'train':
        # configs[3] is the 1-GPU case; N > 1 = data parallel (one process per GPU, 128 samples per rank, ONE
        # all-reduce of the flat gradient buffer per step: libs/trainer/trainer.py::allreduce_gradients)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        blk = train_block(dev, args.train_batch, args.steps, peaks())
        t = torch.tensor([blk['ms_per_step']], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            clocks = sampler.stop()
            ms = float(t[0])
            blk['ms_per_step'] = round(ms, 2)
            blk['value'] = round(world * args.train_batch / ms * 1e3, 2)
            line = {'metric': blk['metric'], 'value': blk['value'], 'unit': blk['unit'], 'n_gpus': world, 'steps': args.steps,
                    'warmup': 2, 'ms_per_step': blk['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                    'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                    'config': {'workload': 'configs[3]: train_IGRs path, heat-map forward + backward with the MSE loss + Adam, '
                                           'batch %d per GPU%s' % (args.train_batch, ', gradients all-reduced (1 collective / step)' if dist else ''),
                               **blk}, 'clocks': clocks}
            print(json.dumps(line))
        if dist:
            dist.destroy_process_group()
        return
    from egonet_b200 import synth
    cfgs = synth.demo_cfgs()
    cfgs['heatmapModel']['b200_precision'] = args.precision
    ego = build_model(cfgs, dev)
    B = args.batch
    K = synth.KITTI_K
    # rotating set of distinct input batches, total > L2, so no step finds its input cached
    n_sets = max(2, int(np.ceil(1.5 * L2_BYTES / (B * 3 * 256 * 256 * 4))))
    recs = synth.boxes(B, cfgs, 2 + rank)
    centers = torch.as_tensor(np.array([r['center'] for r in recs]), dtype=torch.float64, device=dev)
    scales = torch.as_tensor(np.array([r['scale'] for r in recs]), dtype=torch.float64, device=dev)
    host_sets = [synth.crops(B, cfgs, 100 * rank + i).pin_memory() for i in range(n_sets)]
    dev_sets = [h.to(dev) for h in host_sets]
    gathered = torch.empty((world * B, 7), device=dev, dtype=torch.float64) if dist else None

    def step(i):
        pose = one_step(ego, dev_sets[i % n_sets], centers, scales, K)
        if dist:
            dist.all_gather_into_tensor(gathered, pose)
        return pose

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(args.steps):
            step(i)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        # ---- e2e: public API from pinned host memory; every step's H2D copy of its crops and the D2H
        # copy of its pose records are inside the timed region.  The upload of step i+1 is issued on a
        # copy stream while step i computes (two device input slots), as a streaming caller would.
        out_host = torch.empty((B, 7), dtype=torch.float64).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        slots = [torch.empty_like(dev_sets[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            k = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[k])
                slots[k].copy_(host_sets[i % n_sets], non_blocking=True)
                ready[k].record(copy_stream)

        def e2e_run(n):
            cur = torch.cuda.current_stream()
            for k in range(2):
                freed[k].record(cur)
            upload(0)
            for i in range(n):
                k = i & 1
                if i + 1 < n:
                    upload(i + 1)
                cur.wait_event(ready[k])
                pose = ego.forward_crops(slots[k], centers, scales, K=K, alpha_mode='proj')
                freed[k].record(cur)
                if dist:
                    dist.all_gather_into_tensor(gathered, pose)
                out_host.copy_(pose, non_blocking=True)

        e2e_run(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        # ---- BASELINE configs[1] batch size (64) for reference, same path
        b64 = None
        if rank == 0 and B != 64:
            x64 = [d[:64].contiguous() for d in dev_sets]
            c64, s64 = centers[:64].contiguous(), scales[:64].contiguous()
            for i in range(3):
                one_step(ego, x64[i % n_sets], c64, s64, K)
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for i in range(10):
                one_step(ego, x64[i % n_sets], c64, s64, K)
            q1.record()
            torch.cuda.synchronize()
            b64 = 64 * 10 / (q0.elapsed_time(q1) * 1e-3)
            # the same batch through EgoNet.forward_crops: eager launches vs ONE CUDA-graph replay per step
            b64_fc = {}
            try:
                for name, fn in (('eager', ego.forward_crops), ('cuda_graph', ego.forward_crops_graphed)):
                    for i in range(3):
                        fn(x64[i % n_sets], c64, s64, K=K, alpha_mode='proj')
                    torch.cuda.synchronize()
                    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    q0.record()
                    for i in range(10):
                        fn(x64[i % n_sets], c64, s64, K=K, alpha_mode='proj')
                    q1.record()
                    torch.cuda.synchronize()
                    b64_fc[name] = round(64 * 10 / (q0.elapsed_time(q1) * 1e-3), 1)
            except Exception as e:
                b64_fc['error'] = '%s: %s' % (type(e).__name__, e)
        # ---- BASELINE configs[4]: the 4096-crop stream with a single gather (all ranks take part)
        stream = stream4096_block(ego, dev_sets, centers, scales, K, world, rank, dist, B)
        # ---- per-kernel-class timing for the roofline block (rank 0)
        classes, hc_ms = profile_hc(ego, dev_sets[0]) if rank == 0 else ({}, 0.0)
        fast_mode = e2e_img = None
        if rank == 0:
            try:
                if args.precision == 'fp16x2':
                    fast_mode = fast_mode_block(cfgs, dev, dev_sets, centers, scales, K, ego, B)
                e2e_img = e2e_images_block(ego, cfgs, dev, K, max(3, min(args.steps, 10)))
            except Exception as e:  # reported next to the headline numbers, never instead of them
                fast_mode = fast_mode or {'error': '%s: %s' % (type(e).__name__, e)}
                e2e_img = e2e_img or {'error': '%s: %s' % (type(e).__name__, e)}
        # ---- stages either side of the path, timed alone for reference
        extras = None
        if rank == 0:
            try:
                extras = side_stages(ego, recs, B, K, dev, pk_hbm=peaks()['hbm_gbs'])
            except Exception as e:  # reported next to the headline numbers, never instead of them
                extras = {'error': '%s: %s' % (type(e).__name__, e)}

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    pk = peaks()
    total_crops = world * B * args.steps
    value = total_crops / (ms * 1e-3)
    e2e = total_crops / (ms_e2e * 1e-3)
    st = ego.HC.stats()
    line = {
        'metric': METRIC, 'value': round(value, 1), 'unit': 'crops/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms / args.steps, 4), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': {'fp16': 'f16', 'fp32': 'f32', 'fp16x2': 'f16x2'}[args.precision],
        'data': 'synthetic',
        'config': {'workload': 'configs[2]: full per-crop inference (HRNet-W48 heatmap+coords, argmax+soft-argmax '
                               'decode, inverse affine, lifter, pose solve), 256x256 crops, batch %d per GPU' % B,
                   'batch64': {'crops_per_s': round(b64, 1), 'forward_crops': b64_fc,
                               'note': 'configs[1] batch size, same path, 1 GPU; forward_crops = HC + affine + lifter + pose '
                                       '(no heat-map decode) launched eagerly vs replayed from one CUDA graph'} if b64 else None,
                   'batch_per_gpu': B, 'global_batch': world * B, 'sharding': 'crops block-partitioned across ranks, '
                   'NCCL all_gather of [B,7] poses per step' if world > 1 else 'single GPU',
                   'l2': 'rotating %d distinct input batches (%.0f MB) > 126 MB L2; activation workspace %.0f MB' % (
                       n_sets, n_sets * B * 3 * 256 * 256 * 4 / 1e6,
                       ego.HC._workspace.numel() / 1e6 if ego.HC._workspace is not None else 0),
                   'hc_precision': args.precision, 'tc_conv_launches': st['tc_launches'], 'hc_launches': st['launches']},
        'e2e': {'value': round(e2e, 1), 'unit': 'crops/s', 'h2d_bytes_per_step': B * 3 * 256 * 256 * 4,
                'd2h_bytes_per_step': B * 7 * 8, 'ms_per_step': round(ms_e2e / args.steps, 4)},
        'gpu_launches': launches_per_step(ego) * args.steps,
        'clocks': clocks,
    }
    if extras:
        line['config']['side_stages'] = extras
    line['config']['stream4096'] = stream
    if world == 1 and not args.no_train:
        try:
            del ego, dev_sets, host_sets
            torch.cuda.empty_cache()
            line['config']['train_configs3'] = train_block(dev, args.train_batch, 3, pk)
        except Exception as e:
            line['config']['train_configs3'] = {'error': '%s: %s' % (type(e).__name__, e)}
    if fast_mode:
        line['config']['fast_mode_fp16'] = fast_mode
    if e2e_img:
        line['config']['e2e_images'] = e2e_img
    line['config']['parity'] = ('hc_precision=%s; ' % args.precision) + {
        'fp16x2': 'meets the 1e-4 bound vs the reference goldens (coords 4.8e-6, Euler 8.4e-6 rad on the demo config: '
                  'profiles/r02_parity_report.json, tests/test_gpu_parity.py)',
        'fp32': 'CUDA-core comparator, meets the 1e-4 bound',
        'fp16': 'fast mode: ~1e-3 on coordinates, misses the 1e-4 bound'}[args.precision]
    if classes:
        line['roofline'] = roofline_block(classes, hc_ms, pk, B, args.precision)
        line['hc_roofline'] = {
            'tensor_frac': round(value / world * 2 * st['macs_per_crop'] / (pk['bf16_tflops_sustained'] * 1e12), 4),
            'hbm_frac': round(value / world * (st['act_bytes_per_crop'] + st['weight_bytes'] / B) / (pk['hbm_gbs'] * 1e9), 4),
            'flops_per_crop': 2 * st['macs_per_crop'], 'algorithmic_bytes_per_crop': st['act_bytes_per_crop'],
            'weight_bytes': st['weight_bytes'], 'peaks': pk, 'hc_ms_per_batch_profiled': round(hc_ms, 3)}
    if world == 1 and not args.no_cpu_baseline:
        # reported baseline: the oracle port (reference algorithm on torch-CPU).  torch's CPU convs do not
        # scale to every core of the box, so a short calibration picks the best thread count first.
        ncpu = os.cpu_count() or 1
        best_t, best_r = ncpu, 0.0
        for t in sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)}):
            r, _ = cpu_pipeline_rate(cfgs, 4, 1, t)
            if r > best_r:
                best_t, best_r = t, r
        rate, times = cpu_pipeline_rate(cfgs, args.cpu_baseline_crops, 5, best_t)
        line['cpu_baseline'] = {'value': round(rate, 3), 'unit': 'crops/s', 'cores': best_t, 'kind': 'port',
                                'sample': '5 timed passes of %d crops (+1 warm-up; the reference arm uses the same batch) of the same workload, oracle port on '
                                          'torch-CPU fp32 with the best of {8,16,32,64,%d} threads on a %d-core host '
                                          '(median %.2f s/pass)' % (args.cpu_baseline_crops, ncpu, ncpu, float(np.median(times)))}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
