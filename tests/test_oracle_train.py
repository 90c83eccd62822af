"""Training-step oracle (BASELINE configs[3]) against goldens made by the reference module
in train mode: Gaussian targets, loss, every parameter gradient (norm + sum, a few in full), BN running stats.
The CUDA train path is not built yet; this pins the target it will be held to."""
import numpy as np
import torch

from oracle import configs, egonet_ref, hrnet_ref, train_ref


def test_generate_target_matches_reference_golden(golden):
    g = golden('train_tiny.npz')
    hm = configs.tiny_cfgs('heatmap')['heatmapModel']
    for b in range(len(g['joints'])):
        t, w = train_ref.generate_target(g['joints'][b], g['vis'][b], hm['num_joints'], hm['input_size'],
                                         hm['heatmap_size'], int(g['sigma']))
        np.testing.assert_array_equal(t, g['target'][b])
        np.testing.assert_array_equal(w, g['target_weight'][b])
    assert (g['target_weight'].sum() < g['vis'].sum())            # some dots fall outside the map and are dropped
    assert g['target'].max() == 1.0


def test_train_step_matches_reference_golden(golden):
    g = golden('train_tiny.npz')
    cfgs = configs.tiny_cfgs('heatmap')
    torch.set_num_threads(4)
    sd = hrnet_ref.make_weights(cfgs, int(g['seed_w']))
    x = egonet_ref.synth_crops(len(g['joints']), cfgs, int(g['seed_x']))
    loss, grads, new_sd = train_ref.train_forward_backward(sd, cfgs, x, torch.from_numpy(g['target']),
                                                           torch.from_numpy(g['target_weight']))
    # same torch-CPU kernels as the reference module; thread-count dependent summation order only
    np.testing.assert_allclose(loss, float(g['loss']), rtol=1e-5)
    names = [str(n) for n in g['grad_names']]
    assert names == list(grads.keys())                            # every trainable parameter, reference order
    norms = np.array([grads[k].double().norm().item() for k in names])
    np.testing.assert_allclose(norms, g['grad_norms'], rtol=2e-3, atol=1e-9)
    for k in g:
        if k.startswith('grad__'):
            ref = g[k]
            np.testing.assert_allclose(grads[k[6:]].numpy(), ref, rtol=0, atol=2e-3 * np.abs(ref).max())
        elif k.startswith('stat__'):
            np.testing.assert_allclose(new_sd[k[6:]].numpy(), g[k], rtol=1e-4, atol=1e-6)
    # the input state dict is untouched, running statistics moved
    assert torch.equal(sd['bn1.running_mean'], hrnet_ref.make_weights(cfgs, int(g['seed_w']))['bn1.running_mean'])
    assert not torch.equal(new_sd['bn1.running_mean'], sd['bn1.running_mean'])


def test_target_kernel_source_on_host_vs_reference_golden(golden):
    """egonet_b200/csrc/target_math.h compiled for the host: visibility weights exact, dot geometry exact,
    Gaussian values within 2 ulp of numpy's float32 exp (rtol 3e-7)."""
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(root, 'tests', 'native', 'libtarget_host.so')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-I', os.path.join(root, 'egonet_b200', 'csrc'),
                           os.path.join(root, 'tests', 'native', 'target_host.cpp'), '-o', so])
    L = ctypes.CDLL(so)
    L.host_generate_target.argtypes = ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double] + [ctypes.c_void_p] * 2)
    g = golden('train_tiny.npz')
    hm = configs.tiny_cfgs('heatmap')['heatmapModel']
    joints, vis = np.ascontiguousarray(g['joints']), np.ascontiguousarray(g['vis'])
    n, k = vis.shape
    tgt = np.zeros((n, k, hm['heatmap_size'][0], hm['heatmap_size'][1]), np.float32)
    wgt = np.zeros((n, k), np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    L.host_generate_target(vp(joints), vp(vis), n, k, hm['input_size'][0], hm['input_size'][1], hm['heatmap_size'][0],
                           hm['heatmap_size'][1], float(g['sigma']), vp(tgt), vp(wgt))
    np.testing.assert_array_equal(wgt[..., None], g['target_weight'])
    np.testing.assert_array_equal(tgt == 0, g['target'] == 0)
    np.testing.assert_allclose(tgt, g['target'], rtol=3e-7, atol=0)
