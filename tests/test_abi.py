"""CPU-side checks of the C-ABI boundary: the library loads without a GPU/driver,
exports every symbol the header declares, builds the parameter inventory, and
refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from egonet_b200 import _native as N
from oracle import configs, hrnet_ref, lifter_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'egonet_b200.h')).read()
    return sorted(set(re.findall(r'EGN_API\s+[\w\s\*]+?\b(egn_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    L = N.lib()
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), 'symbol %s declared in the header but not exported' % n
    assert sorted(N.SIGNATURES) == names, 'python binding and header disagree'
    assert L.egn_version() == 100


def test_library_has_no_driver_dependency():
    import subprocess
    out = subprocess.run(['ldd', N.LIB_PATH], capture_output=True, text=True).stdout
    assert 'libcuda.so' not in out and 'libcudart' not in out and 'libtorch' not in out


@pytest.mark.parametrize('name', ['demo', 'tiny', 'ped', 'demo_heatmap'])
def test_engine_inventory_matches_reference_state_dict(name):
    cfgs = {'demo': configs.demo_cfgs(), 'tiny': configs.tiny_cfgs(), 'ped': configs.ped_cfgs(),
            'demo_heatmap': configs.demo_cfgs('heatmap')}[name]
    from egonet_b200.libs.model.heatmapModel.hrnet import get_pose_net
    m = get_pose_net(cfgs, is_train=False)
    spec = hrnet_ref.state_dict_spec(cfgs)          # pinned against the reference module (golden tests)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)
    assert m.stats()['macs_per_crop'] == hrnet_ref.macs_per_crop(cfgs)
    m.load_state_dict(hrnet_ref.make_weights(cfgs, 1), strict=True)


def test_lifter_inventory_matches_reference_state_dict():
    cfgs = configs.demo_cfgs()
    from egonet_b200.libs.model.FCmodel import get_fc_model
    m = get_fc_model(1, cfgs, 66, 96)
    assert list(m.state_dict().keys()) == list(lifter_ref.state_dict_spec(cfgs).keys())
    m.load_state_dict(lifter_ref.make_weights(cfgs), strict=True)


def test_bad_configs_are_rejected():
    L = N.lib()
    from egonet_b200.libs.model.heatmapModel.hrnet import _engine_cfg
    cfgs = configs.demo_cfgs()
    c = _engine_cfg(cfgs, 3, N.PREC_FP16, N.CONV_AUTO, False)
    c.input_w = 250
    h = ctypes.c_void_p()
    assert L.egn_hrnet_create(ctypes.byref(c), ctypes.byref(h)) == -1
    assert b'multiple of 32' in L.egn_last_error()
    bad = configs.demo_cfgs()
    bad['heatmapModel']['head_type'] = 'angleregression'
    from egonet_b200.libs.model.heatmapModel.hrnet import get_pose_net
    with pytest.raises(NotImplementedError):
        get_pose_net(bad, is_train=False)


def test_set_weight_validates_keys_and_shapes():
    L = N.lib()
    from egonet_b200.libs.model.heatmapModel.hrnet import _engine_cfg
    c = _engine_cfg(configs.tiny_cfgs(), 3, N.PREC_FP32, N.CONV_AUTO, False)
    h = ctypes.c_void_p()
    N.check(L.egn_hrnet_create(ctypes.byref(c), ctypes.byref(h)))
    w = np.zeros((64, 3, 3, 3), dtype=np.float32)
    shp = (ctypes.c_int64 * 4)(64, 3, 3, 3)
    assert L.egn_hrnet_set_weight(h, b'conv1.weight', w.ctypes.data_as(ctypes.c_void_p), shp, 4) == 0
    assert L.egn_hrnet_set_weight(h, b'nope.weight', w.ctypes.data_as(ctypes.c_void_p), shp, 4) == -1
    bad = (ctypes.c_int64 * 4)(64, 4, 3, 3)
    assert L.egn_hrnet_set_weight(h, b'conv1.weight', w.ctypes.data_as(ctypes.c_void_p), bad, 4) == -1
    assert L.egn_hrnet_set_weight(h, b'bn1.num_batches_tracked', None, None, 0) == 0
    # forward before finalize is a state error, never a silent fallback
    assert L.egn_hrnet_forward(h, ctypes.c_void_p(8), 1, None, None, None, None, 0, None) == -4
    L.egn_hrnet_destroy(h)


@pytest.mark.skipif(HAS_GPU, reason='checks the no-GPU behaviour')
def test_no_cpu_fallback_without_device():
    L = N.lib()
    assert L.egn_device_ok() == 0
    rc = L.egn_argmax2d(ctypes.c_void_p(8), 1, 1, 4, 4, None, ctypes.c_void_p(8), ctypes.c_void_p(8), None)
    assert rc == -2 and b'no CPU fallback' in L.egn_last_error()
    from egonet_b200.libs.common import img_proc
    with pytest.raises(RuntimeError):
        img_proc.get_max_preds(torch.zeros(1, 1, 4, 4))
    from egonet_b200.libs.model.egonet import EgoNet
    m = EgoNet(configs.tiny_cfgs()).eval()
    with pytest.raises(RuntimeError):
        m.HC(torch.zeros(1, 3, 128, 128))
    with pytest.raises(RuntimeError):
        m.get_keypoints(torch.zeros(1, 3, 128, 128), [], is_cuda=False)
