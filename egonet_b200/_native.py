"""ctypes binding of ``include/egonet_b200.h`` (the C-ABI drop-in boundary).

PyTorch is used for device memory and streams only; every computation goes
through ``libegonet_b200.so``.  There is deliberately no fallback: if the
library is missing, or a compute entry is called on a machine without an
sm_100 device, an exception is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libegonet_b200.so')

EGN_MAX_BRANCHES = 4
HEAD_HEATMAP, HEAD_COORDINATES = 0, 1
PREC_FP32, PREC_FP16, PREC_FP16X2 = 0, 1, 2
CONV_AUTO, CONV_SIMT = 0, 1
SOFTARGMAX_SOFTMAX, SOFTARGMAX_SUM = 0, 1
ALPHA_TRANS, ALPHA_PROJ = 0, 1


class EgnError(RuntimeError):
    """A C-ABI entry returned a negative status (message from egn_last_error)."""


class OpInfo(ctypes.Structure):
    _fields_ = [('kind', c_int), ('use_tc', c_int), ('Cin', c_int), ('Cout', c_int), ('H', c_int), ('W', c_int),
                ('OH', c_int), ('OW', c_int), ('ksize', c_int), ('stride', c_int), ('has_res', c_int),
                ('macs', c_int64), ('act_bytes', c_int64), ('weight_bytes', c_int64), ('name', ctypes.c_char * 96)]


class Image(ctypes.Structure):
    """egn_image: one decoded 8-bit RGB image resident in device memory."""
    _fields_ = [('data', c_void_p), ('height', c_int32), ('width', c_int32), ('pitch', c_int32), ('channels', c_int32)]


class HRNetCfg(ctypes.Structure):
    _fields_ = [
        ('in_channels', c_int), ('input_w', c_int), ('input_h', c_int),
        ('heatmap_w', c_int), ('heatmap_h', c_int), ('num_joints', c_int),
        ('head_type', c_int), ('final_conv_kernel', c_int), ('num_stages', c_int),
        ('stage_modules', c_int * 3), ('stage_branches', c_int * 3),
        ('stage_blocks', (c_int * EGN_MAX_BRANCHES) * 3),
        ('stage_channels', (c_int * EGN_MAX_BRANCHES) * 3),
        ('precision', c_int), ('conv_impl', c_int), ('keep_taps', c_int),
    ]


# name -> (restype, argtypes); must list every symbol declared in include/egonet_b200.h
SIGNATURES = {
    'egn_version': (c_int, []),
    'egn_last_error': (c_char_p, []),
    'egn_device_ok': (c_int, []),
    'egn_hrnet_create': (c_int, [POINTER(HRNetCfg), POINTER(c_void_p)]),
    'egn_hrnet_destroy': (None, [c_void_p]),
    'egn_hrnet_num_weights': (c_int, [c_void_p]),
    'egn_hrnet_weight_key': (c_char_p, [c_void_p, c_int]),
    'egn_hrnet_weight_shape': (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    'egn_hrnet_set_weight': (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    'egn_hrnet_finalize': (c_int, [c_void_p]),
    'egn_hrnet_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'egn_hrnet_forward': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    'egn_hrnet_read_tap': (c_int, [c_void_p, c_char_p, c_int, c_void_p, c_void_p, POINTER(c_int), c_void_p]),
    'egn_hrnet_macs_per_crop': (c_int64, [c_void_p]),
    'egn_hrnet_num_launches': (c_int, [c_void_p]),
    'egn_hrnet_num_tc_launches': (c_int, [c_void_p]),
    'egn_hrnet_act_bytes_per_crop': (c_int64, [c_void_p]),
    'egn_hrnet_weight_bytes': (c_int64, [c_void_p]),
    'egn_argmax2d': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_soft_argmax2d': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'egn_local_to_screen': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p]),
    'egn_lifter_create': (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    'egn_lifter_destroy': (None, [c_void_p]),
    'egn_lifter_set_weight': (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    'egn_lifter_set_stats': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_lifter_finalize': (c_int, [c_void_p]),
    'egn_lifter_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'egn_lifter_forward': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    'egn_pose_solve': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_double, c_double, c_int,
                               c_void_p, c_void_p, c_void_p]),
    'egn_hrnet_op_info': (c_int, [c_void_p, c_int, POINTER(OpInfo)]),
    'egn_hrnet_profile': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    'egn_conv2d_bench': (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p, c_int, POINTER(c_float)]),
    'egn_debug_umma_rate': (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_double)]),
    'egn_debug_umma_seq': (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_double)]),
    'egn_debug_conv_plan': (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_char_p, c_int]),
    'egn_debug_umma_probe': (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_conv2d_fused': (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    'egn_debug_conv_acc': (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    'egn_mse_hm_fwd_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_crop_instances': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                   POINTER(c_float), POINTER(c_float), c_void_p, c_void_p, c_void_p]),
    'egn_pnp_refine': (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_rigid_transform': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_similarity_transform': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_refine_with_bbox': (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_double, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'egn_hrnet_train_create': (c_int, [POINTER(HRNetCfg), POINTER(c_void_p)]),
    'egn_hrnet_train_destroy': (None, [c_void_p]),
    'egn_hrnet_train_flat_size': (c_int64, [c_void_p]),
    'egn_hrnet_train_param_offset': (c_int64, [c_void_p, c_int]),
    'egn_hrnet_train_param_trainable': (c_int, [c_void_p, c_int]),
    'egn_hrnet_train_workspace_bytes': (c_size_t, [c_void_p, c_int]),
    'egn_hrnet_train_flops_per_sample': (c_int64, [c_void_p]),
    'egn_hrnet_forward_train': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    'egn_hrnet_backward': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    'egn_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int] + [c_float] * 5 + [c_void_p]),
    'egn_sgd_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_float, c_int, c_void_p]),
    'egn_coord_loss_fwd_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int, c_float, c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    'egn_box_overlaps': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'egn_image_box_overlaps': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'egn_generate_target': (c_int, [c_void_p, c_void_p] + [c_int] * 6 + [c_double, c_void_p, c_void_p, c_void_p]),
    'egn_observation_angle': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the native library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EgnError('native library %s is missing: run `python -m egonet_b200.build` '
                           '(there is no Python/CPU fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise EgnError('egonet_b200 error %d: %s' % (rc, lib().egn_last_error().decode()))


def current_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device/host pointer of a contiguous tensor (or None)."""
    if t is None:
        return None
    assert t.is_contiguous()
    return c_void_p(t.data_ptr())
