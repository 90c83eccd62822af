"""Model-section config dictionaries used by tests, smoke and bench.

``demo_cfgs()`` restates the ``FCModel`` / ``heatmapModel`` blocks of the
reference's ``configs/KITTI_inference:demo.yml:61-151`` (HRNet-W48, 256x256
crops, 64x64 heat-maps, 33 joints, coordinate head).  ``tiny_cfgs()`` is a
shrunken variant with the same topology (used so CPU-side tests finish in
seconds and so the CUDA path is exercised on a second set of tile shapes).
"""
import copy


def _stage(modules, channels):
    n = len(channels)
    return {'num_modules': modules, 'num_branches': n, 'block': 'basic',
            'num_blocks': [4] * n, 'num_channels': list(channels),
            'fuse_method': 'sum'}


def make_cfgs(widths=(48, 96, 192, 384), input_size=(256, 256),
              heatmap_size=(64, 64), modules=(1, 4, 3), head_type='coordinates',
              num_joints=33, neurons=1024, lifter_blocks=2):
    return {
        'FCModel': {'name': 'lifter', 'refine_3d': False, 'norm_twoD': False,
                    'num_blocks': lifter_blocks, 'input_size': 2 * num_joints,
                    'output_size': 3 * (num_joints - 1), 'num_neurons': neurons,
                    'dropout': 0.5, 'leaky': False},
        'heatmapModel': {
            'name': 'hrnet', 'add_xy': False,
            'input_size': list(input_size), 'head_type': head_type,
            'pixel_shuffle': False, 'heatmap_size': list(heatmap_size),
            'init_weights': True, 'num_joints': num_joints, 'pretrained': '',
            'extra': {
                'pretrained_layers': ['*'], 'final_conv_kernel': 1,
                'stage2': _stage(modules[0], widths[:2]),
                'stage3': _stage(modules[1], widths[:3]),
                'stage4': _stage(modules[2], widths[:4]),
            },
        },
    }


def demo_cfgs(head_type='coordinates'):
    return make_cfgs(head_type=head_type)


def tiny_cfgs(head_type='coordinates'):
    """W16 widths, 128x128 crops -> 32x32 heat-maps, one module per stage."""
    return make_cfgs(widths=(16, 32, 64, 128), input_size=(128, 128),
                     heatmap_size=(32, 32), modules=(1, 1, 1),
                     head_type=head_type, neurons=256, lifter_blocks=2)


def ped_cfgs():
    """W32 widths, 192(w)x256(h) crops, 48x64 maps: the shape family of
    ``configs/KITTI_train_IGRs_Ped.yml:72-83,127-156`` (non-square tiles)."""
    return make_cfgs(widths=(32, 64, 128, 256), input_size=(192, 256),
                     heatmap_size=(48, 64), modules=(1, 4, 3))


def clone(cfgs):
    return copy.deepcopy(cfgs)
