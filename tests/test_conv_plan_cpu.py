"""Plan of the tcgen05 convolutions (kernel choice and tiling per layer shape) -- geometry only, runs without a GPU
through egn_debug_conv_plan.  Pins which kernel serves each HRNet-W48 layer class in both tensor-core modes and the
invariants the kernels rely on (shared memory / TMEM budgets, ring depths)."""
import ctypes

import pytest

from egonet_b200 import _native as N


def plan(dtype, Cin, Cout, H, W, k, stride):
    buf = ctypes.create_string_buffer(512)
    N.check(N.lib().egn_debug_conv_plan(dtype, Cin, Cout, H, W, k, stride, buf, 512))
    d = {}
    for item in buf.value.decode().split():
        key, val = item.split('=')
        d[key] = val
    return d


# (Cin, Cout, H, W, k, stride) -> kernel in fp16x2 / fp16 (demo config, 256 x 256 crops)
HRNET_LAYERS = [
    ((48, 48, 64, 64, 3, 1), 'v3-persist', 'v3-persist'),
    ((96, 96, 32, 32, 3, 1), 'v4-tapwin', 'v3-persist'),
    ((192, 192, 16, 16, 3, 1), 'v4-tapwin', 'v1-tap'),
    ((384, 384, 8, 8, 3, 1), 'v1-tap', 'v1-tap'),
    ((64, 64, 64, 64, 3, 1), 'v3-persist', 'v3-persist'),
    ((64, 256, 64, 64, 1, 1), 'v3-persist', 'v3-persist'),
    ((256, 64, 64, 64, 1, 1), 'v1-tap', 'v3-persist'),
    ((64, 64, 128, 128, 3, 2), 'v1-tap', 'v1-tap'),
    ((48, 96, 64, 64, 3, 2), 'v1-tap', 'v1-tap'),
    ((256, 48, 64, 64, 3, 1), 'v4-tapwin', 'v1-tap'),
]


@pytest.mark.parametrize('shape,k_split,k_plain', HRNET_LAYERS, ids=['%dx%d_%dx%d_k%ds%d' % s[0] for s in HRNET_LAYERS])
def test_kernel_choice_per_layer(shape, k_split, k_plain, monkeypatch):
    for var in ('EGN_TC_V3', 'EGN_TC_V2', 'EGN_TC_V4', 'EGN_TC_BLK', 'EGN_TC_V4_PAIR', 'EGN_TC_V4_FOLD', 'EGN_TC_V4_PERSIST'):
        monkeypatch.delenv(var, raising=False)
    assert plan(2, *shape)['kernel'] == k_split
    assert plan(1, *shape)['kernel'] == k_plain


def test_fp16x2_plans_of_the_hot_layers(monkeypatch):
    for var in ('EGN_TC_V3', 'EGN_TC_V4', 'EGN_TC_V4_PAIR', 'EGN_TC_V4_FOLD', 'EGN_TC_V4_PERSIST', 'EGN_TC_ASW', 'EGN_TC_ASLOTS'):
        monkeypatch.delenv(var, raising=False)
    p48 = plan(2, 48, 48, 64, 64, 3, 1)
    # block-shaped windows, 64-byte window rows (three exact chunks), two window slots, staged TMA epilogue
    assert (p48['blk'], p48['a_sw'], p48['a_slots'], p48['resident'], p48['stage']) == ('8x16', '64', '2', '1', '2')
    p96 = plan(2, 96, 96, 32, 32, 3, 1)
    assert (p96['persist'], p96['pair'], p96['fold'], p96['staged']) == ('1', '1', '1', '1')
    p192 = plan(2, 192, 192, 16, 16, 3, 1)
    # two 96-channel parts per work unit: two accumulator sets fit, so the tile loop is persistent
    assert (p192['n_tile'], p192['fold'], p192['persist'], p192['pair']) == ('96', '2', '1', '1')


SHAPES = [(Cin, Cout, H, W, 3, 1) for (Cin, Cout) in ((16, 16), (32, 32), (48, 48), (64, 64), (66, 66), (96, 96), (128, 128),
                                                       (192, 192), (256, 48), (48, 256))
          for (H, W) in ((16, 16), (32, 32), (16, 8), (24, 20), (64, 64))]


@pytest.mark.parametrize('dtype', [1, 2])
def test_budgets_and_ring_invariants(dtype, monkeypatch):
    monkeypatch.setenv('EGN_TC_V4', '2')                # tap-window plans for plain fp16 too
    for force_v4 in (False, True):
        if force_v4:
            monkeypatch.setenv('EGN_TC_V3', '0')
            monkeypatch.setenv('EGN_TC_V2', '0')
        for shape in SHAPES:
            p = plan(dtype, *shape)
            assert int(p['smem']) <= 227 * 1024, (shape, p)
            assert int(p['tmem']) <= 512 and int(p['tmem']) & (int(p['tmem']) - 1) == 0, (shape, p)
            if p['kernel'] == 'v4-tapwin':
                assert 1 <= int(p['na']) <= 4 and 2 <= int(p['nb']) <= 16, (shape, p)
                if p['persist'] == '1':
                    # a persistent CTA walks over many tiles: a single window slot would be refilled while the MMAs of
                    # its last tap are still waiting for their weight tile (deadlock)
                    assert int(p['na']) >= 2, (shape, p)
                    assert p['staged'] == '1', (shape, p)
                if p['pair'] == '1':
                    assert int(p['n_tile']) % 32 == 0, (shape, p)
            if p['kernel'] == 'v3-persist':
                assert p['a_slots'] in ('1', '2'), (shape, p)
