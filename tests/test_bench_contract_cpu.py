"""bench.py contract pieces that run without a GPU: the reference arm (CPU oracle port) prints one valid JSON
line with the keys the driver reads, and the roofline block is assembled correctly from per-class timings."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_json():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1', '--ref-batch', '1'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'crops/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('crops/sec') and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'crops/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and d['vs_baseline'] is None and 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_roofline_block_from_class_timings():
    sys.path.insert(0, ROOT)
    import bench
    pk = {'hbm_gbs': 6500.0, 'bf16_tflops': 1600.0, 'bf16_tflops_sustained': 1400.0, 'source': 'measured'}
    classes = {
        'conv_tc 3x3 s1 48->48 @64x64+res': {'ms': 3.2, 'launches': 32, 'macs': 32 * 10 ** 10, 'act_bytes': 32 * 3 * 10 ** 8,
                                            'weight_bytes': 32 * 5 * 10 ** 4},
        'fuse': {'ms': 1.0, 'launches': 28, 'macs': 0, 'act_bytes': 28 * 2 * 10 ** 8, 'weight_bytes': 0},
    }
    blk = bench.roofline_block(classes, 24.0, pk, 256)
    assert blk['kernel'].startswith('conv_tc 3x3 s1 48->48') and blk['bound'] == 'hbm' and blk['unit'] == 'GB/s'
    assert abs(blk['avg_launch_us'] - 100.0) < 1e-6
    assert abs(blk['achieved'] - (3e8 + 5e4) / 100e-6 / 1e9) < 0.1
    assert abs(blk['frac'] - blk['achieved'] / 6500.0) < 1e-3 and blk['peak'] == 6500.0
    assert blk['traffic'] is None or blk['traffic'] > 0          # from profiles/ncu_traffic.json when the batch matches
    assert abs(blk['share_of_hc_time'] - 3.2 / 24.0) < 1e-3
    assert [c['kernel'] for c in blk['top_classes']][0] == blk['kernel']
