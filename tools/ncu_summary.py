"""Summarise an .ncu-rep (ncu --set full) into a small markdown table under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xyz.md "title"
"""
import csv
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__t_sectors_srcunit_tex.sum', 'L2 sectors from SMs (x32 B)'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/TEX throughput %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem)'),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, 'w') as f:
        f.write('# %s\n\nSource: `%s` (`ncu --set full --clock-control none --import-source on`).\n\n' % (title, rep))
        for r in rows[2:]:
            f.write('## `%s` grid %s block %s\n\n| metric | value |\n|---|---|\n' % (
                r[idx['Kernel Name']][:90], r[idx['Grid Size']], r[idx['Block Size']]))
            for key, label in METRICS:
                if key in idx:
                    f.write('| %s (`%s`) | %s %s |\n' % (label, key, r[idx[key]], units[idx[key]]))
            f.write('\n')
    print('wrote', out)


if __name__ == '__main__':
    main()
