#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small, committed summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r11_launches.csv profiles/r1_launches.md
    python tools/ncu_summary.py kernel   gpurun_out/r10_prof_star.ncu-rep profiles/r1_gat_kn_star.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(row['Metric Unit'], v)
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, '
                'serialised launches: compare SHARES, not absolutes.\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n')
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f'| `{k[:90]}` | {cnt[k]} | {v / 1e3:.3f} | {v / cnt[k]:.1f} | {100 * v / T:.1f}% |\n')


def kernel(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full summary ({src})\n')
        for r in rows[2:]:
            f.write(f"\n## `{r[idx['Kernel Name']][:120]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f'| {k} | {r[idx[k]]} | {units[idx[k]]} |\n')
            stalls = []
            for h in hdr:
                if 'issue_stalled' in h and h.endswith('_per_warp_active.pct') and r[idx[h]]:
                    stalls.append((float(r[idx[h]]), h.split('issue_stalled_')[1].split('_per_warp')[0]))
            if stalls:
                f.write('\nTop warp stall reasons (% of warp-active): ' +
                        ', '.join(f'{n} {v:.1f}' for v, n in sorted(stalls, reverse=True)[:6]) + '\n')


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
