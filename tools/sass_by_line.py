#!/usr/bin/env python
"""Attribute ncu's per-SASS-instruction counters to source lines without needing the source on the GPU box:

    python tools/sass_by_line.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [min_pct]

ncu's SASS page lists the kernel's instructions in address order; `nvdisasm -g` lists the same instructions with
`//## File ..., line N` markers.  The two are zipped by position."""
import csv
import io
import re
import subprocess
import sys


def main():
    rep, cubin, key = sys.argv[1:4]
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
    hdr = rows[hi]
    iI, iS, iSrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    inst = [(r[iSrc].strip(), int(r[iI]), int(r[iS]), [int(r[i] or 0) for i, _ in stall_cols]) for r in rows[hi + 1:] if len(r) > iI and r[iI].isdigit()]
    dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
    srcname = cubin.split('/')[-1].split('.')[0] + '.cu'
    start = next(i for i, l in enumerate(dis) if '.type' in l and key in l and '@function' in l)
    lines, cur = [], None
    for l in dis[start + 1:]:
        m = re.search(r'//## File "(.*?)", line (\d+)', l)
        if m:
            if m.group(1).endswith(srcname):          # lines of inlined library headers keep the enclosing source line
                cur = int(m.group(2))
            continue
        if '.type' in l and '@function' in l:
            break
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
            lines.append(cur)
    if len(lines) != len(inst):
        print(f'warning: {len(lines)} disassembled instructions vs {len(inst)} profiled', file=sys.stderr)
    tot = sum(i[1] for i in inst)
    tot_s = sum(i[2] for i in inst)
    by = {}
    for ln, (_, n, s, st) in zip(lines, inst):
        a = by.setdefault(ln, [0, 0, [0] * len(stall_cols)])
        a[0] += n; a[1] += s
        a[2] = [x + y for x, y in zip(a[2], st)]
    src = open('/root/repo/gnngls_b200/csrc/' + srcname).read().splitlines()
    print(f'total warp instructions {tot}, samples {tot_s}')
    if len(sys.argv) > 5:                               # phase totals: "name:first-last,name:first-last,..."
        for spec in sys.argv[5].split(','):
            name, rng = spec.split(':')
            a, b = (int(x) for x in rng.split('-'))
            n = sum(v[0] for k, v in by.items() if k is not None and a <= k <= b)
            sm = sum(v[1] for k, v in by.items() if k is not None and a <= k <= b)
            print(f'  phase {name:12s} lines {a:4d}-{b:4d}: {100 * n / tot:5.1f}% inst  {100 * sm / max(tot_s, 1):5.1f}% samples')
    for ln in sorted(k for k in by if k is not None):
        n, s, st = by[ln]
        if 100 * n / tot >= min_pct or 100 * s / max(tot_s, 1) >= min_pct:
            top = sorted(zip(st, [h for _, h in stall_cols]), reverse=True)[:2]
            tops = ' '.join(f'{h[6:]}={v}' for v, h in top if v)
            print(f'{ln:5d} {100 * n / tot:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% smpl  [{tops:34s}] {src[ln - 1].strip()[:100]}')


if __name__ == '__main__':
    main()
