#!/usr/bin/env python
"""DRAM traffic of the K_n aggregate kernel from an `ncu --set full` report -> profiles/gat_kn_traffic.json, which bench.py
quotes as `roofline.traffic`.

    python tools/ncu_traffic.py <report.ncu-rep> <n> <instances in the profiled launch> [out.json]

The JSON names the kernel it was measured on; bench.py ignores it when the name or n differs from what it runs."""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, n, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    out = sys.argv[4] if len(sys.argv) > 4 else 'profiles/gat_kn_traffic.json'
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def nbytes(key):
        u, v = col[key]
        return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
    kernel = col['Kernel Name'][1].split('(')[0].split('::')[-1]
    rd, wr = nbytes('dram__bytes_read.sum'), nbytes('dram__bytes_write.sum')
    u, t = col['gpu__time_duration.sum']
    ms = float(t) * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}[u]
    d = {'kernel': kernel, 'n': n, 'instances': B, 'dram_bytes_read': rd, 'dram_bytes_written': wr,
         'dram_bytes_per_instance_layer': (rd + wr) / B, 'kernel_ms_under_ncu': ms, 'source': rep.split('/')[-1]}
    json.dump(d, open(out, 'w'), indent=1)
    print(json.dumps(d))


if __name__ == '__main__':
    main()
