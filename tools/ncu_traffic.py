"""profiles/<tag>_trace_kernel_traffic.json from an `ncu --set full` capture of the trace kernel.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv
    python tools/ncu_traffic.py X_raw.csv profiles/r02_trace_kernel_traffic.json "what was captured"

The file carries the hash of the CUDA sources at the time of writing (bench.source_sha16): bench.py reports
`roofline.traffic` only when that hash equals the hash of the sources of the build it is running.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if 'trace' in d.get('Kernel Name', '') and (best is None or float(d['gpu__time_duration.sum']) > float(best['gpu__time_duration.sum'])):
            best = d
    u = dict(zip(hdr, units))

    def in_bytes(key):
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[key]]
        return float(best[key])*scale

    out = {'kernel': best['Kernel Name'], 'dram_bytes_per_launch': int(in_bytes('dram__bytes_read.sum') + in_bytes('dram__bytes_write.sum')),
           'dram_bytes_read': int(in_bytes('dram__bytes_read.sum')), 'dram_bytes_write': int(in_bytes('dram__bytes_write.sum')),
           'launch_ms': float(best['gpu__time_duration.sum']), 'source_sha16': bench.source_sha16(),
           'note': sys.argv[3] if len(sys.argv) > 3 else ''}
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
    print(out)


if __name__ == '__main__':
    main()
