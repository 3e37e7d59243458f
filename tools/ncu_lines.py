"""Dynamic per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda,sass`.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X_src.csv
    python tools/ncu_lines.py X_src.csv [min_pct]
"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    cut = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
    per, tot_i, tot_s = {}, 0, 0
    for k, h in enumerate(hi):
        f = rows[h - 2][1].split('/')[-1]
        hdr = rows[h]
        iI, iS, iT = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
        end = hi[k + 1] - 2 if k + 1 < len(hi) else len(rows)
        for r in rows[h + 1:end]:
            if len(r) <= iT or not r[0].strip().isdigit():
                continue
            try:
                ins, s, th = int(r[iI] or 0), int(r[iS] or 0), int(r[iT] or 0)
            except ValueError:
                continue
            key = (f, int(r[0]))
            a = per.setdefault(key, [0, 0, 0, r[1].strip()[:95]])
            a[0] += ins; a[1] += s; a[2] += th
            tot_i += ins; tot_s += s
    print(f'total warp instructions {tot_i:.4e}  stall samples {tot_s}')
    for (f, l), (ins, s, th, src) in per.items():
        if 100 * ins / tot_i > cut or 100 * s / max(tot_s, 1) > cut:
            print(f'{f}:{l:4d} inst {100*ins/tot_i:5.2f}% samp {100*s/max(tot_s,1):5.2f}% lanes {th/max(ins,1):5.1f} | {src}')


if __name__ == '__main__':
    main()
