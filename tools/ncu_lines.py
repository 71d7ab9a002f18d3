"""Aggregate the SASS rows of an `ncu --import-source on` capture per CUDA source line.

    ncu -i X.ncu-rep --page source --csv --print-source sass,cuda > /tmp/src.csv
    python tools/ncu_lines.py /tmp/src.csv [top]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    out, cur, hdr, idx, line = {}, None, None, None, None
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            idx = {h: i for i, h in enumerate(hdr)}
            continue
        if not hdr or len(r) != len(hdr):
            continue
        if r[0].isdigit():
            line = (cur, int(r[0]), r[1].strip()[:72])
            out.setdefault(line, [0, 0, 0, 0])
            continue
        g = lambda k: int(float(r[idx[k]])) if r[idx[k]] not in ('', '-') else 0
        acc = out[line]
        acc[0] += g('# Samples'); acc[1] += g('Instructions Executed')
        acc[2] += g('L1 Wavefronts Shared'); acc[3] += g('L1 Wavefronts Shared Ideal')
    tot = sum(v[0] for v in out.values()) or 1
    toti = sum(v[1] for v in out.values()) or 1
    print('samples', tot, 'warp instructions', toti)
    for k, v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k[0]:18s} {k[1]:4d} smp {100*v[0]/tot:5.1f}% inst {100*v[1]/toti:5.1f}% "
              f"smem wf {v[2]/1e6:6.2f}/{v[3]/1e6:6.2f}  {k[2]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
