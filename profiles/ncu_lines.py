"""Per-CUDA-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on):
    python profiles/ncu_lines.py file.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys


def _i(v):
    try:
        return int(v)
    except ValueError:
        return 0


def main(path, top=40):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    agg = collections.OrderedDict()
    fname, hdr, total = '', None, 0
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = {h: i for i, h in enumerate(r)}
            cur = None
        elif hdr is not None and len(r) >= len(hdr):
            if r[0] != '':
                cur = (fname, r[0], r[1].strip()[:100])
            samples = _i(r[hdr['# Samples']])
            inst = _i(r[hdr['Instructions Executed']])
            a = agg.setdefault(cur, [0, 0, collections.Counter()])
            a[0] += samples
            a[1] += inst
            for h, i in hdr.items():
                if h.startswith('stall_') and 'Not Issued' not in h and r[i] not in ('', '0'):
                    a[2][h[6:]] += _i(r[i])
            total += samples
    print('total samples', total)
    for k, (s, n, why) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        w = ' '.join(f'{a}:{b}' for a, b in why.most_common(3))
        print(f'{s:6d} {100 * s / max(total, 1):5.1f}% inst {n:9d}  {k[0]}:{k[1]:>4}  {k[2][:80]:80} {w}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
