"""Stall-reason totals and hottest SASS instructions of an .ncu-rep (needs --import-source on / -lineinfo):
    python profiles/ncu_stalls.py file.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    print(rows[0][1][:120] if len(rows[0]) > 1 else '')
    hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) >= len(rows[start])]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = {s: 0 for s in stalls}
    samp = 0
    inst = 0
    for r in data:
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
        samp += int(r[ix['# Samples']] or 0)
        inst += int(r[ix['Instructions Executed']] or 0)
    print('samples', samp, 'warp-instructions', inst)
    print('  '.join(f'{s[6:]} {100 * v / max(1, samp):.1f}%' for s, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
        why = {s[6:]: r[ix[s]] for s in stalls if r[ix[s]] not in ('0', '')}
        print(f"{r[ix['# Samples']]:>6} {r[ix['Instructions Executed']]:>9}  {r[ix['Source']][:70]:70} {why}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
