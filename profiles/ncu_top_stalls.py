"""Top stalled SASS instructions of a kernel from an .ncu-rep (source page).  usage: ncu_top_stalls.py rep [n]"""
import csv
import subprocess
import sys


def main(path, n=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    body = [r for r in body if r[ci['# Samples']].replace('.', '').isdigit()]
    tot = sum(float(r[ci['# Samples']]) for r in body)
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(float(r[ci[h]]) for r in body) for h in stall_cols}
    print('total samples', tot)
    print('by reason:', ', '.join(f'{k[6:]}={100 * v / tot:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(body, key=lambda r: -float(r[ci['# Samples']]))[:n]:
        top = max(stall_cols, key=lambda h: float(r[ci[h]]))
        print(f"{100 * float(r[ci['# Samples']]) / tot:5.1f}%  {top[6:]:12s} {r[ci['Source']].strip()[:90]}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
