"""Warp-stall samples of an .ncu-rep aggregated by CUDA source line (the CSV source page only carries SASS addresses; the
line of every SASS instruction comes from nvdisasm -g on the cubin inside the in-tree .so, built with -lineinfo).
usage: python profiles/ncu_stalls_by_line.py rep.ncu-rep <cubin name, e.g. vit_attention> <kernel substring> [n]"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'attentionshift_b200', 'csrc', 'libattnshift_b200.so')


def line_table(cubin_name, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', SO], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.startswith(cubin_name)][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
    table, on, cur = {}, False, None
    for ln in txt:
        if ln.startswith('//---') and '.text.' in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), 'inlined' in m.group(3))
            continue
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*);', ln)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main(rep, cubin_name, kernel, n=40):
    table = line_table(cubin_name, kernel)
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr) and r[ci['# Samples']].replace('.', '').isdigit()]
    base = min(int(r[ci['Address']], 16) for r in body)
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg, tot = {}, 0.0
    for r in body:
        off = int(r[ci['Address']], 16) - base
        key = table.get(off, ('?', 0, False))[:2]
        s = float(r[ci['# Samples']])
        tot += s
        a = agg.setdefault(key, {'n': 0.0, 'inst': 0.0})
        a['n'] += s
        a['inst'] += float(r[ci['Instructions Executed']] or 0)
        for h in stall_cols:
            a[h] = a.get(h, 0.0) + float(r[ci[h]])
    src = {}
    print(f'total samples {tot:.0f}')
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]['n'])[:n]:
        if f not in src:
            pth = os.path.join(ROOT, 'attentionshift_b200', 'csrc', f)
            src[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
        text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ''
        top = sorted(stall_cols, key=lambda h: -a.get(h, 0))[:2]
        why = ' '.join(f'{h[6:]}={100 * a[h] / a["n"]:.0f}%' for h in top)
        print(f'{100 * a["n"] / tot:5.1f}%  {f}:{ln:<4d} inst {a["inst"]:>10.0f}  {why:32s} | {text}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
