"""Per-kernel census of the Blackwell-native SASS opcodes in libattnshift_b200.so (cuobjdump -sass):
UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA), LDTM / STTM (tcgen05.ld / st), HMMA (legacy mma.sync: expected 0).
    python profiles/sass_census.py > profiles/sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'attentionshift_b200', 'csrc', 'libattnshift_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ops = ('UTCHMMA', 'UTCQMMA', 'UTCMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'LDTM', 'STTM', 'HMMA', 'SYNCS', 'UTCBAR')
cur = None
tab = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        cur = re.sub(r'\(.*', '', name)
        tab[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        for o in ops:
            if op.startswith(o):
                key = o + ('.2CTA' if '.2CTA' in op and o == 'UTCHMMA' else '')
                tab[cur][key] += 1
        tab[cur]['_all'] += 1
cols = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'SYNCS', 'HMMA']
print('%-44s %s %8s' % ('kernel', ' '.join('%12s' % c for c in cols), 'instrs'))
tot = collections.Counter()
for k, c in tab.items():
    if not any(c[x] for x in cols):
        continue
    print('%-44s %s %8d' % (k[:44], ' '.join('%12d' % c[x] for x in cols), c['_all']))
    tot.update(c)
print('%-44s %s %8d' % ('TOTAL (kernels with any of the above)', ' '.join('%12d' % tot[x] for x in cols), tot['_all']))
print('kernels in the library: %d' % len(tab))
