#!/usr/bin/env python
"""Blackwell evidence: per kernel of libjlm_b200.so, how many tcgen05 / TMEM / TMA instructions its SASS holds.
    python scripts/sass_summary.py > profiles/r02/sass_summary.txt
SASS mnemonics (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG,
tcgen05.commit -> UTCBAR, cp.async -> LDGSTS, mbarrier -> SYNCS."""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         'jlm_b200', 'csrc', 'libjlm_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
PAT = ['UTCHMMA.2CTA', 'UTCHMMA', 'LDTM', 'UTMALDG.2D.2CTA', 'UTMALDG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'DFMA', 'MUFU.EX2', 'HMMA']
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(anonymous namespace\)::', '', name)
        name = re.sub(r'\(.*$', '', name)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if not m:
        continue
    op = m.group(1)
    counts[name]['instructions'] += 1
    for p in PAT:
        if op.startswith(p):
            counts[name][p] += 1
            break
print('arch: sm_100a   library: %s' % os.path.basename(so))
print('%-62s %6s' % ('kernel', 'instr') + ''.join(' %8s' % p[:8] for p in PAT))
for n, c in counts.items():
    if c['instructions'] == 0:
        continue
    print('%-62s %6d' % (n[:62], c['instructions']) + ''.join(' %8d' % c[p] for p in PAT))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print('%-62s %6d' % ('TOTAL', tot['instructions']) + ''.join(' %8d' % tot[p] for p in PAT))
