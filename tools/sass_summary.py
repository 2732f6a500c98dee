"""Per-kernel count of the Blackwell mnemonics (tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM, cp.async.bulk.tensor = UTMALDG,
cp.async.bulk = UBLKCP, tcgen05.commit = UTCBAR, mbarrier = SYNCS) in the built librelnet_b200.so, with registers / spills
from the ptxas -v log of the same build.  Runs without a GPU:

    python tools/sass_summary.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'relation-networks-for-object-detection_b200')
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS']
PAT = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCIMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UBLKCP|SYNCS|MUFU\.\w+)\b')


def short(mangled):
    d = subprocess.run(['c++filt', mangled], capture_output=True, text=True).stdout.strip()
    d = re.sub(r'^void ', '', d.replace('(anonymous namespace)::', ''))
    depth = 0
    for i, ch in enumerate(d):
        depth += (ch == '<') - (ch == '>')
        if ch == '(' and depth == 0:
            return d[:i]
    return d


def main():
    sass = subprocess.run(['cuobjdump', '-sass', os.path.join(PKG, 'librelnet_b200.so')], capture_output=True, text=True).stdout
    cnt, kern = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            kern = m.group(1)
        elif kern:
            for t in PAT.findall(line):
                cnt[kern][t] += 1
    regs, cur = {}, None
    for line in open(os.path.join(PKG, 'build', 'ptxas.log')):
        m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', line)
        if m and cur:
            regs.setdefault(cur, {})['spill'] = (int(m.group(2)), int(m.group(3)))
        m = re.search(r'Used (\d+) registers', line)
        if m and cur:
            regs.setdefault(cur, {})['regs'] = int(m.group(1))
    tot, rows = collections.Counter(), []
    for k, c in cnt.items():
        for x in KEYS:
            tot[x] += c[x]
        if any(c[x] for x in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'UTCBAR')):
            r = regs.get(k, {})
            rows.append((short(k), [c[x] for x in KEYS], {m: v for m, v in c.items() if m.startswith('MUFU')}, r.get('regs'), r.get('spill')))
    rows.sort()
    print('SASS mnemonics of the shipped librelnet_b200.so (cuobjdump -sass, sm_100a; built by build.py from this tree) for every kernel')
    print('that touches the tensor core (UTCHMMA = tcgen05.mma kind::f16 / kind::tf32), TMEM (LDTM = tcgen05.ld), TMA (UTMALDG =')
    print('cp.async.bulk.tensor, UBLKCP = cp.async.bulk) or tcgen05.commit (UTCBAR); SYNCS = mbarrier ops.  regs / spill bytes from ptxas -v.\n')
    print('%-44s ' % 'kernel' + ' '.join('%7s' % x for x in KEYS) + '  regs spill(st,ld)  MUFU')
    for n, v, mu, rg, sp in rows:
        print('%-44s ' % n[:44] + ' '.join('%7d' % x for x in v) + '  %4s %-12s  ' % (rg, '%d,%d' % sp if sp else '-') +
              ', '.join('%s:%d' % (a.replace('MUFU.', ''), b) for a, b in sorted(mu.items())))
    print('%-44s ' % 'TOTAL (whole library)' + ' '.join('%7d' % tot[x] for x in KEYS))
    print('\nNo STTM / UTMASTG: P goes to shared memory through registers (it needs the per-row rescale), outputs are plain vector stores.')
    print('kind::tf32 and kind::f16 share the UTCHMMA mnemonic; gemm_tf32_tc_kernel is the tf32 engine of the backward.')


if __name__ == '__main__':
    main()
