"""SASS census of libddk.so: per kernel, counts of the mnemonics that show which hardware path it uses.
usage: cuobjdump -sass disco_diffdock_b200/libddk.so | python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import re
import sys

cur = None
cnt = collections.OrderedDict()
pat = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        cnt[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        cnt[cur][m.group(1).split('.')[0]] += 1
keys = ['UTCHMMA', 'UTCBAR', 'STTM', 'LDTM', 'UBLKCP', 'LDGSTS', 'SYNCS', 'ELECT', 'FFMA2', 'FFMA', 'LDS', 'STS', 'SHFL', 'BAR', 'NANOSLEEP']
print('SASS census of disco_diffdock_b200/libddk.so (cuobjdump -sass, sm_100a): instructions per kernel by mnemonic (prefix before the first dot).')
print('UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, STTM / LDTM = tcgen05.st / ld (tensor memory), UBLKCP = cp.async.bulk (TMA engine),')
print('LDGSTS = cp.async, SYNCS = mbarrier operations, ELECT = elect.sync, FFMA2 = packed fp32 FMA')
print(f"{'kernel':44s} {'total':>7s} " + ' '.join(f'{k:>7s}' for k in keys))
for f, c in cnt.items():
    name = re.sub(r'^_ZN3ddk\d+', '', f)
    name = re.sub(r'ILi(\d+)EEEv.*', r'<\1>', name)
    name = re.sub(r'ILb(\d+)EEEv.*', r'<\1>', name)
    name = re.sub(r'Ev?NS_.*|Ev.*', '', name)
    print(f'{name[:44]:44s} {sum(c.values()):7d} ' + ' '.join(f'{c.get(k, 0):7d}' for k in keys))
