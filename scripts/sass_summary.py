"""Per-kernel SASS summary of the built library (CPU only): instruction totals and the Blackwell-specific opcodes.
   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, 'deephumor_b200', 'libdeephumor_sm100.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
OPS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'FFMA2', 'FADD2', 'FMUL2', 'HMMA',
       'MUFU.TANH', 'STAS']
print('# cuobjdump -sass deephumor_b200/libdeephumor_sm100.so (sm_100a): per-kernel instruction totals and the Blackwell-specific opcodes')
print('# UTCHMMA = tcgen05.mma kind::f16, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA tensor load / store,')
print('# UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc, FFMA2 / FADD2 / FMUL2 = packed fp32 pairs, HMMA = mma.sync,')
print('# STAS = asynchronous store into another CTA of the cluster (st.async with transaction bytes, remote mbarrier arrive)')
print()
cur, body = None, []
def flush():
    if cur is None:
        return
    name = subprocess.run(['c++filt', cur], capture_output=True, text=True).stdout.strip()
    name = name.replace('(anonymous namespace)::', '').replace('void ', '')
    name = re.sub(r'\(CUtensorMap_st.*', '(...)', name)
    insts = [l for l in body if re.search(r'/\*[0-9a-f]{4,}\*/', l)]
    cnt = collections.Counter()
    first = {}
    for l in insts:
        m = re.search(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if not m:
            continue
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + '.') or (o == 'MUFU.TANH' and op.startswith('MUFU.TANH')):
                cnt[o] += 1
                first.setdefault(o, re.sub(r'\s+', ' ', l.split('*/', 1)[1].split('/*')[0]).strip(' ;'))
    print(name)
    print(f'    instructions {len(insts):6d}  ' + '  '.join(f'{o} {cnt[o]}' for o in OPS if cnt[o]))
    for o in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'STAS'):
        if o in first:
            print(f'      {first[o]}')
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        flush()
        cur, body = m.group(1), []
    elif cur is not None:
        body.append(line)
flush()
