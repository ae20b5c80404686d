"""Opcode histogram of every kernel in libmultinn_sm100.so (cuobjdump -sass), written to profiles/<tag>_sass_opcodes.md:
the evidence that the tensor-core kernels are tcgen05 / TMEM / TMA code (UTCHMMA / UTCQMMA, LDTM, UTMALDG, UTMASTG,
UTMAREDG, SYNCS) and which kernels are SIMT. Runs without a GPU.   python tools/sass_histograms.py [tag]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r2'
so = os.path.join(ROOT, 'multinn_b200', 'libmultinn_sm100.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', line)
    if m and cur is not None:
        cur[m.group(1).split('.')[0]] += 1
demangle = subprocess.run(['c++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()
KEY = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTCBAR', 'SYNCS', 'MUFU', 'FFMA',
       'FFMA2', 'FADD', 'HMMA', 'LDS', 'STS', 'LDG', 'STG', 'SHFL', 'BAR', 'ATOMS', 'RED']
out = [f'# SASS opcode counts per kernel ({tag}; cuobjdump -sass of the shipped libmultinn_sm100.so, sm_100a)\n',
       'Static instruction counts. Tensor-core kernels show UTCHMMA (tcgen05.mma kind::tf32 / kind::f16), LDTM '
       '(tcgen05.ld), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce), SYNCS (mbarrier); SIMT kernels show none.\n',
       '| kernel | total | ' + ' | '.join(KEY) + ' |', '|---|---|' + '---|' * len(KEY)]
for (mangled, c), name in zip(kernels.items(), demangle):
    short = re.sub(r'\(.*', '', name)[:90]
    out.append(f'| `{short}` | {sum(c.values())} | ' + ' | '.join(str(c.get(k, 0)) for k in KEY) + ' |')
path = os.path.join(ROOT, 'profiles', f'{tag}_sass_opcodes.md')
open(path, 'w').write('\n'.join(out) + '\n')
print(path, len(kernels), 'kernels')
