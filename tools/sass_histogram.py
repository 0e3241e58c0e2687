"""SASS opcode histogram of the shipped library: `python tools/sass_histogram.py > profiles/rNN_sass_opcodes.txt` (needs cuobjdump, no GPU).

Per kernel: instruction count and the opcodes that prove what the kernel runs on - DMMA (FP64 tensor core), UTMALDG / UBLKCP (TMA tiled and
bulk copies), SYNCS (mbarrier), USETMAXREG (register hand-over), LDGSTS (cp.async), DFMA / DADD / DMUL (FP64 pipe), MUFU - then the
library-wide histogram."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / 'rom-comma_b200' / 'csrc' / 'librc_b200.so'
sass = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True, check=True).stdout
total, per, fn = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and fn:
        total[m.group(1)] += 1
        per[fn][m.group(1)] += 1
demangle = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(per, demangle)) if len(demangle) == len(per) else {k: k for k in per}
KEYS = ['DMMA', 'UTMALDG', 'UBLKCP', 'SYNCS', 'USETMAXREG', 'LDGSTS', 'DFMA', 'DADD', 'DMUL', 'MUFU', 'ATOM', 'RED', 'SHFL', 'LDS', 'STS', 'LDG', 'STG']
print(f'{lib.name}: {len(per)} kernels, {sum(total.values())} SASS instructions (sm_100a)\n')
print('per kernel: total | ' + ' '.join(KEYS))
for f in sorted(per, key=lambda k: -sum(per[k].values())):
    counts = {k: sum(v for op, v in per[f].items() if op.split('.')[0] == k) for k in KEYS}
    short = re.sub(r'\(.*', '', names[f])
    print(f'{sum(per[f].values()):7d} | ' + ' '.join(f'{k}={counts[k]}' for k in KEYS if counts[k]) + f'   {short}')
print('\nlibrary-wide histogram (opcode with modifiers):')
for op, v in total.most_common():
    print(f'{v:7d} {op}')
