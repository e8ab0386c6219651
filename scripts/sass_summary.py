#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library: counts of the mnemonics that show which hardware path a kernel
uses (B200_PROFILING.md, "What proves a Blackwell-native kernel") and its spill instructions.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt

UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = tensor-map TMA load, UBLKCP = 1-D bulk TMA,
HMMA = mma.sync (legacy tensor path), LDSM/STSM = ldmatrix/stmatrix, STL/LDL = local-memory (spill) traffic.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'slotformer_b200', 'lib', 'libsfb200.so')
PATTERNS = [('UTC*MMA', r'\bUTC\w*MMA\b'), ('LDTM', r'\bLDTM\b'), ('STTM', r'\bSTTM\b'), ('UTMALDG', r'\bUTMALDG\b'),
            ('UBLKCP', r'\bUBLKCP\b'), ('HMMA', r'\bHMMA\b'), ('LDSM', r'\bLDSM\b'), ('STSM', r'\bSTSM\b'),
            ('FFMA2', r'\bFFMA2\b'), ('STL', r'\bSTL\b'), ('LDL', r'\bLDL\b')]


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    kernels, name, body = [], None, []
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            if name:
                kernels.append((name, '\n'.join(body)))
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        kernels.append((name, '\n'.join(body)))
    demangle = subprocess.run(['c++filt'], input='\n'.join(k for k, _ in kernels), capture_output=True, text=True).stdout.splitlines()
    print(f'# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels (cuobjdump -sass); linked libraries: ' +
          ', '.join(sorted(set(re.findall(r'NEEDED\)\s+Shared library: \[(.*?)\]',
                                          subprocess.run(['readelf', '-d', LIB], capture_output=True, text=True).stdout)))))
    hdr = f'{"kernel":<78}' + ''.join(f'{n:>8}' for n, _ in PATTERNS) + f'{"instrs":>8}'
    print(hdr)
    for (mangled, text), nice in sorted(zip(kernels, demangle), key=lambda t: t[1]):
        nice = re.sub(r'\(.*$', '', nice).replace('sfb::', '').replace('(anonymous namespace)::', '')
        n_instr = len(re.findall(r'/\*[0-9a-f]{4,}\*/\s+\S', text))
        print(f'{nice[:77]:<78}' + ''.join(f'{len(re.findall(p, text)):>8}' for _, p in PATTERNS) + f'{n_instr:>8}')


if __name__ == '__main__':
    sys.exit(main())
