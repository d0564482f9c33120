#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libchord_b200.so (no GPU needed):

    python profiles/tools/sass_opcode_histogram.py chord_b200/libchord_b200.so resolve k_lightingILb1

Prints the instruction count, the opcode histogram and the resource usage (registers / stack) of the kernel whose mangled
name contains the given substring.  Used for before / after tables when a change cannot be timed (profiles/*_after_budget.md)."""
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, tu, kernel = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3]
    tmp = tempfile.mkdtemp()
    subprocess.run(f'cd {tmp} && cuobjdump -xelf all {so} > /dev/null 2>&1 && nvdisasm -c {tu}.sm_100a.cubin > all.sass', shell=True, check=True)
    txt = open(os.path.join(tmp, 'all.sass')).read()
    parts = txt.split('//--------------------- .text.')
    body = [p for p in parts[1:] if kernel in p.split('\n')[0]][0]
    ops = {}
    for l in body.split('\n'):
        m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', l)
        if m:
            ops[m.group(1)] = ops.get(m.group(1), 0) + 1
    print('kernel', body.split('\n')[0].strip(' -'))
    print('instructions', sum(ops.values()))
    print(' '.join(f'{k}:{v}' for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))
    ru = subprocess.run(['cuobjdump', '--dump-resource-usage', so], capture_output=True, text=True).stdout.split('\n')
    for i, l in enumerate(ru):
        if kernel in l and i + 1 < len(ru):
            print(ru[i + 1].strip())


if __name__ == '__main__':
    main()
