#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel of libchord_b200.so (no GPU needed):

    python profiles/tools/sass_static_lines.py chord_b200/libchord_b200.so raster k_raster_fastILi32ELi7 600 900

Arguments: the .so, the translation unit (raster / cull / hzb / resolve / shadow / ctx), a substring of the mangled kernel
name, and the source-line range to print.  For an issue-bound kernel with a flat profile the executed-instruction count is
the cost model, so counting the straight-line parts before and after a change is a usable proxy when GPU time is short."""
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, tu, kernel, lo, hi = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    tmp = tempfile.mkdtemp()
    subprocess.run(f'cd {tmp} && cuobjdump -xelf all {so} > /dev/null 2>&1 && nvdisasm -g -c {tu}.sm_100a.cubin > all.sass', shell=True, check=True)
    txt = open(os.path.join(tmp, 'all.sass')).read()
    i = txt.index('.text.' + txt.split('.text.')[[k for k, s in enumerate(txt.split('.text.')) if kernel in s.split('\n')[0]][0]].split('\n')[0].rstrip(' -:'))
    j = txt.find('//--------------------- .text.', i + 100)
    cur, per = None, {}
    for l in txt[i:j if j > 0 else None].split('\n'):
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
            per[cur] = per.get(cur, 0) + 1
    src_dir = os.path.join(os.path.dirname(so), 'csrc')
    cache = {}
    total = 0
    for k in sorted((k for k in per if k), key=lambda k: (k[0], k[1])):
        if lo <= k[1] <= hi:
            if k[0] not in cache:
                p = os.path.join(src_dir, k[0])
                cache[k[0]] = open(p).read().split('\n') if os.path.exists(p) else []
            text = cache[k[0]][k[1] - 1].strip()[:110] if len(cache[k[0]]) >= k[1] else ''
            print(f'{k[0]}:{k[1]:5d} {per[k]:4d} | {text}')
            total += per[k]
    print(f'total in range {total}, kernel total {sum(per.values())}')


if __name__ == '__main__':
    main()
