#!/usr/bin/env python
"""Per-source-line (and per-phase) instruction / stall-sample shares of one kernel from an ncu report captured with
`--set full --import-source on` (the binary must be built with -lineinfo).

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python profiles/tools/ncu_source_lines.py src.csv [top_n] ['{"vertex phase": ["raster.cu", 740, 770], ...}']

The third argument groups line ranges of a file into named phases.  Used for profiles/r01_v5_raster_warp_per_meshlet.md and
profiles/r01_frame_deferred_large_source.md."""
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    cur, hdr, func, out = None, None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) == 2 and r[0] == 'Function Name':
            func = r[1].split('(')[0].split('::')[-1]
            continue
        if r and r[0] == 'Line No':
            hdr = r
            continue
        if hdr and r and r[0] not in ('',):
            try:
                out.append(dict(func=func, file=cur, line=int(r[0]), src=r[1], samples=int(r[hdr.index('# Samples')]),
                                inst=int(r[hdr.index('Instructions Executed')]), thr=int(r[hdr.index('Thread Instructions Executed')])))
            except (ValueError, IndexError):
                pass
    return out


def main():
    recs = load(sys.argv[1])
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    phases = json.loads(sys.argv[3]) if len(sys.argv) > 3 else None
    for func in sorted({r['func'] for r in recs}):
        rs = [r for r in recs if r['func'] == func]
        ti, ts = sum(r['inst'] for r in rs), max(sum(r['samples'] for r in rs), 1)
        print(f'===== {func}: {ti} warp instructions, {ts} stall samples')
        if phases:
            acc = {k: [0, 0] for k in phases}
            other = [0, 0]
            for r in rs:
                for k, (f, a, b) in phases.items():
                    if r['file'] == f and a <= r['line'] <= b:
                        acc[k][0] += r['inst']; acc[k][1] += r['samples']
                        break
                else:
                    other[0] += r['inst']; other[1] += r['samples']
            for k, v in list(acc.items()) + [('other', other)]:
                print(f'{k:32s} inst {100 * v[0] / ti:5.1f}%  samples {100 * v[1] / ts:5.1f}%')
        for r in sorted(rs, key=lambda r: -r['inst'])[:top]:
            print(f"{r['file']}:{r['line']:5d} inst {100 * r['inst'] / ti:5.1f}% samp {100 * r['samples'] / ts:5.1f}% "
                  f"thr/inst {r['thr'] / max(r['inst'], 1):4.1f} | {r['src'][:110]}")


if __name__ == '__main__':
    main()
