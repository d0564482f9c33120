#!/usr/bin/env python
"""Regenerates tests/golden/golden_v1.json from the CPU oracle (python tests/golden/make_golden.py).

The reference (qiutang98/chord) ships no golden vectors for this path and cannot run offline, so these fixtures pin
the ORACLE's outputs on the seeded cases of tests/cases.py: sha256 of the post-cull list, the packed 64-bit
visibility surface, the HZB chains and the tile marker, plus a few human-checkable counts.  CPU tests compare the
oracle with them (so an accidental change of the oracle's arithmetic is caught); GPU tests compare the CUDA path."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
from cases import case_masked, case_ragged, case_s1, case_stream, city_frames, digest, gbuffer_digest  # noqa: E402


def entry_static(c, flags=0b111, prebuilt=False):
    fr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], c['w'], c['h'], flags)
    post = fr.instance_culling(c['lod0'])
    cmds = c['cmds'] if prebuilt else post
    vis, stats = fr.render_mesh(cmds)
    hz = fr.build_hzb(vis, True, True, True)
    extra = {}
    if c.get('name') == 's1' and flags == 0b111:
        extra['gbuffer'] = gbuffer_digest(fr, vis, post)              # K11: PBR + sky, six fp32 planes
    return dict(extra, post=digest(post), post_count=len(post), vis64=digest(vis), covered=int((vis != 0).sum()),
                hzb_min=digest(hz['min']), hzb_max=digest(hz['max']), valid_range=[int(x) for x in hz['range']],
                marker=digest(fr.visibility_mark(vis, post)) if not prebuilt else None,
                culled=stats['trianglesCulled'], clipped=stats['trianglesClipped'], submitted=stats['trianglesSubmitted'],
                alpha_killed=stats['fragmentsAlphaKilled'])


def main():
    g = {}
    for flags in (0b101, 0b111, 0b000):
        g[f's1_flags{flags}'] = entry_static(case_s1(), flags)
    for (n, w, h, s) in ((3000, 640, 360, 1.0), (1500, 512, 512, 40.0), (600, 256, 256, 2000.0)):
        g[f'stream_{n}_{w}x{h}_{s}'] = entry_static(case_stream(n, w, h, s), prebuilt=True)
    g['masked'] = entry_static(case_masked())
    g['ragged_400'] = entry_static(case_ragged(), prebuilt=True)
    for low in (False, True):
        scene, info, frames = city_frames(low=low)
        prev = None
        for fi, f in enumerate(frames):
            ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], prev)
            e = dict(post=digest(ref['post_cull']), post_count=len(ref['post_cull']), vis64=digest(ref['vis64']),
                     covered=int((ref['vis64'] != 0).sum()), hzb_min=digest(ref['hzb']['min']), hzb_max=digest(ref['hzb']['max']),
                     valid_range=[int(x) for x in ref['hzb']['range']])
            if not low and fi < 2:
                e['gbuffer'] = gbuffer_digest(ref['frame'], ref['vis64'], ref['post_cull'])
            if 'stage0_visible' in ref:
                e.update(stage0_visible=len(ref['stage0_visible']), stage0_rejected=len(ref['stage0_rejected']),
                         stage1_visible=len(ref['stage1_visible']))
            g[f'city_{"low" if low else "street"}_frame{fi}'] = e
            prev = ref['hzb']
    with open(os.path.join(HERE, 'golden_v1.json'), 'w') as fh:
        json.dump(g, fh, indent=1, sort_keys=True)
    print(f'wrote {len(g)} cases')


if __name__ == '__main__':
    main()
