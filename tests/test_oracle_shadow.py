"""SURVEY §8f N3 — cascade shadow depth views, CPU leg: the cascade view mirror (cascade_setup.hlsl), the depth-only raster
flags (renderMeshDepth, mesh_raster.cpp:159-206) and hzbCullingGeneric (hzb_culling_generic.hlsl:37-172) of the oracle."""
import numpy as np

import oracle
from cases import city_frames, oracle_render_shadow
from chord_b200 import abi, camera as cam, synth

LIGHT = (0.3, -0.8, 0.2)


def _cfg(**kw):
    d = dict(cascadeCount=4, realtimeCascadeCount=2, cascadeDim=256, cascadeEndDistance=80.0, farCascadeEndDistance=400.0)
    d.update(kw)
    return cam.CascadeShadowMapConfig(**d)


def test_cascade_views():
    scene, info, frames = city_frames(frames=1)
    view = frames[0]['view']
    cfg = _cfg()
    views = cam.cascade_views(view, LIGHT, cfg)
    light = np.asarray(LIGHT, dtype=np.float64) / np.linalg.norm(LIGHT)
    last_r = 0.0
    for cid in range(cfg.cascadeCount):
        m = views[cid]['translatedWorldToClip'].T.astype(np.float64)
        assert m[3, 3] == 1.0 and not m[3, :3].any()                          # orthographic: isOrthoProjection (base.hlsli:241-244)
        # reversed z: moving along the light direction (away from the shadow camera) decreases clip z
        assert (m @ np.append(light, 0.0))[2] < 0
        # texel snapping (cascade_setup.hlsl:283-299): the projected world origin sits on a texel corner
        o = (m @ np.array([0, 0, 0, 1.0]))[:2] * cfg.cascadeDim / 2.0
        assert np.allclose(o, np.rint(o), atol=2e-2), o
        r = 1.0 / m[0, 0] if m[0, 0] else 0.0
        r = 1.0 / np.linalg.norm(m[0, :3])                                      # half extent of the ortho box
        assert r > last_r
        last_r = r
        assert np.allclose(views[cid]['renderDimension'], [256, 256, 1 / 256, 1 / 256])
        inv = views[cid]['clipToTranslatedWorld'].T.astype(np.float64)
        assert np.allclose(inv @ m, np.eye(4), atol=1e-3)
        # the box centre is inside all six planes' positive side or all negative consistently with the main view convention
        c = (inv @ np.array([0, 0, 0.5, 1.0]))[:3]
        d = [np.dot(p[:3], c) + p[3] for p in views[cid]['frustumPlanesRS'][:5]]
        assert all(x > 0 for x in d) or all(x < 0 for x in d), d
    # SDSM: a valid depth range tightens the realtime cascades, the far ones are untouched
    vr = np.array([0.00002, 0.0005], dtype=np.float32).view(np.uint32)
    v2 = cam.cascade_views(view, LIGHT, cfg, valid_depth_range=vr)
    assert not np.array_equal(v2[0]['translatedWorldToClip'], views[0]['translatedWorldToClip'])
    assert np.array_equal(v2[3]['translatedWorldToClip'], views[3]['translatedWorldToClip'])
    # cascade cache (cascade_setup.hlsl:8-22): far cascades update round-robin, realtime ones every frame
    assert [cam.is_cascade_cache_valid(True, 2, 4, t, c) for t in (0, 1) for c in range(4)] == [False, False, False, True, False, False, True, False]
    assert not any(cam.is_cascade_cache_valid(False, 2, 4, 0, c) for c in range(4))


def test_depth_pass_flags():
    """renderMeshDepth: every material two-sided (cull NONE, no mesh-shader cull #0), depth clamp instead of z clipping."""
    scene, camera, objects = synth.single_meshlet_scene()
    view = cam.fill_view(camera)
    cfg = _cfg(cascadeCount=2, realtimeCascadeCount=1, cascadeDim=128, cascadeEndDistance=10.0, farCascadeEndDistance=40.0)
    for light in ((0.0, -0.2, -1.0), (0.0, 0.2, 1.0)):           # lit from the front / from the back of the one-sided patch
        iv = cam.cascade_views(view, light, cfg)[0]
        fr = oracle.Frame(scene, view, iv, objects, 128, 128)
        post = fr.instance_culling(1)
        assert len(post) == 1
        main_like, st0 = fr.render_mesh(post)                     # PASS_TYPE_CLUSTER rules: one-sided
        depth, st1 = fr.render_mesh_depth(post)
        assert (depth != 0).sum() > 50
        if light[2] > 0:                                          # seen from behind: culled unless two-sided
            assert (main_like != 0).sum() == 0 and st0['trianglesCulled'] == 128 and st1['trianglesCulled'] < 128
        else:
            assert np.array_equal(main_like, depth)
    # depth clamp: pull the shadow camera so close that the patch is in front of its near plane (clip z > 1): with the
    # clamp the fragments survive with depth 1.0, without it they are clipped away
    iv = cam.cascade_views(view, (0.0, -0.2, -1.0), cfg)[0].copy()
    m = iv['translatedWorldToClip'].T.copy()
    m[2, 3] += np.float32(0.75)
    iv['translatedWorldToClip'] = m.T
    fr = oracle.Frame(scene, view, iv, objects, 128, 128)
    post = np.zeros(1, dtype=abi.DrawCmd)
    clamped, _ = fr.render_mesh_depth(post, pass_flags=abi.PASS_DEPTH | abi.PASS_DEPTH_CLAMP)
    clipped, _ = fr.render_mesh_depth(post, pass_flags=abi.PASS_DEPTH)
    dz = (clamped >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert (clamped != 0).sum() > 50 and dz.max() == 1.0
    assert (clipped != 0).sum() < (clamped != 0).sum()


STREAM_LIGHT = (0.05, -0.05, -1.0)          # along the view direction: the near patches of the stream occlude the far ones


def _stream_cfg():
    return _cfg(cascadeEndDistance=30.0, farCascadeEndDistance=120.0)


def test_hzb_culling_generic():
    from cases import case_stream
    c = case_stream(3000, 256, 256, 100.0)
    scene, lod0 = c['scene'], c['lod0']
    f = dict(view=c['view'], objects=c['objects'])
    cfg = _stream_cfg()
    views = cam.cascade_views(f['view'], STREAM_LIGHT, cfg)
    far = oracle.Frame(scene, f['view'], views[3], f['objects'], 256, 256)
    post_far = far.instance_culling(lod0)
    depth_far, _ = far.render_mesh_depth(post_far)
    hzb = far.build_hzb(depth_far, True, False, False)
    near = oracle.Frame(scene, f['view'], views[2], f['objects'], 256, 256)
    post = near.instance_culling(lod0)
    kept = near.hzb_culling_generic(hzb, views[3], 1.5, False, post)
    assert 0 < len(kept) < len(post) - 100
    # survivors keep their input order and are a subsequence of the input
    it = iter(post.tolist())
    assert all(any(k == p for p in it) for k in kept.tolist())
    # nothing written yet (HZB all 0 = far): nothing can be occluded; switch flag off: pass-through
    empty = dict(hzb, min=np.zeros_like(hzb['min']))
    assert len(near.hzb_culling_generic(empty, views[3], 1.5, False, post)) == len(post)
    off = oracle.Frame(scene, f['view'], views[2], f['objects'], 256, 256, switch_flags=0b101)
    assert len(off.hzb_culling_generic(hzb, views[3], 1.5, False, post)) == len(post)
    # an all-near HZB (fp16 1.0) occludes every box that is entirely inside the HZB view
    ones = dict(hzb, min=np.full_like(hzb['min'], 0x3C00))
    assert len(near.hzb_culling_generic(ones, views[3], 1.5, False, post)) < len(kept)
    # widening the rect (extentScale) can only keep more
    assert len(near.hzb_culling_generic(hzb, views[3], 8.0, False, post)) >= len(kept)


def test_render_shadow_sequence():
    scene, info, frames = city_frames(frames=3)
    cfg = _cfg()
    hist = None
    covered = []
    for fi, f in enumerate(frames):
        res = oracle_render_shadow(scene, f['view'], f['objects'], info['lod0MeshletCount'], cfg, LIGHT, hist)
        assert res['views'] is not None
        rendered = [c for c in range(4) if res['lists'][c] is not None]
        assert rendered == ([0, 1, 2, 3] if fi == 0 else [0, 1, 2 + (fi % 2)])        # realtime every frame, far round-robin
        if hist is not None:
            for c in range(4):
                if c not in rendered:
                    assert res['depth'][c] is hist['depth'][c] and np.array_equal(res['views'][c], hist['views'][c])
        covered.append([int((d != 0).sum()) for d in res['depth']])
        hist = res
    assert all(c > 0 for c in covered[0][1:])
