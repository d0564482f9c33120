"""Seeded parity cases shared by the golden-fixture generator, the CPU tests and the GPU tests."""
import hashlib

import numpy as np

from chord_b200 import camera as cam, synth


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def gbuffer_digest(frame, vis64, post_cull):
    """sha256 over the six fp32 G-buffer planes of the ORACLE's resolve (PBR branch + sky branch over the colour plane) with the
    seeded LUTs of chord_b200/synth.py - the golden fixtures' K11 entry."""
    import oracle
    tr, ir = synth.synthetic_luts()
    sc, _ = synth.synthetic_scattering_luts()
    g = oracle.lighting(frame, vis64, post_cull, tr, ir)
    oracle.sky(frame, vis64, tr, ir, sc, None, color=g[0])
    return hashlib.sha256(b''.join(np.ascontiguousarray(p).tobytes() for p in g)).hexdigest()


def case_s1():
    scene, camera, objects = synth.single_meshlet_scene()
    view = cam.fill_view(camera)
    return dict(name='s1', scene=scene, view=view, iview=cam.fill_instance_view(view), objects=objects, w=256, h=256,
                lod0=1, groups=1)


def case_stream(n=3000, w=640, h=360, scale=1.0, seed=5):
    scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=seed, overdraw_scale=scale)
    view = cam.fill_view(camera)
    return dict(name=f'stream_{n}_{w}x{h}_{scale}', scene=scene, view=view, iview=cam.fill_instance_view(view),
                objects=objects, w=w, h=h, lod0=n, groups=(n + 3) // 4, cmds=cmds)


def case_ragged(n=400, w=640, h=360, seed=9):
    """Every meshlet shape the format allows (1..128 triangles, 3..255 vertices, odd data offsets, one- and two-sided)."""
    scene, camera, objects, cmds = synth.ragged_stream_scene(n, w, h, seed=seed)
    view = cam.fill_view(camera)
    return dict(name=f'ragged_{n}', scene=scene, view=view, iview=cam.fill_instance_view(view), objects=objects, w=w, h=h,
                lod0=2 * n, groups=2 * ((n + 3) // 4), cmds=cmds)         # two objects instance the same primitive


def city_frames(n_buildings=9, w=640, h=360, frames=4, seed=2, lods=3, low=False):
    """camera path through the procedural city; yields per-frame inputs with genuine last-frame data."""
    scene, camera, objects, info = synth.city_scene(n_buildings=n_buildings, patches=4, lods=lods, width=w, height=h, seed=seed)
    out, last_view, last_cam = [], None, None
    for fi in range(frames):
        if low:   # skim the ground inside the city: ground triangles cross the near plane, walls fill the view
            pos = np.array([30.0 + 1.5 * fi, 0.35, 60.0 - 2.0 * fi])
            c = cam.Camera(pos, (0.25, -0.05, -1.0), width=w, height=h)
        else:
            c = cam.Camera(camera.position + np.array([0.6 * fi, 0.05 * fi, -1.5 * fi]), camera.front, width=w, height=h)
        objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
        view = cam.fill_view(c, last_view=last_view, frame_counter=fi)
        out.append(dict(view=view, iview=cam.fill_instance_view(view), objects=objs, camera=c))
        last_view, last_cam = view, c
    return scene, info, out


def terrain_case(n=48, w=640, h=360, seed=1, native=False):
    """A curved mesh through the meshlet-DAG builder (chord_b200/nanite_builder.py): genuine multi-level DAG
    with geometric simplification, three cameras from close to far so that the LOD cut moves through the levels."""
    from chord_b200 import nanite_builder as nb
    pos, tris, uv = nb.terrain_mesh(n=n, seed=seed)
    asset = (nb.build_native if native else nb.build)(pos, tris, uv0=uv)
    scene = nb.host_scene(asset)
    frames = []
    for (p, f) in (((2.0, 6.0, 14.0), (0.0, -0.45, -1.0)), ((6.0, 30.0, 70.0), (-0.05, -0.4, -1.0)), ((0.0, 90.0, 260.0), (0.0, -0.33, -1.0))):
        c = cam.Camera(p, f, width=w, height=h)
        view = cam.fill_view(c)
        frames.append(dict(view=view, iview=cam.fill_instance_view(view), objects=np.array([cam.fill_object(np.eye(4), c, 0, 0)]), camera=c))
    return scene, asset, frames


def case_masked(w=512, h=384, seed=11):
    """alpha-tested textured materials (masked raster bucket), chord_b200/synth.py::masked_scene."""
    scene, camera, objects = synth.masked_scene(w, h, seed)
    view = cam.fill_view(camera)
    return dict(name='masked', scene=scene, view=view, iview=cam.fill_instance_view(view), objects=objects, w=w, h=h,
                lod0=len(objects), groups=len(objects), camera=camera)


def oracle_render_shadow(scene, view, objects, lod0, cfg, light, history=None, valid_depth_range=None, hzb_culling=True, extent_scale=1.5):
    """chord::renderShadow (mesh_raster.cpp:332-546) on the CPU oracle, same sequencing as chord_b200/renderer.py::renderShadow.
    ``history`` / result: dict(views, depth=[vis64 per cascade], lists, direction, cfg)."""
    import oracle
    direction = np.asarray(light, dtype=np.float32)
    cache_valid = history is not None and np.array_equal(history['direction'], direction) and history['cfg'] is cfg
    tick = int(view['basicData']['frameCounter'])
    views = cam.cascade_views(view, direction, cfg, valid_depth_range, history['views'] if cache_valid else None, cache_valid)
    dim = cfg.cascadeDim
    depth = list(history['depth']) if cache_valid else [np.zeros((dim, dim), dtype=np.uint64) for _ in range(cfg.cascadeCount)]
    lists = [None] * cfg.cascadeCount
    hzbs = [None] * cfg.cascadeCount
    prev_hzb, prev_id = None, None
    for cid in range(cfg.cascadeCount - 1, -1, -1):
        if cam.is_cascade_cache_valid(cache_valid, cfg.realtimeCascadeCount, cfg.cascadeCount, tick, cid):
            continue
        fr = oracle.Frame(scene, view, views[cid], objects, dim, dim)
        post = fr.instance_culling(lod0)
        if prev_hzb is None:
            if cache_valid and hzb_culling:
                cached = fr.build_hzb(depth[cid], True, False, False)
                post = fr.hzb_culling_generic(cached, views[cid], extent_scale, False, post)
        elif hzb_culling:
            post = fr.hzb_culling_generic(prev_hzb, views[prev_id], extent_scale, False, post)
        depth[cid] = np.zeros((dim, dim), dtype=np.uint64)
        fr.render_mesh_depth(post, depth[cid])
        lists[cid] = post
        if cid != 0:
            prev_hzb, prev_id = fr.build_hzb(depth[cid], True, False, False), cid
            hzbs[cid] = prev_hzb
    return dict(views=views, depth=depth, lists=lists, hzbs=hzbs, direction=direction, cfg=cfg)


def case_tall_frame(w=64, h=4096):
    """a 64 x 4096 frame (HZB mip 0 = 32 x 2048, mips 6.. are ONE texel wide) with a near occluder over the top rows
    and a far, tall meshlet (> 1000 px high, so K3 selects a mip >= 6) lower down with nothing in front of it: it must stay
    visible.  With `pitch = mip0W >> mip` (0 beyond mip 5) every fetch reads texel 0 of the mip - the occluder's depth - and
    the meshlet is wrongly culled; `Texture2D.Load` in the reference reads the right row (hzb_mainview_culling.hlsl:150-185)."""
    import math
    scene, camera, objects, cmds = synth.patch_stream_scene(2, w, h, seed=3)
    d = scene.datas[0]
    pos = d['positions'].reshape(2, 81, 3)
    tan_half = math.tan(math.radians(22.5))
    g = np.linspace(0.0, 1.0, 9, dtype=np.float32)
    gx, gy = np.meshgrid(g, g)
    for m, (depth, y0, y1) in enumerate(((2.0, -40.0, 700.0), (50.0, 2000.0, 3200.0))):
        pxw = 2.0 * tan_half / h * depth
        sx = (-8.0 + gx.reshape(-1) * (w + 16.0)).astype(np.float32)
        sy = (y0 + gy.reshape(-1) * (y1 - y0)).astype(np.float32)
        pos[m, :, 0] = (sx - w / 2.0) * pxw
        pos[m, :, 1] = -(sy - h / 2.0) * pxw
        pos[m, :, 2] = -depth
        d['meshlets']['posMin'][m] = pos[m].min(axis=0)
        d['meshlets']['posMax'][m] = pos[m].max(axis=0)
        d['meshlets']['coneApex'][m] = pos[m].mean(axis=0)
    scene.primitives['posMin'][0] = pos.reshape(-1, 3).min(axis=0)
    scene.primitives['posMax'][0] = pos.reshape(-1, 3).max(axis=0)
    view = cam.fill_view(camera)
    return dict(name='tall_frame', scene=scene, view=view, iview=cam.fill_instance_view(view), objects=objects, w=w, h=h,
                lod0=2, groups=1, cmds=cmds)


def case_textured(w=512, h=384, seed=12):
    """SURVEY 8 row a11 widened: textured PBR materials in the resolve (chord_b200/synth.py::textured_scene)."""
    scene, camera, objects = synth.textured_scene(w, h, seed)
    view = cam.fill_view(camera)
    return dict(name='textured', scene=scene, view=view, iview=cam.fill_instance_view(view), objects=objects, w=w, h=h,
                lod0=len(objects), groups=len(objects), camera=camera)
