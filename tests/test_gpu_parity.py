"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact (integer / byte / index work and
depth bits).  Sizes are chosen so that the oracle finishes in seconds."""
import numpy as np
import pytest

import oracle
from chord_b200 import abi, camera as cam, renderer, runtime, synth

pytestmark = pytest.mark.gpu


def _assert_lists_equal(a, b, what):
    assert len(a) == len(b), f'{what}: count {len(a)} != {len(b)}'
    for k in ('objectId', 'meshletId', 'slot'):
        assert np.array_equal(a[k], b[k]), f'{what}: field {k} differs'


def _assert_vis_equal(gpu, ref, what):
    if np.array_equal(gpu, ref):
        return
    bad = np.argwhere(gpu != ref)
    y, x = bad[0]
    raise AssertionError(f'{what}: {len(bad)} texels differ, first at ({x},{y}): gpu={int(gpu[y, x]):#018x} ref={int(ref[y, x]):#018x}')


def _assert_hzb_equal(h_gpu, h_ref, which='min'):
    for m in range(h_ref['mips']):
        g = h_gpu.readback(which, m).reshape(-1)
        o = int(h_ref['offsets'][m])
        r = h_ref[which][o:o + g.size]
        assert np.array_equal(g, r), f'hzb {which} mip {m}: {int((g != r).sum())} texels differ'


def test_s1_single_meshlet(gpu_ctx):
    scene, camera, objects = synth.single_meshlet_scene()
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    for flags in (0b101, 0b111, 0b000):
        gpu_ctx.set_switch_flags(flags)
        fr = runtime.Frame(gpu_ctx, 256, 256)
        fr.begin(view, iview, objects, 1, 1)
        post = runtime.CmdList(gpu_ctx, 4)
        fr.instance_culling(post)
        ofr = oracle.Frame(scene, view, iview, objects, 256, 256, flags)
        opost = ofr.instance_culling(4)
        _assert_lists_equal(post.readback(), opost, 'post-cull list')
        fr.render_mesh(post)
        ovis, ostats = ofr.render_mesh(opost)
        _assert_vis_equal(fr.vis64(), ovis, 'vis64')
        st = fr.raster_stats()
        for k in ('trianglesSubmitted', 'trianglesCulled', 'trianglesClipped', 'meshletsRastered'):
            assert st[k] == ostats[k], (k, st, ostats)
        hz = runtime.Hzb(gpu_ctx, 256, 256)
        fr.build_hzb(hz, True, True, True)
        oh = ofr.build_hzb(ovis, True, True, True)
        assert hz.mip0 == oh['mip0'] and hz.mips == oh['mips']
        _assert_hzb_equal(hz, oh, 'min')
        _assert_hzb_equal(hz, oh, 'max')
        assert np.array_equal(hz.valid_range(), oh['range'])
        # decode + split images
        vis32 = fr.visibility32()
        assert np.array_equal(vis32, (ovis & 0xFFFFFFFF).astype(np.uint32))
        assert np.array_equal(fr.depth32().view(np.uint32), (ovis >> 32).astype(np.uint32))
        assert np.array_equal(fr.decoded(post), oracle.decode(ovis, opost))
        fr.visibility_mark(post)
        assert np.array_equal(fr.tile_marker(), ofr.visibility_mark(ovis, opost))
    gpu_ctx.set_switch_flags(0b111)


def test_context_destroyed_before_its_children(gpu_ctx):
    """Handle lifetime (reference: pool refs kept alive by CommandBuffer::addReferenceResource, command_list.h:23-38): a
    context destroyed FIRST must stay alive - and its children fully usable - until the last child is gone, in any order.
    Round 1 dereferenced a freed context here (the driver's smoke() died with SIGSEGV after printing its result)."""
    make = type(gpu_ctx)                    # the same Context class the fixture uses (product library or host executor)
    scene, camera, objects = synth.single_meshlet_scene()
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    ofr = oracle.Frame(scene, view, iview, objects, 256, 256)
    opost = ofr.instance_culling(4)
    ovis, _ = ofr.render_mesh(opost)
    for order in ((0, 1, 2, 3), (3, 2, 1, 0), (2, 0, 3, 1)):
        ctx = make(gpu_ctx.device)
        ctx.upload_scene(scene)
        assert ctx.refcount == 1
        fr = runtime.Frame(ctx, 256, 256)
        post = runtime.CmdList(ctx, 4)
        hz = runtime.Hzb(ctx, 256, 256)
        cv = runtime.CascadeViews(ctx)
        assert ctx.refcount == 5
        L, raw = runtime.lib(), ctx.h.value
        ctx.close()                                                    # owner's reference dropped; 4 children keep it
        import ctypes as C
        assert L.chord_ctx_refcount(C.c_void_p(raw)) == 4
        # the children still work: a whole frame after the context handle is gone
        fr.begin(view, iview, objects, 1, 1)
        fr.instance_culling(post)
        fr.render_mesh(post)
        fr.build_hzb(hz, True, True, True)
        fr.visibility_mark(post)
        _assert_lists_equal(post.readback(), opost, 'post-cull list after chord_ctx_destroy')
        _assert_vis_equal(fr.vis64(), ovis, 'vis64 after chord_ctx_destroy')
        children = [fr, post, hz, cv]
        for k, i in enumerate(order):
            children[i].close()
            if k < 3:
                assert L.chord_ctx_refcount(C.c_void_p(raw)) == 3 - k
        # the last close tore the context down; nothing left to touch


@pytest.mark.parametrize('n,w,h,scale', [(3000, 640, 360, 1.0), (1500, 512, 512, 40.0), (600, 256, 256, 2000.0)])
def test_patch_stream_raster(gpu_ctx, n, w, h, scale):
    """raster-only (S5 shape): tiny, medium and large triangles."""
    scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=5, overdraw_scale=scale)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(view, iview, objects, n, (n + 3) // 4)
    lst = runtime.CmdList(gpu_ctx, n)
    lst.upload(cmds)
    fr.render_mesh(lst)
    ofr = oracle.Frame(scene, view, iview, objects, w, h)
    ovis, ostats = ofr.render_mesh(cmds)
    _assert_vis_equal(fr.vis64(), ovis, f'vis64 n={n} scale={scale}')
    st = fr.raster_stats()
    for k in ('trianglesSubmitted', 'trianglesCulled', 'trianglesClipped', 'meshletsRastered'):
        assert st[k] == ostats[k], (k, st, ostats)
    assert st['largeQueueOverflow'] == 0
    # the full instance-cull path must reproduce the prebuilt list for this scene (all LOD-0 roots, camera facing)
    post = runtime.CmdList(gpu_ctx, n)
    fr.instance_culling(post)
    _assert_lists_equal(post.readback(), ofr.instance_culling(n), 'post-cull list')


def test_vertex_record_selftest(gpu_ctx):
    """k_raster_fast shares ONE reciprocal between x / w, y / w and z / w (the fast path of the IEEE division with a range
    guard); its packed per-vertex record must equal the plain three-division statement bit for bit on
    arbitrary inputs: 2^24 pseudo-random clip-space vertices per configuration incl. NaN / infinities / denormals, extreme
    exponent gaps and vertices exactly on pixel boundaries."""
    import ctypes as C
    L = runtime.lib()
    n = 1 << 24
    for seed, (w, h, clamp) in enumerate(((3840, 2160, 0), (640, 360, 0), (2048, 2048, 1), (4096, 64, 0))):
        out = (C.c_uint64 * 3)()
        L.chord_selftest_vertex_records(gpu_ctx.h, C.c_uint32(n), C.c_uint32(1000 + seed), C.c_uint32(w), C.c_uint32(h), C.c_uint32(clamp), out)
        assert out[0] == 0, f'{out[0]} vertex records differ at {w}x{h} clamp={clamp}, first at index {out[1]}'
        assert out[2] > n // 8, 'the sample must contain plenty of vertices inside the guard band'


def test_ragged_meshlets(gpu_ctx):
    """Edge cases of the meshlet format: 1..128 triangles, 3..255 vertices (V > 128 takes the large-vertex instantiation of the
    fast raster kernel through the second work list), odd data offsets, one- and two-sided materials in one list."""
    from cases import case_ragged
    c = case_ragged()
    V = c['scene'].datas[0]['meshlets']['vertexTriangleCount'] & 0xFF
    assert (V > 128).sum() > 50 and V.max() == 255 and V.min() == 3
    gpu_ctx.upload_scene(c['scene'])
    ofr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], c['w'], c['h'])
    n = len(c['cmds'])
    for cmds in (c['cmds'], c['cmds'][::-1].copy(), c['cmds'][V[c['cmds']['meshletId']] > 128].copy(), c['cmds'][:1].copy()):
        fr = runtime.Frame(gpu_ctx, c['w'], c['h'])
        fr.begin(c['view'], c['iview'], c['objects'], c['lod0'], c['groups'])
        lst = runtime.CmdList(gpu_ctx, n)
        lst.upload(cmds)
        fr.render_mesh(lst)
        ovis, ostats = ofr.render_mesh(cmds)
        _assert_vis_equal(fr.vis64(), ovis, f'ragged list of {len(cmds)}')
        st = fr.raster_stats()
        for k in ('trianglesSubmitted', 'trianglesCulled', 'trianglesClipped', 'meshletsRastered'):
            assert st[k] == ostats[k], (k, st, ostats)
        assert st['largeQueueOverflow'] == 0


def test_city_frames(gpu_ctx):
    """S2-shaped city: LOD-DAG cut, cone + frustum cull, two-phase HZB with genuine last-frame history, near-plane
    clipping, large triangles; 4 frames of camera motion through the reference frame sequence."""
    w, h = 640, 360
    scene, camera, objects, info = synth.city_scene(n_buildings=9, patches=4, lods=3, width=w, height=h, seed=2)
    gpu_ctx.upload_scene(scene)
    gpu_ctx.set_switch_flags(0b111)
    r = renderer.DeferredRenderer(gpu_ctx, w, h)
    prev_hzb, last_view, last_cam = None, None, None
    for fi in range(4):
        c = cam.Camera(camera.position + np.array([0.6 * fi, 0.05 * fi, -1.5 * fi]), camera.front, width=w, height=h)
        objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
        view = cam.fill_view(c, last_view=last_view, frame_counter=fi)
        iview = cam.fill_instance_view(view)
        post = r.render(c, objs, info['lod0MeshletCount'], info['meshletGroupCount'], view=view, iview=iview)
        ref = oracle.run_frame(scene, view, iview, objs, w, h, info['lod0MeshletCount'], prev_hzb)
        _assert_lists_equal(post.readback(), ref['post_cull'], f'frame {fi} post-cull')
        _assert_vis_equal(r.frame.vis64(), ref['vis64'], f'frame {fi} vis64')
        _assert_hzb_equal(r.history_hzb, ref['hzb'], 'min')
        _assert_hzb_equal(r.history_hzb, ref['hzb'], 'max')
        assert np.array_equal(r.history_hzb.valid_range(), ref['hzb']['range'])
        if fi > 0:
            assert 'stage0_visible' in ref and len(ref['stage0_rejected']) > 0, 'scene should exercise occlusion'
        assert len(ref['post_cull']) > 0
        prev_hzb, last_view, last_cam = ref['hzb'], view, c


def test_hzb_culling_lists(gpu_ctx):
    """K3 in isolation: both phases, visible and rejected lists in canonical order."""
    w, h = 512, 288
    scene, camera, objects, info = synth.city_scene(n_buildings=6, patches=4, lods=2, width=w, height=h, seed=11)
    gpu_ctx.upload_scene(scene)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    cap = info['lod0MeshletCount']
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(view, iview, objects, cap, info['meshletGroupCount'])
    post = runtime.CmdList(gpu_ctx, cap)
    fr.instance_culling(post)
    fr.render_mesh(post)
    hz = runtime.Hzb(gpu_ctx, w, h)
    fr.build_hzb(hz, True, False, False)
    ofr = oracle.Frame(scene, view, iview, objects, w, h)
    opost = ofr.instance_culling(cap)
    ovis, _ = ofr.render_mesh(opost)
    oh = ofr.build_hzb(ovis, True, False, False)
    _assert_hzb_equal(hz, oh, 'min')
    for first in (True, False):
        vis, rej = runtime.CmdList(gpu_ctx, cap), runtime.CmdList(gpu_ctx, cap)
        fr.hzb_culling(hz, first, post, vis, rej if first else None)
        ov, orj = ofr.hzb_culling(oh, first, opost)
        _assert_lists_equal(vis.readback(), ov, f'visible first={first}')
        if first:
            _assert_lists_equal(rej.readback(), orj, 'rejected')
            assert len(orj) > 0 and len(ov) > 0


def test_tile_scissor(gpu_ctx):
    """screen-tile partition: union of tile renders == full render."""
    w, h = 512, 256
    scene, camera, objects, cmds = synth.patch_stream_scene(800, w, h, seed=9, overdraw_scale=30.0)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    lst = runtime.CmdList(gpu_ctx, len(cmds))
    lst.upload(cmds)
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(view, iview, objects, len(cmds), (len(cmds) + 3) // 4)
    fr.render_mesh(lst)
    full = fr.vis64()
    acc = np.zeros_like(full)
    for tile in ((0, 0, 256, 128), (256, 0, 512, 128), (0, 128, 256, 256), (256, 128, 512, 256)):
        fr.begin(view, iview, objects, len(cmds), (len(cmds) + 3) // 4, tile=tile)
        fr.render_mesh(lst)
        part = fr.vis64()
        x0, y0, x1, y1 = tile
        mask = np.zeros_like(full, dtype=bool)
        mask[y0:y1, x0:x1] = True
        assert not part[~mask].any(), 'wrote outside the tile'
        acc |= part
        ot, _ = oracle.Frame(scene, view, iview, objects, w, h, tile=tile).render_mesh(cmds)
        _assert_vis_equal(part, ot, f'tile {tile}')
        # tile classification of a tiled frame: 8x8 tiles outside the rect are marked 'no geometry' without being read
        fr.visibility_mark(lst)
        assert np.array_equal(fr.tile_marker(), oracle.Frame(scene, view, iview, objects, w, h, tile=tile).visibility_mark(ot, cmds)), f'marker of tile {tile}'
    _assert_vis_equal(acc, full, 'union of tiles')


def test_lighting_resolve(gpu_ctx):
    """K9-K11: tile classification + PBR resolve. Floating point: |gpu - oracle| <= 1e-5 (north_star tolerance) on the
    pre-quantisation fp32 G-buffer planes; untouched pixels must stay cleared."""
    from cases import city_frames
    w, h = 640, 360
    scene, info, frames = city_frames(frames=2, w=w, h=h)
    gpu_ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    gpu_ctx.upload_luts(tr, ir)
    r = renderer.DeferredRenderer(gpu_ctx, w, h, do_resolve=True)
    prev = None
    for f in frames:
        post = r.render(f['camera'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'], view=f['view'], iview=f['iview'])
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
        _assert_vis_equal(r.frame.vis64(), ref['vis64'], 'vis64')
        assert np.array_equal(r.frame.tile_marker(), ref['frame'].visibility_mark(ref['vis64'], ref['post_cull']))
        g = oracle.lighting(ref['frame'], ref['vis64'], ref['post_cull'], tr, ir)
        _assert_gbuffer_equal(r.frame, g)
        assert g[0].max() > 0.05 and np.abs(g[4]).max() > 0 or prev is None
        prev = ref['hzb']


GBUF_NAMES = ['color', 'baseColor', 'vertexNormal', 'pixelNormal', 'motion', 'aoRoughMetal']


def _assert_gbuffer_equal(frame, g):
    """north_star: "shaded RGBA within 1e-5".  The resolve kernel restates the oracle operation for operation (IEEE division /
    square root, the oracle's fma placement), so the contract is checked at its strictest: every value of
    every pre-quantisation fp32 plane EQUAL to the oracle's (max |difference| 0 <= 1e-5, also for HDR values whose fp32 ulp is
    larger than 1e-5)."""
    for i, name in enumerate(GBUF_NAMES):
        got = frame.gbuffer(i)
        if not np.array_equal(got, g[i], equal_nan=True):
            err = np.abs(got - g[i])
            bad = np.argwhere(~((got == g[i]) | (np.isnan(got) & np.isnan(g[i]))))
            raise AssertionError(f'{name}: {len(bad)} values differ from the oracle, max |difference| {np.nanmax(err)} first at {tuple(bad[0])}: '
                                 f'{got[tuple(bad[0])]!r} vs {g[i][tuple(bad[0])]!r}')


def test_lighting_textured_materials(gpu_ctx):
    """Row a11 widened (material.hlsli:30-154): sRGB base colour maps, tangent-space normal map (tangent frame of
    nanite_shared.hlsli:160-176), occlusion / roughness / metallic map, all through SampleGrad with the analytic uv
    derivatives of base.hlsli:457-498 and the sampler pinned in DESIGN.md section 3 - every G-buffer plane equal to the oracle."""
    from cases import case_textured
    c = case_textured()
    w, h = c['w'], c['h']
    gpu_ctx.upload_scene(c['scene'])
    tr, ir = synth.synthetic_luts()
    gpu_ctx.upload_luts(tr, ir)
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(c['view'], c['iview'], c['objects'], c['lod0'], c['groups'])
    post = runtime.CmdList(gpu_ctx, c['lod0'])
    fr.instance_culling(post)
    fr.render_mesh(post)
    fr.visibility_mark(post)
    fr.lighting(post)
    ofr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], w, h)
    opost = ofr.instance_culling(c['lod0'])
    ovis, _ = ofr.render_mesh(opost)
    _assert_lists_equal(post.readback(), opost, 'post-cull list')
    _assert_vis_equal(fr.vis64(), ovis, 'vis64')
    g = oracle.lighting(ofr, ovis, opost, tr, ir)
    _assert_gbuffer_equal(fr, g)
    # the scene must actually exercise the textured paths: normal-mapped pixels differ from their vertex normals, the
    # occlusion / roughness / metallic planes carry texture values, base colour varies inside objects
    assert (np.abs(g[3] - g[2]).max(axis=-1) > 0.05).sum() > 500
    assert len(np.unique(g[5][..., 1])) > 50 and len(np.unique(g[5][..., 0])) > 20
    assert len(np.unique(g[1][..., 0])) > 500
    gpu_ctx.upload_textures(None)


def _sky_frame_on_s1(ctx, view, tr, ir, sc, mie, packed=False, w=256, h=256):
    scene, camera, objects = synth.single_meshlet_scene(w, h)
    iview = cam.fill_instance_view(view)
    ctx.upload_scene(scene)
    fr = runtime.Frame(ctx, w, h)
    if packed:
        fr.set_gbuffer_packed(True)
    fr.begin(view, iview, objects, 1, 1)
    post = runtime.CmdList(ctx, 4)
    fr.instance_culling(post)
    fr.render_mesh(post)
    fr.visibility_mark(post)
    fr.lighting(post)
    ofr = oracle.Frame(scene, view, iview, objects, w, h)
    opost = ofr.instance_culling(4)
    ovis, _ = ofr.render_mesh(opost)
    _assert_vis_equal(fr.vis64(), ovis, 'vis64')
    g = oracle.lighting(ofr, ovis, opost, tr, ir)
    pbr = g[0].copy()
    if sc is not None:
        oracle.sky(ofr, ovis, tr, ir, sc, mie, color=g[0])
    return fr, g, pbr, ovis


def test_sky_branch(gpu_ctx):
    """Row a11, LIGHTING_TYPE == None (lighting.hlsl:75-166; the {None, PBR} loop of lighting.cpp:193-246): with the scattering
    volumes uploaded, chord_lighting writes skyColorEvaluate into every texel without geometry - GetSkyRadiance through the
    4-D scattering table packed in a 3-D volume, the ground branch (GetSunAndSkyIrradiance on the intersection with the
    planet) below the horizon - and nothing but the colour plane there.  Every plane equal to the oracle's, for the combined
    table (single Mie extrapolated from its alpha) and for a separate single-Mie volume; packed colour RGBA16F equal too."""
    from cases import city_frames
    tr, ir = synth.synthetic_luts()
    sc, mie = synth.synthetic_scattering_luts()
    gpu_ctx.upload_luts(tr, ir)
    scene, camera, objects = synth.single_meshlet_scene()
    try:
        for combine, m in ((1, None), (0, mie), (0, None)):
            view = cam.fill_view(camera)
            view['basicData']['atmosphere']['bCombineScattering'] = combine
            gpu_ctx.upload_scattering(sc, m)
            fr, g, pbr, ovis = _sky_frame_on_s1(gpu_ctx, view, tr, ir, sc, m)
            _assert_gbuffer_equal(fr, g)
            empty = (ovis & 0xFFFFFFFF) == 0
            # the case must exercise both halves: rays above the horizon (scattering only) and rays that hit the planet
            assert empty.sum() > 10000 and (g[0][empty][:, :3].sum(axis=-1) > 0).all()
            assert np.array_equal(g[0][~empty], pbr[~empty])                   # the sky branch leaves shaded texels alone
            if combine == 1:
                black = view.copy()
                black['basicData']['atmosphere']['ground_albedo'] = 0.0
                iv = cam.fill_instance_view(black)
                o2 = oracle.Frame(scene, black, iv, objects, 256, 256)
                no_ground = oracle.sky(o2, ovis, tr, ir, sc, m)
                hits = (no_ground != g[0]).any(axis=-1) & empty
                assert hits.sum() > 5000 and (empty & ~hits).sum() > 5000, (hits.sum(), empty.sum())
            for k in range(1, 6):
                assert not g[k][empty].any()
        # viewer in space (GetSkyRadiance moves it to the top of the atmosphere along the ray; rays that miss the atmosphere
        # return black) and viewer just under the top boundary
        from chord_b200 import abi
        for dist_km in (6500.0, 6419.0):
            view = cam.fill_view(camera)
            cam_km = np.array([0.0, dist_km - 6360.0, 0.0])
            view['cameraPositionWS_km'] = abi.double3_to_storage(cam_km)
            view['cameraToEarthCenter_km'] = abi.double3_to_storage(cam_km - np.array([0.0, -6360.0, 0.0]))
            gpu_ctx.upload_scattering(sc, None)
            fr, g, _, ovis = _sky_frame_on_s1(gpu_ctx, view, tr, ir, sc, None)
            _assert_gbuffer_equal(fr, g)
            empty = (ovis & 0xFFFFFFFF) == 0
            if dist_km > 6420.0:
                black = empty & ~g[0][..., :3].any(axis=-1)
                assert black.sum() > 500 and (empty & ~black).sum() > 500, (black.sum(), empty.sum())
        # luminance mode (skySpectralRadianceToLumiance, the 1e-5 exposure factor) and the packed colour format
        view = cam.fill_view(camera)
        view['basicData']['atmosphere']['luminanceMode'] = 1
        gpu_ctx.upload_scattering(sc, None)
        fr, g, _, _ = _sky_frame_on_s1(gpu_ctx, view, tr, ir, sc, None)
        _assert_gbuffer_equal(fr, g)
        fr, g, _, _ = _sky_frame_on_s1(gpu_ctx, cam.fill_view(camera), tr, ir, sc, None, packed=True)
        assert np.array_equal(g[0].astype(np.float16).view(np.uint16), fr.gbuffer_packed(0).view(np.uint16))
        # a frame with real geometry through the renderer: sky between the buildings
        w, h = 640, 360
        scene2, info, frames = city_frames(frames=2, w=w, h=h)
        gpu_ctx.upload_scene(scene2)
        r = renderer.DeferredRenderer(gpu_ctx, w, h, do_resolve=True)
        prev = None
        for f in frames:
            r.render(f['camera'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'], view=f['view'], iview=f['iview'])
            ref = oracle.run_frame(scene2, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
            g = oracle.lighting(ref['frame'], ref['vis64'], ref['post_cull'], tr, ir)
            oracle.sky(ref['frame'], ref['vis64'], tr, ir, sc, None, color=g[0])
            _assert_gbuffer_equal(r.frame, g)
            prev = ref['hzb']
        # the branch goes away with its volumes: unshaded texels read as cleared again
        gpu_ctx.upload_scattering(None)
        fr, g, _, _ = _sky_frame_on_s1(gpu_ctx, cam.fill_view(camera), tr, ir, None, None)
        _assert_gbuffer_equal(fr, g)
    finally:
        gpu_ctx.upload_scattering(None)


def test_bvh_traversal(gpu_ctx):
    """N4: persistent-warp BVH traversal -> same post-cull list as the flat kernel and as the oracle; same candidate set."""
    from cases import case_stream, city_frames
    w, h = 640, 360
    scene, info, frames = city_frames(frames=3, n_buildings=16, w=w, h=h)
    far = cam.Camera((80.0, 900.0, 5200.0), (0.0, -0.17, -1.0), width=w, height=h)      # sub-pixel sub-trees get pruned
    fview = cam.fill_view(far)
    frames = frames + [dict(view=fview, iview=cam.fill_instance_view(fview), objects=synth.refill_objects(info['transforms'], far), camera=far)]
    gpu_ctx.upload_scene(scene)
    cap, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    fr = runtime.Frame(gpu_ctx, w, h)
    pruned = False
    for f in frames:
        for flags in (0b111, 0b000):
            gpu_ctx.set_switch_flags(flags)
            fr.begin(f['view'], f['iview'], f['objects'], cap, groups)
            a, b = runtime.CmdList(gpu_ctx, cap), runtime.CmdList(gpu_ctx, cap)
            fr.instance_culling(a)
            fr.instance_culling_bvh(b)
            ofr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], w, h, flags)
            ref, cand = ofr.instance_culling_bvh(cap)
            _assert_lists_equal(a.readback(), ref, 'flat')
            _assert_lists_equal(b.readback(), ref, 'bvh')
            bits, ctrl = fr.bvh_candidates(groups)
            obits, n, total = ofr.bvh_candidates(groups)
            assert ctrl[4] == 0 and ctrl[2] == 0 and ctrl[3] == 1, ctrl        # no overflow, nothing pending, finished
            assert np.array_equal(bits, obits)
            pruned |= n < total
    assert pruned
    gpu_ctx.set_switch_flags(0b111)
    # a primitive whose groups are all root-level leaves (S5 shape): one node, every group a candidate
    c = case_stream(3000, 640, 360, 1.0)
    gpu_ctx.upload_scene(c['scene'])
    fr.begin(c['view'], c['iview'], c['objects'], 3000, 750)
    b = runtime.CmdList(gpu_ctx, 3000)
    fr.instance_culling_bvh(b)
    _assert_lists_equal(b.readback(), oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 640, 360).instance_culling(3000), 'stream bvh')


def test_full_size_s5_properties(gpu_ctx):
    """BASELINE configs[4] at FULL size (781 250 meshlets x 128 tri, 3840x2160), size-independent properties:
    run-to-run determinism (unordered work lists + atomics must not leak into the result), idempotence, union of screen
    tiles == full frame, and bit-exact parity with the oracle on the head of the stream rendered alone."""
    import torch
    from chord_b200 import synth_device
    M, W, H = 781250, 3840, 2160
    scene, camera, objects, cmds = synth_device.patch_stream_scene_device(M, W, H, seed=5, device='cuda:0')
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    fr = runtime.Frame(gpu_ctx, W, H)
    lst = runtime.CmdList(gpu_ctx, M)
    lst.upload(cmds)
    groups = (M + 3) // 4
    fr.begin(view, iview, objects, M, groups)
    fr.render_mesh(lst)
    a = fr.vis64()
    st = fr.raster_stats()
    assert st['trianglesSubmitted'] == M * 128 and st['largeQueueOverflow'] == 0
    fr.render_mesh(lst)                                   # idempotent: a second pass over the same list changes nothing
    assert np.array_equal(fr.vis64(), a)
    fr.begin(view, iview, objects, M, groups)
    fr.render_mesh(lst)
    assert np.array_equal(fr.vis64(), a), 'two runs of the same frame differ'
    acc = np.zeros_like(a)
    for tile in ((0, 0, 1920, 1088), (1920, 0, 3840, 1088), (0, 1088, 1920, 2160), (1920, 1088, 3840, 2160)):
        fr.begin(view, iview, objects, M, groups, tile=tile)
        fr.render_mesh(lst)
        acc |= fr.vis64()
    assert np.array_equal(acc, a), 'union of 4 screen tiles differs from the full frame'
    # the full instance-cull path reproduces the prebuilt list (all groups are LOD-0 roots facing the camera)
    post = runtime.CmdList(gpu_ctx, M)
    fr.begin(view, iview, objects, M, groups)
    fr.instance_culling(post)
    got = post.readback()
    assert len(got) == M and np.array_equal(got['meshletId'], np.arange(M, dtype=np.uint32)) and np.array_equal(got['slot'], got['meshletId'])
    # oracle parity on the first 60 000 meshlets (7.7 M triangles) at full resolution
    n = 60000
    hs = scene.host_prefix(n)
    head = runtime.CmdList(gpu_ctx, n)
    head.upload(cmds[:n])
    fr.begin(view, iview, objects, M, groups)
    fr.reset_stats()                                      # the counters accumulate over calls: count the head alone
    fr.render_mesh(head)
    ovis, ostats = oracle.Frame(hs, view, iview, objects, W, H).render_mesh(cmds[:n])
    _assert_vis_equal(fr.vis64(), ovis, 'head of the S5 stream at 4K')
    st = fr.raster_stats()
    for k in ('trianglesSubmitted', 'trianglesCulled', 'trianglesClipped', 'meshletsRastered'):
        assert st[k] == ostats[k], (k, st, ostats)
    del scene
    torch.cuda.empty_cache()


def test_city_4k_frame(gpu_ctx):
    """BASELINE configs[2] shape at FULL resolution: 371-building city (3.8 M LOD-0 triangles), 3840x2160, two frames
    through the whole pass graph (two-phase HZB with history, HZB builds, tile marker) bit-exact against the oracle."""
    w, h = 3840, 2160
    scene, camera, objects, info = synth.city_scene(n_buildings=371, patches=4, lods=3, width=w, height=h, seed=3, n_protos=48)
    gpu_ctx.upload_scene(scene)
    r = renderer.DeferredRenderer(gpu_ctx, w, h)
    prev, last_view, last_cam = None, None, None
    for fi in range(2):
        c = cam.Camera(camera.position + np.array([0.15 * fi, 0.01 * fi, -0.6 * fi]), camera.front, width=w, height=h)
        objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
        view = cam.fill_view(c, last_view=last_view, frame_counter=fi)
        iview = cam.fill_instance_view(view)
        post = r.render(c, objs, info['lod0MeshletCount'], info['meshletGroupCount'], view=view, iview=iview)
        ref = oracle.run_frame(scene, view, iview, objs, w, h, info['lod0MeshletCount'], prev)
        _assert_lists_equal(post.readback(), ref['post_cull'], f'4K frame {fi} post-cull')
        _assert_vis_equal(r.frame.vis64(), ref['vis64'], f'4K frame {fi} vis64')
        _assert_hzb_equal(r.history_hzb, ref['hzb'], 'min')
        _assert_hzb_equal(r.history_hzb, ref['hzb'], 'max')
        assert r.frame.raster_stats()['largeQueueOverflow'] == 0
        prev, last_view, last_cam = ref['hzb'], view, c


def test_lighting_packed_formats(gpu_ctx):
    """The reference's storage formats (RGBA16F, A2B10G10R10, RG16F, RGBA8): GPU packed output vs the oracle's fp32
    planes quantised the same way, at most one code value / one fp16 ulp apart (the fp32 inputs agree to 1e-5)."""
    from cases import city_frames
    w, h = 640, 360
    scene, info, frames = city_frames(frames=2, w=w, h=h)
    gpu_ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    gpu_ctx.upload_luts(tr, ir)
    r = renderer.DeferredRenderer(gpu_ctx, w, h, do_resolve=True, packed_gbuffer=True, use_bvh=True)
    prev = None
    for f in frames:
        r.render(f['camera'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'], view=f['view'], iview=f['iview'])
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
        _assert_vis_equal(r.frame.vis64(), ref['vis64'], 'vis64 (BVH cull path)')
        g = oracle.lighting(ref['frame'], ref['vis64'], ref['post_cull'], tr, ir)
        prev = ref['hzb']

    def half_ulps(a, b):
        a16, b16 = a.astype(np.float16).view(np.int16).astype(np.int32), b.view(np.int16).astype(np.int32)
        return np.abs(a16 - b16).max()
    assert half_ulps(g[0], r.frame.gbuffer_packed(0)) <= 1
    assert half_ulps(g[4][..., :2], r.frame.gbuffer_packed(4)) <= 1
    for which in (1, 2, 3):
        p = r.frame.gbuffer_packed(which)
        want = np.rint(np.clip(g[which][..., :3], 0, 1) * 1023.0).astype(np.int32)
        got = np.stack([p & 0x3FF, (p >> 10) & 0x3FF, (p >> 20) & 0x3FF], axis=-1).astype(np.int32)
        assert np.abs(got - want).max() <= 1, which
        assert np.array_equal((p >> 30) & 3, np.rint(np.clip(g[which][..., 3], 0, 1) * 3).astype(np.uint32))
    p8 = r.frame.gbuffer_packed(5).astype(np.int32)
    assert np.abs(p8 - np.rint(np.clip(g[5], 0, 1) * 255.0).astype(np.int32)).max() <= 1


def test_queue_overflow_fallbacks(gpu_ctx):
    """Overflowing the deferred / large-triangle queues must not change a single texel (redo through the all-in-one kernel,
    in-thread rasterisation); error paths return CHORD_ERR_* instead of crashing."""
    from cases import case_stream, city_frames
    for (n, w, h, scale) in ((1500, 512, 512, 40.0), (600, 256, 256, 2000.0)):
        c = case_stream(n, w, h, scale)
        gpu_ctx.upload_scene(c['scene'])
        ovis, _ = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], w, h).render_mesh(c['cmds'])
        lst = runtime.CmdList(gpu_ctx, n)
        lst.upload(c['cmds'])
        cases = [dict(deferred=64), dict(deferred=1, large_tris=1, large_items=1)]
        if scale > 1000:
            cases.append(dict(large_tris=8, large_items=16))                 # only this stream has large triangles
        for limits in cases:
            fr = runtime.Frame(gpu_ctx, w, h)
            fr.set_queue_limits(**limits)
            fr.begin(c['view'], c['iview'], c['objects'], n, (n + 3) // 4)
            fr.render_mesh(lst)
            _assert_vis_equal(fr.vis64(), ovis, f'overflow {limits} scale {scale}')
            assert fr.raster_stats()['largeQueueOverflow'] > 0, limits
    # city with near-plane clipping + big triangles through the overflow paths
    scene, info, frames = city_frames(frames=1, low=True)
    f = frames[0]
    gpu_ctx.upload_scene(scene)
    fr = runtime.Frame(gpu_ctx, 640, 360)
    fr.set_queue_limits(deferred=16, large_tris=4, large_items=8)
    fr.begin(f['view'], f['iview'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'])
    post = runtime.CmdList(gpu_ctx, info['lod0MeshletCount'])
    fr.instance_culling(post)
    fr.render_mesh(post)
    ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], None)
    _assert_vis_equal(fr.vis64(), ref['vis64'], 'city through overflow paths')
    # error behaviour (the reference check()-crashes): statuses + messages, never a crash
    with pytest.raises(runtime.ChordError, match='capacity'):
        small = runtime.CmdList(gpu_ctx, 4)
        small.upload(np.zeros(8, dtype=abi.DrawCmd))
    with pytest.raises(runtime.ChordError, match='64, 4096'):
        runtime.Frame(gpu_ctx, 32, 32)
    with pytest.raises(runtime.ChordError):
        fr.begin(f['view'], f['iview'], f['objects'][:0], 1, 1)          # shouldRenderGLTF: no objects
    with pytest.raises(runtime.ChordError, match='render dimension'):
        bad = runtime.Hzb(gpu_ctx, 256, 256)
        fr.build_hzb(bad)


def _demo_exe():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, 'examples', 'chord_b200_demo')
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build_cpp_demo()
    return exe


def _write_demo_blob(path, scene, info, frames, w, h):
    """examples/chord_b200_demo.cpp input: header, scene arrays, first frame, then (optionally) a count and more frames."""
    d = scene.datas[0]
    f = frames[0]
    with open(path, 'wb') as fh:
        hdr = [w, h, len(f['objects']), info['lod0MeshletCount'], info['meshletGroupCount'], len(d['meshlets']), len(d['positions']),
               len(d['meshletData']), len(d['bvhNodes']), len(d['groups']), len(d['groupIndices']), len(scene.primitives), len(scene.materials)]
        fh.write(np.array(hdr, dtype=np.uint64).tobytes())

        def put(a):
            b = np.ascontiguousarray(a).tobytes()
            fh.write(b + b'\0' * ((-len(b)) % 8))
        for a in (d['meshlets'], d['positions'], d['meshletData'], d['bvhNodes'], d['groups'], d['groupIndices'], scene.primitives,
                  scene.materials, f['view'], f['iview'], f['objects']):
            put(a)
        if len(frames) > 1:
            fh.write(np.array([len(frames) - 1], dtype=np.uint64).tobytes())
            for g in frames[1:]:
                assert len(g['objects']) == len(f['objects'])
                put(g['view']); put(g['iview']); put(g['objects'])


def _read_demo_frames(path, n_frames, w, h):
    raw, off, out = open(path, 'rb').read(), 0, []
    for _ in range(n_frames):
        n = int(np.frombuffer(raw[off:off + 8], dtype=np.uint64)[0]); off += 8
        cmds = np.frombuffer(raw[off:off + 12 * n], dtype=abi.DrawCmd); off += 12 * n
        vis = np.frombuffer(raw[off:off + 8 * w * h], dtype=np.uint64).reshape(h, w); off += 8 * w * h
        out.append((cmds, vis))
    assert off == len(raw)
    return out


def test_native_cpp_host(gpu_ctx, tmp_path, exe=None):
    """C++ host code -> C++ shim (include/chord_b200.hpp) -> C ABI -> CUDA: the path a chord maintainer would bind.
    The demo is compiled by __graft_entry__.build(); here it renders a city frame natively and must match the oracle."""
    import subprocess
    from cases import city_frames
    exe = exe or _demo_exe()
    w, h = 640, 360
    scene, info, frames = city_frames(frames=1, w=w, h=h)
    f = frames[0]
    blob = tmp_path / 'in.blob'
    _write_demo_blob(blob, scene, info, frames, w, h)
    ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], None)
    for mode in ([], ['bvh']):
        out = tmp_path / f'out{len(mode)}.blob'
        res = subprocess.run([exe, str(blob), str(out)] + mode, capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        (cmds, vis), = _read_demo_frames(out, 1, w, h)
        _assert_lists_equal(cmds, ref['post_cull'], f'native host {mode}')
        _assert_vis_equal(vis, ref['vis64'], f'native host {mode}')


def test_builder_asset_parity(gpu_ctx):
    """A curved mesh through the meshlet-DAG builder: per camera the flat and BVH post-cull lists, the packed
    surface and the HZB chain must equal the oracle bit for bit, with the LOD cut moving through >= 3 levels."""
    from cases import terrain_case
    w, h = 640, 360
    scene, asset, frames = terrain_case(w=w, h=h)
    gpu_ctx.upload_scene(scene)
    cap, groups = asset['lod0MeshletCount'] * 2, asset['meshletGroupCount']
    lods = set()
    for flags in (0b111, 0b000):
        gpu_ctx.set_switch_flags(flags)
        for f in frames:
            fr = runtime.Frame(gpu_ctx, w, h)
            fr.begin(f['view'], f['iview'], f['objects'], cap, groups)
            a, b = runtime.CmdList(gpu_ctx, cap), runtime.CmdList(gpu_ctx, cap)
            fr.instance_culling(a)
            fr.instance_culling_bvh(b)
            ofr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], w, h, flags)
            ref = ofr.instance_culling(cap)
            _assert_lists_equal(a.readback(), ref, 'terrain flat')
            _assert_lists_equal(b.readback(), ref, 'terrain bvh')
            lods |= set(scene.datas[0]['meshlets']['lod'][ref['meshletId']].tolist())
            fr.render_mesh(a)
            ovis, ost = ofr.render_mesh(ref)
            _assert_vis_equal(fr.vis64(), ovis, 'terrain vis64')
            assert fr.raster_stats()['trianglesCulled'] == ost['trianglesCulled']
            hz = runtime.Hzb(gpu_ctx, w, h)
            fr.build_hzb(hz, True, True, True)
            _assert_hzb_equal(hz, ofr.build_hzb(ovis, True, True, True), 'min')
    gpu_ctx.set_switch_flags(0b111)
    assert len(lods) >= 3


def test_masked_bucket(gpu_ctx):
    """alpha-tested textured materials (mesh_raster.hlsl:107-112, :198-204) — per-fragment clip() with the
    sampler pinned in DESIGN.md section 3.  Bit-exact ids + depth against the oracle for the whole frame, for every
    meshlet rendered alone (so that hidden instances are checked too), per screen tile, at two resolutions, and with
    the texture table removed (every masked material falls back to the builtin white constant)."""
    from cases import case_masked
    for (w, h) in ((512, 384), (1920, 1080)):
        c = case_masked(w, h)
        scene = c['scene']
        gpu_ctx.upload_scene(scene)
        n = c['lod0']
        fr = runtime.Frame(gpu_ctx, w, h)
        post = runtime.CmdList(gpu_ctx, n)
        ofr = oracle.Frame(scene, c['view'], c['iview'], c['objects'], w, h)
        opost = ofr.instance_culling(n)
        fr.begin(c['view'], c['iview'], c['objects'], n, n)
        fr.instance_culling(post)
        _assert_lists_equal(post.readback(), opost, 'post-cull list')
        fr.reset_stats()
        fr.render_mesh(post)
        ovis, ostats = ofr.render_mesh(opost)
        _assert_vis_equal(fr.vis64(), ovis, f'masked frame {w}x{h}')
        st = fr.raster_stats()
        for k in ('trianglesSubmitted', 'trianglesCulled', 'trianglesClipped', 'meshletsRastered', 'fragmentsAlphaKilled'):
            assert st[k] == ostats[k], (k, st, ostats)
        assert st['fragmentsAlphaKilled'] > 10000 and st['trianglesClipped'] > 0
        if (w, h) == (512, 384):
            one = runtime.CmdList(gpu_ctx, 1)
            for k in range(len(opost)):
                one.upload(opost[k:k + 1])
                fr.begin(c['view'], c['iview'], c['objects'], n, n)
                fr.render_mesh(one)
                ov, _ = ofr.render_mesh(opost[k:k + 1])
                _assert_vis_equal(fr.vis64(), ov, f'meshlet {k} alone (material {int(c["objects"][opost[k]["objectId"]]["GLTFMaterialData"])})')
            acc = np.zeros((h, w), dtype=np.uint64)
            for tile in ((0, 0, 256, 192), (256, 0, 512, 192), (0, 192, 256, 384), (256, 192, 512, 384)):
                fr.begin(c['view'], c['iview'], c['objects'], n, n, tile=tile)
                fr.render_mesh(post)
                part = fr.vis64()
                ot, _ = oracle.Frame(scene, c['view'], c['iview'], c['objects'], w, h, tile=tile).render_mesh(opost)
                _assert_vis_equal(part, ot, f'masked tile {tile}')
                acc = np.maximum(acc, part)
            _assert_vis_equal(acc, ovis, 'union of masked tiles')
            # no texture table: builtin white everywhere
            gpu_ctx.upload_textures(None)
            saved, scene.textures = scene.textures, None
            try:
                fr.begin(c['view'], c['iview'], c['objects'], n, n)
                fr.reset_stats()
                fr.render_mesh(post)
                ow, ows = oracle.Frame(scene, c['view'], c['iview'], c['objects'], w, h).render_mesh(opost)
                _assert_vis_equal(fr.vis64(), ow, 'masked frame without a texture table')
                assert fr.raster_stats()['fragmentsAlphaKilled'] == 0 == ows['fragmentsAlphaKilled']
            finally:
                scene.textures = saved
    gpu_ctx.upload_textures(None)


def test_graph_replay_matches_eager(gpu_ctx):
    """chord_graph_*: a recorded frame (uniform upload, clear, cull, raster, HZB, tile mark) replayed as one CUDA graph gives
    the same bytes as the eager calls, replay after replay, and picks up nothing from the calls made in between."""
    from cases import case_stream, case_s1
    c = case_stream(3000, 640, 360, 40.0)
    gpu_ctx.upload_scene(c['scene'])
    n, w, h = c['lod0'], c['w'], c['h']
    fr = runtime.Frame(gpu_ctx, w, h)
    post = runtime.CmdList(gpu_ctx, n)
    hz = runtime.Hzb(gpu_ctx, w, h)

    def frame_calls():
        fr.begin(c['view'], c['iview'], c['objects'], n, c['groups'])
        fr.instance_culling(post)
        fr.render_mesh(post)
        fr.build_hzb(hz, True, True, True)
        fr.visibility_mark(post)

    frame_calls()
    ref = dict(vis=fr.vis64(), post=post.readback(), hzb=hz.readback_chain('min'), marker=fr.tile_marker(), hmax=hz.readback_chain('max'),
               rng=hz.valid_range().copy())
    with runtime.Graph.capture(gpu_ctx) as g:
        frame_calls()
    assert g.kernel_nodes >= 8
    # dirty everything the graph writes, then replay
    other = cam.fill_view(cam.Camera((0.3, 0.1, 0.5), (0.1, 0.0, -1.0), width=w, height=h))
    fr.begin(other, cam.fill_instance_view(other), c['objects'], n, c['groups'])
    fr.instance_culling(post)
    fr.render_mesh(post)
    assert not np.array_equal(fr.vis64(), ref['vis'])
    for rep in range(3):
        k0 = gpu_ctx.kernel_launches
        g.launch()
        assert gpu_ctx.kernel_launches - k0 == g.kernel_nodes
        _assert_vis_equal(fr.vis64(), ref['vis'], f'graph replay {rep}')
        _assert_lists_equal(post.readback(), ref['post'], 'post-cull list (graph)')
        assert np.array_equal(hz.readback_chain('min'), ref['hzb']) and np.array_equal(fr.tile_marker(), ref['marker'])
        assert np.array_equal(hz.readback_chain('max'), ref['hmax']) and np.array_equal(hz.valid_range(), ref['rng'])
    g.close()
    # a synchronising call inside a capture is reported as an error, and the context stays usable
    with pytest.raises(runtime.ChordError):
        with runtime.Graph.capture(gpu_ctx) as bad:
            frame_calls()
            fr.vis64()
    frame_calls()
    _assert_vis_equal(fr.vis64(), ref['vis'], 'eager after a failed capture')


def _depth_bits(vis64):
    return (vis64 >> np.uint64(32)).astype(np.uint32)


def test_cascade_shadow_depth(gpu_ctx):
    """renderShadow (mesh_raster.cpp:332-546) — cascade views -> instance cull with the orthographic view ->
    hzbCullingGeneric against the previous cascade (or the cached depth) -> depth-only raster (two-sided, depth clamp)
    -> HZB.  Per cascade: lists identical, D32 depth bit-exact (and the packed ids), HZB bit-exact; three frames so
    that the cascade cache (round-robin far cascades, history HZB path) is exercised; plus the occluding patch stream."""
    from cases import case_stream, city_frames, oracle_render_shadow
    from chord_b200 import camera as cam_
    light = (0.3, -0.8, 0.2)
    scene, info, frames = city_frames(frames=3)
    gpu_ctx.upload_scene(scene)
    cfg = cam_.CascadeShadowMapConfig(cascadeCount=4, realtimeCascadeCount=2, cascadeDim=512, cascadeEndDistance=80.0, farCascadeEndDistance=400.0)
    pool = renderer._Pool(gpu_ctx)
    hist_g, hist_o = None, None
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    for fi, f in enumerate(frames):
        g = renderer.renderShadow(gpu_ctx, pool, f['view'], f['objects'], lod0, groups, hist_g, cfg, light)
        o = oracle_render_shadow(scene, f['view'], f['objects'], lod0, cfg, light, hist_o)
        for cid in range(cfg.cascadeCount):
            assert (g.lists[cid] is None) == (o['lists'][cid] is None), (fi, cid)
            if g.lists[cid] is not None:
                _assert_lists_equal(g.lists[cid].readback(), o['lists'][cid], f'frame {fi} cascade {cid} list')
            _assert_vis_equal(g.depths[cid].vis64(), o['depth'][cid], f'frame {fi} cascade {cid} depth')
            assert np.array_equal(g.depths[cid].depth32().view(np.uint32), _depth_bits(o['depth'][cid]))
        hist_g, hist_o = g, o
    # a valid SDSM range (buildHZB's validRange of the main view) moves the realtime cascades only
    vr = np.array([0.00002, 0.0005], dtype=np.float32).view(np.uint32)
    g = renderer.renderShadow(gpu_ctx, pool, frames[0]['view'], frames[0]['objects'], lod0, groups, None, cfg, light, valid_depth_range=vr)
    o = oracle_render_shadow(scene, frames[0]['view'], frames[0]['objects'], lod0, cfg, light, None, valid_depth_range=vr)
    for cid in range(cfg.cascadeCount):
        _assert_vis_equal(g.depths[cid].vis64(), o['depth'][cid], f'SDSM cascade {cid} depth')
    # occlusion along the light: hzbCullingGeneric really rejects, and the HZB it reads is the previous cascade's
    c = case_stream(3000, 256, 256, 100.0)
    gpu_ctx.upload_scene(c['scene'])
    cfg2 = cam_.CascadeShadowMapConfig(cascadeCount=4, realtimeCascadeCount=2, cascadeDim=256, cascadeEndDistance=30.0, farCascadeEndDistance=120.0)
    sl = (0.05, -0.05, -1.0)
    g = renderer.renderShadow(gpu_ctx, pool, c['view'], c['objects'], c['lod0'], c['groups'], None, cfg2, sl)
    o = oracle_render_shadow(c['scene'], c['view'], c['objects'], c['lod0'], cfg2, sl, None)
    rejected = 0
    for cid in range(4):
        _assert_lists_equal(g.lists[cid].readback(), o['lists'][cid], f'stream cascade {cid} list')
        _assert_vis_equal(g.depths[cid].vis64(), o['depth'][cid], f'stream cascade {cid} depth')
        full = oracle.Frame(c['scene'], c['view'], o['views'][cid], c['objects'], 256, 256).instance_culling(c['lod0'])
        rejected += len(full) - len(o['lists'][cid])
    assert rejected > 300
    # the depth-clamp pass is refused for a perspective instance view
    fr = runtime.Frame(gpu_ctx, 256, 256)
    fr.begin(c['view'], c['iview'], c['objects'], c['lod0'], c['groups'])
    lst = runtime.CmdList(gpu_ctx, c['lod0'])
    lst.upload(c['cmds'])
    with pytest.raises(runtime.ChordError):
        fr.render_mesh_depth(lst, abi.PASS_DEPTH | abi.PASS_DEPTH_CLAMP)
    fr.render_mesh_depth(lst, abi.PASS_DEPTH)
    ovis, _ = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 256, 256).render_mesh_depth(c['cmds'], pass_flags=abi.PASS_DEPTH)
    _assert_vis_equal(fr.vis64(), ovis, 'two-sided depth pass through the perspective view')


def test_overlapped_list_uploads(gpu_ctx):
    """chord_cmdlist_upload_async runs on the context's upload stream: events keep it behind the last consumer of the list
    and ahead of the next one.  Alternate two lists with different contents; the next upload is always in flight while the current list is rasterised."""
    import ctypes as C
    import torch
    from cases import case_stream
    c = case_stream(3000, 640, 360, 1.0)
    gpu_ctx.upload_scene(c['scene'])
    n = c['lod0']
    L = runtime.lib()
    fr = runtime.Frame(gpu_ctx, 640, 360)
    lists = [runtime.CmdList(gpu_ctx, n), runtime.CmdList(gpu_ctx, n)]
    subsets = [c['cmds'][k::4].copy() for k in range(4)] + [c['cmds'][:0].copy()]      # incl. an empty list
    pins = []
    for s in subsets:
        t = torch.empty(max(len(s), 1) * 12, dtype=torch.uint8).pin_memory()
        t.numpy()[:len(s) * 12] = np.frombuffer(s.tobytes(), dtype=np.uint8)
        pins.append(t)
    ofr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 640, 360)
    want = [ofr.render_mesh(s)[0] for s in subsets]
    order = [0, 1, 2, 3, 4, 2, 0, 3, 1, 4, 0]
    imgs = [np.empty((360, 640), dtype=np.uint64) for _ in order]
    L.chord_cmdlist_upload_async(lists[0].h, C.c_void_p(pins[order[0]].data_ptr()), C.c_uint32(len(subsets[order[0]])))
    for i, k in enumerate(order):
        if i + 1 < len(order):
            nk = order[i + 1]
            L.chord_cmdlist_upload_async(lists[(i + 1) & 1].h, C.c_void_p(pins[nk].data_ptr()), C.c_uint32(len(subsets[nk])))
        fr.begin(c['view'], c['iview'], c['objects'], n, c['groups'])
        fr.render_mesh(lists[i & 1])
        fr.readback_vis64_into(imgs[i].ctypes.data)
    for i, k in enumerate(order):
        _assert_vis_equal(imgs[i], want[k], f'step {i} (subset {k})')


def test_cascade_setup_on_device(gpu_ctx):
    """chord_cascade_setup (cascade_setup.hlsl:89-371 as a one-warp kernel): the views it writes agree with the host mirror to
    fp32 tolerance (pow() cannot be bit-pinned), including the SDSM branch fed from buildHZB's valid range on the GPU and
    the round-robin cascade cache; the shadow pass driven by the device views is bit-exact against the oracle given the
    very same views (read back only for the check)."""
    from cases import city_frames, oracle_render_shadow
    from chord_b200 import camera as cam_
    light = (0.3, -0.8, 0.2)
    scene, info, frames = city_frames(frames=2)
    gpu_ctx.upload_scene(scene)
    w, h = 640, 360
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    cfg = cam_.CascadeShadowMapConfig(cascadeCount=4, realtimeCascadeCount=2, cascadeDim=512, cascadeEndDistance=80.0, farCascadeEndDistance=400.0)
    main = renderer.DeferredRenderer(gpu_ctx, w, h)
    pool = renderer._Pool(gpu_ctx)

    def close(a, b, what):
        for k in ('translatedWorldToClip', 'clipToTranslatedWorld', 'orthoDepthConvertToView', 'renderDimension', 'frustumPlanesRS'):
            x, y = np.asarray(a[k], dtype=np.float64), np.asarray(b[k], dtype=np.float64)
            scale = max(1.0, np.abs(y).max())
            assert np.abs(x - y).max() <= 2e-3 * scale, (what, k, np.abs(x - y).max(), scale)
        assert np.array_equal(a['cameraWorldPos'], b['cameraWorldPos'])

    hist = None
    for fi, f in enumerate(frames):
        main.render(f['camera'], f['objects'], lod0, groups, view=f['view'], iview=f['iview'])      # builds the HZB + valid range
        for sdsm in (False, True):
            hz = main.history_hzb if sdsm else None
            vr = main.history_hzb.valid_range() if sdsm else None
            g = renderer.renderShadow(gpu_ctx, pool, f['view'], f['objects'], lod0, groups, hist if not sdsm else None, cfg, light,
                                      main_frame=main.frame, valid_range_hzb=hz)
            dv = g.device_views.readback(cfg.cascadeCount)
            host = cam_.cascade_views(f['view'], light, cfg, valid_depth_range=vr)
            for cid in range(cfg.cascadeCount):
                if g.lists[cid] is not None:
                    close(dv[cid], host[cid], f'frame {fi} sdsm {sdsm} cascade {cid}')
                    m = dv[cid]['translatedWorldToClip'].T
                    assert m[3, 3] == 1.0 and not m[3, :3].any()
            # downstream parity with exactly these views: replay the sequence on the oracle
            rendered = [c for c in range(cfg.cascadeCount) if g.lists[c] is not None]
            prev, prev_id = None, None
            for cid in reversed(rendered):
                ofr = oracle.Frame(scene, f['view'], dv[cid], f['objects'], cfg.cascadeDim, cfg.cascadeDim)
                post = ofr.instance_culling(lod0)
                if prev is not None:
                    post = ofr.hzb_culling_generic(prev, dv[prev_id], 1.5, False, post)
                elif hist is not None and not sdsm:
                    cached = ofr.build_hzb(hist_depth[cid], True, False, False)
                    post = ofr.hzb_culling_generic(cached, dv[cid], 1.5, False, post)
                ov, _ = ofr.render_mesh_depth(post)
                _assert_lists_equal(g.lists[cid].readback(), post, f'frame {fi} sdsm {sdsm} cascade {cid} list')
                _assert_vis_equal(g.depths[cid].vis64(), ov, f'frame {fi} sdsm {sdsm} cascade {cid} depth')
                if cid != 0:
                    prev, prev_id = ofr.build_hzb(ov, True, False, False), cid
            if not sdsm:
                hist = g
                hist_depth = [d.vis64() for d in g.depths]
        if fi == 1:
            # second frame, cache valid: realtime cascades + one far cascade (tick 1 -> cascade 3) were re-rendered
            pass


# ---- cases added after the last GPU run of round 1 (host-executor verified); kept at the end of the file
def test_empty_lists_publish_zero_counts(gpu_ctx):
    """The cull kernels publish their list counts themselves (no memset in front of them): an empty input list and a frame in
    which every object is culled must overwrite a stale count with zero."""
    import ctypes as C
    w, h, n = 256, 256, 200
    scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=5)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(view, iview, objects, n, (n + 3) // 4)
    full = runtime.CmdList(gpu_ctx, n)
    full.upload(cmds)
    fr.render_mesh(full)
    hz = runtime.Hzb(gpu_ctx, w, h)
    fr.build_hzb(hz, True, False, False)
    stale_v, stale_r, empty = runtime.CmdList(gpu_ctx, n), runtime.CmdList(gpu_ctx, n), runtime.CmdList(gpu_ctx, n)
    stale_v.upload(cmds)
    stale_r.upload(cmds)
    empty.upload(cmds[:0])
    L = runtime.lib()
    L.chord_hzb_culling(fr.h, hz.h, C.c_int(1), empty.h, stale_v.h, stale_r.h)
    assert stale_v.count() == 0 and stale_r.count() == 0
    stale_v.upload(cmds)
    L.chord_hzb_culling(fr.h, hz.h, C.c_int(0), empty.h, stale_v.h, None)
    assert stale_v.count() == 0
    away = cam.Camera(camera.position, -camera.front, width=w, height=h)
    v2 = cam.fill_view(away)
    iv2 = cam.fill_instance_view(v2)
    fr.begin(v2, iv2, objects, n, (n + 3) // 4)
    post = runtime.CmdList(gpu_ctx, n)
    post.upload(cmds)
    fr.instance_culling(post)
    assert len(oracle.Frame(scene, v2, iv2, objects, w, h).instance_culling(n)) == 0 and post.count() == 0


def test_readback_stream_modes(gpu_ctx):
    """chord_readback_visibility32_async on the shared transfer stream and on a stream of the frame's own
    (chord_frame_set_readback_stream), full frame and screen tiles (linear and pitched copies): the image that arrives is the
    synchronous read-back's, and switching back and forth completes what is in flight."""
    import ctypes as C
    L = runtime.lib()
    w, h, n = 640, 384, 150
    scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=5, overdraw_scale=12.0)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    lst = runtime.CmdList(gpu_ctx, n)
    lst.upload(cmds)
    for tile in (None, (0, 64, 640, 256), (128, 64, 512, 256)):
        fr = runtime.Frame(gpu_ctx, w, h)
        for own in (0, 1, 0):
            L.chord_frame_set_readback_stream(fr.h, C.c_int(own))
            fr.begin(view, iview, objects, n, (n + 3) // 4, tile)
            fr.render_mesh(lst)
            a = np.zeros(w * h, dtype=np.uint32)
            L.chord_readback_visibility32_async(fr.h, C.c_void_p(a.ctypes.data))
            L.chord_frame_wait_readbacks(fr.h)
            b = fr.visibility32().reshape(h, w)
            x0, y0, x1, y1 = tile or (0, 0, w, h)
            assert np.array_equal(a.reshape(h, w)[y0:y1, x0:x1], b[y0:y1, x0:x1]) and (b[y0:y1, x0:x1] != 0).any()


def test_hzb_chain_combinations(gpu_ctx):
    """buildHZB(bMin, bMax, bValidRange) in every combination hzb.cpp:44-47 allows (each is its own kernel instantiation), on a
    frame that is neither a power of two nor a multiple of 64: interior, edge and off-screen 64x64 blocks."""
    w, h, n = 328, 200, 300
    scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=5, overdraw_scale=20.0)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(view, iview, objects, n, (n + 3) // 4)
    lst = runtime.CmdList(gpu_ctx, n)
    lst.upload(cmds)
    fr.render_mesh(lst)
    ofr = oracle.Frame(scene, view, iview, objects, w, h)
    ovis, _ = ofr.render_mesh(cmds)
    _assert_vis_equal(fr.vis64(), ovis, 'vis64')
    for bmin, bmax, brange in ((True, False, False), (False, True, False), (True, True, False), (True, True, True)):
        hz = runtime.Hzb(gpu_ctx, w, h)
        fr.build_hzb(hz, bmin, bmax, brange)
        oh = ofr.build_hzb(ovis, bmin, bmax, brange)
        if bmin:
            _assert_hzb_equal(hz, oh, 'min')
        if bmax:
            _assert_hzb_equal(hz, oh, 'max')
        if brange:
            assert np.array_equal(hz.valid_range(), oh['range'])


@pytest.mark.parametrize('w,h', [(264, 72), (72, 264), (200, 200)])
def test_tile_marker_odd_tile_counts(gpu_ctx, w, h):
    """K9 classifies two 8x8 tiles per warp: frames whose tile count is odd / whose edges cut tiles (clamp-to-edge gather)."""
    n = 300
    scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=5, overdraw_scale=20.0)
    view = cam.fill_view(camera)
    iview = cam.fill_instance_view(view)
    gpu_ctx.upload_scene(scene)
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(view, iview, objects, n, (n + 3) // 4)
    lst = runtime.CmdList(gpu_ctx, n)
    lst.upload(cmds)
    fr.render_mesh(lst)
    fr.visibility_mark(lst)
    ofr = oracle.Frame(scene, view, iview, objects, w, h)
    ovis, _ = ofr.render_mesh(cmds)
    _assert_vis_equal(fr.vis64(), ovis, f'vis64 {w}x{h}')
    assert np.array_equal(fr.tile_marker(), ofr.visibility_mark(ovis, cmds))


def test_native_cpp_renderer_camera_path(gpu_ctx, tmp_path, exe=None):
    """chord_b200::DeferredRenderer (include/chord_b200.hpp), the C++ mirror of renderer.cpp:311-363 + the history hand-off at
    :489, driven by native host code over a camera path: from the second frame on stage 0 culls against the PREVIOUS frame's
    HZB and stage 1 re-tests the rejects - every frame's post-cull list and packed surface must equal the oracle's, which is
    run with the oracle's own history."""
    import subprocess
    from cases import city_frames
    exe = exe or _demo_exe()
    w, h = 640, 360
    scene, info, frames = city_frames(frames=3, w=w, h=h)
    blob, out = tmp_path / 'path.blob', tmp_path / 'path.out'
    _write_demo_blob(blob, scene, info, frames, w, h)
    res = subprocess.run([exe, str(blob), str(out), 'path'], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    got = _read_demo_frames(out, len(frames), w, h)
    prev = None
    for i, (f, (cmds, vis)) in enumerate(zip(frames, got)):
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
        _assert_lists_equal(cmds, ref['post_cull'], f'C++ renderer frame {i}')
        _assert_vis_equal(vis, ref['vis64'], f'C++ renderer frame {i}')
        prev = ref['hzb']


def _is_host_executor():
    return 'hostcc_scratch' in runtime.LIB_PATH


def test_native_renderer_camera_path(gpu_ctx):
    """chord_renderer (the frame sequence of renderer.cpp:311-363 + the history hand-off at :489 as ONE native call per frame):
    every frame's post-cull list, packed surface and history HZB (min, max, valid range) equal the oracle's, which runs with
    the oracle's own history; the per-pass GPU timers at the reference's seven cut points come back through the API."""
    from cases import city_frames
    w, h = 640, 360
    scene, info, frames = city_frames(frames=4, w=w, h=h)
    gpu_ctx.upload_scene(scene)
    for bvh in (False, True):
        r = runtime.Renderer(gpu_ctx, w, h, bvh=bvh, timers=True)
        prev = None
        for i, f in enumerate(frames):
            r.render(f['view'], f['iview'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'])
            ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
            _assert_lists_equal(r.post_cull.readback(), ref['post_cull'], f'native renderer frame {i}')
            _assert_vis_equal(r.frame.vis64(), ref['vis64'], f'native renderer frame {i}')
            hz = r.history_hzb
            _assert_hzb_equal(hz, ref['hzb'], 'min'); _assert_hzb_equal(hz, ref['hzb'], 'max')
            assert np.array_equal(hz.valid_range(), ref['hzb']['range'])
            t = r.pass_times()
            assert list(t) == runtime.PASS_LABELS and all(v >= 0.0 for v in t.values())
            if i > 0:
                assert 'stage1_visible' in ref             # the camera path exercises both phases
            prev = ref['hzb']
        r.close()


def test_many_objects_instance_culling(gpu_ctx):
    """K1 with more objects than one block of k_object_cull_scan holds (> 128): the group counts of several blocks are scanned
    by the block that finishes last; post-cull lists (flat and BVH) and the surface equal the oracle's over two frames."""
    from cases import city_frames
    w, h = 320, 192
    scene, info, frames = city_frames(n_buildings=170, frames=2, w=w, h=h, lods=2)
    assert len(frames[0]['objects']) > 128
    gpu_ctx.upload_scene(scene)
    for bvh in (False, True):
        r = runtime.Renderer(gpu_ctx, w, h, bvh=bvh)
        prev = None
        for i, f in enumerate(frames):
            r.render(f['view'], f['iview'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'])
            ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
            _assert_lists_equal(r.post_cull.readback(), ref['post_cull'], f'{len(f["objects"])} objects, frame {i}, bvh {bvh}')
            _assert_vis_equal(r.frame.vis64(), ref['vis64'], f'frame {i}')
            prev = ref['hzb']
        r.close()


def test_native_renderer_graph_replay(gpu_ctx):
    """CHORD_RENDERER_GRAPH: from the third frame on the recorded sequence is REPLAYED with the frame's uniforms rewritten
    (chord_graph_update_frame_inputs): lists, packed surface, history HZB and all six resolve planes of every frame must be
    bit-identical to the eager renderer's, and lists / surface to the oracle's."""
    if _is_host_executor():
        pytest.skip('CUDA graphs are not modelled by the host executor')
    from cases import city_frames
    w, h = 640, 360
    scene, info, frames = city_frames(frames=7, w=w, h=h)
    gpu_ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    gpu_ctx.upload_luts(tr, ir)
    eager = runtime.Renderer(gpu_ctx, w, h, resolve=True)
    graph = runtime.Renderer(gpu_ctx, w, h, resolve=True, graph=True, timers=True)
    prev = None
    for i, f in enumerate(frames):
        a = (f['view'], f['iview'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'])
        eager.render(*a); graph.render(*a)
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
        _assert_lists_equal(graph.post_cull.readback(), ref['post_cull'], f'graph frame {i}')
        _assert_vis_equal(graph.frame.vis64(), ref['vis64'], f'graph frame {i}')
        _assert_hzb_equal(graph.history_hzb, ref['hzb'], 'min')
        for k in range(6):
            assert np.array_equal(graph.frame.gbuffer(k).view(np.uint32), eager.frame.gbuffer(k).view(np.uint32)), f'resolve plane {k} of frame {i}: replay differs from eager'
        assert graph.frame.gbuffer(0).max() > 0.01
        t = graph.pass_times()
        assert sum(t.values()) > 0.0
        prev = ref['hzb']
    assert graph.graph_replays >= 3, graph.graph_replays
    eager.close(); graph.close()


@pytest.mark.parametrize('use_graph', [False, True])
def test_native_renderer_async_resolve(gpu_ctx, use_graph):
    """CHORD_RENDERER_ASYNC_RESOLVE: visibilityMark + lighting of frame N on a second stream while the context stream runs the
    visibility passes of frame N + 1 (two frames in flight).  Same results as the serial renderer: per frame when every frame
    is read back, and for the last frame + its history after a run of frames enqueued without touching any result."""
    if use_graph and _is_host_executor():
        pytest.skip('CUDA graphs are not modelled by the host executor')
    from cases import city_frames
    w, h = 640, 360
    host = _is_host_executor()                      # (streams are not modelled there: fewer frames, one pass over them)
    scene, info, frames = city_frames(frames=4 if host else 9, w=w, h=h)
    gpu_ctx.upload_scene(scene)
    tr, ir = synth.synthetic_luts()
    sc, _ = synth.synthetic_scattering_luts()
    gpu_ctx.upload_luts(tr, ir)
    gpu_ctx.upload_scattering(sc)
    try:
        serial = runtime.Renderer(gpu_ctx, w, h, resolve=True)
        a_sync = runtime.Renderer(gpu_ctx, w, h, resolve=True, graph=use_graph, async_resolve=True)
        last = None
        for i, f in enumerate(frames):
            a = (f['view'], f['iview'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'])
            serial.render(*a); a_sync.render(*a)
            _assert_lists_equal(a_sync.post_cull.readback(), serial.post_cull.readback(), f'async frame {i}')
            _assert_vis_equal(a_sync.frame.vis64(), serial.frame.vis64(), f'async frame {i}')
            last = [serial.frame.gbuffer(k) for k in range(6)]
            for k in range(6):
                assert np.array_equal(a_sync.frame.gbuffer(k).view(np.uint32), last[k].view(np.uint32)), f'resolve plane {k} of frame {i}'
            assert np.array_equal(a_sync.frame.tile_marker(), serial.frame.tile_marker())
        assert last[0].max() > 0.01
        if use_graph:
            assert a_sync.graph_replays >= 2, a_sync.graph_replays
        a_sync.close()
        if host:
            serial.close()
            return
        # the same frames enqueued back to back: nothing is read until the end, so resolves and visibility passes overlap
        a_sync = runtime.Renderer(gpu_ctx, w, h, resolve=True, graph=use_graph, async_resolve=True)
        for f in frames:
            a_sync.render(f['view'], f['iview'], f['objects'], info['lod0MeshletCount'], info['meshletGroupCount'])
        _assert_lists_equal(a_sync.post_cull.readback(), serial.post_cull.readback(), 'async run, last frame')
        _assert_vis_equal(a_sync.frame.vis64(), serial.frame.vis64(), 'async run, last frame')
        for k in range(6):
            assert np.array_equal(a_sync.frame.gbuffer(k).view(np.uint32), last[k].view(np.uint32)), f'async run: resolve plane {k} of the last frame'
        for which in ('min', 'max'):
            assert np.array_equal(a_sync.history_hzb.readback_chain(which), serial.history_hzb.readback_chain(which)), f'async run: history {which} chain'
        a_sync.close(); serial.close()
    finally:
        gpu_ctx.upload_scattering(None)


def _tile_case(gpu_ctx, world, w, h, frames_n):
    from cases import city_frames
    from chord_b200 import multigpu
    scene, info, frames = city_frames(n_buildings=12, frames=frames_n, w=w, h=h)
    tiles = multigpu.all_tiles(world, w, h)
    return scene, info, frames, tiles, multigpu


def test_renderer_fused_exchange_in_process(gpu_ctx, world=2, use_graph=False, async_resolve=False, n_frames=5):
    """Tile-parallel frame with the FUSED exchange (k_hzb_build stores each rank's part of the history chain straight into every
    rank's chain, arrival flags, wait + top-mip rebuild in front of the next frame's phase 0): `world` ranks as separate
    contexts in one process (peer pointers are plain device pointers here; across processes they come from CUDA IPC).  Every
    frame: each rank's tile equals the single full frame inside the tile (depth bits + decoded ids), nothing outside it, and -
    after chord_renderer_history_sync - every rank holds the single-GPU min / max chains and valid range bit for bit."""
    w, h = 640, 384
    scene, info, frames, tiles, multigpu = _tile_case(gpu_ctx, world, w, h, n_frames)
    if use_graph and _is_host_executor():
        pytest.skip('CUDA graphs are not modelled by the host executor')
    make = type(gpu_ctx)
    ctxs = [make(gpu_ctx.device) for _ in range(world)]
    for c in ctxs:
        c.upload_scene(scene)
        if async_resolve:
            c.upload_luts(*synth.synthetic_luts())
    gpu_ctx.upload_scene(scene)
    single = runtime.Renderer(gpu_ctx, w, h)
    # async_resolve: the resolve of frame N on a second stream, and the put of its history on a third one, beside frame N + 1
    ranks = [runtime.Renderer(c, w, h, graph=use_graph, resolve=async_resolve, async_resolve=async_resolve) for c in ctxs]
    ptrs = [r.exchange_export()[1] for r in ranks]
    for i, r in enumerate(ranks):
        r.exchange_connect(i, world, tiles, local_ptrs=ptrs)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    for fi, f in enumerate(frames):
        single.render(f['view'], f['iview'], f['objects'], lod0, groups)
        for i, r in enumerate(ranks):                     # lock step: every rank renders frame fi before anyone starts fi + 1
            r.render(f['view'], multigpu.tile_instance_view(f['view'], f['camera'], tiles[i]), f['objects'], lod0, groups)
        vs = single.frame.vis64()
        ds = single.frame.decoded(single.post_cull)
        union = np.zeros_like(vs)
        for i, r in enumerate(ranks):
            x0, y0, x1, y1 = tiles[i]
            vt = r.frame.vis64()
            assert np.array_equal(vt[y0:y1, x0:x1] >> np.uint64(32), vs[y0:y1, x0:x1] >> np.uint64(32)), f'frame {fi} rank {i}: depth differs inside the tile'
            assert np.array_equal(r.frame.decoded(r.post_cull)[y0:y1, x0:x1], ds[y0:y1, x0:x1]), f'frame {fi} rank {i}: decoded ids differ inside the tile'
            outside = vt.copy(); outside[y0:y1, x0:x1] = 0
            assert not outside.any(), f'frame {fi} rank {i}: texels outside the tile'
            union[y0:y1, x0:x1] = vt[y0:y1, x0:x1]
        assert (union != 0).sum() == (vs != 0).sum()
        for i, r in enumerate(ranks):
            r.history_sync()
        hs = single.history_hzb
        for i, r in enumerate(ranks):
            assert not r.exchange_timed_out()
            hr = r.history_hzb
            for which in ('min', 'max'):
                assert np.array_equal(hr.readback_chain(which), hs.readback_chain(which)), f'frame {fi} rank {i}: gathered {which} chain differs from the single-GPU chain'
            assert np.array_equal(hr.valid_range(), hs.valid_range()), f'frame {fi} rank {i}: valid range'
    if use_graph:
        assert all(r.graph_replays >= 1 for r in ranks)
    for r in ranks + [single]:
        r.close()
    for c in ctxs:
        c.close()


def test_exchange_wait_gives_up_on_a_dead_peer(gpu_ctx):
    """Failure detection of the fused exchange: a rank whose peer stops rendering must not hang - the consumer-side wait is
    bounded (chord_renderer_set_exchange_timeout), raises the status word (chord_renderer_exchange_status) and the rank renders
    on with the history it has.  Two ranks in one process; rank 1 "dies" after the first frame."""
    w, h = 640, 384
    scene, info, frames, tiles, multigpu = _tile_case(gpu_ctx, 2, w, h, 4)
    make = type(gpu_ctx)
    ctxs = [make(gpu_ctx.device) for _ in range(2)]
    for c in ctxs:
        c.upload_scene(scene)
    ranks = [runtime.Renderer(c, w, h) for c in ctxs]
    ptrs = [r.exchange_export()[1] for r in ranks]
    for i, r in enumerate(ranks):
        r.exchange_connect(i, 2, tiles, local_ptrs=ptrs)
        r.set_exchange_timeout(20000)                   # 20 000 polls x 200 ns = 4 ms instead of the default 4 s
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']

    def render(i, fi):
        f = frames[fi]
        ranks[i].render(f['view'], multigpu.tile_instance_view(f['view'], f['camera'], tiles[i]), f['objects'], lod0, groups)
    render(0, 0); render(1, 0)
    render(0, 1)                                        # waits for rank 1's frame 0: it is there
    ctxs[0].sync()
    assert not ranks[0].exchange_timed_out()
    render(0, 2)                                        # waits for rank 1's frame 1, which never comes
    ctxs[0].sync()                                      # ... and returns
    assert ranks[0].exchange_timed_out()
    # the rank is still alive: its own tile of the next frame is rendered (occlusion culling against what history it has is
    # conservative or not, the rasteriser resolves visibility either way)
    render(0, 3)
    x0, y0, x1, y1 = tiles[0]
    vt = ranks[0].frame.vis64()
    assert vt[y0:y1, x0:x1].any()
    for r in ranks:
        r.close()
    for c in ctxs:
        c.close()


def test_renderer_fused_exchange_four_ranks_graph(gpu_ctx):
    test_renderer_fused_exchange_in_process(gpu_ctx, world=4, use_graph=True)


@pytest.mark.parametrize('world,use_graph', [(2, False), (4, True)])
def test_renderer_fused_exchange_async(gpu_ctx, world, use_graph):
    """The same with CHORD_RENDERER_ASYNC_RESOLVE: the producer-side put (k_hzb_exchange_put) leaves the visibility passes'
    stream order - it runs on the auxiliary stream with a host-passed sequence number while the rank's own sequence number and
    valid-range entry advance in stream order (k_hzb_exchange_advance); replayed graphs from the fourth frame on."""
    test_renderer_fused_exchange_in_process(gpu_ctx, world=world, use_graph=use_graph, async_resolve=True, n_frames=8 if use_graph else 4)


@pytest.mark.parametrize('world,mode', [(2, 'eager'), (4, 'graph')])
def test_native_cpp_multirank(gpu_ctx, tmp_path, world, mode):
    """examples/chord_b200_multigpu.cpp: native C++ host, one THREAD per rank (ranks map round-robin onto the visible GPUs), each
    rank one screen tile through chord_b200::Renderer with the fused history exchange over peer pointers.  Every frame's stitched
    surface (each rank contributes its tile) must equal the oracle's full frame - which only holds when phase 0 of every rank
    culled against the complete gathered chain of the frame before - and rank 0's gathered min chain must equal the oracle's."""
    if _is_host_executor():
        pytest.skip('the host executor is single-threaded')
    import os
    import subprocess
    import __graft_entry__
    from cases import city_frames
    from chord_b200 import multigpu
    exe = __graft_entry__.build_cpp_demo('chord_b200_multigpu')
    w, h = 640, 384
    scene, info, frames = city_frames(n_buildings=12, frames=5, w=w, h=h)
    tiles = multigpu.all_tiles(world, w, h)
    blob, out = tmp_path / 'ranks.blob', tmp_path / 'ranks.out'
    _write_demo_blob(blob, scene, info, frames, w, h)
    with open(blob, 'ab') as fh:
        if len(frames) == 1:
            fh.write(np.array([0], dtype=np.uint64).tobytes())
        fh.write(np.array([world] + [v for t in tiles for v in t], dtype=np.uint64).tobytes())
        for f in frames:
            for t in tiles:
                b = np.ascontiguousarray(multigpu.tile_instance_view(f['view'], f['camera'], t)).tobytes()
                fh.write(b + b'\0' * ((-len(b)) % 8))
    res = subprocess.run([exe, str(blob), str(out)] + (['graph'] if mode == 'graph' else []), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(out, 'rb').read()
    prev, off = None, 0
    for i, f in enumerate(frames):
        ref = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], w, h, info['lod0MeshletCount'], prev)
        img = np.frombuffer(raw[off:off + 8 * w * h], dtype=np.uint64).reshape(h, w); off += 8 * w * h
        # ids carry each rank's own slot numbering: compare depth bits and coverage (the decoded ids are checked in-process by
        # test_renderer_fused_exchange_in_process)
        assert np.array_equal(img >> np.uint64(32), ref['vis64'] >> np.uint64(32)), f'frame {i}: stitched depth differs from the full frame'
        prev = ref['hzb']
    n = int(np.frombuffer(raw[off:off + 8], dtype=np.uint64)[0]); off += 8
    chain = np.frombuffer(raw[off:off + 2 * n], dtype=np.uint16)
    assert np.array_equal(chain, prev['min']), 'rank 0: gathered min chain differs from the single-GPU chain'


def test_extreme_frame_shapes(gpu_ctx, shapes=((64, 64), (4096, 64), (64, 4096), (4096, 128))):
    """Render dimensions at the limits renderer.h:52-53 allows: the 64 x 64 minimum (six HZB mips, the fewest hzb.cpp:119
    accepts) and the 4096 maximum in one axis with the other tiny (twelve mips, the non-square tail of the chain where one
    dimension has run out): instance cull, raster, HZB min / max / range and the tile marker against the oracle."""
    for (w, h) in shapes:
        n = 120
        scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=7, overdraw_scale=30.0)
        view = cam.fill_view(camera)
        iview = cam.fill_instance_view(view)
        gpu_ctx.upload_scene(scene)
        fr = runtime.Frame(gpu_ctx, w, h)
        fr.begin(view, iview, objects, n, (n + 3) // 4)
        post = runtime.CmdList(gpu_ctx, n)
        fr.instance_culling(post)
        ofr = oracle.Frame(scene, view, iview, objects, w, h)
        opost = ofr.instance_culling(n)
        _assert_lists_equal(post.readback(), opost, f'post-cull list {w}x{h}')
        fr.render_mesh(post)
        ovis, _ = ofr.render_mesh(opost)
        _assert_vis_equal(fr.vis64(), ovis, f'vis64 {w}x{h}')
        assert (ovis != 0).sum() > w * h // 8
        hz = runtime.Hzb(gpu_ctx, w, h)
        fr.build_hzb(hz, True, True, True)
        oh = ofr.build_hzb(ovis, True, True, True)
        assert hz.mips == oh['mips'] == (6 if max(w, h) == 64 else 12)
        _assert_hzb_equal(hz, oh, 'min')
        _assert_hzb_equal(hz, oh, 'max')
        assert np.array_equal(hz.valid_range(), oh['range'])
        fr.visibility_mark(post)
        assert np.array_equal(fr.tile_marker(), ofr.visibility_mark(ovis, opost))


def test_tall_frame_hzb_rows(gpu_ctx):
    """64 x 4096: HZB mips 6.. are one texel wide; K3 must read row `sy` of such a mip (Texture2D.Load semantics), not texel 0.
    A far, 1200 px tall meshlet below a near occluder must survive phase 1 - checked against the oracle AND against a
    mip-independent brute-force test on the depth image itself (visible iff the farthest depth under its rect <= its nearest)."""
    from cases import case_tall_frame
    c = case_tall_frame()
    w, h = c['w'], c['h']
    gpu_ctx.upload_scene(c['scene'])
    fr = runtime.Frame(gpu_ctx, w, h)
    fr.begin(c['view'], c['iview'], c['objects'], c['lod0'], c['groups'])
    occluder, both = runtime.CmdList(gpu_ctx, 2), runtime.CmdList(gpu_ctx, 2)
    occluder.upload(c['cmds'][:1]); both.upload(c['cmds'])
    fr.render_mesh(occluder)
    hz = runtime.Hzb(gpu_ctx, w, h)
    fr.build_hzb(hz, True, False, False)
    vis_l = runtime.CmdList(gpu_ctx, 2)
    fr.hzb_culling(hz, False, both, vis_l)
    ofr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], w, h)
    ovis, _ = ofr.render_mesh(c['cmds'][:1])
    _assert_vis_equal(fr.vis64(), ovis, 'occluder surface')
    oh = ofr.build_hzb(ovis, True, False, False)
    _assert_hzb_equal(hz, oh, 'min')
    ov, _ = ofr.hzb_culling(oh, False, c['cmds'])
    got = vis_l.readback()
    _assert_lists_equal(got, ov, 'phase-1 survivors on the tall frame')
    # brute force on the depth image: rows 2000..3199 hold nothing (depth 0 = far), so the far meshlet cannot be occluded
    depth = (ovis >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert depth[2000:3200].max() == 0.0 and depth[:700].min() > 0.0
    assert 1 in got['meshletId'], 'the tall far meshlet was occlusion-culled although nothing covers it'


def test_random_raster_cases(gpu_ctx, count=48, seed=2024):
    """Seeded random cases (tests/fuzz_cases.py): packed surface bit-equal to the oracle in every case; the cull / clip
    counters too where no scissor tile is set (with a tile, whole meshlets are rejected before they are counted).  The set
    must contain near-plane / guard-band clipping and large triangles."""
    from fuzz_cases import random_raster_case
    rng = np.random.default_rng(seed)
    clipped = large = 0
    for i in range(count):
        c = random_raster_case(rng)
        gpu_ctx.upload_scene(c['scene'])
        fr = runtime.Frame(gpu_ctx, c['w'], c['h'])
        fr.begin(c['view'], c['iview'], c['objects'], c['lod0'], c['groups'], c['tile'])
        lst = runtime.CmdList(gpu_ctx, len(c['cmds']))
        lst.upload(c['cmds'])
        fr.render_mesh(lst)
        ovis, ost = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], c['w'], c['h'], tile=c['tile']).render_mesh(c['cmds'])
        _assert_vis_equal(fr.vis64(), ovis, f'random case {i} {c["what"]}')
        st = fr.raster_stats()
        if c['tile'] is None:
            for k in ('trianglesSubmitted', 'trianglesCulled', 'trianglesClipped', 'meshletsRastered'):
                assert st[k] == ost[k], (k, c['what'], st, ost)
        assert st['largeQueueOverflow'] == 0
        clipped += ost['trianglesClipped']
        large += st['trianglesLarge']
    assert clipped > 0 and large > 0, (clipped, large)


def test_random_camera_frames(gpu_ctx, sequences=10, seed=777):
    """Seeded random camera sequences around the city (3 frames each, so that stage 0 culls against a genuine previous-frame
    HZB), random switch flags, flat or BVH instance cull: post-cull list, packed surface and both HZB chains of every frame
    against the oracle run with its own history."""
    rng = np.random.default_rng(seed)
    w, h = 320, 200
    scene, camera0, _, info = synth.city_scene(n_buildings=9, patches=4, lods=3, width=w, height=h, seed=2)
    gpu_ctx.upload_scene(scene)
    lod0, groups = info['lod0MeshletCount'], info['meshletGroupCount']
    try:
        for si in range(sequences):
            flags = int(rng.choice([0b111, 0b111, 0b101, 0b011, 0b000]))
            gpu_ctx.set_switch_flags(flags)
            r = renderer.DeferredRenderer(gpu_ctx, w, h, use_bvh=bool(rng.integers(0, 2)))
            base = camera0.position + rng.normal(0, 1, 3) * np.array([25.0, 4.0, 25.0])
            base[1] = max(base[1], 0.4)
            front = np.asarray(camera0.front, dtype=np.float64) + rng.normal(0, 0.5, 3)
            prev = last_view = last_cam = None
            for fi in range(3):
                c = cam.Camera(base + rng.normal(0, 1, 3) * np.array([1.5, 0.2, 1.5]) * fi, front + rng.normal(0, 0.05, 3) * fi, width=w, height=h)
                objs = synth.refill_objects(info['transforms'], c, camera_last=last_cam)
                view = cam.fill_view(c, last_view=last_view, frame_counter=fi)
                iview = cam.fill_instance_view(view)
                post = r.render(c, objs, lod0, groups, view=view, iview=iview)
                ref = oracle.run_frame(scene, view, iview, objs, w, h, lod0, prev, flags)
                _assert_lists_equal(post.readback(), ref['post_cull'], f'sequence {si} frame {fi} flags {flags:03b}')
                _assert_vis_equal(r.frame.vis64(), ref['vis64'], f'sequence {si} frame {fi}')
                _assert_hzb_equal(r.history_hzb, ref['hzb'], 'min')
                _assert_hzb_equal(r.history_hzb, ref['hzb'], 'max')
                prev, last_view, last_cam = ref['hzb'], view, c
    finally:
        gpu_ctx.set_switch_flags(0b111)
