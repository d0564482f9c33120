"""Oracle self-consistency: the invariants the reference documents as in-shader check()s and
size-independent properties of the path, plus an independent numpy statement of the HZB."""
import numpy as np

import oracle
from cases import case_s1, case_stream, city_frames


def test_fp16_rtz_conversion():
    lib = oracle.lib()
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.random(4000).astype(np.float32), (rng.random(2000) * 1e-4).astype(np.float32),
                           np.array([0.0, 1.0, 6.1e-5, 5.96e-8, 2.9e-8, 65504.0, 1e-9], dtype=np.float32)])
    for v in vals:
        hb = lib.oracle_f32_to_f16_rtz(float(v))
        back = lib.oracle_f16_to_f32(hb)
        assert back <= v                                             # toward zero (conservative for the min HZB)
        assert np.float32(back) == np.uint16(hb).view(np.float16).astype(np.float32)
        nxt = np.uint16(hb + 1).view(np.float16).astype(np.float32)
        assert nxt > v or not np.isfinite(nxt)                      # and it is the nearest such value


def _numpy_hzb_min(depth, mip0w, mip0h, mips):
    """independent statement: every texel of mip k is RTZ_fp16(min over its 2^(k+1) square footprint, edge clamped)."""
    h, w = depth.shape
    out = []
    for k in range(mips):
        s = 2 << k
        mw, mh = max(mip0w >> k, 1), max(mip0h >> k, 1)
        m = np.empty((mh, mw), dtype=np.float32)
        for y in range(mh):
            ys = np.minimum(np.arange(y * s, (y + 1) * s), h - 1)
            rows = depth[ys]
            for x in range(mw):
                xs = np.minimum(np.arange(x * s, (x + 1) * s), w - 1)
                m[y, x] = rows[:, xs].min()
        out.append(m)
    return out


def _rtz_bits(a):
    lib = oracle.lib()
    return np.array([lib.oracle_f32_to_f16_rtz(float(v)) for v in a.reshape(-1)], dtype=np.uint16)


def test_hzb_against_numpy_statement():
    c = case_s1()
    fr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 256, 256)
    vis, _ = fr.render_mesh(fr.instance_culling(4))
    hz = fr.build_hzb(vis, True, True, True)
    assert hz['mip0'] == (128, 128) and hz['mips'] == 8
    assert oracle.hzb_dims(1920, 1080) == (1024, 1024, 11) and oracle.hzb_dims(3840, 2160) == (2048, 2048, 12)   # hzb.cpp:49-56
    depth = (vis >> np.uint64(32)).astype(np.uint32).view(np.float32)
    ref = _numpy_hzb_min(depth, 128, 128, 8)
    for k in range(8):
        o = int(hz['offsets'][k])
        assert np.array_equal(hz['min'][o:o + ref[k].size], _rtz_bits(ref[k])), f'min mip {k}'
    # max chain: raw RTZ below mip 5, +1 ulp at mip 5 (hzb.hlsl:70)
    dmax = depth.reshape(128, 2, 128, 2).max(axis=(1, 3))
    assert np.array_equal(hz['max'][:128 * 128], _rtz_bits(dmax))
    o5 = int(hz['offsets'][5])
    m5 = depth.reshape(4, 64, 4, 64).max(axis=(1, 3))
    assert np.array_equal(hz['max'][o5:o5 + 16], _rtz_bits(m5) + 1)
    valid = depth[depth > 0]
    assert hz['range'][0] == valid.min().view(np.uint32) and hz['range'][1] == valid.max().view(np.uint32)


def _object_meshlets(scene, prim):
    d = scene.datas[0]
    ids = []
    for g in d['groups'][prim['meshletGroupOffset']:prim['meshletGroupOffset'] + prim['meshletGroupCount']]:
        o = prim['meshletGroupIndicesOffset'] + g['meshletOffset']
        ids.append(d['groupIndices'][o:o + g['meshletCount']])
    return np.unique(np.concatenate(ids)) + prim['meshletOffset']


def test_instance_culling_invariants():
    scene, info, frames = city_frames(frames=1)
    f = frames[0]
    for flags in (0b000, 0b001, 0b100, 0b111):
        fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, flags)
        post = fr.instance_culling(info['lod0MeshletCount'])
        assert np.array_equal(post['slot'], np.arange(len(post)))            # check(drawCmd.z == threadId), hzb_mainview_culling.hlsl:52
        assert len(post) <= info['lod0MeshletCount']                         # instance_culling.cpp:140 capacity assumption
        key = post['objectId'].astype(np.int64) << 32 | post['meshletId']
        assert len(np.unique(key)) == len(key)
        assert (np.diff(post['objectId'].astype(np.int64)) >= 0).all()       # canonical object order
    # with every cull off, the LOD cut tiles every face exactly once: the selected meshlets' areas add up to the LOD-0 area
    fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, 0)
    post = fr.instance_culling(info['lod0MeshletCount'])
    ms = scene.datas[0]['meshlets']
    ext = np.sort(ms['posMax'] - ms['posMin'], axis=1)
    area = ext[:, 1] * ext[:, 2]                                              # planar, axis-aligned patches
    lods_seen = set()
    for oid in np.unique(post['objectId']):
        sel = post['meshletId'][post['objectId'] == oid]
        lods_seen |= set(ms['lod'][sel].tolist())
        prim = scene.primitives[f['objects'][oid]['GLTFPrimitiveDetail']]
        all_ids = _object_meshlets(scene, prim)
        lod0_ids = all_ids[ms['lod'][all_ids] == 0]
        assert np.isclose(area[sel].sum(), area[lod0_ids].sum(), rtol=1e-4), f'object {oid}: LOD cut does not tile the surface'
    assert len(lods_seen) >= 2, 'the case should exercise both branches of the LOD test'


def test_two_phase_hzb_is_result_invariant():
    """Occlusion culling only removes meshlets that cannot change a pixel: the two-phase frame equals brute force."""
    scene, info, frames = city_frames(frames=3)
    prev = None
    for f in frames:
        two = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], prev)
        brute = oracle.run_frame(scene, f['view'], f['iview'], f['objects'], 640, 360, info['lod0MeshletCount'], None)
        assert np.array_equal(two['post_cull'], brute['post_cull'])
        assert np.array_equal(two['vis64'], brute['vis64'])
        if prev is not None:
            a, b = two['stage0_visible'], two['stage0_rejected']
            assert len(b) > 0
            merged = np.sort(np.concatenate([a['slot'], b['slot']]))
            assert np.array_equal(merged, two['post_cull']['slot'])           # partition of the input
            assert (np.diff(a['slot'].astype(np.int64)) > 0).all() and (np.diff(b['slot'].astype(np.int64)) > 0).all()
        prev = two['hzb']


def test_hzb_flag_off_keeps_everything():
    scene, info, frames = city_frames(frames=2)
    f = frames[1]
    fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, 0b101)
    post = fr.instance_culling(info['lod0MeshletCount'])
    vis, _ = fr.render_mesh(post)
    hz = fr.build_hzb(vis)
    v, r = fr.hzb_culling(hz, True, post)
    assert len(v) == len(post) and len(r) == 0


def test_tile_union_equals_full_frame():
    c = case_stream(800, 512, 256, 30.0, seed=9)
    fr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 512, 256)
    full, _ = fr.render_mesh(c['cmds'])
    acc = np.zeros_like(full)
    for tile in ((0, 0, 256, 128), (256, 0, 512, 128), (0, 128, 256, 256), (256, 128, 512, 256)):
        part, _ = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 512, 256, tile=tile).render_mesh(c['cmds'])
        acc |= part
    assert np.array_equal(acc, full)


def test_masked_and_blend_materials_filtered():
    """pipeline_filter.hlsl:41-130 + mesh_raster.hlsl:198-204: blend is never drawn, masked is clipped by the cutoff."""
    c = case_s1()
    scene = c['scene']
    fr = oracle.Frame(scene, c['view'], c['iview'], c['objects'], 256, 256)
    post = fr.instance_culling(4)
    base = (fr.render_mesh(post)[0] != 0).sum()
    scene.materials[0]['alphaMode'] = 2
    assert (fr.render_mesh(post)[0] != 0).sum() == 0
    scene.materials[0]['alphaMode'] = 1
    scene.materials[0]['alphaCutOff'] = 0.5
    assert (fr.render_mesh(post)[0] != 0).sum() == base
    scene.materials[0]['baseColorFactor'][3] = 0.25
    assert (fr.render_mesh(post)[0] != 0).sum() == 0


def test_empty_and_ragged_inputs():
    """empty draw list, a list whose last group is ragged, a meshlet stream of length 1."""
    c = case_stream(5, 256, 256, 4.0)            # 5 meshlets -> groups of 4 + 1
    fr = oracle.Frame(c['scene'], c['view'], c['iview'], c['objects'], 256, 256)
    assert len(fr.instance_culling(5)) == 5
    vis, st = fr.render_mesh(c['cmds'][:0])
    assert not vis.any() and st['meshletsRastered'] == 0
    one, _ = fr.render_mesh(c['cmds'][:1])
    assert one.any()


def test_bvh_traversal_equals_flat_loop():
    """N4: walking the cluster-group BVH (never traversed by the reference) must give the flat loop's list."""
    from chord_b200 import camera as cam, synth
    scene, info, frames = city_frames(frames=3, n_buildings=16)
    # a distant view: whole sub-trees of the node hierarchy fall below one pixel and are skipped
    far = cam.Camera((80.0, 900.0, 5200.0), (0.0, -0.17, -1.0), width=640, height=360)
    fview = cam.fill_view(far)
    frames = frames + [dict(view=fview, iview=cam.fill_instance_view(fview), objects=synth.refill_objects(info['transforms'], far), camera=far)]
    pruned_any = False
    for f in frames:
        for flags in (0b111, 0b000):
            fr = oracle.Frame(scene, f['view'], f['iview'], f['objects'], 640, 360, flags)
            flat = fr.instance_culling(info['lod0MeshletCount'])
            via_bvh, cand = fr.instance_culling_bvh(info['lod0MeshletCount'])
            assert np.array_equal(flat, via_bvh)
            bits, n, total = fr.bvh_candidates(info['meshletGroupCount'])
            assert n == cand and n <= total
            pruned_any |= n < total
    assert pruned_any, 'the case should prune at least one subtree'
    # the builder's invariants (nanite_builder.cpp:415): root counts every node; children are all-or-none
    nodes = scene.datas[0]['bvhNodes']
    for p in scene.primitives:
        root = nodes[p['bvhNodeOffset']]
        sub = nodes[p['bvhNodeOffset']:p['bvhNodeOffset'] + root['bvhNodeCount']]
        assert sub['leafMeshletGroupCount'].sum() == p['meshletGroupCount']
        valid = (sub['children'] != 0xFFFFFFFF).sum(axis=1)
        assert set(valid.tolist()) <= {0, 8}
