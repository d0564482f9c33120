"""CPU oracle for chord's visibility-buffer path — TEST INFRASTRUCTURE ONLY (parity unpinned, see chord_oracle.h).

ctypes wrapper over oracle/_build/libchord_oracle.so.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this package; chord_b200 (the product) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from chord_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libchord_oracle.so')


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ('chord_oracle.c', 'chord_oracle.h')]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(['make', '-C', _HERE, '-s'])
    return _SO


REFERENCE = os.environ.get('CHORD_REFERENCE', '/root/reference')
_REF_SO = os.path.join(_HERE, '_ref', 'libmeshopt_ref.so')


REF_LAYOUT_PROBE = os.path.join(_HERE, '_ref', 'ref_layout_probe')


def ref_struct_layout():
    """sizeof / offsetof of the a1 structs as the REFERENCE's own headers (install/resource/shader/base.h, gltf.h, compiled
    with g++ + its glm, oracle/ref_layout_probe.cpp) declare them, or None when the probe was never built here."""
    import json
    build_ref()
    if not os.path.exists(REF_LAYOUT_PROBE):
        return None
    return json.loads(subprocess.check_output([REF_LAYOUT_PROBE], text=True))


def build_ref(force=False):
    """oracle/_ref/libmeshopt_ref.so: the reference's vendored meshoptimizer 0.21 compiled from its sources in place (the
    only part of the reference that builds offline) + oracle/meshopt_ref_wrapper.cpp.  Returns the path, or None when
    neither the reference tree nor a prebuilt library is there (the GPU box only has what travelled with the repo)."""
    have_src = os.path.isdir(os.path.join(REFERENCE, 'source', 'asset', 'meshoptimizer'))
    pairs = ((os.path.join(_HERE, 'meshopt_ref_wrapper.cpp'), _REF_SO), (os.path.join(_HERE, 'tinygltf_ref_wrapper.cpp'), os.path.join(_HERE, '_ref', 'libtinygltf_ref.so')),
             (os.path.join(_HERE, 'ref_layout_probe.cpp'), REF_LAYOUT_PROBE))
    stale = any(not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so) for src, so in pairs)
    if have_src and (force or stale):
        try:
            subprocess.check_call(['make', '-C', _HERE, '-s', 'ref', f'REFERENCE={REFERENCE}'])
        except (subprocess.CalledProcessError, OSError) as e:        # a checker that cannot be built is skipped, never fatal
            print(f'oracle/_ref not built: {e}')
    return _REF_SO if os.path.exists(_REF_SO) else None


_ref_lib = None


def meshopt_ref():
    """ctypes handle of oracle/_ref/libmeshopt_ref.so or None (tests skip then)."""
    global _ref_lib
    if _ref_lib is None:
        so = build_ref()
        if so is None:
            return None
        L = C.CDLL(so)
        L.meshopt_ref_build_meshlets.restype = C.c_size_t
        L.meshopt_ref_build_meshlets_bound.restype = C.c_size_t
        L.meshopt_ref_simplify.restype = C.c_size_t
        L.meshopt_ref_simplify_scale.restype = C.c_float
        _ref_lib = L
    return _ref_lib


_TG_SO = os.path.join(_HERE, '_ref', 'libtinygltf_ref.so')
_tg_lib = None


def tinygltf_ref():
    """ctypes handle of oracle/_ref/libtinygltf_ref.so (the reference's glTF parser behind oracle/tinygltf_ref_wrapper.cpp) or None."""
    global _tg_lib
    if _tg_lib is None:
        build_ref()
        if not os.path.exists(_TG_SO):
            return None
        L = C.CDLL(_TG_SO)
        L.tg_last_error.restype = C.c_char_p
        L.tg_accessor_read.restype = C.c_longlong
        _tg_lib = L
    return _tg_lib


class _Frame(C.Structure):
    _fields_ = [('scene', C.c_void_p), ('view', C.c_void_p), ('iview', C.c_void_p), ('objects', C.c_void_p),
                ('objectCount', C.c_uint32), ('width', C.c_uint32), ('height', C.c_uint32),
                ('tile', abi.Rect), ('switchFlags', C.c_uint32), ('threads', C.c_int),
                ('textures', C.c_void_p), ('textureCount', C.c_uint32), ('samplers', C.c_void_p), ('samplerCount', C.c_uint32),
                ('passFlags', C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_instance_culling.restype = C.c_uint32
        _lib.oracle_f32_to_f16_rtz.restype = C.c_uint16
        _lib.oracle_f32_to_f16_rtz.argtypes = [C.c_float]
        _lib.oracle_f16_to_f32.restype = C.c_float
        _lib.oracle_f16_to_f32.argtypes = [C.c_uint16]
    return _lib


def hzb_dims(w, h):
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().oracle_hzb_dims(C.c_uint32(w), C.c_uint32(h), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def hzb_mip_offsets(mip0w, mip0h, mips):
    """Texel offset of every mip inside one contiguous chain; returns (offsets u64[mips], total texels)."""
    offs, o = [], 0
    for m in range(mips):
        offs.append(o)
        o += max(mip0w >> m, 1) * max(mip0h >> m, 1)
    return np.array(offs, dtype=np.uint64), o


class Frame:
    """One view of one scene (mirrors the arguments chord_frame_begin takes)."""

    def __init__(self, scene, view, iview, objects, width, height, switch_flags=0b111, tile=None, threads=0):
        self.scene, self.width, self.height = scene, int(width), int(height)
        self.view = np.ascontiguousarray(view, dtype=abi.PerframeCameraView).reshape(())
        self.iview = np.ascontiguousarray(iview, dtype=abi.InstanceCullingViewInfo).reshape(())
        self.objects = np.ascontiguousarray(objects, dtype=abi.GPUObjectGLTFPrimitive)
        f = _Frame()
        f.scene = C.addressof(scene.desc)
        f.view, f.iview, f.objects = self.view.ctypes.data, self.iview.ctypes.data, self.objects.ctypes.data
        f.objectCount, f.width, f.height = len(self.objects), self.width, self.height
        t = tile if tile is not None else (0, 0, self.width, self.height)
        f.tile = abi.Rect(*t)
        f.switchFlags, f.threads = switch_flags, threads
        tt = getattr(scene, 'textures', None)
        if tt is not None:
            f.textures, f.textureCount = C.addressof(tt.tex), tt.n_tex
            f.samplers, f.samplerCount = C.addressof(tt.smp), tt.n_smp
        self.c = f

    def instance_culling(self, capacity):
        out = np.zeros(max(capacity, 1), dtype=abi.DrawCmd)
        n = lib().oracle_instance_culling(C.byref(self.c), C.c_void_p(out.ctypes.data), C.c_uint32(capacity))
        assert n <= capacity, f'oracle_instance_culling overflow {n} > {capacity}'
        return out[:n].copy()

    def instance_culling_bvh(self, capacity):
        """Same list through the cluster-group BVH; returns (cmds, candidate group count)."""
        out = np.zeros(max(capacity, 1), dtype=abi.DrawCmd)
        cand = C.c_uint32()
        lib().oracle_instance_culling_bvh.restype = C.c_uint32
        n = lib().oracle_instance_culling_bvh(C.byref(self.c), C.c_void_p(out.ctypes.data), C.c_uint32(capacity), C.byref(cand))
        assert n <= capacity
        return out[:n].copy(), cand.value

    def bvh_candidates(self, total_groups):
        words = (total_groups + 31) // 32 + 1
        bits = np.zeros(words, dtype=np.uint32)
        tot = C.c_uint32()
        lib().oracle_bvh_candidates.restype = C.c_uint32
        n = lib().oracle_bvh_candidates(C.byref(self.c), C.c_void_p(bits.ctypes.data), C.c_uint32(words), C.byref(tot))
        return bits, n, tot.value

    def hzb_culling_generic(self, hzb, hzb_view, extent_scale, use_last_frame, cmds):
        """hzbCullingGeneric against the HZB of another view (hzb_culling_generic.hlsl:37-172)."""
        cmds = np.ascontiguousarray(cmds, dtype=abi.DrawCmd)
        hv = np.ascontiguousarray(hzb_view, dtype=abi.InstanceCullingViewInfo).reshape(())
        vis = np.zeros(max(len(cmds), 1), dtype=abi.DrawCmd)
        nv = C.c_uint32()
        lib().oracle_hzb_culling_generic(C.byref(self.c), C.c_void_p(hv.ctypes.data), C.c_float(extent_scale), C.c_int(1 if use_last_frame else 0),
                                         C.c_void_p(hzb['min'].ctypes.data), C.c_void_p(hzb['offsets'].ctypes.data), C.c_uint32(hzb['mip0'][0]),
                                         C.c_void_p(cmds.ctypes.data), C.c_uint32(len(cmds)), C.c_void_p(vis.ctypes.data), C.byref(nv))
        return vis[:nv.value].copy()

    def render_mesh_depth(self, cmds, vis64=None, pass_flags=abi.PASS_DEPTH | abi.PASS_DEPTH_CLAMP):
        """renderMeshDepth (mesh_raster.cpp:159-206): PASS_TYPE_DEPTH, always two-sided, depth clamp."""
        self.c.passFlags = pass_flags
        try:
            return self.render_mesh(cmds, vis64)
        finally:
            self.c.passFlags = 0

    def render_mesh(self, cmds, vis64=None):
        cmds = np.ascontiguousarray(cmds, dtype=abi.DrawCmd)
        if vis64 is None:
            vis64 = np.zeros((self.height, self.width), dtype=np.uint64)
        stats = abi.RasterStats()
        lib().oracle_render_mesh(C.byref(self.c), C.c_void_p(cmds.ctypes.data), C.c_uint32(len(cmds)),
                                 C.c_void_p(vis64.ctypes.data), C.byref(stats))
        return vis64, stats.as_dict()

    def build_hzb(self, vis64, b_min=True, b_max=False, b_range=False):
        w0, h0, mips = hzb_dims(self.width, self.height)
        offs, total = hzb_mip_offsets(w0, h0, mips)
        mn = np.zeros(total, dtype=np.uint16)
        mx = np.zeros(total, dtype=np.uint16)
        rng = np.zeros(2, dtype=np.uint32)
        lib().oracle_build_hzb(C.c_void_p(vis64.ctypes.data), C.c_uint32(self.width), C.c_uint32(self.height),
                               C.c_int(b_min), C.c_int(b_max), C.c_int(b_range),
                               C.c_void_p(mn.ctypes.data), C.c_void_p(mx.ctypes.data), C.c_void_p(offs.ctypes.data),
                               C.c_uint32(w0), C.c_uint32(h0), C.c_uint32(mips), C.c_void_p(rng.ctypes.data))
        return dict(min=mn, max=mx, offsets=offs, mip0=(w0, h0), mips=mips, range=rng)

    def hzb_culling(self, hzb, first_stage, cmds):
        cmds = np.ascontiguousarray(cmds, dtype=abi.DrawCmd)
        vis = np.zeros(max(len(cmds), 1), dtype=abi.DrawCmd)
        rej = np.zeros(max(len(cmds), 1), dtype=abi.DrawCmd)
        nv, nr = C.c_uint32(), C.c_uint32()
        lib().oracle_hzb_culling(C.byref(self.c), C.c_int(first_stage),
                                 C.c_void_p(hzb['min'].ctypes.data), C.c_void_p(hzb['offsets'].ctypes.data),
                                 C.c_uint32(hzb['mip0'][0]), C.c_uint32(hzb['mip0'][1]), C.c_uint32(hzb['mips']),
                                 C.c_void_p(cmds.ctypes.data), C.c_uint32(len(cmds)),
                                 C.c_void_p(vis.ctypes.data), C.byref(nv), C.c_void_p(rej.ctypes.data), C.byref(nr))
        return vis[:nv.value].copy(), rej[:nr.value].copy()

    def visibility_mark(self, vis64, post_cull):
        post_cull = np.ascontiguousarray(post_cull, dtype=abi.DrawCmd)
        mw, mh = (self.width + 7) // 8, (self.height + 7) // 8
        out = np.zeros((mh, mw, 4), dtype=np.uint32)
        lib().oracle_visibility_mark(C.byref(self.c), C.c_void_p(vis64.ctypes.data), C.c_void_p(post_cull.ctypes.data),
                                     C.c_uint32(len(post_cull)), C.c_void_p(out.ctypes.data))
        return out


def lighting(frame, vis64, post_cull, transmittance, irradiance):
    """K11 on the CPU; returns (6, H, W, 4) float32 planes in CHORD_GBUF_* order."""
    post_cull = np.ascontiguousarray(post_cull, dtype=abi.DrawCmd)
    tr = np.ascontiguousarray(transmittance, dtype=np.float16)
    ir = np.ascontiguousarray(irradiance, dtype=np.float16)
    out = np.zeros((6, frame.height, frame.width, 4), dtype=np.float32)
    ptrs = (C.c_void_p * 6)(*[out[i].ctypes.data for i in range(6)])
    lib().oracle_lighting(C.byref(frame.c), C.c_void_p(vis64.ctypes.data), C.c_void_p(post_cull.ctypes.data),
                          C.c_uint32(len(post_cull)), C.c_void_p(tr.ctypes.data), C.c_void_p(ir.ctypes.data), ptrs)
    return out


def sky(frame, vis64, transmittance, irradiance, scattering, single_mie=None, color=None):
    """The sky branch of the lighting pass (LIGHTING_TYPE == None, lighting.hlsl:75-166) on the CPU: writes the colour plane's
    texels that hold no geometry.  ``scattering`` (32, 128, 256, 4) float16; ``single_mie`` same shape or None (combined)."""
    tr = np.ascontiguousarray(transmittance, dtype=np.float16)
    ir = np.ascontiguousarray(irradiance, dtype=np.float16)
    sc = np.ascontiguousarray(scattering, dtype=np.float16)
    assert sc.shape == (32, 128, 256, 4)
    mie = np.ascontiguousarray(single_mie, dtype=np.float16) if single_mie is not None else None
    if color is None:
        color = np.zeros((frame.height, frame.width, 4), dtype=np.float32)
    lib().oracle_sky(C.byref(frame.c), C.c_void_p(vis64.ctypes.data), C.c_void_p(tr.ctypes.data), C.c_void_p(ir.ctypes.data),
                     C.c_void_p(sc.ctypes.data), C.c_void_p(mie.ctypes.data) if mie is not None else None, C.c_void_p(color.ctypes.data))
    return color


def sample_alpha(table, tex_id, smp_id, u, v, dudx, dvdx, dudy, dvdy):
    """The specified sampler of the masked bucket on one sample."""
    fn = lib().oracle_sample_alpha
    fn.restype = C.c_float
    fn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 6
    t = C.addressof(table.tex[tex_id]) if tex_id < table.n_tex else None
    s_ = C.addressof(table.smp[smp_id]) if smp_id < table.n_smp else None
    return fn(t, s_, u, v, dudx, dvdx, dudy, dvdy)


def decode(vis64, post_cull):
    post_cull = np.ascontiguousarray(post_cull, dtype=abi.DrawCmd)
    out = np.zeros(vis64.shape + (3,), dtype=np.uint32)
    lib().oracle_decode(C.c_void_p(vis64.ctypes.data), C.c_uint32(vis64.size), C.c_void_p(post_cull.ctypes.data),
                        C.c_uint32(len(post_cull)), C.c_void_p(out.ctypes.data))
    return out


def raster_clip_triangle(width, height, clip, vis_id, two_sided, vis64=None, tile=None):
    """Rasterise one clip-space triangle ((3,4) float32) into a packed surface."""
    clip = np.ascontiguousarray(clip, dtype=np.float32).reshape(12)
    if vis64 is None:
        vis64 = np.zeros((height, width), dtype=np.uint64)
    t = C.byref(abi.Rect(*tile)) if tile is not None else None
    lib().oracle_raster_clip_triangle(C.c_uint32(width), C.c_uint32(height), t, C.c_void_p(clip.ctypes.data),
                                      C.c_uint32(vis_id), C.c_int(two_sided), C.c_void_p(vis64.ctypes.data))
    return vis64


def run_frame(scene, view, iview, objects, width, height, lod0_capacity, prev_hzb=None, switch_flags=0b111, threads=0,
              tile=None):
    """The reference frame sequence (renderer.cpp:311-363) on the CPU: instance cull -> stage 0 -> HZB -> stage 1 -> HZB."""
    fr = Frame(scene, view, iview, objects, width, height, switch_flags, tile, threads)
    post = fr.instance_culling(lod0_capacity)
    out = dict(post_cull=post)
    vis64 = np.zeros((height, width), dtype=np.uint64)
    if prev_hzb is not None and (switch_flags & abi.FLAG_HZB):
        vis0, rej0 = fr.hzb_culling(prev_hzb, True, post)
        fr.render_mesh(vis0, vis64)
        mid = fr.build_hzb(vis64, True, False, False)
        vis1, _ = fr.hzb_culling(mid, False, rej0)
        fr.render_mesh(vis1, vis64)
        out.update(stage0_visible=vis0, stage0_rejected=rej0, stage1_visible=vis1, mid_hzb=mid)
    else:
        fr.render_mesh(post, vis64)
    out['hzb'] = fr.build_hzb(vis64, True, True, True)
    out['vis64'] = vis64
    out['frame'] = fr
    return out
