"""ctypes bindings over libchord_b200.so (include/chord_b200.h).  No fallback path exists: a missing library or
a missing sm_100 device raises."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libchord_b200.so')

# every symbol include/chord_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    'chord_last_error', 'chord_abi_version', 'chord_device_count',
    'chord_ctx_create', 'chord_ctx_destroy', 'chord_ctx_refcount', 'chord_ctx_stream', 'chord_ctx_sync', 'chord_ctx_set_switch_flags',
    'chord_ctx_get_switch_flags', 'chord_ctx_set_raster_tuning', 'chord_ctx_kernel_launches', 'chord_scene_upload', 'chord_textures_upload',
    'chord_cmdlist_create', 'chord_cmdlist_destroy', 'chord_cmdlist_upload', 'chord_cmdlist_readback',
    'chord_cmdlist_upload_async', 'chord_readback_visibility32_async', 'chord_frame_wait_readbacks', 'chord_frame_set_readback_stream', 'chord_cmdlist_capacity', 'chord_cmdlist_count_ptr', 'chord_cmdlist_cmds_ptr',
    'chord_hzb_create', 'chord_hzb_destroy', 'chord_hzb_info', 'chord_hzb_readback', 'chord_hzb_readback_valid_range',
    'chord_hzb_chain_ptr', 'chord_hzb_chain_bytes', 'chord_hzb_mip_offset', 'chord_hzb_set_valid', 'chord_hzb_is_valid',
    'chord_hzb_finish_top', 'chord_hzb_tile_bytes', 'chord_hzb_pack_tile', 'chord_hzb_unpack_tiles',
    'chord_frame_create', 'chord_frame_destroy', 'chord_frame_begin', 'chord_instance_culling', 'chord_instance_culling_bvh',
    'chord_readback_bvh_candidates', 'chord_hzb_culling',
    'chord_render_mesh', 'chord_render_mesh_depth', 'chord_hzb_culling_generic', 'chord_visibility_stage0', 'chord_visibility_stage1', 'chord_build_hzb',
    'chord_visibility_mark', 'chord_atmosphere_luts_upload', 'chord_atmosphere_scattering_upload', 'chord_lighting',
    'chord_readback_vis64', 'chord_readback_visibility32', 'chord_readback_depth32', 'chord_readback_decoded',
    'chord_readback_tile_marker', 'chord_readback_gbuffer_f32', 'chord_frame_set_gbuffer_packed', 'chord_readback_gbuffer_packed', 'chord_frame_vis64_ptr', 'chord_frame_dims',
    'chord_frame_raster_stats', 'chord_frame_reset_stats', 'chord_frame_set_queue_limits',
    'chord_cascade_views_create', 'chord_cascade_views_destroy', 'chord_cascade_setup', 'chord_cascade_views_readback',
    'chord_cascade_views_upload', 'chord_frame_begin_cascade', 'chord_hzb_culling_generic_cascade',
    'chord_selftest_vertex_records',
    'chord_renderer_create', 'chord_renderer_destroy', 'chord_renderer_set_tile', 'chord_renderer_render', 'chord_renderer_clear_history',
    'chord_renderer_frame', 'chord_renderer_post_cull', 'chord_renderer_history_hzb', 'chord_renderer_graph_replays', 'chord_renderer_pass_times',
    'chord_renderer_exchange_export', 'chord_renderer_exchange_connect', 'chord_renderer_history_sync', 'chord_renderer_exchange_status', 'chord_renderer_set_exchange_timeout',
    'chord_nccl_unique_id', 'chord_nccl_comm_create', 'chord_nccl_comm_destroy', 'chord_hzb_allgather', 'chord_renderer_exchange_connect_nccl',
    'chord_graph_begin_capture', 'chord_graph_end_capture', 'chord_graph_launch', 'chord_graph_kernel_nodes', 'chord_graph_destroy', 'chord_graph_update_frame_inputs',
]

_lib = None


class ChordError(RuntimeError):
    pass


def lib():
    """Load libchord_b200.so.  Raises if it has not been built (python -m chord_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ChordError(f'{LIB_PATH} is missing: build the CUDA library first (python -m chord_b200.build); '
                             'there is no CPU fallback')
        L = C.CDLL(LIB_PATH)
        L.chord_last_error.restype = C.c_char_p
        L.chord_ctx_stream.restype = C.c_void_p
        L.chord_ctx_kernel_launches.restype = C.c_uint64
        L.chord_ctx_get_switch_flags.restype = C.c_uint32
        L.chord_ctx_refcount.restype = C.c_uint32
        L.chord_cmdlist_capacity.restype = C.c_uint32
        L.chord_renderer_graph_replays.restype = C.c_uint64
        for n in ('chord_renderer_destroy', 'chord_nccl_comm_destroy'):
            getattr(L, n).restype = None
        for n in ('chord_cmdlist_count_ptr', 'chord_cmdlist_cmds_ptr', 'chord_hzb_chain_ptr', 'chord_frame_vis64_ptr', 'chord_renderer_frame',
                  'chord_renderer_post_cull', 'chord_renderer_history_hzb'):
            getattr(L, n).restype = C.c_void_p
        L.chord_graph_kernel_nodes.restype = C.c_uint32
        for n in ('chord_ctx_destroy', 'chord_cmdlist_destroy', 'chord_hzb_destroy', 'chord_frame_destroy', 'chord_graph_destroy',
                  'chord_cascade_views_destroy'):
            getattr(L, n).restype = None
        L.chord_hzb_chain_bytes.restype = C.c_uint64
        L.chord_hzb_mip_offset.restype = C.c_uint64
        for n in SYMBOLS:
            f = getattr(L, n)
            if f.restype is C.c_int and n not in ('chord_abi_version', 'chord_hzb_is_valid', 'chord_device_count'):
                f.errcheck = _check
        _lib = L
    return _lib


def _check(rc, func, args):
    if rc != 0:
        raise ChordError(f'{func.__name__} -> {rc}: {_lib.chord_last_error().decode()}')
    return rc


def _p(a):
    return C.c_void_p(a.ctypes.data)


class Context:
    """graphics::Context + GraphicsQueue stand-in: one CUDA device, one stream."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        lib().chord_ctx_create(C.c_int(device), C.byref(self.h))
        self.device = device
        self._scene = None

    def close(self):
        if self.h:
            lib().chord_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def refcount(self):
        """owner + live child handles (chord_ctx_refcount)"""
        return int(lib().chord_ctx_refcount(self.h))

    @property
    def stream(self):
        return lib().chord_ctx_stream(self.h)

    def sync(self):
        lib().chord_ctx_sync(self.h)

    def set_switch_flags(self, flags):
        lib().chord_ctx_set_switch_flags(self.h, C.c_uint32(flags))

    def set_raster_tuning(self, variant=-1, ctas_per_sm=-1, warp_min=-1):
        """chord_ctx_set_raster_tuning: launch shape of the meshlet rasteriser (negative = keep)."""
        lib().chord_ctx_set_raster_tuning(self.h, C.c_int(variant), C.c_int(ctas_per_sm), C.c_int64(warp_min))

    @property
    def switch_flags(self):
        return lib().chord_ctx_get_switch_flags(self.h)

    @property
    def kernel_launches(self):
        return int(lib().chord_ctx_kernel_launches(self.h))

    def upload_scene(self, scene: abi.HostScene):
        lib().chord_scene_upload(self.h, C.byref(scene.desc))
        self._scene = scene
        self.upload_textures(getattr(scene, 'textures', None))

    def upload_textures(self, table):
        """Material texture table of the masked bucket (abi.TextureTable); None = every id is builtin white."""
        if table is None:
            lib().chord_textures_upload(self.h, None, C.c_uint32(0), None, C.c_uint32(0))
        else:
            lib().chord_textures_upload(self.h, C.byref(table.tex), C.c_uint32(table.n_tex), C.byref(table.smp), C.c_uint32(table.n_smp))

    def upload_luts(self, transmittance, irradiance):
        t = np.ascontiguousarray(transmittance, dtype=np.float16)
        i = np.ascontiguousarray(irradiance, dtype=np.float16)
        assert t.shape == (64, 256, 4) and i.shape == (16, 64, 4)
        luts = abi.AtmosphereLuts(t.ctypes.data, i.ctypes.data)
        lib().chord_atmosphere_luts_upload(self.h, C.byref(luts))

    def upload_scattering(self, scattering, single_mie=None):
        """Scattering volumes of the sky branch, (32, 128, 256, 4) float16 each; ``scattering=None`` removes the branch."""
        if scattering is None:
            lib().chord_atmosphere_scattering_upload(self.h, None, None)
            return
        s = np.ascontiguousarray(scattering, dtype=np.float16)
        assert s.shape == (32, 128, 256, 4)
        m = np.ascontiguousarray(single_mie, dtype=np.float16) if single_mie is not None else None
        assert m is None or m.shape == s.shape
        lib().chord_atmosphere_scattering_upload(self.h, C.c_void_p(s.ctypes.data), C.c_void_p(m.ctypes.data) if m is not None else None)


class CascadeViews:
    """Device buffer of the cascades' InstanceCullingViewInfo records, written by CascadeSetupCS on the GPU."""

    def __init__(self, ctx):
        self.ctx, self.h = ctx, C.c_void_p()
        lib().chord_cascade_views_create(ctx.h, C.byref(self.h))

    def setup(self, main_frame, cfg, light_dir, tick, cache_valid=False, valid_range_hzb=None):
        """chord_cascade_setup: ``cfg`` = camera.CascadeShadowMapConfig; ``valid_range_hzb`` = the Hzb whose valid range (SDSM) to read."""
        d = abi.CascadeSetupDesc()
        d.lightDir = (C.c_float * 3)(*[float(x) for x in light_dir])
        d.cascadeStartDistance, d.cascadeEndDistance = cfg.cascadeStartDistance, cfg.cascadeEndDistance
        d.splitLambda, d.farCascadeSplitLambda, d.farCascadeEndDistance = cfg.splitLambda, cfg.farCascadeSplitLambda, cfg.farCascadeEndDistance
        d.radiusScale = cfg.radiusScaleFixed
        d.cascadeCount, d.realtimeCount, d.cascadeDim = cfg.cascadeCount, cfg.realtimeCascadeCount, cfg.cascadeDim
        d.bCacheValid, d.tickCount = int(bool(cache_valid)), int(tick)
        lib().chord_cascade_setup(main_frame.h, C.byref(d), valid_range_hzb.h if valid_range_hzb is not None else None, self.h)

    def readback(self, count):
        out = np.zeros(count, dtype=abi.InstanceCullingViewInfo)
        lib().chord_cascade_views_readback(self.h, _p(out), C.c_uint32(count))
        return out

    def upload(self, views):
        v = np.ascontiguousarray(views, dtype=abi.InstanceCullingViewInfo)
        lib().chord_cascade_views_upload(self.h, _p(v), C.c_uint32(len(v)))

    def close(self):
        if self.h:
            lib().chord_cascade_views_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Graph:
    """A recorded sequence of calls on one context (chord_graph_*): ``with Graph.capture(ctx) as g: ...calls...`` then
    ``g.launch()``."""

    def __init__(self, ctx):
        self.ctx, self.h = ctx, C.c_void_p()

    @classmethod
    def capture(cls, ctx):
        return cls(ctx)

    def __enter__(self):
        lib().chord_graph_begin_capture(self.ctx.h)
        return self

    def __exit__(self, et, ev, tb):
        if et is not None:
            try:
                lib().chord_graph_end_capture(self.ctx.h, C.byref(self.h))
            except ChordError:
                pass
            return False
        lib().chord_graph_end_capture(self.ctx.h, C.byref(self.h))
        return False

    @property
    def kernel_nodes(self):
        return int(lib().chord_graph_kernel_nodes(self.h))

    def launch(self):
        lib().chord_graph_launch(self.h)

    def update_frame_inputs(self, frame, view, iview, objects):
        """chord_graph_update_frame_inputs: the next replay renders this camera / these object transforms."""
        view = np.ascontiguousarray(view, dtype=abi.PerframeCameraView).reshape(())
        iview = np.ascontiguousarray(iview, dtype=abi.InstanceCullingViewInfo).reshape(())
        objects = np.ascontiguousarray(objects, dtype=abi.GPUObjectGLTFPrimitive)
        lib().chord_graph_update_frame_inputs(self.h, frame.h, _p(view), _p(iview), _p(objects), C.c_uint32(len(objects)))

    def close(self):
        if self.h:
            lib().chord_graph_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CmdList:
    """CountAndCmdBuffer (postprocessing.h:48)."""

    def __init__(self, ctx, capacity, _borrowed=None):
        self.ctx, self.capacity = ctx, int(capacity)
        self.h = C.c_void_p()
        self._borrowed = _borrowed is not None
        if self._borrowed:
            self.h = C.c_void_p(_borrowed)                # owned by a Renderer
        else:
            lib().chord_cmdlist_create(ctx.h, C.c_uint32(capacity), C.byref(self.h))

    def close(self):
        if self._borrowed:
            self.h = C.c_void_p()
        if self.h:
            lib().chord_cmdlist_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, cmds):
        cmds = np.ascontiguousarray(cmds, dtype=abi.DrawCmd)
        lib().chord_cmdlist_upload(self.h, _p(cmds), C.c_uint32(len(cmds)))

    def readback(self):
        n = C.c_uint32()
        out = np.zeros(self.capacity, dtype=abi.DrawCmd)
        lib().chord_cmdlist_readback(self.h, _p(out), C.c_uint32(self.capacity), C.byref(n))
        return out[:n.value].copy()

    def count(self):
        n = C.c_uint32()
        lib().chord_cmdlist_readback(self.h, None, C.c_uint32(0), C.byref(n))
        return n.value

    @property
    def cmds_ptr(self):
        return lib().chord_cmdlist_cmds_ptr(self.h)

    @property
    def count_ptr(self):
        return lib().chord_cmdlist_count_ptr(self.h)


class Hzb:
    """HZBContext (render_helper.h:411-451)."""

    def __init__(self, ctx, width, height, _borrowed=None):
        self.ctx, self.width, self.height = ctx, int(width), int(height)
        self.h = C.c_void_p()
        self._borrowed = _borrowed is not None
        if self._borrowed:
            self.h = C.c_void_p(_borrowed)                # owned by a Renderer
        else:
            lib().chord_hzb_create(ctx.h, C.c_uint32(width), C.c_uint32(height), C.byref(self.h))
        a, b, c, d, e = (C.c_uint32() for _ in range(5))
        lib().chord_hzb_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e))
        self.mip0 = (a.value, b.value)
        self.mips = c.value
        self.valid_area = (d.value, e.value)

    def close(self):
        if self._borrowed:
            self.h = C.c_void_p()
        if self.h:
            lib().chord_hzb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def valid(self):
        return bool(lib().chord_hzb_is_valid(self.h))

    def set_valid(self, v):
        lib().chord_hzb_set_valid(self.h, C.c_int(1 if v else 0))

    def mip_dims(self, mip):
        return max(self.mip0[0] >> mip, 1), max(self.mip0[1] >> mip, 1)

    def readback(self, which='min', mip=0):
        w, h = self.mip_dims(mip)
        out = np.zeros((h, w), dtype=np.uint16)
        lib().chord_hzb_readback(self.h, C.c_int(0 if which == 'min' else 1), C.c_uint32(mip), _p(out))
        return out

    def readback_chain(self, which='min'):
        return np.concatenate([self.readback(which, m).reshape(-1) for m in range(self.mips)])

    def valid_range(self):
        out = np.zeros(2, dtype=np.uint32)
        lib().chord_hzb_readback_valid_range(self.h, _p(out))
        return out

    def tile_bytes(self, tile):
        b = C.c_uint64()
        lib().chord_hzb_tile_bytes(self.h, C.byref(abi.Rect(*tile)), C.byref(b))
        return int(b.value)

    def pack_tile(self, which, tile, dst_ptr):
        lib().chord_hzb_pack_tile(self.h, C.c_int(0 if which == 'min' else 1), C.byref(abi.Rect(*tile)), C.c_void_p(dst_ptr))

    def unpack_tiles(self, which, tiles, src_ptr, slot_bytes):
        arr = (abi.Rect * len(tiles))(*[abi.Rect(*t) for t in tiles])
        lib().chord_hzb_unpack_tiles(self.h, C.c_int(0 if which == 'min' else 1), arr, C.c_uint32(len(tiles)),
                                     C.c_void_p(src_ptr), C.c_uint64(slot_bytes))

    def chain_ptr(self, which='min'):
        return lib().chord_hzb_chain_ptr(self.h, C.c_int(0 if which == 'min' else 1))

    @property
    def chain_bytes(self):
        return int(lib().chord_hzb_chain_bytes(self.h))

    def mip_offset_bytes(self, mip):
        return int(lib().chord_hzb_mip_offset(self.h, C.c_uint32(mip)))


class Frame:
    """GBufferTextures + per-view uniforms of one render dimension."""

    def __init__(self, ctx, width, height, _borrowed=None):
        self.ctx, self.width, self.height = ctx, int(width), int(height)
        self.h = C.c_void_p()
        self._borrowed = _borrowed is not None
        if self._borrowed:
            self.h = C.c_void_p(_borrowed)                # owned by a Renderer
        else:
            lib().chord_frame_create(ctx.h, C.c_uint32(width), C.c_uint32(height), C.byref(self.h))
        self._keep = None

    def close(self):
        if self._borrowed:
            self.h = C.c_void_p()
        if self.h:
            lib().chord_frame_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def begin(self, view, iview, objects, lod0_meshlet_count, meshlet_group_count, tile=None):
        view = np.ascontiguousarray(view, dtype=abi.PerframeCameraView).reshape(())
        iview = np.ascontiguousarray(iview, dtype=abi.InstanceCullingViewInfo).reshape(())
        objects = np.ascontiguousarray(objects, dtype=abi.GPUObjectGLTFPrimitive)
        self._keep = (view, iview, objects)       # async copies read these
        t = C.byref(abi.Rect(*tile)) if tile is not None else None
        lib().chord_frame_begin(self.h, _p(view), _p(iview), _p(objects), C.c_uint32(len(objects)),
                                C.c_uint32(lod0_meshlet_count), C.c_uint32(meshlet_group_count), t)

    def instance_culling(self, out: CmdList):
        lib().chord_instance_culling(self.h, out.h)

    def instance_culling_bvh(self, out: CmdList):
        lib().chord_instance_culling_bvh(self.h, out.h)

    def bvh_candidates(self, total_groups):
        words = (total_groups + 31) // 32 + 1
        bits = np.zeros(words, dtype=np.uint32)
        ctrl = np.zeros(5, dtype=np.uint32)
        lib().chord_readback_bvh_candidates(self.h, _p(bits), C.c_uint32(words), _p(ctrl))
        return bits, ctrl

    def hzb_culling(self, hzb: Hzb, first_stage, inp: CmdList, out_visible: CmdList, out_rejected: CmdList = None):
        lib().chord_hzb_culling(self.h, hzb.h, C.c_int(1 if first_stage else 0), inp.h, out_visible.h,
                                out_rejected.h if out_rejected is not None else None)

    def begin_cascade(self, view, cascade_views, cascade_id, objects, lod0_meshlet_count, meshlet_group_count):
        """chord_frame_begin with the instance view taken from the device cascade buffer."""
        v = np.ascontiguousarray(view, dtype=abi.PerframeCameraView).reshape(())
        o = np.ascontiguousarray(objects, dtype=abi.GPUObjectGLTFPrimitive)
        self._keep = (v, o)
        lib().chord_frame_begin_cascade(self.h, _p(v), cascade_views.h, C.c_uint32(cascade_id), _p(o), C.c_uint32(len(o)),
                                        C.c_uint32(lod0_meshlet_count), C.c_uint32(meshlet_group_count))

    def hzb_culling_generic_cascade(self, hzb: Hzb, cascade_views, hzb_view_id, extent_scale, use_last_frame, inp: CmdList, out_visible: CmdList):
        lib().chord_hzb_culling_generic_cascade(self.h, hzb.h, cascade_views.h, C.c_uint32(hzb_view_id), C.c_float(extent_scale),
                                                C.c_int(1 if use_last_frame else 0), inp.h, out_visible.h)

    def render_mesh_depth(self, lst: CmdList, pass_flags=abi.PASS_DEPTH | abi.PASS_DEPTH_CLAMP):
        lib().chord_render_mesh_depth(self.h, lst.h, C.c_uint32(pass_flags))

    def hzb_culling_generic(self, hzb: Hzb, hzb_view, extent_scale, use_last_frame, inp: CmdList, out_visible: CmdList):
        hv = np.ascontiguousarray(hzb_view, dtype=abi.InstanceCullingViewInfo).reshape(())
        lib().chord_hzb_culling_generic(self.h, hzb.h, _p(hv), C.c_float(extent_scale), C.c_int(1 if use_last_frame else 0), inp.h, out_visible.h)

    def render_mesh(self, lst: CmdList):
        lib().chord_render_mesh(self.h, lst.h)

    def visibility_stage0(self, last_hzb, inp, scratch_visible, out_rejected):
        need = C.c_int(0)
        lib().chord_visibility_stage0(self.h, last_hzb.h if last_hzb is not None else None, inp.h,
                                      scratch_visible.h if scratch_visible is not None else None,
                                      out_rejected.h if out_rejected is not None else None, C.byref(need))
        return bool(need.value)

    def visibility_stage1(self, hzb, rejected, scratch_visible):
        lib().chord_visibility_stage1(self.h, hzb.h, rejected.h, scratch_visible.h)

    def build_hzb(self, dst: Hzb, b_min=True, b_max=False, b_range=False):
        lib().chord_build_hzb(self.h, C.c_int(b_min), C.c_int(b_max), C.c_int(b_range), dst.h)

    def hzb_finish_top(self, dst: Hzb, b_min=True, b_max=False):
        lib().chord_hzb_finish_top(self.h, C.c_int(b_min), C.c_int(b_max), dst.h)

    def visibility_mark(self, post: CmdList):
        lib().chord_visibility_mark(self.h, post.h)

    def lighting(self, post: CmdList):
        lib().chord_lighting(self.h, post.h)

    # read-backs
    def vis64(self):
        out = np.zeros((self.height, self.width), dtype=np.uint64)
        lib().chord_readback_vis64(self.h, _p(out))
        return out

    def readback_vis64_into(self, host_ptr):
        lib().chord_readback_vis64(self.h, C.c_void_p(host_ptr))

    def visibility32(self, out=None):
        if out is None:
            out = np.zeros((self.height, self.width), dtype=np.uint32)
        lib().chord_readback_visibility32(self.h, _p(out))
        return out

    def depth32(self):
        out = np.zeros((self.height, self.width), dtype=np.float32)
        lib().chord_readback_depth32(self.h, _p(out))
        return out

    def decoded(self, post: CmdList):
        out = np.zeros((self.height, self.width, 3), dtype=np.uint32)
        lib().chord_readback_decoded(self.h, post.h, _p(out))
        return out

    def tile_marker(self):
        out = np.zeros(((self.height + 7) // 8, (self.width + 7) // 8, 4), dtype=np.uint32)
        lib().chord_readback_tile_marker(self.h, _p(out))
        return out

    def gbuffer(self, which):
        out = np.zeros((self.height, self.width, 4), dtype=np.float32)
        lib().chord_readback_gbuffer_f32(self.h, C.c_int(which), _p(out))
        return out

    def set_queue_limits(self, deferred=0, large_tris=0, large_items=0):
        lib().chord_frame_set_queue_limits(self.h, C.c_uint32(deferred), C.c_uint32(large_tris), C.c_uint32(large_items))

    def set_gbuffer_packed(self, packed):
        lib().chord_frame_set_gbuffer_packed(self.h, C.c_int(1 if packed else 0))

    def gbuffer_packed(self, which):
        """Reference storage formats: which 0 -> (H, W, 4) float16; 1..3 -> (H, W) uint32 A2B10G10R10; 4 -> (H, W, 2) float16; 5 -> (H, W, 4) uint8."""
        shapes = {0: ((self.height, self.width, 4), np.float16), 4: ((self.height, self.width, 2), np.float16),
                  5: ((self.height, self.width, 4), np.uint8)}
        shape, dt = shapes.get(which, ((self.height, self.width), np.uint32))
        out = np.zeros(shape, dtype=dt)
        lib().chord_readback_gbuffer_packed(self.h, C.c_int(which), _p(out))
        return out

    def raster_stats(self):
        s = abi.RasterStats()
        lib().chord_frame_raster_stats(self.h, C.byref(s))
        return s.as_dict()

    def reset_stats(self):
        lib().chord_frame_reset_stats(self.h)

    @property
    def vis64_ptr(self):
        return lib().chord_frame_vis64_ptr(self.h)


RENDERER_RESOLVE, RENDERER_PACKED_GBUFFER, RENDERER_BVH, RENDERER_GRAPH, RENDERER_TIMERS, RENDERER_ASYNC_RESOLVE = 1, 2, 4, 8, 16, 32
PASS_LABELS = ['GLTF Instance Culling', 'GLTF Visibility Stage0', 'BuildHZB Post Prepass Stage0', 'GLTF Visibility Stage1',
               'BuildHZB', 'VisibilityTileMarker', 'lightingTile']           # renderer.cpp:322-363 timer labels


class NcclComm:
    """chord_nccl_comm: an NCCL communicator created inside libchord_b200.so (libnccl.so.2 resolved with dlopen)."""

    def __init__(self, ctx, world, rank, unique_id):
        self.ctx, self.world, self.rank = ctx, world, rank
        self.h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id))
        lib().chord_nccl_comm_create(ctx.h, C.c_uint32(world), C.c_uint32(rank), buf, C.byref(self.h))

    @staticmethod
    def unique_id():
        buf = (C.c_ubyte * 128)()
        lib().chord_nccl_unique_id(buf)
        return bytes(buf)

    def close(self):
        if self.h:
            lib().chord_nccl_comm_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Renderer:
    """chord_renderer: DeferredRenderer::render (renderer.cpp:311-363 + :489) as one native call per frame, optionally recorded
    as CUDA graphs and replayed, optionally one rank of a tile-parallel frame (fused peer-store or NCCL exchange of the
    history HZB)."""

    def __init__(self, ctx, width, height, resolve=False, packed_gbuffer=False, bvh=False, graph=False, timers=False, async_resolve=False):
        self.ctx, self.width, self.height = ctx, int(width), int(height)
        flags = (RENDERER_RESOLVE if resolve else 0) | (RENDERER_PACKED_GBUFFER if packed_gbuffer else 0) | (RENDERER_BVH if bvh else 0) \
            | (RENDERER_GRAPH if graph else 0) | (RENDERER_TIMERS if timers else 0) | (RENDERER_ASYNC_RESOLVE if async_resolve else 0)
        self.h = C.c_void_p()
        lib().chord_renderer_create(ctx.h, C.c_uint32(width), C.c_uint32(height), C.c_uint32(flags), C.byref(self.h))
        self._frames = {}
        self.tile = (0, 0, self.width, self.height)
        self.rank, self.world = 0, 1
        self._keep = None
        self._lod0 = 0

    @property
    def frame(self):
        """The frame of the most recent render() (with async_resolve: one of two; the context stream is ordered after its resolve)."""
        p = lib().chord_renderer_frame(self.h)
        if p not in self._frames:
            self._frames[p] = Frame(self.ctx, self.width, self.height, _borrowed=p)
        return self._frames[p]

    def close(self):
        if self.h:
            for f in self._frames.values():
                f.close()
            self._frames = {}
            lib().chord_renderer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tile(self, tile):
        if tile is None:
            lib().chord_renderer_set_tile(self.h, None)
            self.tile = (0, 0, self.width, self.height)
        else:
            lib().chord_renderer_set_tile(self.h, C.byref(abi.Rect(*tile)))
            self.tile = tuple(tile)

    def render(self, view, iview, objects, lod0_meshlet_count, meshlet_group_count):
        view = np.ascontiguousarray(view, dtype=abi.PerframeCameraView).reshape(())
        iview = np.ascontiguousarray(iview, dtype=abi.InstanceCullingViewInfo).reshape(())
        objects = np.ascontiguousarray(objects, dtype=abi.GPUObjectGLTFPrimitive)
        self._keep = (view, iview, objects)
        self._lod0 = int(lod0_meshlet_count)
        lib().chord_renderer_render(self.h, _p(view), _p(iview), _p(objects), C.c_uint32(len(objects)),
                                    C.c_uint32(lod0_meshlet_count), C.c_uint32(meshlet_group_count))

    def clear_history(self):
        lib().chord_renderer_clear_history(self.h)

    @property
    def post_cull(self):
        p = lib().chord_renderer_post_cull(self.h)
        return CmdList(self.ctx, lib().chord_cmdlist_capacity(C.c_void_p(p)), _borrowed=p) if p else None

    @property
    def history_hzb(self):
        p = lib().chord_renderer_history_hzb(self.h)
        return Hzb(self.ctx, self.width, self.height, _borrowed=p) if p else None

    @property
    def graph_replays(self):
        return int(lib().chord_renderer_graph_replays(self.h))

    def pass_times(self):
        """{reference timer label: GPU microseconds} of the last frame (needs timers=True; synchronises)."""
        out = (C.c_float * 7)()
        lib().chord_renderer_pass_times(self.h, out)
        return dict(zip(PASS_LABELS, [float(x) for x in out]))

    # ---- multi-GPU exchange
    def exchange_export(self):
        """-> (64-byte CUDA IPC handle, device pointer) of this rank's exchange block."""
        h = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        lib().chord_renderer_exchange_export(self.h, h, C.byref(ptr))
        return bytes(h), int(ptr.value)

    def exchange_connect(self, rank, world, tiles, handles=None, local_ptrs=None):
        arr = (abi.Rect * world)(*[abi.Rect(*t) for t in tiles])
        hs = None
        if handles is not None:
            hs = (C.c_ubyte * (64 * world))()
            for i, b in enumerate(handles):
                hs[64 * i:64 * i + 64] = list(b)
        lp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in local_ptrs]) if local_ptrs is not None else None
        lib().chord_renderer_exchange_connect(self.h, C.c_uint32(rank), C.c_uint32(world), arr, hs, lp)
        self.rank, self.world, self.tile = rank, world, tuple(tiles[rank])

    def exchange_connect_nccl(self, tiles, comm: NcclComm):
        arr = (abi.Rect * comm.world)(*[abi.Rect(*t) for t in tiles])
        lib().chord_renderer_exchange_connect_nccl(self.h, arr, comm.h)
        self.rank, self.world, self.tile = comm.rank, comm.world, tuple(tiles[comm.rank])
        self._comm = comm

    def history_sync(self):
        lib().chord_renderer_history_sync(self.h)

    def set_exchange_timeout(self, max_polls):
        """Bound of the wait for the peers' arrival flags: max_polls x 200 ns (0 = default, ~4 s)."""
        lib().chord_renderer_set_exchange_timeout(self.h, C.c_uint32(int(max_polls)))

    def exchange_timed_out(self):
        v = C.c_uint32()
        lib().chord_renderer_exchange_status(self.h, C.byref(v))
        return bool(v.value)
