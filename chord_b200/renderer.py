"""Host mirror of the reference's pass graph for the visibility-buffer path.

Same free functions, same argument meaning and same sequencing as source/renderer (gltf_rendering.h:37-108,
postprocessing.h:41-54, visibility_tile.h:22-34, lighting.h:37, renderer.cpp:311-363 + history at :489), with
``GraphicsQueue&`` -> the context stream and pooled buffers/textures -> objects from a small pool.  Every pass is
one call into libchord_b200.so; nothing is computed on the host.
"""
from dataclasses import dataclass, field

import numpy as np

from . import abi, camera as cam
from .runtime import CascadeViews, CmdList, Context, Frame, Hzb


class _Pool:
    """getContext().getBufferPool()/getTexturePool() stand-in: reuse lists / HZBs of the same shape across frames."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.free_lists, self.free_hzb = {}, {}

    def cmdlist(self, capacity):
        b = self.free_lists.get(capacity)
        return b.pop() if b else CmdList(self.ctx, capacity)

    def hzb(self, w, h):
        b = self.free_hzb.get((w, h))
        return b.pop() if b else Hzb(self.ctx, w, h)

    def release(self, obj):
        if obj is None:
            return
        if isinstance(obj, CmdList):
            self.free_lists.setdefault(obj.capacity, []).append(obj)
        else:
            self.free_hzb.setdefault((obj.width, obj.height), []).append(obj)


@dataclass
class GLTFRenderContext:
    """gltf_rendering.h:15-35."""
    frame: Frame
    gltfObjectCount: int
    lod0MeshletCount: int
    meshletGroupCount: int
    pool: _Pool
    timers: list = field(default_factory=list)


def shouldRenderGLTF(ctx: GLTFRenderContext):
    """mesh_raster.cpp:45-48."""
    return ctx.gltfObjectCount != 0


def instanceCulling(ctx: GLTFRenderContext, use_bvh=False):
    """instance_culling.cpp:83-161 -> CountAndCmdBuffer sized by lod0MeshletCount (:140-141).  ``use_bvh`` walks the
    cluster-group BVH first (same list; the reference builds the BVH but loops flat)."""
    if not shouldRenderGLTF(ctx):
        return None
    out = ctx.pool.cmdlist(ctx.lod0MeshletCount)
    if use_bvh:
        ctx.frame.instance_culling_bvh(out)
    else:
        ctx.frame.instance_culling(out)
    return out


def gltfVisibilityRenderingStage0(ctx: GLTFRenderContext, hzb_last, in_list):
    """mesh_raster.cpp:269-312. Returns (bShouldInvokeStage1, rejected CountAndCmdBuffer)."""
    if not shouldRenderGLTF(ctx):
        return False, None
    if hzb_last is not None and hzb_last.valid and (ctx.frame.ctx.switch_flags & abi.FLAG_HZB):
        visible = ctx.pool.cmdlist(ctx.lod0MeshletCount)
        rejected = ctx.pool.cmdlist(ctx.lod0MeshletCount)
        need = ctx.frame.visibility_stage0(hzb_last, in_list, visible, rejected)
        ctx.pool.release(visible)        # stream-ordered: the next user is enqueued after the raster that reads it
        return need, rejected
    ctx.frame.visibility_stage0(None, in_list, None, None)
    return False, None


def gltfVisibilityRenderingStage1(ctx: GLTFRenderContext, hzb_this, rejected):
    """mesh_raster.cpp:314-329."""
    visible = ctx.pool.cmdlist(ctx.lod0MeshletCount)
    ctx.frame.visibility_stage1(hzb_this, rejected, visible)
    ctx.pool.release(visible)


def buildHZB(ctx: GLTFRenderContext, bBuildMin, bBuildMax, bBuildValidRange):
    """hzb.cpp:38-227 -> HZBContext."""
    h = ctx.pool.hzb(ctx.frame.width, ctx.frame.height)
    ctx.frame.build_hzb(h, bBuildMin, bBuildMax, bBuildValidRange)
    return h


def visibilityMark(ctx: GLTFRenderContext, drawMeshletCmdBuffer):
    """visibility_tile.cpp:20-56."""
    ctx.frame.visibility_mark(drawMeshletCmdBuffer)


def lighting(ctx: GLTFRenderContext, drawMeshletCmdBuffer):
    """lighting.cpp:193-246."""
    ctx.frame.lighting(drawMeshletCmdBuffer)


class DeferredRenderer:
    """renderer.cpp:142-499 restricted to the visibility-buffer path, headless; keeps last frame's HZB as history."""

    def __init__(self, ctx: Context, width, height, do_resolve=False, packed_gbuffer=False, use_bvh=False):
        self.ctx, self.width, self.height = ctx, width, height
        self.frame = Frame(ctx, width, height)
        self.frame.set_gbuffer_packed(packed_gbuffer)
        self.use_bvh = use_bvh
        self.pool = _Pool(ctx)
        self.history_hzb = None            # m_rendererHistory.hzbCtx (renderer.cpp:326, :489)
        self.last_view = None
        self.post_cull = None
        self.do_resolve = do_resolve
        self.frame_counter = 0

    def clear_history(self):
        """renderer.cpp:95-105 (resize / camera cut)."""
        self.pool.release(self.history_hzb)
        self.history_hzb = None
        self.last_view = None

    def render(self, camera, objects, lod0_meshlet_count, meshlet_group_count, tile=None, view=None, iview=None):
        if view is None:
            view = cam.fill_view(camera, last_view=self.last_view, frame_counter=self.frame_counter)
        if iview is None:
            iview = cam.fill_instance_view(view)
        f = self.frame
        f.begin(view, iview, objects, lod0_meshlet_count, meshlet_group_count, tile)      # uniforms + addClearGbufferPass
        rc = GLTFRenderContext(f, len(objects), lod0_meshlet_count, meshlet_group_count, self.pool)
        self.pool.release(self.post_cull)
        post = instanceCulling(rc, self.use_bvh)
        need1, rejected = gltfVisibilityRenderingStage0(rc, self.history_hzb, post)
        if need1:
            tmp = buildHZB(rc, True, False, False)
            gltfVisibilityRenderingStage1(rc, tmp, rejected)
            self.pool.release(tmp)
        self.pool.release(rejected)
        hzb = buildHZB(rc, True, True, True)
        if self.do_resolve:
            visibilityMark(rc, post)
            lighting(rc, post)
        self.pool.release(self.history_hzb)
        self.history_hzb = hzb
        self.last_view = view
        self.post_cull = post
        self.frame_counter += 1
        return post


# ----------------------------------------------------------------------------------------------- cascade shadow depth (N3)
kShadowHZBCullingEnable = True              # r.shadow.hzbCulling (mesh_raster.cpp:28-33)
kShadowExtentScaleForHZBCulling = 1.5       # mesh_raster.cpp:35-40


@dataclass
class CascadeShadowContext:
    """render_helper.h CascadeShadowContext / CascadeShadowHistory: per-cascade depth targets (chord_frames of
    cascadeDim^2 whose packed surface holds the D32 depth in its high word), the cascade views, the light direction."""
    config: object
    direction: np.ndarray
    views: np.ndarray
    depths: list
    lists: list = field(default_factory=list)      # post-cull lists per cascade (kept for inspection / tests)
    device_views: object = None                    # CascadeViews when the setup ran on the GPU (chord_cascade_setup)


def hzbCullingGeneric(frame: Frame, pool: _Pool, hzb, extent_scale, hzb_view, use_last_frame, in_list):
    """detail::hzbCullingGeneric (hzb_culling_generic.hlsl:37-172) -> new CountAndCmdBuffer."""
    out = pool.cmdlist(in_list.capacity)
    frame.hzb_culling_generic(hzb, hzb_view, extent_scale, use_last_frame, in_list, out)
    return out


def renderShadow(ctx: Context, pool: _Pool, view, objects, lod0_meshlet_count, meshlet_group_count, history, config, light_direction,
                 valid_depth_range=None, depth_targets=None, main_frame=None, valid_range_hzb=None):
    """chord::renderShadow (mesh_raster.cpp:332-546): cascade views (cascade_setup.hlsl, host mirror), then for every
    cascade that is not cache-valid, far to near: instanceCulling with the cascade view -> hzbCullingGeneric against the
    previous (farther) cascade's HZB, or against the cached depth of the same cascade when there is none -> clear ->
    renderMeshDepth (two-sided, depth clamp) -> buildHZB(min) for the next cascade.  ``history`` is the
    CascadeShadowContext of the last frame or None."""
    cfg = config
    direction = np.asarray(light_direction, dtype=np.float32)
    cache_valid = (history is not None and len(history.depths) == cfg.cascadeCount
                   and np.array_equal(history.direction, direction) and history.config is cfg)
    tick = int(view['basicData']['frameCounter'])
    dviews = None
    if main_frame is not None:
        # CascadeSetupCS on the device (chord_cascade_setup): reads the main frame's PerframeCameraView and, for SDSM, the
        # valid depth range of ``valid_range_hzb`` straight from HBM; nothing is read back, the whole pass is asynchronous
        dviews = history.device_views if (cache_valid and history.device_views is not None) else CascadeViews(ctx)
        dviews.setup(main_frame, cfg, direction, tick, cache_valid, valid_range_hzb)
        views = None
    else:
        views = cam.cascade_views(view, direction, cfg, valid_depth_range, history.views if cache_valid else None, cache_valid)
    if cache_valid:
        depths = list(history.depths)
    elif depth_targets is not None:              # pooled depth textures (getTexturePool().create, mesh_raster.cpp:381-389)
        depths = list(depth_targets)
    else:
        depths = [Frame(ctx, cfg.cascadeDim, cfg.cascadeDim) for _ in range(cfg.cascadeCount)]
    out = CascadeShadowContext(cfg, direction, views, depths, [None] * cfg.cascadeCount, dviews)

    def begin(f, cid):
        if dviews is not None:
            f.begin_cascade(view, dviews, cid, objects, lod0_meshlet_count, meshlet_group_count)
        else:
            f.begin(view, views[cid], objects, lod0_meshlet_count, meshlet_group_count)

    def cull(f, hzb, view_id, in_list):
        o = pool.cmdlist(in_list.capacity)
        if dviews is not None:
            f.hzb_culling_generic_cascade(hzb, dviews, view_id, kShadowExtentScaleForHZBCulling, False, in_list, o)
        else:
            f.hzb_culling_generic(hzb, views[view_id], kShadowExtentScaleForHZBCulling, False, in_list, o)
        return o

    prev_hzb, prev_id = None, None
    for cid in range(cfg.cascadeCount - 1, -1, -1):
        if cam.is_cascade_cache_valid(cache_valid, cfg.realtimeCascadeCount, cfg.cascadeCount, tick, cid):
            continue
        f = depths[cid]
        cached_hzb = None
        if prev_hzb is None and cache_valid and kShadowHZBCullingEnable:
            # the cascade's depth of last frame is still in f: build its HZB before the target is cleared
            cached_hzb = pool.hzb(cfg.cascadeDim, cfg.cascadeDim)
            f.build_hzb(cached_hzb, True, False, False)
        begin(f, cid)                                                                      # uniforms + clearDepthStencil(0.0)
        post = pool.cmdlist(lod0_meshlet_count)
        f.instance_culling(post)
        if prev_hzb is None:
            if cached_hzb is not None:
                # mesh_raster.cpp:463-479 passes cascadeHistory.viewsSRV, which is the SAME buffer CascadeSetupCS has just
                # rewritten for this cascade (mesh_raster.cpp:411-414): the cached depth is tested with the NEW view
                culled = cull(f, cached_hzb, cid, post)
                pool.release(post); post = culled
                pool.release(cached_hzb)
        elif kShadowHZBCullingEnable:
            culled = cull(f, prev_hzb, prev_id, post)
            pool.release(post); post = culled
        f.render_mesh_depth(post, abi.PASS_DEPTH | abi.PASS_DEPTH_CLAMP)
        out.lists[cid] = post
        if cid != 0:
            pool.release(prev_hzb)
            prev_hzb = pool.hzb(cfg.cascadeDim, cfg.cascadeDim)
            f.build_hzb(prev_hzb, True, False, False)
            prev_id = cid
    pool.release(prev_hzb)
    return out
