"""Screen-tile partition of one frame across the GPUs of one NVSwitch box.

The reference is single-GPU; this is the B200-native scaling layer around the same pass graph:

* the frame is cut into N tiles on multiples of 64 px (one HZB mip-5 texel = one 64x64 block = one owner);
* the scene is replicated; every rank culls against its tile's sub-frustum (the side planes of
  InstanceCullingViewInfo.frustumPlanesRS are replaced, K2 does the test unchanged) and rasterises with the tile
  as scissor, so its work is proportional to the meshlets touching the tile;
* phase 0 of the two-phase occlusion cull needs last frame's FULL min-HZB on every rank: after the final buildHZB
  each rank packs its tile's part of mips 0..5, ONE all-gather (NCCL over NVLink; gloo in the CPU tests) exchanges
  the parts, every rank unpacks them and rebuilds mips 6.. locally (chord_hzb_finish_top).  No mid-frame collective:
  the phase-1 HZB is tile-local (texels outside the tile read 0 = far = never occludes);
* resolve is per-pixel and stays tile-local.

Everything here is host logic; pack / unpack / finish run as CUDA kernels in libchord_b200.so.  The numpy versions
(`pack_tile_np` ...) restate the packed layout for the gloo tests and for documentation.
"""
import numpy as np

from . import abi, camera as cam

LAYOUTS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}


def tile_for_rank(rank, n, width, height):
    """Tile rect (x0, y0, x1, y1) of ``rank``; 2 -> halves, 4 -> 2x2, 8 -> 4x2; inner cuts on multiples of 64 px."""
    cols, rows = LAYOUTS[n]
    cx, cy = rank % cols, rank // cols

    def cut(i, k, dim):
        return dim if i >= k else min(dim, (dim * i // k + 63) // 64 * 64)
    return (cut(cx, cols, width), cut(cy, rows, height), cut(cx + 1, cols, width), cut(cy + 1, rows, height))


def all_tiles(n, width, height):
    return [tile_for_rank(r, n, width, height) for r in range(n)]


def tile_instance_view(view, camera, tile, margin_px=1.0):
    """InstanceCullingViewInfo whose four side planes bound the tile's sub-frustum (camera.cpp:80-154 construction with
    the tile's NDC extent instead of [-1, 1]); near/far planes and the clip matrix are the main view's."""
    iv = cam.fill_instance_view(view)
    W, H = float(view['renderDimension'][0]), float(view['renderDimension'][1])
    x0, y0, x1, y1 = tile
    nx0 = 2.0 * (x0 - margin_px) / W - 1.0
    nx1 = 2.0 * (x1 + margin_px) / W - 1.0
    ny_top = 1.0 - 2.0 * (y0 - margin_px) / H
    ny_bot = 1.0 - 2.0 * (y1 + margin_px) / H
    F = np.float32
    fwd, up, right = camera.front.astype(np.float64), camera.up.astype(np.float64), camera.right.astype(np.float64)
    t = np.tan(camera.fovy * 0.5)
    a = float(camera.aspect)

    def corner(z, nx, ny):
        return fwd * z + right * (nx * z * t * a) + up * (ny * z * t)
    zn, zf = camera.z_near, camera.z_far
    nlt, nld, nrt, nrd = corner(zn, nx0, ny_top), corner(zn, nx0, ny_bot), corner(zn, nx1, ny_top), corner(zn, nx1, ny_bot)
    flt, fld, frt, frd = corner(zf, nx0, ny_top), corner(zf, nx0, ny_bot), corner(zf, nx1, ny_top), corner(zf, nx1, ny_bot)

    def plane(pc, p1, p2):
        n = np.cross(p1 - pc, p2 - pc)
        n = n / np.linalg.norm(n)
        return np.array([n[0], n[1], n[2], -np.dot(n, pc)], dtype=F)
    planes = iv['frustumPlanesRS'].copy()
    planes[0] = plane(fld, flt, nld)      # left
    planes[1] = plane(frd, fld, nrd)      # down
    planes[2] = plane(frt, frd, nrt)      # right
    planes[3] = plane(flt, frt, nlt)      # top
    iv['frustumPlanesRS'] = planes
    return iv


# ------------------------------------------------------------------------------------------------ packed HZB layout
def hzb_geometry(width, height):
    """hzb.cpp:49-56: (mip0W, mip0H, mips, texel offsets)."""
    def next_pot(v):
        p = 1
        while p < v:
            p <<= 1
        return p
    w0, h0 = next_pot(width) // 2, next_pot(height) // 2
    mips = int(np.floor(np.log2(max(w0, h0)))) + 1
    offs, o = [], 0
    for m in range(mips):
        offs.append(o)
        o += max(w0 >> m, 1) * max(h0 >> m, 1)
    return w0, h0, mips, offs, o


def tile_blocks(tile, width, height):
    """64x64-pixel block range owned by a tile; edge tiles also own the clamped off-screen blocks (hzb.cu tile_blocks)."""
    w0, h0, _, _, _ = hzb_geometry(width, height)
    gw, gh = (w0 + 31) // 32, (h0 + 31) // 32
    x0, y0, x1, y1 = tile
    return x0 // 64, y0 // 64, (gw if x1 >= width else x1 // 64), (gh if y1 >= height else y1 // 64)


def tile_texels(tile, width, height):
    bx0, by0, bx1, by1 = tile_blocks(tile, width, height)
    return sum((((bx1 - bx0) * 32) >> m) * (((by1 - by0) * 32) >> m) for m in range(6))


def pack_tile_np(chain, tile, width, height):
    """mips 0..5 of the tile's region, row-major per mip, mips concatenated (k_hzb_tile_copy<true>)."""
    w0, h0, mips, offs, _ = hzb_geometry(width, height)
    bx0, by0, bx1, by1 = tile_blocks(tile, width, height)
    parts = []
    for m in range(6):
        mw, mh = w0 >> m, h0 >> m
        img = chain[offs[m]:offs[m] + mw * mh].reshape(mh, mw)
        parts.append(img[(by0 * 32) >> m:(by1 * 32) >> m, (bx0 * 32) >> m:(bx1 * 32) >> m].reshape(-1))
    return np.concatenate(parts)


def unpack_tile_np(chain, tile, packed, width, height):
    w0, h0, mips, offs, _ = hzb_geometry(width, height)
    bx0, by0, bx1, by1 = tile_blocks(tile, width, height)
    k = 0
    for m in range(6):
        mw, mh = w0 >> m, h0 >> m
        img = chain[offs[m]:offs[m] + mw * mh].reshape(mh, mw)
        rw, rh = ((bx1 - bx0) * 32) >> m, ((by1 - by0) * 32) >> m
        img[(by0 * 32) >> m:(by1 * 32) >> m, (bx0 * 32) >> m:(bx1 * 32) >> m] = packed[k:k + rw * rh].reshape(rh, rw)
        k += rw * rh


def finish_top_np(chain, width, height, reduce=np.minimum):
    """mips 6.. from mip 5 on fp16 bit patterns of non-negative values (monotone, so integer min/max is exact)."""
    w0, h0, mips, offs, _ = hzb_geometry(width, height)
    for m in range(6, mips):
        pw, ph = max(w0 >> (m - 1), 1), max(h0 >> (m - 1), 1)
        src = chain[offs[m - 1]:offs[m - 1] + pw * ph].reshape(ph, pw)
        cols = reduce(src[:, 0::2], src[:, 1::2]) if pw > 1 else src
        r = reduce(cols[0::2, :], cols[1::2, :]) if ph > 1 else cols
        chain[offs[m]:offs[m] + r.size] = r.reshape(-1)


# ------------------------------------------------------------------------------------------------ device all-gather
class _DevPtr:
    """Wraps a raw device pointer so torch can view it (``__cuda_array_interface__``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (int(ptr), False), 'version': 2}


def allgather_hzb(frame, hzb, rank, world, dist, which=('min', 'max'), staging=None):
    """Exchange the tiles' parts of mips 0..5 of ``hzb`` between all ranks and rebuild mips 6.. everywhere (for callers that
    sequence the passes themselves through torch.distributed; the renderer-level paths are ``connect_fused`` / ``connect_nccl``).
    The tiles' valid depth ranges are NOT reduced here: a chain gathered this way must not be used as ``validRangeSource`` of
    chord_cascade_setup (the fused exchange of chord_renderer does reduce them).

    ``dist`` is torch.distributed (NCCL); the collective is enqueued on the library's stream via an ExternalStream so
    that pack -> all-gather -> unpack -> finish_top is stream-ordered without host synchronisation."""
    import torch
    if staging is None:
        tiles = all_tiles(world, frame.width, frame.height)
        slot = max(hzb.tile_bytes(t) for t in tiles)
        slot = (slot + 255) // 256 * 256
        dev = torch.device('cuda', frame.ctx.device)
        staging = dict(tiles=tiles, slot=slot, stream=torch.cuda.ExternalStream(frame.ctx.stream, device=dev),
                       send=torch.zeros(slot, dtype=torch.uint8, device=dev), recv=torch.zeros(slot * world, dtype=torch.uint8, device=dev))
    tiles, slot, stream, send, recv = staging['tiles'], staging['slot'], staging['stream'], staging['send'], staging['recv']
    with torch.cuda.stream(stream):
        for w in which:
            hzb.pack_tile(w, tiles[rank], send.data_ptr())
            dist.all_gather_into_tensor(recv, send)
            hzb.unpack_tiles(w, tiles, recv.data_ptr(), slot)
        frame.hzb_finish_top(hzb, 'min' in which, 'max' in which)
    hzb.set_valid(True)
    return staging


def connect_fused(renderer, rank, world, dist, tiles=None):
    """Fused exchange between processes: every rank exports its exchange block (CUDA IPC), the 64-byte handles travel through
    torch.distributed once, chord_renderer_exchange_connect maps the peers.  After this the frames need no collective call."""
    tiles = tiles or all_tiles(world, renderer.width, renderer.height)
    handle, _ = renderer.exchange_export()
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    renderer.exchange_connect(rank, world, tiles, handles=handles)
    dist.barrier()                       # nobody renders before every rank has mapped every peer
    return tiles


def connect_nccl(renderer, rank, world, dist, tiles=None):
    """NCCL exchange: a communicator created inside libchord_b200.so (rank 0's unique id is broadcast once)."""
    from .runtime import NcclComm
    tiles = tiles or all_tiles(world, renderer.width, renderer.height)
    box = [NcclComm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = NcclComm(renderer.ctx, world, rank, box[0])
    renderer.exchange_connect_nccl(tiles, comm)
    dist.barrier()
    return tiles


class TileRenderer:
    """One rank of a tile-parallel frame: the native renderer (chord_renderer, one C-ABI call per frame) with its history HZB
    exchanged by the fused peer-store path (``exchange='fused'``, default) or by the NCCL all-gather inside the library
    (``exchange='nccl'``).  Both leave min chain, max chain and valid range of the history identical to a single-GPU frame on
    every rank (round 1 gathered the min chain only and left max / valid range rank-local: ADVICE r1)."""

    def __init__(self, ctx, width, height, rank, world, dist, exchange='fused', resolve=False, packed_gbuffer=False, graph=False):
        from .runtime import Renderer
        self.r = Renderer(ctx, width, height, resolve=resolve, packed_gbuffer=packed_gbuffer, graph=graph)
        self.rank, self.world, self.dist = rank, world, dist
        self.tile = tile_for_rank(rank, world, width, height)
        if world > 1:
            (connect_fused if exchange == 'fused' else connect_nccl)(self.r, rank, world, dist)

    def render(self, camera, objects, lod0, groups, view):
        iview = tile_instance_view(view, camera, self.tile) if self.world > 1 else cam.fill_instance_view(view)
        self.r.render(view, iview, objects, lod0, groups)
        return self.r.post_cull
