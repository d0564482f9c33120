"""Pins the oracle's fixed-function rasteriser (PARITY UNPINNED by the reference, which delegates this to Vulkan
hardware) against an independent exact-arithmetic statement of the Vulkan rules: pixel-centre sampling, 8 sub-pixel
bits, top-left fill rule, CCW-front with the flipped viewport, screen-space-linear depth."""
from fractions import Fraction

import numpy as np
import pytest

import oracle


def _snap(v):
    """round-half-even to 1/256 px, like lrintf(v * 256)."""
    return int(np.rint(np.float32(v) * np.float32(256.0)))


def exact_coverage(sx, sy, two_sided, w, h):
    """Independent implementation with Python ints: returns set of covered (px, py)."""
    X = [_snap(v) for v in sx]
    Y = [_snap(v) for v in sy]
    area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0])
    if area2 == 0 or (not two_sided and area2 > 0):
        return set()
    pts = list(zip(X, Y))
    if area2 < 0:                       # make it clockwise in y-down screen space
        pts = [pts[0], pts[2], pts[1]]
    cov = set()
    for py in range(h):
        for px in range(w):
            SX, SY = px * 256 + 128, py * 256 + 128
            inside = True
            for i in range(3):
                (ax, ay), (bx, by) = pts[i], pts[(i + 1) % 3]
                ex, ey = bx - ax, by - ay
                e = ex * (SY - ay) - ey * (SX - ax)
                top_left = (ey == 0 and ex > 0) or ey < 0
                if e < 0 or (e == 0 and not top_left):
                    inside = False
                    break
            if inside:
                cov.add((px, py))
    return cov


def _clip_from_screen(sx, sy, z, w, h, wclip=1.0):
    """clip coordinates whose viewport transform gives the requested pixel positions."""
    ndc_x = np.float32(sx) / np.float32(w / 2) - np.float32(1.0)
    ndc_y = np.float32(1.0) - np.float32(sy) / np.float32(h / 2)
    return np.stack([ndc_x * wclip, ndc_y * wclip, np.asarray(z, np.float32) * wclip, np.full(3, wclip, np.float32)], axis=1)


def _screen_from_clip(clip, w, h):
    """the oracle's own viewport transform restated in numpy float32 (fma emulated in float64: exact for these ranges)."""
    x = (clip[:, 0] / clip[:, 3]).astype(np.float32)
    y = (clip[:, 1] / clip[:, 3]).astype(np.float32)
    sx = np.float32(np.float64(x) * (w / 2) + (w / 2))
    sy = np.float32(np.float64(y) * (-h / 2) + (h / 2))
    return sx, sy


@pytest.mark.parametrize('seed', range(6))
def test_coverage_matches_exact_rule(seed):
    rng = np.random.default_rng(seed)
    w = h = 24
    for _ in range(150):
        kind = rng.integers(0, 4)
        if kind == 0:      # vertices on the 1/2-pixel lattice: many samples exactly on edges (fill-rule cases)
            sx, sy = rng.integers(0, 2 * w, 3) / 2.0, rng.integers(0, 2 * h, 3) / 2.0
        elif kind == 1:    # axis-aligned edges
            sx = np.array([rng.integers(0, w), rng.integers(0, w), rng.integers(0, w)]) + 0.5
            sy = np.array([rng.integers(0, h)] * 2 + [rng.integers(0, h)]) + 0.5
        elif kind == 2:    # tiny triangles
            c = rng.uniform(2, w - 2, 2)
            sx, sy = c[0] + rng.uniform(-1, 1, 3), c[1] + rng.uniform(-1, 1, 3)
        else:
            sx, sy = rng.uniform(-4, w + 4, 3), rng.uniform(-4, h + 4, 3)
        clip = _clip_from_screen(sx, sy, [0.5, 0.25, 0.75], w, h, wclip=rng.uniform(0.5, 4.0))
        rsx, rsy = _screen_from_clip(clip, w, h)
        for two_sided in (0, 1):
            vis = oracle.raster_clip_triangle(w, h, clip, 0x100, two_sided)
            got = set(zip(*np.nonzero(vis.T)))
            want = exact_coverage(rsx, rsy, two_sided, w, h)
            assert got == want, (seed, kind, two_sided, sx, sy)


def test_facing_convention():
    """CCW in y-up NDC is front facing (helper.h:400 + flipped viewport helper.h:304-323); cull mode BACK."""
    w = h = 16
    ccw = np.array([[-0.5, -0.5, 0.5, 1], [0.5, -0.5, 0.5, 1], [0.0, 0.5, 0.5, 1]], dtype=np.float32)
    cw = ccw[[0, 2, 1]]
    assert (oracle.raster_clip_triangle(w, h, ccw, 0x100, 0) != 0).sum() > 10
    assert (oracle.raster_clip_triangle(w, h, cw, 0x100, 0) != 0).sum() == 0
    a = oracle.raster_clip_triangle(w, h, cw, 0x100, 1)
    b = oracle.raster_clip_triangle(w, h, ccw, 0x100, 1)
    assert np.array_equal(a != 0, b != 0)


def test_watertight_fan():
    """A fan of triangles sharing edges covers every pixel exactly once (top-left rule)."""
    rng = np.random.default_rng(3)
    w = h = 40
    for _ in range(20):
        c = rng.uniform(10, 30, 2)
        ang = (np.arange(7) + rng.uniform(-0.35, 0.35, 7)) * (2 * np.pi / 7)      # every wedge < 180 degrees
        rad = rng.uniform(6, 18, 7)
        ring = np.stack([c[0] + rad * np.cos(ang), c[1] - rad * np.sin(ang)], axis=1)     # CCW seen with y up
        count = np.zeros((h, w), dtype=int)
        union = np.zeros((h, w), dtype=np.uint64)
        for i in range(7):
            p = np.array([c, ring[i], ring[(i + 1) % 7]])
            clip = _clip_from_screen(p[:, 0], p[:, 1], [0.5, 0.5, 0.5], w, h)
            v = oracle.raster_clip_triangle(w, h, clip, 0x100 + i, 1)
            count += (v != 0)
            oracle.raster_clip_triangle(w, h, clip, 0x100 + i, 1, union)
        assert count.max() <= 1, 'a pixel was covered by two triangles of a watertight fan'
        assert np.array_equal(union != 0, count == 1)


def test_depth_is_screen_linear_and_reversed_z_wins():
    w = h = 32
    sx, sy = np.array([2.0, 30.0, 2.0]), np.array([2.0, 2.0, 30.0])
    z = np.array([0.2, 0.6, 0.9], dtype=np.float32)
    clip = _clip_from_screen(sx, sy, z, w, h)
    vis = oracle.raster_clip_triangle(w, h, clip, 0x100, 1)
    depth = (vis >> np.uint64(32)).astype(np.uint32).view(np.float32)
    ys, xs = np.nonzero(vis)
    # exact plane through the snapped vertices
    for px, py in list(zip(xs, ys))[::17]:
        A = np.array([[sx[0], sy[0], 1], [sx[1], sy[1], 1], [sx[2], sy[2], 1]])
        coef = np.linalg.solve(A, z.astype(np.float64))
        want = coef @ np.array([px + 0.5, py + 0.5, 1.0])
        assert abs(depth[py, px] - want) < 2e-6
    # nearer (larger reversed-Z depth) wins regardless of submission order; equal depth -> larger id wins
    near = _clip_from_screen(sx, sy, z + 0.05, w, h)
    a = oracle.raster_clip_triangle(w, h, clip, 0x200, 1)
    oracle.raster_clip_triangle(w, h, near, 0x100, 1, a)
    b = oracle.raster_clip_triangle(w, h, near, 0x100, 1)
    oracle.raster_clip_triangle(w, h, clip, 0x200, 1, b)
    assert np.array_equal(a, b) and ((a & np.uint64(0xFFFFFFFF))[a != 0] == 0x100).all()
    t = oracle.raster_clip_triangle(w, h, clip, 0x100, 1)
    oracle.raster_clip_triangle(w, h, clip, 0x300, 1, t)
    assert ((t & np.uint64(0xFFFFFFFF))[t != 0] == 0x300).all()


def test_near_plane_clip_matches_unclipped_part():
    """A triangle crossing the near plane (z > w behind it) must cover exactly what its in-front part covers."""
    w = h = 64
    # view-space triangle in a reversed-Z infinite projection: clip = (x*f, y*f, zn, -zv)
    zn, f = 0.1, 1.0

    def proj(p):
        return np.array([p[0] * f, p[1] * f, zn, -p[2]], dtype=np.float32)
    tri = [(-0.5, -0.3, -1.0), (0.6, -0.3, -1.0), (0.0, 0.4, 0.5)]       # third vertex behind the camera
    clip = np.stack([proj(p) for p in tri])
    vis = oracle.raster_clip_triangle(w, h, clip, 0x100, 1)
    assert (vis != 0).sum() > 50
    depth = (vis >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert depth.max() <= 1.0 and depth[vis != 0].min() > 0.0
    # bottom edge (both endpoints in front) must be rasterised exactly as in an unclipped triangle sharing it
    tri2 = [(-0.5, -0.3, -1.0), (0.6, -0.3, -1.0), (0.0, 0.0, -0.9)]
    both = oracle.raster_clip_triangle(w, h, np.stack([proj(p) for p in tri2]), 0x100, 1)
    row = int(np.nonzero(both.any(axis=1))[0].max())
    assert np.array_equal(vis[row] != 0, both[row] != 0)


def test_scissor_and_idempotence():
    w = h = 32
    clip = _clip_from_screen([1.0, 31.0, 1.0], [1.0, 1.0, 31.0], [0.3, 0.4, 0.5], w, h)
    full = oracle.raster_clip_triangle(w, h, clip, 0x100, 1)
    part = oracle.raster_clip_triangle(w, h, clip, 0x100, 1, tile=(8, 4, 20, 16))
    mask = np.zeros((h, w), bool); mask[4:16, 8:20] = True
    assert np.array_equal(part[mask], full[mask]) and not part[~mask].any()
    again = oracle.raster_clip_triangle(w, h, clip, 0x100, 1, full.copy())
    assert np.array_equal(again, full)
