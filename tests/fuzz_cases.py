"""Seeded random raster cases (shared by the GPU suite and the host-executor suite): patch streams and ragged meshlet streams
at odd frame sizes, seen from a camera that is pushed into the stream and turned, so that triangles cross the near plane, the
guard band and the frustum sides, with and without a screen tile as scissor."""
import numpy as np

from chord_b200 import camera as cam, synth


def random_raster_case(rng):
    seed = int(rng.integers(1, 10 ** 6))
    w = int(rng.choice([64, 72, 200, 256, 328, 640]))
    h = int(rng.choice([64, 72, 136, 256, 360]))
    kind = int(rng.integers(0, 2))
    n = int(rng.integers(20, 300))
    if kind == 0:
        scale = float(rng.choice([1.0, 5.0, 40.0, 300.0, 3000.0]))
        scene, camera, objects, cmds = synth.patch_stream_scene(n, w, h, seed=seed, overdraw_scale=scale)
    else:
        scene, camera, objects, cmds = synth.ragged_stream_scene(n, w, h, seed=seed)
    fwd = np.asarray(camera.front, dtype=np.float64)
    pos = camera.position + fwd * float(rng.choice([0.0, 0.0, 2.0, 20.0, 60.0])) + rng.normal(0, 1.0, 3) * float(rng.choice([0.0, 0.5, 5.0]))
    front = fwd + rng.normal(0, 1.0, 3) * float(rng.choice([0.0, 0.05, 0.4]))
    c2 = cam.Camera(pos, front, width=w, height=h)
    view = cam.fill_view(c2)
    objs = np.array([cam.fill_object(np.eye(4), c2, int(o['GLTFPrimitiveDetail']), int(o['GLTFMaterialData'])) for o in objects])
    tile = None
    if rng.random() < 0.3 and w >= 256 and h >= 136:
        tile = (64, 64, w - 64 if (w - 64) % 64 == 0 else w, h)
    return dict(scene=scene, view=view, iview=cam.fill_instance_view(view), objects=objs, cmds=cmds, w=w, h=h, tile=tile,
                lod0=len(cmds) if kind == 0 else 2 * n, groups=((n + 3) // 4) * (1 if kind == 0 else 2), what=dict(seed=seed, w=w, h=h, kind=kind, n=n, tile=tile))
