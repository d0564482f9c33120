"""chord_b200/gltf_loader.py held against the glTF parser the reference imports assets with: tinygltf (vendored header-only in
the reference, compiled from there into oracle/_ref/libtinygltf_ref.so by oracle/Makefile target `ref`).  Same files, same
accessor bytes, same primitive / material records.  Skipped when oracle/_ref is not available."""
import ctypes as C

import numpy as np
import pytest

import oracle
from chord_b200 import gltf_loader
from test_gltf_loader import _fixture

TG = oracle.tinygltf_ref()
pytestmark = pytest.mark.skipif(TG is None, reason='oracle/_ref/libtinygltf_ref.so not available (no /root/reference)')

_DT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}


def _tg_accessor(index):
    info = (C.c_int * 4)()
    TG.tg_accessor_info(C.c_int(index), info)
    count, ctype, width, normalized = info
    arr = np.zeros((count, width), dtype=_DT[ctype])
    n = TG.tg_accessor_read(C.c_int(index), C.c_void_p(arr.ctypes.data), C.c_longlong(arr.nbytes))
    assert n == arr.nbytes
    return arr, bool(normalized)


@pytest.mark.parametrize('which', ['glb', 'gltf'])
def test_loader_reads_what_tinygltf_reads(tmp_path, which):
    glb, gltf, _ = _fixture(tmp_path)
    path = str(glb if which == 'glb' else gltf)
    assert TG.tg_load(path.encode()) == 0, TG.tg_last_error()
    ours = gltf_loader.load(path)
    # every triangle primitive, in document order, against our de-duplicated import
    seen = {}
    for mi in range(TG.tg_mesh_count()):
        for pi in range(TG.tg_primitive_count(C.c_int(mi))):
            info = (C.c_int * 6)()
            TG.tg_primitive_info(C.c_int(mi), C.c_int(pi), info)
            mode, indices, material, a_pos, a_nrm, a_uv = info
            if mode != 4:
                continue
            key = (a_pos, a_nrm, a_uv, indices)
            if key not in seen:
                seen[key] = len(seen)
            prim = ours['primitives'][seen[key]]
            pos, _ = _tg_accessor(a_pos)
            assert np.array_equal(prim['positions'], pos)
            if indices >= 0:
                idx, _ = _tg_accessor(indices)
                assert np.array_equal(prim['triangles'].reshape(-1), idx.reshape(-1).astype(np.int64))
            else:
                assert np.array_equal(prim['triangles'].reshape(-1), np.arange(len(pos)))
            if a_nrm >= 0:
                assert np.array_equal(prim['normals'], _tg_accessor(a_nrm)[0])
            if a_uv >= 0:
                raw, normalized = _tg_accessor(a_uv)
                want = raw.astype(np.float32) / np.float32(np.iinfo(raw.dtype).max) if normalized else raw
                assert np.array_equal(prim['uv0'], want)
    assert len(seen) == len(ours['primitives']) == 2
    # materials: tinygltf fills the glTF defaults the same way
    assert TG.tg_material_count() == len(ours['materials']) - 1
    for i in range(TG.tg_material_count()):
        m = (C.c_double * 12)()
        TG.tg_material_info(C.c_int(i), m)
        rec = ours['materials'][i]
        assert np.allclose(rec['baseColorFactor'], m[0:4]) and np.isclose(rec['metallicFactor'], m[4]) and np.isclose(rec['roughnessFactor'], m[5])
        assert np.isclose(rec['alphaCutOff'], m[6]) and rec['alphaMode'] == int(m[7]) and rec['bTwoSided'] == int(m[8])
        assert np.allclose(rec['emissiveFactor'], m[9:12])
