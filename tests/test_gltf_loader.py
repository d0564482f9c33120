"""glTF 2.0 importer -> meshlet-DAG builder -> the reference's asset layout (asset_gltf_helper.cpp:423-590).
Fixtures are written by the test itself (.glb with an interleaved, strided vertex buffer; .gltf with base64 and external
buffers); the imported scene must render exactly like the same mesh handed to the builder directly."""
import base64
import json
import struct

import numpy as np
import pytest

import oracle
from chord_b200 import abi, camera as cam, gltf_loader, nanite_builder as nb, synth

F = np.float32


def _pad4(b, fill=b'\0'):
    return b + fill * ((4 - len(b) % 4) % 4)


def _write_glb(path, doc, blob):
    js = _pad4(json.dumps(doc).encode('utf-8'), b' ')
    blob = _pad4(blob)
    with open(path, 'wb') as fh:
        fh.write(struct.pack('<4sII', b'glTF', 2, 12 + 8 + len(js) + 8 + len(blob)))
        fh.write(struct.pack('<II', len(js), 0x4E4F534A) + js)
        fh.write(struct.pack('<II', len(blob), 0x004E4942) + blob)


def _quat_y(angle):
    return [0.0, float(np.sin(angle / 2)), 0.0, float(np.cos(angle / 2))]


def _fixture(tmp_path):
    pos, tris, uv = nb.terrain_mesh(n=20, seed=3)
    nrm = gltf_loader.smooth_normals(pos, tris)
    # buffer view 0: interleaved position | normal | uv16 (normalized), stride 28 (+ 0 pad); view 1: u16 indices
    uv16 = np.rint(np.clip(uv, 0, 1) * 65535).astype(np.uint16)
    inter = np.zeros((len(pos), 28), dtype=np.uint8)
    inter[:, 0:12] = pos.view(np.uint8).reshape(-1, 12)
    inter[:, 12:24] = nrm.view(np.uint8).reshape(-1, 12)
    inter[:, 24:28] = uv16.view(np.uint8).reshape(-1, 4)
    idx16 = tris.astype(np.uint16).reshape(-1)
    # a box without normals / uvs / indices (triangle soup), u32-free
    c = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=F) * F(2.0)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    soup = np.concatenate([c[[q[0], q[1], q[2]]] for q in quads] + [c[[q[0], q[2], q[3]]] for q in quads]).astype(F)
    blob = inter.tobytes() + _pad4(idx16.tobytes()) + soup.tobytes()
    o_idx = len(inter.tobytes())
    o_soup = o_idx + len(_pad4(idx16.tobytes()))
    doc = {
        'asset': {'version': '2.0'},
        'buffers': [{'byteLength': len(blob)}],
        'bufferViews': [
            {'buffer': 0, 'byteOffset': 0, 'byteLength': inter.size, 'byteStride': 28},
            {'buffer': 0, 'byteOffset': o_idx, 'byteLength': idx16.nbytes},
            {'buffer': 0, 'byteOffset': o_soup, 'byteLength': soup.nbytes},
        ],
        'accessors': [
            {'bufferView': 0, 'byteOffset': 0, 'componentType': 5126, 'count': len(pos), 'type': 'VEC3'},
            {'bufferView': 0, 'byteOffset': 12, 'componentType': 5126, 'count': len(pos), 'type': 'VEC3'},
            {'bufferView': 0, 'byteOffset': 24, 'componentType': 5123, 'normalized': True, 'count': len(pos), 'type': 'VEC2'},
            {'bufferView': 1, 'componentType': 5123, 'count': len(idx16), 'type': 'SCALAR'},
            {'bufferView': 2, 'componentType': 5126, 'count': len(soup), 'type': 'VEC3'},
        ],
        'materials': [
            {'name': 'ground', 'pbrMetallicRoughness': {'baseColorFactor': [0.4, 0.5, 0.3, 1.0], 'metallicFactor': 0.0, 'roughnessFactor': 0.8}},
            {'name': 'leaf', 'alphaMode': 'MASK', 'alphaCutoff': 0.3, 'doubleSided': True},
            {'name': 'glass', 'alphaMode': 'BLEND'},
        ],
        'meshes': [
            {'name': 'terrain', 'primitives': [{'attributes': {'POSITION': 0, 'NORMAL': 1, 'TEXCOORD_0': 2}, 'indices': 3, 'material': 0},
                                                {'attributes': {'POSITION': 4}, 'mode': 1}]},            # a line list: skipped
            {'name': 'box', 'primitives': [{'attributes': {'POSITION': 4}}]},                             # no material -> default
            {'name': 'terrain_again', 'primitives': [{'attributes': {'POSITION': 0, 'NORMAL': 1, 'TEXCOORD_0': 2}, 'indices': 3, 'material': 1}]},
        ],
        'nodes': [
            {'name': 'root', 'translation': [1.0, 0.0, -2.0], 'rotation': _quat_y(0.3), 'scale': [1.0, 1.0, 1.0], 'children': [1, 2], 'mesh': 0},
            {'name': 'box', 'matrix': [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 3.0, 4.0, 5.0, 1], 'mesh': 1},
            {'name': 'far', 'translation': [60.0, 0.0, 0.0], 'scale': [0.5, 0.5, 0.5], 'mesh': 2},
            {'name': 'unused', 'mesh': 1},                                                                # not in the scene
        ],
        'scenes': [{'nodes': [0]}], 'scene': 0,
    }
    glb = tmp_path / 'scene.glb'
    _write_glb(glb, doc, blob)
    # the same document as .gltf: the vertex data in an external .bin, indices + soup as a base64 data uri
    ext = dict(doc)
    (tmp_path / 'verts.bin').write_bytes(inter.tobytes())
    rest = blob[o_idx:]
    ext['buffers'] = [{'byteLength': inter.size, 'uri': 'verts.bin'},
                      {'byteLength': len(rest), 'uri': 'data:application/octet-stream;base64,' + base64.b64encode(rest).decode()}]
    ext['bufferViews'] = [dict(doc['bufferViews'][0]),
                          {'buffer': 1, 'byteOffset': 0, 'byteLength': idx16.nbytes},
                          {'buffer': 1, 'byteOffset': o_soup - o_idx, 'byteLength': soup.nbytes}]
    gltf = tmp_path / 'scene.gltf'
    gltf.write_text(json.dumps(ext))
    return glb, gltf, dict(pos=pos, tris=tris, uv=uv, nrm=nrm, soup=soup)


def test_import_matches_the_source_arrays(tmp_path):
    glb, gltf, src = _fixture(tmp_path)
    for path in (glb, gltf):
        g = gltf_loader.load(str(path))
        assert [p['name'] for p in g['primitives']] == ['terrain.0', 'box.0']            # the second terrain re-uses the first import
        t, b = g['primitives']
        assert np.array_equal(t['positions'], src['pos']) and np.array_equal(t['triangles'], src['tris']) and np.array_equal(t['normals'], src['nrm'])
        assert np.abs(t['uv0'] - src['uv']).max() <= 0.5 / 65535 + 1e-7                    # normalized u16
        assert np.array_equal(b['positions'], src['soup']) and np.array_equal(b['triangles'], np.arange(36).reshape(12, 3))
        assert np.allclose(np.linalg.norm(b['normals'], axis=1), 1.0, atol=1e-6)           # generated
        assert len(g['materials']) == 4 and g['materials'][1]['alphaMode'] == 1 and g['materials'][1]['bTwoSided'] == 1
        assert g['materials'][2]['alphaMode'] == 2 and np.isclose(g['materials'][1]['alphaCutOff'], 0.3)
        inst = {(p, m): w for (w, p, m) in g['instances']}
        assert set(inst) == {(0, 0), (1, 3), (0, 1)}                                       # 'unused' is not in the default scene
        root = np.eye(4); a = 0.3
        root[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]; root[:3, 3] = (1, 0, -2)
        assert np.allclose(inst[(0, 0)], root)
        assert np.allclose(inst[(1, 3)], root @ np.array([[1, 0, 0, 3.0], [0, 1, 0, 4.0], [0, 0, 1, 5.0], [0, 0, 0, 1]]))
        far = np.diag([0.5, 0.5, 0.5, 1.0]); far[:3, 3] = (60, 0, 0)
        assert np.allclose(inst[(0, 1)], root @ far)


def test_imported_scene_renders_like_the_direct_build(tmp_path):
    glb, _, src = _fixture(tmp_path)
    scene, info = gltf_loader.build_scene(str(glb))
    assert info['triangles'] == 2 * len(src['tris']) + 12 and len(info['transforms']) == 3
    w, h = 320, 180
    c = cam.Camera((2.0, 9.0, 22.0), (0.0, -0.4, -1.0), width=w, height=h)
    view = cam.fill_view(c)
    iview = cam.fill_instance_view(view)
    objects = synth.refill_objects(info['transforms'], c)
    fr = oracle.Frame(scene, view, iview, objects, w, h)
    post = fr.instance_culling(info['lod0MeshletCount'])
    vis, st = fr.render_mesh(post)
    assert (vis != 0).mean() > 0.3 and st['trianglesSubmitted'] > 0
    assert set(np.unique(post['objectId'])) >= {0, 1}                                       # terrain and box are in view
    # the root instance alone == the same mesh through the builder directly with the same transform and material
    direct = nb.build_native(src['pos'], src['tris'], normals=src['nrm'], uv0=scene.datas[0]['uv0'])
    dscene = nb.host_scene(direct, materials=[scene.materials[0]])
    m0 = info['transforms'][[i for i, t in enumerate(info['transforms']) if t[1] == 0 and t[2] == 0][0]][0]
    dobjs = np.array([cam.fill_object(m0, c, 0, 0)])
    dfr = oracle.Frame(dscene, view, iview, dobjs, w, h)
    dvis, _ = dfr.render_mesh(dfr.instance_culling(direct['lod0MeshletCount']))
    only = np.nonzero((objects['GLTFPrimitiveDetail'] == 0) & (objects['GLTFMaterialData'] == 0))[0]
    sfr = oracle.Frame(scene, view, iview, objects[only], w, h)
    svis, _ = sfr.render_mesh(sfr.instance_culling(info['lod0MeshletCount']))
    assert np.array_equal(svis >> np.uint64(32), dvis >> np.uint64(32))                    # same depth image
    assert np.array_equal(oracle.decode(svis, sfr.instance_culling(info['lod0MeshletCount']))[..., 1:], oracle.decode(dvis, dfr.instance_culling(direct['lod0MeshletCount']))[..., 1:])


def test_malformed_files_are_rejected(tmp_path):
    glb, gltf, _ = _fixture(tmp_path)
    doc = json.loads(gltf.read_text())
    bad = dict(doc); bad['asset'] = {'version': '1.0'}
    (tmp_path / 'v1.gltf').write_text(json.dumps(bad))
    with pytest.raises(gltf_loader.GltfError):
        gltf_loader.load(str(tmp_path / 'v1.gltf'))
    bad = json.loads(gltf.read_text()); bad['accessors'][3]['count'] = 10 ** 6
    (tmp_path / 'overrun.gltf').write_text(json.dumps(bad))
    with pytest.raises(gltf_loader.GltfError):
        gltf_loader.load(str(tmp_path / 'overrun.gltf'))
    bad = json.loads(gltf.read_text()); bad['materials'][1]['alphaMode'] = 'DITHER'
    (tmp_path / 'mode.gltf').write_text(json.dumps(bad))
    with pytest.raises(gltf_loader.GltfError):
        gltf_loader.load(str(tmp_path / 'mode.gltf'))
    raw = bytearray(glb.read_bytes()); raw[4:8] = struct.pack('<I', 1)
    (tmp_path / 'v1.glb').write_bytes(bytes(raw))
    with pytest.raises(gltf_loader.GltfError):
        gltf_loader.load(str(tmp_path / 'v1.glb'))
    empty = dict(doc); empty['scenes'] = [{'nodes': [3]}]; empty['nodes'] = [dict(n) for n in doc['nodes']]; empty['nodes'][3] = {'name': 'nothing'}
    (tmp_path / 'empty.gltf').write_text(json.dumps(empty))
    with pytest.raises(gltf_loader.GltfError):
        gltf_loader.build_scene(str(tmp_path / 'empty.gltf'))


def test_accessor_variants():
    """sparse accessors over a zero-initialised base, normalized signed bytes, tightly packed scalars."""
    idx = np.array([1, 3], dtype=np.uint16)
    vals = np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], dtype=F)
    nb8 = np.array([[-128, 127], [0, 64]], dtype=np.int8)
    blob = _pad4(idx.tobytes()) + vals.tobytes() + nb8.tobytes()
    doc = {'bufferViews': [{'buffer': 0, 'byteOffset': 0, 'byteLength': 4}, {'buffer': 0, 'byteOffset': 4, 'byteLength': 24},
                           {'buffer': 0, 'byteOffset': 28, 'byteLength': 4}],
           'accessors': [{'componentType': 5126, 'count': 4, 'type': 'VEC3',
                          'sparse': {'count': 2, 'indices': {'bufferView': 0, 'componentType': 5123}, 'values': {'bufferView': 1}}},
                         {'bufferView': 2, 'componentType': 5120, 'normalized': True, 'count': 2, 'type': 'VEC2'}]}
    a = gltf_loader._accessor(doc, [blob], 0)
    assert np.array_equal(a, np.array([[0, 0, 0], [1, 2, 3], [0, 0, 0], [4, 5, 6]], dtype=F))
    b = gltf_loader._accessor(doc, [blob], 1)
    assert b[0, 0] == -1.0 and b[0, 1] == 1.0 and b[1, 0] == 0.0 and np.isclose(b[1, 1], 64 / 127)
