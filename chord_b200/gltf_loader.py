"""Minimal glTF 2.0 importer feeding the meshlet-DAG builder: real assets -> chord's GPU asset layout.

Mirrors what the reference's importer keeps of a glTF file for this path (``source/asset/gltf/asset_gltf_helper.cpp:423-590``
processMesh / ``:226-330`` material import, via tinygltf there): per mesh primitive the triangle list with POSITION, NORMAL
(smooth normals are generated when absent, like ``bGenerateSmoothNormal``), TEXCOORD_0; primitives with the same attribute
accessors are imported once and shared; per material the metallic-roughness factors, alpha mode / cutoff and double
sidedness (textures are not imported: the resolve is untextured); per node the world matrix of the
default scene (matrix or TRS, parents applied).  Only ``mode == 4`` (triangles) is taken, as in the reference (``:427-431``).

Reads ``.gltf`` (JSON + external or base64 buffers) and ``.glb`` (binary container).  Pure numpy / json, offline tool: the
per-frame path never calls it.
"""
import base64
import json
import os
import struct

import numpy as np

from . import abi, nanite_builder
from .synth import default_material

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {'SCALAR': 1, 'VEC2': 2, 'VEC3': 3, 'VEC4': 4, 'MAT2': 4, 'MAT3': 9, 'MAT4': 16}
_ALPHA_MODE = {'OPAQUE': 0, 'MASK': 1, 'BLEND': 2}


class GltfError(ValueError):
    pass


def _read_container(path):
    """-> (json dict, binary chunk of a .glb or None)."""
    raw = open(path, 'rb').read()
    if raw[:4] == b'glTF':
        magic, version, length = struct.unpack_from('<4sII', raw, 0)
        if version != 2:
            raise GltfError(f'{path}: glb version {version} (only glTF 2.0 is supported)')
        off, doc, blob = 12, None, None
        while off + 8 <= min(length, len(raw)):
            clen, ctype = struct.unpack_from('<II', raw, off)
            chunk = raw[off + 8:off + 8 + clen]
            if ctype == 0x4E4F534A:
                doc = json.loads(chunk.decode('utf-8'))
            elif ctype == 0x004E4942 and blob is None:
                blob = chunk
            off += 8 + clen + ((4 - clen % 4) % 4)
        if doc is None:
            raise GltfError(f'{path}: glb without a JSON chunk')
        return doc, blob
    return json.loads(raw.decode('utf-8')), None


def _buffers(doc, blob, base_dir):
    out = []
    for i, b in enumerate(doc.get('buffers', [])):
        uri = b.get('uri')
        if uri is None:
            if blob is None:
                raise GltfError(f'buffer {i} has no uri and the file has no binary chunk')
            data = blob
        elif uri.startswith('data:'):
            data = base64.b64decode(uri.split(',', 1)[1])
        else:
            data = open(os.path.join(base_dir, uri), 'rb').read()
        if len(data) < b.get('byteLength', 0):
            raise GltfError(f'buffer {i}: {len(data)} bytes, {b["byteLength"]} declared')
        out.append(data)
    return out


def _accessor(doc, buffers, index, normalize=True):
    """-> (count, width) array; normalized integer accessors become floats in [0, 1] / [-1, 1]."""
    a = doc['accessors'][index]
    dt, width, count = np.dtype(_COMPONENT[a['componentType']]), _WIDTH[a['type']], a['count']
    if 'bufferView' not in a:
        arr = np.zeros((count, width), dtype=dt)
    else:
        bv = doc['bufferViews'][a['bufferView']]
        start = bv.get('byteOffset', 0) + a.get('byteOffset', 0)
        stride = bv.get('byteStride', 0) or dt.itemsize * width
        buf = buffers[bv['buffer']]
        need = start + (count - 1) * stride + dt.itemsize * width if count else start
        if need > len(buf):
            raise GltfError(f'accessor {index} reads past the end of buffer {bv["buffer"]}')
        arr = np.lib.stride_tricks.as_strided(np.frombuffer(buf, dtype=np.uint8, offset=start, count=max(need - start, 0)),
                                              shape=(count, width * dt.itemsize), strides=(stride, 1)).copy().view(dt).reshape(count, width)
    if 'sparse' in a:
        sp = a['sparse']
        ia = dict(bufferView=sp['indices']['bufferView'], byteOffset=sp['indices'].get('byteOffset', 0),
                  componentType=sp['indices']['componentType'], count=sp['count'], type='SCALAR')
        va = dict(bufferView=sp['values']['bufferView'], byteOffset=sp['values'].get('byteOffset', 0),
                  componentType=a['componentType'], count=sp['count'], type=a['type'])
        tmp = dict(doc, accessors=[ia, va])
        arr = arr.copy()
        arr[_accessor(tmp, buffers, 0)[:, 0].astype(np.int64)] = _accessor(tmp, buffers, 1, normalize=False)
    if normalize and a.get('normalized') and dt.kind in 'iu':
        info = np.iinfo(dt)
        arr = np.maximum(arr.astype(np.float32) / np.float32(info.max), -1.0) if dt.kind == 'i' else arr.astype(np.float32) / np.float32(info.max)
    return arr


def _node_matrix(n):
    if 'matrix' in n:
        return np.array(n['matrix'], dtype=np.float64).reshape(4, 4).T            # glTF stores column-major
    m = np.eye(4)
    x, y, z, w = n.get('rotation', (0.0, 0.0, 0.0, 1.0))
    r = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    m[:3, :3] = r * np.array(n.get('scale', (1.0, 1.0, 1.0)))[None, :]
    m[:3, 3] = n.get('translation', (0.0, 0.0, 0.0))
    return m


def smooth_normals(pos, tris):
    """area-weighted vertex normals (the reference generates them with the same weighting when NORMAL is absent)."""
    acc = np.zeros(pos.shape, dtype=np.float64)
    fn = np.cross(pos[tris[:, 1]].astype(np.float64) - pos[tris[:, 0]], pos[tris[:, 2]].astype(np.float64) - pos[tris[:, 0]])
    for k in range(3):
        np.add.at(acc, tris[:, k], fn)
    return (acc / np.maximum(np.linalg.norm(acc, axis=1, keepdims=True), 1e-30)).astype(np.float32)


def load(path):
    """-> dict(primitives=[dict(positions, triangles, normals, uv0, name)], materials=[GLTFMaterialGPUData],
    instances=[(world matrix 4x4 float64, primitive index, material index)]); material index len(materials) - 1 is the
    default material for primitives without one."""
    doc, blob = _read_container(path)
    if doc.get('asset', {}).get('version', '2.0')[0] != '2':
        raise GltfError(f'{path}: glTF version {doc["asset"].get("version")} (only 2.0 is supported)')
    buffers = _buffers(doc, blob, os.path.dirname(os.path.abspath(path)))
    materials = []
    for m in doc.get('materials', []):
        pbr = m.get('pbrMetallicRoughness', {})
        mode = m.get('alphaMode', 'OPAQUE')
        if mode not in _ALPHA_MODE:
            raise GltfError(f'material {m.get("name")}: alphaMode {mode}')
        rec = default_material(two_sided=1 if m.get('doubleSided') else 0, alpha_mode=_ALPHA_MODE[mode],
                               base_color=tuple(pbr.get('baseColorFactor', (1.0, 1.0, 1.0, 1.0))), metallic=pbr.get('metallicFactor', 1.0),
                               roughness=pbr.get('roughnessFactor', 1.0), alpha_cutoff=m.get('alphaCutoff', 0.5))
        rec['emissiveFactor'] = m.get('emissiveFactor', (0.0, 0.0, 0.0))
        materials.append(rec)
    materials.append(default_material(base_color=(1.0, 1.0, 1.0, 1.0), metallic=1.0, roughness=1.0))      # glTF's default material
    prims, cache, mesh_prims = [], {}, []
    for mi, mesh in enumerate(doc.get('meshes', [])):
        entries = []
        for pi, p in enumerate(mesh['primitives']):
            name = f'{mesh.get("name", f"mesh{mi}")}.{pi}'
            if p.get('mode', 4) != 4:
                continue                                   # only triangle lists, like the reference
            if 'POSITION' not in p['attributes']:
                raise GltfError(f'{name}: no POSITION')
            key = (tuple(sorted(p['attributes'].items())), p.get('indices'))
            if key not in cache:
                pos = _accessor(doc, buffers, p['attributes']['POSITION']).astype(np.float32)
                if 'indices' in p:
                    idx = _accessor(doc, buffers, p['indices'], normalize=False)[:, 0].astype(np.int64)
                else:
                    idx = np.arange(len(pos), dtype=np.int64)
                if len(idx) % 3 or (len(idx) and idx.max() >= len(pos)):
                    raise GltfError(f'{name}: malformed index list')
                tris = idx.reshape(-1, 3)
                tris = tris[(tris[:, 0] != tris[:, 1]) & (tris[:, 1] != tris[:, 2]) & (tris[:, 0] != tris[:, 2])]
                nrm = _accessor(doc, buffers, p['attributes']['NORMAL']).astype(np.float32) if 'NORMAL' in p['attributes'] else smooth_normals(pos, tris)
                uv = _accessor(doc, buffers, p['attributes']['TEXCOORD_0']).astype(np.float32) if 'TEXCOORD_0' in p['attributes'] else np.zeros((len(pos), 2), np.float32)
                cache[key] = len(prims)
                prims.append(dict(positions=pos, triangles=tris, normals=nrm, uv0=uv[:, :2], name=name))
            entries.append((cache[key], p.get('material', len(materials) - 1)))
        mesh_prims.append(entries)
    instances = []
    nodes = doc.get('nodes', [])
    scenes = doc.get('scenes', [])
    roots = scenes[doc.get('scene', 0)]['nodes'] if scenes else [i for i in range(len(nodes)) if not any(i in n.get('children', []) for n in nodes)]
    stack = [(r, np.eye(4)) for r in reversed(roots)]                  # depth first, in document order
    seen = 0
    while stack:
        ni, parent = stack.pop()
        seen += 1
        if seen > 16 * max(len(nodes), 1):
            raise GltfError('node hierarchy has a cycle')
        n = nodes[ni]
        world = parent @ _node_matrix(n)
        if 'mesh' in n:
            for (prim, material) in mesh_prims[n['mesh']]:
                instances.append((world, prim, material))
        stack += [(c, world) for c in reversed(n.get('children', []))]
    return dict(primitives=prims, materials=materials, instances=instances)


def build_scene(path, max_lods=nanite_builder.MAX_LODS, native=True):
    """glTF file -> (abi.HostScene, info) with the meshlet DAG of every imported primitive.  ``info['transforms']`` is what
    ``synth.refill_objects`` takes, ``lod0MeshletCount`` / ``meshletGroupCount`` size the per-frame work lists."""
    g = load(path)
    if not g['primitives'] or not g['instances']:
        raise GltfError(f'{path}: no triangle geometry in the default scene')
    datas, prims, lod0 = [], [], []
    for i, p in enumerate(g['primitives']):
        if len(p['triangles']) == 0:
            raise GltfError(f'{p["name"]}: empty primitive')
        # native: the C++ core (csrc/nanite_builder.cpp), what makes Sponza / Bistro-sized primitives practical
        asset = (nanite_builder.build_native if native else nanite_builder.build)(p['positions'], p['triangles'], normals=p['normals'], uv0=p['uv0'], max_lods=max_lods)
        rec = asset['primitive'].copy()
        rec['primitiveDatasBufferId'] = i
        datas.append(asset['datas'])
        prims.append(rec[0])
        lod0.append(asset['lod0MeshletCount'])
    scene = abi.HostScene(datas, np.array(prims, dtype=abi.GLTFPrimitiveBuffer), np.array(g['materials']))
    transforms = [(m, prim, mat) for (m, prim, mat) in g['instances']]
    info = dict(transforms=transforms, names=[p['name'] for p in g['primitives']],
                lod0MeshletCount=int(sum(lod0[p] for (_, p, _) in transforms)),
                meshletGroupCount=int(sum(int(prims[p]['meshletGroupCount']) for (_, p, _) in transforms)),
                triangles=int(sum(len(g['primitives'][p]['triangles']) for (_, p, _) in transforms)))
    return scene, info
