"""Regenerates tests/golden/ref_struct_layout.json: sizeof / offsetof of every struct of SURVEY.md §8 row a1 read from the
reference's OWN headers (/root/reference/install/resource/shader/{base.h, gltf.h}), compiled in place by
oracle/Makefile (target `ref`: g++ -std=c++20 + the reference's header-only glm + oracle/ref_shim/utils/utils.h).
Run in the development container (the GPU box has no /root/reference); tests/test_abi.py compares the fixture - and, where
the reference tree exists, a fresh probe run - with include/chord_abi_types.h field by field."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

if __name__ == '__main__':
    layout = oracle.ref_struct_layout()
    assert layout is not None, 'oracle/_ref/ref_layout_probe could not be built (no /root/reference?)'
    out = os.path.join(ROOT, 'tests', 'golden', 'ref_struct_layout.json')
    with open(out, 'w') as f:
        json.dump(layout, f, indent=1, sort_keys=True)
        f.write('\n')
    print(out, len(layout) - 1, 'structs')
