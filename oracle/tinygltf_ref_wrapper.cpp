// oracle/tinygltf_ref_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The glTF parser the reference imports assets with (tinygltf, vendored header-only under external/include/tinygltf and
// configured like source/pch.h:131-135 / pch.cpp:13, minus the image codecs) behind a thin C interface, compiled by
// oracle/Makefile (target `ref`) from the headers WHERE THEY LIE under /root/reference into oracle/_ref/libtinygltf_ref.so.
// tests/test_gltf_loader_ref.py holds chord_b200/gltf_loader.py against it.  The product never loads this library.
#include <nlohmann/json.hpp>

#define TINYGLTF_IMPLEMENTATION
#define TINYGLTF_NO_STB_IMAGE
#define TINYGLTF_NO_STB_IMAGE_WRITE
#define TINYGLTF_NO_INCLUDE_JSON
#define TINYGLTF_USE_CPP14
#include <tinygltf/tiny_gltf.h>

#include <cstring>
#include <string>

static tinygltf::Model g_model;
static std::string g_err;

extern "C" {

// 0 = ok; the error / warning text is kept for tg_last_error
int tg_load(const char* path)
{
    tinygltf::TinyGLTF loader;
    std::string err, warn;
    g_model = tinygltf::Model();
    const std::string p(path);
    const bool binary = p.size() > 4 && p.substr(p.size() - 4) == ".glb";
    const bool ok = binary ? loader.LoadBinaryFromFile(&g_model, &err, &warn, p) : loader.LoadASCIIFromFile(&g_model, &err, &warn, p);
    g_err = err + warn;
    return ok ? 0 : 1;
}
const char* tg_last_error() { return g_err.c_str(); }

int tg_mesh_count() { return (int)g_model.meshes.size(); }
int tg_primitive_count(int mesh) { return (int)g_model.meshes[mesh].primitives.size(); }

// out = {mode, indices accessor, material, POSITION, NORMAL, TEXCOORD_0} (-1 = absent)
void tg_primitive_info(int mesh, int prim, int* out6)
{
    const tinygltf::Primitive& p = g_model.meshes[mesh].primitives[prim];
    auto attr = [&](const char* name) { auto it = p.attributes.find(name); return it == p.attributes.end() ? -1 : it->second; };
    out6[0] = p.mode; out6[1] = p.indices; out6[2] = p.material; out6[3] = attr("POSITION"); out6[4] = attr("NORMAL"); out6[5] = attr("TEXCOORD_0");
}

// out = {count, componentType, components per element, normalized}
void tg_accessor_info(int accessor, int* out4)
{
    const tinygltf::Accessor& a = g_model.accessors[accessor];
    out4[0] = (int)a.count; out4[1] = a.componentType; out4[2] = tinygltf::GetNumComponentsInType((uint32_t)a.type); out4[3] = a.normalized ? 1 : 0;
}

// tightly packed copy of the accessor's elements, honouring the buffer view's stride (what the reference's importer does with
// accessor / bufferView / buffer, asset_gltf_helper.cpp:470-520); returns the bytes written or -1
long long tg_accessor_read(int accessor, unsigned char* dst, long long dstBytes)
{
    const tinygltf::Accessor& a = g_model.accessors[accessor];
    if (a.bufferView < 0) return -1;
    const tinygltf::BufferView& bv = g_model.bufferViews[a.bufferView];
    const tinygltf::Buffer& buf = g_model.buffers[bv.buffer];
    const int elem = tinygltf::GetComponentSizeInBytes((uint32_t)a.componentType) * tinygltf::GetNumComponentsInType((uint32_t)a.type);
    const int stride = a.ByteStride(bv);
    if (elem <= 0 || stride <= 0 || (long long)a.count * elem > dstBytes) return -1;
    const unsigned char* src = buf.data.data() + bv.byteOffset + a.byteOffset;
    for (size_t i = 0; i < a.count; i++) memcpy(dst + i * elem, src + i * (size_t)stride, (size_t)elem);
    return (long long)a.count * elem;
}

int tg_material_count() { return (int)g_model.materials.size(); }

// out = {baseColorFactor[4], metallicFactor, roughnessFactor, alphaCutoff, alphaMode (0 opaque, 1 mask, 2 blend), doubleSided, emissiveFactor[3]}
void tg_material_info(int material, double* out12)
{
    const tinygltf::Material& m = g_model.materials[material];
    for (int i = 0; i < 4; i++) out12[i] = m.pbrMetallicRoughness.baseColorFactor[i];
    out12[4] = m.pbrMetallicRoughness.metallicFactor; out12[5] = m.pbrMetallicRoughness.roughnessFactor; out12[6] = m.alphaCutoff;
    out12[7] = m.alphaMode == "MASK" ? 1 : m.alphaMode == "BLEND" ? 2 : 0;
    out12[8] = m.doubleSided ? 1 : 0;
    for (int i = 0; i < 3; i++) out12[9 + i] = m.emissiveFactor[i];
}

}  // extern "C"
