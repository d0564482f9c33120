// oracle/meshopt_ref_wrapper.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The one part of the reference that compiles offline: its vendored meshoptimizer 0.21 (source/asset/meshoptimizer/*.cpp,
// plain C++).  oracle/Makefile compiles those sources WHERE THEY LIE under /root/reference together with this thin C wrapper
// into oracle/_ref/libmeshopt_ref.so (nothing of the reference is copied into the repository).  The wrapper exposes, with
// plain pointers, exactly the calls NaniteBuilder makes (source/asset/nanite_builder.cpp:443-486 buildMeshlets,
// :811-835 simplify, :888 simplifyScale) so that tests can hold chord_b200/nanite_builder.py against the real thing.
// The product (chord_b200/) never loads this library.
#include "meshoptimizer.h"

#include <cstddef>
#include <vector>

extern "C" {

int meshopt_ref_version() { return MESHOPTIMIZER_VERSION; }

// nanite_builder.cpp:443-478: meshopt_buildMeshlets (255 vertices / 128 triangles, cone weight) + meshopt_optimizeMeshlet.
// meshlets: 4 x u32 per meshlet {vertex_offset, triangle_offset, vertex_count, triangle_count}; returns the meshlet count.
size_t meshopt_ref_build_meshlets(const unsigned int* indices, size_t indexCount, const float* positions, size_t vertexCount, size_t strideBytes,
                                  size_t maxVertices, size_t maxTriangles, float coneWeight,
                                  unsigned int* meshlets, size_t meshletCapacity, unsigned int* meshletVertices, unsigned char* meshletTriangles)
{
    const size_t bound = meshopt_buildMeshletsBound(indexCount, maxVertices, maxTriangles);
    if (bound > meshletCapacity) return (size_t)-1;
    std::vector<meshopt_Meshlet> ms(bound);
    const size_t n = meshopt_buildMeshlets(ms.data(), meshletVertices, meshletTriangles, indices, indexCount, positions, vertexCount, strideBytes,
                                           maxVertices, maxTriangles, coneWeight);
    for (size_t i = 0; i < n; i++) {
        meshopt_optimizeMeshlet(&meshletVertices[ms[i].vertex_offset], &meshletTriangles[ms[i].triangle_offset], ms[i].triangle_count, ms[i].vertex_count);
        meshlets[4 * i + 0] = ms[i].vertex_offset; meshlets[4 * i + 1] = ms[i].triangle_offset;
        meshlets[4 * i + 2] = ms[i].vertex_count;  meshlets[4 * i + 3] = ms[i].triangle_count;
    }
    return n;
}

size_t meshopt_ref_build_meshlets_bound(size_t indexCount, size_t maxVertices, size_t maxTriangles)
{ return meshopt_buildMeshletsBound(indexCount, maxVertices, maxTriangles); }

// nanite_builder.cpp:480-486: out = {center[3], radius, cone_apex[3], cone_axis[3], cone_cutoff} (11 floats)
void meshopt_ref_meshlet_bounds(const unsigned int* meshletVertices, const unsigned char* meshletTriangles, size_t triangleCount,
                                const float* positions, size_t vertexCount, size_t strideBytes, float* out11)
{
    const meshopt_Bounds b = meshopt_computeMeshletBounds(meshletVertices, meshletTriangles, triangleCount, positions, vertexCount, strideBytes);
    for (int i = 0; i < 3; i++) { out11[i] = b.center[i]; out11[4 + i] = b.cone_apex[i]; out11[7 + i] = b.cone_axis[i]; }
    out11[3] = b.radius; out11[10] = b.cone_cutoff;
}

// nanite_builder.cpp:811-835: meshopt_simplifyWithAttributes, LockBorder | Sparse | ErrorAbsolute; returns the new index count
size_t meshopt_ref_simplify(unsigned int* destination, const unsigned int* indices, size_t indexCount, const float* positions, size_t vertexCount,
                            size_t positionStrideBytes, const float* attributes, size_t attributeStrideBytes, const float* attributeWeights,
                            size_t attributeCount, size_t targetIndexCount, float targetError, float* resultError)
{
    const unsigned int options = meshopt_SimplifyLockBorder | meshopt_SimplifySparse | meshopt_SimplifyErrorAbsolute;
    return meshopt_simplifyWithAttributes(destination, indices, indexCount, positions, vertexCount, positionStrideBytes, attributes, attributeStrideBytes,
                                          attributeWeights, attributeCount, nullptr, targetIndexCount, targetError, options, resultError);
}

float meshopt_ref_simplify_scale(const float* positions, size_t vertexCount, size_t strideBytes)
{ return meshopt_simplifyScale(positions, vertexCount, strideBytes); }

}  // extern "C"
