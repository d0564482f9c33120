// Test infrastructure only.  The 20-line stand-in for the reference's <utils/utils.h> that lets its dual C++/HLSL headers
// install/resource/shader/{base.h, gltf.h} compile with g++ against the reference's own header-only glm
// (external/include/glm), so that oracle/ref_layout_probe.cpp can read sizeof / offsetof of the a1 structs
// from the REFERENCE's declarations.  The real utils.h drags in pch.h (Win32, spdlog, ...); base.h only needs the names below.
#pragma once
#include <cstdint>
#include <glm/glm.hpp>
namespace chord {
    using uint32 = std::uint32_t;
    namespace math {
        using namespace glm;
        using double3 = glm::dvec3;
    }
}
#define ARCHIVE_DECLARE
