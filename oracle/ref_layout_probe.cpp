// Test infrastructure only (never linked into the product).  Compiled by oracle/Makefile (target `ref`) against the
// reference's OWN headers where they lie: -I $(REFERENCE)/install/resource/shader (base.h, gltf.h), -I its header-only glm,
// -I oracle/ref_shim (utils/utils.h stand-in).  Prints sizeof / offsetof of every struct of SURVEY.md §8 row a1 as JSON;
// tests/test_abi.py compares it field by field with include/chord_abi_types.h, and tests/golden/ref_struct_layout.json keeps
// the output so that the comparison also runs where /root/reference does not exist (the GPU box).
#include <cstddef>
#include <cstdio>
#include <gltf.h>

#define BEGIN(T) { using S = T; std::printf("%s\"%s\": {\"sizeof\": %zu, \"fields\": {", first ? "" : ",\n", #T, sizeof(T)); first = false; bool f0 = true;
#define F(m) std::printf("%s\"%s\": [%zu, %zu]", f0 ? "" : ", ", #m, offsetof(S, m), sizeof(S::m)); f0 = false;
#define END std::printf("}}"); }

int main()
{
    bool first = true;
    std::printf("{\n");
    BEGIN(GPUBVHNode) F(sphere) F(children) F(bvhNodeCount) F(leafMeshletGroupOffset) F(leafMeshletGroupCount) END
    BEGIN(GPUGLTFMeshletGroup) F(clusterPosCenter) F(parentError) F(parentPosCenter) F(error) F(meshletOffset) F(meshletCount) END
    BEGIN(GPUGLTFMeshlet) F(posMin) F(dataOffset) F(posMax) F(vertexTriangleCount) F(coneAxis) F(coneCutOff) F(coneApex) F(lod) END
    BEGIN(GLTFPrimitiveBuffer) F(posMin) F(primitiveDatasBufferId) F(posMax) F(vertexOffset) F(posAverage) F(vertexCount) F(meshletOffset)
        F(color0Offset) F(smoothNormalOffset) F(textureCoord1Offset) F(bvhNodeOffset) F(meshletGroupOffset) F(meshletGroupIndicesOffset)
        F(meshletGroupCount) F(lod0IndicesOffset) F(lod0IndicesCount) END
    BEGIN(GLTFPrimitiveDatasBuffer) F(meshletBuffer) F(positionBuffer) F(normalBuffer) F(textureCoord0Buffer) F(tangentBuffer) F(textureCoord1Buffer)
        F(color0Buffer) F(smoothNormalsBuffer) F(meshletDataBuffer) F(bvhNodeBuffer) F(meshletGroupBuffer) F(meshletGroupIndicesBuffer) F(lod0IndicesBuffer) END
    BEGIN(GLTFMaterialGPUData) F(alphaMode) F(alphaCutOff) F(bTwoSided) F(baseColorId) F(baseColorFactor) F(emissiveFactor) F(emissiveTexture)
        F(metallicFactor) F(roughnessFactor) F(metallicRoughnessTexture) F(normalTexture) F(baseColorSampler) F(emissiveSampler) F(normalSampler)
        F(metallicRoughnessSampler) F(normalFactorScale) F(bExistOcclusion) F(occlusionTextureStrength) F(materialType) END
    BEGIN(GPUStorageDouble4) F(x) F(y) F(z) F(w) END
    BEGIN(InstanceCullingViewInfo) F(translatedWorldToClip) F(clipToTranslatedWorld) F(cameraWorldPos) F(orthoDepthConvertToView) F(renderDimension) F(frustumPlanesRS) END
    BEGIN(AtmosphereParameters) F(solar_irradiance) F(sun_angular_radius) F(bottom_radius) F(top_radius) F(luminanceMode) F(rayleigh_density)
        F(rayleigh_scattering) F(bCombineScattering) F(mie_density) F(mie_scattering) F(mie_extinction) F(mie_phase_function_g) F(absorption_density)
        F(absorption_extinction) F(ground_albedo) F(mu_s_min) F(skySpectralRadianceToLumiance) F(sunSpectralRadianceToLumiance) F(earthCenterKm) END
    BEGIN(SkyLightInfo) F(direction) F(cascadeConfig) END
    BEGIN(GPUBasicData) F(blueNoiseCtx) F(atmosphere) F(sunInfo) F(frameCounter) F(frameCounterMod8) F(GLTFPrimitiveDetailBuffer) F(GLTFPrimitiveDataBuffer)
        F(GLTFMaterialBuffer) F(GLTFObjectCount) F(GLTFObjectBuffer) F(halton23_16tap) F(brdfLut) END
    BEGIN(PerframeCameraView) F(basicData) F(translatedWorldToView) F(viewToTranslatedWorld) F(viewToClip) F(clipToView) F(translatedWorldToClip)
        F(translatedWorldToClip_NoJitter) F(clipToTranslatedWorld) F(frustumPlane) F(translatedWorldToClipLastFrame) F(translatedWorldToClipLastFrame_NoJitter)
        F(clipToTranslatedWorld_LastFrame) F(frustumPlaneLastFrame) F(renderDimension) F(cameraToEarthCenter_km) F(cameraPositionWS_km) F(cameraWorldPos)
        F(cameraWorldPosLastFrame) F(clipToTranslatedWorldWithZFar_NoJitter) F(cameraFovy) F(zNear) F(zFar) F(bCameraCut) F(jitterData) F(forwardRS) END
    BEGIN(GPUObjectBasicData) F(localToTranslatedWorld) F(translatedWorldToLocal) F(localToTranslatedWorldLastFrame) F(scaleExtractFromMatrix) END
    BEGIN(GPUObjectGLTFPrimitive) F(basicData) F(GLTFPrimitiveDetail) F(GLTFMaterialData) END
    std::printf(",\n\"constants\": {\"kNaniteMeshletMaxVertices\": %d, \"kNaniteMeshletMaxTriangle\": %d, \"kNaniteMaxLODCount\": %d, \"kNaniteMaxBVHLevelCount\": %d, "
                "\"kNaniteBVHLevelNodeCount\": %d, \"kClusterGroupMergeMaxCount\": %d, \"kHZBMaxMipmapCount\": %d, \"kMaxInstanceIdCount\": %d, \"kMaxCascadeCount\": %d, "
                "\"subpixelPrecisionBits\": %d, \"kFrustumCullingEnableBit\": %d, \"kHZBCullingEnableBit\": %d, \"kMeshletConeCullEnableBit\": %d}\n}\n",
                kNaniteMeshletMaxVertices, kNaniteMeshletMaxTriangle, kNaniteMaxLODCount, kNaniteMaxBVHLevelCount, kNaniteBVHLevelNodeCount, kClusterGroupMergeMaxCount,
                kHZBMaxMipmapCount, kMaxInstanceIdCount, kMaxCascadeCount, subpixelPrecisionBits, kFrustumCullingEnableBit, kHZBCullingEnableBit, kMeshletConeCullEnableBit);
    return 0;
}
