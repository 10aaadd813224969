/* kernels.h - host-callable launchers of the CUDA stages (one .cu per stage family). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lh2b
{
struct DevScene;
struct RenderParams;
struct PathSet;

extern int g_wideBlocksPerSM, g_triThreshold, g_refillThreshold, g_triThresholdShadow, g_raysPerLane;
void LaunchExtend( const DevScene& scene, const float4* O4, const float4* D4, float4* hits, int n, uint32_t* workCounter, int smCount, cudaStream_t s );
void LaunchOcclude( const DevScene& scene, const float4* O4, const float4* D4, uint8_t* occluded, int n, uint32_t* workCounter, int smCount, cudaStream_t s );

void LaunchGenerateExtend( const DevScene& scene, const RenderParams& p, const PathSet& out, float4* hits, uint32_t* workCounter, int smCount, cudaStream_t s );
void LaunchExtendCounted( const DevScene& scene, const PathSet& in, float4* hits, const uint32_t* countPtr, uint32_t* workCounter, uint32_t maxRays, int smCount, cudaStream_t s );
void LaunchConnect( const DevScene& scene, const PathSet& conn, float4* accumulator, const uint32_t* countPtr, uint32_t* workCounter, uint32_t maxRays, int smCount, cudaStream_t s );
void LaunchShade( const RenderParams& p, const PathSet& in, const PathSet& out, const float4* hits, const PathSet& conn,
	int pathLength, uint32_t R0, bool useNEE, uint32_t maxPaths, int smCount, cudaStream_t s );
/* SVGF / TAA chain (filter_kernels.cu). Buffers are float4[w*h] unless noted. */
struct FilterBuffers
{
	const float4* accumulator;	// [2*w*h]: direct, then indirect
	uint4* features; const float4* worldPos; const float4* prevWorldPos; const float4* deltaDepth;
	float4* shading; float2* motion; float4* moments; const float4* prevMoments;
	float4* filteredIN; float4* filteredOUT;	// IN holds last frame's phase-1 output on entry and this frame's phase-2 output on exit
	const float4* prevPixels; float4* taaOut; float4* target;
};
struct FilterSettings
{
	int w, h, samplesTaken, camIsStationary, taa;
	float directClamp, indirectClamp, j0, j1, prevj0, prevj1;
	float prevView[17];
};
void LaunchFilterChain( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, cudaEvent_t* stageEvents = nullptr );	// stageEvents: 7 events recorded at the stage boundaries
void LaunchFilterChainStaged( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, float* hPrepare, float* hP1, float* hP2, float* hP3 );
void LaunchTagTriangles( float4* tris, int triCount, uint32_t inst, cudaStream_t s );
// skin_kernels.cu
void LaunchCaptureBindPose( const float4* coreTris, float4* bindNormals, int triCount, cudaStream_t s );
void LaunchSkin( const float4* bindVerts, const float4* bindNormals, const uint4* joints, const float4* weights, const float4* jointMats, int jointCount,
	float4* verts, float4* coreTris, int triCount, cudaStream_t s );
void LaunchMorph( const float4* bindVerts, const float4* bindNormals, const float4* deltas, const float4* normals, const float* weights, int targetCount,
	float4* verts, float4* coreTris, int triCount, cudaStream_t s );
void LaunchFinalize( const float4* accumulator, float4* out, int n, int samplesTaken, cudaStream_t s );
// the IEEE / libm-accurate builds of the shade and filter stages (Setting "preciseMath", see shade_kernels.cu)
void LaunchShadePrecise( const RenderParams& p, const PathSet& in, const PathSet& out, const float4* hits, const PathSet& conn,
	int pathLength, uint32_t R0, bool useNEE, uint32_t maxPaths, int smCount, cudaStream_t s );
void LaunchFilterChainPrecise( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, cudaEvent_t* stageEvents = nullptr );
void LaunchFilterChainStagedPrecise( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, float* hPrepare, float* hP1, float* hP2, float* hP3 );

} // namespace lh2b
