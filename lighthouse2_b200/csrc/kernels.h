/* kernels.h - host-callable launchers of the CUDA stages (one .cu per stage family). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lh2b
{
struct DevScene;
struct RenderParams;
struct PathSet;

extern int g_traversalVariant, g_wideBlocksPerSM, g_triThreshold, g_refillThreshold, g_triThresholdShadow;
void LaunchExtend( const DevScene& scene, const float4* O4, const float4* D4, float4* hits, int n, uint32_t* workCounter, int smCount, cudaStream_t s );
void LaunchOcclude( const DevScene& scene, const float4* O4, const float4* D4, uint8_t* occluded, int n, uint32_t* workCounter, int smCount, cudaStream_t s );

void LaunchGenerateExtend( const DevScene& scene, const RenderParams& p, const PathSet& out, float4* hits, uint32_t* workCounter, int smCount, cudaStream_t s );
void LaunchExtendCounted( const DevScene& scene, const PathSet& in, float4* hits, const uint32_t* countPtr, uint32_t* workCounter, uint32_t maxRays, int smCount, cudaStream_t s );
void LaunchConnect( const DevScene& scene, const PathSet& conn, float4* accumulator, const uint32_t* countPtr, uint32_t* workCounter, uint32_t maxRays, int smCount, cudaStream_t s );
void LaunchShade( const RenderParams& p, const PathSet& in, const PathSet& out, const float4* hits, const PathSet& conn,
	int pathLength, uint32_t R0, bool useNEE, uint32_t maxPaths, int smCount, cudaStream_t s );
void LaunchTagTriangles( float4* tris, int triCount, uint32_t inst, cudaStream_t s );
void LaunchFinalize( const float4* accumulator, float4* out, int n, int samplesTaken, cudaStream_t s );

} // namespace lh2b
