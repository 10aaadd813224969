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
/* The filter chain of ONE frame sharded over the GPUs of a box (tile_gather.cu): this rank runs every stage on the rows of its band plus
   16 halo rows on either side (prepare +-16 -> a-trous 1 +-14 -> a-trous 2 +-10 -> a-trous 3 +-2 -> TAA +-1 -> present: the band), and the
   history buffers, which are read at reprojected positions anywhere in the frame, are read row by row from the rank that owns the row -
   loads over NVLink through the peer mappings below, issued by the filter kernels themselves. */
#define LH2B_MAX_SHARDS 8
struct FilterShard
{
	int world, rowsPerBand;			// row y of a history buffer lives on rank min( y / rowsPerBand, world - 1 ); rowsPerBand is a multiple of 16
	int rowFirst, rowEnd;				// rows every stage computes here (band + halo, clipped to the frame)
	int presentFirst, presentEnd;		// rows the present pass writes: the band
	int worldPosMargin;				// rows beyond the band for which this rank holds last frame's world positions itself (the renderers send a wider strip of this
									// input than the 16 halo rows: the diamond search of prepare wanders, and its dependent gathers should not cross NVLink)
	const float4* const* prevWorldPos, * const* prevMoments, * const* filteredIN, * const* prevPixels;	// DEVICE tables of 'world' pointers: every rank's buffer (peer mappings; [own rank] = the local buffer)
	float4* phase2Out;				// output of the second a-trous pass (the single-GPU chain reuses filteredIN for it)
};
/* SVGF / TAA chain (filter_kernels.cu). Buffers are float4[w*h] unless noted. */
struct FilterBuffers
{
	const float4* accumulator;	// [2*w*h]: direct, then indirect
	uint4* features; const float4* worldPos; const float4* prevWorldPos; const float4* deltaDepth;
	float4* shading; float2* motion; float4* moments; const float4* prevMoments;
	float4* filteredIN; float4* filteredOUT;	// IN holds last frame's phase-1 output on entry and this frame's phase-2 output on exit
	const float4* prevPixels; float4* taaOut; float4* target;
	const FilterShard* shard = nullptr;	// set: prevWorldPos / prevMoments / filteredIN (as history) / prevPixels above are ignored in favour of the per-rank tables
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
