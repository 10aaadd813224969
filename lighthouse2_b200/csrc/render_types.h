/* render_types.h - launch constants and buffer views of the wavefront stages.

   Counterpart of the reference's Params (lib/rendercore_optix7/core_settings.h:113-133), Counters
   (core_settings.h:96-110) and the __constant__ globals of lib/rendercore_optix7/kernels/.cuda.cu:22-43,
   passed by value as kernel arguments instead of staged constant-memory copies.

   Path state layout in HBM (SoA float4), per ray i:
     O[i]  = origin.xyz, w = (pathIdx << 6) | flags      (core_settings.h:71-79, .optix.cu:121)
     D[i]  = direction.xyz, w = packed normal of the previous vertex (pathtracer.h:236)
     T[i]  = throughput.rgb, w = postponed bsdf pdf       (pathtracer.h:237)
     hit[i]= (u16|v16<<16, instance, primitive, t)        (.optix.cu:174-184)
   Shade reads set 'in' and writes the compacted extension rays to set 'out' (ping-pong: the
   reference compacts in place, which races - SURVEY.md 0.7).
   Connections: cO = origin, cD = L.xyz + tmax, cE = potential radiance.rgb + pixel index bits.
*/
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace lh2b
{

#define LH2B_MAXPATHLENGTH 16	// upper bound for the runtime setting "maxPathLength" (reference: compile-time 3)

// path state flags (lib/rendercore_optix7/kernels/pathtracer.h:29-33)
#define S_SPECULAR      1
#define S_BOUNCED       2
#define S_VIASPECULAR   4
#define S_BOUNCEDTWICE  8

struct DevCounters
{
	uint32_t extensionRays[LH2B_MAXPATHLENGTH + 2];	// [L] = rays produced by shade at path length L
	uint32_t shadowRays[LH2B_MAXPATHLENGTH + 2];		// [L] = shadow rays produced by shade at path length L
	uint32_t workFetch[2 * LH2B_MAXPATHLENGTH + 8];	// dynamic work counters of the persistent kernels
	// probe: written by one path per frame with a single 16-byte store (keep the four words together and 16-byte aligned); NOT cleared
	// per frame - like the reference's counters (kernels/.cuda.cu:191-202 never touches them) a frame without a probe hit keeps the last one
	int32_t probedInstid, probedTriid;
	float probedDist;
	uint32_t pad;
};
static_assert( offsetof( DevCounters, probedInstid ) % 16 == 0, "probe words must be 16-byte aligned" );

struct PathSet { float4* O; float4* D; float4* T; };

/* CUDAMaterial as 8 x uint4 (core_settings.h:136-185): baseData4, parameters, tex0, tex1, nmap0, nmap1, smap, rmap */
struct DevMaterial { uint4 q[8]; };

struct RenderParams
{
	// camera (Params, core_settings.h:122-127; filled as rendercore.cpp:857-864)
	float4 posLensSize;
	float3 right, up, p1;
	float distortion, spreadAngle;
	int w, h, spp;
	int pass;					// samplesTaken before this frame
	uint32_t shift;				// blue-noise tile shift (rendercore.cpp:855,860)
	uint32_t sampleBase;		// first sample index of this core's shard (multi-GPU), 0 otherwise
	uint32_t stride;			// paths of this frame on this core: w * (bandY1 - bandY0) * spp
	int bandY0, bandY1, bandStep;	// rows rendered by this core (tile-sharded frames, lh2b_set_row_band): the 4-row tile rows bandY0/4 + j * bandStep
								// below row bandY1; the whole frame: 0, h, 1
	float geometryEpsilon, clampValue;
	int probePixelIdx;
	int maxPathLength;			// reference MAXPATHLENGTH (3)
	int bsdfModel;				// 0: lambert.h (Lambert + pure specular + dielectric), 1: disney.h (principled); Setting "bsdf"
	uint32_t enoughBounces;		// reference ENOUGH_BOUNCES flag mask (S_BOUNCED); 0 = never stop on bounce count
	// scene (the __constant__ block of kernels/.cuda.cu:22-43)
	const void* instDesc;		// CoreInstanceDesc[]
	const DevMaterial* materials;
	const float4* triLights;	// CoreLightTri4: 6 float4 each
	const float4* pointLights;	// 2 float4 each
	const float4* spotLights;	// 3 float4 each
	const float4* dirLights;	// 2 float4 each
	int4 lightCounts;
	const uchar4* argb32;
	const float4* argb128;
	const uchar4* nrm32;
	const float4* skyPixels;
	int skyW, skyH;
	float worldToSky[12];
	const uint32_t* blueNoise;
	float4* accumulator;
	DevCounters* counters;
	// filter mode (Setting "filter" = 1): per-pixel features of the first diffuse vertex for the SVGF chain, and the
	// accumulator split into direct [0, w*h) and indirect [w*h, 2*w*h) light (lib/RenderCore_Optix7Filter/kernels/pathtracer.h:44-58,95-125,195-234,286-300)
	uint4* features;			// null: filter off
	float4* worldPos;
	float4* deltaDepth;
};

/* number of 4-row tile rows this core renders (tile-sharded frames) */
static __host__ __device__ inline uint32_t BandTileRows( const RenderParams& p )
{
	const uint32_t first = (uint32_t)p.bandY0 / 4, end = ((uint32_t)p.bandY1 + 3) / 4;
	return end > first ? (end - first + (uint32_t)p.bandStep - 1) / (uint32_t)p.bandStep : 0;
}


} // namespace lh2b
