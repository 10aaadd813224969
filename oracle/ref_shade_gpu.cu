/* ref_shade_gpu.cu - TEST INFRASTRUCTURE ONLY. Harness that compiles the REFERENCE's own shading stack for
   sm_100a straight from the reference tree (nothing is copied): lib/rendercore_optix7/kernels/pathtracer.h
   (shadeKernel) with lib/CUDA/shared_kernel_code/{tools,sampling,material,lights}_shared.h and
   lib/sharedBSDFs/{compatibility,lambert}.h - or, with -DREF_BSDF_DISNEY, {ggxmdf,frosted,disney}.h (second library). It re-declares the __constant__ globals that
   lib/rendercore_optix7/kernels/.cuda.cu:22-43,190 owns (that file cannot be compiled with CUDA 12: it pulls in the
   legacy surface reference of .cuda.h:54) and exposes one C entry point that runs shadeKernel on caller data.
   Built by oracle/Makefile into oracle/_ref/libref_shade_gpu.so when /root/reference is present.

   The reference kernel compacts extension rays into the buffer it is reading (pathtracer.h:65-67 vs :234-237).
   To get a deterministic answer from the UNMODIFIED kernel the harness starts the extension counter at
   'pathCount' and uses stride >= 2 * pathCount, so writes land in [pathCount, 2*pathCount) and never overlap the reads.
*/
#include <vector>
#include <cstdio>
#include <cstring>
typedef unsigned int uint;
typedef unsigned char uchar;
#define LH2_DEVFUNC static __forceinline__ __device__
#include "helper_math.h"
#include "cuda_fp16.h"
#include "common_settings.h"
#include "common_classes.h"
#include "common_functions.h"
#include "common_types.h"
#include "core_settings.h"
#define THREADMASK __activemask()
#define NEXTMULTIPLEOF(a,b) (((a)+((b)-1))&(0x7fffffff-((b)-1)))

namespace lh2core
{
__constant__ CoreInstanceDesc* instanceDescriptors;
__constant__ CUDAMaterial* materials;
__constant__ CoreLightTri* triLights;
__constant__ CorePointLight* pointLights;
__constant__ CoreSpotLight* spotLights;
__constant__ CoreDirectionalLight* directionalLights;
__constant__ int4 lightCounts;
__constant__ uchar4* argb32;
__constant__ float4* argb128;
__constant__ uchar4* nrm32;
__constant__ float4* skyPixels;
__constant__ int skywidth;
__constant__ int skyheight;
__constant__ PathState* pathStates;
__constant__ float4* debugData;
__constant__ LightCluster* lightTree;
__constant__ mat4 worldToSky;
__constant__ __device__ float geometryEpsilon;
__constant__ __device__ float clampValue;
static __device__ Counters* counters;

#include "tools_shared.h"
#include "sampling_shared.h"
#include "material_shared.h"
#include "lights_shared.h"
#include "compatibility.h"
#ifdef REF_BSDF_DISNEY	// what kernels/bsdf.h:7-21 of the stock cores selects
#include "ggxmdf.h"
#include "frosted.h"
#include "disney.h"
#else
#include "lambert.h"
#endif
#include "pathtracer.h"
} // namespace lh2core

struct RefShadeIn
{
	int meshCount; const void* const* coreTris; const int* triCounts;
	int instanceCount; const int* instMesh; const float* instInverse16;	// 4x4 inverse per instance, row major
	const void* materials128; int materialCount;
	const void* triLights; int triLightCount; const void* pointLights; int pointLightCount;
	const void* spotLights; int spotLightCount; const void* dirLights; int dirLightCount;
	const void* argb32; int argb32Count; const void* argb128; int argb128Count; const void* nrm32; int nrm32Count;
	const void* skyPixels4; int skyPixelCount, skyW, skyH; float worldToSky[16];
	const uint* blueNoise;	// 5 * 65536
	float geometryEpsilon, clampValue;
	int pathCount, stride;	// stride >= 2 * pathCount
	float* pathStates;		// float4[3 * stride], in/out
	float* hits;			// float4[stride], in
	float* connections;		// float4[6 * stride], out
	float* accumulator;		// float4[w * h], in/out
	uint R0, shift; int pass, probePixelIdx, pathLength, w, h; float spreadAngle; int useNEE;
	uint countersOut[12];
	int timingRuns;			// > 0: after the checked run, time that many more launches of the kernel on the same inputs
	float timingMsMin, timingMsMean;	// CUDA events around the reference's own shade() launch (grid ceil(n/128), block 128: pathtracer.h:244-252)
};

#define CK( x ) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf( stderr, "ref_shade_gpu: %s: %s\n", #x, cudaGetErrorString( e ) ); return 1; } } while (0)
template <typename T> static T* Up( const void* src, size_t bytes )
{
	void* d = nullptr;
	cudaMalloc( &d, bytes > 0 ? bytes : 16 );
	if (bytes) cudaMemcpy( d, src, bytes, cudaMemcpyHostToDevice );
	return (T*)d;
}

extern "C" __attribute__( ( visibility( "default" ) ) ) int refshade_run( RefShadeIn* in )
{
	using namespace lh2core;
	std::vector<void*> owned;
	std::vector<CoreTri4*> dTris( in->meshCount );
	for (int i = 0; i < in->meshCount; i++) dTris[i] = Up<CoreTri4>( in->coreTris[i], (size_t)in->triCounts[i] * sizeof( CoreTri ) ), owned.push_back( dTris[i] );
	std::vector<CoreInstanceDesc> desc( in->instanceCount );
	for (int i = 0; i < in->instanceCount; i++)
	{
		desc[i].triangles = dTris[in->instMesh[i]];
		memcpy( &desc[i].invTransform, in->instInverse16 + i * 16, 64 );
	}
	CoreInstanceDesc* dDesc = Up<CoreInstanceDesc>( desc.data(), desc.size() * sizeof( CoreInstanceDesc ) ); owned.push_back( dDesc );
	CUDAMaterial* dMat = Up<CUDAMaterial>( in->materials128, (size_t)in->materialCount * 128 ); owned.push_back( dMat );
	CoreLightTri* dTL = Up<CoreLightTri>( in->triLights, (size_t)in->triLightCount * sizeof( CoreLightTri ) ); owned.push_back( dTL );
	CorePointLight* dPL = Up<CorePointLight>( in->pointLights, (size_t)in->pointLightCount * sizeof( CorePointLight ) ); owned.push_back( dPL );
	CoreSpotLight* dSL = Up<CoreSpotLight>( in->spotLights, (size_t)in->spotLightCount * sizeof( CoreSpotLight ) ); owned.push_back( dSL );
	CoreDirectionalLight* dDL = Up<CoreDirectionalLight>( in->dirLights, (size_t)in->dirLightCount * sizeof( CoreDirectionalLight ) ); owned.push_back( dDL );
	uchar4* d32 = Up<uchar4>( in->argb32, (size_t)in->argb32Count * 4 ); owned.push_back( d32 );
	float4* d128 = Up<float4>( in->argb128, (size_t)in->argb128Count * 16 ); owned.push_back( d128 );
	uchar4* dN = Up<uchar4>( in->nrm32, (size_t)in->nrm32Count * 4 ); owned.push_back( dN );
	float4* dSky = Up<float4>( in->skyPixels4, (size_t)in->skyPixelCount * 16 ); owned.push_back( dSky );
	uint* dBN = Up<uint>( in->blueNoise, (65536 * 5 + 16) * 4 ); owned.push_back( dBN );	// 16 zero words of padding, see binding.py
	const int4 lc = make_int4( in->triLightCount, in->pointLightCount, in->spotLightCount, in->dirLightCount );
	mat4 w2s;
	memcpy( &w2s, in->worldToSky, 64 );
	CK( cudaMemcpyToSymbol( instanceDescriptors, &dDesc, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( materials, &dMat, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( triLights, &dTL, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( pointLights, &dPL, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( spotLights, &dSL, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( directionalLights, &dDL, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( lightCounts, &lc, sizeof( int4 ) ) );
	CK( cudaMemcpyToSymbol( argb32, &d32, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( argb128, &d128, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( nrm32, &dN, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( skyPixels, &dSky, sizeof( void* ) ) );
	CK( cudaMemcpyToSymbol( skywidth, &in->skyW, sizeof( int ) ) );
	CK( cudaMemcpyToSymbol( skyheight, &in->skyH, sizeof( int ) ) );
	CK( cudaMemcpyToSymbol( worldToSky, &w2s, sizeof( mat4 ) ) );
	CK( cudaMemcpyToSymbol( geometryEpsilon, &in->geometryEpsilon, sizeof( float ) ) );
	CK( cudaMemcpyToSymbol( clampValue, &in->clampValue, sizeof( float ) ) );
	Counters hc;
	memset( &hc, 0, sizeof( hc ) );
	hc.extensionRays = in->pathCount;	// see header: keeps the in-place compaction from overlapping its input
	hc.probedTriid = -1;
	Counters* dC = Up<Counters>( &hc, sizeof( Counters ) ); owned.push_back( dC );
	CK( cudaMemcpyToSymbol( counters, &dC, sizeof( void* ) ) );
	const size_t stride = in->stride;
	float4* dPS = Up<float4>( in->pathStates, stride * 3 * 16 ); owned.push_back( dPS );
	float4* dHits = Up<float4>( in->hits, stride * 16 ); owned.push_back( dHits );
	float4* dPS0 = nullptr, * dHits0 = nullptr;	// pristine inputs for the timed launches (the kernel overwrites both)
	if (in->timingRuns > 0)
	{
		dPS0 = Up<float4>( in->pathStates, stride * 3 * 16 ), owned.push_back( dPS0 );
		dHits0 = Up<float4>( in->hits, stride * 16 ), owned.push_back( dHits0 );
	}
	const Counters hc0 = hc;
	float4* dConn = Up<float4>( nullptr, 0 );
	cudaFree( dConn );
	CK( cudaMalloc( &dConn, stride * 6 * 16 ) ); owned.push_back( dConn );
	CK( cudaMemset( dConn, 0, stride * 6 * 16 ) );
	float4* dAcc = Up<float4>( in->accumulator, (size_t)in->w * in->h * 16 ); owned.push_back( dAcc );
	shade( in->pathCount, dAcc, (uint)stride, dPS, dHits, in->useNEE ? dConn : 0, in->R0, in->shift, dBN, in->pass,
		in->probePixelIdx, in->pathLength, in->w, in->h, in->spreadAngle );
	CK( cudaGetLastError() );
	CK( cudaDeviceSynchronize() );
	CK( cudaMemcpy( in->pathStates, dPS, stride * 3 * 16, cudaMemcpyDeviceToHost ) );
	CK( cudaMemcpy( in->connections, dConn, stride * 6 * 16, cudaMemcpyDeviceToHost ) );
	CK( cudaMemcpy( in->accumulator, dAcc, (size_t)in->w * in->h * 16, cudaMemcpyDeviceToHost ) );
	CK( cudaMemcpy( &hc, dC, sizeof( Counters ), cudaMemcpyDeviceToHost ) );
	memcpy( in->countersOut, &hc, sizeof( Counters ) );
	in->timingMsMin = in->timingMsMean = 0;
	if (in->timingRuns > 0)
	{
		cudaEvent_t e0, e1;
		CK( cudaEventCreate( &e0 ) ); CK( cudaEventCreate( &e1 ) );
		float best = 1e30f, sum = 0;
		for (int r = 0; r < in->timingRuns + 1; r++)	// the first one warms up
		{
			CK( cudaMemcpy( dPS, dPS0, stride * 3 * 16, cudaMemcpyDeviceToDevice ) );
			CK( cudaMemcpy( dHits, dHits0, stride * 16, cudaMemcpyDeviceToDevice ) );
			CK( cudaMemcpy( dC, &hc0, sizeof( Counters ), cudaMemcpyHostToDevice ) );
			CK( cudaDeviceSynchronize() );
			CK( cudaEventRecord( e0 ) );
			shade( in->pathCount, dAcc, (uint)stride, dPS, dHits, in->useNEE ? dConn : 0, in->R0, in->shift, dBN, in->pass,
				in->probePixelIdx, in->pathLength, in->w, in->h, in->spreadAngle );
			CK( cudaEventRecord( e1 ) );
			CK( cudaEventSynchronize( e1 ) );
			float ms = 0;
			CK( cudaEventElapsedTime( &ms, e0, e1 ) );
			if (r > 0) best = ms < best ? ms : best, sum += ms;
		}
		in->timingMsMin = best, in->timingMsMean = sum / in->timingRuns;
		cudaEventDestroy( e0 ), cudaEventDestroy( e1 );
	}
	for (void* p : owned) cudaFree( p );
	return 0;
}
