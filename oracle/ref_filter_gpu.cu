/* ref_filter_gpu.cu - TEST INFRASTRUCTURE ONLY. Compiles the REFERENCE's own SVGF / TAA kernels for sm_100a straight
   from lib/CUDA/shared_kernel_code/finalize_shared.h (prepareFilterKernel :217-314, applyFilterKernel :320-484,
   TAApassKernel :498-548, unsharpenTAAKernel :554-583, finalizeNoTAAKernel :589-600) and runs the chain exactly as
   RenderCore::FinalizeRender of the filter core does (lib/RenderCore_Optix7Filter/rendercore.cpp:897-948).
   The header writes its final image through the legacy `surface<>` reference `renderTarget`, which CUDA 12 no longer
   has; here `renderTarget` is a cudaSurfaceObject_t bound to a float4 array, which the unmodified surf2Dwrite calls
   accept. Built by oracle/Makefile into oracle/_ref/libref_filter_gpu.so when /root/reference is present. */
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
#define FILTERINGCORE
#include "core_settings.h"
#define NEXTMULTIPLEOF(a,b) (((a)+((b)-1))&(0x7fffffff-((b)-1)))

namespace lh2core
{
__constant__ uchar4* argb32;
__constant__ float4* argb128;
__constant__ uchar4* nrm32;
__constant__ float4* skyPixels;
__constant__ int skywidth;
__constant__ int skyheight;
__constant__ float4* debugData;
static __device__ cudaSurfaceObject_t renderTarget;
#include "tools_shared.h"
#include "sampling_shared.h"
#include "finalize_shared.h"
} // namespace lh2core

struct RefFilterIO
{
	int w, h, samplesTaken, camIsStationary, taa;
	float directClamp, indirectClamp, j0, j1, prevj0, prevj1;
	float prevView[17];			// ViewPyramid of the previous frame
	// inputs (host)
	const float* accumulator;	// float4[2 * w * h]: direct, then indirect
	const uint* features;		// uint4[w * h] (in/out: history counter)
	const float* worldPos; const float* prevWorldPos; const float* deltaDepth;	// float4[w * h]
	const float* prevMoments;	// float4[w * h]
	const float* filteredIN;	// float4[w * h]: last frame's phase-1 output (history of the temporal blend)
	const float* prevPixels;	// float4[w * h]: last frame's TAA output
	// outputs (host)
	uint* featuresOut; float* shadingAfterPrepare; float* motion; float* moments;
	float* phase1; float* phase2; float* phase3; float* taaPixels; float* target;
	int timingRuns;				// > 0: after the checked run, time that many more runs of the whole chain on the same inputs
	float stageMs[8];			// mean ms: prepare, a-trous 1, 2, 3, TAA, unsharp / finalizeNoTAA, whole chain, 0 (CUDA events between the reference's own launches)
};

#define CK( x ) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf( stderr, "ref_filter_gpu: %s: %s\n", #x, cudaGetErrorString( e ) ); return 1; } } while (0)
template <typename T> static T* Up( const void* src, size_t bytes )
{
	void* d = nullptr;
	cudaMalloc( &d, bytes > 0 ? bytes : 16 );
	if (src) cudaMemcpy( d, src, bytes, cudaMemcpyHostToDevice ); else cudaMemset( d, 0, bytes );
	return (T*)d;
}

extern "C" __attribute__( ( visibility( "default" ) ) ) int reffilter_run( RefFilterIO* io )
{
	using namespace lh2core;
	const int w = io->w, h = io->h;
	const size_t px = (size_t)w * h, b16 = px * 16;
	float4* acc = Up<float4>( io->accumulator, 2 * b16 );
	uint4* feat = Up<uint4>( io->features, b16 );
	float4* wp = Up<float4>( io->worldPos, b16 ), * pwp = Up<float4>( io->prevWorldPos, b16 ), * dd = Up<float4>( io->deltaDepth, b16 );
	float4* shading = Up<float4>( nullptr, b16 ), * moments = Up<float4>( nullptr, b16 ), * pmom = Up<float4>( io->prevMoments, b16 );
	float2* motion = Up<float2>( nullptr, px * 8 );
	float4* fIN = Up<float4>( io->filteredIN, b16 ), * fOUT = Up<float4>( nullptr, b16 ), * prevPixels = Up<float4>( io->prevPixels, b16 );
	float4* dbg = Up<float4>( nullptr, b16 );
	CK( cudaMemcpyToSymbol( debugData, &dbg, sizeof( void* ) ) );
	cudaArray_t arr;
	cudaChannelFormatDesc cd = cudaCreateChannelDesc<float4>();
	CK( cudaMallocArray( &arr, &cd, w, h, cudaArraySurfaceLoadStore ) );
	std::vector<float> zero( px * 4, 0.0f );
	CK( cudaMemcpy2DToArray( arr, 0, 0, zero.data(), w * 16, w * 16, h, cudaMemcpyHostToDevice ) );
	cudaResourceDesc rd = {};
	rd.resType = cudaResourceTypeArray, rd.res.array.array = arr;
	cudaSurfaceObject_t surf;
	CK( cudaCreateSurfaceObject( &surf, &rd ) );
	CK( cudaMemcpyToSymbol( renderTarget, &surf, sizeof( surf ) ) );
	ViewPyramid pv;
	memcpy( &pv, io->prevView, 68 );
	// --- the chain, as RenderCore::FinalizeRender (filter core) ---
	prepareFilter( acc, feat, wp, pwp, shading, motion, moments, pmom, dd, pv, io->j0, io->j1, io->prevj0, io->prevj1,
		w, h, io->samplesTaken, io->directClamp, io->indirectClamp, io->camIsStationary );
	CK( cudaDeviceSynchronize() );
	CK( cudaMemcpy( io->shadingAfterPrepare, shading, b16, cudaMemcpyDeviceToHost ) );
	CK( cudaMemcpy( io->motion, motion, px * 8, cudaMemcpyDeviceToHost ) );
	CK( cudaMemcpy( io->moments, moments, b16, cudaMemcpyDeviceToHost ) );
	CK( cudaMemcpy( io->featuresOut, feat, b16, cudaMemcpyDeviceToHost ) );
	applyFilter( feat, pwp, wp, dd, motion, moments, shading, fIN, fOUT, w, h, 1, 0 );
	CK( cudaDeviceSynchronize() );
	CK( cudaMemcpy( io->phase1, fOUT, b16, cudaMemcpyDeviceToHost ) );
	applyFilter( feat, pwp, wp, dd, motion, moments, fOUT, 0, fIN, w, h, 2, 0 );
	CK( cudaDeviceSynchronize() );
	CK( cudaMemcpy( io->phase2, fIN, b16, cudaMemcpyDeviceToHost ) );
	applyFilter( feat, pwp, wp, dd, motion, moments, fIN, 0, shading, w, h, 3, 1 );
	CK( cudaDeviceSynchronize() );
	CK( cudaMemcpy( io->phase3, shading, b16, cudaMemcpyDeviceToHost ) );
	if (io->taa)
	{
		TAApass( shading, prevPixels, 0, 0, wp, pwp, motion, w, h );
		CK( cudaDeviceSynchronize() );
		CK( cudaMemcpy( io->taaPixels, shading, b16, cudaMemcpyDeviceToHost ) );
		unsharpenTAA( shading, w, h );
	}
	else finalizeNoTAA( shading, w, h );
	CK( cudaDeviceSynchronize() );
	CK( cudaMemcpy2DFromArray( io->target, w * 16, arr, 0, 0, w * 16, h, cudaMemcpyDeviceToHost ) );
	for (int k = 0; k < 8; k++) io->stageMs[k] = 0;
	if (io->timingRuns > 0)
	{
		// launch shapes are the reference's own (the host wrappers of finalize_shared.h, as lib/RenderCore_Optix7Filter/rendercore.cpp:897-948 calls them)
		uint4* feat0 = Up<uint4>( io->features, b16 );
		cudaEvent_t ev[7];
		for (auto& evt : ev) CK( cudaEventCreate( &evt ) );
		for (int r = 0; r < io->timingRuns + 1; r++)	// the first one warms up
		{
			CK( cudaMemcpy( feat, feat0, b16, cudaMemcpyDeviceToDevice ) );	// the history counter in the features is updated in place
			CK( cudaDeviceSynchronize() );
			CK( cudaEventRecord( ev[0] ) );
			prepareFilter( acc, feat, wp, pwp, shading, motion, moments, pmom, dd, pv, io->j0, io->j1, io->prevj0, io->prevj1,
				w, h, io->samplesTaken, io->directClamp, io->indirectClamp, io->camIsStationary );
			CK( cudaEventRecord( ev[1] ) );
			applyFilter( feat, pwp, wp, dd, motion, moments, shading, fIN, fOUT, w, h, 1, 0 );
			CK( cudaEventRecord( ev[2] ) );
			applyFilter( feat, pwp, wp, dd, motion, moments, fOUT, 0, fIN, w, h, 2, 0 );
			CK( cudaEventRecord( ev[3] ) );
			applyFilter( feat, pwp, wp, dd, motion, moments, fIN, 0, shading, w, h, 3, 1 );
			CK( cudaEventRecord( ev[4] ) );
			if (io->taa) TAApass( shading, prevPixels, 0, 0, wp, pwp, motion, w, h );
			CK( cudaEventRecord( ev[5] ) );
			if (io->taa) unsharpenTAA( shading, w, h ); else finalizeNoTAA( shading, w, h );
			CK( cudaEventRecord( ev[6] ) );
			CK( cudaEventSynchronize( ev[6] ) );
			if (r == 0) continue;
			for (int k = 0; k < 6; k++) { float ms = 0; CK( cudaEventElapsedTime( &ms, ev[k], ev[k + 1] ) ); io->stageMs[k] += ms / io->timingRuns; }
			float ms = 0;
			CK( cudaEventElapsedTime( &ms, ev[0], ev[6] ) );
			io->stageMs[6] += ms / io->timingRuns;
		}
		for (auto& evt : ev) cudaEventDestroy( evt );
		cudaFree( feat0 );
	}
	cudaDestroySurfaceObject( surf ), cudaFreeArray( arr );
	for (void* p : { (void*)acc, (void*)feat, (void*)wp, (void*)pwp, (void*)dd, (void*)shading, (void*)moments, (void*)pmom, (void*)motion, (void*)fIN, (void*)fOUT, (void*)prevPixels, (void*)dbg }) cudaFree( p );
	return 0;
}
