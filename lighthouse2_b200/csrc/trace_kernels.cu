/* trace_kernels.cu - standalone extend / connect kernels over ray buffers in HBM.

   extendKernel  = setupSecondaryRay   (lib/rendercore_optix7/optix/.optix.cu:131-140)
   occludeKernel = the traversal half of generateShadowRay (.optix.cu:142-149), reporting a flag per ray.
   Both read O4/D4 as float4 (32 B per ray) and write 16 B (hit) or 1 B (flag).
*/
#include "kernels.h"
#include "traverse.cuh"

namespace lh2b
{

__device__ __forceinline__ float4 PackHit( const bool hit, const TraceResult& r )
{
	if (!hit) return make_float4( 0, 0, __int_as_float( -1 ), 1e34f );
	const uint32_t uv = (uint32_t)(65535.0f * r.u) + ((uint32_t)(65535.0f * r.v) << 16);
	return make_float4( __uint_as_float( uv ), __uint_as_float( r.inst ), __uint_as_float( r.prim ), r.t );
}

__global__ void __launch_bounds__( 128 ) extendKernel( const DevScene scene, const float4* __restrict__ O4, const float4* __restrict__ D4,
	float4* __restrict__ hits, const int n )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 o = O4[i], d = D4[i];
	TraceResult r;
	const bool hit = Traverse<false>( scene, make_float3( o.x, o.y, o.z ), make_float3( d.x, d.y, d.z ), 0.0f, 1e34f, r );
	hits[i] = PackHit( hit, r );
}

__global__ void __launch_bounds__( 128 ) occludeKernel( const DevScene scene, const float4* __restrict__ O4, const float4* __restrict__ D4,
	uint8_t* __restrict__ occluded, const int n )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 o = O4[i], d = D4[i];
	TraceResult r;
	occluded[i] = Traverse<true>( scene, make_float3( o.x, o.y, o.z ), make_float3( d.x, d.y, d.z ), 0.0f, d.w, r ) ? 1 : 0;
}

void LaunchExtend( const DevScene& scene, const float4* O4, const float4* D4, float4* hits, int n, cudaStream_t s )
{
	if (n <= 0) return;
	extendKernel<<<(n + 127) / 128, 128, 0, s>>>( scene, O4, D4, hits, n );
}

void LaunchOcclude( const DevScene& scene, const float4* O4, const float4* D4, uint8_t* occluded, int n, cudaStream_t s )
{
	if (n <= 0) return;
	occludeKernel<<<(n + 127) / 128, 128, 0, s>>>( scene, O4, D4, occluded, n );
}

} // namespace lh2b
