/* filter_kernels.cu - the SVGF temporal + a-trous filter and the TAA pass as shared-memory-tiled kernels (sm_100a, built
   with --use_fast_math like the reference).

   Restates lib/CUDA/shared_kernel_code/finalize_shared.h of the reference's Optix7Filter core:
     prepareFilterKernel :217-314 (albedo demodulation, clamps, reprojection incl. the 25-step diamond search for
                                   specular pixels, luminance moments with history)         -> prepareKernel
     applyFilterKernel   :320-484 (5x5-minus-corners a-trous at step 1,2,4 with normal^128, depth-gradient, luminance
                                   variance and albedo/material weights; phase 1 adds the temporal blend with a YCoCg
                                   neighbourhood clamp; the last pass remodulates and takes the square root)  -> atrousKernel
     TAApassKernel       :498-548 (Mitchell-Netravali history fetch, variance clip, 0.1/0.9 blend)  -> taaKernel
     unsharpenTAAKernel  :554-583 / finalizeNoTAAKernel :589-600                                     -> presentKernel
   plus the helpers they use from sampling_shared.h:111-215 and tools_shared.h:122-177,237-262.
   The reference reads every tap from global memory (its block is 32x2). Here each block stages its tile plus the
   a-trous halo (2 * step pixels on every side) of the shading and feature buffers in shared memory once, so the
   21 taps per pixel come from on-chip memory; history lookups stay gathers.
   Parity is checked stage by stage against the reference's own kernels (oracle/ref_filter_gpu.cu).
*/
#include "kernels.h"
#include "common.cuh"
#include <algorithm>

/* Second build (build/filter_kernels_precise.o, -DLH2B_PRECISE, no --use_fast_math, -fmad=false), selected by Setting "preciseMath" 1:
   see shade_kernels.cu. */
#ifdef LH2B_PRECISE
#define LaunchFilterChain LaunchFilterChainPrecise
#define LaunchFilterChainStaged LaunchFilterChainStagedPrecise
#define prepareKernel prepareKernelPrecise
#define prepareSearchKernel prepareSearchKernelPrecise
#define prepareFinishKernel prepareFinishKernelPrecise
#define atrousKernel atrousKernelPrecise
#define taaKernel taaKernelPrecise
#define presentKernel presentKernelPrecise
#define __expf expf
#endif

namespace lh2b
{

__device__ __forceinline__ float3 min3f( const float3 a, const float b ) { return make_float3( fminf( a.x, b ), fminf( a.y, b ), fminf( a.z, b ) ); }
__device__ __forceinline__ float3 max3v( const float3 a, const float3 b ) { return make_float3( fmaxf( a.x, b.x ), fmaxf( a.y, b.y ), fmaxf( a.z, b.z ) ); }
__device__ __forceinline__ float3 clamp3( const float3 v, const float3 lo, const float3 hi ) { return make_float3( fminf( fmaxf( v.x, lo.x ), hi.x ), fminf( fmaxf( v.y, lo.y ), hi.y ), fminf( fmaxf( v.z, lo.z ), hi.z ) ); }
__device__ __forceinline__ float sqrLen( const float3 a ) { return dot( a, a ); }
__device__ __forceinline__ float sqrf( const float x ) { return x * x; }
__device__ __forceinline__ float4 operator-( const float4 a, const float4 b ) { return make_float4( a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w ); }
__device__ __forceinline__ float4 operator+( const float4 a, const float4 b ) { return make_float4( a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w ); }
__device__ __forceinline__ float4 operator*( const float4 a, const float s ) { return make_float4( a.x * s, a.y * s, a.z * s, a.w * s ); }
__device__ __forceinline__ float4 operator*( const float s, const float4 a ) { return make_float4( a.x * s, a.y * s, a.z * s, a.w * s ); }

/* tools_shared.h:122-177 */
__device__ __forceinline__ float3 UnpackNormal2( const uint32_t pi )
{
	const uint32_t x = (pi >> 2u) & 1023u, y = (pi >> 12u) & 1023u, z = pi >> 22u;
	return make_float3( x * (1.0f / 511.0f) - 1, y * (1.0f / 511.0f) - 1, z * (1.0f / 511.0f) - 1 );
}
__device__ __forceinline__ float3 RGBToYCoCg( const float3 RGB )
{
	const float3 rgb = min3f( RGB, 4.0f );
	const float Y = (rgb.x + 2 * rgb.y + rgb.z) * 0.25f;
	const float Co = (2 * rgb.x - 2 * rgb.z) * 0.25f + (0.5f * 256.0f / 255.0f);
	const float Cg = (-rgb.x + 2 * rgb.y - rgb.z) * 0.25f + (0.5f * 256.0f / 255.0f);
	return make_float3( Y, Co, Cg );
}
__device__ __forceinline__ float3 YCoCgToRGB( const float3 c )
{
	const float Y = c.x, Co = c.y - (0.5f * 256.0f / 255.0f), Cg = c.z - (0.5f * 256.0f / 255.0f);
	return make_float3( Y + Co - Cg, Y + Cg, Y - Co - Cg );
}
__device__ __forceinline__ float Luminance( const float3 rgb ) { return 0.299f * fminf( rgb.x, 10.0f ) + 0.587f * fminf( rgb.y, 10.0f ) + 0.114f * fminf( rgb.z, 10.0f ); }
__device__ __forceinline__ float3 RGB32toHDR( const uint32_t c )
{
	return make_float3( (float)(c >> 22) * (1.0f / 1023.0f), (float)((c >> 11) & 2047) * (1.0f / 2047.0f), (float)(c & 2047) * (1.0f / 2047.0f) );
}
__device__ __forceinline__ float3 RGB32toHDRmin1( const uint32_t c )
{
	return make_float3( (float)max( 1u, c >> 22 ) * (1.0f / 1023.0f), (float)max( 1u, (c >> 11) & 2047 ) * (1.0f / 2047.0f), (float)max( 1u, c & 2047 ) * (1.0f / 2047.0f) );
}
/* tools_shared.h:237-262: two rgb triples in 5.11 fixed point */
__device__ __forceinline__ float4 CombineToFloat4( const float3 A, const float3 B )
{
	const uint32_t Ar = (uint32_t)(fminf( A.x, 31.999f ) * 2048.0f), Ag = (uint32_t)(fminf( A.y, 31.999f ) * 2048.0f), Ab = (uint32_t)(fminf( A.z, 31.999f ) * 2048.0f);
	const uint32_t Br = (uint32_t)(fminf( B.x, 31.999f ) * 2048.0f), Bg = (uint32_t)(fminf( B.y, 31.999f ) * 2048.0f), Bb = (uint32_t)(fminf( B.z, 31.999f ) * 2048.0f);
	return make_float4( __uint_as_float( (Ar << 16) + Ag ), __uint_as_float( Ab ), __uint_as_float( (Br << 16) + Bg ), __uint_as_float( Bb ) );
}
__device__ __forceinline__ float3 DirectOf( const float4 X )
{
	const uint32_t v0 = __float_as_uint( X.x ), v1 = __float_as_uint( X.y );
	return make_float3( (float)(v0 >> 16) * (1.0f / 2048.0f), (float)(v0 & 65535) * (1.0f / 2048.0f), (float)v1 * (1.0f / 2048.0f) );
}
__device__ __forceinline__ float3 IndirectOf( const float4 X )
{
	const uint32_t v2 = __float_as_uint( X.z ), v3 = __float_as_uint( X.w );
	return make_float3( (float)(v2 >> 16) * (1.0f / 2048.0f), (float)(v2 & 65535) * (1.0f / 2048.0f), (float)v3 * (1.0f / 2048.0f) );
}
__device__ __forceinline__ float OneOverPow2( const int p ) { return __uint_as_float( (uint32_t)(127 - p) << 23 ); }
__device__ __forceinline__ float MitchellNetravali( const float v )
{
	const float B = 1.0f / 3.0f, C = 1.0f / 3.0f, x = fabsf( v ), x2 = x * x, x3 = x2 * x;
	if (x < 1) return (1.0f / 6.0f) * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B));
	else if (x < 2) return 1.0f / 6.0f * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C));
	return 0.0f;
}

/* History buffers (the previous frame's world positions, moments, phase-1 output and TAA image) are read at reprojected positions,
   i.e. anywhere in the frame. When the filter chain of one frame is sharded over the GPUs of a box (FilterShard, kernels.h; tile_gather.cu)
   every rank keeps the history rows of its own band and the kernels read a row from the rank that owns it - plain loads through the
   peer mappings over NVLink. On one GPU the same code runs with lo = 0, hi = h (every row is local): ONE instantiation of every kernel
   serves both cases, because the sharded frame has to equal the single-GPU frame bit for bit and two instantiations of the same source
   need not round alike (measured: a second instantiation of the TAA kernel differed by 1-10 ulp where the 4x4 history footprint touches
   the left image edge - the compiler had versioned that loop differently). The price on one GPU is a compare and a select per history read. */
struct HistBuf { const float4* local; const float4* const* perRank; int lo; uint32_t span; };	// rows [lo, lo + span) of this rank's own copy are valid (its band plus the halo rows
																						// it computes itself); perRank: device table of every rank's buffer (null on one GPU)
struct ShardMap { uint32_t magic; int last; };	// owner of row y: min( y / rowsPerBand, last ), the division as a multiplication by magic = 2^32 / rowsPerBand + 1
/* The common case - a row this rank holds itself - must stay as cheap as the plain load: measured on C5 (one GPU, where every row is
   local), the owner arithmetic plus an indexed constant load in front of every gather cost 0.4 ms of the 3.3 ms chain, an out-of-line
   call for the rare case 0.5 ms (the 16 gathers of a search step no longer overlap). Here the owner lookup sits under the predicate of
   the rare case and reads its table from global memory; the local address needs one subtract and one compare. */
/* row y of a history buffer (the test is per row: the lookups below fetch two to four texels from each row they touch) */
__device__ __forceinline__ const float4* HistRow( const HistBuf& b, const ShardMap& m, const int y, const int w )
{
	// halo rows are computed on both sides of a band boundary with identical results: whatever this rank computed itself it reads locally
	const float4* base = b.local;
	if ((uint32_t)(y - b.lo) >= b.span) base = (const float4*)__ldg( (const unsigned long long*)b.perRank + min( (int)__umulhi( (uint32_t)y, m.magic ), m.last ) );
	return base + y * w;
}
__device__ __forceinline__ float4 HistAt( const HistBuf& b, const ShardMap& m, const int x, const int y, const int w ) { return __ldg( HistRow( b, m, y, w ) + x ); }

/* sampling_shared.h:111-215 */
__device__ __forceinline__ float4 ReadTexelConsistent( const HistBuf& buffer, const HistBuf& prevWorldPos, const ShardMap& m, const float4 localPos,
	const float3 localNormal, const float u, const float v, const int w, const int h )
{
	const int iu1 = (int)floorf( u ), iv1 = (int)floorf( v ), iu0 = max( 0, iu1 - 1 ), iv0 = max( 0, iv1 - 1 );
	if (iu1 >= w || iv1 >= h || iu1 < 0 || iv1 < 0) return make_float4( -1, -1, -1, -1 );
	const float fx = u - floorf( u ), fy = v - floorf( v );
	const float4* b0 = HistRow( buffer, m, iv0, w ), * b1 = HistRow( buffer, m, iv1, w ), * q0 = HistRow( prevWorldPos, m, iv0, w ), * q1 = HistRow( prevWorldPos, m, iv1, w );
	const float4 p0 = __ldg( b0 + iu0 ), p1 = __ldg( b0 + iu1 ), p2 = __ldg( b1 + iu0 ), p3 = __ldg( b1 + iu1 );
	const uint32_t n0 = __float_as_uint( __ldg( q0 + iu0 ).w ), n1 = __float_as_uint( __ldg( q0 + iu1 ).w );
	const uint32_t n2 = __float_as_uint( __ldg( q1 + iu0 ).w ), n3 = __float_as_uint( __ldg( q1 + iu1 ).w );
	const uint32_t spec = __float_as_uint( localPos.w ) & 3;
	float w0 = (1 - fx) * (1 - fy), w1 = fx * (1 - fy), w2 = (1 - fx) * fy, w3 = 1 - (w0 + w1 + w2);
	if (dot( UnpackNormal2( n0 ), localNormal ) < 0.95f || (n0 & 3) != spec) w0 = 0;
	if (dot( UnpackNormal2( n1 ), localNormal ) < 0.95f || (n1 & 3) != spec) w1 = 0;
	if (dot( UnpackNormal2( n2 ), localNormal ) < 0.95f || (n2 & 3) != spec) w2 = 0;
	if (dot( UnpackNormal2( n3 ), localNormal ) < 0.95f || (n3 & 3) != spec) w3 = 0;
	const float sum = w0 + w1 + w2 + w3;
	if (sum == 0) return make_float4( -1, -1, -1, -1 );
	return (w0 * p0 + w1 * p1 + w2 * p2 + w3 * p3) * (1.0f / sum);
}
__device__ __forceinline__ void ReadTexelConsistent2( const HistBuf& buffer, const HistBuf& prevWorldPos, const ShardMap& m, const float4 localPos,
	const float3 localNormal, const float u, const float v, const int w, const int h, float3& direct, float3& indirect )
{
	direct.x = -1;
	const int iu1 = (int)floorf( u ), iv1 = (int)floorf( v ), iu0 = max( 0, iu1 - 1 ), iv0 = max( 0, iv1 - 1 );
	if (iu1 >= w || iv1 >= h || iu1 < 0 || iv1 < 0) return;
	const float fx = u - floorf( u ), fy = v - floorf( v );
	const float4* b0 = HistRow( buffer, m, iv0, w ), * b1 = HistRow( buffer, m, iv1, w ), * q0 = HistRow( prevWorldPos, m, iv0, w ), * q1 = HistRow( prevWorldPos, m, iv1, w );
	const float4 p0 = __ldg( b0 + iu0 ), p1 = __ldg( b0 + iu1 ), p2 = __ldg( b1 + iu0 ), p3 = __ldg( b1 + iu1 );
	const uint32_t n0 = __float_as_uint( __ldg( q0 + iu0 ).w ), n1 = __float_as_uint( __ldg( q0 + iu1 ).w );
	const uint32_t n2 = __float_as_uint( __ldg( q1 + iu0 ).w ), n3 = __float_as_uint( __ldg( q1 + iu1 ).w );
	const uint32_t spec = __float_as_uint( localPos.w ) & 3;
	float w0 = (1 - fx) * (1 - fy), w1 = fx * (1 - fy), w2 = (1 - fx) * fy, w3 = 1 - (w0 + w1 + w2);
	if (dot( UnpackNormal2( n0 ), localNormal ) < 0.975f || (n0 & 3) != spec) w0 = 0;
	if (dot( UnpackNormal2( n1 ), localNormal ) < 0.975f || (n1 & 3) != spec) w1 = 0;
	if (dot( UnpackNormal2( n2 ), localNormal ) < 0.975f || (n2 & 3) != spec) w2 = 0;
	if (dot( UnpackNormal2( n3 ), localNormal ) < 0.975f || (n3 & 3) != spec) w3 = 0;
	const float sum = w0 + w1 + w2 + w3;
	if (sum == 0) return;
	const float r = 1.0f / sum;
	direct = (w0 * DirectOf( p0 ) + w1 * DirectOf( p1 ) + w2 * DirectOf( p2 ) + w3 * DirectOf( p3 )) * r;
	indirect = (w0 * IndirectOf( p0 ) + w1 * IndirectOf( p1 ) + w2 * IndirectOf( p2 ) + w3 * IndirectOf( p3 )) * r;
}

/* ---- prepare (finalize_shared.h:169-314) ------------------------------------------------------------------------ */
/* curN = UnpackNormal2( cur.w ), unpacked once per pixel by the caller instead of once per texel (16 texels per search step) */
/* row: the texel's row of prevWorldPos, or null when that row lies outside the frame (ReadWorldPos then yields an invalid texel) */
__device__ __forceinline__ float WorldDistance( const int x, const float4* __restrict__ row, const float4 cur, const float3 curN, const int w )
{
	const float4 p = (row != nullptr && x >= 0 && x < w) ? __ldg( row + x ) : make_float4( 1e20f, 1e20f, 1e20f, __uint_as_float( 0 ) );
	if ((__float_as_uint( p.w ) & 3) != 1) return 1e21f;
	if (dot( curN, UnpackNormal2( __float_as_uint( p.w ) ) ) < 0.85f) return 1e21f;
	return sqrtf( sqrLen( make_float3( cur.x - p.x, cur.y - p.y, cur.z - p.z ) ) );
}
__device__ __forceinline__ float FineWorldDistance( const float px, const float py, const float4 cur, const float3 curN, const HistBuf& prevWorldPos, const ShardMap& m, const int w, const int h )
{
	const int x0 = (int)px, y0 = (int)py;
	const float4* r0 = (y0 >= 0 && y0 < h) ? HistRow( prevWorldPos, m, y0, w ) : nullptr, * r1 = (y0 + 1 >= 0 && y0 + 1 < h) ? HistRow( prevWorldPos, m, y0 + 1, w ) : nullptr;
	const float d0 = WorldDistance( x0, r0, cur, curN, w ), d1 = WorldDistance( x0 + 1, r0, cur, curN, w );
	const float d2 = WorldDistance( x0, r1, cur, curN, w ), d3 = WorldDistance( x0 + 1, r1, cur, curN, w );
	const float fx = px - floorf( px ), fy = py - floorf( py );
	const float w0 = (1 - fx) * (1 - fy), w1 = fx * (1 - fy), w2 = (1 - fx) * fy, w3 = fx * fy;
	float totalWeight = 0, totalDist = 0;
	if (d0 < 1e20f) totalDist += d0 * w0, totalWeight += w0;
	if (d1 < 1e20f) totalDist += d1 * w1, totalWeight += w1;
	if (d2 < 1e20f) totalDist += d2 * w2, totalWeight += w2;
	if (d3 < 1e20f) totalDist += d3 * w3, totalWeight += w3;
	return totalWeight == 0 ? 1e20f : totalDist / totalWeight;
}

struct PrepareArgs
{
	const float4* accumulator; uint4* features; const float4* worldPos; HistBuf prevWorldPos;
	float4* shading; float2* motion; float4* moments; HistBuf prevMoments; const float4* deltaDepth;
	ShardMap map; int yBlock0;	// first block row of the launch (a band of the frame when the chain is sharded)
	float4 prevPos, prevE, prevRight, prevUp;
	float j0, j1, prevj0, prevj1;
	int w, h; float pixelValueScale, directClamp, indirectClamp; int camIsStationary;
};

/* Per-pixel inputs of prepare: demodulated, clamped light (written to 'shading' by the first pass only) and its luminances. */
struct PrepareLocal { float lumDirect, lumDirect2, lumIndirect, lumIndirect2; };
__device__ __forceinline__ PrepareLocal PrepareLight( const PrepareArgs& a, const int pixelIdx, const uint4 feat, const bool store )
{
	const float3 direct = xyz( a.accumulator[pixelIdx] ) * a.pixelValueScale;
	const float3 albedo = RGB32toHDRmin1( feat.x );
	const float3 indirect = xyz( a.accumulator[pixelIdx + a.w * a.h] ) * a.pixelValueScale;
	const float3 reci = make_float3( 1.0f / albedo.x, 1.0f / albedo.y, 1.0f / albedo.z );
	const float3 directLight = min3f( direct * reci, a.directClamp ), indirectLight = min3f( indirect * reci, a.indirectClamp );
	if (store) a.shading[pixelIdx] = CombineToFloat4( directLight, indirectLight );
	PrepareLocal l;
	l.lumDirect = Luminance( directLight ), l.lumDirect2 = l.lumDirect * l.lumDirect;
	l.lumIndirect = Luminance( indirectLight ), l.lumIndirect2 = l.lumIndirect * l.lumIndirect;
	return l;
}

/* History lookup at the reprojected position, moments, history counter, motion vector (finalize_shared.h:286-313). */
__device__ __forceinline__ void PrepareFinish( const PrepareArgs& a, const int pixelIdx, const uint4 feat, const float4 lwp, float2 prev, PrepareLocal l )
{
	prev.x += 0.5f, prev.y += 0.5f;
	uint32_t fw = feat.w;
	if (prev.x >= 0 && prev.x < a.w && prev.y >= 0 && prev.y < a.h)
	{
		const float4 history = ReadTexelConsistent( a.prevMoments, a.prevWorldPos, a.map, lwp, UnpackNormal2( feat.y ), prev.x, prev.y, a.w, a.h );
		if (history.x > -1)
		{
			l.lumDirect = 0.2f * l.lumDirect + 0.8f * history.x, l.lumDirect2 = 0.2f * l.lumDirect2 + 0.8f * history.y;
			l.lumIndirect = 0.2f * l.lumIndirect + 0.8f * history.z, l.lumIndirect2 = 0.2f * l.lumIndirect2 + 0.8f * history.w;
			if ((fw & 15) < 15) fw++;
		}
		else fw &= 0xfffffff0u;
	}
	else fw &= 0xfffffff0u;
	if (fw != feat.w) a.features[pixelIdx].w = fw;
	a.motion[pixelIdx] = prev;
	a.moments[pixelIdx] = make_float4( l.lumDirect, l.lumDirect2, l.lumIndirect, l.lumIndirect2 );
}

/* specular pixels under a moving camera: diamond search for the world position of this pixel in the previous frame (:255-284) */
__device__ __forceinline__ float2 DiamondSearch( const PrepareArgs& a, const int pixelIdx, const int x, const int y, const float4 lwp )
{
	float2 prev = make_float2( (float)x, (float)y );
	const float4 pw = HistAt( a.prevWorldPos, a.map, x, y, a.w );
	const float3 lwpN = UnpackNormal2( __float_as_uint( lwp.w ) );
	float bestDist = sqrtf( sqrLen( make_float3( lwp.x - pw.x, lwp.y - pw.y, lwp.z - pw.z ) ) ), stepSize = 5.0f;
	const float ox = a.j0 - a.prevj0, oy = a.j1 - a.prevj1;
	int iter = 0;
	while (1)
	{
		int tap = 0;
		const float cx = prev.x, cy = prev.y;
		float d;
		d = FineWorldDistance( cx - stepSize + ox, cy + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
		if (d < bestDist) bestDist = d, prev = make_float2( cx - stepSize, cy ), tap = 1;
		d = FineWorldDistance( cx + stepSize + ox, cy + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
		if (d < bestDist) bestDist = d, prev = make_float2( cx + stepSize, cy ), tap = 2;
		d = FineWorldDistance( cx + ox, cy - stepSize + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
		if (d < bestDist) bestDist = d, prev = make_float2( cx, cy - stepSize ), tap = 3;
		d = FineWorldDistance( cx + ox, cy + stepSize + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
		if (d < bestDist) bestDist = d, prev = make_float2( cx, cy + stepSize ), tap = 4;
		if (tap == 0) { stepSize *= 0.45f; if (stepSize < 0.05f) break; }
		if (++iter == 25) break;
	}
	return prev;
}

/* First pass over all pixels. The diamond search is up to 25 x 16 dependent gathers that only specular pixels under a moving
   camera run; inside a 16 x 2-pixel warp that left 13 of 32 lanes working for 93 % of the kernel's instructions (ncu, C5).
   With a queue (DEFER) those pixels are appended - one warp-aggregated atomicAdd - and finished by prepareSearchKernel
   with full warps; everything a pixel writes is its own, so the split changes no value. */
template <bool DEFER> __global__ void __launch_bounds__( 256 ) prepareKernel( const PrepareArgs a, uint32_t* __restrict__ queue, uint32_t* __restrict__ queueCount )
{
	const int x = threadIdx.x + blockIdx.x * blockDim.x, y = threadIdx.y + (blockIdx.y + a.yBlock0) * blockDim.y;
	const bool inside = x < a.w && y < a.h;
	const int pixelIdx = inside ? x + y * a.w : 0;
	bool defer = false;
	if (inside)
	{
		const uint4 feat = a.features[pixelIdx];
		const float4 lwp = a.worldPos[pixelIdx];
		const PrepareLocal l = PrepareLight( a, pixelIdx, feat, true );
		float2 prev;
		if (((feat.w >> 4) & 3) == 0)
		{
			// diffuse: analytic reprojection into the previous view
			const float3 D = xyz( lwp ) - xyz( a.prevPos );
			const float il = rsqrtf( dot( D, D ) );
			const float3 Dn = D * il;
			const float t = a.prevPos.w / dot( xyz( a.prevE ), Dn );
			const float3 S = xyz( a.prevPos ) + Dn * t;
			prev = make_float2( dot( S, xyz( a.prevRight ) ) - a.prevRight.w - a.j0, dot( S, xyz( a.prevUp ) ) - a.prevUp.w - a.j1 );
		}
		else if (a.camIsStationary) prev = make_float2( (float)x, (float)y );
		else if (DEFER) defer = true;
		else prev = DiamondSearch( a, pixelIdx, x, y, lwp );
		if (!defer) PrepareFinish( a, pixelIdx, feat, lwp, prev, l );
	}
	if (DEFER)
	{
		// warp-aggregated append (all 32 lanes take part)
		const uint32_t mask = __ballot_sync( 0xffffffffu, defer );
		if (mask != 0)
		{
			const int lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31, leader = __ffs( mask ) - 1;
			uint32_t base = 0;
			if (lane == leader) base = atomicAdd( queueCount, (uint32_t)__popc( mask ) );
			base = __shfl_sync( 0xffffffffu, base, leader );
			if (defer) queue[base + __popc( mask & ((1u << lane) - 1) )] = (uint32_t)pixelIdx;
		}
	}
}

/* Second pass: the diamond search for the queued pixels, persistent-thread style. The number of search iterations varies from
   7 to 25 per pixel, so a lane that finishes early takes the next queued pixel (one warp-aggregated atomicAdd once 8 lanes are
   idle) instead of idling until the slowest lane of its warp is done. The result (the reprojected position) is parked in the
   motion buffer; prepareFinishKernel completes those pixels with full warps. counters: [0] = queue length, [1] = fetch cursor.
   (r2, measured and dropped: keeping the distances of the 4 x 4 texels around the search centre in shared memory once the step is below
   one pixel - the last five steps of every search revisit them - made the chain of C5 1.0 ms SLOWER: the centre changes its texel cell on
   most improving steps, and a refill costs as much as the step it replaces.) */
__global__ void __launch_bounds__( 256 ) prepareSearchKernel( const PrepareArgs a, const uint32_t* __restrict__ queue, uint32_t* __restrict__ counters )
{
	const uint32_t n = counters[0];
	const int lane = threadIdx.x & 31;
	bool active = false, exhausted = false;
	int pixelIdx = 0, iter = 0;
	float4 lwp = make_float4( 0, 0, 0, 0 );
	float3 lwpN = make_float3( 0, 0, 0 );
	float2 prev = make_float2( 0, 0 );
	float bestDist = 0, stepSize = 0;
	const float ox = a.j0 - a.prevj0, oy = a.j1 - a.prevj1;
	while (true)
	{
		const uint32_t idle = __ballot_sync( 0xffffffffu, !active );
		if (!exhausted && __popc( idle ) >= 8)
		{
			const int leader = __ffs( idle ) - 1;
			uint32_t base = 0;
			if (lane == leader) base = atomicAdd( counters + 1, (uint32_t)__popc( idle ) );
			base = __shfl_sync( 0xffffffffu, base, leader );
			if (base >= n) exhausted = true;
			const uint32_t mine = base + __popc( idle & ((1u << lane) - 1) );
			if (!active && mine < n)
			{
				pixelIdx = (int)queue[mine];
				const int y = pixelIdx / a.w, x = pixelIdx - y * a.w;
				lwp = a.worldPos[pixelIdx];
				lwpN = UnpackNormal2( __float_as_uint( lwp.w ) );
				const float4 pw = HistAt( a.prevWorldPos, a.map, x, y, a.w );
				prev = make_float2( (float)x, (float)y );
				bestDist = sqrtf( sqrLen( make_float3( lwp.x - pw.x, lwp.y - pw.y, lwp.z - pw.z ) ) ), stepSize = 5.0f, iter = 0;
				active = true;
			}
		}
		if (__ballot_sync( 0xffffffffu, active ) == 0)
		{
			if (exhausted) break;
			continue;
		}
		if (active)
		{
			// one step of the search loop of DiamondSearch above
			int tap = 0;
			const float cx = prev.x, cy = prev.y;
			float d;
			d = FineWorldDistance( cx - stepSize + ox, cy + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
			if (d < bestDist) bestDist = d, prev = make_float2( cx - stepSize, cy ), tap = 1;
			d = FineWorldDistance( cx + stepSize + ox, cy + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
			if (d < bestDist) bestDist = d, prev = make_float2( cx + stepSize, cy ), tap = 2;
			d = FineWorldDistance( cx + ox, cy - stepSize + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
			if (d < bestDist) bestDist = d, prev = make_float2( cx, cy - stepSize ), tap = 3;
			d = FineWorldDistance( cx + ox, cy + stepSize + oy, lwp, lwpN, a.prevWorldPos, a.map, a.w, a.h );
			if (d < bestDist) bestDist = d, prev = make_float2( cx, cy + stepSize ), tap = 4;
			bool done = false;
			if (tap == 0) { stepSize *= 0.45f; if (stepSize < 0.05f) done = true; }
			if (++iter == 25) done = true;
			if (done) a.motion[pixelIdx] = prev, active = false;
		}
	}
}

/* Third pass: history lookup, moments, history counter and the final motion vector of the queued pixels. */
__global__ void __launch_bounds__( 256 ) prepareFinishKernel( const PrepareArgs a, const uint32_t* __restrict__ queue, const uint32_t* __restrict__ counters )
{
	const uint32_t n = counters[0];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		const int pixelIdx = (int)queue[i];
		const uint4 feat = a.features[pixelIdx];
		const float4 lwp = a.worldPos[pixelIdx];
		const PrepareLocal l = PrepareLight( a, pixelIdx, feat, false );
		PrepareFinish( a, pixelIdx, feat, lwp, a.motion[pixelIdx], l );
	}
}

/* ---- a-trous (finalize_shared.h:320-484) ------------------------------------------------------------------------ */
#define AT_BX 32
struct AtrousArgs
{
	const uint4* features; HistBuf prevWorldPos; const float4* worldPos; const float4* deltaDepth; const float2* motion; const float4* moments;
	const float4* A; HistBuf B; float4* C;
	int w, h, phase, lastPass;
	ShardMap map; int yBlock0;
};

/* The block stages its tile plus the halo UNPACKED: every tile pixel is decoded once (5.11 fixed point -> float, luminances,
   10-bit normal, 10/11/11-bit albedo) instead of once per tap of every pixel that reads it (20 taps). Four float4 planes:
     dir = direct.rgb, luminance | ind = indirect.rgb, luminance | nrm = normal.xyz, depth | alb = albedo.rgb, feature word w
   ncu (profiles/r1_v4_filter_kernels_full.txt) had these kernels at 86-91 % of the L1 / shared-memory throughput (80 LDS.128 per
   pixel) with the XU pipe next (five MUFU per tap: lg2 + ex2 of powf( x, 128 ), two exp, one reciprocal). Hence:
     - a thread can filter PIX = 2 pixels STEP rows apart: four of the five tap rows of one are tap rows of the other, so a tap is
       loaded once and weighed for both centres - 48 instead of 80 LDS.128 per pixel, at twice the registers. */
struct AtrousCentre
{
	float3 normal, color, dirSum, indSum;
	float depth, ddx, ddy, lumDir, lumInd, rdir, rind, dirW, indW;
	int matID;
};

template <int STEP> __device__ __forceinline__ void AtrousTap( AtrousCentre& c, const float4 nDir, const float4 nInd, const float4 nNrm, const float4 nAlb,
	const int uu, const int vv, const int phase )
{
	const float w_dist = (uu * uu + vv * vv) * (-1.0f / 7.5f);
	// x^128 stays powf (lg2 + ex2 on the XU pipe under --use_fast_math): the kernel is bound by the FMA pipe (ncu r2: FMUL / FFMA / FADD are
	// 60 % of its instructions), seven squarings there measured slower than two MUFU here
	const float p = powf( fmaxf( 0.0f, dot( xyz( nNrm ), c.normal ) ), 128 );
	const float expected = c.depth + c.ddx * (float)(uu * STEP) + c.ddy * (float)(vv * STEP);
	const float depthError = fabsf( expected - nNrm.w );
	const float expectedDiff = fabsf( expected - c.depth );
	const float w_depth = depthError / fmaxf( 0.00001f, (0.5f + phase * 0.5f) * expectedDiff );
	const float w_normal = p * (((int)(__float_as_uint( nAlb.w ) >> 4) != c.matID) ? 0.0001f : dot( c.color, xyz( nAlb ) ));
	float wd = w_normal * __expf( fabsf( c.lumDir - nDir.w ) * c.rdir + w_dist - w_depth );
	float wi = w_normal * __expf( fabsf( c.lumInd - nInd.w ) * c.rind + w_dist - w_depth );
	if (!isfinite( wd )) wd = 0;
	if (!isfinite( wi )) wi = 0;
	c.dirSum += xyz( nDir ) * wd, c.dirW += wd;
	c.indSum += xyz( nInd ) * wi, c.indW += wi;
}

/* block: AT_BX x BY threads filter AT_BX x PIX*BY pixels; with PIX = 2 thread row ty owns the pixel rows base and base + STEP,
   base = (ty / STEP) * 2 * STEP + ty % STEP */
template <int STEP, int BY, int PIX> __global__ void __launch_bounds__( AT_BX * BY ) atrousKernel( const AtrousArgs a )
{
	constexpr int HALO = 2 * STEP, ROWS = PIX * BY, TW = AT_BX + 2 * HALO, TH = ROWS + 2 * HALO;
	static_assert( PIX == 1 || BY % STEP == 0, "rows of a block must pair up" );
	extern __shared__ float4 tile[];
	float4* tDir = tile, * tInd = tile + TW * TH, * tNrm = tile + 2 * TW * TH, * tAlb = tile + 3 * TW * TH;
	const int x0 = blockIdx.x * AT_BX, y0 = (blockIdx.y + a.yBlock0) * ROWS;
	for (int i = threadIdx.y * AT_BX + threadIdx.x; i < TW * TH; i += AT_BX * BY)
	{
		const int ty = i / TW, tx = i - ty * TW;
		const int gx = min( max( x0 - HALO + tx, 0 ), a.w - 1 ), gy = y0 - HALO + ty;
		if (gy >= 0 && gy < a.h)
		{
			const float4 c = a.A[gx + gy * a.w];
			const uint4 f = a.features[gx + gy * a.w];
			const float3 d = DirectOf( c ), in = IndirectOf( c ), n = UnpackNormal2( f.y ), al = RGB32toHDR( f.x );
			tDir[i] = f4( d, Luminance( d ) ), tInd[i] = f4( in, Luminance( in ) );
			tNrm[i] = f4( n, __uint_as_float( f.z ) ), tAlb[i] = f4( al, __uint_as_float( f.w ) );
		}
	}
	__syncthreads();
	const int x = x0 + threadIdx.x;
	const int rowBase = PIX == 1 ? (int)threadIdx.y : ((int)threadIdx.y / STEP) * 2 * STEP + (int)threadIdx.y % STEP;	// row of the first pixel inside the block
	if (x >= a.w || y0 + rowBase >= a.h) return;
	const int phase = a.phase, cx = threadIdx.x + HALO;
	const float sigma = 10.0f * OneOverPow2( phase - 1 );
	AtrousCentre c[PIX];
	int cy[PIX], py[PIX];
	bool live[PIX];
	uint32_t lfw[PIX];
#pragma unroll
	for (int k = 0; k < PIX; k++)
	{
		py[k] = y0 + rowBase + k * STEP, cy[k] = rowBase + k * STEP + HALO, live[k] = py[k] < a.h;
		const int pix = x + min( py[k], a.h - 1 ) * a.w;
		const float4 cDir = tDir[cx + cy[k] * TW], cInd = tInd[cx + cy[k] * TW], cNrm = tNrm[cx + cy[k] * TW], cAlb = tAlb[cx + cy[k] * TW];
		lfw[k] = __float_as_uint( cAlb.w );
		c[k].normal = xyz( cNrm ), c[k].color = xyz( cAlb ), c[k].matID = lfw[k] >> 4;
		c[k].dirW = 1, c[k].indW = 1, c[k].dirSum = xyz( cDir ), c[k].indSum = xyz( cInd );
		c[k].lumDir = cDir.w, c[k].lumInd = cInd.w, c[k].depth = cNrm.w;
		const float4 dd = a.deltaDepth[pix];
		c[k].ddx = dd.z, c[k].ddy = dd.w;
		const float factor = (lfw[k] & 15) == 0 ? 400.0f : 1.0f;
		const float4 m = a.moments[pix];
		const float var_dir = m.y - m.x * m.x, var_ind = m.w - m.z * m.z;
		c[k].rdir = -1.0f / (sigma * factor * sqrtf( var_dir + 0.00001f ) + 0.00001f);
		c[k].rind = -1.0f / (sigma * factor * sqrtf( var_ind + 0.00001f ) + 0.00001f);
	}
	// tap rows -2 .. 3 relative to the first pixel (in units of STEP): row r is row r of pixel 0 and row r - 1 of pixel 1
#pragma unroll
	for (int r = -2; r <= 1 + PIX; r++)
	{
		const int v = py[0] + r * STEP;
		if (v >= 0 && v < a.h)
		{
#pragma unroll
			for (int uu = -2; uu <= 2; uu++)
			{
				const int vv0 = r, vv1 = r - 1;
				const bool use0 = vv0 <= 2 && abs( uu ) <= (abs( vv0 ) == 2 ? 1 : 2) && (uu != 0 || vv0 != 0);
				const bool use1 = PIX == 2 && vv1 >= -2 && abs( uu ) <= (abs( vv1 ) == 2 ? 1 : 2) && (uu != 0 || vv1 != 0);
				if (!use0 && !use1) continue;
				// columns are clamped to the image by the tile loader, like the reference clamps u
				const int ti = (cx + uu * STEP) + (cy[0] + r * STEP) * TW;
				const float4 nDir = tDir[ti], nInd = tInd[ti], nNrm = tNrm[ti], nAlb = tAlb[ti];
				if (use0) AtrousTap<STEP>( c[0], nDir, nInd, nNrm, nAlb, uu, vv0, phase );
				if (use1) AtrousTap<STEP>( c[PIX - 1], nDir, nInd, nNrm, nAlb, uu, vv1, phase );
			}
		}
	}
#pragma unroll
	for (int k = 0; k < PIX; k++)
	{
		if (!live[k]) continue;
		const int y = py[k], pixelIdx = x + y * a.w;
		const float3 localNormal = c[k].normal, localColor = c[k].color;
		float3 dirF = c[k].dirSum * (1.0f / fmaxf( 0.0001f, c[k].dirW )), indF = c[k].indSum * (1.0f / fmaxf( 0.0001f, c[k].indW ));
		if (STEP == 1 && phase == 1)
		{
			// temporal blend with the previous frame's phase-1 output, clamped to the 3x3 YCoCg neighbourhood
			const float2 pp = a.motion[pixelIdx];
			const int px = (int)pp.x, pyy = (int)pp.y;
			if (px >= 0 && px < a.w && pyy >= 0 && pyy < a.h)
			{
				float3 prevDirect, prevIndirect;
				const float4 localPos = a.worldPos[pixelIdx];
				ReadTexelConsistent2( a.B, a.prevWorldPos, a.map, localPos, localNormal, pp.x, pp.y, a.w, a.h, prevDirect, prevIndirect );
				if (prevDirect.x != -1)
				{
					prevDirect = RGBToYCoCg( prevDirect ), prevIndirect = RGBToYCoCg( prevIndirect );
					float3 dirAvg = RGBToYCoCg( dirF ), dirVar = dirAvg * dirAvg, indAvg = RGBToYCoCg( indF ), indVar = indAvg * indAvg;
					auto tap = [&]( const int ox, const int oy ) {
						const int ti = (cx + ox) + (cy[k] + oy) * TW;
						const float3 f = RGBToYCoCg( xyz( tDir[ti] ) ), g = RGBToYCoCg( xyz( tInd[ti] ) );
						dirAvg += f, dirVar += f * f, indAvg += g, indVar += g * g;
					};
					if (x > 1)
					{
						if (y > 1) tap( -1, -1 );
						tap( -1, 0 );
						if (y < a.h - 1) tap( -1, 1 );
					}
					if (y > 1) tap( 0, -1 );
					if (y < a.h - 1) tap( 0, 1 );
					if (x < a.w - 1)
					{
						if (y > 1) tap( 1, -1 );
						tap( 1, 0 );
						if (y < a.h - 1) tap( 1, 1 );
					}
					dirAvg *= 1.0f / 9.0f, dirVar *= 1.0f / 9.0f, indAvg *= 1.0f / 9.0f, indVar *= 1.0f / 9.0f;
					float3 sDir = max3v( f3( 0 ), dirVar - dirAvg * dirAvg ), sInd = max3v( f3( 0 ), indVar - indAvg * indAvg );
					sDir = make_float3( sqrtf( sDir.x ), sqrtf( sDir.y ), sqrtf( sDir.z ) ), sInd = make_float3( sqrtf( sInd.x ), sqrtf( sInd.y ), sqrtf( sInd.z ) );
					prevDirect = clamp3( prevDirect, dirAvg - 0.75f * sDir, dirAvg + 0.75f * sDir );
					prevIndirect = clamp3( prevIndirect, indAvg - 0.75f * sInd, indAvg + 0.75f * sInd );
					dirF = dirF * 0.1f + YCoCgToRGB( prevDirect ) * 0.9f;
					indF = indF * 0.1f + YCoCgToRGB( prevIndirect ) * 0.9f;
				}
			}
		}
		if (a.lastPass)
		{
			const float3 o = (dirF + indF) * localColor;
			a.C[pixelIdx] = make_float4( sqrtf( o.x ), sqrtf( o.y ), sqrtf( o.z ), 1 );
		}
		else a.C[pixelIdx] = CombineToFloat4( dirF, indF );
	}
}

/* ---- tile staging with the TMA engine ---------------------------------------------------------------------------------------
   The 3x3-neighbourhood kernels (TAA, present) read a 34 x 10 float4 tile of one buffer. A tile row is 544 contiguous bytes in
   global memory, so the rows are fetched with bulk asynchronous copies (cp.async.bulk, SASS UBLKCP - the TMA engine without a
   tensor map) that signal an mbarrier: ten threads issue one copy each, nobody moves data through registers. Rows are clamped to
   the image by choosing the source row; blocks that touch the left / right image edge need clamped columns and take the plain
   path. (An out-of-image halo texel is never used by these kernels - their taps are guarded - so any value will do there.) */
__device__ __forceinline__ uint32_t SmemAddr( const void* p ) { return (uint32_t)__cvta_generic_to_shared( p ); }

__device__ __forceinline__ void LoadTile34x10( float4 (*tile)[34], uint64_t* bar, const float4* __restrict__ src, const int x0, const int y0, const int w, const int h )
{
	const int t = threadIdx.y * 32 + threadIdx.x;
	if (x0 >= 1 && x0 + 33 <= w)	// block-uniform
	{
		constexpr uint32_t ROW_BYTES = 34 * sizeof( float4 );
		if (t == 0)
		{
			asm volatile( "mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"( SmemAddr( bar ) ) );
			asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
			asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"( SmemAddr( bar ) ), "r"( 10 * ROW_BYTES ) : "memory" );
		}
		__syncthreads();
		if (t < 10)
		{
			const int gy = min( max( y0 - 1 + t, 0 ), h - 1 );
			const float4* row = src + (x0 - 1) + (size_t)gy * w;
			asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"( SmemAddr( &tile[t][0] ) ), "l"( row ), "r"( ROW_BYTES ), "r"( SmemAddr( bar ) ) : "memory" );
		}
		uint32_t done = 0;
		while (!done)
			asm volatile( "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"( done ) : "r"( SmemAddr( bar ) ) : "memory" );
	}
	else
	{
		for (int i = t; i < 340; i += 256)
		{
			const int ty = i / 34, tx = i - ty * 34;
			const int gx = min( max( x0 - 1 + tx, 0 ), w - 1 ), gy = min( max( y0 - 1 + ty, 0 ), h - 1 );
			tile[ty][tx] = src[gx + gy * w];
		}
		__syncthreads();
	}
}

/* ---- TAA (finalize_shared.h:498-548) ------------------------------------------------------------------------------ */
__global__ void __launch_bounds__( 256 ) taaKernel( const float4* __restrict__ pixelsIn, float4* __restrict__ pixelsOut, const __grid_constant__ HistBuf prevPixels, const __grid_constant__ ShardMap map,
	const float2* __restrict__ motion, const int w, const int h, const int yBlock0 )
{
	// (r2, measured on the same box: reading the 3x3 neighbourhood straight from global memory instead of the staged tile: 0.291 vs 0.297 ms
	// at 4K - the kernel's time is the 16 history gathers and the arithmetic, not the neighbourhood)
	__shared__ __align__( 16 ) float4 tile[10][34];
	__shared__ __align__( 8 ) uint64_t tileBar;
	const int x0 = blockIdx.x * 32, y0 = (blockIdx.y + yBlock0) * 8;
	LoadTile34x10( tile, &tileBar, pixelsIn, x0, y0, w, h );
	const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
	if (x >= w || y >= h) return;
	const int pixelIdx = x + y * w, cx = threadIdx.x + 1, cy = threadIdx.y + 1;
	float3 pixel = xyz( tile[cy][cx] );
	const float2 mv = motion[pixelIdx];
	const float pu = mv.x - 0.5f, pv = mv.y - 0.5f;
	if (pu >= 0 && pu < w && pv >= 0 && pv < h)
	{
		const float3 newPixel = RGBToYCoCg( pixel );
		// Mitchell-Netravali history fetch (sampling_shared.h:117-134)
		float3 hist;
		{
			const int x1 = (int)(pu - 2.0f), y1 = (int)(pv - 2.0f);
			float totalWeight = 0;
			float4 total = make_float4( 0, 0, 0, 0 );
			for (int yy = y1; yy < y1 + 4; yy++) if (yy > 0 && yy < h)
			{
				const float4* row = HistRow( prevPixels, map, yy, w );
				for (int xx = x1; xx < x1 + 4; xx++) if (xx >= 0 && xx < w)
				{
					const float weight = MitchellNetravali( (float)xx - pu ) * MitchellNetravali( (float)yy - pv );
					total = total + __ldg( row + xx ) * weight, totalWeight += weight;
				}
			}
			hist = xyz( total * (1.0f / totalWeight) );
		}
		float3 history = RGBToYCoCg( hist );
		float3 avg = newPixel, var = newPixel * newPixel;
		auto tap = [&]( const int ox, const int oy ) { const float3 f = RGBToYCoCg( xyz( tile[cy + oy][cx + ox] ) ); avg += f, var += f * f; };
		if (x > 1)
		{
			if (y > 1) tap( -1, -1 );
			tap( -1, 0 );
			if (y < h - 1) tap( -1, 1 );
		}
		if (y > 1) tap( 0, -1 );
		if (y < h - 1) tap( 0, 1 );
		if (x < w - 1)
		{
			if (y > 1) tap( 1, -1 );
			tap( 1, 0 );
			if (y < h - 1) tap( 1, 1 );
		}
		avg *= 1.0f / 9.0f, var *= 1.0f / 9.0f;
		float3 sigma = max3v( f3( 0 ), var - avg * avg );
		sigma = make_float3( sqrtf( sigma.x ), sqrtf( sigma.y ), sqrtf( sigma.z ) );
		history = clamp3( history, avg - 1.25f * sigma, avg + 1.25f * sigma );
		pixel = YCoCgToRGB( newPixel * 0.1f + history * 0.9f );
		if (isnan( pixel.x + pixel.y + pixel.z )) pixel = YCoCgToRGB( newPixel );
	}
	const float3 o = min3f( pixel, 10.0f );
	pixelsOut[pixelIdx] = make_float4( o.x, o.y, o.z, 0 );
}

/* ---- present: unsharpenTAAKernel (:554-583) or finalizeNoTAAKernel (:589-600); border pixels are left untouched ---- */
/* rows [rowFirst, rowEnd) only (the whole frame, or the band of this rank when the chain is sharded) */
__global__ void __launch_bounds__( 256 ) presentKernel( const float4* __restrict__ pixels, float4* __restrict__ target, const int w, const int h, const int taa,
	const int yBlock0, const int rowFirst, const int rowEnd )
{
	__shared__ __align__( 16 ) float4 tile[10][34];
	__shared__ __align__( 8 ) uint64_t tileBar;
	const int x0 = blockIdx.x * 32, y0 = (blockIdx.y + yBlock0) * 8;
	LoadTile34x10( tile, &tileBar, pixels, x0, y0, w, h );
	const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
	if (x == 0 || y == 0 || x >= w - 1 || y >= h - 1 || y < rowFirst || y >= rowEnd) return;
	const int cx = threadIdx.x + 1, cy = threadIdx.y + 1;
	const float4 c = tile[cy][cx];
	if (!taa)
	{
		target[x + y * w] = make_float4( sqrtf( c.x ), sqrtf( c.y ), sqrtf( c.z ), 0 );
		return;
	}
	const float4 p0 = tile[cy - 1][cx - 1], p1 = tile[cy - 1][cx], p2 = tile[cy - 1][cx + 1], p3 = tile[cy][cx + 1];
	const float4 p4 = tile[cy + 1][cx + 1], p5 = tile[cy + 1][cx], p6 = tile[cy + 1][cx - 1], p7 = tile[cy][cx - 1];
	const float4 blur = 0.35f * p0 + 0.5f * p1 + 0.35f * p2 + 0.5f * p3 + 0.35f * p4 + 0.5f * p5 + 0.35f * p6 + 0.5f * p7;
	const float4 sharp = c * 2.7f - 0.5f * blur;
	const float4 px = make_float4( fmaxf( c.x, sharp.x ), fmaxf( c.y, sharp.y ), fmaxf( c.z, sharp.z ), fmaxf( c.w, sharp.w ) );
	target[x + y * w] = make_float4( px.x * px.x, px.y * px.y, px.z * px.z, 0 );
}

/* ---- host side --------------------------------------------------------------------------------------------------- */
static void FilterChainImpl( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, float* hPrepare, float* hP1, float* hP2, float* hP3, cudaEvent_t* ev = nullptr )
{
	auto mark = [&]( int k ) { if (ev) cudaEventRecord( ev[k], st ); };
	mark( 0 );
	auto snap = [&]( float* dst, const float4* src ) { if (dst) cudaMemcpyAsync( dst, src, (size_t)s.w * s.h * 16, cudaMemcpyDeviceToHost, st ); };
	const int w = s.w, h = s.h;
	// rows this launch works on: the frame, or (sharded chain) this rank's band plus the halo - rowFirst is a multiple of 16
	const FilterShard* sh = b.shard;
	const int rowFirst = sh ? sh->rowFirst : 0, rowEnd = sh ? sh->rowEnd : h, rows = rowEnd - rowFirst;
	const int presentFirst = sh ? sh->presentFirst : 0, presentEnd = sh ? sh->presentEnd : h;
	const dim3 grid( (w + 31) / 32, (rows + 7) / 8 ), block( 32, 8 );
	ShardMap map = { 0, 0 };
	if (sh) map.magic = (uint32_t)((1ull << 32) / (uint32_t)sh->rowsPerBand + 1), map.last = sh->world - 1;
	// margin: how many rows beyond the band this rank's own copy of last frame's buffer is valid (16: every halo row, 14: phase-1 output, 1: TAA image)
	auto hist = [&]( const float4* local, const float4* const* perRank, const int margin ) {
		HistBuf hb = {};
		hb.local = local;
		if (!sh) { hb.lo = 0, hb.span = (uint32_t)h; return hb; }	// one GPU: every row is local
		hb.perRank = perRank;
		hb.lo = std::max( margin > 16 ? 0 : rowFirst, sh->presentFirst - margin );
		hb.span = (uint32_t)(std::min( margin > 16 ? h : rowEnd, sh->presentEnd + margin ) - hb.lo);
		return hb; };
	// reprojection constants of the previous view (finalize_shared.h:301-313)
	const float* pv = s.prevView;
	const float3 pos = make_float3( pv[0], pv[1], pv[2] ), p1 = make_float3( pv[3], pv[4], pv[5] ), p2 = make_float3( pv[6], pv[7], pv[8] ), p3 = make_float3( pv[9], pv[10], pv[11] );
	auto norm = []( const float3 v ) { const float l = 1.0f / sqrtf( v.x * v.x + v.y * v.y + v.z * v.z ); return v * l; };
	auto len = []( const float3 v ) { return sqrtf( v.x * v.x + v.y * v.y + v.z * v.z ); };
	const float3 centre = 0.5f * (p2 + p3), direction = norm( centre - pos ), right = norm( p2 - p1 ), up = norm( p3 - p1 );
	const float lenReci = h / len( p3 - p1 );
	PrepareArgs pa;
	pa.accumulator = b.accumulator, pa.features = b.features, pa.worldPos = b.worldPos, pa.prevWorldPos = hist( b.prevWorldPos, sh ? sh->prevWorldPos : nullptr, sh ? sh->worldPosMargin : 16 );
	pa.shading = b.shading, pa.motion = b.motion, pa.moments = b.moments, pa.prevMoments = hist( b.prevMoments, sh ? sh->prevMoments : nullptr, 16 ), pa.deltaDepth = b.deltaDepth;
	pa.map = map, pa.yBlock0 = rowFirst / 8;
	pa.prevPos = f4( pos, -(dot( pos, direction ) - dot( centre, direction )) ), pa.prevE = f4( direction, 0 );
	pa.prevRight = f4( right * lenReci, dot( p1, right ) * lenReci ), pa.prevUp = f4( up * lenReci, dot( p1, up ) * lenReci );
	pa.j0 = s.j0, pa.j1 = s.j1, pa.prevj0 = s.prevj0, pa.prevj1 = s.prevj1;
	pa.w = w, pa.h = h, pa.pixelValueScale = 1.0f / (float)s.samplesTaken, pa.directClamp = s.directClamp, pa.indirectClamp = s.indirectClamp;
	pa.camIsStationary = s.camIsStationary;
	if (b.taaOut && !s.camIsStationary)
	{
		// queue of deferred (specular) pixels in the not yet used TAA output buffer: uint32[w*h], then the counter
		uint32_t* queue = (uint32_t*)b.taaOut, * count = queue + (size_t)w * h;	// count[0]: queue length, count[1]: fetch cursor
		cudaMemsetAsync( count, 0, 2 * sizeof( uint32_t ), st );
		prepareKernel<true><<<grid, block, 0, st>>>( pa, queue, count );
		int dev = 0, sms = 148;
		cudaGetDevice( &dev ), cudaDeviceGetAttribute( &sms, cudaDevAttrMultiProcessorCount, dev );
		prepareSearchKernel<<<sms * 5, 256, 0, st>>>( pa, queue, count );
		prepareFinishKernel<<<sms * 8, 256, 0, st>>>( pa, queue, count );
	}
	else prepareKernel<false><<<grid, block, 0, st>>>( pa, nullptr, nullptr );
	mark( 1 );
	snap( hPrepare, b.shading );
	AtrousArgs aa;
	aa.features = b.features, aa.prevWorldPos = pa.prevWorldPos, aa.worldPos = b.worldPos, aa.deltaDepth = b.deltaDepth, aa.motion = b.motion, aa.moments = b.moments;
	aa.w = w, aa.h = h, aa.map = map;
	// step 1: 32 x 8 threads, one pixel each (62 registers, 8 blocks / SM); steps 2 and 4: 32 x 8 threads, two pixels STEP rows apart each
	// (the halo of 2 * STEP pixels is amortised over 16 rows and every tap is weighed for both) - measured per pass at 4K in
	// profiles/r2_reference_kernels.json; 64 bytes of tile per pixel
	auto smem = []( int step, int rows ) { return (size_t)(AT_BX + 4 * step) * (rows + 4 * step) * 64; };
	static bool attr = false;
	if (!attr)
	{
		cudaFuncSetAttribute( atrousKernel<2, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem( 2, 16 ) );
		cudaFuncSetAttribute( atrousKernel<4, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem( 4, 16 ) );
		attr = true;
	}
	const dim3 grid16( (w + 31) / 32, (rows + 15) / 16 );
	// sharded: phase 2 must not overwrite filteredIN - the other ranks read it (last frame's phase-1 output) until they are through phase 1
	float4* phase2Out = sh ? sh->phase2Out : b.filteredIN;
	aa.A = b.shading, aa.B = hist( b.filteredIN, sh ? sh->filteredIN : nullptr, 14 ), aa.C = b.filteredOUT, aa.phase = 1, aa.lastPass = 0, aa.yBlock0 = rowFirst / 8;
	atrousKernel<1, 8, 1><<<grid, block, smem( 1, 8 ), st>>>( aa );
	mark( 2 );
	snap( hP1, b.filteredOUT );
	aa.A = b.filteredOUT, aa.B = HistBuf{}, aa.C = phase2Out, aa.phase = 2, aa.yBlock0 = rowFirst / 16;
	atrousKernel<2, 8, 2><<<grid16, block, smem( 2, 16 ), st>>>( aa );
	mark( 3 );
	snap( hP2, phase2Out );
	aa.A = phase2Out, aa.C = b.shading, aa.phase = 3, aa.lastPass = 1;
	atrousKernel<4, 8, 2><<<grid16, block, smem( 4, 16 ), st>>>( aa );
	mark( 4 );
	snap( hP3, b.shading );
	const dim3 gridPresent( (w + 31) / 32, (presentEnd - presentFirst + 7) / 8 );
	if (s.taa)
	{
		// (sharded: the present pass needs one TAA row beyond the band - one 8-row block either side; their history taps are remote reads)
		const int taaFirst = sh ? std::max( rowFirst, presentFirst - 8 ) : 0, taaEnd = sh ? std::min( rowEnd, presentEnd + 8 ) : h;
		const dim3 gridTaa( (w + 31) / 32, (taaEnd - taaFirst + 7) / 8 );
		taaKernel<<<gridTaa, block, 0, st>>>( b.shading, b.taaOut, hist( b.prevPixels, sh ? sh->prevPixels : nullptr, 1 ), map, b.motion, w, h, taaFirst / 8 );
		mark( 5 );
		presentKernel<<<gridPresent, block, 0, st>>>( b.taaOut, b.target, w, h, 1, presentFirst / 8, presentFirst, presentEnd );
	}
	else
	{
		mark( 5 );
		presentKernel<<<gridPresent, block, 0, st>>>( b.shading, b.target, w, h, 0, presentFirst / 8, presentFirst, presentEnd );
	}
	mark( 6 );
}

void LaunchFilterChain( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, cudaEvent_t* stageEvents )
{
	FilterChainImpl( b, s, st, nullptr, nullptr, nullptr, nullptr, stageEvents );
}
void LaunchFilterChainStaged( const FilterBuffers& b, const FilterSettings& s, cudaStream_t st, float* hPrepare, float* hP1, float* hP2, float* hP3 )
{
	FilterChainImpl( b, s, st, hPrepare, hP1, hP2, hP3 );
}

} // namespace lh2b
