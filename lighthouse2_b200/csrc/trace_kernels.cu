/* trace_kernels.cu - the ray-tracing stages: generate + extend, extend, connect, and the stand-alone ray queries, all instances of
   the persistent warp-cooperative traversal (traverse_wide.cuh) over ray buffers in HBM: 32 B per ray in (O4, D4), 16 B out (hit
   record) or a 16-byte accumulator update per unoccluded shadow ray.
*/
#include "kernels.h"
#include "render_types.h"
#include "traverse.cuh"
#ifndef WIDE_MIN_BLOCKS
#define WIDE_MIN_BLOCKS 9		// resident blocks per SM the traversal kernels are compiled for (56 registers per thread, no spills)
#endif
#include "traverse_wide.cuh"

namespace lh2b
{

__device__ __forceinline__ float4 PackHit( const bool hit, const TraceResult& r )
{
	if (!hit) return make_float4( 0, 0, __int_as_float( -1 ), 1e34f );
	const uint32_t uv = (uint32_t)(65535.0f * r.u) + ((uint32_t)(65535.0f * r.v) << 16);
	return make_float4( __uint_as_float( uv ), __uint_as_float( r.inst ), __uint_as_float( r.prim ), r.t );
}

/* ---- wavefront stages ------------------------------------------------------------------- */

__device__ __forceinline__ uint32_t WangHashT( uint32_t s ) { s = (s ^ 61) ^ (s >> 16), s *= 9, s = s ^ (s >> 4), s *= 0x27d4eb2d, s = s ^ (s >> 15); return s; }
__device__ __forceinline__ float RandomFloatT( uint32_t& s ) { s ^= s << 13, s ^= s >> 17, s ^= s << 5; return s * 2.3283064365387e-10f; }

/* ---- persistent warp-cooperative kernels (traverse_wide.cuh) ---------------------------------------------------- */

struct BufferRaySource
{
	const float4* __restrict__ O4; const float4* __restrict__ D4; bool shadow;
	__device__ __forceinline__ bool Load( const uint32_t i, WideRay& r, uint32_t& tag ) const
	{
		const float4 o = __ldcs( O4 + i ), d = __ldcs( D4 + i );	// streamed once: evict-first, the BVH keeps the L2
		r.O = make_float3( o.x, o.y, o.z ), r.D = make_float3( d.x, d.y, d.z );
		r.tmin = 0.0f, r.tmax = shadow ? d.w : 1e34f;
		tag = i;
		return true;
	}
};

struct HitBufferSink
{
	float4* __restrict__ hits;
	__device__ __forceinline__ void Closest( const uint32_t i, const bool hit, const TraceResult& r ) const { __stcs( hits + i, PackHit( hit, r ) ); }
	__device__ __forceinline__ void AnyHit( const uint32_t, const bool ) const {}
};

struct OccludedFlagSink
{
	uint8_t* __restrict__ flags;
	__device__ __forceinline__ void Closest( const uint32_t, const bool, const TraceResult& ) const {}
	__device__ __forceinline__ void AnyHit( const uint32_t i, const bool occluded ) const { flags[i] = occluded ? 1 : 0; }
};

struct ConnectSink
{
	const float4* __restrict__ E4; float4* __restrict__ accumulator;
	__device__ __forceinline__ void Closest( const uint32_t, const bool, const TraceResult& ) const {}
	__device__ __forceinline__ void AnyHit( const uint32_t i, const bool occluded ) const
	{
		if (occluded) return;
		const float4 e = __ldcs( E4 + i );
		atomicAdd( accumulator + __float_as_int( e.w ), make_float4( e.x, e.y, e.z, 1 ) );
	}
};

/* Primary rays: work items are 8x4 pixel tiles per warp (coherent rays per warp), per sample. */
/* Exact n / d for any 32-bit n with m = ceil( 2^64 / d ), d >= 2 (n * (d * m - 2^64) < 2^64 always holds). */
__device__ __forceinline__ uint32_t DivMagic( const uint32_t n, const uint64_t m ) { return (uint32_t)__umul64hi( (uint64_t)n, m ); }
__host__ __device__ __forceinline__ uint64_t MagicOf( const uint32_t d ) { return d < 2 ? 0 : 0xffffffffffffffffull / d + 1; }

/* (sx, sy) pixel, s = sample index within this frame's shard; pathIdx = sx + sy * w + s * w * h */
__device__ __forceinline__ void GeneratePrimaryAt( const RenderParams& p, const int sx, const int sy, const uint32_t s, const uint32_t pathIdx, float3& O, float3& D )
{
	const uint32_t pixels = p.w * p.h;
	const uint32_t seedIdx = pathIdx + p.sampleBase * pixels;
	const uint32_t sampleIdx = s + p.sampleBase + p.pass;
	uint32_t seed = WangHashT( seedIdx * 16789 + p.pass * 1791 );
	float4 r4;
	if (sampleIdx < 64)
	{
		const int x = (sx + (p.shift & 127)) & 127, y = (sy + (p.shift >> 24)) & 127;
		const uint32_t* bn = p.blueNoise;
		const uint4 rank = *(const uint4*)(bn + (x + y * 128) * 8 + 65536 * 3);
		const uint32_t v0 = bn[0 + ((sampleIdx ^ rank.x) & 255) * 256], v1 = bn[1 + ((sampleIdx ^ rank.y) & 255) * 256];
		const uint32_t v2 = bn[2 + ((sampleIdx ^ rank.z) & 255) * 256], v3 = bn[3 + ((sampleIdx ^ rank.w) & 255) * 256];
		const uint4 scr = *(const uint4*)(bn + (x + y * 128) * 8 + 65536);
		r4 = make_float4( (0.5f + (int)(v0 ^ scr.x)) * (1.0f / 256.0f), (0.5f + (int)(v1 ^ scr.y)) * (1.0f / 256.0f),
			(0.5f + (int)(v2 ^ scr.z)) * (1.0f / 256.0f), (0.5f + (int)(v3 ^ scr.w)) * (1.0f / 256.0f) );
	}
	else r4.x = RandomFloatT( seed ), r4.y = RandomFloatT( seed ), r4.z = RandomFloatT( seed ), r4.w = RandomFloatT( seed );
	const float ap = p.posLensSize.w;
	O = make_float3( p.posLensSize.x, p.posLensSize.y, p.posLensSize.z );
	if (ap != 0)	// pinhole: the lens offset is 0 * finite, skip it (launch-uniform branch, same result)
	{
		const float blade = (float)(int)(r4.x * 9);
		float r1 = r4.z, r2 = (r4.x - blade * (1.0f / 9.0f)) * 9.0f;
		float x1, y1, x2, y2;
		const float PI_T = 3.14159265358979323846264f;
		__sincosf( blade * PI_T / 4.5f, &x1, &y1 );
		__sincosf( (blade + 1.0f) * PI_T / 4.5f, &x2, &y2 );
		if ((r1 + r2) > 1) r1 = 1.0f - r1, r2 = 1.0f - r2;
		const float xr = x1 * r1 + x2 * r2, yr = y1 * r1 + y2 * r2;
		O = make_float3( p.posLensSize.x + ap * (p.right.x * xr + p.up.x * yr), p.posLensSize.y + ap * (p.right.y * xr + p.up.y * yr),
			p.posLensSize.z + ap * (p.right.z * xr + p.up.z * yr) );
	}
	float fu, fv;
	if (p.distortion == 0) fu = ((float)sx + r4.y) * (1.0f / p.w), fv = ((float)sy + r4.w) * (1.0f / p.h);
	else
	{
		const float tx = sx / (float)p.w - 0.5f, ty = sy / (float)p.h - 0.5f;
		const float rr = tx * tx + ty * ty;
		const float rq = sqrtf( rr ) * (1.0f + p.distortion * rr + p.distortion * rr * rr);
		const float theta = atan2f( tx, ty );
		const float bx = (sinf( theta ) * rq + 0.5f) * p.w, by = (cosf( theta ) * rq + 0.5f) * p.h;
		fu = (bx + r4.y) / (float)p.w, fv = (by + r4.w) / (float)p.h;
	}
	D = make_float3( p.p1.x + fu * p.right.x + fv * p.up.x - O.x, p.p1.y + fu * p.right.y + fv * p.up.y - O.y, p.p1.z + fu * p.right.z + fv * p.up.z - O.z );
	const float il = rsqrtf( D.x * D.x + D.y * D.y + D.z * D.z );
	D.x *= il, D.y *= il, D.z *= il;
}

__device__ __forceinline__ void GeneratePrimary( const RenderParams& p, const uint32_t pathIdx, float3& O, float3& D )
{
	const uint32_t pixels = p.w * p.h, s = pathIdx / pixels, pixelIdx = pathIdx - s * pixels;
	const int sy = pixelIdx / p.w, sx = pixelIdx - sy * p.w;
	GeneratePrimaryAt( p, sx, sy, s, pathIdx, O, D );
}

template <bool BAND> struct TiledPrimarySource	// BAND: tile-sharded frame (this core renders some tile rows only); false keeps the whole-frame index math
{
	const RenderParams* p; float4* __restrict__ outO; float4* __restrict__ outD; uint32_t* __restrict__ pathOf;	// pathOf: smem-free mapping kept in registers by the sink
	uint32_t tilesX, itemsPerSample, tileY0, tileStep;	// this core's tile rows: tileY0 + j * tileStep
	uint64_t tilesXMagic, itemsMagic;	// MagicOf( tilesX ), MagicOf( itemsPerSample ): the index arithmetic below runs per ray at ~1/3 SIMT width
	__device__ __forceinline__ bool Load( const uint32_t work, WideRay& r, uint32_t& tag ) const
	{
		// work item -> (sample, 8x4 pixel tile, lane in tile)
		const uint32_t s = p->spp == 1 ? 0 : DivMagic( work, itemsMagic ), w = work - s * itemsPerSample;
		const uint32_t tile = w >> 5, l = w & 31;
		const uint32_t ty = tilesX == 1 ? tile : DivMagic( tile, tilesXMagic ), tx = tile - ty * tilesX;
		const uint32_t x = tx * 8 + (l & 7), y = (BAND ? tileY0 + ty * tileStep : ty) * 4 + (l >> 3);
		if (x >= (uint32_t)p->w || (BAND ? ((int)y < p->bandY0 || (int)y >= p->bandY1) : y >= (uint32_t)p->h)) return false;
		const uint32_t pathIdx = x + y * p->w + s * (p->w * p->h);
		tag = pathIdx;
		GeneratePrimaryAt( *p, (int)x, (int)y, s, pathIdx, r.O, r.D );
		r.tmin = 0.0f, r.tmax = 1e34f;
		__stcs( outO + pathIdx, make_float4( r.O.x, r.O.y, r.O.z, __uint_as_float( (pathIdx << 6) + 1 /* S_SPECULAR */ ) ) );
		__stcs( outD + pathIdx, make_float4( r.D.x, r.D.y, r.D.z, 0 ) );
		return true;
	}
};

struct TiledHitSink
{
	float4* __restrict__ hits;
	__device__ __forceinline__ void Closest( const uint32_t pathIdx, const bool hit, const TraceResult& r ) const { __stcs( hits + pathIdx, PackHit( hit, r ) ); }
	__device__ __forceinline__ void AnyHit( const uint32_t, const bool ) const {}
};

/* generate + extend for path length 1: setupPrimaryRay (.optix.cu:112-129) with generateEyeRay (:85-104), RandomPointOnLens (:71-83),
   blueNoiseSampler4 (:56-69) and RayTarget (lib/RenderSystem/common_functions.h:29-50), fused into the work fetch of the traversal. */
template <bool TWO_LEVEL, bool BAND, bool STATS> __global__ void __launch_bounds__( WIDE_BLOCK, WIDE_MIN_BLOCKS ) wideGenerateExtendKernel( const DevScene scene, const RenderParams p, const PathSet out, float4* __restrict__ hits,
	uint32_t* workCounter, const WideTuning tune )
{
	TiledPrimarySource<BAND> src;
	src.p = &p, src.outO = out.O, src.outD = out.D, src.pathOf = nullptr;
	src.tilesX = (p.w + 7) / 8, src.tileY0 = (uint32_t)p.bandY0 / 4, src.tileStep = (uint32_t)p.bandStep;
	src.itemsPerSample = src.tilesX * (BAND ? BandTileRows( p ) : ((uint32_t)p.h + 3) / 4) * 32;
	src.tilesXMagic = MagicOf( src.tilesX ), src.itemsMagic = MagicOf( src.itemsPerSample );
	TiledHitSink sink = { hits };
	TraverseWide<false, TWO_LEVEL, STATS>( scene, src, sink, src.itemsPerSample * p.spp, workCounter, tune, scene.stats );
}

/* extend for path length > 1 (setupSecondaryRay, .optix.cu:131-140): the ray count comes from a device counter (frames) or is fixed (queries) */
template <bool TWO_LEVEL, bool STATS> __global__ void __launch_bounds__( WIDE_BLOCK, WIDE_MIN_BLOCKS ) wideExtendKernel( const DevScene scene, const float4* __restrict__ O4, const float4* __restrict__ D4,
	float4* __restrict__ hits, const uint32_t* __restrict__ countPtr, const uint32_t fixedCount, uint32_t* workCounter, const WideTuning tune )
{
	BufferRaySource src = { O4, D4, false };
	HitBufferSink sink = { hits };
	TraverseWide<false, TWO_LEVEL, STATS>( scene, src, sink, countPtr ? *countPtr : fixedCount, workCounter, tune, scene.stats );
}

/* the traversal half of generateShadowRay (.optix.cu:142-149), reporting a flag per ray (ray queries) */
template <bool TWO_LEVEL, bool STATS> __global__ void __launch_bounds__( WIDE_BLOCK, WIDE_MIN_BLOCKS ) wideOccludeKernel( const DevScene scene, const float4* __restrict__ O4, const float4* __restrict__ D4,
	uint8_t* __restrict__ flags, const uint32_t fixedCount, uint32_t* workCounter, const WideTuning tune )
{
	BufferRaySource src = { O4, D4, true };
	OccludedFlagSink sink = { flags };
	TraverseWide<true, TWO_LEVEL, STATS>( scene, src, sink, fixedCount, workCounter, tune, scene.stats );
}

/* connect (generateShadowRay, .optix.cu:142-154): unoccluded shadow rays deposit their potential contribution */
template <bool TWO_LEVEL, bool STATS> __global__ void __launch_bounds__( WIDE_BLOCK, WIDE_MIN_BLOCKS ) wideConnectKernel( const DevScene scene, const PathSet conn, float4* __restrict__ accumulator,
	const uint32_t* __restrict__ countPtr, uint32_t* workCounter, const WideTuning tune )
{
	BufferRaySource src = { conn.O, conn.D, true };
	ConnectSink sink = { conn.T, accumulator };
	TraverseWide<true, TWO_LEVEL, STATS>( scene, src, sink, *countPtr, workCounter, tune, scene.stats );
}

/* flat scenes: write the owning instance index into the spare word of every traversal triangle */
__global__ void tagTrianglesKernel( float4* tris, const int triCount, const uint32_t inst )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < triCount) tris[(size_t)i * 3 + 1].w = __uint_as_float( inst );
}

void LaunchTagTriangles( float4* tris, int triCount, uint32_t inst, cudaStream_t s )
{
	if (triCount > 0) tagTrianglesKernel<<<(triCount + 255) / 256, 256, 0, s>>>( tris, triCount, inst );
}

static uint32_t PersistentGrid( uint32_t maxItems, int smCount, int blocksPerSM )
{
	const uint32_t need = (maxItems + WIDE_BLOCK - 1) / WIDE_BLOCK, cap = (uint32_t)(smCount * blocksPerSM);
	return need < cap ? (need ? need : 1) : cap;
}

int g_wideBlocksPerSM = WIDE_MIN_BLOCKS;
int g_triThreshold = WIDE_TRI_THRESHOLD, g_refillThreshold = WIDE_REFILL_THRESHOLD, g_triThresholdShadow = WIDE_TRI_THRESHOLD, g_raysPerLane = 1;
#define TUNE WideTuning{ g_triThreshold, g_refillThreshold, g_raysPerLane, 0x47800000u }
#define TUNE_SHADOW WideTuning{ g_triThresholdShadow, g_refillThreshold, g_raysPerLane, 0x47800000u }
/* picks the instantiation: two-level or flat scene, work counters on (scene.stats set by lh2b_trace_stats) or off */
#define WIDE_LAUNCH( kernel, grid, s, ... ) do { \
	if (scene.singleIdentity) { if (scene.stats) kernel<false, true><<<grid, WIDE_BLOCK, 0, s>>>( __VA_ARGS__ ); else kernel<false, false><<<grid, WIDE_BLOCK, 0, s>>>( __VA_ARGS__ ); } \
	else { if (scene.stats) kernel<true, true><<<grid, WIDE_BLOCK, 0, s>>>( __VA_ARGS__ ); else kernel<true, false><<<grid, WIDE_BLOCK, 0, s>>>( __VA_ARGS__ ); } } while (0)

void LaunchGenerateExtend( const DevScene& scene, const RenderParams& p, const PathSet& out, float4* hits, uint32_t* workCounter, int smCount, cudaStream_t s )
{
	const uint32_t items = ((p.w + 7) / 8) * BandTileRows( p ) * 32 * p.spp, grid = PersistentGrid( items, smCount, g_wideBlocksPerSM );
	const bool band = !(p.bandY0 == 0 && p.bandY1 == p.h && p.bandStep == 1);
	const bool stats = scene.stats != nullptr;
#define GEN_LAUNCH( TL, BAND, ST ) wideGenerateExtendKernel<TL, BAND, ST><<<grid, WIDE_BLOCK, 0, s>>>( scene, p, out, hits, workCounter, TUNE )
	if (scene.singleIdentity)
	{
		if (band) { if (stats) GEN_LAUNCH( false, true, true ); else GEN_LAUNCH( false, true, false ); }
		else { if (stats) GEN_LAUNCH( false, false, true ); else GEN_LAUNCH( false, false, false ); }
	}
	else if (band) { if (stats) GEN_LAUNCH( true, true, true ); else GEN_LAUNCH( true, true, false ); }
	else { if (stats) GEN_LAUNCH( true, false, true ); else GEN_LAUNCH( true, false, false ); }
#undef GEN_LAUNCH
}

void LaunchExtendCounted( const DevScene& scene, const PathSet& in, float4* hits, const uint32_t* countPtr, uint32_t* workCounter, uint32_t maxRays, int smCount, cudaStream_t s )
{
	const uint32_t grid = PersistentGrid( maxRays, smCount, g_wideBlocksPerSM );
	WIDE_LAUNCH( wideExtendKernel, grid, s, scene, in.O, in.D, hits, countPtr, 0, workCounter, TUNE );
}

void LaunchConnect( const DevScene& scene, const PathSet& conn, float4* accumulator, const uint32_t* countPtr, uint32_t* workCounter, uint32_t maxRays, int smCount, cudaStream_t s )
{
	const uint32_t grid = PersistentGrid( maxRays, smCount, g_wideBlocksPerSM );
	WIDE_LAUNCH( wideConnectKernel, grid, s, scene, conn, accumulator, countPtr, workCounter, TUNE_SHADOW );
}

void LaunchExtend( const DevScene& scene, const float4* O4, const float4* D4, float4* hits, int n, uint32_t* workCounter, int smCount, cudaStream_t s )
{
	if (n <= 0) return;
	cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), s );
	const uint32_t grid = PersistentGrid( (uint32_t)n, smCount, g_wideBlocksPerSM );
	WIDE_LAUNCH( wideExtendKernel, grid, s, scene, O4, D4, hits, nullptr, (uint32_t)n, workCounter, TUNE );
}

void LaunchOcclude( const DevScene& scene, const float4* O4, const float4* D4, uint8_t* occluded, int n, uint32_t* workCounter, int smCount, cudaStream_t s )
{
	if (n <= 0) return;
	cudaMemsetAsync( workCounter, 0, sizeof( uint32_t ), s );
	const uint32_t grid = PersistentGrid( (uint32_t)n, smCount, g_wideBlocksPerSM );
	WIDE_LAUNCH( wideOccludeKernel, grid, s, scene, O4, D4, occluded, (uint32_t)n, workCounter, TUNE_SHADOW );
}

} // namespace lh2b
