/* traverse.cuh - device-side ray traversal of the two-level CWBVH (sm_100a).

   Replaces optixTrace (reference call sites lib/rendercore_optix7/optix/.optix.cu:125 primary,
   :136 secondary, :148 shadow). Contract taken from those call sites:
     - closest hit in (tmin, tmax), no face culling, no any-hit program;
     - single-level instancing, ray transformed by the instance's inverse 3x4, t preserved;
     - hit record = (u16 | v16 << 16, instance index, primitive index, t), u/v truncated to 16 bit
       (.optix.cu:174-184), u weighs vertex1, v weighs vertex2;
     - shadow rays: terminate on first hit (.optix.cu:147-149).

   Arithmetic is spelled with explicit round-to-nearest intrinsics so that the CPU oracle
   (oracle/lh2_oracle.cpp, same operation order, -ffp-contract=off) produces bit-identical t, u, v.
   Ties in t are resolved toward the smaller (instance, primitive) pair, so hit IDs are a pure
   function of the ray and the scene, independent of traversal order.
*/
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lh2b
{

/* All BLASes and the TLAS live in ONE node arena and ONE triangle arena; child / triangle base indices inside the
   nodes are absolute arena indices (baked at build time), so traversal needs no per-ray pointers. */
struct InstTrav
{
	float4 r0, r1, r2;		// rows of the world->object 3x4
	uint32_t rootNode;		// arena index of the BLAS root node
	uint32_t flags;			// bit 0: identity transform
	uint32_t pad0, pad1;
};

struct DevScene
{
	const uint4* nodes;				// node arena (5 x uint4 per node)
	const float4* tris;				// triangle arena (3 x float4 per triangle)
	const uint32_t* tlasLeafIds;	// instance index per top-level leaf slot
	const InstTrav* instances;
	uint32_t tlasRoot;				// arena index of the top-level root node
	int instanceCount;
	int singleIdentity;				// 1: exactly one instance with identity transform -> skip the top level
	uint32_t singleRoot;			// its BLAS root
};

#define LH2B_STACK 48

__device__ __forceinline__ uint32_t SignExtendS8x4( const uint32_t x )
{
	uint32_t r;
	asm( "prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"( r ) : "r"( x ) );
	return r;
}

__device__ __forceinline__ float ByteToFloat( const uint32_t v, const int j )
{
	return (float)((v >> (8 * j)) & 255u);
}

/* exact-order helpers (mirrored in the oracle) */
__device__ __forceinline__ float Dot3( const float ax, const float ay, const float az, const float bx, const float by, const float bz )
{
	return __fmaf_rn( ax, bx, __fmaf_rn( ay, by, __fmul_rn( az, bz ) ) );
}
#define CROSS_X( ax, ay, az, bx, by, bz ) __fmaf_rn( ay, bz, -__fmul_rn( az, by ) )
#define CROSS_Y( ax, ay, az, bx, by, bz ) __fmaf_rn( az, bx, -__fmul_rn( ax, bz ) )
#define CROSS_Z( ax, ay, az, bx, by, bz ) __fmaf_rn( ax, by, -__fmul_rn( ay, bx ) )

__device__ __forceinline__ float SafeRcpDir( const float d )
{
	// avoid inf/NaN slabs for axis-parallel rays; 1e-20 keeps (p - o) * idir finite
	const float a = fabsf( d ) > 1e-20f ? d : copysignf( 1e-20f, d );
	return __frcp_rn( a );
}

struct TraceResult { float t; uint32_t inst, prim; float u, v; };

template <bool ANYHIT>
__device__ __forceinline__ bool Traverse( const DevScene& scene, const float3 wO, const float3 wD, const float tmin, float tmax, TraceResult& res )
{
	uint2 stack[LH2B_STACK];
	int sp = 0;
	float3 O = wO, D = wD;
	const uint4* __restrict__ nodes = scene.nodes;
	const float4* __restrict__ tris = scene.tris;
	bool inBlas = scene.singleIdentity != 0;
	uint32_t curInst = 0;
	float idx = SafeRcpDir( D.x ), idy = SafeRcpDir( D.y ), idz = SafeRcpDir( D.z );
	uint32_t octinv = (D.x < 0 ? 0 : 4) | (D.y < 0 ? 0 : 2) | (D.z < 0 ? 0 : 1);
	uint32_t octinv4 = octinv * 0x01010101u;
	uint2 ng = make_uint2( scene.singleIdentity ? scene.singleRoot : scene.tlasRoot, 0x80000000u ), tg = make_uint2( 0, 0 );
	uint32_t bestInst = 0xffffffffu, bestPrim = 0xffffffffu;
	float bestU = 0, bestV = 0;
	while (true)
	{
		if (ng.y > 0x00ffffffu)
		{
			const uint32_t hits = ng.y;
			const int bit = 31 - __clz( hits );
			ng.y &= ~(1u << bit);
			if (ng.y > 0x00ffffffu) stack[sp++] = ng;
			const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
			const uint32_t rel = __popc( hits & ~(0xffffffffu << slot) & 0xffu );
			const uint4* np = nodes + (size_t)(ng.x + rel) * 5;
			const uint4 n0 = __ldg( np ), n1 = __ldg( np + 1 ), n2 = __ldg( np + 2 ), n3 = __ldg( np + 3 ), n4 = __ldg( np + 4 );
			const float sx = __uint_as_float( (n0.w & 255u) << 23 ) * idx;
			const float sy = __uint_as_float( ((n0.w >> 8) & 255u) << 23 ) * idy;
			const float sz = __uint_as_float( ((n0.w >> 16) & 255u) << 23 ) * idz;
			const float cx = (__uint_as_float( n0.x ) - O.x) * idx;
			const float cy = (__uint_as_float( n0.y ) - O.y) * idy;
			const float cz = (__uint_as_float( n0.z ) - O.z) * idz;
			ng.x = n1.x, tg.x = n1.y;
			uint32_t hitmask = 0;
#pragma unroll
			for (int half = 0; half < 2; half++)
			{
				const uint32_t meta4 = half ? n1.w : n1.z;
				const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
				const uint32_t innerMask4 = SignExtendS8x4( isInner4 << 3 );
				const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
				const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
				const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
				const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
				const uint32_t nx = D.x < 0 ? qhix : qlox, fx = D.x < 0 ? qlox : qhix;
				const uint32_t ny = D.y < 0 ? qhiy : qloy, fy = D.y < 0 ? qloy : qhiy;
				const uint32_t nz = D.z < 0 ? qhiz : qloz, fz = D.z < 0 ? qloz : qhiz;
#pragma unroll
				for (int j = 0; j < 4; j++)
				{
					const float t0x = fmaf( ByteToFloat( nx, j ), sx, cx ), t1x = fmaf( ByteToFloat( fx, j ), sx, cx );
					const float t0y = fmaf( ByteToFloat( ny, j ), sy, cy ), t1y = fmaf( ByteToFloat( fy, j ), sy, cy );
					const float t0z = fmaf( ByteToFloat( nz, j ), sz, cz ), t1z = fmaf( ByteToFloat( fz, j ), sz, cz );
					const float cmin = fmaxf( fmaxf( t0x, t0y ), fmaxf( t0z, tmin ) );
					// far side padded by a few ulp: the slab arithmetic differs from the exact triangle test
					const float cmax = fminf( fminf( t1x, t1y ), fminf( t1z, tmax ) ) * 1.0000005f;
					if (cmin <= cmax)
					{
						const uint32_t cb = (childBits4 >> (8 * j)) & 255u, bi = (bitIndex4 >> (8 * j)) & 255u;
						hitmask |= cb << bi;
					}
				}
			}
			ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
			tg.y = hitmask & 0x00ffffffu;
		}
		else
		{
			tg = ng;
			ng = make_uint2( 0, 0 );
		}
		while (tg.y != 0)
		{
			const int bit = 31 - __clz( tg.y );
			tg.y &= ~(1u << bit);
			if (inBlas)
			{
				const float4* tp = tris + (size_t)(tg.x + bit) * 3;
				const float4 v0 = __ldg( tp ), e1 = __ldg( tp + 1 ), e2 = __ldg( tp + 2 );
				const float pvx = CROSS_X( D.x, D.y, D.z, e2.x, e2.y, e2.z );
				const float pvy = CROSS_Y( D.x, D.y, D.z, e2.x, e2.y, e2.z );
				const float pvz = CROSS_Z( D.x, D.y, D.z, e2.x, e2.y, e2.z );
				const float det = Dot3( e1.x, e1.y, e1.z, pvx, pvy, pvz );
				if (det != 0.0f)
				{
					const float inv = __frcp_rn( det );
					const float tvx = __fsub_rn( O.x, v0.x ), tvy = __fsub_rn( O.y, v0.y ), tvz = __fsub_rn( O.z, v0.z );
					const float u = __fmul_rn( Dot3( tvx, tvy, tvz, pvx, pvy, pvz ), inv );
					if (u >= 0.0f && u <= 1.0f)
					{
						const float qvx = CROSS_X( tvx, tvy, tvz, e1.x, e1.y, e1.z );
						const float qvy = CROSS_Y( tvx, tvy, tvz, e1.x, e1.y, e1.z );
						const float qvz = CROSS_Z( tvx, tvy, tvz, e1.x, e1.y, e1.z );
						const float v = __fmul_rn( Dot3( D.x, D.y, D.z, qvx, qvy, qvz ), inv );
						if (v >= 0.0f && __fadd_rn( u, v ) <= 1.0f)
						{
							const float t = __fmul_rn( Dot3( e2.x, e2.y, e2.z, qvx, qvy, qvz ), inv );
							if (ANYHIT)
							{
								if (t > tmin && t < tmax) return true;
							}
							else if (t > tmin)
							{
								const uint32_t prim = __float_as_uint( v0.w );
								const uint32_t inst = scene.singleIdentity ? __float_as_uint( e1.w ) : curInst;	// flat scenes: from the triangle record
								const bool closer = t < tmax || (t == tmax && (inst < bestInst || (inst == bestInst && prim < bestPrim)));
								if (closer) tmax = t, bestInst = inst, bestPrim = prim, bestU = u, bestV = v;
							}
						}
					}
				}
			}
			else
			{
				// top-level leaf: enter the instance
				const uint32_t inst = __ldg( scene.tlasLeafIds + tg.x + bit );
				if (tg.y != 0) stack[sp++] = tg;
				if (ng.y > 0x00ffffffu) stack[sp++] = ng;
				stack[sp++] = make_uint2( 0, 0 ); // sentinel: return to the top level
				const InstTrav& it = scene.instances[inst];
				const float4 r0 = it.r0, r1 = it.r1, r2 = it.r2;
				O.x = __fmaf_rn( r0.x, wO.x, __fmaf_rn( r0.y, wO.y, __fmaf_rn( r0.z, wO.z, r0.w ) ) );
				O.y = __fmaf_rn( r1.x, wO.x, __fmaf_rn( r1.y, wO.y, __fmaf_rn( r1.z, wO.z, r1.w ) ) );
				O.z = __fmaf_rn( r2.x, wO.x, __fmaf_rn( r2.y, wO.y, __fmaf_rn( r2.z, wO.z, r2.w ) ) );
				D.x = __fmaf_rn( r0.x, wD.x, __fmaf_rn( r0.y, wD.y, __fmul_rn( r0.z, wD.z ) ) );
				D.y = __fmaf_rn( r1.x, wD.x, __fmaf_rn( r1.y, wD.y, __fmul_rn( r1.z, wD.z ) ) );
				D.z = __fmaf_rn( r2.x, wD.x, __fmaf_rn( r2.y, wD.y, __fmul_rn( r2.z, wD.z ) ) );
				idx = SafeRcpDir( D.x ), idy = SafeRcpDir( D.y ), idz = SafeRcpDir( D.z );
				octinv = (D.x < 0 ? 0 : 4) | (D.y < 0 ? 0 : 2) | (D.z < 0 ? 0 : 1);
				octinv4 = octinv * 0x01010101u;
				curInst = inst, inBlas = true;
				ng = make_uint2( it.rootNode, 0x80000000u ), tg = make_uint2( 0, 0 );
				break;
			}
		}
		if (ng.y <= 0x00ffffffu)
		{
			bool done = false;
			while (true)
			{
				if (sp == 0) { done = true; break; }
				ng = stack[--sp];
				if (ng.y != 0) break;
				// sentinel: leave the instance, restore the world-space ray
				O = wO, D = wD;
				idx = SafeRcpDir( D.x ), idy = SafeRcpDir( D.y ), idz = SafeRcpDir( D.z );
				octinv = (D.x < 0 ? 0 : 4) | (D.y < 0 ? 0 : 2) | (D.z < 0 ? 0 : 1);
				octinv4 = octinv * 0x01010101u;
				inBlas = false;
			}
			if (done) break;
		}
	}
	if (ANYHIT) return false;
	res.t = tmax, res.inst = bestInst, res.prim = bestPrim, res.u = bestU, res.v = bestV;
	return bestPrim != 0xffffffffu;
}

} // namespace lh2b
