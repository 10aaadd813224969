/* traverse.cuh - scene view, ray-query contract and exact-order arithmetic helpers of the traversal kernels (traverse_wide.cuh).

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
	const uint4* nodes;				// node arena (128-byte nodes, bvh.h)
	const float4* tris;				// triangle arena (3 x float4 per triangle)
	const uint32_t* tlasLeafIds;	// instance index per top-level leaf slot
	const InstTrav* instances;
	uint32_t tlasRoot;				// arena index of the top-level root node
	int instanceCount;
	int singleIdentity;				// 1: exactly one instance with identity transform -> skip the top level
	uint32_t singleRoot;			// its BLAS root
	struct TraceStats* stats;		// non-null: launch the work-counting instantiations (lh2b_trace_stats)
};

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

} // namespace lh2b
