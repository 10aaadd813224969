/* traverse_wide.cuh - persistent, warp-cooperative CWBVH traversal (sm_100a).

   Same contract and bit-identical results as traverse.cuh (see there for the reference call sites it replaces);
   this is the throughput path. What the first ncu profile (profiles/r1_v0_generateExtend_full.txt) showed for the
   per-thread while-while loop: triangle tests ran with 3 of 32 lanes, node steps with 15 of 32, and the byte->float
   conversions (48 I2F.U8 per node step, quarter-rate XU pipe) dominated the stall samples. Hence:

   1. The warp, not the thread, owns the loop. Every iteration all 32 lanes vote (ballot) once on what they hold,
      then the warp runs a triangle step for the lanes holding a triangle group, followed by a node step for the
      lanes holding a node group (the "if-if" shape: both steps run convergent, each at most once per iteration).
      Triangle groups found while a lane still has one pending are parked on a per-lane deferred stack.
      TRI_THRESHOLD can postpone the triangle step until that many lanes have triangles; measured on the B200
      (1 M-triangle terrain: primary / shadow / diffuse rays) the best value is 1, i.e. never postpone.
   2. Persistent threads: lanes whose ray is finished fetch the next ray index from a global counter with one
      warp-aggregated atomicAdd, so the warp stays full until the launch runs out of rays.
   3. Quantised plane bytes are turned into floats with one PRMT each (byte dropped into the mantissa of 65536.0f:
      value 65536 + 2q), folded into the slab FMA; the 2^-9-quantum rounding this adds is covered by separate,
      outward-rounded near/far offsets.
   4. The node stack head lives in shared memory ([entry][thread] layout: conflict-free for any per-lane depth),
      overflow and the deferred-triangle stack in local memory.

   Closest-hit order independence (tie-break on (instance, primitive)) is what makes deferring legal: the result
   does not depend on the order in which triangles are tested.
*/
#pragma once
#include "traverse.cuh"

namespace lh2b
{

#define WIDE_BLOCK 128
#define WIDE_SMEM_STACK 12		// node-stack entries per thread kept in shared memory
#define WIDE_LOCAL_STACK 40		// overflow entries in local memory
#define WIDE_TRI_STACK 12		// deferred triangle groups per lane
#define WIDE_TRI_THRESHOLD 1	// lanes with pending triangles that trigger a triangle step
#define WIDE_REFILL_THRESHOLD 8	// idle lanes that trigger fetching new rays

/* 0x47800000 = 65536.0f; dropping a byte into mantissa bits 8..15 gives exactly 65536 + 2 * byte. 'base' holds the
   constant in a register (see OpaqueBase) so that the selector can be the instruction's immediate: one PRMT per plane. */
template <int J> __device__ __forceinline__ float ByteFloat( const uint32_t word, const uint32_t base )
{
	uint32_t r;
	asm( "prmt.b32 %0, %1, %2, %3;" : "=r"( r ) : "r"( word ), "r"( base ), "n"( 0x7604 | (J << 4) ) );
	return __uint_as_float( r );
}
__device__ __forceinline__ uint32_t OpaqueBase()
{
	uint32_t k;
	asm volatile( "mov.b32 %0, 0x47800000;" : "=r"( k ) );	// volatile: the optimiser must not fold it back into an immediate
	return k;
}

struct WideRay
{
	float3 O, D;
	float tmin, tmax;
};

/* RaySource: bool Load( uint32_t workIdx, WideRay&, uint32_t& tag ) (false: nothing to trace for this index; tag is
              an opaque per-ray word handed back to the sink, e.g. the path index)
   HitSink:   void Closest( uint32_t tag, bool hit, const TraceResult& ) / void AnyHit( uint32_t tag, bool occluded )

   TWO_LEVEL: the lane is either in the top level (curInst == ~0: node groups index TLAS nodes, leaf groups are
   instances) or inside one instance. Entering an instance transforms the ray, pushes the remaining top-level work and
   a sentinel; the sentinel is only popped once every triangle of that instance (pending group + deferred stack) has
   been tested, because deferred triangle groups are meaningless outside their instance. */
struct WideTuning { int triThreshold, refillThreshold; };

template <bool ANYHIT, bool TWO_LEVEL, class RaySource, class HitSink>
__device__ __forceinline__ void TraverseWide( const DevScene& scene, RaySource& src, HitSink& sink, const uint32_t rayCount, uint32_t* workCounter,
	const WideTuning tune )
{
	__shared__ uint2 smemStack[WIDE_SMEM_STACK][WIDE_BLOCK];
	__shared__ float worldRay[TWO_LEVEL ? 6 : 1][WIDE_BLOCK];	// world-space O, D while the lane is inside a transformed instance
	uint2 localStack[WIDE_LOCAL_STACK];
	uint2 triStack[WIDE_TRI_STACK];
	const uint32_t lane = threadIdx.x & 31;
	const uint4* __restrict__ nodes = scene.nodes;
	const float4* __restrict__ tris = scene.tris;
	const uint32_t NO_INST = 0xffffffffu;
	const uint32_t fbase = OpaqueBase();
	// lane state
	bool active = false;
	uint32_t workIdx = 0;
	float3 O = make_float3( 0, 0, 0 ), D = make_float3( 0, 0, 1 );
	float idx = 0, idy = 0, idz = 0, tmin = 0, tmax = 0;
	uint32_t octinv = 0;
	uint2 ng = make_uint2( 0, 0 ), tg = make_uint2( 0, 0 );
	int sp = 0, tsp = 0;
	uint32_t curInst = TWO_LEVEL ? NO_INST : 0u;
	uint32_t bestPrim = 0xffffffffu, bestInst = 0xffffffffu;
	float bestU = 0, bestV = 0;
	bool exhausted = false;	// no more rays in the launch
#define WIDE_PUSH( e ) do { if (sp < WIDE_SMEM_STACK) smemStack[sp][threadIdx.x] = (e); else localStack[sp - WIDE_SMEM_STACK] = (e); sp++; } while (0)
#define WIDE_TOP() (sp <= WIDE_SMEM_STACK ? smemStack[sp - 1][threadIdx.x] : localStack[sp - 1 - WIDE_SMEM_STACK])
	while (true)
	{
		// ---- lanes without a node group take the next one from their stack; finished rays retire -------------
		if (active)
		{
			if (TWO_LEVEL)
			{
				if (curInst != NO_INST && tg.y == 0 && tsp > 0) tg = triStack[--tsp];
				if (ng.y <= 0x00ffffffu) while (sp > 0)
				{
					const uint2 top = WIDE_TOP();
					if (top.y > 0x00ffffffu) { ng = top, sp--; break; }	// node group
					if (top.y != 0)
					{
						// top-level leaf group (instances): only reachable in the top level
						if (tg.y == 0) tg = top, sp--;
						break;
					}
					// sentinel: leave the instance once all of its triangles are done
					if (tg.y != 0 || tsp > 0) break;
					sp--, curInst = NO_INST;
					if (top.x != 0)
					{
						// the instance had a transform: restore the world-space ray
						O = make_float3( worldRay[0][threadIdx.x], worldRay[1][threadIdx.x], worldRay[2][threadIdx.x] );
						D = make_float3( worldRay[3][threadIdx.x], worldRay[4][threadIdx.x], worldRay[5][threadIdx.x] );
						idx = SafeRcpDir( D.x ), idy = SafeRcpDir( D.y ), idz = SafeRcpDir( D.z );
						octinv = (D.x < 0 ? 0 : 4) | (D.y < 0 ? 0 : 2) | (D.z < 0 ? 0 : 1);
					}
				}
				if (curInst == NO_INST && tg.y != 0)
				{
					// enter one instance of the pending top-level leaf group
					const int bit = 31 - __clz( tg.y );
					tg.y &= ~(1u << bit);
					const uint32_t inst = __ldg( scene.tlasLeafIds + tg.x + bit );
					if (tg.y != 0) WIDE_PUSH( tg );
					if (ng.y > 0x00ffffffu) WIDE_PUSH( ng );
					const InstTrav& it = scene.instances[inst];
					const bool transformed = !(it.flags & 1u);
					WIDE_PUSH( make_uint2( transformed ? 1u : 0u, 0 ) );	// sentinel; x tells whether the ray must be restored
					if (transformed)
					{
						const float4 r0 = it.r0, r1 = it.r1, r2 = it.r2;
						const float3 wO = O, wD = D;
						worldRay[0][threadIdx.x] = O.x, worldRay[1][threadIdx.x] = O.y, worldRay[2][threadIdx.x] = O.z;
						worldRay[3][threadIdx.x] = D.x, worldRay[4][threadIdx.x] = D.y, worldRay[5][threadIdx.x] = D.z;
						O.x = __fmaf_rn( r0.x, wO.x, __fmaf_rn( r0.y, wO.y, __fmaf_rn( r0.z, wO.z, r0.w ) ) );
						O.y = __fmaf_rn( r1.x, wO.x, __fmaf_rn( r1.y, wO.y, __fmaf_rn( r1.z, wO.z, r1.w ) ) );
						O.z = __fmaf_rn( r2.x, wO.x, __fmaf_rn( r2.y, wO.y, __fmaf_rn( r2.z, wO.z, r2.w ) ) );
						D.x = __fmaf_rn( r0.x, wD.x, __fmaf_rn( r0.y, wD.y, __fmul_rn( r0.z, wD.z ) ) );
						D.y = __fmaf_rn( r1.x, wD.x, __fmaf_rn( r1.y, wD.y, __fmul_rn( r1.z, wD.z ) ) );
						D.z = __fmaf_rn( r2.x, wD.x, __fmaf_rn( r2.y, wD.y, __fmul_rn( r2.z, wD.z ) ) );
						idx = SafeRcpDir( D.x ), idy = SafeRcpDir( D.y ), idz = SafeRcpDir( D.z );
						octinv = (D.x < 0 ? 0 : 4) | (D.y < 0 ? 0 : 2) | (D.z < 0 ? 0 : 1);
					}
					curInst = inst;
					ng = make_uint2( it.rootNode, 0x80000000u ), tg = make_uint2( 0, 0 );
				}
			}
			if (ng.y <= 0x00ffffffu && tg.y == 0)
			{
				// nothing left for this ray (two-level: the stack is empty here, the loop above ran dry)
				if (ANYHIT) sink.AnyHit( workIdx, false );
				else
				{
					TraceResult r;
					r.t = tmax, r.inst = bestInst, r.prim = bestPrim, r.u = bestU, r.v = bestV;
					sink.Closest( workIdx, bestPrim != 0xffffffffu, r );
				}
				active = false;
			}
		}
		// ---- refill idle lanes -------------------------------------------------------------------------
		const uint32_t idleMask = __ballot_sync( 0xffffffffu, !active );
		if (!exhausted && __popc( idleMask ) >= tune.refillThreshold)
		{
			const int leader = __ffs( idleMask ) - 1;
			uint32_t base = 0;
			if (lane == leader) base = atomicAdd( workCounter, (uint32_t)__popc( idleMask ) );
			base = __shfl_sync( 0xffffffffu, base, leader );
			if (base >= rayCount) exhausted = true;
			if (!active)
			{
				const uint32_t mine = base + __popc( idleMask & ((1u << lane) - 1) );
				WideRay ray;
				if (mine < rayCount && src.Load( mine, ray, workIdx ))
				{
					active = true;
					O = ray.O, D = ray.D, tmin = ray.tmin, tmax = ray.tmax;
					idx = SafeRcpDir( D.x ), idy = SafeRcpDir( D.y ), idz = SafeRcpDir( D.z );
					octinv = (D.x < 0 ? 0 : 4) | (D.y < 0 ? 0 : 2) | (D.z < 0 ? 0 : 1);
					ng = make_uint2( TWO_LEVEL ? scene.tlasRoot : scene.singleRoot, 0x80000000u ), tg = make_uint2( 0, 0 );
					sp = 0, tsp = 0, bestPrim = 0xffffffffu, bestInst = 0xffffffffu, bestU = bestV = 0;
					curInst = TWO_LEVEL ? NO_INST : 0u;
				}
			}
		}
		// ---- vote ------------------------------------------------------------------------------------
		const bool hasNode = active && ng.y > 0x00ffffffu;
		const bool hasTri = active && tg.y != 0 && (!TWO_LEVEL || curInst != NO_INST);
		const uint32_t nodeMask = __ballot_sync( 0xffffffffu, hasNode ), triMask = __ballot_sync( 0xffffffffu, hasTri );
		const uint32_t fullMask = __ballot_sync( 0xffffffffu, tsp >= WIDE_TRI_STACK - 1 );
		if ((nodeMask | triMask) == 0)
		{
			// no lane can do a node or a triangle step: either nothing is active, or (two-level) lanes are about to
			// enter an instance at the top of the next iteration
			if (exhausted && __ballot_sync( 0xffffffffu, active ) == 0) break;
			continue;
		}
		const bool triPhase = triMask != 0 && (nodeMask == 0 || __popc( triMask ) >= tune.triThreshold || fullMask != 0);
		if (triPhase)
		{
			if (hasTri)
			{
				const int bit = 31 - __clz( tg.y );
				tg.y &= ~(1u << bit);
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
								if (t > tmin && t < tmax)
								{
									sink.AnyHit( workIdx, true );
									active = false, ng.y = 0, tg.y = 0, sp = 0, tsp = 0;
								}
							}
							else if (t > tmin)
							{
								const uint32_t prim = __float_as_uint( v0.w );
								const uint32_t inst = TWO_LEVEL ? curInst : __float_as_uint( e1.w );	// flat scenes: from the triangle record
								const bool closer = t < tmax || (t == tmax && (inst < bestInst || (inst == bestInst && prim < bestPrim)));
								if (closer) tmax = t, bestPrim = prim, bestInst = inst, bestU = u, bestV = v;
							}
						}
					}
				}
				if (!TWO_LEVEL && active && tg.y == 0 && tsp > 0) tg = triStack[--tsp];
			}
			// lanes with a pending node take their node step in the same iteration: one vote for both phases
			if (nodeMask == 0 || fullMask != 0) continue;	// (a full deferred-triangle stack drains before it may grow again)
		}
		// ---- node phase --------------------------------------------------------------------------------
		if (hasNode && (!ANYHIT || active))
		{
			const uint32_t hits = ng.y;
			const int bit = 31 - __clz( hits );
			ng.y &= ~(1u << bit);
			if (ng.y > 0x00ffffffu) WIDE_PUSH( ng );
			const uint32_t slot = (uint32_t)(bit - 24) ^ octinv;
			const uint32_t rel = __popc( hits & ~(0xffffffffu << slot) & 0xffu );
			const uint4* np = nodes + (size_t)(ng.x + rel) * 5;
			const uint4 n0 = __ldg( np ), n1 = __ldg( np + 1 ), n2 = __ldg( np + 2 ), n3 = __ldg( np + 3 ), n4 = __ldg( np + 4 );
			// t(q) = (p + q * 2^e - o) * idir = (65536 + 2q) * (2^e * idir / 2) + ((p - o) * idir - 32768 * 2^e * idir)
			const float sx = __uint_as_float( (n0.w & 255u) << 23 ) * idx, sy = __uint_as_float( ((n0.w >> 8) & 255u) << 23 ) * idy;
			const float sz = __uint_as_float( ((n0.w >> 16) & 255u) << 23 ) * idz;
			const float hx = 0.5f * sx, hy = 0.5f * sy, hz = 0.5f * sz;
			const float cx = fmaf( -32768.0f, sx, (__uint_as_float( n0.x ) - O.x) * idx );
			const float cy = fmaf( -32768.0f, sy, (__uint_as_float( n0.y ) - O.y) * idy );
			const float cz = fmaf( -32768.0f, sz, (__uint_as_float( n0.z ) - O.z) * idz );
			// outward slack of 2^-8 quantum on both sides covers the rounding of the shifted offset
			const float ex = 0.00390625f * fabsf( sx ), ey = 0.00390625f * fabsf( sy ), ez = 0.00390625f * fabsf( sz );
			const float cnx = cx - ex, cfx = cx + ex, cny = cy - ey, cfy = cy + ey, cnz = cz - ez, cfz = cz + ez;
			const uint32_t octinv4 = octinv * 0x01010101u;
			ng.x = n1.x;
			uint2 ntg = make_uint2( n1.y, 0 );
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
#define WIDE_CHILD( J ) { \
					const float t0x = fmaf( ByteFloat<J>( nx, fbase ), hx, cnx ), t1x = fmaf( ByteFloat<J>( fx, fbase ), hx, cfx ); \
					const float t0y = fmaf( ByteFloat<J>( ny, fbase ), hy, cny ), t1y = fmaf( ByteFloat<J>( fy, fbase ), hy, cfy ); \
					const float t0z = fmaf( ByteFloat<J>( nz, fbase ), hz, cnz ), t1z = fmaf( ByteFloat<J>( fz, fbase ), hz, cfz ); \
					const float cmin = fmaxf( fmaxf( t0x, t0y ), fmaxf( t0z, tmin ) ); \
					const float cmax = fminf( fminf( t1x, t1y ), fminf( t1z, tmax ) ) * 1.0000005f; \
					if (cmin <= cmax) hitmask |= ((childBits4 >> (8 * J)) & 255u) << ((bitIndex4 >> (8 * J)) & 255u); }
				WIDE_CHILD( 0 ) WIDE_CHILD( 1 ) WIDE_CHILD( 2 ) WIDE_CHILD( 3 )
#undef WIDE_CHILD
			}
			ng.y = (hitmask & 0xff000000u) | (n0.w >> 24);
			ntg.y = hitmask & 0x00ffffffu;
			if (ntg.y != 0)
			{
				if (tg.y == 0) tg = ntg;
				else triStack[tsp++] = ntg;
			}
			if (!TWO_LEVEL && ng.y <= 0x00ffffffu && sp > 0) { ng = WIDE_TOP(); sp--; }
		}
	}
#undef WIDE_PUSH
#undef WIDE_TOP
}

} // namespace lh2b
