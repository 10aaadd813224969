/* traverse_wide.cuh - persistent, warp-cooperative traversal of the 8-wide BVH of bvh.h (sm_100a).

   Contract: traverse.cuh (the reference call sites it replaces are listed there). What shaped it, with the ncu evidence:

   Round 1 (profiles/r1_v*): the per-thread while-while loop ran triangle tests with 3 of 32 lanes and node steps with 15 of
   32, hence the warp-cooperative shape kept here:
   1. The warp, not the thread, owns the loop. Every iteration all 32 lanes vote (ballot) once on what they hold, then the warp
      runs a triangle step for the lanes holding a triangle group, followed by a node step for the lanes holding a node group
      (the "if-if" shape: both steps run convergent, each at most once per iteration). Triangle groups found while a lane still
      has one pending are parked on a per-lane deferred stack.
   2. Persistent threads: lanes whose ray is finished fetch the next ray index from a global counter with one warp-aggregated
      atomicAdd, so the warp stays full until the launch runs out of rays.
   3. The node stack head lives in shared memory ([entry][thread] layout: conflict-free for any per-lane depth), overflow and
      the deferred-triangle stack in local memory.

   Round 2 (profiles/r1_v4_kernels_full.txt read again): the node step was 282 warp instructions, 2/3 of them on the ALU pipe,
   which issues at half rate - ALU pipe 68 % busy, FMA pipe 28 %, math-pipe-throttle stalls: ALU-pipe bound, not "issue bound".
   ~50 of those instructions assembled the hit mask with per-child variable shifts, ~30 decoded exponents and folded the bias.
   A first redesign (128-byte nodes, bfloat16 planes, decoding for free: profiles/r2_v5_*) cut the instructions by 20 % and moved
   the limit to the L1 data pipe (83 % busy: divergent lanes get 16 bytes per unique address per cycle, so bytes per node step
   are cycles) at unchanged speed. The format kept (bvh.h) stays at 80 bytes and spends its instructions where they are cheap:
   4. origin pre-biased and spacings stored as floats by the builder: the per-node set-up is 3 masks / shifts, 6 FMUL, 3 FADD;
      each plane byte becomes a float with one PRMT (byte dropped into the mantissa of 65536.0f), folded into the slab FFMA;
   5. the three per-child rejections (slab empty, behind tmax, behind the origin) are three values whose SIGN bits are OR-ed and
      shifted into an 8-bit miss mask with one funnel shift per child - differences and the far-side padding run on the FMA pipe
      (FFMA / FADD), only the two 3-input min / max per child stay on the ALU pipe;
   6. one triangle per leaf slot: the hit mask is 8 bits in slot order, AND-ed with the node's internal / leaf masks; the
      front-to-back choice among hit children is one shared-memory table look-up per pop (octant x mask -> slot);
   7. box padding against the triangle test's rounding is applied once by the builder, not per node step.

   Closest-hit order independence (tie-break on (instance, primitive)) is what makes deferring legal: the result does not depend
   on the order in which triangles are tested.
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

/* 0x47800000 = 65536.0f; dropping a byte into mantissa bits 8..15 gives exactly 65536 + 2 * byte: one PRMT per plane, with the byte
   selector as the instruction's immediate. For that the constant must NOT be an immediate too (ptxas then puts the selector into a
   register and re-materialises it before every PRMT: 45 extra moves per node step in the SASS of the first round-2 build): it
   travels as a kernel parameter (WideTuning::byteBase), i.e. a constant-bank operand of the PRMT. */
template <int J> __device__ __forceinline__ float ByteFloat( const uint32_t word, const uint32_t base )
{
	uint32_t r;
	asm( "prmt.b32 %0, %1, %2, %3;" : "=r"( r ) : "r"( word ), "r"( base ), "n"( 0x7604 | (J << 4) ) );
	return __uint_as_float( r );
}

struct WideRay
{
	float3 O, D;
	float tmin, tmax;	// the box test assumes tmin >= 0 (every ray source passes 0)
};

/* Work counters of a launch (STATS instantiations only; lh2b_trace_stats): what the issue roofline is computed from. */
struct TraceStats
{
	unsigned long long rays, nodeSteps, triTests, instanceEntries;	// per ray (lane) events
	unsigned long long iterations, nodePhases, triPhases;				// per warp events
	unsigned long long nodeLanes, triLanes;							// lanes active in those phases (SIMT utilisation of each step)
};

/* octant x child-hit mask -> slot to visit next: the hit slot s with the largest s ^ octinv (front to back) */
__device__ __forceinline__ void FillPopTable( uint8_t* table )
{
	for (int i = threadIdx.x; i < 2048; i += blockDim.x)
	{
		const int o = i >> 8, m = i & 255;
		int best = 0, bestPriority = -1;
		for (int s = 0; s < 8; s++) if (((m >> s) & 1) && (s ^ o) > bestPriority) bestPriority = s ^ o, best = s;
		table[i] = (uint8_t)best;
	}
}

/* RaySource: bool Load( uint32_t workIdx, WideRay&, uint32_t& tag ) (false: nothing to trace for this index; tag is
              an opaque per-ray word handed back to the sink, e.g. the path index)
   HitSink:   void Closest( uint32_t tag, bool hit, const TraceResult& ) / void AnyHit( uint32_t tag, bool occluded )

   Stack / group encoding (uint2): node group  = ( childBase, hitInner << 24 | imask )   - y > 0x00ffffff
                                   leaf group  = ( triBase,   hitLeaf | lmask << 8 )      - 0 < y <= 0x00ffffff
                                   sentinel    = ( restoreRay, 0 )                        - two-level only
   TWO_LEVEL: the lane is either in the top level (curInst == ~0: node groups index TLAS nodes, leaf groups are
   instances) or inside one instance. Entering an instance transforms the ray, pushes the remaining top-level work and
   a sentinel; the sentinel is only popped once every triangle of that instance (pending group + deferred stack) has
   been tested, because deferred triangle groups are meaningless outside their instance. */
struct WideTuning { int triThreshold, refillThreshold, raysPerLane; uint32_t byteBase; };	// byteBase: 0x47800000, see ByteFloat;	// raysPerLane: a launch with few rays uses fewer, fuller warps (blocks beyond rays / (128 * raysPerLane) retire at once)

template <bool ANYHIT, bool TWO_LEVEL, bool STATS, class RaySource, class HitSink>
__device__ __forceinline__ void TraverseWide( const DevScene& scene, RaySource& src, HitSink& sink, const uint32_t rayCount, uint32_t* workCounter,
	const WideTuning tune, TraceStats* statsOut = nullptr )
{
	__shared__ uint2 smemStack[WIDE_SMEM_STACK][WIDE_BLOCK];
	__shared__ float worldRay[TWO_LEVEL ? 6 : 1][WIDE_BLOCK];	// world-space O, D while the lane is inside a transformed instance
	__shared__ uint8_t popTable[2048];
	uint2 localStack[WIDE_LOCAL_STACK];
	uint2 triStack[WIDE_TRI_STACK];
	// few rays (small frames, deep path lengths): keep the persistent-thread refill working by giving every lane several rays instead of
	// spreading one ray per lane over the whole grid
	if (blockIdx.x > 0 && (unsigned long long)blockIdx.x * (WIDE_BLOCK * tune.raysPerLane) >= rayCount) return;
	FillPopTable( popTable );
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31;
	const uint4* __restrict__ nodes = scene.nodes;
	const uint32_t fbase = tune.byteBase;
	const float4* __restrict__ tris = scene.tris;
	const uint32_t NO_INST = 0xffffffffu;
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
	unsigned long long stRays = 0, stNodes = 0, stTris = 0, stInst = 0, stIter = 0, stNodePh = 0, stTriPh = 0, stNodeLanes = 0, stTriLanes = 0;
#define WIDE_PUSH( e ) do { if (sp < WIDE_SMEM_STACK) smemStack[sp][threadIdx.x] = (e); else localStack[sp - WIDE_SMEM_STACK] = (e); sp++; } while (0)
#define WIDE_TOP() (sp <= WIDE_SMEM_STACK ? smemStack[sp - 1][threadIdx.x] : localStack[sp - 1 - WIDE_SMEM_STACK])
	/* takes the highest hit leaf slot out of a leaf group and returns the index of its triangle / instance */
#define WIDE_TAKE_LEAF( g, index ) do { const int s_ = 31 - __clz( (g).y & 0xffu ); \
		index = (g).x + __popc( ((g).y >> 8) & ((1u << s_) - 1u) ); \
		(g).y &= ~(1u << s_); if (((g).y & 0xffu) == 0) (g).y = 0; } while (0)
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
					uint32_t leaf;
					WIDE_TAKE_LEAF( tg, leaf );
					const uint32_t inst = __ldg( scene.tlasLeafIds + leaf );
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
					ng = make_uint2( it.rootNode, 0x01000001u ), tg = make_uint2( 0, 0 );
					if (STATS) stInst++;
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
					// the root is entered as "slot 0 of a group whose only child is the root" (imask = 1, hit = 1)
					ng = make_uint2( TWO_LEVEL ? scene.tlasRoot : scene.singleRoot, 0x01000001u ), tg = make_uint2( 0, 0 );
					sp = 0, tsp = 0, bestPrim = 0xffffffffu, bestInst = 0xffffffffu, bestU = bestV = 0;
					curInst = TWO_LEVEL ? NO_INST : 0u;
					if (STATS) stRays++;
				}
			}
		}
		// ---- vote ------------------------------------------------------------------------------------
		const bool hasNode = active && ng.y > 0x00ffffffu;
		const bool hasTri = active && tg.y != 0 && (!TWO_LEVEL || curInst != NO_INST);
		const uint32_t nodeMask = __ballot_sync( 0xffffffffu, hasNode ), triMask = __ballot_sync( 0xffffffffu, hasTri );
		const uint32_t fullMask = __ballot_sync( 0xffffffffu, tsp >= WIDE_TRI_STACK - 1 );
		if (STATS && lane == 0) stIter++;
		if ((nodeMask | triMask) == 0)
		{
			// no lane can do a node or a triangle step: either nothing is active, or (two-level) lanes are about to
			// enter an instance at the top of the next iteration
			if (exhausted && __ballot_sync( 0xffffffffu, active ) == 0) break;
			continue;
		}
		// a triangle step runs when enough lanes hold triangles, or when a lane holds nothing else (it would idle through the node step)
		const bool triPhase = triMask != 0 && (__popc( triMask ) >= tune.triThreshold || (triMask & ~nodeMask) != 0 || fullMask != 0);
		if (triPhase)
		{
			if (STATS && lane == 0) stTriPh++, stTriLanes += __popc( triMask );
			if (hasTri)
			{
				uint32_t triIdx;
				WIDE_TAKE_LEAF( tg, triIdx );
				if (STATS) stTris++;
				const float4* tp = tris + (size_t)triIdx * 3;
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
		if (STATS && lane == 0) stNodePh++, stNodeLanes += __popc( nodeMask );
		if (hasNode && (!ANYHIT || active))
		{
			if (STATS) stNodes++;
			const uint32_t hits = ng.y;
			const uint32_t slot = popTable[(octinv << 8) | (hits >> 24)];
			ng.y = hits & ~(0x01000000u << slot);
			if (ng.y > 0x00ffffffu) WIDE_PUSH( ng );
			const uint32_t rel = __popc( hits & 0xffu & ((1u << slot) - 1u) );
			const uint4* np = nodes + (size_t)(ng.x + rel) * 5;
			const uint4 n0 = __ldg( np ), n1 = __ldg( np + 1 ), n2 = __ldg( np + 2 ), n3 = __ldg( np + 3 ), n4 = __ldg( np + 4 );
			// t( q ) = ( pb + (32768 + q) * 2^e - o ) * idir = f * h + c,  f = 65536 + 2q (ByteFloat), h = 2^(e-1) * idir, c = (pb - o) * idir
			const float hx = __uint_as_float( n0.w & 0xffff0000u ) * idx, hy = __uint_as_float( n0.w << 16 ) * idy, hz = __uint_as_float( n1.z & 0xffff0000u ) * idz;
			const float cx = (__uint_as_float( n0.x ) - O.x) * idx, cy = (__uint_as_float( n0.y ) - O.y) * idy, cz = (__uint_as_float( n0.z ) - O.z) * idz;
			const bool negx = !(octinv & 4), negy = !(octinv & 2), negz = !(octinv & 1);
			uint32_t miss = 0;
#pragma unroll
			for (int half = 1; half >= 0; half--)
			{
				// plane words: qlo x n2.xy, y n2.zw, z n3.xy; qhi x n3.zw, y n4.xy, z n4.zw (slots 0-3, 4-7); near = the plane the ray meets first
				const uint32_t qlox = half ? n2.y : n2.x, qloy = half ? n2.w : n2.z, qloz = half ? n3.y : n3.x;
				const uint32_t qhix = half ? n3.w : n3.z, qhiy = half ? n4.y : n4.x, qhiz = half ? n4.w : n4.z;
				const uint32_t nx = negx ? qhix : qlox, fx = negx ? qlox : qhix;
				const uint32_t ny = negy ? qhiy : qloy, fy = negy ? qloy : qhiy;
				const uint32_t nz = negz ? qhiz : qloz, fz = negz ? qloz : qhiz;
#define WIDE_CHILD( J ) { \
				const float t0x = fmaf( ByteFloat<J>( nx, fbase ), hx, cx ), t1x = fmaf( ByteFloat<J>( fx, fbase ), hx, cx ); \
				const float t0y = fmaf( ByteFloat<J>( ny, fbase ), hy, cy ), t1y = fmaf( ByteFloat<J>( fy, fbase ), hy, cy ); \
				const float t0z = fmaf( ByteFloat<J>( nz, fbase ), hz, cz ), t1z = fmaf( ByteFloat<J>( fz, fbase ), hz, cz ); \
				const float cmin = fmaxf( fmaxf( t0x, t0y ), t0z ), cmax = fminf( fminf( t1x, t1y ), t1z ); \
				/* miss <=> the slab interval is empty (far side padded by a few ulp: this arithmetic differs from the exact triangle \
				   test), or starts behind tmax, or ends behind the origin: the OR of three sign bits */ \
				const float d1 = fmaf( cmax, 1.0000005f, -cmin ), d2 = tmax - cmin; \
				const uint32_t sign = __float_as_uint( d1 ) | __float_as_uint( d2 ) | __float_as_uint( cmax ); \
				miss = __funnelshift_l( sign, miss, 1 ); }	/* slots arrive 7, 6, .. 0: slot s ends up in bit s */
				WIDE_CHILD( 3 ) WIDE_CHILD( 2 ) WIDE_CHILD( 1 ) WIDE_CHILD( 0 )
#undef WIDE_CHILD
			}
			uint32_t h[8];
			h[3] = n1.z, h[4] = n1.x, h[5] = n1.y;	// slot masks (imask | lmask << 8 in the low half), childBase, triBase
			const uint32_t hitInner = ~miss & h[3] & 0xffu, hitLeaf = ~miss & (h[3] >> 8) & 0xffu;
			ng = make_uint2( h[4], (hitInner << 24) | (h[3] & 0xffu) );
			if (hitLeaf != 0)
			{
				const uint2 ntg = make_uint2( h[5], hitLeaf | (h[3] & 0xff00u) );
				if (tg.y == 0) tg = ntg;
				else triStack[tsp++] = ntg;
			}
			if (!TWO_LEVEL && ng.y <= 0x00ffffffu && sp > 0) { ng = WIDE_TOP(); sp--; }
		}
	}
	if (STATS && statsOut)
	{
		for (int o = 16; o > 0; o >>= 1)
			stRays += __shfl_xor_sync( 0xffffffffu, stRays, o ), stNodes += __shfl_xor_sync( 0xffffffffu, stNodes, o ),
			stTris += __shfl_xor_sync( 0xffffffffu, stTris, o ), stInst += __shfl_xor_sync( 0xffffffffu, stInst, o );
		if (lane == 0)
		{
			atomicAdd( &statsOut->rays, stRays ), atomicAdd( &statsOut->nodeSteps, stNodes ), atomicAdd( &statsOut->triTests, stTris );
			atomicAdd( &statsOut->instanceEntries, stInst ), atomicAdd( &statsOut->iterations, stIter );
			atomicAdd( &statsOut->nodePhases, stNodePh ), atomicAdd( &statsOut->triPhases, stTriPh );
			atomicAdd( &statsOut->nodeLanes, stNodeLanes ), atomicAdd( &statsOut->triLanes, stTriLanes );
		}
	}
#undef WIDE_PUSH
#undef WIDE_TOP
#undef WIDE_TAKE_LEAF
}

} // namespace lh2b
