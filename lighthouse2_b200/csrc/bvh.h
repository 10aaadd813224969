/* bvh.h - acceleration-structure data layout shared by host builders and device kernels.

   The reference has no BVH code of its own: build and traversal live inside OptiX
   (optixAccelBuild at lib/rendercore_optix7/core_mesh.cpp:67-129 and rendercore.cpp:767-797,
   optixTrace at lib/rendercore_optix7/optix/.optix.cu:125,136,148). This file defines the
   B200 replacement: an 8-wide compressed BVH ("CWBVH", Ylitie, Karras, Laine 2017) - eight child
   slots in octant order, child boxes quantised to 8 bits on a per-node grid - with the node
   re-laid-out for the instruction stream and the L1 data path of sm_100a (DESIGN.md section 3 has the
   ncu numbers behind each choice):

   * 80 bytes per node, read as five 16-byte loads. Bytes matter: for divergent rays the L1 returns
     16 bytes per unique address per cycle, so traversal time tracks the bytes a node step loads.
     (A 128-byte variant with bfloat16 planes - free decoding - was built and measured in round 2:
     20 % fewer instructions, but the L1 data pipe went from 48 % to 83 % busy and the time did not move.)
   * every leaf slot holds exactly ONE triangle (or one instance in a top-level node): the per-node
     hit mask is 8 bits in slot order, AND-ed with two slot masks - no per-child variable shifts;
   * the node origin is stored pre-biased and the axis scales as ready-made floats, so that a plane byte
     dropped into a float mantissa (one PRMT) is all the decoding there is.

   Node layout (20 x uint32):
     w[0..2]   pb.xyz     float bits: biased grid origin, pb = p - 32768 * 2^e per axis (p = padded node minimum)
     w[3]      bfloat16( 2^(ex-1) ) << 16 | bfloat16( 2^(ey-1) )        half the grid spacing of x and y, as float halves
     w[4]      childBase  arena index of the first internal child; the child in slot s is
                          childBase + popcount( imask & ((1 << s) - 1) )
     w[5]      triBase    arena index of the first leaf triangle (top level: first leaf id);
                          the leaf in slot s is triBase + popcount( lmask & ((1 << s) - 1) )
     w[6]      bfloat16( 2^(ez-1) ) << 16 | lmask << 8 | imask           bit s of imask: slot s is an internal node; lmask: a leaf
     w[7]      reserved (0)
     w[8..9]   qlo.x[8]   w[10..11] qlo.y[8]   w[12..13] qlo.z[8]   w[14..15] qhi.x[8]   w[16..17] qhi.y[8]   w[18..19] qhi.z[8]
               one byte per slot, slot s in byte s
   Decoding (traverse_wide.cuh): plane( q ) = pb + (32768 + q) * 2^e, evaluated as f * h + pb with f = 65536 + 2q (the byte
   dropped into the mantissa of 65536.0f) and h = 2^(e-1). The encoder quantises every child box against exactly these decoded
   planes, after padding the box by delta (2^-14 of the node extent + 2^-21 of the coordinate magnitude; the traversal's own rounding of
   (pb - o) * idir is below 2^-16 of the extent): lo planes never above,
   hi planes never below the padded box. That padding keeps the float slab test consistent with the exact-order triangle test;
   traversal adds only a relative far-side pad for its own rounding. Empty slots (neither mask bit set) carry qlo = 255, qhi = 0.

   Slot s carries the child that sits on the {+/-x,+/-y,+/-z} side named by the bits of s
   (bit2 = +x, bit1 = +y, bit0 = +z), so "s ^ octinv" is a front-to-back priority.
*/
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <vector>

#ifdef __CUDACC__
#define LH2B_HD __host__ __device__
#else
#define LH2B_HD
#endif

namespace lh2b
{

#define CW_NODE_WORDS 20
#define CW_NODE_QUADS 5		// uint4 per node
struct CwNode { uint32_t w[CW_NODE_WORDS]; };	// 80 bytes
static_assert( sizeof( CwNode ) == 80, "CwNode must be 80 bytes" );

/* Leaf triangle record used by traversal: 48 bytes, Moeller-Trumbore form.
   v0.w holds the primitive index (triangle index in the mesh) as int bits. */
struct CwTri { float v0[3]; int32_t prim; float e1[3]; uint32_t inst; float e2[3]; float pad2; };	// inst: owning instance in flat scenes (see core.cu)
static_assert( sizeof( CwTri ) == 48, "CwTri must be 48 bytes" );

LH2B_HD inline uint32_t CwFloatBits( const float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }
LH2B_HD inline float CwBitsFloat( const uint32_t u ) { float f; memcpy( &f, &u, 4 ); return f; }

/* the plane a byte decodes to, with the float operations of the traversal's arithmetic model (exact product, one rounding) */
LH2B_HD inline float CwPlane( const float pb, const float spacing, const int q ) { return pb + (32768.0f + (float)q) * spacing; }

/* Writes the geometry part of a node - w[0..2], the three spacings in w[3] / w[6] and the plane bytes w[8..19] - from the node
   box and the boxes of the children in its slots (validMask: which slots are occupied). Keeps the low 16 bits of w[6] (the
   slot masks). The only encoder: host collapse, GPU collapse and GPU refit all call it. */
LH2B_HD inline void CwEncodePlanes( uint32_t* w, const float* nodeLo, const float* nodeHi, const float (*childLo)[3], const float (*childHi)[3], const uint32_t validMask )
{
	uint32_t halfSpacingBits[3];
	for (int k = 8; k < 20; k++) w[k] = 0;
	for (int a = 0; a < 3; a++)
	{
		const float ext = nodeHi[a] - nodeLo[a], mag = fmaxf( fabsf( nodeLo[a] ), fabsf( nodeHi[a] ) );
		const float delta = ext * (1.0f / 16384.0f) + mag * (1.0f / 2097152.0f) + 1e-30f;
		const float p = nodeLo[a] - delta, span = ext + 2 * delta;
		// grid spacing 2^e: 255 steps cover the padded extent
		int e = (int)ceilf( log2f( span * (1.0f / 255.0f) ) );
		if (e < -120) e = -120;
		if (e > 120) e = 120;
		while (e < 120 && ldexpf( 255.0f, e ) < span) e++;
		const float spacing = ldexpf( 1.0f, e ), pb = p - 32768.0f * spacing;
		w[a] = CwFloatBits( pb );
		halfSpacingBits[a] = CwFloatBits( ldexpf( 1.0f, e - 1 ) ) >> 16;	// a power of two: the low 16 bits are zero
		for (int s = 0; s < 8; s++)
		{
			int ql = 255, qh = 0;
			if ((validMask >> s) & 1)
			{
				const float lo = childLo[s][a] - delta, hi = childHi[s][a] + delta;
				ql = (int)floorf( (lo - p) / spacing ), qh = (int)ceilf( (hi - p) / spacing );
				ql = ql < 0 ? 0 : (ql > 255 ? 255 : ql), qh = qh < 0 ? 0 : (qh > 255 ? 255 : qh);
				// against the planes as they are decoded; half the padding absorbs the roundings of this check itself
				while (ql > 0 && CwPlane( pb, spacing, ql ) > lo + 0.5f * delta) ql--;
				while (qh < 255 && CwPlane( pb, spacing, qh ) < hi - 0.5f * delta) qh++;
			}
			w[8 + a * 2 + (s >> 2)] |= (uint32_t)ql << (8 * (s & 3));
			w[14 + a * 2 + (s >> 2)] |= (uint32_t)qh << (8 * (s & 3));
		}
	}
	w[3] = (halfSpacingBits[0] << 16) | halfSpacingBits[1];
	w[6] = (halfSpacingBits[2] << 16) | (w[6] & 0xffffu);
}

/* The decoded box of a slot, for checks (the oracle's reader is written independently from the layout comment above). */
LH2B_HD inline void CwDecodeChildBox( const uint32_t* w, const int s, float* lo, float* hi )
{
	const float spacing[3] = { 2 * CwBitsFloat( w[3] & 0xffff0000u ), 2 * CwBitsFloat( w[3] << 16 ), 2 * CwBitsFloat( w[6] & 0xffff0000u ) };
	for (int a = 0; a < 3; a++)
	{
		const int ql = (w[8 + a * 2 + (s >> 2)] >> (8 * (s & 3))) & 255, qh = (w[14 + a * 2 + (s >> 2)] >> (8 * (s & 3))) & 255;
		lo[a] = CwPlane( CwBitsFloat( w[a] ), spacing[a], ql ), hi[a] = CwPlane( CwBitsFloat( w[a] ), spacing[a], qh );
	}
}

/* Plain binary BVH used as the builder intermediate (host SAH builder and GPU LBVH both emit this). */
struct Bvh2Node
{
	float lo[3]; int32_t left;		// left >= 0: internal, children at left/right; left < 0: leaf
	float hi[3]; int32_t right;		// leaf: first = ~left, count = right
};

struct Aabb { float lo[3], hi[3]; };

struct CwBvh
{
	std::vector<CwNode> nodes;
	std::vector<CwTri> tris;		// for a BLAS
	std::vector<uint32_t> leafIds;	// for a TLAS: instance index per leaf slot
	Aabb bounds;
};

/* Host builders (bvh_build_cpu.cpp). Leaves of the binary tree hold exactly one primitive. */
void BuildBvh2SAH( const float* verts4, int triCount, std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& primIdx );
void BuildBvh2FromBoxes( const Aabb* boxes, int count, std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& primIdx );
/* linkedRoots (top level only): when given, leaf 'prim' is emitted as an INTERNAL child holding a copy of
   linkedRoots[prim] (the root node of that instance's BLAS, indices already absolute) - the flat-scene layout. */
void CollapseToCwBvh( const std::vector<Bvh2Node>& bvh2, const std::vector<uint32_t>& primIdx,
	const float* verts4 /* null for TLAS */, CwBvh& out, uint32_t nodeOffset = 0, uint32_t triOffset = 0, const CwNode* linkedRoots = nullptr );
/* nodeOffset / triOffset are added to every childBase / triBase this call writes (absolute arena indices). */

} // namespace lh2b
