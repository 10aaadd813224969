/* bvh.h - acceleration-structure data layout shared by host builders and device kernels.

   The reference has no BVH code of its own: build and traversal live inside OptiX
   (optixAccelBuild at lib/rendercore_optix7/core_mesh.cpp:67-129 and rendercore.cpp:767-797,
   optixTrace at lib/rendercore_optix7/optix/.optix.cu:125,136,148). This file defines the
   B200 replacement: an 8-wide BVH in the spirit of the compressed wide BVH ("CWBVH", Ylitie,
   Karras, Laine 2017) - eight child slots in octant order, child boxes stored relative to the
   node - re-laid-out for sm_100a:

   * a node is ONE 128-byte cache line, read as four 32-byte sectors with four 256-bit loads
     (LDG.E.256, new on sm_100): header | x planes | y planes | z planes;
   * child planes are bfloat16 OFFSETS from the node's (padded) minimum corner, so a plane is
     turned into a float with zero instructions (odd children: the word as it is) or one shift
     (even children) - the first design's 8-bit planes cost one PRMT each (48 per node) on the
     ALU pipe, which ncu showed to be the pipe that limits traversal (profiles/r1_v4_*);
   * every leaf slot holds exactly ONE triangle (or one instance in a top-level node), so the
     per-node hit mask is 8 bits in slot order and needs no per-child variable shifts.

   Node layout (32 x uint32):
     w[0..2]   pmin.xyz   float bits: node minimum, moved out by the padding (see CwEncodePlanes)
     w[3]      imask | lmask << 8     bit s of imask: slot s holds an internal node; of lmask: a leaf
     w[4]      childBase  arena index of the first internal child; the child in slot s is
                          childBase + popcount( imask & ((1 << s) - 1) )
     w[5]      triBase    arena index of the first leaf triangle (top level: first leaf id);
                          the leaf in slot s is triBase + popcount( lmask & ((1 << s) - 1) )
     w[6..7]   reserved (0)
     w[8..11]  lo.x   w[12..15] hi.x   w[16..19] lo.y   w[20..23] hi.y   w[24..27] lo.z   w[28..31] hi.z
               each group: 8 bfloat16, slot s in 16-bit lane s (word s / 2, low half for even s)
   Decoding (traverse_wide.cuh): offset( even s ) = float( word << 16 ); offset( odd s ) = float( word ), i.e. the
   low half of the word - the even neighbour's plane - rides along as extra mantissa bits. The encoder knows those
   bits and makes every decoded plane conservative: lo planes never above, hi planes never below the true box, after
   the box was padded by delta (2^-16 of the node extent + 2^-21 of the coordinate magnitude). That padding is what
   keeps the float slab test consistent with the exact-order triangle test; traversal adds only a relative far-side
   pad for its own rounding.
   Empty slots carry lo = +3.39e38, hi = -3.39e38 (bf16 0x7F7F / 0xFF7F): they fail the slab test by themselves.

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

#define CW_NODE_WORDS 32
#define CW_NODE_QUADS 8		// uint4 per node
struct CwNode { uint32_t w[CW_NODE_WORDS]; };	// 128 bytes
static_assert( sizeof( CwNode ) == 128, "CwNode must be 128 bytes" );

/* Leaf triangle record used by traversal: 48 bytes, Moeller-Trumbore form.
   v0.w holds the primitive index (triangle index in the mesh) as int bits. */
struct CwTri { float v0[3]; int32_t prim; float e1[3]; uint32_t inst; float e2[3]; float pad2; };	// inst: owning instance in flat scenes (see core.cu)
static_assert( sizeof( CwTri ) == 48, "CwTri must be 48 bytes" );

LH2B_HD inline uint32_t CwFloatBits( const float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }
LH2B_HD inline float CwBitsFloat( const uint32_t u ) { float f; memcpy( &f, &u, 4 ); return f; }
/* largest bfloat16 <= x and smallest bfloat16 >= x, for x >= 0 */
LH2B_HD inline uint32_t CwBf16Down( const float x ) { return CwFloatBits( x ) >> 16; }
LH2B_HD inline uint32_t CwBf16Up( const float x ) { const uint32_t b = CwFloatBits( x ); return (b >> 16) + ((b & 0xffffu) ? 1u : 0u); }

/* Writes the geometry part of a node - w[0..2] and w[8..31] - from the node box and the boxes of the children in its
   slots (validMask: which slots are occupied). The only encoder: host collapse, GPU collapse and GPU refit all call it. */
LH2B_HD inline void CwEncodePlanes( uint32_t* w, const float* nodeLo, const float* nodeHi, const float (*childLo)[3], const float (*childHi)[3], const uint32_t validMask )
{
	for (int a = 0; a < 3; a++)
	{
		const float ext = nodeHi[a] - nodeLo[a], mag = fmaxf( fabsf( nodeLo[a] ), fabsf( nodeHi[a] ) );
		const float delta = ext * (1.0f / 65536.0f) + mag * (1.0f / 2097152.0f) + 1e-30f;
		const float ps = nodeLo[a] - 2 * delta;
		w[a] = CwFloatBits( ps );
		uint32_t L[8], H[8];
		float loOff[8];
		for (int s = 0; s < 8; s++)
		{
			if (!((validMask >> s) & 1)) { L[s] = 0x7F7Fu, H[s] = 0xFF7Fu, loOff[s] = 3e38f; continue; }
			float lo = (childLo[s][a] - ps) - delta;
			const float hi = (childHi[s][a] - ps) + delta;
			if (!(lo > 0)) lo = 0;
			loOff[s] = lo, L[s] = CwBf16Down( lo ), H[s] = CwBf16Up( hi > 0 ? hi : 0 );
		}
		// odd slots are decoded with the even neighbour's 16 bits as extra mantissa: a lo plane must stay below the true offset
		for (int s = 1; s < 8; s += 2) if ((validMask >> s) & 1)
			if (L[s] > 0 && CwBitsFloat( (L[s] << 16) | L[s - 1] ) > loOff[s]) L[s]--;
		for (int k = 0; k < 4; k++) w[8 + a * 8 + k] = L[2 * k] | (L[2 * k + 1] << 16), w[12 + a * 8 + k] = H[2 * k] | (H[2 * k + 1] << 16);
	}
}

/* The same decoding the traversal kernels do, for checks (host builder self-test, oracle reader is written independently). */
LH2B_HD inline void CwDecodeChildBox( const uint32_t* w, const int s, float* lo, float* hi )
{
	for (int a = 0; a < 3; a++)
	{
		const uint32_t wl = w[8 + a * 8 + (s >> 1)], wh = w[12 + a * 8 + (s >> 1)];
		const float p = CwBitsFloat( w[a] );
		lo[a] = p + CwBitsFloat( (s & 1) ? wl : (wl << 16) ), hi[a] = p + CwBitsFloat( (s & 1) ? wh : (wh << 16) );
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
