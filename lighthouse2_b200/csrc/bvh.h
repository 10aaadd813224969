/* bvh.h - acceleration-structure data layout shared by host builders and device kernels.

   The reference has no BVH code of its own: build and traversal live inside OptiX
   (optixAccelBuild at lib/rendercore_optix7/core_mesh.cpp:67-129 and rendercore.cpp:767-797,
   optixTrace at lib/rendercore_optix7/optix/.optix.cu:125,136,148). This file defines the
   B200 replacement: an 8-wide compressed BVH ("CWBVH", Ylitie, Karras, Laine 2017) with
   quantised child boxes, 80 bytes per node, read as five 16-byte loads.

   Node layout (uint4 x 5):
     q0: px, py, pz (float bits: node box minimum), {ex, ey, ez, imask} bytes
     q1: childBase (index of first internal child), triBase (index of first leaf triangle),
         meta[0..3], meta[4..7]
     q2: qlox[0..7]      q3.xy: qloz[0..7]   (see CW_* accessors below)
     ...
   Exact byte order:
     bytes  0..11  p (float3)
     bytes 12..14  e[3]   biased exponents: float(2^e) == uint_as_float(e << 23)
     byte  15      imask  bit s set <=> slot s holds an internal node
     bytes 16..19  childBase
     bytes 20..23  triBase
     bytes 24..31  meta[8]
     bytes 32..39  qlox[8]   40..47 qloy[8]   48..55 qloz[8]
     bytes 56..63  qhix[8]   64..71 qhiy[8]   72..79 qhiz[8]
   meta[s]: 0 = empty slot; internal: 0b001_11000 | s; leaf: (unary tri count in bits 5..7)
   | offset of first triangle relative to triBase (0..23).

   Slot s carries the child that sits on the {+/-x,+/-y,+/-z} side named by the bits of s
   (bit2 = +x, bit1 = +y, bit0 = +z), so "s ^ octinv" is a front-to-back priority.
*/
#pragma once
#include <stdint.h>
#include <vector>

namespace lh2b
{

struct CwNode { uint32_t w[20]; };	// 80 bytes
static_assert( sizeof( CwNode ) == 80, "CwNode must be 80 bytes" );

/* Leaf triangle record used by traversal: 48 bytes, Moeller-Trumbore form.
   v0.w holds the primitive index (triangle index in the mesh) as int bits. */
struct CwTri { float v0[3]; int32_t prim; float e1[3]; uint32_t inst; float e2[3]; float pad2; };	// inst: owning instance in flat scenes (see core.cu)
static_assert( sizeof( CwTri ) == 48, "CwTri must be 48 bytes" );

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

/* Host builders (bvh_build_cpu.cpp). */
void BuildBvh2SAH( const float* verts4, int triCount, std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& primIdx );
void BuildBvh2FromBoxes( const Aabb* boxes, int count, int maxLeaf, std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& primIdx );
/* linkedRoots (top level only): when given, leaf 'prim' is emitted as an INTERNAL child holding a copy of
   linkedRoots[prim] (the root node of that instance's BLAS, indices already absolute) - the flat-scene layout. */
void CollapseToCwBvh( const std::vector<Bvh2Node>& bvh2, const std::vector<uint32_t>& primIdx,
	const float* verts4 /* null for TLAS */, CwBvh& out, uint32_t nodeOffset = 0, uint32_t triOffset = 0, const CwNode* linkedRoots = nullptr );
/* nodeOffset / triOffset are added to every childBase / triBase this call writes (absolute arena indices). */

} // namespace lh2b
