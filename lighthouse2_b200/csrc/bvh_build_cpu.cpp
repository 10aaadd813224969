/* bvh_build_cpu.cpp - host-side BVH construction: binned-SAH binary build (one primitive per leaf), SAH-optimal collapse
   to 8-wide, node encoding (bvh.h).

   Replaces optixAccelBuild (reference call sites: lib/rendercore_optix7/core_mesh.cpp:105,123 for
   triangle meshes, lib/rendercore_optix7/rendercore.cpp:795 for the instance level). The GPU LBVH
   builder (bvh_gpu.cu) emits the same Bvh2Node array and uses the same encoding rules; this host
   builder is the quality yardstick and the path for "build once" static meshes when the caller asks
   for it (Setting("bvhBuilder", 1)).
*/
#include "bvh.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <future>
#include <cfloat>

namespace lh2b
{

static inline float HalfArea( const float* lo, const float* hi )
{
	const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
	return ex * ey + ey * ez + ez * ex;
}

static inline void GrowBox( float* lo, float* hi, const float* blo, const float* bhi )
{
	for (int a = 0; a < 3; a++) lo[a] = std::min( lo[a], blo[a] ), hi[a] = std::max( hi[a], bhi[a] );
}

struct BuildCtx
{
	const Aabb* boxes;
	std::vector<float> cent;		// 3 per primitive
	uint32_t* idx;
	Bvh2Node* nodes;
	std::atomic<int> nodePtr;
};

static const int BINS = 16;

static void Subdivide( BuildCtx& c, const int nodeIdx, const int first, const int count, const int depth )
{
	Bvh2Node& node = c.nodes[nodeIdx];
	float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
	float clo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, chi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
	for (int i = 0; i < count; i++)
	{
		const uint32_t p = c.idx[first + i];
		GrowBox( lo, hi, c.boxes[p].lo, c.boxes[p].hi );
		for (int a = 0; a < 3; a++) clo[a] = std::min( clo[a], c.cent[p * 3 + a] ), chi[a] = std::max( chi[a], c.cent[p * 3 + a] );
	}
	memcpy( node.lo, lo, 12 ), memcpy( node.hi, hi, 12 );
	auto makeLeaf = [&]() { node.left = ~first; node.right = count; };
	if (count == 1) { makeLeaf(); return; }
	// binned SAH over the three axes
	float bestCost = FLT_MAX;
	int bestAxis = -1, bestSplit = 0;
	for (int a = 0; a < 3; a++)
	{
		const float ext = chi[a] - clo[a];
		if (!(ext > 0)) continue;
		const float scale = BINS / ext;
		int cnt[BINS] = {};
		float blo[BINS][3], bhi[BINS][3];
		for (int b = 0; b < BINS; b++) for (int k = 0; k < 3; k++) blo[b][k] = FLT_MAX, bhi[b][k] = -FLT_MAX;
		for (int i = 0; i < count; i++)
		{
			const uint32_t p = c.idx[first + i];
			const int b = std::min( BINS - 1, (int)((c.cent[p * 3 + a] - clo[a]) * scale) );
			cnt[b]++;
			GrowBox( blo[b], bhi[b], c.boxes[p].lo, c.boxes[p].hi );
		}
		float rightArea[BINS];
		int rightCnt[BINS];
		float rlo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, rhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
		int rc = 0;
		for (int b = BINS - 1; b > 0; b--)
		{
			if (cnt[b]) GrowBox( rlo, rhi, blo[b], bhi[b] );
			rc += cnt[b];
			rightCnt[b] = rc, rightArea[b] = rc ? HalfArea( rlo, rhi ) : 0;
		}
		float llo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, lhi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
		int lc = 0;
		for (int b = 0; b < BINS - 1; b++)
		{
			if (cnt[b]) GrowBox( llo, lhi, blo[b], bhi[b] );
			lc += cnt[b];
			if (lc == 0 || rightCnt[b + 1] == 0) continue;
			const float cost = HalfArea( llo, lhi ) * lc + rightArea[b + 1] * rightCnt[b + 1];
			if (cost < bestCost) bestCost = cost, bestAxis = a, bestSplit = b + 1;
		}
	}
	int mid;
	if (bestAxis >= 0)
	{
		const float scale = BINS / (chi[bestAxis] - clo[bestAxis]);
		const float base = clo[bestAxis];
		uint32_t* b = c.idx + first, * e = c.idx + first + count;
		uint32_t* m = std::partition( b, e, [&]( uint32_t p ) {
			return std::min( BINS - 1, (int)((c.cent[p * 3 + bestAxis] - base) * scale) ) < bestSplit; } );
		mid = (int)(m - b);
	}
	else mid = 0;
	if (mid == 0 || mid == count)
	{
		// degenerate (coincident centroids): median split on index order
		mid = count / 2;
	}
	const int l = c.nodePtr.fetch_add( 2 );
	node.left = l, node.right = l + 1;
	if (count > 32768 && depth < 6)
	{
		auto fut = std::async( std::launch::async, [&c, l, first, mid, depth]() { Subdivide( c, l, first, mid, depth + 1 ); } );
		Subdivide( c, l + 1, first + mid, count - mid, depth + 1 );
		fut.get();
	}
	else
	{
		Subdivide( c, l, first, mid, depth + 1 );
		Subdivide( c, l + 1, first + mid, count - mid, depth + 1 );
	}
}

void BuildBvh2FromBoxes( const Aabb* boxes, int count, std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& primIdx )
{
	nodes.assign( count > 0 ? 2 * count : 1, Bvh2Node{} );
	primIdx.resize( count );
	if (count == 0)
	{
		Bvh2Node& n = nodes[0];
		for (int a = 0; a < 3; a++) n.lo[a] = 0, n.hi[a] = 0;
		n.left = ~0, n.right = 0;
		nodes.resize( 1 );
		return;
	}
	BuildCtx c;
	c.boxes = boxes, c.idx = primIdx.data(), c.nodes = nodes.data();
	c.cent.resize( (size_t)count * 3 );
	for (int i = 0; i < count; i++)
	{
		primIdx[i] = i;
		for (int a = 0; a < 3; a++) c.cent[i * 3 + a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);
	}
	c.nodePtr = 1;
	Subdivide( c, 0, 0, count, 0 );
	nodes.resize( c.nodePtr.load() );
}

void BuildBvh2SAH( const float* verts4, int triCount, std::vector<Bvh2Node>& nodes, std::vector<uint32_t>& primIdx )
{
	std::vector<Aabb> boxes( triCount );
	for (int i = 0; i < triCount; i++)
	{
		const float* v = verts4 + (size_t)i * 12;
		for (int a = 0; a < 3; a++)
			boxes[i].lo[a] = std::min( v[a], std::min( v[4 + a], v[8 + a] ) ),
			boxes[i].hi[a] = std::max( v[a], std::max( v[4 + a], v[8 + a] ) );
	}
	BuildBvh2FromBoxes( boxes.data(), triCount, nodes, primIdx );
}

/* ---- collapse + encode ------------------------------------------------------------------ */

static inline bool IsLeaf( const Bvh2Node& n ) { return n.left < 0; }

/* Assign up to 8 children to slots 0..7 so that slot bits match the side of the node the child sits on. */
static void AssignSlots( const Bvh2Node* bvh2, const int* child, int n, const float* nodeLo, const float* nodeHi, int* slotOf )
{
	float score[8][8];
	const float cx = 0.5f * (nodeLo[0] + nodeHi[0]), cy = 0.5f * (nodeLo[1] + nodeHi[1]), cz = 0.5f * (nodeLo[2] + nodeHi[2]);
	for (int i = 0; i < n; i++)
	{
		const Bvh2Node& c = bvh2[child[i]];
		const float dx = 0.5f * (c.lo[0] + c.hi[0]) - cx, dy = 0.5f * (c.lo[1] + c.hi[1]) - cy, dz = 0.5f * (c.lo[2] + c.hi[2]) - cz;
		for (int s = 0; s < 8; s++)
			score[i][s] = ((s & 4) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 1) ? dz : -dz);
	}
	bool slotUsed[8] = {}, childDone[8] = {};
	for (int i = 0; i < n; i++) slotOf[i] = -1;
	for (int round = 0; round < n; round++)
	{
		float best = -FLT_MAX;
		int bi = -1, bs = -1;
		for (int i = 0; i < n; i++) if (!childDone[i]) for (int s = 0; s < 8; s++) if (!slotUsed[s])
			if (score[i][s] > best) best = score[i][s], bi = i, bs = s;
		slotOf[bi] = bs, slotUsed[bs] = true, childDone[bi] = true;
	}
}

void CollapseToCwBvh( const std::vector<Bvh2Node>& bvh2v, const std::vector<uint32_t>& primIdx, const float* verts4, CwBvh& out, uint32_t nodeOffset, uint32_t triOffset, const CwNode* linkedRoots )
{
	const Bvh2Node* bvh2 = bvh2v.data();
	out.nodes.clear(), out.tris.clear(), out.leafIds.clear();
	memcpy( out.bounds.lo, bvh2[0].lo, 12 ), memcpy( out.bounds.hi, bvh2[0].hi, 12 );
	struct Task { int bvh2Node, cwNode; };
	struct Link { int cwNode; uint32_t prim; };
	std::vector<Link> links;
	std::vector<Task> queue;
	out.nodes.push_back( CwNode{} );
	queue.push_back( { 0, 0 } );
	size_t head = 0;
	auto emitLeafPrim = [&]( uint32_t prim ) {
		if (verts4)
		{
			const float* v = verts4 + (size_t)prim * 12;
			CwTri t;
			for (int a = 0; a < 3; a++) t.v0[a] = v[a], t.e1[a] = v[4 + a] - v[a], t.e2[a] = v[8 + a] - v[a];
			t.prim = (int32_t)prim, t.inst = 0, t.pad2 = 0;
			out.tris.push_back( t );
		}
		else out.leafIds.push_back( prim );
	};
	// SAH-optimal collapse (same dynamic program as bvh_gpu.cu fitKernel; children are allocated after their parent, so a
	// descending sweep sees children first): T[n][i] = cheapest way to present subtree n as at most i slots, split[n][i-1] = the
	// number of slots the left child gets (0: T[n][i-1] kept), split[n][7] = the same for the node's own eight slots
	const float SAH_NODE = 1.0f, SAH_LEAF = 0.3f;
	std::vector<float> T( bvh2v.size() * 7 );
	std::vector<uint8_t> split( bvh2v.size() * 8, 0 );
	for (int i = (int)bvh2v.size() - 1; i >= 0; i--)
	{
		const Bvh2Node& nd = bvh2[i];
		if (IsLeaf( nd )) { for (int k = 0; k < 7; k++) T[(size_t)i * 7 + k] = HalfArea( nd.lo, nd.hi ) * SAH_LEAF; continue; }
		const float* tl = &T[(size_t)nd.left * 7], * tr = &T[(size_t)nd.right * 7];
		float d[9];
		uint8_t dk[9];
		for (int j = 2; j <= 8; j++)
		{
			float best = 3e38f;
			int bk = 1;
			for (int k = std::max( 1, j - 7 ); k <= std::min( 7, j - 1 ); k++)
			{
				const float v = tl[k - 1] + tr[j - k - 1];
				if (v < best) best = v, bk = k;
			}
			d[j] = best, dk[j] = (uint8_t)bk;
		}
		float* t = &T[(size_t)i * 7];
		uint8_t* sp = &split[(size_t)i * 8];
		t[0] = HalfArea( nd.lo, nd.hi ) * SAH_NODE + d[8], sp[0] = 0, sp[7] = dk[8];
		for (int k = 2; k <= 7; k++)
		{
			if (d[k] < t[k - 2]) t[k - 1] = d[k], sp[k - 1] = dk[k];
			else t[k - 1] = t[k - 2], sp[k - 1] = 0;
		}
	}
	while (head < queue.size())
	{
		const Task task = queue[head++];
		const Bvh2Node& root = bvh2[task.bvh2Node];
		// the node's children: unfold the choices of the dynamic program - (binary node, slots granted) pairs until each is one slot
		int child[8], n = 0;
		if (IsLeaf( root )) { if (root.right > 0) child[n++] = task.bvh2Node; }	// (an empty mesh is one node without children)
		else
		{
			int stackNode[8], stackSlots[8], sp = 0;
			const int k8 = split[(size_t)task.bvh2Node * 8 + 7];
			stackNode[sp] = root.right, stackSlots[sp++] = 8 - k8, stackNode[sp] = root.left, stackSlots[sp++] = k8;
			while (sp > 0)
			{
				const int nd = stackNode[--sp];
				int slots = stackSlots[sp];
				if (IsLeaf( bvh2[nd] )) { child[n++] = nd; continue; }
				while (slots > 1 && split[(size_t)nd * 8 + slots - 1] == 0) slots--;
				if (slots == 1) { child[n++] = nd; continue; }
				const int k = split[(size_t)nd * 8 + slots - 1];
				stackNode[sp] = bvh2[nd].right, stackSlots[sp++] = slots - k, stackNode[sp] = bvh2[nd].left, stackSlots[sp++] = k;
			}
		}
		int slotOf[8];
		AssignSlots( bvh2, child, n, root.lo, root.hi, slotOf );
		int childInSlot[8];
		for (int s = 0; s < 8; s++) childInSlot[s] = -1;
		for (int i = 0; i < n; i++) childInSlot[slotOf[i]] = child[i];
		// header + planes (bvh.h): one primitive per leaf slot, children of a kind are stored in slot order
		CwNode node = {};
		uint32_t imask = 0, lmask = 0;
		for (int s = 0; s < 8; s++) if (childInSlot[s] >= 0)
		{
			if (linkedRoots || !IsLeaf( bvh2[childInSlot[s]] )) imask |= 1u << s; else lmask |= 1u << s;
		}
		int internalCount = 0;
		for (int s = 0; s < 8; s++) internalCount += (imask >> s) & 1;
		const uint32_t childBase = (uint32_t)out.nodes.size();
		const uint32_t triBase = (uint32_t)(verts4 ? out.tris.size() : out.leafIds.size());
		node.w[4] = childBase + nodeOffset, node.w[5] = triBase + triOffset, node.w[6] = imask | (lmask << 8);
		float clo[8][3] = {}, chi[8][3] = {};
		int nextInternal = 0;
		for (int s = 0; s < 8; s++)
		{
			const int ci = childInSlot[s];
			if (ci < 0) continue;
			const Bvh2Node& c = bvh2[ci];
			memcpy( clo[s], c.lo, 12 ), memcpy( chi[s], c.hi, 12 );
			if (IsLeaf( c ) && linkedRoots)
			{
				// flat scene: the instance's BLAS root is copied in as an internal child
				links.push_back( { (int)childBase + nextInternal, primIdx[~c.left] } );
				nextInternal++;
			}
			else if (IsLeaf( c )) emitLeafPrim( primIdx[~c.left] );	// exactly one primitive (BuildBvh2FromBoxes)
			else
			{
				queue.push_back( { ci, (int)childBase + nextInternal } );
				nextInternal++;
			}
		}
		CwEncodePlanes( node.w, root.lo, root.hi, clo, chi, imask | lmask );
		for (int k = 0; k < internalCount; k++) out.nodes.push_back( CwNode{} );
		out.nodes[task.cwNode] = node;
	}
	for (const Link& l : links) out.nodes[l.cwNode] = linkedRoots[l.prim];
}

} // namespace lh2b

/* Device-free entry: the host builder's output for one mesh (binned-SAH binary tree -> 8-wide collapse -> CWBVH encoding), copied
   out for inspection. tests/test_host_bvh_cpu.py decodes it from the layout documented in bvh.h and checks it on the CPU.
   counts: [0] nodes, [1] triangle records. Returns 1 if the caller's arrays are too small (counts are still filled in). */
#include "../../include/lh2b.h"
extern "C" int lh2b_host_bvh_build( const float* verts4, int triCount, void* nodesOut, int maxNodes, void* trisOut, int maxTris, int* counts )
{
	try
	{
		std::vector<lh2b::Bvh2Node> bvh2;
		std::vector<uint32_t> prim;
		lh2b::BuildBvh2SAH( verts4, triCount, bvh2, prim );
		lh2b::CwBvh cw;
		lh2b::CollapseToCwBvh( bvh2, prim, verts4, cw );
		counts[0] = (int)cw.nodes.size(), counts[1] = (int)cw.tris.size();
		if (counts[0] > maxNodes || counts[1] > maxTris) return 1;
		memcpy( nodesOut, cw.nodes.data(), cw.nodes.size() * sizeof( lh2b::CwNode ) );
		memcpy( trisOut, cw.tris.data(), cw.tris.size() * sizeof( lh2b::CwTri ) );
		return 0;
	}
	catch (...) { return 2; }
}
