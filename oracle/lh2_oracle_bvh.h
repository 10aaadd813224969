/* lh2_oracle_bvh.h - TEST INFRASTRUCTURE ONLY. Optional acceleration structure of the CPU oracle.

   The oracle's definition of a ray query is the exhaustive search of lh2_oracle_geom.h (ClosestHit /
   Occluded over instances x triangles). This file only *prunes* that search: a plain binary BVH per mesh
   (binned SAH, 16 bins, leaves of <= 4 triangles, own code - nothing shared with the product's builders in
   lighthouse2_b200/csrc) whose boxes are padded per ray, so that every triangle the exhaustive search could
   accept is still tested with the same TriTest, the same float operation order and the same tie rule
   (equal t -> smaller (instance, primitive)). tests/test_oracle_cpu.py checks BVH == exhaustive search
   bit for bit; with that established the BVH oracle is what makes parity checks at BASELINE.json's full
   sizes (2 M rays x 1 M triangles) and an honest CPU baseline (a CPU tracer *with* a BVH, as any CPU
   implementation of optixTrace would have; SURVEY.md 8d "if the simple-BVH variant is used say so") possible.

   Reference call sites restated: as lh2_oracle_geom.h (.optix.cu:125,136,148,174-184; rendercore.cpp:405-416).
*/
#pragma once
#include "lh2_oracle_geom.h"
#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace orc
{

struct BvhNode { float lo[3]; uint32_t leftFirst; float hi[3]; uint32_t count; };	// count > 0: leaf over prims[leftFirst .. +count)

struct MeshBvh
{
	std::vector<BvhNode> nodes;
	std::vector<uint32_t> prims;
};

namespace bvhdetail
{
struct Box
{
	float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
	void Grow( const float* p ) { for (int a = 0; a < 3; a++) lo[a] = std::min( lo[a], p[a] ), hi[a] = std::max( hi[a], p[a] ); }
	void Grow( const Box& b ) { for (int a = 0; a < 3; a++) lo[a] = std::min( lo[a], b.lo[a] ), hi[a] = std::max( hi[a], b.hi[a] ); }
	float Area() const
	{
		const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
		return x < 0 ? 0.0f : x * y + y * z + z * x;
	}
};
} // namespace bvhdetail

static inline void BuildMeshBvh( const Mesh& m, MeshBvh& out )
{
	using bvhdetail::Box;
	const int n = m.triCount;
	out.prims.resize( n );
	out.nodes.clear();
	out.nodes.reserve( (size_t)n + 2 );
	std::vector<Box> tb( n );
	std::vector<float> cen( (size_t)n * 3 );
	for (int i = 0; i < n; i++)
	{
		out.prims[i] = i;
		const float* v = m.verts4 + (size_t)i * 12;
		tb[i].Grow( v ), tb[i].Grow( v + 4 ), tb[i].Grow( v + 8 );
		for (int a = 0; a < 3; a++) cen[(size_t)i * 3 + a] = 0.5f * (tb[i].lo[a] + tb[i].hi[a]);
	}
	struct Task { uint32_t node, first, count; };
	std::vector<Task> todo;
	out.nodes.push_back( BvhNode{} );
	if (n == 0) { BvhNode& r = out.nodes[0]; memset( &r, 0, sizeof( r ) ); r.count = 0; r.leftFirst = 0; r.lo[0] = r.lo[1] = r.lo[2] = 1; r.hi[0] = r.hi[1] = r.hi[2] = -1; return; }
	todo.push_back( { 0, 0, (uint32_t)n } );
	const int BINS = 16;
	while (!todo.empty())
	{
		const Task t = todo.back();
		todo.pop_back();
		Box nb, cb;
		for (uint32_t i = t.first; i < t.first + t.count; i++)
		{
			const uint32_t p = out.prims[i];
			nb.Grow( tb[p] ), cb.Grow( &cen[(size_t)p * 3] );
		}
		BvhNode& node = out.nodes[t.node];
		for (int a = 0; a < 3; a++) node.lo[a] = nb.lo[a], node.hi[a] = nb.hi[a];
		node.leftFirst = t.first, node.count = t.count;
		if (t.count <= 4) continue;
		// binned SAH over the centroid box
		int bestAxis = -1, bestSplit = 0;
		float bestCost = 3e38f;
		for (int a = 0; a < 3; a++)
		{
			const float ext = cb.hi[a] - cb.lo[a];
			if (!(ext > 0)) continue;
			Box bb[BINS];
			uint32_t bc[BINS] = {};
			const float scale = BINS / ext;
			for (uint32_t i = t.first; i < t.first + t.count; i++)
			{
				const uint32_t p = out.prims[i];
				const int b = std::min( BINS - 1, (int)((cen[(size_t)p * 3 + a] - cb.lo[a]) * scale) );
				bb[b].Grow( tb[p] ), bc[b]++;
			}
			float rightArea[BINS];
			uint32_t rightCount[BINS];
			Box acc;
			uint32_t cnt = 0;
			for (int b = BINS - 1; b > 0; b--) acc.Grow( bb[b] ), cnt += bc[b], rightArea[b] = acc.Area(), rightCount[b] = cnt;
			acc = Box(), cnt = 0;
			for (int b = 0; b < BINS - 1; b++)
			{
				acc.Grow( bb[b] ), cnt += bc[b];
				if (cnt == 0 || rightCount[b + 1] == 0) continue;
				const float cost = acc.Area() * cnt + rightArea[b + 1] * rightCount[b + 1];
				if (cost < bestCost) bestCost = cost, bestAxis = a, bestSplit = b + 1;
			}
		}
		uint32_t mid;
		if (bestAxis < 0)
		{
			// all centroids coincide: split down the middle
			mid = t.first + t.count / 2;
		}
		else
		{
			const float ext = cb.hi[bestAxis] - cb.lo[bestAxis], scale = BINS / ext, lo = cb.lo[bestAxis];
			uint32_t* b0 = out.prims.data() + t.first;
			uint32_t* pm = std::partition( b0, b0 + t.count, [&]( uint32_t p ) {
				return std::min( BINS - 1, (int)((cen[(size_t)p * 3 + bestAxis] - lo) * scale) ) < bestSplit; } );
			mid = (uint32_t)(pm - out.prims.data());
			if (mid == t.first || mid == t.first + t.count) mid = t.first + t.count / 2;
		}
		const uint32_t left = (uint32_t)out.nodes.size();
		out.nodes.push_back( BvhNode{} ), out.nodes.push_back( BvhNode{} );
		BvhNode& nd = out.nodes[t.node];	// (the vector may have moved)
		nd.leftFirst = left, nd.count = 0;
		todo.push_back( { left, t.first, mid - t.first } );
		todo.push_back( { left + 1, mid, t.first + t.count - mid } );
	}
}

/* Padded slab test. The exhaustive search accepts a triangle from float arithmetic whose barycentric error, seen as a
   distance, is a few ulp of the coordinate magnitudes involved; the pad (2^-16 of the largest magnitude among origin and
   box corners, about 250 ulp) is far above that, and the comparison itself is written so that a NaN (0 * inf) visits. */
static inline bool RayBox( const BvhNode& n, const float* O, const float* invD, const float oMag, const float tmin, const float tmax, float& tnear )
{
	float mag = oMag;
	for (int a = 0; a < 3; a++) mag = std::max( mag, std::max( fabsf( n.lo[a] ), fabsf( n.hi[a] ) ) );
	const float pad = mag * (1.0f / 65536.0f);
	float tn = tmin, tf = tmax;
	for (int a = 0; a < 3; a++)
	{
		const float t0 = (n.lo[a] - pad - O[a]) * invD[a], t1 = (n.hi[a] + pad - O[a]) * invD[a];
		const float lo = t0 < t1 ? t0 : t1, hi = t0 < t1 ? t1 : t0;	// NaN: both comparisons false -> lo = hi = t0 = NaN
		if (lo > tn) tn = lo;	// NaN never tightens the interval
		if (hi < tf) tf = hi;
	}
	tnear = tn;
	return !(tn > tf * 1.00001f + 0.0f) || !(tf == tf);
}

struct Accel
{
	std::vector<std::shared_ptr<MeshBvh>> meshBvh;	// per mesh
	std::vector<float> instBox;							// 6 floats per instance: world-space box of the mesh root (padded at use)
};

namespace bvhdetail
{
static inline uint64_t HashBytes( const void* p, size_t bytes )
{
	// 64-bit FNV-1a over 8-byte words (tail bytes folded in one by one)
	const uint64_t* w = (const uint64_t*)p;
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < bytes / 8; i++) h = (h ^ w[i]) * 1099511628211ull;
	const uint8_t* b = (const uint8_t*)p + (bytes / 8) * 8;
	for (size_t i = 0; i < bytes % 8; i++) h = (h ^ b[i]) * 1099511628211ull;
	return h;
}
static std::mutex cacheLock;
static std::map<std::pair<uint64_t, int>, std::shared_ptr<MeshBvh>> cache;	// (hash of vertex bytes, triCount) -> BVH
} // namespace bvhdetail

/* BVHs are cached by vertex content: bench.py and the tests render many crops / frames of one scene. */
static inline std::shared_ptr<MeshBvh> CachedMeshBvh( const Mesh& m )
{
	const std::pair<uint64_t, int> key( bvhdetail::HashBytes( m.verts4, (size_t)m.triCount * 48 ), m.triCount );
	{
		std::lock_guard<std::mutex> g( bvhdetail::cacheLock );
		auto it = bvhdetail::cache.find( key );
		if (it != bvhdetail::cache.end()) return it->second;
	}
	std::shared_ptr<MeshBvh> b = std::make_shared<MeshBvh>();
	BuildMeshBvh( m, *b );
	std::lock_guard<std::mutex> g( bvhdetail::cacheLock );
	if (bvhdetail::cache.size() >= 16) bvhdetail::cache.clear();
	bvhdetail::cache[key] = b;
	return b;
}

static inline Accel* BuildAccel( const Mesh* meshes, int meshCount, const Instance* instances, int instanceCount )
{
	Accel* a = new Accel;
	a->meshBvh.resize( meshCount );
	for (int i = 0; i < meshCount; i++) a->meshBvh[i] = CachedMeshBvh( meshes[i] );
	a->instBox.resize( (size_t)instanceCount * 6 );
	for (int i = 0; i < instanceCount; i++)
	{
		const BvhNode& r = a->meshBvh[instances[i].mesh]->nodes[0];
		bvhdetail::Box wb;
		const float* x = instances[i].xform;
		for (int c = 0; c < 8; c++)
		{
			const float p[3] = { c & 1 ? r.hi[0] : r.lo[0], c & 2 ? r.hi[1] : r.lo[1], c & 4 ? r.hi[2] : r.lo[2] };
			float w[3];
			for (int k = 0; k < 3; k++) w[k] = x[k * 4] * p[0] + x[k * 4 + 1] * p[1] + x[k * 4 + 2] * p[2] + x[k * 4 + 3];
			wb.Grow( w );
		}
		for (int k = 0; k < 3; k++) a->instBox[(size_t)i * 6 + k] = wb.lo[k], a->instBox[(size_t)i * 6 + 3 + k] = wb.hi[k];
	}
	return a;
}

static inline void SafeInvDir( const float* D, float* invD )
{
	for (int a = 0; a < 3; a++) invD[a] = 1.0f / D[a];	// +-inf for 0: the slab test handles it (and NaN) conservatively
}

/* Closest hit inside one mesh (object-space ray); updates best with the exhaustive search's acceptance and tie rule. */
static inline void MeshClosest( const Mesh& m, const MeshBvh& b, const int inst, const float* O, const float* D, const float tmin, Hit& best )
{
	float invD[3];
	SafeInvDir( D, invD );
	const float oMag = std::max( fabsf( O[0] ), std::max( fabsf( O[1] ), fabsf( O[2] ) ) );
	uint32_t stack[128];
	int sp = 0;
	uint32_t cur = 0;
	float tn;
	if (!RayBox( b.nodes[0], O, invD, oMag, tmin, best.t, tn )) return;
	while (true)
	{
		const BvhNode& n = b.nodes[cur];
		if (n.count > 0)
		{
			for (uint32_t i = 0; i < n.count; i++)
			{
				const int p = (int)b.prims[n.leftFirst + i];
				const float* v = m.verts4 + (size_t)p * 12;
				float t, u, w;
				if (!TriTest( O, D, v, v + 4, v + 8, t, u, w )) continue;
				if (!(t > tmin)) continue;
				if (t < best.t || (t == best.t && best.prim >= 0 && (inst < best.inst || (inst == best.inst && p < best.prim))))
					best.t = t, best.u = u, best.v = w, best.inst = inst, best.prim = p;
			}
		}
		else
		{
			float t0, t1;
			const bool h0 = RayBox( b.nodes[n.leftFirst], O, invD, oMag, tmin, best.t, t0 );
			const bool h1 = RayBox( b.nodes[n.leftFirst + 1], O, invD, oMag, tmin, best.t, t1 );
			if (h0 && h1)
			{
				const bool firstLeft = !(t1 < t0);
				if (sp < 127) stack[sp++] = n.leftFirst + (firstLeft ? 1 : 0);
				cur = n.leftFirst + (firstLeft ? 0 : 1);
				continue;
			}
			if (h0 || h1) { cur = n.leftFirst + (h0 ? 0 : 1); continue; }
		}
		// pop; re-test against the (possibly shrunk) best.t. '<=': equal-t candidates must still be reached
		bool found = false;
		while (sp > 0)
		{
			cur = stack[--sp];
			if (RayBox( b.nodes[cur], O, invD, oMag, tmin, best.t, tn )) { found = true; break; }
		}
		if (!found) return;
	}
}

static inline bool MeshOccluded( const Mesh& m, const MeshBvh& b, const float* O, const float* D, const float tmin, const float tmax )
{
	float invD[3];
	SafeInvDir( D, invD );
	const float oMag = std::max( fabsf( O[0] ), std::max( fabsf( O[1] ), fabsf( O[2] ) ) );
	uint32_t stack[128];
	int sp = 0;
	stack[sp++] = 0;
	float tn;
	while (sp > 0)
	{
		const BvhNode& n = b.nodes[stack[--sp]];
		if (!RayBox( n, O, invD, oMag, tmin, tmax, tn )) continue;
		if (n.count > 0)
		{
			for (uint32_t i = 0; i < n.count; i++)
			{
				const float* v = m.verts4 + (size_t)b.prims[n.leftFirst + i] * 12;
				float t, u, w;
				if (TriTest( O, D, v, v + 4, v + 8, t, u, w ) && t > tmin && t < tmax) return true;
			}
		}
		else if (sp < 126) stack[sp++] = n.leftFirst, stack[sp++] = n.leftFirst + 1;
	}
	return false;
}

static inline bool InstanceBoxHit( const Accel& a, const int i, const float* O, const float* D, const float tmin, const float tmax )
{
	BvhNode n;
	for (int k = 0; k < 3; k++) n.lo[k] = a.instBox[(size_t)i * 6 + k], n.hi[k] = a.instBox[(size_t)i * 6 + 3 + k];
	// world boxes come from transformed corners (rounded): widen by 2^-12 of their magnitude on top of the per-ray pad
	for (int k = 0; k < 3; k++)
	{
		const float w = std::max( fabsf( n.lo[k] ), fabsf( n.hi[k] ) ) * (1.0f / 4096.0f);
		n.lo[k] -= w, n.hi[k] += w;
	}
	float invD[3], tn;
	SafeInvDir( D, invD );
	const float oMag = std::max( fabsf( O[0] ), std::max( fabsf( O[1] ), fabsf( O[2] ) ) );
	return RayBox( n, O, invD, oMag, tmin, tmax, tn );
}

static inline Accel* BuildSceneAccel( const Scene& s ) { return BuildAccel( s.meshes, s.meshCount, s.instances, s.instanceCount ); }
static inline void FreeSceneAccel( Accel* a ) { delete a; }

static inline bool AccelClosestHit( const Scene& s, const float* O, const float* D, const float tmin, float tmax, Hit& best )
{
	best.t = tmax, best.inst = -1, best.prim = -1, best.u = best.v = 0;
	for (int i = 0; i < s.instanceCount; i++)
	{
		if (s.instanceCount > 1 && !InstanceBoxHit( *s.accel, i, O, D, tmin, best.t )) continue;
		const int mi = s.instances[i].mesh;
		float oO[3], oD[3];
		ToObjectSpace( s.inverses + i * 12, O, D, oO, oD );
		MeshClosest( s.meshes[mi], *s.accel->meshBvh[mi], i, oO, oD, tmin, best );
	}
	return best.prim >= 0;
}

static inline bool AccelOccluded( const Scene& s, const float* O, const float* D, const float tmin, const float tmax )
{
	for (int i = 0; i < s.instanceCount; i++)
	{
		if (s.instanceCount > 1 && !InstanceBoxHit( *s.accel, i, O, D, tmin, tmax )) continue;
		const int mi = s.instances[i].mesh;
		float oO[3], oD[3];
		ToObjectSpace( s.inverses + i * 12, O, D, oO, oD );
		if (MeshOccluded( s.meshes[mi], *s.accel->meshBvh[mi], oO, oD, tmin, tmax )) return true;
	}
	return false;
}

} // namespace orc
