/* lh2_oracle_cwbvh.h - TEST INFRASTRUCTURE ONLY. An independent CPU reader of the product's acceleration-structure format: it decodes
   the 8-wide compressed BVH exactly as lighthouse2_b200/csrc/bvh.h documents it (80-byte nodes: biased grid origin, three half
   spacings stored as the upper halves of floats, slot masks imask / lmask, child / triangle base, 6 x 8 plane bytes; one triangle
   per leaf slot; 48-byte triangle records v0 / e1 / e2 + primitive index) and
     Check:        walks the tree and verifies that it is a correct acceleration structure for the mesh - every node and every
                   triangle record reachable exactly once, records equal to the mesh triangles, every decoded child box contains
                   everything below it (the quantisation is conservative), slot masks well-formed;
     ClosestHits:  traverses it with plain float slab tests (boxes padded like lh2_oracle_bvh.h) and the oracle's triangle test and
                   tie rule, so the hits must equal the exhaustive search of lh2_oracle_geom.h bit for bit.
   Used by tests/test_host_bvh_cpu.py on the output of the product's host builder (lh2b_host_bvh_build) - no GPU involved.
   The reference has no BVH code of its own (optixAccelBuild / optixTrace: core_mesh.cpp:105,123, .optix.cu:125,136,148). */
#pragma once
#include "lh2_oracle_geom.h"
#include <vector>
#include <algorithm>

namespace orcw
{

static const int NODE_BYTES = 80;

struct Node
{
	double pb[3], spacing[3]; uint32_t imask, lmask, childBase, triBase; uint8_t qlo[3][8], qhi[3][8];
};

static inline float BitsToFloat( uint32_t u ) { float f; memcpy( &f, &u, 4 ); return f; }

static inline Node Decode( const uint8_t* b )
{
	Node n;
	uint32_t w[20];
	memcpy( w, b, 80 );
	for (int a = 0; a < 3; a++) n.pb[a] = BitsToFloat( w[a] );
	// half spacings: upper 16 bits of a float each - x in the high half of w[3], y in its low half, z in the high half of w[6]
	n.spacing[0] = 2.0 * BitsToFloat( w[3] & 0xffff0000u ), n.spacing[1] = 2.0 * BitsToFloat( w[3] << 16 ), n.spacing[2] = 2.0 * BitsToFloat( w[6] & 0xffff0000u );
	n.imask = w[6] & 255u, n.lmask = (w[6] >> 8) & 255u, n.childBase = w[4], n.triBase = w[5];
	memcpy( n.qlo, b + 32, 24 ), memcpy( n.qhi, b + 56, 24 );
	return n;
}

/* plane( q ) = pb + (32768 + q) * spacing, in double: exact for the float inputs */
static inline void ChildBox( const Node& n, int s, float* lo, float* hi )
{
	for (int a = 0; a < 3; a++)
	{
		const double l = n.pb[a] + (32768.0 + n.qlo[a][s]) * n.spacing[a], h = n.pb[a] + (32768.0 + n.qhi[a][s]) * n.spacing[a];
		lo[a] = (float)l, hi[a] = (float)h;
		if ((double)lo[a] > l) lo[a] = nextafterf( lo[a], -3e38f );	// round outwards: the check below must not pass by rounding
		if ((double)hi[a] < h) hi[a] = nextafterf( hi[a], 3e38f );
	}
}

static inline int Rank( uint32_t mask, int s ) { return __builtin_popcount( mask & ((1u << s) - 1u) ); }

struct Report { int nodesVisited, trisVisited, maxDepth, emptySlots, leafSlots, innerSlots, errors, firstError; };

/* returns the exact bounds of everything below node 'idx' through lo / hi */
static inline void CheckNode( const uint8_t* nodes, int nNodes, const float* tris, int nTris, const float* verts4, int triCount,
	std::vector<uint8_t>& nodeSeen, std::vector<int>& primSeen, std::vector<uint8_t>& recSeen, int idx, int depth, float* lo, float* hi, Report& r )
{
	auto fail = [&]( int code ) { if (r.errors++ == 0) r.firstError = code; };
	for (int a = 0; a < 3; a++) lo[a] = 3e38f, hi[a] = -3e38f;
	if (idx < 0 || idx >= nNodes) { fail( 1 ); return; }
	if (nodeSeen[idx]++) { fail( 2 ); return; }
	r.nodesVisited++, r.maxDepth = std::max( r.maxDepth, depth );
	const Node n = Decode( nodes + (size_t)idx * NODE_BYTES );
	if (n.imask & n.lmask) fail( 3 );
	for (int s = 0; s < 8; s++)
	{
		const bool isInner = (n.imask >> s) & 1, isLeaf = (n.lmask >> s) & 1;
		float clo[3], chi[3], blo[3], bhi[3];
		ChildBox( n, s, blo, bhi );
		if (!isInner && !isLeaf)
		{
			r.emptySlots++;
			continue;
		}
		if (isInner)
		{
			r.innerSlots++;
			CheckNode( nodes, nNodes, tris, nTris, verts4, triCount, nodeSeen, primSeen, recSeen, (int)n.childBase + Rank( n.imask, s ), depth + 1, clo, chi, r );
		}
		else
		{
			r.leafSlots++;
			for (int a = 0; a < 3; a++) clo[a] = 3e38f, chi[a] = -3e38f;
			const int rec = (int)n.triBase + Rank( n.lmask, s );	// exactly one triangle per leaf slot
			if (rec < 0 || rec >= nTris) { fail( 6 ); continue; }
			if (recSeen[rec]++) fail( 7 );
			r.trisVisited++;
			const float* t = tris + (size_t)rec * 12;
			int prim; memcpy( &prim, t + 3, 4 );
			if (prim < 0 || prim >= triCount) { fail( 8 ); continue; }
			primSeen[prim]++;
			const float* v = verts4 + (size_t)prim * 12;
			for (int a = 0; a < 3; a++)
			{
				// the record is the Moeller-Trumbore form of exactly this triangle
				if (t[a] != v[a] || t[4 + a] != v[4 + a] - v[a] || t[8 + a] != v[8 + a] - v[a]) fail( 9 );
				clo[a] = std::min( clo[a], std::min( v[a], std::min( v[4 + a], v[8 + a] ) ) );
				chi[a] = std::max( chi[a], std::max( v[a], std::max( v[4 + a], v[8 + a] ) ) );
			}
		}
		// conservative encoding: the decoded box of the slot contains everything below it
		for (int a = 0; a < 3; a++) if (clo[a] <= chi[a] && (blo[a] > clo[a] || bhi[a] < chi[a])) fail( 10 );
		for (int a = 0; a < 3; a++) lo[a] = std::min( lo[a], clo[a] ), hi[a] = std::max( hi[a], chi[a] );
	}
}

static inline Report Check( const uint8_t* nodes, int nNodes, const float* tris, int nTris, const float* verts4, int triCount )
{
	Report r = {};
	std::vector<uint8_t> nodeSeen( nNodes, 0 ), recSeen( nTris, 0 );
	std::vector<int> primSeen( triCount, 0 );
	float lo[3], hi[3];
	CheckNode( nodes, nNodes, tris, nTris, verts4, triCount, nodeSeen, primSeen, recSeen, 0, 1, lo, hi, r );
	for (int i = 0; i < triCount; i++) if (primSeen[i] != 1) { if (r.errors++ == 0) r.firstError = 11; }
	if (r.nodesVisited != nNodes) { if (r.errors++ == 0) r.firstError = 12; }
	return r;
}

/* the oracle's triangle test on a record (v0, e1, e2): same operations in the same order as orc::TriTest */
static inline bool RecordTest( const float* O, const float* D, const float* t, float& tt, float& u, float& v )
{
	const float e1x = t[4], e1y = t[5], e1z = t[6], e2x = t[8], e2y = t[9], e2z = t[10];
	const float pvx = orc::CrossX( D[0], D[1], D[2], e2x, e2y, e2z ), pvy = orc::CrossY( D[0], D[1], D[2], e2x, e2y, e2z ), pvz = orc::CrossZ( D[0], D[1], D[2], e2x, e2y, e2z );
	const float det = orc::Dot3( e1x, e1y, e1z, pvx, pvy, pvz );
	if (!(det != 0.0f)) return false;
	const float inv = 1.0f / det;
	const float tvx = O[0] - t[0], tvy = O[1] - t[1], tvz = O[2] - t[2];
	u = orc::Dot3( tvx, tvy, tvz, pvx, pvy, pvz ) * inv;
	if (!(u >= 0.0f && u <= 1.0f)) return false;
	const float qvx = orc::CrossX( tvx, tvy, tvz, e1x, e1y, e1z ), qvy = orc::CrossY( tvx, tvy, tvz, e1x, e1y, e1z ), qvz = orc::CrossZ( tvx, tvy, tvz, e1x, e1y, e1z );
	v = orc::Dot3( D[0], D[1], D[2], qvx, qvy, qvz ) * inv;
	if (!(v >= 0.0f && u + v <= 1.0f)) return false;
	tt = orc::Dot3( e2x, e2y, e2z, qvx, qvy, qvz ) * inv;
	return true;
}

static inline bool SlabHit( const float* lo, const float* hi, const float* O, const float* invD, float tmin, float tmax )
{
	float mag = std::max( fabsf( O[0] ), std::max( fabsf( O[1] ), fabsf( O[2] ) ) );
	for (int a = 0; a < 3; a++) mag = std::max( mag, std::max( fabsf( lo[a] ), fabsf( hi[a] ) ) );
	const float pad = mag * (1.0f / 65536.0f);
	float tn = tmin, tf = tmax;
	for (int a = 0; a < 3; a++)
	{
		const float t0 = (lo[a] - pad - O[a]) * invD[a], t1 = (hi[a] + pad - O[a]) * invD[a];
		const float l = t0 < t1 ? t0 : t1, h = t0 < t1 ? t1 : t0;
		if (l > tn) tn = l;
		if (h < tf) tf = h;
	}
	return !(tn > tf * 1.00001f) || !(tf == tf);
}

static inline bool ClosestHit( const uint8_t* nodes, const float* tris, const float* O, const float* D, orc::Hit& best )
{
	best.t = 1e34f, best.inst = -1, best.prim = -1, best.u = best.v = 0;
	const float invD[3] = { 1.0f / D[0], 1.0f / D[1], 1.0f / D[2] };
	int stack[256], sp = 0;
	stack[sp++] = 0;
	while (sp > 0)
	{
		const Node n = Decode( nodes + (size_t)stack[--sp] * NODE_BYTES );
		for (int s = 0; s < 8; s++)
		{
			const bool isInner = (n.imask >> s) & 1, isLeaf = (n.lmask >> s) & 1;
			if (!isInner && !isLeaf) continue;
			float lo[3], hi[3];
			ChildBox( n, s, lo, hi );
			if (!SlabHit( lo, hi, O, invD, 0.0f, best.t )) continue;
			if (isInner) { if (sp < 255) stack[sp++] = (int)n.childBase + Rank( n.imask, s ); continue; }
			const float* t = tris + (size_t)(n.triBase + Rank( n.lmask, s )) * 12;
			float tt, u, v;
			if (!RecordTest( O, D, t, tt, u, v ) || !(tt > 0.0f)) continue;
			int prim; memcpy( &prim, t + 3, 4 );
			if (tt < best.t || (tt == best.t && best.prim >= 0 && prim < best.prim)) best.t = tt, best.u = u, best.v = v, best.inst = 0, best.prim = prim;
		}
	}
	return best.prim >= 0;
}

} // namespace orcw
