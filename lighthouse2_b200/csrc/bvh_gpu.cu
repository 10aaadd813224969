/* bvh_gpu.cu - GPU construction of the acceleration structures: Morton-code LBVH (Karras 2012) -> bottom-up bounds ->
   greedy collapse to 8-wide -> node encoding (bvh.h), all on the core's stream without host round trips. Also the "refit"
   path (same topology, new bounds, re-collapse) and the per-frame top level over instance boxes.

   Replaces the closed OptiX builders: optixAccelBuild on triangles (lib/rendercore_optix7/core_mesh.cpp:67-129, always
   a full rebuild there) and on instances (lib/rendercore_optix7/rendercore.cpp:767-797, every Render).

   Stages (n primitives; triangles for a BLAS, instances for the TLAS):
     1. boundsKernel ......... primitive boxes + scene box (block reduce + float atomics)             24n B read
     2. mortonKernel ......... 63-bit Morton code of the box centre                                     8n B written
     3. cub::DeviceRadixSort . (code, primitive index) pairs - the one library call in the builder
     4. radixTreeKernel ...... Karras' binary radix tree over the sorted codes (ties broken by index)
     5. fitKernel ............ leaf boxes, then parents bottom-up (second arrival continues)            = refit
     6. collapseKernel ....... persistent cooperative kernel, one BFS level per grid.sync(): each task turns one binary
                               subtree root into one 80-byte wide node (<= 8 children, octant slots, 8-bit child planes),
                               emits leaf triangles in Moeller-Trumbore form and queues the internal children.
   Encoding rules are identical to the host builder (bvh_build_cpu.cpp); both are checked against the brute-force oracle.
*/
#include "core.h"
#include "kernels.h"
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cfloat>

namespace cg = cooperative_groups;

namespace lh2b
{

struct BuildTask { int bvh2Node; uint32_t cwNode; };

struct GpuBuildScratch
{
	DevBuf<float4> primLo, primHi;			// primitive boxes
	DevBuf<uint64_t> keys, keysAlt;
	DevBuf<uint32_t> idx, idxAlt;
	DevBuf<uint8_t> cubTemp;
	DevBuf<float4> nodeLo, nodeHi;			// 2n-1 binary nodes: internal [0, n-1), leaves [n-1, 2n-1)
	DevBuf<int2> children;					// per internal node
	DevBuf<uint32_t> subtree;				// per internal node: number of primitives below it
	DevBuf<int> clusterA, clusterB, mergeTmp;	// PLOC working sets
	DevBuf<uint32_t> blockCount;
	DevBuf<int> parent;						// per binary node
	DevBuf<uint32_t> visit;					// per internal node arrival counter
	DevBuf<BuildTask> queue;
	DevBuf<float> dpCost;					// per internal node: T[1..7] of the optimal collapse (see fitKernel)
	DevBuf<uint8_t> dpSplit;				// per internal node: the choice behind T[1..7] and behind the node's own 8 slots
	DevBuf<uint32_t> ctrl;					// [0] node counter, [1] leaf counter, [2] queue tail, [3..8] scene box as ordered ints, [9] error flag
};

/* ---- float atomics through order-preserving integer keys ---- */
__device__ __forceinline__ int FloatOrdered( const float f ) { const int i = __float_as_int( f ); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float OrderedFloat( const int i ) { return __int_as_float( i >= 0 ? i : i ^ 0x7fffffff ); }

__global__ void initCtrlKernel( uint32_t* ctrl )
{
	ctrl[0] = 1, ctrl[1] = 0, ctrl[2] = 1, ctrl[9] = 0;
	int* box = (int*)(ctrl + 3);
	box[0] = box[1] = box[2] = FloatOrdered( FLT_MAX ), box[3] = box[4] = box[5] = FloatOrdered( -FLT_MAX );
}

/* stage 1 for triangles */
__global__ void triBoundsKernel( const float4* __restrict__ verts, const int n, float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* ctrl )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	float3 a = make_float3( FLT_MAX, FLT_MAX, FLT_MAX ), b = make_float3( -FLT_MAX, -FLT_MAX, -FLT_MAX );
	if (i < n)
	{
		const float4 v0 = verts[i * 3], v1 = verts[i * 3 + 1], v2 = verts[i * 3 + 2];
		a = make_float3( fminf( v0.x, fminf( v1.x, v2.x ) ), fminf( v0.y, fminf( v1.y, v2.y ) ), fminf( v0.z, fminf( v1.z, v2.z ) ) );
		b = make_float3( fmaxf( v0.x, fmaxf( v1.x, v2.x ) ), fmaxf( v0.y, fmaxf( v1.y, v2.y ) ), fmaxf( v0.z, fmaxf( v1.z, v2.z ) ) );
		lo[i] = make_float4( a.x, a.y, a.z, 0 ), hi[i] = make_float4( b.x, b.y, b.z, 0 );
	}
	// warp reduce, then one atomic set per block (six atomics per warp serialised on one L2 line: 127 us for 1 M triangles)
	for (int o = 16; o > 0; o >>= 1)
	{
		a.x = fminf( a.x, __shfl_xor_sync( 0xffffffffu, a.x, o ) ), a.y = fminf( a.y, __shfl_xor_sync( 0xffffffffu, a.y, o ) ), a.z = fminf( a.z, __shfl_xor_sync( 0xffffffffu, a.z, o ) );
		b.x = fmaxf( b.x, __shfl_xor_sync( 0xffffffffu, b.x, o ) ), b.y = fmaxf( b.y, __shfl_xor_sync( 0xffffffffu, b.y, o ) ), b.z = fmaxf( b.z, __shfl_xor_sync( 0xffffffffu, b.z, o ) );
	}
	__shared__ float red[8][6];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0) red[warp][0] = a.x, red[warp][1] = a.y, red[warp][2] = a.z, red[warp][3] = b.x, red[warp][4] = b.y, red[warp][5] = b.z;
	__syncthreads();
	if (threadIdx.x < 6)
	{
		float v = red[0][threadIdx.x];
		const int warps = (blockDim.x + 31) >> 5;
		for (int w = 1; w < warps; w++) v = threadIdx.x < 3 ? fminf( v, red[w][threadIdx.x] ) : fmaxf( v, red[w][threadIdx.x] );
		int* box = (int*)(ctrl + 3);
		if (threadIdx.x < 3) atomicMin( box + threadIdx.x, FloatOrdered( v ) ); else atomicMax( box + threadIdx.x, FloatOrdered( v ) );
	}
}

/* stage 1 for instances: world box of the transformed mesh box (8 corners), padded like the host path */
struct InstBuildIn { float xform[12]; const float4* bounds; uint64_t pad; };	// bounds: the mesh's device-resident {lo, hi}
__global__ void instBoundsKernel( const InstBuildIn* __restrict__ inst, const int n, float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* ctrl )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const InstBuildIn in = inst[i];
	const float4 mlo = in.bounds[0], mhi = in.bounds[1];
	float a[3] = { 1e34f, 1e34f, 1e34f }, b[3] = { -1e34f, -1e34f, -1e34f };
	for (int k = 0; k < 8; k++)
	{
		const float x = (k & 1) ? mhi.x : mlo.x, y = (k & 2) ? mhi.y : mlo.y, z = (k & 4) ? mhi.z : mlo.z;
		for (int c = 0; c < 3; c++)
		{
			const float v = in.xform[c * 4] * x + in.xform[c * 4 + 1] * y + in.xform[c * 4 + 2] * z + in.xform[c * 4 + 3];
			a[c] = fminf( a[c], v ), b[c] = fmaxf( b[c], v );
		}
	}
	for (int c = 0; c < 3; c++)
	{
		const float pad = 1e-5f * fmaxf( fabsf( a[c] ), fabsf( b[c] ) ) + 1e-30f;
		a[c] -= pad, b[c] += pad;
	}
	lo[i] = make_float4( a[0], a[1], a[2], 0 ), hi[i] = make_float4( b[0], b[1], b[2], 0 );
	int* box = (int*)(ctrl + 3);
	atomicMin( box + 0, FloatOrdered( a[0] ) ), atomicMin( box + 1, FloatOrdered( a[1] ) ), atomicMin( box + 2, FloatOrdered( a[2] ) );
	atomicMax( box + 3, FloatOrdered( b[0] ) ), atomicMax( box + 4, FloatOrdered( b[1] ) ), atomicMax( box + 5, FloatOrdered( b[2] ) );
}

/* stage 2 */
__device__ __forceinline__ uint64_t Spread21( uint64_t x )
{
	x &= 0x1fffffull;
	x = (x | x << 32) & 0x1f00000000ffffull;
	x = (x | x << 16) & 0x1f0000ff0000ffull;
	x = (x | x << 8) & 0x100f00f00f00f00full;
	x = (x | x << 4) & 0x10c30c30c30c30c3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}
__global__ void mortonKernel( const float4* __restrict__ lo, const float4* __restrict__ hi, const int n, const uint32_t* __restrict__ ctrl,
	uint64_t* __restrict__ keys, uint32_t* __restrict__ idx )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int* box = (const int*)(ctrl + 3);
	const float3 smin = make_float3( OrderedFloat( box[0] ), OrderedFloat( box[1] ), OrderedFloat( box[2] ) );
	const float3 smax = make_float3( OrderedFloat( box[3] ), OrderedFloat( box[4] ), OrderedFloat( box[5] ) );
	const float4 a = lo[i], b = hi[i];
	// one scale for all axes (the largest extent): code bits then measure real distance, which matters for flat scenes
	const float ext = fmaxf( smax.x - smin.x, fmaxf( smax.y - smin.y, smax.z - smin.z ) );
	const float sx = ext > 0 ? 2097151.0f / ext : 0, sy = sx, sz = sx;
	const uint64_t x = (uint64_t)fminf( fmaxf( (0.5f * (a.x + b.x) - smin.x) * sx, 0.0f ), 2097151.0f );
	const uint64_t y = (uint64_t)fminf( fmaxf( (0.5f * (a.y + b.y) - smin.y) * sy, 0.0f ), 2097151.0f );
	const uint64_t z = (uint64_t)fminf( fmaxf( (0.5f * (a.z + b.z) - smin.z) * sz, 0.0f ), 2097151.0f );
	keys[i] = (Spread21( x ) << 2) | (Spread21( y ) << 1) | Spread21( z );
	idx[i] = i;
}

/* stage 4 */
__device__ __forceinline__ int Delta( const uint64_t* __restrict__ keys, const int n, const int i, const int j )
{
	if (j < 0 || j >= n) return -1;
	const uint64_t a = keys[i], b = keys[j];
	if (a == b) return 64 + __clz( i ^ j );
	return __clzll( a ^ b );
}
__global__ void radixTreeKernel( const uint64_t* __restrict__ keys, const int n, int2* __restrict__ children, uint32_t* __restrict__ subtree,
	int* __restrict__ parent, uint32_t* __restrict__ visit )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	const int d = Delta( keys, n, i, i + 1 ) - Delta( keys, n, i, i - 1 ) >= 0 ? 1 : -1;
	const int deltaMin = Delta( keys, n, i, i - d );
	int lmax = 2;
	while (Delta( keys, n, i, i + lmax * d ) > deltaMin) lmax *= 2;
	int l = 0;
	for (int t = lmax / 2; t >= 1; t /= 2) if (Delta( keys, n, i, i + (l + t) * d ) > deltaMin) l += t;
	const int j = i + l * d;
	const int deltaNode = Delta( keys, n, i, j );
	int s = 0, t = l;
	do
	{
		t = (t + 1) >> 1;
		if (Delta( keys, n, i, i + (s + t) * d ) > deltaNode) s += t;
	} while (t > 1);
	const int gamma = i + s * d + min( d, 0 );
	const int first = min( i, j ), last = max( i, j );
	const int left = first == gamma ? (n - 1) + gamma : gamma;
	const int right = last == gamma + 1 ? (n - 1) + gamma + 1 : gamma + 1;
	children[i] = make_int2( left, right );
	subtree[i] = (uint32_t)(last - first + 1);
	parent[left] = i, parent[right] = i;
	visit[i] = 0;
	if (i == 0) parent[0] = -1;
}

/* stage 5: leaf boxes from the (possibly updated) primitive boxes, parents bottom-up.

   With dpCost / dpSplit the same bottom-up sweep also fills the tables of the SAH-optimal collapse to 8-wide nodes (the dynamic
   program of Ylitie, Karras, Laine 2017, section 3.2, restated for nodes whose leaf slots hold one primitive): for a binary node n
     T[n][i]  = cheapest way to present the subtree of n to a parent as at most i slots          (i = 1..7)
     D[n][j]  = min over k of T[left][k] + T[right][j - k]                                        (both sides get a slot)
     T[n][1]  = area( n ) * LH2B_SAH_NODE + D[n][8]      - n becomes a wide node and fills its own eight slots
     T[n][i]  = min( T[n][i - 1], D[n][i] )
     T[leaf][i] = area( leaf ) * LH2B_SAH_LEAF
   (every triangle ends up in exactly one leaf slot whatever is chosen, so the leaf term is the same for all collapses: the program minimises
   the summed area of the wide nodes; measured: LH2B_SAH_LEAF 0.1 .. 2.0 gives the same tree)
   dpSplit[n][i - 1] (i = 2..7) = the k behind T[n][i], 0 if T[n][i - 1] was kept; dpSplit[n][7] = the k of the node's own slots.
   collapseKernel then unfolds the choices top-down instead of opening the largest child greedily. */
#define LH2B_SAH_NODE 1.0f
#define LH2B_SAH_LEAF 0.3f
__global__ void fitKernel( const float4* __restrict__ primLo, const float4* __restrict__ primHi, const uint32_t* __restrict__ idx, const int n,
	const int2* __restrict__ children, const int* __restrict__ parent, uint32_t* __restrict__ visit, float4* __restrict__ nodeLo, float4* __restrict__ nodeHi,
	float* __restrict__ dpCost, uint8_t* __restrict__ dpSplit )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t p = idx[i];
	int node = (n - 1) + i;
	float4 lo = primLo[p], hi = primHi[p];
	nodeLo[node] = lo, nodeHi[node] = hi;
	if (n == 1) return;
	float tCur[7];	// T[node][1..7] of the node this thread carries upwards
	{
		const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z, leaf = (ex * ey + ey * ez + ez * ex) * LH2B_SAH_LEAF;
		for (int k = 0; k < 7; k++) tCur[k] = leaf;
	}
	int cur = parent[node];
	while (cur >= 0)
	{
		__threadfence();
		if (atomicAdd( visit + cur, 1u ) == 0) return;	// first child to arrive: the sibling will carry on
		const int2 c = children[cur];
		const int other = c.x == node ? c.y : c.x;
		const float4 olo = __ldcg( nodeLo + other ), ohi = __ldcg( nodeHi + other );	// written by another SM: bypass L1
		if (dpCost)
		{
			float tOther[7];
			if (other >= n - 1)
			{
				const float ex = ohi.x - olo.x, ey = ohi.y - olo.y, ez = ohi.z - olo.z, leaf = (ex * ey + ey * ez + ez * ex) * LH2B_SAH_LEAF;
				for (int k = 0; k < 7; k++) tOther[k] = leaf;
			}
			else for (int k = 0; k < 7; k++) tOther[k] = __ldcg( dpCost + (size_t)other * 7 + k );
			const bool nodeIsLeft = c.x == node;
			const float* tl = nodeIsLeft ? tCur : tOther;
			const float* tr = nodeIsLeft ? tOther : tCur;
			float d[9];
			uint8_t dk[9];
			for (int j = 2; j <= 8; j++)
			{
				float best = 3e38f;
				int bk = 1;
				for (int k = (j - 7 > 1 ? j - 7 : 1); k <= (j - 1 < 7 ? j - 1 : 7); k++)
				{
					const float v = tl[k - 1] + tr[j - k - 1];
					if (v < best) best = v, bk = k;
				}
				d[j] = best, dk[j] = (uint8_t)bk;
			}
			const float ulx = fminf( lo.x, olo.x ), uly = fminf( lo.y, olo.y ), ulz = fminf( lo.z, olo.z );
			const float uhx = fmaxf( hi.x, ohi.x ), uhy = fmaxf( hi.y, ohi.y ), uhz = fmaxf( hi.z, ohi.z );
			const float ex = uhx - ulx, ey = uhy - uly, ez = uhz - ulz;
			float tNew[7];
			uint8_t split[8];
			tNew[0] = (ex * ey + ey * ez + ez * ex) * LH2B_SAH_NODE + d[8], split[0] = 0, split[7] = dk[8];
			for (int k = 2; k <= 7; k++)
			{
				if (d[k] < tNew[k - 2]) tNew[k - 1] = d[k], split[k - 1] = dk[k];
				else tNew[k - 1] = tNew[k - 2], split[k - 1] = 0;
			}
			for (int k = 0; k < 7; k++) dpCost[(size_t)cur * 7 + k] = tNew[k], tCur[k] = tNew[k];
			*(uint2*)(dpSplit + (size_t)cur * 8) = make_uint2( split[0] | (split[1] << 8) | (split[2] << 16) | (split[3] << 24), split[4] | (split[5] << 8) | (split[6] << 16) | (split[7] << 24) );
		}
		lo = make_float4( fminf( lo.x, olo.x ), fminf( lo.y, olo.y ), fminf( lo.z, olo.z ), 0 );
		hi = make_float4( fmaxf( hi.x, ohi.x ), fmaxf( hi.y, ohi.y ), fmaxf( hi.z, ohi.z ), 0 );
		nodeLo[cur] = lo, nodeHi[cur] = hi;
		node = cur, cur = parent[cur];
	}
}

__global__ void resetVisitKernel( uint32_t* visit, const int n )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) visit[i] = 0;
}

/* stage 6 */
struct CollapseArgs
{
	int n;
	const int2* children; const uint32_t* subtree; const float4* nodeLo; const float4* nodeHi; const uint32_t* idx;
	const float4* verts;			// BLAS: triangle vertices; null for the TLAS
	uint4* outNodes;				// arena base
	float4* outTris;				// arena base (BLAS)
	uint32_t* outLeafIds;			// TLAS
	const uint32_t* linkedRootOf;	// TLAS, flat scenes: arena index of the BLAS root to copy per instance; null otherwise
	uint32_t nodeOffset, triOffset, nodeCapacity, triCapacity;
	BuildTask* queue; uint32_t* ctrl;
	float4* boundsOut;				// [0] = lo, [1] = hi of the root (kept on the device for the top-level build)
	uint32_t* countsOut;			// [0] node count, [1] leaf count, [2] overflow flag
	int* wideChild; int* wideSelf;	// BLAS: binary-tree node behind every slot of every wide node / behind the node itself (for refits)
	const uint8_t* dpSplit;			// choices of the SAH-optimal collapse (fitKernel); null: greedy largest-area opening
};

__device__ __forceinline__ float HalfAreaD( const float4 lo, const float4 hi )
{
	const float ex = hi.x - lo.x, ey = hi.y - lo.y, ez = hi.z - lo.z;
	return ex * ey + ey * ez + ez * ex;
}

__device__ void CollapseTask( const CollapseArgs& a, const BuildTask task )
{
	const int n = a.n;
	auto leafLike = [&]( const int node ) -> bool { return node >= n - 1; };	// binary leaves hold one primitive: one leaf slot each
	int child[8], cnt = 0;
	const int root = task.bvh2Node;
	if (leafLike( root )) child[cnt++] = root;
	else if (a.dpSplit)
	{
		// unfold the dynamic program: (binary node, slots granted) pairs until every pair is one slot
		int stackNode[8], stackSlots[8], sp = 0;
		{
			const int2 c = a.children[root];
			const int k = a.dpSplit[(size_t)root * 8 + 7];
			stackNode[sp] = c.y, stackSlots[sp++] = 8 - k, stackNode[sp] = c.x, stackSlots[sp++] = k;
		}
		while (sp > 0)
		{
			const int nd = stackNode[--sp];
			int slots = stackSlots[sp];
			if (leafLike( nd )) { child[cnt++] = nd; continue; }
			while (slots > 1 && a.dpSplit[(size_t)nd * 8 + slots - 1] == 0) slots--;
			if (slots == 1) { child[cnt++] = nd; continue; }	// stays a wide node of its own
			const int k = a.dpSplit[(size_t)nd * 8 + slots - 1];
			const int2 c = a.children[nd];
			stackNode[sp] = c.y, stackSlots[sp++] = slots - k, stackNode[sp] = c.x, stackSlots[sp++] = k;
		}
	}
	else
	{
		const int2 c = a.children[root];
		child[cnt++] = c.x, child[cnt++] = c.y;
		while (cnt < 8)
		{
			float bestA = -1;
			int bi = -1;
			for (int i = 0; i < cnt; i++) if (!leafLike( child[i] ))
			{
				const float ar = HalfAreaD( a.nodeLo[child[i]], a.nodeHi[child[i]] );
				if (ar > bestA) bestA = ar, bi = i;
			}
			if (bi < 0) break;
			const int2 c2 = a.children[child[bi]];
			child[bi] = c2.x, child[cnt++] = c2.y;
		}
	}
	const float4 rlo = a.nodeLo[root], rhi = a.nodeHi[root];
	// octant slot assignment (same greedy rule as the host builder)
	int slotChild[8];
	for (int s = 0; s < 8; s++) slotChild[s] = -1;
	{
		const float cx = 0.5f * (rlo.x + rhi.x), cy = 0.5f * (rlo.y + rhi.y), cz = 0.5f * (rlo.z + rhi.z);
		float dx[8], dy[8], dz[8];
		for (int i = 0; i < cnt; i++)
		{
			const float4 lo = a.nodeLo[child[i]], hi = a.nodeHi[child[i]];
			dx[i] = 0.5f * (lo.x + hi.x) - cx, dy[i] = 0.5f * (lo.y + hi.y) - cy, dz[i] = 0.5f * (lo.z + hi.z) - cz;
		}
		uint32_t slotUsed = 0, childDone = 0;
		for (int round = 0; round < cnt; round++)
		{
			float best = -FLT_MAX;
			int bi = -1, bs = -1;
			for (int i = 0; i < cnt; i++) if (!(childDone >> i & 1)) for (int s = 0; s < 8; s++) if (!(slotUsed >> s & 1))
			{
				const float sc = ((s & 4) ? dx[i] : -dx[i]) + ((s & 2) ? dy[i] : -dy[i]) + ((s & 1) ? dz[i] : -dz[i]);
				if (sc > best) best = sc, bi = i, bs = s;
			}
			slotChild[bs] = child[bi], slotUsed |= 1u << bs, childDone |= 1u << bi;
		}
	}
	if (a.wideChild)
	{
		a.wideSelf[task.cwNode] = root;
		for (int s = 0; s < 8; s++) a.wideChild[(size_t)task.cwNode * 8 + s] = slotChild[s];
	}
	// counts and allocation: internal children and leaf primitives are stored in slot order
	int internalCount = 0, leafPrims = 0;
	uint32_t imask = 0, lmask = 0;
	for (int s = 0; s < 8; s++) if (slotChild[s] >= 0)
	{
		const bool isLeaf = leafLike( slotChild[s] ) && !a.linkedRootOf;
		if (!isLeaf) imask |= 1u << s, internalCount++; else lmask |= 1u << s, leafPrims++;
	}
	const uint32_t childBase = internalCount ? atomicAdd( a.ctrl + 0, (uint32_t)internalCount ) : 0;
	const uint32_t leafBase = leafPrims ? atomicAdd( a.ctrl + 1, (uint32_t)leafPrims ) : 0;
	if (childBase + internalCount > a.nodeCapacity || leafBase + leafPrims > a.triCapacity) { a.ctrl[9] = 1; return; }
	uint32_t queueBase = 0;
	{
		int realInternal = 0;
		for (int s = 0; s < 8; s++) if ((imask >> s & 1) && !(a.linkedRootOf && slotChild[s] >= n - 1)) realInternal++;
		if (realInternal) queueBase = atomicAdd( a.ctrl + 2, (uint32_t)realInternal );
	}
	float clo[8][3], chi[8][3];
	int nextInternal = 0, nextQueued = 0, triCursor = 0;
	for (int s = 0; s < 8; s++)
	{
		const int c = slotChild[s];
		if (c < 0) continue;
		const float4 l4 = a.nodeLo[c], h4 = a.nodeHi[c];
		clo[s][0] = l4.x, clo[s][1] = l4.y, clo[s][2] = l4.z, chi[s][0] = h4.x, chi[s][1] = h4.y, chi[s][2] = h4.z;
		if (imask >> s & 1)
		{
			const uint32_t dst = childBase + nextInternal;
			if (a.linkedRootOf && c >= n - 1)
			{
				// flat scene: copy the BLAS root of this instance in as the child node
				const uint32_t inst = a.idx[c - (n - 1)];
				const uint4* src = a.outNodes + (size_t)a.linkedRootOf[inst] * CW_NODE_QUADS;
				uint4* d = a.outNodes + (size_t)(a.nodeOffset + dst) * CW_NODE_QUADS;
				for (int k = 0; k < CW_NODE_QUADS; k++) d[k] = src[k];
			}
			else a.queue[queueBase + nextQueued++] = BuildTask{ c, dst };
			nextInternal++;
		}
		else
		{
			const uint32_t prim = a.idx[c - (n - 1)];
			const uint32_t at = leafBase + triCursor++;
			if (a.verts)
			{
				const float4 v0 = a.verts[prim * 3], v1 = a.verts[prim * 3 + 1], v2 = a.verts[prim * 3 + 2];
				float4* t = a.outTris + (size_t)(a.triOffset + at) * 3;
				t[0] = make_float4( v0.x, v0.y, v0.z, __uint_as_float( prim ) );
				t[1] = make_float4( v1.x - v0.x, v1.y - v0.y, v1.z - v0.z, __uint_as_float( 0u ) );
				t[2] = make_float4( v2.x - v0.x, v2.y - v0.y, v2.z - v0.z, 0 );
			}
			else a.outLeafIds[at] = prim;
		}
	}
	uint32_t w[CW_NODE_WORDS];
	const float nlo[3] = { rlo.x, rlo.y, rlo.z }, nhi[3] = { rhi.x, rhi.y, rhi.z };
	w[4] = a.nodeOffset + childBase, w[5] = a.triOffset + leafBase, w[6] = imask | (lmask << 8), w[7] = 0;
	CwEncodePlanes( w, nlo, nhi, clo, chi, imask | lmask );
	uint4* out = a.outNodes + (size_t)(a.nodeOffset + task.cwNode) * CW_NODE_QUADS;
	for (int k = 0; k < CW_NODE_QUADS; k++) out[k] = make_uint4( w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3] );
}

__global__ void __launch_bounds__( 128 ) collapseKernel( const CollapseArgs a )
{
	cg::grid_group grid = cg::this_grid();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	if (tid == 0)
	{
		a.queue[0] = BuildTask{ 0, 0 };
		if (a.boundsOut) a.boundsOut[0] = a.nodeLo[0], a.boundsOut[1] = a.nodeHi[0];
	}
	grid.sync();
	uint32_t head = 0, tail = 1;
	while (head < tail)
	{
		for (uint32_t t = head + tid; t < tail; t += stride) CollapseTask( a, a.queue[t] );
		grid.sync();
		head = tail;
		tail = *((volatile uint32_t*)(a.ctrl + 2));
		grid.sync();
	}
	if (tid == 0 && a.countsOut) a.countsOut[0] = a.ctrl[0], a.countsOut[1] = a.ctrl[1], a.countsOut[2] = a.ctrl[9];
}


/* ---- refit of the wide tree in place -------------------------------------------------------------------------------
   After the binary tree has been refitted (fitKernel over the kept topology) every wide node is independent: its frame
   (origin, exponents) comes from the box of the binary node it was collapsed from, its eight quantised child boxes from the
   binary nodes recorded per slot at collapse time. Child links, triangle ranges, meta bytes and the octant slot order stay.
   Triangle records are rewritten from the new vertices; the primitive index and the instance tag they carry stay. */
__global__ void __launch_bounds__( 128 ) requantKernel( const int nodeCount, const int* __restrict__ wideChild, const int* __restrict__ wideSelf,
	const float4* __restrict__ nodeLo, const float4* __restrict__ nodeHi, uint4* __restrict__ outNodes, float4* __restrict__ boundsOut )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nodeCount) return;
	const int self = wideSelf[i];
	const float4 rlo = nodeLo[self], rhi = nodeHi[self];
	if (i == 0 && boundsOut) boundsOut[0] = rlo, boundsOut[1] = rhi;
	float clo[8][3], chi[8][3];
	uint32_t valid = 0;
	for (int s = 0; s < 8; s++)
	{
		const int c = wideChild[(size_t)i * 8 + s];
		if (c < 0) continue;
		const float4 l4 = nodeLo[c], h4 = nodeHi[c];
		clo[s][0] = l4.x, clo[s][1] = l4.y, clo[s][2] = l4.z, chi[s][0] = h4.x, chi[s][1] = h4.y, chi[s][2] = h4.z;
		valid |= 1u << s;
	}
	uint4* out = outNodes + (size_t)i * CW_NODE_QUADS;
	uint32_t w[CW_NODE_WORDS];
	const float nlo[3] = { rlo.x, rlo.y, rlo.z }, nhi[3] = { rhi.x, rhi.y, rhi.z };
	const uint4 keep = out[1];	// childBase, triBase and the slot masks stay
	w[6] = keep.z;
	CwEncodePlanes( w, nlo, nhi, clo, chi, valid );
	out[0] = make_uint4( w[0], w[1], w[2], w[3] );
	out[1] = make_uint4( keep.x, keep.y, w[6], keep.w );
	for (int k = 2; k < CW_NODE_QUADS; k++) out[k] = make_uint4( w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3] );
}

__global__ void triRewriteKernel( const int n, const float4* __restrict__ verts, float4* __restrict__ outTris )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4* t = outTris + (size_t)i * 3;
	const uint32_t prim = __float_as_uint( t[0].w );
	const float tag = t[1].w;
	const float4 v0 = verts[prim * 3], v1 = verts[prim * 3 + 1], v2 = verts[prim * 3 + 2];
	t[0] = make_float4( v0.x, v0.y, v0.z, __uint_as_float( prim ) );
	t[1] = make_float4( v1.x - v0.x, v1.y - v0.y, v1.z - v0.z, tag );
	t[2] = make_float4( v2.x - v0.x, v2.y - v0.y, v2.z - v0.z, 0 );
}

/* ---- stage 4': PLOC (parallel locally-ordered clustering, Meister & Bittner 2018) instead of the radix tree ------------
   Clusters start as the Morton-sorted leaves. Each round every cluster looks PLOC_RADIUS positions left and right for the
   neighbour whose union box has the smallest area; mutual nearest neighbours merge into a new binary node; the cluster
   array is compacted in order. One persistent cooperative kernel runs all rounds (three grid.sync() per round).
   Internal node ids are handed out downwards from n-2, so the last merge - the root - is node 0, like the radix tree. */
#define PLOC_BLOCK 256

struct PlocArgs
{
	int n, radius;
	int* clusterA; int* clusterB; int* nn; int* mergeTmp;
	float4* nodeLo; float4* nodeHi; int2* children; int* parent; uint32_t* subtree; uint32_t* visit;
	uint32_t* blockCount;	// per block valid count
	uint32_t* ctrl;			// [10] next internal id + 1, [11] current cluster count
};

__global__ void __launch_bounds__( PLOC_BLOCK ) plocKernel( const PlocArgs a )
{
	cg::grid_group grid = cg::this_grid();
	__shared__ uint32_t warpSums[PLOC_BLOCK / 32];
	__shared__ uint32_t blockBase;
	const int n = a.n;
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	int* in = a.clusterA;
	int* out = a.clusterB;
	for (uint32_t i = tid; i < (uint32_t)n; i += stride) in[i] = (n - 1) + i;
	if (tid == 0) a.ctrl[10] = n - 1, a.ctrl[11] = n, a.parent[0] = -1;
	grid.sync();
	int m = n;
	while (m > 1)
	{
		// phase 1: nearest neighbour within the window
		for (int i = tid; i < m; i += stride)
		{
			const int ci = in[i];
			const float4 lo = a.nodeLo[ci], hi = a.nodeHi[ci];
			float best = FLT_MAX;
			int bj = -1;
			const int j0 = max( 0, i - a.radius ), j1 = min( m - 1, i + a.radius );
			for (int j = j0; j <= j1; j++) if (j != i)
			{
				const int cj = in[j];
				const float4 l2 = a.nodeLo[cj], h2 = a.nodeHi[cj];
				const float ex = fmaxf( hi.x, h2.x ) - fminf( lo.x, l2.x ), ey = fmaxf( hi.y, h2.y ) - fminf( lo.y, l2.y ), ez = fmaxf( hi.z, h2.z ) - fminf( lo.z, l2.z );
				const float area = ex * ey + ey * ez + ez * ex;
				if (area < best) best = area, bj = j;
			}
			a.nn[i] = bj;
		}
		grid.sync();
		// phase 2: mutual pairs merge; every block owns a contiguous chunk of positions so that compaction keeps the order
		const int chunk = (m + gridDim.x - 1) / gridDim.x;
		const int c0 = blockIdx.x * chunk, c1 = min( m, c0 + chunk );
		uint32_t blockTotal = 0;
		for (int base = c0; base < c1; base += PLOC_BLOCK)
		{
			const int i = base + threadIdx.x;
			int result = -1;	// -1: this position disappears
			if (i < c1)
			{
				const int j = a.nn[i];
				if (a.nn[j] == i)
				{
					if (i < j)
					{
						const int ci = in[i], cj = in[j];
						const int id = (int)atomicSub( a.ctrl + 10, 1u ) - 1;
						const float4 l1 = a.nodeLo[ci], h1 = a.nodeHi[ci], l2 = a.nodeLo[cj], h2 = a.nodeHi[cj];
						a.nodeLo[id] = make_float4( fminf( l1.x, l2.x ), fminf( l1.y, l2.y ), fminf( l1.z, l2.z ), 0 );
						a.nodeHi[id] = make_float4( fmaxf( h1.x, h2.x ), fmaxf( h1.y, h2.y ), fmaxf( h1.z, h2.z ), 0 );
						a.children[id] = make_int2( ci, cj );
						a.parent[ci] = id, a.parent[cj] = id;
						a.subtree[id] = (ci >= n - 1 ? 1u : a.subtree[ci]) + (cj >= n - 1 ? 1u : a.subtree[cj]);
						a.visit[id] = 0;
						result = id;
					}
				}
				else result = in[i];
				a.mergeTmp[i] = result;
			}
			blockTotal += __syncthreads_count( result >= 0 );
		}
		if (threadIdx.x == 0) a.blockCount[blockIdx.x] = blockTotal;
		grid.sync();
		// phase 3: ordered compaction. Block offset = sum of the counts of the blocks before it.
		{
			uint32_t partial = 0;
			for (uint32_t b = threadIdx.x; b < blockIdx.x; b += PLOC_BLOCK) partial += a.blockCount[b];
			for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync( 0xffffffffu, partial, o );
			if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = partial;
			__syncthreads();
			if (threadIdx.x == 0)
			{
				uint32_t t = 0;
				for (int w = 0; w < PLOC_BLOCK / 32; w++) t += warpSums[w];
				blockBase = t;
			}
			__syncthreads();
		}
		uint32_t running = blockBase;
		for (int base = c0; base < c1; base += PLOC_BLOCK)
		{
			const int i = base + threadIdx.x;
			const int v = i < c1 ? a.mergeTmp[i] : -1;
			const bool keep = v >= 0;
			const uint32_t ballot = __ballot_sync( 0xffffffffu, keep );
			const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
			__syncthreads();
			if (lane == 0) warpSums[warp] = __popc( ballot );
			__syncthreads();
			uint32_t before = 0, total = 0;
			for (int w = 0; w < PLOC_BLOCK / 32; w++) { const uint32_t c = warpSums[w]; if (w < (int)warp) before += c; total += c; }
			if (keep) out[running + before + __popc( ballot & ((1u << lane) - 1) )] = v;
			running += total;
		}
		// new cluster count
		uint32_t total = 0;
		{
			uint32_t partial = 0;
			for (uint32_t b = threadIdx.x; b < gridDim.x; b += PLOC_BLOCK) partial += a.blockCount[b];
			for (int o = 16; o > 0; o >>= 1) partial += __shfl_xor_sync( 0xffffffffu, partial, o );
			__syncthreads();
			if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = partial;
			__syncthreads();
			for (int w = 0; w < PLOC_BLOCK / 32; w++) total += warpSums[w];
		}
		grid.sync();
		m = (int)total;
		int* t = in; in = out, out = t;
	}
}

static int PlocGrid( lh2b_core* core )
{
	static int blocksPerSM = 0;
	if (!blocksPerSM) CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSM, plocKernel, PLOC_BLOCK, 0 ) );
	return (int)core->stats.SMcount * (blocksPerSM > 0 ? blocksPerSM : 1);
}

static void BuildPloc( lh2b_core* core, GpuBuildScratch& s, const int n );

static GpuBuildScratch& Scratch( lh2b_core* core )
{
	if (!core->gpuBuild) core->gpuBuild = new GpuBuildScratch();
	return *(GpuBuildScratch*)core->gpuBuild;
}

void ReleaseGpuBuildScratch( lh2b_core* core )
{
	delete (GpuBuildScratch*)core->gpuBuild;
	core->gpuBuild = nullptr;
}

static int CollapseGrid( lh2b_core* core )
{
	static int blocksPerSM = 0;
	if (!blocksPerSM) CUDA_CHECK( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &blocksPerSM, collapseKernel, 128, 0 ) );
	return (int)core->stats.SMcount * (blocksPerSM > 0 ? blocksPerSM : 1);
}

/* Shared tail of BLAS and TLAS builds: stages 2-6 over n primitive boxes already in scratch.primLo/primHi.
   sortTopology = false: keep the sorted order and the radix tree of the previous build (refit). */
static void BuildFromBoxes( lh2b_core* core, GpuBuildScratch& s, const int n, const bool sortTopology, const bool usePloc, CollapseArgs args, const bool keepWideTree = false, const int wideNodes = 0 )
{
	cudaStream_t st = core->stream;
	const int blocks = (n + 255) / 256;
	if (sortTopology)
	{
		s.keys.Resize( n ), s.keysAlt.Resize( n ), s.idx.Resize( n ), s.idxAlt.Resize( n );
		mortonKernel<<<blocks, 256, 0, st>>>( s.primLo.ptr, s.primHi.ptr, n, s.ctrl.ptr, s.keys.ptr, s.idx.ptr );
		size_t tempBytes = 0;
		cub::DeviceRadixSort::SortPairs( nullptr, tempBytes, s.keys.ptr, s.keysAlt.ptr, s.idx.ptr, s.idxAlt.ptr, n, 0, 63, st );
		s.cubTemp.Resize( tempBytes + 16 );
		cub::DeviceRadixSort::SortPairs( s.cubTemp.ptr, tempBytes, s.keys.ptr, s.keysAlt.ptr, s.idx.ptr, s.idxAlt.ptr, n, 0, 63, st );
		s.children.Resize( n ), s.subtree.Resize( n ), s.parent.Resize( 2 * n ), s.visit.Resize( n );
		if (n > 1 && usePloc) BuildPloc( core, s, n );
		else if (n > 1) radixTreeKernel<<<blocks, 256, 0, st>>>( s.keysAlt.ptr, n, s.children.ptr, s.subtree.ptr, s.parent.ptr, s.visit.ptr );
	}
	else if (n > 1) resetVisitKernel<<<blocks, 256, 0, st>>>( s.visit.ptr, n - 1 );
	s.nodeLo.Resize( 2 * n ), s.nodeHi.Resize( 2 * n );
	const bool optimal = core->bvhCollapse == 1 && !keepWideTree && n > 1;
	if (optimal) s.dpCost.Resize( (size_t)n * 7 ), s.dpSplit.Resize( (size_t)n * 8 );
	fitKernel<<<blocks, 256, 0, st>>>( s.primLo.ptr, s.primHi.ptr, s.idxAlt.ptr, n, s.children.ptr, s.parent.ptr, s.visit.ptr, s.nodeLo.ptr, s.nodeHi.ptr,
		optimal ? s.dpCost.ptr : nullptr, optimal ? s.dpSplit.ptr : nullptr );
	if (keepWideTree)
	{
		requantKernel<<<(wideNodes + 127) / 128, 128, 0, st>>>( wideNodes, args.wideChild, args.wideSelf, s.nodeLo.ptr, s.nodeHi.ptr,
			args.outNodes + (size_t)args.nodeOffset * CW_NODE_QUADS, args.boundsOut );
		triRewriteKernel<<<blocks, 256, 0, st>>>( n, args.verts, args.outTris + (size_t)args.triOffset * 3 );
		return;
	}
	s.queue.Resize( (size_t)n + 8 );
	args.n = n, args.children = s.children.ptr, args.subtree = s.subtree.ptr, args.nodeLo = s.nodeLo.ptr, args.nodeHi = s.nodeHi.ptr;
	args.idx = s.idxAlt.ptr, args.queue = s.queue.ptr, args.ctrl = s.ctrl.ptr, args.dpSplit = optimal ? s.dpSplit.ptr : nullptr;
	void* params[] = { &args };
	CUDA_CHECK( cudaLaunchCooperativeKernel( (void*)collapseKernel, dim3( CollapseGrid( core ) ), dim3( 128 ), params, 0, st ) );
}

__global__ void leafBoxKernel( const float4* __restrict__ primLo, const float4* __restrict__ primHi, const uint32_t* __restrict__ idx, const int n,
	float4* __restrict__ nodeLo, float4* __restrict__ nodeHi )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t p = idx[i];
	nodeLo[(n - 1) + i] = primLo[p], nodeHi[(n - 1) + i] = primHi[p];
}

static void BuildPloc( lh2b_core* core, GpuBuildScratch& s, const int n )
{
	cudaStream_t st = core->stream;
	s.nodeLo.Resize( 2 * n ), s.nodeHi.Resize( 2 * n );
	s.clusterA.Resize( n ), s.clusterB.Resize( n ), s.mergeTmp.Resize( 2 * (size_t)n );
	const int grid = PlocGrid( core );
	s.blockCount.Resize( grid );
	leafBoxKernel<<<(n + 255) / 256, 256, 0, st>>>( s.primLo.ptr, s.primHi.ptr, s.idxAlt.ptr, n, s.nodeLo.ptr, s.nodeHi.ptr );
	PlocArgs a;
	a.n = n, a.radius = core->plocRadius, a.clusterA = s.clusterA.ptr, a.clusterB = s.clusterB.ptr, a.nn = s.mergeTmp.ptr + n, a.mergeTmp = s.mergeTmp.ptr;
	a.nodeLo = s.nodeLo.ptr, a.nodeHi = s.nodeHi.ptr, a.children = s.children.ptr, a.parent = s.parent.ptr, a.subtree = s.subtree.ptr, a.visit = s.visit.ptr;
	a.blockCount = s.blockCount.ptr, a.ctrl = s.ctrl.ptr;
	void* params[] = { &a };
	CUDA_CHECK( cudaLaunchCooperativeKernel( (void*)plocKernel, dim3( grid ), dim3( PLOC_BLOCK ), params, 0, st ) );
}

/* BLAS build (or refit) of one mesh from its device-resident vertices. Per-mesh topology (sorted order + radix tree) is
   kept in the mesh so that a later SetGeometry with the same triangle count can refit. */
void GpuBuildMesh( lh2b_core* core, Mesh& mesh, const int refit )
{
	GpuBuildScratch& s = Scratch( core );
	cudaStream_t st = core->stream;
	const int n = mesh.triCount;
	s.ctrl.Resize( 16 );
	initCtrlKernel<<<1, 1, 0, st>>>( s.ctrl.ptr );
	if (n == 0)
	{
		// empty mesh: one node without children, zero box
		CUDA_CHECK( cudaMemsetAsync( core->arenaNodes.ptr + (size_t)mesh.nodeOff * CW_NODE_QUADS, 0, sizeof( CwNode ), st ) );
		CUDA_CHECK( cudaMemsetAsync( mesh.devBounds.ptr, 0, 32, st ) );
		CUDA_CHECK( cudaMemsetAsync( mesh.devCounts.ptr, 0, 16, st ) );
		return;
	}
	s.primLo.Resize( n ), s.primHi.Resize( n );
	triBoundsKernel<<<(n + 255) / 256, 256, 0, st>>>( mesh.verts.ptr, n, s.primLo.ptr, s.primHi.ptr, s.ctrl.ptr );
	// per-mesh topology lives in the mesh's own buffers: swap them into the scratch for the duration of the build
	s.idxAlt.Swap( mesh.topoIdx ), s.children.Swap( mesh.topoChildren ), s.subtree.Swap( mesh.topoSubtree );
	s.parent.Swap( mesh.topoParent ), s.visit.Swap( mesh.topoVisit );
	CollapseArgs a = {};
	a.verts = mesh.verts.ptr, a.outNodes = core->arenaNodes.ptr, a.outTris = core->arenaTris.ptr;
	a.nodeOffset = mesh.nodeOff, a.triOffset = mesh.triOff, a.nodeCapacity = mesh.nodeCap, a.triCapacity = mesh.triCap;
	a.boundsOut = mesh.devBounds.ptr, a.countsOut = mesh.devCounts.ptr;
	if (mesh.topoWideSelf.count < mesh.nodeCap) mesh.topoWideSelf.Resize( mesh.nodeCap ), mesh.topoWideChild.Resize( (size_t)mesh.nodeCap * 8 );
	a.wideChild = mesh.topoWideChild.ptr, a.wideSelf = mesh.topoWideSelf.ptr;
	// refit 0: full build | 1: binary-tree refit + requantisation of the wide tree in place | 2: binary-tree refit + new collapse
	BuildFromBoxes( core, s, n, refit == 0, core->bvhBuilder != 2, a, refit == 1, (int)mesh.nodeCount );
	s.idxAlt.Swap( mesh.topoIdx ), s.children.Swap( mesh.topoChildren ), s.subtree.Swap( mesh.topoSubtree );
	s.parent.Swap( mesh.topoParent ), s.visit.Swap( mesh.topoVisit );
}

/* TLAS build over instance boxes; everything stays on the device. */
void GpuBuildTlas( lh2b_core* core, const void* dInstIn, const int n, const uint32_t* dLinkedRoots )
{
	GpuBuildScratch& s = Scratch( core );
	cudaStream_t st = core->stream;
	s.ctrl.Resize( 16 );
	initCtrlKernel<<<1, 1, 0, st>>>( s.ctrl.ptr );
	if (n == 0) return;
	s.primLo.Resize( n ), s.primHi.Resize( n );
	instBoundsKernel<<<(n + 127) / 128, 128, 0, st>>>( (const InstBuildIn*)dInstIn, n, s.primLo.ptr, s.primHi.ptr, s.ctrl.ptr );
	CollapseArgs a = {};
	a.verts = nullptr, a.outNodes = core->arenaNodes.ptr, a.outLeafIds = core->tlasLeafIds.ptr, a.linkedRootOf = dLinkedRoots;
	a.nodeOffset = core->tlasOff, a.triOffset = 0, a.nodeCapacity = core->tlasCap, a.triCapacity = (uint32_t)core->tlasLeafIds.capacity;
	BuildFromBoxes( core, s, n, true, false, a );
}

} // namespace lh2b
