/* core.cu - C ABI of the B200 render core: scene ownership, acceleration-structure upkeep and
   the ray-query entry points. The render loop itself lives in render.cu.

   Mirrors, per entry point, the reference Optix7 core (lib/rendercore_optix7/rendercore.cpp);
   see include/lh2b.h for the 1:1 citation list.
*/
#include "core.h"
#include "kernels.h"
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace lh2b
{

void InitRenderState( lh2b_core* core );
void ReleaseRenderState( lh2b_core* core );

static thread_local std::string g_lastError;
void SetLastError( const std::string& msg ) { g_lastError = msg; }

static double NowMs()
{
	return std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now().time_since_epoch() ).count();
}

/* 3x4 affine inverse, float arithmetic, cofactor form (the reference inverts the full 4x4 with the
   MESA routine, lib/RenderSystem/common_types.h:545-590; for an affine matrix both give the same
   rows up to rounding; the oracle restates this exact routine so traversal parity is bit-exact). */
void InvertAffine( const float* m, float* inv )
{
	const float a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
	const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
	const float det = a * A + b * B + c * C;
	const float id = det != 0 ? 1.0f / det : 0.0f;
	inv[0] = A * id, inv[1] = -(b * i - c * h) * id, inv[2] = (b * f - c * e) * id;
	inv[4] = B * id, inv[5] = (a * i - c * g) * id, inv[6] = -(a * f - c * d) * id;
	inv[8] = C * id, inv[9] = -(a * h - b * g) * id, inv[10] = (a * e - b * d) * id;
	inv[3] = -(inv[0] * m[3] + inv[1] * m[7] + inv[2] * m[11]);
	inv[7] = -(inv[4] * m[3] + inv[5] * m[7] + inv[6] * m[11]);
	inv[11] = -(inv[8] * m[3] + inv[9] * m[7] + inv[10] * m[11]);
}

static bool IsIdentity( const float* m )
{
	static const float id[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
	return memcmp( m, id, sizeof( id ) ) == 0;
}

/* Arena slots: bump allocation with 25% slack; a mesh that outgrows its slot gets a new one at the top
   (the old slot is not recycled - meshes that change triangle count every frame should be rare). */
static uint32_t ArenaAllocNodes( lh2b_core* core, uint32_t count, uint32_t& cap )
{
	cap = count + count / 4 + 4;
	const uint32_t off = core->arenaNodeTop;
	core->arenaNodeTop += cap;
	const size_t need = (size_t)core->arenaNodeTop * 5;
	if (need > core->arenaNodes.capacity)
	{
		core->arenaNodes.count = core->arenaNodes.capacity;
		core->arenaNodes.Reserve( need + need / 2, true );
	}
	core->arenaNodes.count = need;
	return off;
}

static uint32_t ArenaAllocTris( lh2b_core* core, uint32_t count, uint32_t& cap )
{
	cap = count + count / 4 + 4;
	const uint32_t off = core->arenaTriTop;
	core->arenaTriTop += cap;
	const size_t need = (size_t)core->arenaTriTop * 3;
	if (need > core->arenaTris.capacity)
	{
		core->arenaTris.count = core->arenaTris.capacity;
		core->arenaTris.Reserve( need + need / 2, true );
	}
	core->arenaTris.count = need;
	return off;
}

/* childBase / triBase of every node become absolute arena indices. */
static void BakeOffsets( std::vector<CwNode>& nodes, uint32_t nodeOff, uint32_t triOff )
{
	for (auto& n : nodes) n.w[4] += nodeOff, n.w[5] += triOff;
}

static void RebuildMeshHost( lh2b_core* core, Mesh& mesh )
{
	const double t0 = NowMs();
	std::vector<Bvh2Node> bvh2;
	std::vector<uint32_t> primIdx;
	BuildBvh2SAH( mesh.hostVerts.data(), mesh.triCount, bvh2, primIdx );
	CwBvh cw;
	CollapseToCwBvh( bvh2, primIdx, mesh.hostVerts.data(), cw );
	mesh.bounds = cw.bounds;
	mesh.nodeCount = (uint32_t)cw.nodes.size();
	if (cw.tris.empty()) cw.tris.push_back( CwTri{} );
	if (mesh.nodeCount > mesh.nodeCap) mesh.nodeOff = ArenaAllocNodes( core, mesh.nodeCount, mesh.nodeCap );
	if (cw.tris.size() > mesh.triCap) mesh.triOff = ArenaAllocTris( core, (uint32_t)cw.tris.size(), mesh.triCap );
	BakeOffsets( cw.nodes, mesh.nodeOff, mesh.triOff );	// slots are known only after the build: bake now
	mesh.rootNode = cw.nodes[0], mesh.taggedInst = 0;	// triangle records are emitted with inst = 0
	CUDA_CHECK( cudaMemcpyAsync( core->arenaNodes.ptr + (size_t)mesh.nodeOff * 5, cw.nodes.data(), cw.nodes.size() * sizeof( CwNode ), cudaMemcpyHostToDevice, core->stream ) );
	CUDA_CHECK( cudaMemcpyAsync( core->arenaTris.ptr + (size_t)mesh.triOff * 3, cw.tris.data(), cw.tris.size() * sizeof( CwTri ), cudaMemcpyHostToDevice, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	mesh.buildMs = (float)(NowMs() - t0);
	mesh.dirty = false;
}

static void TransformBounds( const Aabb& b, const float* m, Aabb& out )
{
	for (int a = 0; a < 3; a++) out.lo[a] = 1e34f, out.hi[a] = -1e34f;
	for (int k = 0; k < 8; k++)
	{
		const float x = (k & 1) ? b.hi[0] : b.lo[0], y = (k & 2) ? b.hi[1] : b.lo[1], z = (k & 4) ? b.hi[2] : b.lo[2];
		for (int a = 0; a < 3; a++)
		{
			const float v = m[a * 4] * x + m[a * 4 + 1] * y + m[a * 4 + 2] * z + m[a * 4 + 3];
			out.lo[a] = fminf( out.lo[a], v ), out.hi[a] = fmaxf( out.hi[a], v );
		}
	}
	// pad: the object-space traversal re-derives the ray with rounding, keep the world box conservative
	for (int a = 0; a < 3; a++)
	{
		const float pad = 1e-5f * fmaxf( fabsf( out.lo[a] ), fabsf( out.hi[a] ) ) + 1e-30f;
		out.lo[a] -= pad, out.hi[a] += pad;
	}
}

void UpdateAccelerationStructures( lh2b_core* core )
{
	for (auto& m : core->meshes) if (m->dirty) RebuildMeshHost( core, *m );
	const double t0 = NowMs();
	const int n = (int)core->instances.size();
	std::vector<InstTrav> trav( n > 0 ? n : 1 );
	std::vector<lh2abi::CoreInstanceDesc> desc( n > 0 ? n : 1 );
	std::vector<Aabb> boxes( n );
	for (int i = 0; i < n; i++)
	{
		const Instance& inst = core->instances[i];
		const Mesh& mesh = *core->meshes[inst.mesh];
		float inv[12];
		InvertAffine( inst.xform, inv );
		trav[i].r0 = make_float4( inv[0], inv[1], inv[2], inv[3] );
		trav[i].r1 = make_float4( inv[4], inv[5], inv[6], inv[7] );
		trav[i].r2 = make_float4( inv[8], inv[9], inv[10], inv[11] );
		trav[i].rootNode = mesh.nodeOff, trav[i].flags = IsIdentity( inst.xform ) ? 1u : 0u, trav[i].pad0 = trav[i].pad1 = 0;
		// shading-side descriptor: triangle array + inverse transform (rendercore.cpp:403-417)
		desc[i].triangles = mesh.coreTris.ptr, desc[i].dummy1 = desc[i].dummy2 = 0;
		desc[i].invTransform.A = { inv[0], inv[1], inv[2], inv[3] }, desc[i].invTransform.B = { inv[4], inv[5], inv[6], inv[7] };
		desc[i].invTransform.C = { inv[8], inv[9], inv[10], inv[11] }, desc[i].invTransform.D = { 0, 0, 0, 1 };
		TransformBounds( mesh.bounds, inst.xform, boxes[i] );
	}
	// flat scene: every instance has the identity transform and no mesh is instanced twice. Then no ray ever needs
	// transforming: the top level links copies of the BLAS roots as internal children and traversal is single-level;
	// the instance index of a hit comes from the triangle record.
	bool flat = n > 0;
	{
		std::vector<int> uses( core->meshes.size(), 0 );
		for (int i = 0; i < n; i++) if (!IsIdentity( core->instances[i].xform ) || ++uses[core->instances[i].mesh] > 1) flat = false;
	}
	std::vector<CwNode> linked;
	if (flat)
	{
		linked.resize( n );
		for (int i = 0; i < n; i++)
		{
			Mesh& mesh = *core->meshes[core->instances[i].mesh];
			linked[i] = mesh.rootNode;
			if (mesh.taggedInst != i)
			{
				LaunchTagTriangles( core->arenaTris.ptr + (size_t)mesh.triOff * 3, mesh.triCount, (uint32_t)i, core->stream );
				mesh.taggedInst = i;
			}
		}
	}
	std::vector<Bvh2Node> bvh2;
	std::vector<uint32_t> primIdx;
	BuildBvh2FromBoxes( boxes.data(), n, 1, bvh2, primIdx );
	// the top level's slot must exist before encoding (linked roots are absolute already and must not be re-based)
	const uint32_t tlasNeed = (uint32_t)(2 * n + 2);
	if (tlasNeed > core->tlasCap) core->tlasOff = ArenaAllocNodes( core, tlasNeed, core->tlasCap );
	CwBvh cw;
	CollapseToCwBvh( bvh2, primIdx, nullptr, cw, core->tlasOff, 0, flat ? linked.data() : nullptr );
	if (cw.leafIds.empty()) cw.leafIds.push_back( 0 );
	core->tlasNodeCount = (uint32_t)cw.nodes.size();
	if (core->tlasNodeCount > core->tlasCap) throw CoreError( "internal: top-level node estimate too small" );
	CUDA_CHECK( cudaMemcpyAsync( core->arenaNodes.ptr + (size_t)core->tlasOff * 5, cw.nodes.data(), cw.nodes.size() * sizeof( CwNode ), cudaMemcpyHostToDevice, core->stream ) );
	core->tlasLeafIds.Upload( cw.leafIds.data(), cw.leafIds.size(), core->stream );
	core->instTrav.Upload( trav.data(), trav.size(), core->stream );
	core->instDesc.Upload( desc.data(), desc.size(), core->stream );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	core->scene.nodes = core->arenaNodes.ptr, core->scene.tris = core->arenaTris.ptr;
	core->scene.tlasRoot = core->tlasOff;
	// flat scenes start at the top-level root (whose children are BLAS-root copies), or directly at the BLAS root
	core->scene.singleRoot = !flat ? 0 : (n == 1 ? core->meshes[core->instances[0].mesh]->nodeOff : core->tlasOff);
	core->scene.tlasLeafIds = core->tlasLeafIds.ptr;
	core->scene.instances = core->instTrav.ptr;
	core->scene.instanceCount = n;
	core->scene.singleIdentity = flat ? 1 : 0;
	core->tlasBuildMs = (float)(NowMs() - t0);
}

} // namespace lh2b

using namespace lh2b;

#define API_BEGIN if (!core) { SetLastError( "null core handle" ); return 1; } try { CUDA_CHECK( cudaSetDevice( core->device ) );
#define API_END } catch (const std::exception& e) { SetLastError( e.what() ); return 1; } return 0;

extern "C" {

const char* lh2b_last_error( void ) { return g_lastError.c_str(); }

int lh2b_create( lh2b_core** out, int device )
{
	if (!out) { SetLastError( "null out pointer" ); return 1; }
	*out = nullptr;
	try
	{
		int count = 0;
		cudaError_t e = cudaGetDeviceCount( &count );
		if (e != cudaSuccess || count == 0)
			throw CoreError( std::string( "no CUDA device available (" ) + cudaGetErrorString( e ) + "); this core has no CPU fallback" );
		if (device < 0)
		{
			const char* lr = getenv( "LOCAL_RANK" );
			device = lr ? atoi( lr ) % count : 0;
		}
		if (device >= count) throw CoreError( "device index out of range" );
		CUDA_CHECK( cudaSetDevice( device ) );
		cudaDeviceProp prop;
		CUDA_CHECK( cudaGetDeviceProperties( &prop, device ) );
		std::unique_ptr<lh2b_core> core( new lh2b_core() );
		core->device = device;
		CUDA_CHECK( cudaStreamCreateWithFlags( &core->stream, cudaStreamNonBlocking ) );
		CUDA_CHECK( cudaEventCreate( &core->evA ) );
		CUDA_CHECK( cudaEventCreate( &core->evB ) );
		// device fields of CoreStats (rendercore.cpp:231-239)
		core->stats.SMcount = prop.multiProcessorCount;
		core->stats.ccMajor = prop.major, core->stats.ccMinor = prop.minor;
		core->stats.VRAM = (uint32_t)(prop.totalGlobalMem >> 20);
		core->stats.deviceName = new char[strlen( prop.name ) + 1];
		strcpy( core->stats.deviceName, prop.name );
		core->stats.probedTriid = -1;
		core->queryCounter.Resize( 4 );
		InitRenderState( core.get() );
		*out = core.release();
	}
	catch (const std::exception& e) { SetLastError( e.what() ); return 1; }
	return 0;
}

int lh2b_destroy( lh2b_core* core )
{
	if (!core) return 0;
	cudaSetDevice( core->device );
	cudaStreamSynchronize( core->stream );
	ReleaseRenderState( core );
	cudaEventDestroy( core->evA ), cudaEventDestroy( core->evB );
	cudaStreamDestroy( core->stream );
	delete[] core->stats.deviceName;
	delete core;
	return 0;
}

int lh2b_stream( lh2b_core* core, void** streamOut )
{
	API_BEGIN
	*streamOut = (void*)core->stream;
	API_END
}

int lh2b_set_geometry( lh2b_core* core, int meshIdx, const float* vertexData, int vertexCount, int triangleCount, const void* triangles )
{
	API_BEGIN
	if (meshIdx < 0 || meshIdx > (int)core->meshes.size()) throw CoreError( "SetGeometry: meshes must be introduced in sequential order" );
	if (vertexCount != triangleCount * 3) throw CoreError( "SetGeometry: vertexCount must be 3 * triangleCount" );
	if (meshIdx == (int)core->meshes.size()) core->meshes.emplace_back( new Mesh() );
	Mesh& mesh = *core->meshes[meshIdx];
	mesh.triCount = triangleCount;
	mesh.hostVerts.assign( vertexData, vertexData + (size_t)vertexCount * 4 );
	mesh.verts.Upload( (const float4*)vertexData, (size_t)vertexCount, core->stream );
	if (triangles) mesh.coreTris.Upload( (const float4*)triangles, (size_t)triangleCount * 13, core->stream );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) ); // caller may free its arrays on return
	mesh.dirty = true;
	API_END
}

int lh2b_set_instance( lh2b_core* core, int instanceIdx, int meshIdx, const float* transform )
{
	API_BEGIN
	if (instanceIdx < 0) throw CoreError( "SetInstance: negative instance index" );
	if (meshIdx == -1)
	{
		if ((int)core->instances.size() > instanceIdx) core->instances.resize( instanceIdx );
		return 0;
	}
	if (meshIdx < 0 || meshIdx >= (int)core->meshes.size()) throw CoreError( "SetInstance: unknown mesh" );
	if (instanceIdx > (int)core->instances.size()) throw CoreError( "SetInstance: instances must be introduced in sequential order" );
	if (instanceIdx == (int)core->instances.size()) core->instances.emplace_back();
	Instance& inst = core->instances[instanceIdx];
	inst.mesh = meshIdx;
	if (transform) memcpy( inst.xform, transform, 12 * sizeof( float ) );
	API_END
}

int lh2b_finalize_instances( lh2b_core* core )
{
	API_BEGIN
	UpdateAccelerationStructures( core );
	core->sceneReady = true;
	API_END
}

int lh2b_trace_rays_device( lh2b_core* core, const void* dO, const void* dD, int n, void* dHits, int repeat, float* msOut )
{
	API_BEGIN
	if (!core->sceneReady) throw CoreError( "trace: FinalizeInstances has not been called" );
	if (repeat < 1) repeat = 1;
	if (msOut) CUDA_CHECK( cudaEventRecord( core->evA, core->stream ) );
	for (int r = 0; r < repeat; r++) LaunchExtend( core->scene, (const float4*)dO, (const float4*)dD, (float4*)dHits, n, core->queryCounter.ptr, (int)core->stats.SMcount, core->stream );
	CUDA_CHECK( cudaGetLastError() );
	if (msOut)
	{
		CUDA_CHECK( cudaEventRecord( core->evB, core->stream ) );
		CUDA_CHECK( cudaEventSynchronize( core->evB ) );
		CUDA_CHECK( cudaEventElapsedTime( msOut, core->evA, core->evB ) );
	}
	API_END
}

int lh2b_trace_shadow_rays_device( lh2b_core* core, const void* dO, const void* dD, int n, void* dOcc, int repeat, float* msOut )
{
	API_BEGIN
	if (!core->sceneReady) throw CoreError( "trace: FinalizeInstances has not been called" );
	if (repeat < 1) repeat = 1;
	if (msOut) CUDA_CHECK( cudaEventRecord( core->evA, core->stream ) );
	for (int r = 0; r < repeat; r++) LaunchOcclude( core->scene, (const float4*)dO, (const float4*)dD, (uint8_t*)dOcc, n, core->queryCounter.ptr, (int)core->stats.SMcount, core->stream );
	CUDA_CHECK( cudaGetLastError() );
	if (msOut)
	{
		CUDA_CHECK( cudaEventRecord( core->evB, core->stream ) );
		CUDA_CHECK( cudaEventSynchronize( core->evB ) );
		CUDA_CHECK( cudaEventElapsedTime( msOut, core->evA, core->evB ) );
	}
	API_END
}

int lh2b_trace_rays( lh2b_core* core, const float* origins, const float* directions, int n, float* hitsOut )
{
	API_BEGIN
	if (!core->sceneReady) throw CoreError( "trace: FinalizeInstances has not been called" );
	core->qO.Upload( (const float4*)origins, n, core->stream );
	core->qD.Upload( (const float4*)directions, n, core->stream );
	core->qHits.Resize( n );
	LaunchExtend( core->scene, core->qO.ptr, core->qD.ptr, core->qHits.ptr, n, core->queryCounter.ptr, (int)core->stats.SMcount, core->stream );
	CUDA_CHECK( cudaGetLastError() );
	CUDA_CHECK( cudaMemcpyAsync( hitsOut, core->qHits.ptr, (size_t)n * 16, cudaMemcpyDeviceToHost, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_trace_shadow_rays( lh2b_core* core, const float* origins, const float* directions, int n, uint8_t* occludedOut )
{
	API_BEGIN
	if (!core->sceneReady) throw CoreError( "trace: FinalizeInstances has not been called" );
	core->qO.Upload( (const float4*)origins, n, core->stream );
	core->qD.Upload( (const float4*)directions, n, core->stream );
	core->qOcc.Resize( n );
	LaunchOcclude( core->scene, core->qO.ptr, core->qD.ptr, core->qOcc.ptr, n, core->queryCounter.ptr, (int)core->stats.SMcount, core->stream );
	CUDA_CHECK( cudaGetLastError() );
	CUDA_CHECK( cudaMemcpyAsync( occludedOut, core->qOcc.ptr, (size_t)n, cudaMemcpyDeviceToHost, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_get_bvh_stats( lh2b_core* core, int meshIdx, lh2b_bvh_stats* out )
{
	API_BEGIN
	memset( out, 0, sizeof( *out ) );
	if (meshIdx == -1)
	{
		out->nodes = core->tlasNodeCount, out->triangles = (uint32_t)core->instances.size();
		out->bytes = (uint32_t)(core->tlasNodeCount * 80 + core->tlasLeafIds.Bytes() + core->instTrav.Bytes());
		out->buildMs = core->tlasBuildMs;
	}
	else
	{
		if (meshIdx < 0 || meshIdx >= (int)core->meshes.size()) throw CoreError( "unknown mesh" );
		const Mesh& m = *core->meshes[meshIdx];
		out->nodes = m.nodeCount, out->triangles = m.triCount;
		out->bytes = (uint32_t)(m.nodeCount * 80 + (size_t)m.triCount * 48);
		out->buildMs = m.buildMs, out->sahCost = m.sahCost;
	}
	API_END
}

} // extern "C"
