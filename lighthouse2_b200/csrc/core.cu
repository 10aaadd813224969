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

void GpuBuildMesh( lh2b_core* core, Mesh& mesh, const int refit );
void GpuBuildTlas( lh2b_core* core, const void* dInstIn, const int n, const uint32_t* dLinkedRoots );
void ReleaseGpuBuildScratch( lh2b_core* core );

/* Arena slots: bump allocation; a mesh that outgrows its slot gets a new one at the top (the old slot is not
   recycled - meshes that change triangle count every frame should be rare). Growing the arena drains the stream
   first: builds in flight write through the old pointer. */
template <typename T> static void GrowArena( lh2b_core* core, DevBuf<T>& buf, size_t need )
{
	if (need <= buf.capacity) { buf.count = need; return; }
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	const size_t newCap = need + need / 2;
	T* np = nullptr;
	CUDA_CHECK( cudaMalloc( &np, newCap * sizeof( T ) ) );
	if (buf.ptr && buf.capacity) CUDA_CHECK( cudaMemcpy( np, buf.ptr, buf.capacity * sizeof( T ), cudaMemcpyDeviceToDevice ) );
	if (buf.ptr) cudaFree( buf.ptr );
	buf.ptr = np, buf.capacity = newCap, buf.count = need;
}

static uint32_t ArenaAllocNodes( lh2b_core* core, uint32_t cap )
{
	const uint32_t off = core->arenaNodeTop;
	core->arenaNodeTop += cap;
	GrowArena( core, core->arenaNodes, (size_t)core->arenaNodeTop * CW_NODE_QUADS );
	return off;
}

static uint32_t ArenaAllocTris( lh2b_core* core, uint32_t cap )
{
	const uint32_t off = core->arenaTriTop;
	core->arenaTriTop += cap;
	GrowArena( core, core->arenaTris, (size_t)core->arenaTriTop * 3 );
	return off;
}

static void EnsureMeshSlots( lh2b_core* core, Mesh& mesh, uint32_t nodesNeeded, uint32_t trisNeeded )
{
	if (nodesNeeded > mesh.nodeCap) mesh.nodeCap = nodesNeeded + nodesNeeded / 8 + 4, mesh.nodeOff = ArenaAllocNodes( core, mesh.nodeCap );
	if (trisNeeded > mesh.triCap) mesh.triCap = trisNeeded + 4, mesh.triOff = ArenaAllocTris( core, mesh.triCap );
	if (mesh.devBounds.count == 0) mesh.devBounds.Resize( 2 ), mesh.devCounts.Resize( 4 );
}

/* Host builder (Setting "bvhBuilder" = 1): binned SAH on the CPU, uploaded into the arena. */
static void RebuildMeshHost( lh2b_core* core, Mesh& mesh )
{
	const double t0 = NowMs();
	if (mesh.hostVerts.size() != (size_t)mesh.triCount * 12)
	{
		// the mesh was uploaded while another builder was selected: fetch the positions back
		mesh.hostVerts.resize( (size_t)mesh.triCount * 12 );
		if (mesh.triCount > 0) CUDA_CHECK( cudaMemcpyAsync( mesh.hostVerts.data(), mesh.verts.ptr, (size_t)mesh.triCount * 48, cudaMemcpyDeviceToHost, core->stream ) );
		CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	}
	std::vector<Bvh2Node> bvh2;
	std::vector<uint32_t> primIdx;
	BuildBvh2SAH( mesh.hostVerts.data(), mesh.triCount, bvh2, primIdx );
	CwBvh probe;
	CollapseToCwBvh( bvh2, primIdx, mesh.hostVerts.data(), probe );	// sizes first: the slot must exist before indices can be absolute
	EnsureMeshSlots( core, mesh, (uint32_t)probe.nodes.size(), (uint32_t)std::max<size_t>( probe.tris.size(), 1 ) );
	for (auto& n : probe.nodes) n.w[4] += mesh.nodeOff, n.w[5] += mesh.triOff;
	if (probe.tris.empty()) probe.tris.push_back( CwTri{} );
	mesh.nodeCount = (uint32_t)probe.nodes.size(), mesh.taggedInst = 0, mesh.hasTopology = false;
	const float4 b[2] = { make_float4( probe.bounds.lo[0], probe.bounds.lo[1], probe.bounds.lo[2], 0 ), make_float4( probe.bounds.hi[0], probe.bounds.hi[1], probe.bounds.hi[2], 0 ) };
	CUDA_CHECK( cudaMemcpyAsync( core->arenaNodes.ptr + (size_t)mesh.nodeOff * CW_NODE_QUADS, probe.nodes.data(), probe.nodes.size() * sizeof( CwNode ), cudaMemcpyHostToDevice, core->stream ) );
	CUDA_CHECK( cudaMemcpyAsync( core->arenaTris.ptr + (size_t)mesh.triOff * 3, probe.tris.data(), probe.tris.size() * sizeof( CwTri ), cudaMemcpyHostToDevice, core->stream ) );
	CUDA_CHECK( cudaMemcpyAsync( mesh.devBounds.ptr, b, sizeof( b ), cudaMemcpyHostToDevice, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	mesh.buildMs = (float)(NowMs() - t0);
	mesh.dirty = false;
}

/* GPU builder (default): LBVH + collapse, or a refit when only the vertex positions changed. Returns false if the
   node slot overflowed (the caller grows it and retries). */
static void RebuildMeshGpu( lh2b_core* core, Mesh& mesh )
{
	const int refit = (core->bvhRefit && mesh.hasTopology && mesh.builtTriCount == mesh.triCount && mesh.triCount > 0) ? core->bvhRefit : 0;
	if (!mesh.evStart) { CUDA_CHECK( cudaEventCreate( &mesh.evStart ) ); CUDA_CHECK( cudaEventCreate( &mesh.evEnd ) ); }
	if (refit == 1)
	{
		// in-place refit: node and triangle counts cannot change, so nothing is read back and the host does not wait
		CUDA_CHECK( cudaEventRecord( mesh.evStart, core->stream ) );
		GpuBuildMesh( core, mesh, 1 );
		CUDA_CHECK( cudaEventRecord( mesh.evEnd, core->stream ) );
		CUDA_CHECK( cudaGetLastError() );
		mesh.timingPending = true, mesh.dirty = false;
		return;
	}
	uint32_t nodeGuess = std::max( mesh.nodeCap, (uint32_t)(mesh.triCount / 2 + 64) );
	for (int attempt = 0; attempt < 4; attempt++)
	{
		EnsureMeshSlots( core, mesh, nodeGuess, (uint32_t)std::max( mesh.triCount, 1 ) );
		CUDA_CHECK( cudaEventRecord( mesh.evStart, core->stream ) );
		GpuBuildMesh( core, mesh, refit );
		CUDA_CHECK( cudaEventRecord( mesh.evEnd, core->stream ) );
		uint32_t counts[4] = { 0, 0, 0, 0 };
		CUDA_CHECK( cudaMemcpyAsync( counts, mesh.devCounts.ptr, 16, cudaMemcpyDeviceToHost, core->stream ) );
		CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
		CUDA_CHECK( cudaGetLastError() );
		if (counts[2] == 0)
		{
			mesh.nodeCount = mesh.triCount ? counts[0] : 1;
			CUDA_CHECK( cudaEventElapsedTime( &mesh.buildMs, mesh.evStart, mesh.evEnd ) );
			mesh.timingPending = false;
			mesh.hasTopology = mesh.triCount > 0, mesh.builtTriCount = mesh.triCount, mesh.taggedInst = 0, mesh.dirty = false;
			return;
		}
		nodeGuess = mesh.nodeCap * 2;	// slot too small for this tree: take a bigger one and rebuild
	}
	throw CoreError( "GPU BVH build: node slot overflow persists" );
}

/* Keep the BVH in L2 across the streaming traffic of a frame (path state, accumulator, peer pushes on rank 0): a persisting
   access window over the node arena on the launch stream, backed by a set-aside of L2 (Setting "l2Persist", default on).
   Everything else the stream touches is marked streaming by the window's miss property. */
static void ApplyL2Policy( lh2b_core* core )
{
	const size_t nodeBytes = core->arenaNodes.count * sizeof( uint4 );
	if (core->l2Persist == core->l2Applied && core->l2Base == (void*)core->arenaNodes.ptr && core->l2Bytes == nodeBytes) return;
	cudaDeviceProp prop;
	CUDA_CHECK( cudaGetDeviceProperties( &prop, core->device ) );
	cudaStreamAttrValue attr = {};
	if (core->l2Persist && prop.persistingL2CacheMaxSize > 0 && nodeBytes > 0)
	{
		const size_t window = std::min( nodeBytes, (size_t)prop.accessPolicyMaxWindowSize );
		const size_t setAside = std::min( window, (size_t)prop.persistingL2CacheMaxSize );
		CUDA_CHECK( cudaDeviceSetLimit( cudaLimitPersistingL2CacheSize, setAside ) );
		attr.accessPolicyWindow.base_ptr = core->arenaNodes.ptr;
		attr.accessPolicyWindow.num_bytes = window;
		attr.accessPolicyWindow.hitRatio = window <= setAside ? 1.0f : (float)setAside / (float)window;
		attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
	}
	else attr.accessPolicyWindow.num_bytes = 0;
	CUDA_CHECK( cudaStreamSetAttribute( core->stream, cudaStreamAttributeAccessPolicyWindow, &attr ) );
	CUDA_CHECK( cudaStreamSetAttribute( core->connectStream, cudaStreamAttributeAccessPolicyWindow, &attr ) );
	core->l2Applied = core->l2Persist, core->l2Base = core->arenaNodes.ptr, core->l2Bytes = nodeBytes;
}

struct InstBuildHost { float xform[12]; const float4* bounds; uint64_t pad; };

void UpdateAccelerationStructures( lh2b_core* core )
{
	for (auto& m : core->meshes) if (m->dirty)
	{
		if (core->bvhBuilder == 1) RebuildMeshHost( core, *m ); else RebuildMeshGpu( core, *m );
	}
	CUDA_CHECK( cudaEventRecord( core->evA, core->stream ) );
	const int n = (int)core->instances.size();
	std::vector<InstTrav> trav( n > 0 ? n : 1 );
	std::vector<lh2abi::CoreInstanceDesc> desc( n > 0 ? n : 1 );
	std::vector<InstBuildHost> buildIn( n > 0 ? n : 1 );
	std::vector<uint32_t> linked( n > 0 ? n : 1, 0 );
	// flat scene: every instance has the identity transform and no mesh is instanced twice. Then no ray ever needs
	// transforming: the top level links copies of the BLAS roots as internal children and traversal is single-level;
	// the instance index of a hit comes from the triangle record.
	bool flat = n > 0;
	{
		std::vector<int> uses( core->meshes.size(), 0 );
		for (int i = 0; i < n; i++) if (!IsIdentity( core->instances[i].xform ) || ++uses[core->instances[i].mesh] > 1) flat = false;
	}
	for (int i = 0; i < n; i++)
	{
		const Instance& inst = core->instances[i];
		Mesh& mesh = *core->meshes[inst.mesh];
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
		memcpy( buildIn[i].xform, inst.xform, 48 ), buildIn[i].bounds = mesh.devBounds.ptr, buildIn[i].pad = 0;
		linked[i] = mesh.nodeOff;
		if (flat && mesh.taggedInst != i)
		{
			LaunchTagTriangles( core->arenaTris.ptr + (size_t)mesh.triOff * 3, mesh.triCount, (uint32_t)i, core->stream );
			mesh.taggedInst = i;
		}
	}
	// top level: slot for 2n+2 nodes (n internal at most, n linked copies, root), leaf ids for n instances
	const uint32_t tlasNeed = (uint32_t)(2 * n + 2);
	if (tlasNeed > core->tlasCap) core->tlasCap = tlasNeed + tlasNeed / 2, core->tlasOff = ArenaAllocNodes( core, core->tlasCap );
	core->tlasLeafIds.Reserve( (size_t)n + 8 ), core->tlasLeafIds.count = (size_t)n + 8;
	core->instTrav.Upload( trav.data(), trav.size(), core->stream );
	core->instDesc.Upload( desc.data(), desc.size(), core->stream );
	core->instBuildIn.Upload( (const uint8_t*)buildIn.data(), buildIn.size() * sizeof( InstBuildHost ), core->stream );
	core->linkedRoots.Upload( linked.data(), linked.size(), core->stream );
	if (n == 0) CUDA_CHECK( cudaMemsetAsync( core->arenaNodes.ptr + (size_t)core->tlasOff * CW_NODE_QUADS, 0, sizeof( CwNode ), core->stream ) );
	else GpuBuildTlas( core, core->instBuildIn.ptr, n, flat ? core->linkedRoots.ptr : nullptr );
	CUDA_CHECK( cudaEventRecord( core->evB, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );	// the host vectors above were the copy sources
	CUDA_CHECK( cudaGetLastError() );
	CUDA_CHECK( cudaEventElapsedTime( &core->tlasBuildMs, core->evA, core->evB ) );
	core->tlasNodeCount = 0;
	core->scene.nodes = core->arenaNodes.ptr, core->scene.tris = core->arenaTris.ptr;
	core->scene.tlasRoot = core->tlasOff;
	// flat scenes start at the top-level root (whose children are BLAS-root copies), or directly at the BLAS root
	core->scene.singleRoot = !flat ? 0 : (n == 1 ? core->meshes[core->instances[0].mesh]->nodeOff : core->tlasOff);
	core->scene.tlasLeafIds = core->tlasLeafIds.ptr;
	core->scene.instances = core->instTrav.ptr;
	core->scene.instanceCount = n;
	core->scene.singleIdentity = flat ? 1 : 0;
	core->scene.stats = core->traceStatsOn ? (TraceStats*)core->traceStats.ptr : nullptr;
	ApplyL2Policy( core );
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
		{
			// the launch stream gets the highest priority: copy / gather work on other streams only fills what the frame leaves idle
			int least = 0, greatest = 0;
			CUDA_CHECK( cudaDeviceGetStreamPriorityRange( &least, &greatest ) );
			CUDA_CHECK( cudaStreamCreateWithPriority( &core->stream, cudaStreamNonBlocking, greatest ) );
			CUDA_CHECK( cudaStreamCreateWithPriority( &core->connectStream, cudaStreamNonBlocking, greatest ) );
		}
		CUDA_CHECK( cudaStreamCreateWithFlags( &core->copyStream, cudaStreamNonBlocking ) );
		CUDA_CHECK( cudaEventCreateWithFlags( &core->frameDone, cudaEventDisableTiming ) );
		for (int k = 0; k < 2; k++) CUDA_CHECK( cudaEventCreateWithFlags( &core->copyDone[k], cudaEventDisableTiming ) );
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
	ReleaseGpuBuildScratch( core );
	cudaEventDestroy( core->evA ), cudaEventDestroy( core->evB );
	if (core->copyStream) cudaStreamSynchronize( core->copyStream ), cudaStreamDestroy( core->copyStream );
	if (core->connectStream) cudaStreamSynchronize( core->connectStream ), cudaStreamDestroy( core->connectStream );
	if (core->frameDone) cudaEventDestroy( core->frameDone );
	for (int k = 0; k < 2; k++) if (core->copyDone[k]) cudaEventDestroy( core->copyDone[k] );
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
	if (core->bvhBuilder == 1) mesh.hostVerts.assign( vertexData, vertexData + (size_t)vertexCount * 4 );
	mesh.verts.Upload( (const float4*)vertexData, (size_t)vertexCount, core->stream );
	if (triangles) mesh.coreTris.Upload( (const float4*)triangles, (size_t)triangleCount * 13, core->stream );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) ); // caller may free its arrays on return
	mesh.dirty = true;
	API_END
}

int lh2b_set_geometry_device( lh2b_core* core, int meshIdx, const void* dVertexData, int vertexCount, int triangleCount, const void* dTriangles )
{
	API_BEGIN
	if (meshIdx < 0 || meshIdx > (int)core->meshes.size()) throw CoreError( "SetGeometry: meshes must be introduced in sequential order" );
	if (vertexCount != triangleCount * 3) throw CoreError( "SetGeometry: vertexCount must be 3 * triangleCount" );
	if (meshIdx == (int)core->meshes.size()) core->meshes.emplace_back( new Mesh() );
	Mesh& mesh = *core->meshes[meshIdx];
	mesh.triCount = triangleCount;
	mesh.verts.Resize( (size_t)vertexCount );
	CUDA_CHECK( cudaMemcpyAsync( mesh.verts.ptr, dVertexData, (size_t)vertexCount * 16, cudaMemcpyDeviceToDevice, core->stream ) );
	if (dTriangles)
	{
		mesh.coreTris.Resize( (size_t)triangleCount * 13 );
		CUDA_CHECK( cudaMemcpyAsync( mesh.coreTris.ptr, dTriangles, (size_t)triangleCount * 208, cudaMemcpyDeviceToDevice, core->stream ) );
	}
	if (core->bvhBuilder == 1)
	{
		mesh.hostVerts.resize( (size_t)vertexCount * 4 );
		CUDA_CHECK( cudaMemcpyAsync( mesh.hostVerts.data(), mesh.verts.ptr, (size_t)vertexCount * 16, cudaMemcpyDeviceToHost, core->stream ) );
		CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	}
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );	// the caller may reuse its device buffers on return
	mesh.dirty = true;
	API_END
}

static Mesh& AnimatedMesh( lh2b_core* core, int meshIdx, const char* what )
{
	if (meshIdx < 0 || meshIdx >= (int)core->meshes.size()) throw CoreError( std::string( what ) + ": unknown mesh" );
	Mesh& mesh = *core->meshes[meshIdx];
	if (mesh.coreTris.count != (size_t)mesh.triCount * 13 || mesh.triCount == 0) throw CoreError( std::string( what ) + ": the mesh needs SetGeometry with triangle records first" );
	return mesh;
}

static void CaptureBindPose( lh2b_core* core, Mesh& mesh )
{
	const size_t n = (size_t)mesh.triCount * 3;
	mesh.bindVerts.Resize( n ), mesh.bindNormals.Resize( n );
	CUDA_CHECK( cudaMemcpyAsync( mesh.bindVerts.ptr, mesh.verts.ptr, n * 16, cudaMemcpyDeviceToDevice, core->stream ) );
	LaunchCaptureBindPose( mesh.coreTris.ptr, mesh.bindNormals.ptr, mesh.triCount, core->stream );
}

static void AfterPose( lh2b_core* core, Mesh& mesh )
{
	if (core->bvhBuilder == 1)
	{
		mesh.hostVerts.resize( (size_t)mesh.triCount * 12 );
		CUDA_CHECK( cudaMemcpyAsync( mesh.hostVerts.data(), mesh.verts.ptr, (size_t)mesh.triCount * 48, cudaMemcpyDeviceToHost, core->stream ) );
		CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	}
	CUDA_CHECK( cudaGetLastError() );
	mesh.dirty = true;	// same triangle count: FinalizeInstances refits
}

int lh2b_set_skin( lh2b_core* core, int meshIdx, const uint32_t* joints4, const float* weights4, int vertexCount )
{
	API_BEGIN
	Mesh& mesh = AnimatedMesh( core, meshIdx, "SetSkin" );
	if (vertexCount != mesh.triCount * 3) throw CoreError( "SetSkin: one joint quadruple and one weight quadruple per vertex (3 per triangle)" );
	mesh.skinJoints.Upload( (const uint4*)joints4, (size_t)vertexCount, core->stream );
	mesh.skinWeights.Upload( (const float4*)weights4, (size_t)vertexCount, core->stream );
	CaptureBindPose( core, mesh );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_set_pose( lh2b_core* core, int meshIdx, const float* jointMatrices16, int jointCount )
{
	API_BEGIN
	Mesh& mesh = AnimatedMesh( core, meshIdx, "SetPose" );
	if (mesh.skinJoints.count != (size_t)mesh.triCount * 3) throw CoreError( "SetPose: call SetSkin first" );
	if (jointCount <= 0) throw CoreError( "SetPose: no joints" );
	// the matrices go through a pageable copy owned by the mesh: cudaMemcpyAsync stages pageable memory before it returns, so the
	// caller's array is free on return and the host never waits for the device here
	mesh.hostJointMats.assign( jointMatrices16, jointMatrices16 + (size_t)jointCount * 16 );
	mesh.jointMats.Upload( (const float4*)mesh.hostJointMats.data(), (size_t)jointCount * 4, core->stream );
	LaunchSkin( mesh.bindVerts.ptr, mesh.bindNormals.ptr, mesh.skinJoints.ptr, mesh.skinWeights.ptr, mesh.jointMats.ptr, jointCount,
		mesh.verts.ptr, mesh.coreTris.ptr, mesh.triCount, core->stream );
	AfterPose( core, mesh );
	API_END
}

int lh2b_set_morph_targets( lh2b_core* core, int meshIdx, const float* deltas4, const float* normals4, int targetCount, int vertexCount )
{
	API_BEGIN
	Mesh& mesh = AnimatedMesh( core, meshIdx, "SetMorphTargets" );
	if (vertexCount != mesh.triCount * 3 || targetCount <= 0) throw CoreError( "SetMorphTargets: targetCount arrays of 3 * triangleCount float4 each" );
	mesh.morphDeltas.Upload( (const float4*)deltas4, (size_t)targetCount * vertexCount, core->stream );
	mesh.morphNormals.Upload( (const float4*)normals4, (size_t)targetCount * vertexCount, core->stream );
	mesh.morphTargets = targetCount;
	CaptureBindPose( core, mesh );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_set_morph_weights( lh2b_core* core, int meshIdx, const float* weights, int targetCount )
{
	API_BEGIN
	Mesh& mesh = AnimatedMesh( core, meshIdx, "SetMorphWeights" );
	if (mesh.morphTargets == 0 || targetCount != mesh.morphTargets) throw CoreError( "SetMorphWeights: one weight per target of SetMorphTargets" );
	mesh.hostJointMats.assign( weights, weights + targetCount );
	mesh.morphWeights.Upload( mesh.hostJointMats.data(), (size_t)targetCount, core->stream );
	LaunchMorph( mesh.bindVerts.ptr, mesh.bindNormals.ptr, mesh.morphDeltas.ptr, mesh.morphNormals.ptr, mesh.morphWeights.ptr, targetCount,
		mesh.verts.ptr, mesh.coreTris.ptr, mesh.triCount, core->stream );
	AfterPose( core, mesh );
	API_END
}

int lh2b_read_geometry( lh2b_core* core, int meshIdx, float* vertexDataOut, void* trianglesOut )
{
	API_BEGIN
	if (meshIdx < 0 || meshIdx >= (int)core->meshes.size()) throw CoreError( "ReadGeometry: unknown mesh" );
	Mesh& mesh = *core->meshes[meshIdx];
	if (vertexDataOut) CUDA_CHECK( cudaMemcpyAsync( vertexDataOut, mesh.verts.ptr, (size_t)mesh.triCount * 48, cudaMemcpyDeviceToHost, core->stream ) );
	if (trianglesOut && mesh.coreTris.count) CUDA_CHECK( cudaMemcpyAsync( trianglesOut, mesh.coreTris.ptr, (size_t)mesh.triCount * 208, cudaMemcpyDeviceToHost, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
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

int lh2b_trace_stats_enable( lh2b_core* core, int on )
{
	API_BEGIN
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	if (on && core->traceStats.count == 0)
	{
		core->traceStats.Resize( 16 );
		CUDA_CHECK( cudaMemsetAsync( core->traceStats.ptr, 0, 16 * sizeof( unsigned long long ), core->stream ) );
	}
	core->traceStatsOn = on != 0;
	core->scene.stats = core->traceStatsOn ? (TraceStats*)core->traceStats.ptr : nullptr;
	API_END
}

int lh2b_trace_stats_read( lh2b_core* core, unsigned long long* out9, int reset )
{
	API_BEGIN
	if (core->traceStats.count == 0) throw CoreError( "trace_stats_read: counters were never enabled" );
	CUDA_CHECK( cudaMemcpyAsync( out9, core->traceStats.ptr, 9 * sizeof( unsigned long long ), cudaMemcpyDeviceToHost, core->stream ) );
	if (reset) CUDA_CHECK( cudaMemsetAsync( core->traceStats.ptr, 0, 16 * sizeof( unsigned long long ), core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_get_bvh_stats( lh2b_core* core, int meshIdx, lh2b_bvh_stats* out )
{
	API_BEGIN
	memset( out, 0, sizeof( *out ) );
	if (meshIdx == -1)
	{
		out->nodes = (uint32_t)core->instances.size() + 1, out->triangles = (uint32_t)core->instances.size();
		out->bytes = (uint32_t)(core->tlasCap * sizeof( CwNode ) + core->tlasLeafIds.Bytes() + core->instTrav.Bytes());
		out->buildMs = core->tlasBuildMs;
	}
	else
	{
		if (meshIdx < 0 || meshIdx >= (int)core->meshes.size()) throw CoreError( "unknown mesh" );
		Mesh& m = *core->meshes[meshIdx];
		if (m.timingPending)
		{
			CUDA_CHECK( cudaEventSynchronize( m.evEnd ) );
			CUDA_CHECK( cudaEventElapsedTime( &m.buildMs, m.evStart, m.evEnd ) );
			m.timingPending = false;
		}
		out->nodes = m.nodeCount, out->triangles = m.triCount;
		out->bytes = (uint32_t)(m.nodeCount * sizeof( CwNode ) + (size_t)m.triCount * 48);
		out->buildMs = m.buildMs, out->sahCost = m.sahCost;
	}
	API_END
}

} // extern "C"
