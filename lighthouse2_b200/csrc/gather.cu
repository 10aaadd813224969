/* gather.cu - the one exchange step of the sharded path, as our own collective over NVLink peer memory.

   One process per GPU renders its sample shard of a frame; the float4 accumulators have to be summed and finalized on
   rank 0. Instead of a library reduce (whose kernels compete with the persistent traversal kernels for SMs and spin while
   they wait for the slowest rank) the transfer uses the copy engines and the wait uses stream memory operations:

     rank r > 0, frame k:   [core stream]  snapshot accumulator -> snap[k&1]              (device-to-device, ~15 us at 1080p)
                            [comm stream]  wait( ack >= k-1 )                              rank 0 has consumed frame k-2 (same slot)
                                           copy snap[k&1] -> rank0.slot[r][k&1]            peer copy over NVLink, no SM involved
                                           set  rank0.arrived[r] = k+1                     32-bit memset into peer memory
     rank 0, frame k:       [core stream]  snapshot accumulator -> slot[0][k&1]
                            [comm stream]  wait( arrived[r] >= k+1 ) for every r           cuStreamWaitValue32: no SM spins
                                           sumFinalizeKernel: pixels = sum_r slot[r][k&1] / samples   (reduce + finalize fused)
                                           copy pixels -> pinned host (optional)
                                           set  rank r .ack = k+1 for every r              peers may reuse the slot
   Nothing blocks the host; frame k+1 renders on the core stream meanwhile. Buffers are shared between the processes
   with CUDA IPC handles, exchanged by the caller (lighthouse2_b200/distributed.py uses torch.distributed for that).
   The driver entry points for the stream memory operations are resolved at run time (no link dependency on libcuda).
*/
#include "core.h"
#include "kernels.h"
#include <cuda.h>
#include <cstring>

namespace lh2b
{

typedef CUresult( *WaitValue32Fn )( CUstream, CUdeviceptr, cuuint32_t, unsigned int );
typedef CUresult( *MemsetD32AsyncFn )( CUdeviceptr, unsigned int, size_t, CUstream );

struct GatherHandles { cudaIpcMemHandle_t slots, arrived, ack; };	// slots / arrived are meaningful for rank 0, ack for every rank

#define GATHER_MAX_RANKS 16
struct PeerSlots { const float4* p[GATHER_MAX_RANKS]; };

__global__ void __launch_bounds__( 256 ) sumFinalizeKernel( const PeerSlots slots, const int world, float4* __restrict__ out, const int n, const float scale )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 a = slots.p[0][i];
	for (int r = 1; r < world; r++)
	{
		const float4 b = slots.p[r][i];
		a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
	}
	out[i] = make_float4( a.x * scale, a.y * scale, a.z * scale, a.w * scale );
}

} // namespace lh2b

using namespace lh2b;

struct lh2b_gather
{
	lh2b_core* core = nullptr;
	int rank = 0, world = 1;
	size_t pixels = 0;
	uint32_t frame = 0;
	bool snapshotInFrame = false;		// the frame enqueued last already wrote its snapshot (GatherSnapshotTarget)
	cudaStream_t comm = nullptr;
	cudaEvent_t snapReady = nullptr;		// core stream: the snapshot of this frame is complete
	cudaEvent_t slotFree[2] = { nullptr, nullptr };	// comm stream: the local buffer of that parity has been consumed (kernel on rank 0, push elsewhere)
	// local allocations
	float4* slots = nullptr;			// rank 0: [world][2][pixels]
	uint32_t* arrived = nullptr;		// rank 0: [world]
	uint32_t* ack = nullptr;			// every rank: [1]
	float4* snap = nullptr;				// rank > 0: [2][pixels]
	float4* image = nullptr;			// rank 0: [2][pixels]
	// peer mappings
	float4* rootSlots = nullptr;		// rank > 0: rank 0's slots
	uint32_t* rootArrived = nullptr;	// rank > 0: rank 0's arrived[]
	uint32_t* peerAck[GATHER_MAX_RANKS] = {};	// rank 0: every peer's ack
	WaitValue32Fn waitValue = nullptr;
	MemsetD32AsyncFn memsetD32 = nullptr;
};

float4* GatherSnapshotTarget( lh2b_gather* g, cudaStream_t coreStream )
{
	const uint32_t k = g->frame, slot = k & 1;
	float4* local = g->rank == 0 ? g->slots + (size_t)slot * g->pixels : g->snap + (size_t)slot * g->pixels;
	if (k >= 2) CUDA_CHECK( cudaStreamWaitEvent( coreStream, g->slotFree[slot], 0 ) );
	g->snapshotInFrame = true;
	return local;
}

#define API_BEGIN try {
#define API_END } catch (const std::exception& e) { SetLastError( e.what() ); return 1; } return 0;
#define CU_CHECK( call ) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { char b_[256]; snprintf( b_, sizeof( b_ ), "%s failed at %s:%d: CUresult %d", #call, __FILE__, __LINE__, (int)r_ ); throw lh2b::CoreError( b_ ); } } while (0)

extern "C" {

int lh2b_gather_create( lh2b_core* core, int rank, int world, lh2b_gather** out )
{
	API_BEGIN
	if (!core || !out) throw CoreError( "gather_create: null argument" );
	if (world < 1 || world > GATHER_MAX_RANKS || rank < 0 || rank >= world) throw CoreError( "gather_create: rank / world out of range (at most 16 ranks)" );
	if (core->width == 0) throw CoreError( "gather_create: SetTarget first" );
	CUDA_CHECK( cudaSetDevice( core->device ) );
	lh2b_gather* g = new lh2b_gather();
	g->core = core, g->rank = rank, g->world = world, g->pixels = (size_t)core->width * core->height;
	CUDA_CHECK( cudaStreamCreateWithFlags( &g->comm, cudaStreamNonBlocking ) );
	CUDA_CHECK( cudaEventCreateWithFlags( &g->snapReady, cudaEventDisableTiming ) );
	for (int i = 0; i < 2; i++) CUDA_CHECK( cudaEventCreateWithFlags( &g->slotFree[i], cudaEventDisableTiming ) );
	cudaDriverEntryPointQueryResult q;
	void* fn = nullptr;
	CUDA_CHECK( cudaGetDriverEntryPoint( "cuStreamWaitValue32", &fn, cudaEnableDefault, &q ) );
	if (!fn || q != cudaDriverEntryPointSuccess) throw CoreError( "gather_create: cuStreamWaitValue32 is not available" );
	g->waitValue = (WaitValue32Fn)fn;
	CUDA_CHECK( cudaGetDriverEntryPoint( "cuMemsetD32Async", &fn, cudaEnableDefault, &q ) );
	if (!fn || q != cudaDriverEntryPointSuccess) throw CoreError( "gather_create: cuMemsetD32Async is not available" );
	g->memsetD32 = (MemsetD32AsyncFn)fn;
	CUDA_CHECK( cudaMalloc( &g->ack, 256 ) );
	CUDA_CHECK( cudaMemset( g->ack, 0, 256 ) );
	if (rank == 0)
	{
		CUDA_CHECK( cudaMalloc( &g->slots, (size_t)world * 2 * g->pixels * sizeof( float4 ) ) );
		CUDA_CHECK( cudaMalloc( &g->arrived, 256 ) );
		CUDA_CHECK( cudaMemset( g->arrived, 0, 256 ) );
		CUDA_CHECK( cudaMalloc( &g->image, 2 * g->pixels * sizeof( float4 ) ) );
	}
	else CUDA_CHECK( cudaMalloc( &g->snap, 2 * g->pixels * sizeof( float4 ) ) );
	CUDA_CHECK( cudaDeviceSynchronize() );
	core->gather = g;
	*out = g;
	API_END
}

int lh2b_gather_export( lh2b_gather* g, void* handlesOut )
{
	API_BEGIN
	GatherHandles h;
	memset( &h, 0, sizeof( h ) );
	CUDA_CHECK( cudaIpcGetMemHandle( &h.ack, g->ack ) );
	if (g->rank == 0)
	{
		CUDA_CHECK( cudaIpcGetMemHandle( &h.slots, g->slots ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.arrived, g->arrived ) );
	}
	memcpy( handlesOut, &h, sizeof( h ) );
	API_END
}

int lh2b_gather_import( lh2b_gather* g, const void* handlesOfAllRanks )
{
	API_BEGIN
	const GatherHandles* h = (const GatherHandles*)handlesOfAllRanks;
	CUDA_CHECK( cudaSetDevice( g->core->device ) );
	if (g->rank == 0)
	{
		for (int r = 1; r < g->world; r++) CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->peerAck[r], h[r].ack, cudaIpcMemLazyEnablePeerAccess ) );
	}
	else
	{
		CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootSlots, h[0].slots, cudaIpcMemLazyEnablePeerAccess ) );
		CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootArrived, h[0].arrived, cudaIpcMemLazyEnablePeerAccess ) );
	}
	API_END
}

/* One frame: call after lh2b_render( ..., async ) of that frame on every rank. samplesTotal: samples accumulated over all ranks
   (the divisor of finalize). pinnedOut (rank 0, may be null): page-locked float4[w*h] receiving the image asynchronously. */
int lh2b_gather_frame( lh2b_gather* g, int samplesTotal, float* pinnedOut )
{
	API_BEGIN
	lh2b_core* core = g->core;
	CUDA_CHECK( cudaSetDevice( core->device ) );
	const uint32_t k = g->frame++, slot = k & 1;
	const size_t bytes = g->pixels * sizeof( float4 );
	float4* local = g->rank == 0 ? g->slots + (size_t)slot * g->pixels : g->snap + (size_t)slot * g->pixels;	// rank 0 owns slots[0][*]
	// the local buffer of this parity was last used by frame k-2: its consumer (sum kernel on rank 0, peer copy elsewhere) runs on
	// the comm stream, the snapshot on the core stream. Normally the frame's own last kernel has written the snapshot already.
	if (!g->snapshotInFrame)
	{
		if (k >= 2) CUDA_CHECK( cudaStreamWaitEvent( core->stream, g->slotFree[slot], 0 ) );
		CUDA_CHECK( cudaMemcpyAsync( local, core->accumulator.ptr, bytes, cudaMemcpyDeviceToDevice, core->stream ) );
	}
	g->snapshotInFrame = false;
	CUDA_CHECK( cudaEventRecord( g->snapReady, core->stream ) );
	CUDA_CHECK( cudaStreamWaitEvent( g->comm, g->snapReady, 0 ) );
	if (g->rank > 0)
	{
		if (k >= 2) CU_CHECK( g->waitValue( (CUstream)g->comm, (CUdeviceptr)g->ack, k - 1, CU_STREAM_WAIT_VALUE_GEQ ) );
		float4* dst = g->rootSlots + ((size_t)g->rank * 2 + slot) * g->pixels;
		CUDA_CHECK( cudaMemcpyAsync( dst, local, bytes, cudaMemcpyDeviceToDevice, g->comm ) );
		CUDA_CHECK( cudaEventRecord( g->slotFree[slot], g->comm ) );
		CU_CHECK( g->memsetD32( (CUdeviceptr)(g->rootArrived + g->rank), k + 1, 1, (CUstream)g->comm ) );
	}
	else
	{
		for (int r = 1; r < g->world; r++) CU_CHECK( g->waitValue( (CUstream)g->comm, (CUdeviceptr)(g->arrived + r), k + 1, CU_STREAM_WAIT_VALUE_GEQ ) );
		PeerSlots ps;
		for (int r = 0; r < g->world; r++) ps.p[r] = g->slots + ((size_t)r * 2 + slot) * g->pixels;
		float4* img = g->image + (size_t)slot * g->pixels;
		if (samplesTotal <= 0) throw CoreError( "gather_frame: samplesTotal must be positive" );
		sumFinalizeKernel<<<(unsigned)((g->pixels + 255) / 256), 256, 0, g->comm>>>( ps, g->world, img, (int)g->pixels, 1.0f / (float)samplesTotal );
		CUDA_CHECK( cudaGetLastError() );
		CUDA_CHECK( cudaEventRecord( g->slotFree[slot], g->comm ) );
		if (pinnedOut) CUDA_CHECK( cudaMemcpyAsync( pinnedOut, img, bytes, cudaMemcpyDeviceToHost, g->comm ) );
		for (int r = 1; r < g->world; r++) CU_CHECK( g->memsetD32( (CUdeviceptr)g->peerAck[r], k + 1, 1, (CUstream)g->comm ) );
	}
	API_END
}

/* Blocks until everything enqueued by lh2b_gather_frame on this rank has completed (rank 0: the images are in host memory). */
int lh2b_gather_wait( lh2b_gather* g )
{
	API_BEGIN
	CUDA_CHECK( cudaStreamSynchronize( g->comm ) );
	API_END
}

/* Makes 'stream' wait for the communication work enqueued so far (for device-side timing of whole frames). */
int lh2b_gather_join( lh2b_gather* g, void* stream )
{
	API_BEGIN
	cudaEvent_t ev = nullptr;
	CUDA_CHECK( cudaEventCreateWithFlags( &ev, cudaEventDisableTiming ) );
	CUDA_CHECK( cudaEventRecord( ev, g->comm ) );
	CUDA_CHECK( cudaStreamWaitEvent( (cudaStream_t)stream, ev, 0 ) );
	CUDA_CHECK( cudaEventDestroy( ev ) );
	API_END
}

int lh2b_gather_image_device_ptr( lh2b_gather* g, void** ptrOut )
{
	API_BEGIN
	if (g->rank != 0 || g->frame == 0) throw CoreError( "gather_image_device_ptr: rank 0 after the first frame only" );
	*ptrOut = g->image + (size_t)((g->frame - 1) & 1) * g->pixels;
	API_END
}

int lh2b_gather_destroy( lh2b_gather* g )
{
	API_BEGIN
	if (!g) return 0;
	cudaSetDevice( g->core->device );
	cudaStreamSynchronize( g->comm );
	if (g->core->gather == g) g->core->gather = nullptr;
	if (g->rank == 0) { for (int r = 1; r < g->world; r++) if (g->peerAck[r]) cudaIpcCloseMemHandle( g->peerAck[r] ); }
	else { if (g->rootSlots) cudaIpcCloseMemHandle( g->rootSlots ); if (g->rootArrived) cudaIpcCloseMemHandle( g->rootArrived ); }
	cudaFree( g->slots ), cudaFree( g->arrived ), cudaFree( g->ack ), cudaFree( g->snap ), cudaFree( g->image );
	cudaEventDestroy( g->snapReady ), cudaEventDestroy( g->slotFree[0] ), cudaEventDestroy( g->slotFree[1] ), cudaStreamDestroy( g->comm );
	delete g;
	API_END
}

int lh2b_gather_handle_bytes() { return (int)sizeof( GatherHandles ); }

} // extern "C"
