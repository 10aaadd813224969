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
   That is mode 0 ("root gather"). At 8 GPUs rank 0 then receives 7 x 33 MB and sums 8 buffers per frame while it renders its own
   frame, and being the slowest rank it sets the step time. Mode 1 ("reduce-scatter", Setting "gatherMode" 1 before
   lh2b_gather_create) spreads that work: the image is cut into `world` slices; every rank pushes slice j of its snapshot to rank
   j (copy engines), sums the `world` copies of its own slice and writes the finalized slice straight into rank 0's image buffer
   with peer stores from the same kernel (sum + finalize + transfer fused); rank 0 only waits for the slice flags and copies the
   image to the host. Per rank and frame: (world-1)/world x 33 MB in, the same out, 1/world of the summation.
     every rank r, frame k:  [comm]  for j != r: wait( stageAck[j] >= k-1 ); copy snap[slice j] -> rank j .stage[k&1][r]; set rank j .arrived[r] = k+1
                                     wait( arrived[j] >= k+1 ) for every j != r;  r != 0: wait( imgAck >= k-1 )
                                     sliceSumFinalizeKernel: rank0.image[k&1][slice r] = sum_j stage[k&1][j] / samples
                                     set rank j .stageAck[r] = k+1 for every j != r;   r != 0: set rank0.imgArrived[r] = k+1
             rank 0:                 wait( imgArrived[j] >= k+1 ) for every j; copy image -> pinned host; set rank j .imgAck = k+1
   Nothing blocks the host; frame k+1 renders on the core stream meanwhile. Buffers are shared between the processes
   with CUDA IPC handles, exchanged by the caller (lighthouse2_b200/distributed.py uses torch.distributed for that).
   The driver entry points for the stream memory operations are resolved at run time (no link dependency on libcuda).
*/
#include "core.h"
#include "kernels.h"
#include <cuda.h>
#include <cstring>
#include <algorithm>

namespace lh2b
{

typedef CUresult( *WaitValue32Fn )( CUstream, CUdeviceptr, cuuint32_t, unsigned int );
typedef CUresult( *MemsetD32AsyncFn )( CUdeviceptr, unsigned int, size_t, CUstream );

struct GatherHandles { cudaIpcMemHandle_t slots, arrived, ack, stage, flags, image; };	// mode 0: slots / arrived (rank 0), ack; mode 1: stage, flags (every rank), image (rank 0)

// mode 1 flag words (uint32 indices into the per-rank flags allocation)
#define GF_ARRIVED 0		// [world] written by peer r: its slice for frame k has landed in my stage buffer (k+1)
#define GF_STAGEACK 16		// [world] written by peer j: it has consumed what I pushed for frame k (k+1)
#define GF_IMGARRIVED 32	// [world] rank 0 only, written by peer r: its finalized slice of frame k is in my image (k+1)
#define GF_IMGACK 48		// written by rank 0: image slot of frame k has been copied out (k+1)

#define GATHER_MAX_RANKS 16
struct PeerSlots { const float4* p[GATHER_MAX_RANKS]; };

__global__ void __launch_bounds__( 256 ) sumFinalizeKernel( const PeerSlots slots, const int world, float4* __restrict__ out, const int n, const float scale )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 a = __ldcs( slots.p[0] + i );	// streamed once: evict-first, the BVH of the frame rendering next to this kernel keeps the L2
	for (int r = 1; r < world; r++)
	{
		const float4 b = __ldcs( slots.p[r] + i );
		a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
	}
	__stcs( out + i, make_float4( a.x * scale, a.y * scale, a.z * scale, a.w * scale ) );
}

/* mode 1: this rank's slice. in[j] = the slice as rendered by rank j (own snapshot for j == rank, staged copies otherwise); out points
   into rank 0's image buffer (a peer mapping unless this is rank 0): the stores travel over NVLink. */
__global__ void __launch_bounds__( 256 ) sliceSumFinalizeKernel( const PeerSlots in, const int world, float4* __restrict__ out, const int n, const float scale )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	float4 a = __ldcs( in.p[0] + i );
	for (int r = 1; r < world; r++)
	{
		const float4 b = __ldcs( in.p[r] + i );
		a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
	}
	__stcs( out + i, make_float4( a.x * scale, a.y * scale, a.z * scale, a.w * scale ) );
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
	// mode 1 (reduce-scatter)
	int mode = 0;
	size_t slicePixels = 0;					// pixels per slice (the last slice may be shorter)
	float4* stage = nullptr;				// [2][world][slicePixels]: slices of my part of the image as rendered by every peer
	uint32_t* flags = nullptr;				// GF_* words
	float4* peerStage[GATHER_MAX_RANKS] = {};
	uint32_t* peerFlags[GATHER_MAX_RANKS] = {};
	float4* rootImage = nullptr;			// rank > 0: rank 0's image buffer
	WaitValue32Fn waitValue = nullptr;
	MemsetD32AsyncFn memsetD32 = nullptr;
};

float4* GatherSnapshotTarget( lh2b_gather* g, cudaStream_t coreStream )
{
	const uint32_t k = g->frame, slot = k & 1;
	float4* local = (g->rank == 0 && g->mode == 0) ? g->slots + (size_t)slot * g->pixels : g->snap + (size_t)slot * g->pixels;
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
	g->mode = core->gatherMode;
	if (g->mode == 1)
	{
		g->slicePixels = (g->pixels + world - 1) / world;
		CUDA_CHECK( cudaMalloc( &g->stage, (size_t)2 * world * g->slicePixels * sizeof( float4 ) ) );
		CUDA_CHECK( cudaMalloc( &g->flags, 256 ) );
		CUDA_CHECK( cudaMemset( g->flags, 0, 256 ) );
		CUDA_CHECK( cudaMalloc( &g->snap, 2 * g->pixels * sizeof( float4 ) ) );
		if (rank == 0) CUDA_CHECK( cudaMalloc( &g->image, 2 * g->pixels * sizeof( float4 ) ) );
	}
	else if (rank == 0)
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
	if (g->mode == 1)
	{
		CUDA_CHECK( cudaIpcGetMemHandle( &h.stage, g->stage ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.flags, g->flags ) );
		if (g->rank == 0) CUDA_CHECK( cudaIpcGetMemHandle( &h.image, g->image ) );
	}
	else if (g->rank == 0)
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
	if (g->mode == 1)
	{
		for (int r = 0; r < g->world; r++) if (r != g->rank)
		{
			CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->peerStage[r], h[r].stage, cudaIpcMemLazyEnablePeerAccess ) );
			CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->peerFlags[r], h[r].flags, cudaIpcMemLazyEnablePeerAccess ) );
		}
		if (g->rank != 0) CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootImage, h[0].image, cudaIpcMemLazyEnablePeerAccess ) );
	}
	else if (g->rank == 0)
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
	float4* local = (g->rank == 0 && g->mode == 0) ? g->slots + (size_t)slot * g->pixels : g->snap + (size_t)slot * g->pixels;	// mode 0: rank 0 owns slots[0][*]
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
	if (g->mode == 1)
	{
		if (samplesTotal <= 0) throw CoreError( "gather_frame: samplesTotal must be positive" );
		const int W = g->world, me = g->rank;
		const size_t sp = g->slicePixels;
		auto sliceLen = [&]( int j ) { const size_t b = (size_t)j * sp; return b >= g->pixels ? (size_t)0 : std::min( sp, g->pixels - b ); };
		CUstream cs = (CUstream)g->comm;
		// 1. scatter: slice j of my snapshot goes to rank j (its stage[slot][me]) once rank j has consumed what I sent for frame k-2
		for (int d = 1; d < W; d++)
		{
			const int j = (me + d) % W;	// staggered start so that the ranks do not all push to the same peer first
			if (sliceLen( j ) == 0) continue;
			if (k >= 2) CU_CHECK( g->waitValue( cs, (CUdeviceptr)(g->flags + GF_STAGEACK + j), k - 1, CU_STREAM_WAIT_VALUE_GEQ ) );
			float4* dst = g->peerStage[j] + ((size_t)slot * W + me) * sp;
			CUDA_CHECK( cudaMemcpyAsync( dst, local + (size_t)j * sp, sliceLen( j ) * sizeof( float4 ), cudaMemcpyDeviceToDevice, g->comm ) );
			CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[j] + GF_ARRIVED + me), k + 1, 1, cs ) );
		}
		// 2. reduce my slice once every peer's copy of it has arrived; the finalized slice goes straight into rank 0's image
		const size_t myLen = sliceLen( me );
		if (myLen > 0)
		{
			for (int j = 0; j < W; j++) if (j != me) CU_CHECK( g->waitValue( cs, (CUdeviceptr)(g->flags + GF_ARRIVED + j), k + 1, CU_STREAM_WAIT_VALUE_GEQ ) );
			if (me != 0 && k >= 2) CU_CHECK( g->waitValue( cs, (CUdeviceptr)(g->flags + GF_IMGACK), k - 1, CU_STREAM_WAIT_VALUE_GEQ ) );
			PeerSlots ps;
			for (int j = 0; j < W; j++) ps.p[j] = j == me ? local + (size_t)me * sp : g->stage + ((size_t)slot * W + j) * sp;
			float4* img = (me == 0 ? g->image : g->rootImage) + (size_t)slot * g->pixels + (size_t)me * sp;
			sliceSumFinalizeKernel<<<(unsigned)((myLen + 255) / 256), 256, 0, g->comm>>>( ps, W, img, (int)myLen, 1.0f / (float)samplesTotal );
			CUDA_CHECK( cudaGetLastError() );
		}
		CUDA_CHECK( cudaEventRecord( g->slotFree[slot], g->comm ) );	// my snapshot buffer of this parity has been read (copies + kernel)
		for (int j = 0; j < W; j++) if (j != me && myLen > 0) CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[j] + GF_STAGEACK + me), k + 1, 1, cs ) );
		if (me != 0) CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[0] + GF_IMGARRIVED + me), k + 1, 1, cs ) );
		else
		{
			// 3. rank 0: the image is complete when every slice flag is up
			for (int j = 1; j < W; j++) CU_CHECK( g->waitValue( cs, (CUdeviceptr)(g->flags + GF_IMGARRIVED + j), k + 1, CU_STREAM_WAIT_VALUE_GEQ ) );
			if (pinnedOut) CUDA_CHECK( cudaMemcpyAsync( pinnedOut, g->image + (size_t)slot * g->pixels, bytes, cudaMemcpyDeviceToHost, g->comm ) );
			for (int j = 1; j < W; j++) CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[j] + GF_IMGACK), k + 1, 1, cs ) );
		}
	}
	else if (g->rank > 0)
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
	if (g->mode == 1)
	{
		for (int r = 0; r < g->world; r++) { if (g->peerStage[r]) cudaIpcCloseMemHandle( g->peerStage[r] ); if (g->peerFlags[r]) cudaIpcCloseMemHandle( g->peerFlags[r] ); }
		if (g->rootImage) cudaIpcCloseMemHandle( g->rootImage );
	}
	else if (g->rank == 0) { for (int r = 1; r < g->world; r++) if (g->peerAck[r]) cudaIpcCloseMemHandle( g->peerAck[r] ); }
	else { if (g->rootSlots) cudaIpcCloseMemHandle( g->rootSlots ); if (g->rootArrived) cudaIpcCloseMemHandle( g->rootArrived ); }
	cudaFree( g->slots ), cudaFree( g->arrived ), cudaFree( g->ack ), cudaFree( g->snap ), cudaFree( g->image ), cudaFree( g->stage ), cudaFree( g->flags );
	cudaEventDestroy( g->snapReady ), cudaEventDestroy( g->slotFree[0] ), cudaEventDestroy( g->slotFree[1] ), cudaStreamDestroy( g->comm );
	delete g;
	API_END
}

int lh2b_gather_handle_bytes() { return (int)sizeof( GatherHandles ); }

} // extern "C"
