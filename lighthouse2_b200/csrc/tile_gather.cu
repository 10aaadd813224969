/* tile_gather.cu - tile (row-band) sharding of ONE frame over the GPUs of a box: the strong-scaling mode of SURVEY.md 8e
   (partitioning 2) for BASELINE.json configs[4], "4K 1 spp + SVGF filter, real-time frame-time mode at 1/2/4/8 B200".

   Sample sharding (gather.cu) cannot shorten a 1-spp frame. Here rank r path-traces rows [y_r, y_(r+1)) of the frame
   (lh2b_set_row_band: same path indices, seeds and buffers as the whole frame, so every row is bit-identical to what a single
   GPU produces) and the rows are gathered on rank 0, which runs the tail of the frame - the SVGF / TAA chain with its temporal
   history, or the plain finalize - on the complete buffers. A row band is one contiguous range of every per-pixel buffer, so
   the exchange is a handful of peer copies per rank, straight into rank 0's own frame buffers:

   Layout of the bands: rank 0 takes rows [0, rootRows) (Setting "tileRootShare" x an equal share: it also runs the tail), the other
   ranks interleave the remaining 4-row tile rows (rank r: tile rows rootRows/4 + (r-1) + j * (world-1)), because the cost of a
   row varies over the image (sky vs. near geometry) and contiguous bands left the slowest peer to set the frame time. An
   interleaved band is a 2D range (tile row x pitch) of every buffer: still one (2D) copy per buffer.

     rank r > 0, frame k:  [core stream]  render its tile rows into the local buffers (tail deferred)
                           [comm stream]  wait( ack >= k-1 )                        rank 0 has finished the tail of frame k-2
                                          copy rows of: accumulator (direct), and in filter mode accumulator (indirect),
                                          deltaDepth -> the same rows of rank 0's buffer SET k & 1 (r2: rank 0 swaps two sets of these
                                          buffers every frame, so frame k arrives while the tail of frame k-1 runs);
                                          features, worldPos -> rank 0's staging set k & 1
                                          set rank0.arrived[r] = k+1
     rank 0, frame k:      [core stream]  render its own rows
                                          wait( arrived[r] >= k+1 ) for every r     cuStreamWaitValue32 on the core's own stream
                                          mergeFeaturesKernel (filter mode): staged feature rows -> features, keeping the history
                                          counter bits that rank 0's prepare pass owns; staged world positions -> current buffer
                                          tail: filter chain / finalize -> pixels;  set rank r .ack = k+1 for every r
   80 B per pixel cross NVLink in filter mode (16 B without the filter): 4K -> 663 MB x (N-1)/N per frame into rank 0. The result
   equals the single-GPU frame bit for bit at 1 spp (one path per pixel: no accumulation-order freedom), which is what
   tests/multigpu_worker.py checks. Buffers are shared through CUDA IPC handles like gather.cu's; create the gatherer after
   SetTarget and after Setting( "filter" ) (the handles name the buffers those calls allocate).
*/
#include "core.h"
#include "kernels.h"
#include <cuda.h>
#include <cstring>
#include <algorithm>

namespace lh2b
{

void EnsureFilterBuffersForSharing( lh2b_core* core );	// render.cu
void RunDeferredTail( lh2b_core* core );				// render.cu

typedef CUresult( *TgWaitValue32Fn )( CUstream, CUdeviceptr, cuuint32_t, unsigned int );
typedef CUresult( *TgMemsetD32AsyncFn )( CUdeviceptr, unsigned int, size_t, CUstream );

#define TILE_MAX_RANKS 16
struct TileHandles
{
	cudaIpcMemHandle_t accumulator[2], featStage[2], worldPosStage[2], deltaDepth[2], arrived, ack;	// all but 'ack' are meaningful for rank 0 only; [k & 1]: the set of frame k
	int filter, flip0, pad[2];																// rank 0: filter mode and the worldPos buffer index of its next frame
};

__global__ void __launch_bounds__( 256 ) mergeFeaturesKernel( uint4* __restrict__ features, const uint4* __restrict__ staged, const int first, const int n )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint4 s = staged[first + i];
	const uint32_t history = features[first + i].w & 15u;
	features[first + i] = make_uint4( s.x, s.y, s.z, (s.w & ~15u) | history );
}

} // namespace lh2b

using namespace lh2b;

struct lh2b_tile_gather
{
	lh2b_core* core = nullptr;
	int rank = 0, world = 1, filter = 0, flip0 = 0;
	size_t pixels = 0;
	int rootRows = 0;							// rank 0 renders rows [0, rootRows); multiples of 4 rows: the generate kernel works on 8x4-pixel tiles
	int bandY0 = 0, bandY1 = 0, bandStep = 1, tileRows = 0;	// this rank: tile rows bandY0/4 + j * bandStep below bandY1 (tileRows of them)
	uint32_t frame = 0;
	cudaStream_t comm = nullptr;
	cudaEvent_t rendered = nullptr;
	uint32_t* arrived = nullptr;				// rank 0: [world]
	uint32_t* ack = nullptr;					// every rank
	uint4* featStage[2] = { nullptr, nullptr };	// rank 0: staged feature rows of the peers, per frame parity
	float4* wpStage[2] = { nullptr, nullptr };	// rank 0: staged world positions of the peers (copied into the core's current buffer before the tail)
	// peer mappings (rank > 0: rank 0's buffers; rank 0: every peer's ack)
	float4* rootAccumulator[2] = { nullptr, nullptr }; uint4* rootFeatStage[2] = { nullptr, nullptr }; float4* rootWorldPos[2] = { nullptr, nullptr }; float4* rootDeltaDepth[2] = { nullptr, nullptr };
	uint32_t* rootArrived = nullptr;
	uint32_t* peerAck[TILE_MAX_RANKS] = {};
	TgWaitValue32Fn waitValue = nullptr;
	TgMemsetD32AsyncFn memsetD32 = nullptr;
};

#define API_BEGIN try {
#define API_END } catch (const std::exception& e) { SetLastError( e.what() ); return 1; } return 0;
#define CU_CHECK( call ) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { char b_[256]; snprintf( b_, sizeof( b_ ), "%s failed at %s:%d: CUresult %d", #call, __FILE__, __LINE__, (int)r_ ); throw lh2b::CoreError( b_ ); } } while (0)

extern "C" {

int lh2b_tile_handle_bytes() { return (int)sizeof( TileHandles ); }

/* Band layout (pure host arithmetic, no device needed): rank 0 renders rows [0, rootRows), rootRows = an equal share x rootShare,
   at least one tile row; rank r > 0 renders the 4-row tile rows rootRows/4 + (r-1) + j * (world-1) below the last row. Every row
   belongs to exactly one rank. Returns the band as (y0, y1, step) for lh2b_set_row_band_strided. */
int lh2b_tile_layout( int height, int world, float rootShare, int rank, int* y0, int* y1, int* stepTileRows )
{
	API_BEGIN
	if (world < 1 || world > TILE_MAX_RANKS || rank < 0 || rank >= world || height < 4 * world) throw CoreError( "tile_layout: arguments out of range" );
	if ((height & 3) && world > 2) throw CoreError( "tile_layout: interleaved bands need a frame height that is a multiple of 4" );
	const float share = rootShare < 0 ? 0 : (rootShare > 1 ? 1 : rootShare);
	const int rootRows = world == 1 ? height : std::max( 4, (int)((double)height / world * share) & ~3 );
	if (rank == 0) *y0 = 0, *y1 = rootRows, *stepTileRows = 1;
	else *y0 = rootRows + 4 * (rank - 1), *y1 = height, *stepTileRows = world - 1;
	API_END
}

int lh2b_tile_create( lh2b_core* core, int rank, int world, lh2b_tile_gather** out )
{
	API_BEGIN
	if (!core || !out) throw CoreError( "tile_create: null argument" );
	if (world < 1 || world > TILE_MAX_RANKS || rank < 0 || rank >= world) throw CoreError( "tile_create: rank / world out of range (at most 16 ranks)" );
	if (core->width == 0) throw CoreError( "tile_create: SetTarget first" );
	if (core->height < 4 * world) throw CoreError( "tile_create: fewer than 4 rows per rank" );
	CUDA_CHECK( cudaSetDevice( core->device ) );
	lh2b_tile_gather* g = new lh2b_tile_gather();
	g->core = core, g->rank = rank, g->world = world, g->pixels = (size_t)core->width * core->height;
	g->filter = core->filterEnabled ? 1 : 0;
	// rank 0 also runs the tail of every frame while the peers already render the next one: Setting "tileRootShare" (0..1, default 1)
	// scales its band relative to an equal share; the other ranks split the remaining rows evenly. Boundaries: multiples of 4 rows.
	int r0y0, r0y1, r0step;
	if (lh2b_tile_layout( core->height, world, core->tileRootShare, 0, &r0y0, &r0y1, &r0step ) != 0) throw CoreError( lh2b_last_error() );
	if (lh2b_tile_layout( core->height, world, core->tileRootShare, rank, &g->bandY0, &g->bandY1, &g->bandStep ) != 0) throw CoreError( lh2b_last_error() );
	g->rootRows = r0y1;
	g->tileRows = g->bandY1 > g->bandY0 ? (((g->bandY1 + 3) / 4 - g->bandY0 / 4) + g->bandStep - 1) / g->bandStep : 0;
	if (g->filter) EnsureFilterBuffersForSharing( core );
	g->flip0 = core->filterFlip;
	CUDA_CHECK( cudaStreamCreateWithFlags( &g->comm, cudaStreamNonBlocking ) );
	CUDA_CHECK( cudaEventCreateWithFlags( &g->rendered, cudaEventDisableTiming ) );
	cudaDriverEntryPointQueryResult q;
	void* fn = nullptr;
	CUDA_CHECK( cudaGetDriverEntryPoint( "cuStreamWaitValue32", &fn, cudaEnableDefault, &q ) );
	if (!fn || q != cudaDriverEntryPointSuccess) throw CoreError( "tile_create: cuStreamWaitValue32 is not available" );
	g->waitValue = (TgWaitValue32Fn)fn;
	CUDA_CHECK( cudaGetDriverEntryPoint( "cuMemsetD32Async", &fn, cudaEnableDefault, &q ) );
	if (!fn || q != cudaDriverEntryPointSuccess) throw CoreError( "tile_create: cuMemsetD32Async is not available" );
	g->memsetD32 = (TgMemsetD32AsyncFn)fn;
	CUDA_CHECK( cudaMalloc( &g->ack, 256 ) );
	CUDA_CHECK( cudaMemset( g->ack, 0, 256 ) );
	if (rank == 0)
	{
		CUDA_CHECK( cudaMalloc( &g->arrived, 256 ) );
		CUDA_CHECK( cudaMemset( g->arrived, 0, 256 ) );
		if (world > 1)
		{
			// double-buffered destinations of the peers' rows: the bands of frame k + 1 arrive while the tail of frame k runs
			core->accumulatorAlt.Resize( core->accumulator.count );
			CUDA_CHECK( cudaMemsetAsync( core->accumulatorAlt.ptr, 0, core->accumulatorAlt.count * sizeof( float4 ), core->stream ) );
			if (g->filter)
			{
				core->deltaDepthAlt.Resize( core->deltaDepth.count );
				for (int i = 0; i < 2; i++)
				{
					CUDA_CHECK( cudaMalloc( &g->featStage[i], g->pixels * sizeof( uint4 ) ) );
					CUDA_CHECK( cudaMalloc( &g->wpStage[i], g->pixels * sizeof( float4 ) ) );
				}
			}
			core->tileDouble = true, core->tileFrames = 0;
		}
	}
	// this core renders its band only; the tail of the frame runs in lh2b_tile_frame on rank 0
	const int rc = lh2b_set_row_band_strided( core, g->bandY0, g->bandY1, g->bandStep );
	if (rc != 0) throw CoreError( lh2b_last_error() );
	core->deferTail = true;
	CUDA_CHECK( cudaDeviceSynchronize() );
	*out = g;
	API_END
}

int lh2b_tile_export( lh2b_tile_gather* g, void* handlesOut )
{
	API_BEGIN
	TileHandles h;
	memset( &h, 0, sizeof( h ) );
	CUDA_CHECK( cudaIpcGetMemHandle( &h.ack, g->ack ) );
	if (g->rank == 0)
	{
		lh2b_core* c = g->core;
		CUDA_CHECK( cudaIpcGetMemHandle( &h.arrived, g->arrived ) );
		if (g->world > 1)
		{
			if (c->tileFrames != 0) throw CoreError( "tile_export: export the handles before the first frame" );
			CUDA_CHECK( cudaIpcGetMemHandle( &h.accumulator[0], c->accumulator.ptr ) );	// set of the even frames (the first frame does not swap)
			CUDA_CHECK( cudaIpcGetMemHandle( &h.accumulator[1], c->accumulatorAlt.ptr ) );
			if (g->filter)
			{
				CUDA_CHECK( cudaIpcGetMemHandle( &h.deltaDepth[0], c->deltaDepth.ptr ) );
				CUDA_CHECK( cudaIpcGetMemHandle( &h.deltaDepth[1], c->deltaDepthAlt.ptr ) );
				for (int i = 0; i < 2; i++)
				{
					CUDA_CHECK( cudaIpcGetMemHandle( &h.featStage[i], g->featStage[i] ) );
					CUDA_CHECK( cudaIpcGetMemHandle( &h.worldPosStage[i], g->wpStage[i] ) );
				}
			}
		}
	}
	h.filter = g->filter, h.flip0 = g->flip0;
	memcpy( handlesOut, &h, sizeof( h ) );
	API_END
}

int lh2b_tile_import( lh2b_tile_gather* g, const void* handlesOfAllRanks )
{
	API_BEGIN
	const TileHandles* h = (const TileHandles*)handlesOfAllRanks;
	CUDA_CHECK( cudaSetDevice( g->core->device ) );
	for (int r = 0; r < g->world; r++) if (h[r].filter != g->filter) throw CoreError( "tile_import: ranks disagree on the filter setting" );
	if (g->rank == 0)
	{
		for (int r = 1; r < g->world; r++) CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->peerAck[r], h[r].ack, cudaIpcMemLazyEnablePeerAccess ) );
	}
	else
	{
		const unsigned f = cudaIpcMemLazyEnablePeerAccess;
		CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootArrived, h[0].arrived, f ) );
		for (int i = 0; i < 2; i++)
		{
			CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootAccumulator[i], h[0].accumulator[i], f ) );
			if (!g->filter) continue;
			CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootFeatStage[i], h[0].featStage[i], f ) );
			CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootWorldPos[i], h[0].worldPosStage[i], f ) );
			CUDA_CHECK( cudaIpcOpenMemHandle( (void**)&g->rootDeltaDepth[i], h[0].deltaDepth[i], f ) );
		}
		g->flip0 = h[0].flip0;
	}
	API_END
}

/* One frame: call after lh2b_render( ..., async = 1 ) of that frame on every rank. Rank 0 ends with the complete frame in its pixel
   buffer (lh2b_read_pixels / lh2b_read_pixels_async as usual). */
int lh2b_tile_frame( lh2b_tile_gather* g )
{
	API_BEGIN
	lh2b_core* core = g->core;
	CUDA_CHECK( cudaSetDevice( core->device ) );
	const uint32_t k = g->frame++;
	const size_t w = (size_t)core->width, first = (size_t)g->bandY0 * w;
	if (g->rank > 0)
	{
		CUstream cs = (CUstream)g->comm;
		CUDA_CHECK( cudaEventRecord( g->rendered, core->stream ) );
		CUDA_CHECK( cudaStreamWaitEvent( g->comm, g->rendered, 0 ) );
		const uint32_t set = k & 1;	// destinations are double-buffered on rank 0: frame k may arrive while the tail of frame k - 1 runs
		if (k >= 2) CU_CHECK( g->waitValue( cs, (CUdeviceptr)g->ack, k - 1, CU_STREAM_WAIT_VALUE_GEQ ) );	// rank 0 is done with frame k - 2 (the same set)
		// this rank's rows of one per-pixel buffer: tileRows chunks of 4 rows, bandStep tile rows apart (one contiguous range if bandStep = 1)
		auto push = [&]( void* dst, const void* src, size_t elem ) {
			const size_t chunk = 4 * w * elem, pitch = chunk * (size_t)g->bandStep;
			if (g->bandStep == 1) CUDA_CHECK( cudaMemcpyAsync( (char*)dst + first * elem, (const char*)src + first * elem, (size_t)(g->bandY1 - g->bandY0) * w * elem, cudaMemcpyDeviceToDevice, g->comm ) );
			else CUDA_CHECK( cudaMemcpy2DAsync( (char*)dst + first * elem, pitch, (const char*)src + first * elem, pitch, chunk, (size_t)g->tileRows, cudaMemcpyDeviceToDevice, g->comm ) ); };
		push( g->rootAccumulator[set], core->accumulator.ptr, 16 );
		if (g->filter)
		{
			push( g->rootAccumulator[set] + g->pixels, core->accumulator.ptr + g->pixels, 16 );	// indirect half
			push( g->rootFeatStage[set], core->features.ptr, 16 );
			push( g->rootWorldPos[set], core->worldPosBuf[core->filterFlip].ptr, 16 );	// staged: rank 0's two world-position buffers are both in use by the tail of the frame before
			push( g->rootDeltaDepth[set], core->deltaDepth.ptr, 16 );
		}
		CU_CHECK( g->memsetD32( (CUdeviceptr)(g->rootArrived + g->rank), k + 1, 1, cs ) );
		// the next frame of this rank overwrites the rows just pushed: it has to wait for the copies
		CUDA_CHECK( cudaEventRecord( g->rendered, g->comm ) );
		CUDA_CHECK( cudaStreamWaitEvent( core->stream, g->rendered, 0 ) );
	}
	else
	{
		CUstream cs = (CUstream)core->stream;
		for (int r = 1; r < g->world; r++) CU_CHECK( g->waitValue( cs, (CUdeviceptr)(g->arrived + r), k + 1, CU_STREAM_WAIT_VALUE_GEQ ) );
		if (g->filter && g->world > 1)
		{
			const int peerFirst = (int)((size_t)g->rootRows * w), peerCount = (int)(g->pixels - (size_t)peerFirst);
			mergeFeaturesKernel<<<(peerCount + 255) / 256, 256, 0, core->stream>>>( core->features.ptr, g->featStage[k & 1], peerFirst, peerCount );
			CUDA_CHECK( cudaGetLastError() );
			// the peers' world positions of this frame: from the staging set into the buffer the frame's tail reads as 'current'
			CUDA_CHECK( cudaMemcpyAsync( core->worldPosBuf[core->filterFlip].ptr + peerFirst, g->wpStage[k & 1] + peerFirst, (size_t)peerCount * sizeof( float4 ),
				cudaMemcpyDeviceToDevice, core->stream ) );
		}
		RunDeferredTail( core );
		for (int r = 1; r < g->world; r++) CU_CHECK( g->memsetD32( (CUdeviceptr)g->peerAck[r], k + 1, 1, cs ) );
	}
	API_END
}

int lh2b_tile_wait( lh2b_tile_gather* g )
{
	API_BEGIN
	CUDA_CHECK( cudaStreamSynchronize( g->comm ) );
	CUDA_CHECK( cudaStreamSynchronize( g->core->stream ) );
	API_END
}

/* this rank's band: tile rows y0/4 + j * step below row y1 */
int lh2b_tile_rows( lh2b_tile_gather* g, int* y0, int* y1, int* stepTileRows )
{
	API_BEGIN
	*y0 = g->bandY0, *y1 = g->bandY1, *stepTileRows = g->bandStep;
	API_END
}

int lh2b_tile_destroy( lh2b_tile_gather* g )
{
	API_BEGIN
	if (!g) return 0;
	cudaSetDevice( g->core->device );
	cudaStreamSynchronize( g->comm ), cudaStreamSynchronize( g->core->stream );
	g->core->deferTail = false;
	lh2b_set_row_band( g->core, 0, 0 );
	if (g->rank == 0) { for (int r = 1; r < g->world; r++) if (g->peerAck[r]) cudaIpcCloseMemHandle( g->peerAck[r] ); }
	else
	{
		void* maps[] = { g->rootAccumulator[0], g->rootAccumulator[1], g->rootFeatStage[0], g->rootFeatStage[1], g->rootWorldPos[0], g->rootWorldPos[1],
			g->rootDeltaDepth[0], g->rootDeltaDepth[1], g->rootArrived };
		for (void* m : maps) if (m) cudaIpcCloseMemHandle( m );
	}
	if (g->core->tileDouble)
	{
		g->core->tileDouble = false, g->core->tileFrames = 0;
		g->core->accumulatorAlt.Free(), g->core->deltaDepthAlt.Free();	// whichever set is not current
	}
	cudaFree( g->arrived ), cudaFree( g->ack );
	for (int i = 0; i < 2; i++) cudaFree( g->featStage[i] ), cudaFree( g->wpStage[i] );
	cudaEventDestroy( g->rendered ), cudaStreamDestroy( g->comm );
	delete g;
	API_END
}

} // extern "C"
