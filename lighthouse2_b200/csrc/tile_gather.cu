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

   SHARDED FILTER CHAIN (Setting "tileFilterShard" 1, filter mode, 2..8 ranks). With the tail on rank 0 the frame time stops at that
   rank's filter chain (3.3 ms at 4K) however many GPUs render. Here every rank also FILTERS a band: rank r owns the rows
   [r * B, (r+1) * B) (B = an equal share rounded up to 16 rows) of the image and of the four history buffers (world positions, moments,
   phase-1 output, TAA image). It runs every stage of the chain on its band plus 16 halo rows on either side (FilterShard, kernels.h),
   and the history lookups - reprojected, i.e. anywhere in the frame - load each row from the rank that owns it, through peer mappings,
   from inside the filter kernels (filter_kernels.cu, SHARD = true). Rendering is sharded independently of that: interleaved 4-row tile
   rows (Setting "tileInterleave" 1, balanced) or the filter band itself (0).

     rank r, frame k:  [core stream]  render its tile rows: accumulator / deltaDepth into set k & 1, features and world positions into staging set
                                      k & 1 (the next frame's rendering starts as soon as the chain and the pushes of frame k-1 are through: every
                                      buffer a frame writes exists twice, by frame parity)
                       [comm stream]  for every rank d whose band + halo (world positions: band + 64 rows) holds rows r rendered:
                                      wait( done[d] >= k-1 ) (d is through the chain of frame k-2, which read set k & 1); copy those rows of
                                      deltaDepth, features, world positions once the last shade pass is through, of the accumulator halves once
                                      the frame is; d.arrived[r] = k+1
                       [tail stream]  (high priority, next to the core stream) wait( arrived[s] >= k+1 ) for every sender s; wait( done[s] >= k )
                                      for EVERY s: all ranks are through the chain of frame k-1, so its outputs - this frame's history - are
                                      complete everywhere and nobody reads the history of frame k-2 any more, which this frame overwrites;
                                      merge the staged feature rows (history counter bits stay), staged world positions -> current buffer;
                                      the chain on band + halo; present: rank 0 into its pixel buffer, rank r > 0 with peer stores straight into
                                      rank 0's staging image k & 1 (after wait( outAck >= k-1 ));
                                      s.done[r] = k+1 for every s;  rank r > 0: rank0.outArrived[r] = k+1
     rank 0            [2nd comm stream]  wait( outArrived[s] >= k+1 ) for every s; staging image -> pixel buffer (the peers' bands); s.outAck = k+1;
                                      readers of the pixel buffer (lh2b_read_pixels*, lh2b_present_gl) wait for this copy
   A converging frame (camera at rest) stores no features: its staging set is not merged and 'features' stays what the last restarted frame left.
   Every value a pixel depends on is computed by the same instructions from the same inputs as on one GPU (halo rows are computed twice,
   identically), so the frame is bit-identical to the single-GPU frame at 1 spp - checked by tests/multigpu_worker.py.
*/
#include "core.h"
#include "kernels.h"
#include <cuda.h>
#include <cstring>
#include <algorithm>
#include <vector>

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
	int filter, flip0, shard, interleave;													// rank 0: filter mode and the worldPos buffer index of its next frame
	// sharded filter chain: every rank exports the sets above plus
	cudaIpcMemHandle_t hist[4][2], flags, outStage[2];										// history buffers [worldPos, moments, filtered, taa][flip]; the flag words; rank 0: staging images
};

// flag words of the sharded chain (uint32 each, in one allocation per rank)
enum { FLAG_ARRIVED = 0, FLAG_DONE = 16, FLAG_OUT_ARRIVED = 32, FLAG_OUT_ACK = 48, FLAG_WORDS = 64 };

#define WP_MARGIN 64	// rows of world positions a rank holds beyond its band (a multiple of 4, at least the 16 halo rows)
struct RowSpan { int tile0, count, step; };	// the 4-row tile rows tile0 + j * step, j < count
/* tile rows of the band (y0, y1, step) inside the rows [e0, e1) (e0, e1 multiples of 4) */
static RowSpan BandInside( const int y0, const int y1, const int step, const int e0, const int e1 )
{
	const int t0 = y0 / 4, tEnd = (std::min( y1, e1 ) + 3) / 4, tLo = e0 / 4;
	const int j0 = t0 >= tLo ? 0 : (tLo - t0 + step - 1) / step, first = t0 + j0 * step;
	return RowSpan{ first, first < tEnd ? (tEnd - first + step - 1) / step : 0, step };
}

/* staged feature rows (this rank's own - its shade pass writes into the staging set of the frame - and the peers') -> features, rows
   [rowFirst, rowEnd); the history-counter bits belong to the prepare pass of this rank and stay */
__global__ void __launch_bounds__( 256 ) mergeShardFeaturesKernel( uint4* __restrict__ features, const uint4* __restrict__ staged, const int w, const int rowFirst, const int rowEnd )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= (rowEnd - rowFirst) * w) return;
	const int idx = rowFirst * w + i;
	const uint4 s = staged[idx];
	features[idx] = make_uint4( s.x, s.y, s.z, (s.w & ~15u) | (features[idx].w & 15u) );
}

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
	// sharded filter chain
	int shard = 0, interleave = 0;
	FilterShard fs = {};								// band [presentFirst, presentEnd), band + halo [rowFirst, rowEnd)
	int bands[TILE_MAX_RANKS][3] = {}, exts[TILE_MAX_RANKS][2] = {};	// every rank's rendered band (y0, y1, step) and filter band + halo
	int wpExts[TILE_MAX_RANKS][2] = {};					// ... and the wider strip of world positions it receives (band + WP_MARGIN rows)
	uint32_t* flags = nullptr; uint32_t* peerFlags[TILE_MAX_RANKS] = {};
	float4* peerAccumulator[TILE_MAX_RANKS][2] = {}; uint4* peerFeatStage[TILE_MAX_RANKS][2] = {}; float4* peerWpStage[TILE_MAX_RANKS][2] = {}; float4* peerDeltaDepth[TILE_MAX_RANKS][2] = {};
	float4* phase2 = nullptr;
	float4* outStage[2] = { nullptr, nullptr }, * rootOutStage[2] = { nullptr, nullptr };
	cudaStream_t comm2 = nullptr, tail = nullptr;		// tail: the filter chain of frame k runs here, next to the path tracing of frame k + 1 on the core's stream
	cudaEvent_t chainDone[2] = { nullptr, nullptr }, pushed[2] = { nullptr, nullptr }, tailEvent = nullptr;	// per staging set
	std::vector<void*> maps;							// every peer mapping opened (closed in lh2b_tile_destroy)
	std::vector<cudaEvent_t> stageTiming;				// and the 7 stage boundaries of every chain
	std::vector<cudaEvent_t> timing;					// LH2B_TILE_TIMING=1: per frame {render done, inputs complete, chain done} on the core stream
};

#define API_BEGIN try {
#define API_END } catch (const std::exception& e) { SetLastError( e.what() ); return 1; } return 0;
#define CU_CHECK( call ) do { CUresult r_ = (call); if (r_ != CUDA_SUCCESS) { char b_[256]; snprintf( b_, sizeof( b_ ), "%s failed at %s:%d: CUresult %d", #call, __FILE__, __LINE__, (int)r_ ); throw lh2b::CoreError( b_ ); } } while (0)

/* ---- sharded filter chain ------------------------------------------------------------------------------------------------------ */
static void ShardLayout( const int height, const int world, const int interleave, const int rank, int* band, int* ext, int* rowsPerBandOut )
{
	const int rowsPerBand = (((height + world - 1) / world) + 15) / 16 * 16;
	const int fb0 = rank * rowsPerBand, fb1 = std::min( height, (rank + 1) * rowsPerBand );
	ext[0] = std::max( 0, fb0 - 16 ), ext[1] = std::min( height, fb1 + 16 );
	if (interleave) band[0] = 4 * rank, band[1] = height, band[2] = world; else band[0] = fb0, band[1] = fb1, band[2] = 1;
	if (rowsPerBandOut) *rowsPerBandOut = rowsPerBand;
}

static void ShardCreate( lh2b_tile_gather* g )
{
	lh2b_core* core = g->core;
	const int h = core->height, world = g->world;
	if (world > LH2B_MAX_SHARDS) throw CoreError( "tile_create: the sharded filter chain takes at most 8 ranks" );
	if (h & 3) throw CoreError( "tile_create: the sharded filter chain needs a frame height that is a multiple of 4" );
	int rowsPerBand = 0;
	for (int r = 0; r < world; r++)
	{
		ShardLayout( h, world, core->tileInterleave, r, g->bands[r], g->exts[r], &rowsPerBand );
		g->wpExts[r][0] = std::max( 0, r * rowsPerBand - WP_MARGIN ), g->wpExts[r][1] = std::min( h, (r + 1) * rowsPerBand + WP_MARGIN );
	}
	if ((world - 1) * rowsPerBand >= h || h / 4 < world) throw CoreError( "tile_create: frame too small for this many filter bands (16-row granularity)" );
	g->shard = 1, g->interleave = core->tileInterleave;
	g->bandY0 = g->bands[g->rank][0], g->bandY1 = g->bands[g->rank][1], g->bandStep = g->bands[g->rank][2];
	g->tileRows = BandInside( g->bandY0, g->bandY1, g->bandStep, 0, h ).count;
	g->rootRows = 0;
	EnsureFilterBuffersForSharing( core );
	g->flip0 = core->filterFlip;
	const size_t bytes = g->pixels * sizeof( float4 );
	CUDA_CHECK( cudaMalloc( &g->flags, FLAG_WORDS * sizeof( uint32_t ) ) );
	CUDA_CHECK( cudaMemset( g->flags, 0, FLAG_WORDS * sizeof( uint32_t ) ) );
	core->accumulatorAlt.Resize( core->accumulator.count );
	CUDA_CHECK( cudaMemset( core->accumulatorAlt.ptr, 0, core->accumulatorAlt.count * sizeof( float4 ) ) );
	core->deltaDepthAlt.Resize( core->deltaDepth.count );
	CUDA_CHECK( cudaMemset( core->deltaDepthAlt.ptr, 0, core->deltaDepthAlt.count * sizeof( float4 ) ) );
	for (int i = 0; i < 2; i++)
	{
		CUDA_CHECK( cudaMalloc( &g->featStage[i], bytes ) );
		CUDA_CHECK( cudaMalloc( &g->wpStage[i], bytes ) );
		CUDA_CHECK( cudaMemset( g->featStage[i], 0, bytes ) );
		CUDA_CHECK( cudaMemset( g->wpStage[i], 0, bytes ) );		// like the world-position buffers themselves (EnsureFilterBuffers)
		if (g->rank == 0)
		{
			CUDA_CHECK( cudaMalloc( &g->outStage[i], bytes ) );
			CUDA_CHECK( cudaMemset( g->outStage[i], 0, bytes ) );	// the present pass leaves the frame's border pixels untouched: zero, like the pixel buffer after SetTarget
		}
	}
	CUDA_CHECK( cudaMalloc( &g->phase2, bytes ) );
	CUDA_CHECK( cudaMemset( g->phase2, 0, bytes ) );
	CUDA_CHECK( cudaStreamCreateWithFlags( &g->comm2, cudaStreamNonBlocking ) );
	int prioLow = 0, prioHigh = 0;
	CUDA_CHECK( cudaDeviceGetStreamPriorityRange( &prioLow, &prioHigh ) );
	CUDA_CHECK( cudaStreamCreateWithPriority( &g->tail, cudaStreamNonBlocking, prioHigh ) );	// the chain gates the other ranks: its blocks go first when an SM frees up
	for (int i = 0; i < 2; i++)
	{
		CUDA_CHECK( cudaEventCreateWithFlags( &g->chainDone[i], cudaEventDisableTiming ) );
		CUDA_CHECK( cudaEventCreateWithFlags( &g->pushed[i], cudaEventDisableTiming ) );
	}
	CUDA_CHECK( cudaEventCreateWithFlags( &g->tailEvent, cudaEventDisableTiming ) );
	g->fs.world = world, g->fs.rowsPerBand = rowsPerBand;
	g->fs.rowFirst = g->exts[g->rank][0], g->fs.rowEnd = g->exts[g->rank][1];
	g->fs.presentFirst = g->rank * rowsPerBand, g->fs.presentEnd = std::min( h, (g->rank + 1) * rowsPerBand );
	g->fs.phase2Out = g->phase2, g->fs.worldPosMargin = WP_MARGIN;
	core->tileDouble = true, core->tileFrames = 0;
	core->filterShard = &g->fs;
	core->worldPosOverride = g->wpStage[0], core->featuresOverride = g->featStage[0];
	core->tailStream = getenv( "LH2B_SHARD_NO_OVERLAP" ) ? core->stream : g->tail;	// (the environment switch is for A/B measurements: chain behind the frame on one stream)
	core->shardTarget = nullptr;
	if (g->rank == 0) core->tailEvent = g->tailEvent;
}

static void ShardExport( lh2b_tile_gather* g, TileHandles& h )
{
	lh2b_core* c = g->core;
	if (c->tileFrames != 0) throw CoreError( "tile_export: export the handles before the first frame" );
	CUDA_CHECK( cudaIpcGetMemHandle( &h.accumulator[0], c->accumulator.ptr ) );
	CUDA_CHECK( cudaIpcGetMemHandle( &h.accumulator[1], c->accumulatorAlt.ptr ) );
	CUDA_CHECK( cudaIpcGetMemHandle( &h.deltaDepth[0], c->deltaDepth.ptr ) );
	CUDA_CHECK( cudaIpcGetMemHandle( &h.deltaDepth[1], c->deltaDepthAlt.ptr ) );
	for (int i = 0; i < 2; i++)
	{
		CUDA_CHECK( cudaIpcGetMemHandle( &h.featStage[i], g->featStage[i] ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.worldPosStage[i], g->wpStage[i] ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.hist[0][i], c->worldPosBuf[i].ptr ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.hist[1][i], c->momentsBuf[i].ptr ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.hist[2][i], c->filteredBuf[i].ptr ) );
		CUDA_CHECK( cudaIpcGetMemHandle( &h.hist[3][i], c->taaBuf[i].ptr ) );
		if (g->rank == 0) CUDA_CHECK( cudaIpcGetMemHandle( &h.outStage[i], g->outStage[i] ) );
	}
	CUDA_CHECK( cudaIpcGetMemHandle( &h.flags, g->flags ) );
	h.shard = 1, h.interleave = g->interleave;
}

static void ShardImport( lh2b_tile_gather* g, const TileHandles* h )
{
	lh2b_core* c = g->core;
	const unsigned f = cudaIpcMemLazyEnablePeerAccess;
	auto open = [&]( const cudaIpcMemHandle_t& handle ) { void* p = nullptr; CUDA_CHECK( cudaIpcOpenMemHandle( &p, handle, f ) ); g->maps.push_back( p ); return p; };
	for (int r = 0; r < g->world; r++)
	{
		if (h[r].shard != 1 || h[r].interleave != g->interleave || h[r].flip0 != g->flip0) throw CoreError( "tile_import: ranks disagree on the sharded-chain settings" );
		for (int i = 0; i < 2; i++)
		{
			if (r == g->rank)
			{
				c->shardHist[0][i][r] = c->worldPosBuf[i].ptr, c->shardHist[1][i][r] = c->momentsBuf[i].ptr;
				c->shardHist[2][i][r] = c->filteredBuf[i].ptr, c->shardHist[3][i][r] = c->taaBuf[i].ptr;
				continue;
			}
			g->peerAccumulator[r][i] = (float4*)open( h[r].accumulator[i] ), g->peerDeltaDepth[r][i] = (float4*)open( h[r].deltaDepth[i] );
			g->peerFeatStage[r][i] = (uint4*)open( h[r].featStage[i] ), g->peerWpStage[r][i] = (float4*)open( h[r].worldPosStage[i] );
			for (int kind = 0; kind < 4; kind++) c->shardHist[kind][i][r] = (const float4*)open( h[r].hist[kind][i] );
			if (r == 0) g->rootOutStage[i] = (float4*)open( h[0].outStage[i] );
		}
		if (r != g->rank) g->peerFlags[r] = (uint32_t*)open( h[r].flags );
	}
	// the filter kernels look a row's owner up in device memory
	CUDA_CHECK( cudaMalloc( (void**)&c->shardHistDev, sizeof( c->shardHist ) ) );
	CUDA_CHECK( cudaMemcpy( (void*)c->shardHistDev, c->shardHist, sizeof( c->shardHist ), cudaMemcpyHostToDevice ) );
}

static void ShardFrame( lh2b_tile_gather* g )
{
	lh2b_core* core = g->core;
	const uint32_t k = g->frame++, set = k & 1;
	const size_t w = (size_t)core->width;
	const int me = g->rank;
	cudaStream_t tail = core->tailStream;
	CUstream cs = (CUstream)g->comm, ts = (CUstream)tail;
	// a converging frame (camera at rest, samples added to the frame before) stores no features: shade writes them for the first sample only
	// (sampleIdx == 0), so the staging set of such a frame holds nothing new and 'features' stays what the last restarted frame merged
	const bool restarted = core->samplesTaken == core->spp;
	static const bool timed = getenv( "LH2B_TILE_TIMING" ) != nullptr;
	auto stamp = [&]( cudaStream_t st ) { if (!timed) return; cudaEvent_t e; cudaEventCreate( &e ); cudaEventRecord( e, st ); g->timing.push_back( e ); };
	// ---- comm stream: this rank's rendered rows go to every rank whose band + halo holds some of them
	stamp( core->stream );
	CUDA_CHECK( cudaEventRecord( g->rendered, core->stream ) );
	// two rounds: features, world positions and depth derivatives are final once the LAST shade pass is through (paths that have only met
	// specular surfaces so far still update them at deeper path lengths; its event in the frame's launch sequence, render.cu) - they travel
	// while the last connect pass runs; the accumulator halves follow when the frame is done
	for (int round = 0; round < 2; round++)
	{
		CUDA_CHECK( cudaStreamWaitEvent( g->comm, round == 0 ? core->events[5 * core->maxPathLength + 3] : g->rendered, 0 ) );
		for (int d = 0; d < g->world; d++)
		{
			if (d == me) continue;
			const RowSpan rows = BandInside( g->bandY0, g->bandY1, g->bandStep, g->exts[d][0], g->exts[d][1] );
			const RowSpan wpRows = BandInside( g->bandY0, g->bandY1, g->bandStep, g->wpExts[d][0], g->wpExts[d][1] );
			if (wpRows.count == 0) continue;	// (the wider strip: no rows there, no rows in band + halo either)
			if (round == 0 && k >= 2) CU_CHECK( g->waitValue( cs, (CUdeviceptr)(g->flags + FLAG_DONE + d), k - 1, CU_STREAM_WAIT_VALUE_GEQ ) );	// d is through the chain of frame k - 2 (the same set)
			auto push = [&]( void* dst, const void* src, const RowSpan& r ) {
				if (r.count == 0) return;
				const size_t first = (size_t)r.tile0 * 4 * w * 16, chunk = 4 * w * 16, pitch = chunk * (size_t)r.step;
				if (r.step == 1) CUDA_CHECK( cudaMemcpyAsync( (char*)dst + first, (const char*)src + first, chunk * r.count, cudaMemcpyDeviceToDevice, g->comm ) );
				else CUDA_CHECK( cudaMemcpy2DAsync( (char*)dst + first, pitch, (const char*)src + first, pitch, chunk, (size_t)r.count, cudaMemcpyDeviceToDevice, g->comm ) ); };
			if (round == 0)
			{
				push( g->peerDeltaDepth[d][set], core->deltaDepth.ptr, rows );
				if (restarted) push( g->peerFeatStage[d][set], g->featStage[set], rows );
				push( g->peerWpStage[d][set], g->wpStage[set], wpRows );
			}
			else
			{
				push( g->peerAccumulator[d][set], core->accumulator.ptr, rows );
				push( g->peerAccumulator[d][set] + g->pixels, core->accumulator.ptr + g->pixels, rows );
				CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[d] + FLAG_ARRIVED + me), k + 1, 1, cs ) );
			}
		}
	}
	CUDA_CHECK( cudaEventRecord( g->pushed[set], g->comm ) );
	// ---- tail stream: inputs complete, every rank through the previous chain, then this rank's band of the chain. The core's own stream is
	// free for the path tracing of the next frame meanwhile: everything a frame writes is double-buffered by frame parity.
	CUDA_CHECK( cudaStreamWaitEvent( tail, g->rendered, 0 ) );
	for (int s = 0; s < g->world; s++)
	{
		if (s == me) continue;
		if (BandInside( g->bands[s][0], g->bands[s][1], g->bands[s][2], g->wpExts[me][0], g->wpExts[me][1] ).count > 0)
			CU_CHECK( g->waitValue( ts, (CUdeviceptr)(g->flags + FLAG_ARRIVED + s), k + 1, CU_STREAM_WAIT_VALUE_GEQ ) );
		if (k >= 1) CU_CHECK( g->waitValue( ts, (CUdeviceptr)(g->flags + FLAG_DONE + s), k, CU_STREAM_WAIT_VALUE_GEQ ) );
	}
	stamp( tail );
	const int extRows = g->fs.rowEnd - g->fs.rowFirst, extPixels = extRows * (int)w;
	if (restarted) mergeShardFeaturesKernel<<<(extPixels + 255) / 256, 256, 0, tail>>>( core->features.ptr, g->featStage[set], (int)w, g->fs.rowFirst, g->fs.rowEnd );
	CUDA_CHECK( cudaGetLastError() );
	const size_t wpFirst = (size_t)g->wpExts[me][0] * w, wpPixels = (size_t)(g->wpExts[me][1] - g->wpExts[me][0]) * w;
	CUDA_CHECK( cudaMemcpyAsync( core->worldPosBuf[core->filterFlip].ptr + wpFirst, g->wpStage[set] + wpFirst, wpPixels * sizeof( float4 ), cudaMemcpyDeviceToDevice, tail ) );
	if (me > 0)
	{
		if (k >= 2) CU_CHECK( g->waitValue( ts, (CUdeviceptr)(g->flags + FLAG_OUT_ACK), k - 1, CU_STREAM_WAIT_VALUE_GEQ ) );	// rank 0 has copied frame k - 2 out of that staging image
		core->shardTarget = g->rootOutStage[set];
	}
	if (timed)
	{
		for (int i = 0; i < 7; i++) { cudaEvent_t e; cudaEventCreate( &e ); g->stageTiming.push_back( e ); }
		core->filterStageEvents = g->stageTiming.data() + g->stageTiming.size() - 7;
	}
	RunDeferredTail( core );
	core->filterStageEvents = nullptr;
	stamp( tail );
	for (int s = 0; s < g->world; s++) if (s != me) CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[s] + FLAG_DONE + me), k + 1, 1, ts ) );
	CUDA_CHECK( cudaEventRecord( g->chainDone[set], tail ) );
	if (me > 0) CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[0] + FLAG_OUT_ARRIVED + me), k + 1, 1, ts ) );
	else
	{
		// rank 0: the peers' bands, presented into the staging image by their own kernels, join this rank's band in the pixel buffer
		CUstream c2 = (CUstream)g->comm2;
		CUDA_CHECK( cudaStreamWaitEvent( g->comm2, g->chainDone[set], 0 ) );
		for (int s = 1; s < g->world; s++) CU_CHECK( g->waitValue( c2, (CUdeviceptr)(g->flags + FLAG_OUT_ARRIVED + s), k + 1, CU_STREAM_WAIT_VALUE_GEQ ) );
		const size_t first = (size_t)g->fs.presentEnd * w;
		CUDA_CHECK( cudaMemcpyAsync( core->pixels.ptr + first, g->outStage[set] + first, (g->pixels - first) * sizeof( float4 ), cudaMemcpyDeviceToDevice, g->comm2 ) );
		for (int s = 1; s < g->world; s++) CU_CHECK( g->memsetD32( (CUdeviceptr)(g->peerFlags[s] + FLAG_OUT_ACK), k + 1, 1, c2 ) );
		CUDA_CHECK( cudaEventRecord( g->tailEvent, g->comm2 ) );
	}
	// the next frame of this rank writes the OTHER staging set: it has to wait for the chain and the pushes of the frame before this one
	if (k >= 1)
	{
		CUDA_CHECK( cudaStreamWaitEvent( core->stream, g->chainDone[set ^ 1], 0 ) );
		CUDA_CHECK( cudaStreamWaitEvent( core->stream, g->pushed[set ^ 1], 0 ) );
	}
	core->worldPosOverride = g->wpStage[(k + 1) & 1], core->featuresOverride = g->featStage[(k + 1) & 1];
}

static void ShardDestroy( lh2b_tile_gather* g )
{
	lh2b_core* core = g->core;
	cudaStreamSynchronize( g->comm2 );
	if (g->timing.size() >= 6)
	{
		// averages over the second half of the frames: period (render done -> next render done), stall (render done -> inputs and peers ready), chain
		const size_t frames = g->timing.size() / 3, f0 = frames / 2;
		double render = 0, stall = 0, chain = 0;
		for (size_t f = f0; f < frames; f++)
		{
			float a = 0, b = 0, c = 0;
			cudaEventElapsedTime( &a, g->timing[3 * f - 3], g->timing[3 * f] ), cudaEventElapsedTime( &b, g->timing[3 * f], g->timing[3 * f + 1] ), cudaEventElapsedTime( &c, g->timing[3 * f + 1], g->timing[3 * f + 2] );
			render += a, stall += b, chain += c;
		}
		double stage[6] = {};
		for (size_t f = f0; f < frames && 7 * f + 6 < g->stageTiming.size(); f++)
			for (int i = 0; i < 6; i++) { float t = 0; cudaEventElapsedTime( &t, g->stageTiming[7 * f + i], g->stageTiming[7 * f + i + 1] ); stage[i] += t / (frames - f0); }
		fprintf( stderr, "[tile timing] rank %d: prepare %.3f, a-trous %.3f %.3f %.3f, TAA %.3f, present %.3f ms\n", g->rank, stage[0], stage[1], stage[2], stage[3], stage[4], stage[5] );
		for (cudaEvent_t e : g->stageTiming) cudaEventDestroy( e );
		fprintf( stderr, "[tile timing] rank %d: period %.3f ms, stall %.3f ms, merge + chain %.3f ms per frame (%zu frames)\n", g->rank, render / (frames - f0), stall / (frames - f0), chain / (frames - f0), frames - f0 );
		for (cudaEvent_t e : g->timing) cudaEventDestroy( e );
	}
	cudaStreamSynchronize( g->tail );
	core->filterShard = nullptr, core->worldPosOverride = nullptr, core->featuresOverride = nullptr, core->tailStream = nullptr, core->shardTarget = nullptr, core->tailEvent = nullptr;
	memset( core->shardHist, 0, sizeof( core->shardHist ) );
	if (core->shardHistDev) cudaFree( (void*)core->shardHistDev ), core->shardHistDev = nullptr;
	for (void* m : g->maps) cudaIpcCloseMemHandle( m );
	cudaFree( g->flags ), cudaFree( g->phase2 );
	for (int i = 0; i < 2; i++) cudaFree( g->outStage[i] );
	for (int i = 0; i < 2; i++) cudaEventDestroy( g->chainDone[i] ), cudaEventDestroy( g->pushed[i] );
	cudaEventDestroy( g->tailEvent ), cudaStreamDestroy( g->comm2 ), cudaStreamDestroy( g->tail );
}

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

/* Layout of the sharded filter chain (Setting "tileFilterShard"; pure host arithmetic, no device needed): the rows a rank renders
   (band: y0, y1, step in tile rows), the rows it filters and presents (filterBand), the rows every filter stage computes there
   (withHalo = filterBand +- 16, clipped) and the strip of world positions it receives (worldPosStrip = filterBand +- 64, clipped). */
int lh2b_tile_shard_layout( int height, int world, int interleave, int rank, int* band, int* filterBand, int* withHalo, int* worldPosStrip )
{
	API_BEGIN
	if (world < 2 || world > LH2B_MAX_SHARDS || rank < 0 || rank >= world || height < 4 * world || (height & 3)) throw CoreError( "tile_shard_layout: arguments out of range" );
	int rowsPerBand = 0;
	ShardLayout( height, world, interleave, rank, band, withHalo, &rowsPerBand );
	if ((world - 1) * rowsPerBand >= height) throw CoreError( "tile_shard_layout: frame too small for this many filter bands (16-row granularity)" );
	filterBand[0] = rank * rowsPerBand, filterBand[1] = std::min( height, (rank + 1) * rowsPerBand );
	worldPosStrip[0] = std::max( 0, filterBand[0] - WP_MARGIN ), worldPosStrip[1] = std::min( height, rank * rowsPerBand + rowsPerBand + WP_MARGIN );
	API_END
}

/* The 4-row tile rows of the rendered band (y0, y1, step) that lie inside the rows [e0, e1): first tile row and count (what one
   strided peer copy moves). */
int lh2b_tile_rows_inside( int y0, int y1, int stepTileRows, int e0, int e1, int* firstTileRow, int* count )
{
	API_BEGIN
	if (stepTileRows < 1 || (e0 & 3) || (y0 & 3)) throw CoreError( "tile_rows_inside: arguments out of range" );
	const RowSpan r = BandInside( y0, y1, stepTileRows, e0, e1 );
	*firstTileRow = r.tile0, *count = r.count;
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
	const bool shard = core->tileFilterShard && g->filter && world > 1;
	// rank 0 also runs the tail of every frame while the peers already render the next one: Setting "tileRootShare" (0..1, default 1)
	// scales its band relative to an equal share; the other ranks split the remaining rows evenly. Boundaries: multiples of 4 rows.
	int r0y0, r0y1, r0step;
	if (lh2b_tile_layout( core->height, world, core->tileRootShare, 0, &r0y0, &r0y1, &r0step ) != 0) throw CoreError( lh2b_last_error() );
	if (lh2b_tile_layout( core->height, world, core->tileRootShare, rank, &g->bandY0, &g->bandY1, &g->bandStep ) != 0) throw CoreError( lh2b_last_error() );
	g->rootRows = r0y1;
	g->tileRows = g->bandY1 > g->bandY0 ? (((g->bandY1 + 3) / 4 - g->bandY0 / 4) + g->bandStep - 1) / g->bandStep : 0;
	if (g->filter) EnsureFilterBuffersForSharing( core );
	g->flip0 = core->filterFlip;
	if (shard) ShardCreate( g );
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
	if (rank == 0 && !shard)
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
	if (g->shard) ShardExport( g, h );
	else if (g->rank == 0)
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
	if (g->shard) ShardImport( g, h );
	else if (g->rank == 0)
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
	if (g->shard) { ShardFrame( g ); return 0; }
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
	if (g->tail) CUDA_CHECK( cudaStreamSynchronize( g->tail ) );
	if (g->comm2) CUDA_CHECK( cudaStreamSynchronize( g->comm2 ) );
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
	if (g->shard) ShardDestroy( g );
	else if (g->rank == 0) { for (int r = 1; r < g->world; r++) if (g->peerAck[r]) cudaIpcCloseMemHandle( g->peerAck[r] ); }
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
