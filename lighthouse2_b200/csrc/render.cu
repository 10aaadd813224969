/* render.cu - the render-target, scene-table and frame entry points of the C ABI, and the wavefront
   loop: generate+extend, shade, connect per path length, finalize. No host read-back happens between
   bounces (the reference blocks on a 48-byte counter copy per bounce, rendercore.cpp:903): every stage
   takes its ray count from device counters, so a frame is one fixed sequence of launches on one stream.

   Reference behaviour restated here: lib/rendercore_optix7/rendercore.cpp
     SetTarget :284-326 | SetTextures/SyncStorageType :438-502 | SetMaterials :508-565 | SetLights :698-710
     SetSkyData :716-740 | Setting :746-760 | Render/RenderImpl :819-938 | WaitForRender :945-957
     FinalizeRender :963-979 | GetCoreStats :1000-1003 | Init (blue noise expansion) :247-256
*/
#include "core.h"
#include "kernels.h"
#include <chrono>
#include <cmath>
#include <cstring>
#include <utility>
#include <cuda_fp16.h>

extern "C" const unsigned char lh2b_bluenoise_bytes[];

namespace lh2b
{

static double NowMs()
{
	return std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now().time_since_epoch() ).count();
}

static uint32_t RandomUInt( uint32_t& seed ) { seed ^= seed << 13, seed ^= seed >> 17, seed ^= seed << 5; return seed; }	// lib/platform/system.cpp:51

void InitRenderState( lh2b_core* core )
{
	// blue noise: three byte tables expanded to one uint per entry at offsets 0 / 65536 / 3*65536 (rendercore.cpp:247-254).
	// The sampler (tools_shared.h:324-337, like Heitz's published code) indexes the ranking tile with the UNWRAPPED dimension: from the
	// third path length on (dimensions 8..11) the four rank words of a pixel are those of the next tile pixel, and for tile pixel (127, 127)
	// they lie up to 12 words past the end of the reference's buffer (undefined there). Here those words exist and are zero - without the
	// padding the read returned whatever allocation followed the table: two core instances of one process disagreed in 1-4 pixels of a 4K
	// frame (tools/determinism_probe.py).
	std::vector<uint32_t> bn( 65536 * 5 + 16, 0 );
	const unsigned char* b = lh2b_bluenoise_bytes;
	for (int i = 0; i < 65536; i++) bn[i] = b[i];
	for (int i = 0; i < 128 * 128 * 8; i++) bn[i + 65536] = b[65536 + i];
	for (int i = 0; i < 128 * 128 * 8; i++) bn[i + 3 * 65536] = b[65536 + 131072 + i];
	core->blueNoise.Upload( bn.data(), bn.size(), core->stream );
	core->counters.Resize( 1 );
	{
		DevCounters init;
		memset( &init, 0, sizeof( init ) );
		init.probedTriid = -1;	// nothing probed yet
		CUDA_CHECK( cudaMemcpyAsync( core->counters.ptr, &init, sizeof( init ), cudaMemcpyHostToDevice, core->stream ) );
		CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	}
	CUDA_CHECK( cudaMallocHost( &core->hostCounters, sizeof( DevCounters ) ) );
	CUDA_CHECK( cudaMallocHost( &core->hostCountersB, sizeof( DevCounters ) ) );
	memset( core->hostCounters, 0, sizeof( DevCounters ) ), memset( core->hostCountersB, 0, sizeof( DevCounters ) );
	// events: 0 frame start, then per path length L (1-based): 5 * L + {0: trace start, 1: trace end, 2: shade start (connect( L - 1 ) has finished),
	// 3: shade end / connect start, 4: connect end - on the connect stream when the overlap is on}
	core->events.resize( 5 * (LH2B_MAXPATHLENGTH + 1) + 4 );
	for (auto& e : core->events) CUDA_CHECK( cudaEventCreate( &e ) );
	core->eventsB.resize( core->events.size() );
	for (auto& e : core->eventsB) CUDA_CHECK( cudaEventCreate( &e ) );
	// one-texel placeholders so table pointers are never null
	const uchar4 z4 = make_uchar4( 0, 0, 0, 0 );
	const float4 zf = make_float4( 0, 0, 0, 0 );
	core->argb32.Upload( &z4, 1, core->stream ), core->nrm32.Upload( &z4, 1, core->stream );
	core->argb128.Upload( &zf, 1, core->stream ), core->skyPixels.Upload( &zf, 1, core->stream );
	core->triLights.Upload( &zf, 1, core->stream ), core->pointLights.Upload( &zf, 1, core->stream );
	core->spotLights.Upload( &zf, 1, core->stream ), core->dirLights.Upload( &zf, 1, core->stream );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
}

void ReleaseRenderState( lh2b_core* core )
{
	if (core->glResource) cudaGraphicsUnregisterResource( core->glResource ), core->glResource = nullptr;
	for (auto& e : core->events) cudaEventDestroy( e );
	for (auto& e : core->eventsB) cudaEventDestroy( e );
	core->events.clear(), core->eventsB.clear();
	if (core->hostCounters) cudaFreeHost( core->hostCounters ), core->hostCounters = nullptr;
	if (core->hostCountersB) cudaFreeHost( core->hostCountersB ), core->hostCountersB = nullptr;
}

static void EnsureFilterBuffers( lh2b_core* core )
{
	const size_t px = core->maxPixels;
	if (core->features.count >= px) return;
	cudaStream_t s = core->stream;
	core->features.Resize( px ), core->deltaDepth.Resize( px ), core->shading.Resize( px ), core->motion.Resize( px );
	CUDA_CHECK( cudaMemsetAsync( core->features.ptr, 0, px * 16, s ) );
	for (int k = 0; k < 2; k++)
	{
		core->worldPosBuf[k].Resize( px ), core->momentsBuf[k].Resize( px ), core->filteredBuf[k].Resize( px ), core->taaBuf[k].Resize( px );
		CUDA_CHECK( cudaMemsetAsync( core->worldPosBuf[k].ptr, 0, px * 16, s ) );
		CUDA_CHECK( cudaMemsetAsync( core->momentsBuf[k].ptr, 0, px * 16, s ) );
		CUDA_CHECK( cudaMemsetAsync( core->filteredBuf[k].ptr, 0, px * 16, s ) );
		CUDA_CHECK( cudaMemsetAsync( core->taaBuf[k].ptr, 0, px * 16, s ) );
	}
	core->filterHistoryValid = false;
}

/* The SVGF / TAA tail of FinalizeRender (lib/RenderCore_Optix7Filter/rendercore.cpp:904-934), including its buffer rotation. */
static void RunFilter( lh2b_core* core )
{
	const int cur = core->filterFlip, prev = cur ^ 1;
	FilterBuffers b;
	b.accumulator = core->accumulator.ptr, b.features = core->features.ptr, b.worldPos = core->worldPosBuf[cur].ptr, b.prevWorldPos = core->worldPosBuf[prev].ptr;
	b.deltaDepth = core->deltaDepth.ptr, b.shading = core->shading.ptr, b.motion = core->motion.ptr;
	b.moments = core->momentsBuf[cur].ptr, b.prevMoments = core->momentsBuf[prev].ptr;
	b.filteredIN = core->filteredBuf[cur].ptr, b.filteredOUT = core->filteredBuf[prev].ptr;
	b.prevPixels = core->taaBuf[prev].ptr, b.taaOut = core->taaBuf[cur].ptr, b.target = core->pixels.ptr;
	FilterShard shard;
	if (core->filterShard)
	{
		// this rank filters its band of the frame; history rows come from the rank that owns them (all ranks flip in step)
		shard = *core->filterShard;
		auto table = [&]( int kind, int flip ) { return (const float4* const*)(core->shardHistDev + (kind * 2 + flip) * LH2B_MAX_SHARDS); };
		shard.prevWorldPos = table( 0, prev ), shard.prevMoments = table( 1, prev ), shard.filteredIN = table( 2, cur ), shard.prevPixels = table( 3, prev );
		b.shard = &shard;
		if (core->shardTarget) b.target = core->shardTarget;
	}
	FilterSettings fs;
	fs.w = core->width, fs.h = core->height, fs.samplesTaken = core->samplesTaken;
	fs.camIsStationary = core->samplesTaken == core->spp ? 0 : 1;
	fs.taa = core->taaEnabled ? 1 : 0, fs.directClamp = core->clampDirect, fs.indirectClamp = core->clampIndirect;
	fs.j0 = fs.j1 = fs.prevj0 = fs.prevj1 = 0;	// sub-pixel jitter comes from the blue-noise sampler of generate, not from the view
	memcpy( fs.prevView, core->filterHistoryValid ? &core->prevView : &core->lastView, sizeof( fs.prevView ) );
	cudaStream_t st = core->tailStream ? core->tailStream : core->stream;
	if (core->preciseMath) LaunchFilterChainPrecise( b, fs, st, core->filterStageEvents ); else LaunchFilterChain( b, fs, st, core->filterStageEvents );
	if (!core->taaEnabled)	// without TAA the history of the next frame's TAA pass is this frame's filtered image (swap( shading, prevPixels ))
	{
		const size_t first = core->filterShard ? (size_t)shard.rowFirst * core->width : 0, rows = core->filterShard ? shard.rowEnd - shard.rowFirst : core->height;
		CUDA_CHECK( cudaMemcpyAsync( core->taaBuf[cur].ptr + first, core->shading.ptr + first, rows * core->width * 16, cudaMemcpyDeviceToDevice, st ) );
	}
	// rotation: this frame's phase-1 output (in filteredOUT = filteredBuf[prev]) is the next frame's temporal history (filteredIN)
	core->filterFlip = prev;
	core->prevView = core->lastView, core->filterHistoryValid = true;
}

static uint32_t BandRowsOfCore( const lh2b_core* core );

static void FinishFrame( lh2b_core* core )
{
	if (!core->frameInFlight) return;
	// wait for THIS frame (its last event), not for the stream: with pipelining the next frame is already queued behind it
	CUDA_CHECK( cudaEventSynchronize( core->events.back() ) );
	core->frameInFlight = false;
	const DevCounters& c = *core->hostCounters;
	const int maxLen = core->maxPathLength;
	lh2abi::CoreStats& st = core->stats;
	lh2b_frame_stats& fs = core->frameStats;
	const uint32_t stride = (uint32_t)core->width * (core->bandY1 > core->bandY0 ? BandRowsOfCore( core ) : (uint32_t)core->height) * core->spp;
	st.primaryRayCount = stride;
	st.bounce1RayCount = maxLen >= 2 ? c.extensionRays[1] : 0;
	st.deepRayCount = 0;
	for (int L = 2; L < maxLen; L++) st.deepRayCount += c.extensionRays[L];
	st.totalExtensionRays = st.primaryRayCount + st.bounce1RayCount + st.deepRayCount;
	st.totalShadowRays = 0;
	for (int L = 1; L <= maxLen; L++) st.totalShadowRays += c.shadowRays[L];
	st.totalRays = st.totalExtensionRays + st.totalShadowRays;
	auto elapsed = [&]( int a, int b ) { float ms = 0; cudaEventElapsedTime( &ms, core->events[a], core->events[b] ); return ms; };
	st.traceTime0 = elapsed( 5, 6 );
	st.traceTime1 = maxLen >= 2 ? elapsed( 10, 11 ) : 0;
	st.traceTimeX = 0, st.shadeTime = 0, st.shadowTraceTime = 0;
	for (int L = 3; L <= maxLen; L++) st.traceTimeX += elapsed( 5 * L, 5 * L + 1 );
	// (with the connect overlap the stage times overlap too: connect( L ) runs next to extend( L + 1 ), their sum exceeds the frame time)
	for (int L = 1; L <= maxLen; L++) st.shadeTime += elapsed( 5 * L + 2, 5 * L + 3 ), st.shadowTraceTime += elapsed( 5 * L + 3, 5 * L + 4 );
	st.probedInstid = c.probedInstid, st.probedTriid = c.probedTriid, st.probedDist = c.probedDist;
	// probe world position (rendercore.cpp:935-937)
	{
		const lh2abi::ViewPyramid& v = core->lastView;
		const float rx = v.p2.x - v.p1.x, ry = v.p2.y - v.p1.y, rz = v.p2.z - v.p1.z;
		const float ux = v.p3.x - v.p1.x, uy = v.p3.y - v.p1.y, uz = v.p3.z - v.p1.z;
		const float fu = (core->probeX + 0.5f) / core->width, fv = (core->probeY + 0.5f) / core->height;
		float dx = v.p1.x + fu * rx + fv * ux - v.pos.x, dy = v.p1.y + fu * ry + fv * uy - v.pos.y, dz = v.p1.z + fu * rz + fv * uz - v.pos.z;
		const float il = 1.0f / sqrtf( dx * dx + dy * dy + dz * dz );
		st.probedWorldPos.x = v.pos.x + c.probedDist * dx * il, st.probedWorldPos.y = v.pos.y + c.probedDist * dy * il, st.probedWorldPos.z = v.pos.z + c.probedDist * dz * il;
	}
	const int finEv = (int)core->events.size() - 2;
	const double now = NowMs();
	st.renderTime = (float)((now - core->renderStartMs) * 0.001);	// the reference Timer reports seconds
	st.frameOverhead = core->lastFrameEndMs > 0 ? fmaxf( 0.0f, (float)((core->renderStartMs - core->lastFrameEndMs) * 0.001) ) : 0;
	core->lastFrameEndMs = now;
	fs.generateExtendMs = st.traceTime0, fs.extendMs = st.traceTime1 + st.traceTimeX, fs.shadeMs = st.shadeTime, fs.connectMs = st.shadowTraceTime;
	fs.finalizeMs = core->filterEnabled ? 0 : elapsed( finEv, finEv + 1 ), fs.filterMs = core->filterEnabled ? elapsed( finEv, finEv + 1 ) : 0;
	st.filterTime = fs.filterMs * 0.001f;
	fs.totalMs = elapsed( 0, finEv + 1 );
	fs.primaryRays = st.primaryRayCount, fs.extensionRays = st.totalExtensionRays, fs.shadowRays = st.totalShadowRays;
	fs.kernelLaunches = 3 * maxLen + 1;
	fs.pathLengthReached = 1;
	for (int L = 1; L < maxLen; L++) if (c.extensionRays[L] > 0) fs.pathLengthReached = L + 1;
	// traversal statistics are per second in the reference apps (times in seconds there); we keep milliseconds in frame stats
	st.traceTime0 *= 0.001f, st.traceTime1 *= 0.001f, st.traceTimeX *= 0.001f, st.shadeTime *= 0.001f, st.shadowTraceTime *= 0.001f;
}

static int BandY0( const lh2b_core* core ) { return core->bandY1 > core->bandY0 ? core->bandY0 : 0; }
static int BandY1( const lh2b_core* core ) { return core->bandY1 > core->bandY0 ? std::min( core->bandY1, core->height ) : core->height; }
static int BandStep( const lh2b_core* core ) { return core->bandY1 > core->bandY0 ? std::max( 1, core->bandStep ) : 1; }
static uint32_t BandTileRowsOf( const lh2b_core* core )
{
	const uint32_t first = (uint32_t)BandY0( core ) / 4, end = ((uint32_t)BandY1( core ) + 3) / 4, step = (uint32_t)BandStep( core );
	return end > first ? (end - first + step - 1) / step : 0;
}
static bool WholeFrame( const lh2b_core* core ) { return BandY0( core ) == 0 && BandY1( core ) == core->height && BandStep( core ) == 1; }
/* job slots of path length 1 (tile rows x 4 rows x width; rows outside the band are skipped by the kernels) */
static uint32_t BandSlots( const lh2b_core* core ) { return WholeFrame( core ) ? (uint32_t)core->width * core->height : BandTileRowsOf( core ) * 4u * core->width; }
/* rows this core really renders */
static uint32_t BandRows( const lh2b_core* core )
{
	uint32_t rows = 0;
	for (uint32_t j = 0, n = BandTileRowsOf( core ); j < n; j++)
	{
		const int t0 = ((BandY0( core ) / 4) + (int)j * BandStep( core )) * 4;
		rows += (uint32_t)std::max( 0, std::min( t0 + 4, BandY1( core ) ) - std::max( t0, BandY0( core ) ) );
	}
	return rows;
}
static uint32_t BandRowsOfCore( const lh2b_core* core ) { return BandRows( core ); }

static RenderParams BuildParams( lh2b_core* core, const lh2abi::ViewPyramid& view )
{
	const uint32_t stride = BandSlots( core ) * core->spp;
	RenderParams p = {};
	p.posLensSize = make_float4( view.pos.x, view.pos.y, view.pos.z, view.aperture );
	p.right = make_float3( view.p2.x - view.p1.x, view.p2.y - view.p1.y, view.p2.z - view.p1.z );
	p.up = make_float3( view.p3.x - view.p1.x, view.p3.y - view.p1.y, view.p3.z - view.p1.z );
	p.p1 = make_float3( view.p1.x, view.p1.y, view.p1.z );
	p.distortion = view.distortion, p.spreadAngle = view.spreadAngle;
	p.w = core->width, p.h = core->height, p.spp = core->spp;
	p.pass = core->samplesTaken, p.shift = core->shiftSeed;
	p.sampleBase = core->sampleShardTotal > 0 ? core->sampleShardFirst : 0;
	p.stride = stride, p.bandY0 = BandY0( core ), p.bandY1 = BandY1( core ), p.bandStep = BandStep( core );
	p.geometryEpsilon = core->geometryEpsilon, p.clampValue = core->clampValue;
	p.probePixelIdx = core->probeX + core->width * core->probeY;
	p.maxPathLength = core->maxPathLength, p.enoughBounces = core->enoughBounces, p.bsdfModel = core->bsdfModel;
	p.instDesc = core->instDesc.ptr, p.materials = core->materials.ptr;
	p.triLights = core->triLights.ptr, p.pointLights = core->pointLights.ptr, p.spotLights = core->spotLights.ptr, p.dirLights = core->dirLights.ptr;
	p.lightCounts = make_int4( core->lightCounts[0], core->lightCounts[1], core->lightCounts[2], core->lightCounts[3] );
	p.argb32 = core->argb32.ptr, p.argb128 = core->argb128.ptr, p.nrm32 = core->nrm32.ptr;
	p.skyPixels = core->skyPixels.ptr, p.skyW = core->skyW, p.skyH = core->skyH;
	memcpy( p.worldToSky, core->worldToSky, sizeof( p.worldToSky ) );
	p.blueNoise = core->blueNoise.ptr, p.accumulator = core->accumulator.ptr, p.counters = core->counters.ptr;
	if (core->filterEnabled && core->features.count) p.features = core->featuresOverride ? core->featuresOverride : core->features.ptr, p.worldPos = core->worldPosOverride ? core->worldPosOverride : core->worldPosBuf[core->filterFlip].ptr, p.deltaDepth = core->deltaDepth.ptr;
	return p;
}

/* While an asynchronous read-back of 'pixels' is in flight the next frame presents into the other buffer; if that one is still
   being copied too (two reads outstanding), the launch stream waits for that copy on the device - the host never blocks. */
static void RotatePixelBuffers( lh2b_core* core )
{
	if (!core->copyPending[0]) return;
	if (core->pixelsAlt.count < core->pixels.count)
	{
		core->pixelsAlt.Resize( core->pixels.count );
		CUDA_CHECK( cudaMemsetAsync( core->pixelsAlt.ptr, 0, core->pixelsAlt.count * sizeof( float4 ), core->stream ) );	// as SetTarget does for 'pixels' (the filter's present pass skips the border)
	}
	core->pixels.Swap( core->pixelsAlt );
	std::swap( core->copyDone[0], core->copyDone[1] );
	std::swap( core->copyPending[0], core->copyPending[1] );
	if (core->copyPending[0]) { CUDA_CHECK( cudaStreamWaitEvent( core->stream, core->copyDone[0], 0 ) ); core->copyPending[0] = false; }
}

/* Frame slots: A (the members used everywhere) holds the frame being enqueued / harvested, B a frame that stays in flight. */
static void SwapFrameSlots( lh2b_core* core )
{
	std::swap( core->hostCounters, core->hostCountersB ), core->events.swap( core->eventsB );
	std::swap( core->frameInFlight, core->frameInFlightB ), std::swap( core->renderStartMs, core->renderStartMsB );
	std::swap( core->lastView, core->lastViewB );
}

static void RenderFrame( lh2b_core* core, const lh2abi::ViewPyramid& view )
{
	cudaStream_t s = core->stream;
	const uint32_t stride = BandSlots( core ) * core->spp;
	core->lastView = view;
	CUDA_CHECK( cudaEventRecord( core->events[0], s ) );
	RotatePixelBuffers( core );
	if (core->tileDouble && core->tileFrames++ > 0)
	{
		// rank 0 of a tile-sharded frame: the peers push the rows of frame k into set k & 1 while the tail of frame k - 1 still reads the
		// other set. A converging frame continues the sums of this rank's own rows, which live in the set used last.
		core->accumulator.Swap( core->accumulatorAlt ), core->deltaDepth.Swap( core->deltaDepthAlt );
		if (core->samplesTaken != 0)
		{
			// (only this rank's own rows: the peers' rows of this frame may have arrived in the new set already)
			const size_t px = (size_t)core->width * core->height, first = (size_t)BandY0( core ) * core->width, rowBytes = (size_t)core->width * sizeof( float4 );
			auto own = [&]( float4* dst, const float4* src ) {
				if (BandStep( core ) == 1) CUDA_CHECK( cudaMemcpyAsync( dst + first, src + first, (size_t)(BandY1( core ) - BandY0( core )) * rowBytes, cudaMemcpyDeviceToDevice, s ) );
				else CUDA_CHECK( cudaMemcpy2DAsync( dst + first, (size_t)BandStep( core ) * 4 * rowBytes, src + first, (size_t)BandStep( core ) * 4 * rowBytes, 4 * rowBytes, BandTileRowsOf( core ), cudaMemcpyDeviceToDevice, s ) ); };
			own( core->accumulator.ptr, core->accumulatorAlt.ptr ), own( core->accumulator.ptr + px, core->accumulatorAlt.ptr + px );
			if (core->deltaDepth.count) own( core->deltaDepth.ptr, core->deltaDepthAlt.ptr );
		}
	}
	if (core->samplesTaken == 0)
	{
		// Restart: clear this core's rows of both accumulator halves (tile-sharded frames: the other rows belong to the peers' pushes)
		const size_t px = (size_t)core->width * core->height, rowBytes = (size_t)core->width * sizeof( float4 );
		if (BandStep( core ) == 1)
		{
			const size_t first = (size_t)BandY0( core ) * core->width, n = (size_t)(BandY1( core ) - BandY0( core )) * core->width;
			CUDA_CHECK( cudaMemsetAsync( core->accumulator.ptr + first, 0, n * sizeof( float4 ), s ) );
			CUDA_CHECK( cudaMemsetAsync( core->accumulator.ptr + px + first, 0, n * sizeof( float4 ), s ) );
		}
		else for (int half = 0; half < 2; half++)	// strided tile rows (lh2b_set_row_band_strided guarantees whole tile rows)
			CUDA_CHECK( cudaMemset2DAsync( core->accumulator.ptr + half * px + (size_t)BandY0( core ) * core->width, (size_t)BandStep( core ) * 4 * rowBytes, 0,
				4 * rowBytes, BandTileRowsOf( core ), s ) );
	}
	if (core->filterEnabled) EnsureFilterBuffers( core );
	CUDA_CHECK( cudaMemsetAsync( core->counters.ptr, 0, offsetof( DevCounters, probedInstid ), s ) );	// the probe words keep the last pick
	RandomUInt( core->shiftSeed );
	RenderParams p = BuildParams( core, view );
	const bool useNEE = (core->lightCounts[0] + core->lightCounts[1] + core->lightCounts[2] + core->lightCounts[3]) > 0;
	const PathSet conn = { core->connBuf[0].ptr, core->connBuf[1].ptr, core->connBuf[2].ptr };
	const int sm = (int)core->stats.SMcount;
	// connect( L ) only feeds the accumulator; extend( L + 1 ) only needs shade( L ): the two traversal launches run side by side
	// on two streams (the tail of one persistent grid is filled by the other - what counts for small frames and deep, thin path
	// lengths). shade( L + 1 ) waits for both, so deposits still reach the accumulator in the serial order: generate-time
	// determinism (bit-identical frames at 1 spp) is kept.
	const bool overlap = core->overlapConnect && useNEE && core->connectStream;
	cudaStream_t cs = overlap ? core->connectStream : s;
	for (int L = 1; L <= core->maxPathLength; L++)
	{
		const PathSet in = { core->pathBuf[(L - 1) & 1][0].ptr, core->pathBuf[(L - 1) & 1][1].ptr, core->pathBuf[(L - 1) & 1][2].ptr };
		const PathSet out = { core->pathBuf[L & 1][0].ptr, core->pathBuf[L & 1][1].ptr, core->pathBuf[L & 1][2].ptr };
		CUDA_CHECK( cudaEventRecord( core->events[5 * L], s ) );
		if (L == 1) LaunchGenerateExtend( core->scene, p, in, core->hitBuf.ptr, &core->counters.ptr->workFetch[2 * L], sm, s );
		else LaunchExtendCounted( core->scene, in, core->hitBuf.ptr, &core->counters.ptr->extensionRays[L - 1], &core->counters.ptr->workFetch[2 * L], stride, sm, s );
		CUDA_CHECK( cudaEventRecord( core->events[5 * L + 1], s ) );
		if (overlap && L > 1) CUDA_CHECK( cudaStreamWaitEvent( s, core->events[5 * (L - 1) + 4], 0 ) );	// connect( L - 1 ) is done
		CUDA_CHECK( cudaEventRecord( core->events[5 * L + 2], s ) );
		const uint32_t R0 = RandomUInt( core->camRNGseed ) + L * 91771;
		if (core->preciseMath) LaunchShadePrecise( p, in, out, core->hitBuf.ptr, conn, L, R0, useNEE, stride, sm, s );
		else LaunchShade( p, in, out, core->hitBuf.ptr, conn, L, R0, useNEE, stride, sm, s );
		CUDA_CHECK( cudaEventRecord( core->events[5 * L + 3], s ) );
		if (overlap) CUDA_CHECK( cudaStreamWaitEvent( cs, core->events[5 * L + 3], 0 ) );
		if (useNEE) LaunchConnect( core->scene, conn, core->accumulator.ptr, &core->counters.ptr->shadowRays[L], &core->counters.ptr->workFetch[2 * L + 1], stride, sm, cs );
		CUDA_CHECK( cudaEventRecord( core->events[5 * L + 4], cs ) );
	}
	if (overlap) CUDA_CHECK( cudaStreamWaitEvent( s, core->events[5 * core->maxPathLength + 4], 0 ) );
	CUDA_CHECK( cudaGetLastError() );
	CUDA_CHECK( cudaMemcpyAsync( core->hostCounters, core->counters.ptr, sizeof( DevCounters ), cudaMemcpyDeviceToHost, s ) );
	// FinalizeRender (rendercore.cpp:963-979), enqueued behind the last stage: nothing of it depends on host-visible results,
	// so the whole frame is one uninterrupted sequence of launches and WaitForRender is a single stream synchronisation
	const int total = core->sampleShardTotal > 0 ? core->sampleShardTotal : core->spp;
	core->samplesTaken += total;
	const int localSamples = (int)((long long)core->samplesTaken * core->spp / total);
	const int finEv = (int)core->events.size() - 2;
	CUDA_CHECK( cudaEventRecord( core->events[finEv], s ) );
	if (core->deferTail) { /* tile-sharded frame: lh2b_tile_frame enqueues the tail once the peers' rows are in place */ }
	else if (core->gather)	// sharded frame: rank 0 finalizes the sum of all shards (gather.cu); this rank's last kernel is the snapshot for it
		LaunchFinalize( core->accumulator.ptr, GatherSnapshotTarget( core->gather, s ), core->width * core->height, 1, s );
	else if (core->filterEnabled && core->features.count) RunFilter( core );
	else LaunchFinalize( core->accumulator.ptr, core->pixels.ptr, core->width * core->height, localSamples, s );
	CUDA_CHECK( cudaEventRecord( core->events[finEv + 1], s ) );
	core->frameInFlight = true;
}

void EnsureFilterBuffersForSharing( lh2b_core* core ) { EnsureFilterBuffers( core ); }

/* The tail of a frame whose finalize / filter was deferred (core->deferTail): called by the tile gatherer on rank 0, on the
   core's stream, after every peer's rows have landed in this core's buffers. */
void RunDeferredTail( lh2b_core* core )
{
	if (core->filterEnabled && core->features.count) RunFilter( core );
	else LaunchFinalize( core->accumulator.ptr, core->pixels.ptr, core->width * core->height, core->samplesTaken, core->stream );
}

} // namespace lh2b

using namespace lh2b;

#define API_BEGIN if (!core) { SetLastError( "null core handle" ); return 1; } try { CUDA_CHECK( cudaSetDevice( core->device ) );
#define API_END } catch (const std::exception& e) { SetLastError( e.what() ); return 1; } return 0;

static uint32_t ToChar( float a ) { return (uint32_t)(a * 255.0f); }	// truncation, as TOCHAR (rendercore.cpp:510)
static uint32_t Pack4( float a, float b, float c, float d ) { return ToChar( a ) + (ToChar( b ) << 8) + (ToChar( c ) << 16) + (ToChar( d ) << 24); }
static uint32_t HalfBits( float f ) { const __half h = __float2half_rn( f ); unsigned short u; memcpy( &u, &h, 2 ); return u; }

static void ReleaseGlTarget( lh2b_core* core )
{
	if (core->glResource) cudaGraphicsUnregisterResource( core->glResource );
	core->glResource = nullptr, core->glRegisteredTexture = 0;
}

template <typename V> static uint4 MakeMap( const lh2b_core* core, const V& v )
{
	// CUDAMaterial::Map { short width, height; half uscale, vscale, uoffs, voffs; uint addr } (core_settings.h:141, rendercore.h:78-85)
	if (v.textureID < 0 || v.textureID >= (int)core->texDescs.size()) throw CoreError( "SetMaterials: texture id out of range (call SetTextures first)" );
	const lh2abi::CoreTexDesc& t = core->texDescs[v.textureID];
	uint4 m;
	m.x = (t.width & 0xffff) | ((t.height & 0xffff) << 16);
	m.y = HalfBits( v.uvscale.x ) | (HalfBits( v.uvscale.y ) << 16);
	m.z = HalfBits( v.uvoffset.x ) | (HalfBits( v.uvoffset.y ) << 16);
	m.w = t.firstPixel;
	return m;
}

extern "C" {

int lh2b_set_target( lh2b_core* core, int width, int height, int spp )
{
	API_BEGIN
	if (width <= 0 || height <= 0 || spp <= 0) throw CoreError( "SetTarget: width, height and spp must be positive" );
	if ((unsigned long long)width * height * spp >= (1ull << 26)) throw CoreError( "SetTarget: w*h*spp must stay below 2^26 (path index is packed above 6 flag bits)" );
	if ((core->gather || core->deferTail) && (width != core->width || height != core->height || spp != core->spp))
		throw CoreError( "SetTarget: a multi-GPU gatherer is attached to this core (its peer buffers are sized for the current target); destroy it first" );
	FinishFrame( core );
	CUDA_CHECK( cudaStreamSynchronize( core->copyStream ) );
	core->copyPending[0] = core->copyPending[1] = false;
	if (core->glResource && (width != core->glRegisteredW || height != core->glRegisteredH)) ReleaseGlTarget( core );
	core->width = width, core->height = height, core->spp = spp;
	const size_t pixels = (size_t)width * height;
	if (pixels > core->maxPixels || spp != core->allocatedSpp)
	{
		core->maxPixels = pixels + (pixels >> 4);	// slack against frequent reallocation (rendercore.cpp:298-299)
		core->allocatedSpp = spp;
		const size_t rays = core->maxPixels * spp;
		for (int b = 0; b < 2; b++) for (int k = 0; k < 3; k++) core->pathBuf[b][k].Free(), core->pathBuf[b][k].Resize( rays );
		for (int k = 0; k < 3; k++) core->connBuf[k].Free(), core->connBuf[k].Resize( rays );
		core->hitBuf.Free(), core->hitBuf.Resize( rays );
		core->accumulator.Free(), core->accumulator.Resize( core->maxPixels * 2 );	// second half: indirect light in filter mode
		core->pixels.Free(), core->pixels.Resize( core->maxPixels ), core->pixelsAlt.Free();
	}
	CUDA_CHECK( cudaMemsetAsync( core->accumulator.ptr, 0, 2 * pixels * sizeof( float4 ), core->stream ) );
	CUDA_CHECK( cudaMemsetAsync( core->pixels.ptr, 0, pixels * sizeof( float4 ), core->stream ) );
	core->samplesTaken = 0, core->filterHistoryValid = false;
	API_END
}

int lh2b_setting( lh2b_core* core, const char* name, float value )
{
	API_BEGIN
	if (!name) throw CoreError( "Setting: null name" );
	if (!strcmp( name, "epsilon" )) core->geometryEpsilon = value;
	else if (!strcmp( name, "clampValue" )) core->clampValue = value;
	else if (!strcmp( name, "noiseShift" )) { /* accepted and unused, as in the reference (rendercore.cpp:756-759) */ }
	// the filter core's settings (lib/RenderCore_Optix7Filter/rendercore.cpp:656-678); RenderSystem sends them to every core
	else if (!strcmp( name, "filter" ))
	{
		const bool on = value > 0;
		if (on != core->filterEnabled && (core->gather || core->deferTail)) throw CoreError( "Setting filter: a multi-GPU gatherer is attached to this core; destroy it first" );
		if (on != core->filterEnabled) core->filterEnabled = on, core->filterHistoryValid = false, core->samplesTaken = 0;
	}
	else if (!strcmp( name, "TAA" )) core->taaEnabled = value > 0;
	else if (!strcmp( name, "clampDirect" )) core->clampDirect = value;
	else if (!strcmp( name, "clampIndirect" )) core->clampIndirect = value;
	// extensions of this core (the reference fixes these at compile time: core_settings.h:25, pathtracer.h:33)
	else if (!strcmp( name, "maxPathLength" )) core->maxPathLength = value < 1 ? 1 : (value > LH2B_MAXPATHLENGTH ? LH2B_MAXPATHLENGTH : (int)value);
	else if (!strcmp( name, "tileRootShare" )) core->tileRootShare = value < 0 ? 0 : (value > 1 ? 1 : value);	// read by lh2b_tile_create
	else if (!strcmp( name, "tileFilterShard" )) core->tileFilterShard = value > 0 ? 1 : 0;	// read by lh2b_tile_create: every rank filters its own band (filter mode only)
	else if (!strcmp( name, "tileInterleave" )) core->tileInterleave = value > 0 ? 1 : 0;	// read by lh2b_tile_create (with tileFilterShard): rendered rows interleaved over the ranks (1) or the filter band (0)
	else if (!strcmp( name, "gatherMode" )) core->gatherMode = value > 0 ? 1 : 0;	// read by lh2b_gather_create
	else if (!strcmp( name, "l2Persist" )) core->l2Persist = value > 0 ? 1 : 0;	// takes effect at the next FinalizeInstances
	else if (!strcmp( name, "preciseMath" )) { const int m = value > 0 ? 1 : 0; if (m != core->preciseMath) core->preciseMath = m, core->samplesTaken = 0; }
	else if (!strcmp( name, "overlapConnect" )) { FinishFrame( core ); core->overlapConnect = value > 0 ? 1 : 0; }
	else if (!strcmp( name, "pipeline" )) { FinishFrame( core ); core->pipeline = value > 0; }
	else if (!strcmp( name, "bsdf" )) { const int m = value >= 0.5f ? 1 : 0; if (m != core->bsdfModel) core->bsdfModel = m, core->samplesTaken = 0; }
	else if (!strcmp( name, "maxDiffuseBounces" )) core->enoughBounces = value <= 0 ? 0 : (value < 2 ? S_BOUNCED : S_BOUNCEDTWICE);
	else if (!strcmp( name, "bvhBuilder" )) core->bvhBuilder = (int)value;	// 0: GPU PLOC (default), 1: host binned SAH, 2: GPU LBVH
	else if (!strcmp( name, "bvhRefit" )) core->bvhRefit = (int)value;
	else if (!strcmp( name, "bvhCollapse" )) core->bvhCollapse = value > 0 ? 1 : 0;
	else if (!strcmp( name, "plocRadius" )) core->plocRadius = value < 1 ? 1 : (value > 64 ? 64 : (int)value);
	else if (!strcmp( name, "wideBlocksPerSM" )) g_wideBlocksPerSM = value < 1 ? 1 : (int)value;
	else if (!strcmp( name, "triThreshold" )) g_triThreshold = (int)value;
	else if (!strcmp( name, "triThresholdShadow" )) g_triThresholdShadow = (int)value;
	else if (!strcmp( name, "refillThreshold" )) g_refillThreshold = (int)value;
	else if (!strcmp( name, "raysPerLane" )) g_raysPerLane = value < 1 ? 1 : (int)value;
	// unknown names are ignored
	API_END
}

int lh2b_set_probe_pos( lh2b_core* core, int x, int y )
{
	API_BEGIN
	core->probeX = x, core->probeY = y;
	API_END
}

int lh2b_set_textures( lh2b_core* core, const void* tex, int count )
{
	API_BEGIN
	FinishFrame( core );
	core->texDescs.assign( (const lh2abi::CoreTexDesc*)tex, (const lh2abi::CoreTexDesc*)tex + (count > 0 ? count : 0) );
	for (int storage = 0; storage < 3; storage++)
	{
		size_t total = 0;
		for (auto& t : core->texDescs) if ((int)t.storage == storage) total += t.pixelCount;
		const size_t alloc = total > 16 ? total : 16;
		if (storage == lh2abi::ARGB128)
		{
			std::vector<float4> host( alloc, make_float4( 0, 0, 0, 0 ) );
			size_t at = 0;
			for (auto& t : core->texDescs) if ((int)t.storage == storage)
				memcpy( host.data() + at, t.fdata, (size_t)t.pixelCount * sizeof( float4 ) ), t.firstPixel = (uint32_t)at, at += t.pixelCount;
			core->argb128.Upload( host.data(), alloc, core->stream );
			core->stats.argb128TexelCount = (uint32_t)alloc;
		}
		else
		{
			std::vector<uchar4> host( alloc, make_uchar4( 0, 0, 0, 0 ) );
			size_t at = 0;
			for (auto& t : core->texDescs) if ((int)t.storage == storage)
				memcpy( host.data() + at, t.idata, (size_t)t.pixelCount * sizeof( uchar4 ) ), t.firstPixel = (uint32_t)at, at += t.pixelCount;
			(storage == lh2abi::ARGB32 ? core->argb32 : core->nrm32).Upload( host.data(), alloc, core->stream );
			(storage == lh2abi::ARGB32 ? core->stats.argb32TexelCount : core->stats.nrm32TexelCount) = (uint32_t)alloc;
		}
		CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	}
	API_END
}

int lh2b_set_materials( lh2b_core* core, const void* matPtr, int count )
{
	API_BEGIN
	FinishFrame( core );
	const lh2abi::CoreMaterial* mat = (const lh2abi::CoreMaterial*)matPtr;
	std::vector<DevMaterial> host( count > 0 ? count : 1 );
	memset( host.data(), 0, host.size() * sizeof( DevMaterial ) );
	for (int i = 0; i < count; i++)
	{
		const lh2abi::CoreMaterial& m = mat[i];
		DevMaterial& g = host[i];
		const float tr = 1 - m.absorption.value.x, tg = 1 - m.absorption.value.y, tb = 1 - m.absorption.value.z;
		uint32_t flags = (m.eta.value < 1 ? 1u : 0) + (m.color.textureID != -1 ? (1u << 2) : 0) + (m.normals.textureID != -1 ? (1u << 3) : 0) +
			(m.specular.textureID != -1 ? (1u << 4) : 0) + (m.roughness.textureID != -1 ? (1u << 5) : 0) +
			(m.detailNormals.textureID != -1 ? (1u << 7) : 0) + (m.detailColor.textureID != -1 ? (1u << 9) : 0) +
			((m.flags & 1) ? (1u << 11) : 0) + ((m.flags & 2) ? (1u << 12) : 0);
		if (m.color.textureID != -1 && m.color.textureID < (int)core->texDescs.size() && (core->texDescs[m.color.textureID].flags & 8)) flags += 1u << 1;
		g.q[0].x = HalfBits( m.color.value.x ) | (HalfBits( m.color.value.y ) << 16);
		g.q[0].y = HalfBits( m.color.value.z ) | (HalfBits( tr ) << 16);
		g.q[0].z = HalfBits( tg ) | (HalfBits( tb ) << 16);
		g.q[0].w = flags;
		g.q[1].x = Pack4( m.metallic.value, m.subsurface.value, m.specular.value, m.roughness.value );
		g.q[1].y = Pack4( m.specularTint.value, m.anisotropic.value, m.sheen.value, m.sheenTint.value );
		g.q[1].z = Pack4( m.clearcoat.value, m.clearcoatGloss.value, m.transmission.value, 0 );
		memcpy( &g.q[1].w, &m.eta.value, 4 );
		if (m.color.textureID != -1) g.q[2] = MakeMap( core, m.color );
		if (m.detailColor.textureID != -1) g.q[3] = MakeMap( core, m.detailColor );
		if (m.normals.textureID != -1) g.q[4] = MakeMap( core, m.normals );
		if (m.detailNormals.textureID != -1) g.q[5] = MakeMap( core, m.detailNormals );
		if (m.specular.textureID != -1) g.q[6] = MakeMap( core, m.specular );
		if (m.roughness.textureID != -1) g.q[7] = MakeMap( core, m.roughness );
	}
	core->materials.Upload( host.data(), host.size(), core->stream );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

static void UploadLights( lh2b_core* core, DevBuf<float4>& buf, const void* src, int count, int float4sPer )
{
	if (count > 0) buf.Upload( (const float4*)src, (size_t)count * float4sPer, core->stream );
}

int lh2b_set_lights( lh2b_core* core, const void* tri, int nTri, const void* point, int nPoint, const void* spot, int nSpot, const void* dir, int nDir )
{
	API_BEGIN
	FinishFrame( core );
	// (more than MAXISLIGHTS = 64 lights in total: the shade kernel picks uniformly instead of by potential - shade_kernels.cu LightPickProb)
	UploadLights( core, core->triLights, tri, nTri, 6 ), UploadLights( core, core->pointLights, point, nPoint, 2 );
	UploadLights( core, core->spotLights, spot, nSpot, 3 ), UploadLights( core, core->dirLights, dir, nDir, 2 );
	core->lightCounts[0] = nTri, core->lightCounts[1] = nPoint, core->lightCounts[2] = nSpot, core->lightCounts[3] = nDir;
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_set_sky( lh2b_core* core, const float* pixels, int width, int height, const float* worldToLight )
{
	API_BEGIN
	FinishFrame( core );
	const int w = width >> 6, h = height >> 6;
	std::vector<float4> host( (size_t)width * height + (size_t)w * h + 1 );
	for (size_t i = 0; i < (size_t)width * height; i++) host[i] = make_float4( pixels[i * 3], pixels[i * 3 + 1], pixels[i * 3 + 2], 0 );
	float4* scaled = host.data() + (size_t)width * height;
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
	{
		// 64x64 box filter (rendercore.cpp:727-737)
		float4 total = make_float4( 0, 0, 0, 0 );
		const float4* tile = host.data() + x * 64 + (size_t)y * 64 * width;
		for (int v = 0; v < 64; v++) for (int u = 0; u < 64; u++)
		{
			const float4 t = tile[u + (size_t)v * width];
			total.x += t.x, total.y += t.y, total.z += t.z, total.w += t.w;
		}
		const float r = 1.0f / (64 * 64);
		scaled[x + y * w] = make_float4( total.x * r, total.y * r, total.z * r, total.w * r );
	}
	core->skyPixels.Upload( host.data(), host.size(), core->stream );
	core->skyW = width, core->skyH = height;
	if (worldToLight) memcpy( core->worldToSky, worldToLight, 12 * sizeof( float ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_render( lh2b_core* core, const void* viewPtr, int converge, int async )
{
	API_BEGIN
	if (!core->sceneReady) return 0;	// silent no-op before the first FinalizeInstances (rendercore.cpp:821)
	if (core->width == 0) throw CoreError( "Render: SetTarget has not been called" );
	if (core->materials.count == 0) throw CoreError( "Render: SetMaterials has not been called" );
	// pipelined mode (Setting "pipeline" 1, async calls): the frame in flight keeps running while this one is enqueued behind
	// it; the older frame is harvested afterwards, so the device never idles between frames. GetCoreStats / frame stats then
	// describe the previous frame; WaitForRender harvests the one still in flight.
	const bool pipelined = core->pipeline && async && core->frameInFlight;
	if (pipelined) SwapFrameSlots( core ); else FinishFrame( core );
	if (converge == 1 || core->firstConvergingFrame)
	{
		core->samplesTaken = 0;
		core->firstConvergingFrame = true;
		core->camRNGseed = 0x12345678u;
	}
	if (converge == 0) core->firstConvergingFrame = false;
	core->renderStartMs = NowMs();
	RenderFrame( core, *(const lh2abi::ViewPyramid*)viewPtr );
	if (pipelined)
	{
		SwapFrameSlots( core );
		FinishFrame( core );
		SwapFrameSlots( core );
	}
	if (!async) FinishFrame( core );
	API_END
}

int lh2b_wait_for_render( lh2b_core* core )
{
	API_BEGIN
	FinishFrame( core );
	API_END
}

int lh2b_get_stats( lh2b_core* core, void* out )
{
	API_BEGIN
	core->stats.bvhBuildTime = 0;
	for (auto& m : core->meshes) core->stats.bvhBuildTime += m->buildMs * 0.001f;
	memcpy( out, &core->stats, sizeof( lh2abi::CoreStats ) );
	API_END
}

int lh2b_get_frame_stats( lh2b_core* core, lh2b_frame_stats* out )
{
	API_BEGIN
	core->frameStats.buildMs = core->tlasBuildMs;
	*out = core->frameStats;
	API_END
}

int lh2b_read_pixels( lh2b_core* core, float* rgbaOut )
{
	API_BEGIN
	FinishFrame( core );
	if (core->tailEvent) CUDA_CHECK( cudaStreamWaitEvent( core->stream, core->tailEvent, 0 ) );	// sharded filter chain: the peers' bands of the image
	CUDA_CHECK( cudaMemcpyAsync( rgbaOut, core->pixels.ptr, (size_t)core->width * core->height * sizeof( float4 ), cudaMemcpyDeviceToHost, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_read_pixels_async( lh2b_core* core, float* pinnedOut )
{
	API_BEGIN
	// the frame (finalize included) is already enqueued on the launch stream: order the copy behind it on the copy stream
	CUDA_CHECK( cudaEventRecord( core->frameDone, core->stream ) );
	CUDA_CHECK( cudaStreamWaitEvent( core->copyStream, core->frameDone, 0 ) );
	if (core->tailEvent) CUDA_CHECK( cudaStreamWaitEvent( core->copyStream, core->tailEvent, 0 ) );
	CUDA_CHECK( cudaMemcpyAsync( pinnedOut, core->pixels.ptr, (size_t)core->width * core->height * sizeof( float4 ), cudaMemcpyDeviceToHost, core->copyStream ) );
	CUDA_CHECK( cudaEventRecord( core->copyDone[0], core->copyStream ) );
	core->copyPending[0] = true;
	API_END
}

int lh2b_wait_read_pixels( lh2b_core* core )
{
	API_BEGIN
	CUDA_CHECK( cudaStreamSynchronize( core->copyStream ) );
	core->copyPending[0] = core->copyPending[1] = false;
	API_END
}

/* declared here instead of including cuda_gl_interop.h, which needs the OpenGL headers; the symbol lives in the CUDA runtime */
extern "C" cudaError_t cudaGraphicsGLRegisterImage( struct cudaGraphicsResource** resource, unsigned int image, unsigned int target, unsigned int flags );

int lh2b_present_gl( lh2b_core* core, unsigned int glTextureId )
{
	API_BEGIN
	if (glTextureId == 0 || core->width <= 0) throw CoreError( "present_gl: no target" );
	FinishFrame( core );
	if (core->tailEvent) CUDA_CHECK( cudaStreamWaitEvent( core->stream, core->tailEvent, 0 ) );	// sharded filter chain: the peers' bands of the image
	if (core->glResource && (core->glRegisteredTexture != glTextureId || core->glRegisteredW != core->width || core->glRegisteredH != core->height)) ReleaseGlTarget( core );
	if (!core->glResource)
	{
		// interoptexture.cpp:53: cudaGraphicsGLRegisterImage( &res, ID, GL_TEXTURE_2D, cudaGraphicsMapFlagsWriteDiscard )
		const cudaError_t e = cudaGraphicsGLRegisterImage( &core->glResource, glTextureId, 0x0DE1 /* GL_TEXTURE_2D */, cudaGraphicsRegisterFlagsWriteDiscard );
		if (e != cudaSuccess)
		{
			core->glResource = nullptr;
			cudaGetLastError();	// not sticky: clear it
			throw CoreError( std::string( "present_gl: cudaGraphicsGLRegisterImage: " ) + cudaGetErrorString( e ) );
		}
		core->glRegisteredTexture = glTextureId, core->glRegisteredW = core->width, core->glRegisteredH = core->height;
	}
	cudaStream_t s = core->stream;
	CUDA_CHECK( cudaGraphicsMapResources( 1, &core->glResource, s ) );
	cudaArray_t array = nullptr;
	cudaError_t e = cudaGraphicsSubResourceGetMappedArray( &array, core->glResource, 0, 0 );
	if (e == cudaSuccess) e = cudaMemcpy2DToArrayAsync( array, 0, 0, core->pixels.ptr, (size_t)core->width * sizeof( float4 ), (size_t)core->width * sizeof( float4 ),
		(size_t)core->height, cudaMemcpyDeviceToDevice, s );
	cudaGraphicsUnmapResources( 1, &core->glResource, s );
	if (e != cudaSuccess) throw CoreError( std::string( "present_gl: copy into the mapped texture: " ) + cudaGetErrorString( e ) );
	CUDA_CHECK( cudaStreamSynchronize( s ) );
	API_END
}

int lh2b_read_accumulator( lh2b_core* core, float* rgbaOut )
{
	API_BEGIN
	FinishFrame( core );
	CUDA_CHECK( cudaMemcpyAsync( rgbaOut, core->accumulator.ptr, (size_t)core->width * core->height * sizeof( float4 ), cudaMemcpyDeviceToHost, core->stream ) );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_accumulator_device_ptr( lh2b_core* core, void** ptrOut, int* samplesTakenOut )
{
	API_BEGIN
	FinishFrame( core );
	if (ptrOut) *ptrOut = core->accumulator.ptr;
	if (samplesTakenOut) *samplesTakenOut = core->samplesTaken;
	API_END
}

int lh2b_snapshot_accumulator( lh2b_core* core, void* dDst )
{
	API_BEGIN
	if (!dDst) throw CoreError( "snapshot_accumulator: null destination" );
	// stream-ordered behind the frame passed to lh2b_render last (which may still be in flight) and in front of the next one
	CUDA_CHECK( cudaMemcpyAsync( dDst, core->accumulator.ptr, (size_t)core->width * core->height * sizeof( float4 ), cudaMemcpyDeviceToDevice, core->stream ) );
	API_END
}

int lh2b_finalize_external_on( lh2b_core* core, const void* dAccumulator, int samples, void* dPixelsOut, void* stream )
{
	API_BEGIN
	if (samples <= 0) throw CoreError( "finalize_external: samples must be positive" );
	LaunchFinalize( (const float4*)dAccumulator, (float4*)dPixelsOut, core->width * core->height, samples, (cudaStream_t)stream );
	CUDA_CHECK( cudaGetLastError() );
	API_END
}

int lh2b_finalize_external( lh2b_core* core, const void* dAccumulator, int samples )
{
	API_BEGIN
	FinishFrame( core );
	if (samples <= 0) throw CoreError( "finalize_external: samples must be positive" );
	LaunchFinalize( (const float4*)dAccumulator, core->pixels.ptr, core->width * core->height, samples, core->stream );
	CUDA_CHECK( cudaStreamSynchronize( core->stream ) );
	API_END
}

int lh2b_shade_paths( lh2b_core* core, int pathLength, int n, const float* O4, const float* D4, const float* T4, const float* hits,
	uint32_t R0, uint32_t shift, int pass, float* extO, float* extD, float* extT, int* extCount,
	float* shO, float* shD, float* shE, int* shCount, float* accumulator )
{
	API_BEGIN
	FinishFrame( core );
	if (!core->sceneReady || core->width == 0) throw CoreError( "shade_paths: scene and target must be set" );
	const uint32_t stride = (uint32_t)core->width * core->height * core->spp;
	if (n < 0 || (uint32_t)n > stride) throw CoreError( "shade_paths: n exceeds w*h*spp" );
	if (pathLength < 1 || pathLength > core->maxPathLength) throw CoreError( "shade_paths: pathLength out of range" );
	cudaStream_t s = core->stream;
	const size_t bytes = (size_t)n * sizeof( float4 ), accBytes = (size_t)core->width * core->height * sizeof( float4 );
	CUDA_CHECK( cudaMemcpyAsync( core->pathBuf[0][0].ptr, O4, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->pathBuf[0][1].ptr, D4, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->pathBuf[0][2].ptr, T4, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->hitBuf.ptr, hits, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->accumulator.ptr, accumulator, accBytes, cudaMemcpyHostToDevice, s ) );
	DevCounters hc;
	memset( &hc, 0, sizeof( hc ) );
	hc.extensionRays[pathLength - 1] = (uint32_t)n;
	CUDA_CHECK( cudaMemcpyAsync( core->counters.ptr, &hc, offsetof( DevCounters, probedInstid ), cudaMemcpyHostToDevice, s ) );
	lh2abi::ViewPyramid view = core->lastView;
	RenderParams p = BuildParams( core, view );
	p.stride = pathLength == 1 ? (uint32_t)n : stride;
	p.shift = shift, p.pass = pass;
	const bool useNEE = (core->lightCounts[0] + core->lightCounts[1] + core->lightCounts[2] + core->lightCounts[3]) > 0;
	const PathSet in = { core->pathBuf[0][0].ptr, core->pathBuf[0][1].ptr, core->pathBuf[0][2].ptr };
	const PathSet out = { core->pathBuf[1][0].ptr, core->pathBuf[1][1].ptr, core->pathBuf[1][2].ptr };
	const PathSet conn = { core->connBuf[0].ptr, core->connBuf[1].ptr, core->connBuf[2].ptr };
	if (core->preciseMath) LaunchShadePrecise( p, in, out, core->hitBuf.ptr, conn, pathLength, R0, useNEE, (uint32_t)n, (int)core->stats.SMcount, s );
	else LaunchShade( p, in, out, core->hitBuf.ptr, conn, pathLength, R0, useNEE, (uint32_t)n, (int)core->stats.SMcount, s );
	CUDA_CHECK( cudaGetLastError() );
	CUDA_CHECK( cudaMemcpyAsync( &hc, core->counters.ptr, sizeof( hc ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaStreamSynchronize( s ) );
	const uint32_t ne = hc.extensionRays[pathLength], ns = hc.shadowRays[pathLength];
	*extCount = (int)ne, *shCount = (int)ns;
	CUDA_CHECK( cudaMemcpyAsync( extO, out.O, ne * sizeof( float4 ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaMemcpyAsync( extD, out.D, ne * sizeof( float4 ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaMemcpyAsync( extT, out.T, ne * sizeof( float4 ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaMemcpyAsync( shO, conn.O, ns * sizeof( float4 ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaMemcpyAsync( shD, conn.D, ns * sizeof( float4 ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaMemcpyAsync( shE, conn.T, ns * sizeof( float4 ), cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaMemcpyAsync( accumulator, core->accumulator.ptr, accBytes, cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaStreamSynchronize( s ) );
	API_END
}

int lh2b_read_filter_buffers( lh2b_core* core, uint32_t* features, float* worldPos, float* deltaDepth, float* accumulator2 )
{
	API_BEGIN
	FinishFrame( core );
	if (!core->filterEnabled || core->features.count == 0) throw CoreError( "read_filter_buffers: filter mode is off" );
	const size_t px = (size_t)core->width * core->height;
	const int cur = core->filterFlip ^ 1;	// FinishFrame already rotated: the frame just rendered sits in the other set
	cudaStream_t s = core->stream;
	if (features) CUDA_CHECK( cudaMemcpyAsync( features, core->features.ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (worldPos) CUDA_CHECK( cudaMemcpyAsync( worldPos, core->worldPosBuf[cur].ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (deltaDepth) CUDA_CHECK( cudaMemcpyAsync( deltaDepth, core->deltaDepth.ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (accumulator2) CUDA_CHECK( cudaMemcpyAsync( accumulator2, core->accumulator.ptr, 2 * px * 16, cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaStreamSynchronize( s ) );
	API_END
}

/* Filter mode only: the outputs of the last frame's filter chain that the next frame reads as history (any pointer may be null;
   float4[w*h] each, motion float2[w*h]): moments, phase-1 a-trous output, TAA image, phase-3 output, motion vectors. For tests / debugging. */
int lh2b_read_filter_history( lh2b_core* core, float* moments, float* phase1, float* taa, float* phase3, float* motion )
{
	API_BEGIN
	FinishFrame( core );
	if (!core->filterEnabled || core->features.count == 0) throw CoreError( "read_filter_history: filter mode is off" );
	const size_t px = (size_t)core->width * core->height;
	const int cur = core->filterFlip ^ 1;	// the chain already rotated
	cudaStream_t s = core->stream;
	if (moments) CUDA_CHECK( cudaMemcpyAsync( moments, core->momentsBuf[cur].ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (phase1) CUDA_CHECK( cudaMemcpyAsync( phase1, core->filteredBuf[cur ^ 1].ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (taa) CUDA_CHECK( cudaMemcpyAsync( taa, core->taaBuf[cur].ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (phase3) CUDA_CHECK( cudaMemcpyAsync( phase3, core->shading.ptr, px * 16, cudaMemcpyDeviceToHost, s ) );
	if (motion) CUDA_CHECK( cudaMemcpyAsync( motion, core->motion.ptr, px * 8, cudaMemcpyDeviceToHost, s ) );
	CUDA_CHECK( cudaStreamSynchronize( s ) );
	API_END
}

/* Debugging aid: copy one of the core's device tables to host memory (at most maxBytes; *bytesOut = its size). Names: materials, triLights,
   pointLights, spotLights, dirLights, instDesc, blueNoise, sky, argb32, argb128, nrm32, instTrav, nodes, tris. */
int lh2b_debug_read_table( lh2b_core* core, const char* name, void* out, size_t maxBytes, size_t* bytesOut )
{
	API_BEGIN
	FinishFrame( core );
	const void* src = nullptr;
	size_t bytes = 0;
#define TABLE( NAME, BUF ) if (!strcmp( name, NAME )) src = core->BUF.ptr, bytes = core->BUF.count * sizeof( *core->BUF.ptr );
	TABLE( "materials", materials ) TABLE( "triLights", triLights ) TABLE( "pointLights", pointLights ) TABLE( "spotLights", spotLights )
	TABLE( "dirLights", dirLights ) TABLE( "instDesc", instDesc ) TABLE( "blueNoise", blueNoise ) TABLE( "sky", skyPixels )
	TABLE( "argb32", argb32 ) TABLE( "argb128", argb128 ) TABLE( "nrm32", nrm32 ) TABLE( "instTrav", instTrav )
	// wavefront state as the last frame left it: path0* / path1* = the two ping-pong path-state sets (O, D, T), hits, conn* = shadow rays
	TABLE( "path0O", pathBuf[0][0] ) TABLE( "path0D", pathBuf[0][1] ) TABLE( "path0T", pathBuf[0][2] )
	TABLE( "path1O", pathBuf[1][0] ) TABLE( "path1D", pathBuf[1][1] ) TABLE( "path1T", pathBuf[1][2] )
	TABLE( "hits", hitBuf ) TABLE( "connO", connBuf[0] ) TABLE( "connD", connBuf[1] ) TABLE( "connE", connBuf[2] ) TABLE( "counters", counters )
#undef TABLE
	if (!strcmp( name, "nodes" )) src = core->arenaNodes.ptr, bytes = (size_t)core->arenaNodeTop * CW_NODE_QUADS * sizeof( uint4 );
	if (!strcmp( name, "tris" )) src = core->arenaTris.ptr, bytes = (size_t)core->arenaTriTop * 3 * sizeof( float4 );
	if (!src && bytes == 0 && strcmp( name, "nodes" ) && strcmp( name, "tris" )) { bool known = false; for (const char* n : { "materials", "triLights", "pointLights", "spotLights", "dirLights", "instDesc", "blueNoise", "sky", "argb32", "argb128", "nrm32", "instTrav", "path0O", "path0D", "path0T", "path1O", "path1D", "path1T", "hits", "connO", "connD", "connE", "counters" }) known |= !strcmp( name, n ); if (!known) throw CoreError( "debug_read_table: unknown table" ); }
	if (bytesOut) *bytesOut = bytes;
	if (out && src && bytes) CUDA_CHECK( cudaMemcpy( out, src, std::min( bytes, maxBytes ), cudaMemcpyDeviceToHost ) );
	API_END
}

int lh2b_filter_chain( lh2b_core* core, lh2b_filter_io* io )
{
	API_BEGIN
	FinishFrame( core );
	const int w = io->w, h = io->h;
	const size_t px = (size_t)w * h;
	cudaStream_t s = core->stream;
	DevBuf<float4> acc, wp, pwp, dd, shading, moments, pmom, fIN, fOUT, prevPixels, taaOut, target;
	DevBuf<uint4> feat;
	DevBuf<float2> motion;
	acc.Upload( (const float4*)io->accumulator, 2 * px, s ), feat.Upload( (const uint4*)io->features, px, s );
	wp.Upload( (const float4*)io->worldPos, px, s ), pwp.Upload( (const float4*)io->prevWorldPos, px, s ), dd.Upload( (const float4*)io->deltaDepth, px, s );
	pmom.Upload( (const float4*)io->prevMoments, px, s ), fIN.Upload( (const float4*)io->filteredIN, px, s ), prevPixels.Upload( (const float4*)io->prevPixels, px, s );
	shading.Resize( px ), moments.Resize( px ), fOUT.Resize( px ), taaOut.Resize( px ), target.Resize( px ), motion.Resize( px );
	CUDA_CHECK( cudaMemsetAsync( target.ptr, 0, px * 16, s ) );
	CUDA_CHECK( cudaMemsetAsync( taaOut.ptr, 0, px * 16, s ) );
	FilterBuffers b = { acc.ptr, feat.ptr, wp.ptr, pwp.ptr, dd.ptr, shading.ptr, motion.ptr, moments.ptr, pmom.ptr, fIN.ptr, fOUT.ptr, prevPixels.ptr, taaOut.ptr, target.ptr };
	FilterSettings fs;
	fs.w = w, fs.h = h, fs.samplesTaken = io->samplesTaken, fs.camIsStationary = io->camIsStationary, fs.taa = io->taa;
	fs.directClamp = io->directClamp, fs.indirectClamp = io->indirectClamp, fs.j0 = io->j0, fs.j1 = io->j1, fs.prevj0 = io->prevj0, fs.prevj1 = io->prevj1;
	memcpy( fs.prevView, io->prevView, sizeof( fs.prevView ) );
	// the chain overwrites 'shading' three times and filteredIN/OUT once each: run it stage by stage to hand back every output
	FilterSettings one = fs;
	if (core->preciseMath) LaunchFilterChainStagedPrecise( b, one, s, io->shadingAfterPrepare, io->phase1, io->phase2, io->phase3 );
	else LaunchFilterChainStaged( b, one, s, io->shadingAfterPrepare, io->phase1, io->phase2, io->phase3 );
	CUDA_CHECK( cudaGetLastError() );
	auto down = [&]( void* dst, const void* src, size_t bytes ) { if (dst) CUDA_CHECK( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToHost, s ) ); };
	down( io->featuresOut, feat.ptr, px * 16 ), down( io->motion, motion.ptr, px * 8 ), down( io->moments, moments.ptr, px * 16 );
	down( io->taaPixels, taaOut.ptr, px * 16 ), down( io->target, target.ptr, px * 16 );
	CUDA_CHECK( cudaStreamSynchronize( s ) );
	for (int k = 0; k < 8; k++) io->stageMs[k] = 0;
	if (io->timingRuns > 0)
	{
		DevBuf<uint4> feat0;
		feat0.Upload( (const uint4*)io->features, px, s );
		cudaEvent_t ev[7];
		for (auto& e : ev) CUDA_CHECK( cudaEventCreate( &e ) );
		for (int r = 0; r < io->timingRuns + 1; r++)	// the first one warms up
		{
			CUDA_CHECK( cudaMemcpyAsync( feat.ptr, feat0.ptr, px * 16, cudaMemcpyDeviceToDevice, s ) );	// the history counter is updated in place
			CUDA_CHECK( cudaStreamSynchronize( s ) );
			LaunchFilterChain( b, fs, s, ev );
			CUDA_CHECK( cudaEventSynchronize( ev[6] ) );
			if (r == 0) continue;
			for (int k = 0; k < 6; k++) { float ms = 0; CUDA_CHECK( cudaEventElapsedTime( &ms, ev[k], ev[k + 1] ) ); io->stageMs[k] += ms / io->timingRuns; }
			float ms = 0;
			CUDA_CHECK( cudaEventElapsedTime( &ms, ev[0], ev[6] ) );
			io->stageMs[6] += ms / io->timingRuns;
		}
		for (auto& e : ev) cudaEventDestroy( e );
	}
	API_END
}

int lh2b_shade_paths_time( lh2b_core* core, int pathLength, int n, const float* O4, const float* D4, const float* T4, const float* hits,
	uint32_t R0, uint32_t shift, int pass, int runs, float* msOut )
{
	API_BEGIN
	FinishFrame( core );
	if (!core->sceneReady || core->width == 0) throw CoreError( "shade_paths_time: scene and target must be set" );
	const uint32_t stride = (uint32_t)core->width * core->height * core->spp;
	if (n < 0 || (uint32_t)n > stride) throw CoreError( "shade_paths_time: n exceeds w*h*spp" );
	if (pathLength < 1 || pathLength > core->maxPathLength || runs < 1) throw CoreError( "shade_paths_time: pathLength / runs out of range" );
	cudaStream_t s = core->stream;
	const size_t bytes = (size_t)n * sizeof( float4 );
	CUDA_CHECK( cudaMemcpyAsync( core->pathBuf[0][0].ptr, O4, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->pathBuf[0][1].ptr, D4, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->pathBuf[0][2].ptr, T4, bytes, cudaMemcpyHostToDevice, s ) );
	CUDA_CHECK( cudaMemcpyAsync( core->hitBuf.ptr, hits, bytes, cudaMemcpyHostToDevice, s ) );
	DevCounters hc;
	memset( &hc, 0, sizeof( hc ) );
	hc.extensionRays[pathLength - 1] = (uint32_t)n;
	RenderParams p = BuildParams( core, core->lastView );
	p.stride = pathLength == 1 ? (uint32_t)n : stride;
	p.shift = shift, p.pass = pass;
	const bool useNEE = (core->lightCounts[0] + core->lightCounts[1] + core->lightCounts[2] + core->lightCounts[3]) > 0;
	const PathSet in = { core->pathBuf[0][0].ptr, core->pathBuf[0][1].ptr, core->pathBuf[0][2].ptr };
	const PathSet out = { core->pathBuf[1][0].ptr, core->pathBuf[1][1].ptr, core->pathBuf[1][2].ptr };
	const PathSet conn = { core->connBuf[0].ptr, core->connBuf[1].ptr, core->connBuf[2].ptr };
	float best = 1e30f, sum = 0;
	for (int r = 0; r < runs + 1; r++)	// the first one warms up
	{
		CUDA_CHECK( cudaMemcpyAsync( core->counters.ptr, &hc, offsetof( DevCounters, probedInstid ), cudaMemcpyHostToDevice, s ) );
		CUDA_CHECK( cudaStreamSynchronize( s ) );
		CUDA_CHECK( cudaEventRecord( core->evA, s ) );
		LaunchShade( p, in, out, core->hitBuf.ptr, conn, pathLength, R0, useNEE, (uint32_t)n, (int)core->stats.SMcount, s );
		CUDA_CHECK( cudaEventRecord( core->evB, s ) );
		CUDA_CHECK( cudaEventSynchronize( core->evB ) );
		CUDA_CHECK( cudaGetLastError() );
		float ms = 0;
		CUDA_CHECK( cudaEventElapsedTime( &ms, core->evA, core->evB ) );
		if (r > 0) best = ms < best ? ms : best, sum += ms;
	}
	msOut[0] = best, msOut[1] = sum / runs;
	API_END
}

/* Tile-sharded frames (partitioning 2): this core renders rows [y0, y1) of the frame only - same path indices,
   seeds and buffers as the whole frame, so the rows are bit-identical to what one GPU would produce. y0 = y1 = 0: whole frame. */
int lh2b_set_row_band( lh2b_core* core, int y0, int y1 ) { return lh2b_set_row_band_strided( core, y0, y1, 1 ); }

/* The same with a stride: this core renders the 4-row tile rows y0/4 + j * stepTileRows below row y1 (interleaved bands balance the
   load between ranks when the cost of a row varies over the image). stepTileRows > 1 needs y0, y1 and the frame height to be
   multiples of 4. */
int lh2b_set_row_band_strided( lh2b_core* core, int y0, int y1, int stepTileRows )
{
	API_BEGIN
	if (y0 < 0 || y1 < y0 || (core->height > 0 && y1 > core->height) || stepTileRows < 1) throw CoreError( "set_row_band: rows out of range" );
	if (stepTileRows > 1 && ((y0 | y1 | core->height) & 3)) throw CoreError( "set_row_band_strided: strided bands need rows in multiples of 4" );
	if (y1 > y0 && (y0 & 3)) throw CoreError( "set_row_band: the first row of a band must be a multiple of 4 (bands are made of 4-row tile rows)" );
	FinishFrame( core );
	core->bandY0 = y0, core->bandY1 = y1, core->bandStep = stepTileRows, core->samplesTaken = 0;
	API_END
}

int lh2b_set_sample_shard( lh2b_core* core, int firstSample, int totalSpp )
{
	API_BEGIN
	if (firstSample < 0 || totalSpp < 0) throw CoreError( "set_sample_shard: negative argument" );
	core->sampleShardFirst = firstSample, core->sampleShardTotal = totalSpp;
	API_END
}

} // extern "C"
