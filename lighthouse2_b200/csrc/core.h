/* core.h - host-side state of one B200 render core (one CUDA device, one stream).

   Behavioural counterpart of the reference's RenderCore class
   (lib/rendercore_optix7/rendercore.h:39-176) and CoreMesh (lib/rendercore_optix7/core_mesh.h),
   re-designed around device-resident SoA buffers and a hand-built acceleration structure.
*/
#pragma once
#include "../../include/lh2_core_api.h"
#include "../../include/lh2b.h"
#include "bvh.h"
#include "common.cuh"
#include "traverse.cuh"
#include "render_types.h"
#include "kernels.h"
#include <memory>
#include <vector>

namespace lh2b
{

struct Mesh
{
	int triCount = 0;
	DevBuf<float4> coreTris;		// CoreTri4 view: 13 float4 per triangle (shading data)
	DevBuf<float4> verts;			// float4[3 * triCount] positions as passed to SetGeometry
	uint32_t nodeOff = 0, nodeCap = 0;	// slot in the node arena (in nodes)
	uint32_t triOff = 0, triCap = 0;	// slot in the triangle arena (in triangles)
	std::vector<float> hostVerts;	// kept only when the host builder is selected
	bool dirty = true;
	float buildMs = 0, sahCost = 0;
	uint32_t nodeCount = 0;
	int taggedInst = -1;			// instance index currently written into the triangle records (flat scenes)
	// GPU builder state: topology of the last full build (sorted order + binary radix tree) for refits
	DevBuf<uint32_t> topoIdx, topoVisit;
	DevBuf<int2> topoChildren;
	DevBuf<uint32_t> topoSubtree;
	DevBuf<int> topoParent;
	DevBuf<float4> devBounds;		// {lo, hi} of the mesh, device resident (feeds the top-level build)
	DevBuf<uint32_t> devCounts;		// node count, leaf count, overflow flag of the last build
	DevBuf<int> topoWideChild, topoWideSelf;	// binary node behind each slot of each wide node / behind the wide node (in-place refit)
	bool hasTopology = false;
	int builtTriCount = -1;
	cudaEvent_t evStart = nullptr, evEnd = nullptr;	// device time of the last refit, read lazily (no sync on the refit path)
	bool timingPending = false;
	~Mesh() { if (evStart) cudaEventDestroy( evStart ), cudaEventDestroy( evEnd ); }
	// device-side animation (skin_kernels.cu): bind pose captured by lh2b_set_skin / lh2b_set_morph_targets
	DevBuf<float4> bindVerts, bindNormals;	// 3 per triangle
	DevBuf<uint4> skinJoints;
	DevBuf<float4> skinWeights, jointMats;
	DevBuf<float4> morphDeltas, morphNormals;	// [target][vertex]
	DevBuf<float> morphWeights;
	int morphTargets = 0;
	std::vector<float> hostJointMats;		// pageable staging copy of the per-frame pose parameters
};

struct Instance { int mesh = 0; float xform[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 }; };

} // namespace lh2b

struct lh2b_gather;
/* gather.cu: where the frame being enqueued must leave its accumulator snapshot (also orders the stream behind the consumer of
   that buffer two frames ago) */
float4* GatherSnapshotTarget( lh2b_gather* g, cudaStream_t coreStream );

struct lh2b_core
{
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t evA = nullptr, evB = nullptr;
	lh2abi::CoreStats stats = {};
	std::vector<std::unique_ptr<lh2b::Mesh>> meshes;
	std::vector<lh2b::Instance> instances;
	bool sceneReady = false;
	// top level
	lh2b::DevBuf<uint4> arenaNodes;			// every BLAS + the TLAS, 5 uint4 per node
	lh2b::DevBuf<float4> arenaTris;			// every BLAS triangle, 3 float4 each
	uint32_t arenaNodeTop = 0, arenaTriTop = 0;
	uint32_t tlasOff = 0, tlasCap = 0;
	lh2b::DevBuf<uint32_t> tlasLeafIds;
	lh2b::DevBuf<lh2b::InstTrav> instTrav;
	lh2b::DevScene scene = {};
	float tlasBuildMs = 0;
	uint32_t tlasNodeCount = 0;
	// scratch for the host-buffer query entry points
	lh2b::DevBuf<float4> qO, qD, qHits;
	lh2b::DevBuf<uint8_t> qOcc;
	lh2b::DevBuf<uint32_t> queryCounter;
	lh2b::DevBuf<unsigned long long> traceStats;	// 9 counters (TraceStats, traverse_wide.cuh); scene.stats points here while enabled
	bool traceStatsOn = false;
	void* gpuBuild = nullptr;				// GpuBuildScratch (bvh_gpu.cu)
	lh2b::DevBuf<uint8_t> instBuildIn;		// per-frame top-level build input
	lh2b::DevBuf<uint32_t> linkedRoots;
	int bvhCollapse = 1;	// 1: SAH-optimal collapse to 8-wide (dynamic program), 0: greedy largest-area opening
	int plocRadius = 8;	// swept on the 1M-triangle terrain (tools/quality_sweep.py)
	struct lh2b_gather* gather = nullptr;		// attached multi-GPU gather (gather.cu): frames end with a snapshot for it instead of the local finalize
	int bandY0 = 0, bandY1 = 0, bandStep = 1;	// rows this core renders (lh2b_set_row_band[_strided]; 0, 0 = the whole frame)
	float tileRootShare = 1.0f;				// lh2b_tile_create: rank 0's band relative to an equal share (it also runs the frame's tail)
	bool tileDouble = false;					// tile-sharded frames on rank 0 with peers: the buffers the peers push into are double-buffered (tile_gather.cu)
	uint32_t tileFrames = 0;					// frames rendered since the tile gatherer was attached (parity = which set is current)
	lh2b::DevBuf<float4> accumulatorAlt, deltaDepthAlt;	// the other set (swapped with accumulator / deltaDepth every frame while tileDouble)
	// filter chain sharded over the ranks of a tile gatherer (tile_gather.cu; Settings "tileFilterShard", "tileInterleave", read by lh2b_tile_create)
	int tileFilterShard = 0, tileInterleave = 1;
	const lh2b::FilterShard* filterShard = nullptr;		// set while attached: band, halo rows, phase2Out; RunFilter fills the history tables from shardHist
	const float4* shardHist[4][2][LH2B_MAX_SHARDS] = {};	// [worldPos, moments, filtered, taa][flip][rank]: every rank's history buffers (peer mappings, own = local)
	const float4** shardHistDev = nullptr;				// the same table in device memory (the filter kernels index it by the owner of a row)
	float4* worldPosOverride = nullptr;				// the frame being rendered writes its world positions here (staging set of that frame) instead of worldPosBuf[filterFlip]
	uint4* featuresOverride = nullptr;				// ... and its features here (its history-counter bits are then meaningless: the gatherer merges them into 'features')
	cudaStream_t tailStream = nullptr;				// when set, the filter chain runs on this stream (next to the following frame's path tracing on 'stream')
	float4* shardTarget = nullptr;					// where the present pass of the next tail writes (a peer mapping of rank 0's staging image on ranks > 0)
	cudaEvent_t* filterStageEvents = nullptr;			// when set, the next RunFilter records its 7 stage boundaries here (tile timing)
	cudaEvent_t tailEvent = nullptr;					// rank 0: recorded when the peers' rows of the last frame are in 'pixels' too; readers of 'pixels' wait for it
	bool deferTail = false;					// tile-sharded frames: finalize / filter is enqueued by the gatherer once every band has arrived
	int gatherMode = 0;						// lh2b_gather_create: 0 root gather, 1 reduce-scatter (csrc/gather.cu)
	int l2Persist = 1, l2Applied = -1;			// persisting-L2 window over the node arena (ApplyL2Policy)
	void* l2Base = nullptr; size_t l2Bytes = 0;
	int bvhRefit = 1;						// same triangle count re-sent: 1 refit in place (binary + wide tree), 2 refit the binary tree and collapse again, 0 rebuild	// work counter of the persistent query kernels
	// settings
	int bvhBuilder = 0;				// 0: GPU PLOC (default), 1: host binned SAH, 2: GPU LBVH
	float geometryEpsilon = 1e-4f, clampValue = 10.0f;	// reference defaults: stageClampValue(10) at rendercore.cpp:243; epsilon comes from RenderSystem (rendersystem.h:65-72)
	int maxPathLength = 3;			// reference MAXPATHLENGTH (core_settings.h:25)
	int bsdfModel = 0;					// Setting "bsdf": 0 lambert.h, 1 disney.h (what kernels/bsdf.h of the stock cores selects)
	uint32_t enoughBounces = S_BOUNCED;	// reference ENOUGH_BOUNCES (pathtracer.h:33)
	// render target + wavefront buffers (rendercore.cpp:284-326)
	int width = 0, height = 0, spp = 1;
	size_t maxPixels = 0; int allocatedSpp = 0;
	lh2b::DevBuf<float4> pathBuf[2][3];		// ping-pong O/D/T
	lh2b::DevBuf<float4> hitBuf, connBuf[3], accumulator, pixels;
	// asynchronous read-back (lh2b_read_pixels_async): a second pixel buffer and a copy stream, so that the device->host copy
	// of frame k overlaps the kernels of frame k+1. Index 0 of the pair arrays belongs to 'pixels', 1 to 'pixelsAlt'.
	lh2b::DevBuf<float4> pixelsAlt;
	cudaStream_t copyStream = nullptr;
	cudaStream_t connectStream = nullptr;	// connect( L ) runs here, next to extend( L + 1 ) on the launch stream (Setting "overlapConnect")
	int overlapConnect = 1;
	int preciseMath = 0;	// 1: shade and filter stages run their IEEE / libm-accurate builds (parity tests)
	// CUDA-GL interop present path (lh2b_present_gl): the registered target texture
	struct cudaGraphicsResource* glResource = nullptr;
	unsigned glRegisteredTexture = 0; int glRegisteredW = 0, glRegisteredH = 0;
	cudaEvent_t frameDone = nullptr, copyDone[2] = { nullptr, nullptr };
	bool copyPending[2] = { false, false };
	lh2b::DevBuf<lh2b::DevCounters> counters;
	lh2b::DevCounters* hostCounters = nullptr;	// pinned
	// second frame slot for pipelined rendering (Setting "pipeline"): the frame that is still in flight while the next one is enqueued
	lh2b::DevCounters* hostCountersB = nullptr;
	std::vector<cudaEvent_t> eventsB;
	bool frameInFlightB = false, pipeline = false;
	double renderStartMsB = 0;
	lh2abi::ViewPyramid lastViewB = {};
	int maxPathLengthB = 0, filterB = 0;
	int samplesTaken = 0;
	int sampleShardFirst = 0, sampleShardTotal = 0;	// 0 total = not sharded
	bool firstConvergingFrame = true, frameInFlight = false;
	uint32_t camRNGseed = 0x12345678u, shiftSeed = 0x11331445u;	// rendercore.h:122-123
	int probeX = 0, probeY = 0;
	// scene tables
	lh2b::DevBuf<lh2abi::CoreInstanceDesc> instDesc;
	lh2b::DevBuf<lh2b::DevMaterial> materials;
	lh2b::DevBuf<float4> triLights, pointLights, spotLights, dirLights;
	int lightCounts[4] = { 0, 0, 0, 0 };
	lh2b::DevBuf<uchar4> argb32, nrm32;
	lh2b::DevBuf<float4> argb128, skyPixels;
	std::vector<lh2abi::CoreTexDesc> texDescs;
	int skyW = 0, skyH = 0;
	float worldToSky[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
	lh2b::DevBuf<uint32_t> blueNoise;
	// SVGF / TAA (Setting "filter", "TAA", "clampDirect", "clampIndirect"; lib/RenderCore_Optix7Filter/rendercore.cpp:656-678,897-948)
	bool filterEnabled = false, taaEnabled = false, filterHistoryValid = false;
	float clampDirect = 15.0f, clampIndirect = 15.0f;	// RenderSettings defaults (lib/RenderSystem/rendersystem.h:65-72)
	lh2b::DevBuf<uint4> features;
	lh2b::DevBuf<float4> worldPosBuf[2], deltaDepth, shading, momentsBuf[2], filteredBuf[2], taaBuf[2];
	lh2b::DevBuf<float2> motion;
	int filterFlip = 0;				// which of the double-buffered history sets is "current"
	lh2abi::ViewPyramid prevView = {};
	// timing
	std::vector<cudaEvent_t> events;	// per frame: see render.cu
	lh2b_frame_stats frameStats = {};
	double renderStartMs = 0, lastFrameEndMs = 0;
	lh2abi::ViewPyramid lastView = {};
};
