/* lh2b.h - C ABI of the B200 Lighthouse 2 render core (libRenderCore_B200.so).

   One opaque handle, plain pointers and sizes, int return codes (0 = ok, nonzero = failed,
   message via lh2b_last_error()). Nothing is thrown across this boundary and no call aborts the
   process (the reference cores call FatalError/exit; a library that other languages bind
   should not).

   Each entry point replaces one member of the reference's CoreAPI_Base
   (lib/RenderSystem/core_api_base.h:81-119) as implemented by the Optix7 core
   (lib/rendercore_optix7/rendercore.cpp). Struct arguments are passed as const void* to the
   reference's own PODs (layouts restated in include/lh2_core_api.h). The C++ class behind
   CreateCore() (csrc/core_api.cpp) forwards 1:1 to these functions.
*/
#ifndef LH2B_H
#define LH2B_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LH2B_API __attribute__( ( visibility( "default" ) ) )

typedef struct lh2b_core lh2b_core;

/* CoreAPI_Base::Init (core_api_base.h:89, rendercore.cpp:219-278). device < 0: use LOCAL_RANK or 0. */
LH2B_API int lh2b_create( lh2b_core** out, int device );
/* CoreAPI_Base::Shutdown (core_api_base.h:101, rendercore.cpp:985-994). */
LH2B_API int lh2b_destroy( lh2b_core* core );
/* Last error message of the calling thread ("" if none). */
LH2B_API const char* lh2b_last_error( void );

/* CoreAPI_Base::SetTarget (core_api_base.h:93, rendercore.cpp:284-326): only width/height of the
   GLTexture are used; the image is presented in a linear RGBA32F device buffer (lh2b_read_pixels). */
LH2B_API int lh2b_set_target( lh2b_core* core, int width, int height, int spp );
/* CoreAPI_Base::Setting (core_api_base.h:95, rendercore.cpp:746-760). Unknown names are ignored. Names:
     reference (Optix7 core)        epsilon, clampValue, noiseShift (accepted, unused as there)
     reference (Optix7Filter core)  filter, TAA, clampDirect, clampIndirect (lib/RenderCore_Optix7Filter/rendercore.cpp:656-678)
     path tracer (compile-time in the reference)
                                    maxPathLength (1..LH2B_MAXPATHLENGTH, default 3 = MAXPATHLENGTH), maxDiffuseBounces (1 = ENOUGH_BOUNCES
                                    S_BOUNCED default, 2, 0 = unlimited), bsdf (0 lambert.h model, 1 principled model of disney.h)
     acceleration structure         bvhBuilder (0 GPU PLOC default, 1 host binned SAH, 2 GPU LBVH), bvhRefit (1 in-place refit default, 2 refit +
                                    re-collapse, 0 rebuild), plocRadius (1..64, default 8), bvhCollapse (1 SAH-optimal collapse to 8-wide, default; 0 greedy),
                                    l2Persist (1 default: persisting-L2 window over the node arena)
     frame scheduling               pipeline (1: Render( async ) enqueues frame k+1 behind frame k; statistics lag one frame),
                                    overlapConnect (1 default: connect( L ) runs next to extend( L + 1 ) on a second stream),
                                    gatherMode (read by lh2b_gather_create: 0 root gather default, 1 reduce-scatter),
                                    tileRootShare (read by lh2b_tile_create: rank 0's band relative to an equal share, 0..1),
                                    tileFilterShard (read by lh2b_tile_create, filter mode, 2..8 ranks: 1 = every rank also filters a band
                                    of the frame, see below; default 0 = the tail runs on rank 0), tileInterleave (with tileFilterShard:
                                    1 default = rendered rows interleaved over the ranks, 0 = a rank renders its filter band)
     numerics                       preciseMath (1: the shade and filter stages run their IEEE / libm-accurate builds - no fast math, no FMA
                                    contraction; default 0 = the reference's -use_fast_math behaviour)
     kernel tuning (measurement)    wideBlocksPerSM, triThreshold, triThresholdShadow, refillThreshold */
LH2B_API int lh2b_setting( lh2b_core* core, const char* name, float value );
/* CoreAPI_Base::SetProbePos (core_api_base.h:91, rendercore.cpp:85-88). */
LH2B_API int lh2b_set_probe_pos( lh2b_core* core, int x, int y );
/* CoreAPI_Base::SetTextures (core_api_base.h:103, rendercore.cpp:438-502). tex: CoreTexDesc[count]. */
LH2B_API int lh2b_set_textures( lh2b_core* core, const void* tex, int count );
/* CoreAPI_Base::SetMaterials (core_api_base.h:105, rendercore.cpp:508-565). mat: CoreMaterial[count]. */
LH2B_API int lh2b_set_materials( lh2b_core* core, const void* mat, int count );
/* CoreAPI_Base::SetLights (core_api_base.h:107-110, rendercore.cpp:698-710). */
LH2B_API int lh2b_set_lights( lh2b_core* core, const void* triLights, int triLightCount,
	const void* pointLights, int pointLightCount, const void* spotLights, int spotLightCount,
	const void* directionalLights, int directionalLightCount );
/* CoreAPI_Base::SetSkyData (core_api_base.h:112, rendercore.cpp:716-740). pixels: float3[w*h]; worldToLight: 16 floats row major. */
LH2B_API int lh2b_set_sky( lh2b_core* core, const float* pixels, int width, int height, const float* worldToLight );
/* CoreAPI_Base::SetGeometry (core_api_base.h:114, rendercore.cpp:332-340, core_mesh.cpp:34-61).
   vertexData: float4[vertexCount] (3 per triangle, not indexed); triangles: CoreTri[triangleCount]. */
LH2B_API int lh2b_set_geometry( lh2b_core* core, int meshIdx, const float* vertexData, int vertexCount,
	int triangleCount, const void* triangles );
/* The same with DEVICE pointers (e.g. the output of the host's own simulation kernels): device-to-device copies, nothing
   crosses PCIe. Pointers must be valid on the core's device. */
LH2B_API int lh2b_set_geometry_device( lh2b_core* core, int meshIdx, const void* dVertexData, int vertexCount,
	int triangleCount, const void* dTriangles );
/* Device-side animation: replaces HostMesh::SetPose + the SetGeometry upload that follows it in
   RenderSystem::UpdateSceneGraph (lib/RenderSystem/host_mesh.cpp:711-741 morph targets, :748-906 skinning). The current
   geometry of the mesh becomes the bind pose; every later pose call rewrites positions, CoreTri::vertex0..2, vN0..2 (and
   Nx/Ny/Nz for skinning) on the device and marks the mesh for a refit at the next FinalizeInstances.
   joints4: uint4 per vertex; weights4: float4 per vertex; jointMatrices16: row-major 4x4 per joint (HostSkin::jointMat).
   deltas4 / normals4: float4[targetCount][vertexCount] (poses 1..n of HostMesh::poses); weights: one per target. */
LH2B_API int lh2b_set_skin( lh2b_core* core, int meshIdx, const uint32_t* joints4, const float* weights4, int vertexCount );
LH2B_API int lh2b_set_pose( lh2b_core* core, int meshIdx, const float* jointMatrices16, int jointCount );
LH2B_API int lh2b_set_morph_targets( lh2b_core* core, int meshIdx, const float* deltas4, const float* normals4, int targetCount, int vertexCount );
LH2B_API int lh2b_set_morph_weights( lh2b_core* core, int meshIdx, const float* weights, int targetCount );
/* Read back the current geometry of a mesh (tests, debugging): float4[3 * triangleCount] and / or CoreTri[triangleCount]. */
LH2B_API int lh2b_read_geometry( lh2b_core* core, int meshIdx, float* vertexDataOut, void* trianglesOut );
/* CoreAPI_Base::SetInstance (core_api_base.h:116, rendercore.cpp:346-376). transform: 16 floats row major;
   meshIdx == -1 truncates the instance list at instanceIdx. */
LH2B_API int lh2b_set_instance( lh2b_core* core, int instanceIdx, int meshIdx, const float* transform );
/* CoreAPI_Base::FinalizeInstances (core_api_base.h:118, rendercore.cpp:382-432). */
LH2B_API int lh2b_finalize_instances( lh2b_core* core );
/* CoreAPI_Base::Render (core_api_base.h:97, rendercore.cpp:819-938). view: ViewPyramid (68 bytes);
   converge: 0 = Converge, 1 = Restart. */
LH2B_API int lh2b_render( lh2b_core* core, const void* view, int converge, int async );
/* CoreAPI_Base::WaitForRender (core_api_base.h:99, rendercore.cpp:945-957). */
LH2B_API int lh2b_wait_for_render( lh2b_core* core );
/* CoreAPI_Base::GetCoreStats (core_api_base.h:87, rendercore.cpp:1000-1003). out: CoreStats (120 bytes). */
LH2B_API int lh2b_get_stats( lh2b_core* core, void* outCoreStats );

/* ---- headless extras (no reference counterpart: the reference presents through OpenGL) ---- */

/* Copy the finalized image (accumulator / samplesTaken, finalize_shared.h:29-45) to host: float4[w*h]. */
LH2B_API int lh2b_read_pixels( lh2b_core* core, float* rgbaOut );
/* Pipelined read-back: enqueues the device->host copy of the frame last passed to lh2b_render (it may still be in flight) on
   the core's copy stream and returns at once; 'pinnedOut' must be page-locked and stay valid until lh2b_wait_read_pixels.
   The next lh2b_render presents into a second pixel buffer, so the copy overlaps the next frame's kernels. */
LH2B_API int lh2b_read_pixels_async( lh2b_core* core, float* pinnedOut );
LH2B_API int lh2b_wait_read_pixels( lh2b_core* core );
/* Copy the raw accumulator to host: float4[w*h]. */
LH2B_API int lh2b_read_accumulator( lh2b_core* core, float* rgbaOut );
/* Device pointer of the accumulator (float4[w*h]) for zero-copy collectives; samples taken so far. */
LH2B_API int lh2b_accumulator_device_ptr( lh2b_core* core, void** ptrOut, int* samplesTakenOut );
/* Finalize an externally reduced accumulator (float4[w*h] on this device, e.g. the NCCL sum over the
   shards) into this core's pixel buffer: pixels = accum / samples. */
LH2B_API int lh2b_finalize_external( lh2b_core* core, const void* dAccumulator, int samples );
/* Pipelined multi-GPU frames (lighthouse2_b200/distributed.py): lh2b_snapshot_accumulator enqueues a device-to-device copy of
   the accumulator on the core's stream - behind the frame last passed to lh2b_render, in front of the next one - so that the
   reduce of frame k can run on another stream while frame k+1 renders; lh2b_finalize_external_on launches the finalize kernel
   on the caller's stream (cudaStream_t), reading the reduced accumulator and writing float4[w*h] to dPixelsOut. */
LH2B_API int lh2b_snapshot_accumulator( lh2b_core* core, void* dDst );
LH2B_API int lh2b_finalize_external_on( lh2b_core* core, const void* dAccumulator, int samples, void* dPixelsOut, void* stream );

/* ---- host builder introspection (needs no device) -------------------------------------------------------------------------
   The wide BVH the host builder (Setting "bvhBuilder" 1) produces for one mesh: 128-byte nodes and 48-byte triangle records in the
   layout documented in csrc/bvh.h, root = node 0, indices relative to the returned arrays. counts[0] / counts[1] receive the
   node / triangle-record counts; returns 1 when maxNodes / maxTris are too small. */
LH2B_API int lh2b_host_bvh_build( const float* verts4, int triCount, void* nodesOut, int maxNodes, void* trisOut, int maxTris, int* counts );

/* ---- tile (row-band) sharding of one frame: strong scaling for real-time frames (csrc/tile_gather.cu) -----------
   lh2b_set_row_band: this core renders rows [y0, y1) only (y0 = y1 = 0: the whole frame), with the path indices, seeds and buffers
   of the whole frame. The tile gatherer (one per rank, created after lh2b_set_target and the filter setting) assigns the bands,
   defers the frame's tail and, per frame - lh2b_render( ..., async = 1 ) then lh2b_tile_frame( g ) on every rank - moves the
   peers' rows into rank 0's buffers over NVLink and runs the filter chain / finalize there. Handles are exchanged like the
   gather's: lh2b_tile_handle_bytes() per rank, all-gathered in rank order. While a tile gatherer is attached only rank 0 presents:
   the peers' pixel buffers are not updated (their frames end after the last connect), and a probe pixel outside a rank's band is
   not probed by that rank.
   Setting "tileFilterShard" 1 (before lh2b_tile_create; filter mode): the SVGF / TAA chain is sharded as well - rank r filters the rows
   [r * B, (r+1) * B) (B: an equal share rounded up to 16 rows) plus 16 halo rows, reads the history rows of other bands from their
   owners over NVLink from inside the filter kernels, and presents its band into rank 0's image; the chain of frame k runs on its own
   stream next to the path tracing of frame k + 1. Same calls, same result on rank 0 (bit-identical to one GPU at 1 spp); lh2b_read_pixels*
   on rank 0 wait for the peers' bands. The frame height must be a multiple of 4 and give every rank a non-empty band. Every rank must
   have finished its frames (lh2b_tile_wait, then a barrier of the caller's) before any rank calls lh2b_tile_destroy: the peers read this
   rank's history buffers until their own chain is through. */
typedef struct lh2b_tile_gather lh2b_tile_gather;
LH2B_API int lh2b_set_row_band( lh2b_core* core, int y0, int y1 );
LH2B_API int lh2b_set_row_band_strided( lh2b_core* core, int y0, int y1, int stepTileRows );	/* tile rows y0/4 + j * step below row y1 */
LH2B_API int lh2b_tile_handle_bytes( void );
LH2B_API int lh2b_tile_layout( int height, int world, float rootShare, int rank, int* y0, int* y1, int* stepTileRows );	/* the band of a rank; needs no device */
/* layout of the sharded filter chain (tileFilterShard; no device needed): band = { y0, y1, stepTileRows } rendered by 'rank', the rows it filters
   and presents, those rows plus the 16 halo rows, and the strip of world positions it receives; and the tile rows of a band inside a row range */
LH2B_API int lh2b_tile_shard_layout( int height, int world, int interleave, int rank, int* band3, int* filterBand2, int* withHalo2, int* worldPosStrip2 );
LH2B_API int lh2b_tile_rows_inside( int y0, int y1, int stepTileRows, int e0, int e1, int* firstTileRow, int* count );
LH2B_API int lh2b_tile_create( lh2b_core* core, int rank, int world, lh2b_tile_gather** out );
LH2B_API int lh2b_tile_export( lh2b_tile_gather* g, void* handlesOut );
LH2B_API int lh2b_tile_import( lh2b_tile_gather* g, const void* handlesOfAllRanks );
LH2B_API int lh2b_tile_frame( lh2b_tile_gather* g );
LH2B_API int lh2b_tile_wait( lh2b_tile_gather* g );
LH2B_API int lh2b_tile_rows( lh2b_tile_gather* g, int* y0, int* y1, int* stepTileRows );
LH2B_API int lh2b_tile_destroy( lh2b_tile_gather* g );

/* ---- multi-GPU frame gather over NVLink peer memory (csrc/gather.cu) ----------------------------------------
   One process per GPU renders its sample shard (lh2b_set_sample_shard); rank 0 ends every frame with the summed, finalized
   image. Peer copies by the copy engines + stream memory operations for the hand-shake + one fused sum/finalize kernel; no
   library collective on the data path, nothing blocks the host. Set-up: every rank creates its end, exports
   lh2b_gather_handle_bytes() bytes of CUDA IPC handles, the caller all-gathers them (rank order) and every rank imports the
   whole array. Per frame, on every rank: lh2b_render( ..., async = 1 ) then lh2b_gather_frame( g, samplesOfAllRanks, pinnedOut )
   (pinnedOut: rank 0 only, page-locked float4[w*h] or null). lh2b_gather_wait blocks until this rank's part is complete. */
typedef struct lh2b_gather lh2b_gather;
LH2B_API int lh2b_gather_handle_bytes( void );
LH2B_API int lh2b_gather_create( lh2b_core* core, int rank, int world, lh2b_gather** out );
LH2B_API int lh2b_gather_export( lh2b_gather* g, void* handlesOut );
LH2B_API int lh2b_gather_import( lh2b_gather* g, const void* handlesOfAllRanks );
LH2B_API int lh2b_gather_frame( lh2b_gather* g, int samplesTotal, float* pinnedOut );
LH2B_API int lh2b_gather_wait( lh2b_gather* g );
LH2B_API int lh2b_gather_join( lh2b_gather* g, void* stream );	/* make a cudaStream_t wait for the work enqueued so far */
LH2B_API int lh2b_gather_image_device_ptr( lh2b_gather* g, void** ptrOut );	/* rank 0: device image of the newest frame */
LH2B_API int lh2b_gather_destroy( lh2b_gather* g );
/* Multi-GPU sample sharding: this core renders sample indices [first, first+spp) of each pass and
   seeds as if it were part of a 'total'-spp frame. Default: first 0, total = spp. */
LH2B_API int lh2b_set_sample_shard( lh2b_core* core, int firstSample, int totalSpp );

/* Closest-hit query over the current scene with HOST buffers (the traversal-only parity and
   benchmark entry; same semantics as setupSecondaryRay, .optix.cu:131-140).
   origins/directions: float4[n] (w ignored); hitsOut: float4[n] = (u16|v16<<16, inst, prim, t) bit
   patterns, prim = -1 on a miss. */
LH2B_API int lh2b_trace_rays( lh2b_core* core, const float* origins, const float* directions, int n, float* hitsOut );
/* Occlusion query (generateShadowRay, .optix.cu:142-154): directions.w = tmax; occludedOut[i] = 1 if blocked. */
LH2B_API int lh2b_trace_shadow_rays( lh2b_core* core, const float* origins, const float* directions, int n, uint8_t* occludedOut );
/* Same queries on DEVICE buffers, asynchronous on the core's stream; msOut (optional) receives the
   kernel time of 'repeat' back-to-back launches measured with CUDA events on that stream. */
LH2B_API int lh2b_trace_rays_device( lh2b_core* core, const void* dOrigins, const void* dDirections, int n, void* dHitsOut, int repeat, float* msOut );
LH2B_API int lh2b_trace_shadow_rays_device( lh2b_core* core, const void* dOrigins, const void* dDirections, int n, void* dOccludedOut, int repeat, float* msOut );
/* Traversal work counters (measurement; the issue roofline of bench.py is computed from them): while enabled, every traversal
   launch of this core - frame stages and ray queries - runs its counting instantiation (a few per cent slower) and adds to
   nine 64-bit device counters. lh2b_trace_stats_read copies them out (and clears them when 'reset' is set):
     [0] rays  [1] node steps  [2] triangle tests  [3] instance entries            (per ray)
     [4] loop iterations  [5] node phases  [6] triangle phases                      (per warp)
     [7] lanes active over all node phases  [8] lanes active over all triangle phases */
LH2B_API int lh2b_trace_stats_enable( lh2b_core* core, int on );
LH2B_API int lh2b_trace_stats_read( lh2b_core* core, unsigned long long* out9, int reset );
/* The CUDA stream the core launches on (cudaStream_t as void*). */
LH2B_API int lh2b_stream( lh2b_core* core, void** streamOut );
/* Per-stage device times of the last Render in ms and ray counts; see lh2b_frame_stats below. */
typedef struct lh2b_frame_stats
{
	float generateExtendMs, extendMs, shadeMs, connectMs, finalizeMs, buildMs, filterMs, totalMs;
	uint32_t primaryRays, extensionRays, shadowRays, kernelLaunches;
	uint32_t pathLengthReached, reserved[3];
} lh2b_frame_stats;
LH2B_API int lh2b_get_frame_stats( lh2b_core* core, lh2b_frame_stats* out );
/* Acceleration-structure statistics of a mesh (meshIdx >= 0) or the top level (meshIdx = -1). */
typedef struct lh2b_bvh_stats { uint32_t nodes, triangles, bytes; float buildMs; float sahCost; uint32_t reserved[3]; } lh2b_bvh_stats;
LH2B_API int lh2b_get_bvh_stats( lh2b_core* core, int meshIdx, lh2b_bvh_stats* out );

/* Parity hook: run the shade stage alone (shadeKernel, lib/rendercore_optix7/kernels/pathtracer.h:54-238) on HOST
   buffers of n paths at the given path length, with explicit per-frame random state (R0, shift, pass). Outputs: the
   compacted extension rays and shadow rays (order unspecified; match by the path / pixel index they carry) and the
   accumulator (float4[w*h], in/out). Uses the current target, scene tables and settings. */
LH2B_API int lh2b_shade_paths( lh2b_core* core, int pathLength, int n, const float* O4, const float* D4, const float* T4, const float* hits,
	uint32_t R0, uint32_t shift, int pass, float* extO, float* extD, float* extT, int* extCount,
	float* shO, float* shD, float* shE, int* shCount, float* accumulator );

/* Parity hook: run the SVGF / TAA chain alone on HOST buffers, exactly as the filter core's FinalizeRender does
   (lib/RenderCore_Optix7Filter/rendercore.cpp:897-948): prepareFilter, applyFilter phases 1-3, then TAApass + unsharpenTAA
   (taa = 1) or finalizeNoTAA. All image buffers are float4[w*h] except accumulator (float4[2*w*h]: direct, indirect),
   features (uint4[w*h]) and motion (float2[w*h]). Every stage's output is returned. */
typedef struct lh2b_filter_io
{
	int w, h, samplesTaken, camIsStationary, taa;
	float directClamp, indirectClamp, j0, j1, prevj0, prevj1;
	float prevView[17];
	const float* accumulator; const uint32_t* features; const float* worldPos; const float* prevWorldPos; const float* deltaDepth;
	const float* prevMoments; const float* filteredIN; const float* prevPixels;
	uint32_t* featuresOut; float* shadingAfterPrepare; float* motion; float* moments;
	float* phase1; float* phase2; float* phase3; float* taaPixels; float* target;
	int timingRuns;		/* > 0: after the checked run, time that many more runs of the whole chain on the same inputs */
	float stageMs[8];	/* mean ms: prepare, a-trous 1, 2, 3, TAA, present, whole chain, 0 (CUDA events on the core's stream) */
} lh2b_filter_io;
LH2B_API int lh2b_filter_chain( lh2b_core* core, lh2b_filter_io* io );
/* Measurement twin of lh2b_shade_paths: uploads the same inputs once and times 'runs' launches of the shade kernel on them with
   CUDA events on the core's stream (the kernel reads one path-state set and writes the other, so nothing needs restoring but the
   counters). msOut[0] = fastest, msOut[1] = mean launch. */
LH2B_API int lh2b_shade_paths_time( lh2b_core* core, int pathLength, int n, const float* O4, const float* D4, const float* T4, const float* hits,
	uint32_t R0, uint32_t shift, int pass, int runs, float* msOut );
/* Filter mode only: copy the per-pixel filter inputs of the last frame to host (any pointer may be null):
   features uint4[w*h], worldPos / deltaDepth float4[w*h], accumulator2 float4[2*w*h] (direct, then indirect). */
LH2B_API int lh2b_read_filter_buffers( lh2b_core* core, uint32_t* features, float* worldPos, float* deltaDepth, float* accumulator2 );
/* ... and what the chain of the last frame left for the next one (any pointer may be null; float4[w*h], motion float2[w*h]):
   luminance moments, phase-1 a-trous output, TAA image, phase-3 output, motion vectors. */
LH2B_API int lh2b_read_filter_history( lh2b_core* core, float* moments, float* phase1, float* taa, float* phase3, float* motion );
/* Debugging aid: one of the core's device tables (materials, triLights, pointLights, spotLights, dirLights, instDesc, blueNoise, sky, argb32,
   argb128, nrm32, instTrav, nodes, tris) copied to host memory; *bytesOut = its size, at most maxBytes are copied (out may be null). */
LH2B_API int lh2b_debug_read_table( lh2b_core* core, const char* name, void* out, size_t maxBytes, size_t* bytesOut );

/* Presenting through CUDA-OpenGL interop, as the reference's InteropTexture does (lib/CUDA/shared_host_code/interoptexture.cpp:53-61:
   cudaGraphicsGLRegisterImage on GLTexture::ID, map, write, unmap): copies the finished frame from the core's linear RGBA32F
   pixel buffer into the GL_RGBA32F texture 'glTextureId' on the device - no host round trip. Must be called from the thread that owns
   the GL context, after the frame is finished (lh2b_wait_for_render or a synchronous lh2b_render). The texture is registered on
   first use and re-registered when the id or the target size changes. Returns non-zero (see lh2b_last_error) when the process has no
   usable GL context or the texture cannot be registered - the C++ class then falls back to glTexSubImage2D from host memory. */
LH2B_API int lh2b_present_gl( lh2b_core* core, unsigned int glTextureId );

/* The C handle behind a CoreAPI_Base* obtained from CreateCore() (for headless read-back and statistics). */
LH2B_API lh2b_core* lh2b_handle_of( void* coreApiBase );

#ifdef __cplusplus
}
#endif
#endif
