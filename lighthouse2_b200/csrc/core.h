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
#include <memory>
#include <vector>

namespace lh2b
{

struct Mesh
{
	int triCount = 0;
	DevBuf<float4> coreTris;		// CoreTri4 view: 13 float4 per triangle (shading data)
	DevBuf<float4> verts;			// float4[3 * triCount] positions as passed to SetGeometry
	DevBuf<uint4> nodes;			// CWBVH nodes, 5 uint4 each
	DevBuf<float4> cwTris;			// traversal triangles, 3 float4 each
	std::vector<float> hostVerts;	// kept for host rebuilds
	Aabb bounds = {};
	bool dirty = true;
	float buildMs = 0, sahCost = 0;
	uint32_t nodeCount = 0;
};

struct Instance { int mesh = 0; float xform[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 }; };

} // namespace lh2b

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
	lh2b::DevBuf<uint4> tlasNodes;
	lh2b::DevBuf<uint32_t> tlasLeafIds;
	lh2b::DevBuf<lh2b::InstTrav> instTrav;
	lh2b::DevScene scene = {};
	float tlasBuildMs = 0;
	uint32_t tlasNodeCount = 0;
	// scratch for the host-buffer query entry points
	lh2b::DevBuf<float4> qO, qD, qHits;
	lh2b::DevBuf<uint8_t> qOcc;
	// settings
	int bvhBuilder = 0;				// 0: GPU LBVH (default), 1: host binned SAH
};
