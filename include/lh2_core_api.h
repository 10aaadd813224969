/* lh2_core_api.h - ABI mirror of the Lighthouse 2 core-facing interface.

   This header restates, for the B200 render core, the binary interface that the
   reference RenderSystem uses to talk to a render core. Nothing here is product
   logic: it is the boundary. Each declaration cites the reference declaration it
   must stay layout-compatible with (paths relative to the reference root):

     vector PODs ............ lib/RenderSystem/common_types.h:59-68
     mat4 ................... lib/RenderSystem/common_types.h:460-479 (row major, 16 floats)
     Convergence ............ lib/RenderSystem/common_classes.h:38-42
     CoreTri ................ lib/RenderSystem/common_classes.h:57-97   (208 B)
     CoreInstanceDesc ....... lib/RenderSystem/common_classes.h:172-179 (80 B)
     CoreMaterial ........... lib/RenderSystem/common_classes.h:216-330 (1344 B)
     CoreTexDesc ............ lib/RenderSystem/common_classes.h:338-362 (40 B)
     CoreLightTri ........... lib/RenderSystem/common_classes.h:365-386 (96 B)
     CorePointLight ......... lib/RenderSystem/common_classes.h:393-406 (32 B)
     CoreSpotLight .......... lib/RenderSystem/common_classes.h:413-430 (48 B)
     CoreDirectionalLight ... lib/RenderSystem/common_classes.h:437-450 (32 B)
     ViewPyramid ............ lib/RenderSystem/common_classes.h:452-475 (68 B)
     GLTexture (prefix) ..... lib/platform/system.h:220-238 (ID, width, height)
     CoreStats .............. lib/RenderSystem/core_api_base.h:30-64    (120 B)
     CoreAPI_Base ........... lib/RenderSystem/core_api_base.h:81-119   (15 virtuals, this order)
     CreateCore ............. lib/rendercore_optix7/core_api.cpp:18-22

   tests/test_abi.py compiles this header next to the reference headers (when the
   reference tree is present) and compares every sizeof/offsetof.
*/
#pragma once
#include <stdint.h>
#include <string.h>

namespace lh2abi
{
typedef unsigned int uint;
typedef unsigned char uchar;

struct alignas( 8 ) int2 { int x, y; };
struct alignas( 8 ) uint2 { uint x, y; };
struct alignas( 8 ) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas( 16 ) float4 { float x, y, z, w; };
struct alignas( 4 ) uchar4 { uchar x, y, z, w; };

/* row-major 4x4; translation lives in cell[3], cell[7], cell[11] */
struct mat4 { float cell[16]; };

enum Convergence { Converge = 0, Restart = 1 };

/* 13 x 16 bytes per triangle. */
struct CoreTri
{
	float u0, u1, u2; int ltriIdx;        /* uv layer 0 (u), emissive tri -> index in tri-light list (else -1) */
	float v0, v1, v2; uint material;      /* uv layer 0 (v), material index */
	float3 vN0; float Nx;                 /* vertex normals; geometric normal is (Nx,Ny,Nz) */
	float3 vN1; float Ny;
	float3 vN2; float Nz;
	float3 T; float area;                 /* tangent */
	float3 B; float invArea;              /* bitangent */
	float3 alpha; float LOD;              /* consistent-normal alphas, texture LOD bias */
	float3 vertex0; float dummy0;
	float3 vertex1; float dummy1;
	float3 vertex2; float dummy2;
	float u1_0, u1_1, u1_2, dummy3;       /* uv layer 1 */
	float v1_0, v1_1, v1_2, dummy4;
};

struct float4x4 { float4 A, B, C, D; };
struct CoreInstanceDesc { void* triangles; int dummy1, dummy2; float4x4 invTransform; };

struct CoreMaterial
{
	struct Vec3Value { float3 value; int textureID; float scale; float2 uvscale, uvoffset; uint2 size; };
	struct ScalarValue { float value; int textureID; int component; float scale; float2 uvscale, uvoffset; uint2 size; };
	Vec3Value color, detailColor, normals, detailNormals;
	uint flags;                           /* bit 0: smooth normals, bit 1: has alpha */
	Vec3Value absorption;
	ScalarValue metallic, subsurface, specular, roughness, specularTint, anisotropic,
		sheen, sheenTint, clearcoat, clearcoatGloss, transmission, eta;
	ScalarValue reflection, refraction, ior;
	char pbrtMaterialType;
	ScalarValue urough, vrough;
	Vec3Value Ks, eta_rgb;
	ScalarValue sigma;
	bool thin;
	ScalarValue specTrans, diffTrans;
	Vec3Value scatterDistance;
	ScalarValue flatness;
	Vec3Value Kr, opacity;
};

enum TexelStorage { ARGB32 = 0, ARGB128 = 1, NRM32 = 2 };
struct CoreTexDesc
{
	union { float4* fdata; uchar4* idata; };
	uint width, height, flags, pixelCount, firstPixel, MIPlevels;
	TexelStorage storage;
};

struct CoreLightTri
{
	float3 centre; float energy;
	float3 N; float area;
	float3 radiance; int dummy2;
	float3 vertex0; int triIdx;
	float3 vertex1; int instIdx;
	float3 vertex2; int dummy1;
};
struct CorePointLight { float3 position; float energy; float3 radiance; int dummy; };
struct CoreSpotLight { float3 position; float cosInner; float3 radiance; float cosOuter; float3 direction; int dummy; };
struct CoreDirectionalLight { float3 direction; float energy; float3 radiance; int dummy; };

struct ViewPyramid
{
	float3 pos, p1, p2, p3;               /* eye; focal-plane corners: top-left, top-right, bottom-left */
	float aperture, spreadAngle, imagePlane, focalDistance, distortion;
};

/* Only the data members the cores read; the reference class has no virtuals. */
struct GLTexture { uint ID; uint width, height; };

struct CoreStats
{
	char* deviceName;
	uint SMcount, ccMajor, ccMinor, VRAM;
	uint argb32TexelCount, argb128TexelCount, nrm32TexelCount;
	float bvhBuildTime;
	uint totalRays, totalExtensionRays, totalShadowRays;
	float renderTime, frameOverhead;
	uint primaryRayCount; float traceTime0;
	uint bounce1RayCount; float traceTime1;
	uint deepRayCount; float traceTimeX;
	float shadowTraceTime, shadeTime, filterTime;
	int probedInstid, probedTriid; float probedDist;
	float3 probedWorldPos;
};

/* The vtable order below IS the interface: the reference calls by slot. */
class CoreAPI_Base
{
public:
	virtual CoreStats GetCoreStats() const = 0;
	virtual void Init() = 0;
	virtual void SetProbePos( const int2 pos ) = 0;
	virtual void SetTarget( GLTexture* target, const uint spp ) = 0;
	virtual void Setting( const char* name, float value ) = 0;
	virtual void Render( const ViewPyramid& view, const Convergence converge, bool async ) = 0;
	virtual void WaitForRender() = 0;
	virtual void Shutdown() = 0;
	virtual void SetTextures( const CoreTexDesc* tex, const int textureCount ) = 0;
	virtual void SetMaterials( CoreMaterial* mat, const int materialCount ) = 0;
	virtual void SetLights( const CoreLightTri* triLights, const int triLightCount,
		const CorePointLight* pointLights, const int pointLightCount,
		const CoreSpotLight* spotLights, const int spotLightCount,
		const CoreDirectionalLight* directionalLights, const int directionalLightCount ) = 0;
	virtual void SetSkyData( const float3* pixels, const uint width, const uint height, const mat4& worldToLight ) = 0;
	virtual void SetGeometry( const int meshIdx, const float4* vertexData, const int vertexCount, const int triangleCount, const CoreTri* triangles ) = 0;
	virtual void SetInstance( const int instanceIdx, const int modelIdx, const mat4& transform ) = 0;
	virtual void FinalizeInstances() = 0;
};

static_assert( sizeof( CoreTri ) == 208, "CoreTri" );
static_assert( sizeof( CoreInstanceDesc ) == 80, "CoreInstanceDesc" );
static_assert( sizeof( CoreMaterial ) == 1344, "CoreMaterial" );
static_assert( sizeof( CoreTexDesc ) == 40, "CoreTexDesc" );
static_assert( sizeof( CoreLightTri ) == 96, "CoreLightTri" );
static_assert( sizeof( CorePointLight ) == 32, "CorePointLight" );
static_assert( sizeof( CoreSpotLight ) == 48, "CoreSpotLight" );
static_assert( sizeof( CoreDirectionalLight ) == 32, "CoreDirectionalLight" );
static_assert( sizeof( ViewPyramid ) == 68, "ViewPyramid" );
static_assert( sizeof( CoreStats ) == 120, "CoreStats" );
static_assert( sizeof( mat4 ) == 64, "mat4" );

} // namespace lh2abi

/* The one exported symbol the reference loader resolves (core_api_base.cpp:124). */
extern "C" __attribute__( ( visibility( "default" ) ) ) lh2abi::CoreAPI_Base* CreateCore();
