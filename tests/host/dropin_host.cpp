/* dropin_host.cpp - TEST INFRASTRUCTURE. A minimal stand-in for the reference RenderSystem: loads a core the way
   CoreAPI_Base::CreateCoreAPI does (dlopen + dlsym "CreateCore", lib/RenderSystem/core_api_base.cpp:97-130)
   and drives it through the virtual interface only.

   Built twice from this one source:
     -DLH2_REFERENCE_HEADERS : against the reference's own headers where they lie under /root/reference
                               (recipe: oracle/Makefile -> oracle/_ref/dropin_host_ref). The vtable slots and POD
                               layouts used are then the reference's, which is the drop-in proof.
     (default)               : against include/lh2_core_api.h.
   Modes:  abi            -> prints "name value" lines (sizeof / offsetof) for tests/test_abi.py
           render <lib>   -> renders a small scene through CoreAPI_Base and prints one JSON line
*/
#include <cstring>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <vector>
#include <string>
#include <dlfcn.h>
#ifdef LH2_REFERENCE_HEADERS
#include <immintrin.h>
typedef unsigned int uint;
typedef unsigned char uchar;
typedef unsigned short ushort;
using namespace std;
#include "common_settings.h"
#include "common_types.h"
#include "common_classes.h"
namespace lighthouse2 { class GLTexture { public: uint ID = 0; uint width = 0, height = 0; }; }	// data members of lib/platform/system.h:236-237
#include "core_api_base.h"
namespace api { using namespace lighthouse2; using ::float3; using ::float4; using ::int2; using ::mat4; }	// vector types are global there
static const char* flavour = "reference-headers";
#else
#include "lh2_core_api.h"
namespace api = lh2abi;
static const char* flavour = "own-header";
#endif
#include "lh2b.h"

#define SZ( T ) printf( "sizeof." #T " %zu\n", sizeof( api::T ) )
#define OFF( T, m ) printf( "offsetof." #T "." #m " %zu\n", offsetof( api::T, m ) )

static int PrintAbi()
{
	SZ( CoreTri ); SZ( CoreInstanceDesc ); SZ( CoreMaterial ); SZ( CoreTexDesc ); SZ( CoreLightTri ); SZ( CorePointLight );
	SZ( CoreSpotLight ); SZ( CoreDirectionalLight ); SZ( ViewPyramid ); SZ( CoreStats ); SZ( mat4 ); SZ( GLTexture );
	OFF( CoreTri, ltriIdx ); OFF( CoreTri, material ); OFF( CoreTri, vN0 ); OFF( CoreTri, Nx ); OFF( CoreTri, T ); OFF( CoreTri, area );
	OFF( CoreTri, B ); OFF( CoreTri, alpha ); OFF( CoreTri, LOD ); OFF( CoreTri, vertex0 ); OFF( CoreTri, vertex2 ); OFF( CoreTri, u1_0 );
	OFF( CoreMaterial, color ); OFF( CoreMaterial, detailColor ); OFF( CoreMaterial, normals ); OFF( CoreMaterial, detailNormals );
	OFF( CoreMaterial, flags ); OFF( CoreMaterial, absorption ); OFF( CoreMaterial, metallic ); OFF( CoreMaterial, roughness );
	OFF( CoreMaterial, transmission ); OFF( CoreMaterial, eta ); OFF( CoreMaterial, ior ); OFF( CoreMaterial, urough ); OFF( CoreMaterial, Ks );
	OFF( CoreMaterial, sigma ); OFF( CoreMaterial, specTrans ); OFF( CoreMaterial, flatness ); OFF( CoreMaterial, opacity );
	printf( "sizeof.Vec3Value %zu\nsizeof.ScalarValue %zu\n", sizeof( api::CoreMaterial::Vec3Value ), sizeof( api::CoreMaterial::ScalarValue ) );
	printf( "offsetof.Vec3Value.uvscale %zu\noffsetof.ScalarValue.uvscale %zu\n", offsetof( api::CoreMaterial::Vec3Value, uvscale ), offsetof( api::CoreMaterial::ScalarValue, uvscale ) );
	OFF( CoreTexDesc, pixelCount ); OFF( CoreTexDesc, firstPixel ); OFF( CoreTexDesc, storage );
	OFF( CoreLightTri, energy ); OFF( CoreLightTri, radiance ); OFF( CoreLightTri, triIdx ); OFF( CoreLightTri, instIdx );
	OFF( ViewPyramid, p1 ); OFF( ViewPyramid, aperture ); OFF( ViewPyramid, spreadAngle ); OFF( ViewPyramid, distortion );
	OFF( CoreStats, SMcount ); OFF( CoreStats, bvhBuildTime ); OFF( CoreStats, totalRays ); OFF( CoreStats, renderTime ); OFF( CoreStats, primaryRayCount );
	OFF( CoreStats, traceTime0 ); OFF( CoreStats, deepRayCount ); OFF( CoreStats, shadeTime ); OFF( CoreStats, probedInstid ); OFF( CoreStats, probedWorldPos );
	OFF( CoreInstanceDesc, invTransform );
	return 0;
}

static api::float3 F3( float x, float y, float z ) { api::float3 r; r.x = x, r.y = y, r.z = z; return r; }

static void AddQuad( std::vector<api::float4>& verts, std::vector<api::CoreTri>& tris, api::float3 c, float sx, float sz, float ny, uint material )
{
	// axis-aligned quad in the xz-plane, normal (0, ny, 0)
	const api::float3 p[4] = { F3( c.x - sx, c.y, c.z - sz ), F3( c.x + sx, c.y, c.z - sz ), F3( c.x + sx, c.y, c.z + sz ), F3( c.x - sx, c.y, c.z + sz ) };
	const int order[2][3] = { { 0, 2, 1 }, { 0, 3, 2 } };
	for (int t = 0; t < 2; t++)
	{
		api::CoreTri tri;
		memset( &tri, 0, sizeof( tri ) );
		const int a = order[t][0], b = ny > 0 ? order[t][1] : order[t][2], d = ny > 0 ? order[t][2] : order[t][1];
		tri.vertex0 = p[a], tri.vertex1 = p[b], tri.vertex2 = p[d];
		tri.Nx = 0, tri.Ny = ny, tri.Nz = 0;
		tri.vN0 = tri.vN1 = tri.vN2 = F3( 0, ny, 0 );
		tri.T = F3( 1, 0, 0 ), tri.B = F3( 0, 0, ny );
		tri.material = material, tri.ltriIdx = -1;
		tri.area = 2 * sx * sz, tri.invArea = 1.0f / tri.area;
		tris.push_back( tri );
		const api::float3 q[3] = { p[a], p[b], p[d] };
		for (int k = 0; k < 3; k++) { api::float4 v; v.x = q[k].x, v.y = q[k].y, v.z = q[k].z, v.w = 0; verts.push_back( v ); }
	}
}

static int Render( const char* libPath )
{
	void* mod = dlopen( libPath, RTLD_NOW | RTLD_GLOBAL );
	if (!mod) { printf( "{\"error\": \"dlopen failed: %s\"}\n", dlerror() ); return 2; }
	typedef api::CoreAPI_Base* (*createCoreFunction)();
	createCoreFunction createCore = (createCoreFunction)dlsym( mod, "CreateCore" );
	if (!createCore) { printf( "{\"error\": \"CreateCore not found\"}\n" ); return 2; }
	api::CoreAPI_Base* core = createCore();
	core->Init();
	const int W = 96, H = 64;
	api::GLTexture target;
	target.ID = 0, target.width = W, target.height = H;
	core->SetTarget( &target, 2 );
	core->Setting( "epsilon", 1e-3f );
	core->Setting( "clampValue", 10.0f );
	core->Setting( "someUnknownSetting", 1.0f );	// must be ignored
	// sky: constant
	std::vector<api::float3> sky( 128 * 64, F3( 0.4f, 0.5f, 0.7f ) );
	api::mat4 ident;
	memset( &ident, 0, sizeof( ident ) );
	ident.cell[0] = ident.cell[5] = ident.cell[10] = ident.cell[15] = 1;
	core->SetSkyData( sky.data(), 128, 64, ident );
	core->SetTextures( nullptr, 0 );
	// materials: grey diffuse, mirror-ish, emissive
	std::vector<api::CoreMaterial> mats( 3 );
	memset( mats.data(), 0, mats.size() * sizeof( api::CoreMaterial ) );
	for (auto& m : mats)
	{
		m.color.textureID = m.detailColor.textureID = m.normals.textureID = m.detailNormals.textureID = -1;
		m.specular.textureID = m.roughness.textureID = -1;
		m.color.value = F3( 0.7f, 0.7f, 0.7f ), m.roughness.value = 1.0f, m.eta.value = 1.0f;
	}
	mats[1].color.value = F3( 0.9f, 0.6f, 0.3f ), mats[1].roughness.value = 0.2f;
	mats[2].color.value = F3( 20, 20, 16 );
	core->SetMaterials( mats.data(), (int)mats.size() );
	// mesh 0: floor + raised plate; mesh 1: light quad facing down
	std::vector<api::float4> v0, v1;
	std::vector<api::CoreTri> t0, t1;
	AddQuad( v0, t0, F3( 0, 0, 0 ), 10, 10, 1, 0 );
	AddQuad( v0, t0, F3( 2, 1, 1 ), 2, 2, 1, 1 );
	AddQuad( v1, t1, F3( 0, 8, 0 ), 1.5f, 1.5f, -1, 2 );
	std::vector<api::CoreLightTri> lights;
	for (size_t i = 0; i < t1.size(); i++)
	{
		api::CoreLightTri l;
		memset( &l, 0, sizeof( l ) );
		l.vertex0 = t1[i].vertex0, l.vertex1 = t1[i].vertex1, l.vertex2 = t1[i].vertex2;
		l.centre = F3( (l.vertex0.x + l.vertex1.x + l.vertex2.x) / 3, 8, (l.vertex0.z + l.vertex1.z + l.vertex2.z) / 3 );
		l.N = F3( 0, -1, 0 ), l.area = t1[i].area, l.radiance = mats[2].color.value;
		l.energy = (l.radiance.x + l.radiance.y + l.radiance.z) * l.area;
		l.triIdx = (int)i, l.instIdx = 1;
		t1[i].ltriIdx = (int)i;
		lights.push_back( l );
	}
	core->SetGeometry( 0, v0.data(), (int)v0.size(), (int)t0.size(), t0.data() );
	core->SetGeometry( 1, v1.data(), (int)v1.size(), (int)t1.size(), t1.data() );
	core->SetLights( lights.data(), (int)lights.size(), nullptr, 0, nullptr, 0, nullptr, 0 );
	core->SetInstance( 0, 0, ident );
	core->SetInstance( 1, 1, ident );
	core->SetInstance( 2, -1, ident );
	core->FinalizeInstances();
	api::int2 probe;
	probe.x = W / 2, probe.y = H / 2 + 8;
	core->SetProbePos( probe );
	api::ViewPyramid view;
	memset( &view, 0, sizeof( view ) );
	view.pos = F3( 0, 5, -14 );
	// looking along +z, slightly down: focal plane at distance 5
	const float fd = 5, hs = tanf( 20.0f * 3.14159265f / 180.0f ) * fd, ws = hs * W / H;
	const api::float3 fwd = F3( 0, -0.2425356f, 0.9701425f ), up = F3( 0, 0.9701425f, 0.2425356f );
	const api::float3 C = F3( view.pos.x + fd * fwd.x, view.pos.y + fd * fwd.y, view.pos.z + fd * fwd.z );
	// right = cross( fwd, (0,1,0) ) normalised = (-1, 0, 0) for this forward vector (reference LookAt convention)
	view.p1 = F3( C.x + ws + hs * up.x, C.y + hs * up.y, C.z + hs * up.z );
	view.p2 = F3( C.x - ws + hs * up.x, C.y + hs * up.y, C.z + hs * up.z );
	view.p3 = F3( C.x + ws - hs * up.x, C.y - hs * up.y, C.z - hs * up.z );
	view.aperture = 0, view.spreadAngle = 0.01f, view.imagePlane = 1, view.focalDistance = fd, view.distortion = 0;
	core->Render( view, api::Restart, false );
	core->Render( view, api::Converge, true );
	core->WaitForRender();
	const api::CoreStats stats = core->GetCoreStats();
	// headless read-back through the C handle
	typedef lh2b_core* (*handleFn)(void*);
	typedef int (*readFn)(lh2b_core*, float*);
	handleFn handleOf = (handleFn)dlsym( mod, "lh2b_handle_of" );
	readFn readPixels = (readFn)dlsym( mod, "lh2b_read_pixels" );
	std::vector<float> img( (size_t)W * H * 4, 0.0f );
	int rc = -1;
	if (handleOf && readPixels) rc = readPixels( handleOf( core ), img.data() );
	double sum[3] = { 0, 0, 0 };
	for (int i = 0; i < W * H; i++) for (int k = 0; k < 3; k++) sum[k] += img[i * 4 + k];
	printf( "{\"flavour\": \"%s\", \"device\": \"%s\", \"sm\": %u, \"readback_rc\": %d, \"mean\": [%.6f, %.6f, %.6f], "
		"\"primary\": %u, \"extension\": %u, \"shadow\": %u, \"total\": %u, \"probe\": [%d, %d, %.5f], \"render_time\": %.6f}\n",
		flavour, stats.deviceName ? stats.deviceName : "", stats.SMcount, rc, sum[0] / (W * H), sum[1] / (W * H), sum[2] / (W * H),
		stats.primaryRayCount, stats.totalExtensionRays, stats.totalShadowRays, stats.totalRays,
		stats.probedInstid, stats.probedTriid, stats.probedDist, stats.renderTime );
	core->Shutdown();
	return 0;
}

int main( int argc, char** argv )
{
	if (argc >= 2 && !strcmp( argv[1], "abi" )) return PrintAbi();
	if (argc >= 3 && !strcmp( argv[1], "render" )) return Render( argv[2] );
	fprintf( stderr, "usage: %s abi | render <path to libRenderCore_B200.so>\n", argv[0] );
	return 1;
}
