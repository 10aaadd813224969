/* ref_recorder_core.cpp - TEST INFRASTRUCTURE ONLY. A CoreAPI_Base implementation, compiled against the reference's own
   headers where they lie (oracle/Makefile -> oracle/_ref/libRenderCore_Recorder.so), that renders nothing: it stores every
   byte the reference RenderSystem hands to a core through the 15 virtuals (lib/RenderSystem/core_api_base.h:87-118) and, on
   Render, writes them to the file named by $LH2_RECORD_PATH. tests/ load that file (oracle/binding.py: load_recording) to feed
   the CPU oracle - and our core through its Python mirror - with exactly the scene RenderSystem built from the reference's
   assets (BASELINE.json configs[0]: the tinyapp scene).

   File format: sections of  char tag[16] | uint64 a | uint64 b | uint64 bytes | payload , in call order; 'a' and 'b' carry
   small per-call integers (mesh index, counts ...). Later sections with the same tag and 'a' supersede earlier ones.
*/
#include <cstring>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
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
using namespace lighthouse2;

namespace
{

struct Section { char tag[16]; uint64_t a, b; std::vector<uint8_t> data; };

class RecorderCore : public CoreAPI_Base
{
public:
	CoreStats GetCoreStats() const override { CoreStats s; memset( &s, 0, sizeof( s ) ); s.probedTriid = -1; return s; }
	void Init() override {}
	void SetProbePos( const int2 pos ) override { const int p[2] = { pos.x, pos.y }; Put( "probe", 0, 0, p, sizeof( p ) ); }
	void SetTarget( GLTexture* target, const uint spp ) override
	{
		const uint32_t t[3] = { target->width, target->height, spp };
		Put( "target", 0, 0, t, sizeof( t ) );
	}
	void Setting( const char* name, float value ) override
	{
		char buf[64] = {};
		strncpy( buf, name, 59 );
		memcpy( buf + 60, &value, 4 );
		Put( "setting", sections.size(), 0, buf, sizeof( buf ) );
	}
	void Render( const ViewPyramid& view, const Convergence converge, bool ) override
	{
		Put( "view", 0, (uint64_t)converge, &view, sizeof( view ) );
		const char* path = getenv( "LH2_RECORD_PATH" );
		if (!path) return;
		FILE* f = fopen( path, "wb" );
		if (!f) { fprintf( stderr, "recorder core: cannot write %s\n", path ); return; }
		for (const Section& s : sections)
		{
			const uint64_t bytes = s.data.size();
			fwrite( s.tag, 1, 16, f ), fwrite( &s.a, 8, 1, f ), fwrite( &s.b, 8, 1, f ), fwrite( &bytes, 8, 1, f );
			if (bytes) fwrite( s.data.data(), 1, bytes, f );
		}
		fclose( f );
	}
	void WaitForRender() override {}
	void Shutdown() override {}
	void SetTextures( const CoreTexDesc* tex, const int textureCount ) override
	{
		Put( "texdescs", 0, textureCount, tex, sizeof( CoreTexDesc ) * (size_t)textureCount );
		for (int i = 0; i < textureCount; i++)
		{
			const size_t texel = tex[i].storage == ARGB128 ? 16 : 4;
			Put( "texels", i, tex[i].storage, tex[i].idata, texel * tex[i].pixelCount );
		}
	}
	void SetMaterials( CoreMaterial* mat, const int materialCount ) override { Put( "materials", 0, materialCount, mat, sizeof( CoreMaterial ) * (size_t)materialCount ); }
	void SetLights( const CoreLightTri* triLights, const int triLightCount, const CorePointLight* pointLights, const int pointLightCount,
		const CoreSpotLight* spotLights, const int spotLightCount, const CoreDirectionalLight* directionalLights, const int directionalLightCount ) override
	{
		Put( "trilights", 0, triLightCount, triLights, sizeof( CoreLightTri ) * (size_t)triLightCount );
		Put( "pointlights", 0, pointLightCount, pointLights, sizeof( CorePointLight ) * (size_t)pointLightCount );
		Put( "spotlights", 0, spotLightCount, spotLights, sizeof( CoreSpotLight ) * (size_t)spotLightCount );
		Put( "dirlights", 0, directionalLightCount, directionalLights, sizeof( CoreDirectionalLight ) * (size_t)directionalLightCount );
	}
	void SetSkyData( const float3* pixels, const uint width, const uint height, const mat4& worldToLight ) override
	{
		Put( "sky", width, height, pixels, sizeof( float3 ) * (size_t)width * height );
		Put( "skyxform", 0, 0, worldToLight.cell, sizeof( float ) * 16 );
	}
	void SetGeometry( const int meshIdx, const float4* vertexData, const int vertexCount, const int triangleCount, const CoreTri* triangles ) override
	{
		Put( "verts", meshIdx, vertexCount, vertexData, sizeof( float4 ) * (size_t)vertexCount );
		Put( "tris", meshIdx, triangleCount, triangles, sizeof( CoreTri ) * (size_t)triangleCount );
	}
	void SetInstance( const int instanceIdx, const int modelIdx, const mat4& transform ) override
	{
		Put( "instance", instanceIdx, (uint64_t)(int64_t)modelIdx, transform.cell, sizeof( float ) * 16 );
	}
	void FinalizeInstances() override { Put( "finalize", 0, 0, nullptr, 0 ); }

private:
	void Put( const char* tag, uint64_t a, uint64_t b, const void* p, size_t bytes )
	{
		Section s;
		memset( s.tag, 0, sizeof( s.tag ) );
		strncpy( s.tag, tag, 15 );
		s.a = a, s.b = b;
		if (bytes && p) s.data.assign( (const uint8_t*)p, (const uint8_t*)p + bytes );
		sections.push_back( std::move( s ) );
	}
	std::vector<Section> sections;
};

} // namespace

extern "C" __attribute__( ( visibility( "default" ) ) ) CoreAPI_Base* CreateCore() { return new RecorderCore(); }
