/* core_api.cpp - the C++ face of the drop-in boundary: a CoreAPI_Base implementation whose 15
   virtuals forward 1:1 to the C ABI (include/lh2b.h), and the exported CreateCore() that the
   reference loader resolves with dlsym (lib/RenderSystem/core_api_base.cpp:97-130; the Optix7 core's
   version is lib/rendercore_optix7/core_api.cpp:18-22).

   Error behaviour: the reference cores print and exit (FatalError). A library should not take the
   host process down, so failures are printed to stderr once per call site and the call becomes a
   no-op; Render on a core that failed to initialise produces no image.

   Presenting: the reference writes into an OpenGL texture through CUDA-GL interop
   (lib/CUDA/shared_host_code/interoptexture.cpp:53). This core renders to a linear RGBA32F device
   buffer. If the host process has OpenGL loaded and the target texture ID is non-zero, the finished
   frame is copied into that texture on the device through the same interop (lh2b_present_gl:
   cudaGraphicsGLRegisterImage + map + cudaMemcpy2DToArray); if the texture cannot be registered
   the frame is uploaded with glTexSubImage2D from host memory instead (resolved at run time; no
   link dependency on OpenGL either way).

   Personas: the reference ships the plain path tracer and the filtering path tracer as two libraries
   (RenderCore_Optix7, RenderCore_Optix7Filter) and RenderSystem sends "filter", "TAA", "clampDirect",
   "clampIndirect" to whichever is loaded (rendersystem.cpp:220-226); the plain core ignores those names
   (rendercore.cpp:746-760). One library serves both here: CreateCore() - loaded as RenderCore_B200 -
   behaves as the plain core and drops those four names, CreateCoreFilter() - what libRenderCore_B200Filter.so's
   CreateCore forwards to - passes them on to the SVGF / TAA chain. LH2B_PERSONA=filter|plain overrides.
*/
#include "../../include/lh2_core_api.h"
#include "../../include/lh2b.h"
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace lh2abi;

namespace
{

typedef void (*glBindTextureFn)(unsigned, unsigned);
typedef void (*glTexSubImage2DFn)(unsigned, int, int, int, int, int, unsigned, unsigned, const void*);

class RenderCoreB200 : public CoreAPI_Base
{
public:
	explicit RenderCoreB200( bool filterPersona ) : filterCore( filterPersona ) {}
	CoreStats GetCoreStats() const override
	{
		CoreStats s = {};
		s.probedTriid = -1;
		if (core) lh2b_get_stats( core, &s );
		return s;
	}
	void Init() override { Check( lh2b_create( &core, -1 ), "Init" ); }
	void SetProbePos( const int2 pos ) override { if (core) Check( lh2b_set_probe_pos( core, pos.x, pos.y ), "SetProbePos" ); }
	void SetTarget( GLTexture* target, const uint spp ) override
	{
		if (!core || !target) return;
		glTexture = target->ID, width = (int)target->width, height = (int)target->height;
		Check( lh2b_set_target( core, width, height, (int)spp ), "SetTarget" );
	}
	void Setting( const char* name, float value ) override
	{
		if (!core || !name) return;
		if (!filterCore && (!strcmp( name, "filter" ) || !strcmp( name, "TAA" ) || !strcmp( name, "clampDirect" ) || !strcmp( name, "clampIndirect" ))) return;
		Check( lh2b_setting( core, name, value ), "Setting" );
	}
	void Render( const ViewPyramid& view, const Convergence converge, bool async ) override
	{
		if (!core) return;
		Check( lh2b_render( core, &view, (int)converge, async ? 1 : 0 ), "Render" );
		if (!async) Present();
	}
	void WaitForRender() override
	{
		if (!core) return;
		Check( lh2b_wait_for_render( core ), "WaitForRender" );
		Present();
	}
	void Shutdown() override { if (core) lh2b_destroy( core ), core = nullptr; }
	void SetTextures( const CoreTexDesc* tex, const int textureCount ) override { if (core) Check( lh2b_set_textures( core, tex, textureCount ), "SetTextures" ); }
	void SetMaterials( CoreMaterial* mat, const int materialCount ) override { if (core) Check( lh2b_set_materials( core, mat, materialCount ), "SetMaterials" ); }
	void SetLights( const CoreLightTri* triLights, const int triLightCount, const CorePointLight* pointLights, const int pointLightCount,
		const CoreSpotLight* spotLights, const int spotLightCount, const CoreDirectionalLight* directionalLights, const int directionalLightCount ) override
	{
		if (core) Check( lh2b_set_lights( core, triLights, triLightCount, pointLights, pointLightCount, spotLights, spotLightCount,
			directionalLights, directionalLightCount ), "SetLights" );
	}
	void SetSkyData( const float3* pixels, const uint w, const uint h, const mat4& worldToLight ) override
	{
		if (core) Check( lh2b_set_sky( core, (const float*)pixels, (int)w, (int)h, worldToLight.cell ), "SetSkyData" );
	}
	void SetGeometry( const int meshIdx, const float4* vertexData, const int vertexCount, const int triangleCount, const CoreTri* triangles ) override
	{
		if (core) Check( lh2b_set_geometry( core, meshIdx, (const float*)vertexData, vertexCount, triangleCount, triangles ), "SetGeometry" );
	}
	void SetInstance( const int instanceIdx, const int modelIdx, const mat4& transform ) override
	{
		if (core) Check( lh2b_set_instance( core, instanceIdx, modelIdx, transform.cell ), "SetInstance" );
	}
	void FinalizeInstances() override { if (core) Check( lh2b_finalize_instances( core ), "FinalizeInstances" ); }
	lh2b_core* Handle() const { return core; }

private:
	void Check( int rc, const char* what ) const
	{
		if (rc != 0) fprintf( stderr, "RenderCore_B200: %s failed: %s\n", what, lh2b_last_error() );
	}
	void Present()
	{
		if (glTexture == 0 || width <= 0) return;
		// first choice: CUDA-GL interop, device to device, like the reference's InteropTexture (interoptexture.cpp:53-61); it needs a current GL
		// context - when there is none (or registration fails once) fall back to an upload from host memory
		if (interop && dlsym( RTLD_DEFAULT, "glBindTexture" ))
		{
			if (lh2b_present_gl( core, glTexture ) == 0) return;
			interop = false;
		}
		static glBindTextureFn bind = (glBindTextureFn)dlsym( RTLD_DEFAULT, "glBindTexture" );
		static glTexSubImage2DFn sub = (glTexSubImage2DFn)dlsym( RTLD_DEFAULT, "glTexSubImage2D" );
		if (!bind || !sub) return;	// headless process: the image stays in the linear buffer
		staging.resize( (size_t)width * height * 4 );
		if (lh2b_read_pixels( core, staging.data() ) != 0) return;
		bind( 0x0DE1 /* GL_TEXTURE_2D */, glTexture );
		sub( 0x0DE1, 0, 0, 0, width, height, 0x1908 /* GL_RGBA */, 0x1406 /* GL_FLOAT */, staging.data() );
	}
	lh2b_core* core = nullptr;
	bool filterCore = false, interop = true;
	unsigned glTexture = 0;
	int width = 0, height = 0;
	std::vector<float> staging;
};

} // namespace

static bool PersonaOverride( bool filter )
{
	const char* e = getenv( "LH2B_PERSONA" );
	if (e && !strcmp( e, "filter" )) return true;
	if (e && !strcmp( e, "plain" )) return false;
	return filter;
}

extern "C" __attribute__( ( visibility( "default" ) ) ) lh2abi::CoreAPI_Base* CreateCore()
{
	return new RenderCoreB200( PersonaOverride( false ) );
}

extern "C" __attribute__( ( visibility( "default" ) ) ) lh2abi::CoreAPI_Base* CreateCoreFilter()
{
	return new RenderCoreB200( PersonaOverride( true ) );
}

/* Accessor for hosts that hold the C++ object but want the C handle (headless read-back, statistics). */
extern "C" lh2b_core* lh2b_handle_of( void* api )
{
	return api ? static_cast<RenderCoreB200*>( static_cast<lh2abi::CoreAPI_Base*>( api ) )->Handle() : nullptr;
}
