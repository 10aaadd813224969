/* ref_tinyapp_host.cpp - TEST INFRASTRUCTURE ONLY. The reference's tinyapp (apps/tinyapp/main.cpp:33-42,72-112) without a
   window, linked against the reference's own RenderSystem + platform sources compiled where they lie under /root/reference
   (oracle/Makefile -> oracle/_ref/tinyapp_ref_host). It is the end-to-end drop-in check BASELINE.json configs[0] describes:

       RenderAPI::CreateRenderAPI( <core> )      the unmodified reference loader: dlopen + dlsym "CreateCore"
       DeserializeCamera( camera.xml ), PrepareScene()   pica glTF scene + light quad + legocar.obj, as tinyapp does
       SetTarget( GLTexture 640x360, 1 spp ), SynchronizeSceneData(), Render( Restart )  x frames, car animated per frame

   with <core> = our libRenderCore_B200.so (GPU box) or oracle/_ref/libRenderCore_Recorder.so (CPU: captures the scene for
   the CPU oracle). The window system is replaced by two functions this executable exports, glBindTexture / glTexSubImage2D:
   our core presents into "the GL texture" by resolving exactly these at run time (csrc/core_api.cpp: Present), so the frame
   arrives here and is written to <out> as raw float32 RGBA, rows top to bottom as the core produced them.

   usage: tinyapp_ref_host <core name or path> <shareddata dir/> <camera.xml> <out.bin> [frames=1] [width=640] [height=360] [spp=1] [anim.glb]

   anim.glb (optional, e.g. CesiumMan.glb): additionally loads that skinned / animated glTF scene the way apps/imguiapp/main.cpp:90
   does and advances every animation by 0.05 s per frame (apps/imguiapp/main.cpp:257, viewerapp/main.cpp:189): RenderSystem then
   re-skins the mesh on the host and re-sends it with SetGeometry (same triangle count) every frame - the core's refit path.
*/
#include "platform.h"
#include "rendersystem.h"
#include <cstdio>
#include <vector>
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

static void CrashHandler( int sig )
{
	void* frames[48];
	const int n = backtrace( frames, 48 );
	fprintf( stderr, "tinyapp_ref_host: signal %d\n", sig );
	backtrace_symbols_fd( frames, n, 2 );
	_exit( 128 + sig );
}

extern "C" const float* FakeGL_LastFrame( int* w, int* h, int* count );	// ref_shims/fake_gl.cpp: glBindTexture / glTexSubImage2D

/* the only pieces of lib/platform/platform.cpp a headless host needs: a GLTexture is its ID and size (system.h:236-237) */
namespace lighthouse2
{
GLTexture::GLTexture( uint w, uint h, uint ) { ID = 1, width = w, height = h; }
GLTexture::~GLTexture() {}
}

int main( int argc, char** argv )
{
	if (argc < 5) { fprintf( stderr, "usage: %s <core> <shareddata dir/> <camera.xml> <out.bin> [frames] [w] [h] [spp]\n", argv[0] ); return 2; }
	signal( SIGSEGV, CrashHandler ), signal( SIGABRT, CrashHandler );
	const char* coreName = argv[1];
	const std::string data = argv[2];
	const int frames = argc > 5 ? atoi( argv[5] ) : 1, w = argc > 6 ? atoi( argv[6] ) : 640, h = argc > 7 ? atoi( argv[7] ) : 360;
	const int spp = argc > 8 ? atoi( argv[8] ) : 1;
	const char* animFile = argc > 9 ? argv[9] : nullptr;
	RenderAPI* renderer = RenderAPI::CreateRenderAPI( coreName );
	renderer->DeserializeCamera( argv[3] );
	// PrepareScene of the tinyapp
	renderer->AddScene( "scene.gltf", (data + "pica/").c_str() );
	renderer->SetNodeTransform( renderer->FindNode( "RootNode (gltf orientation matrix)" ), mat4::RotateX( -PI / 2 ) );
	const int lightMat = renderer->AddMaterial( make_float3( 100, 100, 80 ) );
	const int lightQuad = renderer->AddQuad( make_float3( 0, -1, 0 ), make_float3( 0, 26.0f, 0 ), 6.9f, 6.9f, lightMat );
	renderer->AddInstance( lightQuad );
	const int car = renderer->AddInstance( renderer->AddMesh( "legocar.obj", data.c_str(), 10.0f ) );
	if (animFile) renderer->AddScene( animFile, data.c_str(), mat4::Translate( -14, 6, 24 ) * mat4::Scale( 3.0f ) );	// in front of the camera
	GLTexture* target = new GLTexture( w, h, GLTexture::FLOAT );
	renderer->SetTarget( target, spp );
	float r = 0;
	for (int f = 0; f < frames; f++)
	{
		renderer->SynchronizeSceneData();
		renderer->Render( Restart );
		// the tinyapp's rigid animation of the car, applied to the next frame as there
		mat4 M = mat4::RotateY( r * 2.0f ) * mat4::RotateZ( 0.2f * sinf( r * 8.0f ) ) * mat4::Translate( make_float3( 0, 5, 0 ) );
		renderer->SetNodeTransform( car, M );
		if ((r += 0.025f * 0.3f) > 2 * PI) r -= 2 * PI;
		if (animFile) for (int i = 0; i < renderer->AnimationCount(); i++) renderer->UpdateAnimation( i, 0.05f );
	}
	const CoreStats stats = renderer->GetCoreStats();
	int lastW = 0, lastH = 0, presented = 0;
	const float* lastFrame = FakeGL_LastFrame( &lastW, &lastH, &presented );
	FILE* out = fopen( argv[4], "wb" );
	if (out && lastFrame) fwrite( lastFrame, sizeof( float ), (size_t)lastW * lastH * 4, out );
	if (out) fclose( out );
	printf( "{\"frames\": %d, \"presented\": %d, \"width\": %d, \"height\": %d, \"primaryRays\": %u, \"totalRays\": %u, \"shadowRays\": %u, "
		"\"probedInst\": %d, \"probedTri\": %d, \"renderTime\": %f, \"animations\": %d}\n", frames, presented, lastW, lastH, stats.primaryRayCount, stats.totalRays,
		stats.totalShadowRays, stats.probedInstid, stats.probedTriid, stats.renderTime, renderer->AnimationCount() );
	renderer->Shutdown();
	return 0;
}
