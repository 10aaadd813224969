/* fake_gl.cpp - TEST INFRASTRUCTURE ONLY. The "window system" of the headless reference host (oracle/ref_tinyapp_host.cpp):
   the two OpenGL entry points our core resolves at run time to present a frame into the host's GL texture
   (lighthouse2_b200/csrc/core_api.cpp: Present), exported from the executable so that the frame lands in memory here.
   Own translation unit: glad.h turns these names into macros. */
#include <vector>
#include <stddef.h>

static std::vector<float> lastFrame;
static int lastW = 0, lastH = 0, presented = 0;

extern "C" __attribute__( ( visibility( "default" ) ) ) void glBindTexture( unsigned, unsigned ) {}
extern "C" __attribute__( ( visibility( "default" ) ) ) void glTexSubImage2D( unsigned, int, int, int, int w, int h, unsigned, unsigned, const void* pixels )
{
	lastW = w, lastH = h, presented++;
	lastFrame.assign( (const float*)pixels, (const float*)pixels + (size_t)w * h * 4 );
}
extern "C" const float* FakeGL_LastFrame( int* w, int* h, int* count )
{
	*w = lastW, *h = lastH, *count = presented;
	return lastFrame.empty() ? nullptr : lastFrame.data();
}
