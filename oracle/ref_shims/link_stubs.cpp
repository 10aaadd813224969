/* link_stubs.cpp - TEST INFRASTRUCTURE ONLY. Symbols the reference's lib/platform/system.cpp and lib/RenderSystem/host_scene.cpp
   reference but a headless Linux build has no library for: FreeType (text rendering, system.cpp:133-163), the GL shader wrapper
   (platform.cpp) and the PBRT scene parser (lib/RenderSystem/materials/pbrt). None is reachable from the tinyapp scene path;
   each aborts loudly if it is ever called. */
#include "platform.h"
#include <cstdio>
#include <cstdlib>

static void Unreachable( const char* what ) { fprintf( stderr, "oracle/_ref host: %s is not available in the headless build\n", what ); abort(); }

extern "C" {
int FT_Init_FreeType( void* ) { Unreachable( "FreeType" ); return 1; }
int FT_New_Face( void*, const char*, long, void* ) { Unreachable( "FreeType" ); return 1; }
int FT_Set_Pixel_Sizes( void*, unsigned, unsigned ) { Unreachable( "FreeType" ); return 1; }
int FT_Load_Char( void*, unsigned long, int ) { Unreachable( "FreeType" ); return 1; }
int FT_Done_Face( void* ) { Unreachable( "FreeType" ); return 1; }
int FT_Done_FreeType( void* ) { Unreachable( "FreeType" ); return 1; }
}

void PBRTInit() { Unreachable( "the PBRT parser" ); }
void ParsePBRTScene( std::string ) { Unreachable( "the PBRT parser" ); }

namespace lighthouse2
{
Shader::Shader( const char*, const char* ) { Unreachable( "Shader" ); }
void Shader::Bind() { Unreachable( "Shader" ); }
}

/* The reference's skinning code (lib/RenderSystem/host_mesh.cpp:777-877) uses aligned AVX loads (_mm256_load_ps) on
   std::vector<mat4> storage, which the MSVC runtime happens to satisfy and glibc's 16-byte malloc alignment does not: every
   allocation of the headless host is 64-byte aligned. */
#include <new>
void* operator new( std::size_t n )
{
	void* p = nullptr;
	if (posix_memalign( &p, 64, n ? n : 1 ) != 0) throw std::bad_alloc();
	return p;
}
void* operator new[]( std::size_t n ) { return operator new( n ); }
void operator delete( void* p ) noexcept { free( p ); }
void operator delete[]( void* p ) noexcept { free( p ); }
void operator delete( void* p, std::size_t ) noexcept { free( p ); }
void operator delete[]( void* p, std::size_t ) noexcept { free( p ); }
