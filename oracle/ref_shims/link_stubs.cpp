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
