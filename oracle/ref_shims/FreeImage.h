/* FreeImage.h - TEST INFRASTRUCTURE ONLY (shim). The reference's RenderSystem (lib/RenderSystem/host_texture.cpp:209-257,
   host_skydome.cpp:106-131) loads images through FreeImage, whose library is a Windows binary in the reference tree.
   oracle/Makefile compiles the reference RenderSystem where it lies and puts this directory in front of the include path;
   the dozen entry points it uses are implemented in freeimage_stb.cpp on top of stb_image (the copy the reference vendors
   under lib/tinygltf, compiled from there). Semantics kept: bottom-up scanlines, 32-bit LDR / 96-bit float HDR bitmaps. */
#pragma once
#include <stdint.h>
typedef uint8_t BYTE;
typedef int32_t BOOL;
struct FIBITMAP;
enum FREE_IMAGE_FORMAT { FIF_UNKNOWN = -1, FIF_BMP = 0, FIF_JPEG = 2, FIF_PNG = 13, FIF_TARGA = 17, FIF_HDR = 26 };
/* byte order of a 32-bit pixel as stored by this shim (the reference indexes with these macros, host_texture.cpp:234) */
#define FI_RGBA_RED 0
#define FI_RGBA_GREEN 1
#define FI_RGBA_BLUE 2
#define FI_RGBA_ALPHA 3
extern "C" {
FREE_IMAGE_FORMAT FreeImage_GetFileType( const char* filename, int size );
FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename( const char* filename );
FIBITMAP* FreeImage_Load( FREE_IMAGE_FORMAT fif, const char* filename, int flags = 0 );
FIBITMAP* FreeImage_ConvertTo32Bits( FIBITMAP* dib );
void FreeImage_Unload( FIBITMAP* dib );
unsigned FreeImage_GetWidth( FIBITMAP* dib );
unsigned FreeImage_GetHeight( FIBITMAP* dib );
unsigned FreeImage_GetPitch( FIBITMAP* dib );
unsigned FreeImage_GetBPP( FIBITMAP* dib );
BYTE* FreeImage_GetBits( FIBITMAP* dib );
BYTE* FreeImage_GetScanLine( FIBITMAP* dib, int scanline );
BOOL FreeImage_Invert( FIBITMAP* dib );
}
