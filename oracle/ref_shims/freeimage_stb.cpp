/* freeimage_stb.cpp - TEST INFRASTRUCTURE ONLY. See FreeImage.h in this directory. stb_image's implementation is already
   part of the reference build (lib/RenderSystem/host_meshloaders.cpp defines STB_IMAGE_IMPLEMENTATION); only declarations here. */
#include "FreeImage.h"
#include "stb_image.h"
#include <stdlib.h>
#include <string.h>
#include <strings.h>

struct FIBITMAP { unsigned w, h, bpp; BYTE* bits; };

static FREE_IMAGE_FORMAT ByExtension( const char* fn )
{
	const char* dot = strrchr( fn, '.' );
	if (!dot) return FIF_UNKNOWN;
	if (!strcasecmp( dot, ".png" )) return FIF_PNG;
	if (!strcasecmp( dot, ".jpg" ) || !strcasecmp( dot, ".jpeg" )) return FIF_JPEG;
	if (!strcasecmp( dot, ".bmp" )) return FIF_BMP;
	if (!strcasecmp( dot, ".tga" )) return FIF_TARGA;
	if (!strcasecmp( dot, ".hdr" )) return FIF_HDR;
	return FIF_UNKNOWN;
}

extern "C" {

FREE_IMAGE_FORMAT FreeImage_GetFileType( const char* filename, int )
{
	int w, h, c;
	if (!stbi_info( filename, &w, &h, &c )) return FIF_UNKNOWN;
	return stbi_is_hdr( filename ) ? FIF_HDR : (ByExtension( filename ) == FIF_UNKNOWN ? FIF_PNG : ByExtension( filename ));
}

FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename( const char* filename ) { return ByExtension( filename ); }

static void FlipRows( BYTE* bits, unsigned pitch, unsigned h )
{
	BYTE* tmp = (BYTE*)malloc( pitch );
	for (unsigned y = 0; y < h / 2; y++)
	{
		memcpy( tmp, bits + (size_t)y * pitch, pitch );
		memcpy( bits + (size_t)y * pitch, bits + (size_t)(h - 1 - y) * pitch, pitch );
		memcpy( bits + (size_t)(h - 1 - y) * pitch, tmp, pitch );
	}
	free( tmp );
}

FIBITMAP* FreeImage_Load( FREE_IMAGE_FORMAT, const char* filename, int )
{
	int w, h, c;
	FIBITMAP* b = (FIBITMAP*)calloc( 1, sizeof( FIBITMAP ) );
	if (stbi_is_hdr( filename ))
	{
		float* f = stbi_loadf( filename, &w, &h, &c, 3 );
		if (!f) { free( b ); return 0; }
		b->w = w, b->h = h, b->bpp = 96, b->bits = (BYTE*)f;
	}
	else
	{
		stbi_uc* p = stbi_load( filename, &w, &h, &c, 4 );
		if (!p) { free( b ); return 0; }
		b->w = w, b->h = h, b->bpp = 32, b->bits = p;
	}
	FlipRows( b->bits, b->w * (b->bpp / 8), b->h );	// FreeImage keeps scanline 0 at the bottom
	return b;
}

/* real FreeImage returns a new bitmap for LDR input and null for float formats (the reference then uses the original and
   unloads both only when bpp == 32, host_texture.cpp:215,257) */
FIBITMAP* FreeImage_ConvertTo32Bits( FIBITMAP* dib )
{
	if (!dib || dib->bpp != 32) return 0;
	FIBITMAP* b = (FIBITMAP*)calloc( 1, sizeof( FIBITMAP ) );
	*b = *dib;
	b->bits = (BYTE*)malloc( (size_t)dib->w * dib->h * 4 );
	memcpy( b->bits, dib->bits, (size_t)dib->w * dib->h * 4 );
	return b;
}

void FreeImage_Unload( FIBITMAP* dib ) { if (dib) { free( dib->bits ); free( dib ); } }
unsigned FreeImage_GetWidth( FIBITMAP* dib ) { return dib->w; }
unsigned FreeImage_GetHeight( FIBITMAP* dib ) { return dib->h; }
unsigned FreeImage_GetPitch( FIBITMAP* dib ) { return dib->w * (dib->bpp / 8); }
unsigned FreeImage_GetBPP( FIBITMAP* dib ) { return dib->bpp; }
BYTE* FreeImage_GetBits( FIBITMAP* dib ) { return dib->bits; }
BYTE* FreeImage_GetScanLine( FIBITMAP* dib, int scanline ) { return dib->bits + (size_t)scanline * FreeImage_GetPitch( dib ); }
BOOL FreeImage_Invert( FIBITMAP* dib )
{
	if (dib->bpp != 32) return 0;
	for (size_t i = 0; i < (size_t)dib->w * dib->h * 4; i++) dib->bits[i] = 255 - dib->bits[i];
	return 1;
}

}
