/* lh2_oracle_render.h - TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's wavefront path
   tracer (Optix7 core, Lambert+specular BSDF build), one path at a time, plain float arithmetic.

   What is restated, and from where (paths relative to the reference root):
     frame loop, seeds ........ lib/rendercore_optix7/rendercore.cpp:819-938 (RenderImpl), :963-979 (FinalizeRender)
     primary ray generation ... lib/rendercore_optix7/optix/.optix.cu:56-129, lib/RenderSystem/common_functions.h:29-50
     shading .................. lib/rendercore_optix7/kernels/pathtracer.h:54-238
     shading data ............. lib/CUDA/shared_kernel_code/material_shared.h:42-206
     texture fetch ............ lib/CUDA/shared_kernel_code/sampling_shared.h:35-104
     lights, MIS .............. lib/CUDA/shared_kernel_code/lights_shared.h:37-114,174-190,225-313
     BSDF ..................... lib/sharedBSDFs/lambert.h:32-125
     RNG, packing, sky ........ lib/CUDA/shared_kernel_code/tools_shared.h:60-62,101-120,185-216,231-235,324-337
     material conversion ...... lib/rendercore_optix7/rendercore.cpp:508-549 (fp16 colours, 8-bit parameters)
     connect .................. lib/rendercore_optix7/optix/.optix.cu:142-154
     finalize ................. lib/CUDA/shared_kernel_code/finalize_shared.h:29-45
   Ray queries go through the brute-force search of lh2_oracle_geom.h (traversal itself is closed
   source in the reference: parity for it is pinned to the call-site contract, see that header).

   Pinning: oracle/_ref/ holds the reference's own shading headers compiled (a) for the host through
   oracle/ref_cpu_shim.h and (b) for sm_100a; tests/test_oracle_vs_reference*.py compare this
   restatement (and the CUDA core) against them path by path.

   Float semantics: CUDA saturates float->int/uint conversions and maps NaN to 0; F2U/F2I below do the
   same so that index arithmetic (sky lookup, texel addressing, PackNormal) matches the device code.
   The device code is built with -use_fast_math; this code uses libm. Decisions that hinge on the last
   ulp (light pick, Russian roulette) can flip for isolated paths, hence radiance parity is a
   relative-RMSE bound, not bit equality (tolerances are written in the tests).
*/
#pragma once
#include "lh2_oracle_geom.h"
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <vector>

namespace orc
{

struct V3 { float x, y, z; };
static inline V3 v3( float x, float y, float z ) { V3 r = { x, y, z }; return r; }
static inline V3 v3( float s ) { return v3( s, s, s ); }
static inline V3 operator+( V3 a, V3 b ) { return v3( a.x + b.x, a.y + b.y, a.z + b.z ); }
static inline V3 operator-( V3 a, V3 b ) { return v3( a.x - b.x, a.y - b.y, a.z - b.z ); }
static inline V3 operator*( V3 a, V3 b ) { return v3( a.x * b.x, a.y * b.y, a.z * b.z ); }
static inline V3 operator*( V3 a, float s ) { return v3( a.x * s, a.y * s, a.z * s ); }
static inline V3 operator*( float s, V3 a ) { return v3( a.x * s, a.y * s, a.z * s ); }
static inline float dot( V3 a, V3 b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 normalize( V3 a ) { const float il = 1.0f / sqrtf( dot( a, a ) ); return a * il; }
static inline V3 reflect( V3 i, V3 n ) { return i - 2.0f * n * dot( n, i ); }
static inline float sqr( float x ) { return x * x; }

static inline uint32_t F2U( float x ) { if (!(x > 0)) return 0; if (x >= 4294967296.0f) return 0xffffffffu; return (uint32_t)x; }
static inline int32_t F2I( float x ) { if (x != x) return 0; if (x >= 2147483648.0f) return 2147483647; if (x <= -2147483648.0f) return (-2147483647 - 1); return (int32_t)x; }
static inline uint32_t FBits( float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }
static inline float BitsF( uint32_t u ) { float f; memcpy( &f, &u, 4 ); return f; }

/* IEEE half -> float (material colours are stored as fp16, core_settings.h:144) and float -> half, round to nearest even */
static inline float HalfToFloat( uint16_t h )
{
	const uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
	if (e == 0) return (s ? -1.0f : 1.0f) * ldexpf( (float)m, -24 );
	if (e == 31) return m ? NAN : (s ? -INFINITY : INFINITY);
	return (s ? -1.0f : 1.0f) * ldexpf( (float)(m + 1024), (int)e - 25 );
}
static inline uint16_t FloatToHalf( float f )
{
	const uint32_t x = FBits( f ), s = (x >> 16) & 0x8000;
	const int32_t e = (int32_t)((x >> 23) & 255) - 127 + 15;
	uint32_t m = x & 0x7fffff;
	if (((x >> 23) & 255) == 255) return (uint16_t)(s | 0x7c00 | (m ? 0x200 : 0));
	if (e >= 31) return (uint16_t)(s | 0x7c00);
	if (e <= 0)
	{
		if (e < -10) return (uint16_t)s;
		m |= 0x800000;
		const int shift = 14 - e;
		uint32_t r = m >> shift;
		const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
		if (rem > half || (rem == half && (r & 1))) r++;
		return (uint16_t)(s | r);
	}
	uint32_t r = (uint32_t)(e << 10) | (m >> 13);
	const uint32_t rem = m & 0x1fff;
	if (rem > 0x1000 || (rem == 0x1000 && (r & 1))) r++;
	return (uint16_t)(s | r);
}

/* ---- inputs: the reference PODs, addressed as float arrays to stay independent of the product headers ---- */
struct TexDesc { const void* data; uint32_t width, height, flags, pixelCount, firstPixel, mipLevels; int32_t storage; uint32_t pad; };	// CoreTexDesc (40 B)

struct Material		// the CUDAMaterial content (core_settings.h:136-150) after conversion
{
	float color[3], transmittance[3];
	uint32_t flags, params[4];
	struct Map { int w, h; float uscale, vscale, uoffs, voffs; uint32_t addr; } tex0, tex1, nmap0, nmap1, smap, rmap;
	uint32_t baseZ;	// packed (transmittance_g | transmittance_b << 16), read by the normal-map scale quirk (material_shared.h:162-163)
};

struct RenderScene
{
	Scene geo;									// meshes (positions) + instances
	const float* const* coreTris;				// per mesh: CoreTri as 52 floats per triangle
	std::vector<Material> materials;
	const float* triLights; int triLightCount;	// CoreLightTri: 24 floats
	const float* pointLights; int pointLightCount;	// 8 floats
	const float* spotLights; int spotLightCount;	// 12 floats
	const float* dirLights; int dirLightCount;		// 8 floats
	std::vector<float> sky; int skyW, skyH;		// float4 pixels + 64x downscaled copy appended (rendercore.cpp:716-737)
	float worldToSky[16];
	std::vector<uint8_t> argb32, nrm32;			// uchar4 texels, contiguous per storage class
	std::vector<float> argb128;
	const uint32_t* blueNoise;					// expanded table, 5 * 65536 uints (rendercore.cpp:247-254)
};

struct Settings
{
	int w, h, spp;
	int pass;					// samplesTaken before the frame
	uint32_t shift;				// params.shift for this frame
	uint32_t sampleBase;
	uint32_t R0[32];			// per path length (index 1..): RandomUInt(camRNGseed) + pathLength * 91771 (rendercore.cpp:900)
	float geometryEpsilon, clampValue;
	int maxPathLength;
	uint32_t enoughBounces;
	float view[17];				// ViewPyramid
	int bsdfModel;				// 0: lambert.h, 1: disney.h (lh2_oracle_disney.h)
};

enum { S_SPECULAR = 1, S_BOUNCED = 2, S_VIASPECULAR = 4, S_BOUNCEDTWICE = 8 };
static const float PI_ = 3.14159265358979323846264f, INVPI_ = 0.31830988618379067153777f, INV2PI_ = 0.15915494309189533576888f, TWOPI_ = 6.28318530717958647692528f;

static inline uint32_t WangHash( uint32_t s ) { s = (s ^ 61) ^ (s >> 16), s *= 9, s = s ^ (s >> 4), s *= 0x27d4eb2d, s = s ^ (s >> 15); return s; }
static inline uint32_t RandomInt( uint32_t& s ) { s ^= s << 13, s ^= s >> 17, s ^= s << 5; return s; }
static inline float RandomFloat( uint32_t& s ) { return RandomInt( s ) * 2.3283064365387e-10f; }

static inline void BlueNoise4( const uint32_t* bn, int x, int y, int sampleIndex, int dim, float* r )
{
	// tools_shared.h:324-337 / .optix.cu:56-69
	const uint32_t* rank = bn + dim + (x + y * 128) * 8 + 65536 * 3;
	const uint32_t* scr = bn + (dim & 7) + (x + y * 128) * 8 + 65536;
	for (int k = 0; k < 4; k++)
	{
		const int rsi = (sampleIndex ^ rank[k]) & 255;
		const int v = (int)bn[dim + k + rsi * 256];
		r[k] = (0.5f + (float)(int)(v ^ scr[k])) * (1.0f / 256.0f);
	}
}

static inline uint32_t PackNormal( V3 N )
{
	const float f = 65535.0f / fmaxf( sqrtf( 8.0f * N.z + 8.0f ), 0.0001f );
	return F2U( N.x * f + 32767.0f ) + (F2U( N.y * f + 32767.0f ) << 16);
}
static inline V3 UnpackNormal( uint32_t p )
{
	float nx = (float)(p & 65535) * (2.0f / 65535.0f) - 1.0f, ny = (float)(p >> 16) * (2.0f / 65535.0f) - 1.0f;
	float l = -(nx * nx) - (ny * ny) + 1.0f;
	const float z = l;
	l = sqrtf( l ), nx *= l, ny *= l;
	return v3( nx * 2.0f, ny * 2.0f, z * 2.0f - 1.0f );
}
static inline V3 SafeOrigin( V3 O, V3 R, V3 N, float eps ) { return O + N * (dot( N, R ) > 0 ? eps : -eps); }

static inline V3 SampleSky( const RenderScene& s, V3 D, bool small )
{
	const uint32_t w = small ? (uint32_t)s.skyW >> 6 : (uint32_t)s.skyW, h = small ? (uint32_t)s.skyH >> 6 : (uint32_t)s.skyH;
	float phi = atan2f( D.y, D.x );
	if (phi < 0) phi += 2 * PI_;
	const float theta = acosf( fminf( 1.0f, fmaxf( -1.0f, D.z ) ) );
	const uint32_t u = F2U( w * phi * INV2PI_ - 0.5f ), v = F2U( h * theta * INVPI_ - 0.5f );
	const uint32_t idx = u + v * w;
	if (idx >= w * h) return v3( 0 );
	const float* px = s.sky.data() + 4 * ((size_t)idx + (small ? (size_t)s.skyW * s.skyH : 0));
	return v3( px[0], px[1], px[2] );
}

/* ---- textures ---- */
struct F4 { float x, y, z, w; };
static inline F4 Texel( const RenderScene& s, int storage, size_t idx )
{
	if (storage == 1) { const float* p = s.argb128.data() + idx * 4; F4 r = { p[0], p[1], p[2], p[3] }; return r; }
	const uint8_t* p = (storage == 0 ? s.argb32.data() : s.nrm32.data()) + idx * 4;
	const float r = 1.0f / 256.0f;
	F4 o = { p[0] * r, p[1] * r, p[2] * r, p[3] * r };
	return o;
}
static inline F4 FetchTexel( const RenderScene& s, float u, float v, int o, int w, int h, int storage )
{
	const float tx = fmaxf( u + 1000, 0.0f ) * w - 0.5f, ty = fmaxf( v + 1000, 0.0f ) * h - 0.5f;
	const int iu = F2I( tx ) % w, iv = F2I( ty ) % h;
	const float fu = tx - floorf( tx ), fv = ty - floorf( ty );
	const float w0 = (1 - fu) * (1 - fv), w1 = fu * (1 - fv), w2 = (1 - fu) * fv, w3 = 1 - (w0 + w1 + w2);
	const uint32_t iu1 = (uint32_t)(iu + 1) % (uint32_t)w, iv1 = (uint32_t)(iv + 1) % (uint32_t)h;
	const F4 p0 = Texel( s, storage, o + iu + iv * w ), p1 = Texel( s, storage, o + iu1 + iv * w );
	const F4 p2 = Texel( s, storage, o + iu + iv1 * w ), p3 = Texel( s, storage, o + iu1 + iv1 * w );
	F4 r = { p0.x * w0 + p1.x * w1 + p2.x * w2 + p3.x * w3, p0.y * w0 + p1.y * w1 + p2.y * w2 + p3.y * w3,
		p0.z * w0 + p1.z * w1 + p2.z * w2 + p3.z * w3, p0.w * w0 + p1.w * w1 + p2.w * w2 + p3.w * w3 };
	return r;
}
static inline F4 FetchTexelTrilinear( const RenderScene& s, float lambda, float u, float v, int offset, int width, int height )
{
	int level0 = 0, level1 = 0;
	float f = 0;
	if (lambda >= 0)
	{
		level0 = F2I( lambda ) < 4 ? F2I( lambda ) : 4, level1 = level0 + 1 < 4 ? level0 + 1 : 4;
		f = lambda - floorf( lambda );
	}
	const float scale = (float)(width * height) * 1.3333333333f;
	const int o0 = offset + F2I( scale * (1 - BitsF( (uint32_t)(127 - 2 * level0) << 23 )) );
	const int o1 = offset + F2I( scale * (1 - BitsF( (uint32_t)(127 - 2 * level1) << 23 )) );
	const F4 p0 = FetchTexel( s, u, v, o0, width >> level0, height >> level0, 0 ), p1 = FetchTexel( s, u, v, o1, width >> level1, height >> level1, 0 );
	F4 r = { (1 - f) * p0.x + f * p1.x, (1 - f) * p0.y + f * p1.y, (1 - f) * p0.z + f * p1.z, (1 - f) * p0.w + f * p1.w };
	return r;
}

/* ---- shading data ---- */
struct Shading { V3 color, transmittance; int flags; uint32_t params[4]; V3 tint; float lum; };	// tint/lum: material_shared.h:118-119 (principled model)
static inline float Char2Flt( uint32_t a, int s ) { return (float)((a >> s) & 255) * (1.0f / 255.0f); }
static inline float Roughness( const Shading& s ) { return fmaxf( 0.001f, Char2Flt( s.params[0], 24 ) ); }
static inline float Transmission( const Shading& s ) { return Char2Flt( s.params[2], 16 ); }
static inline float Eta( const Shading& s ) { return BitsF( s.params[3] ); }

static inline void TintOf( V3 c, V3& tint, float& lum );
static inline void GetShadingData( const RenderScene& sc, V3 D, float u, float v, float coneWidth, const float* tri, const float* invT,
	Shading& sh, V3& N, V3& iN, V3& fN, V3& T )
{
	const Material& mat = sc.materials[FBits( tri[7] )];
	const uint32_t flags = mat.flags;
	sh.color = v3( mat.color[0], mat.color[1], mat.color[2] ), sh.flags = 0;
	sh.transmittance = v3( mat.transmittance[0], mat.transmittance[1], mat.transmittance[2] );
	memcpy( sh.params, mat.params, 16 );
	TintOf( sh.color, sh.tint, sh.lum );
	N = v3( tri[11], tri[15], tri[19] ), iN = N;
	T = v3( tri[20], tri[21], tri[22] );
	const float w = 1 - (u + v);
	if (flags & (1 << 11)) iN = normalize( w * v3( tri[8], tri[9], tri[10] ) + u * v3( tri[12], tri[13], tri[14] ) + v * v3( tri[16], tri[17], tri[18] ) );
	const V3 A = v3( invT[0], invT[1], invT[2] ), B = v3( invT[4], invT[5], invT[6] ), C = v3( invT[8], invT[9], invT[10] );
	N = normalize( N.x * A + N.y * B + N.z * C );
	iN = normalize( iN.x * A + iN.y * B + iN.z * C );
	fN = iN;
	float tu = 0, tv = 0;
	if (flags & ((1 << 2) | (1 << 9) | (1 << 4) | (1 << 3) | (1 << 7) | (1 << 5)))
		tu = w * tri[0] + u * tri[1] + v * tri[2], tv = w * tri[4] + u * tri[5] + v * tri[6];
	if (flags & (1 << 2))
	{
		const float lambda = tri[31] + log2f( coneWidth * (1.0f / fabsf( dot( D, N ) )) );
		const Material::Map& m = mat.tex0;
		const F4 texel = FetchTexelTrilinear( sc, lambda, m.uscale * (m.uoffs + tu), m.vscale * (m.voffs + tv), m.addr, m.w, m.h );
		if (texel.w < 0.5f) { sh.flags |= 1; return; }
		sh.color = sh.color * v3( texel.x, texel.y, texel.z );
		if (flags & (1 << 9))
		{
			const Material::Map& m1 = mat.tex1;
			const F4 t1 = FetchTexel( sc, m1.uscale * (m1.uoffs + tu), m1.vscale * (m1.voffs + tv), m1.addr, m1.w, m1.h, 0 );
			sh.color = sh.color + v3( t1.x, t1.y, t1.z ) - v3( 0.5f );
		}
	}
	if (flags & (1 << 3))
	{
		const V3 Bt = v3( tri[24], tri[25], tri[26] );
		const Material::Map& m = mat.nmap0;
		const float sb = (float)((mat.baseZ >> 8) & 255) - 128.0f;
		const float n0scale = copysignf( -0.0001f + 0.0001f * expf( 0.1f * fabsf( sb ) ), sb );
		const F4 t0 = FetchTexel( sc, m.uscale * (m.uoffs + tu), m.vscale * (m.voffs + tv), m.addr, m.w, m.h, 2 );
		V3 sn = (v3( t0.x, t0.y, t0.z ) - v3( 0.5f )) * 2.0f;
		sn.x *= n0scale, sn.y *= n0scale;
		if (flags & (1 << 7))
		{
			const Material::Map& m1 = mat.nmap1;
			const float sb1 = (float)((mat.baseZ >> 16) & 255) - 128.0f;
			const float n1scale = copysignf( -0.0001f + 0.0001f * expf( 0.1f * sb1 ), sb1 );
			const F4 t1 = FetchTexel( sc, m1.uscale * (m1.uoffs + tu), m1.vscale * (m1.voffs + tv), m1.addr, m1.w, m1.h, 2 );
			V3 l1 = (v3( t1.x, t1.y, t1.z ) - v3( 0.5f )) * 2.0f;
			l1.x *= n1scale, l1.y *= n1scale;
			sn = sn + l1;
		}
		sn = normalize( sn );
		fN = normalize( sn.x * T + sn.y * Bt + sn.z * iN );
	}
	if (flags & (1 << 5))
	{
		const Material::Map& m = mat.rmap;
		const F4 t = FetchTexel( sc, m.uscale * (m.uoffs + tu), m.vscale * (m.voffs + tv), m.addr, m.w, m.h, 0 );
		sh.params[0] = (sh.params[0] & 0x00ffffff) + ((uint32_t)F2I( t.y * 255.0f ) << 24);
		sh.params[0] = (sh.params[0] & 0xffffff00) + (uint32_t)F2I( t.x * 255.0f );
	}
}

/* ---- helpers shared by host and device in the reference (common_functions.h) ---- */
static inline V3 RandomBarycentrics( float r0 )
{
	const uint32_t uf = F2U( r0 * 4294967296.0f );
	float Ax = 1, Ay = 0, Bx = 0, By = 1, Cx = 0, Cy = 0;
	for (int i = 0; i < 16; ++i)
	{
		const int d = (uf >> (2 * (15 - i))) & 3;
		float Anx, Any, Bnx, Bny, Cnx, Cny;
		switch (d)
		{
		case 0: Anx = (Bx + Cx) * 0.5f, Any = (By + Cy) * 0.5f, Bnx = (Ax + Cx) * 0.5f, Bny = (Ay + Cy) * 0.5f, Cnx = (Ax + Bx) * 0.5f, Cny = (Ay + By) * 0.5f; break;
		case 1: Anx = Ax, Any = Ay, Bnx = (Ax + Bx) * 0.5f, Bny = (Ay + By) * 0.5f, Cnx = (Ax + Cx) * 0.5f, Cny = (Ay + Cy) * 0.5f; break;
		case 2: Anx = (Bx + Ax) * 0.5f, Any = (By + Ay) * 0.5f, Bnx = Bx, Bny = By, Cnx = (Bx + Cx) * 0.5f, Cny = (By + Cy) * 0.5f; break;
		default: Anx = (Cx + Ax) * 0.5f, Any = (Cy + Ay) * 0.5f, Bnx = (Cx + Bx) * 0.5f, Bny = (Cy + By) * 0.5f, Cnx = Cx, Cny = Cy; break;
		}
		Ax = Anx, Ay = Any, Bx = Bnx, By = Bny, Cx = Cnx, Cy = Cny;
	}
	const float rx = (Ax + Bx + Cx) * 0.3333333f, ry = (Ay + By + Cy) * 0.3333333f;
	return v3( rx, ry, 1 - rx - ry );
}
static inline V3 Tangent2World( V3 V, V3 N )
{
	const float sign = copysignf( 1.0f, N.z ), a = -1.0f / (sign + N.z), b = N.x * N.y * a;
	const V3 B = v3( 1.0f + sign * N.x * N.x * a, sign * b, -sign * N.x ), T = v3( b, sign + N.y * N.y * a, -N.y );
	return V.x * T + V.y * B + V.z * N;
}
static inline V3 DiffuseReflectionCosWeighted( float r0, float r1 )
{
	const float term1 = TWOPI_ * r0, term2 = sqrtf( 1 - r1 );
	return v3( cosf( term1 ) * term2, sinf( term1 ) * term2, sqrtf( r1 ) );
}

/* ---- lights ---- */
static inline int LightTotal( const RenderScene& s ) { return (s.triLightCount & 0xffff) + s.pointLightCount + s.spotLightCount + s.dirLightCount; }
static inline float LightPotentials( const RenderScene& s, float* potential, V3 O, V3 N, V3 I, V3 bary )
{
	float sum = 0;
	int lights = 0;
	for (int i = 0; i < (s.triLightCount & 0xffff); i++)
	{
		const float* l = s.triLights + i * 24;
		V3 L = I;
		if (bary.x >= 0) L = v3( bary.x * l[12] + bary.y * l[16] + bary.z * l[20], bary.x * l[13] + bary.y * l[17] + bary.z * l[21], bary.x * l[14] + bary.y * l[18] + bary.z * l[22] );
		L = L - O;
		const float att = 1.0f / dot( L, L );
		L = normalize( L );
		const float LNdotL = fmaxf( 0.0f, -dot( v3( l[4], l[5], l[6] ), L ) ), NdotL = fmaxf( 0.0f, dot( N, L ) );
		const float c = l[3] * LNdotL * NdotL * att;
		potential[lights++] = c, sum += c;
	}
	for (int i = 0; i < s.pointLightCount; i++)
	{
		const float* l = s.pointLights + i * 8;
		const V3 L = v3( l[0], l[1], l[2] ) - O;
		const float NdotL = fmaxf( 0.0f, dot( N, normalize( L ) ) ), att = 1.0f / dot( L, L );
		const float c = l[3] * NdotL * att;
		potential[lights++] = c, sum += c;
	}
	for (int i = 0; i < s.spotLightCount; i++)
	{
		const float* l = s.spotLights + i * 12;
		V3 L = v3( l[0], l[1], l[2] ) - O;
		const float att = 1.0f / dot( L, L );
		L = normalize( L );
		const float d = (fmaxf( 0.0f, -dot( L, v3( l[8], l[9], l[10] ) ) ) - l[7]) / (l[3] - l[7]);
		const float NdotL = fmaxf( 0.0f, dot( N, L ) ), LNdotL = fmaxf( 0.0f, fminf( 1.0f, d ) );
		const float c = (l[4] + l[5] + l[6]) * LNdotL * NdotL * att;
		potential[lights++] = c, sum += c;
	}
	for (int i = 0; i < s.dirLightCount; i++)
	{
		const float* l = s.dirLights + i * 8;
		const float c = l[3] * fmaxf( 0.0f, -(l[0] * N.x + l[1] * N.y + l[2] * N.z) );
		potential[lights++] = c, sum += c;
	}
	return sum;
}
/* More than MAXISLIGHTS (64) lights: the reference's importance sampling would overrun its potential[64] array (lights_shared.h:225-240,
   undefined behaviour). Both the CUDA core and this oracle then take the reference's own other branch (its #else of ISLIGHTS,
   lights_shared.h:256-260): a uniform pick, pickProb = 1 / lightCount. */
static inline float LightPickProb( const RenderScene& s, int idx, V3 O, V3 N, V3 I )
{
	if (LightTotal( s ) > 64) return 1.0f / (float)LightTotal( s );
	float potential[64];
	const float sum = LightPotentials( s, potential, O, N, I, v3( -1 ) );
	if (sum <= 0) return 0;
	return potential[idx] / sum;
}
static inline V3 RandomPointOnLight( const RenderScene& s, float r0, float r1, V3 I, V3 N, float& pickProb, float& lightPdf, V3& lightColor )
{
	const int nTri = s.triLightCount & 0xffff, nPoint = s.pointLightCount, nSpot = s.spotLightCount;
	const int lightCount = LightTotal( s );
	const V3 bary = RandomBarycentrics( r0 );
	int lightIdx = 0;
	if (lightCount > 64) pickProb = 1.0f / (float)lightCount, lightIdx = (int)(r1 * (float)lightCount);
	else
	{
		float potential[64];
		const float sum = LightPotentials( s, potential, I, N, I, bary );
		if (sum <= 0) { lightPdf = 0; return v3( 1 ); }
		r1 *= sum;
		float total = 0;
		for (int i = 0; i < lightCount; i++) { total += potential[i]; if (total >= r1) { lightIdx = i; break; } }
		pickProb = potential[lightIdx] / sum;
	}
	if (lightIdx > lightCount - 1) lightIdx = lightCount - 1;
	if (lightIdx < 0) lightIdx = 0;
	if (lightIdx < nTri)
	{
		const float* l = s.triLights + lightIdx * 24;
		lightColor = v3( l[8], l[9], l[10] );
		const V3 P = v3( bary.x * l[12] + bary.y * l[16] + bary.z * l[20], bary.x * l[13] + bary.y * l[17] + bary.z * l[21], bary.x * l[14] + bary.y * l[18] + bary.z * l[22] );
		V3 L = I - P;
		const float sqDist = dot( L, L );
		L = normalize( L );
		const float LNdotL = L.x * l[4] + L.y * l[5] + L.z * l[6];
		const float reciSolidAngle = sqDist / (l[7] * LNdotL);
		lightPdf = (LNdotL > 0 && dot( L, N ) < 0) ? reciSolidAngle : 0;
		return P;
	}
	if (lightIdx < nTri + nPoint)
	{
		const float* l = s.pointLights + (lightIdx - nTri) * 8;
		const V3 P = v3( l[0], l[1], l[2] ), L = P - I;
		const float sqDist = dot( L, L );
		lightColor = v3( l[4], l[5], l[6] ) * (1.0f / sqDist);
		lightPdf = dot( L, N ) > 0 ? 1 : 0;
		return P;
	}
	if (lightIdx < nTri + nPoint + nSpot)
	{
		const float* l = s.spotLights + (lightIdx - (nTri + nPoint)) * 12;
		const V3 P = v3( l[0], l[1], l[2] );
		V3 L = I - P;
		const float sqDist = dot( L, L );
		L = normalize( L );
		const float d = (fmaxf( 0.0f, L.x * l[8] + L.y * l[9] + L.z * l[10] ) - l[7]) / (l[3] - l[7]);
		const float LNdotL = fminf( 1.0f, d );
		lightPdf = (LNdotL > 0 && dot( L, N ) < 0) ? (sqDist / LNdotL) : 0;
		lightColor = v3( l[4], l[5], l[6] );
		return P;
	}
	const float* l = s.dirLights + (lightIdx - (nTri + nPoint + nSpot)) * 8;
	const V3 L = v3( l[0], l[1], l[2] );
	lightColor = v3( l[4], l[5], l[6] );
	lightPdf = dot( L, N ) < 0 ? 1 : 0;
	return I - 1000.0f * L;
}

/* ---- BSDF (lambert.h) ---- */
static inline float Fr_L( float VDotN, float eio )
{
	if (VDotN < 0.0f) eio = 1.0f / eio, VDotN = fabsf( VDotN );
	const float SinThetaT2 = sqr( eio ) * (1.0f - VDotN * VDotN);
	if (SinThetaT2 > 1.0f) return 1.0f;
	const float LDotN = sqrtf( 1.0f - SinThetaT2 );
	const float r1 = (VDotN - eio * LDotN) / (VDotN + eio * LDotN), r2 = (LDotN - eio * VDotN) / (LDotN + eio * VDotN);
	return 0.5f * (sqr( r1 ) + sqr( r2 ));
}
static inline bool Refract_L( V3 wi, V3 n, float eta, V3& wt )
{
	const float cosThetaI = fabsf( dot( n, wi ) );
	const float sin2ThetaI = fmaxf( 0.0f, 1.0f - cosThetaI * cosThetaI ), sin2ThetaT = eta * eta * sin2ThetaI;
	if (sin2ThetaT >= 1) return false;
	const float cosThetaT = sqrtf( 1.0f - sin2ThetaT );
	wt = eta * (wi * -1.0f) + (eta * cosThetaI - cosThetaT) * n;
	return true;
}
static inline V3 EvaluateBSDF( const Shading& s, V3 iN, V3 wi, float& pdf )
{
	if (Transmission( s ) > 0.999f || Roughness( s ) <= 0.001f) { pdf = 0; return v3( 0 ); }
	pdf = fabsf( dot( wi, iN ) ) * INVPI_;
	return s.color * INVPI_;
}
static inline V3 SampleBSDF( const Shading& s, V3 iN, V3 N, V3 wo, float distance, float r3, float r4, V3& wi, float& pdf, bool& specular )
{
	const float flip = (dot( wo, N ) < 0) ? -1.0f : 1.0f;
	iN = iN * flip;
	specular = true, pdf = 1;
	V3 bsdf;
	const float transmission = Transmission( s );
	if (r4 < transmission)
	{
		const float eio = flip < 0 ? (1.0f / Eta( s )) : Eta( s ), F = Fr_L( dot( iN, wo ), eio );
		const V3 beer = v3( expf( -s.transmittance.x * distance * 2.0f ), expf( -s.transmittance.y * distance * 2.0f ), expf( -s.transmittance.z * distance * 2.0f ) );
		if (r3 < F)
		{
			wi = reflect( wo * -1.0f, iN );
			bsdf = s.color * beer * (1 / fabsf( dot( iN, wi ) ));
		}
		else
		{
			if (!Refract_L( wo, iN, eio, wi )) return v3( 0 );
			return s.color * beer * (1 / fabsf( dot( iN, wi ) ));
		}
	}
	else
	{
		const float pReflect = 1 - Roughness( s );
		if (r3 < pReflect)
		{
			wi = reflect( wo * -1.0f, iN );
			bsdf = s.color * (1.0f / fabsf( dot( iN, wi ) ));
		}
		else
		{
			const float r5 = (r3 - pReflect) / (1 - pReflect), r6 = (r4 - transmission) / (1 - transmission);
			wi = normalize( Tangent2World( DiffuseReflectionCosWeighted( r5, r6 ), iN ) );
			pdf = fmaxf( 0.0f, dot( wi, iN ) ) * INVPI_;
			specular = false;
			bsdf = s.color * INVPI_;
		}
	}
	if (dot( N * flip, wi ) <= 0) pdf = 0;
	return bsdf;
}

} // namespace orc
#include "lh2_oracle_disney.h"
namespace orc
{

static inline void ClampIntensity( V3& c, float clampValue )
{
	const float v = fmaxf( c.x, fmaxf( c.y, c.z ) );
	if (v > clampValue) { const float m = clampValue / v; c.x *= m, c.y *= m, c.z *= m; }
}
static inline void FixNan( V3& a ) { if (!isfinite( a.x + a.y + a.z )) a = v3( 0 ); }

/* ---- primary ray (generateEyeRay / RandomPointOnLens / RayTarget) ---- */
static inline void GeneratePrimary( const RenderScene& sc, const Settings& st, uint32_t pathIdx, V3& O, V3& D )
{
	const float* vw = st.view;
	const V3 pos = v3( vw[0], vw[1], vw[2] ), p1 = v3( vw[3], vw[4], vw[5] ), p2 = v3( vw[6], vw[7], vw[8] ), p3 = v3( vw[9], vw[10], vw[11] );
	const float aperture = vw[12], distortion = vw[16];
	const V3 right = p2 - p1, up = p3 - p1;
	const uint32_t pixels = (uint32_t)st.w * st.h;
	const uint32_t pixelIdx = pathIdx % pixels, seedIdx = pathIdx + st.sampleBase * pixels;
	const uint32_t sampleIdx = seedIdx / pixels + st.pass;
	uint32_t seed = WangHash( seedIdx * 16789 + st.pass * 1791 );
	const int sx = pixelIdx % st.w, sy = pixelIdx / st.w;
	float r[4];
	if (sampleIdx < 64) BlueNoise4( sc.blueNoise, (sx + (st.shift & 127)) & 127, (sy + (st.shift >> 24)) & 127, sampleIdx, 0, r );
	else r[0] = RandomFloat( seed ), r[1] = RandomFloat( seed ), r[2] = RandomFloat( seed ), r[3] = RandomFloat( seed );
	const float blade = (float)F2I( r[0] * 9 );
	float r1 = r[2], r2 = (r[0] - blade * (1.0f / 9.0f)) * 9.0f;
	const float x1 = sinf( blade * PI_ / 4.5f ), y1 = cosf( blade * PI_ / 4.5f );
	const float x2 = sinf( (blade + 1.0f) * PI_ / 4.5f ), y2 = cosf( (blade + 1.0f) * PI_ / 4.5f );
	if ((r1 + r2) > 1) r1 = 1.0f - r1, r2 = 1.0f - r2;
	const float xr = x1 * r1 + x2 * r2, yr = y1 * r1 + y2 * r2;
	O = pos + aperture * (right * xr + up * yr);
	V3 target;
	if (distortion == 0)
	{
		const float u = ((float)sx + r[1]) * (1.0f / st.w), v = ((float)sy + r[3]) * (1.0f / st.h);
		target = p1 + u * right + v * up;
	}
	else
	{
		const float tx = sx / (float)st.w - 0.5f, ty = sy / (float)st.h - 0.5f;
		const float rr = tx * tx + ty * ty;
		const float rq = sqrtf( rr ) * (1.0f + distortion * rr + distortion * rr * rr);
		const float theta = atan2f( tx, ty );
		const float bx = (sinf( theta ) * rq + 0.5f) * st.w, by = (cosf( theta ) * rq + 0.5f) * st.h;
		target = p1 + (bx + r[1]) * (right * (1.0f / (float)st.w)) + (by + r[3]) * (up * (1.0f / (float)st.h));
	}
	D = normalize( target - O );
}

struct PathRecord		// optional per-path trace for path-by-path comparison with the device buffers
{
	uint32_t hit[4];	// primary hit record
	float firstShadow[8];	// first NEE connection: origin.xyz, valid flag, L.xyz, tmax
};

/* Filter mode (Setting "filter"): per-pixel features of the first diffuse vertex and the direct / indirect split, following
   lib/RenderCore_Optix7Filter/kernels/pathtracer.h:44-58,95-125,195-234,286-300 (feature rules) on top of the Optix7 core's
   estimator and random numbers (that is what the CUDA core implements). */
struct FilterArrays { uint32_t* features; float* worldPos; float* deltaDepth; };	// uint4 / float4 / float4 per pixel
static inline uint32_t PackNormal2( V3 N )
{
	const uint32_t x = F2U( (N.x + 1) * 511 ) > 1023u ? 1023u : F2U( (N.x + 1) * 511 ), y = F2U( (N.y + 1) * 511 ) > 1023u ? 1023u : F2U( (N.y + 1) * 511 );
	const uint32_t z = F2U( (N.z + 1) * 511 ) > 1023u ? 1023u : F2U( (N.z + 1) * 511 );
	return (x << 2) + (y << 12) + (z << 22);
}
static inline uint32_t HDRtoRGB32( V3 c ) { return (F2U( 1023.0f * fminf( 1.0f, c.x ) ) << 22) + (F2U( 2047.0f * fminf( 1.0f, c.y ) ) << 11) + F2U( 2047.0f * fminf( 1.0f, c.z ) ); }
static inline V3 RGB32toHDR( uint32_t c ) { return v3( (float)(c >> 22) * (1.0f / 1023.0f), (float)((c >> 11) & 2047) * (1.0f / 2047.0f), (float)(c & 2047) * (1.0f / 2047.0f) ); }
static inline void StoreFeatures( const FilterArrays& fa, uint32_t pathIdx, uint32_t albedo, uint32_t packedNormal, float t, uint32_t isSpecular, uint32_t matid )
{
	uint32_t* f = fa.features + (size_t)pathIdx * 4;
	f[0] = albedo, f[1] = packedNormal, f[2] = FBits( t ), f[3] = (isSpecular << 4) + (matid << 6) + (f[3] & 15);
}
static inline void StoreWorldPos( const FilterArrays& fa, uint32_t pathIdx, V3 P, uint32_t packedNormal )
{
	float* w = fa.worldPos + (size_t)pathIdx * 4;
	w[0] = P.x, w[1] = P.y, w[2] = P.z, w[3] = BitsF( packedNormal );
}

/* One path vertex, exactly the content of the device buffers. */
struct PathState { V3 O; uint32_t data; V3 D; uint32_t packedN; V3 T; float bsdfPdf; };
struct ShadeOut
{
	bool deposit; V3 contribution; uint32_t pixelIdx;		// direct accumulator write (sky / emissive)
	bool shadow; V3 sO, sD; float sTmax; V3 E;				// NEE connection (pathtracer.h:206-209)
	bool extend; PathState next;							// extension ray (pathtracer.h:234-237)
	bool probe; int probeInst, probePrim; float probeDist;
};

/* shadeKernel for one path (pathtracer.h:54-238). hit = (u16|v16<<16, instance, primitive, t bits). */
static inline void ShadeStep( const RenderScene& sc, const Settings& st, int pathLength, const PathState& in, const uint32_t* hit, int probePixelIdx, ShadeOut& out,
	const FilterArrays* fa = nullptr )
{
	memset( &out, 0, sizeof( out ) );
	const uint32_t pixels = (uint32_t)st.w * st.h;
	uint32_t data = in.data;
	const uint32_t pathIdx = data >> 6;
	const bool filter = fa != nullptr;
	const uint32_t pixelIdx = pathIdx % pixels + ((filter && (data & S_BOUNCED)) ? pixels : 0), seedIdx = pathIdx + st.sampleBase * pixels;
	const uint32_t sampleIdx = seedIdx / pixels + st.pass;
	const bool useNEE = LightTotal( sc ) > 0;
	const V3 O = in.O, D = in.D;
	V3 throughput = pathLength == 1 ? v3( 1 ) : in.T;
	const float bsdfPdf = pathLength == 1 ? 1.0f : in.bsdfPdf;
	out.pixelIdx = pixelIdx;
	const bool firstHitToBeStored = filter && (data & S_BOUNCED) == 0 && sampleIdx == 0;
	if (pathLength == 1 && firstHitToBeStored)
	{
		StoreFeatures( *fa, pathIdx, 0, 0, 1e34f, 0, 0 ), StoreWorldPos( *fa, pathIdx, v3( 0 ), 0 );
		memset( fa->deltaDepth + (size_t)pathIdx * 4, 0, 16 );
	}
	auto storeDepthDerivatives = [&]( float depth, const float* tri ) {
		const float* vw = st.view;
		const V3 pos = v3( vw[0], vw[1], vw[2] ), p1 = v3( vw[3], vw[4], vw[5] ), right = v3( vw[6], vw[7], vw[8] ) - p1, up = v3( vw[9], vw[10], vw[11] ) - p1;
		const int x = pathIdx % st.w, y = pathIdx / st.w;
		const V3 triN = v3( tri[11], tri[15], tri[19] ), v0 = v3( tri[32], tri[33], tri[34] );
		const V3 dX = normalize( p1 + (x + 0.5f + 1) * (1.0f / st.w) * right + (y + 0.5f) * (1.0f / st.h) * up - pos );
		const V3 dY = normalize( p1 + (x + 0.5f) * (1.0f / st.w) * right + (y + 0.5f + 1) * (1.0f / st.h) * up - pos );
		const float num = dot( v0 - pos, triN );
		float* d = fa->deltaDepth + (size_t)pathIdx * 4;
		d[0] = d[1] = 0, d[2] = num / dot( triN, dX ) - depth, d[3] = num / dot( triN, dY ) - depth;
	};
	const int prim = (int)hit[2], instIdx = (int)hit[1];
	if (prim == -1)
	{
		const float* m = sc.worldToSky;
		const V3 tD = v3( -(m[0] * D.x + m[1] * D.y + m[2] * D.z), -(m[4] * D.x + m[5] * D.y + m[6] * D.z), -(m[8] * D.x + m[9] * D.y + m[10] * D.z) );
		V3 contribution = throughput * SampleSky( sc, tD, !filter && (data & S_BOUNCED) != 0 ) * (1.0f / bsdfPdf);
		ClampIntensity( contribution, st.clampValue );
		FixNan( contribution );
		out.deposit = true, out.contribution = contribution;
		if (firstHitToBeStored)
		{
			const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( D * -1.0f ) + isSpecular;
			StoreFeatures( *fa, pathIdx, HDRtoRGB32( contribution ), packedNormal, BitsF( hit[3] ), isSpecular, 0 );
			StoreWorldPos( *fa, pathIdx, O + 50000 * D, packedNormal );
			memset( fa->deltaDepth + (size_t)pathIdx * 4, 0, 16 );
		}
		return;
	}
	const float hu = (float)(hit[0] & 65535) * (1.0f / 65535.0f), hv = (float)(hit[0] >> 16) * (1.0f / 65535.0f);
	const float ht = BitsF( hit[3] );
	// picking: the reference lets every sample of the probed pixel write (pathtracer.h:97-102, a race at spp > 1); the defined outcome we
	// restate is "the frame's first sample of that pixel" (pathIdx < w*h), which is what the CUDA core writes
	if ((int)pathIdx == probePixelIdx && pathLength == 1 && (!filter || sampleIdx == 0)) out.probe = true, out.probeInst = instIdx, out.probePrim = prim, out.probeDist = ht;
	const float* tri = sc.coreTris[sc.geo.instances[instIdx].mesh] + (size_t)prim * 52;
	const float* invT = sc.geo.inverses + instIdx * 12;
	Shading sh;
	V3 N, iN, fN, T;
	const V3 I = O + ht * D;
	const float spreadAngle = st.view[13];
	GetShadingData( sc, D, hu, hv, spreadAngle * ht, tri, invT, sh, N, iN, fN, T );
	if (filter && !(sh.flags & 1)) sh.color = v3( fmaxf( 0.05f, sh.color.x ), fmaxf( 0.05f, sh.color.y ), fmaxf( 0.05f, sh.color.z ) );	// FILTERINGCORE
	uint32_t seed = WangHash( seedIdx * 17 + st.R0[pathLength] );
	if (sh.flags & 1)
	{
		if (pathLength == st.maxPathLength)
		{
			if (firstHitToBeStored)
			{
				const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( N ) + isSpecular;
				StoreFeatures( *fa, pathIdx, 0, packedNormal, ht, isSpecular, 0 ), StoreWorldPos( *fa, pathIdx, I, packedNormal );
			}
		}
		else
		{
			out.extend = true, out.next = in;
			out.next.O = I + D * st.geometryEpsilon;
			if (pathLength == 1) out.next.T = v3( 1 ), out.next.bsdfPdf = 1;
			if (!isfinite( out.next.T.x + out.next.T.y + out.next.T.z )) out.next.T = v3( 0 );
		}
		return;
	}
	if (sh.color.x > 1.0f || sh.color.y > 1.0f || sh.color.z > 1.0f)
	{
		const float DdotNL = -dot( D, N );
		if (DdotNL > 0)
		{
			V3 contribution = v3( 0 );
			if (pathLength == 1 || (data & S_SPECULAR) || !useNEE) contribution = sh.color;
			else
			{
				const V3 lastN = UnpackNormal( in.packedN );
				const float area = tri[23];
				const int ltriIdx = (int)FBits( tri[3] );
				const float lightPdf = (ht * ht) / (fabsf( dot( D, N ) ) * area);
				const float pickProb = LightPickProb( sc, ltriIdx, O, lastN, I );
				if ((bsdfPdf + lightPdf * pickProb) > 0) contribution = throughput * sh.color * (1.0f / (bsdfPdf + lightPdf * pickProb));
			}
			ClampIntensity( contribution, st.clampValue );
			FixNan( contribution );
			out.deposit = true, out.contribution = contribution;
		}
		if (firstHitToBeStored)
		{
			const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( N ) + isSpecular;
			StoreFeatures( *fa, pathIdx, HDRtoRGB32( sh.color ), packedNormal, ht, isSpecular, 0 ), StoreWorldPos( *fa, pathIdx, I, packedNormal );
			storeDepthDerivatives( ht, tri );
		}
		return;
	}
	if (!filter && (data & S_BOUNCED)) sh.params[0] |= 255u << 24;
	const float roughness = Roughness( sh );
	if (roughness <= 0.001f || Transmission( sh ) > (filter ? 0.999f : 0.5f)) data |= S_SPECULAR; else data &= ~(uint32_t)S_SPECULAR;
	const float faceDir = (dot( D, N ) > 0) ? -1.0f : 1.0f;
	if (firstHitToBeStored)
	{
		if (data & S_SPECULAR) fa->features[(size_t)pathIdx * 4] = HDRtoRGB32( sh.color );
		else
		{
			V3 albedo = sh.color;
			if (data & S_VIASPECULAR) albedo = albedo * RGB32toHDR( fa->features[(size_t)pathIdx * 4] );
			const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( fN * -1.0f ) + isSpecular;	// sic: always -fN (Optix7Filter pathtracer.h:229)
			StoreFeatures( *fa, pathIdx, HDRtoRGB32( albedo ), packedNormal, ht, isSpecular, 0 ), StoreWorldPos( *fa, pathIdx, I, packedNormal );
			storeDepthDerivatives( ht, tri );
		}
	}
	if (faceDir == 1) sh.transmittance = v3( 0 );
	throughput = throughput * (1.0f / bsdfPdf);
	float r4[4];
	if (sampleIdx < 64)
		BlueNoise4( sc.blueNoise, ((seedIdx % st.w) + (st.shift & 127)) & 127, ((seedIdx / st.w) + (st.shift >> 24)) & 127, sampleIdx, 4 * pathLength - 4, r4 );
	else r4[0] = RandomFloat( seed ), r4[1] = RandomFloat( seed ), r4[2] = RandomFloat( seed ), r4[3] = RandomFloat( seed );
	if ((data & S_SPECULAR) == 0 && useNEE)
	{
		float pickProb = 0, lightPdf = 0;
		V3 lightColor = v3( 0 );
		V3 L = RandomPointOnLight( sc, r4[0], r4[1], I, fN * faceDir, pickProb, lightPdf, lightColor ) - I;
		const float dist = sqrtf( dot( L, L ) );
		L = L * (1.0f / dist);
		const float NdotL = dot( L, fN * faceDir );
		if (NdotL > 0 && lightPdf > 0)
		{
			float lobePdf;
			// BSDF_HAS_PURE_SPECULARS (lambert.h:30) scales the NEE term by ROUGHNESS; the principled model does not (pathtracer.h:194-198)
			const V3 f = st.bsdfModel == 0 ? EvaluateBSDF( sh, fN, L, lobePdf ) * roughness : EvaluatePrincipled( MakePrincipled( sh ), fN, T, D * -1.0f, L, lobePdf );
			if (lobePdf > 0)
			{
				V3 contribution = throughput * f * lightColor * (NdotL / (pickProb * lightPdf + lobePdf));
				FixNan( contribution );
				ClampIntensity( contribution, st.clampValue );
				out.shadow = true, out.sO = SafeOrigin( I, L, N, st.geometryEpsilon ), out.sD = L;
				out.sTmax = dist - 2 * st.geometryEpsilon, out.E = contribution;
			}
		}
	}
	if (data & (filter ? (uint32_t)S_BOUNCED : st.enoughBounces)) return;
	if (pathLength == st.maxPathLength)
	{
		if (firstHitToBeStored)
		{
			const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( N ) + isSpecular;
			StoreFeatures( *fa, pathIdx, 0, packedNormal, ht, isSpecular, 0 ), StoreWorldPos( *fa, pathIdx, I, packedNormal );
		}
		return;
	}
	V3 R;
	float newPdf;
	bool specular = false;
	const float r5 = RandomFloat( seed );	// third random number of the reference SampleBSDF call (unused by the Lambert model)
	const V3 bsdf = st.bsdfModel == 0 ? SampleBSDF( sh, fN, N, D * -1.0f, ht, r4[2], r4[3], R, newPdf, specular ) :
		SamplePrincipled( MakePrincipled( sh ), fN, N, T, D * -1.0f, ht, r4[2], r4[3], r5, R, newPdf, specular );
	if (newPdf < 0.0001f || newPdf != newPdf) return;
	if (specular) data |= S_SPECULAR;
	const float p = (filter || (data & S_SPECULAR) || ((data & S_BOUNCED) == 0)) ? 1 : fminf( 1.0f, fmaxf( fmaxf( bsdf.x, bsdf.y ), bsdf.z ) );
	if (p < RandomFloat( seed )) return;
	throughput = throughput * (1 / p);
	const uint32_t packedNormal = PackNormal( fN * faceDir );
	if (!(data & S_SPECULAR)) data |= (data & S_BOUNCED) ? S_BOUNCEDTWICE : S_BOUNCED; else data |= S_VIASPECULAR;
	out.extend = true;
	out.next.O = SafeOrigin( I, R, N, st.geometryEpsilon ), out.next.data = data;
	out.next.D = R, out.next.packedN = packedNormal;
	FixNan( throughput );
	out.next.T = throughput * bsdf * fabsf( dot( fN, R ) ), out.next.bsdfPdf = newPdf;
}

/* Trace one complete path: generate, then (closest hit, shade, connect) per path length. Deposits into accum
   (4 doubles per pixel). rayCounts[0] += extension rays, [1] += shadow rays. */
static inline void TracePath( const RenderScene& sc, const Settings& st, uint32_t pathIdx, double* accum, uint32_t* rayCounts, PathRecord* rec, const FilterArrays* fa = nullptr )
{
	PathState ps;
	memset( &ps, 0, sizeof( ps ) );
	GeneratePrimary( sc, st, pathIdx, ps.O, ps.D );
	ps.data = (pathIdx << 6) + S_SPECULAR, ps.T = v3( 1 ), ps.bsdfPdf = 1;
	for (int pathLength = 1; pathLength <= st.maxPathLength; pathLength++)
	{
		Hit h;
		const float o[3] = { ps.O.x, ps.O.y, ps.O.z }, d[3] = { ps.D.x, ps.D.y, ps.D.z };
		const bool hit = ClosestHit( sc.geo, o, d, 0.0f, 1e34f, h );
		rayCounts[0]++;
		uint32_t rec4[4];
		PackHit( hit, h, rec4 );
		if (rec && pathLength == 1) memcpy( rec->hit, rec4, 16 );
		ShadeOut so;
		ShadeStep( sc, st, pathLength, ps, rec4, -1, so, fa );
		double* px = accum + (size_t)so.pixelIdx * 4;
		if (so.deposit) px[0] += so.contribution.x, px[1] += so.contribution.y, px[2] += so.contribution.z;
		if (so.shadow)
		{
			rayCounts[1]++;
			const float s3[3] = { so.sO.x, so.sO.y, so.sO.z }, l3[3] = { so.sD.x, so.sD.y, so.sD.z };
			if (rec && rec->firstShadow[3] == 0)
				rec->firstShadow[0] = s3[0], rec->firstShadow[1] = s3[1], rec->firstShadow[2] = s3[2], rec->firstShadow[3] = 1,
				rec->firstShadow[4] = l3[0], rec->firstShadow[5] = l3[1], rec->firstShadow[6] = l3[2], rec->firstShadow[7] = so.sTmax;
			if (!Occluded( sc.geo, s3, l3, 0.0f, so.sTmax )) px[0] += so.E.x, px[1] += so.E.y, px[2] += so.E.z, px[3] += 1;
		}
		if (!so.extend) return;
		ps = so.next;
	}
}

} // namespace orc
