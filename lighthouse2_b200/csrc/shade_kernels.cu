/* shade_kernels.cu - the shade and finalize stages of the wavefront loop (sm_100a, built with
   --use_fast_math like the reference: lib/CUDA/shared_host_code/cudatools.h:171).

   shadeKernel restates lib/rendercore_optix7/kernels/pathtracer.h:54-238 with the helpers it pulls in:
     material fetch ....... lib/CUDA/shared_kernel_code/material_shared.h:42-206
     texture fetch ........ lib/CUDA/shared_kernel_code/sampling_shared.h:35-104
     light sampling / MIS . lib/CUDA/shared_kernel_code/lights_shared.h:37-114,174-190,225-313
     BSDF ................. lib/sharedBSDFs/lambert.h:32-125 (BSDF_HAS_PURE_SPECULARS)
     RNG, packing, sky .... lib/CUDA/shared_kernel_code/tools_shared.h:60-62,101-120,196-216,231-235,324-337
     shared host/device ... lib/RenderSystem/common_functions.h:52-116
   Differences that are design, not behaviour:
     - ping-pong path-state sets instead of in-place compaction (no read/write race);
     - warp-aggregated compaction (one atomic per warp and stream, ballot + popc prefix);
     - red.global.add.v4.f32 for the accumulator instead of a racy read-modify-write;
     - per-path-length counters on the device, so the host never reads back between bounces.
*/
#include "kernels.h"
#include "render_types.h"
#include "common.cuh"
#include <cuda_fp16.h>

/* Second build of this file (csrc/Makefile: build/shade_kernels_precise.o, -DLH2B_PRECISE, no --use_fast_math, -fmad=false): the same
   kernels with IEEE division / sqrt, libm-accurate transcendentals and no FMA contraction, under other names. Setting "preciseMath" 1
   selects it - not for speed: it is what lets the end-to-end parity tests against the CPU oracle (libm, no contraction) use tight bounds. */
#ifdef LH2B_PRECISE
#define LaunchShade LaunchShadePrecise
#define LaunchFinalize LaunchFinalizePrecise
#define shadeKernel shadeKernelPrecise
#define finalizeKernel finalizeKernelPrecise
#define __expf expf
#define __sincosf sincosf
#endif

namespace lh2b
{

#define PI_F      3.14159265358979323846264f
#define INVPI_F   0.31830988618379067153777f
#define INV2PI_F  0.15915494309189533576888f
#define TWOPI_F   6.28318530717958647692528f
#define MIPLEVELS 5
#define MAXISLIGHTS 64

__device__ __forceinline__ float3 normalize3( const float3 v ) { return v * rsqrtf( dot( v, v ) ); }
__device__ __forceinline__ float3 reflect3( const float3 i, const float3 n ) { return i - 2.0f * n * dot( n, i ); }
__device__ __forceinline__ float sqr( const float x ) { return x * x; }
__device__ __forceinline__ float char2flt( const uint32_t a, const int s ) { return (float)((a >> s) & 255u) * (1.0f / 255.0f); }
__device__ __forceinline__ float2 half2x( const uint32_t v ) { return __half22float2( *(const __half2*)&v ); }

__device__ __forceinline__ uint32_t WangHash( uint32_t s ) { s = (s ^ 61) ^ (s >> 16), s *= 9, s = s ^ (s >> 4), s *= 0x27d4eb2d, s = s ^ (s >> 15); return s; }
__device__ __forceinline__ uint32_t RandomInt( uint32_t& s ) { s ^= s << 13, s ^= s >> 17, s ^= s << 5; return s; }
__device__ __forceinline__ float RandomFloat( uint32_t& s ) { return RandomInt( s ) * 2.3283064365387e-10f; }

__device__ __forceinline__ float4 BlueNoise4( const uint32_t* __restrict__ bn, const int x, const int y, const int sampleIndex, const int dim )
{
	// tools_shared.h:324-337 (Heitz ranking/scrambling tiles expanded to one uint per entry)
	const uint4 rank = *(const uint4*)(bn + dim + (x + y * 128) * 8 + 65536 * 3);
	const int i0 = (sampleIndex ^ rank.x) & 255, i1 = (sampleIndex ^ rank.y) & 255;
	const int i2 = (sampleIndex ^ rank.z) & 255, i3 = (sampleIndex ^ rank.w) & 255;
	const uint32_t v0 = bn[dim + 0 + i0 * 256], v1 = bn[dim + 1 + i1 * 256];
	const uint32_t v2 = bn[dim + 2 + i2 * 256], v3 = bn[dim + 3 + i3 * 256];
	const uint4 scr = *(const uint4*)(bn + (dim & 7) + (x + y * 128) * 8 + 65536);
	return make_float4( (0.5f + (int)(v0 ^ scr.x)) * (1.0f / 256.0f), (0.5f + (int)(v1 ^ scr.y)) * (1.0f / 256.0f),
		(0.5f + (int)(v2 ^ scr.z)) * (1.0f / 256.0f), (0.5f + (int)(v3 ^ scr.w)) * (1.0f / 256.0f) );
}

__device__ __forceinline__ uint32_t PackNormal( const float3 N )
{
	const float f = 65535.0f / fmaxf( sqrtf( 8.0f * N.z + 8.0f ), 0.0001f );
	return (uint32_t)(N.x * f + 32767.0f) + ((uint32_t)(N.y * f + 32767.0f) << 16);
}
__device__ __forceinline__ float3 UnpackNormal( const uint32_t p )
{
	float nx = (float)(p & 65535) * (2.0f / 65535.0f) - 1.0f, ny = (float)(p >> 16) * (2.0f / 65535.0f) - 1.0f;
	const float nz = 1.0f, nw = -1.0f;
	float l = nx * -nx + ny * -ny + nz * -nw;
	const float z = l;
	l = sqrtf( l ), nx *= l, ny *= l;
	return make_float3( nx * 2.0f, ny * 2.0f, z * 2.0f - 1.0f );
}

__device__ __forceinline__ float3 SafeOrigin( const float3 O, const float3 R, const float3 N, const float eps )
{
	return O + N * (dot( N, R ) > 0 ? eps : -eps);
}

/* ---- sky (tools_shared.h:185-211) ---- */
__device__ __forceinline__ float SphericalTheta( const float3 v ) { return acosf( fminf( 1.0f, fmaxf( -1.0f, v.z ) ) ); }
__device__ __forceinline__ float SphericalPhi( const float3 v ) { const float p = atan2f( v.y, v.x ); return p < 0 ? p + 2 * PI_F : p; }
__device__ __forceinline__ float3 SampleSky( const RenderParams& p, const float3 D, const bool small )
{
	const uint32_t w = small ? (uint32_t)p.skyW >> 6 : (uint32_t)p.skyW, h = small ? (uint32_t)p.skyH >> 6 : (uint32_t)p.skyH;
	const uint32_t u = (uint32_t)(w * SphericalPhi( D ) * INV2PI_F - 0.5f);
	const uint32_t v = (uint32_t)(h * SphericalTheta( D ) * INVPI_F - 0.5f);
	const uint32_t idx = u + v * w;
	if (idx >= w * h) return f3( 0 );
	return xyz( p.skyPixels[idx + (small ? p.skyW * p.skyH : 0)] );
}

/* ---- textures (sampling_shared.h:35-104) ---- */
__device__ __forceinline__ float4 U8ToFloat4( const uchar4 v ) { const float r = 1.0f / 256.0f; return make_float4( v.x * r, v.y * r, v.z * r, v.w * r ); }
__device__ __forceinline__ float4 operator*( const float4 a, const float s ) { return make_float4( a.x * s, a.y * s, a.z * s, a.w * s ); }
__device__ __forceinline__ float4 operator+( const float4 a, const float4 b ) { return make_float4( a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w ); }
enum { TEX_ARGB32 = 0, TEX_ARGB128 = 1, TEX_NRM32 = 2 };
__device__ __forceinline__ float4 FetchTexel( const RenderParams& p, const float2 uv, const int o, const int w, const int h, const int storage )
{
	const float tx = fmaxf( uv.x + 1000, 0.0f ) * w - 0.5f, ty = fmaxf( uv.y + 1000, 0.0f ) * h - 0.5f;
	const int iu = ((int)tx) % w, iv = ((int)ty) % h;
	const float fu = tx - floorf( tx ), fv = ty - floorf( ty );
	const float w0 = (1 - fu) * (1 - fv), w1 = fu * (1 - fv), w2 = (1 - fu) * fv, w3 = 1 - (w0 + w1 + w2);
	const uint32_t iu1 = (iu + 1) % w, iv1 = (iv + 1) % h;
	float4 p0, p1, p2, p3;
	if (storage == TEX_ARGB128)
		p0 = p.argb128[o + iu + iv * w], p1 = p.argb128[o + iu1 + iv * w], p2 = p.argb128[o + iu + iv1 * w], p3 = p.argb128[o + iu1 + iv1 * w];
	else
	{
		const uchar4* t = storage == TEX_ARGB32 ? p.argb32 : p.nrm32;
		p0 = U8ToFloat4( t[o + iu + iv * w] ), p1 = U8ToFloat4( t[o + iu1 + iv * w] );
		p2 = U8ToFloat4( t[o + iu + iv1 * w] ), p3 = U8ToFloat4( t[o + iu1 + iv1 * w] );
	}
	return p0 * w0 + p1 * w1 + p2 * w2 + p3 * w3;
}
__device__ __forceinline__ float4 FetchTexelTrilinear( const RenderParams& p, const float lambda, const float2 uv, const int offset, const int width, const int height )
{
	int level0 = 0, level1 = 0;
	float f = 0;
	if (lambda >= 0) level0 = min( MIPLEVELS - 1, (int)lambda ), level1 = min( MIPLEVELS - 1, level0 + 1 ), f = lambda - floorf( lambda );
	const float scale = (float)(width * height) * 1.3333333333f;
	const int o0 = offset + (int)(scale * (1 - __uint_as_float( (127 - 2 * level0) << 23 )));
	const int o1 = offset + (int)(scale * (1 - __uint_as_float( (127 - 2 * level1) << 23 )));
	const float4 p0 = FetchTexel( p, uv, o0, width >> level0, height >> level0, TEX_ARGB32 );
	const float4 p1 = FetchTexel( p, uv, o1, width >> level1, height >> level1, TEX_ARGB32 );
	return p0 * (1 - f) + p1 * f;
}

/* ---- material ---- */
#define MAT_ISDIELECTRIC     (1 << 0)
#define MAT_DIFFUSEMAPISHDR  (1 << 1)
#define MAT_HASDIFFUSEMAP    (1 << 2)
#define MAT_HASNORMALMAP     (1 << 3)
#define MAT_HASSPECULARITYMAP (1 << 4)
#define MAT_HASROUGHNESSMAP  (1 << 5)
#define MAT_HAS2NDNORMALMAP  (1 << 7)
#define MAT_HAS2NDDIFFUSEMAP (1 << 9)
#define MAT_HASSMOOTHNORMALS (1 << 11)

struct Shading
{
	float3 color; int flags;		// flags bit 0: alpha-rejected texel
	float3 transmittance;
	uint4 parameters;				// 0.8 fixed point Disney parameter block + eta (core_settings.h:146)
	float4 tint;					// hue of the base colour at unit luminance + its luminance (material_shared.h:118-119); Disney model only
};
#define SH_ROUGHNESS( s ) (fmaxf( 0.001f, char2flt( (s).parameters.x, 24 ) ))
#define SH_TRANSMISSION( s ) char2flt( (s).parameters.z, 16 )
#define SH_ETA( s ) __uint_as_float( (s).parameters.w )

struct InstDesc { const float4* triangles; int d1, d2; float4 A, B, C, D; };	// CoreInstanceDesc

__device__ __forceinline__ float2 MapUV( const uint4 m, const float tu, const float tv )
{
	const float2 sc = half2x( m.y ), of = half2x( m.z );
	return make_float2( sc.x * (of.x + tu), sc.y * (of.y + tv) );
}

__device__ __forceinline__ void GetShadingData( const RenderParams& p, const float3 D, const float u, const float v, const float coneWidth,
	const float4* __restrict__ tri, const InstDesc& inst, Shading& s, float3& N, float3& iN, float3& fN, float3& T )
{
	const float4 t1 = __ldg( tri + 1 ), t2 = __ldg( tri + 2 ), t3 = __ldg( tri + 3 ), t4 = __ldg( tri + 4 ), t5 = __ldg( tri + 5 );
	const DevMaterial& mat = p.materials[__float_as_int( t1.w )];
	const uint4 base = mat.q[0];
	const uint32_t flags = base.w;
	const float2 rg = half2x( base.x ), bm = half2x( base.y ), gb = half2x( base.z );
	s.color = make_float3( rg.x, rg.y, bm.x ), s.flags = 0;
	s.transmittance = make_float3( bm.y, gb.x, gb.y );
	s.parameters = mat.q[1];
	{
		// CIE XYZ round trip of the untextured base colour (material_shared.h:19-33,118-119)
		const float3 c = s.color;
		const float X = fmaxf( 0.0f, 0.412453f * c.x + 0.357580f * c.y + 0.180423f * c.z ), Y = fmaxf( 0.0f, 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z );
		const float Z = fmaxf( 0.0f, 0.019334f * c.x + 0.119193f * c.y + 0.950227f * c.z );
		float3 t = f3( 1 );
		if (Y > 0)
		{
			const float r = 1.0f / Y, x = X * r, y = Y * r, z = Z * r;
			t = make_float3( fmaxf( 0.0f, 3.240479f * x - 1.537150f * y - 0.498535f * z ), fmaxf( 0.0f, -0.969256f * x + 1.875992f * y + 0.041556f * z ),
				fmaxf( 0.0f, 0.055648f * x - 0.204043f * y + 1.057311f * z ) );
		}
		s.tint = make_float4( t.x, t.y, t.z, Y );
	}
	// SetupFrame (material_shared.h:42-85)
	N = make_float3( t2.w, t3.w, t4.w ), iN = N;
	T = xyz( t5 );
	const float w = 1 - (u + v);
	if (flags & MAT_HASSMOOTHNORMALS) iN = normalize3( w * xyz( t2 ) + u * xyz( t3 ) + v * xyz( t4 ) );
	const float3 A = xyz( inst.A ), B = xyz( inst.B ), C = xyz( inst.C );
	N = normalize3( N.x * A + N.y * B + N.z * C );
	iN = normalize3( iN.x * A + iN.y * B + iN.z * C );
	fN = iN;
	// texturing (material_shared.h:127-199)
	float tu = 0, tv = 0;
	if (flags & (MAT_HASDIFFUSEMAP | MAT_HAS2NDDIFFUSEMAP | MAT_HASSPECULARITYMAP | MAT_HASNORMALMAP | MAT_HAS2NDNORMALMAP | MAT_HASROUGHNESSMAP))
	{
		const float4 t0 = __ldg( tri );
		tu = w * t0.x + u * t0.y + v * t0.z, tv = w * t1.x + u * t1.y + v * t1.z;
	}
	if (flags & MAT_HASDIFFUSEMAP)
	{
		const float triLOD = __ldg( tri + 7 ).w;
		const float lambda = triLOD + log2f( coneWidth * (1.0f / fabsf( dot( D, N ) )) );
		const uint4 m = mat.q[2];
		const float4 texel = FetchTexelTrilinear( p, lambda, MapUV( m, tu, tv ), m.w, m.x & 0xffff, m.x >> 16 );
		if (texel.w < 0.5f) { s.flags |= 1; return; }
		s.color = s.color * xyz( texel );
		if (flags & MAT_HAS2NDDIFFUSEMAP)
		{
			const uint4 m1 = mat.q[3];
			s.color += xyz( FetchTexel( p, MapUV( m1, tu, tv ), m1.w, m1.x & 0xffff, m1.x >> 16, TEX_ARGB32 ) ) - f3( 0.5f );
		}
	}
	if (flags & MAT_HASNORMALMAP)
	{
		const float3 Bt = xyz( __ldg( tri + 6 ) );
		const uint4 m = mat.q[4];
		const float sb = (float)((base.z >> 8) & 255) - 128.0f;
		const float n0scale = copysignf( -0.0001f + 0.0001f * __expf( 0.1f * fabsf( sb ) ), sb );
		float3 sn = (xyz( FetchTexel( p, MapUV( m, tu, tv ), m.w, m.x & 0xffff, m.x >> 16, TEX_NRM32 ) ) - f3( 0.5f )) * 2.0f;
		sn.x *= n0scale, sn.y *= n0scale;
		if (flags & MAT_HAS2NDNORMALMAP)
		{
			const uint4 m1 = mat.q[5];
			const float sb1 = (float)((base.z >> 16) & 255) - 128.0f;
			const float n1scale = copysignf( -0.0001f + 0.0001f * __expf( 0.1f * sb1 ), sb1 );
			float3 l1 = (xyz( FetchTexel( p, MapUV( m1, tu, tv ), m1.w, m1.x & 0xffff, m1.x >> 16, TEX_NRM32 ) ) - f3( 0.5f )) * 2.0f;
			l1.x *= n1scale, l1.y *= n1scale;
			sn += l1;
		}
		sn = normalize3( sn );
		fN = normalize3( sn.x * T + sn.y * Bt + sn.z * iN );
	}
	if (flags & MAT_HASROUGHNESSMAP)
	{
		const uint4 m = mat.q[7];
		const float4 texel = FetchTexel( p, MapUV( m, tu, tv ), m.w, m.x & 0xffff, m.x >> 16, TEX_ARGB32 );
		s.parameters.x = (s.parameters.x & 0x00ffffff) + ((int)(texel.y * 255.0f) << 24);
		s.parameters.x = (s.parameters.x & 0xffffff00) + (int)(texel.x * 255.0f);
	}
}

/* ---- shared host/device helpers (common_functions.h:52-116) ---- */
__device__ __forceinline__ float3 RandomBarycentrics( const float r0 )
{
	const uint32_t uf = (uint32_t)(r0 * 4294967296.0f);
	float2 A = make_float2( 1, 0 ), B = make_float2( 0, 1 ), C = make_float2( 0, 0 );
	for (int i = 0; i < 16; ++i)
	{
		const int d = (uf >> (2 * (15 - i))) & 3;
		float2 An, Bn, Cn;
		switch (d)
		{
		case 0: An = make_float2( (B.x + C.x) * 0.5f, (B.y + C.y) * 0.5f ), Bn = make_float2( (A.x + C.x) * 0.5f, (A.y + C.y) * 0.5f ), Cn = make_float2( (A.x + B.x) * 0.5f, (A.y + B.y) * 0.5f ); break;
		case 1: An = A, Bn = make_float2( (A.x + B.x) * 0.5f, (A.y + B.y) * 0.5f ), Cn = make_float2( (A.x + C.x) * 0.5f, (A.y + C.y) * 0.5f ); break;
		case 2: An = make_float2( (B.x + A.x) * 0.5f, (B.y + A.y) * 0.5f ), Bn = B, Cn = make_float2( (B.x + C.x) * 0.5f, (B.y + C.y) * 0.5f ); break;
		default: An = make_float2( (C.x + A.x) * 0.5f, (C.y + A.y) * 0.5f ), Bn = make_float2( (C.x + B.x) * 0.5f, (C.y + B.y) * 0.5f ), Cn = C; break;
		}
		A = An, B = Bn, C = Cn;
	}
	const float rx = (A.x + B.x + C.x) * 0.3333333f, ry = (A.y + B.y + C.y) * 0.3333333f;
	return make_float3( rx, ry, 1 - rx - ry );
}
__device__ __forceinline__ float3 Tangent2World( const float3 V, const float3 N )
{
	const float sign = copysignf( 1.0f, N.z );
	const float a = -1.0f / (sign + N.z), b = N.x * N.y * a;
	const float3 B = make_float3( 1.0f + sign * N.x * N.x * a, sign * b, -sign * N.x );
	const float3 T = make_float3( b, sign + N.y * N.y * a, -N.y );
	return V.x * T + V.y * B + V.z * N;
}
__device__ __forceinline__ float3 DiffuseReflectionCosWeighted( const float r0, const float r1 )
{
	const float term1 = TWOPI_F * r0, term2 = sqrtf( 1 - r1 );
	float s, c;
	sincosf( term1, &s, &c );
	return make_float3( c * term2, s * term2, sqrtf( r1 ) );
}

/* ---- lights (lights_shared.h) ---- */
__device__ __forceinline__ float PotentialTriLight( const RenderParams& p, const int idx, const float3 O, const float3 N, const float3 I, const float3 bary )
{
	const float4* l = p.triLights + idx * 6;
	const float4 centre4 = l[0], LN = l[1];
	float3 L = I;
	if (bary.x >= 0)
	{
		const float4 V0 = l[3], V1 = l[4], V2 = l[5];
		L = make_float3( bary.x * V0.x + bary.y * V1.x + bary.z * V2.x, bary.x * V0.y + bary.y * V1.y + bary.z * V2.y, bary.x * V0.z + bary.y * V1.z + bary.z * V2.z );
	}
	L -= O;
	const float att = 1.0f / dot( L, L );
	L = normalize3( L );
	const float LNdotL = fmaxf( 0.0f, -dot( xyz( LN ), L ) ), NdotL = fmaxf( 0.0f, dot( N, L ) );
	return centre4.w * LNdotL * NdotL * att;
}
__device__ __forceinline__ float PotentialPointLight( const RenderParams& p, const int idx, const float3 I, const float3 N )
{
	const float4 pos4 = p.pointLights[idx * 2];
	const float3 L = xyz( pos4 ) - I;
	const float NdotL = fmaxf( 0.0f, dot( N, normalize3( L ) ) ), att = 1.0f / dot( L, L );
	return pos4.w * NdotL * att;
}
__device__ __forceinline__ float PotentialSpotLight( const RenderParams& p, const int idx, const float3 I, const float3 N )
{
	const float4 pos4 = p.spotLights[idx * 3], rad4 = p.spotLights[idx * 3 + 1], dir4 = p.spotLights[idx * 3 + 2];
	float3 L = xyz( pos4 ) - I;
	const float att = 1.0f / dot( L, L );
	L = normalize3( L );
	const float d = (fmaxf( 0.0f, -dot( L, xyz( dir4 ) ) ) - rad4.w) / (pos4.w - rad4.w);
	const float NdotL = fmaxf( 0.0f, dot( N, L ) ), LNdotL = fmaxf( 0.0f, fminf( 1.0f, d ) );
	return (rad4.x + rad4.y + rad4.z) * LNdotL * NdotL * att;
}
__device__ __forceinline__ float PotentialDirLight( const RenderParams& p, const int idx, const float3 I, const float3 N )
{
	const float4 dir4 = p.dirLights[idx * 2];
	const float LNdotL = fmaxf( 0.0f, -(dir4.x * N.x + dir4.y * N.y + dir4.z * N.z) );
	return dir4.w * LNdotL;
}
#define NTRI( p ) ((p).lightCounts.x & 0xffff)

/* potential[] of every light seen from (O, N); returns the sum. Slot order: tri, point, spot, directional. */
__device__ __forceinline__ float LightPotentials( const RenderParams& p, float* potential, const float3 O, const float3 N, const float3 I, const float3 bary )
{
	float sum = 0;
	int lights = 0;
	for (int i = 0; i < NTRI( p ); i++) { const float c = PotentialTriLight( p, i, O, N, I, bary ); potential[lights++] = c, sum += c; }
	for (int i = 0; i < p.lightCounts.y; i++) { const float c = PotentialPointLight( p, i, O, N ); potential[lights++] = c, sum += c; }
	for (int i = 0; i < p.lightCounts.z; i++) { const float c = PotentialSpotLight( p, i, O, N ); potential[lights++] = c, sum += c; }
	for (int i = 0; i < p.lightCounts.w; i++) { const float c = PotentialDirLight( p, i, O, N ); potential[lights++] = c, sum += c; }
	return sum;
}

/* More than MAXISLIGHTS lights: the reference's importance sampling overruns potential[MAXISLIGHTS] (undefined behaviour); this core
   then takes the reference's other branch - the #else of ISLIGHTS, lights_shared.h:256-260: uniform pick, pickProb = 1 / lightCount. */
__device__ __forceinline__ int LightTotal( const RenderParams& p ) { return NTRI( p ) + p.lightCounts.y + p.lightCounts.z + p.lightCounts.w; }
__device__ __forceinline__ float LightPickProb( const RenderParams& p, const int idx, const float3 O, const float3 N, const float3 I )
{
	if (LightTotal( p ) > MAXISLIGHTS) return 1.0f / (float)LightTotal( p );
	float potential[MAXISLIGHTS];
	const float sum = LightPotentials( p, potential, O, N, I, f3( -1 ) );
	if (sum <= 0) return 0;
	return potential[idx] / sum;
}

__device__ __forceinline__ float3 RandomPointOnLight( const RenderParams& p, const float r0, float r1, const float3 I, const float3 N,
	float& pickProb, float& lightPdf, float3& lightColor )
{
	const int nTri = NTRI( p ), nPoint = p.lightCounts.y, nSpot = p.lightCounts.z, nDir = p.lightCounts.w;
	const float lightCount = nTri + nPoint + nSpot + nDir;
	const float3 bary = RandomBarycentrics( r0 );
	int lightIdx = 0;
	if (lightCount > MAXISLIGHTS) pickProb = 1.0f / lightCount, lightIdx = (int)(r1 * lightCount);
	else
	{
		float potential[MAXISLIGHTS];
		const float sum = LightPotentials( p, potential, I, N, I, bary );
		if (sum <= 0) { lightPdf = 0; return f3( 1 ); }
		const int lights = (int)lightCount;
		r1 *= sum;
		float total = 0;
		for (int i = 0; i < lights; i++) { total += potential[i]; if (total >= r1) { lightIdx = i; break; } }
		pickProb = potential[lightIdx] / sum;
	}
	lightIdx = max( 0, min( lightIdx, (int)lightCount - 1 ) );
	if (lightIdx < nTri)
	{
		const float4* l = p.triLights + lightIdx * 6;
		const float4 V0 = l[3], V1 = l[4], V2 = l[5], LN = l[1];
		lightColor = xyz( l[2] );
		const float3 P = make_float3( bary.x * V0.x + bary.y * V1.x + bary.z * V2.x, bary.x * V0.y + bary.y * V1.y + bary.z * V2.y, bary.x * V0.z + bary.y * V1.z + bary.z * V2.z );
		float3 L = I - P;
		const float sqDist = dot( L, L );
		L = normalize3( L );
		const float LNdotL = L.x * LN.x + L.y * LN.y + L.z * LN.z;
		const float reciSolidAngle = sqDist / (LN.w * LNdotL);
		lightPdf = (LNdotL > 0 && dot( L, N ) < 0) ? reciSolidAngle : 0;
		return P;
	}
	else if (lightIdx < nTri + nPoint)
	{
		const float4* l = p.pointLights + (lightIdx - nTri) * 2;
		const float3 P = xyz( l[0] ), L = P - I;
		const float sqDist = dot( L, L );
		lightColor = xyz( l[1] ) * (1.0f / sqDist);
		lightPdf = dot( L, N ) > 0 ? 1 : 0;
		return P;
	}
	else if (lightIdx < nTri + nPoint + nSpot)
	{
		const float4* l = p.spotLights + (lightIdx - (nTri + nPoint)) * 3;
		const float4 V0 = l[0], V1 = l[1], Dl = l[2];
		const float3 P = xyz( V0 );
		float3 L = I - P;
		const float sqDist = dot( L, L );
		L = normalize3( L );
		const float d = (fmaxf( 0.0f, L.x * Dl.x + L.y * Dl.y + L.z * Dl.z ) - V1.w) / (V0.w - V1.w);
		const float LNdotL = fminf( 1.0f, d );
		lightPdf = (LNdotL > 0 && dot( L, N ) < 0) ? (sqDist / LNdotL) : 0;
		lightColor = xyz( V1 );
		return P;
	}
	else
	{
		const float4* l = p.dirLights + (lightIdx - (nTri + nPoint + nSpot)) * 2;
		const float3 L = xyz( l[0] );
		lightColor = xyz( l[1] );
		lightPdf = dot( L, N ) < 0 ? 1 : 0;
		return I - 1000.0f * L;
	}
}

/* ---- BSDF: Lambert + pure specular + dielectric (lambert.h:32-125) ---- */
__device__ __forceinline__ float Fr_L( float VDotN, float eio )
{
	if (VDotN < 0.0f) eio = 1.0f / eio, VDotN = fabsf( VDotN );
	const float SinThetaT2 = sqr( eio ) * (1.0f - VDotN * VDotN);
	if (SinThetaT2 > 1.0f) return 1.0f;
	const float LDotN = sqrtf( 1.0f - SinThetaT2 );
	const float r1 = (VDotN - eio * LDotN) / (VDotN + eio * LDotN), r2 = (LDotN - eio * VDotN) / (LDotN + eio * VDotN);
	return 0.5f * (sqr( r1 ) + sqr( r2 ));
}
__device__ __forceinline__ bool Refract_L( const float3 wi, const float3 n, const float eta, float3& wt )
{
	const float cosThetaI = fabsf( dot( n, wi ) );
	const float sin2ThetaI = fmaxf( 0.0f, 1.0f - cosThetaI * cosThetaI ), sin2ThetaT = eta * eta * sin2ThetaI;
	if (sin2ThetaT >= 1) return false;
	const float cosThetaT = sqrtf( 1.0f - sin2ThetaT );
	wt = eta * (wi * -1.0f) + (eta * cosThetaI - cosThetaT) * n;
	return true;
}
__device__ __forceinline__ float3 EvaluateBSDF( const Shading& s, const float3 iN, const float3 wi, float& pdf )
{
	if (SH_TRANSMISSION( s ) > 0.999f || SH_ROUGHNESS( s ) <= 0.001f) { pdf = 0; return f3( 0 ); }
	pdf = fabsf( dot( wi, iN ) ) * INVPI_F;
	return s.color * INVPI_F;
}
__device__ __forceinline__ float3 SampleBSDF( const Shading& s, float3 iN, const float3 N, const float3 wo, const float distance,
	const float r3, const float r4, float3& wi, float& pdf, bool& specular )
{
	const float flip = (dot( wo, N ) < 0) ? -1 : 1;
	iN *= flip;
	specular = true, pdf = 1;
	float3 bsdf;
	const float transmission = SH_TRANSMISSION( s );
	if (r4 < transmission)
	{
		const float eio = flip < 0 ? (1.0f / SH_ETA( s )) : SH_ETA( s ), F = Fr_L( dot( iN, wo ), eio );
		const float3 beer = make_float3( expf( -s.transmittance.x * distance * 2.0f ), expf( -s.transmittance.y * distance * 2.0f ), expf( -s.transmittance.z * distance * 2.0f ) );
		if (r3 < F)
		{
			wi = reflect3( wo * -1.0f, iN );
			bsdf = s.color * beer * (1 / fabsf( dot( iN, wi ) ));
		}
		else
		{
			if (!Refract_L( wo, iN, eio, wi )) return f3( 0 );
			return s.color * beer * (1 / fabsf( dot( iN, wi ) ));
		}
	}
	else
	{
		const float pReflect = 1 - SH_ROUGHNESS( s );
		if (r3 < pReflect)
		{
			wi = reflect3( wo * -1.0f, iN );
			bsdf = s.color * (1.0f / fabsf( dot( iN, wi ) ));
		}
		else
		{
			const float r5 = (r3 - pReflect) / (1 - pReflect), r6 = (r4 - transmission) / (1 - transmission);
			wi = normalize3( Tangent2World( DiffuseReflectionCosWeighted( r5, r6 ), iN ) );
			pdf = fmaxf( 0.0f, dot( wi, iN ) ) * INVPI_F;
			specular = false;
			bsdf = s.color * INVPI_F;
		}
	}
	if (dot( N * flip, wi ) <= 0) pdf = 0;
	return bsdf;
}

} // namespace lh2b
#include "bsdf_disney.cuh"
namespace lh2b
{

/* ---- filter features (Optix7Filter pathtracer.h:44-58; tools_shared.h:122-130,160-166) ---- */
__device__ __forceinline__ uint32_t PackNormal2( const float3 N )
{
	const uint32_t x = min( max( (uint32_t)((N.x + 1) * 511), 0u ), 1023u ), y = min( max( (uint32_t)((N.y + 1) * 511), 0u ), 1023u );
	const uint32_t z = min( max( (uint32_t)((N.z + 1) * 511), 0u ), 1023u );
	return (x << 2u) + (y << 12u) + (z << 22u);
}
__device__ __forceinline__ uint32_t HDRtoRGB32( const float3 c )
{
	return ((uint32_t)(1023.0f * fminf( 1.0f, c.x )) << 22) + ((uint32_t)(2047.0f * fminf( 1.0f, c.y )) << 11) + (uint32_t)(2047.0f * fminf( 1.0f, c.z ));
}
__device__ __forceinline__ float3 RGB32toHDRs( const uint32_t c )
{
	return make_float3( (float)(c >> 22) * (1.0f / 1023.0f), (float)((c >> 11) & 2047) * (1.0f / 2047.0f), (float)(c & 2047) * (1.0f / 2047.0f) );
}
__device__ __forceinline__ void StoreFeatures( const RenderParams& p, const uint32_t pathIdx, const uint32_t albedo, const uint32_t packedNormal, const float t,
	const uint32_t isSpecular, const uint32_t matid )
{
	const uint32_t history = p.features[pathIdx].w & 15;	// the history count survives (prepareFilter owns it)
	p.features[pathIdx] = make_uint4( albedo, packedNormal, __float_as_uint( t ), (isSpecular << 4) + (matid << 6) + history );
}
__device__ __forceinline__ void StoreDepthDerivatives( const RenderParams& p, const uint32_t pathIdx, const float depth, const float4* __restrict__ tri )
{
	// depth of the triangle's plane along the rays through (x+1, y) and (x, y+1) minus the depth at (x, y)
	const int x = pathIdx % p.w, y = pathIdx / p.w;
	const float3 triN = make_float3( __ldg( tri + 2 ).w, __ldg( tri + 3 ).w, __ldg( tri + 4 ).w ), v0 = xyz( __ldg( tri + 8 ) ), pos = xyz( p.posLensSize );
	const float3 dX = normalize3( p.p1 + (x + 0.5f + 1) * (1.0f / p.w) * p.right + (y + 0.5f) * (1.0f / p.h) * p.up - pos );
	const float3 dY = normalize3( p.p1 + (x + 0.5f) * (1.0f / p.w) * p.right + (y + 0.5f + 1) * (1.0f / p.h) * p.up - pos );
	const float num = dot( v0 - pos, triN );
	p.deltaDepth[pathIdx] = make_float4( 0, 0, num / dot( triN, dX ) - depth, num / dot( triN, dY ) - depth );
}

__device__ __forceinline__ void ClampIntensity( float3& c, const float clampValue )
{
	const float v = fmaxf( c.x, fmaxf( c.y, c.z ) );
	if (v > clampValue) { const float m = clampValue / v; c.x *= m, c.y *= m, c.z *= m; }
}
__device__ __forceinline__ void FixNan( float3& a ) { if (!isfinite( a.x + a.y + a.z )) a = f3( 0 ); }

__device__ __forceinline__ void Accumulate( float4* acc, const float3 c )
{
	atomicAdd( acc, make_float4( c.x, c.y, c.z, 0 ) );	// red.global.add.v4.f32 on sm_90+
}

/* Warp-aggregated slot allocation: one atomic per warp, lanes take consecutive slots. */
__device__ __forceinline__ uint32_t WarpAlloc( uint32_t* counter, const bool want )
{
	const uint32_t mask = __ballot_sync( 0xffffffffu, want );
	if (mask == 0) return 0;
	const int lane = threadIdx.x & 31, leader = __ffs( mask ) - 1;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd( counter, __popc( mask ) );
	base = __shfl_sync( 0xffffffffu, base, leader );
	return base + __popc( mask & ((1u << lane) - 1) );
}

/* BSDF: 0 = lambert.h model (a14), 1 = Disney principled model (bsdf_disney.cuh); the only line of pathtracer.h that depends on
   the model is the ROUGHNESS factor on the NEE term (BSDF_HAS_PURE_SPECULARS, pathtracer.h:194-198) */
template <int BSDF, int MINB> __global__ void __launch_bounds__( 128, MINB ) shadeKernel( const RenderParams p, const PathSet in, const PathSet out,
	const float4* __restrict__ hits, const PathSet conn, const int pathLength, const uint32_t R0, const int useNEE )
{
	const uint32_t pathCount = pathLength == 1 ? p.stride : p.counters->extensionRays[pathLength - 1];
	const uint32_t rounds = (pathCount + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
	for (uint32_t round = 0; round < rounds; round++)
	{
		uint32_t job = (round * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
		bool inRange = job < pathCount;
		if (pathLength == 1 && p.stride != (uint32_t)(p.w * p.h * p.spp) && inRange)
		{
			// tile-sharded frame: primary paths live at their global path index; job -> (sample, tile row of the band, pixel in it)
			const uint32_t tileRowPixels = 4u * p.w, bandPixels = BandTileRows( p ) * tileRowPixels, smp = job / bandPixels, i = job - smp * bandPixels;
			const uint32_t j = i / tileRowPixels, rem = i - j * tileRowPixels, row = ((uint32_t)p.bandY0 / 4 + j * p.bandStep) * 4 + rem / p.w;
			inRange = row < (uint32_t)p.bandY1 && row >= (uint32_t)p.bandY0;
			job = smp * (p.w * p.h) + row * p.w + rem % p.w;
		}
		// every lane takes part in the two warp-wide allocations below; inactive lanes carry 'false'
		bool emitShadow = false, emitExt = false;
		float4 cO, cD, cE, eO, eD, eT;
		if (inRange) do
		{
			const float4 O4 = in.O[job], D4 = in.D[job];
			float4 T4 = pathLength == 1 ? make_float4( 1, 1, 1, 1 ) : in.T[job];
			const float4 hit = hits[job];
			const float bsdfPdf = T4.w;
			uint32_t data = __float_as_uint( O4.w );
			const float3 D = xyz( D4 );
			float3 throughput = xyz( T4 );
			const int prim = __float_as_int( hit.z ), instIdx = __float_as_int( hit.y );
			const uint32_t pathIdx = data >> 6;
			const bool filter = p.features != nullptr;
			// filter mode: light arriving after the first diffuse bounce goes to the second (indirect) accumulator half
			const uint32_t pixelIdx = pathIdx % (p.w * p.h) + ((filter && (data & S_BOUNCED)) ? p.w * p.h : 0);
			const uint32_t seedIdx = pathIdx + p.sampleBase * (p.w * p.h);	// path index within the whole (multi-GPU) frame
			const uint32_t sampleIdx = seedIdx / (p.w * p.h) + p.pass;
			const bool firstHitToBeStored = filter && (data & S_BOUNCED) == 0 && sampleIdx == 0;
			if (pathLength == 1 && firstHitToBeStored)
			{
				StoreFeatures( p, pathIdx, 0, 0, 1e34f, 0, 0 );
				p.worldPos[pathIdx] = make_float4( 0, 0, 0, __uint_as_float( 0u ) ), p.deltaDepth[pathIdx] = make_float4( 0, 0, 0, 0 );
			}
			if (prim == -1)
			{
				// sky (pathtracer.h:85-94)
				const float* m = p.worldToSky;
				const float3 tD = make_float3( -(m[0] * D.x + m[1] * D.y + m[2] * D.z), -(m[4] * D.x + m[5] * D.y + m[6] * D.z), -(m[8] * D.x + m[9] * D.y + m[10] * D.z) );
				const float3 sky = SampleSky( p, tD, !filter && (data & S_BOUNCED) != 0 );	// the filter core always reads the full-size sky
				float3 contribution = throughput * sky * (1.0f / bsdfPdf);
				ClampIntensity( contribution, p.clampValue );
				FixNan( contribution );
				Accumulate( p.accumulator + pixelIdx, contribution );
				if (firstHitToBeStored)
				{
					const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( D * -1.0f ) + isSpecular;
					StoreFeatures( p, pathIdx, HDRtoRGB32( contribution ), packedNormal, hit.w, isSpecular, 0 );
					const float3 far = xyz( O4 ) + 50000 * D;
					p.worldPos[pathIdx] = make_float4( far.x, far.y, far.z, __uint_as_float( packedNormal ) ), p.deltaDepth[pathIdx] = make_float4( 0, 0, 0, 0 );
				}
				break;
			}
			const float hitU = (__float_as_uint( hit.x ) & 65535) * (1.0f / 65535.0f), hitV = (__float_as_uint( hit.x ) >> 16) * (1.0f / 65535.0f);
			const float hitT = hit.w;
			// object picking (pathtracer.h:97-102; filter core pathtracer.h:130-133). The reference lets every sample of the probed pixel
			// store the three words (a race at spp > 1): here exactly one path - the frame's first sample of that pixel -
			// writes them, with one 16-byte store, so the triple is deterministic and cannot tear.
			if (pathIdx == (uint32_t)p.probePixelIdx && pathLength == 1 && (!filter || sampleIdx == 0))
				*(int4*)&p.counters->probedInstid = make_int4( instIdx, prim, __float_as_int( hitT ), 0 );
			const InstDesc& inst = ((const InstDesc*)p.instDesc)[instIdx];
			const float4* tri = inst.triangles + (size_t)prim * 13;
			Shading sh;
			float3 N, iN, fN, T;
			const float3 I = xyz( O4 ) + hitT * D;
			GetShadingData( p, D, hitU, hitV, p.spreadAngle * hitT, tri, inst, sh, N, iN, fN, T );
			if (filter && !(sh.flags & 1))	// FILTERINGCORE: albedo never reaches zero so that it can be divided out (material_shared.h:200-205)
				sh.color = make_float3( fmaxf( 0.05f, sh.color.x ), fmaxf( 0.05f, sh.color.y ), fmaxf( 0.05f, sh.color.z ) );
			uint32_t seed = WangHash( seedIdx * 17 + R0 );
			if (sh.flags & 1)
			{
				// alpha-rejected texel: continue the same ray behind the surface (pathtracer.h:113-124)
				if (pathLength == p.maxPathLength)
				{
					if (firstHitToBeStored)
					{
						const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( N ) + isSpecular;
						StoreFeatures( p, pathIdx, 0, packedNormal, hitT, isSpecular, 0 );
						p.worldPos[pathIdx] = make_float4( I.x, I.y, I.z, __uint_as_float( packedNormal ) );
					}
				}
				else
				{
					const float3 nO = I + D * p.geometryEpsilon;
					if (!isfinite( T4.x + T4.y + T4.z )) T4 = make_float4( 0, 0, 0, T4.w );
					eO = make_float4( nO.x, nO.y, nO.z, O4.w ), eD = D4, eT = T4, emitExt = true;
				}
				break;
			}
			if (sh.color.x > 1.0f || sh.color.y > 1.0f || sh.color.z > 1.0f)
			{
				// emissive surface: terminate (pathtracer.h:127-153)
				const float DdotNL = -dot( D, N );
				if (DdotNL > 0)
				{
					float3 contribution = f3( 0 );
					if (pathLength == 1 || (data & S_SPECULAR) || !useNEE) contribution = sh.color;
					else
					{
						const float3 lastN = UnpackNormal( __float_as_uint( D4.w ) );
						const float area = __ldg( tri + 5 ).w;
						const int ltriIdx = __float_as_int( __ldg( tri ).w );
						const float lightPdf = (hitT * hitT) / (fabsf( dot( D, N ) ) * area);
						const float pickProb = LightPickProb( p, ltriIdx, xyz( O4 ), lastN, I );
						if ((bsdfPdf + lightPdf * pickProb) > 0) contribution = throughput * sh.color * (1.0f / (bsdfPdf + lightPdf * pickProb));
					}
					ClampIntensity( contribution, p.clampValue );
					FixNan( contribution );
					Accumulate( p.accumulator + pixelIdx, contribution );
				}
				if (firstHitToBeStored)
				{
					const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( N ) + isSpecular;
					StoreFeatures( p, pathIdx, HDRtoRGB32( sh.color ), packedNormal, hitT, isSpecular, 0 );
					p.worldPos[pathIdx] = make_float4( I.x, I.y, I.z, __uint_as_float( packedNormal ) );
					StoreDepthDerivatives( p, pathIdx, hitT, tri );
				}
				break;
			}
			if (!filter && (data & S_BOUNCED)) sh.parameters.x |= 255u << 24;	// path regularisation (not in the filter core)
			const float roughness = SH_ROUGHNESS( sh );
			if (roughness <= 0.001f || SH_TRANSMISSION( sh ) > (filter ? 0.999f : 0.5f)) data |= S_SPECULAR; else data &= ~S_SPECULAR;
			const float faceDir = (dot( D, N ) > 0) ? -1 : 1;
			if (firstHitToBeStored)
			{
				if (data & S_SPECULAR) p.features[pathIdx].x = HDRtoRGB32( sh.color );	// modulated by the first diffuse hit later
				else
				{
					float3 albedo = sh.color;
					if (data & S_VIASPECULAR) albedo *= RGB32toHDRs( p.features[pathIdx].x );
					const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0;
					// the reference writes 'flip ? -fN : fN' with flip = +-1.0f, i.e. always -fN (Optix7Filter pathtracer.h:229): kept
					const uint32_t packedNormal = PackNormal2( fN * -1.0f ) + isSpecular;
					StoreFeatures( p, pathIdx, HDRtoRGB32( albedo ), packedNormal, hitT, isSpecular, 0 );
					p.worldPos[pathIdx] = make_float4( I.x, I.y, I.z, __uint_as_float( packedNormal ) );
					StoreDepthDerivatives( p, pathIdx, hitT, tri );
				}
			}
			if (faceDir == 1) sh.transmittance = f3( 0 );
			throughput *= 1.0f / bsdfPdf;
			float4 r4;
			if (sampleIdx < 64)
			{
				const uint32_t x = ((seedIdx % p.w) + (p.shift & 127)) & 127, y = ((seedIdx / p.w) + (p.shift >> 24)) & 127;
				r4 = BlueNoise4( p.blueNoise, x, y, sampleIdx, 4 * pathLength - 4 );
			}
			else r4.x = RandomFloat( seed ), r4.y = RandomFloat( seed ), r4.z = RandomFloat( seed ), r4.w = RandomFloat( seed );
			// next event estimation (pathtracer.h:183-212)
			if ((data & S_SPECULAR) == 0 && useNEE)
			{
				float pickProb, lightPdf = 0;
				float3 lightColor;
				float3 L = RandomPointOnLight( p, r4.x, r4.y, I, fN * faceDir, pickProb, lightPdf, lightColor ) - I;
				const float dist = sqrtf( dot( L, L ) );
				L *= 1.0f / dist;
				const float NdotL = dot( L, fN * faceDir );
				if (NdotL > 0 && lightPdf > 0)
				{
					float lobePdf;
					const float3 f = BSDF == 0 ? EvaluateBSDF( sh, fN, L, lobePdf ) * roughness :
						EvaluateDisney( UnpackDisney( sh.color, sh.transmittance, sh.tint, sh.parameters ), fN, T, D * -1.0f, L, lobePdf );
					if (lobePdf > 0)
					{
						float3 contribution = throughput * f * lightColor * (NdotL / (pickProb * lightPdf + lobePdf));
						FixNan( contribution );
						ClampIntensity( contribution, p.clampValue );
						const float3 so = SafeOrigin( I, L, N, p.geometryEpsilon );
						cO = make_float4( so.x, so.y, so.z, 0 );
						cD = make_float4( L.x, L.y, L.z, dist - 2 * p.geometryEpsilon );
						cE = make_float4( contribution.x, contribution.y, contribution.z, __int_as_float( (int)pixelIdx ) );
						emitShadow = true;
					}
				}
			}
			if (data & (filter ? (uint32_t)S_BOUNCED : p.enoughBounces)) break;	// filter core: one diffuse bounce, always
			if (pathLength == p.maxPathLength)
			{
				if (firstHitToBeStored)
				{
					// nothing diffuse was found within the path length: leave something sensible for the filter
					const uint32_t isSpecular = (data & S_VIASPECULAR) ? 1 : 0, packedNormal = PackNormal2( N ) + isSpecular;
					StoreFeatures( p, pathIdx, 0, packedNormal, hitT, isSpecular, 0 );
					p.worldPos[pathIdx] = make_float4( I.x, I.y, I.z, __uint_as_float( packedNormal ) );
				}
				break;
			}
			float3 R;
			float newPdf;
			bool specular = false;
			const float r5 = RandomFloat( seed );	// third random number of the reference SampleBSDF call (unused by Lambert)
			const float3 bsdf = BSDF == 0 ? SampleBSDF( sh, fN, N, D * -1.0f, hitT, r4.z, r4.w, R, newPdf, specular ) :
				SampleDisney( UnpackDisney( sh.color, sh.transmittance, sh.tint, sh.parameters ), fN, N, T, D * -1.0f, hitT, r4.z, r4.w, r5, R, newPdf, specular );
			if (newPdf < 0.0001f || isnan( newPdf )) break;
			if (specular) data |= S_SPECULAR;
			const float rr = (filter || (data & S_SPECULAR) || ((data & S_BOUNCED) == 0)) ? 1 : fminf( 1.0f, fmaxf( fmaxf( bsdf.x, bsdf.y ), bsdf.z ) );
			if (rr < RandomFloat( seed )) break;
			throughput *= 1 / rr;
			const uint32_t packedNormal = PackNormal( fN * faceDir );
			if (!(data & S_SPECULAR)) data |= (data & S_BOUNCED) ? S_BOUNCEDTWICE : S_BOUNCED; else data |= S_VIASPECULAR;
			const float3 so = SafeOrigin( I, R, N, p.geometryEpsilon );
			FixNan( throughput );
			const float3 nt = throughput * bsdf * fabsf( dot( fN, R ) );
			eO = make_float4( so.x, so.y, so.z, __uint_as_float( data ) );
			eD = make_float4( R.x, R.y, R.z, __uint_as_float( packedNormal ) );
			eT = make_float4( nt.x, nt.y, nt.z, newPdf );
			emitExt = true;
		} while (0);
		const uint32_t si = WarpAlloc( &p.counters->shadowRays[pathLength], emitShadow );
		if (emitShadow) conn.O[si] = cO, conn.D[si] = cD, conn.T[si] = cE;
		const uint32_t ei = WarpAlloc( &p.counters->extensionRays[pathLength], emitExt );
		if (emitExt) out.O[ei] = eO, out.D[ei] = eD, out.T[ei] = eT;
	}
}

/* finalize (lib/CUDA/shared_kernel_code/finalize_shared.h:29-45): out = accumulator * (1 / samplesTaken), to a linear buffer. */
__global__ void finalizeKernel( const float4* __restrict__ accumulator, float4* __restrict__ out, const int n, const float scale )
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float4 a = __ldcs( accumulator + i );
	__stcs( out + i, make_float4( a.x * scale, a.y * scale, a.z * scale, a.w * scale ) );
}

/* resident blocks per SM the shade kernels are compiled for: Lambert 5 (102 registers, 54 B of spills: measured 9 % faster than 4 - 123
   registers, none - and 4 % faster than 6 on the B200, round 1), principled model 4 */
void LaunchShade( const RenderParams& p, const PathSet& in, const PathSet& out, const float4* hits, const PathSet& conn,
	int pathLength, uint32_t R0, bool useNEE, uint32_t maxPaths, int smCount, cudaStream_t s )
{
	// persistent-style grid: enough blocks to cover maxPaths, capped at 16 resident waves
	uint32_t blocks = (maxPaths + 127) / 128;
	const uint32_t cap = (uint32_t)smCount * (uint32_t)(p.bsdfModel == 1 ? 4 : 5) * 16;
	if (blocks > cap) blocks = cap;
	if (blocks == 0) return;
	if (p.bsdfModel == 1) shadeKernel<1, 4><<<blocks, 128, 0, s>>>( p, in, out, hits, conn, pathLength, R0, useNEE ? 1 : 0 );
	else shadeKernel<0, 5><<<blocks, 128, 0, s>>>( p, in, out, hits, conn, pathLength, R0, useNEE ? 1 : 0 );
}

void LaunchFinalize( const float4* accumulator, float4* out, int n, int samplesTaken, cudaStream_t s )
{
	if (n <= 0) return;
	finalizeKernel<<<(n + 255) / 256, 256, 0, s>>>( accumulator, out, n, 1.0f / (float)samplesTaken );
}

} // namespace lh2b
