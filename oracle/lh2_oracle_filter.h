/* lh2_oracle_filter.h - TEST INFRASTRUCTURE ONLY. CPU restatement of the reference's SVGF / TAA chain, i.e. the
   tail of RenderCore::FinalizeRender of the filtering core (lib/RenderCore_Optix7Filter/rendercore.cpp:904-934):

     prepareFilter( ... )                 lib/CUDA/shared_kernel_code/finalize_shared.h:169-314   -> Prepare
     applyFilter( 1 / 2 / 3 )             finalize_shared.h:320-484                                  -> ApplyFilter
     TAApass / unsharpenTAA               finalize_shared.h:498-583                                  -> TaaPass, UnsharpenTaa
     finalizeNoTAA                        finalize_shared.h:589-600                                  -> FinalizeNoTaa
   with the helpers they use: sampling_shared.h:22-27,111-215 (mitchellNetravali, ReadWorldPos, ReadTexelBmitchellNetravali,
   ReadTexelConsistent, ReadTexelConsistent2) and tools_shared.h:122-177,237-262 (normal / colour packing, YCoCg, luminance,
   5.11 fixed-point pairs).

   Pinned by tests/golden/filter_reference_vectors.npz: outputs of the reference's own kernels (compiled unmodified for sm_100a,
   oracle/_ref/libref_filter_gpu.so, run on the B200 by tools/make_golden_filter.py) for seeded inputs; tests/test_oracle_golden.py
   replays those inputs through this file. The reference kernels are fast-math builds (__expf, powf through exp2 / log2,
   approximate division and sqrt), this file uses libm: values agree to the tolerances written in the test, not bit for bit.
   One deliberate difference: the reference's TAApassKernel reads its 3x3 neighbourhood from the buffer it is overwriting (a
   race - its output changes from run to run, which is why the golden file stores a per-pixel [min, max] over 12 executions);
   TaaPass below reads the unmodified input, as the product's taaKernel does.
*/
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace orcf
{

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct U4 { uint32_t x, y, z, w; };
struct V2 { float x, y; };

static inline V3 operator+( V3 a, V3 b ) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
static inline V3 operator-( V3 a, V3 b ) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
static inline V3 operator*( V3 a, V3 b ) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
static inline V3 operator*( V3 a, float s ) { return { a.x * s, a.y * s, a.z * s }; }
static inline V3 operator*( float s, V3 a ) { return { a.x * s, a.y * s, a.z * s }; }
static inline float Dot( V3 a, V3 b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float Len( V3 a ) { return sqrtf( Dot( a, a ) ); }
static inline V3 Norm( V3 a ) { const float l = 1.0f / Len( a ); return a * l; }
static inline V3 Xyz( V4 a ) { return { a.x, a.y, a.z }; }
static inline V3 Min3( V3 a, float b ) { return { std::min( a.x, b ), std::min( a.y, b ), std::min( a.z, b ) }; }
static inline V3 Max3( V3 a, V3 b ) { return { std::max( a.x, b.x ), std::max( a.y, b.y ), std::max( a.z, b.z ) }; }
static inline V3 Clamp3( V3 v, V3 lo, V3 hi ) { return { std::min( std::max( v.x, lo.x ), hi.x ), std::min( std::max( v.y, lo.y ), hi.y ), std::min( std::max( v.z, lo.z ), hi.z ) }; }
static inline uint32_t Bits( float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }
static inline float AsFloat( uint32_t u ) { float f; memcpy( &f, &u, 4 ); return f; }
static inline V4 operator*( V4 a, float s ) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
static inline V4 operator+( V4 a, V4 b ) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }

/* tools_shared.h:131-137 */
static inline V3 UnpackNormal2( uint32_t pi )
{
	const uint32_t x = (pi >> 2u) & 1023u, y = (pi >> 12u) & 1023u, z = pi >> 22u;
	return { x * (1.0f / 511.0f) - 1, y * (1.0f / 511.0f) - 1, z * (1.0f / 511.0f) - 1 };
}
/* tools_shared.h:141-157 */
static inline V3 RGBToYCoCg( V3 RGB )
{
	const V3 rgb = Min3( RGB, 4.0f );
	return { Dot( rgb, { 1, 2, 1 } ) * 0.25f, Dot( rgb, { 2, 0, -2 } ) * 0.25f + (0.5f * 256.0f / 255.0f), Dot( rgb, { -1, 2, -1 } ) * 0.25f + (0.5f * 256.0f / 255.0f) };
}
static inline V3 YCoCgToRGB( V3 c )
{
	const float Y = c.x, Co = c.y - (0.5f * 256.0f / 255.0f), Cg = c.z - (0.5f * 256.0f / 255.0f);
	return { Y + Co - Cg, Y + Cg, Y - Co - Cg };
}
/* tools_shared.h:159-162 */
static inline float Luminance( V3 rgb ) { return 0.299f * std::min( rgb.x, 10.0f ) + 0.587f * std::min( rgb.y, 10.0f ) + 0.114f * std::min( rgb.z, 10.0f ); }
/* tools_shared.h:171-186 */
static inline V3 RGB32toHDR( uint32_t c ) { return { (float)(c >> 22) * (1.0f / 1023.0f), (float)((c >> 11) & 2047) * (1.0f / 2047.0f), (float)(c & 2047) * (1.0f / 2047.0f) }; }
static inline V3 RGB32toHDRmin1( uint32_t c )
{
	return { (float)std::max( 1u, c >> 22 ) * (1.0f / 1023.0f), (float)std::max( 1u, (c >> 11) & 2047 ) * (1.0f / 2047.0f), (float)std::max( 1u, c & 2047 ) * (1.0f / 2047.0f) };
}
/* tools_shared.h:237-262 */
static inline V4 CombineToFloat4( V3 A, V3 B )
{
	const uint32_t Ar = (uint32_t)(std::min( A.x, 31.999f ) * 2048.0f), Ag = (uint32_t)(std::min( A.y, 31.999f ) * 2048.0f), Ab = (uint32_t)(std::min( A.z, 31.999f ) * 2048.0f);
	const uint32_t Br = (uint32_t)(std::min( B.x, 31.999f ) * 2048.0f), Bg = (uint32_t)(std::min( B.y, 31.999f ) * 2048.0f), Bb = (uint32_t)(std::min( B.z, 31.999f ) * 2048.0f);
	return { AsFloat( (Ar << 16) + Ag ), AsFloat( Ab ), AsFloat( (Br << 16) + Bg ), AsFloat( Bb ) };
}
static inline V3 GetDirect( V4 X ) { const uint32_t v0 = Bits( X.x ), v1 = Bits( X.y ); return { (float)(v0 >> 16) * (1.0f / 2048.0f), (float)(v0 & 65535) * (1.0f / 2048.0f), (float)v1 * (1.0f / 2048.0f) }; }
static inline V3 GetIndirect( V4 X ) { const uint32_t v2 = Bits( X.z ), v3 = Bits( X.w ); return { (float)(v2 >> 16) * (1.0f / 2048.0f), (float)(v2 & 65535) * (1.0f / 2048.0f), (float)v3 * (1.0f / 2048.0f) }; }
static inline float OneOverPow2( int p ) { return AsFloat( (uint32_t)(127 - p) << 23 ); }
/* sampling_shared.h:22-27 */
static inline float MitchellNetravali( float v )
{
	const float B = 1.0f / 3.0f, C = 1.0f / 3.0f, x = fabsf( v ), x2 = x * x, x3 = x2 * x;
	if (x < 1) return (1.0f / 6.0f) * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B));
	else if (x < 2) return 1.0f / 6.0f * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C));
	return 0.0f;
}
/* sampling_shared.h:111-115 */
static inline V4 ReadWorldPos( const V4* buffer, int x, int y, int w, int h )
{
	if (x >= 0 && y >= 0 && x < w && y < h) return buffer[x + y * w];
	return { 1e20f, 1e20f, 1e20f, AsFloat( 0 ) };
}
/* sampling_shared.h:117-134 */
static inline V3 ReadTexelBmitchellNetravali( const V4* buffer, float u, float v, int w, int h )
{
	const int x1 = (int)(u - 2.0f), y1 = (int)(v - 2.0f);
	float totalWeight = 0;
	V4 total = { 0, 0, 0, 0 };
	for (int y = y1; y < y1 + 4; y++) for (int x = x1; x < x1 + 4; x++) if (x >= 0 && y > 0 && x < w && y < h)
	{
		const float weight = MitchellNetravali( (float)x - u ) * MitchellNetravali( (float)y - v );
		total = total + buffer[x + y * w] * weight, totalWeight += weight;
	}
	return Xyz( total * (1.0f / totalWeight) );
}
/* sampling_shared.h:136-173 (the "relax" branch that is compiled: normal threshold 0.95) */
static inline V4 ReadTexelConsistent( const V4* buffer, const V4* prevWorldPos, V4 localPos, V3 localNormal, float u, float v, int w, int h )
{
	const int iu1 = (int)floorf( u ), iv1 = (int)floorf( v ), iu0 = std::max( 0, iu1 - 1 ), iv0 = std::max( 0, iv1 - 1 );
	if (iu1 >= w || iv1 >= h || iu1 < 0 || iv1 < 0) return { -1, -1, -1, -1 };
	const float fx = u - floorf( u ), fy = v - floorf( v );
	const V4 p0 = buffer[iu0 + iv0 * w], p1 = buffer[iu1 + iv0 * w], p2 = buffer[iu0 + iv1 * w], p3 = buffer[iu1 + iv1 * w];
	const uint32_t n0 = Bits( prevWorldPos[iu0 + iv0 * w].w ), n1 = Bits( prevWorldPos[iu1 + iv0 * w].w );
	const uint32_t n2 = Bits( prevWorldPos[iu0 + iv1 * w].w ), n3 = Bits( prevWorldPos[iu1 + iv1 * w].w );
	const uint32_t spec = Bits( localPos.w ) & 3;
	float w0 = (1 - fx) * (1 - fy), w1 = fx * (1 - fy), w2 = (1 - fx) * fy, w3 = 1 - (w0 + w1 + w2);
	if (Dot( UnpackNormal2( n0 ), localNormal ) < 0.95f || (n0 & 3) != spec) w0 = 0;
	if (Dot( UnpackNormal2( n1 ), localNormal ) < 0.95f || (n1 & 3) != spec) w1 = 0;
	if (Dot( UnpackNormal2( n2 ), localNormal ) < 0.95f || (n2 & 3) != spec) w2 = 0;
	if (Dot( UnpackNormal2( n3 ), localNormal ) < 0.95f || (n3 & 3) != spec) w3 = 0;
	const float sum = w0 + w1 + w2 + w3;
	if (sum == 0) return { -1, -1, -1, -1 };
	return (p0 * w0 + p1 * w1 + p2 * w2 + p3 * w3) * (1.0f / sum);
}
/* sampling_shared.h:175-215 (compiled branch: normal threshold 0.975) */
static inline bool ReadTexelConsistent2( const V4* buffer, const V4* prevWorldPos, V4 localPos, V3 localNormal, float u, float v, int w, int h, V3& direct, V3& indirect )
{
	const int iu1 = (int)floorf( u ), iv1 = (int)floorf( v ), iu0 = std::max( 0, iu1 - 1 ), iv0 = std::max( 0, iv1 - 1 );
	if (iu1 >= w || iv1 >= h || iu1 < 0 || iv1 < 0) return false;
	const float fx = u - floorf( u ), fy = v - floorf( v );
	const V4 p0 = buffer[iu0 + iv0 * w], p1 = buffer[iu1 + iv0 * w], p2 = buffer[iu0 + iv1 * w], p3 = buffer[iu1 + iv1 * w];
	const uint32_t n0 = Bits( prevWorldPos[iu0 + iv0 * w].w ), n1 = Bits( prevWorldPos[iu1 + iv0 * w].w );
	const uint32_t n2 = Bits( prevWorldPos[iu0 + iv1 * w].w ), n3 = Bits( prevWorldPos[iu1 + iv1 * w].w );
	const uint32_t spec = Bits( localPos.w ) & 3;
	float w0 = (1 - fx) * (1 - fy), w1 = fx * (1 - fy), w2 = (1 - fx) * fy, w3 = 1 - (w0 + w1 + w2);
	if (Dot( UnpackNormal2( n0 ), localNormal ) < 0.975f || (n0 & 3) != spec) w0 = 0;
	if (Dot( UnpackNormal2( n1 ), localNormal ) < 0.975f || (n1 & 3) != spec) w1 = 0;
	if (Dot( UnpackNormal2( n2 ), localNormal ) < 0.975f || (n2 & 3) != spec) w2 = 0;
	if (Dot( UnpackNormal2( n3 ), localNormal ) < 0.975f || (n3 & 3) != spec) w3 = 0;
	const float sum = w0 + w1 + w2 + w3;
	if (sum == 0) return false;
	const float r = 1.0f / sum;
	direct = (w0 * GetDirect( p0 ) + w1 * GetDirect( p1 ) + w2 * GetDirect( p2 ) + w3 * GetDirect( p3 )) * r;
	indirect = (w0 * GetIndirect( p0 ) + w1 * GetIndirect( p1 ) + w2 * GetIndirect( p2 ) + w3 * GetIndirect( p3 )) * r;
	return true;
}

/* finalize_shared.h:169-178 */
static inline float WorldDistance( int x, int y, V4 cur, const V4* prevWorldPos, int w, int h )
{
	const V4 p = ReadWorldPos( prevWorldPos, x, y, w, h );
	if ((Bits( p.w ) & 3) != 1) return 1e21f;
	if (Dot( UnpackNormal2( Bits( cur.w ) ), UnpackNormal2( Bits( p.w ) ) ) < 0.85f) return 1e21f;
	return Len( { cur.x - p.x, cur.y - p.y, cur.z - p.z } );
}
/* finalize_shared.h:179-199 */
static inline float FineWorldDistance( float px, float py, V4 cur, const V4* prevWorldPos, int w, int h )
{
	const int x0 = (int)px, y0 = (int)py;
	const float fx = px - floorf( px ), fy = py - floorf( py );
	const float w0 = (1 - fx) * (1 - fy), w1 = fx * (1 - fy), w2 = (1 - fx) * fy, w3 = fx * fy;
	const float d0 = WorldDistance( x0, y0, cur, prevWorldPos, w, h ), d1 = WorldDistance( x0 + 1, y0, cur, prevWorldPos, w, h );
	const float d2 = WorldDistance( x0, y0 + 1, cur, prevWorldPos, w, h ), d3 = WorldDistance( x0 + 1, y0 + 1, cur, prevWorldPos, w, h );
	float totalWeight = 0, totalDist = 0;
	if (d0 < 1e20f) totalDist += d0 * w0, totalWeight += w0;
	if (d1 < 1e20f) totalDist += d1 * w1, totalWeight += w1;
	if (d2 < 1e20f) totalDist += d2 * w2, totalWeight += w2;
	if (d3 < 1e20f) totalDist += d3 * w3, totalWeight += w3;
	return totalWeight == 0 ? 1e20f : totalDist / totalWeight;
}

struct Chain
{
	int w, h, samplesTaken, camIsStationary, taa;
	float directClamp, indirectClamp, j0, j1, prevj0, prevj1;
	float prevView[17];
	const V4* accumulator; U4* features; const V4* worldPos; const V4* prevWorldPos; const V4* deltaDepth; const V4* prevMoments;
	V4* shading; V2* motion; V4* moments;
};

/* prepareFilter + prepareFilterKernel (finalize_shared.h:217-314) */
static inline void Prepare( const Chain& c )
{
	const int w = c.w, h = c.h;
	const float* pv = c.prevView;
	const V3 pos = { pv[0], pv[1], pv[2] }, p1 = { pv[3], pv[4], pv[5] }, p2 = { pv[6], pv[7], pv[8] }, p3 = { pv[9], pv[10], pv[11] };
	const V3 centre = 0.5f * (p2 + p3), direction = Norm( centre - pos ), right = Norm( p2 - p1 ), up = Norm( p3 - p1 );
	const float lenReci = h / Len( p3 - p1 );
	const V4 prevPos = { pos.x, pos.y, pos.z, -(Dot( pos, direction ) - Dot( centre, direction )) };
	const V3 prevE = direction;
	const V4 prevRight = { right.x * lenReci, right.y * lenReci, right.z * lenReci, Dot( p1, right ) * lenReci };
	const V4 prevUp = { up.x * lenReci, up.y * lenReci, up.z * lenReci, Dot( p1, up ) * lenReci };
	const float pixelValueScale = 1.0f / (float)c.samplesTaken;
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
	{
		const int pixelIdx = x + y * w;
		const V3 direct = Xyz( c.accumulator[pixelIdx] ) * pixelValueScale;
		const U4 feat = c.features[pixelIdx];
		const V4 lwp = c.worldPos[pixelIdx];
		const V3 albedo = RGB32toHDRmin1( feat.x );
		const V3 indirect = Xyz( c.accumulator[pixelIdx + w * h] ) * pixelValueScale;
		const V3 reci = { 1.0f / albedo.x, 1.0f / albedo.y, 1.0f / albedo.z };
		const V3 directLight = Min3( direct * reci, c.directClamp ), indirectLight = Min3( indirect * reci, c.indirectClamp );
		c.shading[pixelIdx] = CombineToFloat4( directLight, indirectLight );
		float lumDirect = Luminance( directLight ), lumDirect2 = lumDirect * lumDirect;
		float lumIndirect = Luminance( indirectLight ), lumIndirect2 = lumIndirect * lumIndirect;
		V2 prev;
		if (((feat.w >> 4) & 3) == 0)
		{
			const V3 D = Norm( Xyz( lwp ) - Xyz( prevPos ) );
			const V3 S = Xyz( prevPos ) + D * (prevPos.w / Dot( prevE, D ));
			prev = { Dot( S, Xyz( prevRight ) ) - prevRight.w - c.j0, Dot( S, Xyz( prevUp ) ) - prevUp.w - c.j1 };
		}
		else
		{
			prev = { (float)x, (float)y };
			if (!c.camIsStationary)
			{
				const V4 pw = c.prevWorldPos[pixelIdx];
				float bestDist = Len( { lwp.x - pw.x, lwp.y - pw.y, lwp.z - pw.z } ), stepSize = 5.0f;
				const float ox = c.j0 - c.prevj0, oy = c.j1 - c.prevj1;
				int iter = 0;
				while (1)
				{
					// RefineHistoryPos (finalize_shared.h:200-216)
					int tap = 0;
					const float cx = prev.x, cy = prev.y;
					float d;
					d = FineWorldDistance( cx - stepSize + ox, cy + oy, lwp, c.prevWorldPos, w, h );
					if (d < bestDist) bestDist = d, prev = { cx - stepSize, cy }, tap = 1;
					d = FineWorldDistance( cx + stepSize + ox, cy + oy, lwp, c.prevWorldPos, w, h );
					if (d < bestDist) bestDist = d, prev = { cx + stepSize, cy }, tap = 2;
					d = FineWorldDistance( cx + ox, cy - stepSize + oy, lwp, c.prevWorldPos, w, h );
					if (d < bestDist) bestDist = d, prev = { cx, cy - stepSize }, tap = 3;
					d = FineWorldDistance( cx + ox, cy + stepSize + oy, lwp, c.prevWorldPos, w, h );
					if (d < bestDist) bestDist = d, prev = { cx, cy + stepSize }, tap = 4;
					if (tap == 0) { stepSize *= 0.45f; if (stepSize < 0.05f) break; }
					if (++iter == 25) break;
				}
			}
		}
		prev.x += 0.5f, prev.y += 0.5f;
		uint32_t fw = feat.w;
		if (prev.x >= 0 && prev.x < w && prev.y >= 0 && prev.y < h)
		{
			const V4 history = ReadTexelConsistent( c.prevMoments, c.prevWorldPos, lwp, UnpackNormal2( feat.y ), prev.x, prev.y, w, h );
			if (history.x > -1)
			{
				lumDirect = 0.2f * lumDirect + 0.8f * history.x, lumDirect2 = 0.2f * lumDirect2 + 0.8f * history.y;
				lumIndirect = 0.2f * lumIndirect + 0.8f * history.z, lumIndirect2 = 0.2f * lumIndirect2 + 0.8f * history.w;
				if ((fw & 15) < 15) fw++;
			}
			else fw &= 0xfffffff0u;
		}
		else fw &= 0xfffffff0u;
		c.features[pixelIdx].w = fw;
		c.motion[pixelIdx] = prev;
		c.moments[pixelIdx] = { lumDirect, lumDirect2, lumIndirect, lumIndirect2 };
	}
}

/* applyFilterKernel (finalize_shared.h:320-476): A -> C, phase 1 also blends with B, the previous frame's phase-1 output */
static inline void ApplyFilter( const Chain& c, const V4* A, const V4* B, V4* C, int phase, int lastPass )
{
	const int w = c.w, h = c.h, step = 1 << (phase - 1);
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
	{
		const int pixelIdx = x + y * w;
		const U4 lf = c.features[pixelIdx];
		const V4 localPos = c.worldPos[pixelIdx];
		const V3 localNormal = UnpackNormal2( lf.y ), localColor = RGB32toHDR( lf.x );
		const uint32_t localMatID = lf.w >> 4;
		float dirW = 1, indW = 1;
		const V4 combined = A[pixelIdx];
		V3 dirSum = GetDirect( combined ), indSum = GetIndirect( combined );
		const float localDirect = Luminance( GetDirect( combined ) ), localIndirect = Luminance( GetIndirect( combined ) );
		const float localDepth = AsFloat( lf.z );
		const float localDdx = c.deltaDepth[pixelIdx].z, localDdy = c.deltaDepth[pixelIdx].w;
		const float sigma = 10.0f * OneOverPow2( phase - 1 );
		const float factor = (lf.w & 15) == 0 ? 400.0f : 1.0f;
		const V4 m = c.moments[pixelIdx];
		const float var_dir = m.y - m.x * m.x, var_ind = m.w - m.z * m.z;
		const float rdir = -1.0f / (sigma * factor * sqrtf( var_dir + 0.00001f ) + 0.00001f);
		const float rind = -1.0f / (sigma * factor * sqrtf( var_ind + 0.00001f ) + 0.00001f);
		for (int vv = -2; vv <= 2; vv++)
		{
			const int v = vv * step + y, r = abs( vv ) == 2 ? 1 : 2;
			if (v >= 0 && v < h) for (int uu = -r; uu <= r; uu++) if (uu != 0 || vv != 0)
			{
				const int u = std::min( std::max( uu * step + x, 0 ), w - 1 );
				const int li = u + v * w;
				const V4 nc = A[li];
				const U4 nf = c.features[li];
				const float w_dist = (uu * uu + vv * vv) * (-1.0f / 7.5f);
				const V3 nDirect = GetDirect( nc ), nIndirect = GetIndirect( nc );
				float w_normal = powf( std::max( 0.0f, Dot( UnpackNormal2( nf.y ), localNormal ) ), 128 );
				const float expected = localDepth + localDdx * (float)(uu * step) + localDdy * (float)(vv * step);
				const float depthError = fabsf( expected - AsFloat( nf.z ) );
				const float expectedDiff = fabsf( expected - localDepth );
				const float w_depth = depthError / std::max( 0.00001f, (0.5f + phase * 0.5f) * expectedDiff );
				w_normal *= ((nf.w >> 4) != localMatID) ? 0.0001f : Dot( localColor, RGB32toHDR( nf.x ) );
				float wd = w_normal * expf( fabsf( localDirect - Luminance( nDirect ) ) * rdir + w_dist - w_depth );
				float wi = w_normal * expf( fabsf( localIndirect - Luminance( nIndirect ) ) * rind + w_dist - w_depth );
				if (!std::isfinite( wd )) wd = 0;
				if (!std::isfinite( wi )) wi = 0;
				dirSum = dirSum + nDirect * wd, dirW += wd;
				indSum = indSum + nIndirect * wi, indW += wi;
			}
		}
		V3 dirF = dirSum * (1.0f / std::max( 0.0001f, dirW )), indF = indSum * (1.0f / std::max( 0.0001f, indW ));
		if (phase == 1)
		{
			const V2 pp = c.motion[pixelIdx];
			const int px = (int)pp.x, py = (int)pp.y;
			if (px >= 0 && px < w && py >= 0 && py < h)
			{
				V3 prevDirect, prevIndirect;
				if (ReadTexelConsistent2( B, c.prevWorldPos, localPos, localNormal, pp.x, pp.y, w, h, prevDirect, prevIndirect ))
				{
					prevDirect = RGBToYCoCg( prevDirect ), prevIndirect = RGBToYCoCg( prevIndirect );
					V3 dirAvg = RGBToYCoCg( dirF ), dirVar = dirAvg * dirAvg, indAvg = RGBToYCoCg( indF ), indVar = indAvg * indAvg;
					auto tap = [&]( int idx ) {
						const V3 f = RGBToYCoCg( GetDirect( A[idx] ) ), g = RGBToYCoCg( GetIndirect( A[idx] ) );
						dirAvg = dirAvg + f, dirVar = dirVar + f * f, indAvg = indAvg + g, indVar = indVar + g * g; };
					if (x > 1)
					{
						if (y > 1) tap( pixelIdx - w - 1 );
						tap( pixelIdx - 1 );
						if (y < h - 1) tap( pixelIdx + w - 1 );
					}
					if (y > 1) tap( pixelIdx - w );
					if (y < h - 1) tap( pixelIdx + w );
					if (x < w - 1)
					{
						if (y > 1) tap( pixelIdx + 1 - w );
						tap( pixelIdx + 1 );
						if (y < h - 1) tap( pixelIdx + 1 + w );
					}
					dirAvg = dirAvg * (1.0f / 9.0f), dirVar = dirVar * (1.0f / 9.0f), indAvg = indAvg * (1.0f / 9.0f), indVar = indVar * (1.0f / 9.0f);
					V3 sDir = Max3( { 0, 0, 0 }, dirVar - dirAvg * dirAvg ), sInd = Max3( { 0, 0, 0 }, indVar - indAvg * indAvg );
					sDir = { sqrtf( sDir.x ), sqrtf( sDir.y ), sqrtf( sDir.z ) }, sInd = { sqrtf( sInd.x ), sqrtf( sInd.y ), sqrtf( sInd.z ) };
					prevDirect = Clamp3( prevDirect, dirAvg - 0.75f * sDir, dirAvg + 0.75f * sDir );
					prevIndirect = Clamp3( prevIndirect, indAvg - 0.75f * sInd, indAvg + 0.75f * sInd );
					dirF = dirF * 0.1f + YCoCgToRGB( prevDirect ) * 0.9f;
					indF = indF * 0.1f + YCoCgToRGB( prevIndirect ) * 0.9f;
				}
			}
		}
		if (lastPass)
		{
			const V3 comb = (dirF + indF) * RGB32toHDR( lf.x );
			C[pixelIdx] = { sqrtf( comb.x ), sqrtf( comb.y ), sqrtf( comb.z ), 1 };
		}
		else C[pixelIdx] = CombineToFloat4( dirF, indF );
	}
}

/* TAApassKernel (finalize_shared.h:498-541); reads 'in', writes 'out' (see the header comment about the reference's in-place race) */
static inline void TaaPass( const V4* in, V4* out, const V4* prevPixels, const V2* motion, int w, int h )
{
	for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
	{
		const int pixelIdx = x + y * w;
		V3 pixel = Xyz( in[pixelIdx] );
		const float pu = motion[pixelIdx].x - 0.5f, pvv = motion[pixelIdx].y - 0.5f;
		if (pu >= 0 && pu < w && pvv >= 0 && pvv < h)
		{
			const V3 newPixel = RGBToYCoCg( pixel );
			V3 history = RGBToYCoCg( ReadTexelBmitchellNetravali( prevPixels, pu, pvv, w, h ) );
			V3 avg = newPixel, var = newPixel * newPixel;
			auto tap = [&]( int idx ) { const V3 f = RGBToYCoCg( Xyz( in[idx] ) ); avg = avg + f, var = var + f * f; };
			if (x > 1)
			{
				if (y > 1) tap( pixelIdx - w - 1 );
				tap( pixelIdx - 1 );
				if (y < h - 1) tap( pixelIdx + w - 1 );
			}
			if (y > 1) tap( pixelIdx - w );
			if (y < h - 1) tap( pixelIdx + w );
			if (x < w - 1)
			{
				if (y > 1) tap( pixelIdx + 1 - w );
				tap( pixelIdx + 1 );
				if (y < h - 1) tap( pixelIdx + 1 + w );
			}
			avg = avg * (1.0f / 9.0f), var = var * (1.0f / 9.0f);
			V3 sigma = Max3( { 0, 0, 0 }, var - avg * avg );
			sigma = { sqrtf( sigma.x ), sqrtf( sigma.y ), sqrtf( sigma.z ) };
			history = Clamp3( history, avg - 1.25f * sigma, avg + 1.25f * sigma );
			pixel = YCoCgToRGB( newPixel * 0.1f + history * 0.9f );
			if (std::isnan( pixel.x + pixel.y + pixel.z )) pixel = YCoCgToRGB( newPixel );
		}
		const V3 o = Min3( pixel, 10.0f );
		out[pixelIdx] = { o.x, o.y, o.z, 0 };
	}
}

/* unsharpenTAAKernel (finalize_shared.h:554-583); border pixels of the target are left untouched */
static inline void UnsharpenTaa( const V4* px, V4* target, int w, int h )
{
	for (int y = 1; y < h - 1; y++) for (int x = 1; x < w - 1; x++)
	{
		const V4 c = px[x + y * w];
		const V4 p0 = px[x - 1 + (y - 1) * w], p1 = px[x + (y - 1) * w], p2 = px[x + 1 + (y - 1) * w], p3 = px[x + 1 + y * w];
		const V4 p4 = px[x + 1 + (y + 1) * w], p5 = px[x + (y + 1) * w], p6 = px[x - 1 + (y + 1) * w], p7 = px[x - 1 + y * w];
		const V4 blur = p0 * 0.35f + p1 * 0.5f + p2 * 0.35f + p3 * 0.5f + p4 * 0.35f + p5 * 0.5f + p6 * 0.35f + p7 * 0.5f;
		const V4 sharp = c * 2.7f + blur * -0.5f;
		const V4 q = { std::max( c.x, sharp.x ), std::max( c.y, sharp.y ), std::max( c.z, sharp.z ), std::max( c.w, sharp.w ) };
		target[x + y * w] = { q.x * q.x, q.y * q.y, q.z * q.z, 0 };
	}
}

/* finalizeNoTAAKernel (finalize_shared.h:589-600) */
static inline void FinalizeNoTaa( const V4* px, V4* target, int w, int h )
{
	for (int y = 1; y < h - 1; y++) for (int x = 1; x < w - 1; x++)
	{
		const V4 c = px[x + y * w];
		target[x + y * w] = { sqrtf( c.x ), sqrtf( c.y ), sqrtf( c.z ), 0 };
	}
}

} // namespace orcf
