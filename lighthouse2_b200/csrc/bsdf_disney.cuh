/* bsdf_disney.cuh - the "principled" material model of the reference cores (Setting "bsdf" = 1), device code.

   Behaviour restated from lib/sharedBSDFs/disney.h:151-360 (lobe mix: diffuse/subsurface, sheen, anisotropic GGX
   specular, GTR1 clearcoat; rough dielectric when TRANSMISSION wins the coin), lib/sharedBSDFs/ggxmdf.h:32-227
   (GGX with visible-normal sampling, GTR1, roughness -> alpha) and lib/sharedBSDFs/frosted.h:20-118 (rough glass).
   It is what kernels/bsdf.h:7-21 of every stock reference core compiles; lambert.h (a14, bsdf = 0) is the other option.

   Organisation here: the twelve 8-bit material parameters are unpacked once into DisneyParams; a Frame carries the
   tangent basis; every lobe is a small function returning its value and density in tangent space. Quirks of the
   reference that are visible in its output are kept and marked "(ref)":
     - in EvaluateBSDF the sheen lobe, when present, REPLACES the diffuse value (both write the same variable);
     - the clearcoat density ignores the view direction (GTR1 pdf = D * |m.z|);
     - the dielectric branch hands eta (not 1/eta) to the refracted-direction helper.
   Where the reference leaves an output unassigned (early returns of its sampling helpers) this code returns
   pdf = 0 / value = 0, which ends the path.
*/
#pragma once

namespace lh2b
{

struct DisneyParams
{
	float metallic, subsurface, specular, roughness, specTint, anisotropic, sheen, sheenTint, clearcoat, clearcoatGloss, transmission, eta;
	float3 color, tint, transmittance; float luminance;
};

struct Frame { float3 n, t, b; };	// t, b as the reference builds them: b = normalize( n x iT ), t = normalize( n x b )

__device__ __forceinline__ Frame MakeFrame( const float3 n, const float3 iT )
{
	Frame f;
	f.n = n, f.b = normalize3( cross( n, iT ) ), f.t = normalize3( cross( n, f.b ) );
	return f;
}
__device__ __forceinline__ float3 ToLocal( const Frame& f, const float3 v ) { return make_float3( dot( v, f.t ), dot( v, f.b ), dot( v, f.n ) ); }
__device__ __forceinline__ float3 ToWorld( const Frame& f, const float3 v ) { return v.x * f.t + v.y * f.b + v.z * f.n; }

__device__ __forceinline__ float Lerp( const float a, const float b, const float t ) { return a + t * (b - a); }
__device__ __forceinline__ float MixClamped( const float a, const float b, const float t ) { return t <= 0 ? a : t >= 1 ? b : Lerp( a, b, t ); }	// tools_shared.h:98
__device__ __forceinline__ float SchlickWeight( const float u ) { const float m = fmaxf( 0.0f, fminf( 1.0f, 1.0f - u ) ), m2 = m * m; return m2 * m2 * m; }

__device__ __forceinline__ DisneyParams UnpackDisney( const float3 color, const float3 transmittance, const float4 tint, const uint4 q )
{
	DisneyParams d;
	d.metallic = char2flt( q.x, 0 ), d.subsurface = char2flt( q.x, 8 ), d.specular = char2flt( q.x, 16 ), d.roughness = fmaxf( 0.001f, char2flt( q.x, 24 ) );
	d.specTint = char2flt( q.y, 0 ), d.anisotropic = char2flt( q.y, 8 ), d.sheen = char2flt( q.y, 16 ), d.sheenTint = char2flt( q.y, 24 );
	d.clearcoat = char2flt( q.z, 0 ), d.clearcoatGloss = char2flt( q.z, 8 ), d.transmission = char2flt( q.z, 16 ), d.eta = __uint_as_float( q.w );
	d.color = color, d.transmittance = transmittance, d.tint = make_float3( tint.x, tint.y, tint.z ), d.luminance = tint.w;
	return d;
}

/* ---- microfacet distributions (ggxmdf.h) -------------------------------------------------------------------- */
__device__ __forceinline__ void AlphaFromRoughness( const float roughness, const float anisotropy, float& ax, float& ay )
{
	const float r2 = roughness * roughness, aspect = sqrtf( 1.0f + anisotropy * (anisotropy < 0 ? 0.9f : -0.9f) );
	ax = fmaxf( 0.001f, r2 / aspect ), ay = fmaxf( 0.001f, r2 * aspect );
}
__device__ __forceinline__ float GgxD( const float3 m, const float ax, const float ay )
{
	if (m.z == 0) return ax * ax * INVPI_F;
	const float c2 = m.z * m.z, s = sqrtf( fmaxf( 0.0f, 1 - c2 ) ), t2 = (1.0f - c2) / c2;
	float stretch;
	if (ax == ay || s == 0.0f) stretch = 1.0f / (ax * ax);
	else stretch = sqr( m.x / (s * ax) ) + sqr( m.y / (s * ay) );
	return 1.0f / (PI_F * ax * ay * sqr( c2 ) * sqr( 1.0f + t2 * stretch ));
}
__device__ __forceinline__ float GgxLambda( const float3 v, const float ax, const float ay )
{
	if (v.z == 0) return 0;
	const float c2 = v.z * v.z, s = sqrtf( fmaxf( 0.0f, 1 - c2 ) );
	float proj;
	if (ax == ay || s == 0.0f) proj = ax;
	else proj = sqrtf( sqr( (v.x * ax) / s ) + sqr( (v.y * ay) / s ) );
	const float t2 = (s * s) / c2;
	return (-1.0f + sqrtf( 1.0f + proj * proj * t2 )) * 0.5f;
}
__device__ __forceinline__ float GgxG( const float3 wi, const float3 wo, const float ax, const float ay ) { return 1.0f / (1.0f + GgxLambda( wo, ax, ay ) + GgxLambda( wi, ax, ay )); }
__device__ __forceinline__ float GgxVisiblePdf( const float3 v, const float3 m, const float ax, const float ay )
{
	if (v.z == 0.0f) return 0;
	return (1.0f / (1.0f + GgxLambda( v, ax, ay ))) * fabsf( dot( v, m ) ) * GgxD( m, ax, ay ) / fabsf( v.z );
}
/* visible-normal sampling, Heitz 2017 (ggxmdf.h:72-100) */
__device__ __forceinline__ float3 GgxSample( const float3 v, const float r0, const float r1, const float ax, const float ay )
{
	const float sgn = v.z < 0.0f ? -1.0f : 1.0f;
	const float3 s = normalize3( make_float3( sgn * v.x * ax, sgn * v.y * ay, sgn * v.z ) );
	const float3 t1 = v.z < 0.9999f ? normalize3( cross( s, make_float3( 0, 0, 1 ) ) ) : make_float3( 1, 0, 0 );
	const float3 t2 = cross( t1, s );
	const float a = 1.0f / (1.0f + s.z), r = sqrtf( r0 );
	const float phi = r1 < a ? (r1 / a * PI_F) : (PI_F + (r1 - a) / (1.0f - a) * PI_F);
	float p1, p2;
	sincosf( phi, &p2, &p1 );
	p1 *= r, p2 *= r * (r1 < a ? 1.0f : s.z);
	const float3 h = p1 * t1 + p2 * t2 + sqrtf( fmaxf( 0.0f, 1.0f - p1 * p1 - p2 * p2 ) ) * s;
	return normalize3( make_float3( h.x * ax, h.y * ay, fmaxf( 0.0f, h.z ) ) );
}
__device__ __forceinline__ float Gtr1D( const float3 m, const float alpha )
{
	const float a = fmaxf( 0.001f, fminf( alpha, 0.999f ) ), a2 = a * a;
	return ((a2 - 1.0f) / (PI_F * logf( a2 ))) * (1 / (1 + (a2 - 1) * sqr( m.z )));
}
__device__ __forceinline__ float Gtr1Lambda( const float3 v, const float alpha )
{
	if (v.z == 0) return 0;
	const float c2 = v.z * v.z, s = sqrtf( fmaxf( 0.0f, 1.0f - c2 ) );
	if (s == 0) return 0;
	const float cot2 = c2 / (s * s), cot = sqrtf( cot2 ), a2 = sqr( fmaxf( 0.001f, fminf( alpha, 0.999f ) ) );
	const float a = sqrtf( cot2 + a2 ), b = sqrtf( cot2 + 1.0f ), c = logf( cot + b ), d = logf( cot + a );
	return (a - b + cot * (c - d)) / (cot * logf( a2 ));
}
__device__ __forceinline__ float Gtr1G( const float3 wi, const float3 wo, const float alpha ) { return 1.0f / (1.0f + Gtr1Lambda( wo, alpha ) + Gtr1Lambda( wi, alpha )); }
__device__ __forceinline__ float3 Gtr1Sample( const float r0, const float r1, const float alpha )
{
	const float a = fmaxf( 0.001f, fminf( alpha, 0.999f ) ), a2 = a * a;
	const float c2 = (1.0f - powf( a2, 1.0f - r0 )) / (1.0f - a2), s = sqrtf( fmaxf( 0.0f, 1.0f - c2 ) );
	float sp, cp;
	sincosf( TWOPI_F * r1, &sp, &cp );
	return make_float3( cp * s, sp * s, sqrtf( c2 ) );
}

/* ---- reflective lobes (disney.h:33-50,92-150) ----------------------------------------------------------------- */
enum { LOBE_SPECULAR = 0, LOBE_CLEARCOAT = 1 };

template <int LOBE> __device__ __forceinline__ float3 LobeFresnel( const DisneyParams& d, const float3 o, const float3 h )
{
	const float fh = SchlickWeight( fabsf( dot( o, h ) ) );
	if (LOBE == LOBE_CLEARCOAT) return f3( MixClamped( 0.04f, 1.0f, fh ) * 0.25f * d.clearcoat );
	float3 v = (f3( 1.0f - d.specTint ) + d.specTint * d.tint) * (d.specular * 0.08f);
	v = (1.0f - d.metallic) * v + d.metallic * d.color;
	return (1.0f - fh) * v + f3( fh );
}
__device__ __forceinline__ float ClearcoatAlpha( const DisneyParams& d ) { return MixClamped( 0.1f, 0.001f, d.clearcoatGloss ); }

/* density and value of a reflective lobe for a given pair of directions; false: the lobe contributes nothing */
template <int LOBE> __device__ __forceinline__ bool EvalReflective( const DisneyParams& d, const float ax, const float ay, const float3 wo, const float3 wi, const float3 m,
	float3& value, float& pdf )
{
	if (wo.z == 0 || wi.z == 0) return false;
	const float cosOH = dot( wo, m );
	if (cosOH == 0) return false;
	const float D = LOBE == LOBE_SPECULAR ? GgxD( m, ax, ay ) : Gtr1D( m, ax );
	const float G = LOBE == LOBE_SPECULAR ? GgxG( wi, wo, ax, ay ) : Gtr1G( wi, wo, ax );
	value = LobeFresnel<LOBE>( d, wo, m ) * (D * G / fabsf( 4.0f * wo.z * wi.z ));
	pdf = (LOBE == LOBE_SPECULAR ? GgxVisiblePdf( wo, m, ax, ay ) : Gtr1D( m, ax ) * fabsf( m.z )) / fabsf( 4.0f * cosOH );
	return true;
}

/* sample a reflective lobe: wi, density and value WITHOUT the 1 / |4 wo.z wi.z| factor (the caller applies it) */
template <int LOBE> __device__ __forceinline__ void SampleReflective( const DisneyParams& d, const float r0, const float r1, const float ax, const float ay, const float3 wo,
	float3& wi, float& pdf, float3& value )
{
	value = f3( 0 ), pdf = 0, wi = make_float3( 0, 0, 1 );
	if (wo.z == 0) return;
	const float3 m = LOBE == LOBE_SPECULAR ? GgxSample( wo, r0, r1, ax, ay ) : Gtr1Sample( r0, r1, ax );
	wi = reflect3( wo * -1.0f, m );
	if (wi.z == 0) return;
	const float cosOH = dot( wo, m );
	pdf = (LOBE == LOBE_SPECULAR ? GgxVisiblePdf( wo, m, ax, ay ) : Gtr1D( m, ax ) * fabsf( m.z )) / fabsf( 4.0f * cosOH );
	if (pdf < 1.0e-6f) return;
	const float D = LOBE == LOBE_SPECULAR ? GgxD( m, ax, ay ) : Gtr1D( m, ax );
	const float G = LOBE == LOBE_SPECULAR ? GgxG( wi, wo, ax, ay ) : Gtr1G( wi, wo, ax );
	value = LobeFresnel<LOBE>( d, wo, m ) * (D * G);
}

/* ---- diffuse / subsurface and sheen, in world space (disney.h:113-149) ---------------------------------------- */
__device__ __forceinline__ float EvalDiffuse( const DisneyParams& d, const float3 n, const float3 wo, const float3 wi, const float3 m, float3& value )
{
	const float cosON = dot( n, wo ), cosIN = dot( n, wi ), cosIH = dot( wi, m );
	const float fl = SchlickWeight( cosIN ), fv = SchlickWeight( cosON );
	float fd = 0;
	if (d.subsurface != 1.0f)
	{
		const float fd90 = 0.5f + 2.0f * cosIH * cosIH * d.roughness;
		fd = MixClamped( 1.0f, fd90, fl ) * MixClamped( 1.0f, fd90, fv );
	}
	if (d.subsurface > 0)
	{
		const float fss90 = cosIH * cosIH * d.roughness;
		const float fss = MixClamped( 1.0f, fss90, fl ) * MixClamped( 1.0f, fss90, fv );
		const float ss = 1.25f * (fss * (1.0f / (fabsf( cosON ) + fabsf( cosIN )) - 0.5f) + 0.5f);
		fd = MixClamped( fd, ss, d.subsurface );
	}
	value = d.color * fd * INVPI_F * (1.0f - d.metallic);
	return fabsf( cosIN ) * INVPI_F;
}
__device__ __forceinline__ float EvalSheen( const DisneyParams& d, const float3 wi, const float3 m, float3& value )
{
	const float fh = SchlickWeight( dot( wi, m ) );
	value = (f3( 1.0f - d.sheenTint ) + d.sheenTint * d.tint) * (fh * d.sheen * (1.0f - d.metallic));
	return 1.0f / (2 * PI_F);
}

/* ---- rough dielectric (frosted.h) ------------------------------------------------------------------------------ */
__device__ __forceinline__ float DielectricFresnel( const float cosI, const float eta, float& cosT )
{
	const float sinT2 = (1 - cosI * cosI) * (eta * eta);
	if (sinT2 > 1) { cosT = 0; return 1; }
	cosT = fminf( sqrtf( fmaxf( 1 - sinT2, 0.0f ) ), 1.0f );
	const float ci = fabsf( cosI );
	if (ci == 0 && cosT == 0) return 1;
	const float k0 = eta * cosT, k1 = eta * ci;
	return 0.5f * (sqr( (ci - k0) / (ci + k0) ) + sqr( (cosT - k1) / (cosT + k1) ));
}
__device__ __forceinline__ float3 GlassReflection( const float3 color, const float3 wo, const float3 wi, const float3 m, const float ax, const float ay, const float F )
{
	const float denom = fabsf( 4 * wo.z * wi.z );
	if (denom == 0) return f3( 0 );
	return color * (F * GgxD( m, ax, ay ) * GgxG( wi, wo, ax, ay ) / denom);
}
__device__ __forceinline__ float3 GlassRefraction( const float eta, const float3 color, const float3 wo, const float3 wi, const float3 m, const float ax, const float ay, const float T )
{
	if (wo.z == 0 || wi.z == 0) return f3( 0 );
	const float cosIH = dot( m, wi ), cosOH = dot( m, wo );
	const float dots = (cosIH * cosOH) / (wi.z * wo.z), sd = cosOH + eta * cosIH;
	if (fabsf( sd ) < 1.0e-6f) return f3( 0 );
	return color * (fabsf( dots ) * T * GgxD( m, ax, ay ) * GgxG( wi, wo, ax, ay ) / (sd * sd) * (eta * eta));	// never adjoint here
}
__device__ __forceinline__ float ReflectionJacobian( const float cosOH ) { return cosOH == 0 ? 0 : 1 / (4 * fabsf( cosOH )); }
__device__ __forceinline__ float RefractionJacobian( const float3 wo, const float3 wi, const float3 m, const float eta )
{
	const float cosIH = dot( m, wi ), cosOH = dot( m, wo ), sd = cosOH + eta * cosIH;
	if (fabsf( sd ) < 1.0e-6f) return 0;
	return fabsf( cosIH ) * sqr( eta / sd );
}
__device__ __forceinline__ float3 UpperHalf( const float3 h ) { return h.z < 0 ? h * -1.0f : h; }

/* ---- the two entry points the shade kernel calls ----------------------------------------------------------------- */
__device__ __forceinline__ void LobeWeights( const DisneyParams& d, float& wDiff, float& wSheen, float& wSpec, float& wCoat )
{
	wDiff = Lerp( d.luminance, 0, d.metallic ), wSheen = Lerp( d.sheen, 0, d.metallic ), wSpec = Lerp( d.specular, 1, d.metallic ), wCoat = d.clearcoat * 0.25f;
	const float r = 1.0f / (wDiff + wSheen + wSpec + wCoat);
	wDiff *= r, wSheen *= r, wSpec *= r, wCoat *= r;
}

/* EvaluateBSDF (disney.h:290-358): value and solid-angle density for the pair (wo, wi), world space */
__device__ __forceinline__ float3 EvaluateDisney( const DisneyParams& d, const float3 iN, const float3 iT, const float3 wo, const float3 wi, float& pdf )
{
	pdf = 0;
	if (d.transmission > 0.5f)
	{
		const Frame f = MakeFrame( iN, iT );
		const float3 o = ToLocal( f, wo ), i = ToLocal( f, wi );
		const float eta = o.z > 0 ? d.eta : (1.0f / d.eta);
		if (eta == 1) return f3( 0 );
		float ax, ay, jac, cosT;
		AlphaFromRoughness( d.roughness, d.anisotropic, ax, ay );
		float3 m, value;
		if (i.z * o.z >= 0)
		{
			m = UpperHalf( normalize3( i + o ) );
			const float c = dot( o, m ), F = DielectricFresnel( c, 1 / eta, cosT );
			value = GlassReflection( d.color, o, i, m, ax, ay, F );
			const float sum = F + (1 - F);
			pdf = sum != 0 ? F / sum : 1, jac = ReflectionJacobian( c );
		}
		else
		{
			m = UpperHalf( normalize3( o + eta * i ) );
			const float c = dot( o, m ), F = DielectricFresnel( c, 1 / eta, cosT );
			value = GlassRefraction( eta, d.color, o, i, m, ax, ay, 1 - F );
			const float sum = F + (1 - F);
			pdf = 1 - (sum != 0 ? F / sum : 1), jac = RefractionJacobian( o, i, m, eta );
		}
		pdf *= jac * GgxVisiblePdf( o, m, ax, ay );
		return value;
	}
	if (d.roughness <= 0.001f) return f3( 0 );	// specular vertices take no explicit connections
	float wDiff, wSheen, wSpec, wCoat;
	LobeWeights( d, wDiff, wSheen, wSpec, wCoat );
	float3 value = f3( 0 );
	if (wDiff + wSheen > 0)
	{
		const float3 m = normalize3( wi + wo );
		if (wDiff > 0) pdf += wDiff * EvalDiffuse( d, iN, wo, wi, m, value );
		if (wSheen > 0) pdf += wSheen * EvalSheen( d, wi, m, value );	// (ref) overwrites the diffuse value
	}
	if (wSpec + wCoat > 0)
	{
		const Frame f = MakeFrame( iN, iT );
		const float3 o = ToLocal( f, wo ), i = ToLocal( f, wi ), m = normalize3( o + i );
		float3 c;
		float lp;
		if (wSpec > 0)
		{
			float ax, ay;
			AlphaFromRoughness( d.roughness, d.anisotropic, ax, ay );
			if (EvalReflective<LOBE_SPECULAR>( d, ax, ay, o, i, m, c, lp )) if (lp > 0) pdf += wSpec * lp, value += c;
		}
		if (wCoat > 0)
		{
			const float a = ClearcoatAlpha( d );
			if (EvalReflective<LOBE_CLEARCOAT>( d, a, a, o, i, m, c, lp )) if (lp > 0) pdf += wCoat * lp, value += c;
		}
	}
	return value;
}

/* SampleBSDF (disney.h:151-288). r0 picks dielectric vs. opaque and then the lobe; r1 and the lobe-local remap of r0
   drive the lobe's sampler; r2 is the Fresnel coin of the dielectric. */
__device__ __forceinline__ float3 SampleDisney( const DisneyParams& d, float3 iN, const float3 N, const float3 iT, const float3 wo, const float distance,
	const float r0, const float r1, const float r2, float3& wiOut, float& pdf, bool& specular )
{
	pdf = 0;
	const float flip = (dot( wo, N ) < 0) ? -1 : 1;
	iN *= flip;
	const Frame f = MakeFrame( iN, iT );
	if (r0 < d.transmission)
	{
		specular = true;
		const float r3 = r0 / d.transmission;
		const float3 o = ToLocal( f, wo );
		const float eta = flip < 0 ? (1 / d.eta) : d.eta;
		if (eta == 1) return f3( 0 );
		const float3 beer = make_float3( expf( -d.transmittance.x * distance * 2.0f ), expf( -d.transmittance.y * distance * 2.0f ), expf( -d.transmittance.z * distance * 2.0f ) );
		float ax, ay, cosT, jac;
		AlphaFromRoughness( d.roughness, d.anisotropic, ax, ay );
		const float3 m = GgxSample( o, r1, r3, ax, ay );
		const float rcpEta = 1 / eta, c = fmaxf( -1.0f, fminf( dot( o, m ), 1.0f ) );
		const float F = DielectricFresnel( c, eta, cosT );
		float3 i, value;
		if (r2 < F)
		{
			i = reflect3( o * -1.0f, m );
			if (i.z * o.z <= 0) return f3( 0 );
			value = GlassReflection( d.color, o, i, m, ax, ay, F );
			pdf = F, jac = ReflectionJacobian( c );
		}
		else
		{
			// (ref) eta, not 1 / eta, goes into the direction formula (frosted.h:46-52 called from disney.h:196)
			const float3 w = c > 0 ? (eta * c - cosT) * m - eta * o : (eta * c + cosT) * m - eta * o;
			i = w * ((3 - dot( w, w )) * 0.5f);
			if (i.z * o.z > 0) return f3( 0 );
			value = GlassRefraction( rcpEta, d.color, o, i, m, ax, ay, 1 - F );
			pdf = 1 - F, jac = RefractionJacobian( o, i, m, rcpEta );
		}
		pdf *= jac * GgxVisiblePdf( o, m, ax, ay );
		if (pdf > 1.0e-6f) wiOut = ToWorld( f, i );
		return value * beer;
	}
	const float r3 = (r0 - d.transmission) / (1 - d.transmission);
	float wDiff, wSheen, wSpec, wCoat;
	LobeWeights( d, wDiff, wSheen, wSpec, wCoat );
	const float cdfDiff = wDiff, cdfSheen = wDiff + wSheen, cdfSpec = wDiff + wSheen + wSpec;
	float probability, lp;
	float3 value = f3( 0 ), c = f3( 0 ), wi;
	if (r3 < cdfSheen)
	{
		// cosine-weighted direction around the shading normal (common_functions.h:118-124), shared by diffuse and sheen
		const float ra = r3 / cdfSheen, term2 = sqrtf( 1 - r1 );
		float s, co;
		sincosf( TWOPI_F * ra, &s, &co );
		wi = (co * term2 * f.t) + (s * term2) * f.b + sqrtf( r1 ) * f.n;
		const float3 m = normalize3( wi + wo );
		if (r3 < cdfDiff) lp = EvalDiffuse( d, iN, wo, wi, m, value ), probability = wDiff * lp, wDiff = 0;
		else lp = EvalSheen( d, wi, m, value ), probability = wSheen * lp, wSheen = 0;
	}
	else
	{
		const float3 o = ToLocal( f, wo );
		float3 i;
		if (r3 < cdfSpec)
		{
			const float ra = (r3 - cdfSheen) / (cdfSpec - cdfSheen);
			float ax, ay;
			AlphaFromRoughness( d.roughness, d.anisotropic, ax, ay );
			SampleReflective<LOBE_SPECULAR>( d, ra, r1, ax, ay, o, i, lp, value );
			probability = wSpec * lp, wSpec = 0;
		}
		else
		{
			const float ra = (r3 - cdfSpec) / (1 - cdfSpec), a = ClearcoatAlpha( d );
			SampleReflective<LOBE_CLEARCOAT>( d, ra, r1, a, a, o, i, lp, value );
			probability = wCoat * lp, wCoat = 0;
		}
		value *= 1.0f / fabsf( 4.0f * o.z * i.z );
		wi = ToWorld( f, i );
	}
	// the lobes that were not sampled add their value and density for the chosen direction
	if (wDiff + wSheen > 0)
	{
		const float3 m = normalize3( wi + wo );
		if (wDiff > 0) probability += wDiff * EvalDiffuse( d, iN, wo, wi, m, c ), value += c;
		if (wSheen > 0) probability += wSheen * EvalSheen( d, wi, m, c ), value += c;
	}
	if (wSpec + wCoat > 0)
	{
		const float3 o = ToLocal( f, wo ), i = ToLocal( f, wi ), m = normalize3( o + i );
		if (wSpec > 0)
		{
			float ax, ay;
			AlphaFromRoughness( d.roughness, d.anisotropic, ax, ay );
			c = f3( 0 ), lp = 0;
			EvalReflective<LOBE_SPECULAR>( d, ax, ay, o, i, m, c, lp );
			probability += wSpec * lp, value += c;
		}
		if (wCoat > 0)
		{
			const float a = ClearcoatAlpha( d );
			c = f3( 0 ), lp = 0;
			EvalReflective<LOBE_CLEARCOAT>( d, a, a, o, i, m, c, lp );
			probability += wCoat * lp, value += c;
		}
	}
	wiOut = wi;
	pdf = probability > 1.0e-6f ? probability : 0;
	return value;
}

} // namespace lh2b
