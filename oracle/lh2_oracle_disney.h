/* lh2_oracle_disney.h - TEST INFRASTRUCTURE ONLY (CPU oracle). Never included, linked or called by the product.

   CPU restatement (plain float, libm) of the reference's principled material model, the BSDF its stock cores
   compile (lib/rendercore_optix7/kernels/bsdf.h:7-21):
     lib/sharedBSDFs/disney.h   :33-50 Fresnel terms | :92-111 evaluate_mf | :72-90 sample_mf | :113-149 diffuse, sheen
                                :151-288 SampleBSDF   | :290-358 EvaluateBSDF
     lib/sharedBSDFs/ggxmdf.h   :32-100 anisotropic GGX (D, lambda, G, visible-normal sampling) | :160-213 GTR1 | :220-227 alpha
     lib/sharedBSDFs/frosted.h  :20-118 rough dielectric helpers
   Pinned by: the reference's own shadeKernel compiled with these headers (oracle/_ref/libref_shade_disney_gpu.so, GPU
   tests) and the committed vectors tests/golden/shade_disney_reference_vectors.npz generated from it.
   Outputs the reference leaves unassigned on early returns are defined here as pdf = 0, value = 0 (path ends).
*/
#pragma once

namespace orc
{

struct Principled
{
	float metallic, subsurface, specular, roughness, specTint, anisotropic, sheen, sheenTint, clearcoat, clearcoatGloss, transmission, eta, lum;
	V3 color, tint, transmittance;
};

/* material_shared.h:19-33,118-119: hue of the base colour at unit luminance, and that luminance */
static inline void TintOf( V3 c, V3& tint, float& lum )
{
	const float X = fmaxf( 0.0f, 0.412453f * c.x + 0.357580f * c.y + 0.180423f * c.z ), Y = fmaxf( 0.0f, 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z );
	const float Z = fmaxf( 0.0f, 0.019334f * c.x + 0.119193f * c.y + 0.950227f * c.z );
	lum = Y, tint = v3( 1 );
	if (Y > 0)
	{
		const float r = 1.0f / Y, x = X * r, y = Y * r, z = Z * r;
		tint = v3( fmaxf( 0.0f, 3.240479f * x - 1.537150f * y - 0.498535f * z ), fmaxf( 0.0f, -0.969256f * x + 1.875992f * y + 0.041556f * z ),
			fmaxf( 0.0f, 0.055648f * x - 0.204043f * y + 1.057311f * z ) );
	}
}

static inline Principled MakePrincipled( const Shading& s )
{
	Principled p;
	p.metallic = Char2Flt( s.params[0], 0 ), p.subsurface = Char2Flt( s.params[0], 8 ), p.specular = Char2Flt( s.params[0], 16 ), p.roughness = Roughness( s );
	p.specTint = Char2Flt( s.params[1], 0 ), p.anisotropic = Char2Flt( s.params[1], 8 ), p.sheen = Char2Flt( s.params[1], 16 ), p.sheenTint = Char2Flt( s.params[1], 24 );
	p.clearcoat = Char2Flt( s.params[2], 0 ), p.clearcoatGloss = Char2Flt( s.params[2], 8 ), p.transmission = Transmission( s ), p.eta = Eta( s );
	p.color = s.color, p.transmittance = s.transmittance, p.tint = s.tint, p.lum = s.lum;
	return p;
}

namespace pr
{
static inline V3 cross( V3 a, V3 b ) { return v3( a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x ); }
static inline float lerpf( float a, float b, float t ) { return a + t * (b - a); }
static inline float mixc( float a, float b, float t ) { return t <= 0 ? a : t >= 1 ? b : lerpf( a, b, t ); }	// tools_shared.h:98
static inline float schlick( float u ) { const float m = fmaxf( 0.0f, fminf( 1.0f, 1.0f - u ) ), m2 = m * m; return m2 * m2 * m; }	// disney.h:33
static inline V3 local( V3 v, V3 n, V3 t, V3 b ) { return v3( dot( v, t ), dot( v, b ), dot( v, n ) ); }
static inline V3 world( V3 v, V3 n, V3 t, V3 b ) { return v.x * t + v.y * b + v.z * n; }

/* ggxmdf.h:220-227 */
static inline void alphas( float roughness, float aniso, float& ax, float& ay )
{
	const float sq = roughness * roughness, aspect = sqrtf( 1.0f + aniso * (aniso < 0 ? 0.9f : -0.9f) );
	ax = fmaxf( 0.001f, sq / aspect ), ay = fmaxf( 0.001f, sq * aspect );
}
/* ggxmdf.h:32-43 */
static inline float ggxD( V3 m, float ax, float ay )
{
	if (m.z == 0) return sqr( ax ) * INVPI_;
	const float c2 = sqr( m.z ), st = sqrtf( fmaxf( 0.0f, 1 - c2 ) ), t2 = (1.0f - c2) / c2;
	const float sr = (ax == ay || st == 0.0f) ? 1.0f / sqr( ax ) : sqr( m.x / (st * ax) ) + sqr( m.y / (st * ay) );
	return 1.0f / (PI_ * ax * ay * sqr( c2 ) * sqr( 1.0f + t2 * sr ));
}
/* ggxmdf.h:45-56 */
static inline float ggxLambda( V3 v, float ax, float ay )
{
	if (v.z == 0) return 0;
	const float c2 = v.z * v.z, st = sqrtf( fmaxf( 0.0f, 1 - c2 ) );
	const float pr_ = (ax == ay || st == 0.0f) ? ax : sqrtf( sqr( (v.x * ax) / st ) + sqr( (v.y * ay) / st ) );
	return (-1.0f + sqrtf( 1.0f + sqr( pr_ ) * (sqr( st ) / c2) )) * 0.5f;
}
static inline float ggxG( V3 wi, V3 wo, float ax, float ay ) { return 1.0f / (1.0f + ggxLambda( wo, ax, ay ) + ggxLambda( wi, ax, ay )); }	// :58-61
/* ggxmdf.h:140-150: density of visible normals */
static inline float ggxPdf( V3 v, V3 m, float ax, float ay )
{
	if (v.z == 0.0f) return 0;
	return (1.0f / (1.0f + ggxLambda( v, ax, ay ))) * fabsf( dot( v, m ) ) * ggxD( m, ax, ay ) / fabsf( v.z );
}
/* ggxmdf.h:67-100 */
static inline V3 ggxSample( V3 v, float r0, float r1, float ax, float ay )
{
	const float sg = v.z < 0.0f ? -1.0f : 1.0f;
	const V3 st = normalize( v3( sg * v.x * ax, sg * v.y * ay, sg * v.z ) );
	const V3 t1 = v.z < 0.9999f ? normalize( cross( st, v3( 0, 0, 1 ) ) ) : v3( 1, 0, 0 ), t2 = cross( t1, st );
	const float a = 1.0f / (1.0f + st.z), r = sqrtf( r0 );
	const float phi = r1 < a ? (r1 / a * PI_) : (PI_ + (r1 - a) / (1.0f - a) * PI_);
	const float p1 = r * cosf( phi ), p2 = r * sinf( phi ) * (r1 < a ? 1.0f : st.z);
	const V3 h = p1 * t1 + p2 * t2 + sqrtf( fmaxf( 0.0f, 1.0f - p1 * p1 - p2 * p2 ) ) * st;
	return normalize( v3( h.x * ax, h.y * ay, fmaxf( 0.0f, h.z ) ) );
}
/* ggxmdf.h:160-213 (GTR1; alpha_y unused there) */
static inline float gtrD( V3 m, float alpha )
{
	const float a2 = sqr( fmaxf( 0.001f, fminf( alpha, 0.999f ) ) );
	return ((a2 - 1.0f) / (PI_ * logf( a2 ))) * (1 / (1 + (a2 - 1) * sqr( m.z )));
}
static inline float gtrLambda( V3 v, float alpha )
{
	if (v.z == 0) return 0;
	const float c2 = sqr( v.z ), st = sqrtf( fmaxf( 0.0f, 1.0f - c2 ) );
	if (st == 0) return 0;
	const float ct2 = c2 / sqr( st ), ct = sqrtf( ct2 ), a2 = sqr( fmaxf( 0.001f, fminf( alpha, 0.999f ) ) );
	const float a = sqrtf( ct2 + a2 ), b = sqrtf( ct2 + 1.0f );
	return (a - b + ct * (logf( ct + b ) - logf( ct + a ))) / (ct * logf( a2 ));
}
static inline float gtrG( V3 wi, V3 wo, float alpha ) { return 1.0f / (1.0f + gtrLambda( wo, alpha ) + gtrLambda( wi, alpha )); }
static inline float gtrPdf( V3 m, float alpha ) { return gtrD( m, alpha ) * fabsf( m.z ); }
static inline V3 gtrSample( float r0, float r1, float alpha )
{
	const float a2 = sqr( fmaxf( 0.001f, fminf( alpha, 0.999f ) ) );
	const float c2 = (1.0f - powf( a2, 1.0f - r0 )) / (1.0f - a2), st = sqrtf( fmaxf( 0.0f, 1.0f - c2 ) ), phi = TWOPI_ * r1;
	return v3( cosf( phi ) * st, sinf( phi ) * st, sqrtf( c2 ) );
}

/* disney.h:39-51 */
static inline V3 fresnelSpec( const Principled& p, V3 o, V3 h )
{
	V3 v = (v3( 1.0f - p.specTint ) + p.specTint * p.tint) * (p.specular * 0.08f);
	v = (1.0f - p.metallic) * v + p.metallic * p.color;
	const float f = schlick( fabsf( dot( o, h ) ) );
	return (1.0f - f) * v + v3( f );
}
static inline V3 fresnelCoat( const Principled& p, V3 o, V3 h ) { return v3( mixc( 0.04f, 1.0f, schlick( fabsf( dot( o, h ) ) ) ) * 0.25f * p.clearcoat ); }
static inline float coatAlpha( const Principled& p ) { return mixc( 0.1f, 0.001f, p.clearcoatGloss ); }	// disney.h:38

/* disney.h:92-111 evaluate_mf; coat selects GTR1 + clearcoat Fresnel. Returns the density, 0 = nothing (value untouched) */
static inline float evalMf( const Principled& p, bool coat, float ax, float ay, V3 wo, V3 wi, V3 m, V3& value )
{
	if (wo.z == 0 || wi.z == 0) return 0;
	const float coh = dot( wo, m );
	if (coh == 0) return 0;
	const float D = coat ? gtrD( m, ax ) : ggxD( m, ax, ay ), G = coat ? gtrG( wi, wo, ax ) : ggxG( wi, wo, ax, ay );
	value = (coat ? fresnelCoat( p, wo, m ) : fresnelSpec( p, wo, m )) * (D * G / fabsf( 4.0f * wo.z * wi.z ));
	return (coat ? gtrPdf( m, ax ) : ggxPdf( wo, m, ax, ay )) / fabsf( 4.0f * coh );
}
/* disney.h:72-90 sample_mf (value lacks the 1 / |4 wo.z wi.z| factor, applied by the caller) */
static inline void sampleMf( const Principled& p, bool coat, float r0, float r1, float ax, float ay, V3 wo, V3& wi, float& pdf, V3& value )
{
	value = v3( 0 ), pdf = 0, wi = v3( 0, 0, 1 );
	if (wo.z == 0) return;
	const V3 m = coat ? gtrSample( r0, r1, ax ) : ggxSample( wo, r0, r1, ax, ay );
	wi = reflect( wo * -1.0f, m );
	if (wi.z == 0) return;
	pdf = (coat ? gtrPdf( m, ax ) : ggxPdf( wo, m, ax, ay )) / fabsf( 4.0f * dot( wo, m ) );
	if (pdf < 1.0e-6f) return;
	const float D = coat ? gtrD( m, ax ) : ggxD( m, ax, ay ), G = coat ? gtrG( wi, wo, ax ) : ggxG( wi, wo, ax, ay );
	value = (coat ? fresnelCoat( p, wo, m ) : fresnelSpec( p, wo, m )) * (D * G);
}
/* disney.h:113-137 */
static inline float evalDiffuse( const Principled& p, V3 n, V3 wo, V3 wi, V3 m, V3& value )
{
	const float con = dot( n, wo ), cin = dot( n, wi ), cih = dot( wi, m ), fl = schlick( cin ), fv = schlick( con );
	float fd = 0;
	if (p.subsurface != 1.0f) { const float fd90 = 0.5f + 2.0f * sqr( cih ) * p.roughness; fd = mixc( 1.0f, fd90, fl ) * mixc( 1.0f, fd90, fv ); }
	if (p.subsurface > 0)
	{
		const float fss90 = sqr( cih ) * p.roughness, fss = mixc( 1.0f, fss90, fl ) * mixc( 1.0f, fss90, fv );
		fd = mixc( fd, 1.25f * (fss * (1.0f / (fabsf( con ) + fabsf( cin )) - 0.5f) + 0.5f), p.subsurface );
	}
	value = p.color * fd * INVPI_ * (1.0f - p.metallic);
	return fabsf( cin ) * INVPI_;
}
/* disney.h:139-149 */
static inline float evalSheen( const Principled& p, V3 wi, V3 m, V3& value )
{
	value = (v3( 1.0f - p.sheenTint ) + p.sheenTint * p.tint) * (schlick( dot( wi, m ) ) * p.sheen * (1.0f - p.metallic));
	return 1.0f / (2 * PI_);
}

/* frosted.h:20-36 */
static inline float fresnelDielectric( float ci, float eta, float& ct )
{
	const float s2 = (1 - sqr( ci )) * sqr( eta );
	if (s2 > 1) { ct = 0; return 1; }
	ct = fminf( sqrtf( fmaxf( 1 - s2, 0.0f ) ), 1.0f );
	const float a = fabsf( ci );
	if (a == 0 && ct == 0) return 1;
	const float k0 = eta * ct, k1 = eta * a;
	return 0.5f * (sqr( (a - k0) / (a + k0) ) + sqr( (ct - k1) / (ct + k1) ));
}
/* frosted.h:66-74 */
static inline V3 glassReflect( V3 color, V3 wo, V3 wi, V3 m, float ax, float ay, float F )
{
	const float den = fabsf( 4 * wo.z * wi.z );
	return den == 0 ? v3( 0 ) : color * (F * ggxD( m, ax, ay ) * ggxG( wi, wo, ax, ay ) / den);
}
/* frosted.h:80-93 (adjoint = false) */
static inline V3 glassRefract( float eta, V3 color, V3 wo, V3 wi, V3 m, float ax, float ay, float T )
{
	if (wo.z == 0 || wi.z == 0) return v3( 0 );
	const float cih = dot( m, wi ), coh = dot( m, wo ), dots = (cih * coh) / (wi.z * wo.z), sd = coh + eta * cih;
	if (fabsf( sd ) < 1.0e-6f) return v3( 0 );
	float mul = fabsf( dots ) * T * ggxD( m, ax, ay ) * ggxG( wi, wo, ax, ay ) / sqr( sd );
	mul *= sqr( eta );
	return color * mul;
}
static inline float jacReflect( float coh ) { return coh == 0 ? 0 : 1 / (4 * fabsf( coh )); }	// frosted.h:94-98
static inline float jacRefract( V3 wo, V3 wi, V3 m, float eta )	// frosted.h:99-105
{
	const float cih = dot( m, wi ), coh = dot( m, wo ), sd = coh + eta * cih;
	return fabsf( sd ) < 1.0e-6f ? 0 : fabsf( cih ) * sqr( eta / sd );
}
static inline float chooseReflect( float F ) { const float s = F + (1 - F); return s != 0 ? F / s : 1; }	// frosted.h:53-59 with weights 1, 1
static inline void weights( const Principled& p, float w[4] )	// disney.h:213-215,321-322
{
	w[0] = lerpf( p.lum, 0, p.metallic ), w[1] = lerpf( p.sheen, 0, p.metallic ), w[2] = lerpf( p.specular, 1, p.metallic ), w[3] = p.clearcoat * 0.25f;
	const float r = 1.0f / (w[0] + w[1] + w[2] + w[3]);
	for (int i = 0; i < 4; i++) w[i] *= r;
}
} // namespace pr

/* disney.h:290-358 */
static inline V3 EvaluatePrincipled( const Principled& p, V3 iN, V3 iT, V3 wow, V3 wiw, float& pdf )
{
	using namespace pr;
	pdf = 0;
	const V3 B = normalize( pr::cross( iN, iT ) ), T = normalize( pr::cross( iN, B ) );
	if (p.transmission > 0.5f)
	{
		const V3 wo = local( wow, iN, T, B ), wi = local( wiw, iN, T, B );
		const float eta = wo.z > 0 ? p.eta : (1.0f / p.eta);
		if (eta == 1) return v3( 0 );
		float ax, ay, jac, ct;
		alphas( p.roughness, p.anisotropic, ax, ay );
		V3 value, m;
		if (wi.z * wo.z >= 0)
		{
			m = normalize( wi + wo );
			if (m.z < 0) m = m * -1.0f;
			const float c = dot( wo, m ), F = fresnelDielectric( c, 1 / eta, ct );
			value = glassReflect( p.color, wo, wi, m, ax, ay, F ), pdf = chooseReflect( F ), jac = jacReflect( c );
		}
		else
		{
			m = normalize( wo + eta * wi );
			if (m.z < 0) m = m * -1.0f;
			const float c = dot( wo, m ), F = fresnelDielectric( c, 1 / eta, ct );
			value = glassRefract( eta, p.color, wo, wi, m, ax, ay, 1 - F ), pdf = 1 - chooseReflect( F ), jac = jacRefract( wo, wi, m, eta );
		}
		pdf *= jac * ggxPdf( wo, m, ax, ay );
		return value;
	}
	if (p.roughness <= 0.001f) return v3( 0 );
	float w[4];
	weights( p, w );
	V3 value = v3( 0 );
	if (w[0] + w[1] > 0)
	{
		const V3 m = normalize( wiw + wow );
		if (w[0] > 0) pdf += w[0] * evalDiffuse( p, iN, wow, wiw, m, value );
		if (w[1] > 0) pdf += w[1] * evalSheen( p, wiw, m, value );	// the reference writes the same variable: sheen replaces diffuse
	}
	if (w[2] + w[3] > 0)
	{
		const V3 wo = local( wow, iN, T, B ), wi = local( wiw, iN, T, B ), m = normalize( wo + wi );
		if (w[2] > 0)
		{
			float ax, ay;
			alphas( p.roughness, p.anisotropic, ax, ay );
			V3 c = v3( 0 );
			const float sp = evalMf( p, false, ax, ay, wo, wi, m, c );
			if (sp > 0) pdf += w[2] * sp, value = value + c;
		}
		if (w[3] > 0)
		{
			const float a = coatAlpha( p );
			V3 c = v3( 0 );
			const float cp = evalMf( p, true, a, a, wo, wi, m, c );
			if (cp > 0) pdf += w[3] * cp, value = value + c;
		}
	}
	return value;
}

/* disney.h:151-288 */
static inline V3 SamplePrincipled( const Principled& p, V3 iN, V3 N, V3 iT, V3 wow, float distance, float r0, float r1, float r2, V3& wiw, float& pdf, bool& specular )
{
	using namespace pr;
	pdf = 0;
	const float flip = (dot( wow, N ) < 0) ? -1.0f : 1.0f;
	iN = iN * flip;
	const V3 B = normalize( pr::cross( iN, iT ) ), T = normalize( pr::cross( iN, B ) );
	if (r0 < p.transmission)
	{
		specular = true;
		const float r3 = r0 / p.transmission;
		const V3 wo = local( wow, iN, T, B );
		const float eta = flip < 0 ? (1 / p.eta) : p.eta;
		if (eta == 1) return v3( 0 );
		const V3 beer = v3( expf( -p.transmittance.x * distance * 2.0f ), expf( -p.transmittance.y * distance * 2.0f ), expf( -p.transmittance.z * distance * 2.0f ) );
		float ax, ay, ct, jac;
		alphas( p.roughness, p.anisotropic, ax, ay );
		const V3 m = ggxSample( wo, r1, r3, ax, ay );
		const float rcp = 1 / eta, c = fmaxf( -1.0f, fminf( dot( wo, m ), 1.0f ) ), F = fresnelDielectric( c, eta, ct );
		V3 wi, value;
		if (r2 < F)
		{
			wi = reflect( wo * -1.0f, m );
			if (wi.z * wo.z <= 0) return v3( 0 );
			value = glassReflect( p.color, wo, wi, m, ax, ay, F ), pdf = F, jac = jacReflect( c );
		}
		else
		{
			// frosted.h:46-52 is called with eta where it expects 1 / eta (disney.h:196); restated as called
			const V3 d = c > 0 ? (eta * c - ct) * m - eta * wo : (eta * c + ct) * m - eta * wo;
			wi = d * ((3 - dot( d, d )) * 0.5f);
			if (wi.z * wo.z > 0) return v3( 0 );
			value = glassRefract( rcp, p.color, wo, wi, m, ax, ay, 1 - F ), pdf = 1 - F, jac = jacRefract( wo, wi, m, rcp );
		}
		pdf *= jac * ggxPdf( wo, m, ax, ay );
		if (pdf > 1.0e-6f) wiw = world( wi, iN, T, B );
		return value * beer;
	}
	const float r3 = (r0 - p.transmission) / (1 - p.transmission);
	float w[4];
	weights( p, w );
	const float cdf0 = w[0], cdf1 = w[0] + w[1], cdf2 = w[0] + w[1] + w[2];
	float prob, cp;
	V3 value = v3( 0 ), c = v3( 0 );
	if (r3 < cdf1)
	{
		const float ra = r3 / cdf1, t1 = TWOPI_ * ra, t2 = sqrtf( 1 - r1 );	// common_functions.h:118-124
		wiw = (cosf( t1 ) * t2 * T) + (sinf( t1 ) * t2) * B + sqrtf( r1 ) * iN;
		const V3 m = normalize( wiw + wow );
		if (r3 < cdf0) cp = evalDiffuse( p, iN, wow, wiw, m, value ), prob = w[0] * cp, w[0] = 0;
		else cp = evalSheen( p, wiw, m, value ), prob = w[1] * cp, w[1] = 0;
	}
	else
	{
		const V3 wo = local( wow, iN, T, B );
		V3 wi;
		if (r3 < cdf2)
		{
			float ax, ay;
			alphas( p.roughness, p.anisotropic, ax, ay );
			sampleMf( p, false, (r3 - cdf1) / (cdf2 - cdf1), r1, ax, ay, wo, wi, cp, value );
			prob = w[2] * cp, w[2] = 0;
		}
		else
		{
			const float a = coatAlpha( p );
			sampleMf( p, true, (r3 - cdf2) / (1 - cdf2), r1, a, a, wo, wi, cp, value );
			prob = w[3] * cp, w[3] = 0;
		}
		value = value * (1.0f / fabsf( 4.0f * wo.z * wi.z ));
		wiw = world( wi, iN, T, B );
	}
	if (w[0] + w[1] > 0)
	{
		const V3 m = normalize( wiw + wow );
		if (w[0] > 0) prob += w[0] * evalDiffuse( p, iN, wow, wiw, m, c ), value = value + c;
		if (w[1] > 0) prob += w[1] * evalSheen( p, wiw, m, c ), value = value + c;
	}
	if (w[2] + w[3] > 0)
	{
		const V3 wo = local( wow, iN, T, B ), wi = local( wiw, iN, T, B ), m = normalize( wo + wi );
		if (w[2] > 0)
		{
			float ax, ay;
			alphas( p.roughness, p.anisotropic, ax, ay );
			c = v3( 0 );
			prob += w[2] * evalMf( p, false, ax, ay, wo, wi, m, c ), value = value + c;
		}
		if (w[3] > 0)
		{
			const float a = coatAlpha( p );
			c = v3( 0 );
			prob += w[3] * evalMf( p, true, a, a, wo, wi, m, c ), value = value + c;
		}
	}
	pdf = prob > 1.0e-6f ? prob : 0;
	return value;
}

} // namespace orc
