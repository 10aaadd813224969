/* lh2_oracle_capi.cpp - TEST INFRASTRUCTURE ONLY. C entry points (ctypes) of the CPU oracle.
   See lh2_oracle_geom.h / lh2_oracle_shade.h for what is restated and from where. */
#include "lh2_oracle_geom.h"
#include "lh2_oracle_bvh.h"
#include <thread>
#include <vector>
#include <algorithm>

#define ORC_API extern "C" __attribute__( ( visibility( "default" ) ) )

template <typename F> static void ParallelFor( int n, int threads, F f )
{
	if (threads <= 1 || n < 64) { f( 0, n ); return; }
	std::vector<std::thread> pool;
	const int chunk = (n + threads - 1) / threads;
	for (int t = 0; t < threads; t++)
	{
		const int a = t * chunk, b = std::min( n, a + chunk );
		if (a >= b) break;
		pool.emplace_back( [=]() { f( a, b ); } );
	}
	for (auto& th : pool) th.join();
}

/* 0: exhaustive search (the definition, default); 1: the same search pruned by a per-mesh BVH (lh2_oracle_bvh.h). */
ORC_API void orc_set_accel( int mode ) { orc::AccelMode() = mode; }
ORC_API int orc_get_accel() { return orc::AccelMode(); }

/* hits4: uint32[4*n] = (u16|v16<<16, inst, prim, t bits); prim = 0xffffffff on a miss. */
ORC_API void orc_closest_hits( const orc::Mesh* meshes, int meshCount, const orc::Instance* instances, int instanceCount,
	const float* O4, const float* D4, int n, uint32_t* hits4, int threads )
{
	orc::Scene s{ meshes, meshCount, instances, instanceCount, nullptr };
	s.Prepare();
	ParallelFor( n, threads, [&]( int a, int b ) {
		for (int i = a; i < b; i++)
		{
			orc::Hit h;
			const bool hit = orc::ClosestHit( s, O4 + i * 4, D4 + i * 4, 0.0f, 1e34f, h );
			orc::PackHit( hit, h, hits4 + i * 4 );
		}
	} );
	s.Release();
}

/* D4.w = tmax; occluded[i] = 1 if any triangle is hit with t in (0, tmax). */
ORC_API void orc_occluded( const orc::Mesh* meshes, int meshCount, const orc::Instance* instances, int instanceCount,
	const float* O4, const float* D4, int n, uint8_t* occluded, int threads )
{
	orc::Scene s{ meshes, meshCount, instances, instanceCount, nullptr };
	s.Prepare();
	ParallelFor( n, threads, [&]( int a, int b ) {
		for (int i = a; i < b; i++) occluded[i] = orc::Occluded( s, O4 + i * 4, D4 + i * 4, 0.0f, D4[i * 4 + 3] ) ? 1 : 0;
	} );
	s.Release();
}

/* ---- full frame ---------------------------------------------------------------------------- */
#include "lh2_oracle_render.h"
#include "lh2_oracle_anim.h"

struct OrcMaterialIn
{
	float color[3], absorption[3];
	float params[12];		// metallic, subsurface, specular, roughness, specularTint, anisotropic, sheen, sheenTint, clearcoat, clearcoatGloss, transmission, eta
	uint32_t flags;			// CoreMaterial::flags (bit 0 smooth normals, bit 1 alpha)
	int32_t tex[6];			// texture ids: color, detailColor, normals, detailNormals, specular, roughness (-1 = none)
	float uvscale[6][2], uvoffset[6][2];
};

struct OrcFrameIn
{
	const orc::Mesh* meshes; const float* const* coreTris; int meshCount;
	const orc::Instance* instances; int instanceCount;
	const OrcMaterialIn* materials; int materialCount;
	const orc::TexDesc* textures; int textureCount;
	const float* triLights; int triLightCount;
	const float* pointLights; int pointLightCount;
	const float* spotLights; int spotLightCount;
	const float* dirLights; int dirLightCount;
	const float* skyPixels; int skyW, skyH;		// float3 pixels
	float worldToSky[16];
	const uint8_t* blueNoiseBytes;				// sob | scr | rnk byte tables (327680 bytes)
	int w, h, spp, pass;
	uint32_t sampleBase;
	uint32_t shiftSeed, camRNGseed;				// generator states BEFORE this frame (rendercore.h:122-123)
	float geometryEpsilon, clampValue;
	int maxPathLength; uint32_t enoughBounces;
	float view[17];
	int threads;
	int bsdfModel;								// 0 lambert.h, 1 disney.h
};

static uint32_t XorShift( uint32_t& s ) { s ^= s << 13, s ^= s >> 17, s ^= s << 5; return s; }

static void ConvertMaterials( const OrcFrameIn& in, std::vector<orc::TexDesc>& tex, orc::RenderScene& sc )
{
	// texture packing per storage class (rendercore.cpp:458-502), then material conversion (rendercore.cpp:508-549)
	tex.assign( in.textures, in.textures + in.textureCount );
	for (int storage = 0; storage < 3; storage++)
	{
		size_t at = 0;
		for (auto& t : tex) if (t.storage == storage)
		{
			t.firstPixel = (uint32_t)at;
			if (storage == 1) sc.argb128.insert( sc.argb128.end(), (const float*)t.data, (const float*)t.data + (size_t)t.pixelCount * 4 );
			else (storage == 0 ? sc.argb32 : sc.nrm32).insert( (storage == 0 ? sc.argb32 : sc.nrm32).end(), (const uint8_t*)t.data, (const uint8_t*)t.data + (size_t)t.pixelCount * 4 );
			at += t.pixelCount;
		}
	}
	sc.argb32.resize( sc.argb32.size() + 64, 0 ), sc.nrm32.resize( sc.nrm32.size() + 64, 0 ), sc.argb128.resize( sc.argb128.size() + 64, 0 );
	auto toChar = []( float a ) { return (uint32_t)orc::F2U( a * 255.0f ); };
	for (int i = 0; i < in.materialCount; i++)
	{
		const OrcMaterialIn& m = in.materials[i];
		orc::Material g;
		memset( &g, 0, sizeof( g ) );
		uint16_t th[3];
		for (int k = 0; k < 3; k++)
		{
			g.color[k] = orc::HalfToFloat( orc::FloatToHalf( m.color[k] ) );
			th[k] = orc::FloatToHalf( 1 - m.absorption[k] );
			g.transmittance[k] = orc::HalfToFloat( th[k] );
		}
		g.baseZ = (uint32_t)th[1] | ((uint32_t)th[2] << 16);
		const float* p = m.params;
		g.params[0] = toChar( p[0] ) + (toChar( p[1] ) << 8) + (toChar( p[2] ) << 16) + (toChar( p[3] ) << 24);
		g.params[1] = toChar( p[4] ) + (toChar( p[5] ) << 8) + (toChar( p[6] ) << 16) + (toChar( p[7] ) << 24);
		g.params[2] = toChar( p[8] ) + (toChar( p[9] ) << 8) + (toChar( p[10] ) << 16);
		g.params[3] = orc::FBits( p[11] );
		const int bit[6] = { 2, 9, 3, 7, 4, 5 };
		orc::Material::Map* maps[6] = { &g.tex0, &g.tex1, &g.nmap0, &g.nmap1, &g.smap, &g.rmap };
		g.flags = (p[11] < 1 ? 1u : 0) + ((m.flags & 1) ? (1u << 11) : 0) + ((m.flags & 2) ? (1u << 12) : 0);
		for (int k = 0; k < 6; k++) if (m.tex[k] != -1)
		{
			g.flags += 1u << bit[k];
			const orc::TexDesc& t = tex[m.tex[k]];
			maps[k]->w = (int16_t)t.width, maps[k]->h = (int16_t)t.height;
			maps[k]->uscale = orc::HalfToFloat( orc::FloatToHalf( m.uvscale[k][0] ) ), maps[k]->vscale = orc::HalfToFloat( orc::FloatToHalf( m.uvscale[k][1] ) );
			maps[k]->uoffs = orc::HalfToFloat( orc::FloatToHalf( m.uvoffset[k][0] ) ), maps[k]->voffs = orc::HalfToFloat( orc::FloatToHalf( m.uvoffset[k][1] ) );
			maps[k]->addr = t.firstPixel;
		}
		if (m.tex[0] != -1 && (tex[m.tex[0]].flags & 8)) g.flags += 1u << 1;
		sc.materials.push_back( g );
	}
}

/* Renders one frame exactly as RenderCore::Render(view, converge) + FinalizeRender would; accumulates into
   accum (float4 per pixel; caller zeroes it for a Restart frame). rayCounts[0] = extension rays, [1] = shadow rays.
   seedsOut[0] = shiftSeed after, seedsOut[1] = camRNGseed after. records (optional): one PathRecord per path. */
struct FrameCtx { orc::RenderScene sc; std::vector<orc::TexDesc> tex; std::vector<uint32_t> bn; };

static void SetupScene( const OrcFrameIn& in, FrameCtx& ctx )
{
	orc::RenderScene& sc = ctx.sc;
	std::vector<orc::TexDesc>& tex = ctx.tex;
	std::vector<uint32_t>& bn = ctx.bn;
	sc.geo = orc::Scene{ in.meshes, in.meshCount, in.instances, in.instanceCount, nullptr };
	sc.geo.Prepare();
	sc.coreTris = in.coreTris;
	ConvertMaterials( in, tex, sc );
	sc.triLights = in.triLights, sc.triLightCount = in.triLightCount;
	sc.pointLights = in.pointLights, sc.pointLightCount = in.pointLightCount;
	sc.spotLights = in.spotLights, sc.spotLightCount = in.spotLightCount;
	sc.dirLights = in.dirLights, sc.dirLightCount = in.dirLightCount;
	// sky: float3 -> float4 plus the 64x64 box-filtered copy (rendercore.cpp:716-737)
	sc.skyW = in.skyW, sc.skyH = in.skyH;
	const int sw = in.skyW >> 6, sh = in.skyH >> 6;
	sc.sky.assign( ((size_t)in.skyW * in.skyH + (size_t)sw * sh + 1) * 4, 0.0f );
	for (size_t i = 0; i < (size_t)in.skyW * in.skyH; i++) for (int k = 0; k < 3; k++) sc.sky[i * 4 + k] = in.skyPixels[i * 3 + k];
	for (int y = 0; y < sh; y++) for (int x = 0; x < sw; x++)
	{
		float total[4] = { 0, 0, 0, 0 };
		for (int v = 0; v < 64; v++) for (int u = 0; u < 64; u++)
		{
			const float* t = sc.sky.data() + 4 * ((size_t)(x * 64 + u) + (size_t)(y * 64 + v) * in.skyW);
			for (int k = 0; k < 4; k++) total[k] += t[k];
		}
		float* o = sc.sky.data() + 4 * ((size_t)in.skyW * in.skyH + x + (size_t)y * sw);
		for (int k = 0; k < 4; k++) o[k] = total[k] * (1.0f / (64 * 64));
	}
	memcpy( sc.worldToSky, in.worldToSky, sizeof( sc.worldToSky ) );
	bn.assign( 65536 * 5 + 16, 0 );	// 16 zero words behind the ranking tile: the sampler reads past it for tile pixel (127, 127) at dimensions >= 8 (as the reference does)
	for (int i = 0; i < 65536; i++) bn[i] = in.blueNoiseBytes[i];
	for (int i = 0; i < 128 * 128 * 8; i++) bn[i + 65536] = in.blueNoiseBytes[65536 + i], bn[i + 3 * 65536] = in.blueNoiseBytes[65536 + 131072 + i];
	sc.blueNoise = bn.data();
}

/* filter (optional): features uint4[w*h], worldPos / deltaDepth float4[w*h] (in/out: history bits persist); when given, accum is float4[2*w*h]. */
ORC_API void orc_render_frame( const OrcFrameIn* inp, float* accum, uint64_t* rayCounts, uint32_t* seedsOut, orc::PathRecord* records,
	uint32_t* features, float* worldPos, float* deltaDepth )
{
	const OrcFrameIn& in = *inp;
	FrameCtx ctx;
	SetupScene( in, ctx );
	orc::RenderScene& sc = ctx.sc;
	orc::Settings st;
	memset( &st, 0, sizeof( st ) );
	st.w = in.w, st.h = in.h, st.spp = in.spp, st.pass = in.pass, st.sampleBase = in.sampleBase;
	uint32_t shiftSeed = in.shiftSeed, camSeed = in.camRNGseed;
	XorShift( shiftSeed );
	st.shift = shiftSeed;
	for (int L = 1; L <= in.maxPathLength; L++) st.R0[L] = XorShift( camSeed ) + L * 91771;
	st.geometryEpsilon = in.geometryEpsilon, st.clampValue = in.clampValue;
	st.maxPathLength = in.maxPathLength, st.enoughBounces = in.enoughBounces, st.bsdfModel = in.bsdfModel;
	memcpy( st.view, in.view, sizeof( st.view ) );
	if (seedsOut) seedsOut[0] = shiftSeed, seedsOut[1] = camSeed;
	const int pixels = in.w * in.h;
	const orc::FilterArrays fa = { features, worldPos, deltaDepth };
	const bool filter = features != nullptr;
	std::vector<double> acc( (size_t)pixels * 4 * (filter ? 2 : 1), 0.0 );
	const int threads = in.threads > 0 ? in.threads : 1;
	std::vector<uint64_t> counts( (size_t)threads * 2, 0 );
	std::vector<std::thread> pool;
	// interleave rows over threads: cost varies smoothly over the image
	for (int t = 0; t < threads; t++) pool.emplace_back( [&, t]() {
		uint32_t rc[2] = { 0, 0 };
		for (int y = t; y < in.h; y += threads) for (int x = 0; x < in.w; x++) for (int s = 0; s < in.spp; s++)
		{
			const uint32_t pathIdx = (uint32_t)(x + y * in.w) + (uint32_t)s * pixels;
			orc::PathRecord* rec = records ? records + pathIdx : nullptr;
			if (rec) memset( rec, 0, sizeof( *rec ) );
			orc::TracePath( sc, st, pathIdx, acc.data(), rc, rec, filter ? &fa : nullptr );
			counts[t * 2] += rc[0], counts[t * 2 + 1] += rc[1], rc[0] = rc[1] = 0;
		}
	} );
	for (auto& th : pool) th.join();
	for (size_t i = 0; i < acc.size(); i++) accum[i] += (float)acc[i];
	if (rayCounts) { rayCounts[0] = rayCounts[1] = 0; for (int t = 0; t < threads; t++) rayCounts[0] += counts[t * 2], rayCounts[1] += counts[t * 2 + 1]; }
	sc.geo.Release();
}


/* CUDAMaterial records (128 bytes each, core_settings.h:136-150) as the reference host code would build them
   (rendercore.cpp:508-549); used to feed the reference shadeKernel harness. */
ORC_API void orc_convert_materials( const OrcFrameIn* inp, uint32_t* out32PerMaterial )
{
	FrameCtx ctx;
	SetupScene( *inp, ctx );
	for (size_t i = 0; i < ctx.sc.materials.size(); i++)
	{
		const orc::Material& g = ctx.sc.materials[i];
		uint32_t* o = out32PerMaterial + i * 32;
		memset( o, 0, 128 );
		o[0] = orc::FloatToHalf( g.color[0] ) | ((uint32_t)orc::FloatToHalf( g.color[1] ) << 16);
		o[1] = orc::FloatToHalf( g.color[2] ) | ((uint32_t)orc::FloatToHalf( g.transmittance[0] ) << 16);
		o[2] = g.baseZ, o[3] = g.flags;
		memcpy( o + 4, g.params, 16 );
		const orc::Material::Map* maps[6] = { &g.tex0, &g.tex1, &g.nmap0, &g.nmap1, &g.smap, &g.rmap };
		for (int k = 0; k < 6; k++)
		{
			uint32_t* m = o + 8 + k * 4;
			m[0] = ((uint32_t)maps[k]->w & 0xffff) | (((uint32_t)maps[k]->h & 0xffff) << 16);
			m[1] = orc::FloatToHalf( maps[k]->uscale ) | ((uint32_t)orc::FloatToHalf( maps[k]->vscale ) << 16);
			m[2] = orc::FloatToHalf( maps[k]->uoffs ) | ((uint32_t)orc::FloatToHalf( maps[k]->voffs ) << 16);
			m[3] = maps[k]->addr;
		}
	}
	ctx.sc.geo.Release();
}

/* The packed texel arrays and sky table exactly as the core keeps them, for the reference harness. */
ORC_API void orc_scene_tables( const OrcFrameIn* inp, uint8_t* argb32, uint32_t* argb32Count, float* argb128, uint32_t* argb128Count,
	uint8_t* nrm32, uint32_t* nrm32Count, float* sky4, uint32_t* skyCount, float* inverses16 )
{
	FrameCtx ctx;
	SetupScene( *inp, ctx );
	const orc::RenderScene& sc = ctx.sc;
	*argb32Count = (uint32_t)(sc.argb32.size() / 4), *argb128Count = (uint32_t)(sc.argb128.size() / 4), *nrm32Count = (uint32_t)(sc.nrm32.size() / 4);
	*skyCount = (uint32_t)(sc.sky.size() / 4);
	if (argb32) memcpy( argb32, sc.argb32.data(), sc.argb32.size() );
	if (argb128) memcpy( argb128, sc.argb128.data(), sc.argb128.size() * 4 );
	if (nrm32) memcpy( nrm32, sc.nrm32.data(), sc.nrm32.size() );
	if (sky4) memcpy( sky4, sc.sky.data(), sc.sky.size() * 4 );
	if (inverses16) for (int i = 0; i < inp->instanceCount; i++)
	{
		float* o = inverses16 + i * 16;
		memcpy( o, sc.geo.inverses + i * 12, 48 );
		o[12] = o[13] = o[14] = 0, o[15] = 1;
	}
	ctx.sc.geo.Release();
}

/* shadeKernel on n paths (the stage-level oracle). Inputs/outputs mirror lh2b_shade_paths, except that outputs are
   NOT compacted: slot i belongs to input path i, flags[i] bit0 = extension ray, bit1 = shadow ray, bit2 = deposit. */
ORC_API void orc_shade_paths( const OrcFrameIn* inp, int pathLength, int n, const float* O4, const float* D4, const float* T4, const uint32_t* hits,
	uint32_t R0, uint32_t shift, int pass, float* extO, float* extD, float* extT, float* shO, float* shD, float* shE, float* deposit4, uint8_t* flags )
{
	const OrcFrameIn& in = *inp;
	FrameCtx ctx;
	SetupScene( in, ctx );
	orc::Settings st;
	memset( &st, 0, sizeof( st ) );
	st.w = in.w, st.h = in.h, st.spp = in.spp, st.pass = pass, st.sampleBase = in.sampleBase, st.shift = shift;
	st.R0[pathLength] = R0;
	st.geometryEpsilon = in.geometryEpsilon, st.clampValue = in.clampValue;
	st.maxPathLength = in.maxPathLength, st.enoughBounces = in.enoughBounces, st.bsdfModel = in.bsdfModel;
	memcpy( st.view, in.view, sizeof( st.view ) );
	ParallelFor( n, in.threads, [&]( int a, int b ) {
		for (int i = a; i < b; i++)
		{
			orc::PathState ps;
			ps.O = orc::v3( O4[i * 4], O4[i * 4 + 1], O4[i * 4 + 2] ), ps.data = orc::FBits( O4[i * 4 + 3] );
			ps.D = orc::v3( D4[i * 4], D4[i * 4 + 1], D4[i * 4 + 2] ), ps.packedN = orc::FBits( D4[i * 4 + 3] );
			ps.T = orc::v3( T4[i * 4], T4[i * 4 + 1], T4[i * 4 + 2] ), ps.bsdfPdf = T4[i * 4 + 3];
			orc::ShadeOut so;
			orc::ShadeStep( ctx.sc, st, pathLength, ps, hits + i * 4, -1, so );
			flags[i] = (so.extend ? 1 : 0) | (so.shadow ? 2 : 0) | (so.deposit ? 4 : 0);
			float* o;
			o = extO + i * 4, o[0] = so.next.O.x, o[1] = so.next.O.y, o[2] = so.next.O.z, o[3] = orc::BitsF( so.next.data );
			o = extD + i * 4, o[0] = so.next.D.x, o[1] = so.next.D.y, o[2] = so.next.D.z, o[3] = orc::BitsF( so.next.packedN );
			o = extT + i * 4, o[0] = so.next.T.x, o[1] = so.next.T.y, o[2] = so.next.T.z, o[3] = so.next.bsdfPdf;
			o = shO + i * 4, o[0] = so.sO.x, o[1] = so.sO.y, o[2] = so.sO.z, o[3] = 0;
			o = shD + i * 4, o[0] = so.sD.x, o[1] = so.sD.y, o[2] = so.sD.z, o[3] = so.sTmax;
			o = shE + i * 4, o[0] = so.E.x, o[1] = so.E.y, o[2] = so.E.z, o[3] = orc::BitsF( so.pixelIdx );
			o = deposit4 + i * 4, o[0] = so.contribution.x, o[1] = so.contribution.y, o[2] = so.contribution.z, o[3] = orc::BitsF( so.pixelIdx );
		}
	} );
	ctx.sc.geo.Release();
}

/* mesh animation (lh2_oracle_anim.h) */
ORC_API void orc_skin_mesh( const float* bindVerts, const float* bindNormals, const uint32_t* joints, const float* weights, const float* jointMats,
	int triCount, float* verts, float* tris )
{
	orc::SkinMesh( bindVerts, bindNormals, joints, weights, jointMats, triCount, verts, tris );
}
ORC_API void orc_morph_mesh( const float* bindVerts, const float* bindNormals, const float* deltas, const float* normals, const float* w, int targetCount,
	int triCount, float* verts, float* tris )
{
	orc::MorphMesh( bindVerts, bindNormals, deltas, normals, w, targetCount, triCount, verts, tris );
}

/* ---- SVGF / TAA chain (lh2_oracle_filter.h) -------------------------------------------------------------------------- */
#include "lh2_oracle_filter.h"

/* same layout as RefFilterIO of oracle/ref_filter_gpu.cu (ctypes: binding.FilterIO), so one set of buffers drives the reference
   kernels, the product's parity hook and this restatement */
struct OrcFilterIO
{
	int w, h, samplesTaken, camIsStationary, taa;
	float directClamp, indirectClamp, j0, j1, prevj0, prevj1;
	float prevView[17];
	const float* accumulator; const uint32_t* features; const float* worldPos; const float* prevWorldPos; const float* deltaDepth;
	const float* prevMoments; const float* filteredIN; const float* prevPixels;
	uint32_t* featuresOut; float* shadingAfterPrepare; float* motion; float* moments;
	float* phase1; float* phase2; float* phase3; float* taaPixels; float* target;
};

ORC_API int orc_filter_chain( const OrcFilterIO* io )
{
	using namespace orcf;
	const int w = io->w, h = io->h;
	const size_t px = (size_t)w * h;
	std::vector<U4> feat( px );
	memcpy( feat.data(), io->features, px * 16 );
	std::vector<V4> shading( px ), moments( px ), fOUT( px ), fIN( px ), taa( px ), target( px, V4{ 0, 0, 0, 0 } );
	std::vector<V2> motion( px );
	memcpy( fIN.data(), io->filteredIN, px * 16 );
	Chain c;
	c.w = w, c.h = h, c.samplesTaken = io->samplesTaken, c.camIsStationary = io->camIsStationary, c.taa = io->taa;
	c.directClamp = io->directClamp, c.indirectClamp = io->indirectClamp, c.j0 = io->j0, c.j1 = io->j1, c.prevj0 = io->prevj0, c.prevj1 = io->prevj1;
	memcpy( c.prevView, io->prevView, sizeof( c.prevView ) );
	c.accumulator = (const V4*)io->accumulator, c.features = feat.data(), c.worldPos = (const V4*)io->worldPos, c.prevWorldPos = (const V4*)io->prevWorldPos;
	c.deltaDepth = (const V4*)io->deltaDepth, c.prevMoments = (const V4*)io->prevMoments;
	c.shading = shading.data(), c.motion = motion.data(), c.moments = moments.data();
	Prepare( c );
	memcpy( io->shadingAfterPrepare, shading.data(), px * 16 ), memcpy( io->motion, motion.data(), px * 8 );
	memcpy( io->moments, moments.data(), px * 16 ), memcpy( io->featuresOut, feat.data(), px * 16 );
	ApplyFilter( c, shading.data(), fIN.data(), fOUT.data(), 1, 0 );
	memcpy( io->phase1, fOUT.data(), px * 16 );
	ApplyFilter( c, fOUT.data(), nullptr, fIN.data(), 2, 0 );
	memcpy( io->phase2, fIN.data(), px * 16 );
	ApplyFilter( c, fIN.data(), nullptr, shading.data(), 3, 1 );
	memcpy( io->phase3, shading.data(), px * 16 );
	if (io->taa)
	{
		TaaPass( shading.data(), taa.data(), (const V4*)io->prevPixels, motion.data(), w, h );
		memcpy( io->taaPixels, taa.data(), px * 16 );
		UnsharpenTaa( taa.data(), target.data(), w, h );
	}
	else FinalizeNoTaa( shading.data(), target.data(), w, h );
	memcpy( io->target, target.data(), px * 16 );
	return 0;
}

/* ---- independent reader of the product's CWBVH format (lh2_oracle_cwbvh.h) ------------------------------------------------ */
#include "lh2_oracle_cwbvh.h"

ORC_API void orc_cwbvh_check( const uint8_t* nodes, int nNodes, const float* tris, int nTris, const float* verts4, int triCount, int* report8 )
{
	const orcw::Report r = orcw::Check( nodes, nNodes, tris, nTris, verts4, triCount );
	memcpy( report8, &r, sizeof( r ) );
}

ORC_API void orc_cwbvh_closest_hits( const uint8_t* nodes, const float* tris, const float* O4, const float* D4, int n, uint32_t* hits4, int threads )
{
	ParallelFor( n, threads, [&]( int a, int b ) {
		for (int i = a; i < b; i++)
		{
			orc::Hit h;
			const bool hit = orcw::ClosestHit( nodes, tris, O4 + i * 4, D4 + i * 4, h );
			orc::PackHit( hit, h, hits4 + i * 4 );
		}
	} );
}
