/* lh2_oracle_capi.cpp - TEST INFRASTRUCTURE ONLY. C entry points (ctypes) of the CPU oracle.
   See lh2_oracle_geom.h / lh2_oracle_shade.h for what is restated and from where. */
#include "lh2_oracle_geom.h"
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
