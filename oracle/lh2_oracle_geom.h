/* lh2_oracle_geom.h - TEST INFRASTRUCTURE ONLY. CPU oracle for the ray/scene queries.

   Parity status: UNPINNED BY REFERENCE SOURCE for traversal. The reference delegates ray/triangle
   intersection to NVIDIA OptiX 7.4.0 (ABI 55; headers only under lib/OptiX7, runtime = display
   driver, absent here) - call sites lib/rendercore_optix7/optix/.optix.cu:125,136,148 - and ships
   no test, golden vector or fixture for it. The oracle therefore restates the *published semantic
   contract* of those call sites as an exhaustive brute-force search:
     - closest hit over all instances x triangles, t in (0, 1e34), no culling (.optix.cu:125,136);
     - ray moved to object space with the instance's inverse 3x4 (rendercore.cpp:405-416 builds the
       inverse; OptiX applies it), direction not renormalised so t is shared between spaces;
     - hit record (u16 | v16<<16, instanceIdx, primIdx, t), barycentrics truncated (.optix.cu:174-184);
     - occlusion query = any hit with t in (0, tmax) (.optix.cu:147-149).
   Equal-t candidates are resolved toward the smaller (instance, primitive) pair; OptiX leaves
   that case undefined, so the CUDA core adopts the same rule and parity is bit-exact everywhere.

   Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
   this code. The product (lighthouse2_b200/) never includes, links or calls it.
*/
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace orc
{

struct Mesh { const float* verts4; int triCount; };			// float4[3 * triCount]
struct Instance { int mesh; float xform[12]; };				// row-major 3x4 object->world

/* float helpers with a fixed operation order; build with -ffp-contract=off */
static inline float Dot3( float ax, float ay, float az, float bx, float by, float bz )
{
	return fmaf( ax, bx, fmaf( ay, by, az * bz ) );
}
static inline float CrossX( float ax, float ay, float az, float bx, float by, float bz ) { (void)ax, (void)bx; return fmaf( ay, bz, -(az * by) ); }
static inline float CrossY( float ax, float ay, float az, float bx, float by, float bz ) { (void)ay, (void)by; return fmaf( az, bx, -(ax * bz) ); }
static inline float CrossZ( float ax, float ay, float az, float bx, float by, float bz ) { (void)az, (void)bz; return fmaf( ax, by, -(ay * bx) ); }

/* affine inverse, same cofactor order as the core's host code (csrc/core.cu InvertAffine) */
static inline void InvertAffine( const float* m, float* inv )
{
	const float a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
	const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
	const float det = a * A + b * B + c * C;
	const float id = det != 0 ? 1.0f / det : 0.0f;
	inv[0] = A * id, inv[1] = -(b * i - c * h) * id, inv[2] = (b * f - c * e) * id;
	inv[4] = B * id, inv[5] = (a * i - c * g) * id, inv[6] = -(a * f - c * d) * id;
	inv[8] = C * id, inv[9] = -(a * h - b * g) * id, inv[10] = (a * e - b * d) * id;
	inv[3] = -(inv[0] * m[3] + inv[1] * m[7] + inv[2] * m[11]);
	inv[7] = -(inv[4] * m[3] + inv[5] * m[7] + inv[6] * m[11]);
	inv[11] = -(inv[8] * m[3] + inv[9] * m[7] + inv[10] * m[11]);
}

static inline bool IsIdentity( const float* m )
{
	static const float id[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
	return memcmp( m, id, sizeof( id ) ) == 0;
}

static inline void ToObjectSpace( const float* inv, const float* O, const float* D, float* oO, float* oD )
{
	for (int r = 0; r < 3; r++)
	{
		const float* row = inv + r * 4;
		oO[r] = fmaf( row[0], O[0], fmaf( row[1], O[1], fmaf( row[2], O[2], row[3] ) ) );
		oD[r] = fmaf( row[0], D[0], fmaf( row[1], D[1], row[2] * D[2] ) );
	}
}

/* Moeller-Trumbore, two-sided. Returns true and (t,u,v) if the supporting plane is crossed inside the triangle. */
static inline bool TriTest( const float* O, const float* D, const float* v0, const float* v1, const float* v2, float& t, float& u, float& v )
{
	const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
	const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
	const float pvx = CrossX( D[0], D[1], D[2], e2x, e2y, e2z );
	const float pvy = CrossY( D[0], D[1], D[2], e2x, e2y, e2z );
	const float pvz = CrossZ( D[0], D[1], D[2], e2x, e2y, e2z );
	const float det = Dot3( e1x, e1y, e1z, pvx, pvy, pvz );
	if (!(det != 0.0f)) return false;
	const float inv = 1.0f / det;
	const float tvx = O[0] - v0[0], tvy = O[1] - v0[1], tvz = O[2] - v0[2];
	u = Dot3( tvx, tvy, tvz, pvx, pvy, pvz ) * inv;
	if (!(u >= 0.0f && u <= 1.0f)) return false;
	const float qvx = CrossX( tvx, tvy, tvz, e1x, e1y, e1z );
	const float qvy = CrossY( tvx, tvy, tvz, e1x, e1y, e1z );
	const float qvz = CrossZ( tvx, tvy, tvz, e1x, e1y, e1z );
	v = Dot3( D[0], D[1], D[2], qvx, qvy, qvz ) * inv;
	if (!(v >= 0.0f && u + v <= 1.0f)) return false;
	t = Dot3( e2x, e2y, e2z, qvx, qvy, qvz ) * inv;
	return true;
}

struct Hit { float t, u, v; int inst, prim; };

/* Optional pruning structure (lh2_oracle_bvh.h). AccelMode() = 0 (default): exhaustive search, the definition.
   1: the same search pruned by a per-mesh BVH; identical results (tests/test_oracle_cpu.py), used for full-size checks
   and as the CPU baseline. */
struct Accel;
static inline int& AccelMode() { static int mode = 0; return mode; }
struct Scene;
static inline Accel* BuildSceneAccel( const Scene& s );
static inline void FreeSceneAccel( Accel* a );

struct Scene
{
	const Mesh* meshes; int meshCount;
	const Instance* instances; int instanceCount;
	float* inverses;	// 12 floats per instance, filled by Prepare
	Accel* accel = nullptr;
	void Prepare()
	{
		inverses = new float[(size_t)(instanceCount > 0 ? instanceCount : 1) * 12];
		for (int i = 0; i < instanceCount; i++) InvertAffine( instances[i].xform, inverses + i * 12 );
		accel = AccelMode() ? BuildSceneAccel( *this ) : nullptr;
	}
	void Release() { delete[] inverses; inverses = 0; if (accel) FreeSceneAccel( accel ); accel = nullptr; }
};

static inline bool AccelClosestHit( const Scene& s, const float* O, const float* D, const float tmin, float tmax, Hit& best );
static inline bool AccelOccluded( const Scene& s, const float* O, const float* D, const float tmin, const float tmax );

static inline bool ClosestHit( const Scene& s, const float* O, const float* D, const float tmin, float tmax, Hit& best )
{
	if (s.accel) return AccelClosestHit( s, O, D, tmin, tmax, best );
	best.t = tmax, best.inst = -1, best.prim = -1, best.u = best.v = 0;
	for (int i = 0; i < s.instanceCount; i++)
	{
		const Mesh& m = s.meshes[s.instances[i].mesh];
		float oO[3], oD[3];
		ToObjectSpace( s.inverses + i * 12, O, D, oO, oD );
		for (int p = 0; p < m.triCount; p++)
		{
			const float* v = m.verts4 + (size_t)p * 12;
			float t, u, w;
			if (!TriTest( oO, oD, v, v + 4, v + 8, t, u, w )) continue;
			if (!(t > tmin)) continue;
			// (inst, prim) ascends in this loop, so a strict '<' keeps the smallest pair on equal t
			if (t < best.t) best.t = t, best.u = u, best.v = w, best.inst = i, best.prim = p;
		}
	}
	return best.prim >= 0;
}

static inline bool Occluded( const Scene& s, const float* O, const float* D, const float tmin, const float tmax )
{
	if (s.accel) return AccelOccluded( s, O, D, tmin, tmax );
	for (int i = 0; i < s.instanceCount; i++)
	{
		const Mesh& m = s.meshes[s.instances[i].mesh];
		float oO[3], oD[3];
		ToObjectSpace( s.inverses + i * 12, O, D, oO, oD );
		for (int p = 0; p < m.triCount; p++)
		{
			const float* v = m.verts4 + (size_t)p * 12;
			float t, u, w;
			if (TriTest( oO, oD, v, v + 4, v + 8, t, u, w ) && t > tmin && t < tmax) return true;
		}
	}
	return false;
}

static inline void PackHit( const bool hit, const Hit& h, uint32_t* out4 )
{
	if (!hit)
	{
		const float t = 1e34f;
		out4[0] = 0, out4[1] = 0, out4[2] = 0xffffffffu;
		memcpy( out4 + 3, &t, 4 );
		return;
	}
	out4[0] = (uint32_t)(65535.0f * h.u) + ((uint32_t)(65535.0f * h.v) << 16);
	out4[1] = (uint32_t)h.inst, out4[2] = (uint32_t)h.prim;
	memcpy( out4 + 3, &h.t, 4 );
}

} // namespace orc
