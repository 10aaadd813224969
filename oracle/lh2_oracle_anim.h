/* lh2_oracle_anim.h - TEST INFRASTRUCTURE ONLY (CPU oracle). Never included, linked or called by the product.

   CPU restatement of the reference's host-side mesh animation, the step in front of SetGeometry for animated meshes:
     lib/RenderSystem/host_mesh.cpp:748-906  HostMesh::SetPose( const HostSkin* )  (linear-blend skinning; the scalar branch
                                             :884-904 states the arithmetic, the AVX branch :762-882 is what the stock build runs:
                                             both position AND normal are multiplied as M * v, normals normalised, the
                                             geometric normal comes from the skinned corners)
     lib/RenderSystem/host_mesh.cpp:711-741  HostMesh::SetPose( const vector<float>& )  (morph targets; target normals are
                                             added unweighted; Nx/Ny/Nz untouched)
   Data: verts = float4[3 * triCount]; tris = CoreTri[triCount] as 52 floats each (float4 #2..4 = vN0..2 + Nx/Ny/Nz,
   #8..10 = vertex0..2). Parity unpinned by reference execution (RenderSystem does not build here: FreeImage, GL); pinned to
   the source lines above. The reference normalises with _mm_rsqrt_ps (12-bit estimate); this restatement uses exact
   square roots, so comparisons against a real reference run would need 4e-4 relative on normals.
*/
#pragma once
#include <math.h>
#include <stdint.h>

namespace orc
{

static inline void Normalize3( float* v ) { const float l = 1.0f / sqrtf( v[0] * v[0] + v[1] * v[1] + v[2] * v[2] ); v[0] *= l, v[1] *= l, v[2] *= l; }

static inline void SkinMesh( const float* bindVerts, const float* bindNormals /* float4 per vertex */, const uint32_t* joints, const float* weights,
	const float* jointMats /* 16 per joint, row major */, int triCount, float* verts, float* tris )
{
	for (int t = 0; t < triCount; t++)
	{
		float P[3][3];
		for (int k = 0; k < 3; k++)
		{
			const int v = 3 * t + k;
			float M[12] = { 0 };
			for (int q = 0; q < 4; q++)
			{
				const float* J = jointMats + (size_t)joints[v * 4 + q] * 16;
				for (int c = 0; c < 12; c++) M[c] += weights[v * 4 + q] * J[c];
			}
			const float* p = bindVerts + v * 4, * n = bindNormals + v * 4;
			float N[3];
			for (int r = 0; r < 3; r++)
			{
				P[k][r] = M[r * 4] * p[0] + M[r * 4 + 1] * p[1] + M[r * 4 + 2] * p[2] + M[r * 4 + 3];	// w = 1: HostMesh vertices are make_float4( pos, 1 )
				N[r] = M[r * 4] * n[0] + M[r * 4 + 1] * n[1] + M[r * 4 + 2] * n[2];
			}
			Normalize3( N );
			float* out = verts + v * 4;
			out[0] = P[k][0], out[1] = P[k][1], out[2] = P[k][2], out[3] = 1;
			float* rec = tris + (size_t)t * 52;
			rec[(2 + k) * 4] = N[0], rec[(2 + k) * 4 + 1] = N[1], rec[(2 + k) * 4 + 2] = N[2];
			rec[(8 + k) * 4] = P[k][0], rec[(8 + k) * 4 + 1] = P[k][1], rec[(8 + k) * 4 + 2] = P[k][2];
		}
		const float a[3] = { P[1][0] - P[0][0], P[1][1] - P[0][1], P[1][2] - P[0][2] }, b[3] = { P[2][0] - P[0][0], P[2][1] - P[0][1], P[2][2] - P[0][2] };
		float G[3] = { a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0] };
		Normalize3( G );
		float* rec = tris + (size_t)t * 52;
		rec[2 * 4 + 3] = G[0], rec[3 * 4 + 3] = G[1], rec[4 * 4 + 3] = G[2];
	}
}

static inline void MorphMesh( const float* bindVerts, const float* bindNormals, const float* deltas /* [target][vertex] float4 */, const float* normals,
	const float* w, int targetCount, int triCount, float* verts, float* tris )
{
	const size_t vc = (size_t)triCount * 3;
	for (size_t v = 0; v < vc; v++)
	{
		float p[3] = { bindVerts[v * 4], bindVerts[v * 4 + 1], bindVerts[v * 4 + 2] }, n[3] = { bindNormals[v * 4], bindNormals[v * 4 + 1], bindNormals[v * 4 + 2] };
		for (int j = 0; j < targetCount; j++) for (int c = 0; c < 3; c++)
			p[c] += w[j] * deltas[(j * vc + v) * 4 + c], n[c] += normals[(j * vc + v) * 4 + c];
		Normalize3( n );
		verts[v * 4] = p[0], verts[v * 4 + 1] = p[1], verts[v * 4 + 2] = p[2], verts[v * 4 + 3] = 1;
		float* rec = tris + (v / 3) * 52;
		const int k = (int)(v % 3);
		for (int c = 0; c < 3; c++) rec[(2 + k) * 4 + c] = n[c], rec[(8 + k) * 4 + c] = p[c];
	}
}

} // namespace orc
