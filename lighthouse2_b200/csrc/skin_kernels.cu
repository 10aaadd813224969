/* skin_kernels.cu - mesh animation on the device: the vertex work the reference does on the host
   before every SetGeometry of an animated mesh, moved in front of the refit so that no triangle data crosses PCIe.

   Restates lib/RenderSystem/host_mesh.cpp
     HostMesh::SetPose( const HostSkin* )        :748-906  linear-blend skinning: M = sum_k w_k * jointMat[j_k] per vertex,
                                                           position = M * (p, 1), vertex normal = normalize( M3x3 * n ),
                                                           geometric normal from the skinned corners; written to the
                                                           vertex list, CoreTri::vertex0..2, vN0..2 and Nx/Ny/Nz
     HostMesh::SetPose( const vector<float>& )   :711-741  morph targets: p = base + sum_j w_j * delta_j; vertex normals =
                                                           normalize( base normal + sum_j normal_j ) (the reference adds the
                                                           target normals UNWEIGHTED: kept); Nx/Ny/Nz are not touched there
   One thread per triangle; the outputs are the two per-mesh buffers SetGeometry fills (float4 positions for the BVH,
   CoreTri records for shading). Area / tangent / LOD of the records stay as uploaded, like in the reference.
*/
#include "kernels.h"
#include "common.cuh"

namespace lh2b
{

__device__ __forceinline__ float3 Norm3( const float3 v ) { const float l = 1.0f / sqrtf( dot( v, v ) ); return v * l; }

__global__ void captureBindPoseKernel( const float4* __restrict__ coreTris, float4* __restrict__ bindNormals, const int triCount )
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= triCount) return;
	for (int k = 0; k < 3; k++) bindNormals[3 * t + k] = coreTris[(size_t)t * 13 + 2 + k];
}

__global__ void __launch_bounds__( 128 ) skinKernel( const float4* __restrict__ bindVerts, const float4* __restrict__ bindNormals,
	const uint4* __restrict__ joints, const float4* __restrict__ weights, const float4* __restrict__ jointMats /* 4 rows per joint */,
	const int jointCount, float4* __restrict__ verts, float4* __restrict__ coreTris, const int triCount )
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= triCount) return;
	float3 P[3], Nv[3];
	for (int k = 0; k < 3; k++)
	{
		const int v = 3 * t + k;
		const uint4 j = joints[v];
		const float4 w = weights[v];
		float4 r0 = make_float4( 0, 0, 0, 0 ), r1 = r0, r2 = r0;
		const uint32_t ji[4] = { j.x, j.y, j.z, j.w };
		const float wi[4] = { w.x, w.y, w.z, w.w };
		for (int q = 0; q < 4; q++)
		{
			const uint32_t jq = min( ji[q], (uint32_t)(jointCount - 1) );
			const float4 a = __ldg( jointMats + jq * 4 ), b = __ldg( jointMats + jq * 4 + 1 ), c = __ldg( jointMats + jq * 4 + 2 );
			r0.x += wi[q] * a.x, r0.y += wi[q] * a.y, r0.z += wi[q] * a.z, r0.w += wi[q] * a.w;
			r1.x += wi[q] * b.x, r1.y += wi[q] * b.y, r1.z += wi[q] * b.z, r1.w += wi[q] * b.w;
			r2.x += wi[q] * c.x, r2.y += wi[q] * c.y, r2.z += wi[q] * c.z, r2.w += wi[q] * c.w;
		}
		const float4 p = bindVerts[v], n = bindNormals[v];
		// positions are points: w = 1 (HostMesh stores make_float4( pos, 1 ), host_mesh.cpp:570-572), whatever the caller left in w
		P[k] = make_float3( r0.x * p.x + r0.y * p.y + r0.z * p.z + r0.w, r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w, r2.x * p.x + r2.y * p.y + r2.z * p.z + r2.w );
		Nv[k] = Norm3( make_float3( r0.x * n.x + r0.y * n.y + r0.z * n.z, r1.x * n.x + r1.y * n.y + r1.z * n.z, r2.x * n.x + r2.y * n.y + r2.z * n.z ) );
	}
	const float3 N = Norm3( cross( P[1] - P[0], P[2] - P[0] ) );
	const float Nk[3] = { N.x, N.y, N.z };
	float4* rec = coreTris + (size_t)t * 13;
	for (int k = 0; k < 3; k++)
	{
		verts[3 * t + k] = make_float4( P[k].x, P[k].y, P[k].z, 1 );
		rec[2 + k] = make_float4( Nv[k].x, Nv[k].y, Nv[k].z, Nk[k] );
		const float keep = rec[8 + k].w;
		rec[8 + k] = make_float4( P[k].x, P[k].y, P[k].z, keep );
	}
}

__global__ void __launch_bounds__( 128 ) morphKernel( const float4* __restrict__ bindVerts, const float4* __restrict__ bindNormals,
	const float4* __restrict__ targetDeltas /* [target][vertex] */, const float4* __restrict__ targetNormals, const float* __restrict__ targetWeights,
	const int targetCount, float4* __restrict__ verts, float4* __restrict__ coreTris, const int triCount )
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= triCount) return;
	const size_t vertexCount = (size_t)triCount * 3;
	float4* rec = coreTris + (size_t)t * 13;
	for (int k = 0; k < 3; k++)
	{
		const int v = 3 * t + k;
		float4 p = bindVerts[v], n = bindNormals[v];
		p.w = 1;
		for (int j = 0; j < targetCount; j++)
		{
			const float wj = targetWeights[j];
			const float4 d = targetDeltas[j * vertexCount + v], dn = targetNormals[j * vertexCount + v];
			p.x += wj * d.x, p.y += wj * d.y, p.z += wj * d.z;
			n.x += dn.x, n.y += dn.y, n.z += dn.z;	// unweighted, as host_mesh.cpp:731-734
		}
		const float3 nn = Norm3( make_float3( n.x, n.y, n.z ) );
		verts[v] = p;
		rec[2 + k] = make_float4( nn.x, nn.y, nn.z, rec[2 + k].w );
		rec[8 + k] = make_float4( p.x, p.y, p.z, rec[8 + k].w );
	}
}

void LaunchCaptureBindPose( const float4* coreTris, float4* bindNormals, int triCount, cudaStream_t s )
{
	if (triCount > 0) captureBindPoseKernel<<<(triCount + 255) / 256, 256, 0, s>>>( coreTris, bindNormals, triCount );
}
void LaunchSkin( const float4* bindVerts, const float4* bindNormals, const uint4* joints, const float4* weights, const float4* jointMats, int jointCount,
	float4* verts, float4* coreTris, int triCount, cudaStream_t s )
{
	if (triCount > 0) skinKernel<<<(triCount + 127) / 128, 128, 0, s>>>( bindVerts, bindNormals, joints, weights, jointMats, jointCount, verts, coreTris, triCount );
}
void LaunchMorph( const float4* bindVerts, const float4* bindNormals, const float4* deltas, const float4* normals, const float* weights, int targetCount,
	float4* verts, float4* coreTris, int triCount, cudaStream_t s )
{
	if (triCount > 0) morphKernel<<<(triCount + 127) / 128, 128, 0, s>>>( bindVerts, bindNormals, deltas, normals, weights, targetCount, verts, coreTris, triCount );
}

} // namespace lh2b
