/* common.cuh - small host/device utilities of the B200 core: error plumbing, device buffers,
   float3/float4 helpers. Behavioural counterpart of the reference's CoreBuffer<T> / CHK_CUDA
   (lib/CUDA/shared_host_code/cudatools.h:23-28,180-330), minus host mirrors and GL interop. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace lh2b
{

void SetLastError( const std::string& msg );

struct CoreError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CUDA_CHECK( call ) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
	char buf_[512]; snprintf( buf_, sizeof( buf_ ), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString( e_ ) ); \
	throw lh2b::CoreError( buf_ ); } } while (0)

template <typename T> struct DevBuf
{
	T* ptr = nullptr;
	size_t count = 0, capacity = 0;
	DevBuf() = default;
	DevBuf( const DevBuf& ) = delete;
	DevBuf& operator=( const DevBuf& ) = delete;
	~DevBuf() { if (ptr) cudaFree( ptr ); }
	void Reserve( size_t n, bool keep = false )
	{
		if (n <= capacity) return;
		T* np = nullptr;
		CUDA_CHECK( cudaMalloc( &np, n * sizeof( T ) ) );
		if (keep && ptr && count) CUDA_CHECK( cudaMemcpy( np, ptr, count * sizeof( T ), cudaMemcpyDeviceToDevice ) );
		if (ptr) cudaFree( ptr );
		ptr = np, capacity = n;
	}
	void Resize( size_t n ) { Reserve( n ); count = n; }
	void Upload( const T* src, size_t n, cudaStream_t s )
	{
		Resize( n );
		if (n) CUDA_CHECK( cudaMemcpyAsync( ptr, src, n * sizeof( T ), cudaMemcpyHostToDevice, s ) );
	}
	void Free() { if (ptr) cudaFree( ptr ); ptr = nullptr, count = capacity = 0; }
	void Swap( DevBuf& o ) { T* p = ptr; ptr = o.ptr, o.ptr = p; size_t t = count; count = o.count, o.count = t; t = capacity; capacity = o.capacity, o.capacity = t; }
	size_t Bytes() const { return count * sizeof( T ); }
};

__host__ __device__ __forceinline__ float3 operator+( const float3 a, const float3 b ) { return make_float3( a.x + b.x, a.y + b.y, a.z + b.z ); }
__host__ __device__ __forceinline__ float3 operator-( const float3 a, const float3 b ) { return make_float3( a.x - b.x, a.y - b.y, a.z - b.z ); }
__host__ __device__ __forceinline__ float3 operator-( const float3 a ) { return make_float3( -a.x, -a.y, -a.z ); }
__host__ __device__ __forceinline__ float3 operator*( const float3 a, const float3 b ) { return make_float3( a.x * b.x, a.y * b.y, a.z * b.z ); }
__host__ __device__ __forceinline__ float3 operator*( const float3 a, const float s ) { return make_float3( a.x * s, a.y * s, a.z * s ); }
__host__ __device__ __forceinline__ float3 operator*( const float s, const float3 a ) { return make_float3( a.x * s, a.y * s, a.z * s ); }
__host__ __device__ __forceinline__ void operator+=( float3& a, const float3 b ) { a.x += b.x, a.y += b.y, a.z += b.z; }
__host__ __device__ __forceinline__ void operator-=( float3& a, const float3 b ) { a.x -= b.x, a.y -= b.y, a.z -= b.z; }
__host__ __device__ __forceinline__ void operator*=( float3& a, const float s ) { a.x *= s, a.y *= s, a.z *= s; }
__host__ __device__ __forceinline__ void operator*=( float3& a, const float3 b ) { a.x *= b.x, a.y *= b.y, a.z *= b.z; }
__host__ __device__ __forceinline__ float dot( const float3 a, const float3 b ) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ float3 cross( const float3 a, const float3 b ) { return make_float3( a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x ); }
__host__ __device__ __forceinline__ float3 xyz( const float4 a ) { return make_float3( a.x, a.y, a.z ); }
__host__ __device__ __forceinline__ float4 f4( const float3 a, const float w ) { return make_float4( a.x, a.y, a.z, w ); }
__host__ __device__ __forceinline__ float3 f3( const float s ) { return make_float3( s, s, s ); }

} // namespace lh2b
