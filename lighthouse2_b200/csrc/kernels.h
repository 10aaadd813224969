/* kernels.h - host-callable launchers of the CUDA stages (one .cu per stage family). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lh2b
{
struct DevScene;

void LaunchExtend( const DevScene& scene, const float4* O4, const float4* D4, float4* hits, int n, cudaStream_t s );
void LaunchOcclude( const DevScene& scene, const float4* O4, const float4* D4, uint8_t* occluded, int n, cudaStream_t s );

} // namespace lh2b
