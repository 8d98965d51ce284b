// ssao.cuh -- Renderer::post_process_ssao_SIMD (renderer/renderer.cpp:1229-1434) as CUDA kernels; the per-pixel arithmetic is
// ssao_device.h's.  The G-buffers are written by the shade stage (kernels.cuh: shade_prepare) or by the hybrid rasterizer
// (raster.cuh); k_ssao_occlusion counts the occluded samples of every pixel, k_ssao_apply blurs the counts and darkens the
// image, BEFORE the SSAA resolve (Renderer::post_process, renderer.cpp:1118-1124).
#pragma once

#include "ssao_device.h"

namespace rtb {

__global__ void __launch_bounds__(256)
k_gbuffer_clear(float* z, V3* n, size_t count)                                       // clear_z_buffer / clear_normal_buffer, renderer.cpp:165-173
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        z[i] = INFINITY;
        n[i] = v3(0.0f, 0.0f, 0.0f);
    }
}

__global__ void __launch_bounds__(128)
k_ssao_occlusion(SsaoView f, int* ao)
{
    const size_t total = (size_t)f.rw * f.rh;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) ao[i] = ssao_count_pixel(f, i);
}

__global__ void __launch_bounds__(256)
k_ssao_apply(SsaoView f, const int* ao, uint32_t* image)
{
    const size_t total = (size_t)f.rw * f.rh;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) ssao_apply_pixel(f, ao, image, i);
}

} // namespace rtb
