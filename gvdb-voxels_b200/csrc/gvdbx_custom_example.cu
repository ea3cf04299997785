// gvdbx_custom_example.cu — the reference's custom-kernel sample (source/gRenderKernel/render_custom.cu:33-63: trilinear
// surface hit, then a "custom look" of diffuse + sky reflection) written against gvdbx_plugin.cuh.  Built into
// libgvdbx_custom_example.so; tests/test_parity_gpu.py checks it bit for bit against the reference's own sample kernel
// launched through VolumeGVDB::RenderKernel.
#include "../../include/gvdbx.h"
#include "gvdbx_plugin.cuh"

__device__ __forceinline__ float3 reflect3(float3 i, float3 n) { return i - 2.0f * n * gx_dot(n, i); }

__global__ void gvdbx_example_raycast_kernel(const __grid_constant__ GxParams P)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.width || y >= P.height) return;

    float3 hit = make_float3(GX_NOHIT, GX_NOHIT, GX_NOHIT);
    float4 clr = make_float4(1, 1, 1, 1);
    float3 norm = make_float3(0, 0, 0);
    float3 rdir = gx_normalize(gvdbx_view_ray(P, (float(x) + 0.5) / P.width, (float(y) + 0.5) / P.height));

    // the sample starts its rays at scn.campos (no grid transform)
    gvdbx_ray_cast<GX_MODE_TRILINEAR>(P, P.campos, rdir, hit, norm, clr, x, y);

    if (hit.z != GX_NOHIT) {
        float3 lightdir = gx_normalize(P.light_pos - hit);
        float3 eyedir = gx_normalize(P.campos - hit);
        float3 R = gx_normalize(reflect3(eyedir, norm));
        float diffuse = max(0.0f, gx_dot(norm, lightdir));
        float refl = min(1.0f, max(0.0f, R.y));
        const float d = diffuse * 0.6;
        clr = make_float4(d + refl * 0.f, d + refl * 0.3f, d + refl * 0.7f, d + refl * 1.0f);
    } else {
        clr = make_float4(0.0, 0.0, 0.1, 1.0);
    }
    P.out[y * P.width + x] = make_uchar4(clr.x * 255, clr.y * 255, clr.z * 255, 255);
}

// host side: VolumeGVDB::RenderKernel(user_kernel, chan, rbuf) — 8x8 CTAs like the reference's launch (:4325)
extern "C" int gvdbx_example_render_custom(gvdbx_t* h, const void* scninfo, int chan, uint64_t outbuf_d, void* cuda_stream)
{
    GxParams P;
    int rc = gvdbx_kernel_params(h, scninfo, GVDBX_SHADE_TRILINEAR, chan, outbuf_d, &P, sizeof P);
    if (rc) return rc;
    dim3 block(8, 8, 1), grid((P.width + 7) / 8, (P.height + 7) / 8, 1);
    gvdbx_example_raycast_kernel<<<grid, block, GVDBX_KERNEL_SMEM(64), (cudaStream_t)cuda_stream>>>(P);
    return cudaGetLastError() == cudaSuccess ? GVDBX_OK : GVDBX_E_CUDA;
}
