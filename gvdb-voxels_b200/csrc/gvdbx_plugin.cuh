// gvdbx_plugin.cuh — device-side API for USER kernels on the B200 layout: the counterpart of the reference's
// RenderKernel plugin point (src/gvdb_volume_gvdb.cpp:4309-4333), where an application compiles its own kernel against
// the five kernel headers (source/gRenderKernel/render_custom.cu:17-23) and launches it through VolumeGVDB::RenderKernel.
//
// Here: include this header, write
//     __global__ void my_kernel(const __grid_constant__ GxParams P) { ... gvdbx_ray_cast<GX_MODE_TRILINEAR>(P, ...) ... }
// compile with nvcc (-gencode arch=compute_100a,code=sm_100a --use_fast_math, like the library), ask the library for
// the frame's parameter block with gvdbx_kernel_params() and launch with GVDBX_KERNEL_SMEM(threads) bytes of dynamic
// shared memory (the traversal stack).  csrc/gvdbx_custom_example.cu is the reference's own sample kernel done this way.
#pragma once
#include "gvdbx_extra.cuh"

#define GVDBX_KERNEL_SMEM(threads_per_block) (size_t(threads_per_block) * GX_STACK_BYTES_PER_THREAD)

// ray origin / direction of the camera ray through normalised image coordinates (u, v)      cuda_gvdb_geom.cuh:48-63
__device__ __forceinline__ float3 gvdbx_view_pos(const GxParams& P) { return gx_mmult(P.invxform, P.campos); }
__device__ __forceinline__ float3 gvdbx_view_ray(const GxParams& P, float u, float v)
{
    float3 vv = u * P.camu + v * P.camv + P.cams;
    return gx_normalize(gx_mmult(P.invxrot, vv));
}

// rayCast(gvdb, chan, pos, dir, hit, norm, clr, brickFunc) with brickFunc chosen by MODE (GX_MODE_VOXEL / TRILINEAR /
// LEVELSET / DEEP / TRICUBIC / EMPTYSKIP / SHADOW).  hit / norm / clr carry the reference's in-out semantics: set
// hit = (NOHIT, NOHIT, NOHIT), clr = (1,1,1,1) for the surface modes, hit = (0, 0, NOHIT), clr = (0,0,0,1) for deep.
// (px, py) = pixel of this thread, only used when a depth buffer is bound.
template <int MODE, int SAMPLER = GX_SAMPLER_TEX>
__device__ __forceinline__ void gvdbx_ray_cast(const GxParams& P, float3 pos, float3 dir, float3& hit, float3& norm, float4& clr,
                                               int px = 0, int py = 0)
{
    GxSampler<SAMPLER, false> smp(P);
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    GxHit h;
    h.hit = hit; h.norm = norm; h.clr = clr;
    h.t = 0; h.leaf = -1; h.vox = make_int3(0, 0, 0);
    h.cpos = make_float3(0, 0, 0);
    gx_raycast<MODE, true>(P, smp, pos, dir, h, cnt, px, py);
    if (MODE == GX_MODE_VOXEL || MODE == GX_MODE_TRILINEAR || MODE == GX_MODE_LEVELSET || MODE == GX_MODE_TRICUBIC) gx_hit_color(P, h);
    hit = h.hit; norm = h.norm; clr = h.clr;
}
