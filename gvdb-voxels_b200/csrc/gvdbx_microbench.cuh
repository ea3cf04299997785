// gvdbx_microbench.cuh — fetch-only microbenchmarks behind gvdbx_measure_sampler_ab: how many fp32 trilinear samples per second
// each way of reading a brick delivers on THIS GPU for the access pattern of a ray packet marching through bricks (north_star:
// "brick atlas re-laid out for coalesced, vectorised loads and staged through shared memory or TMA per coherent ray tile, the
// layout and the texture-versus-linear-load choice justified by counters").  All variants take the same samples: every warp is an
// 8x4 packet `spacing` voxels apart, 32 samples per lane and brick, a handful of bricks per warp visited in turn.
//   variant 0  texture unit on the caller's 3-D array (the production sampler; gx_tex_peak_kernel in gvdbx_import.cuh)
//   variant 1  brick-major copy (4 KB contiguous per 10^3 brick), 8 scalar read-only loads + the software model of the unit's
//              1.8 fixed-point filter (GxSampler<GX_SAMPLER_LINEAR>, the production texture-free sampler)
//   variant 2  x-pair layout (float2 {T[x], T[x+1]} per texel, 8 KB per brick): 4 vector loads of 8 bytes + the same filter
//   variant 3  the brick-major block of the warp's current brick staged into shared memory by TMA (cp.async.bulk, 4 KB per copy,
//              mbarrier completion, two stages per warp so that the copy of the next brick overlaps the samples of this one),
//              8 shared-memory loads + the same filter
// The filter arithmetic (~45 integer / float instructions per sample) is identical in 1-3; what differs is where the texels come from.
#pragma once
#include "gvdbx_trace.cuh"

// weights of the unit's filter for brick-local texel coordinates (see GxSampler<GX_SAMPLER_LINEAR>::tri)
struct GxSoftW { int ix, iy, iz, ix1, iy1, iz1; float w[8]; };
__device__ __forceinline__ void gx_soft_split(float c, int& i, int& a)
{
    const float cb = c - 0.5f, f = floorf(cb);
    a = __float2int_rd(fmaf(cb - f, 256.0f, 0.5f));
    i = int(f);
    if (a >= 256) { a = 0; i++; }
}
__device__ __forceinline__ GxSoftW gx_soft_weights(float x, float y, float z)
{
    GxSoftW s;
    int ax, ay, az;
    gx_soft_split(x, s.ix, ax); gx_soft_split(y, s.iy, ay); gx_soft_split(z, s.iz, az);
    const int BD = GX_BRICK_DIM;
    s.ix1 = min(max(s.ix + 1, 0), BD - 1); s.iy1 = min(max(s.iy + 1, 0), BD - 1); s.iz1 = min(max(s.iz + 1, 0), BD - 1);
    s.ix = min(max(s.ix, 0), BD - 1); s.iy = min(max(s.iy, 0), BD - 1); s.iz = min(max(s.iz, 0), BD - 1);
    const int by = 256 - ay, s0 = 256 - az, s1 = az;
    const int x1a = (s0 * ax + 128) >> 8, x0a = s0 - x1a, x1b = (s1 * ax + 128) >> 8, x0b = s1 - x1b;
    const int w110 = (x1a * ay + 128) >> 8, w100 = x1a - w110, w000 = (x0a * by + 128) >> 8, w010 = x0a - w000;
    const int w111 = (x1b * ay + 128) >> 8, w101 = x1b - w111, w001 = (x0b * by + 128) >> 8, w011 = x0b - w001;
    s.w[0] = float(w000); s.w[1] = float(w100); s.w[2] = float(w010); s.w[3] = float(w110);
    s.w[4] = float(w001); s.w[5] = float(w101); s.w[6] = float(w011); s.w[7] = float(w111);
    return s;
}
__device__ __forceinline__ float gx_soft_sum(const GxSoftW& s, float c000, float c100, float c010, float c110, float c001, float c101, float c011, float c111)
{
    float acc = s.w[0] * c000;
    acc = fmaf(s.w[1], c100, acc); acc = fmaf(s.w[2], c010, acc); acc = fmaf(s.w[3], c110, acc);
    acc = fmaf(s.w[4], c001, acc); acc = fmaf(s.w[5], c101, acc); acc = fmaf(s.w[6], c011, acc); acc = fmaf(s.w[7], c111, acc);
    return acc * (1.0f / 256.0f);
}

// x-pair layout of the first `n` leaf blocks of the brick-major copy
__global__ void gx_build_pairs(const float* __restrict__ bricks, float2* __restrict__ pairs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 1000) return;
    const int b = i / 1000, t = i % 1000, x = t % 10;
    const float* src = bricks + size_t(b) * GX_BRICK_STRIDE;
    pairs[size_t(b) * GX_BRICK_STRIDE + t] = make_float2(src[t], src[x < 9 ? t + 1 : t]);
}

#define GX_MB_SAMPLES 32        // samples per lane and brick visit
#define GX_MB_BRICKS  4         // bricks a warp cycles through

__device__ __forceinline__ unsigned gx_smem_addr(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }

template <int VARIANT>
__global__ void __launch_bounds__(128) gx_linear_peak_kernel(const float* __restrict__ bricks, const float2* __restrict__ pairs, int nbricks, int visits,
                                                              float spacing, float* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char gx_mb_smem[];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gwarp = tid >> 5, lane = threadIdx.x & 31, lwarp = threadIdx.x >> 5;
    const float x0 = 1.0f + spacing * float(lane & 7), y0 = 1.0f + 1.07f * spacing * float(lane >> 3), z0 = 1.0f;
    const float dx = 0.11f, dy = 0.07f, dz = 0.22f;
    float acc = 0.f;

    // variant 3: two 4 KB stages + two mbarriers per warp
    float* stage = reinterpret_cast<float*>(gx_mb_smem) + lwarp * 2 * GX_BRICK_STRIDE;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(gx_mb_smem + (blockDim.x >> 5) * 2 * GX_BRICK_STRIDE * sizeof(float)) + lwarp * 2;
    auto brick_of = [&](int v) { return (gwarp * GX_MB_BRICKS + (v % GX_MB_BRICKS)) % nbricks; };
    auto issue = [&](int v) {           // one lane arms the stage's barrier with the byte count and starts the bulk copy
        const int s = v & 1;
        const unsigned b = gx_smem_addr(bar + s), d = gx_smem_addr(stage + s * GX_BRICK_STRIDE);
        const float* src = bricks + size_t(brick_of(v)) * GX_BRICK_STRIDE;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(GX_BRICK_STRIDE * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(d), "l"(src), "r"(GX_BRICK_STRIDE * 4), "r"(b) : "memory");
    };
    if (VARIANT == 3) {
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gx_smem_addr(bar)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gx_smem_addr(bar + 1)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(0);
        }
        __syncwarp();
    }

    for (int v = 0; v < visits; v++) {
        const float* blk = bricks + size_t(brick_of(v)) * GX_BRICK_STRIDE;
        const float2* pblk = pairs + size_t(brick_of(v)) * GX_BRICK_STRIDE;
        const float* sblk = stage + (v & 1) * GX_BRICK_STRIDE;
        if (VARIANT == 3) {
            if (lane == 0 && v + 1 < visits) issue(v + 1);          // the other stage was released by the __syncwarp below
            const unsigned b = gx_smem_addr(bar + (v & 1)), parity = (v >> 1) & 1;
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(b), "r"(parity) : "memory");
        }
        float x = x0, y = y0, z = z0;
        #pragma unroll 4
        for (int k = 0; k < GX_MB_SAMPLES; k++) {
            const GxSoftW s = gx_soft_weights(x, y, z);
            const int BD = GX_BRICK_DIM;
            const int r00 = (s.iz * BD + s.iy) * BD, r10 = (s.iz * BD + s.iy1) * BD, r01 = (s.iz1 * BD + s.iy) * BD, r11 = (s.iz1 * BD + s.iy1) * BD;
            if (VARIANT == 1) {
                acc += gx_soft_sum(s, __ldg(blk + r00 + s.ix), __ldg(blk + r00 + s.ix1), __ldg(blk + r10 + s.ix), __ldg(blk + r10 + s.ix1),
                                   __ldg(blk + r01 + s.ix), __ldg(blk + r01 + s.ix1), __ldg(blk + r11 + s.ix), __ldg(blk + r11 + s.ix1));
            } else if (VARIANT == 2) {
                const float2 a = __ldg(pblk + r00 + s.ix), b = __ldg(pblk + r10 + s.ix), c = __ldg(pblk + r01 + s.ix), d = __ldg(pblk + r11 + s.ix);
                acc += gx_soft_sum(s, a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y);
            } else {
                acc += gx_soft_sum(s, sblk[r00 + s.ix], sblk[r00 + s.ix1], sblk[r10 + s.ix], sblk[r10 + s.ix1],
                                   sblk[r01 + s.ix], sblk[r01 + s.ix1], sblk[r11 + s.ix], sblk[r11 + s.ix1]);
            }
            x += dx; y += dy; z += dz;
        }
        if (VARIANT == 3) __syncwarp();     // every lane has finished reading this stage before it is refilled
    }
    out[tid] = acc;
}

// ------------------------------------------------------------------------------------------------ deep sample loop alone
// Ceiling of the deep marcher's INNER loop on this GPU: the four-sample round of gx_raycast_deep_q — four texture fetches, four
// transfer-index computations, four 16-byte table gathers, four emission / absorption updates with their separately rounded
// products — and nothing else: no traversal, no brick entries, no partial rounds, every lane always active, bricks L1-resident,
// the same 8x4 packet geometry as gx_tex_peak_kernel.  deep achieved / this = what traversal, brick changes, lane imbalance and
// cache misses cost; this / texture peak = what the per-sample arithmetic and the table gather cost.
// LUTMODE 0: the table gather is a 16-byte read-only load (the production path); 1: a float4 point fetch through a linear
// texture object over the same table (A/B: does the texture path gather 16-byte entries cheaper than the LSU path?)
template <int LUTMODE>
__global__ void __launch_bounds__(128, 7) gx_deep_loop_peak_kernel(cudaTextureObject_t tex, const float4* __restrict__ lut, cudaTextureObject_t lut_tex, int ares_x, int ares_y, int ares_z,
                                                                   int rounds, float spacing, float thresh, float inv_range, float minval, float albedo,
                                                                   float* __restrict__ out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const float bx = float((warp * 10) % max(ares_x - 10, 1)) + 1.0f, by = float(((warp / 7) * 10) % max(ares_y - 10, 1)) + 1.0f,
                bz = float(((warp / 53) * 10) % max(ares_z - 10, 1)) + 1.0f;
    float3 p = make_float3(bx + spacing * float(lane & 7), by + 1.07f * spacing * float(lane >> 3), bz);
    const float3 wpt = make_float3(0.11f, 0.07f, 0.22f);
    GxParams P;                                  // only the field gx_deep_accumulate_pre reads
    P.extinct.y = albedo;
    float4 clr = make_float4(0.f, 0.f, 0.f, 1.f);
    for (int r = 0; r < rounds; r++) {
        float3 p1, p2, p3;
        GX_STEP_ADD(p1, p); GX_STEP_ADD(p2, p1); GX_STEP_ADD(p3, p2);
        const float v0 = tex3D<float>(tex, p.x, p.y, p.z), v1 = tex3D<float>(tex, p1.x, p1.y, p1.z);
        const float v2 = tex3D<float>(tex, p2.x, p2.y, p2.z), v3 = tex3D<float>(tex, p3.x, p3.y, p3.z);
        float4 c0, c1, c2, c3;
        if (LUTMODE == 0) {
            c0 = gx_lut(lut, gx_transfer_index(v0, thresh, inv_range)); c1 = gx_lut(lut, gx_transfer_index(v1, thresh, inv_range));
            c2 = gx_lut(lut, gx_transfer_index(v2, thresh, inv_range)); c3 = gx_lut(lut, gx_transfer_index(v3, thresh, inv_range));
        } else {
            c0 = tex1Dfetch<float4>(lut_tex, int(gx_transfer_index(v0, thresh, inv_range))); c1 = tex1Dfetch<float4>(lut_tex, int(gx_transfer_index(v1, thresh, inv_range)));
            c2 = tex1Dfetch<float4>(lut_tex, int(gx_transfer_index(v2, thresh, inv_range))); c3 = tex1Dfetch<float4>(lut_tex, int(gx_transfer_index(v3, thresh, inv_range)));
        }
        if (v0 >= minval) gx_deep_accumulate_pre(P, clr, c0);
        if (v1 >= minval) gx_deep_accumulate_pre(P, clr, c1);
        if (v2 >= minval) gx_deep_accumulate_pre(P, clr, c2);
        if (v3 >= minval) gx_deep_accumulate_pre(P, clr, c3);
        if (!(clr.w > 0.01f)) clr.w = 1.0f;      // the marcher would stop here; the benchmark keeps the lane busy
        GX_STEP_ADD(p, p3);
        if (p.z > bz + 7.5f) p = make_float3(bx + spacing * float(lane & 7), by + 1.07f * spacing * float(lane >> 3), bz);
    }
    out[tid] = clr.x + clr.y + clr.z + clr.w;
}
