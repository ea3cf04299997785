// gvdbx_import.cuh — non-template kernels used by gvdbx_api.cu only: reference pools / atlas -> compact tables and the
// brick-major atlas, tile assembly, sampler calibration.  (Kept out of gvdbx_device.cuh, which every gvdbx_k_*.cu includes.)
#pragma once
#include "gvdbx_extra.cuh"

// ------------------------------------------------------------------------------------------------ import kernels
// pool-0 / pool-1 (reference layout) -> compact tables.  One thread per child cell.
//   child list entry = Elem(0, lev-1, ndx) = grp | lev << 8 | ndx << 16, or 0xFFFFFFFFFFFFFFFF (src/gvdb_allocator.h:59-62,
//   gvdb_volume_gvdb.cpp:3015-3023); node->mChildList = Elem(1, lev, ndx) or ID_UNDEFL.
__global__ void gx_build_child_table(const char* __restrict__ nodelist, int nodewid, int nodecnt,
                                     const char* __restrict__ childlist, int childwid, int cells,
                                     int* __restrict__ child_out, int4* __restrict__ npos_out)
{
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t total = size_t(nodecnt) * cells;
    if (i >= total) return;
    int n = int(i / cells), b = int(i % cells);
    const GxNode* node = reinterpret_cast<const GxNode*>(nodelist + size_t(n) * nodewid);
    uint64_t listid = node->mChildList;
    int c = -1;
    if (listid != GX_ID_UNDEFL) {
        uint64_t cndx = listid >> 16;
        const uint64_t* clist = reinterpret_cast<const uint64_t*>(childlist + cndx * size_t(childwid));
        c = int(clist[b] >> 16);
    }
    child_out[i] = c;
    if (b == 0) npos_out[n] = make_int4(node->mPos.x, node->mPos.y, node->mPos.z, 0);
}

__global__ void gx_build_leaf_table(const char* __restrict__ nodelist, int nodewid, int nodecnt, int brick_res,
                                    int apron, int cnt_x, int cnt_y, GxLeafRec* __restrict__ out)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nodecnt) return;
    const GxNode* node = reinterpret_cast<const GxNode*>(nodelist + size_t(n) * nodewid);
    GxLeafRec r;
    r.px = node->mPos.x; r.py = node->mPos.y; r.pz = node->mPos.z;
    r.vx = node->mValue.x; r.vy = node->mValue.y; r.vz = node->mValue.z;
    int sx = (r.vx - apron) / brick_res, sy = (r.vy - apron) / brick_res, sz = (r.vz - apron) / brick_res;
    r.base = (r.vx < 0) ? 0 : ((sz * cnt_y + sy) * cnt_x + sx) * GX_BRICK_STRIDE;
    r.pad = 0;
    out[n] = r;
}

// atlas (x-fastest linear image, as cuMemcpy3D array->linear delivers it) -> brick-major; one CTA per brick slot.
// Also reduces the slot's value range (NaN-ignoring min / max over the 10^3 texels).
__global__ void gx_repack_atlas(const float* __restrict__ lin, int rx, int ry, int rz, int cnt_x, int cnt_y,
                                float* __restrict__ bricks, GxRange* __restrict__ range)
{
    const int slot = blockIdx.x;
    const int sx = slot % cnt_x, sy = (slot / cnt_x) % cnt_y, sz = slot / (cnt_x * cnt_y);
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < GX_BRICK_STRIDE; i += blockDim.x) {
        float v = 0.f;
        if (i < GX_BRICK_DIM * GX_BRICK_DIM * GX_BRICK_DIM) {
            int x = i % GX_BRICK_DIM, y = (i / GX_BRICK_DIM) % GX_BRICK_DIM, z = i / (GX_BRICK_DIM * GX_BRICK_DIM);
            size_t ax = size_t(sx) * GX_BRICK_DIM + x, ay = size_t(sy) * GX_BRICK_DIM + y, az = size_t(sz) * GX_BRICK_DIM + z;
            v = lin[(az * ry + ay) * rx + ax];
            lo = fminf(lo, v); hi = fmaxf(hi, v);
        }
        bricks[size_t(slot) * GX_BRICK_STRIDE + i] = v;
    }
    __shared__ float slo[8], shi[8];
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); w++) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
        range[slot].lo = lo; range[slot].hi = hi;
    }
}

// value range per leaf = range of the brick slot the leaf's mValue points at (run when both topology and atlas are in)
__global__ void gx_leaf_ranges(const GxLeafRec* __restrict__ leaf, int nleaf, const GxRange* __restrict__ slot_range,
                               int nslots, GxRange* __restrict__ out)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nleaf) return;
    const int slot = leaf[n].base / GX_BRICK_STRIDE;
    GxRange r;
    r.lo = -INFINITY; r.hi = INFINITY;                     // unknown slot: never culled
    if (leaf[n].vx >= 0 && slot >= 0 && slot < nslots) r = slot_range[slot];
    out[n] = r;
}

// scatter gathered tile buffers [nranks][slots][ts*ts] back into a row-major frame
__global__ void gx_assemble_tiles(const uchar4* __restrict__ gathered, uchar4* __restrict__ frame, int width, int height,
                                  int ts, int tiles_x, int ntiles, int nranks, int slots)
{
    const int tile = blockIdx.y;
    if (tile >= ntiles) return;
    const int r = tile % nranks, k = tile / nranks;
    const uchar4* src = gathered + (size_t(r) * slots + k) * ts * ts;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ts * ts; i += gridDim.x * blockDim.x) {
        int lx = i % ts, ly = i / ts;
        int x = (tile % tiles_x) * ts + lx, y = (tile / tiles_x) * ts + ly;
        if (x < width && y < height) frame[size_t(y) * width + x] = src[i];
    }
}

// calibration: hardware filter vs software model at arbitrary atlas-space points
__global__ void gx_sample_points_kernel(GxParams P, const float* __restrict__ xyz, int n, int cnt_x, int cnt_y,
                                        float* __restrict__ out_tex, float* __restrict__ out_lin)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    out_tex[i] = tex3D<float>(P.tex, x, y, z);
    GxSampler<GX_SAMPLER_LINEAR, false> s(P);
    GxLeafRec L;
    int sx = int(x) / GX_BRICK_DIM, sy = int(y) / GX_BRICK_DIM, sz = int(z) / GX_BRICK_DIM;
    L.vx = sx * GX_BRICK_DIM + 1; L.vy = sy * GX_BRICK_DIM + 1; L.vz = sz * GX_BRICK_DIM + 1;
    L.base = ((sz * cnt_y + sy) * cnt_x + sx) * GX_BRICK_STRIDE;
    L.px = L.py = L.pz = L.pad = 0;
    s.enter(L);
    out_lin[i] = s.tri(x, y, z);
}

// ------------------------------------------------------------------------------------------------ deep-mode transfer table
// {rgb, exp(EXTINCT * alpha * DIRECTSTEP)} for every transfer-function entry: the expression of rayDeepBrick
// (cuda_gvdb_raycast.cuh:515) evaluated once per entry and frame instead of once per sample (gx_deep_accumulate_pre).
__global__ void gx_build_deep_lut(const float4* __restrict__ src, float4* __restrict__ dst, int n, float extinct, float step)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 val = src[i];
    val.w = exp(extinct * val.w * step);
    dst[i] = val;
}

// ------------------------------------------------------------------------------------------------ apron update
// VolumeGVDB::UpdateApron for a float channel with apron 1 (gvdb_volume_gvdb.cpp:4418-4453, kernel
// cuda_gvdb_operators.cuh:72-126): every texel of the six 10x10 faces of every brick takes the value of the voxel that
// occupies its index-space position in whichever leaf contains it (top-down point query), else `boundval`.  One CTA per
// leaf; reads only interior voxels and writes only apron texels, so the pass is race-free.  Both copies of the atlas
// are kept coherent: the 3-D array (through a surface object) and the brick-major layout.
__global__ void gx_update_apron_kernel(const __grid_constant__ GxParams P, cudaSurfaceObject_t surf, float* __restrict__ bricks,
                                       int nleaf, float boundval)
{
    const int leaf = blockIdx.x;
    if (leaf >= nleaf) return;
    const GxLeafRec L = P.leaf[leaf];
    if (L.vx < 0) return;
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < 6 * GX_BRICK_DIM * GX_BRICK_DIM; i += blockDim.x) {
        const int side = i / (GX_BRICK_DIM * GX_BRICK_DIM), u = (i / GX_BRICK_DIM) % GX_BRICK_DIM, v = i % GX_BRICK_DIM;
        int bx, by, bz;                                   // texel inside the 10^3 brick
        switch (side) {
        case 0:  bx = 0; by = u; bz = v; break;
        case 1:  bx = u; by = 0; bz = v; break;
        case 2:  bx = u; by = v; bz = 0; break;
        case 3:  bx = GX_BRICK_DIM - 1; by = u; bz = v; break;
        case 4:  bx = u; by = GX_BRICK_DIM - 1; bz = v; break;
        default: bx = u; by = v; bz = GX_BRICK_DIM - 1; break;
        }
        // index-space centre of the texel (getAtlasToWorld, cuda_gvdb_nodes.cuh:145-153)
        const float3 wpos = make_float3(float(L.px) + float(bx - 1) + 0.5f, float(L.py) + float(by - 1) + 0.5f, float(L.pz) + float(bz - 1) + 0.5f);
        float value = boundval;
        const int n = gx_node_at_point<GxSampler<GX_SAMPLER_TEX, false>>(P, wpos, cnt);
        if (n >= 0) {
            const GxLeafRec N = P.leaf[n];
            const float3 offs = make_float3(float(N.vx), float(N.vy), float(N.vz)) + (wpos - make_float3(float(N.px), float(N.py), float(N.pz)));
            value = surf3Dread<float>(surf, int(unsigned(offs.x)) * int(sizeof(float)), int(unsigned(offs.y)), int(unsigned(offs.z)));
        }
        surf3Dwrite(value, surf, (L.vx - 1 + bx) * int(sizeof(float)), L.vy - 1 + by, L.vz - 1 + bz);
        bricks[size_t(L.base) + (bz * GX_BRICK_DIM + by) * GX_BRICK_DIM + bx] = value;
    }
}

// value range per brick slot from the brick-major layout (after the aprons changed)
__global__ void gx_brick_ranges(const float* __restrict__ bricks, GxRange* __restrict__ range)
{
    const float* b = bricks + size_t(blockIdx.x) * GX_BRICK_STRIDE;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < GX_BRICK_DIM * GX_BRICK_DIM * GX_BRICK_DIM; i += blockDim.x) { const float v = b[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
    __shared__ float slo[8], shi[8];
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (blockDim.x >> 5); w++) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
        range[blockIdx.x].lo = lo; range[blockIdx.x].hi = hi;
    }
}

// SHADE_VOXEL occupancy bits: for every leaf, bit (z, y * 8 + x) of word z = (interior voxel (x, y, z) > thresh).  The
// comparison is the one raySurfaceVoxelBrick makes per DDA step (cuda_gvdb_raycast.cuh:241) on the value a texel-centre
// fetch returns, i.e. the stored texel itself.  One 64-thread CTA per leaf, one byte (a row of 8 voxels) per thread.
__global__ void gx_build_voxel_mask(const GxLeafRec* __restrict__ leaf, int nleaf, const float* __restrict__ bricks, float thresh,
                                    unsigned char* __restrict__ out)
{
    const int n = blockIdx.x;
    if (n >= nleaf) return;
    const int z = threadIdx.x >> 3, y = threadIdx.x & 7;
    unsigned bits = 0;
    if (leaf[n].vx >= 0) {
        const float* row = bricks + size_t(leaf[n].base) + ((z + 1) * GX_BRICK_DIM + (y + 1)) * GX_BRICK_DIM + 1;
        #pragma unroll
        for (int x = 0; x < 8; x++) bits |= (row[x] > thresh ? 1u : 0u) << x;
    }
    out[size_t(n) * 64 + z * 8 + y] = (unsigned char)bits;
}

// ------------------------------------------------------------------------------------------------ cross-GPU flags
// one thread: release-store of a sequence number (the frame's pixels were stored by kernels earlier in the stream)
__global__ void gx_signal_kernel(unsigned int* flag, unsigned int value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(value) : "memory");
}
__global__ void gx_signal_add_kernel(unsigned int* flag, unsigned int inc)
{
    __threadfence_system();
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(flag), "r"(inc) : "memory");
}
struct GxFlagList { unsigned int* p[16]; };
__global__ void gx_signal_many_kernel(GxFlagList L, int n, unsigned int value)
{
    __threadfence_system();
    if (threadIdx.x < n) asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(L.p[threadIdx.x]), "r"(value) : "memory");
}
// one thread: polls a LOCAL flag until it reaches `value`; bounded (~20 s) so that a lost peer cannot wedge the GPU —
// on timeout flag[1] is set to 0xDEAD and the stream continues
__global__ void gx_wait_kernel(unsigned int* flag, unsigned int value)
{
    const long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (int(v - value) >= 0) return;
        __nanosleep(256);
        if (clock64() - t0 > 40000000000ll) { flag[1] = 0xDEADu; return; }
    }
}
