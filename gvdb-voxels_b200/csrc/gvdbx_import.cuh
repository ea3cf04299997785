// gvdbx_import.cuh — non-template kernels used by gvdbx_api.cu only: reference pools / atlas -> compact tables and the
// brick-major atlas, tile assembly, sampler calibration.  (Kept out of gvdbx_device.cuh, which every gvdbx_k_*.cu includes.)
#pragma once
#include "gvdbx_extra.cuh"

// ------------------------------------------------------------------------------------------------ import kernels
// Error bits the import kernels raise (checked by the host after the import; a malformed VBX file or a stale VDBInfo must
// not turn into out-of-bounds reads in the render kernels).
#define GX_IMPORT_E_CHILDLIST 1      // node->mChildList points past the child-list pool
#define GX_IMPORT_E_CHILD     2      // a child entry points past the node pool of the level below
#define GX_IMPORT_E_LEAFSLOT  4      // a leaf's mValue brick does not lie inside the atlas

__global__ void gx_clear_import_bits(int* err) { atomicAnd(err, ~(GX_IMPORT_E_CHILDLIST | GX_IMPORT_E_CHILD | GX_IMPORT_E_LEAFSLOT)); }

// pool-0 / pool-1 (reference layout) -> compact tables.  One thread per child cell.
//   child list entry = Elem(0, lev-1, ndx) = grp | lev << 8 | ndx << 16, or 0xFFFFFFFFFFFFFFFF (src/gvdb_allocator.h:59-62,
//   gvdb_volume_gvdb.cpp:3015-3023); node->mChildList = Elem(1, lev, ndx) or ID_UNDEFL.
__global__ void gx_build_child_table(const char* __restrict__ nodelist, int nodewid, int nodecnt,
                                     const char* __restrict__ childlist, int childwid, int cells, unsigned long long listcnt, int childcnt,
                                     int* __restrict__ child_out, float4* __restrict__ npos_out, int* __restrict__ err)
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
        if (cndx >= listcnt) { atomicOr(err, GX_IMPORT_E_CHILDLIST); }
        else {
            const uint64_t* clist = reinterpret_cast<const uint64_t*>(childlist + cndx * size_t(childwid));
            c = int(clist[b] >> 16);
            if (c < -1 || c >= childcnt) { atomicOr(err, GX_IMPORT_E_CHILD); c = -1; }
        }
    }
    child_out[i] = c;
    if (b == 0) npos_out[n] = make_float4(float(node->mPos.x), float(node->mPos.y), float(node->mPos.z), 0.f);
}

// leaf records; a leaf without a brick (mValue.x < 0) keeps vx < 0 and is never sampled
__global__ void gx_build_leaf_table(const char* __restrict__ nodelist, int nodewid, int nodecnt, int brick_dim, int3 atlas_res,
                                    GxLeafRec* __restrict__ out, int* __restrict__ err)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nodecnt) return;
    const GxNode* node = reinterpret_cast<const GxNode*>(nodelist + size_t(n) * nodewid);
    GxLeafRec r;
    r.px = node->mPos.x; r.py = node->mPos.y; r.pz = node->mPos.z;
    r.vx = node->mValue.x; r.vy = node->mValue.y; r.vz = node->mValue.z;
    r.idx = n;
    r.pad = 0;
    if (r.vx >= 0 && (r.vx < 1 || r.vy < 1 || r.vz < 1 || r.vx - 1 + brick_dim > atlas_res.x || r.vy - 1 + brick_dim > atlas_res.y ||
                      r.vz - 1 + brick_dim > atlas_res.z)) {
        atomicOr(err, GX_IMPORT_E_LEAFSLOT);
        r.vx = -1;
    }
    out[n] = r;
}

// CTA-wide NaN-ignoring {min, max}
__device__ __forceinline__ void gx_block_minmax(float& lo, float& hi)
{
    __shared__ float slo[32], shi[32];
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < (int)((blockDim.x + 31) >> 5); w++) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
}

// Value range of every LEAF over all brick_dim^3 texels of its brick (interior + apron), read straight from the atlas
// array through a point-filter texture (exact texel values); one CTA per leaf.  A leaf without a brick is never culled.
__global__ void gx_leaf_ranges_array(cudaTextureObject_t point_tex, const GxLeafRec* __restrict__ leaf, int nleaf, int brick_dim,
                                     GxRange* __restrict__ out)
{
    const int n = blockIdx.x;
    if (n >= nleaf) return;
    const GxLeafRec L = leaf[n];
    float lo = INFINITY, hi = -INFINITY;
    if (L.vx >= 0) {
        const int total = brick_dim * brick_dim * brick_dim;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int x = i % brick_dim, y = (i / brick_dim) % brick_dim, z = i / (brick_dim * brick_dim);
            const float v = tex3D<float>(point_tex, float(L.vx - 1 + x) + 0.5f, float(L.vy - 1 + y) + 0.5f, float(L.vz - 1 + z) + 0.5f);
            lo = fminf(lo, v); hi = fmaxf(hi, v);
        }
    }
    gx_block_minmax(lo, hi);
    if (threadIdx.x == 0) {
        GxRange r;
        r.lo = (L.vx >= 0) ? lo : -INFINITY; r.hi = (L.vx >= 0) ? hi : INFINITY;
        out[n] = r;
    }
}

// brick-major copy for the linear sampler: block n = the brick_dim^3 texels of leaf n, x fastest; one CTA per leaf
__global__ void gx_copy_bricks_array(cudaTextureObject_t point_tex, const GxLeafRec* __restrict__ leaf, int nleaf, int brick_dim,
                                     int brick_stride, float* __restrict__ bricks)
{
    const int n = blockIdx.x;
    if (n >= nleaf) return;
    const GxLeafRec L = leaf[n];
    float* dst = bricks + size_t(n) * size_t(brick_stride);
    const int total = brick_dim * brick_dim * brick_dim;
    for (int i = threadIdx.x; i < brick_stride; i += blockDim.x) {
        float v = 0.f;
        if (i < total && L.vx >= 0) {
            const int x = i % brick_dim, y = (i / brick_dim) % brick_dim, z = i / (brick_dim * brick_dim);
            v = tex3D<float>(point_tex, float(L.vx - 1 + x) + 0.5f, float(L.vy - 1 + y) + 0.5f, float(L.vz - 1 + z) + 0.5f);
        }
        dst[i] = v;
    }
}

// brick slot -> leaf index (calibration aid only: gvdbx_sample_points takes atlas-space points)
__global__ void gx_build_slot_map(const GxLeafRec* __restrict__ leaf, int nleaf, int brick_dim, int cnt_x, int cnt_y, int* __restrict__ slot_leaf)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= nleaf || leaf[n].vx < 0) return;
    const int sx = (leaf[n].vx - 1) / brick_dim, sy = (leaf[n].vy - 1) / brick_dim, sz = (leaf[n].vz - 1) / brick_dim;
    slot_leaf[(size_t(sz) * cnt_y + sy) * cnt_x + sx] = n;
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
                                        const int* __restrict__ slot_leaf, float* __restrict__ out_tex, float* __restrict__ out_lin)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    out_tex[i] = tex3D<float>(P.tex, x, y, z);
    GxSampler<GX_SAMPLER_LINEAR, false> s(P);
    const int bd = P.brick_dim;
    int sx = int(x) / bd, sy = int(y) / bd, sz = int(z) / bd;
    const int lf = slot_leaf[(size_t(sz) * cnt_y + sy) * cnt_x + sx];
    if (lf < 0) { out_lin[i] = 0.f; return; }
    s.enter(P.leaf[lf]);
    out_lin[i] = s.tri(x, y, z);
}

// ------------------------------------------------------------------------------------------------ texture-unit peak
// Roofline denominator of the filtered modes: how many fp32 trilinear samples per second the texture units deliver on THIS
// GPU when nothing else is in the way — every thread marches a short segment through one brick of the imported atlas (so the
// texels are L1-resident, like a ray marching a brick), eight independent fetches in flight, nothing but the fetches and one
// add per sample.  Measured by gvdbx_measure_tex_peak, used by bench.py as `roofline.peak` for the TEX-bound deep mode.
__global__ void __launch_bounds__(256) gx_tex_peak_kernel(cudaTextureObject_t tex, int ares_x, int ares_y, int ares_z, int rounds, float spacing,
                                                          float* __restrict__ out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    // one brick-sized region per warp, lanes `spacing` voxels apart in x and y (an 8x4-pixel ray packet inside a brick: the
    // unit's rate depends on how many distinct texel quads one warp request touches)
    const int warp = tid >> 5, lane = tid & 31;
    const float bx = float((warp * 10) % max(ares_x - 10, 1)) + 1.0f, by = float(((warp / 7) * 10) % max(ares_y - 10, 1)) + 1.0f,
                bz = float(((warp / 53) * 10) % max(ares_z - 10, 1)) + 1.0f;
    float x = bx + spacing * float(lane & 7), y = by + 1.07f * spacing * float(lane >> 3), z = bz;
    const float dx = 0.11f, dy = 0.07f, dz = 0.22f;
    float acc = 0.f;
    for (int r = 0; r < rounds; r++) {
        float v[8];
        #pragma unroll
        for (int k = 0; k < 8; k++) v[k] = tex3D<float>(tex, x + dx * float(k), y + dy * float(k), z + dz * float(k));
        #pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k];
        x += 8.f * dx; y += 8.f * dy; z += 8.f * dz;
        if (z > bz + 7.5f) { x = bx + spacing * float(lane & 7); y = by + 1.07f * spacing * float(lane >> 3); z = bz; }
    }
    out[tid] = acc;
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
// cuda_gvdb_operators.cuh:72-126): every texel of the six faces of every brick takes the value of the voxel that occupies
// its index-space position in whichever leaf contains it (top-down point query), else `boundval`.  One CTA per leaf;
// reads only interior voxels and writes only apron texels, so the pass is race-free.  Both copies of the atlas are kept
// coherent: the 3-D array (through a surface object) and, if it has been built, the brick-major copy.
__global__ void gx_update_apron_kernel(const __grid_constant__ GxParams P, cudaSurfaceObject_t surf, float* __restrict__ bricks,
                                       int nleaf, float boundval)
{
    const int leaf = blockIdx.x;
    if (leaf >= nleaf) return;
    const GxLeafRec L = P.leaf[leaf];
    if (L.vx < 0) return;
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    const int BD = P.brick_dim;
    for (int i = threadIdx.x; i < 6 * BD * BD; i += blockDim.x) {
        const int side = i / (BD * BD), u = (i / BD) % BD, v = i % BD;
        int bx, by, bz;                                   // texel inside the brick
        switch (side) {
        case 0:  bx = 0; by = u; bz = v; break;
        case 1:  bx = u; by = 0; bz = v; break;
        case 2:  bx = u; by = v; bz = 0; break;
        case 3:  bx = BD - 1; by = u; bz = v; break;
        case 4:  bx = u; by = BD - 1; bz = v; break;
        default: bx = u; by = v; bz = BD - 1; break;
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
        if (bricks != nullptr) bricks[size_t(leaf) * size_t(P.brick_stride) + (bz * BD + by) * BD + bx] = value;
    }
}

// VolumeGVDB::UpdateApronFaces (gvdb_volume_gvdb.cpp:4461-4496, kernel gvdbUpdateApronFacesF cuda_gvdb_operators.cuh:27-61):
// the cheap apron update after a compute pass — face-adjacent bricks swap their boundary layers: for each of the three
// lower faces (-x, -y, -z) of a brick that has a neighbour there, the brick's first voxel layer goes into the neighbour's
// upper apron and the neighbour's last voxel layer into the brick's lower apron (res0 x res0 texels per face; edge and corner
// texels and faces without a neighbour are left alone).  The reference looks the neighbour up in a host-built table
// (UpdateNeighbors, :2527-2556: leaf at centre - brick width); here the point query does it on the fly.
// NOTE: the reference's table stores a MISSING neighbour as ElemNdx(ID_UNDEFL) = 0xFFFF while its kernel tests for 0xFFFFFFFF,
// so the stock kernel reads leaf 65535 for every boundary brick (out of bounds below 65536 leaves); the intended semantics
// are implemented, and on faces that do have a neighbour the result equals UpdateApron's (tested against the reference atlas).
__global__ void gx_update_apron_faces_kernel(const __grid_constant__ GxParams P, cudaSurfaceObject_t surf, float* __restrict__ bricks, int nleaf)
{
    const int leaf = blockIdx.x;
    if (leaf >= nleaf) return;
    const GxLeafRec L = P.leaf[leaf];
    if (L.vx < 0) return;
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    const int R = P.res[0], BD = P.brick_dim;
    const float d = float(R);
    const float3 ct = make_float3(float(L.px) + d * 0.5f, float(L.py) + d * 0.5f, float(L.pz) + d * 0.5f);
    __shared__ int nbr[3];
    if (threadIdx.x < 3) {
        const float3 q = make_float3(ct.x - (threadIdx.x == 0 ? d : 0.f), ct.y - (threadIdx.x == 1 ? d : 0.f), ct.z - (threadIdx.x == 2 ? d : 0.f));
        int n = gx_node_at_point<GxSampler<GX_SAMPLER_TEX, false>>(P, q, cnt);
        if (n >= 0 && P.leaf[n].vx < 0) n = -1;
        nbr[threadIdx.x] = n;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * R * R; i += blockDim.x) {
        const int side = i / (R * R), u = (i / R) % R, v = i % R;
        const int n = nbr[side];
        if (n < 0) continue;
        const GxLeafRec N = P.leaf[n];
        int3 vox, inc;
        if (side == 0)      { vox = make_int3(0, u, v); inc = make_int3(1, 0, 0); }
        else if (side == 1) { vox = make_int3(u, 0, v); inc = make_int3(0, 1, 0); }
        else                { vox = make_int3(u, v, 0); inc = make_int3(0, 0, 1); }
        const int3 vn = make_int3(vox.x + inc.x * (R - 1), vox.y + inc.y * (R - 1), vox.z + inc.z * (R - 1));   // neighbour's last layer
        const float mine = surf3Dread<float>(surf, (L.vx + vox.x) * int(sizeof(float)), L.vy + vox.y, L.vz + vox.z);
        const float theirs = surf3Dread<float>(surf, (N.vx + vn.x) * int(sizeof(float)), N.vy + vn.y, N.vz + vn.z);
        surf3Dwrite(mine, surf, (N.vx + vn.x + inc.x) * int(sizeof(float)), N.vy + vn.y + inc.y, N.vz + vn.z + inc.z);      // neighbour's upper apron
        surf3Dwrite(theirs, surf, (L.vx + vox.x - inc.x) * int(sizeof(float)), L.vy + vox.y - inc.y, L.vz + vox.z - inc.z); // own lower apron
        if (bricks != nullptr) {        // brick-local texel = interior voxel + 1
            bricks[size_t(n) * size_t(P.brick_stride) + ((vn.z + inc.z + 1) * BD + (vn.y + inc.y + 1)) * BD + (vn.x + inc.x + 1)] = mine;
            bricks[size_t(leaf) * size_t(P.brick_stride) + ((vox.z - inc.z + 1) * BD + (vox.y - inc.y + 1)) * BD + (vox.x - inc.x + 1)] = theirs;
        }
    }
}

// SHADE_VOXEL occupancy bits (8^3 bricks): for every leaf, bit (y * 8 + x) of word z = (interior voxel (x, y, z) > thresh).
// The comparison is the one raySurfaceVoxelBrick makes per DDA step (cuda_gvdb_raycast.cuh:241) on the value a texel-centre
// fetch returns, i.e. the stored texel itself.  One 64-thread CTA per leaf, one byte (a row of 8 voxels) per thread.
__global__ void gx_build_voxel_mask(cudaTextureObject_t point_tex, const GxLeafRec* __restrict__ leaf, int nleaf, float thresh,
                                    unsigned char* __restrict__ out)
{
    const int n = blockIdx.x;
    if (n >= nleaf) return;
    const int z = threadIdx.x >> 3, y = threadIdx.x & 7;
    unsigned bits = 0;
    const GxLeafRec L = leaf[n];
    if (L.vx >= 0) {
        #pragma unroll
        for (int x = 0; x < 8; x++)
            bits |= (tex3D<float>(point_tex, float(L.vx + x) + 0.5f, float(L.vy + y) + 0.5f, float(L.vz + z) + 0.5f) > thresh ? 1u : 0u) << x;
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
// on timeout flag[1] is set to 0xDEAD, a sticky error bit is raised in the context (gvdbx_sync then fails) and the stream continues
#define GX_ERR_WAIT_TIMEOUT 0x100
__global__ void gx_wait_kernel(unsigned int* flag, unsigned int value, int* err)
{
    const long long t0 = clock64();
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (int(v - value) >= 0) return;
        __nanosleep(256);
        if (clock64() - t0 > 40000000000ll) { flag[1] = 0xDEADu; if (err) atomicOr(err, GX_ERR_WAIT_TIMEOUT); return; }
    }
}
