// gvdbx_device.cuh — device side of the B200-native GVDB ray-cast render path (sm_100a).
//
// What is computed (reference, relative to source/gvdb_library/kernels/):
//   ray generation      cuda_gvdb_geom.cuh:48-63   (getViewPos / getViewRay)
//   slab test           cuda_gvdb_geom.cuh:85-98   (rayBoxIntersect)
//   hierarchical DDA    cuda_gvdb_dda.cuh:38-91    (HDDAState), cuda_gvdb_raycast.cuh:543-611 (rayCast)
//   brick samplers      cuda_gvdb_raycast.cuh:227-265 (voxel), :281-300 (trilinear), :389-410 + :186-197 (level set),
//                       :485-533 (deep) + cuda_gvdb_dda.cuh:20-23 (transfer)
//   gradients           cuda_gvdb_raycast.cuh:132-157
//   shading + packing   cuda_gvdb_module.cu:38-57, :60-181
//
// How it is laid out here (B200-first, not a translation):
//   * all per-frame state (scene + tree geometry + table pointers) travels as ONE __grid_constant__ kernel parameter
//     -> constant-bank reads, instead of the reference's VDBInfo struct in global memory (42 LDG sites per kernel);
//   * the tree is traversed through compact tables built at import time: per level one int32 child table indexed by
//     NODE index (4 B per cell, one dependent load per DDA step instead of node->mChildList->clist[b] = 3 loads of
//     64 B + 8 B), one float4 position record per internal node and one 32-B record per leaf;
//   * the brick atlas is available both as the caller's 3-D array through a texture object (bit-exact hardware
//     trilinear) and re-laid out brick-major (4 KB per 10^3 brick, contiguous) for plain vectorisable loads with a
//     software emulation of the texture unit's 1.8 fixed-point filtering;
//   * warps are 8x4 pixel tiles; the per-level traversal stack lives in registers (levels 1..4), not local memory.
//
// Floating-point contract: SHADE_VOXEL output must be BIT-EXACT against the reference module compiled with its own
// flag (--use_fast_math).  This file is therefore compiled with --use_fast_math too, and every expression on the
// exact path keeps the reference's operand order, literal types (several literals are double) and division /
// rsqrt / exp forms so that nvcc's contraction and approximation choices coincide.  Do not "simplify" them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gvdbx_types.h"

#define GX_MODE_VOXEL     0
#define GX_MODE_TRILINEAR 1
#define GX_MODE_LEVELSET  2
#define GX_MODE_DEEP      3
#define GX_MODE_TRICUBIC  4      // SHADE_TRICUBIC  (brick function + pixel mode)
#define GX_MODE_EMPTYSKIP 5      // SHADE_EMPTYSKIP (brick function + pixel mode)
#define GX_MODE_SHADOW    6      // rayShadowBrick: brick function only (secondary ray of GX_MODE_DEEPSHADOW)
#define GX_MODE_SECTION2D 7      // SHADE_SECTION2D (pixel mode, no ray cast)
#define GX_MODE_SECTION3D 8      // SHADE_SECTION3D (pixel mode: section plane + trilinear surface)
#define GX_MODE_DEEPSHADOW 9     // deep march + shadow march towards the light (BASELINE.json config 4)

#define GX_SAMPLER_TEX    0
#define GX_SAMPLER_LINEAR 1

// brick-major copy of the atlas (linear sampler): one block of brick_dim^3 floats per LEAF, brick_dim = res0 + 2 (apron 1),
// padded to a multiple of 256 floats.  8^3 bricks (Configure(.., 3), the UNI kernels): 10^3 in 1024 floats = one 4 KB block.
#define GX_BRICK_STRIDE   1024
#define GX_BRICK_DIM      10

// ------------------------------------------------------------------------------------------------ parameters
struct alignas(16) GxLeafRec {     // 32 B, one per level-0 node
    int px, py, pz;                // mPos   : index-space min corner
    int idx;                       // the leaf's own index = its block in the brick-major copy (offset = idx * brick_stride, 64-bit)
    int vx, vy, vz;                // mValue : atlas texel of the first interior voxel
    int pad;
};
// Value range of every brick slot over all 10^3 texels (interior + apron), built when the atlas is imported.  Hardware
// and software trilinear filtering are convex combinations of texels (non-negative 8-bit weights summing to 256), so no
// sample taken inside the brick can leave [lo, hi]: a brick whose range cannot satisfy the mode's acceptance test is
// skipped without changing any result.  (The reference keeps an unused mVRange field in every node for this purpose,
// src/gvdb_node.h:34.)
struct GxRange { float lo, hi; };

struct GxParams {
    // ---- scene (ScnInfo fields the path reads)
    int      width, height;
    float    camnear, camfar;
    float3   campos, cams, camu, camv;
    float3   light_pos;
    float3   slice_pnt, slice_norm; // SCN_SLICE_PNT / SCN_SLICE_NORM (section modes)
    float3   shadow_params;        // x = SHADOWAMT, y = SHADOWBIAS
    float4   backclr;
    float    xform[16], invxform[16], invxrot[16];
    float3   extinct;              // x = EXTINCT, y = ALBEDO
    float3   steps;                // x = DIRECTSTEP, y = SHADOWSTEP, z = FINESTEP
    float3   cutoff;               // x = MINVAL, y = ALPHACUT
    float3   thresh;               // x = THRESH, y = VMIN, z = VMAX
    const float4* transfer;
    const float4* transfer_deep;       // per frame: {rgb, exp(EXTINCT * alpha * DIRECTSTEP)} of every entry (gx_build_deep_lut); set whenever `transfer` is, for the
                                       // deep modes and for every block handed to a plugin kernel (gvdbx_kernel_params); the deep marchers require it
    const float*  dbuf;
    // ---- tree geometry (VDBInfo fields the path reads)
    int      dim[GX_MAXLEV];
    int      res[GX_MAXLEV];
    float3   vdel[GX_MAXLEV];
    int3     noderange[GX_MAXLEV];
    int      top_lev;
    float    epsilon;
    float3   bmin, bmax;
    // ---- compact traversal tables
    const int*       child[GX_MAXLEV];   // child[lev][node * cells(lev) + b] = index at lev-1, or -1
    const float4*    npos[GX_MAXLEV];    // npos[lev][node] = {float(mPos), 0}: every user converts the corner to float first (lev >= 1)
    const GxLeafRec* leaf;               // leaf[node]                              (lev == 0)
    // ---- the reference's own pools (GX_REF_LAYOUT builds only): VDBInfo::nodelist / nodewid / childlist / childwid
    const char* ref_nodes[GX_MAXLEV];
    const char* ref_clist[GX_MAXLEV];
    int         ref_nodewid[GX_MAXLEV], ref_childwid[GX_MAXLEV];
    // ---- atlas
    cudaTextureObject_t tex;             // caller's 3-D array, linear filter, unnormalised, clamp
    cudaTextureObject_t clr_tex;         // colour channel (uchar4 atlas, VDBInfo::clr_chan); 0 = CHAN_UNDEF
    const float*        bricks;          // brick-major copy, one block per leaf (built on first use of the linear sampler)
    int                 brick_dim;       // res0 + 2 (texels per brick edge incl. apron)
    int                 brick_stride;    // floats per block
    const GxRange*      range;           // value range per LEAF (same index as `leaf`); null = no culling
    const unsigned long long* vmask;     // SHADE_VOXEL: per leaf 8 x 64 bits, bit (z, y * 8 + x) = voxel value > THRESH; null = fetch
    // ---- output
    uchar4*  out;
    float4*  dbg;                        // 3 x 16 B per pixel (debug variant only)
    unsigned long long* counters;        // 6 x u64 (count variant only)
    int      out_stride;                 // pixels per output row
    int      x0, y0, x1, y1;             // pixel rectangle [x0,x1) x [y0,y1)
    // ---- sub-pixel sampling (GX_FLAG_SPP): spp rays per pixel on a spp_grid x spp_grid pattern, averaged before packing
    int      spp, spp_grid;
    float    spp_inv_grid, spp_inv;
    // ---- tile-list mode (multi-GPU): tiles of tile_w x tile_h pixels with id % nranks == rank, packed tile after tile
    // (square tiles for the peer frame ring; full-width bands — tile_w >= width, tiles_x == 1 — for the host frame ring)
    int      tile_w, tile_h, tiles_x, ntiles, rank, nranks;
};

// ------------------------------------------------------------------------------------------------ small vector algebra
// component-wise, written out so that operand order is explicit on the exact path
__device__ __forceinline__ float3 gx3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ float3 operator+(float3 a, float b)  { return make_float3(a.x + b, a.y + b, a.z + b); }
__device__ __forceinline__ float3 operator-(float3 a, float b)  { return make_float3(a.x - b, a.y - b, a.z - b); }
__device__ __forceinline__ float3 operator*(float3 a, float b)  { return make_float3(a.x * b, a.y * b, a.z * b); }
__device__ __forceinline__ float3 operator*(float b, float3 a)  { return make_float3(b * a.x, b * a.y, b * a.z); }
__device__ __forceinline__ float3 operator/(float b, float3 a)  { return make_float3(b / a.x, b / a.y, b / a.z); }
__device__ __forceinline__ float3 operator-(float b, float3 a)  { return make_float3(b - a.x, b - a.y, b - a.z); }
__device__ __forceinline__ void   operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ void   operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__device__ __forceinline__ float  gx_dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 gx_normalize(float3 v) { float inv = rsqrtf(gx_dot(v, v)); return v * inv; }
__device__ __forceinline__ float3 gx_floor(float3 a) { return make_float3(floorf(a.x), floorf(a.y), floorf(a.z)); }
__device__ __forceinline__ float3 gx_fabs(float3 a) { return make_float3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
__device__ __forceinline__ float3 gx_f3(int3 a) { return make_float3(float(a.x), float(a.y), float(a.z)); }
__device__ __forceinline__ int3   gx_i3(float3 a) { return make_int3(int(a.x), int(a.y), int(a.z)); }
// column-major 4x4 times (v,1)  (cuda_math.cuh:1472)
__device__ __forceinline__ float3 gx_mmult(const float* m, float3 v)
{
    float3 p;
    p.x = v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12];
    p.y = v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13];
    p.z = v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14];
    return p;
}

// MUFU.RCP — what div.approx.ftz lowers to on sm_100a (x / d == x * rcp(d) bit for bit under --use_fast_math), so a
// reciprocal that is reused many times per ray can be taken once without changing any result.
__device__ __forceinline__ float gx_rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------------ work counters
struct GxCount { unsigned int s_tri, s_pt, n_dda, n_desc, s_lut, rays; };

// ------------------------------------------------------------------------------------------------ samplers
// Both take ATLAS-space coordinates exactly as the reference passes them to tex3D (p + o).
// UNI = every level of the tree has log2dim 3 (Configure(3,3,3,3,3), the reference's default and all its samples):
// res = 8, dim = 3, vdel[lev] = 8^lev become compile-time constants.
template <int SAMPLER, bool UNI_> struct GxSampler;

template <bool UNI_> struct GxSampler<GX_SAMPLER_TEX, UNI_> {
    static constexpr bool UNI = UNI_;
    cudaTextureObject_t tex;
    __device__ __forceinline__ GxSampler(const GxParams& P) : tex(P.tex) {}
    __device__ __forceinline__ void enter(const GxLeafRec&) {}
    // filtered fetch at atlas coordinate (x,y,z)
    __device__ __forceinline__ float tri(float x, float y, float z) const { return tex3D<float>(tex, x, y, z); }
    // exact voxel value at integer atlas texel (ix,iy,iz): texel centres sit at +0.5 -> filter weights are 0
    __device__ __forceinline__ float point(float x, float y, float z) const { return tex3D<float>(tex, x, y, z); }
};

// Software model of the texture unit's trilinear filter on the brick-major layout, fitted on a B200 against
// tex3D<float> (tests/calib_trilinear.py, profiles/r01_trilinear_calibration.md).  The unit does NOT lerp with three
// 8-bit weights; it builds eight 8-bit corner weights that sum to 256 by a hierarchical split with round-half-up:
//   a  = rhu(frac(x - 0.5) * 256) per axis (carry into the texel index when it reaches 256)
//   S1 = az, S0 = 256 - az                                   (z slices, exact)
//   per slice S:  X1 = rhu(S * ax / 256), X0 = S - X1        (x split)
//                 w11 = rhu(X1 * ay / 256), w10 = X1 - w11   (y split of the x+1 column: far corner rounded)
//                 w00 = rhu(X0 * (256 - ay) / 256), w01 = X0 - w00   (y split of the x column: near corner rounded)
//   result = sum(w_c * T_c) / 256
// This reproduces all eight weights on 262144 calibration points; the weighted sum differs from the hardware's by
// at most 1 ulp-of-result (the unit accumulates wider than fp32).
// The two x-neighbours of a sample are adjacent floats; a brick is one contiguous 4 KB block, so a warp marching
// through one brick touches at most 32 consecutive 128-B lines.
template <bool UNI_> struct GxSampler<GX_SAMPLER_LINEAR, UNI_> {
    static constexpr bool UNI = UNI_;
    const float* bricks;
    const float* b;      // current brick
    int ox, oy, oz;      // atlas texel index of the brick's texel (0,0,0) = mValue - apron
    int bd_, stride_;    // brick edge / block size for trees with other brick sizes (compile-time 10 / 1024 when UNI)
    __device__ __forceinline__ GxSampler(const GxParams& P) : bricks(P.bricks), b(P.bricks), ox(0), oy(0), oz(0), bd_(P.brick_dim), stride_(P.brick_stride) {}
    __device__ __forceinline__ int bd() const { return UNI ? GX_BRICK_DIM : bd_; }
    __device__ __forceinline__ void enter(const GxLeafRec& L)
    {
        b = bricks + size_t(L.idx) * size_t(UNI ? GX_BRICK_STRIDE : stride_);
        ox = L.vx - 1; oy = L.vy - 1; oz = L.vz - 1;
    }
    static __device__ __forceinline__ void split(float c, int o, int& i, int& a)
    {
        float cb = c - 0.5f;
        float f = floorf(cb);
        a = __float2int_rd(fmaf(cb - f, 256.0f, 0.5f));
        i = int(f) - o;
        if (a >= 256) { a = 0; i++; }
    }
    __device__ __forceinline__ float tri(float x, float y, float z) const
    {
        int ix, iy, iz, ax, ay, az;
        split(x, ox, ix, ax); split(y, oy, iy, ay); split(z, oz, iz, az);
        // in-brick samples only ever need texels 0..bd-1; clamp so that zero-weight neighbours stay inside the brick
        const int BD = bd();
        const int ix1 = min(max(ix + 1, 0), BD - 1), iy1 = min(max(iy + 1, 0), BD - 1), iz1 = min(max(iz + 1, 0), BD - 1);
        ix = min(max(ix, 0), BD - 1); iy = min(max(iy, 0), BD - 1); iz = min(max(iz, 0), BD - 1);
        const float* r00 = b + (iz * BD + iy) * BD;
        const float* r10 = b + (iz * BD + iy1) * BD;
        const float* r01 = b + (iz1 * BD + iy) * BD;
        const float* r11 = b + (iz1 * BD + iy1) * BD;
        const float c000 = __ldg(r00 + ix), c100 = __ldg(r00 + ix1);
        const float c010 = __ldg(r10 + ix), c110 = __ldg(r10 + ix1);
        const float c001 = __ldg(r01 + ix), c101 = __ldg(r01 + ix1);
        const float c011 = __ldg(r11 + ix), c111 = __ldg(r11 + ix1);
        const int by = 256 - ay;
        const int s0 = 256 - az, s1 = az;
        const int x1a = (s0 * ax + 128) >> 8, x0a = s0 - x1a;
        const int x1b = (s1 * ax + 128) >> 8, x0b = s1 - x1b;
        const int w110 = (x1a * ay + 128) >> 8, w100 = x1a - w110;
        const int w000 = (x0a * by + 128) >> 8, w010 = x0a - w000;
        const int w111 = (x1b * ay + 128) >> 8, w101 = x1b - w111;
        const int w001 = (x0b * by + 128) >> 8, w011 = x0b - w001;
        float acc = float(w000) * c000;
        acc = fmaf(float(w100), c100, acc);
        acc = fmaf(float(w010), c010, acc);
        acc = fmaf(float(w110), c110, acc);
        acc = fmaf(float(w001), c001, acc);
        acc = fmaf(float(w101), c101, acc);
        acc = fmaf(float(w011), c011, acc);
        acc = fmaf(float(w111), c111, acc);
        return acc * (1.0f / 256.0f);
    }
    __device__ __forceinline__ float point(float x, float y, float z) const
    {
        int ix = int(x) - ox, iy = int(y) - oy, iz = int(z) - oz;    // x = texel + 0.5 -> truncation gives the texel
        return __ldg(b + (iz * bd() + iy) * bd() + ix);
    }
};

// ------------------------------------------------------------------------------------------------ tree geometry accessors
template <class S> __device__ __forceinline__ int gx_res(const GxParams& P, int lev) { return S::UNI ? 8 : P.res[lev]; }
template <class S> __device__ __forceinline__ int gx_dim(const GxParams& P, int lev) { return S::UNI ? 3 : P.dim[lev]; }
template <class S> __device__ __forceinline__ float3 gx_vdel(const GxParams& P, int lev)
{
    if (S::UNI) { const float v = __int_as_float((127 + 3 * lev) << 23); return make_float3(v, v, v); }     // 8^lev, exact, from exponent bits
    return P.vdel[lev];
}

// ------------------------------------------------------------------------------------------------ tree tables
// Two layouts behind four accessors.  Default: the compact tables built at import (§2 of DESIGN.md).  GX_REF_LAYOUT (the
// module-level drop-in, csrc/gvdbx_module.cu): the reference's own pools as VDBInfo points at them — 64-byte node
// records (mPos@4, mValue@16, mChildList@48) and per-node lists of 64-bit child entries (index = entry >> 16, all ones =
// no child), kernels/cuda_gvdb_nodes.cuh:24-35, :115-129, :184-196.
typedef float4 gx_npos_t;          // index-space min corner of a node, converted to float once at import
#ifdef GX_REF_LAYOUT
typedef const unsigned long long* gx_ctab_t;
__device__ __forceinline__ const GxNode* gx_ref_node(const GxParams& P, int lev, int node)
{
    return reinterpret_cast<const GxNode*>(P.ref_nodes[lev] + size_t(node) * P.ref_nodewid[lev]);
}
__device__ __forceinline__ gx_ctab_t gx_table(const GxParams& P, int lev, int node, int /*dim*/)
{
    const unsigned long long listid = gx_ref_node(P, lev, node)->mChildList;
    if (listid == GX_ID_UNDEFL) return nullptr;
    return reinterpret_cast<gx_ctab_t>(P.ref_clist[lev] + size_t(listid >> 16) * P.ref_childwid[lev]);
}
__device__ __forceinline__ int gx_child(gx_ctab_t t, int b) { return t ? int(__ldg(t + b) >> 16) : -1; }
__device__ __forceinline__ gx_npos_t gx_node_pos(const GxParams& P, int lev, int node)
{
    const GxNode* n = gx_ref_node(P, lev, node);
    return make_float4(float(n->mPos.x), float(n->mPos.y), float(n->mPos.z), 0.f);
}
__device__ __forceinline__ GxLeafRec gx_leaf(const GxParams& P, int node)
{
    const GxNode* n = gx_ref_node(P, 0, node);
    GxLeafRec r;
    r.px = n->mPos.x; r.py = n->mPos.y; r.pz = n->mPos.z; r.idx = node;
    r.vx = n->mValue.x; r.vy = n->mValue.y; r.vz = n->mValue.z; r.pad = 0;
    return r;
}
#else
typedef const int* gx_ctab_t;
__device__ __forceinline__ gx_ctab_t gx_table(const GxParams& P, int lev, int node, int dim) { return P.child[lev] + (size_t(node) << (3 * dim)); }
__device__ __forceinline__ int gx_child(gx_ctab_t t, int b) { return __ldg(t + b); }
__device__ __forceinline__ gx_npos_t gx_node_pos(const GxParams& P, int lev, int node) { return __ldg(&P.npos[lev][node]); }
__device__ __forceinline__ GxLeafRec gx_leaf(const GxParams& P, int node) { return P.leaf[node]; }
#endif

// ------------------------------------------------------------------------------------------------ geometry
// slab test: (tnear clamped to >= 0, tfar, 0 | NOHIT)                       cuda_gvdb_geom.cuh:85-98
__device__ __forceinline__ float3 gx_ray_box(float3 rpos, float3 rdir, float3 vmin, float3 vmax)
{
    float t0 = (vmin.x - rpos.x) / rdir.x;
    float t1 = (vmax.x - rpos.x) / rdir.x;
    float t2 = (vmin.y - rpos.y) / rdir.y;
    float t3 = (vmax.y - rpos.y) / rdir.y;
    float t4 = (vmin.z - rpos.z) / rdir.z;
    float t5 = (vmax.z - rpos.z) / rdir.z;
    float tn = fmaxf(fmaxf(fminf(t0, t1), fminf(t2, t3)), fminf(t4, t5));
    float tf = fminf(fminf(fmaxf(t0, t1), fmaxf(t2, t3)), fmaxf(t4, t5));
    tn = (tn < 0) ? 0.0 : tn;
    return make_float3(tn, tf, (tf < tn || tf < 0) ? GX_NOHIT : 0);
}

// ------------------------------------------------------------------------------------------------ hierarchical DDA
struct GxDDA {
    float3 pos, dir;
    float3 inv;     // MUFU.RCP of dir, taken once per ray
    float3 tDel;
    float3 t;
    int3   p;
    float3 tSide;
    bool   mx, my, mz;      // axis mask of the pending step (the reference's int3 mask, kept as predicates)

    __device__ __forceinline__ void set_ray(float3 startPos, float3 startDir, float3 startT)
    {
        pos = startPos; dir = startDir;
        inv = make_float3(gx_rcp_approx(dir.x), gx_rcp_approx(dir.y), gx_rcp_approx(dir.z));
        t = startT;
    }
    // pStep = isign3(dir): +1 for dir > 0, else -1 (also for 0)           cuda_math.cuh:1541-1545
    // (not stored: one compare + select where it is used costs less than three live registers)
    __device__ __forceinline__ float3 fstep() const { return make_float3((dir.x > 0) ? 1.0f : -1.0f, (dir.y > 0) ? 1.0f : -1.0f, (dir.z > 0) ? 1.0f : -1.0f); }
    // cuda_gvdb_dda.cuh:61-66
    __device__ __forceinline__ void prepare(float3 vmin, float3 vdel)
    {
        tDel = gx_fabs(vdel * inv);                 // vdel / dir
        float3 pFlt = (pos + t.x * dir - vmin) / vdel;
        tSide = ((gx_floor(pFlt) - pFlt + 0.5f) * fstep() + 0.5) * tDel + t.x;
        p = gx_i3(gx_floor(pFlt));
    }
    // the same with `inv` holding |1 / dir| (the walker, gvdbx_walk.cuh): |vdel * inv| == vdel * |inv| bit for bit, vdel > 0
    __device__ __forceinline__ void prepare_abs(float3 vmin, float3 vdel)
    {
        tDel = vdel * inv;
        float3 pFlt = (pos + t.x * dir - vmin) / vdel;
        tSide = ((gx_floor(pFlt) - pFlt + 0.5f) * fstep() + 0.5) * tDel + t.x;
        p = gx_i3(gx_floor(pFlt));
    }
    // cuda_gvdb_dda.cuh:70-75 (brick: child size 1, no "+ t.x")
    __device__ __forceinline__ void prepare_leaf(float3 vmin)
    {
        tDel = gx_fabs(inv);                        // 1.0f / dir
        float3 pFlt = pos + t.x * dir - vmin;
        tSide = ((gx_floor(pFlt) - pFlt + 0.5f) * fstep() + 0.5) * tDel;
        p = gx_i3(gx_floor(pFlt));
    }
    // cuda_gvdb_dda.cuh:78-83 (tie rules: x beats z on <=, y beats x, z beats y)
    __device__ __forceinline__ void next()
    {
        mx = (tSide.x < tSide.y) & (tSide.x <= tSide.z);
        my = (tSide.y < tSide.z) & (tSide.y <= tSide.x);
        mz = (tSide.z < tSide.x) & (tSide.z <= tSide.y);
        t.y = mx ? tSide.x : (my ? tSide.y : tSide.z);
    }
    // Next (:78-83) + Step (:86-90) in one predicated block — (sx, sy, sz) = isign3(dir) as integers; t.y = the selected side,
    // tSide += float(mask) * tDel (a 0 mask still multiplies: 0 * inf = NaN on an axis-parallel ray exactly like the reference),
    // p += mask * pStep.  The caller moves t.y into t.x.  6 FSETP, 5 FSEL, 3 FFMA, 3 predicated IADD.
    __device__ __forceinline__ void next_step(int sx, int sy, int sz)
    {
        asm("{\n\t"
        ".reg .pred mx, my, mz;\n\t"
        ".reg .f32 fm;\n\t"
        "setp.lt.ftz.f32 mx, %1, %2;\n\t"
        "setp.le.and.ftz.f32 mx, %1, %3, mx;\n\t"
        "setp.lt.ftz.f32 my, %2, %3;\n\t"
        "setp.le.and.ftz.f32 my, %2, %1, my;\n\t"
        "setp.lt.ftz.f32 mz, %3, %1;\n\t"
        "setp.le.and.ftz.f32 mz, %3, %2, mz;\n\t"
        "selp.f32 %0, %2, %3, my;\n\t"
        "selp.f32 %0, %1, %0, mx;\n\t"
        "selp.f32 fm, 0f3F800000, 0f00000000, mx;\n\t"
        "fma.rn.ftz.f32 %1, fm, %7, %1;\n\t"
        "selp.f32 fm, 0f3F800000, 0f00000000, my;\n\t"
        "fma.rn.ftz.f32 %2, fm, %8, %2;\n\t"
        "selp.f32 fm, 0f3F800000, 0f00000000, mz;\n\t"
        "fma.rn.ftz.f32 %3, fm, %9, %3;\n\t"
        "@mx add.s32 %4, %4, %10;\n\t"
        "@my add.s32 %5, %5, %11;\n\t"
        "@mz add.s32 %6, %6, %12;\n\t"
        "}"
        : "=&f"(t.y), "+f"(tSide.x), "+f"(tSide.y), "+f"(tSide.z), "+r"(p.x), "+r"(p.y), "+r"(p.z)
        : "f"(tDel.x), "f"(tDel.y), "f"(tDel.z), "r"(sx), "r"(sy), "r"(sz));
    }
    // cuda_gvdb_dda.cuh:86-90: tSide += float(mask) * tDel (a 0 mask still multiplies: 0 * inf = NaN on an axis-parallel
    // ray, exactly like the reference), p += mask * pStep.  The float mask comes from a select, not an int -> float
    // conversion (quarter-rate pipe).
    __device__ __forceinline__ void step()
    {
        t.x = t.y;
        tSide.x = fmaf(mx ? 1.0f : 0.0f, tDel.x, tSide.x);
        tSide.y = fmaf(my ? 1.0f : 0.0f, tDel.y, tSide.y);
        tSide.z = fmaf(mz ? 1.0f : 0.0f, tDel.z, tSide.z);
        p.x += mx ? ((dir.x > 0) ? 1 : -1) : 0;
        p.y += my ? ((dir.y > 0) ? 1 : -1) : 0;
        p.z += mz ? ((dir.z > 0) ? 1 : -1) : 0;
    }
};

// ------------------------------------------------------------------------------------------------ transfer function
// cuda_gvdb_dda.cuh:20-23 — note the double-precision clamp and multiply.
__device__ __forceinline__ float4 gx_transfer(const GxParams& P, float v)
{
    return __ldg(&P.transfer[int(min(1.0, max(0.0, (v - P.thresh.x) / (P.thresh.z - P.thresh.y))) * 16300.0f)]);
}

// colour channel fetch at atlas position p (truncated to the texel): getColorF, cuda_gvdb_raycast.cuh:206-209 +
// make_float4(uchar4), cuda_math.cuh:221-224
__device__ __forceinline__ float4 gx_color(const GxParams& P, float3 p)
{
    const uchar4 a = tex3D<uchar4>(P.clr_tex, (int)p.x, (int)p.y, (int)p.z);
    return make_float4(float(a.x) / 255.0f, float(a.y) / 255.0f, float(a.z) / 255.0f, float(a.w) / 255.0f);
}

// depth-buffer clip: cuda_gvdb_raycast.cuh:343-370
__device__ __forceinline__ float gx_depth_max(const GxParams& P, float3 rayDir, int px, int py)
{
    if (P.dbuf != nullptr) {
        float3 w;
        w.x = rayDir.x * P.xform[0] + rayDir.y * P.xform[4] + rayDir.z * P.xform[8];
        w.y = rayDir.x * P.xform[1] + rayDir.y * P.xform[5] + rayDir.z * P.xform[9];
        w.z = rayDir.x * P.xform[2] + rayDir.y * P.xform[6] + rayDir.z * P.xform[10];
        float z = P.dbuf[(P.height - 1 - py) * P.width + px];
        float n = P.camnear, f = P.camfar;
        float lin = (-n * f / (f - n)) / (z - (f / (f - n)));
        return lin / sqrtf(gx_dot(w, w));
    }
    return INFINITY;
}

// ------------------------------------------------------------------------------------------------ gradients
// central differences at +-0.5 texel (the apron is one texel wide)   cuda_gvdb_raycast.cuh:132-157
template <class S>
__device__ __forceinline__ float3 gx_gradient(const S& smp, float3 p, GxCount& cnt, bool levelset)
{
    float3 g;
    float xm = smp.tri(p.x - .5, p.y, p.z), xp = smp.tri(p.x + .5, p.y, p.z);
    float ym = smp.tri(p.x, p.y - .5, p.z), yp = smp.tri(p.x, p.y + .5, p.z);
    float zm = smp.tri(p.x, p.y, p.z - .5), zp = smp.tri(p.x, p.y, p.z + .5);
    cnt.s_tri += 6;
    if (levelset) { g.x = 1.0 * (xp - xm); g.y = 1.0 * (yp - ym); g.z = 1.0 * (zp - zm); }   // positive gradient
    else          { g.x = 1.0 * (xm - xp); g.y = 1.0 * (ym - yp); g.z = 1.0 * (zm - zp); }   // negative gradient
    return gx_normalize(g);
}

// ------------------------------------------------------------------------------------------------ brick functions
// cpos = atlas position of a surface hit (p + o exactly as the brick function holds it): the colour channel is fetched
// there AFTER the ray cast returns (gx_hit_color), so that the constant surface colour (1,1,1,1) does not have to live in
// registers through the whole traversal
struct GxHit { float3 hit, norm; float4 clr; float t; int leaf; int3 vox; float3 cpos; };

// SHADE_VOXEL: per-voxel DDA inside the brick                           cuda_gvdb_raycast.cuh:227-265
template <class S>
__device__ __forceinline__ void gx_brick_voxel(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                               GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    if (P.range != nullptr && !(__ldg(&P.range[nodeid].hi) > P.thresh.x)) return;             // no voxel above THRESH
    smp.enter(L);
    float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const int res0 = gx_res<S>(P, 0);

    GxDDA dda;
    dda.set_ray(pos, dir, t);
    dda.prepare_leaf(vmin);
    // isign3(dir) from the sign bit: differs from (dir > 0 ? 1 : -1) only for a zero component, whose axis never steps
    const int sx = (__float_as_int(dir.x) >> 31) | 1, sy = (__float_as_int(dir.y) >> 31) | 1, sz = (__float_as_int(dir.z) >> 31) | 1;

    // Occupancy bits instead of one dependent point fetch per voxel step: `value > THRESH` was evaluated once per voxel
    // when THRESH was set (gx_build_voxel_mask, on the exact texel values a centre fetch returns), so the per-step test is
    // a bit test on a 64-bit z-slice that is reloaded only when the ray changes slice.
    const unsigned long long* mk = (P.vmask != nullptr && res0 == 8) ? P.vmask + size_t(nodeid) * 8 : nullptr;
    unsigned long long slice = 0;
    int slice_z = -1;
    // A brick DDA only ever moves away from its entry voxel in the direction of the ray, so every voxel it can visit
    // lies in the "forward octant" box of the entry voxel.  No occupancy bit in that box = no hit in this brick, whatever
    // the exact path: skip the voxel walk (the work counters of the reference semantics are kept by the COUNT variant,
    // which does not take this shortcut: P.range == nullptr there).
    if (mk != nullptr && P.range != nullptr && unsigned(dda.p.x | dda.p.y | dda.p.z) < 8u) {
        const unsigned xm = dir.x > 0 ? (0xFFu << dda.p.x) & 0xFFu : 0xFFu >> (7 - dda.p.x);
        const unsigned long long ym = dir.y > 0 ? ~0ull << (8 * dda.p.y) : ~0ull >> (8 * (7 - dda.p.y));
        const unsigned long long m2 = (0x0101010101010101ull * xm) & ym;
        const int z0 = dir.z > 0 ? dda.p.z : 0, z1 = dir.z > 0 ? 7 : dda.p.z;
        unsigned long long any = 0;
        for (int z = z0; z <= z1; z++) any |= __ldg(mk + z) & m2;
        if (any == 0) return;
    }
    // 0 <= p < res0 on every axis (res0 is a power of two) == one unsigned compare on the OR of the coordinates
    for (int iter = 0; iter < GX_MAX_ITER && unsigned(dda.p.x | dda.p.y | dda.p.z) < unsigned(res0); iter++) {
        cnt.s_pt++;
        bool solid;
        if (mk != nullptr) {
            if (dda.p.z != slice_z) { slice_z = dda.p.z; slice = __ldg(mk + slice_z); }
            solid = (slice >> (dda.p.y * 8 + dda.p.x)) & 1ull;
        } else {
            solid = smp.point(dda.p.x + o.x + .5, dda.p.y + o.y + .5, dda.p.z + o.z + .5) > P.thresh.x;
        }
        if (solid) {
            vmin += gx_f3(dda.p);
            dda.t = gx_ray_box(pos, dir, vmin, vmin + 1);
            if (dda.t.z == GX_NOHIT) {      // reference quirk: no step, vmin keeps accumulating (raycast.cuh:244-247)
                h.hit.z = GX_NOHIT;
                continue;
            }
            h.hit = pos + dda.t.x * dir;
            float3 fromVoxelCenter = (h.hit - vmin) - 0.5f;
            fromVoxelCenter -= 0.01 * dir;
            const float maxCoordinate = fmaxf(fmaxf(fabsf(fromVoxelCenter.x), fabsf(fromVoxelCenter.y)), fabsf(fromVoxelCenter.z));
            h.norm.x = (fabsf(fromVoxelCenter.x) == maxCoordinate ? copysignf(1.0f, fromVoxelCenter.x) : 0.0f);
            h.norm.y = (fabsf(fromVoxelCenter.y) == maxCoordinate ? copysignf(1.0f, fromVoxelCenter.y) : 0.0f);
            h.norm.z = (fabsf(fromVoxelCenter.z) == maxCoordinate ? copysignf(1.0f, fromVoxelCenter.z) : 0.0f);
            h.t = dda.t.x; h.leaf = nodeid; h.vox = gx_i3(vmin);
            h.cpos = gx_f3(dda.p) + o;
            return;
        }
        dda.next_step(sx, sy, sz);
    }
}

// SHADE_TRILINEAR: fixed-step march, first sample >= THRESH             cuda_gvdb_raycast.cuh:281-300
template <class S>
__device__ __forceinline__ void gx_brick_trilinear(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                   GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    t.x = P.steps.x * ceilf(t.x / P.steps.x);
    float3 p = pos + t.x * dir - vmin;

    for (int iter = 0; iter < GX_MAX_ITER && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0; iter++) {
        cnt.s_tri++;
        if (smp.tri(p.x + o.x, p.y + o.y, p.z + o.z) >= P.thresh.x) {
            h.hit = p + vmin;
            h.norm = gx_gradient(smp, p + o, cnt, false);
            h.t = t.x; h.leaf = nodeid; h.vox = gx_i3(gx_floor(h.hit));
            h.cpos = p + o;
            return;
        }
        p += P.steps.x * dir;
        t.x += P.steps.x;
    }
}

// SHADE_LEVELSET: first sample < THRESH                                 cuda_gvdb_raycast.cuh:389-410, :186-197
// (p is taken from the UNSNAPPED t.x; the fine march re-tests the same point and returns at i = 0; bounds inclusive)
template <class S>
__device__ __forceinline__ void gx_brick_levelset(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                  GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    float3 p = pos + t.x * dir - vmin;

    for (int iter = 0; iter < GX_MAX_ITER && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x <= res0 && p.y <= res0 && p.z <= res0; iter++) {
        cnt.s_tri++;
        if (smp.tri(p.x + o.x, p.y + o.y, p.z + o.z) < P.thresh.x) {
            // rayLevelSet(): the first fine sample is this very point, so it returns immediately with p unchanged
            cnt.s_tri++;
            h.hit = p + vmin;
            if (h.hit.z != GX_NOHIT) {
                h.norm = gx_gradient(smp, p + o, cnt, true);
                h.t = t.x; h.leaf = nodeid; h.vox = gx_i3(gx_floor(h.hit));
                h.cpos = p + o;
                return;
            }
        }
        p += P.steps.x * dir;
    }
}

// SHADE_VOLUME: emission / absorption through the transfer function    cuda_gvdb_raycast.cuh:485-533
template <class S>
__device__ __forceinline__ void gx_brick_deep(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                              GxHit& h, GxCount& cnt, float tDepth)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    t.x = P.steps.x * ceilf(t.x / P.steps.x);
    float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    float3 wp = pos + t.x * dir;
    float3 p = wp - vmin;
    const float3 wpt = P.steps.x * dir;
    const float dt = sqrtf(gx_dot(wpt, wpt));
    const float res0 = float(gx_res<S>(P, 0));
    float4& clr = h.clr;

    if (h.hit.x == 0) h.hit.x = t.x;

    for (int iter = 0; clr.w > P.cutoff.y && iter < GX_MAX_ITER && p.x >= 0 && p.y >= 0 && p.z >= 0
                       && p.x < res0 && p.y < res0 && p.z < res0; iter++) {
        if (t.x > tDepth) {
            float3 d = wp - pos;
            h.hit.y = sqrtf(gx_dot(d, d));
            h.hit.z = 1;
            clr = make_float4(fminf(clr.x, 1.f), fminf(clr.y, 1.f), fminf(clr.z, 1.f), fmaxf(clr.w, 0.f));
            return;
        }
        cnt.s_tri++;
        const float rawSample = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);
        if (rawSample >= P.cutoff.x) {
            cnt.s_lut++;
            float4 val = gx_transfer(P, rawSample);
            val.w = exp(P.extinct.x * val.w * P.steps.x);
            const float4 hclr = P.clr_tex ? gx_color(P, p + o) : make_float4(1, 1, 1, 1);
            // reference: clr.x += val.x * clr.w * (1 - val.w) * ALBEDO * hclr.x with hclr a run-time select, so the
            // final add pairs with "* hclr.x" (= 1): every product is rounded on its own.  Pinned with _rn intrinsics.
            const float om = 1 - val.w;
            clr.x = __fadd_rn(clr.x, __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(val.x, clr.w), om), P.extinct.y), hclr.x));
            clr.y = __fadd_rn(clr.y, __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(val.y, clr.w), om), P.extinct.y), hclr.y));
            clr.z = __fadd_rn(clr.z, __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(val.z, clr.w), om), P.extinct.y), hclr.z));
            clr.w *= val.w;
        }
        p += wpt;
        wp += wpt;
        t.x += dt;
    }
    h.hit.y = t.x;
    clr = make_float4(fminf(clr.x, 1.f), fminf(clr.y, 1.f), fminf(clr.z, 1.f), fmaxf(clr.w, 0.f));
}

// ------------------------------------------------------------------------------------------------ master ray cast
// Iterative hierarchical 3-D DDA over the compact tables.               cuda_gvdb_raycast.cuh:543-611
// The per-level stack (node id, tMax) is kept in registers: only levels 1..4 can hold state.
struct GxStackReg {            // register-resident variant (used by the A/B packet traversal)
    int   n1, n2, n3, n4;
    float m1, m2, m3, m4;
    __device__ __forceinline__ void set(int lev, int n, float m)
    {
        if (lev == 1) { n1 = n; m1 = m; } else if (lev == 2) { n2 = n; m2 = m; }
        else if (lev == 3) { n3 = n; m3 = m; } else { n4 = n; m4 = m; }
    }
    __device__ __forceinline__ int   node(int lev) const { return lev == 1 ? n1 : (lev == 2 ? n2 : (lev == 3 ? n3 : n4)); }
    __device__ __forceinline__ float tmax(int lev) const { return lev == 1 ? m1 : (lev == 2 ? m2 : (lev == 3 ? m3 : m4)); }
};
// Default: the (node, tMax) pair of every level 1..4 lives in dynamic shared memory, [level - 1][thread] — one STS / LDS
// per access, conflict-free, no select chains and eight registers fewer than the register-resident form (the reference:
// two dynamically indexed local-memory arrays).  GX_STACK_BYTES_PER_THREAD of dynamic shared memory per thread.
#define GX_STACK_BYTES_PER_THREAD 36     // 32 used by GxStack; the walker (gvdbx_walk.cuh) keeps one row of 9 words per thread
struct GxStack {
#ifdef GX_REF_LAYOUT    // launched by the reference's own RenderKernel / Render (no dynamic shared memory): room for 16 x 16 CTAs
    static __device__ __forceinline__ int* base() { __shared__ int gx_stack_static[8 * 256]; return gx_stack_static; }
#else
    static __device__ __forceinline__ int* base() { extern __shared__ int gx_stack_smem[]; return gx_stack_smem; }
#endif
    // this thread's column and the row pitch, computed ONCE per ray: a level change is then one multiply-add and two
    // shared-memory accesses (level changes happen in more than half of all DDA iterations of a sparse volume)
    int* col;
    int  nt;
    __device__ __forceinline__ GxStack() : nt(blockDim.x * blockDim.y) { col = base() + threadIdx.y * blockDim.x + threadIdx.x; }
    __device__ __forceinline__ void  set(int lev, int n, float m) const { int* p = col + (lev - 1) * nt; p[0] = n; p[4 * nt] = __float_as_int(m); }
    __device__ __forceinline__ int   node(int lev) const { return col[(lev - 1) * nt]; }
    __device__ __forceinline__ float tmax(int lev) const { return __int_as_float(col[(lev + 3) * nt]); }
};

#include "gvdbx_walk.cuh"
#ifndef GX_QK
#define GX_QK 2              // bricks queued per round by the brick-queue ray casts (gvdbx_trace.cuh)
#endif

// four-samples-per-round brick marchers (gvdbx_trace.cuh)
template <class S> __device__ __forceinline__ void gx2_brick_trilinear(const GxParams&, S&, int, float3, float3, float3, GxHit&, GxCount&);
template <class S> __device__ __forceinline__ void gx2_brick_levelset(const GxParams&, S&, int, float3, float3, float3, GxHit&, GxCount&);
template <class S> __device__ __forceinline__ void gx2_brick_deep(const GxParams&, S&, int, float3, float3, float3, GxHit&, GxCount&, float);

// brick functions of the remaining shade modes (gvdbx_extra.cuh)
template <class S> __device__ __forceinline__ void gx_brick_tricubic(const GxParams&, S&, int, float3, float3, float3, GxHit&, GxCount&);
template <class S> __device__ __forceinline__ void gx_brick_shadow(const GxParams&, S&, int, float3, float3, float3, GxHit&, GxCount&);

// lane-state-machine form of the ray cast for the fixed-step marchers (gvdbx_trace.cuh).  Bit-identical output, but
// MEASURED SLOWER than the literal nesting on every workload (cfg1 3102 vs 3862, cfg2 2099 vs 2499, cfg4 411 vs 496
// Mrays/s, profiles/r01_ncu_summary.md): compiled only with -DGX_STATE_MACHINE=1.
#ifndef GX_STATE_MACHINE
#define GX_STATE_MACHINE 0
#endif
template <int MODE, class S>
__device__ __forceinline__ void gx_raycast_sm(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt, int px, int py);
// brick-queue form of the deep ray cast (gvdbx_trace.cuh): BATCH == 2
template <class S>
__device__ __forceinline__ void gx_raycast_deep_q(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt);
template <int MODE, class S>
__device__ __forceinline__ void gx_raycast_surface_q(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt);

// rayCast on the lean walker (gvdbx_walk.cuh): the reference's nesting — the brick function runs inside the iteration that
// found the brick — with the walker's bookkeeping.  (Measured, cfg3 voxel 4K: walking to the next brick first and calling the
// brick function behind the loop makes the lanes of a warp wait for each other per brick: 4.34 vs 4.29 ms.)
template <int MODE, int BATCH, class S>
__device__ __forceinline__ bool gx_lean_brick(const GxParams& P, S& smp, int c, float3 t, float3 pos, float3 dir, GxHit& h, GxCount& cnt)
{
    if constexpr (MODE == GX_MODE_VOXEL)          gx_brick_voxel(P, smp, c, t, pos, dir, h, cnt);
    else if constexpr (MODE == GX_MODE_TRILINEAR) gx2_brick_trilinear(P, smp, c, t, pos, dir, h, cnt);
    else if constexpr (MODE == GX_MODE_LEVELSET)  gx2_brick_levelset(P, smp, c, t, pos, dir, h, cnt);
    else if constexpr (MODE == GX_MODE_DEEP)      { if (!P.clr_tex) gx2_brick_deep(P, smp, c, t, pos, dir, h, cnt, INFINITY);   // per-sample colour: literal marcher
                                                    else            gx_brick_deep(P, smp, c, t, pos, dir, h, cnt, INFINITY); }
    else if constexpr (MODE == GX_MODE_TRICUBIC)  gx_brick_tricubic(P, smp, c, t, pos, dir, h, cnt);
    else if constexpr (MODE == GX_MODE_EMPTYSKIP) h.hit = pos + t.x * dir;
    else                                          gx_brick_shadow(P, smp, c, t, pos, dir, h, cnt);
    // rayCast's tests behind the brick call (:584-590); true = the ray is finished
    if (h.clr.w <= 0) { h.clr.w = 0; return true; }
    if (h.hit.z != GX_NOHIT) return true;
    if (MODE == GX_MODE_DEEP && h.clr.w <= P.cutoff.y) return true;
    return false;
}
template <int MODE, int BATCH, class S>
__device__ __forceinline__ void gx_raycast_lean(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt)
{
    typedef GxWalk<S, (BATCH == 2 ? GX_WALK_WORDS + 2 * GX_QK : GX_WALK_WORDS)> Walk;      // queue kernels: the row also holds the brick queue
    Walk w;
    if (!w.start(P, pos, dir, cnt)) return;
    w.walk(P, cnt, [&](int leaf, float t_enter, float t_exit) {
        return gx_lean_brick<MODE, BATCH>(P, smp, leaf, make_float3(t_enter, t_exit, 0.f), pos, dir, h, cnt);
    });
}

template <int MODE, int BATCH, class S>
__device__ __forceinline__ void gx_raycast(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt,
                                           int px, int py)
{
    if constexpr (GX_STATE_MACHINE && BATCH && (MODE == GX_MODE_TRILINEAR || MODE == GX_MODE_LEVELSET || MODE == GX_MODE_DEEP)) {
        if (MODE != GX_MODE_DEEP || (P.dbuf == nullptr && !P.clr_tex)) { gx_raycast_sm<MODE>(P, smp, pos, dir, h, cnt, px, py); return; }
    }
#ifndef GX_REF_LAYOUT
    // Every ray cast of a kernel that is not one of the A/B baselines runs on the walker (gvdbx_walk.cuh) — it and GxStack lay
    // the shared-memory stack out differently, so the choice must be uniform over the block: it depends on the kernel variant
    // and on whether the frame binds a depth buffer (the depth clip lives in the literal loop below).
    if constexpr (BATCH >= 1) {
        if (P.dbuf == nullptr) {
            if constexpr (BATCH == 2 && MODE == GX_MODE_DEEP) {
                if (!P.clr_tex) { gx_raycast_deep_q(P, smp, pos, dir, h, cnt); return; }     // per-sample colour: one brick at a time
            }
            if constexpr (BATCH == 2 && (MODE == GX_MODE_TRILINEAR || MODE == GX_MODE_LEVELSET)) {
                gx_raycast_surface_q<MODE>(P, smp, pos, dir, h, cnt); return;
            }
            gx_raycast_lean<MODE, BATCH>(P, smp, pos, dir, h, cnt);
            return;
        }
    }
#endif
    GxStack st;
    int lev = P.top_lev;
    cnt.rays++;
    float3 tStart = gx_ray_box(pos, dir, P.bmin, P.bmax);
    if (tStart.z == GX_NOHIT) return;
    if (lev < 1 || lev >= GX_MAXLEV) return;        // single-brick volume: the reference loop never runs either
    gx_npos_t np = gx_node_pos(P, lev, 0);
    cnt.n_desc++;
    float3 vmin = make_float3(float(np.x), float(np.y), float(np.z));

    tStart.x += P.epsilon;
    st.set(lev, 0, tStart.y - P.epsilon);
    // the state of the CURRENT level is mirrored in plain registers so that the per-step code has no select chains:
    float      cur_tmax = tStart.y - P.epsilon;       // exit parameter of the current node
    gx_ctab_t  ctab = gx_table(P, lev, 0, gx_dim<S>(P, lev));   // child table of the current node (node 0 of the top level)
    unsigned   res = unsigned(gx_res<S>(P, lev));

    GxDDA dda;
    dda.set_ray(pos, dir, tStart);
    dda.prepare(vmin, gx_vdel<S>(P, lev));
    const float tDepth = gx_depth_max(P, dir, px, py);

    // Level changes.  A descent (:592-597) and every pop of the ascent loop (:603-608) end with HDDAState::Prepare, a pure
    // function of (ray, t.x, node corner, cell size): when several happen in one iteration only the LAST one survives, so
    // the level bookkeeping runs where the reference has it and ONE Prepare runs at the end of the iteration — at a single
    // code site, where the lanes that descended and the lanes that ascended execute it together.
    int node = 0;
    bool moved = false;

    // loop guard of the reference (cuda_gvdb_raycast.cuh:567): 0 <= p <= res on every axis == unsigned(p) <= res
    for (int iter = 0; iter < GX_MAX_ITER && lev > 0 && lev <= P.top_lev
                       && unsigned(dda.p.x) <= res && unsigned(dda.p.y) <= res && unsigned(dda.p.z) <= res; iter++) {
        dda.next();
        if (dda.t.x > tDepth) { h.hit.z = 0; return; }

        const int dm = gx_dim<S>(P, lev);
        const int b = (((int(dda.p.z) << dm) + int(dda.p.y)) << dm) + int(dda.p.x);
        // cells with a coordinate == res can only be reached through the reference's inclusive loop bound; they hold no
        // child.  res is a power of two, so "all three coordinates < res" is one compare on their OR.
        int c = -1;
        if (unsigned(dda.p.x | dda.p.y | dda.p.z) < res) c = gx_child(ctab, b);
        cnt.n_dda++;
        if (c != -1) {
            if (lev == 1) {
                dda.t.x += P.epsilon;
                if constexpr (MODE == GX_MODE_VOXEL)          gx_brick_voxel(P, smp, c, dda.t, pos, dir, h, cnt);
                else if constexpr (MODE == GX_MODE_TRILINEAR) { if (BATCH) gx2_brick_trilinear(P, smp, c, dda.t, pos, dir, h, cnt);
                                                      else       gx_brick_trilinear(P, smp, c, dda.t, pos, dir, h, cnt); }
                else if constexpr (MODE == GX_MODE_LEVELSET)  { if (BATCH) gx2_brick_levelset(P, smp, c, dda.t, pos, dir, h, cnt);
                                                      else       gx_brick_levelset(P, smp, c, dda.t, pos, dir, h, cnt); }
                else if constexpr (MODE == GX_MODE_DEEP)      { if (BATCH && !P.clr_tex) gx2_brick_deep(P, smp, c, dda.t, pos, dir, h, cnt, tDepth);   // per-sample colour: literal marcher
                                                      else       gx_brick_deep(P, smp, c, dda.t, pos, dir, h, cnt, tDepth); }
                else if constexpr (MODE == GX_MODE_TRICUBIC)  gx_brick_tricubic(P, smp, c, dda.t, pos, dir, h, cnt);
                else if constexpr (MODE == GX_MODE_EMPTYSKIP) h.hit = pos + dda.t.x * dir;       // rayEmptySkipBrick, cuda_gvdb_raycast.cuh:425-428
                else                                gx_brick_shadow(P, smp, c, dda.t, pos, dir, h, cnt);
                if (h.clr.w <= 0) { h.clr.w = 0; return; }
                if (h.hit.z != GX_NOHIT) return;
                // deep mode: once transmittance is at or below ALPHACUT no later brick can change the colour
                // (their sample loops do not run and the exit clamp is idempotent), so stop here.
                if (MODE == GX_MODE_DEEP && h.clr.w <= P.cutoff.y) return;
                dda.step();
            } else {
                lev--;
                cnt.n_desc++;
                dda.t.x += P.epsilon;
                cur_tmax = dda.t.y - P.epsilon;
                st.set(lev, c, cur_tmax);
                node = c; moved = true;
            }
        } else {
            dda.step();
        }
        while (dda.t.x > cur_tmax && lev <= P.top_lev) {
            lev++;
            if (lev <= P.top_lev) {
                node = st.node(lev);
                cur_tmax = st.tmax(lev);
                cnt.n_desc++;
                moved = true;
            }
        }
        if (moved && lev <= P.top_lev) {
            moved = false;
            ctab = gx_table(P, lev, node, gx_dim<S>(P, lev));
            res = unsigned(gx_res<S>(P, lev));
            np = gx_node_pos(P, lev, node);
            vmin = make_float3(float(np.x), float(np.y), float(np.z));
            dda.prepare(vmin, gx_vdel<S>(P, lev));
        }
    }
}

// hclr of the surface brick functions (`if (gvdb->clr_chan != CHAN_UNDEF) hclr = getColorF(gvdb, gvdb->clr_chan, p + o)`,
// cuda_gvdb_raycast.cuh:259, :294, :333, :403), applied after the ray cast: a hit ends the ray, so nothing reads hclr in
// between except rayCast's `clr.w <= 0` test, which can only confirm the return the hit causes anyway
__device__ __forceinline__ void gx_hit_color(const GxParams& P, GxHit& h)
{
    if (P.clr_tex && h.hit.z != GX_NOHIT) h.clr = gx_color(P, h.cpos);
}

// ------------------------------------------------------------------------------------------------ shading
// Phong + optional shadow ray with the same brick function              cuda_gvdb_module.cu:38-57
template <int MODE, int BATCH, class S>
__device__ __forceinline__ float4 gx_phong(const GxParams& P, S& smp, float3 shit, float3 snorm, float4 sclr, GxCount& cnt,
                                           int px, int py)
{
    if (shit.z == GX_NOHIT) return P.backclr;
    float3 lightdir = gx_normalize(P.light_pos - shit);
    float diff = 0.9 * fmaxf(0.0f, gx_dot(snorm, lightdir));
    float amb = 0.1f;
    if (P.shadow_params.x > 0) {
        GxHit h2;
        h2.hit = make_float3(0, 0, GX_NOHIT);
        h2.clr = make_float4(0, 0, 0, 1);
        h2.norm = make_float3(0, 0, 0); h2.t = 0; h2.leaf = -1; h2.vox = make_int3(0, 0, 0);
        gx_raycast<MODE, BATCH>(P, smp, shit + snorm * P.shadow_params.y, lightdir, h2, cnt, px, py);
        diff = (h2.hit.z == GX_NOHIT ? diff : diff * (1.0 - P.shadow_params.x));
    }
    return make_float4(sclr.x * (diff + amb), sclr.y * (diff + amb), sclr.z * (diff + amb), 1.0);
}

// ------------------------------------------------------------------------------------------------ kernels
#define GX_FLAG_DEBUG   1
#define GX_FLAG_COUNT   2
#define GX_FLAG_TILES   4
#define GX_FLAG_LITERAL 8      // reference-shaped loops: literal nesting, one sample at a time (A/B baseline)
#define GX_FLAG_PACKET  16     // vote-converged two-phase traversal of gvdbx_trace.cuh (A/B)
#define GX_FLAG_QUEUE   64     // deep modes: brick-queue traversal (gx_raycast_deep_q, gvdbx_trace.cuh)
// default (neither flag): literal nesting of traversal and brick visit, four-samples-per-round brick marchers

template <int MODE, class S>
__device__ __forceinline__ float4 gx2_trace_pixel(const GxParams& P, S& smp, float3 rpos, float3 rdir, int px, int py,
                                                  GxCount& cnt, GxHit& prim, float4& raw, bool valid);

// section / deep-shadow pixel functions (gvdbx_extra.cuh)
template <class S> __device__ __forceinline__ float4 gx_pixel_section2d(const GxParams&, S&, int, int, GxCount&);
template <int BATCH, class S> __device__ __forceinline__ float4 gx_pixel_section3d(const GxParams&, S&, float3, float3, int, int, GxCount&, GxHit&);
template <int BATCH, class S> __device__ __forceinline__ float4 gx_pixel_deepshadow(const GxParams&, S&, float3, float3, int, int, GxCount&, GxHit&, float4&);

// One camera ray through pixel (x,y) at sub-pixel offset (ox,oy) — the reference kernels use (0.5, 0.5) — shaded to the
// float colour the reference packs with make_uchar4(clr*255).           cuda_gvdb_module.cu:60-181, :184-207, :225-298
template <int MODE, int BATCH, class S>
__device__ __forceinline__ float4 gx_shade_pixel(const GxParams& P, S& smp, int x, int y, float ox, float oy,
                                                 GxCount& cnt, GxHit& h, float4& raw)
{
    if constexpr (MODE == GX_MODE_SECTION2D) { return gx_pixel_section2d(P, smp, x, y, cnt); } else {
    // cuda_gvdb_geom.cuh:48-63
    float3 rpos = gx_mmult(P.invxform, P.campos);
    float u = float(x + ox) / float(P.width), v = float(y + oy) / float(P.height);
    float3 vv = u * P.camu + v * P.camv + P.cams;
    float3 rdir = gx_normalize(gx_mmult(P.invxrot, vv));
    float4 clr;
    if constexpr (MODE == GX_MODE_DEEP) {
        h.clr = make_float4(0, 0, 0, 1);
        h.hit = make_float3(0, 0, GX_NOHIT);
        gx_raycast<GX_MODE_DEEP, BATCH>(P, smp, rpos, rdir, h, cnt, x, y);
        raw = h.clr;
        clr = h.clr;
        float a = 1.0 - clr.w;
        clr = make_float4(P.backclr.x + a * (clr.x - P.backclr.x), P.backclr.y + a * (clr.y - P.backclr.y),
                          P.backclr.z + a * (clr.z - P.backclr.z), 1.0 - clr.w);
    } else if constexpr (MODE == GX_MODE_DEEPSHADOW) {
        clr = gx_pixel_deepshadow<BATCH>(P, smp, rpos, rdir, x, y, cnt, h, raw);
    } else if constexpr (MODE == GX_MODE_SECTION3D) {
        clr = gx_pixel_section3d<BATCH>(P, smp, rpos, rdir, x, y, cnt, h);
    } else if constexpr (MODE == GX_MODE_EMPTYSKIP) {                               // cuda_gvdb_module.cu:184-207
        h.clr = make_float4(1, 1, 1, 1);
        h.hit = make_float3(GX_NOHIT, GX_NOHIT, GX_NOHIT);
        gx_raycast<GX_MODE_EMPTYSKIP, BATCH>(P, smp, rpos, rdir, h, cnt, x, y);
        if (h.hit.z != GX_NOHIT) { const float3 c = h.hit * 0.01; clr = make_float4(c.x, c.y, c.z, 1); }
        else clr = P.backclr;
        clr.w = 1.0f;                                                     // the kernel packs alpha as the literal 255
    } else {
        h.clr = make_float4(1, 1, 1, 1);
        h.hit = (MODE == GX_MODE_LEVELSET) ? make_float3(0, 0, GX_NOHIT) : make_float3(GX_NOHIT, GX_NOHIT, GX_NOHIT);
        gx_raycast<MODE, BATCH>(P, smp, rpos, rdir, h, cnt, x, y);
        // colour channel: the hit voxel's colour is fetched now but kept as its four BYTES (one register) while the shadow
        // ray runs; performPhongShading's sclr.rgb * (diff + amb) is then (diff + amb) — what the call returns for the
        // constant colour 1 — times the colour: the same single rounding per channel
        const bool tint = P.clr_tex && h.hit.z != GX_NOHIT;
        uchar4 cb = make_uchar4(255, 255, 255, 255);
        if (tint) cb = tex3D<uchar4>(P.clr_tex, (int)h.cpos.x, (int)h.cpos.y, (int)h.cpos.z);
        // the tricubic kernel shades its shadow ray with the trilinear brick function (cuda_gvdb_module.cu:136)
        clr = gx_phong<(MODE == GX_MODE_TRICUBIC ? GX_MODE_TRILINEAR : MODE), BATCH>(P, smp, h.hit, h.norm, h.clr, cnt, x, y);
        if (tint) {
            clr.x = (float(cb.x) / 255.0f) * clr.x; clr.y = (float(cb.y) / 255.0f) * clr.y; clr.z = (float(cb.z) / 255.0f) * clr.z;
        }
    }
    return clr;
    }
}

#define GX_FLAG_SPP     32     // P.spp sub-pixel samples per pixel, averaged in float before packing

// Launch bounds.  CTAs are at most GX_MAXTHREADS = 128 threads (default 8x8 = two 8x4-pixel warps); the register budget follows
// from the number of such CTAs per SM: 8 -> 64 registers (surface modes, SHADE_VOXEL: issue-bound, occupancy pays),
// the deep brick-queue kernels included since the walker gives six registers back while the bricks are marched (GxWalk::resume):
// measured cfg4 deep + shadow 4K 20.17 ms at 72 registers / 28 warps per SM, 19.43 ms at 64 registers / 32 warps (deep: 18.93 -> 18.66);
// before that step 64 registers spilled (23.1 ms against 19.7 at 72 and 20.1 at 80).
template <int MODE, int SAMPLER, int FLAGS, bool UNI>
#ifndef GX_MAXTHREADS
#define GX_MAXTHREADS 128
#endif
#ifndef GX_MINBLOCKS
#define GX_MINBLOCKS 8
#endif
#ifndef GX_QUEUE_MINBLOCKS
#define GX_QUEUE_MINBLOCKS 8
#endif
#ifndef GX_SURFQ_MINBLOCKS
#define GX_SURFQ_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(GX_MAXTHREADS, !(FLAGS & GX_FLAG_QUEUE) ? GX_MINBLOCKS : ((MODE == GX_MODE_DEEP || MODE == GX_MODE_DEEPSHADOW) ? GX_QUEUE_MINBLOCKS : GX_SURFQ_MINBLOCKS))
gx_render_kernel(const __grid_constant__ GxParams P)
{
    int x, y;
    size_t opix;
    bool valid;
    if (FLAGS & GX_FLAG_TILES) {
        // blockIdx.y = tile slot of this rank, blockIdx.x = sub-block inside the tile
        const int tw = P.tile_w, th = P.tile_h;
        const int sub_x = tw / blockDim.x;
        const int tile = blockIdx.y * P.nranks + P.rank;
        if (tile >= P.ntiles) return;
        const int lx = (blockIdx.x % sub_x) * blockDim.x + threadIdx.x;
        const int ly = (blockIdx.x / sub_x) * blockDim.y + threadIdx.y;
        x = (tile % P.tiles_x) * tw + lx;
        y = (tile / P.tiles_x) * th + ly;
        // direct mode (out_stride > 0): pixels go straight into a row-major frame (possibly a peer GPU's, over NVLink);
        // packed mode: tile after tile into this rank's own buffer
        opix = P.out_stride > 0 ? size_t(y) * P.out_stride + x : (size_t(blockIdx.y) * th + ly) * tw + lx;
        valid = (x < P.width && y < P.height);
    } else {
        x = P.x0 + blockIdx.x * blockDim.x + threadIdx.x;
        y = P.y0 + blockIdx.y * blockDim.y + threadIdx.y;
        valid = (x < P.x1 && y < P.y1);
        opix = size_t(y) * P.out_stride + x;
    }
    if (!(FLAGS & GX_FLAG_PACKET) && !valid) return;     // the packet traversal keeps whole warps alive for its votes

    GxSampler<SAMPLER, UNI> smp(P);
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    GxHit h;
    h.norm = make_float3(0, 0, 0); h.t = 0; h.leaf = -1; h.vox = make_int3(0, 0, 0);
    h.hit = make_float3(0, 0, GX_NOHIT); h.clr = make_float4(0, 0, 0, 0);

    float4 clr;
    float4 raw = make_float4(0, 0, 0, 0);
    constexpr int BATCH = (FLAGS & GX_FLAG_LITERAL) ? 0 : ((FLAGS & GX_FLAG_QUEUE) ? 2 : 1);   // 0 literal loops, 1 four-sample rounds, 2 + brick queue
    if constexpr ((FLAGS & GX_FLAG_PACKET) != 0) {
        float3 rpos = gx_mmult(P.invxform, P.campos);
        float u = float(x + 0.5f) / float(P.width), v = float(y + 0.5f) / float(P.height);
        float3 vv = u * P.camu + v * P.camv + P.cams;
        float3 rdir = gx_normalize(gx_mmult(P.invxrot, vv));
        clr = gx2_trace_pixel<MODE>(P, smp, rpos, rdir, x, y, cnt, h, raw, valid);
        if (!valid) return;
    } else if constexpr ((FLAGS & GX_FLAG_SPP) != 0) {
        // sample s sits at ((s % g) + 0.5) / g, ((s / g) + 0.5) / g inside the pixel; colours are summed in sample order
        float4 sum = make_float4(0, 0, 0, 0);
        for (int s = 0; s < P.spp; s++) {
            const float ox = __fmul_rn(float(s % P.spp_grid) + 0.5f, P.spp_inv_grid);
            const float oy = __fmul_rn(float(s / P.spp_grid) + 0.5f, P.spp_inv_grid);
            const float4 c = gx_shade_pixel<MODE, BATCH>(P, smp, x, y, ox, oy, cnt, h, raw);
            sum.x = __fadd_rn(sum.x, c.x); sum.y = __fadd_rn(sum.y, c.y); sum.z = __fadd_rn(sum.z, c.z); sum.w = __fadd_rn(sum.w, c.w);
        }
        clr = make_float4(__fmul_rn(sum.x, P.spp_inv), __fmul_rn(sum.y, P.spp_inv), __fmul_rn(sum.z, P.spp_inv), __fmul_rn(sum.w, P.spp_inv));
    } else {
        clr = gx_shade_pixel<MODE, BATCH>(P, smp, x, y, 0.5f, 0.5f, cnt, h, raw);
    }
    if (MODE == GX_MODE_EMPTYSKIP || MODE == GX_MODE_SECTION2D || MODE == GX_MODE_SECTION3D)
        P.out[opix] = make_uchar4(clr.x * 255, clr.y * 255, clr.z * 255, 255);
    else
        P.out[opix] = make_uchar4(clr.x * 255, clr.y * 255, clr.z * 255, clr.w * 255);

    if (FLAGS & GX_FLAG_DEBUG) {
        float4* d = P.dbg + 3 * (size_t(y) * P.width + x);
        if (MODE == GX_MODE_DEEP || MODE == GX_MODE_DEEPSHADOW) {
            d[0] = raw;
            d[1] = make_float4(h.hit.x, h.hit.y, h.hit.z, 0.f);
            d[2] = make_float4(0, 0, 0, 0);
        } else if (MODE == GX_MODE_SECTION2D) {
            d[0] = clr; d[1] = make_float4(0, 0, 0, 0); d[2] = make_float4(0, 0, 0, 0);
        } else {
            bool miss = (h.hit.z == GX_NOHIT);
            d[0] = make_float4(h.hit.x, h.hit.y, h.hit.z, miss ? 0.f : h.t);
            d[1] = miss ? make_float4(0, 0, 0, __int_as_float(-1)) : make_float4(h.norm.x, h.norm.y, h.norm.z, __int_as_float(h.leaf));
            d[2] = miss ? make_float4(0, 0, 0, 0)
                        : make_float4(__int_as_float(h.vox.x), __int_as_float(h.vox.y), __int_as_float(h.vox.z), 0.f);
        }
    }
    if (FLAGS & GX_FLAG_COUNT) {
        unsigned int v6[6] = {cnt.s_tri, cnt.s_pt, cnt.n_dda, cnt.n_desc, cnt.s_lut, cnt.rays};
        #pragma unroll
        for (int i = 0; i < 6; i++) atomicAdd(&P.counters[i], (unsigned long long)v6[i]);
    }
}

// ------------------------------------------------------------------------------------------------ explicit ray bundles
// gvdbRaytrace (cuda_gvdb_module.cu:211-222): one thread per 64-byte ScnRay record {hit@0, normal@12, orig@24, dir@36,
// clr@48, pnode@52, pndx@56}; surface hit with the trilinear brick function, hit pulled back by `bias` along the ray.
template <int SAMPLER, bool UNI>
__global__ void __launch_bounds__(64) gx_raytrace_kernel(const __grid_constant__ GxParams P, float* __restrict__ rays, int num_rays, float bias)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_rays) return;
    float* r = rays + 16 * size_t(i);
    GxSampler<SAMPLER, UNI> smp(P);
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    GxHit h;
    h.hit = make_float3(GX_NOHIT, GX_NOHIT, GX_NOHIT);
    h.norm = make_float3(r[3], r[4], r[5]);                 // left as it was when nothing is hit
    h.clr = make_float4(1, 1, 1, 1);
    h.t = 0; h.leaf = -1; h.vox = make_int3(0, 0, 0);
    const float3 orig = make_float3(r[6], r[7], r[8]), dir = make_float3(r[9], r[10], r[11]);
    gx_raycast<GX_MODE_TRILINEAR, true>(P, smp, orig, dir, h, cnt, 0, 0);
    if (h.hit.z != GX_NOHIT) h.hit -= dir * bias;
    r[0] = h.hit.x; r[1] = h.hit.y; r[2] = h.hit.z;
    r[3] = h.norm.x; r[4] = h.norm.y; r[5] = h.norm.z;
}

