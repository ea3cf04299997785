// oracle_kernels.cu — TEST INFRASTRUCTURE (oracle/_ref), not product.
//
// Thin __global__ wrappers around the reference's OWN, UNMODIFIED device functions, launched through the
// reference's plugin point VolumeGVDB::RenderKernel (gvdb_volume_gvdb.cpp:4309-4333), exactly like the
// reference sample source/gRenderKernel/render_custom.cu:17-23 includes the headers.  The reference headers are
// compiled in place from /root/reference (-I path); nothing is copied.  The wrappers only expose values the
// native kernels keep internal (hit point, normal, raw deep colour) so parity can be checked on them bit-exactly.
//
// Output layout for the *Hit kernels: 32 B per pixel = float4{hit.xyz, 0} float4{norm.xyz, 0}
// (render buffer created with AddRenderBuf(chan, w, h, 32)).
#include <stdio.h>
#include "cuda_math.cuh"

#define CUDA_PATHWAY
#include "cuda_gvdb_scene.cuh"
#include "cuda_gvdb_nodes.cuh"
#include "cuda_gvdb_geom.cuh"
#include "cuda_gvdb_dda.cuh"
#include "cuda_gvdb_raycast.cuh"

#define ORACLE_PIXEL()                                                            \
    int x = blockIdx.x * blockDim.x + threadIdx.x;                                \
    int y = blockIdx.y * blockDim.y + threadIdx.y;                                \
    if (x >= scn.width || y >= scn.height) return;                                \
    float3 rpos = getViewPos();                                                   \
    float3 rdir = getViewRay(float(x + 0.5f) / float(scn.width), float(y + 0.5f) / float(scn.height));

static __device__ void oracleStoreHit(uchar4* outBuf, int x, int y, float3 hit, float3 norm)
{
    float4* o = (float4*)outBuf + 2 * (y * scn.width + x);
    bool miss = (hit.z == NOHIT);
    o[0] = make_float4(hit.x, hit.y, hit.z, 0.f);
    o[1] = miss ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(norm.x, norm.y, norm.z, 0.f);
}

// primary-ray hit point + normal, SHADE_VOXEL brick function (module.cu:83-96 without shading)
extern "C" __global__ void oracleHitVoxel(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float3 hit = make_float3(NOHIT, NOHIT, NOHIT), norm = make_float3(0, 0, 0);
    float4 clr = make_float4(1, 1, 1, 1);
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, raySurfaceVoxelBrick);
    oracleStoreHit(outBuf, x, y, hit, norm);
}
// SHADE_TRILINEAR brick function (module.cu:103-116)
extern "C" __global__ void oracleHitTrilinear(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float3 hit = make_float3(NOHIT, NOHIT, NOHIT), norm = make_float3(0, 0, 0);
    float4 clr = make_float4(1, 1, 1, 1);
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, raySurfaceTrilinearBrick);
    oracleStoreHit(outBuf, x, y, hit, norm);
}
// SHADE_LEVELSET brick function (module.cu:164-177)
extern "C" __global__ void oracleHitLevelSet(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float3 hit = make_float3(0, 0, NOHIT), norm = make_float3(0, 0, 0);
    float4 clr = make_float4(1, 1, 1, 1);
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, rayLevelSetBrick);
    oracleStoreHit(outBuf, x, y, hit, norm);
}
// raw (unpacked) deep colour: float4{clr} float4{hit.x, hit.y, hit.z, 0} (module.cu:60-74 before the composite)
extern "C" __global__ void oracleDeepRaw(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float4 clr = make_float4(0, 0, 0, 1);
    float3 hit = make_float3(0, 0, NOHIT), norm;
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, rayDeepBrick);
    float4* o = (float4*)outBuf + 2 * (y * scn.width + x);
    o[0] = clr;
    o[1] = make_float4(hit.x, hit.y, hit.z, 0.f);
}
// BASELINE.json config 4: deep + shadow ray.  Not a native reference path (rayShadowBrick is defined at
// cuda_gvdb_raycast.cuh:445 but never called); composed here only from reference device functions:
// primary deep march, then from the first sample position a shadow march towards the light that attenuates rgb.
extern "C" __global__ void oracleDeepShadow(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float4 clr = make_float4(0, 0, 0, 1);
    float3 hit = make_float3(0, 0, NOHIT), norm;
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, rayDeepBrick);
    if (hit.x != 0.f) {
        float3 spos = rpos + rdir * hit.x;
        float3 ldir = normalize(scn.light_pos - spos);
        float4 sclr = make_float4(0, 0, 0, 0);
        float3 shit = make_float3(0, 0, NOHIT), snorm;
        rayCast(gvdb, chan, spos, ldir, shit, snorm, sclr, rayShadowBrick);
        float lit = 1.0f - sclr.w;
        clr.x *= lit; clr.y *= lit; clr.z *= lit;
    }
    clr = make_float4(lerp3(SCN_BACKCLR, clr, 1.0 - clr.w), 1.0 - clr.w);
    outBuf[y * scn.width + x] = make_uchar4(clr.x * 255, clr.y * 255, clr.z * 255, clr.w * 255);
}

// SHADE_TRICUBIC brick function (module.cu:122-139): primary hit point + normal
extern "C" __global__ void oracleHitTricubic(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float3 hit = make_float3(NOHIT, NOHIT, NOHIT), norm = make_float3(0, 0, 0);
    float4 clr = make_float4(1, 1, 1, 1);
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, raySurfaceTricubicBrick);
    oracleStoreHit(outBuf, x, y, hit, norm);
}
// SHADE_EMPTYSKIP brick function (module.cu:184-207): first brick entry point
extern "C" __global__ void oracleHitEmptySkip(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float3 hit = make_float3(NOHIT, NOHIT, NOHIT), norm = make_float3(0, 0, 0);
    float4 clr = make_float4(1, 1, 1, 1);
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, rayEmptySkipBrick);
    oracleStoreHit(outBuf, x, y, hit, make_float3(0, 0, 0));
}
// BASELINE.json config 5: deep render with several rays per pixel.  One launch per sample: scn.frame = samples per
// pixel n, scn.samples = index s of this launch (the two ScnInfo fields PrepareRender fills from Scene::SetFrame /
// SetSample and no native kernel reads).  Sample s sits at ((s % g) + .5) / g, ((s / g) + .5) / g with g = ceil(sqrt(n))
// instead of the native kernels' (.5, .5); everything else is gvdbRayDeep (module.cu:60-78).  The composited FLOAT
// colour goes to a 16 B/pixel render buffer; the harness sums the samples in order, scales by 1/n and packs.
extern "C" __global__ void oracleDeepSample(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= scn.width || y >= scn.height) return;
    int n = scn.frame < 1 ? 1 : scn.frame, s = scn.samples, g = 1;
    while (g * g < n) g++;
    const float inv_g = __fdiv_rn(1.0f, float(g));
    const float ox = __fmul_rn(float(s % g) + 0.5f, inv_g), oy = __fmul_rn(float(s / g) + 0.5f, inv_g);
    float3 rpos = getViewPos();
    float3 rdir = getViewRay(float(x + ox) / float(scn.width), float(y + oy) / float(scn.height));
    float4 clr = make_float4(0, 0, 0, 1);
    float3 hit = make_float3(0, 0, NOHIT), norm;
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, rayDeepBrick);
    clr = make_float4(lerp3(SCN_BACKCLR, clr, 1.0 - clr.w), 1.0 - clr.w);
    ((float4*)outBuf)[y * scn.width + x] = clr;
}

// voxel-hit ID and depth.  The native SHADE_VOXEL kernel keeps both internal (it only emits RGBA8), so this
// is raySurfaceVoxelBrick (cuda_gvdb_raycast.cuh:227-265) restated statement by statement with two extra outputs smuggled
// through the brick-function signature: norm <- bit patterns of int3(vmin) (the index-space voxel) and hclr <- {dda.t.x (the
// depth), bit pattern of the leaf id, 0, 1}.  The traversal around it is the reference's own rayCast.
// Output: 32 B per pixel = float4{hit.xyz, t} int4{voxel.xyz, leaf}; a miss writes {hit, 0} {0, 0, 0, -1}.
__device__ void oracleVoxelIdBrick(VDBInfo* gvdb, uchar chan, int nodeid, float3 t, float3 pos, float3 dir, float3& hit, float3& norm, float4& hclr)
{
    float3 vmin;
    VDBNode* node = getNode(gvdb, 0, nodeid, &vmin);
    float3 o = make_float3(node->mValue);
    HDDAState dda;
    dda.SetFromRay(pos, dir, t);
    dda.PrepareLeaf(vmin);
    for (int iter = 0; iter < MAX_ITER && dda.p.x >= 0 && dda.p.y >= 0 && dda.p.z >= 0
                       && dda.p.x < gvdb->res[0] && dda.p.y < gvdb->res[0] && dda.p.z < gvdb->res[0]; iter++) {
        if (tex3D<float>(gvdb->volIn[chan], dda.p.x + o.x + .5, dda.p.y + o.y + .5, dda.p.z + o.z + .5) > SCN_THRESH) {
            vmin += make_float3(dda.p);
            dda.t = rayBoxIntersect(pos, dir, vmin, vmin + 1);
            if (dda.t.z == NOHIT) { hit.z = NOHIT; continue; }
            hit = getRayPoint(pos, dir, dda.t.x);
            int3 v = make_int3(vmin);
            norm = make_float3(__int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z));
            hclr = make_float4(dda.t.x, __int_as_float(nodeid), 0.f, 1.f);
            return;
        }
        dda.Next();
        dda.Step();
    }
}
extern "C" __global__ void oracleVoxelId(VDBInfo* gvdb, uchar chan, uchar4* outBuf)
{
    ORACLE_PIXEL();
    float3 hit = make_float3(NOHIT, NOHIT, NOHIT), norm = make_float3(0, 0, 0);
    float4 clr = make_float4(1, 1, 1, 1);
    rayCast(gvdb, chan, rpos, rdir, hit, norm, clr, oracleVoxelIdBrick);
    float4* o = (float4*)outBuf + 2 * (y * scn.width + x);
    const bool miss = (hit.z == NOHIT);
    o[0] = make_float4(hit.x, hit.y, hit.z, miss ? 0.f : clr.x);
    o[1] = miss ? make_float4(0.f, 0.f, 0.f, __int_as_float(-1)) : make_float4(norm.x, norm.y, norm.z, clr.y);
}
