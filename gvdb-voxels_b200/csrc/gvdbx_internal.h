// gvdbx_internal.h — the context object behind gvdbx_t and the CUDA-context guard, shared by the translation units that
// implement the C ABI (gvdbx_api.cu: import / render; gvdbx_multi.cu: multi-GPU rings).  Not part of the public interface.
#pragma once
#include "../../include/gvdbx.h"
#include "gvdbx_device.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <dlfcn.h>

// ------------------------------------------------------------------------------------------------ CUDA context
// The reference lives in a context it creates itself (StartCuda: cuCtxCreate, gvdb_allocator.cpp:1105-1109) and brackets
// every entry point with cuCtxPushCurrent / cuCtxPopCurrent (gvdb_volume_gvdb.cpp:41-42).  Its pools, CUarrays and render
// buffers belong to THAT context, so this library must run in it too: gvdbx_create adopts the context that is current on
// the calling thread (only when none is current does it bind the device's primary context, as a stand-alone CUDA-runtime
// host expects), and every entry point makes the adopted context current for its own duration — the same push / pop
// discipline as the reference.  The four driver entry points are taken from libcuda at run time (no link dependency:
// the library still loads, and fails with GVDBX_E_CUDA, on a machine without a driver).
typedef int (*gx_cuCtxGetCurrent_t)(void**);
typedef int (*gx_cuCtxPushCurrent_t)(void*);
typedef int (*gx_cuCtxPopCurrent_t)(void**);
typedef int (*gx_cuCtxGetDevice_t)(int*);
struct GxDriver {
    gx_cuCtxGetCurrent_t  get = nullptr;
    gx_cuCtxPushCurrent_t push = nullptr;
    gx_cuCtxPopCurrent_t  pop = nullptr;
    gx_cuCtxGetDevice_t   dev = nullptr;
    bool ok = false, looked = false;
    bool load()
    {
        if (looked) return ok;
        looked = true;
        void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        get = (gx_cuCtxGetCurrent_t)dlsym(lib, "cuCtxGetCurrent");
        push = (gx_cuCtxPushCurrent_t)dlsym(lib, "cuCtxPushCurrent_v2");
        pop = (gx_cuCtxPopCurrent_t)dlsym(lib, "cuCtxPopCurrent_v2");
        dev = (gx_cuCtxGetDevice_t)dlsym(lib, "cuCtxGetDevice");
        ok = get && push && pop && dev;
        return ok;
    }
};
inline GxDriver gx_drv;

#define GX_MAX_BANDS 16
struct gvdbx_ctx {
    int          device = 0;
    void*        cuctx = nullptr;       // the CUcontext every entry point runs in (adopted at creation)
    cudaStream_t stream = nullptr;
    std::string  err;
    // options
    // banded render + overlapped read-back (gvdbx_render_banded / gvdbx_read_banded)
    cudaStream_t band_streams[2] = {nullptr, nullptr}, band_copy = nullptr;
    cudaEvent_t  band_join[2] = {nullptr, nullptr}, band_fork = nullptr, band_ev[GX_MAX_BANDS] = {};
    int          band_n = 0, band_rows = 0, band_w = 0, band_h = 0;
    uint64_t     band_buf = 0;
    int sampler = GX_SAMPLER_TEX, block_w = 8, block_h = 8, count = 0, literal = 0, spp = 1, deep_shadow = 0, memops = 0;
    // topology
    bool       have_topo = false, uniform3 = false;
    GxVDBInfo  vdb;
    int*       d_child[GX_MAXLEV] = {};
    float4*    d_npos[GX_MAXLEV] = {};
    GxLeafRec* d_leaf = nullptr;
    // atlas
    bool                have_atlas = false;
    cudaArray_t         own_array = nullptr;
    cudaTextureObject_t tex = 0;
    cudaTextureObject_t tex_point = 0;  // same array, point filter: exact texel values for the import kernels
    cudaSurfaceObject_t surf = 0;       // same array, for UpdateApron (needs CUDA_ARRAY3D_SURFACE_LDST like the reference's volOut)
    cudaArray_t         array = nullptr; // the array the objects sit on (caller's or own_array)
    float*              d_bricks = nullptr;     // brick-major copy, one block per leaf: built on first use of the linear sampler
    int                 brick_dim = GX_BRICK_DIM, brick_stride = GX_BRICK_STRIDE;
    GxRange*            d_leaf_range = nullptr; // per leaf (valid when topology and atlas are both imported)
    int*                d_err = nullptr;        // error bits raised by the import kernels
    cudaEvent_t         build_ev = nullptr;     // orders lazily built tables (occupancy bits, brick-major copy) against all lanes
    unsigned long long* d_vmask = nullptr;      // SHADE_VOXEL occupancy bits per leaf for THRESH == vmask_thresh
    uint32_t            vmask_thresh_bits = 0;
    bool                vmask_valid = false;
    int                 use_vmask = 1;
    int                 cull = 1;
    int                 ares[3] = {0, 0, 0};
    // colour channel (VDBInfo::clr_chan): uchar4 atlas with the slot layout of channel 0
    cudaTextureObject_t clr_tex = 0;
    cudaArray_t         clr_own = nullptr;
    // transfer function
    float4* d_transfer = nullptr;
    std::vector<float4*> deep_lut;      // per lane (+1 for the creation stream): this frame's {rgb, exp(EXTINCT * alpha * DIRECTSTEP)}
    int     cur_lane = -1;
    // counters
    unsigned long long* d_counters = nullptr;
    // frame lanes: internal streams that consecutive frames alternate between (the tail of frame j overlaps frame j + 1)
    cudaStream_t base_stream = nullptr;
    std::vector<cudaStream_t> lanes;
    std::vector<cudaEvent_t>  lane_ev;
    cudaEvent_t  base_ev = nullptr;
};

#define GX_CUDA(h, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            (h)->err = std::string(#call) + ": " + cudaGetErrorName(e_) + " - " + cudaGetErrorString(e_); \
            return GVDBX_E_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

static inline int gx_fail(gvdbx_t* h, int code, const std::string& msg) { if (h) h->err = msg; return code; }

// makes the handle's context current for the lifetime of the object (no-op when it already is)
struct GxCtx {
    bool pushed = false;
    explicit GxCtx(const gvdbx_t* h)
    {
        if (!h || !h->cuctx || !gx_drv.ok) return;
        void* cur = nullptr;
        if (gx_drv.get(&cur) == 0 && cur == h->cuctx) return;
        pushed = gx_drv.push(h->cuctx) == 0;
    }
    ~GxCtx() { if (pushed) { void* p = nullptr; gx_drv.pop(&p); } }
    GxCtx(const GxCtx&) = delete;
    GxCtx& operator=(const GxCtx&) = delete;
};

