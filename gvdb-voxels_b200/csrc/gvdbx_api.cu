// gvdbx_api.cu — C ABI (include/gvdbx.h) over the sm_100a kernels in gvdbx_device.cuh.
//
// Host-side counterpart of the reference's PrepareVDB / PrepareRender / Render / ReadRenderBuf
// (src/gvdb_volume_gvdb.cpp:3946-3989, 4254-4306, 4336-4381, 4241-4251).  CUDA runtime API only; works under the
// caller's current context; everything is stream-ordered on the stream given to gvdbx_create.
#include "gvdbx_internal.h"
#include "gvdbx_import.cuh"
#include "gvdbx_microbench.cuh"


extern "C" int gvdbx_create(gvdbx_t** out, int cuda_device, void* cuda_stream)
{
    if (!out) return GVDBX_E_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || cuda_device < 0 || cuda_device >= n) {
        fprintf(stderr, "gvdbx_create: no usable CUDA device %d (%s); this library has no CPU path\n", cuda_device,
                e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
        return GVDBX_E_CUDA;
    }
    // adopt the caller's current context when it lives on `cuda_device` (the reference's own, when called from the VolumeGVDB
    // shim); otherwise bind the primary context of `cuda_device` and leave the caller's current device as it was
    void* cur = nullptr;
    int curdev = -1;
    const bool have_cur = gx_drv.load() && gx_drv.get(&cur) == 0 && cur != nullptr && gx_drv.dev(&curdev) == 0;
    if (!have_cur || curdev != cuda_device) {
        int prev = -1;
        if (have_cur) cudaGetDevice(&prev);
        if (cudaSetDevice(cuda_device) != cudaSuccess || cudaFree(0) != cudaSuccess) return GVDBX_E_CUDA;
        cur = nullptr;
        if (gx_drv.ok) gx_drv.get(&cur);
        if (prev >= 0 && prev != cuda_device) cudaSetDevice(prev);
    }
    gvdbx_t* h = new gvdbx_ctx;
    h->device = cuda_device;
    h->cuctx = cur;
    h->stream = (cudaStream_t)cuda_stream;
    h->base_stream = h->stream;
    GxCtx ctx_(h);
    if (cudaMalloc(&h->d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess) { delete h; return GVDBX_E_CUDA; }
    cudaMemset(h->d_counters, 0, 8 * sizeof(unsigned long long));
    if (cudaMalloc(&h->d_err, sizeof(int)) != cudaSuccess) { cudaFree(h->d_counters); delete h; return GVDBX_E_CUDA; }
    cudaMemset(h->d_err, 0, sizeof(int));
    *out = h;
    return GVDBX_OK;
}

static void gx_free_derived(gvdbx_t* h)       // tables that depend on topology AND atlas
{
    if (h->d_leaf_range) cudaFree(h->d_leaf_range);
    if (h->d_vmask) cudaFree(h->d_vmask);
    if (h->d_bricks) cudaFree(h->d_bricks);
    h->d_leaf_range = nullptr; h->d_vmask = nullptr; h->d_bricks = nullptr;
    h->vmask_valid = false;
}
static void gx_free_topology(gvdbx_t* h)
{
    gx_free_derived(h);
    for (int l = 0; l < GX_MAXLEV; l++) {
        if (h->d_child[l]) cudaFree(h->d_child[l]);
        if (h->d_npos[l]) cudaFree(h->d_npos[l]);
        h->d_child[l] = nullptr; h->d_npos[l] = nullptr;
    }
    if (h->d_leaf) cudaFree(h->d_leaf);
    h->d_leaf = nullptr;
    h->have_topo = false;
}
static void gx_free_atlas(gvdbx_t* h)
{
    gx_free_derived(h);
    if (h->tex) cudaDestroyTextureObject(h->tex);
    if (h->tex_point) cudaDestroyTextureObject(h->tex_point);
    if (h->surf) cudaDestroySurfaceObject(h->surf);
    h->surf = 0; h->array = nullptr;
    if (h->own_array) cudaFreeArray(h->own_array);
    h->tex = 0; h->tex_point = 0; h->own_array = nullptr; h->have_atlas = false;
}

extern "C" int gvdbx_destroy(gvdbx_t* h)
{
    if (!h) return GVDBX_E_ARG;
    {
    GxCtx ctx_(h);
    cudaStreamSynchronize(h->stream);
    gvdbx_lanes(h, 0);
    gx_free_topology(h);
    gx_free_atlas(h);
    if (h->clr_tex) cudaDestroyTextureObject(h->clr_tex);
    if (h->clr_own) cudaFreeArray(h->clr_own);
    if (h->d_transfer) cudaFree(h->d_transfer);
    for (float4* p : h->deep_lut) if (p) cudaFree(p);
    if (h->d_counters) cudaFree(h->d_counters);
    if (h->d_err) cudaFree(h->d_err);
    if (h->build_ev) cudaEventDestroy(h->build_ev);
    if (h->band_streams[0]) {
        for (int k = 0; k < 2; k++) { cudaStreamSynchronize(h->band_streams[k]); cudaStreamDestroy(h->band_streams[k]); cudaEventDestroy(h->band_join[k]); }
        cudaStreamDestroy(h->band_copy); cudaEventDestroy(h->band_fork);
        for (int b = 0; b < GX_MAX_BANDS; b++) cudaEventDestroy(h->band_ev[b]);
    }
    }
    delete h;
    return GVDBX_OK;
}

extern "C" const char* gvdbx_last_error(const gvdbx_t* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int gvdbx_set_option(gvdbx_t* h, int option, int value)
{
    if (!h) return GVDBX_E_ARG;
    switch (option) {
    case GVDBX_OPT_SAMPLER:  if (value != 0 && value != 1) return gx_fail(h, GVDBX_E_ARG, "sampler must be 0 or 1"); h->sampler = value; break;
    case GVDBX_OPT_BLOCK_W:  if (value < 1 || value > 32) return gx_fail(h, GVDBX_E_ARG, "block_w"); h->block_w = value; break;
    case GVDBX_OPT_BLOCK_H:  if (value < 1 || value > 32) return gx_fail(h, GVDBX_E_ARG, "block_h"); h->block_h = value; break;
    case GVDBX_OPT_COUNTERS: if (value < 0 || value > 2) return gx_fail(h, GVDBX_E_ARG, "counters must be 0, 1 or 2"); h->count = value; break;
    case GVDBX_OPT_CULL: h->cull = value ? 1 : 0; break;
    case GVDBX_OPT_SPP: if (value < 1 || value > 64) return gx_fail(h, GVDBX_E_ARG, "spp must be 1..64"); h->spp = value; break;
    case GVDBX_OPT_DEEP_SHADOW: h->deep_shadow = value ? 1 : 0; break;
    case GVDBX_OPT_STREAM_MEMOPS: h->memops = value ? 1 : 0; break;
    case GVDBX_OPT_VOXEL_MASK: h->use_vmask = value ? 1 : 0; break;
    case GVDBX_OPT_TRAVERSAL: if (value < 0 || value > 4) return gx_fail(h, GVDBX_E_ARG, "traversal must be 0..4"); h->literal = value; break;
    default: return gx_fail(h, GVDBX_E_ARG, "unknown option");
    }
    if (h->block_w * h->block_h > GX_MAXTHREADS || (h->block_w * h->block_h) % 32 != 0) {
        const bool pending = (option == GVDBX_OPT_BLOCK_W);     // width is set first, height second: only judge the pair
        if (!pending) { h->block_w = 8; h->block_h = 8; return gx_fail(h, GVDBX_E_ARG, "CTA tile must be a multiple of 32 and at most 128 threads"); }
    }
    return GVDBX_OK;
}

// error bits of the import kernels -> return code (synchronises: imports are not the per-frame path)
static int gx_check_import(gvdbx_t* h)
{
    int e = 0;
    GX_CUDA(h, cudaMemcpyAsync(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!e) return GVDBX_OK;
    e &= ~GX_ERR_WAIT_TIMEOUT;
    if (!e) return GVDBX_OK;
    gx_clear_import_bits<<<1, 1, 0, h->stream>>>(h->d_err);
    std::string msg = "malformed pools:";
    if (e & GX_IMPORT_E_CHILDLIST) msg += " a node's mChildList points past the child-list pool;";
    if (e & GX_IMPORT_E_CHILD) msg += " a child entry points past the node pool of the level below;";
    if (e & GX_IMPORT_E_LEAFSLOT) msg += " a leaf's mValue brick lies outside the atlas (VDBInfo.atlas_res);";
    return gx_fail(h, GVDBX_E_ARG, msg);
}

// Tables derived from topology AND atlas: built by whichever import comes last.  Per-leaf value ranges are reduced straight
// from the atlas array (no staging copy of the atlas exists); occupancy bits and the brick-major copy are built on first use.
static int gx_update_leaf_ranges(gvdbx_t* h)
{
    gx_free_derived(h);
    if (!h->have_topo || !h->tex_point || !h->d_leaf) return GVDBX_OK;
    const GxVDBInfo& v = h->vdb;
    if (v.atlas_res.x != h->ares[0] || v.atlas_res.y != h->ares[1] || v.atlas_res.z != h->ares[2])
        return gx_fail(h, GVDBX_E_ARG, "VDBInfo.atlas_res does not match the imported atlas (stale VDBInfo, or atlas of another volume)");
    const int n = v.nodecnt[0];
    GX_CUDA(h, cudaMalloc(&h->d_leaf_range, size_t(n) * sizeof(GxRange)));
    gx_leaf_ranges_array<<<n, 128, 0, h->stream>>>(h->tex_point, h->d_leaf, n, h->brick_dim, h->d_leaf_range);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

// A table that frames in flight on OTHER lanes may be reading (or are about to read) is (re)built in stream order: the
// building stream first waits for everything enqueued on every lane and on the creation stream, and afterwards every lane
// waits for the build — events only, no host synchronisation.
static int gx_build_begin(gvdbx_t* h)
{
    if (!h->build_ev) GX_CUDA(h, cudaEventCreateWithFlags(&h->build_ev, cudaEventDisableTiming));
    for (size_t i = 0; i < h->lanes.size(); i++) {
        if (h->lanes[i] == h->stream) continue;
        GX_CUDA(h, cudaEventRecord(h->lane_ev[i], h->lanes[i]));
        GX_CUDA(h, cudaStreamWaitEvent(h->stream, h->lane_ev[i], 0));
    }
    if (h->base_stream != h->stream && h->base_ev) {
        GX_CUDA(h, cudaEventRecord(h->base_ev, h->base_stream));
        GX_CUDA(h, cudaStreamWaitEvent(h->stream, h->base_ev, 0));
    }
    return GVDBX_OK;
}
static int gx_build_end(gvdbx_t* h)
{
    GX_CUDA(h, cudaEventRecord(h->build_ev, h->stream));
    for (cudaStream_t s : h->lanes) if (s != h->stream) GX_CUDA(h, cudaStreamWaitEvent(s, h->build_ev, 0));
    if (h->base_stream != h->stream) GX_CUDA(h, cudaStreamWaitEvent(h->base_stream, h->build_ev, 0));
    return GVDBX_OK;
}

// SHADE_VOXEL occupancy bits for the frame's THRESH (8^3 bricks): rebuilt — one pass over the interior texels — only when
// THRESH, the atlas or the topology changed; stream-ordered against the frames of all lanes (gx_build_begin / _end).
static int gx_ensure_voxel_mask(gvdbx_t* h, float thresh)
{
    uint32_t bits;
    memcpy(&bits, &thresh, 4);
    if (h->vmask_valid && bits == h->vmask_thresh_bits) return GVDBX_OK;
    const int n = h->vdb.nodecnt[0];
    if (!h->d_vmask) GX_CUDA(h, cudaMalloc(&h->d_vmask, size_t(n) * 64));
    int rc = gx_build_begin(h);
    if (rc) return rc;
    gx_build_voxel_mask<<<n, 64, 0, h->stream>>>(h->tex_point, h->d_leaf, n, thresh, (unsigned char*)h->d_vmask);
    GX_CUDA(h, cudaGetLastError());
    rc = gx_build_end(h);
    if (rc) return rc;
    h->vmask_thresh_bits = bits;
    h->vmask_valid = true;
    return GVDBX_OK;
}

// brick-major copy of the atlas (one block per leaf) for the linear sampler: built when that sampler is first used
static int gx_ensure_bricks(gvdbx_t* h)
{
    if (h->d_bricks) return GVDBX_OK;
    const int n = h->vdb.nodecnt[0];
    GX_CUDA(h, cudaMalloc(&h->d_bricks, size_t(n) * size_t(h->brick_stride) * sizeof(float)));
    int rc = gx_build_begin(h);
    if (rc) return rc;
    gx_copy_bricks_array<<<n, 256, 0, h->stream>>>(h->tex_point, h->d_leaf, n, h->brick_dim, h->brick_stride, h->d_bricks);
    GX_CUDA(h, cudaGetLastError());
    return gx_build_end(h);
}

// ------------------------------------------------------------------------------------------------ topology
// listcnt[l] = number of child lists in pool 1 of level l when known (host import: pool bytes / childwid), else 0 = bounded
// by the node count of that level (every node owns at most one list)
static int gx_import_topology(gvdbx_t* h, const void* vdbinfo, const uint64_t* listcnt)
{
    if (!h || !vdbinfo) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GxVDBInfo v;
    memcpy(&v, vdbinfo, sizeof v);
    if (v.top_lev < 0 || v.top_lev >= GX_MAXLEV) return gx_fail(h, GVDBX_E_ARG, "top_lev out of range (0..4)");
    // bricks: Configure(.., q0) with q0 = 2..5 (4^3 .. 32^3 voxels; the reference's samples use 3, 4 and 5), apron 1
    if (v.atlas_apron != 1) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only apron 1 is supported (the reference's default, gvdb_volume_gvdb.cpp:73)");
    if (v.dim[0] < 2 || v.dim[0] > 5) return gx_fail(h, GVDBX_E_UNSUPPORTED, "brick log2dim must be 2..5");
    if (v.brick_res != v.res[0] + 2 * v.atlas_apron) return gx_fail(h, GVDBX_E_ARG, "brick_res != res[0] + 2 * apron");
    if (v.atlas_res.x <= 0 || v.atlas_res.y <= 0 || v.atlas_res.z <= 0 || v.atlas_res.x % v.brick_res || v.atlas_res.y % v.brick_res || v.atlas_res.z % v.brick_res)
        return gx_fail(h, GVDBX_E_ARG, "atlas_res must be a positive multiple of brick_res");
    for (int l = 0; l <= v.top_lev; l++) {
        if (v.dim[l] < 1 || v.dim[l] > 8 || v.res[l] != (1 << v.dim[l])) return gx_fail(h, GVDBX_E_ARG, "inconsistent dim/res");
        if (v.nodecnt[l] <= 0 || v.nodelist[l] == 0) return gx_fail(h, GVDBX_E_ARG, "empty node pool below top_lev");
        if (v.nodewid[l] < (int)sizeof(GxNode)) return gx_fail(h, GVDBX_E_ARG, "nodewid smaller than a node record");
        if (l >= 1 && (v.childlist[l] == 0 || v.childwid[l] != 8 * (1 << (3 * v.dim[l]))))
            return gx_fail(h, GVDBX_E_UNSUPPORTED, "child lists must be dense (bitmasks off): childwid = 8*res^3");
    }
    gx_free_topology(h);
    for (int l = 1; l <= v.top_lev; l++) {
        const int cells = 1 << (3 * v.dim[l]);
        const size_t total = size_t(v.nodecnt[l]) * cells;
        GX_CUDA(h, cudaMalloc(&h->d_child[l], total * sizeof(int)));
        GX_CUDA(h, cudaMalloc(&h->d_npos[l], size_t(v.nodecnt[l]) * sizeof(float4)));
        const int threads = 256;
        const unsigned blocks = (unsigned)((total + threads - 1) / threads);
        const unsigned long long lists = (listcnt && listcnt[l]) ? listcnt[l] : (unsigned long long)v.nodecnt[l];
        gx_build_child_table<<<blocks, threads, 0, h->stream>>>((const char*)v.nodelist[l], v.nodewid[l], v.nodecnt[l],
                                                               (const char*)v.childlist[l], v.childwid[l], cells, lists, v.nodecnt[l - 1],
                                                               h->d_child[l], h->d_npos[l], h->d_err);
        GX_CUDA(h, cudaGetLastError());
    }
    GX_CUDA(h, cudaMalloc(&h->d_leaf, size_t(v.nodecnt[0]) * sizeof(GxLeafRec)));
    gx_build_leaf_table<<<(v.nodecnt[0] + 255) / 256, 256, 0, h->stream>>>((const char*)v.nodelist[0], v.nodewid[0], v.nodecnt[0], v.brick_res,
                                                                          make_int3(v.atlas_res.x, v.atlas_res.y, v.atlas_res.z), h->d_leaf, h->d_err);
    GX_CUDA(h, cudaGetLastError());
    int rc = gx_check_import(h);
    if (rc) { gx_free_topology(h); return rc; }
    h->vdb = v;
    h->have_topo = true;
    h->brick_dim = v.brick_res;
    h->brick_stride = (v.brick_res * v.brick_res * v.brick_res + 255) / 256 * 256;
    h->uniform3 = true;
    for (int l = 0; l <= v.top_lev; l++) if (v.dim[l] != 3 || v.vdel[l].x != float(1 << (3 * l)) || v.vdel[l].y != v.vdel[l].x || v.vdel[l].z != v.vdel[l].x) h->uniform3 = false;
    // VDBInfo is the authority on the atlas geometry: an atlas imported earlier with another resolution is stale (UpdateAtlas
    // re-allocates the reference's array when the volume grows) and is dropped; import it again
    if (h->have_atlas && (v.atlas_res.x != h->ares[0] || v.atlas_res.y != h->ares[1] || v.atlas_res.z != h->ares[2])) gx_free_atlas(h);
    return gx_update_leaf_ranges(h);
}

extern "C" int gvdbx_import_topology(gvdbx_t* h, const void* vdbinfo) { return gx_import_topology(h, vdbinfo, nullptr); }

extern "C" int gvdbx_import_topology_host(gvdbx_t* h, const void* vdbinfo, const void* const* pool0, const void* const* pool1,
                                          const uint64_t* pool1_bytes)
{
    if (!h || !vdbinfo || !pool0 || !pool1 || !pool1_bytes) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GxVDBInfo v;
    memcpy(&v, vdbinfo, sizeof v);
    if (v.top_lev < 0 || v.top_lev >= GX_MAXLEV) return gx_fail(h, GVDBX_E_ARG, "top_lev out of range (0..4)");
    std::vector<void*> tmp;
    int rc = GVDBX_OK;
    auto up = [&](const void* src, size_t bytes, uint64_t* dst) -> bool {
        void* d = nullptr;
        if (bytes == 0 || !src) { *dst = 0; return true; }
        if (cudaMalloc(&d, bytes) != cudaSuccess) return false;
        tmp.push_back(d);
        if (cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return false;
        *dst = (uint64_t)d;
        return true;
    };
    for (int l = 0; l <= v.top_lev && rc == GVDBX_OK; l++) {
        if (!up(pool0[l], size_t(v.nodecnt[l]) * v.nodewid[l], &v.nodelist[l])) rc = GVDBX_E_CUDA;
        if (l >= 1 && !up(pool1[l], (size_t)pool1_bytes[l], &v.childlist[l])) rc = GVDBX_E_CUDA;
    }
    if (rc == GVDBX_OK) {
        uint64_t listcnt[GX_MAXLEV] = {0, 0, 0, 0, 0};
        for (int l = 1; l <= v.top_lev; l++) if (v.childwid[l] > 0) listcnt[l] = pool1_bytes[l] / (uint64_t)v.childwid[l];
        rc = gx_import_topology(h, &v, listcnt);
    } else h->err = "pool upload failed";
    cudaStreamSynchronize(h->stream);
    for (void* d : tmp) cudaFree(d);
    return rc;
}

// ------------------------------------------------------------------------------------------------ atlas
static int gx_make_texture(gvdbx_t* h, cudaArray_t arr, int rx, int ry, int rz)
{
    // same descriptor as SetupAtlasAccess (gvdb_volume_gvdb.cpp:753-783): linear filter, element-type reads,
    // unnormalised coordinates, clamp addressing; plus a point-filter view of the same array for the import kernels
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.normalizedCoords = 0;
    GX_CUDA(h, cudaCreateTextureObject(&h->tex, &rd, &td, nullptr));
    td.filterMode = cudaFilterModePoint;
    GX_CUDA(h, cudaCreateTextureObject(&h->tex_point, &rd, &td, nullptr));
    h->array = arr;
    if (cudaCreateSurfaceObject(&h->surf, &rd) != cudaSuccess) { h->surf = 0; cudaGetLastError(); }   // array without the surface flag: no UpdateApron
    h->ares[0] = rx; h->ares[1] = ry; h->ares[2] = rz;
    return GVDBX_OK;
}

// The atlas is sampled where it lies: the caller's array (gvdbx_import_atlas_array) or one array owned by the library
// (_host / _device).  No second copy is made at import; what is derived from it (value ranges now, occupancy bits and the
// brick-major copy on first use) is read straight from the array.
extern "C" int gvdbx_import_atlas_array(gvdbx_t* h, int chan, void* cuarray, int rx, int ry, int rz)
{
    if (!h || !cuarray || rx <= 0 || ry <= 0 || rz <= 0) return GVDBX_E_ARG;
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 (T_FLOAT) is supported");
    GxCtx ctx_(h);
    gx_free_atlas(h);
    int rc = gx_make_texture(h, (cudaArray_t)cuarray, rx, ry, rz);
    if (rc) return rc;
    h->have_atlas = true;
    rc = gx_update_leaf_ranges(h);
    if (rc) gx_free_atlas(h);                           // an atlas that contradicts the topology is not kept
    return rc;
}

static int gx_import_atlas_linear(gvdbx_t* h, const void* texels, int rx, int ry, int rz, cudaMemcpyKind kind)
{
    gx_free_atlas(h);
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<float>();
    GX_CUDA(h, cudaMalloc3DArray(&h->own_array, &fd, make_cudaExtent(rx, ry, rz), cudaArraySurfaceLoadStore));
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr((void*)texels, size_t(rx) * sizeof(float), rx, ry);
    cp.dstArray = h->own_array;
    cp.extent = make_cudaExtent(rx, ry, rz);
    cp.kind = kind;
    GX_CUDA(h, cudaMemcpy3DAsync(&cp, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));       // the caller may free / reuse its image now
    int rc = gx_make_texture(h, h->own_array, rx, ry, rz);
    if (rc) return rc;
    h->have_atlas = true;
    rc = gx_update_leaf_ranges(h);
    if (rc) gx_free_atlas(h);                           // an atlas that contradicts the topology is not kept
    return rc;
}

extern "C" int gvdbx_import_atlas_host(gvdbx_t* h, int chan, const float* texels, int rx, int ry, int rz)
{
    if (!h || !texels || rx <= 0 || ry <= 0 || rz <= 0) return GVDBX_E_ARG;
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 (T_FLOAT) is supported");
    GxCtx ctx_(h);
    return gx_import_atlas_linear(h, texels, rx, ry, rz, cudaMemcpyHostToDevice);
}

// ------------------------------------------------------------------------------------------------ UpdateApron / atlas read-back
static void gx_tree_params(gvdbx_t* h, GxParams& P)
{
    memset(&P, 0, sizeof P);
    const GxVDBInfo& v = h->vdb;
    for (int l = 0; l < GX_MAXLEV; l++) {
        P.dim[l] = v.dim[l]; P.res[l] = v.res[l]; P.vdel[l] = make_float3(v.vdel[l].x, v.vdel[l].y, v.vdel[l].z);
        P.noderange[l] = make_int3(v.noderange[l].x, v.noderange[l].y, v.noderange[l].z);
        P.child[l] = h->d_child[l]; P.npos[l] = h->d_npos[l];
    }
    P.top_lev = v.top_lev; P.epsilon = v.epsilon;
    P.leaf = h->d_leaf;
    P.tex = h->tex; P.bricks = h->d_bricks;
    P.brick_dim = h->brick_dim; P.brick_stride = h->brick_stride;
}

extern "C" int gvdbx_update_apron(gvdbx_t* h, int chan, float boundval)
{
    if (!h) return GVDBX_E_ARG;
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 (T_FLOAT) is supported");
    if (!h->have_topo || !h->have_atlas) return gx_fail(h, GVDBX_E_STATE, "UpdateApron needs topology and atlas");
    if (!h->surf) return gx_fail(h, GVDBX_E_UNSUPPORTED, "the atlas array was created without surface load/store");
    GxCtx ctx_(h);
    GxParams P;
    gx_tree_params(h, P);
    const int n = h->vdb.nodecnt[0];
    int rc = gx_build_begin(h);                         // frames in flight on other lanes read the texels this pass rewrites
    if (rc) return rc;
    gx_update_apron_kernel<<<n, 128, 0, h->stream>>>(P, h->surf, h->d_bricks, n, boundval);
    GX_CUDA(h, cudaGetLastError());
    // value ranges over interior + apron change with the aprons; occupancy bits (interior voxels only) do not
    gx_leaf_ranges_array<<<n, 128, 0, h->stream>>>(h->tex_point, h->d_leaf, n, h->brick_dim, h->d_leaf_range);
    GX_CUDA(h, cudaGetLastError());
    return gx_build_end(h);
}

extern "C" int gvdbx_update_apron_faces(gvdbx_t* h, int chan)
{
    if (!h) return GVDBX_E_ARG;
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 (T_FLOAT) is supported");
    if (!h->have_topo || !h->have_atlas) return gx_fail(h, GVDBX_E_STATE, "UpdateApronFaces needs topology and atlas");
    if (!h->surf) return gx_fail(h, GVDBX_E_UNSUPPORTED, "the atlas array was created without surface load/store");
    GxCtx ctx_(h);
    GxParams P;
    gx_tree_params(h, P);
    const int n = h->vdb.nodecnt[0];
    int rc = gx_build_begin(h);
    if (rc) return rc;
    gx_update_apron_faces_kernel<<<n, 192, 0, h->stream>>>(P, h->surf, h->d_bricks, n);
    GX_CUDA(h, cudaGetLastError());
    gx_leaf_ranges_array<<<n, 128, 0, h->stream>>>(h->tex_point, h->d_leaf, n, h->brick_dim, h->d_leaf_range);
    GX_CUDA(h, cudaGetLastError());
    return gx_build_end(h);
}

// host image of the atlas array, x fastest (the inverse of gvdbx_import_atlas_host; Allocator::AtlasRetrieveSlice per slice)
extern "C" int gvdbx_export_atlas_host(gvdbx_t* h, int chan, float* texels, int rx, int ry, int rz)
{
    if (!h || !texels) return GVDBX_E_ARG;
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 (T_FLOAT) is supported");
    if (!h->have_atlas || !h->array) return gx_fail(h, GVDBX_E_STATE, "no atlas imported");
    if (rx != h->ares[0] || ry != h->ares[1] || rz != h->ares[2]) return gx_fail(h, GVDBX_E_ARG, "atlas resolution mismatch");
    GxCtx ctx_(h);
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcArray = h->array;
    cp.dstPtr = make_cudaPitchedPtr(texels, size_t(rx) * sizeof(float), rx, ry);
    cp.extent = make_cudaExtent(rx, ry, rz);
    cp.kind = cudaMemcpyDeviceToHost;
    GX_CUDA(h, cudaMemcpy3DAsync(&cp, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));
    return GVDBX_OK;
}

// Same from a DEVICE image of the atlas (x fastest): what a rank receives when the volume is replicated with a
// broadcast over NVLink instead of being rebuilt / re-uploaded by every process.
extern "C" int gvdbx_import_atlas_device(gvdbx_t* h, int chan, uint64_t texels_d, int rx, int ry, int rz)
{
    if (!h || !texels_d || rx <= 0 || ry <= 0 || rz <= 0) return GVDBX_E_ARG;
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 (T_FLOAT) is supported");
    GxCtx ctx_(h);
    return gx_import_atlas_linear(h, (const void*)texels_d, rx, ry, rz, cudaMemcpyDeviceToDevice);
}

// ------------------------------------------------------------------------------------------------ colour channel
static void gx_free_color(gvdbx_t* h)
{
    if (h->clr_tex) cudaDestroyTextureObject(h->clr_tex);
    if (h->clr_own) cudaFreeArray(h->clr_own);
    h->clr_tex = 0; h->clr_own = nullptr;
}
static int gx_make_color_texture(gvdbx_t* h, cudaArray_t arr, int filter)
{
    // same descriptor as SetupAtlasAccess builds for a T_UCHAR4 channel (gvdb_volume_gvdb.cpp:753-783): integer reads,
    // unnormalised coordinates, the channel's filter mode (F_POINT for a usable colour channel: CUDA rejects linear
    // filtering of integer reads, for the reference's cuTexObjectCreate as for this call)
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.filterMode = filter ? cudaFilterModeLinear : cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.normalizedCoords = 0;
    GX_CUDA(h, cudaCreateTextureObject(&h->clr_tex, &rd, &td, nullptr));
    return GVDBX_OK;
}
extern "C" int gvdbx_import_color_array(gvdbx_t* h, void* cuarray, int filter)
{
    if (!h || !cuarray) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    gx_free_color(h);
    return gx_make_color_texture(h, (cudaArray_t)cuarray, filter);
}
extern "C" int gvdbx_import_color_host(gvdbx_t* h, const void* rgba8_texels, int rx, int ry, int rz, int filter)
{
    if (!h || !rgba8_texels || rx <= 0 || ry <= 0 || rz <= 0) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    gx_free_color(h);
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    GX_CUDA(h, cudaMalloc3DArray(&h->clr_own, &fd, make_cudaExtent(rx, ry, rz), 0));
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr((void*)rgba8_texels, size_t(rx) * 4, rx, ry);
    cp.dstArray = h->clr_own;
    cp.extent = make_cudaExtent(rx, ry, rz);
    cp.kind = cudaMemcpyHostToDevice;
    GX_CUDA(h, cudaMemcpy3DAsync(&cp, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));
    return gx_make_color_texture(h, h->clr_own, filter);
}
extern "C" int gvdbx_clear_color(gvdbx_t* h)
{
    if (!h) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    gx_free_color(h);
    return GVDBX_OK;
}

extern "C" int gvdbx_set_transfer(gvdbx_t* h, const float* rgba_host)
{
    if (!h || !rgba_host) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    if (!h->d_transfer) GX_CUDA(h, cudaMalloc(&h->d_transfer, GVDBX_TRANSFER_ENTRIES * sizeof(float4)));
    GX_CUDA(h, cudaMemcpyAsync(h->d_transfer, rgba_host, GVDBX_TRANSFER_ENTRIES * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));      // the host buffer may be reused by the caller
    return GVDBX_OK;
}

// ------------------------------------------------------------------------------------------------ render
typedef void (*gx_kernel_t)(const GxParams);
// one translation unit per shade mode (gvdbx_k_*.cu, gvdbx_pick.cuh)
gx_kernel_t gx_pick_voxel(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_trilinear(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_levelset(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_deep(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_deepshadow(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_tricubic(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_emptyskip(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_section2d(int sampler, int flags, bool uni);
gx_kernel_t gx_pick_section3d(int sampler, int flags, bool uni);

static gx_kernel_t gx_pick(int mode, int sampler, int flags, bool uni)
{
    switch (mode) {
    case GX_MODE_VOXEL:      return gx_pick_voxel(sampler, flags, uni);
    case GX_MODE_TRILINEAR:  return gx_pick_trilinear(sampler, flags, uni);
    case GX_MODE_LEVELSET:   return gx_pick_levelset(sampler, flags, uni);
    case GX_MODE_DEEP:       return gx_pick_deep(sampler, flags, uni);
    case GX_MODE_DEEPSHADOW: return gx_pick_deepshadow(sampler, flags, uni);
    case GX_MODE_TRICUBIC:   return gx_pick_tricubic(sampler, flags, uni);
    case GX_MODE_EMPTYSKIP:  return gx_pick_emptyskip(sampler, flags, uni);
    case GX_MODE_SECTION2D:  return gx_pick_section2d(sampler, flags, uni);
    case GX_MODE_SECTION3D:  return gx_pick_section3d(sampler, flags, uni);
    }
    return nullptr;
}

static inline float3 f3(const GxF3& a) { return make_float3(a.x, a.y, a.z); }

// brick-queue traversal (gx_raycast_deep_q / gx_raycast_surface_q): default for the deep modes (cfg4 4K: 24.4 -> 20.2 ms) and
// SHADE_TRILINEAR (cfg1: +6 %); SHADE_LEVELSET keeps the literal nesting (the queue measured -2 % .. +2 % there) unless
// GVDBX_OPT_TRAVERSAL = 3 asks for it; 4 switches the queue off everywhere (A/B)
static int gx_queue_flag(const gvdbx_t* h, int mode)
{
    if (mode == GX_MODE_DEEP || mode == GX_MODE_DEEPSHADOW || mode == GX_MODE_TRILINEAR || mode == GX_MODE_LEVELSET)
        return (h->literal == 0 || h->literal == 3) ? GX_FLAG_QUEUE : 0;
    return 0;
}
// dynamic shared memory: the traversal stack, plus the brick queue of the queue variants
static size_t gx_smem_bytes(dim3 block, int flags)
{
    return size_t(block.x) * block.y * (GX_STACK_BYTES_PER_THREAD + ((flags & GX_FLAG_QUEUE) ? GX_QUEUE_BYTES_PER_THREAD : 0));
}

static int gx_fill_params(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, GxParams& P, int& mode, bool force_deep_lut = false)
{
    if (!h || !scninfo) return GVDBX_E_ARG;
    if (!h->have_topo) return gx_fail(h, GVDBX_E_STATE, "render before gvdbx_import_topology");
    if (!h->have_atlas) return gx_fail(h, GVDBX_E_STATE, "render before gvdbx_import_atlas_*");
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 is supported");
    switch (shade_mode) {
    case GVDBX_SHADE_VOXEL:     mode = GX_MODE_VOXEL; break;
    case GVDBX_SHADE_TRILINEAR: mode = GX_MODE_TRILINEAR; break;
    case GVDBX_SHADE_LEVELSET:  mode = GX_MODE_LEVELSET; break;
    case GVDBX_SHADE_VOLUME:    mode = h->deep_shadow ? GX_MODE_DEEPSHADOW : GX_MODE_DEEP; break;
    case GVDBX_SHADE_TRICUBIC:  mode = GX_MODE_TRICUBIC; break;
    case GVDBX_SHADE_EMPTYSKIP: mode = GX_MODE_EMPTYSKIP; break;
    case GVDBX_SHADE_SECTION2D: mode = GX_MODE_SECTION2D; break;
    case GVDBX_SHADE_SECTION3D: mode = GX_MODE_SECTION3D; break;
    default: return gx_fail(h, GVDBX_E_UNSUPPORTED, "unknown shade mode (the reference's Render switch handles 0..7, gvdb_volume_gvdb.cpp:4363-4372)");
    }
    if (h->sampler != GX_SAMPLER_TEX && (mode == GX_MODE_TRICUBIC || mode == GX_MODE_EMPTYSKIP || mode == GX_MODE_SECTION2D || mode == GX_MODE_SECTION3D))
        return gx_fail(h, GVDBX_E_UNSUPPORTED, "this shade mode needs the texture sampler (tricubic taps reach beyond the brick apron)");
    GxScnInfo s;
    memcpy(&s, scninfo, sizeof s);
    if (s.width <= 0 || s.height <= 0) return gx_fail(h, GVDBX_E_ARG, "ScnInfo width/height");
    memset(&P, 0, sizeof P);
    P.width = s.width; P.height = s.height; P.camnear = s.camnear; P.camfar = s.camfar;
    P.campos = f3(s.campos); P.cams = f3(s.cams); P.camu = f3(s.camu); P.camv = f3(s.camv);
    P.light_pos = f3(s.light_pos); P.shadow_params = f3(s.shadow_params);
    P.slice_pnt = f3(s.slice_pnt); P.slice_norm = f3(s.slice_norm);
    P.backclr = make_float4(s.backclr.x, s.backclr.y, s.backclr.z, s.backclr.w);
    memcpy(P.xform, s.xform, sizeof P.xform);
    memcpy(P.invxform, s.invxform, sizeof P.invxform);
    memcpy(P.invxrot, s.invxrot, sizeof P.invxrot);
    P.extinct = f3(s.extinct); P.steps = f3(s.steps); P.cutoff = f3(s.cutoff); P.thresh = f3(s.thresh);
    P.transfer = h->d_transfer ? h->d_transfer : (const float4*)s.transfer;
    if ((mode == GX_MODE_DEEP || mode == GX_MODE_DEEPSHADOW || mode == GX_MODE_SECTION2D || mode == GX_MODE_SECTION3D) && !P.transfer)
        return gx_fail(h, GVDBX_E_STATE, "transfer function not on GPU (reference: 'Must call CommitTransferFunc')");
    P.dbuf = (const float*)s.dbuf;
    P.transfer_deep = nullptr;
    if (mode == GX_MODE_DEEP || mode == GX_MODE_DEEPSHADOW || (force_deep_lut && P.transfer)) {
        // this frame's derived table, in the buffer of the stream / lane the frame is enqueued on (stream-ordered, ~2 us)
        const size_t slot = size_t(h->cur_lane + 1);
        if (h->deep_lut.size() <= slot) h->deep_lut.resize(slot + 1, nullptr);
        if (!h->deep_lut[slot]) GX_CUDA(h, cudaMalloc(&h->deep_lut[slot], GVDBX_TRANSFER_ENTRIES * sizeof(float4)));
        gx_build_deep_lut<<<GVDBX_TRANSFER_ENTRIES / 256, 256, 0, h->stream>>>(P.transfer, h->deep_lut[slot], GVDBX_TRANSFER_ENTRIES, P.extinct.x, P.steps.x);
        GX_CUDA(h, cudaGetLastError());
        P.transfer_deep = h->deep_lut[slot];
    }
    const GxVDBInfo& v = h->vdb;
    for (int l = 0; l < GX_MAXLEV; l++) {
        P.dim[l] = v.dim[l]; P.res[l] = v.res[l]; P.vdel[l] = f3(v.vdel[l]);
        P.noderange[l] = make_int3(v.noderange[l].x, v.noderange[l].y, v.noderange[l].z);
        P.child[l] = h->d_child[l]; P.npos[l] = h->d_npos[l];
    }
    P.top_lev = v.top_lev; P.epsilon = v.epsilon; P.bmin = f3(v.bmin); P.bmax = f3(v.bmax);
    P.leaf = h->d_leaf;
    if (h->sampler == GX_SAMPLER_LINEAR) { const int rc = gx_ensure_bricks(h); if (rc) return rc; }
    P.tex = h->tex; P.bricks = h->d_bricks;
    P.brick_dim = h->brick_dim; P.brick_stride = h->brick_stride;
    // the colour channel is used exactly when the reference would: VDBInfo::clr_chan set (SetColorChannel)
    P.clr_tex = 0;
    if (v.clr_chan != GX_CHAN_UNDEF) {
        if (!h->clr_tex) return gx_fail(h, GVDBX_E_STATE, "VDBInfo.clr_chan is set but no colour atlas was imported (gvdbx_import_color_*)");
        P.clr_tex = h->clr_tex;
    }
    P.range = h->cull ? h->d_leaf_range : nullptr;
    P.vmask = nullptr;
    if (mode == GX_MODE_VOXEL && h->use_vmask && v.res[0] == 8) {
        const int rc = gx_ensure_voxel_mask(h, s.thresh.x);
        if (rc) return rc;
        P.vmask = h->d_vmask;
    }
    P.counters = h->d_counters;
    P.out_stride = s.width;
    P.x0 = 0; P.y0 = 0; P.x1 = s.width; P.y1 = s.height;
    // sub-pixel pattern: g x g grid with g = ceil(sqrt(spp)); sample s at ((s % g) + .5) / g, ((s / g) + .5) / g
    int g = 1;
    while (g * g < h->spp) g++;
    P.spp = h->spp; P.spp_grid = g; P.spp_inv_grid = 1.0f / float(g); P.spp_inv = 1.0f / float(h->spp);
    return GVDBX_OK;
}

extern "C" int gvdbx_render(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d,
                            int tx0, int ty0, int tw, int th)
{
    if (!h) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    if (shade_mode == GVDBX_SHADE_OFF) {    // gvdb_volume_gvdb.cpp:4340-4346
        const GxScnInfo* s = (const GxScnInfo*)scninfo;
        if (!s || !outbuf_d) return GVDBX_E_ARG;
        GX_CUDA(h, cudaMemsetAsync((void*)outbuf_d, 0, size_t(s->width) * s->height * 4, h->stream));
        return GVDBX_OK;
    }
    GxParams P; int mode = 0;
    int rc = gx_fill_params(h, scninfo, shade_mode, chan, P, mode);
    if (rc) return rc;
    if (!outbuf_d) return gx_fail(h, GVDBX_E_ARG, "null output buffer");
    if (tw > 0 && th > 0) {
        if (tx0 < 0 || ty0 < 0 || tx0 + tw > P.width || ty0 + th > P.height) return gx_fail(h, GVDBX_E_ARG, "tile outside the frame");
        P.x0 = tx0; P.y0 = ty0; P.x1 = tx0 + tw; P.y1 = ty0 + th;
    }
    P.out = (uchar4*)outbuf_d;
    const bool core = (mode <= GX_MODE_DEEP);          // the A/B traversal variants exist for the four core modes only
    const int qf = gx_queue_flag(h, mode);
    const int flags = h->count ? (GX_FLAG_DEBUG | GX_FLAG_COUNT)
                    : (h->spp > 1 ? (GX_FLAG_SPP | qf)
                    : (core && h->literal == 1 ? GX_FLAG_LITERAL : (core && h->literal == 2 ? GX_FLAG_PACKET : qf)));
    // counters = 1 and the A/B baseline follow the reference's own work (no brick culling); counters = 2 count what the
    // production kernel really does
    if ((h->count == 1) || (flags & GX_FLAG_LITERAL)) P.range = nullptr;
    if (h->count && h->spp > 1) return gx_fail(h, GVDBX_E_ARG, "work counters are taken at 1 ray per pixel: set GVDBX_OPT_SPP to 1 for the counted render");
    gx_kernel_t k = gx_pick(mode, h->sampler, flags, h->uniform3);
    if (!k) return gx_fail(h, GVDBX_E_UNSUPPORTED, "no kernel variant for this mode / sampler / option combination");
    float4* dbg_tmp = nullptr;
    if (h->count) {     // counted renders reuse the debug variant; give it a scratch debug buffer
        GX_CUDA(h, cudaMalloc(&dbg_tmp, size_t(P.width) * P.height * 48));
        P.dbg = dbg_tmp;
        GX_CUDA(h, cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(unsigned long long), h->stream));
    }
    dim3 block(h->block_w, h->block_h, 1);
    dim3 grid((P.x1 - P.x0 + block.x - 1) / block.x, (P.y1 - P.y0 + block.y - 1) / block.y, 1);
    k<<<grid, block, gx_smem_bytes(block, flags), h->stream>>>(P);
    GX_CUDA(h, cudaGetLastError());
    if (dbg_tmp) { cudaStreamSynchronize(h->stream); cudaFree(dbg_tmp); }
    return GVDBX_OK;
}

// RenderKernel plugin point: the parameter block of this frame for a user kernel built against csrc/gvdbx_plugin.cuh
extern "C" int gvdbx_kernel_params(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d, void* params_out,
                                   size_t params_bytes)
{
    if (!h || !params_out) return GVDBX_E_ARG;
    if (params_bytes != sizeof(GxParams)) return gx_fail(h, GVDBX_E_ARG, "params_bytes != sizeof(GxParams): plugin built against other headers");
    GxCtx ctx_(h);
    GxParams P; int mode = 0;
    // a plugin kernel may march deep bricks whatever `shade_mode` says: the derived table is always prepared
    int rc = gx_fill_params(h, scninfo, shade_mode, chan, P, mode, true);
    if (rc) return rc;
    P.out = (uchar4*)outbuf_d;
    memcpy(params_out, &P, sizeof P);
    return GVDBX_OK;
}

extern "C" int gvdbx_render_debug(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d, uint64_t dbg_d)
{
    if (!h) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GxParams P; int mode = 0;
    int rc = gx_fill_params(h, scninfo, shade_mode, chan, P, mode);
    if (rc) return rc;
    if (!outbuf_d || !dbg_d) return gx_fail(h, GVDBX_E_ARG, "null output buffer");
    P.out = (uchar4*)outbuf_d;
    P.dbg = (float4*)dbg_d;
    P.range = nullptr;                                  // debug + counters follow the reference's own work (no brick culling)
    GX_CUDA(h, cudaMemsetAsync(h->d_counters, 0, 8 * sizeof(unsigned long long), h->stream));
    gx_kernel_t k = gx_pick(mode, h->sampler, GX_FLAG_DEBUG | GX_FLAG_COUNT, h->uniform3);
    if (!k) return gx_fail(h, GVDBX_E_UNSUPPORTED, "no kernel variant for this mode / sampler combination");
    dim3 block(h->block_w, h->block_h, 1);
    dim3 grid((P.width + block.x - 1) / block.x, (P.height + block.y - 1) / block.y, 1);
    k<<<grid, block, size_t(block.x) * block.y * GX_STACK_BYTES_PER_THREAD, h->stream>>>(P);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

extern "C" int gvdbx_tiles_per_rank(int width, int height, int tile_size, int nranks)
{
    if (width <= 0 || height <= 0 || tile_size <= 0 || nranks <= 0) return GVDBX_E_ARG;
    const int tiles = ((width + tile_size - 1) / tile_size) * ((height + tile_size - 1) / tile_size);
    return (tiles + nranks - 1) / nranks;
}

// tile-list launch: the tiles (tile_w x tile_h pixels, numbered row-major) with id % nranks == rank.  direct: every pixel at
// its place in a row-major frame (possibly a peer GPU's: the stores travel over NVLink from inside the render kernel and no
// gather / assemble step exists); otherwise packed tile after tile into this rank's own buffer.
static int gx_render_tile_list(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t out_d, int tile_w, int tile_h,
                               int rank, int nranks, bool direct)
{
    if (!h) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GxParams P; int mode = 0;
    int rc = gx_fill_params(h, scninfo, shade_mode, chan, P, mode);
    if (rc) return rc;
    if (!out_d || nranks <= 0 || rank < 0 || rank >= nranks) return gx_fail(h, GVDBX_E_ARG, "rank/nranks/buffer");
    if (tile_w <= 0 || tile_h <= 0 || tile_w % h->block_w || tile_h % h->block_h)
        return gx_fail(h, GVDBX_E_ARG, "tile size must be a multiple of the CTA tile");
    P.out = (uchar4*)out_d;
    P.out_stride = direct ? P.width : 0;                // > 0 selects direct addressing in the tile-list kernels
    P.tile_w = tile_w; P.tile_h = tile_h;
    P.tiles_x = (P.width + tile_w - 1) / tile_w;
    P.ntiles = P.tiles_x * ((P.height + tile_h - 1) / tile_h);
    P.rank = rank; P.nranks = nranks;
    const int slots = (P.ntiles + nranks - 1) / nranks;
    const int flags = GX_FLAG_TILES | (h->spp > 1 ? GX_FLAG_SPP : 0) | gx_queue_flag(h, mode);
    gx_kernel_t k = gx_pick(mode, h->sampler, flags, h->uniform3);
    if (!k) return gx_fail(h, GVDBX_E_UNSUPPORTED, "no kernel variant for this mode / sampler combination");
    dim3 block(h->block_w, h->block_h, 1);
    dim3 grid((tile_w / h->block_w) * (tile_h / h->block_h), slots, 1);
    k<<<grid, block, gx_smem_bytes(block, flags), h->stream>>>(P);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

extern "C" int gvdbx_render_tiles(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t packed_d,
                                  int tile_size, int rank, int nranks)
{
    return gx_render_tile_list(h, scninfo, shade_mode, chan, packed_d, tile_size, tile_size, rank, nranks, false);
}

extern "C" int gvdbx_render_tiles_direct(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t frame_d,
                                         int tile_size, int rank, int nranks)
{
    return gx_render_tile_list(h, scninfo, shade_mode, chan, frame_d, tile_size, tile_size, rank, nranks, true);
}

// Full-width bands of `band_rows` rows, band b of the frame owned by rank b % nranks, packed band after band with a row
// pitch of gvdbx_band_pitch(width) pixels: every band is one contiguous block of rows, so a rank can copy its share of the
// frame to the host with one 2-D copy per band (the host frame ring, gvdbx_hostring_*).
static int gx_band_pitch(const gvdbx_t* h, int width) { return (width + h->block_w - 1) / h->block_w * h->block_w; }
extern "C" int gvdbx_render_bands(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t packed_d, int band_rows,
                                  int rank, int nranks)
{
    if (!h || !scninfo) return GVDBX_E_ARG;
    int width = 0;
    memcpy(&width, scninfo, sizeof width);              // ScnInfo.width @ 0
    return gx_render_tile_list(h, scninfo, shade_mode, chan, packed_d, gx_band_pitch(h, width), band_rows, rank, nranks, false);
}

// One call per frame and rank for the peer frame ring: [wait until *wait_flag_d >= wait_value] -> this rank's tiles into
// frame_d -> *done_flag_d += 1, all on the context's current stream / lane (wait_flag_d == 0: no wait).
extern "C" int gvdbx_render_tiles_ring(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t frame_d,
                                       int tile_size, int rank, int nranks, uint64_t wait_flag_d, uint32_t wait_value,
                                       uint64_t done_flag_d)
{
    if (!h || !done_flag_d) return GVDBX_E_ARG;
    int rc = GVDBX_OK;
    if (wait_flag_d) rc = gvdbx_stream_wait(h, nullptr, wait_flag_d, wait_value);
    if (rc == GVDBX_OK) rc = gvdbx_render_tiles_direct(h, scninfo, shade_mode, chan, frame_d, tile_size, rank, nranks);
    if (rc == GVDBX_OK) rc = gvdbx_stream_signal_add(h, nullptr, done_flag_d, 1);
    return rc;
}

// ------------------------------------------------------------------------------------------------ peer memory + flags
extern "C" int gvdbx_peer_alloc(gvdbx_t* h, size_t bytes, uint64_t* dptr, void* handle64)
{
    if (!h || !dptr || !handle64 || bytes == 0) return GVDBX_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == GVDBX_IPC_HANDLE_BYTES, "IPC handle size");
    GxCtx ctx_(h);
    void* p = nullptr;
    GX_CUDA(h, cudaMalloc(&p, bytes));
    GX_CUDA(h, cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) { cudaFree(p); h->err = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); return GVDBX_E_CUDA; }
    memcpy(handle64, &hd, sizeof hd);
    *dptr = (uint64_t)p;
    return GVDBX_OK;
}
extern "C" int gvdbx_peer_free(gvdbx_t* h, uint64_t dptr)
{
    if (!h || !dptr) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GX_CUDA(h, cudaFree((void*)dptr));
    return GVDBX_OK;
}
extern "C" int gvdbx_peer_open(gvdbx_t* h, const void* handle64, uint64_t* dptr)
{
    if (!h || !dptr || !handle64) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, sizeof hd);
    void* p = nullptr;
    GX_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    *dptr = (uint64_t)p;
    return GVDBX_OK;
}
extern "C" int gvdbx_peer_close(gvdbx_t* h, uint64_t dptr)
{
    if (!h || !dptr) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GX_CUDA(h, cudaIpcCloseMemHandle((void*)dptr));
    return GVDBX_OK;
}

extern "C" int gvdbx_set_stream(gvdbx_t* h, void* cuda_stream)
{
    if (!h) return GVDBX_E_ARG;
    h->stream = (cudaStream_t)cuda_stream;
    h->base_stream = h->stream;
    return GVDBX_OK;
}

// ------------------------------------------------------------------------------------------------ frame lanes
// A frame's kernel ends with a tail of a few long rays during which most SMs idle (1080p: ~15 % of the frame; a rank of
// an 8-GPU run renders 1/8 of the rays and is ALL tail).  Rendering consecutive frames on alternating streams lets the
// next frame fill those SMs.  The library owns the streams; the caller owns one output buffer per lane.
extern "C" int gvdbx_lanes(gvdbx_t* h, int n)
{
    if (!h || n < 0 || n > 16) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    for (cudaStream_t s : h->lanes) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (cudaEvent_t e : h->lane_ev) cudaEventDestroy(e);
    if (h->base_ev) { cudaEventDestroy(h->base_ev); h->base_ev = nullptr; }
    h->lanes.clear(); h->lane_ev.clear();
    h->stream = h->base_stream;
    h->cur_lane = -1;
    for (int i = 0; i < n; i++) {
        cudaStream_t s; cudaEvent_t e;
        GX_CUDA(h, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        GX_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->lanes.push_back(s); h->lane_ev.push_back(e);
    }
    if (n) GX_CUDA(h, cudaEventCreateWithFlags(&h->base_ev, cudaEventDisableTiming));
    return GVDBX_OK;
}
extern "C" int gvdbx_lane_select(gvdbx_t* h, int lane)
{
    if (!h) return GVDBX_E_ARG;
    if (lane < 0 || h->lanes.empty()) { h->stream = h->base_stream; h->cur_lane = -1; return GVDBX_OK; }
    h->cur_lane = lane % (int)h->lanes.size();
    h->stream = h->lanes[h->cur_lane];
    return GVDBX_OK;
}
extern "C" void* gvdbx_lane_stream(gvdbx_t* h, int lane)
{
    if (!h || lane < 0 || h->lanes.empty()) return h ? (void*)h->base_stream : nullptr;
    return (void*)h->lanes[lane % (int)h->lanes.size()];
}
// every lane waits for what the creation stream has enqueued so far (inputs ready); stream-ordered, no host sync
extern "C" int gvdbx_lanes_fork(gvdbx_t* h)
{
    if (!h) return GVDBX_E_ARG;
    if (h->lanes.empty()) return GVDBX_OK;
    GxCtx ctx_(h);
    GX_CUDA(h, cudaEventRecord(h->base_ev, h->base_stream));
    for (cudaStream_t s : h->lanes) GX_CUDA(h, cudaStreamWaitEvent(s, h->base_ev, 0));
    return GVDBX_OK;
}
// the creation stream waits for everything enqueued on the lanes so far; selects the creation stream again
extern "C" int gvdbx_lanes_join(gvdbx_t* h)
{
    if (!h) return GVDBX_E_ARG;
    h->stream = h->base_stream;
    h->cur_lane = -1;
    if (h->lanes.empty()) return GVDBX_OK;
    GxCtx ctx_(h);
    for (size_t i = 0; i < h->lanes.size(); i++) {
        GX_CUDA(h, cudaEventRecord(h->lane_ev[i], h->lanes[i]));
        GX_CUDA(h, cudaStreamWaitEvent(h->base_stream, h->lane_ev[i], 0));
    }
    return GVDBX_OK;
}

// stream-ordered flag write: everything enqueued before it on `stream` (incl. stores to peer memory) is complete and
// visible system-wide before the flag changes
extern "C" int gvdbx_stream_signal(gvdbx_t* h, void* cuda_stream, uint64_t flag_d, uint32_t value)
{
    if (!h || !flag_d) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    gx_signal_kernel<<<1, 1, 0, st>>>((unsigned int*)flag_d, value);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

// same, but adds `inc` to a counter (several producers share one counter: the waiter targets producers * uses)
extern "C" int gvdbx_stream_signal_add(gvdbx_t* h, void* cuda_stream, uint64_t flag_d, uint32_t inc)
{
    if (!h || !flag_d) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    gx_signal_add_kernel<<<1, 1, 0, st>>>((unsigned int*)flag_d, inc);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}
// same value to up to 16 flags (one per peer) with one launch
extern "C" int gvdbx_stream_signal_many(gvdbx_t* h, void* cuda_stream, const uint64_t* flags_d, int n, uint32_t value)
{
    if (!h || !flags_d || n <= 0 || n > 16) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    GxFlagList L;
    for (int i = 0; i < 16; i++) L.p[i] = i < n ? (unsigned int*)flags_d[i] : nullptr;
    gx_signal_many_kernel<<<1, 16, 0, st>>>(L, n, value);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

// stream-ordered wait until *flag >= value (flag in LOCAL device memory).  Default: a one-thread polling kernel that gives
// up after ~20 s (flag[1] = 0xDEAD) so that a lost peer or a mis-ordered submission cannot wedge the GPU.  With
// GVDBX_OPT_STREAM_MEMOPS the driver's cuStreamWaitValue32 is used instead (no SM involved, but unbounded).
// Submission-order rule either way: when waiter and signaller are streams of the SAME process, enqueue the signal first —
// streams can share a hardware queue, where a wait enqueued ahead of its signal would never be passed.
typedef int (*gx_cuStreamWaitValue32_t)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
extern "C" int gvdbx_stream_wait(gvdbx_t* h, void* cuda_stream, uint64_t flag_d, uint32_t value)
{
    if (!h || !flag_d) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (h->memops) {
        static gx_cuStreamWaitValue32_t fn = nullptr;
        static bool looked = false;
        if (!looked) {
            looked = true;
            void* p = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
                fn = (gx_cuStreamWaitValue32_t)p;
            cudaGetLastError();
        }
        if (fn && fn(st, (unsigned long long)flag_d, value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */) == 0) return GVDBX_OK;
    }
    gx_wait_kernel<<<1, 1, 0, st>>>((unsigned int*)flag_d, value, h->d_err);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

extern "C" int gvdbx_assemble_tiles(gvdbx_t* h, uint64_t gathered_d, uint64_t frame_d, int width, int height, int tile_size, int nranks)
{
    if (!h || !gathered_d || !frame_d || width <= 0 || height <= 0 || tile_size <= 0 || nranks <= 0) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    const int tiles_x = (width + tile_size - 1) / tile_size;
    const int ntiles = tiles_x * ((height + tile_size - 1) / tile_size);
    const int slots = (ntiles + nranks - 1) / nranks;
    dim3 grid((tile_size * tile_size + 255) / 256, ntiles, 1);
    gx_assemble_tiles<<<grid, 256, 0, h->stream>>>((const uchar4*)gathered_d, (uchar4*)frame_d, width, height, tile_size, tiles_x, ntiles, nranks, slots);
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

// VolumeGVDB::Raytrace (gvdb_volume_gvdb.cpp:4384-4408): `rays_d` = device array of n 64-byte ScnRay records, updated in
// place (hit, normal).  ScnInfo supplies steps / thresholds exactly like PrepareRender(1,1,0) does for the reference.
extern "C" int gvdbx_raytrace(gvdbx_t* h, const void* scninfo, int chan, uint64_t rays_d, int num_rays, float bias)
{
    if (!h) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GxParams P; int mode = 0;
    int rc = gx_fill_params(h, scninfo, GVDBX_SHADE_TRILINEAR, chan, P, mode);
    if (rc) return rc;
    if (!rays_d || num_rays <= 0) return gx_fail(h, GVDBX_E_ARG, "rays/num_rays");
    P.dbuf = nullptr;                                   // per-pixel depth buffers do not apply to ray bundles
    const unsigned blocks = (num_rays + 63) / 64;       // 64-thread CTAs like the reference launch
    if (h->sampler == GX_SAMPLER_TEX) {
        if (h->uniform3) gx_raytrace_kernel<GX_SAMPLER_TEX, true><<<blocks, 64, 64 * GX_STACK_BYTES_PER_THREAD, h->stream>>>(P, (float*)rays_d, num_rays, bias);
        else             gx_raytrace_kernel<GX_SAMPLER_TEX, false><<<blocks, 64, 64 * GX_STACK_BYTES_PER_THREAD, h->stream>>>(P, (float*)rays_d, num_rays, bias);
    } else {
        if (h->uniform3) gx_raytrace_kernel<GX_SAMPLER_LINEAR, true><<<blocks, 64, 64 * GX_STACK_BYTES_PER_THREAD, h->stream>>>(P, (float*)rays_d, num_rays, bias);
        else             gx_raytrace_kernel<GX_SAMPLER_LINEAR, false><<<blocks, 64, 64 * GX_STACK_BYTES_PER_THREAD, h->stream>>>(P, (float*)rays_d, num_rays, bias);
    }
    GX_CUDA(h, cudaGetLastError());
    return GVDBX_OK;
}

extern "C" int gvdbx_read_buffer(gvdbx_t* h, uint64_t buf_d, void* host, size_t bytes)
{
    if (!h || !buf_d || !host) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GX_CUDA(h, cudaMemcpyAsync(host, (const void*)buf_d, bytes, cudaMemcpyDeviceToHost, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));
    return GVDBX_OK;
}

// ------------------------------------------------------------------------------------------------ banded render + overlapped read-back
// The strict drop-in sequence is Render() (asynchronous) followed by ReadRenderBuf() (synchronous copy into the caller's pageable
// memory): kernel and copy in series, 33 MB per 4K frame.  gvdbx_render_banded renders the same frame as horizontal bands,
// consecutive bands alternating between two internal streams (the tail of band b overlaps band b + 1, like the frame lanes) with an
// event behind each; gvdbx_read_banded copies band after band as they finish, so all but the last band's copy hides behind the
// rendering of the bands below it.  Same bytes in the same buffer; the caller's sequence does not change.
static int gx_band_setup(gvdbx_t* h)
{
    if (h->band_streams[0]) return GVDBX_OK;
    for (int k = 0; k < 2; k++) {
        GX_CUDA(h, cudaStreamCreateWithFlags(&h->band_streams[k], cudaStreamNonBlocking));
        GX_CUDA(h, cudaEventCreateWithFlags(&h->band_join[k], cudaEventDisableTiming));
    }
    GX_CUDA(h, cudaStreamCreateWithFlags(&h->band_copy, cudaStreamNonBlocking));
    GX_CUDA(h, cudaEventCreateWithFlags(&h->band_fork, cudaEventDisableTiming));
    for (int b = 0; b < GX_MAX_BANDS; b++) GX_CUDA(h, cudaEventCreateWithFlags(&h->band_ev[b], cudaEventDisableTiming));
    return GVDBX_OK;
}

extern "C" int gvdbx_render_banded(gvdbx_t* h, const void* scninfo, int shade_mode, int chan, uint64_t outbuf_d, int nbands)
{
    if (!h || !scninfo || !outbuf_d || nbands < 1 || nbands > GX_MAX_BANDS) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GxScnInfo s;
    memcpy(&s, scninfo, sizeof s);
    h->band_n = 0;
    if (shade_mode == GVDBX_SHADE_OFF || nbands == 1 || s.height < nbands * h->block_h || h->count)      // work counters are per launch
        return gvdbx_render(h, scninfo, shade_mode, chan, outbuf_d, 0, 0, 0, 0);
    int rc = gx_band_setup(h);
    if (rc) return rc;
    // whatever the frame needs built (occupancy bits for a new THRESH, brick-major copy) is enqueued on the caller's stream first:
    // the band streams fork behind it
    { GxParams P; int mode = 0; rc = gx_fill_params(h, scninfo, shade_mode, chan, P, mode); if (rc) return rc; }
    cudaStream_t saved_stream = h->stream;
    const int saved_lane = h->cur_lane;
    GX_CUDA(h, cudaEventRecord(h->band_fork, saved_stream));
    for (int k = 0; k < 2; k++) GX_CUDA(h, cudaStreamWaitEvent(h->band_streams[k], h->band_fork, 0));
    const int rows = ((s.height + nbands - 1) / nbands + h->block_h - 1) / h->block_h * h->block_h;
    int nb = 0;
    for (int y0 = 0; y0 < s.height && rc == GVDBX_OK; y0 += rows, nb++) {
        const int k = nb & 1;
        h->stream = h->band_streams[k];
        h->cur_lane = int(h->lanes.size()) + k;            // private slots for the per-frame derived tables of the two band streams
        rc = gvdbx_render(h, scninfo, shade_mode, chan, outbuf_d, 0, y0, s.width, std::min(rows, s.height - y0));
        if (rc == GVDBX_OK && cudaEventRecord(h->band_ev[nb], h->band_streams[k]) != cudaSuccess) rc = GVDBX_E_CUDA;
    }
    h->stream = saved_stream;
    h->cur_lane = saved_lane;
    for (int k = 0; k < 2; k++) {                           // later work on the caller's stream is ordered behind every band
        GX_CUDA(h, cudaEventRecord(h->band_join[k], h->band_streams[k]));
        GX_CUDA(h, cudaStreamWaitEvent(saved_stream, h->band_join[k], 0));
    }
    if (rc) return rc;
    h->band_n = nb; h->band_rows = rows; h->band_w = s.width; h->band_h = s.height; h->band_buf = outbuf_d;
    return GVDBX_OK;
}

extern "C" int gvdbx_read_banded(gvdbx_t* h, uint64_t buf_d, void* host, size_t bytes)
{
    if (!h || !buf_d || !host) return GVDBX_E_ARG;
    if (h->band_n == 0 || h->band_buf != buf_d || bytes != size_t(h->band_w) * h->band_h * 4) return gvdbx_read_buffer(h, buf_d, host, bytes);
    GxCtx ctx_(h);
    for (int b = 0; b < h->band_n; b++) {
        const int y0 = b * h->band_rows, rows = std::min(h->band_rows, h->band_h - y0);
        const size_t off = size_t(y0) * h->band_w * 4, n = size_t(rows) * h->band_w * 4;
        GX_CUDA(h, cudaEventSynchronize(h->band_ev[b]));
        GX_CUDA(h, cudaMemcpyAsync((char*)host + off, (const char*)buf_d + off, n, cudaMemcpyDeviceToHost, h->band_copy));
        GX_CUDA(h, cudaStreamSynchronize(h->band_copy));
    }
    return GVDBX_OK;
}

// same copy without the synchronisation: the host buffer (pinned, for a truly asynchronous copy) is valid once the stream
// the context currently uses has passed this point (gvdbx_sync, or an event of the caller's)
extern "C" int gvdbx_read_buffer_async(gvdbx_t* h, uint64_t buf_d, void* host, size_t bytes)
{
    if (!h || !buf_d || !host) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GX_CUDA(h, cudaMemcpyAsync(host, (const void*)buf_d, bytes, cudaMemcpyDeviceToHost, h->stream));
    return GVDBX_OK;
}

extern "C" int gvdbx_sync(gvdbx_t* h)
{
    if (!h) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    GX_CUDA(h, cudaStreamSynchronize(h->stream));
    // sticky: a stream-ordered wait that ran into its timeout means frames were rendered / consumed out of order
    int e = 0;
    GX_CUDA(h, cudaMemcpy(&e, h->d_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e & GX_ERR_WAIT_TIMEOUT) return gx_fail(h, GVDBX_E_CUDA, "a gvdbx_stream_wait timed out (lost peer or stalled consumer): frames after it are not trustworthy");
    return GVDBX_OK;
}

extern "C" int gvdbx_get_counters(gvdbx_t* h, gvdbx_counters* out)
{
    if (!h || !out) return GVDBX_E_ARG;
    GxCtx ctx_(h);
    unsigned long long v[8];
    GX_CUDA(h, cudaMemcpyAsync(v, h->d_counters, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    GX_CUDA(h, cudaStreamSynchronize(h->stream));
    out->s_tri = v[0]; out->s_pt = v[1]; out->n_dda = v[2]; out->n_desc = v[3]; out->s_lut = v[4]; out->rays = v[5];
    return GVDBX_OK;
}

extern "C" int gvdbx_sample_points(gvdbx_t* h, int chan, uint64_t xyz_d, int n, uint64_t out_tex_d, uint64_t out_lin_d)
{
    if (!h || !xyz_d || !out_tex_d || !out_lin_d || n <= 0) return GVDBX_E_ARG;
    if (!h->have_atlas || !h->have_topo) return gx_fail(h, GVDBX_E_STATE, "needs topology and atlas");
    if (chan != 0) return gx_fail(h, GVDBX_E_UNSUPPORTED, "only channel 0 is supported");
    GxCtx ctx_(h);
    int rc = gx_ensure_bricks(h);
    if (rc) return rc;
    GxParams P;
    gx_tree_params(h, P);
    const int cx = h->ares[0] / h->brick_dim, cy = h->ares[1] / h->brick_dim, cz = h->ares[2] / h->brick_dim;
    int* slot_leaf = nullptr;
    GX_CUDA(h, cudaMalloc(&slot_leaf, size_t(cx) * cy * cz * sizeof(int)));
    GX_CUDA(h, cudaMemsetAsync(slot_leaf, 0xFF, size_t(cx) * cy * cz * sizeof(int), h->stream));
    const int nl = h->vdb.nodecnt[0];
    gx_build_slot_map<<<(nl + 255) / 256, 256, 0, h->stream>>>(h->d_leaf, nl, h->brick_dim, cx, cy, slot_leaf);
    gx_sample_points_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(P, (const float*)xyz_d, n, cx, cy, slot_leaf, (float*)out_tex_d, (float*)out_lin_d);
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(h->stream);
    cudaFree(slot_leaf);
    GX_CUDA(h, e);
    return GVDBX_OK;
}

// fp32 trilinear samples per second of this GPU's texture units on L1-resident bricks of the imported atlas (synchronises)
extern "C" int gvdbx_measure_tex_peak(gvdbx_t* h, float lane_spacing, double* gsamples_per_s)
{
    if (!h || !gsamples_per_s || !(lane_spacing >= 0.f) || lane_spacing > 8.f) return GVDBX_E_ARG;
    if (!h->have_atlas) return gx_fail(h, GVDBX_E_STATE, "no atlas imported");
    GxCtx ctx_(h);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int blocks = sms * 8, threads = 256, rounds = 512;
    float* out = nullptr;
    GX_CUDA(h, cudaMalloc(&out, size_t(blocks) * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    GX_CUDA(h, cudaEventCreate(&e0));
    GX_CUDA(h, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {             // first repetition warms up
        GX_CUDA(h, cudaEventRecord(e0, h->stream));
        gx_tex_peak_kernel<<<blocks, threads, 0, h->stream>>>(h->tex, h->ares[0], h->ares[1], h->ares[2], rounds, lane_spacing, out);
        GX_CUDA(h, cudaEventRecord(e1, h->stream));
        GX_CUDA(h, cudaEventSynchronize(e1));
        float ms = 0.f;
        GX_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        const double g = double(blocks) * threads * rounds * 8.0 / (double(ms) * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *gsamples_per_s = best;
    return GVDBX_OK;
}

// The four ways of reading a brick (csrc/gvdbx_microbench.cuh), fetch + filter only, on the imported atlas: Gsamples/s of
// [0] the texture unit, [1] brick-major copy + scalar loads, [2] x-pair layout + 8-byte loads, [3] TMA-staged shared memory.
// Builds the brick-major copy if it does not exist yet (synchronises).
extern "C" int gvdbx_measure_sampler_ab(gvdbx_t* h, float lane_spacing, double* gsamples_per_s4)
{
    if (!h || !gsamples_per_s4 || !(lane_spacing >= 0.f) || lane_spacing > 8.f) return GVDBX_E_ARG;
    if (!h->have_atlas || !h->have_topo) return gx_fail(h, GVDBX_E_STATE, "needs topology and atlas");
    if (h->brick_dim != GX_BRICK_DIM) return gx_fail(h, GVDBX_E_UNSUPPORTED, "the sampler A/B microbenchmark is written for 8^3 bricks");
    int rc = gvdbx_measure_tex_peak(h, lane_spacing, &gsamples_per_s4[0]);
    if (rc) return rc;
    GxCtx ctx_(h);
    rc = gx_ensure_bricks(h);
    if (rc) return rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int nb = h->vdb.nodecnt[0] < 4096 ? h->vdb.nodecnt[0] : 4096;
    const int blocks = sms * 8, threads = 128, visits = 64;
    float* out = nullptr; float2* pairs = nullptr;
    GX_CUDA(h, cudaMalloc(&out, size_t(blocks) * threads * sizeof(float)));
    GX_CUDA(h, cudaMalloc(&pairs, size_t(nb) * GX_BRICK_STRIDE * sizeof(float2)));
    gx_build_pairs<<<(nb * 1000 + 255) / 256, 256, 0, h->stream>>>(h->d_bricks, pairs, nb);
    const size_t smem3 = size_t(threads / 32) * 2 * GX_BRICK_STRIDE * sizeof(float) + size_t(threads / 32) * 2 * sizeof(unsigned long long);
    cudaEvent_t e0, e1;
    GX_CUDA(h, cudaEventCreate(&e0));
    GX_CUDA(h, cudaEventCreate(&e1));
    for (int variant = 1; variant <= 3; variant++) {
        double best = 0.0;
        for (int rep = 0; rep < 4; rep++) {             // first repetition warms up
            GX_CUDA(h, cudaEventRecord(e0, h->stream));
            if (variant == 1)      gx_linear_peak_kernel<1><<<blocks, threads, 0, h->stream>>>(h->d_bricks, pairs, nb, visits, lane_spacing, out);
            else if (variant == 2) gx_linear_peak_kernel<2><<<blocks, threads, 0, h->stream>>>(h->d_bricks, pairs, nb, visits, lane_spacing, out);
            else                   gx_linear_peak_kernel<3><<<blocks, threads, smem3, h->stream>>>(h->d_bricks, pairs, nb, visits, lane_spacing, out);
            GX_CUDA(h, cudaEventRecord(e1, h->stream));
            GX_CUDA(h, cudaEventSynchronize(e1));
            GX_CUDA(h, cudaGetLastError());
            float ms = 0.f;
            GX_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
            const double g = double(blocks) * threads * visits * GX_MB_SAMPLES / (double(ms) * 1e-3) / 1e9;
            if (rep > 0 && g > best) best = g;
        }
        gsamples_per_s4[variant] = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out); cudaFree(pairs);
    return GVDBX_OK;
}

// Gsamples/s of the deep marcher's inner loop alone (csrc/gvdbx_microbench.cuh): fetch + transfer index + table gather + colour
// update, four samples per round, no traversal.  Needs the atlas and a transfer function (gvdbx_set_transfer); synchronises.
// table_through_texture = 1 reads the transfer table through a float4 texture object instead of 16-byte loads (measured slower:
// 274 against 349 Gsamples/s — the production kernels use the loads).
extern "C" int gvdbx_measure_deep_loop_peak(gvdbx_t* h, float lane_spacing, int table_through_texture, double* gsamples_per_s)
{
    if (!h || !gsamples_per_s || !(lane_spacing >= 0.f) || lane_spacing > 8.f) return GVDBX_E_ARG;
    if (!h->have_atlas) return gx_fail(h, GVDBX_E_STATE, "no atlas imported");
    if (!h->d_transfer) return gx_fail(h, GVDBX_E_STATE, "no transfer function (gvdbx_set_transfer)");
    GxCtx ctx_(h);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int blocks = sms * 14, threads = 128, rounds = 1024;
    float* out = nullptr;
    GX_CUDA(h, cudaMalloc(&out, size_t(blocks) * threads * sizeof(float)));
    cudaTextureObject_t lut_tex = 0;
    if (table_through_texture) {                    // A/B: table entries fetched through a linear float4 texture instead of 16-byte loads
        cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = h->d_transfer;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = GVDBX_TRANSFER_ENTRIES * sizeof(float4);
        cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; td.filterMode = cudaFilterModePoint; td.addressMode[0] = cudaAddressModeClamp;
        GX_CUDA(h, cudaCreateTextureObject(&lut_tex, &rd, &td, nullptr));
    }
    cudaEvent_t e0, e1;
    GX_CUDA(h, cudaEventCreate(&e0));
    GX_CUDA(h, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {             // first repetition warms up
        GX_CUDA(h, cudaEventRecord(e0, h->stream));
        if (!lut_tex) gx_deep_loop_peak_kernel<0><<<blocks, threads, 0, h->stream>>>(h->tex, h->d_transfer, 0, h->ares[0], h->ares[1], h->ares[2], rounds, lane_spacing,
                                                                                    0.1f, 1.0f, 0.005f, 1.5f, out);
        else          gx_deep_loop_peak_kernel<1><<<blocks, threads, 0, h->stream>>>(h->tex, h->d_transfer, lut_tex, h->ares[0], h->ares[1], h->ares[2], rounds, lane_spacing,
                                                                                    0.1f, 1.0f, 0.005f, 1.5f, out);
        GX_CUDA(h, cudaEventRecord(e1, h->stream));
        GX_CUDA(h, cudaEventSynchronize(e1));
        GX_CUDA(h, cudaGetLastError());
        float ms = 0.f;
        GX_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        const double g = double(blocks) * threads * rounds * 4.0 / (double(ms) * 1e-3) / 1e9;
        if (rep > 0 && g > best) best = g;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (lut_tex) cudaDestroyTextureObject(lut_tex);
    cudaFree(out);
    *gsamples_per_s = best;
    return GVDBX_OK;
}

// device-buffer helpers for the host mirror (gvdbx_host.cpp is plain C++)
extern "C" int gvdbx_internal_alloc(gvdbx_t* h, uint64_t* ptr, size_t bytes)
{
    GxCtx ctx_(h);
    void* d = nullptr;
    if (!ptr || cudaMalloc(&d, bytes) != cudaSuccess) return GVDBX_E_CUDA;
    *ptr = (uint64_t)d;
    return GVDBX_OK;
}
extern "C" int gvdbx_internal_free(gvdbx_t* h, uint64_t ptr)
{
    GxCtx ctx_(h);
    return cudaFree((void*)ptr) == cudaSuccess ? GVDBX_OK : GVDBX_E_CUDA;
}
