// gvdbx_types.h — byte layouts of the reference's interface structs, as seen by this library.
//
// These are the *wire formats* of the drop-in boundary (north_star: "same node-pool and brick-atlas layout"):
//   GxVDBInfo  <-> VDBInfo  (kernels/cuda_gvdb_nodes.cuh:42-67, host twin src/gvdb_volume_gvdb.h:65-90), 1232 B
//   GxScnInfo  <-> ScnInfo  (kernels/cuda_gvdb_scene.cuh:35-64, host twin src/gvdb_volume_gvdb.h:92-121), 416 B
//   GxNode     <-> VDBNode / Node (kernels/cuda_gvdb_nodes.cuh:24-35, src/gvdb_node.h:27-40), 64 B
// Offsets are asserted below against the values probed from the reference build.
#pragma once
#include <stdint.h>
#include <stddef.h>

struct GxF3 { float x, y, z; };
struct GxI3 { int x, y, z; };
struct GxF4 { float x, y, z, w; };

#define GX_MAXLEV       5          // kernels/cuda_gvdb_raycast.cuh:23
#define GX_MAX_ITER     256        // kernels/cuda_gvdb_raycast.cuh:24
#define GX_NOHIT        1.0e10f    // kernels/cuda_gvdb_scene.cuh:20
#define GX_ID_UNDEFL    0xFFFFFFFFull
#define GX_CHAN_UNDEF   255

struct alignas(16) GxVDBInfo {
    int       dim[10];
    int       res[10];
    GxF3      vdel[10];
    GxI3      noderange[10];
    int       nodecnt[10];
    int       nodewid[10];
    int       childwid[10];
    uint64_t  nodelist[10];
    uint64_t  childlist[10];
    uint64_t  atlas_map;
    GxI3      atlas_cnt;
    GxI3      atlas_res;
    int       atlas_apron;
    int       brick_res;
    int       apron_table[8];
    int       top_lev;
    int       max_iter;
    float     epsilon;
    uint8_t   update;
    uint8_t   clr_chan;
    GxF3      bmin;
    GxF3      bmax;
    uint64_t  volIn[32];
    uint64_t  volOut[32];
};
static_assert(sizeof(GxVDBInfo) == 1232, "VDBInfo size");
static_assert(offsetof(GxVDBInfo, vdel) == 80 && offsetof(GxVDBInfo, noderange) == 200, "VDBInfo layout");
static_assert(offsetof(GxVDBInfo, nodelist) == 440 && offsetof(GxVDBInfo, childlist) == 520, "VDBInfo layout");
static_assert(offsetof(GxVDBInfo, atlas_map) == 600 && offsetof(GxVDBInfo, atlas_cnt) == 608, "VDBInfo layout");
static_assert(offsetof(GxVDBInfo, top_lev) == 672 && offsetof(GxVDBInfo, epsilon) == 680, "VDBInfo layout");
static_assert(offsetof(GxVDBInfo, clr_chan) == 685 && offsetof(GxVDBInfo, bmin) == 688, "VDBInfo layout");
static_assert(offsetof(GxVDBInfo, volIn) == 712 && offsetof(GxVDBInfo, volOut) == 968, "VDBInfo layout");

struct alignas(16) GxScnInfo {
    int       width, height;
    float     camnear, camfar;
    GxF3      campos, cams, camu, camv;
    GxF3      light_pos, slice_pnt, slice_norm, shadow_params;
    GxF4      backclr;
    float     xform[16], invxform[16], invxrot[16];
    float     bias;
    char      shading, filtering;
    int       frame, samples;
    GxF3      extinct, steps, cutoff, thresh;
    uint64_t  transfer;
    uint64_t  outbuf;
    uint64_t  dbuf;
};
static_assert(sizeof(GxScnInfo) == 416, "ScnInfo size");
static_assert(offsetof(GxScnInfo, campos) == 16 && offsetof(GxScnInfo, light_pos) == 64, "ScnInfo layout");
static_assert(offsetof(GxScnInfo, backclr) == 112 && offsetof(GxScnInfo, xform) == 128, "ScnInfo layout");
static_assert(offsetof(GxScnInfo, bias) == 320 && offsetof(GxScnInfo, frame) == 328, "ScnInfo layout");
static_assert(offsetof(GxScnInfo, extinct) == 336 && offsetof(GxScnInfo, thresh) == 372, "ScnInfo layout");
static_assert(offsetof(GxScnInfo, transfer) == 384 && offsetof(GxScnInfo, dbuf) == 400, "ScnInfo layout");

struct alignas(16) GxNode {
    uint8_t   mLev, mFlags, mPriority, pad;
    GxI3      mPos;
    GxI3      mValue;
    GxF3      mVRange;
    uint64_t  mParent;
    uint64_t  mChildList;
    uint64_t  mMask;
};
static_assert(sizeof(GxNode) == 64, "Node size");
static_assert(offsetof(GxNode, mPos) == 4 && offsetof(GxNode, mValue) == 16, "Node layout");
static_assert(offsetof(GxNode, mParent) == 40 && offsetof(GxNode, mChildList) == 48, "Node layout");
