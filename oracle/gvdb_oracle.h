/* gvdb_oracle.h — CPU restatement of the reference's algorithms for the render hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or executed by the product
 * (gvdb-voxels_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Pinning status: the integer side (topology pools, atlas slot assignment, apron contents, VDBInfo) is pinned
 * byte-for-byte against the UNMODIFIED reference library run on a B200 (oracle/_ref/ref_harness dumps; see
 * tests/golden/ and tests/test_oracle_vs_ref_golden.py).  The floating-point ray marcher can only be pinned to
 * tolerance, because the reference's own arithmetic uses GPU approximate instructions (--use_fast_math: MUFU
 * rcp/rsq/ex2) and the texture unit's fixed-point filtering: the bit-exact oracle for those is oracle/_ref running
 * on the GPU box (tests/refcmp.py).
 */
#ifndef GVDB_ORACLE_H
#define GVDB_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_tree ora_tree;

/* ---- topology: Configure / ActivateSpace / FinishTopology / UpdateAtlas (gvdb_volume_gvdb.cpp:2380, 2804, 1579, 2630) */
ora_tree* ora_tree_create(int levs, const int* logdim, const int* initcnt, int atlas_cx, int atlas_cy, int atlas_cz, int apron);
void      ora_tree_destroy(ora_tree* t);
int64_t   ora_activate_space(ora_tree* t, int x, int y, int z);   /* returns leaf index or -1 */
void      ora_activate_bricks(ora_tree* t, const int32_t* pos, int n);
void      ora_finish_topology(ora_tree* t);                        /* ComputeBounds */
void      ora_update_atlas(ora_tree* t);                           /* slot assignment + atlas map */
void      ora_set_epsilon(ora_tree* t, float eps, int maxiter);

/* pool access (reference layout: pool 0 = 64-B node records, pool 1 = dense 8-B child lists) */
uint64_t  ora_pool_count(const ora_tree* t, int grp, int lev);
uint64_t  ora_pool_width(const ora_tree* t, int grp, int lev);
const void* ora_pool_data(const ora_tree* t, int grp, int lev);
int       ora_num_levels(const ora_tree* t);
void      ora_atlas_res(const ora_tree* t, int res[3]);
const void* ora_atlas_map(const ora_tree* t, uint64_t* bytes);
/* PrepareVDB (gvdb_volume_gvdb.cpp:3946): fills the 1232-byte VDBInfo; pointer fields are left 0 */
void      ora_fill_vdbinfo(const ora_tree* t, void* vdbinfo1232);

/* ---- atlas: brick upload + UpdateApron semantics (cuda_gvdb_operators.cuh:72-126) */
/* writes the 512 interior values of every leaf (leaf n <- values + 512*n) into a zeroed atlas image, then fills aprons */
void      ora_fill_atlas(const ora_tree* t, const float* values, float* atlas, float boundval);

/* ---- CPU ray caster (cuda_gvdb_raycast.cuh / cuda_gvdb_dda.cuh / cuda_gvdb_geom.cuh / cuda_gvdb_module.cu:38-181) */
typedef struct ora_volume {
    const void*  vdbinfo;        /* 1232 B (host copy) */
    const void*  pool0[10];      /* node records per level */
    const void*  pool1[10];      /* child lists per level */
    const float* atlas;          /* x fastest */
    int          atlas_res[3];
    const float* transfer;       /* 16384 x float4 */
} ora_volume;

/* renders rows [y0,y1) of the frame described by the 416-byte ScnInfo; out_rgba is the full frame buffer
 * (width*height*4); hit_norm (optional, 8 floats per pixel: hit.xyz,0,norm.xyz,0) may be NULL.
 * threads <= 0: all OpenMP threads.  Returns 0 or -1 on unsupported input. */
int       ora_render(const ora_volume* v, const void* scninfo416, int shade, int y0, int y1,
                     uint8_t* out_rgba, float* hit_norm, int threads);
/* every shade mode of Render()'s switch (0 voxel, 1 section 2-D, 2 section 3-D, 3 empty skip, 4 trilinear, 5 tricubic,
 * 6 level set, 7 deep) + the compositions of BASELINE.json configs 4 / 5 (deep_shadow, spp rays per pixel) */
int       ora_render_ex(const ora_volume* v, const void* scninfo416, int shade, int y0, int y1,
                        uint8_t* rgba_out, float* hit_norm_out, int threads, int deep_shadow, int spp);
/* software model of the texture unit: trilinear fetch at atlas coordinate (x,y,z) */
float     ora_tex3d(const ora_volume* v, float x, float y, float z);
int       ora_max_threads(void);
void      ora_set_num_threads(int n);

/* ---- scenes (oracle/scenes.h) */
int       ora_scene_preset(const char* name, void* preset_out, size_t preset_bytes);
int       ora_scene_generate(const void* preset, int* nbricks, int32_t** brick_pos, float** values);
void      ora_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
