/* gvdb_oracle.c — CPU restatement of the reference's algorithms for the render hot path.
 * TEST INFRASTRUCTURE ONLY — see gvdb_oracle.h for the rules and the pinning status.
 *
 * Every function cites the reference code it restates (paths relative to /root/reference/source/gvdb_library/).
 * Plain C11, single precision where the reference is single precision, no FMA contraction (-ffp-contract=off).
 */
#include "gvdb_oracle.h"
#include "scenes.h"

/* the remaining values of the reference's SHADE_* enum (src/gvdb_types.h:105-112) */
#define ORA_SHADE_SECTION2D 1
#define ORA_SHADE_SECTION3D 2
#define ORA_SHADE_EMPTYSKIP 3
#define ORA_SHADE_TRICUBIC  5

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ============================================================================================ layouts */
typedef struct { int x, y, z; } i3;
typedef struct { float x, y, z; } f3;
typedef struct { float x, y, z, w; } f4;

/* Node: src/gvdb_node.h:27-40 (64 bytes) */
typedef struct {
    uint8_t  mLev, mFlags, mPriority, pad;
    i3       mPos;
    i3       mValue;
    f3       mVRange;
    uint64_t mParent;
    uint64_t mChildList;
    uint64_t mMask;
} ora_node;
_Static_assert(sizeof(ora_node) == 64, "node size");

/* AtlasNode: src/gvdb_volume_gvdb.h (mPos, mLeafNode), 16 bytes */
typedef struct { i3 mPos; int mLeafNode; } ora_atlas_node;

/* VDBInfo: src/gvdb_volume_gvdb.h:65-90 (1232 bytes) */
typedef struct {
    int      dim[10], res[10];
    f3       vdel[10];
    i3       noderange[10];
    int      nodecnt[10], nodewid[10], childwid[10];
    uint64_t nodelist[10], childlist[10];
    uint64_t atlas_map;
    i3       atlas_cnt, atlas_res;
    int      atlas_apron, brick_res;
    int      apron_table[8];
    int      top_lev, max_iter;
    float    epsilon;
    uint8_t  update, clr_chan;
    f3       bmin, bmax;
    uint64_t volIn[32], volOut[32];
} __attribute__((aligned(16))) ora_vdbinfo;
_Static_assert(sizeof(ora_vdbinfo) == 1232, "vdbinfo size");

/* ScnInfo: src/gvdb_volume_gvdb.h:92-121 (416 bytes) */
typedef struct {
    int      width, height;
    float    camnear, camfar;
    f3       campos, cams, camu, camv, light_pos, slice_pnt, slice_norm, shadow_params;
    f4       backclr;
    float    xform[16], invxform[16], invxrot[16];
    float    bias;
    char     shading, filtering;
    int      frame, samples;
    f3       extinct, steps, cutoff, thresh;
    uint64_t transfer, outbuf, dbuf;
} __attribute__((aligned(16))) ora_scninfo;
_Static_assert(sizeof(ora_scninfo) == 416, "scninfo size");

#define ID_UNDEFL  0xFFFFFFFFull
#define ID_UNDEF64 0xFFFFFFFFFFFFFFFFull
#define ORA_MAXLEV 10                       /* host MAXLEV, src/gvdb_volume_gvdb.h:40 */
#define NOHIT      1.0e10f

/* pool element reference: src/gvdb_allocator.h:59-62 */
static inline uint64_t Elem(uint64_t grp, uint64_t lev, uint64_t ndx) { return grp | (lev << 8) | (ndx << 16); }
static inline int      ElemLev(uint64_t e) { return (int)((e >> 8) & 0xFF); }
static inline uint64_t ElemNdx(uint64_t e) { return e >> 16; }

/* ============================================================================================ pools */
typedef struct { char* cpu; uint64_t lastEle, usedNum, max, stride, size; } ora_pool;

struct ora_tree {
    int       levs;
    int       logdim[ORA_MAXLEV];
    ora_pool  pool[2][ORA_MAXLEV];
    uint64_t  root;
    /* atlas bookkeeping (Allocator::mAtlas[0]) */
    i3        atlas_cnt;
    int       apron, leafdim;
    uint64_t  atlas_last, atlas_max;
    ora_atlas_node* amap; uint64_t amap_cnt;
    /* bounds, epsilon */
    f3        vmin, vmax;
    float     epsilon; int maxiter;
};

/* Allocator::PoolCreate, src/gvdb_allocator.cpp:55-91 */
static void pool_create(ora_pool* p, uint64_t width, uint64_t initmax)
{
    memset(p, 0, sizeof *p);
    p->max = initmax; p->stride = width; p->size = width * initmax;
    if (p->size) p->cpu = (char*)calloc(p->size, 1);
}
/* Allocator::PoolAlloc, src/gvdb_allocator.cpp:163-192 (doubling growth, new memory zeroed) */
static uint64_t pool_alloc(ora_tree* t, int grp, int lev)
{
    if (lev >= t->levs) return ID_UNDEFL;
    ora_pool* p = &t->pool[grp][lev];
    if (p->lastEle >= p->max) {
        p->max *= 2;
        p->size = p->stride * p->max;
        if (p->cpu) {
            char* n = (char*)calloc(p->size, 1);
            memcpy(n, p->cpu, p->stride * p->lastEle);
            free(p->cpu);
            p->cpu = n;
        }
    }
    p->lastEle++; p->usedNum++;
    return Elem((uint64_t)grp, (uint64_t)lev, p->lastEle - 1);
}
static inline ora_node* node_at(ora_tree* t, uint64_t id)
{
    return (ora_node*)(t->pool[0][ElemLev(id)].cpu + 64 * ElemNdx(id));
}
static inline uint64_t* clist_at(ora_tree* t, uint64_t id)
{
    ora_pool* p = &t->pool[1][ElemLev(id)];
    return (uint64_t*)(p->cpu + p->stride * ElemNdx(id));
}

/* getRes / getRange / getVoxCnt: src/gvdb_volume_gvdb.h:606-625 */
static inline int lev_res(const ora_tree* t, int lev) { return 1 << t->logdim[lev]; }
static inline int lev_range(const ora_tree* t, int lev)
{
    if (lev == -1) return 1;
    int r = lev_res(t, 0);
    for (int l = 1; l <= lev; l++) r *= lev_res(t, l);
    return r;
}
static inline uint64_t lev_voxcnt(const ora_tree* t, int lev) { uint64_t r = (uint64_t)lev_res(t, lev); return r * r * r; }

/* VolumeGVDB::Configure, src/gvdb_volume_gvdb.cpp:2380-2434 + AddChannel :2437-2453 (atlas bookkeeping only) */
ora_tree* ora_tree_create(int levs, const int* logdim, const int* initcnt, int cx, int cy, int cz, int apron)
{
    ora_tree* t = (ora_tree*)calloc(1, sizeof *t);
    t->levs = levs;
    for (int n = 0; n < levs; n++) t->logdim[n] = logdim[n] == 0 ? 1 : logdim[n];
    for (int n = 0; n < levs; n++) pool_create(&t->pool[0][n], 64, initcnt[n] == 0 ? 1 : (uint64_t)initcnt[n]);
    pool_create(&t->pool[1][0], 0, 0);
    for (int n = 1; n < levs; n++) pool_create(&t->pool[1][n], 8 * lev_voxcnt(t, n), initcnt[n] == 0 ? 1 : (uint64_t)initcnt[n]);
    t->root = ID_UNDEFL;
    t->atlas_cnt.x = cx; t->atlas_cnt.y = cy; t->atlas_cnt.z = cz;
    t->apron = apron; t->leafdim = lev_res(t, 0);
    t->atlas_max = (uint64_t)cx * cy * cz;
    t->epsilon = 0.001f; t->maxiter = 256;          /* src/gvdb_volume_gvdb.cpp:71-72 */
    return t;
}
void ora_tree_destroy(ora_tree* t)
{
    if (!t) return;
    for (int g = 0; g < 2; g++) for (int l = 0; l < ORA_MAXLEV; l++) free(t->pool[g][l].cpu);
    free(t->amap);
    free(t);
}
void ora_set_epsilon(ora_tree* t, float eps, int maxiter) { t->epsilon = eps; t->maxiter = maxiter; }

/* SetupNode, src/gvdb_volume_gvdb.cpp:2512-2525 */
static void setup_node(ora_tree* t, uint64_t id, int lev, i3 pos)
{
    ora_node* n = node_at(t, id);
    n->mLev = (uint8_t)lev; n->mPos = pos;
    n->mChildList = ID_UNDEFL; n->mParent = ID_UNDEFL;
    n->mValue.x = n->mValue.y = n->mValue.z = -1;
    n->mFlags = 1;
}
/* InsertChild (bitmasks off), src/gvdb_volume_gvdb.cpp:2980-3027 */
static uint64_t insert_child(ora_tree* t, uint64_t nodeid, uint64_t childid, uint32_t i)
{
    node_at(t, childid)->mParent = nodeid;
    ora_node* curr = node_at(t, nodeid);
    if (curr->mChildList == ID_UNDEFL) {
        uint64_t cl = pool_alloc(t, 1, curr->mLev);
        curr = node_at(t, nodeid);
        curr->mChildList = cl;
        memset(clist_at(t, cl), 0xFF, 8 * lev_voxcnt(t, curr->mLev));
    }
    clist_at(t, curr->mChildList)[i] = childid;
    return childid;
}
static uint64_t get_child_node(ora_tree* t, uint64_t nodeid, uint32_t b)     /* :3079-3092 */
{
    ora_node* curr = node_at(t, nodeid);
    if (curr->mChildList == ID_UNDEFL) return ID_UNDEF64;
    return clist_at(t, curr->mChildList)[b];
}
/* getPosInNode, :2858-2875 */
static int pos_in_node(ora_tree* t, uint64_t id, i3 pos, uint32_t* bit)
{
    ora_node* c = node_at(t, id);
    int res = lev_res(t, c->mLev), range = lev_range(t, c->mLev);
    i3 p = { pos.x - c->mPos.x, pos.y - c->mPos.y, pos.z - c->mPos.z };
    if (p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < range && p.y < range && p.z < range) {
        p.x = p.x * res / range; p.y = p.y * res / range; p.z = p.z * res / range;
        *bit = (uint32_t)((p.z * res + p.y) * res + p.x);
        return 1;
    }
    *bit = 0;
    return 0;
}
/* GetCoveringNode, :2915-2931 */
static i3 covering_node(ora_tree* t, int lev, i3 pos, int* range_out)
{
    int range = lev_range(t, lev);
    *range_out = range;
    i3 np = {0, 0, 0};
    if (lev == ORA_MAXLEV - 1) return np;
    np.x = pos.x / range * range; np.y = pos.y / range * range; np.z = pos.z / range * range;
    if (pos.x < np.x) np.x -= range;
    if (pos.y < np.y) np.y -= range;
    if (pos.z < np.z) np.z -= range;
    return np;
}
static uint64_t activate(ora_tree* t, uint64_t nodeid, i3 pos, int* bNew, uint64_t stopnode, int stoplev);

/* AddChildNode, :2707-2723 */
static uint64_t add_child_node(ora_tree* t, uint64_t nodeid, i3 ppos, int plev, uint32_t i)
{
    uint64_t child = pool_alloc(t, 0, plev - 1);
    int range = lev_range(t, plev - 1), logr = t->logdim[plev];
    uint32_t mask = (1u << logr) - 1;
    i3 p = { (int)(i & mask), (int)((i & (mask << logr)) >> logr), (int)((i & (mask << (2 * logr))) >> (2 * logr)) };
    p.x = p.x * range + ppos.x; p.y = p.y * range + ppos.y; p.z = p.z * range + ppos.z;
    setup_node(t, child, plev - 1, p);
    return insert_child(t, nodeid, child, i);
}
/* Reparent, :2726-2763 */
static uint64_t reparent(ora_tree* t, int lev, uint64_t prevroot, i3 pos, int* bNew)
{
    i3 prev_pos = node_at(t, prevroot)->mPos, pos1 = {0, 0, 0}, pos2;
    int cover = 0, range;
    while (!cover && lev < ORA_MAXLEV) {
        lev++;
        pos1 = covering_node(t, lev, pos, &range);
        pos2 = covering_node(t, lev, prev_pos, &range);
        cover = (pos1.x == pos2.x && pos1.y == pos2.y && pos1.z == pos2.z);
    }
    if (lev >= ORA_MAXLEV) return ID_UNDEFL;
    uint64_t newroot = pool_alloc(t, 0, lev);
    if (newroot == ID_UNDEFL) return ID_UNDEFL;
    setup_node(t, newroot, lev, pos1);
    int bn = 0;
    activate(t, newroot, prev_pos, &bn, prevroot, 0);
    uint64_t leaf = activate(t, newroot, pos, bNew, ID_UNDEFL, 0);
    t->root = newroot;
    return node_at(t, leaf)->mParent;
}
/* ActivateSpace (recursive), :2804-2855 */
static uint64_t activate(ora_tree* t, uint64_t nodeid, i3 pos, int* bNew, uint64_t stopnode, int stoplev)
{
    uint32_t b;
    if (t->root == ID_UNDEFL && nodeid == t->root) {
        int range;
        i3 p = covering_node(t, stoplev, pos, &range);
        t->root = pool_alloc(t, 0, stoplev);
        setup_node(t, t->root, stoplev, p);
        nodeid = t->root;
    }
    ora_node* curr = node_at(t, nodeid);
    if (pos_in_node(t, nodeid, pos, &b)) {
        if (stopnode != ID_UNDEFL) {
            ora_node* sn = node_at(t, stopnode);
            if (pos.x == sn->mPos.x && pos.y == sn->mPos.y && pos.z == sn->mPos.z &&
                get_child_node(t, nodeid, b) == ID_UNDEF64 && curr->mLev == sn->mLev + 1)
                return insert_child(t, nodeid, stopnode, b);
        }
        if (curr->mLev == stoplev) return nodeid;
        uint64_t childid;
        if (get_child_node(t, nodeid, b) == ID_UNDEF64) {
            int lev = curr->mLev;
            childid = add_child_node(t, nodeid, curr->mPos, lev, b);
            if (lev == 1) *bNew = 1;
        } else {
            childid = get_child_node(t, nodeid, b);
        }
        if (ElemLev(childid) == 0) return childid;
        return activate(t, childid, pos, bNew, stopnode, stoplev);
    } else {
        uint64_t parent = curr->mParent;
        if (parent == ID_UNDEFL) {
            parent = reparent(t, curr->mLev, nodeid, pos, bNew);
            if (parent == ID_UNDEFL) return ID_UNDEFL;
        }
        return activate(t, parent, pos, bNew, stopnode, stoplev);
    }
}
/* ActivateSpace(Vector3DF), :2766-2774 */
int64_t ora_activate_space(ora_tree* t, int x, int y, int z)
{
    int bnew = 0;
    i3 p = {x, y, z};
    uint64_t id = activate(t, t->root, p, &bnew, ID_UNDEFL, 0);
    if (id == ID_UNDEFL) return -1;
    return (int64_t)ElemNdx(id);
}
/* the activation loop of a volume build: ActivateSpace for n brick corners (xyz triples), in order */
void ora_activate_bricks(ora_tree* t, const int32_t* pos, int n)
{
    for (int i = 0; i < n; i++) ora_activate_space(t, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
}
/* ComputeBounds, :1792-1816 (called by FinishTopology :1579-1593) */
void ora_finish_topology(ora_tree* t)
{
    int range = lev_range(t, 0);
    ora_pool* p = &t->pool[0][0];
    if (p->lastEle == 0) return;
    ora_node* c = (ora_node*)p->cpu;
    t->vmin.x = (float)c->mPos.x; t->vmin.y = (float)c->mPos.y; t->vmin.z = (float)c->mPos.z;
    t->vmax = t->vmin;
    for (uint64_t n = 0; n < p->lastEle; n++) {
        c = (ora_node*)(p->cpu + 64 * n);
        if (!c->mFlags) continue;
        if (c->mPos.x < t->vmin.x) t->vmin.x = (float)c->mPos.x;
        if (c->mPos.y < t->vmin.y) t->vmin.y = (float)c->mPos.y;
        if (c->mPos.z < t->vmin.z) t->vmin.z = (float)c->mPos.z;
        if (c->mPos.x + range > t->vmax.x) t->vmax.x = (float)(c->mPos.x + range);
        if (c->mPos.y + range > t->vmax.y) t->vmax.y = (float)(c->mPos.y + range);
        if (c->mPos.z + range > t->vmax.z) t->vmax.z = (float)(c->mPos.z + range);
    }
}
/* Allocator::getAtlasPos, src/gvdb_allocator.cpp:705-715 */
static i3 atlas_pos(const ora_tree* t, uint64_t id)
{
    i3 p;
    int a2 = t->atlas_cnt.x * t->atlas_cnt.y;
    p.z = (int)(id / (uint64_t)a2); id -= (uint64_t)p.z * a2;
    p.y = (int)(id / (uint64_t)t->atlas_cnt.x); id -= (uint64_t)p.y * t->atlas_cnt.x;
    p.x = (int)id;
    int s = t->leafdim + 2 * t->apron;
    p.x = p.x * s + t->apron; p.y = p.y * s + t->apron; p.z = p.z * s + t->apron;
    return p;
}
/* UpdateAtlas, src/gvdb_volume_gvdb.cpp:2630-2704 (+ AtlasResize gvdb_allocator.cpp:622-650, AtlasAlloc :692-703,
 * ClearMapping :2559-2580, AssignMapping :2583-2588) */
void ora_update_atlas(ora_tree* t)
{
    ora_pool* p = &t->pool[0][0];
    uint64_t total = p->lastEle, used = p->usedNum;
    t->atlas_last = 0;
    if (used > t->atlas_max) {
        t->atlas_cnt.z = (int)ceil(used / (float)(t->atlas_cnt.x * t->atlas_cnt.y));
        t->atlas_max = (uint64_t)t->atlas_cnt.x * t->atlas_cnt.y * t->atlas_cnt.z;
    }
    for (uint64_t n = 0; n < total; n++) {
        ora_node* nd = (ora_node*)(p->cpu + 64 * n);
        if (!nd->mFlags) continue;
        if (t->atlas_last >= t->atlas_max) {
            uint64_t want = t->atlas_last + (uint64_t)t->atlas_cnt.x * t->atlas_cnt.y;
            t->atlas_cnt.z = (int)ceil(want / (float)(t->atlas_cnt.x * t->atlas_cnt.y));
            t->atlas_max = (uint64_t)t->atlas_cnt.x * t->atlas_cnt.y * t->atlas_cnt.z;
        }
        nd->mValue = atlas_pos(t, t->atlas_last++);
    }
    uint64_t cnt = (uint64_t)t->atlas_cnt.x * t->atlas_cnt.y * t->atlas_cnt.z;
    if (cnt != t->amap_cnt) { free(t->amap); t->amap = (ora_atlas_node*)malloc(cnt * sizeof(ora_atlas_node)); t->amap_cnt = cnt; }
    for (uint64_t i = 0; i < cnt; i++) { t->amap[i].mLeafNode = -1; t->amap[i].mPos.x = t->amap[i].mPos.y = t->amap[i].mPos.z = -1; }
    int leafres = t->leafdim + 2 * t->apron;
    for (uint64_t n = 0; n < total; n++) {
        ora_node* nd = (ora_node*)(p->cpu + 64 * n);
        if (!nd->mFlags) continue;
        int ix = nd->mValue.x / leafres, iy = nd->mValue.y / leafres, iz = nd->mValue.z / leafres;
        ora_atlas_node* an = &t->amap[((size_t)iz * t->atlas_cnt.y + iy) * t->atlas_cnt.x + ix];
        an->mPos = nd->mPos; an->mLeafNode = (int)n;
    }
}

uint64_t ora_pool_count(const ora_tree* t, int g, int l) { return t->pool[g][l].lastEle; }
uint64_t ora_pool_width(const ora_tree* t, int g, int l) { return t->pool[g][l].stride; }
const void* ora_pool_data(const ora_tree* t, int g, int l) { return t->pool[g][l].cpu; }
int ora_num_levels(const ora_tree* t) { return t->levs; }
void ora_atlas_res(const ora_tree* t, int r[3])
{
    int s = t->leafdim + 2 * t->apron;
    r[0] = t->atlas_cnt.x * s; r[1] = t->atlas_cnt.y * s; r[2] = t->atlas_cnt.z * s;
}
const void* ora_atlas_map(const ora_tree* t, uint64_t* bytes) { *bytes = t->amap_cnt * sizeof(ora_atlas_node); return t->amap; }

/* PrepareVDB, src/gvdb_volume_gvdb.cpp:3946-3989 */
void ora_fill_vdbinfo(const ora_tree* t, void* out)
{
    ora_vdbinfo v;
    memset(&v, 0, sizeof v);
    int tlev = 1;
    for (int n = t->levs - 1; n >= 0; n--) {
        v.dim[n] = t->logdim[n];
        v.res[n] = lev_res(t, n);
        float rg = (float)lev_range(t, n), rs = (float)lev_res(t, n);
        v.vdel[n].x = v.vdel[n].y = v.vdel[n].z = rg / rs;
        v.noderange[n].x = v.noderange[n].y = v.noderange[n].z = lev_range(t, n);
        v.nodecnt[n] = (int)t->pool[0][n].lastEle;
        v.nodewid[n] = (int)t->pool[0][n].stride;
        v.childwid[n] = (int)t->pool[1][n].stride;
        if (v.nodecnt[n] == 1) tlev = n;
    }
    v.atlas_apron = t->apron;
    v.atlas_cnt = t->atlas_cnt;
    int s = t->leafdim + 2 * t->apron;
    v.atlas_res.x = t->atlas_cnt.x * s; v.atlas_res.y = t->atlas_cnt.y * s; v.atlas_res.z = t->atlas_cnt.z * s;
    v.brick_res = s;
    for (int n = 0; n < t->apron; n++) { v.apron_table[n] = n; v.apron_table[(t->apron * 2 - 1) - n] = (s - 1) - n; }
    v.top_lev = tlev; v.epsilon = t->epsilon; v.max_iter = t->maxiter;
    v.bmin = t->vmin; v.bmax = t->vmax;
    v.clr_chan = 255;                                 /* CHAN_UNDEF (gvdb_volume_gvdb.cpp constructor) */
    for (int c = 0; c < 32; c++) { v.volIn[c] = ID_UNDEFL; v.volOut[c] = ID_UNDEFL; }
    memcpy(out, &v, sizeof v);
}

/* ============================================================================================ atlas + apron */
/* point query: getNode(lev,start,pos) kernels/cuda_gvdb_nodes.cuh:199-226, integer form (all inputs are voxel centres) */
static const ora_node* leaf_at_point(const ora_tree* t, int x, int y, int z)
{
    if (t->root == ID_UNDEFL) return NULL;
    ora_vdbinfo dummy; (void)dummy;
    int lev = ElemLev(t->root);
    const ora_node* n = (const ora_node*)(t->pool[0][lev].cpu + 64 * ElemNdx(t->root));
    while (lev > 0 && n) {
        int range = lev_range(t, lev), res = lev_res(t, lev), cdel = range / res;
        int px = x - n->mPos.x, py = y - n->mPos.y, pz = z - n->mPos.z;
        if (px < 0 || py < 0 || pz < 0 || px >= range || py >= range || pz >= range) return NULL;
        int b = ((pz / cdel) * res + (py / cdel)) * res + (px / cdel);
        if (n->mChildList == ID_UNDEFL) return NULL;
        const ora_pool* cp = &t->pool[1][ElemLev(n->mChildList)];
        uint64_t c = ((const uint64_t*)(cp->cpu + cp->stride * ElemNdx(n->mChildList)))[b];
        if (c == ID_UNDEF64) return NULL;
        lev--;
        n = (const ora_node*)(t->pool[0][lev].cpu + 64 * ElemNdx(c));
    }
    return n;
}
/* brick upload (LoadVBX-style slice writes, src/gvdb_volume_gvdb.cpp:666-673) followed by
 * UpdateApron<float> (kernels/cuda_gvdb_operators.cuh:72-126): every apron texel takes the value of the voxel at
 * the same index-space position in whichever active brick contains it, else `boundval`. */
void ora_fill_atlas(const ora_tree* t, const float* values, float* atlas, float boundval)
{
    int res[3]; ora_atlas_res(t, res);
    const ora_pool* p = &t->pool[0][0];
    const int R = t->leafdim, A = t->apron, S = R + 2 * A;
    size_t rx = (size_t)res[0], ry = (size_t)res[1];
    memset(atlas, 0, sizeof(float) * rx * ry * (size_t)res[2]);
    for (uint64_t n = 0; n < p->lastEle; n++) {
        const ora_node* nd = (const ora_node*)(p->cpu + 64 * n);
        const float* v = values + (size_t)R * R * R * n;
        for (int k = 0; k < R; k++) for (int j = 0; j < R; j++)
            memcpy(&atlas[((size_t)(nd->mValue.z + k) * ry + (nd->mValue.y + j)) * rx + nd->mValue.x], v + (k * R + j) * R, sizeof(float) * R);
    }
    #pragma omp parallel for schedule(dynamic, 256)
    for (int64_t n = 0; n < (int64_t)p->lastEle; n++) {
        const ora_node* nd = (const ora_node*)(p->cpu + 64 * n);
        for (int k = 0; k < S; k++) for (int j = 0; j < S; j++) for (int i = 0; i < S; i++) {
            if (i > 0 && i < S - 1 && j > 0 && j < S - 1 && k > 0 && k < S - 1) continue;   /* interior */
            int wx = nd->mPos.x + i - A, wy = nd->mPos.y + j - A, wz = nd->mPos.z + k - A;
            const ora_node* src = leaf_at_point(t, wx, wy, wz);
            float val = boundval;
            if (src) val = atlas[((size_t)(src->mValue.z + (wz - src->mPos.z)) * ry + (src->mValue.y + (wy - src->mPos.y))) * rx
                                 + (src->mValue.x + (wx - src->mPos.x))];
            atlas[((size_t)(nd->mValue.z - A + k) * ry + (nd->mValue.y - A + j)) * rx + (nd->mValue.x - A + i)] = val;
        }
    }
}

/* ============================================================================================ CPU ray caster */
static inline f3 F3(float x, float y, float z) { f3 r = {x, y, z}; return r; }
static inline f3 add3(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub3(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul3(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 div3(f3 a, f3 b) { return F3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline f3 scl3(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 nrm3(f3 v) { float inv = 1.0f / sqrtf(dot3(v, v)); return scl3(v, inv); }       /* cuda_math.cuh:1367 */
static inline f3 flr3(f3 a) { return F3(floorf(a.x), floorf(a.y), floorf(a.z)); }
static inline f3 mmult(const float* m, f3 v)                                                      /* cuda_math.cuh:1472 */
{
    return F3(v.x * m[0] + v.y * m[4] + v.z * m[8] + m[12], v.x * m[1] + v.y * m[5] + v.z * m[9] + m[13],
              v.x * m[2] + v.y * m[6] + v.z * m[10] + m[14]);
}

typedef struct {
    const ora_volume*  v;
    const ora_vdbinfo* g;
    const ora_scninfo* s;
} rc_ctx;

/* texture unit model: see gvdbx_device.cuh GxSampler<LINEAR> / profiles/r01_trilinear_calibration.md.
 * (texel centres at +0.5, 8-bit corner weights from a z -> x -> y hierarchical split with round-half-up) */
static inline void tex_split(float c, int* i, int* a)
{
    float cb = c - 0.5f, f = floorf(cb);
    *a = (int)floorf((cb - f) * 256.0f + 0.5f);
    *i = (int)f;
    if (*a >= 256) { *a = 0; (*i)++; }
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
float ora_tex3d(const ora_volume* v, float x, float y, float z)
{
    int ix, iy, iz, ax, ay, az;
    tex_split(x, &ix, &ax); tex_split(y, &iy, &ay); tex_split(z, &iz, &az);
    const int rx = v->atlas_res[0], ry = v->atlas_res[1], rz = v->atlas_res[2];
    int x0 = clampi(ix, 0, rx - 1), x1 = clampi(ix + 1, 0, rx - 1);
    int y0 = clampi(iy, 0, ry - 1), y1 = clampi(iy + 1, 0, ry - 1);
    int z0 = clampi(iz, 0, rz - 1), z1 = clampi(iz + 1, 0, rz - 1);
    const float* A = v->atlas;
#define TX(X, Y, Z) A[((size_t)(Z) * ry + (Y)) * rx + (X)]
    int by = 256 - ay, s0 = 256 - az, s1 = az;
    int x1a = (s0 * ax + 128) >> 8, x0a = s0 - x1a, x1b = (s1 * ax + 128) >> 8, x0b = s1 - x1b;
    int w110 = (x1a * ay + 128) >> 8, w100 = x1a - w110, w000 = (x0a * by + 128) >> 8, w010 = x0a - w000;
    int w111 = (x1b * ay + 128) >> 8, w101 = x1b - w111, w001 = (x0b * by + 128) >> 8, w011 = x0b - w001;
    double acc = (double)w000 * TX(x0, y0, z0) + (double)w100 * TX(x1, y0, z0) + (double)w010 * TX(x0, y1, z0) + (double)w110 * TX(x1, y1, z0)
               + (double)w001 * TX(x0, y0, z1) + (double)w101 * TX(x1, y0, z1) + (double)w011 * TX(x0, y1, z1) + (double)w111 * TX(x1, y1, z1);
#undef TX
    return (float)(acc * (1.0 / 256.0));
}

/* rayBoxIntersect, kernels/cuda_gvdb_geom.cuh:85-98 */
static inline f3 ray_box(f3 rpos, f3 rdir, f3 vmin, f3 vmax)
{
    float h0 = (vmin.x - rpos.x) / rdir.x, h1 = (vmax.x - rpos.x) / rdir.x;
    float h2 = (vmin.y - rpos.y) / rdir.y, h3 = (vmax.y - rpos.y) / rdir.y;
    float h4 = (vmin.z - rpos.z) / rdir.z, h5 = (vmax.z - rpos.z) / rdir.z;
    float tn = fmaxf(fmaxf(fminf(h0, h1), fminf(h2, h3)), fminf(h4, h5));
    float tf = fminf(fminf(fmaxf(h0, h1), fmaxf(h2, h3)), fmaxf(h4, h5));
    if (tn < 0) tn = 0.0f;
    return F3(tn, tf, (tf < tn || tf < 0) ? NOHIT : 0);
}

/* HDDAState, kernels/cuda_gvdb_dda.cuh:38-91 */
typedef struct { f3 pos, dir; i3 pStep; f3 tDel, t; i3 p; f3 tSide; i3 mask; } dda_t;
static inline void dda_set(dda_t* d, f3 pos, f3 dir, f3 t)
{
    d->pos = pos; d->dir = dir; d->t = t;
    d->pStep.x = dir.x > 0 ? 1 : -1; d->pStep.y = dir.y > 0 ? 1 : -1; d->pStep.z = dir.z > 0 ? 1 : -1;
}
static inline void dda_prepare(dda_t* d, f3 vmin, f3 vdel, int leaf)
{
    f3 q = div3(leaf ? F3(1, 1, 1) : vdel, d->dir);
    d->tDel = F3(fabsf(q.x), fabsf(q.y), fabsf(q.z));
    f3 pf = sub3(add3(d->pos, scl3(d->dir, d->t.x)), vmin);
    if (!leaf) pf = div3(pf, vdel);
    f3 fl = flr3(pf);
    f3 ps = F3((float)d->pStep.x, (float)d->pStep.y, (float)d->pStep.z);
    f3 a = add3(mul3(add3(sub3(fl, pf), F3(0.5f, 0.5f, 0.5f)), ps), F3(0.5f, 0.5f, 0.5f));
    d->tSide = mul3(a, d->tDel);
    if (!leaf) d->tSide = add3(d->tSide, F3(d->t.x, d->t.x, d->t.x));
    d->p.x = (int)fl.x; d->p.y = (int)fl.y; d->p.z = (int)fl.z;
}
static inline void dda_next(dda_t* d)
{
    d->mask.x = (d->tSide.x < d->tSide.y) & (d->tSide.x <= d->tSide.z);
    d->mask.y = (d->tSide.y < d->tSide.z) & (d->tSide.y <= d->tSide.x);
    d->mask.z = (d->tSide.z < d->tSide.x) & (d->tSide.z <= d->tSide.y);
    d->t.y = d->mask.x ? d->tSide.x : (d->mask.y ? d->tSide.y : d->tSide.z);
}
static inline void dda_step(dda_t* d)
{
    d->t.x = d->t.y;
    d->tSide.x += (float)d->mask.x * d->tDel.x; d->tSide.y += (float)d->mask.y * d->tDel.y; d->tSide.z += (float)d->mask.z * d->tDel.z;
    d->p.x += d->mask.x * d->pStep.x; d->p.y += d->mask.y * d->pStep.y; d->p.z += d->mask.z * d->pStep.z;
}

static inline const ora_node* get_node(const rc_ctx* c, int lev, int n)            /* nodes.cuh:184-196 */
{
    return (const ora_node*)((const char*)c->v->pool0[lev] + (size_t)n * c->g->nodewid[lev]);
}
static inline int get_child(const rc_ctx* c, const ora_node* node, int b)           /* nodes.cuh:115-124 */
{
    uint64_t listid = node->mChildList;
    if (listid == ID_UNDEFL) return -1;
    int clev = (int)((listid >> 8) & 0xFF);
    uint64_t cndx = listid >> 16;
    const uint64_t* cl = (const uint64_t*)((const char*)c->v->pool1[clev] + cndx * (size_t)c->g->childwid[clev]);
    return (int)(cl[b] >> 16);
}
static inline float fetch(const rc_ctx* c, float x, float y, float z) { return ora_tex3d(c->v, x, y, z); }

/* transfer(), kernels/cuda_gvdb_dda.cuh:20-23 */
static inline f4 transfer(const rc_ctx* c, float v)
{
    double u = (double)((v - c->s->thresh.x) / (c->s->thresh.z - c->s->thresh.y));
    u = u < 0.0 ? 0.0 : u; u = u > 1.0 ? 1.0 : u;
    int idx = (int)(u * (double)16300.0f);
    const float* T = c->v->transfer + 4 * (size_t)idx;
    f4 r = {T[0], T[1], T[2], T[3]};
    return r;
}
/* getGradient / getGradientLevelSet, kernels/cuda_gvdb_raycast.cuh:132-157 */
static inline f3 gradient(const rc_ctx* c, f3 p, int levelset)
{
    float xm = fetch(c, p.x - .5f, p.y, p.z), xp = fetch(c, p.x + .5f, p.y, p.z);
    float ym = fetch(c, p.x, p.y - .5f, p.z), yp = fetch(c, p.x, p.y + .5f, p.z);
    float zm = fetch(c, p.x, p.y, p.z - .5f), zp = fetch(c, p.x, p.y, p.z + .5f);
    f3 g = levelset ? F3(xp - xm, yp - ym, zp - zm) : F3(xm - xp, ym - yp, zm - zp);
    return nrm3(g);
}

typedef struct { f3 hit, norm; f4 clr; } rc_out;

/* raySurfaceVoxelBrick, kernels/cuda_gvdb_raycast.cuh:227-265 */
static void brick_voxel(const rc_ctx* c, int nodeid, f3 t, f3 pos, f3 dir, rc_out* o)
{
    const ora_node* node = get_node(c, 0, nodeid);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    f3 a = F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z);
    int res0 = c->g->res[0];
    dda_t d;
    dda_set(&d, pos, dir, t);
    dda_prepare(&d, vmin, F3(1, 1, 1), 1);
    for (int it = 0; it < 256 && d.p.x >= 0 && d.p.y >= 0 && d.p.z >= 0 && d.p.x < res0 && d.p.y < res0 && d.p.z < res0; it++) {
        if (fetch(c, d.p.x + a.x + .5f, d.p.y + a.y + .5f, d.p.z + a.z + .5f) > c->s->thresh.x) {
            vmin = add3(vmin, F3((float)d.p.x, (float)d.p.y, (float)d.p.z));
            d.t = ray_box(pos, dir, vmin, add3(vmin, F3(1, 1, 1)));
            if (d.t.z == NOHIT) { o->hit.z = NOHIT; continue; }
            o->hit = add3(pos, scl3(dir, d.t.x));
            f3 fc = sub3(sub3(o->hit, vmin), F3(0.5f, 0.5f, 0.5f));
            fc = sub3(fc, scl3(dir, 0.01f));
            float mx = fmaxf(fmaxf(fabsf(fc.x), fabsf(fc.y)), fabsf(fc.z));
            o->norm.x = fabsf(fc.x) == mx ? copysignf(1.0f, fc.x) : 0.0f;
            o->norm.y = fabsf(fc.y) == mx ? copysignf(1.0f, fc.y) : 0.0f;
            o->norm.z = fabsf(fc.z) == mx ? copysignf(1.0f, fc.z) : 0.0f;
            return;
        }
        dda_next(&d);
        dda_step(&d);
    }
}
/* raySurfaceTrilinearBrick, :281-300 */
static void brick_trilinear(const rc_ctx* c, int nodeid, f3 t, f3 pos, f3 dir, rc_out* o)
{
    const ora_node* node = get_node(c, 0, nodeid);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    f3 a = F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z);
    float res0 = (float)c->g->res[0], step = c->s->steps.x;
    t.x = step * ceilf(t.x / step);
    f3 p = sub3(add3(pos, scl3(dir, t.x)), vmin);
    for (int it = 0; it < 256 && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0; it++) {
        if (fetch(c, p.x + a.x, p.y + a.y, p.z + a.z) >= c->s->thresh.x) {
            o->hit = add3(p, vmin);
            o->norm = gradient(c, add3(p, a), 0);
            return;
        }
        p = add3(p, scl3(dir, step));
        t.x += step;
    }
}
/* rayLevelSetBrick + rayLevelSet, :389-410, :186-197 */
static void brick_levelset(const rc_ctx* c, int nodeid, f3 t, f3 pos, f3 dir, rc_out* o)
{
    const ora_node* node = get_node(c, 0, nodeid);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    f3 a = F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z);
    float res0 = (float)c->g->res[0], step = c->s->steps.x;
    f3 p = sub3(add3(pos, scl3(dir, t.x)), vmin);
    for (int it = 0; it < 256 && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x <= res0 && p.y <= res0 && p.z <= res0; it++) {
        if (fetch(c, p.x + a.x, p.y + a.y, p.z + a.z) < c->s->thresh.x) {
            o->hit = add3(p, vmin);          /* the fine march re-tests this point and returns at i = 0 */
            if (o->hit.z != NOHIT) { o->norm = gradient(c, add3(p, a), 1); return; }
        }
        p = add3(p, scl3(dir, step));
    }
}
/* rayDeepBrick, :485-533 (no depth buffer on the CPU checker) */
static void brick_deep(const rc_ctx* c, int nodeid, f3 t, f3 pos, f3 dir, rc_out* o)
{
    const ora_node* node = get_node(c, 0, nodeid);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    f3 a = F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z);
    float res0 = (float)c->g->res[0], step = c->s->steps.x;
    t.x = step * ceilf(t.x / step);
    f3 wp = add3(pos, scl3(dir, t.x));
    f3 p = sub3(wp, vmin);
    f3 wpt = scl3(dir, step);
    float dt = sqrtf(dot3(wpt, wpt));
    if (o->hit.x == 0) o->hit.x = t.x;
    f4* k = &o->clr;
    for (int it = 0; k->w > c->s->cutoff.y && it < 256 && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0; it++) {
        float raw = fetch(c, p.x + a.x, p.y + a.y, p.z + a.z);
        if (raw >= c->s->cutoff.x) {
            f4 val = transfer(c, raw);
            val.w = expf(c->s->extinct.x * val.w * step);
            k->x += val.x * k->w * (1 - val.w) * c->s->extinct.y;
            k->y += val.y * k->w * (1 - val.w) * c->s->extinct.y;
            k->z += val.z * k->w * (1 - val.w) * c->s->extinct.y;
            k->w *= val.w;
        }
        p = add3(p, wpt); wp = add3(wp, wpt);
        t.x += dt;
    }
    o->hit.y = t.x;
    k->x = fminf(k->x, 1.f); k->y = fminf(k->y, 1.f); k->z = fminf(k->z, 1.f); k->w = fmaxf(k->w, 0.f);
}

/* getTricubic, kernels/cuda_gvdb_raycast.cuh:32-96: 27 fetches at texel corners, quadratic B-spline weights */
static float tricubic(const rc_ctx* c, f3 p, f3 offs)
{
    f3 q = sub3(flr3(add3(p, offs)), F3(1, 1, 1));
    f3 fr = sub3(p, flr3(p));
    f3 tb = F3(fr.x * 0.5f + 0.25f, fr.y * 0.5f + 0.25f, fr.z * 0.5f + 0.25f);
    f3 ta = F3(1.0f - tb.x, 1.0f - tb.y, 1.0f - tb.z);
    f3 ta2 = mul3(ta, ta), tb2 = mul3(tb, tb), tab = scl3(mul3(ta, tb), 2.0f);
    float plane[3][3];
    for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) {
        float t0 = fetch(c, q.x, q.y + (float)j, q.z + (float)k), t1 = fetch(c, q.x + 1, q.y + (float)j, q.z + (float)k),
              t2 = fetch(c, q.x + 2, q.y + (float)j, q.z + (float)k);
        plane[k][j] = t0 * ta2.x + t1 * tab.x + t2 * tb2.x;
    }
    float col[3];
    for (int k = 0; k < 3; k++) col[k] = plane[k][0] * ta2.y + plane[k][1] * tab.y + plane[k][2] * tb2.y;
    return col[0] * ta2.z + col[1] * tab.z + col[2] * tb2.z;
}
/* getGradientTricubic, :159-169 */
static f3 gradient_tricubic(const rc_ctx* c, f3 p, f3 offs)
{
    const float vs = 0.5f;
    f3 g;
    g.x = (tricubic(c, add3(p, F3(-vs, 0, 0)), offs) - tricubic(c, add3(p, F3(vs, 0, 0)), offs)) / (2 * vs);
    g.y = (tricubic(c, add3(p, F3(0, -vs, 0)), offs) - tricubic(c, add3(p, F3(0, vs, 0)), offs)) / (2 * vs);
    g.z = (tricubic(c, add3(p, F3(0, 0, -vs)), offs) - tricubic(c, add3(p, F3(0, 0, vs)), offs)) / (2 * vs);
    return nrm3(g);
}
/* raySurfaceTricubicBrick, :316-339 */
static void brick_tricubic(const rc_ctx* c, int nodeid, f3 t, f3 pos, f3 dir, rc_out* o)
{
    const ora_node* node = get_node(c, 0, nodeid);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    f3 a = F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z);
    float res0 = (float)c->g->res[0];
    f3 p = sub3(add3(pos, scl3(dir, t.x)), vmin);
    for (int it = 0; it < 256 && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0; it++) {
        float vz = tricubic(c, p, a);
        if (vz >= c->s->thresh.x) {
            float vx = tricubic(c, sub3(p, scl3(dir, c->s->steps.z)), a);
            float vy = (vz - c->s->thresh.x) / (vz - vx);
            p = add3(p, scl3(dir, -vy * c->s->steps.z));
            o->hit = add3(p, vmin);
            o->norm = gradient_tricubic(c, p, a);
            return;
        }
        p = add3(p, scl3(dir, c->s->steps.x));
        t.x += c->s->steps.x;
    }
}
/* rayShadowBrick, :445-463: opacity accumulates in clr.w; the attenuation is evaluated in double like the literals make it */
static void brick_shadow(const rc_ctx* c, int nodeid, f3 t, f3 pos, f3 dir, rc_out* o)
{
    const ora_node* node = get_node(c, 0, nodeid);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    t.x += c->g->epsilon;
    t.y -= c->g->epsilon;
    f3 a = F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z);
    f3 p = sub3(add3(pos, scl3(dir, t.x)), vmin);
    f3 pt = scl3(dir, c->s->steps.x);
    float res0 = (float)c->g->res[0];
    for (; o->clr.w < 1 && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0;) {
        f4 T = transfer(c, fetch(c, p.x + a.x, p.y + a.y, p.z + a.z));
        float val = (float)exp((double)(c->s->extinct.x * T.w * c->s->steps.y) / (1.0 + (double)t.x * 0.4));
        o->clr.w = (float)(1.0 - (1.0 - (double)o->clr.w) * (double)val);
        p = add3(p, pt);
        t.x += c->s->steps.y;
    }
}
/* getNode(lev, start, pos) / getNodeAtPoint, kernels/cuda_gvdb_nodes.cuh:199-253: leaf index at an index-space point or -1 */
static int node_at_point(const rc_ctx* c, f3 pos)
{
    const ora_vdbinfo* g = c->g;
    int lev = g->top_lev, id = 0;
    const ora_node* node = get_node(c, lev, 0);
    while (lev > 0) {
        f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
        f3 vmax = add3(vmin, F3((float)g->noderange[lev].x, (float)g->noderange[lev].y, (float)g->noderange[lev].z));
        if (pos.x < vmin.x || pos.y < vmin.y || pos.z < vmin.z || pos.x >= vmax.x || pos.y >= vmax.y || pos.z >= vmax.z) return -1;
        f3 q = div3(sub3(pos, vmin), g->vdel[lev]);
        int px = (int)q.x, py = (int)q.y, pz = (int)q.z;
        int b = (((pz << g->dim[lev]) + py) << g->dim[lev]) + px;
        lev--;
        id = get_child(c, node, b);
        if (id == -1) return -1;
        node = get_node(c, lev, id);
    }
    return id;
}

#define ORA_BRICK_SHADOW 100    /* brick function selector for rayShadowBrick (not a shade mode) */

/* rayCast, kernels/cuda_gvdb_raycast.cuh:543-611 */
static void ray_cast(const rc_ctx* c, int shade, f3 pos, f3 dir, rc_out* o)
{
    const ora_vdbinfo* g = c->g;
    int nodeid[ORA_MAXLEV]; float tMax[ORA_MAXLEV];
    int lev = g->top_lev;
    nodeid[lev] = 0;
    f3 tStart = ray_box(pos, dir, g->bmin, g->bmax);
    const ora_node* node = get_node(c, lev, 0);
    f3 vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
    if (tStart.z == NOHIT) return;
    tStart.x += g->epsilon;
    tMax[lev] = tStart.y - g->epsilon;
    dda_t d;
    dda_set(&d, pos, dir, tStart);
    dda_prepare(&d, vmin, g->vdel[lev], 0);
    for (int it = 0; it < 256 && lev > 0 && lev <= g->top_lev && d.p.x >= 0 && d.p.y >= 0 && d.p.z >= 0 &&
                     d.p.x <= g->res[lev] && d.p.y <= g->res[lev] && d.p.z <= g->res[lev]; it++) {
        dda_next(&d);
        int b = (((d.p.z << g->dim[lev]) + d.p.y) << g->dim[lev]) + d.p.x;
        int child = -1;
        if (d.p.x < g->res[lev] && d.p.y < g->res[lev] && d.p.z < g->res[lev]) child = get_child(c, node, b);
        if (child != -1) {
            if (lev == 1) {
                nodeid[0] = child;
                d.t.x += g->epsilon;
                switch (shade) {
                case SCN_SHADE_VOXEL:     brick_voxel(c, child, d.t, pos, dir, o); break;
                case SCN_SHADE_TRILINEAR: brick_trilinear(c, child, d.t, pos, dir, o); break;
                case SCN_SHADE_LEVELSET:  brick_levelset(c, child, d.t, pos, dir, o); break;
                case ORA_SHADE_TRICUBIC:  brick_tricubic(c, child, d.t, pos, dir, o); break;
                case ORA_SHADE_EMPTYSKIP: o->hit = add3(pos, scl3(dir, d.t.x)); break;          /* rayEmptySkipBrick, :425-428 */
                case ORA_BRICK_SHADOW:    brick_shadow(c, child, d.t, pos, dir, o); break;
                default:                  brick_deep(c, child, d.t, pos, dir, o); break;
                }
                if (o->clr.w <= 0) { o->clr.w = 0; return; }
                if (o->hit.z != NOHIT) return;
                dda_step(&d);
            } else {
                lev--;
                nodeid[lev] = child;
                node = get_node(c, lev, child);
                vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
                d.t.x += g->epsilon;
                tMax[lev] = d.t.y - g->epsilon;
                dda_prepare(&d, vmin, g->vdel[lev], 0);
            }
        } else {
            dda_step(&d);
        }
        while (d.t.x > tMax[lev] && lev <= g->top_lev) {
            lev++;
            if (lev <= g->top_lev) {
                node = get_node(c, lev, nodeid[lev]);
                vmin = F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z);
                dda_prepare(&d, vmin, g->vdel[lev], 0);
            }
        }
    }
}

/* performPhongShading, kernels/cuda_gvdb_module.cu:38-57 */
static f4 phong(const rc_ctx* c, int shade, f3 shit, f3 snorm, f4 sclr)
{
    if (shit.z == NOHIT) return c->s->backclr;
    f3 ld = nrm3(sub3(c->s->light_pos, shit));
    float diff = (float)(0.9 * (double)fmaxf(0.0f, dot3(snorm, ld)));
    float amb = 0.1f;
    if (c->s->shadow_params.x > 0) {
        rc_out o2;
        o2.hit = F3(0, 0, NOHIT); o2.clr.x = o2.clr.y = o2.clr.z = 0; o2.clr.w = 1; o2.norm = F3(0, 0, 0);
        ray_cast(c, shade, add3(shit, scl3(snorm, c->s->shadow_params.y)), ld, &o2);
        if (o2.hit.z != NOHIT) diff = (float)((double)diff * (1.0 - (double)c->s->shadow_params.x));
    }
    f4 r = { sclr.x * (diff + amb), sclr.y * (diff + amb), sclr.z * (diff + amb), 1.0f };
    return r;
}

/* one camera ray through pixel (x, y) at sub-pixel offset (ox, oy), shaded to the float colour the kernels pack:
 * gvdbRayDeep / gvdbRaySurfaceVoxel / ...Trilinear / ...Tricubic / gvdbRayLevelSet / gvdbRayEmptySkip / gvdbSection2D /
 * gvdbSection3D, kernels/cuda_gvdb_module.cu:60-298; deep_shadow = the composition of SURVEY.md 8c (ii) */
static f4 shade_pixel(const rc_ctx* c, int shade, int deep_shadow, int x, int y, float ox, float oy, rc_out* o)
{
    const ora_scninfo* s = c->s;
    const int W = s->width, H = s->height;
    f4 clr;
    if (shade == ORA_SHADE_SECTION2D) {                                                /* module.cu:272-298 */
        f3 spnt = F3((float)((double)(float)x * 2.0 / W - 1.0), 0, (float)((double)(float)y * 2.0 / H - 1.0));
        f3 wpos = add3(s->slice_pnt, mul3(spnt, s->slice_norm));
        f4 r = {0, 0, 0, 1};
        int n = node_at_point(c, wpos);
        if (n < 0) return r;
        const ora_node* node = get_node(c, 0, n);
        f3 p = add3(F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z),
                    sub3(wpos, F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z)));
        f4 t = transfer(c, fetch(c, p.x, p.y, p.z));
        r.x = t.w * t.x; r.y = t.w * t.y; r.z = t.w * t.z;
        return r;
    }
    f3 rpos = mmult(s->invxform, s->campos);                                       /* geom.cuh:48-51 */
    float u = (float)(x + ox) / (float)W, w = (float)(y + oy) / (float)H;
    f3 vv = add3(add3(scl3(s->camu, u), scl3(s->camv, w)), s->cams);               /* geom.cuh:55-63 */
    f3 rdir = nrm3(mmult(s->invxrot, vv));
    o->norm = F3(0, 0, 0);
    if (shade == SCN_SHADE_VOLUME) {
        o->clr.x = o->clr.y = o->clr.z = 0; o->clr.w = 1;
        o->hit = F3(0, 0, NOHIT);
        ray_cast(c, shade, rpos, rdir, o);
        if (deep_shadow && o->hit.x != 0.f) {
            f3 spos = add3(rpos, scl3(rdir, o->hit.x));
            f3 ldir = nrm3(sub3(s->light_pos, spos));
            rc_out o2;
            o2.hit = F3(0, 0, NOHIT); o2.norm = F3(0, 0, 0);
            o2.clr.x = o2.clr.y = o2.clr.z = o2.clr.w = 0;
            ray_cast(c, ORA_BRICK_SHADOW, spos, ldir, &o2);
            float lit = 1.0f - o2.clr.w;
            o->clr.x *= lit; o->clr.y *= lit; o->clr.z *= lit;
        }
        float a = (float)(1.0 - (double)o->clr.w);
        clr.x = s->backclr.x + a * (o->clr.x - s->backclr.x);
        clr.y = s->backclr.y + a * (o->clr.y - s->backclr.y);
        clr.z = s->backclr.z + a * (o->clr.z - s->backclr.z);
        clr.w = a;
    } else if (shade == ORA_SHADE_EMPTYSKIP) {                                          /* module.cu:184-207 */
        o->clr.x = o->clr.y = o->clr.z = o->clr.w = 1;
        o->hit = F3(NOHIT, NOHIT, NOHIT);
        ray_cast(c, shade, rpos, rdir, o);
        if (o->hit.z != NOHIT) { clr.x = o->hit.x * 0.01f; clr.y = o->hit.y * 0.01f; clr.z = o->hit.z * 0.01f; }
        else clr = s->backclr;
        clr.w = 1.0f;
    } else if (shade == ORA_SHADE_SECTION3D) {                                          /* module.cu:225-269 */
        f4 k = {1, 1, 1, 0};
        f3 wpos = rpos;
        const f3 pn = s->slice_norm, pp = s->slice_pnt;                               /* rayPlaneIntersect, geom.cuh:67-71 */
        float t = ((pp.x - wpos.x) * pn.x + (pp.y - wpos.y) * pn.y + (pp.z - wpos.z) * pn.z) / (rdir.x * pn.x + rdir.y * pn.y + rdir.z * pn.z);
        t = t > 0 ? t : NOHIT;
        if (t > 0) {
            wpos = add3(wpos, scl3(rdir, t));
            int n = node_at_point(c, wpos);
            if (n >= 0) {
                const ora_node* node = get_node(c, 0, n);
                f3 p = add3(F3((float)node->mValue.x, (float)node->mValue.y, (float)node->mValue.z),
                            sub3(wpos, F3((float)node->mPos.x, (float)node->mPos.y, (float)node->mPos.z)));
                t = fetch(c, p.x, p.y, p.z);
                k = transfer(c, t);
            } else t = 0;
        }
        o->hit = F3(NOHIT, NOHIT, NOHIT);
        o->clr.x = o->clr.y = o->clr.z = o->clr.w = 1;
        ray_cast(c, SCN_SHADE_TRILINEAR, wpos, rdir, o);
        f4 a4;
        if (o->hit.z != NOHIT) {
            f3 ld = nrm3(sub3(s->light_pos, o->hit));
            float ds = (t > s->thresh.x) ? 1.0f : (float)(0.8 * (double)fmaxf(0.0f, dot3(o->norm, ld)));
            a4.x = o->clr.x * ds; a4.y = o->clr.y * ds; a4.z = o->clr.z * ds; a4.w = o->clr.w * ds;
        } else a4 = s->backclr;
        clr.x = a4.x + k.w * (k.x - a4.x); clr.y = a4.y + k.w * (k.y - a4.y); clr.z = a4.z + k.w * (k.z - a4.z);
        clr.w = 1.0f;
    } else {
        o->clr.x = o->clr.y = o->clr.z = o->clr.w = 1;
        o->hit = shade == SCN_SHADE_LEVELSET ? F3(0, 0, NOHIT) : F3(NOHIT, NOHIT, NOHIT);
        ray_cast(c, shade, rpos, rdir, o);
        /* the tricubic kernel shades its shadow ray with the trilinear brick function (module.cu:136) */
        clr = phong(c, shade == ORA_SHADE_TRICUBIC ? SCN_SHADE_TRILINEAR : shade, o->hit, o->norm, o->clr);
    }
    return clr;
}

/* Render(shade) for every mode of the reference's switch (gvdb_volume_gvdb.cpp:4363-4372); deep_shadow and spp are the
 * compositions of BASELINE.json configs 4 and 5 (spp rays per pixel on a g x g sub-pixel grid, colours summed in sample
 * order, scaled by 1/spp, packed once) */
int ora_render_ex(const ora_volume* v, const void* scninfo, int shade, int y0, int y1, uint8_t* out, float* hit_norm, int threads,
                  int deep_shadow, int spp)
{
    const ora_scninfo* s = (const ora_scninfo*)scninfo;
    const ora_vdbinfo* g = (const ora_vdbinfo*)v->vdbinfo;
    if (shade < 0 || shade > SCN_SHADE_VOLUME) return -1;
    if (g->top_lev < 1 || g->top_lev >= 5) return -1;
    if (spp < 1) spp = 1;
    rc_ctx c = { v, g, s };
    const int W = s->width, H = s->height;
    if (y0 < 0) y0 = 0;
    if (y1 > H) y1 = H;
    int grid = 1;
    while (grid * grid < spp) grid++;
    const float inv_grid = 1.0f / (float)grid, inv_spp = 1.0f / (float)spp;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (int y = y0; y < y1; y++) {
        for (int x = 0; x < W; x++) {
            rc_out o;
            o.hit = F3(0, 0, NOHIT); o.norm = F3(0, 0, 0);
            f4 clr = {0, 0, 0, 0};
            if (spp == 1) clr = shade_pixel(&c, shade, deep_shadow, x, y, 0.5f, 0.5f, &o);
            else {
                for (int k = 0; k < spp; k++) {
                    f4 q = shade_pixel(&c, shade, deep_shadow, x, y, ((float)(k % grid) + 0.5f) * inv_grid, ((float)(k / grid) + 0.5f) * inv_grid, &o);
                    clr.x += q.x; clr.y += q.y; clr.z += q.z; clr.w += q.w;
                }
                clr.x *= inv_spp; clr.y *= inv_spp; clr.z *= inv_spp; clr.w *= inv_spp;
            }
            uint8_t* px = out + 4 * ((size_t)y * W + x);
            px[0] = (uint8_t)(unsigned)(clr.x * 255); px[1] = (uint8_t)(unsigned)(clr.y * 255);
            px[2] = (uint8_t)(unsigned)(clr.z * 255);
            px[3] = (shade == ORA_SHADE_EMPTYSKIP || shade == ORA_SHADE_SECTION2D || shade == ORA_SHADE_SECTION3D) ? 255 : (uint8_t)(unsigned)(clr.w * 255);
            if (hit_norm) {
                float* hn = hit_norm + 8 * ((size_t)y * W + x);
                int miss = (o.hit.z == NOHIT);
                hn[0] = o.hit.x; hn[1] = o.hit.y; hn[2] = o.hit.z; hn[3] = 0;
                hn[4] = miss ? 0 : o.norm.x; hn[5] = miss ? 0 : o.norm.y; hn[6] = miss ? 0 : o.norm.z; hn[7] = 0;
            }
        }
    }
    return 0;
}
int ora_render(const ora_volume* v, const void* scninfo, int shade, int y0, int y1, uint8_t* out, float* hit_norm, int threads)
{
    return ora_render_ex(v, scninfo, shade, y0, y1, out, hit_norm, threads, 0, 1);
}

int ora_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1: let the caller ask for the host's cores explicitly */
void ora_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ============================================================================================ scenes */
int ora_scene_preset(const char* name, void* out, size_t bytes)
{
    if (bytes < sizeof(scene_preset)) return -2;
    return scene_get_preset(name, (scene_preset*)out);
}
int ora_scene_generate(const void* preset, int* nbricks, int32_t** brick_pos, float** values)
{
    scene_data d;
    int rc = scene_generate((const scene_preset*)preset, &d);
    *nbricks = d.nbricks; *brick_pos = d.brick_pos; *values = d.values;
    return rc;
}
void ora_free(void* p) { free(p); }
