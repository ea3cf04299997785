/* scenes.h — deterministic synthetic volumes + render presets (TEST INFRASTRUCTURE, not product).
 *
 * Header-only plain C so the same generator is compiled into
 *   - oracle/liboracle.so        (CPU restatement, used by tests / bench.py cpu_baseline leg)
 *   - oracle/_ref/ref_harness    (drives the UNMODIFIED reference through its own public API)
 * Both sides therefore see bit-identical brick lists and voxel values.
 *
 * Inputs follow SURVEY.md §8(d): tree Configure(3,3,3,3,3) (8^3 bricks), channel 0 T_FLOAT, apron 1.
 * A scene is (a) an ordered list of brick min-corners (index space, multiples of 8) — the order in which
 * ActivateSpace is called (reference: gvdb_volume_gvdb.cpp:2766) — and (b) 512 voxel values per brick
 * (x fastest, then y, then z), the brick's *interior* voxels.
 */
#ifndef GVDBX_SCENES_H
#define GVDBX_SCENES_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SCN_KIND_SPHERE = 0, SCN_KIND_SDF = 1, SCN_KIND_BALLS = 2, SCN_KIND_CLOUD = 3 };

/* shade modes — numeric values of the reference enum (gvdb_types.h:73-81 / cuda_gvdb_scene.cuh:25-32) */
enum { SCN_SHADE_VOXEL = 0, SCN_SHADE_TRILINEAR = 4, SCN_SHADE_LEVELSET = 6, SCN_SHADE_VOLUME = 7 };

typedef struct scene_preset {
    char  name[32];
    int   kind;          /* SCN_KIND_* */
    int   N;             /* index-space extent (cube N^3), multiple of 8 */
    float a, b, c;       /* kind-specific parameters (see generators) */
    int   width, height; /* render size */
    int   shade;         /* SCN_SHADE_* */
    float fov;
    float cam_angs[3], cam_target[3], cam_dist;
    float light_angs[3], light_target[3], light_dist;
    float steps[3], extinct[3], thresh[3], cutoff[3], backclr[4], shadow[3];
    float epsilon;
    int   transfer;      /* 0 = Initialize() default ramp, 1 = gRenderToFile 4-ramp table */
} scene_preset;

typedef struct scene_data {
    int      nbricks;
    int32_t* brick_pos;  /* 3*nbricks */
    float*   values;     /* 512*nbricks */
} scene_data;

/* ------------------------------------------------------------------ PRNG / hashing */
static inline uint64_t scn_splitmix64(uint64_t* s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline float scn_u01(uint64_t* s) { return (float)(scn_splitmix64(s) >> 40) * (1.0f / 16777216.0f); }

static inline float scn_lattice(int x, int y, int z, uint32_t seed)
{
    uint32_t h = (uint32_t)x * 0x8da6b343u ^ (uint32_t)y * 0xd8163841u ^ (uint32_t)z * 0xcb1ab31fu ^ seed;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12; h *= 0x297a2d39u; h ^= h >> 15;
    return (float)(h >> 8) * (1.0f / 16777216.0f);   /* [0,1) */
}
/* trilinear value noise with smoothstep fade, period `per` voxels; returns [0,1) */
static inline float scn_vnoise(float x, float y, float z, float per, uint32_t seed)
{
    float fx = x / per, fy = y / per, fz = z / per;
    int ix = (int)floorf(fx), iy = (int)floorf(fy), iz = (int)floorf(fz);
    float tx = fx - ix, ty = fy - iy, tz = fz - iz;
    tx = tx * tx * (3.f - 2.f * tx); ty = ty * ty * (3.f - 2.f * ty); tz = tz * tz * (3.f - 2.f * tz);
    float c000 = scn_lattice(ix, iy, iz, seed),     c100 = scn_lattice(ix + 1, iy, iz, seed);
    float c010 = scn_lattice(ix, iy + 1, iz, seed), c110 = scn_lattice(ix + 1, iy + 1, iz, seed);
    float c001 = scn_lattice(ix, iy, iz + 1, seed), c101 = scn_lattice(ix + 1, iy, iz + 1, seed);
    float c011 = scn_lattice(ix, iy + 1, iz + 1, seed), c111 = scn_lattice(ix + 1, iy + 1, iz + 1, seed);
    float x00 = c000 + tx * (c100 - c000), x10 = c010 + tx * (c110 - c010);
    float x01 = c001 + tx * (c101 - c001), x11 = c011 + tx * (c111 - c011);
    float y0 = x00 + ty * (x10 - x00), y1 = x01 + ty * (x11 - x01);
    return y0 + tz * (y1 - y0);
}
static inline float scn_fbm3(float x, float y, float z, float per, uint32_t seed)
{   /* 3 octaves, amplitudes 1, 1/2, 1/4, normalised to [0,1) */
    float v = scn_vnoise(x, y, z, per, seed) + 0.5f * scn_vnoise(x, y, z, per * 0.5f, seed + 1u)
            + 0.25f * scn_vnoise(x, y, z, per * 0.25f, seed + 2u);
    return v * (1.0f / 1.75f);
}

/* ------------------------------------------------------------------ per-kind voxel functions
 * voxel (i,j,k) is evaluated at its centre (i+.5, j+.5, k+.5). */
static inline float scn_sphere_val(const scene_preset* p, int i, int j, int k)
{   /* density d = clamp(1 - r/R, 0, 1); a = R */
    float c = 0.5f * (float)p->N;
    float dx = i + 0.5f - c, dy = j + 0.5f - c, dz = k + 0.5f - c;
    float d = 1.0f - sqrtf(dx * dx + dy * dy + dz * dz) / p->a;
    return d < 0.f ? 0.f : (d > 1.f ? 1.f : d);
}
static inline float scn_sdf_raw(const scene_preset* p, float x, float y, float z)
{   /* signed distance to sphere radius a, displaced by fbm amplitude b, base period c */
    float cc = 0.5f * (float)p->N;
    float dx = x - cc, dy = y - cc, dz = z - cc;
    float r = sqrtf(dx * dx + dy * dy + dz * dz);
    return r - p->a - p->b * (scn_fbm3(x, y, z, p->c, 0x5EEDu) - 0.5f);
}
static inline float scn_cloud_val(const scene_preset* p, int i, int j, int k)
{   /* ball radius a with fbm-modulated density, period c; b = noise weight in [0,1] */
    float c = 0.5f * (float)p->N;
    float x = i + 0.5f, y = j + 0.5f, z = k + 0.5f;
    float dx = x - c, dy = y - c, dz = z - c;
    float r = sqrtf(dx * dx + dy * dy + dz * dz);
    float fall = 1.0f - r / p->a;
    if (fall <= 0.f) return 0.f;
    if (fall > 1.f) fall = 1.f;
    float n = scn_fbm3(x, y, z, p->c, 0xC10Du);
    float d = fall * ((1.0f - p->b) + p->b * 2.0f * n);
    return d < 0.f ? 0.f : (d > 1.f ? 1.f : d);
}

/* ------------------------------------------------------------------ generators */
typedef struct { float x, y, z, r; } scn_ball;

static inline void scn_push_brick(scene_data* d, int* cap, int bx, int by, int bz, const float* vals)
{
    if (d->nbricks == *cap) {
        *cap = *cap ? *cap * 2 : 1024;
        d->brick_pos = (int32_t*)realloc(d->brick_pos, sizeof(int32_t) * 3 * (size_t)*cap);
        d->values    = (float*)realloc(d->values, sizeof(float) * 512 * (size_t)*cap);
    }
    d->brick_pos[3 * d->nbricks + 0] = bx; d->brick_pos[3 * d->nbricks + 1] = by; d->brick_pos[3 * d->nbricks + 2] = bz;
    memcpy(d->values + 512 * (size_t)d->nbricks, vals, sizeof(float) * 512);
    d->nbricks++;
}

static inline int scene_generate(const scene_preset* p, scene_data* out)
{
    int cap = 0, nb = p->N / 8;
    float vals[512];
    out->nbricks = 0; out->brick_pos = NULL; out->values = NULL;

    if (p->kind == SCN_KIND_BALLS) {
        /* union of `a` random solid balls, radius in [b, c]; value 1 inside, 0 outside */
        int nballs = (int)p->a;
        uint64_t s = 0x9E3779B97F4A7C15ULL;
        scn_ball* balls = (scn_ball*)malloc(sizeof(scn_ball) * (size_t)nballs);
        for (int n = 0; n < nballs; n++) {
            float r = p->b + (p->c - p->b) * scn_u01(&s);
            balls[n].r = r;
            balls[n].x = r + 1.f + ((float)p->N - 2.f * r - 2.f) * scn_u01(&s);
            balls[n].y = r + 1.f + ((float)p->N - 2.f * r - 2.f) * scn_u01(&s);
            balls[n].z = r + 1.f + ((float)p->N - 2.f * r - 2.f) * scn_u01(&s);
        }
        /* CSR of balls per brick */
        size_t nb3 = (size_t)nb * nb * nb;
        uint32_t* cnt = (uint32_t*)calloc(nb3 + 1, sizeof(uint32_t));
        for (int pass = 0; pass < 2; pass++) {
            uint32_t* lst = NULL; uint32_t* fill = NULL;
            if (pass == 1) {
                uint32_t acc = 0;
                for (size_t i = 0; i <= nb3; i++) { uint32_t t = cnt[i]; cnt[i] = acc; acc += t; }
                lst = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(cnt[nb3] + 1));
                fill = (uint32_t*)calloc(nb3, sizeof(uint32_t));
            }
            for (int n = 0; n < nballs; n++) {
                int x0 = (int)floorf((balls[n].x - balls[n].r) / 8.f), x1 = (int)floorf((balls[n].x + balls[n].r) / 8.f);
                int y0 = (int)floorf((balls[n].y - balls[n].r) / 8.f), y1 = (int)floorf((balls[n].y + balls[n].r) / 8.f);
                int z0 = (int)floorf((balls[n].z - balls[n].r) / 8.f), z1 = (int)floorf((balls[n].z + balls[n].r) / 8.f);
                if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0; if (z0 < 0) z0 = 0;
                if (x1 >= nb) x1 = nb - 1; if (y1 >= nb) y1 = nb - 1; if (z1 >= nb) z1 = nb - 1;
                for (int z = z0; z <= z1; z++) for (int y = y0; y <= y1; y++) for (int x = x0; x <= x1; x++) {
                    size_t id = ((size_t)z * nb + y) * nb + x;
                    if (pass == 0) cnt[id]++; else lst[cnt[id] + fill[id]++] = (uint32_t)n;
                }
            }
            if (pass == 1) {
                for (int bz = 0; bz < nb; bz++) for (int by = 0; by < nb; by++) for (int bx = 0; bx < nb; bx++) {
                    size_t id = ((size_t)bz * nb + by) * nb + bx;
                    uint32_t b0 = cnt[id], b1 = cnt[id + 1];
                    if (b0 == b1) continue;
                    int any = 0;
                    for (int k = 0; k < 8; k++) for (int j = 0; j < 8; j++) for (int i = 0; i < 8; i++) {
                        float x = bx * 8 + i + 0.5f, y = by * 8 + j + 0.5f, z = bz * 8 + k + 0.5f, v = 0.f;
                        for (uint32_t q = b0; q < b1; q++) {
                            const scn_ball* B = &balls[lst[q]];
                            float dx = x - B->x, dy = y - B->y, dz = z - B->z;
                            if (dx * dx + dy * dy + dz * dz <= B->r * B->r) { v = 1.f; any = 1; break; }
                        }
                        vals[(k * 8 + j) * 8 + i] = v;
                    }
                    if (any) scn_push_brick(out, &cap, bx * 8, by * 8, bz * 8, vals);
                }
                free(lst); free(fill);
            }
        }
        free(cnt); free(balls);
        return 0;
    }

    /* one z-slab of bricks per task (OpenMP when the including file is compiled with it); slabs are concatenated in z
     * order afterwards, so the brick order — the ActivateSpace order — does not depend on the thread count */
    scene_data* slab = (scene_data*)calloc((size_t)nb, sizeof(scene_data));
    #pragma omp parallel for schedule(dynamic, 1)
    for (int bz = 0; bz < nb; bz++) {
        int scap = 0;
        float v8[512];
        scene_data* out_s = &slab[bz];
        for (int by = 0; by < nb; by++) for (int bx = 0; bx < nb; bx++) {
        if (p->kind == SCN_KIND_SPHERE || p->kind == SCN_KIND_CLOUD) {
            /* cheap reject: brick farther than R + brick diagonal from the centre */
            float c = 0.5f * (float)p->N;
            float dx = bx * 8 + 4.f - c, dy = by * 8 + 4.f - c, dz = bz * 8 + 4.f - c;
            float R = p->a + 7.0f;
            if (dx * dx + dy * dy + dz * dz > R * R) continue;
            int any = 0;
            for (int k = 0; k < 8; k++) for (int j = 0; j < 8; j++) for (int i = 0; i < 8; i++) {
                float v = (p->kind == SCN_KIND_SPHERE) ? scn_sphere_val(p, bx * 8 + i, by * 8 + j, bz * 8 + k)
                                                       : scn_cloud_val(p, bx * 8 + i, by * 8 + j, bz * 8 + k);
                v8[(k * 8 + j) * 8 + i] = v;
                any |= (v > 0.f);
            }
            if (any) scn_push_brick(out_s, &scap, bx * 8, by * 8, bz * 8, v8);
        } else { /* SCN_KIND_SDF: narrow band, half-width `band` = 12 voxels tested at the brick centre */
            const float band = 12.0f;
            float c = 0.5f * (float)p->N;
            float dx = bx * 8 + 4.f - c, dy = by * 8 + 4.f - c, dz = bz * 8 + 4.f - c;
            float r = sqrtf(dx * dx + dy * dy + dz * dz);
            if (fabsf(r - p->a) > band + 0.5f * p->b + 7.0f) continue;
            float sc = scn_sdf_raw(p, bx * 8 + 4.f, by * 8 + 4.f, bz * 8 + 4.f);
            if (fabsf(sc) > band) continue;
            for (int k = 0; k < 8; k++) for (int j = 0; j < 8; j++) for (int i = 0; i < 8; i++) {
                float v = scn_sdf_raw(p, bx * 8 + i + 0.5f, by * 8 + j + 0.5f, bz * 8 + k + 0.5f);
                v = v < -band ? -band : (v > band ? band : v);
                v8[(k * 8 + j) * 8 + i] = v;
            }
            scn_push_brick(out_s, &scap, bx * 8, by * 8, bz * 8, v8);
        }
        }
    }
    size_t total = 0;
    for (int bz = 0; bz < nb; bz++) total += (size_t)slab[bz].nbricks;
    out->brick_pos = (int32_t*)malloc(sizeof(int32_t) * 3 * (total ? total : 1));
    out->values = (float*)malloc(sizeof(float) * 512 * (total ? total : 1));
    for (int bz = 0; bz < nb; bz++) {
        size_t n = (size_t)slab[bz].nbricks;
        if (n) {
            memcpy(out->brick_pos + 3 * (size_t)out->nbricks, slab[bz].brick_pos, sizeof(int32_t) * 3 * n);
            memcpy(out->values + 512 * (size_t)out->nbricks, slab[bz].values, sizeof(float) * 512 * n);
            out->nbricks += (int)n;
        }
        free(slab[bz].brick_pos); free(slab[bz].values);
    }
    free(slab);
    (void)cap; (void)vals;
    return 0;
}

static inline void scene_free(scene_data* d) { free(d->brick_pos); free(d->values); memset(d, 0, sizeof(*d)); }

/* ------------------------------------------------------------------ presets
 * cfg1..cfg4 follow BASELINE.json configs[0..3] / SURVEY.md §8(d); *_small are the same generators at sizes the
 * CPU oracle renders in seconds. */
static inline void scn_set3(float* d, float x, float y, float z) { d[0] = x; d[1] = y; d[2] = z; }

static inline int scene_get_preset(const char* name, scene_preset* p)
{
    memset(p, 0, sizeof(*p));
    strncpy(p->name, name, sizeof(p->name) - 1);
    /* defaults = reference Scene()/Camera3D() defaults (gvdb_scene.cpp:28-53, gvdb_camera.cpp:42-54) */
    p->fov = 40.f; p->epsilon = 0.001f;
    scn_set3(p->steps, 1.0f, 16.f, 0.1f); scn_set3(p->extinct, -1.1f, 1.5f, 0.f);
    scn_set3(p->thresh, 0.1f, 0.f, 1.f);  scn_set3(p->cutoff, 0.005f, 0.01f, 0.f);
    scn_set3(p->shadow, 0.8f, 1.0f, 0.f);
    p->backclr[0] = 0.1f; p->backclr[1] = 0.2f; p->backclr[2] = 0.4f; p->backclr[3] = 1.0f;
    scn_set3(p->light_angs, 299.f, 57.3f, 0.f); p->light_dist = 200.f;

    int small = strstr(name, "_small") != NULL;
    int tiny  = strstr(name, "_tiny") != NULL;
    if (!strncmp(name, "cfg1", 4)) {            /* sphere density, SHADE_TRILINEAR, 1024x768 */
        p->kind = SCN_KIND_SPHERE; p->N = tiny ? 64 : (small ? 128 : 256);
        float s = p->N / 256.f;
        p->a = 100.f * s; p->width = tiny ? 96 : (small ? 256 : 1024); p->height = tiny ? 72 : (small ? 192 : 768);
        p->shade = SCN_SHADE_TRILINEAR; p->fov = 50.f;
        scn_set3(p->cam_angs, 20.f, 30.f, 0.f); scn_set3(p->cam_target, 128.f * s, 128.f * s, 128.f * s); p->cam_dist = 500.f * s;
        scn_set3(p->light_target, 132.f * s, -20.f * s, 50.f * s); p->light_dist = 200.f * s;
        scn_set3(p->steps, .25f, 16.f, .25f); scn_set3(p->thresh, 0.1f, 0.f, 1.f); scn_set3(p->cutoff, .005f, .01f, 0.f);
    } else if (!strncmp(name, "cfg2", 4)) {     /* noise-displaced sphere SDF, SHADE_LEVELSET, 1920x1080 */
        p->kind = SCN_KIND_SDF; p->N = tiny ? 64 : (small ? 128 : 1024);
        float s = p->N / 1024.f;
        p->a = 300.f * s; p->b = 24.f * (small || tiny ? 0.25f : 1.f); p->c = 64.f * (tiny ? 0.25f : (small ? 0.5f : 1.f));
        p->width = tiny ? 96 : (small ? 320 : 1920); p->height = tiny ? 54 : (small ? 180 : 1080);
        p->shade = SCN_SHADE_LEVELSET; p->fov = 40.f;
        scn_set3(p->cam_angs, 35.f, 25.f, 0.f); scn_set3(p->cam_target, 512.f * s, 512.f * s, 512.f * s); p->cam_dist = 1800.f * s;
        scn_set3(p->light_target, 528.f * s, -80.f * s, 200.f * s); p->light_dist = 800.f * s;
        scn_set3(p->steps, .25f, 16.f, .25f); scn_set3(p->thresh, 0.0f, -3.f, 3.f); p->epsilon = 0.01f;
    } else if (!strncmp(name, "cfg3", 4)) {     /* 600 random solid balls r in [12,40] (~1 % occupancy, ~170 k bricks), SHADE_VOXEL, 3840x2160 */
        p->kind = SCN_KIND_BALLS; p->N = tiny ? 64 : (small ? 256 : 2048);
        float s = p->N / 2048.f;
        p->a = tiny ? 6.f : (small ? 48.f : 600.f); p->b = tiny ? 4.f : (small ? 8.f : 12.f); p->c = tiny ? 9.f : (small ? 20.f : 40.f);
        p->width = tiny ? 96 : (small ? 384 : 3840); p->height = tiny ? 54 : (small ? 216 : 2160);
        p->shade = SCN_SHADE_VOXEL; p->fov = 40.f;
        scn_set3(p->cam_angs, 60.f, 20.f, 0.f); scn_set3(p->cam_target, 1024.f * s, 1024.f * s, 1024.f * s); p->cam_dist = 3600.f * s;
        scn_set3(p->light_target, 1056.f * s, -160.f * s, 400.f * s); p->light_dist = 1600.f * s;
        scn_set3(p->thresh, 0.5f, 0.f, 1.f);
    } else if (!strncmp(name, "cfg4", 4)) {     /* noise cloud of radius 400 in a 1024^3 index space, deep volume (SHADE_VOLUME), 3840x2160 */
        p->kind = SCN_KIND_CLOUD; p->N = tiny ? 64 : (small ? 128 : 1024);
        float s = p->N / 512.f;
        p->a = 200.f * s; p->b = 0.6f; p->c = 64.f * (tiny ? 0.25f : (small ? 0.5f : 1.f));
        p->width = tiny ? 96 : (small ? 384 : 3840); p->height = tiny ? 54 : (small ? 216 : 2160);
        p->shade = SCN_SHADE_VOLUME; p->fov = 50.f;
        scn_set3(p->cam_angs, 20.f, 30.f, 0.f); scn_set3(p->cam_target, 256.f * s, 256.f * s, 256.f * s); p->cam_dist = 1000.f * s;
        scn_set3(p->light_target, 264.f * s, -40.f * s, 100.f * s); p->light_dist = 400.f * s;
        scn_set3(p->steps, .25f, 16.f, .25f); scn_set3(p->extinct, -1.0f, 1.5f, 0.f);
        scn_set3(p->thresh, 0.1f, 0.f, 1.f); scn_set3(p->cutoff, .005f, .01f, 0.f);
        p->transfer = 1;
    } else if (!strncmp(name, "cfg5", 4)) {     /* large noise cloud (~2.0 M bricks, ~8 GB atlas at full size), deep, 4 rays per pixel, 7680x4320 */
        p->kind = SCN_KIND_CLOUD; p->N = tiny ? 64 : (small ? 256 : 4096);
        float s = p->N / 4096.f;
        p->a = 625.f * s; p->b = 0.6f; p->c = 256.f * (tiny ? 0.0625f : (small ? 0.125f : 1.f));
        p->width = tiny ? 96 : (small ? 480 : 7680); p->height = tiny ? 54 : (small ? 270 : 4320);
        p->shade = SCN_SHADE_VOLUME; p->fov = 40.f;
        scn_set3(p->cam_angs, 30.f, 25.f, 0.f); scn_set3(p->cam_target, 2048.f * s, 2048.f * s, 2048.f * s); p->cam_dist = 7000.f * s;
        scn_set3(p->light_target, 2112.f * s, -320.f * s, 800.f * s); p->light_dist = 3200.f * s;
        scn_set3(p->steps, .5f, 16.f, .5f); scn_set3(p->extinct, -1.0f, 1.5f, 0.f);
        scn_set3(p->thresh, 0.1f, 0.f, 1.f); scn_set3(p->cutoff, .005f, .01f, 0.f);
        p->transfer = 1;
    } else {
        return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------ explicit ray bundles (VolumeGVDB::Raytrace)
 * n ScnRay records of 64 bytes (src/gvdb_volume_gvdb.h:300-308): hit@0 normal@12 orig@24 dir@36 clr@48 pnode@52 pndx@56.
 * Origins on a shell of radius 0.9 N around the volume centre, directions towards random points of the central half. */
static inline void scene_make_rays(const scene_preset* p, int n, void* out)
{
    uint64_t s = 0xA5A5A5A55A5A5A5AULL;
    float c = 0.5f * (float)p->N, R = 0.9f * (float)p->N;
    for (int i = 0; i < n; i++) {
        float* r = (float*)((char*)out + 64 * (size_t)i);
        memset(r, 0, 64);
        float u = 2.f * scn_u01(&s) - 1.f, ph = 6.2831853f * scn_u01(&s);
        float q = sqrtf(1.f - u * u);
        float ox = c + R * q * cosf(ph), oy = c + R * u, oz = c + R * q * sinf(ph);
        float tx = c + 0.5f * (float)p->N * (scn_u01(&s) - 0.5f), ty = c + 0.5f * (float)p->N * (scn_u01(&s) - 0.5f),
              tz = c + 0.5f * (float)p->N * (scn_u01(&s) - 0.5f);
        float dx = tx - ox, dy = ty - oy, dz = tz - oz;
        float il = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
        r[3] = 7.f; r[4] = 8.f; r[5] = 9.f;               /* normal: left untouched by the kernel on a miss */
        r[6] = ox; r[7] = oy; r[8] = oz;
        r[9] = dx * il; r[10] = dy * il; r[11] = dz * il;
        ((uint32_t*)r)[12] = 0xFF00FF00u + (uint32_t)i;
    }
}

#ifdef __cplusplus
}
#endif
#endif /* GVDBX_SCENES_H */
