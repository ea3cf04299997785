// gvdbx_extra.cuh — the remaining shade modes of VolumeGVDB::Render (gvdb_volume_gvdb.cpp:4363-4372) and the composed
// deep + shadow mode of BASELINE.json config 4, on the compact tables of gvdbx_device.cuh.
//
//   point query          cuda_gvdb_nodes.cuh:199-253   (getNode(lev,start,pos) / getNodeAtPoint)
//   tricubic sampling    cuda_gvdb_raycast.cuh:32-96   (getTricubic), :159-169 (getGradientTricubic), :316-339 (brick)
//   shadow accumulation  cuda_gvdb_raycast.cuh:445-463 (rayShadowBrick)
//   section kernels      cuda_gvdb_module.cu:225-298   (gvdbSection3D / gvdbSection2D), cuda_gvdb_geom.cuh:67-71
//
// Same floating-point contract as gvdbx_device.cuh: operand order, literal types and division forms of the reference
// are kept so that nvcc --use_fast_math makes the same contraction / approximation choices.
#pragma once
#include "gvdbx_device.cuh"
#include "gvdbx_trace.cuh"

// ------------------------------------------------------------------------------------------------ point query
// Top-down descent to the leaf that contains index-space point `pos`; returns the leaf index or -1.
// One int4 position record + one 4-B child-table entry per level instead of the reference's node record (64 B) + child
// list pointer chase.  top_lev == 0 (single brick): the reference returns leaf 0 without any bounds test; so do we.
template <class S>
__device__ __forceinline__ int gx_node_at_point(const GxParams& P, float3 pos, GxCount& cnt)
{
    int lev = P.top_lev;
    if (lev < 0 || lev >= GX_MAXLEV) return -1;
    int n = 0;
    if (lev == 0) return 0;
    int4 np = gx_node_pos(P, lev, 0);
    cnt.n_desc++;
    float3 vmin = make_float3(float(np.x), float(np.y), float(np.z));
    while (lev > 0) {
        const int3 nr = P.noderange[lev];
        const float3 vmax = vmin + make_float3(float(nr.x), float(nr.y), float(nr.z));
        if (pos.x < vmin.x || pos.y < vmin.y || pos.z < vmin.z || pos.x >= vmax.x || pos.y >= vmax.y || pos.z >= vmax.z) return -1;
        const float3 q = (pos - vmin) / gx_vdel<S>(P, lev);
        const int3 p = make_int3(int(q.x), int(q.y), int(q.z));
        const int dm = gx_dim<S>(P, lev);
        const int b = (((p.z << dm) + p.y) << dm) + p.x;
        // the approximate division can round a coordinate just below the upper face up to res: the reference then reads a
        // neighbouring cell (or past the list); such a point is treated as outside here
        if (unsigned(b) >= (1u << (3 * dm))) return -1;
        const int c = gx_child(gx_table(P, lev, n, dm), b);
        cnt.n_dda++;
        lev--;
        if (c == -1) return -1;
        n = c;
        if (lev > 0) {
            np = gx_node_pos(P, lev, n);
            cnt.n_desc++;
            vmin = make_float3(float(np.x), float(np.y), float(np.z));
        }
    }
    return n;
}

// ------------------------------------------------------------------------------------------------ tricubic
// 27 fetches at texel CORNERS (integer atlas coordinates: each is the mean of 8 texels), quadratic B-spline weights.
// Reaches one texel beyond the apron for p < 1, so neither the brick-major layout nor value-range culling applies.
template <class S>
__device__ __forceinline__ float gx_tricubic(const S& smp, float3 p, float3 offs, GxCount& cnt)
{
    const float MID = 1.0;
    const float HI = 2.0;
    float3 q = gx_floor(p + offs) - MID;
    float3 tb = (p - gx_floor(p)) * 0.5 + 0.25;
    float3 ta = (1.0 - tb);
    float3 ta2 = ta * ta;
    float3 tb2 = tb * tb;
    float3 tab = ta * tb * 2.0;
    cnt.s_tri += 27;

    float tv[9];
    tv[0] = smp.tri(q.x,       q.y,       q.z);
    tv[1] = smp.tri(q.x + MID, q.y,       q.z);
    tv[2] = smp.tri(q.x + HI,  q.y,       q.z);
    tv[3] = smp.tri(q.x,       q.y + MID, q.z);
    tv[4] = smp.tri(q.x + MID, q.y + MID, q.z);
    tv[5] = smp.tri(q.x + HI,  q.y + MID, q.z);
    tv[6] = smp.tri(q.x,       q.y + HI,  q.z);
    tv[7] = smp.tri(q.x + MID, q.y + HI,  q.z);
    tv[8] = smp.tri(q.x + HI,  q.y + HI,  q.z);
    float3 abc = make_float3(tv[0] * ta2.x + tv[1] * tab.x + tv[2] * tb2.x,
                             tv[3] * ta2.x + tv[4] * tab.x + tv[5] * tb2.x,
                             tv[6] * ta2.x + tv[7] * tab.x + tv[8] * tb2.x);
    tv[0] = smp.tri(q.x,       q.y,       q.z + MID);
    tv[1] = smp.tri(q.x + MID, q.y,       q.z + MID);
    tv[2] = smp.tri(q.x + HI,  q.y,       q.z + MID);
    tv[3] = smp.tri(q.x,       q.y + MID, q.z + MID);
    tv[4] = smp.tri(q.x + MID, q.y + MID, q.z + MID);
    tv[5] = smp.tri(q.x + HI,  q.y + MID, q.z + MID);
    tv[6] = smp.tri(q.x,       q.y + HI,  q.z + MID);
    tv[7] = smp.tri(q.x + MID, q.y + HI,  q.z + MID);
    tv[8] = smp.tri(q.x + HI,  q.y + HI,  q.z + MID);
    float3 def = make_float3(tv[0] * ta2.x + tv[1] * tab.x + tv[2] * tb2.x,
                             tv[3] * ta2.x + tv[4] * tab.x + tv[5] * tb2.x,
                             tv[6] * ta2.x + tv[7] * tab.x + tv[8] * tb2.x);
    tv[0] = smp.tri(q.x,       q.y,       q.z + HI);
    tv[1] = smp.tri(q.x + MID, q.y,       q.z + HI);
    tv[2] = smp.tri(q.x + HI,  q.y,       q.z + HI);
    tv[3] = smp.tri(q.x,       q.y + MID, q.z + HI);
    tv[4] = smp.tri(q.x + MID, q.y + MID, q.z + HI);
    tv[5] = smp.tri(q.x + HI,  q.y + MID, q.z + HI);
    tv[6] = smp.tri(q.x,       q.y + HI,  q.z + HI);
    tv[7] = smp.tri(q.x + MID, q.y + HI,  q.z + HI);
    tv[8] = smp.tri(q.x + HI,  q.y + HI,  q.z + HI);
    float3 ghi = make_float3(tv[0] * ta2.x + tv[1] * tab.x + tv[2] * tb2.x,
                             tv[3] * ta2.x + tv[4] * tab.x + tv[5] * tb2.x,
                             tv[6] * ta2.x + tv[7] * tab.x + tv[8] * tb2.x);
    float3 jkl = make_float3(abc.x * ta2.y + abc.y * tab.y + abc.z * tb2.y,
                             def.x * ta2.y + def.y * tab.y + def.z * tb2.y,
                             ghi.x * ta2.y + ghi.y * tab.y + ghi.z * tb2.y);
    return jkl.x * ta2.z + jkl.y * tab.z + jkl.z * tb2.z;
}

template <class S>
__device__ __forceinline__ float3 gx_gradient_tricubic(const S& smp, float3 p, float3 offs, GxCount& cnt)
{
    const float vs = 0.5;
    float3 g;
    g.x = (gx_tricubic(smp, p + make_float3(-vs, 0, 0), offs, cnt) - gx_tricubic(smp, p + make_float3(vs, 0, 0), offs, cnt)) / (2 * vs);
    g.y = (gx_tricubic(smp, p + make_float3(0, -vs, 0), offs, cnt) - gx_tricubic(smp, p + make_float3(0, vs, 0), offs, cnt)) / (2 * vs);
    g.z = (gx_tricubic(smp, p + make_float3(0, 0, -vs), offs, cnt) - gx_tricubic(smp, p + make_float3(0, 0, vs), offs, cnt)) / (2 * vs);
    return gx_normalize(g);
}

// SHADE_TRICUBIC brick function: fixed-step march (no start snap), first tricubic sample >= THRESH, one secant
// refinement step back along the ray.                                    cuda_gvdb_raycast.cuh:316-339
template <class S>
__device__ __forceinline__ void gx_brick_tricubic(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                  GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    float3 p = pos + t.x * dir - vmin;
    float3 v;

    for (int iter = 0; iter < GX_MAX_ITER && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0; iter++) {
        v.z = gx_tricubic(smp, p, o, cnt);
        if (v.z >= P.thresh.x) {
            v.x = gx_tricubic(smp, p - P.steps.z * dir, o, cnt);
            v.y = (v.z - P.thresh.x) / (v.z - v.x);
            p += -v.y * P.steps.z * dir;
            h.hit = p + vmin;
            h.norm = gx_gradient_tricubic(smp, p, o, cnt);
            h.t = t.x; h.leaf = nodeid; h.vox = gx_i3(gx_floor(h.hit));
            h.cpos = p + o;
            return;
        }
        p += P.steps.x * dir;
        t.x += P.steps.x;
    }
}

// ------------------------------------------------------------------------------------------------ shadow accumulation
// rayShadowBrick: opacity accumulates in clr.w from 0 towards 1; the ray parameter advances by DIRECTSTEP while the
// attenuation uses SHADOWSTEP / (1 + 0.4 t) — in DOUBLE, like the reference's literals make it.  No iteration cap.
template <class S>
__device__ __forceinline__ void gx_brick_shadow(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    t.x += P.epsilon;
    t.y -= P.epsilon;
    float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    float3 p = pos + t.x * dir - vmin;
    float3 pt = P.steps.x * dir;
    const float res0 = float(gx_res<S>(P, 0));
    const float inv_range = gx_rcp_approx(P.thresh.z - P.thresh.y);
    float4& clr = h.clr;
    float val = 0;

    for (; clr.w < 1 && p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < res0 && p.y < res0 && p.z < res0;) {
        cnt.s_tri++; cnt.s_lut++;
        const float4 T = __ldg(&P.transfer[gx_transfer_index(smp.tri(p.x + o.x, p.y + o.y, p.z + o.z), P.thresh.x, inv_range)]);
        val = exp(P.extinct.x * T.w * P.steps.y / (1.0 + t.x * 0.4));
        clr.w = 1.0 - (1.0 - clr.w) * val;
        p += pt;
        t.x += P.steps.y;
    }
}

// ------------------------------------------------------------------------------------------------ composed deep + shadow
// Primary emission/absorption march (gvdbRayDeep up to the composite), then — if any brick was entered — one shadow
// march from the first sample position towards the light whose accumulated opacity darkens the colour; composite over
// the background as gvdbRayDeep does.  The reference defines rayShadowBrick but never calls it; the composition is the one
// SURVEY.md §8c (ii) specifies, and the parity tests build the same thing from the reference's own device functions.
template <bool BATCH, class S>
__device__ __forceinline__ float4 gx_pixel_deepshadow(const GxParams& P, S& smp, float3 rpos, float3 rdir, int x, int y,
                                                      GxCount& cnt, GxHit& h, float4& raw)
{
    h.clr = make_float4(0, 0, 0, 1);
    h.hit = make_float3(0, 0, GX_NOHIT);
    gx_raycast<GX_MODE_DEEP, BATCH>(P, smp, rpos, rdir, h, cnt, x, y);
    float4 clr = h.clr;
    if (h.hit.x != 0.f) {
        float3 spos = rpos + rdir * h.hit.x;
        float3 ldir = gx_normalize(P.light_pos - spos);
        GxHit h2;
        h2.clr = make_float4(0, 0, 0, 0);
        h2.hit = make_float3(0, 0, GX_NOHIT);
        h2.norm = make_float3(0, 0, 0); h2.t = 0; h2.leaf = -1; h2.vox = make_int3(0, 0, 0);
        gx_raycast<GX_MODE_SHADOW, BATCH>(P, smp, spos, ldir, h2, cnt, x, y);
        float lit = 1.0f - h2.clr.w;
        clr.x *= lit; clr.y *= lit; clr.z *= lit;
    }
    raw = clr;
    float a = 1.0 - clr.w;
    return make_float4(P.backclr.x + a * (clr.x - P.backclr.x), P.backclr.y + a * (clr.y - P.backclr.y),
                       P.backclr.z + a * (clr.z - P.backclr.z), 1.0 - clr.w);
}

// ------------------------------------------------------------------------------------------------ sections
// gvdbSection2D: the frame is the slice_pnt + (u, 0, v) * slice_norm plane with u, v in [-1, 1); each pixel is the
// transfer colour of the trilinear value at that point over black.       cuda_gvdb_module.cu:272-298
template <class S>
__device__ __forceinline__ float4 gx_pixel_section2d(const GxParams& P, S& smp, int x, int y, GxCount& cnt)
{
    float3 bgclr = make_float3(0, 0, 0);
    float3 spnt = make_float3(float(x) * 2.0 / P.width - 1.0, 0, float(y) * 2.0 / P.height - 1.0);
    float3 wpos = P.slice_pnt + spnt * P.slice_norm;
    const int n = gx_node_at_point<S>(P, wpos, cnt);
    if (n < 0) return make_float4(bgclr.x, bgclr.y, bgclr.z, 1);
    const GxLeafRec L = gx_leaf(P, n);
    cnt.n_desc++;
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    const float3 offs = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float3 p = offs + (wpos - vmin);
    cnt.s_tri++; cnt.s_lut++;
    const float4 clr = __ldg(&P.transfer[gx_transfer_index(smp.tri(p.x, p.y, p.z), P.thresh.x, gx_rcp_approx(P.thresh.z - P.thresh.y))]);
    bgclr = make_float3(bgclr.x + clr.w * (clr.x - bgclr.x), bgclr.y + clr.w * (clr.y - bgclr.y), bgclr.z + clr.w * (clr.z - bgclr.z));
    return make_float4(bgclr.x, bgclr.y, bgclr.z, 1);
}

// gvdbSection3D: transfer colour on the section plane blended over the trilinear surface found behind it.
//                                                                         cuda_gvdb_module.cu:225-269
template <bool BATCH, class S>
__device__ __forceinline__ float4 gx_pixel_section3d(const GxParams& P, S& smp, float3 rpos, float3 rdir, int x, int y,
                                                     GxCount& cnt, GxHit& h)
{
    float4 clr = make_float4(1, 1, 1, 0);
    float3 wpos = rpos;
    // rayPlaneIntersect (cuda_gvdb_geom.cuh:67-71): NOHIT (> 0!) when the plane lies behind the ray
    const float3 pn = P.slice_norm, pp = P.slice_pnt;
    float t = ((pp.x - wpos.x) * pn.x + (pp.y - wpos.y) * pn.y + (pp.z - wpos.z) * pn.z) / (rdir.x * pn.x + rdir.y * pn.y + rdir.z * pn.z);
    t = (t > 0 ? t : GX_NOHIT);
    if (t > 0) {
        wpos += t * rdir;
        const int n = gx_node_at_point<S>(P, wpos, cnt);
        if (n >= 0) {
            const GxLeafRec L = gx_leaf(P, n);
            cnt.n_desc++;
            smp.enter(L);
            const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
            const float3 offs = make_float3(float(L.vx), float(L.vy), float(L.vz));
            const float3 p = offs + (wpos - vmin);
            cnt.s_tri++; cnt.s_lut++;
            t = smp.tri(p.x, p.y, p.z);
            clr = __ldg(&P.transfer[gx_transfer_index(t, P.thresh.x, gx_rcp_approx(P.thresh.z - P.thresh.y))]);
            if (P.clr_tex) {                // the section plane takes the voxel's colour too (cuda_gvdb_module.cu:249-252); alpha * 1.0
                const float4 c = gx_color(P, p);
                clr.x *= c.x; clr.y *= c.y; clr.z *= c.z;
            }
        } else {
            t = 0;
        }
    }
    h.hit = make_float3(GX_NOHIT, GX_NOHIT, GX_NOHIT);
    h.clr = make_float4(1, 1, 1, 1);
    gx_raycast<GX_MODE_TRILINEAR, BATCH>(P, smp, wpos, rdir, h, cnt, x, y);
    gx_hit_color(P, h);
    if (h.hit.z != GX_NOHIT) {
        float3 lightdir = gx_normalize(P.light_pos - h.hit);
        float ds = (t > P.thresh.x) ? 1 : 0.8 * fmaxf(0.0f, gx_dot(h.norm, lightdir));
        const float4 a = make_float4(h.clr.x * ds, h.clr.y * ds, h.clr.z * ds, h.clr.w * ds);
        clr = make_float4(a.x + clr.w * (clr.x - a.x), a.y + clr.w * (clr.y - a.y), a.z + clr.w * (clr.z - a.z), a.w + clr.w * (clr.w - a.w));
    } else {
        const float4 a = P.backclr;
        clr = make_float4(a.x + clr.w * (clr.x - a.x), a.y + clr.w * (clr.y - a.y), a.z + clr.w * (clr.z - a.z), a.w + clr.w * (clr.w - a.w));
    }
    return clr;
}
