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
// One float4 position record + one 4-B child-table entry per level instead of the reference's node record (64 B) + child
// list pointer chase.  top_lev == 0 (single brick): the reference returns leaf 0 without any bounds test; so do we.
template <class S>
__device__ __forceinline__ int gx_node_at_point(const GxParams& P, float3 pos, GxCount& cnt)
{
    int lev = P.top_lev;
    if (lev < 0 || lev >= GX_MAXLEV) return -1;
    int n = 0;
    if (lev == 0) return 0;
    gx_npos_t np = gx_node_pos(P, lev, 0);
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
// Quadratic B-spline reconstruction over a 3x3x3 stencil of texel CORNERS (integer atlas coordinates: each corner fetch is
// the hardware mean of its eight texels).  Separable: weights w(t) = {(1-t)^2, 2 t (1-t), t^2} with t = frac/2 + 1/4 per axis,
// reduced along x inside a row, along y inside a slice, along z across the three slices.  For p < 1 the stencil reaches one
// texel beyond the apron, so neither the brick-major layout nor value-range culling applies to this mode.
// (what is computed: getTricubic, cuda_gvdb_raycast.cuh:32-96; bit-exactness fixes the operand order of every sum)
struct GxSplineW { float3 lo, mid, hi; };              // per axis: weight of stencil index 0, 1, 2

__device__ __forceinline__ GxSplineW gx_spline_weights(float3 p)
{
    const float3 t = (p - gx_floor(p)) * 0.5 + 0.25;
    const float3 u = (1.0 - t);
    GxSplineW w;
    w.lo = u * u; w.hi = t * t; w.mid = u * t * 2.0;
    return w;
}
// one row of three corners along x at (y, z), reduced with the x weights
template <class S>
__device__ __forceinline__ float gx_spline_row(const S& smp, float x0, float y, float z, const GxSplineW& w)
{
    const float c0 = smp.tri(x0, y, z), c1 = smp.tri(x0 + 1.0f, y, z), c2 = smp.tri(x0 + 2.0f, y, z);
    return c0 * w.lo.x + c1 * w.mid.x + c2 * w.hi.x;
}
// one z slice: three rows, reduced with the y weights
template <class S>
__device__ __forceinline__ float gx_spline_slice(const S& smp, float3 q, float z, const GxSplineW& w)
{
    const float r0 = gx_spline_row(smp, q.x, q.y, z, w);
    const float r1 = gx_spline_row(smp, q.x, q.y + 1.0f, z, w);
    const float r2 = gx_spline_row(smp, q.x, q.y + 2.0f, z, w);
    return r0 * w.lo.y + r1 * w.mid.y + r2 * w.hi.y;
}
template <class S>
__device__ __forceinline__ float gx_tricubic(const S& smp, float3 p, float3 offs, GxCount& cnt)
{
    const float3 q = gx_floor(p + offs) - 1.0f;        // corner (0,0,0) of the stencil
    const GxSplineW w = gx_spline_weights(p);
    cnt.s_tri += 27;
    const float s0 = gx_spline_slice(smp, q, q.z, w);
    const float s1 = gx_spline_slice(smp, q, q.z + 1.0f, w);
    const float s2 = gx_spline_slice(smp, q, q.z + 2.0f, w);
    return s0 * w.lo.z + s1 * w.mid.z + s2 * w.hi.z;
}

// central difference of the tricubic field over +-0.5 voxel along one axis (backward minus forward: the surface normal
// points against the gradient), cuda_gvdb_raycast.cuh:159-169
template <int AXIS, class S>
__device__ __forceinline__ float gx_tricubic_slope(const S& smp, float3 p, float3 offs, GxCount& cnt)
{
    const float h = 0.5;
    float3 back = p, fwd = p;
    if (AXIS == 0) { back.x = p.x + -h; fwd.x = p.x + h; }
    if (AXIS == 1) { back.y = p.y + -h; fwd.y = p.y + h; }
    if (AXIS == 2) { back.z = p.z + -h; fwd.z = p.z + h; }
    return (gx_tricubic(smp, back, offs, cnt) - gx_tricubic(smp, fwd, offs, cnt)) / (2 * h);
}
template <class S>
__device__ __forceinline__ float3 gx_gradient_tricubic(const S& smp, float3 p, float3 offs, GxCount& cnt)
{
    float3 g;
    g.x = gx_tricubic_slope<0>(smp, p, offs, cnt);
    g.y = gx_tricubic_slope<1>(smp, p, offs, cnt);
    g.z = gx_tricubic_slope<2>(smp, p, offs, cnt);
    return gx_normalize(g);
}

// SHADE_TRICUBIC brick function (what: raySurfaceTricubicBrick, cuda_gvdb_raycast.cuh:316-339): fixed-step march from the
// unsnapped entry point; the first sample at or above THRESH is pulled back along the ray by the secant through it and the
// sample one FINESTEP before it.  In-brick test on the float bit patterns like the other marchers (gvdbx_trace.cuh).
template <class S>
__device__ __forceinline__ void gx_brick_tricubic(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                  GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    const float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    const float thr = P.thresh.x, fine = P.steps.z;
    float3 p = gx_poszero(pos + t.x * dir - vmin);
    for (int iter = 0; iter < GX_MAX_ITER && GX_INB(p, res0); iter++) {
        const float here = gx_tricubic(smp, p, o, cnt);
        if (here >= thr) {
            const float before = gx_tricubic(smp, p - fine * dir, o, cnt);
            const float back = (here - thr) / (here - before);          // fraction of a fine step to retreat
            p += -back * fine * dir;
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
// Opacity towards the light (what: rayShadowBrick, cuda_gvdb_raycast.cuh:445-463): every sample is an opaque layer of
// transparency exp(EXTINCT * alpha(v) * SHADOWSTEP / (1 + 0.4 s)) — evaluated in DOUBLE, as the reference's literals make it —
// where the attenuation parameter s advances by SHADOWSTEP per sample while the position advances by DIRECTSTEP; opacity
// accumulates in clr.w from 0 towards 1; no iteration cap.  Four samples per round like the other marchers: positions by
// the same chain of additions, the four texture fetches and four transfer-table reads in flight before the first layer is
// applied, layers applied strictly in order, a round cut short exactly where the one-at-a-time loop would stop.
template <class S>
__device__ __forceinline__ void gx_brick_shadow(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    const float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    const float inv_range = gx_rcp_approx(P.thresh.z - P.thresh.y);
    const float3 wpt = P.steps.x * dir;
    float s = t.x + P.epsilon;                          // the brick is entered epsilon further in than rayCast handed over
    float3 p = gx_poszero(pos + s * dir - vmin);
    float& opacity = h.clr.w;

    while (opacity < 1 && GX_INB(p, res0)) {
        float3 p1, p2, p3;
        GX_STEP_ADD(p1, p); GX_STEP_ADD(p2, p1); GX_STEP_ADD(p3, p2);
        const bool k1 = GX_INB(p1, res0), k2 = GX_INB(p2, res0), k3 = GX_INB(p3, res0);
        const float v0 = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);       // the loop condition holds for the first sample
        const float v1 = GX_TRI_IF(k1, p1), v2 = GX_TRI_IF(k2, p2), v3 = GX_TRI_IF(k3, p3);
        const float a0 = gx_lut(P.transfer, gx_transfer_index(v0, P.thresh.x, inv_range)).w;
#if GX_PRED_FETCH
        const float a1 = k1 ? gx_lut(P.transfer, gx_transfer_index(v1, P.thresh.x, inv_range)).w : 0.f;
        const float a2 = k2 ? gx_lut(P.transfer, gx_transfer_index(v2, P.thresh.x, inv_range)).w : 0.f;
        const float a3 = k3 ? gx_lut(P.transfer, gx_transfer_index(v3, P.thresh.x, inv_range)).w : 0.f;
#else
        const float a1 = gx_lut(P.transfer, gx_transfer_index(v1, P.thresh.x, inv_range)).w;
        const float a2 = gx_lut(P.transfer, gx_transfer_index(v2, P.thresh.x, inv_range)).w;
        const float a3 = gx_lut(P.transfer, gx_transfer_index(v3, P.thresh.x, inv_range)).w;
#endif
        // one layer: clr.w = 1 - (1 - clr.w) * exp(...), then the parameter step
        #define GX_SHADOW_LAYER(alpha) { cnt.s_tri++; cnt.s_lut++;                                              \
            const float val = exp(P.extinct.x * (alpha) * P.steps.y / (1.0 + s * 0.4));                          \
            opacity = 1.0 - (1.0 - opacity) * val;                                                             \
            s += P.steps.y; }
        int done = 1;
        GX_SHADOW_LAYER(a0);
        if (opacity < 1 && k1) { GX_SHADOW_LAYER(a1); done = 2;
            if (opacity < 1 && k2) { GX_SHADOW_LAYER(a2); done = 3;
                if (opacity < 1 && k3) { GX_SHADOW_LAYER(a3); done = 4; } } }
        #undef GX_SHADOW_LAYER
        if (done < 4) break;            // the loop condition failed inside this round
        GX_STEP_ADD(p, p3);
    }
}

// ------------------------------------------------------------------------------------------------ composed deep + shadow
// Primary emission/absorption march (gvdbRayDeep up to the composite), then — if any brick was entered — one shadow
// march from the first sample position towards the light whose accumulated opacity darkens the colour; composite over
// the background as gvdbRayDeep does.  The reference defines rayShadowBrick but never calls it; the composition is the one
// SURVEY.md §8c (ii) specifies, and the parity tests build the same thing from the reference's own device functions.
template <int BATCH, class S>
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
    const float4 clr = gx_lut(P.transfer, gx_transfer_index(smp.tri(p.x, p.y, p.z), P.thresh.x, gx_rcp_approx(P.thresh.z - P.thresh.y)));
    bgclr = make_float3(bgclr.x + clr.w * (clr.x - bgclr.x), bgclr.y + clr.w * (clr.y - bgclr.y), bgclr.z + clr.w * (clr.z - bgclr.z));
    return make_float4(bgclr.x, bgclr.y, bgclr.z, 1);
}

// gvdbSection3D: transfer colour on the section plane blended over the trilinear surface found behind it.
//                                                                         cuda_gvdb_module.cu:225-269
template <int BATCH, class S>
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
            clr = gx_lut(P.transfer, gx_transfer_index(t, P.thresh.x, gx_rcp_approx(P.thresh.z - P.thresh.y)));
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
