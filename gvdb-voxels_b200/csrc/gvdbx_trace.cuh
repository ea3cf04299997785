// gvdbx_trace.cuh — warp-coherent ("while-while") form of the GVDB ray cast for sm_100a.
//
// Same per-ray arithmetic as gvdbx_device.cuh's literal form (and therefore as the reference: every lane executes the
// identical sequence of floating-point operations on its own ray), but the control flow is organised for a warp:
//
//   * PHASE A / PHASE B: a lane advances its hierarchical DDA until it stands in front of a brick (or dies); only
//     then does the warp reconverge and all lanes sample their bricks together.  In the literal nesting
//     (rayCast -> brickFunc inside the loop body, cuda_gvdb_raycast.cuh:567-610) brick visits of different lanes fall
//     into different outer iterations and the sample loop runs at ~12/32 active lanes (profiles/r01_*).
//   * ONE traversal instance serves primary AND shadow rays: a lane whose primary ray is finished turns it into its
//     shadow ray (performPhongShading, cuda_gvdb_module.cu:38-57) and keeps going instead of idling until the whole
//     warp has finished the primary pass.
//   * Inside a brick the fixed-step marchers take four samples per round with the four texture fetches in flight
//     together; every sample's in-brick test is exact (three integer instructions on the float bit patterns).  Sample
//     positions are produced by the same chain of roundings (p = fma(step, dir, p) resp. p = p + wpt), so every fetched
//     value, threshold test and accumulated colour is bit-identical to the one-at-a-time loop.
#pragma once
#include "gvdbx_device.cuh"

// ------------------------------------------------------------------------------------------------ brick samplers
// Fixed-step marchers, four samples per round.  Every sample of a round is checked exactly as the reference loop
// condition does (GX_INB / GX_INB_LE below).  The four fetches of a round are issued before the first result is consumed;
// results are then consumed strictly in order, so the sample at which the loop ends, the hit position and the
// accumulated colour are those of the one-at-a-time loop.  Fetches behind the end of the loop are discarded (a fetch just outside
// the brick reads apron / neighbour texels, never unmapped memory).
// Exact in-brick tests in three integer instructions.  Non-negative floats order like their bit patterns, and every
// negative number or NaN has a larger pattern than any non-negative finite one, so for hi > 0
//     0 <= x && x < hi   <=>   bits(x) < bits(hi)        (likewise <=),
// and the three axes fold into one unsigned maximum.  The only float the pattern test gets wrong is -0.0 (>= 0 is true,
// its pattern is huge): the marchers therefore replace a -0.0 in the brick-entry point by +0.0 (gx_poszero) — no later
// point of the march can be -0.0 then (x + y is -0.0 only if both are), and +-0.0 are interchangeable in every
// comparison, sum and texture coordinate the march evaluates, so the samples taken are the reference's.
__device__ __forceinline__ float gx_poszero1(float x) { const unsigned u = __float_as_uint(x); return __uint_as_float(u == 0x80000000u ? 0u : u); }
__device__ __forceinline__ float3 gx_poszero(float3 p) { return make_float3(gx_poszero1(p.x), gx_poszero1(p.y), gx_poszero1(p.z)); }
__device__ __forceinline__ unsigned gx_maxbits(float3 q) { return max(max(__float_as_uint(q.x), __float_as_uint(q.y)), __float_as_uint(q.z)); }
#define GX_INB(q, hi)    (gx_maxbits(q) <  __float_as_uint(hi))
#define GX_INB_LE(q, hi) (gx_maxbits(q) <= __float_as_uint(hi))
// filtered fetch of a sample that is only consumed if it lies inside the brick (`k`): predicated off otherwise
#ifndef GX_PRED_FETCH
#define GX_PRED_FETCH 1
#endif
#if GX_PRED_FETCH
#define GX_TRI_IF(k, q) ((k) ? smp.tri((q).x + o.x, (q).y + o.y, (q).z + o.z) : 0.f)
#else
#define GX_TRI_IF(k, q) (smp.tri((q).x + o.x, (q).y + o.y, (q).z + o.z))
#endif
#define GX_STEP_FMA(dst, src) { (dst).x = __fmaf_rn(st, dir.x, (src).x); (dst).y = __fmaf_rn(st, dir.y, (src).y); (dst).z = __fmaf_rn(st, dir.z, (src).z); }
#define GX_STEP_ADD(dst, src) { (dst).x = __fadd_rn((src).x, wpt.x); (dst).y = __fadd_rn((src).y, wpt.y); (dst).z = __fadd_rn((src).z, wpt.z); }

// SHADE_TRILINEAR                                                        cuda_gvdb_raycast.cuh:281-300
template <class S>
__device__ __forceinline__ void gx2_brick_trilinear(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                    GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    const float st = P.steps.x, thr = P.thresh.x;
    if (P.range != nullptr && !(__ldg(&P.range[nodeid].hi) >= thr)) return;                // no sample can reach THRESH
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    const float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    t.x = st * ceilf(t.x / st);
    float3 p = gx_poszero(pos + t.x * dir - vmin);
    for (int iter = 0; iter < GX_MAX_ITER; iter += 4) {
        float3 p1, p2, p3;
        GX_STEP_FMA(p1, p); GX_STEP_FMA(p2, p1); GX_STEP_FMA(p3, p2);
        const bool k0 = GX_INB(p, res0), k1 = GX_INB(p1, res0), k2 = GX_INB(p2, res0), k3 = GX_INB(p3, res0);
        const float v0 = GX_TRI_IF(k0, p);
        const float v1 = GX_TRI_IF(k1, p1);
        const float v2 = GX_TRI_IF(k2, p2);
        const float v3 = GX_TRI_IF(k3, p3);
        int k = -1;                      // index of the sample that ends the loop, hit = it passed the threshold
        bool hit = false;
        if (!k0) k = 0; else if (v0 >= thr) { k = 0; hit = true; }
        else if (!k1) k = 1; else if (v1 >= thr) { k = 1; hit = true; p = p1; }
        else if (!k2) k = 2; else if (v2 >= thr) { k = 2; hit = true; p = p2; }
        else if (!k3) k = 3; else if (v3 >= thr) { k = 3; hit = true; p = p3; }
        if (k >= 0) {
            cnt.s_tri += k + (hit ? 1 : 0);
            if (hit) {
                h.hit = p + vmin;
                h.norm = gx_gradient(smp, p + o, cnt, false);
                h.t = t.x; h.leaf = nodeid; h.vox = gx_i3(gx_floor(h.hit));
                h.cpos = p + o;
            }
            return;
        }
        cnt.s_tri += 4;
        GX_STEP_FMA(p, p3);
    }
}

// SHADE_LEVELSET                                                         cuda_gvdb_raycast.cuh:389-410, :186-197
// (p from the UNSNAPPED t.x; inclusive bounds; the fine march re-tests the same point and returns at once)
template <class S>
__device__ __forceinline__ void gx2_brick_levelset(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                                   GxHit& h, GxCount& cnt)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    const float st = P.steps.x, thr = P.thresh.x;
    if (P.range != nullptr && !(__ldg(&P.range[nodeid].lo) < thr)) return;                 // no sample can fall below THRESH
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    const float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    const float res0 = float(gx_res<S>(P, 0));
    float3 p = gx_poszero(pos + t.x * dir - vmin);
    for (int iter = 0; iter < GX_MAX_ITER; iter += 4) {
        float3 p1, p2, p3;
        GX_STEP_FMA(p1, p); GX_STEP_FMA(p2, p1); GX_STEP_FMA(p3, p2);
        const bool k0 = GX_INB_LE(p, res0), k1 = GX_INB_LE(p1, res0), k2 = GX_INB_LE(p2, res0), k3 = GX_INB_LE(p3, res0);
        const float v0 = GX_TRI_IF(k0, p);
        const float v1 = GX_TRI_IF(k1, p1);
        const float v2 = GX_TRI_IF(k2, p2);
        const float v3 = GX_TRI_IF(k3, p3);
        int k = -1;
        bool hit = false;
        if (!k0) k = 0; else if (v0 < thr) { k = 0; hit = true; }
        else if (!k1) k = 1; else if (v1 < thr) { k = 1; hit = true; p = p1; }
        else if (!k2) k = 2; else if (v2 < thr) { k = 2; hit = true; p = p2; }
        else if (!k3) k = 3; else if (v3 < thr) { k = 3; hit = true; p = p3; }
        if (k >= 0) {
            cnt.s_tri += k + (hit ? 2 : 0);
            if (hit) {
                h.hit = p + vmin;       // always != NOHIT for finite coordinates: the reference accepts it and returns
                h.norm = gx_gradient(smp, p + o, cnt, true);
                h.t = t.x; h.leaf = nodeid; h.vox = gx_i3(gx_floor(h.hit));
                h.cpos = p + o;
            }
            return;
        }
        cnt.s_tri += 4;
        GX_STEP_FMA(p, p3);
    }
}

// one emission/absorption update for a sample that passed the MINVAL test     cuda_gvdb_raycast.cuh:514-523
// (the reference multiplies by hclr = 1 last; x * 1 is exact, so only the five separately rounded products remain)
__device__ __forceinline__ void gx_deep_accumulate(const GxParams& P, float4& clr, float4 val)
{
    val.w = exp(P.extinct.x * val.w * P.steps.x);
    const float om = 1 - val.w;
    clr.x = __fadd_rn(clr.x, __fmul_rn(__fmul_rn(__fmul_rn(val.x, clr.w), om), P.extinct.y));
    clr.y = __fadd_rn(clr.y, __fmul_rn(__fmul_rn(__fmul_rn(val.y, clr.w), om), P.extinct.y));
    clr.z = __fadd_rn(clr.z, __fmul_rn(__fmul_rn(__fmul_rn(val.z, clr.w), om), P.extinct.y));
    clr.w *= val.w;
}
#ifdef GX_REF_LAYOUT
#define GX_DEEP_LUT false
#else
#define GX_DEEP_LUT true
#endif
// The same update with the per-sample transparency exp(EXTINCT * alpha * DIRECTSTEP) already in val.w: it depends only on
// the table entry and two frame constants, so gx_build_deep_lut evaluates it once per entry and frame with the very
// expression above (same instruction sequence, same bits) instead of once per sample.
__device__ __forceinline__ void gx_deep_accumulate_pre(const GxParams& P, float4& clr, float4 val)
{
    const float om = 1 - val.w;
    clr.x = __fadd_rn(clr.x, __fmul_rn(__fmul_rn(__fmul_rn(val.x, clr.w), om), P.extinct.y));
    clr.y = __fadd_rn(clr.y, __fmul_rn(__fmul_rn(__fmul_rn(val.y, clr.w), om), P.extinct.y));
    clr.z = __fadd_rn(clr.z, __fmul_rn(__fmul_rn(__fmul_rn(val.z, clr.w), om), P.extinct.y));
    clr.w *= val.w;
}
// transfer-function index (cuda_gvdb_dda.cuh:20-23): int(min(1.0, max(0.0, u)) * 16300.0f) with u = (v - THRESH) /
// (VMAX - VMIN) a float (divide = multiply by the approximate reciprocal) and the clamp / scale / truncation in DOUBLE,
// i.e. floor of the EXACT product x = clamp(u) * 16300 (24 x 14 bits fit a double).  Same integer without the FP64 pipe and
// with a single conversion: round the product TOWARDS ZERO in fp32 and truncate.  rz(x) is the largest float <= x, and
// floor(x) < 2^24 is itself a float <= x, so floor(x) <= rz(x) <= x and floor(rz(x)) = floor(x).
__device__ __forceinline__ unsigned gx_transfer_index(float v, float thresh, float inv_range)
{
    const float u = fminf(fmaxf((v - thresh) * inv_range, 0.0f), 1.0f);      // NaN -> 0 like max(0.0, NaN)
    return unsigned(__float2int_rz(__fmul_rz(u, 16300.0f)));
}
// table entry by 32-bit byte offset from the (uniform) table base: one IMAD.WIDE.U32 instead of a sign-extended 64-bit index
__device__ __forceinline__ float4 gx_lut(const float4* table, unsigned idx)
{
    return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(table) + (size_t(idx) << 4)));
}

// SHADE_VOLUME                                                           cuda_gvdb_raycast.cuh:485-533
template <class S>
__device__ __forceinline__ void gx2_brick_deep(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir,
                                               GxHit& h, GxCount& cnt, float tDepth)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    const float st = P.steps.x;
    t.x = st * ceilf(t.x / st);
    if (h.hit.x == 0) h.hit.x = t.x;    // parameter of the first sample of the first brick entered (deep + shadow starts there)
    // every sample below MINVAL is skipped by the reference: such a brick only refreshes hit.y (not part of the image)
    // and re-applies an idempotent clamp.  Not taken when a depth buffer is bound (hit.z bookkeeping).
    if (P.range != nullptr && P.dbuf == nullptr && !(__ldg(&P.range[nodeid].hi) >= P.cutoff.x)) return;
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    const float3 o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    float3 wp = pos + t.x * dir;
    float3 p = gx_poszero(wp - vmin);
    const float3 wpt = make_float3(__fmul_rn(st, dir.x), __fmul_rn(st, dir.y), __fmul_rn(st, dir.z));
    const float dt = sqrtf(gx_dot(wpt, wpt));
    const float res0 = float(gx_res<S>(P, 0));
    const float minval = P.cutoff.x, acut = P.cutoff.y, thresh = P.thresh.x;
    const float inv_range = gx_rcp_approx(P.thresh.z - P.thresh.y);
    float4& clr = h.clr;

    if (P.dbuf != nullptr) {            // depth-buffer compositing: one sample at a time, as written in the reference
        for (int iter = 0; clr.w > acut && iter < GX_MAX_ITER && GX_INB(p, res0); iter++) {
            if (t.x > tDepth) {
                float3 d = wp - pos;
                h.hit.y = sqrtf(gx_dot(d, d));
                h.hit.z = 1;
                clr = make_float4(fminf(clr.x, 1.f), fminf(clr.y, 1.f), fminf(clr.z, 1.f), fmaxf(clr.w, 0.f));
                return;
            }
            cnt.s_tri++;
            const float raw = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);
            if (raw >= minval) { cnt.s_lut++; gx_deep_accumulate(P, clr, gx_lut(P.transfer, gx_transfer_index(raw, thresh, inv_range))); }
            GX_STEP_ADD(p, p); GX_STEP_ADD(wp, wp);
            t.x += dt;
        }
    } else {
        for (int iter = 0; iter < GX_MAX_ITER && clr.w > acut; iter += 4) {
            float3 p1, p2, p3;
            GX_STEP_ADD(p1, p); GX_STEP_ADD(p2, p1); GX_STEP_ADD(p3, p2);
            const bool k0 = GX_INB(p, res0), k1 = GX_INB(p1, res0), k2 = GX_INB(p2, res0), k3 = GX_INB(p3, res0);
            const float v0 = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);
            const float v1 = smp.tri(p1.x + o.x, p1.y + o.y, p1.z + o.z);
            const float v2 = smp.tri(p2.x + o.x, p2.y + o.y, p2.z + o.z);
            const float v3 = smp.tri(p3.x + o.x, p3.y + o.y, p3.z + o.z);
            const bool a0 = v0 >= minval, a1 = v1 >= minval, a2 = v2 >= minval, a3 = v3 >= minval;
            // transfer-function reads of the whole round in flight together (the index is valid for every sample; rejected ones are unused)
            // {rgb, exp(EXTINCT * alpha * DIRECTSTEP)} built for this frame by the host side; the module-level drop-in has no
            // host side: plain table, exp per sample
            constexpr bool pre = GX_DEEP_LUT;
            const float4* lut = pre ? P.transfer_deep : P.transfer;
            const float4 c0 = gx_lut(lut, gx_transfer_index(v0, thresh, inv_range));
            const float4 c1 = gx_lut(lut, gx_transfer_index(v1, thresh, inv_range));
            const float4 c2 = gx_lut(lut, gx_transfer_index(v2, thresh, inv_range));
            const float4 c3 = gx_lut(lut, gx_transfer_index(v3, thresh, inv_range));
            // consume in order; `done` = samples processed (each is followed by one position / t step in the reference)
            int done = 0;
            bool more = k0;             // loop condition for sample 0 (alpha was checked by the for statement)
            if (more) { done = 1; cnt.s_tri++; if (a0) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c0); else gx_deep_accumulate(P, clr, c0); } more = k1 && clr.w > acut; }
            if (more) { done = 2; cnt.s_tri++; if (a1) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c1); else gx_deep_accumulate(P, clr, c1); } more = k2 && clr.w > acut; }
            if (more) { done = 3; cnt.s_tri++; if (a2) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c2); else gx_deep_accumulate(P, clr, c2); } more = k3 && clr.w > acut; }
            if (more) { done = 4; cnt.s_tri++; if (a3) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c3); else gx_deep_accumulate(P, clr, c3); } }
            for (int q = 0; q < done; q++) t.x += dt;
            if (done < 4) break;        // left the brick or fell below ALPHACUT inside this round
            GX_STEP_ADD(p, p3);
        }
    }
    h.hit.y = t.x;
    clr = make_float4(fminf(clr.x, 1.f), fminf(clr.y, 1.f), fminf(clr.z, 1.f), fmaxf(clr.w, 0.f));
}

// ------------------------------------------------------------------------------------------------ deep mode: brick queue
// In the literal nesting (brick march inside the hierarchical-DDA loop) the lanes of a warp enter their bricks in different
// DDA iterations and with different chord lengths, so the sample rounds — 85 % of the instructions of a deep frame — run at
// 14-15 of 32 lanes (profiles/r02b_cfg4_deep*.raw.csv).  Here a ray alternates between two phases:
//   A  walk the hierarchical DDA exactly as gx_raycast does, but instead of marching a brick, append {leaf, entry parameter}
//      to a small per-thread queue in shared memory (bricks whose value range cannot contribute are dropped right here);
//      until GX_QK bricks are queued or the traversal ends;
//   B  march the queued bricks in order in ONE flat loop of four-sample rounds: a lane that leaves a brick starts the next
//      queued one inside the same loop (a ~30-instruction brick entry), so all lanes that still have work sample together
//      and chord-length differences average out over GX_QK bricks instead of costing idle lanes per brick.
// Every lane still executes, on its own ray, exactly the reference's sequence of floating-point operations: the DDA never
// depends on sample results (only the decision to stop does), each brick is entered with the same parameter, samples are
// consumed in order and the colour clamp is applied at every brick end like rayDeepBrick does.  What changes is only how
// many DDA iterations run before the ray is found to be opaque (at most GX_QK - 1 bricks of look-ahead), never the image.
#ifndef GX_QK
#define GX_QK 2              // measured on cfg4 deep 4K: 1 -> 20.9 ms, 2 -> 20.2, 3 / 4 -> 20.6, 8 -> 21.1, 16 -> 21.7 (look-ahead waste and shared
                             // memory taken from L1 grow with the depth; most of the gain is the flat loop itself)
#endif
// GX_PRED_FETCH = 1 (default): fetches / table reads of samples behind the brick's end are predicated off (0: issued and discarded,
// the A/B baseline; measured cfg4 deep 4K 19.67 -> 18.91 ms, deep + shadow 20.62 -> 20.16 ms)
#ifndef GX_Q_PREFETCH
#define GX_Q_PREFETCH 0      // measured: prefetching the next round costs registers (spills at 64, 22.5 vs 21.1 ms at 80): off
#endif
#define GX_QUEUE_BYTES_PER_THREAD (8 * GX_QK)
#define GX_WALK_WORDS_Q (GX_WALK_WORDS + 2 * GX_QK)        // stack + queue, odd stride (9 + even)

template <class S>
__device__ __forceinline__ void gx_raycast_deep_q(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt)
{
    typedef GxWalk<S, GX_WALK_WORDS_Q> Walk;
    Walk w;
    if (!w.start(P, pos, dir, cnt)) return;
    // per-thread queue behind the traversal stack in the thread's shared-memory row
    int*   q_leaf = w.row + 8;
    float* q_tx = reinterpret_cast<float*>(w.row + 8 + GX_QK);
    constexpr int nt = 1;                // queue entries of a thread are adjacent words

    const float stp = P.steps.x;
    const float3 wpt = make_float3(__fmul_rn(stp, dir.x), __fmul_rn(stp, dir.y), __fmul_rn(stp, dir.z));
    const float res0 = float(gx_res<S>(P, 0));
    const float minval = P.cutoff.x, acut = P.cutoff.y, thresh = P.thresh.x;
    const float inv_range = gx_rcp_approx(P.thresh.z - P.thresh.y);
    const float4* lut = GX_DEEP_LUT ? P.transfer_deep : P.transfer;
    float4& clr = h.clr;
    bool walking = true;                 // the hierarchical DDA has not left the volume / spent its iteration budget yet

    while (walking) {
        // ---- phase A: queue the next bricks (the walker = cuda_gvdb_raycast.cuh:567-610 without the brick call)
        int qn = 0;
        w.resume(P);
        walking = w.walk(P, cnt, [&](int leaf, float t_enter, float) {
            // brick entry of rayDeepBrick that depends on the entry parameter only (:490, first-sample bookkeeping)
            const float ts = stp * ceilf(t_enter / stp);
            if (h.hit.x == 0) h.hit.x = ts;
            cnt.n_desc++;
            // a brick without a sample >= MINVAL only re-applies an idempotent clamp: not queued
            if (P.range == nullptr || __ldg(&P.range[leaf].hi) >= minval) { q_leaf[qn * nt] = leaf; q_tx[qn * nt] = ts; qn++; }
            return qn >= GX_QK;
        });
        if (qn == 0) break;

        // ---- phase B: one flat loop of four-sample rounds over the queued bricks.  (GX_Q_PREFETCH = 1, an A/B build option:
        // the four fetches of the NEXT round are issued before this round's table reads and colour updates.)
        int    qi = 0, it = 0;
        bool   fresh = true;             // the current queue entry has not been entered yet
        float3 p = make_float3(0, 0, 0), o = make_float3(0, 0, 0);
#if GX_Q_PREFETCH
        bool   have = false;             // w0..w3 hold the values of the round that starts at p
        float  w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
#endif
        while (qi < qn) {
            if (fresh) {                 // brick entry (rayDeepBrick :487-496): sample position from the snapped parameter
                const GxLeafRec L = gx_leaf(P, q_leaf[qi * nt]);
                smp.enter(L);
                const float tx = q_tx[qi * nt];
                o = make_float3(float(L.vx), float(L.vy), float(L.vz));
                p = gx_poszero(pos + tx * dir - make_float3(float(L.px), float(L.py), float(L.pz)));
                it = 0;
                fresh = false;
#if GX_Q_PREFETCH
                have = false;
#endif
            }
            bool leave = true;           // this brick ends in this round
            if (clr.w > acut) {
                float3 p1, p2, p3, pn;
                GX_STEP_ADD(p1, p); GX_STEP_ADD(p2, p1); GX_STEP_ADD(p3, p2); GX_STEP_ADD(pn, p3);
                const bool k0 = GX_INB(p, res0), k1 = GX_INB(p1, res0), k2 = GX_INB(p2, res0), k3 = GX_INB(p3, res0);
                const bool kn = GX_INB(pn, res0) && it + 4 < GX_MAX_ITER;      // the next round has a first sample
                float v0, v1, v2, v3;
#if GX_Q_PREFETCH
                if (have) { v0 = w0; v1 = w1; v2 = w2; v3 = w3; }
                else
#endif
                {
#if GX_PRED_FETCH
                    // a sample behind the brick's end is never consumed (samples are taken in order and `more` stops at the first
                    // one outside): its fetch and its table read are predicated off instead of issued and discarded
                    v0 = k0 ? smp.tri(p.x + o.x, p.y + o.y, p.z + o.z) : 0.f;
                    v1 = k1 ? smp.tri(p1.x + o.x, p1.y + o.y, p1.z + o.z) : 0.f;
                    v2 = k2 ? smp.tri(p2.x + o.x, p2.y + o.y, p2.z + o.z) : 0.f;
                    v3 = k3 ? smp.tri(p3.x + o.x, p3.y + o.y, p3.z + o.z) : 0.f;
#else
                    v0 = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);
                    v1 = smp.tri(p1.x + o.x, p1.y + o.y, p1.z + o.z);
                    v2 = smp.tri(p2.x + o.x, p2.y + o.y, p2.z + o.z);
                    v3 = smp.tri(p3.x + o.x, p3.y + o.y, p3.z + o.z);
#endif
                }
#if GX_Q_PREFETCH
                if (kn) {
                    float3 q1, q2, q3;
                    GX_STEP_ADD(q1, pn); GX_STEP_ADD(q2, q1); GX_STEP_ADD(q3, q2);
                    w0 = smp.tri(pn.x + o.x, pn.y + o.y, pn.z + o.z);
                    w1 = smp.tri(q1.x + o.x, q1.y + o.y, q1.z + o.z);
                    w2 = smp.tri(q2.x + o.x, q2.y + o.y, q2.z + o.z);
                    w3 = smp.tri(q3.x + o.x, q3.y + o.y, q3.z + o.z);
                }
#endif
#if GX_PRED_FETCH
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 c0 = (k0 && v0 >= minval) ? gx_lut(lut, gx_transfer_index(v0, thresh, inv_range)) : z4;
                const float4 c1 = (k1 && v1 >= minval) ? gx_lut(lut, gx_transfer_index(v1, thresh, inv_range)) : z4;
                const float4 c2 = (k2 && v2 >= minval) ? gx_lut(lut, gx_transfer_index(v2, thresh, inv_range)) : z4;
                const float4 c3 = (k3 && v3 >= minval) ? gx_lut(lut, gx_transfer_index(v3, thresh, inv_range)) : z4;
#else
                const float4 c0 = gx_lut(lut, gx_transfer_index(v0, thresh, inv_range));
                const float4 c1 = gx_lut(lut, gx_transfer_index(v1, thresh, inv_range));
                const float4 c2 = gx_lut(lut, gx_transfer_index(v2, thresh, inv_range));
                const float4 c3 = gx_lut(lut, gx_transfer_index(v3, thresh, inv_range));
#endif
                int done = 0;
                bool more = k0;
                #define GX_Q_SAMPLE(v, c, knext) { done++; cnt.s_tri++; if ((v) >= minval) { cnt.s_lut++; if (GX_DEEP_LUT) gx_deep_accumulate_pre(P, clr, c); else gx_deep_accumulate(P, clr, c); } more = (knext) && clr.w > acut; }
                if (more) GX_Q_SAMPLE(v0, c0, k1)
                if (more) GX_Q_SAMPLE(v1, c1, k2)
                if (more) GX_Q_SAMPLE(v2, c2, k3)
                if (more) GX_Q_SAMPLE(v3, c3, true)
                #undef GX_Q_SAMPLE
                it += 4;
                // a next round exists only if all four samples were taken, transmittance is still above the cut and its first
                // sample lies inside the brick within the sample budget (otherwise that round would consume nothing)
                if (done == 4 && kn && clr.w > acut) {
                    p = pn; leave = false;
#if GX_Q_PREFETCH
                    have = true;
#endif
                }
            }
            if (leave) {                 // exit of rayDeepBrick (:532) and rayCast's tests behind the brick call (:584-590)
                clr = make_float4(fminf(clr.x, 1.f), fminf(clr.y, 1.f), fminf(clr.z, 1.f), fmaxf(clr.w, 0.f));
                if (clr.w <= 0) { clr.w = 0; return; }
                if (clr.w <= acut) return;          // no later brick can change the colour
                qi++;
                fresh = true;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ surface modes: brick queue
// The same two-phase scheme for SHADE_TRILINEAR and SHADE_LEVELSET (primary and shadow rays): phase A queues the next bricks
// whose value range can contain the surface, phase B marches them in one flat loop of four-sample rounds; the first sample
// that passes the threshold test ends the ray exactly where raySurfaceTrilinearBrick / rayLevelSetBrick end it
// (cuda_gvdb_raycast.cuh:281-300, :389-410): bricks are queued in ray order, each is entered with the reference's entry
// parameter, samples are tested in order.  A hit in the first queued brick wastes the DDA steps that found the others —
// bricks that can hold the surface are neighbours, so that is a step or two.
template <int MODE, class S>
__device__ __forceinline__ void gx_raycast_surface_q(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt)
{
    constexpr bool LS = (MODE == GX_MODE_LEVELSET);
    typedef GxWalk<S, GX_WALK_WORDS_Q> Walk;
    Walk w;
    if (!w.start(P, pos, dir, cnt)) return;
    int*   q_leaf = w.row + 8;
    float* q_tx = reinterpret_cast<float*>(w.row + 8 + GX_QK);
    constexpr int nt = 1;
    const float st = P.steps.x, thr = P.thresh.x;       // `st`: the step GX_STEP_FMA advances by
    const float res0 = float(gx_res<S>(P, 0));
    bool walking = true;

    while (walking) {
        // ---- phase A (see gx_raycast_deep_q)
        int qn = 0;
        w.resume(P);
        walking = w.walk(P, cnt, [&](int leaf, float t_enter, float) {
            cnt.n_desc++;
            bool keep = true;       // value-range culling: no sample of this brick can pass the threshold test
            if (P.range != nullptr) keep = LS ? (__ldg(&P.range[leaf].lo) < thr) : (__ldg(&P.range[leaf].hi) >= thr);
            if (keep) { q_leaf[qn * nt] = leaf; q_tx[qn * nt] = t_enter; qn++; }
            return qn >= GX_QK;
        });
        if (qn == 0) break;

        // ---- phase B: flat loop of four-sample rounds over the queued bricks
        int    qi = 0, it = 0;
        bool   fresh = true;
        float3 p = make_float3(0, 0, 0), o = make_float3(0, 0, 0);
        float  tx = 0.f;
        while (qi < qn) {
            if (fresh) {
                const GxLeafRec L = gx_leaf(P, q_leaf[qi * nt]);
                smp.enter(L);
                tx = q_tx[qi * nt];
                if (!LS) tx = st * ceilf(tx / st);            // trilinear: start snapped to the step grid (:286); level set: not (:394)
                o = make_float3(float(L.vx), float(L.vy), float(L.vz));
                p = gx_poszero(pos + tx * dir - make_float3(float(L.px), float(L.py), float(L.pz)));
                it = 0;
                fresh = false;
            }
            float3 p1, p2, p3;
            GX_STEP_FMA(p1, p); GX_STEP_FMA(p2, p1); GX_STEP_FMA(p3, p2);
            bool k0, k1, k2, k3;
            if (LS) { k0 = GX_INB_LE(p, res0); k1 = GX_INB_LE(p1, res0); k2 = GX_INB_LE(p2, res0); k3 = GX_INB_LE(p3, res0); }
            else    { k0 = GX_INB(p, res0); k1 = GX_INB(p1, res0); k2 = GX_INB(p2, res0); k3 = GX_INB(p3, res0); }
            const float v0 = GX_TRI_IF(k0, p);
            const float v1 = GX_TRI_IF(k1, p1);
            const float v2 = GX_TRI_IF(k2, p2);
            const float v3 = GX_TRI_IF(k3, p3);
            const bool t0 = LS ? v0 < thr : v0 >= thr, t1 = LS ? v1 < thr : v1 >= thr, t2 = LS ? v2 < thr : v2 >= thr, t3 = LS ? v3 < thr : v3 >= thr;
            int k = -1;                      // index of the sample that ends this brick; hit = it passed the threshold test
            bool hit = false;
            if (!k0) k = 0; else if (t0) { k = 0; hit = true; }
            else if (!k1) k = 1; else if (t1) { k = 1; hit = true; p = p1; }
            else if (!k2) k = 2; else if (t2) { k = 2; hit = true; p = p2; }
            else if (!k3) k = 3; else if (t3) { k = 3; hit = true; p = p3; }
            if (hit) {
                cnt.s_tri += k + (LS ? 2 : 1);
                const int leaf = q_leaf[qi * nt];
                const GxLeafRec L = gx_leaf(P, leaf);
                h.hit = p + make_float3(float(L.px), float(L.py), float(L.pz));
                h.norm = gx_gradient(smp, p + o, cnt, LS);
                h.t = tx; h.leaf = leaf; h.vox = gx_i3(gx_floor(h.hit));
                h.cpos = p + o;
                return;
            }
            if (k >= 0) { cnt.s_tri += k; qi++; fresh = true; }
            else {
                cnt.s_tri += 4;
                it += 4;
                GX_STEP_FMA(p, p3);
                if (it >= GX_MAX_ITER) { qi++; fresh = true; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ lane state machine
// The literal nesting (rayCast -> brickFunc inside the loop body) makes a warp pay, for every brick visit, as many
// sample rounds as its LONGEST chord needs: lanes with a short chord (or none) idle, and ncu shows the marchers at 10-15
// of 32 lanes.  Here a lane is either TRAVERSING (one hierarchical-DDA iteration per turn) or MARCHING (one round of four
// samples per turn); the warp loops over turns and a lane that finishes its brick goes back to traversing while its
// neighbours keep marching theirs — nobody waits for anybody.  Every lane still executes exactly the reference's
// sequence of floating-point operations on its own ray (same roundings, same order), only the interleaving across lanes
// differs, so the output is unchanged (all parity tests pass with it).
// RESULT: slower by 17-20 % (A/B, -DGX_STATE_MACHINE=1).  With 8^3 bricks and a quarter-voxel step a brick visit is only
// ~5 rounds; the traversal turn + brick entry + post-brick bookkeeping (~150 instructions) now runs in nearly every warp
// turn for the few lanes that need it instead of once per brick visit for all lanes, which costs more than the marcher
// lanes it recovers.  Kept as the measured alternative to the literal nesting, not compiled by default.
struct GxMarch {
    float3 p;           // current sample, brick-local
    float3 o;           // mValue: atlas texel of the first interior voxel
    float  tx, dt;      // ray parameter of the current sample (deep) / of the brick entry (surface), parameter step (deep)
    int    node, it;    // leaf index, samples taken so far in this brick
};

// brick entry: everything the brick functions do before their sample loop.  Returns false when there is nothing to
// march (culled brick, or deep mode already below ALPHACUT): the visit is complete.
template <int MODE, class S>
__device__ __forceinline__ bool gx3_begin(const GxParams& P, S& smp, int nodeid, float3 t, float3 pos, float3 dir, GxHit& h,
                                          GxCount& cnt, GxMarch& M)
{
    const GxLeafRec L = gx_leaf(P, nodeid);
    cnt.n_desc++;
    const float st = P.steps.x;
    if constexpr (MODE == GX_MODE_TRILINEAR) {
        if (P.range != nullptr && !(__ldg(&P.range[nodeid].hi) >= P.thresh.x)) return false;
        t.x = st * ceilf(t.x / st);
    } else if constexpr (MODE == GX_MODE_LEVELSET) {
        if (P.range != nullptr && !(__ldg(&P.range[nodeid].lo) < P.thresh.x)) return false;
    } else {
        t.x = st * ceilf(t.x / st);
        if (h.hit.x == 0) h.hit.x = t.x;
        if (P.range != nullptr && !(__ldg(&P.range[nodeid].hi) >= P.cutoff.x)) return false;
    }
    smp.enter(L);
    const float3 vmin = make_float3(float(L.px), float(L.py), float(L.pz));
    M.o = make_float3(float(L.vx), float(L.vy), float(L.vz));
    M.node = nodeid; M.it = 0; M.tx = t.x; M.dt = 0.f;
    if constexpr (MODE == GX_MODE_DEEP) {
        const float3 wp = pos + t.x * dir;
        M.p = gx_poszero(wp - vmin);
        const float3 wpt = make_float3(__fmul_rn(st, dir.x), __fmul_rn(st, dir.y), __fmul_rn(st, dir.z));
        M.dt = sqrtf(gx_dot(wpt, wpt));
        if (!(h.clr.w > P.cutoff.y)) {          // the sample loop would not run: exit bookkeeping of rayDeepBrick only
            h.hit.y = t.x;
            h.clr = make_float4(fminf(h.clr.x, 1.f), fminf(h.clr.y, 1.f), fminf(h.clr.z, 1.f), fmaxf(h.clr.w, 0.f));
            return false;
        }
    } else {
        M.p = gx_poszero(pos + t.x * dir - vmin);
    }
    return true;
}

// one round of four samples; returns true when the brick visit is complete (hit, left the brick, sample budget spent)
template <int MODE, class S>
__device__ __forceinline__ bool gx3_round(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt, GxMarch& M)
{
    const float res0 = float(gx_res<S>(P, 0));
    const float st = P.steps.x;
    const float3 o = M.o;
    float3 p = M.p, p1, p2, p3;
    if constexpr (MODE == GX_MODE_DEEP) {
        const float3 wpt = make_float3(__fmul_rn(st, dir.x), __fmul_rn(st, dir.y), __fmul_rn(st, dir.z));
        const float minval = P.cutoff.x, acut = P.cutoff.y, thresh = P.thresh.x;
        const float inv_range = gx_rcp_approx(P.thresh.z - P.thresh.y);
        float4& clr = h.clr;
        GX_STEP_ADD(p1, p); GX_STEP_ADD(p2, p1); GX_STEP_ADD(p3, p2);
        const bool k0 = GX_INB(p, res0), k1 = GX_INB(p1, res0), k2 = GX_INB(p2, res0), k3 = GX_INB(p3, res0);
        const float v0 = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);
        const float v1 = smp.tri(p1.x + o.x, p1.y + o.y, p1.z + o.z);
        const float v2 = smp.tri(p2.x + o.x, p2.y + o.y, p2.z + o.z);
        const float v3 = smp.tri(p3.x + o.x, p3.y + o.y, p3.z + o.z);
        const bool a0 = v0 >= minval, a1 = v1 >= minval, a2 = v2 >= minval, a3 = v3 >= minval;
        constexpr bool pre = GX_DEEP_LUT;
        const float4* lut = pre ? P.transfer_deep : P.transfer;
        const float4 c0 = gx_lut(lut, gx_transfer_index(v0, thresh, inv_range));
        const float4 c1 = gx_lut(lut, gx_transfer_index(v1, thresh, inv_range));
        const float4 c2 = gx_lut(lut, gx_transfer_index(v2, thresh, inv_range));
        const float4 c3 = gx_lut(lut, gx_transfer_index(v3, thresh, inv_range));
        int done = 0;
        bool more = k0;
        if (more) { done = 1; cnt.s_tri++; if (a0) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c0); else gx_deep_accumulate(P, clr, c0); } more = k1 && clr.w > acut; }
        if (more) { done = 2; cnt.s_tri++; if (a1) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c1); else gx_deep_accumulate(P, clr, c1); } more = k2 && clr.w > acut; }
        if (more) { done = 3; cnt.s_tri++; if (a2) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c2); else gx_deep_accumulate(P, clr, c2); } more = k3 && clr.w > acut; }
        if (more) { done = 4; cnt.s_tri++; if (a3) { cnt.s_lut++; if (pre) gx_deep_accumulate_pre(P, clr, c3); else gx_deep_accumulate(P, clr, c3); } }
        for (int q = 0; q < done; q++) M.tx += M.dt;
        M.it += 4;
        if (done == 4 && M.it < GX_MAX_ITER && clr.w > acut) { GX_STEP_ADD(M.p, p3); return false; }
        h.hit.y = M.tx;
        clr = make_float4(fminf(clr.x, 1.f), fminf(clr.y, 1.f), fminf(clr.z, 1.f), fmaxf(clr.w, 0.f));
        return true;
    } else {
        const float thr = P.thresh.x;
        GX_STEP_FMA(p1, p); GX_STEP_FMA(p2, p1); GX_STEP_FMA(p3, p2);
        bool k0, k1, k2, k3;
        if constexpr (MODE == GX_MODE_TRILINEAR) { k0 = GX_INB(p, res0); k1 = GX_INB(p1, res0); k2 = GX_INB(p2, res0); k3 = GX_INB(p3, res0); }
        else                                     { k0 = GX_INB_LE(p, res0); k1 = GX_INB_LE(p1, res0); k2 = GX_INB_LE(p2, res0); k3 = GX_INB_LE(p3, res0); }
        const float v0 = smp.tri(p.x + o.x, p.y + o.y, p.z + o.z);
        const float v1 = smp.tri(p1.x + o.x, p1.y + o.y, p1.z + o.z);
        const float v2 = smp.tri(p2.x + o.x, p2.y + o.y, p2.z + o.z);
        const float v3 = smp.tri(p3.x + o.x, p3.y + o.y, p3.z + o.z);
        const bool ls = (MODE == GX_MODE_LEVELSET);
        const bool t0 = ls ? v0 < thr : v0 >= thr, t1 = ls ? v1 < thr : v1 >= thr, t2 = ls ? v2 < thr : v2 >= thr, t3 = ls ? v3 < thr : v3 >= thr;
        int k = -1;
        bool hit = false;
        if (!k0) k = 0; else if (t0) { k = 0; hit = true; }
        else if (!k1) k = 1; else if (t1) { k = 1; hit = true; p = p1; }
        else if (!k2) k = 2; else if (t2) { k = 2; hit = true; p = p2; }
        else if (!k3) k = 3; else if (t3) { k = 3; hit = true; p = p3; }
        if (k >= 0) {
            cnt.s_tri += k + (hit ? (ls ? 2 : 1) : 0);
            if (hit) {
                const GxLeafRec L = gx_leaf(P, M.node);
                h.hit = p + make_float3(float(L.px), float(L.py), float(L.pz));
                h.norm = gx_gradient(smp, p + o, cnt, ls);
                h.t = M.tx; h.leaf = M.node; h.vox = gx_i3(gx_floor(h.hit));
                h.cpos = p + o;
            }
            return true;
        }
        cnt.s_tri += 4;
        M.it += 4;
        GX_STEP_FMA(M.p, p3);
        return M.it >= GX_MAX_ITER;
    }
}

template <int MODE, class S>
__device__ __forceinline__ void gx_raycast_sm(const GxParams& P, S& smp, float3 pos, float3 dir, GxHit& h, GxCount& cnt, int px, int py)
{
    GxStack st;
    int lev = P.top_lev;
    cnt.rays++;
    float3 tStart = gx_ray_box(pos, dir, P.bmin, P.bmax);
    if (tStart.z == GX_NOHIT) return;
    if (lev < 1 || lev >= GX_MAXLEV) return;
    gx_npos_t np = gx_node_pos(P, lev, 0);
    cnt.n_desc++;
    float3 vmin = make_float3(float(np.x), float(np.y), float(np.z));
    tStart.x += P.epsilon;
    st.set(lev, 0, tStart.y - P.epsilon);
    float      cur_tmax = tStart.y - P.epsilon;
    gx_ctab_t  ctab = gx_table(P, lev, 0, gx_dim<S>(P, lev));
    unsigned   res = unsigned(gx_res<S>(P, lev));
    GxDDA dda;
    dda.set_ray(pos, dir, tStart);
    dda.prepare(vmin, gx_vdel<S>(P, lev));
    const float tDepth = gx_depth_max(P, dir, px, py);

    // tail of a reference loop iteration: pop the levels whose exit has been passed (cuda_gvdb_raycast.cuh:603-609)
    auto ascend = [&]() {
        while (dda.t.x > cur_tmax && lev <= P.top_lev) {
            lev++;
            if (lev <= P.top_lev) {
                const int n = st.node(lev);
                cur_tmax = st.tmax(lev);
                ctab = gx_table(P, lev, n, gx_dim<S>(P, lev));
                res = unsigned(gx_res<S>(P, lev));
                const gx_npos_t q = gx_node_pos(P, lev, n);
                cnt.n_desc++;
                dda.prepare(make_float3(float(q.x), float(q.y), float(q.z)), gx_vdel<S>(P, lev));
            }
        }
    };
    // what rayCast does after a brick function returns (:584-602); false = the ray is finished
    auto after_brick = [&]() -> bool {
        if (h.clr.w <= 0) { h.clr.w = 0; return false; }
        if (h.hit.z != GX_NOHIT) return false;
        if (MODE == GX_MODE_DEEP && h.clr.w <= P.cutoff.y) return false;
        dda.step();
        return true;
    };

    enum { TRAVERSE = 0, MARCH = 1, DONE = 2 };
    int state = TRAVERSE, iter = 0;
    GxMarch M;
    while (state != DONE) {
        if (state == TRAVERSE) {
            if (!(iter < GX_MAX_ITER && lev > 0 && lev <= P.top_lev
                  && unsigned(dda.p.x) <= res && unsigned(dda.p.y) <= res && unsigned(dda.p.z) <= res)) { state = DONE; }
            else {
                dda.next();
                if (dda.t.x > tDepth) { h.hit.z = 0; state = DONE; }
                else {
                    const int dm = gx_dim<S>(P, lev);
                    const int b = (((int(dda.p.z) << dm) + int(dda.p.y)) << dm) + int(dda.p.x);
                    int c = -1;
                    if (unsigned(dda.p.x | dda.p.y | dda.p.z) < res) c = gx_child(ctab, b);
                    cnt.n_dda++;
                    bool tail = true;                   // finish the reference iteration now (ascend, iter++)
                    if (c != -1) {
                        if (lev == 1) {
                            dda.t.x += P.epsilon;
                            if (gx3_begin<MODE>(P, smp, c, dda.t, pos, dir, h, cnt, M)) { state = MARCH; tail = false; }
                            else if (!after_brick()) { state = DONE; tail = false; }
                        } else {
                            lev--;
                            np = gx_node_pos(P, lev, c);
                            cnt.n_desc++;
                            dda.t.x += P.epsilon;
                            cur_tmax = dda.t.y - P.epsilon;
                            st.set(lev, c, cur_tmax);
                            ctab = gx_table(P, lev, c, gx_dim<S>(P, lev));
                            res = unsigned(gx_res<S>(P, lev));
                            dda.prepare(make_float3(float(np.x), float(np.y), float(np.z)), gx_vdel<S>(P, lev));
                        }
                    } else {
                        dda.step();
                    }
                    if (tail) { ascend(); iter++; }
                }
            }
        }
        if (state == MARCH) {
            if (gx3_round<MODE>(P, smp, pos, dir, h, cnt, M)) {
                if (after_brick()) { ascend(); iter++; state = TRAVERSE; }
                else state = DONE;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ traversal state
struct GxTrav {
    GxDDA   dda;
    GxStackReg st;
    int     lev, iter;
    float   tDepth;
    bool    alive;
};

// entry of rayCast: slab test, root, first Prepare                      cuda_gvdb_raycast.cuh:551-565
template <class S>
__device__ __forceinline__ void gx2_start(const GxParams& P, GxTrav& T, float3 pos, float3 dir, GxCount& cnt, int px, int py)
{
    T.alive = false; T.iter = 0;
    T.lev = P.top_lev;
    T.st.n1 = T.st.n2 = T.st.n3 = T.st.n4 = 0; T.st.m1 = T.st.m2 = T.st.m3 = T.st.m4 = 0.f;
    cnt.rays++;
    float3 tStart = gx_ray_box(pos, dir, P.bmin, P.bmax);
    if (tStart.z == GX_NOHIT) return;
    if (T.lev < 1 || T.lev >= GX_MAXLEV) return;
    const gx_npos_t np = gx_node_pos(P, T.lev, 0);
    cnt.n_desc++;
    const float3 vmin = make_float3(float(np.x), float(np.y), float(np.z));
    tStart.x += P.epsilon;
    T.st.set(T.lev, 0, tStart.y - P.epsilon);
    T.dda.set_ray(pos, dir, tStart);
    T.dda.prepare(vmin, gx_vdel<S>(P, T.lev));
    T.tDepth = gx_depth_max(P, dir, px, py);
    T.alive = true;
}

// pop levels whose exit has been passed                                  cuda_gvdb_raycast.cuh:603-609
template <class S>
__device__ __forceinline__ void gx2_ascend(const GxParams& P, GxTrav& T, GxCount& cnt)
{
    while (T.dda.t.x > T.st.tmax(T.lev) && T.lev <= P.top_lev) {
        T.lev++;
        if (T.lev <= P.top_lev) {
            const gx_npos_t np = gx_node_pos(P, T.lev, T.st.node(T.lev));
            cnt.n_desc++;
            T.dda.prepare(make_float3(float(np.x), float(np.y), float(np.z)), gx_vdel<S>(P, T.lev));
        }
    }
}

// PHASE A, one iteration: one pass of the reference loop body (cuda_gvdb_raycast.cuh:567-602) minus the brick call.
// Returns the leaf index when the ray now stands in front of a brick (dda.t.x already moved inside by epsilon; the
// iteration is completed later by gx2_leave_brick), otherwise -1 (T.alive cleared when the ray is finished).
template <class S>
__device__ __forceinline__ int gx2_dda_iteration(const GxParams& P, GxTrav& T, GxHit& h, GxCount& cnt)
{
    GxDDA& d = T.dda;
    const int lev = T.lev;
    if (!(T.iter < GX_MAX_ITER && lev > 0 && lev <= P.top_lev && d.p.x >= 0 && d.p.y >= 0 && d.p.z >= 0
          && d.p.x <= gx_res<S>(P, lev) && d.p.y <= gx_res<S>(P, lev) && d.p.z <= gx_res<S>(P, lev))) { T.alive = false; return -1; }
    d.next();
    if (d.t.x > T.tDepth) { h.hit.z = 0; T.alive = false; return -1; }
    const int dm = gx_dim<S>(P, lev);
    const int b = (((int(d.p.z) << dm) + int(d.p.y)) << dm) + int(d.p.x);
    int c = -1;
    if (d.p.x < gx_res<S>(P, lev) && d.p.y < gx_res<S>(P, lev) && d.p.z < gx_res<S>(P, lev))
        c = gx_child(gx_table(P, lev, T.st.node(lev), dm), b);
    cnt.n_dda++;
    if (c != -1) {
        d.t.x += P.epsilon;
        if (lev == 1) return c;
        T.lev = lev - 1;
        const gx_npos_t np = gx_node_pos(P, lev - 1, c);
        cnt.n_desc++;
        T.st.set(lev - 1, c, d.t.y - P.epsilon);
        d.prepare(make_float3(float(np.x), float(np.y), float(np.z)), gx_vdel<S>(P, lev - 1));
    } else {
        d.step();
    }
    gx2_ascend<S>(P, T, cnt);
    T.iter++;
    return -1;
}

// second half of the reference iteration that visited a brick: termination tests, Step, ascend    :584-609
template <int MODE, class S>
__device__ __forceinline__ void gx2_leave_brick(const GxParams& P, GxTrav& T, GxHit& h, GxCount& cnt)
{
    if (h.clr.w <= 0) { h.clr.w = 0; T.alive = false; return; }
    if (h.hit.z != GX_NOHIT) { T.alive = false; return; }
    if (MODE == GX_MODE_DEEP && h.clr.w <= P.cutoff.y) { T.alive = false; return; }   // later bricks cannot change the colour
    T.dda.step();
    gx2_ascend<S>(P, T, cnt);
    T.iter++;
}

// ------------------------------------------------------------------------------------------------ per-pixel driver
// All 32 lanes of the warp call this together (`valid` = lane owns a pixel); convergence is forced with warp votes at
// the head of every DDA iteration and every ray-state transition, so that PHASE B really runs with all lanes that
// have a brick pending.  Returns the final float colour of the pixel (before 8-bit packing); `prim` receives the
// primary ray's hit record.
template <int MODE, class S>
__device__ __forceinline__ float4 gx2_trace_pixel(const GxParams& P, S& smp, float3 rpos, float3 rdir, int px, int py,
                                                  GxCount& cnt, GxHit& prim, float4& raw, bool valid)
{
    const unsigned FULL = 0xffffffffu;
    GxTrav T;
    GxHit h;
    h.norm = make_float3(0, 0, 0); h.t = 0; h.leaf = -1; h.vox = make_int3(0, 0, 0);
    if (MODE == GX_MODE_DEEP) { h.clr = make_float4(0, 0, 0, 1); h.hit = make_float3(0, 0, GX_NOHIT); }
    else { h.clr = make_float4(1, 1, 1, 1);
           h.hit = (MODE == GX_MODE_LEVELSET) ? make_float3(0, 0, GX_NOHIT) : make_float3(GX_NOHIT, GX_NOHIT, GX_NOHIT); }
    prim = h;
    float3 pos = rpos, dir = rdir;
    int phase = 0;                      // 0 = primary ray, 1 = shadow ray
    float diff = 0.f;
    float4 result = make_float4(0, 0, 0, 0);
    bool done = !valid;
    T.alive = false; T.iter = 0; T.lev = 0; T.tDepth = 0;
    if (valid) gx2_start<S>(P, T, pos, dir, cnt, px, py);

    while (!__all_sync(FULL, done)) {
        // ---- PHASE A: every live lane walks its DDA until it stands in front of a brick or its ray ends
        int brick = -1;
        while (__any_sync(FULL, !done && T.alive && brick < 0)) {
            if (!done && T.alive && brick < 0) brick = gx2_dda_iteration<S>(P, T, h, cnt);
        }
        // ---- PHASE B: the warp samples its pending bricks together
        if (brick >= 0) {
            if (MODE == GX_MODE_VOXEL)          gx_brick_voxel(P, smp, brick, T.dda.t, pos, dir, h, cnt);
            else if (MODE == GX_MODE_TRILINEAR) gx2_brick_trilinear(P, smp, brick, T.dda.t, pos, dir, h, cnt);
            else if (MODE == GX_MODE_LEVELSET)  gx2_brick_levelset(P, smp, brick, T.dda.t, pos, dir, h, cnt);
            else                                gx2_brick_deep(P, smp, brick, T.dda.t, pos, dir, h, cnt, T.tDepth);
            gx2_leave_brick<MODE, S>(P, T, h, cnt);
        }
        __syncwarp();
        if (done || T.alive) continue;

        // ---- the current ray of this lane is finished
        if (MODE == GX_MODE_DEEP) {
            raw = h.clr;
            prim = h;
            const float a = 1.0 - h.clr.w;
            result = make_float4(P.backclr.x + a * (h.clr.x - P.backclr.x), P.backclr.y + a * (h.clr.y - P.backclr.y),
                                 P.backclr.z + a * (h.clr.z - P.backclr.z), 1.0 - h.clr.w);
            done = true;
        } else if (phase == 0) {                                 // performPhongShading, cuda_gvdb_module.cu:38-57
            gx_hit_color(P, h);
            prim = h;
            if (h.hit.z == GX_NOHIT) { result = P.backclr; done = true; }
            else {
                const float3 lightdir = gx_normalize(P.light_pos - h.hit);
                diff = 0.9 * fmaxf(0.0f, gx_dot(h.norm, lightdir));
                if (P.shadow_params.x > 0) {
                    pos = h.hit + h.norm * P.shadow_params.y;
                    dir = lightdir;
                    result = h.clr;                              // surface colour, kept while the shadow ray runs
                    h.hit = make_float3(0, 0, GX_NOHIT);
                    h.clr = make_float4(0, 0, 0, 1);
                    phase = 1;
                    gx2_start<S>(P, T, pos, dir, cnt, px, py);
                    if (!T.alive) {                              // shadow ray misses the volume box entirely
                        result = make_float4(result.x * (diff + 0.1f), result.y * (diff + 0.1f), result.z * (diff + 0.1f), 1.0);
                        done = true;
                    }
                } else {
                    result = make_float4(h.clr.x * (diff + 0.1f), h.clr.y * (diff + 0.1f), h.clr.z * (diff + 0.1f), 1.0);
                    done = true;
                }
            }
        } else {                                                 // shadow ray finished
            diff = (h.hit.z == GX_NOHIT ? diff : diff * (1.0 - P.shadow_params.x));
            result = make_float4(result.x * (diff + 0.1f), result.y * (diff + 0.1f), result.z * (diff + 0.1f), 1.0);
            done = true;
        }
    }
    return result;
}
