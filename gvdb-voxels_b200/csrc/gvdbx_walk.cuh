// gvdbx_walk.cuh — the hierarchical DDA of rayCast (cuda_gvdb_raycast.cuh:543-611, cuda_gvdb_dda.cuh:38-91) as a lean,
// resumable walker: "give me the next brick this ray enters".  Included by gvdbx_device.cuh (needs GxDDA, the table
// accessors and GxParams).
//
// Same per-ray arithmetic as the reference loop — every t, tSide and cell index is produced by the same sequence of
// floating-point operations — with the bookkeeping around it reduced to what one iteration needs (ncu source view of the
// literal form, SHADE_VOXEL 4K, profiles/r02_summary.md: ~70 issued instructions per iteration on the no-level-change path, 13 of them
// recomputing a shared-memory address, 12 stepping the cell index, 6 re-testing loop guards):
//
//   * Next() and Step() are FUSED.  The reference steps the DDA of the current level only if the cell is empty or after the
//     brick function returns; when it descends instead, the level's (tSide, p) are dead — the Prepare after the ascent
//     recomputes them from t.x.  So the step can be applied in every iteration, right behind Next(), while the three
//     axis predicates are still live: no mask has to be kept across the child lookup.
//   * the loop guard `0 <= p <= res` (inclusive, :567) shares one test with "p addresses a cell" (`p < res`, all axes):
//     only when that test fails are the exact guard comparisons made.  `lev > 0 && lev <= top_lev` holds by construction
//     except right after the root was popped, where the walker stops.
//   * the step signs (isign3(dir)) are three integer registers, so a cell step is three predicated adds;
//   * the (node, tMax) stack of a thread is ONE row of GX_WALK_WORDS words in shared memory, odd stride = conflict-free,
//     addressed as row + level with immediate offsets: a push or pop is one address instruction and two accesses;
//   * level changes of one iteration (descent or pops) end in ONE Prepare at one code site (as before);
//   * resumable: walk() returns after the iteration that found a brick is complete; the brick-queue ray casts call resume()
//     before walking on, which recomputes tDel and the step signs so that they are not live across the sample loop.
//
// The depth-buffer clip (`t.x > tDepth`, :571) is not part of the walker: rays of a frame with a depth buffer bound take the
// literal loop of gx_raycast.
#pragma once

// words per thread of dynamic shared memory: 4 node ids (levels 1..4) + 4 exit parameters + padding to an odd stride;
// queue kernels append 2 * GX_QK words (leaf, entry parameter per queued brick) and pad again
#define GX_WALK_WORDS 9
#ifndef GX_WALK_RESUME
#define GX_WALK_RESUME 1      // tDel and the step signs are recomputed when the walk resumes behind a sample loop (0: kept live, A/B)
#endif

template <class S, int ROW_WORDS>
struct GxWalk {
    GxDDA     d;            // d.t.x = entry parameter of the current cell, d.t.y = its exit parameter (after advance())
    int       sx, sy, sz;   // isign3(dir): cell step per axis
    float     cur_tmax;     // exit parameter of the current node
    gx_ctab_t ctab;         // child table of the current node
    unsigned  res_;         // cells per axis of the current node (compile-time 8 when S::UNI)
    int       lev, iter;
    int       node;         // node to prepare (valid when a level change is pending)
    int*      row;          // this thread's stack row
    // result of advance() when it returns BRICK
    int       leaf;
    float     t_enter, t_exit;

    enum { CONT = 0, DESCENDED = 1, BRICK = 2, END = 3 };

    static __device__ __forceinline__ int* smem_row()
    {
        extern __shared__ int gx_stack_smem[];
        return gx_stack_smem + (threadIdx.y * blockDim.x + threadIdx.x) * ROW_WORDS;
    }
    __device__ __forceinline__ unsigned res() const { return S::UNI ? 8u : res_; }
    __device__ __forceinline__ void push(int l, int n, float m) const { int* p = row + l; p[-1] = n; p[3] = __float_as_int(m); }

    __device__ __forceinline__ void prepare(const GxParams& P)
    {
        ctab = gx_table(P, lev, node, gx_dim<S>(P, lev));
        res_ = unsigned(gx_res<S>(P, lev));
        const gx_npos_t np = gx_node_pos(P, lev, node);
        d.prepare_abs(make_float3(float(np.x), float(np.y), float(np.z)), gx_vdel<S>(P, lev));
    }

    // entry of rayCast (:551-565).  false = nothing to walk (miss, or a single-brick volume: the reference loop never runs)
    __device__ __forceinline__ bool start(const GxParams& P, float3 pos, float3 dir, GxCount& cnt)
    {
        lev = P.top_lev;
        cnt.rays++;
        float3 tStart = gx_ray_box(pos, dir, P.bmin, P.bmax);
        if (tStart.z == GX_NOHIT) return false;
        if (lev < 1 || lev >= GX_MAXLEV) return false;
        row = smem_row();
        cnt.n_desc++;
        tStart.x += P.epsilon;
        cur_tmax = tStart.y - P.epsilon;
        push(lev, 0, cur_tmax);
        d.set_ray(pos, dir, tStart);
        d.inv = gx_fabs(d.inv);             // only |1 / dir| is needed from here on (prepare_abs)
        // isign3(dir) (cuda_math.cuh:1541-1545: +1 for dir > 0, else -1) from the sign bit.  The two differ only for a zero
        // (or flushed subnormal) component, and that axis never steps: its tDel is infinite, its tSide +inf or NaN, and both
        // comparisons of its mask are false
        // (volatile: keeps the front end from rematerialising the three integers from dir at every step; ptxas still may under pressure)
        asm volatile("shr.s32 %0, %1, 31;\n\tor.b32 %0, %0, 1;" : "=r"(sx) : "r"(__float_as_int(dir.x)));
        asm volatile("shr.s32 %0, %1, 31;\n\tor.b32 %0, %0, 1;" : "=r"(sy) : "r"(__float_as_int(dir.y)));
        asm volatile("shr.s32 %0, %1, 31;\n\tor.b32 %0, %0, 1;" : "=r"(sz) : "r"(__float_as_int(dir.z)));
        node = 0; iter = 0;
        prepare(P);
        return true;
    }

    // first half of one iteration of the reference loop (:567-598): guard, Next, child lookup, Step / descent bookkeeping.
    __device__ __forceinline__ int advance(const GxParams& P, GxCount& cnt)
    {
        if (!(iter < GX_MAX_ITER)) return END;          // `lev <= top_lev` of the guard: settle() spends the budget when the root is popped
        const unsigned R = res();
        // one unsigned maximum serves both tests: every axis < res = the cell exists; any axis > res = outside the
        // inclusive guard (:567); a coordinate == res passes the guard and holds no child
        const unsigned m = max(max(unsigned(d.p.x), unsigned(d.p.y)), unsigned(d.p.z));
        if (m > R) return END;
        int c = -1;
        if (m < R) {
            const int dm = gx_dim<S>(P, lev);
            c = gx_child(ctab, (((d.p.z << dm) + d.p.y) << dm) + d.p.x);
        }
        cnt.n_dda++;
        iter++;
        // Next (cuda_gvdb_dda.cuh:78-83) + Step (:86-90) in one predicated block: mask = (x < y & x <= z, y < z & y <= x,
        // z < x & z <= y); t.y = the selected side; tSide += float(mask) * tDel — a 0 mask still multiplies, 0 * inf = NaN on
        // an axis-parallel ray exactly like the reference; p += mask * pStep
        d.next_step(sx, sy, sz);
        if (c == -1) { d.t.x = d.t.y; return CONT; }
        if (lev == 1) {
            leaf = c; t_enter = d.t.x + P.epsilon; t_exit = d.t.y;
            d.t.x = d.t.y;
            return BRICK;
        }
        lev--;
        cnt.n_desc++;
        d.t.x += P.epsilon;
        cur_tmax = d.t.y - P.epsilon;
        push(lev, c, cur_tmax);
        node = c;
        return DESCENDED;
    }

    // second half (:603-609): pop the levels whose exit has been passed; one Prepare if the level changed
    __device__ __forceinline__ void settle(const GxParams& P, GxCount& cnt, bool changed)
    {
        if (changed || d.t.x > cur_tmax) {
            while (d.t.x > cur_tmax && lev <= P.top_lev) {
                lev++;
                if (lev <= P.top_lev) {
                    const int* p = row + lev;
                    node = p[-1];
                    cur_tmax = __int_as_float(p[3]);
                    cnt.n_desc++;
                    changed = true;
                }
            }
            if (lev > P.top_lev) iter = GX_MAX_ITER;          // the root was popped: the next guard ends the walk
            else if (changed) prepare(P);
        }
    }

    // Before walk() is called again after a long pause (the sample loop of the brick-queue ray casts): the per-level step tDel and
    // the step signs are recomputed from (level, |1/dir|, dir) — the same expressions, the same bits — so that these six
    // registers are not live while the bricks are marched.
    __device__ __forceinline__ void resume(const GxParams& P)
    {
#if GX_WALK_RESUME
        if (lev <= P.top_lev) d.tDel = gx_vdel<S>(P, lev) * d.inv;
        asm volatile("shr.s32 %0, %1, 31;\n\tor.b32 %0, %0, 1;" : "=r"(sx) : "r"(__float_as_int(d.dir.x)));
        asm volatile("shr.s32 %0, %1, 31;\n\tor.b32 %0, %0, 1;" : "=r"(sy) : "r"(__float_as_int(d.dir.y)));
        asm volatile("shr.s32 %0, %1, 31;\n\tor.b32 %0, %0, 1;" : "=r"(sz) : "r"(__float_as_int(d.dir.z)));
#endif
    }

    // The reference loop with the brick visit as a functor, called INSIDE the iteration that found the brick (the reference's
    // nesting): on_brick(leaf, t_enter, t_exit) returns true to stop.  The iteration is completed (settle) before returning,
    // so the walk can be resumed by calling walk() again.  Returns false when the traversal is over.
    template <class F>
    __device__ __forceinline__ bool walk(const GxParams& P, GxCount& cnt, F&& on_brick)
    {
        for (;;) {
            const int s = advance(P, cnt);
            if (s == END) return false;
            bool stop = false;
            if (s == BRICK) stop = on_brick(leaf, t_enter, t_exit);
            settle(P, cnt, s == DESCENDED);
            if (stop) return true;
        }
    }
};
