// td_obs.cuh -- (f) the observation writer: element types, dense planes, sparse one-hots / enemy statistics,
// the in-place (incremental) update.  TDBoard.py:355-365, 85-144.
#pragma once
#include "td_common.cuh"

namespace td {

// ------------------------------------------------------------------------------------------------
// (f) observation: dense planes with streaming float4 stores, then the sparse one-hots / enemy
//     statistics as 4-byte stores on top (ordered after the dense pass by __syncwarp).

// ------------------------------------------------------------------------------------------------
// Observation element types (td_step_io.obs_format): float32 is the reference layout and the default; bfloat16 and
// unorm8 are the opt-in reduced-precision planes of SURVEY.md 8(f) f4 -- same (45, L, L) layout, 2 / 1 bytes per
// element.  bf16 = round-to-nearest-even of the float32 value; u8 = rint(min(v * 255, 255)) (values above 1 saturate).
// Four consecutive elements ("quad") go out in one store: 16 / 8 / 4 bytes.
template <class OT> struct ObsType;
template <> struct ObsType<float> { static constexpr int kFormat = TD_OBS_F32; };
template <> struct ObsType<__nv_bfloat16> { static constexpr int kFormat = TD_OBS_BF16; };
template <> struct ObsType<uint8_t> { static constexpr int kFormat = TD_OBS_U8; };

__device__ __forceinline__ uint32_t obs_u8(float v) { return __float2uint_rn(fminf(__fmul_rn(v, 255.f), 255.f)); }

__device__ __forceinline__ void obs_store4(float *o, size_t quad, float4 v) { TD_ST(reinterpret_cast<float4 *>(o) + quad, v); }
__device__ __forceinline__ void obs_store4(__nv_bfloat16 *o, size_t quad, float4 v)
{
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t *>(&lo);
    u.y = *reinterpret_cast<const uint32_t *>(&hi);
    reinterpret_cast<uint2 *>(o)[quad] = u;
}
__device__ __forceinline__ void obs_store4(uint8_t *o, size_t quad, float4 v)
{
    reinterpret_cast<uint32_t *>(o)[quad] = obs_u8(v.x) | (obs_u8(v.y) << 8) | (obs_u8(v.z) << 16) | (obs_u8(v.w) << 24);
}
__device__ __forceinline__ void obs_store1(float *o, size_t i, float v) { o[i] = v; }
__device__ __forceinline__ void obs_store1(__nv_bfloat16 *o, size_t i, float v) { o[i] = __float2bfloat16_rn(v); }
__device__ __forceinline__ void obs_store1(uint8_t *o, size_t i, float v) { o[i] = (uint8_t)obs_u8(v); }

__device__ __forceinline__ void fill_planes(float *o, int first_plane, int n_planes, int cells, float v, int lane, int stride)
{
    float4 *p = reinterpret_cast<float4 *>(o + (size_t)first_plane * cells);
    const int n4 = (n_planes * cells) >> 2;
    const float4 x = make_float4(v, v, v, v);
    for (int q = lane; q < n4; q += stride) TD_ST(p + q, x);
}

__device__ __forceinline__ void fill_planes_scalar(float *o, int first_plane, int n_planes, int cells, float v, int lane, int stride)
{
    float *p = o + (size_t)first_plane * cells;
    for (int q = lane; q < n_planes * cells; q += stride) TD_ST(p + q, v);
}

// N4 consecutive float4 of one value, fully unrolled: one STG.128 with an immediate offset per 512 bytes.
template <int N4, int G, class OT>
__device__ __forceinline__ void store_run(OT *o, int first_quad, float v, int lane)
{
    const float4 x = make_float4(v, v, v, v);
    constexpr int kFullIters = N4 / G, kRem = N4 % G;
    const size_t q0 = (size_t)first_quad + lane;
#pragma unroll
    for (int k = 0; k < kFullIters; ++k) obs_store4(o, q0 + G * k, x);
    if (kRem != 0 && lane < kRem) obs_store4(o, q0 + G * kFullIters, x);
}

// CELLS > 0: compile-time board size (runs of equal planes are unrolled stores with immediate offsets).
// CELLS == 0: run-time board size, plane by plane (also handles L*L not divisible by 4).
// Step 1 of the observation: the 12 broadcast plane values, one lane each, parked behind ratio[64] in scratch.
template <class W>
__device__ __forceinline__ void obs_prepare(W &w)
{
    const DevConfig &cc = w.pp->cfg;
    const int lane = w.lane;
    // The 12 broadcast values (f64 quotients rounded once to f32, TDBoard.py:115-125,134-142), one per lane:
    // lane 0 -> plane 5, 1 -> 11, 2 -> 12, 3 -> 13, 4..7 -> 41..44 (cost_def / enemy_cost / 8), 8..11 -> 21..24.
    float *pv = reinterpret_cast<float *>(w.scratch()) + 64;         // [48], behind ratio[64]
    {
        double num = 1.0, den = 1.0;      // idle lanes divide 1 by 1: a zero numerator takes the division's slow-path call
        int plane = 47;
        if (lane == 0) { num = (double)w.base_LP; den = (double)cc.base_LP; plane = 5; }
        else if (lane == 1) { num = w.cost_def; den = cc.max_cost; plane = 11; }
        else if (lane == 2) { num = w.cost_atk; den = cc.max_cost; plane = 12; }
        else if (lane == 3) { num = (double)w.steps; den = (double)cc.max_steps; plane = 13; }
        else if (lane < 8) { num = w.cost_def; den = cc.enemy_cost[lane - 4][0]; plane = 41 + lane - 4; }
        else if (lane < 12) { plane = 21 + lane - 8; }
        double qv = num / den;
        if (lane >= 4 && lane < 8) qv *= 0.125;                    // "/ max_cluster_length": exact power of two
        float val = (float)qv;
        if (lane == 0 && cc.base_LP < 0) val = 1.f;
        if (lane >= 8 && lane < 12) val = w.cost_def >= cc.tower_cost[lane - 8][0] ? 1.f : 0.f;
        gsync(w);
        for (int q = lane; q < 48; q += W::G) pv[q] = 0.f;
        gsync(w);
        if (lane < 12) pv[plane] = val;
        gsync(w);
    }
}

// Step 2: the dense planes of the env whose record sits in w.slice, written by NT cooperating threads
// (tid in [0, NT)): NT = W::G for one group per env, NT = the CTA size for the CTA-cooperative sweep.
// Only the slice pointers of `w` are used.
// dist / maxd for the distance plane, correctly rounded like the IEEE division the reference's float32 array
// performs, without the division's range check: a zero numerator (every off-road cell) sends __fdiv_rn through
// its slow-path call.  One refined reciprocal per env, then quotient + exact remainder + correction per cell
// (Markstein); tests/test_host.py proves it for all 0 <= a <= 255, 1 <= b <= 256 and any 1-ulp reciprocal.
struct SmallDiv {
    float b, r;
    __device__ __forceinline__ explicit SmallDiv(float den) : b(den)
    {
        float x;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(den));
        r = __fmaf_rn(x, __fmaf_rn(-den, x, 1.f), x);
    }
    __device__ __forceinline__ float operator()(float a) const
    {
        const float q = __fmul_rn(a, r);
        return __fmaf_rn(__fmaf_rn(-b, q, a), r, q);
    }
};

// INTERLEAVE (one group per env, NT == W::G): the sparse pieces follow the zero runs they land on; returns true
// when they were written here (the unrolled path), false when the caller still has to run obs_sparse.
template <int NT, class W, class OT, bool INTERLEAVE = false>
__device__ __forceinline__ bool obs_dense(W &w, OT *o, int tid)
{
    constexpr int CELLS = W::kCells;
    constexpr bool kF32 = ObsType<OT>::kFormat == TD_OBS_F32;
    static_assert(kF32 || CELLS > 0, "reduced-precision observations exist for the specialised board sizes");
    const int cells = CELLS > 0 ? CELLS : w.ncells();
    const bool vec = (cells & 3) == 0 && ((reinterpret_cast<uintptr_t>(o) & (4 * sizeof(OT) - 1)) == 0);
    const SmallDiv by_maxd((float)w.mh()->maxd_p1);
    const float *pv = reinterpret_cast<const float *>(w.scratch()) + 64;
    if (CELLS > 0 && (vec || !kF32)) {
        constexpr int C4 = CELLS > 0 ? CELLS / 4 : 1;
        constexpr int kIters = (C4 + NT - 1) / NT;
        const uchar4 *cb = reinterpret_cast<const uchar4 *>(w.cells());
        const uchar4 *db = reinterpret_cast<const uchar4 *>(w.dist());
        const uchar4 *mb = reinterpret_cast<const uchar4 *>(w.map6());
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int q = tid + NT * it;
            if (q < C4) {
                const uchar4 c = cb[q], d = db[q], m = mb[q];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    obs_store4(o, (size_t)(k * C4 + q), make_float4((float)((c.x >> k) & 1), (float)((c.y >> k) & 1),
                                                                     (float)((c.z >> k) & 1), (float)((c.w >> k) & 1)));
                obs_store4(o, (size_t)(9 * C4 + q), make_float4(by_maxd((float)d.x), by_maxd((float)d.y),
                                                                 by_maxd((float)d.z), by_maxd((float)d.w)));
                obs_store4(o, (size_t)(14 * C4 + q), make_float4(m.x == 0 ? 1.f : 0.f, m.y == 0 ? 1.f : 0.f,
                                                                  m.z == 0 ? 1.f : 0.f, m.w == 0 ? 1.f : 0.f));
            }
        }
        store_run<C4, NT>(o, 4 * C4, 0.f, tid);
        store_run<C4, NT>(o, 5 * C4, pv[5], tid);
        store_run<3 * C4, NT>(o, 6 * C4, 0.f, tid);
        if constexpr (INTERLEAVE) { gsync(w); obs_sparse_static(w, o); }
        store_run<C4, NT>(o, 10 * C4, 0.f, tid);
#pragma unroll
        for (int k = 11; k < 14; ++k) store_run<C4, NT>(o, k * C4, pv[k], tid);
        store_run<6 * C4, NT>(o, 15 * C4, 0.f, tid);
        if constexpr (INTERLEAVE) { gsync(w); obs_sparse_towers(w, o); }
#pragma unroll
        for (int k = 21; k < 25; ++k) store_run<C4, NT>(o, k * C4, pv[k], tid);
        store_run<16 * C4, NT>(o, 25 * C4, 0.f, tid);
        if constexpr (INTERLEAVE) { gsync(w); obs_sparse_enemies(w, o); }
#pragma unroll
        for (int k = 41; k < 45; ++k) store_run<C4, NT>(o, k * C4, pv[k], tid);
        return INTERLEAVE;
    } else if constexpr (kF32) {
      if (vec) {
        const int c4 = cells >> 2;
        float4 *o4 = reinterpret_cast<float4 *>(o);
        const uchar4 *cb = reinterpret_cast<const uchar4 *>(w.cells());
        const uchar4 *db = reinterpret_cast<const uchar4 *>(w.dist());
        const uchar4 *mb = reinterpret_cast<const uchar4 *>(w.map6());
        for (int q = tid; q < c4; q += NT) {
            uchar4 c = cb[q];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                TD_ST(o4 + k * c4 + q, make_float4((float)((c.x >> k) & 1), (float)((c.y >> k) & 1),
                                                    (float)((c.z >> k) & 1), (float)((c.w >> k) & 1)));
        }
        fill_planes(o, 4, 1, cells, 0.f, tid, NT);
        fill_planes(o, 5, 1, cells, pv[5], tid, NT);
        fill_planes(o, 6, 3, cells, 0.f, tid, NT);
        for (int q = tid; q < c4; q += NT) {
            uchar4 d = db[q];
            TD_ST(o4 + 9 * c4 + q, make_float4(by_maxd((float)d.x), by_maxd((float)d.y),
                                                by_maxd((float)d.z), by_maxd((float)d.w)));
        }
        fill_planes(o, 10, 1, cells, 0.f, tid, NT);
        for (int k = 11; k < 14; ++k) fill_planes(o, k, 1, cells, pv[k], tid, NT);
        for (int q = tid; q < c4; q += NT) {
            uchar4 m = mb[q];
            TD_ST(o4 + 14 * c4 + q, make_float4(m.x == 0 ? 1.f : 0.f, m.y == 0 ? 1.f : 0.f,
                                                 m.z == 0 ? 1.f : 0.f, m.w == 0 ? 1.f : 0.f));
        }
        fill_planes(o, 15, 6, cells, 0.f, tid, NT);
        for (int k = 21; k < 25; ++k) fill_planes(o, k, 1, cells, pv[k], tid, NT);
        fill_planes(o, 25, 16, cells, 0.f, tid, NT);
        for (int k = 41; k < 45; ++k) fill_planes(o, k, 1, cells, pv[k], tid, NT);
      } else {
        for (int q = tid; q < cells; q += NT) {
            uint8_t c = w.cells()[q];
#pragma unroll
            for (int k = 0; k < 4; ++k) TD_ST(o + (size_t)k * cells + q, (float)((c >> k) & 1));
            TD_ST(o + (size_t)9 * cells + q, by_maxd((float)w.dist()[q]));
            TD_ST(o + (size_t)14 * cells + q, w.map6()[q] == 0 ? 1.f : 0.f);
        }
        for (int k = 4; k < TD_NCHANNELS; ++k)
            if (k != 9 && k != 14) fill_planes_scalar(o, k, 1, cells, pv[k], tid, NT);
      }
    }
    return false;
}

// Step 3: the sparse one-hots and enemy statistics, 4-byte stores on top of the dense planes, in three pieces: the
// static one-hots (end, starts: planes 4, 6-8), the towers (planes 15-20), the enemy statistics (planes 25-40).
// Every piece must be ordered after the dense stores of its planes by the caller (gsync for one group per env).
// obs_sparse runs all three behind the whole dense pass; obs_dense<INTERLEAVE> runs each right behind the zero
// run it lands on -- a fix-up that follows its line by microseconds instead of a whole observation costs less
// (tools/compressbench.cu: 0.1869 -> 0.1783 ms on compressible memory, 0.2106 -> 0.2051 ms on ordinary memory).
template <class W, class OT>
__device__ __forceinline__ void obs_sparse_static(W &w, OT *o)
{
#ifdef TD_EXP_NO_SPARSE      // timing experiment only (wrong observations): what do the sparse fix-ups cost?
    return;
#endif
    const int lane = w.lane, cells = W::kCells > 0 ? W::kCells : w.ncells();
    if (lane == 0) obs_store1(o, (size_t)4 * cells + w.mh()->end, 1.f);
    if (lane < w.mh()->num_roads) obs_store1(o, (size_t)(6 + lane) * cells + w.mh()->start[lane], 1.f);
}

template <class W, class OT>
__device__ __forceinline__ void obs_sparse_towers(W &w, OT *o)
{
#ifdef TD_EXP_NO_SPARSE
    return;
#endif
    const int cells = W::kCells > 0 ? W::kCells : w.ncells();
    for (int t = w.lane; t < w.nt; t += W::G) {
        const td_tower_rec &T = w.tw()[t];
        obs_store1(o, (size_t)(15 + (T.type_lv >> 2)) * cells + T.loc, 1.f);
        obs_store1(o, (size_t)(17 + (T.type_lv & 3)) * cells + T.loc, 1.f);
    }
}

// enemy statistics per (type, cell) group in list order, float32 (TDBoard.py:355-365, NumPy-2 casts)
template <class W, class OT>
__device__ __forceinline__ void obs_sparse_enemies(W &w, OT *o)
{
#ifdef TD_EXP_NO_SPARSE
    return;
#endif
    const DevConfig &cc = w.pp->cfg;
    constexpr int CELLS = W::kCells;
    const int lane = w.lane, cells = CELLS > 0 ? CELLS : w.ncells();
    const int ne = w.ne;
    if (ne == 0) return;                                         // no live enemy (about half of all env-steps): nothing to fold
    const bool one_pass = W::G == 32 && ne <= 32;                // every enemy has its own lane
    if (one_pass) {
        // lanes of one (cell, type) group find each other with one match instruction; every lane then folds its
        // group's ratios in list order (ascending lane): as many rounds as the largest group has members
        const bool have = lane < ne;
        float mine = 0.f;
        int loc = 0, ty = 0;
        if (have) {
            const td_enemy_rec &x = w.en()[lane];
            mine = (float)(x.LP / cc.enemy_LP[x.type_lv & 3][x.type_lv >> 2]);
            loc = x.loc;
            ty = x.type_lv & 3;
        }
        TD_CHECK(w, loc < cells && ne <= w.ecap);
        const unsigned group = __match_any_sync(kFull, have ? (unsigned)(loc * 4 + ty) : 0x80000000u + lane);
        unsigned todo = group;
        float mn = 1.f, mx = 0.f, sum = 0.f;
        while (__any_sync(kFull, todo != 0u)) {
            const int j = todo ? __ffs(todo) - 1 : lane;
            const float r = __shfl_sync(kFull, mine, j);
            if (todo) {
                mn = r < mn ? r : mn;
                mx = r > mx ? r : mx;
                sum = __fadd_rn(sum, r);
                todo &= todo - 1u;
            }
        }
        if (have && lane == __ffs(group) - 1) {
            const float cnt = (float)__popc(group);
            obs_store1(o, (size_t)(25 + ty) * cells + loc, mn);
            obs_store1(o, (size_t)(29 + ty) * cells + loc, mx);
            obs_store1(o, (size_t)(33 + ty) * cells + loc, __fdiv_rn(sum, cnt));
            obs_store1(o, (size_t)(37 + ty) * cells + loc, cnt * 0.125f);
        }
        return;
    }
    float *ratio = reinterpret_cast<float *>(w.scratch());       // [64]
    for (int e = lane; e < ne; e += W::G) {
        const td_enemy_rec &x = w.en()[e];
        ratio[e] = (float)(x.LP / cc.enemy_LP[x.type_lv & 3][x.type_lv >> 2]);
    }
    gsync(w);
    for (int e = lane; e < ne; e += W::G) {
        const int loc = w.en()[e].loc, ty = w.en()[e].type_lv & 3;
        TD_CHECK(w, loc < cells && ne <= w.ecap);
        float mn = 1.f, mx = 0.f, sum = 0.f, cnt = 0.f;
        bool leader = true;
        for (int j = 0; j < ne; ++j) {
            if (w.en()[j].loc == loc && (w.en()[j].type_lv & 3) == ty) {
                if (j < e) leader = false;
                float r = ratio[j];
                mn = r < mn ? r : mn;
                mx = r > mx ? r : mx;
                sum = __fadd_rn(sum, r);
                cnt += 1.f;
            }
        }
        if (leader) {
            obs_store1(o, (size_t)(25 + ty) * cells + loc, mn);
            obs_store1(o, (size_t)(29 + ty) * cells + loc, mx);
            obs_store1(o, (size_t)(33 + ty) * cells + loc, __fdiv_rn(sum, cnt));
            obs_store1(o, (size_t)(37 + ty) * cells + loc, cnt * 0.125f);
        }
    }
}

template <class W, class OT>
__device__ __forceinline__ void obs_sparse(W &w, OT *o)
{
    gsync(w);   // orders the dense stores (or the clears of the in-place update) before the sparse stores below
    obs_sparse_static(w, o);
    obs_sparse_towers(w, o);
    obs_sparse_enemies(w, o);
}

// The observation as an update of the previous one in the same buffer (td_step_io.obs_incremental): the 12 planes
// that broadcast a per-step scalar and the buildable plane are rewritten, the cells where towers / enemies stood
// before the step are cleared, the sparse entries of the new state are written on top.  Static map planes and
// zeros that stayed zeros are not touched: 5.2 KB instead of 18 KB of dense stores on a 10x10 board.
// Stands in for obs_dense between obs_prepare and obs_sparse.
template <class W>
__device__ __forceinline__ void obs_dense_incremental(W &w, float *o)
{
    constexpr int CELLS = W::kCells;
    constexpr int C4 = CELLS > 0 ? CELLS / 4 : 1;
    static_assert(CELLS > 0 && CELLS % 4 == 0, "specialised board sizes only");
    const int lane = w.lane;
    const float *pv = reinterpret_cast<const float *>(w.scratch()) + 64;
    float4 *o4 = reinterpret_cast<float4 *>(o);
    const uchar4 *mb = reinterpret_cast<const uchar4 *>(w.map6());
    store_run<C4, W::G>(o, 5 * C4, pv[5], lane);
#pragma unroll
    for (int k = 11; k < 14; ++k) store_run<C4, W::G>(o, k * C4, pv[k], lane);
    constexpr int kIters = (C4 + W::G - 1) / W::G;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        const int q = lane + W::G * it;
        if (q < C4) {
            const uchar4 m = mb[q];
            TD_ST(o4 + 14 * C4 + q, make_float4(m.x == 0 ? 1.f : 0.f, m.y == 0 ? 1.f : 0.f,
                                                 m.z == 0 ? 1.f : 0.f, m.w == 0 ? 1.f : 0.f));
        }
    }
#pragma unroll
    for (int k = 21; k < 25; ++k) store_run<C4, W::G>(o, k * C4, pv[k], lane);
#pragma unroll
    for (int k = 41; k < 45; ++k) store_run<C4, W::G>(o, k * C4, pv[k], lane);
    const uint32_t *old = w.old_lists();
    const int nt0 = (int)old[0], ne0 = (int)old[1];
    for (int t = lane; t < nt0; t += W::G) {
        const uint32_t key = old[4 + t];
        const int loc = key & 0xffff, tl = key >> 16;
        o[(size_t)(15 + (tl >> 2)) * CELLS + loc] = 0.f;
        o[(size_t)(17 + (tl & 3)) * CELLS + loc] = 0.f;
    }
    for (int e = lane; e < ne0; e += W::G) {
        const uint32_t key = old[4 + TD_CAP_TOWERS + e];
        const int loc = key & 0xffff, ty = key >> 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) o[(size_t)(25 + 4 * k + ty) * CELLS + loc] = 0.f;
    }
    // obs_sparse starts with the group barrier that orders these clears before the new entries
}

template <class W, class OT>
__device__ __forceinline__ void write_obs(W &w, OT *o)
{
    obs_prepare(w);
    if (!obs_dense<W::G, W, OT, true>(w, o, w.lane)) obs_sparse(w, o);
}

} // namespace td
