// td_kernels.cuh -- sm_100a kernels for the batched gym-TD board step.
//
// One warp advances one game instance.  The env record (header, tower list, enemy list, map6)
// is staged from HBM into the warp's private shared-memory slice with asynchronous 16-byte copies
// (cp.async, one round trip), all game rules run warp-cooperatively on that slice (lanes = towers
// or enemies, ballots, matches and shuffles instead of loops where the reference loops), the
// (45, L, L) float32 observation is written with unrolled 128-bit stores (or updated in place,
// INC kernels), and the live prefix of the record is written back.  Nothing stays in registers
// across the observation stores.  No block-level synchronisation exists anywhere: warps never
// communicate.  DESIGN.md section 7 lists what was measured to get here, including the dead ends.
//
// Reference semantics followed (file:line under gym_TD/envs/):
//   (a) defender decode + build/LvUp/destruct  TDDefense.py:38-77, TDMulti.py:65-115, TDBoard.py:226-293,
//                                               TDElements.py:134-170 (incl. the lvup argument swap)
//   (b) cluster summon                          TDBoard.py:199-224, TDAttack.py:36-46, TDMulti.py:88-98
//   (c) advance / leak / end condition          TDBoard.py:319-346, 370-385
//   (d) sort, target selection, damage          TDBoard.py:305-317, TDElements.py:19-28, 67-132
//   (e) reward + economy                        TDBoard.py:298-299, 315, 337-338, 348-353
//   (f) enemy statistics + observation          TDBoard.py:355-365, 85-144
//   scripted opponents                          TDGymBasic.py:81-196 (CPython `random` semantics)
//
// Compile with -fmad=false: the reference never fuses a multiply-add, and costs, LP, margins and
// rewards are compared bit for bit.
#pragma once
#include "../../include/td_b200.h"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Observation store flavour (experiments): 0 = st.global.cs (streaming), 1 = plain st.global, 2 = st.global.wt
#ifndef TD_STORE_MODE
#define TD_STORE_MODE 1
#endif
#if TD_STORE_MODE == 0
#define TD_ST(p, v) __stcs((p), (v))
#elif TD_STORE_MODE == 1
#define TD_ST(p, v) (*(p) = (v))
#elif TD_STORE_MODE == 2
#define TD_ST(p, v) __stwt((p), (v))
#else
// L2 evict_first policy on the observation stream: the env records keep their place in L2 (tools/storebench_l2.cu)
__device__ __forceinline__ void td_st_first(float4 *a, float4 v)
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void td_st_first(float *a, float v)
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}
#define TD_ST(p, v) td_st_first((p), (v))
#endif
// Debug build (-DTD_DEBUG_BOUNDS): index invariants are checked on the device and a violation sets the sticky
// flag bit 2 of the env, which TDVecEnv.stats() / the parity tests surface.  (compute-sanitizer is closed on
// the B200 pool, so this is the memory-safety net next to the bit-exact parity runs.)
#ifdef TD_DEBUG_BOUNDS
#define TD_CHECK(w, cond) do { if (!(cond)) (w).flags |= 4; } while (0)
#else
#define TD_CHECK(w, cond) do { } while (0)
#endif
#ifndef TD_WARPS_PER_CTA
#define TD_WARPS_PER_CTA 4
#endif

namespace td {

constexpr int kWarpsPerCta = TD_WARPS_PER_CTA;
constexpr unsigned kFull = 0xffffffffu;
// Env record in HBM (mirrored byte for byte in the warp's shared-memory slice):
//   [ td_env_header 64 | opponent word cache 64 | map6 cells_pad | static map (MapHdr 16, cells, dist) |
//     towers 32 x 16 | enemies 64 x 24 ]
// Everything a step normally needs sits in two contiguous prefixes, fetched in ONE round trip: the first
// runs from the header to tower kSpecTowers, the second covers enemies [0, kSpecEnemies).
constexpr int kOffRngCache = 64;          // td_env_header, then the cached generator words, then map6
constexpr int kRngCacheDef = 16;          // words cached per env: the defender env's attacker draws ~4 per step,
constexpr int kRngCacheAtk = 64;          // the attacker env's scripted defender up to ~60 (shuffle of the road cells)
constexpr int kTowerBytes = 16;
constexpr int kEnemyBytes = 24;
constexpr int kMapHdrBytes = 16;
constexpr int kMtWords = 624;
// behind the scratch area of a slice: the tower / enemy cells of the env before the step (incremental observation)
constexpr int kOldListBytes = 16 + 4 * TD_CAP_TOWERS + 4 * TD_CAP_ENEMIES;
constexpr int kTwistStageBytes = kMtWords * 4;   // staging area of the generator regeneration (tail of a slice)
// Speculatively staged list prefixes.  Lists are short in practice (tools/state_hist.py: towers p99 = 10 in the
// defender env, 11-13 in the attacker env whose scripted defender keeps building; live enemies p99 = 6): 12 towers
// and 8 enemies cover almost every env and read 256 B less per env-step than 16 / 16 (def-small 0.2156 -> 0.2132 ms,
// atk-small 0.2435 -> 0.2397 ms; 8 / 8 is best for def-small alone, 0.2126 ms, and neutral for atk-small).
#ifndef TD_SPEC_TOWERS
#define TD_SPEC_TOWERS 12
#endif
#ifndef TD_SPEC_ENEMIES
#define TD_SPEC_ENEMIES 8
#endif
constexpr int kSpecTowers = TD_SPEC_TOWERS;     // speculatively staged list prefixes
constexpr int kSpecEnemies = TD_SPEC_ENEMIES;

// Derived constant tables (uploaded by td_set_config).
struct DevConfig {
    double enemy_LP[TD_NTYPES][TD_NLV];
    double enemy_speed[TD_NTYPES][TD_NLV];
    double enemy_defense[TD_NTYPES][TD_NLV];
    double enemy_cost[TD_NTYPES][TD_NLV];
    double tower_attack[TD_NTYPES][TD_NLV];
    double tower_cost[TD_NTYPES][TD_NLV];
    double tower_intv[TD_NTYPES][TD_NLV];     // effective interval by (type, lv): lv1 = tower_cost[t][1] (sic)
    double tower_refund[TD_NTYPES][TD_NLV];   // Tower.cost by (type, lv): lv1 = cost[t][0] + interval[t][1] (sic)
    int tower_range[TD_NTYPES][TD_NLV];
    int tower_splash[TD_NTYPES][TD_NLV];
    double destruct_return, frozen_ratio, atk_init_cost, def_init_cost, max_cost;
    double reward_kill, penalty_leak, reward_time, rate_init, rate_final, def_rate, upgrade_at;
    int frozen_time, base_LP, tower_distance, atk_interval, def_interval, max_steps;
    int upgrade_step;       // smallest s with (double)s / max_steps >= enemy_upgrade_at (TDBoard.py:201 without the division)
    double min_enemy_cost[TD_NLV];   // cheapest enemy type per level: below it every remaining cluster slot fails
};

struct MapHdr {            // 16 bytes, head of a map-pool record
    uint16_t start[3];
    uint16_t end;
    uint8_t num_roads;
    uint8_t maxd_p1;       // max(map[4]) + 1
    uint8_t pad[6];
};

struct EnvStats {          // 32 bytes per env, written only when an episode ends
    double return_sum;
    uint32_t episodes, wins, length_sum, kills, leaks, flags;
};

struct StepParams {
    uint8_t *records;          // [n][record_bytes]
    const uint8_t *maps;       // [n_maps][map_bytes]
    uint32_t *mt;              // [n][624] scripted-opponent generator words (may be NULL)
    EnvStats *stats;           // [n]
    int n_envs, n_maps, map_stride;
    int env_begin;             // the step kernel covers envs [env_begin, n_envs) (chunked host-path launches)
    int L, cells, cells_pad, record_bytes, map_bytes, smem_per_warp, scratch_off;
    int off_static, off_towers, off_enemies, rng_cache_words;
    int difficulty;
    int opponent_seeded;
    int old_lists_off;         // offset of the pre-step tower / enemy cell lists inside a slice
    int act_stage_off;         // offset of the attacker's (3, 8) int64 action / RealAction inside a slice (ATK, 2P)
    td_step_io io;
    DevConfig cfg;             // per handle: travels with every launch in the kernel-parameter constant bank
};

// ------------------------------------------------------------------------------------------------
// per-warp context: pointers into the warp's shared-memory slice + uniform registers

// CELLS > 0: board size known at compile time -> every pointer into the slice is base + constant.
// CELLS == 0: layout read from the kernel parameters (constant bank).
// GW = lanes per game instance: 32 (one warp per env) or 16 (two envs per warp: the uniform bookkeeping
// of both is issued once, and twice as many envs are in flight per SM at the same warp count).
template <int CELLS, int GW, int RC = 0>
struct Ctx {
    static constexpr int kCells = CELLS;
    static constexpr int G = GW;
    static constexpr int kRngWords = RC;
    unsigned gmask;            // the warp lanes of this env's group
    int gbase;                 // first warp lane of the group
    static constexpr int kL = CELLS == 100 ? 10 : CELLS == 400 ? 20 : CELLS == 900 ? 30 : 0;
    static constexpr int kPad = (CELLS + 15) & ~15;
    uint8_t *slice;            // the warp's shared-memory slice: [record | scratch]
    const StepParams *pp;
    int lane, ecap;            // lane = index inside the group
    // uniform copies of hot header fields (identical in all lanes of the group)
    double cost_def, cost_atk;
    int nt, ne, base_LP, steps, def_cd, atk_cd, fail, flags;
    // opponent generator cursor
    uint32_t *mt;
    int mt_pos;
    int ck, cn;                 // consumed / valid words of the (tempered) word cache
    bool static_dirty;          // the record's static map was replaced (reset)
    bool cache_dirty;           // the word cache changed this step (it goes back to the record)
    bool cache_raw;             // ... by the asynchronous top-up: its words still have to be tempered

    __device__ __forceinline__ int L() const { return CELLS ? kL : pp->L; }
    __device__ __forceinline__ int ncells() const { return CELLS ? CELLS : pp->cells; }
    __device__ __forceinline__ int cells_pad() const { return CELLS ? kPad : pp->cells_pad; }
    __device__ __forceinline__ int map_bytes() const { return kMapHdrBytes + 2 * cells_pad(); }
    // RC > 0: cached generator words known at compile time; RC == 0: read from the kernel parameters
    __device__ __forceinline__ int rng_words() const { return RC ? RC : pp->rng_cache_words; }
    __device__ __forceinline__ int hdr_bytes() const { return kOffRngCache + 4 * rng_words(); }
    __device__ __forceinline__ int off_static() const { return hdr_bytes() + cells_pad(); }
    __device__ __forceinline__ int off_towers() const { return off_static() + map_bytes(); }
    __device__ __forceinline__ int off_enemies() const { return off_towers() + TD_CAP_TOWERS * kTowerBytes; }
    __device__ __forceinline__ int record_bytes() const { return off_enemies() + TD_CAP_ENEMIES * kEnemyBytes; }
    __device__ __forceinline__ td_env_header *hdr() const { return reinterpret_cast<td_env_header *>(slice); }
    __device__ __forceinline__ const uint32_t *rng_cache() const { return reinterpret_cast<const uint32_t *>(slice + kOffRngCache); }
    __device__ __forceinline__ uint8_t *map6() const { return slice + hdr_bytes(); }
    __device__ __forceinline__ MapHdr *mh() const { return reinterpret_cast<MapHdr *>(slice + off_static()); }
    __device__ __forceinline__ uint8_t *cells() const { return slice + off_static() + kMapHdrBytes; }
    __device__ __forceinline__ uint8_t *dist() const { return slice + off_static() + kMapHdrBytes + cells_pad(); }
    __device__ __forceinline__ td_tower_rec *tw() const { return reinterpret_cast<td_tower_rec *>(slice + off_towers()); }
    __device__ __forceinline__ td_enemy_rec *en() const { return reinterpret_cast<td_enemy_rec *>(slice + off_enemies()); }
    __device__ __forceinline__ uint8_t *scratch() const { return slice + record_bytes(); }
    __device__ __forceinline__ uint32_t *old_lists() const
    {
        return reinterpret_cast<uint32_t *>(slice + pp->old_lists_off);
    }
    // the attacker's cluster action, later its RealAction: 24 int64 staged with the record (ATK / 2P envs)
    __device__ __forceinline__ long long *act_stage() const
    {
        return reinterpret_cast<long long *>(slice + pp->act_stage_off);
    }
    // tail of the slice (envs with a scripted opponent only): staging area of the generator regeneration
    __device__ __forceinline__ uint32_t *twist_stage() const
    {
        return reinterpret_cast<uint32_t *>(slice + pp->smem_per_warp - kTwistStageBytes);
    }
};

// group-level primitives: ballots are returned relative to the group (bit 0 = group lane 0)
template <class W> __device__ __forceinline__ unsigned gballot(const W &w, bool pred)
{
    const unsigned b = __ballot_sync(w.gmask, pred);
    if (W::G == 32) return b;
    return (b >> w.gbase) & 0xffffu;
}
template <class W, class T> __device__ __forceinline__ T gshfl(const W &w, T v, int src)
{
    return __shfl_sync(w.gmask, v, src, W::G);
}
template <class W> __device__ __forceinline__ bool gall(const W &w, bool pred) { return __all_sync(w.gmask, pred); }
template <class W> __device__ __forceinline__ bool gany(const W &w, bool pred) { return __any_sync(w.gmask, pred); }
template <class W> __device__ __forceinline__ void gsync(const W &w) { __syncwarp(w.gmask); }

// n16 <= MAXN int4 with a compile-time bound: one predicated load/store pair per pass instead of a loop.
template <int MAXN, int G>
__device__ __forceinline__ void copy16_upto(void *dst, const void *src, int n16, int lane)
{
    int4 *d = reinterpret_cast<int4 *>(dst);
    const int4 *s = reinterpret_cast<const int4 *>(src);
#pragma unroll
    for (int k = 0; k < (MAXN + G - 1) / G; ++k) {
        const int q = lane + G * k;
        if (q < n16) d[q] = s[q];
    }
}

__device__ __forceinline__ void warp_copy16(void *dst, const void *src, int n16, int lane, int stride)
{
    int4 *d = reinterpret_cast<int4 *>(dst);
    const int4 *s = reinterpret_cast<const int4 *>(src);
    for (int q = lane; q < n16; q += stride) d[q] = s[q];
}

// global -> shared, 16 bytes per lane per instruction, asynchronous (LDGSTS): every segment of a stage
// is in flight at once and no register is held while the data travels.
__device__ __forceinline__ void async_copy16(void *smem_dst, const void *gmem_src, int n16, int lane, int stride)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const char *s = reinterpret_cast<const char *>(gmem_src);
    for (int q = lane; q < n16; q += stride)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * q), "l"(s + 16 * q) : "memory");
}
template <class W>
__device__ __forceinline__ void async_wait_all(const W &w)
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    gsync(w);
}

template <class W>
__device__ __forceinline__ void ctx_bind(W &w, uint8_t *slice, const StepParams &p)
{
    w.slice = slice;
    w.pp = &p;
    const int wl = threadIdx.x & 31;
    w.lane = wl & (W::G - 1);
    w.gbase = wl - w.lane;
    w.gmask = W::G == 32 ? kFull : (0xffffu << w.gbase);
    w.ecap = TD_CAP_ENEMIES;
    w.mt = nullptr;
    w.mt_pos = 0;
    w.ck = 0; w.cn = 0;
    w.static_dirty = false;
    w.cache_dirty = false;
    w.cache_raw = false;
}

template <class W>
__device__ __forceinline__ void load_static_map(W &w, const StepParams &p, int map_id)
{
    async_copy16(w.mh(), p.maps + (size_t)map_id * w.map_bytes(), w.map_bytes() >> 4, w.lane, W::G);
    w.static_dirty = true;
}

template <class W>
__device__ __forceinline__ void pull_header(W &w)
{
    const td_env_header *h = w.hdr();
    w.cost_def = h->cost_def;
    w.cost_atk = h->cost_atk;
    w.nt = h->n_towers;
    w.ne = h->n_enemies;
    w.base_LP = h->base_LP;
    w.steps = h->steps;
    w.def_cd = h->defender_cd;
    w.atk_cd = h->attacker_cd;
    w.flags = h->flags;
    w.fail = TD_FC_SUCCESS;
    w.mt_pos = h->rng_pos;
    w.ck = h->pad1;             // cached generator words already consumed / valid (the cache holds tempered words)
    w.cn = h->pad0;
}

template <class W>
__device__ __forceinline__ void push_header(W &w)
{
    if (w.lane == 0) {
        td_env_header *h = w.hdr();
        h->cost_def = w.cost_def;
        h->cost_atk = w.cost_atk;
        h->n_towers = (uint8_t)w.nt;
        h->n_enemies = (uint8_t)w.ne;
        h->base_LP = w.base_LP;
        h->steps = w.steps;
        h->defender_cd = (int16_t)w.def_cd;
        h->attacker_cd = (int16_t)w.atk_cd;
        h->flags = (uint8_t)w.flags;
        h->rng_pos = w.mt_pos;
        h->pad0 = (uint8_t)w.cn;
        h->pad1 = w.ck;
    }
}

// ------------------------------------------------------------------------------------------------
// CPython-compatible MT19937 consumer (random.Random): one tempered window of <= 32 words per fill

// Regenerate the 624 words in place.  mt[k] = mt[(k+397)%624] ^ f(mt[k], mt[k+1]); chunks of 32 words in
// ascending order keep every operand in the state (old / new) the sequential algorithm sees.
__device__ __noinline__ void mt_twist(uint32_t *mt, int lane, int stride, unsigned gmask)
{
#pragma unroll 1
    for (int base = 0; base < kMtWords; base += stride) {
        const int k = base + lane;
        uint32_t v = 0;
        if (k < kMtWords - 1) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
            v = mt[k < 227 ? k + 397 : k - 227] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp(gmask);
        if (k < kMtWords - 1) mt[k] = v;
        __syncwarp(gmask);
    }
    if (lane == 0) {
        uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    __syncwarp(gmask);
}

// The same through shared memory: the 20 dependent chunks of the regeneration cost one HBM round trip each when
// run on the state in place (40 us per twist under load); staged, the state travels once in and once out.
// gmt is 16-byte aligned (624 words per env), smt is a 2496-byte staging area in the group's slice.
__device__ __noinline__ void mt_twist_staged(uint32_t *gmt, uint32_t *smt, int lane, int stride, unsigned gmask)
{
    for (int q = lane; q < kMtWords / 4; q += stride) reinterpret_cast<int4 *>(smt)[q] = reinterpret_cast<const int4 *>(gmt)[q];
    __syncwarp(gmask);
    mt_twist(smt, lane, stride, gmask);
    for (int q = lane; q < kMtWords / 4; q += stride) reinterpret_cast<int4 *>(gmt)[q] = reinterpret_cast<const int4 *>(smt)[q];
    __syncwarp(gmask);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// The generator words of a step are consumed from the record's word cache, which holds TEMPERED words: a draw is
// one broadcast read from shared memory.  The cache is topped up behind the observation stores only when less
// than half of it is left (refill / finish_refill: most steps neither read the generator state nor write the cache
// back).  Only when a step needs more words than the cache holds (the scripted defender's shuffle, a few percent
// of its steps) the out-of-line refill fetches them from the generator state in HBM, twisting it when exhausted.
struct MtRefill { int cn, mt_pos; };
__device__ __noinline__ MtRefill mt_refill(uint32_t *cache, uint32_t *mt, uint32_t *stage, int lane, int G, unsigned gmask,
                                           int mt_pos, int words)
{
    if (mt_pos >= kMtWords) { mt_twist_staged(mt, stage, lane, G, gmask); mt_pos = 0; }
    const int n = min(words, kMtWords - mt_pos);
    __syncwarp(gmask);                                  // every earlier read of the cache is done
    for (int q = lane; q < n; q += G) cache[q] = mt_temper(mt[mt_pos + q]);
    __syncwarp(gmask);
    MtRefill r;
    r.cn = n;
    r.mt_pos = mt_pos;
    return r;
}

template <class W>
__device__ __forceinline__ void mt_more_words(W &w)
{
    const MtRefill r = mt_refill(const_cast<uint32_t *>(w.rng_cache()), w.mt, w.twist_stage(), w.lane, W::G, w.gmask,
                                 w.mt_pos, w.rng_words());
    w.cn = r.cn;
    w.mt_pos = r.mt_pos;
    w.ck = 0;
    w.cache_dirty = true;
}

// raw words just copied from the generator state -> tempered, in place
template <class W>
__device__ __forceinline__ void temper_cache(W &w)
{
    uint32_t *cache = const_cast<uint32_t *>(w.rng_cache());
    constexpr int kIters = W::kRngWords > 0 ? (W::kRngWords + W::G - 1) / W::G : 0;
    if (kIters > 0) {
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int q = w.lane + W::G * it;
            if (q < w.cn) cache[q] = mt_temper(cache[q]);
        }
    } else {
        for (int q = w.lane; q < w.cn; q += W::G) cache[q] = mt_temper(cache[q]);
    }
    gsync(w);
}

template <class W>
__device__ __forceinline__ uint32_t mt_next(W &w)
{
    if (__builtin_expect(w.ck >= w.cn, 0)) mt_more_words(w);
    TD_CHECK(w, w.ck >= 0 && w.ck < w.cn && w.cn <= w.rng_words() && w.mt_pos < kMtWords);
    const uint32_t r = w.rng_cache()[w.ck];
    ++w.ck;
    ++w.mt_pos;
    return r;
}

// Stage one env in one round trip: [header .. tower kSpecTowers) and enemies [0, kSpecEnemies) are fetched
// speculatively; only envs with longer lists pay a second trip for the rest.
template <class W>
__device__ __forceinline__ void issue_env_load(W &w, const uint8_t *rec)
{
    async_copy16(w.slice, rec, (w.off_towers() + kSpecTowers * kTowerBytes) >> 4, w.lane, W::G);
    async_copy16(w.en(), rec + w.off_enemies(), (kSpecEnemies * kEnemyBytes) >> 4, w.lane, W::G);
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// The speculative part has landed (caller waited): read the header, fetch the rare remainder.
template <class W>
__device__ __forceinline__ void finish_env_load(W &w, const StepParams &p, const uint8_t *rec, uint32_t *mt_base)
{
    pull_header(w);
    w.mt = mt_base;
    if (__builtin_expect(w.nt > kSpecTowers || w.ne > kSpecEnemies, 0)) {
        if (w.nt > kSpecTowers)
            async_copy16(w.tw() + kSpecTowers, rec + w.off_towers() + kSpecTowers * kTowerBytes, w.nt - kSpecTowers, w.lane, W::G);
        if (w.ne > kSpecEnemies)
            async_copy16(w.en() + kSpecEnemies, rec + w.off_enemies() + kSpecEnemies * kEnemyBytes,
                         ((w.ne - kSpecEnemies) * 3 + 1) >> 1, w.lane, W::G);
        async_wait_all(w);
    }
}

template <class W>
__device__ __forceinline__ void load_env(W &w, const StepParams &p, const uint8_t *rec, uint32_t *mt_base = nullptr)
{
    issue_env_load(w, rec);
    async_wait_all(w);
    finish_env_load(w, p, rec, mt_base);
}

// Write back the header block (incl. the word cache), the changed maps and the live list prefixes.
// UNROLLED: predicated single-pass copies instead of loops -- 45 fewer instructions per env-step.  Measured on
// B200: attacker env 0.291 -> 0.273 ms, in-place observation update (def-small) 0.191 -> 0.167 ms, but the
// full-write defender step 0.2175 -> 0.2220 ms (same box, twice), so the caller chooses per kernel variant.
template <bool UNROLLED = false, class W>
__device__ __forceinline__ void store_env(W &w, const StepParams &p, uint8_t *rec, bool map6_dirty, bool push = true)
{
    if (push) push_header(w);
    gsync(w);
    const int head = w.static_dirty ? w.off_towers() : (map6_dirty ? w.off_static() : w.hdr_bytes());
    // the word cache block [kOffRngCache, hdr_bytes) goes back only when it was refilled
    const int skip_lo = w.cache_dirty ? 0 : (kOffRngCache >> 4), skip_hi = w.cache_dirty ? 0 : (w.hdr_bytes() >> 4);
    if (!UNROLLED) {
        for (int q = w.lane; q < (head >> 4); q += W::G)
            if (q < skip_lo || q >= skip_hi) reinterpret_cast<int4 *>(rec)[q] = reinterpret_cast<const int4 *>(w.slice)[q];
        warp_copy16(rec + w.off_towers(), w.tw(), w.nt, w.lane, W::G);
        warp_copy16(rec + w.off_enemies(), w.en(), (w.ne * 3 + 1) >> 1, w.lane, W::G);
        return;
    }
    if (W::kCells > 0 && W::kRngWords > 0) {
        constexpr int kHeadMax = (kOffRngCache + 4 * W::kRngWords + 3 * W::kPad + kMapHdrBytes) / 16;   // = off_towers / 16
#pragma unroll
        for (int k = 0; k < (kHeadMax + W::G - 1) / W::G; ++k) {
            const int q = w.lane + W::G * k;
            if (q < (head >> 4) && (q < skip_lo || q >= skip_hi))
                reinterpret_cast<int4 *>(rec)[q] = reinterpret_cast<const int4 *>(w.slice)[q];
        }
    } else {
        for (int q = w.lane; q < (head >> 4); q += W::G)
            if (q < skip_lo || q >= skip_hi) reinterpret_cast<int4 *>(rec)[q] = reinterpret_cast<const int4 *>(w.slice)[q];
    }
    copy16_upto<TD_CAP_TOWERS, W::G>(rec + w.off_towers(), w.tw(), w.nt, w.lane);
    copy16_upto<(TD_CAP_ENEMIES * 3 + 1) / 2, W::G>(rec + w.off_enemies(), w.en(), (w.ne * 3 + 1) >> 1, w.lane);
}

// TDGymBasic.reset (:37-55) + TDBoard.__init__ (:63-79): fresh episode on map `map_id`.
template <class W>
__device__ __forceinline__ void reset_env(W &w, const StepParams &p, int map_id, bool reload_map)
{
    const DevConfig &cc = w.pp->cfg;
    if (reload_map) {
        gsync(w);
        load_static_map(w, p, map_id);
        async_wait_all(w);
    }
    for (int q = w.lane; q < (w.cells_pad() >> 2); q += W::G) {
        uint32_t c4 = reinterpret_cast<const uint32_t *>(w.cells())[q];
        reinterpret_cast<uint32_t *>(w.map6())[q] = c4 & 0x01010101u;           // map[6] = 1 on road cells
    }
    w.cost_def = cc.def_init_cost;
    w.cost_atk = cc.atk_init_cost;
    w.nt = 0;
    w.ne = 0;
    w.base_LP = cc.base_LP;
    w.steps = 0;
    w.def_cd = 0;
    w.atk_cd = 0;
    w.fail = TD_FC_SUCCESS;
    if (w.lane == 0) {
        w.hdr()->map_id = map_id;
        w.hdr()->ep_return = 0.0;
        w.hdr()->ep_kills = 0;
        w.hdr()->ep_leaks = 0;
    }
    gsync(w);
}

// random._randbelow_with_getrandbits(n), 1 <= n < 2^31
template <class W>
__device__ __forceinline__ int py_randbelow(W &w, int n)
{
    const int shift = __clz(n);      // 32 - bit_length(n)
    uint32_t r;
    do { r = mt_next(w) >> shift; } while (r >= (uint32_t)n);
    return (int)r;
}

// random.shuffle(list) (for i in reversed(range(1, n)): j = randbelow(i + 1); swap) on a uint16 list in shared
// memory.  The draws are serial by definition (rejections shift every later draw), so one lane runs the whole
// loop alone, straight over the tempered word cache -- a fifth of the instructions of the same loop with a
// group-wide draw per element.
template <class W>
__device__ __forceinline__ void py_shuffle_u16(W &w, uint16_t *list, int n)
{
    int i = n - 1;
    while (i >= 1) {
        if (w.ck >= w.cn) mt_more_words(w);
        TD_CHECK(w, w.cn <= w.rng_words() && w.mt_pos + (w.cn - w.ck) <= kMtWords);
        int k = w.ck;
        if (w.lane == 0) {
            const uint32_t *words = w.rng_cache();
            const int end = w.cn;
            while (i >= 1 && k < end) {
                const uint32_t r = words[k++] >> __clz(i + 1);
                if (r <= (uint32_t)i) {
                    const uint16_t t = list[i];
                    list[i] = list[r];
                    list[r] = t;
                    --i;
                }
            }
        }
        k = gshfl(w, k, 0);
        i = gshfl(w, i, 0);
        w.mt_pos += k - w.ck;
        w.ck = k;
        gsync(w);
    }
}

template <class W>
__device__ __forceinline__ double py_random(W &w)
{
    uint32_t a = mt_next(w) >> 5, b = mt_next(w) >> 6;
    return __dmul_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------
// (a) defender operations -- all arguments and results are warp-uniform

// map[6] += delta on the Manhattan diamond around `loc` (TDBoard.py:239-245, 281-287).  Out of line: it is
// reached from several build / destruct sites and only on the rare successful operation.
__device__ __noinline__ void diamond_add_cells(uint8_t *map6, int loc, int delta, int L, int D, int lane, int stride,
                                               unsigned gmask)
{
    const int WD = 2 * D + 1;
    const int r0 = loc / L, c0 = loc - r0 * L;
    for (int k = lane; k < WD * WD; k += stride) {
        int i = k / WD - D, j = k % WD - D;
        int r = r0 + i, c = c0 + j;
        if (abs(i) + abs(j) <= D && r >= 0 && r < L && c >= 0 && c < L)
            map6[r * L + c] = (uint8_t)(map6[r * L + c] + delta);
    }
    __syncwarp(gmask);
}

template <class W>
__device__ __forceinline__ void diamond_add(W &w, int loc, int delta)
{
    diamond_add_cells(w.map6(), loc, delta, w.L(), w.pp->cfg.tower_distance, w.lane, W::G, w.gmask);
}

template <class W>
__device__ __forceinline__ bool tower_build(W &w, int t, int loc, bool &map6_dirty)   // TDBoard.py:226-247
{
    const DevConfig &cc = w.pp->cfg;
    const double cost = cc.tower_cost[t][0];
    if (w.cost_def < cost) { w.fail = TD_FC_COST_SHORTAGE; return false; }
    if (w.map6()[loc] > 0) { w.fail = TD_FC_INVALID_POSITION; return false; }
    if (w.nt >= TD_CAP_TOWERS) { w.flags |= 2; w.fail = TD_FC_INVALID_POSITION; return false; }
    TD_CHECK(w, loc >= 0 && loc < w.ncells() && t >= 0 && t < TD_NTYPES);
    if (w.lane == 0) {
        td_tower_rec &r = w.tw()[w.nt];
        r.cd = 0.0;
        r.loc = (uint16_t)loc;
        r.type_lv = (uint8_t)t;
    }
    w.nt += 1;
    w.cost_def = __dsub_rn(w.cost_def, cost);
    diamond_add(w, loc, +1);
    map6_dirty = true;
    w.fail = TD_FC_SUCCESS;
    return true;
}

template <class W>
__device__ __forceinline__ int find_tower(const W &w, int loc)
{
    int idx = -1;
    for (int base = 0; base < w.nt; base += W::G) {
        const int t = base + w.lane;
        const unsigned b = gballot(w, t < w.nt && w.tw()[t].loc == loc);
        if (b && idx < 0) idx = base + __ffs(b) - 1;
    }
    return idx;
}

template <class W>
__device__ __forceinline__ bool tower_lvup(W &w, int loc)                              // TDBoard.py:249-271
{
    const DevConfig &cc = w.pp->cfg;
    int idx = find_tower(w, loc);
    if (idx < 0) { w.fail = TD_FC_UNKNOWN_TARGET; return false; }
    int tl = w.tw()[idx].type_lv, ty = tl & 3, lv = tl >> 2;
    if (lv >= TD_NLV - 1) { w.fail = TD_FC_LV_MAX; return false; }
    const double cost = cc.tower_cost[ty][lv + 1];
    if (w.cost_def < cost) { w.fail = TD_FC_COST_SHORTAGE; return false; }
    gsync(w);
    if (w.lane == 0) w.tw()[idx].type_lv = (uint8_t)(ty | ((lv + 1) << 2));
    gsync(w);
    w.cost_def = __dsub_rn(w.cost_def, cost);
    w.fail = TD_FC_SUCCESS;
    return true;
}

template <class W>
__device__ __forceinline__ bool tower_destruct(W &w, int loc, bool &map6_dirty)        // TDBoard.py:273-293
{
    const DevConfig &cc = w.pp->cfg;
    int idx = find_tower(w, loc);
    if (idx < 0) { w.fail = TD_FC_UNKNOWN_TARGET; return false; }
    int tl = w.tw()[idx].type_lv;
    double c = __dadd_rn(w.cost_def, __dmul_rn(cc.tower_refund[tl & 3][tl >> 2], cc.destruct_return));
    w.cost_def = cc.max_cost < c ? cc.max_cost : c;
    // remove from the list, keeping the order of the rest
    constexpr int kPasses = TD_CAP_TOWERS / W::G;
    td_tower_rec mine[kPasses];
#pragma unroll
    for (int q = 0; q < kPasses; ++q) {
        const int t = w.lane + W::G * q;
        if (t > idx && t < w.nt) mine[q] = w.tw()[t];
    }
    gsync(w);
#pragma unroll
    for (int q = 0; q < kPasses; ++q) {
        const int t = w.lane + W::G * q;
        if (t > idx && t < w.nt) w.tw()[t - 1] = mine[q];
    }
    gsync(w);
    w.nt -= 1;
    diamond_add(w, loc, -1);
    map6_dirty = true;
    w.fail = TD_FC_SUCCESS;
    return true;
}

// Discrete action (TDDefense.py:61-77, TDMulti.py:100-115).  Returns success.
template <class W>
__device__ __forceinline__ bool decode_discrete(W &w, long long a, long long &real, int &failcode, bool &dirty)
{
    const DevConfig &cc = w.pp->cfg;
    const long long nop = 6ll * w.ncells();
    real = nop;
    failcode = 0;
    if (w.def_cd != 0 || a == nop || (unsigned long long)a > (unsigned long long)nop) return false;
    int ai = (int)a;
    int act = ai / w.ncells(), loc = ai - act * w.ncells();
    bool res;
    if (act < TD_NTYPES) res = tower_build(w, act, loc, dirty);
    else if (act == TD_NTYPES) res = tower_lvup(w, loc);
    else res = tower_destruct(w, loc, dirty);
    if (res) { w.def_cd = cc.def_interval; real = a; }
    failcode = w.fail;
    return res;
}

// Multi-action Box(6, L, L) (TDDefense.py:40-60, TDMulti.py:65-84): r-major, c, then build 0..3, LvUp,
// destruct inside a cell, every operation seeing the state left by the previous one.  32 cells are
// screened per pass; a cell is skipped when none of its flagged operations can succeed in the
// current state (no tower on it, and no flagged build that is both affordable and placeable).  The
// screen is recomputed after every success because cost and map6 then change.
template <class W>
__device__ __forceinline__ void decode_multi(W &w, const long long *act, long long *real, bool &dirty)
{
    const DevConfig &cc = w.pp->cfg;
    const int cells = w.ncells();
    uint8_t *tower_at = w.scratch();      // cells bytes: 1 where a tower stands (scratch >= cells_pad here)
    const bool enabled = w.def_cd == 0;
    for (int q = w.lane; q < (w.cells_pad() >> 2); q += W::G) reinterpret_cast<uint32_t *>(tower_at)[q] = 0u;
    gsync(w);
    for (int t = w.lane; t < w.nt; t += W::G) tower_at[w.tw()[t].loc] = 1;
    gsync(w);
    for (int base = 0; base < cells; base += W::G) {
        const int cell = base + w.lane;
        unsigned flags = 0;          // bit ch set when action[ch][cell] == 1
        if (cell < cells) {
#pragma unroll
            for (int ch = 0; ch < 6; ++ch) {
                long long v = __ldcs(act + (size_t)ch * cells + cell);
                flags |= (v == 1 ? 1u : 0u) << ch;
            }
        }
        unsigned done_mask = 0;      // successes of this lane's cell
        if (enabled) {
            unsigned pending = gballot(w, flags != 0);
            while (pending) {
                // screen with the current state
                bool can = false;
                if (flags) {
                    if (tower_at[cell]) can = (flags & 0x30u) != 0 || false;
                    if (!can && (flags & 0x0fu) && w.map6()[cell] == 0) {
#pragma unroll
                        for (int t = 0; t < TD_NTYPES; ++t)
                            can = can || (((flags >> t) & 1u) && !(w.cost_def < cc.tower_cost[t][0]));
                    }
                    // a flagged build on a free cell can create the tower that a flagged LvUp/destruct then hits
                }
                unsigned cand = gballot(w, can) & pending;
                if (!cand) break;
                int src = __ffs(cand) - 1;
                unsigned f = gshfl(w, flags, src);
                int loc = base + src;
                unsigned ok = 0;
                for (int t = 0; t < TD_NTYPES; ++t)
                    if ((f >> t) & 1u) if (tower_build(w, t, loc, dirty)) { ok |= 1u << t; if (w.lane == 0) tower_at[loc] = 1; gsync(w); }
                if ((f >> 4) & 1u) if (tower_lvup(w, loc)) ok |= 1u << 4;
                if ((f >> 5) & 1u) if (tower_destruct(w, loc, dirty)) { ok |= 1u << 5; if (w.lane == 0) tower_at[loc] = 0; gsync(w); }
                if (ok) w.def_cd = cc.def_interval;
                if (w.lane == src) done_mask = ok;
                // cells up to and including src are finished
                pending &= ~((2u << src) - 1u);
            }
        }
        if (cell < cells && real) {
#pragma unroll
            for (int ch = 0; ch < 6; ++ch) __stcs(real + (size_t)ch * cells + cell, (long long)((done_mask >> ch) & 1u));
        }
    }
    gsync(w);
}

// ------------------------------------------------------------------------------------------------
// (b) summon

template <class W>
__device__ __forceinline__ void append_enemy(W &w, int t, int lv, int start)
{
    const DevConfig &cc = w.pp->cfg;
    if (w.ne >= w.ecap) { w.flags |= 1; return; }
    if (w.lane == 0) {
        td_enemy_rec &e = w.en()[w.ne];
        e.LP = cc.enemy_LP[t][lv];
        e.margin = 0.0;
        e.loc = (uint16_t)start;
        e.type_lv = (uint8_t)(t | (lv << 2));
        e.slowdown = 0;
    }
    w.ne += 1;
}

// TDBoard.py:199-224 for one road.  `mine` is this lane's slot value (lanes lane_base..lane_base+7 hold the
// cluster); updated in place to the RealAction value.  Returns the bool of the (bool, list) tuple.
// The eight types are packed into three ballots, the f64 cost chain runs on uniform registers, and the
// affordable slots append their enemies in one parallel store (list order = slot order).
template <class W>
__device__ __forceinline__ bool summon_cluster(W &w, int road, long long &mine, int lane_base)
{
    const DevConfig &cc = w.pp->cfg;
    const int start = w.mh()->start[road];
    const int lv = w.steps >= cc.upgrade_step ? 1 : 0;      // progress >= enemy_upgrade_at
    const int tv = (mine < 0 || mine >= TD_NTYPES) ? TD_NTYPES : (int)mine;     // 4 == enemy_types: empty slot
    const unsigned b0 = gballot(w, tv & 1) >> lane_base, b1 = gballot(w, tv & 2) >> lane_base,
                   b2 = gballot(w, tv & 4) >> lane_base;
    unsigned todo = ~b2 & 0xffu;                     // slots holding a real type (0..3)
    const bool tried = todo != 0;
    unsigned ok = 0, poor = 0;
    const double cheapest = cc.min_enemy_cost[lv];
    while (todo) {
        if (w.cost_atk < cheapest) { poor |= todo; break; }      // an empty purse fails every remaining slot alike
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int t = ((b0 >> k) & 1) | (((b1 >> k) & 1) << 1);
        const double cost = cc.enemy_cost[t][lv];
        if (w.cost_atk < cost) poor |= 1u << k;
        else { w.cost_atk = __dsub_rn(w.cost_atk, cost); ok |= 1u << k; }
    }
    int n = __popc(ok);
    if (n > w.ecap - w.ne) { w.flags |= 1; n = w.ecap - w.ne; }
    const int k = w.lane - lane_base;
    if (k >= 0 && k < TD_CLUSTER) {
        if ((poor >> k) & 1u) mine = TD_NTYPES;
        const int idx = __popc(ok & ((1u << k) - 1u));
        if (((ok >> k) & 1u) && idx < n) {
            td_enemy_rec &e = w.en()[w.ne + idx];
            e.LP = cc.enemy_LP[tv][lv];
            e.margin = 0.0;
            e.loc = (uint16_t)start;
            e.type_lv = (uint8_t)(tv | (lv << 2));
            e.slowdown = 0;
        }
    }
    w.ne += n;
    gsync(w);
    if (ok == 0 && tried) { w.fail = TD_FC_COST_SHORTAGE; return false; }
    w.fail = TD_FC_SUCCESS;
    return true;
}

// scripted attacker of the defender env: 8 x type t on one road (TDGymBasic.py:95-108 -> TDBoard.py:199-224).
// All eight slots cost the same, so the first unaffordable slot ends the cluster; the summoned enemies are
// appended by eight lanes at once.
template <class W>
__device__ __forceinline__ void summon_uniform(W &w, int t, int road)
{
    const DevConfig &cc = w.pp->cfg;
    const int start = w.mh()->start[road];
    const int lv = w.steps >= cc.upgrade_step ? 1 : 0;      // progress >= enemy_upgrade_at
    const double cost = cc.enemy_cost[t][lv];
    int n = 0;
#pragma unroll 1
    for (int k = 0; k < TD_CLUSTER; ++k) {
        if (w.cost_atk < cost) break;
        w.cost_atk = __dsub_rn(w.cost_atk, cost);
        ++n;
    }
    w.fail = n == 0 ? TD_FC_COST_SHORTAGE : TD_FC_SUCCESS;
    if (n > w.ecap - w.ne) { w.flags |= 1; n = w.ecap - w.ne; }
    if (w.lane < n) {
        td_enemy_rec &e = w.en()[w.ne + w.lane];
        e.LP = cc.enemy_LP[t][lv];
        e.margin = 0.0;
        e.loc = (uint16_t)start;
        e.type_lv = (uint8_t)(t | (lv << 2));
        e.slowdown = 0;
    }
    w.ne += n;
    gsync(w);
}

// ------------------------------------------------------------------------------------------------
// scripted opponents on the device generator (TDGymBasic.py:81-196, random_agent=True)

// host_cluster != 0xffffffff: the eight types and the road were drawn by the host (td_step_io.opponent_cluster_dev)
template <class W>
__device__ __forceinline__ void opponent_enemy(W &w, int difficulty, unsigned host_cluster)
{
    const DevConfig &cc = w.pp->cfg;
    if (w.atk_cd != 0) return;
    if (difficulty == 0) {                                   // random_enemy_lv0
        long long mine = 0;
        int road;
        if (host_cluster != 0xffffffffu) {
            mine = w.lane < TD_CLUSTER ? (long long)((host_cluster >> (2 * w.lane)) & 3u) : 0ll;
            road = min((int)((host_cluster >> 16) & 3u), w.mh()->num_roads - 1);
        } else {
            for (int k = 0; k < TD_CLUSTER; ++k) { int t = py_randbelow(w, TD_NTYPES + 1); if (w.lane == k) mine = t; }
            road = py_randbelow(w, w.mh()->num_roads);
        }
        summon_cluster(w, road, mine, 0);
    } else {                                                 // random_enemy_lv1
        int t = py_randbelow(w, TD_NTYPES);
        int road = py_randbelow(w, w.mh()->num_roads);
        summon_uniform(w, t, road);
    }
    w.atk_cd = cc.atk_interval;                              // the returned tuple is always truthy
}

template <class W>
__device__ __forceinline__ void opponent_tower(W &w, int difficulty, bool &dirty)
{
    const DevConfig &cc = w.pp->cfg;
    if (w.def_cd != 0) return;
    const int L = w.L();
    if (difficulty == 0) {                                   // random_tower_lv0
        int r = py_randbelow(w, L), c = py_randbelow(w, L), t = py_randbelow(w, TD_NTYPES);
        if (tower_build(w, t, r * L + c, dirty)) w.def_cd = cc.def_interval;
        return;
    }
    int act = py_randbelow(w, 3);                            // random_tower_lv1 / lv2
    if (act == 0) {
        int t = 0;
        if (difficulty == 2) {
            // TDGymBasic.py:216-240: counter the enemy type drawn in proportion to the live enemies
            if (w.ne == 0) return;
            int cnt[TD_NTYPES] = {0, 0, 0, 0};
            for (int base = 0; base < w.ne; base += W::G) {
                const int e = base + w.lane;
                const int ty = e < w.ne ? (w.en()[e].type_lv & 3) : -1;
#pragma unroll
                for (int q = 0; q < TD_NTYPES; ++q) cnt[q] += __popc(gballot(w, ty == q));
            }
            double p = py_random(w);
            int pick = -1, last = 0;
#pragma unroll
            for (int q = 0; q < TD_NTYPES; ++q) {
                if (cnt[q] == 0 || pick >= 0) continue;
                const double ratio = (double)(float)cnt[q] / (double)w.ne;   // float32 counts / np.int64 sum -> f64
                last = q;
                if (p < ratio) pick = q;
                else p = __dsub_rn(p, ratio);
            }
            if (pick < 0) pick = last;
            t = pick == 0 ? 2 : pick == 2 ? 1 : 0;           // [2, 0, 1, 0][type]
            if (py_random(w) < 0.2) t = 3;
        }
        // road cells in row-major order
        uint16_t *list = reinterpret_cast<uint16_t *>(w.scratch());
        int n = 0;
        for (int base = 0; base < w.ncells(); base += W::G) {
            int c = base + w.lane;
            bool on = c < w.ncells() && (w.cells()[c] & 1);
            unsigned b = gballot(w, on);
            TD_CHECK(w, 2 * (n + __popc(b)) <= max(768, w.cells_pad()));
            if (on) list[n + __popc(b & ((1u << w.lane) - 1u))] = (uint16_t)c;
            n += __popc(b);
        }
        gsync(w);
        py_shuffle_u16(w, list, n);
        if (difficulty != 2) t = py_randbelow(w, TD_NTYPES);
        for (int k = 0; k < n; ++k) {
            int di = py_randbelow(w, 25);
            int cell = list[k];
            int r = cell / L + (di / 5 - 2), c = cell % L + (di % 5 - 2);
            if (r < 0 || r >= L || c < 0 || c >= L) continue;
            if (tower_build(w, t, r * L + c, dirty)) { w.def_cd = cc.def_interval; return; }
            if (w.fail == TD_FC_COST_SHORTAGE) return;
        }
    } else {
        if (w.nt == 0) return;
        if (act == 2 && py_random(w) > .01) return;
        int id = py_randbelow(w, w.nt);
        int loc = w.tw()[id].loc;
        bool ok = act == 1 ? tower_lvup(w, loc) : tower_destruct(w, loc, dirty);
        if (ok) w.def_cd = cc.def_interval;
    }
}

// ------------------------------------------------------------------------------------------------
// (c)(d)(e) TDBoard.step, returns the defender reward; kills/leaks for the statistics

struct EnemyRegs {
    double LP, margin;
    int loc, tl, slow, r, c;
    bool valid;
};

template <int NCHUNK, class W>
__device__ __forceinline__ double board_step(W &w, int &kills_out, int &leaks_out)
{
    const DevConfig &cc = w.pp->cfg;
    const int L = w.L(), lane = w.lane;
    double reward = __dadd_rn(0.0, cc.reward_time);
    w.steps += 1;
    const double progress = (double)w.steps / (double)cc.max_steps;

    const int ne = w.ne, nt = w.nt;
    EnemyRegs E[NCHUNK];
    double *keys = reinterpret_cast<double *>(w.scratch());    // [64]
    uint8_t *erow = w.scratch() + 512, *ecol = w.scratch() + 576;  // [64] each

    // ---- load enemies into registers, sort key = dist - margin (TDBoard.py:305)
    bool unsorted = false;
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) {
        int e = lane + W::G * k;
        E[k].valid = e < ne;
        if (E[k].valid) {
            const td_enemy_rec &x = w.en()[e];
            E[k].LP = x.LP; E[k].margin = x.margin; E[k].loc = x.loc; E[k].tl = x.type_lv; E[k].slow = x.slowdown;
            keys[e] = __dsub_rn((double)w.dist()[E[k].loc], E[k].margin);
        }
    }
    gsync(w);
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) {
        int e = lane + W::G * k;
        bool inv = E[k].valid && e > 0 && keys[e - 1] > keys[e];
        unsorted = unsorted || inv;
    }
    unsorted = gany(w, unsorted);
    if (unsorted) {
        // stable rank = #(key smaller) + #(equal key, earlier position)
        int rank[NCHUNK];
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) rank[k] = 0;
        for (int j = 0; j < ne; ++j) {
            double kj = keys[j];
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k) {
                int e = lane + W::G * k;
                if (E[k].valid) { double ke = keys[e]; rank[k] += (kj < ke || (kj == ke && j < e)) ? 1 : 0; }
            }
        }
        gsync(w);
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k)
            if (E[k].valid) {
                TD_CHECK(w, rank[k] >= 0 && rank[k] < ne);
                td_enemy_rec &x = w.en()[rank[k]];
                x.LP = E[k].LP; x.margin = E[k].margin; x.loc = (uint16_t)E[k].loc; x.type_lv = (uint8_t)E[k].tl;
                x.slowdown = (uint8_t)E[k].slow;
            }
        gsync(w);
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) {
            int e = lane + W::G * k;
            if (E[k].valid) {
                const td_enemy_rec &x = w.en()[e];
                E[k].LP = x.LP; E[k].margin = x.margin; E[k].loc = x.loc; E[k].tl = x.type_lv; E[k].slow = x.slowdown;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) {
        int e = lane + W::G * k;
        if (E[k].valid) {
            E[k].r = E[k].loc / L; E[k].c = E[k].loc - E[k].r * L;
            erow[e] = (uint8_t)E[k].r; ecol[e] = (uint8_t)E[k].c;
        }
    }
    gsync(w);

    // ---- towers choose targets: first enemy in list order within Chebyshev range, corpses included
    //      (TDBoard.py:306-312, TDElements.py:72-132).  Positions do not change inside the tower loop, so
    //      every tower's choice is independent: lane = tower.
    uint8_t *fire = w.scratch() + 640, *vict = w.scratch() + 672;     // [32] each
    for (int tt = lane; tt < nt; tt += W::G) {
        td_tower_rec &T = w.tw()[tt];
        const int ty = T.type_lv & 3, lv = T.type_lv >> 2;
        double cd = __dsub_rn(T.cd, 1.0);
        int target = -1, victim = -1;
        if (!(cd > 0.0)) {
            const int rge = cc.tower_range[ty][lv];
            const int tr = T.loc / L, tc = T.loc - tr * L;
            for (int j = 0; j < ne; ++j) {
                int dr = abs((int)erow[j] - tr), dc = abs((int)ecol[j] - tc);
                if (max(dr, dc) <= rge) { target = j; break; }
            }
            if (target >= 0) {
                cd = __dadd_rn(cd, cc.tower_intv[ty][lv]);
                victim = target;
                if (ty == 3) {                                // Frozen: first enemy within splash of the target
                    const int sp = cc.tower_splash[ty][lv];
                    if (sp > 0) {
                        const int r0 = erow[target], c0 = ecol[target];
                        for (int j = 0; j < ne; ++j) {
                            int dr = abs((int)erow[j] - r0), dc = abs((int)ecol[j] - c0);
                            if (max(dr, dc) <= sp) { victim = j; break; }
                        }
                    }
                }
            }
            if (cd < 0.0) cd = 0.0;
        }
        T.cd = cd;
        fire[tt] = (uint8_t)(target < 0 ? 0xff : target);
        vict[tt] = (uint8_t)(victim < 0 ? 0xff : victim);
    }
    gsync(w);

    // ---- damage in tower order: lane = enemy (TDElements.py:19-28)
    bool hit[NCHUNK];
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) hit[k] = false;
    if (ne > 0) {
        double defense[NCHUNK];
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) defense[k] = E[k].valid ? cc.enemy_defense[E[k].tl & 3][E[k].tl >> 2] : 0.0;
        for (int t = 0; t < nt; ++t) {
            const int f = fire[t];
            if (f == 0xff) continue;
            const int tl = w.tw()[t].type_lv, ty = tl & 3, lv = tl >> 2;
            const double atk = cc.tower_attack[ty][lv];
            const double floor_ = __dmul_rn(atk, .05);
            const bool magic = (ty == 1 || ty == 3);
            const int sp = cc.tower_splash[ty][lv];
            const int fr = erow[f], fc = ecol[f], v = vict[t];
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k) {
                int e = lane + W::G * k;
                bool h;
                if (ty == 2) h = E[k].valid && max(abs(E[k].r - fr), abs(E[k].c - fc)) <= sp;
                else if (ty == 3) h = E[k].valid && e == v;
                else h = E[k].valid && e == f;
                if (h) {
                    double dmg;
                    if (magic) dmg = atk;
                    else { dmg = __dsub_rn(atk, defense[k]); if (!(dmg > 0.0)) dmg = 0.0; }
                    if (dmg < floor_) dmg = floor_;
                    E[k].LP = __dsub_rn(E[k].LP, dmg);
                    if (E[k].LP <= 0.0) E[k].LP = 0.0;
                    if (ty == 3) E[k].slow = cc.frozen_time;
                    hit[k] = true;
                }
            }
        }
    }

    // ---- remove the killed, move the rest, remove the leaked (TDBoard.py:313-346)
    int kills = 0, leaks = 0, kept_before = 0;
    const int end = w.mh()->end;
    int newidx[NCHUNK];
    bool keep[NCHUNK];
    gsync(w);
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) {
        bool killed = E[k].valid && hit[k] && !(E[k].LP > 0.0);
        bool leaked = false;
        if (E[k].valid && !killed) {
            const double speed = cc.enemy_speed[E[k].tl & 3][E[k].tl >> 2];
            if (E[k].slow > 0) { E[k].margin = __dadd_rn(E[k].margin, __dmul_rn(speed, cc.frozen_ratio)); E[k].slow -= 1; }
            else E[k].margin = __dadd_rn(E[k].margin, speed);
            while (E[k].margin >= 1.0) {
                E[k].margin = __dsub_rn(E[k].margin, 1.0);
                int d = (w.cells()[E[k].loc] >> 4) & 3;
                E[k].loc += (d == 0) ? 1 : (d == 1) ? -1 : (d == 2) ? L : -L;
                TD_CHECK(w, E[k].loc >= 0 && E[k].loc < w.ncells() && (w.cells()[E[k].loc] & 1));
                if (E[k].loc == end) { leaked = true; break; }
            }
        }
        keep[k] = E[k].valid && !killed && !leaked;
        unsigned bk = gballot(w, killed), bl = gballot(w, leaked), bs = gballot(w, keep[k]);
        kills += __popc(bk);
        leaks += __popc(bl);
        newidx[k] = kept_before + __popc(bs & ((1u << lane) - 1u));
        kept_before += __popc(bs);
    }
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k)
        if (keep[k]) {
            TD_CHECK(w, newidx[k] >= 0 && newidx[k] < w.ecap);
            td_enemy_rec &x = w.en()[newidx[k]];
            x.LP = E[k].LP; x.margin = E[k].margin; x.loc = (uint16_t)E[k].loc; x.type_lv = (uint8_t)E[k].tl;
            x.slowdown = (uint8_t)E[k].slow;
        }
    w.ne = kept_before;
    gsync(w);

    reward = __dadd_rn(reward, __dmul_rn(cc.reward_kill, (double)kills));
    const bool has_base = cc.base_LP >= 0;
    for (int i = 0; i < leaks; ++i) {
        if (has_base && w.base_LP > 0) reward = __dsub_rn(reward, cc.penalty_leak);
        if (has_base) w.base_LP = max(w.base_LP - 1, 0);
    }

    // ---- economy (TDBoard.py:348-353)
    double rate;
    if (progress >= 0.5) rate = cc.rate_final;
    else rate = __dadd_rn(__dmul_rn(cc.rate_init, __dsub_rn(1.0, progress)), __dmul_rn(cc.rate_final, progress));
    double ca = __dadd_rn(w.cost_atk, rate);
    w.cost_atk = cc.max_cost < ca ? cc.max_cost : ca;
    double cd = __dadd_rn(w.cost_def, cc.def_rate);
    w.cost_def = cc.max_cost < cd ? cc.max_cost : cd;

    kills_out = kills;
    leaks_out = leaks;
    return reward;
}

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

template <int NT, class W, class OT>
__device__ __forceinline__ void obs_dense(const W &w, OT *o, int tid)
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
        store_run<C4, NT>(o, 10 * C4, 0.f, tid);
#pragma unroll
        for (int k = 11; k < 14; ++k) store_run<C4, NT>(o, k * C4, pv[k], tid);
        store_run<6 * C4, NT>(o, 15 * C4, 0.f, tid);
#pragma unroll
        for (int k = 21; k < 25; ++k) store_run<C4, NT>(o, k * C4, pv[k], tid);
        store_run<16 * C4, NT>(o, 25 * C4, 0.f, tid);
#pragma unroll
        for (int k = 41; k < 45; ++k) store_run<C4, NT>(o, k * C4, pv[k], tid);
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
}

// Step 3: the sparse one-hots and enemy statistics, 4-byte stores on top of the dense planes (the caller
// orders them after every dense store to this env: __syncwarp for one group, __syncthreads for a CTA sweep).
template <class W, class OT>
__device__ __forceinline__ void obs_sparse(W &w, OT *o)
{
    const DevConfig &cc = w.pp->cfg;
    constexpr int CELLS = W::kCells;
    const int lane = w.lane, cells = CELLS > 0 ? CELLS : w.ncells();
    // ---- enemy statistics per (type, cell) group in list order, float32 (TDBoard.py:355-365, NumPy-2 casts)
    const int ne = w.ne;
    const bool one_pass = W::G == 32 && ne <= 32;                // every enemy has its own lane
    float *ratio = reinterpret_cast<float *>(w.scratch());       // [64], only for the general path
    float mine = 0.f;
    if (one_pass) {
        if (lane < ne) {
            const td_enemy_rec &x = w.en()[lane];
            mine = (float)(x.LP / cc.enemy_LP[x.type_lv & 3][x.type_lv >> 2]);
        }
    } else {
        for (int e = lane; e < ne; e += W::G) {
            const td_enemy_rec &x = w.en()[e];
            ratio[e] = (float)(x.LP / cc.enemy_LP[x.type_lv & 3][x.type_lv >> 2]);
        }
    }
    gsync(w);   // also orders the dense stores above before the sparse stores below
    if (lane == 0) obs_store1(o, (size_t)4 * cells + w.mh()->end, 1.f);
    if (lane < w.mh()->num_roads) obs_store1(o, (size_t)(6 + lane) * cells + w.mh()->start[lane], 1.f);
    for (int t = lane; t < w.nt; t += W::G) {
        const td_tower_rec &T = w.tw()[t];
        obs_store1(o, (size_t)(15 + (T.type_lv >> 2)) * cells + T.loc, 1.f);
        obs_store1(o, (size_t)(17 + (T.type_lv & 3)) * cells + T.loc, 1.f);
    }
    if (one_pass) {
        // lanes of one (cell, type) group find each other with one match instruction; every lane then folds its
        // group's ratios in list order (ascending lane): as many rounds as the largest group has members
        const bool have = lane < ne;
        const int loc = have ? w.en()[lane].loc : 0, ty = have ? (w.en()[lane].type_lv & 3) : 0;
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
    obs_dense<W::G>(w, o, w.lane);
    obs_sparse(w, o);
}

// ------------------------------------------------------------------------------------------------
// kernels

extern __shared__ __align__(16) uint8_t td_smem[];

#ifndef TD_MIN_BLOCKS
#define TD_MIN_BLOCKS 6
#endif
#ifndef TD_MIN_BLOCKS_ATK
#define TD_MIN_BLOCKS_ATK 8      // 10x10 boards: the attacker env is latency-bound (scripted defender): 32 warps per SM at 64 registers
#endif                           // (24 B of spills): 6 / 7 / 8 CTAs per SM = 0.2490 / 0.2440 / 0.2376 ms on B200; the other
                                 // kinds gain nothing or lose (multi-action: +5 % at 7); its in-place-observation variant is
                                 // best at 7 (0.2075 vs 0.2195 ms at 8)

// One env's rules for one step: load the record, apply the actions / scripted opponent, advance the board, emit
// the per-env outputs, auto-reset.  Leaves the updated record in the slice and starts the asynchronous copy of
// the next step's generator words into the slice's word cache (the caller waits for it before store_env).
template <int KIND, bool MULTI, int NCHUNK, bool INC, class W>
__device__ __forceinline__ void env_rules(const StepParams &p, const int env, W &w, uint8_t *rec, bool &dirty)
{
    const DevConfig &cc = p.cfg;
    constexpr int GW = W::G;
    const int lane = w.lane;
    const td_step_io &io = p.io;
    const bool host_opponent = (KIND == TD_KIND_DEF && (io.opponent_dev != nullptr || io.opponent_cluster_dev != nullptr)) ||
                               (KIND == TD_KIND_ATK && io.def_action_dev != nullptr);
    const bool device_opponent = (KIND != TD_KIND_2P) && p.opponent_seeded && p.mt != nullptr && !host_opponent;
    // the record and the inputs are requested together: one round trip
    issue_env_load(w, rec);
    long long in_def = 0;
    int in_opp = 0xff;
    if (KIND != TD_KIND_ATK && !MULTI) in_def = io.def_action_dev[env];
    if (KIND == TD_KIND_ATK && io.def_action_dev != nullptr) in_def = io.def_action_dev[env];     // host-resolved build
    unsigned in_cluster = 0xffffffffu;
    if (KIND == TD_KIND_DEF && io.opponent_cluster_dev != nullptr) in_cluster = io.opponent_cluster_dev[env];
    if (KIND != TD_KIND_DEF) {
        // the attacker's (3, 8) int64 action travels with the record: 12 asynchronous 16-byte copies into the slice,
        // where summon_cluster turns it into the RealAction in place (no registers held across the step)
        const long long *src = reinterpret_cast<const long long *>(io.atk_action_dev) + (size_t)env * TD_ROADS * TD_CLUSTER;
        if ((reinterpret_cast<uintptr_t>(io.atk_action_dev) & 15) == 0) {
            async_copy16(w.act_stage(), src, TD_ROADS * TD_CLUSTER / 2, lane, GW);
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            for (int q = lane; q < TD_ROADS * TD_CLUSTER; q += GW) w.act_stage()[q] = src[q];
        }
    }
    if (KIND == TD_KIND_DEF && io.opponent_dev != nullptr) in_opp = io.opponent_dev[env];
    w.ecap = GW * NCHUNK;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    gsync(w);
    finish_env_load(w, p, rec, device_opponent ? p.mt + (size_t)env * kMtWords : nullptr);
    if (INC) {
        // remember where towers and enemies stand in the observation the caller's buffer still holds
        uint32_t *old = w.old_lists();
        if (lane == 0) { old[0] = (uint32_t)w.nt; old[1] = (uint32_t)w.ne; }
        for (int t = lane; t < w.nt; t += GW) old[4 + t] = w.tw()[t].loc | ((uint32_t)w.tw()[t].type_lv << 16);
        for (int e = lane; e < w.ne; e += GW) old[4 + TD_CAP_TOWERS + e] = w.en()[e].loc | ((uint32_t)(w.en()[e].type_lv & 3) << 16);
    }

    // cooldowns (TDDefense.py:38-39)
    w.atk_cd = max(w.atk_cd - 1, 0);
    w.def_cd = max(w.def_cd - 1, 0);

    long long real_def = 6ll * w.ncells();
    int fail_def = 0;
    bool def_ok = false;
    int fail_atk[TD_ROADS] = {0, 0, 0}, n_fail_atk = 0;

    auto defender = [&]() {
        if (MULTI) {
            decode_multi(w, reinterpret_cast<const long long *>(io.def_action_dev) + (size_t)env * 6 * w.ncells(),
                         io.real_def_dev ? reinterpret_cast<long long *>(io.real_def_dev) + (size_t)env * 6 * w.ncells() : nullptr,
                         dirty);
        } else {
            def_ok = decode_discrete(w, in_def, real_def, fail_def, dirty);
        }
    };
    auto attacker = [&]() {
        if (w.atk_cd == 0) {
            const int nr = w.mh()->num_roads;
            // one instance of the cluster code for all roads (kept rolled: the kernel is instruction-cache bound)
#pragma unroll 1
            for (int i = 0; i < nr; ++i) {
                long long cur = lane < TD_CLUSTER ? w.act_stage()[i * TD_CLUSTER + lane] : (long long)TD_NTYPES;
                int code = 0;
                bool skip = false;
                if (!(KIND == TD_KIND_2P && MULTI))
                    skip = gall(w, lane >= TD_CLUSTER || cur == TD_NTYPES);           // TDAttack.py:39-41
                if (!skip) {
                    const long long before = cur;
                    const bool res = summon_cluster(w, i, cur, 0);
                    if (KIND == TD_KIND_2P) { cur = before; w.atk_cd = cc.atk_interval; }   // tuple truthiness
                    else if (res) w.atk_cd = cc.atk_interval;
                    code = w.fail;
                    if (lane < TD_CLUSTER) w.act_stage()[i * TD_CLUSTER + lane] = cur;          // RealAction
                }
                if (n_fail_atk == 0) fail_atk[0] = code; else if (n_fail_atk == 1) fail_atk[1] = code; else fail_atk[2] = code;
                ++n_fail_atk;
            }
            if (KIND == TD_KIND_2P && MULTI) n_fail_atk = 0;
        }
    };

    if (KIND == TD_KIND_DEF) {
        defender();
        if (io.opponent_dev != nullptr) {
            const int o = in_opp;
            if (o != 0xff && w.atk_cd == 0) {
                summon_uniform(w, o & 3, min((o >> 4) & 3, w.mh()->num_roads - 1));
                w.atk_cd = cc.atk_interval;
            }
        } else if (io.opponent_cluster_dev != nullptr) {
            if (in_cluster != 0xffffffffu) opponent_enemy(w, 0, in_cluster);
        } else if (device_opponent) opponent_enemy(w, p.difficulty, 0xffffffffu);
    } else if (KIND == TD_KIND_ATK) {
        attacker();
        if (io.def_action_dev != nullptr) {                    // random_tower_lv0 resolved by the host (np_random)
            const long long top = (long long)TD_NTYPES * w.ncells();
            if (w.def_cd == 0 && in_def >= 0 && in_def < top) {
                const int t = (int)(in_def / w.ncells()), loc = (int)(in_def - (long long)t * w.ncells());
                if (tower_build(w, t, loc, dirty)) w.def_cd = cc.def_interval;
            }
        } else if (device_opponent) opponent_tower(w, p.difficulty, dirty);
    } else {
        attacker();
        defender();
    }
    auto refill = [&]() {
        // Generator words for the next step: copied global -> shared straight into the slice's word cache (this
        // step's draws are done with it), asynchronously, so that no register waits for them behind the observation.
        if (w.mt != nullptr && w.cn - w.ck < (w.rng_words() >> 1)) {
            w.cache_dirty = true;
            w.cache_raw = true;
            w.ck = 0;
            w.cn = min(w.rng_words(), max(kMtWords - w.mt_pos, 0));
            constexpr int kRefill = (W::kRngWords + GW - 1) / GW;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(w.rng_cache());
#pragma unroll
            for (int q = 0; q < kRefill; ++q) {
                const int k = lane + GW * q;
                if (k < w.cn)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * k), "l"(w.mt + w.mt_pos + k) : "memory");
                else if (k < W::kRngWords)
                    const_cast<uint32_t *>(w.rng_cache())[k] = 0u;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (INC) refill();          // in-place observation kernels write the record back first (see td_step_kernel)
    gsync(w);

    int kills, leaks;
    double reward = board_step<NCHUNK>(w, kills, leaks);
    if (KIND == TD_KIND_ATK) reward = -reward;

    const bool has_base = cc.base_LP >= 0;
    const bool done = (has_base && w.base_LP <= 0) || w.steps >= cc.max_steps;
    const bool def_wins = !has_base || w.base_LP > 0;
    const bool atk_wins = !has_base || w.base_LP <= 0;
    const bool my_win = KIND == TD_KIND_ATK ? atk_wins : def_wins;

    if (lane == 0) {
        if (io.reward_dev) io.reward_dev[env] = reward;
        if (io.done_dev) io.done_dev[env] = done ? 1 : 0;
        if (io.win_dev) io.win_dev[env] = done ? (my_win ? 1 : 0) : -1;
        if (io.allow_next_dev) io.allow_next_dev[env] = (uint8_t)((w.def_cd <= 1 ? 1 : 0) | (w.atk_cd <= 1 ? 2 : 0));
        if (!MULTI && KIND != TD_KIND_ATK) {
            if (io.real_def_dev) io.real_def_dev[env] = real_def;
            if (io.fail_def_dev) io.fail_def_dev[env] = fail_def;
        } else if (io.fail_def_dev) io.fail_def_dev[env] = 0;
        if (KIND != TD_KIND_DEF && io.fail_atk_dev) {
            int4 f = make_int4(n_fail_atk, fail_atk[0], fail_atk[1], fail_atk[2]);
            reinterpret_cast<int4 *>(io.fail_atk_dev)[env] = f;
        }
        td_env_header *h = w.hdr();
        h->ep_return = __dadd_rn(h->ep_return, reward);
        h->ep_kills = (uint16_t)(h->ep_kills + kills);
        h->ep_leaks = (uint16_t)(h->ep_leaks + leaks);
        if (done) {
            EnvStats &s = p.stats[env];
            s.return_sum = __dadd_rn(s.return_sum, h->ep_return);
            s.episodes += 1;
            s.wins += my_win ? 1 : 0;
            s.length_sum += (uint32_t)w.steps;
            s.kills += h->ep_kills;
            s.leaks += h->ep_leaks;
        }
        if (w.flags) p.stats[env].flags |= (uint32_t)w.flags;
    }
    if (KIND != TD_KIND_DEF && io.real_atk_dev) {
        for (int q = lane; q < TD_ROADS * TD_CLUSTER; q += GW)
            io.real_atk_dev[(size_t)env * TD_ROADS * TD_CLUSTER + q] = w.act_stage()[q];
    }
    if (io.packed_out_dev != nullptr) {
        // every small output of the env in one record and one coalesced store (the buffer may be host memory)
        constexpr int kStride = KIND == TD_KIND_DEF ? 32 : 256;
        int4 *po = reinterpret_cast<int4 *>(static_cast<uint8_t *>(io.packed_out_dev) + (size_t)env * kStride);
        const long long rd = (MULTI || KIND == TD_KIND_ATK) ? 0ll : real_def;
        const long long rbits = __double_as_longlong(reward);
        const int win_v = done ? (my_win ? 1 : 0) : -1;
        const unsigned fl = (done ? 1u : 0u) | ((unsigned)(win_v & 0xff) << 8) |
                            ((unsigned)((w.def_cd <= 1 ? 1 : 0) | (w.atk_cd <= 1 ? 2 : 0)) << 16);
        int4 v = make_int4(0, 0, 0, 0);
        if (lane == 0) v = make_int4((int)rbits, (int)(rbits >> 32), (int)rd, (int)(rd >> 32));
        if (lane == 1) v = make_int4((MULTI || KIND == TD_KIND_ATK) ? 0 : fail_def, (int)fl, 0, 0);
        if (KIND != TD_KIND_DEF) {
            if (lane == 2) v = make_int4(n_fail_atk, fail_atk[0], fail_atk[1], fail_atk[2]);
            // lanes 4..15 carry real_atk[2q], real_atk[2q + 1] (q = lane - 4), straight from the slice
            if (lane >= 4 && lane < 16) v = reinterpret_cast<const int4 *>(w.act_stage())[lane - 4];
        }
        if (lane < kStride / 16) po[lane] = v;
    }
    (void)def_ok;
    gsync(w);

    if (__builtin_expect(done && io.auto_reset, 0)) {
        int next = (w.hdr()->map_id + p.map_stride) % p.n_maps;
        if (lane == 0) w.hdr()->episode += 1;
        reset_env(w, p, next, true);
        dirty = true;
    }
    if (!INC) refill();
}

// INC: the observation is an in-place update of the previous one (td_step_io.obs_incremental, vouched for by the
// engine); a separate instantiation, so that the full-write kernels carry none of its code.
template <int KIND, bool MULTI, int CELLS, int NCHUNK, int GW, bool INC, class OT = float>
__global__ void __launch_bounds__(kWarpsPerCta * 32, (KIND == TD_KIND_ATK && CELLS == 100) ? (INC ? 7 : TD_MIN_BLOCKS_ATK) : TD_MIN_BLOCKS) td_step_kernel(const __grid_constant__ StepParams p)
{
    // one group of GW lanes per game instance (GW = 16: two instances share a warp)
    const int group = threadIdx.x / GW, lane = threadIdx.x & (GW - 1);
    const int env = p.env_begin + blockIdx.x * (kWarpsPerCta * 32 / GW) + group;
    if (env >= p.n_envs) return;
    constexpr int RC = KIND == TD_KIND_ATK ? kRngCacheAtk : kRngCacheDef;
    Ctx<CELLS, GW, RC> w;
    ctx_bind(w, td_smem + (size_t)group * p.smem_per_warp, p);               // [record | scratch] per instance
    uint8_t *rec = p.records + (size_t)env * w.record_bytes();
    bool dirty = false;
    env_rules<KIND, MULTI, NCHUNK, INC>(p, env, w, rec, dirty);
    // The header scalars go back to the slice before the observation is written: their registers are free
    // during the store phase (a spilled one cost a local-memory reload behind 18 KB of stores: 8 % of the step).
    push_header(w);
    // In-place observation kernels write the record back before the observation (the generator words were
    // requested before board_step): 64 instead of 79 registers, def-small 0.167 -> 0.161 ms.  The full-write
    // kernels keep the record for last: 0.2175 vs 0.2206 ms on def-small.
    if (INC) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (w.cache_raw) { gsync(w); temper_cache(w); }
        store_env<true>(w, p, rec, dirty, false);
    }
    if (p.io.obs_dev) {
        OT *o = reinterpret_cast<OT *>(p.io.obs_dev) + (size_t)env * TD_NCHANNELS * w.ncells();
        if constexpr (INC && CELLS > 0) {
            obs_prepare(w);
            if (!w.static_dirty && (reinterpret_cast<uintptr_t>(o) & 15) == 0) obs_dense_incremental(w, o);
            else obs_dense<GW>(w, o, w.lane);              // envs restarted inside this step get all 45 planes
            obs_sparse(w, o);
        } else {
            write_obs(w, o);
        }
    }
    if (!INC) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // the next steps' generator words are in the slice
        if (w.cache_raw) { gsync(w); temper_cache(w); }
        store_env<(KIND == TD_KIND_ATK)>(w, p, rec, dirty, false);
    }
}

// reset (mask / explicit map ids) and observation-only kernels
__global__ void __launch_bounds__(kWarpsPerCta * 32)
td_reset_kernel(const __grid_constant__ StepParams p, const uint8_t *mask, const int32_t *map_ids, float *obs)
{
    const int warp = threadIdx.x >> 5;
    const int env = blockIdx.x * kWarpsPerCta + warp;
    if (env >= p.n_envs) return;
    if (mask && !mask[env]) return;
    Ctx<0, 32> w;
    ctx_bind(w, td_smem + (size_t)warp * p.smem_per_warp, p);
    uint8_t *rec = p.records + (size_t)env * p.record_bytes;
    if (w.lane < (w.hdr_bytes() >> 4)) reinterpret_cast<int4 *>(w.hdr())[w.lane] = reinterpret_cast<const int4 *>(rec)[w.lane];
    gsync(w);
    pull_header(w);
    int id = map_ids ? map_ids[env] : env % p.n_maps;
    id = ((id % p.n_maps) + p.n_maps) % p.n_maps;
    reset_env(w, p, id, true);
    if (obs) write_obs(w, obs + (size_t)env * TD_NCHANNELS * w.ncells());
    gsync(w);
    store_env(w, p, rec, true);
}

template <int CELLS, class OT = float>
__global__ void __launch_bounds__(kWarpsPerCta * 32) td_observe_kernel(const __grid_constant__ StepParams p, OT *obs)
{
    const int warp = threadIdx.x >> 5;
    const int env = blockIdx.x * kWarpsPerCta + warp;
    if (env >= p.n_envs) return;
    Ctx<CELLS, 32> w;
    ctx_bind(w, td_smem + (size_t)warp * p.smem_per_warp, p);
    load_env(w, p, p.records + (size_t)env * p.record_bytes);
    write_obs(w, obs + (size_t)env * TD_NCHANNELS * w.ncells());
}

// deterministic reduction of the per-env statistics: one block, fixed order
__global__ void __launch_bounds__(256) td_stats_kernel(const EnvStats *s, int n, long long steps, td_stats *out)
{
    __shared__ double r[256];
    __shared__ long long acc[256][6];
    const int t = threadIdx.x;
    double rs = 0.0;
    long long a[6] = {0, 0, 0, 0, 0, 0};
    for (int i = t; i < n; i += 256) {
        rs += s[i].return_sum;
        a[0] += s[i].episodes; a[1] += s[i].length_sum; a[2] += s[i].wins;
        a[3] += s[i].kills; a[4] += s[i].leaks; a[5] += s[i].flags ? 1 : 0;
    }
    r[t] = rs;
    for (int k = 0; k < 6; ++k) acc[t][k] = a[k];
    __syncthreads();
    for (int stride = 128; stride > 0; stride >>= 1) {
        if (t < stride) {
            r[t] += r[t + stride];
            for (int k = 0; k < 6; ++k) acc[t][k] += acc[t + stride][k];
        }
        __syncthreads();
    }
    if (t == 0) {
        out->return_sum = r[0];
        out->episodes = acc[0][0]; out->length_sum = acc[0][1]; out->wins = acc[0][2];
        out->kills = acc[0][3]; out->leaks = acc[0][4]; out->overflow_envs = acc[0][5];
        out->steps = steps;
    }
}

} // namespace td
