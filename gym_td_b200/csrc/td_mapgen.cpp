// td_mapgen.cpp -- host-side road generator for the B200 gym-TD simulator.
//
// Produces, for a NumPy-legacy MT19937 stream RandomState(seed), exactly the map the reference
// builds in TDGymBasic.reset -> TDBoard.__init__ -> TDRoadGen.create_road_v2
// (gym_TD/envs/TDGymBasic.py:42-51, TDBoard.py:31-59, TDRoadGen.py:4-199), consuming the
// stream draw for draw.  Written from the behaviour of that generator, for many maps in
// parallel on host threads; parity is checked map-for-map in tests/test_mapgen.py.
//
// Seed-skip rule: the reference generator raises or never terminates for a few
// percent of 10x10 seeds.  A seed is invalid when the generator would raise, or would make more
// than `budget` randint() calls; invalid seeds are skipped by the caller (s <- s + 1).
#include "../../include/td_b200.h"

#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

namespace {

struct InvalidSeed {};

// MT19937 as seeded by numpy.random.RandomState(int): init_genrand(seed)
class LegacyStream {
public:
    LegacyStream(const uint32_t *state625, int budget) : calls_(0), budget_(budget)
    {
        memcpy(mt_, state625, sizeof(mt_));
        pos_ = (int)state625[624];
        if (pos_ < 0 || pos_ > 624) pos_ = 624;
    }
    void save(uint32_t *state625) const
    {
        memcpy(state625, mt_, sizeof(mt_));
        state625[624] = (uint32_t)pos_;
    }
    LegacyStream(uint32_t seed, int budget) : pos_(624), calls_(0), budget_(budget)
    {
        mt_[0] = seed;
        for (int i = 1; i < 624; ++i)
            mt_[i] = 1812433253u * (mt_[i - 1] ^ (mt_[i - 1] >> 30)) + (uint32_t)i;
    }
    // RandomState.randint(low, high): masked rejection on 32-bit words; a zero-width range is free
    int randint(int low, int high)
    {
        if (++calls_ > budget_) throw InvalidSeed();
        if (low >= high) throw InvalidSeed();          // numpy: ValueError("low >= high")
        uint32_t range = (uint32_t)(high - 1 - low);
        if (range == 0) return low;
        uint32_t mask = range;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next32() & mask; } while (v > range);
        return low + (int)v;
    }
    int calls() const { return calls_; }

private:
    uint32_t next32()
    {
        if (pos_ >= 624) twist();
        uint32_t y = mt_[pos_++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    void twist()
    {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt_[k] & 0x80000000u) | (mt_[(k + 1) % 624] & 0x7fffffffu);
            mt_[k] = mt_[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        pos_ = 0;
    }
    uint32_t mt_[624];
    int pos_, calls_, budget_;
};

typedef std::pair<int, int> Cell;          // (row, col)
typedef std::vector<Cell> Road;

class RoadBuilder {
public:
    RoadBuilder(LegacyStream &rng, int L) : rng_(rng), L_(L), field_((size_t)L * L, 0), turn_((size_t)L * L, 0) {}

    // TDRoadGen.py:4-199
    std::vector<Road> build(int num_roads)
    {
        const int L = L_;
        const int lo = L / 3, hi = (L * 2 + 2) / 3;
        Cell center;
        center.first = rng_.randint(lo, hi);
        center.second = rng_.randint(lo, hi);
        at(field_, center) = 1;
        const int dir0 = rng_.randint(0, 4);

        Road to_end, to_start;
        for (;;) {                                   // centre -> end, shorter than L cells
            bool ok = grow(center, dir0, to_end);
            if (ok && (int)to_end.size() < L) break;
            erase(to_end);
        }
        for (;;) {                                   // centre -> start, far enough from the end
            bool ok = grow(center, (dir0 + 2) % 4, to_start);
            if (ok) {
                if ((int)(to_end.size() + to_start.size()) + 1 >= 2 * L) ok = false;
                else if (manhattan(to_start.back(), to_end.back()) < L * 3 / 4) ok = false;
            }
            if (ok) break;
            erase(to_start);
        }
        Road main_road(to_start.rbegin(), to_start.rend());
        main_road.push_back(center);
        main_road.insert(main_road.end(), to_end.begin(), to_end.end());

        // junction candidates: straight stretches of the main road (TDRoadGen.py:162-170)
        std::vector<int> junction;
        for (size_t i = 0; i < main_road.size();) {
            if (!at(turn_, main_road[i])) {
                if (i + 1 < main_road.size() && !at(turn_, main_road[i + 1])) junction.push_back((int)i);
                ++i;
            } else i += 2;
        }

        std::vector<Road> roads;
        roads.push_back(main_road);
        const int nj = (int)junction.size();
        for (int k = 1; k < num_roads; ++k) {
            Road branch;
            int at_index = 0;
            for (;;) {
                int pick = rng_.randint(nj * 2 / 5, nj * 4 / 5);
                int dir = rng_.randint(0, 4);
                at_index = junction[(size_t)pick];
                bool ok = grow(main_road[(size_t)at_index], dir, branch);
                if (ok) {
                    if ((int)branch.size() + (int)main_road.size() - at_index >= 2 * L) ok = false;
                    else {
                        if (branch.empty()) throw InvalidSeed();   // reference: new_road[-1] IndexError
                        if (manhattan(branch.back(), main_road.back()) < L * 3 / 4) ok = false;
                    }
                }
                if (ok) break;
                erase(branch);
            }
            Road full(branch.rbegin(), branch.rend());
            full.insert(full.end(), main_road.begin() + at_index, main_road.end());
            roads.push_back(full);
        }
        return roads;
    }

private:
    static int manhattan(const Cell &a, const Cell &b) { return std::abs(a.first - b.first) + std::abs(a.second - b.second); }
    uint8_t &at(std::vector<uint8_t> &g, const Cell &c) { return g[(size_t)c.first * L_ + c.second]; }
    bool inner(const Cell &c) const { return c.first > 0 && c.first < L_ - 1 && c.second > 0 && c.second < L_ - 1; }
    void erase(const Road &r)
    {
        for (size_t i = 0; i < r.size(); ++i) { at(field_, r[i]) = 0; at(turn_, r[i]) = 0; }
    }

    // Lay up to `n` cells from `pos` along direction d.  Returns true when an occupied cell blocked the way.
    // `moved` reports whether at least one cell was laid (the reference clears its cross flag only then).
    bool lay(Cell &pos, int d, int n, Road &road, bool &moved)
    {
        static const int dr[4] = {1, 0, -1, 0}, dc[4] = {0, -1, 0, 1};   // TDRoadGen.py:15
        for (int s = 0; s < n; ++s) {
            Cell nxt(pos.first + dr[d], pos.second + dc[d]);
            if (at(field_, nxt) != 0) return true;
            pos = nxt;
            road.push_back(pos);
            at(field_, pos) = 1;
            moved = true;
            if (!inner(pos)) return false;
        }
        return false;
    }

    // TDRoadGen.py:31-119: random walk of straight pieces and S-turns until the border is reached.
    bool grow(const Cell &start, int dir, Road &road)
    {
        static const int dr[4] = {1, 0, -1, 0}, dc[4] = {0, -1, 0, 1};
        road.clear();
        Cell pos = start;
        int pending_turn = 0;            // 0 = none, else the side (-1/+1) of the next turn
        int loops = 0;
        while (inner(pos) && loops < 100) {
            ++loops;
            const int shape = rng_.randint(0, 2);
            const int seg = rng_.randint(L_ * 3 / 20, L_ / 4);
            bool blocked, moved = false;
            if (shape <= 0) {
                blocked = lay(pos, dir, seg * 2, road, moved);
            } else {
                blocked = lay(pos, dir, seg, road, moved);
                if (!inner(pos)) break;
                int side;
                if (pending_turn != 0) { side = pending_turn; pending_turn = 0; }
                else { side = rng_.randint(0, 2) * 2 - 1; pending_turn = -side; }
                at(turn_, pos) = 1;
                dir = (dir + 4 + side) % 4;
                bool moved2 = false;
                bool blocked2 = lay(pos, dir, seg, road, moved2);
                // the reference resets `cross` after every laid cell of the second leg only
                if (blocked2) blocked = true;
                else if (moved2) blocked = false;
            }
            if (blocked) {
                int open[4], n_open = 0;
                for (int d = 0; d < 4; ++d) {
                    Cell nb(pos.first + dr[d], pos.second + dc[d]);
                    if (at(field_, nb) == 0) open[n_open++] = d;
                }
                if (n_open == 0) return false;
                dir = open[rng_.randint(0, n_open)];
                pending_turn = 0;
                at(turn_, pos) = 1;
            }
        }
        return loops < 100;
    }

    LegacyStream &rng_;
    int L_;
    std::vector<uint8_t> field_, turn_;
};

// TDBoard.py:35-59: planes from the road lists
void fill_map(const std::vector<Road> &roads, int L, td_map *m)
{
    m->map_size = L;
    m->num_roads = (int)roads.size();
    memset(m->cells, 0, sizeof(m->cells));
    memset(m->dist, 0, sizeof(m->dist));
    for (int i = 0; i < TD_ROADS; ++i) m->start[i] = 0;
    for (size_t i = 0; i < roads.size(); ++i) {
        const Road &rd = roads[i];
        m->start[i] = rd.front().first * L + rd.front().second;
        if (i == 0) m->end = rd.back().first * L + rd.back().second;
        for (size_t k = 0; k < rd.size(); ++k) {
            int c = rd[k].first * L + rd[k].second;
            m->cells[c] |= (uint8_t)(1u | (2u << i));
            if (k > 0) {
                int prev = rd[k - 1].first * L + rd[k - 1].second;
                int drow = rd[k].first - rd[k - 1].first, dcol = rd[k].second - rd[k - 1].second;
                int d = drow == 0 ? (dcol == 1 ? 0 : 1) : (drow == 1 ? 2 : 3);
                m->cells[prev] = (uint8_t)((m->cells[prev] & 0x0f) | (d << 4));
            }
        }
        int dist = 0;
        for (size_t k = rd.size(); k-- > 0; ++dist) m->dist[rd[k].first * L + rd[k].second] = (uint8_t)dist;
    }
    int maxd = 0;
    for (int c = 0; c < L * L; ++c) if (m->dist[c] > maxd) maxd = m->dist[c];
    m->max_dist = maxd;
}

int generate_from(LegacyStream &rng, int L, int num_roads, td_map *out)
{
    try {
        if (num_roads <= 0) num_roads = rng.randint(1, TD_ROADS + 1);   // TDGymBasic.py:42
        RoadBuilder b(rng, L);
        std::vector<Road> roads = b.build(num_roads);
        for (size_t i = 0; i < roads.size(); ++i)
            if (roads[i].empty() || roads[i].size() > 255) return 0;
        fill_map(roads, L, out);
        out->n_randint = rng.calls();
        return 1;
    } catch (const InvalidSeed &) {
        out->n_randint = rng.calls();
        return 0;
    } catch (...) {
        return TD_E_ALLOC;
    }
}

int generate_one(uint32_t seed, int L, int num_roads, int budget, td_map *out)
{
    if (L < 4 || L > TD_MAX_L || num_roads > TD_ROADS || !out) return TD_E_INVALID;
    LegacyStream rng(seed, budget > 0 ? budget : 100000);
    return generate_from(rng, L, num_roads, out);
}

} // namespace

extern "C" int td_mapgen(uint32_t seed, int map_size, int num_roads, int budget, td_map *out)
{
    return generate_one(seed, map_size, num_roads, budget, out);
}

extern "C" int td_mapgen_stream(uint32_t *state625, int map_size, int num_roads, int budget, td_map *out)
{
    if (!state625 || map_size < 4 || map_size > TD_MAX_L || num_roads > TD_ROADS || !out) return TD_E_INVALID;
    LegacyStream rng(state625, budget > 0 ? budget : 100000);
    int rc = generate_from(rng, map_size, num_roads, out);
    rng.save(state625);
    return rc;
}

extern "C" int td_mapgen_batch(uint32_t *seeds, int n, int map_size, int num_roads, int budget,
                               int skip_invalid, int threads, td_map *out, int32_t *valid_out)
{
    if (!seeds || !out || n < 0) return TD_E_INVALID;
    if (threads < 1) threads = 1;
    if (threads > n) threads = n > 0 ? n : 1;
    std::vector<int> status((size_t)threads, 0);
    auto work = [&](int tid) {
        for (int i = tid; i < n; i += threads) {
            int v;
            for (;;) {
                v = generate_one(seeds[i], map_size, num_roads, budget, &out[i]);
                if (v != 0 || !skip_invalid) break;
                seeds[i] += 1;
            }
            if (v < 0) { status[(size_t)tid] = v; return; }
            if (valid_out) valid_out[i] = v;
        }
    };
    try {
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
        work(0);
        for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
    } catch (...) {
        return TD_E_ALLOC;
    }
    for (int t = 0; t < threads; ++t) if (status[(size_t)t] < 0) return status[(size_t)t];
    return TD_OK;
}
