// TEST INFRASTRUCTURE ONLY -- host emulation of the CUDA forward kernel.
//
// Compiles bayesair_b200/csrc/ba_sim.cuh (the very code the sm_100a kernel runs) with
// g++ and executes the two phases of every tick lane by lane on the CPU, in a
// SHUFFLED lane order per tick so that the cross-lane protocol (atomic bag append +
// per-destination re-ordering) is exercised under arbitrary interleavings.  It lets
// the build container (which has no GPU) check the kernel logic against the oracle
// before any GPU time is spent.  It is not part of the product: nothing in
// bayesair_b200/ loads it, and the product has no CPU path.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../bayesair_b200/csrc/ba_plan_build.h"
#include "../../bayesair_b200/csrc/ba_sim.cuh"
#include "../../bayesair_b200/csrc/ba_scan2.cuh"
#include <chrono>
#include <thread>

struct EmuPlan { BaHostPlan host; };

extern "C" int ba_emu_plan_create(const BaDayTable *days, int n_days, int n_airports, float dt,
                                  float max_t, int canc, int max_ticks, EmuPlan **out)
{
    EmuPlan *p = new EmuPlan();
    std::string err;
    int rc = ba_build_host_plan(days, n_days, n_airports, dt, max_t, canc, max_ticks, &p->host, &err);
    if (rc != 0) { fprintf(stderr, "emu plan: %s\n", err.c_str()); delete p; return rc; }
    *out = p;
    return 0;
}
extern "C" void ba_emu_plan_destroy(EmuPlan *p) { delete p; }
extern "C" void ba_emu_plan_info(const EmuPlan *p, int *out /* D,N,Fmax,R,C,stride */)
{
    const BaPlanView &v = p->host.view;
    out[0] = v.D; out[1] = v.N; out[2] = v.Fmax; out[3] = v.R; out[4] = v.C; out[5] = v.grad_stride;
}
extern "C" void ba_emu_plan_routes(const EmuPlan *p, int32_t *o, int32_t *d)
{
    memcpy(o, p->host.route_origin.data(), p->host.route_origin.size() * 4);
    memcpy(d, p->host.route_dest.data(), p->host.route_dest.size() * 4);
}

template <bool EXACT, bool PREDICT = false, bool TRK = true>
static void run_item(const BaPlanView &plan, const BaDayView &day, int p, int d, int P,
                     const float *lat_airport, const float *lat_route, const float *scales,
                     const float *noise, uint64_t seed, int value_mode, int want_grad,
                     int replay_waits, unsigned shuffle_seed, float *logp, uint32_t *status, int32_t *ticks,
                     float *tape, int32_t *pop_tick, const float *noise2 = nullptr,
                     float *sim = nullptr)
{
    (void)P;
    const int N = plan.N, C = plan.C, R = plan.R, D = plan.D;
    std::vector<uint32_t> state((size_t)BA_NWORDS_STATE_TRACK * N + 8, 0);
    std::vector<uint32_t> pool(ba_shared_list_words(day.q_total_s) + 2 * BA_DSTAGE_WORDS * (plan.max_dslice + 2) + 8, 0);
    std::vector<uint32_t> fixed(BA_SHARED_FIXED_WORDS + 8, 0);
    std::vector<uint32_t> gs(ba_scratch_words(day, true, plan.max_ticks, plan.max_dslice) + 8, 0);
    uint32_t ctaw[BA_CTA_WORDS] = {0, 0, 0, 0};
    BaCtx c;
    memset(&c, 0, sizeof(c));
    ba_carve_state(c, (uint32_t *)(((uintptr_t)state.data() + 7) & ~(uintptr_t)7), N, plan.any_track != 0);
    c.sigma = scales[0]; c.v_t = scales[1]; c.v_u = scales[2];
    c.inv_sigma = 1.0f / c.sigma; c.inv_sigma2 = c.inv_sigma * c.inv_sigma;
    c.inv_sigma3 = c.inv_sigma2 * c.inv_sigma; c.log_sigma = logf(c.sigma);
    c.inv_vt = 1.0f / c.v_t; c.inv_vu = 1.0f / c.v_u;
    c.c0_lp = ba_relaxed_bernoulli_lp(0.0f, 0.0f, nullptr);
    c.c1_lp = ba_relaxed_bernoulli_lp(0.0f, 1.0f, nullptr);
    c.value_mode = value_mode; c.want_grad = want_grad && tape;
    c.cancel_mode = plan.include_cancellations; c.max_t = plan.max_t; c.seed = seed;
    c.replay_always = replay_waits;
    c.route_base = C * N;
    c.tick_time = plan.tick_time; c.cta = ctaw;
    const size_t pd = (size_t)p * D + d;
    c.N = N; c.F = day.F; c.flights = day.flights; c.route = day.route; c.ap = day.airports;
    c.dep_sched = day.dep_sched; c.dep_word = day.dep_word; c.pos_dep = day.pos_dep;
    c.pos_arr = day.pos_arr; c.rt_gid = day.rt_gid; c.rt_off = day.rt_off;
    c.rt_flt = day.rt_flt; c.Rd = day.Rd;
    c.lat_route = lat_route + (size_t)p * R;
    c.noise = noise ? noise + pd * (size_t)plan.Fmax * 4 : nullptr;
    c.noise2 = (PREDICT && noise2) ? noise2 + pd * (size_t)plan.Fmax * 4 : nullptr;
    c.sim = PREDICT ? sim + pd * (size_t)plan.Fmax * 3 : nullptr;
    c.particle = p; c.dayidx = d;
    c.grow = c.want_grad ? tape + pd * (size_t)plan.grad_stride : nullptr;
    c.pop_tick = pop_tick ? pop_tick + pd * (size_t)plan.Fmax * 2 : nullptr;
    ba_carve_lists<EXACT>(c, day, plan.max_ticks, plan.max_dslice,
                          (uint32_t *)(((uintptr_t)fixed.data() + 15) & ~(uintptr_t)15), pool.data(), gs.data());
    for (uint32_t k = 0; k < c.cal_cap; ++k) c.cal_cnt[k] = 0;

    const float *la = lat_airport + (size_t)p * C * N;
    for (int a = 0; a < N; ++a) ba_init_airport<EXACT, TRK>(c, a, la, N);
    if (c.grow) for (int i = 0; i < plan.grad_stride; ++i) c.grow[i] = 0.0f;
    if (c.pop_tick) for (int i = 0; i < 2 * plan.Fmax; ++i) c.pop_tick[i] = -1;

    // one accumulator per lane, reduced in lane order (as the kernel does per warp)
    std::vector<BaAcc> acc(N, BaAcc{0, 0, 0, 0, 0});
    std::mt19937 rng(shuffle_seed + 977u * (unsigned)pd);
    std::vector<int> forder(day.F);
    for (int f = 0; f < day.F; ++f) forder[f] = f;
    if (shuffle_seed) std::shuffle(forder.begin(), forder.end(), rng);
    for (int f : forder) ba_prologue_flight<PREDICT, TRK>(c, acc[f % N], f);
    if (!PREDICT) {
        ba_calendar_scan(c, -1);
        if (shuffle_seed) std::shuffle(forder.begin(), forder.end(), rng);
        for (int f : forder) ba_calendar_scatter(c, f);
    }
    std::vector<int> order(N);
    for (int i = 0; i < N; ++i) order[i] = i;
    float t = 0.0f;
    int tick = day.first_tick - 1;
    // staging runs ahead of the scan exactly as in the kernel (ba_kernels.cu)
    auto stage_departures = [&](int k) {
        if (k > day.last_ready_tick) return;
        std::vector<uint32_t> sl;
        for (int i = day.dcal_off[k]; i < day.dcal_off[k + 1]; ++i) sl.push_back((uint32_t)i);
        if (shuffle_seed) std::shuffle(sl.begin(), sl.end(), rng);
        for (uint32_t i : sl)
            ba_stage_departure<true>(c, acc[i % N], i, (uint32_t)day.dcal_off[k], (k - 1) & 1);
    };
    bool busy = day.F > 0;
    uint32_t st = 0;
    if (!EXACT && !PREDICT) {
        // single-barrier scan, exactly as in the kernel (ba_kernels.cu).  Asynchronous copies
        // either land at once (odd shuffle seeds) or only when the tick's barrier is reached,
        // in shuffled order (even seeds): the two extremes of cp.async.
        const uint32_t cap = (uint32_t)BA_DUE_CAP_S;
        std::vector<BaEmuCopy> pending;
        ba_emu_defer = (shuffle_seed % 2 == 0) ? &pending : nullptr;
        auto land = [&]() {
            if (shuffle_seed) std::shuffle(pending.begin(), pending.end(), rng);
            for (const BaEmuCopy &cp : pending) memcpy(cp.dst, cp.src, 4 * (size_t)cp.words);
            pending.clear();
        };
        auto slice_lo = [&](int k) { return k >= 1 && (uint32_t)k < c.cal_cap ? c.cal_cnt[k - 1] : 0u; };
        auto slice_n = [&](int k) {
            if (k < 1 || (uint32_t)k >= c.cal_cap) return 0u;
            const uint32_t n = c.cal_cnt[k] - c.cal_cnt[k - 1];
            return n > cap ? cap : n;
        };
        auto stage_words = [&](int k) {
            if (k < 1 || (uint32_t)k >= c.cal_cap) return;
            const uint32_t lo = c.cal_cnt[k - 1], n = c.cal_cnt[k] - lo;
            if (n > cap) acc[0].status |= BA_S_OVERFLOW;
            std::vector<uint32_t> sl;
            for (uint32_t i = lo; i < lo + (n > cap ? cap : n); ++i) sl.push_back(i);
            if (shuffle_seed) std::shuffle(sl.begin(), sl.end(), rng);
            for (uint32_t i : sl) ba_stage_word(c, i, lo, k & 1);
        };
        auto stage_departures_async = [&](int k) {          // slice k -> buffer (k - 1) & 1
            if (k > day.last_ready_tick) return;
            std::vector<uint32_t> sl;
            for (int i = day.dcal_off[k]; i < day.dcal_off[k + 1]; ++i) sl.push_back((uint32_t)i);
            if (shuffle_seed) std::shuffle(sl.begin(), sl.end(), rng);
            for (uint32_t i : sl)
                ba_stage_departure<false>(c, acc[i % N], i, (uint32_t)day.dcal_off[k], (k - 1) & 1);
        };
        stage_departures_async(tick + 1);
        stage_words(tick + 1);
        land();
        while (busy) {
            if (tick >= plan.max_ticks) { st |= BA_S_TICK_CAP; break; }
            ++tick;
            t = c.tick_time[tick];
            const int prev = (tick - 1) & 1, cur = tick & 1;
            const uint32_t n_cur = slice_n(tick), n_prev = slice_n(tick - 1);
            uint32_t n_late = c.cta[8 + (tick - 1) % 3];
            if (n_late > cap - n_prev) n_late = cap - n_prev;
            c.cta[8 + (tick + 1) % 3] = 0;
            {
                const uint32_t lo = slice_lo(tick);
                std::vector<uint32_t> sl;
                for (uint32_t i = 0; i < n_cur; ++i) sl.push_back(i);
                if (shuffle_seed) std::shuffle(sl.begin(), sl.end(), rng);
                for (uint32_t i : sl) ba_complete_arrival(c, i, lo, cur);
            }
            stage_words(tick + 1);
            stage_departures_async(tick + 1);
            bool more = tick < day.last_ready_tick;
            if (shuffle_seed) std::shuffle(order.begin(), order.end(), rng);
            for (int i = 0; i < N; ++i) {
                const int a = day.lane_airport[order[i]];
                ba_phase_c_pull_fast(c, acc[order[i]], a, prev, n_prev, n_late, tick);
                ba_phase_a<EXACT, PREDICT, TRK>(c, acc[order[i]], a, t, tick, n_cur);
                more |= ba_airport_busy<TRK>(c, a);
            }
            land();                                          // cp.async.wait_all + barrier
            bool ovf = false;
            for (int i = 0; i < N; ++i) ovf |= (acc[i].status & BA_S_OVERFLOW) != 0;
            if (ovf) break;                                  // the kernel aborts: re-run exactly
            busy = more;
        }
        land();
        ba_emu_defer = nullptr;
    } else {
        auto stage_arrivals = [&](int k) {
            uint32_t n = 0;
            if (!PREDICT && (uint32_t)k < c.cal_cap) {
                const uint32_t lo = c.cal_cnt[k - 1], hi = c.cal_cnt[k];
                n = hi - lo;
                if (n > c.due_cap) { n = c.due_cap; acc[0].status |= BA_S_OVERFLOW; }
                std::vector<uint32_t> sl;
                for (uint32_t i = lo; i < lo + n; ++i) sl.push_back(i);
                if (shuffle_seed) std::shuffle(sl.begin(), sl.end(), rng);
                for (uint32_t i : sl) ba_stage_arrival<true>(c, i, lo, k & 1);
            }
            c.cta[k & 1] = n;
        };
        stage_departures(tick + 1);
        stage_departures(tick + 2);
        stage_arrivals(tick + 1);
        while (busy) {
            if (tick >= plan.max_ticks) { st |= BA_S_TICK_CAP; break; }
            ++tick;
            t = c.tick_time[tick];
            const int prev = (tick - 1) & 1, cur = tick & 1;
            uint32_t n_due = c.cta[prev];
            if (n_due > c.due_cap) n_due = c.due_cap;
            bool more = tick < day.last_ready_tick;
            if (shuffle_seed) std::shuffle(order.begin(), order.end(), rng);
            for (int i = 0; i < N; ++i) {
                const int a = day.lane_airport[order[i]];
                ba_phase_c_pull<EXACT>(c, acc[order[i]], a, prev, n_due, tick);
                ba_phase_a<EXACT, PREDICT, TRK>(c, acc[order[i]], a, t, tick);
                more |= ba_airport_busy<TRK>(c, a);
            }
            if (!EXACT) {
                bool ovf = false;
                for (int i = 0; i < N; ++i) ovf |= (acc[i].status & BA_S_OVERFLOW) != 0;
                if (ovf) break;                      // the kernel aborts: the item is re-run exactly
            }
            for (int i = 0; i < N; ++i) ba_run_advance(c, day.lane_airport[order[i]], tick);
            if (PREDICT) {
                std::vector<uint32_t> slice;
                for (uint32_t i = c.cta[4]; i < c.cta[2]; ++i) slice.push_back(i);
                if (shuffle_seed) std::shuffle(slice.begin(), slice.end(), rng);
                for (uint32_t i : slice) ba_phase_b_pool(c, acc[i % N], i, t, cur);
                ba_pool_advance(c, 7u);
            } else if ((uint32_t)tick < c.cal_cap) {
                uint32_t n = c.cal_cnt[tick] - c.cal_cnt[tick - 1];
                if (n > c.due_cap) n = c.due_cap;
                std::vector<uint32_t> slice;
                for (uint32_t i = 0; i < n; ++i) slice.push_back(i);
                if (shuffle_seed) std::shuffle(slice.begin(), slice.end(), rng);
                for (uint32_t i : slice) ba_phase_b_arrival(c, i, cur);
            }
            stage_arrivals(tick + 1);
            stage_departures(tick + 2);
            busy = more;
        }
    }
    if (!PREDICT) for (int f = 0; f < day.F; ++f) ba_epilogue_flight<TRK>(c, acc[f % N], f);
    if (c.grow && !PREDICT) {
        for (int i = 0; i < N; ++i) ba_epilogue_airport<TRK>(c, day.lane_airport[i], N);
        for (int r = 0; r < day.Rd; ++r) ba_epilogue_route(c, r);
    }
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < N; ++i) {
        s0 += acc[i].lp; s1 += acc[i].g_sigma; s2 += acc[i].g_vt; s3 += acc[i].g_vu;
        st |= acc[i].status;
    }
    logp[pd] = (float)s0; status[pd] = st;
    if (ticks) ticks[pd] = tick;
    if (c.grow) {
        c.grow[c.route_base + R + 0] = (float)s1;
        c.grow[c.route_base + R + 1] = (float)s2;
        c.grow[c.route_base + R + 2] = (float)s3;
    }
}


// ---------------------------------------------------------------------------
// Ring-less scan (ba_scan2.cuh): one warp per particle-day.  The emulation runs the lanes of
// every step in a SHUFFLED order (flights in the prologue, airports inside a round of a
// tick), the two extremes of what lockstep lanes may observe of each other's writes.
// ---------------------------------------------------------------------------
static void run_item2(const BaPlanView &plan, const BaDayView &day, int p, int d,
                      const float *lat_airport, const float *lat_route, const float *scales,
                      const float *noise, uint64_t seed, int value_mode, int want_grad,
                      unsigned shuffle_seed, float *logp, uint32_t *status, int32_t *ticks,
                      float *tape, int32_t *pop_tick, float *vtape = nullptr)
{
    const int N = plan.N, R = plan.R, D = plan.D;
    const size_t pd = (size_t)p * D + d;
    std::vector<uint32_t> sh(ba2_smem_words(N) + 64, 0);
    std::vector<uint32_t> gs(ba2_scratch_words(day.F, N, R) + 16, 0xDEADBEEFu);
    Ba2Ctx c;
    memset(&c, 0, sizeof(c));
    c.nslot = (uint32_t)N;
    c.un = sh.data();
    c.s_rate = (float *)(sh.data() + BA2_KC + N); c.s_mt = c.s_rate + N;
    c.sigma = scales[0]; c.v_t = scales[1]; c.v_u = scales[2];
    c.inv_sigma = 1.0f / c.sigma; c.inv_sigma2 = c.inv_sigma * c.inv_sigma;
    c.inv_sigma3 = c.inv_sigma2 * c.inv_sigma; c.log_sigma = logf(c.sigma);
    c.inv_vt = 1.0f / c.v_t; c.inv_vu = 1.0f / c.v_u;
    c.c0_lp = ba_relaxed_bernoulli_lp(0.0f, 0.0f, nullptr);
    c.c1_lp = ba_relaxed_bernoulli_lp(0.0f, 1.0f, nullptr);
    c.value_mode = value_mode; c.want_grad = want_grad && tape;
    c.seed = seed; c.route_base = 2 * N; c.tick_time = plan.tick_time; c.max_ticks = plan.max_ticks;
    c.N = N; c.F = day.F;
    c.flights = day.flights; c.route = day.route; c.ap = day.airports;
    c.pos_live = day.pos_live; c.pos_dep = day.pos_dep; c.pos_arr = day.pos_arr;
    c.rt_gid = day.rt_gid; c.rt_off = day.rt_off; c.rt_flt = day.rt_flt; c.Rd = day.Rd;
    c.lq = day.lq; c.drun = day.drun; c.drun_off = day.drun_off; c.order = day.lane_sorted;
    c.first_tick = day.first_tick;
    c.lat_airport = lat_airport + (size_t)p * 2 * N;
    c.lat_route = lat_route + (size_t)p * R;
    c.noise = noise ? noise + pd * (size_t)plan.Fmax * 4 : nullptr;
    c.particle = p; c.dayidx = d;
    c.grow = c.want_grad ? tape + pd * (size_t)plan.grad_stride : nullptr;
    c.pop_tick = pop_tick ? pop_tick + pd * (size_t)plan.Fmax * 2 : nullptr;
    c.vrow = vtape ? vtape + pd * (size_t)plan.Fmax * 4 : nullptr;
    if (c.vrow) for (int i = 0; i < 4 * plan.Fmax; ++i) c.vrow[i] = 0.0f;
    ba2_carve_scratch(c, (uint32_t *)(((uintptr_t)gs.data() + 15) & ~(uintptr_t)15), day.F, N, R);
    const uint32_t n_arr = (uint32_t)day.n_arr;

    for (int a = 0; a < N; ++a) ba2_init_airport(c, a, N, false);
    for (int i = 0; i < BA2_KC; ++i) c.un[i] = 0;
    if (c.grow) for (int i = 0; i < R; ++i) c.grow[c.route_base + i] = 0.0f;
    if (c.pop_tick) for (int i = 2 * day.F; i < 2 * plan.Fmax; ++i) c.pop_tick[i] = -1;

    std::vector<BaAcc> acc(32, BaAcc{0, 0, 0, 0, 0});
    std::mt19937 rng(shuffle_seed + 977u * (unsigned)pd);
    std::vector<int> forder(day.F);
    for (int f = 0; f < day.F; ++f) forder[f] = f;
    if (shuffle_seed) std::shuffle(forder.begin(), forder.end(), rng);
    for (int f : forder) ba2_prologue_flight(c, acc[f % 32], f);
    ba2_bucket_scan(c, -1);
    if (shuffle_seed) std::shuffle(forder.begin(), forder.end(), rng);
    for (int f : forder) ba2_scatter_flight(c, f);
    for (int a = 0; a < N; ++a) ba2_split_cursor(c, a);
    ba2_split(c, -1, n_arr);
    uint32_t st = 0;
    for (int i = 0; i < 32; ++i) st |= acc[i].status;
    const bool aborted = (st & BA_S_OVERFLOW) != 0;

    uint32_t k = (uint32_t)(day.first_tick - 1);
    if (!aborted) {
        int nleft = 0;
        for (int slot = 0; slot < N; ++slot) {
            const int a = c.order[slot];
            ba2_state_init(c, (uint32_t)slot, a);
            nleft += (c.ap[a].dep_live + c.ap[a].arr_cnt) != 0 ? 1 : 0;
        }
        std::vector<int> order(32);
        while (nleft > 0) {
            if ((int)k >= plan.max_ticks) { st |= BA_S_TICK_CAP; break; }
            ++k;
            const float t = c.tick_time[k];
            for (int base = 0; base < N; base += 32) {          // rounds
                for (int l = 0; l < 32; ++l) order[l] = l;
                if (shuffle_seed) std::shuffle(order.begin(), order.end(), rng);
                for (int l : order) {
                    const int slot = base + l;
                    if (slot < N && ba2_airport_tick(c, (uint32_t)slot, k, t, acc[l].status)) nleft -= 1;
                }
            }
        }
    }
    int tick = (int)k;
    if (day.F > 0 && tick < day.last_ready_tick) tick = day.last_ready_tick;
    c.s_rate = (float *)sh.data(); c.s_mt = c.s_rate + N;
    c.accA = (unsigned long long *)(sh.data() + ((2 * N + 1) & ~1));
    for (int a = 0; a < N; ++a) ba2_init_airport(c, a, N, true);
    for (int i = 0; i < 3 * N; ++i) c.accA[i] = 0ull;
    for (int r = 0; r < R; ++r) c.accR[r] = 0ull;
    if (!aborted) for (int f = 0; f < day.F; ++f) ba2_epilogue_flight(c, acc[f % 32], f);
    if (c.vrow && !aborted) {
        for (int a = 0; a < N; ++a) ba2_value_grad_airport(c, a);
        for (int f = 0; f < day.F; ++f) ba2_value_grad_finish(c, f);
    }
    if (c.grow && !aborted) {
        for (int a = 0; a < N; ++a) ba2_epilogue_airport(c, a, N);
        for (int r = 0; r < R; ++r) ba2_epilogue_route(c, r);
    }
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < 32; ++i) {
        s0 += acc[i].lp; s1 += acc[i].g_sigma; s2 += acc[i].g_vt; s3 += acc[i].g_vu;
        st |= acc[i].status;
    }
    logp[pd] = (float)s0; status[pd] = st;
    if (ticks) ticks[pd] = tick;
    if (c.grow) {
        c.grow[c.route_base + R + 0] = (float)s1;
        c.grow[c.route_base + R + 1] = (float)s2;
        c.grow[c.route_base + R + 2] = (float)s3;
    }
}

extern "C" int ba_emu_scan2_ok(const EmuPlan *pl) { return pl->host.view.scan2_ok; }

// same selection as ba_forward: the ring-less scan, overflowed items re-run with exact lists
extern "C" int ba_emu_forward2(const EmuPlan *pl, int P, const float *lat_airport,
                               const float *lat_route, const float *scales, const float *noise,
                               uint64_t seed, int value_mode, int want_grad, int rerun_overflow,
                               unsigned shuffle_seed, float *logp, uint32_t *status,
                               int32_t *ticks, float *tape, int32_t *pop_tick, float *vtape)
{
    const BaPlanView &plan = pl->host.view;
    if (!plan.scan2_ok) return 1;
    for (int d = 0; d < plan.D; ++d)
        for (int p = 0; p < P; ++p) {
            const BaDayView &day = plan.days[d];
            run_item2(plan, day, p, d, lat_airport, lat_route, scales, noise, seed, value_mode,
                      want_grad, shuffle_seed, logp, status, ticks, tape, pop_tick, vtape);
            if (rerun_overflow && !vtape && (status[(size_t)p * plan.D + d] & BA_S_OVERFLOW))
                run_item<true, false, false>(plan, day, p, d, P, lat_airport, lat_route, scales, noise,
                                             seed, value_mode, want_grad, 0, shuffle_seed, logp, status,
                                             ticks, tape, pop_tick);
        }
    return 0;
}

extern "C" int ba_emu_forward(const EmuPlan *pl, int P, const float *lat_airport,
                              const float *lat_route, const float *scales, const float *noise,
                              uint64_t seed, int value_mode, int want_grad, int exact,
                              int rerun_overflow, int replay_waits, unsigned shuffle_seed, float *logp,
                              uint32_t *status, int32_t *ticks, float *tape, int32_t *pop_tick)
{
    const BaPlanView &plan = pl->host.view;
    for (int d = 0; d < plan.D; ++d)
        for (int p = 0; p < P; ++p) {
            const BaDayView &day = plan.days[d];
            // same kernel selection as ba_forward: built without the aircraft-bookkeeping code
            // when no airport of the plan needs it
#define BA_EMU_RUN(EX, TRK) run_item<EX, false, TRK>(plan, day, p, d, P, lat_airport, lat_route, scales, noise, seed, \
                               value_mode, want_grad, replay_waits, shuffle_seed, logp, status, ticks, tape, pop_tick)
            const bool trk = plan.any_track != 0;
            if (exact) {
                if (trk) BA_EMU_RUN(true, true); else BA_EMU_RUN(true, false);
            } else {
                if (trk) BA_EMU_RUN(false, true); else BA_EMU_RUN(false, false);
                if (rerun_overflow && (status[(size_t)p * plan.D + d] & BA_S_OVERFLOW)) {
                    if (trk) BA_EMU_RUN(true, true); else BA_EMU_RUN(true, false);
                }
            }
#undef BA_EMU_RUN
        }
    return 0;
}

extern "C" void ba_emu_backward(const EmuPlan *pl, int P, const float *dlogp, const float *tape,
                                float *g_airport, float *g_route, float *g_scales)
{
    const BaPlanView &v = pl->host.view;
    const int CN = v.C * v.N, S = v.grad_stride;
    for (int p = 0; p < P; ++p)
        for (int j = 0; j < S; ++j) {
            float s = 0.0f;
            for (int d = 0; d < v.D; ++d)
                s = fmaf(dlogp[(size_t)p * v.D + d], tape[((size_t)p * v.D + d) * S + j], s);
            if (j < CN) g_airport[(size_t)p * CN + j] = s;
            else if (j < CN + v.R) g_route[(size_t)p * v.R + (j - CN)] = s;
            else g_scales[(size_t)p * 3 + (j - CN - v.R)] = s;
        }
}

extern "C" void ba_emu_noise(const EmuPlan *pl, int P, uint64_t seed, float *noise)
{
    const BaPlanView &v = pl->host.view;
    for (int p = 0; p < P; ++p)
        for (int d = 0; d < v.D; ++d)
            for (int f = 0; f < v.Fmax; ++f) {
                BaNoise4 z = ba_philox_noise(seed, p, d, f);
                float *o = noise + (((size_t)p * v.D + d) * v.Fmax + f) * 4;
                o[0] = z.e_dep; o[1] = z.eps_travel; o[2] = z.e_arr; o[3] = z.eps_turn;
            }
}

extern "C" int ba_emu_predict(const EmuPlan *pl, int P, const float *lat_airport,
                              const float *lat_route, const float *scales, const float *noise,
                              const float *noise2, uint64_t seed, unsigned shuffle_seed,
                              float *sim, uint32_t *status, int32_t *ticks, int32_t *pop_tick)
{
    const BaPlanView &plan = pl->host.view;
    std::vector<float> logp((size_t)P * plan.D);
    for (int d = 0; d < plan.D; ++d)
        for (int p = 0; p < P; ++p)
            run_item<true, true>(plan, plan.days[d], p, d, P, lat_airport, lat_route, scales, noise,
                                 seed, 0, 0, 0, shuffle_seed, logp.data(), status, ticks, nullptr,
                                 pop_tick, noise2, sim);
    return 0;
}
