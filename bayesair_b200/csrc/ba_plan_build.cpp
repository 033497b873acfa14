// Host-side schedule compiler: BaDayTable (flights in stable scheduled-departure
// order, airports in day order) -> static tables of the kernel.
//
// What it fixes at plan time (none of it depends on a particle's latents):
//   * per-airport departure lists in pending order (the reference scans ALL pending
//     flights every tick and buckets them by origin, bayes_air/network.py:57-68;
//     restricted to one origin the scan order is this list);
//   * the compact route table (dense [N,N] travel times -> [R]);
//   * which airports need aircraft bookkeeping at all (ba_plan.h:
//     BA_SIMPLE_RATIO);
//   * list capacities: a small shared-memory ring per airport sized from the
//     airport's daily traffic (fast variant) and exact-capacity rings in global
//     scratch (the variant a particle-day is re-run with if a ring overflows);
//   * the lane -> airport assignment, sorted by daily traffic so that the airports
//     of one warp have similar trip counts.
#include "ba_plan_build.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

static uint32_t pow2ceil(uint32_t v)
{
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

static double env_double(const char *name, double dflt)
{
    const char *s = std::getenv(name);
    return s ? std::atof(s) : dflt;
}

int ba_build_host_plan(const BaDayTable *days, int n_days, int n_airports, float delta_t,
                       float max_t, int include_cancellations, int max_ticks,
                       BaHostPlan *out, std::string *err, double q_div_arg, double q_min_arg)
{
    auto fail = [&](int code, const std::string &m) { if (err) *err = m; return code; };
    if (!days || n_days <= 0 || n_airports <= 0 || !out)
        return fail(BA_ERR_BAD_ARG, "ba_plan_create: null / empty argument");
    if (!(delta_t > 0.0f) || max_ticks <= 0)
        return fail(BA_ERR_BAD_ARG, "ba_plan_create: delta_t and max_ticks must be positive");
    if (n_airports > BA_MAX_AIRPORTS)
        return fail(BA_ERR_LIMIT, "ba_plan_create: more than 4096 airports");

    // tunables of the shared-memory variant (documented in DESIGN.md)
    const double q_div = q_div_arg > 0.0 ? q_div_arg : env_double("BA_Q_DIV", 32.0);   // ring = traffic / q_div
    const double q_min = q_min_arg > 0.0 ? q_min_arg : env_double("BA_Q_MIN", 4.0);
    if (max_ticks > 65000)
        return fail(BA_ERR_LIMIT, "ba_plan_create: max_ticks must be <= 65000");

    const int N = n_airports;
    // the clock: t_k = fl32(t_{k-1} + fl32(delta_t))  (model.py:158-161)
    out->tick_time.assign((size_t)max_ticks + 2, 0.0f);
    for (int k = 1; k <= max_ticks + 1; ++k) out->tick_time[k] = out->tick_time[k - 1] + delta_t;
    out->days.clear();
    out->days.resize(n_days);
    out->max_dslice = 0;

    // ---- route table over all days ----------------------------------------
    std::map<int64_t, int32_t> route_id;
    for (int d = 0; d < n_days; ++d) {
        const BaDayTable &T = days[d];
        if (T.n_airports != N)
            return fail(BA_ERR_BAD_ARG, "ba_plan_create: every day must have the same airports");
        if (T.n_flights < 0 || T.n_flights >= BA_MAX_FLIGHTS)
            return fail(BA_ERR_LIMIT, "ba_plan_create: flights per day out of range");
        for (int a = 0; a < N; ++a)
            if (T.h_airport_global[a] < 0 || T.h_airport_global[a] >= N)
                return fail(BA_ERR_BAD_ARG, "ba_plan_create: airport_global out of range");
        for (int f = 0; f < T.n_flights; ++f) {
            int o = T.h_origin[f], e = T.h_dest[f];
            if (o < 0 || o >= N || e < 0 || e >= N || o == e)
                return fail(BA_ERR_BAD_ARG, "ba_plan_create: bad origin/destination index");
            if (f > 0 && T.h_sched_dep[f] < T.h_sched_dep[f - 1])
                return fail(BA_ERR_BAD_ARG,
                            "ba_plan_create: flights must be sorted by scheduled departure");
            int64_t key = (int64_t)T.h_airport_global[o] * N + T.h_airport_global[e];
            route_id.emplace(key, 0);
        }
    }
    out->route_origin.clear();
    out->route_dest.clear();
    {
        int32_t r = 0;
        for (auto &kv : route_id) {
            kv.second = r++;
            out->route_origin.push_back((int32_t)(kv.first / N));
            out->route_dest.push_back((int32_t)(kv.first % N));
        }
    }
    const int R = (int)route_id.size();

    int Fmax = 0;
    uint64_t scratch_fast = 0, scratch_exact = 0;
    uint32_t smem_words = 0;
    for (int d = 0; d < n_days; ++d) {
        const BaDayTable &T = days[d];
        BaHostDay &H = out->days[d];
        const int F = T.n_flights;
        Fmax = std::max(Fmax, F);
        H.flights.resize(F);
        H.route.resize(F);
        H.airports.assign(N, BaAirport{});
        std::vector<int> dep_cnt(N, 0), arr_cnt(N, 0);
        for (int f = 0; f < F; ++f) {
            int o = T.h_origin[f], e = T.h_dest[f];
            int c = T.h_cancel_obs[f];
            uint32_t cc = c < 0 ? 3u : (c ? 1u : 0u);
            H.flights[f].sched_dep = T.h_sched_dep[f];
            H.flights[f].act_dep = T.h_act_dep[f];
            H.flights[f].act_arr = T.h_act_arr[f];
            H.flights[f].od = (uint32_t)o | ((uint32_t)e << 12) | (cc << 24);
            H.route[f] = route_id[(int64_t)T.h_airport_global[o] * N + T.h_airport_global[e]];
            dep_cnt[o]++;
            // a flight observed as cancelled never reaches a queue (network.py:127-129)
            if (cc != 1u) arr_cnt[e]++;
        }
        // departures grouped by origin, ascending flight index (= pending order)
        H.dep_fid.resize(F);
        {
            std::vector<int> off(N + 1, 0);
            for (int a = 0; a < N; ++a) off[a + 1] = off[a] + dep_cnt[a];
            std::vector<int> cur(off.begin(), off.end() - 1);
            for (int f = 0; f < F; ++f) H.dep_fid[cur[T.h_origin[f]]++] = f;
            for (int a = 0; a < N; ++a) { H.airports[a].dep_off = off[a]; }
        }
        H.dep_sched.resize(F); H.dep_word.resize(F); H.pos_dep.resize(F);
        for (int k = 0; k < F; ++k) {
            const int f = H.dep_fid[k];
            H.pos_dep[f] = k;
            H.dep_sched[k] = T.h_sched_dep[f];
            H.dep_word[k] = (uint32_t)f | (((H.flights[f].od >> 12) & 0xFFFu) << BA_Q_DST_SHIFT) |
                            (((H.flights[f].od >> 24) & 3u) << BA_DW_CANCEL_SHIFT);
        }
        // arrival order: flights grouped by destination, ascending flight index
        // (cancelled-as-observed flights never arrive); and the day's route CSR
        {
            std::vector<int> off(N + 1, 0);
            for (int a = 0; a < N; ++a) off[a + 1] = off[a] + arr_cnt[a];
            H.pos_arr.assign(F, -1);
            std::vector<int> cur(off.begin(), off.end() - 1);
            for (int f = 0; f < F; ++f)
                if (((H.flights[f].od >> 24) & 3u) != 1u) H.pos_arr[f] = cur[T.h_dest[f]]++;
            for (int a = 0; a < N; ++a) H.airports[a].arr_off = off[a];
            std::map<int32_t, std::vector<int32_t>> by_route;
            for (int f = 0; f < F; ++f)
                if (((H.flights[f].od >> 24) & 3u) != 1u) by_route[H.route[f]].push_back(f);
            H.rt_gid.clear(); H.rt_off.assign(1, 0); H.rt_flt.clear();
            for (auto &kv : by_route) {
                H.rt_gid.push_back(kv.first);
                for (int32_t f : kv.second) H.rt_flt.push_back(f);
                H.rt_off.push_back((int32_t)H.rt_flt.size());
            }
        }
        // largest number of departures of one airport that become ready in the same tick
        std::vector<int> max_batch(N, 0);
        std::vector<int> ready_tick(F, 1), batch_rank(F, 0);
        int last_ready_tick = 0;
        {
            std::vector<int> last_tick(N, -1), run(N, 0);
            const std::vector<float> &tt = out->tick_time;
            int k = 1;
            for (int f = 0; f < F; ++f) {          // flights are sorted by scheduled departure
                while (k < max_ticks + 1 && tt[k] < T.h_sched_dep[f]) ++k;
                const int o = T.h_origin[f];
                if (last_tick[o] == k) run[o]++; else { last_tick[o] = k; run[o] = 1; }
                max_batch[o] = std::max(max_batch[o], run[o]);
                ready_tick[f] = k; batch_rank[f] = run[o] - 1;
                last_ready_tick = std::max(last_ready_tick, k);
            }
        }
        std::vector<int> dep_live(N, 0);
        for (int f = 0; f < F; ++f)
            if (((H.flights[f].od >> 24) & 3u) != 1u) dep_live[T.h_origin[f]]++;
        uint32_t q_s = 0, q_x = 0, turn = 0, av = 0, dl = 0;
        uint32_t any_t = 0, any_a = 0;
        for (int a = 0; a < N; ++a) {
            BaAirport &A = H.airports[a];
            A.dep_cnt = dep_cnt[a];
            A.arr_cnt = arr_cnt[a];
            A.dep_live = dep_live[a];
            A.global = T.h_airport_global[a];
            uint32_t flags = 0;
            if (include_cancellations) flags = BA_AP_TRACK_T | BA_AP_TRACK_A;
            else {
                const bool simple = dep_cnt[a] < 1000 &&
                    (double)(1000 - dep_cnt[a]) >= BA_SIMPLE_RATIO * (double)std::max(1, max_batch[a]);
                if (!simple) flags |= BA_AP_TRACK_T;
                if (dep_cnt[a] >= 1000) flags |= BA_AP_TRACK_T | BA_AP_TRACK_A;
            }
            A.flags = flags;
            if (flags & BA_AP_TRACK_T) any_t = 1;
            if (flags & BA_AP_TRACK_A) any_a = 1;
            const uint32_t traffic = (uint32_t)(dep_cnt[a] + arr_cnt[a]);
            // exact variant
            uint32_t qx = pow2ceil(std::max(1u, traffic));
            A.q_off_x = q_x; A.q_mask_x = qx - 1; q_x += qx;
            // shared variant
            uint32_t want_q = (uint32_t)std::ceil(std::max(q_min, traffic / q_div));
            uint32_t qs = std::min(qx, pow2ceil(std::max(1u, want_q)));
            A.q_off_s = q_s; A.q_mask_s = qs - 1; q_s += qs;
            // always-global lists (only for airports that track aircraft)
            A.turn_off = turn; A.av_off = av; A.dl_off = dl;
            A.av_mask = 0; A.dl_mask = 0;
            if (flags & BA_AP_TRACK_T) {
                turn += std::max(1, arr_cnt[a]);
                uint32_t dc = pow2ceil(std::max(1, dep_cnt[a]));
                A.dl_mask = dc - 1; dl += dc;
            }
            if (flags & BA_AP_TRACK_A) {
                // releases + run-length-encoded reserve markers (ba_sim.cuh)
                uint32_t ac = pow2ceil(2u * (uint32_t)arr_cnt[a] + 4u);
                A.av_mask = ac - 1; av += ac;
            }
        }
        // departure calendar: flights of bookkeeping-free airports, by ready tick
        {
            H.dcal_off.assign((size_t)last_ready_tick + 2, 0);
            std::vector<int> cnt((size_t)last_ready_tick + 2, 0);
            auto on_calendar = [&](int f) {
                // (an unobserved cancellation at such an airport can only come out "not
                //  cancelled": its probability is exactly 0, see BA_SIMPLE_RATIO)
                return !(H.airports[T.h_origin[f]].flags & BA_AP_TRACK_T) &&
                       ((H.flights[f].od >> 24) & 3u) != 1u;
            };
            for (int f = 0; f < F; ++f) if (on_calendar(f)) cnt[ready_tick[f]]++;
            for (int k = 0; k <= last_ready_tick; ++k) H.dcal_off[k + 1] = H.dcal_off[k] + cnt[k];
            H.dcal_rec.assign(2 * (size_t)H.dcal_off[last_ready_tick + 1], 0u);
            H.dep_cpos.assign((size_t)F, -1);
            // within a tick: grouped by origin (day order), pending order inside a group
            std::vector<std::vector<int>> per_tick((size_t)last_ready_tick + 1);
            for (int f = 0; f < F; ++f) if (on_calendar(f)) per_tick[ready_tick[f]].push_back(f);
            std::vector<std::vector<uint32_t>> runs(N);
            int max_slice = 0;
            for (int k = 0; k <= last_ready_tick; ++k) {
                std::vector<int> &v = per_tick[k];
                std::stable_sort(v.begin(), v.end(),
                                 [&](int x, int y) { return T.h_origin[x] < T.h_origin[y]; });
                max_slice = std::max(max_slice, (int)v.size());
                int pos = H.dcal_off[k];
                for (size_t i = 0; i < v.size(); ++i, ++pos) {
                    const int f = v[i], o = T.h_origin[f];
                    const int kd = H.pos_dep[f];
                    const float ready = std::max(0.0f, H.dep_sched[kd]);     // network.py:60-68
                    uint32_t ready_bits;
                    memcpy(&ready_bits, &ready, 4);
                    H.dcal_rec[2 * pos] = BA_Q_DEP | (H.dep_word[kd] & BA_DW_ENTRY_MASK);
                    H.dcal_rec[2 * pos + 1] = ready_bits;
                    H.dep_cpos[kd] = pos;
                    if (i == 0 || T.h_origin[v[i - 1]] != o) {
                        runs[o].push_back((uint32_t)k);
                        runs[o].push_back((uint32_t)(pos - H.dcal_off[k]));
                        runs[o].push_back(1u);
                    } else {
                        runs[o].back() += 1u;
                    }
                }
            }
            H.drun_off.assign(N + 1, 0);
            H.drun.clear();
            for (int a = 0; a < N; ++a) {
                // two words per run: tick, first record relative to the slice << 16 | records
                for (size_t r = 0; r + 2 < runs[a].size(); r += 3) {
                    H.drun.push_back(runs[a][r]);
                    H.drun.push_back((runs[a][r + 1] << 16) | (runs[a][r + 2] & 0xFFFFu));
                }
                // sentinel run: never matches a tick
                H.drun.push_back(0xFFFFFFFFu); H.drun.push_back(0u);
                H.drun_off[a + 1] = (int32_t)(H.drun.size() / 2);
            }
            H.view.max_dslice = max_slice;
            out->max_dslice = std::max(out->max_dslice, max_slice);
        }
        // ring-less scan (ba_scan2.cuh): live departures grouped by origin (day order), pending
        // order inside a group, one sentinel record per airport; arrival lists of arr_cnt + 1
        // slots.  The list position of a flight ("lpos") is also the static part of the
        // in-transit ordering key: at bookkeeping-free airports departures leave in this order.
        {
            H.lq.clear();
            H.pos_live.assign((size_t)F, -1);
            uint32_t as = 0;
            for (int a = 0; a < N; ++a) {
                BaAirport &A = H.airports[a];
                A.lq_off = (uint32_t)H.lq.size();
                A.as_off = as;
                as += (uint32_t)arr_cnt[a] + 1u;
                for (int i = 0; i < dep_cnt[a]; ++i) {
                    const int f = H.dep_fid[(size_t)A.dep_off + i];
                    if (((H.flights[f].od >> 24) & 3u) == 1u) continue;
                    H.pos_live[f] = (int32_t)H.lq.size();
                    BaLiveDep r;
                    r.tick = (uint32_t)ready_tick[f];
                    r.qstart = std::max(0.0f, T.h_sched_dep[f]);
                    r.dslot = (uint32_t)T.h_dest[f];        // airport for now, slot below
                    r.das_off = 0;
                    H.lq.push_back(r);
                }
                BaLiveDep sentinel;
                sentinel.tick = 0xFFFFFFFFu; sentinel.qstart = 0.0f; sentinel.dslot = 0; sentinel.das_off = 0;
                H.lq.push_back(sentinel);
            }
            H.view.n_live = (int32_t)H.lq.size() - N;
            H.view.n_arr = (int32_t)as - N;
        }
        // lane -> airport, heaviest first
        H.lane_airport.resize(N);
        std::iota(H.lane_airport.begin(), H.lane_airport.end(), 0);
        std::stable_sort(H.lane_airport.begin(), H.lane_airport.end(), [&](int x, int y) {
            return dep_cnt[x] + arr_cnt[x] > dep_cnt[y] + arr_cnt[y];
        });
        H.lane_sorted = H.lane_airport;
        {
            // ... and dealt round-robin over the full warps: a warp's time per tick is the sum
            // over its loops of the longest trip count among its lanes, and the maxima over
            // 1/k of the hubs are smaller than over all of them (measured: +8 % with
            // saturated hubs, +2 % otherwise)
            const int full = N / 32;
            const char *lo = std::getenv("BA_LANE_ORDER");        // "0": keep the sorted order
            if (full > 1 && !(lo && std::atoi(lo) == 0)) {
                const std::vector<int32_t> srt = H.lane_airport;
                std::vector<int32_t> lanes;
                for (int w = 0; w < full; ++w)
                    for (int l = 0; l < 32; ++l) lanes.push_back(srt[(size_t)l * full + w]);
                for (int r = 32 * full; r < N; ++r) lanes.push_back(srt[r]);
                H.lane_airport = lanes;
            }
        }
        for (int sl = 0; sl < N; ++sl) H.airports[(size_t)H.lane_airport[sl]].slot = (uint32_t)sl;
        {
            std::vector<uint32_t> sorted_slot((size_t)N, 0);
            for (int sl = 0; sl < N; ++sl) sorted_slot[(size_t)H.lane_sorted[sl]] = (uint32_t)sl;
            for (BaLiveDep &r : H.lq) {
                if (r.tick == 0xFFFFFFFFu) continue;
                const uint32_t dest = r.dslot;
                r.das_off = H.airports[dest].as_off;
                r.dslot = sorted_slot[dest];
            }
        }
        // idle prefix of the day: ticks before the first scheduled departure do nothing
        // (no pending flight is ready, nothing is queued or in transit) as long as the
        // reserve injection (t >= max_t, model.py:170-174) has not started.
        int first_tick = 1;
        float t_before = 0.0f;
        if (F > 0) {
            float t = 0.0f;
            const float first_sched = T.h_sched_dep[0];
            for (int k = 1; k <= max_ticks; ++k) {
                float tn = t + delta_t;             // fp32, model.py:161
                if (first_sched <= tn || tn >= max_t) { first_tick = k; t_before = t; break; }
                t = tn;
                first_tick = k + 1; t_before = t;
            }
        }
        BaDayView &V = H.view;
        V.N = N; V.F = F;
        V.flights = H.flights.data(); V.route = H.route.data(); V.dep_fid = H.dep_fid.data();
        V.dep_sched = H.dep_sched.data(); V.dep_word = H.dep_word.data();
        V.pos_dep = H.pos_dep.data(); V.pos_arr = H.pos_arr.data();
        V.rt_gid = H.rt_gid.data(); V.rt_off = H.rt_off.data(); V.rt_flt = H.rt_flt.data();
        V.Rd = (int32_t)H.rt_gid.size();
        V.dcal_off = H.dcal_off.data(); V.dcal_rec = H.dcal_rec.data();
        V.dep_cpos = H.dep_cpos.data();
        V.last_ready_tick = last_ready_tick;
        V.drun_off = H.drun_off.data(); V.drun = H.drun.data();
        V.any_track_t = any_t; V.any_track_a = any_a;
        V.airports = H.airports.data(); V.lane_airport = H.lane_airport.data();
        V.lane_sorted = H.lane_sorted.data();
        V.q_total_s = q_s; V.q_total_x = q_x;
        V.turn_total = turn; V.av_total = av; V.dl_total = dl;
        V.first_tick = first_tick; V.t_before_first = t_before;
        V.lq = H.lq.data(); V.pos_live = H.pos_live.data();
        smem_words = std::max(smem_words, 6u * q_s);
    }

    out->day_views.resize(n_days);
    for (int d = 0; d < n_days; ++d) {
        const BaDayView &V = out->days[d].view;
        out->day_views[d] = V;
        // the staged-departure buffers are strided by the plan-wide largest calendar slice
        scratch_fast = std::max(scratch_fast, ba_scratch_words(V, false, max_ticks, out->max_dslice));
        scratch_exact = std::max(scratch_exact, ba_scratch_words(V, true, max_ticks, out->max_dslice));
    }
    BaPlanView &P = out->view;
    P.D = n_days; P.N = N; P.Fmax = Fmax; P.R = R;
    P.C = include_cancellations ? 4 : 2;
    P.include_cancellations = include_cancellations ? 1 : 0;
    P.max_dslice = out->max_dslice;
    P.any_track = 0;
    for (int d = 0; d < n_days; ++d) P.any_track |= (int32_t)out->days[d].view.any_track_t;
    P.max_ticks = max_ticks; P.delta_t = delta_t; P.max_t = max_t;
    // the ring-less scan handles fully bookkeeping-free plans with one lane per airport
    // (list positions are 16-bit: flights + airports of a day below 65536)
    P.scan2_ok = (!P.any_track && N <= 1024 && Fmax + N < (1 << 16)) ? 1 : 0;
    P.grad_stride = P.C * N + R + 3;
    P.days = out->day_views.data();
    P.tick_time = out->tick_time.data();
    out->scratch_words_fast = scratch_fast;
    out->scratch_words_exact = scratch_exact;
    out->smem_list_words = smem_words;
    return BA_OK;
}
