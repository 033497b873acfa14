// sm_100a kernels of the BayesAir hot path.
//
//   ba_forward_kernel<EXACT>   persistent CTAs; each pulls (particle, day) work items
//                              from an atomic counter and runs the tick loop of
//                              ba_sim.cuh with one lane per airport.  EXACT=false keeps
//                              the runway-queue rings and inbound bags in shared memory
//                              (plan-time capacities); EXACT=true keeps them in per-CTA
//                              global scratch with capacities no schedule can overflow.
//   ba_forward2_kernel<NW>     the ring-less scan of ba_scan2.cuh for plans without aircraft
//                              bookkeeping and with every flight observed (the training
//                              configurations): NW warps per (particle, day), one for
//                              networks up to 128 airports.
//   ba_backward_kernel         contracts the per-(particle, day) gradient rows written by
//                              the forward kernels with the cotangent d logp.
//   ba_backward_values_kernel  the same for the value tape of a value-mode forward.
//   ba_noise_kernel            materialises the Philox per-flight noise (tests / callers
//                              that want to keep their draws).
//   ba_collect_kernel          compacts the particle-days that must be re-run.
//
// Nothing here is a dense contraction: no tensor cores, by design.
#include <algorithm>
#include "ba_kernels.cuh"
#include "ba_sim.cuh"
#include "ba_scan2.cuh"

#define BA_RED_WORDS (32 * 6 + 4)
#define BA_MAX_DEVICES 64

int ba_forward_block_threads(int n_airports)
{
    int t = ((n_airports + 31) / 32) * 32;
    if (t < 32) t = 32;
    if (t > 1024) t = 1024;
    return t;
}

static __host__ __device__ size_t ba_state_words(int n_airports, bool tracked)
{
    const size_t per = tracked ? BA_NWORDS_STATE_TRACK : BA_NWORDS_STATE_SIMPLE;
    return (per * (size_t)n_airports + 3) & ~(size_t)3;
}

size_t ba_forward_smem_bytes(int n_airports, uint32_t list_words, bool exact, bool tracked,
                             int max_dslice)
{
    size_t w = ba_state_words(n_airports, tracked) + BA_RED_WORDS + BA_CTA_WORDS;
    // list_words = 6 words per queue entry; plus calendar counters, due buffers, staged departures
    if (!exact) w += list_words + BA_SHARED_FIXED_WORDS +
                     2 * BA_DSTAGE_WORDS * (size_t)((max_dslice + 1) & ~1);
    return w * 4;
}

// resident CTAs per SM the register allocation aims for, per block-size class
// (measured on the bench workload: 7 CTAs/SM at 72 registers 7.4e5, 8 at 64 registers 7.8e5,
//  9 at 56 registers 7.5e5 particle-days/s; the plan sizes its shared rings to match)
#ifndef BA_MIN_BLOCKS_128
#define BA_MIN_BLOCKS_128 8
#endif
#define BA_MIN_BLOCKS(MAXT) ((MAXT) == 128 ? BA_MIN_BLOCKS_128 : ((MAXT) == 256 ? 3 : 1))

// resident CTAs per SM the plan should size its shared memory for
int ba_forward_target_ctas(int block_threads)
{
    return block_threads <= 128 ? BA_MIN_BLOCKS(128) : (block_threads <= 256 ? BA_MIN_BLOCKS(256) : 1);
}

template <bool EXACT, int MAXT, bool PREDICT, bool TRK>
__global__ void __launch_bounds__(MAXT, BA_MIN_BLOCKS(MAXT)) ba_forward_kernel(const BaFwdParams prm)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = (nthr + 31) >> 5;
    const BaPlanView &plan = prm.plan;
    const int N = plan.N, D = plan.D, C = plan.C, R = plan.R;

    // shared memory: everything of fixed size first (immediate addresses), then the
    // per-airport records, then the size-dependent lists of the fast variant
    uint32_t *red = smem;                                    // [BA_RED_WORDS], 16 B aligned
    uint32_t *ctaw = red + BA_RED_WORDS;                     // [BA_CTA_WORDS]
    uint32_t *fixed = ctaw + BA_CTA_WORDS;                   // [BA_SHARED_FIXED_WORDS] (fast)
    uint32_t *state_w = fixed + (EXACT ? 0 : BA_SHARED_FIXED_WORDS);
    uint32_t *pool = state_w + ba_state_words(N, plan.any_track != 0);
    uint32_t *gscr = prm.scratch + (size_t)blockIdx.x * prm.scratch_stride;

    BaCtx c;
    ba_carve_state(c, state_w, N, plan.any_track != 0);
    c.sigma = prm.scales[0]; c.v_t = prm.scales[1]; c.v_u = prm.scales[2];
    c.inv_sigma = 1.0f / c.sigma;
    c.inv_sigma2 = c.inv_sigma * c.inv_sigma;
    c.inv_sigma3 = c.inv_sigma2 * c.inv_sigma;
    c.log_sigma = logf(c.sigma);
    c.inv_vt = 1.0f / c.v_t; c.inv_vu = 1.0f / c.v_u;
    c.c0_lp = ba_relaxed_bernoulli_lp(0.0f, 0.0f, nullptr);
    c.c1_lp = ba_relaxed_bernoulli_lp(0.0f, 1.0f, nullptr);
    c.value_mode = prm.value_mode; c.want_grad = prm.want_grad && prm.tape != nullptr;
    c.cancel_mode = plan.include_cancellations;
    c.replay_always = prm.replay_waits;
    c.max_t = plan.max_t;
    c.seed = prm.seed_dev ? *prm.seed_dev : prm.seed;
    c.route_base = C * N;
    c.tick_time = plan.tick_time;
    c.cta = ctaw;

    // optional phase timing (thread 0 reads the SM clock after each barrier)
    long long pc_last = 0;
    const bool prof = prm.phase_cycles != nullptr && tid == 0;
    // (accumulated straight into global memory: no per-thread counters held in registers)
#define BA_MARK(slot) do { if (prof) { long long now_ = clock64(); atomicAdd(prm.phase_cycles + (slot), (unsigned long long)(now_ - pc_last)); pc_last = now_; } } while (0)

    for (;;) {
        if (prof) pc_last = clock64();
        // ---- next work item ------------------------------------------------
        if (tid == 0) red[BA_RED_WORDS - 1] = atomicAdd(prm.work_counter, 1u);
        __syncthreads();
        const int wi = (int)red[BA_RED_WORDS - 1];
        const int n_work = prm.n_work_dev ? (int)*prm.n_work_dev : prm.n_work;
        if (wi >= n_work) break;
        int p, d;
        if (prm.work_list) { const int pd = prm.work_list[wi]; p = pd / D; d = pd - p * D; }
        else { d = wi / prm.P; p = wi - d * prm.P; }     // same-day items adjacent in time
        const size_t pd = (size_t)p * D + d;
        const BaDayView day = plan.days[d];

        c.N = N; c.F = day.F;
        c.flights = day.flights; c.route = day.route; c.ap = day.airports;
        c.dep_sched = day.dep_sched; c.dep_word = day.dep_word; c.pos_dep = day.pos_dep;
        c.pos_arr = day.pos_arr; c.rt_gid = day.rt_gid; c.rt_off = day.rt_off;
        c.rt_flt = day.rt_flt; c.Rd = day.Rd;
        c.lat_route = prm.lat_route + (size_t)p * R;
        c.noise = prm.noise ? prm.noise + pd * (size_t)plan.Fmax * 4 : nullptr;
        c.noise2 = (PREDICT && prm.noise2) ? prm.noise2 + pd * (size_t)plan.Fmax * 4 : nullptr;
        c.sim = PREDICT ? prm.sim + pd * (size_t)plan.Fmax * 3 : nullptr;
        c.particle = (uint32_t)p + prm.particle_offset; c.dayidx = (uint32_t)d + prm.day_offset;
        c.grow = c.want_grad ? prm.tape + pd * (size_t)plan.grad_stride : nullptr;
        c.pop_tick = prm.pop_tick ? prm.pop_tick + pd * (size_t)plan.Fmax * 2 : nullptr;
        ba_carve_lists<EXACT>(c, day, plan.max_ticks, plan.max_dslice, fixed, pool, gscr);

        // ---- setup -----------------------------------------------------------
        const float *lat_airport = prm.lat_airport + (size_t)p * C * N;
        for (int a = tid; a < N; a += nthr) ba_init_airport<EXACT, TRK>(c, a, lat_airport, N);
        if (c.grow)
            for (int i = tid; i < plan.grad_stride; i += nthr) c.grow[i] = 0.0f;
        if (c.pop_tick)
            for (int i = tid; i < 2 * plan.Fmax; i += nthr) c.pop_tick[i] = -1;
        BaAcc acc = {0.0, 0.0f, 0.0f, 0.0f, 0u};
        for (uint32_t k = tid; k < c.cal_cap; k += nthr) c.cal_cnt[k] = 0;
        if (tid < BA_CTA_WORDS) c.cta[tid] = 0;
        __syncthreads();

        BA_MARK(0);
        // ---- prologue: flight-parallel, coalesced ------------------------------
        for (int f = tid; f < day.F; f += nthr) ba_prologue_flight<PREDICT, TRK>(c, acc, f);
        __syncthreads();
        BA_MARK(1);
        if (!PREDICT) {
            if (tid < 32) ba_calendar_scan(c, tid);
            __syncthreads();
            for (int f = tid; f < day.F; f += nthr) ba_calendar_scatter(c, f);
        }
        // Staging runs ahead of the scan (the records are static once the prologue is done):
        // arrival-calendar slice k -> due buffer k & 1 during tick k - 1, departure-calendar
        // slice k -> staging buffer (k - 1) & 1 during tick k - 2.
        int tick = day.first_tick - 1;
        __syncthreads();                                        // calendar records are written
        auto stage_departures = [&](int k, uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo + tid; i < hi; i += nthr)
                ba_stage_departure<EXACT>(c, acc, i, lo, (k - 1) & 1);
        };
        // calendar offsets are fetched a tick before they are needed (no load on the path)
        auto dcal_at = [&](int k) {
            return (uint32_t)day.dcal_off[k <= day.last_ready_tick + 1 ? k : day.last_ready_tick + 1];
        };
        const int a0 = tid < N ? day.lane_airport[tid] : 0;     // this thread's first airport
        bool busy = day.F > 0;
        if (!EXACT && !PREDICT) {
            // ---- tick loop (model.py:158-200), ONE barrier per tick (ba_sim.cuh: single-
            //      barrier scan).  During tick j every thread first issues the asynchronous
            //      copies that assemble the due buffer of tick j and stage what tick j + 1
            //      will read, then runs its airport lane: phase C of tick j - 1, phase A of
            //      tick j.  Nothing a thread reads in tick j is written in tick j. ----------
            const uint32_t cap = (uint32_t)BA_DUE_CAP_S;
            // the copy-issuing work goes to the threads in reverse order: the last warp holds
            // the fewest (and lightest) airports, so it has the time
            const uint32_t rtid = (uint32_t)(nthr - 1 - tid);
            auto slice_lo = [&](int k) { return k >= 1 && (uint32_t)k < c.cal_cap ? c.cal_cnt[k - 1] : 0u; };
            auto slice_n = [&](int k) {
                if (k < 1 || (uint32_t)k >= c.cal_cap) return 0u;
                const uint32_t n = c.cal_cnt[k] - c.cal_cnt[k - 1];
                return n > cap ? cap : n;
            };
            auto stage_words = [&](int k) {
                if (k < 1 || (uint32_t)k >= c.cal_cap) return;
                const uint32_t lo = c.cal_cnt[k - 1], n = c.cal_cnt[k] - lo;
                if (n > cap) acc.status |= BA_S_OVERFLOW;       // more arrivals in one tick than slots
                for (uint32_t i = lo + rtid; i < lo + (n > cap ? cap : n); i += nthr) ba_stage_word(c, i, lo, k & 1);
            };
            uint32_t d_lo = dcal_at(tick + 1), d_hi = dcal_at(tick + 2), d_nx = dcal_at(tick + 3);
            stage_departures(tick + 1, d_lo, d_hi);             // slice tick + 1 -> buffer tick & 1
            stage_words(tick + 1);
            ba_copy_wait();
            __syncthreads();
            BA_MARK(2);
            float t_next = c.tick_time[tick + 1];
            while (busy) {
                if (tick >= plan.max_ticks) { acc.status |= BA_S_TICK_CAP; break; }
                ++tick;
                const float t = t_next;                         // model.py:161
                t_next = c.tick_time[tick + 1];
                d_lo = d_hi; d_hi = d_nx; d_nx = dcal_at(tick + 3);   // slice tick + 1 = [d_lo, d_hi)
                const int prev = (tick - 1) & 1, cur = tick & 1;
                const uint32_t n_cur = slice_n(tick), n_prev = slice_n(tick - 1);
                uint32_t n_late = c.cta[8 + (tick - 1) % 3];
                if (n_late > cap - n_prev) n_late = cap - n_prev;   // overflow already flagged
                if (tid == 0) c.cta[8 + (tick + 1) % 3] = 0;    // read two ticks ago, next used next tick
                // (1) this tick's arrival slice -> due buffer `cur` (read next tick)
                {
                    const uint32_t lo = slice_lo(tick);
                    for (uint32_t i = rtid; i < n_cur; i += nthr) ba_complete_arrival(c, i, lo, cur);
                }
                // (2) what the next tick reads: its slice's queue words, its departures
                stage_words(tick + 1);
                for (uint32_t i = d_lo + rtid; i < d_hi; i += nthr)     // slice tick + 1 -> buffer `cur`
                    ba_stage_departure<EXACT>(c, acc, i, d_lo, cur);
                // (3) the airport lanes
                int more = tick < day.last_ready_tick;
                for (int s = tid; s < N; s += nthr) {
                    const int a = s == tid ? a0 : day.lane_airport[s];
                    ba_phase_c_pull_fast(c, acc, a, prev, n_prev, n_late, tick);
                    ba_phase_a<EXACT, PREDICT, TRK>(c, acc, a, t, tick, n_cur);
                    more |= ba_airport_busy<TRK>(c, a) ? 1 : 0;
                }
                if (acc.status & BA_S_OVERFLOW) c.cta[3] = 1u;
                ba_copy_wait();
                busy = __syncthreads_or(more) != 0;
                BA_MARK(3);
                if (c.cta[3]) break;                            // re-run with exact lists anyway
            }
            ba_copy_wait();
        } else {
            auto stage_arrivals = [&](int k) {                      // returns nothing; counter by tid 0
                uint32_t n = 0;
                if (!PREDICT && (uint32_t)k < c.cal_cap) {
                    const uint32_t lo = c.cal_cnt[k - 1], hi = c.cal_cnt[k];
                    n = hi - lo;
                    if (n > c.due_cap) { n = c.due_cap; acc.status |= BA_S_OVERFLOW; }
                    for (uint32_t i = lo + tid; i < lo + n; i += nthr) ba_stage_arrival<EXACT>(c, i, lo, k & 1);
                }
                if (tid == 0) c.cta[k & 1] = n;                     // late entries go behind the slice
            };
            uint32_t d_lo = dcal_at(tick + 1), d_mid = dcal_at(tick + 2), d_hi = dcal_at(tick + 3);
            stage_departures(tick + 1, d_lo, d_mid);
            stage_departures(tick + 2, d_mid, d_hi);
            stage_arrivals(tick + 1);
            ba_copy_wait();
            __syncthreads();
            BA_MARK(2);

            // ---- tick loop (model.py:158-200): two barriers per tick ------------------
            //   interval 1 (one lane per airport): phase C of the previous tick, then phase A
            //   interval 2 (flight-parallel):      phase B, completes the due buffer of this tick
            float t_next = c.tick_time[tick + 1];
            while (busy) {
                if (tick >= plan.max_ticks) { acc.status |= BA_S_TICK_CAP; break; }
                ++tick;
                const float t = t_next;                             // model.py:161
                t_next = c.tick_time[tick + 1];
                d_lo = d_hi; d_hi = dcal_at(tick + 3);              // slice tick + 2, used in interval 2
                const int prev = (tick - 1) & 1, cur = tick & 1;
                uint32_t n_due = c.cta[prev];
                if (n_due > c.due_cap) n_due = c.due_cap;           // overflow already flagged
                int more = tick < day.last_ready_tick;
                for (int s = tid; s < N; s += nthr) {
                    const int a = s == tid ? a0 : day.lane_airport[s];
                    ba_phase_c_pull<EXACT>(c, acc, a, prev, n_due, tick);
                    ba_phase_a<EXACT, PREDICT, TRK>(c, acc, a, t, tick);
                    more |= ba_airport_busy<TRK>(c, a) ? 1 : 0;
                }
                if (!EXACT && (acc.status & BA_S_OVERFLOW)) c.cta[3] = 1u;
                ba_copy_wait();                                     // copies issued a tick ago
                __syncthreads();
                BA_MARK(3);
                if (!EXACT && c.cta[3]) break;                      // a ring overflowed: this
                                                                    // particle-day is re-run anyway
                for (int s = tid; s < N; s += nthr)
                    ba_run_advance(c, s == tid ? a0 : day.lane_airport[s], tick);
                if (PREDICT) {
                    const uint32_t lo = c.cta[4], hi = c.cta[2];
                    for (uint32_t i = lo + tid; i < hi; i += nthr) ba_phase_b_pool(c, acc, i, t, cur);
                } else if ((uint32_t)tick < c.cal_cap) {
                    uint32_t n = c.cal_cnt[tick] - c.cal_cnt[tick - 1];
                    if (n > c.due_cap) n = c.due_cap;
                    for (uint32_t i = tid; i < n; i += nthr) ba_phase_b_arrival(c, i, cur);
                }
                stage_arrivals(tick + 1);                           // due buffer `prev`: consumed
                stage_departures(tick + 2, d_lo, d_hi);             // staging buffer `prev`: consumed
                busy = __syncthreads_or(more) != 0;
                if (PREDICT && tid == 0) ba_pool_advance(c, 64u);   // next reader: phase B, after a barrier
                BA_MARK(4);
            }
            ba_copy_wait();

        }

        // ---- epilogue ----------------------------------------------------------
        __syncthreads();          // every wait of the tape is written
        const bool aborted = PREDICT || (!EXACT && c.cta[3] != 0u);   // predictive: outcomes only
        if (!aborted)
            for (int f = tid; f < day.F; f += nthr) ba_epilogue_flight<TRK>(c, acc, f);
        BA_MARK(5);
        if (c.grow && !aborted) {
            __syncthreads();
            for (int s = tid; s < N; s += nthr) ba_epilogue_airport<TRK>(c, day.lane_airport[s], N);
            for (int r = tid; r < day.Rd; r += nthr) ba_epilogue_route(c, r);
        }
        // deterministic block reduction: lanes -> warp (shuffle tree) -> warps in order
        double v0 = acc.lp;
        float v1 = acc.g_sigma, v2 = acc.g_vt, v3 = acc.g_vu;
        uint32_t st = acc.status;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_down_sync(0xffffffffu, v0, o);
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
            v2 += __shfl_down_sync(0xffffffffu, v2, o);
            v3 += __shfl_down_sync(0xffffffffu, v3, o);
            st |= __shfl_down_sync(0xffffffffu, st, o);
        }
        float *redf = (float *)red;
        double *redd = (double *)red;          // red is 16-byte aligned; 6 words per warp
        if (lane == 0) {
            redd[warp * 3 + 0] = v0;
            redf[warp * 6 + 2] = v1; redf[warp * 6 + 3] = v2;
            redf[warp * 6 + 4] = v3; red[warp * 6 + 5] = st;
        }
        __syncthreads();
        if (tid == 0) {
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            uint32_t ss = 0;
            for (int w = 0; w < nwarp; ++w) {
                s0 += redd[w * 3 + 0]; s1 += redf[w * 6 + 2]; s2 += redf[w * 6 + 3];
                s3 += redf[w * 6 + 4]; ss |= red[w * 6 + 5];
            }
            prm.logp[pd] = (float)s0;
            prm.status[pd] = ss;
            if (prm.ticks) prm.ticks[pd] = tick;
            if (c.grow) {
                c.grow[c.route_base + R + 0] = (float)s1;
                c.grow[c.route_base + R + 1] = (float)s2;
                c.grow[c.route_base + R + 2] = (float)s3;
            }
        }
        __syncthreads();     // red[] and the state arrays are reused by the next item
        BA_MARK(6);
    }
}


// ---------------------------------------------------------------------------
// ba_forward2_kernel<NW>: the ring-less scan (ba_scan2.cuh) for bookkeeping-free, fully
// observed plans.  Persistent small CTAs: NW warps (1 for networks up to ~128 airports) own a
// (particle, day) from the prologue to the epilogue, so an SM runs ~30 independent
// particle-days.  With NW = 1 nothing ever waits for another warp; larger networks spread the
// rounds of a tick over NW warps (one CTA barrier per tick) so that an SM still holds ~30
// warps although a particle-day's state takes more shared memory.
// ---------------------------------------------------------------------------
#define BA2_RED_WORDS (8 * 6 + 8)

template <int NW> __device__ __forceinline__ void ba2_sync()
{
    if (NW == 1) __syncwarp(); else __syncthreads();
}
template <int NW> __device__ __forceinline__ bool ba2_any(bool v)
{
    if (NW == 1) return __any_sync(0xffffffffu, v) != 0;
    return __syncthreads_or(v ? 1 : 0) != 0;
}
// rounds of 32 airports are dealt to the warps in snake order (0 1 2 3 3 2 1 0 ...): the
// airports are sorted by traffic, so every warp gets heavy and light rounds
template <int NW> __device__ __forceinline__ bool ba2_my_round(int j, int warp)
{
    if (NW == 1) return true;
    const int pos = j % (2 * NW);
    return (pos < NW ? pos : 2 * NW - 1 - pos) == warp;
}

template <int NW>
__global__ void __launch_bounds__(32 * NW, 32 / NW) ba_forward2_kernel(const BaFwdParams prm)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NT = 32 * NW;
    const BaPlanView &plan = prm.plan;
    const int N = plan.N, D = plan.D, R = plan.R;

    Ba2Ctx c;
    c.nslot = (uint32_t)N;
    c.un = smem + BA2_RED_WORDS;
    uint32_t *red = smem;
    uint32_t *gscr = prm.scratch + (size_t)blockIdx.x * prm.scratch_stride;

    c.sigma = prm.scales[0]; c.v_t = prm.scales[1]; c.v_u = prm.scales[2];
    c.inv_sigma = 1.0f / c.sigma;
    c.inv_sigma2 = c.inv_sigma * c.inv_sigma;
    c.inv_sigma3 = c.inv_sigma2 * c.inv_sigma;
    c.log_sigma = logf(c.sigma);
    c.inv_vt = 1.0f / c.v_t; c.inv_vu = 1.0f / c.v_u;
    c.c0_lp = ba_relaxed_bernoulli_lp(0.0f, 0.0f, nullptr);
    c.c1_lp = ba_relaxed_bernoulli_lp(0.0f, 1.0f, nullptr);
    c.value_mode = prm.value_mode; c.want_grad = prm.want_grad && prm.tape != nullptr;
    c.seed = prm.seed_dev ? *prm.seed_dev : prm.seed;
    c.route_base = 2 * N;
    c.tick_time = plan.tick_time;
    c.max_ticks = plan.max_ticks;
    c.N = N;
    // optional phase timing (BA_PHASE_PROFILE)
    long long pc_last = 0;
    const bool prof = prm.phase_cycles != nullptr && tid == 0;
#define BA2_MARK(slot) do { if (prof) { long long now_ = clock64(); atomicAdd(prm.phase_cycles + (slot), (unsigned long long)(now_ - pc_last)); pc_last = now_; } } while (0)

    for (;;) {
        if (prof) pc_last = clock64();
        // ---- next work item ------------------------------------------------
        int wi = 0;
        if (NW == 1) {
            if (lane == 0) wi = (int)atomicAdd(prm.work_counter, 1u);
            wi = __shfl_sync(0xffffffffu, wi, 0);
        } else {
            if (tid == 0) red[BA2_RED_WORDS - 1] = atomicAdd(prm.work_counter, 1u);
            __syncthreads();
            wi = (int)red[BA2_RED_WORDS - 1];
        }
        const int n_work = prm.n_work_dev ? (int)*prm.n_work_dev : prm.n_work;
        if (wi >= n_work) break;
        int p, d;
        if (prm.work_list) { const int pd = prm.work_list[wi]; p = pd / D; d = pd - p * D; }
        else { const int g = wi + prm.wi_begin; d = g / prm.P; p = g - d * prm.P; }   // same-day items adjacent in time
        const size_t pd = (size_t)p * D + d;
        const BaDayView day = plan.days[d];

        c.F = day.F;
        c.flights = day.flights; c.route = day.route; c.ap = day.airports;
        c.pos_live = day.pos_live; c.pos_dep = day.pos_dep; c.pos_arr = day.pos_arr;
        c.rt_gid = day.rt_gid; c.rt_off = day.rt_off; c.rt_flt = day.rt_flt; c.Rd = day.Rd;
        c.lq = day.lq; c.drun = day.drun; c.drun_off = day.drun_off; c.order = day.lane_sorted;
        c.first_tick = day.first_tick;
        c.lat_airport = prm.lat_airport + (size_t)p * 2 * N;
        c.lat_route = prm.lat_route + (size_t)p * R;
        c.noise = prm.noise ? prm.noise + pd * (size_t)plan.Fmax * 4 : nullptr;
        c.particle = (uint32_t)p + prm.particle_offset; c.dayidx = (uint32_t)d + prm.day_offset;
        c.grow = c.want_grad ? prm.tape + pd * (size_t)plan.grad_stride : nullptr;
        c.pop_tick = prm.pop_tick ? prm.pop_tick + pd * (size_t)plan.Fmax * 2 : nullptr;
        c.vrow = prm.vtape ? prm.vtape + pd * (size_t)plan.Fmax * 4 : nullptr;
        ba2_carve_scratch(c, gscr, day.F, N, R);
        const uint32_t n_arr = (uint32_t)day.n_arr;

        // ---- setup -----------------------------------------------------------
        c.s_rate = (float *)(c.un + BA2_KC + N); c.s_mt = c.s_rate + N;   // behind buckets and cursors
        for (int a = tid; a < N; a += NT) ba2_init_airport(c, a, N, false);
        for (int i = tid; i < BA2_KC; i += NT) c.un[i] = 0;
        if (c.pop_tick)
            for (int i = 2 * day.F + tid; i < 2 * plan.Fmax; i += NT) c.pop_tick[i] = -1;
        if (c.vrow)
            for (int i = 4 * day.F + tid; i < 4 * plan.Fmax; i += NT) c.vrow[i] = 0.0f;
        BaAcc acc = {0.0, 0.0f, 0.0f, 0.0f, 0u};
        ba2_sync<NW>();
        BA2_MARK(0);

        // ---- prologue ----------------------------------------------------------
        // two flights per lane and step: both flights' loads are in flight together
        for (int f = tid; f < day.F; f += 2 * NT) {
            const int f1 = f + NT;
            const bool two = f1 < day.F;
            const Ba2FlightIn in0 = ba2_prologue_load(c, f);
            const Ba2FlightIn in1 = ba2_prologue_load(c, two ? f1 : f);
            const float T0 = c.lat_route[in0.rt], T1 = c.lat_route[in1.rt];
            ba2_prologue_compute(c, acc, f, in0, T0);
            if (two) ba2_prologue_compute(c, acc, f1, in1, T1);
        }
        ba2_sync<NW>();
        BA2_MARK(1);
        if (warp == 0) ba2_bucket_scan(c, lane);
        ba2_sync<NW>();
        for (int f = tid; f < day.F; f += NT) ba2_scatter_flight(c, f);
        for (int a = tid; a < N; a += NT) ba2_split_cursor(c, a);
        ba2_sync<NW>();
        // stable split by destination: all warps when their cursor tables fit behind the
        // prologue's shared tables, else one warp walks the order
        if (NW > 1 && (size_t)(BA2_KC + 3 * N + NW * N) <= ba2_smem_words(N)) {
            uint32_t *cnt = c.un + BA2_KC + 3 * N;
            for (int i = tid; i < NW * N; i += NT) cnt[i] = 0u;
            __syncthreads();
            ba2_split_count(c, lane, warp, NW, cnt, n_arr);
            __syncthreads();
            for (int a = tid; a < N; a += NT) ba2_split_prefix(c, a, NW, cnt);
            __syncthreads();
            ba2_split_walk(c, lane, warp, NW, cnt, n_arr);
        } else if (warp == 0) ba2_split(c, lane, n_arr);
        const bool aborted = ba2_any<NW>((acc.status & BA_S_OVERFLOW) != 0u);
        BA2_MARK(2);

        // ---- scan: the lanes walk the airports in rounds, one clock ------------------
        uint32_t k = (uint32_t)(day.first_tick - 1);
        if (!aborted) {
            uint32_t nleft = 0;
            for (uint32_t slot = tid; slot < (uint32_t)N; slot += NT) {
                const int a = c.order[slot];
                ba2_state_init(c, slot, a);
            }
            for (int j = 0; 32 * j < N; ++j) {
                const uint32_t slot = 32u * (uint32_t)j + (uint32_t)lane;
                if (!ba2_my_round<NW>(j, warp) || slot >= (uint32_t)N) continue;
                const int a = c.order[slot];
                nleft += (c.ap[a].dep_live + c.ap[a].arr_cnt) != 0 ? 1u : 0u;
            }
            bool go = ba2_any<NW>(nleft != 0u);
            while (go) {
                if ((int)k >= plan.max_ticks) { acc.status |= BA_S_TICK_CAP; break; }
                ++k;
                const float t = c.tick_time[k];                 // model.py:161
                for (int j = 0; 32 * j < N; ++j) {
                    const uint32_t slot = 32u * (uint32_t)j + (uint32_t)lane;
                    if (!ba2_my_round<NW>(j, warp) || slot >= (uint32_t)N) continue;
                    if (ba2_airport_tick(c, slot, k, t, acc.status)) nleft -= 1u;
                }
                // every airport's tick k is in memory before tick k + 1
                go = ba2_any<NW>(nleft != 0u);
            }
        }
        BA2_MARK(3);
        int tick = (int)k;
        if (day.F > 0 && tick < day.last_ready_tick) tick = day.last_ready_tick;

        // ---- epilogue ----------------------------------------------------------
        ba2_sync<NW>();
        c.s_rate = (float *)c.un; c.s_mt = c.s_rate + N;          // the scan state is dead
        c.accA = (unsigned long long *)(c.un + ((2 * N + 1) & ~1));
        for (int a = tid; a < N; a += NT) ba2_init_airport(c, a, N, true);
        if (c.grow) {
            for (int i = tid; i < 3 * N; i += NT) c.accA[i] = 0ull;
            for (int r = tid; r < R; r += NT) c.accR[r] = 0ull;
        }
        ba2_sync<NW>();
        if (!aborted)
            for (int f = tid; f < day.F; f += 2 * NT) {
                const int f1 = f + NT;
                const bool two = f1 < day.F;
                const Ba2EpiIn in0 = ba2_epilogue_load(c, f);
                const Ba2EpiIn in1 = ba2_epilogue_load(c, two ? f1 : f);
                const Ba2Pay pay0 = ba2_epilogue_pay(c, in0), pay1 = ba2_epilogue_pay(c, in1);
                const float T0 = c.lat_route[in0.rt], T1 = c.lat_route[in1.rt];
                ba2_epilogue_compute(c, acc, f, in0, pay0, T0);
                if (two) ba2_epilogue_compute(c, acc, f1, in1, pay1, T1);
            }
        if (c.grow && !aborted) {
            ba2_sync<NW>();
            for (int s = tid; s < N; s += NT) ba2_epilogue_airport(c, s, N);
            for (int r = tid; r < R; r += NT) ba2_epilogue_route(c, r);
        }
        if (c.vrow && !aborted) {
            // value mode: wait-mediated part of the gradient w.r.t. the service-time values
            ba2_sync<NW>();
            for (int s = tid; s < N; s += NT) ba2_value_grad_airport(c, s);
            ba2_sync<NW>();
            for (int f = tid; f < day.F; f += NT) ba2_value_grad_finish(c, f);
        }
        // deterministic reduction: shuffle tree over the lanes, warps in order
        double v0 = acc.lp;
        float v1 = acc.g_sigma, v2 = acc.g_vt, v3 = acc.g_vu;
        uint32_t st = acc.status;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_down_sync(0xffffffffu, v0, o);
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
            v2 += __shfl_down_sync(0xffffffffu, v2, o);
            v3 += __shfl_down_sync(0xffffffffu, v3, o);
            st |= __shfl_down_sync(0xffffffffu, st, o);
        }
        if (NW > 1) {
            float *redf = (float *)red;
            double *redd = (double *)red;          // 16-byte aligned; 6 words per warp
            if (lane == 0) {
                redd[warp * 3 + 0] = v0;
                redf[warp * 6 + 2] = v1; redf[warp * 6 + 3] = v2;
                redf[warp * 6 + 4] = v3; red[warp * 6 + 5] = st;
            }
            __syncthreads();
            if (tid == 0) {
                v0 = 0; v1 = v2 = v3 = 0; st = 0;
                for (int w = 0; w < NW; ++w) {
                    v0 += redd[w * 3 + 0]; v1 += redf[w * 6 + 2]; v2 += redf[w * 6 + 3];
                    v3 += redf[w * 6 + 4]; st |= red[w * 6 + 5];
                }
            }
        }
        if (tid == 0) {
            prm.logp[pd] = (float)v0;
            prm.status[pd] = st;
            if (prm.ticks) prm.ticks[pd] = tick;
            if (c.grow) {
                c.grow[c.route_base + R + 0] = v1;
                c.grow[c.route_base + R + 1] = v2;
                c.grow[c.route_base + R + 2] = v3;
            }
        }
        ba2_sync<NW>();        // the shared tables are reused by the next item
        BA2_MARK(4);
    }
#undef BA2_MARK
}

size_t ba_forward2_smem_bytes(int n_airports) { return (BA2_RED_WORDS + ba2_smem_words(n_airports)) * 4; }

uint64_t ba_forward2_scratch_words(int Fmax, int n_airports, int n_routes) { return ba2_scratch_words(Fmax, n_airports, n_routes); }

// warps per particle-day: one up to 128 airports; beyond that the state takes more shared
// memory per particle-day, and more warps keep the SM's warp count up
int ba_forward2_warps(int n_airports)
{
    if (const char *v = getenv("BA_WARPS")) { const int w = atoi(v); if (w == 1 || w == 2 || w == 4 || w == 8) return w; }
    // (measured on B200: N = 100 nominal 1.32e6 / 1.34e6 pd/s at 1 / 2 warps, congested
    //  7.9e5 / 7.0e5; N = 500: 0.88 / 1.23 / 1.47 / 1.38 e5 at 1 / 2 / 4 / 8 warps)
    return n_airports <= 128 ? 1 : (n_airports <= 256 ? 2 : 4);
}

#define BA2_DISPATCH(NWV, STMT)                                              \
    do {                                                                     \
        if ((NWV) == 1) { auto kfn = ba_forward2_kernel<1>; STMT; }           \
        else if ((NWV) == 2) { auto kfn = ba_forward2_kernel<2>; STMT; }      \
        else if ((NWV) == 4) { auto kfn = ba_forward2_kernel<4>; STMT; }      \
        else { auto kfn = ba_forward2_kernel<8>; STMT; }                      \
    } while (0)

// sets the dynamic shared-memory limit (on the current device) and returns the resident grid
int ba_forward2_configure(int nw, size_t smem, int device)
{
    // the limit is a property of the kernel function on one device, shared by every plan of
    // the process there: only ever raise it
    static size_t cur[BA_MAX_DEVICES][4] = {};
    const int cls = nw == 1 ? 0 : (nw == 2 ? 1 : (nw == 4 ? 2 : 3));
    cudaError_t e = cudaSuccess;
    if (device >= 0 && device < BA_MAX_DEVICES && smem > cur[device][cls] && smem > 48 * 1024) {
        BA2_DISPATCH(nw, e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (e != cudaSuccess) return -1;
        cur[device][cls] = smem;
    }
    int per_sm = 0, sms = 0;
    BA2_DISPATCH(nw, e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, 32 * nw, smem));
    if (e != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
    if (const char *v = getenv("BA_MAX_CTAS_PER_SM")) { const int cap = atoi(v); if (cap > 0 && cap < per_sm) per_sm = cap; }
    return per_sm * sms;
}

cudaError_t ba_launch_forward2(const BaFwdParams &p, int nw, int grid, size_t smem, cudaStream_t s)
{
    BA2_DISPATCH(nw, (kfn<<<grid, 32 * nw, smem, s>>>(p)));
    return cudaGetLastError();
}

// g[p, j] = sum_d dlogp[p, d] * tape[p, d, j]
__global__ void __launch_bounds__(256) ba_backward_kernel(const BaBwdParams prm)
{
    const int CN = prm.C * prm.N;
    const int stride = prm.stride;
    for (int p = blockIdx.y; p < prm.P; p += gridDim.y) {
        const float *tp = prm.tape + (size_t)p * prm.D * stride;
        const float *dl = prm.dlogp + (size_t)p * prm.D;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < stride; j += gridDim.x * blockDim.x) {
            float s = 0.0f;
            for (int d = 0; d < prm.D; ++d) s = fmaf(__ldg(dl + d), __ldg(tp + (size_t)d * stride + j), s);
            if (j < CN) prm.g_airport[(size_t)p * CN + j] = s;
            else if (j < CN + prm.R) prm.g_route[(size_t)p * prm.R + (j - CN)] = s;
            else prm.g_scales[(size_t)p * 3 + (j - CN - prm.R)] = s;
        }
    }
}

__global__ void __launch_bounds__(256) ba_noise_kernel(int P, int D, int Fmax, uint64_t seed,
                                                      float4 *noise)
{
    const size_t total = (size_t)P * D * Fmax;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t f = (uint32_t)(i % Fmax);
        const size_t pd = i / Fmax;
        const uint32_t d = (uint32_t)(pd % D), p = (uint32_t)(pd / D);
        const BaNoise4 z = ba_philox_noise(seed, p, d, f);
        noise[i] = make_float4(z.e_dep, z.eps_travel, z.e_arr, z.eps_turn);
    }
}

__global__ void __launch_bounds__(256) ba_collect_kernel(const uint32_t *status, int n,
                                                        uint32_t bit, int32_t *list,
                                                        unsigned int *count)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        if (status[i] & bit) list[atomicAdd(count, 1u)] = i;
}

// ---------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------
// block-size classes: the register budget follows the CTA size actually used
#define BA_DISPATCH_T(EX, TRK, BLOCK, STMT)                                                 \
    do {                                                                                    \
        if ((BLOCK) <= 128) { auto kfn = ba_forward_kernel<EX, 128, false, TRK>; STMT; }      \
        else if ((BLOCK) <= 256) { auto kfn = ba_forward_kernel<EX, 256, false, TRK>; STMT; } \
        else if ((BLOCK) <= 512) { auto kfn = ba_forward_kernel<EX, 512, false, TRK>; STMT; } \
        else { auto kfn = ba_forward_kernel<EX, 1024, false, TRK>; STMT; }                    \
    } while (0)
// TRKV (runtime): does any airport of the plan keep aircraft books (ba_sim.cuh: ba_ap_flags)
#define BA_DISPATCH(EX, TRKV, BLOCK, STMT)                                                  \
    do {                                                                                    \
        if (TRKV) BA_DISPATCH_T(EX, true, BLOCK, STMT); else BA_DISPATCH_T(EX, false, BLOCK, STMT); \
    } while (0)

// The dynamic-shared-memory limit is a property of the kernel function, shared by every plan
// of the process: only ever raise it (a later, smaller plan must not lower it under an
// earlier one).
cudaError_t ba_forward_configure(int block, size_t smem_fast, size_t smem_exact, bool tracked)
{
    // (cudaFuncSetAttribute applies to the current device: one cache per device)
    static size_t cur_dev[BA_MAX_DEVICES][2][2][4] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= BA_MAX_DEVICES) dev = 0;
    size_t (*cur)[4] = cur_dev[dev][tracked ? 1 : 0];
    const int cls = block <= 128 ? 0 : (block <= 256 ? 1 : (block <= 512 ? 2 : 3));
    cudaError_t e = cudaSuccess;
    // tuning knob: share of the SM's L1/shared array given to shared memory (percent); the
    // rest caches the per-CTA scratch.  Default: the driver's choice for full occupancy.
    static int carve = -2;
    if (carve == -2) { const char *v = getenv("BA_CARVEOUT"); carve = v ? atoi(v) : -1; }
    if (carve >= 0) {
        BA_DISPATCH(false, tracked, block, e = cudaFuncSetAttribute(
            kfn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        if (e != cudaSuccess) return e;
    }
    if (smem_fast > cur[0][cls]) {
        BA_DISPATCH(false, tracked, block, e = cudaFuncSetAttribute(
            kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
        if (e != cudaSuccess) return e;
        cur[0][cls] = smem_fast;
    }
    if (smem_exact > cur[1][cls]) {
        BA_DISPATCH(true, tracked, block, e = cudaFuncSetAttribute(
            kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_exact));
        if (e != cudaSuccess) return e;
        cur[1][cls] = smem_exact;
    }
    return e;
}

int ba_forward_max_resident(int block_threads, size_t smem_bytes, bool exact, bool tracked, int device)
{
    int per_sm = 0, sms = 0;
    cudaError_t e;
    if (exact)
        BA_DISPATCH(true, tracked, block_threads, e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, kfn, block_threads, smem_bytes));
    else
        BA_DISPATCH(false, tracked, block_threads, e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &per_sm, kfn, block_threads, smem_bytes));
    if (e != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
        return -1;
    if (const char *v = getenv("BA_MAX_CTAS_PER_SM")) { const int cap = atoi(v); if (cap > 0 && cap < per_sm) per_sm = cap; }
    return per_sm * sms;
}

cudaError_t ba_launch_forward(const BaFwdParams &p, bool exact, bool tracked, int grid, int block,
                              size_t smem, cudaStream_t s)
{
    if (exact) BA_DISPATCH(true, tracked, block, (kfn<<<grid, block, smem, s>>>(p)));
    else BA_DISPATCH(false, tracked, block, (kfn<<<grid, block, smem, s>>>(p)));
    return cudaGetLastError();
}

// posterior-predictive variant: exact-capacity lists only
#define BA_DISPATCH_PREDICT(BLOCK, STMT)                                                 \
    do {                                                                                 \
        if ((BLOCK) <= 128) { auto kfn = ba_forward_kernel<true, 128, true, true>; STMT; }     \
        else if ((BLOCK) <= 256) { auto kfn = ba_forward_kernel<true, 256, true, true>; STMT; } \
        else if ((BLOCK) <= 512) { auto kfn = ba_forward_kernel<true, 512, true, true>; STMT; } \
        else { auto kfn = ba_forward_kernel<true, 1024, true, true>; STMT; }                   \
    } while (0)

cudaError_t ba_launch_predict(const BaFwdParams &p, int grid, int block, size_t smem,
                              cudaStream_t s)
{
    static size_t cur_dev[BA_MAX_DEVICES][4] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= BA_MAX_DEVICES) dev = 0;
    size_t *cur = cur_dev[dev];
    const int cls = block <= 128 ? 0 : (block <= 256 ? 1 : (block <= 512 ? 2 : 3));
    if (smem > cur[cls]) {
        cudaError_t e = cudaSuccess;
        BA_DISPATCH_PREDICT(block, e = cudaFuncSetAttribute(
            kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (e != cudaSuccess) return e;
        cur[cls] = smem;
    }
    BA_DISPATCH_PREDICT(block, (kfn<<<grid, block, smem, s>>>(p)));
    return cudaGetLastError();
}

cudaError_t ba_launch_backward(const BaBwdParams &p, cudaStream_t s)
{
    dim3 grid((unsigned)((p.stride + 255) / 256), (unsigned)(p.P < 65535 ? p.P : 65535));
    if (grid.x > 64) grid.x = 64;
    ba_backward_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

// g_noise[pd, f, :] = dlogp[pd] * vtape[pd, f, :]
__global__ void __launch_bounds__(256) ba_backward_values_kernel(const float *__restrict__ dlogp,
                                                                 const float4 *__restrict__ vtape,
                                                                 float4 *__restrict__ g_noise,
                                                                 long long n_pd, int row)
{
    const long long total = n_pd * row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const float w = __ldg(dlogp + i / row);
        const float4 v = vtape[i];
        g_noise[i] = make_float4(w * v.x, w * v.y, w * v.z, w * v.w);
    }
}

cudaError_t ba_launch_backward_values(const float *dlogp, const float *vtape, float *g_noise,
                                      long long n_pd, int row, cudaStream_t s)
{
    const long long total = n_pd * row;
    int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
    if (grid < 1) grid = 1;
    ba_backward_values_kernel<<<grid, 256, 0, s>>>(dlogp, (const float4 *)vtape, (float4 *)g_noise, n_pd, row);
    return cudaGetLastError();
}

cudaError_t ba_launch_noise(const BaPlanView &plan, int P, uint64_t seed, float *noise,
                            cudaStream_t s)
{
    const size_t total = (size_t)P * plan.D * plan.Fmax;
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    ba_noise_kernel<<<grid, 256, 0, s>>>(P, plan.D, plan.Fmax, seed, (float4 *)noise);
    return cudaGetLastError();
}

cudaError_t ba_launch_collect(const uint32_t *status, int n, uint32_t bit, int32_t *list,
                              unsigned int *count, cudaStream_t s)
{
    int grid = (n + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    ba_collect_kernel<<<grid, 256, 0, s>>>(status, n, bit, list, count);
    return cudaGetLastError();
}
