// Ring-less scan ("scan2") of the hot path for plans without aircraft bookkeeping and with
// every flight observed -- the training configurations (BASELINE configs[1], [2], [4]).
//
// ONE WARP PER (particle, day).  Same three stages as ba_sim.cuh (prologue / scan /
// epilogue) and the same reference semantics (bayes_air/model.py:158-200,
// network.py:43-198, types/airport.py:86-221), but
//
//   * The 32 lanes walk the airports in rounds (airport 32 r + lane in round r, heaviest
//     airports first), all on one clock; the per-airport state (ten packed words) lives in
//     shared memory between ticks.  A particle-day needs ~5 KB of shared memory and 64
//     registers x 32 threads, so an SM holds ~30 of them: 30 independent serial chains per
//     SM hide each other's memory latency, and no warp ever waits for another one (no
//     barrier, no progress words, no spinning).
//   * The runway queue of an airport is VIRTUAL.  Its departures join in a static order
//     (pending order of its live departures; BaDayView::lq), its arrivals in in-transit list
//     order.  FIFO order is the merge of the two sequences by joining tick (arrivals of the
//     end of tick k-1 before the departures of tick k, model.py:181-196), so the queue is two
//     read pointers and nothing is ever copied into a ring.
//   * Per flight there is ONE 16-byte payload record, indexed by the list position of its
//     departure ("lpos"): {departure service draw, arrival time, arrival service draw, tick
//     at which the departure was served}.  The draws are overwritten by the total waits when
//     the entries are served (the tape of the gradient), so the record is prologue output,
//     cross-airport message (q) and tape at once.
//   * Arrivals: the prologue leaves, per destination, the list of its arrivals as 4-byte
//     keys (arrival tick << 16 | lpos) SORTED BY ARRIVAL TICK (counting sort by tick, then a
//     stable split by destination).  At tick k the airport takes the keys whose tick has
//     come and whose departure has been served (q <= k-1), orders that small block by
//     (q, lpos) = the reference's in-transit list order (network.py:136-198; at
//     bookkeeping-free airports departures leave in list order), and the block simply becomes
//     part of the FIFO prefix of the same array.  A flight whose departure is served only at
//     or after its arrival tick (late path, network.py:186) is dropped from the list when its
//     tick comes; the ORIGIN hands it over when it serves the departure (an entry in the
//     destination's late box), and it joins at the end of that tick.  Nobody polls.
//   * Waits are EXACT by construction: the total wait of an entry is the left fold of the
//     service draws assigned while it was queued (types/airport.py:150-153).  Entries that
//     joined in the same tick share the fold's start, so one running fp32 accumulator serves
//     the whole head group; when an older accumulator cannot be continued (a new group
//     reaches the head) the fold is replayed from a history of draws that is only written
//     when a younger group is waiting (rare outside saturated queues).
//
// The same source is compiled by nvcc (sm_100a kernel) and by g++ (tests/hostemu).
#pragma once
#include "ba_sim.cuh"

#ifndef BA2_KC
#define BA2_KC 768                    // arrival-tick buckets kept in shared memory
#endif
#define BA2_NONE 0xFFFFFFFFu
#define BA2_TICK_NONE 0xFFFFu
#define BA2_STATE_WORDS 12            // packed words per airport

// per-flight payload, indexed by the list position of the flight's departure
struct alignas(16) Ba2Pay {
    float xw;        // departure service draw; once the departure is served: its total wait
    uint32_t q;      // tick at which the departure left its runway queue (BA2_NONE: not yet)
    float at;        // arrival time at the destination = queue start of the arrival entry
                     //   (diagnostics: overwritten by the arrival's pop tick when it is served)
    float xa;        // arrival service draw; once the arrival is served: its total wait
};

#if defined(__CUDA_ARCH__)
BA_DEV Ba2Pay ba2_ld_pay(const Ba2Pay *p)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    Ba2Pay r; r.xw = __uint_as_float(v.x); r.q = v.y; r.at = __uint_as_float(v.z); r.xa = __uint_as_float(v.w);
    return r;
}
BA_DEV void ba2_st_pay(Ba2Pay *p, const Ba2Pay &r)
{
    *reinterpret_cast<uint4 *>(p) = make_uint4(__float_as_uint(r.xw), r.q, __float_as_uint(r.at), __float_as_uint(r.xa));
}
struct Ba2U2 { uint32_t x, y; };
BA_DEV Ba2U2 ba2_ld2(const uint32_t *p) { const uint2 v = *reinterpret_cast<const uint2 *>(p); Ba2U2 r = {v.x, v.y}; return r; }
BA_DEV void ba2_st2(uint32_t *p, uint32_t a, uint32_t b) { *reinterpret_cast<uint2 *>(p) = make_uint2(a, b); }
#else
struct Ba2U2 { uint32_t x, y; };
BA_DEV Ba2Pay ba2_ld_pay(const Ba2Pay *p) { return *p; }
BA_DEV void ba2_st_pay(Ba2Pay *p, const Ba2Pay &r) { *p = r; }
BA_DEV Ba2U2 ba2_ld2(const uint32_t *p) { Ba2U2 v; v.x = p[0]; v.y = p[1]; return v; }
BA_DEV void ba2_st2(uint32_t *p, uint32_t a, uint32_t b) { p[0] = a; p[1] = b; }
#endif

#if defined(__CUDA_ARCH__)
#define BA2_STREAM_ST(p, v) __stcs((p), (v))      // written once, never re-read by this kernel
#else
#define BA2_STREAM_ST(p, v) (*(p) = (v))
#endif
#if defined(__CUDA_ARCH__) && !defined(BA2_NO_PREFETCH)
#define BA2_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#else
#define BA2_PREFETCH_L2(p) ((void)0)
#endif

struct Ba2Ctx {
    // static day tables
    int N, F;
    const BaFlight *flights;
    const int32_t *route, *pos_live, *pos_dep, *pos_arr, *rt_gid, *rt_off, *rt_flt;
    int Rd;
    const BaAirport *ap;
    const BaLiveDep *lq;
    const uint32_t *drun;
    const int32_t *drun_off;
    const int32_t *order;                     // [N] airport run in slot i (heaviest first)
    const float *tick_time;
    int first_tick, max_ticks;
    // particle inputs
    const float *lat_airport;                 // [2, N] of this particle
    const float *lat_route;
    const float *noise;
    uint64_t seed;
    uint32_t particle, dayidx;
    float sigma, v_t, v_u, inv_sigma, inv_sigma2, inv_sigma3, log_sigma, inv_vt, inv_vu;
    float c0_lp, c1_lp;
    int value_mode, want_grad;
    // shared memory of this warp
    uint32_t *un;                             // union: prologue {kcnt [BA2_KC], cursors [N], rate [N], mean turnaround [N]} /
                                              //   scan {state [BA2_STATE_WORDS][nslot]} /
                                              //   epilogue {rate [N], mean turnaround [N]}
    float *s_rate, *s_mt;                     // [N] per airport (day-local index), inside `un`
    uint32_t nslot;                           // stride of the state arrays (= N)
    // global scratch of this warp
    Ba2Pay *P;                                // [n_live + N] payload / tape / q by lpos
    uint32_t *Ak;                             // [n_arr + N] arrival keys per destination (+ sentinels)
    uint32_t *Lb;                             // [2][n_arr + N] late boxes per destination (even /
                                              //   odd ticks): lpos of the flights served at or
                                              //   after their arrival tick
    uint32_t lb_stride;
    uint32_t *S0, *S1;                        // prologue: key by flight [F] / {key, destination} [2 F]
                                              //   sorted by arrival tick
    float *hx;                                // [2 F] history of draws by assignment index
    uint32_t *glog;                           // [4 F] (tick, assignments before it) of ticks that
                                              //   ended with younger entries still waiting
    float *eps;                               // [F] travel-time draw of every flight, kept for the epilogue
    // epilogue: owner sums as 64-bit fixed-point accumulators (integer adds commute: any
    // order gives the same bits) -- per airport in shared memory, per route in scratch
    unsigned long long *accA;                 // [3][N]: departures, arrivals, turnaround (value mode)
    unsigned long long *accR;                 // [R] by global route id
    // outputs
    float *grow;
    int32_t *pop_tick;
    int route_base;
    // value mode: gradient w.r.t. the four per-flight site values (nullable)
    float *vrow;                              // [Fmax, 4] of this particle-day
    float *rD, *rA;                           // [n_live + N] residual / gradient of the departure
                                              //   and arrival entry of every live position
};

// words of per-warp global scratch
static inline uint64_t ba2_scratch_words(int F, int N, int R)
{
    const uint64_t f = (uint64_t)F, n = (uint64_t)N;
    return (f + 4) + 4 * (f + n + 4) + 3 * (f + n + 4) + 4 * f + 2 * f + 4 * f + 2 * (f + n + 4) +
           2 * (uint64_t)R + 64;
}

BA_DEV void ba2_carve_scratch(Ba2Ctx &c, uint32_t *g, int F, int N, int R)
{
    c.accR = (unsigned long long *)g; g += 2 * (((uint32_t)R + 1u) & ~1u);   // 16-byte aligned
    const uint32_t f = (uint32_t)F, fn = ((uint32_t)(F + N) + 3u) & ~3u;
    c.eps = (float *)g; g += (f + 3u) & ~3u;              // 16-byte aligned blocks first
    c.P = (Ba2Pay *)g; g += 4 * fn;
    c.S0 = g; c.S1 = g + ((f + 1u) & ~1u);
    g += 4 * f;
    c.Ak = g; g += fn;
    c.Lb = g; g += 2 * fn; c.lb_stride = fn;
    c.hx = (float *)g; g += 2 * f;
    c.glog = g; g += 4 * f;
    c.rD = (float *)g; c.rA = (float *)g + fn;
}

// shared-memory words per warp
#if defined(__CUDACC__)
__host__ __device__
#endif
static inline size_t ba2_smem_words(int N)
{
    const size_t nslot = (size_t)N;
    const size_t pro = BA2_KC + 3 * (size_t)N;
    const size_t un = BA2_STATE_WORDS * nslot > pro ? BA2_STATE_WORDS * nslot : pro;
    return (un + 3) & ~(size_t)3;
}

BA_DEV BaFlight ba2_flight(const Ba2Ctx &c, int f)
{
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(c.flights) + f);
    BaFlight fl;
    fl.sched_dep = __uint_as_float(v.x); fl.act_dep = __uint_as_float(v.y);
    fl.act_arr = __uint_as_float(v.z); fl.od = v.w;
    return fl;
#else
    return c.flights[f];
#endif
}

BA_DEV BaNoise4 ba2_noise(const Ba2Ctx &c, int f)
{
    if (c.noise) {
#if defined(__CUDA_ARCH__)
        // read once per phase: streaming (evict-first), the L2 is for the payload records
        const float4 v = __ldcs(reinterpret_cast<const float4 *>(c.noise) + f);
        BaNoise4 z = {v.x, v.y, v.z, v.w};
#else
        BaNoise4 z = {c.noise[4 * f], c.noise[4 * f + 1], c.noise[4 * f + 2], c.noise[4 * f + 3]};
#endif
        return z;
    }
    return ba_philox_noise(c.seed, c.particle, c.dayidx, (uint32_t)f);
}

// ---------------------------------------------------------------------------
// per-airport constants (model.py:139-146)
// ---------------------------------------------------------------------------
BA_DEV void ba2_init_airport(Ba2Ctx &c, int a, int Ng, bool epilogue)
{
    const int g = c.ap[a].global;
    (void)epilogue;
    c.s_mt[a] = c.lat_airport[0 * Ng + g];
    c.s_rate[a] = BA_DIV(1.0f, c.lat_airport[1 * Ng + g]);     // airport.py:146
}

// ---------------------------------------------------------------------------
// PROLOGUE sweep 1: one flight -> its draws, arrival time and arrival tick
// ---------------------------------------------------------------------------
// What a flight's prologue / epilogue step reads, gathered up front so that the loads of two
// flights are in flight together (the loops are latency-bound: one warp per particle-day).
struct Ba2FlightIn {
    BaFlight fl;
    BaNoise4 nz;
    int rt, lpos;
};

BA_DEV Ba2FlightIn ba2_prologue_load(const Ba2Ctx &c, int f)
{
    Ba2FlightIn in;
    in.fl = ba2_flight(c, f);
    in.nz = ba2_noise(c, f);
    in.rt = c.route[f];
    in.lpos = c.pos_live[f];
    return in;
}

// The flight's draws, arrival time and arrival tick; and the part of its log-probability that
// does not depend on any wait: the four latent per-flight sites (airport.py:144-147,
// network.py:158-163, airport.py:217-220).  The wait-dependent observed sites are the
// epilogue's.
BA_DEV void ba2_prologue_compute(Ba2Ctx &c, BaAcc &acc, int f, const Ba2FlightIn &in, float T)
{
    const BaFlight &fl = in.fl;
    const BaNoise4 &nz = in.nz;
    const uint32_t cobs = (fl.od >> 24) & 3u;
    const int o = (int)(fl.od & 0xFFFu), d = (int)((fl.od >> 12) & 0xFFFu);
#if defined(__CUDA_ARCH__)
    __stcs(c.eps + f, nz.eps_travel);
#else
    c.eps[f] = nz.eps_travel;
#endif
    uint32_t key = BA2_NONE;
    if (cobs != 1u) {
        const float rate_o = c.s_rate[o], rate_d = c.s_rate[d], mt = c.s_mt[d];
        // Exponential(1/m).rsample = e / rate  (airport.py:144-147)
        const float xd = c.value_mode ? nz.e_dep : BA_DIV(nz.e_dep, rate_o);
        const float xa = c.value_mode ? nz.e_arr : BA_DIV(nz.e_arr, rate_d);
        // Normal(T, T v_t).rsample = T + eps * (T v_t)  (network.py:158-163)
        const float sc = BA_MUL(T, c.v_t);
        const float tau = c.value_mode ? nz.eps_travel : BA_ADD(T, BA_MUL(nz.eps_travel, sc));
        const float tstd = BA_MUL(c.v_u, mt);
        const float u = c.value_mode ? nz.eps_turn : BA_ADD(mt, BA_MUL(nz.eps_turn, tstd));
        const float dz = (tau - T) / sc, du = (u - mt) / tstd;
        const float lp1 = (BA_LOG(rate_o) - rate_o * xd) + (BA_LOG(rate_d) - rate_d * xa);
        const float lp2 = (-0.5f * dz * dz - BA_LOG(sc) - BA_LOG_SQRT_2PI)
                        + (-0.5f * du * du - BA_LOG(tstd) - BA_LOG_SQRT_2PI);
        acc.lp += (double)lp1 + (double)lp2;
        float y_dep = fl.act_dep;
        if (cobs == 3u || isnan(y_dep) || isnan(fl.act_arr)) {
            acc.status |= BA_S_UNOBSERVED;      // this entry point is the fully observed path
            y_dep = fl.sched_dep;
        }
        const float at = BA_ADD(y_dep, tau);    // network.py:166-168 with the observed departure
        // arrival tick: the first tick whose clock reaches `at` (network.py:186)
        const int kmax = BA2_KC - 1 < c.max_ticks + 1 ? BA2_KC - 1 : c.max_ticks + 1;
        const float guess = ceilf(at / c.tick_time[1]);
        int kk = guess < 1.0f ? 1 : (guess > (float)kmax ? kmax : (int)guess);
        while (kk > 1 && c.tick_time[kk - 1] >= at) --kk;
        while (kk < kmax && c.tick_time[kk] < at) ++kk;
        if (c.tick_time[kk] < at) acc.status |= BA_S_OVERFLOW;      // beyond the buckets: exact lists
        if (kk < c.first_tick) kk = c.first_tick;
        Ba2Pay pay; pay.xw = xd; pay.at = at; pay.xa = xa; pay.q = BA2_NONE;
        ba2_st_pay(c.P + in.lpos, pay);
        key = ((uint32_t)kk << 16) | (uint32_t)in.lpos;
        BA_ATOMIC_INC(&c.un[kk]);
    }
    c.S0[f] = key;
}

// PROLOGUE sweep 1, one flight (host emulation; the kernel runs the two halves itself)
BA_DEV void ba2_prologue_flight(Ba2Ctx &c, BaAcc &acc, int f)
{
    const Ba2FlightIn in = ba2_prologue_load(c, f);
    ba2_prologue_compute(c, acc, f, in, c.lat_route[in.rt]);
}

// exclusive prefix over the tick buckets (the warp; host: lane = -1, serial)
BA_DEV void ba2_bucket_scan(Ba2Ctx &c, int lane)
{
    uint32_t *kcnt = c.un;
#if defined(__CUDA_ARCH__)
    const uint32_t per = (BA2_KC + 31u) / 32u;
    const uint32_t b = (uint32_t)lane * per, e = min(b + per, (uint32_t)BA2_KC);
    uint32_t sum = 0;
    for (uint32_t k = b; k < e; ++k) sum += kcnt[k];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t run = incl - sum;
    for (uint32_t k = b; k < e; ++k) { const uint32_t n = kcnt[k]; kcnt[k] = run; run += n; }
#else
    (void)lane;
    uint32_t run = 0;
    for (uint32_t k = 0; k < BA2_KC; ++k) { const uint32_t n = kcnt[k]; kcnt[k] = run; run += n; }
#endif
}

// PROLOGUE sweep 2: flight -> its place in the tick-sorted order
BA_DEV void ba2_scatter_flight(Ba2Ctx &c, int f)
{
    const uint32_t key = c.S0[f];
    if (key == BA2_NONE) return;
    const uint32_t d = (c.flights[f].od >> 12) & 0xFFFu;   // static table: an L2 hit
    const uint32_t pos = BA_ATOMIC_INC(&c.un[key >> 16]);
    ba2_st2(c.S1 + 2 * pos, key, d);
}

// start cursor of a destination's list (static: the lists have static sizes) + sentinel
BA_DEV void ba2_split_cursor(Ba2Ctx &c, int a)
{
    const BaAirport A = c.ap[a];
    c.un[BA2_KC + a] = A.as_off;
    c.Ak[A.as_off + (uint32_t)A.arr_cnt] = BA2_NONE;
}

// PROLOGUE sweep 3: stable split of the tick-sorted keys by destination.  Device: a warp
// walks its range of the order 32 keys at a time (match.any ranks equal destinations inside
// the chunk) with its own cursor per destination.
BA_DEV void ba2_split_range(Ba2Ctx &c, int lane, uint32_t *cur, uint32_t begin, uint32_t end)
{
#if defined(__CUDA_ARCH__)
    for (uint32_t base = begin; base < end; base += 32u) {
        const uint32_t i = base + (uint32_t)lane;
        const bool ok = i < end;
        const uint32_t act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const Ba2U2 kd = ba2_ld2(c.S1 + 2 * i);
            const uint32_t m = __match_any_sync(act, kd.y);
            const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
            const uint32_t pos = cur[kd.y] + rank;
            __syncwarp(act);
            if (rank + 1u == (uint32_t)__popc(m)) cur[kd.y] = pos + 1u;
            c.Ak[pos] = kd.x;
            __syncwarp(act);
        }
    }
#else
    (void)lane;
    for (uint32_t i = begin; i < end; ++i) c.Ak[cur[c.S1[2 * i + 1]]++] = c.S1[2 * i];
#endif
}

BA_DEV void ba2_split(Ba2Ctx &c, int lane, uint32_t n_arr)
{
    ba2_split_range(c, lane, c.un + BA2_KC, 0u, n_arr);
}

#if defined(__CUDA_ARCH__)
// With several warps on a particle-day the split runs on all of them: every warp takes a
// contiguous range of the tick-sorted order, the ranges' per-destination counts are turned
// into per-warp start cursors (a prefix over the warps, in order: still a stable split).
// `cnt` = [nw][N] words of shared memory behind the prologue's tables.
BA_DEV uint32_t ba2_split_chunk(uint32_t n_arr, int nw) { return ((n_arr + (uint32_t)nw - 1u) / (uint32_t)nw + 31u) & ~31u; }

BA_DEV void ba2_split_count(Ba2Ctx &c, int lane, int warp, int nw, uint32_t *cnt, uint32_t n_arr)
{
    const uint32_t chunk = ba2_split_chunk(n_arr, nw);
    const uint32_t begin = min(n_arr, (uint32_t)warp * chunk), end = min(n_arr, begin + chunk);
    uint32_t *mine = cnt + (size_t)warp * c.N;
    for (uint32_t i = begin + (uint32_t)lane; i < end; i += 32u) BA_ATOMIC_INC(&mine[c.S1[2 * i + 1]]);
}

BA_DEV void ba2_split_prefix(Ba2Ctx &c, int a, int nw, uint32_t *cnt)
{
    uint32_t run = c.un[BA2_KC + a];
    for (int w = 0; w < nw; ++w) { const uint32_t t = cnt[(size_t)w * c.N + a]; cnt[(size_t)w * c.N + a] = run; run += t; }
}

BA_DEV void ba2_split_walk(Ba2Ctx &c, int lane, int warp, int nw, uint32_t *cnt, uint32_t n_arr)
{
    const uint32_t chunk = ba2_split_chunk(n_arr, nw);
    const uint32_t begin = min(n_arr, (uint32_t)warp * chunk), end = min(n_arr, begin + chunk);
    ba2_split_range(c, lane, cnt + (size_t)warp * c.N, begin, end);
}
#endif

// ---------------------------------------------------------------------------
// SCAN: the warp's lanes walk the airports in rounds, all on one clock
// ---------------------------------------------------------------------------
// Per-airport state between ticks, packed (all list indices and ticks fit 16 bits):
//   [0] r | w << 16        arrival keys of this airport in Ak: [.., ha) served, [ha, w) queued
//   [1] ha | left << 16      (FIFO order), [w, r) deferred (arrival tick passed, departure not
//                            served), [r, ..) unexamined; left = entries still to serve
//   [2] hd | td << 16      departures in lq / P: [.., hd) served, [hd, td) queued
//   [3] rp | rt << 16      next run of the departure calendar and its tick
//   [4] key of Ak[r], fetched ahead
//   [5] A                  assigned service time of the head (airport.py:156)
//   [6] wacc               fold of the draws assigned since group g joined
//   [7] g | gnew << 16     joining tick of the accumulator's group / of the youngest group
//   [8] cnt | asg << 16    assignments made so far; 0: head not assigned, 1: arrival, 2: departure
//   [9] lw | lr << 16      log of (tick, cnt) pairs: written / consumed
//  [10] entries of this airport's two late boxes (even | odd << 16 ticks), incremented by
//       the origins
//  [11] hxo                this airport's slice of hx / glog
BA_DEV void ba2_state_init(const Ba2Ctx &c, uint32_t slot, int a)
{
    const BaAirport A = c.ap[a];
    uint32_t *s = c.un + slot;
    const uint32_t ns = c.nslot;
    const uint32_t rp = (uint32_t)c.drun_off[a];
    s[0] = A.as_off | (A.as_off << 16);
    s[1 * ns] = A.as_off | ((uint32_t)(A.dep_live + A.arr_cnt) << 16);
    s[2 * ns] = A.lq_off | (A.lq_off << 16);
    s[3 * ns] = rp | (c.drun[2 * rp] << 16);
    s[4 * ns] = c.Ak[A.as_off];
    s[5 * ns] = 0; s[6 * ns] = 0;
    s[7 * ns] = BA2_TICK_NONE | (BA2_TICK_NONE << 16);
    s[8 * ns] = 0; s[9 * ns] = 0;
    s[10 * ns] = 0;
    s[11 * ns] = A.lq_off + A.as_off - 2u * (uint32_t)a;
}

// in-transit list order of an arrival: (tick at which its departure was served, origin,
// position in the origin's queue) = (q, list position of the departure)
BA_DEV uint32_t ba2_key(uint32_t q, uint32_t lpos) { return (q << 16) | lpos; }

// Block of this tick's joiners while an out-of-line path works on it.
struct Ba2Blk { uint32_t w, pk; };

// Inserts the arrival `key` (its departure was served at tick q <= km1; `okey` = its place
// in in-transit list order) into the block [wb, w) of this tick's joiners, kept sorted; the
// slot Ak[w] is free (w <= r: the flights that went to the late box left holes).
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
void ba2_insert(uint32_t *Ak, const Ba2Pay *P, Ba2Blk &B, uint32_t key, uint32_t okey, uint32_t wb)
{
    uint32_t j = B.w;
    while (j > wb) {
        const uint32_t e = Ak[j - 1];
        const uint32_t el = e & 0xFFFFu;
        if (ba2_key(P[el].q, el) <= okey) break;
        Ak[j] = e;
        --j;
    }
    Ak[j] = key;
    if (okey > B.pk || B.w == wb) B.pk = okey;
    B.w += 1;
}

// A new group (joining tick p) reaches the head: its waits are the fold of the draws
// assigned since it joined (types/airport.py:150-153) -- replay them from the history.
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
float ba2_replay(const uint32_t *glog, const float *hx, uint32_t lw, uint32_t *lr_io, uint32_t cnt,
                 uint32_t p, uint32_t k, uint32_t cnt0, uint32_t *status)
{
    uint32_t lo = cnt0;
    if (p != k) {
        uint32_t lr = *lr_io;
        while (lr < lw && glog[2u * lr] != p) lr += 1;
        if (lr == lw) { *status |= BA_S_INTERNAL; return 0.0f; }
        lo = glog[2u * lr + 1u];
        *lr_io = lr + 1;
    }
    float w = 0.0f;
    for (uint32_t j = lo; j < cnt; ++j) w = BA_ADD(w, hx[j]);
    return w;
}

// One tick of the airport in `slot`: joins (model.py:181-183,186-196), then
// update_runway_queue (types/airport.py:86-128).  Returns true when the airport has just
// served its last entry.
BA_DEV bool ba2_airport_tick(const Ba2Ctx &c, uint32_t slot, uint32_t k, float t, uint32_t &status)
{
    uint32_t *s = c.un + slot;
    const uint32_t ns = c.nslot;
    const uint32_t s1 = s[1 * ns];
    uint32_t left = s1 >> 16;
    if (left == 0u) return false;
    const uint32_t s0 = s[0], s2 = s[2 * ns], s3 = s[3 * ns], s8 = s[8 * ns];
    uint32_t nkey = s[4 * ns];
    uint32_t r = s0 & 0xFFFFu, w = s0 >> 16, ha = s1 & 0xFFFFu;
    uint32_t hd = s2 & 0xFFFFu, td = s2 >> 16, rp = s3 & 0xFFFFu, rt = s3 >> 16;
    uint32_t cnt = s8 & 0xFFFFu, asg = s8 >> 16;
    const uint32_t km1 = k - 1u;
    // late entries handed over during tick k - 1 (the box of that tick's parity)
    const uint32_t lsh = (km1 & 1u) * 16u;
    const uint32_t nlate = (*(volatile uint32_t *)(s + 10 * ns) >> lsh) & 0xFFFFu;
    // nothing queued, nothing joins: most airport-ticks of the light airports
    if ((nkey >> 16) > km1 && nlate == 0u && rt != k && ha == w && hd == td && asg == 0u) return false;
    // What this tick will certainly touch first is requested up front, as independent loads
    // (one memory round trip instead of a chain): the q of the first due arrival and the key
    // behind it, the record of the departure at the head of its list.
    uint32_t q_first = BA2_NONE, nkey2 = BA2_NONE;
    if ((nkey >> 16) <= km1) { q_first = c.P[nkey & 0xFFFFu].q; nkey2 = c.Ak[r + 1u]; }
    uint32_t dk = BA2_NONE;                   // lq[hd] and its draw, valid within this tick; bit 31:
    float dqs = 0.0f, dx = 0.0f;              //   its arrival time has passed already
    if (hd < td || rt == k) {
        const Ba2Pay pay = ba2_ld_pay(c.P + hd);
        const Ba2U2 lq2 = ba2_ld2((const uint32_t *)(c.lq + hd));
        dk = lq2.x | (pay.at <= t ? 0x80000000u : 0u); dqs = __uint_as_float_portable(lq2.y); dx = pay.xw;
    }
    float A = __uint_as_float_portable(s[5 * ns]), wacc = __uint_as_float_portable(s[6 * ns]);
    const uint32_t s7 = s[7 * ns];
    uint32_t g = s7 & 0xFFFFu, gnew = s7 >> 16;

    bool joined = false;
    const uint32_t wb = w;
    uint32_t pk = 0;
    // ---- arrivals that joined at the end of tick k - 1 ------------------------------
    // (1) from the list: arrival tick k - 1, departure served before that tick.  A flight
    //     whose departure had not been served by then comes through the late box instead.
    while ((nkey >> 16) <= km1) {
        const uint32_t lpos = nkey & 0xFFFFu;
        const uint32_t q = q_first != BA2_NONE ? q_first : c.P[lpos].q;
        q_first = BA2_NONE;
        if (q < (nkey >> 16)) {
            const uint32_t okey = ba2_key(q, lpos);
            if (w == r && (w == wb || okey >= pk)) {
                pk = okey; w += 1;                           // already in place
            } else {
                Ba2Blk B; B.w = w; B.pk = pk;
                ba2_insert(c.Ak, c.P, B, nkey, okey, wb);
                w = B.w; pk = B.pk;
            }
            joined = true;
        }
        r += 1;
        nkey = nkey2 != BA2_NONE ? nkey2 : c.Ak[r];
        nkey2 = BA2_NONE;
    }
    // (2) from the late box: departures served in tick k - 1 at or after the arrival tick
    //     (network.py:186); they sort behind (1): their q is larger
    if (nlate != 0u) {
        const uint32_t *lb = c.Lb + (km1 & 1u) * c.lb_stride + c.ap[c.order[slot]].as_off;
        Ba2Blk B; B.w = w; B.pk = pk;
        for (uint32_t i = 0; i < nlate; ++i) {
            const uint32_t lpos = lb[i];
            ba2_insert(c.Ak, c.P, B, (km1 << 16) | lpos, ba2_key(km1, lpos), wb);
        }
        w = B.w; pk = B.pk;
        BA_ATOMIC_AND(s + 10 * ns, ~(0xFFFFu << lsh));       // the box is empty again
        joined = true;
    }
    // ---- departures that become ready at tick k (static calendar run) -----------------
    if (rt == k) {
        td += c.drun[2 * rp + 1] & 0xFFFFu;
        rp += 1;
        rt = c.drun[2 * rp] & 0xFFFFu;
        joined = true;
    }
    if (joined) gnew = k;
    // ---- service loop -----------------------------------------------------------------
    const uint32_t cnt0 = cnt;
    uint32_t hkey = BA2_NONE;                 // Ak[ha] and its payload, valid within this tick
    float hat = 0.0f, hxa = 0.0f;
    uint32_t lwr = BA2_NONE;                  // log indices, loaded on demand
    for (;;) {
        if (asg == 0u) {
            const bool ha_ok = ha < w, hd_ok = hd < td;
            if (!ha_ok && !hd_ok) break;
            if (ha_ok && hkey == BA2_NONE) {
                hkey = c.Ak[ha];
                const Ba2Pay pay = ba2_ld_pay(c.P + (hkey & 0xFFFFu));
                hat = pay.at; hxa = pay.xa;
            }
            if (hd_ok && dk == BA2_NONE) {
                const Ba2Pay pay = ba2_ld_pay(c.P + hd);
                const Ba2U2 lq2 = ba2_ld2((const uint32_t *)(c.lq + hd));
                dk = lq2.x | (pay.at <= t ? 0x80000000u : 0u); dqs = __uint_as_float_portable(lq2.y); dx = pay.xw;
            }
            const uint32_t pa = ha_ok ? (hkey >> 16) + 1u : BA2_NONE;
            const uint32_t pd = hd_ok ? (dk & 0xFFFFu) : BA2_NONE;
            const bool arr = pa <= pd;                       // arrivals first on ties
            const uint32_t p = arr ? pa : pd;
            const float x = arr ? hxa : dx;
            const float qs = arr ? hat : dqs;
            if (p != g) {
                if (p == k && cnt == cnt0) wacc = 0.0f;
                else {
                    const uint32_t hxo = s[11 * ns];
                    if (lwr == BA2_NONE) lwr = s[9 * ns];
                    uint32_t lr = lwr >> 16;
                    wacc = ba2_replay(c.glog + 2u * hxo, c.hx + hxo, lwr & 0xFFFFu, &lr, cnt, p, k, cnt0, &status);
                    lwr = (lwr & 0xFFFFu) | (lr << 16);
                }
                g = p;
            }
            if (gnew != p) c.hx[s[11 * ns] + cnt] = x; // a younger group is waiting
            cnt += 1;
            wacc = BA_ADD(wacc, x);                          // airport.py:150-153
            A = BA_ADD(qs, wacc);                            // airport.py:156
            asg = arr ? 1u : 2u;
        }
        if (!(A <= t)) break;                                // airport.py:111
        if (asg == 1u) {
            if (hkey == BA2_NONE) hkey = c.Ak[ha];
            Ba2Pay *pp = c.P + (hkey & 0xFFFFu);
            pp->xa = wacc;                                   // tape: total wait of the arrival
            if (c.pop_tick || c.vrow) pp->at = __uint_as_float_portable(k);
            ha += 1; hkey = BA2_NONE;
        } else {
            // tape: total wait of the departure, and the message to the destination
            // (an entry assigned in an earlier tick: look its arrival time up again)
            const bool late = dk == BA2_NONE ? c.P[hd].at <= t : (dk >> 31) != 0u;
            ba2_st2((uint32_t *)(c.P + hd), __float_as_uint_portable(wacc), k);
            if (late) {
                // served at or after its arrival time: joins the destination's queue at the
                // end of this very tick (update_in_transit_flights, network.py:186)
                const BaLiveDep ld = c.lq[hd];
                const uint32_t sh = (k & 1u) * 16u;
                const uint32_t n = (BA_ATOMIC_ADD(&c.un[10u * ns + ld.dslot], 1u << sh) >> sh) & 0xFFFFu;
                c.Lb[(k & 1u) * c.lb_stride + ld.das_off + n] = hd;
            }
            hd += 1; dk = BA2_NONE;
        }
        left -= 1; asg = 0u;
    }
    // a tick that ends with entries still waiting behind an assigned head: the entries that
    // joined in it will need the number of assignments made before it (ba2_replay)
    if (joined && asg != 0u) {
        if (lwr == BA2_NONE) lwr = s[9 * ns];
        uint32_t *lg = c.glog + 2u * (s[11 * ns] + (lwr & 0xFFFFu));
        lg[0] = k; lg[1] = cnt0;
        lwr += 1u;
    }
    // what the next tick touches first: on its way to L2 while the other airports run
    if ((nkey >> 16) <= k) BA2_PREFETCH_L2(c.P + (nkey & 0xFFFFu));
    if (hd < td || rt == k + 1u) BA2_PREFETCH_L2(c.P + hd);
    s[0] = r | (w << 16);
    s[1 * ns] = ha | (left << 16);
    s[2 * ns] = hd | (td << 16);
    s[3 * ns] = rp | (rt << 16);
    s[4 * ns] = nkey;
    s[5 * ns] = __float_as_uint_portable(A); s[6 * ns] = __float_as_uint_portable(wacc);
    s[7 * ns] = g | (gnew << 16);
    s[8 * ns] = cnt | (asg << 16);
    if (lwr != BA2_NONE) s[9 * ns] = lwr;
    return left == 0u;
}

// ---------------------------------------------------------------------------
// EPILOGUE (flight-parallel; owner sums through fixed-point accumulators); SURVEY.md
// App. A.3 / A.5
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// VALUE MODE: gradient of the log-probability w.r.t. the four per-flight site values (the
// sites a guide over them proposes, wn_network.py:285-302).  Direct terms per flight here;
// the wait-mediated terms per airport below.
// ---------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
void ba2_value_grad_flight(Ba2Ctx &c, int f, int lpos, int o, int d, float T, float tau, float r_d, float r_a)
{
    const BaNoise4 nz = ba2_noise(c, f);                  // the values themselves
    const float mt = c.s_mt[d], tstd = BA_MUL(c.v_u, mt), sc = BA_MUL(T, c.v_t);
    const float dz = (tau - T) / sc, du = (nz.eps_turn - mt) / tstd;
    // Exponential(x; rate): -rate; Normal travel time: -dz / sc, and it moves the arrival's
    // queue start (airport.py:197-204): + r_a; Normal turnaround time: -du / tstd
    c.vrow[4 * f + 0] = -c.s_rate[o];
    c.vrow[4 * f + 1] = -dz / sc + r_a;
    c.vrow[4 * f + 2] = -c.s_rate[d];
    c.vrow[4 * f + 3] = -du / tstd;
    c.rD[lpos] = r_d;
    c.rA[lpos] = r_a;
}

// The total wait of a queue entry is the sum of the service draws assigned while it was
// queued (types/airport.py:150-153): the draws of the entries served from the tick it joined
// up to and including its own.  So d logp / d (draw of the m-th served entry) is the sum of
// the residuals r_e of the entries e whose waiting interval [lo_e, e] contains m.  The service
// order is the FIFO merge of the airport's arrivals and departures by joining tick (arrivals
// first on ties), an entry is assigned its draw when its predecessor is served or when it
// joins an empty queue, and lo_e counts the assignments made before the tick e joined -- all
// recoverable from what the scan left: joining and serving ticks of every entry.
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
void ba2_value_grad_airport(Ba2Ctx &c, int a)
{
    const BaAirport A = c.ap[a];
    const uint32_t n_dep = (uint32_t)A.dep_live, n_arr = (uint32_t)A.arr_cnt, n = n_dep + n_arr;
    if (n == 0u) return;
    const uint32_t off = A.lq_off + A.as_off - 2u * (uint32_t)a;      // this airport's slices
    float *D = c.hx + off;                     // [n] difference array, then the gradient
    uint32_t *ident = c.glog + 2u * off;       // [n] {live position | departure << 31, assignment tick}
    for (uint32_t m = 0; m < n; ++m) D[m] = 0.0f;
    uint32_t ia = 0, id = 0, prev_pop = 0, lo = 0;
    for (uint32_t m = 0; m < n; ++m) {
        const bool a_ok = ia < n_arr, d_ok = id < n_dep;
        uint32_t ja = BA2_NONE, jd = BA2_NONE, akey = 0;
        if (a_ok) { akey = c.Ak[A.as_off + ia]; ja = (akey >> 16) + 1u; }
        if (d_ok) jd = c.lq[A.lq_off + id].tick;
        const bool arr = ja <= jd;                         // arrivals first on ties
        const uint32_t join = arr ? ja : jd;
        const uint32_t lpos = arr ? (akey & 0xFFFFu) : A.lq_off + id;
        const uint32_t pop = arr ? __float_as_uint_portable(c.P[lpos].at) : c.P[lpos].q;
        const float r = arr ? c.rA[lpos] : c.rD[lpos];
        const uint32_t asg = m == 0u ? join : (prev_pop > join ? prev_pop : join);
        ident[2u * m] = lpos | (arr ? 0u : 0x80000000u);
        ident[2u * m + 1u] = asg;
        while (lo < m && ident[2u * lo + 1u] < join) ++lo;   // assignments before the joining tick
        D[lo] += r;
        if (m + 1u < n) D[m + 1u] -= r;
        prev_pop = pop;
        if (arr) ++ia; else ++id;
    }
    float run = 0.0f;
    for (uint32_t m = 0; m < n; ++m) {
        run += D[m];
        const uint32_t w = ident[2u * m];
        if (w & 0x80000000u) c.rD[w & 0x7FFFFFFFu] = run; else c.rA[w] = run;
    }
}

// adds the wait-mediated terms to the two service-time values of a flight
BA_DEV void ba2_value_grad_finish(Ba2Ctx &c, int f)
{
    const int lpos = c.pos_live[f];
    if (lpos < 0) return;
    c.vrow[4 * f + 0] += c.rD[lpos];
    c.vrow[4 * f + 2] += c.rA[lpos];
}

// fixed point of the owner sums: 2^-28 resolution, +-3.4e10 range
#define BA2_FX_SCALE 268435456.0f
BA_DEV long long ba2_fx(float x) { return (long long)llrintf(x * BA2_FX_SCALE); }
BA_DEV float ba2_unfx(unsigned long long v) { return (float)((double)(long long)v * (1.0 / (double)BA2_FX_SCALE)); }

struct Ba2EpiIn {
    BaFlight fl;
    int lpos, rt;
    float eps;
};

BA_DEV Ba2EpiIn ba2_epilogue_load(const Ba2Ctx &c, int f)
{
    Ba2EpiIn in;
    in.fl = ba2_flight(c, f);
    in.lpos = c.pos_live[f]; in.rt = c.route[f];
#if defined(__CUDA_ARCH__)
    in.eps = __ldcs(c.eps + f);
#else
    in.eps = c.eps[f];
#endif
    return in;
}

// second round of loads: the flight's payload record (waits, arrival time) and travel time
BA_DEV Ba2Pay ba2_epilogue_pay(const Ba2Ctx &c, const Ba2EpiIn &in)
{
    if (((in.fl.od >> 24) & 3u) == 1u) { Ba2Pay z; z.xw = 0.0f; z.q = BA2_NONE; z.at = 0.0f; z.xa = 0.0f; return z; }
    return ba2_ld_pay(c.P + in.lpos);
}

// The wait-dependent sites of a flight (*_simulated_departure_time, *_simulated_arrival_time,
// airport.py:175-182,197-204), its cancellation site, and its gradient contributions.
BA_DEV void ba2_epilogue_compute(Ba2Ctx &c, BaAcc &acc, int f, const Ba2EpiIn &in, const Ba2Pay &pay, float T)
{
    const BaFlight &fl = in.fl;
    const uint32_t cobs = (fl.od >> 24) & 3u;
    const int o = (int)(fl.od & 0xFFFu), d = (int)((fl.od >> 12) & 0xFFFu);
    // *_cancelled (network.py:100-109): the cancellation probability is exactly 0 at every
    // airport of such a plan (ba_plan.h: BA_SIMPLE_RATIO)
    acc.lp += (double)(cobs == 1u ? c.c1_lp : c.c0_lp);
    if (cobs == 1u) {                 // observed as cancelled: never queued anywhere
        if (c.pop_tick) { c.pop_tick[2 * f] = -1; c.pop_tick[2 * f + 1] = -1; }
        if (c.vrow) { c.vrow[4 * f] = 0.0f; c.vrow[4 * f + 1] = 0.0f; c.vrow[4 * f + 2] = 0.0f; c.vrow[4 * f + 3] = 0.0f; }
        return;
    }
    if (c.pop_tick) {
        c.pop_tick[2 * f] = pay.q == BA2_NONE ? -1 : (int32_t)pay.q;
        c.pop_tick[2 * f + 1] = (int32_t)__float_as_uint_portable(pay.at);
    }
    const bool grad = c.want_grad != 0;
    const float half_inv_var = 0.5f * c.inv_sigma2;
    const float lp_const = -c.log_sigma - BA_LOG_SQRT_2PI;
    float y_dep = fl.act_dep, y_arr = fl.act_arr;
    if (cobs == 3u || isnan(y_dep) || isnan(y_arr)) { y_dep = fl.sched_dep; y_arr = fl.sched_dep; }

    // departure side (airport.py:175-182): Normal(queue_start + total_wait, sigma)
    const float wdep = pay.xw;
    const float qstart_d = fmaxf(0.0f, fl.sched_dep);
    const float mu_d = BA_ADD(qstart_d, wdep);
    const float dy_d = y_dep - mu_d;
    // arrival side (airport.py:197-204): queue_start = departure + travel time
    const float sc = BA_MUL(T, c.v_t);
    const float tau = c.value_mode ? in.eps : BA_ADD(T, BA_MUL(in.eps, sc));
    const float qstart_a = BA_ADD(y_dep, tau);
    const float warr = pay.xa;
    const float mu_a = BA_ADD(qstart_a, warr);
    const float dy_a = y_arr - mu_a;
    const float lp = (lp_const - dy_d * dy_d * half_inv_var) + (lp_const - dy_a * dy_a * half_inv_var);
    acc.lp += (double)lp;
    if (c.vrow) ba2_value_grad_flight(c, f, in.lpos, o, d, T, tau, dy_d * c.inv_sigma2, dy_a * c.inv_sigma2);

    if (grad) {
        const float r_d = dy_d * c.inv_sigma2, r_a = dy_a * c.inv_sigma2;
        acc.g_sigma += (dy_d * dy_d + dy_a * dy_a) * c.inv_sigma3 - 2.0f * c.inv_sigma;
        // owner sums: per origin / destination airport in shared memory, per route in scratch
        const int N = c.N;
        float a_dep, a_arr, a_rt;
        if (!c.value_mode) {
            a_dep = r_d * wdep;
            a_arr = r_a * warr;
            // arrival queue_start = act_dep + tau, tau = T (1 + v_t eps)
            a_rt = r_a * (1.0f + c.v_t * in.eps) - 1.0f / T;
            acc.g_vt += r_a * in.eps * T - c.inv_vt;
            acc.g_vu += -c.inv_vu;
        } else {
            // value mode (tests, callers that pin the per-flight site values): the values
            // are read again; the draws are not functions of the latents here
            const BaNoise4 nz = ba2_noise(c, f);
            const float mt = c.s_mt[d], tstd = BA_MUL(c.v_u, mt);
            const float dz = (tau - T) / sc;
            const float du = (nz.eps_turn - mt) / tstd;
            a_dep = nz.e_dep;
            a_arr = nz.e_arr;
            a_rt = dz / sc + (dz * dz - 1.0f) / T;
            acc.g_vt += (dz * dz - 1.0f) * c.inv_vt;
            BA_ATOMIC_ADD64(c.accA + 2 * N + d, ba2_fx(du / tstd + (du * du - 1.0f) / mt));
            acc.g_vu += (du * du - 1.0f) * c.inv_vu;
        }
        BA_ATOMIC_ADD64(c.accA + o, ba2_fx(a_dep));
        BA_ATOMIC_ADD64(c.accA + N + d, ba2_fx(a_arr));
        BA_ATOMIC_ADD64(c.accR + in.rt, ba2_fx(a_rt));
    }
}

// EPILOGUE of one flight (host emulation; the kernel runs the stages itself)
BA_DEV void ba2_epilogue_flight(Ba2Ctx &c, BaAcc &acc, int f)
{
    const Ba2EpiIn in = ba2_epilogue_load(c, f);
    const Ba2Pay pay = ba2_epilogue_pay(c, in);
    ba2_epilogue_compute(c, acc, f, in, pay, c.lat_route[in.rt]);
}

BA_DEV void ba2_epilogue_airport(Ba2Ctx &c, int a, int Ng)
{
    const BaAirport A = c.ap[a];
    const float mt = c.s_mt[a], m = c.lat_airport[1 * Ng + A.global];
    const float s_dep = ba2_unfx(c.accA[a]), s_arr = ba2_unfx(c.accA[c.N + a]);
    const float n_ent = (float)(A.dep_live + A.arr_cnt), n_land = (float)A.arr_cnt;
    float g_serv, g_turn;
    if (!c.value_mode) {
        // x = e m  =>  d/dm = sum r W / m - (#entries)/m ;  u = m~ (1 + v_u eps)
        g_serv = (s_dep + s_arr) / m - n_ent / m;
        g_turn = -n_land / mt;
    } else {
        // log Exponential(x; 1/m) = -log m - x/m
        g_serv = -n_ent / m + (s_dep + s_arr) / (m * m);
        g_turn = ba2_unfx(c.accA[2 * c.N + a]);
    }
    const int g = A.global;
    BA2_STREAM_ST(c.grow + 0 * Ng + g, g_turn);
    BA2_STREAM_ST(c.grow + 1 * Ng + g, g_serv);
}

// one GLOBAL route (routes without a flight on this day keep a zero accumulator)
BA_DEV void ba2_epilogue_route(Ba2Ctx &c, int gid)
{
    BA2_STREAM_ST(c.grow + c.route_base + gid, ba2_unfx(c.accR[gid]));
}
