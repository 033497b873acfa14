// Ring-less scan ("scan2") of the hot path for plans without aircraft bookkeeping and with
// every flight observed -- the training configurations (BASELINE configs[1], [2], [4]).
//
// Same three stages as ba_sim.cuh (prologue / scan / epilogue) and the same reference
// semantics (bayes_air/model.py:158-200, network.py:43-198, types/airport.py:86-221), but
// the scan keeps no queue rings, no due buffers and takes no CTA barrier per tick:
//
//   * The runway queue of an airport is VIRTUAL.  Its departures join in a static order
//     (pending order of its live departures; BaDayView::lq), its arrivals in in-transit list
//     order.  FIFO order is the merge of the two sequences by joining tick (arrivals of the
//     end of tick k-1 before the departures of tick k, model.py:181-196), so the queue is two
//     read pointers and nothing is ever copied into a ring.
//   * Arrivals: the prologue leaves, per destination, the list of its arrivals SORTED BY
//     ARRIVAL TICK (counting sort by tick, then a stable split by destination).  At tick k the
//     airport's lane takes the records whose tick has come and whose departure has been served
//     (q_dep <= k-1, read from a shared u16 table the origin lanes fill), orders that small
//     block by (q_dep, static departure position) = the reference's in-transit list order
//     (network.py:136-198), and the block simply becomes part of the FIFO prefix of the same
//     array.  Flights whose departure is still queued stay behind in a "deferred" zone
//     between the FIFO prefix and the unexamined suffix and are polled each tick (late path).
//   * Waits are EXACT by construction: the total wait of an entry is the left fold of the
//     service draws assigned while it was queued (types/airport.py:150-153).  Entries that
//     joined in the same tick share the fold's start, so one running fp32 accumulator serves
//     the whole head group; when an older accumulator cannot be continued (a new group
//     reaches the head) the fold is replayed from a history of draws that is only written
//     when a younger group is waiting (rare outside saturated queues).
//   * Warps run ASYNCHRONOUSLY (the lanes of a warp share a clock): a lane needs another
//     airport's state only when a departure it is waiting for has not been served yet; only
//     then does it look at how far the ORIGIN's warp has come (per-warp progress words in
//     shared memory) and, if that warp is behind, suspends on the origin warp's mbarrier
//     (one phase per tick) -- nobody spins.  The normal case (the flight left its origin
//     many ticks ago) touches no synchronisation at all.
//
// The same source is compiled by nvcc (sm_100a kernel) and by g++ (tests/hostemu: one
// std::thread per warp, same protocol).
#pragma once
#include "ba_sim.cuh"

#if !defined(__CUDACC__)
#include <atomic>
#include <thread>
#endif

#ifndef BA2_KC
#define BA2_KC 768                    // arrival-tick buckets kept in shared memory
#endif
#define BA2_MAX_RANGES 8              // ranges (= warps) of the stable split by destination
#define BA2_Q_INVALID 0xFFFFu         // departure not served yet
#define BA2_NONE 0xFFFFFFFFu
#define BA2_PROG_DONE 0x7FFFFFFF

// arrival record: list position of the flight's departure (16 bits: the static part of the
// ordering key; the high half carries the destination until the split is done, afterwards the
// warp that runs the origin airport), arrival time at the destination (= queue start),
// arrival service draw, arrival tick (late joiners: the tick at whose end they joined; until
// the split is done the high half carries the origin's warp)
struct alignas(16) Ba2Rec { uint32_t lpos; float at; float xa; uint32_t ka; };

#if defined(__CUDA_ARCH__)
BA_DEV Ba2Rec ba2_ld_rec(const Ba2Rec *p)
{
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    Ba2Rec r; r.lpos = v.x; r.at = __uint_as_float(v.y); r.xa = __uint_as_float(v.z); r.ka = v.w;
    return r;
}
BA_DEV void ba2_st_rec(Ba2Rec *p, const Ba2Rec &r)
{
    *reinterpret_cast<uint4 *>(p) = make_uint4(r.lpos, __float_as_uint(r.at), __float_as_uint(r.xa), r.ka);
}
#else
BA_DEV Ba2Rec ba2_ld_rec(const Ba2Rec *p) { return *p; }
BA_DEV void ba2_st_rec(Ba2Rec *p, const Ba2Rec &r) { *p = r; }
#endif
#if defined(__CUDA_ARCH__) && !defined(BA2_NO_PREFETCH)
#define BA2_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
#else
#define BA2_PREFETCH(p) ((void)0)
#endif
#if defined(__CUDACC__)
#define BA2_FENCE() __threadfence_block()
#define BA2_PAUSE(ns) __nanosleep(ns)
#else
#define BA2_FENCE() std::atomic_thread_fence(std::memory_order_seq_cst)
#define BA2_PAUSE(ns) std::this_thread::yield()
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
    const float *tick_time;
    int first_tick, max_ticks;
    // particle inputs
    const float *lat_route;
    const float *noise;
    uint64_t seed;
    uint32_t particle, dayidx;
    float sigma, v_t, v_u, inv_sigma, inv_sigma2, inv_sigma3, log_sigma, inv_vt, inv_vu;
    float c0_lp, c1_lp;
    int value_mode, want_grad;
    // shared memory
    float *s_m, *s_rate, *s_mt, *s_tstd;      // [N] per airport (day-local index)
    uint32_t *kcnt;                           // [BA2_KC] arrival-tick buckets
    uint32_t *hist;                           // [n_ranges * N] split cursors
    volatile uint16_t *qsh;                   // [n_live + N] tick at which each live departure
                                              //   left its runway queue (BA2_Q_INVALID: not yet)
    volatile int *prog;                       // [nwarp] last tick each warp has completed
    uint64_t *bars;                           // [nwarp] mbarrier of each warp (one phase per tick)
    int nwarp, n_ranges;
    uint32_t range_len;
    // global scratch of this CTA
    Ba2Rec *As;                               // [n_arr + N] arrival lists (one sentinel each)
    Ba2Rec *tmpf;                             // [F] prologue: records by flight
    uint32_t *idx1;                           // [F] prologue: flights sorted by arrival tick
    float *xdep;                              // [n_live + N] departure service draw by lpos
    float *wt;                                // [2 (n_live + N)] tape: total wait of the departure
                                              //   (even) / arrival (odd) entry of the flight at lpos
    uint32_t *qarr;                           // [n_live + N] tick at which the arrival entry left
    float *hx;                                // [2 F] history of draws by assignment index
    uint32_t *glog;                           // [4 F] (tick, assignments before it) of ticks that
                                              //   ended with younger entries still waiting
    float *nzs;                               // [4 F] generated noise kept for the epilogue
    float *xdk, *xak, *atk, *ct;              // epilogue contributions (alias tmpf)
    // outputs
    float *grow;
    int32_t *pop_tick;
    int route_base;
};

// words of per-CTA global scratch
static inline uint64_t ba2_scratch_words(int F, int N)
{
    const uint64_t f = (uint64_t)F, n = (uint64_t)N;
    return 4 * f + 4 * (f + n) + 4 * f + f + 4 * (f + n) + 2 * f + 4 * f + 64;
}

BA_DEV void ba2_carve_scratch(Ba2Ctx &c, uint32_t *g, int F, int N)
{
    const uint32_t f = (uint32_t)F, fn = (uint32_t)(F + N);
    c.nzs = (float *)g; g += 4 * f;                       // 16-byte aligned blocks first
    c.As = (Ba2Rec *)g; g += 4 * fn;
    c.tmpf = (Ba2Rec *)g;
    c.xdk = (float *)g; c.xak = (float *)g + f; c.atk = (float *)g + 2 * f; c.ct = (float *)g + 3 * f;
    g += 4 * f;
    c.idx1 = g; g += f;
    c.xdep = (float *)g; g += fn;
    c.wt = (float *)g; g += 2 * fn;
    c.qarr = g; g += fn;
    c.hx = (float *)g; g += 2 * f;
    c.glog = g;
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
        const float4 v = __ldg(reinterpret_cast<const float4 *>(c.noise) + f);
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
BA_DEV void ba2_init_airport(Ba2Ctx &c, int a, const float *lat_airport /* [2, N] */, int Ng)
{
    const int g = c.ap[a].global;
    const float mt = lat_airport[0 * Ng + g];
    const float m = lat_airport[1 * Ng + g];
    c.s_mt[a] = mt; c.s_m[a] = m;
    c.s_rate[a] = BA_DIV(1.0f, m);               // airport.py:146
    c.s_tstd[a] = BA_MUL(c.v_u, mt);             // model.py:143-145
}

// ---------------------------------------------------------------------------
// PROLOGUE sweep 1: one flight -> its draws, arrival time and arrival tick
// ---------------------------------------------------------------------------
BA_DEV void ba2_prologue_flight(Ba2Ctx &c, BaAcc &acc, int f)
{
    const BaFlight fl = ba2_flight(c, f);
    const uint32_t cobs = (fl.od >> 24) & 3u;
    const int o = (int)(fl.od & 0xFFFu), d = (int)((fl.od >> 12) & 0xFFFu);
    const BaNoise4 nz = ba2_noise(c, f);
    if (!c.noise) {
#if defined(__CUDA_ARCH__)
        reinterpret_cast<float4 *>(c.nzs)[f] = make_float4(nz.e_dep, nz.eps_travel, nz.e_arr, nz.eps_turn);
#else
        c.nzs[4 * f] = nz.e_dep; c.nzs[4 * f + 1] = nz.eps_travel; c.nzs[4 * f + 2] = nz.e_arr; c.nzs[4 * f + 3] = nz.eps_turn;
#endif
    }
    Ba2Rec rec;
    rec.lpos = 0; rec.at = 0.0f; rec.xa = 0.0f; rec.ka = BA2_NONE;
    if (cobs != 1u) {
        const float T = c.lat_route[c.route[f]];
        // Exponential(1/m).rsample = e / rate  (airport.py:144-147)
        const float xd = c.value_mode ? nz.e_dep : BA_DIV(nz.e_dep, c.s_rate[o]);
        const float xa = c.value_mode ? nz.e_arr : BA_DIV(nz.e_arr, c.s_rate[d]);
        // Normal(T, T v_t).rsample = T + eps * (T v_t)  (network.py:158-163)
        const float sc = BA_MUL(T, c.v_t);
        const float tau = c.value_mode ? nz.eps_travel : BA_ADD(T, BA_MUL(nz.eps_travel, sc));
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
        const int lpos = c.pos_live[f];
        c.xdep[lpos] = xd;
        rec.lpos = (uint32_t)lpos | ((uint32_t)d << 16);
        rec.at = at; rec.xa = xa; rec.ka = (uint32_t)kk | ((c.ap[o].slot >> 5) << 16);
        BA_ATOMIC_INC(&c.kcnt[kk]);
    }
    ba2_st_rec(c.tmpf + f, rec);
}

// exclusive prefix over the tick buckets (one warp; host: lane = -1, serial)
BA_DEV void ba2_bucket_scan(Ba2Ctx &c, int lane)
{
#if defined(__CUDA_ARCH__)
    const uint32_t per = (BA2_KC + 31u) / 32u;
    const uint32_t b = (uint32_t)lane * per, e = min(b + per, (uint32_t)BA2_KC);
    uint32_t sum = 0;
    for (uint32_t k = b; k < e; ++k) sum += c.kcnt[k];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t run = incl - sum;
    for (uint32_t k = b; k < e; ++k) { const uint32_t n = c.kcnt[k]; c.kcnt[k] = run; run += n; }
#else
    (void)lane;
    uint32_t run = 0;
    for (uint32_t k = 0; k < BA2_KC; ++k) { const uint32_t n = c.kcnt[k]; c.kcnt[k] = run; run += n; }
#endif
}

// PROLOGUE sweep 2: flight -> its place in the tick-sorted order; histogram of destinations
// per range of that order
BA_DEV void ba2_scatter_flight(Ba2Ctx &c, int f)
{
    const uint32_t ka = c.tmpf[f].ka;
    if (ka == BA2_NONE) return;
    const uint32_t pos = BA_ATOMIC_INC(&c.kcnt[ka & 0xFFFFu]);
    c.idx1[pos] = (uint32_t)f;
    const uint32_t d = c.tmpf[f].lpos >> 16;
    BA_ATOMIC_INC(&c.hist[(pos / c.range_len) * (uint32_t)c.N + d]);
}

// counts -> start cursors (one thread per destination); writes the list's sentinel
BA_DEV void ba2_split_cursors(Ba2Ctx &c, int a)
{
    uint32_t run = c.ap[a].as_off;
    for (int r = 0; r < c.n_ranges; ++r) {
        const uint32_t n = c.hist[(uint32_t)r * (uint32_t)c.N + (uint32_t)a];
        c.hist[(uint32_t)r * (uint32_t)c.N + (uint32_t)a] = run;
        run += n;
    }
    Ba2Rec s; s.lpos = 0; s.at = 0.0f; s.xa = 0.0f; s.ka = BA2_NONE;
    ba2_st_rec(c.As + run, s);
}

// PROLOGUE sweep 3: stable split of range `r` by destination.  Device: one warp walks the
// range 32 records at a time (match.any ranks equal destinations inside the chunk).  Host:
// serial.
BA_DEV void ba2_split_range(Ba2Ctx &c, int r, int lane, uint32_t n_arr)
{
    const uint32_t lo = (uint32_t)r * c.range_len;
    uint32_t hi = lo + c.range_len;
    if (hi > n_arr) hi = n_arr;
    uint32_t *cur = c.hist + (uint32_t)r * (uint32_t)c.N;
#if defined(__CUDA_ARCH__)
    for (uint32_t base = lo; base < hi; base += 32u) {
        const uint32_t i = base + (uint32_t)lane;
        const bool ok = i < hi;
        const uint32_t act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const uint32_t f = c.idx1[i];
            Ba2Rec rec = ba2_ld_rec(c.tmpf + f);
            const uint32_t d = rec.lpos >> 16;
            const uint32_t m = __match_any_sync(act, d);
            const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
            const uint32_t pos = cur[d] + rank;
            __syncwarp(act);
            if (rank + 1u == (uint32_t)__popc(m)) cur[d] = pos + 1u;
            rec.lpos = (rec.lpos & 0xFFFFu) | (rec.ka & 0xFFFF0000u);
            rec.ka &= 0xFFFFu;
            ba2_st_rec(c.As + pos, rec);
            __syncwarp(act);
        }
    }
#else
    (void)lane;
    for (uint32_t i = lo; i < hi; ++i) {
        Ba2Rec rec = c.tmpf[c.idx1[i]];
        const uint32_t d = rec.lpos >> 16;
        rec.lpos = (rec.lpos & 0xFFFFu) | (rec.ka & 0xFFFF0000u);
        rec.ka &= 0xFFFFu;
        c.As[cur[d]++] = rec;
    }
#endif
}

// ---------------------------------------------------------------------------
// SCAN: one lane per airport, the lanes of a warp on a common clock, warps asynchronous
// ---------------------------------------------------------------------------
struct Ba2Lane {
    // arrival list of this airport: [as0, ha) served, [ha, w) queued (FIFO order),
    // [w, r) deferred (arrival tick passed, departure not served), [r, ...) unexamined
    uint32_t r, w, ha;
    uint32_t n_w0, n_ka;      // first and last word of As[r], fetched ahead
    Ba2Rec h;                 // As[ha] (h.ka == BA2_NONE: not loaded)
    // departures: [.., hd) served, [hd, td) queued; next calendar run (tick << 16 | records)
    uint32_t hd, td, rp, rtc;
    uint32_t dk;              // ready tick of lq[hd]
    float dqs, dx;            // its queue start and service draw
    // service state
    uint32_t asg;             // 0: head not assigned, 1: arrival assigned, 2: departure assigned
    float A;                  // assigned service time of the head (airport.py:156)
    float wacc;               // fold of the draws assigned since group g joined
    uint32_t g;               // joining tick of the group the accumulator belongs to
    uint32_t gnew;            // joining tick of the youngest group in the queue
    uint32_t cnt;             // assignments made so far
    uint32_t pk;              // ordering key of As[w - 1] (current block)
    uint32_t left;            // entries still to serve
};

// Rarely touched lane state.  Its address is taken by the out-of-line slow paths, so it
// lives in local memory; everything in Ba2Lane stays in registers.
struct Ba2Cold {
    uint32_t hxo;             // this airport's slice of hx / glog
    uint32_t lw, lr;          // log of (tick, cnt) pairs: written / consumed
    uint32_t status;
};

// Block of this tick's joiners while an out-of-line path works on it.
struct Ba2Blk {
    uint32_t w, hka, pk;
};

BA_DEV void ba2_lane_init(const Ba2Ctx &c, Ba2Lane &L, Ba2Cold &Q, int a)
{
    const BaAirport A = c.ap[a];
    L.r = L.w = L.ha = A.as_off;
    { const Ba2Rec n = ba2_ld_rec(c.As + L.r); L.n_w0 = n.lpos; L.n_ka = n.ka; }
    L.h.lpos = 0; L.h.at = 0.0f; L.h.xa = 0.0f; L.h.ka = BA2_NONE;
    L.hd = L.td = A.lq_off;
    L.rp = (uint32_t)c.drun_off[a];
    L.rtc = (c.drun[2 * L.rp] << 16) | (c.drun[2 * L.rp + 1] & 0xFFFFu);
    { const BaLiveDep q = c.lq[L.hd]; L.dk = q.tick; L.dqs = q.qstart; }
    L.dx = c.xdep[L.hd];
    L.asg = 0; L.A = 0.0f; L.wacc = 0.0f; L.g = BA2_NONE; L.gnew = BA2_NONE; L.cnt = 0;
    Q.hxo = A.lq_off + A.as_off - 2u * (uint32_t)a;
    Q.lw = Q.lr = 0; Q.status = 0; L.pk = 0;
    L.left = (uint32_t)(A.dep_live + A.arr_cnt);
}

// in-transit list order of an arrival: (tick at which its departure was served, origin,
// position in the origin's queue) = (q, list position of the departure)
BA_DEV uint32_t ba2_key(uint32_t q, uint32_t lpos) { return (q << 16) | lpos; }

// ---- waiting for another warp ----------------------------------------------------------
// Every warp owns an mbarrier it arrives on once per tick (after publishing the tick in
// prog[]); a warp that needs another one to get further suspends on that barrier's current
// phase instead of spinning.  The host emulation yields.
#if defined(__CUDACC__)
BA_DEV void ba2_bar_init(uint64_t *bar, int first)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    if (!first) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a) : "memory");
}
BA_DEV void ba2_bar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
// suspends until the phase of that parity has completed (the hardware wakes the thread on
// any barrier traffic of the SM or after the time limit: stay in the short loop; bounded,
// the caller re-reads the progress word and comes back)
BA_DEV void ba2_bar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
    for (int i = 0; i < 64; ++i) {
        uint32_t ok;
        asm volatile(
            "{\n .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            " selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity), "r"(100000u) : "memory");
        if (ok) break;
    }
}
#else
BA_DEV void ba2_bar_init(uint64_t *, int) {}
BA_DEV void ba2_bar_arrive(uint64_t *) {}
BA_DEV void ba2_bar_wait(uint64_t *, uint32_t) { std::this_thread::yield(); }
#endif

// The departure behind `word` (list position | origin's warp << 16) shows "not served".
// That is final for tick km1 only once the origin's warp has finished that tick.
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
uint32_t ba2_resolve(volatile const int *prog, uint64_t *bars, volatile const uint16_t *qsh,
                     uint32_t word, uint32_t km1, uint32_t t0)
{
    const uint32_t lpos = word & 0xFFFFu, ow = word >> 16;
    for (;;) {
        const int p = prog[ow];
        BA2_FENCE();
        const uint32_t q = (uint32_t)qsh[lpos];
        if (q != BA2_Q_INVALID || p >= (int)km1) return q;
        // the origin's warp is working on tick p + 1: suspend until that phase completes
        ba2_bar_wait(bars + ow, ((uint32_t)p - t0) & 1u);
    }
}

// Inserts `rec` (its departure was served at tick q <= km1) into the block [wb, w) of this
// tick's joiners, kept sorted by in-transit list order; the slot As[w] is free.
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
void ba2_insert(Ba2Rec *As, volatile const uint16_t *qsh, Ba2Blk &B, Ba2Rec rec, uint32_t key,
                uint32_t wb, uint32_t ha)
{
    uint32_t j = B.w;
    while (j > wb) {
        const Ba2Rec e = ba2_ld_rec(As + j - 1);
        const uint32_t el = e.lpos & 0xFFFFu;
        if (ba2_key((uint32_t)qsh[el], el) <= key) break;
        ba2_st_rec(As + j, e);
        --j;
    }
    ba2_st_rec(As + j, rec);
    if (key > B.pk || B.w == wb) B.pk = key;
    if (ha >= wb) B.hka = BA2_NONE;                 // the cached head may have moved
    B.w += 1;
}

// Deferred zone [w, r): flights whose arrival tick has passed but whose departure had not
// been served.  Those served by the end of tick km1 join now (late path, network.py:186).
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
bool ba2_recheck_deferred(Ba2Rec *As, volatile const uint16_t *qsh, volatile const int *prog,
                          uint64_t *bars, Ba2Blk &B, uint32_t r, uint32_t km1, uint32_t wb,
                          uint32_t ha, uint32_t t0)
{
    bool joined = false;
    for (uint32_t i = B.w; i < r; ++i) {
        Ba2Rec rec = ba2_ld_rec(As + i);
        const uint32_t lpos = rec.lpos & 0xFFFFu;
        uint32_t q = (uint32_t)qsh[lpos];
        if (q == BA2_Q_INVALID) q = ba2_resolve(prog, bars, qsh, rec.lpos, km1, t0);
        if (q > km1) continue;
        // the vacated slot i takes the (already examined) record at w
        if (i != B.w) ba2_st_rec(As + i, ba2_ld_rec(As + B.w));
        rec.ka = km1;                               // joined at the end of tick km1
        ba2_insert(As, qsh, B, rec, ba2_key(q, lpos), wb, ha);
        joined = true;
    }
    return joined;
}

// A new group (joining tick p) reaches the head: its waits are the fold of the draws
// assigned since it joined (types/airport.py:150-153) -- replay them from the history.
#if defined(__CUDACC__)
__device__ __noinline__
#else
static
#endif
float ba2_replay(const uint32_t *glog, const float *hxb, Ba2Cold &Q, uint32_t cnt, uint32_t p,
                 uint32_t k, uint32_t cnt0)
{
    uint32_t lo = cnt0;
    if (p != k) {
        const uint32_t *lg = glog + 2u * Q.hxo;
        while (Q.lr < Q.lw && lg[2u * Q.lr] != p) Q.lr += 1;
        if (Q.lr == Q.lw) { Q.status |= BA_S_INTERNAL; return 0.0f; }
        lo = lg[2u * Q.lr + 1u];
        Q.lr += 1;
    }
    float w = 0.0f;
    const float *hx = hxb + Q.hxo;
    for (uint32_t j = lo; j < cnt; ++j) w = BA_ADD(w, hx[j]);
    return w;
}

// One tick of one airport: joins (model.py:181-183,186-196), then update_runway_queue
// (types/airport.py:86-128).
BA_DEV void ba2_lane_tick(const Ba2Ctx &c, Ba2Lane &L, Ba2Cold &Q, uint32_t k, float t)
{
    const uint32_t km1 = k - 1u;
    bool joined = false;
    const uint32_t wb = L.w;
    // ---- arrivals that joined at the end of tick k - 1 ------------------------------
    if (L.w != L.r) {
        Ba2Blk B; B.w = L.w; B.hka = L.h.ka; B.pk = L.pk;
        joined = ba2_recheck_deferred(c.As, c.qsh, c.prog, c.bars, B, L.r, km1, wb, L.ha,
                                      (uint32_t)(c.first_tick - 1));
        L.w = B.w; L.h.ka = B.hka; L.pk = B.pk;
    }
    while (L.n_ka <= km1) {
        const uint32_t lpos = L.n_w0 & 0xFFFFu;
        uint32_t q = (uint32_t)c.qsh[lpos];
        if (q == BA2_Q_INVALID)
            q = ba2_resolve(c.prog, c.bars, c.qsh, L.n_w0, km1, (uint32_t)(c.first_tick - 1));
        if (q <= km1) {
            const uint32_t key = ba2_key(q, lpos);
            if (L.w == L.r && (L.w == wb || key >= L.pk)) {
                L.pk = key; L.w += 1;                        // already in place
            } else {
                const Ba2Rec rec = ba2_ld_rec(c.As + L.r);
                if (L.w != L.r) ba2_st_rec(c.As + L.r, ba2_ld_rec(c.As + L.w));
                Ba2Blk B; B.w = L.w; B.hka = L.h.ka; B.pk = L.pk;
                ba2_insert(c.As, c.qsh, B, rec, key, wb, L.ha);
                L.w = B.w; L.h.ka = B.hka; L.pk = B.pk;
            }
            joined = true;
        }
        L.r += 1;
        { const Ba2Rec n = ba2_ld_rec(c.As + L.r); L.n_w0 = n.lpos; L.n_ka = n.ka; }
        BA2_PREFETCH(c.As + L.r + 2);
    }
    // ---- departures that become ready at tick k (static calendar run) -----------------
    if ((L.rtc >> 16) == k) {
        L.td += L.rtc & 0xFFFFu;
        L.rp += 1;
        L.rtc = (c.drun[2 * L.rp] << 16) | (c.drun[2 * L.rp + 1] & 0xFFFFu);
        joined = true;
    }
    if (joined) L.gnew = k;
    // ---- service loop -----------------------------------------------------------------
    const uint32_t cnt0 = L.cnt;
    for (;;) {
        if (L.asg == 0u) {
            const bool ha_ok = L.ha < L.w, hd_ok = L.hd < L.td;
            if (!ha_ok && !hd_ok) break;
            if (ha_ok && L.h.ka == BA2_NONE) L.h = ba2_ld_rec(c.As + L.ha);
            const uint32_t pa = ha_ok ? L.h.ka + 1u : BA2_NONE;
            const uint32_t pd = hd_ok ? L.dk : BA2_NONE;
            const bool arr = pa <= pd;                       // arrivals first on ties
            const uint32_t p = arr ? pa : pd;
            const float x = arr ? L.h.xa : L.dx;
            const float qs = arr ? L.h.at : L.dqs;
            if (p != L.g) {
                L.wacc = (p == k && L.cnt == cnt0) ? 0.0f : ba2_replay(c.glog, c.hx, Q, L.cnt, p, k, cnt0);
                L.g = p;
            }
            if (L.gnew != p) c.hx[Q.hxo + L.cnt] = x;        // a younger group is waiting
            L.cnt += 1;
            L.wacc = BA_ADD(L.wacc, x);                      // airport.py:150-153
            L.A = BA_ADD(qs, L.wacc);                        // airport.py:156
            L.asg = arr ? 1u : 2u;
        }
        if (!(L.A <= t)) break;                              // airport.py:111
        if (L.asg == 1u) {
            const uint32_t lpos = L.h.lpos & 0xFFFFu;
            c.wt[2u * lpos + 1u] = L.wacc;
            if (c.pop_tick) c.qarr[lpos] = k;
            L.ha += 1;
            if (L.ha < L.w) L.h = ba2_ld_rec(c.As + L.ha); else L.h.ka = BA2_NONE;
            BA2_PREFETCH(c.As + L.ha + 2);
        } else {
            c.wt[2u * L.hd] = L.wacc;
            c.qsh[L.hd] = (uint16_t)k;
            L.hd += 1;
            { const BaLiveDep q = c.lq[L.hd]; L.dk = q.tick; L.dqs = q.qstart; }
            L.dx = c.xdep[L.hd];
            BA2_PREFETCH(c.lq + L.hd + 4);
            BA2_PREFETCH(c.xdep + L.hd + 8);
        }
        L.left -= 1; L.asg = 0u;
    }
    // a tick that ends with entries still waiting behind an assigned head: the entries that
    // joined in it will need the number of assignments made before it (ba2_replay)
    if (joined && L.asg != 0u) {
        uint32_t *lg = c.glog + 2u * (Q.hxo + Q.lw);
        lg[0] = k; lg[1] = cnt0;
        Q.lw += 1;
    }
}

// ---------------------------------------------------------------------------
// EPILOGUE (flight-parallel, then owner sums); SURVEY.md App. A.3 / A.5
// ---------------------------------------------------------------------------
BA_DEV void ba2_epilogue_flight(Ba2Ctx &c, BaAcc &acc, int f)
{
    const BaFlight fl = ba2_flight(c, f);
    const uint32_t cobs = (fl.od >> 24) & 3u;
    const int k = c.pos_dep[f];
    const int o = (int)(fl.od & 0xFFFu), d = (int)((fl.od >> 12) & 0xFFFu);
    // *_cancelled (network.py:100-109): the cancellation probability is exactly 0 at every
    // airport of such a plan (ba_plan.h: BA_SIMPLE_RATIO)
    acc.lp += (double)(cobs == 1u ? c.c1_lp : c.c0_lp);
    if (cobs == 1u) {                 // observed as cancelled: never queued anywhere
        if (c.want_grad) { c.xdk[k] = 0.0f; c.atk[f] = 0.0f; }
        if (c.pop_tick) { c.pop_tick[2 * f] = -1; c.pop_tick[2 * f + 1] = -1; }
        return;
    }
    const int j = c.pos_arr[f];
    const int lpos = c.pos_live[f];
    BaNoise4 nz;
    if (c.noise) nz = ba2_noise(c, f);
    else {
#if defined(__CUDA_ARCH__)
        const float4 v = reinterpret_cast<const float4 *>(c.nzs)[f];
        nz.e_dep = v.x; nz.eps_travel = v.y; nz.e_arr = v.z; nz.eps_turn = v.w;
#else
        nz.e_dep = c.nzs[4 * f]; nz.eps_travel = c.nzs[4 * f + 1]; nz.e_arr = c.nzs[4 * f + 2]; nz.eps_turn = c.nzs[4 * f + 3];
#endif
    }
    if (c.pop_tick) {
        const uint32_t qd = (uint32_t)c.qsh[lpos];
        c.pop_tick[2 * f] = qd == BA2_Q_INVALID ? -1 : (int32_t)qd;
        c.pop_tick[2 * f + 1] = (int32_t)c.qarr[lpos];
    }
    const float T = c.lat_route[c.route[f]];
    const bool grad = c.want_grad != 0;
    const float half_inv_var = 0.5f * c.inv_sigma2;
    const float lp_const = -c.log_sigma - BA_LOG_SQRT_2PI;
    float y_dep = fl.act_dep, y_arr = fl.act_arr;
    if (cobs == 3u || isnan(y_dep) || isnan(y_arr)) { y_dep = fl.sched_dep; y_arr = fl.sched_dep; }

    // departure side: *_departure_service_time, *_simulated_departure_time
    // (airport.py:144-147,175-182)
    const float rate_o = c.s_rate[o];
    const float xd = c.value_mode ? nz.e_dep : BA_DIV(nz.e_dep, rate_o);
    const float wdep = c.wt[2 * lpos];
    const float qstart_d = fmaxf(0.0f, fl.sched_dep);
    const float mu_d = BA_ADD(qstart_d, wdep);              // queue_start + total_wait
    const float dy_d = y_dep - mu_d;
    float lp = (BA_LOG(rate_o) - rate_o * xd) + (lp_const - dy_d * dy_d * half_inv_var);

    // arrival side: *_travel_time, *_arrival_service_time, *_simulated_arrival_time,
    // *_turnaround_time (network.py:158-163; airport.py:144-147,197-204,217-220)
    const float rate_d = c.s_rate[d], mt = c.s_mt[d], tstd = c.s_tstd[d];
    const float sc = BA_MUL(T, c.v_t);
    const float tau = c.value_mode ? nz.eps_travel : BA_ADD(T, BA_MUL(nz.eps_travel, sc));
    const float qstart_a = BA_ADD(y_dep, tau);
    const float warr = c.wt[2 * lpos + 1];
    const float mu_a = BA_ADD(qstart_a, warr);
    const float dy_a = y_arr - mu_a;
    const float xa = c.value_mode ? nz.e_arr : BA_DIV(nz.e_arr, rate_d);
    const float u = c.value_mode ? nz.eps_turn : BA_ADD(mt, BA_MUL(nz.eps_turn, tstd));
    const float dz = (tau - T) / sc;
    const float du = (u - mt) / tstd;
    lp += (BA_LOG(rate_d) - rate_d * xa) + (lp_const - dy_a * dy_a * half_inv_var);
    const float lp2 = (-0.5f * dz * dz - BA_LOG(sc) - BA_LOG_SQRT_2PI)
                    + (-0.5f * du * du - BA_LOG(tstd) - BA_LOG_SQRT_2PI);
    acc.lp += (double)lp + (double)lp2;

    if (grad) {
        const float r_d = dy_d * c.inv_sigma2, r_a = dy_a * c.inv_sigma2;
        acc.g_sigma += (dy_d * dy_d + dy_a * dy_a) * c.inv_sigma3 - 2.0f * c.inv_sigma;
        if (!c.value_mode) {
            c.xdk[k] = r_d * wdep;
            c.xak[j] = r_a * warr;
            // arrival queue_start = act_dep + tau, tau = T (1 + v_t eps)
            c.atk[f] = r_a * (1.0f + c.v_t * nz.eps_travel) - 1.0f / T;
            acc.g_vt += r_a * nz.eps_travel * T - c.inv_vt;
            acc.g_vu += -c.inv_vu;
        } else {
            c.xdk[k] = xd;
            c.xak[j] = xa;
            c.atk[f] = dz / sc + (dz * dz - 1.0f) / T;
            acc.g_vt += (dz * dz - 1.0f) * c.inv_vt;
            c.ct[j] = du / tstd + (du * du - 1.0f) / mt;     // arrival order: no other writer
            acc.g_vu += (du * du - 1.0f) * c.inv_vu;
        }
    }
}

BA_DEV void ba2_epilogue_airport(Ba2Ctx &c, int a, int Ng)
{
    const BaAirport A = c.ap[a];
    const float m = c.s_m[a], mt = c.s_mt[a];
    float s_dep = 0.0f, s_arr = 0.0f, s_turn = 0.0f;
    for (int i = 0; i < A.dep_cnt; ++i) s_dep += c.xdk[A.dep_off + i];
    for (int i = 0; i < A.arr_cnt; ++i) s_arr += c.xak[A.arr_off + i];
    const float n_ent = (float)(A.dep_live + A.arr_cnt), n_land = (float)A.arr_cnt;
    float g_serv, g_turn;
    if (!c.value_mode) {
        // x = e m  =>  d/dm = sum r W / m - (#entries)/m ;  u = m~ (1 + v_u eps)
        g_serv = (s_dep + s_arr) / m - n_ent / m;
        g_turn = -n_land / mt;
    } else {
        for (int i = 0; i < A.arr_cnt; ++i) s_turn += c.ct[A.arr_off + i];
        // log Exponential(x; 1/m) = -log m - x/m
        g_serv = -n_ent / m + (s_dep + s_arr) / (m * m);
        g_turn = s_turn;
    }
    const int g = A.global;
    c.grow[0 * Ng + g] = g_turn;
    c.grow[1 * Ng + g] = g_serv;
}

BA_DEV void ba2_epilogue_route(Ba2Ctx &c, int r)
{
    float s = 0.0f;
    const int b = c.rt_off[r], e = c.rt_off[r + 1];
    for (int i = b; i < e; ++i) s += c.atk[c.rt_flt[i]];
    c.grow[c.route_base + c.rt_gid[r]] = s;
}
