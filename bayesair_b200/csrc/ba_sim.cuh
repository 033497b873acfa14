// Core of the hot path: one (particle, day) of the air-traffic-network simulation
// with the log-joint and its pathwise gradient fused into the scan.
//
// Decomposition: one CTA per (particle, day), three stages.
//
//   PROLOGUE (flight-parallel, coalesced 16-byte loads of the flight table and the
//       noise tensor): everything about a flight that does not depend on queue
//       dynamics -- its two service-time draws, its travel time and hence (the
//       departure time being observed) the time it reaches its destination and the
//       tick at which that time passes -- is derived once and written to L2-resident
//       per-CTA scratch; the flights are counting-sorted by that tick (the ARRIVAL
//       CALENDAR, exact: it uses the same fp32 clock table as the scan).
//   SCAN (one lane per airport; the serial part): the tick loop of model.py:158-200.
//       It touches shared memory only, apart from fire-and-forget stores of each
//       entry's total wait W (the "tape") and of the in-transit ordering keys.
//       phase A (per airport, no cross-airport reads): turnaround release -> reserve
//           injection -> pop-ready / cancellation -> enqueue departures -> runway
//           service loop (model.py:163-196).  A served departure records its
//           in-transit ordering key (tick, origin, per-origin sequence).
//       phase B (flight-parallel): the reference's in-transit list (network.py:178-198)
//           is never materialised.  The flights of this tick's calendar slice whose
//           departure has been served get a slot in a small "due" buffer; a departure
//           served at or after its own arrival tick is put there by phase A itself.
//       phase C (owner pull, at the start of the next tick): each airport lane takes
//           the due entries addressed to it in key order -- this IS the reference's
//           in-transit list order restricted to that destination (model.py:186-196) --
//           appends them to its runway queue, then its newly ready departures.
//       Shared-list kernel: one barrier per tick -- phase B is a set of asynchronous
//       copies issued next to the lanes' work (see "SINGLE-BARRIER SCAN" below).
//       Exact-list / predictive kernels: lanes, barrier, phase B, barrier.
//   EPILOGUE (flight-parallel, then one owner per airport / route over static
//       slices): mu = queue_start + W for every entry, the seven per-flight
//       log-densities and the pathwise gradient sums.  Deterministic: every sum has a
//       single owner and a fixed order; no floating-point atomics anywhere.
//
// All arithmetic that feeds a decision (a `<=` against the clock) is IEEE fp32 in
// the reference's operation order, with contraction disabled (BA_ADD / BA_MUL /
// BA_DIV), see SURVEY.md H1.
//
// The same source is compiled (a) by nvcc into the sm_100a kernel and (b) by g++
// into tests/hostemu (a debugging harness that runs the phases lane by lane on the
// CPU; it is not part of the product and not reachable from the Python package).
#pragma once
#include <math.h>
#include <stdint.h>
#if !defined(__CUDACC__)
#include <vector>
#endif

#include "ba_plan.h"

#if defined(__CUDACC__)
#define BA_DEV __device__ __forceinline__
#else
#define BA_DEV static inline
#endif
#if defined(__CUDACC__)
#define BA_MEMBER __device__ __forceinline__
#else
#define BA_MEMBER inline
#endif

#if defined(__CUDA_ARCH__)
#define BA_ADD(a, b) __fadd_rn((a), (b))
#define BA_MUL(a, b) __fmul_rn((a), (b))
#define BA_DIV(a, b) __fdiv_rn((a), (b))
#define BA_ATOMIC_INC(p) atomicAdd((p), 1u)
#define BA_ATOMIC_OR(p, v) atomicOr((p), (v))
#define BA_ATOMIC_AND(p, v) atomicAnd((p), (v))
#define BA_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define BA_ATOMIC_DEC(p) atomicSub((p), 1u)
#define BA_ATOMIC_ADD64(p, v) atomicAdd((p), (unsigned long long)(v))
#define ba_ctz64(x) (__ffsll((long long)(x)) - 1)
#define __float_as_uint_portable(x) __float_as_uint(x)
#define __uint_as_float_portable(x) __uint_as_float(x)
#define BA_LOG(x) logf(x)
#define BA_EXP(x) expf(x)
#define BA_LOG1P(x) log1pf(x)
#else
// host emulation: build with -ffp-contract=off
#define BA_ADD(a, b) ((float)(a) + (float)(b))
#define BA_MUL(a, b) ((float)(a) * (float)(b))
#define BA_DIV(a, b) ((float)(a) / (float)(b))
#include <string.h>
static inline uint32_t ba_host_inc(uint32_t *p) { return (*p)++; }
static inline uint32_t ba_host_dec(uint32_t *p) { return (*p)--; }
static inline uint32_t __float_as_uint_portable(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static inline float __uint_as_float_portable(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
#define BA_ATOMIC_INC(p) ba_host_inc(p)
#define BA_ATOMIC_OR(p, v) (*(p) |= (v))
#define BA_ATOMIC_AND(p, v) (*(p) &= (v))
static inline uint32_t ba_host_add(uint32_t *p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
#define BA_ATOMIC_ADD(p, v) ba_host_add((p), (v))
#define BA_ATOMIC_DEC(p) ba_host_dec(p)
#define BA_ATOMIC_ADD64(p, v) (*(p) += (unsigned long long)(v))
#define ba_ctz64(x) __builtin_ctzll((unsigned long long)(x))
#define BA_LOG(x) logf(x)
#define BA_EXP(x) expf(x)
#define BA_LOG1P(x) log1pf(x)
#endif

// status bits (mirror include/bayesair_b200.h)
#define BA_S_TICK_CAP 1u
#define BA_S_OVERFLOW 2u
#define BA_S_UNOBSERVED 4u
#define BA_S_INTERNAL 8u

#define BA_LOG_SQRT_2PI 0.918938533204672741780329736406f
#define BA_F32_EPS 1.1920928955078125e-07f
#define BA_F32_TINY 1.17549435082228750797e-38f

struct BaNoise4 { float e_dep, eps_travel, e_arr, eps_turn; };

// 8-byte pair accesses (LDS.64 / STS.64 on the device)
struct alignas(8) BaPairU { uint32_t a, b; };
struct alignas(8) BaPairF { float a, b; };
struct alignas(8) BaPairUF { uint32_t a; float b; };
#if defined(__CUDA_ARCH__)
#define BA_LD_PAIR(T, p) (*reinterpret_cast<const T *>(p))
#define BA_ST_PAIR(T, p, v) (*reinterpret_cast<T *>(p) = (v))
#else
template <typename T> static inline T ba_ld_pair(const uint32_t *p) { T v; memcpy(&v, p, 8); return v; }
template <typename T> static inline void ba_st_pair(uint32_t *p, T v) { memcpy(p, &v, 8); }
#define BA_LD_PAIR(T, p) ba_ld_pair<T>(p)
#define BA_ST_PAIR(T, p, v) ba_st_pair<T>(p, v)
#endif

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011): counter = (flight, day, particle lo, hi),
// key = seed.  One call yields the four per-flight draws.
// ---------------------------------------------------------------------------
BA_DEV void ba_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                      uint32_t k1, uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

BA_DEV float ba_u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

BA_DEV BaNoise4 ba_philox_noise(uint64_t seed, uint32_t particle, uint32_t day, uint32_t flight)
{
    uint32_t r[4];
    ba_philox(flight, day, particle, 0x42413230u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    BaNoise4 z;
    float u1 = ba_u01(r[0]), u2 = ba_u01(r[1]), u3 = ba_u01(r[2]), u4 = ba_u01(r[3]);
    z.e_dep = -BA_LOG(u1);
    z.e_arr = -BA_LOG(u4);
    float rad = sqrtf(-2.0f * BA_LOG(u2));
    float ang = 6.283185307179586f * u3;
    z.eps_travel = rad * cosf(ang);
    z.eps_turn = rad * sinf(ang);
    return z;
}

// ---------------------------------------------------------------------------
// torch.distributions arithmetic (fp32)
// ---------------------------------------------------------------------------
BA_DEV float ba_softplus(float x) { return x > 20.0f ? x : BA_LOG1P(BA_EXP(x)); }

// RelaxedBernoulli(temperature 0.5, probs=p).log_prob(v), and d/dp
// (bayes_air/network.py:102-109; torch relaxed_bernoulli.py + SigmoidTransform)
BA_DEV float ba_relaxed_bernoulli_lp(float p, float v, float *dlp_dp)
{
    float q = fminf(fmaxf(p, BA_F32_EPS), 1.0f - BA_F32_EPS);
    float logits = BA_LOG(q) - BA_LOG1P(-q);
    float y = fminf(fmaxf(v, BA_F32_TINY), 1.0f - BA_F32_EPS);
    float x = BA_LOG(y) - BA_LOG1P(-y);
    float diff = logits - 0.5f * x;
    float ed = BA_EXP(diff);
    float lp = -0.69314718055994530942f + diff - 2.0f * BA_LOG1P(ed);
    lp += ba_softplus(-x) + ba_softplus(x);
    if (dlp_dp) {
        float pass = (p >= BA_F32_EPS && p <= 1.0f - BA_F32_EPS) ? 1.0f : 0.0f;
        float sig = isinf(ed) ? 1.0f : ed / (1.0f + ed);
        *dlp_dp = pass * (1.0f / q + 1.0f / (1.0f - q)) * (1.0f - 2.0f * sig);
    }
    return lp;
}

// ---------------------------------------------------------------------------
// per-CTA context
// ---------------------------------------------------------------------------
// One field of the per-airport records (see ba_carve_state): indexed by airport.
template <typename T> struct BaCol {
    T *p;
    uint32_t s;
    BA_MEMBER T &operator[](int a) const { return p[(uint32_t)a * s]; }
    BA_MEMBER T &operator[](uint32_t a) const { return p[a * s]; }
};

struct BaCtx {
    // static day tables
    int N, F;
    const BaFlight *flights;
    const int32_t *route;
    const float *dep_sched;
    const uint32_t *dep_word;
    const int32_t *pos_dep;
    const int32_t *pos_arr, *rt_gid, *rt_off, *rt_flt;
    int Rd;
    const BaAirport *ap;
    // particle inputs
    const float *lat_route;       // [R]
    const float *noise;           // [Fmax*4] of this (particle, day) or nullptr
    // posterior-predictive mode only (flights with obs=None are sampled):
    const float *noise2;          // [Fmax*4] U(0,1) cancel, N(0,1) departure, N(0,1) arrival, -
    float *sim;                   // [Fmax*3] out: cancelled, simulated departure, arrival
    uint32_t *pool;               // in-transit pool {word, key, seq, arrival, x} x F (aliases
                                  //   the calendar arrays, unused in this mode)
    uint64_t seed;
    uint32_t particle, dayidx;
    float sigma, v_t, v_u;        // the three scales
    float inv_sigma2, inv_sigma3, inv_sigma, log_sigma;
    float inv_vt, inv_vu;
    float c0_lp, c1_lp;           // cancellation log-prob at p == 0 for v = 0 / 1
    int value_mode, want_grad, cancel_mode, replay_always;
    float max_t;
    // per-airport state (shared memory), indexed by day-local airport
    BaCol<float> m, rate, mt, tstd, base, n_avail, n0, next_sched;
    BaCol<float> asg_lo, asg_hi, w_head;  // head entry: interval holding its exact fp32
                                      //   service time (NaN = unassigned) and its total wait
    double *psum;                     // running sum of all service draws assigned so far
    BaCol<float> acc_ln0, acc_bc;
    BaCol<uint32_t> next_dep, q_head, q_tail, q_fill, dep_pops, turn_n;
    BaCol<uint32_t> arr_left;             // arrivals this airport still waits for
    BaCol<uint32_t> run_ptr, run_tick;    // next run of the departure calendar, its tick,
    BaCol<uint32_t> run_info;             //   first staged record << 16 | records
    BaCol<uint32_t> inmask;               // two words per due buffer (even: [0], [1]; odd: [2], [3]):
                                          //   bit s = slot s < 64 of that buffer's slice is
                                          //   addressed to this airport
    BaCol<uint32_t> av_head, av_tail, zeros_left, dl_head, dl_tail;
    BaCol<uint32_t> s_flags, s_qw, s_goff, s_depoff, s_depcnt;
    // CTA-wide scalars (shared memory): [0], [1] entries in the even / odd due buffer,
    // [3] abort flag (a shared ring overflowed), [2] / [4] entries in / first live entry of
    // the predictive in-transit pool; single-barrier scan: [8], [9], [10] late entries of the
    // due buffer of tick j at [8 + j % 3]
    uint32_t *cta;
    const int32_t *dcal_off;      // static departure calendar (ba_plan.h)
    const uint32_t *dcal_rec;
    const int32_t *dep_cpos;      // dep order -> calendar record
    float *nzs;                   // [4 F] generated noise kept for the epilogue (Philox mode)
    uint32_t *dpair;              // [2 F] {x_dep, arrival time} in calendar-record order
                                  //   (global scratch): the per-particle half of a staged entry
    const int32_t *drun_off;      // per-airport runs of the departure calendar
    const uint32_t *drun;
    int last_ready_tick;
    // staged departures of the next tick (two buffers, 4 words per record, in calendar-slice
    // order): shared (fast variant) or global (exact variant)
    uint32_t *dstage_base;
    uint32_t dstage_stride, dstage_cap;
    // arrival calendar
    const float *tick_time;       // [max_ticks + 2]
    uint32_t *cal_cnt;            // [cal_cap] bucket end offsets (shared or global)
    uint32_t cal_cap;             // ticks covered by the calendar
    uint32_t *karr;               // [F] arrival tick of every flight (global scratch)
    uint32_t *cal;                // [3F] records {word, arrival time, x_arr} sorted by tick
    uint32_t *popkey, *popseq;    // [F] ordering key of served departures (global scratch)
    // two due buffers (even / odd ticks), 6 words per entry: {queue word, key}
    // {owner << 16 | seq, queue start} {x, y}
    uint32_t *due_base;           // buffer p at due_base + p * due_stride (no pointer arrays:
    uint32_t due_stride;          //   a runtime-indexed member would push BaCtx to local memory)
    uint32_t due_cap;
    // single-barrier scan (shared-list variant): queue words of the next arrival slice, two
    // buffers of BA_DUE_CAP_S words, staged a tick ahead
    uint32_t *wbuf;
    int first_tick;
    // list pools (shared memory in the fast variant, global scratch in the exact one)
    // runway queues, 6 words per entry (three 8-byte pairs): {word, lo} {queue start,
    // prefix sum at joining} {own service draw x, arrival time at the destination}
    uint32_t *qe;
    // spill area (global scratch, fast variant): when an airport's shared ring is full, new
    // entries go to gq at their absolute index and are refilled into the ring as the head
    // advances.  q_fill[a] = first index not (yet) present in the shared ring.
    uint32_t *gq;
    float *gx;                    // service draw of every pushed entry at its absolute index
                                  //   (global): the history the exact replay reads
    // per-flight global scratch
    float *xdk, *xak, *atk;       // prologue outputs, dep order (reused by the epilogue)
    float *xaf;                   // x_arr by flight (for departures served after arrival)
    float *atf;                   // arrival time at the destination by flight (observed mode)
    float *ct;                    // epilogue: per-flight d/d mean_turnaround contributions
    float *wd, *wa;               // tape: total wait of the departure / arrival entry
    float *trl, *tre;             // turnaround release time and its noise (tracked days)
    float *qd, *se, *sw;          // departure queue start, aircraft-source noise, weight
    // always-global lists
    float *tt, *te;               // turnaround lists
    float *at, *ae;               // aircraft FIFOs
    uint32_t *dl;                 // delayed deques
    // outputs
    float *grow;                  // gradient row [C*N + R + 3] or nullptr
    int32_t *pop_tick;            // [Fmax*2] of this (particle, day) or nullptr
    int route_base;               // C*N
};

struct BaAcc {
    double lp;                    // ~7F terms of mixed sign per particle-day: keep the sum exact-ish
    float g_sigma, g_vt, g_vu;
    uint32_t status;
};

// Per-airport shared state is an array of records (one per airport, odd word stride so
// that consecutive lanes hit different banks): every field of an airport is then the
// airport's record address plus a compile-time offset, which costs no address arithmetic
// per access.  Only psum (8-byte) lives in its own array in front of the records.
#define BA_REC_WORDS_SIMPLE 23    // 23 words, no aircraft bookkeeping
#define BA_REC_WORDS_TRACK 37     // 37 words, some airport of the plan tracks aircraft
#define BA_NWORDS_STATE_SIMPLE (2 + BA_REC_WORDS_SIMPLE)
#define BA_NWORDS_STATE_TRACK (2 + BA_REC_WORDS_TRACK)
#define BA_Q_WORDS 6              // words per runway-queue entry
#define BA_CTA_WORDS 12           // CTA-wide scalars

// Ring placement of an airport in one word: first entry of its ring in the pool (27 bits)
// and log2 of the ring's power-of-two capacity (5 bits).
BA_DEV uint32_t ba_qw_pack(uint32_t off, uint32_t mask)
{
#if defined(__CUDA_ARCH__)
    const uint32_t lg = (uint32_t)__popc(mask);          // mask = 2^lg - 1
#else
    uint32_t lg = 0;
    while ((1u << lg) <= mask && lg < 31u) ++lg;
#endif
    return off | (lg << 27);
}
BA_DEV uint32_t ba_qw_off(uint32_t qw) { return qw & 0x07FFFFFFu; }
BA_DEV uint32_t ba_qw_mask(uint32_t qw) { return (1u << (qw >> 27)) - 1u; }

// Assigns the per-airport fields out of one contiguous word buffer.  The fields that only
// airports with aircraft bookkeeping touch exist only when the plan has such airports.
BA_DEV void ba_carve_state(BaCtx &c, uint32_t *w, int N, bool tracked)
{
    c.psum = (double *)w;                      // 2 N words, 8-byte aligned
    w += 2 * N;
    float *f = (float *)w;
    const uint32_t s = tracked ? BA_REC_WORDS_TRACK : BA_REC_WORDS_SIMPLE;
#define BA_COLF(name, k) c.name.p = f + (k); c.name.s = s
#define BA_COLU(name, k) c.name.p = w + (k); c.name.s = s
    BA_COLF(m, 0); BA_COLF(rate, 1); BA_COLF(mt, 2); BA_COLF(tstd, 3);
    BA_COLF(w_head, 4); BA_COLF(asg_lo, 5); BA_COLF(asg_hi, 6);
    BA_COLU(q_head, 7); BA_COLU(q_tail, 8); BA_COLU(dep_pops, 9);
    BA_COLU(arr_left, 10); BA_COLU(s_flags, 11); BA_COLU(s_qw, 12);
    BA_COLU(run_info, 13); BA_COLF(next_sched, 14);
    BA_COLU(q_fill, 15); BA_COLU(s_goff, 16); BA_COLU(run_ptr, 17); BA_COLU(run_tick, 18);
    BA_COLU(inmask, 19);                       // words 19 .. 22
    if (tracked) {
        BA_COLF(base, 23); BA_COLF(n_avail, 24); BA_COLF(n0, 25);
        BA_COLF(acc_ln0, 26); BA_COLF(acc_bc, 27);
        BA_COLU(next_dep, 28); BA_COLU(turn_n, 29); BA_COLU(av_head, 30);
        BA_COLU(av_tail, 31); BA_COLU(zeros_left, 32); BA_COLU(dl_head, 33);
        BA_COLU(dl_tail, 34); BA_COLU(s_depoff, 35); BA_COLU(s_depcnt, 36);
    } else {
        c.base.p = c.n_avail.p = c.n0.p = c.acc_ln0.p = c.acc_bc.p = nullptr;
        c.next_dep.p = c.turn_n.p = c.av_head.p = c.av_tail.p = c.zeros_left.p = nullptr;
        c.dl_head.p = c.dl_tail.p = c.s_depoff.p = c.s_depcnt.p = nullptr;
    }
#undef BA_COLF
#undef BA_COLU
}

// Fixed-size shared structures of the fast variant (calendar counters + two due buffers):
// they sit at compile-time offsets in front of the per-airport state, so every access to
// them is a shared load with an immediate address.
#define BA_DUE_STRIDE_S (BA_DUE_WORDS * BA_DUE_CAP_S + BA_DUE_CAP_S / 2)
#define BA_SHARED_FIXED_WORDS (BA_KCAL + 2 * BA_DUE_STRIDE_S + 2 * BA_DUE_CAP_S)

// Words of shared memory the size-dependent list structures of the fast variant need
// (runway rings; the staged departures are added by the caller).
BA_DEV uint32_t ba_shared_list_words(uint32_t q_total_s)
{
    return BA_Q_WORDS * q_total_s;
}

// Carves the queue pool / calendar counters / due buffer (shared or global) and the
// per-flight scratch.  `shared_fixed` = BA_SHARED_FIXED_WORDS of shared memory (fast
// variant), `shared_pool` = shared memory after the state records, `g` = start of this
// CTA's global scratch.
template <bool EXACT>
BA_DEV void ba_carve_lists(BaCtx &c, const BaDayView &day, int max_ticks, int max_dslice_plan,
                           uint32_t *shared_fixed, uint32_t *shared_pool, uint32_t *g)
{
    const uint32_t F = (uint32_t)day.F;
    c.nzs = (float *)g; g += 4 * F;           // first: 16-byte aligned
    c.dpair = g; g += 2 * F;                  // 8-byte aligned
    c.xdk = (float *)g; g += F; c.xak = (float *)g; g += F; c.atk = (float *)g; g += F;
    c.ct = (float *)g; g += F; c.xaf = (float *)g; g += F; c.atf = (float *)g; g += F;
    c.wd = (float *)g; g += F; c.wa = (float *)g; g += F;
    c.karr = g; g += F;
    c.cal = g; g += 3 * F;
    c.popkey = g; g += F; c.popseq = g; g += F;
    c.pool = c.cal;                           // cal (3F) + popkey + popseq = 5F words
    c.trl = c.tre = nullptr;
    if (day.any_track_t) { c.trl = (float *)g; g += F; c.tre = (float *)g; g += F; }
    c.qd = c.se = c.sw = nullptr;
    if (day.any_track_a) {
        c.qd = (float *)g; g += F; c.se = (float *)g; g += F; c.sw = (float *)g; g += F;
    }
    c.gx = (float *)g; g += day.q_total_x;
    c.gq = nullptr;
    if (!EXACT) { c.gq = g; g += BA_Q_WORDS * day.q_total_x; }
    uint32_t *s = EXACT ? g : shared_pool;
    c.dstage_cap = (uint32_t)day.max_dslice;
    c.dstage_stride = BA_DSTAGE_WORDS * (((uint32_t)max_dslice_plan + 1u) & ~1u);
    c.dstage_base = s; s += 2u * c.dstage_stride;
    const uint32_t qt = EXACT ? day.q_total_x : day.q_total_s;
    c.qe = s; s += BA_Q_WORDS * qt;
    if (EXACT) {
        c.due_cap = F;
        const uint32_t own_words = (c.due_cap + 7u) / 8u * 4u; // u16 per entry, 16-byte groups
        s += (4u - ((uint32_t)(s - g) & 3u)) & 3u;             // g is 16-byte aligned
        c.due_base = s; c.due_stride = BA_DUE_WORDS * c.due_cap + own_words;
        s += 2u * c.due_stride;
        c.cal_cap = (uint32_t)max_ticks + 2u;
        c.cal_cnt = s; s += c.cal_cap;
        g = s;
    } else {
        c.due_cap = (uint32_t)BA_DUE_CAP_S;
        c.cal_cap = (uint32_t)BA_KCAL;
        c.cal_cnt = shared_fixed;
        c.due_base = shared_fixed + BA_KCAL; c.due_stride = (uint32_t)BA_DUE_STRIDE_S;
        c.wbuf = c.due_base + 2 * BA_DUE_STRIDE_S;
    }
    c.tt = (float *)g; g += day.turn_total; c.te = (float *)g; g += day.turn_total;
    c.at = (float *)g; g += day.av_total; c.ae = (float *)g; g += day.av_total;
    c.dl = g;
    c.first_tick = day.first_tick;
    c.dcal_off = day.dcal_off; c.dcal_rec = day.dcal_rec; c.last_ready_tick = day.last_ready_tick;
    c.drun_off = day.drun_off; c.drun = day.drun;
    c.dep_cpos = day.dep_cpos;
}

BA_DEV BaNoise4 ba_noise(const BaCtx &c, int f)
{
    if (c.noise) {
#if defined(__CUDA_ARCH__)
        float4 v = __ldg(reinterpret_cast<const float4 *>(c.noise) + f);
        BaNoise4 z = {v.x, v.y, v.z, v.w};
        return z;
#else
        BaNoise4 z = {c.noise[4 * f], c.noise[4 * f + 1], c.noise[4 * f + 2], c.noise[4 * f + 3]};
        return z;
#endif
    }
    return ba_philox_noise(c.seed, c.particle, c.dayidx, (uint32_t)f);
}

// The epilogue needs the same four draws again.  Generated noise is kept in scratch by the
// prologue (16 bytes per flight) instead of running Philox + the four transforms twice.
BA_DEV void ba_noise_keep(BaCtx &c, int f, const BaNoise4 &z)
{
    if (c.noise) return;
#if defined(__CUDA_ARCH__)
    reinterpret_cast<float4 *>(c.nzs)[f] = make_float4(z.e_dep, z.eps_travel, z.e_arr, z.eps_turn);
#else
    c.nzs[4 * f] = z.e_dep; c.nzs[4 * f + 1] = z.eps_travel; c.nzs[4 * f + 2] = z.e_arr; c.nzs[4 * f + 3] = z.eps_turn;
#endif
}

BA_DEV BaNoise4 ba_noise_again(const BaCtx &c, int f)
{
    if (c.noise) return ba_noise(c, f);
#if defined(__CUDA_ARCH__)
    const float4 v = reinterpret_cast<const float4 *>(c.nzs)[f];
    BaNoise4 z = {v.x, v.y, v.z, v.w};
#else
    BaNoise4 z = {c.nzs[4 * f], c.nzs[4 * f + 1], c.nzs[4 * f + 2], c.nzs[4 * f + 3]};
#endif
    return z;
}

struct BaObsNoise { float u_cancel, eps_dep, eps_arr; };

BA_DEV BaObsNoise ba_noise2(const BaCtx &c, int f)
{
    BaObsNoise z;
    if (c.noise2) {
        z.u_cancel = c.noise2[4 * f]; z.eps_dep = c.noise2[4 * f + 1]; z.eps_arr = c.noise2[4 * f + 2];
        return z;
    }
    uint32_t r[4];
    ba_philox((uint32_t)f, c.dayidx, c.particle, 0x42413231u, (uint32_t)c.seed,
              (uint32_t)(c.seed >> 32), r);
    z.u_cancel = ba_u01(r[0]);
    const float rad = sqrtf(-2.0f * BA_LOG(ba_u01(r[1])));
    const float ang = 6.283185307179586f * ba_u01(r[2]);
    z.eps_dep = rad * cosf(ang); z.eps_arr = rad * sinf(ang);
    return z;
}

// RelaxedBernoulliStraightThrough(T = 0.5, probs = p) sample driven by the uniform u: hard 0/1
// (torch LogitRelaxedBernoulli.rsample + SigmoidTransform + Pyro's straight-through round;
// bayes_air/network.py:100-109 with obs=None)
BA_DEV float ba_relaxed_bernoulli_draw(float p, float u)
{
    const float q = fminf(fmaxf(p, BA_F32_EPS), 1.0f - BA_F32_EPS);
    const float uu = fminf(fmaxf(u, BA_F32_EPS), 1.0f - BA_F32_EPS);
    const float x = (BA_LOG(uu) - BA_LOG1P(-uu) + BA_LOG(q) - BA_LOG1P(-q)) / 0.5f;
    float y = 1.0f / (1.0f + BA_EXP(-x));
    y = fminf(fmaxf(y, BA_F32_TINY), 1.0f - BA_F32_EPS);
    y = fminf(fmaxf(y, BA_F32_EPS), 1.0f - BA_F32_EPS);
    return rintf(y);
}

BA_DEV BaFlight ba_flight(const BaCtx &c, int f)
{
#if defined(__CUDA_ARCH__)
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(c.flights) + f);
    BaFlight fl;
    fl.sched_dep = __uint_as_float(v.x); fl.act_dep = __uint_as_float(v.y);
    fl.act_arr = __uint_as_float(v.z); fl.od = v.w;
    return fl;
#else
    return c.flights[f];
#endif
}

// ---------------------------------------------------------------------------
// per-airport initialisation (model.py:139-155)
// ---------------------------------------------------------------------------
// Plans without aircraft bookkeeping (no cancellations, every airport "simple") run kernels
// built with TRK = false: the tracking flags read as clear at compile time, so all of that
// code (and its registers) drops out.
#define BA_AP_TRACK_ANY (BA_AP_TRACK_T | BA_AP_TRACK_A)
template <bool TRK>
BA_DEV uint32_t ba_ap_flags(const BaCtx &c, int a)
{
    const uint32_t f = c.s_flags[a];
    return TRK ? f : (f & ~(uint32_t)BA_AP_TRACK_ANY);
}

template <bool EXACT, bool TRK = true>
BA_DEV void ba_init_airport(BaCtx &c, int a, const float *lat_airport /* [C,N] */, int Ng)
{
    const BaAirport A = c.ap[a];
    const int g = A.global;
    const float mt = lat_airport[0 * Ng + g];
    const float m = lat_airport[1 * Ng + g];
    c.mt[a] = mt;
    c.m[a] = m;
    const float rate = BA_DIV(1.0f, m);          // airport.py:146
    const float tstd = BA_MUL(c.v_u, mt);        // model.py:143-145
    c.rate[a] = rate; c.tstd[a] = tstd;
    c.asg_lo[a] = NAN; c.asg_hi[a] = NAN; c.w_head[a] = 0.0f; c.psum[a] = 0.0;
    c.next_sched[a] = A.dep_cnt > 0 ? c.dep_sched[A.dep_off] : INFINITY;
    c.q_head[a] = 0; c.q_tail[a] = 0; c.q_fill[a] = 0; c.dep_pops[a] = 0;
    c.arr_left[a] = (uint32_t)A.arr_cnt;
    (&c.inmask[a])[0] = 0; (&c.inmask[a])[1] = 0; (&c.inmask[a])[2] = 0; (&c.inmask[a])[3] = 0;
    c.s_goff[a] = A.q_off_x;
    {
        const uint32_t rp = (uint32_t)c.drun_off[a];
        c.run_ptr[a] = rp;
        c.run_tick[a] = c.drun[2 * rp];
        c.run_info[a] = c.drun[2 * rp + 1];
    }
    if (TRK && c.n_avail.p) {
        float n0 = 1000.0f, base = 0.0f;             // model.py:116-121
        if (c.cancel_mode) {
            n0 = BA_EXP(lat_airport[2 * Ng + g]);    // model.py:91-99
            base = BA_EXP(lat_airport[3 * Ng + g]);  // model.py:103-111
        }
        c.base[a] = base;
        c.n_avail[a] = n0;
        c.n0[a] = n0;
        // while i < n0: append(0.0)  -> ceil(n0) zero-time aircraft (model.py:152-155)
        uint32_t nz = 0;
        if (n0 > 0.0f) {
            float cz = ceilf(n0);
            nz = cz > 4.0e9f ? 4000000000u : (uint32_t)cz;
        }
        c.zeros_left[a] = nz;
        c.acc_ln0[a] = 0.0f; c.acc_bc[a] = 0.0f;
        c.next_dep[a] = 0;
        c.turn_n[a] = 0; c.av_head[a] = 0; c.av_tail[a] = 0; c.dl_head[a] = 0; c.dl_tail[a] = 0;
        c.s_depoff[a] = (uint32_t)A.dep_off;
        c.s_depcnt[a] = (uint32_t)A.dep_cnt;
    }
    c.s_flags[a] = A.flags;
    c.s_qw[a] = ba_qw_pack(EXACT ? A.q_off_x : A.q_off_s, EXACT ? A.q_mask_x : A.q_mask_s);
}

// ---------------------------------------------------------------------------
// PROLOGUE: one flight (call for f = tid, tid + nthreads, ...)
// ---------------------------------------------------------------------------
// The per-particle half {x_dep, arrival time} of the flight's departure-queue entry, stored
// where the static half lies in the departure calendar (ba_stage_departure copies both).
BA_DEV void ba_store_dpair(BaCtx &c, int k, float xd, float at)
{
    const int cp = c.dep_cpos[k];
    if (cp < 0) return;
    BaPairF xy = {xd, at};
    BA_ST_PAIR(BaPairF, c.dpair + 2 * cp, xy);
}

template <bool PREDICT, bool TRK = true>
BA_DEV void ba_prologue_flight(BaCtx &c, BaAcc &acc, int f)
{
    const BaFlight fl = ba_flight(c, f);
    const uint32_t cobs = (fl.od >> 24) & 3u;
    const int o = (int)(fl.od & 0xFFFu), d = (int)((fl.od >> 12) & 0xFFFu);
    const BaNoise4 nz = ba_noise(c, f);
    ba_noise_keep(c, f, nz);
    const float T = c.lat_route[c.route[f]];
    // Exponential(1/m).rsample = e / rate  (airport.py:144-147)
    const float xd = c.value_mode ? nz.e_dep : BA_DIV(nz.e_dep, c.rate[o]);
    const float xa = c.value_mode ? nz.e_arr : BA_DIV(nz.e_arr, c.rate[d]);
    // Normal(T, T v_t).rsample = T + eps * (T v_t)  (network.py:158-163)
    const float sc = BA_MUL(T, c.v_t);
    const float tau = c.value_mode ? nz.eps_travel : BA_ADD(T, BA_MUL(nz.eps_travel, sc));
    float y_dep = fl.act_dep, y_arr = fl.act_arr;
    const int k = c.pos_dep[f];
    if (PREDICT) {
        // departure / arrival times may be sampled: the arrival time is only known when the
        // departure is served.  Stage the travel time and the turnaround duration instead.
        c.xdk[k] = xd; c.xak[k] = xa; c.atk[k] = tau; c.xaf[f] = xa;
        ba_store_dpair(c, k, xd, tau);
        c.sim[3 * f] = cobs == 1u ? 1.0f : 0.0f; c.sim[3 * f + 1] = NAN; c.sim[3 * f + 2] = NAN;
        if (TRK && c.trl) {
            const float mt = c.mt[d];
            c.trl[f] = c.value_mode ? nz.eps_turn : BA_ADD(mt, BA_MUL(nz.eps_turn, c.tstd[d]));
            c.tre[f] = nz.eps_turn;
        }
        return;
    }
    if (cobs == 3u || (cobs != 1u && (isnan(y_dep) || isnan(y_arr)))) {
        acc.status |= BA_S_UNOBSERVED;      // this entry point is the fully observed path
        y_dep = fl.sched_dep; y_arr = fl.sched_dep;
    }
    const float at = BA_ADD(y_dep, tau);    // network.py:166-168 with the observed departure
    // dep-order copies: only the aircraft-bookkeeping departures read them before the
    // epilogue reuses the arrays
    if (TRK) { c.xdk[k] = xd; c.xak[k] = xa; c.atk[k] = at; }
    c.xaf[f] = xa; c.atf[f] = at;
    ba_store_dpair(c, k, xd, at);
    c.popkey[f] = BA_KEY_INVALID;
    // arrival tick: the first tick whose clock reaches `at` (exact: same fp32 clock table
    // the scan uses; network.py:186 `arrival_time <= time`)
    uint32_t ka = 0xFFFFFFFFu;
    if (cobs != 1u) {
        const int kmax = (int)c.cal_cap - 1;
        float guess = ceilf(at / c.tick_time[1]);
        int kk = guess < 1.0f ? 1 : (guess > (float)kmax ? kmax : (int)guess);
        while (kk > 1 && c.tick_time[kk - 1] >= at) --kk;
        while (kk < kmax && c.tick_time[kk] < at) ++kk;
        // beyond the calendar: the shared-memory variant hands the item to the exact one
        if (c.tick_time[kk] < at) acc.status |= c.gq ? BA_S_OVERFLOW : BA_S_TICK_CAP;
        if (kk < c.first_tick) kk = c.first_tick;
        ka = (uint32_t)kk;
        BA_ATOMIC_INC(&c.cal_cnt[ka]);
    }
    c.karr[f] = ka;
    if (TRK && c.trl) {
        // Normal(m~, v_u m~).rsample (airport.py:217-221)
        const float mt = c.mt[d];
        const float u = c.value_mode ? nz.eps_turn : BA_ADD(mt, BA_MUL(nz.eps_turn, c.tstd[d]));
        c.trl[f] = BA_ADD(y_arr, u);
        c.tre[f] = nz.eps_turn;
    }
}

// ---------------------------------------------------------------------------
// runway-queue append (model.py:181-183 / network.py:186-191)
// ---------------------------------------------------------------------------
// The entry remembers how much service had been handed out at this airport when it
// joined (psum, rounded to fp32) and the index of the first later assignment (lo):
// its total wait is the sum of the draws lo..own index (airport.py:150-153).
BA_DEV void ba_q_push(BaCtx &c, BaAcc &acc, int a, uint32_t word, float qstart, float x, float y)
{
    const uint32_t head = c.q_head[a], tail = c.q_tail[a], mask = ba_qw_mask(c.s_qw[a]);
    uint32_t *e;
    if (c.q_fill[a] == tail && tail - head <= mask) {
        e = c.qe + BA_Q_WORDS * (ba_qw_off(c.s_qw[a]) + (tail & mask));       // room in the ring
        c.q_fill[a] = tail + 1;
    } else if (c.gq) {
        e = c.gq + BA_Q_WORDS * (c.s_goff[a] + tail);                // spill (absolute index)
    } else {
        acc.status |= BA_S_OVERFLOW; return;                          // cannot happen: exact ring
    }
    BaPairU wl = {word, head + (isnan(c.asg_hi[a]) ? 0u : 1u)};
    BaPairF sp = {qstart, (float)c.psum[a]};
    BaPairF xy = {x, y};
    BA_ST_PAIR(BaPairU, e, wl); BA_ST_PAIR(BaPairF, e + 2, sp); BA_ST_PAIR(BaPairF, e + 4, xy);
    c.gx[c.s_goff[a] + tail] = x;
    c.q_tail[a] = tail + 1;
}

// Moves spilled entries into the shared ring while there is room (owner lane only).
BA_DEV void ba_q_refill(BaCtx &c, int a, uint32_t head, uint32_t tail)
{
    uint32_t fill = c.q_fill[a];
    if (fill == tail) return;
    const uint32_t qw = c.s_qw[a], mask = ba_qw_mask(qw), qoff = ba_qw_off(qw), goff = c.s_goff[a];
    while (fill != tail && fill - head <= mask) {
        const uint32_t *src = c.gq + BA_Q_WORDS * (goff + fill);
        uint32_t *dst = c.qe + BA_Q_WORDS * (qoff + (fill & mask));
        BA_ST_PAIR(BaPairU, dst, BA_LD_PAIR(BaPairU, src));
        BA_ST_PAIR(BaPairU, dst + 2, BA_LD_PAIR(BaPairU, src + 2));
        BA_ST_PAIR(BaPairU, dst + 4, BA_LD_PAIR(BaPairU, src + 4));
        ++fill;
    }
    c.q_fill[a] = fill;
}

BA_DEV void ba_due_append(BaCtx &c, BaAcc &acc, int parity, uint32_t word, uint32_t key,
                          uint32_t owner, uint32_t seq, float qstart, float x, float y)
{
    const uint32_t slot = BA_ATOMIC_INC(&c.cta[parity]);
    if (slot >= c.due_cap) { acc.status |= BA_S_OVERFLOW; return; }
    uint32_t *buf = c.due_base + (uint32_t)parity * c.due_stride;
    uint32_t *de = buf + BA_DUE_WORDS * slot;
    BaPairU wk = {word, key};
    BaPairUF os = {(owner << 16) | (seq & 0xFFFFu), qstart};
    BaPairF xy = {x, y};
    BA_ST_PAIR(BaPairU, de, wk); BA_ST_PAIR(BaPairUF, de + 2, os); BA_ST_PAIR(BaPairF, de + 4, xy);
    ((uint16_t *)(buf + BA_DUE_WORDS * c.due_cap))[slot] = (uint16_t)owner;
}

// available_aircraft.pop(0) for one departing flight (network.py:118-126).
// Returns the ready time max(aircraft, sched); *weight / *src_eps describe the
// aircraft branch for the adjoint (0 none, 0.5 tie, 1 aircraft time strictly later).
BA_DEV float ba_take_aircraft(BaCtx &c, BaAcc &acc, int a, const BaAirport *A, float sched,
                              float t, float *weight, float *src_eps)
{
    *weight = 0.0f; *src_eps = 0.0f;
    c.n_avail[a] = c.n_avail[a] - 1.0f;
    float ac_time = 0.0f;
    bool has_src = false;
    if (c.s_flags[a] & BA_AP_TRACK_A) {
        if (c.zeros_left[a] > 0) {
            c.zeros_left[a] -= 1;
        } else {
            uint32_t h = c.av_head[a];
            if (h == c.av_tail[a]) { acc.status |= BA_S_INTERNAL; return sched; }
            uint32_t i = A->av_off + (h & A->av_mask);
            float tm = c.at[i];
            if (isnan(tm)) {
                // run of reserve aircraft: the reference appended the clock tensor
                // itself (model.py:174), so its value is the clock at pop time
                ac_time = t;
                float left = c.ae[i] - 1.0f;
                if (left <= 0.0f) c.av_head[a] = h + 1; else c.ae[i] = left;
            } else {
                ac_time = tm; *src_eps = c.ae[i]; has_src = !c.value_mode;
                c.av_head[a] = h + 1;
            }
        }
    }
    float ready = fmaxf(ac_time, sched);          // torch.maximum
    if (has_src) *weight = ac_time > sched ? 1.0f : (ac_time == sched ? 0.5f : 0.0f);
    return ready;
}

// one ready departure of a tracked airport: take an aircraft, enqueue, record the branch
BA_DEV void ba_depart_tracked(BaCtx &c, BaAcc &acc, int a, const BaAirport *A, uint32_t k,
                              float t)
{
    const uint32_t w = c.dep_word[k];
    const int f = (int)(w & BA_FID_MASK);
    const float sched = c.dep_sched[k];
    float weight, src_eps;
    const float ready = ba_take_aircraft(c, acc, a, A, sched, t, &weight, &src_eps);
    if (c.qd) { c.qd[f] = ready; c.se[f] = src_eps; c.sw[f] = weight; }
    ba_q_push(c, acc, a, BA_Q_DEP | (w & BA_DW_ENTRY_MASK), ready, c.xdk[k], c.atk[k]);
}

// ---------------------------------------------------------------------------
// SCAN, phase A
// ---------------------------------------------------------------------------
BA_DEV void ba_late_append(BaCtx &c, BaAcc &acc, int tick, uint32_t n_slice, uint32_t word,
                           uint32_t key, uint32_t owner, uint32_t seq, float qstart, float x);

template <bool EXACT, bool PREDICT, bool TRK = true>
BA_DEV void ba_phase_a(BaCtx &c, BaAcc &acc, int a, float t, int tick, uint32_t n_slice_cur = 0)
{
    const uint32_t flags = ba_ap_flags<TRK>(c, a);
    const BaAirport *A = &c.ap[a];      // only dereferenced on the tracked paths

    // 1. update_available_aircraft (airport.py:66-84)
    if (flags & BA_AP_TRACK_T) {
        uint32_t n = c.turn_n[a];
        if (n) {
            const uint32_t off = A->turn_off;
            uint32_t w = 0;
            float navail = c.n_avail[a];
            for (uint32_t i = 0; i < n; ++i) {
                float tm = c.tt[off + i], ep = c.te[off + i];
                if (tm <= t) {
                    navail = navail + 1.0f;
                    if (flags & BA_AP_TRACK_A) {
                        uint32_t tl = c.av_tail[a];
                        if (tl - c.av_head[a] > A->av_mask) acc.status |= BA_S_INTERNAL;
                        else {
                            uint32_t j = A->av_off + (tl & A->av_mask);
                            c.at[j] = tm; c.ae[j] = ep; c.av_tail[a] = tl + 1;
                        }
                    }
                } else {
                    if (w != i) { c.tt[off + w] = tm; c.te[off + w] = ep; }
                    ++w;
                }
            }
            c.turn_n[a] = w;
            c.n_avail[a] = navail;
        }
        // 2. reserve injection (model.py:170-174)
        if (t >= c.max_t) {
            c.n_avail[a] = c.n_avail[a] + 1.0f;
            if (flags & BA_AP_TRACK_A) {
                uint32_t tl = c.av_tail[a], hd = c.av_head[a];
                bool merged = false;
                if (tl != hd) {
                    uint32_t j = A->av_off + ((tl - 1) & A->av_mask);
                    if (isnan(c.at[j])) { c.ae[j] = c.ae[j] + 1.0f; merged = true; }
                }
                if (!merged) {
                    if (tl - hd > A->av_mask) acc.status |= BA_S_INTERNAL;
                    else {
                        uint32_t j = A->av_off + (tl & A->av_mask);
                        c.at[j] = NAN; c.ae[j] = 1.0f; c.av_tail[a] = tl + 1;
                    }
                }
            }
        }
    }

    // 3. pop_ready_to_depart_flights restricted to this origin (network.py:57-134).
    //    Airports without aircraft bookkeeping never cancel or delay: their departures
    //    arrive through the static departure calendar (phase B) instead.
    if (flags & BA_AP_TRACK_T) {
        const bool due = c.next_sched[a] <= t;
        const uint32_t dep_cnt = c.s_depcnt[a], dep_off = c.s_depoff[a];
        const uint32_t nd = c.next_dep[a];
        uint32_t n_new = 0;
        if (due) {
            // the cached head of this origin's pending list is due: count the due prefix
            n_new = 1;
            float nxt = INFINITY;
            while (nd + n_new < dep_cnt) {
                nxt = c.dep_sched[dep_off + nd + n_new];
                if (!(nxt <= t)) break;
                ++n_new;
                nxt = INFINITY;
            }
            c.next_sched[a] = nxt;
        }
        const uint32_t k0 = dep_off + nd;
        const uint32_t n_old = c.dl_tail[a] - c.dl_head[a];
        const uint32_t n_ready = n_new + n_old;
        if (n_ready) {
            const float n_start = c.n_avail[a];
            const float base = c.base[a];
            float lp0 = 0.0f, lp1 = 0.0f, dl0 = 0.0f, dl1 = 0.0f, sg = 1.0f, dp_dn = 0.0f;
            if (n_new) {
                // network.py:80-84, fp32 in the reference's operation order
                const float ratio = BA_DIV(n_start, (float)n_ready);
                const float z = BA_MUL(10.0f, BA_ADD(ratio, -0.25f));
                sg = 1.0f / (1.0f + BA_EXP(-z));
                const float p = 1.0f - (1.0f - base) * sg;
                lp0 = ba_relaxed_bernoulli_lp(p, 0.0f, &dl0);
                lp1 = ba_relaxed_bernoulli_lp(p, 1.0f, &dl1);
                // d p / d base = sg ; d p / d n = -(1-base) sg (1-sg) * 10 / n_ready
                dp_dn = -(1.0f - base) * sg * (1.0f - sg) * 10.0f / (float)n_ready;
            }
            float g_p = 0.0f;          // sum over newly scored flights of d lp / d p
            uint32_t first_delayed = n_new;
            for (uint32_t i = 0; i < n_new; ++i) {
                const uint32_t k = k0 + i;
                uint32_t cobs = (c.dep_word[k] >> BA_DW_CANCEL_SHIFT) & 3u;
                if (cobs == 3u) {
                    if (!PREDICT) { acc.status |= BA_S_UNOBSERVED; continue; }
                    // obs=None: the site is sampled (network.py:100-109)
                    const uint32_t w = c.dep_word[k];
                    const int f = (int)(w & BA_FID_MASK);
                    const float pp = 1.0f - (1.0f - base) * sg;
                    cobs = ba_relaxed_bernoulli_draw(pp, ba_noise2(c, f).u_cancel) != 0.0f ? 1u : 0u;
                    c.sim[3 * f] = (float)cobs;
                    if (cobs)   // it will never land: one arrival less to wait for
                        BA_ATOMIC_DEC(&c.arr_left[(w >> BA_Q_DST_SHIFT) & 0xFFFu]);
                }
                acc.lp += cobs ? lp1 : lp0;
                g_p += cobs ? dl1 : dl0;
                if (cobs) continue;                       // cancelled -> completed
                if (c.n_avail[a] <= 0.0f) {               // delayed (network.py:112-114)
                    if (first_delayed == n_new) first_delayed = i;
                    continue;
                }
                ba_depart_tracked(c, acc, a, A, k, t);
            }
            if (c.cancel_mode && c.want_grad && n_new) {
                c.acc_bc[a] += g_p * sg * base;           // b = exp(c)
                // n = exp(l) + integers: d n / d l = exp(l); multiplied in the epilogue
                c.acc_ln0[a] += g_p * dp_dn;
            }
            // newly delayed flights go IN FRONT of the previously delayed ones
            // (they precede them in the re-built pending list, network.py:60-68,112-114)
            if (first_delayed < n_new) {
                uint32_t hd = c.dl_head[a];
                for (uint32_t i = n_new; i-- > first_delayed;) {
                    const uint32_t k = k0 + i;
                    uint32_t cobs = (c.dep_word[k] >> BA_DW_CANCEL_SHIFT) & 3u;
                    if (PREDICT && cobs == 3u) cobs = c.sim[3 * (c.dep_word[k] & BA_FID_MASK)] != 0.0f;
                    if (cobs != 0u) continue;
                    hd -= 1;
                    c.dl[A->dl_off + (hd & A->dl_mask)] = k;      // dep-order position
                }
                c.dl_head[a] = hd;
            } else {
                // previously delayed flights, oldest first, while aircraft remain
                uint32_t hd = c.dl_head[a];
                const uint32_t tl = c.dl_tail[a];
                while (hd != tl && c.n_avail[a] > 0.0f) {
                    ba_depart_tracked(c, acc, a, A, c.dl[A->dl_off + (hd & A->dl_mask)], t);
                    ++hd;
                }
                c.dl_head[a] = hd;
            }
            c.next_dep[a] = nd + n_new;
        }
    }

    // 5. update_runway_queue (airport.py:101-128): shared memory only.
    //
    // The reference adds every service draw x to the wait of EVERY queued entry
    // (airport.py:150-153), O(queue length) per assignment.  Here an entry's wait is
    // psum(now) - psum(at joining), O(1).  That difference R (double) is not the fp32
    // sequential sum W the reference forms, so the decision `queue_start + W <= t` is
    // taken on an interval [asg_lo, asg_hi] that provably contains the reference's fp32
    // value:  |W - R| <= (n-1) u R  (sequential fp32 sum of n non-negative terms),
    // |fl32(psum) - psum| <= u psum,  |fl32(qs + W) - (qs + W)| <= u |qs + W|,
    // u = 2^-24.  Only when the clock falls INSIDE the interval (probability ~1e-5 per
    // check) is W replayed with the reference's sequential fp32 adds.
    {
        uint32_t head = c.q_head[a];
        const uint32_t tail = c.q_tail[a];
        if (head == tail) return;
        const uint32_t qw = c.s_qw[a], qoff = ba_qw_off(qw), qmask = ba_qw_mask(qw);
        float lo = c.asg_lo[a], hi = c.asg_hi[a], wait = c.w_head[a];
        double psum = c.psum[a];
        if (c.gq) ba_q_refill(c, a, head, tail);
        while (head != tail) {
            if (c.gq && head == c.q_fill[a]) { c.q_head[a] = head; ba_q_refill(c, a, head, tail); }
            const uint32_t *e = c.qe + BA_Q_WORDS * (qoff + (head & qmask));
            const BaPairU wl = BA_LD_PAIR(BaPairU, e);          // word, lo
            if (isnan(hi)) {
                // _assign_service_time (airport.py:130-158)
                const BaPairF sp = BA_LD_PAIR(BaPairF, e + 2);  // queue start, psum at joining
                psum += (double)BA_LD_PAIR(BaPairF, e + 4).a;
                const double R = psum - (double)sp.b;
                const double A_ = (double)sp.a + R;
                wait = (float)R;
                // margin in fp32: 4u (n R + |pj| + |A|) covers the three bounds above plus the
                // roundings of this very computation
                const float Af = (float)A_;
                const float Bf = 2.384185791015625e-07f *
                    ((float)(head - wl.b + 1u) * wait + fabsf(sp.b) + fabsf(Af)) + 1e-30f;
                lo = Af - Bf; hi = Af + Bf;
                if (PREDICT || c.replay_always) lo = -INFINITY, hi = INFINITY;
            }
            if (!(hi <= t)) {
                if (lo > t || isnan(lo)) break;             // certainly still being served
                // ambiguous: replay the reference's arithmetic exactly from the history of
                // service draws (global, absolute index: never overwritten)
                float w = 0.0f;
                const float *hx = c.gx + c.s_goff[a];
                for (uint32_t j = wl.b; j != head + 1u; ++j) w = BA_ADD(w, hx[j]);
                wait = w;
                lo = hi = BA_ADD(BA_LD_PAIR(BaPairF, e + 2).a, w);
                if (!(hi <= t)) break;
            }
            // served (airport.py:111-126)
            const uint32_t word = wl.a;
            const int f = (int)(word & BA_FID_MASK);
            if (PREDICT) {
                // outcomes: observed values where present, else Normal(mu, sigma_r).rsample
                // (airport.py:175-182,197-204); hi == lo == the exact fp32 queue_start + wait
                const BaFlight fl = ba_flight(c, f);
                const BaObsNoise on = ba_noise2(c, f);
                if (word & BA_Q_DEP) {
                    float y = fl.act_dep;
                    if (isnan(y)) y = BA_ADD(hi, BA_MUL(on.eps_dep, c.sigma));
                    c.sim[3 * f + 1] = y;
                    const uint32_t key = ((uint32_t)tick << BA_KEY_TICK_SHIFT) | (uint32_t)a;
                    const uint32_t seq = c.dep_pops[a];
                    c.dep_pops[a] = seq + 1;
                    const float av = BA_ADD(y, BA_LD_PAIR(BaPairF, e + 4).b);   // + travel time
                    const uint32_t dst = (word >> BA_Q_DST_SHIFT) & 0xFFFu;
                    if (av <= t) {
                        ba_due_append(c, acc, tick & 1, word & BA_FID_MASK, key, dst, seq, av,
                                      c.xaf[f], 0.0f);
                    } else {
                        const uint32_t slot = BA_ATOMIC_INC(&c.cta[2]);
                        uint32_t *pe = c.pool + 5u * slot;
                        pe[0] = word & BA_DW_ENTRY_MASK; pe[1] = key; pe[2] = seq;
                        pe[3] = __float_as_uint_portable(av);
                        pe[4] = __float_as_uint_portable(c.xaf[f]);
                    }
                    if (c.pop_tick) c.pop_tick[2 * f] = tick;
                } else {
                    float y = fl.act_arr;
                    if (isnan(y)) y = BA_ADD(hi, BA_MUL(on.eps_arr, c.sigma));
                    c.sim[3 * f + 2] = y;
                    BA_ATOMIC_DEC(&c.arr_left[a]);     // origin lanes may decrement it too (sampled cancellation)
                    if (flags & BA_AP_TRACK_T) {
                        const uint32_t n = c.turn_n[a];
                        c.tt[A->turn_off + n] = BA_ADD(y, c.trl[f]);
                        c.te[A->turn_off + n] = c.tre[f];
                        c.turn_n[a] = n + 1;
                    }
                    if (c.pop_tick) c.pop_tick[2 * f + 1] = tick;
                }
            } else if (word & BA_Q_DEP) {
                c.wd[f] = wait;
                // add_in_transit_flights (network.py:136-168).  The in-transit list is
                // implicit: its order is the key (tick, origin, per-origin sequence).
                const uint32_t key = ((uint32_t)tick << BA_KEY_TICK_SHIFT) | (uint32_t)a;
                const uint32_t seq = c.dep_pops[a];
                c.dep_pops[a] = seq + 1;
                const float av = BA_LD_PAIR(BaPairF, e + 4).b;
                if (av <= t) {
                    // already past its arrival time: joins the destination queue at the
                    // end of this very tick (update_in_transit_flights, network.py:186)
                    if (!EXACT)     // single-barrier scan: slots from the top of this tick's buffer
                        ba_late_append(c, acc, tick, n_slice_cur, word & BA_FID_MASK, key,
                                       (word >> BA_Q_DST_SHIFT) & 0xFFFu, seq, av, c.xaf[f]);
                    else
                        ba_due_append(c, acc, tick & 1, word & BA_FID_MASK, key,
                                      (word >> BA_Q_DST_SHIFT) & 0xFFFu, seq, av, c.xaf[f], 0.0f);
                } else {
                    c.popkey[f] = key; c.popseq[f] = seq;
                }
                if (c.pop_tick) c.pop_tick[2 * f] = tick;
            } else {
                c.wa[f] = wait;
                c.arr_left[a] -= 1;
                // _assign_turnaround_time (airport.py:210-221)
                if (flags & BA_AP_TRACK_T) {
                    const uint32_t n = c.turn_n[a];
                    c.tt[A->turn_off + n] = c.trl[f];
                    c.te[A->turn_off + n] = c.tre[f];
                    c.turn_n[a] = n + 1;
                }
                if (c.pop_tick) c.pop_tick[2 * f + 1] = tick;
            }
            ++head;
            lo = hi = NAN;
        }
        c.q_head[a] = head;
        c.asg_lo[a] = lo; c.asg_hi[a] = hi; c.w_head[a] = wait; c.psum[a] = psum;
    }
}

// ---------------------------------------------------------------------------
// PROLOGUE, arrival-calendar construction (after every flight's arrival tick is counted):
//   ba_calendar_scan     exclusive prefix over the bucket counts, by one warp
//                        (`lane` = 0..31; the host emulation passes lane = -1: serial)
//   ba_calendar_scatter  one flight: write its record into its bucket
// Afterwards cal_cnt[k] is the END offset of bucket k (= start of bucket k+1).
// ---------------------------------------------------------------------------
BA_DEV void ba_calendar_scan(BaCtx &c, int lane)
{
#if defined(__CUDA_ARCH__)
    const uint32_t per = (c.cal_cap + 31u) / 32u;
    const uint32_t b = (uint32_t)lane * per, e = min(b + per, c.cal_cap);
    uint32_t sum = 0;
    for (uint32_t k = b; k < e; ++k) sum += c.cal_cnt[k];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    uint32_t run = incl - sum;
    for (uint32_t k = b; k < e; ++k) { const uint32_t n = c.cal_cnt[k]; c.cal_cnt[k] = run; run += n; }
#else
    (void)lane;
    uint32_t run = 0;
    for (uint32_t k = 0; k < c.cal_cap; ++k) { const uint32_t n = c.cal_cnt[k]; c.cal_cnt[k] = run; run += n; }
#endif
}

BA_DEV void ba_calendar_scatter(BaCtx &c, int f)
{
    const uint32_t ka = c.karr[f];
    if (ka == 0xFFFFFFFFu) return;
    const uint32_t pos = BA_ATOMIC_INC(&c.cal_cnt[ka]);
    // one word per record: dest << 18 | flight; the arrival time and the service draw are
    // fetched by flight (atf, xaf) when the slice is assembled
    c.cal[pos] = (uint32_t)f | (((ba_flight(c, f).od >> 12) & 0xFFFu) << BA_Q_DST_SHIFT);
}

// ---------------------------------------------------------------------------
// SCAN, phase B (flight-parallel).
//   ba_stage_arrival      (one tick ahead) the static part of one record of an
//       arrival-calendar slice goes to its slot of that tick's due buffer
//   ba_phase_b_arrival    one slot of this tick's slice: a flight whose departure has been
//       served gets its place in the in-transit order; one whose departure is still queued
//       leaves its slot empty -- phase A hands it over itself when it is served
//   ba_stage_departure    (two ticks ahead) one record of a departure-calendar slice
//   ba_run_advance        per airport: the next run of the departure calendar
// ---------------------------------------------------------------------------
// Copies of static records into the staging areas.  Fast variant: asynchronous global ->
// shared copies (cp.async), issued one or two ticks before the data is read, so that no
// thread ever waits for them; ba_copy_wait() before the next barrier makes them visible.
// Exact variant / host: plain copies at the same place.
#if !defined(__CUDACC__)
// Host emulation of the asynchronous copies: with a hook installed they are queued and land
// when the emulated barrier is reached (tests/hostemu), otherwise they land at once -- the
// two extremes of what cp.async may do.
struct BaEmuCopy { uint32_t *dst; const uint32_t *src; int words; };
static thread_local std::vector<BaEmuCopy> *ba_emu_defer = nullptr;
#endif

template <bool EXACT>
BA_DEV void ba_copy8(uint32_t *dst, const uint32_t *src)
{
#if defined(__CUDA_ARCH__)
    if (!EXACT) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::
                     "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        return;
    }
#endif
#if !defined(__CUDACC__)
    if (!EXACT && ba_emu_defer) { ba_emu_defer->push_back({dst, src, 2}); return; }
#endif
    BA_ST_PAIR(BaPairU, dst, BA_LD_PAIR(BaPairU, src));
}

template <bool EXACT>
BA_DEV void ba_copy4(uint32_t *dst, const uint32_t *src)
{
#if defined(__CUDA_ARCH__)
    if (!EXACT) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::
                     "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
        return;
    }
#endif
#if !defined(__CUDACC__)
    if (!EXACT && ba_emu_defer) { ba_emu_defer->push_back({dst, src, 1}); return; }
#endif
    *dst = *src;
}

BA_DEV void ba_copy_wait()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;\n" ::: "memory");
#endif
}

// Stage record i of an arrival-calendar slice into its slot of a due buffer: {queue word,
// arrival time (= queue start), service draw}.  The slot is completed (ordering key,
// owner) by ba_phase_b_arrival at the slice's tick.
template <bool EXACT>
BA_DEV void ba_stage_arrival(BaCtx &c, uint32_t i, uint32_t slice_lo, int parity)
{
    const uint32_t slot = i - slice_lo;
    if (slot >= c.due_cap) return;                  // flagged by the caller
    uint32_t *de = c.due_base + (uint32_t)parity * c.due_stride + BA_DUE_WORDS * slot;
    const uint32_t word = c.cal[i];
    const uint32_t f = word & BA_FID_MASK;
    de[0] = word;
    de[3] = __float_as_uint_portable(c.atf[f]);
    de[4] = __float_as_uint_portable(c.xaf[f]);
}

BA_DEV void ba_phase_b_arrival(BaCtx &c, uint32_t slot, int parity)
{
    uint32_t *buf = c.due_base + (uint32_t)parity * c.due_stride;
    uint32_t *de = buf + BA_DUE_WORDS * slot;
    uint16_t *own = (uint16_t *)(buf + BA_DUE_WORDS * c.due_cap);
    const uint32_t word = de[0];
    const int f = (int)(word & BA_FID_MASK);
    const uint32_t key = c.popkey[f];
    if (key == BA_KEY_INVALID) {
        // departure still queued (phase A hands the flight over itself when it is served):
        // the slot stays empty
        own[slot] = (uint16_t)0xFFFFu; de[2] = 0xFFFF0000u;
        return;
    }
    c.popkey[f] = BA_KEY_INVALID;          // consumed
    const uint32_t owner = (word >> BA_Q_DST_SHIFT) & 0xFFFu;
    de[1] = key; de[2] = (owner << 16) | (c.popseq[f] & 0xFFFFu);
    own[slot] = (uint16_t)owner;
}

// Stage record i of a departure-calendar slice: the flight becomes ready at the slice's
// tick (network.py:60-68) at an airport that can neither cancel nor delay it.  Its queue
// entry {queue word, queue start} {x, arrival time} lands at the record's position in the
// slice; the origin lane knows (static runs) where its records are and appends them behind
// the arrivals of the tick before (model.py:181-183).
template <bool EXACT>
BA_DEV void ba_stage_departure(BaCtx &c, BaAcc &acc, uint32_t i, uint32_t slice_lo, int parity)
{
    if (i - slice_lo >= c.dstage_cap) { acc.status |= BA_S_INTERNAL; return; }
    uint32_t *d = c.dstage_base + (uint32_t)parity * c.dstage_stride + BA_DSTAGE_WORDS * (i - slice_lo);
    ba_copy8<EXACT>(d, c.dcal_rec + 2 * i);
    ba_copy8<EXACT>(d + 2, c.dpair + 2 * i);
}

// The run of the departure calendar this airport consumed at `tick` is done: fetch its next
// one (interval 2: off the lanes' critical path).
BA_DEV void ba_run_advance(BaCtx &c, int a, int tick)
{
    if (c.run_tick[a] != (uint32_t)tick) return;
    const uint32_t rp = c.run_ptr[a] + 1u;
    const uint32_t *r = c.drun + 2 * rp;
    c.run_ptr[a] = rp;
    c.run_tick[a] = r[0];
    c.run_info[a] = r[1];
}

// Posterior-predictive mode: arrival times are only known once a departure is served, so
// in-transit flights sit in an append-only pool that all threads scan each tick
// (update_in_transit_flights, network.py:178-198); consumed entries are tombstoned.
BA_DEV void ba_phase_b_pool(BaCtx &c, BaAcc &acc, uint32_t i, float t, int parity)
{
    uint32_t *pe = c.pool + 5u * i;
    const float av = __uint_as_float_portable(pe[3]);
    if (!(av <= t)) return;                        // not yet arrived, or tombstone (NaN)
    const uint32_t word = pe[0];
    ba_due_append(c, acc, parity, word & BA_FID_MASK, pe[1], (word >> BA_Q_DST_SHIFT) & 0xFFFu,
                  pe[2], av, __uint_as_float_portable(pe[4]), 0.0f);
    pe[3] = __float_as_uint_portable(NAN);
}

// advances the first-live pointer over leading tombstones (one thread; up to `budget`)
BA_DEV void ba_pool_advance(BaCtx &c, uint32_t budget)
{
    uint32_t lo = c.cta[4];
    const uint32_t n = c.cta[2];
    while (lo < n && budget-- && isnan(__uint_as_float_portable(c.pool[5u * lo + 3]))) ++lo;
    c.cta[4] = lo;
}

// ---------------------------------------------------------------------------
// SINGLE-BARRIER SCAN (shared-list variant, observed flights).  Which flights of the arrival
// slice of tick j have left their origin is settled when tick j - 1 ends (a departure served
// in tick j itself at or after its arrival time takes the late path), so the slice's due
// buffer can be assembled DURING tick j, next to the lanes' work, by asynchronous copies that
// nobody waits for until the tick's only barrier:
//   ba_stage_word        (tick j - 1) queue word of record i of slice j -> word buffer j & 1
//   ba_complete_arrival  (tick j) slot s of slice j: {word} by a plain store, {ordering key,
//       sequence, arrival time, service draw} by cp.async straight from where phase A and the
//       prologue left them; owner list entry.  A flight that has not departed keeps the
//       invalid key it was initialised with: the owner sees it sort last and stops there.
//   ba_late_append       (tick j, phase A) entries served after their arrival time take slots
//       from the top of the same buffer downwards (counter [8 + j % 3])
//   ba_run_advance_async the airport's next calendar run, copied into its state record
// ---------------------------------------------------------------------------
BA_DEV void ba_stage_word(BaCtx &c, uint32_t i, uint32_t slice_lo, int parity)
{
    const uint32_t s = i - slice_lo;
    if (s >= (uint32_t)BA_DUE_CAP_S) return;
    ba_copy4<false>(c.wbuf + (uint32_t)parity * BA_DUE_CAP_S + s, c.cal + i);
}

BA_DEV void ba_complete_arrival(BaCtx &c, uint32_t s, uint32_t slice_lo, int parity)
{
    const uint32_t word = c.wbuf[(uint32_t)parity * BA_DUE_CAP_S + s];
    const uint32_t f = word & BA_FID_MASK;
    uint32_t *buf = c.due_base + (uint32_t)parity * c.due_stride;
    uint32_t *de = buf + BA_DUE_WORDS * s;
    de[0] = word;
    ba_copy4<false>(de + 1, c.popkey + f);
    ba_copy4<false>(de + 2, c.popseq + f);
    ba_copy4<false>(de + 3, (const uint32_t *)c.atf + f);
    ba_copy4<false>(de + 4, (const uint32_t *)c.xaf + f);
    const uint32_t owner = (word >> BA_Q_DST_SHIFT) & 0xFFFu;
    ((uint16_t *)(buf + BA_DUE_WORDS * c.due_cap))[s] = (uint16_t)owner;
    if (s < 64u) BA_ATOMIC_OR(&c.inmask[owner] + 2 * parity + (s >> 5), 1u << (s & 31u));
}

BA_DEV void ba_late_append(BaCtx &c, BaAcc &acc, int tick, uint32_t n_slice, uint32_t word,
                           uint32_t key, uint32_t owner, uint32_t seq, float qstart, float x)
{
    const uint32_t l = BA_ATOMIC_INC(&c.cta[8 + tick % 3]);
    if (l + n_slice >= c.due_cap) { acc.status |= BA_S_OVERFLOW; return; }
    const uint32_t slot = c.due_cap - 1u - l;
    uint32_t *buf = c.due_base + (uint32_t)(tick & 1) * c.due_stride;
    uint32_t *de = buf + BA_DUE_WORDS * slot;
    BaPairU wk = {word, key};
    BaPairUF os = {seq & 0xFFFFu, qstart};
    BA_ST_PAIR(BaPairU, de, wk); BA_ST_PAIR(BaPairUF, de + 2, os);
    de[4] = __float_as_uint_portable(x);
    ((uint16_t *)(buf + BA_DUE_WORDS * c.due_cap))[slot] = (uint16_t)owner;
}

BA_DEV void ba_run_advance_async(BaCtx &c, int a)
{
    const uint32_t rp = c.run_ptr[a] + 1u;
    c.run_ptr[a] = rp;
    ba_copy4<false>(&c.run_tick[a], c.drun + 2 * rp);
    ba_copy4<false>(&c.run_info[a], c.drun + 2 * rp + 1);
}

// ---------------------------------------------------------------------------
// SCAN, phase C (owner pull): airport `a` appends the due entries addressed to it to its
// runway queue, in key order = the reference's in-transit list order restricted to this
// destination (model.py:186-196, network.py:186-191), then the newly ready departures in
// pending order.  Runs at the start of the next tick's interval, before phase A.
// ---------------------------------------------------------------------------
BA_DEV uint64_t ba_due_order(const uint32_t *de)
{
    return ((uint64_t)de[1] << 16) | (uint64_t)(de[2] & 0xFFFFu);
}

// Ring counters of one airport held in registers while a batch of entries is appended.
// Every entry appended between two service loops shares lo and the prefix sum at joining.
struct BaPush {
    uint32_t head, tail, fill, qoff, qmask, goff, lo;
    float pj;
};

BA_DEV void ba_push_begin(const BaCtx &c, int a, BaPush &P)
{
    P.head = c.q_head[a]; P.tail = c.q_tail[a]; P.fill = c.q_fill[a];
    { const uint32_t qw = c.s_qw[a]; P.qoff = ba_qw_off(qw); P.qmask = ba_qw_mask(qw); }
    P.goff = c.s_goff[a];
    P.lo = P.head + (isnan(c.asg_hi[a]) ? 0u : 1u);
    P.pj = (float)c.psum[a];
}

BA_DEV void ba_push_one(BaCtx &c, BaAcc &acc, BaPush &P, uint32_t word, float qstart, float x, float y)
{
    uint32_t *e;
    if (P.fill == P.tail && P.tail - P.head <= P.qmask) {
        e = c.qe + BA_Q_WORDS * (P.qoff + (P.tail & P.qmask));
        P.fill = P.tail + 1;
    } else if (c.gq) {
        e = c.gq + BA_Q_WORDS * (P.goff + P.tail);
    } else {
        acc.status |= BA_S_OVERFLOW; return;
    }
    BaPairU wl = {word, P.lo};
    BaPairF sp = {qstart, P.pj};
    BaPairF xy = {x, y};
    BA_ST_PAIR(BaPairU, e, wl); BA_ST_PAIR(BaPairF, e + 2, sp); BA_ST_PAIR(BaPairF, e + 4, xy);
    c.gx[P.goff + P.tail] = x;
    P.tail += 1;
}

BA_DEV void ba_push_end(BaCtx &c, int a, const BaPush &P)
{
    c.q_tail[a] = P.tail; c.q_fill[a] = P.fill;
}

static_assert(BA_DUE_CAP_S <= 128, "the owner scan of the shared due buffer uses two 64-bit masks");
template <bool EXACT = true>
BA_DEV void ba_phase_c_pull(BaCtx &c, BaAcc &acc, int a, int parity, uint32_t n_due, int tick)
{
    const bool has_run = c.run_tick[a] == (uint32_t)tick;
    uint64_t m0 = 0, m1 = 0;
    const uint32_t *due = c.due_base + (uint32_t)parity * c.due_stride;
    if (n_due && n_due <= 128u) {
        // owners are packed u16, eight per 16-byte load (broadcast: every lane of the warp
        // reads the same words)
        const uint32_t *ow = due + BA_DUE_WORDS * c.due_cap;
        const uint32_t aa = (uint32_t)a | ((uint32_t)a << 16);
        const uint32_t groups = (n_due + 7u) >> 3;
        for (uint32_t g = 0; g < groups; ++g) {
            uint32_t w0 = ow[4 * g], w1 = ow[4 * g + 1], w2 = ow[4 * g + 2], w3 = ow[4 * g + 3];
            uint32_t bits = 0;
            uint32_t x;
            x = w0 ^ aa; bits |= ((x & 0xFFFFu) == 0u) | (((x >> 16) == 0u) << 1);
            x = w1 ^ aa; bits |= (((x & 0xFFFFu) == 0u) << 2) | (((x >> 16) == 0u) << 3);
            x = w2 ^ aa; bits |= (((x & 0xFFFFu) == 0u) << 4) | (((x >> 16) == 0u) << 5);
            x = w3 ^ aa; bits |= (((x & 0xFFFFu) == 0u) << 6) | (((x >> 16) == 0u) << 7);
            const uint32_t left = n_due - 8u * g;
            if (left < 8u) bits &= (1u << left) - 1u;
            if (g < 8u) m0 |= (uint64_t)bits << (8u * g); else m1 |= (uint64_t)bits << (8u * (g - 8u));
        }
    }
    const bool general = EXACT && n_due > 128u;     // the shared due buffer holds 128 entries
    if (!(m0 | m1) && !has_run && !general) return;
    BaPush P;
    ba_push_begin(c, a, P);
    // ---- arrivals addressed to this airport, in in-transit list order ------------------
    while (m0 | m1) {
        uint64_t best = ~0ull; uint32_t bj = 0;
        for (uint64_t m = m0; m; m &= m - 1) {
            const uint32_t j = (uint32_t)ba_ctz64(m);
            const uint64_t o = ba_due_order(due + BA_DUE_WORDS * j);
            if (o < best) { best = o; bj = j; }
        }
        for (uint64_t m = m1; m; m &= m - 1) {
            const uint32_t j = 64u + (uint32_t)ba_ctz64(m);
            const uint64_t o = ba_due_order(due + BA_DUE_WORDS * j);
            if (o < best) { best = o; bj = j; }
        }
        const uint32_t *de = due + BA_DUE_WORDS * bj;
        ba_push_one(c, acc, P, de[0] & BA_FID_MASK, __uint_as_float_portable(de[3]), BA_LD_PAIR(BaPairF, de + 4).a, 0.0f);
        if (bj < 64u) m0 &= ~(1ull << bj); else m1 &= ~(1ull << (bj - 64u));
    }
    if (general) {
        // very busy ticks (exact variant): repeated selection over the whole buffer
        uint64_t last = 0; bool first = true;
        for (;;) {
            uint64_t best = ~0ull; uint32_t bj = 0xFFFFFFFFu;
            for (uint32_t j = 0; j < n_due; ++j) {
                const uint32_t *de = due + BA_DUE_WORDS * j;
                if ((de[2] >> 16) != (uint32_t)a) continue;
                const uint64_t o = ba_due_order(de);
                if ((first || o > last) && o < best) { best = o; bj = j; }
            }
            if (bj == 0xFFFFFFFFu) break;
            const uint32_t *de = due + BA_DUE_WORDS * bj;
            ba_push_one(c, acc, P, de[0] & BA_FID_MASK, __uint_as_float_portable(de[3]), BA_LD_PAIR(BaPairF, de + 4).a, 0.0f);
            last = best; first = false;
        }
    }
    // ---- this airport's newly ready departures: a static run of the staged slice ---------
    if (has_run) {
        const uint32_t ri = c.run_info[a];
        const uint32_t cnt = ri & 0xFFFFu;
        const uint32_t *d = c.dstage_base + (uint32_t)parity * c.dstage_stride +
                            BA_DSTAGE_WORDS * (ri >> 16);
        for (uint32_t i = 0; i < cnt; ++i, d += BA_DSTAGE_WORDS) {
            const BaPairUF wq = BA_LD_PAIR(BaPairUF, d);
            const BaPairF xy = BA_LD_PAIR(BaPairF, d + 2);
            ba_push_one(c, acc, P, wq.a, wq.b, xy.a, xy.b);
        }
    }
    ba_push_end(c, a, P);
}

// Owner pull of the single-barrier scan: the slice's slots sit at [0, n_slice) of the buffer,
// the late entries at [cap - n_late, cap).  Slots of flights that had not departed carry the
// invalid (= largest) key: they sort last and end the selection.  After the arrivals the
// airport appends its staged departures and asks for its next calendar run.
BA_DEV void ba_phase_c_pull_fast(BaCtx &c, BaAcc &acc, int a, int parity, uint32_t n_slice,
                                 uint32_t n_late, int tick)
{
    const bool has_run = c.run_tick[a] == (uint32_t)tick;
    uint64_t m0 = 0, m1 = 0;
    const uint32_t *due = c.due_base + (uint32_t)parity * c.due_stride;
    const uint32_t *ow = due + BA_DUE_WORDS * c.due_cap;
    const uint32_t aa = (uint32_t)a | ((uint32_t)a << 16);
    const uint32_t cap = (uint32_t)BA_DUE_CAP_S;
#define BA_OWNER_BITS(g, bits)                                                                  \
    do {                                                                                        \
        const uint32_t w0 = ow[4 * (g)], w1 = ow[4 * (g) + 1], w2 = ow[4 * (g) + 2], w3 = ow[4 * (g) + 3]; \
        uint32_t x;                                                                             \
        bits = 0;                                                                               \
        x = w0 ^ aa; bits |= ((x & 0xFFFFu) == 0u) | (((x >> 16) == 0u) << 1);                  \
        x = w1 ^ aa; bits |= (((x & 0xFFFFu) == 0u) << 2) | (((x >> 16) == 0u) << 3);           \
        x = w2 ^ aa; bits |= (((x & 0xFFFFu) == 0u) << 4) | (((x >> 16) == 0u) << 5);           \
        x = w3 ^ aa; bits |= (((x & 0xFFFFu) == 0u) << 6) | (((x >> 16) == 0u) << 7);           \
    } while (0)
    // the slice: with at most 64 slots, the bit mask the copy-issuing threads filled says it
    // all; else groups from the bottom (slots at and beyond n_slice hold stale owners)
    uint32_t *imp = &c.inmask[a] + 2 * parity;
    const uint32_t im0 = imp[0], im1 = imp[1];
    if (im0) imp[0] = 0;
    if (im1) imp[1] = 0;
    m0 = n_slice <= 64u ? ((uint64_t)im1 << 32) | (uint64_t)im0 : 0ull;
    const uint32_t g_lo_end = n_slice <= 64u ? 0u : (n_slice + 7u) >> 3;
    for (uint32_t g = 0; g < g_lo_end; ++g) {
        uint32_t bits;
        BA_OWNER_BITS(g, bits);
        const uint32_t left = n_slice - 8u * g;
        if (left < 8u) bits &= (1u << left) - 1u;
        if (g < 8u) m0 |= (uint64_t)bits << (8u * g); else m1 |= (uint64_t)bits << (8u * (g - 8u));
    }
    // the late entries (usually none): groups from the top; slots below cap - n_late are stale
    if (n_late) {
        const uint32_t first_late = cap - n_late;
        for (uint32_t g = first_late >> 3; g < (cap >> 3); ++g) {
            uint32_t bits;
            BA_OWNER_BITS(g, bits);
            if (first_late > 8u * g) bits &= ~((1u << (first_late - 8u * g)) - 1u);
            if (g < 8u) m0 |= (uint64_t)bits << (8u * g); else m1 |= (uint64_t)bits << (8u * (g - 8u));
        }
    }
#undef BA_OWNER_BITS
    if (!(m0 | m1) && !has_run) return;
    BaPush P;
    ba_push_begin(c, a, P);
    // ---- arrivals addressed to this airport, in in-transit list order ------------------
    while (m0 | m1) {
        uint64_t best = ~0ull; uint32_t bj = 0;
        for (uint64_t m = m0; m; m &= m - 1) {
            const uint32_t j = (uint32_t)ba_ctz64(m);
            const uint64_t o = ba_due_order(due + BA_DUE_WORDS * j);
            if (o < best) { best = o; bj = j; }
        }
        for (uint64_t m = m1; m; m &= m - 1) {
            const uint32_t j = 64u + (uint32_t)ba_ctz64(m);
            const uint64_t o = ba_due_order(due + BA_DUE_WORDS * j);
            if (o < best) { best = o; bj = j; }
        }
        if ((uint32_t)(best >> 16) == BA_KEY_INVALID) break;      // only not-yet-departed flights left
        const uint32_t *de = due + BA_DUE_WORDS * bj;
        ba_push_one(c, acc, P, de[0] & BA_FID_MASK, __uint_as_float_portable(de[3]),
                    __uint_as_float_portable(de[4]), 0.0f);
        if (bj < 64u) m0 &= ~(1ull << bj); else m1 &= ~(1ull << (bj - 64u));
    }
    // ---- this airport's newly ready departures: a static run of the staged slice ---------
    if (has_run) {
        const uint32_t ri = c.run_info[a];
        const uint32_t cnt = ri & 0xFFFFu;
        const uint32_t *d = c.dstage_base + (uint32_t)parity * c.dstage_stride +
                            BA_DSTAGE_WORDS * (ri >> 16);
        for (uint32_t i = 0; i < cnt; ++i, d += BA_DSTAGE_WORDS) {
            const BaPairUF wq = BA_LD_PAIR(BaPairUF, d);
            const BaPairF xy = BA_LD_PAIR(BaPairF, d + 2);
            ba_push_one(c, acc, P, wq.a, wq.b, xy.a, xy.b);
        }
        ba_run_advance_async(c, a);
    }
    ba_push_end(c, a, P);
}

// Does this airport still hold work (completion vote, network.py:34-41)?  Every flight
// that is not observed as cancelled eventually lands here, so "in transit or queued
// somewhere" is "not yet landed at its destination".
template <bool TRK = true>
BA_DEV bool ba_airport_busy(const BaCtx &c, int a)
{
    return c.q_head[a] != c.q_tail[a] || c.arr_left[a] != 0u ||
           ((ba_ap_flags<TRK>(c, a) & BA_AP_TRACK_T) &&
            (c.next_dep[a] < c.s_depcnt[a] || c.dl_head[a] != c.dl_tail[a]));
}

// ---------------------------------------------------------------------------
// EPILOGUE, stage 1 (flight-parallel, coalesced): the seven per-flight log-densities
// and every per-flight gradient contribution.  Global sums
// (log-prob, d/d scales) stay in the thread's accumulator; contributions owned by an
// airport or a route are written to scratch for stage 2:
//   xdk[k]   (dep order)     origin's  sum r W        (value mode: x_dep)
//   xak[j]   (arrival order) dest's    sum r W        (value mode: x_arr)
//   atk[f]                   route's   d/dT term
//   ct[k|j]                  d/d mean_turnaround term (noise mode: aircraft branch of the
//                            departure, dep order; value mode: landing term, arrival order)
// ---------------------------------------------------------------------------
template <bool TRK = true>
BA_DEV void ba_epilogue_flight(BaCtx &c, BaAcc &acc, int f)
{
    const BaFlight fl = ba_flight(c, f);
    const uint32_t cobs = (fl.od >> 24) & 3u;
    const int k = c.pos_dep[f];
    const int o = (int)(fl.od & 0xFFFu), d = (int)((fl.od >> 12) & 0xFFFu);
    // *_cancelled (network.py:100-109) at an airport whose cancellation probability is
    // exactly 0 (ba_plan.h: BA_SIMPLE_RATIO); other airports score it in the scan
    if (!(ba_ap_flags<TRK>(c, o) & BA_AP_TRACK_T)) acc.lp += (double)(cobs == 1u ? c.c1_lp : c.c0_lp);
    if (cobs == 1u) {                 // observed as cancelled: never queued anywhere
        // (value mode keeps the turnaround terms in ARRIVAL order in ct: position k of that
        //  array belongs to another, live flight there -- leave it alone)
        if (c.want_grad) { c.xdk[k] = 0.0f; if (!c.value_mode) c.ct[k] = 0.0f; c.atk[f] = 0.0f; }
        return;
    }
    const int j = c.pos_arr[f];
    const BaNoise4 nz = ba_noise_again(c, f);
    const int rid = c.route[f];
    const float T = c.lat_route[rid];
    const bool grad = c.want_grad != 0;
    const float half_inv_var = 0.5f * c.inv_sigma2;
    const float lp_const = -c.log_sigma - BA_LOG_SQRT_2PI;

    // ---- departure side: *_departure_service_time, *_simulated_departure_time
    //      (airport.py:144-147,175-182)
    const float rate_o = c.rate[o];
    const float xd = c.value_mode ? nz.e_dep : BA_DIV(nz.e_dep, rate_o);
    const float wdep = c.wd[f];
    const bool track_a = (ba_ap_flags<TRK>(c, o) & BA_AP_TRACK_A) != 0u;
    const float qstart_d = track_a ? c.qd[f] : fmaxf(0.0f, fl.sched_dep);
    const float mu_d = BA_ADD(qstart_d, wdep);              // queue_start + total_wait
    const float dy_d = fl.act_dep - mu_d;
    float lp = (BA_LOG(rate_o) - rate_o * xd) + (lp_const - dy_d * dy_d * half_inv_var);

    // ---- arrival side: *_travel_time, *_arrival_service_time,
    //      *_simulated_arrival_time, *_turnaround_time
    //      (network.py:158-163; airport.py:144-147,197-204,217-220)
    const float rate_d = c.rate[d], mt = c.mt[d], tstd = c.tstd[d];
    const float sc = BA_MUL(T, c.v_t);
    const float tau = c.value_mode ? nz.eps_travel : BA_ADD(T, BA_MUL(nz.eps_travel, sc));
    const float qstart_a = BA_ADD(fl.act_dep, tau);
    const float warr = c.wa[f];
    const float mu_a = BA_ADD(qstart_a, warr);
    const float dy_a = fl.act_arr - mu_a;
    const float xa = c.value_mode ? nz.e_arr : BA_DIV(nz.e_arr, rate_d);
    const float u = c.value_mode ? nz.eps_turn : BA_ADD(mt, BA_MUL(nz.eps_turn, tstd));
    const float dz = (tau - T) / sc;
    const float du = (u - mt) / tstd;
    lp += (BA_LOG(rate_d) - rate_d * xa) + (lp_const - dy_a * dy_a * half_inv_var);
    float lp2 = (-0.5f * dz * dz - BA_LOG(sc) - BA_LOG_SQRT_2PI)
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
            float gvu = -c.inv_vu, turn = 0.0f;
            if (track_a) {
                const float wgt = c.sw[f];
                if (wgt > 0.0f) {
                    // ready = max(act_arr_g + u_g, sched), u_g = m~ (1 + v_u eps_g)
                    const float se = c.se[f];
                    turn = wgt * r_d * (1.0f + c.v_u * se);
                    gvu += wgt * r_d * se * c.mt[o];
                }
            }
            if (TRK) c.ct[k] = turn;          // only summed for airports that track aircraft
            acc.g_vu += gvu;
        } else {
            c.xdk[k] = xd;
            c.xak[j] = xa;
            c.atk[f] = dz / sc + (dz * dz - 1.0f) / T;
            acc.g_vt += (dz * dz - 1.0f) * c.inv_vt;
            c.ct[j] = du / tstd + (du * du - 1.0f) / mt;
            acc.g_vu += (du * du - 1.0f) * c.inv_vu;
        }
    }
}

// ---------------------------------------------------------------------------
// EPILOGUE, stage 2: owner sums over contiguous slices, fixed order, no atomics.
// ---------------------------------------------------------------------------
template <bool TRK = true>
BA_DEV void ba_epilogue_airport(BaCtx &c, int a, int Ng)
{
    const BaAirport A = c.ap[a];
    const float m = c.m[a], mt = c.mt[a];
    float s_dep = 0.0f, s_arr = 0.0f, s_turn = 0.0f;
    for (int i = 0; i < A.dep_cnt; ++i) s_dep += c.xdk[A.dep_off + i];
    for (int i = 0; i < A.arr_cnt; ++i) s_arr += c.xak[A.arr_off + i];
    const float n_ent = (float)(A.dep_live + A.arr_cnt), n_land = (float)A.arr_cnt;
    float g_serv, g_turn;
    if (!c.value_mode) {
        if (TRK && (A.flags & BA_AP_TRACK_A))
            for (int i = 0; i < A.dep_cnt; ++i) s_turn += c.ct[A.dep_off + i];
        // x = e m  =>  d/dm = sum r W / m - (#entries)/m ;  u = m~ (1 + v_u eps)
        g_serv = (s_dep + s_arr) / m - n_ent / m;
        g_turn = s_turn - n_land / mt;
    } else {
        for (int i = 0; i < A.arr_cnt; ++i) s_turn += c.ct[A.arr_off + i];
        // log Exponential(x; 1/m) = -log m - x/m
        g_serv = -n_ent / m + (s_dep + s_arr) / (m * m);
        g_turn = s_turn;
    }
    const int g = A.global;
    c.grow[0 * Ng + g] = g_turn;
    c.grow[1 * Ng + g] = g_serv;
    if (c.cancel_mode) {
        c.grow[2 * Ng + g] = c.acc_ln0[a] * c.n0[a];
        c.grow[3 * Ng + g] = c.acc_bc[a];
    }
}

BA_DEV void ba_epilogue_route(BaCtx &c, int r)
{
    float s = 0.0f;
    const int b = c.rt_off[r], e = c.rt_off[r + 1];
    for (int i = b; i < e; ++i) s += c.atk[c.rt_flt[i]];
    c.grow[c.route_base + c.rt_gid[r]] = s;
}
