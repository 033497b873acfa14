// Internal: static per-day tables ("plan") shared by all particles.
//
// Built once on the host from BaDayTable (include/bayesair_b200.h) by
// ba_plan_build.cpp, uploaded to the device by ba_capi.cu.  The same POD views
// are used by the CUDA kernel and by the host emulation harness in tests/.
#pragma once
#include <stdint.h>

// ---- packing limits -------------------------------------------------------
#define BA_MAX_AIRPORTS 4096          // 12 bits in BaFlight::od
#define BA_MAX_FLIGHTS (1 << 18)      // 18 bits of flight id in queue / bag words
#define BA_FID_MASK 0x3FFFFu

// queue entry word: [31] departure, [29:18] destination airport, [17:0] flight
#define BA_Q_DEP 0x80000000u
#define BA_Q_DST_SHIFT 18
// dep_word: [31:30] observed cancellation (0, 1, 3 = None), [29:18] destination, [17:0] flight
#define BA_DW_CANCEL_SHIFT 30
#define BA_DW_ENTRY_MASK 0x3FFFFFFFu
// in-transit ordering key: [27:12] tick of the departure's service, [11:0] origin airport
#define BA_KEY_TICK_SHIFT 12
#define BA_KEY_INVALID 0xFFFFFFFFu
// arrival calendar: buckets (ticks) kept in shared memory by the fast variant
#ifndef BA_KCAL
#define BA_KCAL 384
#endif
#ifndef BA_DUE_CAP_S
#define BA_DUE_CAP_S 128              // entries per due buffer in shared memory (two buffers)
#endif
#define BA_KEY_DEPARTURE 0xF0000000u   // keys of newly ready departures sort after arrivals
#define BA_DUE_WORDS 6               // word, key, owner<<16|seq, queue start, x, y
#define BA_DSTAGE_WORDS 4            // staged departure: queue word, queue start, x, arrival time

// airport flags
#define BA_AP_TRACK_T 1u              // turnaround list needed (n_avail can matter)
#define BA_AP_TRACK_A 2u              // available-aircraft FIFO contents needed

// No-cancellation mode (1000 reserve aircraft per airport, model.py:116-118): the
// cancellation probability of a batch of n_ready flights is
//   p = 1 - sigmoid(10 (n_avail / n_ready - 1/4)),
// which is EXACTLY 0 in fp32 once n_avail / n_ready >= 1.93 (sigmoid rounds to 1.0f).
// n_avail never drops below 1000 - (departures of the day) and n_ready never exceeds the
// largest number of departures scheduled inside one tick, both known at plan time.  An
// airport that satisfies (1000 - D_a) >= BA_SIMPLE_RATIO * max_batch_a can therefore
// neither cancel nor delay anything and needs no aircraft bookkeeping at all
// (SURVEY.md App. A.2 Q4).
#define BA_SIMPLE_RATIO 1.95

struct BaFlight {                     // 16 B, one LDG.128
    float sched_dep;
    float act_dep;                    // NaN = unobserved
    float act_arr;                    // NaN = unobserved
    uint32_t od;                      // [11:0] origin, [23:12] dest, [25:24] cancel obs (3 = None)
};

struct BaAirport {                    // 80 B static row per (day, airport)
    int32_t dep_off, dep_cnt;         // slice of BaDayView::dep_fid
    int32_t arr_cnt;
    int32_t global;                   // latent index
    uint32_t flags;
    uint32_t q_off_s, q_mask_s;       // runway-queue ring, shared-memory variant
    uint32_t q_off_x, q_mask_x;       // exact-capacity variant (global scratch)
    uint32_t lq_off;                  // ring-less scan (ba_scan2.cuh): first live departure of
                                      //   this airport in BaDayView::lq (one sentinel per airport)
    uint32_t as_off;                  // ... first slot of its arrival list (arr_cnt + 1 slots)
    uint32_t slot;                    // lane slot that runs this airport (BaDayView::lane_airport)
    uint32_t pad5;
    uint32_t turn_off;                // turnaround list (cap = arr_cnt), global scratch
    uint32_t av_off, av_mask;         // aircraft FIFO ring, global scratch
    uint32_t dl_off, dl_mask;         // delayed-departure deque ring, global scratch
    int32_t arr_off;                  // slice of arrival order (arr_cnt entries)
    int32_t dep_live;                 // departures not observed as cancelled
};

// One live (= not observed as cancelled) departure of a bookkeeping-free airport, in
// (origin, pending order) order: the tick at which it becomes ready (network.py:60-68) and
// its queue start max(0, scheduled departure) (network.py:118-126 with 1000 reserves).
struct BaLiveDep {
    uint32_t tick; float qstart;
    uint32_t dslot;                   // slot (BaDayView::lane_sorted) that runs the destination
    uint32_t das_off;                 // BaAirport::as_off of the destination
};

struct BaDayView {
    int32_t N, F;
    const BaFlight *flights;          // [F]
    const int32_t *route;             // [F] compact global route id
    const int32_t *dep_fid;           // [F] flights grouped by origin, ascending ("dep order")
    const float *dep_sched;           // [F] scheduled departure, dep order
    const uint32_t *dep_word;         // [F] flight | dest << 18 | cancel obs << 30, dep order
    const int32_t *pos_dep;           // [F] flight -> position in dep order
    const int32_t *pos_arr;           // [F] flight -> position in arrival order (-1: cancelled)
    const int32_t *rt_gid;            // [Rd] global route id of the day's r-th used route
    const int32_t *rt_off;            // [Rd+1] CSR offsets into rt_flt
    const int32_t *rt_flt;            // [sum arr_cnt] flights of each used route, ascending
    int32_t Rd;                       // routes used on this day
    const int32_t *dcal_off;          // [last_ready_tick + 2] departure calendar: slice of tick k
    const uint32_t *dcal_rec;         //   = flights of bookkeeping-free airports that become
                                      //   ready at tick k, grouped by origin airport, pending
                                      //   order inside a group; record = the static half of the
                                      //   flight's queue entry {queue word, queue start (fp32)}
    const int32_t *dep_cpos;          // [F] dep order -> calendar record (-1: not on the calendar)
    int32_t last_ready_tick;          // tick at which the day's last scheduled flight is ready
    const int32_t *drun_off;          // [N+1] per-airport slices of drun (in units of runs)
    const uint32_t *drun;             // runs {tick, first record relative to the tick's slice
                                      //   << 16 | count}: the contiguous part of a tick's
                                      //   calendar slice that belongs to one airport;
                                      //   every airport's list ends with a sentinel run
    int32_t max_dslice;               // largest calendar slice of any tick of the day
    const BaAirport *airports;        // [N] day order
    const int32_t *lane_airport;      // [N] airport handled by lane slot i (load-sorted)
    const int32_t *lane_sorted;       // [N] airports by decreasing traffic (ring-less scan)
    uint32_t q_total_s;               // queue pool size (entries) of the shared variant
    uint32_t q_total_x;               // queue pool size of the exact variant
    uint32_t turn_total, av_total, dl_total;
    uint32_t any_track_t, any_track_a; // does any airport of the day track aircraft
    int32_t first_tick;               // ticks [1, first_tick) are provably idle
    float t_before_first;             // clock after first_tick-1 sequential fp32 adds
    // ring-less scan (ba_scan2.cuh)
    const BaLiveDep *lq;              // [n_live + N] live departures grouped by origin + sentinels
    const int32_t *pos_live;          // [F] flight -> index into lq ("lpos"), -1: cancelled
    int32_t n_live, n_arr;            // live departures / arrivals of the day (equal)
};

struct BaPlanView {
    int32_t D, N, Fmax, R, C;         // days, airports, max flights, routes, latent rows
    int32_t include_cancellations;
    int32_t any_track;                // some airport of some day tracks aircraft
    int32_t max_dslice;               // largest departure-calendar slice over days and ticks
    int32_t max_ticks;
    float delta_t, max_t;
    int32_t grad_stride;              // C*N + R + 3
    int32_t scan2_ok;                 // the ring-less scan (ba_scan2.cuh) can run this plan
    const float *tick_time;           // [max_ticks + 2] clock after k sequential fp32 adds
    const BaDayView *days;            // [D]
};

// Per-CTA global scratch layout (4-byte words), in this order:
//   xdk, xak, atk   [F] each, dep order : service draws x_dep, x_arr and the arrival
//                                         time y_dep + tau, derived by the prologue;
//                                         dead after the scan and reused by the epilogue
//                                         for per-flight gradient contributions
//   ct              [F]                 : epilogue contributions to d/d mean_turnaround
//   xaf             [F]                 : x_arr by flight (departures served after arrival)
//   wd, wa          [F] each            : waits recorded by the scan ("tape")
//   karr            [F]                 : arrival tick of every flight
//   cal             [3F]                : arrival calendar records {flight|dest, arrival
//                                         time, x_arr}, counting-sorted by arrival tick
//   popkey, popseq  [F] each            : in-transit ordering key of served departures
//   trl, tre        [F] each if any_track_t : turnaround release time / its noise
//   qd, se, sw      [F] each if any_track_a : departure queue start, aircraft-source
//                                         noise and branch weight (adjoint of max())
//   runway queues 6 x q_total_x (exact variant: the rings; fast variant: spill area)
//   exact variant only: two due buffers 6 x F each,
//                       calendar counters max_ticks + 2
//   turnaround lists 2 x turn_total, aircraft FIFOs 2 x av_total, delayed deques dl_total
static inline uint64_t ba_scratch_words(const BaDayView &d, bool exact, int max_ticks,
                                        int max_dslice_plan)
{
    uint64_t w = 20ull * (uint64_t)d.F + 2;   // incl. kept noise and the calendar-ordered {x_dep, arrival} pairs
    if (d.any_track_t) w += 2ull * (uint64_t)d.F;
    if (d.any_track_a) w += 3ull * (uint64_t)d.F;
    w += 7ull * d.q_total_x;          // draw history + exact rings (exact) / spill area (fast)
    if (exact) w += 2ull * BA_DSTAGE_WORDS * ((uint64_t)max_dslice_plan + 2) + 4 +
                    2ull * (uint64_t)BA_DUE_WORDS * (uint64_t)d.F +
                    2ull * (((uint64_t)d.F + 7) / 8 * 4) + (uint64_t)max_ticks + 2;
    w += 2ull * d.turn_total + 2ull * d.av_total + 1ull * d.dl_total;
    return w + 16;
}
