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

// queue entry word: [31] departure, [30:29] aircraft-branch weight code, [17:0] flight
#define BA_Q_DEP 0x80000000u
#define BA_Q_W_SHIFT 29
// bag entry word: [31:18] origin airport (day order), [17:0] flight
#define BA_BAG_O_SHIFT 18

// airport flags
#define BA_AP_TRACK_T 1u              // turnaround list needed (n_avail can matter)
#define BA_AP_TRACK_A 2u              // available-aircraft FIFO contents needed

// no-cancellation mode: an airport with at most this many departures a day has
// cancellation probability exactly 0 in fp32 at every tick (1000 reserves,
// sigmoid(10*(n/nr - 1/4)) == 1.0f for n/nr >= 1.93), so neither the turnaround
// list nor the aircraft FIFO can influence anything.
#define BA_SIMPLE_MAX_DEPARTURES 340

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
    uint32_t bag_off_s, bag_cap_s;    // inbound bag,       shared-memory variant
    uint32_t q_off_x, q_mask_x;       // exact-capacity variants (global scratch)
    uint32_t bag_off_x, bag_cap_x;
    uint32_t turn_off;                // turnaround list (cap = arr_cnt), global scratch
    uint32_t av_off, av_mask;         // aircraft FIFO ring, global scratch
    uint32_t dl_off, dl_mask;         // delayed-departure deque ring, global scratch
    uint32_t pad0, pad1;
};

struct BaDayView {
    int32_t N, F;
    const BaFlight *flights;          // [F]
    const int32_t *route;             // [F] compact global route id
    const int32_t *dep_fid;           // [F] flights grouped by origin, ascending
    const BaAirport *airports;        // [N] day order
    const int32_t *lane_airport;      // [N] airport handled by lane slot i (load-sorted)
    uint32_t q_total_s, bag_total_s;  // pool sizes (entries) of the shared variant
    uint32_t q_total_x, bag_total_x;  // pool sizes of the exact variant
    uint32_t turn_total, av_total, dl_total;
    int32_t first_tick;               // ticks [1, first_tick) are provably idle
    float t_before_first;             // clock after first_tick-1 sequential fp32 adds
};

struct BaPlanView {
    int32_t D, N, Fmax, R, C;         // days, airports, max flights, routes, latent rows
    int32_t include_cancellations;
    int32_t max_ticks;
    float delta_t, max_t;
    int32_t grad_stride;              // C*N + R + 3
    const BaDayView *days;            // [D]
};

// scratch words (4 B) one CTA needs for the exact variant / the always-global lists
static inline uint64_t ba_scratch_words(const BaDayView &d, bool exact)
{
    uint64_t w = 0;
    if (exact) w += 4ull * d.q_total_x + 2ull * d.bag_total_x;
    w += 2ull * d.turn_total + 2ull * d.av_total + 1ull * d.dl_total;
    return w;
}
