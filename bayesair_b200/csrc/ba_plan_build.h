// Internal: host-side schedule compiler (BaDayTable -> plan tables).
#pragma once
#include <string>
#include <vector>

#include "../../include/bayesair_b200.h"
#include "ba_plan.h"

struct BaHostDay {
    std::vector<BaFlight> flights;
    std::vector<int32_t> route, dep_fid, lane_airport, lane_sorted, pos_dep, pos_arr, rt_gid, rt_off, rt_flt;
    std::vector<float> dep_sched;
    std::vector<uint32_t> dep_word, dcal_rec;
    std::vector<int32_t> dcal_off, drun_off, dep_cpos;
    std::vector<uint32_t> drun;
    std::vector<BaAirport> airports;
    std::vector<BaLiveDep> lq;
    std::vector<int32_t> pos_live;
    BaDayView view;   // pointers into the vectors above (host view)
};

struct BaHostPlan {
    std::vector<BaHostDay> days;
    std::vector<BaDayView> day_views;   // host views, contiguous
    std::vector<int32_t> route_origin, route_dest;   // global airport indices per route
    BaPlanView view;                    // .days -> day_views.data()
    uint64_t scratch_words_fast = 0;    // per-CTA scratch, shared-memory variant
    uint64_t scratch_words_exact = 0;   // per-CTA scratch, exact variant
    uint32_t smem_list_words = 0;       // max over days: 6*q_total_s
    int32_t max_dslice = 0;             // max over days and ticks of the departure slice
    std::vector<float> tick_time;       // [max_ticks + 2]
};

// Returns BA_OK or an error status with *err filled.
int ba_build_host_plan(const BaDayTable *days, int n_days, int n_airports, float delta_t,
                       float max_t, int include_cancellations, int max_ticks,
                       BaHostPlan *out, std::string *err, double q_div = 0.0, double q_min = 0.0);
