// Internal: kernel parameter blocks + launch wrappers (implemented in ba_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ba_plan.h"

struct BaFwdParams {
    BaPlanView plan;              // .days = DEVICE array of BaDayView with device pointers
    int32_t P;
    const float *lat_airport;     // [P, C, N]
    const float *lat_route;       // [P, R]
    const float *scales;          // [3]
    const float *noise;           // [P, D, Fmax, 4] or nullptr
    const float *noise2;          // predictive: [P, D, Fmax, 4] or nullptr (Philox)
    float *sim;                   // predictive: [P, D, Fmax, 3] out
    uint64_t seed;
    const uint64_t *seed_dev;     // if set, the seed is read from device memory
    uint32_t particle_offset;     // added to the particle index in the Philox counters
    uint32_t day_offset;          // added to the day index in the Philox counters
    int32_t value_mode, want_grad, replay_waits;
    float *logp;                  // [P, D]
    uint32_t *status;             // [P, D]
    int32_t *ticks;               // [P, D] or nullptr
    float *tape;                  // [P, D, grad_stride] or nullptr
    int32_t *pop_tick;            // [P, D, Fmax, 2] or nullptr
    float *vtape;                 // value mode: [P, D, Fmax, 4] d logp / d site values, or nullptr
    uint32_t *scratch;            // per-CTA global scratch
    uint64_t scratch_stride;      // words per CTA
    unsigned int *work_counter;   // dynamic work distribution
    const int32_t *work_list;     // optional explicit list of pd = p*D + d
    int32_t n_work;
    const unsigned int *n_work_dev; // if set, overrides n_work (device-side count)
    unsigned long long *phase_cycles; // optional [8] per-phase cycle totals (debug, BA_PHASE_PROFILE)
    int32_t wi_begin;             // first work item of this launch (a batch run as two launches)
};

struct BaBwdParams {
    int32_t P, D, C, N, R, stride;
    const float *dlogp;           // [P, D]
    const float *tape;            // [P, D, stride]
    float *g_airport;             // [P, C*N]
    float *g_route;               // [P, R]
    float *g_scales;              // [P, 3]
};

int ba_forward_block_threads(int n_airports);
int ba_forward_target_ctas(int block_threads);
size_t ba_forward_smem_bytes(int n_airports, uint32_t list_words, bool exact, bool tracked,
                             int max_dslice);
cudaError_t ba_forward_configure(int block, size_t smem_fast, size_t smem_exact, bool tracked);
int ba_forward_max_resident(int block_threads, size_t smem_bytes, bool exact, bool tracked, int device);

cudaError_t ba_launch_forward(const BaFwdParams &p, bool exact, bool tracked, int grid, int block,
                              size_t smem, cudaStream_t s);
// ring-less scan (ba_scan2.cuh): one-warp CTAs
size_t ba_forward2_smem_bytes(int n_airports);
uint64_t ba_forward2_scratch_words(int Fmax, int n_airports, int n_routes);
int ba_forward2_warps(int n_airports);
int ba_forward2_configure(int nw, size_t smem, int device);
cudaError_t ba_launch_forward2(const BaFwdParams &p, int nw, int grid, size_t smem, cudaStream_t s);
// mean-field guide kernels (ba_guide.cu)
struct BaMfParams {
    int32_t P, D, N, R, n, stride;
    uint32_t particle_offset;
    uint64_t seed;
    const uint64_t *seed_dev;
    const float *params;          // [2n] loc | log_scale
    const int32_t *route_of;      // [N*N] route index of (origin, dest) or -1
    float *lat_airport, *lat_route, *lp0, *z;
    const float *tape, *logp;
    float inv_count_flights;
    float prior_weight;           // share of the per-particle terms (prior, -log q) counted here
    float *work;                  // [chunks, 2n + 4]
    int32_t chunks;
    float *flat;
};
cudaError_t ba_launch_mf_sample(const BaMfParams &p, cudaStream_t s);
cudaError_t ba_launch_mf_backward(const BaMfParams &p, cudaStream_t s);
int ba_mf_chunks(int P);
// rational-quadratic spline of the flow guide (ba_guide.cu)
cudaError_t ba_launch_rqs_forward(const float *x, const float *params, float *y, float *ld,
                                  long long M, int bins, float bound, cudaStream_t s);
cudaError_t ba_launch_rqs_backward(const float *x, const float *params, const float *gy,
                                   const float *gld, float *gx, float *gparams, long long M,
                                   int bins, float bound, cudaStream_t s);
cudaError_t ba_launch_predict(const BaFwdParams &p, int grid, int block, size_t smem,
                              cudaStream_t s);
cudaError_t ba_launch_backward(const BaBwdParams &p, cudaStream_t s);
cudaError_t ba_launch_backward_values(const float *dlogp, const float *vtape, float *g_noise,
                                      long long n_pd, int row, cudaStream_t s);
cudaError_t ba_launch_noise(const BaPlanView &plan, int P, uint64_t seed, float *noise,
                            cudaStream_t s);
// gathers the particle-days whose status has `bit` set into list (device), count in *n
cudaError_t ba_launch_collect(const uint32_t *status, int n, uint32_t bit, int32_t *list,
                              unsigned int *count, cudaStream_t s);
