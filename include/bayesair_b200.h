/*
 * bayesair_b200 -- C ABI of the B200-native (sm_100a) hot path of dawsonc/BayesAir.
 *
 * The reference has no native code and no FFI for this path: the
 * boundary it exposes is the Python callable
 *     bayes_air.model.air_traffic_network_model(states, delta_t, max_t, device,
 *                                               include_cancellations)
 * (/root/reference/bayes_air/model.py:14-20) evaluated under Pyro effect handlers
 * and followed by Trace.log_prob_sum() + backward()
 * (/root/reference/scripts/wn/train.py:301-316).  The entry points below are what
 * a ctypes / cffi binding on the reference side binds to replace that evaluation
 * (see INTEGRATION.md); each comment names the reference lines it stands for.
 *
 * Conventions
 *   - plain C, no exceptions across the boundary: every call returns a BaStatus,
 *     ba_last_error() gives the message of the last failure on the calling thread;
 *   - all pointers named d_* are DEVICE pointers owned by the caller (allocate them
 *     with any allocator: cudaMalloc, the PyTorch caching allocator ...); pointers
 *     named h_* are host pointers; the library allocates nothing per call except
 *     inside the opaque plan (static day tables + per-CTA scratch);
 *   - all work is enqueued on the caller's stream (passed as void* = cudaStream_t),
 *     no hidden synchronisation in ba_forward / ba_backward (a batch of just over one
 *     wave of particle-days forks part of its work to a stream of the plan and joins it
 *     back with events: still ordered on the caller's stream, and capturable);
 *   - a plan is immutable after creation and bound to one device; calls on one plan
 *     must be stream-ordered (they share the plan's scratch).
 *   - all model arithmetic is fp32, like the reference's.
 */
#ifndef BAYESAIR_B200_H
#define BAYESAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BA_VERSION 200

typedef enum BaStatus {
    BA_OK = 0,
    BA_ERR_BAD_ARG = 1,
    BA_ERR_CUDA = 2,
    BA_ERR_NO_DEVICE = 3,
    BA_ERR_LIMIT = 4          /* a static limit of the kernel (airports, flights) exceeded */
} BaStatus;

/* per-(particle, day) status word written by ba_forward (bit set) */
#define BA_PD_OK 0u
#define BA_PD_TICK_CAP 1u       /* simulation did not complete within max_ticks            */
#define BA_PD_OVERFLOW 2u       /* a shared-memory list overflowed; ba_forward re-runs these
                                   particle-days with exact-capacity lists automatically,
                                   so callers only see this bit if that re-run is disabled */
#define BA_PD_UNOBSERVED 4u     /* an observation was missing in an observed-mode call     */
#define BA_PD_INTERNAL 8u

/* One network day, host side, flights ALREADY in stable scheduled-departure order
 * (NetworkState.__post_init__, bayes_air/network.py:30-32) and airports in that
 * day's dict order (bayes_air/schedule.py:100-104).  Replaces the per-day
 * NetworkState object graph handed to the model (bayes_air/model.py:15,126-127). */
typedef struct BaDayTable {
    int32_t n_airports;
    int32_t n_flights;
    const float *h_sched_dep;       /* [F] scheduled_departure_time                    */
    const float *h_act_dep;         /* [F] actual_departure_time, NaN = None           */
    const float *h_act_arr;         /* [F] actual_arrival_time,   NaN = None           */
    const int8_t *h_cancel_obs;     /* [F] actually_cancelled 0/1, -1 = None           */
    const int32_t *h_origin;        /* [F] day-local airport index                     */
    const int32_t *h_dest;          /* [F]                                             */
    const int32_t *h_airport_global;/* [N] day-local -> latent index (states[0] order) */
} BaDayTable;

typedef struct BaPlan BaPlan;

typedef struct BaPlanInfo {
    int32_t n_days;
    int32_t n_airports;         /* N (identical on all days, model.py:26-27)            */
    int32_t max_flights;        /* Fmax over days: row stride of the noise tensor       */
    int32_t n_routes;           /* R: distinct (origin, dest) pairs used on any day     */
    int32_t n_airport_latents;  /* C: 2 without cancellations, 4 with                   */
    int32_t grad_stride;        /* floats per (particle, day) row of the gradient tape  */
    int32_t include_cancellations;
    int32_t device;
    float delta_t, max_t;
    int32_t max_ticks;
    int32_t block_threads;      /* CTA size the forward kernel is launched with         */
    int32_t smem_bytes;         /* dynamic shared memory per CTA (fast variant)         */
    int32_t resident_ctas;      /* persistent grid size                                 */
} BaPlanInfo;

/* Options of a forward call. */
typedef struct BaForwardOpts {
    int32_t noise_is_value;     /* 0: d_noise holds standard draws (Exp(1), N(0,1), Exp(1),
                                      N(0,1)) pushed through the reference's rsample
                                      arithmetic in-kernel (bayes_air samples these sites
                                      in-model: airport.py:144-147,217-220, network.py:158-163)
                                   1: d_noise holds the site VALUES (service time, travel
                                      time, service time, turnaround time), i.e. the sites
                                      were fixed by a guide / condition()                    */
    int32_t want_grad;          /* 1: also fill the gradient tape for ba_backward          */
    int32_t force_exact_lists;  /* 1: skip the shared-memory variant (testing)             */
    int32_t no_overflow_rerun;  /* 1: leave BA_PD_OVERFLOW particle-days unresolved (testing) */
    int32_t replay_waits;       /* 1: always form queue waits with the reference's sequential
                                      fp32 adds instead of only when a decision is within
                                      rounding distance of the clock (testing; same results) */
    int32_t force_ring_lists;   /* 1: skip the ring-less scan, use the shared-ring kernel (testing) */
    uint64_t seed;              /* Philox seed, used when d_noise == NULL                  */
    int32_t *d_pop_tick;        /* nullable diagnostics [P, D, Fmax, 2]: tick at which each
                                   flight left its departure / arrival runway queue (-1 =
                                   never); the "queue decisions" parity tests compare      */
    const uint64_t *d_seed;     /* nullable: the Philox seed is read from DEVICE memory when the
                                   kernel starts (a captured CUDA graph replays with a new seed
                                   each step); overrides `seed`                             */
    uint32_t particle_offset;   /* index of this call's first particle in the global batch:
                                   Philox counters use particle_offset + p, so ranks that share
                                   a seed draw different per-flight noise                   */
    uint32_t day_offset;        /* index of the plan's first day in the whole data set: Philox
                                   counters use day_offset + d, so a rank whose plan holds a
                                   subset of the days draws what the full plan would       */
    float *d_value_tape;        /* nullable, needs noise_is_value: [P, D, Fmax, 4] gradient of
                                   logp[p, d] w.r.t. the four per-flight site values the noise
                                   tensor holds (what a guide over those sites needs; SURVEY.md
                                   App. A.5; wn_network.py:285-302).  Plans without
                                   cancellations only; particle-days whose lists overflow are
                                   then NOT re-run (their status keeps the overflow bit)     */
} BaForwardOpts;

int ba_version(void);
const char *ba_last_error(void);

/* Compile a list of network days into device tables ("schedule compiler",
 * replaces deepcopy(states) + the per-call Python object walk, model.py:38,126-155).
 * delta_t / max_t / include_cancellations: the model's keyword arguments
 * (model.py:16-19).  max_ticks: hard cap on ticks per day (the reference has none;
 * see BA_PD_TICK_CAP). */
BaStatus ba_plan_create(const BaDayTable *days, int32_t n_days, int32_t n_airports,
                        float delta_t, float max_t, int32_t include_cancellations,
                        int32_t max_ticks, int32_t device, BaPlan **out);
void ba_plan_destroy(BaPlan *plan);
BaStatus ba_plan_info(const BaPlan *plan, BaPlanInfo *out);
/* Compact route table: h_origin[r], h_dest[r] = global airport indices of route r.
 * The dense [N,N] travel-time latents (scripts/wn/train.py:261-267) are gathered
 * down to [R] by the caller. */
BaStatus ba_plan_routes(const BaPlan *plan, int32_t *h_origin, int32_t *h_dest);

/* Bytes of the gradient tape for n_particles (d_tape of ba_forward / ba_backward). */
size_t ba_tape_bytes(const BaPlan *plan, int32_t n_particles);

/* Forward simulation + fused log-joint for n_particles x n_days particle-days.
 * Replaces, per particle: the day loop of air_traffic_network_model
 * (model.py:126-209) and the sum over all per-flight sites of
 * Trace.log_prob_sum() (scripts/wn/train.py:314).
 *   d_lat_airport [P, C, N]  rows: mean_turnaround_time, mean_service_time
 *                            (+ log_initial_available_aircraft, base_cancel_logprob)
 *                            in global airport order              (model.py:59-113)
 *   d_lat_route   [P, R]     travel_time_{o}_{d} of the used routes (model.py:77-87)
 *   d_scales      [3]        runway_use_time_std_dev, travel_time_variation,
 *                            turnaround_time_variation (constrained values, model.py:41-55)
 *   d_noise       [P, D, Fmax, 4] or NULL (-> Philox from opts->seed)
 *   d_logp        [P, D]     out: sum of the 7F per-flight site log-probs of the day
 *   d_status      [P, D]     out: BA_PD_* bits
 *   d_ticks       [P, D]     out (nullable): ticks simulated
 *   d_tape        ba_tape_bytes() or NULL when !want_grad */
BaStatus ba_forward(BaPlan *plan, int32_t n_particles,
                    const float *d_lat_airport, const float *d_lat_route,
                    const float *d_scales, const float *d_noise,
                    const BaForwardOpts *opts,
                    float *d_logp, uint32_t *d_status, int32_t *d_ticks,
                    void *d_tape, void *stream);

/* Pathwise VJP (queue decisions held fixed, as torch autograd does for the
 * reference: scripts/training.py:289): contracts the tape with the cotangent.
 *   d_dlogp       [P, D]   cotangent of d_logp
 *   d_g_airport   [P, C, N], d_g_route [P, R], d_g_scales [P, 3]   out */
BaStatus ba_backward(BaPlan *plan, int32_t n_particles, const float *d_dlogp,
                     const void *d_tape, float *d_g_airport, float *d_g_route,
                     float *d_g_scales, void *stream);

/* Posterior-predictive forward simulation: flights whose observations are
 * missing (h_act_dep / h_act_arr NaN, h_cancel_obs -1, i.e. obs=None at
 * bayes_air/network.py:100-109 and bayes_air/types/airport.py:175-204) are SAMPLED; their
 * sampled departure time drives the downstream arrival like in the reference.  No
 * log-prob, no gradient.
 *   d_noise  [P, D, Fmax, 4] as in ba_forward, or NULL (Philox from opts->seed)
 *   d_noise2 [P, D, Fmax, 4] U(0,1) for `_cancelled`, N(0,1) for `_simulated_departure_time`,
 *            N(0,1) for `_simulated_arrival_time`, unused; or NULL (Philox)
 *   d_sim    [P, D, Fmax, 3] out: cancelled (0/1), departure time, arrival time (NaN where
 *            the flight was cancelled) */
BaStatus ba_predict(BaPlan *plan, int32_t n_particles,
                    const float *d_lat_airport, const float *d_lat_route,
                    const float *d_scales, const float *d_noise, const float *d_noise2,
                    const BaForwardOpts *opts, float *d_sim, uint32_t *d_status,
                    int32_t *d_ticks, void *stream);

/* Writes the standard per-flight noise the kernel would draw from `seed`
 * ([P, D, Fmax, 4]); for tests and for callers that want to keep the draws. */
BaStatus ba_generate_noise(BaPlan *plan, int32_t n_particles, uint64_t seed,
                           float *d_noise, void *stream);

/* Host-buffer convenience: the same forward + backward through pinned/pageable
 * HOST arrays (copies in, runs, copies out, synchronises).  This is the call a
 * non-PyTorch host (and bench.py's end-to-end leg) makes.  h_g_* may be NULL
 * (forward only). */
BaStatus ba_logjoint_host(BaPlan *plan, int32_t n_particles,
                          const float *h_lat_airport, const float *h_lat_route,
                          const float *h_scales, const float *h_noise,
                          const BaForwardOpts *opts,
                          float *h_logp, uint32_t *h_status,
                          float *h_g_airport, float *h_g_route, float *h_g_scales);

/* d_g_noise[p, d, f, :] = d_dlogp[p, d] * d_value_tape[p, d, f, :]: the cotangent w.r.t. the
 * per-flight site values of a value-mode forward (BaForwardOpts.d_value_tape).  The
 * value-gradient half of the reverse pass; ba_backward gives the latent half. */
BaStatus ba_backward_values(BaPlan *plan, int32_t n_particles, const float *d_dlogp,
                            const float *d_value_tape, float *d_g_noise, void *stream);

/* ---- mean-field guide, fused with the simulation's gradient ------------
 * The guide of scripts/wn/train.py's objective when it is a diagonal Normal over the
 * n = 2N + N^2 log-latents [turnaround N | service N | travel N x N] (train.py:228-235,
 * 254-267): z = loc + exp(log_scale) * eps, latents = exp(z).  Two kernels bracket
 * ba_forward so that nothing of the ELBO's backward pass runs in the host framework:
 *
 * ba_mf_sample    draws eps (Philox: seed, particle_offset + p, latent index), writes the
 *                 kernel-layout latents d_lat_airport [P,2,N] / d_lat_route [P,R], and per
 *                 particle d_lp0 [P] = log prior(latents) - log q(z)  (model.py:58-87 priors;
 *                 train.py:301-316 single_particle_elbo).  d_z [P,n] nullable (diagnostics).
 * ba_mf_backward  contracts the tape of ba_forward with the guide's Jacobian and writes ONE
 *                 flat buffer d_flat [2n + 4] ready for a sum all-reduce:
 *                   [0,n)    d loss / d loc         [n,2n)  d loss / d log_scale
 *                   [2n,2n+3) d loss / d scales     [2n+3]  loss
 *                 with loss = -(1 / (count * flights)) sum_p ELBO_p  (train.py:318-326);
 *                 `inv_count_flights` = 1 / (global particle count * total flights).
 *                 `prior_weight`: share of the per-particle terms (prior, -log q) this call
 *                 accounts for -- 1 when the plan holds every day of the data set; when the
 *                 DAYS are sharded over ranks too (each rank's plan holds a subset of the
 *                 days and the ranks of one particle shard share seed and particle_offset),
 *                 the shares of those ranks must sum to 1.
 *                 d_work: ba_mf_work_bytes() bytes of scratch.
 * d_params [2n] = loc | log_scale.  Only plans without cancellations (C = 2). */
size_t ba_mf_work_bytes(const BaPlan *plan, int32_t n_particles);
BaStatus ba_mf_sample(BaPlan *plan, int32_t n_particles, const float *d_params,
                      uint64_t seed, const uint64_t *d_seed, uint32_t particle_offset,
                      float *d_lat_airport, float *d_lat_route, float *d_lp0, float *d_z,
                      void *stream);
BaStatus ba_mf_backward(BaPlan *plan, int32_t n_particles, const float *d_params,
                        uint64_t seed, const uint64_t *d_seed, uint32_t particle_offset,
                        const void *d_tape, const float *d_logp, const float *d_lp0,
                        float inv_count_flights, float prior_weight, void *d_work, float *d_flat,
                        void *stream);

/* ---- conditional flow guide: monotonic rational-quadratic spline -------------------------
 * The elementwise transform of the CalNF guide (the reference's guide is zuko.flows.NSF,
 * scripts/wn/train.py:622-626; its surface is used by scripts/training.py:184-294).  Each
 * of the M = samples x transformed-features elements has 3*bins-1 unconstrained parameters
 * (d_params [M, 3*bins-1], the conditioner GEMM's output): bin-width logits, bin-height
 * logits, interior knot-slope logits; knots span [-bound, bound], identity outside.
 * ba_rqs_forward  -> d_y [M], d_logdet [M] = log dy/dx.
 * ba_rqs_backward -> d_gx [M], d_gparams [M, 3*bins-1]: the gradient of
 *                    sum(d_gy * y + d_glogdet * logdet) w.r.t. x and the parameters.
 * bins in {4, 8, 12}.  Device pointers, fp32. */
BaStatus ba_rqs_forward(const float *d_x, const float *d_params, float *d_y, float *d_logdet,
                        int64_t n_elements, int32_t bins, float bound, void *stream);
BaStatus ba_rqs_backward(const float *d_x, const float *d_params, const float *d_gy,
                         const float *d_glogdet, float *d_gx, float *d_gparams,
                         int64_t n_elements, int32_t bins, float bound, void *stream);

/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
uint64_t ba_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BAYESAIR_B200_H */
