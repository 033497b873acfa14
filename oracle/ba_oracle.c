/*
 * ORACLE / TEST INFRASTRUCTURE ONLY -- never linked, imported or called by the
 * product path (bayesair_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * A plain, sequential, list-based C restatement of the hot path of
 * dawsonc/BayesAir: one (particle, day) of the discrete-time simulation inside
 * bayes_air.model.air_traffic_network_model together with the log-probability
 * of every per-flight sample site and the pathwise gradient of that sum with the
 * queue decisions held fixed (what torch autograd gives the reference, since
 * comparisons are not differentiable).  It deliberately keeps the reference's
 * data structures (one global pending list scanned in full every tick, one
 * global in-transit list, per-airport runway / turnaround / available-aircraft
 * lists) instead of the per-airport decomposition the CUDA kernel uses, so the
 * two are independent statements of the same semantics.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_golden.py checks this file
 * against tests/golden/ *.npz, which hold log_prob_sum / per-site-class sums /
 * autograd gradients produced by the UNMODIFIED reference files run in the
 * build container through oracle/reference_runner.py (generator:
 * tests/golden/make_golden.py).  The reference's own unit tests hold no golden
 * vectors for this path.
 *
 * Reference lines restated (relative to /root/reference):
 *   clock + tick order ............ bayes_air/model.py:158-200
 *   initial aircraft .............. bayes_air/model.py:148-155
 *   reserve injection ............. bayes_air/model.py:170-174
 *   turnaround release ............ bayes_air/types/airport.py:66-84
 *   pop-ready / cancellation ...... bayes_air/network.py:43-134
 *   runway service loop ........... bayes_air/types/airport.py:86-128
 *   service-time assignment ....... bayes_air/types/airport.py:130-158
 *   departure / arrival sites ..... bayes_air/types/airport.py:164-204
 *   turnaround site ............... bayes_air/types/airport.py:210-221
 *   travel-time site .............. bayes_air/network.py:136-168
 *   in-transit -> arrival queue ... bayes_air/network.py:178-198
 *   completion test ............... bayes_air/network.py:34-41
 * Third-party arithmetic (not under /root/reference): torch.distributions
 * {Normal, Exponential, RelaxedBernoulli}.log_prob / rsample (torch 2.1.2 pinned by
 * the reference, formulas read from the installed torch 2.11) and
 * pyro-ppl 1.9.0 Trace.log_prob_sum -- see SURVEY.md App. A.3.
 *
 * All model arithmetic that feeds a decision is IEEE fp32 in the reference's
 * operation order (compile with -ffp-contract=off); log-prob terms are formed
 * in fp32 like torch does and accumulated in double.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BA_ORACLE_OK 0
#define BA_ORACLE_TICK_CAP 1
#define BA_ORACLE_BAD_ARG 2
#define BA_ORACLE_INTERNAL 3

typedef struct {
    int n_airports;              /* N: airports of this day, in the day's dict order      */
    int n_flights;               /* F: flights, already in stable sched-dep order           */
    const float *sched_dep;      /* [F]                                                     */
    const float *act_dep;        /* [F]  NaN = unobserved (obs=None)                        */
    const float *act_arr;        /* [F]  NaN = unobserved                                   */
    const int8_t *cancel_obs;    /* [F]  0 / 1, -1 = unobserved                             */
    const int32_t *origin;       /* [F]  day-local airport index                            */
    const int32_t *dest;         /* [F]                                                     */
    const int32_t *airport_global; /* [N] day-local -> index into the latent arrays         */
} BaOracleDay;

typedef struct {
    int n_global;                /* Ng: airports in states[0] order (latent indexing)       */
    const float *mean_turnaround;   /* [Ng]                                                 */
    const float *mean_service;      /* [Ng]                                                 */
    const float *travel;            /* [Ng*Ng] row = origin                                 */
    const float *log_init_aircraft; /* [Ng] or NULL (no-cancellation mode)                  */
    const float *base_cancel_logprob; /* [Ng] or NULL                                       */
    float scales[3];             /* runway_use_time_std_dev, travel_time_variation,
                                    turnaround_time_variation (constrained values)           */
} BaOracleLatents;

typedef struct {
    double logp;                 /* sum of all per-flight site log-probs of this day        */
    double class_sums[7];        /* dep_service, travel, arr_service, turnaround,
                                    cancelled, sim_departure, sim_arrival                   */
    int n_ticks;
    int status;
} BaOracleResult;

/* gradient buffers (double, caller-zeroed, accumulated into) */
typedef struct {
    double *g_turn;     /* [Ng]    */
    double *g_serv;     /* [Ng]    */
    double *g_travel;   /* [Ng*Ng] */
    double *g_logn0;    /* [Ng] or NULL */
    double *g_basecl;   /* [Ng] or NULL */
    double *g_scales;   /* [3]     */
} BaOracleGrads;

/* ------------------------------------------------------------------------- */
/* torch.distributions arithmetic in fp32                                    */
/* ------------------------------------------------------------------------- */
static const float kLogSqrt2Pi = 0.918938533204672741780329736406f;
static const float kEps = 1.1920928955078125e-07f; /* torch.finfo(float32).eps  */
static const float kTiny = 1.17549435082228750797e-38f; /* finfo.tiny          */

/* Normal(loc, scale).log_prob(v)  (torch/distributions/normal.py::log_prob) */
static float normal_logpdf(float v, float loc, float scale)
{
    float var = scale * scale;
    float d = v - loc;
    return -(d * d) / (2.0f * var) - logf(scale) - kLogSqrt2Pi;
}

/* Exponential(rate).log_prob(v) */
static float expo_logpdf(float v, float rate) { return logf(rate) - rate * v; }

static float softplus_f(float x) /* F.softplus, beta=1, threshold=20 */
{
    return x > 20.0f ? x : log1pf(expf(x));
}

static float sigmoid_f(float z) { return 1.0f / (1.0f + expf(-z)); }

/* RelaxedBernoulli(temperature=0.5, probs=p).log_prob(v) and d/dp of it.
 * relaxed_bernoulli.py::LogitRelaxedBernoulli.log_prob + SigmoidTransform. */
static float relaxed_bernoulli_logpdf(float p, float v, float *dlp_dp)
{
    const float temp = 0.5f;
    float q = p < kEps ? kEps : (p > 1.0f - kEps ? 1.0f - kEps : p);
    float logits = logf(q) - log1pf(-q);
    float y = v < kTiny ? kTiny : (v > 1.0f - kEps ? 1.0f - kEps : v);
    float x = logf(y) - log1pf(-y);
    float diff = logits - x * temp;
    float lp = logf(temp) + diff - 2.0f * log1pf(expf(diff));
    lp = lp + softplus_f(-x) + softplus_f(x);
    if (dlp_dp) {
        /* clamp passes gradient on the closed interval [eps, 1-eps] */
        float pass = (p >= kEps && p <= 1.0f - kEps) ? 1.0f : 0.0f;
        float dlogits = 1.0f / q + 1.0f / (1.0f - q);
        float ed = expf(diff);
        float sig = isinf(ed) ? 1.0f : ed / (1.0f + ed);
        *dlp_dp = pass * dlogits * (1.0f - 2.0f * sig);
    }
    return lp;
}

/* ------------------------------------------------------------------------- */
/* simulation state                                                          */
/* ------------------------------------------------------------------------- */
typedef struct {
    int flight;
    float queue_start;
    float total_wait;
    int assigned;          /* assigned_service_time is not None */
    float assigned_time;
    /* where queue_start came from, for the adjoint of max(aircraft, sched): */
    float src_eps;         /* turnaround noise of the landing that freed the aircraft */
    float src_weight;      /* 1 aircraft branch, 0.5 tie, 0 schedule / constant       */
} QEntry;

typedef struct { float time; float eps; int is_clock_alias; int has_src; } Aircraft;

typedef struct {
    QEntry *q; int q_head, q_tail;             /* runway_queue (cap = entries/day)  */
    Aircraft *turn; int n_turn;                /* turnaround_queue                  */
    Aircraft *avail; int a_head, a_tail;       /* available_aircraft (FIFO)         */
    float n_avail;                             /* num_available_aircraft (real)     */
    float mean_service, mean_turn, turn_std, base_cancel;
    int g;                                     /* global latent index               */
} AirportSim;

typedef struct { int flight; float arrival_time; } Transit;

static float flight_noise(const float *noise, int f, int col) { return noise[4 * f + col]; }

/* RelaxedBernoulliStraightThrough(T=0.5, probs=p).rsample() driven by the uniform u
 * (torch relaxed_bernoulli.py::LogitRelaxedBernoulli.rsample + SigmoidTransform._call,
 * then Pyro's straight-through rounding).  Returns the hard 0/1 value, *soft = the
 * relaxed value the site's log_prob is evaluated at. */
static float relaxed_bernoulli_draw(float p, float u, float *soft)
{
    const float temp = 0.5f;
    float q = p < kEps ? kEps : (p > 1.0f - kEps ? 1.0f - kEps : p);
    float uu = u < kEps ? kEps : (u > 1.0f - kEps ? 1.0f - kEps : u);
    float x = (logf(uu) - log1pf(-uu) + logf(q) - log1pf(-q)) / temp;
    float y = sigmoid_f(x);
    y = y < kTiny ? kTiny : (y > 1.0f - kEps ? 1.0f - kEps : y);   /* SigmoidTransform clamp */
    y = y < kEps ? kEps : (y > 1.0f - kEps ? 1.0f - kEps : y);     /* clamp_probs            */
    *soft = y;
    return rintf(y);                                                /* torch.round: half-even */
}

int ba_oracle_run_day(const BaOracleDay *day, const BaOracleLatents *lat,
                      const float *noise,       /* [F,4] e_dep, eps_travel, e_arr, eps_turn */
                      const float *obs_noise,   /* [F,4] u_cancel, eps_dep_obs, eps_arr_obs, -
                                                   for flights with obs=None; nullable      */
                      int noise_is_value,       /* 1: columns are the site VALUES x,tau,x,u */
                      float delta_t, float max_t, int include_cancellations,
                      int max_ticks,
                      BaOracleResult *res, BaOracleGrads *grads,
                      float *sim_out,           /* [F,3] cancelled, sim_dep, sim_arr; nullable */
                      int32_t *pop_tick         /* [F,2] dep pop tick, arr pop tick; nullable */)
{
    const int N = day->n_airports, F = day->n_flights, Ng = lat->n_global;
    const float sigma_r = lat->scales[0], v_t = lat->scales[1], v_u = lat->scales[2];
    int status = BA_ORACLE_OK;
    memset(res, 0, sizeof(*res));
    if (N <= 0 || F < 0 || !noise) { res->status = BA_ORACLE_BAD_ARG; return BA_ORACLE_BAD_ARG; }

    /* per-airport capacities from the schedule */
    int *n_dep = calloc(N, sizeof(int)), *n_arr = calloc(N, sizeof(int));
    for (int f = 0; f < F; ++f) { n_dep[day->origin[f]]++; n_arr[day->dest[f]]++; }

    AirportSim *ap = calloc(N, sizeof(AirportSim));
    int reserve_cap = max_ticks + 8;
    for (int a = 0; a < N; ++a) {
        AirportSim *A = &ap[a];
        A->g = day->airport_global[a];
        A->q = malloc(sizeof(QEntry) * (size_t)(n_dep[a] + n_arr[a] + 1));
        A->turn = malloc(sizeof(Aircraft) * (size_t)(n_arr[a] + 1));
        /* model.py:139-146 */
        A->mean_service = lat->mean_service[A->g];
        A->mean_turn = lat->mean_turnaround[A->g];
        A->turn_std = v_u * A->mean_turn;
        float n0;
        if (include_cancellations) {
            n0 = expf(lat->log_init_aircraft[A->g]);          /* model.py:91-99  */
            A->base_cancel = expf(lat->base_cancel_logprob[A->g]); /* :103-111     */
        } else {
            n0 = 1000.0f;                                      /* model.py:116-121 */
            A->base_cancel = 0.0f;
        }
        A->n_avail = n0;
        /* model.py:152-155: while i < num_available_aircraft: append(0.0) */
        int n_init = 0;
        while ((float)n_init < n0) n_init++;
        A->avail = malloc(sizeof(Aircraft) * (size_t)(n_init + n_arr[a] + reserve_cap));
        for (int i = 0; i < n_init; ++i) {
            A->avail[i].time = 0.0f; A->avail[i].eps = 0.0f;
            A->avail[i].is_clock_alias = 0; A->avail[i].has_src = 0;
        }
        A->a_head = 0; A->a_tail = n_init;
    }

    int *pending = malloc(sizeof(int) * (size_t)(F + 1)), n_pending = F;
    int *new_pending = malloc(sizeof(int) * (size_t)(F + 1));
    int *ready = malloc(sizeof(int) * (size_t)(F + 1));
    int *bucket_order = malloc(sizeof(int) * (size_t)(N + 1));
    char *bucket_seen = malloc((size_t)N);
    Transit *transit = malloc(sizeof(Transit) * (size_t)(F + 1));
    int n_transit = 0;
    int8_t *sim_cancelled = malloc((size_t)F + 1);   /* -1 = not yet scored */
    float *sim_dep = malloc(sizeof(float) * (size_t)(F + 1));
    float *sim_arr = malloc(sizeof(float) * (size_t)(F + 1));
    for (int f = 0; f < F; ++f) {
        pending[f] = f; sim_cancelled[f] = -1; sim_dep[f] = NAN; sim_arr[f] = NAN;
        if (pop_tick) { pop_tick[2 * f] = -1; pop_tick[2 * f + 1] = -1; }
    }

    double lp_class[7] = {0, 0, 0, 0, 0, 0, 0};
    double g_sigma = 0, g_vt = 0, g_vu = 0;

    float t = 0.0f;
    int tick = 0;
    for (;;) {
        /* NetworkState.complete, network.py:34-41 (turnaround queues ignored) */
        int complete = (n_pending == 0 && n_transit == 0);
        for (int a = 0; a < N && complete; ++a)
            if (ap[a].q_head != ap[a].q_tail) complete = 0;
        if (complete) break;
        if (tick >= max_ticks) { status = BA_ORACLE_TICK_CAP; break; }

        t = t + delta_t;   /* model.py:161, in-place fp32 add */
        tick++;

        /* 1. airport.update_available_aircraft(t), airport.py:66-84 */
        for (int a = 0; a < N; ++a) {
            AirportSim *A = &ap[a];
            int w = 0;
            for (int i = 0; i < A->n_turn; ++i) {
                if (A->turn[i].time <= t) {
                    A->avail[A->a_tail++] = A->turn[i];
                    A->n_avail = A->n_avail + 1.0f;
                } else {
                    A->turn[w++] = A->turn[i];
                }
            }
            A->n_turn = w;
        }

        /* 2. reserve injection, model.py:170-174.  The appended object is the clock
         * tensor itself, mutated in place afterwards: its value is read at pop time. */
        if (t >= max_t) {
            for (int a = 0; a < N; ++a) {
                AirportSim *A = &ap[a];
                A->n_avail = A->n_avail + 1.0f;
                Aircraft r = {0.0f, 0.0f, 1, 0};
                A->avail[A->a_tail++] = r;
            }
        }

        /* 3. pop_ready_to_depart_flights, network.py:57-134 */
        int n_ready = 0, n_new_pending = 0, n_buckets = 0;
        memset(bucket_seen, 0, (size_t)N);
        for (int i = 0; i < n_pending; ++i) {
            int f = pending[i];
            if (day->sched_dep[f] <= t) {
                int o = day->origin[f];
                if (!bucket_seen[o]) { bucket_seen[o] = 1; bucket_order[n_buckets++] = o; }
                ready[n_ready++] = f;
            } else {
                new_pending[n_new_pending++] = f;
            }
        }
        for (int b = 0; b < n_buckets; ++b) {
            int o = bucket_order[b];
            AirportSim *A = &ap[o];
            int n_to_depart = 0;
            for (int i = 0; i < n_ready; ++i) if (day->origin[ready[i]] == o) n_to_depart++;
            /* network.py:80-84 (fp32, same operation order) */
            float n_at_start = A->n_avail;
            float ratio = n_at_start / (float)n_to_depart;
            float z = 10.0f * (ratio - 0.25f);
            float sg = sigmoid_f(z);
            float p_cancel = 1.0f - (1.0f - A->base_cancel) * sg;
            for (int i = 0; i < n_ready; ++i) {
                int f = ready[i];
                if (day->origin[f] != o) continue;
                if (sim_cancelled[f] < 0) {               /* first time seen, :100-109 */
                    float v, v_hard;
                    if (day->cancel_obs[f] < 0) {             /* obs=None: sampled */
                        if (!obs_noise || grads) { status = BA_ORACLE_BAD_ARG; goto done; }
                        v_hard = relaxed_bernoulli_draw(p_cancel, obs_noise[4 * f], &v);
                    } else {
                        v = v_hard = (float)day->cancel_obs[f];
                    }
                    float dlp_dp = 0.0f;
                    float lpc = relaxed_bernoulli_logpdf(p_cancel, v, &dlp_dp);
                    lp_class[4] += (double)lpc;
                    sim_cancelled[f] = v_hard != 0.0f ? 1 : 0;
                    if (grads && include_cancellations) {
                        /* p = 1 - (1-b) sigma(z), z = 10 (n/nr - 1/4),
                         * n = exp(l) + integers, b = exp(c) */
                        double dp_db = (double)sg;                 /* dp/db = sigma(z) */
                        double dp_dz = -(1.0 - (double)A->base_cancel) * sg * (1.0 - sg);
                        double dz_dn = 10.0 / (double)n_to_depart;
                        double n0 = exp((double)lat->log_init_aircraft[A->g]);
                        grads->g_basecl[A->g] += dlp_dp * dp_db * (double)A->base_cancel;
                        grads->g_logn0[A->g] += dlp_dp * dp_dz * dz_dn * n0;
                    }
                }
                if (sim_cancelled[f] == 0) {
                    if (A->n_avail <= 0.0f) {
                        new_pending[n_new_pending++] = f;  /* delayed: END of pending */
                    } else {
                        A->n_avail = A->n_avail - 1.0f;
                        if (A->a_head == A->a_tail) { status = BA_ORACLE_INTERNAL; goto done; }
                        Aircraft ac = A->avail[A->a_head++];           /* pop(0) */
                        float ac_time = ac.is_clock_alias ? t : ac.time;
                        float sd = day->sched_dep[f];
                        float ready_time = ac_time > sd ? ac_time : sd; /* torch.maximum */
                        /* 4. model.py:181-183 */
                        QEntry e;
                        e.flight = f; e.queue_start = ready_time; e.total_wait = 0.0f;
                        e.assigned = 0; e.assigned_time = 0.0f;
                        e.src_eps = ac.eps;
                        e.src_weight = !ac.has_src ? 0.0f
                                       : (ac_time > sd ? 1.0f : (ac_time == sd ? 0.5f : 0.0f));
                        A->q[A->q_tail++] = e;
                    }
                }   /* else: cancelled -> completed, never queued (:127-129) */
            }
        }
        { int *tmp = pending; pending = new_pending; new_pending = tmp; n_pending = n_new_pending; }

        /* 5. update_runway_queue for every airport in day order, airport.py:101-128 */
        for (int a = 0; a < N; ++a) {
            AirportSim *A = &ap[a];
            float rate = 1.0f / A->mean_service;     /* airport.py:146 */
            while (A->q_head != A->q_tail &&
                   (!A->q[A->q_head].assigned || A->q[A->q_head].assigned_time <= t)) {
                QEntry *h = &A->q[A->q_head];
                int f = h->flight;
                int departing = (day->origin[f] == a);   /* airport.py:115,140 */
                if (!h->assigned) {
                    /* _assign_service_time, airport.py:130-158 */
                    int col = departing ? 0 : 2;
                    float x = noise_is_value ? flight_noise(noise, f, col)
                                             : flight_noise(noise, f, col) / rate; /* e / rate */
                    lp_class[col] += (double)expo_logpdf(x, rate);
                    for (int j = A->q_head; j < A->q_tail; ++j)
                        A->q[j].total_wait = A->q[j].total_wait + x;
                    h->assigned_time = h->queue_start + h->total_wait;
                    h->assigned = 1;
                    if (grads) {
                        /* value mode: d/dm [log(1/m) - x/m] ; noise mode: x = e*m, -1/m only */
                        double m = A->mean_service;
                        grads->g_serv[A->g] += noise_is_value ? (-1.0 / m + (double)x / (m * m))
                                                              : (-1.0 / m);
                    }
                }
                if (h->assigned_time <= t) {
                    QEntry e = *h;
                    A->q_head++;
                    float mu = e.queue_start + e.total_wait;
                    if (departing) {
                        /* _assign_departure_time, airport.py:164-182 */
                        float y = day->act_dep[f];
                        if (isnan(y)) {                        /* obs=None: Normal.rsample */
                            if (!obs_noise || grads) { status = BA_ORACLE_BAD_ARG; goto done; }
                            y = mu + obs_noise[4 * f + 1] * sigma_r;
                        }
                        lp_class[5] += (double)normal_logpdf(y, mu, sigma_r);
                        sim_dep[f] = y;
                        if (pop_tick) pop_tick[2 * f] = tick;
                        /* add_in_transit_flights, network.py:136-168 */
                        float T = lat->travel[(size_t)A->g * Ng + ap[day->dest[f]].g];
                        float sc = T * v_t;
                        float tau = noise_is_value ? flight_noise(noise, f, 1)
                                                   : T + flight_noise(noise, f, 1) * sc;
                        lp_class[1] += (double)normal_logpdf(tau, T, sc);
                        transit[n_transit].flight = f;
                        transit[n_transit].arrival_time = sim_dep[f] + tau;
                        n_transit++;
                        if (grads) {
                            double r = ((double)y - mu) / ((double)sigma_r * sigma_r);
                            double dy = (double)y - mu;
                            g_sigma += dy * dy / pow((double)sigma_r, 3) - 1.0 / sigma_r;
                            size_t ti = (size_t)A->g * Ng + ap[day->dest[f]].g;
                            if (!noise_is_value) {
                                grads->g_serv[A->g] += r * (double)e.total_wait / A->mean_service;
                                grads->g_travel[ti] += -1.0 / T;
                                g_vt += -1.0 / v_t;
                                if (e.src_weight > 0.0f) {
                                    /* ready = max(act_arr_g + u_g, sched), u = m~ (1 + v_u eps) */
                                    grads->g_turn[A->g] += e.src_weight * r * (1.0 + (double)v_u * e.src_eps);
                                    g_vu += e.src_weight * r * (double)e.src_eps * A->mean_turn;
                                }
                            } else {
                                double zt = ((double)tau - T) / sc;
                                grads->g_travel[ti] += zt / sc + (zt * zt - 1.0) / T;
                                g_vt += (zt * zt - 1.0) / v_t;
                            }
                        }
                    } else {
                        /* _assign_arrival_time, airport.py:188-204 */
                        float y = day->act_arr[f];
                        if (isnan(y)) {
                            if (!obs_noise || grads) { status = BA_ORACLE_BAD_ARG; goto done; }
                            y = mu + obs_noise[4 * f + 2] * sigma_r;
                        }
                        lp_class[6] += (double)normal_logpdf(y, mu, sigma_r);
                        sim_arr[f] = y;
                        if (pop_tick) pop_tick[2 * f + 1] = tick;
                        /* _assign_turnaround_time, airport.py:210-221 */
                        float u = noise_is_value ? flight_noise(noise, f, 3)
                                                 : A->mean_turn + flight_noise(noise, f, 3) * A->turn_std;
                        lp_class[3] += (double)normal_logpdf(u, A->mean_turn, A->turn_std);
                        Aircraft ac;
                        ac.time = sim_arr[f] + u;
                        ac.eps = noise_is_value ? 0.0f : flight_noise(noise, f, 3);
                        ac.is_clock_alias = 0;
                        ac.has_src = noise_is_value ? 0 : 1;
                        A->turn[A->n_turn++] = ac;
                        if (grads) {
                            double r = ((double)y - mu) / ((double)sigma_r * sigma_r);
                            double dy = (double)y - mu;
                            g_sigma += dy * dy / pow((double)sigma_r, 3) - 1.0 / sigma_r;
                            int o = day->origin[f];
                            size_t ti = (size_t)ap[o].g * Ng + A->g;
                            float T = lat->travel[ti];
                            if (!noise_is_value) {
                                double eps_t = flight_noise(noise, f, 1);
                                grads->g_serv[A->g] += r * (double)e.total_wait / A->mean_service;
                                grads->g_travel[ti] += r * (1.0 + (double)v_t * eps_t);
                                g_vt += r * eps_t * T;
                                grads->g_turn[A->g] += -1.0 / A->mean_turn;
                                g_vu += -1.0 / v_u;
                            } else {
                                double zu = ((double)u - A->mean_turn) / A->turn_std;
                                grads->g_turn[A->g] += zu / A->turn_std + (zu * zu - 1.0) / A->mean_turn;
                                g_vu += (zu * zu - 1.0) / v_u;
                            }
                        }
                    }
                }
            }
        }

        /* 6. update_in_transit_flights, network.py:178-198 (list order preserved) */
        {
            int w = 0;
            for (int i = 0; i < n_transit; ++i) {
                if (transit[i].arrival_time <= t) {
                    AirportSim *A = &ap[day->dest[transit[i].flight]];
                    QEntry e;
                    e.flight = transit[i].flight; e.queue_start = transit[i].arrival_time;
                    e.total_wait = 0.0f; e.assigned = 0; e.assigned_time = 0.0f;
                    e.src_eps = 0.0f; e.src_weight = 0.0f;
                    A->q[A->q_tail++] = e;
                } else {
                    transit[w++] = transit[i];
                }
            }
            n_transit = w;
        }
    }

done:
    res->n_ticks = tick;
    res->status = status;
    res->logp = 0.0;
    for (int c = 0; c < 7; ++c) { res->class_sums[c] = lp_class[c]; res->logp += lp_class[c]; }
    if (grads) {
        grads->g_scales[0] += g_sigma; grads->g_scales[1] += g_vt; grads->g_scales[2] += g_vu;
    }
    if (sim_out)
        for (int f = 0; f < F; ++f) {
            sim_out[3 * f] = (float)sim_cancelled[f];
            sim_out[3 * f + 1] = sim_dep[f];
            sim_out[3 * f + 2] = sim_arr[f];
        }
    for (int a = 0; a < N; ++a) { free(ap[a].q); free(ap[a].turn); free(ap[a].avail); }
    free(ap); free(n_dep); free(n_arr); free(pending); free(new_pending); free(ready);
    free(bucket_order); free(bucket_seen); free(transit); free(sim_cancelled);
    free(sim_dep); free(sim_arr);
    return status;
}

/* ------------------------------------------------------------------------- */
/* batch driver: P particles x D days over a pthread pool                    */
/* (the reference's only parallelism is one process per seed,                */
/*  scripts/wn/experiments.sh:3-13; particle-days are independent)           */
/* ------------------------------------------------------------------------- */
typedef struct {
    const BaOracleDay *days; int D; int P; int Ng; int Fmax;
    const float *mean_turnaround, *mean_service, *travel, *log_init, *base_cl; /* [P,...] */
    const float *scales; const float *noise; /* [P,D,Fmax,4] */
    const float *obs_noise; float *sim; /* [P,D,Fmax,4] nullable; [P,D,Fmax,3] nullable */
    int noise_is_value; float delta_t, max_t; int include_cancellations, max_ticks;
    int want_grad;
    double *logp;      /* [P,D] */
    int32_t *status;   /* [P,D] */
    int32_t *n_ticks;  /* [P,D] */
    double *g_turn, *g_serv, *g_travel, *g_logn0, *g_basecl, *g_scales; /* [P,...], day-summed */
    int next; pthread_mutex_t mu;
} Batch;

static void *batch_worker(void *arg)
{
    Batch *b = (Batch *)arg;
    for (;;) {
        pthread_mutex_lock(&b->mu);
        int p = b->next++;
        pthread_mutex_unlock(&b->mu);
        if (p >= b->P) break;
        size_t Ng = (size_t)b->Ng;
        BaOracleLatents lat;
        lat.n_global = b->Ng;
        lat.mean_turnaround = b->mean_turnaround + p * Ng;
        lat.mean_service = b->mean_service + p * Ng;
        lat.travel = b->travel + p * Ng * Ng;
        lat.log_init_aircraft = b->log_init ? b->log_init + p * Ng : NULL;
        lat.base_cancel_logprob = b->base_cl ? b->base_cl + p * Ng : NULL;
        memcpy(lat.scales, b->scales, sizeof(lat.scales));
        BaOracleGrads g, *gp = NULL;
        if (b->want_grad) {
            g.g_turn = b->g_turn + p * Ng; g.g_serv = b->g_serv + p * Ng;
            g.g_travel = b->g_travel + p * Ng * Ng;
            g.g_logn0 = b->g_logn0 ? b->g_logn0 + p * Ng : NULL;
            g.g_basecl = b->g_basecl ? b->g_basecl + p * Ng : NULL;
            g.g_scales = b->g_scales + (size_t)p * 3;
            gp = &g;
        }
        for (int d = 0; d < b->D; ++d) {
            BaOracleResult r;
            const float *nz = b->noise + ((size_t)p * b->D + d) * (size_t)b->Fmax * 4;
            const size_t row = ((size_t)p * b->D + d) * (size_t)b->Fmax;
            const float *on = b->obs_noise ? b->obs_noise + row * 4 : NULL;
            float *so = b->sim ? b->sim + row * 3 : NULL;
            ba_oracle_run_day(&b->days[d], &lat, nz, on, b->noise_is_value, b->delta_t, b->max_t,
                              b->include_cancellations, b->max_ticks, &r, gp, so, NULL);
            b->logp[(size_t)p * b->D + d] = r.logp;
            b->status[(size_t)p * b->D + d] = r.status;
            b->n_ticks[(size_t)p * b->D + d] = r.n_ticks;
        }
    }
    return NULL;
}

int ba_oracle_run_batch(const BaOracleDay *days, int D, int P, int Ng, int Fmax,
                        const float *mean_turnaround, const float *mean_service,
                        const float *travel, const float *log_init, const float *base_cl,
                        const float *scales, const float *noise, const float *obs_noise,
                        float *sim, int noise_is_value,
                        float delta_t, float max_t, int include_cancellations, int max_ticks,
                        int want_grad, int n_threads,
                        double *logp, int32_t *status, int32_t *n_ticks,
                        double *g_turn, double *g_serv, double *g_travel,
                        double *g_logn0, double *g_basecl, double *g_scales)
{
    Batch b;
    memset(&b, 0, sizeof(b));
    b.days = days; b.D = D; b.P = P; b.Ng = Ng; b.Fmax = Fmax;
    b.mean_turnaround = mean_turnaround; b.mean_service = mean_service; b.travel = travel;
    b.log_init = log_init; b.base_cl = base_cl; b.scales = scales; b.noise = noise;
    b.obs_noise = obs_noise; b.sim = sim;
    b.noise_is_value = noise_is_value; b.delta_t = delta_t; b.max_t = max_t;
    b.include_cancellations = include_cancellations; b.max_ticks = max_ticks;
    b.want_grad = want_grad; b.logp = logp; b.status = status; b.n_ticks = n_ticks;
    b.g_turn = g_turn; b.g_serv = g_serv; b.g_travel = g_travel;
    b.g_logn0 = g_logn0; b.g_basecl = g_basecl; b.g_scales = g_scales;
    pthread_mutex_init(&b.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, batch_worker, &b);
    for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
    pthread_mutex_destroy(&b.mu);
    return 0;
}
