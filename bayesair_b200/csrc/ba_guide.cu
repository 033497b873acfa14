// Mean-field guide fused with the simulation's gradient (SURVEY.md §8e "fusion point").
//
// The guide of the reference's WN objective, when it is a diagonal Normal over the
// n = 2N + N^2 log-latents [turnaround N | service N | travel N x N]
// (scripts/wn/train.py:228-235,254-267), has an elementwise backward pass:
//     d loss / d loc_j       = s * sum_p gz[p, j]
//     d loss / d log_scale_j = s * sum_p (gz[p, j] * eps[p, j] * sigma_j + 1)
// with gz = d (log prior + simulation log-joint) / d z and s = -1 / (particles * flights)
// (train.py:301-326: ELBO = model log-prob - guide log-prob, loss = -mean ELBO / flights).
// Two kernels bracket ba_forward so that none of this runs as framework autograd:
//
//   ba_mf_sample_kernel    eps by Philox4x32-10 (counter: latent index / 4, particle; key:
//                          seed), z = loc + sigma eps, kernel-layout latents exp(z), and per
//                          particle log prior(exp z) - log q(z)  (model.py:58-87 priors)
//   ba_mf_backward_kernel  regenerates eps, gathers the particle-day gradient rows the forward
//                          kernel wrote (the tape), applies the chain rule of exp and the
//                          prior, and reduces over a chunk of particles in a fixed order;
//   ba_mf_reduce_kernel    sums the chunks (fixed order: deterministic) into the flat buffer
//                          [d loc | d log_scale | d scales (3) | loss] that goes straight
//                          into ONE sum all-reduce.
//
// HBM-bound streaming kernels (the tape is read once, coalesced along the latent index).
#include "ba_kernels.cuh"
#include "ba_sim.cuh"

#define BA_MF_TAG 0x4D463031u         // "MF01": Philox counter word of the guide's draws
#define BA_LOG_SQRT_2PI_D 0.918938533204672741780329736406

// four standard normals for latents 4*j4 .. 4*j4+3 of particle `particle`
__device__ __forceinline__ void ba_mf_eps4(uint64_t seed, uint32_t particle, uint32_t j4, float e[4])
{
    uint32_t r[4];
    ba_philox(j4, particle, BA_MF_TAG, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const float u0 = ba_u01(r[0]), u1 = ba_u01(r[1]), u2 = ba_u01(r[2]), u3 = ba_u01(r[3]);
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincosf(6.283185307179586f * u1, &s0, &c0);
    sincosf(6.283185307179586f * u3, &s1, &c1);
    e[0] = ra * c0; e[1] = ra * s0; e[2] = rb * c1; e[3] = rb * s1;
}

// Gamma(a, b) prior of the latent x = exp(z) (model.py:62-64,71-73,80-82): log-density at x
// and its derivative w.r.t. z; the diagonal of the travel matrix has no site
__device__ __forceinline__ void ba_mf_prior(int j, int N, float z, float x, float *lp, float *dlp_dz)
{
    float a, b, lnorm;
    if (j < N) { a = 1.0f; b = 2.0f; lnorm = 0.6931471805599453f; }                 // a ln b - lgamma(a)
    else if (j < 2 * N) { a = 1.5f; b = 10.0f; lnorm = 3.4538776394910684f + 0.12078223763524522f; }
    else {
        const int od = j - 2 * N;
        if (od / N == od % N) { *lp = 0.0f; *dlp_dz = 0.0f; return; }
        a = 4.0f; b = 1.25f; lnorm = 0.8925742052568390f - 1.791759469228055f;
    }
    *lp = lnorm + (a - 1.0f) * z - b * x;
    *dlp_dz = (a - 1.0f) - b * x;
}

__global__ void __launch_bounds__(256) ba_mf_sample_kernel(const BaMfParams prm)
{
    const int p = blockIdx.x;
    const int N = prm.N, n = prm.n;
    const uint64_t seed = prm.seed_dev ? *prm.seed_dev : prm.seed;
    const float *loc = prm.params, *lsc = prm.params + n;
    double acc = 0.0;
    for (int j4 = threadIdx.x; 4 * j4 < n; j4 += blockDim.x) {
        float e[4];
        ba_mf_eps4(seed, prm.particle_offset + (uint32_t)p, (uint32_t)j4, e);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = 4 * j4 + i;
            if (j >= n) break;
            const float ls = lsc[j];
            const float z = loc[j] + expf(ls) * e[i];
            const float x = expf(z);
            if (prm.z) prm.z[(size_t)p * n + j] = z;
            if (j < 2 * N) prm.lat_airport[(size_t)p * 2 * N + j] = x;
            else {
                const int r = prm.route_of[j - 2 * N];
                if (r >= 0) prm.lat_route[(size_t)p * prm.R + r] = x;
            }
            float lp, dl;
            ba_mf_prior(j, N, z, x, &lp, &dl);
            // - log q(z) = 0.5 eps^2 + log sigma + log sqrt(2 pi)
            acc += (double)lp + (double)(0.5f * e[i] * e[i] + ls) + BA_LOG_SQRT_2PI_D;
        }
    }
    // block reduction, fixed order
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        prm.lp0[p] = (float)s;
    }
}

// grid (ceil(n / 4 / 256), chunks): thread = four consecutive latents, block row = a chunk of
// particles walked in order
__global__ void __launch_bounds__(256) ba_mf_backward_kernel(const BaMfParams prm)
{
    const int N = prm.N, n = prm.n, D = prm.D, stride = prm.stride;
    const uint64_t seed = prm.seed_dev ? *prm.seed_dev : prm.seed;
    const int j4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunk = blockIdx.y;
    const int per = (prm.P + prm.chunks - 1) / prm.chunks;
    const int p0 = chunk * per, p1 = min(prm.P, p0 + per);
    float *out = prm.work + (size_t)chunk * (2 * n + 4);
    if (4 * j4 < n) {
        int col[4];
        float lc[4], sg[4], a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = 4 * j4 + i;
            col[i] = -1; lc[i] = 0.0f; sg[i] = 0.0f;
            if (j < n) {
                lc[i] = prm.params[j]; sg[i] = expf(prm.params[n + j]);
                if (j < 2 * N) col[i] = j;
                else { const int r = prm.route_of[j - 2 * N]; col[i] = r >= 0 ? 2 * N + r : -1; }
            }
        }
        for (int p = p0; p < p1; ++p) {
            float e[4];
            ba_mf_eps4(seed, prm.particle_offset + (uint32_t)p, (uint32_t)j4, e);
            const float *tp = prm.tape + (size_t)p * D * stride;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = 4 * j4 + i;
                if (j >= n) break;
                const float z = lc[i] + sg[i] * e[i];
                const float x = expf(z);
                float g = 0.0f;
                if (col[i] >= 0)
                    for (int d = 0; d < D; ++d) g += __ldg(tp + (size_t)d * stride + col[i]);
                float lp, dl;
                ba_mf_prior(j, N, z, x, &lp, &dl);
                const float gz = x * g + prm.prior_weight * dl;   // chain rule of exp + the prior
                a1[i] += gz;
                a2[i] += gz * e[i] * sg[i] + prm.prior_weight;    // + 1: d(-log q)/d log_scale
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = 4 * j4 + i;
            if (j < n) { out[j] = a1[i]; out[n + j] = a2[i]; }
        }
    }
    // the three scale gradients and the ELBO sum of this chunk: first block column only
    if (blockIdx.x == 0) {
        double acc[4] = {0, 0, 0, 0};
        for (int i = p0 * D + threadIdx.x; i < p1 * D; i += blockDim.x) {
            const float *row = prm.tape + (size_t)i * stride + 2 * N + prm.R;
            acc[0] += row[0]; acc[1] += row[1]; acc[2] += row[2];
            acc[3] += prm.logp[i];
        }
        for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) acc[3] += (double)prm.prior_weight * prm.lp0[p];
        __shared__ double red[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_down_sync(0xffffffffu, acc[q], o);
            if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = acc[q];
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
            out[2 * n + threadIdx.x] = (float)s;
        }
    }
}

__global__ void __launch_bounds__(256) ba_mf_reduce_kernel(const BaMfParams prm)
{
    const int m = 2 * prm.n + 4;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    float s = 0.0f;
    for (int c = 0; c < prm.chunks; ++c) s += prm.work[(size_t)c * m + j];
    prm.flat[j] = -prm.inv_count_flights * s;         // loss = -mean ELBO / flights
}

int ba_mf_chunks(int P)
{
    int c = (P + 31) / 32;
    if (c > 64) c = 64;
    if (c < 1) c = 1;
    return c;
}

cudaError_t ba_launch_mf_sample(const BaMfParams &p, cudaStream_t s)
{
    ba_mf_sample_kernel<<<p.P, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t ba_launch_mf_backward(const BaMfParams &p, cudaStream_t s)
{
    const int n4 = (p.n + 3) / 4;
    dim3 grid((unsigned)((n4 + 255) / 256), (unsigned)p.chunks);
    ba_mf_backward_kernel<<<grid, 256, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int m = 2 * p.n + 4;
    ba_mf_reduce_kernel<<<(m + 255) / 256, 256, 0, s>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Monotonic rational-quadratic spline of the conditional flow guide (bayesair_b200/flows.py;
// the reference's guide is zuko.flows.NSF, scripts/wn/train.py:622-626 -- same spline family,
// Durkan et al. 2019).  Elementwise over M = samples x transformed features; each element has
// its own NP = 3 BINS - 1 unconstrained parameters produced by the conditioner GEMM:
//     widths  = MIN + (1 - MIN BINS) softmax(p[0:BINS])        (x knots, x 2 bound - bound)
//     heights = MIN + (1 - MIN BINS) softmax(p[BINS:2 BINS])   (y knots)
//     slopes  = MIN + softplus(p[2 BINS:] + c), 1 at both ends (identity tails)
// ba_rqs_forward_kernel:  y, log dy/dx.   ba_rqs_backward_kernel: d/dx and d/dparams of
// (gy . y + gld . logdet) by reverse differentiation of the same expressions, one pass.
// HBM-bound streaming kernels: a block of 256 elements moves its 256 x NP parameter floats
// through shared memory with coalesced 16-byte accesses (the rows are 92 bytes for BINS = 8);
// threads then read / write their row at an odd stride (bank-conflict free).
// ---------------------------------------------------------------------------------------
#define BA_RQS_MIN_BIN 1e-3f
#define BA_RQS_MIN_SLOPE 1e-3f
#define BA_RQS_SOFTPLUS_C 0.53974242f      // log(expm1(1 - 1e-3))
#define BA_RQS_BLOCK 256

template <int BINS>
struct BaRqsBin {
    int k;
    float x0, wk, y0, hk, d0, d1;
    float sw[BINS], sh[BINS];            // softmax of the width / height logits
};

template <int BINS>
__device__ __forceinline__ void ba_rqs_knots(const float *p, float x, float bound, BaRqsBin<BINS> &b)
{
    float mw = p[0], mh = p[BINS];
#pragma unroll
    for (int i = 1; i < BINS; ++i) { mw = fmaxf(mw, p[i]); mh = fmaxf(mh, p[BINS + i]); }
    float zw = 0.0f, zh = 0.0f;
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
        b.sw[i] = __expf(p[i] - mw); zw += b.sw[i];
        b.sh[i] = __expf(p[BINS + i] - mh); zh += b.sh[i];
    }
    const float izw = 1.0f / zw, izh = 1.0f / zh, span = 2.0f * bound, sc = 1.0f - BA_RQS_MIN_BIN * BINS;
    float cx = 0.0f, cy = 0.0f;
    b.k = BINS - 1; b.x0 = 0.0f; b.y0 = 0.0f;
    bool found = false;
#pragma unroll
    for (int i = 0; i < BINS; ++i) {
        b.sw[i] *= izw; b.sh[i] *= izh;
        const float w = BA_RQS_MIN_BIN + sc * b.sw[i], h = BA_RQS_MIN_BIN + sc * b.sh[i];
        const float x1 = (i == BINS - 1) ? bound : (cx + w) * span - bound;
        if (!found && (x < x1 || i == BINS - 1)) {
            found = true; b.k = i;
            b.x0 = cx * span - bound; b.y0 = cy * span - bound;
            b.wk = x1 - b.x0;
            b.hk = ((i == BINS - 1) ? bound : (cy + h) * span - bound) - b.y0;
        }
        cx += w; cy += h;
    }
    const int k = b.k;
    // slopes at the two knots of the bin (dynamic index into the staged row: shared memory)
    const float pd0 = k > 0 ? p[2 * BINS + k - 1] : 0.0f, pd1 = k < BINS - 1 ? p[2 * BINS + k] : 0.0f;
    const float a0 = pd0 + BA_RQS_SOFTPLUS_C, a1 = pd1 + BA_RQS_SOFTPLUS_C;
    b.d0 = k > 0 ? BA_RQS_MIN_SLOPE + (a0 > 20.0f ? a0 : log1pf(__expf(a0))) : 1.0f;
    b.d1 = k < BINS - 1 ? BA_RQS_MIN_SLOPE + (a1 > 20.0f ? a1 : log1pf(__expf(a1))) : 1.0f;
}

template <int BINS>
__device__ __forceinline__ void ba_rqs_stage_in(float *sm, const float *params, long long base, long long M)
{
    constexpr int NP = 3 * BINS - 1;
    const long long n = min((long long)BA_RQS_BLOCK, M - base) * NP;
    const float *src = params + base * NP;          // base is a multiple of 256: 16-byte aligned
    const int n4 = (int)(n >> 2);
    for (int i = threadIdx.x; i < n4; i += BA_RQS_BLOCK)
        reinterpret_cast<float4 *>(sm)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
    for (int i = 4 * n4 + threadIdx.x; i < n; i += BA_RQS_BLOCK) sm[i] = __ldg(src + i);
    __syncthreads();
}

template <int BINS>
__global__ void __launch_bounds__(BA_RQS_BLOCK) ba_rqs_forward_kernel(
    const float *__restrict__ x, const float *__restrict__ params, float *__restrict__ y,
    float *__restrict__ ld, long long M, float bound)
{
    constexpr int NP = 3 * BINS - 1;
    __shared__ __align__(16) float sm[BA_RQS_BLOCK * NP];
    const long long base = (long long)blockIdx.x * BA_RQS_BLOCK;
    ba_rqs_stage_in<BINS>(sm, params, base, M);
    const long long i = base + threadIdx.x;
    if (i >= M) return;
    const float xv = x[i];
    if (!(xv > -bound && xv < bound)) { y[i] = xv; ld[i] = 0.0f; return; }
    BaRqsBin<BINS> b;
    ba_rqs_knots<BINS>(sm + threadIdx.x * NP, xv, bound, b);
    const float th = (xv - b.x0) / b.wk, s = b.hk / b.wk, t1 = th * (1.0f - th);
    const float den = s + (b.d1 + b.d0 - 2.0f * s) * t1;
    y[i] = b.y0 + b.hk * (s * th * th + b.d0 * t1) / den;
    const float nd = b.d1 * th * th + 2.0f * s * t1 + b.d0 * (1.0f - th) * (1.0f - th);
    ld[i] = 2.0f * __logf(s) + __logf(nd) - 2.0f * __logf(den);
}

template <int BINS>
__global__ void __launch_bounds__(BA_RQS_BLOCK) ba_rqs_backward_kernel(
    const float *__restrict__ x, const float *__restrict__ params, const float *__restrict__ gy,
    const float *__restrict__ gld, float *__restrict__ gx, float *__restrict__ gparams,
    long long M, float bound)
{
    constexpr int NP = 3 * BINS - 1;
    __shared__ __align__(16) float sm[BA_RQS_BLOCK * NP];
    const long long base = (long long)blockIdx.x * BA_RQS_BLOCK;
    ba_rqs_stage_in<BINS>(sm, params, base, M);
    const long long i = base + threadIdx.x;
    float *row = sm + threadIdx.x * NP;
    if (i < M) {
        const float xv = x[i], gyv = gy[i], glv = gld[i];
        if (!(xv > -bound && xv < bound)) {
            gx[i] = gyv;
#pragma unroll
            for (int j = 0; j < NP; ++j) row[j] = 0.0f;
        } else {
            BaRqsBin<BINS> b;
            ba_rqs_knots<BINS>(row, xv, bound, b);
            const int k = b.k;
            const float pd0 = k > 0 ? row[2 * BINS + k - 1] : 0.0f, pd1 = k < BINS - 1 ? row[2 * BINS + k] : 0.0f;
            const float iw = 1.0f / b.wk;
            const float th = (xv - b.x0) * iw, s = b.hk * iw, t1 = th * (1.0f - th), omt = 1.0f - th;
            const float e = b.d1 + b.d0 - 2.0f * s;
            const float den = s + e * t1, iden = 1.0f / den;
            const float inner = s * th * th + b.d0 * t1, numy = b.hk * inner;
            const float nd = b.d1 * th * th + 2.0f * s * t1 + b.d0 * omt * omt;
            // reverse sweep
            const float g_numy = gyv * iden;
            const float g_den = -gyv * numy * iden * iden - 2.0f * glv * iden;
            const float g_nd = glv / nd;
            float g_s = 2.0f * glv / s + g_nd * 2.0f * t1 + g_numy * b.hk * th * th + g_den;
            float g_d1 = g_nd * th * th;
            float g_d0 = g_nd * omt * omt + g_numy * b.hk * t1;
            float g_t1 = g_nd * 2.0f * s + g_numy * b.hk * b.d0 + g_den * e;
            float g_th = g_nd * (2.0f * b.d1 * th - 2.0f * b.d0 * omt) + g_numy * b.hk * s * 2.0f * th;
            float g_hk = g_numy * inner;
            const float g_e = g_den * t1;
            g_d1 += g_e; g_d0 += g_e; g_s -= 2.0f * g_e;
            g_th += g_t1 * (1.0f - 2.0f * th);
            g_hk += g_s * iw;
            const float g_wk = -g_s * s * iw - g_th * th * iw;
            const float g_x = g_th * iw, g_x0 = -g_x, g_y0 = gyv;
            gx[i] = g_x;
            // knots -> widths / heights -> logits
            const float span = 2.0f * bound, sc = 1.0f - BA_RQS_MIN_BIN * BINS;
            float dotw = 0.0f, doth = 0.0f;
#pragma unroll
            for (int j = 0; j < BINS; ++j) {
                const float gw = j < k ? g_x0 : (j == k ? g_wk : 0.0f);
                const float gh = j < k ? g_y0 : (j == k ? g_hk : 0.0f);
                dotw += gw * b.sw[j]; doth += gh * b.sh[j];
            }
#pragma unroll
            for (int j = 0; j < BINS; ++j) {
                const float gw = j < k ? g_x0 : (j == k ? g_wk : 0.0f);
                const float gh = j < k ? g_y0 : (j == k ? g_hk : 0.0f);
                row[j] = span * sc * b.sw[j] * (gw - dotw);
                row[BINS + j] = span * sc * b.sh[j] * (gh - doth);
            }
#pragma unroll
            for (int j = 0; j < BINS - 1; ++j) row[2 * BINS + j] = 0.0f;
            if (k > 0) row[2 * BINS + k - 1] = g_d0 / (1.0f + __expf(-(pd0 + BA_RQS_SOFTPLUS_C)));
            if (k < BINS - 1) row[2 * BINS + k] = g_d1 / (1.0f + __expf(-(pd1 + BA_RQS_SOFTPLUS_C)));
        }
    }
    __syncthreads();
    const long long n = min((long long)BA_RQS_BLOCK, M - base) * NP;
    float *dst = gparams + base * NP;
    const int n4 = (int)(n >> 2);
    for (int j = threadIdx.x; j < n4; j += BA_RQS_BLOCK)
        reinterpret_cast<float4 *>(dst)[j] = reinterpret_cast<const float4 *>(sm)[j];
    for (int j = 4 * n4 + threadIdx.x; j < n; j += BA_RQS_BLOCK) dst[j] = sm[j];
}

cudaError_t ba_launch_rqs_forward(const float *x, const float *params, float *y, float *ld,
                                  long long M, int bins, float bound, cudaStream_t s)
{
    const unsigned grid = (unsigned)((M + BA_RQS_BLOCK - 1) / BA_RQS_BLOCK);
    switch (bins) {
    case 4: ba_rqs_forward_kernel<4><<<grid, BA_RQS_BLOCK, 0, s>>>(x, params, y, ld, M, bound); break;
    case 8: ba_rqs_forward_kernel<8><<<grid, BA_RQS_BLOCK, 0, s>>>(x, params, y, ld, M, bound); break;
    case 12: ba_rqs_forward_kernel<12><<<grid, BA_RQS_BLOCK, 0, s>>>(x, params, y, ld, M, bound); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t ba_launch_rqs_backward(const float *x, const float *params, const float *gy,
                                   const float *gld, float *gx, float *gparams, long long M,
                                   int bins, float bound, cudaStream_t s)
{
    const unsigned grid = (unsigned)((M + BA_RQS_BLOCK - 1) / BA_RQS_BLOCK);
    switch (bins) {
    case 4: ba_rqs_backward_kernel<4><<<grid, BA_RQS_BLOCK, 0, s>>>(x, params, gy, gld, gx, gparams, M, bound); break;
    case 8: ba_rqs_backward_kernel<8><<<grid, BA_RQS_BLOCK, 0, s>>>(x, params, gy, gld, gx, gparams, M, bound); break;
    case 12: ba_rqs_backward_kernel<12><<<grid, BA_RQS_BLOCK, 0, s>>>(x, params, gy, gld, gx, gparams, M, bound); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
