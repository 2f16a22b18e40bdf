// MAPPO GAE / return scan and advantage normalisation (src/learners/mappo_gnn_sat_learner.py:504-532).
// HBM-bound streaming kernels: one thread per env column, time steps walked in reverse with the next
// chunk of loads issued before the current chunk's dependent chain is evaluated.
#include "internal.h"

namespace msat {

constexpr int GAE_CHUNK = 8;

// for t = T-1..0: nt = 1-done; delta = r + gamma*v_next*nt - v; gae = delta + (gamma*lambda)*nt*gae
// (learner:515-516); adv[t] = gae; targets = adv + value (learner:526).  No FMA contraction so the
// rounding sequence is the reference's mul/add sequence.
__global__ void __launch_bounds__(64) gae_kernel(const float* __restrict__ reward, long long rs_t, long long rs_b,
                                                 const uint8_t* __restrict__ done, const float* __restrict__ value,
                                                 const float* __restrict__ last_val, float gamma, float gl,
                                                 float* __restrict__ adv, float* __restrict__ targets, int T, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float gae = 0.0f;
    float next_value = last_val[b];
    float r[GAE_CHUNK], v[GAE_CHUNK], rn[GAE_CHUNK], vn[GAE_CHUNK];
    uint8_t dn[GAE_CHUNK], dnn[GAE_CHUNK];

    auto load = [&](int t_hi, float* rr, float* vv, uint8_t* dd) {
#pragma unroll
        for (int i = 0; i < GAE_CHUNK; ++i) {
            const int t = t_hi - i;
            if (t >= 0) {
                rr[i] = __ldcs(reward + (long long)t * rs_t + (long long)b * rs_b);
                vv[i] = __ldcs(value + (size_t)t * B + b);
                dd[i] = __ldcs(done + (size_t)t * B + b);
            }
        }
    };
    load(T - 1, r, v, dn);
    for (int t_hi = T - 1; t_hi >= 0; t_hi -= GAE_CHUNK) {
        load(t_hi - GAE_CHUNK, rn, vn, dnn);   // prefetch the next (earlier) chunk
#pragma unroll
        for (int i = 0; i < GAE_CHUNK; ++i) {
            const int t = t_hi - i;
            if (t >= 0) {
                const float nt = dn[i] ? 0.0f : 1.0f;
                const float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(__fmul_rn(gamma, next_value), nt)), v[i]);
                gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nt), gae));
                __stcs(adv + (size_t)t * B + b, gae);
                __stcs(targets + (size_t)t * B + b, __fadd_rn(gae, v[i]));
                next_value = v[i];
            }
        }
#pragma unroll
        for (int i = 0; i < GAE_CHUNK; ++i) { r[i] = rn[i]; v[i] = vn[i]; dn[i] = dnn[i]; }
    }
}

__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ adv, long long count,
                                                        double* __restrict__ stats) {
    double s = 0.0, ss = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const double x = (double)adv[i];
        s += x;
        ss += x * x;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    __shared__ double sh[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; c += sh[1][w]; }
        atomicAdd(&stats[1], a);
        atomicAdd(&stats[2], c);
        if (blockIdx.x == 0) atomicAdd(&stats[0], (double)count);
    }
}

// adv = (adv - mean) / (std + 1e-8), population std over all elements (learner:530-532)
__global__ void __launch_bounds__(256) adv_normalize_kernel(float* __restrict__ adv, long long count,
                                                            const double* __restrict__ stats) {
    const double n = stats[0];
    const double mean = stats[1] / n;
    double var = stats[2] / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float fm = (float)mean;
    const float denom = __fadd_rn((float)sqrt(var), 1e-8f);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        adv[i] = __fdiv_rn(__fsub_rn(adv[i], fm), denom);
}

cudaError_t launch_gae(const float* reward, long long rs_t, long long rs_b, const uint8_t* done, const float* value,
                       const float* last_val, float gamma, float gl, float* adv, float* targets, int T, int B,
                       cudaStream_t s) {
    if (T == 0 || B == 0) return cudaSuccess;
    gae_kernel<<<(B + 63) / 64, 64, 0, s>>>(reward, rs_t, rs_b, done, value, last_val, gamma, gl, adv, targets, T, B);
    return cudaGetLastError();
}
cudaError_t launch_adv_stats(const float* adv, long long count, double* stats, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    long long blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adv_stats_kernel<<<(int)blocks, 256, 0, s>>>(adv, count, stats);
    return cudaGetLastError();
}
cudaError_t launch_adv_normalize(float* adv, long long count, const double* stats, cudaStream_t s) {
    if (count == 0) return cudaSuccess;
    long long blocks = (count + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adv_normalize_kernel<<<(int)blocks, 256, 0, s>>>(adv, count, stats);
    return cudaGetLastError();
}

}  // namespace msat
