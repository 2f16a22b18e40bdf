// MAPPO GAE / return scan and advantage normalisation (src/learners/mappo_gnn_sat_learner.py:504-532).
// HBM-bound streaming kernels: lane = env column (coalesced), the time axis walked in reverse and, when
// the batch alone cannot fill the GPU, split into segments that are scanned as affine maps and composed.
#include "internal.h"

namespace msat {

constexpr int GAE_CHUNK = 8;      // loads in flight per array and thread in the segmented scan
constexpr int GAE_CHUNK_1 = 16;   // ... and in the plain (S = 1) scan, which has fewer warps per SM

// Reverse recurrence over one time segment [t0, t1) of one env column (learner:515-516):
//   nt = 1-done; delta = r + gamma*v_next*nt - v; gae = delta + (gamma*lambda)*nt*gae.
// No FMA contraction, so within a segment the rounding sequence is the reference's mul/add sequence.
// WRITE = false: only the segment's affine map gae_out = bsum + aprod * gae_in is accumulated.
template <bool WRITE, int CH>
__device__ __forceinline__ void gae_segment(const float* __restrict__ reward, long long rs_t, long long rs_b,
                                            const uint8_t* __restrict__ done, const float* __restrict__ value,
                                            float gamma, float gl, float* __restrict__ adv,
                                            float* __restrict__ targets, int B, int b, int t0, int t1,
                                            float next_value, float& gae, float& aprod, double& ssum,
                                            double& ssq) {
    for (int t_hi = t1 - 1; t_hi >= t0; t_hi -= CH) {
        float r[CH], v[CH];
        uint8_t dn[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {          // all loads of the chunk first (memory-level parallelism)
            const int t = t_hi - i;
            if (t >= t0) {
                r[i] = __ldg(reward + (long long)t * rs_t + (long long)b * rs_b);
                v[i] = __ldg(value + (size_t)t * B + b);
                dn[i] = __ldg(done + (size_t)t * B + b);
            }
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int t = t_hi - i;
            if (t >= t0) {
                const float nt = dn[i] ? 0.0f : 1.0f;
                const float c = __fmul_rn(gl, nt);
                const float delta = __fsub_rn(__fadd_rn(r[i], __fmul_rn(__fmul_rn(gamma, next_value), nt)), v[i]);
                gae = __fadd_rn(delta, __fmul_rn(c, gae));
                if (WRITE) {
                    ssum += (double)gae;
                    ssq += (double)gae * (double)gae;
                    __stcs(adv + (size_t)t * B + b, gae);
                    __stcs(targets + (size_t)t * B + b, __fadd_rn(gae, v[i]));   // learner:526
                } else {
                    aprod = __fmul_rn(aprod, c);
                }
                next_value = v[i];
            }
        }
    }
}

// One CTA = 32 env columns (lane = column, coalesced) x S time segments (one warp each).  With S > 1
// every warp first reduces its segment to an affine map (inputs read once from HBM), the maps are
// composed back to front in shared memory, then each warp replays its segment from its true
// incoming gae (inputs now L2/L1 hits) and writes advantages and targets.  S = 1 is the plain scan.
template <int S>
__global__ void __launch_bounds__(32 * S) gae_kernel(const float* __restrict__ reward, long long rs_t,
                                                     long long rs_b, const uint8_t* __restrict__ done,
                                                     const float* __restrict__ value,
                                                     const float* __restrict__ last_val, float gamma, float gl,
                                                     float* __restrict__ adv, float* __restrict__ targets, int T,
                                                     int B, int seg_len, double* __restrict__ stats) {
    __shared__ float sA[S][32], sB[S][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + lane;
    const bool ok = b < B;
    const int t0 = min(T, w * seg_len), t1 = min(T, (w + 1) * seg_len);
    float next_value = 0.0f;
    if (ok && t0 < t1) next_value = (t1 < T) ? __ldg(value + (size_t)t1 * B + b) : __ldg(last_val + b);
    float gae_in = 0.0f;
    double ssum = 0.0, ssq = 0.0;      // fused advantage statistics (learner:530-531)
    if (S > 1) {
        float bsum = 0.0f, aprod = 1.0f;
        if (ok && t0 < t1)
            gae_segment<false, GAE_CHUNK>(reward, rs_t, rs_b, done, value, gamma, gl, adv, targets, B, b, t0, t1, next_value, bsum,
                                          aprod, ssum, ssq);
        sA[w][lane] = aprod;
        sB[w][lane] = bsum;
        __syncthreads();
        for (int s2 = S - 1; s2 > w; --s2) gae_in = __fadd_rn(sB[s2][lane], __fmul_rn(sA[s2][lane], gae_in));
    }
    float dummy = 1.0f;
    if (ok && t0 < t1)
        gae_segment<true, (S == 1 ? GAE_CHUNK_1 : GAE_CHUNK)>(reward, rs_t, rs_b, done, value, gamma, gl, adv, targets, B,
                                                              b, t0, t1, next_value, gae_in, dummy, ssum, ssq);
    if (stats) {       // one atomic pair per warp; the element count is added once
        for (int o = 16; o > 0; o >>= 1) {
            ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
            ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
        }
        if (lane == 0) {
            atomicAdd(&stats[1], ssum);
            atomicAdd(&stats[2], ssq);
            if (blockIdx.x == 0 && w == 0) atomicAdd(&stats[0], (double)T * (double)B);
        }
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
                       double* stats, cudaStream_t s) {
    if (T == 0 || B == 0) return cudaSuccess;
    const int grid = (B + 31) / 32;
    // enough warps to cover HBM latency: >= 8 per SM, segments of at least one load chunk
    int S = 1;
    while (S < 32 && grid * S < 148 * 8 && T / (2 * S) >= GAE_CHUNK) S *= 2;
    const int seg_len = (T + S - 1) / S;
#define MSAT_GAE_LAUNCH(SS)                                                                                         \
    gae_kernel<SS><<<grid, 32 * SS, 0, s>>>(reward, rs_t, rs_b, done, value, last_val, gamma, gl, adv, targets, T, B, \
                                            seg_len, stats)
    switch (S) {
        case 1: MSAT_GAE_LAUNCH(1); break;
        case 2: MSAT_GAE_LAUNCH(2); break;
        case 4: MSAT_GAE_LAUNCH(4); break;
        case 8: MSAT_GAE_LAUNCH(8); break;
        case 16: MSAT_GAE_LAUNCH(16); break;
        default: MSAT_GAE_LAUNCH(32); break;
    }
#undef MSAT_GAE_LAUNCH
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
