// MAPPO GAE / return scan and advantage normalisation (src/learners/mappo_gnn_sat_learner.py:504-532).
// HBM-bound streaming kernels: lane = env column (coalesced), the time axis walked in reverse and, when
// the batch alone cannot fill the GPU, split into segments that are scanned as affine maps and composed.
#include "internal.h"

namespace msat {

int g_gae_force_plain = 0;   // msat_debug_option("gae_plain", 1): keep the register-chunked scan (A/B measurements)

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

// ---- software-pipelined plain scan (S = 1) ------------------------------------------------------------
// When the batch alone fills the GPU the register-chunked scan above stops at ~0.78 of the copy bandwidth:
// not for lack of loads in flight but because every warp request covers only 128 contiguous bytes of a
// [T, B] row, so HBM pages are opened for half-page bursts.  Here a lane owns FOUR adjacent env columns:
// every warp instruction moves 512 contiguous bytes (16-byte cp.async / LDS.128 / STG.128; the done row as
// 4-byte copies), the two warps of a CTA cover 1 KB of each row, and a private ring of GP_NS stages x
// GP_CH time steps per warp keeps ~27 KB per warp in flight without holding registers.  Four independent
// recurrences per lane give the instruction-level parallelism that the lower warp count takes away.  No
// cross-warp synchronisation; per column the arithmetic (and its rounding order) is gae_segment's.
int g_gae_variant = 0;                      // msat_tune("gae_variant", 4 | 2 | 1): pin the columns per lane (sweeps)
int g_gae_warps_per_sm = 0;                // segmented scan: split time until this many warps per SM (0 = by batch); msat_tune("gae_warps_per_sm", w)
int g_gae_pipe_min_cols = 24576;            // smallest batch that takes the pipelined scan; msat_tune("gae_pipe_min_cols", B)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void gae_update(float r, float v, uint32_t dn, float gamma, float gl, float& next_value,
                                           float& gae, double& ssum, double& ssq, float& adv, float& tgt) {
    const float nt = dn ? 0.0f : 1.0f;
    const float cc = __fmul_rn(gl, nt);
    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, next_value), nt)), v);
    gae = __fadd_rn(delta, __fmul_rn(cc, gae));
    ssum += (double)gae;
    ssq += (double)gae * (double)gae;
    adv = gae;
    tgt = __fadd_rn(gae, v);                               // learner:526
    next_value = v;
}

// VEC = env columns per lane: 4 (16-byte accesses, 512 contiguous bytes per warp request) for batches of
// ~50k columns and more, 2 for half of that so that enough warps stay in flight (one warp's 512-step scan is
// latency-bound on its own: ~70-85 us whatever the batch, profiles/r2_gae_sweep.txt); smaller batches take
// the time-segmented kernel above.
template <int VEC>
__device__ __forceinline__ void cp_async_vec(void* smem_dst, const void* gmem_src) {
    if (VEC == 4) cp_async16(smem_dst, gmem_src);
    else if (VEC == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
    else cp_async4(smem_dst, gmem_src);
}

template <int GP_CH, int GP_NS, int VEC>
__global__ void __launch_bounds__(32) gae_pipe_kernel(const float* __restrict__ reward, long long rs_t,
                                                      const uint8_t* __restrict__ done,
                                                      const float* __restrict__ value,
                                                      const float* __restrict__ last_val, float gamma, float gl,
                                                      float* __restrict__ adv, float* __restrict__ targets, int T, int B,
                                                      double* __restrict__ stats) {
    extern __shared__ __align__(16) uint8_t gp_smem[];
    constexpr int COLS = 32 * VEC;
    constexpr int GP_STAGE = GP_CH * (COLS * 4 + COLS * 4 + COLS);
    const int lane = threadIdx.x;
    const int b = blockIdx.x * COLS + VEC * lane;          // first of this lane's VEC columns
    uint8_t* ring = gp_smem;
    const bool ok = b < B;                                 // B % 4 == 0 and VEC | 4: all VEC columns or none
    // the done row travels in 4-byte pieces: every lane (VEC 4), every 2nd (VEC 2) or every 4th lane (VEC 1)
    const bool d_lane = (lane % (4 / VEC)) == 0;
    const int nchunks = (T + GP_CH - 1) / GP_CH;

    // chunk c = time steps T-1-c*GP_CH down to T-c*GP_CH-GP_CH; row i of a stage = step t_hi - i
    auto issue = [&](int c) {
        if (c < nchunks && ok) {
            uint8_t* st = ring + (c % GP_NS) * GP_STAGE;
            const int t_hi = T - 1 - c * GP_CH;
#pragma unroll
            for (int i = 0; i < GP_CH; ++i) {
                const int t = t_hi - i;
                if (t >= 0) {
                    cp_async_vec<VEC>(st + i * (COLS * 4) + lane * (4 * VEC), reward + (long long)t * rs_t + b);
                    cp_async_vec<VEC>(st + GP_CH * COLS * 4 + i * (COLS * 4) + lane * (4 * VEC), value + (size_t)t * B + b);
                    if (d_lane) cp_async4(st + 2 * GP_CH * COLS * 4 + i * COLS + lane * VEC, done + (size_t)t * B + b);
                }
            }
        }
        cp_async_commit();                                 // one group per call keeps the group count uniform
    };

#pragma unroll
    for (int c = 0; c < GP_NS - 1; ++c) issue(c);
    float nv[VEC], gae[VEC];
    double ssum[VEC], ssq[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        nv[j] = ok ? __ldg(last_val + b + j) : 0.0f;
        gae[j] = 0.0f;
        ssum[j] = 0.0;
        ssq[j] = 0.0;
    }
    for (int c = 0; c < nchunks; ++c) {
        issue(c + GP_NS - 1);
        cp_async_wait<GP_NS - 1>();                        // chunk c has landed (this lane's copies) ...
        if (VEC < 4) __syncwarp();                         // ... and the neighbour lane's piece of the done row
        const uint8_t* st = ring + (c % GP_NS) * GP_STAGE;
        const int t_hi = T - 1 - c * GP_CH;
        if (ok) {
#pragma unroll
            for (int i = 0; i < GP_CH; ++i) {
                const int t = t_hi - i;
                if (t >= 0) {
                    float r[VEC], v[VEC], a[VEC], g[VEC];
                    uint32_t dn[VEC];
                    const uint8_t* pr = st + i * (COLS * 4) + lane * (4 * VEC);
                    const uint8_t* pv = st + GP_CH * COLS * 4 + i * (COLS * 4) + lane * (4 * VEC);
                    const uint8_t* pd = st + 2 * GP_CH * COLS * 4 + i * COLS + lane * VEC;
                    if (VEC == 4) {
                        const float4 r4 = *reinterpret_cast<const float4*>(pr), v4 = *reinterpret_cast<const float4*>(pv);
                        const uint32_t d4 = *reinterpret_cast<const uint32_t*>(pd);
                        r[0] = r4.x; r[1 % VEC] = r4.y; r[2 % VEC] = r4.z; r[3 % VEC] = r4.w;
                        v[0] = v4.x; v[1 % VEC] = v4.y; v[2 % VEC] = v4.z; v[3 % VEC] = v4.w;
                        dn[0] = d4 & 0xFFu; dn[1 % VEC] = d4 & 0xFF00u; dn[2 % VEC] = d4 & 0xFF0000u; dn[3 % VEC] = d4 & 0xFF000000u;
                    } else if (VEC == 2) {
                        const float2 r2 = *reinterpret_cast<const float2*>(pr), v2 = *reinterpret_cast<const float2*>(pv);
                        const uint32_t d2 = *reinterpret_cast<const uint16_t*>(pd);
                        r[0] = r2.x; r[1 % VEC] = r2.y;
                        v[0] = v2.x; v[1 % VEC] = v2.y;
                        dn[0] = d2 & 0xFFu; dn[1 % VEC] = d2 & 0xFF00u;
                    } else {
                        r[0] = *reinterpret_cast<const float*>(pr);
                        v[0] = *reinterpret_cast<const float*>(pv);
                        dn[0] = *pd;
                    }
#pragma unroll
                    for (int j = 0; j < VEC; ++j)
                        gae_update(r[j], v[j], dn[j], gamma, gl, nv[j], gae[j], ssum[j], ssq[j], a[j], g[j]);
                    float* pa = adv + (size_t)t * B + b;
                    float* pt = targets + (size_t)t * B + b;
                    if (VEC == 4) {
                        __stcs(reinterpret_cast<float4*>(pa), make_float4(a[0], a[1 % VEC], a[2 % VEC], a[3 % VEC]));
                        __stcs(reinterpret_cast<float4*>(pt), make_float4(g[0], g[1 % VEC], g[2 % VEC], g[3 % VEC]));
                    } else if (VEC == 2) {
                        __stcs(reinterpret_cast<float2*>(pa), make_float2(a[0], a[1 % VEC]));
                        __stcs(reinterpret_cast<float2*>(pt), make_float2(g[0], g[1 % VEC]));
                    } else {
                        __stcs(pa, a[0]);
                        __stcs(pt, g[0]);
                    }
                }
            }
        }
        // the r / v pieces of a stage are refilled only by the lane that read them (program order); the done
        // pieces of VEC < 4 are shared between neighbouring lanes
        if (VEC < 4) __syncwarp();
    }
    if (stats) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s1 += ssum[j]; s2 += ssq[j]; }
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            atomicAdd(&stats[1], s1);
            atomicAdd(&stats[2], s2);
            if (blockIdx.x == 0) atomicAdd(&stats[0], (double)T * (double)B);
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

// adv = (adv - mean) / (std + 1e-8), population std over all elements (learner:530-532).
// 8 B per element of pure streaming: 16-byte accesses, four independent chunks per thread in flight,
// grid sized by the element count; the <= 3 elements before the first 16-byte boundary and the tail
// are handled by the first threads of block 0.
constexpr int NORM_UNROLL = 4;
__global__ void __launch_bounds__(256) adv_normalize_kernel(float* __restrict__ adv, long long count,
                                                            const double* __restrict__ stats) {
    const double n = stats[0];
    const double mean = stats[1] / n;
    double var = stats[2] / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float fm = (float)mean;
    const float denom = __fadd_rn((float)sqrt(var), 1e-8f);
    const long long head = min(count, (long long)((16 - (reinterpret_cast<uintptr_t>(adv) & 15)) & 15) / 4);
    const long long n4 = (count - head) / 4;
    float4* a4 = reinterpret_cast<float4*>(adv + head);
    const long long base = ((long long)blockIdx.x * blockDim.x) * NORM_UNROLL + threadIdx.x;
    float4 x[NORM_UNROLL];
#pragma unroll
    for (int u = 0; u < NORM_UNROLL; ++u) {
        const long long i = base + (long long)u * blockDim.x;
        if (i < n4) x[u] = a4[i];
    }
#pragma unroll
    for (int u = 0; u < NORM_UNROLL; ++u) {
        const long long i = base + (long long)u * blockDim.x;
        if (i < n4) {
            float4 y;
            y.x = __fdiv_rn(__fsub_rn(x[u].x, fm), denom);
            y.y = __fdiv_rn(__fsub_rn(x[u].y, fm), denom);
            y.z = __fdiv_rn(__fsub_rn(x[u].z, fm), denom);
            y.w = __fdiv_rn(__fsub_rn(x[u].w, fm), denom);
            a4[i] = y;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
        const long long tail0 = head + 4 * n4;
        const long long i = threadIdx.x < 4 ? (long long)threadIdx.x : tail0 + (threadIdx.x - 4);
        const bool mine = threadIdx.x < 4 ? i < head : i < count;
        if (mine) adv[i] = __fdiv_rn(__fsub_rn(adv[i], fm), denom);
    }
}

template <int CH, int NS, int VEC>
static cudaError_t launch_gae_pipe(const float* reward, long long rs_t, const uint8_t* done, const float* value,
                                   const float* last_val, float gamma, float gl, float* adv, float* targets, int T, int B,
                                   double* stats, cudaStream_t s) {
    constexpr int kSmem = NS * CH * (32 * VEC * 9);
    static_assert(kSmem <= 48 * 1024, "ring must fit the default dynamic shared memory limit");
    gae_pipe_kernel<CH, NS, VEC><<<(B + 32 * VEC - 1) / (32 * VEC), 32, kSmem, s>>>(reward, rs_t, done, value, last_val,
                                                                                    gamma, gl, adv, targets, T, B, stats);
    return cudaGetLastError();
}

cudaError_t launch_gae(const float* reward, long long rs_t, long long rs_b, const uint8_t* done, const float* value,
                       const float* last_val, float gamma, float gl, float* adv, float* targets, int T, int B,
                       double* stats, cudaStream_t s) {
    if (T == 0 || B == 0) return cudaSuccess;
    const int grid = (B + 31) / 32;
    // enough warps to cover HBM latency: >= 8 per SM, segments of at least one load chunk
    int S = 1;
    // measured (profiles/r2_gae_sweep.txt, "segments"): 8 warps per SM up to 8,192 columns, 16 above
    const int wps = g_gae_warps_per_sm > 0 ? g_gae_warps_per_sm : (grid > 256 ? 16 : 8);
    while (S < 32 && grid * S < 148 * wps && T / (2 * S) >= GAE_CHUNK) S *= 2;
    const int seg_len = (T + S - 1) / S;
    // the batch fills the GPU by itself and the rows are 16-byte copyable: pipelined plain scan
    const bool rows16 = rs_b == 1 && (rs_t % 4) == 0 && (B % 4) == 0 &&
                        ((reinterpret_cast<uintptr_t>(reward) | reinterpret_cast<uintptr_t>(value) |
                          reinterpret_cast<uintptr_t>(last_val) | reinterpret_cast<uintptr_t>(adv) |
                          reinterpret_cast<uintptr_t>(targets)) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(done) & 3) == 0;
    // pipelined scan when the batch keeps enough one-warp CTAs in flight at some column vector width
    if (rows16 && !g_gae_force_plain && B >= g_gae_pipe_min_cols) {
        int vec = g_gae_variant;               // msat_tune("gae_variant", 4 | 2 | 1) pins the width (sweeps)
        if (vec != 4 && vec != 2 && vec != 1) vec = B >= 2 * g_gae_pipe_min_cols ? 4 : 2;
#define MSAT_PIPE(V) launch_gae_pipe<8, 4, V>(reward, rs_t, done, value, last_val, gamma, gl, adv, targets, T, B, stats, s)
        return vec == 4 ? MSAT_PIPE(4) : (vec == 2 ? MSAT_PIPE(2) : MSAT_PIPE(1));
#undef MSAT_PIPE
    }
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
    const long long per_block = 256LL * NORM_UNROLL * 4;
    const long long blocks = (count + per_block - 1) / per_block + 1;     // +1: head/tail re-basing can shift by < 4
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    adv_normalize_kernel<<<(int)blocks, 256, 0, s>>>(adv, count, stats);
    return cudaGetLastError();
}

}  // namespace msat
