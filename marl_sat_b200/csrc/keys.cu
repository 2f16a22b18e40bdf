// Rollout RNG chain on device: JAX 0.4.29 Threefry-2x32 split / randint as used by
// src/learners/mappo_gnn_sat_learner.py:397,416-417,426-434 and src/runners/mappo_runner.py:289-295.
#include "internal.h"

namespace msat {

// rng,act = split(rng); rng,step = split(rng); rng,prob,reset = split(rng,3)   (learner:397,416,426)
__global__ void rng_chain_kernel(const uint32_t* __restrict__ rng_in, uint32_t* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t r[2] = {rng_in[0], rng_in[1]}, a[2], b[2];
    split2(r[0], r[1], a, b);                 // rng <- a, act_key <- b
    const uint32_t act0 = b[0], act1 = b[1];
    r[0] = a[0]; r[1] = a[1];
    split2(r[0], r[1], a, b);                 // rng <- a, step_key <- b
    const uint32_t st0 = b[0], st1 = b[1];
    r[0] = a[0]; r[1] = a[1];
    // split(rng, 3): threefry_2x32(rng, arange(6)).reshape(3, 2)
    uint32_t w[6];
    for (uint32_t i = 0; i < 6; ++i) w[i] = bits32_at(r[0], r[1], 6u, i);
    out[0] = w[0]; out[1] = w[1];             // rng'
    out[2] = act0; out[3] = act1;
    out[4] = st0;  out[5] = st1;
    out[6] = w[2]; out[7] = w[3];             // prob_key
    out[8] = w[4]; out[9] = w[5];             // reset_key
}

__global__ void rng_split2_kernel(const uint32_t* __restrict__ key_in, uint32_t* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t a[2], b[2];
    split2(key_in[0], key_in[1], a, b);
    out[0] = a[0]; out[1] = a[1]; out[2] = b[0]; out[3] = b[1];
}

// problem_idx = randint(prob_key, (Bg,), 0, P)[off + b]; reset_keys = split(reset_key, Bg)[off + b]
// (jax._src.random._randint: two 32-bit draws combined with multiplier (2^16 % span)^2 % span).
__global__ void __launch_bounds__(256) env_keys_kernel(const uint32_t* __restrict__ prob_key,
                                                       const uint32_t* __restrict__ reset_key, uint32_t Bg,
                                                       uint32_t off, uint32_t Bl, uint32_t P,
                                                       int32_t* __restrict__ idx, uint32_t* __restrict__ keys) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= Bl) return;
    const uint32_t g = off + b;
    if (idx) {
        uint32_t k1[2], k2[2];
        split2(prob_key[0], prob_key[1], k1, k2);
        const uint32_t hi = bits32_at(k1[0], k1[1], Bg, g);
        const uint32_t lo = bits32_at(k2[0], k2[1], Bg, g);
        const uint32_t span = P > 0u ? P : 1u;
        uint32_t mult = 65536u % span;
        mult = (mult * mult) % span;
        const uint32_t o = ((hi % span) * mult + (lo % span)) % span;
        idx[b] = (int32_t)o;
    }
    if (keys) {
        keys[2 * b + 0] = bits32_at(reset_key[0], reset_key[1], 2u * Bg, 2u * g);
        keys[2 * b + 1] = bits32_at(reset_key[0], reset_key[1], 2u * Bg, 2u * g + 1u);
    }
}

cudaError_t launch_rng_chain(const uint32_t* rng_in, uint32_t* chain_out, cudaStream_t s) {
    rng_chain_kernel<<<1, 32, 0, s>>>(rng_in, chain_out);
    return cudaGetLastError();
}
cudaError_t launch_rng_split2(const uint32_t* key_in, uint32_t* out, cudaStream_t s) {
    rng_split2_kernel<<<1, 32, 0, s>>>(key_in, out);
    return cudaGetLastError();
}
cudaError_t launch_env_keys(const uint32_t* prob_key, const uint32_t* reset_key, int Bg, int off, int Bl, int P,
                            int32_t* idx, uint32_t* keys, cudaStream_t s) {
    if (Bl == 0) return cudaSuccess;
    env_keys_kernel<<<(Bl + 255) / 256, 256, 0, s>>>(prob_key, reset_key, (uint32_t)Bg, (uint32_t)off, (uint32_t)Bl,
                                                     (uint32_t)P, idx, keys);
    return cudaGetLastError();
}

}  // namespace msat
