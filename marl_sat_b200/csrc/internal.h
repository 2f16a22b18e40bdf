// Private declarations shared by the translation units of libmarlsat_b200.so.
#pragma once
#include "common.cuh"

struct msat_plan {
    msat::Dims d;
    int group_threads;     // GS: threads cooperating on one env (32/64/128/256)
    int group_smem_bytes;  // shared memory per env group (multiple of 128)
    int smem_bytes;        // dynamic shared memory per 256-thread CTA
    int group_threads_noobs;   // GS used when a launch writes no observations (little per-env work: small groups)
    int smem_bytes_noobs;
    int compile_smem_bytes;
};

namespace msat {

constexpr int kCtaThreads = 256;

// Shared-memory carve-up of one env group (byte offsets from the group base).
struct GroupLayout {
    int rec, st, satw, x, smx, bar, misc, total;
};
__host__ __device__ inline GroupLayout group_layout(const Dims& d) {
    GroupLayout L;
    int o = 0;
    L.rec = o;  o += d.rec_bytes;                 // bank record (TMA destination, 128-byte aligned)
    L.st = o;   o += 4 * d.state_words;           // state record (multiple of 16 bytes)
    L.satw = o; o += 4 * d.sw;
    L.x = o;    o += 4 * d.xw;
    o = (o + 7) & ~7;
    L.smx = o;  o += 8 * (d.fw + 2);              // {mask word, value word} per 32 observation ints
    L.bar = o;  o += 8;                           // mbarrier
    L.misc = o; o += 16;                          // [0] #unsatisfied accumulator, [1] new problem idx, [2..3] reset key
    L.total = (o + 127) & ~127;
    return L;
}

struct EnvArgs {
    const uint8_t* bank;
    int P;
    const uint32_t* state_in;
    uint32_t* state_out;
    const int32_t* actions;
    const int32_t* prob_idx;
    const uint32_t* keys;
    const uint32_t* rng_in;     // fused key derivation (msat_rollout_step): rollout rng before this step
    uint32_t* chain_out;        // ... and the advanced chain {rng', act, step, prob, reset}
    uint32_t Bg, env_off;       // global batch size and this shard's first global env index
    int auto_reset;
    int32_t* obs;
    int32_t* gnn_assign;        // optional GNN-input outputs of the state the caller continues from
    float* gnn_cf;              // (learner:165-195): assignment int32[B,n], clause_features float[B,m,3]
    float* reward;
    int reward_cols;
    uint8_t* done;
    int done_cols;
    uint8_t* solved;
    int32_t* num_unsat;
    int32_t* episode_step;
    int B;
};

enum EnvMode { MODE_RESET = 0, MODE_STEP = 1, MODE_OBS = 2 };

struct ExportArgs {
    const uint8_t* bank;
    int P;
    const uint32_t* state;
    int B;
    int32_t* assign;
    uint8_t* sat;
    int32_t* num_unsat;
    int32_t* step;
    uint8_t* done;
    int32_t* clauses;
    int32_t* acm;
    int32_t* anm;
    int32_t* l2a;
    int32_t* pidx;
};

cudaError_t launch_compile_bank(const msat_plan* plan, const int32_t* clauses, int P, uint8_t* bank, cudaStream_t s);
cudaError_t launch_env(const msat_plan* plan, EnvMode mode, const EnvArgs& a, cudaStream_t s);
cudaError_t launch_export(const msat_plan* plan, const ExportArgs& a, cudaStream_t s);
cudaError_t launch_rng_chain(const uint32_t* rng_in, uint32_t* chain_out, cudaStream_t s);
cudaError_t launch_rng_split2(const uint32_t* key_in, uint32_t* out, cudaStream_t s);
cudaError_t launch_env_keys(const uint32_t* prob_key, const uint32_t* reset_key, int Bg, int off, int Bl, int P,
                            int32_t* idx, uint32_t* keys, cudaStream_t s);
cudaError_t launch_gae(const float* reward, long long rs_t, long long rs_b, const uint8_t* done, const float* value,
                       const float* last_val, float gamma, float gamma_lambda, float* adv, float* targets, int T, int B,
                       double* stats, cudaStream_t s);
cudaError_t launch_adv_stats(const float* adv, long long count, double* stats, cudaStream_t s);
cudaError_t launch_adv_normalize(float* adv, long long count, const double* stats, cudaStream_t s);
cudaError_t launch_gnn_static(const msat_plan* plan, const uint8_t* bank, int P, float* svf, float* a_pos, float* a_neg,
                              cudaStream_t s);
cudaError_t launch_gnn_dynamic(const msat_plan* plan, const uint8_t* bank, int P, const uint32_t* state, int B,
                               int32_t* assign, float* cf, cudaStream_t s);
cudaError_t launch_rollout_metrics(const float* reward, long long rs_t, long long rs_b, const uint8_t* done,
                                   const uint8_t* solved, const int32_t* num_unsat, const int32_t* episode_step, int T,
                                   int B, double* sums, cudaStream_t s);
cudaError_t launch_flip_gains(const msat_plan* plan, const uint8_t* bank, int P, const uint32_t* state, int B, float tau,
                              int32_t* delta, int32_t* labels, cudaStream_t s);
cudaError_t launch_eval_track(const msat_plan* plan, const uint32_t* state, const uint8_t* solved, int t, int B,
                              uint8_t* ever, int32_t* steps, int32_t* solution, cudaStream_t s);

}  // namespace msat
