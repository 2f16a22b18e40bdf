// extern "C" boundary of libmarlsat_b200.so (include/marl_sat_b200.h): plan construction, argument
// validation and kernel enqueue.  No allocation, no synchronisation (except msat_step_host, which is
// documented to synchronise), no torch types.
#include <math.h>
#include <mutex>
#include <new>

#include "../../include/marl_sat_b200.h"
#include "internal.h"

using namespace msat;

namespace {

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? MSAT_OK : (int)e; }
constexpr int kMaxSmem = 227 * 1024;

// Two internal streams + fork/join events used by msat_rollout_step_host to overlap PCIe copies with the
// kernel; created lazily for the current device (one process drives one GPU in this design).
struct HostPipe {
    std::mutex mu;
    int dev = -1;
    cudaStream_t ws[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    cudaError_t init() {
        int cur = 0;
        cudaError_t e = cudaGetDevice(&cur);
        if (e != cudaSuccess || cur == dev) return e;
        for (int w = 0; w < 2 && e == cudaSuccess; ++w) {
            e = cudaStreamCreateWithFlags(&ws[w], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join[w], cudaEventDisableTiming);
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e == cudaSuccess) dev = cur;
        return e;
    }
};
HostPipe g_pipe;

}  // namespace

extern "C" {

const char* msat_version(void) { return "marl_sat_b200 0.1 (sm_100a)"; }

int32_t msat_num_agents_for(int32_t n, int32_t vars_per_agent) {
    if (n <= 0) return MSAT_EINVAL;
    if (vars_per_agent > 0) return (n + vars_per_agent - 1) / vars_per_agent;   // env:299
    if (n % 4 == 0) return n / 4;                                               // env:315-323
    int r = (int)floor(sqrt((double)n));
    while ((long long)r * r > n) --r;
    while ((long long)(r + 1) * (r + 1) <= n) ++r;
    return r > 2 ? r : 2;                                                       // env:326
}

int msat_plan_create(msat_plan** out, int32_t n, int32_t m, int32_t k, int32_t A, int32_t action_mode,
                     int32_t max_steps, int32_t group_threads) {
    if (!out) return MSAT_EINVAL;
    *out = nullptr;
    if (n <= 0 || m <= 0 || k <= 0 || A <= 0 || A > n || n > 32767) return MSAT_EINVAL;
    if (action_mode != 0 && action_mode != 1) return MSAT_EINVAL;
    if ((long long)A * (2LL * n + m) > (1LL << 26)) return MSAT_EUNSUPPORTED;
    msat_plan* p = new (std::nothrow) msat_plan();
    if (!p) return MSAT_EINVAL;
    Dims& d = p->d;
    d.n = n; d.m = m; d.k = k; d.A = A;
    d.base = n / A; d.rem = n % A;
    d.V = d.base + (d.rem > 0 ? 1 : 0);
    d.D = 2 * n + m;
    d.AD = A * d.D;
    d.action_mode = action_mode;
    d.max_steps = max_steps;
    d.aw = (n + 31) / 32;
    d.sw = (m + 31) / 32;
    d.xw = (d.D + 31) / 32 + 1;
    d.fw = (d.AD + 31) / 32;
    d.agw = (A + 31) / 32;
    d.inv_D = (uint32_t)(((1ULL << 32) + (uint64_t)d.D - 1) / (uint64_t)d.D);
    d.lits_bytes = (m * k * 2 + 15) & ~15;
    d.rec_bytes = (d.lits_bytes + 4 * (d.fw + 1) + 127) & ~127;
    d.state_words = (d.aw + 4 + 3) & ~3;

    const int chunks = (d.AD + 3) / 4;
    int gs = group_threads;
    // measured on B200 (profiles/r1_group_size_sweep.md): about 16-24 store iterations per thread is the sweet
    // spot -- larger groups idle most lanes in the short per-env phases, smaller ones serialise the stores
    if (gs == 0) gs = chunks <= 768 ? 32 : (chunks <= 1536 ? 64 : (chunks <= 3072 ? 128 : 256));
    if (gs != 32 && gs != 64 && gs != 128 && gs != 256) { delete p; return MSAT_EINVAL; }
    const GroupLayout L = group_layout(d);
    // grow the group until one CTA's groups fit in shared memory
    while (gs < 256 && (long long)L.total * (kCtaThreads / gs) > kMaxSmem) gs *= 2;
    if ((long long)L.total * (kCtaThreads / gs) > kMaxSmem) { delete p; return MSAT_EUNSUPPORTED; }
    p->group_threads = gs;
    p->group_smem_bytes = L.total;
    p->smem_bytes = L.total * (kCtaThreads / gs);
    // launches that write no observations (emit_obs off, GNN-input mode) have ~m clause evaluations of work per
    // env: one warp per env unless the caller pinned the group size or eight groups do not fit in shared memory
    int gn = group_threads ? gs : 32;
    while (gn < gs && (long long)L.total * (kCtaThreads / gn) > kMaxSmem) gn *= 2;
    p->group_threads_noobs = gn;
    p->smem_bytes_noobs = L.total * (kCtaThreads / gn);
    p->compile_smem_bytes = 4 * (m + n) * d.agw;
    if (p->compile_smem_bytes > kMaxSmem) { delete p; return MSAT_EUNSUPPORTED; }
    *out = p;
    return MSAT_OK;
}

void msat_plan_destroy(msat_plan* plan) { delete plan; }

int msat_plan_dims(const msat_plan* plan, msat_dims* o) {
    if (!plan || !o) return MSAT_EINVAL;
    const Dims& d = plan->d;
    o->n = d.n; o->m = d.m; o->k = d.k; o->A = d.A; o->V = d.V; o->D = d.D;
    o->action_mode = d.action_mode; o->max_steps = d.max_steps;
    o->rec_bytes = d.rec_bytes; o->state_words = d.state_words;
    o->group_threads = plan->group_threads; o->smem_bytes = plan->smem_bytes;
    return MSAT_OK;
}

int msat_compile_bank(const msat_plan* plan, const int32_t* clauses, int32_t P, void* bank, void* stream) {
    if (!plan || P < 0 || (P > 0 && (!clauses || !bank))) return MSAT_EINVAL;
    if (!aligned(bank, 128) || !aligned(clauses, 4)) return MSAT_EALIGN;
    return cuda_rc(launch_compile_bank(plan, clauses, P, static_cast<uint8_t*>(bank), (cudaStream_t)stream));
}

int msat_reset(const msat_plan* plan, const void* bank, int32_t P, const int32_t* problem_idx, const uint32_t* keys,
               uint32_t* state, int32_t* obs, int32_t B, void* stream) {
    if (!plan || B < 0 || P <= 0) return MSAT_EINVAL;
    if (B > 0 && (!bank || !problem_idx || !keys || !state)) return MSAT_EINVAL;
    if (!aligned(bank, 128) || !aligned(state, 16) || !aligned(obs, 16)) return MSAT_EALIGN;
    EnvArgs a{};
    a.bank = static_cast<const uint8_t*>(bank); a.P = P;
    a.state_out = state; a.prob_idx = problem_idx; a.keys = keys; a.obs = obs; a.B = B;
    return cuda_rc(launch_env(plan, MODE_RESET, a, (cudaStream_t)stream));
}

int msat_step(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state_in, uint32_t* state_out,
              const int32_t* actions, int32_t auto_reset, const int32_t* new_problem_idx, const uint32_t* reset_keys,
              int32_t* obs, float* reward, int32_t reward_cols, uint8_t* done, int32_t done_cols, uint8_t* solved,
              int32_t* num_unsatisfied, int32_t* episode_step, int32_t B, void* stream) {
    if (!plan || B < 0 || P <= 0 || (done && done_cols <= 0) || (reward && reward_cols <= 0)) return MSAT_EINVAL;
    if (B > 0 && (!bank || !state_in || !state_out || !actions)) return MSAT_EINVAL;
    if (auto_reset && B > 0 && (!new_problem_idx || !reset_keys)) return MSAT_EINVAL;
    if (!aligned(bank, 128) || !aligned(state_in, 16) || !aligned(state_out, 16) || !aligned(obs, 16))
        return MSAT_EALIGN;
    EnvArgs a{};
    a.bank = static_cast<const uint8_t*>(bank); a.P = P;
    a.state_in = state_in; a.state_out = state_out; a.actions = actions;
    a.auto_reset = auto_reset ? 1 : 0; a.prob_idx = new_problem_idx; a.keys = reset_keys;
    a.obs = obs; a.reward = reward; a.reward_cols = reward_cols; a.done = done; a.done_cols = done_cols;
    a.solved = solved;
    a.num_unsat = num_unsatisfied; a.episode_step = episode_step; a.B = B;
    return cuda_rc(launch_env(plan, MODE_STEP, a, (cudaStream_t)stream));
}

int msat_rollout_step(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state_in,
                      uint32_t* state_out, const int32_t* actions, const uint32_t* rng_in, uint32_t* chain_out,
                      int32_t Bg, int32_t env_offset, int32_t* obs, float* reward, int32_t reward_cols, uint8_t* done,
                      int32_t done_cols, uint8_t* solved, int32_t* num_unsatisfied, int32_t* episode_step, int32_t B,
                      void* stream) {
    if (!plan || B < 0 || P <= 0 || (done && done_cols <= 0) || (reward && reward_cols <= 0)) return MSAT_EINVAL;
    if (!rng_in || !chain_out || Bg <= 0 || env_offset < 0 || (long long)env_offset + B > Bg) return MSAT_EINVAL;
    if (Bg > (1 << 30)) return MSAT_EUNSUPPORTED;
    {   // the advanced chain is written while other CTAs still read rng_in: the buffers must not overlap
        const uintptr_t r = reinterpret_cast<uintptr_t>(rng_in), c = reinterpret_cast<uintptr_t>(chain_out);
        if (r + 8 > c && c + 40 > r) return MSAT_EINVAL;
    }
    if (B > 0 && (!bank || !state_in || !state_out || !actions)) return MSAT_EINVAL;
    if (!aligned(bank, 128) || !aligned(state_in, 16) || !aligned(state_out, 16) || !aligned(obs, 16))
        return MSAT_EALIGN;
    EnvArgs a{};
    a.bank = static_cast<const uint8_t*>(bank); a.P = P;
    a.state_in = state_in; a.state_out = state_out; a.actions = actions;
    a.auto_reset = 1; a.rng_in = rng_in; a.chain_out = chain_out; a.Bg = (uint32_t)Bg; a.env_off = (uint32_t)env_offset;
    a.obs = obs; a.reward = reward; a.reward_cols = reward_cols; a.done = done; a.done_cols = done_cols;
    a.solved = solved;
    a.num_unsat = num_unsatisfied; a.episode_step = episode_step; a.B = B;
    if (B == 0) return cuda_rc(launch_rng_chain(rng_in, chain_out, (cudaStream_t)stream));
    return cuda_rc(launch_env(plan, MODE_STEP, a, (cudaStream_t)stream));
}

int msat_rollout_step_gnn(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state_in,
                          uint32_t* state_out, const int32_t* actions, const uint32_t* rng_in, uint32_t* chain_out,
                          int32_t Bg, int32_t env_offset, int32_t* assignment, float* clause_features, float* reward,
                          int32_t reward_cols, uint8_t* done, int32_t done_cols, uint8_t* solved,
                          int32_t* num_unsatisfied, int32_t* episode_step, int32_t B, void* stream) {
    if (!plan || B < 0 || P <= 0 || (done && done_cols <= 0) || (reward && reward_cols <= 0)) return MSAT_EINVAL;
    if (!rng_in || !chain_out || Bg <= 0 || env_offset < 0 || (long long)env_offset + B > Bg) return MSAT_EINVAL;
    if (Bg > (1 << 30)) return MSAT_EUNSUPPORTED;
    {
        const uintptr_t r = reinterpret_cast<uintptr_t>(rng_in), c = reinterpret_cast<uintptr_t>(chain_out);
        if (r + 8 > c && c + 40 > r) return MSAT_EINVAL;
    }
    if (B > 0 && (!bank || !state_in || !state_out || !actions)) return MSAT_EINVAL;
    if (!aligned(bank, 128) || !aligned(state_in, 16) || !aligned(state_out, 16)) return MSAT_EALIGN;
    EnvArgs a{};
    a.bank = static_cast<const uint8_t*>(bank); a.P = P;
    a.state_in = state_in; a.state_out = state_out; a.actions = actions;
    a.auto_reset = 1; a.rng_in = rng_in; a.chain_out = chain_out; a.Bg = (uint32_t)Bg; a.env_off = (uint32_t)env_offset;
    a.gnn_assign = assignment; a.gnn_cf = clause_features;
    a.reward = reward; a.reward_cols = reward_cols; a.done = done; a.done_cols = done_cols; a.solved = solved;
    a.num_unsat = num_unsatisfied; a.episode_step = episode_step; a.B = B;
    if (B == 0) return cuda_rc(launch_rng_chain(rng_in, chain_out, (cudaStream_t)stream));
    return cuda_rc(launch_env(plan, MODE_STEP, a, (cudaStream_t)stream));
}

int msat_get_obs(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state, int32_t* obs, int32_t B,
                 void* stream) {
    if (!plan || B < 0 || P <= 0) return MSAT_EINVAL;
    if (B > 0 && (!bank || !state || !obs)) return MSAT_EINVAL;
    if (!aligned(bank, 128) || !aligned(state, 16) || !aligned(obs, 16)) return MSAT_EALIGN;
    EnvArgs a{};
    a.bank = static_cast<const uint8_t*>(bank); a.P = P;
    a.state_in = state; a.obs = obs; a.B = B;
    return cuda_rc(launch_env(plan, MODE_OBS, a, (cudaStream_t)stream));
}

int msat_export_state(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state, int32_t B,
                      int32_t* variable_assignments, uint8_t* clauses_satisfied_status, int32_t* num_unsatisfied,
                      int32_t* step, uint8_t* done, int32_t* clauses, int32_t* agent_clause_masks,
                      int32_t* agent_neighbor_masks, int32_t* literal_to_agent_idx, int32_t* problem_idx,
                      void* stream) {
    if (!plan || B < 0 || P <= 0) return MSAT_EINVAL;
    if (B > 0 && (!bank || !state)) return MSAT_EINVAL;
    ExportArgs a{};
    a.bank = static_cast<const uint8_t*>(bank); a.P = P; a.state = state; a.B = B;
    a.assign = variable_assignments; a.sat = clauses_satisfied_status; a.num_unsat = num_unsatisfied;
    a.step = step; a.done = done; a.clauses = clauses; a.acm = agent_clause_masks; a.anm = agent_neighbor_masks;
    a.l2a = literal_to_agent_idx; a.pidx = problem_idx;
    return cuda_rc(launch_export(plan, a, (cudaStream_t)stream));
}

int msat_rollout_step_host(const msat_plan* plan, const void* bank, int32_t P, uint32_t* state,
                           const int32_t* actions_host, int32_t* actions_dev, const uint32_t* rng_in,
                           uint32_t* chain_out, int32_t Bg, int32_t env_offset, int32_t* obs_dev, float* reward_dev,
                           int32_t reward_cols, uint8_t* done_dev, int32_t done_cols, uint8_t* solved_dev,
                           int32_t* num_unsatisfied_dev, int32_t* episode_step_dev, float* reward_host,
                           uint8_t* done_host, uint8_t* solved_host, int32_t* num_unsatisfied_host,
                           int32_t* episode_step_host, int32_t B, void* stream) {
    if (!plan || !actions_host || !actions_dev || B < 0) return MSAT_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const Dims& d = plan->d;
    const size_t act_per_env = (size_t)d.A * (d.action_mode == 0 ? 1 : d.V);
    const size_t obs_per_env = (size_t)d.A * d.D;

    // One slice [b0, b0 + bc) of the batch on stream `w`: actions in, fused step, results out.
    auto run_slice = [&](int b0, int bc, cudaStream_t w) -> int {
        cudaError_t e = cudaSuccess;
        if (bc > 0 && act_per_env)
            e = cudaMemcpyAsync(actions_dev + b0 * act_per_env, actions_host + b0 * act_per_env,
                                (size_t)bc * act_per_env * sizeof(int32_t), cudaMemcpyHostToDevice, w);
        if (e != cudaSuccess) return (int)e;
        uint32_t* st = state + (size_t)b0 * d.state_words;
        int rc = msat_rollout_step(plan, bank, P, st, st, actions_dev + b0 * act_per_env, rng_in, chain_out, Bg,
                                   env_offset + b0, obs_dev ? obs_dev + b0 * obs_per_env : nullptr,
                                   reward_dev ? reward_dev + (size_t)b0 * reward_cols : nullptr, reward_cols,
                                   done_dev ? done_dev + (size_t)b0 * done_cols : nullptr, done_cols,
                                   solved_dev ? solved_dev + b0 : nullptr,
                                   num_unsatisfied_dev ? num_unsatisfied_dev + b0 : nullptr,
                                   episode_step_dev ? episode_step_dev + b0 : nullptr, bc, (void*)w);
        if (rc != MSAT_OK || bc == 0) return rc;
        // device->host result copies; outputs that are adjacent in both address spaces (the Python layer
        // carves them out of one device block and one pinned block) travel as ONE copy
        struct Seg { const char* dev; char* host; size_t bytes; };
        Seg seg[5];
        int ns = 0;
        auto add = [&](const void* dv, void* hs, size_t off, size_t bytes) {
            if (dv && hs && bytes) seg[ns++] = Seg{static_cast<const char*>(dv) + off, static_cast<char*>(hs) + off, bytes};
        };
        add(reward_dev, reward_host, (size_t)b0 * reward_cols * sizeof(float), (size_t)bc * reward_cols * sizeof(float));
        add(num_unsatisfied_dev, num_unsatisfied_host, (size_t)b0 * 4, (size_t)bc * 4);
        add(episode_step_dev, episode_step_host, (size_t)b0 * 4, (size_t)bc * 4);
        add(done_dev, done_host, (size_t)b0 * done_cols, (size_t)bc * done_cols);
        add(solved_dev, solved_host, (size_t)b0, (size_t)bc);
        for (int i = 1; i < ns; ++i)          // insertion sort by device address
            for (int j = i; j > 0 && seg[j].dev < seg[j - 1].dev; --j) { Seg t = seg[j]; seg[j] = seg[j - 1]; seg[j - 1] = t; }
        for (int i = 0; i < ns && e == cudaSuccess;) {
            Seg cur = seg[i++];
            while (i < ns && seg[i].dev == cur.dev + cur.bytes && seg[i].host == cur.host + cur.bytes) cur.bytes += seg[i++].bytes;
            e = cudaMemcpyAsync(cur.host, cur.dev, cur.bytes, cudaMemcpyDeviceToHost, w);
        }
        return (int)e;
    };

    constexpr int kSlices = 4;
    if (B < 32768) {   // one slice on the caller's stream (slicing only pays when the kernel is long)
        int rc = run_slice(0, B, s);
        if (rc != MSAT_OK) return rc;
        return cuda_rc(cudaStreamSynchronize(s));
    }
    // Large batch: slices alternate between two internal streams so that the action upload of slice
    // i+1 and the result download of slice i-1 overlap the kernel of slice i (separate copy engines).
    {
        std::lock_guard<std::mutex> lock(g_pipe.mu);
        cudaError_t e = g_pipe.init();
        if (e != cudaSuccess) return (int)e;
        e = cudaEventRecord(g_pipe.fork, s);
        for (int w = 0; w < 2 && e == cudaSuccess; ++w) e = cudaStreamWaitEvent(g_pipe.ws[w], g_pipe.fork, 0);
        if (e != cudaSuccess) return (int)e;
        const int per = (((B + kSlices - 1) / kSlices) + 63) & ~63;   // multiple of 64 envs keeps every slice aligned
        for (int c = 0, b0 = 0; b0 < B; ++c, b0 += per) {
            const int bc = B - b0 < per ? B - b0 : per;
            int rc = run_slice(b0, bc, g_pipe.ws[c & 1]);
            if (rc != MSAT_OK) return rc;
        }
        for (int w = 0; w < 2 && e == cudaSuccess; ++w) {
            e = cudaEventRecord(g_pipe.join[w], g_pipe.ws[w]);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s, g_pipe.join[w], 0);
        }
        if (e != cudaSuccess) return (int)e;
    }
    return cuda_rc(cudaStreamSynchronize(s));
}

int msat_rng_chain(const uint32_t* rng_in, uint32_t* chain_out, void* stream) {
    if (!rng_in || !chain_out) return MSAT_EINVAL;
    return cuda_rc(launch_rng_chain(rng_in, chain_out, (cudaStream_t)stream));
}

int msat_rng_split2(const uint32_t* key_in, uint32_t* out, void* stream) {
    if (!key_in || !out) return MSAT_EINVAL;
    return cuda_rc(launch_rng_split2(key_in, out, (cudaStream_t)stream));
}

int msat_env_keys(const uint32_t* prob_key, const uint32_t* reset_key, int32_t Bg, int32_t off, int32_t Bl, int32_t P,
                  int32_t* problem_idx, uint32_t* reset_keys, void* stream) {
    if (Bg < 0 || off < 0 || Bl < 0 || (long long)off + Bl > Bg) return MSAT_EINVAL;
    if (problem_idx && (!prob_key || P <= 0)) return MSAT_EINVAL;
    if (reset_keys && !reset_key) return MSAT_EINVAL;
    if (Bg > (1 << 30)) return MSAT_EUNSUPPORTED;
    return cuda_rc(launch_env_keys(prob_key, reset_key, Bg, off, Bl, P, problem_idx, reset_keys, (cudaStream_t)stream));
}

int msat_gae(const float* reward, int64_t rs_t, int64_t rs_b, const uint8_t* done, const float* value,
             const float* last_val, double gamma, double gae_lambda, float* advantages, float* targets, double* stats,
             int32_t T, int32_t B, void* stream) {
    if (T < 0 || B < 0) return MSAT_EINVAL;
    if (T > 0 && B > 0 && (!reward || !done || !value || !last_val || !advantages || !targets)) return MSAT_EINVAL;
    return cuda_rc(launch_gae(reward, rs_t, rs_b, done, value, last_val, (float)gamma, (float)(gamma * gae_lambda),
                              advantages, targets, T, B, stats, (cudaStream_t)stream));
}

int msat_adv_stats(const float* adv, int64_t count, double* stats, void* stream) {
    if (count < 0 || !stats || (count > 0 && !adv)) return MSAT_EINVAL;
    return cuda_rc(launch_adv_stats(adv, count, stats, (cudaStream_t)stream));
}

int msat_adv_normalize(float* adv, int64_t count, const double* stats, void* stream) {
    if (count < 0 || !stats || (count > 0 && !adv)) return MSAT_EINVAL;
    return cuda_rc(launch_adv_normalize(adv, count, stats, (cudaStream_t)stream));
}

int msat_gnn_static(const msat_plan* plan, const void* bank, int32_t P, float* static_var_features, float* a_pos,
                    float* a_neg, void* stream) {
    if (!plan || P < 0 || (P > 0 && !bank)) return MSAT_EINVAL;
    if (2 * plan->d.n * (int)sizeof(int) > 48 * 1024) return MSAT_EUNSUPPORTED;
    return cuda_rc(launch_gnn_static(plan, static_cast<const uint8_t*>(bank), P, static_var_features, a_pos, a_neg,
                                     (cudaStream_t)stream));
}

int msat_gnn_dynamic(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state, int32_t B,
                     int32_t* assignment, float* clause_features, void* stream) {
    if (!plan || B < 0 || P <= 0 || (B > 0 && (!bank || !state))) return MSAT_EINVAL;
    return cuda_rc(launch_gnn_dynamic(plan, static_cast<const uint8_t*>(bank), P, state, B, assignment,
                                      clause_features, (cudaStream_t)stream));
}

int msat_rollout_metrics(const float* reward, int64_t rs_t, int64_t rs_b, const uint8_t* done, const uint8_t* solved,
                         const int32_t* num_unsatisfied, const int32_t* episode_step, int32_t T, int32_t B,
                         double* sums, void* stream) {
    if (T < 0 || B < 0 || !sums) return MSAT_EINVAL;
    if (T > 0 && B > 0 && (!reward || !done || !solved || !num_unsatisfied || !episode_step)) return MSAT_EINVAL;
    return cuda_rc(launch_rollout_metrics(reward, rs_t, rs_b, done, solved, num_unsatisfied, episode_step, T, B, sums,
                                          (cudaStream_t)stream));
}

int msat_flip_gains(const msat_plan* plan, const void* bank, int32_t P, const uint32_t* state, int32_t B, double tau,
                    int32_t* delta_unsat, int32_t* greedy_labels, void* stream) {
    if (!plan || B < 0 || P <= 0 || (B > 0 && (!bank || !state))) return MSAT_EINVAL;
    if (plan->d.n * (int)sizeof(int) > 48 * 1024) return MSAT_EUNSUPPORTED;
    return cuda_rc(launch_flip_gains(plan, static_cast<const uint8_t*>(bank), P, state, B, (float)tau, delta_unsat,
                                     greedy_labels, (cudaStream_t)stream));
}

int msat_eval_track(const msat_plan* plan, const uint32_t* state, const uint8_t* solved, int32_t t, int32_t B,
                    uint8_t* ever_solved, int32_t* steps_to_solve, int32_t* solution, void* stream) {
    if (!plan || B < 0 || t < 0) return MSAT_EINVAL;
    if (B > 0 && (!state || !solved || !ever_solved || !steps_to_solve || !solution)) return MSAT_EINVAL;
    return cuda_rc(launch_eval_track(plan, state, solved, t, B, ever_solved, steps_to_solve, solution,
                                     (cudaStream_t)stream));
}

}  // extern "C"
