// Batched SATEnv kernels for sm_100a: formula-bank compiler, reset / step(+auto-reset) / get_obs,
// and the SATState exporter.  Reference semantics: src/envs/multi_agent_sat_env.py (cited as env:LINE)
// and src/learners/mappo_gnn_sat_learner.py:422-464 (auto-reset).  See DESIGN.md section 4.
#include "internal.h"

namespace msat {

// =====================================================================================
// K_compile: one CTA per formula.  Evaluates the agent<->clause and agent<->neighbour
// relations of _compute_observation_maps (env:99-128) once per formula and stores them as a
// flat A*D-bit mask stream in observation order [own(n) | clauses(m) | neighbours(n)] per agent,
// next to the literals packed literal-major as u16 codes ((var << 1) | negated, 0xFFFF for a 0 padding literal).
// =====================================================================================
__global__ void __launch_bounds__(256) compile_bank_kernel(Dims d, const int32_t* __restrict__ clauses,
                                                           uint8_t* __restrict__ bank) {
    extern __shared__ uint32_t csm[];
    uint32_t* clause_agents = csm;                 // [m][agw]  agents related to clause c
    uint32_t* var_agents = csm + d.m * d.agw;      // [n][agw]  agents that see variable v in a related clause
    const int tid = threadIdx.x, nt = blockDim.x;
    const int32_t* cl = clauses + (size_t)blockIdx.x * d.m * d.k;
    uint8_t* rec = bank + (size_t)blockIdx.x * d.rec_bytes;
    uint16_t* lits = reinterpret_cast<uint16_t*>(rec);
    uint32_t* mflat = reinterpret_cast<uint32_t*>(rec + d.lits_bytes);

    for (int i = tid; i < (d.m + d.n) * d.agw; i += nt) csm[i] = 0u;
    for (int i = d.m * d.k + tid; i < d.lits_bytes / 2; i += nt) lits[i] = LIT_PAD;
    for (int i = d.lits_bytes / 4 + d.fw + 1 + tid; i < d.rec_bytes / 4; i += nt)
        reinterpret_cast<uint32_t*>(rec)[i] = 0u;
    __syncthreads();

    for (int c = tid; c < d.m; c += nt) {
        uint32_t* ca = clause_agents + c * d.agw;
        for (int j = 0; j < d.k; ++j) {
            const int lit = cl[c * d.k + j];
            uint16_t code;
            if (lit == 0) {
                // env:100,106: |0|-1 = -1 equals the -1 padding of every agent that owns fewer than
                // V variables, so the clause becomes "related" to all of those agents.
                code = LIT_PAD;
                if (d.rem > 0)
                    for (int a = d.rem; a < d.A; ++a) ca[a >> 5] |= 1u << (a & 31);
            } else {
                int v = (lit < 0 ? -lit : lit) - 1;
                v = v < d.n ? v : d.n - 1;
                code = (uint16_t)((v << 1) | (lit < 0 ? 1 : 0));
                const int a = var_to_agent(d, v);
                ca[a >> 5] |= 1u << (a & 31);
            }
            lits[lit_index(d.m, c, j)] = code;
        }
    }
    __syncthreads();
    // env:116-121: every real variable of a related clause is a candidate neighbour.
    for (int c = tid; c < d.m; c += nt) {
        for (int j = 0; j < d.k; ++j) {
            const uint16_t code = lits[lit_index(d.m, c, j)];
            if (code == LIT_PAD) continue;
            const int v = code >> 1;
            for (int w = 0; w < d.agw; ++w) {
                const uint32_t x = clause_agents[c * d.agw + w];
                if (x) atomicOr(&var_agents[v * d.agw + w], x);
            }
        }
    }
    __syncthreads();
    for (int w = tid; w < d.fw + 1; w += nt) {
        uint32_t bits = 0u;
        for (int b = 0; b < 32; ++b) {
            const int j = w * 32 + b;
            if (j >= d.AD) break;
            const int a = j / d.D, p = j - a * d.D;
            uint32_t bit;
            if (p < d.n) {
                bit = var_to_agent(d, p) == a;                                         // env:355
            } else if (p < d.n + d.m) {
                bit = (clause_agents[(p - d.n) * d.agw + (a >> 5)] >> (a & 31)) & 1u;  // env:112
            } else {
                const int v = p - d.n - d.m;
                bit = ((var_agents[v * d.agw + (a >> 5)] >> (a & 31)) & 1u) && var_to_agent(d, v) != a;   // env:123
            }
            bits |= bit << b;
        }
        mflat[w] = bits;
    }
}

// =====================================================================================
// Device pieces of the env kernel.  One group of GS threads owns one env.
// =====================================================================================

// Clause truth via warp ballots (env:130-156): lane = clause, ballot word = 32 clause-status bits.
// One literal: true <=> assignment bit != negation flag; a 0 padding literal is never true (env:141-144).
__device__ __forceinline__ bool literal_true(uint32_t code, const uint32_t* assign) {
    const uint32_t v = code >> 1;
    return code != LIT_PAD && (((assign[v >> 5] >> (v & 31)) ^ code) & 1u) != 0u;
}
template <int K>
__device__ __forceinline__ bool clause_true_fixed(const uint16_t* lits, int m, int c, const uint32_t* assign) {
    uint32_t code[K];
#pragma unroll
    for (int j = 0; j < K; ++j) code[j] = lits[lit_index(m, c, j)];     // independent loads first
    bool sat = false;
#pragma unroll
    for (int j = 0; j < K; ++j) sat |= literal_true(code[j], assign);
    return sat;
}

template <int GS>
__device__ __forceinline__ void eval_clauses(const Dims& d, const uint16_t* lits, const uint32_t* assign,
                                             uint32_t* satw, int* nunsat, int gt) {
    const int lane = gt & 31;
    int local = 0;
    for (int w = gt >> 5; w < d.sw; w += GS / 32) {
        const int c = w * 32 + lane;
        const bool valid = c < d.m;
        bool sat = false;
        if (valid) {
            if (d.k == 3) {
                sat = clause_true_fixed<3>(lits, d.m, c, assign);
            } else {
#pragma unroll 4
                for (int j = 0; j < d.k; ++j) sat |= literal_true(lits[lit_index(d.m, c, j)], assign);
            }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, valid && sat);
        const uint32_t bad = __ballot_sync(0xffffffffu, valid && !sat);
        if (lane == 0) {
            satw[w] = word;
            local += __popc(bad);
        }
    }
    if (lane == 0 && local) atomicAdd(nunsat, local);
}

// assign = randint(key, (n,), 0, 2) (env:162): bit 0 of threefry_2x32(split(key)[1], arange(n)).
template <int GS>
__device__ __forceinline__ void threefry_assign(const Dims& d, uint32_t k0, uint32_t k1, uint32_t* assign, int gt) {
    uint32_t ka[2], kb[2];
    split2(k0, k1, ka, kb);
    const int half = (d.n + 1) >> 1;
    for (int i = gt; i < half; i += GS) {
        uint32_t x0 = (uint32_t)i;
        uint32_t x1 = (half + i < d.n) ? (uint32_t)(half + i) : 0u;
        threefry2x32(kb[0], kb[1], x0, x1);
        if (x0 & 1u) atomicOr(&assign[i >> 5], 1u << (i & 31));
        const int v = half + i;
        if (v < d.n && (x1 & 1u)) atomicOr(&assign[v >> 5], 1u << (v & 31));
    }
}

// Apply all agents' flips simultaneously (env:233-250) on the packed assignment in shared memory.
template <int GS>
__device__ __forceinline__ void apply_actions(const Dims& d, const int32_t* __restrict__ actions, int e,
                                              uint32_t* assign, int gt) {
    if (d.action_mode == 0) {
        const int32_t* act = actions + (size_t)e * d.A;
        for (int a = gt; a < d.A; a += GS) {
            const int x = act[a];
            const int size = group_size(d, a);
            if (x >= size) continue;                       // env:236 no-op
            int idx = x < size - 1 ? x : size - 1;         // env:238
            if (idx < 0) idx += d.V;                       // JAX gather: wrap once, then clamp
            idx = idx < 0 ? 0 : (idx > d.V - 1 ? d.V - 1 : idx);
            if (idx >= size) continue;                     // landed on a -1 pad: one_hot(-1) = 0 (env:243)
            const int v = group_start(d, a) + idx;
            atomicXor(&assign[v >> 5], 1u << (v & 31));
        }
    } else {
        const int32_t* act = actions + (size_t)e * d.A * d.V;
        for (int i = gt; i < d.A * d.V; i += GS) {
            const int a = i / d.V, j = i - a * d.V;
            if (j < group_size(d, a) && (act[i] & 1)) {    // env:246-250 (actions are 0/1)
                const int v = group_start(d, a) + j;
                atomicXor(&assign[v >> 5], 1u << (v & 31));
            }
        }
    }
}

// Four observation ints from a 4-bit mask nibble and a 4-bit value nibble: out = mask ? value : -1.
// The nibbles are spread to one byte per element with a multiply, combined bytewise into
// {0xFF, 0x00, 0x01} and sign-extended to int32 with one PRMT each.
__device__ __forceinline__ int4 expand_nibble(uint32_t mn, uint32_t xn) {
    const uint32_t m4 = (mn * 0x00204081u) & 0x01010101u;          // bit i -> byte i
    const uint32_t r = 0xFFFFFFFFu - m4 * 0xFFu;                   // byte: mask ? 0x00 : 0xFF
    const uint32_t c = ((xn * 0x00204081u) & 0x01010101u) | r;     // byte: mask ? value : 0xFF
    // prmt.b32 selector nibble: bits 2:0 pick the byte, bit 3 replicates its sign instead (the
    // __byte_perm intrinsic masks bit 3 away, hence inline PTX): byte i sign-extended to 32 bits.
    int4 v;
    asm("prmt.b32 %0, %1, %1, 0x8880;" : "=r"(v.x) : "r"(c));
    asm("prmt.b32 %0, %1, %1, 0x9991;" : "=r"(v.y) : "r"(c));
    asm("prmt.b32 %0, %1, %1, 0xAAA2;" : "=r"(v.z) : "r"(c));
    asm("prmt.b32 %0, %1, %1, 0xBBB3;" : "=r"(v.w) : "r"(c));
    return v;
}

// Observation writer (env:345-398).  Per env the A*D output ints are one flat range of the
// [B,A,D] tensor.  out[i] = mask[i] ? value[i] : -1 where mask is the bank's flat bit stream and
// value is [assign | clause status | assign] repeated per agent.  Both streams are re-based to the
// 128-byte aligned start of the range so every lane owns whole 16-byte chunks (st.global.cs.v4).
template <int GS>
__device__ __forceinline__ void emit_obs(const Dims& d, int e, const uint32_t* assign, const uint32_t* satw,
                                         uint32_t* X, uint2* smx, const uint32_t* mflat,
                                         int32_t* __restrict__ obs, int gid, int gt) {
    for (int w = gt; w < d.xw; w += GS) {
        const int pos = 32 * w;
        X[w] = extract32(assign, d.aw, pos) | extract32(satw, d.sw, pos - d.n) |
               extract32(assign, d.aw, pos - d.n - d.m);
    }
    group_sync<GS>(gid);

    const long long g_start = (long long)e * d.AD;
    const int s = (int)(g_start & 31);
    const int nw = (s + d.AD + 31) >> 5;
    for (int w = gt; w < nw; w += GS) {
        const int j0 = 32 * w - s;
        const uint32_t mw = extract32(mflat, d.fw + 1, j0);
        uint32_t xw = 0u;
        int j = j0 > 0 ? j0 : 0;
        const int jend = (j0 + 32 < d.AD) ? j0 + 32 : d.AD;
        if (j < jend) {
            int arow = (int)__umulhi((uint32_t)j, d.inv_D);     // j / D with a rounded-up reciprocal (+0/+1)
            if (arow * d.D > j) --arow;
            int p = j - arow * d.D;
            while (j < jend) {
                int len = d.D - p;
                if (len > jend - j) len = jend - j;
                uint32_t bits = extract32(X, d.xw, p);
                if (len < 32) bits &= (1u << len) - 1u;
                xw |= bits << (j - j0);
                j += len;
                p = 0;
            }
        }
        smx[w] = make_uint2(mw, xw);
    }
    group_sync<GS>(gid);

    // ---- streaming stores: chunk q covers ints [4q, 4q+4) of the re-based range ----
    int32_t* out = obs + (g_start - s);
    const int lo_valid = s, hi_valid = s + d.AD;
    const int q_lo = (lo_valid + 3) >> 2;       // first chunk that lies completely inside the env's range
    const int q_hi = hi_valid >> 2;             // one past the last complete chunk
    {
        // GS is a multiple of 8, so a lane keeps the same nibble position in every iteration.
        int q = q_lo + gt;
        const int sh = (q & 7) * 4;
        const uint2* sp = smx + (q >> 3);
        int4* op = reinterpret_cast<int4*>(out) + q;
#pragma unroll 4
        for (; q < q_hi; q += GS, sp += GS / 8, op += GS) {
            const uint2 mx = *sp;
            __stcs(op, expand_nibble((mx.x >> sh) & 0xFu, (mx.y >> sh) & 0xFu));
        }
    }
    // at most one partial chunk at each end (the env's range is only 4-byte aligned): scalar stores
    if (gt < 8) {
        const int q = (gt < 4) ? q_lo - 1 : q_hi;
        const int i = 4 * q + (gt & 3);
        const bool partial = (gt < 4) ? (4 * q_lo != lo_valid) : (4 * q_hi != hi_valid);
        if (partial && q >= 0 && i >= lo_valid && i < hi_valid) {
            const uint2 mx = smx[i >> 5];
            out[i] = ((mx.x >> (i & 31)) & 1u) ? (int)((mx.y >> (i & 31)) & 1u) : -1;
        }
    }
}

// Dynamic GNN input of the env's (post-reset) state straight from the shared-memory formula record
// (learner:165-195): assignment int32[n]; clause_features float[m][3] = {is_sat, #true literals / 3.0, 1}.
// Lets a GNN-style consumer skip the local observations altogether.
template <int GS>
__device__ __forceinline__ void emit_gnn(const Dims& d, int e, const uint16_t* lits, const uint32_t* assign,
                                         int32_t* __restrict__ gnn_assign, float* __restrict__ gnn_cf, int gt) {
    if (gnn_assign)
        for (int v = gt; v < d.n; v += GS) gnn_assign[(size_t)e * d.n + v] = (int)((assign[v >> 5] >> (v & 31)) & 1u);
    if (gnn_cf)
        for (int c = gt; c < d.m; c += GS) {
            int ntrue = 0;
            for (int j = 0; j < d.k; ++j) ntrue += literal_true(lits[lit_index(d.m, c, j)], assign) ? 1 : 0;
            float* o = gnn_cf + ((size_t)e * d.m + c) * 3;
            o[0] = ntrue > 0 ? 1.0f : 0.0f;
            o[1] = __fdiv_rn((float)ntrue, 3.0f);       // the literal 3.0 of learner:185 for every clause width
            o[2] = 1.0f;
        }
}

// =====================================================================================
// K_env<GS, MODE>: reset / step (+ fused auto-reset) / get_obs.  256-thread CTAs, 256/GS envs each.
// =====================================================================================
template <int GS, int MODE>
__global__ void __launch_bounds__(kCtaThreads) env_kernel(const Dims d, const EnvArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int gid = threadIdx.x / GS, gt = threadIdx.x % GS;
    const int e = blockIdx.x * (kCtaThreads / GS) + gid;
    if (e >= a.B) return;   // whole group leaves together

    const GroupLayout L = group_layout(d);
    uint8_t* base = smem_raw + (size_t)gid * L.total;
    uint8_t* rec = base + L.rec;
    uint32_t* st = reinterpret_cast<uint32_t*>(base + L.st);
    uint32_t* satw = reinterpret_cast<uint32_t*>(base + L.satw);
    uint32_t* X = reinterpret_cast<uint32_t*>(base + L.x);
    uint2* smx = reinterpret_cast<uint2*>(base + L.smx);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + L.bar);
    int* misc = reinterpret_cast<int*>(base + L.misc);
    const uint16_t* lits = reinterpret_cast<const uint16_t*>(rec);
    const uint32_t* mflat = reinterpret_cast<const uint32_t*>(rec + d.lits_bytes);
    uint32_t* st_tail = st + d.aw;

    // ---- stage inputs: state record (plain loads) and bank record (TMA bulk copy) ----
    if (MODE == MODE_RESET) {
        for (int i = gt; i < d.state_words; i += GS) st[i] = 0u;
    } else {
        const uint32_t* sin = a.state_in + (size_t)e * d.state_words;
        for (int i = gt; i < d.state_words; i += GS) st[i] = sin[i];
    }
    if (gt == 0) {
        misc[0] = 0;
        mbar_init(bar, 1);
    }
    group_sync<GS>(gid);

    // Every thread takes its copy of the scalar state fields NOW: thread 0 rewrites them in shared memory
    // further down, and a warp that read `step` after that write would disagree with the others about
    // `done` one step before the time-out and wait at a group barrier nobody else reaches.
    const int step_old = (MODE == MODE_RESET) ? 0 : (int)st_tail[ST_STEP];
    int pidx;
    if (MODE == MODE_RESET) pidx = a.prob_idx[e];
    else pidx = (int)st_tail[ST_PIDX];
    pidx = pidx < 0 ? 0 : (pidx >= a.P ? a.P - 1 : pidx);
    if (gt == 0) {
        mbar_expect_tx(bar, (uint32_t)d.rec_bytes);
        tma_load_1d(rec, a.bank + (size_t)pidx * d.rec_bytes, (uint32_t)d.rec_bytes, bar);
    }

    // ---- new assignment while the record is in flight ----
    if (MODE == MODE_RESET) {
        threefry_assign<GS>(d, a.keys[2 * (size_t)e], a.keys[2 * (size_t)e + 1], st, gt);
    } else if (MODE == MODE_STEP) {
        apply_actions<GS>(d, a.actions, e, st, gt);
    }
    group_sync<GS>(gid);
    mbar_wait(bar, 0);

    eval_clauses<GS>(d, lits, st, satw, &misc[0], gt);
    group_sync<GS>(gid);
    int nunsat = misc[0];

    if (MODE == MODE_STEP) {
        if (a.rng_in && blockIdx.x == 0 && threadIdx.x == 0) {
            // advance the rollout rng once per step (learner:397,416,426); chain_out never aliases rng_in
            uint32_t c[10];
            rng_chain_compute(a.rng_in[0], a.rng_in[1], c);
            for (int i = 0; i < 10; ++i) a.chain_out[i] = c[i];
        }
        const bool solved = nunsat == 0;                                  // env:257
        const bool done = solved || (step_old + 1 >= d.max_steps);        // env:258-259
        // pre-reset outputs stored in the Transition (learner:467-478)
        if (a.reward)
            for (int i = gt; i < a.reward_cols; i += GS)
                a.reward[(size_t)e * a.reward_cols + i] = solved ? 1.0f : 0.0f;                     // env:193
        if (a.done)
            for (int i = gt; i < a.done_cols; i += GS) a.done[(size_t)e * a.done_cols + i] = done ? 1 : 0;
        if (gt == 0) {
            if (a.solved) a.solved[e] = solved ? 1 : 0;
            if (a.num_unsat) a.num_unsat[e] = nunsat;
            if (a.episode_step) a.episode_step[e] = step_old + 1;        // env:281
        }
        if (done && a.auto_reset) {
            // learner:425-464: swap in a fresh episode on a newly drawn formula (group-uniform branch)
            group_sync<GS>(gid);   // everyone is done reading the old record / misc
            uint32_t rk0, rk1;
            if (a.rng_in) {
                // fused key derivation (learner:426-434) from the rollout rng, global env index
                if (gt == 0) {
                    uint32_t c[10], k[2];
                    rng_chain_compute(a.rng_in[0], a.rng_in[1], c);
                    misc[1] = (int)env_problem_index(c[6], c[7], a.Bg, a.env_off + (uint32_t)e, (uint32_t)a.P);
                    env_reset_key(c[8], c[9], a.Bg, a.env_off + (uint32_t)e, k);
                    misc[2] = (int)k[0];
                    misc[3] = (int)k[1];
                }
                group_sync<GS>(gid);
                pidx = misc[1];
                rk0 = (uint32_t)misc[2];
                rk1 = (uint32_t)misc[3];
            } else {
                pidx = a.prob_idx[e];
                rk0 = a.keys[2 * (size_t)e];
                rk1 = a.keys[2 * (size_t)e + 1];
            }
            pidx = pidx < 0 ? 0 : (pidx >= a.P ? a.P - 1 : pidx);
            if (gt == 0) {
                misc[0] = 0;
                fence_proxy_async();
                mbar_expect_tx(bar, (uint32_t)d.rec_bytes);
                tma_load_1d(rec, a.bank + (size_t)pidx * d.rec_bytes, (uint32_t)d.rec_bytes, bar);
            }
            for (int i = gt; i < d.aw; i += GS) st[i] = 0u;
            group_sync<GS>(gid);
            threefry_assign<GS>(d, rk0, rk1, st, gt);
            group_sync<GS>(gid);
            mbar_wait(bar, 1);
            eval_clauses<GS>(d, lits, st, satw, &misc[0], gt);
            group_sync<GS>(gid);
            nunsat = misc[0];
            if (gt == 0) {
                st_tail[ST_STEP] = 0u;
                st_tail[ST_PIDX] = (uint32_t)pidx;
                st_tail[ST_NUNSAT] = (uint32_t)nunsat;
                st_tail[ST_FLAGS] = 0u;
            }
        } else if (gt == 0) {
            st_tail[ST_STEP] = (uint32_t)(step_old + 1);                 // env:269
            st_tail[ST_NUNSAT] = (uint32_t)nunsat;
            st_tail[ST_FLAGS] = done ? 1u : 0u;                          // env:270
        }
    } else if (MODE == MODE_RESET) {
        if (gt == 0) {
            st_tail[ST_STEP] = 0u;                                       // env:170
            st_tail[ST_PIDX] = (uint32_t)pidx;
            st_tail[ST_NUNSAT] = (uint32_t)nunsat;
            st_tail[ST_FLAGS] = 0u;                                      // env:171
        }
    }
    if (MODE != MODE_OBS) {
        group_sync<GS>(gid);
        uint32_t* sout = a.state_out + (size_t)e * d.state_words;
        for (int i = gt; i < d.state_words; i += GS) sout[i] = st[i];
    }
    if (a.gnn_assign || a.gnn_cf) emit_gnn<GS>(d, e, lits, st, a.gnn_assign, a.gnn_cf, gt);
    if (a.obs) emit_obs<GS>(d, e, st, satw, X, smx, mflat, a.obs, gid, gt);
}

template <int GS>
static cudaError_t launch_env_gs(const msat_plan* plan, EnvMode mode, const EnvArgs& a, cudaStream_t s, int smem_bytes) {
    const int groups = kCtaThreads / GS;
    const int grid = (a.B + groups - 1) / groups;
    if (grid == 0) return cudaSuccess;
    const void* fn = nullptr;
    switch (mode) {
        case MODE_RESET: fn = (const void*)env_kernel<GS, MODE_RESET>; break;
        case MODE_STEP: fn = (const void*)env_kernel<GS, MODE_STEP>; break;
        default: fn = (const void*)env_kernel<GS, MODE_OBS>; break;
    }
    if (smem_bytes > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (err != cudaSuccess) return err;
    }
    Dims d = plan->d;
    EnvArgs args = a;
    void* params[] = {&d, &args};
    return cudaLaunchKernel(fn, dim3(grid), dim3(kCtaThreads), params, (size_t)smem_bytes, s);
}

cudaError_t launch_env(const msat_plan* plan, EnvMode mode, const EnvArgs& a, cudaStream_t s) {
    // the group size follows the work per env: observation-writing launches use the plan's GS, launches
    // without observations the small-group variant
    const bool noobs = a.obs == nullptr;
    const int gs = noobs ? plan->group_threads_noobs : plan->group_threads;
    const int smem = noobs ? plan->smem_bytes_noobs : plan->smem_bytes;
    switch (gs) {
        case 32: return launch_env_gs<32>(plan, mode, a, s, smem);
        case 64: return launch_env_gs<64>(plan, mode, a, s, smem);
        case 128: return launch_env_gs<128>(plan, mode, a, s, smem);
        default: return launch_env_gs<256>(plan, mode, a, s, smem);
    }
}

cudaError_t launch_compile_bank(const msat_plan* plan, const int32_t* clauses, int P, uint8_t* bank, cudaStream_t s) {
    if (P == 0) return cudaSuccess;
    if (plan->compile_smem_bytes > 48 * 1024) {
        cudaError_t err = cudaFuncSetAttribute((const void*)compile_bank_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, plan->compile_smem_bytes);
        if (err != cudaSuccess) return err;
    }
    compile_bank_kernel<<<P, 256, plan->compile_smem_bytes, s>>>(plan->d, clauses, bank);
    return cudaGetLastError();
}

// =====================================================================================
// K_export: reference-shaped SATState leaves (env:13-24) from packed state + bank.  Off the hot
// path (API fidelity, parity tests); one 128-thread CTA per env, global-memory reads only.
// =====================================================================================
__global__ void __launch_bounds__(128) export_kernel(const Dims d, const ExportArgs a) {
    const int e = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t* st = a.state + (size_t)e * d.state_words;
    const uint32_t* tail = st + d.aw;
    int pidx = (int)tail[ST_PIDX];
    pidx = pidx < 0 ? 0 : (pidx >= a.P ? a.P - 1 : pidx);
    const uint8_t* rec = a.bank + (size_t)pidx * d.rec_bytes;
    const uint16_t* lits = reinterpret_cast<const uint16_t*>(rec);
    const uint32_t* mflat = reinterpret_cast<const uint32_t*>(rec + d.lits_bytes);
    if (a.assign)
        for (int v = tid; v < d.n; v += nt) a.assign[(size_t)e * d.n + v] = (st[v >> 5] >> (v & 31)) & 1u;
    if (a.sat)
        for (int c = tid; c < d.m; c += nt) {
            bool sat = false;
            for (int j = 0; j < d.k; ++j) {
                const uint32_t code = lits[lit_index(d.m, c, j)];
                if (code != LIT_PAD) {
                    const uint32_t v = code >> 1;
                    sat |= (((st[v >> 5] >> (v & 31)) ^ code) & 1u) != 0u;
                }
            }
            a.sat[(size_t)e * d.m + c] = sat ? 1 : 0;
        }
    if (tid == 0) {
        if (a.num_unsat) a.num_unsat[e] = (int)tail[ST_NUNSAT];
        if (a.step) a.step[e] = (int)tail[ST_STEP];
        if (a.pidx) a.pidx[e] = pidx;
    }
    if (a.done)
        for (int i = tid; i < d.A; i += nt) a.done[(size_t)e * d.A + i] = (uint8_t)(tail[ST_FLAGS] & 1u);
    if (a.clauses || a.l2a)
        for (int i = tid; i < d.m * d.k; i += nt) {
            const uint32_t code = lits[lit_index(d.m, i / d.k, i % d.k)];
            const int v = (code == LIT_PAD) ? -1 : (int)(code >> 1);
            if (a.clauses) a.clauses[(size_t)e * d.m * d.k + i] = v < 0 ? 0 : ((code & 1u) ? -(v + 1) : (v + 1));
            // env:160: index -1 wraps to the last variable
            if (a.l2a) a.l2a[(size_t)e * d.m * d.k + i] = var_to_agent(d, v < 0 ? d.n - 1 : v);
        }
    if (a.acm)
        for (int i = tid; i < d.A * d.m; i += nt) {
            const int ag = i / d.m, c = i - ag * d.m;
            const int j = ag * d.D + d.n + c;
            a.acm[(size_t)e * d.A * d.m + i] = ((mflat[j >> 5] >> (j & 31)) & 1u) ? 1 : -1;
        }
    if (a.anm)
        for (int i = tid; i < d.A * d.n; i += nt) {
            const int ag = i / d.n, v = i - ag * d.n;
            const int j = ag * d.D + d.n + d.m + v;
            a.anm[(size_t)e * d.A * d.n + i] = ((mflat[j >> 5] >> (j & 31)) & 1u) ? 1 : -1;
        }
}

cudaError_t launch_export(const msat_plan* plan, const ExportArgs& a, cudaStream_t s) {
    if (a.B == 0) return cudaSuccess;
    export_kernel<<<a.B, 128, 0, s>>>(plan->d, a);
    return cudaGetLastError();
}

}  // namespace msat
