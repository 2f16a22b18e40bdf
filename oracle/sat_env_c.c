/*
 * Plain-C restatement of the reference SATEnv hot path (TEST ORACLE / CPU BASELINE ONLY).
 *
 * Second, independent restatement next to the NumPy one (oracle/sat_env.py); the two are
 * cross-checked against each other in tests/test_oracle_c.py.  It follows
 * /root/reference/src/envs/multi_agent_sat_env.py (env:LINE) and the rollout step of
 * /root/reference/src/learners/mappo_gnn_sat_learner.py:418-464 (learner:LINE) with the
 * reference's own structure: every env is stepped, every env is reset on its newly drawn
 * formula, and every state/observation leaf is then selected by done (no incremental
 * shortcuts), so that it can serve as the compiled multi-threaded CPU baseline of bench.py.
 * Pinned by the reference-generated fixtures tests/golden/env_*.npz (tests/test_golden_env.py); the PRNG follows
 * oracle/threefry.py (pinned by known-answer vectors).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/liboracle_c.so oracle/sat_env_c.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- Threefry-2x32 (jax 0.4.29 default PRNG) ---------------------------------------- */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t* x0, uint32_t* x1) {
    static const int RA[4] = {13, 15, 26, 6}, RB[4] = {17, 29, 16, 24};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    uint32_t a = *x0 + ks[0], b = *x1 + ks[1];
    for (int g = 0; g < 5; ++g) {
        const int* R = (g % 2 == 0) ? RA : RB;
        for (int i = 0; i < 4; ++i) { a += b; b = rotl32(b, R[i]); b ^= a; }
        a += ks[(g + 1) % 3];
        b += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    *x0 = a; *x1 = b;
}

/* element i of threefry_2x32(key, arange(N)) with the odd-length zero pad */
static uint32_t bits32_at(const uint32_t key[2], uint32_t N, uint32_t i) {
    uint32_t half = (N + 1u) >> 1;
    int lo = i < half;
    uint32_t blk = lo ? i : i - half;
    uint32_t x0 = blk, x1 = (half + blk < N) ? half + blk : 0u;
    threefry2x32(key[0], key[1], &x0, &x1);
    return lo ? x0 : x1;
}

static void split2(const uint32_t key[2], uint32_t a[2], uint32_t b[2]) {
    a[0] = bits32_at(key, 4, 0); a[1] = bits32_at(key, 4, 1);
    b[0] = bits32_at(key, 4, 2); b[1] = bits32_at(key, 4, 3);
}

/* ---- static env description (env:29-97, 286-338) --------------------------------------- */
typedef struct {
    int n, m, k, A, V, max_steps, action_mode;
    const int32_t* agent_vars;   /* [A][V], -1 padded (env:61) */
    const int32_t* var2agent;    /* [n]                (env:92-97) */
} EnvDesc;

static inline int jax_index(int idx, int size) {   /* x[idx]: wrap negatives once, then clamp */
    if (idx < 0) idx += size;
    return idx < 0 ? 0 : (idx >= size ? size - 1 : idx);
}

/* env:130-156 */
static int satisfaction(const EnvDesc* d, const int32_t* assign, const int32_t* cl, uint8_t* status) {
    int unsat = 0;
    for (int c = 0; c < d->m; ++c) {
        int sat = 0;
        for (int j = 0; j < d->k; ++j) {
            int lit = cl[c * d->k + j];
            int a = assign[jax_index(abs(lit) - 1, d->n)];
            sat |= (lit > 0 && a == 1) || (lit < 0 && a == 0);
        }
        status[c] = (uint8_t)sat;
        unsat += !sat;
    }
    return unsat;
}

/* env:99-128, including the comparison against the -1 padded agent_vars rows */
static void observation_maps(const EnvDesc* d, const int32_t* cl, int32_t* acm, int32_t* anm, uint8_t* present) {
    for (int a = 0; a < d->A; ++a) {
        const int32_t* row = d->agent_vars + a * d->V;
        memset(present, 0, (size_t)d->n);
        for (int c = 0; c < d->m; ++c) {
            int related = 0;
            for (int j = 0; j < d->k && !related; ++j) {
                int vi = abs(cl[c * d->k + j]) - 1;
                for (int s = 0; s < d->V; ++s)
                    if (vi == row[s]) { related = 1; break; }
            }
            acm[a * d->m + c] = related ? 1 : -1;
            if (related)
                for (int j = 0; j < d->k; ++j) {
                    int vi = abs(cl[c * d->k + j]) - 1;
                    if (vi >= 0 && vi < d->n) present[vi] = 1;
                }
        }
        for (int v = 0; v < d->n; ++v) {
            int own = 0;
            for (int s = 0; s < d->V; ++s) own |= (row[s] == v);
            anm[a * d->n + v] = (present[v] && !own) ? 1 : -1;
        }
    }
}

/* env:345-398 */
static void get_obs(const EnvDesc* d, const int32_t* assign, const uint8_t* status, const int32_t* acm,
                    const int32_t* anm, int32_t* obs) {
    const int D = 2 * d->n + d->m;
    for (int a = 0; a < d->A; ++a) {
        int32_t* o = obs + (size_t)a * D;
        for (int v = 0; v < d->n; ++v) o[v] = (d->var2agent[v] == a) ? assign[v] : -1;
        for (int c = 0; c < d->m; ++c) o[d->n + c] = (acm[a * d->m + c] == 1) ? (status[c] ? 1 : 0) : -1;
        for (int v = 0; v < d->n; ++v) {
            int nm = anm[a * d->n + v];
            o[d->n + d->m + v] = (nm != -1) ? nm * assign[v] : -1;
        }
    }
}

/* env:158-181 for one env */
static void reset_one(const EnvDesc* d, const int32_t* cl, const uint32_t key[2], int32_t* assign, uint8_t* status,
                      int32_t* nunsat, int32_t* step, uint8_t* done, int32_t* clauses, int32_t* acm, int32_t* anm,
                      int32_t* l2a, int32_t* obs, uint8_t* scratch) {
    uint32_t k1[2], k2[2];
    split2(key, k1, k2);
    for (int v = 0; v < d->n; ++v) assign[v] = (int32_t)(bits32_at(k2, (uint32_t)d->n, (uint32_t)v) & 1u);   /* env:162 */
    memcpy(clauses, cl, sizeof(int32_t) * (size_t)d->m * d->k);
    for (int i = 0; i < d->m * d->k; ++i) l2a[i] = d->var2agent[jax_index(abs(cl[i]) - 1, d->n)];          /* env:160 */
    observation_maps(d, cl, acm, anm, scratch);
    *nunsat = satisfaction(d, assign, cl, status);
    *step = 0;
    memset(done, 0, (size_t)d->A);
    get_obs(d, assign, status, acm, anm, obs);
}

/* ---- batched entry points (ctypes) ----------------------------------------------------------- */
/* All arrays carry a leading B axis, C-contiguous, same dtypes as oracle/sat_env.py's SATState. */

int oracle_c_reset(int n, int m, int k, int A, int V, int max_steps, int action_mode, const int32_t* agent_vars,
                   const int32_t* var2agent, int B, const int32_t* clauses_in, const uint32_t* keys, int32_t* assign,
                   uint8_t* status, int32_t* nunsat, int32_t* step, uint8_t* done, int32_t* clauses, int32_t* acm,
                   int32_t* anm, int32_t* l2a, int32_t* obs) {
    EnvDesc d = {n, m, k, A, V, max_steps, action_mode, agent_vars, var2agent};
    const size_t D = (size_t)2 * n + m;
#pragma omp parallel
    {
        uint8_t* scratch = (uint8_t*)malloc((size_t)n + 1);
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b)
            reset_one(&d, clauses_in + (size_t)b * m * k, keys + 2 * (size_t)b, assign + (size_t)b * n,
                      status + (size_t)b * m, nunsat + b, step + b, done + (size_t)b * A,
                      clauses + (size_t)b * m * k, acm + (size_t)b * A * m, anm + (size_t)b * A * n,
                      l2a + (size_t)b * m * k, obs + (size_t)b * A * D, scratch);
        free(scratch);
    }
    return 0;
}

/* One rollout step structured like the reference (learner:418-464): step_env for every env (env:225-284),
 * reset of every env on problems[new_idx[b]] with reset_keys[b] when auto_reset != 0, then a per-leaf
 * select by done.  The state arrays are updated in place; reward/done_all/solved/nunsat_out/episode_step
 * are the pre-reset values (learner:467-478).  With auto_reset == 0 this is plain vmapped step_env. */
int oracle_c_step(int n, int m, int k, int A, int V, int max_steps, int action_mode, const int32_t* agent_vars,
                  const int32_t* var2agent, int B, const int32_t* actions, int auto_reset, int P,
                  const int32_t* problems, const int32_t* new_idx, const uint32_t* reset_keys, int32_t* assign,
                  uint8_t* status, int32_t* nunsat, int32_t* step, uint8_t* done, int32_t* clauses, int32_t* acm,
                  int32_t* anm, int32_t* l2a, int32_t* obs, float* reward, uint8_t* done_all, uint8_t* solved_out,
                  int32_t* nunsat_out, int32_t* episode_step) {
    EnvDesc d = {n, m, k, A, V, max_steps, action_mode, agent_vars, var2agent};
    const size_t D = (size_t)2 * n + m, mk = (size_t)m * k;
    (void)P;
#pragma omp parallel
    {
        /* private copies of every leaf of the freshly reset env (the reference materialises them all) */
        int32_t* r_assign = (int32_t*)malloc(sizeof(int32_t) * n);
        uint8_t* r_status = (uint8_t*)malloc((size_t)m);
        uint8_t* r_done = (uint8_t*)malloc((size_t)A);
        int32_t* r_clauses = (int32_t*)malloc(sizeof(int32_t) * mk);
        int32_t* r_acm = (int32_t*)malloc(sizeof(int32_t) * (size_t)A * m);
        int32_t* r_anm = (int32_t*)malloc(sizeof(int32_t) * (size_t)A * n);
        int32_t* r_l2a = (int32_t*)malloc(sizeof(int32_t) * mk);
        int32_t* r_obs = (int32_t*)malloc(sizeof(int32_t) * (size_t)A * D);
        uint8_t* scratch = (uint8_t*)malloc((size_t)n + 1);
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            int32_t* as = assign + (size_t)b * n;
            const int32_t* cl = clauses + (size_t)b * mk;
            /* --- flips (env:233-250) --- */
            if (action_mode == 0) {
                const int32_t* act = actions + (size_t)b * A;
                for (int a = 0; a < A; ++a) {
                    int nv = 0;
                    for (int s = 0; s < V; ++s) nv += agent_vars[a * V + s] >= 0;   /* sum(action_mask[a]) */
                    if (act[a] >= nv) continue;                                      /* no-op, env:236 */
                    int safe = act[a] < nv - 1 ? act[a] : nv - 1;                    /* env:238 */
                    int var = agent_vars[a * V + jax_index(safe, V)];
                    if (var >= 0 && var < n) as[var] = (as[var] != 0) ^ 1;           /* one_hot + logical_xor */
                }
            } else {
                const int32_t* act = actions + (size_t)b * A * V;
                for (int a = 0; a < A; ++a)
                    for (int s = 0; s < V; ++s) {
                        int var = agent_vars[a * V + s];
                        if (var >= 0) as[var] ^= act[a * V + s];                     /* env:246-250 */
                    }
            }
            /* --- satisfaction, done, reward, info (env:252-282) --- */
            int nu = satisfaction(&d, as, cl, status + (size_t)b * m);
            int is_solved = nu == 0;
            int is_done = is_solved || (step[b] + 1 >= max_steps);
            nunsat[b] = nu;
            episode_step[b] = step[b] + 1;
            step[b] += 1;
            memset(done + (size_t)b * A, is_done, (size_t)A);
            for (int a = 0; a < A; ++a) reward[(size_t)b * A + a] = is_solved ? 1.0f : 0.0f;
            done_all[b] = (uint8_t)is_done;
            solved_out[b] = (uint8_t)is_solved;
            nunsat_out[b] = nu;
            get_obs(&d, as, status + (size_t)b * m, acm + (size_t)b * A * m, anm + (size_t)b * A * n,
                    obs + (size_t)b * A * D);
            if (!auto_reset) continue;
            /* --- reset EVERY env on its new formula, then select by done (learner:431-464) --- */
            int32_t r_nunsat, r_step;
            reset_one(&d, problems + (size_t)new_idx[b] * mk, reset_keys + 2 * (size_t)b, r_assign, r_status, &r_nunsat,
                      &r_step, r_done, r_clauses, r_acm, r_anm, r_l2a, r_obs, scratch);
            if (is_done) {
                memcpy(as, r_assign, sizeof(int32_t) * n);
                memcpy(status + (size_t)b * m, r_status, (size_t)m);
                nunsat[b] = r_nunsat;
                step[b] = r_step;
                memcpy(done + (size_t)b * A, r_done, (size_t)A);
                memcpy(clauses + (size_t)b * mk, r_clauses, sizeof(int32_t) * mk);
                memcpy(acm + (size_t)b * A * m, r_acm, sizeof(int32_t) * (size_t)A * m);
                memcpy(anm + (size_t)b * A * n, r_anm, sizeof(int32_t) * (size_t)A * n);
                memcpy(l2a + (size_t)b * mk, r_l2a, sizeof(int32_t) * mk);
                memcpy(obs + (size_t)b * A * D, r_obs, sizeof(int32_t) * (size_t)A * D);
            }
        }
        free(r_assign); free(r_status); free(r_done); free(r_clauses); free(r_acm); free(r_anm); free(r_l2a);
        free(r_obs); free(scratch);
    }
    return 0;
}

/* rollout key chain (learner:397,416,426-434): chain[10] = {rng', act, step, prob, reset};
 * new_idx[B] = randint(prob_key, (B,), 0, P); reset_keys[B][2] = split(reset_key, B) */
int oracle_c_rollout_keys(const uint32_t rng[2], int B, int P, uint32_t chain[10], int32_t* new_idx, uint32_t* reset_keys) {
    uint32_t r[2] = {rng[0], rng[1]}, a[2], b[2];
    split2(r, a, b); chain[2] = b[0]; chain[3] = b[1]; r[0] = a[0]; r[1] = a[1];
    split2(r, a, b); chain[4] = b[0]; chain[5] = b[1]; r[0] = a[0]; r[1] = a[1];
    uint32_t w[6];
    for (uint32_t i = 0; i < 6; ++i) w[i] = bits32_at(r, 6, i);
    chain[0] = w[0]; chain[1] = w[1]; chain[6] = w[2]; chain[7] = w[3]; chain[8] = w[4]; chain[9] = w[5];
    uint32_t k1[2], k2[2];
    split2(&chain[6], k1, k2);
    uint32_t span = P > 0 ? (uint32_t)P : 1u, mult = 65536u % span;
    mult = (mult * mult) % span;
    for (int i = 0; i < B; ++i) {
        uint32_t hi = bits32_at(k1, (uint32_t)B, (uint32_t)i), lo = bits32_at(k2, (uint32_t)B, (uint32_t)i);
        new_idx[i] = (int32_t)(((hi % span) * mult + (lo % span)) % span);
        reset_keys[2 * i] = bits32_at(&chain[8], 2u * (uint32_t)B, 2u * (uint32_t)i);
        reset_keys[2 * i + 1] = bits32_at(&chain[8], 2u * (uint32_t)B, 2u * (uint32_t)i + 1u);
    }
    return 0;
}

void oracle_c_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_c_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
