/*
 * oea_oracle.c — CPU ORACLE (test infrastructure; NOT part of the product path).
 *
 * Plain-C restatement of the reference's algorithm for hot path (i): the TensorFlow-1 graph that
 * nju-websoft/OpenEA builds for negative-sampled triple scoring.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * PARITY: pinned to the reference's own graph code — its _define_* methods and session.run executed on a TF-1 graph
 * interpreter (oracle/tf1_shim.py) give tests/golden/path_i_reference_graphs.npz, which this oracle reproduces
 * (tests/test_reference_graph_goldens.py::test_c_oracle_reproduces_the_reference_graph).  UNPINNED remains what that
 * interpreter itself restates, TF's op and optimiser semantics: the arithmetic lives in TensorFlow 1.x (tested 1.8/1.12 per the
 * reference README.md:110; un-pinned in setup.py:13), which is absent from /root/reference and cannot be
 * installed here (no cp312 wheel, no network).  The reference ships no tests or golden vectors for this
 * path.  This file restates TF's documented semantics at the reference's own call sites:
 *   - tf.nn.l2_normalize(whole_table, 1)            modules/base/initializers.py:26,34,41,50
 *   - tf.nn.embedding_lookup of h, r, t rows        models/basic_model.py:88-94
 *   - score / loss                                  modules/base/losses.py:15-73, approaches/bootea.py:197
 *   - compute_gradients (dense w.r.t. the variable) modules/base/optimizers.py:4-7
 *   - Adagrad / Adam / SGD apply                    modules/base/optimizers.py:10-20
 * and is itself cross-checked against float64 torch-autograd of the same formulas in
 * tests/test_oracle_triple.py (an independent derivation of the backward pass).
 *
 * It deliberately keeps TF's DENSE structure (normalise the whole table, dense gradient, dense optimiser
 * pass) so that (a) it times what the TF-CPU reference executes per step and (b) tests can show that the
 * engine's row-sparse update equals the dense one on every row.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { SCORE_L1 = 0, SCORE_L2SQ = 1 };
enum { LOSS_MARGIN = 0, LOSS_LIMITED = 1, LOSS_LOGISTIC = 2, LOSS_POSITIVE = 3, LOSS_LOGSIGMOID = 4 };
enum { OPT_SGD = 0, OPT_ADAGRAD = 1, OPT_ADAM = 2 };

static const float NORM_EPS = 1e-12f; /* tf.nn.l2_normalize default epsilon */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* y = x * rsqrt(max(sum(x^2), eps)) for every row (initializers.py:26: tf.nn.l2_normalize(embeddings, 1)). */
void orc_l2_normalize_rows(const float* x, int rows, int d, float* y, float* inv_out, float* ss_out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; ++i) {
        const float* xi = x + (size_t)i * d;
        float ss = 0.f;
        for (int c = 0; c < d; ++c) ss += xi[c] * xi[c];
        const float inv = 1.0f / sqrtf(ss > NORM_EPS ? ss : NORM_EPS);
        float* yi = y + (size_t)i * d;
        for (int c = 0; c < d; ++c) yi[c] = xi[c] * inv;
        if (inv_out) inv_out[i] = inv;
        if (ss_out) ss_out[i] = ss;
    }
}

static inline float sgnf(float v) { return (float)(v > 0.f) - (float)(v < 0.f); } /* tf.sign(0) = 0 */
static inline float softplusf(float v) { return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v))); }

/* per-triple loss and d loss / d score (losses.py:26,53-55,70-72; bootea.py:197); relu'(0) = 0 */
static inline void loss_of(int kind, int is_neg, float s, float margin, float neg_margin, float balance,
                           float* L, float* g) {
    switch (kind) {
        case LOSS_LIMITED:
            if (!is_neg) { *L = fmaxf(s - margin, 0.f); *g = s > margin ? 1.f : 0.f; }
            else { *L = balance * fmaxf(neg_margin - s, 0.f); *g = s < neg_margin ? -balance : 0.f; }
            break;
        case LOSS_LOGISTIC:
            if (!is_neg) { *L = softplusf(s); *g = 1.f / (1.f + expf(-s)); }
            else { *L = softplusf(-s); *g = -1.f / (1.f + expf(s)); }
            break;
        case LOSS_LOGSIGMOID:
            *L = softplusf(s); *g = 1.f / (1.f + expf(-s));
            break;
        default:
            *L = s; *g = 1.f;
            break;
    }
}

/*
 * Forward + backward of the triple-scoring graph, dense as TF builds it.
 *   ent_w [N,d], rel_w [R,d]: raw variables.  g_ent [N,d], g_rel [R,d]: d loss / d variable (overwritten).
 * Returns the batch loss (sum).  scores_out (optional) receives n_pos + n_neg scores.
 */
double orc_triple_fwd_bwd(const float* ent_w, int N, const float* rel_w, int R, int d, int ent_norm, int rel_norm,
                          const int32_t* ph, const int32_t* pr, const int32_t* pt, int n_pos,
                          const int32_t* nh, const int32_t* nr, const int32_t* nt, int n_neg,
                          int score_kind, int loss_kind, float margin, float neg_margin, float balance,
                          float* g_ent, float* g_rel, float* scores_out) {
    const int total = n_pos + n_neg;
    float* ent_hat = (float*)malloc((size_t)N * d * sizeof(float));
    float* rel_hat = (float*)malloc((size_t)R * d * sizeof(float));
    float* ent_inv = (float*)malloc((size_t)N * sizeof(float));
    float* rel_inv = (float*)malloc((size_t)R * sizeof(float));
    float* ent_ss = (float*)malloc((size_t)N * sizeof(float));
    float* rel_ss = (float*)malloc((size_t)R * sizeof(float));
    float* score = (float*)malloc((size_t)(total > 0 ? total : 1) * sizeof(float));
    float* gscore = (float*)calloc((size_t)(total > 0 ? total : 1), sizeof(float));
    /* the table-level l2_normalize TF evaluates every step */
    if (ent_norm) orc_l2_normalize_rows(ent_w, N, d, ent_hat, ent_inv, ent_ss);
    else { memcpy(ent_hat, ent_w, (size_t)N * d * sizeof(float)); for (int i = 0; i < N; ++i) { ent_inv[i] = 1.f; ent_ss[i] = 1.f; } }
    if (rel_norm) orc_l2_normalize_rows(rel_w, R, d, rel_hat, rel_inv, rel_ss);
    else { memcpy(rel_hat, rel_w, (size_t)R * d * sizeof(float)); for (int i = 0; i < R; ++i) { rel_inv[i] = 1.f; rel_ss[i] = 1.f; } }

    /* scores */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < total; ++i) {
        const int neg = i >= n_pos, j = neg ? i - n_pos : i;
        const float* h = ent_hat + (size_t)(neg ? nh[j] : ph[j]) * d;
        const float* r = rel_hat + (size_t)(neg ? nr[j] : pr[j]) * d;
        const float* t = ent_hat + (size_t)(neg ? nt[j] : pt[j]) * d;
        float s = 0.f;
        for (int c = 0; c < d; ++c) {
            const float u = h[c] + r[c] - t[c];
            s += score_kind == SCORE_L1 ? fabsf(u) : u * u;
        }
        score[i] = s;
    }
    /* loss + d loss / d score */
    double loss = 0.0;
    if (loss_kind == LOSS_MARGIN) { /* Σ relu(m + s⁺_i − s⁻_i), n_neg == n_pos */
        for (int i = 0; i < n_pos; ++i) {
            const float v = margin + score[i] - score[n_pos + i];
            if (v > 0.f) { loss += v; gscore[i] = 1.f; gscore[n_pos + i] = -1.f; }
        }
    } else {
        for (int i = 0; i < total; ++i) {
            float L, g;
            loss_of(loss_kind, i >= n_pos, score[i], margin, neg_margin, balance, &L, &g);
            loss += L;
            gscore[i] = g;
        }
    }
    if (scores_out) memcpy(scores_out, score, (size_t)total * sizeof(float));

    /* gradient w.r.t. the NORMALISED tables: scatter-add, row-partitioned across threads (deterministic) */
    float* gh_ent = (float*)calloc((size_t)N * d, sizeof(float));
    float* gh_rel = (float*)calloc((size_t)R * d, sizeof(float));
#pragma omp parallel
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num(), nth = omp_get_num_threads();
#else
        const int tid = 0, nth = 1;
#endif
        for (int i = 0; i < total; ++i) {
            const float g = gscore[i];
            if (g == 0.f) continue;
            const int neg = i >= n_pos, j = neg ? i - n_pos : i;
            const int hi = neg ? nh[j] : ph[j], ri = neg ? nr[j] : pr[j], ti = neg ? nt[j] : pt[j];
            const int mine_h = hi % nth == tid, mine_r = ri % nth == tid, mine_t = ti % nth == tid;
            if (!(mine_h || mine_r || mine_t)) continue;
            const float* h = ent_hat + (size_t)hi * d;
            const float* r = rel_hat + (size_t)ri * d;
            const float* t = ent_hat + (size_t)ti * d;
            for (int c = 0; c < d; ++c) {
                const float u = h[c] + r[c] - t[c];
                const float du = g * (score_kind == SCORE_L1 ? sgnf(u) : 2.f * u);
                if (mine_h) gh_ent[(size_t)hi * d + c] += du;
                if (mine_r) gh_rel[(size_t)ri * d + c] += du;
                if (mine_t) gh_ent[(size_t)ti * d + c] -= du;
            }
        }
    }
    /* back through the table-level normalisation: dx = (dy − y·<y,dy>)·inv  (or dy·inv when Σx² < eps) */
    for (int pass = 0; pass < 2; ++pass) {
        const int rows = pass ? R : N, on = pass ? rel_norm : ent_norm;
        const float* hat = pass ? rel_hat : ent_hat;
        const float* inv = pass ? rel_inv : ent_inv;
        const float* ss = pass ? rel_ss : ent_ss;
        const float* gh = pass ? gh_rel : gh_ent;
        float* out = pass ? g_rel : g_ent;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < rows; ++i) {
            const float* y = hat + (size_t)i * d;
            const float* gy = gh + (size_t)i * d;
            float* o = out + (size_t)i * d;
            if (!on) { memcpy(o, gy, (size_t)d * sizeof(float)); continue; }
            float dot = 0.f;
            for (int c = 0; c < d; ++c) dot += y[c] * gy[c];
            if (ss[i] < NORM_EPS) dot = 0.f;
            for (int c = 0; c < d; ++c) o[c] = (gy[c] - y[c] * dot) * inv[i];
        }
    }
    free(ent_hat); free(rel_hat); free(ent_inv); free(rel_inv); free(ent_ss); free(rel_ss);
    free(score); free(gscore); free(gh_ent); free(gh_rel);
    return loss;
}

/* Dense optimiser pass (optimizers.py:10-20, TF1 update rules). t is the 1-based Adam step. */
void orc_opt_dense(int kind, float* w, const float* g, float* s1, float* s2, size_t n, float lr,
                   float beta1, float beta2, float eps, int t) {
    if (kind == OPT_ADAGRAD) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; ++i) {
            s1[i] += g[i] * g[i];
            w[i] -= lr * g[i] / sqrtf(s1[i]);
        }
    } else if (kind == OPT_SGD) {
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; ++i) w[i] -= lr * g[i];
    } else {
        const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t)));
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; ++i) {
            s1[i] = beta1 * s1[i] + (1.f - beta1) * g[i];
            s2[i] = beta2 * s2[i] + (1.f - beta2) * g[i] * g[i];
            w[i] -= lr_t * s1[i] / (sqrtf(s2[i]) + eps);
        }
    }
}

/* One full training step as session.run([loss, optimizer]) executes it (basic_model.py:224-230). */
double orc_triple_step(float* ent_w, int N, float* rel_w, int R, int d, int ent_norm, int rel_norm,
                       const int32_t* ph, const int32_t* pr, const int32_t* pt, int n_pos,
                       const int32_t* nh, const int32_t* nr, const int32_t* nt, int n_neg,
                       int score_kind, int loss_kind, float margin, float neg_margin, float balance,
                       int opt_kind, float lr, float* ent_s1, float* ent_s2, float* rel_s1, float* rel_s2, int t) {
    float* g_ent = (float*)malloc((size_t)N * d * sizeof(float));
    float* g_rel = (float*)malloc((size_t)R * d * sizeof(float));
    const double loss = orc_triple_fwd_bwd(ent_w, N, rel_w, R, d, ent_norm, rel_norm, ph, pr, pt, n_pos, nh, nr, nt,
                                           n_neg, score_kind, loss_kind, margin, neg_margin, balance, g_ent, g_rel, 0);
    orc_opt_dense(opt_kind, ent_w, g_ent, ent_s1, ent_s2, (size_t)N * d, lr, 0.9f, 0.999f, 1e-8f, t);
    orc_opt_dense(opt_kind, rel_w, g_rel, rel_s1, rel_s2, (size_t)R * d, lr, 0.9f, 0.999f, 1e-8f, t);
    free(g_ent); free(g_rel);
    return loss;
}
