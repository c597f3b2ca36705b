/*
 * qldpc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU, float64 restatement of the reference's (michelebanfi/qLDPC)
 * BP + OSD decode path.  It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time the
 * CUDA path against it.  Nothing under qldpc_b200/ may import, link or call it.
 *
 * Parity status: PINNED.  tools/make_golden.py runs the unmodified Python
 * reference (imported from /root/reference in the build container) on seeded
 * inputs and stores its outputs under tests/golden/; tests/test_oracle.py checks
 * this file against those vectors bit-for-bit (min-sum hard/LLR/iteration, OSD-0,
 * OSD-w, checks) and within libm-vs-NumPy last-ulp tolerance for sum-product
 * (NumPy's tanh/arctanh are not glibc's; oracle/bp_numpy.py is the bit-exact
 * sum-product restatement).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  The reference works on dense (m, n) float64 arrays; here
 * the same arithmetic is done on the edge list only.  Non-edges contribute
 * exact identities in the reference (x*1.0, x+0.0), so the results are
 * bit-identical PROVIDED the per-variable order of the posterior additions
 * matches NumPy's (sequential for C-ordered arrays, pairwise-8 for the
 * Fortran-ordered Hx of the BB .npz files).  That order is an
 * input here: var_edge0 (iteration 0) and var_edge1 (iterations >= 1) list each
 * variable's edges in the order they must be added, left to right.
 *
 * Build: see oracle/Makefile  (-O2 -ffp-contract=off: no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MIN_SUM 0      /* rework/decoding.py:5-75   performMinSum_Symmetric            */
#define ORC_SUM_PRODUCT 1  /* decoding/beliefPropagation.py:88-144 performBeliefPropagationFast
                              (== loop version :6-85 and rework/decoding.py:77-129)         */
#define ORC_SUM_PRODUCT_SYM 2 /* rework/decoding.py:131-191 performBeliefPropagation_Symmetric */

typedef struct {
    int m, n, E;
    const int32_t *row_ptr;   /* [m+1] CSR, edges of check c are row_ptr[c]..row_ptr[c+1]-1     */
    const int32_t *col_idx;   /* [E]   ascending column index within a row                      */
    const int32_t *var_ptr;   /* [n+1]                                                          */
    const int32_t *var_edge0; /* [E]   edge ids of variable v in ADD ORDER, iteration 0         */
    const int32_t *var_edge1; /* [E]   same for iterations >= 1                                 */
} orc_graph;

/* --------------------------------------------------------------------------
 * One BP decode of one shot.  Returns 1 if the syndrome was matched.
 *   hard[n] int8, llr[n] float64 (posterior `values`), *iters = 0-based exit
 *   iteration (max_iter-1 on failure), as the reference's 4-tuples.
 * work: caller-provided scratch of at least 3*E + 2*m doubles.
 * -------------------------------------------------------------------------- */
static int bp_one(const orc_graph *g, int variant, const uint8_t *synd, const double *prior,
                  int max_iter, double alpha, double damping, double clip,
                  int8_t *hard, double *llr, int *iters, double *work,
                  double *r_dump, int dump_iter)
{
    const int m = g->m, n = g->n, E = g->E;
    double *Q = work, *R = work + E, *T = work + 2 * E; /* T: tanh cache (sum-product) */
    const double CLIP_VAL = 0.9999999;                  /* beliefPropagation.py:110 */
    const double one_minus_d = 1.0 - damping;           /* decoding.py:65 `(1 - damping)` in float64 */

    /* Q = np.where(mask, initialBelief, 0)  (beliefPropagation.py:107, decoding.py:21);
       Q_old == Q at all times between iterations (decoding.py:22,67). */
    for (int c = 0; c < m; ++c)
        for (int e = g->row_ptr[c]; e < g->row_ptr[c + 1]; ++e) Q[e] = prior[g->col_idx[e]];

    int it = 0, ok = 0;
    for (it = 0; it < max_iter; ++it) {
        /* ---------------- horizontal step ---------------- */
        for (int c = 0; c < m; ++c) {
            const int e0 = g->row_ptr[c], e1 = g->row_ptr[c + 1];
            const double ssign = synd[c] ? -1.0 : 1.0; /* (1 - 2*syndrome) */
            if (variant == ORC_MIN_SUM) {
                /* decoding.py:28-55 */
                double sprod = 1.0, min1 = INFINITY, min2 = INFINITY;
                int imin = -1;
                for (int e = e0; e < e1; ++e) {
                    double q = Q[e], a = fabs(q);
                    if (q < 0.0) sprod = -sprod;      /* sign 0 -> +1 (:30) */
                    if (a < min1) { min1 = a; imin = e; } /* np.argmin: first minimum (:41) */
                }
                for (int e = e0; e < e1; ++e)          /* min of the rest (:44-46) */
                    if (e != imin) { double a = fabs(Q[e]); if (a < min2) min2 = a; }
                for (int e = e0; e < e1; ++e) {
                    double q = Q[e];
                    double mag = (fabs(q) == min1) ? min2 : min1;  /* :51-53 compares VALUES */
                    double sg = (q < 0.0) ? -sprod : sprod;       /* r_signs (:35) */
                    /* R_new = alpha * syndrome_sign * r_signs * magnitudes (:55): the first two
                       products are exact (+-alpha), so one rounding: (+-alpha) * mag. */
                    R[e] = ((alpha * ssign) * sg) * mag;
                }
            } else {
                /* beliefPropagation.py:114-126 / decoding.py:156-166 */
                double prod = 1.0;
                for (int e = e0; e < e1; ++e) { T[e] = tanh(Q[e] * 0.5); prod *= T[e]; } /* ascending col */
                for (int e = e0; e < e1; ++e) {
                    double ts = (fabs(T[e]) < 1e-15) ? 1e-15 : T[e];
                    double x = (prod / ts) * ssign;
                    if (x < -CLIP_VAL) x = -CLIP_VAL;
                    if (x > CLIP_VAL) x = CLIP_VAL;
                    R[e] = 2.0 * atanh(x);
                }
            }
        }
        if (r_dump && it == dump_iter) { /* alpha_estimation return paths (decoding.py:58-59,168-169) */
            for (int e = 0; e < E; ++e)
                r_dump[e] = (variant == ORC_MIN_SUM) ? R[e] / alpha : R[e];
            *iters = 0;
            return 0;
        }
        if (variant == ORC_SUM_PRODUCT_SYM)
            for (int e = 0; e < E; ++e) R[e] = R[e] * alpha; /* R_scaled (decoding.py:171) */

        /* ---------------- vertical step ---------------- */
        const int32_t *ve = (it == 0) ? g->var_edge0 : g->var_edge1;
        for (int v = 0; v < n; ++v) {
            const int a0 = g->var_ptr[v], a1 = g->var_ptr[v + 1];
            double s = 0.0;
            if (a1 > a0) { s = R[ve[a0]]; for (int a = a0 + 1; a < a1; ++a) s = s + R[ve[a]]; }
            double val = s + prior[v];                 /* values = R_sum + initialBelief */
            llr[v] = val;
            hard[v] = (val < 0.0) ? 1 : 0;
            for (int a = a0; a < a1; ++a) {
                int e = ve[a];
                double qn = val - R[e];                /* Q_new = values - R */
                if (variant == ORC_SUM_PRODUCT) { Q[e] = qn; continue; }
                double q = damping * qn + one_minus_d * Q[e]; /* decoding.py:65 / :179 */
                if (q < -clip) q = -clip;              /* np.clip (:66 / :181) */
                if (q > clip) q = clip;
                Q[e] = q;
            }
        }
        /* ---------------- syndrome check ---------------- */
        ok = 1;
        for (int c = 0; c < m && ok; ++c) {
            int par = 0;
            for (int e = g->row_ptr[c]; e < g->row_ptr[c + 1]; ++e) par ^= hard[g->col_idx[e]];
            if (par != (synd[c] & 1)) ok = 0;
        }
        /* `and not alpha_estimation` (decoding.py:72,188): no early exit while dumping messages */
        if (ok && !r_dump) { *iters = it; return 1; }
    }
    *iters = max_iter - 1;
    return 0;
}

int orc_bp_decode(int variant, int m, int n, const int32_t *row_ptr, const int32_t *col_idx,
                  const int32_t *var_ptr, const int32_t *var_edge0, const int32_t *var_edge1,
                  const uint8_t *synd, const double *prior, int max_iter, double alpha,
                  double damping, double clip, int8_t *hard, double *llr, int *iters)
{
    orc_graph g = {m, n, row_ptr[m], row_ptr, col_idx, var_ptr, var_edge0, var_edge1};
    double *work = (double *)malloc(sizeof(double) * (3 * (size_t)g.E + 8));
    int ok = bp_one(&g, variant, synd, prior, max_iter, alpha, damping, clip, hard, llr, iters, work, NULL, -1);
    free(work);
    return ok;
}

/* alpha_estimation=True return value: check-to-variable messages per edge (CSR edge order):
   min-sum R_new/alpha after the first check pass (decoding.py:58-59), sum-product-symmetric
   R at currentIter == 10 (decoding.py:168-169). */
void orc_bp_alpha_messages(int variant, int m, int n, const int32_t *row_ptr, const int32_t *col_idx,
                           const int32_t *var_ptr, const int32_t *var_edge0, const int32_t *var_edge1,
                           const uint8_t *synd, const double *prior, int max_iter, double alpha,
                           double damping, double clip, double *r_edges)
{
    orc_graph g = {m, n, row_ptr[m], row_ptr, col_idx, var_ptr, var_edge0, var_edge1};
    double *work = (double *)malloc(sizeof(double) * (3 * (size_t)g.E + 8));
    int8_t *hard = (int8_t *)malloc(n);
    double *llr = (double *)malloc(sizeof(double) * n);
    int iters;
    int dump_iter = (variant == ORC_MIN_SUM) ? 0 : 10;
    memset(r_edges, 0, sizeof(double) * g.E);
    bp_one(&g, variant, synd, prior, max_iter, alpha, damping, clip, hard, llr, &iters, work, r_edges, dump_iter);
    free(work); free(hard); free(llr);
}

/* --------------------------------------------------------------------------
 * NumPy's pairwise summation of a contiguous float64 vector (np.sum on 1-D /
 * along the contiguous axis): umath loops_utils.h DOUBLE_pairwise_sum.  Needed
 * for compute_metric's np.sum(solution * np.abs(llr)) (OSD_enhanced.py:174).
 * -------------------------------------------------------------------------- */
static double np_pairwise_sum(const double *a, long n)
{
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}
double orc_np_pairwise_sum(const double *a, long n) { return np_pairwise_sum(a, n); }

/* --------------------------------------------------------------------------
 * OSD.  H is given dense, row-major uint8 (m x n).  Rows are packed into
 * uint64 words in PERMUTED column order, with the residual syndrome kept in a
 * separate array b[], and the reference's Gauss-Jordan is replayed step by
 * step, including the row swaps (OSD.py:31-72 == OSD_enhanced.py:180-224).
 *
 * Ordering contract: STABLE ascending sort of |llr| (ties ->
 * lower column index).  The reference's np.argsort(kind='quicksort') order on
 * ties is unspecified; parity with it is tested on tie-free inputs.
 * -------------------------------------------------------------------------- */
typedef struct { double key; int idx; } kv;
static int kv_cmp(const void *a, const void *b)
{
    const kv *x = (const kv *)a, *y = (const kv *)b;
    /* NaN sorts last, as in np.argsort */
    int xn = isnan(x->key), yn = isnan(y->key);
    if (xn || yn) { if (xn != yn) return xn - yn; return x->idx - y->idx; }
    if (x->key < y->key) return -1;
    if (x->key > y->key) return 1;
    return x->idx - y->idx;
}

static inline int getbit(const uint64_t *row, int j) { return (int)((row[j >> 6] >> (j & 63)) & 1u); }

typedef struct {
    int m, n, W;
    uint64_t *A;      /* m x W reduced rows (permuted column order)          */
    uint64_t *A0;     /* m x W UNREDUCED permuted rows (recompute_solution)  */
    uint8_t *b;       /* m reduced residual syndrome                         */
    int *ordering;    /* n   permuted position -> original column            */
    int *piv_row, *piv_col, npiv;
    uint8_t *e_perm;  /* n */
} osd_ws;

static void osd_ws_alloc(osd_ws *w, int m, int n)
{
    w->m = m; w->n = n; w->W = (n + 63) / 64;
    w->A = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m * w->W);
    w->A0 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)m * w->W);
    w->b = (uint8_t *)malloc(m);
    w->ordering = (int *)malloc(sizeof(int) * n);
    w->piv_row = (int *)malloc(sizeof(int) * (m + 1));
    w->piv_col = (int *)malloc(sizeof(int) * (m + 1));
    w->e_perm = (uint8_t *)malloc(n);
}
static void osd_ws_free(osd_ws *w)
{
    free(w->A); free(w->A0); free(w->b); free(w->ordering); free(w->piv_row); free(w->piv_col); free(w->e_perm);
}

/* syndrome of a candidate x (uint8[n]) under dense H; returns number of checks that differ from synd */
static int synd_mismatch(int m, int n, const uint8_t *H, const uint8_t *x, const uint8_t *synd)
{
    int bad = 0;
    for (int r = 0; r < m; ++r) {
        int par = 0;
        const uint8_t *h = H + (size_t)r * n;
        for (int j = 0; j < n; ++j) par ^= (h[j] & x[j]);
        bad += (par != (synd[r] & 1));
    }
    return bad;
}

/* compute_metric (OSD_enhanced.py:158-177) */
static double osd_metric(int m, int n, const uint8_t *H, const uint8_t *sol, const double *llr,
                         const uint8_t *synd, double *tmp)
{
    int sw = synd_mismatch(m, n, H, sol, synd);
    double metric = (sw > 0) ? (1e10 + (double)sw * 1e8) : 0.0;
    for (int j = 0; j < n; ++j) tmp[j] = (double)sol[j] * fabs(llr[j]);
    metric += np_pairwise_sum(tmp, n);
    return metric;
}

/* OSD-0 front half shared by performOSD (OSD.py:3-28) and performOSD_enhanced (:31-55). */
static void osd0_core(osd_ws *w, const uint8_t *H, const uint8_t *synd, const double *llr,
                      const uint8_t *hard, uint8_t *sol)
{
    const int m = w->m, n = w->n, W = w->W;
    /* residual syndrome (OSD.py:7-8) */
    for (int r = 0; r < m; ++r) {
        int par = 0;
        const uint8_t *h = H + (size_t)r * n;
        for (int j = 0; j < n; ++j) par ^= (h[j] & hard[j]);
        w->b[r] = (uint8_t)((synd[r] & 1) ^ par);
    }
    /* ordering = argsort(|llr|) ascending, stable (OSD.py:10-11) */
    kv *keys = (kv *)malloc(sizeof(kv) * n);
    for (int j = 0; j < n; ++j) { keys[j].key = fabs(llr[j]); keys[j].idx = j; }
    qsort(keys, n, sizeof(kv), kv_cmp);
    for (int j = 0; j < n; ++j) w->ordering[j] = keys[j].idx;
    free(keys);
    /* H_permuted = H[:, ordering] (OSD.py:12) */
    memset(w->A, 0, sizeof(uint64_t) * (size_t)m * W);
    for (int r = 0; r < m; ++r)
        for (int j = 0; j < n; ++j)
            if (H[(size_t)r * n + w->ordering[j]]) w->A[(size_t)r * W + (j >> 6)] |= (uint64_t)1 << (j & 63);
    memcpy(w->A0, w->A, sizeof(uint64_t) * (size_t)m * W);

    /* gf2_elimination (OSD.py:31-72) */
    int row = 0;
    w->npiv = 0;
    for (int col = 0; col < n; ++col) {
        if (row >= m) break;                                    /* :43 */
        int pr = -1;
        for (int r = row; r < m; ++r) if (getbit(w->A + (size_t)r * W, col)) { pr = r; break; } /* :47-50 */
        if (pr == -1) continue;
        if (pr != row) {                                        /* swap (:56-58) */
            for (int k = 0; k < W; ++k) { uint64_t t = w->A[(size_t)row * W + k]; w->A[(size_t)row * W + k] = w->A[(size_t)pr * W + k]; w->A[(size_t)pr * W + k] = t; }
            uint8_t t = w->b[row]; w->b[row] = w->b[pr]; w->b[pr] = t;
        }
        w->piv_row[w->npiv] = row; w->piv_col[w->npiv] = col; w->npiv++;
        for (int r = 0; r < m; ++r)                             /* eliminate all other rows (:64-68) */
            if (r != row && getbit(w->A + (size_t)r * W, col)) {
                for (int k = 0; k < W; ++k) w->A[(size_t)r * W + k] ^= w->A[(size_t)row * W + k];
                w->b[r] ^= w->b[row];
            }
        row++;
    }
    /* e_permuted[c] = s_reduced[r]; unpermute; solution = hard ^ e (OSD.py:16-26) */
    memset(w->e_perm, 0, n);
    for (int i = 0; i < w->npiv; ++i) w->e_perm[w->piv_col[i]] = w->b[w->piv_row[i]];
    for (int j = 0; j < n; ++j) sol[w->ordering[j]] = (uint8_t)(hard[w->ordering[j]] ^ w->e_perm[j]);
}

/* performOSD (OSD.py:3-28).  out[n] uint8 */
void orc_osd0(int m, int n, const uint8_t *H, const uint8_t *synd, const double *llr,
              const uint8_t *hard, uint8_t *out)
{
    osd_ws w;
    osd_ws_alloc(&w, m, n);
    osd0_core(&w, H, synd, llr, hard, out);
    osd_ws_free(&w);
}

/* performOSD_enhanced (OSD_enhanced.py:5-131).  max_combinations <= 0 means None.
   Returns: 0 = OSD-0 solution returned by the early exits (:58-64,:71-72), 1 = sweep ran. */
int orc_osd_enhanced(int m, int n, const uint8_t *H, const uint8_t *synd, const double *llr,
                     const uint8_t *hard, int order, long max_combinations, uint8_t *out)
{
    osd_ws w;
    osd_ws_alloc(&w, m, n);
    osd0_core(&w, H, synd, llr, hard, out);
    int swept = 0;
    if (synd_mismatch(m, n, H, out, synd) == 0 || order == 0) goto done;   /* :58-64 */
    {
        const int W = w.W;
        /* non-pivot permuted positions, ascending; the reference re-sorts them by |llr| (:75-77),
           which is the identity under the stable-sort contract (they are already ascending). */
        uint8_t *is_piv = (uint8_t *)calloc(n, 1);
        for (int i = 0; i < w.npiv; ++i) is_piv[w.piv_col[i]] = 1;
        int *np_pos = (int *)malloc(sizeof(int) * n);
        int nnp = 0;
        for (int j = 0; j < n; ++j) if (!is_piv[j]) np_pos[nnp++] = j;
        free(is_piv);
        if (nnp == 0) { free(np_pos); goto done; }                          /* :71-72 */
        int T = nnp < order + 10 ? nnp : order + 10;                         /* :80 */
        swept = 1;

        uint8_t *best = (uint8_t *)malloc(n), *cand = (uint8_t *)malloc(n), *e_full = (uint8_t *)malloc(n);
        double *tmp = (double *)malloc(sizeof(double) * n);
        memcpy(best, out, n);
        double best_metric = osd_metric(m, n, H, out, llr, synd, tmp);       /* :84 */
        int found_valid = 0;                                                 /* :85 (False here) */
        long tested = 0;
        int *comb = (int *)malloc(sizeof(int) * (order + 1));
        int wmax = order < T ? order : T;                                    /* :89 range(1, min(order+1, len+1)) */
        for (int wt = 1; wt <= wmax; ++wt) {
            if (max_combinations > 0 && tested >= max_combinations) break;  /* :91 */
            for (int i = 0; i < wt; ++i) comb[i] = i;                        /* itertools.combinations: lexicographic */
            while (1) {
                if (max_combinations > 0 && tested >= max_combinations) break; /* :96 */
                /* e_test = e_permuted with flips (:100-102) */
                memcpy(e_full, w.e_perm, n);
                for (int i = 0; i < wt; ++i) e_full[np_pos[comb[i]]] ^= 1;
                /* recompute_solution (:134-155): UNREDUCED H_permuted rows, REDUCED syndrome,
                   Gauss-Seidel over the pivots in order */
                for (int i = 0; i < w.npiv; ++i) {
                    int r = w.piv_row[i], c = w.piv_col[i];
                    const uint64_t *hr = w.A0 + (size_t)r * W;
                    int contrib = 0;
                    for (int j = 0; j < n; ++j) if (j != c && getbit(hr, j)) contrib ^= e_full[j];
                    e_full[c] = (uint8_t)(w.b[r] ^ contrib);
                }
                for (int j = 0; j < n; ++j) cand[w.ordering[j]] = (uint8_t)(hard[w.ordering[j]] ^ e_full[j]); /* :109-111 */
                int valid = (synd_mismatch(m, n, H, cand, synd) == 0);      /* :114-115 */
                if (valid && !found_valid) {                                 /* :117-121 */
                    memcpy(best, cand, n);
                    best_metric = osd_metric(m, n, H, cand, llr, synd, tmp);
                    found_valid = 1;
                } else if (valid || !found_valid) {                          /* :122-127 */
                    double tm = osd_metric(m, n, H, cand, llr, synd, tmp);
                    if (tm < best_metric) { memcpy(best, cand, n); best_metric = tm; }
                }
                tested++;
                /* next combination */
                int i = wt - 1;
                while (i >= 0 && comb[i] == T - wt + i) --i;
                if (i < 0) break;
                comb[i]++;
                for (int k = i + 1; k < wt; ++k) comb[k] = comb[k - 1] + 1;
            }
        }
        memcpy(out, best, n);
        free(best); free(cand); free(e_full); free(tmp); free(comb); free(np_pos);
    }
done:
    osd_ws_free(&w);
    return swept;
}

/* debugging / test hook: OSD-0 internals (ordering, pivots, reduced syndrome) */
void orc_osd0_internals(int m, int n, const uint8_t *H, const uint8_t *synd, const double *llr,
                        const uint8_t *hard, uint8_t *out, int32_t *ordering, int32_t *npiv,
                        int32_t *piv_row, int32_t *piv_col, uint8_t *s_reduced)
{
    osd_ws w;
    osd_ws_alloc(&w, m, n);
    osd0_core(&w, H, synd, llr, hard, out);
    for (int j = 0; j < n; ++j) ordering[j] = w.ordering[j];
    *npiv = w.npiv;
    for (int i = 0; i < w.npiv; ++i) { piv_row[i] = w.piv_row[i]; piv_col[i] = w.piv_col[i]; }
    memcpy(s_reduced, w.b, m);
    osd_ws_free(&w);
}

/* --------------------------------------------------------------------------
 * Batch driver == the per-shot body of the reference's Monte-Carlo loops
 * (paperResults.py:57-100, rework/main.py:75-112): BP, OSD on BP failure,
 * then the caller-visible per-shot outputs.  Shots are independent; OpenMP
 * over shots gives the "all host cores" CPU baseline.
 *   synd   [B][m] uint8         corr [B][n] uint8 (final detection)
 *   conv   [B]    uint8 (BP converged)      iters [B] int32
 *   llr_out [B][n] float64 or NULL
 *   osd_order < 0: BP only (no OSD);  0: performOSD;  > 0: performOSD_enhanced(order)
 * -------------------------------------------------------------------------- */
void orc_decode_batch(int variant, int m, int n, const int32_t *row_ptr, const int32_t *col_idx,
                      const int32_t *var_ptr, const int32_t *var_edge0, const int32_t *var_edge1,
                      const uint8_t *Hdense, long B, const uint8_t *synd, const double *prior,
                      int max_iter, double alpha, double damping, double clip, int osd_order,
                      uint8_t *corr, uint8_t *conv, int32_t *iters, double *llr_out, int nthreads)
{
    orc_graph g = {m, n, row_ptr[m], row_ptr, col_idx, var_ptr, var_edge0, var_edge1};
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * (3 * (size_t)g.E + 8));
        double *llr = (double *)malloc(sizeof(double) * n);
        int8_t *hard = (int8_t *)malloc(n);
        uint8_t *sol = (uint8_t *)malloc(n);
#pragma omp for schedule(dynamic, 16)
        for (long i = 0; i < B; ++i) {
            int it;
            int ok = bp_one(&g, variant, synd + i * m, prior, max_iter, alpha, damping, clip,
                            hard, llr, &it, work, NULL, -1);
            conv[i] = (uint8_t)ok;
            iters[i] = it;
            if (!ok && osd_order >= 0) {
                if (osd_order == 0) orc_osd0(m, n, Hdense, synd + i * m, llr, (const uint8_t *)hard, sol);
                else orc_osd_enhanced(m, n, Hdense, synd + i * m, llr, (const uint8_t *)hard, osd_order, 0, sol);
                memcpy(corr + i * n, sol, n);
            } else {
                memcpy(corr + i * n, hard, n);
            }
            if (llr_out) memcpy(llr_out + i * n, llr, sizeof(double) * n);
        }
        free(work); free(llr); free(hard); free(sol);
    }
}

/* Per-shot checks of the drivers (paperResults.py:83-100, rework/main.py:90-112):
   logical[i] = any(L @ (corr ^ err) % 2), valid[i] = (H @ corr % 2 == synd), weight[i] = sum(corr ^ err) */
void orc_check_batch(int m, int n, int k, const uint8_t *Hdense, const uint8_t *L, long B,
                     const uint8_t *err, const uint8_t *corr, const uint8_t *synd,
                     uint8_t *logical, uint8_t *valid, int32_t *weight)
{
    for (long i = 0; i < B; ++i) {
        const uint8_t *e = err + i * n, *c = corr + i * n;
        int wt = 0, lg = 0;
        for (int j = 0; j < n; ++j) wt += (e[j] ^ c[j]) & 1;
        for (int r = 0; r < k; ++r) {
            int par = 0;
            for (int j = 0; j < n; ++j) par ^= L[(size_t)r * n + j] & (e[j] ^ c[j]);
            lg |= par;
        }
        logical[i] = (uint8_t)lg;
        valid[i] = (uint8_t)(synd_mismatch(m, n, Hdense, c, synd + i * m) == 0);
        weight[i] = wt;
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
