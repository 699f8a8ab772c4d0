/*
 * oracle/bluest_oracle.c -- TEST INFRASTRUCTURE ONLY (not the product).
 *
 * Plain-C, single-threaded CPU restatement of the five scalar loop nests of the
 * BLUEST native extension (reference: bluest/cmisc.cpp:10-97).  It exists so the
 * CUDA path can be checked against an independent statement of the algorithm
 * on the GPU box, where /root/reference is not available.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (bluest_b200/) never links or calls it.
 *
 * Conventions follow the reference: a size class holds Lk groups of exactly k
 * models; `groups` is the (Lk,k) int64 table of sorted model ids, `invcovs` the
 * (Lk,k,k) row-major table of per-group inverse covariances; every output is
 * caller-allocated, pre-zeroed and accumulated in place ("+=") except the
 * cleanup matrix which is *assigned* (see orc_cleanup_class).
 *
 * Pinned against: the reference's own compiled cmisc.cpp (oracle/_ref) and the
 * golden vectors under tests/golden/ produced by running the reference Python in
 * the build container (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -o liboracle.so bluest_oracle.c   (no -ffast-math:
 * summation order below is the program order, so results are reproducible).
 */
#include <stdint.h>
#include <stddef.h>

/* psi[(N*a + b), i] += Cinv_i[j,l] with a=g_i[j], b=g_i[l]; psi is (N*N, Lk) row-major.
 * Restates assemble_psi_c, bluest/cmisc.cpp:10-23. */
void orc_psi_class(double *psi, int64_t N, int64_t k, int64_t Lk,
                   const int64_t *groups, const double *invcovs)
{
    for (int64_t i = 0; i < Lk; ++i) {
        const int64_t *g = groups + i * k;
        const double *ci = invcovs + i * k * k;
        for (int64_t j = 0; j < k; ++j) {
            double *row = psi + (N * g[j]) * Lk + i;
            for (int64_t l = 0; l < k; ++l)
                row[g[l] * Lk] += ci[j * k + l];
        }
    }
}

/* PHI[N*a + b] += m_i * Cinv_i[j,l]; restates objectiveK_c<double>, bluest/cmisc.cpp:25-40. */
void orc_phi_class(double *phi, int64_t N, int64_t k, int64_t Lk, const double *m,
                   const int64_t *groups, const double *invcovs)
{
    for (int64_t i = 0; i < Lk; ++i) {
        const int64_t *g = groups + i * k;
        const double *ci = invcovs + i * k * k;
        const double w = m[i];
        for (int64_t j = 0; j < k; ++j) {
            double *row = phi + N * g[j];
            for (int64_t l = 0; l < k; ++l)
                row[g[l]] += w * ci[j * k + l];
        }
    }
}

/* Integer-weight overload (objectiveK_c<long int>, bluest/cmisc.cpp:105). */
void orc_phi_class_i64(double *phi, int64_t N, int64_t k, int64_t Lk, const int64_t *m,
                       const int64_t *groups, const double *invcovs)
{
    for (int64_t i = 0; i < Lk; ++i) {
        const int64_t *g = groups + i * k;
        const double *ci = invcovs + i * k * k;
        for (int64_t j = 0; j < k; ++j)
            for (int64_t l = 0; l < k; ++l)
                phi[N * g[j] + g[l]] += (double)m[i] * ci[j * k + l];
    }
}

/* grad[i] += sum_{j,l} x[g[j]] Cinv_i[j,l] x[g[l]]   (sign applied by the caller,
 * misc.py:493); restates gradK_c, bluest/cmisc.cpp:58-72.  The summation order is the
 * reference's: j outer, l inner, one running sum per group. */
void orc_grad_class(double *grad, int64_t k, int64_t Lk, const int64_t *groups,
                    const double *invcovs, const double *x)
{
    for (int64_t i = 0; i < Lk; ++i) {
        const int64_t *g = groups + i * k;
        const double *ci = invcovs + i * k * k;
        double acc = grad[i];
        for (int64_t j = 0; j < k; ++j)
            for (int64_t l = 0; l < k; ++l)
                acc += x[g[j]] * ci[j * k + l] * x[g[l]];
        grad[i] = acc;
    }
}

/* Reference cleanup matrix, bug-compatible: X[g[j], i] = Cinv_i[j,l] * x[g[l]] is an
 * ASSIGNMENT inside the l loop, so only l = k-1 survives (bluest/cmisc.cpp:42-56,
 * the "=" at line 51).  X is (N, Lk) row-major. */
void orc_cleanup_class(double *X, int64_t k, int64_t Lk, const int64_t *groups,
                       const double *invcovs, const double *x)
{
    for (int64_t i = 0; i < Lk; ++i) {
        const int64_t *g = groups + i * k;
        const double *ci = invcovs + i * k * k;
        for (int64_t j = 0; j < k; ++j)
            for (int64_t l = 0; l < k; ++l)
                X[Lk * g[j] + i] = ci[j * k + l] * x[g[l]];
    }
}

/* What the cleanup matrix was meant to be: column i is u_i = R_i^T Cinv_i R_i x
 * (SURVEY.md section 0.3).  Not in the reference; used to check the U factor. */
void orc_ufactor_class(double *U, int64_t k, int64_t Lk, const int64_t *groups,
                       const double *invcovs, const double *x)
{
    for (int64_t i = 0; i < Lk; ++i) {
        const int64_t *g = groups + i * k;
        const double *ci = invcovs + i * k * k;
        for (int64_t j = 0; j < k; ++j) {
            double acc = 0.0;
            for (int64_t l = 0; l < k; ++l)
                acc += ci[j * k + l] * x[g[l]];
            U[Lk * g[j] + i] += acc;
        }
    }
}

/* hess[ik,iq] += sum x[gk[lk]] Ck[lk,jk] P[gk[jk],gq[jq]] Cq[jq,lq] x[gq[lq]];
 * restates hessKQ_c, bluest/cmisc.cpp:74-97, same 6-deep order so that the
 * rounding is the reference's (without -ffast-math re-association). hess is (Lk,Lq). */
void orc_hess_block(double *hess, int64_t N, int64_t k, int64_t q, int64_t Lk, int64_t Lq,
                    const int64_t *groupsk, const int64_t *groupsq,
                    const double *invcovsk, const double *invcovsq, const double *P)
{
    const double *x = P; /* first row of Phi^+ */
    for (int64_t ik = 0; ik < Lk; ++ik) {
        const int64_t *gk = groupsk + ik * k;
        const double *ck = invcovsk + ik * k * k;
        for (int64_t iq = 0; iq < Lq; ++iq) {
            const int64_t *gq = groupsq + iq * q;
            const double *cq = invcovsq + iq * q * q;
            double acc = hess[ik * Lq + iq];
            for (int64_t lk = 0; lk < k; ++lk)
                for (int64_t jk = 0; jk < k; ++jk)
                    for (int64_t jq = 0; jq < q; ++jq)
                        for (int64_t lq = 0; lq < q; ++lq)
                            acc += x[gk[lk]] * ck[lk * k + jk] * P[N * gk[jk] + gq[jq]]
                                   * cq[jq * q + lq] * x[gq[lq]];
            hess[ik * Lq + iq] = acc;
        }
    }
}

/* One-pass biased covariance of pilot samples: C_hat = S2/n - s1 s1^T / n^2, with
 * s1[i] = sum_s Y[s,i], S2[i,j] = sum_s Y[s,i] Y[s,j]  (blue_fn.py:159-167 default inner
 * product a*b; blue_models.py:333).  Y is (n, N) row-major.  Plain sequential sums. */
void orc_pilot_cov(const double *Y, int64_t n, int64_t N, double *s1, double *S2, double *C)
{
    for (int64_t i = 0; i < N; ++i) s1[i] = 0.0;
    for (int64_t i = 0; i < N * N; ++i) S2[i] = 0.0;
    for (int64_t s = 0; s < n; ++s) {
        const double *y = Y + s * N;
        for (int64_t i = 0; i < N; ++i) {
            s1[i] += y[i];
            for (int64_t j = 0; j < N; ++j) S2[i * N + j] += y[i] * y[j];
        }
    }
    const double dn = (double)n;
    for (int64_t i = 0; i < N; ++i)
        for (int64_t j = 0; j < N; ++j)
            C[i * N + j] = S2[i * N + j] / dn - (s1[i] * s1[j]) / (dn * dn);
}
