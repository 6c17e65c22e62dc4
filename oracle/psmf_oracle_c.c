/* C / OpenMP restatement of the masked PSMF / rPSMF filter step  --  TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithm as oracle/psmf_oracle.py (which is pinned against the reference's golden vectors and
 * cross-checks this file in tests/test_oracle_golden.py); used as the multi-threaded CPU baseline of
 * bench.py where numpy temporaries would dominate.  Random-walk dynamics (f = identity), R = rho I.
 *
 * Reference lines restated:  ExperimentImpute/rPSMF.py:81-135 (step), :30-36 (compute_Sinv, via the
 * O(d r^2) statistics), PSMF.py:60-84 (robust = 0).
 *
 * Build: make -C oracle   ->  oracle/libpsmf_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXR 16

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline sets its thread count explicitly */
void psmf_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int psmf_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* solve A X = B (r x r, r x nb) with partial pivoting; A and B are overwritten, X returned in B */
static void solve(int r, int nb, double* A, double* B) {
    for (int k = 0; k < r; ++k) {
        int p = k;
        for (int i = k + 1; i < r; ++i)
            if (fabs(A[i * r + k]) > fabs(A[p * r + k])) p = i;
        if (p != k) {
            for (int j = 0; j < r; ++j) { double t = A[k * r + j]; A[k * r + j] = A[p * r + j]; A[p * r + j] = t; }
            for (int j = 0; j < nb; ++j) { double t = B[k * nb + j]; B[k * nb + j] = B[p * nb + j]; B[p * nb + j] = t; }
        }
        const double inv = 1.0 / A[k * r + k];
        for (int i = k + 1; i < r; ++i) {
            const double f = A[i * r + k] * inv;
            if (f == 0.0) continue;
            for (int j = k; j < r; ++j) A[i * r + j] -= f * A[k * r + j];
            for (int j = 0; j < nb; ++j) B[i * nb + j] -= f * B[k * nb + j];
        }
    }
    for (int k = r - 1; k >= 0; --k) {
        const double inv = 1.0 / A[k * r + k];
        for (int j = 0; j < nb; ++j) {
            double s = B[k * nb + j];
            for (int i = k + 1; i < r; ++i) s -= A[k * r + i] * B[i * nb + j];
            B[k * nb + j] = s * inv;
        }
    }
}

/* One launch = n_steps filter steps.  C (d, r) row-major is updated in place; state arrays are in/out.
 * Y, M time-major (n_steps, d).  X_out (n_steps, r) optional.  scal = {rho, lambda} in/out.
 * Returns the first step with a non-finite N, or -1. */
int64_t psmf_oracle_run(int64_t d, int r, int robust, int cupdate_vt, double* C, double* x, double* P, double* V, double* Q,
                        double* scal, const double* Y, const uint8_t* M, int64_t n_steps, double* X_out) {
    if (r > MAXR) return -2;
    double* e = (double*)malloc(sizeof(double) * (size_t)d);
    double rho = scal[0], lam = scal[1];
    int64_t bad = -1;
    const int nst = r * r + r + 4;
    for (int64_t t = 0; t < n_steps; ++t) {
        const double* y = Y + t * d;
        const uint8_t* m = M ? M + t * d : NULL;
        double xb[MAXR], vx[MAXR], vxt[MAXR], Pb[MAXR * MAXR], a = 0.0;
        for (int j = 0; j < r; ++j) xb[j] = x[j];                                  /* rPSMF.py:86 */
        for (int i = 0; i < r * r; ++i) Pb[i] = P[i] + Q[i];                       /* rPSMF.py:87 */
        for (int j = 0; j < r; ++j) {
            double s1 = 0, s2 = 0;
            for (int k = 0; k < r; ++k) { s1 += V[j * r + k] * xb[k]; s2 += V[k * r + j] * xb[k]; }
            vx[j] = s1; vxt[j] = s2;
        }
        for (int j = 0; j < r; ++j) a += xb[j] * vx[j];                            /* rPSMF.py:93 */
        const double w1 = 1.0 / (rho + a), w0 = 1.0 / a;                           /* rPSMF.py:92,98,32 */
        double tot[MAXR * MAXR + MAXR + 4];
        memset(tot, 0, sizeof(tot));
#pragma omp parallel
        {
            double loc[MAXR * MAXR + MAXR + 4];
            memset(loc, 0, sizeof(loc));
#pragma omp for schedule(static) nowait
            for (int64_t i = 0; i < d; ++i) {
                const double* c = C + i * r;
                const int mi = m ? (m[i] != 0) : 1;
                double yh = 0.0;
                for (int j = 0; j < r; ++j) yh += c[j] * xb[j];                    /* rPSMF.py:89 */
                const double ei = y[i] - (mi ? yh : 0.0);                          /* rPSMF.py:101 */
                e[i] = ei;
                if (mi) {
                    for (int j = 0; j < r; ++j) {
                        const double cw = c[j] * w1;
                        for (int k = j; k < r; ++k) loc[j * r + k] += cw * c[k];   /* CM' Ri CM */
                        loc[r * r + j] += ei * w1 * c[j];                          /* CM' Ri diff */
                    }
                    loc[r * r + r + 0] += w1 * ei * ei;
                    loc[r * r + r + 1] += ei * ei;
                    loc[r * r + r + 3] += 1.0;
                } else {
                    loc[r * r + r + 0] += w0 * ei * ei;
                    loc[r * r + r + 2] += ei * ei;
                }
            }
#pragma omp critical
            for (int i = 0; i < nst; ++i) tot[i] += loc[i];
        }
        double G[MAXR * MAXR];
        for (int j = 0; j < r; ++j)
            for (int k = j; k < r; ++k) G[j * r + k] = G[k * r + j] = tot[j * r + k];
        const double* b = tot + r * r;
        const double s = tot[r * r + r], q1 = tot[r * r + r + 1], q0 = tot[r * r + r + 2], nobs = tot[r * r + r + 3];
        /* K = (I + Pb G)^-1 Pb, Kb = K b */
        double A[MAXR * MAXR], B[MAXR * (MAXR + 1)];
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < r; ++j) {
                double acc = (i == j) ? 1.0 : 0.0;
                for (int k = 0; k < r; ++k) acc += Pb[i * r + k] * G[k * r + j];
                A[i * r + j] = acc;
                B[i * (r + 1) + j] = Pb[i * r + j];
            }
        for (int i = 0; i < r; ++i) {
            double acc = 0.0;
            for (int k = 0; k < r; ++k) acc += Pb[i * r + k] * b[k];
            B[i * (r + 1) + r] = acc;
        }
        solve(r, r + 1, A, B);
        double bkb = 0.0, trpg = 0.0;
        for (int j = 0; j < r; ++j) {
            const double kb = B[j * (r + 1) + r];
            x[j] = xb[j] + kb;                                                     /* rPSMF.py:104 */
            bkb += b[j] * kb;
            for (int i = 0; i < r; ++i) trpg += Pb[i * r + j] * G[i * r + j];
        }
        const double sSe = s - bkb;
        const double eta = (rho * nobs + (rho + a) * trpg) / (double)d;            /* rPSMF.py:108 */
        const double omega = robust ? (lam + sSe) / (lam + (double)d) : 1.0;       /* rPSMF.py:105 */
        const double N = a + eta;                                                  /* rPSMF.py:109 */
        const double phi = robust ? (lam + q1 / (a + eta) + (q0 != 0.0 ? q0 / eta : 0.0)) / (lam + (double)d) : 1.0;
        if (!isfinite(N) || N == 0.0 || !isfinite(omega) || !isfinite(phi)) { if (bad < 0) bad = t; }
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < r; ++j) {
                P[i * r + j] = omega * B[i * (r + 1) + j];                         /* rPSMF.py:106 */
                V[i * r + j] = phi * (V[i * r + j] - vx[i] * vxt[j] / N);          /* rPSMF.py:115 */
                Q[i * r + j] = omega * Q[i * r + j];                               /* rPSMF.py:133 */
            }
        double g[MAXR];
        for (int j = 0; j < r; ++j) g[j] = (cupdate_vt ? vx[j] : vxt[j]) / N;
        rho = omega * rho;                                                         /* rPSMF.py:134 */
        if (robust) lam += (double)d;                                              /* rPSMF.py:135 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < d; ++i) {                                          /* rPSMF.py:111 */
            double* c = C + i * r;
            const double ei = e[i];
            for (int j = 0; j < r; ++j) c[j] += ei * g[j];
        }
        if (X_out) for (int j = 0; j < r; ++j) X_out[t * r + j] = x[j];
    }
    scal[0] = rho; scal[1] = lam;
    free(e);
    return bad;
}
