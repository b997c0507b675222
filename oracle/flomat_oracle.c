/* flomat_oracle.c -- TEST INFRASTRUCTURE ONLY (parity checker, never the product path).
 *
 * A plain-C, CPU restatement of the arithmetic that soegaard/sci's flomat reaches through
 * its FFI boundary on the dense FP64 hot path.  The reference has no native code of its own:
 * the arithmetic lives in whatever system CBLAS/LAPACK `ffi-lib` finds (flomat/flomat.rkt:92-135,
 * un-vendored and un-pinned; flomat/info.rkt:14-16 names no version).  This file therefore
 * restates the *published* reference-BLAS / reference-LAPACK (netlib) algorithms for exactly
 * the symbols flomat binds on this path, with the argument conventions of the Racket `_fun`
 * types (column-major, 32-bit ints, 1-based int32 ipiv, LAPACK scalars by reference):
 *
 *   ora_dgemm   <- cblas_dgemm   flomat.rkt:855-875 (call site :888)
 *   ora_daxpy   <- cblas_daxpy   flomat.rkt:751-757 (call site :810)
 *   ora_dcopy   <- cblas_dcopy   flomat.rkt:483-488 (call sites :500, :544 incX=0 broadcast)
 *   ora_dscal   <- cblas_dscal   flomat.rkt:1014-1018 (call site :1026)
 *   ora_dlange  <- dlange_       flomat.rkt:1317-1324 (call site :1340)
 *   ora_dgetrf  <- dgetrf_       flomat.rkt:1505-1523 (call site :1527)
 *   ora_dgesv   <- dgesv_        flomat.rkt:2112-2124 (call sites :2131, :2140)
 *   ora_dgetri  <- dgetri_       flomat.rkt:1754-1768 (call site :1778)
 *
 * plus the Racket-side per-element loops of the pointwise macros (flomat.rkt:3021-3146):
 *   ora_ewise   <- .+ .- .* ./   (one IEEE operation per element => bit-exact parity expected)
 *
 * Parity pinning: checked against every golden vector the reference's own rackunit suite holds
 * for this path (flomat.rkt:3335-3353, 3405-3430, 3547-3569, 3609-3617, 3655-3660); see
 * tests/golden/ and tests/test_oracle_golden.py.  expm has NO reference test (expm.rkt has no test
 * submodule) -> for expm parity is "unpinned" beyond the manual's two doc examples.
 *
 * Build: make -C oracle   (gcc -O3 -shared -fPIC -pthread; -ffp-contract=off so that
 * no FMA contraction changes the single-rounding-per-operation semantics of the restatement).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <stddef.h>
#include <pthread.h>

#define A_(i, j) a[(size_t)(i) + (size_t)(j) * lda]
#define B_(i, j) b[(size_t)(i) + (size_t)(j) * ldb]
#define C_(i, j) c[(size_t)(i) + (size_t)(j) * ldc]

enum { ORA_NOTRANS = 111, ORA_TRANS = 112 };

/* Reference BLAS DGEMM, column major (order=102 is the only order flomat.rkt:888 passes). */
static void dgemm_cols(int j0, int j1, int ta, int tb, int M, int K, double alpha, const double *a, int lda,
                       const double *b, int ldb, double beta, double *c, int ldc) {
    for (int j = j0; j < j1; ++j) {
        if (beta == 0.0) {
            for (int i = 0; i < M; ++i) C_(i, j) = 0.0;
        } else if (beta != 1.0) {
            for (int i = 0; i < M; ++i) C_(i, j) = beta * C_(i, j);
        }
        if (alpha == 0.0) continue;
        if (!ta) {
            for (int l = 0; l < K; ++l) {
                const double temp = alpha * (tb ? B_(j, l) : B_(l, j));
                const double *acol = &A_(0, l);
                double *ccol = &C_(0, j);
                for (int i = 0; i < M; ++i) ccol[i] += temp * acol[i];
            }
        } else {
            for (int i = 0; i < M; ++i) {
                double temp = 0.0;
                for (int l = 0; l < K; ++l) temp += A_(l, i) * (tb ? B_(j, l) : B_(l, j));
                C_(i, j) = alpha * temp + (beta == 0.0 ? 0.0 : C_(i, j));
            }
        }
    }
}

/* Columns of C are independent, so splitting them over threads does not change any rounding. */
typedef struct { int j0, j1, ta, tb, M, K, lda, ldb, ldc; double alpha, beta; const double *a, *b; double *c; } gemm_job;
static void *gemm_worker(void *p) {
    gemm_job *g = (gemm_job *)p;
    dgemm_cols(g->j0, g->j1, g->ta, g->tb, g->M, g->K, g->alpha, g->a, g->lda, g->b, g->ldb, g->beta, g->c, g->ldc);
    return NULL;
}
static int ora_threads = 1;
void ora_set_threads(int t) { ora_threads = t < 1 ? 1 : (t > 64 ? 64 : t); }

void ora_dgemm(int transA, int transB, int M, int N, int K, double alpha, const double *a, int lda,
               const double *b, int ldb, double beta, double *c, int ldc) {
    const int ta = (transA == ORA_TRANS), tb = (transB == ORA_TRANS);
    int nt = ora_threads;
    if ((double)M * N * K < 1e6 || N < nt) nt = 1;
    if (nt == 1) { dgemm_cols(0, N, ta, tb, M, K, alpha, a, lda, b, ldb, beta, c, ldc); return; }
    pthread_t th[64]; gemm_job jobs[64];
    for (int t = 0; t < nt; ++t) {
        gemm_job g = { (int)((long)N * t / nt), (int)((long)N * (t + 1) / nt), ta, tb, M, K, lda, ldb, ldc, alpha, beta, a, b, c };
        jobs[t] = g;
        pthread_create(&th[t], NULL, gemm_worker, &jobs[t]);
    }
    for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

void ora_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy) {
    if (n <= 0 || alpha == 0.0) return;
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0, iy = incy < 0 ? (ptrdiff_t)(1 - n) * incy : 0;
    for (int i = 0; i < n; ++i, ix += incx, iy += incy) y[iy] += alpha * x[ix];
}

void ora_dcopy(int n, const double *x, int incx, double *y, int incy) {
    ptrdiff_t ix = incx < 0 ? (ptrdiff_t)(1 - n) * incx : 0, iy = incy < 0 ? (ptrdiff_t)(1 - n) * incy : 0;
    for (int i = 0; i < n; ++i, ix += incx, iy += incy) y[iy] = x[ix];
}

void ora_dscal(int n, double alpha, double *x, int incx) {
    if (n <= 0 || incx <= 0) return;
    for (int i = 0; i < n; ++i) x[(size_t)i * incx] = alpha * x[(size_t)i * incx];
}

/* Reference LAPACK DLANGE: 'M' max|a|, '1'/'O' max column abs-sum, 'I' max row abs-sum,
 * 'F'/'E' Frobenius via the scaled sum of squares (DLASSQ). NaN propagates as in LAPACK >= 3.x. */
double ora_dlange(const char *norm, const int *pm, const int *pn, const double *a, const int *plda, double *work) {
    const int m = *pm, n = *pn, lda = *plda;
    if (m == 0 || n == 0) return 0.0;
    double value = 0.0;
    const char t = norm[0];
    if (t == 'M' || t == 'm') {
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) {
                double v = fabs(A_(i, j));
                if (value < v || isnan(v)) value = v;
            }
    } else if (t == '1' || t == 'O' || t == 'o') {
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            for (int i = 0; i < m; ++i) sum += fabs(A_(i, j));
            if (value < sum || isnan(sum)) value = sum;
        }
    } else if (t == 'I' || t == 'i') {
        double *w = work ? work : (double *)calloc((size_t)m, sizeof(double));
        for (int i = 0; i < m; ++i) w[i] = 0.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) w[i] += fabs(A_(i, j));
        for (int i = 0; i < m; ++i)
            if (value < w[i] || isnan(w[i])) value = w[i];
        if (!work) free(w);
    } else { /* 'F' / 'E' */
        double scale = 0.0, sumsq = 1.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) {
                double v = fabs(A_(i, j));
                if (v > 0.0 || isnan(v)) {
                    if (scale < v) { sumsq = 1.0 + sumsq * (scale / v) * (scale / v); scale = v; }
                    else sumsq += (v / scale) * (v / scale);
                }
            }
        value = scale * sqrt(sumsq);
    }
    return value;
}

/* DGETF2-style unblocked right-looking LU with partial pivoting (first max = IDAMAX tie-break).
 * ipiv is 1-based (flomat.rkt:1457-1463). info>0: U(info,info) exactly zero (factorisation completes). */
void ora_dgetrf(const int *pm, const int *pn, double *a, const int *plda, int *ipiv, int *info) {
    const int m = *pm, n = *pn, lda = *plda;
    *info = 0;
    if (m < 0) { *info = -1; return; }
    if (n < 0) { *info = -2; return; }
    if (lda < (m > 1 ? m : 1)) { *info = -4; return; }
    const int mn = m < n ? m : n;
    const double sfmin = DBL_MIN;
    for (int j = 0; j < mn; ++j) {
        int jp = j;
        double vmax = fabs(A_(j, j));
        for (int i = j + 1; i < m; ++i) {
            double v = fabs(A_(i, j));
            if (v > vmax) { vmax = v; jp = i; }
        }
        ipiv[j] = jp + 1;
        if (A_(jp, j) != 0.0) {
            if (jp != j)
                for (int k = 0; k < n; ++k) { double t = A_(j, k); A_(j, k) = A_(jp, k); A_(jp, k) = t; }
            if (j < m - 1) {
                if (fabs(A_(j, j)) >= sfmin) {
                    const double r = 1.0 / A_(j, j);
                    for (int i = j + 1; i < m; ++i) A_(i, j) *= r;
                } else {
                    for (int i = j + 1; i < m; ++i) A_(i, j) /= A_(j, j);
                }
            }
        } else if (*info == 0) {
            *info = j + 1;
        }
        if (j < mn - 1 || n > mn) {
            for (int k = j + 1; k < n; ++k) {
                const double t = A_(j, k);
                if (t != 0.0)
                    for (int i = j + 1; i < m; ++i) A_(i, k) -= A_(i, j) * t;
            }
        }
    }
}

/* DGETRS 'N': row interchanges (DLASWP), unit-lower forward solve, upper back solve (reference DTRSM order). */
static void ora_getrs_n(int n, int nrhs, const double *a, int lda, const int *ipiv, double *b, int ldb) {
    for (int i = 0; i < n; ++i) {
        const int ip = ipiv[i] - 1;
        if (ip != i)
            for (int k = 0; k < nrhs; ++k) { double t = B_(i, k); B_(i, k) = B_(ip, k); B_(ip, k) = t; }
    }
    for (int j = 0; j < nrhs; ++j) {
        for (int k = 0; k < n; ++k) {
            const double t = B_(k, j);
            if (t != 0.0)
                for (int i = k + 1; i < n; ++i) B_(i, j) -= t * A_(i, k);
        }
        for (int k = n - 1; k >= 0; --k) {
            if (B_(k, j) != 0.0) {
                B_(k, j) /= A_(k, k);
                const double t = B_(k, j);
                for (int i = 0; i < k; ++i) B_(i, j) -= t * A_(i, k);
            }
        }
    }
}

void ora_dgesv(const int *pn, const int *pnrhs, double *a, const int *plda, int *ipiv, double *b, const int *pldb, int *info) {
    const int n = *pn, nrhs = *pnrhs, lda = *plda, ldb = *pldb;
    *info = 0;
    if (n < 0) { *info = -1; return; }
    if (nrhs < 0) { *info = -2; return; }
    if (lda < (n > 1 ? n : 1)) { *info = -4; return; }
    if (ldb < (n > 1 ? n : 1)) { *info = -7; return; }
    ora_dgetrf(pn, pn, a, plda, ipiv, info);
    if (*info == 0) ora_getrs_n(n, nrhs, a, lda, ipiv, b, ldb);
}

/* DGETRI (unblocked path): inv(U) by DTRTI2, then solve inv(A)*L = inv(U) column by column from the
 * right using `work` (n doubles used), then undo the row interchanges as column swaps in reverse. */
void ora_dgetri(const int *pn, double *a, const int *plda, const int *ipiv, double *work, const int *plwork, int *info) {
    const int n = *pn, lda = *plda;
    (void)plwork;
    *info = 0;
    if (n < 0) { *info = -1; return; }
    if (lda < (n > 1 ? n : 1)) { *info = -3; return; }
    if (n == 0) return;
    for (int j = 0; j < n; ++j)
        if (A_(j, j) == 0.0) { *info = j + 1; return; }
    /* DTRTI2, upper, non-unit */
    for (int j = 0; j < n; ++j) {
        A_(j, j) = 1.0 / A_(j, j);
        const double ajj = -A_(j, j);
        /* x := U(0:j,0:j) * x  (DTRMV upper notrans non-unit) on column j */
        for (int k = 0; k < j; ++k) {
            const double t = A_(k, j);
            if (t != 0.0) {
                for (int i = 0; i < k; ++i) A_(i, j) += t * A_(i, k);
                A_(k, j) = t * A_(k, k);
            }
        }
        for (int i = 0; i < j; ++i) A_(i, j) *= ajj;
    }
    double *w = work ? work : (double *)malloc((size_t)n * sizeof(double));
    for (int j = n - 1; j >= 0; --j) {
        for (int i = j + 1; i < n; ++i) { w[i] = A_(i, j); A_(i, j) = 0.0; }
        if (j < n - 1) {
            /* A(:,j) -= A(:,j+1:n) * w(j+1:n)   (DGEMV notrans, alpha=-1, beta=1) */
            for (int k = j + 1; k < n; ++k) {
                const double t = -w[k];
                if (t != 0.0)
                    for (int i = 0; i < n; ++i) A_(i, j) += t * A_(i, k);
            }
        }
    }
    for (int j = n - 2; j >= 0; --j) {
        const int jp = ipiv[j] - 1;
        if (jp != j)
            for (int i = 0; i < n; ++i) { double t = A_(i, j); A_(i, j) = A_(i, jp); A_(i, jp) = t; }
    }
    if (!work) free(w);
}

/* Pointwise macros flomat.rkt:3021-3146: op 0:+ 1:- 2:* 3:/ ; a or b may be a scalar (sa/sb used when the
 * pointer is NULL).  One IEEE-754 operation per element, so any correct implementation is bit-identical. */
void ora_ewise(int op, int m, int n, const double *a, int lda, double sa, const double *b, int ldb, double sb,
               double *c, int ldc) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            const double x = a ? A_(i, j) : sa, y = b ? B_(i, j) : sb;
            double r;
            switch (op) {
                case 0: r = x + y; break;
                case 1: r = x - y; break;
                case 2: r = x * y; break;
                default: r = x / y; break;
            }
            C_(i, j) = r;
        }
}
