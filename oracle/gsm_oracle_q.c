/*
 * gsm_oracle_q.c -- QUAD-PRECISION (IEEE binary128, ~34 digits) twin of the local-Kriging oracle.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, tools/ reports): never linked or called by the product library.
 *
 * Purpose (VERDICT r01 "next round" item 1): a float64 solver on an ill-conditioned
 * system (UniversalKriging with raw monomials of alpha-scaled coordinates, Gaussian variograms without nugget)
 * carries a forward error far above 1e-10, so "GPU vs float64 oracle" cannot tell which side is wrong.  This twin
 * evaluates the SAME reference formulation
 *     [C F; F' 0] [lambda; nu] = [c0; f0]        (krig.jl:111-138, 165-179; ordinary.jl:33-77; universal.jl:81-137)
 *     mean = lambda' z  (+ SK mean, simple.jl:55-69)      var = C0 - c0' lambda - f0' nu + 1e-10  (krig.jl:264-293)
 * from the SAME float64 inputs (alpha-scaled coordinates, values, neighbour rows), but carries every operation --
 * distances, variogram, monomials, Gaussian elimination with partial pivoting, inner products -- in __float128.
 * With condition numbers up to ~1e12 the result is correct to > 20 digits, i.e. it is the exact answer of the
 * reference's equations as far as any float64 comparison can see.  Both the GPU result and the float64 oracle are
 * then measured against it.
 *
 * Variogram formulas: the ones of gsm_oracle.py (restated GeoStatsFunctions 0.15; Gaussian uses n' = nugget + 1e-6
 * in both terms).
 */
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __float128 q_t;

#define QMAXK 64
#define QMAXCON 10
#define QMAXM (QMAXK + QMAXCON)

typedef struct { int kind; double sill, nugget, range; } ovario;                       /* as in gsm_oracle.c */
typedef struct { int kind; double mean; int ndrift; int exps[QMAXCON][3]; } omodel;   /* 0 SK 1 OK 2 UK */

static q_t gamma_q(const ovario* v, q_t h) {
  q_t s = v->sill, n = v->nugget, r = v->range, t = h / r;
  if (v->kind == 0) {
    n = n + (q_t)1e-6;   /* the double constant 1e-6, like the reference's typeof(s)(1e-6) */
    return (s - n) * (1 - expq(-3 * (t * t))) + (h > 0 ? n : (q_t)0);
  }
  if (v->kind == 1) {
    q_t g = (h < r) ? (s - n) * ((q_t)1.5 * t - (q_t)0.5 * (t * t * t)) : (s - n);
    return g + (h > 0 ? n : (q_t)0);
  }
  return (s - n) * (1 - expq(-3 * t)) + (h > 0 ? n : (q_t)0);
}
static q_t cov_q(const ovario* v, q_t h) { return (q_t)v->sill - gamma_q(v, h); }

static q_t dist_q(const double* a, const double* b, int dim) {
  q_t s = 0;
  for (int d = 0; d < dim; ++d) { q_t df = (q_t)a[d] - (q_t)b[d]; s += df * df; }
  return sqrtq(s);
}
static q_t ipow_q(q_t x, int p) { q_t r = 1; for (int i = 0; i < p; ++i) r *= x; return r; }

/* Gaussian elimination with partial pivoting, nrhs = 1; returns 0 when a pivot is exactly zero */
static int solve_q(q_t* A, q_t* b, int m) {
  for (int c = 0; c < m; ++c) {
    int p = c; q_t best = fabsq(A[c * m + c]);
    for (int r = c + 1; r < m; ++r) { q_t a = fabsq(A[r * m + c]); if (a > best) { best = a; p = r; } }
    if (best == 0) return 0;
    if (p != c) {
      for (int j = 0; j < m; ++j) { q_t t = A[c * m + j]; A[c * m + j] = A[p * m + j]; A[p * m + j] = t; }
      q_t t = b[c]; b[c] = b[p]; b[p] = t;
    }
    for (int r = c + 1; r < m; ++r) {
      q_t f = A[r * m + c] / A[c * m + c];
      if (f == 0) continue;
      for (int j = c; j < m; ++j) A[r * m + j] -= f * A[c * m + j];
      b[r] -= f * b[c];
    }
  }
  for (int r = m - 1; r >= 0; --r) {
    q_t s = b[r];
    for (int j = r + 1; j < m; ++j) s -= A[r * m + j] * b[j];
    b[r] = s / A[r * m + r];
  }
  return 1;
}

/*
 * sX: n x dim scaled sample coordinates, z: n values, sT: M x dim scaled targets,
 * nbr: M x k sample rows (0-based; only the first cnt[t] are read), cnt: M.
 * Outputs (rounded once to double): mean[M], var[M] (with the +1e-10), ok[M] (0: singular / no neighbours).
 * cond_est (optional, M): a cheap conditioning witness = max|pivot| / min|pivot| of the elimination.
 */
int gsmq_local(const double* sX, int dim, int n, const double* z, const ovario* v, const omodel* m,
               const double* sT, long M, int k, const int* nbr, const int* cnt,
               double* mean, double* var, unsigned char* ok, double* cond_est) {
  (void)n;
  int ncon = m->kind == 0 ? 0 : (m->kind == 1 ? 1 : m->ndrift);
  if (k > QMAXK || ncon > QMAXCON) return 1;
  q_t* A = (q_t*)malloc(sizeof(q_t) * QMAXM * QMAXM);
  q_t* A0 = (q_t*)malloc(sizeof(q_t) * QMAXM * QMAXM);
  if (!A || !A0) return 2;
  for (long t = 0; t < M; ++t) {
    const int nn = cnt[t];
    const int* id = nbr + t * (long)k;
    const double* x0 = sT + t * (long)dim;
    ok[t] = 0; mean[t] = 0; var[t] = 0;
    if (cond_est) cond_est[t] = 0;
    if (nn <= 0) continue;
    const int mm = nn + ncon;
    q_t b[QMAXM], b0[QMAXM];
    for (int i = 0; i < nn; ++i) {
      const double* xi = sX + (long)id[i] * dim;
      for (int j = 0; j <= i; ++j) {
        const double* xj = sX + (long)id[j] * dim;
        q_t c = (i == j) ? (q_t)v->sill - gamma_q(v, 0) : cov_q(v, dist_q(xi, xj, dim));
        A[i * mm + j] = c; A[j * mm + i] = c;
      }
      b[i] = cov_q(v, dist_q(xi, x0, dim));
      for (int c = 0; c < ncon; ++c) {
        q_t f = 1;
        if (m->kind == 2) for (int d = 0; d < dim; ++d) f *= ipow_q((q_t)xi[d], m->exps[c][d]);
        A[i * mm + nn + c] = f; A[(nn + c) * mm + i] = f;
      }
    }
    for (int c = 0; c < ncon; ++c) {
      for (int c2 = 0; c2 < ncon; ++c2) A[(nn + c) * mm + nn + c2] = 0;
      q_t f = 1;
      if (m->kind == 2) for (int d = 0; d < dim; ++d) f *= ipow_q((q_t)x0[d], m->exps[c][d]);
      b[nn + c] = f;
    }
    memcpy(A0, A, sizeof(q_t) * mm * mm);
    memcpy(b0, b, sizeof(q_t) * mm);
    if (!solve_q(A, b, mm)) continue;
    if (cond_est) {
      q_t lo = fabsq(A[0]), hi = lo;
      for (int i = 1; i < mm; ++i) { q_t p = fabsq(A[i * mm + i]); if (p < lo) lo = p; if (p > hi) hi = p; }
      cond_est[t] = (double)(hi / lo);
    }
    q_t mu = (m->kind == 0) ? (q_t)m->mean : (q_t)0;
    for (int i = 0; i < nn; ++i) mu += b[i] * ((q_t)z[id[i]] - ((m->kind == 0) ? (q_t)m->mean : (q_t)0));
    q_t s2 = (q_t)v->sill - gamma_q(v, 0);
    for (int i = 0; i < mm; ++i) s2 -= b0[i] * b[i];
    s2 += (q_t)1e-10;
    mean[t] = (double)mu;
    var[t] = (double)s2;
    ok[t] = 1;
  }
  free(A); free(A0);
  return 0;
}
