/*
 * gsm_oracle.c -- C restatement of the reference CPU algorithm (TEST INFRASTRUCTURE ONLY).
 *
 * Twin of oracle/gsm_oracle.py for sizes where a Python loop is too slow, and the timed
 * "reference-algorithm CPU restatement" of bench.py (cpu_baseline.kind = "port").  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * (geostatsmodels.jl_b200/) never does.  PARITY PINNING: validated against gsm_oracle.py (which is
 * pinned to the reference's golden values) in tests/test_oracle_c.py; the GeoStatsFunctions / Meshes /
 * NearestNeighbors arithmetic is restated, not vendored (see gsm_oracle.py header: parity unpinned for
 * interior variogram values, neighbour tie-breaks and scaled-grid centroid arithmetic).
 *
 * Per target it follows /root/reference/src step for step:
 *   models.jl:162-184   search k nearest (KD-tree, like NearestNeighbors behind Meshes.KNearestSearch),
 *                       view the neighbours, fit, predict
 *   krig.jl:88-138      allocate-free assembly of the (k+ncon)^2 system in covariance form
 *   krig.jl:148-162     Bunch-Kaufman LDL' (LAPACK dsytf2 algorithm, upper triangle, partial pivoting)
 *   krig.jl:316-339     solve (dsytrs)          krig.jl:245-258 / simple.jl:55-69   mean
 *   krig.jl:264-293     variance + 1e-10        models.jl:236-247  contiguous target chunks per thread
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -pthread (oracle/Makefile).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXK 64
#define MAXCON 10
#define MAXM (MAXK + MAXCON)

typedef struct { int kind; double sill, nugget, range; } ovario;          /* 0 gaussian 1 spherical 2 exponential */
typedef struct { int kind; double mean; int ndrift; int exps[MAXCON][3]; } omodel; /* 0 SK 1 OK 2 UK */

/* ---------------------------------------------------------------- variogram (gsm_oracle.py twin) */
static double gamma_h(const ovario* v, double h) {
  double s = v->sill, n = v->nugget, r = v->range, t = h / r, g;
  if (v->kind == 0) { n = n + 1e-6; return (s - n) * (1.0 - exp(-3.0 * (t * t))) + (h > 0 ? n : 0.0); }  /* n' everywhere */
  if (v->kind == 1) {
    g = (h < r) ? (s - n) * (1.5 * t - 0.5 * (t * t * t)) : (s - n);
    return g + (h > 0 ? n : 0.0);
  }
  return (s - n) * (1.0 - exp(-3.0 * t)) + (h > 0 ? n : 0.0);
}
static double cov_h(const ovario* v, double h) { return v->sill - gamma_h(v, h); }

static double d2key(const double* a, const double* b, int dim) {
  double s = 0.0;
  for (int d = 0; d < dim; ++d) { double df = a[d] - b[d]; s = s + df * df; }
  return s;
}

/* ---------------------------------------------------------------- bounded max-heap on (d2, idx) */
typedef struct { double d[MAXK]; int i[MAXK]; int n, k; } heap_t;
static int kless(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }
static void heap_push(heap_t* h, double d, int idx) {
  if (h->n < h->k) {
    int c = h->n++;
    while (c > 0) { int p = (c - 1) / 2; if (!kless(h->d[p], h->i[p], d, idx)) break; h->d[c] = h->d[p]; h->i[c] = h->i[p]; c = p; }
    h->d[c] = d; h->i[c] = idx;
  } else if (kless(d, idx, h->d[0], h->i[0])) {
    int c = 0;
    for (;;) {
      int l = 2 * c + 1; if (l >= h->n) break;
      if (l + 1 < h->n && kless(h->d[l], h->i[l], h->d[l + 1], h->i[l + 1])) l++;
      if (!kless(d, idx, h->d[l], h->i[l])) break;
      h->d[c] = h->d[l]; h->i[c] = h->i[l]; c = l;
    }
    h->d[c] = d; h->i[c] = idx;
  }
}
static void heap_sorted(heap_t* h, int* out_idx, double* out_d2) { /* ascending */
  int n = h->n;
  for (int e = n - 1; e >= 0; --e) {
    out_idx[e] = h->i[0]; out_d2[e] = h->d[0];
    double d = h->d[e]; int idx = h->i[e]; int c = 0;
    for (;;) {
      int l = 2 * c + 1; if (l >= e) break;
      if (l + 1 < e && kless(h->d[l], h->i[l], h->d[l + 1], h->i[l + 1])) l++;
      if (!kless(d, idx, h->d[l], h->i[l])) break;
      h->d[c] = h->d[l]; h->i[c] = h->i[l]; c = l;
    }
    h->d[c] = d; h->i[c] = idx;
  }
}

/* ---------------------------------------------------------------- KD-tree */
typedef struct { int lo, hi, left, right, axis; double split; } kdnode;
typedef struct { const double* X; int dim, n; int* perm; kdnode* nodes; int nnodes, leaf; } kdtree;

static void nth_elem(const double* X, int dim, int axis, int* p, int lo, int hi, int nth) {
  while (hi - lo > 1) {
    double pv = X[(size_t)p[(lo + hi) / 2] * dim + axis];
    int i = lo, j = hi - 1;
    while (i <= j) {
      while (X[(size_t)p[i] * dim + axis] < pv) i++;
      while (X[(size_t)p[j] * dim + axis] > pv) j--;
      if (i <= j) { int t = p[i]; p[i] = p[j]; p[j] = t; i++; j--; }
    }
    if (nth <= j) hi = j + 1; else if (nth >= i) lo = i; else return;
  }
}
static int kd_build(kdtree* t, int lo, int hi) {
  int id = t->nnodes++;
  kdnode* nd = &t->nodes[id];
  nd->lo = lo; nd->hi = hi; nd->left = nd->right = -1; nd->axis = 0; nd->split = 0;
  if (hi - lo <= t->leaf) return id;
  double bestw = -1; int ax = 0;
  for (int d = 0; d < t->dim; ++d) {
    double mn = 1e300, mx = -1e300;
    for (int i = lo; i < hi; ++i) { double v = t->X[(size_t)t->perm[i] * t->dim + d]; if (v < mn) mn = v; if (v > mx) mx = v; }
    if (mx - mn > bestw) { bestw = mx - mn; ax = d; }
  }
  int mid = (lo + hi) / 2;
  nth_elem(t->X, t->dim, ax, t->perm, lo, hi, mid);
  double sp = t->X[(size_t)t->perm[mid] * t->dim + ax];
  int l = kd_build(t, lo, mid);
  int r = kd_build(t, mid, hi);
  nd = &t->nodes[id];
  nd->axis = ax; nd->split = sp; nd->left = l; nd->right = r;
  return id;
}
static void kd_query(const kdtree* t, int id, const double* q, heap_t* h) {
  const kdnode* nd = &t->nodes[id];
  if (nd->left < 0) {
    for (int i = nd->lo; i < nd->hi; ++i) { int idx = t->perm[i]; heap_push(h, d2key(t->X + (size_t)idx * t->dim, q, t->dim), idx); }
    return;
  }
  double diff = q[nd->axis] - nd->split;
  int near = diff < 0 ? nd->left : nd->right, far = diff < 0 ? nd->right : nd->left;
  kd_query(t, near, q, h);
  if (h->n < h->k || diff * diff <= h->d[0]) kd_query(t, far, q, h);
}

/* ---------------------------------------------------------------- Bunch-Kaufman (dsytf2/dsytrs, UPLO='U') */
#define A_(i, j) A[(size_t)(j) * lda + (i)]
static int bk_factor(double* A, int n, int lda, int* ipiv) {
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  int info = 0, k = n - 1;
  while (k >= 0) {
    int kstep = 1, kp, imax = 0;
    double absakk = fabs(A_(k, k)), colmax = 0.0;
    if (k > 0) { for (int i = 0; i < k; ++i) if (fabs(A_(i, k)) > colmax) { colmax = fabs(A_(i, k)); imax = i; } }
    if (fmax(absakk, colmax) == 0.0 || isnan(absakk)) { if (!info) info = k + 1; kp = k; }
    else {
      if (absakk >= alpha * colmax) kp = k;
      else {
        double rowmax = 0.0;
        for (int j = imax + 1; j <= k; ++j) if (fabs(A_(imax, j)) > rowmax) rowmax = fabs(A_(imax, j));
        for (int i = 0; i < imax; ++i) if (fabs(A_(i, imax)) > rowmax) rowmax = fabs(A_(i, imax));
        if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
        else if (fabs(A_(imax, imax)) >= alpha * rowmax) kp = imax;
        else { kp = imax; kstep = 2; }
      }
      int kk = k - kstep + 1;
      if (kp != kk) {
        for (int i = 0; i < kp; ++i) { double t = A_(i, kk); A_(i, kk) = A_(i, kp); A_(i, kp) = t; }
        for (int j = kp + 1; j < kk; ++j) { double t = A_(j, kk); A_(j, kk) = A_(kp, j); A_(kp, j) = t; }
        { double t = A_(kk, kk); A_(kk, kk) = A_(kp, kp); A_(kp, kp) = t; }
        if (kstep == 2) { double t = A_(k - 1, k); A_(k - 1, k) = A_(kp, k); A_(kp, k) = t; }
      }
      if (kstep == 1) {
        double r1 = 1.0 / A_(k, k);
        for (int j = 0; j < k; ++j) {
          double xj = A_(j, k);
          if (xj != 0.0) { double tmp = -r1 * xj; for (int i = 0; i <= j; ++i) A_(i, j) += A_(i, k) * tmp; }
        }
        for (int i = 0; i < k; ++i) A_(i, k) *= r1;
      } else if (k > 1) {
        double d12 = A_(k - 1, k), d22 = A_(k - 1, k - 1) / d12, d11 = A_(k, k) / d12, t = 1.0 / (d11 * d22 - 1.0);
        d12 = t / d12;
        for (int j = k - 2; j >= 0; --j) {
          double wkm1 = d12 * (d11 * A_(j, k - 1) - A_(j, k));
          double wk = d12 * (d22 * A_(j, k) - A_(j, k - 1));
          for (int i = j; i >= 0; --i) A_(i, j) = A_(i, j) - A_(i, k) * wk - A_(i, k - 1) * wkm1;
          A_(j, k) = wk; A_(j, k - 1) = wkm1;
        }
      }
    }
    if (kstep == 1) ipiv[k] = kp + 1; else { ipiv[k] = -(kp + 1); ipiv[k - 1] = -(kp + 1); }
    k -= kstep;
  }
  return info;
}
static void bk_solve(const double* A, int n, int lda, const int* ipiv, double* b) {
  int k = n - 1;
  while (k >= 0) { /* solve U*D*x = b */
    if (ipiv[k] > 0) {
      int kp = ipiv[k] - 1;
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      for (int i = 0; i < k; ++i) b[i] -= A_(i, k) * b[k];
      b[k] /= A_(k, k);
      k -= 1;
    } else {
      int kp = -ipiv[k] - 1;
      if (kp != k - 1) { double t = b[k - 1]; b[k - 1] = b[kp]; b[kp] = t; }
      for (int i = 0; i < k - 1; ++i) { b[i] -= A_(i, k) * b[k]; }
      for (int i = 0; i < k - 1; ++i) { b[i] -= A_(i, k - 1) * b[k - 1]; }
      double akm1k = A_(k - 1, k), akm1 = A_(k - 1, k - 1) / akm1k, ak = A_(k, k) / akm1k, denom = akm1 * ak - 1.0;
      double bkm1 = b[k - 1] / akm1k, bk = b[k] / akm1k;
      b[k - 1] = (ak * bkm1 - bk) / denom;
      b[k] = (akm1 * bk - bkm1) / denom;
      k -= 2;
    }
  }
  k = 0;
  while (k < n) { /* solve U'*x = b */
    if (ipiv[k] > 0) {
      double s = 0; for (int i = 0; i < k; ++i) s += A_(i, k) * b[i];
      b[k] -= s;
      int kp = ipiv[k] - 1;
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k += 1;
    } else {
      double s0 = 0, s1 = 0;
      for (int i = 0; i < k; ++i) { s0 += A_(i, k) * b[i]; s1 += A_(i, k + 1) * b[i]; }
      b[k] -= s0; b[k + 1] -= s1;
      int kp = -ipiv[k] - 1;
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k += 2;
    }
  }
}
#undef A_

/* ---------------------------------------------------------------- one prediction (fit + predictprob) */
static double ipow(double x, int p) { double r = 1.0; for (int i = 0; i < p; ++i) r = r * x; return r; }
static double drift(const omodel* m, int n, const double* x, int dim) {
  double r = 1.0;
  for (int d = 0; d < dim; ++d) r = r * ipow(x[d], m->exps[n][d]);
  return r;
}
static int krige_one(const ovario* v, const omodel* m, int dim, int n, const double* Xn, const double* zn,
                     const double* x0, double* mu, double* var) {
  int ncon = m->kind == 0 ? 0 : (m->kind == 1 ? 1 : m->ndrift), M = n + ncon;
  double A[MAXM * MAXM], rhs[MAXM], w[MAXM];
  int ipiv[MAXM];
  for (int j = 0; j < n; ++j)
    for (int i = 0; i <= j; ++i) { double c = cov_h(v, sqrt(d2key(Xn + i * dim, Xn + j * dim, dim))); A[j * M + i] = c; A[i * M + j] = c; }
  for (int c = 0; c < ncon; ++c) {
    for (int i = 0; i < n; ++i) { double f = m->kind == 1 ? 1.0 : drift(m, c, Xn + i * dim, dim); A[(n + c) * M + i] = f; A[i * M + n + c] = f; }
    for (int c2 = 0; c2 < ncon; ++c2) A[(n + c) * M + n + c2] = 0.0;
  }
  int info = bk_factor(A, M, M, ipiv);
  for (int i = 0; i < n; ++i) rhs[i] = cov_h(v, sqrt(d2key(Xn + i * dim, x0, dim)));
  for (int c = 0; c < ncon; ++c) rhs[n + c] = m->kind == 1 ? 1.0 : drift(m, c, x0, dim);
  memcpy(w, rhs, sizeof(double) * M);
  bk_solve(A, M, M, ipiv, w);
  double s = 0.0;
  if (m->kind == 0) { for (int i = 0; i < n; ++i) s += w[i] * (zn[i] - m->mean); s += m->mean; }
  else for (int i = 0; i < n; ++i) s += w[i] * zn[i];
  double cl = 0.0, cn = 0.0;
  for (int i = 0; i < n; ++i) cl += rhs[i] * w[i];
  for (int c = 0; c < ncon; ++c) cn += rhs[n + c] * w[n + c];
  *mu = s;
  *var = (v->sill - gamma_h(v, 0.0)) - cl - cn + 1e-10;
  return info;
}

/* ---------------------------------------------------------------- driver with contiguous chunks */
typedef struct {
  const kdtree* tree; const double* X; const double* z; int dim, n; const ovario* v; const omodel* m;
  int mode; /* 0 knn only, 1 krige, 2 idw */ double exponent;
  const double* T; long lo, hi; int k, minn; double radius;
  double* mean; double* var; unsigned char* status; int* nbr; int* cnt;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  int idx[MAXK]; double d2[MAXK], Xn[MAXK * 3], zn[MAXK];
  for (long t = j->lo; t < j->hi; ++t) {
    const double* x0 = j->T + t * j->dim;
    heap_t h; h.n = 0; h.k = j->k;
    if (j->tree) kd_query(j->tree, 0, x0, &h);
    else for (int i = 0; i < j->n; ++i) heap_push(&h, d2key(j->X + (size_t)i * j->dim, x0, j->dim), i);
    int n = h.n;
    heap_sorted(&h, idx, d2);
    if (j->radius > 0) { int keep = 0; while (keep < n && sqrt(d2[keep]) <= j->radius) keep++; n = keep; }
    if (j->cnt) j->cnt[t] = n;
    if (j->nbr) for (int i = 0; i < j->k; ++i) j->nbr[t * j->k + i] = i < n ? idx[i] : -1;
    if (j->mode == 0) continue;
    if (n < j->minn || n == 0) { j->mean[t] = NAN; if (j->var) j->var[t] = NAN; if (j->status) j->status[t] = 1; continue; }
    for (int i = 0; i < n; ++i) { memcpy(Xn + i * j->dim, j->X + (size_t)idx[i] * j->dim, sizeof(double) * j->dim); zn[i] = j->z[idx[i]]; }
    if (j->mode == 1) {
      double mu, var;
      int info = krige_one(j->v, j->m, j->dim, n, Xn, zn, x0, &mu, &var);
      j->mean[t] = mu; if (j->var) j->var[t] = var; if (j->status) j->status[t] = info ? 2 : 0;
    } else {
      double sw = 0; int hit = -1; double w[MAXK];
      for (int i = 0; i < n; ++i) {
        double d = sqrt(d2[i]);
        double p; int ei = (int)j->exponent;
        if ((double)ei == j->exponent) { p = 1.0; for (int q = 0; q < ei; ++q) p = p * d; } else p = pow(d, j->exponent);
        w[i] = 1.0 / p; sw += w[i];
        if (isinf(w[i]) && hit < 0) hit = i;
      }
      if (hit >= 0) j->mean[t] = zn[hit];
      else { double s = 0; for (int i = 0; i < n; ++i) s += (w[i] / sw) * zn[i]; j->mean[t] = s; }
      if (j->status) j->status[t] = 0;
    }
  }
  return 0;
}

/* Exported: targets T are explicit points (M x dim), already scaled like the samples. */
int gsmo_local(const double* X, int dim, int n, const double* z, const ovario* v, const omodel* m, int mode,
               double exponent, const double* T, long M, int k, int minn, double radius, int use_tree,
               int nthreads, double* mean, double* var, unsigned char* status, int* nbr, int* cnt) {
  if (k > MAXK || k < 1 || dim < 1 || dim > 3) return 1;
  if (k > n) k = n;
  kdtree tree; memset(&tree, 0, sizeof(tree));
  if (use_tree) {
    tree.X = X; tree.dim = dim; tree.n = n; tree.leaf = 10;
    tree.perm = (int*)malloc(sizeof(int) * n);
    for (int i = 0; i < n; ++i) tree.perm[i] = i;
    tree.nodes = (kdnode*)malloc(sizeof(kdnode) * (2 * (size_t)n + 4));
    kd_build(&tree, 0, n);
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; job_t jobs[256];
  long per = (M + nthreads - 1) / nthreads;
  int started = 0;
  for (int c = 0; c < nthreads; ++c) {
    long lo = c * per, hi = lo + per > M ? M : lo + per;
    if (lo >= hi) break;
    job_t jb = {use_tree ? &tree : 0, X, z, dim, n, v, m, mode, exponent, T, lo, hi, k, minn, radius, mean, var, status, nbr, cnt};
    jobs[c] = jb;
    pthread_create(&th[c], 0, worker, &jobs[c]);
    started++;
  }
  for (int c = 0; c < started; ++c) pthread_join(th[c], 0);
  if (use_tree) { free(tree.perm); free(tree.nodes); }
  return 0;
}
