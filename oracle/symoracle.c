/*
 * symoracle.c -- C restatement of the packed vector contraction, for full-size checks and CPU timing.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py): only tests/, __graft_entry__.smoke() and bench.py's CPU arms
 * load this; the product never does.
 *
 * The enumeration follows the reference's generator semantics literally -- NOT the closed form used by the
 * CUDA kernels -- so the two are independent statements of the storage order:
 *   _sub_σindex_iter / σindex_iter   symtensor/permcls_symtensor.py:288-347
 *     position k takes ascending values not used by earlier positions; if m_k == m_{k-1} then v_k > v_{k-1}
 *   class order / multiplicity       symtensor/utils.py:839-856, 1000-1002, 760-776 (passed in by the caller)
 *   s = sum_classes gamma * sum_p A[p] * prod x[v]^m   == symalg.contract_all_indices_with_vector
 *                                                     (symtensor/symalg.py:505-527)
 * Accumulation is in long double (64-bit mantissa) so the oracle is accurate to ~1e-18 relative per term.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SO_MAX_RANK 16

typedef struct {
  int l;                 /* distinct values */
  int dim;
  int mult[SO_MAX_RANK];
  const double* x;
} so_cls;

/* number of stored components below a partial assignment (positions k..l-1 still free) */
static int64_t count_rec(const so_cls* c, int k, int* vals, unsigned char* used) {
  const int lo = (k > 0 && c->mult[k] == c->mult[k - 1]) ? vals[k - 1] + 1 : 0;
  int64_t n = 0;
  for (int v = lo; v < c->dim; ++v) {
    if (used[v]) continue;
    if (k == c->l - 1) { ++n; continue; }
    used[v] = 1; vals[k] = v;
    n += count_rec(c, k + 1, vals, used);
    used[v] = 0;
  }
  return n;
}

static long double powi(double x, int m) { long double p = 1.0L; for (int i = 0; i < m; ++i) p *= (long double)x; return p; }

static long double sum_rec(const so_cls* c, int k, int* vals, unsigned char* used, long double w, const double** data) {
  const int lo = (k > 0 && c->mult[k] == c->mult[k - 1]) ? vals[k - 1] + 1 : 0;
  long double acc = 0.0L;
  if (k == c->l - 1) {
    const int m = c->mult[k];
    const double* d = *data;
    for (int v = lo; v < c->dim; ++v) {
      if (used[v]) continue;
      acc += (long double)(*d++) * powi(c->x[v], m);
    }
    *data = d;
    return acc * w;
  }
  for (int v = lo; v < c->dim; ++v) {
    if (used[v]) continue;
    used[v] = 1; vals[k] = v;
    acc += sum_rec(c, k + 1, vals, used, w * powi(c->x[v], c->mult[k]), data);
    used[v] = 0;
  }
  return acc;
}

/* One class: data[size] in storage order.  Parallel over the first value v0 (block offsets by counting). */
double so_class_contract_vec_f64(int l, const int32_t* mult, int dim, const double* data, const double* x, double gamma,
                                 int64_t* size_out) {
  so_cls c;
  c.l = l; c.dim = dim; c.x = x;
  for (int i = 0; i < l; ++i) c.mult[i] = mult[i];
  if (l == 0) { if (size_out) *size_out = 1; return data[0] * gamma; }
  if (l > dim) { if (size_out) *size_out = 0; return 0.0; }
  int64_t* start = (int64_t*)calloc((size_t)dim + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 1)
  for (int v0 = 0; v0 < dim; ++v0) {
    int vals[SO_MAX_RANK]; unsigned char* used = (unsigned char*)calloc((size_t)dim, 1);
    vals[0] = v0; used[v0] = 1;
    start[v0 + 1] = (l == 1) ? 1 : count_rec(&c, 1, vals, used);
    free(used);
  }
  for (int v0 = 0; v0 < dim; ++v0) start[v0 + 1] += start[v0];
  long double total = 0.0L;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int v0 = 0; v0 < dim; ++v0) {
    int vals[SO_MAX_RANK]; unsigned char* used = (unsigned char*)calloc((size_t)dim, 1);
    vals[0] = v0; used[v0] = 1;
    const double* d = data + start[v0];
    const long double w = powi(x[v0], c.mult[0]);
    if (l == 1) total += (long double)d[0] * w;
    else total += sum_rec(&c, 1, vals, used, w, &d);
    free(used);
  }
  if (size_out) *size_out = start[dim];
  free(start);
  return (double)(total * (long double)gamma);
}

/* Representative multi-indices of one class in storage order: out[size * rank] (int32). */
static void dump_rec(const so_cls* c, int k, int* vals, unsigned char* used, int32_t** out) {
  const int lo = (k > 0 && c->mult[k] == c->mult[k - 1]) ? vals[k - 1] + 1 : 0;
  for (int v = lo; v < c->dim; ++v) {
    if (used[v]) continue;
    vals[k] = v;
    if (k == c->l - 1) {
      int32_t* o = *out;
      for (int i = 0; i < c->l; ++i) for (int m = 0; m < c->mult[i]; ++m) *o++ = vals[i];
      *out = o;
    } else {
      used[v] = 1;
      dump_rec(c, k + 1, vals, used, out);
      used[v] = 0;
    }
  }
}

int64_t so_class_dump_index(int l, const int32_t* mult, int dim, int32_t* out) {
  so_cls c;
  c.l = l; c.dim = dim; c.x = 0;
  int rank = 0;
  for (int i = 0; i < l; ++i) { c.mult[i] = mult[i]; rank += mult[i]; }
  if (l == 0) return 1;
  if (l > dim) return 0;
  int vals[SO_MAX_RANK]; unsigned char* used = (unsigned char*)calloc((size_t)dim, 1);
  int32_t* o = out;
  dump_rec(&c, 0, vals, used, &o);
  free(used);
  return (int64_t)(o - out) / rank;
}

int so_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
