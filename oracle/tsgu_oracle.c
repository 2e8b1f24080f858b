/*
 * tsgu_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Plain-C, single-threaded CPU restatement of the sparse_mm hot path of
 * cai4cai/torchsparsegradutils, used as the parity checker for the sm_100a
 * kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * below against fixtures in tests/golden/ that were produced by importing the
 * real reference (tests/golden/make_golden.py) in the build container.
 *
 * What each function restates (paths relative to the reference checkout):
 *   orc_spmm_csr_*      torch.sparse.mm(A, B)            sparse_matmul.py:155
 *   orc_sddmm_*         index_select x2, mul, sum        sparse_matmul.py:201-205
 *   orc_spmm_t_*        torch.sparse.mm(A.t(), grad)     sparse_matmul.py:229
 *   orc_coo_sort        unique(sorted)+argsort(inverse)  utils/utils.py:148-149
 *   orc_compress_rows   bincount + cumsum                utils/utils.py:228-231
 *   orc_decompress_crow repeat_interleave(arange, diff)  utils/utils.py:461-464,
 *                                                        sparse_matmul.py:190-192
 *   orc_csr_transpose   A.t().to_sparse_csr() (implicit) sparse_matmul.py:229
 *
 * The arithmetic itself lives in PyTorch ATen (third party, pyproject.toml:23
 * "torch>=2.5", 2.11.0+cu128 installed here).  Its published semantics are the
 * textbook ones restated here: C[i,:] = sum_e vals[e] * B[col[e],:] over the
 * stored entries of row i, duplicates included.
 *
 * All index arrays are int64.  `acc64` selects double accumulation (a tighter
 * "truth" than the reference's own fp32 arithmetic) or accumulation in the
 * value type in CSR order (what ATen's non-MKL loop does).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ SpMM */
#define DEFINE_SPMM(NAME, T)                                                         \
  ORC_EXPORT void NAME(int64_t n, int64_t K, const int64_t* rowptr,                  \
                       const int64_t* col, const T* vals, const T* B,                \
                       int64_t b_rs, int64_t b_cs, T* C, int acc64) {                \
    double* accd = (double*)malloc(sizeof(double) * (size_t)(K > 0 ? K : 1));        \
    T* acct = (T*)malloc(sizeof(T) * (size_t)(K > 0 ? K : 1));                       \
    for (int64_t i = 0; i < n; ++i) {                                                \
      for (int64_t k = 0; k < K; ++k) { accd[k] = 0.0; acct[k] = (T)0; }             \
      for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {                          \
        const T a = vals[e];                                                         \
        const T* brow = B + col[e] * b_rs;                                           \
        if (acc64)                                                                   \
          for (int64_t k = 0; k < K; ++k) accd[k] += (double)a * (double)brow[k * b_cs]; \
        else                                                                         \
          for (int64_t k = 0; k < K; ++k) acct[k] += a * brow[k * b_cs];             \
      }                                                                              \
      for (int64_t k = 0; k < K; ++k) C[i * K + k] = acc64 ? (T)accd[k] : acct[k];   \
    }                                                                                \
    free(accd);                                                                      \
    free(acct);                                                                      \
  }
DEFINE_SPMM(orc_spmm_csr_f32, float)
DEFINE_SPMM(orc_spmm_csr_f64, double)

/* ----------------------------------------------------------------- SDDMM */
/* out[e] = <G[row[e],:], B[col[e],:]> for every stored entry, in storage order. */
#define DEFINE_SDDMM(NAME, T)                                                        \
  ORC_EXPORT void NAME(int64_t nnz, int64_t K, const int64_t* row,                   \
                       const int64_t* col, const T* G, int64_t g_rs, const T* B,     \
                       int64_t b_rs, int64_t b_cs, T* out, int acc64) {              \
    for (int64_t e = 0; e < nnz; ++e) {                                              \
      const T* g = G + row[e] * g_rs;                                                \
      const T* b = B + col[e] * b_rs;                                                \
      if (acc64) {                                                                   \
        double s = 0.0;                                                              \
        for (int64_t k = 0; k < K; ++k) s += (double)g[k] * (double)b[k * b_cs];     \
        out[e] = (T)s;                                                               \
      } else {                                                                       \
        T s = (T)0;                                                                  \
        for (int64_t k = 0; k < K; ++k) s += g[k] * b[k * b_cs];                     \
        out[e] = s;                                                                  \
      }                                                                              \
    }                                                                                \
  }
DEFINE_SDDMM(orc_sddmm_f32, float)
DEFINE_SDDMM(orc_sddmm_f64, double)

/* ------------------------------------------------- gradB = A^T * G (scatter) */
#define DEFINE_SPMM_T(NAME, T)                                                       \
  ORC_EXPORT void NAME(int64_t nnz, int64_t m, int64_t K, const int64_t* row,        \
                       const int64_t* col, const T* vals, const T* G, int64_t g_rs,  \
                       T* out, int acc64) {                                          \
    size_t tot = (size_t)m * (size_t)K;                                              \
    double* acc = (double*)calloc(tot ? tot : 1, sizeof(double));                    \
    for (size_t t = 0; t < tot; ++t) out[t] = (T)0;                                  \
    for (int64_t e = 0; e < nnz; ++e) {                                              \
      const T a = vals[e];                                                           \
      const T* g = G + row[e] * g_rs;                                                \
      if (acc64)                                                                     \
        for (int64_t k = 0; k < K; ++k) acc[col[e] * K + k] += (double)a * (double)g[k]; \
      else                                                                           \
        for (int64_t k = 0; k < K; ++k) out[col[e] * K + k] += a * g[k];             \
    }                                                                                \
    if (acc64)                                                                       \
      for (size_t t = 0; t < tot; ++t) out[t] = (T)acc[t];                           \
    free(acc);                                                                       \
  }
DEFINE_SPMM_T(orc_spmm_t_f32, float)
DEFINE_SPMM_T(orc_spmm_t_f64, double)

/* ------------------------------------------------------------- COO sort */
/* Stable lexicographic sort of the columns of idx (ndim x nnz, row-major).
 * perm[k] = original position of the k-th sorted coordinate.  For duplicate-free
 * input this equals argsort(unique(...).inverse) of utils/utils.py:148-149; for
 * duplicates the tie order is "original position" (documented extension). */
static int64_t g_ndim, g_nnz;
static const int64_t* g_idx;
static int lex_less_eq(int64_t a, int64_t b) {
  for (int64_t d = 0; d < g_ndim; ++d) {
    int64_t x = g_idx[d * g_nnz + a], y = g_idx[d * g_nnz + b];
    if (x != y) return x < y;
  }
  return 1; /* equal keys: keep left first (stable) */
}
static void merge_sort(int64_t* a, int64_t* tmp, int64_t lo, int64_t hi) {
  if (hi - lo < 2) return;
  int64_t mid = lo + (hi - lo) / 2;
  merge_sort(a, tmp, lo, mid);
  merge_sort(a, tmp, mid, hi);
  int64_t i = lo, j = mid, k = lo;
  while (i < mid && j < hi) tmp[k++] = lex_less_eq(a[i], a[j]) ? a[i++] : a[j++];
  while (i < mid) tmp[k++] = a[i++];
  while (j < hi) tmp[k++] = a[j++];
  memcpy(a + lo, tmp + lo, sizeof(int64_t) * (size_t)(hi - lo));
}
ORC_EXPORT void orc_coo_sort(int64_t ndim, int64_t nnz, const int64_t* idx,
                             int64_t* sorted, int64_t* perm) {
  g_ndim = ndim; g_nnz = nnz; g_idx = idx;
  int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
  for (int64_t e = 0; e < nnz; ++e) perm[e] = e;
  merge_sort(perm, tmp, 0, nnz);
  for (int64_t d = 0; d < ndim; ++d)
    for (int64_t e = 0; e < nnz; ++e) sorted[d * nnz + e] = idx[d * nnz + perm[e]];
  free(tmp);
}

/* ------------------------------------------------ row compress / expand */
ORC_EXPORT void orc_compress_rows(int64_t nnz, const int64_t* rows, int64_t n, int64_t* crow) {
  for (int64_t i = 0; i <= n; ++i) crow[i] = 0;
  for (int64_t e = 0; e < nnz; ++e) crow[rows[e] + 1] += 1; /* bincount */
  for (int64_t i = 0; i < n; ++i) crow[i + 1] += crow[i];   /* cumsum   */
}
ORC_EXPORT void orc_decompress_crow(int64_t n, const int64_t* crow, int64_t* rows) {
  for (int64_t i = 0; i < n; ++i)
    for (int64_t e = crow[i]; e < crow[i + 1]; ++e) rows[e] = i;
}

/* -------------------------------------------------------- CSR transpose */
/* Counting sort by column, stable in storage order: permT[k] = position in A's
 * storage of the k-th entry of A^T; colT[k] = its row in A. */
ORC_EXPORT void orc_csr_transpose(int64_t n, int64_t m, const int64_t* rowptr,
                                  const int64_t* col, int64_t* rowptrT, int64_t* colT,
                                  int64_t* permT) {
  int64_t nnz = rowptr[n];
  for (int64_t j = 0; j <= m; ++j) rowptrT[j] = 0;
  for (int64_t e = 0; e < nnz; ++e) rowptrT[col[e] + 1] += 1;
  for (int64_t j = 0; j < m; ++j) rowptrT[j + 1] += rowptrT[j];
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m > 0 ? m : 1));
  memcpy(cur, rowptrT, sizeof(int64_t) * (size_t)m);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      int64_t dst = cur[col[e]]++;
      colT[dst] = i;
      permT[dst] = e;
    }
  free(cur);
}
