// spmm.cu -- row-split CSR SpMM for sm_100a:  C[t] = A[t] * B[t].
//
// Replaces torch.sparse.mm at reference sparse_matmul.py:155 (forward) and :229 (grad_B, with
// the transposed structure from tsgu_csr_transpose + a value permutation).
//
// Mapping: a group of LPR lanes owns one row of A; the lanes tile the dense K dimension with
// 128-bit loads (EPV elements each), VPL vectors per lane, so one B-row fetch is a single fully
// coalesced request per vector slot.  The group loads LPR (col, val) pairs with one coalesced
// request, then broadcasts them with shuffles and keeps U independent B-row loads in flight
// before the FMA chain (HBM/L2-latency hiding by ILP; no tensor cores: 2 flop per 4-16 B).
// Accumulation is in CSR order inside a row (deterministic, no atomics).
#include "common.cuh"

namespace tsgu {

template <typename V, typename I>
struct SpmmParams {
  const I* rowptr;
  const I* colind;
  const V* vals;
  const I* perm;  // nullable
  const V* B;
  V* C;
  int64_t batch, n, K;
  int64_t rowptr_bstride, nnz_bstride;
  int64_t b_bs, b_rs, b_cs, c_bs, ldc;
};

// EPV: elements per vector load (1 = scalar path that also honours b_cs != 1)
// LPR: lanes per row (power of two <= 32); VPL: vectors per lane
template <typename V, typename I, int EPV, int LPR, int VPL>
__global__ void __launch_bounds__(256) spmm_rowsplit_kernel(const SpmmParams<V, I> p) {
  using Acc = typename VT<V>::Acc;
  static_assert(EPV == 1 || EPV * sizeof(V) == 16, "vector path is 128-bit");
  constexpr int U = (VPL >= 4) ? 2 : (VPL == 2 ? 4 : 8);  // B-row loads in flight per lane = U*VPL
  constexpr int CHUNK = LPR * VPL * EPV;                   // K elements covered per pass

  const int lane = threadIdx.x & 31;
  const int gl = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const int64_t groups_per_block = blockDim.x / LPR;
  const int64_t total_rows = p.batch * p.n;

  for (int64_t r = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / LPR; r < total_rows;
       r += (int64_t)gridDim.x * groups_per_block) {
    const int64_t item = (p.batch == 1) ? 0 : r / p.n;
    const int64_t lr = r - item * p.n;
    const I* rp = p.rowptr + item * p.rowptr_bstride + lr;
    const int64_t e0 = (int64_t)__ldg(rp) + item * p.nnz_bstride;
    const int64_t e1 = (int64_t)__ldg(rp + 1) + item * p.nnz_bstride;
    const V* Bi = p.B + item * p.b_bs;
    V* Crow = p.C + item * p.c_bs + lr * p.ldc;

    for (int64_t k0 = 0; k0 < p.K; k0 += CHUNK) {
      Acc acc[VPL][EPV];
#pragma unroll
      for (int w = 0; w < VPL; ++w)
#pragma unroll
        for (int i = 0; i < EPV; ++i) acc[w][i] = Acc(0);

      for (int64_t base = e0; base < e1; base += LPR) {
        const int64_t e = base + gl;
        const bool ok = e < e1;
        I c = ok ? __ldg(p.colind + e) : I(0);
        Acc v = Acc(0);
        if (ok) v = load_scalar<V>(p.vals + (p.perm ? (int64_t)__ldg(p.perm + e) : e));
        const int cnt = (int)min((int64_t)LPR, e1 - base);
        for (int j0 = 0; j0 < cnt; j0 += U) {
          Raw<V, EPV> b[U][VPL];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int j = j0 + u;
            const int64_t cj = (int64_t)shfl_idx(gmask, c, j, LPR);
            const V* brow = Bi + cj * p.b_rs;
#pragma unroll
            for (int w = 0; w < VPL; ++w) {
              const int64_t kk = k0 + (int64_t)(w * LPR + gl) * EPV;
              if (j < cnt && kk < p.K)
                b[u][w] = raw_ldg<V, EPV>(brow + (EPV == 1 ? kk * p.b_cs : kk));
              else
                b[u][w] = raw_zero<V, EPV>();
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const Acc vj = shfl_idx(gmask, v, j0 + u, LPR);  // lanes >= cnt hold v = 0
#pragma unroll
            for (int w = 0; w < VPL; ++w) {
              Acc x[EPV];
              raw_unpack<V, EPV>(b[u][w], x);
#pragma unroll
              for (int i = 0; i < EPV; ++i) acc[w][i] = fma(vj, x[i], acc[w][i]);
            }
          }
        }
      }
#pragma unroll
      for (int w = 0; w < VPL; ++w) {
        const int64_t kk = k0 + (int64_t)(w * LPR + gl) * EPV;
        if (kk < p.K) store_vec<V, EPV>(Crow + kk, acc[w]);
      }
    }
  }
}

template <typename V, typename I, int EPV, int LPR, int VPL>
static int launch_rowsplit(const SpmmParams<V, I>& p, cudaStream_t s) {
  const int64_t total_rows = p.batch * p.n;
  const int threads = 256;
  const int64_t gpb = threads / LPR;
  int64_t blocks = (total_rows + gpb - 1) / gpb;
  if (blocks > 0x7fffffffLL) blocks = 0x7fffffffLL;
  spmm_rowsplit_kernel<V, I, EPV, LPR, VPL><<<(unsigned)blocks, threads, 0, s>>>(p);
  count_launch();
  return launch_status();
}

template <typename V, typename I>
static int spmm_dispatch(const SpmmParams<V, I>& p, cudaStream_t s) {
  constexpr int EPVF = 16 / sizeof(V);
  const bool vec_ok = p.b_cs == 1 && (p.K % EPVF) == 0 && (p.b_rs % EPVF) == 0 && (p.b_bs % EPVF) == 0 &&
                      (p.ldc % EPVF) == 0 && (p.c_bs % EPVF) == 0 && aligned16(p.B) && aligned16(p.C);
  if (vec_ok) {
    const int64_t kv = p.K / EPVF;  // vectors per row
    if (kv <= 4) return launch_rowsplit<V, I, EPVF, 4, 1>(p, s);
    if (kv <= 8) return launch_rowsplit<V, I, EPVF, 8, 1>(p, s);
    if (kv <= 16) return launch_rowsplit<V, I, EPVF, 16, 1>(p, s);
    if (kv <= 32) return launch_rowsplit<V, I, EPVF, 32, 1>(p, s);
    if (kv <= 64) return launch_rowsplit<V, I, EPVF, 32, 2>(p, s);
    return launch_rowsplit<V, I, EPVF, 32, 4>(p, s);  // K > 128 vectors: chunk loop inside
  }
  if (p.K == 1) return launch_rowsplit<V, I, 1, 1, 1>(p, s);
  if (p.K <= 4) return launch_rowsplit<V, I, 1, 4, 1>(p, s);
  return launch_rowsplit<V, I, 1, 32, 1>(p, s);
}

}  // namespace tsgu

extern "C" int tsgu_spmm_csr(const void* rowptr, const void* colind, const void* vals, const void* perm,
                             const void* B, void* C, int64_t batch, int64_t n, int64_t m, int64_t K,
                             int64_t rowptr_bstride, int64_t nnz_bstride, int64_t nnz_total,
                             int64_t b_bs, int64_t b_rs, int64_t b_cs, int64_t c_bs, int64_t ldc,
                             int val_dtype, int idx_dtype, int algo, void* workspace,
                             size_t workspace_bytes, void* stream) {
  (void)m; (void)nnz_total; (void)workspace; (void)workspace_bytes;
  if (batch < 0 || n < 0 || K < 0) return TSGU_ERR_SHAPE;
  if (algo != TSGU_ALGO_AUTO && algo != TSGU_ALGO_ROWSPLIT && algo != TSGU_ALGO_MERGE) return TSGU_ERR_ALGO;
  if (batch == 0 || n == 0 || K == 0) return 0;
  TSGU_DISPATCH_VAL(val_dtype, TSGU_DISPATCH_IDX(idx_dtype, {
    tsgu::SpmmParams<V, I> p;
    p.rowptr = (const I*)rowptr; p.colind = (const I*)colind; p.vals = (const V*)vals;
    p.perm = (const I*)perm; p.B = (const V*)B; p.C = (V*)C;
    p.batch = batch; p.n = n; p.K = K;
    p.rowptr_bstride = rowptr_bstride; p.nnz_bstride = nnz_bstride;
    p.b_bs = b_bs; p.b_rs = b_rs; p.b_cs = b_cs; p.c_bs = c_bs; p.ldc = ldc;
    return tsgu::spmm_dispatch<V, I>(p, tsgu::as_stream(stream));
  }));
  return 0;
}

extern "C" size_t tsgu_spmm_workspace_bytes(int64_t batch, int64_t n, int64_t K, int64_t nnz_total,
                                            int val_dtype, int algo) {
  (void)batch; (void)n; (void)K; (void)nnz_total; (void)val_dtype; (void)algo;
  return 0;
}
