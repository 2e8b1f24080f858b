// merge.cuh -- entry points of the merge-path kernels (defined in merge.cu).
#pragma once
#include "common.cuh"

namespace tsgu {
size_t spmm_merge_workspace_bytes(int64_t rows, int64_t K, int64_t nnz, int val_dtype);
size_t sddmm_merge_workspace_bytes(int64_t rows, int64_t nnz);
template <typename V, typename I>
int spmm_merge_dispatch(const I* rowptr, const I* colind, const V* vals, const I* perm, const V* B, V* C, int64_t rows,
                        int64_t K, int64_t nnz, int64_t b_rs, int64_t ldc, void* ws, size_t ws_bytes, cudaStream_t s);
template <typename V, typename I>
int sddmm_merge_dispatch(const I* rowptr, const I* colind, const I* out_index, const V* G, const V* B, V* out,
                         int64_t rows, int64_t K, int64_t nnz, int64_t g_rs, int64_t b_rs, void* ws, size_t ws_bytes,
                         cudaStream_t s);
}  // namespace tsgu
