// l2_gather_probe.cu -- what is the ceiling for "gather random 512-B rows" on this GPU?
// Pure gather (no FMAs beyond an XOR fold), maximal memory-level parallelism, no sparse structure.
// Used to put the SpMM/SDDMM kernels' L2->SM rates (profiles/) next to a measured ceiling.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_probe scripts/l2_gather_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

template <int U>
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ table, const uint32_t* __restrict__ idx,
                                                     long n_idx, uint32_t row_vec, uint4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long warp = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (long base = warp * 32; base < n_idx; base += nwarps * 32) {
    const uint32_t mine = idx[base + lane];  // n_idx is a multiple of 32
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += U) {
      uint4 b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t r = __shfl_sync(0xffffffffu, mine, j0 + u);
        b[u] = __ldg(table + (size_t)r * row_vec + lane);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x ^= b[u].x; acc.y ^= b[u].y; acc.z ^= b[u].z; acc.w ^= b[u].w; }
    }
  }
  if (acc.x == 0x12345678u) out[warp] = acc;  // keep the loads alive
}

template <int U>
float run(const uint4* table, const uint32_t* idx, long n_idx, uint4* out, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) gather_kernel<U><<<blocks, 256>>>(table, idx, n_idx, 32, out);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) gather_kernel<U><<<blocks, 256>>>(table, idx, n_idx, 32, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const long n_idx = 8388608;  // config 2: nnz per pass
  for (long rows : {65536L, 524288L, 4194304L}) {  // 32 MiB (L2 resident), 256 MiB, 2 GiB tables of 512-B rows
    uint4* table; uint32_t* idx; uint4* out;
    cudaMalloc(&table, rows * 512); cudaMemset(table, 1, rows * 512);
    cudaMalloc(&idx, n_idx * 4); cudaMalloc(&out, 1 << 24);
    std::vector<uint32_t> h(n_idx);
    uint64_t s = 88172645463325252ull;
    for (long i = 0; i < n_idx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % (uint64_t)rows); }
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    for (int bps : {2, 4, 8}) {
      const int blocks = 148 * bps;
      const double gb = n_idx * 512.0 / 1e9;
      float a = run<4>(table, idx, n_idx, out, blocks), b = run<8>(table, idx, n_idx, out, blocks), c = run<16>(table, idx, n_idx, out, blocks);
      printf("table %5ld MiB  CTAs/SM %d : U=4 %.1f GB/s  U=8 %.1f GB/s  U=16 %.1f GB/s\n", rows * 512 >> 20, bps, gb / a * 1e3, gb / b * 1e3, gb / c * 1e3);
    }
    cudaFree(table); cudaFree(idx); cudaFree(out);
  }
  return 0;
}
