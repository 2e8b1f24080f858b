// tma_gather_probe.cu -- ceiling of "gather random 512-B rows" through the TMA unit on sm_100a:
// cp.async.bulk.tensor.2d ... tile::gather4 (4 rows per instruction) into a shared-memory ring,
// completion on mbarriers, consumers read the rows back with LDS.128 (as an SpMM would).
// Companion of l2_gather_probe.cu (LDG path).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_probe scripts/tma_gather_probe.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(s32(dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(bar)) : "memory");
}

constexpr int SLOT_BYTES = 2048;  // 4 rows x 512 B
template <int SLOTS, bool READ>
__global__ void __launch_bounds__(288) tma_gather_kernel(const __grid_constant__ CUtensorMap map, const int4* __restrict__ idx4,
                                                         long n_g4, uint4* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SLOTS * SLOT_BYTES);
  uint64_t* empty = full + SLOTS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long per = (n_g4 + gridDim.x - 1) / gridDim.x;
  const long g_lo = (long)blockIdx.x * per;
  const long g_hi = g_lo + per < n_g4 ? g_lo + per : n_g4;
  if (warp == 0) {  // producer: lane l feeds slots l, l+32, ...
    long k = 0;
    for (long g = g_lo + lane; g < g_hi; g += 32, ++k) {
      const int slot = (int)((k * 32 + lane) % SLOTS);
      const long use = (k * 32 + lane) / SLOTS;
      if (use > 0) mbar_wait(&empty[slot], (uint32_t)((use - 1) & 1));
      const int4 r = __ldg(idx4 + g);
      mbar_expect(&full[slot], SLOT_BYTES);
      tma_gather4(smem + slot * SLOT_BYTES, &map, 0, r.x, r.y, r.z, r.w, &full[slot]);
    }
  } else {  // 8 consumer warps: warp c takes gathers c-1, c-1+8, ...
    uint4 acc = make_uint4(0, 0, 0, 0);
    const long n_local = g_hi - g_lo;
    for (long q = warp - 1; q < n_local; q += 8) {
      const int slot = (int)(q % SLOTS);
      const long use = q / SLOTS;
      mbar_wait(&full[slot], (uint32_t)(use & 1));
      if (READ) {
        const uint4* rows = reinterpret_cast<const uint4*>(smem + slot * SLOT_BYTES);
#pragma unroll
        for (int j = 0; j < 4; ++j) { uint4 v = rows[j * 32 + lane]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
    if (acc.x == 0x12345678u) out[blockIdx.x * 8 + warp] = acc;
  }
}

template <int SLOTS, bool READ>
float run(const CUtensorMap& map, const int4* idx4, long n_g4, uint4* out) {
  const int smem = SLOTS * SLOT_BYTES + 2 * SLOTS * 8;
  cudaFuncSetAttribute(tma_gather_kernel<SLOTS, READ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) tma_gather_kernel<SLOTS, READ><<<148, 288, smem>>>(map, idx4, n_g4, out);
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) tma_gather_kernel<SLOTS, READ><<<148, 288, smem>>>(map, idx4, n_g4, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1.f; }
  return ms / 10;
}

int main() {
  const long n_idx = 8388608;
  cuInit(0);
  for (long rows : {65536L, 524288L}) {
    float* table; int* idx; uint4* out;
    cudaMalloc(&table, rows * 512); cudaMemset(table, 1, rows * 512);
    cudaMalloc(&idx, n_idx * 4); cudaMalloc(&out, 1 << 20);
    std::vector<int> h(n_idx);
    uint64_t s = 88172645463325252ull;
    for (long i = 0; i < n_idx; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)rows); }
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t gdim[2] = {128, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {512};
    cuuint32_t box[2] = {128, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstride, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    const double gb = n_idx * 512.0 / 1e9;
    float a = run<32, false>(map, (const int4*)idx, n_idx / 4, out);
    float b = run<64, false>(map, (const int4*)idx, n_idx / 4, out);
    float c = run<96, false>(map, (const int4*)idx, n_idx / 4, out);
    float d = run<64, true>(map, (const int4*)idx, n_idx / 4, out);
    float e = run<96, true>(map, (const int4*)idx, n_idx / 4, out);
    printf("table %4ld MiB: TMA gather4 only  32/64/96 slots: %.1f / %.1f / %.1f GB/s ; + LDS read-back 64/96 slots: %.1f / %.1f GB/s\n",
           rows * 512 >> 20, gb / a * 1e3, gb / b * 1e3, gb / c * 1e3, gb / d * 1e3, gb / e * 1e3);
    cudaFree(table); cudaFree(idx); cudaFree(out);
  }
  return 0;
}
