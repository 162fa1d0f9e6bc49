// Microbenchmark: cycles per tcgen05.mma (kind::f16, SS mode, SW128 K-major operands in static smem) vs shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../3d-condtional-stable-diffusion_b200/csrc/ptx.cuh"

template <int M, int N, int SBO_A, int PATTERN, int LAYOUT = 2, int SBO_B = 1024>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = ptx::smem_u32(smem);
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  ptx::fence_proxy_async();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::make_idesc_act(M, N);
    const uint64_t da0 = ptx::make_smem_desc(base, 16, SBO_A, (uint64_t)LAYOUT);
    const uint64_t db0 = ptx::make_smem_desc(base + 48 * 1024, 16, SBO_B, (uint64_t)LAYOUT);
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // walk different tiles so operands are not trivially cached: 8 A offsets x 128 B rows, 4 k-steps
          const int accsel = PATTERN == 0 ? (k & 1) : (PATTERN == 1 ? 0 : (k >> 2));
          if (LAYOUT == 4)   // SWIZZLE_64B: 64-byte rows, two K = 16 steps per row; odd k: next tap (+1 row)
            ptx::tc_mma_f16(tm + accsel * N, da0 + (uint64_t)((k & 1) * 2 + (k >> 1) * 4), db0 + (uint64_t)((k & 1) * 2 + (k >> 1) * 384), idesc, 1u);
          else
            ptx::tc_mma_f16(tm + accsel * N, da0 + (uint64_t)((k & 3) * 2 + (k >> 2) * 8), db0 + (uint64_t)((k & 3) * 2), idesc, 1u);
        }
      }
      ptx::tc_commit(ptx::smem_u32(&bar));
    }
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0, nullptr, 0);
    t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

template <int M, int N, int SBO_A, int PATTERN = 0, int LAYOUT = 2, int SBO_B = 1024>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 8);
  auto k = probe<M, N, SBO_A, PATTERN, LAYOUT, SBO_B>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  k<<<grid, 128, 100 * 1024>>>(d, 10);
  k<<<grid, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  double cyc = (double)h / (iters * 8);
  printf("%-28s grid %3d: %7.1f cycles/MMA  -> %6.0f FLOP/cycle/SM (%s)\n", name, grid, cyc, 2.0 * M * N * 16 / cyc, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {148}) {
    run<128, 96, 640, 1, 4, 512>("M128 N96 SW64 same acc", grid);
    run<128, 96, 640, 2, 4, 512>("M128 N96 SW64 4+4 acc", grid);
    run<128, 32, 640, 1, 4, 512>("M128 N32 SW64 same acc", grid);
    run<128, 64, 640, 1, 4, 512>("M128 N64 SW64 same acc", grid);
    run<128, 96, 1280, 1, 2, 1024>("M128 N96 SW128 same acc", grid);
    run<128, 192, 1280, 1, 2, 1024>("M128 N192 SW128 same acc", grid);
    run<64, 8, 1024>("M64 N8 (issue floor?)", grid);
    run<64, 16, 1024>("M64 N16", grid);
    run<64, 32, 1024>("M64 N32", grid);
    run<128, 64, 1280, 1>("M128 N64 same acc", grid);
    run<128, 64, 1280, 2>("M128 N64 4+4 acc", grid);
    run<128, 128, 1280, 1>("M128 N128 same acc", grid);
    run<128, 128, 1280, 2>("M128 N128 4+4 acc", grid);
    run<128, 32, 1280, 2>("M128 N32 4+4 acc", grid);
    run<128, 16, 1024>("M128 N16", grid);
    run<128, 32, 1024>("M128 N32", grid);
    run<128, 64, 1024>("M128 N64", grid);
    run<128, 64, 1280>("M128 N64 (A SBO=1280)", grid);
    run<128, 128, 1024>("M128 N128", grid);
    run<128, 128, 1280>("M128 N128 (A SBO=1280)", grid);
    run<128, 256, 1024>("M128 N256", grid);
    run<64, 64, 1024>("M64 N64", grid);
    run<64, 128, 1024>("M64 N128", grid);
    run<64, 256, 1024>("M64 N256", grid);
  }
  return 0;
}
