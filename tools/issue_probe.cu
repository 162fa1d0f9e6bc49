// Microbenchmark: cost of one weight-stage iteration of the halo kernel's MMA-issue loop
// (barrier wait + fence + elect + TPS*TD*4 UTCHMMA + commit + syncwarp).  Variants isolate each piece.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../3d-condtional-stable-diffusion_b200/csrc/ptx.cuh"

// VARIANT bit0: barrier wait, bit1: commit, bit2: runtime-varying descriptor bases (else loop-invariant)
template <int VARIANT, int N, int TPS, int TD>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters, int rt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[8];
  __shared__ uint32_t tptr;
  const uint32_t base = ptx::smem_u32(smem);
  for (int i = threadIdx.x; i < (215 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(ptx::smem_u32(&tptr), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  ptx::fence_proxy_async();
  const uint32_t tm = tptr;
  const uint32_t bar0 = ptx::smem_u32(&bars[0]);
  constexpr int kSlab = 23552, kTap = N * 128, kStage = TPS * kTap;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint64_t a0 = ptx::make_smem_desc(base, 16, 1280, ptx::kLayoutSw128);
    const uint64_t b0 = ptx::make_smem_desc(base + 6 * kSlab, 16, 1024, ptx::kLayoutSw128);
    long long t0 = clock64();
    uint32_t sb = 0, q = rt;
    for (int i = 0; i < iters; ++i) {
      if (VARIANT & 1) ptx::mbar_wait(bar0 + 8 * (sb & 3), 1, nullptr, 0);   // parity 1 on a fresh barrier: already complete
      ptx::tc_fence_after();
      const int kh = i % 3;
      uint64_t a_pl[TD];
#pragma unroll
      for (int pl = 0; pl < TD; ++pl) a_pl[pl] = a0 + (uint64_t)(((VARIANT & 4) ? (q + pl) % 4 : pl) * (kSlab >> 4));
      if (ptx::elect_one()) {
        const uint64_t db0 = b0 + (uint64_t)(((VARIANT & 4) ? sb % 3 : 0) * (kStage >> 4));
#pragma unroll
        for (int u = 0; u < TPS; ++u)
#pragma unroll
          for (int pl = 0; pl < TD; ++pl) {
            const uint64_t da = a_pl[pl] + (uint64_t)(((VARIANT & 4) ? kh : 1) * 80 + u * 8);
            const uint64_t db = db0 + (uint64_t)(u * (kTap >> 4));
            ptx::tc_mma_f16(tm + pl * N, da, db, idesc, 1u);
            ptx::tc_mma_f16(tm + pl * N, da + 2, db + 2, idesc, 1u);
            ptx::tc_mma_f16(tm + pl * N, da + 4, db + 4, idesc, 1u);
            ptx::tc_mma_f16(tm + pl * N, da + 6, db + 6, idesc, 1u);
          }
        if (VARIANT & 2) ptx::tc_commit(bar0 + 8 * 4 + 8 * (sb & 1));
      }
      __syncwarp();
      ++sb; ++q;
    }
    if (ptx::elect_one()) ptx::tc_commit(bar0 + 8 * 7);
    __syncwarp();
    ptx::mbar_wait(bar0 + 8 * 7, 0, nullptr, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { ptx::tc_fence_after(); ptx::tmem_dealloc(tm, 512); }
}

template <int VARIANT, int N, int TPS, int TD>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  auto k = probe<VARIANT, N, TPS, TD>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 3000;
  k<<<148, 128, 220 * 1024>>>(d, 10, 0);
  k<<<148, 128, 220 * 1024>>>(d, iters, 1);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const int nm = TPS * TD * 4, fl = nm * (N == 64 ? 48 : (N == 128 ? 64 : 40));
  printf("%-52s %8.1f cycles/stage (%2d MMAs: %5.1f each; tensor floor %d) %s\n", name, (double)h / iters, nm, (double)h / iters / nm, fl, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<0, 64, 3, 2>("N64 TPS3 TD2: invariant descs, no wait/commit");
  run<4, 64, 3, 2>("N64 TPS3 TD2: runtime descs, no wait/commit");
  run<5, 64, 3, 2>("N64 TPS3 TD2: runtime descs + wait");
  run<7, 64, 3, 2>("N64 TPS3 TD2: runtime descs + wait + commit");
  run<7, 64, 1, 2>("N64 TPS1 TD2: all");
  run<7, 128, 1, 2>("N128 TPS1 TD2: all");
  run<7, 32, 3, 2>("N32 TPS3 TD2: all");
  return 0;
}
