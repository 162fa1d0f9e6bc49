// 3^3 stride-1 'same' Conv3D with ONE output channel and 32 input channels (the vqgan_attn_cp decoder's head,
// networks/vqgan_attn_cp.py:424-427: Conv3D(out_channels=1, 3, padding='same') on the 128^3 x 32 feature volume).
//
// HBM-bound: 64 B read + 4 B written per voxel, 1728 FLOP.  GEMM-N = 1 wastes an N = 16 tensor-core tile 16x and re-reads
// the A operand from shared memory once per tap (the halo kernel ran this layer at 0.56 TB/s).  Here the channel
// contraction of ALL 27 taps is one small GEMM per voxel, P[v][tap] = sum_c x[v][c] * w[tap][c]  (M = voxels of a halo
// plane, N = 27 -> 32, K = 32: warp-level mma.sync.m16n8k16, A straight from the TMA-written plane via ldmatrix, weights
// resident in registers), and the spatial part is a 27-point shifted sum of P -- every input voxel is read from shared
// memory once.  A CTA sweeps a (8 h x 32 w) column along d: input plane p contributes kd = 0/1/2 to the output planes
// p+1 / p / p-1, whose partial sums live in registers (one output position per thread), so there is no halo in d.
//   per input plane:  wait TMA  ->  P = X W^T (8 warps x <=3 m16 tiles)  ->  sync  ->  (refill the plane buffer) 27 LDS + adds
//                     -> store the finished plane  ->  sync
// Two CTAs share an SM (84 KB each): one runs its MMA phase while the other gathers.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace stencil {

constexpr int kTH = 8, kTW = 32;              // output tile of one plane
constexpr int kHH = kTH + 2, kHW = kTW + 2;   // halo plane 10 x 34
constexpr int kNV = kHH * kHW;                // 340 voxels
constexpr int kMT = (kNV + 15) / 16;          // 22 m16 tiles (rows 340..351 are padding)
constexpr int kPV = 356;                      // P row stride in words: >= 16 * kMT and == 4 (mod 32) -> conflict-free fragment stores
constexpr int kC = 32;                        // input channels = K of the GEMM
constexpr int kInBytes = kMT * 16 * kC * 2;   // 22528 (multiple of 512: SWIZZLE_64B atoms stay aligned)
constexpr int kBoxBytes = kNV * kC * 2;       // 21760 bytes per TMA box
constexpr int kThreads = 256;
constexpr size_t kSmem = 1024 + 2 * kInBytes + 27 * kPV * 4 + 64;

struct Params {
  int batch, D, H, W;
  int tiles_h, tiles_w, dsplit, dlen;   // work item = (sample, h tile, w tile, d range of dlen planes)
  int items;
  const act_t* w;      // packed weights of output channel 0: [tap][64] (channels 0..31 used)
  const float* bias;   // (1) or null
  int act, y_f32;
  void* y;
  int* dbg;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef B200DM_ACT_FP16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
#endif
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct Item { int n, h0, w0, d0, d1; };
__device__ __forceinline__ Item decode(const Params& p, int item) {
  Item it;
  int r = item;
  const int ds = r % p.dsplit; r /= p.dsplit;
  const int tw = r % p.tiles_w; r /= p.tiles_w;
  const int th = r % p.tiles_h; r /= p.tiles_h;
  it.n = r; it.h0 = th * kTH; it.w0 = tw * kTW;
  it.d0 = ds * p.dlen; it.d1 = min(p.D, it.d0 + p.dlen);
  return it;
}

__global__ void __launch_bounds__(kThreads, 2) conv_stencil_c1_kernel(const __grid_constant__ CUtensorMap mapX, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t in_addr = base;                                  // 2 plane buffers
  float* P = reinterpret_cast<float*>(sm + 2 * kInBytes);         // [27][kPV]
  const uint32_t bar0 = base + 2 * kInBytes + 27 * kPV * 4;       // 2 mbarriers
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;

  pdl_launch_dependents();
  if (tid == 0) {
    ptx::prefetch_tmap(&mapX);
    ptx::mbar_init(bar0, 1);
    ptx::mbar_init(bar0 + 8, 1);
    ptx::fence_barrier_init();
  }
  // weights -> B fragments (col-major K x N: b0 = W[tap = nt*8 + g][c = ks*16 + 2*t4 .. +1], b1 = the same 8 channels up)
  uint32_t bw[4][2][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int tap = nt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(p.w + (size_t)min(tap, 26) * 64 + ks * 16 + 2 * t4);
      bw[nt][ks][0] = tap < 27 ? __ldg(wp) : 0u;
      bw[nt][ks][1] = tap < 27 ? __ldg(wp + 4) : 0u;
    }
  }
  const float bias = p.bias ? __ldg(p.bias) : 0.0f;
  __syncthreads();
  pdl_wait();

  // flattened (item, plane) sequence of this CTA; the producer (thread 0) runs two planes ahead of the consumers
  int pr_item = blockIdx.x, pr_plane = 0, pr_count = 0;   // producer cursor; plane index relative to d0 - 1
  Item pr_it = pr_item < p.items ? decode(p, pr_item) : Item{0, 0, 0, 0, 0};
  auto produce = [&]() {   // thread 0 only
    if (pr_item >= p.items) return;
    const uint32_t bar = bar0 + (pr_count & 1) * 8;
    ptx::mbar_expect_tx(bar, kBoxBytes);
    ptx::tma_load_5d(in_addr + (pr_count & 1) * kInBytes, &mapX, bar, 0, pr_it.w0 - 1, pr_it.h0 - 1, pr_it.d0 - 1 + pr_plane, pr_it.n);
    ++pr_count;
    if (++pr_plane == pr_it.d1 - pr_it.d0 + 2) {
      pr_plane = 0;
      pr_item += gridDim.x;
      if (pr_item < p.items) pr_it = decode(p, pr_item);
    }
  };
  if (tid == 0) { produce(); produce(); }

  const int oh = wi, ow = lane;   // this thread's output position inside the tile
  int count = 0;
  for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
    const Item it = decode(p, item);
    const int nplanes = it.d1 - it.d0 + 2;
    const bool inside = it.h0 + oh < p.H && it.w0 + ow < p.W;
    float a_prev = 0.0f, a_cur = 0.0f;
    for (int pl = 0; pl < nplanes; ++pl, ++count) {
      const int b = count & 1;
      if (!ptx::mbar_wait(bar0 + b * 8, (count >> 1) & 1, p.dbg, 0x5701)) return;
      // ---- P[tap][v] = sum_c X[v][c] W[tap][c] for the 340 voxels of the halo plane
      const uint32_t xin = in_addr + b * kInBytes;
      for (int mt = wi; mt < kMT; mt += 8) {
        float c[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.0f; }
        const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t a[4];
          const int chunk = ks * 2 + (lane >> 4);
          ldmatrix_x4(xin + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4), a);   // SWIZZLE_64B: 16-byte chunk ^= (row / 2) % 4
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma16816(c[nt], a, bw[nt][ks][0], bw[nt][ks][1]);
        }
        float* pr = P + mt * 16 + g;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int tap = nt * 8 + 2 * t4;
          if (tap < 27) { pr[tap * kPV] = c[nt][0]; pr[tap * kPV + 8] = c[nt][2]; }
          if (tap + 1 < 27) { pr[(tap + 1) * kPV] = c[nt][1]; pr[(tap + 1) * kPV + 8] = c[nt][3]; }
        }
      }
      __syncthreads();              // P complete, plane buffer b consumed
      if (tid == 0) produce();      // refill buffer b with the plane after next
      // ---- 27-point shifted sum: input plane q = d0 - 1 + pl feeds output planes q + 1 (kd = 0), q (kd = 1), q - 1 (kd = 2)
      float s[3];
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        float acc = 0.0f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) acc += P[(kd * 9 + kh * 3 + kw) * kPV + (oh + kh) * kHW + ow + kw];
        s[kd] = acc;
      }
      const int j = it.d0 + pl - 2;   // the plane that receives its last (kd = 2) contribution now
      const float done = a_prev + s[2];
      a_prev = a_cur + s[1];
      a_cur = s[0];
      if (pl >= 2 && inside) {
        const float v = apply_act(done + bias, p.act);
        const size_t o = (((size_t)it.n * p.D + j) * p.H + it.h0 + oh) * p.W + it.w0 + ow;
        if (p.y_f32) reinterpret_cast<float*>(p.y)[o] = v;
        else reinterpret_cast<act_t*>(p.y)[o] = float_to_act(v);
      }
      __syncthreads();              // P free for the next plane
    }
  }
}

}  // namespace stencil
