// K1-K5 (+ the attention matmuls): Conv3D as an implicit GEMM on the 5th-gen tensor cores.
//
//   GEMM view   M = output voxels (tile = a 128-voxel box bw x bh x bd x bn of the NDHWC tensor)
//               N = C_out (BLOCK_N columns of TMEM, fp32 accumulators)
//               K = taps x C_in, walked as k-blocks of (tap, 64-channel chunk)
//   A operand   TMA (cp.async.bulk.tensor.5d, SWIZZLE_128B) loads the box shifted by the tap offset straight
//               from the bf16 NDHWC activation: out-of-bounds voxels / channels are zero-filled by the TMA
//               unit, which IS TF 'same' padding (incl. the asymmetric (0,1) of stride-2) -- no im2col
//               buffer, no padded copy.  Two K-segments (two tensor maps) read [x, skip] without ever
//               materialising layers.Concatenate.
//   B operand   packed bf16 weights [group][n_pad][chunk][tap][64], one 2-D TMA per k-block.
//   MMA         tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16 x4 per k-block, issued by one
//               thread; accumulators live in TMEM; tcgen05.commit releases smem stages / signals the epilogue.
//   Epilogue    4 warps: tcgen05.ld -> +bias +temb[t][n] -> PReLU/act -> +residual -> act -> bf16|fp32 NDHWC
//               (optionally transposed per sample, for V^T of the attention blocks).
//   PARITY mode nearest-upsample(2)+Conv3 and ConvTranspose(k4,s2) both become 8 output-parity
//               sub-convolutions of 2^3 taps on the low-res input (blockIdx.z = parity): 27 -> 8 taps.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <new>
#include <vector>
#include "conv_common.cuh"

// Tuning / experiment switches (B200DM_NO_PAIR, B200DM_CLUSTER, B200DM_KSPLIT, ...) are read only when the process also sets
// B200DM_TUNING=1: a production process's environment cannot reconfigure the kernels by accident.
static const char* tuning_env(const char* name) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("B200DM_TUNING"); on = (e && e[0] == '1') ? 1 : 0; }
  return on ? getenv(name) : nullptr;
}

#include "conv_halo.cuh"
#include "conv_halo_up.cuh"
#include "conv_stencil.cuh"
#include "conv_sweep.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16

// CMODE 0: plain; 1: cluster with TMA-multicast operands (experiment); 2: split-K cluster (K range per CTA, partial
// accumulators reduced through distributed shared memory by the cluster's first CTA).
// (two-stage instantiations: registers capped so that four CTAs share an SM -- see conv_plan_create)
template <int BLOCK_N, int NSTAGE, int CMODE>
__global__ void __launch_bounds__(kThreads, NSTAGE == 2 ? 4 : 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                  const __grid_constant__ CUtensorMap mapB, const __grid_constant__ ConvOutMaps om, const ConvParams p) {
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr int kResBytes = BLOCK_N >= 64 ? BLOCK_N * 256 : 0;   // staged epilogue: 128 rows x BLOCK_N bf16 residual tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const bool staged = BLOCK_N >= 64 && CMODE != 1 && p.tma_epi != 0;
  const bool staged_res = staged && p.residual != nullptr;
  uint8_t* res_smem = smem + NSTAGE * kStageBytes;
  const bool side_on = BLOCK_N >= 64 && CMODE == 0 && p.side != 0 && blockIdx.y == 0;   // n-tile 0 writes the side output
  uint8_t* side_smem = res_smem + (staged_res ? kResBytes : 0);                         // [128 rows][128 B] staging + affine table
  float* side_tab = reinterpret_cast<float*>(side_smem + 16384);                        // [2][side_c]
  uint64_t* bars = reinterpret_cast<uint64_t*>(side_smem + (p.side ? 16384 + 2 * p.side_c * 4 : 0));
  // bars[0..NSTAGE) full, [NSTAGE..2NSTAGE) empty, [2NSTAGE] tmem_full, [2NSTAGE+1] residual tile; then tmem ptr
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 3);   // (+3 keeps bias_s 16-byte aligned)
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_smem + 2);   // [2][BLOCK_N]: bias (+temb), output-affine scale

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar_base = ptx::smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * NSTAGE);
  const uint32_t res_bar = bar_base + 8u * (2 * NSTAGE + 1);

  // tile decode
  pdl_launch_dependents();
  int tix = 0;   // timeline index of this thread's region (tuning aid: b200dm_conv_plan_set_trace)
  if (threadIdx.x == 0) trace_ev(p, 3, tix, 40);
  constexpr bool CLUSTER = CMODE == 1;
  constexpr bool KSPLIT = CMODE == 2;
  const int ks = KSPLIT ? p.ksplit : 1;
  const int krank = KSPLIT ? (int)ptx::cluster_ctaid_x() : 0;
  int bx = KSPLIT ? blockIdx.x / ks : blockIdx.x;
  const int tw = bx % p.tiles_w; bx /= p.tiles_w;
  const int th = bx % p.tiles_h; bx /= p.tiles_h;
  const int td = bx % p.tiles_d; bx /= p.tiles_d;
  const int tn = bx;
  const int w0 = tw * p.box_w, h0 = th * p.box_h, d0 = td * p.box_d, n0 = tn * p.box_n;
  const int n_tile = blockIdx.y;
  const int parity = blockIdx.z;
  const int pw = parity & 1, ph = (parity >> 1) & 1, pd = (parity >> 2) & 1;
  const int nkb_all = (p.nch0 + p.nch1) * p.ntaps;
  const int kb0 = KSPLIT ? (int)((long long)krank * nkb_all / ks) : 0;            // this CTA's K range
  const int kb1 = KSPLIT ? (int)((long long)(krank + 1) * nkb_all / ks) : nkb_all;

  // Cluster of cl_m x cl_n CTAs (x = m-tiles, y = n-tiles): the cl_n CTAs of one m-tile each fetch 1/cl_n of the A tile and
  // TMA-multicast it to all of them; the cl_m CTAs of one n-tile do the same with the B tile.  A stage may be refilled
  // only when EVERY CTA of the cluster has consumed it (their tcgen05.commit multicasts onto all empty barriers).
  // Why: one SM's TMA path sustains ~35 B/cycle (measured), a 128x128 tile wants 128 B/cycle of operands.
  const uint32_t cl_m = CLUSTER ? (uint32_t)p.cl_m : 1u, cl_n = CLUSTER ? (uint32_t)p.cl_n : 1u;
  const uint32_t rank_m = CLUSTER ? ptx::cluster_ctaid_x() : 0u, rank_n = CLUSTER ? ptx::cluster_ctaid_y() : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), cl_m * cl_n + (side_on ? 1u : 0u)); }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::mbar_init(res_bar, 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tmap(&mapA0);
    ptx::prefetch_tmap(&mapB);
    if (staged) ptx::prefetch_tmap(&om.y[blockIdx.z]);
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_smem), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CLUSTER) ptx::cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
  bool mid_synced = false;                // split-K: every thread of the cluster passes ONE mid-kernel cluster barrier
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) trace_ev(p, 3, tix, 41);
  pdl_wait();   // everything above overlapped the previous kernel's tail; from here on we touch its outputs
  if (threadIdx.x == 0) trace_ev(p, 3, tix, 42);

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      const int b_row = (p.mode == B200DM_CONV_PARITY ? parity : (p.mode == B200DM_CONV_BATCHED_GEMM ? n0 : 0)) * p.n_pad +
                        n_tile * BLOCK_N;
      uint32_t s = 0, phase = 0;
      int chunk = kb0 / p.ntaps, tap = kb0 % p.ntaps;
      if (BLOCK_N >= 64 && staged_res && krank == 0 && ptx::elect_one()) {
        // residual tile (same box as the output tile) -> smem, consumed by the epilogue long after it has landed
        int ng = 0;
        for (int g = 0; g < BLOCK_N / 64; ++g) if (n_tile * BLOCK_N + g * 64 < p.c_out) ++ng;
        ptx::mbar_expect_tx(res_bar, (uint32_t)ng * 16384u);
        for (int g = 0; g < ng; ++g)
          ptx::tma_load_5d(ptx::smem_u32(res_smem) + g * 16384, &om.r, res_bar, n_tile * BLOCK_N + g * 64, w0, h0, d0, n0);
      }
      __syncwarp();
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!ptx::mbar_wait(empty_bar(s), phase ^ 1, p.dbg, 1)) break;
        if (lane == 0 && kb == kb0) trace_ev(p, 3, tix, 43);
        if (ptx::elect_one()) {
          int ow, oh, od;
          if (p.mode == B200DM_CONV_PARITY) {
            ow = (tap & 1) - 1 + pw; oh = ((tap >> 1) & 1) - 1 + ph; od = ((tap >> 2) & 1) - 1 + pd;
          } else {
            const int k = p.ksize;
            ow = tap % k - p.pad; oh = (tap / k) % k - p.pad; od = tap / (k * k) - p.pad;
          }
          const uint32_t a_dst = smem_base + s * kStageBytes;
          const uint32_t b_dst = a_dst + kABytes;
          ptx::mbar_expect_tx(full_bar(s), kStageBytes);
          const CUtensorMap* mapA = chunk < p.nch0 ? &mapA0 : &mapA1;
          const int cch = (chunk < p.nch0 ? chunk : chunk - p.nch0) * 64;
          int cw = w0 * p.stride + ow, chh = h0 * p.stride + oh, cd = d0 * p.stride + od, cn = n0;
          if (CLUSTER && cl_n > 1) {
            const int off = (int)rank_n * p.a_split_ext;
            if (p.a_split_dim == 1) cw += off; else if (p.a_split_dim == 2) chh += off; else if (p.a_split_dim == 3) cd += off; else cn += off;
            uint16_t mask = 0;
            for (uint32_t j = 0; j < cl_n; ++j) mask |= (uint16_t)(1u << (rank_m + j * cl_m));
            ptx::tma_load_5d_mc(a_dst + rank_n * (kABytes / cl_n), mapA, full_bar(s), cch, cw, chh, cd, cn, mask);
          } else {
            ptx::tma_load_5d(a_dst, mapA, full_bar(s), cch, cw, chh, cd, cn);
          }
          if (CLUSTER && cl_m > 1) {
            uint16_t mask = 0;
            for (uint32_t i = 0; i < cl_m; ++i) mask |= (uint16_t)(1u << (i + rank_n * cl_m));
            ptx::tma_load_2d_mc(b_dst + rank_m * (kBBytes / cl_m), &mapB, full_bar(s), kb * 64, b_row + (int)rank_m * (BLOCK_N / (int)cl_m), mask);
          } else {
            ptx::tma_load_2d(b_dst, &mapB, full_bar(s), kb * 64, b_row);
          }
        }
        __syncwarp();
        if (++s == NSTAGE) { s = 0; phase ^= 1; }
        if (++tap == p.ntaps) { tap = 0; ++chunk; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    {  // whole warp walks the loop; one elected lane issues (no per-lane serialisation loops around UTCHMMA)
      constexpr uint32_t idesc = ptx::make_idesc_act(128, BLOCK_N);
      const uint64_t a_desc0 = ptx::make_smem_desc(smem_base, 16, 1024, ptx::kLayoutSw128);
      const uint64_t b_desc0 = ptx::make_smem_desc(smem_base + kABytes, 16, 1024, ptx::kLayoutSw128);
      bool ok = true;
      uint32_t s = 0, phase = 0;
      int mchunk = kb0 / p.ntaps, mtap = kb0 % p.ntaps;
      for (int kb = kb0; kb < kb1 && ok; ++kb) {
        ok = ptx::mbar_wait(full_bar(s), phase, p.dbg, 2);
        if (!ok) break;
        if (lane == 0) trace_ev(p, 0, tix, 44);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          // swap_ab (BLOCK_N = 128, both tiles [128 rows][64 K]): the weight tile is the A operand -> D^T in TMEM
          const uint64_t da = (p.swap_ab ? b_desc0 : a_desc0) + (uint64_t)(s * (kStageBytes >> 4));
          const uint64_t db = (p.swap_ab ? a_desc0 : b_desc0) + (uint64_t)(s * (kStageBytes >> 4));
          // a ragged last chunk of a K segment (e.g. C_in = 32 or 8) issues only the K=16 steps that hold real channels
          const int ksn = mchunk == p.nch0 - 1 ? p.ksteps0_last : (mchunk == p.nch0 + p.nch1 - 1 && p.nch1 > 0 ? p.ksteps1_last : 4);
          ptx::tc_mma_f16(tmem_base, da, db, idesc, kb != kb0 ? 1u : 0u);
          if (ksn > 1) ptx::tc_mma_f16(tmem_base, da + 2, db + 2, idesc, 1u);
          if (ksn > 2) ptx::tc_mma_f16(tmem_base, da + 4, db + 4, idesc, 1u);
          if (ksn > 3) ptx::tc_mma_f16(tmem_base, da + 6, db + 6, idesc, 1u);
          if (CLUSTER) ptx::tc_commit_mc(empty_bar(s), (uint16_t)((1u << (cl_m * cl_n)) - 1u));
          else ptx::tc_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == NSTAGE) { s = 0; phase ^= 1; }
        if (++mtap == p.ntaps) { mtap = 0; ++mchunk; }
      }
      if (ptx::elect_one()) ptx::tc_commit(tmem_full_bar);
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int t = r;
    const int iw = t % p.box_w; t /= p.box_w;
    const int ih = t % p.box_h; t /= p.box_h;
    const int id = t % p.box_d; t /= p.box_d;
    const int in_ = t;
    const int mw = w0 + iw, mh = h0 + ih, md = d0 + id, n = n0 + in_;
    const bool valid = (mw < p.m_w) && (mh < p.m_h) && (md < p.m_d) && (n < p.batch);
    int ow = mw, oh = mh, od = md;
    if (p.mode == B200DM_CONV_PARITY) { ow = 2 * mw + pw; oh = 2 * mh + ph; od = 2 * md + pd; }
    const int64_t vox = ((int64_t)od * p.out_h + oh) * p.out_w + ow;           // voxel index within the sample
    const int64_t vox_per = (int64_t)p.out_d * p.out_h * p.out_w;
    const int64_t row_off = ((int64_t)n * vox_per + vox) * p.c_out;             // NDHWC element offset of channel 0
    // bias (+ temb row when the tile lies inside one sample) -> smem while the MMAs run
    const float* cb = nullptr;
    const float* cbrow = nullptr;
    if (p.chan_bias) {
      const int tt = p.t_dev ? p.t_dev[0] : 0;
      const bool uniform = p.box_n == 1 || p.chan_bias_rows <= 1;
      const float* row = p.chan_bias + ((int64_t)tt * p.chan_bias_rows + (p.chan_bias_rows > 1 ? (uniform ? n0 : n) : 0)) * p.c_out;
      if (uniform) cbrow = n0 < p.batch ? row : nullptr;
      else if (valid) cb = row;
    }
    const bool has_bs = p.bias != nullptr || cbrow != nullptr || p.out_scale != nullptr;
    float* scale_s = bias_s + BLOCK_N;
    if (has_bs) {
      stage_bias(p, bias_s, scale_s, n_tile * BLOCK_N, BLOCK_N, cbrow, threadIdx.x - 64);
      epilogue_bar_sync();
    }
    if (BLOCK_N >= 64 && CMODE == 0 && side_on) {
      // ---- side output: every A tile (one 64-channel chunk of x or skip, 1^3 conv => k-block == chunk) is read back from
      // shared memory, normalised (+ swish) and TMA-stored at its channel offset of the concatenated tensor; the stage is
      // released to the producer only after these 128 threads are done with it (empty barrier count 2)
      const int et = threadIdx.x - 64;
      for (int c = et; c < p.side_c; c += 128) { side_tab[c] = __ldg(p.side_scale + c); side_tab[p.side_c + c] = __ldg(p.side_shift + c); }
      epilogue_bar_sync();
      uint32_t ss = 0, sph = 0;
      for (int kb = 0; kb < nkb_all; ++kb) {
        if (!ptx::mbar_wait(full_bar(ss), sph, p.dbg, 5)) break;
        const int cbase = kb < p.nch0 ? kb * 64 : p.nch0 * 64 + (kb - p.nch0) * 64;   // (c0 % 64 == 0 whenever c1 > 0)
        if (kb > 0) {   // the previous chunk's store has finished reading the staging tile
          if (et == 0) ptx::bulk_wait_read_all();
          epilogue_bar_sync();
        }
        const uint8_t* arow = smem + ss * kStageBytes + (uint32_t)r * 128u;
        uint8_t* srow = side_smem + (uint32_t)r * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t o = (uint32_t)((j ^ (r & 7)) << 4);
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(arow + o), f);
          const int c = cbase + j * 8;
          if (c < p.side_c) {   // (side_c % 8 == 0)
            const float4 a0 = *reinterpret_cast<const float4*>(side_tab + c), a1 = *reinterpret_cast<const float4*>(side_tab + c + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(side_tab + p.side_c + c), b1 = *reinterpret_cast<const float4*>(side_tab + p.side_c + c + 4);
            f[0] = apply_act(fmaf(f[0], a0.x, b0.x), p.side_act); f[1] = apply_act(fmaf(f[1], a0.y, b0.y), p.side_act);
            f[2] = apply_act(fmaf(f[2], a0.z, b0.z), p.side_act); f[3] = apply_act(fmaf(f[3], a0.w, b0.w), p.side_act);
            f[4] = apply_act(fmaf(f[4], a1.x, b1.x), p.side_act); f[5] = apply_act(fmaf(f[5], a1.y, b1.y), p.side_act);
            f[6] = apply_act(fmaf(f[6], a1.z, b1.z), p.side_act); f[7] = apply_act(fmaf(f[7], a1.w, b1.w), p.side_act);
          }
          *reinterpret_cast<bf16x8*>(srow + o) = pack8(f);
        }
        ptx::fence_proxy_async();
        epilogue_bar_sync();
        if (et == 0) {
          ptx::tma_store_5d(&om.s, ptx::smem_u32(side_smem), cbase, w0, h0, d0, n0);
          ptx::bulk_commit_group();
          ptx::mbar_arrive(empty_bar(ss));
        }
        if (++ss == NSTAGE) { ss = 0; sph ^= 1; }
      }
      if (et == 0) ptx::bulk_wait_read_all();
    }
    if (warp == 2 && lane == 0) trace_ev(p, 1, tix, 45);
    const bool ok = ptx::mbar_wait(tmem_full_bar, 0, p.dbg, 3);
    ptx::tc_fence_after();
    if (warp == 2 && lane == 0) trace_ev(p, 1, tix, 46);
    // split-K staging: partial accumulators as fp32 [col][row] in this CTA's (now idle) pipeline stages
    float* stage_f = reinterpret_cast<float*>(smem);
    if (KSPLIT && krank != 0) {
      if (ok) {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
          uint32_t rr[16];
          ptx::tc_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, rr);
          ptx::tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j += 4)   // [col / 4][row] float4: conflict-free 16-byte accesses on both sides
            *reinterpret_cast<float4*>(stage_f + ((c0 + j) * 128 + r * 4)) =
                make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]), __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
        }
      }
      ptx::cluster_sync_all();   // release: the leader may now read the staging through DSMEM
      mid_synced = true;
    } else {
      if (KSPLIT) { ptx::cluster_sync_all(); mid_synced = true; }   // acquire: every peer's partial tile is staged
      if (BLOCK_N == 128 && staged && p.swap_ab) {
        // ---- per-sample transposed output (V^T of the attention blocks) by operand swap: this thread's TMEM lane is
        // output channel n_tile*128 + r, the 128 columns are the tile's voxels = 256 contiguous bytes of y[n][channel][:]
        if (ok) {
          const int co = n_tile * BLOCK_N + r;
          const float bsv = (p.bias && co < p.c_out) ? __ldg(p.bias + co) : 0.f;
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
          const int vox0 = (d0 * p.m_h + h0) * p.m_w + w0;
#pragma unroll 1
          for (int g = 0; g < 2; ++g) {
            uint8_t* stg = smem + g * kStageBytes;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t ra[16], rb[16];
              ptx::tc_ld_32x32b_x16(trow + (uint32_t)(g * 64 + h * 32), ra);
              ptx::tc_ld_32x32b_x16(trow + (uint32_t)(g * 64 + h * 32 + 16), rb);
              ptx::tc_wait_ld();
              float v[32];
#pragma unroll
              for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(ra[j]) + bsv; v[16 + j] = __uint_as_float(rb[j]) + bsv; }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<bf16x8*>(stg + (uint32_t)r * 128u + (uint32_t)(((h * 4 + u) ^ (r & 7)) << 4)) =
                    pack8(*reinterpret_cast<float(*)[8]>(&v[8 * u]));
            }
            ptx::fence_proxy_async();
            epilogue_bar_sync();
            if (threadIdx.x == 64) {
              ptx::tma_store_5d(&om.y[0], ptx::smem_u32(stg), vox0 + g * 64, n_tile * BLOCK_N, 0, 0, n0);
              ptx::bulk_commit_group();
            }
          }
          if (threadIdx.x == 64) ptx::bulk_wait_read_all();
        }
      } else if (BLOCK_N >= 64 && staged) {
        // ---- staged epilogue: 64-column groups -> bf16 SWIZZLE_128B tile in the (idle) pipeline stages -> one TMA store
        // per group.  Row-per-thread global stores touch 32 lines per instruction (measured ~1000 cycles per 16-column
        // chunk); the TMA store writes whole 128-byte rows and clips the ragged edges itself.
        if (ok) {
          if (staged_res) ptx::mbar_wait(res_bar, 0, p.dbg, 4);
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
          for (int g = 0; g < BLOCK_N / 64; ++g) {
            const int colg = n_tile * BLOCK_N + g * 64;
            if (colg >= p.c_out) break;   // CTA-uniform (c_out % 64 == 0 on this path)
            uint8_t* stg = smem + g * kStageBytes;
            const uint8_t* rs = staged_res ? res_smem + g * 16384 : nullptr;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int c0 = g * 64 + h * 32;
              uint32_t ra[16], rb[16];
              ptx::tc_ld_32x32b_x16(trow + (uint32_t)c0, ra);
              ptx::tc_ld_32x32b_x16(trow + (uint32_t)c0 + 16, rb);
              ptx::tc_wait_ld();
              if (KSPLIT) {
                for (int pr = 1; pr < ks; ++pr) {
                  const uint32_t remote = ptx::mapa_shared(ptx::smem_u32(stage_f + c0 * 128 + r * 4), (uint32_t)pr);
#pragma unroll
                  for (int j = 0; j < 16; j += 4) {
                    const float4 a4 = ptx::ld_dsmem_f32x4(remote + j * 512), b4 = ptx::ld_dsmem_f32x4(remote + (16 + j) * 512);
                    ra[j] = __float_as_uint(__uint_as_float(ra[j]) + a4.x); ra[j + 1] = __float_as_uint(__uint_as_float(ra[j + 1]) + a4.y);
                    ra[j + 2] = __float_as_uint(__uint_as_float(ra[j + 2]) + a4.z); ra[j + 3] = __float_as_uint(__uint_as_float(ra[j + 3]) + a4.w);
                    rb[j] = __float_as_uint(__uint_as_float(rb[j]) + b4.x); rb[j + 1] = __float_as_uint(__uint_as_float(rb[j + 1]) + b4.y);
                    rb[j + 2] = __float_as_uint(__uint_as_float(rb[j + 2]) + b4.z); rb[j + 3] = __float_as_uint(__uint_as_float(rb[j + 3]) + b4.w);
                  }
                }
              }
              conv_epilogue16_staged(p, ra, r, h * 32, n_tile * BLOCK_N + c0, has_bs ? bias_s + c0 : nullptr, cb,
                                     p.out_scale ? scale_s + c0 : nullptr, rs, stg);
              conv_epilogue16_staged(p, rb, r, h * 32 + 16, n_tile * BLOCK_N + c0 + 16, has_bs ? bias_s + c0 + 16 : nullptr, cb,
                                     p.out_scale ? scale_s + c0 + 16 : nullptr, rs, stg);
            }
            ptx::fence_proxy_async();
            epilogue_bar_sync();
            if (threadIdx.x == 64) {
              ptx::tma_store_5d(&om.y[parity], ptx::smem_u32(stg), colg, w0, h0, d0, n0);
              ptx::bulk_commit_group();
            }
          }
          if (threadIdx.x == 64) ptx::bulk_wait_read_all();
        }
      } else if (ok) {
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
          const int col0 = n_tile * BLOCK_N + c0;
          if (col0 >= p.c_out) break;  // warp-uniform
          uint32_t rr[16];
          if (p.epi_dbg < 2) {
            ptx::tc_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, rr);
            ptx::tc_wait_ld();
          }
          if (p.epi_dbg != 0) continue;
          if (KSPLIT) {
            for (int pr = 1; pr < ks; ++pr) {
              const uint32_t remote = ptx::mapa_shared(ptx::smem_u32(stage_f + c0 * 128 + r * 4), (uint32_t)pr);
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 a4 = ptx::ld_dsmem_f32x4(remote + j * 512);
                rr[j] = __float_as_uint(__uint_as_float(rr[j]) + a4.x); rr[j + 1] = __float_as_uint(__uint_as_float(rr[j + 1]) + a4.y);
                rr[j + 2] = __float_as_uint(__uint_as_float(rr[j + 2]) + a4.z); rr[j + 3] = __float_as_uint(__uint_as_float(rr[j + 3]) + a4.w);
              }
            }
          }
          if (!valid) continue;
          conv_epilogue16(p, rr, col0, n, vox, vox_per, row_off, has_bs ? bias_s + c0 : nullptr, cb,
                          p.out_scale ? scale_s + c0 : nullptr);
        }
      }
    }
  }
  if (warp == 2 && lane == 0) trace_ev(p, 1, tix, 47);
  if (KSPLIT && !mid_synced) ptx::cluster_sync_all();   // producer / MMA warps: their half of the mid-kernel barrier
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
  if (CLUSTER || KSPLIT) ptx::cluster_sync_all();   // no CTA exits while a peer may still write / read its shared memory
  if (threadIdx.x == 0) trace_ev(p, 3, tix, 48);
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

int pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }

struct Geometry {
  int ntaps, pad, groups, nch0, nch1, n_pad, block_n;
  int out_d, out_h, out_w, m_d, m_h, m_w;
  int box_w, box_h, box_d, box_n;
  size_t ktot;
};

int compute_geometry(const b200dm_conv_desc* d, Geometry* g) {
  B2_CHECK_ARG(d, "conv: null desc");
  B2_CHECK_ARG(d->mode >= 0 && d->mode <= 2, "conv: bad mode %d", d->mode);
  B2_CHECK_ARG(d->batch > 0 && d->in_d > 0 && d->in_h > 0 && d->in_w > 0, "conv: empty input");
  B2_CHECK_ARG(d->c0 > 0 && d->c0 % 8 == 0 && d->c1 >= 0 && d->c1 % 8 == 0, "conv: C_in segments must be multiples of 8 (got %d,%d)", d->c0, d->c1);
  B2_CHECK_ARG(d->c_out > 0, "conv: c_out must be > 0");
  g->nch0 = (d->c0 + 63) / 64;
  g->nch1 = (d->c1 + 63) / 64;
  if (d->mode == B200DM_CONV_DIRECT) {
    B2_CHECK_ARG(d->ksize == 1 || d->ksize == 3 || d->ksize == 4, "conv: ksize %d unsupported", d->ksize);
    B2_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv: stride %d unsupported", d->stride);
    g->ntaps = d->ksize * d->ksize * d->ksize;
    g->groups = 1;
    g->out_d = (d->in_d + d->stride - 1) / d->stride;
    g->out_h = (d->in_h + d->stride - 1) / d->stride;
    g->out_w = (d->in_w + d->stride - 1) / d->stride;
    // TF 'same': total = max((out-1)*s + k - in, 0), before = total/2.  Cubic volumes: same pad on all axes.
    B2_CHECK_ARG(d->in_d == d->in_h || d->stride == 1, "conv: strided conv needs equal D,H,W");
    int total = (g->out_w - 1) * d->stride + d->ksize - d->in_w;
    if (total < 0) total = 0;
    g->pad = total / 2;
    if (d->stride == 2) {
      int td = (g->out_d - 1) * 2 + d->ksize - d->in_d, th = (g->out_h - 1) * 2 + d->ksize - d->in_h;
      B2_CHECK_ARG((td < 0 ? 0 : td) / 2 == g->pad && (th < 0 ? 0 : th) / 2 == g->pad, "conv: strided conv needs the same 'same' padding on all axes");
    }
    g->m_d = g->out_d; g->m_h = g->out_h; g->m_w = g->out_w;
  } else if (d->mode == B200DM_CONV_PARITY) {
    B2_CHECK_ARG(d->ksize == 3 || d->ksize == 4, "conv: parity mode needs ksize 3 (upsample fold) or 4 (transposed conv)");
    g->ntaps = 8; g->groups = 8; g->pad = 0;
    g->out_d = 2 * d->in_d; g->out_h = 2 * d->in_h; g->out_w = 2 * d->in_w;
    g->m_d = d->in_d; g->m_h = d->in_h; g->m_w = d->in_w;
  } else {
    B2_CHECK_ARG(d->in_d == 1 && d->in_h == 1, "conv: batched GEMM uses in_d = in_h = 1, in_w = rows");
    B2_CHECK_ARG(d->c1 == 0, "conv: batched GEMM takes one K segment");
    g->ntaps = 1; g->groups = d->batch; g->pad = 0;
    g->out_d = 1; g->out_h = 1; g->out_w = d->in_w;
    g->m_d = 1; g->m_h = 1; g->m_w = d->in_w;
  }
  if (d->c_out <= 16) g->n_pad = 16;
  else if (d->c_out <= 32) g->n_pad = 32;
  else if (d->c_out <= 64) g->n_pad = 64;
  else g->n_pad = ((d->c_out + 127) / 128) * 128;
  g->block_n = g->n_pad < 128 ? g->n_pad : 128;
  if (const char* e = tuning_env("B200DM_IGEMM_BN")) {   // tuning aid: narrower N tiles (more CTAs, shorter epilogues)
    const int v = atoi(e);
    if ((v == 64 || v == 32) && g->n_pad >= 128 && !(d->mode == B200DM_CONV_DIRECT && d->ksize == 3 && d->stride == 1 && d->in_w >= 8 && d->in_h >= 16))
      g->block_n = v;
  }
  g->ktot = (size_t)(g->nch0 + g->nch1) * g->ntaps * 64;
  // M tile box, product 128
  if (d->mode == B200DM_CONV_BATCHED_GEMM) {
    g->box_w = 128; g->box_h = 1; g->box_d = 1; g->box_n = 1;
  } else {
    int bw = pow2_ceil(g->m_w); if (bw > 8) bw = 8;
    int bh = pow2_ceil(g->m_h); if (bh > 128 / bw) bh = 128 / bw; if (bh > 16) bh = 16;
    int bd = pow2_ceil(g->m_d); if (bd > 128 / (bw * bh)) bd = 128 / (bw * bh);
    int bn = 128 / (bw * bh * bd);
    g->box_w = bw; g->box_h = bh; g->box_d = bd; g->box_n = bn;
  }
  return B200DM_OK;
}

uint16_t f32_to_bf16_rne(float f) {   // fp32 -> the library's 16-bit storage type (host side), round to nearest even
#ifdef B200DM_ACT_FP16
  const __half h = __float2half_rn(f);
  uint16_t r;
  memcpy(&r, &h, 2);
  return r;
#endif
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace

struct b200dm_conv_plan {
  b200dm_conv_desc desc;
  Geometry g;
  ConvParams p;
  CUtensorMap mapA0, mapA1, mapB;
  ConvOutMaps om;
  dim3 grid;
  size_t smem;
  int nstage;
  double flops;
  bool halo = false;
  bool ups = false;    // x2 up-convolution (PARITY mode) on the halo / CTA-pair machinery (conv_halo_up.cuh)
  bool cg2 = false;    // halo kernel, BLOCK_N = 64, two planes, staged: CTA pairs (cta_group::2), see conv_halo.cuh
  bool wide = false;   // halo kernel, BLOCK_N = 128, staged: 3-tap weight stages x 2, 4 slabs (see conv_plan_create)
  bool pair = false;   // halo kernel on 8 x 8 planes: 8w x 8h x 2d tiles from pair slabs (conv_halo.cuh)
  int halo_td = 1, halo_nb = 4, halo_tps = 1;
  int halo_ns = 0;     // slab ring depth when a variant fixes it (0 = by epilogue kind)
  const CUtensorMap& residual_map() const { return p.residual ? om.r : om.y[0]; }   // (a valid map either way: unused without a residual)
  bool fuse = false;   // CTA-pair N = 128 fp32 conv with the reverse-diffusion update fused into its epilogue (set_fused_update)
  bool stencil = false;   // C_out = 1, C_in = 32 3^3 conv: HBM-bound stencil-reduce kernel (conv_stencil.cuh)
  stencil::Params sp;
  CUtensorMap mapS;
  bool sweep = false;     // C_in = C_out = 32 3^3 conv: d-sweeping N = 96 kernel (conv_sweep.cuh); mapS = input, mapB = weights, om.y[0] = output
  sweep::Params wp;
};

static int* g_dbg_flag = nullptr;
static int ensure_dbg_flag() {
  if (!g_dbg_flag) {
    B2_CHECK_CUDA(cudaMalloc(&g_dbg_flag, sizeof(int)));
    B2_CHECK_CUDA(cudaMemset(g_dbg_flag, 0, sizeof(int)));
  }
  return B200DM_OK;
}

int* b200dm_dbg_flag_ptr() { return ensure_dbg_flag() == B200DM_OK ? g_dbg_flag : nullptr; }

extern "C" int b200dm_debug_flag_read_reset(int32_t* flag_out) {
  B2_CHECK_ARG(flag_out, "debug_flag: null");
  *flag_out = 0;
  if (!g_dbg_flag) return B200DM_OK;
  B2_CHECK_CUDA(cudaMemcpy(flag_out, g_dbg_flag, sizeof(int), cudaMemcpyDeviceToHost));
  B2_CHECK_CUDA(cudaMemset(g_dbg_flag, 0, sizeof(int)));
  return B200DM_OK;
}

extern "C" size_t b200dm_conv_packed_weight_bytes(const b200dm_conv_desc* d) {
  Geometry g;
  if (compute_geometry(d, &g) != B200DM_OK) return 0;
  if (d->mode == B200DM_CONV_BATCHED_GEMM) return 0;
  return (size_t)g.groups * g.n_pad * g.ktot * 2;
}

extern "C" int b200dm_conv_pack_weights(const b200dm_conv_desc* d, const float* w, int32_t transposed, void* packed) {
  Geometry g;
  int rc = compute_geometry(d, &g);
  if (rc) return rc;
  B2_CHECK_ARG(w && packed, "conv_pack_weights: null pointer");
  B2_CHECK_ARG(d->mode != B200DM_CONV_BATCHED_GEMM, "conv_pack_weights: batched GEMM has no packed weights");
  const int k = d->ksize, cin = d->c0 + d->c1, cout = d->c_out;
  uint16_t* out = reinterpret_cast<uint16_t*>(packed);
  memset(out, 0, (size_t)g.groups * g.n_pad * g.ktot * 2);
  // Keras kernel index: ((kd*k + kh)*k + kw) * cin*cout + (transposed ? co*cin + ci : ci*cout + co)
  auto W = [&](int kd, int kh, int kw, int ci, int co) -> float {
    const size_t base = ((size_t)(kd * k + kh) * k + kw) * (size_t)cin * cout;
    return transposed ? w[base + (size_t)co * cin + ci] : w[base + (size_t)ci * cout + co];
  };
  const int nch = g.nch0 + g.nch1;
  for (int grp = 0; grp < g.groups; ++grp) {
    const int pw = grp & 1, ph = (grp >> 1) & 1, pd = (grp >> 2) & 1;
    for (int co = 0; co < cout; ++co) {
      uint16_t* row = out + ((size_t)grp * g.n_pad + co) * g.ktot;
      for (int ch = 0; ch < nch; ++ch) {
        for (int tap = 0; tap < g.ntaps; ++tap) {
          for (int c = 0; c < 64; ++c) {
            int ci;
            if (ch < g.nch0) { ci = ch * 64 + c; if (ci >= d->c0) continue; }
            else { ci = (ch - g.nch0) * 64 + c; if (ci >= d->c1) continue; ci += d->c0; }
            float v = 0.f;
            if (d->mode == B200DM_CONV_DIRECT) {
              v = W(tap / (k * k), (tap / k) % k, tap % k, ci, co);
            } else {
              const int tw = tap & 1, th = (tap >> 1) & 1, tdp = (tap >> 2) & 1;
              if (k == 3) {
                // nearest-upsample fold: parity 0: slot0 (offset -1) <- {k0}, slot1 (offset 0) <- {k1,k2};
                //                        parity 1: slot0 (offset 0) <- {k0,k1}, slot1 (offset +1) <- {k2}
                auto lo = [](int par, int slot) { return par == 0 ? (slot == 0 ? 0 : 1) : (slot == 0 ? 0 : 2); };
                auto hi = [](int par, int slot) { return par == 0 ? (slot == 0 ? 0 : 2) : (slot == 0 ? 1 : 2); };
                double acc = 0.0;
                for (int kd = lo(pd, tdp); kd <= hi(pd, tdp); ++kd)
                  for (int kh = lo(ph, th); kh <= hi(ph, th); ++kh)
                    for (int kw = lo(pw, tw); kw <= hi(pw, tw); ++kw) acc += (double)W(kd, kh, kw, ci, co);
                v = (float)acc;
              } else {
                // transposed conv k4 s2 'same': out[2m+p] = sum_off in[m+off] * w[p + 1 - 2*off], off = slot - 1 + p
                auto kk = [](int par, int slot) { return par + 1 - 2 * (slot - 1 + par); };
                v = W(kk(pd, tdp), kk(ph, th), kk(pw, tw), ci, co);
              }
            }
            row[((size_t)ch * g.ntaps + tap) * 64 + c] = f32_to_bf16_rne(v);
          }
        }
      }
    }
  }
  return B200DM_OK;
}

static size_t conv_smem_bytes(int block_n, int nstage, bool staged_res, int side_c = 0) {
  return 1024 + (size_t)nstage * (kABytes + block_n * 128) + (staged_res ? (size_t)block_n * 256 : 0) +
         (side_c > 0 ? 16384 + (size_t)2 * side_c * 4 : 0) + (2 * nstage + 3) * 8 + 16 + 2 * block_n * 4;
}

static size_t conv_smem_attr(int block_n, int nstage) {   // opt-in ceiling of one kernel instantiation
  const size_t v = conv_smem_bytes(block_n, nstage, true, 1024);
  return v > 232448 ? 232448 : v;
}

template <int BLOCK_N, int NSTAGE>
static int launch_conv(const b200dm_conv_plan* pl, cudaStream_t s) {
  if (pl->p.ksplit > 1) {
    auto kern = conv_igemm_kernel<BLOCK_N, NSTAGE, 2>;
    static bool attr_set = false;
    if (!attr_set) {
      B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_attr(BLOCK_N, NSTAGE)));
      attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = pl->grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl->p.ksplit; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = b200dm_pdl_enabled() ? 2 : 1;
    B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA0, pl->mapA1, pl->mapB, pl->om, pl->p));
    return B200DM_OK;
  }
  if (pl->p.cl_m * pl->p.cl_n > 1) {
    auto kern = conv_igemm_kernel<BLOCK_N, NSTAGE, 1>;
    static bool attr_set = false;
    if (!attr_set) {
      B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_attr(BLOCK_N, NSTAGE)));
      attr_set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = pl->grid; cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl->p.cl_m; at[0].val.clusterDim.y = pl->p.cl_n; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = b200dm_pdl_enabled() ? 2 : 1;
    B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA0, pl->mapA1, pl->mapB, pl->om, pl->p));
    return B200DM_OK;
  }
  auto kern = conv_igemm_kernel<BLOCK_N, NSTAGE, 0>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem_attr(BLOCK_N, NSTAGE)));
    attr_set = true;
  }
  B2_CHECK_CUDA(b2_launch(kern, pl->grid, dim3(kThreads), pl->smem, s, pl->mapA0, pl->mapA1, pl->mapB, pl->om, pl->p));
  return B200DM_OK;
}

constexpr int kHaloNS = 6;         // slab ring depth, direct-store epilogue
constexpr int kHaloNSStaged = 5;   // staged epilogue: one slab less, the room holds the output staging tiles
// N <= 32 tiles have small weight stages and 16 KB of staging: room for the 6th slab.  With 5 (4 live per 2-plane tile) the
// issuers waited ~1000 cycles per 8000-cycle tile for the first two slabs of the next tile (tools/conv_trace.py, 128^3 32 -> 32).
constexpr int halo_ns_for(int block_n, bool staged) { return staged && block_n > 32 ? kHaloNSStaged : kHaloNS; }

template <int BLOCK_N, int TD, int NB, int TPS, bool STAGED>
static int launch_halo(const b200dm_conv_plan* pl, cudaStream_t s) {
  auto kern = halo::conv_halo_kernel<BLOCK_N, TD, halo_ns_for(BLOCK_N, STAGED), NB, TPS, STAGED>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  B2_CHECK_CUDA(b2_launch(kern, pl->grid, dim3(halo::kThreads), pl->smem, s, pl->mapA0, pl->mapA1, pl->mapB, pl->om.y[0], pl->om.y[1], pl->p));
  return B200DM_OK;
}

template <int BLOCK_N, int NS, int NB, bool PAIR = false>
static int launch_halo_up(const b200dm_conv_plan* pl, cudaStream_t s) {
  auto kern = halo::conv_halo_up_kernel<BLOCK_N, NS, NB, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = pl->grid; cfg.blockDim = dim3(halo::kThreads); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = b200dm_pdl_enabled() ? 2 : 1;
  B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA0, pl->mapA1, pl->mapB, pl->om, pl->p));
  return B200DM_OK;
}

constexpr int kHaloNSPair = 4;   // pair slabs are 25 KB; three per channel chunk are live
constexpr int kHaloNSWide = 4;

template <int BLOCK_N, int TD, int NS, int NB, bool PAIR, bool STAGED = true>
static int launch_halo_cg2(const b200dm_conv_plan* pl, cudaStream_t s) {
  auto kern = halo::conv_halo_kernel<BLOCK_N, TD, NS, NB, 3, STAGED, PAIR, true>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = pl->grid; cfg.blockDim = dim3(halo::kThreads); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = b200dm_pdl_enabled() ? 2 : 1;
  B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA0, pl->mapA1, pl->mapB, pl->om.y[0], pl->om.y[1], pl->p));
  return B200DM_OK;
}

constexpr int kHaloNSFused = 4;   // the 16-bit staging tiles of the fused update take the room of the fifth slab

template <int TD>
static int launch_halo_cg2_fused(const b200dm_conv_plan* pl, cudaStream_t s) {
  auto kern = halo::conv_halo_kernel<128, TD, kHaloNSFused, 3, 3, true, false, true, true>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = pl->grid; cfg.blockDim = dim3(halo::kThreadsFused); cfg.dynamicSmemBytes = pl->smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = b200dm_pdl_enabled() ? 2 : 1;
  B2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, pl->mapA0, pl->mapA1, pl->mapB, pl->om.y[0], pl->om.y[1], pl->p));
  return B200DM_OK;
}

template <int TD>
static int launch_halo_wide(const b200dm_conv_plan* pl, cudaStream_t s) {
  auto kern = halo::conv_halo_kernel<128, TD, kHaloNSWide, 2, 3, true>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  B2_CHECK_CUDA(b2_launch(kern, pl->grid, dim3(halo::kThreads), pl->smem, s, pl->mapA0, pl->mapA1, pl->mapB, pl->om.y[0], pl->om.y[1], pl->p));
  return B200DM_OK;
}

static int launch_halo_pair(const b200dm_conv_plan* pl, cudaStream_t s) {
  auto kern = halo::conv_halo_kernel<64, 1, kHaloNSPair, 3, 3, true, true>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    attr_set = true;
  }
  B2_CHECK_CUDA(b2_launch(kern, pl->grid, dim3(halo::kThreads), pl->smem, s, pl->mapA0, pl->mapA1, pl->mapB, pl->om.y[0], pl->om.y[1], pl->p));
  return B200DM_OK;
}

// weight-ring depth / taps per stage by BLOCK_N (smem: 6 slabs x 23 KB + NB x TPS x BLOCK_N x 128 B <= 227 KB)
static void halo_ring_config(int block_n, int* nb, int* tps) {
  if (block_n >= 128) { *nb = 4; *tps = 1; }
  else if (block_n == 64) { *nb = 3; *tps = 3; }
  else { *nb = 4; *tps = 3; }
}

template <int BLOCK_N, int NB, int TPS>
static int dispatch_halo(const b200dm_conv_plan* pl, cudaStream_t s) {
  if constexpr (BLOCK_N >= 32) {
    if (pl->p.tma_epi) return pl->halo_td == 2 ? launch_halo<BLOCK_N, 2, NB, TPS, true>(pl, s) : launch_halo<BLOCK_N, 1, NB, TPS, true>(pl, s);
  }
  return pl->halo_td == 2 ? launch_halo<BLOCK_N, 2, NB, TPS, false>(pl, s) : launch_halo<BLOCK_N, 1, NB, TPS, false>(pl, s);
}

static size_t halo_smem_bytes(const b200dm_conv_plan* pl) {
  const bool st = pl->p.tma_epi != 0;
  const int ns = pl->fuse ? kHaloNSFused : pl->halo_ns ? pl->halo_ns : (pl->pair ? kHaloNSPair : (pl->wide ? kHaloNSWide : halo_ns_for(pl->g.block_n, st)));   // (cg2: 5, or 4 for pair slabs)
  return 1024 + (size_t)ns * (pl->pair ? 25 * 1024 : halo::kSlabBytes) +
         (size_t)pl->halo_nb * pl->halo_tps * (pl->cg2 ? pl->g.block_n / 2 : pl->g.block_n) * 128 +
         (size_t)halo::stage_bytes(pl->g.block_n, st) + (pl->fuse ? halo::kFuseStage16Bytes + 32 : 0) + (2 * ns + 2 * pl->halo_nb + 4) * 8 + 16 + 4 * pl->g.block_n * 4;
}

extern "C" int b200dm_conv_plan_create(const b200dm_conv_desc* d, const void* x0, const void* x1, const void* w_packed,
                                       const float* bias, const float* chan_bias, const int32_t* t_dev,
                                       const void* residual, const void* prelu_alpha, void* y,
                                       b200dm_conv_plan** out) {
  B2_CHECK_ARG(out, "conv_plan_create: null out");
  *out = nullptr;
  Geometry g;
  int rc = compute_geometry(d, &g);
  if (rc) return rc;
  B2_CHECK_ARG(x0 && w_packed && y, "conv_plan_create: null tensor pointer");
  B2_CHECK_ARG((d->c1 > 0) == (x1 != nullptr), "conv_plan_create: x1 must be given iff c1 > 0");
  B2_CHECK_ARG(d->y_dtype == B200DM_BF16 || d->y_dtype == B200DM_F32, "conv_plan_create: bad y dtype");
  B2_CHECK_ARG(((uintptr_t)x0 & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0, "conv_plan_create: pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { b200dm_set_error("conv_plan_create: cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return B200DM_ERR_CUDA; }
  rc = ensure_dbg_flag();
  if (rc) return rc;

  b200dm_conv_plan* pl = new (std::nothrow) b200dm_conv_plan();
  B2_CHECK_ARG(pl, "conv_plan_create: out of memory");
  pl->desc = *d;
  pl->g = g;
  const int st = d->mode == B200DM_CONV_DIRECT ? d->stride : 1;
  // C_out = 1, C_in = 32, 3^3 stride 1 (the vqgan_attn_cp decoder's head): HBM-bound stencil-reduce kernel instead of an
  // N = 16 tensor-core tile (conv_stencil.cuh); bias + activation epilogue only
  if (d->mode == B200DM_CONV_DIRECT && d->ksize == 3 && d->stride == 1 && d->c_out == 1 && d->c0 == stencil::kC && d->c1 == 0 &&
      !chan_bias && !residual && !prelu_alpha && d->reserved[0] == B200DM_ACT_NONE && d->reserved[1] == 0 && d->use_halo >= 0 &&
      !tuning_env("B200DM_NO_STENCIL")) {
    cuuint64_t dims[5] = {(cuuint64_t)d->c0, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h, (cuuint64_t)d->in_d, (cuuint64_t)d->batch};
    cuuint64_t strides[4] = {(cuuint64_t)d->c0 * 2, (cuuint64_t)d->in_w * d->c0 * 2, (cuuint64_t)d->in_h * d->in_w * d->c0 * 2,
                             (cuuint64_t)d->in_d * d->in_h * d->in_w * d->c0 * 2};
    cuuint32_t box[5] = {(cuuint32_t)stencil::kC, (cuuint32_t)stencil::kHW, (cuuint32_t)stencil::kHH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (enc(&pl->mapS, kTmapAct16, 5, const_cast<void*>(x0), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      delete pl; b200dm_set_error("cuTensorMapEncodeTiled(stencil X) failed"); return B200DM_ERR_CUDA;
    }
    stencil::Params& sp = pl->sp;
    memset(&sp, 0, sizeof(sp));
    sp.batch = d->batch; sp.D = d->in_d; sp.H = d->in_h; sp.W = d->in_w;
    sp.tiles_h = (d->in_h + stencil::kTH - 1) / stencil::kTH; sp.tiles_w = (d->in_w + stencil::kTW - 1) / stencil::kTW;
    // d ranges: enough work items for ~6 per resident CTA (two CTAs per SM), each range >= 8 planes (2 halo planes per range)
    const long long cols = (long long)d->batch * sp.tiles_h * sp.tiles_w, ctas = 2LL * b2_num_sms();
    int dsplit = 1;
    while (cols * dsplit < 6 * ctas && d->in_d / (dsplit * 2) >= 8) dsplit *= 2;
    sp.dlen = (d->in_d + dsplit - 1) / dsplit; sp.dsplit = (d->in_d + sp.dlen - 1) / sp.dlen;
    sp.items = (int)(cols * sp.dsplit);
    sp.w = (const act_t*)w_packed; sp.bias = bias; sp.act = d->act; sp.y_f32 = d->y_dtype == B200DM_F32; sp.y = y; sp.dbg = g_dbg_flag;
    pl->stencil = true;
    pl->grid = dim3((unsigned)(sp.items < ctas ? sp.items : ctas), 1, 1);
    pl->smem = stencil::kSmem;
    pl->flops = 2.0 * 27 * d->c0 * (double)d->batch * d->in_d * d->in_h * d->in_w;
    *out = pl;
    return B200DM_OK;
  }
  // C_in = C_out = 32, 3^3 stride 1, 16-bit output (the decoders' 128^3 / 64^3 residual units): d-sweeping kernel with the three
  // kd taps of an input plane as one N = 96 MMA and the weight set resident in shared memory (conv_sweep.cuh)
  if (d->mode == B200DM_CONV_DIRECT && d->ksize == 3 && d->stride == 1 && d->c_out == 32 && d->c0 == 32 && d->c1 == 0 &&
      d->y_dtype == B200DM_BF16 && !chan_bias && !prelu_alpha && d->reserved[1] == 0 && d->use_halo >= 0 && g.n_pad == 32 &&
      g.ktot == 27 * 64 && d->in_w >= 8 && d->in_h >= 16 && d->in_d >= 2 && !tuning_env("B200DM_NO_SWEEP")) {
    const cuuint64_t C = 32;
    cuuint64_t dims[5] = {C, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h, (cuuint64_t)d->in_d, (cuuint64_t)d->batch};
    cuuint64_t strides[4] = {C * 2, (cuuint64_t)d->in_w * C * 2, (cuuint64_t)d->in_h * d->in_w * C * 2,
                             (cuuint64_t)d->in_d * d->in_h * d->in_w * C * 2};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    cuuint32_t boxx[5] = {32, 10, 18, 2, 1}, boxy[5] = {32, 8, 16, 1, 1};
    cuuint64_t wdims[3] = {64, 32, 27};
    cuuint64_t wstrides[2] = {(cuuint64_t)g.ktot * 2, 128};
    cuuint32_t wbox[3] = {32, 32, 1};
    const bool okm =
        enc(&pl->mapS, kTmapAct16, 5, const_cast<void*>(x0), dims, strides, boxx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
        enc(&pl->om.y[0], kTmapAct16, 5, y, dims, strides, boxy, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
        enc(&pl->mapB, kTmapAct16, 3, const_cast<void*>(w_packed), wdims, wstrides, wbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    const bool okr = !residual || enc(&pl->om.r, kTmapAct16, 5, const_cast<void*>(residual), dims, strides, boxy, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    if (!okm || !okr) { delete pl; b200dm_set_error("cuTensorMapEncodeTiled(sweep) failed"); return B200DM_ERR_CUDA; }
    sweep::Params& wp = pl->wp;
    memset(&wp, 0, sizeof(wp));
    wp.D = d->in_d; wp.H = d->in_h; wp.W = d->in_w; wp.batch = d->batch;
    wp.tiles_w = (d->in_w + 7) / 8; wp.tiles_h = (d->in_h + 15) / 16;
    const long long cols = (long long)d->batch * wp.tiles_w * wp.tiles_h, ctas = b2_num_sms();
    int dsplit = 1;
    while (cols * dsplit < 6 * ctas && d->in_d / (dsplit * 2) >= 8) dsplit *= 2;   // >= ~6 items per CTA, d ranges of >= 8 planes
    wp.dlen = (d->in_d + dsplit - 1) / dsplit; wp.dsplit = (d->in_d + wp.dlen - 1) / wp.dlen;
    wp.items = (int)(cols * wp.dsplit);
    ConvParams& p = pl->p;
    memset(&p, 0, sizeof(p));
    p.batch = d->batch; p.in_d = p.out_d = d->in_d; p.in_h = p.out_h = d->in_h; p.in_w = p.out_w = d->in_w;
    p.c_out = 32; p.n_pad = 32; p.act = d->act; p.post_act = d->reserved[0];
    p.bias = bias; p.residual = (const act_t*)residual; p.y = y; p.dbg = g_dbg_flag; p.tma_epi = 1;
    pl->sweep = true; pl->halo = true;
    pl->grid = dim3((unsigned)(wp.items < ctas ? wp.items : ctas), 1, 1);
    pl->smem = sweep::kSmem;
    pl->flops = 2.0 * 27 * 32 * 32 * (double)d->batch * d->in_d * d->in_h * d->in_w;
    *out = pl;
    return B200DM_OK;
  }
  // halo-reuse kernel: 3^3 stride-1 convs on volumes that fill its 8w x 16h tile (use_halo = -1 forces it off)
  pl->halo = d->mode == B200DM_CONV_DIRECT && d->ksize == 3 && d->stride == 1 && d->in_w >= 8 && d->in_h >= 16 &&
             d->reserved[1] == 0 && d->use_halo >= 0;
  // 8 x 8 planes (the U-Net's deepest level): pair-slab tiles, 64 output channels per CTA so that >= 128 CTAs share the
  // weight stream; bf16 staged output only
  pl->pair = !pl->halo && d->mode == B200DM_CONV_DIRECT && d->ksize == 3 && d->stride == 1 && d->in_w == 8 && d->in_h == 8 &&
             d->in_d >= 2 && d->reserved[1] == 0 && d->use_halo >= 0 && d->y_dtype == B200DM_BF16 && d->c_out % 64 == 0 &&
             !prelu_alpha && !(tuning_env("B200DM_TMA_EPI") && atoi(tuning_env("B200DM_TMA_EPI")) == 0) && !tuning_env("B200DM_NO_PAIR");
  if (pl->pair) { pl->halo = true; g.block_n = 64; pl->g.block_n = 64; }
  // x2 up-convolutions (nearest-upsample + conv3, ConvT k4 s2) on low-resolution planes that fill the 8 x 16 halo tile
  {
    const long long per_up = (long long)((d->in_w + 7) / 8) * ((d->in_h + 15) / 16) * ((d->in_d + 1) / 2) * d->batch;
    // (C_out = 32 -- the decoders' last ConvT, 64 -> 32 at 64^3 -> 128^3 -- runs N = 32 tiles with 64-byte staged rows)
    const bool up32 = d->c_out == 32 && g.block_n == 32 && !tuning_env("B200DM_NO_UP32");
    pl->ups = d->mode == B200DM_CONV_PARITY && d->in_w >= 8 && d->in_h >= 16 && d->in_d >= 2 && (d->c_out % 64 == 0 || up32) &&
              d->y_dtype == B200DM_BF16 && !residual && !prelu_alpha && d->reserved[1] == 0 && d->use_halo >= 0 && per_up % 2 == 0 &&
              (g.block_n == 64 || g.block_n == 128 || up32) && !(tuning_env("B200DM_TMA_EPI") && atoi(tuning_env("B200DM_TMA_EPI")) == 0) &&
              !(tuning_env("B200DM_CG2") && atoi(tuning_env("B200DM_CG2")) == 0) && !tuning_env("B200DM_NO_UPS");
    if (pl->ups) pl->halo = true;
    // the same on 8 x 8 low-resolution planes (8^3 -> 16^3): pair-slab tiles
    const long long per_pair = (long long)((d->in_d + 1) / 2) * d->batch;
    if (!pl->ups && d->mode == B200DM_CONV_PARITY && d->in_w == 8 && d->in_h == 8 && d->in_d >= 2 && d->in_d % 2 == 0 && d->c_out % 64 == 0 &&
        d->y_dtype == B200DM_BF16 && !residual && !prelu_alpha && d->reserved[1] == 0 && d->use_halo >= 0 && per_pair % 2 == 0 &&
        (g.block_n == 64 || g.block_n == 128) && !(tuning_env("B200DM_TMA_EPI") && atoi(tuning_env("B200DM_TMA_EPI")) == 0) &&
        !(tuning_env("B200DM_CG2") && atoi(tuning_env("B200DM_CG2")) == 0) && !tuning_env("B200DM_NO_UPS")) {
      pl->ups = true; pl->pair = true; pl->halo = true;
    }
  }
  // igemm cluster (see the kernel): cl_n n-tiles share A, cl_m m-tiles share B; B200DM_CLUSTER="m,n" switches it on.
  int cl_m = 1, cl_n = 1, a_split_dim = 0, a_split_ext = 0;
  if (!pl->halo && st == 1) {
    const long long mt = (long long)((g.m_w + g.box_w - 1) / g.box_w) * ((g.m_h + g.box_h - 1) / g.box_h) *
                         ((g.m_d + g.box_d - 1) / g.box_d) * ((d->batch + g.box_n - 1) / g.box_n);
    const int nt = d->mode == B200DM_CONV_BATCHED_GEMM ? (d->c_out + g.block_n - 1) / g.block_n : g.n_pad / g.block_n;
    // Default OFF: measured on B200 (8^3 256->256 conv, 47 us unclustered) 2x1 68 us, 1x2 72 us, 2x2 91 us, 4x2 97 us --
    // the operand stream is capped by what ONE SM can take from L2 (~40 B/clk), which multicast does not raise at
    // cluster sizes <= 4, and the cluster-wide stage hand-shake adds latency.  Kept for experiments and tests.
    int want_m = 1, want_n = 1;
    if (const char* e = tuning_env("B200DM_CLUSTER")) { if (sscanf(e, "%d,%d", &want_m, &want_n) != 2) { want_m = 1; want_n = 1; } }
    // B is shared by m-tiles of the same sample only in GEMM mode (x = tile within sample fastest)
    const long long m_share = d->mode == B200DM_CONV_BATCHED_GEMM ? (g.m_w + g.box_w - 1) / g.box_w : mt;
    for (cl_m = want_m; cl_m > 1 && (m_share % cl_m != 0 || mt % cl_m != 0 || g.block_n / cl_m < 8); cl_m >>= 1) {}
    for (cl_n = want_n; cl_n > 1 && nt % cl_n != 0; cl_n >>= 1) {}
    if (cl_m < 1) cl_m = 1;
    if (cl_n < 1) cl_n = 1;
    if (cl_n > 1) {  // split the outermost non-unit box dim of the A tile between the sharers
      int ext[5] = {0, g.box_w, g.box_h, g.box_d, g.box_n};
      for (int dim = 4; dim >= 1; --dim)
        if (ext[dim] > 1) { if (ext[dim] % cl_n == 0) { a_split_dim = dim; a_split_ext = ext[dim] / cl_n; } break; }
      if (a_split_dim == 0) cl_n = 1;
    }
  }
  auto encodeA = [&](CUtensorMap* m, const void* ptr, int C) -> int {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h, (cuuint64_t)d->in_d, (cuuint64_t)d->batch};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)d->in_w * C * 2, (cuuint64_t)d->in_h * d->in_w * C * 2,
                             (cuuint64_t)d->in_d * d->in_h * d->in_w * C * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)(g.box_w * st), (cuuint32_t)(g.box_h * st), (cuuint32_t)(g.box_d * st), (cuuint32_t)g.box_n};
    if (pl->pair) {   // dims listed as (C, W, D, H, N): the box {64, 10, 2, 10, 1} lands as rows [h][d][w]
      dims[2] = (cuuint64_t)d->in_d; dims[3] = (cuuint64_t)d->in_h;
      const cuuint64_t sh = strides[1], sd = strides[2];
      strides[1] = sd; strides[2] = sh;
      box[1] = 10; box[2] = 2; box[3] = 10; box[4] = 1;
    } else if (pl->halo) { box[1] = 10; box[2] = 18; box[3] = 1; box[4] = 1; }   // one halo d-plane slab
    else if (cl_n > 1) box[a_split_dim] = (cuuint32_t)a_split_ext;          // this CTA's share of the multicast A tile

    cuuint32_t es[5] = {1, (cuuint32_t)st, (cuuint32_t)st, (cuuint32_t)st, 1};
    CUresult r = enc(m, kTmapAct16, 5, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { b200dm_set_error("cuTensorMapEncodeTiled(A) failed: %d (C=%d dims %d,%d,%d,%d box %d,%d,%d,%d stride %d)", (int)r, C, d->in_w, d->in_h, d->in_d, d->batch, g.box_w, g.box_h, g.box_d, g.box_n, st); return B200DM_ERR_CUDA; }
    return B200DM_OK;
  };
  rc = encodeA(&pl->mapA0, x0, d->c0);
  if (rc) { delete pl; return rc; }
  rc = encodeA(&pl->mapA1, d->c1 > 0 ? x1 : x0, d->c1 > 0 ? d->c1 : d->c0);
  if (rc) { delete pl; return rc; }
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.ktot, (cuuint64_t)g.groups * g.n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)g.ktot * 2};
    if (d->mode == B200DM_CONV_BATCHED_GEMM) {
      // B is a raw activation tensor [batch][c_out rows][K = c0]; rows beyond c_out of a sample alias the next sample
      // (masked by the epilogue's col < c_out), K beyond c0 is zero-filled by TMA.
      dims[0] = (cuuint64_t)d->c0; dims[1] = (cuuint64_t)d->batch * d->c_out; strides[0] = (cuuint64_t)d->c0 * 2;
    }
    cuuint32_t box[2] = {64, (cuuint32_t)(g.block_n / cl_m)};   // this CTA's share of the multicast B tile
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r;
    if (pl->halo) {
      const long long t1 = (long long)((d->in_w + 7) / 8) * ((d->in_h + 15) / 16) * ((d->in_d + 1) / 2) * d->batch;
      {  // planes per tile: 2 (weight stage shared by two planes, two MMA issuer warps) unless that costs a wave
        const long long sms = b2_num_sms(), nt = g.n_pad / g.block_n;
        const long long t_one = (long long)((d->in_w + 7) / 8) * ((d->in_h + 15) / 16) * d->in_d * d->batch;
        const long long cost2 = (t1 * nt + sms - 1) / sms * 2, cost1 = (t_one * nt + sms - 1) / sms;
        pl->halo_td = d->in_d >= 2 && cost2 <= cost1 ? 2 : 1;
        if (pl->pair) pl->halo_td = 2;   // d step of a pair tile (the kernel runs with one accumulator, TD = 1)
      }
      halo_ring_config(g.block_n, &pl->halo_nb, &pl->halo_tps);
      // packed weights [n_pad rows][chunk*27 + tap][64] viewed as 3-D {64, rows, tap-chunks}: one box = TPS consecutive taps
      cuuint64_t dims3[3] = {64, (cuuint64_t)g.n_pad, (cuuint64_t)(g.ktot / 64)};
      cuuint64_t strides3[2] = {(cuuint64_t)g.ktot * 2, 128};
      cuuint32_t box3[3] = {64, (cuuint32_t)g.block_n, (cuuint32_t)pl->halo_tps};
      r = enc(&pl->mapB, kTmapAct16, 3, const_cast<void*>(w_packed), dims3, strides3, box3, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      r = enc(&pl->mapB, kTmapAct16, 2, const_cast<void*>(w_packed), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { b200dm_set_error("cuTensorMapEncodeTiled(B) failed: %d (halo %d)", (int)r, (int)pl->halo); delete pl; return B200DM_ERR_CUDA; }
  }
  ConvParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.mode = d->mode; p.batch = d->batch;
  p.in_d = d->in_d; p.in_h = d->in_h; p.in_w = d->in_w;
  p.out_d = g.out_d; p.out_h = g.out_h; p.out_w = g.out_w;
  p.m_d = g.m_d; p.m_h = g.m_h; p.m_w = g.m_w;
  p.box_w = g.box_w; p.box_h = g.box_h; p.box_d = g.box_d; p.box_n = g.box_n;
  p.tiles_w = (g.m_w + g.box_w - 1) / g.box_w; p.tiles_h = (g.m_h + g.box_h - 1) / g.box_h;
  p.tiles_d = (g.m_d + g.box_d - 1) / g.box_d; p.tiles_n = (d->batch + g.box_n - 1) / g.box_n;
  p.nch0 = g.nch0; p.nch1 = g.nch1; p.ntaps = g.ntaps; p.ksize = d->ksize; p.stride = st; p.pad = g.pad;
  p.ksteps0_last = ((d->c0 - (g.nch0 - 1) * 64) + 15) / 16;
  p.ksteps1_last = d->c1 > 0 ? ((d->c1 - (g.nch1 - 1) * 64) + 15) / 16 : 4;
  p.c_out = d->c_out; p.n_pad = d->mode == B200DM_CONV_BATCHED_GEMM ? d->c_out : g.n_pad;
  p.act = d->act; p.post_act = d->reserved[0]; p.y_f32 = d->y_dtype == B200DM_F32; p.transposed_store = d->reserved[1];
  p.chan_bias_rows = d->chan_bias_rows;
  p.bias = bias; p.chan_bias = chan_bias; p.t_dev = t_dev;
  p.residual = (const act_t*)residual; p.prelu_alpha = (const act_t*)prelu_alpha;
  p.y = y; p.dbg = g_dbg_flag;
  if (const char* e = tuning_env("B200DM_EPI_DBG")) p.epi_dbg = atoi(e);
  p.cl_m = cl_m; p.cl_n = cl_n; p.a_split_dim = a_split_dim; p.a_split_ext = a_split_ext;
  const long long mtiles = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  if (mtiles > 0x7fffffffLL) { delete pl; b200dm_set_error("conv_plan_create: too many tiles"); return B200DM_ERR_INVALID; }
  const int ntiles = d->mode == B200DM_CONV_BATCHED_GEMM ? (d->c_out + g.block_n - 1) / g.block_n : g.n_pad / g.block_n;
  pl->grid = dim3((unsigned)mtiles, (unsigned)ntiles, d->mode == B200DM_CONV_PARITY ? 8 : 1);
  // split-K: when the tile grid leaves most SMs idle and K is long, ksplit CTAs (one cluster) share a tile.  Each CTA
  // streams its K range at the per-SM operand rate (~40 B/clk), so an under-filled grid is K-latency-bound otherwise.
  p.ksplit = 1;
  if (!pl->halo && cl_m * cl_n == 1 && d->reserved[1] == 0) {
    const long long ctas = mtiles * ntiles * (d->mode == B200DM_CONV_PARITY ? 8 : 1);
    const int nkb = (g.nch0 + g.nch1) * g.ntaps;
    int ksp = 1;
    while (ksp < 8 && ctas * (ksp * 2) <= b2_num_sms() && nkb / (ksp * 2) >= 13) ksp *= 2;   // >= 13 k-blocks per CTA: 3^3 convs only
    if (const char* e = tuning_env("B200DM_KSPLIT")) { const int v = atoi(e); if (v >= 1 && v <= 8 && (v & (v - 1)) == 0 && nkb / v >= 1) ksp = v; }
    if (ksp > 1) { p.ksplit = ksp; pl->grid.x = (unsigned)(mtiles * ksp); }
  }
  // pipeline depth 4; measured on B200: a deeper ring (6 stages at BLOCK_N=128, 8 below) does not help at BLOCK_N=128 (the
  // operand stream is bandwidth-, not latency-bound) and hurts below 128, where 4 stages let two CTAs share an SM
  pl->nstage = 4;
  // short K loops (1^3 convs: 1-3 k-blocks) never fill four stages; two stages let 3-4 CTAs share an SM, which hides the
  // per-CTA prologue / TMA latency / epilogue of these HBM-bound launches
  if ((g.nch0 + g.nch1) * g.ntaps <= 3 && g.block_n >= 64 && cl_m * cl_n == 1 && !tuning_env("B200DM_IGEMM_NO2")) pl->nstage = 2;
  // grids of several waves: two stages x three co-resident CTAs keep more operand bytes in flight per SM than one CTA with
  // four stages, and hide every CTA's prologue / epilogue (measured: 16^3 128->128 parity conv 123 -> 77 us, 8^3 MLP GEMM
  // 17.5 -> 12.7 us; a single-wave split-K grid gets slower, so it keeps four stages)
  if (!pl->halo && (long long)pl->grid.x * pl->grid.y * pl->grid.z >= 2LL * b2_num_sms() && g.block_n >= 64 && p.ksplit <= 1 &&
      cl_m * cl_n == 1 && !tuning_env("B200DM_IGEMM_NO2"))
    pl->nstage = 2;
  if (const char* e = tuning_env("B200DM_IGEMM_STAGES")) { if (atoi(e) == 2 && g.block_n >= 64 && cl_m * cl_n == 1) pl->nstage = 2; if (atoi(e) == 4) pl->nstage = 4; }
  // staged (TMA-store) epilogue of the per-tap GEMM kernel: bf16 NDHWC output with whole 64-channel groups
  memset(&pl->om, 0, sizeof(pl->om));
  {
    const bool f32 = d->y_dtype == B200DM_F32;   // fp32 output (the U-Net's eps): halo kernel only, 32-column groups
    // (halo kernel, C_out = 32: one 32-channel group of 64-byte rows, SWIZZLE_64B -- the decoders' 128^3 convs)
    const bool n32 = pl->halo && !pl->pair && d->c_out == 32 && g.block_n == 32 && !f32 && (pl->ups || d->mode == B200DM_CONV_DIRECT);
    bool want = cl_m * cl_n == 1 && (!f32 || pl->halo) && d->reserved[1] == 0 && !prelu_alpha &&
                ((d->c_out % 64 == 0 && g.block_n >= 64) || n32) && !(d->mode == B200DM_CONV_PARITY && residual);
    if (const char* e = tuning_env("B200DM_TMA_EPI")) { if (atoi(e) == 0) want = false; }
    if (want) {
      const int par = d->mode == B200DM_CONV_PARITY ? 8 : 1;
      const int ps = d->mode == B200DM_CONV_PARITY ? 2 : 1;   // output voxel stride of one M-space step
      const cuuint64_t C = (cuuint64_t)d->c_out;
      auto encodeY = [&](CUtensorMap* m, const void* base) -> bool {
        cuuint64_t dims[5] = {C, (cuuint64_t)g.m_w, (cuuint64_t)g.m_h, (cuuint64_t)g.m_d, (cuuint64_t)d->batch};
        const cuuint64_t eb = f32 ? 4 : 2;
        cuuint64_t strides[4] = {C * eb * ps, (cuuint64_t)g.out_w * C * eb * ps, (cuuint64_t)g.out_h * g.out_w * C * eb * ps,
                                 (cuuint64_t)g.out_d * g.out_h * g.out_w * C * eb};
        cuuint32_t box[5] = {f32 || n32 ? 32u : 64u, (cuuint32_t)g.box_w, (cuuint32_t)g.box_h, (cuuint32_t)g.box_d, (cuuint32_t)g.box_n};
        if (pl->pair) {   // (C, W, D, H, N) order, one 8w x 2d x 8h tile
          dims[2] = (cuuint64_t)g.m_d; dims[3] = (cuuint64_t)g.m_h;
          const cuuint64_t sh = strides[1], sd = strides[2];
          strides[1] = sd; strides[2] = sh;
          box[1] = 8; box[2] = 2; box[3] = 8; box[4] = 1;
        } else if (pl->halo) { box[1] = 8; box[2] = 16; box[3] = 1; box[4] = 1; }   // one output plane of the halo kernel's tile
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        return enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : kTmapAct16, 5, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   n32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      };
      bool okm = true;
      for (int q = 0; q < par && okm; ++q) {
        const size_t off = d->mode == B200DM_CONV_PARITY
                               ? ((size_t)((q >> 2) & 1) * g.out_h * g.out_w + (size_t)((q >> 1) & 1) * g.out_w + (size_t)(q & 1)) * d->c_out * 2
                               : 0;
        okm = encodeY(&pl->om.y[q], (const char*)y + off);
      }
      if (okm && residual && !pl->halo) okm = encodeY(&pl->om.r, residual);
      if (!okm) { delete pl; b200dm_set_error("cuTensorMapEncodeTiled(Y) failed"); return B200DM_ERR_CUDA; }
      p.tma_epi = 1;
    }
  }
  // per-sample transposed store by operand swap (see the kernel): tiles of whole planes inside one sample, bias only
  if (d->reserved[1] != 0 && !pl->halo && cl_m * cl_n == 1 && d->mode == B200DM_CONV_DIRECT && d->ksize == 1 && st == 1 &&
      d->y_dtype == B200DM_BF16 && g.block_n == 128 && g.box_n == 1 && g.box_w == g.m_w && g.box_h == g.m_h && g.m_d % g.box_d == 0 &&
      !residual && !prelu_alpha && !chan_bias && d->act == B200DM_ACT_NONE && d->reserved[0] == B200DM_ACT_NONE &&
      !(tuning_env("B200DM_TMA_EPI") && atoi(tuning_env("B200DM_TMA_EPI")) == 0)) {
    const cuuint64_t Lv = (cuuint64_t)g.m_w * g.m_h * g.m_d;
    cuuint64_t dims[5] = {Lv, (cuuint64_t)d->c_out, 1, 1, (cuuint64_t)d->batch};
    cuuint64_t strides[4] = {Lv * 2, Lv * d->c_out * 2, Lv * d->c_out * 2, Lv * d->c_out * 2};
    cuuint32_t box[5] = {64, 128, 1, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    if (Lv % 8 == 0 && enc(&pl->om.y[0], kTmapAct16, 5, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
      p.swap_ab = 1;
      p.tma_epi = 1;
    }
  }
  pl->smem = conv_smem_bytes(g.block_n, pl->nstage, p.tma_epi && residual);
  if (pl->smem > 232448) { p.tma_epi = 0; pl->smem = conv_smem_bytes(g.block_n, pl->nstage, false); }
  // BLOCK_N = 128 with the staged epilogue: one-tap weight stages hold only 8 MMAs (4 per issuer warp) and the per-stage
  // barrier round trip (~200 cycles) made those convs issue-bound (45 % of the MMA floor); 3-tap stages x 2 fit once the
  // slab ring is the 4 slabs a channel chunk needs (the next chunk's slab i reloads as soon as slab i is released).
  if (pl->halo && !pl->pair && g.block_n == 128 && p.tma_epi && !tuning_env("B200DM_NO_WIDE")) {
    cuuint64_t dims3[3] = {64, (cuuint64_t)g.n_pad, (cuuint64_t)(g.ktot / 64)};
    cuuint64_t strides3[2] = {(cuuint64_t)g.ktot * 2, 128};
    cuuint32_t box3[3] = {64, 128, 3};
    cuuint32_t es3[3] = {1, 1, 1};
    if (enc(&pl->mapB, kTmapAct16, 3, const_cast<void*>(w_packed), dims3, strides3, box3, es3,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { delete pl; b200dm_set_error("cuTensorMapEncodeTiled(B wide) failed"); return B200DM_ERR_CUDA; }
    pl->wide = true; pl->halo_nb = 2; pl->halo_tps = 3;
  }
  // CTA pairs (cta_group::2) for every staged halo configuration with 64- or 128-channel tiles: needs an even tile count per
  // n-tile so that the two CTAs of a pair always work on the same n-tile and run the same number of tiles
  // the U-Net's input conv (256 -> 32, direct stores).  Short-K N=32 convs (the decoder's 32 -> 32 at 128^3) are bound by their
  // direct-store epilogue; pairing them only couples two epilogues (measured 4.18 -> 4.79 ms), so they stay single-CTA.
  // With the SWIZZLE_64B staged epilogue (n32 above) those convs run 3.78 ms, paired or not gated on tma_epi any more.
  const bool cg2_short = p.tma_epi && tuning_env("B200DM_CG2_N32S") && atoi(tuning_env("B200DM_CG2_N32S")) != 0;   // experiment: pair short-K staged N=32 convs too
  const bool cg2_n32 = g.block_n == 32 && !pl->pair && !pl->ups && pl->halo_td == 2 && !p.y2 && (g.nch0 + g.nch1 >= 2 || cg2_short);
  if (pl->halo && (((g.block_n == 64 || g.block_n == 128) && p.tma_epi && (pl->pair || pl->halo_td == 2 || g.block_n == 128)) || cg2_n32) &&
      !(tuning_env("B200DM_CG2") && atoi(tuning_env("B200DM_CG2")) == 0)) {
    const int td = pl->halo_td;   // (pair: d step 2)
    const long long per = (long long)((d->in_w + 7) / 8) * (pl->pair ? (d->in_h + 7) / 8 : (d->in_h + 15) / 16) * ((d->in_d + td - 1) / td) * d->batch;
    if (per % 2 == 0) {
      cuuint64_t dims3[3] = {64, (cuuint64_t)g.n_pad, (cuuint64_t)(g.ktot / 64)};
      cuuint64_t strides3[2] = {(cuuint64_t)g.ktot * 2, 128};
      cuuint32_t box3[3] = {64, (cuuint32_t)(g.block_n / 2), 3};   // this CTA's half of the rows of a 3-tap weight stage
      cuuint32_t es3[3] = {1, 1, 1};
      if (enc(&pl->mapB, kTmapAct16, 3, const_cast<void*>(w_packed), dims3, strides3, box3, es3,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { delete pl; b200dm_set_error("cuTensorMapEncodeTiled(B cg2) failed"); return B200DM_ERR_CUDA; }
      pl->cg2 = true; pl->wide = false; pl->halo_nb = cg2_n32 ? 4 : 3; pl->halo_tps = 3;
    }
  }
  if (pl->ups) {
    if (!p.tma_epi) { delete pl; b200dm_set_error("conv_plan_create: internal: up-conv plan without staged epilogue"); return B200DM_ERR_UNSUPPORTED; }
    // packed [parity][n_pad rows][chunk*8 + tap][64]: one box = this CTA's half of the rows x the 4 in-plane taps of one td
    cuuint64_t dims3[3] = {64, (cuuint64_t)g.groups * g.n_pad, (cuuint64_t)(g.ktot / 64)};
    cuuint64_t strides3[2] = {(cuuint64_t)g.ktot * 2, 128};
    cuuint32_t box3[3] = {64, (cuuint32_t)(g.block_n / 2), 4};
    cuuint32_t es3[3] = {1, 1, 1};
    if (enc(&pl->mapB, kTmapAct16, 3, const_cast<void*>(w_packed), dims3, strides3, box3, es3,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { delete pl; b200dm_set_error("cuTensorMapEncodeTiled(B up) failed"); return B200DM_ERR_CUDA; }
    pl->cg2 = true; pl->wide = false; pl->halo_td = 2;
    pl->halo_nb = g.block_n == 32 ? 4 : (g.block_n == 64 ? 3 : 2); pl->halo_tps = 4; pl->halo_ns = pl->pair ? 4 : 5;
  }
  if (pl->halo) {
    const int td = pl->halo_td;
    p.tiles_w = (d->in_w + 7) / 8; p.tiles_h = pl->pair ? (d->in_h + 7) / 8 : (d->in_h + 15) / 16; p.tiles_d = (d->in_d + td - 1) / td;
    p.tiles_n = d->batch;
    const long long per = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
    p.halo_td = td; p.halo_tiles_per_ntile = (int)per; p.halo_ntn = ntiles; p.halo_total_tiles = (int)(per * ntiles * (pl->ups ? 8 : 1));
    int ctas = b2_num_sms();
    if (ctas > p.halo_total_tiles) ctas = p.halo_total_tiles;
    if (pl->cg2) ctas &= ~1;   // whole pairs
    pl->grid = dim3((unsigned)ctas, 1, 1);
    pl->smem = halo_smem_bytes(pl);
  }
  // algorithmic FLOPs: 2*k^3*Cin*Cout*B*out_voxels; convT: 2*64*Cin*Cout*B*in_voxels; GEMM: 2*M*N*K
  if (d->mode == B200DM_CONV_PARITY)
    pl->flops = d->ksize == 3 ? 2.0 * 27 * (d->c0 + d->c1) * d->c_out * (double)d->batch * g.out_d * g.out_h * g.out_w
                              : 2.0 * 64 * (d->c0 + d->c1) * d->c_out * (double)d->batch * d->in_d * d->in_h * d->in_w;
  else
    pl->flops = 2.0 * g.ntaps * (d->c0 + d->c1) * d->c_out * (double)d->batch * g.out_d * g.out_h * g.out_w;
  *out = pl;
  return B200DM_OK;
}

extern "C" int b200dm_conv_plan_run(b200dm_conv_plan* pl, void* stream) {
  B2_CHECK_ARG(pl, "conv_plan_run: null plan");
  cudaStream_t s = (cudaStream_t)stream;
  if (pl->stencil) {
    auto kern = stencil::conv_stencil_c1_kernel;
    static bool attr_set = false;
    if (!attr_set) {
      B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stencil::kSmem));
      attr_set = true;
    }
    B2_CHECK_CUDA(b2_launch(kern, pl->grid, dim3(stencil::kThreads), pl->smem, s, pl->mapS, pl->sp));
    return B200DM_OK;
  }
  if (pl->sweep) {
    auto kern = sweep::conv_sweep32_kernel;
    static bool attr_set = false;
    if (!attr_set) {
      B2_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep::kSmem));
      attr_set = true;
    }
    B2_CHECK_CUDA(b2_launch(kern, pl->grid, dim3(sweep::kThreads), pl->smem, s, pl->mapS, pl->mapB, pl->om.y[0], pl->residual_map(), pl->p, pl->wp));
    return B200DM_OK;
  }
  if (pl->ups && pl->pair) return pl->g.block_n == 64 ? launch_halo_up<64, 4, 3, true>(pl, s) : launch_halo_up<128, 4, 2, true>(pl, s);
  if (pl->ups && pl->g.block_n == 32) return launch_halo_up<32, 5, 4>(pl, s);
  if (pl->ups) return pl->g.block_n == 64 ? launch_halo_up<64, 5, 3>(pl, s) : launch_halo_up<128, 5, 2>(pl, s);
  if (pl->cg2) {
    if (pl->fuse) return pl->halo_td == 2 ? launch_halo_cg2_fused<2>(pl, s) : launch_halo_cg2_fused<1>(pl, s);
    if (pl->pair) return launch_halo_cg2<64, 1, kHaloNSPair, 3, true>(pl, s);
    if (pl->g.block_n == 32) return pl->p.tma_epi ? launch_halo_cg2<32, 2, kHaloNS, 4, false, true>(pl, s)
                                                  : launch_halo_cg2<32, 2, kHaloNS, 4, false, false>(pl, s);
    if (pl->g.block_n == 64) return launch_halo_cg2<64, 2, kHaloNSStaged, 3, false>(pl, s);
    return pl->halo_td == 2 ? launch_halo_cg2<128, 2, kHaloNSStaged, 3, false>(pl, s) : launch_halo_cg2<128, 1, kHaloNSStaged, 3, false>(pl, s);
  }
  if (pl->pair) return launch_halo_pair(pl, s);
  if (pl->halo) {
    switch (pl->g.block_n) {
      case 16: return dispatch_halo<16, 4, 3>(pl, s);
      case 32: return dispatch_halo<32, 4, 3>(pl, s);
      case 64: return dispatch_halo<64, 3, 3>(pl, s);
      case 128:
        if (pl->wide) return pl->halo_td == 2 ? launch_halo_wide<2>(pl, s) : launch_halo_wide<1>(pl, s);
        return dispatch_halo<128, 4, 1>(pl, s);
    }
  }
  if (pl->nstage == 2) {
    switch (pl->g.block_n) {
      case 64: return launch_conv<64, 2>(pl, s);
      case 128: return launch_conv<128, 2>(pl, s);
    }
  }
  switch (pl->g.block_n) {
    case 16: return launch_conv<16, 4>(pl, s);
    case 32: return launch_conv<32, 4>(pl, s);
    case 64: return launch_conv<64, 4>(pl, s);
    case 128: return launch_conv<128, 4>(pl, s);
  }
  b200dm_set_error("conv_plan_run: unsupported BLOCK_N %d", pl->g.block_n);
  return B200DM_ERR_UNSUPPORTED;
}

extern "C" void b200dm_conv_plan_destroy(b200dm_conv_plan* p) { delete p; }

// tuning aid: device buffer of 4 * 2048 int64 receiving CTA 0's per-role timeline ((clock64 << 8) | tag); nullptr = off
extern "C" int b200dm_conv_plan_set_trace(b200dm_conv_plan* p, void* trace) {
  B2_CHECK_ARG(p, "conv_plan_set_trace: null plan");
  p->p.trace = (long long*)trace;
  return B200DM_OK;
}

extern "C" double b200dm_conv_plan_flops(const b200dm_conv_plan* p) { return p ? p->flops : 0.0; }

extern "C" int b200dm_conv_plan_add_output(b200dm_conv_plan* p, void* y_extra, const float* scale, const float* shift, int32_t act) {
  B2_CHECK_ARG(p && y_extra && scale && shift, "conv_plan_add_output: null argument");
  B2_CHECK_ARG(p->desc.c_out % 16 == 0 && p->desc.reserved[1] == 0, "conv_plan_add_output: needs c_out %% 16 == 0 and a plain (non-transposed) store");
  B2_CHECK_ARG(((uintptr_t)y_extra & 15) == 0 && ((uintptr_t)scale & 15) == 0 && ((uintptr_t)shift & 15) == 0, "conv_plan_add_output: pointers must be 16-byte aligned");
  if (p->pair || p->wide || p->cg2 || p->stencil || p->sweep) { b200dm_set_error("conv_plan_add_output: not available on pair-slab / wide-stage halo / stencil plans"); return B200DM_ERR_UNSUPPORTED; }
  if (p->p.tma_epi) { p->p.tma_epi = 0; p->smem = p->halo ? halo_smem_bytes(p) : conv_smem_bytes(p->g.block_n, p->nstage, false); }
  if (!p->p.y2) { p->p.y2 = (act_t*)y_extra; p->p.scale2 = scale; p->p.shift2 = shift; p->p.act2 = act; }
  else if (!p->p.y3) { p->p.y3 = (act_t*)y_extra; p->p.scale3 = scale; p->p.shift3 = shift; p->p.act3 = act; }
  else { b200dm_set_error("conv_plan_add_output: at most two extra outputs"); return B200DM_ERR_UNSUPPORTED; }
  return B200DM_OK;
}

extern "C" int b200dm_conv_plan_set_fused_update(b200dm_conv_plan* pl, const b200dm_update_desc* u, const float* x_t, float* x_prev,
                                                 void* x_prev_16) {
  B2_CHECK_ARG(pl && u && x_t && x_prev && x_prev_16, "conv_plan_set_fused_update: null argument");
  const b200dm_conv_desc* d = &pl->desc;
  const ConvParams& p = pl->p;
  // the CTA-pair N = 128 halo kernel with the fp32 staged epilogue, no other epilogue term
  if (!(pl->halo && pl->cg2 && !pl->pair && !pl->ups && !pl->sweep && !pl->stencil && pl->g.block_n == 128 &&
        p.y_f32 && p.tma_epi && d->c_out % 128 == 0 && !p.residual && !p.prelu_alpha && !p.y2 && p.act == B200DM_ACT_NONE &&
        p.post_act == B200DM_ACT_NONE)) {
    b200dm_set_error("conv_plan_set_fused_update: only on the CTA-pair fp32-output 3^3 conv with C_out %% 128 == 0");
    return B200DM_ERR_UNSUPPORTED;
  }
  B2_CHECK_ARG(u->sampler == 0 || u->sampler == 1, "conv_plan_set_fused_update: sampler must be 0 (ddpm) or 1 (ddim)");
  B2_CHECK_ARG(u->beta && u->sqrt_alpha && u->alpha_bar && u->alpha_bar_prev && u->sqrt_alpha_bar && u->sqrt_alpha_bar_prev &&
                   u->sqrt_one_minus_alpha_bar, "conv_plan_set_fused_update: null schedule table");
  B2_CHECK_ARG(u->batch == d->batch && u->n_per_sample == (int64_t)d->in_d * d->in_h * d->in_w * d->c_out,
               "conv_plan_set_fused_update: update geometry (%d x %lld) differs from the conv output", u->batch, (long long)u->n_per_sample);
  B2_CHECK_ARG((u->n_per_sample >> 2) < (1ll << 32), "conv_plan_set_fused_update: sample too large for the 32-bit Philox counter");
  B2_CHECK_ARG(x_t == x_prev, "conv_plan_set_fused_update: x_prev must alias x_t (the tile is updated in place through one tensor map)");
  B2_CHECK_ARG(((uintptr_t)x_t & 15) == 0 && ((uintptr_t)x_prev & 15) == 0 && ((uintptr_t)x_prev_16 & 15) == 0,
               "conv_plan_set_fused_update: pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode_fn();
  const cuuint64_t C = (cuuint64_t)d->c_out;
  cuuint64_t dims[5] = {C, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h, (cuuint64_t)d->in_d, (cuuint64_t)d->batch};
  cuuint32_t box[5] = {16, 8, 16, 1, 1}, es[5] = {1, 1, 1, 1, 1};   // one round of the fused epilogue: 16 columns x one 8 x 16 plane tile
  cuuint64_t st4[4] = {C * 4, (cuuint64_t)d->in_w * C * 4, (cuuint64_t)d->in_h * d->in_w * C * 4, (cuuint64_t)d->in_d * d->in_h * d->in_w * C * 4};
  cuuint64_t st2[4] = {C * 2, (cuuint64_t)d->in_w * C * 2, (cuuint64_t)d->in_h * d->in_w * C * 2, (cuuint64_t)d->in_d * d->in_h * d->in_w * C * 2};
  CUtensorMap my, mz;
  if (!enc ||
      enc(&my, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, x_prev, dims, st4, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
      enc(&mz, kTmapAct16, 5, x_prev_16, dims, st2, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    b200dm_set_error("conv_plan_set_fused_update: cuTensorMapEncodeTiled failed");
    return B200DM_ERR_CUDA;
  }
  pl->om.y[0] = my; pl->om.y[1] = mz;
  pl->p.upd = *u; pl->p.upd_x = x_t;
  pl->fuse = true;
  pl->smem = halo_smem_bytes(pl);
  if (pl->smem > 232448) { b200dm_set_error("conv_plan_set_fused_update: shared memory"); return B200DM_ERR_UNSUPPORTED; }
  return B200DM_OK;
}

extern "C" int b200dm_conv_plan_set_side_norm(b200dm_conv_plan* pl, void* y_side, const float* scale, const float* shift, int32_t act) {
  B2_CHECK_ARG(pl && y_side && scale && shift, "conv_plan_set_side_norm: null argument");
  const b200dm_conv_desc* d = &pl->desc;
  const int C = d->c0 + d->c1;
  if (pl->halo || pl->stencil || !pl->p.tma_epi || pl->p.swap_ab || pl->p.ksplit > 1 || pl->p.cl_m * pl->p.cl_n != 1 || d->mode != B200DM_CONV_DIRECT ||
      d->ksize != 1 || d->stride != 1 || pl->g.block_n < 64 || (d->c1 > 0 && d->c0 % 64 != 0) || C % 8 != 0 || C > 1024 ||
      ((uintptr_t)y_side & 15) != 0) {
    b200dm_set_error("conv_plan_set_side_norm: only on staged 1^3 stride-1 convs (c0 %% 64 == 0 when c1 > 0)");
    return B200DM_ERR_UNSUPPORTED;
  }
  EncodeTiledFn enc = get_encode_fn();
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h, (cuuint64_t)d->in_d, (cuuint64_t)d->batch};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)d->in_w * C * 2, (cuuint64_t)d->in_h * d->in_w * C * 2,
                           (cuuint64_t)d->in_d * d->in_h * d->in_w * C * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)pl->g.box_w, (cuuint32_t)pl->g.box_h, (cuuint32_t)pl->g.box_d, (cuuint32_t)pl->g.box_n};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  if (!enc || enc(&pl->om.s, kTmapAct16, 5, y_side, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
    b200dm_set_error("conv_plan_set_side_norm: cuTensorMapEncodeTiled failed");
    return B200DM_ERR_CUDA;
  }
  const size_t smem = conv_smem_bytes(pl->g.block_n, pl->nstage, pl->p.residual != nullptr, C);
  if (smem > conv_smem_attr(pl->g.block_n, pl->nstage)) { b200dm_set_error("conv_plan_set_side_norm: shared memory"); return B200DM_ERR_UNSUPPORTED; }
  pl->smem = smem;
  pl->p.side = 1; pl->p.side_act = act; pl->p.side_c = C; pl->p.side_scale = scale; pl->p.side_shift = shift;
  return B200DM_OK;
}

extern "C" int b200dm_conv_plan_info(const b200dm_conv_plan* p, int32_t* halo, int32_t* block_n, int32_t* ksplit) {
  B2_CHECK_ARG(p, "conv_plan_info: null plan");
  if (halo) *halo = p->stencil ? 2 : (p->sweep ? 3 : (p->halo ? 1 : 0));   // 2 = HBM-bound stencil-reduce kernel (C_out = 1), 3 = d-sweeping N = 96 kernel
  if (block_n) *block_n = p->g.block_n;
  if (ksplit) *ksplit = p->p.ksplit;
  return B200DM_OK;
}

extern "C" int b200dm_conv_plan_set_out_affine(b200dm_conv_plan* p, const float* scale, const float* shift) {
  B2_CHECK_ARG(p, "conv_plan_set_out_affine: null plan");
  B2_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_plan_set_out_affine: give both scale and shift, or neither");
  B2_CHECK_ARG(!scale || !p->p.prelu_alpha, "conv_plan_set_out_affine: not combinable with PReLU");
  B2_CHECK_ARG(!scale || !p->p.swap_ab, "conv_plan_set_out_affine: not combinable with a transposed store");
  B2_CHECK_ARG(!scale || !p->stencil, "conv_plan_set_out_affine: not available on the C_out = 1 stencil plan");
  p->p.out_scale = scale;
  p->p.out_shift = shift;
  return B200DM_OK;
}

// GroupNorm statistics as a by-product of the producing conv (d-sweeping kernel only): the epilogue accumulates per-(item, channel)
// partial sums of the STORED values into `workspace` (float [rows][C][2], rows = batch * rows_per_sample); b200dm_gn_finalize
// turns them into (mean, rstd) in a fixed order.  Returns the workspace size, 0 when the plan cannot produce partials.
extern "C" size_t b200dm_conv_plan_gn_partials_bytes(const b200dm_conv_plan* p, int32_t* rows_per_sample) {
  if (!p || !p->sweep) return 0;
  if (rows_per_sample) *rows_per_sample = p->wp.items / p->wp.batch * 2;
  return (size_t)p->wp.items * 2 * 32 * 2 * sizeof(float);
}

extern "C" int b200dm_conv_plan_set_gn_partials(b200dm_conv_plan* p, float* workspace, size_t ws_bytes) {
  B2_CHECK_ARG(p && workspace, "conv_plan_set_gn_partials: null argument");
  const size_t need = b200dm_conv_plan_gn_partials_bytes(p, nullptr);
  if (need == 0) { b200dm_set_error("conv_plan_set_gn_partials: this plan cannot produce GroupNorm partial sums"); return B200DM_ERR_UNSUPPORTED; }
  B2_CHECK_ARG(ws_bytes >= need, "conv_plan_set_gn_partials: workspace too small (%zu < %zu)", ws_bytes, need);
  p->wp.gn_part = workspace;
  return B200DM_OK;
}

extern "C" int b200dm_gn_finalize(const float* partials, int32_t batch, int32_t rows_per_sample, int32_t c, int32_t groups,
                                  int64_t voxels_per_sample, float eps, float* mean_rstd, void* stream) {
  B2_CHECK_ARG(partials && mean_rstd && batch > 0 && rows_per_sample > 0 && c > 0 && groups > 0 && c % groups == 0 && voxels_per_sample > 0,
               "gn_finalize: bad argument");
  const double count = (double)voxels_per_sample * (c / groups);
  B2_CHECK_CUDA(b2_launch(sweep::gn_finalize_kernel, dim3((unsigned)(batch * groups)), dim3(128), 0, (cudaStream_t)stream, partials,
                          rows_per_sample, 1, c, groups, count, eps, mean_rstd));
  return B200DM_OK;
}

// Fold the consumer-side GroupNorm + activation into the conv's operand path (d-sweeping kernel only): the conv then computes
// conv(act(gamma * (x - mean) * rstd + beta)) from the RAW x, mean / rstd per (sample, group) read from `mean_rstd` at run time.
extern "C" int b200dm_conv_plan_set_input_norm(b200dm_conv_plan* p, const float* mean_rstd, const float* gamma, const float* beta,
                                               int32_t groups, int32_t act) {
  B2_CHECK_ARG(p && mean_rstd && gamma && beta, "conv_plan_set_input_norm: null argument");
  if (!p->sweep) { b200dm_set_error("conv_plan_set_input_norm: this plan's kernel has no input transform"); return B200DM_ERR_UNSUPPORTED; }
  B2_CHECK_ARG(groups > 0 && 32 % groups == 0, "conv_plan_set_input_norm: groups must divide 32");
  p->wp.in_mr = mean_rstd; p->wp.in_gamma = gamma; p->wp.in_beta = beta; p->wp.in_groups = groups; p->wp.in_act = act;
  return B200DM_OK;
}
