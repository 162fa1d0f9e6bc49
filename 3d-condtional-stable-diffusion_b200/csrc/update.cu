// K10: fused reverse-diffusion update with an in-register Philox4x32-10 noise stream.
// One HBM pass: read x_t (fp32) + eps (bf16|fp32) [+ injected noise], write x_{t-1} (fp32) [+ bf16 copy
// for the next U-Net input conv].  Replaces ~10 TF elementwise ops + tf.random.normal per step
// (networks/dm3d.py:477-508, 516-530).  Arithmetic follows the reference op by op in fp32
// (__fmul_rn/__fadd_rn/__fdiv_rn forbid FMA contraction so the result matches separate TF ops).
#include "common.cuh"
#include "update_math.cuh"

namespace {

using namespace upd;

// 8 elements per thread per trip (two Philox counters), two trips in flight: 4x LDG.128 x_t, 2-4x LDG.128 eps issued before
// any use; grid = (blocks, batch) so the loop has no index division.
template <bool kEpsBf16>
__global__ void __launch_bounds__(256, 4) update_kernel(b200dm_update_desc d, const float* __restrict__ x_t,
                                                     const void* __restrict__ eps, const float* __restrict__ noise,
                                                     float* __restrict__ x_prev, act_t* __restrict__ x_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const Coef k = load_coef(d);
  const uint32_t n8 = (uint32_t)(d.n_per_sample >> 3);
  const int64_t b = blockIdx.y, base = b * d.n_per_sample;
  const bool gen = (noise == nullptr) && d.sampler == 0 && k.t > 0;
  const bool inj = noise != nullptr && k.t > 0;
  // reserved bit 0: the Philox key and the global index of sample 0 live in device memory (t_dev[4..7]), so a captured step
  // graph serves every seed / shard without re-capture; desc.sample_id0 is then this launch's offset inside the call's batch
  uint64_t seed = d.seed;
  int64_t sid0 = d.sample_id0;
  if ((d.reserved & 1) && d.t_dev) {
    seed = (uint64_t)(uint32_t)d.t_dev[4] | ((uint64_t)(uint32_t)d.t_dev[5] << 32);
    sid0 += (int64_t)((uint64_t)(uint32_t)d.t_dev[6] | ((uint64_t)(uint32_t)d.t_dev[7] << 32));
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  constexpr int U = 2;
  for (uint32_t j0 = blockIdx.x * blockDim.x + threadIdx.x; j0 < n8; j0 += stride * U) {
    float x[U][8], e[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t j = j0 + u * stride;
      if (j >= n8) break;
      const int64_t off = base + ((int64_t)j << 3);
      *reinterpret_cast<float4*>(&x[u][0]) = __ldg(reinterpret_cast<const float4*>(x_t + off));
      *reinterpret_cast<float4*>(&x[u][4]) = __ldg(reinterpret_cast<const float4*>(x_t + off + 4));
      if (kEpsBf16) {
        unpack8(ldg_bf16x8(reinterpret_cast<const bf16x8*>(reinterpret_cast<const act_t*>(eps) + off)), e[u]);
      } else {
        *reinterpret_cast<float4*>(&e[u][0]) = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eps) + off));
        *reinterpret_cast<float4*>(&e[u][4]) = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eps) + off + 4));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t j = j0 + u * stride;
      if (j >= n8) break;
      const int64_t off = base + ((int64_t)j << 3);
      float z[8];
      if (gen) {
        float za[4], zb[4];
        normal4(2 * j, (uint32_t)k.t, (uint32_t)(sid0 + b), 0u, seed, za);
        normal4(2 * j + 1, (uint32_t)k.t, (uint32_t)(sid0 + b), 0u, seed, zb);
#pragma unroll
        for (int q = 0; q < 4; ++q) { z[q] = za[q]; z[4 + q] = zb[q]; }
      } else if (inj) {
        *reinterpret_cast<float4*>(&z[0]) = __ldg(reinterpret_cast<const float4*>(noise + off));
        *reinterpret_cast<float4*>(&z[4]) = __ldg(reinterpret_cast<const float4*>(noise + off + 4));
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) z[q] = 0.0f;
      }
      float y[8];
      step_vec(k, d.sampler, x[u], e[u], z, y);
      *reinterpret_cast<float4*>(x_prev + off) = *reinterpret_cast<float4*>(&y[0]);
      *reinterpret_cast<float4*>(x_prev + off + 4) = *reinterpret_cast<float4*>(&y[4]);
      if (x_bf16) *reinterpret_cast<bf16x8*>(x_bf16 + off) = pack8(y);
    }
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ x, act_t* __restrict__ xb,
                                                            int64_t n_per_sample, int batch, uint64_t seed,
                                                            int64_t sample_id0, int step, int stream_id) {
  const int64_t n4 = n_per_sample >> 2, total = n4 * batch;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / n4, j = i - b * n4;
    float z[4];
    normal4((uint32_t)j, (uint32_t)step, (uint32_t)(sample_id0 + b), (uint32_t)stream_id, seed, z);
    const int64_t off = b * n_per_sample + (j << 2);
    *reinterpret_cast<float4*>(x + off) = *reinterpret_cast<float4*>(&z[0]);
    if (xb) {
      *reinterpret_cast<act2_t*>(xb + off) = floats_to_act2(z[0], z[1]);
      *reinterpret_cast<act2_t*>(xb + off + 2) = floats_to_act2(z[2], z[3]);
    }
  }
}

__global__ void step_advance_kernel(int32_t* t_dev, int32_t delta) {
  pdl_launch_dependents();
  pdl_wait();
  t_dev[0] += delta;
  t_dev[1] += delta;
}

// t_dev = [t, t_prev, idx, -]: move to the next entry of a device-resident timestep sequence (terminated by -1), so one
// captured step graph replays ANY sequence -- every DDPM range, strided DDIM, non-uniform spacings -- with no host work.
__global__ void step_advance_seq_kernel(int32_t* t_dev, const int32_t* __restrict__ seq) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = t_dev[2] + 1;
  t_dev[2] = i;
  const int t = seq[i];
  t_dev[0] = t < 0 ? 0 : t;                 // past the end: stay on a valid table row (the step's result is unused)
  t_dev[1] = t < 0 ? -1 : seq[i + 1];
}

}  // namespace

extern "C" int b200dm_ddpm_update(const b200dm_update_desc* d, const float* x_t, const void* eps,
                                  const float* noise, float* x_prev, void* x_prev_bf16, void* stream) {
  B2_CHECK_ARG(d && x_t && eps && x_prev, "ddpm_update: null argument");
  B2_CHECK_ARG(d->n_per_sample > 0 && d->n_per_sample % 8 == 0, "ddpm_update: n_per_sample must be a positive multiple of 8");
  B2_CHECK_ARG(d->batch > 0, "ddpm_update: batch must be > 0");
  B2_CHECK_ARG(d->sampler == 0 || d->sampler == 1, "ddpm_update: sampler must be 0 (ddpm) or 1 (ddim)");
  B2_CHECK_ARG(d->beta && d->sqrt_alpha && d->alpha_bar && d->alpha_bar_prev && d->sqrt_alpha_bar &&
                   d->sqrt_alpha_bar_prev && d->sqrt_one_minus_alpha_bar, "ddpm_update: null schedule table");
  B2_CHECK_ARG(d->t_dev || d->t >= 0, "ddpm_update: negative timestep");
  B2_CHECK_ARG((d->n_per_sample >> 3) < (1ll << 31) && d->batch <= 65535, "ddpm_update: sample too large / batch > 65535");
  const int64_t per = (d->n_per_sample >> 3), want = (per + 2 * 256 - 1) / (2 * 256);
  int64_t cap = ((int64_t)b2_num_sms() * 4 + d->batch - 1) / d->batch;   // one resident wave: 4 blocks per SM (62 registers)
  if (cap < 1) cap = 1;
  const dim3 grid((unsigned)(want < cap ? want : cap), (unsigned)d->batch);
  cudaStream_t s = (cudaStream_t)stream;
  if (d->eps_dtype == B200DM_BF16)
    B2_CHECK_CUDA(b2_launch(update_kernel<true>, grid, dim3(256), 0, s, *d, x_t, eps, noise, x_prev, (act_t*)x_prev_bf16));
  else
    B2_CHECK_CUDA(b2_launch(update_kernel<false>, grid, dim3(256), 0, s, *d, x_t, eps, noise, x_prev, (act_t*)x_prev_bf16));
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}

extern "C" int b200dm_philox_normal(float* x, void* x_bf16, int64_t n_per_sample, int32_t batch, uint64_t seed,
                                    int64_t sample_id0, int32_t step, int32_t stream_id, void* stream) {
  B2_CHECK_ARG(x && n_per_sample > 0 && n_per_sample % 4 == 0 && batch > 0, "philox_normal: bad arguments");
  const int64_t total = (n_per_sample >> 2) * batch;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)b2_num_sms() * 8 ? want : (int64_t)b2_num_sms() * 8);
  philox_normal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (act_t*)x_bf16, n_per_sample, batch, seed,
                                                              sample_id0, step, stream_id);
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}

extern "C" int b200dm_step_advance_seq(int32_t* t_dev, const int32_t* seq, void* stream) {
  B2_CHECK_ARG(t_dev && seq, "step_advance_seq: null argument");
  B2_CHECK_CUDA(b2_launch(step_advance_seq_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, t_dev, seq));
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}

extern "C" int b200dm_step_advance(int32_t* t_dev, int32_t delta, void* stream) {
  B2_CHECK_ARG(t_dev, "step_advance: null t_dev");
  B2_CHECK_CUDA(b2_launch(step_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, t_dev, delta));
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}
