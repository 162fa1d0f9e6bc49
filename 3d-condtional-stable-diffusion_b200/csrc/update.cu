// K10: fused reverse-diffusion update with an in-register Philox4x32-10 noise stream.
// One HBM pass: read x_t (fp32) + eps (bf16|fp32) [+ injected noise], write x_{t-1} (fp32) [+ bf16 copy
// for the next U-Net input conv].  Replaces ~10 TF elementwise ops + tf.random.normal per step
// (networks/dm3d.py:477-508, 516-530).  Arithmetic follows the reference op by op in fp32
// (__fmul_rn/__fadd_rn/__fdiv_rn forbid FMA contraction so the result matches separate TF ops).
#include "common.cuh"

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& z0, float& z1) {
  const float u1 = __fadd_rn(__fmul_rn((float)(ra >> 8), 5.9604644775390625e-8f), 2.98023223876953125e-8f);
  const float u2 = __fmul_rn((float)(rb >> 8), 5.9604644775390625e-8f);
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

__device__ __forceinline__ void normal4(uint32_t ctr, uint32_t step, uint32_t sample, uint32_t stream, uint64_t seed,
                                        float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(ctr, step, sample, stream, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  box_muller(r[0], r[1], z[0], z[1]);
  box_muller(r[2], r[3], z[2], z[3]);
}

struct Coef {
  float sq1ab, sqab, c1, c2, sigma, sqab_p, sq1ab_p;
  int t, t_prev;
};

__device__ __forceinline__ Coef load_coef(const b200dm_update_desc& d) {
  Coef k;
  k.t = d.t_dev ? d.t_dev[0] : d.t;
  k.t_prev = d.t_dev ? d.t_dev[1] : d.t_prev;
  const int t = k.t;
  const float b = d.beta[t], sqa = d.sqrt_alpha[t], ab = d.alpha_bar[t], abp = d.alpha_bar_prev[t];
  const float sqabp = d.sqrt_alpha_bar_prev[t];
  k.sqab = d.sqrt_alpha_bar[t];
  k.sq1ab = d.sqrt_one_minus_alpha_bar[t];
  const float om = __fsub_rn(1.0f, ab);
  k.c1 = __fdiv_rn(__fmul_rn(b, sqabp), om);                     // b*sqab_prev/(1-ab)
  k.c2 = __fdiv_rn(__fmul_rn(__fsub_rn(1.0f, abp), sqa), om);     // (1-ab_prev)*sqa/(1-ab)
  const float var = __fdiv_rn(__fmul_rn(__fsub_rn(1.0f, abp), b), om);
  k.sigma = expf(0.5f * logf(fmaxf(var, 1e-20f)));                // exp(0.5*log(max(var,1e-20)))
  if (d.sampler == 1 && k.t_prev >= 0) {
    k.sqab_p = d.sqrt_alpha_bar[k.t_prev];
    k.sq1ab_p = d.sqrt_one_minus_alpha_bar[k.t_prev];
  } else {
    k.sqab_p = 1.0f; k.sq1ab_p = 0.0f;
  }
  return k;
}

__device__ __forceinline__ float step_one(const Coef& k, int sampler, float x, float e, float z) {
  const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sq1ab, e)), k.sqab);
  if (sampler == 0) {
    float mean = __fadd_rn(__fmul_rn(k.c1, x0), __fmul_rn(k.c2, x));
    mean = fminf(fmaxf(mean, -1.0f), 1.0f);
    return (k.t > 0) ? __fadd_rn(mean, __fmul_rn(k.sigma, z)) : mean;
  }
  const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
  if (k.t_prev < 0) return x0c;
  return __fadd_rn(__fmul_rn(k.sqab_p, x0c), __fmul_rn(k.sq1ab_p, e));
}

// 8 elements per thread per iteration (two Philox counters): 2x LDG.128 x_t, 1-2x LDG.128 eps,
// 2x STG.128 fp32 + 1x STG.128 bf16.
template <bool kEpsBf16>
__global__ void __launch_bounds__(256) update_kernel(b200dm_update_desc d, const float* __restrict__ x_t,
                                                     const void* __restrict__ eps, const float* __restrict__ noise,
                                                     float* __restrict__ x_prev, __nv_bfloat16* __restrict__ x_bf16) {
  pdl_launch_dependents();
  pdl_wait();
  const Coef k = load_coef(d);
  const int64_t n8 = d.n_per_sample >> 3;
  const int64_t total = n8 * d.batch;
  const bool gen = (noise == nullptr) && d.sampler == 0 && k.t > 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / n8, j = i - b * n8;
    const int64_t off = b * d.n_per_sample + (j << 3);
    float x[8], e[8], z[8];
    *reinterpret_cast<float4*>(&x[0]) = __ldg(reinterpret_cast<const float4*>(x_t + off));
    *reinterpret_cast<float4*>(&x[4]) = __ldg(reinterpret_cast<const float4*>(x_t + off + 4));
    if (kEpsBf16) {
      bf16x8 p = *reinterpret_cast<const bf16x8*>(reinterpret_cast<const __nv_bfloat16*>(eps) + off);
      unpack8(p, e);
    } else {
      *reinterpret_cast<float4*>(&e[0]) = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eps) + off));
      *reinterpret_cast<float4*>(&e[4]) = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eps) + off + 4));
    }
    if (gen) {
      float za[4], zb[4];
      normal4((uint32_t)(2 * j), (uint32_t)k.t, (uint32_t)(d.sample_id0 + b), 0u, d.seed, za);
      normal4((uint32_t)(2 * j + 1), (uint32_t)k.t, (uint32_t)(d.sample_id0 + b), 0u, d.seed, zb);
#pragma unroll
      for (int q = 0; q < 4; ++q) { z[q] = za[q]; z[4 + q] = zb[q]; }
    } else if (noise != nullptr && k.t > 0) {
      *reinterpret_cast<float4*>(&z[0]) = __ldg(reinterpret_cast<const float4*>(noise + off));
      *reinterpret_cast<float4*>(&z[4]) = __ldg(reinterpret_cast<const float4*>(noise + off + 4));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) z[q] = 0.0f;
    }
    float y[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) y[q] = step_one(k, d.sampler, x[q], e[q], z[q]);
    *reinterpret_cast<float4*>(x_prev + off) = *reinterpret_cast<float4*>(&y[0]);
    *reinterpret_cast<float4*>(x_prev + off + 4) = *reinterpret_cast<float4*>(&y[4]);
    if (x_bf16) *reinterpret_cast<bf16x8*>(x_bf16 + off) = pack8(y);
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
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
      *reinterpret_cast<__nv_bfloat162*>(xb + off) = __floats2bfloat162_rn(z[0], z[1]);
      *reinterpret_cast<__nv_bfloat162*>(xb + off + 2) = __floats2bfloat162_rn(z[2], z[3]);
    }
  }
}

__global__ void step_advance_kernel(int32_t* t_dev, int32_t delta) {
  pdl_launch_dependents();
  pdl_wait();
  t_dev[0] += delta;
  t_dev[1] += delta;
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
  const int64_t total = (d->n_per_sample >> 3) * d->batch;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)b2_num_sms() * 8 ? want : (int64_t)b2_num_sms() * 8);
  cudaStream_t s = (cudaStream_t)stream;
  if (d->eps_dtype == B200DM_BF16)
    B2_CHECK_CUDA(b2_launch(update_kernel<true>, dim3(grid), dim3(256), 0, s, *d, x_t, eps, noise, x_prev, (__nv_bfloat16*)x_prev_bf16));
  else
    B2_CHECK_CUDA(b2_launch(update_kernel<false>, dim3(grid), dim3(256), 0, s, *d, x_t, eps, noise, x_prev, (__nv_bfloat16*)x_prev_bf16));
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}

extern "C" int b200dm_philox_normal(float* x, void* x_bf16, int64_t n_per_sample, int32_t batch, uint64_t seed,
                                    int64_t sample_id0, int32_t step, int32_t stream_id, void* stream) {
  B2_CHECK_ARG(x && n_per_sample > 0 && n_per_sample % 4 == 0 && batch > 0, "philox_normal: bad arguments");
  const int64_t total = (n_per_sample >> 2) * batch;
  const int64_t want = (total + 255) / 256;
  const int grid = (int)(want < (int64_t)b2_num_sms() * 8 ? want : (int64_t)b2_num_sms() * 8);
  philox_normal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)x_bf16, n_per_sample, batch, seed,
                                                              sample_id0, step, stream_id);
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}

extern "C" int b200dm_step_advance(int32_t* t_dev, int32_t delta, void* stream) {
  B2_CHECK_ARG(t_dev, "step_advance: null t_dev");
  B2_CHECK_CUDA(b2_launch(step_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, t_dev, delta));
  B2_CHECK_LAUNCH();
  return B200DM_OK;
}
