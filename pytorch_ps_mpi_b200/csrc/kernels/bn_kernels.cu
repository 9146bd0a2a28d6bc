// Fused training BatchNorm for channels-last (NHWC) bf16 activations on sm_100a.
//
// Why it exists: the launch list of one ResNet-18 step (profiles/resnet18_step_launches_n1.txt)
// shows ATen's channels-last BatchNorm + the separate add / ReLU kernels taking ~45 % of the step
// at ~0.5 TB/s.  BatchNorm is pure HBM traffic, so the framework ships its own:
//   forward : psb_bn_stats   (1 read of x, fp32 sum / sum-of-squares per channel)
//             psb_bn_finalize (C threads: mean, rstd, running stats, fused scale/shift)
//             psb_bn_apply   (read x [+ residual] → y = act(x*scale + shift [+ residual]), 1 write)
//   backward: psb_bn_bwd_reduce (read dy, x, y → Σdy', Σdy'·x̂ with the ReLU mask folded in)
//             psb_bn_bwd_finalize
//             psb_bn_bwd_apply  (write dx and, for residual blocks, the masked skip gradient)
// Every pass moves 16 bytes per thread, a warp touches 512 contiguous bytes, accumulation is fp32.
#include <cuda_bf16.h>

#include "kernels.h"

namespace {
using namespace psb;

constexpr int BN_THREADS = 256;

struct BnGeom {
  long long pixels;   // N*H*W
  int C;              // channels (multiple of 8)
  int groups;         // C / 8
  int lanes;          // BN_THREADS / groups  (pixels processed in parallel by a CTA)
};

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* f) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x8(v, f);
}
__device__ __forceinline__ void ld8_stream(const __nv_bfloat16* p, float* f) {
  uint4 v = ld_stream_v4(p);
  unpack_bf16x8(v, f);
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                             pack_bf16x2(f[6], f[7]));
}

// Block-level reduction of per-thread 8-channel partials over the pixel lanes, then one atomicAdd
// per channel per CTA.  smem layout: [lanes][C] floats.
__device__ __forceinline__ void reduce_lanes_atomic(float* smem, const float* part, int tx, int ty, const BnGeom& g,
                                                    float* out) {
  if (ty < g.lanes) {   // threads beyond lanes*groups (C/8 not dividing 256) hold no partials
    float4* row = reinterpret_cast<float4*>(smem + (size_t)ty * g.C + tx * 8);
    row[0] = make_float4(part[0], part[1], part[2], part[3]);
    row[1] = make_float4(part[4], part[5], part[6], part[7]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < g.C; c += BN_THREADS) {
    float s = 0.f;
    for (int l = 0; l < g.lanes; ++l) s += smem[(size_t)l * g.C + c];
    atomicAdd(out + c, s);
  }
  __syncthreads();
}

// ---- forward ------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS) psb_bn_stats(const __nv_bfloat16* __restrict__ x, float* __restrict__ sums,
                                                            BnGeom g) {
  extern __shared__ float smem[];
  const int tx = threadIdx.x % g.groups, ty = threadIdx.x / g.groups;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const long long per_cta = (g.pixels + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per_cta;
  const long long p1 = p0 + per_cta < g.pixels ? p0 + per_cta : g.pixels;
  if (ty < g.lanes) {
    long long p = p0 + ty;
    for (; p + 3LL * g.lanes < p1; p += 4LL * g.lanes) {   // 4 independent 16-byte loads in flight
      float a[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) ld8_stream(x + (p + (long long)u * g.lanes) * g.C + tx * 8, a[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += a[u][j];
          q[j] = fmaf(a[u][j], a[u][j], q[j]);
        }
    }
    for (; p < p1; p += g.lanes) {
      float a[8];
      ld8_stream(x + p * g.C + tx * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += a[j];
        q[j] = fmaf(a[j], a[j], q[j]);
      }
    }
  }
  reduce_lanes_atomic(smem, s, tx, ty, g, sums);
  reduce_lanes_atomic(smem, q, tx, ty, g, sums + g.C);
}

// sums[0:C]=Σx, sums[C:2C]=Σx² → mean/rstd (saved for backward), running stats, fused scale/shift
__global__ void psb_bn_finalize(const float* __restrict__ sums, const __nv_bfloat16* __restrict__ gamma,
                                const __nv_bfloat16* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                                float* __restrict__ scale, float* __restrict__ shift, float* running_mean,
                                float* running_var, int C, long long pixels, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv_n = 1.f / (float)pixels;
  const float m = sums[c] * inv_n;
  float var = fmaf(-m, m, sums[C + c] * inv_n);
  var = fmaxf(var, 0.f);
  const float r = rsqrtf(var + eps);
  mean[c] = m;
  rstd[c] = r;
  const float gsc = gamma ? __bfloat162float(gamma[c]) : 1.f;
  const float b = beta ? __bfloat162float(beta[c]) : 0.f;
  scale[c] = gsc * r;
  shift[c] = fmaf(-m, gsc * r, b);
  if (running_mean) {
    const float unbiased = pixels > 1 ? var * (float)pixels / (float)(pixels - 1) : var;
    running_mean[c] = fmaf(momentum, m - running_mean[c], running_mean[c]);
    running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
  }
}

template <bool RES, bool RELU>
__global__ void __launch_bounds__(BN_THREADS) psb_bn_apply(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ res,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ mask, BnGeom g) {
  // `mask` (RELU, training): one bit per element, y > 0 — the backward reads this 1/16-size tensor instead of y
  const int tx = threadIdx.x % g.groups, ty = threadIdx.x / g.groups;
  if (ty >= g.lanes) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = scale[tx * 8 + j];
    sh[j] = shift[tx * 8 + j];
  }
  const long long stride = (long long)gridDim.x * g.lanes;
  for (long long p = (long long)blockIdx.x * g.lanes + ty; p < g.pixels; p += stride) {
    const long long off = p * g.C + tx * 8;
    float a[8], r[8];
    ld8_stream(x + off, a);
    if (RES) ld8_stream(res + off, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = fmaf(a[j], sc[j], sh[j]);
      if (RES) v += r[j];
      if (RELU) v = fmaxf(v, 0.f);
      a[j] = v;
    }
    st8(y + off, a);
    if (RELU && mask != nullptr) {
      uint32_t m = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) m |= (__bfloat162float(__float2bfloat16_rn(a[j])) > 0.f ? 1u : 0u) << j;
      mask[p * g.groups + tx] = (uint8_t)m;
    }
  }
}

// ---- backward -----------------------------------------------------------------------------
// ReLU mask of 8 consecutive channels: from the 1-bit mask tensor (MASKED) or from the saved output y
template <bool MASKED>
__device__ __forceinline__ uint32_t relu_bits(const __nv_bfloat16* __restrict__ y, const uint8_t* __restrict__ mask, long long p,
                                              long long off, int groups, int tx) {
  if (MASKED) return mask[p * groups + tx];
  float o[8];
  ld8_stream(y + off, o);
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) m |= (o[j] > 0.f ? 1u : 0u) << j;
  return m;
}

template <bool RELU, bool MASKED>
__global__ void __launch_bounds__(BN_THREADS) psb_bn_bwd_reduce(const __nv_bfloat16* __restrict__ dy,
                                                                 const __nv_bfloat16* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ y,
                                                                 const uint8_t* __restrict__ mask,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 float* __restrict__ sums, BnGeom g) {
  extern __shared__ float smem[];
  const int tx = threadIdx.x % g.groups, ty = threadIdx.x / g.groups;
  float s[8], q[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = q[j] = 0.f;
    mu[j] = mean[tx * 8 + j];
    rs[j] = rstd[tx * 8 + j];
  }
  const long long per_cta = (g.pixels + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per_cta;
  const long long p1 = p0 + per_cta < g.pixels ? p0 + per_cta : g.pixels;
  if (ty < g.lanes) {
    long long p = p0 + ty;
    for (; p + (long long)g.lanes < p1; p += 2LL * g.lanes) {
      float d[2][8], a[2][8];
      uint32_t mk[2] = {0xffu, 0xffu};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long pp = p + (long long)u * g.lanes;
        const long long off = pp * g.C + tx * 8;
        ld8_stream(dy + off, d[u]);
        ld8_stream(x + off, a[u]);
        if (RELU) mk[u] = relu_bits<MASKED>(y, mask, pp, off, g.groups, tx);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dd = (RELU && !(mk[u] >> j & 1u)) ? 0.f : d[u][j];
          s[j] += dd;
          q[j] = fmaf(dd, (a[u][j] - mu[j]) * rs[j], q[j]);
        }
    }
    for (; p < p1; p += g.lanes) {
      const long long off = p * g.C + tx * 8;
      float d[8], a[8];
      ld8_stream(dy + off, d);
      ld8_stream(x + off, a);
      const uint32_t mk = RELU ? relu_bits<MASKED>(y, mask, p, off, g.groups, tx) : 0xffu;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (RELU && !(mk >> j & 1u)) ? 0.f : d[j];
        s[j] += dd;
        q[j] = fmaf(dd, (a[j] - mu[j]) * rs[j], q[j]);
      }
    }
  }
  reduce_lanes_atomic(smem, s, tx, ty, g, sums);
  reduce_lanes_atomic(smem, q, tx, ty, g, sums + g.C);
}

// dx = gamma*rstd * (dy' - mean(dy') - x̂ * mean(dy'·x̂))  =  dy'*a + x*b + c   per channel
__global__ void psb_bn_bwd_finalize(const float* __restrict__ sums, const __nv_bfloat16* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ ca,
                                    float* __restrict__ cb, float* __restrict__ cc, __nv_bfloat16* __restrict__ dgamma,
                                    __nv_bfloat16* __restrict__ dbeta, int C, long long pixels) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv_n = 1.f / (float)pixels;
  const float sdy = sums[c], sdyx = sums[C + c];
  const float g = gamma ? __bfloat162float(gamma[c]) : 1.f;
  const float r = rstd[c], m = mean[c];
  const float a = g * r;
  const float k = sdyx * inv_n * r;           // coefficient of (x - m)
  ca[c] = a;
  cb[c] = -a * k;
  cc[c] = a * (k * m - sdy * inv_n);
  if (dgamma) dgamma[c] = __float2bfloat16_rn(sdyx);
  if (dbeta) dbeta[c] = __float2bfloat16_rn(sdy);
}

template <bool RES, bool RELU, bool MASKED>
__global__ void __launch_bounds__(BN_THREADS) psb_bn_bwd_apply(const __nv_bfloat16* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ y,
                                                                const uint8_t* __restrict__ mask, const float* __restrict__ ca,
                                                                const float* __restrict__ cb, const float* __restrict__ cc,
                                                                __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dres,
                                                                BnGeom g) {
  const int tx = threadIdx.x % g.groups, ty = threadIdx.x / g.groups;
  if (ty >= g.lanes) return;
  float a[8], b[8], c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = ca[tx * 8 + j];
    b[j] = cb[tx * 8 + j];
    c[j] = cc[tx * 8 + j];
  }
  const long long stride = (long long)gridDim.x * g.lanes;
  for (long long p = (long long)blockIdx.x * g.lanes + ty; p < g.pixels; p += stride) {
    const long long off = p * g.C + tx * 8;
    float d[8], xv[8];
    ld8_stream(dy + off, d);
    ld8_stream(x + off, xv);
    const uint32_t mk = RELU ? relu_bits<MASKED>(y, mask, p, off, g.groups, tx) : 0xffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (RELU && !(mk >> j & 1u)) d[j] = 0.f;
      xv[j] = fmaf(d[j], a[j], fmaf(xv[j], b[j], c[j]));
    }
    st8(dx + off, xv);
    if (RES) st8(dres + off, d);
  }
}

// (A fused BatchNorm-apply + ReLU + 3x3/s2 max-pool pair lived here in round 2: per-pixel gather, then a 2x2-quad backward.
//  It was correct but never beat the unfused kernels once those got the 1-bit ReLU mask and the quad max-pool backward —
//  forward 0.39 vs 0.49 ms, forward+backward 1.31 vs 1.19 ms at batch 256 — and was deleted; the measurements are in
//  profiles/bnpool_fusion_REJECTED.jsonl.)
BnGeom geom(long long pixels, int C) {
  BnGeom g;
  g.pixels = pixels;
  g.C = C;
  g.groups = C / 8;
  g.lanes = BN_THREADS / g.groups;
  if (g.lanes < 1) g.lanes = 1;
  return g;
}

int grid_for(long long pixels, const BnGeom& g, int min_iters = 1) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (pixels + (long long)g.lanes * min_iters - 1) / ((long long)g.lanes * min_iters);   // >= min_iters pixel batches per CTA
  long long cap = (long long)sms * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}
// The reducing kernels end with 2C float atomics per CTA: on the small late-stage tensors (12 544 pixels x 512 channels) 1 184
// CTAs of ~3 pixel batches each spent their time in 1.2 M contended atomics — every statistics / backward-reduce launch had
// a ~16 us floor whatever its size (profiles/resnet18_step_launches_r2.txt).  They get at least 8 batches per CTA.
constexpr int REDUCE_MIN_ITERS = 8;

}  // namespace

// C must be a multiple of 8 and <= 2048 (groups <= 256)
void psb_bn_forward(cudaStream_t s, const void* x, const void* res, const void* gamma, const void* beta, void* y,
                    float* sums /*2C, zeroed here*/, float* mean, float* rstd, float* scale, float* shift, float* running_mean,
                    float* running_var, long long pixels, int C, float eps, float momentum, int relu, int training, void* mask) {
  auto MK = reinterpret_cast<uint8_t*>(mask);
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto R = reinterpret_cast<const __nv_bfloat16*>(res);
  auto Y = reinterpret_cast<__nv_bfloat16*>(y);
  psb_count_launch(training ? 3 : 1);
  if (training) {
    cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, s);
    psb_bn_stats<<<grid_for(pixels, g, REDUCE_MIN_ITERS), BN_THREADS, sizeof(float) * g.lanes * C, s>>>(X, sums, g);
    psb_bn_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, reinterpret_cast<const __nv_bfloat16*>(gamma),
                                                     reinterpret_cast<const __nv_bfloat16*>(beta), mean, rstd, scale, shift,
                                                     running_mean, running_var, C, pixels, eps, momentum);
  }
  if (res != nullptr) {
    if (relu) psb_bn_apply<true, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
    else psb_bn_apply<true, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
  } else {
    if (relu) psb_bn_apply<false, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
    else psb_bn_apply<false, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
  }
}

// Training forward whose per-channel sums were produced elsewhere (the fused stem kernel's epilogue, stem_kernels.cu):
// finalize + apply only — the statistics pass over x is skipped (the fused stem, default).
void psb_bn_forward_presummed(cudaStream_t s, const void* x, const void* res, const void* gamma, const void* beta, void* y,
                              const float* sums /*2C, filled by the producer*/, float* mean, float* rstd, float* scale,
                              float* shift, float* running_mean, float* running_var, long long pixels, int C, float eps,
                              float momentum, int relu, void* mask) {
  auto MK = reinterpret_cast<uint8_t*>(mask);
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto R = reinterpret_cast<const __nv_bfloat16*>(res);
  auto Y = reinterpret_cast<__nv_bfloat16*>(y);
  psb_count_launch(2);
  psb_bn_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, reinterpret_cast<const __nv_bfloat16*>(gamma),
                                                   reinterpret_cast<const __nv_bfloat16*>(beta), mean, rstd, scale, shift,
                                                   running_mean, running_var, C, pixels, eps, momentum);
  if (res != nullptr) {
    if (relu) psb_bn_apply<true, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
    else psb_bn_apply<true, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
  } else {
    if (relu) psb_bn_apply<false, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
    else psb_bn_apply<false, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, MK, g);
  }
}

void psb_bn_backward(cudaStream_t s, const void* dy, const void* x, const void* y, const void* gamma, const float* mean,
                     const float* rstd, float* sums /*2C*/, float* coef /*3C*/, void* dx, void* dres, void* dgamma, void* dbeta,
                     long long pixels, int C, int relu, const void* mask) {
  auto MK = reinterpret_cast<const uint8_t*>(mask);
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  auto DY = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto Y = reinterpret_cast<const __nv_bfloat16*>(y);
  psb_count_launch(3);
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, s);
  const size_t sm = sizeof(float) * g.lanes * C;
  const int rgrid = grid_for(pixels, g, REDUCE_MIN_ITERS);
  if (relu && MK) psb_bn_bwd_reduce<true, true><<<rgrid, BN_THREADS, sm, s>>>(DY, X, Y, MK, mean, rstd, sums, g);
  else if (relu) psb_bn_bwd_reduce<true, false><<<rgrid, BN_THREADS, sm, s>>>(DY, X, Y, MK, mean, rstd, sums, g);
  else psb_bn_bwd_reduce<false, false><<<rgrid, BN_THREADS, sm, s>>>(DY, X, Y, MK, mean, rstd, sums, g);
  psb_bn_bwd_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, reinterpret_cast<const __nv_bfloat16*>(gamma), mean, rstd, coef,
                                                       coef + C, coef + 2 * C, reinterpret_cast<__nv_bfloat16*>(dgamma),
                                                       reinterpret_cast<__nv_bfloat16*>(dbeta), C, pixels);
  auto DX = reinterpret_cast<__nv_bfloat16*>(dx);
  auto DR = reinterpret_cast<__nv_bfloat16*>(dres);
#define PSB_APPLY(RES_, RELU_, MASKED_) \
  psb_bn_bwd_apply<RES_, RELU_, MASKED_><<<grid, BN_THREADS, 0, s>>>(DY, X, Y, MK, coef, coef + C, coef + 2 * C, DX, DR, g)
  if (dres != nullptr) {
    if (relu && MK) PSB_APPLY(true, true, true);
    else if (relu) PSB_APPLY(true, true, false);
    else PSB_APPLY(true, false, false);
  } else {
    if (relu && MK) PSB_APPLY(false, true, true);
    else if (relu) PSB_APPLY(false, true, false);
    else PSB_APPLY(false, false, false);
  }
#undef PSB_APPLY
}
