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
                                                            __nv_bfloat16* __restrict__ y, BnGeom g) {
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
  }
}

// ---- backward -----------------------------------------------------------------------------
template <bool RELU>
__global__ void __launch_bounds__(BN_THREADS) psb_bn_bwd_reduce(const __nv_bfloat16* __restrict__ dy,
                                                                 const __nv_bfloat16* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ y,
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
      float d[2][8], a[2][8], o[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long off = (p + (long long)u * g.lanes) * g.C + tx * 8;
        ld8_stream(dy + off, d[u]);
        ld8_stream(x + off, a[u]);
        if (RELU) ld8_stream(y + off, o[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dd = (RELU && !(o[u][j] > 0.f)) ? 0.f : d[u][j];
          s[j] += dd;
          q[j] = fmaf(dd, (a[u][j] - mu[j]) * rs[j], q[j]);
        }
    }
    for (; p < p1; p += g.lanes) {
      const long long off = p * g.C + tx * 8;
      float d[8], a[8], o[8];
      ld8_stream(dy + off, d);
      ld8_stream(x + off, a);
      if (RELU) ld8_stream(y + off, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dd = (RELU && !(o[j] > 0.f)) ? 0.f : d[j];
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

template <bool RES, bool RELU>
__global__ void __launch_bounds__(BN_THREADS) psb_bn_bwd_apply(const __nv_bfloat16* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ y, const float* __restrict__ ca,
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
    float d[8], xv[8], o[8];
    ld8_stream(dy + off, d);
    ld8_stream(x + off, xv);
    if (RELU) ld8_stream(y + off, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (RELU && !(o[j] > 0.f)) d[j] = 0.f;
      xv[j] = fmaf(d[j], a[j], fmaf(xv[j], b[j], c[j]));
    }
    st8(dx + off, xv);
    if (RES) st8(dres + off, d);
  }
}

// ==========================================================================================================
// EXPERIMENTAL (opt-in, PSB200_BNPOOL=fused): BatchNorm-apply + ReLU + 3x3/s2/p1 max-pool in ONE pass, forward and
// backward — the ResNet stem tail.  Unfused, the 411 MB BN output (batch 256) is written, read back by the pool, and
// in backward the pool gradient is materialised (411 MB) and read twice together with the BN output (ReLU mask):
//   forward : psb_bnrelu_pool_fwd   reads x once, writes the pooled tensor + 1-byte window positions
//   backward: psb_bnpool_bwd_reduce / psb_bnpool_bwd_apply GATHER dy from (dy_pooled, arg) on the fly and recompute
//             the ReLU mask from x·scale+shift, so neither the BN output nor the un-pooled gradient ever exists.
// Values are rounded to bf16 before the max exactly as the unfused pair does, so the forward is bit-identical.
// ==========================================================================================================
struct BnPoolGeom {
  int N, H, W, C, OH, OW, groups;
};

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__global__ void __launch_bounds__(256) psb_bnrelu_pool_fwd(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                                           const float* __restrict__ shift, __nv_bfloat16* __restrict__ y,
                                                           uint8_t* __restrict__ arg, BnPoolGeom g) {
  const long long total = (long long)g.N * g.OH * g.OW * g.groups;
  int cur = -1;
  float sc[8], sh[8];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % g.groups);
    if (cg != cur) {
      cur = cg;
#pragma unroll
      for (int j = 0; j < 8; ++j) sc[j] = scale[cg * 8 + j], sh[j] = shift[cg * 8 + j];
    }
    long long p = i / g.groups;
    const int ow = (int)(p % g.OW);
    p /= g.OW;
    const int oh = (int)(p % g.OH);
    const int n = (int)(p / g.OH);
    float best[8];
    uint32_t pos[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY, pos[j] = 255;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = oh * 2 - 1 + kh;
      if (h < 0 || h >= g.H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = ow * 2 - 1 + kw;
        if (w < 0 || w >= g.W) continue;
        float v[8];
        ld8(x + (((long long)n * g.H + h) * g.W + w) * g.C + cg * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float z = bf16_round(fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f));
          if (z > best[j]) {             // strictly greater: the first maximum wins ties (ATen's rule)
            best[j] = z;
            pos[j] = kh * 3 + kw;
          }
        }
      }
    }
    const long long o = (((long long)n * g.OH + oh) * g.OW + ow) * g.C + cg * 8;
    st8(y + o, best);
    *reinterpret_cast<uint2*>(arg + o) =
        make_uint2(pos[0] | (pos[1] << 8) | (pos[2] << 16) | (pos[3] << 24), pos[4] | (pos[5] << 8) | (pos[6] << 16) | (pos[7] << 24));
  }
}

// gradient reaching the (never materialised) BN+ReLU output at input pixel (n, h, w), channels cg*8..+7
__device__ __forceinline__ void gather_pool_grad(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ arg,
                                                 const BnPoolGeom& g, int n, int h, int w, int cg, float* acc) {
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int oh0 = h >> 1, ow0 = w >> 1;          // windows covering (h, w): oh*2-1 <= h <= oh*2+1
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int oh = oh0 + a;
    const int kh = h - (oh * 2 - 1);
    if (oh >= g.OH || kh < 0 || kh > 2) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int ow = ow0 + b;
      const int kw = w - (ow * 2 - 1);
      if (ow >= g.OW || kw < 0 || kw > 2) continue;
      const long long o = (((long long)n * g.OH + oh) * g.OW + ow) * g.C + cg * 8;
      const uint2 pr = *reinterpret_cast<const uint2*>(arg + o);
      float d[8];
      ld8(dy + o, d);
      const uint32_t want = (uint32_t)(kh * 3 + kw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t pj = ((j < 4 ? pr.x : pr.y) >> (8 * (j & 3))) & 0xffu;
        if (pj == want) acc[j] += d[j];
      }
    }
  }
}

// Σ dy' and Σ dy'·x̂ over all input pixels (dy' = pooled gradient gathered back, masked by the recomputed ReLU)
__global__ void __launch_bounds__(BN_THREADS) psb_bnpool_bwd_reduce(const __nv_bfloat16* __restrict__ dy,
                                                                     const uint8_t* __restrict__ arg,
                                                                     const __nv_bfloat16* __restrict__ x,
                                                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                     float* __restrict__ sums, BnGeom g, BnPoolGeom pg) {
  extern __shared__ float smem[];
  const int tx = threadIdx.x % g.groups, ty = threadIdx.x / g.groups;
  float s[8], q[8], mu[8], rs[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = q[j] = 0.f;
    mu[j] = mean[tx * 8 + j];
    rs[j] = rstd[tx * 8 + j];
    sc[j] = scale[tx * 8 + j];
    sh[j] = shift[tx * 8 + j];
  }
  const long long per_cta = (g.pixels + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per_cta;
  const long long p1 = p0 + per_cta < g.pixels ? p0 + per_cta : g.pixels;
  if (ty < g.lanes) {
    for (long long p = p0 + ty; p < p1; p += g.lanes) {
      const int w = (int)(p % pg.W);
      const long long t = p / pg.W;
      const int h = (int)(t % pg.H), n = (int)(t / pg.H);
      float d[8], a[8];
      gather_pool_grad(dy, arg, pg, n, h, w, tx, d);
      ld8_stream(x + p * g.C + tx * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float o = bf16_round(fmaxf(fmaf(a[j], sc[j], sh[j]), 0.f));
        const float dd = o > 0.f ? d[j] : 0.f;
        s[j] += dd;
        q[j] = fmaf(dd, (a[j] - mu[j]) * rs[j], q[j]);
      }
    }
  }
  reduce_lanes_atomic(smem, s, tx, ty, g, sums);
  reduce_lanes_atomic(smem, q, tx, ty, g, sums + g.C);
}

__global__ void __launch_bounds__(BN_THREADS) psb_bnpool_bwd_apply(const __nv_bfloat16* __restrict__ dy,
                                                                    const uint8_t* __restrict__ arg,
                                                                    const __nv_bfloat16* __restrict__ x,
                                                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                                                    const float* __restrict__ ca, const float* __restrict__ cb,
                                                                    const float* __restrict__ cc, __nv_bfloat16* __restrict__ dx,
                                                                    BnGeom g, BnPoolGeom pg) {
  const int tx = threadIdx.x % g.groups, ty = threadIdx.x / g.groups;
  if (ty >= g.lanes) return;
  float a[8], b[8], c[8], sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = ca[tx * 8 + j];
    b[j] = cb[tx * 8 + j];
    c[j] = cc[tx * 8 + j];
    sc[j] = scale[tx * 8 + j];
    sh[j] = shift[tx * 8 + j];
  }
  const long long stride = (long long)gridDim.x * g.lanes;
  for (long long p = (long long)blockIdx.x * g.lanes + ty; p < g.pixels; p += stride) {
    const int w = (int)(p % pg.W);
    const long long t = p / pg.W;
    const int h = (int)(t % pg.H), n = (int)(t / pg.H);
    float d[8], xv[8];
    gather_pool_grad(dy, arg, pg, n, h, w, tx, d);
    const long long off = p * g.C + tx * 8;
    ld8_stream(x + off, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float o = bf16_round(fmaxf(fmaf(xv[j], sc[j], sh[j]), 0.f));
      const float dd = o > 0.f ? d[j] : 0.f;
      xv[j] = fmaf(dd, a[j], fmaf(xv[j], b[j], c[j]));
    }
    st8(dx + off, xv);
  }
}

BnGeom geom(long long pixels, int C) {
  BnGeom g;
  g.pixels = pixels;
  g.C = C;
  g.groups = C / 8;
  g.lanes = BN_THREADS / g.groups;
  if (g.lanes < 1) g.lanes = 1;
  return g;
}

int grid_for(long long pixels, const BnGeom& g) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (pixels + g.lanes - 1) / g.lanes;        // one pixel batch per CTA at least
  long long cap = (long long)sms * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

// C must be a multiple of 8 and <= 2048 (groups <= 256)
void psb_bn_forward(cudaStream_t s, const void* x, const void* res, const void* gamma, const void* beta, void* y,
                    float* sums /*2C, zeroed here*/, float* mean, float* rstd, float* scale, float* shift, float* running_mean,
                    float* running_var, long long pixels, int C, float eps, float momentum, int relu, int training) {
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto R = reinterpret_cast<const __nv_bfloat16*>(res);
  auto Y = reinterpret_cast<__nv_bfloat16*>(y);
  psb_count_launch(training ? 3 : 1);
  if (training) {
    cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, s);
    psb_bn_stats<<<grid, BN_THREADS, sizeof(float) * g.lanes * C, s>>>(X, sums, g);
    psb_bn_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, reinterpret_cast<const __nv_bfloat16*>(gamma),
                                                     reinterpret_cast<const __nv_bfloat16*>(beta), mean, rstd, scale, shift,
                                                     running_mean, running_var, C, pixels, eps, momentum);
  }
  if (res != nullptr) {
    if (relu) psb_bn_apply<true, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
    else psb_bn_apply<true, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
  } else {
    if (relu) psb_bn_apply<false, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
    else psb_bn_apply<false, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
  }
}

// Training forward whose per-channel sums were produced elsewhere (the fused stem kernel's epilogue, stem_kernels.cu):
// finalize + apply only — the statistics pass over x is skipped.  EXPERIMENTAL (PSB200_STEM=fused).
void psb_bn_forward_presummed(cudaStream_t s, const void* x, const void* res, const void* gamma, const void* beta, void* y,
                              const float* sums /*2C, filled by the producer*/, float* mean, float* rstd, float* scale,
                              float* shift, float* running_mean, float* running_var, long long pixels, int C, float eps,
                              float momentum, int relu) {
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
    if (relu) psb_bn_apply<true, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
    else psb_bn_apply<true, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
  } else {
    if (relu) psb_bn_apply<false, true><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
    else psb_bn_apply<false, false><<<grid, BN_THREADS, 0, s>>>(X, R, scale, shift, Y, g);
  }
}

void psb_bn_backward(cudaStream_t s, const void* dy, const void* x, const void* y, const void* gamma, const float* mean,
                     const float* rstd, float* sums /*2C*/, float* coef /*3C*/, void* dx, void* dres, void* dgamma, void* dbeta,
                     long long pixels, int C, int relu) {
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  auto DY = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto Y = reinterpret_cast<const __nv_bfloat16*>(y);
  psb_count_launch(3);
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, s);
  const size_t sm = sizeof(float) * g.lanes * C;
  if (relu) psb_bn_bwd_reduce<true><<<grid, BN_THREADS, sm, s>>>(DY, X, Y, mean, rstd, sums, g);
  else psb_bn_bwd_reduce<false><<<grid, BN_THREADS, sm, s>>>(DY, X, Y, mean, rstd, sums, g);
  psb_bn_bwd_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, reinterpret_cast<const __nv_bfloat16*>(gamma), mean, rstd, coef,
                                                       coef + C, coef + 2 * C, reinterpret_cast<__nv_bfloat16*>(dgamma),
                                                       reinterpret_cast<__nv_bfloat16*>(dbeta), C, pixels);
  auto DX = reinterpret_cast<__nv_bfloat16*>(dx);
  auto DR = reinterpret_cast<__nv_bfloat16*>(dres);
  if (dres != nullptr) {
    if (relu) psb_bn_bwd_apply<true, true><<<grid, BN_THREADS, 0, s>>>(DY, X, Y, coef, coef + C, coef + 2 * C, DX, DR, g);
    else psb_bn_bwd_apply<true, false><<<grid, BN_THREADS, 0, s>>>(DY, X, Y, coef, coef + C, coef + 2 * C, DX, DR, g);
  } else {
    if (relu) psb_bn_bwd_apply<false, true><<<grid, BN_THREADS, 0, s>>>(DY, X, Y, coef, coef + C, coef + 2 * C, DX, DR, g);
    else psb_bn_bwd_apply<false, false><<<grid, BN_THREADS, 0, s>>>(DY, X, Y, coef, coef + C, coef + 2 * C, DX, DR, g);
  }
}

// ---- EXPERIMENTAL fused BN + ReLU + 3x3/s2/p1 max-pool (training) -------------------------------------------------
// `sums_in` != nullptr: Σx / Σx² were produced elsewhere (fused stem epilogue) and the statistics pass is skipped.
void psb_bnpool_forward(cudaStream_t s, const void* x, const void* gamma, const void* beta, void* y, void* arg, float* sums,
                        const float* sums_in, float* mean, float* rstd, float* scale, float* shift, float* running_mean,
                        float* running_var, int N, int H, int W, int C, float eps, float momentum) {
  const long long pixels = (long long)N * H * W;
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  psb_count_launch(sums_in ? 2 : 3);
  if (sums_in == nullptr) {
    cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, s);
    psb_bn_stats<<<grid, BN_THREADS, sizeof(float) * g.lanes * C, s>>>(X, sums, g);
  }
  psb_bn_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums_in ? sums_in : sums, reinterpret_cast<const __nv_bfloat16*>(gamma),
                                                   reinterpret_cast<const __nv_bfloat16*>(beta), mean, rstd, scale, shift,
                                                   running_mean, running_var, C, pixels, eps, momentum);
  BnPoolGeom pg{N, H, W, C, (H + 2 - 3) / 2 + 1, (W + 2 - 3) / 2 + 1, C / 8};
  const long long total = (long long)N * pg.OH * pg.OW * pg.groups;
  long long want = (total + 255) / 256, cap = (long long)grid_for(pixels, g) * 2;
  psb_bnrelu_pool_fwd<<<(int)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, s>>>(
      X, scale, shift, reinterpret_cast<__nv_bfloat16*>(y), reinterpret_cast<uint8_t*>(arg), pg);
}

void psb_bnpool_backward(cudaStream_t s, const void* dy, const void* arg, const void* x, const void* gamma, const float* mean,
                         const float* rstd, const float* scale, const float* shift, float* sums /*2C*/, float* coef /*3C*/,
                         void* dx, void* dgamma, void* dbeta, int N, int H, int W, int C) {
  const long long pixels = (long long)N * H * W;
  const BnGeom g = geom(pixels, C);
  const int grid = grid_for(pixels, g);
  BnPoolGeom pg{N, H, W, C, (H + 2 - 3) / 2 + 1, (W + 2 - 3) / 2 + 1, C / 8};
  auto DY = reinterpret_cast<const __nv_bfloat16*>(dy);
  auto A = reinterpret_cast<const uint8_t*>(arg);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  psb_count_launch(3);
  cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, s);
  psb_bnpool_bwd_reduce<<<grid, BN_THREADS, sizeof(float) * g.lanes * C, s>>>(DY, A, X, scale, shift, mean, rstd, sums, g, pg);
  psb_bn_bwd_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, reinterpret_cast<const __nv_bfloat16*>(gamma), mean, rstd, coef,
                                                       coef + C, coef + 2 * C, reinterpret_cast<__nv_bfloat16*>(dgamma),
                                                       reinterpret_cast<__nv_bfloat16*>(dbeta), C, pixels);
  psb_bnpool_bwd_apply<<<grid, BN_THREADS, 0, s>>>(DY, A, X, scale, shift, coef, coef + C, coef + 2 * C,
                                                   reinterpret_cast<__nv_bfloat16*>(dx), g, pg);
}
