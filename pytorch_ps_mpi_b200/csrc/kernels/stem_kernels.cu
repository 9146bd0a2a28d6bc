// The ResNet stem (default path since round 2; PSB200_STEM=im2col falls back) — 7x7 / stride 2 / pad 3 convolution, 3 → 64 channels —
// as ONE implicit-GEMM kernel on tcgen05, replacing psb_im2col_stem (writes a 1.13 GB patch matrix at batch 256) +
// psb_bcast_gemm2_kernel (reads it back) + the BatchNorm statistics pass over the 411 MB output
// (profiles/resnet18_step_launches_final.txt: 355 + 505 + ~95 us of a 7.9 ms step).
//
// One tile = one output row (n, oh): OW <= 128 pixels x 64 channels, K = 176 (7 kernel rows x 24 columns — 21 real,
// 3 zero — + 8 zero columns: the layout of ops/stem.py, so the same [64,176] weight matrix is used).
//
//   warps 0-3  builders   cp.async the 7 input rows of the NEXT tile into a zero-margined smem patch while building
//                         the CURRENT tile's A operand: thread m copies, for each kernel row, the 42 contiguous
//                         bytes x[n, 2oh-3+kh, 2m-3 .. 2m+3, 0..2] into the 128B-swizzled K-major UMMA layout
//                         (three [128 x 64] bf16 blocks; the same canonical layout TMA would have produced)
//   warp  8    MMA        11 x tcgen05.mma.cta_group::1.kind::f16 (128 x 64 x 16) per tile, accumulators
//                         double-buffered in TMEM; the [64,176] weight operand is TMA-loaded once per CTA
//   warps 4-7  epilogue   tcgen05.ld → bf16 → swizzled staging tile → ONE cp.async.bulk.tensor store per output row
//                         (NHWC: the row is 14 KB contiguous), and the per-channel Σy / Σy² of the staged bf16 values
//                         (exactly what psb_bn_stats would read back) accumulated in registers, 2 x 64 atomics per CTA
//
// Expected (desk estimate, to be measured): ~1000-1500 cycles per tile per SM → 0.15-0.25 ms for batch 256
// versus ~0.95 ms for the three passes it replaces.
#include "gemm_common.cuh"

namespace {

constexpr int SK = 176;                 // GEMM K (ops/stem.py STEM_K)
constexpr int S_THREADS = 288;          // 4 builder + 4 epilogue + 1 MMA warp
constexpr int SA_BLK = 128 * 128;       // one k-block of A: 128 rows x 128 B
constexpr int SA_BYTES = 3 * SA_BLK;    // 48 KB per A buffer
constexpr int SB_BLK = 64 * 128;        // one k-block of B (64 output channels)
constexpr int SB_BYTES = 3 * SB_BLK;    // 24 KB
constexpr int SO_BYTES = 128 * 128;     // staging tile: 128 rows x 64 bf16
constexpr int MARGIN = 48;              // zero bytes left and right of a patch row (>= 18 needed, 16-byte multiple)
constexpr int MAX_W = 256;
constexpr int PATCH_PITCH_MAX = MARGIN + MAX_W * 6 + MARGIN;
constexpr int PATCH_BYTES = 7 * PATCH_PITCH_MAX;             // 11 424
constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + 2 * SA_BYTES;                   // 98 304
constexpr int OFF_O = OFF_B + SB_BYTES;                       // 122 880 (1024-aligned)
constexpr int OFF_P = OFF_O + 2 * SO_BYTES;                   // 155 648
constexpr int OFF_BAR = OFF_P + 2 * PATCH_BYTES;              // 178 496
constexpr int STEM_SMEM = OFF_BAR + 256 + 1024;               // + barriers + alignment slack
static_assert(OFF_O % 1024 == 0 && OFF_B % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
static_assert(STEM_SMEM <= 232448, "shared memory budget");

struct StemParams {
  const __nv_bfloat16* x;     // [N, H, W, 3] bf16 (channels-last, already normalised)
  float* sums;                // [128]: Σy[64] | Σy²[64], accumulated with atomics (caller zeroes); nullable
  int N, H, W, OH, OW;
  // The broadcast gate (as in bcast_gemm*.cu): the weight matrix lives IN the symmetric parameter arena in the [64,176]
  // GEMM layout (layout.py custom placement) and the lane that TMA-loads it first acquires the PS's PARAMS_READY epoch, so
  // this kernel — the first consumer of parameters in a ResNet forward — is the req.Wait() of the broadcast
  // (/root/reference/mpi_comms.py:120-124) and workers queue no separate wait kernel.
  const uint64_t* ready_flag; // nullable
  uint64_t ready_epoch;
  uint64_t* err_slot;
  unsigned long long timeout_ns;
};

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__host__ __device__ constexpr uint32_t stem_idesc() {   // D=f32, A=B=bf16, K-major, M=128, N=64
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// queue the 7 input rows of output row `tile` into `patch` (cp.async; rows outside the image are zero-filled)
__device__ __forceinline__ void load_patch(const StemParams& p, int tile, uint32_t patch, int pitch, int bt) {
  const int n = tile / p.OH, oh = tile - n * p.OH;
  const int chunks = p.W * 6 / 16;                                    // 16-byte pieces per input row
  for (int i = bt; i < 7 * chunks; i += 128) {
    const int kh = i / chunks, c = i - kh * chunks;
    const int ih = 2 * oh - 3 + kh;
    const uint32_t dst = patch + kh * pitch + MARGIN + c * 16;
    if (ih >= 0 && ih < p.H) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.x) + ((size_t)(n * p.H + ih) * p.W) * 6 + (size_t)c * 16;
      cp_async16(dst, src);
    } else {
      sts_v4(dst, 0u, 0u, 0u, 0u);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// A-operand row `bt` (output pixel ow = bt) of one tile: for each kernel row the 42 contiguous patch bytes
// x[ih, 2ow-3 .. 2ow+3, 0..2] → three 16-byte chunks of the 128B-swizzled [128 x 64] k-blocks.  The same tile serves
// the forward (K-major A operand) and the weight gradient (MN-major operand: rows are then the reduction dimension).
__device__ __forceinline__ void build_row(uint32_t patch, uint32_t a_tile, int pitch, int bt) {
  const uint32_t prow = patch + 30 + 12 * bt;                               // MARGIN + (2*ow - 3) * 6 bytes
  const uint32_t arow = a_tile + (bt >> 3) * 1024 + (bt & 7) * 128;
#pragma unroll
  for (int kh = 0; kh < 7; ++kh) {
    const uint32_t q = prow + kh * pitch;              // 2-mod-4 aligned: one 2-byte and ten 4-byte loads
    const uint32_t first = lds_u16(q);
    uint32_t w4[10], v[12];
#pragma unroll
    for (int j = 0; j < 10; ++j) w4[j] = lds_u32(q + 2 + 4 * j);
    v[0] = first | (w4[0] << 16);
#pragma unroll
    for (int j = 1; j < 10; ++j) v[j] = (w4[j - 1] >> 16) | (w4[j] << 16);
    v[10] = w4[9] >> 16;
    v[11] = 0u;
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
      const int qc = kh * 3 + c3;                      // 16-byte chunk index within the 352-byte logical row
      sts_v4(arow + (qc >> 3) * SA_BLK + (((qc & 7) ^ (bt & 7)) << 4), v[4 * c3], v[4 * c3 + 1], v[4 * c3 + 2],
             v[4 * c3 + 3]);
    }
  }
  // chunk 21 (columns 168..175) stays zero from the initial clear
}

__global__ void __launch_bounds__(S_THREADS, 1)
psb_stem_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_y,
                    const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* b_full = bars;            // [1]  weights landed
  uint64_t* a_full = bars + 1;        // [2]  4 builder warps arrive
  uint64_t* a_empty = bars + 3;       // [2]  tcgen05.commit
  uint64_t* t_full = bars + 5;        // [2]  tcgen05.commit (accumulator ready)
  uint64_t* t_empty = bars + 7;       // [2]  4 epilogue warps arrive
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.N * p.OH;
  const int per = (tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(t0 + per, tiles);
  const int pitch = MARGIN + p.W * 6 + MARGIN;
  const uint32_t sA = smem_u32(smem + OFF_A), sB = smem_u32(smem + OFF_B), sO = smem_u32(smem + OFF_O),
                 sP = smem_u32(smem + OFF_P);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_y) : "memory");
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&a_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero both A buffers (rows >= OW and the K padding are never written again) and the patch margins
  for (int i = threadIdx.x; i < (2 * SA_BYTES) / 16; i += S_THREADS) sts_v4(sA + i * 16, 0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < (2 * PATCH_BYTES) / 16; i += S_THREADS) sts_v4(sP + i * 16, 0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the zero rows / columns of A are read by the tensor core
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    // ============================== builders ==============================
    const int bt = threadIdx.x;                        // 0..127 == the A row (output pixel ow) this thread fills
    if (t0 < t1) load_patch(p, t0, sP, pitch, bt);
    for (int t = t0, i = 0; t < t1; ++t, ++i) {
      const int buf = i & 1;
      if (t + 1 < t1) {
        load_patch(p, t + 1, sP + (buf ^ 1) * PATCH_BYTES, pitch, bt);      // patch[buf^1]: last read one tile ago (barrier below)
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      named_bar(1, 128);                               // everybody's copies of patch[buf] have landed
      mbar_wait(&a_empty[buf], ((i >> 1) & 1) ^ 1);    // the MMAs that read A[buf] two tiles ago are done
      if (bt < p.OW) build_row(sP + buf * PATCH_BYTES, sA + buf * SA_BYTES, pitch, bt);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic writes → visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
      named_bar(1, 128);                               // everybody finished READING patch[buf] (refilled during the next tile)
    }
  } else if (warp == 8) {
    // ============================== MMA issuer ==============================
    if (elect_one()) {
      if (p.ready_flag != nullptr) {       // patch loads / A-tile builds of the first tiles already run in the other warps
        psb::spin_until_ge(p.ready_flag, p.ready_epoch, p.err_slot, p.timeout_ns);
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire → async-proxy (TMA) reads
      }
      mbar_expect_tx(b_full, SB_BYTES);
      for (int b = 0; b < 3; ++b) tma_load_2d(&tmap_w, b_full, smem + OFF_B + b * SB_BLK, b * 64, 0);
    }
    __syncwarp();
    mbar_wait(b_full, 0);
    const uint32_t idesc = stem_idesc();
    for (int t = t0, i = 0; t < t1; ++t, ++i) {
      const int buf = i & 1;
      const uint32_t par = (i >> 1) & 1;
      mbar_wait(&t_empty[buf], par ^ 1);
      mbar_wait(&a_full[buf], par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + buf * 64;
#pragma unroll
        for (int kk = 0; kk < SK / UMMA_K; ++kk) {
          const uint64_t da = make_desc(sA + buf * SA_BYTES + (kk >> 2) * SA_BLK) + (uint64_t)(2 * (kk & 3));
          const uint64_t db = make_desc(sB + (kk >> 2) * SB_BLK) + (uint64_t)(2 * (kk & 3));
          umma(d_tmem, da, db, idesc, kk != 0);
        }
        umma_commit(&a_empty[buf]);
        umma_commit(&t_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ============================== epilogue ==============================
    const int quarter = warp & 3;                      // warps 4..7 → TMEM lane quarters 0..3
    const int et = (warp - 4) * 32 + lane;             // 0..127
    const int row = quarter * 32 + lane;               // == et
    const int ch = et & 63, half = et >> 6;
    const int rhalf = (p.OW + 1) >> 1;
    const int r0 = half * rhalf, r1 = min(p.OW, r0 + rhalf);
    float s_sum = 0.f, s_sq = 0.f;
    for (int t = t0, i = 0; t < t1; ++t, ++i) {
      const int buf = i & 1;
      mbar_wait(&t_full[buf], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32], pk[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * 64;
      tmem_ld_32x32b_x32(taddr, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = psb::pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      tmem_ld_32x32b_x32(taddr + 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[16 + j] = psb::pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[buf]);       // the accumulator is free again: the MMAs of tile i+2 may start

      // staging buffer `buf`: the bulk store issued from it two tiles ago must have finished reading it
      if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      named_bar(2, 128);
      const uint32_t orow = sO + buf * SO_BYTES + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) sts_v4(orow + ((j ^ (row & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      named_bar(2, 128);
      if (et == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_y),
                     "r"(sO + buf * SO_BYTES), "r"(0), "r"(t * p.OW)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (p.sums != nullptr) {
        // column sums of the bf16 values just staged (what a separate statistics pass would read back)
        const uint32_t cbase = sO + buf * SO_BYTES + (ch & 7) * 2;
        for (int rr = r0; rr < r1; ++rr) {
          const float v = __uint_as_float(lds_u16(cbase + rr * 128 + (((ch >> 3) ^ (rr & 7)) << 4)) << 16);
          s_sum += v;
          s_sq = fmaf(v, v, s_sq);
        }
      }
    }
    if (p.sums != nullptr && t0 < t1) {
      atomicAdd(p.sums + ch, s_sum);
      atomicAdd(p.sums + 64 + ch, s_sq);
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ==========================================================================================================
// Weight gradient of the stem, implicit too:  dW2d[co][k] = Σ_pixels gy[p][co] · A[p][k].
// The reduction runs over PIXELS, so both operands are "MN-major" for the tensor core — which is exactly how they
// already sit in shared memory: the im2col tile built by build_row ([pixel rows][k], k contiguous) is the M-side
// operand, the TMA-loaded gy tile ([pixel rows][64 channels]) the N-side one.  UMMA M = 128 covers k-blocks {0,1};
// a second MMA over k-blocks {1,2} provides k = 128..175 (rows 64..111 of its accumulator).  Both accumulators stay
// in TMEM for the CTA's whole life; each CTA writes one fp32 partial [176][64] and the host sums the partials
// (deterministic, 148 x 45 KB).
//   warps 0-3  builders (as in the forward), then the final TMEM → global epilogue
//   warp  4    TMA producer of the gy tiles     warp 5   MMA issuer
// ==========================================================================================================
constexpr int WG_THREADS = 192;
constexpr int SG_BYTES = 128 * 128;                              // gy tile: up to 128 pixel rows x 64 bf16
constexpr int WOFF_A = 0;
constexpr int WOFF_G = WOFF_A + 2 * SA_BYTES;                    // 98 304
constexpr int WOFF_P = WOFF_G + 2 * SG_BYTES;                    // 131 072
constexpr int WOFF_BAR = WOFF_P + 2 * PATCH_BYTES;
constexpr int WGRAD_SMEM = WOFF_BAR + 256 + 1024;
static_assert(WOFF_G % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
static_assert(WGRAD_SMEM <= 232448, "shared memory budget");

struct StemWgradParams {
  const __nv_bfloat16* x;     // [N, H, W, 3] bf16
  float* partial;             // [gridDim.x][176][64] fp32
  int N, H, W, OH, OW;
};

// MN-major, SWIZZLE_128B smem descriptor: 64-element MN blocks `lbo` bytes apart, 8-row K groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t stem_idesc_mn() {   // as stem_idesc, both operands MN-major (bits 15 / 16)
  return stem_idesc() | (1u << 15) | (1u << 16);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
psb_stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ StemWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WOFF_BAR);
  uint64_t* a_full = bars;            // [2]  4 builder warps arrive
  uint64_t* g_full = bars + 2;        // [2]  TMA bytes of the gy tile
  uint64_t* empty = bars + 4;         // [2]  tcgen05.commit: A[buf] and G[buf] are free again
  uint64_t* d_full = bars + 6;        // [1]  every MMA of this CTA has completed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.N * p.OH;
  const int per = (tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(t0 + per, tiles);
  const int pitch = MARGIN + p.W * 6 + MARGIN;
  const uint32_t sA = smem_u32(smem + WOFF_A), sG = smem_u32(smem + WOFF_G), sP = smem_u32(smem + WOFF_P);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_g) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&g_full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(d_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows >= OW of A and of the gy tiles take part in the last K step: they must be zero, not stale bits
  for (int i = threadIdx.x; i < (2 * SA_BYTES + 2 * SG_BYTES) / 16; i += WG_THREADS) sts_v4(sA + i * 16, 0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < (2 * PATCH_BYTES) / 16; i += WG_THREADS) sts_v4(sP + i * 16, 0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;
  StemParams lp{};                                     // load_patch only reads x / H / W / OH
  lp.x = p.x, lp.N = p.N, lp.H = p.H, lp.W = p.W, lp.OH = p.OH, lp.OW = p.OW;

  if (warp < 4) {
    // ============================== builders ==============================
    const int bt = threadIdx.x;
    if (t0 < t1) load_patch(lp, t0, sP, pitch, bt);
    for (int t = t0, i = 0; t < t1; ++t, ++i) {
      const int buf = i & 1;
      if (t + 1 < t1) {
        load_patch(lp, t + 1, sP + (buf ^ 1) * PATCH_BYTES, pitch, bt);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      named_bar(1, 128);
      mbar_wait(&empty[buf], ((i >> 1) & 1) ^ 1);
      if (bt < p.OW) build_row(sP + buf * PATCH_BYTES, sA + buf * SA_BYTES, pitch, bt);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[buf]);
      named_bar(1, 128);
    }
    // ---- final epilogue: this CTA's partial dW2d^T [176][64] ----
    float* out = p.partial + (size_t)blockIdx.x * SK * 64;
    const int r = warp * 32 + lane;                    // TMEM lane == accumulator row
    if (t0 < t1) {
      mbar_wait(d_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {           // accumulator 0: k = r;  accumulator 1: k = 64 + r (r >= 64 only)
        const int k = part == 0 ? r : 64 + r;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + part * 64 + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if ((part == 0 || r >= 64) && k < SK) {
            float4* o = reinterpret_cast<float4*>(out + (size_t)k * 64 + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
          }
        }
      }
    } else {
      for (int k = r; k < SK; k += 128)
        for (int c = 0; c < 64; ++c) out[(size_t)k * 64 + c] = 0.f;
    }
  } else if (warp == 4) {
    // ============================== gy TMA producer ==============================
    if (elect_one()) {
      for (int t = t0, i = 0; t < t1; ++t, ++i) {
        const int buf = i & 1;
        mbar_wait(&empty[buf], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&g_full[buf], (uint32_t)p.OW * 128u);
        tma_load_2d(&tmap_g, &g_full[buf], smem + WOFF_G + buf * SG_BYTES, 0, t * p.OW);
      }
    }
  } else {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = stem_idesc_mn();
    const int ksteps = (p.OW + 15) >> 4;               // 16 pixel rows per UMMA K step (rows >= OW are zero)
    for (int t = t0, i = 0; t < t1; ++t, ++i) {
      const int buf = i & 1;
      const uint32_t par = (i >> 1) & 1;
      mbar_wait(&a_full[buf], par);
      mbar_wait(&g_full[buf], par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t a0 = sA + buf * SA_BYTES + ks * 2048;
          const uint64_t dg = make_desc_mn(sG + buf * SG_BYTES + ks * 2048, 16);
          umma(tmem_base, make_desc_mn(a0, SA_BLK), dg, idesc, (i | ks) != 0);                 // k-blocks 0,1 → k 0..127
          umma(tmem_base + 64, make_desc_mn(a0 + SA_BLK, SA_BLK), dg, idesc, (i | ks) != 0);   // k-blocks 1,2 → k 64..191
        }
        umma_commit(&empty[buf]);
        if (t + 1 == t1) umma_commit(d_full);
      }
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// dW2d[co][k] = Σ_cta partial[cta][k][co]  (fp32 sum in CTA order → deterministic), cast to bf16.
// Thread t handles (k = t / 64, co = t % 64): the reads of one warp are 128 contiguous bytes of each partial.
__global__ void __launch_bounds__(256) psb_stem_wgrad_finalize_kernel(const float* __restrict__ partial, int grid,
                                                                      __nv_bfloat16* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= SK * 64) return;
  const int k = t >> 6, co = t & 63;
  float acc = 0.f;
  int g = 0;
  for (; g + 4 <= grid; g += 4) {
    const float a0 = partial[(size_t)(g + 0) * SK * 64 + t], a1 = partial[(size_t)(g + 1) * SK * 64 + t],
                a2 = partial[(size_t)(g + 2) * SK * 64 + t], a3 = partial[(size_t)(g + 3) * SK * 64 + t];
    acc += a0;
    acc += a1;
    acc += a2;
    acc += a3;
  }
  for (; g < grid; ++g) acc += partial[(size_t)g * SK * 64 + t];
  out[co * SK + k] = __float2bfloat16_rn(acc);
}

}  // namespace

void psb_stem_wgrad_finalize_launch(cudaStream_t s, const float* partial, int grid, void* out_bf16) {
  psb_count_launch(1);
  psb_stem_wgrad_finalize_kernel<<<(SK * 64 + 255) / 256, 256, 0, s>>>(partial, grid, reinterpret_cast<__nv_bfloat16*>(out_bf16));
}

int psb_stem_fwd_smem_bytes() { return STEM_SMEM; }

// tmap_w: [64,176] bf16 weights, box 64 x 64, SWIZZLE_128B.  tmap_y: [N*OH*OW, 64] bf16 output, box 64 columns x OW rows,
// SWIZZLE_128B.  `sums` (nullable) must hold 128 zeroed floats.
void psb_stem_fwd_launch(cudaStream_t s, const void* tmap_w, const void* tmap_y, const void* x, float* sums, int N, int H, int W,
                         int num_sms, const uint64_t* ready_flag, uint64_t ready_epoch, unsigned long long timeout_ns) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(psb_stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STEM_SMEM);
    configured = true;
  }
  StemParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.sums = sums;
  p.N = N, p.H = H, p.W = W;
  p.OH = (H - 1) / 2 + 1, p.OW = (W - 1) / 2 + 1;
  p.ready_flag = ready_flag;
  p.ready_epoch = ready_epoch;
  p.err_slot = ready_flag != nullptr ? const_cast<uint64_t*>(ready_flag) - SIG_PARAMS_READY + SIG_ERROR : nullptr;
  p.timeout_ns = timeout_ns;
  const int tiles = N * p.OH;
  const int grid = tiles < num_sms ? tiles : num_sms;
  psb_count_launch(1);
  psb_stem_fwd_kernel<<<grid, S_THREADS, STEM_SMEM, s>>>(*reinterpret_cast<const CUtensorMap*>(tmap_w),
                                                        *reinterpret_cast<const CUtensorMap*>(tmap_y), p);
}

// tmap_g: [N*OH*OW, 64] bf16 output gradient (channels-last), box 64 columns x OW rows, SWIZZLE_128B.
// `partial`: [grid][176][64] fp32, fully written (no zeroing needed); returns the grid size through *grid_out.
int psb_stem_wgrad_grid(int N, int H, int num_sms) {
  const int tiles = N * ((H - 1) / 2 + 1);
  return tiles < num_sms ? tiles : num_sms;
}
void psb_stem_wgrad_launch(cudaStream_t s, const void* tmap_g, const void* x, float* partial, int N, int H, int W, int num_sms) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(psb_stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WGRAD_SMEM);
    configured = true;
  }
  StemWgradParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.partial = partial;
  p.N = N, p.H = H, p.W = W;
  p.OH = (H - 1) / 2 + 1, p.OW = (W - 1) / 2 + 1;
  psb_count_launch(1);
  psb_stem_wgrad_kernel<<<psb_stem_wgrad_grid(N, H, num_sms), WG_THREADS, WGRAD_SMEM, s>>>(
      *reinterpret_cast<const CUtensorMap*>(tmap_g), p);
}
