// psb_bcast_gemm2_kernel — the cta_group::2 tcgen05 GEMM (the production kernel behind bcast_linear / BcastLinear for
// M >= 256): Y[M,N] = act(X[M,K] · W[N,K]^T + bias), bf16 in, fp32 accumulate in TMEM, bf16 out.
//
// A thread-block cluster of two CTAs (one TPC) works on ONE 256 x BNT output tile:
//   * each CTA TMA-loads its own 128 rows of A and its own HALF of the B tile and signals the LEADER's full barrier
//     (cp.async.bulk.tensor ... .cta_group::2, peer-masked mbarrier address);
//   * only the leader issues tcgen05.mma.cta_group::2 (UMMA 256 x BNT x 16): each B byte is fetched from L2 once per PAIR;
//   * tcgen05.commit ... .multicast::cluster releases the smem slot / publishes the accumulator in BOTH CTAs; each CTA's
//     epilogue warps drain their own 128 TMEM lanes and arrive on the leader's tmem_empty.
// The weight tensor map points INTO the symmetric parameter arena and the TMA producer acquires the PARAMS_READY epoch
// (ld.acquire.sys + fence.proxy.async) before its first weight load: the GEMM is the req.Wait() of the PS broadcast
// (/root/reference/mpi_comms.py:120-133).
//
// Epilogues (template EPI).  Round-1's epilogue (EPI 0: lane == row, 16-byte stores at a row stride → 32 LSU wavefronts
// per store instruction) cost a constant ~3.4 us per 256x256 tile and held the K <= 1024 shapes at 0.65x cuBLAS.  The
// round-2 hardware sweep (bench/gemm_variants.py → profiles/gemm_variants_r2.jsonl) decided the defaults:
//   EPI = 3  (default when N % 8 == 0)  TMA-store epilogue: each warp packs 32 rows x 64 columns into a 128B-swizzled staging
//            tile (two per warp, so a store is in flight while the next chunk is packed) and one lane issues
//            cp.async.bulk.tensor.2d.global.shared::cta — no LSU wavefronts at all, bounds clipped by the TMA unit.
//            16384x3072x768: 98.2 → 57.2 us (cuBLAS 62.4);  16384x2304x768: 74.5 → 45.9 us (cuBLAS 51.2).
//   EPI = 1  (default otherwise)  staged epilogue: each warp transposes 32 rows x 64 columns through padded shared memory
//            and writes full 128-byte lines (4 rows per store instruction); handles any N.
//   EPI = 0  the round-1 epilogue — kept selectable (variant bits) as the A/B baseline of that measurement.
// (The sweep also had an eight-epilogue-warp variant and "never store" / "never MMA" diagnostic builds; they decided nothing
//  further — eight warps: 73.8 us on the same shape — and were deleted.  Their rows stay in profiles/gemm_variants_r2.jsonl.)
#include "gemm_common.cuh"

namespace {

template <int BNT, int EPI>
struct CfgX {
  using C = Cfg2<BNT>;
  static constexpr int EPI_WARPS = 4;
  static constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
  static constexpr int ROW_PITCH = 144;                                  // 128 B of bf16 + 16 B pad: conflict-free both ways
  static constexpr int BAR_BYTES = 1024;                                 // keeps the staging area 1024-byte aligned (swizzle)
  static constexpr int STAGING = EPI == 1 ? 4 * 32 * ROW_PITCH           // one padded 32-row buffer per epilogue warp
                                 : (EPI == 3 ? 4 * 2 * 4096 : 0);        // two swizzled 32 x 128 B buffers per warp
  static constexpr int SMEM_BYTES = C::STAGES * C::STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + STAGING;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may opt into");
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// 32 fp32 accumulator columns of one row → (+bias, ReLU) → 16 packed bf16x2 words
__device__ __forceinline__ void finish32(const uint32_t* r, const GemmParams& p, int col0, uint32_t* out16) {
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float a = __uint_as_float(r[j]), b = __uint_as_float(r[j + 1]);
    if (p.bias != nullptr) {
      if (col0 + j < p.N) a += p.bias[col0 + j];
      if (col0 + j + 1 < p.N) b += p.bias[col0 + j + 1];
    }
    if (p.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
    out16[j >> 1] = psb::pack_bf16x2(a, b);
  }
}

template <int BNT, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((CfgX<BNT, EPI>::NTHREADS), 1)
psb_bcast_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                        const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using C = Cfg2<BNT>;
  using X = CfgX<BNT, EPI>;
  constexpr int STAGES2 = C::STAGES, STAGE2_BYTES = C::STAGE_BYTES, A2_BYTES = C::A_BYTES, TMEM_COLS2 = C::TMEM_COLS;
  constexpr int BN2 = BNT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES2 * STAGE2_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES2;
  uint64_t* tmem_full = bars + 2 * STAGES2;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  uint8_t* staging = smem + STAGES2 * STAGE2_BYTES + X::BAR_BYTES;      // EPI 1 / 3 only; 1024-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int ncl = gridDim.x / 2, cl = blockIdx.x / 2;
  const int tiles_m = (p.M + 255) / 256, tiles_n = (p.N + BN2 - 1) / BN2;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    if constexpr (EPI == 3) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * X::EPI_WARPS);       // every epilogue warp of both CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"((uint32_t)TMEM_COLS2)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) — identical to the production kernel =====================
    if (elect_one()) {
      if (p.ready_flag != nullptr) {
        psb::spin_until_ge(p.ready_flag, p.ready_epoch, p.err_slot, p.timeout_ns);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      uint32_t stage = 0, phase = 0;
      for (int tile = cl; tile < num_tiles; tile += ncl) {
        const int m0 = (tile / tiles_n) * 256 + (int)cta * BM;
        const int n0 = (tile % tiles_n) * BN2 + (int)cta * (BN2 / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE2_BYTES;
          const uint32_t lead_full = smem_u32(&full[stage]) & PEER_MASK;
          if (leader) mbar_expect_tx(&full[stage], 2 * STAGE2_BYTES);
          tma_load_2d_2sm(&tmap_a, lead_full, sa, kb * BK, m0);
          tma_load_2d_2sm(&tmap_b, lead_full, sa + A2_BYTES, kb * BK, n0);
          if (++stage == STAGES2) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = make_idesc2<BNT>();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = cl; tile < num_tiles; tile += ncl) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * STAGE2_BYTES);
            const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + A2_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma2(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            umma_commit_2sm(&empty[stage]);
            if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == STAGES2) stage = 0, phase ^= 1;
        }
        if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;                                     // the TMEM lane quarter this warp may read
    constexpr int COLS_PER_WARP = BN2;
    constexpr int cbeg = 0;
    const bool vec_ok = (p.N % 8) == 0;
    constexpr bool store = true;
    uint32_t acc = 0, acc_phase = 0;
    [[maybe_unused]] uint32_t nchunk = 0;                              // EPI 3: staging-buffer parity
    for (int tile = cl; tile < num_tiles; tile += ncl) {
      const int m0 = (tile / tiles_n) * 256 + (int)cta * BM, n0 = (tile % tiles_n) * BN2;
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN2;
      if constexpr (EPI == 1) {
        const uint32_t stg = smem_u32(staging + (warp - 2) * 32 * X::ROW_PITCH);
#pragma unroll 1
        for (int c0 = 0; c0 < BN2; c0 += 64) {
          uint32_t r[32], pk[32];
          tmem_ld_32x32b_x32(t_row + c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish32(r, p, n0 + c0, pk);
          tmem_ld_32x32b_x32(t_row + c0 + 32, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish32(r, p, n0 + c0 + 32, pk + 16);
          // lane == row: 128 B of this row into the padded staging tile (quarter-warps hit 32 distinct banks)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(stg + lane * X::ROW_PITCH + j * 16, make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
          __syncwarp();
          // 8 lanes per row: every store instruction writes 4 full 128-byte lines
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3), ch = lane & 7;
            const uint4 v = ld_shared_v4(stg + rr * X::ROW_PITCH + ch * 16);
            const int grow = m0 + quarter * 32 + rr, gcol = n0 + c0 + ch * 8;
            if (store && grow < p.M && gcol < p.N) {
              __nv_bfloat16* o = p.out + (size_t)grow * p.N + gcol;
              if (vec_ok && gcol + 8 <= p.N) {
                *reinterpret_cast<uint4*>(o) = v;
              } else {
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int t = 0; t < 8; ++t)
                  if (gcol + t < p.N) reinterpret_cast<uint16_t*>(o)[t] = (uint16_t)(w[t >> 1] >> ((t & 1) * 16));
              }
            }
          }
          __syncwarp();                                               // the staging tile is reused by the next chunk
        }
      } else if constexpr (EPI == 3) {
        const uint32_t stg0 = smem_u32(staging + (warp - 2) * 2 * 4096);
#pragma unroll 1
        for (int c0 = 0; c0 < BN2; c0 += 64) {
          const uint32_t stg = stg0 + (nchunk & 1u) * 4096;
          // the store issued two chunks ago has finished READING this buffer (bulk groups are per issuing thread)
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          uint32_t r[32], pk[32];
          tmem_ld_32x32b_x32(t_row + c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish32(r, p, n0 + c0, pk);
          tmem_ld_32x32b_x32(t_row + c0 + 32, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish32(r, p, n0 + c0 + 32, pk + 16);
          // lane == row; 16-byte chunk j of row r lives at chunk (j ^ (r & 7)) — the SWIZZLE_128B pattern of tmap_c
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes → visible to the TMA unit
          __syncwarp();
          if (lane == 0) {
            if (store && n0 + c0 < p.N && m0 + quarter * 32 < p.M)
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tmap_c), "r"(stg),
                           "r"(n0 + c0), "r"(m0 + quarter * 32)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");       // an empty group keeps the buffer parity uniform
          }
          ++nchunk;
        }
      } else {
        const int row = m0 + quarter * 32 + lane;
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + COLS_PER_WARP; c0 += 32) {
          uint32_t r[32], pk[16];
          tmem_ld_32x32b_x32(t_row + c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          finish32(r, p, n0 + c0, pk);
          if (store && row < p.M && n0 + c0 < p.N) {
            __nv_bfloat16* orow = p.out + (size_t)row * p.N + n0 + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = n0 + c0 + 8 * j;
              if (vec_ok && col + 8 <= p.N) {
                *reinterpret_cast<uint4*>(orow + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              } else {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                  if (col + t < p.N)
                    reinterpret_cast<uint16_t*>(orow + 8 * j)[t] = (uint16_t)(pk[4 * j + (t >> 1)] >> ((t & 1) * 16));
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&tmem_empty[acc]) & PEER_MASK);
      if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
    }
    if constexpr (EPI == 3) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output tiles written before exit
      __syncwarp();
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS2)
                 : "memory");
  }
}

template <int BNT, int EPI>
void launch_x(cudaStream_t s, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
              int clusters) {
  using X = CfgX<BNT, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(psb_bcast_gemm2_kernel<BNT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, X::SMEM_BYTES);
    configured = true;
  }
  psb_bcast_gemm2_kernel<BNT, EPI><<<2 * clusters, X::NTHREADS, X::SMEM_BYTES, s>>>(ta, tb, tc, p);
}

template <int BNT>
void launch_e(cudaStream_t s, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
              const GemmParams& p, int clusters) {
  if (epi == 1) launch_x<BNT, 1>(s, ta, tb, tc, p, clusters);
  else if (epi == 3) launch_x<BNT, 3>(s, ta, tb, tc, p, clusters);
  else launch_x<BNT, 0>(s, ta, tb, tc, p, clusters);
}

}  // namespace

// `a.two_cta` must be set (the tensor maps are built for the 2-CTA box shapes).
// epi: -1 = auto (TMA store when N % 8 == 0 and `tmap_out` is given, else staged), 0 / 1 / 3 = force (see top).
// `tmap_out` (EPI 3): CUtensorMap of the [M, N] bf16 output, box 64 columns x 32 rows, SWIZZLE_128B.
void psb_launch_bcast_gemm2(cudaStream_t s, const BcastGemmArgs& a, int num_sms, int epi, const void* tmap_out) {
  if (epi < 0) epi = (tmap_out != nullptr && a.N % 8 == 0) ? 3 : 1;
  GemmParams p{};
  p.bias = a.bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(a.tmap_c));
  p.ready_flag = a.ready_flag;
  p.ready_epoch = a.ready_epoch;
  p.err_slot = a.ready_flag != nullptr ? const_cast<uint64_t*>(a.ready_flag) - SIG_PARAMS_READY + SIG_ERROR : nullptr;
  p.M = a.M, p.N = a.N, p.K = a.K, p.relu = a.relu;
  p.timeout_ns = a.timeout_ns;
  psb_count_launch(1);
  const int bnt = a.N <= 64 ? 64 : (a.N <= 128 ? 128 : 256);
  const int tiles = ((a.M + 255) / 256) * ((a.N + bnt - 1) / bnt);
  int clusters = num_sms / 2;
  if (tiles < clusters) clusters = tiles;
  const CUtensorMap& ta = *reinterpret_cast<const CUtensorMap*>(a.tmap_a);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(a.tmap_b);
  const CUtensorMap& tc = tmap_out != nullptr ? *reinterpret_cast<const CUtensorMap*>(tmap_out) : ta;   // unused unless EPI 3
  if (bnt == 64) launch_e<64>(s, epi, ta, tb, tc, p, clusters);
  else if (bnt == 128) launch_e<128>(s, epi, ta, tb, tc, p, clusters);
  else launch_e<256>(s, epi, ta, tb, tc, p, clusters);
}
