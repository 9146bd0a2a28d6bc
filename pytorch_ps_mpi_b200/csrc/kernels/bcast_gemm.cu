// bcast_gemm: Y[M,N] = act(X[M,K] · W[N,K]^T + bias) in bf16 with fp32 accumulation on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA into 128B-swizzled
// shared memory), whose WEIGHT operand is the tile the parameter server just broadcast.
//
// Fusion with the PS→worker broadcast (K4 in SURVEY §2.6; the Ibcast + Wait + first forward
// matmul of /root/reference/mpi_comms.py:120-133): the weight tensor map points INTO the symmetric
// parameter arena (this rank's copy that the server's multimem.st fills — or, in pull mode, the
// server's own arena mapped over NVLink).  The TMA producer warp acquires the PARAMS_READY epoch
// flag (ld.acquire.sys + fence.proxy.async) immediately before its first weight load, so the
// kernel launches while the broadcast is still in flight: TMEM allocation, barrier init,
// descriptor prefetch and the scheduler prologue overlap the tail of the server's update kernel,
// and this GEMM *is* the req.Wait() for the whole forward pass that follows it on the stream.
//
// Structure (one CTA per SM, persistent over output tiles):
// (M >= 256 runs on the cta_group::2 kernel in bcast_gemm2.cu; this 1-CTA kernel serves small M.)
//   warp 0      TMA producer   cp.async.bulk.tensor.2d → smem ring (4 stages x {A 128x64, B 256x64} = 192 KB)
//   warp 1      MMA issuer     one elected lane: tcgen05.mma.cta_group::1.kind::f16, UMMA 128x256x16,
//                              tcgen05.commit → frees smem slots / publishes the accumulator
//   warps 2..5  epilogue       tcgen05.ld 32x32b → +bias → ReLU → bf16 → 16-byte global stores
//   TMEM        2 accumulator stages x BN fp32 columns (double-buffered against the epilogue)
#include "gemm_common.cuh"

namespace {

__global__ void __launch_bounds__(THREADS, 1)
psb_bcast_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;     // [ACC_STAGES]
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation is warp-wide; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // The broadcast gate: do not touch a weight tile before the server published this epoch.
      if (p.ready_flag != nullptr) {
        psb::spin_until_ge(p.ready_flag, p.ready_epoch, p.err_slot, p.timeout_ns);
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy acquire → async-proxy (TMA) reads
      }
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_2d(&tmap_a, &full[stage], sa, kb * BK, m0);
          tma_load_2d(&tmap_b, &full[stage], sa + A_BYTES, kb * BK, n0);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc();
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = make_desc(a_addr), db = make_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr>>4) units
            umma(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);                         // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tmem_full[acc]);  // accumulator complete → epilogue
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + quarter * 32 + lane;
      const bool vec_ok = (p.N % 8) == 0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < p.M) {
          __nv_bfloat16* orow = p.out + (size_t)row * p.N + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = n0 + c0 + j;
            float v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              v[t] = __uint_as_float(r[j + t]);
              if (p.bias != nullptr && col + t < p.N) v[t] += p.bias[col + t];
              if (p.relu) v[t] = fmaxf(v[t], 0.f);
            }
            if (vec_ok && col + 8 <= p.N) {
              uint4 o = make_uint4(psb::pack_bf16x2(v[0], v[1]), psb::pack_bf16x2(v[2], v[3]),
                                   psb::pack_bf16x2(v[4], v[5]), psb::pack_bf16x2(v[6], v[7]));
              *reinterpret_cast<uint4*>(orow + j) = o;
            } else {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                if (col + t < p.N) orow[j + t] = __float2bfloat16_rn(v[t]);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) acc = 0, acc_phase ^= 1;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}


}  // namespace

int psb_bcast_gemm_smem_bytes() { return SMEM_BYTES; }

void psb_launch_bcast_gemm(cudaStream_t s, const BcastGemmArgs& a, int num_sms) {
  if (a.two_cta) {   // M >= 256: the cta_group::2 kernel with the TMA-store (or staged) epilogue, bcast_gemm2.cu
    psb_launch_bcast_gemm2(s, a, num_sms, -1, a.tmap_out);
    return;
  }
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(psb_bcast_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    configured = true;
  }
  GemmParams p{};
  p.bias = a.bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(a.tmap_c));   // tmap_c carries the output pointer
  p.ready_flag = a.ready_flag;
  p.ready_epoch = a.ready_epoch;
  p.err_slot = a.ready_flag != nullptr ? const_cast<uint64_t*>(a.ready_flag) - SIG_PARAMS_READY + SIG_ERROR : nullptr;
  p.M = a.M, p.N = a.N, p.K = a.K, p.relu = a.relu;
  p.timeout_ns = a.timeout_ns;
  psb_count_launch(1);
  const int tiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  psb_bcast_gemm_kernel<<<grid, THREADS, SMEM_BYTES, s>>>(*reinterpret_cast<const CUtensorMap*>(a.tmap_a),
                                                          *reinterpret_cast<const CUtensorMap*>(a.tmap_b), p);
}
