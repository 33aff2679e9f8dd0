// tc_gemm.cu — dense projections on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
//   Y[M_tok, N] = X[M_tok, K] * W[N, K]^T          (HF nn.Linear, HF:models/llama/modeling_llama.py:238-249,177-183)
//
// computed transposed so the WEIGHT rows ride the 128 TMEM lanes:  D[128 weight rows, TN tokens] += A * B^T with
//   A = W tile  [128 x 64]  (K-major, TMA 128B-swizzled)      -> UMMA M = 128
//   B = X tile  [TN  x 64]  (K-major, TMA 128B-swizzled)      -> UMMA N = TN (16..256)
// Tokens are the UMMA N dimension, so the same kernel serves prefill (TN = 128 / 256, tensor-pipe bound) and batched
// decode (TN = 16/32, HBM bound: weights stream HBM -> TMA -> smem -> tensor core without touching registers).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> fused RoPE/KV-append | SwiGLU | residual, same math as gemv_epilogue).
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "sk_partition.h"
#include "tc_gemm.h"

#include "epilogue.cuh"

constexpr int TC_BM = 128;   // weight rows per tile (UMMA M)
constexpr int TC_BK = 64;    // K elements per stage = one 128-byte swizzle span of bf16
constexpr int TC_THREADS = 192;

SSB_DEVINL void tma_load_2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst_smem)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
SSB_DEVINL void prefetch_tmap(const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }
SSB_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SSB_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SSB_DEVINL void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
SSB_DEVINL void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
SSB_DEVINL void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
SSB_DEVINL void tc_ld8p(uint32_t taddr, uint32_t* v) {  // same, into v[0..7] of a larger register array
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
SSB_DEVINL void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: 8-row groups of 128 B rows, groups 1024 B apart
// (cute::UMMA::SmemDescriptor bit layout: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64))
SSB_DEVINL uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// TC_SK_SMEM_BUDGET (compile-time experiment, variant library "sk2cta": 104 KiB + TC_SK_PREFETCH_RING): shared-memory
// budget of the ring for the decode-sized tiles (TN <= 64).  At ~100 KiB two CTAs fit on an SM, so under PDL the CTAs of
// the NEXT projection become resident while this one is still in its mainloop / epilogue, set up their barriers and TMEM
// and put their first weight tiles in flight; 5 x 20 KiB per CTA still covers the per-SM HBM latency-bandwidth product.
#ifndef TC_SK_SMEM_BUDGET
#define TC_SK_SMEM_BUDGET (200 * 1024)
#endif
template <int TN>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;                       // 16 KiB
  static constexpr int B_BYTES = TN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  static constexpr int BUDGET = TN <= 64 ? TC_SK_SMEM_BUDGET : 200 * 1024;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * TN <= 32) ? 32 : (2 * TN <= 64) ? 64 : (2 * TN <= 128) ? 128 : (2 * TN <= 256) ? 256 : 512;
  static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 2 * TN * 4 /*per-token pos / KV block of a tile*/;
};

#ifndef TC_EPI_V2
#define TC_EPI_V2 1  // prefill epilogue in passes of 32 tokens with the next pass's inputs prefetched (0 = 8 tokens at a time; variant "tcepi1")
#endif
// Epilogue inputs of CG consecutive tokens of one row pair — residual words (EPI_RESID) or packed RoPE cos|sin
// (EPI_QKV_ROPE, positions staged in shared memory).  They do not depend on the accumulator, so a caller can have them in
// flight while the MMAs of the tile (or the previous pass's stores) run; `on` = this lane runs the epilogue of a real row.
template <int EPI, int CG>
SSB_DEVINL void tc_epi_inputs(const GemvArgs& a, int pair, int m0, const int* __restrict__ s_pos, bool on, uint32_t (&w)[CG]) {
  if constexpr (EPI == EPI_RESID) {
#pragma unroll
    for (int j = 0; j < CG; ++j)
      w[j] = (on && m0 + j < a.M) ? __ldcg(reinterpret_cast<const uint32_t*>(a.resid + (size_t)(m0 + j) * a.ld_out + 2 * pair)) : 0u;
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    const int half = a.head_dim >> 1, q_pairs = a.q_rows >> 1, qk_pairs = (a.q_rows + a.kv_rows) >> 1;
    const int pp = pair < q_pairs ? pair : pair - q_pairs;
    const int jj = pp % half;
#pragma unroll
    for (int j = 0; j < CG; ++j) w[j] = (on && pair < qk_pairs && m0 + j < a.M) ? a.rope_cs[(size_t)s_pos[j] * half + jj] : 0u;
  }
}
// eight tokens of one row pair with their inputs already loaded (w = the eight words tc_epi_inputs produced for them)
template <int EPI>
SSB_DEVINL void tc_epi_store8(const GemvArgs& a, int pair, int m0, const int* __restrict__ s_pos, const int* __restrict__ s_blk,
                              const uint32_t* w, const float (&mine)[8], const float (&other)[8]) {
  if constexpr (EPI == EPI_RESID) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (m0 + j < a.M)
        *reinterpret_cast<uint32_t*>(a.out_bf16 + (size_t)(m0 + j) * a.ld_out + 2 * pair) =
            pack_bf16(bf16r(mine[j]) + bf_lo(w[j]), bf16r(other[j]) + bf_hi(w[j]));
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    tc_epilogue8_qkv_cs(a, pair, m0, s_pos, s_blk, w, mine, other);
  } else {
    tc_epilogue8<EPI>(a, pair, m0, mine, other);
  }
}

// One pass of the fused epilogue: CG tokens of one row PAIR, split between the pair's two lanes.  Each lane holds its own
// row's CG accumulators f[]; the even lane finishes tokens [0, CG/2) of the pass, the odd lane tokens [CG/2, CG), after
// swapping the halves the partner needs.  (The first version let the even lane do all tokens while the odd lane idled: with
// one epilogue warp per scheduler the epilogue is a dependent-issue chain, so instructions per warp are what it costs —
// the split halves them.)  w = the inputs of THIS lane's CG/2 tokens (tc_epi_inputs at token m_base + off).  Must be
// called by all 32 lanes (shuffles).  s_pos / s_blk point at the pass's first token inside the staged tile.
template <int EPI, int CG>
SSB_DEVINL void tc_epi_pass(const GemvArgs& a, int pair, bool on, int odd, int m_base, const int* __restrict__ s_pos,
                            const int* __restrict__ s_blk, const float (&f)[CG], const uint32_t (&w)[CG / 2]) {
  constexpr int H = CG / 2;
  float recv[H];
#pragma unroll
  for (int j = 0; j < H; ++j) recv[j] = __shfl_xor_sync(0xffffffffu, odd ? f[j] : f[H + j], 1);
  const int off = odd ? H : 0;
#pragma unroll
  for (int c = 0; c < H; c += 8) {
    float v0[8], v1[8];  // even row (gate / first of the RoPE pair), odd row
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float own = odd ? f[H + c + j] : f[c + j];
      v0[j] = odd ? recv[c + j] : own;
      v1[j] = odd ? own : recv[c + j];
    }
    if (on && m_base + off + c < a.M) tc_epi_store8<EPI>(a, pair, m_base + off + c, s_pos + off + c, s_blk + off + c, &w[c], v0, v1);
  }
}

template <int TN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemvArgs a) {
  using Cfg = TcCfg<TN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;   // [2]
  uint64_t* tempty = tfull + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  [[maybe_unused]] int* s_pos = reinterpret_cast<int*>(bars) + 64;  // [TN] position of each token of the current tile (QKV epilogue)
  [[maybe_unused]] int* s_blk = s_pos + TN;                        // [TN] KV block holding that position

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ntiles = (a.N + TC_BM - 1) / TC_BM;
  const int n_mtiles = (a.M + TN - 1) / TN;
  const int n_tiles = n_ntiles * n_mtiles;
  const int nkb = (a.K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // activations (B operand, residual) come from the previous kernel

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile % n_mtiles, nt = tile / n_mtiles;  // token tiles fastest: CTAs running together share a weight tile (one HBM read)
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_2d(sa, &tmA, kb * TC_BK, nt * TC_BM, &full[stage]);
          tma_load_2d(sa + Cfg::A_BYTES, &tmB, kb * TC_BK, mt * TN, &full[stage]);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (one thread)
    constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, TN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint64_t ad = umma_desc_k_sw128(sa);
          const uint64_t bd = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)  // UMMA_K = 16 bf16 = 32 B: advance the start address inside the span
            tc_mma_bf16(tmem_base + acc * TN, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          tc_commit(&empty[stage]);                    // frees the smem stage when these MMAs retire
          if (kb == nkb - 1) tc_commit(&tfull[acc]);   // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps (TMEM lane quadrant = warp % 4)
    const int quad = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mt = tile % n_mtiles, nt = tile / n_mtiles;  // token tiles fastest: CTAs running together share a weight tile (one HBM read)
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if constexpr (EPI == EPI_QKV_ROPE) {
        // stage the tile's per-token position and KV block while its MMAs run (two dependent loads, once per token)
        named_bar_sync(2, 128);  // the previous tile's epilogue no longer reads s_pos / s_blk
        for (int t = threadIdx.x - 64; t < TN; t += 128) {
          const int m = mt * TN + t;
          int pos = 0, blk = 0;
          if (m < a.M) {
            pos = __ldcg(a.row_pos + m);
            blk = a.block_table[(size_t)a.row_slot[m] * a.bt_stride + pos / a.block_size];
          }
          s_pos[t] = pos;
          s_blk[t] = blk;
        }
        named_bar_sync(2, 128);
      }
      const int row = nt * TC_BM + quad * 32 + lane;  // physical weight row (even = first of a pair)
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN;
#if TC_EPI_V2
      // passes of CG tokens: one TMEM read and one batch of input loads per pass instead of per 8 tokens (each batch is a
      // full L2 round trip, ~0.7 us; a 256-token tile used to pay 32 of them in sequence), and the inputs of pass p+1 are
      // requested before pass p is converted and stored.  The first batch flies while the tile's MMAs finish.
      constexpr int CG = TN < 32 ? TN : 32;
      constexpr int H = CG / 2;  // tokens per lane and pass (tc_epi_pass)
      const bool on = row < a.N;
      const int pair = row >> 1, odd = lane & 1, off = odd ? H : 0;
      uint32_t w[H], wn[H];
      tc_epi_inputs<EPI, H>(a, pair, mt * TN + off, s_pos + off, on, w);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int g0 = 0; g0 < TN; g0 += CG) {
        if (mt * TN + g0 >= a.M) break;  // warp-uniform
        uint32_t v[CG];
#pragma unroll
        for (int q = 0; q < CG / 8; ++q) tc_ld8p(taddr + g0 + 8 * q, &v[8 * q]);
        if (g0 + CG < TN && mt * TN + g0 + CG < a.M)
          tc_epi_inputs<EPI, H>(a, pair, mt * TN + g0 + CG + off, s_pos + g0 + CG + off, on, wn);
        tc_wait_ld();
        float f[CG];
#pragma unroll
        for (int j = 0; j < CG; ++j) f[j] = __uint_as_float(v[j]);
        tc_epi_pass<EPI, CG>(a, pair, on, odd, mt * TN + g0, s_pos + g0, s_blk + g0, f, w);
#pragma unroll
        for (int j = 0; j < H; ++j) w[j] = wn[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
#else
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < TN; c += 8) {
        if (mt * TN + c >= a.M) break;  // warp-uniform
        uint32_t v[8];
        tc_ld8(taddr + c, v);
        tc_wait_ld();
        float mine[8], other[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          mine[j] = __uint_as_float(v[j]);
          other[j] = __shfl_xor_sync(0xffffffffu, mine[j], 1);
        }
        if constexpr (EPI == EPI_QKV_ROPE) {
          if (!(lane & 1) && row < a.N) tc_epilogue8_qkv_staged(a, row >> 1, mt * TN + c, s_pos + c, s_blk + c, mine, other);
        } else {
          if (!(lane & 1) && row < a.N) tc_epilogue8<EPI>(a, row >> 1, mt * TN + c, mine, other);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

// ====================================================================================================================
// stream-K variant for ONE token tile (batched decode, M <= TN <= 64): the (weight tile, k-block) units of the whole
// matrix are laid end to end and every CTA streams an equal contiguous share, so all SMs pull the same number of weight
// bytes no matter how many 128-row tiles N has (a 4096-row matrix is only 32 tiles for 148 SMs).  A tile whose k-range
// is split over several CTAs is reduced through a small fp32 workspace: the parts that do not start at k = 0 are
// written out ("contributors"), the CTA holding the k = 0 part ("owner" — it reaches that part LAST in its own range,
// when the others are already done) adds them to its TMEM accumulator in registers and runs the fused epilogue.
// Dependencies only point from higher to lower CTA index ranges that are processed earlier, so there is no cycle.
// ====================================================================================================================
struct SkWs {
  float* part;      // [slots][TN/8][128][8] fp32
  unsigned* flags;  // [slots], zero on entry and on exit
  unsigned long long* prof;  // [grid][8] globaltimer stamps (TC_SK_PROF builds), may be null
};
#ifndef TC_SK_PROF
#define TC_SK_PROF 0
#endif
#ifndef TC_SK_EPI_V2
#define TC_SK_EPI_V2 1  // batched owner epilogue (0 = the first version, kept for the A/B: variant library "skepi1")
#endif
// phase stamps of one stream-K launch, per CTA: 0 entry, 1 setup done (barriers, TMEM), 2 first weight tile landed,
// 3 last MMA committed, 4 epilogue saw the last accumulator, 5 epilogue done, 6 exit, 7 = number of segments
SSB_DEVINL void sk_stamp(const SkWs& ws, int i) {
#if TC_SK_PROF
  if (ws.prof) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    ws.prof[(size_t)blockIdx.x * 8 + i] = t;
  }
#endif
}

#ifndef TC_SK_PREFETCH_RING
#define TC_SK_PREFETCH_RING 0
#endif
template <int TN, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_sk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemvArgs a, const SkWs ws) {
  using Cfg = TcCfg<TN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::STAGES;
  uint64_t* tfull = empty + Cfg::STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_ntiles = (a.N + TC_BM - 1) / TC_BM;
  const int nkb = (a.K + TC_BK - 1) / TC_BK;
  // 32-bit unit arithmetic (the launcher checks U * grid < 2^31): the 64-bit divisions of the first version were software
  // routines on the critical path of every segment and of the owner's contributor search
  const int U = n_ntiles * nkb;
  const int G = (int)gridDim.x;
  const int u0 = sk_begin(blockIdx.x, U, G), u1 = sk_begin(blockIdx.x + 1, U, G);  // sk_partition.h (checked exhaustively on the host)
  [[maybe_unused]] int* s_pos = reinterpret_cast<int*>(bars) + 64;  // [TN] position of each token (QKV epilogue), staged once
  [[maybe_unused]] int* s_blk = s_pos + TN;                        // [TN] KV block holding that position

  if (threadIdx.x == 0) {
    sk_stamp(ws, 0);
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) sk_stamp(ws, 1);
  pdl_launch_dependents();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer.  The weight (A) stream does not
    // depend on the previous kernel: it starts before griddepcontrol.wait; the activation (B) loads wait for it.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool waited = false;
      int ustart = u0;
#if TC_SK_PREFETCH_RING
      // compile-time experiment for round 2 (never run on hardware): put the weight tiles of the first ring-full of
      // units in flight before the dependency wait instead of only the first one, so the HBM stream of this projection
      // ramps up while the previous kernel drains; the activation tiles follow after the wait on the same barriers.
      {
        const int upre = min(u1, u0 + Cfg::STAGES);
        int st = 0;
        for (int u = u0; u < upre; ++u, ++st) {  // the ring starts empty: no empty-wait in the first pass
          mbar_expect_tx(&full[st], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_2d(smem + (size_t)st * Cfg::STAGE_BYTES, &tmA, (int)(u % nkb) * TC_BK, (int)(u / nkb) * TC_BM, &full[st]);
        }
        pdl_wait();
        waited = true;
        st = 0;
        for (int u = u0; u < upre; ++u, ++st)
          tma_load_2d(smem + (size_t)st * Cfg::STAGE_BYTES + Cfg::A_BYTES, &tmB, (int)(u % nkb) * TC_BK, 0, &full[st]);
        const int n = (int)(upre - u0);
        stage = n % Cfg::STAGES;
        phase = n == Cfg::STAGES ? 1u : 0u;
        ustart = upre;
      }
#endif
      for (int u = ustart; u < u1; ++u) {
        const int nt = u / nkb, kb = u - nt * nkb;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
        tma_load_2d(sa, &tmA, kb * TC_BK, nt * TC_BM, &full[stage]);
        if (!waited) {
          pdl_wait();
          waited = true;
        }
        tma_load_2d(sa + Cfg::A_BYTES, &tmB, kb * TC_BK, 0, &full[stage]);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, TN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int u = u0; u < u1; ++it) {
      const int kb0 = u % nkb;
      const int kb1 = min(nkb, kb0 + (u1 - u));
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          if (TC_SK_PROF && u == u0 && kb == kb0) sk_stamp(ws, 2);
          const uint32_t sa = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint64_t ad = umma_desc_k_sw128(sa);
          const uint64_t bd = umma_desc_k_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc_mma_bf16(tmem_base + acc * TN, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb > kb0) || k != 0);
          tc_commit(&empty[stage]);
          if (kb == kb1 - 1) tc_commit(&tfull[acc]);
          if (TC_SK_PROF && u + (kb1 - kb0) >= u1 && kb == kb1 - 1) sk_stamp(ws, 3);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      u += kb1 - kb0;
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps
    pdl_wait();  // residuals / positions written by the previous kernel
    const int quad = warp & 3;
    const int erow = quad * 32 + lane;  // row inside the 128-row tile
    if constexpr (EPI == EPI_QKV_ROPE && TC_SK_EPI_V2) {
      // per-token position and KV block, once per CTA (two dependent loads) instead of per thread and 8-token chunk
      for (int t = threadIdx.x - 64; t < TN; t += 128) {
        int pos = 0, blk = 0;
        if (t < a.M) {
          pos = __ldcg(a.row_pos + t);
          blk = a.block_table[(size_t)a.row_slot[t] * a.bt_stride + pos / a.block_size];
        }
        s_pos[t] = pos;
        s_blk[t] = blk;
      }
      named_bar_sync(2, 128);
    }
    int it = 0;
    for (int u = u0; u < u1; ++it) {
      const int nt = u / nkb, kb0 = u - nt * nkb;
      const int kb1 = min(nkb, kb0 + (u1 - u));
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool is_first_seg = (u == u0);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (TC_SK_PROF && warp == 2 && lane == 0 && u + (kb1 - kb0) >= u1) sk_stamp(ws, 4);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN;
      const int row = nt * TC_BM + erow;
      if (kb0 != 0) {
        // contributor: dump the partial accumulator to this CTA's slot (slot 0 = first segment of the range, else 1)
        const int slot = blockIdx.x * 2 + (is_first_seg ? 0 : 1);
        float* dst = ws.part + (size_t)slot * (TN * 128);
#pragma unroll 1
        for (int c = 0; c < TN; c += 8) {
          uint32_t v[8];
          tc_ld8(taddr + c, v);
          tc_wait_ld();
          float4* d4 = reinterpret_cast<float4*>(dst + ((size_t)(c / 8) * 128 + erow) * 8);
          d4[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
          d4[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        }
        tc_fence_before();
#if !TC_SK_EPI_V2
        __threadfence();
#endif
        // the barrier orders the 128 threads' stores before warp 2's release store (release is cumulative at gpu scope)
        named_bar_sync(2, 128);
        if (warp == 2 && lane == 0) {
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ws.flags + slot), "r"(1u) : "memory");
          mbar_arrive(&tempty[acc]);
        } else if (lane == 0) {
          mbar_arrive(&tempty[acc]);
        }
      } else {
        // owner: the part starting at k = 0.  Contributors = the CTAs whose ranges cover the rest of this tile: the next
        // n_contrib CTAs (every CTA has a non-empty range: grid <= U), each with its FIRST segment.
#if TC_SK_EPI_V2
        // Version 2.  The first version spent ~11 us here per owner (profiles/r02_sk_phase_stamps_v1.txt) in 12-16 SERIAL L2
        // round trips: flag by flag, then per 8-token chunk the partials, then the residual / RoPE inputs.  Now every load
        // that does not depend on another is issued together: epilogue inputs first (they depend on nothing computed
        // here), all flags in one batch, partials of 32 tokens x 2 contributors in flight at a time.
        const int t_end = (nt + 1) * nkb;  // first unit after this tile
        const int n_contrib = sk_contributors(blockIdx.x, U, G, u + (kb1 - kb0), t_end);
        const int slot0 = (blockIdx.x + 1) * 2;  // contributor i dumped into slot0 + 2 i
        constexpr int CG = TN < 32 ? TN : 32;    // tokens per pass
        constexpr int H = CG / 2;                // tokens per lane and pass (tc_epi_pass)
        const bool on = row < a.N;
        const int pair = row >> 1, odd = lane & 1, off = odd ? H : 0;
        bool polled = false;
#pragma unroll 1
        for (int g0 = 0; g0 < TN; g0 += CG) {
          if (g0 >= a.M) break;  // warp-uniform
          uint32_t v[CG];
#pragma unroll
          for (int q = 0; q < CG / 8; ++q) tc_ld8p(taddr + g0 + 8 * q, &v[8 * q]);
          // ---- epilogue inputs of this lane's half of the pass
          uint32_t w[H];
          tc_epi_inputs<EPI, H>(a, pair, g0 + off, s_pos + g0 + off, on, w);
          // ---- all contributors' flags in one round trip (they are normally up long before the owner gets here)
          if (!polled) {
            polled = true;
            for (;;) {
              unsigned ok = 1u;
#pragma unroll
              for (int i = 0; i < SK_MAX_CONTRIB; ++i) {
                unsigned f = 1u;
                if (i < n_contrib) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(ws.flags + slot0 + 2 * i) : "memory");
                ok &= (f != 0u) ? 1u : 0u;
              }
              if (ok) break;
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");  // acquire: the contributors' partials are visible to this thread
          }
          tc_wait_ld();
          float f[CG];
#pragma unroll
          for (int j = 0; j < CG; ++j) f[j] = __uint_as_float(v[j]);
          const float* pbase = ws.part + (size_t)slot0 * (TN * 128) + ((size_t)(g0 / 8) * 128 + erow) * 8;
#pragma unroll 1
          for (int i = 0; i < n_contrib; i += 2) {
            const bool two = i + 1 < n_contrib;
            const float* p0 = pbase + (size_t)(2 * i) * (TN * 128);
            const float* p1 = p0 + (size_t)2 * (TN * 128);
            float4 x[CG / 4], y[CG / 4];
#pragma unroll
            for (int q = 0; q < CG / 8; ++q) {
              x[2 * q] = __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)q * 128 * 8));
              x[2 * q + 1] = __ldcg(reinterpret_cast<const float4*>(p0 + (size_t)q * 128 * 8) + 1);
            }
#pragma unroll
            for (int q = 0; q < CG / 8; ++q) {
              y[2 * q] = two ? __ldcg(reinterpret_cast<const float4*>(p1 + (size_t)q * 128 * 8)) : make_float4(0.f, 0.f, 0.f, 0.f);
              y[2 * q + 1] = two ? __ldcg(reinterpret_cast<const float4*>(p1 + (size_t)q * 128 * 8) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // same order of additions as the first version: contributor i, then i + 1
#pragma unroll
            for (int q = 0; q < CG / 4; ++q) {
              f[4 * q] += x[q].x; f[4 * q + 1] += x[q].y; f[4 * q + 2] += x[q].z; f[4 * q + 3] += x[q].w;
            }
            if (two) {  // warp-uniform
#pragma unroll
              for (int q = 0; q < CG / 4; ++q) {
                f[4 * q] += y[q].x; f[4 * q + 1] += y[q].y; f[4 * q + 2] += y[q].z; f[4 * q + 3] += y[q].w;
              }
            }
          }
          tc_epi_pass<EPI, CG>(a, pair, on, odd, g0, s_pos + g0, s_blk + g0, f, w);
        }
        tc_fence_before();
        named_bar_sync(2, 128);  // every epilogue thread has consumed the partials
        if (warp == 2 && lane < n_contrib) ws.flags[slot0 + 2 * lane] = 0;  // clean for the next launch
        if (lane == 0) mbar_arrive(&tempty[acc]);
#else
        const int t_end = (nt + 1) * nkb;                     // first unit after this tile
        int cu = u + (kb1 - kb0);                              // first unit not covered by this CTA
        int n_contrib = 0;
        int cslot[12];  // the launcher sizes the grid so that a tile is never split over more than 10 CTAs
        for (int c = blockIdx.x + 1; cu < t_end && c < G && n_contrib < 12; ++c) {
          const int c0 = (int)((unsigned)c * (unsigned)U / (unsigned)G), c1 = (int)((unsigned)(c + 1) * (unsigned)U / (unsigned)G);
          if (c1 <= c0) continue;
          // CTA c's segment inside this tile starts at max(c0, cu) == c0 (ranges are contiguous) and is its FIRST segment
          cslot[n_contrib++] = c * 2 + 0;
          cu = c1;
        }
        for (int i = 0; i < n_contrib; ++i) {
          unsigned f;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(ws.flags + cslot[i]) : "memory");
          } while (f == 0u);
        }
#pragma unroll 1
        for (int c = 0; c < TN; c += 8) {
          if (c >= a.M) break;
          uint32_t v[8];
          tc_ld8(taddr + c, v);
          tc_wait_ld();
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j]);
          for (int i = 0; i < n_contrib; ++i) {
            const float4* p4 = reinterpret_cast<const float4*>(ws.part + (size_t)cslot[i] * (TN * 128) + ((size_t)(c / 8) * 128 + erow) * 8);
            const float4 x = __ldcg(p4), y = __ldcg(p4 + 1);
            f[0] += x.x; f[1] += x.y; f[2] += x.z; f[3] += x.w;
            f[4] += y.x; f[5] += y.y; f[6] += y.z; f[7] += y.w;
          }
          float other[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) other[j] = __shfl_xor_sync(0xffffffffu, f[j], 1);
          if (!(lane & 1) && row < a.N) tc_epilogue8<EPI>(a, row >> 1, c, f, other);
        }
        tc_fence_before();
        named_bar_sync(2, 128);  // every epilogue thread has consumed the partials
        if (warp == 2 && lane == 0)
          for (int i = 0; i < n_contrib; ++i) ws.flags[cslot[i]] = 0;  // clean for the next launch
        if (lane == 0) mbar_arrive(&tempty[acc]);
#endif
      }
      u += kb1 - kb0;
    }
    if (TC_SK_PROF && warp == 2 && lane == 0) {
      sk_stamp(ws, 5);
      if (ws.prof) ws.prof[(size_t)blockIdx.x * 8 + 7] = (unsigned long long)it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TC_SK_PROF && threadIdx.x == 0) sk_stamp(ws, 6);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D K-major bf16 tensor map: global [rows, cols] with row pitch ld elements, box [box_rows, 64 cols], 128B swizzle
cudaError_t tc_make_tmap(TcTensorMap* out, const bf16* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  static_assert(sizeof(TcTensorMap) == sizeof(CUtensorMap), "tensor map size");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Token-tile width (UMMA N).  Decode-sized forwards: the smallest tile that holds the batch (stream-K kernel).
int tc_pick_tn(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : 128; }
// Prefill-sized forwards choose per projection between 128- and 256-token tiles.  A 128 x 256 tile ingests 48 KiB of
// operands per k-block for twice the FLOPs of a 128 x 128 tile's 32 KiB (87 vs 64 FLOP per byte through the L2 -> SM path
// that bounds this kernel, profiles/r01_tcgen05_ncu_summary.txt: tensor pipe 21-36 % active) and halves the re-reads of a
// weight tile (one read per token tile) — but it also halves the tile count, and the persistent grid runs whole waves:
// pick the width with the smaller  waves x per-tile cost  (cost 1.0 vs 1.7: the operand bytes and a longer epilogue), 128 on a tie.
int tc_pick_tn_prefill(int M, int N, int n_sm) {
  if (M <= 128) return tc_pick_tn(M);
  const long long nt = (N + TC_BM - 1) / TC_BM;
  const long long t128 = nt * ((M + 127) / 128), t256 = nt * ((M + 255) / 256);
  const long long w128 = (t128 + n_sm - 1) / n_sm, w256 = (t256 + n_sm - 1) / n_sm;
  return 17 * w256 < 10 * w128 ? 256 : 128;  // measured: 1.5 was break-even at 7B / 512 tokens (TTFT 13.13 vs 12.87 ms), 7 % better at 16 K tokens
}
int tc_weight_box_rows() { return TC_BM; }

template <int TN, int EPI>
static cudaError_t launch_tc_t(const TcTensorMap& tmA, const TcTensorMap& tmB, const GemvArgs& a, const LaunchCfg& lc) {
  using Cfg = TcCfg<TN>;
  static std::atomic<unsigned long long> attr_mask{0};  // per instantiation, per device
  if (first_launch_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<TN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
  }
  const int n_tiles = ((a.N + TC_BM - 1) / TC_BM) * ((a.M + TN - 1) / TN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles < lc.n_sm ? n_tiles : lc.n_sm);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, tc_gemm_kernel<TN, EPI>, *reinterpret_cast<const CUtensorMap*>(&tmA),
                            *reinterpret_cast<const CUtensorMap*>(&tmB), a);
}

template <int EPI>
static cudaError_t launch_tc_e(int tn, const TcTensorMap& tmA, const TcTensorMap& tmB, const GemvArgs& a, const LaunchCfg& lc) {
  switch (tn) {
    case 16: return launch_tc_t<16, EPI>(tmA, tmB, a, lc);
    case 32: return launch_tc_t<32, EPI>(tmA, tmB, a, lc);
    case 64: return launch_tc_t<64, EPI>(tmA, tmB, a, lc);
    case 128: return launch_tc_t<128, EPI>(tmA, tmB, a, lc);
    case 256: return launch_tc_t<256, EPI>(tmA, tmB, a, lc);
  }
  return cudaErrorInvalidValue;
}

template <int TN, int EPI>
static cudaError_t launch_sk_t(const TcTensorMap& tmA, const TcTensorMap& tmB, const GemvArgs& a, const LaunchCfg& lc) {
  using Cfg = TcCfg<TN>;
  static std::atomic<unsigned long long> attr_mask{0};  // per instantiation, per device
  if (first_launch_on_device(attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_sk_kernel<TN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
  }
  const int nkb = (a.K + TC_BK - 1) / TC_BK;
  const long long U = (long long)((a.N + TC_BM - 1) / TC_BM) * nkb;
  const int grid = sk_grid(U, nkb, lc.n_sm);  // >= nkb/8 units per CTA: a tile spans at most 8 full + 2 partial ranges
  if (grid * 2 > lc.sk_slots || !sk_fits(U, grid)) return cudaErrorInvalidValue;
  const SkWs ws = {lc.sk_part, lc.sk_flags, lc.sk_prof};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, tc_gemm_sk_kernel<TN, EPI>, *reinterpret_cast<const CUtensorMap*>(&tmA),
                            *reinterpret_cast<const CUtensorMap*>(&tmB), a, ws);
}

template <int EPI>
static cudaError_t launch_sk_e(int tn, const TcTensorMap& tmA, const TcTensorMap& tmB, const GemvArgs& a, const LaunchCfg& lc) {
  switch (tn) {
    case 16: return launch_sk_t<16, EPI>(tmA, tmB, a, lc);
    case 32: return launch_sk_t<32, EPI>(tmA, tmB, a, lc);
    case 64: return launch_sk_t<64, EPI>(tmA, tmB, a, lc);
  }
  return cudaErrorInvalidValue;
}

static cudaError_t launch_tc_streamk(const TcTensorMap& tmA, const TcTensorMap& tmB, int tn, const GemvArgs& a, int epi, const LaunchCfg& lc) {
  switch (epi) {
    case EPI_QKV_ROPE: return launch_sk_e<EPI_QKV_ROPE>(tn, tmA, tmB, a, lc);
    case EPI_SWIGLU: return launch_sk_e<EPI_SWIGLU>(tn, tmA, tmB, a, lc);
    case EPI_RESID: return launch_sk_e<EPI_RESID>(tn, tmA, tmB, a, lc);
    case EPI_BF16: return launch_sk_e<EPI_BF16>(tn, tmA, tmB, a, lc);
    case EPI_F32: return launch_sk_e<EPI_F32>(tn, tmA, tmB, a, lc);
    case EPI_GELU: return launch_sk_e<EPI_GELU>(tn, tmA, tmB, a, lc);
    case EPI_RESID2: return launch_sk_e<EPI_RESID2>(tn, tmA, tmB, a, lc);
    case EPI_F32_BF16R: return launch_sk_e<EPI_F32_BF16R>(tn, tmA, tmB, a, lc);
  }
  return cudaErrorInvalidValue;
}

// tmB must have been built with box_rows == tn
cudaError_t launch_tc_gemm(const TcTensorMap& tmA, const TcTensorMap& tmB, int tn, const GemvArgs& a, int epi, const LaunchCfg& lc) {
  if ((a.N & 1) || (a.K & 7)) return cudaErrorInvalidValue;
  if (a.M <= tn && tn <= 64 && lc.sk_part && !a.row_map) return launch_tc_streamk(tmA, tmB, tn, a, epi, lc);  // batched decode
  switch (epi) {
    case EPI_QKV_ROPE: return launch_tc_e<EPI_QKV_ROPE>(tn, tmA, tmB, a, lc);
    case EPI_SWIGLU: return launch_tc_e<EPI_SWIGLU>(tn, tmA, tmB, a, lc);
    case EPI_RESID: return launch_tc_e<EPI_RESID>(tn, tmA, tmB, a, lc);
    case EPI_BF16: return launch_tc_e<EPI_BF16>(tn, tmA, tmB, a, lc);
    case EPI_F32: return launch_tc_e<EPI_F32>(tn, tmA, tmB, a, lc);
    case EPI_GELU: return launch_tc_e<EPI_GELU>(tn, tmA, tmB, a, lc);
    case EPI_RESID2: return launch_tc_e<EPI_RESID2>(tn, tmA, tmB, a, lc);
    case EPI_F32_BF16R: return launch_tc_e<EPI_F32_BF16R>(tn, tmA, tmB, a, lc);
  }
  return cudaErrorInvalidValue;
}

// LayerNorm (prefill): y = bf16((x - mean) * rstd * w + b), fp32 statistics   HF:models/falcon/modeling_falcon.py:572-576
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                                                        bf16* __restrict__ out, int K, float eps) {
  __shared__ float red[8];
  pdl_wait();
  pdl_launch_dependents();
  const int m = blockIdx.x, tid = threadIdx.x;
  const bf16* src = x + (size_t)m * K;
  float sm = 0.f;
  for (int k = tid * 8; k < K; k += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + k);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) sm += bf_lo(u[i]) + bf_hi(u[i]);
  }
  sm = warp_sum(sm);
  if ((tid & 31) == 0) red[tid >> 5] = sm;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  __syncthreads();
  const float mean = tot / (float)K;
  float sq = 0.f;
  for (int k = tid * 8; k < K; k += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + k);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = bf_lo(u[i]) - mean, hi = bf_hi(u[i]) - mean;
      sq += lo * lo + hi * hi;
    }
  }
  sq = warp_sum(sq);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / (float)K + eps);
  for (int k = tid * 8; k < K; k += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + k);
    const uint4 wv = ldg128(w + k);
    const uint4 bv = ldg128(b + k);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
    const uint32_t bu[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = pack_bf16((bf_lo(u[i]) - mean) * rstd * bf_lo(wu[i]) + bf_lo(bu[i]), (bf_hi(u[i]) - mean) * rstd * bf_hi(wu[i]) + bf_hi(bu[i]));
    *reinterpret_cast<uint4*>(out + (size_t)m * K + k) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t launch_layernorm(const bf16* x, const bf16* w, const bf16* b, bf16* out, int M, int K, float eps, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(M);
  cfg.blockDim = dim3(256);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, layernorm_kernel, x, w, b, out, K, eps);
}

// ------------------------------------------------------------------------------------------------ RMSNorm (prefill)
// xn[m][:] = w * bf16(x[m][:] * rsqrt(mean(x^2) + eps))     HF:models/llama/modeling_llama.py:62-67
__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ out,
                                                      int K, float eps) {
  __shared__ float red[8];
  pdl_wait();
  pdl_launch_dependents();
  const int m = blockIdx.x, tid = threadIdx.x;
  const bf16* src = x + (size_t)m * K;
  float ss = 0.f;
  for (int k = tid * 8; k < K; k += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + k);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = bf_lo(u[i]), hi = bf_hi(u[i]);
      ss += lo * lo + hi * hi;
    }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / (float)K + eps);
  for (int k = tid * 8; k < K; k += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(src + k);
    const uint4 wv = ldg128(w + k);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      o[i] = pack_bf16(bf16r(bf_lo(u[i]) * rstd) * bf_lo(wu[i]), bf16r(bf_hi(u[i]) * rstd) * bf_hi(wu[i]));
    *reinterpret_cast<uint4*>(out + (size_t)m * K + k) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

cudaError_t launch_rmsnorm(const bf16* x, const bf16* w, bf16* out, int M, int K, float eps, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(M);
  cfg.blockDim = dim3(256);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, rmsnorm_kernel, x, w, out, K, eps);
}
