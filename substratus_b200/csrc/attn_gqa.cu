// attn_gqa.cu — paged-KV decode attention for grouped-query models (G = 4 or 8 query heads per KV head,
// e.g. Llama-2-70B 64/8, Falcon-40B 128/8).  One CTA = one (row, KV head [, head sub-group], context split);
// the K/V rows of the split are staged ONCE in shared memory (cp.async 16 B per thread, all loads in flight at once),
// then warp g computes query head g against the staged tile: dot via 16-lane shuffle reduction, fp32 online softmax,
// fp32 P.V — so the K/V stream is read once per group instead of once per head and no thread carries G accumulators.
// Splits are merged by the last-arriving CTA (same scheme and buffers as attn_decode_kernel).
// Math/rounding: HF:models/llama/modeling_llama.py:187-221 (repeat_kv + eager_attention_forward), see kernels.cu.
#include "common.cuh"
#include "kernels.h"

constexpr int GQ_TOK = 64;  // tokens staged per pass

SSB_DEVINL void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
SSB_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int D, int G>
__global__ void __launch_bounds__(G * 32) attn_gqa_kernel(const AttnArgs a) {
  constexpr int LPR = D / 8, RPW = 32 / LPR;
  constexpr int NT = G * 32;
  __shared__ __align__(16) bf16 sK[GQ_TOK][D];
  __shared__ __align__(16) bf16 sV[GQ_TOK][D];
  __shared__ int sm_last;

  pdl_wait();
  pdl_launch_dependents();
  const int split = blockIdx.x, row = blockIdx.z;
  const int kvh = blockIdx.y / (a.group / G);
  const int head0 = blockIdx.y * G;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const int slot = a.row_slot[row];
  const int ctx = a.row_pos[row] + 1;
  const int BS = a.block_size;
  int chunk = (ctx + a.n_splits - 1) / a.n_splits;
  chunk = ((chunk + BS - 1) / BS) * BS;
  const int n_active = (ctx + chunk - 1) / chunk;
  if (split >= n_active) return;
  const int t_begin = split * chunk;
  const int t_end = min(ctx, t_begin + chunk);
  const int HD = a.n_heads * D;
  const int* bt = a.block_table + (size_t)slot * a.bt_stride;

  float q[8];
  {
    const uint4 v = *reinterpret_cast<const uint4*>(a.q + (size_t)row * HD + (head0 + warp) * D + li * 8);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      q[2 * i] = bf_lo(u[i]);
      q[2 * i + 1] = bf_hi(u[i]);
    }
  }
  float m = -1e30f, l = 0.f, acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  for (int p0 = t_begin; p0 < t_end; p0 += GQ_TOK) {
    const int np = min(GQ_TOK, t_end - p0);
    if (p0 != t_begin) __syncthreads();  // previous tile fully consumed
    for (int i = tid; i < np * LPR; i += NT) {
      const int tt = i / LPR, c = i % LPR;
      const int t = p0 + tt;
      const size_t off = (((size_t)bt[t / BS] * a.kvh + kvh) * BS + (t % BS)) * D + c * 8;
      cp_async16(&sK[tt][c * 8], a.kcache + off);
      cp_async16(&sV[tt][c * 8], a.vcache + off);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int tb = 0; tb < np; tb += RPW) {  // warp-uniform trip count (full-mask shuffles below)
      const int tt = tb + sub;
      const bool tv = tt < np;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (tv) {
        kv = *reinterpret_cast<const uint4*>(&sK[tt][li * 8]);
        vv = *reinterpret_cast<const uint4*>(&sV[tt][li * 8]);
      }
      const uint32_t ku[4] = {kv.x, kv.y, kv.z, kv.w};
      const uint32_t vu[4] = {vv.x, vv.y, vv.z, vv.w};
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        d = fmaf(q[2 * i], bf_lo(ku[i]), d);
        d = fmaf(q[2 * i + 1], bf_hi(ku[i]), d);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (!tv) continue;
      const float s = bf16r(bf16r(d) * a.scale);
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      m = mn;
      l = l * corr + p;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] = fmaf(p, bf_lo(vu[i]), acc[2 * i] * corr);
        acc[2 * i + 1] = fmaf(p, bf_hi(vu[i]), acc[2 * i + 1] * corr);
      }
    }
  }
  // merge the RPW token sub-groups of the warp
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const float ol = __shfl_xor_sync(0xffffffffu, l, o);
    const float mn = fmaxf(m, om);
    const float wa = __expf(m - mn), wb = __expf(om - mn);
    l = l * wa + ol * wb;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float oa = __shfl_xor_sync(0xffffffffu, acc[i], o);
      acc[i] = acc[i] * wa + oa * wb;
    }
    m = mn;
  }
  const size_t pbase = ((size_t)row * gridDim.y + blockIdx.y) * a.n_splits;
  if (n_active == 1) {
    if (sub == 0) {
      const float inv = 1.0f / l;
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = pack_bf16(acc[2 * i] * inv, acc[2 * i + 1] * inv);
      *reinterpret_cast<uint4*>(a.out + (size_t)row * HD + (head0 + warp) * D + li * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    return;
  }
  if (sub == 0) {
    float* po = a.part_o + (pbase + split) * (G * D) + warp * D + li * 8;
    *reinterpret_cast<float4*>(po) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(po + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (li == 0) {
      a.part_ml[(pbase + split) * (2 * G) + warp] = m;
      a.part_ml[(pbase + split) * (2 * G) + G + warp] = l;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sm_last = (atomicAdd(&a.counters[row * gridDim.y + blockIdx.y], 1) == n_active - 1);
  __syncthreads();
  if (!sm_last) return;
  __threadfence();
  // warp g merges head g over the splits; lane handles 4 dims (D = 128) or lanes 0..15 (D = 64)
  for (int dd = lane * 4; dd < D; dd += 128) {
    float M2 = -1e30f;
    for (int s = 0; s < n_active; ++s) M2 = fmaxf(M2, __ldcg(&a.part_ml[(pbase + s) * (2 * G) + warp]));
    float L2 = 0.f;
    float4 O = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n_active; ++s) {
      const float w = __expf(__ldcg(&a.part_ml[(pbase + s) * (2 * G) + warp]) - M2);
      L2 += __ldcg(&a.part_ml[(pbase + s) * (2 * G) + G + warp]) * w;
      const float4 o = __ldcg(reinterpret_cast<const float4*>(a.part_o + (pbase + s) * (G * D) + warp * D + dd));
      O.x += o.x * w;
      O.y += o.y * w;
      O.z += o.z * w;
      O.w += o.w * w;
    }
    const float inv = 1.0f / L2;
    uint2 o;
    o.x = pack_bf16(O.x * inv, O.y * inv);
    o.y = pack_bf16(O.z * inv, O.w * inv);
    *reinterpret_cast<uint2*>(a.out + (size_t)row * HD + (head0 + warp) * D + dd) = o;
  }
  if (tid == 0) a.counters[row * gridDim.y + blockIdx.y] = 0;
}

template <int D, int G>
static cudaError_t launch_gqa_t(const AttnArgs& a, const LaunchCfg& lc) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.n_splits, a.kvh * (a.group / G), a.M);
  cfg.blockDim = dim3(G * 32);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn_gqa_kernel<D, G>, a);
}

// gc = query heads per CTA (4 or 8)
cudaError_t launch_attn_gqa(const AttnArgs& a, int gc, const LaunchCfg& lc) {
  if (a.head_dim == 128) {
    if (gc == 8) return launch_gqa_t<128, 8>(a, lc);
    if (gc == 4) return launch_gqa_t<128, 4>(a, lc);
  } else if (a.head_dim == 64) {
    if (gc == 8) return launch_gqa_t<64, 8>(a, lc);
    if (gc == 4) return launch_gqa_t<64, 4>(a, lc);
  }
  return cudaErrorInvalidValue;
}

// =====================================================================================================================
// prefill attention: a CTA owns a tile of up to 16 consecutive query rows of one sequence and one query head; the
// causal key range [0, pos_last] is streamed through shared memory in 64-token chunks (K and V staged once per chunk
// for all 16 queries instead of once per query), each warp carries two query rows with fp32 online softmax.
// Same rounding points as the decode kernels (scores rounded to bf16, scaled in bf16, fp32 softmax, fp32 P.V).
// HF:models/llama/modeling_llama.py:199-221 with the causal mask of :399-406.
// =====================================================================================================================
constexpr int PF_ROWS = 16, PF_WARPS = 8;

template <int D>
__global__ void __launch_bounds__(PF_WARPS * 32) attn_prefill_kernel(const AttnArgs a) {
  constexpr int LPR = D / 8, RPW = 32 / LPR;
  __shared__ __align__(16) bf16 sK[GQ_TOK][D];
  __shared__ __align__(16) bf16 sV[GQ_TOK][D];
  pdl_wait();
  pdl_launch_dependents();
  const int tile = blockIdx.x, head = blockIdx.y;
  const int kvh = head / a.group;
  const int row0 = a.tile_row0[tile], nrows = a.tile_nrows[tile];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane / LPR, li = lane % LPR;
  const int slot = a.row_slot[row0];
  const int BS = a.block_size;
  const int HD = a.n_heads * D;
  const int* bt = a.block_table + (size_t)slot * a.bt_stride;
  const int ctx = a.row_pos[row0 + nrows - 1] + 1;  // keys needed by the last row of the tile

  int pos[2];
  float q[2][8], m[2], l[2], acc[2][8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int rr = warp * 2 + r;
    pos[r] = rr < nrows ? a.row_pos[row0 + rr] : -1;
    m[r] = -1e30f;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r][i] = 0.f, q[r][i] = 0.f;
    if (rr < nrows) {
      const uint4 v = *reinterpret_cast<const uint4*>(a.q + (size_t)(row0 + rr) * HD + head * D + li * 8);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q[r][2 * i] = bf_lo(u[i]);
        q[r][2 * i + 1] = bf_hi(u[i]);
      }
    }
  }
  for (int p0 = 0; p0 < ctx; p0 += GQ_TOK) {
    const int np = min(GQ_TOK, ctx - p0);
    if (p0) __syncthreads();
    for (int i = tid; i < np * LPR; i += PF_WARPS * 32) {
      const int tt = i / LPR, c = i % LPR;
      const int t = p0 + tt;
      const size_t off = (((size_t)bt[t / BS] * a.kvh + kvh) * BS + (t % BS)) * D + c * 8;
      cp_async16(&sK[tt][c * 8], a.kcache + off);
      cp_async16(&sV[tt][c * 8], a.vcache + off);
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int lim = min(np, pos[r] - p0 + 1);  // keys of this chunk visible to row r (<= 0: none); warp-uniform
      for (int tb = 0; tb < lim; tb += RPW) {
        const int tt = tb + sub;
        const bool tv = tt < lim;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
        if (tv) {
          kv = *reinterpret_cast<const uint4*>(&sK[tt][li * 8]);
          vv = *reinterpret_cast<const uint4*>(&sV[tt][li * 8]);
        }
        const uint32_t ku[4] = {kv.x, kv.y, kv.z, kv.w};
        const uint32_t vu[4] = {vv.x, vv.y, vv.z, vv.w};
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          d = fmaf(q[r][2 * i], bf_lo(ku[i]), d);
          d = fmaf(q[r][2 * i + 1], bf_hi(ku[i]), d);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (!tv) continue;
        const float s = bf16r(bf16r(d) * a.scale);
        const float mn = fmaxf(m[r], s);
        const float corr = __expf(m[r] - mn), p = __expf(s - mn);
        m[r] = mn;
        l[r] = l[r] * corr + p;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[r][2 * i] = fmaf(p, bf_lo(vu[i]), acc[r][2 * i] * corr);
          acc[r][2 * i + 1] = fmaf(p, bf_hi(vu[i]), acc[r][2 * i + 1] * corr);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m[r], o);
      const float ol = __shfl_xor_sync(0xffffffffu, l[r], o);
      const float mn = fmaxf(m[r], om);
      const float wa = __expf(m[r] - mn), wb = __expf(om - mn);
      l[r] = l[r] * wa + ol * wb;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float oa = __shfl_xor_sync(0xffffffffu, acc[r][i], o);
        acc[r][i] = acc[r][i] * wa + oa * wb;
      }
      m[r] = mn;
    }
    const int rr = warp * 2 + r;
    if (rr < nrows && sub == 0) {
      const float inv = 1.0f / l[r];
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = pack_bf16(acc[r][2 * i] * inv, acc[r][2 * i + 1] * inv);
      *reinterpret_cast<uint4*>(a.out + (size_t)(row0 + rr) * HD + head * D + li * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

cudaError_t launch_attn_prefill(const AttnArgs& a, const LaunchCfg& lc) {
  if (a.n_tiles <= 0) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.n_tiles, a.n_heads);
  cfg.blockDim = dim3(PF_WARPS * 32);
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = lc.pdl ? 1 : 0;
  if (a.head_dim == 128) return cudaLaunchKernelEx(&cfg, attn_prefill_kernel<128>, a);
  if (a.head_dim == 64) return cudaLaunchKernelEx(&cfg, attn_prefill_kernel<64>, a);
  return cudaErrorInvalidValue;
}
