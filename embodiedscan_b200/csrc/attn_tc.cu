// esb200 — multi-head attention core of the grounding decoder (text<->3D cross-attention, self-attention
// and text cross-attention of embodiedscan/models/layers/ground_transformer/decoder.py:89-95,151-177; 8 heads of 32 channels,
// 256 queries, up to ~3.5k keys, key padding mask) as flash-attention tiles on tcgen05 / TMEM fed by TMA, forward AND backward.
//
// forward  (one CTA = 128 queries of one (scan, head)): Q, K, V tiles arrive as 64B-swizzled TMA boxes; S = Q K^T is one
//   tcgen05.mma pair (M=128, N=128 keys, K=32) into TMEM; 128 softmax threads (thread = query row = TMEM lane) keep the running
//   maximum / sum in the exp2 domain, write P (bf16) as the K-major 128B-swizzled A operand; O_tile = P V is eight more MMAs
//   with V read as the MN-major B operand (no transpose); O is rescaled in registers. Stores O (bf16) and the log-sum-exp.
// backward (one CTA = 128 keys of one (scan, head), loop over query tiles): S^T = K Q^T and dP^T = V dO^T in TMEM, thread =
//   key row recomputes P^T and dS^T = P^T o (dP^T - delta) * scale, stores both as K-major A tiles; dV += P^T dO and
//   dK += dS^T Q accumulate in TMEM over all query tiles (dO, Q read MN-major); dQ_tile = dS K reads the SAME dS^T tile as the
//   MN-major A operand and K as the MN-major B operand, and is added to dQ (fp32) with 16-byte vector atomics.
// Layout: q (B,H,Lq,32), k / v (B,H,Lk,32) bf16 contiguous; key_pad (B,Lk) uint8, 1 = ignore; lse (B,H,Lq) natural log.
// Roofline: tensor pipe for long key sets, launch / latency bound at decoder sizes (256 x 3.5k x 32 per head).
#include "tc_common.cuh"

using namespace esb_tc;

namespace {

constexpr int AT_D = 32;                    // head dimension
constexpr int AT_THREADS = 192;             // warp 0 TMA, warp 1 MMA, warps 2-5 softmax / gradient math
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t piece) { return row * 128u + ((piece ^ (row & 7u)) << 4); }

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                const __grid_constant__ CUtensorMap tmv, const unsigned char* __restrict__ key_pad,
                __nv_bfloat16* __restrict__ o, float* __restrict__ lse, int H, int Lq, int Lk, float scale) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;                        // 128 x 64 B
  uint8_t* k_s = smem + 8192;                 // 2 stages x 8 KB
  uint8_t* v_s = smem + 8192 * 3;             // 2 stages x 8 KB
  uint8_t* p_s = smem + 8192 * 5;             // 2 chunks x 16 KB (offset 40960: 1024-aligned)
  uint64_t* bars = (uint64_t*)(p_s + 32768);
  uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *s_free = bars + 6,
           *p_full = bars + 7, *o_full = bars + 8;
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  unsigned char* kmask_all = (unsigned char*)(tmem_slot + 1);  // 2 x 128 flags (key tiles alternate buffers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / H;
  const int q0 = blockIdx.x * 128;
  const int n_kt = (Lk + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_s = tmem_base, t_o = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 8192);
      tma_load_2d(&tmq, q_full, smem_u32(q_s), 0, bh * Lq + q0);
      for (int j = 0; j < n_kt; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], 16384);
        tma_load_2d(&tmk, &kv_full[s], smem_u32(k_s + s * 8192), 0, bh * Lk + j * 128);
        tma_load_2d(&tmv, &kv_full[s], smem_u32(v_s + s * 8192), 0, bh * Lk + j * 128);
      }
    }
  } else if (warp == 1) {
    const uint32_t id_s = make_idesc(128, 128, 0, 0), id_o = make_idesc(128, AT_D, 0, 1);
    mbar_wait(q_full, 0);
    for (int j = 0; j < n_kt; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_full[s], (j >> 1) & 1);
      mbar_wait(s_free, (j & 1) ^ 1);                          // the softmax threads have read S of tile j-1 out of TMEM
      tc_fence_after();
      if (lane == 0) {
        const uint32_t qa = smem_u32(q_s), ka = smem_u32(k_s + s * 8192);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16(t_s, make_desc_sw(qa + kk * 32, 16, 512, 4), make_desc_sw(ka + kk * 32, 16, 512, 4), id_s, kk ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);                                // P of tile j is in shared memory, O_tile of j-1 consumed
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pa = smem_u32(p_s), va = smem_u32(v_s + s * 8192);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16(t_o, make_desc_sw(pa + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, 2),
                    make_desc_sw(va + kk * 1024, 2048, 512, 4), id_o, kk ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                            // query row of the tile = TMEM lane
    const int et = r;
    const float sc2 = scale * LOG2E;
    float m = -INFINITY, l = 0.f, acc[AT_D];
#pragma unroll
    for (int e = 0; e < AT_D; ++e) acc[e] = 0.f;
    for (int j = 0; j < n_kt; ++j) {
      unsigned char* kmask = kmask_all + (j & 1) * 128;
      {   // key mask of this tile (identical for every query row)
        const int kj = j * 128 + et;
        kmask[et] = (kj >= Lk || (key_pad != nullptr && key_pad[(long long)b * Lk + kj])) ? 1 : 0;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1 over the S tile in TMEM: running maximum (TMEM reads are cheap; keeping 128 scores in registers is not)
      float mx = m;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (!kmask[c0 + i]) mx = fmaxf(mx, __uint_as_float(v[i]) * sc2);
      }
      const float m_use = mx == -INFINITY ? 0.f : mx;
      const float alpha = exp2f(m - m_use);                    // m = -inf -> 0
      float rs = 0.f;
      // pass 2: probabilities -> bf16 A operand in shared memory
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_s + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
        if (c0 == 96) {
          tc_fence_before();
          mbar_arrive(s_free);
        }
#pragma unroll
        for (int pj = 0; pj < 4; ++pj) {                       // 4 pieces of 8 keys
          uint32_t w[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int i0 = pj * 8 + 2 * h;
            const float p0 = kmask[c0 + i0] ? 0.f : exp2f(__uint_as_float(v[i0]) * sc2 - m_use);
            const float p1 = kmask[c0 + i0 + 1] ? 0.f : exp2f(__uint_as_float(v[i0 + 1]) * sc2 - m_use);
            rs += p0 + p1;
            w[h] = pack_bf16(__float_as_uint(p0), __float_as_uint(p1));
          }
          *reinterpret_cast<uint4*>(p_s + (c0 >> 6) * 16384 + swz128((uint32_t)r, (uint32_t)(((c0 & 63) >> 3) + pj))) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      l = l * alpha + rs;
      m = mx;
      fence_proxy_async();
      mbar_arrive(p_full);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      uint32_t ov[32];
      tmem_ld32(t_o + ((uint32_t)(quad * 32) << 16), ov);
      tc_fence_before();
#pragma unroll
      for (int e = 0; e < AT_D; ++e) acc[e] = acc[e] * alpha + __uint_as_float(ov[e]);
    }
    if (q0 + r < Lq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      uint32_t w[16];
#pragma unroll
      for (int h = 0; h < 16; ++h) w[h] = pack_bf16(__float_as_uint(acc[2 * h] * inv), __float_as_uint(acc[2 * h + 1] * inv));
      uint4* dst = reinterpret_cast<uint4*>(o + ((long long)bh * Lq + q0 + r) * AT_D);
#pragma unroll
      for (int h = 0; h < 4; ++h) dst[h] = make_uint4(w[4 * h], w[4 * h + 1], w[4 * h + 2], w[4 * h + 3]);
      lse[(long long)bh * Lq + q0 + r] = l > 0.f ? (m + log2f(l)) * 0.6931471805599453f : -INFINITY;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// delta[bh, q] = sum_d dO * O  (fp32)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, long long rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float s = 0.f;
  const uint4* a = reinterpret_cast<const uint4*>(o + i * AT_D);
  const uint4* g = reinterpret_cast<const uint4*>(dout + i * AT_D);
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const uint4 x = a[h], y = g[h];
    const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* yp = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 xf = __bfloat1622float2(xp[e]), yf = __bfloat1622float2(yp[e]);
      s = fmaf(xf.x, yf.x, s);
      s = fmaf(xf.y, yf.y, s);
    }
  }
  delta[i] = s;
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
                const unsigned char* __restrict__ key_pad, const float* __restrict__ lse, const float* __restrict__ delta,
                float* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int H, int Lq, int Lk,
                float scale) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* k_s = smem;                        // 8 KB
  uint8_t* v_s = smem + 8192;                 // 8 KB
  uint8_t* q_s = smem + 16384;                // 2 stages x 8 KB
  uint8_t* do_s = smem + 32768;               // 2 stages x 8 KB
  uint8_t* pt_s = smem + 49152;               // P^T: 2 chunks x 16 KB (1024-aligned)
  uint8_t* ds_s = smem + 49152 + 32768;       // dS^T
  float* lse2_all = (float*)(smem + 49152 + 65536);        // 2 x 128: lse * log2e of the query tile (tiles alternate buffers)
  float* delta_all = lse2_all + 256;
  uint64_t* bars = (uint64_t*)(delta_all + 256);
  uint64_t *kv_full = bars, *qd_full = bars + 1, *qd_empty = bars + 3, *sdp_full = bars + 5, *st_free = bars + 6,
           *pds_full = bars + 7, *dq_full = bars + 8, *dq_free = bars + 9, *fin_full = bars + 10;
  uint32_t* tmem_slot = (uint32_t*)(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / H;
  const int k0 = blockIdx.x * 128;
  const int n_qt = (Lq + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&qd_full[s], 1); mbar_init(&qd_empty[s], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(st_free, 128);
    mbar_init(pds_full, 128);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
    mbar_init(fin_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmq); tma_prefetch_desc(&tmk); tma_prefetch_desc(&tmv); tma_prefetch_desc(&tmdo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_st = tmem_base, t_dp = tmem_base + 128, t_dv = tmem_base + 256, t_dk = tmem_base + 288, t_dq = tmem_base + 320;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 16384);
      tma_load_2d(&tmk, kv_full, smem_u32(k_s), 0, bh * Lk + k0);
      tma_load_2d(&tmv, kv_full, smem_u32(v_s), 0, bh * Lk + k0);
      for (int i = 0; i < n_qt; ++i) {
        const int s = i & 1;
        mbar_wait(&qd_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qd_full[s], 16384);
        tma_load_2d(&tmq, &qd_full[s], smem_u32(q_s + s * 8192), 0, bh * Lq + i * 128);
        tma_load_2d(&tmdo, &qd_full[s], smem_u32(do_s + s * 8192), 0, bh * Lq + i * 128);
      }
    }
  } else if (warp == 1) {
    const uint32_t id_s = make_idesc(128, 128, 0, 0);          // S^T, dP^T: A, B K-major
    const uint32_t id_g = make_idesc(128, AT_D, 0, 1);         // dV, dK: A K-major, B MN-major
    const uint32_t id_q = make_idesc(128, AT_D, 1, 1);         // dQ: A MN-major (dS^T read transposed), B MN-major
    mbar_wait(kv_full, 0);
    for (int i = 0; i < n_qt; ++i) {
      const int s = i & 1;
      mbar_wait(&qd_full[s], (i >> 1) & 1);
      mbar_wait(st_free, (i & 1) ^ 1);
      tc_fence_after();
      const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s), qa = smem_u32(q_s + s * 8192), da = smem_u32(do_s + s * 8192);
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16(t_st, make_desc_sw(ka + kk * 32, 16, 512, 4), make_desc_sw(qa + kk * 32, 16, 512, 4), id_s, kk ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_bf16(t_dp, make_desc_sw(va + kk * 32, 16, 512, 4), make_desc_sw(da + kk * 32, 16, 512, 4), id_s, kk ? 1u : 0u);
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, i & 1);                              // P^T and dS^T of this query tile are in shared memory
      mbar_wait(dq_free, (i & 1) ^ 1);                         // dQ of the previous tile has been read out of TMEM
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pa = smem_u32(pt_s), sa = smem_u32(ds_s);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {                       // reduction over the 128 queries of the tile
          const uint32_t a_off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_bf16(t_dv, make_desc_sw(pa + a_off, 16, 1024, 2), make_desc_sw(da + kk * 1024, 2048, 512, 4), id_g,
                    (i > 0 || kk > 0) ? 1u : 0u);
          umma_bf16(t_dk, make_desc_sw(sa + a_off, 16, 1024, 2), make_desc_sw(qa + kk * 1024, 2048, 512, 4), id_g,
                    (i > 0 || kk > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)                         // reduction over the 128 keys of this CTA
          umma_bf16(t_dq, make_desc_sw(sa + kk * 2048, 16384, 1024, 2), make_desc_sw(ka + kk * 1024, 2048, 512, 4), id_q,
                    kk ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&qd_empty[s]);
        if (i == n_qt - 1) umma_commit(fin_full);
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                            // key row of this CTA (S^T lanes) / query row (dQ lanes)
    const int kj = k0 + r;
    const bool k_dead = kj >= Lk || (key_pad != nullptr && key_pad[(long long)b * Lk + kj]);
    const float sc2 = scale * LOG2E;
    for (int i = 0; i < n_qt; ++i) {
      float* lse2_s = lse2_all + (i & 1) * 128;
      float* delta_s = delta_all + (i & 1) * 128;
      {
        const int qi = i * 128 + r;
        lse2_s[r] = qi < Lq ? lse[(long long)bh * Lq + qi] * LOG2E : INFINITY;     // +inf -> p = 0 for absent queries
        delta_s[r] = qi < Lq ? delta[(long long)bh * Lq + qi] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t sv[32], dv_[32];
        tmem_ld32(t_st + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, sv);
        tmem_ld32(t_dp + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, dv_);
        if (c0 == 96) {
          tc_fence_before();
          mbar_arrive(st_free);
        }
#pragma unroll
        for (int pj = 0; pj < 4; ++pj) {
          uint32_t wp[4], wd[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float p[2], ds[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int c = c0 + pj * 8 + 2 * h + u;
              const float e = k_dead ? 0.f : exp2f(__uint_as_float(sv[pj * 8 + 2 * h + u]) * sc2 - lse2_s[c]);
              p[u] = e;
              ds[u] = e * (__uint_as_float(dv_[pj * 8 + 2 * h + u]) - delta_s[c]) * scale;
            }
            wp[h] = pack_bf16(__float_as_uint(p[0]), __float_as_uint(p[1]));
            wd[h] = pack_bf16(__float_as_uint(ds[0]), __float_as_uint(ds[1]));
          }
          const uint32_t off = (uint32_t)(c0 >> 6) * 16384u + swz128((uint32_t)r, (uint32_t)(((c0 & 63) >> 3) + pj));
          *reinterpret_cast<uint4*>(pt_s + off) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
          *reinterpret_cast<uint4*>(ds_s + off) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        }
      }
      fence_proxy_async();
      mbar_arrive(pds_full);
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t qv[32];
      tmem_ld32(t_dq + ((uint32_t)(quad * 32) << 16), qv);
      tc_fence_before();
      mbar_arrive(dq_free);
      const int qi = i * 128 + r;
      if (qi < Lq) {
        float* dst = dq + ((long long)bh * Lq + qi) * AT_D;
#pragma unroll
        for (int h = 0; h < 8; ++h)
          atomicAdd(reinterpret_cast<float4*>(dst + 4 * h),
                    make_float4(__uint_as_float(qv[4 * h]), __uint_as_float(qv[4 * h + 1]), __uint_as_float(qv[4 * h + 2]),
                                __uint_as_float(qv[4 * h + 3])));
      }
    }
    mbar_wait(fin_full, 0);
    tc_fence_after();
    uint32_t a[32], c[32];
    tmem_ld32(t_dv + ((uint32_t)(quad * 32) << 16), a);
    tmem_ld32(t_dk + ((uint32_t)(quad * 32) << 16), c);
    tc_fence_before();
    if (kj < Lk) {
      uint4* pv = reinterpret_cast<uint4*>(dv + ((long long)bh * Lk + kj) * AT_D);
      uint4* pk = reinterpret_cast<uint4*>(dk + ((long long)bh * Lk + kj) * AT_D);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        pv[h] = make_uint4(pack_bf16(a[8 * h], a[8 * h + 1]), pack_bf16(a[8 * h + 2], a[8 * h + 3]), pack_bf16(a[8 * h + 4], a[8 * h + 5]),
                           pack_bf16(a[8 * h + 6], a[8 * h + 7]));
        pk[h] = make_uint4(pack_bf16(c[8 * h], c[8 * h + 1]), pack_bf16(c[8 * h + 2], c[8 * h + 3]), pack_bf16(c[8 * h + 4], c[8 * h + 5]),
                           pack_bf16(c[8 * h + 6], c[8 * h + 7]));
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int rows32_map(CUtensorMap* tm, const void* base, long long rows) {      // (rows, 32) bf16: boxes of 128 rows, 64B swizzle
  unsigned long long dims[2] = {AT_D, (unsigned long long)(rows > 0 ? rows : 1)};
  unsigned long long str[1] = {AT_D * 2};
  unsigned box[2] = {AT_D, 128};
  return esb_tma_encode(tm, base, 2, dims, str, box, nullptr, 64);
}

}  // namespace

// o (B,H,Lq,32) = softmax(q k^T * scale + key padding) v ; lse (B,H,Lq) fp32 = log-sum-exp of the scaled scores (-inf for a query
// whose keys are all padded: its output row is zero).
extern "C" int esb_attn_fwd(const void* q, const void* k, const void* v, const unsigned char* key_pad, void* o, float* lse, int B,
                            int H, int Lq, int Lk, float scale, void* stream_) {
  ESB_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, "esb_attn_fwd: empty problem");
  CUtensorMap tmq, tmk, tmv;
  int rc = rows32_map(&tmq, q, (long long)B * H * Lq);
  if (rc == ESB_OK) rc = rows32_map(&tmk, k, (long long)B * H * Lk);
  if (rc == ESB_OK) rc = rows32_map(&tmv, v, (long long)B * H * Lk);
  if (rc != ESB_OK) return rc;
  const size_t smem = 8192 * 5 + 32768 + 16 * 8 + 16 + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { esb_set_error("esb_attn_fwd: smem attr: %s", cudaGetErrorString(e)); return ESB_ECUDA; }
  dim3 grid((Lq + 127) / 128, B * H);
  attn_fwd_kernel<<<grid, AT_THREADS, smem, (cudaStream_t)stream_>>>(tmq, tmk, tmv, key_pad, (__nv_bfloat16*)o, lse, H, Lq, Lk, scale);
  ESB_CUDA_LAUNCH_CHECK("attn_fwd_kernel");
  return ESB_OK;
}

// dq (B,H,Lq,32) fp32 ZEROED BY THE CALLER (key tiles accumulate with vector atomics); dk, dv (B,H,Lk,32) bf16 written once;
// delta (B,H,Lq) fp32 workspace.
extern "C" int esb_attn_bwd(const void* q, const void* k, const void* v, const unsigned char* key_pad, const void* o,
                            const void* dout, const float* lse, float* delta, float* dq, void* dk, void* dv, int B, int H, int Lq,
                            int Lk, float scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(B > 0 && H > 0 && Lq > 0 && Lk > 0, "esb_attn_bwd: empty problem");
  CUtensorMap tmq, tmk, tmv, tmdo;
  int rc = rows32_map(&tmq, q, (long long)B * H * Lq);
  if (rc == ESB_OK) rc = rows32_map(&tmk, k, (long long)B * H * Lk);
  if (rc == ESB_OK) rc = rows32_map(&tmv, v, (long long)B * H * Lk);
  if (rc == ESB_OK) rc = rows32_map(&tmdo, dout, (long long)B * H * Lq);
  if (rc != ESB_OK) return rc;
  const long long rows = (long long)B * H * Lq;
  attn_delta_kernel<<<esb_div_up(rows, 256), 256, 0, stream>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)dout, delta, rows);
  ESB_CUDA_LAUNCH_CHECK("attn_delta_kernel");
  const size_t smem = 49152 + 65536 + 2048 + 16 * 8 + 16 + 1024;
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { esb_set_error("esb_attn_bwd: smem attr: %s", cudaGetErrorString(e)); return ESB_ECUDA; }
  dim3 grid((Lk + 127) / 128, B * H);
  attn_bwd_kernel<<<grid, AT_THREADS, smem, stream>>>(tmq, tmk, tmv, tmdo, key_pad, lse, delta, dq, (__nv_bfloat16*)dk,
                                                      (__nv_bfloat16*)dv, H, Lq, Lk, scale);
  ESB_CUDA_LAUNCH_CHECK("attn_bwd_kernel");
  return ESB_OK;
}
