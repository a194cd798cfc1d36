// esb200 — sparse 3D convolution on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 in / fp32 accumulate.
// The throughput path for ME.MinkowskiConvolution forward / dgrad / wgrad (†upstream MinkowskiEngine; call sites
// embodiedscan/models/backbones/mink_resnet.py:58-62,104-108, embodiedscan/models/dense_heads/fcaf3d_head.py:919-946).
//
// forward / dgrad (spconv_tc_fwd_kernel): output-stationary implicit GEMM. A CTA owns 128 output rows x N_TILE output
// channels; for every kernel offset k that any row of the tile uses (per-tile bit mask) and every 64-channel slice of
// Cin it stages
//     A = the 128 gathered neighbour rows (cp.async 16 B, zero-filled where the neighbour is absent) and
//     B = W_k^T slice (N_TILE x 64, K-major)
// into 128B-swizzled shared memory; one elected thread issues tcgen05.mma (M=128, N=N_TILE, K=16) accumulating in
// TMEM; the epilogue reads TMEM with tcgen05.ld and writes each output row once (no atomics, deterministic).
// Warp roles: warps 0-3 = gather producers, then epilogue (TMEM lane quadrant = warp id); warp 4 = TMEM allocator +
// MMA issuer. Full/empty mbarrier ring between producers and the MMA thread; tcgen05.commit releases stages.
//
// wgrad (spconv_tc_wgrad_kernel): dW_k = X_k^T dY_k over the compacted pair list of offset k. Both operands are the
// gathered rows themselves, i.e. MN-major (M = Cin, N = Cout contiguous, reduction over pairs), staged in the
// canonical MN-major 128B-swizzle layout; split over pair ranges with fp32 atomics into dW.
//
// Roofline: pair model bytes = P*(Cin+Cout)*2 + 8P + K*Cin*Cout*2; the gather is L2-fed.
#include "tc_common.cuh"
#include <stdlib.h>

using namespace esb_tc;

namespace {

// per-tile (128 rows) bit mask of the kernel offsets that have at least one valid neighbour
__global__ void tile_mask_kernel(const int* __restrict__ nbr, int K, int n, uint32_t* __restrict__ masks) {
  const int tile = blockIdx.x, r = threadIdx.x;
  const int row = tile * TC_M + r;
  uint32_t m = 0;
  if (row < n)
    for (int k = 0; k < K; ++k)
      if (nbr[(long long)k * n + row] >= 0) m |= 1u << k;
  m = __reduce_or_sync(0xffffffffu, m);
  __shared__ uint32_t s[4];
  if ((r & 31) == 0) s[r >> 5] = m;
  __syncthreads();
  if (r == 0) masks[tile] = s[0] | s[1] | s[2] | s[3];
}

// ------------------------------------------------------------------------------------------------------------
// forward / dgrad
// ------------------------------------------------------------------------------------------------------------
// B_MN = false: wt is (K, cout, cin)  (W_k^T, reduction dim contiguous  -> K-major B; used by dgrad with wt = W itself)
// B_MN = true : wt is (K, cin, cout)  (W_k as stored, output dim contiguous -> MN-major B; used by forward, no transpose)
// MT = 128-row tiles per CTA: with MT = 2 one filter stage feeds two accumulators (half the filter traffic per output row:
// at C = 64 the filter re-reads were as large as the gather itself) and 256 producer threads gather, one tile each.
// The filter tile of every stage is ONE tiled TMA box (cp.async.bulk.tensor) issued by thread 0 — the 128B-swizzled layouts
// tcgen05 wants are what the tensor map writes; the gather stays on cp.async (tile::gather4 was measured 3x slower, see
// spconv_tma.cu).
template <int N_TILE, int STAGES, bool B_MN, int MT>
__global__ void __launch_bounds__(MT * 128 + 32)
spconv_tc_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __grid_constant__ CUtensorMap tmw,
                     const int* __restrict__ nbr, const uint32_t* __restrict__ masks, __nv_bfloat16* __restrict__ y,
                     int n_out, int cin, int cout, int K) {
  constexpr int A_BYTES = MT * A_STAGE_BYTES;
  constexpr int B_STAGE_BYTES = N_TILE * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  constexpr int LAG = STAGES - 2;                 // stages kept in flight per producer thread before it signals
  constexpr int NPROD = MT * 128;
  constexpr int TMEM_COLS = MT * N_TILE < 32 ? 32 : MT * N_TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);
  __shared__ int idx_s[2][NPROD];

  const int warp = threadIdx.x >> 5;
  const int tile = blockIdx.x;                     // MT * 128 output rows
  const int n0 = blockIdx.y * N_TILE;
  uint32_t mask = masks[tile * MT];
  if (MT == 2 && (tile * 2 + 1) * TC_M < n_out) mask |= masks[tile * 2 + 1];
  const int nchunk = cin / TC_BK;
  const int total = __popc(mask) * nchunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], NPROD + 1);          // every producer thread + the expect_tx arrival of the filter box
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmw);
  }
  if (warp == NPROD / 32) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NPROD / 32) {
    // ---------------- producers ----------------
    // Lane mapping: 8 consecutive lanes fetch the 8 x 16 B chunks of ONE gathered row, so a warp-wide cp.async touches
    // 4 rows x 128 B (4 L1 wavefronts) instead of 32 different rows (32 wavefronts: the round-1a L1TEX bottleneck).
    // Thread t of tile m = t >> 7 owns chunk (t & 7) of rows ((t & 127) >> 3) + 16 i, i = 0..7.
    const int t = threadIdx.x;
    const int tm = t >> 7, tl = t & 127;
    const int sub = tl & 7, rgrp = tl >> 3;
    const uint32_t sw = (uint32_t)(rgrp & 7);
    const uint32_t a_thread_off = (uint32_t)(tm * A_STAGE_BYTES + (rgrp >> 3) * 1024 + (rgrp & 7) * 128) + ((sub ^ sw) << 4);
    int it = 0, kcount = 0;
    const int my_row = tile * MT * TC_M + t;
    // the neighbour index of the NEXT offset is loaded while the stages of the current one are in flight
    int v_next = (mask && my_row < n_out) ? nbr[(long long)(__ffs(mask) - 1) * n_out + my_row] : -1;
    for (uint32_t mk = mask; mk; mk &= mk - 1, ++kcount) {
      const int k = __ffs(mk) - 1;
      int* idx_buf = idx_s[kcount & 1];
      idx_buf[t] = v_next;
      asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");     // producers only
      {
        const uint32_t rest = mk & (mk - 1);
        v_next = (rest && my_row < n_out) ? nbr[(long long)(__ffs(rest) - 1) * n_out + my_row] : -1;
      }
      int src[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) src[i] = idx_buf[tm * 128 + rgrp + 16 * i];
      for (int c = 0; c < nchunk; ++c, ++it) {
        const int s = it % STAGES;
        if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
        const uint32_t stage_base = smem_u32(smem + s * STAGE_BYTES);
        if (t == 0) {        // the filter tile of this stage: one (B_MN: N_TILE / 64) tiled TMA box(es)
          mbar_expect_tx(&full_bar[s], B_STAGE_BYTES);
          const uint32_t b_base = stage_base + A_BYTES;
          if (B_MN) {        // stored kernel (K*cin rows, cout columns): 64 reduction rows x 64 columns per box
#pragma unroll
            for (int a = 0; a < N_TILE / 64; ++a) tma_load_2d(&tmw, &full_bar[s], b_base + a * 8192, n0 + a * 64, k * cin + c * TC_BK);
          } else {           // (K*cout rows, cin columns): N_TILE output rows x 64 reduction columns
            tma_load_2d(&tmw, &full_bar[s], b_base, c * TC_BK, k * cout + n0);
          }
        }
        const uint32_t a_base = stage_base + a_thread_off;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          cp_async16_ca(a_base + i * 2048, x + (long long)(src[i] >= 0 ? src[i] : 0) * cin + c * TC_BK + sub * 8,
                        src[i] >= 0 ? 16 : 0);
        cp_async_commit();
        if (it >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(&full_bar[(it - LAG) % STAGES]);
        }
      }
    }
    // drain the last LAG stages
    cp_async_wait<0>();
    fence_proxy_async();
    for (int d = (total > LAG ? total - LAG : 0); d < total; ++d) mbar_arrive(&full_bar[d % STAGES]);

    // ---------------- epilogue ----------------
    const int row = tile * MT * TC_M + t;            // TMEM lane = row within its 128-row tile; accumulator = tile tm
    if (total > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    __nv_bfloat16* yrow = y + (long long)row * cout + n0;
#pragma unroll 1
    for (int c0 = 0; c0 < N_TILE; c0 += 32) {
      uint32_t v[32];
      if (total > 0) {
        tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tm * N_TILE + c0), v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (row < n_out) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16(v[8 * q + 0], v[8 * q + 1]);
          o.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
          o.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]);
          o.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
          *reinterpret_cast<uint4*>(yrow + c0 + 8 * q) = o;
        }
      }
    }
    tc_fence_before();
  } else {
    // ---------------- MMA issuer (last warp) ----------------
    const uint32_t idesc = make_idesc(TC_M, N_TILE, 0, B_MN ? 1 : 0);
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if ((threadIdx.x & 31) == 0) {
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < TC_BK / 16; ++kk) {
          const uint64_t bd = B_MN ? make_desc(b_addr + kk * 2048, 8192, 1024) : make_desc(b_addr + kk * 32, 16, 1024);
#pragma unroll
          for (int m = 0; m < MT; ++m)
            umma_bf16(tmem_base + m * N_TILE, make_desc(a_addr + m * A_STAGE_BYTES + kk * 32, 16, 1024), bd, idesc,
                      (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (it == total - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == NPROD / 32) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad: dW[k] (Cin, Cout) += sum over pairs p of offset k of x[pin[p], :]^T dy[pout[p], :]
// CTA = (offset k, pair split) x (128-channel slice of Cin) x (N_TILE slice of Cout); reduction dim = pairs (64 / stage)
// ------------------------------------------------------------------------------------------------------------
template <int N_TILE, int STAGES>
__global__ void __launch_bounds__(160)
spconv_tc_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                       const int* __restrict__ pair_in, const int* __restrict__ pair_out,
                       const int* __restrict__ k_offsets, float* __restrict__ dw, int cin, int cout, int K,
                       int chunk_pairs) {
  constexpr int A_BYTES = 64 * 256;              // 64 pairs x 128 channels (2 M-atoms of 64 ch)
  constexpr int B_BYTES = 64 * N_TILE * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int LAG = STAGES - 2;
  constexpr int NB_ATOMS = N_TILE / 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  // Work item = one chunk of `chunk_pairs` consecutive pairs of ONE offset (uniform work per CTA: no heavy-offset tail).
  // blockIdx.x enumerates chunks offset by offset; CTAs past the last chunk exit.
  int k = 0, chunk = blockIdx.x, p_beg = 0, p_end = 0;
  for (; k < K; ++k) {
    p_beg = k_offsets[k];
    p_end = k_offsets[k + 1];
    const int nch = (p_end - p_beg + chunk_pairs - 1) / chunk_pairs;
    if (chunk < nch) break;
    chunk -= nch;
  }
  if (k >= K) return;                              // uniform for the whole CTA
  const int ci0 = blockIdx.y * 128, co0 = blockIdx.z * N_TILE;
  const int s_beg = p_beg + chunk * chunk_pairs;
  const int s_end = min(p_end, s_beg + chunk_pairs);
  const int total = (s_end - s_beg + 63) / 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(tmem_slot, N_TILE < 32 ? 32 : N_TILE);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int a_ch = min(128, cin - ci0);            // valid channels of the A slice (64 or 128)

  if (warp < 4) {
    const int t = threadIdx.x;
    // thread t moves chunk (t & 15) of pair rows kk = (t >> 4) + 8 q, q = 0..7 (16 consecutive lanes = one 256 B row);
    // the pair indices of stage it+1 are fetched while the copies of stage it are in flight (index-load latency was
    // 55% of the stall samples in the round-1a profile).
    const int mc = t & 15, kk0 = t >> 4;
    const bool a_ok_ch = mc * 8 < a_ch;
    int ri[8], ro[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int p = s_beg + kk0 + 8 * q;
      ri[q] = p < s_end ? pair_in[p] : -1;
      ro[q] = p < s_end ? pair_out[p] : -1;
    }
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&empty_bar[s], ((it / STAGES) - 1) & 1);
      const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t b_base = a_base + A_BYTES;
      // canonical MN-major SW128: atom(mi, kj) at mi*8192 + kj*1024, row kk%8, 16 B chunk index ^ row
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int kk = kk0 + 8 * q;
        const bool ok = ri[q] >= 0 && a_ok_ch;
        const uint32_t dst = a_base + (mc >> 3) * 8192 + (kk >> 3) * 1024 + (kk & 7) * 128 + (((mc & 7) ^ (kk & 7)) << 4);
        cp_async16_cg(dst, x + (long long)(ok ? ri[q] : 0) * cin + ci0 + mc * 8, ok ? 16 : 0);
      }
#pragma unroll
      for (int q = 0; q < N_TILE / 16; ++q) {
        const int idx = q * 128 + t;
        const int kk = idx / (N_TILE / 8), nc = idx % (N_TILE / 8);
        // N_TILE = 128: kk = kk0 + 8 q (register ro[q]); N_TILE = 64: kk = (t >> 3) + 16 q -> fetch through ro when aligned
        int rr;
        if (N_TILE == 128) {
          rr = ro[q];
        } else {
          const int p = s_beg + it * 64 + kk;
          rr = p < s_end ? pair_out[p] : -1;
        }
        const uint32_t dst = b_base + (nc >> 3) * 8192 + (kk >> 3) * 1024 + (kk & 7) * 128 + (((nc & 7) ^ (kk & 7)) << 4);
        cp_async16_cg(dst, dy + (long long)(rr >= 0 ? rr : 0) * cout + co0 + nc * 8, rr >= 0 ? 16 : 0);
      }
      cp_async_commit();
      if (it + 1 < total) {   // prefetch the next stage's pair indices
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int p = s_beg + (it + 1) * 64 + kk0 + 8 * q;
          ri[q] = p < s_end ? pair_in[p] : -1;
          ro[q] = p < s_end ? pair_out[p] : -1;
        }
      }
      if (it >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(&full_bar[(it - LAG) % STAGES]);
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int d = (total > LAG ? total - LAG : 0); d < total; ++d) mbar_arrive(&full_bar[d % STAGES]);

    // epilogue: TMEM lane = ci (within the 128 slice), column = co. The tile is transposed through the (now idle) stage
    // buffers so that a warp adds whole 512-byte runs of one dW row with 16-byte vector atomics (coalesced RED traffic);
    // the round-1 epilogue issued 32 scalar atomics per thread with a lane stride of one row.
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    constexpr int PITCH = N_TILE + 4;                 // floats; +4 keeps the per-lane float4 stores conflict-free
    float* stg = reinterpret_cast<float*>(smem);
#pragma unroll 1
    for (int c0 = 0; c0 < N_TILE; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + threadIdx.x * PITCH + c0 + 4 * q) =
            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                        __uint_as_float(v[4 * q + 3]));
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int lane = threadIdx.x & 31;
    for (int rr = 0; rr < 32; ++rr) {
      const int r = warp * 32 + rr;
      const int ci = ci0 + r;
      if (ci >= cin) break;
      float* dwrow = dw + ((long long)k * cin + ci) * cout + co0;
#pragma unroll
      for (int c = lane * 4; c < N_TILE; c += 128)
        atomicAdd(reinterpret_cast<float4*>(dwrow + c), *reinterpret_cast<const float4*>(stg + r * PITCH + c));
    }
  } else {
    const uint32_t idesc = make_idesc(128, N_TILE, 1, 1);
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if ((threadIdx.x & 31) == 0) {
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {           // 16 pairs per MMA = two 8-row K groups = 2048 B
          uint64_t ad = make_desc(a_addr + kk * 2048, 8192, 1024);
          uint64_t bd = make_desc(b_addr + kk * 2048, 8192, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        if (it == total - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, N_TILE < 32 ? 32 : N_TILE);
  }
  (void)NB_ATOMS;
}

template <int N_TILE, int STAGES, bool B_MN, int MT>
int launch_fwd(const void* x, const CUtensorMap& tmw, const int* nbr, const unsigned* masks, void* y, long long n_out, int cin,
               int cout, int K, cudaStream_t stream) {
  size_t smem = (size_t)STAGES * (MT * A_STAGE_BYTES + N_TILE * 128) + 1024 + 256;
  auto kern = spconv_tc_fwd_kernel<N_TILE, STAGES, B_MN, MT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { esb_set_error("spconv_tc_fwd: smem attr: %s", cudaGetErrorString(e)); return ESB_ECUDA; }
  dim3 grid(esb_div_up(n_out, MT * TC_M), cout / N_TILE);
  kern<<<grid, MT * 128 + 32, smem, stream>>>((const __nv_bfloat16*)x, tmw, nbr, masks, (__nv_bfloat16*)y, (int)n_out, cin,
                                              cout, K);
  return ESB_OK;
}

template <int N_TILE, int STAGES>
int launch_wgrad(const void* x, const void* dy, const int* pin, const int* pout, const int* koff, float* dw, int cin,
                 int cout, int K, int n_chunks, int chunk_pairs, cudaStream_t stream) {
  size_t smem = (size_t)STAGES * (64 * 256 + 64 * N_TILE * 2) + 1024 + 256;
  auto kern = spconv_tc_wgrad_kernel<N_TILE, STAGES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { esb_set_error("spconv_tc_wgrad: smem attr: %s", cudaGetErrorString(e)); return ESB_ECUDA; }
  dim3 grid(n_chunks, esb_div_up(cin, 128), cout / N_TILE);
  kern<<<grid, 160, smem, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, pin, pout, koff, dw, cin, cout, K,
                                    chunk_pairs);
  return ESB_OK;
}

}  // namespace

// masks: ceil(n/128) uint32, bit k set when any row of the tile has a neighbour through offset k
extern "C" int esb_kmap_tile_masks(const int* nbr, int K, long long n, unsigned* masks, void* stream) {
  ESB_CHECK_ARG(K >= 1 && K <= 32, "esb_kmap_tile_masks: K must be in [1,32]");
  if (n == 0) return ESB_OK;
  tile_mask_kernel<<<esb_div_up(n, TC_M), 128, 0, (cudaStream_t)stream>>>(nbr, K, (int)n, masks);
  ESB_CUDA_LAUNCH_CHECK("tile_mask_kernel");
  return ESB_OK;
}

// bf16 tensor-core forward / dgrad. x (n_in,cin) ; nbr (K,n_out) ; masks from esb_kmap_tile_masks(nbr) ; y (n_out,cout).
// w_layout 0: w is (K,cout,cin) (reduction dim contiguous) ; w_layout 1: w is (K,cin,cout) (output dim contiguous).
// Forward passes the stored kernel with w_layout 1; dgrad passes the same tensor with roles swapped and w_layout 0.
// Requires cin % 64 == 0 and cout % 64 == 0.
extern "C" int esb_spconv_tc_fwd(const void* x, const void* wt, const int* nbr, const unsigned* masks, void* y,
                                 long long n_out, int cin, int cout, int K, int w_layout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0 && cin > 0 && cout > 0, "esb_spconv_tc_fwd: channels must be multiples of 64");
  ESB_CHECK_ARG(K >= 1 && K <= 27, "esb_spconv_tc_fwd: K must be in [1,27]");
  if (n_out == 0) return ESB_OK;
  int rc;
  // wide tiles amortise the gather; narrow tiles when there are too few row tiles to fill 148 SMs
  long long row_tiles = (n_out + TC_M - 1) / TC_M;
  const int n_tile = (cout % 256 == 0 && row_tiles * (cout / 256) >= 148) ? 256
                     : (cout % 128 == 0 && row_tiles * (cout / 128) >= 148) ? 128 : 64;
  // Two row tiles per CTA (one filter stage, two accumulators, 256 producer threads) halve the filter traffic, but one
  // 288-thread CTA per SM measured slower than two 160-thread CTAs (C2 step: 0.685 vs 0.700 of HBM peak, 46.8 vs 40.9 ms per
  // step, profiles/r2_bench_ab.txt): opt-in through ESB200_SPCONV_MT2=1.
  const bool two = n_tile <= 128 && (row_tiles / 2) * (cout / n_tile) >= 2 * 148 && getenv("ESB200_SPCONV_MT2") != nullptr;
  CUtensorMap tmw;
  {
    unsigned long long dims[2], str[1];
    unsigned box[2];
    if (w_layout) {   // (K*cin rows, cout columns): 64 x 64 boxes
      dims[0] = (unsigned long long)cout; dims[1] = (unsigned long long)K * cin; str[0] = (unsigned long long)cout * 2;
      box[0] = 64; box[1] = 64;
    } else {          // (K*cout rows, cin columns): N_TILE rows x 64 columns
      dims[0] = (unsigned long long)cin; dims[1] = (unsigned long long)K * cout; str[0] = (unsigned long long)cin * 2;
      box[0] = 64; box[1] = (unsigned)n_tile;
    }
    rc = esb_tma_encode(&tmw, wt, 2, dims, str, box, nullptr, 128);
    if (rc != ESB_OK) return rc;
  }
#define ESB_TC_LAUNCH(NT, ST, MTV)                                                                       \
  (w_layout ? launch_fwd<NT, ST, true, MTV>(x, tmw, nbr, masks, y, n_out, cin, cout, K, stream)         \
            : launch_fwd<NT, ST, false, MTV>(x, tmw, nbr, masks, y, n_out, cin, cout, K, stream))
  if (n_tile == 256)
    rc = ESB_TC_LAUNCH(256, 4, 1);
  else if (n_tile == 128)
    rc = two ? ESB_TC_LAUNCH(128, 4, 2) : ESB_TC_LAUNCH(128, 3, 1);
  else
    rc = two ? ESB_TC_LAUNCH(64, 4, 2) : ESB_TC_LAUNCH(64, 4, 1);
#undef ESB_TC_LAUNCH
  if (rc != ESB_OK) return rc;
  ESB_CUDA_LAUNCH_CHECK("spconv_tc_fwd_kernel");
  return ESB_OK;
}

// bf16 tensor-core wgrad over pair lists; dw (K,cin,cout) fp32, zeroed by the caller (atomic accumulation over splits).
extern "C" int esb_spconv_tc_wgrad(const void* x, const void* dy, const int* pair_in, const int* pair_out,
                                   const int* k_offsets, float* dw, long long n_pairs_hint, int cin, int cout, int K,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0 && cin > 0 && cout > 0, "esb_spconv_tc_wgrad: channels must be multiples of 64");
  int n_tile = (cout % 128 == 0) ? 128 : 64;
  // n_pairs_hint is an upper bound of the pair count (K * n_out); aim at ~4 waves of 2 CTAs/SM over the channel tiles
  long long tiles = (long long)esb_div_up(cin, 128) * (cout / n_tile);
  long long target_chunks = (4LL * 296 + tiles - 1) / tiles;
  long long cp = (n_pairs_hint / 2 + target_chunks - 1) / target_chunks;   // maps are ~40% dense: hint/2 ~ real pairs
  cp = (cp + 63) / 64 * 64;
  if (cp < 512) cp = 512;
  if (cp > 16384) cp = 16384;
  int chunk_pairs = (int)cp;
  int n_chunks = (int)(n_pairs_hint / chunk_pairs) + K + 1;               // upper bound; surplus CTAs exit immediately
  int rc = n_tile == 128
               ? launch_wgrad<128, 3>(x, dy, pair_in, pair_out, k_offsets, dw, cin, cout, K, n_chunks, chunk_pairs, stream)
               : launch_wgrad<64, 4>(x, dy, pair_in, pair_out, k_offsets, dw, cin, cout, K, n_chunks, chunk_pairs, stream);
  if (rc != ESB_OK) return rc;
  ESB_CUDA_LAUNCH_CHECK("spconv_tc_wgrad_kernel");
  return ESB_OK;
}
