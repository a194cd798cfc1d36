// esb200 — sparse 3D convolution, TMA-fed: the forward / dgrad / wgrad kernels of spconv_tc.cu with every operand moved
// by the tensor-memory accelerator instead of per-thread cp.async address generation
// (ME.MinkowskiConvolution, †upstream MinkowskiEngine; call sites embodiedscan/models/backbones/mink_resnet.py:58-62,104-108,
// embodiedscan/models/dense_heads/fcaf3d_head.py:919-946).
//
//   * gathered rows (A of forward/dgrad, both operands of wgrad): cp.async.bulk.tensor ... tile::gather4 — one instruction
//     fetches the 64-channel slice of FOUR rows named by the kernel map into four consecutive 128-byte swizzled smem rows;
//     a missing neighbour (-1) is out of bounds and arrives as zeros. One producer warp keeps every pipeline stage in flight
//     (round 1: 128 threads x 8 cp.async each per stage, at most two stages in flight per thread).
//   * filter tiles (B of forward/dgrad): plain 2-D tiled TMA boxes of the stored (K,Cin,Cout) kernel, MN-major for the forward
//     pass, K-major for dgrad — the same tensor, never transposed.
//   * forward/dgrad CTAs own MT x 128 output rows (MT = 2 when there are enough row tiles): one filter stage feeds two
//     accumulators in TMEM, halving the filter traffic per output row (it was as large as the gather itself at C = 64).
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (TMEM lane quadrant = warp & 3).
// Roofline: pair model bytes = P*(Cin+Cout)*2 + 8P + K*Cin*Cout*2.
#include "tc_common.cuh"

using namespace esb_tc;

namespace {

constexpr int THREADS = 192;

// ------------------------------------------------------------------------------------------------------------
// forward / dgrad
// ------------------------------------------------------------------------------------------------------------
template <int N_TILE, int MT, bool B_MN>
__global__ void __launch_bounds__(THREADS)
spconv_tma_fwd_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmw,
                      const int* __restrict__ nbr, const uint32_t* __restrict__ masks, __nv_bfloat16* __restrict__ y,
                      int n_out, int cin, int cout, int K, int stages) {
  constexpr int A_BYTES = MT * A_STAGE_BYTES;
  constexpr int B_BYTES = N_TILE * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = MT * N_TILE < 32 ? 32 : MT * N_TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int STAGES = stages;
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;                       // MT*128 output rows
  const int n0 = blockIdx.y * N_TILE;
  uint32_t mask = masks[tile * MT];
  if (MT == 2 && (tile * 2 + 1) * TC_M < n_out) mask |= masks[tile * 2 + 1];
  const int nchunk = cin / TC_BK;
  const int total = __popc(mask) * nchunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmw);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer: lane l owns rows 4l..4l+3 of each 128-row A tile ----------------
    const int row0 = tile * MT * TC_M + 4 * lane;
    int cur[MT][4];
    auto load_idx = [&](int k, int (&dst)[MT][4]) {
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = row0 + m * TC_M + j;
          dst[m][j] = r < n_out ? nbr[(long long)k * n_out + r] : -1;
        }
    };
    if (mask) load_idx(__ffs(mask) - 1, cur);
    int it = 0;
    for (uint32_t mk = mask; mk; mk &= mk - 1) {
      const int k = __ffs(mk) - 1;
      int nxt[MT][4];
      const uint32_t rest = mk & (mk - 1);
      if (rest) load_idx(__ffs(rest) - 1, nxt);      // the next offset's indices travel while this one's stages are issued
      for (int c = 0; c < nchunk; ++c, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        if (lane == 0) mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
        __syncwarp();
        const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int m = 0; m < MT; ++m)
          tma_gather4(&tmx, &full_bar[s], a_base + m * A_STAGE_BYTES + lane * 512, c * TC_BK, cur[m][0], cur[m][1], cur[m][2],
                      cur[m][3]);
        if (lane == 0) {
          const uint32_t b_base = a_base + A_BYTES;
          if (B_MN) {        // stored kernel (K*cin rows, cout columns): 64 reduction rows x 64 columns per box
#pragma unroll
            for (int a = 0; a < N_TILE / 64; ++a)
              tma_load_2d(&tmw, &full_bar[s], b_base + a * 8192, n0 + a * 64, k * cin + c * TC_BK);
          } else {           // (K*cout rows, cin columns): N_TILE output rows x 64 reduction columns
            tma_load_2d(&tmw, &full_bar[s], b_base, c * TC_BK, k * cout + n0);
          }
        }
      }
      if (rest) {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int j = 0; j < 4; ++j) cur[m][j] = nxt[m][j];
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = make_idesc(TC_M, N_TILE, 0, B_MN ? 1 : 0);
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
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
  } else {
    // ---------------- epilogue: each output row written once ----------------
    const int quad = warp & 3;
    if (total > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int m = 0; m < MT; ++m) {
      const int row = tile * MT * TC_M + m * TC_M + quad * 32 + lane;
      __nv_bfloat16* yrow = y + (long long)row * cout + n0;
#pragma unroll 1
      for (int c0 = 0; c0 < N_TILE; c0 += 32) {
        uint32_t v[32];
        if (total > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(m * N_TILE + c0), v);
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
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad: dW[k] (Cin, Cout) += sum over pairs p of offset k of x[pin[p], :]^T dy[pout[p], :]
// CTA = (offset k, chunk of pairs) x (128-channel slice of Cin) x (N_TILE slice of Cout); 64 pairs per stage, both operands
// gathered rows (MN-major), fp32 partial sums added to dW with coalesced vector atomics.
// ------------------------------------------------------------------------------------------------------------
template <int N_TILE>
__global__ void __launch_bounds__(THREADS)
spconv_tma_wgrad_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmdy,
                        const int* __restrict__ pair_in, const int* __restrict__ pair_out, const int* __restrict__ k_offsets,
                        float* __restrict__ dw, int cin, int cout, int K, int chunk_pairs, int stages) {
  constexpr int A_BYTES = 64 * 256;              // 64 pairs x 128 channels (2 M-atoms of 64 channels)
  constexpr int B_BYTES = 64 * N_TILE * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = N_TILE < 32 ? 32 : N_TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int STAGES = stages;
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  const int a_atoms = (cin - ci0) >= 128 ? 2 : 1;  // valid 64-channel atoms of the A slice

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmdy);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // lanes 0..15: pair group g = lane (pairs 4g..4g+3 of the 64-pair stage); lanes 16..31 idle (kept for __syncwarp)
    int ri[4], ro[4];
    auto load_idx = [&](int it_) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = s_beg + it_ * 64 + 4 * lane + j;
        const bool ok = lane < 16 && p < s_end;
        ri[j] = ok ? pair_in[p] : -1;
        ro[j] = ok ? pair_out[p] : -1;
      }
    };
    load_idx(0);
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
      if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(a_atoms * 64 * 128 + B_BYTES));
      __syncwarp();
      const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES);
      const uint32_t b_base = a_base + A_BYTES;
      if (lane < 16) {
        // canonical MN-major SW128: atom(mi) at mi*8192, pair row kk at kk*128 inside it; 4 rows = 512 contiguous bytes
        for (int mi = 0; mi < a_atoms; ++mi)
          tma_gather4(&tmx, &full_bar[s], a_base + mi * 8192 + lane * 512, ci0 + mi * 64, ri[0], ri[1], ri[2], ri[3]);
#pragma unroll
        for (int ni = 0; ni < N_TILE / 64; ++ni)
          tma_gather4(&tmdy, &full_bar[s], b_base + ni * 8192 + lane * 512, co0 + ni * 64, ro[0], ro[1], ro[2], ro[3]);
      }
      if (it + 1 < total) load_idx(it + 1);
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, N_TILE, 1, 1);
    for (int it = 0; it < total; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)             // 16 pairs per MMA = two 8-row K groups = 2048 B
          umma_bf16(tmem_base, make_desc(a_addr + kk * 2048, 8192, 1024), make_desc(b_addr + kk * 2048, 8192, 1024), idesc,
                    (it > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (it == total - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
    tc_fence_before();
  } else {
    // epilogue: TMEM lane = ci (within the 128 slice), column = co; transposed through the idle stage buffers so a warp adds
    // contiguous runs of one dW row with 16-byte vector atomics
    const int quad = warp & 3;
    const int et = quad * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    constexpr int PITCH = N_TILE + 4;
    float* stg = reinterpret_cast<float*>(smem);
#pragma unroll 1
    for (int c0 = 0; c0 < N_TILE; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + et * PITCH + c0 + 4 * q) =
            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                        __uint_as_float(v[4 * q + 3]));
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int rr = 0; rr < 32; ++rr) {
      const int r = quad * 32 + rr;
      const int ci = ci0 + r;
      if (ci >= cin || r >= a_atoms * 64) break;     // rows of an absent atom hold stale accumulator garbage
      float* dwrow = dw + ((long long)k * cin + ci) * cout + co0;
#pragma unroll
      for (int c = lane * 4; c < N_TILE; c += 128)
        atomicAdd(reinterpret_cast<float4*>(dwrow + c), *reinterpret_cast<const float4*>(stg + r * PITCH + c));
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int rows_map(CUtensorMap* tm, const void* base, long long rows, int cols) {   // (rows, cols) bf16 row-major, gather4 boxes
  unsigned long long dims[2] = {(unsigned long long)cols, (unsigned long long)(rows > 0 ? rows : 1)};
  unsigned long long str[1] = {(unsigned long long)cols * 2};
  unsigned box[2] = {64, 1};
  return esb_tma_encode(tm, base, 2, dims, str, box, nullptr, 128);
}

template <int N_TILE, int MT, bool B_MN>
int launch_fwd(const CUtensorMap& tmx, const CUtensorMap& tmw, const int* nbr, const unsigned* masks, void* y, long long n_out,
               int cin, int cout, int K, cudaStream_t stream) {
  constexpr int STAGE_BYTES = MT * A_STAGE_BYTES + N_TILE * 128;
  // MT = 2 / N_TILE = 256: one CTA per SM with ~200 KB of stages in flight; small tiles: two CTAs per SM
  int stages = ((MT == 1 && N_TILE <= 128 ? 100 : 200) * 1024) / STAGE_BYTES;
  stages = stages > 6 ? 6 : stages;
  const size_t smem = (size_t)stages * STAGE_BYTES + (2 * stages + 1) * 8 + 16 + 1024;
  auto kern = spconv_tma_fwd_kernel<N_TILE, MT, B_MN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { esb_set_error("spconv_tma_fwd: smem attr: %s", cudaGetErrorString(e)); return ESB_ECUDA; }
  dim3 grid(esb_div_up(n_out, MT * TC_M), cout / N_TILE);
  kern<<<grid, THREADS, smem, stream>>>(tmx, tmw, nbr, masks, (__nv_bfloat16*)y, (int)n_out, cin, cout, K, stages);
  return ESB_OK;
}

template <int N_TILE>
int launch_wgrad(const CUtensorMap& tmx, const CUtensorMap& tmdy, const int* pin, const int* pout, const int* koff, float* dw,
                 int cin, int cout, int K, int n_chunks, int chunk_pairs, cudaStream_t stream) {
  constexpr int STAGE_BYTES = 64 * 256 + 64 * N_TILE * 2;
  int stages = (100 * 1024) / STAGE_BYTES;           // two CTAs per SM
  const int need = (128 * (N_TILE + 4) * 4 + STAGE_BYTES - 1) / STAGE_BYTES;   // the epilogue reuses the stage buffers
  stages = stages < need ? need : stages;
  const size_t smem = (size_t)stages * STAGE_BYTES + (2 * stages + 1) * 8 + 16 + 1024;
  auto kern = spconv_tma_wgrad_kernel<N_TILE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { esb_set_error("spconv_tma_wgrad: smem attr: %s", cudaGetErrorString(e)); return ESB_ECUDA; }
  dim3 grid(n_chunks, esb_div_up(cin, 128), cout / N_TILE);
  kern<<<grid, THREADS, smem, stream>>>(tmx, tmdy, pin, pout, koff, dw, cin, cout, K, chunk_pairs, stages);
  return ESB_OK;
}

}  // namespace

// Same contract as esb_spconv_tc_fwd (x (n_in,cin), nbr (K,n_out), masks per 128-row tile, y (n_out,cout); w_layout 0: w is
// (K,cout,cin), 1: (K,cin,cout)) plus n_in, the row count of x (the extent of the gather's tensor map).
extern "C" int esb_spconv_tma_fwd(const void* x, const void* wt, const int* nbr, const unsigned* masks, void* y, long long n_in,
                                  long long n_out, int cin, int cout, int K, int w_layout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0 && cin > 0 && cout > 0, "esb_spconv_tma_fwd: channels must be multiples of 64");
  ESB_CHECK_ARG(K >= 1 && K <= 27, "esb_spconv_tma_fwd: K must be in [1,27]");
  if (n_out == 0) return ESB_OK;
  CUtensorMap tmx, tmw;
  int rc = rows_map(&tmx, x, n_in, cin);
  if (rc != ESB_OK) return rc;
  const long long row_tiles = (n_out + TC_M - 1) / TC_M;
  int n_tile = (cout % 256 == 0 && row_tiles * (cout / 256) >= 148) ? 256 : (cout % 128 == 0 && row_tiles * (cout / 128) >= 148) ? 128 : 64;
  const bool two = n_tile <= 128 && (row_tiles / 2) * (cout / n_tile) >= 148;      // M = 256 only while the grid still fills the SMs
  {
    unsigned long long dims[2], str[1];
    unsigned box[2];
    if (w_layout) {   // (K*cin rows, cout columns)
      dims[0] = (unsigned long long)cout; dims[1] = (unsigned long long)K * cin; str[0] = (unsigned long long)cout * 2;
      box[0] = 64; box[1] = 64;
    } else {          // (K*cout rows, cin columns)
      dims[0] = (unsigned long long)cin; dims[1] = (unsigned long long)K * cout; str[0] = (unsigned long long)cin * 2;
      box[0] = 64; box[1] = (unsigned)n_tile;
    }
    rc = esb_tma_encode(&tmw, wt, 2, dims, str, box, nullptr, 128);
    if (rc != ESB_OK) return rc;
  }
#define ESB_TM(NT, MTV) (w_layout ? launch_fwd<NT, MTV, true>(tmx, tmw, nbr, masks, y, n_out, cin, cout, K, stream) \
                                  : launch_fwd<NT, MTV, false>(tmx, tmw, nbr, masks, y, n_out, cin, cout, K, stream))
  if (n_tile == 256) rc = ESB_TM(256, 1);
  else if (n_tile == 128) rc = two ? ESB_TM(128, 2) : ESB_TM(128, 1);
  else rc = two ? ESB_TM(64, 2) : ESB_TM(64, 1);
#undef ESB_TM
  if (rc != ESB_OK) return rc;
  ESB_CUDA_LAUNCH_CHECK("spconv_tma_fwd_kernel");
  return ESB_OK;
}

// Same contract as esb_spconv_tc_wgrad plus the row counts of x and dy.
extern "C" int esb_spconv_tma_wgrad(const void* x, const void* dy, const int* pair_in, const int* pair_out,
                                    const int* k_offsets, float* dw, long long n_in, long long n_out, long long n_pairs_hint,
                                    int cin, int cout, int K, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(cin % 64 == 0 && cout % 64 == 0 && cin > 0 && cout > 0, "esb_spconv_tma_wgrad: channels must be multiples of 64");
  CUtensorMap tmx, tmdy;
  int rc = rows_map(&tmx, x, n_in, cin);
  if (rc != ESB_OK) return rc;
  rc = rows_map(&tmdy, dy, n_out, cout);
  if (rc != ESB_OK) return rc;
  const int n_tile = (cout % 128 == 0) ? 128 : 64;
  const long long tiles = (long long)esb_div_up(cin, 128) * (cout / n_tile);
  const long long target_chunks = (4LL * 296 + tiles - 1) / tiles;
  long long cp = (n_pairs_hint / 2 + target_chunks - 1) / target_chunks;   // maps are ~40% dense: hint/2 ~ real pairs
  cp = (cp + 63) / 64 * 64;
  if (cp < 512) cp = 512;
  if (cp > 16384) cp = 16384;
  const int chunk_pairs = (int)cp;
  const int n_chunks = (int)(n_pairs_hint / chunk_pairs) + K + 1;          // upper bound; surplus CTAs exit immediately
  rc = n_tile == 128 ? launch_wgrad<128>(tmx, tmdy, pair_in, pair_out, k_offsets, dw, cin, cout, K, n_chunks, chunk_pairs, stream)
                     : launch_wgrad<64>(tmx, tmdy, pair_in, pair_out, k_offsets, dw, cin, cout, K, n_chunks, chunk_pairs, stream);
  if (rc != ESB_OK) return rc;
  ESB_CUDA_LAUNCH_CHECK("spconv_tma_wgrad_kernel");
  return ESB_OK;
}
