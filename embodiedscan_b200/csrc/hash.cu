// esb200 — voxelisation hashing, coordinate maps and kernel maps (replaces the
// MinkowskiEngine CoordinateManager surface used at
// embodiedscan/models/detectors/sparse_featfusion_single_stage.py:109-118 and inside every
// ME.MinkowskiConvolution / MaxPooling / GenerativeConvolutionTranspose of mink_resnet.py and fcaf3d_head.py).
//
// Integer work, HBM/L2-latency bound: open-addressing hash table with 64-bit packed keys, linear probing.
// Determinism rule (ME CPU semantics): among duplicate coordinates the FIRST row wins and output
// rows keep first-occurrence order. Implemented as atomicMin(row) per slot + flag + exclusive scan.
#include "common.cuh"
#include <cub/cub.cuh>
#include <limits.h>

static thread_local char g_err[512] = "";
extern "C" const char* esb_last_error() { return g_err; }
void esb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// voxelize: coords[i] = (batch, floor(p * inv_voxel)); torch's CUDA `tensor / python_scalar` multiplies
// by the fp32 reciprocal, which is what the reference's GPU path evaluates for `p[:, :3] / self.voxel_size`.
// ------------------------------------------------------------------------------------------------
__global__ void voxelize_kernel(const float* __restrict__ pts, long long n, int pstride, int batch,
                                float inv_voxel, int* __restrict__ coords) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pts + i * pstride;
  int4 c;
  c.x = batch;
  c.y = (int)floorf(__fmul_rn(p[0], inv_voxel));
  c.z = (int)floorf(__fmul_rn(p[1], inv_voxel));
  c.w = (int)floorf(__fmul_rn(p[2], inv_voxel));
  reinterpret_cast<int4*>(coords)[i] = c;
}

extern "C" int esb_voxelize_points(const float* pts, long long n, int pstride, int batch, float inv_voxel,
                                   int* coords, void* stream) {
  ESB_CHECK_ARG(n >= 0 && pstride >= 3, "esb_voxelize_points: bad n/pstride");
  if (n == 0) return ESB_OK;
  voxelize_kernel<<<esb_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, n, pstride, batch, inv_voxel, coords);
  ESB_CUDA_LAUNCH_CHECK("voxelize_kernel");
  return ESB_OK;
}

// ------------------------------------------------------------------------------------------------
// hash table
// ------------------------------------------------------------------------------------------------
extern "C" long long esb_hash_capacity(long long n) {
  long long cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

__global__ void hash_clear_kernel(unsigned long long* keys, int* vals, long long cap) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < cap) {
    keys[i] = ESB_EMPTY_KEY;
    vals[i] = INT_MAX;
  }
}

// Insert row i under key(coord / div * div); vals[slot] = min row index. slot_of[i] remembers the slot.
__global__ void hash_insert_min_kernel(const int* __restrict__ coords, long long n, int div,
                                       unsigned long long* keys, int* vals, uint32_t mask,
                                       int* __restrict__ slot_of, int* __restrict__ err) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = reinterpret_cast<const int4*>(coords)[i];
  if (div > 1) {
    c.y = esb_floor_div(c.y, div) * div;
    c.z = esb_floor_div(c.z, div) * div;
    c.w = esb_floor_div(c.w, div) * div;
  }
  if (!esb_coord_in_range(c.x, c.y, c.z, c.w)) {
    atomicExch(err, 1);
    slot_of[i] = -1;
    return;
  }
  unsigned long long key = esb_pack_key(c.x, c.y, c.z, c.w);
  uint32_t slot = esb_hash64(key) & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long prev = atomicCAS(&keys[slot], ESB_EMPTY_KEY, key);
    if (prev == ESB_EMPTY_KEY || prev == key) {
      atomicMin(&vals[slot], (int)i);
      slot_of[i] = (int)slot;
      return;
    }
    slot = (slot + 1) & mask;
  }
  atomicExch(err, 2);  // table full (cannot happen with cap >= 2n)
  slot_of[i] = -1;
}

__global__ void flag_winner_kernel(const int* __restrict__ slot_of, const int* __restrict__ vals, long long n,
                                   int* __restrict__ flag) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = slot_of[i];
  flag[i] = (s >= 0 && vals[s] == (int)i) ? 1 : 0;
}

// winners write their (quantised) coordinate to the compacted list and publish their output row in the table
__global__ void compact_winner_kernel(const int* __restrict__ coords, long long n, int div,
                                      const int* __restrict__ flag, const int* __restrict__ rank,
                                      const int* __restrict__ slot_of, int* vals, int* __restrict__ out_coords,
                                      int* __restrict__ count, const int* __restrict__ err) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) {
    int4 c = reinterpret_cast<const int4*>(coords)[i];
    if (div > 1) {
      c.y = esb_floor_div(c.y, div) * div;
      c.z = esb_floor_div(c.z, div) * div;
      c.w = esb_floor_div(c.w, div) * div;
    }
    int r = rank[i];
    reinterpret_cast<int4*>(out_coords)[r] = c;
    vals[slot_of[i]] = r;
  }
  // a negative count reports an out-of-range coordinate (-1) or a full table (-2) to the host
  if (i == n - 1) *count = (*err) ? -(*err) : rank[i] + flag[i];
}

__global__ void inverse_map_kernel(const int* __restrict__ slot_of, const int* __restrict__ vals, long long n,
                                   int* __restrict__ in2out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = slot_of[i];
  in2out[i] = s >= 0 ? vals[s] : -1;
}

static size_t scan_temp_bytes(long long n) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int*)nullptr, (int*)nullptr, (int)n);
  return tb;
}

extern "C" size_t esb_coord_unique_workspace_bytes(long long n) {
  // slot_of, flag, rank, err + cub temp
  return 3 * esb_align((size_t)n * 4) + 256 + esb_align(scan_temp_bytes(n > 0 ? n : 1));
}

// Deduplicate (optionally after quantising xyz to multiples of `div`) with first-occurrence order.
//   keys/vals : table of `cap` slots (cap from esb_hash_capacity(n)); on return maps coordinate -> output row.
//   out_coords: (n,4) capacity; first *count rows valid.  in2out: (n) input row -> output row.
//   count_dev : device int32. The caller reads it after synchronising the stream.
extern "C" int esb_coord_unique(const int* coords_in, long long n, int div, unsigned long long* keys, int* vals,
                                long long cap, int* out_coords, int* in2out, int* count_dev, void* ws,
                                size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(n >= 0 && div >= 1, "esb_coord_unique: bad n/div");
  ESB_CHECK_ARG(cap >= 2 * n && (cap & (cap - 1)) == 0, "esb_coord_unique: cap must be a power of two >= 2n");
  ESB_CHECK_ARG(n < INT_MAX, "esb_coord_unique: n too large");
  if (ws_bytes < esb_coord_unique_workspace_bytes(n)) {
    esb_set_error("esb_coord_unique: workspace too small");
    return ESB_ENOMEM;
  }
  hash_clear_kernel<<<esb_div_up(cap, 256), 256, 0, stream>>>(keys, vals, cap);
  if (n == 0) {
    ESB_CUDA_CALL(cudaMemsetAsync(count_dev, 0, 4, stream));
    return ESB_OK;
  }
  char* p = (char*)ws;
  int* slot_of = (int*)p; p += esb_align((size_t)n * 4);
  int* flag = (int*)p;    p += esb_align((size_t)n * 4);
  int* rank = (int*)p;    p += esb_align((size_t)n * 4);
  int* err = (int*)p;     p += 256;
  void* cub_tmp = p;
  size_t cub_bytes = scan_temp_bytes(n);
  ESB_CUDA_CALL(cudaMemsetAsync(err, 0, 4, stream));
  int grid = esb_div_up(n, 256);
  hash_insert_min_kernel<<<grid, 256, 0, stream>>>(coords_in, n, div, keys, vals, (uint32_t)(cap - 1), slot_of, err);
  flag_winner_kernel<<<grid, 256, 0, stream>>>(slot_of, vals, n, flag);
  cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, flag, rank, (int)n, stream);
  compact_winner_kernel<<<grid, 256, 0, stream>>>(coords_in, n, div, flag, rank, slot_of, vals, out_coords, count_dev, err);
  inverse_map_kernel<<<grid, 256, 0, stream>>>(slot_of, vals, n, in2out);
  ESB_CUDA_LAUNCH_CHECK("esb_coord_unique");
  return ESB_OK;
}

// Build a table for coordinates that are already unique: vals[slot(coord_i)] = i.
__global__ void hash_build_kernel(const int* __restrict__ coords, long long n, unsigned long long* keys, int* vals,
                                  uint32_t mask) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int4 c = reinterpret_cast<const int4*>(coords)[i];
  unsigned long long key = esb_pack_key(c.x, c.y, c.z, c.w);
  uint32_t slot = esb_hash64(key) & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    unsigned long long prev = atomicCAS(&keys[slot], ESB_EMPTY_KEY, key);
    if (prev == ESB_EMPTY_KEY || prev == key) {
      atomicMin(&vals[slot], (int)i);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

extern "C" int esb_hash_build(const int* coords, long long n, unsigned long long* keys, int* vals, long long cap,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESB_CHECK_ARG(cap >= 2 * n && (cap & (cap - 1)) == 0, "esb_hash_build: cap must be a power of two >= 2n");
  hash_clear_kernel<<<esb_div_up(cap, 256), 256, 0, stream>>>(keys, vals, cap);
  if (n > 0) hash_build_kernel<<<esb_div_up(n, 256), 256, 0, stream>>>(coords, n, keys, vals, (uint32_t)(cap - 1));
  ESB_CUDA_LAUNCH_CHECK("esb_hash_build");
  return ESB_OK;
}

__global__ void hash_lookup_kernel(const int* __restrict__ q, long long nq, const unsigned long long* __restrict__ keys,
                                   const int* __restrict__ vals, uint32_t mask, int* __restrict__ out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int4 c = reinterpret_cast<const int4*>(q)[i];
  out[i] = esb_coord_in_range(c.x, c.y, c.z, c.w)
               ? esb_hash_find(keys, vals, mask, esb_pack_key(c.x, c.y, c.z, c.w))
               : -1;
}

extern "C" int esb_hash_lookup(const int* query, long long nq, const unsigned long long* keys, const int* vals,
                               long long cap, int* out, void* stream) {
  if (nq == 0) return ESB_OK;
  hash_lookup_kernel<<<esb_div_up(nq, 256), 256, 0, (cudaStream_t)stream>>>(query, nq, keys, vals,
                                                                           (uint32_t)(cap - 1), out);
  ESB_CUDA_LAUNCH_CHECK("hash_lookup_kernel");
  return ESB_OK;
}

// ------------------------------------------------------------------------------------------------
// ME `features_at_coordinates` (†upstream; used by FCAF3D `_prune`, fcaf3d_head.py:1091-1114): multilinear interpolation of
// the rows of a stride-`ts` tensor at integer query coordinates [b,x,y,z]; absent lattice points contribute 0. Arithmetic is
// spelled out operation by operation (no FMA contraction) in the order of the torch expression the oracle uses:
//   frac = q/ts - floor(q/ts); w_k = ((wx*wy)*wz); out = sum_{k=0..7} F[idx_k] * w_k  (k = dx + 2dy + 4dz).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void interp_features_kernel(const int* __restrict__ q, long long nq, const unsigned long long* __restrict__ keys,
                                       const int* __restrict__ vals, uint32_t mask, const T* __restrict__ feats, int C, int ts,
                                       float inv_ts, float* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const int4 c = reinterpret_cast<const int4*>(q)[i];
  float frac[3];
  int base[3];
  const int xyz[3] = {c.y, c.z, c.w};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float v = __fmul_rn((float)xyz[a], inv_ts);      // ts is a power of two: exact, = q / ts
    const float fl = floorf(v);
    frac[a] = __fsub_rn(v, fl);
    base[a] = (int)fl * ts;
  }
  for (int ch = 0; ch < C; ++ch) out[i * C + ch] = 0.f;
  for (int k = 0; k < 8; ++k) {
    const int d[3] = {k & 1, (k >> 1) & 1, (k >> 2) & 1};
    float w = d[0] ? frac[0] : __fsub_rn(1.f, frac[0]);
    w = __fmul_rn(w, d[1] ? frac[1] : __fsub_rn(1.f, frac[1]));
    w = __fmul_rn(w, d[2] ? frac[2] : __fsub_rn(1.f, frac[2]));
    const int x = base[0] + d[0] * ts, y = base[1] + d[1] * ts, z = base[2] + d[2] * ts;
    const int idx = esb_coord_in_range(c.x, x, y, z) ? esb_hash_find(keys, vals, mask, esb_pack_key(c.x, x, y, z)) : -1;
    if (idx >= 0)
      for (int ch = 0; ch < C; ++ch)
        out[i * C + ch] = __fadd_rn(out[i * C + ch], __fmul_rn(esb_to_float(feats[(long long)idx * C + ch]), w));
  }
}

extern "C" int esb_interp_features(const int* query, long long nq, const unsigned long long* keys, const int* vals,
                                   long long cap, const void* feats, int C, int ts, int dtype, float* out, void* stream) {
  ESB_CHECK_ARG(ts >= 1 && (ts & (ts - 1)) == 0 && C >= 1, "esb_interp_features: tensor stride must be a power of two");
  if (nq == 0) return ESB_OK;
  if (dtype == ESB_F32)
    interp_features_kernel<float><<<esb_div_up(nq, 256), 256, 0, (cudaStream_t)stream>>>(
        query, nq, keys, vals, (uint32_t)(cap - 1), (const float*)feats, C, ts, 1.0f / (float)ts, out);
  else
    interp_features_kernel<__nv_bfloat16><<<esb_div_up(nq, 256), 256, 0, (cudaStream_t)stream>>>(
        query, nq, keys, vals, (uint32_t)(cap - 1), (const __nv_bfloat16*)feats, C, ts, 1.0f / (float)ts, out);
  ESB_CUDA_LAUNCH_CHECK("interp_features_kernel");
  return ESB_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel map: nbr[k*n_out + o] = row of (out_coord[o] + offset[k]) in the input table, or -1.
// Output-stationary layout: offset-major so a warp's stores along o coalesce.
// ------------------------------------------------------------------------------------------------
struct OffsetList {
  int n;
  int off[27 * 3];
};

__global__ void kernel_map_kernel(const int* __restrict__ out_coords, long long n_out, OffsetList offs,
                                  const unsigned long long* __restrict__ keys, const int* __restrict__ vals,
                                  uint32_t mask, int* __restrict__ nbr) {
  long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  int4 c = reinterpret_cast<const int4*>(out_coords)[o];
  for (int k = 0; k < offs.n; ++k) {
    int x = c.y + offs.off[3 * k], y = c.z + offs.off[3 * k + 1], z = c.w + offs.off[3 * k + 2];
    int r = -1;
    if (esb_coord_in_range(c.x, x, y, z)) r = esb_hash_find(keys, vals, mask, esb_pack_key(c.x, x, y, z));
    nbr[(long long)k * n_out + o] = r;
  }
}

extern "C" int esb_kernel_map(const int* out_coords, long long n_out, const int* offsets_host, int K,
                              const unsigned long long* keys, const int* vals, long long cap, int* nbr,
                              void* stream) {
  ESB_CHECK_ARG(K >= 1 && K <= 27, "esb_kernel_map: K must be in [1,27]");
  if (n_out == 0) return ESB_OK;
  OffsetList offs;
  offs.n = K;
  for (int i = 0; i < 3 * K; ++i) offs.off[i] = offsets_host[i];
  kernel_map_kernel<<<esb_div_up(n_out, 128), 128, 0, (cudaStream_t)stream>>>(out_coords, n_out, offs, keys, vals,
                                                                              (uint32_t)(cap - 1), nbr);
  ESB_CUDA_LAUNCH_CHECK("kernel_map_kernel");
  return ESB_OK;
}

// input-stationary map (used by dgrad): nbr_in[k*n_in + i] = o  <=>  nbr_out[k*n_out + o] = i.
__global__ void fill_kernel(int* p, long long n, int v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void kernel_map_transpose_kernel(const int* __restrict__ nbr_out, int K, long long n_out, long long n_in,
                                            int* __restrict__ nbr_in) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)K * n_out) return;
  int k = (int)(t / n_out);
  long long o = t - (long long)k * n_out;
  int i = nbr_out[t];
  if (i >= 0) nbr_in[(long long)k * n_in + i] = (int)o;
}

extern "C" int esb_kernel_map_transpose(const int* nbr_out, int K, long long n_out, long long n_in, int* nbr_in,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  long long tot_in = (long long)K * n_in, tot_out = (long long)K * n_out;
  if (tot_in > 0) fill_kernel<<<esb_div_up(tot_in, 256), 256, 0, stream>>>(nbr_in, tot_in, -1);
  if (tot_out > 0)
    kernel_map_transpose_kernel<<<esb_div_up(tot_out, 256), 256, 0, stream>>>(nbr_out, K, n_out, n_in, nbr_in);
  ESB_CUDA_LAUNCH_CHECK("kernel_map_transpose");
  return ESB_OK;
}

// pair lists (offset-major, then output row): compaction of the valid entries of nbr.
__global__ void pair_flag_kernel(const int* __restrict__ nbr, long long tot, int* __restrict__ flag) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < tot) flag[t] = nbr[t] >= 0 ? 1 : 0;
}
__global__ void pair_compact_kernel(const int* __restrict__ nbr, int K, long long n_out, const int* __restrict__ flag,
                                    const int* __restrict__ rank, int* __restrict__ pair_in,
                                    int* __restrict__ pair_out, int* __restrict__ k_offsets) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long tot = (long long)K * n_out;
  if (t >= tot) return;
  int k = (int)(t / n_out);
  long long o = t - (long long)k * n_out;
  if (o == 0) k_offsets[k] = rank[t];
  if (t == tot - 1) k_offsets[K] = rank[t] + flag[t];
  if (flag[t]) {
    int r = rank[t];
    pair_in[r] = nbr[t];
    pair_out[r] = (int)o;
  }
}

extern "C" size_t esb_kmap_pairs_workspace_bytes(int K, long long n_out) {
  long long tot = (long long)K * n_out;
  if (tot < 1) tot = 1;
  return 2 * esb_align((size_t)tot * 4) + esb_align(scan_temp_bytes(tot));
}

extern "C" int esb_kmap_pairs(const int* nbr, int K, long long n_out, int* pair_in, int* pair_out, int* k_offsets,
                              void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  long long tot = (long long)K * n_out;
  ESB_CHECK_ARG(tot < INT_MAX, "esb_kmap_pairs: K*n_out too large");
  if (ws_bytes < esb_kmap_pairs_workspace_bytes(K, n_out)) {
    esb_set_error("esb_kmap_pairs: workspace too small");
    return ESB_ENOMEM;
  }
  if (tot == 0) {
    ESB_CUDA_CALL(cudaMemsetAsync(k_offsets, 0, 4 * (K + 1), stream));
    return ESB_OK;
  }
  char* p = (char*)ws;
  int* flag = (int*)p; p += esb_align((size_t)tot * 4);
  int* rank = (int*)p; p += esb_align((size_t)tot * 4);
  size_t cub_bytes = scan_temp_bytes(tot);
  pair_flag_kernel<<<esb_div_up(tot, 256), 256, 0, stream>>>(nbr, tot, flag);
  cub::DeviceScan::ExclusiveSum((void*)p, cub_bytes, flag, rank, (int)tot, stream);
  pair_compact_kernel<<<esb_div_up(tot, 256), 256, 0, stream>>>(nbr, K, n_out, flag, rank, pair_in, pair_out, k_offsets);
  ESB_CUDA_LAUNCH_CHECK("esb_kmap_pairs");
  return ESB_OK;
}

// generative transpose (k2 s2): child row = parent*8 + k, k = dx + 2*dy + 4*dz, coord = parent + d*half_stride.
__global__ void generative_children_kernel(const int* __restrict__ coords_in, long long n_in, int half,
                                           int* __restrict__ out) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_in * 8) return;
  long long p = t >> 3;
  int k = (int)(t & 7);
  int4 c = reinterpret_cast<const int4*>(coords_in)[p];
  c.y += (k & 1) * half;
  c.z += ((k >> 1) & 1) * half;
  c.w += ((k >> 2) & 1) * half;
  reinterpret_cast<int4*>(out)[t] = c;
}

extern "C" int esb_generative_children(const int* coords_in, long long n_in, int half_stride, int* out_coords,
                                       void* stream) {
  if (n_in == 0) return ESB_OK;
  generative_children_kernel<<<esb_div_up(n_in * 8, 256), 256, 0, (cudaStream_t)stream>>>(coords_in, n_in,
                                                                                        half_stride, out_coords);
  ESB_CUDA_LAUNCH_CHECK("generative_children_kernel");
  return ESB_OK;
}
