// esb200 — batched one-to-one (Hungarian) assignment on the device.
//
// Replaces the reference's per-sample, per-decoder-layer `cost.cpu()` -> scipy.optimize.linear_sum_assignment ->
// `.to(device)` round trip (embodiedscan/models/task_modules/assigners/hungarian_assigner.py:110-126, called 7 x batch
// times per iteration from grounding_head.py:398) by ONE launch that solves every (layer, sample) problem: one CTA per
// problem, shortest-augmenting-path (Jonker-Volgenant style potentials) in fp64 with the column scan spread over the
// CTA's threads. Each problem is `n_pred x n_gt[p]` (n_gt[p] <= n_pred); the n_gt "targets" are the rows that get
// augmented, predictions are the columns, so the work is O(n_gt^2 * n_pred / threads) per problem.
//
// Exactness: integer output. The optimum of a generic float cost matrix is unique, so the result equals scipy's; exact
// ties (measure zero for real costs) resolve to the lowest prediction index.
#include "common.cuh"

namespace {

constexpr int HUNG_THREADS = 256;
constexpr double HUNG_INF = 1e300;

// nan -> 100, +inf -> 100, -inf -> -100: torch.nan_to_num(cost, nan=100.0, posinf=100.0, neginf=-100.0)
__device__ __forceinline__ double sanitize(float c) {
  if (isnan(c)) return 100.0;
  if (isinf(c)) return c > 0 ? 100.0 : -100.0;
  return (double)c;
}

__global__ void __launch_bounds__(HUNG_THREADS)
hungarian_kernel(const float* __restrict__ cost, const int* __restrict__ n_gt, int n_pred, int ld_gt,
                 long long problem_stride, int* __restrict__ pred_to_gt,
                 int* __restrict__ gt_to_pred) {
  extern __shared__ unsigned char smem_raw[];
  const int prob = blockIdx.x;
  const int n = n_gt[prob];          // rows to assign (targets)
  const int m = n_pred;              // columns (predictions)
  const float* a = cost + (long long)prob * problem_stride;   // a[pred * ld_gt + gt]
  int* out = pred_to_gt + (long long)prob * m;
  const int tid = threadIdx.x;

  double* u = reinterpret_cast<double*>(smem_raw);            // [ld_gt + 1] row potentials (1-based)
  double* v = u + (ld_gt + 1);                                // [m + 1] column potentials
  double* minv = v + (m + 1);                                 // [m + 1]
  int* p = reinterpret_cast<int*>(minv + (m + 1));            // [m + 1] row matched to column (0 = free)
  int* way = p + (m + 1);                                     // [m + 1]
  unsigned char* used = reinterpret_cast<unsigned char*>(way + (m + 1));   // [m + 1]
  __shared__ double red_val[HUNG_THREADS / 32];
  __shared__ int red_idx[HUNG_THREADS / 32];
  __shared__ double s_delta;
  __shared__ int s_j1;

  for (int j = tid; j <= m; j += HUNG_THREADS) { v[j] = 0.0; p[j] = 0; way[j] = 0; }
  for (int i = tid; i <= ld_gt; i += HUNG_THREADS) u[i] = 0.0;
  for (int j = tid; j < m; j += HUNG_THREADS) out[j] = -1;
  if (gt_to_pred != nullptr)
    for (int i = tid; i < ld_gt; i += HUNG_THREADS) gt_to_pred[(long long)prob * ld_gt + i] = -1;
  __syncthreads();
  if (n <= 0) return;

  for (int i = 1; i <= n; ++i) {
    for (int j = tid; j <= m; j += HUNG_THREADS) { minv[j] = HUNG_INF; used[j] = 0; }
    if (tid == 0) p[0] = i;
    __syncthreads();
    int j0 = 0;
    while (true) {
      if (tid == 0) used[j0] = 1;
      __syncthreads();
      const int i0 = p[j0];
      const double ui0 = u[i0];
      double best = HUNG_INF;
      int best_j = 0x7fffffff;
      for (int j = tid + 1; j <= m; j += HUNG_THREADS) {
        if (!used[j]) {
          const double cur = sanitize(a[(long long)(j - 1) * ld_gt + (i0 - 1)]) - ui0 - v[j];
          if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
          const double mv = minv[j];
          if (mv < best) { best = mv; best_j = j; }       // ascending j per thread: first minimum kept
        }
      }
      // block argmin, ties -> lowest column
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int oj = __shfl_down_sync(0xffffffffu, best_j, off);
        if (ov < best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
      }
      if ((tid & 31) == 0) { red_val[tid >> 5] = best; red_idx[tid >> 5] = best_j; }
      __syncthreads();
      if (tid == 0) {
        double b = red_val[0];
        int bj = red_idx[0];
        for (int w = 1; w < HUNG_THREADS / 32; ++w)
          if (red_val[w] < b || (red_val[w] == b && red_idx[w] < bj)) { b = red_val[w]; bj = red_idx[w]; }
        s_delta = b;
        s_j1 = bj;
      }
      __syncthreads();
      const double delta = s_delta;
      const int j1 = s_j1;
      for (int j = tid; j <= m; j += HUNG_THREADS) {
        if (used[j]) { u[p[j]] += delta; v[j] -= delta; }     // used columns hold distinct rows: no write conflict
        else minv[j] -= delta;
      }
      __syncthreads();
      j0 = j1;
      const int matched = p[j0];
      __syncthreads();             // every thread has read p[j0] before thread 0 rewrites p along the path
      if (matched == 0) break;
    }
    if (tid == 0) {            // augment along the alternating path
      while (j0) { const int jp = way[j0]; p[j0] = p[jp]; j0 = jp; }
    }
    __syncthreads();
  }
  for (int j = tid + 1; j <= m; j += HUNG_THREADS)
    if (p[j] > 0) {
      out[j - 1] = p[j] - 1;
      if (gt_to_pred != nullptr) gt_to_pred[(long long)prob * ld_gt + p[j] - 1] = j - 1;
    }
}

}  // namespace

extern "C" int esb_hungarian_batch(const float* cost, const int* n_gt, int n_problems, int n_pred, int ld_gt,
                                   int* pred_to_gt, int* gt_to_pred, void* stream) {
  ESB_CHECK_ARG(n_problems >= 0 && n_pred >= 1 && ld_gt >= 1, "esb_hungarian_batch: bad sizes");
  ESB_CHECK_ARG(ld_gt <= n_pred, "esb_hungarian_batch: needs n_gt <= n_pred (got ld_gt=%d, n_pred=%d)", ld_gt, n_pred);
  if (n_problems == 0) return ESB_OK;
  const size_t smem = sizeof(double) * ((size_t)(ld_gt + 1) + 2 * (size_t)(n_pred + 1)) +
                      sizeof(int) * 2 * (size_t)(n_pred + 1) + (size_t)(n_pred + 1);
  ESB_CHECK_ARG(smem <= 200 * 1024, "esb_hungarian_batch: n_pred=%d too large for shared memory", n_pred);
  if (smem > 48 * 1024)
    ESB_CUDA_CALL(cudaFuncSetAttribute(hungarian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hungarian_kernel<<<n_problems, HUNG_THREADS, smem, (cudaStream_t)stream>>>(
      cost, n_gt, n_pred, ld_gt, (long long)n_pred * ld_gt, pred_to_gt, gt_to_pred);
  ESB_CUDA_LAUNCH_CHECK("esb_hungarian_batch");
  return ESB_OK;
}
