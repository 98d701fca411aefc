// dtb_tiefix.cu -- second half of the hybrid sort for wide single keys (float64, wide int64).
//
// An LSD sort of a 64-bit key costs 8 passes over (8-byte key, 4-byte row id) pairs.  For n rows
// only ~log2(n)+few leading bits of the normalised key discriminate; so wide keys are sorted by their
// TOP 32 bits with 4 passes over 32-bit keys, and the rows that still tie on those bits (for 1e9
// random doubles ~20 % of the rows, in runs of 2-3) are put in order by their LOW bits here:
//
//   tie_fix_kernel   : one thread per run head; runs of <= 32 rows are insertion-sorted (stable) by
//                      the low bits, gathered through the row ids; longer runs are queued
//   long_run_kernel  : one warp per queued run: if its rows do not all carry the same low bits the
//                      input is too clustered for this scheme -> *fallback = 1 and the caller redoes
//                      the sort with the plain 8-pass path (the result is always the exact stable
//                      order, sort.cc:27-33; only the time differs)
//
// The reference has no counterpart: its MSD recursion degenerates on float keys.
#include "dtb_common.cuh"

namespace dtb {

constexpr int TF_MAX = 32;                 // longest run sorted by a single thread

__global__ void __launch_bounds__(256)
tie_fix_kernel(const u32* __restrict__ tk, int32_t* __restrict__ order, int64_t start, int64_t n,
               KeyNorm k, u64 lowmask, u32* __restrict__ long_list, u32 long_cap, u32* __restrict__ counters)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
    const u32 t = tk[p];
    const bool head = (p == start || tk[p - 1] != t) && (p + 1 < n && tk[p + 1] == t);
    if (!head) continue;
    int len = 2;
    while (len <= TF_MAX && p + len < n && tk[p + len] == t) len++;
    if (len > TF_MAX) {                                        // queue for the warp-per-run check
      const u32 slot = atomicAdd(&counters[0], 1u);
      if (slot < long_cap) long_list[slot] = (u32)p; else counters[1] = 1;   // list overflow -> fallback
      continue;
    }
    u64 lo[TF_MAX]; int32_t row[TF_MAX];
    for (int j = 0; j < len; j++) {
      row[j] = order[p + j];
      lo[j] = norm_load_dynamic(k, (int64_t)row[j]) & lowmask;
    }
    bool moved = false;
    for (int j = 1; j < len; j++) {                            // stable insertion sort on the low bits
      const u64 l = lo[j]; const int32_t r = row[j];
      int i = j - 1;
      while (i >= 0 && lo[i] > l) { lo[i + 1] = lo[i]; row[i + 1] = row[i]; i--; moved = true; }
      lo[i + 1] = l; row[i + 1] = r;
    }
    if (moved) for (int j = 0; j < len; j++) order[p + j] = row[j];
  }
}

__global__ void __launch_bounds__(256)
long_run_kernel(const u32* __restrict__ tk, const int32_t* __restrict__ order, int64_t n, KeyNorm k,
                u64 lowmask, const u32* __restrict__ long_list, u32 long_cap, u32* __restrict__ counters)
{
  const int lane = threadIdx.x & 31;
  const u32 nlong = counters[0] < long_cap ? counters[0] : long_cap;
  const u32 warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (u32 w = warp0; w < nlong; w += nwarps) {
    const int64_t p = long_list[w];
    const u32 t = tk[p];
    u64 mn = ~0ull, mx = 0;
    for (int64_t q0 = p; q0 < n; q0 += 32) {
      const int64_t q = q0 + lane;
      const bool in = q < n && tk[q] == t;
      if (in) {
        const u64 l = norm_load_dynamic(k, (int64_t)order[q]) & lowmask;
        mn = l < mn ? l : mn; mx = l > mx ? l : mx;
      }
      if (__any_sync(0xffffffffu, !in)) break;                 // the run ends inside this window
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const u64 a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
      mn = a < mn ? a : mn; mx = b > mx ? b : mx;
    }
    if (lane == 0 && mn != mx) counters[1] = 1;                // a long run with different low bits
  }
}

// counters: device u32[2] = {number of long runs, fallback flag}, zeroed by the callee.
// Histogram of the leading 12 bits of the normalised key (feeds the dense rank table of HybridKey).
__global__ void __launch_bounds__(512)
top12_hist_kernel(KeyNorm k, int top_shift, int64_t n, u32* __restrict__ ghist)
{
  __shared__ u32 h[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&h[(u32)(norm_load_dynamic(k, i) >> top_shift)], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (h[i]) atomicAdd(&ghist[i], h[i]);
}

int launch_top12_histogram(const KeyNorm& k, int total_bits, int64_t n, uint32_t* hist, cudaStream_t s)
{
  DTB_CUDA_CHECK(cudaMemsetAsync(hist, 0, 4096 * sizeof(u32), s));
  if (n == 0) return DTB_OK;
  top12_hist_kernel<<<NUM_SMS_B200 * 4, 512, 0, s>>>(k, total_bits - 12, n, hist);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

int launch_tie_fix(const uint32_t* top_keys, int32_t* order, int64_t start, int64_t n, const KeyNorm& k,
                   int low_bits, uint32_t* long_list, uint32_t long_cap, uint32_t* counters, cudaStream_t s)
{
  DTB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(u32), s));
  if (n - start < 2) return DTB_OK;
  const u64 lowmask = low_bits >= 64 ? ~0ull : ((1ull << low_bits) - 1ull);
  int64_t want = (n - start + 255) / 256;
  int grid = (int)(want > NUM_SMS_B200 * 32 ? NUM_SMS_B200 * 32 : want);
  tie_fix_kernel<<<grid, 256, 0, s>>>(top_keys, order, start, n, k, lowmask, long_list, long_cap, counters);
  long_run_kernel<<<NUM_SMS_B200 * 4, 256, 0, s>>>(top_keys, order, n, k, lowmask, long_list, long_cap, counters);
  count_launch(2);
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

}  // namespace dtb
