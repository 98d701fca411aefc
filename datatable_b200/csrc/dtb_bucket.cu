// dtb_bucket.cu -- bucketed multi-reducer: every reducer of one value column in ONE sweep, with
// shared-memory accumulators instead of one L2 atomic per row and reducer.
//
// Replaces, for DT[:, {mean(v), min(v), max(v), count(v), ...}, by(k...)], the reference's one full
// pass per reducer per column (column/latent.cc:103-135 materialising sumprod.h / mean.h / minmax.h /
// count.h one after the other: 12 passes for BASELINE config C4) -- and this engine's own first
// version, which streamed the rows once per reducer and folded them with L2 atomics (1-2 per row and
// reducer: 15e9 atomics at ~190 G/s = the 100 ms of C4's 122 ms in round 1).
//
// When the normalised group key x spans 2^12 .. 2^20 values and no key is hot:
//   count    : rows per bucket, bucket = x >> 11 (<= 512 buckets of 2048 consecutive keys)      [once per call]
//   scatter  : rows of the value column are partitioned by bucket -- (x & 2047 as uint16, raw value) --
//              tile by tile: shared-memory counters hand out the slots of a tile (order inside a bucket does
//              not matter to a reducer), one global atomic per (tile, bucket) reserves the output range,
//              the tile is staged in shared memory and written out in bucket runs
//   aggregate: a CTA walks a fixed-size chunk of the partitioned rows bucket by bucket, folds every
//              requested word (int sum, float sum, count, min, max, NA count) into 2048-slot
//              shared-memory tables and flushes the touched slots once per (chunk, bucket)
// Bytes per row and column: read 4 (x) + V, write 2 + V, read 2 + V  (V = value bytes) -- against
// (4 + V) per reducer before; L2 atomics: ~2048 * words per 262144 rows instead of 1-2 per row.
//
// Bound: LSU wavefronts (shared-memory atomics with ~3.5-way bank conflicts) and HBM, about evenly.
#include <type_traits>
#include "dtb_common.cuh"

namespace dtb {

constexpr int BK_BITS = 11;                       // keys per bucket = 2048
constexpr int BK_KEYS = 1 << BK_BITS;
constexpr int BK_MAXB = 512;                      // buckets: group key domain <= 2^20
constexpr int BK_THREADS = 512;                   // = BK_MAXB: thread t owns bucket t in the tile scan
constexpr int BK_IPT = 8;                         // (16 rows x 256 threads ran at 24 % occupancy: 48 registers of row
constexpr int BK_TILE = BK_THREADS * BK_IPT;      //  state per thread; profiles/r2_c4_launches_1e9.txt) 4096 rows per tile
constexpr int64_t BK_CHUNK = 262144;              // rows per aggregate CTA

static inline int bk_grid(int64_t n, int threads) {
  const int64_t want = (n + threads - 1) / threads;
  return (int)(want > NUM_SMS_B200 * 16 ? NUM_SMS_B200 * 16 : (want < 1 ? 1 : want));
}

// ---- rows per (slab, bucket) -> start[slab][bucket] ------------------------------------------------
// A slab is a contiguous range of tiles.  Every slab gets its own output range inside every bucket
// (buckets stay contiguous: slab 0's rows, then slab 1's, ...), so that the scatter tiles of one slab
// reserve their runs from a cursor only they share: with ONE cursor per bucket the 244 k tiles of a 1e9-row
// column issued 244 k same-address L2 atomics-with-return per bucket (7.4 ms per column, LSU half idle).
__global__ void __launch_bounds__(512)
bucket_count_kernel(const u32* __restrict__ xkeys, int gshift, int64_t n, int64_t slab_rows, u32* __restrict__ hist /*[nslabs][BK_MAXB]*/)
{
  __shared__ u32 h[BK_MAXB];
  for (int i = threadIdx.x; i < BK_MAXB; i += blockDim.x) h[i] = 0;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * slab_rows;
  const int64_t r1 = (r0 + slab_rows < n) ? r0 + slab_rows : n;
  for (int64_t i = r0 + threadIdx.x; i < r1; i += blockDim.x)
    atomicAdd(&h[(xkeys[i] >> gshift) >> BK_BITS], 1u);
  __syncthreads();
  for (int i = threadIdx.x; i < BK_MAXB; i += blockDim.x) hist[(size_t)blockIdx.x * BK_MAXB + i] = h[i];
}

// start[slab][b] = rows of buckets < b (all slabs) + rows of bucket b in slabs < slab; bstart[b] = start[0][b], bstart[nb] = n
__global__ void __launch_bounds__(BK_MAXB)
bucket_scan_kernel(u32* __restrict__ hist /*in: counts, out: starts*/, int nslabs, int nb, u32* __restrict__ bstart /*[nb+1]*/)
{
  __shared__ u32 s[BK_MAXB];
  const int t = threadIdx.x;
  u32 tot = 0;
  if (t < nb) for (int c = 0; c < nslabs; c++) tot += hist[(size_t)c * BK_MAXB + t];
  s[t] = tot;
  __syncthreads();
  for (int d = 1; d < BK_MAXB; d <<= 1) {
    const u32 a = t >= d ? s[t - d] : 0;
    __syncthreads();
    s[t] += a;
    __syncthreads();
  }
  if (t < nb) {
    u32 run = s[t] - tot;
    bstart[t] = run;
    for (int c = 0; c < nslabs; c++) { const u32 v = hist[(size_t)c * BK_MAXB + t]; hist[(size_t)c * BK_MAXB + t] = run; run += v; }
    if (t == nb - 1) bstart[nb] = s[t];
  }
}

int64_t bucket_slab_rows(int64_t n) {
  // ~4 slabs per SM, whole tiles per slab
  int64_t r = (n + NUM_SMS_B200 * 4 - 1) / (NUM_SMS_B200 * 4);
  r = (r + BK_TILE - 1) / BK_TILE * BK_TILE;
  return r < BK_TILE ? BK_TILE : r;
}
int bucket_num_slabs(int64_t n) { const int64_t r = bucket_slab_rows(n); return (int)((n + r - 1) / r); }

// hist: u32[nslabs * 512] (becomes the per-slab cursors' initial values), bstart: u32[nb + 1]
int launch_bucket_starts(const u32* xkeys, int gshift, int64_t n, int nb, u32* hist, u32* bstart, cudaStream_t s)
{
  const int nslabs = bucket_num_slabs(n);
  prof_begin("bucket_count", s);
  bucket_count_kernel<<<nslabs, 512, 0, s>>>(xkeys, gshift, n, bucket_slab_rows(n), hist);
  prof_end(s);
  bucket_scan_kernel<<<1, BK_MAXB, 0, s>>>(hist, nslabs, nb, bstart);
  count_launch(2);
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

// ---- partition one value column by bucket ---------------------------------------------------------
template <typename L>
__global__ void __launch_bounds__(BK_THREADS, 3)
bucket_scatter_kernel(const u32* __restrict__ xkeys, int gshift, const L* __restrict__ v, int64_t n, int64_t slab_rows,
                      u32* __restrict__ cursors /*[nslabs][BK_MAXB]*/, unsigned short* __restrict__ xlow_out, L* __restrict__ v_out)
{
  __shared__ u32 cnt[BK_MAXB];                 // rows of the bucket in this tile; then: tile slot of its first row
  __shared__ u32 gbase[BK_MAXB];               // (reserved global slot) - (tile slot) of the bucket
  __shared__ u32 wsum[BK_THREADS / 32];
  extern __shared__ __align__(16) unsigned char bk_stage[];      // staged tile: values, then group keys
  L* sv = reinterpret_cast<L*>(bk_stage);
  u32* sx = reinterpret_cast<u32*>(bk_stage + sizeof(L) * BK_TILE);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)blockIdx.x * BK_TILE;
  u32* cursor = cursors + (size_t)(base / slab_rows) * BK_MAXB;       // the slab's own reservation cursors
  const int tile_n = (int)((n - base) < (int64_t)BK_TILE ? (n - base) : (int64_t)BK_TILE);
  for (int i = tid; i < BK_MAXB; i += BK_THREADS) cnt[i] = 0;
  __syncthreads();

  u32 x[BK_IPT]; L val[BK_IPT]; unsigned short r[BK_IPT];
#pragma unroll
  for (int i = 0; i < BK_IPT; i++) {
    const int p = tid + i * BK_THREADS;
    x[i] = p < tile_n ? (xkeys[base + p] >> gshift) : 0xffffffffu;
    val[i] = p < tile_n ? v[base + p] : (L)0;
  }
#pragma unroll
  for (int i = 0; i < BK_IPT; i++)
    r[i] = (x[i] != 0xffffffffu) ? (unsigned short)atomicAdd(&cnt[x[i] >> BK_BITS], 1u) : (unsigned short)0;
  __syncthreads();

  // exclusive scan of cnt[] over the buckets (one per thread), one global reservation per non-empty bucket
  const u32 c0 = cnt[tid];
  u32 incl = c0;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  u32 wpre = 0;
#pragma unroll
  for (int w = 0; w < BK_THREADS / 32; w++) if (w < warp) wpre += wsum[w];
  const u32 e0 = wpre + incl - c0;
  __syncthreads();
  cnt[tid] = e0;
  if (c0) gbase[tid] = atomicAdd(&cursor[tid], c0) - e0;
  __syncthreads();

#pragma unroll
  for (int i = 0; i < BK_IPT; i++) {
    if (x[i] != 0xffffffffu) {
      const u32 slot = cnt[x[i] >> BK_BITS] + r[i];
      sx[slot] = x[i]; sv[slot] = val[i];
    }
  }
  __syncthreads();
  for (int p = tid; p < tile_n; p += BK_THREADS) {
    const u32 xx = sx[p];
    const u32 dst = gbase[xx >> BK_BITS] + (u32)p;
    xlow_out[dst] = (unsigned short)(xx & (BK_KEYS - 1));
    v_out[dst] = sv[p];
  }
}

// ---- aggregate ---------------------------------------------------------------------------------------
struct BucketAcc { u64* w[BK_NWORDS]; };        // global accumulator tables, indexed by group key x (NULL = not requested)

template <typename T>
__global__ void __launch_bounds__(512)
bucket_aggregate_kernel(const unsigned short* __restrict__ xlow, const typename RawKey<T>::load_t* __restrict__ v,
                        const u32* __restrict__ start, int nb, int64_t n, BucketAcc acc)
{
  constexpr bool ISF = std::is_floating_point<T>::value;
  extern __shared__ __align__(16) u64 sacc[];   // [requested word][BK_KEYS]
  __shared__ int s_b;
  u64* sw[BK_NWORDS];
  int nw = 0;
#pragma unroll
  for (int w = 0; w < BK_NWORDS; w++) sw[w] = acc.w[w] ? sacc + (size_t)(nw++) * BK_KEYS : nullptr;

  const int64_t c0 = (int64_t)blockIdx.x * BK_CHUNK;
  const int64_t c1 = (c0 + BK_CHUNK < n) ? c0 + BK_CHUNK : n;
  if (threadIdx.x == 0) {                       // bucket that holds row c0: largest b with start[b] <= c0
    int lo = 0, hi = nb;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int64_t)start[mid] <= c0) lo = mid; else hi = mid; }
    s_b = lo;
  }
  __syncthreads();
  for (int b = s_b; b < nb; b++) {
    const int64_t bs = start[b], be = start[b + 1];
    if (bs >= c1) break;
    const int64_t lo = bs > c0 ? bs : c0, hi = be < c1 ? be : c1;
    if (lo >= hi) continue;
    for (int k = threadIdx.x; k < BK_KEYS; k += blockDim.x) {
#pragma unroll
      for (int w = 0; w < BK_NWORDS; w++) if (sw[w]) sw[w][k] = (w == BK_MIN) ? ~0ull : 0ull;
    }
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
      const int k = xlow[i];
      u64 u; const bool valid = RawKey<T>::get(v[i], u);          // u: sign-extended int or float image
      // counts are 32-bit shared-memory atomics on the low word of the slot (a chunk holds 262144 rows)
      if (!valid) { if (sw[BK_CNTNA]) atomicAdd(reinterpret_cast<u32*>(&sw[BK_CNTNA][k]), 1u); continue; }
      if (sw[BK_CNT]) atomicAdd(reinterpret_cast<u32*>(&sw[BK_CNT][k]), 1u);
      if (sw[BK_SUMI]) atomicAdd(&sw[BK_SUMI][k], u);
      if (sw[BK_SUMF]) {
        double d;
        if constexpr (std::is_same<T, float>::value) d = (double)__uint_as_float((u32)v[i]);
        else if constexpr (std::is_same<T, double>::value) d = __longlong_as_double((long long)v[i]);
        else d = (double)(int64_t)u;
        atomicAdd(reinterpret_cast<double*>(&sw[BK_SUMF][k]), d);
      }
      if (sw[BK_MIN] || sw[BK_MAX]) {
        const u64 key = ISF ? u : (u ^ 0x8000000000000000ull);    // same encodings as dtb_reduce.cu:p_add
        // after the first few rows of a key most rows improve neither bound: look before the atomic
        if (sw[BK_MIN]) { const u64 km = ISF ? key : key - 1; if (km < sw[BK_MIN][k]) atomicMin(&sw[BK_MIN][k], km); }
        if (sw[BK_MAX]) { if (key > sw[BK_MAX][k]) atomicMax(&sw[BK_MAX][k], key); }
      }
    }
    __syncthreads();
    const u64 xb = (u64)b << BK_BITS;
    for (int k = threadIdx.x; k < BK_KEYS; k += blockDim.x) {
      if (sw[BK_CNT])   { const u64 a = sw[BK_CNT][k];   if (a) atomicAdd(&acc.w[BK_CNT][xb + k], a); }
      if (sw[BK_CNTNA]) { const u64 a = sw[BK_CNTNA][k]; if (a) atomicAdd(&acc.w[BK_CNTNA][xb + k], a); }
      if (sw[BK_SUMI])  { const u64 a = sw[BK_SUMI][k];  if (a) atomicAdd(&acc.w[BK_SUMI][xb + k], a); }
      if (sw[BK_SUMF])  { const double d = __longlong_as_double((long long)sw[BK_SUMF][k]);
                          if (d != 0.0) atomicAdd(reinterpret_cast<double*>(acc.w[BK_SUMF]) + xb + k, d); }
      if (sw[BK_MIN])   { const u64 a = sw[BK_MIN][k];   if (a != ~0ull) atomicMin(&acc.w[BK_MIN][xb + k], a); }
      if (sw[BK_MAX])   { const u64 a = sw[BK_MAX][k];   if (a != 0ull)  atomicMax(&acc.w[BK_MAX][xb + k], a); }
    }
    __syncthreads();
  }
}

size_t bucket_scratch_bytes(int64_t n, int value_bytes) {
  return ((size_t)n * 2 + 255) / 256 * 256 + ((size_t)n * value_bytes + 255) / 256 * 256 + bucket_starts_bytes(n);
}
size_t bucket_starts_bytes(int64_t n) { return sizeof(u32) * ((size_t)bucket_num_slabs(n) * BK_MAXB + 8); }

// acc_w[w]: global table of (1 << dbits) u64 for every requested word (NULL otherwise), already set to the
// word's identity (~0 for BK_MIN, 0 otherwise).  slab_starts / bstart: from launch_bucket_starts.
int launch_bucketed_reduce(const u32* xkeys, int gshift, int dbits, const void* value, int stype, int64_t n,
                           const u32* slab_starts, const u32* start, unsigned long long* const* acc_w, void* scratch,
                           cudaStream_t s)
{
  if (n == 0) return DTB_OK;
  const int nb = 1 << (dbits > BK_BITS ? dbits - BK_BITS : 0);
  if (nb > BK_MAXB) { set_error("internal: bucketed reducer needs a group key domain of at most 2^20"); return DTB_EINVAL; }
  const int esz = stype_bytes(stype);
  unsigned short* xlow = (unsigned short*)scratch;
  char* vpart = (char*)scratch + ((size_t)n * 2 + 255) / 256 * 256;
  u32* cursor = (u32*)(vpart + ((size_t)n * esz + 255) / 256 * 256);
  const int64_t slab_rows = bucket_slab_rows(n);
  DTB_CUDA_CHECK(cudaMemcpyAsync(cursor, slab_starts, sizeof(u32) * (size_t)bucket_num_slabs(n) * BK_MAXB, cudaMemcpyDeviceToDevice, s));
  const unsigned tiles = (unsigned)((n + BK_TILE - 1) / BK_TILE);
  if (esz == 8) DTB_CUDA_CHECK(cudaFuncSetAttribute(bucket_scatter_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * BK_TILE));
  prof_begin("bucket_scatter", s);
  switch (esz) {
    case 1: bucket_scatter_kernel<uint8_t><<<tiles, BK_THREADS, (sizeof(uint8_t) + 4) * BK_TILE, s>>>(xkeys, gshift, (const uint8_t*)value, n, slab_rows, cursor, xlow, (uint8_t*)vpart); break;
    case 2: bucket_scatter_kernel<uint16_t><<<tiles, BK_THREADS, (sizeof(uint16_t) + 4) * BK_TILE, s>>>(xkeys, gshift, (const uint16_t*)value, n, slab_rows, cursor, xlow, (uint16_t*)vpart); break;
    case 4: bucket_scatter_kernel<u32><<<tiles, BK_THREADS, (sizeof(u32) + 4) * BK_TILE, s>>>(xkeys, gshift, (const u32*)value, n, slab_rows, cursor, xlow, (u32*)vpart); break;
    case 8: bucket_scatter_kernel<u64><<<tiles, BK_THREADS, (sizeof(u64) + 4) * BK_TILE, s>>>(xkeys, gshift, (const u64*)value, n, slab_rows, cursor, xlow, (u64*)vpart); break;
    default: set_error("unsupported stype"); return DTB_ENOTIMPL;
  }
  prof_end(s);
  count_launch();
  BucketAcc acc; int nw = 0;
  for (int w = 0; w < BK_NWORDS; w++) { acc.w[w] = acc_w[w]; nw += acc_w[w] != nullptr; }
  const size_t smem = (size_t)nw * BK_KEYS * sizeof(u64);
  const unsigned chunks = (unsigned)((n + BK_CHUNK - 1) / BK_CHUNK);
  prof_begin("bucket_aggregate", s);
#define DTB_AGG(T) { DTB_CUDA_CHECK(cudaFuncSetAttribute(bucket_aggregate_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                     bucket_aggregate_kernel<T><<<chunks, 512, smem, s>>>(xlow, (const typename RawKey<T>::load_t*)vpart, start, nb, n, acc); }
  switch (stype) {
    case DTB_STYPE_BOOL: case DTB_STYPE_INT8:    DTB_AGG(int8_t)  break;
    case DTB_STYPE_INT16:                        DTB_AGG(int16_t) break;
    case DTB_STYPE_INT32: case DTB_STYPE_DATE32: DTB_AGG(int32_t) break;
    case DTB_STYPE_INT64: case DTB_STYPE_TIME64: DTB_AGG(int64_t) break;
    case DTB_STYPE_FLOAT32:                      DTB_AGG(float)   break;
    case DTB_STYPE_FLOAT64:                      DTB_AGG(double)  break;
    default: set_error("unsupported stype"); return DTB_ENOTIMPL;
  }
#undef DTB_AGG
  prof_end(s);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

}  // namespace dtb
