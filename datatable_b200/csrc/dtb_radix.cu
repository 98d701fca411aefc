// dtb_radix.cu -- stable LSD radix sort passes over normalised composite keys.
//
// Replaces SortContext::build_histogram / reorder_data / radix_psort /
// _radix_recurse (sort.cc:950-1353).  The reference is an MSD recursion over
// a chunk x radix size_t histogram with insertion-sort leaves -- a CPU idiom.
// Here every pass is ONE kernel that reads each (key, row) pair once and
// writes it once ("single sweep"): tiles take ticket numbers from an atomic
// counter, rank their rows with warp match/ballot histogramming in shared
// memory, publish per-digit tile counts and resolve the global digit offsets
// with a decoupled look-back over earlier tiles' status words.  Stability
// (ties keep ascending row index, sort.cc:27-33) follows from ranking rows in
// (item, lane) = position order inside a tile and tiles in ticket order.
//
// Key normalisation (sort.cc:690-845) is evaluated on the fly in the first
// pass and in the histogram kernel: no separate `x` array is materialised for
// single-column keys.
//
// Bound: HBM.  Algorithmic bytes per row per pass = read (key + idx) + write
// (key + idx); first pass reads the raw column only, last pass of a sort-only
// call writes idx only.
#include "dtb_common.cuh"

namespace dtb {

// ===========================================================================
// Composite key materialisation (multi-column keys)
// ===========================================================================
template <typename KeyT>
__global__ void __launch_bounds__(256)
compose_keys_kernel(KeyPlan kp, int64_t n, const int32_t* __restrict__ idx, KeyT* __restrict__ out)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t row = idx ? (int64_t)idx[i] : i;     // later rounds see the rows in the current order
    u64 x = 0;
    for (int c = 0; c < kp.nkeys; c++)
      x |= norm_load_dynamic(kp.k[c], row) << kp.k[c].lshift;
    out[i] = (KeyT)x;
  }
}

int launch_compose_keys(const KeyPlan& kp, int64_t n, const int32_t* idx, void* keys_out, int key_bytes,
                        cudaStream_t s)
{
  if (n == 0) return DTB_OK;
  int64_t want = (n + 255) / 256;
  int grid = (int)(want > NUM_SMS_B200 * 16 ? NUM_SMS_B200 * 16 : want);
  if (key_bytes == 4) compose_keys_kernel<u32><<<grid, 256, 0, s>>>(kp, n, idx, (u32*)keys_out);
  else                compose_keys_kernel<u64><<<grid, 256, 0, s>>>(kp, n, idx, (u64*)keys_out);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

// ===========================================================================
// Digit histograms for all passes (one read of the key source)
// ===========================================================================
struct HistPlan {
  int npasses;
  int shift[MAX_PASSES];
  u32 mask[MAX_PASSES];
};

template <typename KeyT, typename Src, int NBINS>
__global__ void __launch_bounds__(512)
histogram_kernel(Src src, int64_t n, HistPlan hp, u32* __restrict__ ghist)
{
  extern __shared__ u32 shist[];            // [npasses][NBINS]
  const int nwords = hp.npasses * NBINS;
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) shist[i] = 0;
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    KeyT x = src.load(i);
    for (int p = 0; p < hp.npasses; p++) {
      u32 d = (u32)(x >> hp.shift[p]) & hp.mask[p];
      atomicAdd(&shist[p * NBINS + d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) {
    u32 c = shist[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

template <typename KeyT, typename Src>
static int run_hist(Src src, int64_t n, const PassPlan& pp, int nbins_log2, u32* hist, cudaStream_t s)
{
  HistPlan hp; hp.npasses = pp.npasses;
  for (int p = 0; p < pp.npasses; p++) { hp.shift[p] = pp.shift[p]; hp.mask[p] = (1u << pp.bits[p]) - 1; }
  const int nbins = 1 << nbins_log2;
  DTB_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(u32) * pp.npasses * nbins, s));
  if (n == 0) return DTB_OK;
  int64_t want = (n + 512 * 16 - 1) / (512 * 16);
  int grid = (int)(want < 1 ? 1 : (want > NUM_SMS_B200 * 4 ? NUM_SMS_B200 * 4 : want));
  size_t smem = sizeof(u32) * pp.npasses * nbins;
  if (nbins_log2 != 8) { set_error("internal: only 8-bit digit kernels are built"); return DTB_EINVAL; }
  histogram_kernel<KeyT, Src, 256><<<grid, 512, smem, s>>>(src, n, hp, hist);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

template <typename KeyT>
static int run_hist_raw(const KeyPlan& kp, int64_t n, const PassPlan& pp, int nbins_log2, u32* hist,
                        cudaStream_t s)
{
  const KeyNorm& k = kp.k[0];
#define DTB_CASE(T)                                                                          \
  { RawSrc<T, KeyT> src; src.init(k);                                                        \
    return run_hist<KeyT>(src, n, pp, nbins_log2, hist, s); }
  switch (k.stype) {
    case DTB_STYPE_BOOL: case DTB_STYPE_INT8:    DTB_CASE(int8_t)
    case DTB_STYPE_INT16:                        DTB_CASE(int16_t)
    case DTB_STYPE_INT32: case DTB_STYPE_DATE32: DTB_CASE(int32_t)
    case DTB_STYPE_INT64: case DTB_STYPE_TIME64: DTB_CASE(int64_t)
    case DTB_STYPE_FLOAT32:                      DTB_CASE(float)
    case DTB_STYPE_FLOAT64:                      DTB_CASE(double)
  }
#undef DTB_CASE
  set_error("internal: bad stype in histogram"); return DTB_EINVAL;
}

int launch_histograms(int src_kind, const void* packed, const KeyPlan& kp, int key_bytes,
                      int64_t n, const PassPlan& pp, int nbins_log2, uint32_t* hist, cudaStream_t s)
{
  if (src_kind == 0) {
    if (key_bytes == 4) { PackedSrc<u32> src{(const u32*)packed}; return run_hist<u32>(src, n, pp, nbins_log2, hist, s); }
    else                { PackedSrc<u64> src{(const u64*)packed}; return run_hist<u64>(src, n, pp, nbins_log2, hist, s); }
  }
  return key_bytes == 4 ? run_hist_raw<u32>(kp, n, pp, nbins_log2, hist, s)
                        : run_hist_raw<u64>(kp, n, pp, nbins_log2, hist, s);
}

// Exclusive scan of each pass' NBINS counters (one block per pass).
__global__ void scan_hist_kernel(u32* hist, int nbins, u32* hmax)
{
  __shared__ u32 wsum[32];
  __shared__ u32 wmax[32];
  u32* h = hist + (size_t)blockIdx.x * nbins;
  // nbins <= 1024 = blockDim.x
  const int t = threadIdx.x;
  u32 v = t < nbins ? h[t] : 0;
  {
    // largest digit count of this pass (skew detector for the direct-address reducers)
    u32 m = v;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { const u32 o = __shfl_xor_sync(0xffffffffu, m, d); m = o > m ? o : m; }
    if ((t & 31) == 0) wmax[t >> 5] = m;
    __syncthreads();
    if (t == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); w++) m = wmax[w] > m ? wmax[w] : m;
      if (hmax) hmax[blockIdx.x] = m;
    }
  }
  u32 incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 o = __shfl_up_sync(0xffffffffu, incl, d);
    if ((t & 31) >= d) incl += o;
  }
  if ((t & 31) == 31) wsum[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    u32 w = t < (blockDim.x >> 5) ? wsum[t] : 0;
    u32 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 o = __shfl_up_sync(0xffffffffu, wi, d);
      if (t >= d) wi += o;
    }
    wsum[t] = wi - w;
  }
  __syncthreads();
  if (t < nbins) h[t] = incl - v + wsum[t >> 5];
}

int launch_scan_histograms(uint32_t* hist, int npasses, int nbins_log2, uint32_t* hmax, cudaStream_t s)
{
  const int nbins = 1 << nbins_log2;
  int threads = nbins < 32 ? 32 : nbins;
  scan_hist_kernel<<<npasses, threads, 0, s>>>(hist, nbins, hmax);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

// ===========================================================================
// The single-sweep scatter pass
// ===========================================================================
constexpr u32 ST_FLAG_AGG  = 1u << 30;     // tile count published
constexpr u32 ST_FLAG_INCL = 2u << 30;     // inclusive prefix published
constexpr u32 ST_MASK      = (1u << 30) - 1;

template <typename KeyT, typename Src>
struct PassArgs {
  Src            src;
  const int32_t* idx_in;        // NULL = identity
  KeyT*          keys_out;      // NULL = do not write keys
  int32_t*       idx_out;
  int64_t        n;
  int            shift;
  u32            mask;
  const u32*     bin_start;     // [NBINS] global exclusive digit offsets
  u32*           status;        // [ntiles][NBINS]
  u32*           tile_counter;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// One tile.  FULL = the tile holds TILE rows (no per-row validity predicates).
//
// Shared memory: whist[WARPS][NBINS] u16 | bin_dst[NBINS] u32 | skey[TILE] | sidx[TILE] | ridx[TILE]
// Registers hold only the 16 keys and their 16-bit ranks; the incoming row ids are
// prefetched straight into shared memory with cp.async (no register staging), so the
// kernel fits 4 CTAs = 32 warps per SM.
template <typename KeyT, typename Src, int NBINS, int THREADS, int IPT, bool FULL>
__device__ __forceinline__ void radix_pass_tile(const PassArgs<KeyT, Src>& a, unsigned char* smem_raw,
                                                u32* s_wsum, const u32 tile)
{
  constexpr int WARPS = THREADS / 32;
  constexpr int TILE = THREADS * IPT;
  static_assert(NBINS == THREADS, "one digit per thread in the scan phase");

  unsigned short* whist = reinterpret_cast<unsigned short*>(smem_raw);
  u32* bin_dst    = reinterpret_cast<u32*>(smem_raw + sizeof(unsigned short) * WARPS * NBINS);
  KeyT* skey      = reinterpret_cast<KeyT*>(bin_dst + NBINS + 4);
  int32_t* sidx   = reinterpret_cast<int32_t*>(skey + TILE);
  int32_t* ridx   = sidx + TILE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = (int64_t)tile * TILE;
  const int tile_n = FULL ? TILE : (int)(a.n - base);
  const bool have_idx = a.idx_in != nullptr;

  // ---- prefetch the incoming row ids into shared memory (consumed in the reorder phase) ----
  if (have_idx) {
    const int32_t* g = a.idx_in + base;
    if (FULL) {
#pragma unroll
      for (int j = 0; j < IPT / 4; j++) {
        const int c = tid + j * THREADS;                       // 16-byte chunk index
        cp_async16(ridx + 4 * c, g + 4 * c);
      }
    } else {
      for (int p = tid; p < tile_n; p += THREADS) ridx[p] = g[p];
    }
  }

  // ---- load keys (warp-striped: item i of lane l sits at warp_base + i*32 + l) ----
  KeyT key[IPT];
  const int wbase = warp * 32 * IPT;
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    const int lp = wbase + i * 32 + lane;
    key[i] = (FULL || lp < tile_n) ? a.src.load(base + lp) : (KeyT)0;
  }

  // ---- rank inside the warp: rows with equal digits keep (item, lane) order ----
  // The lanes holding the same digit are found with a shared-memory atomicOr on a per-warp
  // mask table, not MATCH.ANY: on sm_100 MATCH.ANY issues once per ~60 SM cycles and bound
  // the whole kernel; the atomicOr sequence costs ~7 (scripts/ubench/match_bench.cu).
  unsigned short rank[IPT];
  unsigned short* myhist = whist + warp * NBINS;
  u32* wmask = reinterpret_cast<u32*>(skey) + warp * NBINS;    // skey/sidx are idle until the reorder phase
  const unsigned lt = lanemask_lt();
  const unsigned lanebit = 1u << lane;
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    const bool valid = FULL || (wbase + i * 32 + lane) < tile_n;
    const u32 d = (u32)(key[i] >> a.shift) & a.mask;
    if (valid) atomicOr(&wmask[d], lanebit);
    __syncwarp();
    unsigned peers = 0; unsigned short cnt = 0;
    if (valid) { peers = wmask[d]; cnt = myhist[d]; }
    const unsigned before = peers & lt;
    rank[i] = cnt + (unsigned short)__popc(before);
    __syncwarp();                                               // all reads of this round precede the update
    if (valid && before == 0) { myhist[d] = cnt + (unsigned short)__popc(peers); wmask[d] = 0; }
    __syncwarp();
  }
  __syncthreads();

  // ---- per digit (thread b owns digit b): prefix over warps, publish, scan over digits, look back ----
  const int b = tid;
  u32 run = 0;
#pragma unroll
  for (int w = 0; w < WARPS; w++) {
    const unsigned short c = whist[w * NBINS + b];
    whist[w * NBINS + b] = (unsigned short)run;
    run += c;
  }
  st_relaxed_u32(&a.status[(size_t)tile * NBINS + b], (tile == 0 ? ST_FLAG_INCL : ST_FLAG_AGG) | run);

  u32 incl = run;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const u32 w = lane < WARPS ? s_wsum[lane] : 0;
    u32 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u32 o = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += o;
    }
    s_wsum[lane] = wi - w;
  }
  __syncthreads();
  const u32 tstart = incl - run + s_wsum[warp];                  // first slot of digit b inside the tile
#pragma unroll
  for (int w = 0; w < WARPS; w++) whist[w * NBINS + b] += (unsigned short)tstart;

  // Decoupled look-back, LB_W predecessors per step: the loads of a window are independent, so a
  // walk of depth D costs ~D/LB_W L2 round trips instead of D.
  constexpr int LB_W = 4;
  u32 prev = 0;
  if (tile > 0) {
    const u32* sp = a.status + (size_t)tile * NBINS + b;         // sp[-j*NBINS] = tile-j
    int64_t left = (int64_t)tile;                                 // predecessors not yet visited
    bool done = false;
    while (!done) {
      u32 sv[LB_W];
#pragma unroll
      for (int j = 0; j < LB_W; j++)
        sv[j] = (j < left) ? ld_relaxed_u32(sp - (size_t)(j + 1) * NBINS) : ST_FLAG_INCL;
#pragma unroll
      for (int j = 0; j < LB_W; j++) {
        if (!done) {
          u32 x = sv[j];
          while ((x & ~ST_MASK) == 0) x = ld_relaxed_u32(sp - (size_t)(j + 1) * NBINS);   // holds a ticket: is running
          prev += x & ST_MASK;
          if ((x & ~ST_MASK) == ST_FLAG_INCL) done = true;
        }
      }
      sp -= (size_t)LB_W * NBINS; left -= LB_W;
    }
    st_relaxed_u32(&a.status[(size_t)tile * NBINS + b], ST_FLAG_INCL | ((prev + run) & ST_MASK));
  }
  bin_dst[b] = a.bin_start[b] + prev - tstart;
  if (have_idx && FULL) cp_async_commit_wait_all();
  __syncthreads();

  // ---- reorder the tile in shared memory ----
#pragma unroll
  for (int i = 0; i < IPT; i++) {
    const int pos = wbase + i * 32 + lane;
    if (FULL || pos < tile_n) {
      const u32 d = (u32)(key[i] >> a.shift) & a.mask;
      const u32 lp = (u32)myhist[d] + rank[i];
      skey[lp] = key[i];
      sidx[lp] = have_idx ? ridx[pos] : (int32_t)(base + pos);
    }
  }
  __syncthreads();

  // ---- coalesced scatter: consecutive threads write consecutive slots of a digit run ----
#pragma unroll 4
  for (int p = tid; p < tile_n; p += THREADS) {
    const KeyT k = skey[p];
    const u32 d = (u32)(k >> a.shift) & a.mask;
    const u32 dst = bin_dst[d] + (u32)p;
    if (a.keys_out) a.keys_out[dst] = k;
    a.idx_out[dst] = sidx[p];
  }
}

template <typename KeyT, typename Src, int NBINS, int THREADS, int IPT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
radix_pass_kernel(const __grid_constant__ PassArgs<KeyT, Src> a)
{
  constexpr int TILE = THREADS * IPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ u32 s_wsum[32];
  __shared__ u32 s_ticket;

  if (threadIdx.x == 0) s_ticket = atomicAdd(a.tile_counter, 1u);
  {
    constexpr int WARPS = THREADS / 32;
    u32* z = reinterpret_cast<u32*>(smem_raw);                              // whist
    for (int i = threadIdx.x; i < WARPS * NBINS / 2; i += THREADS) z[i] = 0;
    u32* m = reinterpret_cast<u32*>(smem_raw + sizeof(unsigned short) * WARPS * NBINS) + NBINS + 4;   // = skey
    for (int i = threadIdx.x; i < WARPS * NBINS; i += THREADS) m[i] = 0;    // per-warp peer masks
  }
  __syncthreads();
  const u32 tile = s_ticket;
  const int64_t base = (int64_t)tile * TILE;
  if (a.n - base >= (int64_t)TILE)
    radix_pass_tile<KeyT, Src, NBINS, THREADS, IPT, true>(a, smem_raw, s_wsum, tile);
  else
    radix_pass_tile<KeyT, Src, NBINS, THREADS, IPT, false>(a, smem_raw, s_wsum, tile);
}

template <typename KeyT> struct PassCfg;
template <> struct PassCfg<u32> { static constexpr int THREADS = 256, IPT = 16, MINB = 4; };
template <> struct PassCfg<u64> { static constexpr int THREADS = 256, IPT = 16, MINB = 3; };

template <typename KeyT, int NBINS>
static constexpr size_t pass_smem_bytes() {
  return sizeof(unsigned short) * (PassCfg<KeyT>::THREADS / 32) * NBINS + sizeof(u32) * (NBINS + 4)
       + (sizeof(KeyT) + 2 * sizeof(int32_t)) * PassCfg<KeyT>::THREADS * PassCfg<KeyT>::IPT;
}

int radix_pass_tile_rows(int key_bytes, int /*nbins_log2*/) {
  return key_bytes == 4 ? PassCfg<u32>::THREADS * PassCfg<u32>::IPT
                        : PassCfg<u64>::THREADS * PassCfg<u64>::IPT;
}

template <typename KeyT, typename Src>
static int run_pass(Src src, const PassIO& io, int64_t n, int shift, int bits,
                    const u32* bin_start, u32* status, u32* tile_counter, cudaStream_t s)
{
  constexpr int NBINS = 256;
  constexpr int THREADS = PassCfg<KeyT>::THREADS, IPT = PassCfg<KeyT>::IPT, MINB = PassCfg<KeyT>::MINB;
  if (n == 0) return DTB_OK;
  if (io.idx_in && (reinterpret_cast<uintptr_t>(io.idx_in) & 15)) {
    set_error("internal: row-id buffer must be 16-byte aligned"); return DTB_EINVAL;
  }
  PassArgs<KeyT, Src> a;
  a.src = src; a.idx_in = io.idx_in; a.keys_out = (KeyT*)io.keys_out; a.idx_out = io.idx_out;
  a.n = n; a.shift = shift; a.mask = (1u << bits) - 1;
  a.bin_start = bin_start; a.status = status; a.tile_counter = tile_counter;
  const int64_t ntiles = (n + THREADS * IPT - 1) / (THREADS * IPT);
  constexpr size_t smem = pass_smem_bytes<KeyT, NBINS>();
  auto kern = radix_pass_kernel<KeyT, Src, NBINS, THREADS, IPT, MINB>;
  static bool configured = false;   // per instantiation
  if (!configured) {
    DTB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  kern<<<(unsigned)ntiles, THREADS, smem, s>>>(a);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

template <typename KeyT>
static int run_pass_raw(const PassIO& io, const KeyPlan& kp, int64_t n, int shift, int bits,
                        const u32* bin_start, u32* status, u32* tile_counter, cudaStream_t s)
{
  const KeyNorm& k = kp.k[0];
#define DTB_CASE(T)                                                                          \
  { RawSrc<T, KeyT> src; src.init(k);                                                        \
    return run_pass<KeyT>(src, io, n, shift, bits, bin_start, status, tile_counter, s); }
  switch (k.stype) {
    case DTB_STYPE_BOOL: case DTB_STYPE_INT8:    DTB_CASE(int8_t)
    case DTB_STYPE_INT16:                        DTB_CASE(int16_t)
    case DTB_STYPE_INT32: case DTB_STYPE_DATE32: DTB_CASE(int32_t)
    case DTB_STYPE_INT64: case DTB_STYPE_TIME64: DTB_CASE(int64_t)
    case DTB_STYPE_FLOAT32:                      DTB_CASE(float)
    case DTB_STYPE_FLOAT64:                      DTB_CASE(double)
  }
#undef DTB_CASE
  set_error("internal: bad stype in radix pass"); return DTB_EINVAL;
}

int launch_radix_pass(const PassIO& io, const KeyPlan& kp, int key_bytes, int64_t n,
                      int shift, int bits, int nbins_log2,
                      const uint32_t* bin_start, uint32_t* status, uint32_t* tile_counter,
                      cudaStream_t s)
{
  if (nbins_log2 != 8) { set_error("internal: only 8-bit digit kernels are built"); return DTB_EINVAL; }
  if (n >= (int64_t)ST_MASK) { set_error("nrows too large for 30-bit look-back words"); return DTB_ENOTIMPL; }
  if (io.src_kind == 0) {
    if (key_bytes == 4) { PackedSrc<u32> src{(const u32*)io.keys_in};
      return run_pass<u32>(src, io, n, shift, bits, bin_start, status, tile_counter, s); }
    else { PackedSrc<u64> src{(const u64*)io.keys_in};
      return run_pass<u64>(src, io, n, shift, bits, bin_start, status, tile_counter, s); }
  }
  return key_bytes == 4 ? run_pass_raw<u32>(io, kp, n, shift, bits, bin_start, status, tile_counter, s)
                        : run_pass_raw<u64>(io, kp, n, shift, bits, bin_start, status, tile_counter, s);
}

__global__ void iota32_kernel(int32_t* out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int32_t)i;
}

int launch_iota32(int32_t* out, int64_t n, cudaStream_t s) {
  if (n == 0) return DTB_OK;
  int64_t want = (n + 255) / 256;
  int grid = (int)(want > NUM_SMS_B200 * 16 ? NUM_SMS_B200 * 16 : want);
  iota32_kernel<<<grid, 256, 0, s>>>(out, n);
  count_launch();
  DTB_CUDA_CHECK(cudaGetLastError());
  return DTB_OK;
}

}  // namespace dtb
