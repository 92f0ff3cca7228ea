// =============================================================================================
// K1w: window trainer -- float32, time-major, groupings with overlapping windows (Grouper("time.dayofyear", 31):
// neighbouring groups share 30/31 of their 930 samples).  (included by xsdba_b200.cu inside its anonymous namespace)
//
// The per-group kernels (K1f / K1b) order the ~930 samples of every (gridpoint, group) from scratch: 365 times per
// gridpoint, 30 x more ordering work than the data has samples.  Here a CTA takes a CHUNK of up to 38
// consecutive groups of 8 gridpoints and orders the UNION of their window rows once (<= 2048 rows: 68 days of year x
// 30 years):
//   1. load the union rows (32-byte row pieces), monotone uint32 keys (NaN / missing -> 0xFFFFFFFF, above +inf);
//      every thread keeps its 32 keys in registers;
//   2. sort the keys of every column in shared memory (generic register-blocked bitonic sorter);
//   3. rank of every union row in its column's order: binary search of its key, equal keys take consecutive slots
//      through a shared-memory counter per run (any bijection inside a run of equal keys is as good as another);
//      stored as the inverse map  sorted position -> union row;
//   4. the window of group j as a BITMAP over the sorted positions: a host-built 64-bit mask per union row says which
//      groups of the chunk contain it, so word (column, j, b) is one warp ballot over the 32 positions of block b;
//      a prefix of the word popcounts per (column, group) turns "order statistic i" into a 6-step search for the
//      word plus a 5-step search for the bit, one thread per (column, group, node); value = sorted[position].
//      n_valid = set bits below the column's first NaN key.
// Semantics: identical to K1f (nbutils._nan_quantile_1d etc., nbutils.py:24-148; window gather base.py:261-265).
// Plain eqm_train / group quantiles only (no jitter, normalisation or frequency adaptation: those keep K1f / K1b).
// =============================================================================================
constexpr int kWinColBits = 2;
constexpr int kWinCols = 1 << kWinColBits;  // gridpoints per CTA: 4 -> 102 KB of shared memory, two CTAs per SM
constexpr int kWinThreads = 512;
constexpr int kWinMaxRows = 2048;   // union rows of a chunk
constexpr int kWinMaxGroups = 38;   // groups of a chunk (bits of the per-row group mask, words of the bitmap area)
constexpr int kWinWords = kWinMaxRows / 32;
constexpr int kWinPerThread = kWinMaxRows * kWinCols / kWinThreads;  // 32 keys per thread

struct WinSmem {
  static constexpr size_t keys = 0;                                              // unsigned [2048][8] sorted keys
  static constexpr size_t inv = keys + (size_t)kWinMaxRows * kWinCols * 4;        // uint16 [2048][8] union row at a sorted position
  static constexpr size_t gmask = inv + (size_t)kWinMaxRows * kWinCols * 2;       // uint64 [2048] groups that contain a union row
  static constexpr size_t pre = inv;                                              // uint16 [8][38][64] inclusive popcount prefix
                                                                                  //   (aliases inv + gmask once the bitmaps exist)
  static constexpr size_t scratch = gmask + (size_t)kWinMaxRows * 8;              // tie counters unsigned [2048][8] (64 KB), then
                                                                                  // bitmap words unsigned [8][38][64] (76 KB)
  static constexpr size_t scratch_bytes = (size_t)kWinCols * kWinMaxGroups * kWinWords * 4;
  static constexpr size_t nv = scratch + scratch_bytes;                           // int [8] keys below the NaN key per column
  static constexpr size_t nval = nv + 32;                                         // int [8][38] valid samples per (column, group)
  static constexpr size_t total = nval + (size_t)kWinCols * kWinMaxGroups * 4;
};
static_assert(WinSmem::scratch_bytes >= (size_t)kWinMaxRows * kWinCols * 4, "tie counters alias the bitmap area");
static_assert((size_t)kWinCols * kWinMaxGroups * kWinWords * 2 <= (size_t)kWinMaxRows * kWinCols * 2 + (size_t)kWinMaxRows * 8,
              "the popcount prefix aliases inv + gmask");

__device__ __forceinline__ unsigned win_key(float v) {   // monotone: -inf < ... < -0 < +0 < ... < +inf < NaN
  const unsigned b = __float_as_uint(v);
  return v != v ? 0xFFFFFFFFu : (b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u));
}
__device__ __forceinline__ float win_value(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// position of the (r + 1)-th set bit of m (0 <= r < popc(m)): five halving steps
__device__ __forceinline__ int nth_set_bit(unsigned m, int r) {
  int pos = 0;
  int t = __popc(m & 0xFFFFu);
  if (r >= t) { r -= t; m >>= 16; pos += 16; }
  t = __popc(m & 0xFFu);
  if (r >= t) { r -= t; m >>= 8; pos += 8; }
  t = __popc(m & 0xFu);
  if (r >= t) { r -= t; m >>= 4; pos += 4; }
  t = __popc(m & 0x3u);
  if (r >= t) { r -= t; m >>= 2; pos += 2; }
  t = (int)(m & 1u);
  if (r >= t) pos += 1;
  return pos;
}

// key at order statistic i of one (column, group): pre = inclusive popcount prefix over the 64 bitmap words.
// PAIR: also the key at i + 1 (which must exist) -- the next set bit, in the same word 13 times out of 14.
template <bool PAIR>
__device__ __forceinline__ unsigned win_select(const unsigned short* __restrict__ pre, const unsigned* __restrict__ words,
                                               const unsigned* __restrict__ keys_col, int i, unsigned& next_key) {
  int w = 0;
#pragma unroll
  for (int step = kWinWords / 2; step > 0; step >>= 1)
    if ((int)pre[w + step - 1] <= i) w += step;
  const int cb = w > 0 ? (int)pre[w - 1] : 0;
  const unsigned m = words[w];
  const int bit = nth_set_bit(m, i - cb);
  if (PAIR) {
    unsigned rest = bit == 31 ? 0u : (m & ~((2u << bit) - 1u));
    int w2 = w;
    while (rest == 0u) rest = words[++w2];
    next_key = keys_col[((w2 << 5) + __ffs(rest) - 1) * kWinCols];
  }
  return keys_col[((w << 5) + bit) * kWinCols];
}

__global__ void __launch_bounds__(kWinThreads, 2)
train_window_kernel(const float* __restrict__ ref, const float* __restrict__ hist_in, long long n_pts, long long st,
                    const int32_t* __restrict__ seg_off, const int32_t* __restrict__ chunk_g,
                    const int32_t* __restrict__ urow_off, const int32_t* __restrict__ urows,
                    const unsigned long long* __restrict__ gmask_g, int n_groups, const float* __restrict__ q, int nq,
                    int kind, int mode, float* __restrict__ af, float* __restrict__ hist_q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw + WinSmem::keys);
  unsigned short* inv = reinterpret_cast<unsigned short*>(smem_raw + WinSmem::inv);
  unsigned long long* gm = reinterpret_cast<unsigned long long*>(smem_raw + WinSmem::gmask);
  unsigned short* pre = reinterpret_cast<unsigned short*>(smem_raw + WinSmem::pre);
  unsigned* tie = reinterpret_cast<unsigned*>(smem_raw + WinSmem::scratch);
  unsigned* bmw = reinterpret_cast<unsigned*>(smem_raw + WinSmem::scratch);
  int* nvs = reinterpret_cast<int*>(smem_raw + WinSmem::nv);
  int* nval = reinterpret_cast<int*>(smem_raw + WinSmem::nval);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long n0 = (long long)blockIdx.x * kWinCols;
  const int ch = blockIdx.y;
  const int g0 = chunk_g[ch], K = chunk_g[ch + 1] - g0;
  const int32_t* __restrict__ rows = urows + urow_off[ch];
  const unsigned long long* __restrict__ gmask = gmask_g + urow_off[ch];
  const int U = urow_off[ch + 1] - urow_off[ch];
  int n_pad = 32;
  while (n_pad < U) n_pad <<= 1;
  const int n_words = n_pad >> 5;
  const long long out_stride = (long long)n_groups * nq;
  const int n_pass = mode == 0 ? 2 : 1;
  const int c_own = tid % kWinCols;            // the thread's keys: rows tid / 8 + 64 i of column tid % 8
  const bool c_own_ok = n0 + c_own < n_pts;
  const float fnan = Num<float>::nan();

  for (int pass = 0; pass < n_pass; ++pass) {
    const float* __restrict__ src = (pass == 0 ? ref : hist_in) + n0 + (c_own_ok ? c_own : 0);
    // ---- 1. union rows -> keys (registers + shared memory), group masks, counters --------------------------
    unsigned kreg[kWinPerThread];
    {
      int t[kWinPerThread];
#pragma unroll
      for (int i = 0; i < kWinPerThread; ++i) {
        const int r = tid / kWinCols + i * (kWinThreads / kWinCols);
        t[i] = r < U ? rows[r] : -1;
      }
#pragma unroll
      for (int i = 0; i < kWinPerThread; ++i) {
        float v = fnan;
        if (t[i] >= 0 && c_own_ok) v = src[(long long)t[i] * st];
        kreg[i] = win_key(v);
      }
#pragma unroll
      for (int i = 0; i < kWinPerThread; ++i) {
        const int idx = tid + i * kWinThreads;
        if (idx < n_pad * kWinCols) { keys[idx] = kreg[i]; tie[idx] = 0u; inv[idx] = 0xFFFFu; }
      }
      for (int r = tid; r < n_pad; r += kWinThreads) gm[r] = r < U ? gmask[r] : 0ull;
    }
    __syncthreads();
    // ---- 2. sort every column ----------------------------------------------------------------------------
    sort_columns<unsigned, kWinCols>(keys, n_pad);
    // ---- 3. ranks -> inverse map ---------------------------------------------------------------------------
    if (tid < kWinCols) {
      int lo = 0, hi = n_pad;  // keys below the NaN key: the valid samples of the union
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid * kWinCols + tid] < 0xFFFFFFFFu) lo = mid + 1; else hi = mid; }
      nvs[tid] = lo;
    }
    {
      const unsigned* col = keys + c_own;
#pragma unroll
      for (int i = 0; i < kWinPerThread; ++i) {   // (fully unrolled: kreg stays in registers)
        const int r = tid / kWinCols + i * (kWinThreads / kWinCols);
        if (r >= U) continue;
        const unsigned k = kreg[i];
        int lo = 0;
        for (int step = n_pad >> 1; step > 0; step >>= 1)     // first position with key >= k (n_pad is a power of two)
          if (col[(lo + step - 1) * kWinCols] < k) lo += step;
        int rank = lo;
        if (lo + 1 < n_pad && col[(lo + 1) * kWinCols] == k) rank += (int)atomicAdd(tie + lo * kWinCols + c_own, 1u);
        inv[rank * kWinCols + c_own] = (unsigned short)r;
      }
    }
    __syncthreads();
    // ---- 4a. bitmap words.  Lane l of a warp holds the group mask of sorted position 32 b + l; the word of group j
    //      over these 32 positions is column j of that 32 x 64 bit matrix: a butterfly transpose of the low halves
    //      (5 shuffle steps) hands lane j the word of group j, ballots do the few groups above 31 ----------------
    for (int task = warp; task < kWinCols * n_words; task += kWinThreads / 32) {
      const int c = task % kWinCols, b = task / kWinCols;
      const unsigned r = inv[((b << 5) + lane) * kWinCols + c];
      const unsigned long long m = r == 0xFFFFu ? 0ull : gm[r];
      unsigned x = (unsigned)m;
      const unsigned hi = (unsigned)(m >> 32);
#pragma unroll
      for (int k = 16; k >= 1; k >>= 1) {
        const unsigned m0 = k == 16 ? 0x0000FFFFu : k == 8 ? 0x00FF00FFu : k == 4 ? 0x0F0F0F0Fu : k == 2 ? 0x33333333u : 0x55555555u;
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, k);
        x = (lane & k) ? ((x & ~m0) | ((y & ~m0) >> k)) : ((x & m0) | ((y & m0) << k));
      }
      if (lane < K) bmw[(c * kWinMaxGroups + lane) * kWinWords + b] = x;
      for (int j = 32; j < K; ++j) {
        const unsigned word = __ballot_sync(0xffffffffu, (hi >> (j - 32)) & 1u);
        if (lane == 0) bmw[(c * kWinMaxGroups + j) * kWinWords + b] = word;
      }
    }
    __syncthreads();   // (inv and gm are dead from here on: pre aliases them)
    // ---- 4b. popcount prefix and valid count per (column, group) ---------------------------------------------
    if (tid < kWinCols * K) {
      const int c = tid % kWinCols, j = tid / kWinCols;
      const unsigned* words = bmw + (c * kWinMaxGroups + j) * kWinWords;
      unsigned short* p = pre + (c * kWinMaxGroups + j) * kWinWords;
      const int nv = nvs[c];
      int run = 0, n = 0;
      for (int w = 0; w < kWinWords; ++w) {
        const unsigned m = w < n_words ? words[w] : 0u;
        run += __popc(m);
        p[w] = (unsigned short)run;
        unsigned mv = m;
        if ((w << 5) + 32 > nv) mv = (w << 5) >= nv ? 0u : (m & ((1u << (nv & 31)) - 1u));
        n += __popc(mv);
      }
      nval[c * kWinMaxGroups + j] = n;
    }
    __syncthreads();
    // ---- 4c. one thread per (group, column, node); node fastest so that a thread's neighbours write neighbouring
    //      table entries.  (cj, k) advance by the block size without divisions. -----------------------------------
    {
      int k = tid % nq, cj = tid / nq;
      const int dk = kWinThreads % nq, dcj = kWinThreads / nq;
      const int n_cj = kWinCols * K;
      for (; cj < n_cj; k += dk, cj += dcj) {
        if (k >= nq) { k -= nq; ++cj; if (cj >= n_cj) break; }
        const int c = cj & (kWinCols - 1), j = cj >> kWinColBits;
        const int g = g0 + j;
        const int S = seg_off[g + 1] - seg_off[g];
        const int n = nval[c * kWinMaxGroups + j];
        const unsigned short* p = pre + (c * kWinMaxGroups + j) * kWinWords;
        const unsigned* words = bmw + (c * kWinMaxGroups + j) * kWinWords;
        const unsigned* kc = keys + c;
        float res = fnan;
        unsigned nk = 0u;
        if (n > 0 && S > 0) {
          const double vi = (double)(n - 1) * (double)q[k];     // nbutils.py:131
          float left, right, gamma;
          if (vi >= (double)(n - 1)) {     // nbutils.py:47-51: position -1 of the full-length sorted row
            left = right = (n < S) ? fnan : win_value(win_select<false>(p, words, kc, n - 1, nk));
            gamma = (float)(vi + 1.0);
          } else if (vi < 0.0) {
            left = right = win_value(win_select<false>(p, words, kc, 0, nk));
            gamma = (float)vi;
          } else {
            const int i = (int)vi;
            left = win_value(win_select<true>(p, words, kc, i, nk));   // i + 1 <= n - 1: the next set bit is a valid sample
            right = win_value(nk);
            gamma = (float)(vi - (double)i);                      // nbutils.py:142
          }
          const float diff = right - left;
          res = gamma >= 0.5f ? __fmaf_rn(-diff, 1.0f - gamma, right) : __fmaf_rn(diff, gamma, left);
          if (res != res) res = win_value(win_select<false>(p, words, kc, n - 1, nk));   // nbutils.py:146
        }
        if (n0 + c >= n_pts) continue;
        const long long o = (n0 + c) * out_stride + (long long)g * nq + k;
        if (mode == 1) {
          af[o] = res;
        } else if (pass == 0) {
          af[o] = res;                                            // ref_q parked in af until the hist pass
        } else {
          const float rq = af[o];
          hist_q[o] = res;
          af[o] = kind == XSDBA_KIND_ADD ? __fsub_rn(rq, res) : __fdiv_rn(rq, res);   // utils.py:130-143
        }
      }
    }
    __syncthreads();
  }
}
