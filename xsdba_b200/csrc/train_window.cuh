// =============================================================================================
// K1w: window trainer -- float32, time-major, groupings with overlapping windows (Grouper("time.dayofyear", 31):
// neighbouring groups share 30/31 of their 930 samples).  (included by xsdba_b200.cu inside its anonymous namespace)
//
// The per-group kernels (K1f / K1b) order the ~930 samples of every (gridpoint, group) from scratch: 365 times per
// gridpoint, 30 x more ordering work than the data has samples.  Here a CTA takes a CHUNK of up to 38
// consecutive groups of 8 gridpoints and orders the UNION of their window rows once (<= 2048 rows: 68 days of year x
// 30 years):
//   1. load the union rows (32-byte row pieces), monotone uint32 keys (NaN / missing -> 0xFFFFFFFF, above +inf);
//   2. sort the keys of every column in shared memory (generic register-blocked bitonic sorter);
//   3. rank of every union row in its column's order: binary search of its key, equal keys take consecutive slots
//      through a shared-memory counter per run (any bijection inside a run of equal keys is as good as another);
//   4. one thread per (column, group): the group's window is a BITMAP over the 2048 ranks (64 words, private to the
//      thread, word-major in shared memory so that a warp never has a bank conflict); the order statistics i, i+1 of
//      the window are the (i+1)-th / (i+2)-th set bits -- one popcount sweep over the 64 words serves all nq nodes,
//      because the nodes ascend -- and their values are sorted[position].  n_valid = set bits below the first NaN key.
// Cost per (gridpoint, group, array): ~930 x 6 (bitmap) + ~3.5 k (sweep) + the chunk's sort amortised over 38 groups
// (~4.6 k) instructions, against ~930 x 95 = 88 k for a sorting network per group.
// Semantics: identical to K1f (nbutils._nan_quantile_1d etc., nbutils.py:24-148; window gather base.py:261-265).
// Plain eqm_train / group quantiles only (no jitter, normalisation or frequency adaptation: those keep K1f / K1b).
// =============================================================================================
constexpr int kWinCols = 8;         // gridpoints per CTA
constexpr int kWinThreads = 512;
constexpr int kWinMaxRows = 2048;   // union rows of a chunk
constexpr int kWinMaxGroups = 38;   // groups of a chunk: kWinCols x 38 = 304 selection threads, 256 bytes of bitmap each
constexpr int kWinWords = kWinMaxRows / 32;
constexpr int kWinSel = 320;        // bitmap stride in words: selection threads rounded up to whole warps (bank = lane)

struct WinSmem {
  static constexpr size_t keys = 0;                                          // unsigned [2048][8] sorted keys
  static constexpr size_t ranks = keys + (size_t)kWinMaxRows * kWinCols * 4;  // uint16   [2048][8] rank of every union row
  static constexpr size_t scratch = ranks + (size_t)kWinMaxRows * kWinCols * 2;  // tie counters unsigned [2048][8] (64 KB),
                                                                                 // then bitmaps unsigned [64][320] (80 KB)
  static constexpr size_t scratch_bytes = (size_t)kWinWords * kWinSel * 4;
  static constexpr size_t nv = scratch + scratch_bytes;                       // int [8] keys below 0xFFFFFFFF per column
  static constexpr size_t total = nv + 64;
};
static_assert(WinSmem::scratch_bytes >= (size_t)kWinMaxRows * kWinCols * 4, "tie counters alias the bitmap area");
static_assert(kWinSel >= kWinCols * kWinMaxGroups && kWinSel % 32 == 0, "one bitmap column per selection thread");

__device__ __forceinline__ unsigned win_key(float v) {   // monotone: -inf < ... < -0 < +0 < ... < +inf < NaN
  const unsigned b = __float_as_uint(v);
  return v != v ? 0xFFFFFFFFu : (b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u));
}
__device__ __forceinline__ float win_value(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// position of the (r + 1)-th set bit of m (0 <= r < popc(m))
__device__ __forceinline__ int nth_set_bit(unsigned m, int r) {
  for (int t = 0; t < r; ++t) m &= m - 1;
  return __ffs(m) - 1;
}

__global__ void __launch_bounds__(kWinThreads, 1)
train_window_kernel(const float* __restrict__ ref, const float* __restrict__ hist_in, long long n_pts, long long st,
                    const int32_t* __restrict__ seg_off, const uint16_t* __restrict__ lseg,
                    const int32_t* __restrict__ chunk_g, const int32_t* __restrict__ urow_off,
                    const int32_t* __restrict__ urows, int n_groups, const float* __restrict__ q, int nq, int kind,
                    int mode, float* __restrict__ af, float* __restrict__ hist_q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw + WinSmem::keys);
  unsigned short* ranks = reinterpret_cast<unsigned short*>(smem_raw + WinSmem::ranks);
  unsigned* tie = reinterpret_cast<unsigned*>(smem_raw + WinSmem::scratch);
  unsigned* bm = reinterpret_cast<unsigned*>(smem_raw + WinSmem::scratch);
  int* nvs = reinterpret_cast<int*>(smem_raw + WinSmem::nv);

  const int tid = threadIdx.x;
  const long long n0 = (long long)blockIdx.x * kWinCols;
  const int ch = blockIdx.y;
  const int g0 = chunk_g[ch], K = chunk_g[ch + 1] - g0;
  const int32_t* __restrict__ rows = urows + urow_off[ch];
  const int U = urow_off[ch + 1] - urow_off[ch];
  int n_pad = 32;
  while (n_pad < U) n_pad <<= 1;
  const long long out_stride = (long long)n_groups * nq;
  const int n_sel = K * kWinCols;               // selection threads: (group j, column c) = (tid / 8, tid % 8)
  const int n_pass = mode == 0 ? 2 : 1;

  for (int pass = 0; pass < n_pass; ++pass) {
    const float* __restrict__ src = pass == 0 ? ref : hist_in;
    // ---- 1. union rows -> keys --------------------------------------------------------------------------
    for (int idx = tid; idx < n_pad * kWinCols; idx += kWinThreads) {
      const int r = idx / kWinCols, c = idx % kWinCols;
      unsigned k = 0xFFFFFFFFu;
      if (r < U && n0 + c < n_pts) k = win_key(src[n0 + c + (long long)rows[r] * st]);
      keys[idx] = k;
    }
    for (int idx = tid; idx < n_pad * kWinCols; idx += kWinThreads) tie[idx] = 0u;
    __syncthreads();
    // ---- 2. sort every column ------------------------------------------------------------------------------
    sort_columns<unsigned, kWinCols>(keys, n_pad);
    // ---- 3. ranks ------------------------------------------------------------------------------------------
    if (tid < kWinCols) {
      int lo = 0, hi = n_pad;  // keys below the NaN key: the valid samples of the union
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid * kWinCols + tid] < 0xFFFFFFFFu) lo = mid + 1; else hi = mid; }
      nvs[tid] = lo;
    }
    for (int idx = tid; idx < U * kWinCols; idx += kWinThreads) {
      const int r = idx / kWinCols, c = idx % kWinCols;
      unsigned k = 0xFFFFFFFFu;
      if (n0 + c < n_pts) k = win_key(src[n0 + c + (long long)rows[r] * st]);
      const unsigned* col = keys + c;
      int lo = 0, hi = n_pad;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid * kWinCols] < k) lo = mid + 1; else hi = mid; }
      int rank = lo;
      if (lo + 1 < n_pad && col[(lo + 1) * kWinCols] == k) rank += (int)atomicAdd(tie + lo * kWinCols + c, 1u);
      ranks[idx] = (unsigned short)rank;
    }
    __syncthreads();
    // ---- 4. one thread per (group, column): window bitmap, popcount sweep ------------------------------------
    if (tid < n_sel) {
      const int j = tid / kWinCols, c = tid % kWinCols;
      const int g = g0 + j;
      unsigned* my = bm + tid;                                  // word w of this thread at my[w * kWinSel]
      const int n_words = n_pad >> 5;
      for (int w = 0; w < n_words; ++w) my[w * kWinSel] = 0u;
      const int s0 = seg_off[g], S = seg_off[g + 1] - s0;
      for (int s_ = 0; s_ < S; ++s_) {
        const unsigned li = lseg[s0 + s_];
        if (li == 0xFFFFu) continue;
        const unsigned rk = ranks[li * kWinCols + c];
        my[(rk >> 5) * kWinSel] |= 1u << (rk & 31);
      }
      // valid samples of the window: set bits below the column's first NaN key
      const int nv = nvs[c];
      int n = 0;
      for (int w = 0; w < n_words; ++w) {
        unsigned m = my[w * kWinSel];
        if ((w << 5) + 32 > nv) m = (w << 5) >= nv ? 0u : (m & ((1u << (nv & 31)) - 1u));
        n += __popc(m);
      }
      // sweep state: `cb` set bits lie in the words before `w`
      int w = 0, cb = 0;
      auto select = [&](int i, int& pos_next) -> unsigned {     // key of order statistic i (0 <= i < n); position of i + 1
        if (i < cb) { w = 0; cb = 0; }
        unsigned m = my[w * kWinSel];
        while (cb + __popc(m) <= i) { cb += __popc(m); ++w; m = my[w * kWinSel]; }
        const int bit = nth_set_bit(m, i - cb);
        const int pos = (w << 5) + bit;
        unsigned rest = bit == 31 ? 0u : (m & ~((2u << bit) - 1u));
        int w2 = w;
        while (rest == 0u && w2 + 1 < n_words) { ++w2; rest = my[w2 * kWinSel]; }
        pos_next = rest ? (w2 << 5) + __ffs(rest) - 1 : -1;
        return keys[pos * kWinCols + c];
      };
      const bool col_ok = n0 + c < n_pts;
      const long long o_col = (n0 + c) * out_stride + (long long)g * nq;
      const float fnan = Num<float>::nan();
      float vmax = fnan;
      if (n > 0) { int pn; vmax = win_value(select(n - 1, pn)); }
      for (int k = 0; k < nq; ++k) {
        float res = fnan;
        if (n > 0 && S > 0) {
          const double vi = (double)(n - 1) * (double)q[k];   // nbutils.py:131
          float left, right, gamma;
          if (vi >= (double)(n - 1)) {   // nbutils.py:47-51: position -1 of the full-length sorted row
            left = right = (n < S) ? fnan : vmax;
            gamma = (float)(vi + 1.0);
          } else if (vi < 0.0) {
            int pn; left = right = win_value(select(0, pn));
            gamma = (float)vi;
          } else {
            const int i = (int)vi;
            int pn;
            left = win_value(select(i, pn));
            right = win_value(keys[pn * kWinCols + c]);      // i + 1 < n: the next set bit exists and is valid
            gamma = (float)(vi - (double)i);                  // nbutils.py:142
          }
          const float diff = right - left;
          res = gamma >= 0.5f ? __fmaf_rn(-diff, 1.0f - gamma, right) : __fmaf_rn(diff, gamma, left);
          if (res != res) res = vmax;                         // nbutils.py:146
        }
        if (!col_ok) continue;
        if (mode == 1) {
          af[o_col + k] = res;
        } else if (pass == 0) {
          af[o_col + k] = res;                                // ref_q parked in af until the hist pass
        } else {
          const float rq = af[o_col + k];
          hist_q[o_col + k] = res;
          af[o_col + k] = kind == XSDBA_KIND_ADD ? __fsub_rn(rq, res) : __fdiv_rn(rq, res);   // utils.py:130-143
        }
      }
    }
    __syncthreads();
  }
}
