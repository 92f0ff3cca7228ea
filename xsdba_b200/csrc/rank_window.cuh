// =============================================================================================
// K3w: window ranks (+ QDM factor lookup) -- float32, time-major, groupings with overlapping windows, rank_window = True
// (qdm_adjust with `group.apply(u.rank, sim, main_only=False, pct=True)`, _adjustment.py:872; utils.py:573-638).
// (included by xsdba_b200.cu inside its anonymous namespace, after K1w)
//
// Every member of group j (30 days of year d) is ranked inside the 930-sample window of its group; the per-group
// kernel K3 sorts that window from scratch, 365 times per gridpoint.  K3w re-uses K1w's structure: the UNION of the
// window rows of a chunk of up to 38 groups is ordered once per CTA (4 gridpoints), the window of group j is a bitmap
// over the sorted positions with a popcount prefix, and
//     #(window samples <  x) = set bits of bitmap j below the first position of x's run of equal keys,
//     #(window samples <= x) = set bits below the end of that run,
// two prefix look-ups per member instead of a sort per group.  The average-tie percentile rank, its (mn, mx)
// normalisation (utils.py:629-634) and the nearest-node factor lookup on the shared quantile axis follow K3 /
// lookup_2d_nearest_n to the letter (float64 arithmetic, left node on distance ties, NaN factors dropped,
// extrapolation first); -0.0 and +0.0 rank as equal.
// =============================================================================================
struct RankWinSmem {
  static constexpr size_t mnmx = WinSmem::total;                                   // double [2][4][38]
  static constexpr size_t vrange = mnmx + (size_t)2 * kWinCols * kWinMaxGroups * 8; // int [3][4][38]: first / last valid node, holes
  static constexpr size_t mpre = vrange + (size_t)3 * kWinCols * kWinMaxGroups * 4; // int [40] member-count prefix of the chunk
  static constexpr size_t qs = mpre + 40 * 4;                                      // float [kWinMaxNq] quantile axis
  static constexpr int kWinMaxNq = 128;
  static constexpr size_t total = qs + (size_t)kWinMaxNq * 4;
};

// set bits of one (column, group) bitmap below sorted position p (0 <= p <= 32 * kWinWords)
__device__ __forceinline__ int win_count_below(const unsigned short* __restrict__ pre, const unsigned* __restrict__ words, int p) {
  const int w = p >> 5, b = p & 31;
  if (w >= kWinWords) return (int)pre[kWinWords - 1];
  const int base = w > 0 ? (int)pre[w - 1] : 0;
  return base + (b ? __popc(words[w] & ((1u << b) - 1u)) : 0);
}
// sorted position of order statistic i of one (column, group)
__device__ __forceinline__ int win_position(const unsigned short* __restrict__ pre, const unsigned* __restrict__ words, int i) {
  int w = 0;
#pragma unroll
  for (int step = kWinWords / 2; step > 0; step >>= 1)
    if ((int)pre[w + step - 1] <= i) w += step;
  const int cb = w > 0 ? (int)pre[w - 1] : 0;
  return (w << 5) + nth_set_bit(words[w], i - cb);
}

__global__ void __launch_bounds__(kWinThreads, 2)
rank_window_kernel(const float* __restrict__ sim, long long n_pts, long long st, const int32_t* __restrict__ mem_off,
                   const int32_t* __restrict__ mem_rows, const int32_t* __restrict__ mem_u,
                   const int32_t* __restrict__ chunk_g, const int32_t* __restrict__ urow_off,
                   const int32_t* __restrict__ urows, const unsigned long long* __restrict__ gmask_g, int n_groups,
                   const float* __restrict__ af, const float* __restrict__ q, int nq, int extrap, int kind, int do_adjust,
                   float* __restrict__ scen, double* __restrict__ sim_q) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned* keys = reinterpret_cast<unsigned*>(smem_raw + WinSmem::keys);
  unsigned* rowpos = keys;  // alias (after the ranks): first | end << 16 of every union row's run of equal keys
  unsigned short* inv = reinterpret_cast<unsigned short*>(smem_raw + WinSmem::inv);
  unsigned long long* gm = reinterpret_cast<unsigned long long*>(smem_raw + WinSmem::gmask);
  unsigned short* pre = reinterpret_cast<unsigned short*>(smem_raw + WinSmem::pre);
  unsigned* tie = reinterpret_cast<unsigned*>(smem_raw + WinSmem::scratch);
  unsigned* bmw = reinterpret_cast<unsigned*>(smem_raw + WinSmem::scratch);
  int* nvs = reinterpret_cast<int*>(smem_raw + WinSmem::nv);
  int* nval = reinterpret_cast<int*>(smem_raw + WinSmem::nval);
  double* mns = reinterpret_cast<double*>(smem_raw + RankWinSmem::mnmx);
  double* mxs = mns + kWinCols * kWinMaxGroups;
  int* vfirst = reinterpret_cast<int*>(smem_raw + RankWinSmem::vrange);
  int* vlast = vfirst + kWinCols * kWinMaxGroups;
  int* vholes = vlast + kWinCols * kWinMaxGroups;
  int* mpre = reinterpret_cast<int*>(smem_raw + RankWinSmem::mpre);
  float* qsm = reinterpret_cast<float*>(smem_raw + RankWinSmem::qs);

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
  const int c_own = tid % kWinCols;            // the thread's keys: rows tid / 4 + 128 i of column tid % 4
  const bool c_own_ok = n0 + c_own < n_pts;
  const float fnan = Num<float>::nan();
  const double dnan = __longlong_as_double(0x7ff8000000000000LL);
  const long long pt_stride = (long long)n_groups * nq;

  // ---- 1. union rows -> keys (registers + shared memory), group masks, counters --------------------------------
  unsigned kreg[kWinPerThread];
  {
    const float* __restrict__ src = sim + n0 + (c_own_ok ? c_own : 0);
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
      kreg[i] = win_key(v + 0.0f);   // (-0.0 -> +0.0: equal values must share a key)
    }
#pragma unroll
    for (int i = 0; i < kWinPerThread; ++i) {
      const int idx = tid + i * kWinThreads;
      if (idx < n_pad * kWinCols) { keys[idx] = kreg[i]; tie[idx] = 0u; inv[idx] = 0xFFFFu; }
    }
    for (int r = tid; r < n_pad; r += kWinThreads) gm[r] = r < U ? gmask[r] : 0ull;
    if (tid <= K) mpre[tid] = mem_off[g0 + tid] - mem_off[g0];
    if (do_adjust) for (int k = tid; k < nq; k += kWinThreads) qsm[k] = q[k];
  }
  __syncthreads();
  // ---- 2. sort every column ----------------------------------------------------------------------------------
  sort_columns<unsigned, kWinCols>(keys, n_pad);
  // ---- 3. ranks -> inverse map; the first position of every row's run of equal keys stays in kreg ---------------
  if (tid < kWinCols) {
    int lo = 0, hi = n_pad;  // keys below the NaN key: the valid samples of the union
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid * kWinCols + tid] < 0xFFFFFFFFu) lo = mid + 1; else hi = mid; }
    nvs[tid] = lo;
  }
  {
    const unsigned* col = keys + c_own;
#pragma unroll
    for (int i = 0; i < kWinPerThread; ++i) {
      const int r = tid / kWinCols + i * (kWinThreads / kWinCols);
      if (r >= U) continue;
      const unsigned k = kreg[i];
      int lo = 0;
      for (int step = n_pad >> 1; step > 0; step >>= 1)     // first position with key >= k (n_pad is a power of two)
        if (col[(lo + step - 1) * kWinCols] < k) lo += step;
      int rank = lo;
      if (lo + 1 < n_pad && col[(lo + 1) * kWinCols] == k) rank += (int)atomicAdd(tie + lo * kWinCols + c_own, 1u);
      inv[rank * kWinCols + c_own] = (unsigned short)r;
      kreg[i] = (unsigned)lo;
    }
  }
  __syncthreads();
  // ---- 3b. end of the run: the tie counter of its first position holds the run length (0: a single key) ---------
#pragma unroll
  for (int i = 0; i < kWinPerThread; ++i) {
    const int r = tid / kWinCols + i * (kWinThreads / kWinCols);
    if (r >= U) continue;
    const unsigned lo = kreg[i];
    const unsigned len = max(tie[lo * kWinCols + c_own], 1u);
    kreg[i] = lo | ((lo + len) << 16);
  }
  __syncthreads();   // (the tie counters are dead: the bitmap words alias them)
  // ---- 4a. bitmap words (as K1w) -------------------------------------------------------------------------------
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
  // ---- 4b. popcount prefix and valid count per (column, group); valid nodes of the factor rows ------------------
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
  if (do_adjust) {   // one warp per (column, group): first / last node with a non-NaN factor, NaNs in between?
    for (int cj = warp; cj < kWinCols * K; cj += kWinThreads / 32) {
      const int c = cj % kWinCols, j = cj / kWinCols;
      int first = nq, last = -1, cnt = 0;
      if (n0 + c < n_pts) {
        const float* row = af + (n0 + c) * pt_stride + (long long)(g0 + j) * nq;
        float a[RankWinSmem::kWinMaxNq / 32];   // (all loads of the row first: four independent requests, not a chain)
#pragma unroll
        for (int b = 0; b < RankWinSmem::kWinMaxNq / 32; ++b) a[b] = b * 32 + lane < nq ? row[b * 32 + lane] : fnan;
#pragma unroll
        for (int b = 0; b < RankWinSmem::kWinMaxNq / 32; ++b) {
          const unsigned ok = __ballot_sync(0xffffffffu, a[b] == a[b]);
          if (ok) {
            if (first == nq) first = b * 32 + __ffs(ok) - 1;
            last = b * 32 + 31 - __clz(ok);
            cnt += __popc(ok);
          }
        }
      }
      if (lane == 0) {
        vfirst[c * kWinMaxGroups + j] = first;
        vlast[c * kWinMaxGroups + j] = last;
        vholes[c * kWinMaxGroups + j] = (last >= first && cnt != last - first + 1) ? 1 : 0;
      }
    }
  }
  __syncthreads();
  // ---- 4c. (mn, mx) of the percentile ranks per (column, group): average ranks of the window's smallest and largest
  //      value over n  (utils.py:629-634; K3) -------------------------------------------------------------------
  if (tid < kWinCols * K) {
    const int c = tid % kWinCols, j = tid / kWinCols;
    const unsigned short* p = pre + (c * kWinMaxGroups + j) * kWinWords;
    const unsigned* words = bmw + (c * kWinMaxGroups + j) * kWinWords;
    const unsigned* col = keys + c;
    const int n = nval[c * kWinMaxGroups + j];
    double mn = dnan, mx = dnan;
    if (n > 0) {
      const unsigned kmin = col[win_position(p, words, 0) * kWinCols];
      const unsigned kmax = col[win_position(p, words, n - 1) * kWinCols];
      int hi_min = 0, lo_max = 0;   // first position with key > kmin / with key >= kmax
      for (int step = n_pad >> 1; step > 0; step >>= 1) {
        if (col[(hi_min + step - 1) * kWinCols] <= kmin) hi_min += step;
        if (col[(lo_max + step - 1) * kWinCols] < kmax) lo_max += step;
      }
      if (hi_min < n_pad && col[hi_min * kWinCols] <= kmin) ++hi_min;   // (the search covers n_pad - 1 positions)
      if (lo_max < n_pad && col[lo_max * kWinCols] < kmax) ++lo_max;
      const int ub = win_count_below(p, words, hi_min);   // multiplicity of the minimum inside the window
      const int lb = win_count_below(p, words, lo_max);   // window samples below the maximum
      mn = ((double)(ub + 1) * 0.5) / (double)n;
      mx = ((double)(lb + n + 1) * 0.5) / (double)n;
    }
    mns[c * kWinMaxGroups + j] = mn;
    mxs[c * kWinMaxGroups + j] = mx;
  }
  __syncthreads();   // (the sorted keys are dead: rowpos aliases them)
  // ---- 5. run bounds of every union row, indexed by row ----------------------------------------------------------
#pragma unroll
  for (int i = 0; i < kWinPerThread; ++i) {
    const int r = tid / kWinCols + i * (kWinThreads / kWinCols);
    if (r < U) rowpos[r * kWinCols + c_own] = kreg[i];
  }
  __syncthreads();
  // ---- 6. one thread per (group, member, column), column fastest ------------------------------------------------
  const int n_items = mpre[K] * kWinCols;
  for (int item = tid; item < n_items; item += kWinThreads) {
    const int c = item & (kWinCols - 1), mi = item >> kWinColBits;
    int j = 0;
#pragma unroll
    for (int step = 32; step > 0; step >>= 1)
      if (j + step <= K && mpre[j + step] <= mi) j += step;
    const long long pt = n0 + c;
    if (pt >= n_pts) continue;
    const int g = g0 + j;
    const int m_glob = mem_off[g0] + mi;
    const long long o = (long long)mem_rows[m_glob] * st + pt;
    const float x = sim[o];
    const unsigned rp = rowpos[mem_u[m_glob] * kWinCols + c];
    const int n = nval[c * kWinMaxGroups + j];
    double sq = dnan;
    if (x == x && n > 0) {
      const unsigned short* p = pre + (c * kWinMaxGroups + j) * kWinWords;
      const unsigned* words = bmw + (c * kWinMaxGroups + j) * kWinWords;
      const int lb = win_count_below(p, words, (int)(rp & 0xffffu));
      const int ub = win_count_below(p, words, (int)(rp >> 16));
      const double mn = mns[c * kWinMaxGroups + j], mx = mxs[c * kWinMaxGroups + j];
      const double r = ((double)(lb + ub + 1) * 0.5) / (double)n;
      sq = __ddiv_rn(__dmul_rn(mx, __dsub_rn(r, mn)), __dsub_rn(mx, mn));
    }
    if (sim_q) sim_q[o] = sq;
    if (!do_adjust) continue;
    // nearest node of the shared quantile axis among the nodes with a factor (lookup_2d_nearest_n, x_shared)
    float f = fnan;
    const int kf = vfirst[c * kWinMaxGroups + j], kl = vlast[c * kWinMaxGroups + j];
    const float* row = af + pt * pt_stride + (long long)g * nq;
    if (sq != sq) {
      f = fnan;
    } else if (sq < (double)qsm[0]) {
      f = (extrap == 0 && kl >= kf) ? row[kf] : fnan;
    } else if (sq > (double)qsm[nq - 1]) {
      f = (extrap == 0 && kl >= kf) ? row[kl] : fnan;
    } else {
      int pos = 0, hi = nq;   // number of nodes < sq
      while (pos < hi) { const int mid = (pos + hi) >> 1; if ((double)qsm[mid] < sq) pos = mid + 1; else hi = mid; }
      int il = pos - 1, ir = pos;   // nearest nodes with a factor on either side
      if (il > kl) il = kl;
      if (ir < kf) ir = kf;
      if (vholes[c * kWinMaxGroups + j]) {
        while (il >= kf && row[il] != row[il]) --il;
        while (ir <= kl && row[ir] != row[ir]) ++ir;
      }
      double dbest = __longlong_as_double(0x7ff0000000000000LL);
      if (il >= kf && kl >= kf) { dbest = fabs(sq - (double)qsm[il]); f = row[il]; }
      if (ir <= kl && kl >= kf) {
        const double d = fabs((double)qsm[ir] - sq);
        if (d < dbest) { dbest = d; f = row[ir]; }
      }
      if (!(dbest < 1.0)) {   // no node with a factor in this row (or a degenerate axis): rows of the neighbouring groups
        Tables<float, kWinCols> tb;
        tb.nq = nq; tb.gx = q; tb.gy = af; tb.x_shared = true; tb.centre_only = true; tb.G = n_groups; tb.pt_stride = pt_stride;
        f = nearest_cross_rows<double, float, kWinCols>(tb, c, pt, g, sq, __dmul_rn(dbest, dbest), f);
      }
    }
    scen[o] = kind == XSDBA_KIND_ADD ? __fadd_rn(x, f) : __fmul_rn(x, f);
  }
}
