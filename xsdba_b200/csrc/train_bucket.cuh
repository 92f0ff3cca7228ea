// =============================================================================================
// K1b: bucket-select train / quantile kernel -- float32, time-major (stride_pt == 1), segments <= 1024 slots.
// (included by xsdba_b200.cu inside its anonymous namespace, after K1f)
//
// Same contract and results as K1f (train_fast_kernel), different algorithm.  The quantile step needs ~2*nq order
// statistics of each ~930-sample segment, not the sorted segment, and a full sorting network costs 21 compare-
// exchanges per sample.  Here every column (gridpoint = lane) is distribution-sorted instead:
//   1. load the segment into registers (32 slots per thread, the 128-byte rows of the time-major input); valid
//      count, column min / max (two-level reduction); NaN and missing slots become +inf keys, which sort behind
//      every valid value and are never read back (positions >= n);
//   2. bucket = ceil((v - min) * scale) in [0, 1022] with ONE round-up FFMA on the magic constant 2^23 (no F2I), +inf
//      keys clamp to bucket 1023.  The map is monotone; bucket 0 holds exactly the samples equal to the column
//      minimum (the product is exact inside the FMA and rounds up) -- the dry-day zeros of precipitation are a
//      constant-time case -- and bucket 1023 holds exactly the +inf keys;
//   3. histogram with shared-memory atomics on [word][lane] counters: a warp's 32 lanes hit 32 different banks,
//      measured 12-13 lane-updates per clock per SM (profiles/microbench_smem.cu) -- 2.5 clocks per warp.  Two 16-bit
//      counters share a word (buckets w and w + 512; a column has exactly 1024 keys), so 1024 buckets cost 64 KB;
//   4. exclusive prefix over the counters of every column (packed adds); a second atomic pass hands every key its
//      slot and scatters it: the column is now sorted by bucket, buckets hold 1-3 samples for continuous data;
//   5. order statistics i, i+1: the bucket of position i is a function of the VALUE stored there (no search); its
//      range comes from the prefix table; up to 8 samples are selected in registers through a 19-exchange
//      network; the few threads (~1e-3) that meet a bucket of 9..16 samples retry out of line with a 63-exchange
//      network on a 16-slot window.  Columns with heavier buckets (ties, multi-scale data) or an infinite range send
//      the tile to K1f's sorter right after the
//      histogram -- same results.
// Semantics: nbutils._nan_quantile_1d / _get_indexes / _linear_interpolation (nbutils.py:24-148), NaNs excluded,
// utils.get_correction (utils.py:130-143), dqm_train's normalisation (_adjustment.py:163-179) -- as K1f.
// =============================================================================================
constexpr int kBktN = 1024;       // buckets per column
constexpr int kBktW = kBktN / 2;  // counter words per column: bucket b lives in half b / 512 of word b % 512
constexpr int kBktAbort = 16;     // a bucket with more keys than this (seen in the histogram) sends the tile to the sorter
constexpr int kBktPitch = 33;     // row pitch of the [warp][column] partial tables (transposed reads are conflict free)

struct BktSmem {
  static constexpr size_t buf = 0;                                   // float    [1024 + 16][32] scattered / sorted column, 16 rows of +inf
  static constexpr size_t hist = buf + (1024 + 16) * 32 * 4;          // unsigned [512][32] packed counters (aliases: psum)
  static constexpr size_t part = hist + (size_t)kBktW * 32 * 4;      // [3][32][33]: pmin, pmax (float), pcnt (int); aliases: tot
  static constexpr size_t col = part + 3 * 32 * kBktPitch * 4;       // [8][32]: cmin, cmax (float), cnt (int), mu[2], scale (float), row stride copies
  static constexpr size_t rows = col + 8 * 32 * 4;                   // int [1024] member rows of the group (-1 past S)
  static constexpr size_t q = rows + 1024 * 4;                       // double [kFastMaxNq]
  static constexpr size_t pos = q + kFastMaxNq * 8;                  // int [kFastMaxNq], float [kFastMaxNq]: node positions of a full column
  static constexpr size_t refq = pos + kFastMaxNq * 8;               // float [nq][33]
  static __host__ __device__ constexpr size_t total(int nq) { return refq + (size_t)nq * 33 * 4; }
};

__device__ __forceinline__ float min3f(float a, float b, float c) {
  float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}

// Bucket key: 2^23 + bucket, as the bits of a float.  Monotone in v for scale >= 0; bucket 0 <=> v == cmin when
// scale > 0; finite v <= cmax land in [0, 1022]; +inf (and the NaN of inf * 0) -> 1023.
__device__ __forceinline__ unsigned bucket_key(float v, float cmin, float scale) {
  return __float_as_uint(fminf(__fmaf_ru(__fsub_rn(v, cmin), scale, 8388608.0f), 8389631.0f));
}
// end[b] (inclusive prefix after the scatter) of one column: half b / 512 of word b % 512, stride 32 words
__device__ __forceinline__ int bucket_end(const unsigned* __restrict__ endp, unsigned b) {
  const unsigned w = endp[(b & (kBktW - 1)) * 32];
  return (int)((b & kBktW) ? (w >> 16) : (w & 0xffffu));
}

// Bucket width of one column.  degenerate: infinite / NaN range -- the bucket map says nothing about such a column.
__device__ __forceinline__ float bucket_scale(float cmin, float cmax, bool& degenerate) {
  const float finf = __int_as_float(0x7f800000);
  const bool nonempty = !(cmin == finf && cmax == -finf);  // (min / max of no samples)
  const float range = __fsub_rn(cmax, cmin);
  const float sc = __fdiv_rn((float)kBktN - 2.5f, range);  // ceil((v - min) * scale) <= 1022 for finite v <= max
  const bool fin = range == range && fabsf(range) != finf;
  const bool ok = fin && range > 0.0f && fabsf(sc) != finf;
  degenerate = nonempty && !(ok || (fin && range == 0.0f));
  return ok ? sc : 0.0f;
}

// base + 128 * row as one IMAD (the C expression becomes shift + mask + add)
__device__ __forceinline__ uint32_t row128(uint32_t base, unsigned row) {
  uint32_t a;
  asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(a) : "r"(row), "r"(base));
  return a;
}

__device__ __forceinline__ void ce8(float& a, float& b) { const float lo = fminf(a, b); b = fmaxf(a, b); a = lo; }
__device__ __forceinline__ float pick8(const float (&x)[8], int r) {  // x[r], 0 <= r < 8: a 3-level select tree
  const bool b0 = r & 1, b1 = r & 2, b2 = r & 4;
  const float y0 = b0 ? x[1] : x[0], y1 = b0 ? x[3] : x[2], y2 = b0 ? x[5] : x[4], y3 = b0 ? x[7] : x[6];
  const float z0 = b1 ? y1 : y0, z1 = b1 ? y3 : y2;
  return b2 ? z1 : z0;
}

// Order statistics i and i + 1 of one bucket-sorted, non-degenerate column (0 <= i, i + 1 < n).  Returns false when a
// bucket holds more than 8 samples (the caller retries with bucket_select_pair16).  Straight-line code (selects, no
// branches), so that the two nodes a thread may own are interleaved by the instruction scheduler.
__device__ __forceinline__ bool bucket_select_pair(const unsigned* __restrict__ endp, const float* __restrict__ col,
                                                   float cmin, float scale, int i, float& left, float& right) {
  const float xi = col[i * 32], xj = col[(i + 1) * 32];
  const unsigned b0 = bucket_key(xi, cmin, scale) & (kBktN - 1), b1 = bucket_key(xj, cmin, scale) & (kBktN - 1);
  // buckets 0 (== column minimum) and 1023 (== +inf) hold one value each: nothing to select (the loads below then
  // look at bucket 1, in range and unused)
  const bool single = b0 == 0 || b0 == kBktN - 1;
  const unsigned bq = single ? 1u : b0;
  const int s0 = bucket_end(endp, bq - 1);
  const int m = bucket_end(endp, bq) - s0;
  const int r = single ? 0 : i - s0;
  // the 8 slots from s0 on: the m samples of the bucket, then samples of later buckets (all larger: the bucket map
  // is monotone) or the +inf rows behind the column -- the first m of the sorted window are the sorted bucket
  float x[8];
  const float* p = col + s0 * 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = p[j * 32];
  ce8(x[0], x[1]); ce8(x[2], x[3]); ce8(x[4], x[5]); ce8(x[6], x[7]);
  ce8(x[0], x[2]); ce8(x[1], x[3]); ce8(x[4], x[6]); ce8(x[5], x[7]);
  ce8(x[1], x[2]); ce8(x[5], x[6]); ce8(x[0], x[4]); ce8(x[3], x[7]);
  ce8(x[1], x[5]); ce8(x[2], x[6]);
  ce8(x[1], x[4]); ce8(x[3], x[6]);
  ce8(x[2], x[4]); ce8(x[3], x[5]);
  ce8(x[3], x[4]);
  const float l8 = pick8(x, r & 7), r8 = pick8(x, (r + 1) & 7);
  left = single ? xi : l8;
  right = single ? xi : r8;
  // position i + 1 in another bucket: that bucket's smallest sample (slots past the bucket hold larger samples: the
  // minimum of the 8-slot window is the minimum of the bucket).  Bucket 1023 holds +inf only.
  const bool next = b1 != b0, next_inf = b1 == kBktN - 1;
  const int m1 = bucket_end(endp, b1) - (i + 1);
  const float* pn = col + (i + 1) * 32;
  float y[8];
#pragma unroll
  for (int j = 1; j < 8; ++j) y[j] = pn[j * 32];
  float mn = min3f(xj, y[1], y[2]);
  mn = min3f(mn, y[3], y[4]); mn = min3f(mn, y[5], y[6]); mn = fminf(mn, y[7]);
  right = next ? (next_inf ? xj : mn) : right;
  return (single || m <= 8) && (!next || next_inf || m1 <= 8);
}

__device__ __forceinline__ float pick16(const float (&x)[16], int r) {  // x[r], 0 <= r < 16
  const bool b0 = r & 1, b1 = r & 2, b2 = r & 4, b3 = r & 8;
  float y[8], z[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) y[j] = b0 ? x[2 * j + 1] : x[2 * j];
#pragma unroll
  for (int j = 0; j < 4; ++j) z[j] = b1 ? y[2 * j + 1] : y[2 * j];
  const float w0 = b2 ? z[1] : z[0], w1 = b2 ? z[3] : z[2];
  return b3 ? w1 : w0;
}

// The same for buckets of up to 16 samples (Batcher's 63-exchange network on a 16-slot window) and a next bucket of
// any size: the out-of-line second try of the few threads (~1e-3) that meet such a bucket.  Returns false for
// m > 16: that node goes to the warp-cooperative path.
__device__ __noinline__ bool bucket_select_pair16(const unsigned* __restrict__ endp, const float* __restrict__ col,
                                                  float cmin, float scale, int i, float& left, float& right) {
  const float xi = col[i * 32], xj = col[(i + 1) * 32];
  const unsigned b0 = bucket_key(xi, cmin, scale) & (kBktN - 1), b1 = bucket_key(xj, cmin, scale) & (kBktN - 1);
  left = right = xi;
  if (b0 != 0 && b0 != kBktN - 1) {
    const int s0 = bucket_end(endp, b0 - 1);
    const int m = bucket_end(endp, b0) - s0, r = i - s0;
    if (m > 16) return false;
    float x[16];
    const float* p = col + s0 * 32;
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = p[j * 32];
    ce8(x[0], x[1]); ce8(x[2], x[3]); ce8(x[0], x[2]); ce8(x[1], x[3]); ce8(x[1], x[2]); ce8(x[4], x[5]);
    ce8(x[6], x[7]); ce8(x[4], x[6]); ce8(x[5], x[7]); ce8(x[5], x[6]); ce8(x[0], x[4]); ce8(x[2], x[6]);
    ce8(x[2], x[4]); ce8(x[1], x[5]); ce8(x[3], x[7]); ce8(x[3], x[5]); ce8(x[1], x[2]); ce8(x[3], x[4]);
    ce8(x[5], x[6]); ce8(x[8], x[9]); ce8(x[10], x[11]); ce8(x[8], x[10]); ce8(x[9], x[11]); ce8(x[9], x[10]);
    ce8(x[12], x[13]); ce8(x[14], x[15]); ce8(x[12], x[14]); ce8(x[13], x[15]); ce8(x[13], x[14]); ce8(x[8], x[12]);
    ce8(x[10], x[14]); ce8(x[10], x[12]); ce8(x[9], x[13]); ce8(x[11], x[15]); ce8(x[11], x[13]); ce8(x[9], x[10]);
    ce8(x[11], x[12]); ce8(x[13], x[14]); ce8(x[0], x[8]); ce8(x[4], x[12]); ce8(x[4], x[8]); ce8(x[2], x[10]);
    ce8(x[6], x[14]); ce8(x[6], x[10]); ce8(x[2], x[4]); ce8(x[6], x[8]); ce8(x[10], x[12]); ce8(x[1], x[9]);
    ce8(x[5], x[13]); ce8(x[5], x[9]); ce8(x[3], x[11]); ce8(x[7], x[15]); ce8(x[7], x[11]); ce8(x[3], x[5]);
    ce8(x[7], x[9]); ce8(x[11], x[13]); ce8(x[1], x[2]); ce8(x[3], x[4]); ce8(x[5], x[6]); ce8(x[7], x[8]);
    ce8(x[9], x[10]); ce8(x[11], x[12]); ce8(x[13], x[14]);
    left = pick16(x, r);
    right = pick16(x, (r + 1) & 15);
  }
  if (b1 != b0) {
    float mn = xj;
    if (b1 != kBktN - 1) {
      const int m1 = bucket_end(endp, b1) - (i + 1);
      const float* pn = col + (i + 1) * 32;
      for (int a = 1; a < m1; ++a) mn = fminf(mn, pn[a * 32]);
    }
    right = mn;
  }
  return true;
}

// nbutils._linear_interpolation (nbutils.py:101-104, 146) on the pair
__device__ __forceinline__ float bucket_interpolate(float left, float right, float gamma, float cmax) {
  const float diff = right - left;
  float r = gamma >= 0.5f ? __fmaf_rn(-diff, 1.0f - gamma, right) : __fmaf_rn(diff, gamma, left);
  if (r != r) r = cmax;  // nbutils.py:146
  return r;
}

// Position of quantile q in a column of n valid samples (nbutils.py:131, 47-56, 142): the index i of the left order
// statistic and the interpolation weight; i = -1: beyond the last sample, i = -2: before the first.
__device__ __forceinline__ void bucket_node_position(int n, double qk, int& i, float& gamma) {
  const double vi = (double)(n - 1) * qk;
  if (vi >= (double)(n - 1)) { i = -1; gamma = (float)(vi + 1.0); }
  else if (vi < 0.0) { i = -2; gamma = (float)vi; }
  else { i = (int)vi; gamma = (float)(vi - (double)i); }
}

// One quantile node of one column from the bucket-sorted column, buckets of up to 8 samples: straight-line code.
// heavy: a larger bucket was met (the result is then garbage and the caller retries with bucket_quantile_node<0>).
__device__ __forceinline__ float bucket_quantile_node_fast(const unsigned* __restrict__ endp, const float* __restrict__ colp,
                                                           int i, float gamma, int n, int S, float cmin, float cmax,
                                                           float scale, bool& heavy) {
  const bool edge_hi = i == -1, edge_lo = i == -2;  // (a non-edge node implies n >= 2 and i + 1 <= n - 1)
  float l, r;
  const bool ok = bucket_select_pair(endp, colp, cmin, scale, (edge_hi || edge_lo) ? 0 : i, l, r);
  const float hi_val = (n < S) ? Num<float>::nan() : cmax;  // nbutils.py:47-51: position -1 of the full-length sorted row
  const float left = edge_hi ? hi_val : edge_lo ? cmin : l;
  const float right = edge_hi ? hi_val : edge_lo ? cmin : r;
  heavy = !(ok || edge_hi || edge_lo) && n > 0;
  const float res = bucket_interpolate(left, right, gamma, cmax);
  return n > 0 ? res : Num<float>::nan();
}

// One quantile node of one column.  MODE 0: second try of a thread after bucket_quantile_node_fast (buckets of up to 16
// samples, the most the histogram check lets through); MODE 2: from the two sorted runs the sorter leaves.
template <int MODE>
__device__ __forceinline__ float bucket_quantile_node(const unsigned* __restrict__ endp, const float* __restrict__ colp,
                                                      int i, float gamma, int n, int S, float cmin, float cmax,
                                                      float scale, bool& heavy) {
  float left, right;
  if (i == -1) {  // nbutils.py:47-51: position -1 of the full-length sorted row
    left = right = (n < S) ? Num<float>::nan() : cmax;
  } else if (i == -2) {
    left = right = cmin;
  } else {
    left = right = cmin;
    if (MODE == 0) {
      heavy = !bucket_select_pair16(endp, colp, cmin, scale, i, left, right);
    } else {
      // both runs in full: the 1024 keys are the valid values plus +inf padding, and in a degenerate column
      // (infinite range: every key in one bucket) the slot order says nothing about which is which
      two_run_pair(colp, 512, colp + 512 * 32, 512, i, left, right);
    }
  }
  return bucket_interpolate(left, right, gamma, cmax);
}

// The sorter path of one pass (heavy buckets / degenerate columns): K1f's sorting network on the 1024 keys of every
// column in buf, quantiles from the two sorted runs.  One out-of-line copy: its register needs (32 keys per thread in
// flight) must not leak into the allocation of the bucket path.
struct BktOut {
  float* af; float* hist_q; float* refq; const double* qs;
  long long o_col; int n_items, nq, S, mode, pass, kind; bool col_ok;
};
__device__ __noinline__ void bucket_sorter_select(float* buf, const BktOut& o, int n, float cmin, float cmax) {
  sort_halves_512(buf, 0);
  const int lane = threadIdx.x & 31;
  const float* colp = buf + lane;
  bool heavy_unused = false;
#pragma unroll 1
  for (int item = threadIdx.x; item < o.n_items; item += kFastThreads) {
    const int k = item >> 5;
    float r = Num<float>::nan();
    if (n > 0) {
      int i; float gamma;
      bucket_node_position(n, o.qs[k], i, gamma);
      r = bucket_quantile_node<2>(nullptr, colp, i, gamma, n, o.S, cmin, cmax, 0.0f, heavy_unused);
    }
    if (o.mode == 1) {
      if (o.col_ok) o.af[o.o_col + k] = r;
    } else if (o.pass == 0) {
      o.refq[k * 33 + lane] = r;
    } else if (o.col_ok) {
      const float rq = o.refq[k * 33 + lane];
      o.hist_q[o.o_col + k] = r;
      o.af[o.o_col + k] = o.kind == XSDBA_KIND_ADD ? __fsub_rn(rq, r) : __fdiv_rn(rq, r);
    }
  }
  __syncthreads();
}

template <bool JITTER, bool NORM>
__global__ void __launch_bounds__(kFastThreads, 1)
train_bucket_kernel(const float* __restrict__ ref, const float* __restrict__ hist_in, long long n_pts, long long st,
                    const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows, int n_groups,
                    const float* __restrict__ q, int nq, int kind, int normalize_arg, int mode, float* __restrict__ af,
                    float* __restrict__ hist_q, float* __restrict__ scaling, JitterParams jp, int use_jitter,
                    const double* __restrict__ q64, int vec_enable, long long n_tiles) {
  const int normalize = NORM ? normalize_arg : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* buf = reinterpret_cast<float*>(smem_raw + BktSmem::buf);
  unsigned* hist = reinterpret_cast<unsigned*>(smem_raw + BktSmem::hist);
  double* psum = reinterpret_cast<double*>(smem_raw + BktSmem::hist);  // alias (NORM reduction, before the histogram)
  float* pmin = reinterpret_cast<float*>(smem_raw + BktSmem::part);
  float* pmax = pmin + 32 * kBktPitch;
  int* pcnt = reinterpret_cast<int*>(pmax + 32 * kBktPitch);
  unsigned* tot = reinterpret_cast<unsigned*>(pmin);                   // alias (prefix, after the min / max reduction)
  float* cminv = reinterpret_cast<float*>(smem_raw + BktSmem::col);
  float* cmaxv = cminv + 32;
  int* cnt = reinterpret_cast<int*>(cmaxv + 32);
  float* mu = reinterpret_cast<float*>(cnt + 32);                      // [2][32]
  float* scalev = cminv + 5 * 32;
  int* rows_tab = reinterpret_cast<int*>(smem_raw + BktSmem::rows);
  double* qs = reinterpret_cast<double*>(smem_raw + BktSmem::q);
  int* pos_i = reinterpret_cast<int*>(smem_raw + BktSmem::pos);        // node positions of a column without NaNs (n == S)
  float* pos_g = reinterpret_cast<float*>(pos_i + kFastMaxNq);
  float* refq = reinterpret_cast<float*>(smem_raw + BktSmem::refq);

  const long long out_stride = (long long)n_groups * nq;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float fnan = Num<float>::nan(), finf = Num<float>::inf();
  if (tid < 32) reinterpret_cast<int*>(cminv + 6 * 32)[tid] = (int)st * 4;
  if (tid < 16 * 32) buf[1024 * 32 + tid] = finf;  // the rows behind the column (the selection reads 8 / 16-slot windows)

  // Persistent: one CTA per SM walks the (group, tile) work items in launch order -- CTA b takes items b, b + grid, ...,
  // so the CTAs of a wave stream neighbouring tiles as before, without a CTA launch (198 KB of shared memory, 1024
  // threads) between two items, and the per-group tables are rebuilt only when the group changes.
  // (The normalising variant is launched with one CTA per item and compiled as a single trip: with the loop its
  // register allocation -- it also carries the float64 group means -- spills 180 bytes and runs 9 % slower.)
  constexpr bool kPersist = !NORM;
  const long long n_work = n_tiles * n_groups;
  int g_loaded = -1;
  long long w = blockIdx.x;
  do {
  const int g = (int)(w / n_tiles);
  const long long n0 = (w - (long long)g * n_tiles) * 32;
  const int S = seg_off[g + 1] - seg_off[g];

  if (S == 0) {  // group without members: NaN rows
    const int nqp = (nq + 31) & ~31;
    for (int item = tid; item < 32 * nqp; item += kFastThreads) {
      const int c = item / nqp, k = item % nqp;
      if (k >= nq || n0 + c >= n_pts) continue;
      const long long o = (n0 + c) * out_stride + (long long)g * nq + k;
      af[o] = fnan;
      if (mode == 0) hist_q[o] = fnan;
    }
    if (mode == 0 && scaling && tid < 32 && n0 + tid < n_pts) scaling[(n0 + tid) * n_groups + g] = fnan;
    continue;
  }
  if (!kPersist || g != g_loaded) {   // (CTA-uniform) the tables of the group: member rows, node positions of a full column
    __syncthreads();
    if (tid < nq) {
      const double qk = q64 ? q64[tid] : (double)q[tid];
      qs[tid] = qk;
      bucket_node_position(S, qk, pos_i[tid], pos_g[tid]);
    }
    rows_tab[tid] = tid < S ? seg_rows[seg_off[g] + tid] : -1;
    g_loaded = g;
    __syncthreads();
  }
  const bool col_ok = n0 + lane < n_pts;
  // whole tile inside the grid, rows 16-byte aligned: the vector load path
  const bool vec_ok = vec_enable && n0 + 32 <= n_pts && (st & 3) == 0 &&
                      ((reinterpret_cast<unsigned long long>(ref) | reinterpret_cast<unsigned long long>(hist_in)) & 15) == 0;
  const int n_pass = mode == 0 ? 2 : 1;
  const int n_items = nq * 32;
  // the row stride, read back from shared memory: a value ptxas cannot prove uniform stays in a vector register, and
  // row x stride + base is then ONE IMAD.WIDE per load instead of IMAD.WIDE (uniform operand) + a 64-bit add
  const int st4 = reinterpret_cast<volatile int*>(cminv + 6 * 32)[lane];  // (32 equal copies, one per lane)

  for (int pass = 0; pass < n_pass; ++pass) {
    // ---- load: warp w takes slots 32 w .. 32 w + 31 (one 128-byte row per instruction); columns past n_pts read
    //      column n0 (always valid memory).  Warps whose 32 slots all hold a row (warp-uniform) load without predicates ----
    float v[32];
    const int slot0 = warp * 32;
    const bool rows_ok = __all_sync(0xffffffffu, rows_tab[slot0 + lane] >= 0);  // (-1: past the segment / missing window slot)
    {
      const char* __restrict__ srcb = reinterpret_cast<const char*>((pass == 0 ? ref : hist_in) + n0 + (col_ok ? lane : 0));
      if (rows_ok && vec_ok) {
        // 16 bytes per lane: one instruction fetches four 128-byte rows (8 lanes each) straight into the warp's own rows of
        // buf (cp.async: no staging registers) -- a quarter of the load instructions, address computations and L1
        // requests of the 4-byte form; the column-wise read-back is 32 conflict-free LDS
        const char* __restrict__ src16 = reinterpret_cast<const char*>((pass == 0 ? ref : hist_in) + n0) + (lane & 7) * 16;
        const uint32_t dst16 = (uint32_t)__cvta_generic_to_shared(buf + (slot0 + (lane >> 3)) * 32 + (lane & 7) * 4);
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {
          const char* pa = row_address(src16, rows_tab[slot0 + 4 * qd + (lane >> 3)], st4);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst16 + qd * 4 * 128), "l"(pa) : "memory");
        }
        if (!normalize) {  // the counters are free (NORM sums in them first): clear them while the rows are in flight
#pragma unroll
          for (int j = 0; j < kBktW * 32 / kFastThreads; ++j) hist[tid + j * kFastThreads] = 0u;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const float* own = buf + slot0 * 32 + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = own[i * 32];
      } else if (rows_ok) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          asm("ld.global.nc.f32 %0, [%1];" : "=f"(v[i]) : "l"(row_address(srcb, rows_tab[slot0 + i], st4)));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int t = rows_tab[slot0 + i];
          const char* pa = row_address(srcb, t, st4);
          asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\tmov.f32 %0, %3;\n\t@p ld.global.nc.f32 %0, [%2];\n\t}"
              : "=f"(v[i]) : "r"(t), "l"(pa), "f"(fnan));
        }
      }
    }
    // ---- L2 prefetch of what this SM loads next, so that the next load phase (during which nothing else runs on
    //      this SM: one CTA per SM, all warps in the same phase) sees L2 latency and bandwidth, and the HBM traffic
    //      overlaps the histogram / scatter / selection phases.  Pass 0 prefetches this tile's hist rows; the last
    //      pass the ref rows of this CTA's next work item.  Lane i of warp w prefetches slot w + 32 i: one instruction per warp. -------------------------
    {
      const float* nxt = nullptr;
      int gn = g;
      if (mode == 0 && pass == 0) {
        nxt = hist_in + n0;
      } else if (w + gridDim.x < n_work) {   // this CTA's next work item
        const long long w2 = w + gridDim.x;
        gn = (int)(w2 / n_tiles);
        nxt = ref + (w2 - (long long)gn * n_tiles) * 32;
      }
      if (nxt) {
        const int slot = slot0 + lane;
        int t = -1;
        if (gn == g) t = rows_tab[slot];
        else if (slot < seg_off[gn + 1] - seg_off[gn]) t = seg_rows[seg_off[gn] + slot];
        if (t >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(row_address(reinterpret_cast<const char*>(nxt), t, st4)));
      }
    }
    float my_mn = finf, my_mx = -finf;
    if (JITTER && use_jitter && pass == 1) {  // hist only, per window slot (_adjustment.py:58-67); see K1f
      const long long seg_base = seg_off[g];
      float* own = buf + slot0 * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) own[i * 32] = v[i];
#pragma unroll 1
      for (int i = 0; i < 32; ++i)
        own[i * 32] = jitter_value<float>(own[i * 32], jp, (unsigned long long)((seg_base + slot0 + i) * n_pts + n0 + lane));
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = own[i * 32];
    }
    if (normalize) {  // dqm_train: x + (-mean) or x * (1/mean)   (_adjustment.py:167-168)
      double my_sum = 0.0;
      int c0 = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) if (v[i] == v[i]) { my_sum += (double)v[i]; ++c0; }
      pcnt[warp * 32 + lane] = c0;
      psum[warp * 32 + lane] = my_sum;
      __syncthreads();
      if (warp == 0) {
        double s = 0.0; int n = 0;
        for (int w = 0; w < 32; ++w) { s += psum[w * 32 + lane]; n += pcnt[w * 32 + lane]; }
        mu[pass * 32 + lane] = (float)(s / (double)n);
      }
      __syncthreads();
      const float m = mu[pass * 32 + lane];
      const float inv = kind == XSDBA_KIND_ADD ? -m : __fdiv_rn(1.0f, m);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float w = kind == XSDBA_KIND_ADD ? __fadd_rn(v[i], inv) : __fmul_rn(v[i], inv);
        v[i] = (v[i] == v[i] && w != w) ? finf : w;  // (a valid value stays a valid sort key, as in K1f)
      }
    }
    // ---- column min / max of the raw values (min / max skip NaN operands).  NaN values and missing slots need no
    //      treatment here: their bucket key is 1023 like +inf's (bucket_key), the scatter stores them as +inf, and the
    //      number of valid samples is 1024 minus the population of bucket 1023 (a real +inf makes the column
    //      degenerate, and that path counts for itself) ----
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      my_mx = max3f(my_mx, v[i], v[i + 1]);
      my_mn = min3f(my_mn, v[i], v[i + 1]);
    }
    pmin[warp * kBktPitch + lane] = my_mn;
    pmax[warp * kBktPitch + lane] = my_mx;
    if (normalize || !(rows_ok && vec_ok)) {
#pragma unroll
      for (int j = 0; j < kBktW * 32 / kFastThreads; ++j) hist[tid + j * kFastThreads] = 0u;
    }
    __syncthreads();
    {  // warp c reduces column c: lane r holds the partial of warp r (pitch 33: conflict free both ways)
      float mn = pmin[lane * kBktPitch + warp], mx = pmax[lane * kBktPitch + warp];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, d));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
      }
      if (lane == 0) { cminv[warp] = mn; cmaxv[warp] = mx; }
    }
    __syncthreads();
    const float cmin = cminv[lane], cmax = cmaxv[lane];
    bool degenerate;  // infinite / NaN range: the bucket map says nothing, the column goes through the sorter
    const float scale = bucket_scale(cmin, cmax, degenerate);
    if (warp == 0) scalev[lane] = scale;  // (read back by the scatter, behind a barrier)
    // ---- histogram: every slot has a key (valid value or +inf), no predicates --------------------------
    // (32-bit shared addresses: counter word = this lane's base + 128 bytes x (bucket mod 512) -- LOP3 + LEA)
    const uint32_t hist_lane = (uint32_t)__cvta_generic_to_shared(hist + lane);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const unsigned u = bucket_key(v[i], cmin, scale);
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(row128(hist_lane, u & (kBktW - 1))),
                   "r"((u & kBktW) ? 65536u : 1u) : "memory");
    }
    __syncthreads();
    int n = 1024 - (int)(hist[(kBktW - 1) * 32 + lane] >> 16);  // valid samples: all keys but those of bucket 1023
    if (warp == 0) cnt[lane] = n;
    // ---- exclusive prefix over the buckets of every column: thread (w, lane) owns words 16w .. 16w+15.  The two
    //      halves of a word are prefixed independently by the packed adds (no carry: totals <= 1024); the upper
    //      halves (buckets 512..1023) then start at the number of keys in buckets 0..511 ----------------------
    {
      // "some bucket holds more than kBktAbort keys" with packed arithmetic: count + (0x8000 - (kBktAbort + 1)) sets
      // bit 15 of its half (counts <= 1024: no carry between the halves); the two single-valued buckets (0: column
      // minimum, 1023: +inf keys) are masked out
      constexpr unsigned kBias = (0x8000u - (unsigned)(kBktAbort + 1)) * 0x00010001u;
      unsigned sum = 0, over = 0;
#pragma unroll
      for (int j = 0; j < kBktW / 32; ++j) {
        const unsigned w = hist[(warp * (kBktW / 32) + j) * 32 + lane];
        sum += w;
        unsigned wm = w;
        if (j == 0) wm = warp == 0 ? (w & 0xffff0000u) : w;
        if (j == kBktW / 32 - 1) wm = warp == 31 ? (wm & 0x0000ffffu) : wm;
        over |= (wm + kBias) & 0x80008000u;
      }
      const unsigned mxc = over ? (unsigned)kBktAbort + 1u : 0u;
      tot[warp * kBktPitch + lane] = sum;
      // Heavy buckets (ties away from the minimum, multi-scale data such as jittered precipitation) or a degenerate
      // column: selecting inside such buckets would cost more than sorting, so the whole tile goes to K1f's sorter
      // now -- keys stored in slot order, no prefix, no scatter.  (The barrier is the one the prefix needs anyway.)
      if (__syncthreads_or((mxc > (unsigned)kBktAbort || degenerate) ? 1 : 0)) {
        int my_cnt = 0;  // exact count here: a real +inf is a valid sample, its key is not told from a NaN's
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          my_cnt += v[i] == v[i] ? 1 : 0;
          buf[(slot0 + i) * 32 + lane] = fminf(v[i], finf);  // NaN -> +inf
        }
        pcnt[warp * kBktPitch + lane] = my_cnt;
        __syncthreads();
        n = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) n += pcnt[w * kBktPitch + lane];
        const BktOut bo{af, hist_q, refq, qs, (n0 + lane) * out_stride + (long long)g * nq, n_items, nq, S, mode, pass, kind,
                        col_ok};
        bucket_sorter_select(buf, bo, n, cmin, cmax);
        continue;
      }
      {  // warp c scans the 32 chunk totals of column c (lane r = chunk r) with packed adds; the upper halves start
         // at the number of keys in buckets 0..511, i.e. the lower half of the grand total
        const unsigned t = tot[lane * kBktPitch + warp];
        unsigned inc = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += y;
        }
        const unsigned all = __shfl_sync(0xffffffffu, inc, 31);
        tot[lane * kBktPitch + warp] = inc - t + (all << 16);
      }
      __syncthreads();
      unsigned run = tot[warp * kBktPitch + lane];
#pragma unroll
      for (int j = 0; j < kBktW / 32; ++j) {
        unsigned* p = hist + (warp * (kBktW / 32) + j) * 32 + lane;
        const unsigned c = *p;
        *p = run;
        run += c;
      }
    }
    __syncthreads();
    // ---- scatter: the atomic hands out the slot, start[b] becomes end[b].  (The keys are recomputed -- 3
    //      instructions -- from cmin / scale re-read from shared memory (volatile: opaque to nvvm AND ptxas): otherwise
    //      the compiler keeps the 64 addresses and increments of the histogram pass alive across the prefix, i.e.
    //      spills them to local memory.) --------
    {
      const uint32_t buf_lane = (uint32_t)__cvta_generic_to_shared(buf + lane);
      const float cmin_b = *reinterpret_cast<volatile float*>(cminv + lane);
      const float scale_b = *reinterpret_cast<volatile float*>(scalev + lane);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned u = bucket_key(v[i], cmin_b, scale_b);
        const bool hi = (u & kBktW) != 0;
        unsigned old;
        asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(row128(hist_lane, u & (kBktW - 1))),
                     "r"(hi ? 65536u : 1u) : "memory");
        const unsigned pos = hi ? (old >> 16) : (old & 0xffffu);
        // (NaN -> +inf: the selection windows may reach into bucket 1023)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(row128(buf_lane, pos)), "f"(fminf(v[i], finf)) : "memory");
      }
    }
    __syncthreads();
    // ---- quantiles: node k = item / 32 is warp-uniform, lane = column.  Results go straight to global memory
    //      (4-byte stores 2400 bytes apart; the 8 nodes of a 32-byte sector are written by 8 warps within the same
    //      round, the L2 write-back merges them) -- no staging buffer, no result registers across the barrier ----
    // (o = offset of node 0 of the column in the trained tables: hoisted out of the node loop)
    auto emit = [&](long long o_col0, bool okc, int k, int c, float r) {
      const long long o = o_col0 + k;
      if (mode == 1) {
        if (okc) af[o] = r;
      } else if (pass == 0) {
        refq[k * 33 + c] = r;
      } else if (okc) {
        const float rq = refq[k * 33 + c];
        hist_q[o] = r;
        af[o] = kind == XSDBA_KIND_ADD ? __fsub_rn(rq, r) : __fdiv_rn(rq, r);
      }
    };
    const long long o_own = (n0 + lane) * out_stride + (long long)g * nq;
    // a thread owns nodes tid / 32 and tid / 32 + 32 (nq = 50: two nodes for 18 warps, one for 14): both go through the
    // same straight-line code in one basic block, so the scheduler overlaps their shared-memory latencies
    auto node_retry = [&](int k, int i, float gamma) {  // a bucket of 9..16 samples (rare, divergent)
      bool heavy = false;   // (stays false: more than kBktAbort = 16 keys in a bucket sent the tile to the sorter)
      const float r = bucket_quantile_node<0>(hist + lane, buf + lane, i, gamma, n, S, cmin, cmax, scale, heavy);
      emit(o_own, col_ok, k, lane, r);
    };
#pragma unroll 1
    for (int item0 = tid; item0 < n_items; item0 += 2 * kFastThreads) {
      const int item1 = item0 + kFastThreads;
      const int k0 = item0 >> 5, k1 = item1 >> 5;
      int i0 = pos_i[k0];
      float g0 = pos_g[k0];
      if (n != S) bucket_node_position(n, qs[k0], i0, g0);  // a column with NaNs: its own positions
      if (item1 < n_items) {  // (warp-uniform)
        int i1 = pos_i[k1];
        float g1 = pos_g[k1];
        if (n != S) bucket_node_position(n, qs[k1], i1, g1);
        bool h0, h1;
        const float r0 = bucket_quantile_node_fast(hist + lane, buf + lane, i0, g0, n, S, cmin, cmax, scale, h0);
        const float r1 = bucket_quantile_node_fast(hist + lane, buf + lane, i1, g1, n, S, cmin, cmax, scale, h1);
        if (h0) node_retry(k0, i0, g0); else emit(o_own, col_ok, k0, lane, r0);
        if (h1) node_retry(k1, i1, g1); else emit(o_own, col_ok, k1, lane, r1);
      } else {
        bool h0;
        const float r0 = bucket_quantile_node_fast(hist + lane, buf + lane, i0, g0, n, S, cmin, cmax, scale, h0);
        if (h0) node_retry(k0, i0, g0); else emit(o_own, col_ok, k0, lane, r0);
      }
    }
    __syncthreads();  // refq complete; hist / buf / the partial tables are free for the next pass
  }
  if (normalize && mode == 0 && scaling && tid < 32 && n0 + tid < n_pts) {
    const float mr = mu[tid], mh = mu[32 + tid];  // scaling = get_correction(mu_hist, mu_ref)
    scaling[(n0 + tid) * n_groups + g] = kind == XSDBA_KIND_ADD ? __fsub_rn(mr, mh) : __fdiv_rn(mr, mh);
  }
  } while (kPersist && (w += gridDim.x) < n_work);  // work items
}
