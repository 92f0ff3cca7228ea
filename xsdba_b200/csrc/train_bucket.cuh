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
//      network, up to 64 by counting, otherwise (heavy ties, multi-scale data, infinite ranges) the CTA runs K1f's
//      sorter on the scattered column -- same results.
// Semantics: nbutils._nan_quantile_1d / _get_indexes / _linear_interpolation (nbutils.py:24-148), NaNs excluded,
// utils.get_correction (utils.py:130-143), dqm_train's normalisation (_adjustment.py:163-179) -- as K1f.
// =============================================================================================
constexpr int kBktN = 1024;       // buckets per column
constexpr int kBktW = kBktN / 2;  // counter words per column: bucket b lives in half b / 512 of word b % 512
constexpr int kBktLoopMax = 64;   // largest bucket selected by counting; larger ones -> sorter fallback
constexpr int kBktAbort = 24;     // a bucket with more keys than this (seen in the histogram) sends the tile to the sorter

struct BktSmem {
  static constexpr size_t buf = 0;                                   // float    [1024][32] scattered / sorted column
  static constexpr size_t hist = buf + 1024 * 32 * 4;                // unsigned [512][32] packed counters (aliases: psum)
  static constexpr size_t part = hist + (size_t)kBktW * 32 * 4;      // [3][32][32]: pmin, pmax (float), pcnt (int); tot aliases pmin
  static constexpr size_t col = part + 3 * 1024 * 4;                 // [8][32]: cmin, cmax (float), cnt (int), mu[2] (float), nlow
  static constexpr size_t rows = col + 8 * 32 * 4;                   // int [1024] member rows of the group (-1 past S)
  static constexpr size_t q = rows + 1024 * 4;                       // double [kFastMaxNq]
  static constexpr size_t refq = q + kFastMaxNq * 8;                 // float [nq][33]
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

__device__ __forceinline__ void ce8(float& a, float& b) { const float lo = fminf(a, b); b = fmaxf(a, b); a = lo; }
__device__ __forceinline__ float pick8(const float (&x)[8], int r) {  // x[r], 0 <= r < 8: a 3-level select tree
  const bool b0 = r & 1, b1 = r & 2, b2 = r & 4;
  const float y0 = b0 ? x[1] : x[0], y1 = b0 ? x[3] : x[2], y2 = b0 ? x[5] : x[4], y3 = b0 ? x[7] : x[6];
  const float z0 = b1 ? y1 : y0, z1 = b1 ? y3 : y2;
  return b2 ? z1 : z0;
}

// slow path of bucket_select_pair: a bucket with 8 < m <= kBktLoopMax samples, r-th and (r+1)-th by counting
__device__ __noinline__ void bucket_count_select(const float* __restrict__ col, int s, int e, int r, float& left,
                                                 float& right) {
  for (int a = s; a < e; ++a) {
    const float xa = col[a * 32];
    int c = 0;
    for (int j = s; j < e; ++j) {
      const float y = col[j * 32];
      c += (y < xa || (y == xa && j < a)) ? 1 : 0;
    }
    if (c == r) left = xa;
    if (c == r + 1) right = xa;
  }
}

// Order statistics i and i + 1 of one bucket-sorted, non-degenerate column (0 <= i, i + 1 < n).  Returns false when a
// bucket is too large for the in-place selection (the caller falls back to the sorter).
__device__ __forceinline__ bool bucket_select_pair(const unsigned* __restrict__ endp, const float* __restrict__ col,
                                                   float cmin, float scale, int i, float& left, float& right) {
  const float inf = __int_as_float(0x7f800000);
  const float xi = col[i * 32], xj = col[(i + 1) * 32];
  const unsigned b0 = bucket_key(xi, cmin, scale) & (kBktN - 1), b1 = bucket_key(xj, cmin, scale) & (kBktN - 1);
  bool ok = true;
  left = right = xi;  // buckets 0 (== column minimum) and 1023 (== +inf) hold one value each
  if (b0 != 0 && b0 != kBktN - 1) {
    const int e0 = bucket_end(endp, b0);
    const int s0 = bucket_end(endp, b0 - 1);
    const int m = e0 - s0, r = i - s0;
    if (m <= 8) {
      float x[8];
      const float* p = col + s0 * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = j < m ? p[j * 32] : inf;
      ce8(x[0], x[1]); ce8(x[2], x[3]); ce8(x[4], x[5]); ce8(x[6], x[7]);
      ce8(x[0], x[2]); ce8(x[1], x[3]); ce8(x[4], x[6]); ce8(x[5], x[7]);
      ce8(x[1], x[2]); ce8(x[5], x[6]); ce8(x[0], x[4]); ce8(x[3], x[7]);
      ce8(x[1], x[5]); ce8(x[2], x[6]);
      ce8(x[1], x[4]); ce8(x[3], x[6]);
      ce8(x[2], x[4]); ce8(x[3], x[5]);
      ce8(x[3], x[4]);
      left = pick8(x, r);
      right = pick8(x, (r + 1) & 7);
    } else if (m <= kBktLoopMax) {
      bucket_count_select(col, s0, e0, r, left, right);
    } else {
      ok = false;
    }
  }
  if (b1 != b0) {  // position i + 1 opens the next non-empty bucket: its smallest sample
    float mn = xj;
    if (b1 != kBktN - 1) {
      const int m1 = bucket_end(endp, b1) - (i + 1);
      const float* p = col + (i + 1) * 32;
      if (m1 <= 8) {
        float y[8];
#pragma unroll
        for (int j = 1; j < 8; ++j) y[j] = j < m1 ? p[j * 32] : inf;
        mn = min3f(mn, y[1], y[2]); mn = min3f(mn, y[3], y[4]); mn = min3f(mn, y[5], y[6]); mn = fminf(mn, y[7]);
      } else {
        for (int a = 1; a < m1; ++a) mn = fminf(mn, p[a * 32]);
      }
    }
    right = mn;
  }
  return ok;
}

// One quantile node of one column from the bucket-sorted column (fast = true) or from the two sorted runs the
// sorter leaves (fast = false).  Returns the node value; sets fb when the bucket path has to give up.
template <bool FAST>
__device__ __forceinline__ float bucket_quantile_node(const unsigned* __restrict__ endp, const float* __restrict__ colp,
                                                      double qk, int n, int S, float cmin, float cmax, float scale,
                                                      bool degenerate, int& fb) {
  const double vi = (double)(n - 1) * qk;  // nbutils.py:131
  float left, right, gamma;
  if (vi >= (double)(n - 1)) {  // nbutils.py:47-51: position -1 of the full-length sorted row
    left = right = (n < S) ? Num<float>::nan() : cmax;
    gamma = (float)(vi + 1.0);
  } else if (vi < 0.0) {
    left = right = cmin;
    gamma = (float)vi;
  } else {
    const int i = (int)vi;
    left = right = cmin;
    if (FAST) {
      if (degenerate || !bucket_select_pair(endp, colp, cmin, scale, i, left, right)) fb = 1;
    } else {
      // both runs in full: the 1024 keys are the valid values plus +inf padding, and in a degenerate column
      // (infinite range: every key in one bucket) the scatter order says nothing about which is which
      two_run_pair(colp, 512, colp + 512 * 32, 512, i, left, right);
    }
    gamma = (float)(vi - (double)i);  // nbutils.py:142
  }
  const float diff = right - left;
  float r = gamma >= 0.5f ? __fmaf_rn(-diff, 1.0f - gamma, right) : __fmaf_rn(diff, gamma, left);
  if (r != r) r = cmax;  // nbutils.py:146
  return r;
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
  int fb_unused = 0;
#pragma unroll 1
  for (int item = threadIdx.x; item < o.n_items; item += kFastThreads) {
    const int k = item >> 5;
    float r = Num<float>::nan();
    if (n > 0) r = bucket_quantile_node<false>(nullptr, colp, o.qs[k], n, o.S, cmin, cmax, 0.0f, false, fb_unused);
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
                    const double* __restrict__ q64) {
  const int normalize = NORM ? normalize_arg : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* buf = reinterpret_cast<float*>(smem_raw + BktSmem::buf);
  unsigned* hist = reinterpret_cast<unsigned*>(smem_raw + BktSmem::hist);
  double* psum = reinterpret_cast<double*>(smem_raw + BktSmem::hist);  // alias (NORM reduction, before the histogram)
  float* pmin = reinterpret_cast<float*>(smem_raw + BktSmem::part);
  float* pmax = pmin + 1024;
  int* pcnt = reinterpret_cast<int*>(pmax + 1024);
  unsigned* tot = reinterpret_cast<unsigned*>(pmin);                   // alias (prefix, after the min / max reduction)
  float* cminv = reinterpret_cast<float*>(smem_raw + BktSmem::col);
  float* cmaxv = cminv + 32;
  int* cnt = reinterpret_cast<int*>(cmaxv + 32);
  float* mu = reinterpret_cast<float*>(cnt + 32);                      // [2][32]
  int* rows_tab = reinterpret_cast<int*>(smem_raw + BktSmem::rows);
  double* qs = reinterpret_cast<double*>(smem_raw + BktSmem::q);
  float* refq = reinterpret_cast<float*>(smem_raw + BktSmem::refq);

  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * 32;
  const int S = seg_off[g + 1] - seg_off[g];
  const long long out_stride = (long long)n_groups * nq;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float fnan = Num<float>::nan(), finf = Num<float>::inf();

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
    return;
  }
  if (tid < nq) qs[tid] = q64 ? q64[tid] : (double)q[tid];
  rows_tab[tid] = tid < S ? seg_rows[seg_off[g] + tid] : -1;
  __syncthreads();
  const bool col_ok = n0 + lane < n_pts;
  const int n_pass = mode == 0 ? 2 : 1;
  const int n_items = nq * 32;
  const int st4 = (int)st * 4;
  const long long n_tiles = gridDim.x;

  for (int pass = 0; pass < n_pass; ++pass) {
    // ---- load: warp w takes slots w, w + 32, ...; columns past n_pts read column n0 (always valid memory) ----
    float v[32];
    {
      const char* __restrict__ srcb = reinterpret_cast<const char*>((pass == 0 ? ref : hist_in) + n0 + (col_ok ? lane : 0));
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int t = rows_tab[warp + 32 * i];
        const char* pa = row_address(srcb, t, st4);
        asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\tmov.f32 %0, %3;\n\t@p ld.global.nc.f32 %0, [%2];\n\t}"
            : "=f"(v[i]) : "r"(t), "l"(pa), "f"(fnan));
      }
    }
    // ---- L2 prefetch of what this SM loads next, so that the next load phase (during which nothing else runs on
    //      this SM: one CTA per SM, all warps in the same phase) sees L2 latency and bandwidth, and the HBM traffic
    //      overlaps the histogram / scatter / selection phases.  Pass 0 prefetches this tile's hist rows; the last
    //      pass the ref rows of the tile 148 blocks ahead in launch order (the one an SM of this wave picks up
    //      next).  Lane i of warp w prefetches slot w + 32 i: one instruction per warp. -------------------------
    {
      const float* nxt = nullptr;
      int gn = g;
      if (mode == 0 && pass == 0) {
        nxt = hist_in + n0;
      } else if (n_tiles >= 148) {
        long long x2 = (long long)blockIdx.x + 148;
        if (x2 >= n_tiles) { x2 -= n_tiles; ++gn; }
        if (gn < n_groups) nxt = ref + x2 * 32;
      }
      if (nxt) {
        const int slot = warp + 32 * lane;
        int t = -1;
        if (gn == g) t = rows_tab[slot];
        else if (slot < seg_off[gn + 1] - seg_off[gn]) t = seg_rows[seg_off[gn] + slot];
        if (t >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(row_address(reinterpret_cast<const char*>(nxt), t, st4)));
      }
    }
    float my_mn = finf, my_mx = -finf;
    int my_cnt = 0;
    if (JITTER && use_jitter && pass == 1) {  // hist only, per window slot (_adjustment.py:58-67); see K1f
      const long long seg_base = seg_off[g];
      float* own = buf + warp * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) own[(size_t)i * 1024] = v[i];
#pragma unroll 1
      for (int i = 0; i < 32; ++i)
        own[(size_t)i * 1024] = jitter_value<float>(
            own[(size_t)i * 1024], jp, (unsigned long long)((seg_base + warp + 32 * i) * n_pts + n0 + lane));
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = own[(size_t)i * 1024];
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
    // ---- column max / count on the raw values (max skips NaN operands), NaN -> +inf keys, column min ----
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      my_mx = max3f(my_mx, v[i], v[i + 1]);
      asm("{\n\t.reg .pred p;\n\tsetp.num.f32 p, %1, %1;\n\t@p add.s32 %0, %0, 1;\n\t@!p mov.f32 %1, %2;\n\t}"
          : "+r"(my_cnt), "+f"(v[i]) : "f"(finf));
      asm("{\n\t.reg .pred p;\n\tsetp.num.f32 p, %1, %1;\n\t@p add.s32 %0, %0, 1;\n\t@!p mov.f32 %1, %2;\n\t}"
          : "+r"(my_cnt), "+f"(v[i + 1]) : "f"(finf));
      my_mn = min3f(my_mn, v[i], v[i + 1]);
    }
    pmin[warp * 32 + lane] = my_mn;
    pmax[warp * 32 + lane] = my_mx;
    pcnt[warp * 32 + lane] = my_cnt;
#pragma unroll
    for (int j = 0; j < kBktW * 32 / kFastThreads; ++j) hist[tid + j * kFastThreads] = 0u;
    __syncthreads();
    if (warp == 0) {
      float mn = finf;
#pragma unroll 8
      for (int w = 0; w < 32; ++w) mn = fminf(mn, pmin[w * 32 + lane]);
      cminv[lane] = mn;
    } else if (warp == 1) {
      float mx = -finf;
#pragma unroll 8
      for (int w = 0; w < 32; ++w) mx = fmaxf(mx, pmax[w * 32 + lane]);
      cmaxv[lane] = mx;
    } else if (warp == 2) {
      int n = 0;
#pragma unroll 8
      for (int w = 0; w < 32; ++w) n += pcnt[w * 32 + lane];
      cnt[lane] = n;
    }
    __syncthreads();
    const float cmin = cminv[lane], cmax = cmaxv[lane];
    const int n = cnt[lane];
    float scale;
    bool degenerate;  // infinite / NaN range: the bucket map says nothing, the column goes through the sorter
    {
      const float range = __fsub_rn(cmax, cmin);
      const float sc = __fdiv_rn((float)kBktN - 2.5f, range);  // ceil((v - min) * scale) <= 1022 for finite v <= max
      const bool fin = range == range && fabsf(range) != finf;
      const bool ok = fin && range > 0.0f && fabsf(sc) != finf;
      scale = ok ? sc : 0.0f;
      degenerate = n > 0 && !(ok || (fin && range == 0.0f));
    }
    // ---- histogram: every slot has a key (valid value or +inf), no predicates --------------------------
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const unsigned u = bucket_key(v[i], cmin, scale);
      atomicAdd(hist + (u & (kBktW - 1)) * 32 + lane, (u & kBktW) ? 65536u : 1u);
    }
    __syncthreads();
    // ---- exclusive prefix over the buckets of every column: thread (w, lane) owns words 16w .. 16w+15.  The two
    //      halves of a word are prefixed independently by the packed adds (no carry: totals <= 1024); the upper
    //      halves (buckets 512..1023) then start at the number of keys in buckets 0..511 ----------------------
    {
      unsigned sum = 0, mxc = 0;
#pragma unroll
      for (int j = 0; j < kBktW / 32; ++j) {
        const unsigned w = hist[(warp * (kBktW / 32) + j) * 32 + lane];
        sum += w;
        // largest bucket of the column, the two single-valued buckets (0: column minimum, 1023: +inf keys) aside
        const unsigned lo = (warp == 0 && j == 0) ? 0u : (w & 0xffffu);
        const unsigned hi = (warp == 31 && j == kBktW / 32 - 1) ? 0u : (w >> 16);
        mxc = max(mxc, max(lo, hi));
      }
      tot[warp * 32 + lane] = sum;
      // Heavy buckets (ties away from the minimum, multi-scale data such as jittered precipitation) or a degenerate
      // column: selecting inside such buckets would cost more than sorting, so the whole tile goes to K1f's sorter
      // now -- keys stored in slot order, no prefix, no scatter.  (The barrier is the one the prefix needs anyway.)
      if (__syncthreads_or((mxc > (unsigned)kBktAbort || degenerate) ? 1 : 0)) {
#pragma unroll
        for (int i = 0; i < 32; ++i) buf[(warp + 32 * i) * 32 + lane] = v[i];
        __syncthreads();
        const BktOut bo{af, hist_q, refq, qs, (n0 + lane) * out_stride + (long long)g * nq, n_items, nq, S, mode, pass, kind,
                        col_ok};
        bucket_sorter_select(buf, bo, n, cmin, cmax);
        continue;
      }
      if (warp == 0) {  // one warp scans the 32 chunk totals of every column
        unsigned all = 0;  // (two sweeps over shared memory: 32 more registers next to v[] would spill)
#pragma unroll 8
        for (int w = 0; w < 32; ++w) all += tot[w * 32 + lane];
        unsigned run = all << 16;
#pragma unroll 4
        for (int w = 0; w < 32; ++w) { const unsigned t = tot[w * 32 + lane]; tot[w * 32 + lane] = run; run += t; }
      }
      __syncthreads();
      unsigned run = tot[warp * 32 + lane];
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
    //      instructions -- from opaque copies of cmin / scale: otherwise the compiler keeps the 64 addresses and
    //      increments of the histogram pass alive across the prefix, i.e. spills them to local memory.) --------
    {
      float cmin_b = cmin, scale_b = scale;
      asm volatile("" : "+f"(cmin_b), "+f"(scale_b));
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const unsigned u = bucket_key(v[i], cmin_b, scale_b);
        const bool hi = (u & kBktW) != 0;
        const unsigned old = atomicAdd(hist + (u & (kBktW - 1)) * 32 + lane, hi ? 65536u : 1u);
        const unsigned pos = hi ? (old >> 16) : (old & 0xffffu);
        buf[pos * 32 + lane] = v[i];
      }
    }
    __syncthreads();
    // ---- quantiles: node k = item / 32 is warp-uniform, lane = column.  Results go straight to global memory
    //      (4-byte stores 2400 bytes apart; the 8 nodes of a 32-byte sector are written by 8 warps within the same
    //      round, the L2 write-back merges them) -- no staging buffer, no result registers across the barrier ----
    const unsigned* endp = hist + lane;
    const float* colp = buf + lane;
    const long long o_col = (n0 + lane) * out_stride + (long long)g * nq;
    int fb = 0;
#pragma unroll 1
    for (int item = tid; item < n_items; item += kFastThreads) {
      const int k = item >> 5;
      float r = fnan;
      if (n > 0) r = bucket_quantile_node<true>(endp, colp, qs[k], n, S, cmin, cmax, scale, degenerate, fb);
      if (mode == 1) {
        if (col_ok) af[o_col + k] = r;
      } else if (pass == 0) {
        refq[k * 33 + lane] = r;
      } else if (col_ok) {
        const float rq = refq[k * 33 + lane];
        hist_q[o_col + k] = r;
        af[o_col + k] = kind == XSDBA_KIND_ADD ? __fsub_rn(rq, r) : __fdiv_rn(rq, r);
      }
    }
    if (__syncthreads_or(fb)) {
      // a bucket too large to select from (heavy ties, multi-scale data) or a degenerate column: the scattered
      // column holds all 1024 keys (valid values, then +inf) -- run K1f's sorter on it, select from the two runs
      const BktOut bo{af, hist_q, refq, qs, o_col, n_items, nq, S, mode, pass, kind, col_ok};
      bucket_sorter_select(buf, bo, n, cmin, cmax);
      }
  }
  if (normalize && mode == 0 && scaling && tid < 32 && n0 + tid < n_pts) {
    const float mr = mu[tid], mh = mu[32 + tid];  // scaling = get_correction(mu_hist, mu_ref)
    scaling[(n0 + tid) * n_groups + g] = kind == XSDBA_KIND_ADD ? __fsub_rn(mr, mh) : __fdiv_rn(mr, mh);
  }
}
