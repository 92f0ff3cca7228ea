// xsdba_b200: CUDA kernels (sm_100a) + C ABI for the quantile-mapping hot path of xsdba.
// See include/xsdba_b200.h for the boundary and DESIGN.md for the kernel inventory.
#include "../../include/xsdba_b200.h"
#include "common.cuh"
#include "sort.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <type_traits>
#include <vector>

using namespace xsdba;

namespace {

std::atomic<long long> g_launches{0};

constexpr int kThreads = 256;
// Narrow tiles (C <= 4 columns: one CTA per SM because the segment fills the shared memory) run with twice the
// threads so that the SM still has 16 warps to hide latencies with.
template <int C>
constexpr int threads_for_cols() { return C <= 4 ? 512 : kThreads; }
constexpr int kSortBytes = 128 * 1024;  // shared-memory budget of the column sorter

struct DevTable {
  int32_t* off = nullptr;   // [G+1]
  int32_t* rows = nullptr;  // time index, or -1 for a window slot outside the series
  int64_t total = 0;
  int32_t max_len = 0;
};

}  // namespace

// Chunks of consecutive groups whose windows share most of their rows (day-of-year groups with a 31-day window share
// 30/31 with their neighbours): the window trainer K1w loads and orders the UNION of a chunk's rows once.
struct WinTables {
  int32_t n_chunks = 0;
  int32_t* chunk_g = nullptr;   // [n_chunks + 1] first group of every chunk
  int32_t* urow_off = nullptr;  // [n_chunks + 1] offsets into urows
  int32_t* urows = nullptr;     // union of the chunk's window rows (time indices, ascending)
  unsigned long long* gmask = nullptr;  // [union rows] bit j: group chunk_g + j has the row in its window
  int32_t* mem_u = nullptr;     // [members] index of every exact group member inside the union rows of its chunk (K3w)
  int32_t max_union = 0, max_groups = 0;
};

struct xsdba_grouping {
  int64_t n_time = 0;
  int32_t n_groups = 0;
  int32_t window = 1;
  int device = 0;
  DevTable members;   // exact group members (window = 1), ascending time
  DevTable segments;  // members x window slots (aliases `members` when window == 1)
  int32_t* gidx = nullptr;  // [n_time] group of every time step (-1: none)
  WinTables win;      // only when window > 1 and the sharing pays (else n_chunks == 0)
};

namespace {

// =============================================================================================
// K1: group-segmented quantiles / train.  grid = (ceil(n_pts / C), n_groups).
// One CTA stages the whole segment of C neighbouring points in shared memory ([n_pad][C], column =
// point, so time-major inputs are read as contiguous C*sizeof(T)-byte rows), sorts each column and
// evaluates the nq type-7 quantiles.  mode 0: train (ref, hist -> af, hist_q[, scaling]);
// mode 1: quantiles of `ref` only -> af (used as `out`).
// =============================================================================================
template <typename T, int C>
__device__ void load_segment(T* sm, const T* __restrict__ src, long long n0, long long n_pts, long long sp,
                             long long st, const int32_t* __restrict__ rows, int S, int n_pad,
                             const JitterParams* jp = nullptr, long long seg_base = 0) {
  const int total = n_pad * C;
  // batches of UL elements per thread: first the row numbers, then the samples, so that UL independent loads are
  // in flight (one element at a time is a chain of two dependent global loads per iteration)
  constexpr int UL = 8;
  for (int base = threadIdx.x; base < total; base += blockDim.x * UL) {
    int t[UL];
#pragma unroll
    for (int u = 0; u < UL; ++u) {
      const int idx = base + u * blockDim.x;
      const int r = idx / C, c = idx % C;
      t[u] = (idx < total && r < S && n0 + c < n_pts) ? rows[r] : -1;
    }
    T v[UL];
#pragma unroll
    for (int u = 0; u < UL; ++u) {
      const int idx = base + u * blockDim.x;
      const int r = idx / C, c = idx % C;
      v[u] = t[u] >= 0 ? src[(n0 + c) * sp + (long long)t[u] * st] : Num<T>::nan();
      (void)r;
    }
#pragma unroll
    for (int u = 0; u < UL; ++u) {
      const int idx = base + u * blockDim.x;
      if (idx >= total) continue;
      const int r = idx / C, c = idx % C;
      T x = v[u];
      if (jp && r < S && n0 + c < n_pts)
        x = jitter_value<T>(x, *jp, (unsigned long long)((seg_base + r) * n_pts + n0 + c));
      sm[idx] = x;
    }
  }
}

// Per-column valid count (and float64 sum of the valid values when want_sum).  blockDim % C == 0, so
// idx % C is constant per thread.
template <typename T, int C>
__device__ void count_columns(const T* sm, int n_pad, int* cnt /*[C]*/, double* sum /*[C]*/, bool want_sum) {
  if (threadIdx.x < C) { cnt[threadIdx.x] = 0; sum[threadIdx.x] = 0.0; }
  __syncthreads();
  const int total = n_pad * C;
  int my_cnt = 0;
  double my_sum = 0.0;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const T v = sm[idx];
    if (!is_nan(v)) {
      ++my_cnt;
      if (want_sum) my_sum += (double)v;
    }
  }
  const int c = threadIdx.x % C;
  atomicAdd(&cnt[c], my_cnt);
  if (want_sum) atomicAdd(&sum[c], my_sum);
  __syncthreads();
}

// NaN -> +inf sort keys; optional DQM normalisation x + (-mu) or x * (1/mu) of the valid values
// (_adjustment.py:167-168 through utils.invert / apply_correction, utils.py:146-177).
template <typename T, int C>
__device__ void make_keys(T* sm, int n_pad, const int* cnt, const double* sum, int normalize, int kind) {
  const int c = threadIdx.x % C;
  T inv = (T)0;
  if (normalize) {
    const T mu = (T)(sum[c] / (double)cnt[c]);
    inv = kind == XSDBA_KIND_ADD ? -mu : Num<T>::div((T)1, mu);
  }
  const int total = n_pad * C;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    T v = sm[idx];
    if (is_nan(v)) {
      v = Num<T>::inf();
    } else if (normalize) {
      v = kind == XSDBA_KIND_ADD ? Num<T>::add(v, inv) : Num<T>::mul(v, inv);
      if (is_nan(v)) v = Num<T>::inf();  // (cannot happen for finite data; keeps the sorter NaN-free)
    }
    sm[idx] = v;
  }
  __syncthreads();
}

struct AdaptParams {
  int on;
  double thresh;
  unsigned long long seed;
  double* P0_ref;   // [n_pts][n_groups]
  double* P0_hist;  // [n_pts][n_groups]
  void* pth;        // [n_pts][n_groups], data dtype
};

// per-column count of valid values and of values <= thresh (utils.ecdf, utils.py:87-106)
template <typename T, int C>
__device__ void count_le_columns(const T* sm, int n_pad, int* n_valid, int* n_le, double thresh) {
  if (threadIdx.x < C) { n_valid[threadIdx.x] = 0; n_le[threadIdx.x] = 0; }
  __syncthreads();
  int a = 0, b = 0;
  for (int idx = threadIdx.x; idx < n_pad * C; idx += blockDim.x) {
    const T v = sm[idx];
    if (!is_nan(v)) { ++a; if ((double)v <= thresh) ++b; }
  }
  atomicAdd(&n_valid[threadIdx.x % C], a);
  atomicAdd(&n_le[threadIdx.x % C], b);
  __syncthreads();
}

// numba's np.nanquantile on a sorted column (see select_kernel mode 0)
template <typename T, int C>
__device__ double numba_nanquantile_sorted(const T* col, int n, double q) {
  if (n <= 0 || q != q) return Num<double>::nan();
  const double pct = q * 100.0;
  if (n == 1) return (double)col[0];
  if (pct == 100.0) return (double)col[(size_t)(n - 1) * C];
  if (pct == 0.0) return (double)col[0];
  const double rank = 1.0 + (double)(n - 1) * (pct / 100.0);
  const double f = floor(rank);
  const double m = rank - f;
  long long kk = (long long)f - 1;
  kk = kk < 0 ? 0 : (kk > n - 2 ? n - 2 : kk);
  const double lower = (double)col[(size_t)kk * C], upper = (double)col[(size_t)(kk + 1) * C];
  return __dadd_rn(__dmul_rn(lower, __dsub_rn(1.0, m)), __dmul_rn(upper, m));
}

template <typename T, int C>
__global__ void __launch_bounds__(threads_for_cols<C>())
train_kernel(const T* __restrict__ ref, const T* __restrict__ hist, long long n_pts, long long sp, long long st,
             const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows, int n_groups,
             const T* __restrict__ q, int nq, int kind, int normalize, int mode, T* __restrict__ af,
             T* __restrict__ hist_q, T* __restrict__ scaling, int n_pad, JitterParams jp, int use_jitter,
             const double* __restrict__ q64, AdaptParams ap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sum = reinterpret_cast<double*>(smem_raw);   // [C]
  double* mu_ref = sum + C;                            // [C]
  int* cnt = reinterpret_cast<int*>(mu_ref + C);       // [C]
  T* refq = reinterpret_cast<T*>(smem_raw + C * 24);   // [nq][C]  (24 = 8 + 8 + 4, padded to 8)
  T* sm = refq + (size_t)nq * C;                       // [n_pad][C]
  // frequency adaptation state (only used when ap.on): per column P0_ref, P0_hist, pth, dP0
  __shared__ double af_p0r[32], af_p0h[32], af_pth[32], af_dp0[32];
  __shared__ int af_le[32], af_nh[32];

  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int S = seg_off[g + 1] - seg_off[g];
  const int32_t* rows = seg_rows + seg_off[g];
  const long long out_stride = (long long)n_groups * nq;

  if (S == 0) {  // group without members: NaN rows (template of base.map_blocks, base.py:652-694)
    for (int item = threadIdx.x; item < C * nq; item += blockDim.x) {
      const int c = item / nq, k = item % nq;
      if (n0 + c >= n_pts) continue;
      const long long o = (n0 + c) * out_stride + (long long)g * nq + k;
      af[o] = Num<T>::nan();
      if (mode == 0) hist_q[o] = Num<T>::nan();
    }
    if (mode == 0 && scaling && threadIdx.x < C && n0 + threadIdx.x < n_pts)
      scaling[(n0 + threadIdx.x) * n_groups + g] = Num<T>::nan();
    return;
  }

  const int n_pass = mode == 0 ? 2 : 1;
  const bool adapt = ap.on && mode == 0;
  if (adapt) {
    // _adapt_freq (_processing.py:75-86): P0_hist = ecdf(hist, thresh) needs a counting pre-pass over the
    // (jittered) hist segment, because pth is taken from the sorted ref at that rank
    load_segment<T, C>(sm, hist, n0, n_pts, sp, st, rows, S, n_pad, use_jitter ? &jp : nullptr, (long long)seg_off[g]);
    __syncthreads();
    count_le_columns<T, C>(sm, n_pad, af_nh, af_le, ap.thresh);
    if (threadIdx.x < C) af_p0h[threadIdx.x] = (double)af_le[threadIdx.x] / (double)af_nh[threadIdx.x];
    __syncthreads();
  }
  for (int pass = 0; pass < n_pass; ++pass) {
    const T* src = pass == 0 ? ref : hist;
    // jitter applies to hist only, after the window gather (_adjustment.py:58-67, 80-81)
    load_segment<T, C>(sm, src, n0, n_pts, sp, st, rows, S, n_pad, (use_jitter && pass == 1) ? &jp : nullptr,
                       (long long)seg_off[g]);
    __syncthreads();
    if (adapt && pass == 0) {
      count_le_columns<T, C>(sm, n_pad, af_nh, af_le, ap.thresh);   // (af_nh is reused: ref counts here)
      if (threadIdx.x < C) af_p0r[threadIdx.x] = (double)af_le[threadIdx.x] / (double)af_nh[threadIdx.x];
      __syncthreads();
    }
    // with frequency adaptation dqm_train normalises the ADAPTED hist (_adjustment.py:155-168: preprocessing comes
    // first) and pth is a quantile of the RAW ref: both passes sort raw values here and normalise afterwards
    const int norm_now = adapt ? 0 : normalize;
    count_columns<T, C>(sm, n_pad, cnt, sum, norm_now != 0);
    make_keys<T, C>(sm, n_pad, cnt, sum, norm_now, kind);
    sort_columns<T, C>(sm, n_pad);
    // mean of the (sorted, NaN-free below cnt) column, then x (+|*) inv(mean): the late normalisation of the adapt path
    auto normalise_sorted = [&]() {
      if (threadIdx.x < C) sum[threadIdx.x] = 0.0;
      __syncthreads();
      const int c = threadIdx.x % C;
      double my_sum = 0.0;
      for (int idx = threadIdx.x; idx < n_pad * C; idx += blockDim.x)
        if (idx / C < cnt[c]) my_sum += (double)sm[idx];
      atomicAdd(&sum[c], my_sum);
      __syncthreads();
      const T mu = (T)(sum[c] / (double)cnt[c]);
      const T inv = kind == XSDBA_KIND_ADD ? -mu : Num<T>::div((T)1, mu);
      for (int idx = threadIdx.x; idx < n_pad * C; idx += blockDim.x) {
        if (idx / C >= cnt[c]) continue;
        T v = kind == XSDBA_KIND_ADD ? Num<T>::add(sm[idx], inv) : Num<T>::mul(sm[idx], inv);
        sm[idx] = is_nan(v) ? Num<T>::inf() : v;
      }
      __syncthreads();
    };
    if (adapt && pass == 0 && threadIdx.x < C) {
      // dP0 and pth = vecquantiles(ref, P0_hist).where(dP0 > 0)   (_processing.py:84-99)
      const int c = threadIdx.x;
      const double p0r = af_p0r[c], p0h = af_p0h[c];
      const double dp0 = p0h == 0.0 ? Num<double>::nan() : (p0h - p0r) / p0h;
      double pth = Num<double>::nan();
      if (dp0 > 0.0) pth = (double)(T)numba_nanquantile_sorted<T, C>(sm + c, cnt[c], p0h);
      af_dp0[c] = dp0; af_pth[c] = pth;
      if (n0 + c < n_pts) {
        const long long o = (n0 + c) * n_groups + g;
        ap.P0_ref[o] = p0r; ap.P0_hist[o] = p0h; reinterpret_cast<T*>(ap.pth)[o] = (T)pth;
      }
    }
    if (adapt && pass == 0 && normalize) {
      __syncthreads();
      normalise_sorted();
      sort_columns<T, C>(sm, n_pad);
    }
    if (adapt && pass == 1) {
      // replace the excess "dry" values: sorted position i has tie-broken percentile rank i/(n-1)
      // (_processing.py:104-122); then restore the order
      __syncthreads();
      for (int idx = threadIdx.x; idx < n_pad * C; idx += blockDim.x) {
        const int i = idx / C, c = idx % C;
        const int n = cnt[c];
        if (i >= n || !(af_dp0[c] > 0.0) || n < 2) continue;
        const double rnk = (double)i / (double)(n - 1);
        const double p0s = af_p0h[c];
        const bool keep = (rnk < (af_p0r[c] / af_p0h[c]) * p0s) || (rnk > p0s);
        if (!keep) {
          const double u = (double)(T)hash_uniform(ap.seed, (unsigned long long)(((long long)seg_off[g] + i) * n_pts + n0 + c) * 2 + 1);
          sm[idx] = (T)((af_pth[c] - ap.thresh) * u + ap.thresh);
        }
      }
      __syncthreads();
      if (normalize) normalise_sorted();
      sort_columns<T, C>(sm, n_pad);
    }
    // quantiles: item -> (point c, node k), k fastest so that global writes are contiguous
    for (int item = threadIdx.x; item < C * nq; item += blockDim.x) {
      const int c = item / nq, k = item % nq;
      const T v = quantile_sorted<T, C>(sm + c, cnt[c], S, q64 ? q64[k] : (double)q[k]);
      if (mode == 1) {
        if (n0 + c < n_pts) af[(n0 + c) * out_stride + (long long)g * nq + k] = v;
      } else if (pass == 0) {
        refq[k * C + c] = v;
      } else if (n0 + c < n_pts) {
        const long long o = (n0 + c) * out_stride + (long long)g * nq + k;
        const T rq = refq[k * C + c];
        hist_q[o] = v;
        af[o] = kind == XSDBA_KIND_ADD ? Num<T>::sub(rq, v) : Num<T>::div(rq, v);  // utils.py:130-143
      }
    }
    if (normalize && mode == 0 && threadIdx.x < C) {
      const int c = threadIdx.x;
      const T mu = (T)(sum[c] / (double)cnt[c]);
      if (pass == 0) {
        mu_ref[c] = (double)mu;
      } else if (scaling && n0 + c < n_pts) {
        const T mr = (T)mu_ref[c];  // scaling = get_correction(mu_hist, mu_ref)  (_adjustment.py:177-179)
        scaling[(n0 + c) * n_groups + g] = kind == XSDBA_KIND_ADD ? Num<T>::sub(mr, mu) : Num<T>::div(mr, mu);
      }
    }
    __syncthreads();
  }
}

// =============================================================================================
// K1f: fast train / quantile kernel -- float32, time-major (stride_pt == 1), segments <= 1024 slots.
// grid = (ceil(n_pts/32), n_groups), 1024 threads, one CTA per SM (128 KB sort buffer).  Column = lane
// = gridpoint: the 128-byte rows of the time-major input land in shared memory as they are, and every
// later shared-memory access is bank-conflict free.  See sort.cuh for the sorter.
// =============================================================================================
constexpr int kFastThreads = 1024;
constexpr int kFastMaxNq = 128;

struct FastSmem {
  static constexpr size_t buf = 0;                                  // float [1024][32]
  static constexpr size_t rows = buf + 1024 * 32 * 4;               // int   [1024]
  static constexpr size_t pcnt = rows + 1024 * 4;                   // int   [32][32]
  static constexpr size_t psum = pcnt + 32 * 32 * 4;                // double[32][32]
  static constexpr size_t cnt = psum + 32 * 32 * 8;                 // int   [2][32]
  static constexpr size_t mu = cnt + 2 * 32 * 4;                    // float [2][32]  (ref, hist means)
  static constexpr size_t q = mu + 2 * 32 * 4;                      // double[kFastMaxNq]
  static constexpr size_t refq = q + kFastMaxNq * 8;                // float [nq][33]
  static __host__ __device__ constexpr size_t total(int nq) { return refq + (size_t)3 * nq * 33 * 4; }
};

// base + row * stride_bytes as ONE IMAD.WIDE (the C expression compiles to a 64 x 64 multiply sequence)
__device__ __forceinline__ const char* row_address(const char* base, int row, int stride_bytes) {
  long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(row), "r"(stride_bytes), "l"(reinterpret_cast<long long>(base)));
  return reinterpret_cast<const char*>(r);
}

template <bool JITTER, bool NORM>
__global__ void __launch_bounds__(kFastThreads, 1)
train_fast_kernel(const float* __restrict__ ref, const float* __restrict__ hist, long long n_pts, long long st,
                  const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows, int n_groups,
                  const float* __restrict__ q, int nq, int kind, int normalize_arg, int mode, float* __restrict__ af,
                  float* __restrict__ hist_q, float* __restrict__ scaling, JitterParams jp, int use_jitter,
                  const double* __restrict__ q64, int stagger_ns) {
  // <false, false> is the lean EQM instantiation; NORM adds dqm_train's normalisation, JITTER the jitter options
  const int normalize = NORM ? normalize_arg : 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* buf = reinterpret_cast<float*>(smem_raw + FastSmem::buf);
  int* rows_tab = reinterpret_cast<int*>(smem_raw + FastSmem::rows);
  int* pcnt = reinterpret_cast<int*>(smem_raw + FastSmem::pcnt);
  double* psum = reinterpret_cast<double*>(smem_raw + FastSmem::psum);
  int* cnt = reinterpret_cast<int*>(smem_raw + FastSmem::cnt);
  float* mu = reinterpret_cast<float*>(smem_raw + FastSmem::mu);
  double* qs = reinterpret_cast<double*>(smem_raw + FastSmem::q);
  float* refq = reinterpret_cast<float*>(smem_raw + FastSmem::refq);
  float* outh = refq + (size_t)nq * 33;
  float* outa = outh + (size_t)nq * 33;

  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * 32;
  const int S = seg_off[g + 1] - seg_off[g];
  const long long out_stride = (long long)n_groups * nq;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float fnan = Num<float>::nan(), finf = Num<float>::inf();
  const int nqp = (nq + 31) & ~31;

  if (S == 0) {  // group without members: NaN rows
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
  {
    const int32_t* rows = seg_rows + seg_off[g];
    rows_tab[tid] = tid < S ? rows[tid] : -1;
    if (tid < nq) qs[tid] = q64 ? q64[tid] : (double)q[tid];
  }
  __syncthreads();

  const bool col_ok = n0 + lane < n_pts;
  const int half = warp & 1;
  const int n_pass = mode == 0 ? 2 : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    // columns past n_pts read column n0 instead (always valid memory); their results are never written out
    const char* __restrict__ srcb = reinterpret_cast<const char*>((pass == 0 ? ref : hist) + n0 + (col_ok ? lane : 0));
    const int st4 = (int)st * 4;  // byte stride between time steps (launcher guarantees it fits)
    // ---- load: warp w takes slots w, w+32, ...; slot s -> half s&1, in-half row s>>1 ------------
    float v[32];
    int my_cnt = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int t = rows_tab[warp + 32 * i];
      const char* pa = row_address(srcb, t, st4);
      v[i] = fnan;
      if (t >= 0) v[i] = *reinterpret_cast<const float*>(pa);
    }
    if (JITTER && use_jitter && pass == 1) {  // hist only, per window slot (_adjustment.py:58-67)
      // The hash needs ~20 registers per value; unrolled over the 32 values of a thread it spilled 400-1000 bytes.
      // The values take a round trip through the thread's own slots of the (free) sort buffer instead and are
      // jittered one at a time in a rolled loop.
      const long long seg_base = seg_off[g];
      float* own = buf + ((size_t)half * 512 + (warp >> 1)) * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) own[(size_t)i * 16 * 32] = v[i];
#pragma unroll 1
      for (int i = 0; i < 32; ++i)
        own[(size_t)i * 16 * 32] = jitter_value<float>(
            own[(size_t)i * 16 * 32], jp, (unsigned long long)((seg_base + warp + 32 * i) * n_pts + n0 + lane));
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = own[(size_t)i * 16 * 32];
    }
    double my_sum = 0.0;
    if (normalize) {
#pragma unroll
      for (int i = 0; i < 32; ++i) if (v[i] == v[i]) my_sum += (double)v[i];
    }
    // lean instantiation: valid count and the sort key (NaN -> +inf, NaNs sort last) from one predicate per element
    if (!NORM) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        asm("{\n\t.reg .pred p;\n\tsetp.num.f32 p, %1, %1;\n\t@p add.s32 %0, %0, 1;\n\t@!p mov.f32 %1, %2;\n\t}"
            : "+r"(my_cnt), "+f"(v[i]) : "f"(finf));
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) my_cnt += (v[i] == v[i]) ? 1 : 0;
    }
    pcnt[warp * 32 + lane] = my_cnt;
    if (normalize) psum[warp * 32 + lane] = my_sum;
    __syncthreads();
    if (tid < 64) {  // (h, c): valid count of each half of each column
      const int h = tid >> 5;
      int n = 0;
#pragma unroll
      for (int w = 0; w < 16; ++w) n += pcnt[(2 * w + h) * 32 + lane];
      cnt[h * 32 + lane] = n;
    }
    if (normalize && tid >= 64 && tid < 96) {
      double s = 0.0; int n = 0;
      for (int w = 0; w < 32; ++w) { s += psum[w * 32 + lane]; n += pcnt[w * 32 + lane]; }
      mu[pass * 32 + lane] = (float)(s / (double)n);
    }
    if (normalize) {
      __syncthreads();
      // dqm_train: x + (-mean) or x * (1/mean)   (_adjustment.py:167-168)
      const float m = mu[pass * 32 + lane];
      const float inv = kind == XSDBA_KIND_ADD ? -m : __fdiv_rn(1.0f, m);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = kind == XSDBA_KIND_ADD ? __fadd_rn(v[i], inv) : __fmul_rn(v[i], inv);
    }
    {
      float* dst = buf + ((size_t)half * 512 + (warp >> 1)) * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[(size_t)i * 16 * 32] = NORM ? ((v[i] == v[i]) ? v[i] : finf) : v[i];
    }
    __syncthreads();
    sort_halves_512(buf, stagger_ns);
    // ---- quantiles from the two sorted runs -------------------------------------------------------
    for (int item = tid; item < nq * 32; item += kFastThreads) {
      const int k = item >> 5;  // warp-uniform node, lane = column
      const int nA = cnt[lane], nB = cnt[32 + lane], n = nA + nB;
      const float* A = buf + lane;
      const float* B = buf + 512 * 32 + lane;
      float res = fnan;
      if (n > 0) {
        const float amax = nA > 0 ? A[(size_t)(nA - 1) * 32] : -finf;
        const float bmax = nB > 0 ? B[(size_t)(nB - 1) * 32] : -finf;
        const float vmax = fmaxf(amax, bmax);
        const double vi = (double)(n - 1) * qs[k];  // nbutils.py:131
        float left, right, gamma;
        if (vi >= (double)(n - 1)) {  // nbutils.py:47-51: position -1 of the full-length sorted row
          left = right = (n < S) ? fnan : vmax;
          gamma = (float)(vi + 1.0);
        } else if (vi < 0.0) {
          left = right = two_run_at(A, nA, B, nB, 0);
          gamma = (float)vi;
        } else {
          const int i = (int)vi;
          two_run_pair(A, nA, B, nB, i, left, right);
          gamma = (float)(vi - (double)i);  // nbutils.py:142
        }
        const float diff = right - left;
        res = gamma >= 0.5f ? __fmaf_rn(-diff, 1.0f - gamma, right) : __fmaf_rn(diff, gamma, left);
        if (res != res) res = vmax;  // nbutils.py:146
      }
      if (mode == 1) {
        outa[k * 33 + lane] = res;
      } else if (pass == 0) {
        refq[k * 33 + lane] = res;
      } else {
        const float rq = refq[k * 33 + lane];
        outh[k * 33 + lane] = res;
        outa[k * 33 + lane] = kind == XSDBA_KIND_ADD ? __fsub_rn(rq, res) : __fdiv_rn(rq, res);
      }
    }
    __syncthreads();
  }
  // ---- contiguous table writes: a warp writes the nq values of one point ---------------------------
  for (int item = tid; item < 32 * nqp; item += kFastThreads) {
    const int c = item / nqp, k = item % nqp;
    if (k >= nq || n0 + c >= n_pts) continue;
    const long long o = (n0 + c) * out_stride + (long long)g * nq + k;
    af[o] = outa[k * 33 + c];
    if (mode == 0) hist_q[o] = outh[k * 33 + c];
  }
  if (normalize && mode == 0 && scaling && tid < 32 && n0 + tid < n_pts) {
    const float mr = mu[tid], mh = mu[32 + tid];  // scaling = get_correction(mu_hist, mu_ref)
    scaling[(n0 + tid) * n_groups + g] = kind == XSDBA_KIND_ADD ? __fsub_rn(mr, mh) : __fdiv_rn(mr, mh);
  }
}

#include "train_bucket.cuh"
#include "train_window.cuh"
#include "rank_window.cuh"

// =============================================================================================
// Table staging shared by the adjust kernels: rows r-1, r, r+1 (cyclic) of the tile's tables into
// shared memory as xs/ys[slot][k][C] (column = point), NaN nodes dropped, bounds / constants of the
// centre row recorded.  Global reads are contiguous (a point's nq nodes are adjacent in memory); the
// transposition goes through a small padded staging buffer so that both sides are conflict free.
// =============================================================================================
template <typename T, int C>
__device__ void stage_tables(Tables<T, C>& tb, T* stage /*[2][C][nq|1]*/, long long n0, long long n_pts, int r,
                             bool grouped) {
  const int nq = tb.nq;
  const int pitch = nq | 1;
  T* sx = stage;
  T* sy = stage + (size_t)C * pitch;
  for (int s = 0; s < (grouped ? 3 : 1); ++s) {
    const int slot = grouped ? s : 1;
    if (tb.centre_only && slot != 1) continue;
    const int g = grouped ? (r + slot - 1 + tb.G) % tb.G : 0;
    int has_nan = 0;
    for (int idx = threadIdx.x; idx < C * nq; idx += blockDim.x) {
      const int c = idx / nq, k = idx - c * nq;
      T xv = Num<T>::nan(), yv = Num<T>::nan();
      if (n0 + c < n_pts) {
        const long long o = (n0 + c) * tb.pt_stride + (long long)g * nq + k;
        xv = tb.x_shared ? tb.gx[k] : tb.gx[o];
        yv = tb.gy[o];
      }
      has_nan |= (is_nan(xv) || is_nan(yv));
      sx[c * pitch + k] = xv;
      sy[c * pitch + k] = yv;
    }
    has_nan = __syncthreads_or(has_nan);
    T* xs = tb.xsl[slot];
    T* ys = tb.ysl[slot];
    if (!has_nan) {  // common case: plain transposed copy by all threads (+inf padding rows)
      for (int idx = threadIdx.x; idx < C * tb.ld; idx += blockDim.x) {
        const int c = idx % C, k = idx / C;
        xs[(size_t)k * C + c] = k < nq ? sx[c * pitch + k] : Num<T>::inf();
        if (k < nq) ys[(size_t)k * C + c] = sy[c * pitch + k];
      }
      if (threadIdx.x < C) {
        const int c = threadIdx.x;
        tb.nvl[slot][c] = nq;
        if (slot == 1) {
          tb.blo[c] = sx[c * pitch]; tb.bhi[c] = sx[c * pitch + nq - 1];
          tb.clo[c] = sy[c * pitch]; tb.chi[c] = sy[c * pitch + nq - 1];
        }
      }
    } else if (threadIdx.x < C) {  // compaction: one thread per point
      const int c = threadIdx.x;
      T blo = Num<T>::nan(), bhi = Num<T>::nan(), clo = Num<T>::nan(), chi = Num<T>::nan();
      bool have_b = false, have_c = false;
      int w = 0;
      for (int k = 0; k < nq; ++k) {
        const T xv = sx[c * pitch + k], yv = sy[c * pitch + k];
        if (!is_nan(xv)) { if (!have_b) { blo = xv; have_b = true; } bhi = xv; }
        if (!is_nan(yv)) { if (!have_c) { clo = yv; have_c = true; } chi = yv; }
        if (!is_nan(xv) && !is_nan(yv)) { xs[(size_t)w * C + c] = xv; ys[(size_t)w * C + c] = yv; ++w; }
      }
      for (int k = w; k < tb.ld; ++k) xs[(size_t)k * C + c] = Num<T>::inf();
      tb.nvl[slot][c] = w;
      if (slot == 1) { tb.blo[c] = blo; tb.bhi[c] = bhi; tb.clo[c] = clo; tb.chi[c] = chi; }
    }
    __syncthreads();
  }
}

// slots = 3: rows g-1, g, g+1 staged; slots = 1: the centre row only (Tables::centre_only)
template <typename T, int C>
__device__ Tables<T, C> carve_tables(unsigned char* base, int nq, int slots = 3) {
  Tables<T, C> tb;
  tb.nq = nq;
  int top = 1;
  while (top * 2 <= nq) top *= 2;
  tb.top = top;
  tb.ld = 2 * top;
  tb.centre_only = slots == 1;
  T* xs = reinterpret_cast<T*>(base);
  T* ys = xs + (size_t)slots * tb.ld * C;
  tb.blo = ys + (size_t)slots * tb.ld * C;
  tb.bhi = tb.blo + C;
  tb.clo = tb.bhi + C;
  tb.chi = tb.clo + C;
  int* nv = reinterpret_cast<int*>(tb.chi + C);
  for (int sl = 0; sl < 3; ++sl) {
    const int pos = slots == 1 ? 0 : sl;
    tb.xsl[sl] = xs + (size_t)pos * tb.ld * C;
    tb.ysl[sl] = ys + (size_t)pos * tb.ld * C;
    tb.nvl[sl] = nv + sl * C;
  }
  return tb;
}
template <typename T, int C>
__host__ __device__ constexpr size_t tables_bytes(int nq, int slots = 3) {
  int top = 1;
  while (top * 2 <= nq) top *= 2;
  return ((size_t)2 * slots * (nq > 0 ? 2 * top : 0) * C + 4 * C) * sizeof(T) + 3 * C * sizeof(int);
}
template <typename T, int C>
__host__ __device__ constexpr size_t stage_bytes(int nq) { return (size_t)2 * C * (nq | 1) * sizeof(T); }

// =============================================================================================
// K2: streaming adjust, EQM / DQM flavour.  grid = (ceil(n_pts/32), n_groups); each warp walks the
// time steps of its group four at a time, lane = point: read sim (coalesced for time-major), look
// the factor up (branch-free binary search in the staged table), apply, write scen.
// =============================================================================================
template <typename T, int C = 32>
__global__ void __launch_bounds__(kThreads)
adjust_kernel(const T* __restrict__ sim, long long n_pts, long long sp, long long st,
              const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows, int n_groups,
              const T* __restrict__ af, const T* __restrict__ hist_q, int nq, int interp, int extrap, int kind,
              T* __restrict__ scen, const unsigned* __restrict__ gate_count = nullptr, unsigned gate_cap = 0) {
  // C = points per CTA: 32 (lane = point) unless the tables of a very fine quantile grid need the shared memory
  constexpr int U = 4;
  if (gate_count && *gate_count <= gate_cap) return;  // overflow fallback of K2t: nothing to redo
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Tables<T, C> tb = carve_tables<T, C>(smem_raw, nq);
  T* stage = reinterpret_cast<T*>(smem_raw + ((tables_bytes<T, C>(nq) + 15) & ~(size_t)15));
  tb.gx = hist_q; tb.gy = af; tb.x_shared = false; tb.G = n_groups; tb.pt_stride = (long long)n_groups * nq;

  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], m1 = mem_off[g + 1];
  if (m0 == m1) return;
  const bool grouped = n_groups > 1;
  stage_tables<T, C>(tb, stage, n0, n_pts, g, grouped);
  if (interp == XSDBA_INTERP_CUBIC && !grouped) stage_cubic<T, C>(tb);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = n0 + lane;
  if (pt >= n_pts || lane >= C) return;
  const T* __restrict__ src = sim + pt * sp;
  T* __restrict__ dst = scen + pt * sp;
  for (int m = m0 + warp * U; m < m1; m += n_warps * U) {
    long long o[U];
    T x[U], f[U];
#pragma unroll
    for (int j = 0; j < U; ++j) o[j] = (long long)mem_rows[min(m + j, m1 - 1)] * st;
#pragma unroll
    for (int j = 0; j < U; ++j) x[j] = src[o[j]];
    if (grouped) lookup_2d_nearest_n<T, T, C, U>(tb, lane, pt, g, x, f, extrap);
    else {
#pragma unroll
      for (int j = 0; j < U; ++j) f[j] = lookup_1d<T, T, C>(tb, lane, x[j], interp, extrap);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) if (m + j < m1) dst[o[j]] = apply_corr<T>(x[j], f[j], kind);
  }
}

// =============================================================================================
// K2f: lean adjust kernel for the headline case -- float32, grouped (month / dayofyear) 2-D nearest
// rule.  Same staging as K2; the per-sample work is branch-free float32: 6-step binary search on the
// staged centre row, nearest of the two bracketing nodes, extrapolation override.  A sample goes to
// the exact float64 / cross-row routine (lookup_2d_nearest_n) only when float32 cannot decide:
// the two candidate distances agree to ~1e-5 relative (possible tie), the nearest in-row node is
// >= ~1 away (a neighbouring row may win), or the sample is NaN.
// =============================================================================================
constexpr int kAdjMaxRows = 2048;  // member rows of one group kept in shared memory by K2f

template <int TOP>
__global__ void __launch_bounds__(kThreads, 3)
adjust_fast_kernel(const float* __restrict__ sim, long long n_pts, long long sp, long long st,
                   const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows, int n_groups,
                   const float* __restrict__ af, const float* __restrict__ hist_q, int nq, int extrap, int kind,
                   float* __restrict__ scen) {
  constexpr int C = 32;
  constexpr int U = 8;         // rows per batch; two batches in flight per warp (software pipeline)
  constexpr int LD = 2 * TOP;  // rows per staged slot (tb.ld)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Tables<float, C> tb = carve_tables<float, C>(smem_raw, nq);
  float* stage = reinterpret_cast<float*>(smem_raw + ((tables_bytes<float, C>(nq) + 15) & ~(size_t)15));
  int* rows_sm = reinterpret_cast<int*>(stage + 2 * C * (nq | 1));
  tb.gx = hist_q; tb.gy = af; tb.x_shared = false; tb.G = n_groups; tb.pt_stride = (long long)n_groups * nq;

  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], n_rows = mem_off[g + 1] - m0;
  if (n_rows == 0) return;
  for (int i = threadIdx.x; i < n_rows; i += blockDim.x) rows_sm[i] = mem_rows[m0 + i];
  stage_tables<float, C>(tb, stage, n0, n_pts, g, true);  // (syncs)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = n0 + lane;
  if (pt >= n_pts) return;
  const float* __restrict__ src = sim + pt * sp;
  float* __restrict__ dst = scen + pt * sp;
  // centre row of this lane's point (column = lane); rows [n, LD) of xs hold +inf
  const float* xs = reinterpret_cast<const float*>(smem_raw) + (size_t)LD * C + lane;
  const float* ys = xs + (size_t)3 * LD * C;
  const int n = tb.nvl[1][lane];
  const float blo = tb.blo[lane], bhi = tb.bhi[lane];
  const float fnan = Num<float>::nan(), finf = Num<float>::inf();
  const float clo = extrap == 0 ? tb.clo[lane] : fnan, chi = extrap == 0 ? tb.chi[lane] : fnan;

  float xn[U];
  int m = warp * U;
  if (m < n_rows) {
#pragma unroll
    for (int j = 0; j < U; ++j) xn[j] = src[(long long)rows_sm[min(m + j, n_rows - 1)] * st];
  }
  for (; m < n_rows; m += n_warps * U) {
    float x[U];
    int pos[U];
#pragma unroll
    for (int j = 0; j < U; ++j) { x[j] = xn[j]; pos[j] = 0; }
    const int mn = m + n_warps * U;  // prefetch the next batch before working on this one
    if (mn < n_rows) {
#pragma unroll
      for (int j = 0; j < U; ++j) xn[j] = src[(long long)rows_sm[min(mn + j, n_rows - 1)] * st];
    }
#pragma unroll
    for (int step = TOP; step > 0; step >>= 1) {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const float v = xs[(size_t)(pos[j] + step - 1) * C];
        pos[j] = v < x[j] ? pos[j] + step : pos[j];
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int i = pos[j];  // #nodes < x  (<= n)
      const int il = i > 0 ? i - 1 : 0;
      const float xl = xs[(size_t)il * C], xh = xs[(size_t)i * C];   // xs[n] = +inf
      const float yl = ys[(size_t)il * C], yh = ys[(size_t)(i < n ? i : il) * C];
      const float dl = i > 0 ? x[j] - xl : finf;
      const float dh = xh - x[j];
      const float dmin = fminf(dl, dh);
      float f = dh < dl ? yh : yl;
      const bool below = x[j] < blo, above = x[j] > bhi;
      // float32 is decisive unless: near-tie, far node (cross-row candidates), NaN sample / empty row
      const bool sure = (fabsf(dl - dh) > 1e-5f * dmin) && (dmin < 0.99f);
      f = below ? clo : f;
      f = above ? chi : f;
      if (!(sure || below || above)) {
        const float xa[1] = {x[j]};
        float fo[1];
        lookup_2d_nearest_n<float, float, C, 1>(tb, lane, pt, g, xa, fo, extrap);
        f = fo[0];
      }
      if (m + j < n_rows)
        dst[(long long)rows_sm[m + j] * st] = kind == XSDBA_KIND_ADD ? __fadd_rn(x[j], f) : __fmul_rn(x[j], f);
    }
  }
}

// =============================================================================================
// K2p + K2t: packed decision tables and the tile-persistent adjust kernel (float32, grouped nearest).
//
// The 2-D nearest rule restricted to one table row is a step function of the sample: node k answers for the
// samples between the midpoints towards nodes k-1 and k+1.  Float32 arithmetic on the sample cannot be trusted near
// a midpoint (tie) or when the nearest in-row node is >= 1 away (a node of a neighbouring row, one unit of group
// coordinate away, may be nearer), so K2p `pack_tables_kernel` turns every (point, group) row of the caller's
// point-major af / hist_q [N][G][nq] into nq + 2 INTERVALS with float32 bounds (directed rounding, towards the node):
//     interval 0        : (-inf, blo)                             -> first factor (constant extrapolation) or NaN
//     interval k + 1    : [lo_k, hi_k] around node k, inside its midpoints (shrunk by 1e-5 of the node gap),
//                          inside node +- 0.99 and inside [blo, bhi]                      -> af[k]
//     interval nq' + 1  : (bhi, +inf]                             -> last factor or NaN
// hi is ascending.  A sample x belongs to interval i = #{hi < x} and is DECIDED iff x >= lo_i; the undecided ones
// (gaps between intervals: ~1e-4 of the samples) go to the exact float64 / cross-row second pass as before.
// Rows are stored column-major per tile of 32 points -- hi[R][32], lo[R][32], y[R][32], R = nq + 2, then nv[32]
// (number of NaN-free nodes, for the second pass) -- the exact shared-memory image of the lookup, so K2t fetches a
// row with three bulk async copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier).
//
// K2t `adjust_tile_kernel`: one CTA per tile of 32 points walks ALL groups; a ring of two slots holds the current
// group's intervals while the next one streams in.  Per sample: a 6-step branch-free search on hi, one compare
// against lo, one load of y -- about half the instructions of a search on the nodes followed by the distance logic.
// =============================================================================================
__host__ __device__ constexpr size_t packed_slot_floats(int nq) { return (size_t)(3 * (nq + 2) + 1) * 32; }

__global__ void __launch_bounds__(kThreads)
pack_tables_kernel(const float* __restrict__ af, const float* __restrict__ hist_q, long long n_pts, int n_groups,
                   int nq, int extrap, float* __restrict__ packed) {
  constexpr int C = 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // the row of every point as it lies in memory (nq contiguous nodes), pitch odd: column-wise reads are conflict free
  const int pitch = nq | 1;
  float* sx = reinterpret_cast<float*>(smem_raw);  // [C][pitch] hist_q, compacted in place when a row has NaNs
  float* sy = sx + C * pitch;                      // [C][pitch] af
  __shared__ int nv[C], bad[C];
  __shared__ float s_blo[C], s_bhi[C], s_clo[C], s_chi[C];
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const long long pt_stride = (long long)n_groups * nq;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const float finf = Num<float>::inf(), fnan = Num<float>::nan();
  if (threadIdx.x < C) bad[threadIdx.x] = 0;
  __syncthreads();
  int has_nan = 0;
  for (int c = warp; c < C; c += n_warps) {  // one point per warp and round: contiguous reads, no index division
    const bool ok = n0 + c < n_pts;
    const long long o = (n0 + c) * pt_stride + (long long)g * nq;
    int flag = 0;
    for (int k = lane; k < nq; k += 32) {
      const float xv = ok ? hist_q[o + k] : fnan, yv = ok ? af[o + k] : fnan;
      sx[c * pitch + k] = xv;
      sy[c * pitch + k] = yv;
      flag |= (xv != xv || yv != yv) ? 1 : 0;
      flag |= (fabsf(xv) == finf) ? 2 : 0;
    }
    flag = __reduce_or_sync(0xffffffffu, flag);
    if (lane == 0 && flag) bad[c] = flag;
    has_nan |= flag & 1;
  }
  has_nan = __syncthreads_or(has_nan);
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    float* x = sx + c * pitch;
    float* y = sy + c * pitch;
    if (!(bad[c] & 1)) {
      nv[c] = nq; s_blo[c] = x[0]; s_bhi[c] = x[nq - 1]; s_clo[c] = y[0]; s_chi[c] = y[nq - 1];
    } else {  // NaN nodes are dropped (utils.py:381-382); bounds / constants from the first / last non-NaN of each table
      float blo = fnan, bhi = fnan, clo = fnan, chi = fnan;
      bool have_b = false, have_c = false;
      int w = 0;
      for (int k = 0; k < nq; ++k) {
        const float xv = x[k], yv = y[k];
        if (xv == xv) { if (!have_b) { blo = xv; have_b = true; } bhi = xv; }
        if (yv == yv) { if (!have_c) { clo = yv; have_c = true; } chi = yv; }
        if (xv == xv && yv == yv) { x[w] = xv; y[w] = yv; ++w; }
      }
      nv[c] = w; s_blo[c] = blo; s_bhi[c] = bhi; s_clo[c] = clo; s_chi[c] = chi;
    }
  }
  __syncthreads();
  const float* xs = sx + lane * pitch;
  const float* ys = sy + lane * pitch;
  const int n = nv[lane];
  const bool usable = n > 0 && !(bad[lane] & 2);  // otherwise every (non-NaN) sample goes to the exact pass
  const float blo = s_blo[lane], bhi = s_bhi[lane];
  const float clo = extrap == 0 ? s_clo[lane] : fnan, chi = extrap == 0 ? s_chi[lane] : fnan;
  const int R = nq + 2;
  float* out = packed + ((size_t)blockIdx.x * n_groups + g) * packed_slot_floats(nq);
  // float32 arithmetic with directed rounding, every bound moved TOWARDS its node: the decided intervals can only be
  // narrower than the exact ones (midpoint +- 1e-5 gap, node +- 0.99), never wider
  const float tiny = __int_as_float(1);  // smallest subnormal: x -+ tiny rounds to the neighbour of x
  // Row t of the image: thread t of a column handles the GAP between nodes t - 1 and t once (one midpoint, one
  // half-width) and writes the two bounds that depend on it, hi[t] (end of node t - 1's interval) and lo[t + 1] (start
  // of node t's), plus the factor y[t].  Rows 0 and 1 also need lo[0] / lo[1].
  float* out_hi = out;
  float* out_lo = out + (size_t)R * C;
  float* out_y = out + (size_t)2 * R * C;
  for (int t = warp; t < R; t += n_warps) {
    float hi = finf, lo_next = finf, y = fnan;
    if (usable) {
      if (t == 0) {
        hi = __fadd_rd(blo, -tiny); y = clo;
        lo_next = fmaxf(blo, __fadd_ru(xs[0], -0.99f));            // lo[1]: node 0 answers from blo on
      } else if (t <= n) {
        const float xl = xs[t - 1];                                 // node t - 1 = interval t
        y = ys[t - 1];
        if (t < n) {
          const float xh = xs[t];
          const float w = __fmul_ru(1e-5f, __fsub_ru(xh, xl));
          hi = fminf(__fsub_rd(__fmul_rd(0.5f, __fadd_rd(xl, xh)), w), __fadd_rd(xl, 0.99f));
          lo_next = fmaxf(__fadd_ru(__fmul_ru(0.5f, __fadd_ru(xl, xh)), w), __fadd_ru(xh, -0.99f));
        } else {
          hi = fminf(bhi, __fadd_rd(xl, 0.99f));                    // the last node answers up to bhi
          lo_next = __fadd_ru(bhi, tiny);                           // lo[n + 1]: beyond bhi
        }
      } else if (t == n + 1) {
        y = chi;
      }
    }
    out_hi[(size_t)t * C + lane] = hi;
    out_y[(size_t)t * C + lane] = y;
    if (t + 1 < R) out_lo[(size_t)(t + 1) * C + lane] = lo_next;
    if (t == 0) out_lo[lane] = usable ? -finf : finf;
  }
  if (warp == 0) reinterpret_cast<int*>(out + (size_t)3 * R * C)[lane] = n;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Best candidate of one table row for the exact 2-D nearest rule (second pass).  `sorted`: the row is NaN free
// (hence ascending): the nearest node is one of the two neighbours of a binary search; otherwise all nq nodes are
// scanned.  Tie rule of the scan in both cases: the first (lowest-index) node among equally near ones, i.e. the left
// neighbour on a distance tie and the first element of a run of duplicated nodes; strict < between rows.
__device__ __forceinline__ void nearest_in_raw_row(const float* __restrict__ gx, const float* __restrict__ gy, int nq,
                                                   bool sorted, double xd, double dg2, double& best_d2, float& best_y) {
  if (!sorted) {
    for (int k = 0; k < nq; ++k) {
      const float xv = gx[k], yv = gy[k];
      if (xv != xv || yv != yv) continue;
      const double d = fabs(xd - (double)xv);
      const double d2 = __dadd_rn(__dmul_rn(d, d), dg2);
      if (d2 < best_d2) { best_d2 = d2; best_y = yv; }
    }
    return;
  }
  const float xf = (float)xd;  // (x is a float32 sample: exact)
  int lo = 0, hi = nq;         // lower bound: first node >= x
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (gx[mid] < xf) lo = mid + 1; else hi = mid; }
  const int i = lo;
  double d = __longlong_as_double(0x7ff0000000000000LL);
  int idx = -1;
  if (i > 0) {
    const float xl = gx[i - 1];
    d = fabs(xd - (double)xl);
    idx = i - 1;
    if (i > 1 && gx[i - 2] == xl) {  // duplicated node: the scan keeps its first occurrence
      int a = 0, b = i - 1;
      while (a < b) { const int mid = (a + b) >> 1; if (gx[mid] < xl) a = mid + 1; else b = mid; }
      idx = a;
    }
  }
  if (i < nq) {
    const double dr = fabs((double)gx[i] - xd);
    if (idx < 0 || dr < d) { d = dr; idx = i; }
  }
  const double d2 = __dadd_rn(__dmul_rn(d, d), dg2);
  if (d2 < best_d2) { best_d2 = d2; best_y = gy[idx]; }
}

struct FixEntry { long long off; long long pt; float x; int g; };

// Second pass of K2t: the deferred samples (typically ~1e-4 of all) get the exact rule -- float64 distances over the
// cyclically padded rows, centre row first, then rows at distance 1, 2, ... (-, +), as lookup_2d_nearest_n -- from
// the caller's raw point-major tables: a point's row is nq contiguous floats (7 sectors), and K2p's nv says whether
// it is NaN free, i.e. searchable.  nv == nullptr: scan everything (debug / no packed image).
__global__ void adjust_fix_kernel(const FixEntry* __restrict__ fix, const unsigned* __restrict__ count, unsigned cap,
                                  const float* __restrict__ af, const float* __restrict__ hist_q, int G, int nq, int extrap,
                                  int kind, const float* __restrict__ packed, float* __restrict__ scen) {
  const unsigned n_fix = min(*count, cap);
  const size_t slot_f = packed_slot_floats(nq);
  for (unsigned e_i = blockIdx.x * blockDim.x + threadIdx.x; e_i < n_fix; e_i += gridDim.x * blockDim.x) {
    const FixEntry e = fix[e_i];
    const float fnan = Num<float>::nan();
    float f = fnan;
    if (e.x == e.x) {
      const long long base = e.pt * (long long)G * nq;
      const int lane = (int)(e.pt & 31);
      const float* tile = packed ? packed + (size_t)(e.pt >> 5) * G * slot_f + (size_t)3 * (nq + 2) * 32 : nullptr;
      auto row_sorted = [&](int gg) {
        return tile && reinterpret_cast<const int*>(tile + (size_t)gg * slot_f)[lane] == nq;
      };
      const float* xr = hist_q + base + (long long)e.g * nq;
      const float* yr = af + base + (long long)e.g * nq;
      float blo = fnan, bhi = fnan, clo = fnan, chi = fnan;
      if (row_sorted(e.g)) {
        blo = xr[0]; bhi = xr[nq - 1]; clo = yr[0]; chi = yr[nq - 1];
      } else {
        bool hb = false, hc = false;
        for (int k = 0; k < nq; ++k) {
          const float xv = xr[k], yv = yr[k];
          if (xv == xv) { if (!hb) { blo = xv; hb = true; } bhi = xv; }
          if (yv == yv) { if (!hc) { clo = yv; hc = true; } chi = yv; }
        }
      }
      const double xd = (double)e.x;
      if (xd < (double)blo) f = extrap == 0 ? clo : fnan;
      else if (xd > (double)bhi) f = extrap == 0 ? chi : fnan;
      else {
        double best_d2 = __longlong_as_double(0x7ff0000000000000LL);
        for (int dist = 0; dist <= G + 1; ++dist) {
          const double dg2 = (double)dist * (double)dist;
          if (dg2 >= best_d2) break;
          for (int sgn = -1; sgn <= 1; sgn += 2) {
            if (dist == 0 && sgn > 0) break;
            const int pr = e.g + 1 + sgn * dist;  // padded row index in [0, G+1]
            if (pr < 0 || pr > G + 1) continue;
            const int gg = (pr - 1 + G) % G;
            nearest_in_raw_row(hist_q + base + (long long)gg * nq, af + base + (long long)gg * nq, nq, row_sorted(gg), xd,
                               dg2, best_d2, f);
          }
        }
      }
    }
    scen[e.off] = kind == XSDBA_KIND_ADD ? __fadd_rn(e.x, f) : __fmul_rn(e.x, f);
  }
}

// shared-memory load by 32-bit shared address (keeps the search pointer a plain 32-bit register: LDS [R + imm])
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

constexpr int kFixStage = 64;       // deferred samples staged per CTA between two flushes (K2t)
constexpr int kTileMaxRows = 1024;  // member rows of one group kept in shared memory by K2t (x2 buffers)

// shared memory of K2t: ring of 2 slots {hi[2 TOP][32], lo[R][32], y[R][32]}, 2 row lists, mbarriers, fix staging
__host__ __device__ constexpr size_t tile_slot_floats(int top, int nq) { return (size_t)(2 * top + 2 * (nq + 2)) * 32; }

template <int TOP>
__global__ void __launch_bounds__(kThreads, 3)
adjust_tile_kernel(const float* __restrict__ sim, long long n_pts, long long sp, int st,
                   const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows, int n_groups, int nq,
                   const float* __restrict__ packed, int kind, float* __restrict__ scen,
                   FixEntry* __restrict__ fix, unsigned* __restrict__ fix_count, unsigned fix_cap) {
  constexpr int C = 32, U = 8, LD = 2 * TOP, RING = 2;  // current group + the next one in flight
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int R = nq + 2;  // rows per array in the packed image; hi rows [R, LD) of the ring keep their initial +inf
  const int slot_fl = (int)tile_slot_floats(TOP, nq);
  const uint32_t off_lo = LD * C * 4, off_y = (LD + R) * C * 4;               // byte offsets inside a ring slot (CTA-uniform)
  float* ring = reinterpret_cast<float*>(smem_raw);                            // [RING]{hi[LD][C], lo[R][C], y[R][C]}
  int* rows_sm = reinterpret_cast<int*>(ring + RING * slot_fl);                // [2][kTileMaxRows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(rows_sm + 2 * kTileMaxRows);    // [RING]
  // deferred samples are collected per CTA and appended to the global list once per group: one global atomic per
  // flush instead of one per sample (a warp otherwise waits a global round trip for its slot number)
  FixEntry* stage_fix = reinterpret_cast<FixEntry*>(bars + 4);                 // [kFixStage]
  unsigned* stage_n = reinterpret_cast<unsigned*>(stage_fix + kFixStage);      // [1] (+ [1] flush base)

  const int tile = blockIdx.x;
  const long long n0 = (long long)tile * C;
  const int G = n_groups;
  const size_t slot_f = packed_slot_floats(nq);
  const float* my = packed + (size_t)tile * G * slot_f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = n0 + lane;
  const bool pt_ok = pt < n_pts;
  const float* __restrict__ src = sim + (pt_ok ? pt : n0) * sp;
  float* __restrict__ dst = scen + (pt_ok ? pt : n0) * sp;
  const float finf = Num<float>::inf();

  for (int i = threadIdx.x; i < RING * slot_fl; i += blockDim.x) ring[i] = finf;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stage_n[0] = 0;
  }
  // (generic-proxy writes above vs. the async-proxy bulk copies below)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // the intervals of group e live in ring slot e % RING (use number e / RING of that slot's mbarrier)
  auto fetch = [&](int e) {
    const int s = e % RING;
    const uint32_t bytes = (uint32_t)R * C * 4;
    mbar_expect_tx(&bars[s], 3 * bytes);
    const float* g0 = my + (size_t)e * slot_f;
    float* s0 = ring + (size_t)s * slot_fl;
    bulk_g2s(s0, g0, bytes, &bars[s]);
    bulk_g2s(s0 + LD * C, g0 + (size_t)R * C, 2 * bytes, &bars[s]);  // lo and y are adjacent on both sides
  };
  if (threadIdx.x == 0) fetch(0);
  {
    // every row list is padded with U copies of its last row, so that a batch never needs an index clamp
    const int r0 = mem_off[0], nr = mem_off[1] - r0;
    if (nr > 0)
      for (int i = threadIdx.x; i < nr + U; i += blockDim.x) rows_sm[i] = mem_rows[r0 + min(i, nr - 1)];
  }

  for (int g = 0; g < G; ++g) {
    const int m0 = mem_off[g], n_rows = mem_off[g + 1] - m0;
    const int* rows = rows_sm + (g & 1) * kTileMaxRows;
    // rows of the next group: loads now, stores at the end of the step
    int nxt[(kTileMaxRows + kThreads - 1) / kThreads];
    int n_next = 0, r_next = 0;
    if (g + 1 < G) { r_next = mem_off[g + 1]; n_next = mem_off[g + 2] - r_next; }
#pragma unroll
    for (int i = 0; i < (kTileMaxRows + kThreads - 1) / kThreads; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      nxt[i] = idx < n_next + U && n_next > 0 ? mem_rows[r_next + min(idx, n_next - 1)] : 0;
    }
    __syncthreads();  // everyone is done with step g-1: its slot is free, rows_sm[g&1] is complete
    if (threadIdx.x == 0 && g + 1 < G) fetch(g + 1);  // streams in while this group is processed
    mbar_wait(&bars[g % RING], (uint32_t)((g / RING) & 1));

    // two register sets (A, B) alternate between "being loaded" and "being processed": the next batch of U rows is
    // always in flight while the current one is searched, and no register-to-register copies are needed
    const uint32_t xb = smem_u32(ring + (size_t)(g % RING) * slot_fl + lane);  // hi[0] of this lane's column
    const char* __restrict__ srcb = reinterpret_cast<const char*>(src);  // byte addressing: one IMAD.WIDE per sample
    char* __restrict__ dstb = reinterpret_cast<char*>(dst);
    const int st4 = st * 4;
    // (the row numbers are read from shared memory again at store time: measured faster than keeping them in 16
    //  registers next to the two value sets -- 1.72 vs 1.83 ms per slab for the whole adjust, same box.  Three CTAs per
    //  SM at 80 registers beat four at 64: 1.72 vs 1.87 ms)
    auto batch_rows = [&](int mb, int (&ov)[U]) {
      const int4 r0 = *reinterpret_cast<const int4*>(rows + mb), r1 = *reinterpret_cast<const int4*>(rows + mb + 4);
      ov[0] = r0.x; ov[1] = r0.y; ov[2] = r0.z; ov[3] = r0.w;
      ov[4] = r1.x; ov[5] = r1.y; ov[6] = r1.z; ov[7] = r1.w;
    };
    auto load_batch = [&](int mb, float (&xv)[U]) {
      int ov[U];
      batch_rows(mb, ov);
#pragma unroll
      for (int j = 0; j < U; ++j) xv[j] = *reinterpret_cast<const float*>(srcb + (long long)ov[j] * st4);
    };
    auto process_batch = [&](int mb, const float (&x)[U]) {
      uint32_t po[U];  // shared-memory address of hi[i]: the search advances it by step rows (128 B each)
#pragma unroll
      for (int j = 0; j < U; ++j) po[j] = xb;
#pragma unroll
      for (int step = TOP; step > 0; step >>= 1) {
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const float v = lds_f32(po[j] + (step - 1) * (C * 4));
          po[j] = v < x[j] ? po[j] + step * (C * 4) : po[j];
        }
      }
      int ov[U];
      batch_rows(mb, ov);
#pragma unroll
      for (int j = 0; j < U; ++j) {
        // po = address of hi[i], i = #{hi < x}: the interval of x; decided iff x >= lo[i] (a NaN sample is "decided":
        // NaN (+|*) anything = NaN)
        const float lo = lds_f32(po[j] + off_lo), f = lds_f32(po[j] + off_y);
        if (mb + j < n_rows && pt_ok) {
          *reinterpret_cast<float*>(dstb + (long long)ov[j] * st4) =
              kind == XSDBA_KIND_ADD ? __fadd_rn(x[j], f) : __fmul_rn(x[j], f);
          if (x[j] < lo) {
            // between two intervals (near-tie / far node / empty row): defer to the exact second pass
            const FixEntry en{(long long)(dst - scen) + (long long)ov[j] * st, pt, x[j], g};
            const unsigned sl = atomicAdd(stage_n, 1u);
            if (sl < kFixStage) {
              stage_fix[sl] = en;
            } else {  // staging list full until the next flush: straight to the global list
              const unsigned slot = atomicAdd(fix_count, 1u);
              if (slot < fix_cap) fix[slot] = en;
            }
          }
        }
      }
    };
    {
      float xa[U], xbv[U];
      const int stride = n_warps * U;
      int m = warp * U;
      if (m < n_rows) load_batch(m, xa);
      while (m < n_rows) {
        if (m + stride < n_rows) load_batch(m + stride, xbv);
        process_batch(m, xa);
        m += stride;
        if (m >= n_rows) break;
        if (m + stride < n_rows) load_batch(m + stride, xa);
        process_batch(m, xbv);
        m += stride;
      }
    }
    __syncthreads();  // every sample of this group has been looked at: flush the staged deferred samples
    {
      const unsigned ns = min(stage_n[0], (unsigned)kFixStage);
      if (ns > 0) {  // (CTA-uniform)
        if (threadIdx.x == 0) stage_n[1] = atomicAdd(fix_count, ns);
        __syncthreads();  // the base is visible, and everybody has read ns
        const unsigned base = stage_n[1];
        for (unsigned i = threadIdx.x; i < ns; i += blockDim.x)
          if (base + i < fix_cap) fix[base + i] = stage_fix[i];
        if (threadIdx.x == 0) stage_n[0] = 0;  // (the barrier at the top of the next group orders this reset)
      }
    }
    {
      int* rn = rows_sm + ((g + 1) & 1) * kTileMaxRows;
#pragma unroll
      for (int i = 0; i < (kTileMaxRows + kThreads - 1) / kThreads; ++i) {
        const int idx = threadIdx.x + i * kThreads;
        if (idx < n_next + U) rn[idx] = nxt[i];
      }
    }
  }
}

// =============================================================================================
// K4: per-(point, group) polynomial trend  (PolyDetrend, detrending.py:165-208 via map_groups).
// grid = (ceil(n_pts/32), n_groups), 256 threads.  y = x (+|*) scaling[point][group] (nullable), window
// slots averaged NaN-skipping first (detrending.py:199-200), least squares of degree <= 4 on the
// normalised time coordinate u = (t - t_first)/(t_last - t_first) by float64 normal equations, then
// evaluated at every member: trend (float64) has the strides of x.
// =============================================================================================
constexpr int kMaxDeg = 4;

template <typename T>
__global__ void __launch_bounds__(kThreads)
poly_trend_kernel(const T* __restrict__ x, long long n_pts, long long sp, long long st,
                  const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows,
                  const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows,
                  const int32_t* __restrict__ gidx, int n_groups, int window, const T* __restrict__ scaling, int kind, int degree, const double* __restrict__ tcoord,
                  double* __restrict__ trend) {
  constexpr int C = 32, NS = 2 * kMaxDeg + 1, NB = kMaxDeg + 1;
  __shared__ double red[kThreads / 32][NS + NB][C];
  __shared__ double coef[NB][C];
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], n_mem = mem_off[g + 1] - m0;
  if (n_mem == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = n0 + lane;
  const bool ok = pt < n_pts;
  const int32_t* segs = seg_rows + seg_off[g];
  const double t0 = tcoord[mem_rows[m0]], t1 = tcoord[mem_rows[m0 + n_mem - 1]];
  const double inv = t1 > t0 ? 1.0 / (t1 - t0) : 1.0;
  T sc = (T)0;
  if (scaling && ok) sc = scaling[pt * n_groups + g];

  double S[NS], B[NB];
#pragma unroll
  for (int k = 0; k < NS; ++k) S[k] = 0.0;
#pragma unroll
  for (int k = 0; k < NB; ++k) B[k] = 0.0;
  for (int m = warp; m < n_mem; m += n_warps) {
    double y;
    if (window == 1) {
      T v = ok ? x[pt * sp + (long long)mem_rows[m0 + m] * st] : Num<T>::nan();
      if (scaling) v = apply_corr<T>(v, sc, kind);
      y = (double)v;
    } else {
      double acc = 0.0; int cnt = 0;
      for (int j = 0; j < window; ++j) {
        const int t = segs[m * window + j];
        if (t < 0 || !ok) continue;
        T v = x[pt * sp + (long long)t * st];
        // every time step is scaled by its OWN group's factor before the window is built (_adjustment.py:748-757)
        if (scaling) v = gidx[t] >= 0 ? apply_corr<T>(v, scaling[pt * n_groups + gidx[t]], kind) : Num<T>::nan();
        if (!is_nan(v)) { acc += (double)v; ++cnt; }
      }
      y = cnt > 0 ? acc / (double)cnt : Num<double>::nan();
    }
    if (y == y) {
      const double u = (tcoord[mem_rows[m0 + m]] - t0) * inv;
      double p = 1.0;
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (k <= 2 * degree) S[k] += p;
        if (k <= degree) B[k < NB ? k : 0] += p * y;
        p *= u;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) red[warp][k][lane] = S[k];
#pragma unroll
  for (int k = 0; k < NB; ++k) red[warp][NS + k][lane] = B[k];
  __syncthreads();
  if (warp == 0) {
    double A[NB][NB + 1];
    double Ssum[NS], Bsum[NB];
    for (int k = 0; k < NS; ++k) { double a = 0; for (int w = 0; w < n_warps; ++w) a += red[w][k][lane]; Ssum[k] = a; }
    for (int k = 0; k < NB; ++k) { double a = 0; for (int w = 0; w < n_warps; ++w) a += red[w][NS + k][lane]; Bsum[k] = a; }
    const int nd = degree + 1;
    for (int i = 0; i < nd; ++i) { for (int j = 0; j < nd; ++j) A[i][j] = Ssum[i + j]; A[i][nd] = Bsum[i]; }
    bool singular = Ssum[0] <= (double)degree;  // not more valid points than the degree
    for (int i = 0; i < nd && !singular; ++i) {  // Gaussian elimination, partial pivoting
      int piv = i;
      for (int r = i + 1; r < nd; ++r) if (fabs(A[r][i]) > fabs(A[piv][i])) piv = r;
      if (A[piv][i] == 0.0) { singular = true; break; }
      if (piv != i) for (int j = 0; j <= nd; ++j) { const double tmp = A[i][j]; A[i][j] = A[piv][j]; A[piv][j] = tmp; }
      for (int r = i + 1; r < nd; ++r) {
        const double f = A[r][i] / A[i][i];
        for (int j = i; j <= nd; ++j) A[r][j] -= f * A[i][j];
      }
    }
    for (int i = nd - 1; i >= 0; --i) {
      double v = A[i][nd];
      for (int j = i + 1; j < nd; ++j) v -= A[i][j] * coef[j][lane];
      coef[i][lane] = singular ? Num<double>::nan() : v / A[i][i];
    }
  }
  __syncthreads();
  if (!ok) return;
  for (int m = warp; m < n_mem; m += n_warps) {
    const int t = mem_rows[m0 + m];
    const double u = (tcoord[t] - t0) * inv;
    double v = coef[degree][lane];
    for (int k = degree - 1; k >= 0; --k) v = v * u + coef[k][lane];
    trend[pt * sp + (long long)t * st] = v;
  }
}

// =============================================================================================
// K5: DQM adjust (dqm_adjust.func, _adjustment.py:748-780) given the trend: per sample, in float64 like
// the reference after detrending,  x = sim (+|*) scaling ; xd = x (+|*) invert(trend) ; f = lookup(xd) ;
// scen = (xd (+|*) f) (+|*) trend.  Same tiling / staging as K2 (generic kernel).
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
dqm_adjust_kernel(const T* __restrict__ sim, long long n_pts, long long sp, long long st,
                  const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows, int n_groups,
                  const T* __restrict__ af, const T* __restrict__ hist_q, const T* __restrict__ scaling,
                  const double* __restrict__ trend, int nq, int interp, int extrap, int kind, T* __restrict__ scen) {
  constexpr int C = 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Tables<T, C> tb = carve_tables<T, C>(smem_raw, nq);
  T* stage = reinterpret_cast<T*>(smem_raw + ((tables_bytes<T, C>(nq) + 15) & ~(size_t)15));
  tb.gx = hist_q; tb.gy = af; tb.x_shared = false; tb.G = n_groups; tb.pt_stride = (long long)n_groups * nq;
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], m1 = mem_off[g + 1];
  if (m0 == m1) return;
  const bool grouped = n_groups > 1;
  stage_tables<T, C>(tb, stage, n0, n_pts, g, grouped);
  if (interp == XSDBA_INTERP_CUBIC && !grouped) stage_cubic<T, C>(tb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = n0 + lane;
  if (pt >= n_pts) return;
  const T sc = scaling[pt * n_groups + g];
  for (int m = m0 + warp; m < m1; m += n_warps) {
    const long long o = pt * sp + (long long)mem_rows[m] * st;
    const T xs = apply_corr<T>(sim[o], sc, kind);           // scaled_sim, data dtype (_adjustment.py:748-757)
    const double tr = trend[o];
    const double itr = kind == XSDBA_KIND_ADD ? -tr : __ddiv_rn(1.0, tr);   // utils.invert
    const double xd = apply_corr<double>((double)xs, itr, kind);             // detrended, float64
    const T f = grouped ? lookup_2d_nearest<double, T, C>(tb, lane, pt, g, xd, extrap)
                        : lookup_1d<double, T, C>(tb, lane, xd, interp, extrap);
    const double sd = apply_corr<double>(xd, (double)f, kind);
    scen[o] = (T)apply_corr<double>(sd, tr, kind);
  }
}

// =============================================================================================
// K6: LOESS trend of a whole series (LoessDetrend(group="time"), detrending.py:211-296 ->
// loess.loess_smoothing / _loess_nb, loess.py:49-179, 244-278), equal-spacing branch, tricube
// weights, niter = 1, local degree d in {0, 1}.
//   K6a loess_compact_kernel: one thread per point walks its series, applies the optional per-group
//       scaling, drops NaNs (loess.py:94-100) and writes the compacted values / time indices
//       time-major into the workspace (lanes = neighbouring points, so rows stay coalesced).
//   K6b loess_smooth_kernel: thread = (point, output index i); the window, bandwidth h and the
//       "weights are only recomputed near the edges" rule follow loess.py:122-150 literally: away from
//       the edges the weights are those computed at i = HW from the first 2*HW+1 valid samples.
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
loess_compact_kernel(const T* __restrict__ x, long long n_pts, long long sp, long long st, int n_time,
                     const int32_t* __restrict__ gidx, int n_groups, const T* __restrict__ scaling, int kind,
                     T* __restrict__ yc, int32_t* __restrict__ tc, int32_t* __restrict__ nvalid,
                     double* __restrict__ trend) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt >= n_pts) return;
  int n = 0;
  for (int t = 0; t < n_time; ++t) {
    T v = x[pt * sp + (long long)t * st];
    if (scaling) v = gidx[t] >= 0 ? apply_corr<T>(v, scaling[pt * n_groups + gidx[t]], kind) : Num<T>::nan();
    trend[pt * sp + (long long)t * st] = Num<double>::nan();
    if (!is_nan(v)) {
      yc[(long long)n * n_pts + pt] = v;
      tc[(long long)n * n_pts + pt] = t;
      ++n;
    }
  }
  nvalid[pt] = n;
}

__device__ __forceinline__ double tricube_w(double u) {  // loess.py:31-35
  const double c = 1.0 - u * u * u;
  return u >= 1.0 ? 0.0 : c * c * c;
}
__device__ __forceinline__ double gaussian_w(double u) {  // loess.py:16-26: the span covers 95 % of the gaussian
  return u >= 1.0 ? 0.0 : exp(-(u * u) / (2.0 * ((1.0 / 1.96) * (1.0 / 1.96))));
}

// LOESS window geometry of one compacted series of n samples (loess.py:104-119)
// (the weight function travels in the sign of f inside the library: f < 0 = gaussian weights on the span |f|)
struct LoessGeom {
  int r, hw, R, HW;
  bool gauss;
  __device__ LoessGeom(int n, double f_signed) {
    gauss = f_signed < 0.0;
    const double f = fabs(f_signed);
    r = (int)(2.0 * floor(f * (double)n / 2.0) + 1.0);
    hw = (r - 1) / 2;
    R = r + 4 < n ? r + 4 : n;
    HW = hw + 2;
  }
  // true when output i uses the weights computed at i = HW on the window [i-HW, i+HW] (loess.py:131-150)
  __device__ bool interior(int i, int n) const { return i > HW && i < n - HW - 1; }
  __device__ double w(double u) const { return gauss ? gaussian_w(u) : tricube_w(u); }
};

// K6c: the "interior" weights of every point: w[k] = tricube(|x[k] - x[HW]| / ((hw+1) dx)), k in [0, 2HW],
// written k-major ([k][point]) so that the smoothing kernel reads them coalesced.
__global__ void __launch_bounds__(kThreads)
loess_weights_kernel(const int32_t* __restrict__ tc, const int32_t* __restrict__ nvalid, long long n_pts, int n_time,
                     const double* __restrict__ xn, double f, int w_rows, double* __restrict__ wtab) {
  const long long pt = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
  if (pt >= n_pts) return;
  const int n = nvalid[pt];
  const LoessGeom gm(n, f);
  if (n < 2 * gm.HW + 3) return;  // no interior outputs for this point
  const double dx = xn[1] - xn[0];
  const double h = (double)(gm.hw + 1) * dx;
  const double xc = xn[tc[(long long)gm.HW * n_pts + pt]];
  for (int k = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); k < w_rows; k += gridDim.y * (blockDim.x >> 5))
    wtab[(long long)k * n_pts + pt] =
        k <= 2 * gm.HW ? gm.w(fabs(xn[tc[(long long)k * n_pts + pt]] - xc) / h) : 0.0;  // zero tail: K6 reads past K
}

// total interior weight of every point, summed in tap order like the reference's w.sum() (loess.py:38-39): all
// interior outputs of a point share it.  Row w_rows of wtab; the complete-series value goes to wsh[w_rows].
__global__ void loess_wsum_kernel(const int32_t* __restrict__ nvalid, long long n_pts, int n_time, double f, int w_rows,
                                  double* __restrict__ wtab, double* __restrict__ wsh) {
  const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pt == n_pts && wsh) {  // one extra thread sums the shared vector
    double s = 0.0;
    const LoessGeom gm(n_time, f);
    for (int k = 0; k <= 2 * gm.HW && k < w_rows; ++k) s += wsh[k];
    wsh[w_rows] = s;
  }
  if (pt >= n_pts) return;
  const LoessGeom gm(nvalid[pt], f);
  double s = 0.0;
  if (nvalid[pt] >= 2 * gm.HW + 3)
    for (int k = 0; k <= 2 * gm.HW && k < w_rows; ++k) s += wtab[(long long)k * n_pts + pt];
  wtab[(long long)w_rows * n_pts + pt] = s;
}

// interior weights of a COMPLETE series (n == n_time): the same for every such point, so a warp whose 32 points are
// all complete reads one broadcast value per tap instead of 32 per-point ones
__global__ void loess_shared_weights_kernel(const double* __restrict__ xn, int n, double f, int w_rows,
                                            double* __restrict__ wsh) {
  const LoessGeom gm(n, f);
  const double dx = xn[1] - xn[0];
  const double h = (double)(gm.hw + 1) * dx;
  const double xc = xn[gm.HW];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < w_rows; k += gridDim.x * blockDim.x)
    wsh[k] = (k <= 2 * gm.HW && n >= 2 * gm.HW + 3) ? gm.w(fabs(xn[k] - xc) / h) : 0.0;
}

// one output by the literal rule (edges, short series)
template <typename T>
__device__ double loess_one(const T* __restrict__ y, const int32_t* __restrict__ tcp, long long n_pts, int n, int i,
                            const LoessGeom& gm, const double* __restrict__ xn, double dx, int degree,
                            const double* __restrict__ delta = nullptr) {
  int lo, hi;                                                 // loess.py:124-135
  if (i < gm.HW) { lo = 0; hi = gm.R; }
  else if (i >= n - gm.HW - 1) { lo = n - gm.R; hi = n; }
  else { lo = i - gm.HW; hi = i + gm.HW + 1; }
  // weights: recomputed for i <= HW or i >= n-HW, otherwise those of the last recomputation (i = HW),
  // i.e. taken from the samples [0, 2HW+1) around sample HW                    (loess.py:136-150)
  const bool edge = (i <= gm.HW) || (i >= n - gm.HW);
  const int ic = edge ? i : gm.HW;
  const int wlo = edge ? lo : 0;
  double h;
  if (ic < gm.hw) h = (double)(gm.r - ic) * dx;
  else if (ic >= n - gm.hw) h = (double)(ic - (n - gm.r) + 1) * dx;
  else h = (double)(gm.hw + 1) * dx;
  const double xc = xn[tcp[(long long)ic * n_pts]];
  const double xi = xn[tcp[(long long)i * n_pts]];
  double sw = 0, swy = 0, swx = 0, swxx = 0, swxy = 0;
  for (int j = lo; j < hi; ++j) {
    const int k = wlo + (j - lo);
    double w = k < n ? gm.w(fabs(xn[tcp[(long long)k * n_pts]] - xc) / h) : 0.0;
    if (delta) w = delta[(long long)j * n_pts] * w;            // robustness weights: w = di * wi (loess.py:150)
    const double yj = (double)y[(long long)j * n_pts];
    sw += w; swy += w * yj;
    if (degree == 1) {
      const double xj = xn[tcp[(long long)j * n_pts]];
      swx += w * xj; swxx += w * xj * xj; swxy += w * yj * xj;
    }
  }
  if (degree == 0) return swy / sw;                           // loess.py:38-39
  double a00 = sw, a01 = swx, a10 = swx, a11 = swxx, b0 = swy, b1 = swxy;  // loess.py:42-46
  if (fabs(a10) > fabs(a00)) { double t_; t_ = a00; a00 = a10; a10 = t_; t_ = a01; a01 = a11; a11 = t_; t_ = b0; b0 = b1; b1 = t_; }
  const double m = a10 / a00;
  a11 -= m * a01; b1 -= m * b0;
  const double beta1 = b1 / a11;
  const double beta0 = (b0 - a01 * beta1) / a00;
  return beta0 + beta1 * xi;
}

// The unequal-spacing branch of _loess_nb (dx == 0, loess.py:107-111, 151-158): r = round(f n) (half to even), the
// window of output i holds at most 2 (r + 2) samples around it, the bandwidth h is the distance of the r-th closest
// sample (np.sort(diffs)[r], diffs[0] = 0 is the sample itself) -- found by walking outwards from i on the sorted
// abscissa instead of sorting -- and the weights are recomputed for every output and iteration.
template <typename T>
__device__ double loess_one_unequal(const T* __restrict__ y, const int32_t* __restrict__ tcp, long long n_pts, int n, int i,
                                    int r, bool gauss, const double* __restrict__ xn, int degree,
                                    const double* __restrict__ delta) {
  const int HW = min(r + 2, n), R = min(2 * HW, n);
  int lo, hi;                                                 // loess.py:124-135
  if (i < HW) { lo = 0; hi = R; }
  else if (i >= n - HW - 1) { lo = n - R; hi = n; }
  else { lo = i - HW; hi = i + HW + 1; }
  if (r >= hi - lo) return __longlong_as_double(0x7ff8000000000000LL);   // (np.sort(diffs)[r] does not exist)
  const double xi = xn[tcp[(long long)i * n_pts]];
  double h = 0.0;
  {
    int l = i - 1, u = i + 1;
    for (int s_ = 0; s_ < r; ++s_) {
      const double dl = l >= lo ? xi - xn[tcp[(long long)l * n_pts]] : __longlong_as_double(0x7ff0000000000000LL);
      const double du = u < hi ? xn[tcp[(long long)u * n_pts]] - xi : __longlong_as_double(0x7ff0000000000000LL);
      if (dl <= du) { h = dl; --l; } else { h = du; ++u; }
    }
  }
  double sw = 0, swy = 0, swx = 0, swxx = 0, swxy = 0;
  for (int j = lo; j < hi; ++j) {
    const double xj = xn[tcp[(long long)j * n_pts]];
    const double uu = fabs(xj - xi) / h;
    double w = gauss ? gaussian_w(uu) : tricube_w(uu);
    if (delta) w = delta[(long long)j * n_pts] * w;            // w = di * weight_func(diffs / h)  (loess.py:158)
    const double yj = (double)y[(long long)j * n_pts];
    sw += w; swy += w * yj;
    if (degree == 1) { swx += w * xj; swxx += w * xj * xj; swxy += w * yj * xj; }
  }
  if (degree == 0) return swy / sw;                           // loess.py:38-39
  double a00 = sw, a01 = swx, a10 = swx, a11 = swxx, b0 = swy, b1 = swxy;  // loess.py:42-46
  if (fabs(a10) > fabs(a00)) { double t_; t_ = a00; a00 = a10; a10 = t_; t_ = a01; a01 = a11; a11 = t_; t_ = b0; b0 = b1; b1 = t_; }
  const double m = a10 / a00;
  a11 -= m * a01; b1 -= m * b0;
  const double beta1 = b1 / a11;
  const double beta0 = (b0 - a01 * beta1) / a00;
  return beta0 + beta1 * xi;
}

// K6n: the unequal-spacing branch, thread = (point, output), lane = point
template <typename T>
__global__ void __launch_bounds__(kThreads)
loess_unequal_kernel(const T* __restrict__ yc, const int32_t* __restrict__ tc, const int32_t* __restrict__ nvalid,
                     long long n_pts, long long sp, long long st, const double* __restrict__ xn, double f_signed,
                     int degree, const double* __restrict__ delta, double* __restrict__ trend) {
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5, rows_per_cta = blockDim.x >> 5;
  const long long pt = (long long)blockIdx.x * 32 + lane;
  if (pt >= n_pts) return;
  const int n = nvalid[pt];
  if (n == 0) return;
  const int r = (int)rint(fabs(f_signed) * (double)n);         // np.round: half to even
  for (int i = blockIdx.y * rows_per_cta + row; i < n; i += gridDim.y * rows_per_cta)
    trend[pt * sp + (long long)tc[(long long)i * n_pts + pt] * st] =
        loess_one_unequal<T>(yc + pt, tc + pt, n_pts, n, i, r, f_signed < 0.0, xn, degree, delta ? delta + pt : nullptr);
}

// K6b: thread = (point, chunk of RO consecutive outputs).  A chunk that is interior for the thread's series
// runs as a register-tiled FIR on the precomputed weights (every y / w value is loaded once per chunk and
// feeds RO accumulators); edge chunks and short series use the literal per-output rule.
// K6e: edge weights shared by every point WITHOUT missing values (n == n_time: the compacted axis is the full axis,
// so the weights of the first HW+1 and the last HW outputs depend on (output, tap) only, loess.py:138-147).  One
// table per call, tap-major: etab[j * NI + e] with e = i for the left outputs, HW+1 + (i - (n-HW)) for the right
// ones, j relative to the shared window [0, R) / [n-R, n); esum[e] = the weights summed in tap order.  38 MB for
// 30 years at f = 0.2 -- L2 resident -- against 10 float64 operations per (tap, output) to recompute them.
__global__ void __launch_bounds__(kThreads)
loess_edge_table_kernel(const double* __restrict__ xn, int n, double f, double* __restrict__ etab) {
  const LoessGeom gm(n, f);
  const int NI = 2 * gm.HW + 1;
  const double dx = xn[1] - xn[0];
  const long long total = (long long)gm.R * NI;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(idx % NI), j = (int)(idx / NI);
    const bool left = e <= gm.HW;
    const int i = left ? e : n - gm.HW + (e - gm.HW - 1);
    const int lo = left ? 0 : n - gm.R;
    double h;
    if (i < gm.hw) h = (double)(gm.r - i) * dx;
    else if (i >= n - gm.hw) h = (double)(i - (n - gm.r) + 1) * dx;
    else h = (double)(gm.hw + 1) * dx;
    // (the gaussian weight jumps from 0.146 to 0 at u = 1: u must be the reference's quotient there, loess.py:153)
    etab[idx] = gm.gauss ? gm.w(fabs(xn[lo + j] - xn[i]) / h) : gm.w(fabs(xn[lo + j] - xn[i]) * (1.0 / h));
  }
}
__global__ void loess_edge_sum_kernel(int n, double f, const double* __restrict__ etab, double* __restrict__ esum) {
  const LoessGeom gm(n, f);
  const int NI = 2 * gm.HW + 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NI) return;
  double s = 0.0;
  for (int j = 0; j < gm.R; ++j) s += etab[(long long)j * NI + e];
  esum[e] = s;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
loess_smooth_kernel(const T* __restrict__ yc, const int32_t* __restrict__ tc, const int32_t* __restrict__ nvalid,
                    long long n_pts, long long sp, long long st, int n_time, const double* __restrict__ xn,
                    double f, int degree, const double* __restrict__ wtab, int w_rows, const double* __restrict__ wsh,
                    const double* __restrict__ etab, const double* __restrict__ esum, const double* __restrict__ delta,
                    int tiled, double* __restrict__ trend) {
  constexpr int RO = 8;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5, rows_per_cta = blockDim.x >> 5;
  const long long pt = (long long)blockIdx.x * 32 + lane;
  // a tile of 32 complete series has its interior outputs computed by K6t (loess_interior_tile_kernel): same vote there
  const bool tile_done = __all_sync(0xffffffffu, pt < n_pts && nvalid[pt < n_pts ? pt : 0] == n_time) && tiled;
  if (tile_done) return;  // K6t + K6u cover every output of a complete tile
  if (pt >= n_pts) return;
  const int n = nvalid[pt];
  if (n == 0) return;
  const double dx = xn[1] - xn[0];                              // loess.py:262-263
  const LoessGeom gm(n, f);
  const T* y = yc + pt;
  const int32_t* tcp = tc + pt;
  const double* w = wtab + pt;
  // (lanes past n_pts / empty columns have left: vote among the remaining ones)
  const bool warp_complete = wsh != nullptr && __all_sync(__activemask(), n == n_time);
  for (int i0 = (blockIdx.y * rows_per_cta + row) * RO; i0 < n; i0 += gridDim.y * rows_per_cta * RO) {
    // robustness iterations (delta != null, niter > 1) take the plain per-output path: every weight is then di * wi
    const bool fast = !delta && degree == 0 && gm.interior(i0, n) && gm.interior(i0 + RO - 1, n);
    if (fast) {
      // output i0+r sums w[k] * y[i0 + r - HW + k], k = 0..2HW.  With j = i0 - HW + m the pair (m, r) uses w[m - r].
      // Blocks of RO taps: the loads of block b+1 are issued before block b is consumed (the loop is otherwise a
      // chain load -> fma -> next load, one memory latency per tap).  The weights sit in a register ring: step u
      // overwrites slot u, so w[m - r] is slot (u - r) mod RO, a compile-time index.  Indices past the last tap are
      // clamped: the weight rows beyond K are zero, so those products vanish.  The denominator is the point's
      // total weight (loess_wsum_kernel).
      const int K = 2 * gm.HW;  // last weight index
      const T* yb = y + (long long)(i0 - gm.HW) * n_pts;
      const double* wp = warp_complete ? wsh : w;
      const long long wstep = warp_complete ? 1 : n_pts;
      const double sw_total = warp_complete ? wsh[w_rows] : wtab[(long long)w_rows * n_pts + pt];
      const int m_last = K + RO - 1;
      double swy[RO], wr[RO], wn[RO], yn[RO];
#pragma unroll
      for (int r = 0; r < RO; ++r) {
        swy[r] = 0; wr[r] = 0;
        wn[r] = wp[(long long)min(r, w_rows - 1) * wstep];
        yn[r] = (double)yb[(long long)min(r, m_last) * n_pts];
      }
      for (int mb = 0; mb <= m_last; mb += RO) {
        double wc[RO], yc[RO];
#pragma unroll
        for (int u = 0; u < RO; ++u) { wc[u] = wn[u]; yc[u] = yn[u]; }
#pragma unroll
        for (int u = 0; u < RO; ++u) {
          wn[u] = wp[(long long)min(mb + RO + u, w_rows - 1) * wstep];
          yn[u] = (double)yb[(long long)min(mb + RO + u, m_last) * n_pts];
        }
#pragma unroll
        for (int u = 0; u < RO; ++u) {
          wr[u] = wc[u];
#pragma unroll
          for (int r = 0; r < RO; ++r) swy[r] = fma(wr[(u - r + RO) % RO], yc[u], swy[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RO; ++r)
        trend[pt * sp + (long long)tcp[(long long)(i0 + r) * n_pts] * st] = swy[r] / sw_total;
    } else if (!delta && degree == 0 && n >= gm.R && i0 + RO <= n &&
               ((i0 + RO - 1 <= gm.HW) || (i0 >= n - gm.HW))) {
      // edge chunk: all RO outputs share the window [0, R) (left) or [n-R, n) (right) and recompute their weights
      // (loess.py:138-147): every y_j / x_j is loaded once and feeds RO (weight, sum) pairs.
      const int lo = (i0 + RO - 1 <= gm.HW) ? 0 : n - gm.R;
      if (etab && n == n_time) {  // complete series: weights and their sums come from the shared table (K6e)
        const int NI = 2 * gm.HW + 1;
        const int e0 = lo == 0 ? i0 : gm.HW + 1 + (i0 - (n - gm.HW));
        const double* et = etab + e0;
        const T* yl = y + (long long)lo * n_pts;
        double swy[RO];
#pragma unroll
        for (int r = 0; r < RO; ++r) swy[r] = 0;
        for (int j = 0; j < gm.R; ++j) {
          const double yj = (double)yl[(long long)j * n_pts];
#pragma unroll
          for (int r = 0; r < RO; ++r) swy[r] = fma(et[r], yj, swy[r]);
          et += NI;
        }
#pragma unroll
        for (int r = 0; r < RO; ++r)
          trend[pt * sp + (long long)tcp[(long long)(i0 + r) * n_pts] * st] = swy[r] / esum[e0 + r];
        continue;
      }
      double xi[RO], ih[RO], sw[RO], swy[RO];
#pragma unroll
      for (int r = 0; r < RO; ++r) {
        const int i = i0 + r;
        double h;
        if (i < gm.hw) h = (double)(gm.r - i) * dx;
        else if (i >= n - gm.hw) h = (double)(i - (n - gm.r) + 1) * dx;
        else h = (double)(gm.hw + 1) * dx;
        ih[r] = gm.gauss ? h : 1.0 / h;  // (gaussian weights jump at u = 1: u must be the reference's quotient, loess.py:153)
        xi[r] = xn[tcp[(long long)i * n_pts]];
        sw[r] = 0; swy[r] = 0;
      }
      for (int j = lo; j < lo + gm.R; ++j) {
        const double xj = xn[tcp[(long long)j * n_pts]];
        const double yj = (double)y[(long long)j * n_pts];
#pragma unroll
        for (int r = 0; r < RO; ++r) {
          const double wgt = gm.gauss ? gm.w(fabs(xj - xi[r]) / ih[r]) : gm.w(fabs(xj - xi[r]) * ih[r]);
          sw[r] += wgt; swy[r] = fma(wgt, yj, swy[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RO; ++r)
        trend[pt * sp + (long long)tcp[(long long)(i0 + r) * n_pts] * st] = swy[r] / sw[r];
    } else {
      for (int r = 0; r < RO && i0 + r < n; ++r)
        trend[pt * sp + (long long)tcp[(long long)(i0 + r) * n_pts] * st] =
            loess_one<T>(y, tcp, n_pts, n, i0 + r, gm, xn, dx, degree, delta ? delta + pt : nullptr);
    }
  }
}

// =============================================================================================
// K3: per-group percentile ranks (+ QDM factor lookup).  grid = (ceil(n_pts / C), n_groups).
// The segment (exact members, or members x window when rank_window) of C points is sorted in shared
// memory; every member then finds its average-tie rank by two binary searches in its sorted column.
//   r = avgrank / n_valid ; sim_q = mx * (r - mn) / (mx - mn)         (utils.py:629-634)
// With do_adjust: af lookup on the shared quantile axis and scen = sim (+|*) af (_adjustment.py:873-881).
// =============================================================================================
template <typename T, int C>
__global__ void __launch_bounds__(threads_for_cols<C>())
rank_kernel(const T* __restrict__ sim, long long n_pts, long long sp, long long st,
            const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows,
            const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows, int n_groups,
            const T* __restrict__ af, const T* __restrict__ q, int nq, int interp, int extrap, int kind,
            int do_adjust, T* __restrict__ scen, double* __restrict__ sim_q, int n_pad, int rank_mode,
            const double* __restrict__ gcoord, const unsigned char* __restrict__ diag, int slots) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* mnmx = reinterpret_cast<double*>(smem_raw);      // [2][C]
  double* sum = mnmx + 2 * C;                              // [C] (unused sum slot of count_columns)
  int* cnt = reinterpret_cast<int*>(sum + C);              // [C]
  unsigned char* p = smem_raw + C * 28 + ((C * 28) % 8 ? 4 : 0);
  T* sm = reinterpret_cast<T*>(p);                         // [n_pad][C]
  Tables<T, C> tb = carve_tables<T, C>(p + (size_t)n_pad * C * sizeof(T), nq, slots);
  T* stage = reinterpret_cast<T*>(p + (size_t)n_pad * C * sizeof(T) + ((tables_bytes<T, C>(nq, slots) + 15) & ~(size_t)15));

  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], m1 = mem_off[g + 1];
  if (m0 == m1) return;
  const int S = seg_off[g + 1] - seg_off[g];
  const bool grouped = n_groups > 1;

  load_segment<T, C>(sm, sim, n0, n_pts, sp, st, seg_rows + seg_off[g], S, n_pad);
  __syncthreads();
  count_columns<T, C>(sm, n_pad, cnt, sum, false);
  make_keys<T, C>(sm, n_pad, cnt, sum, 0, kind);
  sort_columns<T, C>(sm, n_pad);
  if (do_adjust) {
    tb.gx = q; tb.gy = af; tb.x_shared = true; tb.G = n_groups; tb.pt_stride = (long long)n_groups * nq;
    stage_tables<T, C>(tb, stage, n0, n_pts, g, grouped);
    if (interp == XSDBA_INTERP_CUBIC && !grouped) stage_cubic<T, C>(tb);
  }
  if (threadIdx.x < C) {
    const int c = threadIdx.x, n = cnt[c];
    double mn = __longlong_as_double(0x7ff8000000000000LL), mx = mn;
    if (n > 0) {
      const T* col = sm + c;
      const T vmin = col[0], vmax = col[(size_t)(n - 1) * C];
      int ub = 1; while (ub < n && col[(size_t)ub * C] == vmin) ++ub;          // multiplicity of the minimum
      int lb = n - 1; while (lb > 0 && col[(size_t)(lb - 1) * C] == vmax) --lb;  // first index of the maximum
      // rank_mode 0: utils.rank(pct=True): r = avgrank / n_valid        (utils.py:629-634)
      // rank_mode 1: utils._rank_bn:          r = avgrank / max(avgrank)  (utils.py:641-646)
      const double den = rank_mode == 0 ? (double)n : (double)(lb + n + 1) * 0.5;
      mn = ((double)(ub + 1) * 0.5) / den;                // avg rank of ranks 1..ub
      mx = ((double)(lb + n + 1) * 0.5) / den;            // avg rank of ranks lb+1..n
    }
    mnmx[c] = mn; mnmx[C + c] = mx;
    if (n > 0) {  // (re-derive the maximum average rank for _rank_bn; sum[] is free after count_columns)
      const T* col = sm + c;
      const T vmax = col[(size_t)(n - 1) * C];
      int lb = n - 1; while (lb > 0 && col[(size_t)(lb - 1) * C] == vmax) --lb;
      sum[c] = (double)(lb + n + 1) * 0.5;
    }
  }
  __syncthreads();

  // members: item -> (member m, point c), c fastest (coalesced for time-major)
  const int n_mem = m1 - m0;
  for (int item = threadIdx.x; item < n_mem * C; item += blockDim.x) {
    const int c = item % C;
    const long long pt = n0 + c;
    if (pt >= n_pts) continue;
    const long long o = pt * sp + (long long)mem_rows[m0 + item / C] * st;
    const T x = sim[o];
    double sq = __longlong_as_double(0x7ff8000000000000LL);
    const int n = cnt[c];
    if (!is_nan(x) && n > 0) {
      const T* col = sm + c;
      int lo = 0, hi = n;  // lower bound: #values < x
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] < x) lo = mid + 1; else hi = mid; }
      const int lb = lo;
      hi = n;              // upper bound: #values <= x
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] <= x) lo = mid + 1; else hi = mid; }
      const int ub = lo;
      const double mn = mnmx[c], mx = mnmx[C + c];
      if (rank_mode == 0) {
        const double r = ((double)(lb + ub + 1) * 0.5) / (double)n;
        sq = __ddiv_rn(__dmul_rn(mx, __dsub_rn(r, mn)), __dsub_rn(mx, mn));
      } else {  // _rank_bn: rnk / nanmax(rnk), then 1 * (rnk - mn) / (1 - mn)
        const double r = ((double)(lb + ub + 1) * 0.5) / sum[c];
        sq = __ddiv_rn(__dsub_rn(r, mn), __dsub_rn(1.0, mn));
      }
    }
    if (sim_q) sim_q[o] = sq;
    if (do_adjust) {
      T f;
      if (!grouped) f = lookup_1d<double, T, C>(tb, c, sq, interp, extrap);
      else if (interp == XSDBA_INTERP_LINEAR)
        f = lookup_2d_linear_shared<T, C>(tb, c, g, sq, gcoord[mem_rows[m0 + item / C]], diag, extrap);
      else f = lookup_2d_nearest<double, T, C>(tb, c, pt, g, sq, extrap);
      scen[o] = apply_corr<T>(x, f, kind);
    }
  }
}

// =============================================================================================
// Microbenchmark (not on the product path): copy the member rows of every group with the same
// (point tile x group) decomposition as K2f, V floats per lane (tile = 32*V points, 128*V-byte row
// pieces).  profiles/ uses it to measure what HBM gives this access pattern.
// =============================================================================================
template <int V>
__global__ void __launch_bounds__(kThreads)
copy_rows_kernel(const float* __restrict__ src, long long n_pts, long long st, const int32_t* __restrict__ mem_off,
                 const int32_t* __restrict__ mem_rows, float* __restrict__ dst) {
  constexpr int U = 8;
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = ((long long)blockIdx.x * 32 + lane) * V;
  const int m0 = mem_off[g], n_rows = mem_off[g + 1] - m0;
  if (pt >= n_pts) return;
  typedef typename std::conditional<V == 1, float, typename std::conditional<V == 2, float2, float4>::type>::type vec;
  for (int m = warp * U; m < n_rows; m += n_warps * U) {
    vec x[U];
#pragma unroll
    for (int j = 0; j < U; ++j)
      x[j] = *reinterpret_cast<const vec*>(src + (long long)mem_rows[m0 + min(m + j, n_rows - 1)] * st + pt);
#pragma unroll
    for (int j = 0; j < U; ++j)
      if (m + j < n_rows) *reinterpret_cast<vec*>(dst + (long long)mem_rows[m0 + m + j] * st + pt) = x[j];
  }
}

// =============================================================================================
// K7: MBCn / N-pdf transform helpers (_adjustment.py:289-328, 426-464; processing.py:323-350;
// _processing.py:184-247).
//   rotate_kernel      y[v][e] = sum_w R[v][w] x[w][e]   (rot @ x, _adjustment.py:311, 449; V <= 8)
//   standardize_kernel (x - nanmean) / nanstd(ddof=0) along time per (variable, point)
//   reorder_kernel     Schaake shuffle  sort(sim)[argsort(argsort(ref))]  per (point, group), window
//                      segments flattened and the centre column kept (_processing.py:204-211)
// =============================================================================================
constexpr int kMaxVar = 8;
struct RotMat { float r[kMaxVar * kMaxVar]; };

// FUSED = true: acc = fma(R[v][w], x[w], acc), w ascending from acc = 0 -- what `rot @ x` computes through
// OpenBLAS's sgemm / dgemm micro-kernels (_adjustment.py:311).  FUSED = false: acc = acc + R[v][w] * x[w] with the
// product rounded first -- numpy's einsum("ij,j...->i...") sum-of-products loop (_adjustment.py:449, 462), which is
// compiled for the SSE2 baseline (no FMA).  Both were pinned bit for bit against numpy 2.3 / OpenBLAS 0.3.30
// (tests/golden, test_mbcn_*); the N-pdf iteration is chaotic (one ulp here moves af_q by 1e-3 after 15 iterations), so
// the distinction matters.
template <typename T, bool FUSED>
__global__ void rotate_kernel(const T* __restrict__ x, long long n_elem, int n_var, RotMat R, T* __restrict__ y) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += (long long)gridDim.x * blockDim.x) {
    T in[kMaxVar];
#pragma unroll
    for (int w = 0; w < kMaxVar; ++w) if (w < n_var) in[w] = x[(long long)w * n_elem + e];
#pragma unroll
    for (int v = 0; v < kMaxVar; ++v) {
      if (v >= n_var) break;
      T acc = (T)0;
#pragma unroll
      for (int w = 0; w < kMaxVar; ++w)
        if (w < n_var)
          acc = FUSED ? Num<T>::fma((T)R.r[v * kMaxVar + w], in[w], acc)
                      : Num<T>::add(acc, Num<T>::mul((T)R.r[v * kMaxVar + w], in[w]));
      y[(long long)v * n_elem + e] = acc;
    }
  }
}

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum) of f(x[t]) over
// t in [t0, t0 + n): < 8 terms sequentially, <= 128 terms with 8 interleaved accumulators combined as a tree, longer
// runs split at n/2 rounded down to a multiple of 8.  Rounded in T after every operation.
template <typename T, class F>
__device__ T np_pairwise_leaf(const T* __restrict__ x, long long st, int t0, int n, F f) {
  if (n < 8) {
    T res = (T)0;
    for (int i = 0; i < n; ++i) res = Num<T>::add(res, f(x[(long long)(t0 + i) * st]));
    return res;
  }
  T r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = f(x[(long long)(t0 + j) * st]);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = Num<T>::add(r[j], f(x[(long long)(t0 + i + j) * st]));
  }
  T res = Num<T>::add(Num<T>::add(Num<T>::add(r[0], r[1]), Num<T>::add(r[2], r[3])),
                      Num<T>::add(Num<T>::add(r[4], r[5]), Num<T>::add(r[6], r[7])));
  for (; i < n; ++i) res = Num<T>::add(res, f(x[(long long)(t0 + i) * st]));
  return res;
}
// (the recursion of the original, unrolled into an explicit stack: depth log2(n / 128) + 1)
template <typename T, class F>
__device__ T np_pairwise_sum(const T* __restrict__ x, long long st, int n_total, F f) {
  struct Frame { int t0, n, state; T left; };
  Frame stk[24];
  int sp = 1;
  stk[0].t0 = 0; stk[0].n = n_total; stk[0].state = 0; stk[0].left = (T)0;
  T result = (T)0;
  while (sp > 0) {
    Frame& fr = stk[sp - 1];
    if (fr.state == 0) {
      if (fr.n <= 128) { result = np_pairwise_leaf<T>(x, st, fr.t0, fr.n, f); --sp; continue; }
      int n2 = fr.n / 2;
      n2 -= n2 % 8;
      fr.state = 1;
      stk[sp].t0 = fr.t0; stk[sp].n = n2; stk[sp].state = 0; stk[sp].left = (T)0;
      ++sp;
    } else if (fr.state == 1) {
      int n2 = fr.n / 2;
      n2 -= n2 % 8;
      fr.left = result;
      fr.state = 2;
      stk[sp].t0 = fr.t0 + n2; stk[sp].n = fr.n - n2; stk[sp].state = 0; stk[sp].left = (T)0;
      ++sp;
    } else {
      result = Num<T>::add(fr.left, result);
      --sp;
    }
  }
  return result;
}

// (x - nanmean(x)) / nanstd(x) along time, one thread per (variable, point), with numpy's arithmetic
// (processing.standardize is not used by _npdft_train: it calls np.nanmean / np.nanstd itself, _adjustment.py:303-305;
// numpy/lib/_nanfunctions_impl.py: NaNs replaced by 0, pairwise sums in the data dtype, the division by the count done
// in float64 and rounded to the data dtype, deviations and squares rounded to the data dtype).
template <typename T>
__global__ void standardize_kernel(const T* __restrict__ x, long long n_pts, long long sp, long long st, int n_time,
                                   int n_var, long long var_stride, T* __restrict__ y) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_pts * n_var) return;
  const long long pt = id % n_pts;
  const int v = (int)(id / n_pts);
  const T* xv = x + v * var_stride + pt * sp;
  T* yv = y + v * var_stride + pt * sp;
  int n = 0;
  for (int t = 0; t < n_time; ++t) n += is_nan(xv[(long long)t * st]) ? 0 : 1;
  const T tot = np_pairwise_sum<T>(xv, st, n_time, [](T a) { return is_nan(a) ? (T)0 : a; });
  const T mean = (T)((double)tot / (double)n);
  const T ssq = np_pairwise_sum<T>(xv, st, n_time, [mean](T a) {
    if (is_nan(a)) return (T)0;
    const T d = Num<T>::sub(a, mean);
    return Num<T>::mul(d, d);
  });
  T var = (T)((double)ssq / (double)n);
  if (n <= 0) var = Num<T>::nan();
  const T sd = Num<T>::sqrt(var);
  for (int t = 0; t < n_time; ++t) yv[(long long)t * st] = Num<T>::div(Num<T>::sub(xv[(long long)t * st], mean), sd);
}

// ordinal rank of every segment slot of `ref` (ties by slot order, NaNs last) and the sorted `sim`
// segment; member m (its centre slot) receives sorted_sim[rank(centre slot)].
template <typename T, int C>
__global__ void __launch_bounds__(kThreads)
reorder_kernel(const T* __restrict__ sim, const T* __restrict__ ref, long long n_pts, long long sp, long long st,
               const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows,
               const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows, int window, int n_pad,
               T* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sum = reinterpret_cast<double*>(smem_raw);
  int* cnt = reinterpret_cast<int*>(sum + C);
  T* sref = reinterpret_cast<T*>(smem_raw + C * 16);      // sorted ref keys [n_pad][C]
  T* ssim = sref + (size_t)n_pad * C;                     // sorted sim keys [n_pad][C]
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], n_mem = mem_off[g + 1] - m0;
  if (n_mem == 0) return;
  const int S = seg_off[g + 1] - seg_off[g];
  const int32_t* rows = seg_rows + seg_off[g];
  load_segment<T, C>(sref, ref, n0, n_pts, sp, st, rows, S, n_pad);
  load_segment<T, C>(ssim, sim, n0, n_pts, sp, st, rows, S, n_pad);
  __syncthreads();
  count_columns<T, C>(sref, n_pad, cnt, sum, false);
  make_keys<T, C>(sref, n_pad, cnt, sum, 0, XSDBA_KIND_ADD);
  make_keys<T, C>(ssim, n_pad, cnt, sum, 0, XSDBA_KIND_ADD);
  sort_columns<T, C>(sref, n_pad);
  sort_columns<T, C>(ssim, n_pad);
  const int half = window / 2;
  for (int item = threadIdx.x; item < n_mem * C; item += blockDim.x) {
    const int c = item % C, m = item / C;
    const long long pt = n0 + c;
    if (pt >= n_pts) continue;
    const int slot = m * window + half;                   // centre column of member m
    const int t = mem_rows[m0 + m];
    const long long o = pt * sp + (long long)t * st;
    T key = ref[o];
    const bool key_nan = is_nan(key);
    if (key_nan) key = Num<T>::inf();
    const T* col = sref + c;
    int lo = 0, hi = S;  // #keys < key over the whole padded segment (NaN keys are +inf: sorted last)
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] < key) lo = mid + 1; else hi = mid; }
    int rank = lo;
    // ties (and NaNs): np.argsort orders equal keys by position -> count equal keys in earlier slots
    if (lo + 1 < S && col[(size_t)(lo + 1) * C] == key) {
      for (int s2 = 0; s2 < slot; ++s2) {
        const int t2 = rows[s2];
        T k2 = t2 >= 0 ? ref[pt * sp + (long long)t2 * st] : Num<T>::nan();
        if (is_nan(k2)) k2 = Num<T>::inf();
        if (k2 == key) ++rank;
      }
    }
    const T v = ssim[(size_t)rank * C + c];
    // np.sort puts NaNs last: sorted positions >= (#valid sim) are NaN
    int n_sim = 0;  // (#valid sim values: keys below +inf; genuine +inf values are kept as they are)
    {
      int l2 = 0, h2 = S; const T inf = Num<T>::inf();
      const T* cs = ssim + c;
      while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (cs[(size_t)mid * C] < inf) l2 = mid + 1; else h2 = mid; }
      n_sim = l2;
    }
    out[o] = (rank >= n_sim && v == Num<T>::inf()) ? Num<T>::nan() : v;
  }
}

// =============================================================================================
// K8: per-(point, group) selections built on the segment sorter.
//   mode 0  nbutils.vecquantiles (nbutils.py:151-195): numba's np.nanquantile(row, rnk[row]) --
//           rank = 1 + (n-1)*((q*100)/100), val = lower*(1-m) + upper*m in float64 (numba
//           np/arraymath.py _collect_percentiles_inner), NaN rank -> NaN.
//   mode 1  utils.map_cdf / map_cdf_1d / _ecdf_1d (utils.py:35-84): q = (1 + #{y <= v}) / (1 + n_y),
//           then numpy's np.nanquantile(x, q) (float64 lerp, switch at gamma >= 0.5).
// grid = (ceil(n_pts / C), n_groups); out is [n_pts][n_groups] (mode 0) or [n_pts][n_groups][nv] (mode 1).
// =============================================================================================
template <typename T, int C>
__global__ void __launch_bounds__(kThreads)
select_kernel(const T* __restrict__ x, const T* __restrict__ y, long long n_pts, long long sp, long long st,
              const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_rows, int n_groups, int mode,
              const T* __restrict__ rnk, const double* __restrict__ yvals, int nv, T* __restrict__ out, int n_pad) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sum = reinterpret_cast<double*>(smem_raw);   // [C]
  int* cnt = reinterpret_cast<int*>(sum + C);          // [C]
  int* ycnt = cnt + C;                                 // [C]
  T* sm = reinterpret_cast<T*>(smem_raw + C * 16);     // [n_pad][C]
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int S = seg_off[g + 1] - seg_off[g];
  const int32_t* rows = seg_rows + seg_off[g];
  const int n_out = mode == 0 ? 1 : nv;
  if (S == 0) {
    for (int item = threadIdx.x; item < C * n_out; item += blockDim.x) {
      const int c = item / n_out, k = item % n_out;
      if (n0 + c < n_pts) out[((n0 + c) * n_groups + g) * n_out + k] = Num<T>::nan();
    }
    return;
  }
  load_segment<T, C>(sm, x, n0, n_pts, sp, st, rows, S, n_pad);
  __syncthreads();
  count_columns<T, C>(sm, n_pad, cnt, sum, false);
  make_keys<T, C>(sm, n_pad, cnt, sum, 0, XSDBA_KIND_ADD);
  sort_columns<T, C>(sm, n_pad);
  for (int k = 0; k < n_out; ++k) {
    double q = Num<double>::nan();
    if (mode == 1) {  // _ecdf_1d(y, v): y's valid count and #{y <= v}, per column
      if (threadIdx.x < C) { ycnt[threadIdx.x] = 0; cnt[threadIdx.x] = cnt[threadIdx.x]; sum[threadIdx.x] = 0.0; }
      __syncthreads();
      const double v = yvals[k];
      int my_le = 0, my_n = 0;
      for (int idx = threadIdx.x; idx < S * C; idx += blockDim.x) {  // idx % C constant per thread
        const int r = idx / C, c = idx % C;
        if (n0 + c >= n_pts) continue;
        const int t = rows[r];
        if (t < 0) continue;
        const T yv = y[(n0 + c) * sp + (long long)t * st];
        if (!is_nan(yv)) { ++my_n; if ((double)yv <= v) ++my_le; }
      }
      atomicAdd(&ycnt[threadIdx.x % C], my_le);
      atomicAdd(&sum[threadIdx.x % C], (double)my_n);
      __syncthreads();
    }
    if (threadIdx.x < C && n0 + threadIdx.x < n_pts) {
      const int c = threadIdx.x;
      const int n = cnt[c];
      const T* col = sm + c;
      double res = Num<double>::nan();
      if (mode == 0) {
        const T rk = rnk[(n0 + c) * n_groups + g];
        if (!is_nan(rk) && n > 0) {
          const double pct = (double)rk * 100.0;
          if (n == 1) res = (double)col[0];
          else if (pct == 100.0) res = (double)col[(size_t)(n - 1) * C];
          else if (pct == 0.0) res = (double)col[0];
          else {
            const double rank = 1.0 + (double)(n - 1) * (pct / 100.0);
            const double f = floor(rank);
            const double m = rank - f;
            long long kk = (long long)f - 1;
            kk = kk < 0 ? 0 : (kk > n - 2 ? n - 2 : kk);
            const double lower = (double)col[(size_t)kk * C], upper = (double)col[(size_t)(kk + 1) * C];
            res = __dadd_rn(__dmul_rn(lower, __dsub_rn(1.0, m)), __dmul_rn(upper, m));
          }
        }
      } else {
        q = (1.0 + (double)ycnt[c]) / (1.0 + sum[c]);   // utils.py:35-37
        if (n > 0 && q == q) {
          if (q > 1.0) q = 1.0;
          const double vi = q * (double)(n - 1);
          long long i0 = (long long)floor(vi);
          i0 = i0 < 0 ? 0 : (i0 > n - 1 ? n - 1 : i0);
          const long long i1 = i0 + 1 > n - 1 ? n - 1 : i0 + 1;
          const double gm = vi - (double)i0;
          const T aT = col[(size_t)i0 * C], bT = col[(size_t)i1 * C];
          const double a = (double)aT, b = (double)bT;
          const double d = (double)Num<T>::sub(bT, aT);  // numpy _lerp: subtract(b, a) in the array dtype
          res = gm >= 0.5 ? __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, gm))) : __dadd_rn(a, __dmul_rn(d, gm));
        }
      }
      out[((n0 + c) * n_groups + g) * n_out + k] = (T)res;
    }
    __syncthreads();
  }
}

// =============================================================================================
// K9: frequency adaptation of sim with stored factors (the adjust-side _adapt_freq_preprocess,
// _adjustment.py:32-45, 639-646 -> _adapt_freq.func with P0_ref / P0_hist / pth given, _processing.py:75-122)
// and the max_tail_factor mask (_adjustment.py:647-658, 672-673).  grid = (ceil(n_pts / C), n_groups).
// The reference breaks rank ties with numpy's global RNG and fills with np.random.random_sample; here both
// draws are counter-based hashes of (seed, element), so parity of the random part is distributional.
// =============================================================================================
template <typename T, int C>
__global__ void __launch_bounds__(kThreads)
adapt_apply_kernel(const T* __restrict__ sim, long long n_pts, long long sp, long long st,
                   const int32_t* __restrict__ mem_off, const int32_t* __restrict__ mem_rows, int n_groups,
                   double thresh, const double* __restrict__ P0_ref, const double* __restrict__ P0_hist,
                   const T* __restrict__ pth, unsigned long long seed, T* __restrict__ out, int n_pad) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sum = reinterpret_cast<double*>(smem_raw);
  int* cnt = reinterpret_cast<int*>(sum + C);
  int* nle = cnt + C;
  T* sm = reinterpret_cast<T*>(smem_raw + C * 16);
  const int g = blockIdx.y;
  const long long n0 = (long long)blockIdx.x * C;
  const int m0 = mem_off[g], n_mem = mem_off[g + 1] - m0;
  if (n_mem == 0) return;
  load_segment<T, C>(sm, sim, n0, n_pts, sp, st, mem_rows + m0, n_mem, n_pad);
  __syncthreads();
  count_le_columns<T, C>(sm, n_pad, cnt, nle, thresh);
  if (threadIdx.x < C) sum[threadIdx.x] = (double)nle[threadIdx.x] / (double)cnt[threadIdx.x];  // P0_sim
  __syncthreads();
  make_keys<T, C>(sm, n_pad, cnt, sum, 0, XSDBA_KIND_ADD);
  sort_columns<T, C>(sm, n_pad);
  for (int item = threadIdx.x; item < n_mem * C; item += blockDim.x) {
    const int c = item % C, m = item / C;
    const long long pt = n0 + c;
    if (pt >= n_pts) continue;
    const long long o = pt * sp + (long long)mem_rows[m0 + m] * st;
    const T x = sim[o];
    T res = x;
    const int n = cnt[c];
    const long long og = pt * n_groups + g;
    const double p0r = P0_ref[og], p0h = P0_hist[og], p0s = sum[c];
    const double dp0 = p0h == 0.0 ? Num<double>::nan() : (p0h - p0r) / p0h;
    if (!is_nan(x) && dp0 > 0.0 && n > 1) {
      const T* col = sm + c;
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] < x) lo = mid + 1; else hi = mid; }
      const int lb = lo;
      hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] <= x) lo = mid + 1; else hi = mid; }
      const int ub = lo;
      // Random tie-break (utils.py:618-627: ranks of the values plus a small noise, i.e. the tied values take the
      // positions lb .. ub-1 in a random order -- a permutation, every position exactly once).  The position of
      // this sample inside its tie block is its rank among the tied samples ordered by their hash draw (then by
      // address); it only matters when the block straddles a keep / replace boundary, which is checked first.
      const unsigned long long id = (unsigned long long)o;
      const double lim_lo = (p0r / p0h) * p0s;
      auto kept = [&](int r) { const double rk = (double)r / (double)(n - 1); return (rk < lim_lo) || (rk > p0s); };
      bool keep = kept(lb);
      bool mixed = ub - lb > 1 && kept(ub - 1) != keep;
      if (!mixed && keep && ub - lb > 2) {
        // both ends kept: the block may still contain the whole replaced interval [lim_lo, P0_sim] (the rule is
        // "replace" on one interval of positions, so two replaced ends do prove a replaced middle)
        const int r_first = (int)ceil(lim_lo * (double)(n - 1));
        mixed = r_first > lb && r_first < ub - 1 && !kept(r_first);
      }
      if (mixed) {
        const double h_me = hash_uniform(seed, 2 * id);
        int before = 0;
        for (int m2 = 0; m2 < n_mem; ++m2) {
          const long long o2 = pt * sp + (long long)mem_rows[m0 + m2] * st;
          if (o2 == o || !(sim[o2] == x)) continue;
          const double h2 = hash_uniform(seed, 2 * (unsigned long long)o2);
          before += (h2 < h_me || (h2 == h_me && o2 < o)) ? 1 : 0;
        }
        keep = kept(lb + before);
      }
      if (!keep) res = (T)(((double)pth[og] - thresh) * (double)(T)hash_uniform(seed, 2 * id + 1) + thresh);
    }
    out[o] = res;
  }
}

// scen = where(adapted_sim > max_tail_factor * hist_q_raw[..., -1] (broadcast nearest), adapted_sim, scen)
template <typename T>
__global__ void tail_mask_kernel(const T* __restrict__ adapted, long long n_pts, long long sp, long long st, int n_time,
                                 const int32_t* __restrict__ gidx, int n_groups, const T* __restrict__ hq_raw, int nq,
                                 double factor, T* __restrict__ scen) {
  const long long total = n_pts * n_time;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long pt = sp == 1 ? e % n_pts : e / n_time;
    const int t = (int)(sp == 1 ? e / n_pts : e % n_time);
    const int g = gidx[t];
    if (g < 0) continue;
    const long long o = pt * sp + (long long)t * st;
    const T a = adapted[o];
    const T lastq = hq_raw[(pt * n_groups + g) * nq + nq - 1];
    if ((double)a > factor * (double)lastq) scen[o] = a;
  }
}

// =============================================================================================
// K10: energy score (nbutils._escore, nbutils.py:274-372; processing.escore, processing.py:393-489) --
// Szekely-Rizzo e-distance between the K-variate clouds tgt (K x n2) and sim (K x n1) of every gridpoint.
// Arrays are (variable, time, point) with variables var_stride apart; observations with a NaN in any
// variable are removed per array (remove_NaNs).  One CTA per point; O(n1*n2*K) pair distances accumulated
// in float64 (the reference accumulates in the data dtype under fastmath: its summation order is unpinned).
//   out = n1*n2/(n1+n2) * (2*mean|X-Y| - mean|X-X'| - mean|Y-Y'|) / 2
// =============================================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
escore_kernel(const T* __restrict__ tgt, const T* __restrict__ sim, long long n_pts, long long sp, long long st,
              int n_time_t, int n_time_s, int step_t, int step_s, int n_var, long long var_stride_t,
              long long var_stride_s, T* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  int* idx_t = reinterpret_cast<int*>(smem_raw);                 // valid observation rows of tgt
  int* idx_s = idx_t + (n_time_t + step_t - 1) / step_t;         // valid observation rows of sim
  __shared__ int n_t, n_s;
  __shared__ double red[3][kThreads / 32];
  const long long pt = blockIdx.x;
  if (threadIdx.x == 0) {  // compaction of the valid observations (serial: tiny next to the O(n^2) loops)
    int a = 0;
    for (int t = 0; t < n_time_t; t += step_t) {
      bool ok = true;
      for (int v = 0; v < n_var; ++v) ok = ok && !is_nan(tgt[v * var_stride_t + pt * sp + (long long)t * st]);
      if (ok) idx_t[a++] = t;
    }
    n_t = a;
    a = 0;
    for (int t = 0; t < n_time_s; t += step_s) {
      bool ok = true;
      for (int v = 0; v < n_var; ++v) ok = ok && !is_nan(sim[v * var_stride_s + pt * sp + (long long)t * st]);
      if (ok) idx_s[a++] = t;
    }
    n_s = a;
  }
  __syncthreads();
  const int n2 = n_t, n1 = n_s;
  double sxy = 0, sxx = 0, syy = 0;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    T xi[kMaxVar];
    for (int v = 0; v < n_var; ++v) xi[v] = tgt[v * var_stride_t + pt * sp + (long long)idx_t[i] * st];
    for (int j = 0; j < n1; ++j) {
      double d1 = 0;
      for (int v = 0; v < n_var; ++v) {
        const double d = (double)xi[v] - (double)sim[v * var_stride_s + pt * sp + (long long)idx_s[j] * st];
        d1 += d * d;
      }
      sxy += sqrt(d1);
    }
    for (int j = 0; j < i; ++j) {
      double d1 = 0;
      for (int v = 0; v < n_var; ++v) {
        const double d = (double)xi[v] - (double)tgt[v * var_stride_t + pt * sp + (long long)idx_t[j] * st];
        d1 += d * d;
      }
      sxx += sqrt(d1);
    }
  }
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    T yi[kMaxVar];
    for (int v = 0; v < n_var; ++v) yi[v] = sim[v * var_stride_s + pt * sp + (long long)idx_s[i] * st];
    for (int j = 0; j < i; ++j) {
      double d1 = 0;
      for (int v = 0; v < n_var; ++v) {
        const double d = (double)yi[v] - (double)sim[v * var_stride_s + pt * sp + (long long)idx_s[j] * st];
        d1 += d * d;
      }
      syy += sqrt(d1);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sxy += __shfl_xor_sync(0xffffffffu, sxy, o);
    sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
    syy += __shfl_xor_sync(0xffffffffu, syy, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sxy; red[1][threadIdx.x >> 5] = sxx; red[2][threadIdx.x >> 5] = syy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
    double res = Num<double>::nan();
    if (n1 > 0 && n2 > 0) {
      const double mxy = a / ((double)n2 * (double)n1);
      const double mxx = 2.0 * b / ((double)n2 * (double)n2);
      const double myy = 2.0 * c / ((double)n1 * (double)n1);
      const double w = (double)n1 * (double)n2 / (double)(n1 + n2);
      res = w * (mxy + mxy - mxx - myy) / 2.0;
    }
    out[pt] = (T)res;
  }
}

// elementwise jitter (processing.jitter / jitter_under_thresh / jitter_over_thresh, processing.py:124-257)
template <typename T>
__global__ void jitter_kernel(const T* __restrict__ x, long long n, JitterParams jp, T* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = jitter_value<T>(x[i], jp, (unsigned long long)i);
}


// =============================================================================================
// K7n: one variable of one N-pdf iteration, fused (float32 data, one "time" group per block of time steps).
//   train  (ref given):   ref_q, hist_q = quantiles at the float64 nodes (nbutils._quantile, _adjustment.py:315),
//                         af = ref_q - hist_q (:316), hist += interp1d(rank_bn(hist), q, af) (:317-324)
//   adjust (ref == NULL): x += interp1d(rank_bn(x), q, af) with the stored af (_adjustment.py:453-460)
// The reference runs the lookup in float64 (af_q is a float64 array, the nodes are float64) and stores the sum
// into the float32 series.  Before this kernel the step took three launches and three segment sorts (ref, hist,
// and hist again, as float64, for the ranks); here the sorted hist serves quantiles and ranks, the keys stay
// float32 (exact: the data are float32) and only the tables are float64.  grid = ceil(n_pts / C), C columns per CTA.
// =============================================================================================
template <int C>
__global__ void __launch_bounds__(threads_for_cols<C>())
npdft_step_kernel(const float* __restrict__ ref, float* __restrict__ hist, long long n_pts, long long sp, long long st,
                  const int32_t* __restrict__ seg_rows, int S, const double* __restrict__ q64, int nq, int interp,
                  int extrap, float* __restrict__ af_io /*[n_pts][nq]: written (train) or read (adjust)*/, int n_pad) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sum = reinterpret_cast<double*>(smem_raw);       // [C] (count_columns scratch), then max average rank
  double* mnv = sum + C;                                   // [C] smallest normalised rank
  int* cnt = reinterpret_cast<int*>(mnv + C);              // [C]
  unsigned char* p = smem_raw + C * 24;
  float* refq = reinterpret_cast<float*>(p);               // [nq][C]
  float* sm = refq + (size_t)nq * C;                       // [n_pad][C] sort keys
  unsigned char* tp = reinterpret_cast<unsigned char*>(sm + (size_t)n_pad * C);
  tp += (8 - (reinterpret_cast<size_t>(tp) & 7)) & 7;
  Tables<double, C> tb = carve_tables<double, C>(tp, nq, 1);
  const long long n0 = (long long)blockIdx.x * C;
  const bool train = ref != nullptr;

  if (train) {
    load_segment<float, C>(sm, ref, n0, n_pts, sp, st, seg_rows, S, n_pad);
    __syncthreads();
    count_columns<float, C>(sm, n_pad, cnt, sum, false);
    make_keys<float, C>(sm, n_pad, cnt, sum, 0, XSDBA_KIND_ADD);
    sort_columns<float, C>(sm, n_pad);
    for (int item = threadIdx.x; item < C * nq; item += blockDim.x) {
      const int c = item % C, k = item / C;
      refq[k * C + c] = quantile_sorted<float, C>(sm + c, cnt[c], S, q64[k]);
    }
    __syncthreads();
  }
  load_segment<float, C>(sm, hist, n0, n_pts, sp, st, seg_rows, S, n_pad);
  __syncthreads();
  count_columns<float, C>(sm, n_pad, cnt, sum, false);
  make_keys<float, C>(sm, n_pad, cnt, sum, 0, XSDBA_KIND_ADD);
  sort_columns<float, C>(sm, n_pad);
  // factors of this (point, variable): computed and stored (train) or loaded (adjust); staged as float64 tables
  for (int item = threadIdx.x; item < C * nq; item += blockDim.x) {
    const int c = item % C, k = item / C;
    if (n0 + c >= n_pts) { refq[k * C + c] = Num<float>::nan(); continue; }
    float a;
    if (train) {
      a = __fsub_rn(refq[k * C + c], quantile_sorted<float, C>(sm + c, cnt[c], S, q64[k]));
      af_io[(n0 + c) * nq + k] = a;
    } else {
      a = af_io[(n0 + c) * nq + k];
    }
    refq[k * C + c] = a;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x, n = cnt[c];
    // NaN nodes dropped (utils.py:351-352), first / last non-NaN factor for the constant extrapolation (:362-368)
    double clo = Num<double>::nan(), chi = clo;
    bool have = false;
    int w = 0;
    double* xs = tb.xsl[1];
    double* ys = tb.ysl[1];
    for (int k = 0; k < nq; ++k) {
      const double y = (double)refq[k * C + c], x = q64[k];
      if (y == y) { if (!have) { clo = y; have = true; } chi = y; }
      if (y == y && x == x) { xs[(size_t)w * C + c] = x; ys[(size_t)w * C + c] = y; ++w; }
    }
    for (int k = w; k < tb.ld; ++k) xs[(size_t)k * C + c] = Num<double>::inf();
    tb.nvl[1][c] = w;
    tb.blo[c] = q64[0]; tb.bhi[c] = q64[nq - 1]; tb.clo[c] = clo; tb.chi[c] = chi;
    // _rank_bn (utils.py:641-646): rnk / nanmax(rnk), then (rnk - mn) / (1 - mn)
    double mx_rank = Num<double>::nan(), mn = mx_rank;
    if (n > 0) {
      const float* col = sm + c;
      const float vmin = col[0], vmax = col[(size_t)(n - 1) * C];
      int ub = 1; while (ub < n && col[(size_t)ub * C] == vmin) ++ub;
      int lb = n - 1; while (lb > 0 && col[(size_t)(lb - 1) * C] == vmax) --lb;
      mx_rank = (double)(lb + n + 1) * 0.5;
      mn = ((double)(ub + 1) * 0.5) / mx_rank;
    }
    sum[c] = mx_rank; mnv[c] = mn;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < S * C; item += blockDim.x) {
    const int c = item % C;
    const long long pt = n0 + c;
    const int t = seg_rows[item / C];
    if (pt >= n_pts || t < 0) continue;
    const long long o = pt * sp + (long long)t * st;
    const float x = hist[o];
    double sq = Num<double>::nan();
    const int n = cnt[c];
    if (x == x && n > 0) {
      const float* col = sm + c;
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] < x) lo = mid + 1; else hi = mid; }
      const int lb = lo;
      hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[(size_t)mid * C] <= x) lo = mid + 1; else hi = mid; }
      const double r = ((double)(lb + lo + 1) * 0.5) / sum[c];
      sq = __ddiv_rn(__dsub_rn(r, mnv[c]), __dsub_rn(1.0, mnv[c]));
    }
    const double f = lookup_1d<double, double, C>(tb, c, sq, interp, extrap);
    hist[o] = (float)__dadd_rn((double)x, f);   // float32 + float64 -> float64, stored into the float32 series
  }
}

// =============================================================================================
// host side
// =============================================================================================
inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? XSDBA_OK : (int)e; }
#define XS_CUDA(call)                                  \
  do {                                                 \
    cudaError_t _e = (call);                           \
    if (_e != cudaSuccess) return (int)_e;             \
  } while (0)

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

template <typename T> int pick_cols(int n_pad) {
  int c = (int)(kSortBytes / sizeof(T)) / n_pad;
  if (c >= 32) return 32;
  int p = 1; while (p * 2 <= c) p *= 2;
  return c < 1 ? 0 : p;
}

template <typename K> int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) XS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return XSDBA_OK;
}

// Scratch buffers come from the device's stream-ordered pool (cudaMallocAsync).  Its default release threshold of 0
// hands the memory back to the driver at every synchronisation, which costs milliseconds per call (tens for the
// gigabyte-sized LOESS scratch): keep it.  Called by every launcher that allocates scratch.
// The threshold is set once per DEVICE (a process may drive several) and is bounded: up to 8 GiB of freed scratch
// stays in the pool, anything above goes back to the driver so that other allocators of the process (PyTorch's caching
// allocator does not draw from this pool) are not starved after one large call.  xsdba_trim_pool() returns the rest.
void keep_pool_memory() {
  static std::atomic<int> pool_ready[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return; }
  if (pool_ready[dev].exchange(1)) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = 8ull << 30;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
}

// number of SMs of the current device (grid sizing of the grid-stride kernels)
int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
  int n = cached[dev].load();
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    cached[dev].store(n);
  }
  return n;
}

// every launcher checks that the grouping handle lives on the current device
inline bool wrong_device(const xsdba_grouping* grp) {
  int dev = -1;
  return grp && (cudaGetDevice(&dev) != cudaSuccess || dev != grp->device);
}

template <typename T, int C>
int launch_train_c(const T* ref, const T* hist, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                   const T* q, int nq, int kind, int normalize, int mode, T* af, T* hq, T* scaling, int n_pad,
                   cudaStream_t s, const JitterParams& jp, int use_jitter, const double* q64, const AdaptParams& ap) {
  const size_t smem = (size_t)C * 24 + ((size_t)nq * C + (size_t)n_pad * C) * sizeof(T);
  auto kern = train_kernel<T, C>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  dim3 grid((unsigned)((n_pts + C - 1) / C), (unsigned)grp->n_groups);
  kern<<<grid, threads_for_cols<C>(), smem, s>>>(ref, hist, n_pts, sp, st, grp->segments.off, grp->segments.rows,
                                                 grp->n_groups, q, nq, kind, normalize, mode, af, hq, scaling, n_pad, jp,
                                                 use_jitter, q64, ap);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

// float32 / time-major / <= 1024-slot segments take the register-blocked kernel; returns false otherwise
bool launch_train_fast(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                       const xsdba_grouping* grp, const float* q, int nq, int kind, int normalize, int mode, float* af,
                       float* hq, float* scaling, cudaStream_t s, int* rc, const JitterParams& jp, int use_jitter,
                       const double* q64) {
  if (sp != 1 || st < 0 || st > INT32_MAX / 4 || grp->segments.max_len > 1024 || nq > kFastMaxNq ||
      getenv("XSDBA_B200_NO_FAST"))
    return false;
  dim3 grid((unsigned)((n_pts + 31) / 32), (unsigned)grp->n_groups);
  // K1b (bucket select) is the default; XSDBA_B200_TRAIN_ALGO=sort keeps K1f's sorting network for A/B runs
  const char* algo = getenv("XSDBA_B200_TRAIN_ALGO");
  // (its shared-memory image fits the 227 KB of an SM up to nq ~ 110; finer grids keep the sorter)
  if (!(algo && strcmp(algo, "sort") == 0) && BktSmem::total(nq) <= 227 * 1024) {
    const size_t smem_b = BktSmem::total(nq);
    const long long n_tiles_b = (n_pts + 31) / 32;
    // persistent: one CTA per SM (its shared-memory image leaves room for one) walks the (group, tile) items.  The
    // normalising (DQM) variant keeps one CTA per item: measured 9 % faster that way on day-of-year groupings, where a
    // persistent CTA changes group -- and rebuilds the group's tables -- every 2.4 items.
    const long long n_work_b = n_tiles_b * grp->n_groups;
    if (normalize && n_work_b > 0x7fffffffLL) return false;  // (one CTA per item must fit grid.x: the generic kernel otherwise)
    const dim3 grid_b((unsigned)(normalize ? n_work_b : std::min<long long>(n_work_b, (long long)sm_count())));
    static const int vec_enable = getenv("XSDBA_B200_NO_VEC_LOAD") ? 0 : 1;  // (A/B switch of the 16-byte load path)
#define XS_TRAIN_BKT(J, N)                                                                                            \
  do {                                                                                                               \
    *rc = set_smem(train_bucket_kernel<J, N>, smem_b);                                                               \
    if (*rc) return true;                                                                                            \
    train_bucket_kernel<J, N><<<grid_b, kFastThreads, smem_b, s>>>(ref, hist, n_pts, st, grp->segments.off,          \
                                                                 grp->segments.rows, grp->n_groups, q, nq, kind,    \
                                                                 normalize, mode, af, hq, scaling, jp, use_jitter,  \
                                                                 q64, vec_enable, n_tiles_b);                       \
  } while (0)
    if (use_jitter && normalize) XS_TRAIN_BKT(true, true);
    else if (use_jitter) XS_TRAIN_BKT(true, false);
    else if (normalize) XS_TRAIN_BKT(false, true);
    else XS_TRAIN_BKT(false, false);
#undef XS_TRAIN_BKT
    ++g_launches;
    *rc = cuda_status(cudaGetLastError());
    return true;
  }
  const size_t smem = FastSmem::total(nq);
  static const int stagger_ns = getenv("XSDBA_B200_STAGGER_NS") ? atoi(getenv("XSDBA_B200_STAGGER_NS")) : 0;
#define XS_TRAIN_FAST(J, N)                                                                                          \
  do {                                                                                                               \
    *rc = set_smem(train_fast_kernel<J, N>, smem);                                                                   \
    if (*rc) return true;                                                                                            \
    train_fast_kernel<J, N><<<grid, kFastThreads, smem, s>>>(ref, hist, n_pts, st, grp->segments.off,                \
                                                             grp->segments.rows, grp->n_groups, q, nq, kind,        \
                                                             normalize, mode, af, hq, scaling, jp, use_jitter, q64, \
                                                             stagger_ns);                                           \
  } while (0)
  if (use_jitter && normalize) XS_TRAIN_FAST(true, true);
  else if (use_jitter) XS_TRAIN_FAST(true, false);
  else if (normalize) XS_TRAIN_FAST(false, true);
  else XS_TRAIN_FAST(false, false);
#undef XS_TRAIN_FAST
  ++g_launches;
  *rc = cuda_status(cudaGetLastError());
  return true;
}
bool launch_train_fast(const double*, const double*, int64_t, int64_t, int64_t, const xsdba_grouping*, const double*,
                       int, int, int, int, double*, double*, double*, cudaStream_t, int*, const JitterParams&, int,
                       const double*) {
  return false;
}

// groupings with heavily overlapping windows (day-of-year x 31): one ordering per chunk of groups (K1w)
bool launch_train_window(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                         const xsdba_grouping* grp, const float* q, int nq, int kind, int mode, float* af, float* hq,
                         cudaStream_t s, int* rc) {
  if (sp != 1 || st < 0 || grp->win.n_chunks <= 0 || getenv("XSDBA_B200_NO_WINDOW_KERNEL") || getenv("XSDBA_B200_NO_FAST"))
    return false;
  *rc = set_smem(train_window_kernel, WinSmem::total);
  if (*rc) return true;
  dim3 grid((unsigned)((n_pts + kWinCols - 1) / kWinCols), (unsigned)grp->win.n_chunks);
  train_window_kernel<<<grid, kWinThreads, WinSmem::total, s>>>(ref, hist, n_pts, st, grp->segments.off, grp->win.chunk_g,
                                                                grp->win.urow_off, grp->win.urows, grp->win.gmask,
                                                                grp->n_groups, q, nq, kind, mode, af, hq);
  ++g_launches;
  *rc = cuda_status(cudaGetLastError());
  return true;
}
bool launch_train_window(const double*, const double*, int64_t, int64_t, int64_t, const xsdba_grouping*, const double*, int,
                         int, int, double*, double*, cudaStream_t, int*) {
  return false;
}

template <typename T>
int launch_train(const T* ref, const T* hist, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                 const T* q, int nq, int kind, int normalize, int mode, T* af, T* hq, T* scaling, void* stream,
                 const double* jitter = nullptr, unsigned long long seed = 0, const double* q64 = nullptr,
                 const AdaptParams* adapt = nullptr) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  int fast_rc = 0;
  JitterParams jp;
  const double dnan = __builtin_nan("");
  jp.lower = jitter ? jitter[0] : dnan; jp.minimum = jitter ? jitter[1] : 0.0;
  jp.upper = jitter ? jitter[2] : dnan; jp.maximum = jitter ? jitter[3] : 0.0; jp.seed = seed;
  const int use_jitter = jitter && (jitter[0] == jitter[0] || jitter[2] == jitter[2]);
  if ((n_pts > 0 && !ref) || !grp || (n_pts > 0 && !q) || (n_pts > 0 && !af) || n_pts < 0 || nq <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (mode == 0 && n_pts > 0 && (!hist || !hq)) return XSDBA_ERR_INVALID_ARGUMENT;
  if (kind != XSDBA_KIND_ADD && kind != XSDBA_KIND_MUL) return XSDBA_ERR_INVALID_ARGUMENT;
  if (normalize && mode == 0 && n_pts > 0 && !scaling) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n_pts == 0) return XSDBA_OK;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  const int n_pad = std::max(2, next_pow2(grp->segments.max_len));
  const int C = pick_cols<T>(n_pad);
  cudaStream_t s = (cudaStream_t)stream;
  AdaptParams ap{};
  if (adapt) ap = *adapt;
  if (!ap.on && !use_jitter && !normalize && !q64 &&
      launch_train_window(ref, hist, n_pts, sp, st, grp, q, nq, kind, mode, af, hq, s, &fast_rc))
    return fast_rc;
  if (!ap.on &&
      launch_train_fast(ref, hist, n_pts, sp, st, grp, q, nq, kind, normalize, mode, af, hq, scaling, s, &fast_rc, jp, use_jitter, q64))
    return fast_rc;
#define XS_CASE(CC) case CC: return launch_train_c<T, CC>(ref, hist, n_pts, sp, st, grp, q, nq, kind, normalize, mode, af, hq, scaling, n_pad, s, jp, use_jitter, q64, ap)
  switch (C) {
    XS_CASE(32); XS_CASE(16); XS_CASE(8); XS_CASE(4); XS_CASE(2); XS_CASE(1);
    default: return XSDBA_ERR_SEGMENT_TOO_LONG;
  }
#undef XS_CASE
}

template <int TOP>
bool launch_adjust_tile_t(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                          const float* af, const float* hq, int nq, int extrap, int kind, float* scen, cudaStream_t s) {
  const size_t smem_t = 2 * tile_slot_floats(TOP, nq) * sizeof(float) + 2 * kTileMaxRows * sizeof(int) +
                        4 * sizeof(uint64_t) + kFixStage * sizeof(FixEntry) + 2 * sizeof(unsigned) + 128;
  const size_t smem_p = stage_bytes<float, 32>(nq);
  if (smem_t > 220 * 1024 || smem_p > 220 * 1024 || st < 0 || st > INT32_MAX / 4) return false;
  const int64_t tiles = (n_pts + 31) / 32;
  float* packed = nullptr;
  keep_pool_memory();
  if (cudaMallocAsync(&packed, (size_t)tiles * grp->n_groups * packed_slot_floats(nq) * sizeof(float), s) != cudaSuccess) {
    cudaGetLastError();
    return false;  // not enough memory for the packed image: the caller falls back to the staging kernel
  }
  if (set_smem(pack_tables_kernel, smem_p) || set_smem(adjust_tile_kernel<TOP>, smem_t)) {
    cudaFreeAsync(packed, s);
    return false;
  }
  // deferred-sample list: capacity 1/64 of the samples (ties are ~1e-4); on overflow the generic exact kernel
  // redoes the whole block (it exits immediately otherwise), so the result never depends on the capacity
  static const unsigned cap_div = getenv("XSDBA_B200_FIX_CAP_DIV") ? (unsigned)atoi(getenv("XSDBA_B200_FIX_CAP_DIV")) : 64u;
  const unsigned fix_cap = (unsigned)std::min<int64_t>(std::max<int64_t>(n_pts * grp->n_time / cap_div, 1024), 1 << 28);
  FixEntry* fix = nullptr;
  unsigned* fix_count = nullptr;
  if (cudaMallocAsync(&fix, (size_t)fix_cap * sizeof(FixEntry), s) != cudaSuccess ||
      cudaMallocAsync(&fix_count, sizeof(unsigned), s) != cudaSuccess) {
    cudaGetLastError();
    if (fix) cudaFreeAsync(fix, s);
    cudaFreeAsync(packed, s);
    return false;
  }
  if (cudaMemsetAsync(fix_count, 0, sizeof(unsigned), s) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeAsync(packed, s); cudaFreeAsync(fix, s); cudaFreeAsync(fix_count, s);
    return false;
  }
  pack_tables_kernel<<<dim3((unsigned)tiles, (unsigned)grp->n_groups), kThreads, smem_p, s>>>(
      af, hq, n_pts, grp->n_groups, nq, extrap, packed);
  adjust_tile_kernel<TOP><<<(unsigned)tiles, kThreads, smem_t, s>>>(sim, n_pts, sp, (int)st, grp->members.off,
                                                                    grp->members.rows, grp->n_groups, nq, packed, kind,
                                                                    scen, fix, fix_count, fix_cap);
  static const bool fix_scan = getenv("XSDBA_B200_FIX_SCAN") != nullptr;  // (debug: scan every row in the second pass)
  adjust_fix_kernel<<<sm_count() * 6, 256, 0, s>>>(fix, fix_count, fix_cap, af, hq, grp->n_groups, nq, extrap, kind,
                                                   fix_scan ? nullptr : packed, scen);
  {
    const size_t smem_g = ((tables_bytes<float, 32>(nq) + 15) & ~(size_t)15) + stage_bytes<float, 32>(nq);
    auto kern = adjust_kernel<float, 32>;
    if (set_smem(kern, smem_g) == XSDBA_OK)
      kern<<<dim3((unsigned)tiles, (unsigned)grp->n_groups), kThreads, smem_g, s>>>(
          sim, n_pts, sp, st, grp->members.off, grp->members.rows, grp->n_groups, af, hq, nq, XSDBA_INTERP_NEAREST, extrap,
          kind, scen, fix_count, fix_cap);
  }
  g_launches += 4;
  cudaFreeAsync(packed, s); cudaFreeAsync(fix, s); cudaFreeAsync(fix_count, s);
  return true;
}

bool launch_adjust_tile(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                        const float* af, const float* hq, int nq, int top, int extrap, int kind, float* scen,
                        cudaStream_t s) {
  switch (top) {
    case 8: return launch_adjust_tile_t<8>(sim, n_pts, sp, st, grp, af, hq, nq, extrap, kind, scen, s);
    case 16: return launch_adjust_tile_t<16>(sim, n_pts, sp, st, grp, af, hq, nq, extrap, kind, scen, s);
    case 32: return launch_adjust_tile_t<32>(sim, n_pts, sp, st, grp, af, hq, nq, extrap, kind, scen, s);
    case 64: return launch_adjust_tile_t<64>(sim, n_pts, sp, st, grp, af, hq, nq, extrap, kind, scen, s);
    default: return false;
  }
}

bool launch_adjust_fast(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                        const float* af, const float* hq, int nq, int interp, int extrap, int kind, float* scen,
                        size_t smem, dim3 grid, cudaStream_t s) {
  if (grp->n_groups <= 1 || interp != XSDBA_INTERP_NEAREST || getenv("XSDBA_B200_NO_FAST")) return false;
  if (grp->members.max_len > kAdjMaxRows) return false;
  int top = 1;
  while (top * 2 <= nq) top *= 2;
  {
    int top_t = 8;  // K2t searches nq + 2 intervals: 2 * top_t - 1 >= nq + 1
    while (2 * top_t < nq + 2) top_t *= 2;
    if (grp->members.max_len <= kTileMaxRows - 8 && !getenv("XSDBA_B200_NO_TILE") &&
        launch_adjust_tile(sim, n_pts, sp, st, grp, af, hq, nq, top_t, extrap, kind, scen, s))
      return true;
  }
  smem += (size_t)grp->members.max_len * sizeof(int);
#define XS_LAUNCH(TOP)                                                                                          \
  case TOP:                                                                                                     \
    if (set_smem(adjust_fast_kernel<TOP>, smem)) return false;                                                  \
    adjust_fast_kernel<TOP><<<grid, kThreads, smem, s>>>(sim, n_pts, sp, st, grp->members.off, grp->members.rows, \
                                                         grp->n_groups, af, hq, nq, extrap, kind, scen);        \
    break;
  switch (top) {
    XS_LAUNCH(4) XS_LAUNCH(8) XS_LAUNCH(16) XS_LAUNCH(32) XS_LAUNCH(64) XS_LAUNCH(128)
    default: return false;
  }
#undef XS_LAUNCH
  ++g_launches;
  return true;
}
bool launch_adjust_fast(const double*, int64_t, int64_t, int64_t, const xsdba_grouping*, const double*, const double*,
                        int, int, int, int, double*, size_t, dim3, cudaStream_t) { return false; }

template <typename T, int C>
bool launch_adjust_narrow(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* af,
                          const T* hq, int nq, int interp, int extrap, int kind, T* scen, void* stream) {
  const size_t smem = ((tables_bytes<T, C>(nq) + 15) & ~(size_t)15) + stage_bytes<T, C>(nq);
  auto kern = adjust_kernel<T, C>;
  if (smem > 200 * 1024 || set_smem(kern, smem) != XSDBA_OK) return false;
  dim3 grid((unsigned)((n_pts + C - 1) / C), (unsigned)grp->n_groups);
  kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(sim, n_pts, sp, st, grp->members.off, grp->members.rows,
                                                       grp->n_groups, af, hq, nq, interp, extrap, kind, scen, nullptr, 0u);
  ++g_launches;
  return true;
}

template <typename T>
int launch_adjust(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* af,
                  const T* hq, int nq, int interp, int extrap, int kind, T* scen, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !sim) || !grp || (n_pts > 0 && !af) || (n_pts > 0 && !hq) || (n_pts > 0 && !scen) || n_pts < 0 || nq <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (kind != XSDBA_KIND_ADD && kind != XSDBA_KIND_MUL) return XSDBA_ERR_INVALID_ARGUMENT;
  if (interp != XSDBA_INTERP_NEAREST && interp != XSDBA_INTERP_LINEAR && interp != XSDBA_INTERP_CUBIC) return XSDBA_ERR_INVALID_ARGUMENT;
  if (extrap != XSDBA_EXTRAP_CONSTANT && extrap != XSDBA_EXTRAP_NAN) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups > 1 && interp != XSDBA_INTERP_NEAREST) return XSDBA_ERR_UNSUPPORTED;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  const size_t smem = ((tables_bytes<T, 32>(nq) + 15) & ~(size_t)15) + stage_bytes<T, 32>(nq);
  if (smem > 200 * 1024) {
    // very fine quantile grids: fewer points per CTA so that the staged tables still fit
    if (launch_adjust_narrow<T, 16>(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, stream) ||
        launch_adjust_narrow<T, 8>(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, stream) ||
        launch_adjust_narrow<T, 4>(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, stream) ||
        launch_adjust_narrow<T, 2>(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, stream))
      return cuda_status(cudaGetLastError());
    return XSDBA_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)((n_pts + 31) / 32), (unsigned)grp->n_groups);
  if (launch_adjust_fast(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, smem, grid, (cudaStream_t)stream))
    return cuda_status(cudaGetLastError());
  auto kern = adjust_kernel<T, 32>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(sim, n_pts, sp, st, grp->members.off, grp->members.rows,
                                                       grp->n_groups, af, hq, nq, interp, extrap, kind, scen, nullptr, 0u);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T, int C>
int launch_rank_c(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                  const DevTable& seg, const T* af, const T* q, int nq, int interp, int extrap, int kind,
                  int do_adjust, T* scen, double* sim_q, int n_pad, cudaStream_t s, int rank_mode, const double* gcoord,
                  const unsigned char* diag, int slots) {
  const size_t head = (size_t)C * 28 + (((size_t)C * 28) % 8 ? 4 : 0);
  const size_t smem = head + (size_t)n_pad * C * sizeof(T) +
                      (do_adjust ? ((tables_bytes<T, C>(nq, slots) + 15) & ~(size_t)15) + stage_bytes<T, C>(nq)
                                 : tables_bytes<T, C>(0, slots));
  if (smem > 220 * 1024) return XSDBA_ERR_SEGMENT_TOO_LONG;
  auto kern = rank_kernel<T, C>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  dim3 grid((unsigned)((n_pts + C - 1) / C), (unsigned)grp->n_groups);
  kern<<<grid, threads_for_cols<C>(), smem, s>>>(sim, n_pts, sp, st, grp->members.off, grp->members.rows, seg.off,
                                                 seg.rows, grp->n_groups, af, q, do_adjust ? nq : 0, interp, extrap, kind,
                                                 do_adjust, scen, sim_q, n_pad, rank_mode, gcoord, diag, slots);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

// rank_window = True on groupings with heavily overlapping windows: one ordering per chunk of groups (K3w)
bool launch_rank_window(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const float* af,
                        const float* q, int nq, int extrap, int kind, int do_adjust, float* scen, double* sim_q,
                        cudaStream_t s, int* rc) {
  if (sp != 1 || st < 0 || grp->win.n_chunks <= 0 || !grp->win.mem_u || (do_adjust && nq > RankWinSmem::kWinMaxNq) ||
      getenv("XSDBA_B200_NO_WINDOW_KERNEL") || getenv("XSDBA_B200_NO_FAST"))
    return false;
  *rc = set_smem(rank_window_kernel, RankWinSmem::total);
  if (*rc) return true;
  dim3 grid((unsigned)((n_pts + kWinCols - 1) / kWinCols), (unsigned)grp->win.n_chunks);
  rank_window_kernel<<<grid, kWinThreads, RankWinSmem::total, s>>>(
      sim, n_pts, st, grp->members.off, grp->members.rows, grp->win.mem_u, grp->win.chunk_g, grp->win.urow_off,
      grp->win.urows, grp->win.gmask, grp->n_groups, af, q, nq, extrap, kind, do_adjust, scen, sim_q);
  ++g_launches;
  *rc = cuda_status(cudaGetLastError());
  return true;
}
bool launch_rank_window(const double*, int64_t, int64_t, int64_t, const xsdba_grouping*, const double*, const double*, int,
                        int, int, int, double*, double*, cudaStream_t, int*) {
  return false;
}

template <typename T>
int launch_rank(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* af,
                const T* q, int nq, int interp, int extrap, int kind, int rank_window, int do_adjust, T* scen,
                double* sim_q, void* stream, int rank_mode = 0, const double* gcoord = nullptr,
                const unsigned char* diag = nullptr) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !sim) || !grp || n_pts < 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (do_adjust) {
    if ((n_pts > 0 && !af) || (n_pts > 0 && !q) || (n_pts > 0 && !scen) || nq <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
    if (kind != XSDBA_KIND_ADD && kind != XSDBA_KIND_MUL) return XSDBA_ERR_INVALID_ARGUMENT;
    if (interp != XSDBA_INTERP_NEAREST && interp != XSDBA_INTERP_LINEAR && interp != XSDBA_INTERP_CUBIC) return XSDBA_ERR_INVALID_ARGUMENT;
    if (extrap != XSDBA_EXTRAP_CONSTANT && extrap != XSDBA_EXTRAP_NAN) return XSDBA_ERR_INVALID_ARGUMENT;
    if (grp->n_groups > 1 && interp == XSDBA_INTERP_CUBIC) return XSDBA_ERR_UNSUPPORTED;
    if (grp->n_groups > 1 && interp != XSDBA_INTERP_NEAREST && !(gcoord && diag)) return XSDBA_ERR_UNSUPPORTED;
  } else if (!sim_q) {
    return XSDBA_ERR_INVALID_ARGUMENT;
  }
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  {
    int wrc = 0;
    if (rank_window && rank_mode == 0 && grp->n_groups > 1 && (!do_adjust || interp == XSDBA_INTERP_NEAREST) &&
        launch_rank_window(sim, n_pts, sp, st, grp, af, q, nq, extrap, kind, do_adjust, scen, sim_q, (cudaStream_t)stream, &wrc))
      return wrc;
  }
  const DevTable& seg = rank_window ? grp->segments : grp->members;
  const int n_pad = std::max(2, next_pow2(seg.max_len));
  int C = pick_cols<T>(n_pad);
  // the linear rule interpolates between rows: it needs the three staged rows.  The nearest rule almost never
  // leaves the centre row (quantile nodes are < 1 apart, neighbouring rows are 1 away), so it stages that row only
  // and reads the neighbours from global memory in the rare case -- a third of the shared memory, 3x the occupancy
  const int slots = (do_adjust && ((grp->n_groups > 1 && interp == XSDBA_INTERP_LINEAR) || interp == XSDBA_INTERP_CUBIC)) ? 3 : 1;
  // the staged lookup tables share the CTA's shared memory with the sort buffer: narrow the tile until both fit
  auto need = [&](int c) {
    size_t top = 1;
    while (top * 2 <= (size_t)nq) top *= 2;
    const size_t tables = do_adjust ? ((size_t)2 * slots * 2 * top * c + 4 * c) * sizeof(T) + 3 * c * sizeof(int) + 16 +
                                          (size_t)2 * c * (nq | 1) * sizeof(T)
                                    : 0;
    return (size_t)c * 28 + 8 + (size_t)n_pad * c * sizeof(T) + tables;
  };
  while (C > 1 && need(C) > 216 * 1024) C >>= 1;
  cudaStream_t s = (cudaStream_t)stream;
#define XS_CASE(CC) case CC: return launch_rank_c<T, CC>(sim, n_pts, sp, st, grp, seg, af, q, nq, interp, extrap, kind, do_adjust, scen, sim_q, n_pad, s, rank_mode, gcoord, diag, slots)
  switch (C) {
    XS_CASE(32); XS_CASE(16); XS_CASE(8); XS_CASE(4); XS_CASE(2); XS_CASE(1);
    default: return XSDBA_ERR_SEGMENT_TOO_LONG;
  }
#undef XS_CASE
}

template <typename T>
int launch_poly_trend(const T* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* scaling,
                      int kind, int degree, const double* tcoord, double* trend, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !x) || !grp || (n_pts > 0 && !tcoord) || (n_pts > 0 && !trend) || n_pts < 0 || degree < 0 || degree > kMaxDeg) return XSDBA_ERR_INVALID_ARGUMENT;
  if (kind != XSDBA_KIND_ADD && kind != XSDBA_KIND_MUL) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  dim3 grid((unsigned)((n_pts + 31) / 32), (unsigned)grp->n_groups);
  poly_trend_kernel<T><<<grid, kThreads, 0, (cudaStream_t)stream>>>(
      x, n_pts, sp, st, grp->members.off, grp->members.rows, grp->segments.off, grp->segments.rows, grp->gidx,
      grp->n_groups, grp->window, scaling, kind, degree, tcoord, trend);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_dqm_adjust(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* af,
                      const T* hq, const T* scaling, const double* trend, int nq, int interp, int extrap, int kind,
                      T* scen, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !sim) || !grp || (n_pts > 0 && !af) || (n_pts > 0 && !hq) || (n_pts > 0 && !scaling) || (n_pts > 0 && !trend) || (n_pts > 0 && !scen) || n_pts < 0 || nq <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (kind != XSDBA_KIND_ADD && kind != XSDBA_KIND_MUL) return XSDBA_ERR_INVALID_ARGUMENT;
  if (interp != XSDBA_INTERP_NEAREST && interp != XSDBA_INTERP_LINEAR && interp != XSDBA_INTERP_CUBIC) return XSDBA_ERR_INVALID_ARGUMENT;
  if (extrap != XSDBA_EXTRAP_CONSTANT && extrap != XSDBA_EXTRAP_NAN) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups > 1 && interp != XSDBA_INTERP_NEAREST) return XSDBA_ERR_UNSUPPORTED;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  const size_t smem = ((tables_bytes<T, 32>(nq) + 15) & ~(size_t)15) + stage_bytes<T, 32>(nq);
  if (smem > 200 * 1024) return XSDBA_ERR_UNSUPPORTED;
  auto kern = dqm_adjust_kernel<T>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  dim3 grid((unsigned)((n_pts + 31) / 32), (unsigned)grp->n_groups);
  kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(sim, n_pts, sp, st, grp->members.off, grp->members.rows,
                                                       grp->n_groups, af, hq, scaling, trend, nq, interp, extrap, kind, scen);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

// K6t: interior outputs (degree 0, first iteration) of a tile of 32 COMPLETE series.  K6b lets every warp stream its own
// 2 HW + 1 taps from L2; here a CTA of 8 warps x 16 outputs = 128 consecutive outputs shares them: the union of the
// windows goes through shared memory in tiles of 128 taps (loaded once per CTA, coalesced), the weights -- one vector
// for all complete series -- sit in shared memory and are read as broadcasts, and each thread (lane = point) keeps 16
// accumulators and a 16-slot weight ring in registers: 16 DFMA per tap and thread against 2 shared-memory loads.
// Same summation order as K6b (taps ascending), same denominator (total weight in tap order).
constexpr int kLoessTileOut = 16;    // outputs per warp
constexpr int kLoessTileTaps = 128;  // taps per shared-memory tile
template <typename T>
__global__ void __launch_bounds__(kThreads)
loess_interior_tile_kernel(const T* __restrict__ yc, const int32_t* __restrict__ nvalid, long long n_pts, long long sp,
                           long long st, int n_time, double f, const double* __restrict__ wsh, int w_rows,
                           double* __restrict__ trend) {
  constexpr int RO = kLoessTileOut, TT = kLoessTileTaps;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* wsm = reinterpret_cast<double*>(smem_raw);                  // [w_rows + RO] (zero beyond the last weight)
  T* ytile = reinterpret_cast<T*>(wsm + w_rows + RO);                 // [TT][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = (long long)blockIdx.x * 32 + lane;
  if (!__all_sync(0xffffffffu, pt < n_pts && nvalid[pt < n_pts ? pt : 0] == n_time)) return;  // K6b does this tile
  const int n = n_time;
  const LoessGeom gm(n, f);
  const int K = 2 * gm.HW;
  // interior outputs (LoessGeom::interior) plus i = n-HW-1, which keeps the interior weights on the window
  // [n-R, n) = [i-HW, i+HW] (loess.py:128-150: right-hand window, weights not recomputed)
  const int first = gm.HW + 1, last = n - gm.HW - 1;
  const int o_base = first + blockIdx.y * (n_warps * RO);
  if (o_base > last) return;
  for (int k = threadIdx.x; k < w_rows + RO; k += blockDim.x) wsm[k] = k <= K && k < w_rows ? wsh[k] : 0.0;
  const double sw_total = wsh[w_rows];
  const int i0 = o_base + warp * RO;                   // this warp's outputs i0 .. i0 + RO - 1
  const int jb = o_base - gm.HW;                       // first tap of the CTA's window union
  const int je = min(o_base + n_warps * RO - 1, last) + gm.HW;
  const T* y = yc + pt;
  double swy[RO], wr[RO];
#pragma unroll
  for (int r = 0; r < RO; ++r) { swy[r] = 0; wr[r] = 0; }
  for (int t0 = jb; t0 <= je; t0 += TT) {
    __syncthreads();  // the previous tile has been consumed (and wsm is complete before the first use)
    for (int tt = warp; tt < TT; tt += n_warps) ytile[tt * 32 + lane] = y[(long long)min(t0 + tt, n - 1) * n_pts];
    __syncthreads();
    // weight index of the tile's first tap for output i0: a multiple of RO by construction
    const int m0 = t0 - (i0 - gm.HW);
#pragma unroll 1
    for (int b = 0; b < TT; b += RO) {
      const int m = m0 + b;
      if (m < 0 || m > K + RO - 1) continue;           // (warp-uniform) outside this warp's windows
#pragma unroll
      for (int u = 0; u < RO; ++u) {
        wr[u] = wsm[min(m + u, w_rows + RO - 1)];
        const double yj = (double)ytile[(b + u) * 32 + lane];
#pragma unroll
        for (int r = 0; r < RO; ++r) swy[r] = fma(wr[(u - r + RO) % RO], yj, swy[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RO; ++r)
    if (i0 + r <= last) trend[pt * sp + (long long)(i0 + r) * st] = swy[r] / sw_total;
}

// K6u: edge outputs (degree 0, first iteration) of a tile of 32 COMPLETE series.  The HW+1 left (HW right) outputs share
// the window [0, R) ([n-R, n)); their weights come from the (tap, output) table of K6e.  Same structure as K6t: 8 warps
// x 16 outputs per CTA, the window staged through shared memory in 128-tap tiles, 16 accumulators per thread; the 16
// weights of a tap are warp-uniform loads of 128 contiguous bytes of the L2-resident table.
template <typename T>
__global__ void __launch_bounds__(kThreads)
loess_edge_tile_kernel(const T* __restrict__ yc, const int32_t* __restrict__ nvalid, long long n_pts, long long sp,
                       long long st, int n_time, double f, const double* __restrict__ etab,
                       const double* __restrict__ esum, int chunks_per_side, double* __restrict__ trend) {
  constexpr int RO = kLoessTileOut, TT = kLoessTileTaps;
  __shared__ T ytile[TT * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const long long pt = (long long)blockIdx.x * 32 + lane;
  if (!__all_sync(0xffffffffu, pt < n_pts && nvalid[pt < n_pts ? pt : 0] == n_time)) return;  // K6b does this tile
  const int n = n_time;
  const LoessGeom gm(n, f);
  const int NI = 2 * gm.HW + 1;
  const int side = blockIdx.y / chunks_per_side, chunk = blockIdx.y % chunks_per_side;
  const int n_edge = side == 0 ? gm.HW + 1 : gm.HW;          // outputs on this side
  const int e_base = side == 0 ? 0 : gm.HW + 1;              // their first table column
  const int i_base = side == 0 ? 0 : n - gm.HW;              // their first output index
  const int lo = side == 0 ? 0 : n - gm.R;                   // the shared window
  const int o0 = chunk * (n_warps * RO) + warp * RO;         // this warp's first output on the side
  if (chunk * (n_warps * RO) >= n_edge) return;
  const double* et = etab + e_base + min(o0, n_edge - 1);    // (warps past the end compute garbage they never write)
  const int span = min(RO, max(n_edge - o0, 1));             // table columns this warp may read
  const T* y = yc + pt + (long long)lo * n_pts;
  double swy[RO];
#pragma unroll
  for (int r = 0; r < RO; ++r) swy[r] = 0;
  for (int t0 = 0; t0 < gm.R; t0 += TT) {
    __syncthreads();
    for (int tt = warp; tt < TT; tt += n_warps) ytile[tt * 32 + lane] = y[(long long)min(t0 + tt, gm.R - 1) * n_pts];
    __syncthreads();
    const int nt = min(TT, gm.R - t0);
    for (int tt = 0; tt < nt; ++tt) {
      const double* row = et + (long long)(t0 + tt) * NI;
      const double yj = (double)ytile[tt * 32 + lane];
#pragma unroll
      for (int r = 0; r < RO; ++r) swy[r] = fma(row[r < span ? r : 0], yj, swy[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < RO; ++r) {
    const int o = o0 + r;
    if (o < n_edge) trend[pt * sp + (long long)(i_base + o) * st] = swy[r] / esum[e_base + o];
  }
}

// K6d: robustness weights between two LOESS iterations (loess.py:166-176): residuals of the compacted series,
// s = median(|residuals|) (mean of the two middle values for an even count), xres = residuals / (6 s) -- or the
// indicator of a non-zero residual when s == 0 -- delta = (1 - xres^2)^2, 0 where |xres| >= 1.
// One CTA per point; |residuals| are sorted in shared memory (n_pad doubles).
template <typename T>
__global__ void __launch_bounds__(kThreads)
loess_delta_kernel(const T* __restrict__ yc, const int32_t* __restrict__ tc, const int32_t* __restrict__ nvalid,
                   long long n_pts, long long sp, long long st, int n_pad, const double* __restrict__ trend,
                   double* __restrict__ delta) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* a = reinterpret_cast<double*>(smem_raw);
  const long long pt = blockIdx.x;
  const int n = nvalid[pt];
  if (n == 0) return;
  const double dinf = __longlong_as_double(0x7ff0000000000000LL);
  for (int j = threadIdx.x; j < n_pad; j += blockDim.x) {
    double v = dinf;
    if (j < n) {
      v = fabs((double)yc[(long long)j * n_pts + pt] - trend[pt * sp + (long long)tc[(long long)j * n_pts + pt] * st]);
      if (v != v) v = dinf;
    }
    a[j] = v;
  }
  __syncthreads();
  sort_columns<double, 1>(a, n_pad);
  const double s = (n & 1) ? a[n >> 1] : (a[(n >> 1) - 1] + a[n >> 1]) / 2.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double res = (double)yc[(long long)j * n_pts + pt] - trend[pt * sp + (long long)tc[(long long)j * n_pts + pt] * st];
    const double xres = s == 0.0 ? (res != 0.0 ? 1.0 : 0.0) : res / (6.0 * s);
    const double c = 1.0 - xres * xres;
    delta[(long long)j * n_pts + pt] = fabs(xres) >= 1.0 ? 0.0 : c * c;
  }
}

template <typename T>
int launch_loess_trend(const T* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* scaling,
                       int kind, double f, int niter, int degree, const double* xn, double* trend, void* stream,
                       int unequal = 0) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !x) || !grp || (n_pts > 0 && !xn) || (n_pts > 0 && !trend) || n_pts < 0 || !(f != 0.0) || degree < 0 || degree > 1) return XSDBA_ERR_INVALID_ARGUMENT;
  if (niter < 1) return XSDBA_ERR_INVALID_ARGUMENT;
  if (kind != XSDBA_KIND_ADD && kind != XSDBA_KIND_MUL) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n_pts == 0) return XSDBA_OK;
  keep_pool_memory();
  cudaStream_t s = (cudaStream_t)stream;
  const int n_time = (int)grp->n_time;
  T* yc = nullptr; int32_t* tc = nullptr; int32_t* nv = nullptr;
  if (cudaMallocAsync(&yc, sizeof(T) * n_pts * n_time, s) != cudaSuccess ||
      cudaMallocAsync(&tc, sizeof(int32_t) * n_pts * n_time, s) != cudaSuccess ||
      cudaMallocAsync(&nv, sizeof(int32_t) * n_pts, s) != cudaSuccess) {
    cudaGetLastError();
    if (yc) cudaFreeAsync(yc, s);
    if (tc) cudaFreeAsync(tc, s);
    return XSDBA_ERR_OUT_OF_MEMORY;
  }
  loess_compact_kernel<T><<<(unsigned)((n_pts + kThreads - 1) / kThreads), kThreads, 0, s>>>(
      x, n_pts, sp, st, n_time, grp->gidx, grp->n_groups, scaling, kind, yc, tc, nv, trend);
  if (unequal) {   // dx == 0 (loess.py:257-260): every output recomputes its bandwidth and weights, no tables
    double* delta_u = nullptr;
    int rc_u = XSDBA_OK;
    if (niter > 1) {
      const int n_pad = std::max(2, next_pow2(n_time));
      if ((size_t)n_pad * sizeof(double) > 200 * 1024) rc_u = XSDBA_ERR_SEGMENT_TOO_LONG;
      else if (cudaMallocAsync(&delta_u, sizeof(double) * n_pts * n_time, s) != cudaSuccess) { cudaGetLastError(); rc_u = XSDBA_ERR_OUT_OF_MEMORY; }
      else rc_u = set_smem(loess_delta_kernel<T>, (size_t)n_pad * sizeof(double));
    }
    const unsigned chunks_u = (unsigned)std::min<int64_t>(std::max<int64_t>(1, (n_time + 7) / 8), 4096);
    for (int it = 0; it < niter && rc_u == XSDBA_OK; ++it) {
      loess_unequal_kernel<T><<<dim3((unsigned)((n_pts + 31) / 32), chunks_u), kThreads, 0, s>>>(
          yc, tc, nv, n_pts, sp, st, xn, f, degree, it > 0 ? delta_u : nullptr, trend);
      ++g_launches;
      if (it + 1 < niter) {
        const int n_pad = std::max(2, next_pow2(n_time));
        loess_delta_kernel<T><<<(unsigned)n_pts, kThreads, (size_t)n_pad * sizeof(double), s>>>(yc, tc, nv, n_pts, sp, st, n_pad,
                                                                                                trend, delta_u);
        ++g_launches;
      }
    }
    ++g_launches;
    if (delta_u) cudaFreeAsync(delta_u, s);
    cudaFreeAsync(yc, s); cudaFreeAsync(tc, s); cudaFreeAsync(nv, s);
    if (rc_u != XSDBA_OK) return rc_u;
    return cuda_status(cudaGetLastError());
  }
  // interior weight table: at most 2*HW+1 <= f*n_time + 6 rows per point
  const double fa = std::fabs(f);  // (f < 0: gaussian weights, see LoessGeom)
  const int w_rows = (int)std::min<double>((double)n_time, fa * (double)n_time + 8.0);
  double* wtab = nullptr;
  if (cudaMallocAsync(&wtab, sizeof(double) * n_pts * (w_rows + 1), s) != cudaSuccess) {
    cudaGetLastError();
    cudaFreeAsync(yc, s); cudaFreeAsync(tc, s); cudaFreeAsync(nv, s);
    return XSDBA_ERR_OUT_OF_MEMORY;
  }
  loess_weights_kernel<<<dim3((unsigned)((n_pts + 31) / 32), 16), kThreads, 0, s>>>(tc, nv, n_pts, n_time, xn, f, w_rows, wtab);
  double* wsh = nullptr;
  if (cudaMallocAsync(&wsh, sizeof(double) * (w_rows + 1), s) == cudaSuccess) {
    loess_shared_weights_kernel<<<8, kThreads, 0, s>>>(xn, n_time, f, w_rows, wsh);
    ++g_launches;
  } else {
    cudaGetLastError();
    wsh = nullptr;
  }
  loess_wsum_kernel<<<(unsigned)((n_pts + 1 + 127) / 128), 128, 0, s>>>(nv, n_pts, n_time, f, w_rows, wtab, wsh);
  ++g_launches;
  // edge weights of complete series (optional: without the table the kernel recomputes them per point)
  double* etab = nullptr;
  double* esum = nullptr;
  if (degree == 0 && n_time >= 8) {
    const int r_ = (int)(2.0 * std::floor(fa * (double)n_time / 2.0) + 1.0);   // LoessGeom on the host
    const int HW_ = (r_ - 1) / 2 + 2, R_ = std::min(r_ + 4, n_time), NI_ = 2 * HW_ + 1;
    if (n_time >= R_ && n_time > 2 * HW_ + 2 &&
        cudaMallocAsync(&etab, sizeof(double) * (size_t)R_ * NI_ + sizeof(double) * NI_, s) == cudaSuccess) {
      esum = etab + (size_t)R_ * NI_;
      loess_edge_table_kernel<<<sm_count() * 8, kThreads, 0, s>>>(xn, n_time, f, etab);
      loess_edge_sum_kernel<<<(unsigned)((NI_ + 127) / 128), 128, 0, s>>>(n_time, f, etab, esum);
      g_launches += 2;
    } else {
      cudaGetLastError();
      etab = nullptr;
    }
  }
  const unsigned chunks = (unsigned)std::min<int64_t>(std::max<int64_t>(1, (n_time + 63) / 64), 1024);
  double* delta = nullptr;
  int rc_iter = XSDBA_OK;
  if (niter > 1) {
    const int n_pad = std::max(2, next_pow2(n_time));
    auto dk = loess_delta_kernel<T>;
    if ((size_t)n_pad * sizeof(double) > 200 * 1024) rc_iter = XSDBA_ERR_SEGMENT_TOO_LONG;
    else if (cudaMallocAsync(&delta, sizeof(double) * n_pts * n_time, s) != cudaSuccess) {
      cudaGetLastError();
      delta = nullptr;
      rc_iter = XSDBA_ERR_OUT_OF_MEMORY;
    } else rc_iter = set_smem(dk, (size_t)n_pad * sizeof(double));
  }
  // K6t for the interior of complete tiles (first iteration, degree 0); K6b does everything else
  int tiled = 0;
  const size_t smem_t = sizeof(double) * (w_rows + kLoessTileOut) + sizeof(T) * kLoessTileTaps * 32;
  if (degree == 0 && wsh && etab && smem_t <= 200 * 1024 && set_smem(loess_interior_tile_kernel<T>, smem_t) == XSDBA_OK &&
      !getenv("XSDBA_B200_NO_LOESS_TILE"))
    tiled = 1;
  for (int it = 0; it < niter && rc_iter == XSDBA_OK; ++it) {
    const int tiled_now = (it == 0) ? tiled : 0;
    if (tiled_now) {
      const int per_cta = (kThreads / 32) * kLoessTileOut;
      loess_interior_tile_kernel<T><<<dim3((unsigned)((n_pts + 31) / 32), (unsigned)((n_time + per_cta - 1) / per_cta)),
                                      kThreads, smem_t, s>>>(yc, nv, n_pts, sp, st, n_time, f, wsh, w_rows, trend);
      ++g_launches;
      {
        const int r_ = (int)(2.0 * std::floor(fa * (double)n_time / 2.0) + 1.0);
        const int HW_ = (r_ - 1) / 2 + 2;
        const int cps = (HW_ + 1 + per_cta - 1) / per_cta;
        loess_edge_tile_kernel<T><<<dim3((unsigned)((n_pts + 31) / 32), (unsigned)(2 * cps)), kThreads, 0, s>>>(
            yc, nv, n_pts, sp, st, n_time, f, etab, esum, cps, trend);
        ++g_launches;
      }
    }
    loess_smooth_kernel<T><<<dim3((unsigned)((n_pts + 31) / 32), chunks), kThreads, 0, s>>>(
        yc, tc, nv, n_pts, sp, st, n_time, xn, f, degree, wtab, w_rows, wsh, etab, esum, it > 0 ? delta : nullptr,
        tiled_now, trend);
    ++g_launches;
    if (it + 1 < niter) {
      const int n_pad = std::max(2, next_pow2(n_time));
      loess_delta_kernel<T><<<(unsigned)n_pts, kThreads, (size_t)n_pad * sizeof(double), s>>>(yc, tc, nv, n_pts, sp, st, n_pad,
                                                                                              trend, delta);
      ++g_launches;
    }
  }
  g_launches += 2;
  if (delta) cudaFreeAsync(delta, s);
  cudaFreeAsync(yc, s); cudaFreeAsync(tc, s); cudaFreeAsync(nv, s); cudaFreeAsync(wtab, s);
  if (etab) cudaFreeAsync(etab, s);
  if (wsh) cudaFreeAsync(wsh, s);
  if (rc_iter != XSDBA_OK) return rc_iter;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_jitter(const T* x, int64_t n, const double* j4, uint64_t seed, T* out, void* stream) {
  if ((n > 0 && !x) || (n > 0 && !out) || (n > 0 && !j4) || n < 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n == 0) return XSDBA_OK;
  JitterParams jp{j4[0], j4[1], j4[2], j4[3], seed};
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  jitter_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, jp, out);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}
template <typename T, int C>
int launch_reorder_c(const T* sim, const T* ref, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                     T* out, int n_pad, cudaStream_t s) {
  const size_t smem = (size_t)C * 16 + (size_t)2 * n_pad * C * sizeof(T);
  auto kern = reorder_kernel<T, C>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  dim3 grid((unsigned)((n_pts + C - 1) / C), (unsigned)grp->n_groups);
  kern<<<grid, kThreads, smem, s>>>(sim, ref, n_pts, sp, st, grp->members.off, grp->members.rows, grp->segments.off,
                                    grp->segments.rows, grp->window, n_pad, out);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_reorder(const T* sim, const T* ref, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, T* out,
                   void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !sim) || (n_pts > 0 && !ref) || !grp || (n_pts > 0 && !out) || n_pts < 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  const int n_pad = std::max(2, next_pow2(grp->segments.max_len));
  int C = pick_cols<T>(2 * n_pad);  // two sorted arrays share the budget
  cudaStream_t s = (cudaStream_t)stream;
#define XS_CASE(CC) case CC: return launch_reorder_c<T, CC>(sim, ref, n_pts, sp, st, grp, out, n_pad, s)
  switch (C) {
    XS_CASE(32); XS_CASE(16); XS_CASE(8); XS_CASE(4); XS_CASE(2); XS_CASE(1);
    default: return XSDBA_ERR_SEGMENT_TOO_LONG;
  }
#undef XS_CASE
}

template <typename T, int C>
int launch_select_c(const T* x, const T* y, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, int mode,
                    const T* rnk, const double* yvals, int nv, T* out, int n_pad, cudaStream_t s) {
  const size_t smem = (size_t)C * 16 + (size_t)n_pad * C * sizeof(T);
  auto kern = select_kernel<T, C>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  dim3 grid((unsigned)((n_pts + C - 1) / C), (unsigned)grp->n_groups);
  kern<<<grid, kThreads, smem, s>>>(x, y, n_pts, sp, st, grp->segments.off, grp->segments.rows, grp->n_groups, mode, rnk,
                                    yvals, nv, out, n_pad);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_select(const T* x, const T* y, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, int mode,
                  const T* rnk, const double* yvals, int nv, T* out, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !x) || !grp || (n_pts > 0 && !out) || n_pts < 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (mode == 0 && !rnk) return XSDBA_ERR_INVALID_ARGUMENT;
  if (mode == 1 && (!y || !yvals || nv <= 0)) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  const int n_pad = std::max(2, next_pow2(grp->segments.max_len));
  const int C = pick_cols<T>(n_pad);
  cudaStream_t s = (cudaStream_t)stream;
#define XS_CASE(CC) case CC: return launch_select_c<T, CC>(x, y, n_pts, sp, st, grp, mode, rnk, yvals, nv, out, n_pad, s)
  switch (C) {
    XS_CASE(32); XS_CASE(16); XS_CASE(8); XS_CASE(4); XS_CASE(2); XS_CASE(1);
    default: return XSDBA_ERR_SEGMENT_TOO_LONG;
  }
#undef XS_CASE
}

template <typename T, int C>
int launch_adapt_apply_c(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, double thresh,
                         const double* p0r, const double* p0h, const T* pth, unsigned long long seed, T* out, int n_pad,
                         cudaStream_t s) {
  const size_t smem = (size_t)C * 16 + (size_t)n_pad * C * sizeof(T);
  auto kern = adapt_apply_kernel<T, C>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  dim3 grid((unsigned)((n_pts + C - 1) / C), (unsigned)grp->n_groups);
  kern<<<grid, kThreads, smem, s>>>(sim, n_pts, sp, st, grp->members.off, grp->members.rows, grp->n_groups, thresh, p0r,
                                    p0h, pth, seed, out, n_pad);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_adapt_apply(const T* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, double thresh,
                       const double* p0r, const double* p0h, const T* pth, unsigned long long seed, T* out, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !sim) || !grp || (n_pts > 0 && !p0r) || (n_pts > 0 && !p0h) || (n_pts > 0 && !pth) || (n_pts > 0 && !out) || n_pts < 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups > 65535) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  const int n_pad = std::max(2, next_pow2(grp->members.max_len));
  const int C = pick_cols<T>(n_pad);
  cudaStream_t s = (cudaStream_t)stream;
#define XS_CASE(CC) case CC: return launch_adapt_apply_c<T, CC>(sim, n_pts, sp, st, grp, thresh, p0r, p0h, pth, seed, out, n_pad, s)
  switch (C) {
    XS_CASE(32); XS_CASE(16); XS_CASE(8); XS_CASE(4); XS_CASE(2); XS_CASE(1);
    default: return XSDBA_ERR_SEGMENT_TOO_LONG;
  }
#undef XS_CASE
}

template <typename T>
int launch_tail_mask(const T* adapted, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp, const T* hq_raw,
                     int nq, double factor, T* scen, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;  // the handle's tables live on another device
  if ((n_pts > 0 && !adapted) || !grp || (n_pts > 0 && !hq_raw) || (n_pts > 0 && !scen) || n_pts < 0 || nq <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (sp != 1 && st != 1) return XSDBA_ERR_UNSUPPORTED;
  if (n_pts == 0) return XSDBA_OK;
  const int64_t total = n_pts * grp->n_time;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 32);
  tail_mask_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(adapted, n_pts, sp, st, (int)grp->n_time, grp->gidx,
                                                                grp->n_groups, hq_raw, nq, factor, scen);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_escore(const T* tgt, const T* sim, int64_t n_pts, int64_t sp, int64_t st, int64_t nt_t, int64_t nt_s, int n_var,
                  int64_t vs_t, int64_t vs_s, int n_sub, T* out, void* stream) {
  if ((n_pts > 0 && !tgt) || (n_pts > 0 && !sim) || (n_pts > 0 && !out) || n_pts < 0 || n_var < 1 || n_var > kMaxVar || nt_t <= 0 || nt_s <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n_pts == 0) return XSDBA_OK;
  // N > 0: about N evenly spaced observations of each cloud (processing.py:459-464)
  const int step_t = n_sub > 0 ? (int)((nt_t + n_sub - 1) / n_sub) : 1;
  const int step_s = n_sub > 0 ? (int)((nt_s + n_sub - 1) / n_sub) : 1;
  const size_t smem = sizeof(int) * ((nt_t + step_t - 1) / step_t + (nt_s + step_s - 1) / step_s);
  if (smem > 200 * 1024) return XSDBA_ERR_SEGMENT_TOO_LONG;
  auto kern = escore_kernel<T>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  kern<<<(unsigned)n_pts, kThreads, smem, (cudaStream_t)stream>>>(tgt, sim, n_pts, sp, st, (int)nt_t, (int)nt_s, step_t,
                                                                  step_s, n_var, vs_t, vs_s, out);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_rotate(const T* x, int64_t n_elem, int n_var, const float* rot_host, T* y, void* stream, bool fused = true) {
  if ((n_elem > 0 && !x) || (n_elem > 0 && !y) || (n_elem > 0 && !rot_host) || n_var < 1 || n_var > kMaxVar || n_elem < 0 || (n_elem > 0 && x == y)) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n_elem == 0) return XSDBA_OK;
  RotMat R;
  for (int v = 0; v < kMaxVar; ++v) for (int w = 0; w < kMaxVar; ++w)
    R.r[v * kMaxVar + w] = (v < n_var && w < n_var) ? rot_host[v * n_var + w] : 0.f;
  const unsigned blocks = (unsigned)std::min<int64_t>((n_elem + 255) / 256, (int64_t)sm_count() * 32);
  if (fused) rotate_kernel<T, true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n_elem, n_var, R, y);
  else rotate_kernel<T, false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n_elem, n_var, R, y);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <typename T>
int launch_standardize(const T* x, int64_t n_pts, int64_t sp, int64_t st, int64_t n_time, int n_var, int64_t var_stride,
                       T* y, void* stream) {
  if ((n_pts > 0 && !x) || (n_pts > 0 && !y) || n_pts < 0 || n_time <= 0 || n_var < 1) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n_pts == 0) return XSDBA_OK;
  const int64_t n = n_pts * n_var;
  standardize_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, n_pts, sp, st, (int)n_time,
                                                                                       n_var, var_stride, y);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

template <int C>
int launch_npdft_step_c(const float* ref, float* hist, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                        const double* q64, int nq, int interp, int extrap, float* af_io, int n_pad, cudaStream_t s) {
  const size_t smem = (size_t)C * 24 + ((size_t)nq * C + (size_t)n_pad * C) * sizeof(float) + 8 +
                      ((tables_bytes<double, C>(nq, 1) + 15) & ~(size_t)15);
  if (smem > 220 * 1024) return XSDBA_ERR_SEGMENT_TOO_LONG;
  auto kern = npdft_step_kernel<C>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  kern<<<(unsigned)((n_pts + C - 1) / C), threads_for_cols<C>(), smem, s>>>(ref, hist, n_pts, sp, st, grp->segments.rows,
                                                                             (int)grp->segments.total, q64, nq, interp,
                                                                             extrap, af_io, n_pad);
  ++g_launches;
  return cuda_status(cudaGetLastError());
}

int launch_npdft_step(const float* ref, float* hist, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping* grp,
                      const double* q64, int nq, int interp, int extrap, float* af_io, void* stream) {
  if (wrong_device(grp)) return XSDBA_ERR_INVALID_ARGUMENT;
  if ((n_pts > 0 && !hist) || !grp || (n_pts > 0 && !q64) || (n_pts > 0 && !af_io) || n_pts < 0 || nq <= 0) return XSDBA_ERR_INVALID_ARGUMENT;
  if (grp->n_groups != 1 || grp->window != 1) return XSDBA_ERR_INVALID_ARGUMENT;  // a block of time steps is one group
  if (interp != XSDBA_INTERP_NEAREST && interp != XSDBA_INTERP_LINEAR) return XSDBA_ERR_INVALID_ARGUMENT;
  if (extrap != XSDBA_EXTRAP_CONSTANT && extrap != XSDBA_EXTRAP_NAN) return XSDBA_ERR_INVALID_ARGUMENT;
  if (n_pts == 0) return XSDBA_OK;
  const int n_pad = std::max(2, next_pow2(grp->segments.max_len));
  cudaStream_t s = (cudaStream_t)stream;
  // columns per CTA: the keys plus the float64 tables must fit
  for (int C = 32; C >= 1; C >>= 1) {
    const size_t need = (size_t)C * 24 + ((size_t)nq * C + (size_t)n_pad * C) * sizeof(float) + 8 +
                        (size_t)(2 * 2 * 128 * C + 4 * C) * 8 + 64;
    if ((size_t)n_pad * C * sizeof(float) > (size_t)kSortBytes || need > 200 * 1024) continue;
    switch (C) {
      case 32: return launch_npdft_step_c<32>(ref, hist, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, n_pad, s);
      case 16: return launch_npdft_step_c<16>(ref, hist, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, n_pad, s);
      case 8: return launch_npdft_step_c<8>(ref, hist, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, n_pad, s);
      case 4: return launch_npdft_step_c<4>(ref, hist, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, n_pad, s);
      case 2: return launch_npdft_step_c<2>(ref, hist, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, n_pad, s);
      default: return launch_npdft_step_c<1>(ref, hist, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, n_pad, s);
    }
  }
  return XSDBA_ERR_SEGMENT_TOO_LONG;
}

int upload_table(const std::vector<int32_t>& off, const std::vector<int32_t>& rows, DevTable& t) {
  XS_CUDA(cudaMalloc(&t.off, off.size() * sizeof(int32_t)));
  XS_CUDA(cudaMalloc(&t.rows, std::max<size_t>(rows.size(), 1) * sizeof(int32_t)));
  XS_CUDA(cudaMemcpy(t.off, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (!rows.empty()) XS_CUDA(cudaMemcpy(t.rows, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  t.total = (int64_t)rows.size();
  t.max_len = 0;
  for (size_t g = 0; g + 1 < off.size(); ++g) t.max_len = std::max(t.max_len, off[g + 1] - off[g]);
  return XSDBA_OK;
}

// Greedy chunking of consecutive groups for K1w.  Leaves win.n_chunks == 0 (kernel not used) when a group repeats a
// row, a single group exceeds the row budget, or the sharing is below 2x.
int build_window_tables(const std::vector<int32_t>& soff, const std::vector<int32_t>& srows, int32_t n_time, int n_groups,
                        const std::vector<int32_t>& moff, const std::vector<int32_t>& mrows, WinTables& win) {
  std::vector<int32_t> stamp(n_time, -1), local(n_time, 0);
  std::vector<int32_t> mem_u(mrows.size(), -1);
  bool members_inside = true;  // every exact member is a row of its own window (the centre slot)
  std::vector<int32_t> chunk_g{0}, urow_off{0}, urows;
  std::vector<unsigned long long> gmask;
  std::vector<int32_t> cur;  // rows of the open chunk
  int chunk_id = 0, groups_in = 0, max_union = 0, max_groups = 0;
  auto close_chunk = [&](int next_g) {
    std::sort(cur.begin(), cur.end());
    for (size_t i = 0; i < cur.size(); ++i) local[cur[i]] = (int32_t)i;
    const size_t base = gmask.size();
    gmask.resize(base + cur.size(), 0ull);
    for (int g = chunk_g.back(); g < next_g; ++g)
      for (int s_ = soff[g]; s_ < soff[g + 1]; ++s_)
        if (srows[s_] >= 0) gmask[base + local[srows[s_]]] |= 1ull << (g - chunk_g.back());
    for (int g = chunk_g.back(); g < next_g; ++g)
      for (int m = moff[g]; m < moff[g + 1]; ++m) {
        if (stamp[mrows[m]] == chunk_id) mem_u[m] = local[mrows[m]];
        else members_inside = false;
      }
    urows.insert(urows.end(), cur.begin(), cur.end());
    urow_off.push_back((int32_t)urows.size());
    chunk_g.push_back(next_g);
    max_union = std::max<int>(max_union, (int)cur.size());
    max_groups = std::max(max_groups, groups_in);
    cur.clear(); groups_in = 0; ++chunk_id;
  };
  std::vector<int32_t> seen(n_time, -1);
  for (int g = 0; g < n_groups; ++g) {
    int fresh = 0;
    for (int s_ = soff[g]; s_ < soff[g + 1]; ++s_) {
      const int32_t t = srows[s_];
      if (t < 0) continue;
      if (seen[t] == g) return XSDBA_OK;  // a row twice in one window: not a case for the bitmap selection
      seen[t] = g;
      if (stamp[t] != chunk_id) ++fresh;
    }
    int own = 0;
    for (int s_ = soff[g]; s_ < soff[g + 1]; ++s_) own += srows[s_] >= 0 ? 1 : 0;
    if (own > kWinMaxRows) return XSDBA_OK;
    if (groups_in > 0 && ((int)cur.size() + fresh > kWinMaxRows || groups_in >= kWinMaxGroups)) close_chunk(g);
    for (int s_ = soff[g]; s_ < soff[g + 1]; ++s_) {
      const int32_t t = srows[s_];
      if (t >= 0 && stamp[t] != chunk_id) { stamp[t] = chunk_id; cur.push_back(t); }
    }
    ++groups_in;
  }
  if (groups_in > 0) close_chunk(n_groups);
  if (urows.size() * 2 > srows.size()) return XSDBA_OK;  // less than 2x sharing: the per-group kernels are as good
  cudaError_t e = cudaMalloc(&win.chunk_g, chunk_g.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&win.urow_off, urow_off.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&win.urows, std::max<size_t>(urows.size(), 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&win.gmask, std::max<size_t>(gmask.size(), 1) * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpy(win.chunk_g, chunk_g.data(), chunk_g.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(win.urow_off, urow_off.data(), urow_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(win.urows, urows.data(), urows.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(win.gmask, gmask.data(), gmask.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && members_inside && !mem_u.empty()) {
    e = cudaMalloc(&win.mem_u, mem_u.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(win.mem_u, mem_u.data(), mem_u.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) return (int)e;
  win.n_chunks = (int32_t)chunk_g.size() - 1;
  win.max_union = max_union;
  win.max_groups = max_groups;
  return XSDBA_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int xsdba_version(void) { return 100; }

int64_t xsdba_launch_count(void) { return g_launches.load(); }

const char* xsdba_status_string(int status) {
  switch (status) {
    case XSDBA_OK: return "ok";
    case XSDBA_ERR_INVALID_ARGUMENT: return "invalid argument";
    case XSDBA_ERR_UNSUPPORTED: return "valid in xsdba but not supported by xsdba_b200 yet";
    case XSDBA_ERR_SEGMENT_TOO_LONG: return "a (point, group) segment is longer than the in-SM sorter accepts";
    case XSDBA_ERR_NO_DEVICE: return "no CUDA device";
    case XSDBA_ERR_OUT_OF_MEMORY: return "out of memory";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown error";
  }
}

int xsdba_grouping_create(xsdba_grouping_t** out, const int32_t* grp_idx_host, int64_t n_time, int32_t n_groups,
                          int32_t window) {
  if (!out || !grp_idx_host || n_time <= 0 || n_groups <= 0 || window < 1 || n_time > 0x7fffffff)
    return XSDBA_ERR_INVALID_ARGUMENT;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return XSDBA_ERR_NO_DEVICE;
  for (int64_t t = 0; t < n_time; ++t)
    if (grp_idx_host[t] < -1 || grp_idx_host[t] >= n_groups) return XSDBA_ERR_INVALID_ARGUMENT;
  xsdba_grouping* g = new (std::nothrow) xsdba_grouping();
  if (!g) return XSDBA_ERR_OUT_OF_MEMORY;
  g->n_time = n_time; g->n_groups = n_groups; g->window = window;
  cudaGetDevice(&g->device);
  std::vector<int32_t> off(n_groups + 1, 0), rows;
  for (int64_t t = 0; t < n_time; ++t) if (grp_idx_host[t] >= 0) ++off[grp_idx_host[t] + 1];
  for (int i = 0; i < n_groups; ++i) off[i + 1] += off[i];
  rows.resize(off[n_groups]);
  {
    std::vector<int32_t> cur(off.begin(), off.end() - 1);
    for (int64_t t = 0; t < n_time; ++t) if (grp_idx_host[t] >= 0) rows[cur[grp_idx_host[t]]++] = (int32_t)t;
  }
  int rc = upload_table(off, rows, g->members);
  if (rc == XSDBA_OK) {
    cudaError_t e = cudaMalloc(&g->gidx, n_time * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(g->gidx, grp_idx_host, n_time * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = (int)e;
  }
  if (rc == XSDBA_OK) {
    if (window == 1) {
      g->segments = g->members;
    } else {
      // rolling(center=True).construct: slot j of the row centred on t is x[t - window/2 + j]  (base.py:261-265)
      const int half = window / 2;
      std::vector<int32_t> soff(n_groups + 1), srows;
      if ((int64_t)rows.size() * window > 0x7fffffff) { rc = XSDBA_ERR_SEGMENT_TOO_LONG; }
      else {
        srows.reserve(rows.size() * (size_t)window);
        for (int i = 0; i < n_groups; ++i) {
          soff[i] = (int32_t)srows.size();
          for (int m = off[i]; m < off[i + 1]; ++m)
            for (int j = 0; j < window; ++j) {
              const int64_t tt = (int64_t)rows[m] - half + j;
              srows.push_back(tt >= 0 && tt < n_time ? (int32_t)tt : -1);
            }
        }
        soff[n_groups] = (int32_t)srows.size();
        rc = upload_table(soff, srows, g->segments);
        if (rc == XSDBA_OK) rc = build_window_tables(soff, srows, (int32_t)n_time, n_groups, off, rows, g->win);
      }
    }
  }
  if (rc != XSDBA_OK) { xsdba_grouping_destroy(g); return rc; }
  if (g->segments.max_len > XSDBA_MAX_SEGMENT) { xsdba_grouping_destroy(g); return XSDBA_ERR_SEGMENT_TOO_LONG; }
  *out = g;
  return XSDBA_OK;
}

int xsdba_grouping_destroy(xsdba_grouping_t* g) {
  if (!g) return XSDBA_OK;
  if (g->segments.off != g->members.off) { cudaFree(g->segments.off); cudaFree(g->segments.rows); }
  cudaFree(g->members.off);
  cudaFree(g->members.rows);
  cudaFree(g->gidx);
  cudaFree(g->win.chunk_g); cudaFree(g->win.urow_off); cudaFree(g->win.urows); cudaFree(g->win.gmask); cudaFree(g->win.mem_u);
  delete g;
  return XSDBA_OK;
}

int64_t xsdba_grouping_max_segment(const xsdba_grouping_t* g) { return g ? g->segments.max_len : -1; }
int32_t xsdba_grouping_n_groups(const xsdba_grouping_t* g) { return g ? g->n_groups : -1; }

int xsdba_qm_train_f32(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                       const xsdba_grouping_t* grp, const float* q, int32_t nq, int32_t kind, int32_t normalize,
                       float* af, float* hq, float* scaling, void* stream) {
  return launch_train<float>(ref, hist, n_pts, sp, st, grp, q, nq, kind, normalize, 0, af, hq, scaling, stream);
}
int xsdba_qm_train_f64(const double* ref, const double* hist, int64_t n_pts, int64_t sp, int64_t st,
                       const xsdba_grouping_t* grp, const double* q, int32_t nq, int32_t kind, int32_t normalize,
                       double* af, double* hq, double* scaling, void* stream) {
  return launch_train<double>(ref, hist, n_pts, sp, st, grp, q, nq, kind, normalize, 0, af, hq, scaling, stream);
}
int xsdba_group_quantile_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                             const float* q, int32_t nq, float* out, void* stream) {
  return launch_train<float>(x, nullptr, n_pts, sp, st, grp, q, nq, XSDBA_KIND_ADD, 0, 1, out, nullptr, nullptr, stream);
}
int xsdba_group_quantile_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                             const double* q, int32_t nq, double* out, void* stream) {
  return launch_train<double>(x, nullptr, n_pts, sp, st, grp, q, nq, XSDBA_KIND_ADD, 0, 1, out, nullptr, nullptr, stream);
}

int xsdba_qm_adjust_f32(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                        const float* af, const float* hq, int32_t nq, int32_t interp, int32_t extrap, int32_t kind,
                        float* scen, void* stream) {
  return launch_adjust<float>(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, stream);
}
int xsdba_qm_adjust_f64(const double* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                        const double* af, const double* hq, int32_t nq, int32_t interp, int32_t extrap, int32_t kind,
                        double* scen, void* stream) {
  return launch_adjust<double>(sim, n_pts, sp, st, grp, af, hq, nq, interp, extrap, kind, scen, stream);
}

int xsdba_qdm_adjust_f32(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const float* af, const float* q, int32_t nq, int32_t interp, int32_t extrap, int32_t kind,
                         int32_t rank_window, float* scen, double* sim_q, void* stream) {
  return launch_rank<float>(sim, n_pts, sp, st, grp, af, q, nq, interp, extrap, kind, rank_window, 1, scen, sim_q, stream);
}
int xsdba_qdm_adjust_f64(const double* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const double* af, const double* q, int32_t nq, int32_t interp, int32_t extrap, int32_t kind,
                         int32_t rank_window, double* scen, double* sim_q, void* stream) {
  return launch_rank<double>(sim, n_pts, sp, st, grp, af, q, nq, interp, extrap, kind, rank_window, 1, scen, sim_q, stream);
}
int xsdba_group_rank_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         int32_t rank_window, double* rank, void* stream) {
  return launch_rank<float>(x, n_pts, sp, st, grp, nullptr, nullptr, 0, 0, 0, XSDBA_KIND_ADD, rank_window, 0, nullptr, rank, stream);
}
int xsdba_group_rank_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         int32_t rank_window, double* rank, void* stream) {
  return launch_rank<double>(x, n_pts, sp, st, grp, nullptr, nullptr, 0, 0, 0, XSDBA_KIND_ADD, rank_window, 0, nullptr, rank, stream);
}

int xsdba_qm_train_jitter_f32(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                              const xsdba_grouping_t* grp, const float* q, int32_t nq, int32_t kind, int32_t normalize,
                              const double* jitter4_host, uint64_t seed, float* af, float* hq, float* scaling,
                              void* stream) {
  return launch_train<float>(ref, hist, n_pts, sp, st, grp, q, nq, kind, normalize, 0, af, hq, scaling, stream,
                             jitter4_host, seed);
}
int xsdba_qm_train_jitter_f64(const double* ref, const double* hist, int64_t n_pts, int64_t sp, int64_t st,
                              const xsdba_grouping_t* grp, const double* q, int32_t nq, int32_t kind, int32_t normalize,
                              const double* jitter4_host, uint64_t seed, double* af, double* hq, double* scaling,
                              void* stream) {
  return launch_train<double>(ref, hist, n_pts, sp, st, grp, q, nq, kind, normalize, 0, af, hq, scaling, stream,
                              jitter4_host, seed);
}
int xsdba_jitter_f32(const float* x, int64_t n, const double* j4, uint64_t seed, float* out, void* stream) {
  return launch_jitter<float>(x, n, j4, seed, out, stream);
}
int xsdba_jitter_f64(const double* x, int64_t n, const double* j4, uint64_t seed, double* out, void* stream) {
  return launch_jitter<double>(x, n, j4, seed, out, stream);
}

int xsdba_qm_train_q64_f32(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                           const xsdba_grouping_t* grp, const double* q64, int32_t nq, int32_t kind, float* af,
                           float* hq, void* stream) {
  if (!q64) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_train<float>(ref, hist, n_pts, sp, st, grp, reinterpret_cast<const float*>(q64), nq, kind, 0, 0, af, hq,
                             nullptr, stream, nullptr, 0, q64);
}
int xsdba_npdft_step_f32(const float* ref, float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const double* q64, int32_t nq, int32_t interp, int32_t extrap, float* af_io, void* stream) {
  return launch_npdft_step(ref, x, n_pts, sp, st, grp, q64, nq, interp, extrap, af_io, stream);
}
int xsdba_rank_lookup_f32(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                          const float* af, const float* q, int32_t nq, int32_t interp, int32_t extrap, int32_t kind,
                          int32_t rank_window, int32_t rank_mode, float* scen, double* sim_q, void* stream) {
  return launch_rank<float>(sim, n_pts, sp, st, grp, af, q, nq, interp, extrap, kind, rank_window, 1, scen, sim_q, stream,
                            rank_mode);
}
int xsdba_rank_lookup_f64(const double* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                          const double* af, const double* q, int32_t nq, int32_t interp, int32_t extrap, int32_t kind,
                          int32_t rank_window, int32_t rank_mode, double* scen, double* sim_q, void* stream) {
  return launch_rank<double>(sim, n_pts, sp, st, grp, af, q, nq, interp, extrap, kind, rank_window, 1, scen, sim_q, stream,
                             rank_mode);
}
int xsdba_qdm_adjust_linear_f32(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                                const float* af, const float* q, int32_t nq, int32_t extrap, int32_t kind,
                                int32_t rank_window, const double* gcoord, const unsigned char* diag, float* scen,
                                double* sim_q, void* stream) {
  if (!gcoord || !diag) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_rank<float>(sim, n_pts, sp, st, grp, af, q, nq, XSDBA_INTERP_LINEAR, extrap, kind, rank_window, 1, scen,
                            sim_q, stream, 0, gcoord, diag);
}
int xsdba_qdm_adjust_linear_f64(const double* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                                const double* af, const double* q, int32_t nq, int32_t extrap, int32_t kind,
                                int32_t rank_window, const double* gcoord, const unsigned char* diag, double* scen,
                                double* sim_q, void* stream) {
  if (!gcoord || !diag) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_rank<double>(sim, n_pts, sp, st, grp, af, q, nq, XSDBA_INTERP_LINEAR, extrap, kind, rank_window, 1, scen,
                             sim_q, stream, 0, gcoord, diag);
}
int xsdba_escore_f32(const float* tgt, const float* sim, int64_t n_pts, int64_t sp, int64_t st, int64_t n_time_tgt,
                     int64_t n_time_sim, int32_t n_var, int64_t var_stride_tgt, int64_t var_stride_sim, int32_t n_sub,
                     float* out, void* stream) {
  return launch_escore<float>(tgt, sim, n_pts, sp, st, n_time_tgt, n_time_sim, n_var, var_stride_tgt, var_stride_sim, n_sub,
                              out, stream);
}
int xsdba_escore_f64(const double* tgt, const double* sim, int64_t n_pts, int64_t sp, int64_t st, int64_t n_time_tgt,
                     int64_t n_time_sim, int32_t n_var, int64_t var_stride_tgt, int64_t var_stride_sim, int32_t n_sub,
                     double* out, void* stream) {
  return launch_escore<double>(tgt, sim, n_pts, sp, st, n_time_tgt, n_time_sim, n_var, var_stride_tgt, var_stride_sim, n_sub,
                               out, stream);
}
int xsdba_rotate_f32(const float* x, int64_t n_elem, int32_t n_var, const float* rot_host, float* y, void* stream) {
  return launch_rotate<float>(x, n_elem, n_var, rot_host, y, stream);
}
int xsdba_rotate_f64(const double* x, int64_t n_elem, int32_t n_var, const float* rot_host, double* y, void* stream) {
  return launch_rotate<double>(x, n_elem, n_var, rot_host, y, stream);
}
int xsdba_rotate_unfused_f32(const float* x, int64_t n_elem, int32_t n_var, const float* rot_host, float* y, void* stream) {
  return launch_rotate<float>(x, n_elem, n_var, rot_host, y, stream, false);
}
int xsdba_rotate_unfused_f64(const double* x, int64_t n_elem, int32_t n_var, const float* rot_host, double* y, void* stream) {
  return launch_rotate<double>(x, n_elem, n_var, rot_host, y, stream, false);
}
int xsdba_standardize_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, int64_t n_time, int32_t n_var,
                          int64_t var_stride, float* y, void* stream) {
  return launch_standardize<float>(x, n_pts, sp, st, n_time, n_var, var_stride, y, stream);
}
int xsdba_standardize_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, int64_t n_time, int32_t n_var,
                          int64_t var_stride, double* y, void* stream) {
  return launch_standardize<double>(x, n_pts, sp, st, n_time, n_var, var_stride, y, stream);
}
int xsdba_reorder_f32(const float* sim, const float* ref, int64_t n_pts, int64_t sp, int64_t st,
                      const xsdba_grouping_t* grp, float* out, void* stream) {
  return launch_reorder<float>(sim, ref, n_pts, sp, st, grp, out, stream);
}
int xsdba_reorder_f64(const double* sim, const double* ref, int64_t n_pts, int64_t sp, int64_t st,
                      const xsdba_grouping_t* grp, double* out, void* stream) {
  return launch_reorder<double>(sim, ref, n_pts, sp, st, grp, out, stream);
}

int xsdba_group_vecquantile_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                                const float* rnk, float* out, void* stream) {
  return launch_select<float>(x, nullptr, n_pts, sp, st, grp, 0, rnk, nullptr, 0, out, stream);
}
int xsdba_group_vecquantile_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                                const double* rnk, double* out, void* stream) {
  return launch_select<double>(x, nullptr, n_pts, sp, st, grp, 0, rnk, nullptr, 0, out, stream);
}
int xsdba_map_cdf_f32(const float* x, const float* y, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                      const double* yvals, int32_t nv, float* out, void* stream) {
  return launch_select<float>(x, y, n_pts, sp, st, grp, 1, nullptr, yvals, nv, out, stream);
}
int xsdba_map_cdf_f64(const double* x, const double* y, int64_t n_pts, int64_t sp, int64_t st,
                      const xsdba_grouping_t* grp, const double* yvals, int32_t nv, double* out, void* stream) {
  return launch_select<double>(x, y, n_pts, sp, st, grp, 1, nullptr, yvals, nv, out, stream);
}

int xsdba_dqm_train_adapt_f32(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                              const xsdba_grouping_t* grp, const float* q, int32_t nq, int32_t kind,
                              const double* jitter4_host, double adapt_thresh, uint64_t seed, float* af, float* hq,
                              float* scaling, double* P0_ref, double* P0_hist, float* pth, void* stream) {
  if (!P0_ref || !P0_hist || !pth || !(adapt_thresh == adapt_thresh)) return XSDBA_ERR_INVALID_ARGUMENT;
  AdaptParams ap{1, adapt_thresh, seed, P0_ref, P0_hist, pth};
  return launch_train<float>(ref, hist, n_pts, sp, st, grp, q, nq, kind, 1, 0, af, hq, scaling, stream, jitter4_host, seed,
                             nullptr, &ap);
}
int xsdba_dqm_train_adapt_f64(const double* ref, const double* hist, int64_t n_pts, int64_t sp, int64_t st,
                              const xsdba_grouping_t* grp, const double* q, int32_t nq, int32_t kind,
                              const double* jitter4_host, double adapt_thresh, uint64_t seed, double* af, double* hq,
                              double* scaling, double* P0_ref, double* P0_hist, double* pth, void* stream) {
  if (!P0_ref || !P0_hist || !pth || !(adapt_thresh == adapt_thresh)) return XSDBA_ERR_INVALID_ARGUMENT;
  AdaptParams ap{1, adapt_thresh, seed, P0_ref, P0_hist, pth};
  return launch_train<double>(ref, hist, n_pts, sp, st, grp, q, nq, kind, 1, 0, af, hq, scaling, stream, jitter4_host,
                              seed, nullptr, &ap);
}
int xsdba_qm_train_adapt_f32(const float* ref, const float* hist, int64_t n_pts, int64_t sp, int64_t st,
                             const xsdba_grouping_t* grp, const float* q, int32_t nq, int32_t kind,
                             const double* jitter4_host, double adapt_thresh, uint64_t seed, float* af, float* hq,
                             double* P0_ref, double* P0_hist, float* pth, void* stream) {
  if (!P0_ref || !P0_hist || !pth || !(adapt_thresh == adapt_thresh)) return XSDBA_ERR_INVALID_ARGUMENT;
  AdaptParams ap{1, adapt_thresh, seed, P0_ref, P0_hist, pth};
  return launch_train<float>(ref, hist, n_pts, sp, st, grp, q, nq, kind, 0, 0, af, hq, nullptr, stream, jitter4_host, seed,
                             nullptr, &ap);
}
int xsdba_qm_train_adapt_f64(const double* ref, const double* hist, int64_t n_pts, int64_t sp, int64_t st,
                             const xsdba_grouping_t* grp, const double* q, int32_t nq, int32_t kind,
                             const double* jitter4_host, double adapt_thresh, uint64_t seed, double* af, double* hq,
                             double* P0_ref, double* P0_hist, double* pth, void* stream) {
  if (!P0_ref || !P0_hist || !pth || !(adapt_thresh == adapt_thresh)) return XSDBA_ERR_INVALID_ARGUMENT;
  AdaptParams ap{1, adapt_thresh, seed, P0_ref, P0_hist, pth};
  return launch_train<double>(ref, hist, n_pts, sp, st, grp, q, nq, kind, 0, 0, af, hq, nullptr, stream, jitter4_host, seed,
                              nullptr, &ap);
}
int xsdba_adapt_freq_apply_f32(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                               double thresh, const double* P0_ref, const double* P0_hist, const float* pth, uint64_t seed,
                               float* out, void* stream) {
  return launch_adapt_apply<float>(sim, n_pts, sp, st, grp, thresh, P0_ref, P0_hist, pth, seed, out, stream);
}
int xsdba_adapt_freq_apply_f64(const double* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                               double thresh, const double* P0_ref, const double* P0_hist, const double* pth,
                               uint64_t seed, double* out, void* stream) {
  return launch_adapt_apply<double>(sim, n_pts, sp, st, grp, thresh, P0_ref, P0_hist, pth, seed, out, stream);
}
int xsdba_tail_mask_f32(const float* adapted, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                        const float* hq_raw, int32_t nq, double factor, float* scen, void* stream) {
  return launch_tail_mask<float>(adapted, n_pts, sp, st, grp, hq_raw, nq, factor, scen, stream);
}
int xsdba_tail_mask_f64(const double* adapted, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                        const double* hq_raw, int32_t nq, double factor, double* scen, void* stream) {
  return launch_tail_mask<double>(adapted, n_pts, sp, st, grp, hq_raw, nq, factor, scen, stream);
}

int xsdba_poly_trend_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const float* scaling, int32_t kind, int32_t degree, const double* tcoord, double* trend,
                         void* stream) {
  return launch_poly_trend<float>(x, n_pts, sp, st, grp, scaling, kind, degree, tcoord, trend, stream);
}
int xsdba_poly_trend_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const double* scaling, int32_t kind, int32_t degree, const double* tcoord, double* trend,
                         void* stream) {
  return launch_poly_trend<double>(x, n_pts, sp, st, grp, scaling, kind, degree, tcoord, trend, stream);
}
int xsdba_dqm_adjust_f32(const float* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const float* af, const float* hq, const float* scaling, const double* trend, int32_t nq,
                         int32_t interp, int32_t extrap, int32_t kind, float* scen, void* stream) {
  return launch_dqm_adjust<float>(sim, n_pts, sp, st, grp, af, hq, scaling, trend, nq, interp, extrap, kind, scen, stream);
}
int xsdba_dqm_adjust_f64(const double* sim, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                         const double* af, const double* hq, const double* scaling, const double* trend, int32_t nq,
                         int32_t interp, int32_t extrap, int32_t kind, double* scen, void* stream) {
  return launch_dqm_adjust<double>(sim, n_pts, sp, st, grp, af, hq, scaling, trend, nq, interp, extrap, kind, scen, stream);
}

int xsdba_loess_trend_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                          const float* scaling, int32_t kind, double f, int32_t niter, int32_t degree, const double* xn,
                          double* trend, void* stream) {
  if (!(f > 0.0)) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_loess_trend<float>(x, n_pts, sp, st, grp, scaling, kind, f, niter, degree, xn, trend, stream);
}
int xsdba_loess_trend_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                          const double* scaling, int32_t kind, double f, int32_t niter, int32_t degree, const double* xn,
                          double* trend, void* stream) {
  if (!(f > 0.0)) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_loess_trend<double>(x, n_pts, sp, st, grp, scaling, kind, f, niter, degree, xn, trend, stream);
}

// weights: 0 = tricube, 1 = gaussian (loess.py:16-35, 247); equal_spacing: 1 = the dx > 0 branch (xn equally spaced),
// 0 = the dx == 0 branch (loess.py:251-260)
int xsdba_loess_trend_w_f32(const float* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                            const float* scaling, int32_t kind, double f, int32_t niter, int32_t degree, int32_t weights,
                            int32_t equal_spacing, const double* xn, double* trend, void* stream) {
  if (!(f > 0.0) || (weights != 0 && weights != 1)) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_loess_trend<float>(x, n_pts, sp, st, grp, scaling, kind, weights ? -f : f, niter, degree, xn, trend, stream,
                                   equal_spacing ? 0 : 1);
}
int xsdba_loess_trend_w_f64(const double* x, int64_t n_pts, int64_t sp, int64_t st, const xsdba_grouping_t* grp,
                            const double* scaling, int32_t kind, double f, int32_t niter, int32_t degree, int32_t weights,
                            int32_t equal_spacing, const double* xn, double* trend, void* stream) {
  if (!(f > 0.0) || (weights != 0 && weights != 1)) return XSDBA_ERR_INVALID_ARGUMENT;
  return launch_loess_trend<double>(x, n_pts, sp, st, grp, scaling, kind, weights ? -f : f, niter, degree, xn, trend, stream,
                                    equal_spacing ? 0 : 1);
}

// microbenchmark entry (see copy_rows_kernel); time-major float32 only, n_pts % (32*v) == 0 expected
int xsdba_debug_copy_rows_f32(const float* src, int64_t n_pts, int64_t st, const xsdba_grouping_t* grp, float* dst,
                              int32_t v, void* stream) {
  if (!src || !dst || !grp) return XSDBA_ERR_INVALID_ARGUMENT;
  dim3 grid((unsigned)((n_pts + 32 * v - 1) / (32 * v)), (unsigned)grp->n_groups);
  cudaStream_t s = (cudaStream_t)stream;
  if (v == 1) copy_rows_kernel<1><<<grid, kThreads, 0, s>>>(src, n_pts, st, grp->members.off, grp->members.rows, dst);
  else if (v == 2) copy_rows_kernel<2><<<grid, kThreads, 0, s>>>(src, n_pts, st, grp->members.off, grp->members.rows, dst);
  else if (v == 4) copy_rows_kernel<4><<<grid, kThreads, 0, s>>>(src, n_pts, st, grp->members.off, grp->members.rows, dst);
  else return XSDBA_ERR_INVALID_ARGUMENT;
  return cuda_status(cudaGetLastError());
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// host end-to-end entry point (defined in host_pipeline.inc)
// ---------------------------------------------------------------------------------------------
#include "host_pipeline.inc"
