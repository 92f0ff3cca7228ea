// Shared device helpers for the xsdba_b200 kernels (sm_100a).
//
// Arithmetic notes (these pin bit-parity with the reference; measured facts are listed in DESIGN.md):
//  * quantile virtual index  vi = (n-1)*q  rounded once in float64   (nbutils.py:131, LLVM-folded)
//  * gamma cast to the data type; lerp branches are FMAs             (nbutils.py:101-104, contract)
//  * SciPy interp1d / numpy.interp / cKDTree arithmetic is NOT contracted: every product / sum is
//    rounded separately  -> __fmul_rn / __fadd_rn / __dmul_rn / __dadd_rn below.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace xsdba {

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
  static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
  // sort keys are NaN-free (NaN -> +inf before sorting), so one compare + selects will do: fmin / fmax on doubles
  // expand to ~13 instructions each (IEEE NaN handling in integer arithmetic) and were 61 % of the float64 sorter
  static __device__ __forceinline__ double mn(double a, double b) { return b < a ? b : a; }
  static __device__ __forceinline__ double mx(double a, double b) { return b < a ? a : b; }
  static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};

// monotone uint32 sort keys of the window trainer (K1w): only the column sorter's min / max
template <> struct Num<unsigned> {
  static __device__ __forceinline__ unsigned mn(unsigned a, unsigned b) { return a < b ? a : b; }
  static __device__ __forceinline__ unsigned mx(unsigned a, unsigned b) { return a < b ? b : a; }
};
template <typename T> __device__ __forceinline__ bool is_nan(T v) { return v != v; }

// ---------------------------------------------------------------------------------------------
// Jitter (processing.jitter, processing.py:180-257; used on `hist` inside the per-group train
// function, _adjustment.py:58-67): non-NaN values < lower are REPLACED by U(minimum, lower), values
// >= upper by U(upper, maximum).  The reference draws from numpy's global RNG (not reproducible by
// design); here the draw is a counter-based hash of (seed, element id), so a call is
// reproducible and every window slot gets its own draw like in the reference.
// ---------------------------------------------------------------------------------------------
struct JitterParams {
  double lower, minimum, upper, maximum;  // lower / upper = NaN disables that side
  unsigned long long seed;
};

__device__ __forceinline__ double hash_uniform(unsigned long long seed, unsigned long long id) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (id + 1);  // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);  // [0, 1)
}

template <typename T>
__device__ __forceinline__ T jitter_value(T v, const JitterParams& jp, unsigned long long id) {
  if (v != v) return v;
  if ((double)v < jp.lower) return (T)(jp.minimum + (jp.lower - jp.minimum) * hash_uniform(jp.seed, 2 * id));
  if ((double)v >= jp.upper) return (T)(jp.upper + (jp.maximum - jp.upper) * hash_uniform(jp.seed, 2 * id + 1));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Column sort in shared memory.  sm is [n_pad][C] (column c of row r at sm[r*C + c]); n_pad is a
// power of two; NaNs have been replaced by +inf by the caller.  Every column is sorted ascending.
//
// sort_columns_v0: one compare-exchange per thread per step (bitonic network), a __syncthreads and a
// shared-memory round trip per stage -- log2(n)(log2(n)+1)/2 of them.
// sort_columns_rb: the same network, register blocked.  A thread takes the 32 rows of one column that
// differ in 5 consecutive index bits, runs up to 5 stages on them in registers and writes them back, so
// a phase of p stages costs ceil(p/5) round trips and the first five phases (a full sort of every
// 32-row block) cost one: 11 round trips instead of 55 for 1024 rows.  (K1f's sorter in sort.cuh is the
// hand-scheduled float32 / 1024-thread special case of this; this one serves every other kernel.)
// With C < 32 columns several lanes of a warp work on different row sets of the same column; when the
// exchange bits include bit 0 those sets are 32+ rows apart and hit the same banks (32/C-way conflict),
// which is why the narrow tiles of very long segments (C < 8) stay on v0.
// ---------------------------------------------------------------------------------------------
template <typename T, int C>
__device__ void sort_columns_v0(T* sm, int n_pad) {
  const int half_items = (n_pad >> 1) * C;
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int id = threadIdx.x; id < half_items; id += blockDim.x) {
        const int c = id % C;
        const int p = id / C;
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int l = i | j;
        const bool asc = (i & k) == 0;
        const T a = sm[i * C + c];
        const T b = sm[l * C + c];
        const T lo = Num<T>::mn(a, b);
        const T hi = Num<T>::mx(a, b);
        sm[i * C + c] = asc ? lo : hi;
        sm[l * C + c] = asc ? hi : lo;
      }
      __syncthreads();
    }
  }
}

// stages of exchange bits R-1 .. 0 (local) on the 2^R registers, all in direction DESC
template <typename T, int R, bool DESC>
__device__ __forceinline__ void rb_stages(T (&r)[1 << R]) {
#pragma unroll
  for (int lb = R - 1; lb >= 0; --lb) {
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) {
      if (j & (1 << lb)) continue;
      const T lo = Num<T>::mn(r[j], r[j | (1 << lb)]);
      const T hi = Num<T>::mx(r[j], r[j | (1 << lb)]);
      r[j] = DESC ? hi : lo;
      r[j | (1 << lb)] = DESC ? lo : hi;
    }
  }
}

// full bitonic sort of the 2^R registers; the last phase runs in direction `desc`
template <typename T, int R>
__device__ __forceinline__ void rb_sort_block(T (&r)[1 << R], bool desc) {
#pragma unroll
  for (int ph = 1; ph < R; ++ph) {
#pragma unroll
    for (int lb = ph - 1; lb >= 0; --lb) {
#pragma unroll
      for (int j = 0; j < (1 << R); ++j) {
        if (j & (1 << lb)) continue;
        const bool d = ((j >> ph) & 1) != 0;  // compile-time after unrolling
        const T lo = Num<T>::mn(r[j], r[j | (1 << lb)]);
        const T hi = Num<T>::mx(r[j], r[j | (1 << lb)]);
        r[j] = d ? hi : lo;
        r[j | (1 << lb)] = d ? lo : hi;
      }
    }
  }
  if (desc) rb_stages<T, R, true>(r);
  else rb_stages<T, R, false>(r);
}

// One pass over the whole tile: exchange bits [b_lo, b_lo + R) of phase p (FIRST: phases 1..R at once).
// L = log2(n_pad).  Ends with __syncthreads.
template <typename T, int C, int R, bool FIRST>
__device__ __forceinline__ void rb_pass(T* sm, int n_pad, int L, int b_lo, int p) {
  const int n_items = (n_pad >> R) * C;
  for (int item = threadIdx.x; item < n_items; item += blockDim.x) {
    const int c = item % C, s = item / C;
    const int base = ((s >> b_lo) << (b_lo + R)) | (s & ((1 << b_lo) - 1));
    T* col = sm + (size_t)base * C + c;
    const size_t step = (size_t)C << b_lo;
    T r[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) r[j] = col[j * step];
    const bool desc = p < L && ((base >> p) & 1);
    if (FIRST) {
      rb_sort_block<T, R>(r, desc);
    } else if (desc) {
      rb_stages<T, R, true>(r);
    } else {
      rb_stages<T, R, false>(r);
    }
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) col[j * step] = r[j];
  }
  __syncthreads();
}

template <typename T, int C, bool FIRST>
__device__ __forceinline__ void rb_pass_r(T* sm, int n_pad, int L, int b_lo, int p, int R) {
  switch (R) {
    case 1: rb_pass<T, C, 1, FIRST>(sm, n_pad, L, b_lo, p); break;
    case 2: rb_pass<T, C, 2, FIRST>(sm, n_pad, L, b_lo, p); break;
    case 3: rb_pass<T, C, 3, FIRST>(sm, n_pad, L, b_lo, p); break;
    case 4: rb_pass<T, C, 4, FIRST>(sm, n_pad, L, b_lo, p); break;
    default: rb_pass<T, C, 5, FIRST>(sm, n_pad, L, b_lo, p); break;
  }
}

template <typename T, int C>
__device__ void sort_columns_rb(T* sm, int n_pad) {
  int L = 0;
  while ((1 << L) < n_pad) ++L;
  const int R0 = L < 5 ? L : 5;
  rb_pass_r<T, C, true>(sm, n_pad, L, 0, R0, R0);  // phases 1..R0: every 2^R0-row block sorted
  for (int p = R0 + 1; p <= L; ++p) {
    int top = p;  // stages of bits p-1 .. 0, in chunks of up to 5 from the top
    while (top > 0) {
      const int R = top < 5 ? top : 5;
      rb_pass_r<T, C, false>(sm, n_pad, L, top - R, p, R);
      top -= R;
    }
  }
}

// Narrow tiles (C < 8 columns: segments of 4k..32k rows).  The exchange bits 5 and above are register blocked
// as above -- the lanes of a warp then take consecutive rows, which is conflict free for any C -- and the
// stages of bits 4..0 run inside a warp on one 32-row block of all C columns: the block is 32*C contiguous
// floats, lane l holds elements l, l+32, ... (row = element / C), so the low 5 - log2(C) row bits live in
// the lane index (compare-exchange by __shfl_xor) and the remaining ones in the register index.
template <typename T, int C, bool FIRST>
__device__ __forceinline__ void warp_block_pass(T* sm, int n_pad, int L, int p) {
  constexpr int LC = C == 1 ? 0 : (C == 2 ? 1 : (C == 4 ? 2 : (C == 8 ? 3 : (C == 16 ? 4 : 5))));
  constexpr int LANE_BITS = 5 - LC;  // row bits held in the lane index (lane bit = row bit + LC)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int blk = warp; blk < (n_pad >> 5); blk += n_warps) {
    T* base = sm + (size_t)blk * 32 * C;
    T v[C];
#pragma unroll
    for (int k = 0; k < C; ++k) v[k] = base[lane + 32 * k];
    const bool blk_desc = p < L && (((blk << 5) >> p) & 1);
#pragma unroll
    for (int ph = FIRST ? 1 : 5; ph <= 5; ++ph) {
#pragma unroll
      for (int b = (FIRST ? ph : 5) - 1; b >= 0; --b) {
        if (b < LANE_BITS) {
          const bool upper = (lane >> (b + LC)) & 1;
#pragma unroll
          for (int k = 0; k < C; ++k) {
            // direction of this element's sub-sequence: row bit ph inside the block, the block's own bit for ph == 5
            const int row = (lane >> LC) + k * (32 >> LC);
            const bool desc = (FIRST && ph < 5) ? ((row >> ph) & 1) : blk_desc;
            const T o = __shfl_xor_sync(0xffffffffu, v[k], 1 << (b + LC));
            // both lanes evaluate the same ordered pair (lower lane's value, upper lane's value), so that values
            // that compare equal but differ in bits (+-0) are neither duplicated nor lost
            const T p = upper ? o : v[k], q = upper ? v[k] : o;
            const T lo = Num<T>::mn(p, q), hi = Num<T>::mx(p, q);
            v[k] = (upper == desc) ? lo : hi;
          }
        } else {
          const int kb = 1 << (b - LANE_BITS);
#pragma unroll
          for (int k = 0; k < C; ++k) {
            if (k & kb) continue;
            const int row = (lane >> LC) + k * (32 >> LC);
            const bool desc = (FIRST && ph < 5) ? ((row >> ph) & 1) : blk_desc;
            const T lo = Num<T>::mn(v[k], v[k | kb]), hi = Num<T>::mx(v[k], v[k | kb]);
            v[k] = desc ? hi : lo;
            v[k | kb] = desc ? lo : hi;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < C; ++k) base[lane + 32 * k] = v[k];
  }
  __syncthreads();
}

template <typename T, int C>
__device__ void sort_columns_narrow(T* sm, int n_pad) {
  int L = 0;
  while ((1 << L) < n_pad) ++L;
  warp_block_pass<T, C, true>(sm, n_pad, L, 5);  // phases 1..5: every 32-row block sorted
  for (int p = 6; p <= L; ++p) {
    int top = p;  // bits p-1 .. 5 register blocked in chunks of up to 5, then bits 4 .. 0 inside a warp
    while (top > 5) {
      const int R = top - 5 < 5 ? top - 5 : 5;
      rb_pass_r<T, C, false>(sm, n_pad, L, top - R, p, R);
      top -= R;
    }
    warp_block_pass<T, C, false>(sm, n_pad, L, p);
  }
}

template <typename T, int C>
__device__ void sort_columns(T* sm, int n_pad) {
  if (n_pad < 32) sort_columns_v0<T, C>(sm, n_pad);
  else if (C >= 8) sort_columns_rb<T, C>(sm, n_pad);
  else sort_columns_narrow<T, C>(sm, n_pad);
}

// Value at position i (may be negative, python-style) of the reference's sorted full-length row:
// positions [0, n) are the sorted valid values, positions [n, S) are NaN (numba sorts NaNs last).
template <typename T, int C>
__device__ __forceinline__ T sorted_at(const T* col, long long i, int n, int S) {
  if (i < 0) i += S;
  if (i < 0 || i >= n) return Num<T>::nan();
  return col[(size_t)i * C];
}

// Type-7 quantile of one sorted column: _nan_quantile_1d + _get_indexes + _linear_interpolation
// (nbutils.py:24-148).  col points at sm[0*C + c]; n = number of valid values; S = segment length.
template <typename T, int C>
__device__ __forceinline__ T quantile_sorted(const T* col, int n, int S, double qk) {
  const double vi = (double)(n - 1) * qk;   // nbutils.py:131
  long long prev = (long long)floor(vi);
  long long next = prev + 1;
  if (vi >= (double)(n - 1)) { prev = -1; next = -1; }  // nbutils.py:47-51
  if (vi < 0.0) { prev = 0; next = 0; }                 // nbutils.py:53-56
  if (vi != vi) { prev = -1; next = -1; }               // nbutils.py:57-62
  const T left = sorted_at<T, C>(col, prev, n, S);
  const T right = sorted_at<T, C>(col, next, n, S);
  const T gamma = (T)(vi - (double)prev);               // nbutils.py:142
  const T diff = right - left;
  T res = (gamma >= (T)0.5) ? Num<T>::fma(-diff, (T)1 - gamma, right)   // nbutils.py:103-104
                            : Num<T>::fma(diff, gamma, left);           // nbutils.py:101-102
  if (is_nan(res)) res = sorted_at<T, C>(col, (long long)n - 1, n, S);  // nbutils.py:146
  return res;
}

// ---------------------------------------------------------------------------------------------
// Factor lookup tables (interp_on_quantiles).  For a tile of C points the kernel stages up to three
// rows of the cyclically padded tables (group row r-1, r, r+1) in shared memory, NaN nodes dropped
// (utils.py:351-352, 381-382), as xs/ys[slot][k][C]; nv[slot][C] is the number of kept nodes.
// blo/bhi (first/last non-NaN hist_q of the centre row) and clo/chi (first/last non-NaN af) feed
// the extrapolation rule (nbutils.py:375-416; utils.py:362-368).
// ---------------------------------------------------------------------------------------------
template <typename T, int C>
struct Tables {
  T* xsl[3];   // per slot (rows g-1, g, g+1): [ld][C] compacted nodes, +inf padded
  T* ysl[3];   // per slot: [ld][C] factors of the kept nodes
  int* nvl[3]; // per slot: [C] number of kept nodes
  T* blo;      // [C] centre row: first / last non-NaN hist_q, first / last non-NaN af
  T* bhi;
  T* clo;
  T* chi;
  int nq;
  int top;   // largest power of two <= nq (first step of the branch-free searches)
  int ld;    // rows per slot = 2*top >= nq+1; rows [nv, ld) of xs hold +inf so searches need no bound check
  // global fallbacks for rows further than +-1 (rare): raw tables of this tile's points
  const T* gx;      // hist_q (per point) or q (shared, x_shared = true)
  const T* gy;      // af
  bool x_shared;
  bool centre_only; // only slot 1 is staged: rows g-1 / g+1 are read from global memory when (rarely) needed
  int G;            // number of groups (rows are cyclic)
  long long pt_stride;  // G*nq
};

// number of nodes strictly less than x (searchsorted side='left') in compacted column
template <typename TX, typename T, int C>
__device__ __forceinline__ int lower_bound_col(const T* xs, int n, TX x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((TX)xs[mid * C] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// 1-D rule = scipy.interpolate.interp1d as called by utils._interp_on_quantiles_1D (utils.py:350-377)
template <typename TX, typename T, int C>
__device__ T lookup_1d(const Tables<T, C>& tb, int c, TX x, int interp, int extrap) {
  if (is_nan(x)) return Num<T>::nan();
  const int n = tb.nvl[1][c];
  if (n == 0) return Num<T>::nan();
  const T* xs = tb.xsl[1] + c;
  const T* ys = tb.ysl[1] + c;
  // _check_bounds / fill_value (scipy _interpolate.py: interp1d._evaluate)
  if (x < (TX)xs[0]) return extrap == 0 ? tb.clo[c] : Num<T>::nan();
  if (x > (TX)xs[(size_t)(n - 1) * C]) return extrap == 0 ? tb.chi[c] : Num<T>::nan();
  if (interp == 0) {
    // nearest: x_bds = x/2.0 ; x_bds[1:] + x_bds[:-1] in the node dtype, searchsorted side='left'
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const T bd = Num<T>::add(Num<T>::mul(xs[(size_t)(mid + 1) * C], (T)0.5), Num<T>::mul(xs[(size_t)mid * C], (T)0.5));
      if ((TX)bd < x) lo = mid + 1; else hi = mid;
    }
    return ys[(size_t)lo * C];
  }
  if (interp == 2) {
    // cubic: scipy interp1d(kind="cubic") = make_interp_spline(k=3), the not-a-knot cubic spline through the kept nodes
    // (needs four of them; SciPy raises below that).  Second derivatives M from stage_cubic (tb.ysl[0]); on
    // [x_i, x_{i+1}]:  S = A y_i + B y_{i+1} + ((A^3 - A) M_i + (B^3 - B) M_{i+1}) h^2 / 6, float64.
    if (n < 4) return Num<T>::nan();
    const T* ms = tb.ysl[0] + c;
    int idx = lower_bound_col<TX, T, C>(xs, n, x);
    idx = idx < 1 ? 1 : (idx > n - 1 ? n - 1 : idx);
    const double x0 = (double)xs[(size_t)(idx - 1) * C], x1 = (double)xs[(size_t)idx * C];
    const double y0 = (double)ys[(size_t)(idx - 1) * C], y1 = (double)ys[(size_t)idx * C];
    const double m0 = (double)ms[(size_t)(idx - 1) * C], m1 = (double)ms[(size_t)idx * C];
    const double h = x1 - x0, A = (x1 - (double)x) / h, B = ((double)x - x0) / h;
    return (T)(A * y0 + B * y1 + ((A * A * A - A) * m0 + (B * B * B - B) * m1) * (h * h) / 6.0);
  }
  if (n < 2) return Num<T>::nan();
  if (sizeof(T) == 8 && sizeof(TX) == 8) {
    // float64 nodes and values: interp1d delegates to numpy.interp (compiled_base.c arr_interp)
    int j = lower_bound_col<TX, T, C>(xs, n, x);           // #nodes < x
    if (j < n && (TX)xs[(size_t)j * C] == x) {
      // largest index with xs[j] <= x
      while (j + 1 < n && (TX)xs[(size_t)(j + 1) * C] == x) ++j;
      return ys[(size_t)j * C];
    }
    j -= 1;  // xs[j] < x < xs[j+1]
    const double x0 = xs[(size_t)j * C], x1 = xs[(size_t)(j + 1) * C];
    const double y0 = ys[(size_t)j * C], y1 = ys[(size_t)(j + 1) * C];
    const double slope = __ddiv_rn(__dsub_rn(y1, y0), __dsub_rn(x1, x0));
    double r = __dadd_rn(__dmul_rn(slope, __dsub_rn((double)x, x0)), y0);
    if (r != r) {
      r = __dadd_rn(__dmul_rn(slope, __dsub_rn((double)x, x1)), y1);
      if (r != r && y0 == y1) r = y0;
    }
    return (T)r;
  }
  // interp1d._call_linear: searchsorted(x, x_new) clipped to [1, n-1]; two-term form.
  int idx = lower_bound_col<TX, T, C>(xs, n, x);
  idx = idx < 1 ? 1 : (idx > n - 1 ? n - 1 : idx);
  const T x_lo = xs[(size_t)(idx - 1) * C], x_hi = xs[(size_t)idx * C];
  const T y_lo = ys[(size_t)(idx - 1) * C], y_hi = ys[(size_t)idx * C];
  const T den = Num<T>::sub(x_hi, x_lo);  // node dtype
  const TX w_hi = Num<TX>::div(Num<TX>::sub(x, (TX)x_lo), (TX)den);
  const TX w_lo = Num<TX>::div(Num<TX>::sub((TX)x_hi, x), (TX)den);
  return (T)Num<TX>::add(Num<TX>::mul(w_hi, (TX)y_hi), Num<TX>::mul(w_lo, (TX)y_lo));
}

// Second derivatives of the not-a-knot cubic spline through the compacted centre row of every column (1-D tables,
// group = "time"): tridiagonal system in M_1..M_{n-2} with M_0, M_{n-1} eliminated by the not-a-knot conditions
// (third derivative continuous at x_1 and x_{n-2}), Thomas algorithm with float64 arithmetic (recurrence values kept
// in the table dtype), one thread per column.  M goes to the factor array of slot 0, the recurrences to the node
// arrays of slots 0 and 2 -- the neighbour-row slots, which only grouped lookups use.  Needs the three-slot layout.
// Call after stage_tables (ends with a barrier itself).
template <typename T, int C>
__device__ void stage_cubic(const Tables<T, C>& tb) {
  if (threadIdx.x < C) {
    const int c = threadIdx.x, n = tb.nvl[1][c];
    const T* xs = tb.xsl[1] + c;
    const T* ys = tb.ysl[1] + c;
    T* ms = tb.ysl[0] + c;
    T* cp = tb.xsl[0] + c;   // modified super-diagonal
    T* dp = tb.xsl[2] + c;   // modified right-hand side, then the solution
    if (n >= 4) {
      auto X = [&](int k) { return (double)xs[(size_t)k * C]; };
      auto Y = [&](int k) { return (double)ys[(size_t)k * C]; };
      const int m = n - 2;  // unknowns M_1..M_{n-2}, row i <-> M_{i+1}
      for (int i = 0; i < m; ++i) {
        const int k = i + 1;
        const double h0 = X(k) - X(k - 1), h1 = X(k + 1) - X(k);
        double a = h0, b = 2.0 * (h0 + h1), cc = h1;
        const double r = 6.0 * ((Y(k + 1) - Y(k)) / h1 - (Y(k) - Y(k - 1)) / h0);
        if (k == 1) {          // M_0 = ((h0 + h1) M_1 - h0 M_2) / h1
          b += h0 * (h0 + h1) / h1; cc -= h0 * h0 / h1; a = 0.0;
        }
        if (k == n - 2) {      // M_{n-1} = ((h0 + h1) M_{n-2} - h1 M_{n-3}) / h0
          b += h1 * (h0 + h1) / h0; a -= h1 * h1 / h0; cc = 0.0;
        }
        if (i == 0) { cp[0] = (T)(cc / b); dp[0] = (T)(r / b); }
        else {
          const double den = b - a * (double)cp[(size_t)(i - 1) * C];
          cp[(size_t)i * C] = (T)(cc / den);
          dp[(size_t)i * C] = (T)((r - a * (double)dp[(size_t)(i - 1) * C]) / den);
        }
      }
      double next = 0.0;
      for (int i = m - 1; i >= 0; --i) {
        const double v = (double)dp[(size_t)i * C] - (double)cp[(size_t)i * C] * next;
        ms[(size_t)(i + 1) * C] = (T)v;
        dp[(size_t)i * C] = (T)v;
        next = v;
      }
      {
        const double h0 = X(1) - X(0), h1 = X(2) - X(1);
        ms[0] = (T)(((h0 + h1) * (double)dp[0] - h0 * (double)dp[(size_t)1 * C]) / h1);
        const double g0 = X(n - 2) - X(n - 3), g1 = X(n - 1) - X(n - 2);
        ms[(size_t)(n - 1) * C] = (T)(((g0 + g1) * (double)dp[(size_t)(m - 1) * C] - g1 * (double)dp[(size_t)(m - 2) * C]) / g0);
      }
    }
  }
  __syncthreads();
}

// best candidate of one compacted row for the 2-D Euclidean-nearest rule
template <typename TX, typename T, int C>
__device__ __forceinline__ void nearest_in_row(const T* xs, const T* ys, int n, TX x, double dg2, double& best_d2,
                                               T& best_y) {
  if (n == 0) return;
  const int i = lower_bound_col<TX, T, C>(xs, n, x);
  double dbest = 1e300;  // placeholder, replaced below
  int ibest = -1;
  if (i > 0) { dbest = fabs((double)x - (double)xs[(size_t)(i - 1) * C]); ibest = i - 1; }
  if (i < n) {
    const double d = fabs((double)xs[(size_t)i * C] - (double)x);
    if (ibest < 0 || d < dbest) { dbest = d; ibest = i; }
  }
  const double d2 = __dadd_rn(__dmul_rn(dbest, dbest), dg2);
  if (d2 < best_d2) { best_d2 = d2; best_y = ys[(size_t)ibest * C]; }
}

// branch-free searchsorted(side='left') over a compacted column: #nodes < x.  top = tb.top.
template <typename TX, typename T, int C>
__device__ __forceinline__ int lower_bound_bf(const T* xs, int n, TX x, int top) {
  int pos = 0;
  for (int step = top; step > 0; step >>= 1) {
    const int p2 = pos + step;
    if (p2 <= n && (TX)xs[(size_t)(p2 - 1) * C] < x) pos = p2;
  }
  return pos;
}

// Cross-row part of the 2-D nearest rule (rare: only when the in-row nearest node is >= 1 away).
template <typename TX, typename T, int C>
__device__ __noinline__ T nearest_cross_rows(const Tables<T, C>& tb, int c, long long pt, int r, TX x, double best_d2,
                                             T best_y) {
  const int nq = tb.nq;
  // padded row coordinate of the sample is r+1 in [1, G]; padded rows exist for 0..G+1
  for (int dist = 1; dist <= tb.G + 1; ++dist) {
    const double dg2 = (double)dist * (double)dist;
    if (dg2 >= best_d2) break;  // nothing at this row distance can be strictly nearer
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      const int pr = r + 1 + sgn * dist;  // padded row index
      if (pr < 0 || pr > tb.G + 1) continue;
      if (dist == 1 && !tb.centre_only) {
        const int slot = 1 + sgn;
        nearest_in_row<TX, T, C>(tb.xsl[slot] + c, tb.ysl[slot] + c,
                                 tb.nvl[slot][c], x, dg2, best_d2, best_y);
      } else {
        // scan the raw row in global memory (padded row pr is group (pr-1) mod G)
        const int g = (pr - 1 + tb.G) % tb.G;
        const T* gx = tb.x_shared ? tb.gx : tb.gx + pt * tb.pt_stride + (long long)g * nq;
        const T* gy = tb.gy + pt * tb.pt_stride + (long long)g * nq;
        for (int k = 0; k < nq; ++k) {
          const T xv = gx[k], yv = gy[k];
          if (is_nan(xv) || is_nan(yv)) continue;
          const double d = fabs((double)x - (double)xv);
          const double d2 = __dadd_rn(__dmul_rn(d, d), dg2);
          if (d2 < best_d2) { best_d2 = d2; best_y = yv; }
        }
      }
    }
  }
  return best_y;
}

// 2-D rule = scipy griddata(method="nearest") on points (hist_q, group coordinate) in raw units
// + _extrapolate_on_quantiles (utils.py:380-400, 477-513; nbutils.py:392-416), for N samples of the
// same point and group at once (independent searches interleave).  r is the 0-based group; rows are
// the cyclically padded table rows r-1..r+1 (+ global fallback).  The extrapolation override of the
// reference is unconditional, so it is tested first and the search skipped for out-of-range samples
// (_extrapolate_on_quantiles with an integer group coordinate: np.interp returns the row's own bound).
template <typename TX, typename T, int C, int N>
__device__ __forceinline__ void lookup_2d_nearest_n(const Tables<T, C>& tb, int c, long long pt, int r,
                                                    const TX (&x)[N], T (&out)[N], int extrap) {
  const int nq = tb.nq;
  const T* xs = tb.xsl[1] + c;
  const T* ys = tb.ysl[1] + c;
  const int n = tb.nvl[1][c];
  const double blo = (double)tb.blo[c], bhi = (double)tb.bhi[c];
  int pos[N];
#pragma unroll
  for (int j = 0; j < N; ++j) pos[j] = 0;
  for (int step = tb.top; step > 0; step >>= 1) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int p2 = pos[j] + step;
      if ((TX)xs[(size_t)(p2 - 1) * C] < x[j]) pos[j] = p2;  // rows >= n hold +inf
    }
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const double xd = (double)x[j];
    T res;
    if (is_nan(x[j])) {
      res = Num<T>::nan();
    } else if (xd < blo) {
      res = extrap == 0 ? tb.clo[c] : Num<T>::nan();
    } else if (xd > bhi) {
      res = extrap == 0 ? tb.chi[c] : Num<T>::nan();
    } else {
      const int i = pos[j];
      double dbest = __longlong_as_double(0x7ff0000000000000LL);
      T ybest = Num<T>::nan();
      if (i > 0) { dbest = fabs(xd - (double)xs[(size_t)(i - 1) * C]); ybest = ys[(size_t)(i - 1) * C]; }
      if (i < n) {
        const double d = fabs((double)xs[(size_t)i * C] - xd);
        if (d < dbest) { dbest = d; ybest = ys[(size_t)i * C]; }
      }
      res = ybest;
      if (!(dbest < 1.0))  // a node of a neighbouring row (>= 1 away in the group coordinate) may be nearer
        res = nearest_cross_rows<TX, T, C>(tb, c, pt, r, x[j], __dmul_rn(dbest, dbest), ybest);
    }
    out[j] = res;
  }
}

template <typename TX, typename T, int C>
__device__ __forceinline__ T lookup_2d_nearest(const Tables<T, C>& tb, int c, long long pt, int r, TX x, int extrap) {
  const TX xa[1] = {x};
  T o[1];
  lookup_2d_nearest_n<TX, T, C, 1>(tb, c, pt, r, xa, o, extrap);
  return o[0];
}

// 2-D "linear" rule on the shared quantile axis (QDM, _adjustment.py:873-880): scipy griddata(method="linear") =
// LinearNDInterpolator on the regular lattice (q_k, g'), g' the cyclically padded group coordinate.  Qhull splits
// every lattice cell into two triangles; which diagonal it uses is data independent and comes from the host
// (diag[(padded row) * (nq-1) + k], 0: (r,k)-(r+1,k+1), 1: (r,k+1)-(r+1,k)).  Barycentric weights in float64.
// Staged slots 0,1,2 hold padded rows g, g+1, g+2 (no NaN factors: the host checks).  newg is the fractional
// padded coordinate of the sample (Grouper.get_index(interp=True), base.py:306-320).
template <typename T, int C>
__device__ T lookup_2d_linear_shared(const Tables<T, C>& tb, int c, int g, double x, double newg,
                                     const unsigned char* __restrict__ diag, int extrap) {
  if (x != x || newg != newg) return Num<T>::nan();
  const int nq = tb.nq;
  const int pr0 = (int)floor(newg);
  const double v = newg - (double)pr0;
  int s0 = pr0 - g;                       // slot of the lower row
  if (s0 < 0 || s0 > 2 || (v > 0.0 && s0 > 1)) return Num<T>::nan();
  const int s1 = v > 0.0 ? s0 + 1 : s0;
  const T* xs = tb.xsl[1] + c;            // the quantile axis is the same in every row
  const T* y0 = tb.ysl[s0] + c;
  const T* y1 = tb.ysl[s1] + c;
  const double q_first = (double)xs[0], q_last = (double)xs[(size_t)(nq - 1) * C];
  if (x < q_first || x > q_last) {
    // outside the convex hull -> NaN from griddata, then _extrapolate_on_quantiles (nbutils.py:397-416) unless "nan"
    if (extrap != 0) return Num<T>::nan();
    const int k = x < q_first ? 0 : nq - 1;
    const double f0 = (double)y0[(size_t)k * C], f1 = (double)y1[(size_t)k * C];
    if (v == 0.0) return (T)f0;
    const double slope = __ddiv_rn(__dsub_rn(f1, f0), 1.0);   // np.interp between two padded rows
    return (T)__dadd_rn(__dmul_rn(slope, v), f0);
  }
  int k = lower_bound_bf<double, T, C>(xs, nq, x, tb.top) - 1;  // last node <= x ... (nodes < x) - 1
  if (k < 0) k = 0;
  if (k > nq - 2) k = nq - 2;
  const double qk = (double)xs[(size_t)k * C], qk1 = (double)xs[(size_t)(k + 1) * C];
  const double u = (x - qk) / (qk1 - qk);
  const double yA = (double)y0[(size_t)k * C], yB = (double)y0[(size_t)(k + 1) * C];
  if (v == 0.0) return (T)(yA * (1.0 - u) + yB * u);
  const double yD = (double)y1[(size_t)k * C], yC = (double)y1[(size_t)(k + 1) * C];
  double r;
  if (diag[(size_t)pr0 * (nq - 1) + k] == 0) {   // diagonal A-C
    r = v <= u ? yA * (1.0 - u) + yB * (u - v) + yC * v : yA * (1.0 - v) + yD * (v - u) + yC * u;
  } else {                                        // diagonal B-D
    r = u + v <= 1.0 ? yA * (1.0 - u - v) + yB * u + yD * v : yC * (u + v - 1.0) + yB * (1.0 - v) + yD * (1.0 - u);
  }
  return (T)r;
}

template <typename T> __device__ __forceinline__ T apply_corr(T x, T f, int kind) {
  return kind == 43 ? Num<T>::add(x, f) : Num<T>::mul(x, f);  // utils.py:146-162
}

}  // namespace xsdba
