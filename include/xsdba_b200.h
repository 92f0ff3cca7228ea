/*
 * xsdba_b200 -- C ABI of the B200 (sm_100a) quantile-mapping hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy / xarray types.
 * Each entry point names the interface of the reference (Ouranosinc/xsdba v0.7.0, paths relative
 * to src/xsdba/) that it replaces.  The Python host layer (xsdba_b200/_adjustment.py) binds these
 * with ctypes and mirrors the reference's L4 functions (eqm_train, dqm_train, qm_adjust,
 * dqm_adjust, qdm_adjust); INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * Conventions
 *  - "dev" pointers are CUDA device pointers, "host" pointers are ordinary host memory.
 *  - Series arrays are addressed by element strides: element (point p, time t) of an array x is
 *    x[p*stride_pt + t*stride_time].  The reference's natural (time, lat, lon) C order is
 *    stride_pt = 1, stride_time = n_pts ("time-major"); (lat, lon, time) is stride_pt = n_time,
 *    stride_time = 1 ("point-major").  Kernels are tuned for time-major.
 *  - Trained tables are point-major: af / hist_q are [n_pts][n_groups][nq], scaling is
 *    [n_pts][n_groups]  (the reference's output dims (<points>, month|dayofyear|group, quantiles),
 *    base.py:652-694).
 *  - kind: '+' (43) or '*' (42), as in utils.get_correction / apply_correction (utils.py:130-162).
 *  - Every function returns 0 on success, a negative XSDBA_ERR_* for argument errors, or a
 *    positive cudaError_t.  Nothing throws across the ABI.  Functions are re-entrant; the only
 *    state is the caller-owned grouping handle (immutable after creation, bound to the device it was
 *    created on: calls from another current device return XSDBA_ERR_INVALID_ARGUMENT), the caller's
 *    stream and caller-owned workspaces.  Per-call scratch of the device entry points comes from the
 *    stream-ordered memory pool of the current device (cudaMallocAsync on the caller's stream).
 *  - Inputs are never modified (the reference's numba quantile sorts its input in place when the
 *    reshape is a view; this library does not).
 */
#ifndef XSDBA_B200_H
#define XSDBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XSDBA_OK 0
#define XSDBA_ERR_INVALID_ARGUMENT (-1)
#define XSDBA_ERR_UNSUPPORTED (-2)      /* valid in the reference, not built here yet (e.g. grouped linear) */
#define XSDBA_ERR_SEGMENT_TOO_LONG (-3) /* a (point, group) segment exceeds the shared-memory sorter */
#define XSDBA_ERR_NO_DEVICE (-4)
#define XSDBA_ERR_OUT_OF_MEMORY (-5)

#define XSDBA_KIND_ADD 43 /* '+' */
#define XSDBA_KIND_MUL 42 /* '*' */

#define XSDBA_INTERP_NEAREST 0
#define XSDBA_INTERP_LINEAR 1
#define XSDBA_INTERP_CUBIC 2   /* group = "time" only: scipy interp1d(kind="cubic"), the not-a-knot cubic spline */

#define XSDBA_EXTRAP_CONSTANT 0
#define XSDBA_EXTRAP_NAN 1

#define XSDBA_MAX_SEGMENT 32768 /* longest (point, group) segment the in-SM sorter takes (float32) */

typedef struct xsdba_grouping xsdba_grouping_t;

/* Library / build information. */
int xsdba_version(void);
const char* xsdba_status_string(int status);
/* Number of kernels this library has launched in this process (all streams); bench.py reports the
 * delta over its timed region as "gpu_launches". */
int64_t xsdba_launch_count(void);

/*
 * Grouping handle: replaces base.Grouper.group / get_index / apply's membership logic
 * (base.py:232-345, 410-420) and the rolling(center=True).construct window gather (base.py:261-265).
 *
 * grp_idx_host[t] is the 0-based group of time step t (month-1, dayofyear-1, 0 for group="time"),
 * or -1 for "in no group".  window is the Grouper window (odd or even, >= 1): a reducing function
 * sees, for every member t of a group, the samples x[t - window/2 + j], j = 0..window-1, NaN outside
 * [0, n_time) -- positional, like xarray.  The handle owns small device tables derived from this.
 */
int xsdba_grouping_create(xsdba_grouping_t** out, const int32_t* grp_idx_host, int64_t n_time,
                          int32_t n_groups, int32_t window);
int xsdba_grouping_destroy(xsdba_grouping_t* g);
int64_t xsdba_grouping_max_segment(const xsdba_grouping_t* g); /* longest segment incl. window slots */
int32_t xsdba_grouping_n_groups(const xsdba_grouping_t* g);

/*
 * Train: replaces _adjustment.eqm_train.func (_adjustment.py:253-286; normalize = 0) and
 * _adjustment.dqm_train.func (_adjustment.py:150-190; normalize = 1) for every group at once, i.e.
 * Grouper.apply + nbutils.quantile (nbutils.py:108-148, 198-271) + utils.get_correction
 * (utils.py:130-143).  q_dev holds nq nodes already cast to the data dtype (nbutils.py:253).
 * scaling_dev may be NULL when normalize = 0.  Groups without members give NaN rows.
 */
int xsdba_qm_train_f32(const float* ref_dev, const float* hist_dev, int64_t n_pts, int64_t stride_pt,
                       int64_t stride_time, const xsdba_grouping_t* grp, const float* q_dev, int32_t nq,
                       int32_t kind, int32_t normalize, float* af_dev, float* hist_q_dev,
                       float* scaling_dev, void* cuda_stream);
int xsdba_qm_train_f64(const double* ref_dev, const double* hist_dev, int64_t n_pts, int64_t stride_pt,
                       int64_t stride_time, const xsdba_grouping_t* grp, const double* q_dev, int32_t nq,
                       int32_t kind, int32_t normalize, double* af_dev, double* hist_q_dev,
                       double* scaling_dev, void* cuda_stream);

/*
 * Train with the jitter pre-step of _preprocess_dataset (_adjustment.py:48-83): jitter4_host =
 * {lower, minimum, upper, maximum} in the data's units (NaN lower / upper disables that side; minimum
 * is the already next-after'ed lower bound, processing.py:222-224).  As in the reference the noise is
 * applied to `hist` only, after the window gather, independently for every window slot.  The draws are a
 * counter-based hash of (seed, element): reproducible per call, distributionally equal to the
 * reference's numpy.random.uniform draws (which are not reproducible by design).
 */
int xsdba_qm_train_jitter_f32(const float* ref_dev, const float* hist_dev, int64_t n_pts, int64_t stride_pt,
                              int64_t stride_time, const xsdba_grouping_t* grp, const float* q_dev, int32_t nq,
                              int32_t kind, int32_t normalize, const double* jitter4_host, uint64_t seed,
                              float* af_dev, float* hist_q_dev, float* scaling_dev, void* cuda_stream);
int xsdba_qm_train_jitter_f64(const double* ref_dev, const double* hist_dev, int64_t n_pts, int64_t stride_pt,
                              int64_t stride_time, const xsdba_grouping_t* grp, const double* q_dev, int32_t nq,
                              int32_t kind, int32_t normalize, const double* jitter4_host, uint64_t seed,
                              double* af_dev, double* hist_q_dev, double* scaling_dev, void* cuda_stream);
/* Elementwise processing.jitter (processing.py:180-257) over n contiguous elements. */
int xsdba_jitter_f32(const float* x_dev, int64_t n, const double* jitter4_host, uint64_t seed, float* out_dev,
                     void* cuda_stream);
int xsdba_jitter_f64(const double* x_dev, int64_t n, const double* jitter4_host, uint64_t seed, double* out_dev,
                     void* cuda_stream);

/*
 * Frequency adaptation (the adapt_freq_thresh option of the three QM classes).
 *  - xsdba_qm_train_adapt_*: eqm_train with _preprocess_dataset's adapt_freq step (_adjustment.py:69-70 ->
 *    _processing._adapt_freq.func, _processing.py:75-131) on hist inside every group (after jitter, after the
 *    window gather): P0_ref / P0_hist = ecdf at the threshold (float64 [n_pts][n_groups]), pth =
 *    vecquantiles(ref, P0_hist) where dP0 > 0 else NaN (data dtype), the excess dry values of hist replaced
 *    by U(thresh, pth) before its quantiles are taken.  P0_ref, P0_hist, pth are deterministic (bit-exact
 *    against the oracle); which tied values are replaced and the fill values are hashes of (seed, element).
 *  - xsdba_adapt_freq_apply_*: the adjust-side _adapt_freq_preprocess (_adjustment.py:32-45, 639-646) on
 *    sim with the stored P0_ref / P0_hist / pth, per exact group; out has the strides of sim.
 *  - xsdba_tail_mask_*: the max_tail_factor mask (_adjustment.py:647-658, 672-673): scen = adapted sim
 *    wherever adapted sim > factor * last node of hist_q_raw of the sample's group (nearest broadcast).
 */
int xsdba_qm_train_adapt_f32(const float* ref_dev, const float* hist_dev, int64_t n_pts, int64_t stride_pt,
                             int64_t stride_time, const xsdba_grouping_t* grp, const float* q_dev, int32_t nq,
                             int32_t kind, const double* jitter4_host, double adapt_thresh, uint64_t seed,
                             float* af_dev, float* hist_q_dev, double* P0_ref_dev, double* P0_hist_dev,
                             float* pth_dev, void* cuda_stream);
int xsdba_qm_train_adapt_f64(const double* ref_dev, const double* hist_dev, int64_t n_pts, int64_t stride_pt,
                             int64_t stride_time, const xsdba_grouping_t* grp, const double* q_dev, int32_t nq,
                             int32_t kind, const double* jitter4_host, double adapt_thresh, uint64_t seed,
                             double* af_dev, double* hist_q_dev, double* P0_ref_dev, double* P0_hist_dev,
                             double* pth_dev, void* cuda_stream);
/* dqm_train with adapt_freq_thresh (_adjustment.py:95-192): hist is frequency-adapted first, then both series are
 * normalised by their group means (the ADAPTED hist's mean); scaling_dev [n_pts][n_groups] as xsdba_qm_train_*. */
int xsdba_dqm_train_adapt_f32(const float* ref_dev, const float* hist_dev, int64_t n_pts, int64_t stride_pt,
                              int64_t stride_time, const xsdba_grouping_t* grp, const float* q_dev, int32_t nq,
                              int32_t kind, const double* jitter4_host, double adapt_thresh, uint64_t seed,
                              float* af_dev, float* hist_q_dev, float* scaling_dev, double* P0_ref_dev,
                              double* P0_hist_dev, float* pth_dev, void* cuda_stream);
int xsdba_dqm_train_adapt_f64(const double* ref_dev, const double* hist_dev, int64_t n_pts, int64_t stride_pt,
                              int64_t stride_time, const xsdba_grouping_t* grp, const double* q_dev, int32_t nq,
                              int32_t kind, const double* jitter4_host, double adapt_thresh, uint64_t seed,
                              double* af_dev, double* hist_q_dev, double* scaling_dev, double* P0_ref_dev,
                              double* P0_hist_dev, double* pth_dev, void* cuda_stream);
int xsdba_adapt_freq_apply_f32(const float* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                               const xsdba_grouping_t* grp, double thresh, const double* P0_ref_dev,
                               const double* P0_hist_dev, const float* pth_dev, uint64_t seed, float* out_dev,
                               void* cuda_stream);
int xsdba_adapt_freq_apply_f64(const double* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                               const xsdba_grouping_t* grp, double thresh, const double* P0_ref_dev,
                               const double* P0_hist_dev, const double* pth_dev, uint64_t seed, double* out_dev,
                               void* cuda_stream);
int xsdba_tail_mask_f32(const float* adapted_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                        const xsdba_grouping_t* grp, const float* hist_q_raw_dev, int32_t nq, double factor,
                        float* scen_dev, void* cuda_stream);
int xsdba_tail_mask_f64(const double* adapted_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                        const xsdba_grouping_t* grp, const double* hist_q_raw_dev, int32_t nq, double factor,
                        double* scen_dev, void* cuda_stream);

/*
 * Quantiles only: replaces nbutils.quantile over grouped segments (nbutils.py:224-271), used for
 * hist_q_raw (_adjustment.py:254-256) and by callers that want ref_q.  out is [n_pts][n_groups][nq].
 */
int xsdba_group_quantile_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                             const xsdba_grouping_t* grp, const float* q_dev, int32_t nq, float* out_dev,
                             void* cuda_stream);
int xsdba_group_quantile_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                             const xsdba_grouping_t* grp, const double* q_dev, int32_t nq, double* out_dev,
                             void* cuda_stream);

/*
 * Adjust (EQM / DQM flavour): replaces _adjustment.qm_adjust.func (_adjustment.py:660-669) =
 * utils.interp_on_quantiles (utils.py:408-513; 1-D SciPy interp1d rule for group="time", 2-D
 * Euclidean-nearest griddata rule + nbutils._extrapolate_on_quantiles for month/dayofyear groups,
 * with add_cyclic_bounds, utils.py:284-314) followed by utils.apply_correction (utils.py:146-162).
 * grp gives the group of each sim time step (its window is ignored).  n_groups == 1 selects the
 * 1-D rule.  Grouped interp = LINEAR returns XSDBA_ERR_UNSUPPORTED (Qhull path).
 * scen has the same strides as sim.
 */
int xsdba_qm_adjust_f32(const float* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                        const xsdba_grouping_t* grp, const float* af_dev, const float* hist_q_dev,
                        int32_t nq, int32_t interp, int32_t extrap, int32_t kind, float* scen_dev,
                        void* cuda_stream);
int xsdba_qm_adjust_f64(const double* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                        const xsdba_grouping_t* grp, const double* af_dev, const double* hist_q_dev,
                        int32_t nq, int32_t interp, int32_t extrap, int32_t kind, double* scen_dev,
                        void* cuda_stream);

/*
 * Adjust (QDM): replaces _adjustment.qdm_adjust.func (_adjustment.py:872-881): per-group percentile
 * ranks (utils.rank with pct=True -> bottleneck.nanrankdata, utils.py:612-638; Grouper.apply with
 * main_only = !rank_window, base.py:438-439), factor lookup on the shared quantile axis, then
 * apply_correction.  sim_q_dev (float64, same strides as sim) may be NULL.
 */
int xsdba_qdm_adjust_f32(const float* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const float* af_dev, const float* q_dev, int32_t nq,
                         int32_t interp, int32_t extrap, int32_t kind, int32_t rank_window,
                         float* scen_dev, double* sim_q_dev, void* cuda_stream);
int xsdba_qdm_adjust_f64(const double* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const double* af_dev, const double* q_dev, int32_t nq,
                         int32_t interp, int32_t extrap, int32_t kind, int32_t rank_window,
                         double* scen_dev, double* sim_q_dev, void* cuda_stream);

/*
 * Adjust (QDM) with grouped interp="linear" (SURVEY.md 8f rank 2): the point set of SciPy's
 * LinearNDInterpolator is the regular lattice (quantile node, padded group coordinate), identical for every
 * gridpoint (_adjustment.py:873-880), so the Qhull triangulation is computed once on the host
 * (scipy.spatial.Delaunay, the reference's own dependency) and only its per-cell diagonal choice is uploaded:
 * diag_dev[(n_groups+1) * (nq-1)] uint8, 0 = diagonal (r,k)-(r+1,k+1), 1 = (r,k+1)-(r+1,k).  gcoord_dev[n_time]
 * is the fractional padded group coordinate of every time step (Grouper.get_index(interp=True),
 * base.py:306-320).  Factors must be NaN free.  (EQM/DQM grouped linear needs a per-gridpoint triangulation:
 * still XSDBA_ERR_UNSUPPORTED.)
 */
int xsdba_qdm_adjust_linear_f32(const float* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                                const xsdba_grouping_t* grp, const float* af_dev, const float* q_dev, int32_t nq,
                                int32_t extrap, int32_t kind, int32_t rank_window, const double* gcoord_dev,
                                const unsigned char* diag_dev, float* scen_dev, double* sim_q_dev,
                                void* cuda_stream);
int xsdba_qdm_adjust_linear_f64(const double* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                                const xsdba_grouping_t* grp, const double* af_dev, const double* q_dev, int32_t nq,
                                int32_t extrap, int32_t kind, int32_t rank_window, const double* gcoord_dev,
                                const unsigned char* diag_dev, double* scen_dev, double* sim_q_dev,
                                void* cuda_stream);

/*
 * Percentile ranks only: replaces Grouper.apply(utils.rank, x, main_only = !rank_window, pct = True).
 */
int xsdba_group_rank_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, int32_t rank_window, double* rank_dev,
                         void* cuda_stream);
int xsdba_group_rank_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, int32_t rank_window, double* rank_dev,
                         void* cuda_stream);

/*
 * End-to-end host entry points (what the xarray-facing layer calls with numpy buffers; they stand for
 * TrainAdjust.train + .adjust on in-memory data, adjustment.py:226-316): train(ref, hist) + adjust(sim) on HOST
 * arrays in the reference's (time, points) C order (time-major, contiguous).  Points are streamed through the GPU in
 * slabs on three streams created by the call (H2D copy, train, adjust, D2H copy overlapped); af / hist_q / scaling
 * (point-major, may be NULL) and scen (time-major) are written back to host.  Pinned host memory gives full PCIe
 * rate; pageable works.
 *   method 0 = EQM (eqm_train + qm_adjust), 1 = QDM (eqm_train + qdm_adjust; ranks over the window when grp_sim
 *   carries one, i.e. rank_window=True), 2 = DQM (dqm_train + dqm_adjust with PolyDetrend(detrend_degree) on the
 *   adjustment group; grp_sim then carries the Grouper window, sim_tcoord_host the sim time coordinate in days).
 * grp_train carries the Grouper window of the training step; grp_sim describes the sim time axis.
 * State: none.  The staging buffers live in a caller-owned DEVICE workspace of at least
 * xsdba_qm_train_adjust_host_workspace_bytes(...) bytes (any device allocation, e.g. a torch tensor); concurrent
 * calls need distinct workspaces.  xsdba_qm_train_adjust_host_f32 is the convenience form (EQM / QDM, float32) that
 * allocates and frees its workspace itself.
 */
int64_t xsdba_qm_train_adjust_host_workspace_bytes(int64_t n_pts, const xsdba_grouping_t* grp_train,
                                                   const xsdba_grouping_t* grp_sim, int32_t nq, int32_t elem_size,
                                                   int32_t method, int64_t slab_pts);
int xsdba_qm_train_adjust_host_ws_f32(const float* ref_host, const float* hist_host, const float* sim_host,
                                      int64_t n_pts, const xsdba_grouping_t* grp_train,
                                      const xsdba_grouping_t* grp_sim, const float* q_host, int32_t nq, int32_t kind,
                                      int32_t method, int32_t interp, int32_t extrap, int32_t detrend_degree,
                                      const double* sim_tcoord_host, float* scen_host, float* af_host,
                                      float* hist_q_host, float* scaling_host, int64_t slab_pts, void* workspace_dev,
                                      int64_t workspace_bytes);
int xsdba_qm_train_adjust_host_ws_f64(const double* ref_host, const double* hist_host, const double* sim_host,
                                      int64_t n_pts, const xsdba_grouping_t* grp_train,
                                      const xsdba_grouping_t* grp_sim, const double* q_host, int32_t nq, int32_t kind,
                                      int32_t method, int32_t interp, int32_t extrap, int32_t detrend_degree,
                                      const double* sim_tcoord_host, double* scen_host, double* af_host,
                                      double* hist_q_host, double* scaling_host, int64_t slab_pts,
                                      void* workspace_dev, int64_t workspace_bytes);
int xsdba_qm_train_adjust_host_f32(const float* ref_host, const float* hist_host, const float* sim_host,
                                   int64_t n_pts, const xsdba_grouping_t* grp_train,
                                   const xsdba_grouping_t* grp_sim, const float* q_host, int32_t nq,
                                   int32_t kind, int32_t mode, int32_t interp, int32_t extrap,
                                   float* scen_host, float* af_host, float* hist_q_host,
                                   int64_t slab_pts);

/*
 * MBCn / N-pdf transform building blocks (_npdft_train / _npdft_adjust / mbcn_adjust,
 * _adjustment.py:289-328, 426-464, 467-591).  The per-iteration loop is host code
 * (xsdba_b200/mbcn.py); every array op inside it is one of these kernels.
 *  - xsdba_qm_train_q64_f32: eqm_train with float64 quantile nodes on float32 data -- _npdft_train calls
 *    nbutils._quantile with the un-cast float64 nodes (_adjustment.py:315), so the virtual index uses them.
 *  - xsdba_rank_lookup_*: xsdba_qdm_adjust_* with the rank normalisation selectable: rank_mode 0 =
 *    utils.rank(pct=True) (utils.py:629-634), 1 = utils._rank_bn (utils.py:641-646), as used at
 *    _adjustment.py:317-323, 453-459 (x + af looked up at the rank of x).
 *  - xsdba_rotate_*: y[v] = sum_w rot[v][w] * x[w] over n_var stacked variables of n_elem elements each
 *    (rot @ x, _adjustment.py:311, 449; rot_host is n_var x n_var float32 row-major, n_var <= 8; y != x).
 *  - xsdba_standardize_*: (x - nanmean) / nanstd (ddof = 0) along time for every (variable, point)
 *    (processing.standardize, processing.py:323-350; _adjustment.py:303-305); variables var_stride apart.
 *  - xsdba_reorder_*: Schaake shuffle sort(sim)[argsort(argsort(ref))] per (point, group); with a window
 *    the [time, window] segment is flattened and the centre column kept (_processing.py:204-211).
 */
int xsdba_qm_train_q64_f32(const float* ref_dev, const float* hist_dev, int64_t n_pts, int64_t stride_pt,
                           int64_t stride_time, const xsdba_grouping_t* grp, const double* q64_dev, int32_t nq,
                           int32_t kind, float* af_dev, float* hist_q_dev, void* cuda_stream);
/* One variable of one N-pdf iteration, fused (float32 series, one group = one block of time steps):
 *   ref_dev != NULL (_npdft_train, _adjustment.py:313-324): af = quantile(ref) - quantile(x) at the float64 nodes is
 *   WRITTEN to af_io_dev [n_pts][nq], then x += interp1d(rank_bn(x), q, af) in place;
 *   ref_dev == NULL (_npdft_adjust, _adjustment.py:451-460): af is READ from af_io_dev, x updated in place.
 * The lookup and the sum run in float64 like the reference's (float64 af_q and nodes), the keys stay float32. */
int xsdba_npdft_step_f32(const float* ref_dev, float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const double* q64_dev, int32_t nq, int32_t interp,
                         int32_t extrap, float* af_io_dev, void* cuda_stream);
int xsdba_rank_lookup_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                          const xsdba_grouping_t* grp, const float* af_dev, const float* q_dev, int32_t nq,
                          int32_t interp, int32_t extrap, int32_t kind, int32_t rank_window, int32_t rank_mode,
                          float* out_dev, double* rank_dev, void* cuda_stream);
int xsdba_rank_lookup_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                          const xsdba_grouping_t* grp, const double* af_dev, const double* q_dev, int32_t nq,
                          int32_t interp, int32_t extrap, int32_t kind, int32_t rank_window, int32_t rank_mode,
                          double* out_dev, double* rank_dev, void* cuda_stream);
int xsdba_rotate_f32(const float* x_dev, int64_t n_elem, int32_t n_var, const float* rot_host, float* y_dev,
                     void* cuda_stream);
int xsdba_rotate_f64(const double* x_dev, int64_t n_elem, int32_t n_var, const float* rot_host, double* y_dev,
                     void* cuda_stream);
/* The same product with the multiply and the add rounded separately, terms in ascending order: numpy's
 * einsum("ij,j...->i...") as _npdft_adjust calls it (_adjustment.py:449, 462); xsdba_rotate_* is the fused chain of
 * `rot @ x` in _npdft_train (_adjustment.py:311).  The N-pdf iteration amplifies one-ulp differences, so both exist. */
int xsdba_rotate_unfused_f32(const float* x_dev, int64_t n_elem, int32_t n_var, const float* rot_host, float* y_dev,
                             void* cuda_stream);
int xsdba_rotate_unfused_f64(const double* x_dev, int64_t n_elem, int32_t n_var, const float* rot_host, double* y_dev,
                             void* cuda_stream);
int xsdba_standardize_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                          int64_t n_time, int32_t n_var, int64_t var_stride, float* y_dev, void* cuda_stream);
int xsdba_standardize_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                          int64_t n_time, int32_t n_var, int64_t var_stride, double* y_dev, void* cuda_stream);
int xsdba_reorder_f32(const float* sim_dev, const float* ref_dev, int64_t n_pts, int64_t stride_pt,
                      int64_t stride_time, const xsdba_grouping_t* grp, float* out_dev, void* cuda_stream);
int xsdba_reorder_f64(const double* sim_dev, const double* ref_dev, int64_t n_pts, int64_t stride_pt,
                      int64_t stride_time, const xsdba_grouping_t* grp, double* out_dev, void* cuda_stream);

/*
 * Per-(point, group) selections on the segment sorter.
 *  - xsdba_group_vecquantile_*: replaces nbutils.vecquantiles / _vecquantiles (nbutils.py:151-195): one
 *    numba np.nanquantile(segment, rnk[point][group]) per segment (NaN rank -> NaN); rnk / out are
 *    [n_pts][n_groups].  Used by _adapt_freq (_processing.py:107).
 *  - xsdba_map_cdf_*: replaces utils.map_cdf / map_cdf_1d / _ecdf_1d (utils.py:35-84) per group: the value
 *    of x with the same empirical CDF as y_value in y, for nv values (yvals_dev, float64 device array);
 *    out is [n_pts][n_groups][nv].
 */
int xsdba_group_vecquantile_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                                const xsdba_grouping_t* grp, const float* rnk_dev, float* out_dev, void* cuda_stream);
int xsdba_group_vecquantile_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                                const xsdba_grouping_t* grp, const double* rnk_dev, double* out_dev, void* cuda_stream);
int xsdba_map_cdf_f32(const float* x_dev, const float* y_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                      const xsdba_grouping_t* grp, const double* yvals_dev, int32_t nv, float* out_dev,
                      void* cuda_stream);
int xsdba_map_cdf_f64(const double* x_dev, const double* y_dev, int64_t n_pts, int64_t stride_pt,
                      int64_t stride_time, const xsdba_grouping_t* grp, const double* yvals_dev, int32_t nv,
                      double* out_dev, void* cuda_stream);

/*
 * Energy score: replaces processing.escore / nbutils._escore (processing.py:393-489; nbutils.py:274-372;
 * SURVEY.md 8f rank 3) without the optional scaling: tgt and sim are (variable, time, point) arrays, variables
 * var_stride_* elements apart, n_var <= 8; observations with a NaN in any variable are dropped; n_sub > 0
 * keeps about n_sub evenly spaced observations of each cloud.  out_dev[n_pts].
 */
int xsdba_escore_f32(const float* tgt_dev, const float* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                     int64_t n_time_tgt, int64_t n_time_sim, int32_t n_var, int64_t var_stride_tgt,
                     int64_t var_stride_sim, int32_t n_sub, float* out_dev, void* cuda_stream);
int xsdba_escore_f64(const double* tgt_dev, const double* sim_dev, int64_t n_pts, int64_t stride_pt,
                     int64_t stride_time, int64_t n_time_tgt, int64_t n_time_sim, int32_t n_var,
                     int64_t var_stride_tgt, int64_t var_stride_sim, int32_t n_sub, double* out_dev,
                     void* cuda_stream);

/*
 * Polynomial trend: replaces detrending.PolyDetrend.fit(...).ds.trend = _polydetrend_get_trend
 * (detrending.py:165-208; xarray polyfit/polyval per group through map_groups): y = x (+|*)
 * scaling[point][group] when scaling_dev != NULL (the scaled_sim of dqm_adjust, _adjustment.py:748-757),
 * the Grouper window is averaged NaN-skipping first (detrending.py:199-200), least squares of degree
 * 0..4 on tcoord_dev[n_time] (any monotone float64 time coordinate, e.g. days), evaluated at every
 * member.  trend_dev is float64 with the strides of x.
 */
int xsdba_poly_trend_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const float* scaling_dev, int32_t kind, int32_t degree,
                         const double* tcoord_dev, double* trend_dev, void* cuda_stream);
int xsdba_poly_trend_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const double* scaling_dev, int32_t kind, int32_t degree,
                         const double* tcoord_dev, double* trend_dev, void* cuda_stream);

/*
 * LOESS trend over the whole series: replaces detrending.LoessDetrend(group="time").fit(...).ds.trend =
 * loess.loess_smoothing -> numba _loess_nb (loess.py:49-179, 182-279; detrending.py:211-296) in its
 * equal-spacing, tricube, skipna form with local degree d in {0, 1}; niter >= 1 (iterations after the first
 * re-weight every sample with the bisquare of its residual over 6 x the median absolute residual,
 * loess.py:166-176).  y = x (+|*) scaling[point][group(t)] when scaling_dev != NULL (grp
 * supplies group(t); pass a single-group handle otherwise).  xn_dev[n_time] is the time coordinate
 * rescaled to [0, 1] (loess.py:244-245).  trend_dev is float64 with the strides of x; NaN where x is NaN.
 */
int xsdba_loess_trend_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                          const xsdba_grouping_t* grp, const float* scaling_dev, int32_t kind, double f,
                          int32_t niter, int32_t degree, const double* xn_dev, double* trend_dev,
                          void* cuda_stream);
int xsdba_loess_trend_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                          const xsdba_grouping_t* grp, const double* scaling_dev, int32_t kind, double f,
                          int32_t niter, int32_t degree, const double* xn_dev, double* trend_dev,
                          void* cuda_stream);
/* The same with the weight function and the spacing branch selectable: weights 0 = tricube (loess.py:29-35),
 * 1 = gaussian (loess.py:16-26, `LoessDetrend(weights="gaussian")`, loess.py:247); equal_spacing 1 = the dx > 0 branch
 * above, 0 = the dx == 0 branch (loess.py:107-111, 151-158: r = round(f n), bandwidth = distance of the r-th closest
 * sample, weights recomputed for every output), which is what an irregular time axis and every grouped LoessDetrend
 * (the members of a month or season are not equally spaced) take. */
int xsdba_loess_trend_w_f32(const float* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                            const xsdba_grouping_t* grp, const float* scaling_dev, int32_t kind, double f,
                            int32_t niter, int32_t degree, int32_t weights, int32_t equal_spacing,
                            const double* xn_dev, double* trend_dev, void* cuda_stream);
int xsdba_loess_trend_w_f64(const double* x_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                            const xsdba_grouping_t* grp, const double* scaling_dev, int32_t kind, double f,
                            int32_t niter, int32_t degree, int32_t weights, int32_t equal_spacing,
                            const double* xn_dev, double* trend_dev, void* cuda_stream);

/*
 * Adjust (DQM): replaces _adjustment.dqm_adjust.func (_adjustment.py:748-780) once the trend of the
 * scaled sim is known (xsdba_poly_trend_* or xsdba_loess_trend_*): scale, detrend (float64 like the
 * reference), factor lookup as in xsdba_qm_adjust_*, correction, retrend.  scen has the strides of sim.
 */
int xsdba_dqm_adjust_f32(const float* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const float* af_dev, const float* hist_q_dev,
                         const float* scaling_dev, const double* trend_dev, int32_t nq, int32_t interp,
                         int32_t extrap, int32_t kind, float* scen_dev, void* cuda_stream);
int xsdba_dqm_adjust_f64(const double* sim_dev, int64_t n_pts, int64_t stride_pt, int64_t stride_time,
                         const xsdba_grouping_t* grp, const double* af_dev, const double* hist_q_dev,
                         const double* scaling_dev, const double* trend_dev, int32_t nq, int32_t interp,
                         int32_t extrap, int32_t kind, double* scen_dev, void* cuda_stream);

/* Microbenchmark only (profiles/microbench_rows.py): copy every group's member rows with the tiling
 * of the adjust kernel, v in {1,2,4} floats per lane.  Not part of the reference-facing surface. */
int xsdba_debug_copy_rows_f32(const float* src_dev, int64_t n_pts, int64_t stride_time,
                              const xsdba_grouping_t* grp, float* dst_dev, int32_t v, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* XSDBA_B200_H */
