/* dfit.h -- C ABI of the B200-native per-voxel curve-fit engine (libdfit.so).
 *
 * Drop-in boundary for the per-voxel non-linear least-squares hot path of ad12/DOSMA.  The
 * reference has no FFI for this path: its "operator API" is the Python surface
 *
 *   dosma/core/fitting.py:755-768   curve_fit(func, x, y, y_bounds, p0, maxfev, ftol, eps, ...)
 *   dosma/core/fitting.py:1026-1073 _curve_fit(...)            one voxel -> scipy.optimize.curve_fit
 *   dosma/core/fitting.py:109-146   _Fitter._process_params    ufuncs, bounds, r2 threshold, nan fill
 *   dosma/core/fitting.py:205-215   mask scatter
 *   dosma/core/fitting.py:701-718   MonoExponentialFit log-linear ("polyfit") initial guess
 *   dosma/core/fitting.py:734-737   MonoExponentialFit rounding
 *
 * and the entry points below are what a ctypes binding inside that module would call instead of
 * the `for i in range(N)` / `mp.Pool.map` loop at fitting.py:855-868 (INTEGRATION.md shows the
 * stub).  Plain C: pointers and sizes only, no C++ or torch types, no exceptions; every function
 * returns DFIT_OK (0) or a negative dfit_status, and dfit_last_error() returns a thread-local
 * detail string for the last failure.
 *
 * There is NO CPU implementation behind these symbols: without a CUDA device every compute entry
 * point fails with DFIT_ERR_NO_DEVICE.
 */
#ifndef DFIT_H_
#define DFIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFIT_VERSION 300 /* major*10000 + minor*100 + patch */
#define DFIT_MAX_PARAMS 4
#define DFIT_MAX_ECHOES 32

typedef enum dfit_status {
  DFIT_OK = 0,
  DFIT_ERR_BAD_ARG = -1,
  DFIT_ERR_NO_DEVICE = -2,
  DFIT_ERR_CUDA = -3,
  DFIT_ERR_OOM = -4,
  DFIT_ERR_UNSUPPORTED = -5
} dfit_status;

/* Model registry (fitting.py:1016-1023 + the 1-parameter model of tests/core/test_fitting.py:52). */
typedef enum dfit_model {
  DFIT_MODEL_MONOEXP = 0, /* a * exp(b x)                       P = 2 */
  DFIT_MODEL_BIEXP = 1,   /* a1 exp(b1 x) + a2 exp(b2 x)        P = 4 */
  DFIT_MODEL_LINEAR = 2   /* a * x                              P = 1 */
} dfit_model;

typedef enum dfit_dtype {
  DFIT_F32 = 0,
  DFIT_F64 = 1,
  DFIT_I16 = 2, /* DICOM pixel data (dicom_io.py:296-309) is converted in-kernel */
  DFIT_U16 = 3,
  DFIT_I32 = 4,
  DFIT_U8 = 5
} dfit_dtype;

/* Sample layout of `y`: element (echo e, voxel v). */
typedef enum dfit_layout {
  DFIT_PLANAR = 0,      /* y[e * ld + v]  -- the (E, N) array of fitting.py:194-196; ld >= N */
  DFIT_ECHO_FASTEST = 1 /* y[v * ld + e]  -- ld >= E */
} dfit_layout;

typedef enum dfit_init {
  DFIT_INIT_GIVEN = 0,    /* p0 (scalar and/or per voxel) as passed               fitting.py:720 */
  DFIT_INIT_LOGLINEAR = 1 /* mono-exp only: degree-1 fit of ln(y), in-kernel      fitting.py:701-718 */
} dfit_init;

/* Post-processing ufuncs recognised by the fused epilogue (fitting.py:123-128). */
typedef enum dfit_ufunc {
  DFIT_UFUNC_NONE = 0,
  DFIT_UFUNC_INV_ABS = 1, /* 1 / |v|   (MonoExponentialFit, fitting.py:725) */
  DFIT_UFUNC_NEG_INV = 2, /* -1 / v */
  DFIT_UFUNC_ABS = 3,
  DFIT_UFUNC_INV = 4
} dfit_ufunc;

/* Per-voxel status byte (optional output).  1..4 are success, like SciPy's ier in (1,2,3,4). */
typedef enum dfit_voxel_status {
  DFIT_VOX_SKIPPED = 0,   /* outside mask, all-zero samples, or y out of y_bounds (fitting.py:1065-1067) */
  DFIT_VOX_CONV_F = 1,
  DFIT_VOX_CONV_X = 2,
  DFIT_VOX_CONV_FX = 3,
  DFIT_VOX_EXACT = 4,
  DFIT_VOX_MAXITER = 5,   /* -> NaN parameters, r2 = 0 (fitting.py:1069-1073) */
  DFIT_VOX_NONFINITE = 6, /* NaN/Inf sample or initial guess */
  DFIT_VOX_NUMERIC = 7
} dfit_voxel_status;

/* Options.  Initialise with dfit_default_opts(); all fields have reference-equivalent defaults. */
typedef struct dfit_opts {
  int32_t struct_size;   /* sizeof(dfit_opts), for ABI evolution */
  int32_t model;         /* dfit_model */
  int32_t compute_dtype; /* DFIT_F32 or DFIT_F64: arithmetic type of the solver */
  int32_t init_mode;     /* dfit_init */
  int32_t init_linear;   /* start the linear parameters at their least-squares optimum: 1 on, 0 off,
                            -1 auto (on for models with one linear parameter) */
  int32_t maxfev;        /* reference budget (fitting.py:761, default 100), counted like MINPACK:
                            1 per trial step + P per accepted step */
  double ftol;           /* reference tolerance (fitting.py:762, default 1e-5) */
  double ftol_scale;     /* engine stops at ftol*ftol_scale (default 1e-2): see DESIGN.md "Parity definition" */
  double xtol;           /* relative scaled-step tolerance; <= 0 selects the dtype default */
  double lambda0;        /* initial Marquardt damping; <= 0 selects the default 1e-3 */
  double r2_eps;         /* fitting.py:763, default 1e-8 */
  double y_lo, y_hi;     /* y_bounds (fitting.py:758, 1065); default -inf, +inf */
  double p0[DFIT_MAX_PARAMS]; /* scalar initial guess; NaN entries are taken from p0_voxel */
  /* ---- fused epilogue: _process_params + mask fill + rounding ---- */
  int32_t post_enabled;
  int32_t ufunc[DFIT_MAX_PARAMS];
  double lb[DFIT_MAX_PARAMS], ub[DFIT_MAX_PARAMS]; /* inclusive; outside -> NaN (fitting.py:130-138) */
  int32_t has_r2_threshold;
  double r2_threshold;   /* r2 < threshold -> NaN parameters (fitting.py:140-141) */
  int32_t has_nan_fill;
  double nan_fill;       /* np.nan_to_num value (fitting.py:143-144); also fills voxels outside the mask */
  int32_t decimals[DFIT_MAX_PARAMS]; /* np.around per parameter, < 0 = none (fitting.py:736-737) */
  int32_t fast_path;     /* mono-exponential model on uniformly spaced echoes: variable-projection Newton on
                            q = exp(b dx) from a data-driven (Prony) start, falling back per voxel to the LM
                            from p0 whenever it declines.  -1 auto (on), 0 off (always LM from p0), 1 on, 2 on but never
                            with the two-voxels-per-lane kernel (diagnostic) */
  int32_t use_tma;         /* -1 auto, 0 plain coalesced loads, 1 TMA-staged tiles */
  int32_t out_param;       /* -1 (default): popt is [N, P].  i >= 0: only parameter i is written and popt is [N] --
                              what MonoExponentialFit keeps of its fit (fitting.py:734): a third less to write and to
                              bring back over PCIe */
} dfit_opts;

/* Aggregate counters of the last fit on a handle (device-accumulated). */
typedef struct dfit_stats {
  int64_t n_voxels;     /* voxels presented */
  int64_t n_fitted;     /* voxels on which the solver ran */
  int64_t n_failed;     /* solver ran but did not converge (status >= 5) */
  int64_t n_nonfinite;  /* voxels with NaN/Inf input -- the reference raises ValueError for these */
  int64_t n_oob;        /* voxels skipped by y_bounds */
  int64_t sum_iters;    /* total passes over the echoes (Newton passes of the fast path, LM trial steps otherwise) */
  int32_t max_iters;    /* largest per-voxel pass count */
  int32_t n_launches;   /* kernels launched by the call */
  float kernel_ms;      /* device time of the fit kernel(s), CUDA events on the engine's stream */
  float total_ms;       /* device time of the whole call incl. copies (host entry point only) */
  int64_t n_deferred;   /* dense two-voxel TMA kernel: voxels the straight-line fast path turned down (they are queued per
                           warp and fitted 32 at a time by the one-voxel path: Newton loop, then LM from p0 -- inside the
                           kernel, or, where most of a batch is LM-bound (air, background), by the LM-in-rounds kernel
                           launched right behind it: n_launches then counts two) */
} dfit_stats;

typedef struct dfit_handle dfit_handle;

int dfit_version(void);
int dfit_device_count(void);
const char* dfit_strerror(int code);
const char* dfit_last_error(void);
int dfit_default_opts(dfit_opts* opts, int model);
int dfit_model_nparams(int model);

/* One handle per (thread, device): owns streams, pinned staging and device scratch. */
int dfit_create(int device, dfit_handle** out);
int dfit_destroy(dfit_handle* h);

/* Fit N voxels whose samples are already in device memory (HBM-resident path).
 *   x        host, double[n_echo]                      echo / spin-lock times
 *   y        device, dtype y_dtype, layout/ld as above
 *   mask     device uint8[N] or NULL; only mask != 0 is fitted (fitting.py:199-200)
 *   p0_voxel device, p0_dtype (F32|F64) [N, P] or NULL (fitting.py:849-851)
 *   popt     device, out_dtype (F32|F64) [N, P];  r2 device, out_dtype [N]
 *   status   device uint8[N] or NULL;  niter device uint8[N] or NULL
 *   stream   cudaStream_t (as void*) to launch on; NULL = the legacy default stream
 * Asynchronous with respect to the host; outputs are valid once `stream` has drained. */
int dfit_fit_device(dfit_handle* h, const dfit_opts* opts, int n_echo, int64_t n_vox, const double* x, const void* y,
                    int y_dtype, int y_layout, int64_t ld, const uint8_t* mask, const void* p0_voxel, int p0_dtype,
                    void* popt, void* r2, int out_dtype, uint8_t* status, uint8_t* niter, void* stream);

/* Fused all-gather epilogue for multi-GPU runs (one process per GPU).  maps[r], r = 0..world-1, is a
 * device pointer THIS process can store to that addresses rank r's reassembled fp32 map of shape
 * [world * rows_per_rank, P + 1] (its own allocation for r == rank, a CUDA-IPC / peer mapping of
 * rank r's allocation otherwise).  While set, every dfit_fit_device call on this handle additionally
 * stores each voxel's packed row [popt..., r2] into all `world` maps at row rank * rows_per_rank + v,
 * so the reassembly of the parameter map travels over NVLink while the fit is
 * still running; popt / r2 may then be NULL.  Peer stores are complete when the launching stream has
 * drained; ranks synchronise with each other before reading their maps.  world = 0 clears. */
int dfit_set_gather(dfit_handle* h, int world, int rank, void* const* maps, int64_t rows_per_rank);

/* The general form.  Rows carry the parameters named in `param_mask` (bit i = parameter i; 0 = all) followed by r2:
 * a T2 / T1rho map needs [b or tc, r2] -- 8-byte rows instead of 12, and the reassembly is bound by the bytes every
 * GPU must receive.  `multicast`: an NVLS multicast address bound to the same `world` buffers (e.g. from
 * torch.distributed._symmetric_memory): every row then leaves the GPU once (multimem.st) and the NVSwitch replicates
 * it, instead of `world - 1` peer stores.
 * `split_list` (masked fits of ONE volume by all ranks, SURVEY.md section 8e: "shard the mask-compacted voxel list"): every
 * rank passes the mask of the WHOLE volume (n_vox = whole volume, the same on all ranks), fills its own map outside
 * the mask and fits the masked voxels of its voxel span [fit_lo, fit_hi) only -- the host cuts the spans so that each
 * holds the same number of masked voxels however the tissue is distributed (dosma_b200.sharding.masked_spans); `y`
 * then holds only the samples of the span, y_voxel0 <= fit_lo being the voxel index of its first column. */
typedef struct dfit_gather_desc {
  int32_t struct_size; /* sizeof(dfit_gather_desc) */
  int32_t world, rank;
  void* const* maps;   /* world device pointers this process can store to */
  void* multicast;     /* or NULL */
  int64_t rows;        /* rows in every map */
  int64_t row0;        /* row of voxel 0 of this rank's dfit_fit_device calls */
  uint32_t param_mask;
  int32_t split_list;
  int64_t fit_lo, fit_hi;
  int64_t y_voxel0;
} dfit_gather_desc;
int dfit_set_gather_ex(dfit_handle* h, const dfit_gather_desc* g); /* g == NULL or g->world == 0 clears */

/* Plumbing for the maps above: allocate a zeroed device buffer and export its CUDA IPC handle
 * (DFIT_IPC_HANDLE_BYTES bytes, to be sent to the peer processes by any means), and map a peer's
 * buffer into this process for stores from THIS handle's device (peer access over NVLink is enabled
 * lazily by the mapping). */
#define DFIT_IPC_HANDLE_BYTES 64
int dfit_ipc_alloc(dfit_handle* h, size_t bytes, void** dev_ptr, unsigned char* handle_out);
int dfit_ipc_open(dfit_handle* h, const unsigned char* handle, void** dev_ptr);
int dfit_ipc_close(dfit_handle* h, void* dev_ptr);
int dfit_ipc_free(dfit_handle* h, void* dev_ptr);

/* Fit N voxels whose samples are in HOST memory: the reference-facing call.
 *   y_planes  host, n_echo pointers, one contiguous [N] plane per echo (the list of volumes the
 *             reference concatenates at fitting.py:194-196), element type y_dtype
 *   mask, p0_voxel, popt, r2, status, niter: host buffers, same shapes as above (NULL where optional)
 * Copies are chunked and pipelined against the kernel on the handle's streams; pinned buffers are
 * used directly, pageable ones are staged.  Synchronous: outputs are complete on return. */
int dfit_fit_host(dfit_handle* h, const dfit_opts* opts, int n_echo, int64_t n_vox, const double* x,
                  const void* const* y_planes, int y_dtype, const uint8_t* mask, const void* p0_voxel, int p0_dtype,
                  void* popt, void* r2, int out_dtype, uint8_t* status, uint8_t* niter);

/* Counters of the most recent dfit_fit_* call on this handle (synchronises the handle's stream). */
int dfit_get_stats(dfit_handle* h, dfit_stats* out);

/* ---- qDESS analytic T2 map --------------------------------------------------
 * Element-wise restatement of dosma/scan_sequences/mri/qdess.py:225-255:
 *   t2 = nan_to_num(-2000 (TR - TE) / (log(|nan_to_num(S2 / S1)| / k) + c1)), then bounds -> NaN (:237-239),
 *   NaN -> fill (:240-245), around(decimals) (:247-248), optional fat / fluid suppression (:250-255).
 * k and c1 are the scalars of qdess.py:204-223, computed by the caller from the sequence parameters. */
typedef struct dfit_qdess_opts {
  int32_t struct_size;
  double k, c1;
  double tr_minus_te; /* seconds */
  int32_t has_bounds;
  double lb, ub;
  int32_t has_nan_fill;
  double nan_fill;
  int32_t decimals; /* < 0: none */
  int32_t suppress_fat, suppress_fluid;
  double beta;
  int32_t compute_dtype; /* -1 / DFIT_F64: the reference's float64 arithmetic (exact parity, default);
                            DFIT_F32: fp32 + MUFU fast path (HBM-bound, ~2e-6 relative before rounding) */
} dfit_qdess_opts;

int dfit_default_qdess_opts(dfit_qdess_opts* opts);
int dfit_qdess_t2_device(dfit_handle* h, const dfit_qdess_opts* opts, int64_t n_vox, const void* echo1, const void* echo2,
                         int in_dtype, void* t2, int out_dtype, void* stream);
int dfit_qdess_t2_host(dfit_handle* h, const dfit_qdess_opts* opts, int64_t n_vox, const void* echo1, const void* echo2,
                       int in_dtype, void* t2, int out_dtype);

/* ---- region statistics of a fitted map -------------------------------------
 * `QuantitativeValue.to_metrics`, dosma/core/quant_vals.py:145-229: for each region, over the valid voxels
 * (finite, and inside `bounds` with the given closedness, :182-190): number of voxels, np.nanmean,
 * np.nanstd (population) and np.nanmedian, in float64.
 *   map            host, f32 or f64 [N];  labels: host label mask [N] (any dfit_dtype) or NULL
 *   region_labels  n_regions entries: L > 0 selects label L (:218), -1 any positive label ("total" with a
 *                  mask, :216), -2 every voxel ("total" without a mask, :214)
 *   out            n_regions x 4 doubles: count, mean, std, median (NaN where the region is empty) */
int dfit_region_metrics_host(dfit_handle* h, int64_t n_vox, const void* map, int map_dtype, const void* labels,
                             int labels_dtype, int n_regions, const int32_t* region_labels, int has_bounds, double lb,
                             double ub, int closed_left, int closed_right, double* out);

#ifdef __cplusplus
}
#endif
#endif /* DFIT_H_ */
