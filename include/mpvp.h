/*
 * mpvp.h -- C ABI of libmpvp.so: the B200 (sm_100a) implementation of the mpv-prescalers
 * per-pixel hot path (RAVU family + NNEDI3).
 *
 * The reference has no FFI: its "interface" is the mpv user-shader file, consumed by a host
 * player that (a) uploads every //!TEXTURE block once at load time and (b) runs each //!DESC pass
 * per frame at its hook point (reference README.md:33-37).  This header is
 * that contract restated for a C caller:
 *
 *   load time   mpvp_weights_create_lut()      <- //!TEXTURE name / SIZE / FORMAT rgba16f + hex payload
 *                                                 (ravu-lite-ar-r3.hook:198-202)
 *               mpvp_weights_create_nnedi3()   <- inline W(i,..)/WS(..) literals
 *                                                 (nnedi3-nns16-win8x4.hook:31-49)
 *   per batch   mpvp_ravu_lite_launch()        <- passes "RAVU-Lite(-AR) (step1|step2, rN)"
 *                                                 (ravu-lite-ar-r3.hook:15-197), fused
 *               mpvp_ravu_launch()             <- passes "RAVU (step1..4, luma|yuv|rgb, rN)"
 *                                                 (ravu-r2.hook:15-338, ravu-r2-rgb.hook), fused chain
 *               mpvp_ravu3x_launch()           <- pass "RAVU-3x (luma|yuv|rgb, rN)"
 *                                                 (compute/ravu-3x-r2.hook:15-115)
 *               mpvp_ravu_zoom_launch()        <- pass "RAVU-Zoom(-AR) (luma|yuv|rgb, rN)"
 *                                                 (ravu-zoom-r2.hook:15-134, ravu-zoom-ar-r2.hook:15-208)
 *               mpvp_nnedi3_launch()           <- passes "NNEDI3 (double_y|combine_y|double_x|combine_x, ..)"
 *                                                 (nnedi3-nns16-win8x4.hook:15-194)
 *
 * Conventions
 *   - all frame pointers are DEVICE pointers owned by the caller; planes are float32 in [0,1],
 *     layout [n][channels][h][w] with explicit strides given in ELEMENTS;
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) on the device that owns the weights; no hidden synchronisation, no allocation;
 *   - return value 0 = success, negative = error (MPVP_E_*); mpvp_last_error() returns a
 *     thread-local description of the last failure;
 *   - `bucket_out` (nullable, int32) receives the LUT row chosen for every key evaluation -- used by
 *     the parity tests to check the >= 99.99 % bucket-agreement rule.
 *   - the *_host entry points take HOST pointers (pageable or pinned), stage them through device
 *     scratch owned by the library, run the same kernels and copy the result back before returning.
 */
#ifndef MPVP_H
#define MPVP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPVP_ABI_VERSION 1

#define MPVP_OK 0
#define MPVP_E_INVALID (-1)   /* bad argument */
#define MPVP_E_CUDA (-2)      /* CUDA runtime error (see mpvp_last_error) */
#define MPVP_E_UNSUPPORTED (-3)
#define MPVP_E_NOMEM (-4)

typedef struct mpvp_weights mpvp_weights; /* opaque, lives on one device */

/* Key-section constants extracted from the hook text. */
typedef struct mpvp_key_params {
  float gauss[36];          /* sigma=2 Gaussian over the inner g*g gradient points, x-major */
  int32_t n_gauss;          /* g*g: 9, 25 (lite/3x) or 16, 36 (ravu/zoom) */
  float strength_thr[3];    /* lite/zoom: 0.004 0.016 0.05; 3x: 0.005 0.02 */
  int32_t n_strength_thr;   /* 0 => log2 form: clamp(floor(log2(lambda*scale + eps)), 0, n_strength-1) */
  float strength_log2_scale;/* ravu: 2000.0 (ravu-r2.hook:97) */
  int32_t n_strength;       /* 4 (lite/zoom), 9 (ravu), 3 (3x) */
  float coherence_thr[2];   /* 0.25 0.5 */
  /* The same quantisers restated on the eigenvalues so that the kernels need no sqrt / division:
   * strength = #{i : L1 >= l1_thr[i]} with l1_thr[i] = the smallest float32 x for which the shader's
   * strength expression evaluated on lambda = sqrt(x) reaches level i+1 (exact, found by bisection);
   * coherence: mu >= c  <=>  L1 >= L2 * ((1+c)/(1-c))^2 = L2 * coh_ratio. */
  float l1_thr[8];
  int32_t n_l1_thr;
  float coh_ratio[2];
} mpvp_key_params;

/* Key source for 3-channel planes. */
#define MPVP_KEY_LUMA 0 /* 1 channel                                   (ravu-r2.hook)      */
#define MPVP_KEY_YUV 1  /* 3 channels, key from channel 0              (ravu-r2-yuv.hook)  */
#define MPVP_KEY_RGB 2  /* 3 channels, key from BT.709 luma of rgb     (ravu-r2-rgb.hook:21) */

/* ---- plane formats (the wire formats either side of the path) -------------------
 * The reference's shaders never see integers: the host binds 8/10/16-bit video planes as UNORM textures,
 * `HOOKED_tex()` returns HOOKED_mul * texture(HOOKED_raw, pos) (gather/ravu-lite-ar-r3.hook:23), and the
 * pass output goes to the host's FBO format (rgba16f by default) or, for the last pass, to the output surface.
 * The *_io entry points take and produce those formats directly, so integer video never round-trips HBM as
 * float32:  integer input   sample = float(raw) / in_max   (one correctly rounded fp32 division; in_max = 255,
 *                           1023, 65535 ... = UNORM normalisation times HOOKED_mul);
 *           integer output  raw = rint(clamp(v, 0, 1) * out_max)  (round-half-even, the UNORM store rule);
 *           MPVP_FMT_F16    binary16 planes (round-to-nearest-even on store).
 * Strides are always in ELEMENTS of the plane's own type.  A null `io` means float32 in and out. */
#define MPVP_FMT_F32 0
#define MPVP_FMT_F16 1
#define MPVP_FMT_U8 2
#define MPVP_FMT_U16 3
typedef struct mpvp_io {
  int32_t in_format, out_format; /* MPVP_FMT_* */
  float in_max, out_max;         /* integer formats only: 255, 1023, 4095, 65535 ... */
} mpvp_io;

const char* mpvp_last_error(void);
int mpvp_abi_version(void);
/* number of kernels launched by this library since load (all threads); for bench bookkeeping */
uint64_t mpvp_launch_count(void);
/* TEST HOOK: cap the number of persistent CTAs of every later launch (0 = no cap, the default) so that small planes
 * make each CTA walk many tiles; returns the previous cap.  Results never depend on it (bit-identical). */
int mpvp_debug_set_grid_limit(int max_ctas);

/* Fill the derived tables of mpvp_key_params (l1_thr[], n_l1_thr, coh_ratio[]) from the shader constants the caller
 * has set (strength_thr[] / n_strength_thr or strength_log2_scale, n_strength, coherence_thr[]): the launch entry
 * points evaluate the (strength, coherence) quantisers on the eigenvalues through these tables and refuse key params
 * whose tables are missing or inconsistent.  Pure host arithmetic; 0 on success. */
int mpvp_key_params_finalize(mpvp_key_params* key);

/* ---- weights -------------------------------------------------------------------------- */

/* Upload a //!TEXTURE payload: host_rgba32f is [h][w][4] float32 exactly as decoded from the hex line.
 * round_to_fp16 != 0 reproduces rgba16f storage (round-to-nearest-even to binary16). */
int mpvp_weights_create_lut(int device, const float* host_rgba32f, int w, int h, int round_to_fp16,
                            mpvp_weights** out);

/* Upload one NNEDI3 direction.  w1, w2: [nns][8][win_short] float32 in canonical window order
 * (index a = offset -3..4 along the long axis, b = offset -(S/2-1)..S/2 along the short axis);
 * b1, b2: [nns].  The library repacks them into the tcgen05 B-operand layout. */
int mpvp_weights_create_nnedi3(int device, const float* w1, const float* w2, const float* b1,
                               const float* b2, int nns, int win_short, mpvp_weights** out);

int mpvp_weights_destroy(mpvp_weights* w);

/* ---- RAVU-Lite (2x luma) ---------------------------------------------------------------- */
/* in [n][h][w] -> out [n][2h][2w]; phase c of the shader's vec4 goes to (2x + c/2, 2y + c%2). */
int mpvp_ravu_lite_launch(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int ar,
                          float ar_strength, const float* in, float* out, int n, int h, int w,
                          int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                          int64_t out_stride_y, int32_t* bucket_out, void* stream);

/* ---- RAVU (2x, three convolutions + merge, offset -0.5,-0.5) --------------------------------- */
/* in [n][c][h][w] -> out [n][c][2h][2w], c = 1 (MPVP_KEY_LUMA) or 3.  bucket_out: [n][3][h][w]. */
int mpvp_ravu_launch(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                     const float* in, float* out, int n, int h, int w, int64_t in_stride_n,
                     int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n,
                     int64_t out_stride_c, int64_t out_stride_y, int32_t* bucket_out, void* stream);

/* ---- RAVU-3x ---------------------------------------------------------------------------- */
/* in [n][c][h][w] -> out [n][c][3h][3w]. bucket_out: [n][h][w]. */
int mpvp_ravu3x_launch(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                       const float* in, float* out, int n, int h, int w, int64_t in_stride_n,
                       int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n,
                       int64_t out_stride_c, int64_t out_stride_y, int32_t* bucket_out, void* stream);

/* ---- RAVU-Zoom (arbitrary ratio) ------------------------------------------------------------ */
/* in [n][c][h][w] -> out [n][c][out_h][out_w]; lut_ar nullable (non-AR). bucket_out: [n][out_h][out_w]. */
int mpvp_ravu_zoom_launch(const mpvp_weights* lut, const mpvp_weights* lut_ar,
                          const mpvp_key_params* key, int radius, int key_mode, float ar_strength,
                          const float* in, float* out, int n, int h, int w, int out_h, int out_w,
                          int64_t in_stride_n, int64_t in_stride_c, int64_t in_stride_y,
                          int64_t out_stride_n, int64_t out_stride_c, int64_t out_stride_y,
                          int32_t* bucket_out, void* stream);

/* ---- NNEDI3 --------------------------------------------------------------------------------- */
/* One doubling pass.  direction 0 = double_y + combine_y: in [n][h][w] -> out [n][2h][w];
 * direction 1 = double_x + combine_x: in [n][h][w] -> out [n][h][2w].  `nn` must have been created
 * from the weights of that direction. */
int mpvp_nnedi3_launch(const mpvp_weights* nn, int direction, const float* in, float* out, int n,
                       int h, int w, int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                       int64_t out_stride_y, void* stream);

/* ---- the same launches with explicit plane formats (see mpvp_io above) ------------------------------------------
 * `in` / `out` point to planes of io->in_format / io->out_format; everything else is as in the float32 entry
 * points, which are exactly these with io = NULL. */
int mpvp_ravu_lite_launch_io(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int ar,
                             float ar_strength, const void* in, void* out, int n, int h, int w,
                             int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                             int64_t out_stride_y, int32_t* bucket_out, const mpvp_io* io, void* stream);
int mpvp_ravu_launch_io(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                        const void* in, void* out, int n, int h, int w, int64_t in_stride_n,
                        int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n,
                        int64_t out_stride_c, int64_t out_stride_y, int32_t* bucket_out, const mpvp_io* io,
                        void* stream);
int mpvp_ravu3x_launch_io(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                          const void* in, void* out, int n, int h, int w, int64_t in_stride_n,
                          int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n,
                          int64_t out_stride_c, int64_t out_stride_y, int32_t* bucket_out, const mpvp_io* io,
                          void* stream);
int mpvp_ravu_zoom_launch_io(const mpvp_weights* lut, const mpvp_weights* lut_ar,
                             const mpvp_key_params* key, int radius, int key_mode, float ar_strength,
                             const void* in, void* out, int n, int h, int w, int out_h, int out_w,
                             int64_t in_stride_n, int64_t in_stride_c, int64_t in_stride_y,
                             int64_t out_stride_n, int64_t out_stride_c, int64_t out_stride_y,
                             int32_t* bucket_out, const mpvp_io* io, void* stream);
int mpvp_nnedi3_launch_io(const mpvp_weights* nn, int direction, const void* in, void* out, int n, int h,
                          int w, int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                          int64_t out_stride_y, const mpvp_io* io, void* stream);

/* ---- the step after the path: offset-correcting main scaler ---------------------------------
 * ravu and nnedi3 declare //!OFFSET -0.5 -0.5 (ravu-r2.hook:325, nnedi3-nns16-win8x4.hook:95,185): texel X of their output
 * holds the content that belongs half a texel further right / down.  mpv corrects that in its main scaler (--scale),
 * which samples the hooked plane at the shifted position while resizing to the output size.  This entry point is that
 * scaler: separable polyphase resampling of `planes` planes [h][w] -> [out_h][out_w] with
 *     s(o) = (o + 0.5) * in / out - 0.5 + offset      (offset = the accumulated //!OFFSET, in texels of the input plane)
 * clamp-to-edge, weights normalised per output coordinate, kernel not widened when downscaling (mpv's default
 * --correct-downscaling=no; downscaling by more than 2x is refused).  Kernels as in mpv's filter_kernels.c. */
#define MPVP_SCALER_BILINEAR 0
#define MPVP_SCALER_CATMULL_ROM 1 /* bicubic B=0, C=0.5 */
#define MPVP_SCALER_MITCHELL 2    /* bicubic B=C=1/3 */
#define MPVP_SCALER_SPLINE36 3
#define MPVP_SCALER_LANCZOS 4     /* sinc windowed by sinc, radius 3 */
int mpvp_resample_launch_io(int device, int kernel, const void* in, void* out, int planes, int h, int w, int out_h,
                            int out_w, float offset_x, float offset_y, int64_t in_stride_p, int64_t in_stride_y,
                            int64_t out_stride_p, int64_t out_stride_y, const mpvp_io* io, void* stream);

/* ---- host-buffer entry points (the end-to-end path: H2D + kernel(s) + D2H inside the call) ------------------------------
 * HOST pointers (pageable or pinned) to contiguous planes [n][c][h][w] in, [n][c][out_h][out_w] out; `io` as in the *_io
 * launches (NULL = float32).  The call stages the frames through device scratch it allocates, in chunks on two streams,
 * and returns after the last copy has completed.  c = 1 for MPVP_KEY_LUMA, else 3. */
int mpvp_ravu_lite_host(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int ar,
                        float ar_strength, const float* host_in, float* host_out, int n, int h,
                        int w);
int mpvp_ravu_host(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                   const void* host_in, void* host_out, int n, int h, int w, const mpvp_io* io);
int mpvp_ravu3x_host(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                     const void* host_in, void* host_out, int n, int h, int w, const mpvp_io* io);
int mpvp_ravu_zoom_host(const mpvp_weights* lut, const mpvp_weights* lut_ar, const mpvp_key_params* key,
                        int radius, int key_mode, float ar_strength, const void* host_in, void* host_out,
                        int n, int h, int w, int out_h, int out_w, const mpvp_io* io);
/* nn_y / nn_x: weights of double_y / double_x; either may be NULL (per-axis //!WHEN): [n][h][w] -> [n][2h or h][2w or w] */
int mpvp_nnedi3_host(const mpvp_weights* nn_y, const mpvp_weights* nn_x, const void* host_in, void* host_out,
                     int n, int h, int w, const mpvp_io* io);

#ifdef __cplusplus
}
#endif
#endif /* MPVP_H */
