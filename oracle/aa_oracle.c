/*
 * aa_oracle.c -- CPU restatement of the reference's anti-aliased separable resize.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under interpolate_antialiasing_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it,
 * and only as the checker.  The product path has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement
 *   (1) bit-for-bit against the unmodified reference extension compiled into oracle/_ref
 *       (tables via the identity trick, forward outputs, the non-AA backward),
 *   (2) against the reference's own known-answer vectors (notebook 64->10 table,
 *       data/proto_aa_interp_lin_step_two_output.png) committed under tests/golden/.
 *
 * All citations are file:line under /root/reference/step_two_dot_two/ unless noted.
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -mfma, no -ffast-math: the
 * reference is x86-64 "-O3" without FMA, so every multiply and add rounds separately).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AA_BOX 0      /* "nearest_forward" = box filter, HelperInterpNearest  aa_interpolation_impl.h:331-373 */
#define AA_TRIANGLE 1 /* "linear_forward", HelperInterpLinear                aa_interpolation_impl.h:285-329 */
#define AA_CUBIC 2    /* "cubic_forward", HelperInterpCubic                  aa_interpolation_impl.h:375-425 */

static int base_interp_size(int filter) { /* :287, :333, :377 */
  return filter == AA_BOX ? 1 : (filter == AA_TRIANGLE ? 2 : 4);
}

/* ------------------------------------------------------------------ fp32 instantiation ---- */

/* Optional caller-provided scale factors: the `scale_factors` argument of ti_upsample_*2d_cpu
 * (:735, :740-742: get_scale_value), consumed by compute_scales_value in torch ATen/native/UpSample.h:
 * (scale.has_value() && scale.value() > 0.) ? static_cast<T>(1.0 / scale.value()) : T(in) / out.
 * The reference's shim always passes {} (extension_interpolate.cpp:12); the oracle keeps the knob so the
 * product's scale_factors plumbing has something to be checked against.  Not thread-safe (test code). */
static double g_sf_h = 0.0, g_sf_w = 0.0; /* per-call factors for forward/backward */
static double g_axis_sf = 0.0;            /* factor of the axis whose tables are being built */
void aa_oracle_set_scale_factors(double sh, double sw) { g_sf_h = sh; g_sf_w = sw; }
void aa_oracle_set_axis_scale(double s) { g_axis_sf = s; }

/* torch ATen/native/UpSample.h area_pixel_compute_scale<float> (called at :314, :347, :391) */
static float scale_f32(int64_t in, int64_t out, int align) {
  if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
  if (g_axis_sf > 0.) return (float)(1.0 / g_axis_sf);
  return (float)in / (float)out;
}

/* :292-300  triangle; `1.0 - x` is evaluated in double and returned as float */
static float filt_triangle_f32(float x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return (float)(1.0 - (double)x);
  return 0.0f;
}
/* :367-372  box on (-0.5, 0.5] */
static float filt_box_f32(float x) {
  if (x > -0.5 && x <= 0.5) return 1.0f;
  return 0.0f;
}
/* :410-424  Keys cubic a=-0.5; first lobe in double, second lobe polynomial in float then *a in double */
static float filt_cubic_f32(float x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) {
    double xd = (double)x;
    return (float)(((a + 2.0) * xd - (a + 3.0)) * xd * xd + 1);
  }
  if (x < 2.0) {
    float p = x - 5;
    p = p * x;
    p = p + 8;
    p = p * x;
    p = p - 4;
    return (float)((double)p * a);
  }
  return 0.0f;
}
static float filt_f32(int filter, float x) {
  return filter == AA_BOX ? filt_box_f32(x) : (filter == AA_TRIANGLE ? filt_triangle_f32(x) : filt_cubic_f32(x));
}

/* K = padded taps per output (:208-210) */
int aa_oracle_interp_size_f32(int64_t in, int64_t out, int filter, int align) {
  float scale = scale_f32(in, out, align);
  int isz = base_interp_size(filter);
  float support = (scale >= 1.0) ? (float)((isz * 0.5) * (double)scale) : (float)(isz * 0.5);
  return (int)ceilf(support) * 2 + 1;
}

/* HelperInterpBase::_compute_indices_weights_aa  :194-281, scalar_t = float.
 * xmin/xsize: [out] int64, w: [out*K] float (zero padded).  Returns K. */
int aa_oracle_tables_f32(int64_t in, int64_t out, int filter, int align,
                         int64_t* xmin_o, int64_t* xsize_o, float* w) {
  float scale = scale_f32(in, out, align);
  int isz = base_interp_size(filter);
  float support = (scale >= 1.0) ? (float)((isz * 0.5) * (double)scale) : (float)(isz * 0.5); /* :208-209 */
  int K = (int)ceilf(support) * 2 + 1;                                                    /* :210 */
  float invscale = (scale >= 1.0) ? (float)(1.0 / (double)scale) : 1.0f;                      /* :242 */
  for (int64_t i = 0; i < out; i++) {
    float center = (float)((double)scale * ((double)i + 0.5));                                /* :253 */
    int64_t xmin = (int64_t)((double)(center - support) + 0.5);                               /* :254 */
    if (xmin < 0) xmin = 0;
    int64_t xmax = (int64_t)((double)(center + support) + 0.5);                               /* :255-257 */
    if (xmax > in) xmax = in;
    xmax -= xmin;
    xmin_o[i] = xmin;
    xsize_o[i] = xmax;
    float total = 0.0f;
    int64_t j;
    for (j = 0; j < xmax; j++) {
      /* (j + xmin - center + 0.5) * invscale : int64 -> float, float subtract, rest in double  :266 */
      float d = (float)(j + xmin) - center;
      float arg = (float)(((double)d + 0.5) * (double)invscale);
      float wj = filt_f32(filter, arg);
      w[i * K + j] = wj;
      total += wj;                                                                             /* :268 */
    }
    for (j = 0; j < xmax; j++)
      if (total != 0.0) w[i * K + j] /= total;                                                 /* :270-274 */
    for (; j < K; j++) w[i * K + j] = 0.0f;                                                    /* :276-278 */
  }
  return K;
}

/* ------------------------------------------------------------------ fp64 instantiation ---- */

static double scale_f64(int64_t in, int64_t out, int align) {
  if (align) return out > 1 ? (double)(in - 1) / (double)(out - 1) : 0.0;
  if (g_axis_sf > 0.) return 1.0 / g_axis_sf;
  return (double)in / (double)out;
}
static double filt_f64(int filter, double x) {
  const double a = -0.5;
  if (filter == AA_BOX) return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0;
  if (x < 0.0) x = -x;
  if (filter == AA_TRIANGLE) return x < 1.0 ? 1.0 - x : 0.0;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}
int aa_oracle_interp_size_f64(int64_t in, int64_t out, int filter, int align) {
  double scale = scale_f64(in, out, align);
  int isz = base_interp_size(filter);
  double support = (scale >= 1.0) ? (isz * 0.5) * scale : isz * 0.5;
  return (int)ceilf((float)support) * 2 + 1; /* ceilf(double) converts to float first (:210) */
}
int aa_oracle_tables_f64(int64_t in, int64_t out, int filter, int align,
                         int64_t* xmin_o, int64_t* xsize_o, double* w) {
  double scale = scale_f64(in, out, align);
  int isz = base_interp_size(filter);
  double support = (scale >= 1.0) ? (isz * 0.5) * scale : isz * 0.5;
  int K = (int)ceilf((float)support) * 2 + 1;
  double invscale = (scale >= 1.0) ? 1.0 / scale : 1.0;
  for (int64_t i = 0; i < out; i++) {
    double center = scale * (i + 0.5);
    int64_t xmin = (int64_t)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int64_t xmax = (int64_t)(center + support + 0.5);
    if (xmax > in) xmax = in;
    xmax -= xmin;
    xmin_o[i] = xmin;
    xsize_o[i] = xmax;
    double total = 0.0;
    int64_t j;
    for (j = 0; j < xmax; j++) {
      double wj = filt_f64(filter, (j + xmin - center + 0.5) * invscale);
      w[i * K + j] = wj;
      total += wj;
    }
    for (j = 0; j < xmax; j++)
      if (total != 0.0) w[i * K + j] /= total;
    for (; j < K; j++) w[i * K + j] = 0.0;
  }
  return K;
}

/* ------------------------------------------------------------------ forward ---------------- */
/*
 * ti_separable_upsample_generic_Nd_kernel_impl :628-683: pass 1 along W (dim 3) into a
 * default-contiguous temp [N,C,H,oW] (:657-668), pass 2 along H (dim 2) into the output (:677-679).
 * Each output element is `t0*w0` then `+= tj*wj`, ascending j, in scalar_t
 * (interpolate_aa_single_dim :60-87, ..._zero_strides :29-58).
 * Input/output are addressed with element strides so channels_first and channels_last
 * tensors are read/written in place (the reference does the same through TensorIterator).
 */
#define DEFINE_FORWARD(NAME, T, TABLES, KFN)                                                            \
  int NAME(const T* in, int64_t N, int64_t C, int64_t H, int64_t W, int64_t isn, int64_t isc,        \
           int64_t ish, int64_t isw, T* out, int64_t oH, int64_t oW, int64_t osn, int64_t osc,       \
           int64_t osh, int64_t osw, int filter, int align) {                                        \
    if (N == 0) return 0;                                                                            \
    g_axis_sf = g_sf_w;                                                                              \
    int64_t kw_ = KFN(W, oW, filter, align);                                                         \
    g_axis_sf = g_sf_h;                                                                              \
    int64_t kh_ = KFN(H, oH, filter, align);                                                         \
    g_axis_sf = 0.0;                                                                                 \
    int64_t* xminw = malloc(sizeof(int64_t) * oW);                                                   \
    int64_t* xsizew = malloc(sizeof(int64_t) * oW);                                                  \
    int64_t* xminh = malloc(sizeof(int64_t) * oH);                                                   \
    int64_t* xsizeh = malloc(sizeof(int64_t) * oH);                                                  \
    T* ww = malloc(sizeof(T) * oW * kw_);                                                             \
    T* wh = malloc(sizeof(T) * oH * kh_);                                                             \
    T* tmp = malloc(sizeof(T) * (size_t)H * oW);                                                     \
    if (!xminw || !xsizew || !xminh || !xsizeh || !ww || !wh || !tmp) return -1;                     \
    g_axis_sf = g_sf_w;                                                                              \
    int Kw = TABLES(W, oW, filter, align, xminw, xsizew, ww);                                        \
    g_axis_sf = g_sf_h;                                                                              \
    int Kh = TABLES(H, oH, filter, align, xminh, xsizeh, wh);                                        \
    g_axis_sf = 0.0;                                                                                 \
    for (int64_t n = 0; n < N; n++)                                                                  \
      for (int64_t c = 0; c < C; c++) {                                                              \
        const T* ip = in + n * isn + c * isc;                                                        \
        T* op = out + n * osn + c * osc;                                                             \
        for (int64_t y = 0; y < H; y++)                                                              \
          for (int64_t ox = 0; ox < oW; ox++) {                                                      \
            const T* s = ip + y * ish + xminw[ox] * isw;                                             \
            const T* wp = ww + ox * Kw;                                                              \
            T acc = s[0] * wp[0];                                                                    \
            for (int64_t j = 1; j < xsizew[ox]; j++) acc += s[j * isw] * wp[j];                      \
            tmp[y * oW + ox] = acc;                                                                  \
          }                                                                                          \
        for (int64_t oy = 0; oy < oH; oy++)                                                          \
          for (int64_t ox = 0; ox < oW; ox++) {                                                      \
            const T* s = tmp + xminh[oy] * oW + ox;                                                  \
            const T* wp = wh + oy * Kh;                                                              \
            T acc = s[0] * wp[0];                                                                    \
            for (int64_t j = 1; j < xsizeh[oy]; j++) acc += s[j * oW] * wp[j];                       \
            op[oy * osh + ox * osw] = acc;                                                           \
          }                                                                                          \
      }                                                                                              \
    free(xminw); free(xsizew); free(xminh); free(xsizeh); free(ww); free(wh); free(tmp);             \
    return 0;                                                                                        \
  }

DEFINE_FORWARD(aa_oracle_forward_f32, float, aa_oracle_tables_f32, aa_oracle_interp_size_f32)
DEFINE_FORWARD(aa_oracle_forward_f64, double, aa_oracle_tables_f64, aa_oracle_interp_size_f64)

/* ------------------------------------------------------------------ backward --------------- */
/*
 * (1) The reference's exported backward: cpu_upsample_linear_backward loop2d
 *     aa_interpolation_backward_impl.h:80-108 -- the stock NON-antialiased 2-tap bilinear scatter
 *     (the `antialias` flag is dropped at :176-180).  Index/lambda arithmetic follows torch
 *     ATen/native/UpSample.h compute_source_index_and_lambda / area_pixel_compute_source_index /
 *     guard_index_and_lambda as installed (2.11).  It is the adjoint of the AA forward only when
 *     both scales are <= 1, and is kept as a regression target for exactly that case.
 *     grad_out/grad_in are contiguous NCHW planes here (the reference .contiguous()-copies, :38-39).
 */
#define DEFINE_NONAA_BACKWARD(NAME, T)                                                               \
  static void NAME##_idx(T ratio, int64_t o, int64_t in, int64_t out, int align, int64_t* i0,       \
                         int64_t* i1, T* l0, T* l1) {                                                \
    if (out == in) { *i0 = o; *i1 = o; *l0 = (T)1; *l1 = (T)0; return; }                             \
    T real;                                                                                          \
    if (align) real = ratio * o;                                                                     \
    else {                                                                                           \
      real = ratio * (o + (T)0.5) - (T)0.5;                                                          \
      if (real < (T)0) real = (T)0;                                                                  \
    }                                                                                                \
    int64_t idx = (int64_t)floorf((float)real); /* guard_index_and_lambda uses floorf even for double */ \
    if (idx > in - 1) idx = in - 1;                                                                  \
    T lam = real - (T)idx;                                                                           \
    if (lam < (T)0) lam = (T)0;                                                                      \
    if (lam > (T)1) lam = (T)1;                                                                      \
    *i0 = idx; *i1 = idx + ((idx < in - 1) ? 1 : 0); *l1 = lam; *l0 = (T)1 - lam;                    \
  }                                                                                                  \
  int NAME(const T* gout, int64_t planes, int64_t oH, int64_t oW, T* gin, int64_t H, int64_t W,      \
           int align) {                                                                              \
    T hs = align ? (oH > 1 ? (T)(H - 1) / (T)(oH - 1) : (T)0) : (T)H / (T)oH;                         \
    T ws = align ? (oW > 1 ? (T)(W - 1) / (T)(oW - 1) : (T)0) : (T)W / (T)oW;                         \
    memset(gin, 0, sizeof(T) * (size_t)planes * H * W);                                              \
    for (int64_t c = 0; c < planes; c++)                                                             \
      for (int64_t oh = 0; oh < oH; oh++) {                                                          \
        int64_t ih0, ih1; T h0, h1;                                                                  \
        NAME##_idx(hs, oh, H, oH, align, &ih0, &ih1, &h0, &h1);                                      \
        for (int64_t ow = 0; ow < oW; ow++) {                                                        \
          int64_t iw0, iw1; T w0, w1;                                                                \
          NAME##_idx(ws, ow, W, oW, align, &iw0, &iw1, &w0, &w1);                                    \
          T g = gout[(c * oH + oh) * oW + ow];                                                       \
          T* p = gin + c * H * W;                                                                    \
          p[ih0 * W + iw0] += h0 * w0 * g;                                                           \
          p[ih0 * W + iw1] += h0 * w1 * g;                                                           \
          p[ih1 * W + iw0] += h1 * w0 * g;                                                           \
          p[ih1 * W + iw1] += h1 * w1 * g;                                                           \
        }                                                                                            \
      }                                                                                              \
    return 0;                                                                                        \
  }

DEFINE_NONAA_BACKWARD(aa_oracle_backward_nonaa_f32, float)
DEFINE_NONAA_BACKWARD(aa_oracle_backward_nonaa_f64, double)

/*
 * (2) The TRUE adjoint of the forward above (what north-star item (3) and cfg4's gradcheck need;
 *     SURVEY section 0.2).  Not present in the reference: "parity unpinned" against reference code,
 *     pinned instead by (a) fp64 gradcheck of the forward, (b) equality with
 *     Wh^T * g * Ww built from the bit-exact tables, which is literally what this computes:
 *     scatter in the reverse pass order (H^T first, then W^T), accumulating in T.
 */
#define DEFINE_ADJOINT(NAME, T, TABLES, KFN)                                                            \
  int NAME(const T* gout, int64_t planes, int64_t oH, int64_t oW, T* gin, int64_t H, int64_t W,      \
           int filter, int align) {                                                                  \
    g_axis_sf = g_sf_w;                                                                              \
    int64_t kw_ = KFN(W, oW, filter, align);                                                         \
    g_axis_sf = g_sf_h;                                                                              \
    int64_t kh_ = KFN(H, oH, filter, align);                                                         \
    g_axis_sf = 0.0;                                                                                 \
    int64_t* xminw = malloc(sizeof(int64_t) * oW);                                                   \
    int64_t* xsizew = malloc(sizeof(int64_t) * oW);                                                  \
    int64_t* xminh = malloc(sizeof(int64_t) * oH);                                                   \
    int64_t* xsizeh = malloc(sizeof(int64_t) * oH);                                                  \
    T* ww = malloc(sizeof(T) * oW * kw_);                                                             \
    T* wh = malloc(sizeof(T) * oH * kh_);                                                             \
    T* tmp = malloc(sizeof(T) * (size_t)H * oW);                                                     \
    if (!xminw || !xsizew || !xminh || !xsizeh || !ww || !wh || !tmp) return -1;                     \
    g_axis_sf = g_sf_w;                                                                              \
    int Kw = TABLES(W, oW, filter, align, xminw, xsizew, ww);                                        \
    g_axis_sf = g_sf_h;                                                                              \
    int Kh = TABLES(H, oH, filter, align, xminh, xsizeh, wh);                                        \
    g_axis_sf = 0.0;                                                                                 \
    for (int64_t c = 0; c < planes; c++) {                                                           \
      const T* g = gout + c * oH * oW;                                                               \
      T* gi = gin + c * H * W;                                                                       \
      memset(tmp, 0, sizeof(T) * (size_t)H * oW);                                                    \
      memset(gi, 0, sizeof(T) * (size_t)H * W);                                                      \
      for (int64_t oy = 0; oy < oH; oy++)                                                            \
        for (int64_t j = 0; j < xsizeh[oy]; j++)                                                     \
          for (int64_t ox = 0; ox < oW; ox++)                                                        \
            tmp[(xminh[oy] + j) * oW + ox] += wh[oy * Kh + j] * g[oy * oW + ox];                     \
      for (int64_t y = 0; y < H; y++)                                                                \
        for (int64_t ox = 0; ox < oW; ox++)                                                          \
          for (int64_t j = 0; j < xsizew[ox]; j++)                                                   \
            gi[y * W + xminw[ox] + j] += ww[ox * Kw + j] * tmp[y * oW + ox];                         \
    }                                                                                                \
    free(xminw); free(xsizew); free(xminh); free(xsizeh); free(ww); free(wh); free(tmp);             \
    return 0;                                                                                        \
  }

DEFINE_ADJOINT(aa_oracle_backward_adjoint_f32, float, aa_oracle_tables_f32, aa_oracle_interp_size_f32)
DEFINE_ADJOINT(aa_oracle_backward_adjoint_f64, double, aa_oracle_tables_f64, aa_oracle_interp_size_f64)
