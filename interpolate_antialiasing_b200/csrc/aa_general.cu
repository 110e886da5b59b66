// aa_general.cu -- gather-form "banded separable apply" tile kernel (sm_100a).
//
//   out[p, i, (jx, c)] = sum_a sum_b  Ah[i, a] * Aw[jx, b] * in[p, a, (b, c)]
//
// where Ah / Aw are banded matrices stored as (start, size, weights[pitch]) per output index.
// With the FORWARD tables (xmin, xsize, w) this is the reference's separable forward
// (/root/reference/step_two_dot_two/aa_interpolation_impl.h:628-683: W pass, then H pass, each
// element `t0*w0` then `+= tj*wj` ascending j, :60-87 / :29-58); with the ADJOINT tables
// (omin, osize, wT) it is the true adjoint Wh^T g Ww in gather form (no atomics, no zero fill),
// replacing aa_interpolation_backward_impl.h:185-219.
//
// EXACT=true keeps multiply and add separate (__fmul_rn/__fadd_rn): bit-identical to the
// reference's x86-64 "-O3" arithmetic for f32 and f64.  EXACT=false uses FMA (backward).
//
// One CTA produces a TH x TW tile of one plane.  The rows of the input needed by the tile are
// walked in chunks of RC rows: each chunk is first filtered horizontally into shared memory
// (T[RC][TW], rounded to acc_t exactly like the reference's temp tensor), then every thread
// advances the vertical sums of its own output elements over that chunk, in ascending row order.
// This path is the general one (any scale, up or down, f64, box filter, backward); the streaming
// kernel in aa_stream.cu is the bandwidth-optimised forward for downsampling.
#include "aa_common.cuh"

namespace aa {
namespace {

constexpr int TW = 64;   // flat output columns per tile
constexpr int TH = 16;   // output rows per tile
constexpr int NTY = 4;   // thread rows
constexpr int RPT = TH / NTY;  // output rows per thread
constexpr int RC = 32;   // input rows per chunk
constexpr int HPT = RC / NTY;  // horizontal-pass rows per thread per chunk

template <typename T> struct LoadCvt;
template <> struct LoadCvt<float> {
  template <typename A> static __device__ __forceinline__ A ld(const float* p) { return (A)__ldg(p); }
};
template <> struct LoadCvt<double> {
  template <typename A> static __device__ __forceinline__ A ld(const double* p) { return (A)__ldg(p); }
};
template <> struct LoadCvt<uint8_t> {
  template <typename A> static __device__ __forceinline__ A ld(const uint8_t* p) { return (A)__ldg(p); }
};

template <bool EXACT> __device__ __forceinline__ float mac(float acc, float a, float b) {
  return EXACT ? __fadd_rn(acc, __fmul_rn(a, b)) : fmaf(a, b, acc);
}
template <bool EXACT> __device__ __forceinline__ double mac(double acc, double a, double b) {
  return EXACT ? __dadd_rn(acc, __dmul_rn(a, b)) : fma(a, b, acc);
}

struct GParams {
  const void* in;
  void* out;
  OutEpi epi;
  Layout lin, lout;
  const int32_t *h_start, *h_size, *w_start, *w_size;
  const void *h_w, *w_w;
  int h_pitch, w_pitch;
  int64_t in_h, in_w, out_h, out_w;  // spatial sizes (w in pixels, not flat)
  int64_t tiles_x, tiles_y;
};

template <typename in_t, typename acc_t, bool EXACT>
__global__ void __launch_bounds__(TW* NTY) aa_general_kernel(const GParams P) {
  __shared__ acc_t Ts[RC][TW + 1];
  const int tx = threadIdx.x % TW;
  const int ty = threadIdx.x / TW;
  int64_t b = blockIdx.x;
  const int64_t tile_x = b % P.tiles_x; b /= P.tiles_x;
  const int64_t tile_y = b % P.tiles_y; b /= P.tiles_y;
  const int64_t plane = b;
  const int Ci = P.lin.Ci;
  const int64_t owf = P.out_w * Ci;  // flat output width
  const in_t* ip = (const in_t*)P.in + (plane / P.lin.Cp) * P.lin.stride_n + (plane % P.lin.Cp) * P.lin.stride_p;
  acc_t* op = (acc_t*)P.out + (plane / P.lout.Cp) * P.lout.stride_n + (plane % P.lout.Cp) * P.lout.stride_p;
  const acc_t* hw = (const acc_t*)P.h_w;
  const acc_t* ww = (const acc_t*)P.w_w;

  // this thread's flat output column
  const int64_t of = tile_x * TW + tx;
  const bool col_ok = of < owf;
  int64_t jx = 0; int c = 0; int wst = 0, wsz = 0;
  if (col_ok) {
    jx = of / Ci; c = (int)(of % Ci);
    wst = P.w_start[jx]; wsz = P.w_size[jx];
  }
  const acc_t* wrow = ww + jx * P.w_pitch;
  const in_t* icol = ip + (int64_t)wst * Ci + c;

  // this thread's output rows and their vertical windows
  const int64_t oy0 = tile_y * TH;
  int hst[RPT], hsz[RPT];
  acc_t acc[RPT];
#pragma unroll
  for (int r = 0; r < RPT; r++) {
    const int64_t oy = oy0 + ty + r * NTY;
    acc[r] = (acc_t)0;
    if (oy < P.out_h) { hst[r] = P.h_start[oy]; hsz[r] = P.h_size[oy]; }
    else { hst[r] = 0; hsz[r] = 0; }
  }
  // input rows needed by the tile (starts and ends are non-decreasing in the output index)
  const int64_t oy_last = (oy0 + TH <= P.out_h ? oy0 + TH : P.out_h) - 1;
  const int64_t row_begin = P.h_start[oy0];
  const int64_t row_end = (int64_t)P.h_start[oy_last] + P.h_size[oy_last];

  for (int64_t r0 = row_begin; r0 < row_end; r0 += RC) {
    // ---- horizontal pass of rows [r0, r0+RC) for this tile's columns -> Ts
    acc_t t[HPT];
#pragma unroll
    for (int i = 0; i < HPT; i++) t[i] = (acc_t)0;
    if (col_ok) {
      for (int k = 0; k < wsz; k++) {
        const acc_t wk = __ldg(wrow + k);
        const in_t* src = icol + (int64_t)k * Ci;
#pragma unroll
        for (int i = 0; i < HPT; i++) {
          const int64_t row = r0 + ty + i * NTY;
          if (row < row_end) t[i] = mac<EXACT>(t[i], LoadCvt<in_t>::template ld<acc_t>(src + row * P.lin.stride_h), wk);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < HPT; i++) Ts[ty + i * NTY][tx] = t[i];
    __syncthreads();
    // ---- vertical accumulation over this chunk, ascending rows
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      int64_t a = hst[r] > r0 ? hst[r] : r0;
      int64_t e = (int64_t)hst[r] + hsz[r];
      if (e > r0 + RC) e = r0 + RC;
      const acc_t* hrow = hw + (oy0 + ty + r * NTY) * P.h_pitch;
      for (int64_t row = a; row < e; row++)
        acc[r] = mac<EXACT>(acc[r], Ts[row - r0][tx], __ldg(hrow + (row - hst[r])));
    }
    __syncthreads();
  }
  if (col_ok) {
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const int64_t oy = oy0 + ty + r * NTY;
      if (oy < P.out_h) {
        if constexpr (sizeof(acc_t) == 4) {
          if (!P.epi.plain()) {
            const int64_t off = (plane / P.lout.Cp) * P.lout.stride_n + (plane % P.lout.Cp) * P.lout.stride_p + oy * P.lout.stride_h +
                                P.epi.coloff((int)jx, c, Ci);
            aa_store<true>(P.out, off, (float)acc[r], c, P.epi);
            continue;
          }
        }
        op[oy * P.lout.stride_h + of] = acc[r];
      }
    }
  }
}

template <typename in_t, typename acc_t>
int launch_t(const GParams& P, int64_t planes, bool exact, cudaStream_t stream) {
  const int64_t nblocks = P.tiles_x * P.tiles_y * planes;
  if (nblocks <= 0) return AA_OK;
  if (nblocks >= (1ll << 31)) return fail(AA_ERR_UNSUPPORTED, "too many tiles for one launch");
  if (exact) aa_general_kernel<in_t, acc_t, true><<<(unsigned)nblocks, TW * NTY, 0, stream>>>(P);
  else aa_general_kernel<in_t, acc_t, false><<<(unsigned)nblocks, TW * NTY, 0, stream>>>(P);
  AA_LAUNCH_CHECK("aa_general_kernel");
  return AA_OK;
}

// ---- the reference's exported (non-AA) bilinear backward, gather form --------------------------
// cpu_upsample_linear_backward loop2d, aa_interpolation_backward_impl.h:80-108: for every output
// (oh, ow) four scatter-adds  gin[ih_a][iw_b] += (h_a * w_b) * g.  Here every grad_input element
// gathers the same terms in the same (oh, ow, a, b) order, so the sums round identically.
template <typename T> struct Lam { int64_t i0, i1; T l0, l1; };

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

template <typename T>
__device__ __forceinline__ Lam<T> src_index_rn(T ratio, int64_t o, int64_t in, int64_t out, int align) {
  Lam<T> r;
  if (out == in) { r.i0 = o; r.i1 = o; r.l0 = (T)1; r.l1 = (T)0; return r; }
  T real;
  if (align) real = mul_rn(ratio, (T)o);
  else {
    real = sub_rn(mul_rn(ratio, add_rn((T)o, (T)0.5)), (T)0.5);
    if (real < (T)0) real = (T)0;
  }
  // guard_index_and_lambda of the INSTALLED torch (2.11, ATen/native/UpSample.h) -- the header oracle/_ref compiles
  // against -- is min(static_cast<int64_t>(floorf(real)), in - 1): floorf, i.e. through float even when T is double
  // (the 2021 header had static_cast<int64_t>(real); for real >= 0 the two differ only where float rounding crosses an
  // integer).  Kept bit-identical to what the oracle computes here (tests/test_backward_gpu.py, f64 non-AA backward).
  int64_t idx = (int64_t)floorf((float)real);
  if (idx > in - 1) idx = in - 1;
  T lam = sub_rn(real, (T)idx);
  lam = lam < (T)0 ? (T)0 : (lam > (T)1 ? (T)1 : lam);
  r.i0 = idx; r.i1 = idx + (idx < in - 1 ? 1 : 0); r.l1 = lam; r.l0 = sub_rn((T)1, lam);
  return r;
}

template <typename T>
__global__ void aa_nonaa_bilinear_backward(const T* __restrict__ gout, T* __restrict__ gin, Layout lo, Layout li,
                                           int64_t oH, int64_t oW, int64_t H, int64_t W, int align, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Ci = li.Ci;
  const int64_t rowf = W * Ci;
  const int64_t f = idx % rowf;
  const int64_t y = (idx / rowf) % H;
  const int64_t p = idx / (rowf * H);
  const int64_t x = f / Ci;
  const int c = (int)(f % Ci);
  const T* g = gout + (p / lo.Cp) * lo.stride_n + (p % lo.Cp) * lo.stride_p;
  T* gi = gin + (p / li.Cp) * li.stride_n + (p % li.Cp) * li.stride_p;
  const T hs = align ? (oH > 1 ? div_rn((T)(H - 1), (T)(oH - 1)) : (T)0) : div_rn((T)H, (T)oH);
  const T ws = align ? (oW > 1 ? div_rn((T)(W - 1), (T)(oW - 1)) : (T)0) : div_rn((T)W, (T)oW);
  // conservative candidate ranges of outputs that can touch (y, x): invert the source map +-2
  auto range = [](double sc, int64_t v, int64_t out, int64_t in, int64_t* lo_, int64_t* hi_) {
    if (sc <= 0.0 || out == in) {
      if (out == in) { *lo_ = v; *hi_ = v + 1; } else { *lo_ = 0; *hi_ = out; }
      return;
    }
    double a = ((double)v - 1.5) / sc - 2.0, b = ((double)v + 1.5) / sc + 2.0;
    int64_t l = (int64_t)floor(a), h = (int64_t)ceil(b) + 1;
    *lo_ = l < 0 ? 0 : (l > out ? out : l);
    *hi_ = h < 0 ? 0 : (h > out ? out : h);
  };
  int64_t oh0, oh1, ow0, ow1;
  range((double)hs, y, oH, H, &oh0, &oh1);
  range((double)ws, x, oW, W, &ow0, &ow1);
  if (y == H - 1 || y == 0) { if (y == 0) oh0 = 0; if (y == H - 1) oh1 = oH; }
  if (x == W - 1 || x == 0) { if (x == 0) ow0 = 0; if (x == W - 1) ow1 = oW; }
  T acc = (T)0;
  for (int64_t oh = oh0; oh < oh1; oh++) {
    const Lam<T> lh = src_index_rn<T>(hs, oh, H, oH, align);
    if (lh.i0 != y && lh.i1 != y) continue;
    for (int64_t ow = ow0; ow < ow1; ow++) {
      const Lam<T> lw = src_index_rn<T>(ws, ow, W, oW, align);
      if (lw.i0 != x && lw.i1 != x) continue;
      const T gv = g[oh * lo.stride_h + ow * Ci + c];
      if (lh.i0 == y && lw.i0 == x) acc = add_rn(acc, mul_rn(mul_rn(lh.l0, lw.l0), gv));
      if (lh.i0 == y && lw.i1 == x) acc = add_rn(acc, mul_rn(mul_rn(lh.l0, lw.l1), gv));
      if (lh.i1 == y && lw.i0 == x) acc = add_rn(acc, mul_rn(mul_rn(lh.l1, lw.l0), gv));
      if (lh.i1 == y && lw.i1 == x) acc = add_rn(acc, mul_rn(mul_rn(lh.l1, lw.l1), gv));
    }
  }
  gi[y * li.stride_h + f] = acc;
}

}  // namespace

int launch_general(const void* in, int in_dtype, const Layout& lin, void* out, int out_dtype,
                   const Layout& lout, const BandedAxis& ah, const BandedAxis& aw, bool exact,
                   OutEpi epi, cudaStream_t stream) {
  GParams P;
  P.in = in; P.out = out; P.epi = epi; P.lin = lin; P.lout = lout;
  if (!epi.plain()) {
    if (in_dtype == AA_F64) return fail(AA_ERR_UNSUPPORTED, "the fused output epilogue needs u8 or f32 input");
    out_dtype = AA_F32;  // accumulate in f32, convert at the store
  }
  P.h_start = ah.start; P.h_size = ah.size; P.h_w = ah.w; P.h_pitch = ah.pitch;
  P.w_start = aw.start; P.w_size = aw.size; P.w_w = aw.w; P.w_pitch = aw.pitch;
  P.in_h = ah.n_in; P.in_w = aw.n_in; P.out_h = ah.n_out; P.out_w = aw.n_out;
  P.tiles_x = (aw.n_out * lin.Ci + TW - 1) / TW;
  P.tiles_y = (ah.n_out + TH - 1) / TH;
  if (out_dtype == AA_F64) {
    if (in_dtype != AA_F64) return fail(AA_ERR_UNSUPPORTED, "f64 output needs f64 input");
    return launch_t<double, double>(P, lin.planes, exact, stream);
  }
  if (out_dtype != AA_F32) return fail(AA_ERR_UNSUPPORTED, "output dtype must be f32 or f64");
  if (in_dtype == AA_F32) return launch_t<float, float>(P, lin.planes, exact, stream);
  if (in_dtype == AA_U8) return launch_t<uint8_t, float>(P, lin.planes, exact, stream);
  return fail(AA_ERR_UNSUPPORTED, "f32 output needs u8 or f32 input");
}

int launch_backward_nonaa(const void* gout, void* gin, int dtype, const Layout& lout, const Layout& lin,
                          int64_t oH, int64_t oW, int64_t H, int64_t W, int align, cudaStream_t stream) {
  const int64_t total = lin.planes * H * W * lin.Ci;
  if (total == 0) return AA_OK;
  const int NT = 256;
  const int64_t nb = (total + NT - 1) / NT;
  if (nb >= (1ll << 31)) return fail(AA_ERR_UNSUPPORTED, "too many elements for one launch");
  if (dtype == AA_F32)
    aa_nonaa_bilinear_backward<float><<<(unsigned)nb, NT, 0, stream>>>((const float*)gout, (float*)gin, lout, lin, oH, oW, H, W, align, total);
  else if (dtype == AA_F64)
    aa_nonaa_bilinear_backward<double><<<(unsigned)nb, NT, 0, stream>>>((const double*)gout, (double*)gin, lout, lin, oH, oW, H, W, align, total);
  else return fail(AA_ERR_UNSUPPORTED, "non-AA backward: f32/f64 only");
  AA_LAUNCH_CHECK("aa_nonaa_bilinear_backward");
  return AA_OK;
}

}  // namespace aa
