// aa_band.cu -- K3b: few-tap gather kernel for scales around 1 (about 0.6x..1x: as many or more input
// rows than output rows, <= 7 taps; the forward of aa_interpolation_impl.h:60-87 in that range, e.g.
// 1080p -> 720p).  Same arithmetic as aa_tile.cu (horizontal pass into a shared T, then vertical pass,
// FMA, ascending taps) but organised as a walk down the image, because here the input patch of an output
// tile is as large as the tile itself and adjacent tiles share several rows: one CTA owns a column band
// of 256 flat outputs of one plane and walks down it in chunks of TY output rows (a "segment" of
// consecutive chunks per CTA, so the per-CTA setup -- plane decode, column weights -- is paid once and
// every input row goes through the horizontal pass once).  Per chunk:
//   load   the input rows the chunk needs and the band has not seen yet -> shared `patch`
//          (cp.async, 16 bytes per copy when rows are aligned, zero filled past the last column);
//          issued one chunk ahead, so it overlaps the previous chunk's vertical pass;
//   H pass one thread per flat output column, its <= KW weights live in registers: every NEW patch row
//          -> one row of T (shared).  Rows shared with the previous chunk are kept (moved to the top
//          of T), never recomputed;
//   V pass one thread per 4 consecutive flat columns: per output row one broadcast LDS.128 fetches
//          {weights, first T row}, then KH x (LDS.128 + 4 FMA) and one 128-bit store.
// Tap counts KH/KW are template parameters (loops fully unrolled); weights beyond a window's true
// size are zero.  Summation: horizontal then vertical, FMA, ascending taps (the bit-exact non-FMA
// order is aa_general.cu's job).
#include <stdlib.h>

#include <algorithm>

#include "aa_common.cuh"

namespace aa {
namespace {

constexpr int TXV = 64;         // float4 columns per band
constexpr int TXF = TXV * 4;    // flat output columns per band (= threads: one column each in the H pass)
constexpr int NTY = 4;          // thread rows in the V pass
constexpr int NT = TXV * NTY;   // 256 threads
constexpr int NWARP = NT / 32;

struct BParams {
  const void* in;
  void* out;  // float* or uint8_t* (epi.u8)
  OutEpi epi;
  Layout lin, lout;
  int Ci;
  const int32_t *h_start, *h_size, *w_start, *w_size;
  const float *h_w, *w_w;
  int h_pitch, w_pitch;
  int in_h, in_wf, out_h, out_wf;  // flat widths (pixels * Ci)
  int tiles_x;
  int ty;          // output rows per chunk
  int n_chunks;    // chunks per band: ceil(out_h / ty)
  int seg_chunks;  // chunks per CTA
  int pr;          // patch rows (max input rows one chunk spans)
  int tr;          // T rows (pr + KH - 1: the unrolled tap loop may read up to KH-1 rows past a window, with zero weight)
  int pcp;         // patch pitch in floats (multiple of 4)
  int vec_store;   // rows of out are 16-byte aligned -> float4 stores
  int vec_load;    // rows of in are 16-byte (f32) / 4-byte (u8) aligned and whole 4-element groups -> 16-byte cp.async / 32-bit loads
  FastDiv dci, dcp;  // division by Ci (flat column -> pixel) and by lin.Cp (plane -> image)
  int64_t plane0;  // first plane of this launch (planes are launched in slabs of <= 65535)
  RedoList* redo;  // float input: where a CTA that stored a NaN/Inf reports its segment (aa_common.cuh)
};

template <typename in_t> __device__ __forceinline__ float ldf(const in_t* p) { return (float)__ldg(p); }

template <int KH, int KW, bool GEN, typename in_t>
__global__ void __launch_bounds__(NT) aa_band_kernel(const BParams P) {
  constexpr int HR = (KH + 1 + 3) / 4;  // float4 per row record {w[KH], first T row}
  if constexpr (sizeof(in_t) == 4) aa_trigger_drain();
  extern __shared__ __align__(16) float smem[];
  const int TY = P.ty;
  float* Ts = smem;                                                       // [tr][TXF]
  float4* hrec = reinterpret_cast<float4*>(Ts + (size_t)P.tr * TXF);      // [2][TY][HR]
  float* patch = reinterpret_cast<float*>(hrec + 2 * TY * HR);            // [pr][pcp]
  const int tid = threadIdx.x;
  const int tx = tid % TXV, ty = tid / TXV;
  const int warp = tid >> 5, lane = tid & 31;
  // grid = (bands, segments, planes): no index decode; in and out share Cp (same memory format family)
  const int tile_x = blockIdx.x;
  const int cA = blockIdx.y * P.seg_chunks, cB = min(P.n_chunks, cA + P.seg_chunks);
  if (cA >= cB) return;
  const int64_t plane = (int64_t)blockIdx.z + P.plane0;
  const int Ci = P.Ci;
  int64_t pn = plane, pp = 0;
  if (P.lin.Cp > 1) {
    if (plane < (1ll << 32)) { pn = P.dcp.div((uint32_t)plane); pp = plane - pn * P.lin.Cp; }
    else { pn = plane / P.lin.Cp; pp = plane - pn * P.lin.Cp; }
  }
  const in_t* ip = (const in_t*)P.in + pn * P.lin.stride_n + pp * P.lin.stride_p;
  const int64_t op = pn * P.lout.stride_n + pp * P.lout.stride_p;  // element offset

  // ---- band geometry (once per CTA); starts and ends are non-decreasing in the output index
  const int of0 = tile_x * TXF, of1 = min(P.out_wf, of0 + TXF);
  const int oxa = (int)P.dci.div(of0), oxb = (int)P.dci.div(of1 - 1);
  const int c0 = __ldg(P.w_start + oxa) * Ci, c1 = (__ldg(P.w_start + oxb) + __ldg(P.w_size + oxb)) * Ci;
  const int nc = c1 - c0;
  const int pct = nc + (KW - 1) * Ci;  // patch columns touched by the unrolled tap loop
  const bool vload = P.vec_load != 0;
  const int lead = vload ? (c0 & 3) : 0;  // aligned patches start `lead` columns early
  const int pcp = P.pcp;
  const int64_t sh = P.lin.stride_h;
  const in_t* colbase = ip + (c0 - lead);

  // chunk extents: output rows [c*TY, ..) need input rows [ra, rb)
  auto extents = [&](int c, int& ra, int& rb) {
    const int oy0 = c * TY, oy1 = min(P.out_h, oy0 + TY);
    ra = __ldg(P.h_start + oy0);
    rb = __ldg(P.h_start + oy1 - 1) + __ldg(P.h_size + oy1 - 1);
  };
  // rows [rf, re) of the band -> patch buffer pb (row j of the buffer = input row rf + j)
  auto issue_patch = [&](float* pb, int rf, int re) {
    const int n = re - rf;
    const in_t* src = colbase + (int64_t)rf * sh;
    if constexpr (sizeof(in_t) == 4) {
      const uint32_t pbs = (uint32_t)__cvta_generic_to_shared(pb);
      if (vload) {
        // 16-byte copies, one warp per row; groups past the row's end are zero filled (src-size 0).
        // Columns between the band's last column and the row's end hold real (unused) data.
        const int nq = (lead + pct + 3) >> 2;                       // groups the tap loops may touch
        const int nql = min(nq, (P.in_wf - (c0 - lead)) >> 2);      // groups that exist in the row
        for (int r = warp; r < n; r += NWARP) {
          const in_t* srow = src + (int64_t)r * sh;
          uint32_t d = pbs + 4u * (r * pcp + 4 * lane);
          for (int q = lane; q < nq; q += 32, d += 512u) {
            const in_t* g = srow + 4 * min(q, nql - 1);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(q < nql ? 16 : 0) : "memory");
          }
        }
      } else {
        for (int r = warp; r < n; r += NWARP) {
          const in_t* srow = src + (int64_t)r * sh;
          uint32_t d = pbs + 4u * (r * pcp + lane);
          for (int c = lane; c < pct; c += 32, d += 128u) {
            const in_t* g = srow + min(c, nc - 1);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(g), "r"(c < nc ? 4 : 0) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
      // uint8: through registers (converted once here, not per tap), 4 loads in flight per thread
      if (vload) {
        // rows 4-byte aligned and made of whole 4-pixel groups: one 32-bit load = 4 pixels -> one 128-bit shared store
        const int nq = (lead + pct + 3) >> 2;                       // groups the tap loops may touch
        const int nql = min(nq, (P.in_wf - (c0 - lead)) >> 2);      // groups that exist in the row
        for (int r = warp; r < n; r += NWARP) {
          const in_t* srow = src + (int64_t)r * sh;
          float* drow = pb + r * pcp;
          for (int qb = lane; qb < nq; qb += 128) {
            uint32_t wv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int q = qb + 32 * j;
              wv[j] = q < nql ? __ldg(reinterpret_cast<const uint32_t*>(srow) + q) : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const int q = qb + 32 * j;
              if (q < nq) {
                float f[4];
                aa_unpack4(wv[j], f);
                *reinterpret_cast<float4*>(drow + 4 * q) = make_float4(f[0], f[1], f[2], f[3]);
              }
            }
          }
        }
      } else
      for (int r = warp; r < n; r += NWARP) {
        const in_t* srow = src + (int64_t)r * sh;
        float* drow = pb + r * pcp;
        for (int cb = lane; cb < pct; cb += 128) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int c = cb + 32 * j;
            v[j] = c < nc ? ldf(srow + c) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int c = cb + 32 * j;
            if (c < pct) drow[c] = v[j];
          }
        }
      }
    }
  };
  // this thread's row record of chunk c (threads < TY): {w[KH], offset of the first T row}
  auto load_rec = [&](int c, int ra, float (&rec)[HR * 4]) {
#pragma unroll
    for (int k = 0; k < HR * 4; k++) rec[k] = 0.f;
    const int oy = c * TY + tid;
    if (tid < TY && oy < P.out_h) {
      const int st = __ldg(P.h_start + oy), sz = __ldg(P.h_size + oy);
      const float* hr = P.h_w + (int64_t)oy * P.h_pitch;
#pragma unroll
      for (int k = 0; k < KH; k++) rec[k] = k < sz ? __ldg(hr + k) : 0.f;
      rec[KH] = __int_as_float((sz > 0 ? st - ra : 0) * TXF);  // empty window: any initialised rows
    }
  };
  auto store_rec = [&](int c, const float (&rec)[HR * 4]) {
    if (tid < TY) {
      float4* dst = hrec + ((c & 1) * TY + tid) * HR;
#pragma unroll
      for (int q = 0; q < HR; q++) dst[q] = make_float4(rec[4 * q], rec[4 * q + 1], rec[4 * q + 2], rec[4 * q + 3]);
    }
  };

  int ra, rb;
  extents(cA, ra, rb);
  issue_patch(patch, ra, rb);
  int ra_n = ra, rb_n = rb;
  if (cA + 1 < cB) extents(cA + 1, ra_n, rb_n);
  float rec[HR * 4];
  load_rec(cA, ra, rec);

  // ---- H pass setup: this thread's flat output column
  float w[KW];
  int soff = 0;
  {
    const int of = of0 + tid;
    if (of < of1) {
      const int ox = (int)P.dci.div(of);
      const int c = of - ox * Ci;
      const int st = __ldg(P.w_start + ox), sz = __ldg(P.w_size + ox);
      const float* wr = P.w_w + (int64_t)ox * P.w_pitch;
#pragma unroll
      for (int k = 0; k < KW; k++) w[k] = k < sz ? __ldg(wr + k) : 0.f;
      // an empty window (adjoint tables: a grad_in column no grad_out column reaches) starts past the patch: its
      // taps all have zero weight, point them at loaded data
      soff = sz > 0 ? st * Ci + c - c0 + lead : 0;
    } else {
#pragma unroll
      for (int k = 0; k < KW; k++) w[k] = 0.f;
    }
  }
  // ---- V pass setup: this thread's 4 flat output columns
  const int ofv = of0 + 4 * tx;
  const bool full = P.vec_store && (ofv + 4 <= of1);
  int coff[4], cch[4];  // per-column output offset and channel (the generic epilogue may be planar / per-channel)
#pragma unroll
  for (int i = 0; i < 4; i++) { coff[i] = ofv + i; cch[i] = 0; }
  if constexpr (GEN) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int oxi = (int)P.dci.div(min(ofv + i, P.out_wf - 1));
      cch[i] = (ofv + i) - oxi * Ci;
      coff[i] = P.epi.coloff(oxi, cch[i], Ci);
    }
  }
  constexpr bool CHK = sizeof(in_t) == 4;  // float input can carry NaN/Inf (aa_common.cuh: aa_exact_region)
  float4 chk = make_float4(0.f, 0.f, 0.f, 0.f);
  // T starts finite: rows past a window (zero weight) are read by the unrolled tap loop
  for (int r = 0; r < P.tr; r++) Ts[r * TXF + tid] = 0.f;
  store_rec(cA, rec);

  int rbase = ra, rdone = ra;  // T holds input rows [rbase, rdone) in its rows 0..
  int pa = ra;                 // first row of the patch in flight
  for (int c = cA; c < cB; c++) {
    const bool more = c + 1 < cB;
    const int pa_n = max(ra_n, rb);  // first input row the next chunk still needs loaded
    if constexpr (sizeof(in_t) == 4) asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (more) load_rec(c + 1, ra_n, rec);
    __syncthreads();  // patch of chunk c landed; every thread is done with the previous chunk's T rows
    // ---- rows shared with the previous chunk move to the top of T (own column only: no hazard)
    {
      const int keep = max(rdone - ra, 0);
      const int shift = ra - rbase;
      if (shift > 0) {  // 4 rows at a time: the loads of a group are all ahead of its stores (shift >= 1)
        float* tcol = Ts + tid;
        for (int j = 0; j < keep; j += 4) {
          float t[4];
#pragma unroll
          for (int i = 0; i < 4; i++) t[i] = tcol[(min(j + i, keep - 1) + shift) * TXF];
#pragma unroll
          for (int i = 0; i < 4; i++)
            if (j + i < keep) tcol[(j + i) * TXF] = t[i];
        }
      }
    }
    // ---- H pass: new rows [pa, rb) -> T rows (pa - ra)..
    {
      aa_hpass<KW>(patch + soff, pcp, Ci, w, Ts + (pa - ra) * TXF + tid, TXF, rb - pa);
    }
    rbase = ra;
    rdone = rb;
    __syncthreads();
    if (more) {
      issue_patch(patch, pa_n, rb_n);  // the patch is free again: overlaps this chunk's vertical pass
      store_rec(c + 1, rec);
    }
    // ---- V pass + store
    if (ofv < of1) {
      const int oy0 = c * TY, rows = min(P.out_h - oy0, TY);
      int64_t dst = op + (int64_t)(oy0 + ty) * P.lout.stride_h;  // row base; column offsets below
      const int64_t dstep = (int64_t)NTY * P.lout.stride_h;
      const float4* hr = hrec + (c & 1) * TY * HR;
#pragma unroll 2
      for (int oyl = ty; oyl < rows; oyl += NTY, dst += dstep) {
        float rc[HR * 4];
#pragma unroll
        for (int q = 0; q < HR; q++) {
          const float4 t4 = hr[oyl * HR + q];
          rc[4 * q] = t4.x; rc[4 * q + 1] = t4.y; rc[4 * q + 2] = t4.z; rc[4 * q + 3] = t4.w;
        }
        const float4* src = reinterpret_cast<const float4*>(Ts + __float_as_int(rc[KH]) + 4 * tx);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KH; k++) {
          const float4 v = src[k * TXV];
          aa_fma4(a, v, rc[k]);
        }
        if constexpr (CHK) aa_fma4(chk, a, 0.f);  // 0 * a stays 0 unless a holds a NaN/Inf
        if (!GEN && full && P.epi.kind == 0) {
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + dst + ofv) = a;
        } else if (!GEN && full && P.epi.kind == 1) {
          const unsigned int pk = aa_to_u8(a.x, P.epi.round) | (aa_to_u8(a.y, P.epi.round) << 8) |
                                  (aa_to_u8(a.z, P.epi.round) << 16) | (aa_to_u8(a.w, P.epi.round) << 24);
          *reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(P.out) + dst + ofv) = pk;
        } else {
          aa_store<GEN>(P.out, dst + coff[0], a.x, cch[0], P.epi);
          if (ofv + 1 < of1) aa_store<GEN>(P.out, dst + coff[1], a.y, cch[1], P.epi);
          if (ofv + 2 < of1) aa_store<GEN>(P.out, dst + coff[2], a.z, cch[2], P.epi);
          if (ofv + 3 < of1) aa_store<GEN>(P.out, dst + coff[3], a.w, cch[3], P.epi);
        }
      }
    }
    // ---- advance
    pa = pa_n;
    ra = ra_n;
    rb = rb_n;
    if (c + 2 < cB) extents(c + 2, ra_n, rb_n);
  }
  if constexpr (CHK) {  // once per segment: a non-finite value anywhere in it -> the drain kernel redoes it tap-exactly
    if (__syncthreads_or(aa_nonfinite(chk.x + chk.y + chk.z + chk.w)) && tid == 0 && P.redo)
      redo_push(P.redo, plane, -1, cA * TY, min(P.out_h, cB * TY), of0, of1);
  }
}

template <int KH, int KW, bool GEN, typename in_t>
int launch_k(BParams& P, int64_t planes, const BandedAxis& ah, int nc_max, cudaStream_t stream, GeomPlan& G) {
  constexpr int HR = (KH + 1 + 3) / 4;
  P.pcp = (nc_max + (KW - 1) * P.Ci + 3 + 3) & ~3;  // + up to 3 lead columns (aligned 16-byte copies); multiple of 4
  // chunk height: 16 output rows measured best at 0.75x..1x (bilinear and bicubic), 8 when 16 do not fit 4 CTAs/SM
  static const int ty_env = [] { const char* e = getenv("AA_BAND_TY"); return e ? atoi(e) : 0; }();  // tuning knob
  const int tys[2] = {ty_env > 0 ? ty_env : 16, 8};
  const size_t limits[2] = {56 * 1024, 113 * 1024};
  int best_ty = G.ty, best_pr = G.nr;
  size_t best_smem = G.smem;
  for (int li = 0; li < 2 && !best_ty; li++)
    for (int ti = 0; ti < 2 && !best_ty; ti++) {
      const int TY = tys[ti];
      int64_t nr = 1;  // exact row plan for this chunk height from the host mirror
      for (int64_t y0 = 0; y0 < P.out_h; y0 += TY) {
        const int64_t y1 = std::min<int64_t>(P.out_h, y0 + TY) - 1;
        nr = std::max<int64_t>(nr, (int64_t)ah.h_start[y1] + ah.h_size[y1] - ah.h_start[y0]);
      }
      if (nr > 1024) continue;
      const size_t smem = sizeof(float) * ((size_t)(nr + KH - 1) * TXF + (size_t)2 * TY * HR * 4 + (size_t)nr * P.pcp);
      if (smem > limits[li]) continue;
      best_ty = TY; best_pr = (int)nr; best_smem = smem;
    }
  if (!best_ty) return fail(AA_ERR_UNSUPPORTED, "band: input patch too large; use the streaming/general path");
  G.ty = best_ty; G.nr = best_pr; G.smem = best_smem;
  P.ty = best_ty; P.pr = best_pr; P.tr = best_pr + KH - 1;
  P.n_chunks = (P.out_h + P.ty - 1) / P.ty;
  if (planes <= 0 || P.n_chunks <= 0) return AA_OK;
  // segments: enough CTAs to fill the GPU many times over (tail effect), otherwise as long as possible
  static const int cta_env = [] { const char* e = getenv("AA_BAND_CTAS"); return e ? atoi(e) : 0; }();  // tuning knob
  const int64_t target_ctas = 148 * (cta_env > 0 ? cta_env : 64);
  const int64_t bands = (int64_t)P.tiles_x * planes;
  int64_t nseg = std::min<int64_t>(P.n_chunks, std::max<int64_t>(1, (target_ctas + bands - 1) / bands));
  P.seg_chunks = (int)((P.n_chunks + nseg - 1) / nseg);
  nseg = (P.n_chunks + P.seg_chunks - 1) / P.seg_chunks;
  if (nseg > 65535) return fail(AA_ERR_UNSUPPORTED, "band: too many row segments");
  auto kern = aa_band_kernel<KH, KW, GEN, in_t>;
  AA_CUDA_TRY(ensure_smem_attr(kern, ah.device, best_smem));
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
    P.plane0 = p0;
    const dim3 grid((unsigned)P.tiles_x, (unsigned)nseg, (unsigned)std::min<int64_t>(65535, planes - p0));
    kern<<<grid, NT, best_smem, stream>>>(P);
    AA_LAUNCH_CHECK("aa_band_kernel");
  }
  return AA_OK;
}

template <int KH, int KW, typename in_t>
int launch_gen(BParams& P, int64_t planes, const BandedAxis& ah, int nc_max, cudaStream_t stream, GeomPlan& G) {
  if (P.epi.generic()) return launch_k<KH, KW, true, in_t>(P, planes, ah, nc_max, stream, G);  // decode-adjacent epilogue
  return launch_k<KH, KW, false, in_t>(P, planes, ah, nc_max, stream, G);
}

template <int KH, typename in_t>
int launch_kh(BParams& P, int kw, int64_t nb, const BandedAxis& ah, int nc, cudaStream_t s, GeomPlan& G) {
  if (kw <= 2) return launch_gen<KH, 2, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 3) return launch_gen<KH, 3, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 4) return launch_gen<KH, 4, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 5) return launch_gen<KH, 5, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 7) return launch_gen<KH, 7, in_t>(P, nb, ah, nc, s, G);
  return fail(AA_ERR_UNSUPPORTED, "band: more than 7 horizontal taps");
}

template <typename in_t>
int launch_in(BParams& P, int kh, int kw, int64_t nb, const BandedAxis& ah, int nc, cudaStream_t s, GeomPlan& G) {
  if (kh <= 2) return launch_kh<2, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 3) return launch_kh<3, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 4) return launch_kh<4, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 5) return launch_kh<5, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 7) return launch_kh<7, in_t>(P, kw, nb, ah, nc, s, G);
  return fail(AA_ERR_UNSUPPORTED, "band: more than 7 vertical taps");
}

}  // namespace

int launch_band(const void* in, int in_dtype, const Layout& lin, void* out, const Layout& lout,
                const BandedAxis& ah, const BandedAxis& aw, int kh_max, int kw_max, OutEpi epi, cudaStream_t stream) {
  if (in_dtype != AA_F32 && in_dtype != AA_U8) return fail(AA_ERR_UNSUPPORTED, "band: f32/u8 input only");
  if (kh_max > 7 || kw_max > 7) return fail(AA_ERR_UNSUPPORTED, "band: more than 7 taps; use the streaming/general path");
  const int Ci = lin.Ci;
  if (aw.n_out * Ci >= (1ll << 30) || aw.n_in * Ci >= (1ll << 30) || ah.n_out >= (1ll << 30)) return fail(AA_ERR_UNSUPPORTED, "band: size limits");
  BParams P;
  P.in = in; P.out = out; P.epi = epi; P.lin = lin; P.lout = lout; P.Ci = Ci;
  P.h_start = ah.start; P.h_size = ah.size; P.h_w = (const float*)ah.w; P.h_pitch = ah.pitch;
  P.w_start = aw.start; P.w_size = aw.size; P.w_w = (const float*)aw.w; P.w_pitch = aw.pitch;
  P.in_h = (int)ah.n_in; P.in_wf = (int)(aw.n_in * Ci); P.out_h = (int)ah.n_out; P.out_wf = (int)(aw.n_out * Ci);
  P.tiles_x = (P.out_wf + TXF - 1) / TXF;
  const GeomKey gkey{ah.id, aw.id, 2 | (Ci << 8) | (in_dtype << 24) | ((epi.generic() ? 1 : 0) << 28)};
  GeomPlan G;
  const bool planned = geom_lookup(gkey, &G);
  if (planned && !G.ty) return fail(AA_ERR_UNSUPPORTED, "band: input patch too large; use the streaming/general path");
  int64_t nc = G.nc;
  if (!planned) {
    // exact column plan from the host mirror of the table (the row plan depends on the chunk height)
    nc = 1;
    for (int64_t f0 = 0; f0 < P.out_wf; f0 += TXF) {
      const int64_t f1 = std::min<int64_t>(P.out_wf, f0 + TXF) - 1;
      const int64_t x0 = f0 / Ci, x1 = f1 / Ci;
      nc = std::max<int64_t>(nc, ((int64_t)aw.h_start[x1] + aw.h_size[x1] - aw.h_start[x0]) * Ci);
    }
    G.nc = (int)std::min<int64_t>(nc, 1 << 30);
    if (nc > 4096) {
      geom_store(gkey, G);
      return fail(AA_ERR_UNSUPPORTED, "band: input patch too large; use the streaming/general path");
    }
  }
  P.vec_load = ((uintptr_t)in) % (in_dtype == AA_F32 ? 16 : 4) == 0 && lin.stride_h % 4 == 0 && lin.stride_n % 4 == 0 &&
               (lin.Cp == 1 || lin.stride_p % 4 == 0) && P.in_wf % 4 == 0;
  P.dci = FastDiv::make((uint32_t)Ci);
  P.dcp = FastDiv::make((uint32_t)(lin.Cp > 0 ? lin.Cp : 1));
  P.vec_store = (((uintptr_t)out) % (epi.kind == 1 ? 4 : 16) == 0) && (lout.stride_h % 4 == 0) && (lout.stride_n % 4 == 0) &&
                (lout.Cp == 1 || lout.stride_p % 4 == 0);
  P.redo = nullptr;
  if (in_dtype == AA_F32) {
    const int rl = redo_list(ah.device, stream, &P.redo);
    if (rl != AA_OK) return rl;
  }
  int rc = in_dtype == AA_F32 ? launch_in<float>(P, kh_max, kw_max, lin.planes, ah, (int)nc, stream, G)
                              : launch_in<uint8_t>(P, kh_max, kw_max, lin.planes, ah, (int)nc, stream, G);
  if (rc == AA_OK && in_dtype == AA_F32 && lin.planes > 0) {
    const RedoParams R{in, out, epi, lin, lout, Ci, ExactTabs{P.h_start, P.h_size, P.w_start, P.w_size, P.h_w, P.w_w, P.h_pitch, P.w_pitch},
                       P.out_h, P.out_wf, 1, 0, P.redo};
    rc = launch_redo(R, ah.device, stream);
  }
  if (!planned && (rc == AA_OK || rc == AA_ERR_UNSUPPORTED) && lin.planes > 0) {
    if (rc == AA_ERR_UNSUPPORTED) G.ty = 0;
    geom_store(gkey, G);
  }
  return rc;
}

}  // namespace aa
