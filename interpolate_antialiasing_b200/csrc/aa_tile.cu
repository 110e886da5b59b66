// aa_tile.cu -- K3: gather-form tile kernel for banded separable applies with FEW taps per output
// (sm_100a).  This is the shape of
//   * the backward (adjoint) pass of a downsampling forward: grad_in = Wh^T * grad_out * Ww, every
//     grad_input element gathers <= KT (3 bilinear, 5 bicubic) grad_output rows/columns -- no atomics, no
//     zero-fill pass (replaces /root/reference/step_two_dot_two/aa_interpolation_backward_impl.h:185-219,
//     whose scatter loop2d :80-108 + zero_() :215 it supersedes);
//   * the forward pass near scale 1 and when upsampling (K = 3..7 taps, aa_interpolation_impl.h:60-87).
// Such passes are output-(write-)bound, so the kernel is organised around the OUTPUT tile: one CTA
// produces TY x TXF flat outputs of one plane.
//   stage 0  the input patch the tile depends on -> shared memory (coalesced, zero padded so the tap
//            loops below need no bounds checks);
//   stage 1  horizontal pass: one thread per flat output column, its <= KW weights live in registers,
//            walks the patch rows -> T[rows][TXF] in shared memory;
//   stage 2  vertical pass: one thread per 4 consecutive flat columns, per output row one broadcast
//            LDS.128 fetches {weights, first row}, then KH x (LDS.128 + 4 FMA) and one 128-bit store.
// Tap counts KH/KW are template parameters (loops fully unrolled); weights beyond a window's true
// size are zeroed in-kernel.  Summation: horizontal then vertical, FMA, ascending taps (the
// bit-exact non-FMA order is aa_general.cu's job).
#include <algorithm>

#include <map>
#include <mutex>

#include "aa_common.cuh"

namespace aa {
namespace {

constexpr int TXV = 64;         // float4 columns per tile
constexpr int TXF = TXV * 4;    // flat output columns per tile (= threads: one column each in stage 1)
constexpr int NTY = 4;          // thread rows in stages 0 and 2
constexpr int NT = TXV * NTY;   // 256 threads

struct TParams {
  const void* in;
  void* out;  // float* or uint8_t* (epi.u8)
  OutEpi epi;
  Layout lin, lout;
  int Ci;
  const int32_t *h_start, *h_size, *w_start, *w_size;
  const float *h_w, *w_w;
  int h_pitch, w_pitch;
  int in_h, in_wf, out_h, out_wf;  // flat widths (pixels * Ci)
  int tiles_x, tiles_y;
  int ty;    // output rows per tile (64 / 32 / 16)
  int pr;    // patch rows (max input rows one tile spans)
  int tr;    // T rows: pr + the zero rows the unrolled vertical tap loop may touch past the last window
  int pcp;   // patch pitch in floats (max input flat cols per tile + (KW-1)*Ci, padded)
  int vec_store;  // rows of out are 16-byte aligned -> float4 stores
  int vec_load;   // rows of in are 16-byte (f32) / 4-byte (u8) aligned -> 16-byte cp.async / 32-bit loads in stage 0
  FastDiv dci, dcp;  // division by Ci (flat column -> pixel) and by lin.Cp (plane -> image)
  int64_t plane0;  // first plane of this launch (planes are launched in slabs of <= 65535)
  RedoList* redo;  // float input: where a CTA that stored a NaN/Inf reports its tile (aa_common.cuh)
};

template <typename in_t> __device__ __forceinline__ float ldf(const in_t* p) { return (float)__ldg(p); }

// VR = output rows one thread produces per vertical step.  VR = 4 (windows of 4 consecutive output rows start
// within 3 input rows of each other: every upsampling-like gather) loads the union of the 4 windows once
// -- KH+3 T rows and KH+3 broadcast weight quads {w_row0..w_row3} -- instead of 4 x KH rows: the vertical
// pass is shared-memory-bandwidth bound, and this halves its traffic.  VR = 1 handles any window spacing.
template <int KH, int VR> struct VRec {
  static constexpr int HR = (KH + 1 + 3) / 4;  // VR=1: float4 per row record {w[KH], first T row}
  static constexpr int U4 = KH + 3;            // VR=4: T rows a group of 4 output rows spans
  static constexpr int GR = U4 + 1;            // VR=4: float4 per group record {w quad per T row, first T row}
  // zero rows below the last window: what the unrolled tap loop can touch past it, + 1 so that a tile of empty
  // windows only (adjoint tables) still reads initialised rows
  static constexpr int ZR = (VR == 4 ? U4 - 1 : KH - 1) + 1;
  static __host__ __device__ constexpr int rec4(int ty) { return VR == 4 ? (ty / 4) * GR : ty * HR; }
};

template <int KH, int KW, int VR, bool GEN, typename in_t>
// register caps = what the kernels used before the NaN/Inf check was added (32 for the 2-3 tap shapes, whose small
// tiles fit 7 CTAs per SM; 40-48 for the others): left alone, ptxas takes ~48 for all of them and the 2-3 tap shapes
// lose two CTAs per SM (bilinear 2x upsampling 180 -> 199 us)
__global__ void __launch_bounds__(NT, GEN ? 2 : (KH <= 3 && VR == 1 ? 8 : 5)) aa_tile_kernel(const TParams P) {
  using R = VRec<KH, VR>;
  constexpr int HR = R::HR;
  if constexpr (sizeof(in_t) == 4) aa_trigger_drain();
  const int TY = P.ty;
  extern __shared__ __align__(16) float smem[];
  float* Ts = smem;                                                   // [tr][TXF]
  float4* hrec = reinterpret_cast<float4*>(Ts + (size_t)P.tr * TXF);  // row / group records
  float* patch = reinterpret_cast<float*>(hrec + R::rec4(TY));        // [pr][pcp]
  const int tid = threadIdx.x;
  const int tx = tid % TXV, ty = tid / TXV;
  // grid = (tiles_x, tiles_y, planes): no index decode; in and out share Cp (same memory format family)
  const int tile_x = blockIdx.x, tile_y = blockIdx.y;
  const int64_t plane = (int64_t)blockIdx.z + P.plane0;
  const int Ci = P.Ci;
  int64_t pn = plane, pp = 0;
  if (P.lin.Cp > 1) {
    if (plane < (1ll << 32)) { pn = P.dcp.div((uint32_t)plane); pp = plane - pn * P.lin.Cp; }
    else { pn = plane / P.lin.Cp; pp = plane - pn * P.lin.Cp; }
  }
  const in_t* ip = (const in_t*)P.in + pn * P.lin.stride_n + pp * P.lin.stride_p;
  const int64_t op = pn * P.lout.stride_n + pp * P.lout.stride_p;  // element offset

  // tile extents (starts and ends are non-decreasing in the output index)
  const int oy0 = tile_y * TY, oy1 = min(P.out_h, oy0 + TY);
  const int of0 = tile_x * TXF, of1 = min(P.out_wf, of0 + TXF);
  const int r0 = __ldg(P.h_start + oy0), r1 = __ldg(P.h_start + oy1 - 1) + __ldg(P.h_size + oy1 - 1);
  const int oxa = (int)P.dci.div(of0), oxb = (int)P.dci.div(of1 - 1);
  const int c0 = __ldg(P.w_start + oxa) * Ci, c1 = (__ldg(P.w_start + oxb) + __ldg(P.w_size + oxb)) * Ci;
  const int nr = r1 - r0, nc = c1 - c0;
  const int pct = nc + (KW - 1) * Ci;      // patch columns touched
  const int lead = P.vec_load ? (c0 & 3) : 0;  // aligned patches start `lead` columns early

  // ---- per-row (VR = 1) / per-group (VR = 4) records for stage 2
  if (tid < TY) {
    const int oy = oy0 + tid;
    if constexpr (VR == 1) {
      float rec[HR * 4];
#pragma unroll
      for (int k = 0; k < HR * 4; k++) rec[k] = 0.f;
      if (oy < oy1) {
        const int st = __ldg(P.h_start + oy), sz = __ldg(P.h_size + oy);
        const float* hr = P.h_w + (int64_t)oy * P.h_pitch;
#pragma unroll
        for (int k = 0; k < KH; k++) rec[k] = k < sz ? __ldg(hr + k) : 0.f;
        rec[KH] = __int_as_float((sz > 0 ? st - r0 : 0) * TXF);  // empty window: any initialised rows
      }
#pragma unroll
      for (int q = 0; q < HR; q++) hrec[tid * HR + q] = make_float4(rec[4 * q], rec[4 * q + 1], rec[4 * q + 2], rec[4 * q + 3]);
    } else {
      // thread = output row j of group g: column j of the group's [U4][4] weight matrix (dense over the union
      // of the 4 windows, zero where a row's window does not reach)
      const int g = tid >> 2, j = tid & 3;
      const int oyg = oy0 + 4 * g;
      if (oyg < oy1) {
        float* gw = reinterpret_cast<float*>(hrec + g * R::GR);
        const int base = __ldg(P.h_start + oyg);
#pragma unroll
        for (int u = 0; u < R::U4; u++) gw[4 * u + j] = 0.f;
        if (oy < oy1) {
          const int st = __ldg(P.h_start + oy), sz = __ldg(P.h_size + oy);
          const float* hr = P.h_w + (int64_t)oy * P.h_pitch;
          const int d = st - base;  // 0..3 (host-checked)
#pragma unroll
          for (int k = 0; k < KH; k++)
            if (k < sz) gw[4 * (d + k) + j] = __ldg(hr + k);
        }
        if (j == 0) gw[4 * R::U4] = __int_as_float((base - r0 < nr ? base - r0 : 0) * TXF);  // all-empty group: any initialised rows
      }
    }
  }
  // ---- stage 0: input patch -> shared (zero padded).  All copies of a thread are in flight at once:
  // f32 goes global->shared with cp.async (LDGSTS, zero-fill via src-size 0), u8 through registers in
  // batches of 8 loads.
  {
    const in_t* src = ip + (int64_t)r0 * P.lin.stride_h + c0;
    if constexpr (sizeof(in_t) == 4) {
      if (P.vec_load) {
        // 16-byte copies: the patch starts at the aligned column below c0 (`lead` extra elements), its
        // pitch is a multiple of 4 so chunk i of the CTA lands at patch + 4 i; the copy's src-size
        // (0..16 bytes) zero-fills past the last valid column / row.
        const in_t* srca = src - lead;
        const int nq = P.pcp >> 2;      // chunks per patch row
        const int nca = lead + nc;      // valid elements per patch row
        const int total = nr * nq;
        const int dr = NT / nq, dq = NT - dr * nq;
        int r = tid / nq, q = tid - r * nq;
        uint32_t d = (uint32_t)__cvta_generic_to_shared(patch) + 16u * tid;
        for (int i = tid; i < total; i += NT, d += 16u * NT) {
          int bytes = min(max((nca - 4 * q) * 4, 0), 16);
          if (r >= nr) bytes = 0;
          const in_t* g = bytes ? srca + (int64_t)r * P.lin.stride_h + 4 * q : srca;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(bytes) : "memory");
          q += dq; r += dr;
          if (q >= nq) { q -= nq; r++; }
        }
      } else {
      for (int r = ty; r < nr; r += NTY) {
        const in_t* srow = src + (int64_t)r * P.lin.stride_h;
        const uint32_t drow = (uint32_t)__cvta_generic_to_shared(patch + r * P.pcp);
        const bool rok = r < nr;
        for (int c = tx; c < pct; c += TXV) {
          const bool ok = rok && c < nc;
          const in_t* g = ok ? srow + c : src;  // any valid address when zero-filling
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(drow + 4u * c), "l"(g), "r"(ok ? 4 : 0) : "memory");
        }
      }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else if (P.vec_load) {
      // uint8, rows 4-byte aligned: one 32-bit load = 4 pixels -> PRMT/FADD conversion -> one 128-bit shared store
      // (byte loads made the uint8 upsampling shapes patch-load-bound: 0.34-0.42 of peak where fp32 reaches 0.8).
      // Groups that would cross the end of the row (the tensor may end there) are read byte by byte.
      const in_t* srca = src - lead;
      const int nq = P.pcp >> 2;                       // 4-pixel groups per patch row
      const int nca = lead + nc;                       // valid elements per patch row
      const int row_left = P.in_wf - (c0 - lead);      // elements from the patch's first column to the end of the row
      const int total = nr * nq;
      constexpr int B = 4;
      for (int i0 = tid; i0 < total; i0 += NT * B) {
        uint32_t wv[B];
#pragma unroll
        for (int j = 0; j < B; j++) {
          const int i = i0 + j * NT;
          const int r = i / nq, q = i - r * nq;
          wv[j] = 0u;
          if (i < total && 4 * q < nca) {
            const in_t* g = srca + (int64_t)r * P.lin.stride_h + 4 * q;
            if (4 * q + 4 <= row_left) wv[j] = __ldg(reinterpret_cast<const uint32_t*>(g));
            else
              for (int e = 0; 4 * q + e < row_left; e++) wv[j] |= (uint32_t)__ldg(g + e) << (8 * e);
          }
        }
#pragma unroll
        for (int j = 0; j < B; j++) {
          const int i = i0 + j * NT;
          if (i < total) {
            float f[4];
            aa_unpack4(wv[j], f);
            *reinterpret_cast<float4*>(patch + 4 * i) = make_float4(f[0], f[1], f[2], f[3]);
          }
        }
      }
    } else {
      constexpr int B = 8;
      const int ncol_it = (pct + TXV - 1) / TXV;       // column iterations per row
      const int nrow_it = (nr - ty + NTY - 1) / NTY;  // row iterations of this thread
      const int total = nrow_it > 0 ? nrow_it * ncol_it : 0;
      for (int i0 = 0; i0 < total; i0 += B) {
        float v[B];
#pragma unroll
        for (int j = 0; j < B; j++) {
          const int i = i0 + j;
          const int r = ty + (i / ncol_it) * NTY, c = tx + (i % ncol_it) * TXV;
          v[j] = (i < total && r < nr && c < nc) ? ldf(src + (int64_t)r * P.lin.stride_h + c) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < B; j++) {
          const int i = i0 + j;
          const int r = ty + (i / ncol_it) * NTY, c = tx + (i % ncol_it) * TXV;
          if (i < total && c < pct) patch[r * P.pcp + c] = v[j];
        }
      }
    }
  }
  // ---- stage 1 setup: this thread's flat output column
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
  if constexpr (sizeof(in_t) == 4) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- stage 1: horizontal pass -> Ts
  constexpr bool CHK = sizeof(in_t) == 4;  // float input can carry NaN/Inf (aa_common.cuh: aa_exact_region)
  float2 chk = make_float2(0.f, 0.f);
  aa_hpass<KW, CHK>(patch + soff, P.pcp, Ci, w, Ts + tid, TXF, nr, chk);
#pragma unroll
  for (int r = 0; r < R::ZR; r++) Ts[(nr + r) * TXF + tid] = 0.f;  // rows past the last window: zero weight, finite value
  if constexpr (CHK) {  // a non-finite value went into T: the drain kernel redoes this tile tap-exactly (aa_common.cuh)
    if (__syncthreads_or(aa_nonfinite(chk.x + chk.y)) && tid == 0 && P.redo) redo_push(P.redo, plane, -1, oy0, oy1, of0, of1);
  } else {
    __syncthreads();
  }
  // ---- stage 2: vertical pass + store
  const int ofv = of0 + 4 * tx;
  if (ofv < of1) {
    int64_t dst = op + (int64_t)(oy0 + ty) * P.lout.stride_h;  // row base; column offsets below
    const int64_t dstep = (int64_t)NTY * P.lout.stride_h;
    const bool full = P.vec_store && (ofv + 4 <= of1);
    int coff[4], cch[4];  // per-column output offset and channel (the generic epilogue may be planar / per-channel)
#pragma unroll
    for (int i = 0; i < 4; i++) { coff[i] = ofv + i; cch[i] = 0; }
    if constexpr (GEN) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int oxi = (int)P.dci.div(ofv + i);
        cch[i] = (ofv + i) - oxi * Ci;
        coff[i] = P.epi.coloff(oxi, cch[i], Ci);
      }
    }
    auto store4 = [&](int64_t d, const float4& a) {
      if (!GEN && full && P.epi.kind == 0) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + d + ofv) = a;
      } else if (!GEN && full && P.epi.kind == 1) {
        const unsigned int pk = aa_to_u8(a.x, P.epi.round) | (aa_to_u8(a.y, P.epi.round) << 8) |
                                (aa_to_u8(a.z, P.epi.round) << 16) | (aa_to_u8(a.w, P.epi.round) << 24);
        *reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(P.out) + d + ofv) = pk;
      } else {
        aa_store<GEN>(P.out, d + coff[0], a.x, cch[0], P.epi);
        if (ofv + 1 < of1) aa_store<GEN>(P.out, d + coff[1], a.y, cch[1], P.epi);
        if (ofv + 2 < of1) aa_store<GEN>(P.out, d + coff[2], a.z, cch[2], P.epi);
        if (ofv + 3 < of1) aa_store<GEN>(P.out, d + coff[3], a.w, cch[3], P.epi);
      }
    };
    if constexpr (VR == 1) {
#pragma unroll 2
      for (int oyl = ty; oyl < oy1 - oy0; oyl += NTY, dst += dstep) {
        float rec[HR * 4];
#pragma unroll
        for (int q = 0; q < HR; q++) {
          const float4 t4 = hrec[oyl * HR + q];
          rec[4 * q] = t4.x; rec[4 * q + 1] = t4.y; rec[4 * q + 2] = t4.z; rec[4 * q + 3] = t4.w;
        }
        const float4* src = reinterpret_cast<const float4*>(Ts + __float_as_int(rec[KH]) + 4 * tx);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KH; k++) {
          const float4 v = src[k * TXV];
          aa_fma4(a, v, rec[k]);
        }
        store4(dst, a);
      }
    } else {
      const int rows = oy1 - oy0;
      const int64_t sho = P.lout.stride_h;
      for (int g = ty; 4 * g < rows; g += NTY) {
        const float4* gr = hrec + g * R::GR;
        const float4* src = reinterpret_cast<const float4*>(Ts + __float_as_int(reinterpret_cast<const float*>(gr + R::U4)[0]) + 4 * tx);
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
        for (int u = 0; u < R::U4; u++) {
          const float4 wq = gr[u];
          const float4 v = src[u * TXV];
          aa_fma4(a0, v, wq.x); aa_fma4(a1, v, wq.y); aa_fma4(a2, v, wq.z); aa_fma4(a3, v, wq.w);
        }
        const int64_t d = op + (int64_t)(oy0 + 4 * g) * sho;
        store4(d, a0);
        if (4 * g + 1 < rows) store4(d + sho, a1);
        if (4 * g + 2 < rows) store4(d + 2 * sho, a2);
        if (4 * g + 3 < rows) store4(d + 3 * sho, a3);
      }
    }
  }
}

template <int KH, int KW, int VR, bool GEN, typename in_t>
int launch_ty(TParams& P, int TY, int64_t planes, const BandedAxis& ah, size_t smem_limit, cudaStream_t stream, GeomPlan& G) {
  using R = VRec<KH, VR>;
  int64_t nr = 1;
  if (G.ty) {  // planned before: only the chosen tile height is tried, with its recorded patch height
    if (G.ty != TY) return AA_ERR_UNSUPPORTED;
    nr = G.nr;
  } else {
    // exact row plan for this tile height from the host mirror
    for (int64_t y0 = 0; y0 < P.out_h; y0 += TY) {
      const int64_t y1 = std::min<int64_t>(P.out_h, y0 + TY) - 1;
      nr = std::max<int64_t>(nr, (int64_t)ah.h_start[y1] + ah.h_size[y1] - ah.h_start[y0]);
    }
  }
  if (nr > 1024) return fail(AA_ERR_UNSUPPORTED, "tile: input patch too large; use the streaming/general path");
  P.ty = TY;
  P.pr = (int)nr;
  P.tr = (int)nr + R::ZR;
  P.tiles_y = (P.out_h + TY - 1) / TY;
  const size_t smem = sizeof(float) * ((size_t)P.tr * TXF + (size_t)R::rec4(TY) * 4 + (size_t)P.pr * P.pcp);
  if (smem > smem_limit) return fail(AA_ERR_UNSUPPORTED, "tile: input patch too large; use the streaming/general path");
  if (planes <= 0) return AA_OK;
  if (P.tiles_y > 65535) return fail(AA_ERR_UNSUPPORTED, "tile: too many row tiles");
  auto kern = aa_tile_kernel<KH, KW, VR, GEN, in_t>;
  AA_CUDA_TRY(ensure_smem_attr(kern, ah.device, smem));
  G.ty = TY; G.nr = (int)nr; G.smem = smem;
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {
    P.plane0 = p0;
    const dim3 grid((unsigned)P.tiles_x, (unsigned)P.tiles_y, (unsigned)std::min<int64_t>(65535, planes - p0));
    kern<<<grid, NT, smem, stream>>>(P);
    AA_LAUNCH_CHECK("aa_tile_kernel");
  }
  return AA_OK;
}

template <int KH, int KW, int VR, typename in_t>
int launch_vr(TParams& P, int64_t planes, const BandedAxis& ah, cudaStream_t stream, GeomPlan& G) {
  // tall tiles amortise the per-CTA setup when the patch stays small (upsampling-like gathers)
  if (P.epi.generic()) {  // decode-adjacent epilogue: its own instantiations, two tile heights
    int rg = launch_ty<KH, KW, VR, true, in_t>(P, 32, planes, ah, 72 * 1024, stream, G);
    if (rg != AA_ERR_UNSUPPORTED) return rg;
    return launch_ty<KH, KW, VR, true, in_t>(P, 16, planes, ah, 72 * 1024, stream, G);
  }
  int rc = launch_ty<KH, KW, VR, false, in_t>(P, 64, planes, ah, 40 * 1024, stream, G);
  if (rc != AA_ERR_UNSUPPORTED) return rc;
  rc = launch_ty<KH, KW, VR, false, in_t>(P, 32, planes, ah, 72 * 1024, stream, G);
  if (rc != AA_ERR_UNSUPPORTED) return rc;
  return launch_ty<KH, KW, VR, false, in_t>(P, 16, planes, ah, 72 * 1024, stream, G);
}

template <int KH, int KW, typename in_t>
int launch_k(TParams& P, int64_t planes, const BandedAxis& ah, int nc_max, cudaStream_t stream, GeomPlan& G) {
  P.pcp = (nc_max + (KW - 1) * P.Ci + 3 + 3) & ~3;  // + up to 3 lead columns (aligned 16-byte copies); multiple of 4
  // 4 output rows per vertical step when every aligned group of 4 rows starts within 3 input rows and there
  // are enough taps to share (measured: +15-25 % for the bicubic gathers, -7 % for the 2-3 tap bilinear ones)
  if constexpr (KH >= 4) {
    bool vr4 = G.vr4 != 0;
    if (!G.ty) {
      vr4 = true;
      for (int64_t y = 0; y < P.out_h && vr4; y += 4)
        vr4 = ah.h_start[std::min<int64_t>(P.out_h - 1, y + 3)] - ah.h_start[y] <= 3;
      G.vr4 = vr4 ? 1 : 0;
    }
    if (vr4) return launch_vr<KH, KW, 4, in_t>(P, planes, ah, stream, G);
  }
  return launch_vr<KH, KW, 1, in_t>(P, planes, ah, stream, G);
}

template <int KH, typename in_t>
int launch_kh(TParams& P, int kw, int64_t nb, const BandedAxis& ah, int nc, cudaStream_t s, GeomPlan& G) {
  if (kw <= 2) return launch_k<KH, 2, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 3) return launch_k<KH, 3, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 4) return launch_k<KH, 4, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 5) return launch_k<KH, 5, in_t>(P, nb, ah, nc, s, G);
  if (kw <= 7) return launch_k<KH, 7, in_t>(P, nb, ah, nc, s, G);
  return fail(AA_ERR_UNSUPPORTED, "tile: more than 7 horizontal taps");
}

template <typename in_t>
int launch_in(TParams& P, int kh, int kw, int64_t nb, const BandedAxis& ah, int nc, cudaStream_t s, GeomPlan& G) {
  if (kh <= 2) return launch_kh<2, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 3) return launch_kh<3, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 4) return launch_kh<4, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 5) return launch_kh<5, in_t>(P, kw, nb, ah, nc, s, G);
  if (kh <= 7) return launch_kh<7, in_t>(P, kw, nb, ah, nc, s, G);
  return fail(AA_ERR_UNSUPPORTED, "tile: more than 7 vertical taps");
}

std::mutex g_geom_mu;
std::map<GeomKey, GeomPlan> g_geom;

}  // namespace

bool geom_lookup(const GeomKey& k, GeomPlan* p) {
  std::lock_guard<std::mutex> lock(g_geom_mu);
  auto it = g_geom.find(k);
  if (it == g_geom.end()) return false;
  *p = it->second;
  return true;
}
void geom_store(const GeomKey& k, const GeomPlan& p) {
  std::lock_guard<std::mutex> lock(g_geom_mu);
  if (g_geom.size() > 4096) g_geom.clear();
  g_geom[k] = p;
}
cudaError_t ensure_smem_attr_impl(const void* kern, int device, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> have;
  std::lock_guard<std::mutex> lock(mu);
  auto it = have.find({kern, device});
  if (it != have.end() && it->second >= smem) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) have[{kern, device}] = smem;
  return e;
}
void tile_plan_clear() {
  std::lock_guard<std::mutex> lock(g_geom_mu);
  g_geom.clear();
}

int launch_tile(const void* in, int in_dtype, const Layout& lin, void* out, const Layout& lout,
                const BandedAxis& ah, const BandedAxis& aw, int kh_max, int kw_max, OutEpi epi, cudaStream_t stream) {
  if (in_dtype != AA_F32 && in_dtype != AA_U8) return fail(AA_ERR_UNSUPPORTED, "tile: f32/u8 input only");
  if (kh_max > 7 || kw_max > 7) return fail(AA_ERR_UNSUPPORTED, "tile: more than 7 taps; use the streaming/general path");
  const int Ci = lin.Ci;
  if (aw.n_out * Ci >= (1ll << 30) || aw.n_in * Ci >= (1ll << 30) || ah.n_out >= (1ll << 30)) return fail(AA_ERR_UNSUPPORTED, "tile: size limits");
  TParams P;
  P.in = in; P.out = out; P.epi = epi; P.lin = lin; P.lout = lout; P.Ci = Ci;
  P.h_start = ah.start; P.h_size = ah.size; P.h_w = (const float*)ah.w; P.h_pitch = ah.pitch;
  P.w_start = aw.start; P.w_size = aw.size; P.w_w = (const float*)aw.w; P.w_pitch = aw.pitch;
  P.in_h = (int)ah.n_in; P.in_wf = (int)(aw.n_in * Ci); P.out_h = (int)ah.n_out; P.out_wf = (int)(aw.n_out * Ci);
  P.tiles_x = (P.out_wf + TXF - 1) / TXF;
  const GeomKey gkey{ah.id, aw.id, 1 | (Ci << 8) | (in_dtype << 24) | ((epi.generic() ? 1 : 0) << 28)};
  GeomPlan G;
  const bool planned = geom_lookup(gkey, &G);
  if (planned && !G.ty) return fail(AA_ERR_UNSUPPORTED, "tile: input patch too large; use the streaming/general path");
  int64_t nc = G.nc;
  if (!planned) {
    // exact column plan from the host mirror of the table (the row plan depends on the tile height)
    nc = 1;
    for (int64_t f0 = 0; f0 < P.out_wf; f0 += TXF) {
      const int64_t f1 = std::min<int64_t>(P.out_wf, f0 + TXF) - 1;
      const int64_t x0 = f0 / Ci, x1 = f1 / Ci;
      nc = std::max<int64_t>(nc, ((int64_t)aw.h_start[x1] + aw.h_size[x1] - aw.h_start[x0]) * Ci);
    }
    G.nc = (int)std::min<int64_t>(nc, 1 << 30);
    if (nc > 4096) {
      geom_store(gkey, G);  // ty == 0: remembered as not eligible
      return fail(AA_ERR_UNSUPPORTED, "tile: input patch too large; use the streaming/general path");
    }
  }
  P.vec_load = ((uintptr_t)in) % (in_dtype == AA_F32 ? 16 : 4) == 0 && lin.stride_h % 4 == 0 && lin.stride_n % 4 == 0 &&
               (lin.Cp == 1 || lin.stride_p % 4 == 0);
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
