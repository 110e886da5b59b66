// aa_stream_common.cuh -- pieces shared by the two variants of the streaming forward kernel
// (aa_stream.cu: plain 128-bit global loads; aa_stream_tma.cu: cp.async.bulk row staging).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "aa_common.cuh"

namespace aa {
namespace stream_detail {

constexpr int kMaxA = 6;

struct SParams {
  const void* in;
  void* out;    // float* or uint8_t* (epi.u8)
  OutEpi epi;
  Layout lin, lout;
  int Ci;
  int64_t H, oH, oW;
  const float* slot_h;  // [H][RS]
  int RS;
  const int32_t *xmin_h, *xsize_h;
  const int32_t *xmin_w, *xsize_w;
  const float* w_w;  // [oW][Kw]
  int Kw;
  RedoList* redo;  // float input: where a CTA that stored a NaN/Inf reports its unit range (aa_common.cuh)
  int n_strips, strip_ox;
  int64_t total_units;  // planes * n_strips * oH
  int vw;               // columns of a Vs row (strip capacity); the row pitch is vs_pitch(vw)
  int pad;              // q = 0/1/2/4: Vs rows are stored with q pad words per 32 columns (bank-conflict-free gathers at
                        // power-of-two lane strides); pos(f) = f + (f >> 5) * q
  int vr;               // rows of Vs
  int tg;               // buffered rows that trigger a horizontal phase
  int aln;              // alignment (elements) of a strip's first flat element
  int in_pitch;         // TMA variant: bytes between staged input rows in shared memory
  int kp;               // pitch of the strip's weight table, made odd so that lanes (= different columns) reading tap j
                        // hit different banks: pairs: Kw + max(xmin[o+1]-xmin[o]) (float2 units); single columns: Kw (floats)
  int pairs;            // horizontal phase computes pairs of adjacent output columns (wide strips)
  int wtab_bytes;       // bytes reserved for the weight table (pinfo follows, 16-byte aligned)
};

// ---- vector loads (read-once data: bypass L1 allocation) -------------------------------------
// Loads return RAW registers (floats, or packed bytes for uint8 input); expand() turns them into
// floats right before the FMAs, so a batch of U rows in flight costs U*VEC/4 registers for uint8.
template <typename in_t, int VEC> struct Raw;
template <int VEC> struct Raw<float, VEC> { static constexpr int N = VEC; using T = float; };
template <int VEC> struct Raw<uint8_t, VEC> { static constexpr int N = VEC / 4; using T = uint32_t; };

template <typename in_t, int VEC> struct VLoad;
template <> struct VLoad<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  }
};
template <> struct VLoad<float, 2> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
  }
};
template <> struct VLoad<float, 1> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) {
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
  }
};
template <> struct VLoad<uint8_t, 16> {
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&r)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
  }
};
template <> struct VLoad<uint8_t, 8> {
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&r)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(p));
  }
};
template <> struct VLoad<uint8_t, 4> {
  static __device__ __forceinline__ void ld(const uint8_t* p, uint32_t (&r)[1]) {
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r[0]) : "l"(p));
  }
};
__device__ __forceinline__ void unpack4(uint32_t w, float* v) { aa_unpack4(w, v); }  // aa_common.cuh
template <int VEC> __device__ __forceinline__ void expand(const float (&r)[VEC], float (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC; i++) v[i] = r[i];
}
template <int VEC> __device__ __forceinline__ void expand(const uint32_t (&r)[VEC / 4], float (&v)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC / 4; i++) unpack4(r[i], v + 4 * i);
}

// acc[a][:] += w[a] * v[:] for the A open output rows; packed FFMA2 (two fp32 FMAs per issue slot,
// scalar weight broadcast) whenever the thread owns an even number of elements.
template <int A, int VEC>
__device__ __forceinline__ void vfma(float (&acc)[A][VEC], const float (&v)[VEC], const float* rw) {
  if constexpr (VEC % 2 == 0) {
#pragma unroll
    for (int a = 0; a < A; a++) {
      const float2 w2 = make_float2(rw[a], rw[a]);
#pragma unroll
      for (int e = 0; e < VEC; e += 2) {
        const float2 r = __ffma2_rn(w2, make_float2(v[e], v[e + 1]), make_float2(acc[a][e], acc[a][e + 1]));
        acc[a][e] = r.x;
        acc[a][e + 1] = r.y;
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < A; a++)
#pragma unroll
      for (int e = 0; e < VEC; e++) acc[a][e] = fmaf(rw[a], v[e], acc[a][e]);
  }
}

template <int VEC> __device__ __forceinline__ void store_vec(float* dst, const float* a) {
  if constexpr (VEC % 4 == 0) {
#pragma unroll
    for (int i = 0; i < VEC / 4; i++)
      reinterpret_cast<float4*>(dst)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(a[0], a[1]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i++) dst[i] = a[i];
  }
}


// ---- horizontal phase (shared by both variants) ---------------------------------------------
// Adjacent output columns have heavily overlapping windows (stride scale_w, length 2*support_w), so a
// thread computes a PAIR of adjacent columns (same channel) for RPT buffered rows: every Vs element it
// loads feeds both columns, and the two weights arrive in one 64-bit load.  Per strip the kernel builds
//   Wp[pair][KP]     float2 {w_a[j], w_b[j - shift]} zero padded, shift = xmin_b - xmin_a
//   pinfo[pair*Ci+c] int4   {first tap's offset in a Vs row, pair*KP, union window length | has_b<<16 | c<<20,
//                            offset of column a inside an output row (OutEpi::coloff)}
// so the phase itself needs no integer division and no per-column table lookups.
// Row pitch of Vs (floats) and the position of column f inside a row.  With padding, q unused words follow
// every 32 columns: a gather whose lanes are 4, 8, 16, 32... words apart (integer scale factors, times the channel
// interleave) then spreads over all 32 banks instead of 8, 4, 2, 1.  q is chosen per plan (plan_stream) from the
// strip's actual window starts; padded kernels reserve the pitch of the largest q.
constexpr int kPadMaxQ = 4;
__host__ __device__ constexpr int vs_pitch(int vw, bool pad) { return pad ? vw + (vw / 32) * kPadMaxQ : vw; }
__device__ __forceinline__ int vs_pos(int f, int q) { return f + (f >> 5) * q; }

struct HRole {        // which (row group, pair-column) items a thread owns; fixed per strip
  int rg0, rg_par;    // first row group and stride over row groups
  int cf0, cf_step;   // first pair-column and stride over pair-columns
};
__device__ __forceinline__ HRole hphase_role(int t, int nthreads, int npc) {
  HRole r;
  const int nw = nthreads >> 5, warp = t >> 5, lane = t & 31;
  const int wpr = (npc + 31) >> 5;  // warps needed to cover one row group
  if (wpr >= nw) { r.rg0 = 0; r.rg_par = 1; r.cf0 = t; r.cf_step = nthreads; }
  else {
    r.rg_par = nw / wpr;
    r.rg0 = warp / wpr;
    if (r.rg0 >= r.rg_par) r.rg0 = 1 << 20;  // idle warp
    r.cf0 = (warp - (warp / wpr) * wpr) * 32 + lane;
    r.cf_step = wpr * 32;
  }
  return r;
}
// Called by all `nthreads` threads between two barriers when the CTA moves to a new strip.
// P.pairs == 0 builds the single-column form of the same tables instead (narrow strips, where halving
// the number of independent items would starve the phase): Wp is then a float array with pitch KP.
__device__ __forceinline__ void strip_setup(const SParams& P, int t, int nthreads, int ox0, int ox1, float2* Wp, int4* pinfo,
                                            int* fl0_out, int* npc_out) {
  const int Ci = P.Ci, Kw = P.Kw, KP = P.kp;
  const int nox = ox1 - ox0;
  const int fl0 = (__ldg(P.xmin_w + ox0) * Ci) & ~(P.aln - 1);
  *fl0_out = fl0;
  if (!P.pairs) {
    float* Ws = reinterpret_cast<float*>(Wp);
    for (int i = t; i < nox * KP; i += nthreads) {
      const int oxl = i / KP, j = i - oxl * KP;
      Ws[i] = j < Kw ? __ldg(P.w_w + (int64_t)(ox0 + oxl) * Kw + j) : 0.f;
    }
    for (int cf = t; cf < nox * Ci; cf += nthreads) {
      const int oxl = cf / Ci, c = cf - oxl * Ci;
      pinfo[cf] = make_int4(__ldg(P.xmin_w + ox0 + oxl) * Ci + c - fl0, oxl * KP, __ldg(P.xsize_w + ox0 + oxl) | (c << 20),
                            P.epi.coloff(ox0 + oxl, c, Ci));
    }
    *npc_out = nox * Ci;
    return;
  }
  const int np = (nox + 1) >> 1;
  for (int i = t; i < np * KP; i += nthreads) {
    const int p = i / KP, j = i - p * KP;
    const int oa = ox0 + 2 * p, ob = oa + 1;
    const int xa = __ldg(P.xmin_w + oa), sa = __ldg(P.xsize_w + oa);
    float wa = j < sa ? __ldg(P.w_w + (int64_t)oa * Kw + j) : 0.f, wb = 0.f;
    if (ob < ox1) {
      const int jb = j - (__ldg(P.xmin_w + ob) - xa);
      if (jb >= 0 && jb < __ldg(P.xsize_w + ob)) wb = __ldg(P.w_w + (int64_t)ob * Kw + jb);
    }
    Wp[i] = make_float2(wa, wb);
  }
  for (int pc = t; pc < np * Ci; pc += nthreads) {
    const int p = pc / Ci, c = pc - p * Ci;
    const int oa = ox0 + 2 * p, ob = oa + 1;
    const int xa = __ldg(P.xmin_w + oa);
    int len = __ldg(P.xsize_w + oa), hasb = 0;
    if (ob < ox1) { len = max(len, __ldg(P.xmin_w + ob) - xa + __ldg(P.xsize_w + ob)); hasb = 1; }
    pinfo[pc] = make_int4(xa * Ci + c - fl0, p * KP, len | (hasb << 16) | (c << 20), P.epi.coloff(oa, c, Ci));
  }
  *npc_out = np * Ci;
}
// Gather over the buffered rows [0, cnt) of Vs -> output rows gbase..gbase+cnt-1.  Trip counts are
// warp-uniform (longest union window in the warp); past a lane's own window the weights read are the
// zero padding and the data pointer stops advancing, so no element outside the true windows is touched.
template <int RPT, int VW, bool GEN, bool PAD, bool CHK>
__device__ __forceinline__ bool hphase_run_pairs(const float* __restrict__ Vs, const float2* __restrict__ Wp,
                                           const int4* __restrict__ pinfo, void* __restrict__ op, int64_t op_off,
                                           const OutEpi& epi, int64_t out_stride_h, int Ci, int npc, const HRole role, int gbase,
                                           int cnt, int padsh) {
  const int nrg = (cnt + RPT - 1) / RPT;
  bool bad = false;  // CHK: a non-finite value was stored (the caller redoes the rows tap-exactly, aa_common.cuh)
  for (int rg = role.rg0; rg < nrg; rg += role.rg_par) {
    for (int cfb = role.cf0 - (role.cf0 & 31); cfb < npc; cfb += role.cf_step) {
      const int pc = cfb + (role.cf0 & 31);
      const bool act = pc < npc;
      const int4 pi = pinfo[act ? pc : 0];
      const int len = act ? (pi.z & 0xffff) : 1;
      const int lenm = __reduce_max_sync(0xffffffffu, len);
      const float2* wr = Wp + pi.y;
      constexpr int VP = vs_pitch(VW, PAD);
      const float* vrow = Vs + (rg * RPT) * VP;   // warp-uniform
      const float* vp = vrow + pi.x;
      uint32_t fb = 4u * (uint32_t)pi.x;           // PAD: byte offset of the tap's column, before padding
      float2 h[RPT];
#pragma unroll
      for (int r = 0; r < RPT; r++) h[r] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int j = 0; j < lenm; j++) {
        const float2 w2 = wr[j];
        if constexpr (PAD) vp = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vrow) + fb + ((fb >> 7) << padsh));
#pragma unroll
        for (int r = 0; r < RPT; r++) {
          const float v = vp[r * VP];
          h[r] = __ffma2_rn(make_float2(v, v), w2, h[r]);
        }
        if constexpr (PAD) fb += (j + 1 < len) ? 4u * Ci : 0u;
        else vp += (j + 1 < len) ? Ci : 0;
      }
      if (act) {
        const int64_t dst = op_off + (int64_t)(gbase + rg * RPT) * out_stride_h + pi.w;
        const bool hasb = ((pi.z >> 16) & 1) != 0;
        const int c = pi.z >> 20, cstep = epi.colstep(Ci);
#pragma unroll
        for (int r = 0; r < RPT; r++) {
          if (rg * RPT + r < cnt) {
            if constexpr (CHK) bad |= aa_nonfinite(h[r].x + h[r].y);
            aa_store<GEN>(op, dst + (int64_t)r * out_stride_h, h[r].x, c, epi);
            if (hasb) aa_store<GEN>(op, dst + (int64_t)r * out_stride_h + cstep, h[r].y, c, epi);
          }
        }
      }
    }
  }
  return bad;
}
// single-column form (P.pairs == 0): one item = one flat output column x RPT rows, FFMA2 over row pairs
template <int RPT, int VW, bool GEN, bool PAD, bool CHK>
__device__ __forceinline__ bool hphase_run_single(const float* __restrict__ Vs, const float* __restrict__ Ws,
                                                  const int4* __restrict__ pinfo, void* __restrict__ op, int64_t op_off,
                                                  const OutEpi& epi, int64_t out_stride_h, int Ci, int nof, const HRole role,
                                                  int gbase, int cnt, int padsh) {
  const int nrg = (cnt + RPT - 1) / RPT;
  bool bad = false;
  for (int rg = role.rg0; rg < nrg; rg += role.rg_par) {
    for (int cfb = role.cf0 - (role.cf0 & 31); cfb < nof; cfb += role.cf_step) {
      const int cf = cfb + (role.cf0 & 31);
      const bool act = cf < nof;
      const int4 ci = pinfo[act ? cf : 0];
      const int xs = act ? (ci.z & 0xfffff) : 1;
      const int xsm = __reduce_max_sync(0xffffffffu, xs);
      const float* wr = Ws + ci.y;
      constexpr int VP = vs_pitch(VW, PAD);
      const float* vrow = Vs + (rg * RPT) * VP;
      const float* vp = vrow + ci.x;
      uint32_t fb = 4u * (uint32_t)ci.x;
      float h[RPT];
#pragma unroll
      for (int r = 0; r < RPT; r++) h[r] = 0.f;
#pragma unroll 4
      for (int j = 0; j < xsm; j++) {
        const float wj = wr[j];
        const float2 w2 = make_float2(wj, wj);
        if constexpr (PAD) vp = reinterpret_cast<const float*>(reinterpret_cast<const char*>(vrow) + fb + ((fb >> 7) << padsh));
#pragma unroll
        for (int r = 0; r < RPT; r += 2) {
          const float2 q = __ffma2_rn(w2, make_float2(vp[r * VP], vp[(r + 1) * VP]), make_float2(h[r], h[r + 1]));
          h[r] = q.x;
          h[r + 1] = q.y;
        }
        if constexpr (PAD) fb += (j + 1 < xs) ? 4u * Ci : 0u;
        else vp += (j + 1 < xs) ? Ci : 0;
      }
      if (act) {
        const int64_t dst = op_off + (int64_t)(gbase + rg * RPT) * out_stride_h + ci.w;
        const int c = ci.z >> 20;
#pragma unroll
        for (int r = 0; r < RPT; r++)
          if (rg * RPT + r < cnt) {
            if constexpr (CHK) bad |= aa_nonfinite(h[r]);
            aa_store<GEN>(op, dst + (int64_t)r * out_stride_h, h[r], c, epi);
          }
      }
    }
  }
  return bad;
}
template <int RPT, int VW, bool GEN, bool PAD, bool CHK>
__device__ __forceinline__ bool hphase_run(const SParams& P, const float* Vs, const float2* Wp, const int4* pinfo, int64_t op_off,
                                           int npc, const HRole role, int gbase, int cnt) {
  const int padsh = P.pad == 4 ? 4 : P.pad == 2 ? 3 : 2;  // byte shift of (f >> 5) * q
  if (P.pairs) return hphase_run_pairs<RPT, VW, GEN, PAD, CHK>(Vs, Wp, pinfo, P.out, op_off, P.epi, P.lout.stride_h, P.Ci, npc, role, gbase, cnt, padsh);
  return hphase_run_single<RPT, VW, GEN, PAD, CHK>(Vs, reinterpret_cast<const float*>(Wp), pinfo, P.out, op_off, P.epi, P.lout.stride_h, P.Ci, npc, role, gbase, cnt, padsh);
}
// bytes of the strip tables (after Vs) for a plan
inline size_t strip_table_bytes(const SParams& P) {
  const size_t items = P.pairs ? (size_t)((P.strip_ox + 1) / 2) * P.Ci : (size_t)P.strip_ox * P.Ci;
  return (size_t)P.wtab_bytes + 16 + items * sizeof(int4);
}

// Host-side view of the tables for one direction of the banded separable apply (forward, or adjoint when
// the backward is the downsampling-shaped direction).  "Output rows" are the rows the kernel produces.
struct StreamTables {
  uint64_t key_h, key_w;                  // identity for the plan cache (AxisTables::id)
  int dir;                                // 0 forward, 1 adjoint
  const int32_t *hh_start, *hh_size;      // host: per output row, first input row and window length
  int64_t n_out_h;
  const int32_t *hw_start, *hw_size;      // host: per output column, first input column and window length
  int64_t n_in_w, n_out_w;
};

// Strip / row-buffer plan shared by both variants (exact, from the host mirrors of the tables).
//   cap   flat elements one strip may span (threads * VEC)
//   aln   alignment of a strip's first element (VEC, or 16 bytes' worth for the TMA variant)
//   U     input rows processed between two checks of the row buffer
inline int plan_stream(SParams& P, const StreamTables& T, int cap, int aln, int vec, int U, int tg) {
  const int64_t oW = P.oW;
  const int Ci = P.Ci;
  int n_strips = 1, strip_ox = (int)oW;
  int64_t max_extent = 0;
  for (;; n_strips++) {
    if (n_strips > oW) return fail(AA_ERR_UNSUPPORTED, "stream: a single output column spans more than one strip");
    strip_ox = (int)((oW + n_strips - 1) / n_strips);
    bool ok = strip_ox <= 512;
    max_extent = 0;
    for (int64_t a = 0; ok && a < oW; a += strip_ox) {
      const int64_t b = std::min<int64_t>(oW, a + strip_ox) - 1;
      const int64_t f0 = ((int64_t)T.hw_start[a] * Ci) & ~(int64_t)(aln - 1);
      const int64_t f1 = ((int64_t)T.hw_start[b] + T.hw_size[b]) * Ci;
      if (f1 - f0 > cap) ok = false;
      max_extent = std::max(max_extent, f1 - f0);
    }
    if (ok) break;
  }
  n_strips = (int)((oW + strip_ox - 1) / strip_ox);
  P.n_strips = n_strips;
  P.strip_ox = strip_ox;
  P.aln = aln;
  (void)vec;
  P.vw = cap;  // compile-time row pitch of Vs (lets the horizontal phase use immediate row offsets)
  int fmax = 1;
  {
    const int64_t oH = P.oH;
    int64_t lo = 0;
    for (int64_t o = 0; o < oH; o++) {  // ends are non-decreasing
      const int64_t e = (int64_t)T.hh_start[o] + T.hh_size[o];
      while ((int64_t)T.hh_start[lo] + T.hh_size[lo] <= e - U) lo++;
      fmax = std::max<int>(fmax, (int)(o - lo + 1));
    }
  }
  int shift = 0;
  for (int64_t a = 0; a < oW; a += strip_ox)
    for (int64_t o = a; o + 1 < std::min<int64_t>(oW, a + strip_ox); o += 2)
      shift = std::max<int>(shift, T.hw_start[o + 1] - T.hw_start[o]);
  P.pairs = (int64_t)strip_ox * Ci >= 256 ? 1 : 0;  // enough independent items per row group to halve them
  P.kp = (P.pairs ? P.Kw + shift : P.Kw) | 1;
  P.wtab_bytes = P.pairs ? (int)((size_t)((strip_ox + 1) / 2) * P.kp * sizeof(float2)) : (int)((size_t)strip_ox * P.kp * sizeof(float));
  if (P.kp >= (1 << 16) || (int64_t)((strip_ox + 1) / 2) * P.kp >= (1ll << 30)) return fail(AA_ERR_UNSUPPORTED, "stream: window too long");
  // Bank picture of the gather: the first 32 items of the first strip read, at tap 0, the columns below.
  // Pad the rows when that would at least halve a >= 4-way conflict.
  {
    auto worst = [&](int q) {
      int cnt[32] = {0}, w = 0;
      const int64_t f0 = ((int64_t)T.hw_start[0] * Ci) & ~(int64_t)(aln - 1);
      const int nitem = P.pairs ? (strip_ox + 1) / 2 * Ci : strip_ox * Ci;
      for (int l = 0; l < 32 && l < nitem; l++) {
        const int o = P.pairs ? 2 * (l / Ci) : l / Ci, c = l % Ci;
        if (o >= oW) break;
        const int64_t f = (int64_t)T.hw_start[o] * Ci + c - f0;
        w = std::max(w, ++cnt[(f + (f >> 5) * q) & 31]);
      }
      return w;
    };
    const int w0 = worst(0);
    int best_q = 0, best_w = w0;
    for (int q : {1, 2, 4}) {
      const int w = worst(q);
      if (w < best_w) { best_w = w; best_q = q; }
    }
    P.pad = (w0 >= 4 && 2 * best_w <= w0) ? best_q : 0;
  }
  P.tg = tg;
  P.vr = (P.tg - 1 + fmax + 3) / 4 * 4;
  if (P.vr > 32) return fail(AA_ERR_UNSUPPORTED, "stream: too many output rows finish per input batch (upsampling in H)");
  P.total_units = P.lin.planes * n_strips * P.oH;
  return AA_OK;
}

// Launch plans are cached per (tables, interleave, kernel variant): the steady-state host cost of a
// call is one map lookup + one kernel launch (no occupancy queries, no attribute calls).
struct PlanKey {
  uint64_t th, tw;
  int Ci, kid;
  bool operator<(const PlanKey& o) const {
    if (th != o.th) return th < o.th;
    if (tw != o.tw) return tw < o.tw;
    if (Ci != o.Ci) return Ci < o.Ci;
    return kid < o.kid;
  }
};
struct Plan {
  int n_strips, strip_ox, vw, vr, tg, aln, in_pitch, kp, pairs, wtab_bytes, pad;
  size_t smem;
  int max_grid;  // SMs * resident CTAs per SM
};
bool plan_lookup(const PlanKey& k, Plan* p);
void plan_store(const PlanKey& k, const Plan& p);
void plan_clear();
inline void plan_apply(SParams& P, const Plan& pl) {
  P.n_strips = pl.n_strips; P.strip_ox = pl.strip_ox; P.vw = pl.vw; P.vr = pl.vr; P.tg = pl.tg; P.aln = pl.aln;
  P.pad = pl.pad;
  P.in_pitch = pl.in_pitch;
  P.kp = pl.kp;
  P.pairs = pl.pairs;
  P.wtab_bytes = pl.wtab_bytes;
  P.total_units = P.lin.planes * pl.n_strips * P.oH;
}
inline Plan plan_from(const SParams& P, size_t smem, int max_grid) {
  return Plan{P.n_strips, P.strip_ox, P.vw, P.vr, P.tg, P.aln, P.in_pitch, P.kp, P.pairs, P.wtab_bytes, P.pad, smem, max_grid};
}

}  // namespace stream_detail

// TMA variant launcher (aa_stream_tma.cu); AA_ERR_UNSUPPORTED when rows are not 16-byte aligned.
int launch_stream_tma(stream_detail::SParams& P, int A, int in_dtype, const stream_detail::StreamTables& T, int device,
                      cudaStream_t stream);

}  // namespace aa
