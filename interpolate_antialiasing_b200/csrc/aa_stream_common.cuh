// aa_stream_common.cuh -- pieces shared by the two variants of the streaming forward kernel
// (aa_stream.cu: plain 128-bit global loads; aa_stream_tma.cu: cp.async.bulk row staging).
#pragma once
#include <algorithm>

#include "aa_common.cuh"

namespace aa {
namespace stream_detail {

constexpr int kMaxA = 6;

struct SParams {
  const void* in;
  float* out;
  Layout lin, lout;
  int Ci;
  int64_t H, oH, oW;
  const float* slot_h;  // [H][RS]
  int RS;
  const int32_t *xmin_h, *xsize_h;
  const int32_t *xmin_w, *xsize_w;
  const float* w_w;  // [oW][Kw]
  int Kw;
  int n_strips, strip_ox;
  int64_t total_units;  // planes * n_strips * oH
  int vw;               // shared-memory row pitch of Vs in floats
  int vr;               // rows of Vs
  int tg;               // buffered rows that trigger a horizontal phase
  int aln;              // alignment (elements) of a strip's first flat element
  int in_pitch;         // TMA variant: bytes between staged input rows in shared memory
};

// ---- vector loads (read-once data: bypass L1 allocation) -------------------------------------
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
__device__ __forceinline__ void unpack4(uint32_t w, float* v) {
  v[0] = (float)(w & 0xffu);
  v[1] = (float)((w >> 8) & 0xffu);
  v[2] = (float)((w >> 16) & 0xffu);
  v[3] = (float)(w >> 24);
}
template <> struct VLoad<uint8_t, 16> {
  static __device__ __forceinline__ void ld(const uint8_t* p, float (&v)[16]) {
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p));
    unpack4(a, v); unpack4(b, v + 4); unpack4(c, v + 8); unpack4(d, v + 12);
  }
};
template <> struct VLoad<uint8_t, 8> {
  static __device__ __forceinline__ void ld(const uint8_t* p, float (&v)[8]) {
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
    unpack4(a, v); unpack4(b, v + 4);
  }
};
template <> struct VLoad<uint8_t, 4> {
  static __device__ __forceinline__ void ld(const uint8_t* p, float (&v)[4]) {
    uint32_t a;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
    unpack4(a, v);
  }
};

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
// Per strip, every flat output column gets a packed descriptor in shared memory so the phase needs no
// integer division: x = first tap's offset inside a Vs row, y = (offset of the column's weights in
// Ws) | (window length << 20).
struct HRole {        // which (row group, column) items a thread owns; fixed per strip
  int rg0, rg_par;    // first row group and stride over row groups
  int cf0, cf_step;   // first flat column and stride over columns
};
__device__ __forceinline__ HRole hphase_role(int t, int nthreads, int nof) {
  HRole r;
  const int nw = nthreads >> 5, warp = t >> 5, lane = t & 31;
  const int wpr = (nof + 31) >> 5;  // warps needed to cover one row group
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
__device__ __forceinline__ void hphase_build_colinfo(int2* colinfo, const int* sxmin, const int* sxsize, int t, int nthreads,
                                                     int nof, int Ci, int Kw, int fl0) {
  for (int cf = t; cf < nof; cf += nthreads) {
    const int oxl = cf / Ci;
    const int c = cf - oxl * Ci;
    colinfo[cf] = make_int2(sxmin[oxl] * Ci + c - fl0, (oxl * Kw) | (sxsize[oxl] << 20));
  }
}
// Gather over the buffered rows [0, cnt) of Vs -> output rows gbase..gbase+cnt-1.  Trip counts are
// warp-uniform (max window length in the warp) with per-lane predication, so there is no divergence and
// no tap outside a column's true window is ever read.
template <int RPT, int VW>
__device__ __forceinline__ void hphase_run(const float* __restrict__ Vs, const float* __restrict__ Ws,
                                           const int2* __restrict__ colinfo, float* __restrict__ op, int64_t out_stride_h,
                                           int Ci, int nof, const HRole role, int gbase, int cnt) {
  const int nrg = (cnt + RPT - 1) / RPT;
  for (int rg = role.rg0; rg < nrg; rg += role.rg_par) {
    for (int cfb = role.cf0 - (role.cf0 & 31); cfb < nof; cfb += role.cf_step) {
      const int cf = cfb + (role.cf0 & 31);
      const bool act = cf < nof;
      const int2 ci = colinfo[act ? cf : 0];
      const int xs = act ? (ci.y >> 20) : 1;
      const int xsm = __reduce_max_sync(0xffffffffu, xs);
      const float* wr = Ws + (ci.y & 0xfffff);
      const float* vp = Vs + (rg * RPT) * VW + ci.x;
      float h[RPT];
#pragma unroll
      for (int r = 0; r < RPT; r++) h[r] = 0.f;
      // Warp-uniform trip count.  Past a lane's own window the weight read is the table's zero padding
      // (aa_interpolation_impl.h:276-278) and the data pointer stops advancing, so no element outside
      // the true window is touched.
#pragma unroll 4
      for (int j = 0; j < xsm; j++) {
        const float wj = wr[j];
#pragma unroll
        for (int r = 0; r < RPT; r++) h[r] = fmaf(wj, vp[r * VW], h[r]);
        vp += (j + 1 < xs) ? Ci : 0;
      }
      if (act) {
        float* dst = op + (int64_t)(gbase + rg * RPT) * out_stride_h + cf;
#pragma unroll
        for (int r = 0; r < RPT; r++)
          if (rg * RPT + r < cnt) dst[(int64_t)r * out_stride_h] = h[r];
      }
    }
  }
}

// Strip / row-buffer plan shared by both variants (exact, from the host mirrors of the tables).
//   cap   flat elements one strip may span (threads * VEC)
//   aln   alignment of a strip's first element (VEC, or 16 bytes' worth for the TMA variant)
//   U     input rows processed between two checks of the row buffer
inline int plan_stream(SParams& P, const AxisTables* th, const AxisTables* tw, int cap, int aln, int vec, int U, int tg) {
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
      const int64_t f0 = ((int64_t)tw->h_xmin[a] * Ci) & ~(int64_t)(aln - 1);
      const int64_t f1 = ((int64_t)tw->h_xmin[b] + tw->h_xsize[b]) * Ci;
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
      const int64_t e = (int64_t)th->h_xmin[o] + th->h_xsize[o];
      while ((int64_t)th->h_xmin[lo] + th->h_xsize[lo] <= e - U) lo++;
      fmax = std::max<int>(fmax, (int)(o - lo + 1));
    }
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
  const void *th, *tw;
  int Ci, kid;
  bool operator<(const PlanKey& o) const {
    if (th != o.th) return th < o.th;
    if (tw != o.tw) return tw < o.tw;
    if (Ci != o.Ci) return Ci < o.Ci;
    return kid < o.kid;
  }
};
struct Plan {
  int n_strips, strip_ox, vw, vr, tg, aln, in_pitch;
  size_t smem;
  int max_grid;  // SMs * resident CTAs per SM
};
bool plan_lookup(const PlanKey& k, Plan* p);
void plan_store(const PlanKey& k, const Plan& p);
void plan_clear();
inline void plan_apply(SParams& P, const Plan& pl) {
  P.n_strips = pl.n_strips; P.strip_ox = pl.strip_ox; P.vw = pl.vw; P.vr = pl.vr; P.tg = pl.tg; P.aln = pl.aln;
  P.in_pitch = pl.in_pitch;
  P.total_units = P.lin.planes * pl.n_strips * P.oH;
}
inline Plan plan_from(const SParams& P, size_t smem, int max_grid) {
  return Plan{P.n_strips, P.strip_ox, P.vw, P.vr, P.tg, P.aln, P.in_pitch, smem, max_grid};
}

}  // namespace stream_detail

// TMA variant launcher (aa_stream_tma.cu); AA_ERR_UNSUPPORTED when rows are not 16-byte aligned.
int launch_stream_tma(stream_detail::SParams& P, int A, int in_dtype, const AxisTables* th, const AxisTables* tw, int device,
                      cudaStream_t stream);

}  // namespace aa
