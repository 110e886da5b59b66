// aa_stream.cu -- K2: the fused streaming forward kernel (sm_100a), the bandwidth-bound hot path.
//
// Replaces the two TensorIterator passes + temp tensor of
// ti_separable_upsample_generic_Nd_kernel_impl
// (/root/reference/step_two_dot_two/aa_interpolation_impl.h:628-683) with ONE kernel that reads
// every input element from HBM once and writes every output element once.
//
// Shape of the computation (see DESIGN.md section 4 for the why):
//   * A plane row is a flat array of W*Ci elements (Ci = C for channels_last, 1 for channels_first),
//     so the VERTICAL filter is layout-agnostic: thread t owns VEC consecutive flat elements and
//     streams down the rows with 128-bit coalesced loads.  Each input row contributes to at most A
//     output rows (A = 3 bilinear, 5 bicubic when downsampling), so the thread keeps A
//     accumulators per element, ordered by age: acc[k] collects the k-th oldest open output row.
//     The per-row record {wT[y][0..A), first_flush_oy | nflush<<24} is warp-uniform (one 16/32-byte
//     broadcast load).
//     The FMAs are packed FFMA2 (two fp32 FMAs per issue slot, scalar weight broadcast).
//   * When an output row's window ends, its accumulator (one vertically-filtered row, still at full
//     input width) is flushed to shared memory.  Once >= tg rows are buffered the CTA runs the
//     HORIZONTAL filter as a gather over the shared-memory rows (aa_stream_common.cuh: per-strip weight
//     tables in shared memory, pairs of adjacent output columns x 4 rows per thread for wide strips)
//     and writes the output rows with coalesced stores (optionally clamped/rounded to uint8).
//   * Work is split stream-K style: the (plane, column strip, output row) space is cut into
//     gridDim.x equal contiguous ranges; a CTA that starts mid-plane re-reads only the
//     <= 2*support_h halo rows of its first window.
//
// Order of the passes is V then H (the reference is H then V).  Both orders evaluate the same
// separable sum with fp32 rounding of the intermediate; the difference is pure rounding
// (measured <= 1.3e-4 abs / 2.3e-6 rel on a 0..255 scale against the reference, tolerance
// 1e-3 abs / 1e-5 rel -- tests/test_forward_gpu.py).  The bit-exact order lives in aa_general.cu.
#include <stdlib.h>

#include <map>
#include <mutex>

#include "aa_stream_common.cuh"

namespace aa {
using namespace stream_detail;
namespace {

// A     accumulators per element, ordered by age (>= max output rows one input row contributes to)
// VEC   flat elements per thread per row
// NT    threads per CTA
// U     input rows in flight per thread
// MINB  CTAs per SM the register allocation must allow
// Shared memory holds up to P.vr vertically-filtered rows; the horizontal phase runs at the end of a
// U-row batch once at least P.tg rows are buffered (host guarantees tg - 1 + max flushes per batch <= vr).
template <int A, int VEC, typename in_t, int NT, int U, int MINB, bool GEN, bool PAD, bool PF = true>
__global__ void __launch_bounds__(NT, MINB) aa_stream_kernel(const SParams P) {
  extern __shared__ __align__(16) float smem[];
  if constexpr (sizeof(in_t) == 4) aa_trigger_drain();
  constexpr int RPT = 4;
  constexpr int VW = NT * VEC;  // row pitch of Vs in floats (== P.vw)  // rows per thread in the horizontal phase
  constexpr int RS4 = (A + 1 + 3) / 4;  // float4 per row record
  using RawT = typename Raw<in_t, VEC>::T;
  constexpr int RN = Raw<in_t, VEC>::N;
  float* Vs = smem;                                  // [vr][vw]
  float2* Wp = reinterpret_cast<float2*>(Vs + (size_t)P.vr * vs_pitch(P.vw, PAD));                       // [pairs][kp]
  int4* pinfo = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(Wp) + P.wtab_bytes + 15) & ~(uintptr_t)15);      // [pairs * Ci]

  const int t = threadIdx.x;
  const int Ci = P.Ci;
  const int64_t oH = P.oH;
  const int64_t stride_h = P.lin.stride_h;
  constexpr int vw = VW;
  const int64_t u_begin = P.total_units * (int64_t)blockIdx.x / gridDim.x;
  const int64_t u_end = P.total_units * (int64_t)(blockIdx.x + 1) / gridDim.x;
  int cur_strip = -1, strip_fl0 = 0, strip_npc = 0;
  HRole role = {0, 1, 0, 1};
  constexpr bool CHK = sizeof(in_t) == 4;  // float input can carry NaN/Inf (aa_common.cuh: aa_exact_region)
  __shared__ int s_bad;                    // some thread stored a non-finite value (in shared memory, not a register: the
  if (t == 0) s_bad = 0;                   // 64-register shapes have none to spare; ordered by the strip-setup barrier)

  for (int64_t u = u_begin; u < u_end;) {
    // ---- segment = run of output rows [oyA, oyB) inside one (plane, strip) column
    const int64_t col = u / oH;
    const int oyA = (int)(u - col * oH);
    const int64_t seg_end = min(u_end, (col + 1) * oH);
    const int oyB = oyA + (int)(seg_end - u);
    const int64_t plane = col / P.n_strips;
    const int s = (int)(col - plane * P.n_strips);
    const int ox0 = s * P.strip_ox;
    const int ox1 = min((int)P.oW, ox0 + P.strip_ox);
    if (s != cur_strip) {
      __syncthreads();
      strip_setup(P, t, NT, ox0, ox1, Wp, pinfo, &strip_fl0, &strip_npc);
      role = hphase_role(t, NT, strip_npc);
      cur_strip = s;
      __syncthreads();
    }
    const int fl0 = strip_fl0;                                                              // first flat element of the strip
    const int fl_end = (__ldg(P.xmin_w + ox1 - 1) + __ldg(P.xsize_w + ox1 - 1)) * Ci;       // one past the last
    const bool valid = fl0 + VEC * t < fl_end;
    // threads beyond the strip re-read its first vector (a legal address) and never store
    const int fmy = valid ? fl0 + VEC * t : fl0;
    const int64_t yA = __ldg(P.xmin_h + oyA);
    const int64_t yB = (int64_t)__ldg(P.xmin_h + oyB - 1) + __ldg(P.xsize_h + oyB - 1);
    const in_t* ip = (const in_t*)P.in + (plane / P.lin.Cp) * P.lin.stride_n + (plane % P.lin.Cp) * P.lin.stride_p + fmy + yA * stride_h;
    const float4* rp = reinterpret_cast<const float4*>(P.slot_h) + yA * RS4;
    const int64_t op = (plane / P.lout.Cp) * P.lout.stride_n + (plane % P.lout.Cp) * P.lout.stride_p;  // element offset of the plane
    float* vdst = Vs + vs_pos(VEC * t, PAD ? P.pad : 0);  // this thread's columns in Vs row 0
    float* vptr = vdst;          // where the next finished row goes
    const bool vstore = valid;   // (kept in a predicate-friendly local)

    float acc[A][VEC];
#pragma unroll
    for (int a = 0; a < A; a++)
#pragma unroll
      for (int i = 0; i < VEC; i++) acc[a][i] = 0.f;
    int gbase = oyA;  // output row held in Vs[0]
    int cnt = 0;      // rows buffered in Vs

    // one input row: A FMAs per element with warp-uniform weights; acc[k] belongs to the k-th oldest
    // open output row.  When rows finish (rarely: once per scale_h rows) the oldest accumulators are
    // stored to shared memory and the rest shift down.
    auto row = [&](const RawT (&raw)[RN], const float4 (&rq)[RS4]) {
      const float* rw = reinterpret_cast<const float*>(rq);
      float v[VEC];
      expand<VEC>(raw, v);
      vfma<A, VEC>(acc, v, rw);
      const int packed = __float_as_int(rw[A]);
      if (packed >> 24) {
        int nfl = packed >> 24;
        int o = packed & 0xffffff;
        // lean retire loop (it runs once per scale_h input rows: at small scales it rivals the FMAs):
        // a running shared-memory pointer instead of cnt*pitch, the strip predicate kept in `vstore`
#pragma unroll 1
        do {
          if (o >= oyA && o < oyB) {
            if (vstore) {
              if constexpr (PAD) {  // padded rows: the thread's columns stay contiguous but lose their 16-byte alignment
#pragma unroll
                for (int e = 0; e < VEC; e++) vptr[e] = acc[0][e];
              } else {
                store_vec<VEC>(vptr, acc[0]);
              }
            }
            vptr += vs_pitch(VW, PAD);
            cnt++;
          }
#pragma unroll
          for (int a = 0; a + 1 < A; a++)
#pragma unroll
            for (int e = 0; e < VEC; e++) acc[a][e] = acc[a + 1][e];
#pragma unroll
          for (int e = 0; e < VEC; e++) acc[A - 1][e] = 0.f;
          o++;
        } while (--nfl);
      }
    };
    // horizontal filter over the buffered rows [gbase, gbase+cnt)
    auto hphase = [&]() {
      __syncthreads();
      if (hphase_run<RPT, VW, GEN, PAD, CHK>(P, Vs, Wp, pinfo, op, strip_npc, role, gbase, cnt)) s_bad = 1;
      __syncthreads();
      gbase += cnt;
      cnt = 0;
      vptr = vdst;
    };

    int64_t y = yA;
    for (; y + U <= yB; y += U) {
      RawT v[U][RN];
      float4 rq[U][RS4];
#pragma unroll
      for (int i = 0; i < U; i++) VLoad<in_t, VEC>::ld(ip + i * stride_h, v[i]);
#pragma unroll
      for (int i = 0; i < U; i++)
#pragma unroll
        for (int q = 0; q < RS4; q++) rq[i][q] = __ldg(rp + i * RS4 + q);
      ip += U * stride_h;
      rp += U * RS4;
      // L2 prefetch of a later batch: keeps HBM busy across the horizontal phase, when no load of this CTA is in
      // flight -- a hint, no registers (measured +3..7 points of HBM peak wherever the horizontal phase is a large
      // share: scales 0.125x-0.5x, bicubic, uint8, backward of upsampling).  Where the hints are issued is chosen
      // per shape by what ptxas then does with the real loads (scripts/sass_lint.py): for the 3-accumulator
      // shapes at the END of the batch (the batch after the next one), so that they do not queue ahead of this
      // batch's loads; for the others next to the loads (the next batch).
      constexpr bool PF_AT_END = A * VEC <= 12;
      if (PF && !PF_AT_END && y + 2 * U <= yB) {
#pragma unroll
        for (int i = 0; i < U; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ip + i * stride_h));
      }
#pragma unroll
      for (int i = 0; i < U; i++) row(v[i], rq[i]);
      if (PF && PF_AT_END && y + 3 * U <= yB) {
#pragma unroll
        for (int i = 0; i < U; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ip + (U + i) * stride_h));
      }
      if (cnt >= P.tg) hphase();
    }
    for (; y < yB; y++) {
      RawT v[RN];
      float4 rq[RS4];
      VLoad<in_t, VEC>::ld(ip, v);
#pragma unroll
      for (int q = 0; q < RS4; q++) rq[q] = __ldg(rp + q);
      ip += stride_h;
      rp += RS4;
      row(v, rq);
    }
    if (cnt > 0) hphase();
    u = seg_end;
  }
  if constexpr (CHK) {  // (every write of s_bad is followed by a barrier of its horizontal phase; the unit range is
    if (threadIdx.x == 0 && s_bad && P.redo)  // recomputed so that nothing of it stays live through the loop)
      redo_push(P.redo, P.total_units * (int64_t)blockIdx.x / gridDim.x, P.total_units * (int64_t)(blockIdx.x + 1) / gridDim.x, 0, 0, 0, 0);
  }
}

template <int A, int VEC, typename in_t, int NT_, bool PAD = false>
struct Cfg {
  static constexpr int NT = NT_;  // 256, or 128 when a whole row fits 128 threads (narrow images)
  // rows in flight per thread.  uint8 rows are 4x fewer bytes per element, and the wide-tap uint8 shapes are
  // limited to 2 CTAs/SM by shared memory anyway: they (A >= 4) take 8 rows and the 128-register budget that leaves.
  static constexpr bool U8W = sizeof(in_t) == 1 && VEC == 8 && A >= 4;
  static constexpr int U = U8W ? 8 : 4;
  static constexpr int TG = 4;  // buffered rows that trigger a horizontal phase
  // register budget: 64/thread when the accumulators are few (4 CTAs/SM at 256 threads), else fewer CTAs
  // (scripts/sass_lint.py checks that ptxas still issues all U row loads before the first FMA under each budget)
  static constexpr int MINB = (U8W ? 2 : (A * VEC <= 12 && !PAD && A < 6) ? 4 : ((A * VEC <= 32 && (A < 6 || VEC < 4)) ? 3 : 2)) * (256 / NT_);
};

template <int A, int VEC, typename in_t, int NT_ = 256, bool GEN = false>
int launch_cfg(SParams& P, const StreamTables& T, int device, cudaStream_t stream) {
  using C = Cfg<A, VEC, in_t, NT_>;
  if constexpr (!GEN && NT_ == 256 && (VEC == 4 || (VEC == 8 && sizeof(in_t) == 1))) {
    // decode-adjacent epilogue (normalise / half / planar): separate instantiations of the common shapes
    if (P.epi.generic()) return launch_cfg<A, VEC, in_t, NT_, true>(P, T, device, stream);
  } else if constexpr (!GEN) {
    if (P.epi.generic()) return fail(AA_ERR_UNSUPPORTED, "stream: generic epilogue not instantiated for this shape");
  }
  // PAD (bank-conflict-free padded row buffer) exists for the wide-vector shapes; plan_stream decides per table
  constexpr bool PADDABLE = !GEN && NT_ == 256 && ((VEC == 4 && sizeof(in_t) == 4) || (VEC == 8 && sizeof(in_t) == 1));
  auto kern = aa_stream_kernel<A, VEC, in_t, C::NT, C::U, C::MINB, GEN, false>;
  const PlanKey key{T.key_h, T.key_w, P.Ci, (T.dir << 29) | (GEN ? (1 << 28) : 0) | (NT_ << 16) | (A << 8) | (VEC << 2) | (int)sizeof(in_t) % 4};
  Plan pl;
  if (!plan_lookup(key, &pl)) {
    P.in_pitch = 0;
    int rc = plan_stream(P, T, C::NT * VEC, VEC, VEC, C::U, C::TG);
    if (rc != AA_OK) return rc;
    if (!PADDABLE) P.pad = 0;
    const size_t smem_ = sizeof(float) * (size_t)P.vr * vs_pitch(P.vw, P.pad != 0) + strip_table_bytes(P);
    if (smem_ > 200 * 1024) return fail(AA_ERR_UNSUPPORTED, "stream: shared memory plan too large");
    int occ = 0, sms = 0;
    if constexpr (PADDABLE) {
      if (P.pad) {
        using CP = Cfg<A, VEC, in_t, NT_, true>;
        auto kp = aa_stream_kernel<A, VEC, in_t, CP::NT, CP::U, CP::MINB, GEN, true>;
        AA_CUDA_TRY(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        AA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kp, C::NT, smem_));
      }
    }
    if (!P.pad) {
      if constexpr (A == 3 && !GEN && NT_ == 256 && VEC == 4 && sizeof(in_t) == 4)  // the variant without L2 hints (see below)
        AA_CUDA_TRY(cudaFuncSetAttribute(aa_stream_kernel<A, VEC, in_t, C::NT, C::U, C::MINB, GEN, false, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      AA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, C::NT, smem_));
    }
    AA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (occ < 1) return fail(AA_ERR_UNSUPPORTED, "stream: kernel does not fit on an SM");
    pl = plan_from(P, smem_, occ * sms);
    plan_store(key, pl);
  }
  plan_apply(P, pl);
  const size_t smem = pl.smem;
  const int64_t min_units = 4;  // do not cut segments shorter than this many output rows
  // up to 2x more CTAs than fit at once: the hardware block scheduler then evens out SM-to-SM speed differences
  // at the tail (measured +1 % on cfg2 in two same-box A/B runs; 3x was not reproducible; each extra cut costs one window's halo rows)
  // but never cut work finer than about one column (oH output rows) per CTA
  static const int gmul = [] { const char* e = getenv("AA_STREAM_GRID_MUL"); return e ? std::max(1, atoi(e)) : 2; }();  // tuning knob
  // ... and always whole waves: 1.5 waves (cfg2 at 128 images: 896 columns over 592 resident CTAs) left every other SM idle
  // through the second half of the kernel, 0.69 instead of 0.88 of peak -- the shard size of 2-GPU strong scaling
  const int64_t waves = std::max<int64_t>(1, std::min<int64_t>(gmul, (P.total_units / P.oH) / std::max(1, pl.max_grid)));
  const int64_t want = (int64_t)pl.max_grid * waves;
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(want, P.total_units / min_units));
  if constexpr (PADDABLE) {
    if (P.pad) {
      using CP = Cfg<A, VEC, in_t, NT_, true>;
      aa_stream_kernel<A, VEC, in_t, CP::NT, CP::U, CP::MINB, GEN, true><<<(unsigned)grid, C::NT, smem, stream>>>(P);
      AA_LAUNCH_CHECK("aa_stream_kernel");
      return AA_OK;
    }
  }
  // Long uninterrupted bilinear streams (3 accumulators, fewer than 1 output pixel per 32 input pixels, e.g. cfg2: the
  // horizontal phase is a few percent of the CTA's time) run the variant without the L2 hints: there the hints buy
  // little and the kernel with them measured 2 % slower on channels_last cfg2 (same-box A/B).  The bicubic shapes
  // gain from the hints at every scale.
  if constexpr (A == 3 && !GEN && NT_ == 256 && VEC == 4 && sizeof(in_t) == 4) {
    if (!P.pad && P.oH * P.oW * 32 < P.H * T.n_in_w) {
      aa_stream_kernel<A, VEC, in_t, C::NT, C::U, C::MINB, GEN, false, false><<<(unsigned)grid, C::NT, smem, stream>>>(P);
      AA_LAUNCH_CHECK("aa_stream_kernel");
      return AA_OK;
    }
  }
  kern<<<(unsigned)grid, C::NT, smem, stream>>>(P);
  AA_LAUNCH_CHECK("aa_stream_kernel");
  return AA_OK;
}

template <int A>
int launch_A(SParams& P, int in_dtype, int vec, const StreamTables& T, int device, cudaStream_t stream) {
  if (in_dtype == AA_F32) {
    // narrow rows: a 128-thread CTA covers the whole row, so no lanes idle through the vertical pass
    if (vec == 4 && T.n_in_w * P.Ci <= 128 * 4) return launch_cfg<A, 4, float, 128>(P, T, device, stream);
    if (vec == 4) return launch_cfg<A, 4, float>(P, T, device, stream);
    if (vec == 2) return launch_cfg<A, 2, float>(P, T, device, stream);
    return launch_cfg<A, 1, float>(P, T, device, stream);
  }
  if (vec == 8) return launch_cfg<A, 8, uint8_t>(P, T, device, stream);
  return launch_cfg<A, 4, uint8_t>(P, T, device, stream);
}

}  // namespace

namespace stream_detail {
namespace {
std::mutex g_plan_mu;
std::map<PlanKey, Plan> g_plans;
}  // namespace
bool plan_lookup(const PlanKey& k, Plan* p) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  auto it = g_plans.find(k);
  if (it == g_plans.end()) return false;
  *p = it->second;
  return true;
}
void plan_store(const PlanKey& k, const Plan& p) {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  if (g_plans.size() > 4096) g_plans.clear();  // plans of evicted tables: bounded, rebuilt on demand
  g_plans[k] = p;
}
void plan_clear() {
  std::lock_guard<std::mutex> lock(g_plan_mu);
  g_plans.clear();
}
}  // namespace stream_detail

namespace {
// widest vector the addresses allow: base pointer, plane strides and row stride must all be aligned
int pick_vec(const void* in, int in_dtype, const Layout& lin) {
  const int es = in_dtype == AA_F32 ? 4 : 1;
  auto aligned = [&](int vec) {
    const int64_t bytes = (int64_t)vec * es;
    if (((uintptr_t)in) % bytes) return false;
    if ((lin.stride_h % vec) || (lin.stride_n % vec) || (lin.Cp > 1 && lin.stride_p % vec)) return false;
    return true;
  };
  if (in_dtype == AA_F32) { for (int v : {4, 2, 1}) if (aligned(v)) return v; }
  else { for (int v : {8, 4}) if (aligned(v)) return v; }
  return 0;
}
int dispatch_A(SParams& P, int A, int in_dtype, int vec, const StreamTables& T, int device, uint32_t flags, cudaStream_t stream) {
  if (flags & AA_FLAG_STREAM_TMA) return launch_stream_tma(P, A, in_dtype, T, device, stream);
  switch (A) {
    case 3: return launch_A<3>(P, in_dtype, vec, T, device, stream);
    case 4: return launch_A<4>(P, in_dtype, vec, T, device, stream);
    case 5: return launch_A<5>(P, in_dtype, vec, T, device, stream);
    case 6: return launch_A<6>(P, in_dtype, vec, T, device, stream);
  }
  return fail(AA_ERR_UNSUPPORTED, "stream: unsupported accumulator count");
}
}  // namespace

int launch_stream(const void* in, int in_dtype, const Layout& lin, void* out, const Layout& lout,
                  AxisTables* th, AxisTables* tw, int64_t H, int64_t W, int64_t oH, int64_t oW,
                  uint32_t flags, OutEpi epi, cudaStream_t stream) {
  if (in_dtype != AA_F32 && in_dtype != AA_U8) return fail(AA_ERR_UNSUPPORTED, "stream: input must be f32 or u8");
  if (th->dtype != AA_F32 || tw->dtype != AA_F32) return fail(AA_ERR_UNSUPPORTED, "stream: f32 tables only");
  if (th->kt_max > kMaxA) return fail(AA_ERR_UNSUPPORTED, "stream: too many output rows per input row (upsampling in H)");
  if (oH >= (1 << 24) || W * lin.Ci >= (1ll << 30) || tw->K >= (1 << 16) || lin.Ci > 2047) return fail(AA_ERR_UNSUPPORTED, "stream: size limits");
  const int vec = pick_vec(in, in_dtype, lin);
  if (!vec) return fail(AA_ERR_UNSUPPORTED, "stream: input rows are not sufficiently aligned");
  const int A = th->kt_max <= 3 ? 3 : th->kt_max;
  int rc = ensure_slot_tables(th, A, stream);
  if (rc != AA_OK) return rc;

  SParams P;
  P.in = in; P.out = out; P.epi = epi; P.lin = lin; P.lout = lout; P.Ci = lin.Ci;
  P.H = H; P.oH = oH; P.oW = oW;
  P.slot_h = th->slot; P.RS = th->slot_RS;
  P.xmin_h = th->xmin; P.xsize_h = th->xsize;
  P.xmin_w = tw->xmin; P.xsize_w = tw->xsize; P.w_w = (const float*)tw->w; P.Kw = tw->K;
  const StreamTables T{th->id, tw->id, 0, th->h_xmin.data(), th->h_xsize.data(), oH, tw->h_xmin.data(), tw->h_xsize.data(), W, oW};
  P.redo = nullptr;
  if (in_dtype == AA_F32) {
    rc = redo_list(th->device, stream, &P.redo);
    if (rc != AA_OK) return rc;
  }
  rc = dispatch_A(P, A, in_dtype, vec, T, th->device, flags, stream);
  if (rc == AA_OK && in_dtype == AA_F32) {  // drain kernel: tap-exact redo of what stored a NaN/Inf (aa_redo.cu)
    const RedoParams R{in, out, epi, lin, lout, lin.Ci,
                       ExactTabs{P.xmin_h, P.xsize_h, P.xmin_w, P.xsize_w, (const float*)th->w, P.w_w, th->K, P.Kw},
                       (int)oH, (int)(oW * lin.Ci), P.n_strips, P.strip_ox, P.redo};
    rc = launch_redo(R, th->device, stream);
  }
  return rc;
}

int launch_stream_adjoint(const void* gout, const Layout& lo, void* gin, const Layout& li, AxisTables* th, AxisTables* tw,
                          cudaStream_t stream) {
  // grad_out [.., oH, oW] is streamed, grad_in [.., H, W] is produced: the roles of the two table sets swap.
  if (th->dtype != AA_F32 || tw->dtype != AA_F32) return fail(AA_ERR_UNSUPPORTED, "stream: f32 tables only");
  const int64_t H = th->in, oH = th->out, W = tw->in, oW = tw->out;
  if (th->xsize_max > kMaxA) return fail(AA_ERR_UNSUPPORTED, "stream/adjoint: too many grad_in rows per grad_out row");
  if (H >= (1 << 24) || oW * lo.Ci >= (1ll << 30) || tw->KT >= (1 << 16) || lo.Ci > 2047) return fail(AA_ERR_UNSUPPORTED, "stream/adjoint: size limits");
  // every grad_in row / column must be covered by at least one grad_out row / column (else it is never written)
  for (int64_t y = 0; y < H; y++) if (th->h_osize[y] < 1) return fail(AA_ERR_UNSUPPORTED, "stream/adjoint: uncovered rows");
  for (int64_t x = 0; x < W; x++) if (tw->h_osize[x] < 1) return fail(AA_ERR_UNSUPPORTED, "stream/adjoint: uncovered columns");
  const int vec = pick_vec(gout, AA_F32, lo);
  if (!vec) return fail(AA_ERR_UNSUPPORTED, "stream/adjoint: rows are not sufficiently aligned");
  const int A = th->xsize_max <= 3 ? 3 : th->xsize_max;
  int rc = ensure_slot_tables_adj(th, A, stream);
  if (rc != AA_OK) return rc;

  SParams P;
  P.in = gout; P.out = gin; P.epi = OutEpi(); P.lin = lo; P.lout = li; P.Ci = lo.Ci;
  P.H = oH; P.oH = H; P.oW = W;
  P.slot_h = th->slot_adj; P.RS = th->slot_adj_RS;
  P.xmin_h = th->omin; P.xsize_h = th->osize;
  P.xmin_w = tw->omin; P.xsize_w = tw->osize; P.w_w = (const float*)tw->wT; P.Kw = tw->KT;
  const StreamTables T{th->id, tw->id, 1, th->h_omin.data(), th->h_osize.data(), H, tw->h_omin.data(), tw->h_osize.data(), oW, W};
  rc = redo_list(th->device, stream, &P.redo);
  if (rc != AA_OK) return rc;
  rc = dispatch_A(P, A, AA_F32, vec, T, th->device, 0u, stream);
  if (rc == AA_OK) {
    const RedoParams R{gout, gin, P.epi, lo, li, lo.Ci,
                       ExactTabs{P.xmin_h, P.xsize_h, P.xmin_w, P.xsize_w, (const float*)th->wT, P.w_w, th->KT, P.Kw},
                       (int)H, (int)(W * lo.Ci), P.n_strips, P.strip_ox, P.redo};
    rc = launch_redo(R, th->device, stream);
  }
  return rc;
}

}  // namespace aa
