// aa_vmma.cu -- K4: fused forward for uint8 inputs with the VERTICAL pass on the 5th-generation tensor
// cores (tcgen05.mma kind::i8, accumulators in TMEM) and input tiles staged by TMA tensor copies.
//
// Same contract as the streaming kernel (aa_stream.cu): one launch replaces the two TensorIterator
// passes + temp tensor of ti_separable_upsample_generic_Nd_kernel_impl
// (/root/reference/step_two_dot_two/aa_interpolation_impl.h:628-683); every input byte is read from HBM
// once, every output element written once.  Why a different engine for uint8: with 1-byte pixels the
// FP32 pipes need >= 6 issue slots per input byte (unpack + A FMAs per element) while
// HBM delivers a byte every ~0.11 issue slots, so the streaming kernel tops out near 0.3-0.6 of HBM peak on
// wide-tap shapes (cfg3).  Here the per-byte work of the vertical pass costs NO issue slots:
//
//   * the vertical filter of a block of OYB = 32 output rows is a banded matrix: out_v[oy, x] =
//     sum_y Wh[oy, y] * in[y, x] with y in a window of <= 32*KSTEPS input rows.  Written as a GEMM
//     D[x, (limb, oy)] = sum_y A[x, y] * B[y, (limb, oy)] the data operand A is the raw uint8 image tile
//     (128 flat columns x 32 rows per MMA, M-major = exactly how TMA lands it in shared memory with the
//     128-byte swizzle) and B holds the fp32 weights of :194-281 quantised to 24-bit fixed point and
//     split into three balanced base-256 digits (int8 "limbs"; aa_tables.cu).  uint8 x int8 products
//     accumulate exactly in int32 in TMEM, so the vertical pass is exact up to the 2^-s weight
//     quantisation (s = 22..30: <= 7e-5 on a 0..255 scale in the worst case, far inside the
//     1e-3 + 1e-5*|x| contract and below the fp32 rounding of the reference's own pass);
//   * four epilogue groups of four warps read the accumulators (tcgen05.ld), rebuild fp32 from the three limbs with
//     three integer adds + three FFMA2 per two values (magic-number int->float, no I2F), and store the
//     vertically filtered rows TRANSPOSED ([flat column][32 output rows]) in shared memory;
//   * the HORIZONTAL pass is the same pair-of-columns gather as aa_stream_common.cuh, but one 128-bit
//     LDS now feeds four output rows (4x fewer shared-memory loads than the row-major buffer).
//
// Roles (576 threads, one CTA per SM, persistent over a contiguous range of work items):
//   warp 0     TMA producer: cp.async.bulk.tensor (4-D map: flat x, y, channel plane, image), one box = the whole K span
//              of a 128-column tile, into a ring of stages + one bulk copy of the weight limbs whenever the row block
//              changes; mbarrier complete_tx
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma (M=128, N=96, K=32) per 32 rows of a stage and
//              tcgen05.commit to release stages / publish accumulators (a ring of 5 accumulator buffers in TMEM)
//   warps 2-17 epilogue + horizontal pass: 4 independent groups of 4 warps (one per TMEM lane quarter), 8 rows each
// Work item = (column strip of <= 512 flat input columns, plane, block of 32 output rows), numbered in that order with
// the row block fastest: consecutive items of a CTA walk DOWN one strip of one plane (part of the 2*support halo rows an
// item shares with its predecessor are L2 hits), then take the same strip of the next plane, so a CTA's whole range
// lies in one or two strips and the strip's horizontal tables are set up once or twice per CTA.
#include <cuda.h>
#include <stdlib.h>

#include <map>
#include <mutex>

#include "aa_stream_common.cuh"

namespace aa {
using namespace stream_detail;
namespace {

constexpr int OYB = 32;                      // output rows per item
constexpr int NLIMB = 3;                     // int8 digits per weight
constexpr int UMMA_N = OYB * NLIMB;          // 96 accumulator columns per tile
constexpr int TILE_M = 128;                  // flat input columns per tile (= TMEM lanes)
constexpr int KSTEP = 32;                    // input rows per MMA (K of an 8-bit tcgen05.mma)
constexpr int STAGE_BYTES = KSTEP * TILE_M;  // 4096
constexpr int MAX_TILES = 4;                 // tiles per strip
constexpr int VCOLS = MAX_TILES * TILE_M;    // 512 flat columns of vertically filtered data
constexpr int VPITCH = OYB + 4;              // floats per column of V (pad: conflict-free 128-bit stores)
constexpr int VPADC = 32;                    // extra V columns a lane may read past its own window (zero weights, finite data)
constexpr int VALLOC = VCOLS + VPADC;
constexpr int NACC = 5;                      // accumulator buffers in TMEM
constexpr int ACC_COLS = 96;                 // columns per buffer (= UMMA_N); 5 x 96 = 480 of the 512 allocated
constexpr int MAX_KSTEPS = 8;
constexpr int NWC = 16, NTC = NWC * 32, NT = NTC + 64;  // 16 epilogue warps + producer + MMA issuer
constexpr int NGRP = 4, GROWS = OYB / NGRP;             // 4 independent epilogue groups of 4 warps, 8 output rows each
constexpr int BAR_BYTES = 1024;
// Wait-time counters (aa_debug_counters, scripts/vmma_prof.py) are compiled in only with -DAA_VMMA_PROFILE: the
// kernel is sensitive to its code size (see the note above aa_vmma_kernel).
#ifdef AA_VMMA_PROFILE
constexpr bool kProf = true;
#else
constexpr bool kProf = false;
#endif
constexpr int MAX_STAGES = 40;

struct VParams {
  SParams S;            // output, epilogue, horizontal tables and strip plan (strip_setup reads these)
  const int8_t* bq;     // [n_oyb][ksteps*32][128] weight limbs, 128B-swizzled rows (aa_tables.cu)
  const float* qmeta;   // {c0, c1, c2, K0}: fp32 = fma(m2, c2, fma(m1, c1, fma(m0, c0, K0)))
  int ksteps, n_oyb, nstage, b_bytes, Cp_in;
  int oyb;  // output rows per item: OYB (32), or 16 for vertical scales whose 32-row blocks would not fit the K span
  int64_t total_items;
  uint32_t idesc;       // tcgen05 instruction descriptor
  uint64_t desc_tmpl;   // shared-memory matrix descriptor without the start address
  int* dbg;             // device: [0] = first watchdog code (0 = none); 64-bit profile counters from byte 32
  long long timeout_clk;
  int prof;
  int hr;  // rows per thread in the horizontal pass (2, 4 or 8)
  int vt;  // V chunk rotation shift (31 = none), see vchunk()
  int st2;  // fp32 planar output whose base pointer and strides are all even: column pairs go out as 64-bit stores
  int stagger;  // cycles the odd epilogue groups wait after every strip change (de-phases the groups), 0 = none
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a wrong descriptor or a lost transaction must not hang the GPU.  After timeout_clk cycles the
// waiter records `code` and raises the CTA-wide abort flag; every role then drains its loops without waiting
// (results are garbage, the host reports AA_ERR_CUDA from the debug word).
struct Watch {
  volatile int* abort_flag;
  int* dbg;
  long long limit;
  int prof;  // AA_VMMA_PROF=1: the cycles spent in each kind of wait are summed in `acc` and flushed once at kernel end
  long long* acc;  // [16] per-thread accumulators (registers / local)
};
__device__ __forceinline__ bool mbar_wait(const Watch& w, uint32_t bar, uint32_t parity, int code) {
  if (mbar_try(bar, parity)) return true;
  const long long t0 = clock64();
  for (int spin = 0;; spin++) {
    if (mbar_try(bar, parity)) {
      if (kProf && w.prof) w.acc[code] += clock64() - t0;
      return true;
    }
    if ((spin & 63) == 63) {
      if (*w.abort_flag) return false;
      if (clock64() - t0 > w.limit) {
        *w.abort_flag = 1;
        atomicCAS(w.dbg, 0, code);
        return false;
      }
    }
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int c, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y), "r"(c), "r"(n)
      : "memory");
}
// named barriers: 1 + g = the 128 threads of epilogue group g (they own rows [8g, 8g+8) of every item and run
// independently of the other groups, so that one group's TMEM reads overlap another's shared-memory gather and the
// schedulers always have several warps in different phases to pick from); 8 = all epilogue threads (strip changes only)
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(NTC / NGRP) : "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 8, %0;" ::"n"(NTC) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Item {
  int64_t plane;
  int strip, oyb, ox0, ox1, fl0, ntiles;
};
__device__ __forceinline__ void item_strip(const VParams& P, Item& it) {  // geometry of it.strip
  it.ox0 = it.strip * P.S.strip_ox;
  it.ox1 = min((int)P.S.oW, it.ox0 + P.S.strip_ox);
  it.fl0 = (__ldg(P.S.xmin_w + it.ox0) * P.S.Ci) & ~(P.S.aln - 1);  // TMA: the box must start on a 16-byte boundary
  const int fl_end = (__ldg(P.S.xmin_w + it.ox1 - 1) + __ldg(P.S.xsize_w + it.ox1 - 1)) * P.S.Ci;
  it.ntiles = (fl_end - it.fl0 + TILE_M - 1) / TILE_M;
}
// items are numbered (strip, plane, oyb) with oyb fastest: a CTA walks down one strip of one plane (part of the halo rows
// of consecutive items are L2 hits), then takes the same strip of the next plane -- its range of items touches one or two
// strips, so the strip tables (Wp, pinfo) and the all-warps barrier that guards them are set up once or twice per CTA
// instead of once per 16 items (cfg3 1.012 -> 0.96-0.97 ms, the uint8 sweep 0.415 -> 0.435 of peak).  With the plane
// fastest instead ((strip, oyb, plane): the weight matrix B would change once per CTA, not per item) the halo reuse is
// lost: 0.7 % slower on cfg3, 1.2 % on the sweep, same-box A/B.
__device__ __forceinline__ Item item_first(const VParams& P, int64_t i) {
  Item it;
  const int64_t t = i / P.n_oyb;
  it.oyb = (int)(i - t * P.n_oyb);
  it.strip = (int)(t / P.S.lin.planes);
  it.plane = t - (int64_t)it.strip * P.S.lin.planes;
  item_strip(P, it);
  return it;
}
__device__ __forceinline__ void item_next(const VParams& P, Item& it) {
  if (++it.oyb < P.n_oyb) return;
  it.oyb = 0;
  if (++it.plane < P.S.lin.planes) return;
  it.plane = 0;
  if (++it.strip < P.S.n_strips) item_strip(P, it);  // (past the last item: nothing to look up)
}

// Horizontal pass of one epilogue group over its 16 rows of the transposed buffer V[flat column][VPITCH]: one work
// item = one pair of adjacent output columns (one channel) x R output rows (R = 8 or 4); the rows of a tap arrive in
// R/4 LDS.128 and the pair's two weights in one LDS.64.  A warp covers 32*R/16 pairs x 16/R row chunks, so the weight
// loads of the chunks coalesce and the V loads spread over all banks (pitch 36).  Window bookkeeping
// (Wp, pinfo) is the pair form of aa_stream_common.cuh::strip_setup; as there, a lane never reads a column outside
// its own pair's windows (the pointer stops advancing, the weights past the window are zero padding).
// Position of 4-row chunk `cq` inside column x of V.  With vt < 31 the chunk is rotated by (x >> vt): pair columns that
// are a multiple of 8 flat columns apart (integer scales 2, 4, 8: the padded pitch alone maps them to the same banks,
// a 16-way conflict on the 128-bit loads) then land in different bank groups; 2^vt = the largest power of two dividing
// the pair spacing, chosen by the host plan.  vt = 31: no rotation (odd spacings are conflict-free by the pitch).
__device__ __forceinline__ int vchunk(int cq, int x, int vt) { return (cq + (x >> vt)) & 7; }

template <bool GEN, int R, int CI, bool ROT>
__device__ __forceinline__ void hphase_T(const VParams& P, const float* __restrict__ V, const float2* __restrict__ Wp,
                                         const int4* __restrict__ pinfo, int64_t op_off, int npc, int tg, int grp, int oy0, int nrows) {
  constexpr int NCH = GROWS / R;       // row chunks of the group's 8 rows
  constexpr int PPW = 32 / NCH;        // pair-columns per warp iteration
  constexpr int SH = R == 8 ? 5 : (R == 4 ? 4 : 3);   // log2(PPW)
  const int Ci = CI ? CI : P.S.Ci;
  const int vt = P.vt;
  const int64_t osh = P.S.lout.stride_h;
  for (int it = tg; ((it >> 5) << SH) < npc; it += NTC / NGRP) {
    const int pc = ((it >> 5) << SH) | (it & (PPW - 1));
    const int r0 = grp * GROWS + ((it & 31) >> SH) * R;  // first of this thread's R rows inside the block
    const bool act = pc < npc;
    const int4 pi = pinfo[act ? pc : 0];
    // warp-uniform trip count = the longest union window of the warp's pairs.  A lane whose own window is shorter
    // keeps walking: its weights there are the table's zero padding and V holds finite values everywhere (uint8 data,
    // buffer zeroed at start; the host plan keeps pi.x + trip*Ci inside the VALLOC columns), so nothing leaks.
    const int lenm = __reduce_max_sync(0xffffffffu, act ? (pi.z & 0xffff) : 1);
    const float2* wr = Wp + pi.y;
    const float* vp = V + pi.x * VPITCH + r0;
    const int vstep = Ci * VPITCH;
    float2 h[R];
#pragma unroll
    for (int r = 0; r < R; r++) h[r] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < lenm; j++) {
      const float2 w2 = wr[j];
      const float* vj;
      if constexpr (ROT) {
        const int x = pi.x + j * Ci;
        vj = V + x * VPITCH + 4 * vchunk(r0 >> 2, x, vt) + (r0 & 3);
      } else {
        vj = vp + j * vstep;
      }
      if constexpr (R >= 4) {
#pragma unroll
        for (int q = 0; q < R; q += 4) {
          const float4 v = ROT && q ? *reinterpret_cast<const float4*>(V + (pi.x + j * Ci) * VPITCH + 4 * vchunk((r0 >> 2) + 1, pi.x + j * Ci, vt))
                                   : *reinterpret_cast<const float4*>(vj + q);
          h[q + 0] = __ffma2_rn(make_float2(v.x, v.x), w2, h[q + 0]);
          h[q + 1] = __ffma2_rn(make_float2(v.y, v.y), w2, h[q + 1]);
          h[q + 2] = __ffma2_rn(make_float2(v.z, v.z), w2, h[q + 2]);
          h[q + 3] = __ffma2_rn(make_float2(v.w, v.w), w2, h[q + 3]);
        }
      } else {
        const float2 v = *reinterpret_cast<const float2*>(vj);
        h[0] = __ffma2_rn(make_float2(v.x, v.x), w2, h[0]);
        h[1] = __ffma2_rn(make_float2(v.y, v.y), w2, h[1]);
      }
    }
    if (act) {
      const int64_t dst = op_off + (int64_t)(oy0 + r0) * osh + pi.w;
      const bool hasb = ((pi.z >> 16) & 1) != 0;
      const int c = pi.z >> 20, cstep = P.S.epi.colstep(Ci);
      // planar fp32 output, both columns of the pair present, 8-byte aligned: one 64-bit store per row instead of two 32-bit
      // ones (P.st2: base pointer and all strides are even, checked on the host; the pair's own column offset here)
      if (!GEN && P.st2 && hasb && Ci == 1 && ((dst & 1) == 0) && r0 + R <= nrows) {
        float* o = reinterpret_cast<float*>(P.S.out) + dst;
#pragma unroll
        for (int r = 0; r < R; r++) *reinterpret_cast<float2*>(o + (int64_t)r * osh) = h[r];
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (r0 + r < nrows) {
            aa_store<GEN>(P.S.out, dst + (int64_t)r * osh, h[r].x, c, P.S.epi);
            if (hasb) aa_store<GEN>(P.S.out, dst + (int64_t)r * osh + cstep, h[r].y, c, P.S.epi);
          }
        }
      }
    }
  }
}
// GEN: decode-adjacent epilogue; CI: compile-time channel interleave of the flat rows (0 = runtime); HR: rows per thread in
// the horizontal pass; ROT: rotated V chunks (even pair spacings); OYBR: output rows per item (32, or 16 for scales > ~7x).
// One instantiation per combination keeps each kernel's code small: with four epilogue groups in different phases the
// instruction caches hold both the epilogue and the horizontal pass (a kernel that carried every variant behind runtime
// switches measured 7-11 % slower on cfg3).
template <bool GEN, int CI, int HR, bool ROT, int OYBR>
__global__ void __launch_bounds__(NT, 1) aa_vmma_kernel(const __grid_constant__ CUtensorMap tmap, const VParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // the 128-byte swizzle atoms need 1024-byte alignment
  unsigned char* sm = smem_raw + (base - raw);
  // layout: A ring [nstage][b_bytes] (one stage = the whole K span of one tile) | B [2][b_bytes] | V [VCOLS][VPITCH] f32 |
  //         barriers | Wp | pinfo
  const uint32_t sA = base;
  const uint32_t sB = sA + (uint32_t)P.nstage * (uint32_t)P.b_bytes;
  float* V = reinterpret_cast<float*>(sm + ((size_t)P.nstage + 2) * (size_t)P.b_bytes);
  unsigned char* bar_base = reinterpret_cast<unsigned char*>(V) + sizeof(float) * VALLOC * VPITCH;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bar_base);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * MAX_STAGES;
  const uint32_t tfull0 = empty0 + 8 * MAX_STAGES, tempty0 = tfull0 + 8 * NACC;
  const uint32_t bfull0 = tempty0 + 8 * NACC, bempty0 = bfull0 + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_base + 8 * (2 * MAX_STAGES + 2 * NACC + 4));
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float2* Wp = reinterpret_cast<float2*>(bar_base + BAR_BYTES);
  int4* pinfo = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(Wp) + P.S.wtab_bytes + 15) & ~(uintptr_t)15);

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) {
    for (int i = 0; i < P.nstage; i++) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); }
    for (int i = 0; i < NACC; i++) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, NWC); }
    for (int i = 0; i < 2; i++) { mbar_init(bfull0 + 8 * i, 1); mbar_init(bempty0 + 8 * i, 1); }
    *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM by shared-memory footprint)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long wacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const Watch W{abort_flag, P.dbg, P.timeout_clk, P.prof, wacc};
  const long long t_kernel0 = kProf ? clock64() : 0;
  if (warp >= 2) {
    // V: finite everywhere from the start (lanes may read zero-weight taps past their own window)
    for (int i = t - 64; i < VALLOC * VPITCH / 4; i += NTC) reinterpret_cast<float4*>(V)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();

  const int64_t i_begin = P.total_items * (int64_t)blockIdx.x / gridDim.x;
  const int64_t i_end = P.total_items * (int64_t)(blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0, bslot = 0;
      uint32_t phase = 0, bphase = 0;
      bool ok = true;
      Item it = item_first(P, i_begin);
      int b_oyb = -1;  // row block whose weight matrix was copied last
      int pc = 0, pn = 0;  // tensor-map coordinates of the item's plane
      int64_t tm_plane = -1;
      for (int64_t i = i_begin; i < i_end && ok; i++, item_next(P, it)) {
        if (it.oyb != b_oyb) {
          ok = mbar_wait(W, bempty0 + 8 * bslot, bphase ^ 1, 1);
          if (!ok) break;
          mbar_expect_tx(bfull0 + 8 * bslot, (uint32_t)P.b_bytes);
          bulk_g2s(sB + (uint32_t)bslot * P.b_bytes, P.bq + (size_t)it.oyb * P.b_bytes, (uint32_t)P.b_bytes, bfull0 + 8 * bslot);
          if (++bslot == 2) { bslot = 0; bphase ^= 1; }
          b_oyb = it.oyb;
        }
        const int y0 = __ldg(P.S.xmin_h + it.oyb * OYBR);
        if (it.plane != tm_plane) {  // (the 64-bit divisions once per plane, not per item)
          pc = (int)(it.plane % P.Cp_in);
          pn = (int)(it.plane / P.Cp_in);
          tm_plane = it.plane;
        }
        for (int s = 0; s < it.ntiles; s++) {  // one TMA box per tile: 128 flat columns x ksteps*32 rows
          ok = mbar_wait(W, empty0 + 8 * stage, phase ^ 1, 2);
          if (!ok) break;
          mbar_expect_tx(full0 + 8 * stage, (uint32_t)P.b_bytes);
          tma_load_4d(sA + (uint32_t)stage * P.b_bytes, &tmap, full0 + 8 * stage, it.fl0 + s * TILE_M, y0, pc, pn);
          if (++stage == P.nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      int stage = 0, bslot = 0, acc = 0;
      uint32_t phase = 0, bphase = 0, aphase = 0;
      bool ok = true;
      Item it = item_first(P, i_begin);
      int b_oyb = -1, b_cur = -1;  // row block / slot of the weight matrix in use
      uint64_t bdesc0 = 0;
      for (int64_t i = i_begin; i < i_end && ok; i++, item_next(P, it)) {
        if (it.oyb != b_oyb) {
          if (b_cur >= 0) umma_commit(bempty0 + 8 * b_cur);  // the previous matrix is free once the MMAs issued so far are done
          ok = mbar_wait(W, bfull0 + 8 * bslot, bphase, 3);
          if (!ok) break;
          bdesc0 = P.desc_tmpl | (uint64_t)(((sB + (uint32_t)bslot * P.b_bytes) >> 4) & 0x3FFFu);
          b_cur = bslot;
          b_oyb = it.oyb;
          if (++bslot == 2) { bslot = 0; bphase ^= 1; }
        }
        for (int s = 0; s < it.ntiles && ok; s++) {
          ok = mbar_wait(W, tempty0 + 8 * acc, aphase ^ 1, 4);
          if (!ok) break;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * (uint32_t)ACC_COLS;
          ok = mbar_wait(W, full0 + 8 * stage, phase, 5);
          if (!ok) break;
          tc_fence_after();
          const uint64_t adesc0 = P.desc_tmpl | (uint64_t)(((sA + (uint32_t)stage * P.b_bytes) >> 4) & 0x3FFFu);
          for (int ks = 0; ks < P.ksteps; ks++)  // 32 rows (4096 bytes) of both operands per MMA
            umma_i8(d_tmem, adesc0 + (uint64_t)(ks * (STAGE_BYTES >> 4)), bdesc0 + (uint64_t)(ks * (STAGE_BYTES >> 4)), P.idesc, ks > 0 ? 1u : 0u);
          umma_commit(empty0 + 8 * stage);  // the stage is free once these MMAs have read it
          if (++stage == P.nstage) { stage = 0; phase ^= 1; }
          umma_commit(tfull0 + 8 * acc);
          if (++acc == NACC) { acc = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // =============================== epilogue + horizontal pass =============================
    const int tc = t - 64;
    const int q = warp & 3;          // TMEM lane quarter this warp may read (warp id mod 4)
    const int half = (warp - 2) >> 2;  // epilogue group: which 8 of the 32 output rows
    const float c0 = __ldg(P.qmeta + 0), c1 = __ldg(P.qmeta + 1), c2 = __ldg(P.qmeta + 2), k0 = __ldg(P.qmeta + 3);
    int acc = 0;
    uint32_t aphase = 0;
    int cur_strip = -1, strip_fl0 = 0, strip_npc = 0;
    bool ok = true;
    int64_t op = 0, op_plane = -1;  // element offset of the item's output plane
    Item it = item_first(P, i_begin);
    for (int64_t i = i_begin; i < i_end; i++, item_next(P, it)) {
      if (it.strip != cur_strip) {
        consumer_sync();
        strip_setup(P.S, tc, NTC, it.ox0, it.ox1, Wp, pinfo, &strip_fl0, &strip_npc);
        cur_strip = it.strip;
        consumer_sync();
        // The groups leave this barrier in lock-step: all four would read TMEM together and then gather from shared
        // memory together.  Holding the odd groups back puts their gather next to the even groups' epilogue for the 16
        // items until the next strip change.  cfg3, four accumulator buffers: 1.041 -> 1.022 ms at 1000 cycles (2000: 1.025,
        // 3000: 1.032, 4500: 1.051 -- a group further behind than the ring allows stalls the MMA warp); five buffers:
        // 1.046 -> 1.012 ms at 2000-2500 cycles (1000: 1.025, 3000: 1.018).
        if (P.stagger > 0 && (half & 1)) {
          const long long t0 = clock64();
          while (clock64() - t0 < (long long)P.stagger) {}
        }
      }
      for (int s = 0; s < it.ntiles; s++) {
        if (ok) ok = mbar_wait(W, tfull0 + 8 * acc, aphase, 6);
        ok = __all_sync(0xffffffffu, ok);  // the TMEM loads below are warp-collective
        tc_fence_after();
        if (ok && half * GROWS >= OYBR) {  // 16-row items: the upper two groups have no rows; they only release the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
        } else if (ok) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + half * GROWS);
          int l0[GROWS], l1[GROWS], l2[GROWS];
          const long long t_e0 = kProf && P.prof ? clock64() : 0;
          tmem_ld8(taddr, l0);
          tmem_ld8(taddr + OYB, l1);
          tmem_ld8(taddr + 2 * OYB, l2);
          tmem_ld_wait();
          if (kProf && P.prof) wacc[7] += clock64() - t_e0;  // TMEM read
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + 8 * acc);  // accumulators are in registers: the MMA warp may reuse the buffer
          const int vx = s * TILE_M + q * 32 + lane;
          float* vcol = V + (size_t)vx * VPITCH;
          const float2 c0v = make_float2(c0, c0), c1v = make_float2(c1, c1), c2v = make_float2(c2, c2), k0v = make_float2(k0, k0);
#pragma unroll
          for (int e = 0; e < GROWS; e += 4) {
            float o[4];
#pragma unroll
            for (int p = 0; p < 4; p += 2) {
              // int -> float by the 1.5*2^23 magic (|limb sum| < 2^22); the bias is folded into k0 (see aa_tables.cu).
              // (Pre-arming the accumulators with the magic through tcgen05.st instead was measured slower: ptxas
              // re-materialises the 16 constant source registers of every STTM.)
              const float2 m0 = make_float2(__int_as_float(l0[e + p] + 0x4B400000), __int_as_float(l0[e + p + 1] + 0x4B400000));
              const float2 m1 = make_float2(__int_as_float(l1[e + p] + 0x4B400000), __int_as_float(l1[e + p + 1] + 0x4B400000));
              const float2 m2 = make_float2(__int_as_float(l2[e + p] + 0x4B400000), __int_as_float(l2[e + p + 1] + 0x4B400000));
              float2 u = __ffma2_rn(m0, c0v, k0v);
              u = __ffma2_rn(m1, c1v, u);
              u = __ffma2_rn(m2, c2v, u);
              o[p] = u.x;
              o[p + 1] = u.y;
            }
            float* vdst = ROT ? vcol + 4 * vchunk((half * GROWS + e) >> 2, vx, P.vt) : vcol + half * GROWS + e;
            *reinterpret_cast<float4*>(vdst) = make_float4(o[0], o[1], o[2], o[3]);
          }
          if (kProf && P.prof) wacc[11] += clock64() - t_e0;  // whole epilogue of the tile (TMEM read + convert + store)
        }
        if (++acc == NACC) { acc = 0; aphase ^= 1; }
      }
      const long long t_h0 = kProf ? clock64() : 0;
      group_sync(half);  // this group's 16 rows of all tiles are in V
      const long long t_h1 = kProf ? clock64() : 0;
      if (it.plane != op_plane) {  // 16 consecutive items share the plane: the 64-bit divisions once per plane, not per item
        op = (it.plane / P.S.lout.Cp) * P.S.lout.stride_n + (it.plane % P.S.lout.Cp) * P.S.lout.stride_p;
        op_plane = it.plane;
      }
      const int nrows = min(OYBR, (int)P.S.oH - it.oyb * OYBR);
      // rows per thread: the pass is bound by shared-memory bandwidth (2 B of V per FMA for a column pair + 4/R B of
      // weights), so more rows per thread is less traffic; fewer rows only when the strip is too narrow to occupy the warps
      if (half * GROWS < nrows)  // else: no row of this group in the item (last block of the image, or 16-row items)
        hphase_T<GEN, HR, CI, ROT>(P, V, Wp, pinfo, op, strip_npc, tc & (NTC / NGRP - 1), half, it.oyb * OYBR, nrows);
      const long long t_h2 = kProf ? clock64() : 0;
      group_sync(half);  // V rows of this group may be overwritten
      if (kProf && P.prof) {
        wacc[8] += t_h1 - t_h0;        // barrier before the horizontal pass
        wacc[9] += t_h2 - t_h1;        // horizontal pass
        wacc[10] += clock64() - t_h2;  // barrier after it
      }
    }
  }
  if (kProf && P.prof && lane == 0) {
    unsigned long long* pc = reinterpret_cast<unsigned long long*>(P.dbg) + 4;
    for (int k = 1; k < 12; k++)
      if (wacc[k]) atomicAdd(pc + k, (unsigned long long)wacc[k]);
    atomicAdd(pc + 12 + (warp == 0 ? 0 : warp == 1 ? 1 : 2), (unsigned long long)(clock64() - t_kernel0));  // role lifetimes
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

struct VPlanKey {
  uint64_t th, tw;
  int Ci, gen;
  bool operator<(const VPlanKey& o) const {
    if (th != o.th) return th < o.th;
    if (tw != o.tw) return tw < o.tw;
    if (Ci != o.Ci) return Ci < o.Ci;
    return gen < o.gen;
  }
};
struct VPlan {
  int n_strips, strip_ox, kp, wtab_bytes, nstage, sms, vt;
  size_t smem;
};
std::mutex g_vplan_mu;
std::map<VPlanKey, VPlan> g_vplans;
int* g_dbg[64] = {};

int debug_env(const char* name, long long dflt, long long* out) {
  const char* e = getenv(name);
  *out = e ? strtoll(e, nullptr, 0) : dflt;
  return e != nullptr;
}

}  // namespace

namespace {
typedef void (*VKernel)(const CUtensorMap, const VParams);
constexpr size_t cap_total_bytes = 227 * 1024;
template <bool GEN, int CI, int HR, bool ROT>
VKernel pick_oyb(int oyb) {
  return oyb == 16 ? aa_vmma_kernel<GEN, CI, HR, ROT, 16> : aa_vmma_kernel<GEN, CI, HR, ROT, 32>;
}
template <bool GEN, int CI, bool ROT>
VKernel pick_hr(int hr, int oyb) { return hr == 4 ? pick_oyb<GEN, CI, 4, ROT>(oyb) : pick_oyb<GEN, CI, 2, ROT>(oyb); }
template <bool GEN>
VKernel pick_ci(int ci, int hr, bool rot, int oyb) {
  if (rot) return pick_hr<GEN, 0, true>(hr, oyb);  // rotated chunks: per-tap address arithmetic, runtime interleave
  switch (ci) {
    case 1: return pick_hr<GEN, 1, false>(hr, oyb);
    case 3: return pick_hr<GEN, 3, false>(hr, oyb);
    case 4: return pick_hr<GEN, 4, false>(hr, oyb);
    default: return pick_hr<GEN, 0, false>(hr, oyb);
  }
}
VKernel pick_kernel(bool gen, int ci, int hr, bool rot, int oyb) {
  return gen ? pick_ci<true>(ci, hr, rot, oyb) : pick_ci<false>(ci, hr, rot, oyb);
}
}  // namespace

void vmma_plan_clear() {
  std::lock_guard<std::mutex> lock(g_vplan_mu);
  g_vplans.clear();
}

// Polls the watchdog word of `device` (set by a kernel whose mbarrier wait timed out).  Called by tests and by the
// host-buffer entry after its synchronisation point; costs one 4-byte D2H copy.
int vmma_check_watchdog(int device) {
  if (device < 0 || device >= 64 || !g_dbg[device]) return AA_OK;
  int v = 0;
  AA_CUDA_TRY(cudaMemcpy(&v, g_dbg[device], sizeof(int), cudaMemcpyDeviceToHost));
  if (v != 0) {
    int zero = 0;
    cudaMemcpy(g_dbg[device], &zero, sizeof(int), cudaMemcpyHostToDevice);
    return fail(AA_ERR_CUDA, "vmma: mbarrier wait timed out in the tensor-core kernel (watchdog code " + std::to_string(v) + ")");
  }
  return AA_OK;
}

// The 16 profile counters of `device` (AA_VMMA_PROF=1): cycles summed over waiting warps, see aa_debug_counters.
int vmma_read_counters(int device, unsigned long long* out, int reset) {
  for (int i = 0; i < 16; i++) out[i] = 0;
  if (device < 0 || device >= 64 || !g_dbg[device]) return AA_OK;
  AA_CUDA_TRY(cudaMemcpy(out, reinterpret_cast<char*>(g_dbg[device]) + 32, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (reset) AA_CUDA_TRY(cudaMemset(reinterpret_cast<char*>(g_dbg[device]) + 32, 0, 16 * sizeof(unsigned long long)));
  return AA_OK;
}

static int vmma_prepare_device(int device) {
  std::lock_guard<std::mutex> lock(g_vplan_mu);
  if (device < 0 || device >= 64) return fail(AA_ERR_INVALID, "bad device ordinal");
  if (!g_dbg[device]) {
    int* d = nullptr;
    AA_CUDA_TRY(cudaMalloc(&d, 64 * sizeof(int)));
    AA_CUDA_TRY(cudaMemset(d, 0, 64 * sizeof(int)));
    g_dbg[device] = d;
  }
  return AA_OK;
}

int vmma_warm(AxisTables* th, cudaStream_t stream) {
  int rc = vmma_prepare_device(th->device);
  if (rc != AA_OK) return rc;
  return ensure_vq_tables(th, OYB, KSTEP, MAX_KSTEPS, stream);
}

int launch_vmma(const void* in, const Layout& lin, void* out, const Layout& lout, AxisTables* th, AxisTables* tw, int64_t H,
                int64_t W, int64_t oH, int64_t oW, OutEpi epi, cudaStream_t stream) {
  if (th->dtype != AA_F32 || tw->dtype != AA_F32) return fail(AA_ERR_UNSUPPORTED, "vmma: f32 tables only");
  const int Ci = lin.Ci;
  // TMA: 16-byte aligned base and strides (uint8: elements == bytes)
  if (((uintptr_t)in) % 16 || lin.stride_h % 16 || (lin.planes > lin.Cp && lin.stride_n % 16) || (lin.Cp > 1 && lin.stride_p % 16))
    return fail(AA_ERR_UNSUPPORTED, "vmma: input rows are not 16-byte aligned");
  if (W * Ci >= (1ll << 31) || H >= (1ll << 31) || oH >= (1 << 24) || tw->K >= (1 << 16) || Ci > 2047)
    return fail(AA_ERR_UNSUPPORTED, "vmma: size limits");
  if (th->xsize_max > 128) return fail(AA_ERR_UNSUPPORTED, "vmma: vertical window too long for the int32 accumulators");
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(AA_ERR_UNSUPPORTED, "vmma: cuTensorMapEncodeTiled is not available");
  int rc = ensure_vq_tables(th, OYB, KSTEP, MAX_KSTEPS, stream);
  if (rc != AA_OK) return rc;

  VParams P;
  P.S = SParams();
  P.S.in = in; P.S.out = out; P.S.epi = epi; P.S.lin = lin; P.S.lout = lout; P.S.Ci = Ci;
  P.S.H = H; P.S.oH = oH; P.S.oW = oW;
  P.S.xmin_h = th->xmin; P.S.xsize_h = th->xsize;
  P.S.xmin_w = tw->xmin; P.S.xsize_w = tw->xsize; P.S.w_w = (const float*)tw->w; P.S.Kw = tw->K;
  P.S.aln = 16; P.S.pairs = 1; P.S.pad = 0;
  P.bq = th->vq; P.qmeta = th->vq_meta; P.ksteps = th->vq_ksteps; P.n_oyb = th->vq_noyb; P.oyb = th->vq_oyb;
  P.b_bytes = P.ksteps * STAGE_BYTES;
  P.Cp_in = lin.Cp;

  const bool gen = epi.generic();
  const VPlanKey key{th->id, tw->id, Ci, gen ? 1 : 0};
  VPlan pl;
  bool have = false;
  {
    std::lock_guard<std::mutex> lock(g_vplan_mu);
    auto itp = g_vplans.find(key);
    if (itp != g_vplans.end()) { pl = itp->second; have = true; }
  }
  if (!have) {
    const int32_t* xs = tw->h_xmin.data();
    const int32_t* xz = tw->h_xsize.data();
    const size_t fixed = (size_t)2 * P.b_bytes + sizeof(float) * VALLOC * VPITCH + BAR_BYTES + 1024 /*alignment slack*/;
    const size_t stage_bytes = (size_t)P.b_bytes;  // one stage = one tile's K span
    const size_t cap_total = 227 * 1024;
    int n_strips = 1, strip_ox = (int)oW, kp = 0, wtab = 0;
    size_t tab_bytes = 0;
    for (;; n_strips++) {
      if (n_strips > oW) return fail(AA_ERR_UNSUPPORTED, "vmma: a single output column spans more than one strip");
      strip_ox = (int)((oW + n_strips - 1) / n_strips);
      bool ok = strip_ox <= 2048;
      int shift = 0;
      for (int64_t a = 0; ok && a < oW; a += strip_ox) {
        const int64_t b = std::min<int64_t>(oW, a + strip_ox) - 1;
        const int64_t f0 = ((int64_t)xs[a] * Ci) & ~15ll;
        if (((int64_t)xs[b] + xz[b]) * Ci - f0 > VCOLS) ok = false;
        // every pair may walk the strip's longest union window: keep first tap + longest * Ci inside the VALLOC columns
        int64_t longest = 1;
        for (int64_t o = a; o <= b; o += 2) {
          const int64_t e1 = o + 1 <= b ? (int64_t)xs[o + 1] + xz[o + 1] : 0;
          longest = std::max<int64_t>(longest, std::max<int64_t>((int64_t)xs[o] + xz[o], e1) - xs[o]);
        }
        for (int64_t o = a; ok && o <= b; o += 2)
          if (((int64_t)xs[o] + longest) * Ci + (Ci - 1) - f0 > VALLOC) ok = false;
        for (int64_t o = a; o + 1 <= b; o += 2) shift = std::max<int>(shift, xs[o + 1] - xs[o]);
      }
      if (ok) {
        kp = (tw->K + shift) | 1;
        const int np = (strip_ox + 1) / 2;
        wtab = (int)((size_t)np * kp * sizeof(float2));
        tab_bytes = (size_t)wtab + 16 + (size_t)np * Ci * sizeof(int4);
        if (kp >= (1 << 16) || fixed + tab_bytes + 2 * stage_bytes > cap_total) ok = false;
      }
      if (ok) break;
    }
    n_strips = (int)((oW + strip_ox - 1) / strip_ox);
    int nstage = (int)((cap_total - fixed - tab_bytes) / stage_bytes);
    if (nstage > MAX_STAGES) nstage = MAX_STAGES;
    if (nstage < 2) return fail(AA_ERR_UNSUPPORTED, "vmma: shared memory plan too large");
    pl.n_strips = n_strips; pl.strip_ox = strip_ox; pl.kp = kp; pl.wtab_bytes = wtab; pl.nstage = nstage;
    {
      // spacing of adjacent pair columns in flat elements (taken in the middle of the row): even -> rotate the chunks
      const int64_t om = std::min<int64_t>(oW - 1, (oW / 2) & ~1ll);
      const int64_t d = om + 2 < oW ? ((int64_t)xs[om + 2] - xs[om]) * Ci : 1;
      int tz = 31;
      if (d > 0 && (d & 1) == 0) {
        tz = 0;
        while (tz < 5 && ((d >> tz) & 1) == 0) tz++;
      }
      pl.vt = tz;
    }
    pl.smem = fixed + tab_bytes + (size_t)nstage * stage_bytes;
    AA_CUDA_TRY(cudaDeviceGetAttribute(&pl.sms, cudaDevAttrMultiProcessorCount, th->device));
    if ((rc = vmma_prepare_device(th->device)) != AA_OK) return rc;
    std::lock_guard<std::mutex> lock(g_vplan_mu);
    if (g_vplans.size() > 4096) g_vplans.clear();
    g_vplans[key] = pl;
  }
  P.S.n_strips = pl.n_strips; P.S.strip_ox = pl.strip_ox; P.S.kp = pl.kp; P.S.wtab_bytes = pl.wtab_bytes;
  P.nstage = pl.nstage;
  P.vt = pl.vt;
  P.st2 = (!gen && epi.kind == 0 && !epi.planar && Ci == 1 && ((uintptr_t)out) % 8 == 0 && lout.stride_h % 2 == 0 && lout.stride_n % 2 == 0 &&
           (lout.Cp == 1 || lout.stride_p % 2 == 0)) ? 1 : 0;
  P.total_items = lin.planes * pl.n_strips * P.n_oyb;
  P.dbg = g_dbg[th->device];

  // tcgen05 descriptors.  Instruction: D = s32, A = u8 (the image tile), B = s8 (weight limbs), both operands
  // MN-major (the contiguous dimension is M resp. N), M = 128, N = 96.  Shared-memory matrices: 128-byte swizzle,
  // 8 K-rows of 128 bytes per atom -> stride between atoms along K (SBO) = 1024 bytes; one atom wide along M/N.
  long long v;
  P.idesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(UMMA_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  if (debug_env("AA_VMMA_IDESC", 0, &v)) P.idesc = (uint32_t)v;
  long long lbo = STAGE_BYTES, sbo = 1024, lay = 2;
  debug_env("AA_VMMA_LBO", lbo, &lbo);
  debug_env("AA_VMMA_SBO", sbo, &sbo);
  debug_env("AA_VMMA_LAYOUT", lay, &lay);
  P.desc_tmpl = ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)lay << 61);
  // measured on cfg3 (strip of 32 pairs): 4 rows/thread 1.148 ms, 2 rows 1.185 ms, 8 rows 1.554 ms (one busy warp per
  // group: latency-bound); narrow strips take 2 rows so that all four warps of a group have work
  debug_env("AA_VMMA_R", ((pl.strip_ox + 1) / 2) * Ci > 16 ? 4 : 2, &v);
  P.hr = (int)v;
  debug_env("AA_VMMA_STAGGER", 2200, &v);
  P.stagger = (int)v;
  debug_env("AA_VMMA_PROF", 0, &v);
  P.prof = (int)v;
  debug_env("AA_VMMA_TIMEOUT_MS", 2000, &v);
  P.timeout_clk = v * 2000000ll;  // ~2 GHz

  // 4-D tensor map over the uint8 input: {flat x (W*Ci bytes), y, channel plane, image}
  CUtensorMap tmap;
  const int64_t n_img = lin.planes / lin.Cp;
  const cuuint64_t dims[4] = {(cuuint64_t)(W * Ci), (cuuint64_t)H, (cuuint64_t)lin.Cp, (cuuint64_t)n_img};
  const cuuint64_t any16 = (cuuint64_t)lin.stride_h * (cuuint64_t)H;  // placeholder stride of a size-1 dimension
  const cuuint64_t strides[3] = {(cuuint64_t)lin.stride_h, lin.Cp > 1 ? (cuuint64_t)lin.stride_p : any16,
                                 n_img > 1 ? (cuuint64_t)lin.stride_n : any16};
  const cuuint32_t box[4] = {TILE_M, (cuuint32_t)(P.ksteps * KSTEP), 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  debug_env("AA_VMMA_L2PROMO", 1, &v);  // 0 none, 1 64B, 2 128B, 3 256B
  const CUtensorMapL2promotion promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(in), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(AA_ERR_UNSUPPORTED, "vmma: cuTensorMapEncodeTiled failed (" + std::to_string((int)cr) + ")");

  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(pl.sms, P.total_items));
  const VKernel kern = pick_kernel(gen, P.vt < 31 ? 0 : Ci, P.hr, P.vt < 31, P.oyb);
  if (!kern) return fail(AA_ERR_UNSUPPORTED, "vmma: no kernel instantiation for this shape");
  AA_CUDA_TRY(ensure_smem_attr(kern, th->device, cap_total_bytes));
  kern<<<(unsigned)grid, NT, pl.smem, stream>>>(tmap, P);
  AA_LAUNCH_CHECK("aa_vmma_kernel");
  return AA_OK;
}

}  // namespace aa
