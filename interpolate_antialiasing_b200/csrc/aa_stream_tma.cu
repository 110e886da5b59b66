// aa_stream_tma.cu -- K2, TMA variant: the streaming fused forward kernel with the input row band
// staged in shared memory by bulk asynchronous copies (cp.async.bulk, the 1-D TMA path; SASS UBLKCP)
// tracked by mbarriers, instead of per-thread global loads.
//
// Same algorithm, tables, work split and numerics as aa_stream.cu (read that header first).  What
// changes is who moves the bytes:
//   * warp NWC (the producer) walks the CTA's segments; for every chunk of R input rows it waits for a
//     free stage (empty mbarrier), arms the stage's full mbarrier with the byte count and issues one
//     cp.async.bulk per row (a contiguous, 16-byte aligned strip of the row) -- one elected lane, no
//     registers, up to STAGES*R rows in flight per CTA regardless of what the consumer warps are doing
//     (the horizontal phase no longer drains the memory pipe);
//   * the NWC consumer warps wait on the full mbarrier, read their VEC elements per row with one
//     conflict-free 128-/64-bit LDS, run the vertical FMAs exactly as the LDG variant, and release the
//     stage (one arrive per warp on the empty mbarrier).
// Consumer-only synchronisation uses named barrier 1 so the producer warp never takes part.
#include "aa_stream_common.cuh"

namespace aa {
using namespace stream_detail;
namespace {

constexpr int NWC = 8;             // consumer warps
constexpr int NTC = NWC * 32;      // consumer threads
constexpr int NT = NTC + 32;       // + one producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
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
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }

// staged-row reads (raw registers; expand() converts uint8 just before the FMAs)
template <typename in_t, int VEC> struct SLoad;
template <> struct SLoad<float, 4> {
  static __device__ __forceinline__ void ld(const unsigned char* p, float (&v)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
};
template <> struct SLoad<uint8_t, 8> {
  static __device__ __forceinline__ void ld(const unsigned char* p, uint32_t (&r)[2]) {
    const uint2 q = *reinterpret_cast<const uint2*>(p);
    r[0] = q.x; r[1] = q.y;
  }
};

// A accumulators per element, VEC elements per thread, R rows per stage, STAGES stages.
template <int A, int VEC, typename in_t, int R, int STAGES, int MINB>
__global__ void __launch_bounds__(NT, MINB) aa_stream_tma_kernel(const SParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int RPT = 4;
  constexpr int VW = NTC * VEC;  // row pitch of Vs in floats (== P.vw)
  constexpr int RS4 = (A + 1 + 3) / 4;
  using RawT = typename Raw<in_t, VEC>::T;
  constexpr int RN = Raw<in_t, VEC>::N;
  constexpr int ES = (int)sizeof(in_t);
  // layout: [stages][R][in_pitch] staged input | Vs[vr][vw] | Ws[strip_ox][Kw] | sxmin | sxsize | mbarriers
  unsigned char* stage_base = smem_raw;
  float* Vs = reinterpret_cast<float*>(stage_base + (size_t)STAGES * R * P.in_pitch);
  float2* Wp = reinterpret_cast<float2*>(Vs + (size_t)P.vr * P.vw);                   // [pairs][kp]
  int4* pinfo = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(Wp) + P.wtab_bytes + 15) & ~(uintptr_t)15);  // [pairs * Ci]
  uint64_t* bars = reinterpret_cast<uint64_t*>(pinfo + (P.pairs ? (size_t)((P.strip_ox + 1) / 2) : (size_t)P.strip_ox) * P.Ci);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  __shared__ int s_bad;  // some consumer stored a non-finite value
  if (t == 0) {
    s_bad = 0;
    for (int i = 0; i < STAGES; i++) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, NWC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int Ci = P.Ci;
  const int64_t oH = P.oH;
  const int64_t stride_h = P.lin.stride_h;
  const int64_t u_begin = P.total_units * (int64_t)blockIdx.x / gridDim.x;
  const int64_t u_end = P.total_units * (int64_t)(blockIdx.x + 1) / gridDim.x;
  int stage = 0;
  uint32_t phase = 0;

  if (warp == NWC) {
    // =========================== producer: one elected lane issues the bulk copies ===============
    if (lane == 0) {
      for (int64_t u = u_begin; u < u_end;) {
        const int64_t col = u / oH;
        const int oyA = (int)(u - col * oH);
        const int64_t seg_end = min(u_end, (col + 1) * oH);
        const int oyB = oyA + (int)(seg_end - u);
        const int64_t plane = col / P.n_strips;
        const int s = (int)(col - plane * P.n_strips);
        const int ox0 = s * P.strip_ox;
        const int ox1 = min((int)P.oW, ox0 + P.strip_ox);
        const int fl0 = (__ldg(P.xmin_w + ox0) * Ci) & ~(P.aln - 1);
        const int fl_end = (__ldg(P.xmin_w + ox1 - 1) + __ldg(P.xsize_w + ox1 - 1)) * Ci;
        const uint32_t row_bytes = (uint32_t)(((fl_end - fl0) * ES + 15) & ~15);
        const int64_t yA = __ldg(P.xmin_h + oyA);
        const int64_t yB = (int64_t)__ldg(P.xmin_h + oyB - 1) + __ldg(P.xsize_h + oyB - 1);
        const in_t* src = (const in_t*)P.in + (plane / P.lin.Cp) * P.lin.stride_n + (plane % P.lin.Cp) * P.lin.stride_p + fl0 + yA * stride_h;
        for (int64_t y = yA; y < yB; y += R) {
          const int n = (int)min((int64_t)R, yB - y);
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          mbar_expect_tx(full0 + 8 * stage, row_bytes * n);
          const uint32_t dst = smem_u32(stage_base + (size_t)stage * R * P.in_pitch);
          for (int i = 0; i < n; i++) bulk_g2s(dst + i * P.in_pitch, src + i * stride_h, row_bytes, full0 + 8 * stage);
          src += (int64_t)R * stride_h;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        u = seg_end;
      }
    }
    return;
  }

  // =============================== consumers ========================================================
  constexpr bool CHK = sizeof(in_t) == 4;  // float input can carry NaN/Inf (aa_common.cuh: aa_exact_region)
  constexpr int vw = VW;
  int cur_strip = -1, strip_fl0 = 0, strip_npc = 0;
  HRole role = {0, 1, 0, 1};
  for (int64_t u = u_begin; u < u_end;) {
    const int64_t col = u / oH;
    const int oyA = (int)(u - col * oH);
    const int64_t seg_end = min(u_end, (col + 1) * oH);
    const int oyB = oyA + (int)(seg_end - u);
    const int64_t plane = col / P.n_strips;
    const int s = (int)(col - plane * P.n_strips);
    const int ox0 = s * P.strip_ox;
    const int ox1 = min((int)P.oW, ox0 + P.strip_ox);
    if (s != cur_strip) {
      consumer_sync();
      strip_setup(P, t, NTC, ox0, ox1, Wp, pinfo, &strip_fl0, &strip_npc);
      role = hphase_role(t, NTC, strip_npc);
      cur_strip = s;
      consumer_sync();
    }
    const int fl0 = strip_fl0;                                                              // first flat element of the strip
    const int fl_end = (__ldg(P.xmin_w + ox1 - 1) + __ldg(P.xsize_w + ox1 - 1)) * Ci;       // one past the last
    const bool valid = fl0 + VEC * t < fl_end;
    const int64_t yA = __ldg(P.xmin_h + oyA);
    const int64_t yB = (int64_t)__ldg(P.xmin_h + oyB - 1) + __ldg(P.xsize_h + oyB - 1);
    const float4* rp = reinterpret_cast<const float4*>(P.slot_h) + yA * RS4;
    const int64_t op = (plane / P.lout.Cp) * P.lout.stride_n + (plane % P.lout.Cp) * P.lout.stride_p;  // element offset of the plane
    float* vdst = Vs + VEC * t;
    const unsigned char* my_in = stage_base + (size_t)VEC * ES * t;  // + stage*R*in_pitch + i*in_pitch

    float acc[A][VEC];
#pragma unroll
    for (int a = 0; a < A; a++)
#pragma unroll
      for (int i = 0; i < VEC; i++) acc[a][i] = 0.f;
    int gbase = oyA;
    int cnt = 0;

    auto row = [&](const RawT (&raw)[RN], const float4 (&rq)[RS4]) {
      const float* rw = reinterpret_cast<const float*>(rq);
      float v[VEC];
      expand<VEC>(raw, v);
      vfma<A, VEC>(acc, v, rw);
      const int packed = __float_as_int(rw[A]);
      if (packed >> 24) {
        const int nfl = packed >> 24;
        int o = packed & 0xffffff;
#pragma unroll 1
        for (int k = 0; k < nfl; k++, o++) {
          if (o >= oyA && o < oyB) {
            if (valid) store_vec<VEC>(vdst + (size_t)cnt * vw, acc[0]);
            cnt++;
          }
#pragma unroll
          for (int a = 0; a + 1 < A; a++)
#pragma unroll
            for (int e = 0; e < VEC; e++) acc[a][e] = acc[a + 1][e];
#pragma unroll
          for (int e = 0; e < VEC; e++) acc[A - 1][e] = 0.f;
        }
      }
    };
    auto hphase = [&]() {
      consumer_sync();
      if (hphase_run<RPT, VW, false, false, CHK>(P, Vs, Wp, pinfo, op, strip_npc, role, gbase, cnt)) s_bad = 1;
      consumer_sync();
      gbase += cnt;
      cnt = 0;
    };

    for (int64_t y = yA; y < yB; y += R) {
      const int n = (int)min((int64_t)R, yB - y);
      const unsigned char* sp = my_in + (size_t)stage * R * P.in_pitch;
      mbar_wait(full0 + 8 * stage, phase);
      if (n == R) {
        RawT v[R][RN];
        float4 rq[R][RS4];
#pragma unroll
        for (int i = 0; i < R; i++) SLoad<in_t, VEC>::ld(sp + i * P.in_pitch, v[i]);
#pragma unroll
        for (int i = 0; i < R; i++)
#pragma unroll
          for (int q = 0; q < RS4; q++) rq[i][q] = __ldg(rp + i * RS4 + q);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);  // staged rows are in registers: release the stage
#pragma unroll
        for (int i = 0; i < R; i++) row(v[i], rq[i]);
      } else {
        for (int i = 0; i < n; i++) {
          RawT v[RN];
          float4 rq[RS4];
          SLoad<in_t, VEC>::ld(sp + i * P.in_pitch, v);
#pragma unroll
          for (int q = 0; q < RS4; q++) rq[q] = __ldg(rp + i * RS4 + q);
          row(v, rq);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);
      }
      rp += R * RS4;
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
      if (cnt >= P.tg || (y + R >= yB && cnt > 0)) hphase();
    }
    u = seg_end;
  }
  if constexpr (CHK) {
    if (t == 0 && s_bad && P.redo) redo_push(P.redo, u_begin, u_end, 0, 0, 0, 0);
  }
}

template <int A, int VEC, typename in_t>
int launch_tma_cfg(SParams& P, const StreamTables& T, int device, cudaStream_t stream) {
  constexpr int R = 4, STAGES = 4;
  constexpr int MINB = 2;
  constexpr int ES = (int)sizeof(in_t);
  auto kern = aa_stream_tma_kernel<A, VEC, in_t, R, STAGES, MINB>;
  const PlanKey key{T.key_h, T.key_w, P.Ci, (T.dir << 29) | 0x10000 | (A << 8) | (VEC << 2) | ES % 4};
  Plan pl;
  if (!plan_lookup(key, &pl)) {
    int rc = plan_stream(P, T, NTC * VEC, 16 / ES, VEC, R, 4);
    if (rc != AA_OK) return rc;
    P.pad = 0;  // the padded row buffer exists in the plain-load variant only
    P.in_pitch = (P.vw * ES + 15) & ~15;
    const size_t smem_ = (size_t)STAGES * R * P.in_pitch + sizeof(float) * (size_t)P.vr * P.vw + strip_table_bytes(P) + 8 + 16 * STAGES;
    if (smem_ > 200 * 1024) return fail(AA_ERR_UNSUPPORTED, "stream/tma: shared memory plan too large");
    AA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int occ = 0, sms = 0;
    AA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem_));
    AA_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (occ < 1) return fail(AA_ERR_UNSUPPORTED, "stream/tma: kernel does not fit on an SM");
    pl = plan_from(P, smem_, occ * sms);
    plan_store(key, pl);
  }
  plan_apply(P, pl);
  const size_t smem = pl.smem;
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(pl.max_grid, P.total_units / 4));
  kern<<<(unsigned)grid, NT, smem, stream>>>(P);
  AA_LAUNCH_CHECK("aa_stream_tma_kernel");
  return AA_OK;
}

template <int A>
int launch_tma_A(SParams& P, int in_dtype, const StreamTables& T, int device, cudaStream_t stream) {
  if (in_dtype == AA_F32) return launch_tma_cfg<A, 4, float>(P, T, device, stream);
  return launch_tma_cfg<A, 8, uint8_t>(P, T, device, stream);
}

}  // namespace

int launch_stream_tma(SParams& P, int A, int in_dtype, const StreamTables& T, int device, cudaStream_t stream) {
  if (P.epi.generic()) return fail(AA_ERR_UNSUPPORTED, "stream/tma: generic epilogue not instantiated");
  // cp.async.bulk needs 16-byte aligned global addresses and sizes: base, plane and row strides
  const int es = in_dtype == AA_F32 ? 4 : 1;
  const int64_t al = 16 / es;
  if (((uintptr_t)P.in) % 16 || (P.lin.stride_h % al) || (P.lin.stride_n % al) || (P.lin.Cp > 1 && P.lin.stride_p % al))
    return fail(AA_ERR_UNSUPPORTED, "stream/tma: input rows are not 16-byte aligned");
  switch (A) {
    case 3: return launch_tma_A<3>(P, in_dtype, T, device, stream);
    case 4: return launch_tma_A<4>(P, in_dtype, T, device, stream);
    case 5: return launch_tma_A<5>(P, in_dtype, T, device, stream);
    case 6: return launch_tma_A<6>(P, in_dtype, T, device, stream);
  }
  return fail(AA_ERR_UNSUPPORTED, "stream/tma: unsupported accumulator count");
}

}  // namespace aa
