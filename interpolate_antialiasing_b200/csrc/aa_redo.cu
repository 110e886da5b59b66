// aa_redo.cu -- the drain kernel behind every fast float launch (sm_100a): exact NaN/Inf placement.
//
// The reference only touches the taps j < xsize of a window
// (/root/reference/step_two_dot_two/aa_interpolation_impl.h:73-85), the fast kernels also multiply a few
// zero-weight neighbours (see aa_common.cuh, "non-finite inputs").  A CTA of a fast kernel that stored a
// non-finite value appends its region to the stream's RedoList; this kernel re-evaluates the listed regions with
// in-window taps only.  With an empty list -- every finite image -- each CTA reads one counter and exits.
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "aa_common.cuh"

namespace aa {
namespace {

constexpr int NT = 256;
constexpr int ROWS = 32;  // overflow mode: rows per work item

template <bool GEN>
__global__ void __launch_bounds__(NT) aa_redo_kernel(const __grid_constant__ RedoParams R) {
  RedoList* L = R.list;
  // launched with programmatic stream serialization: the CTAs may already be resident while the fast kernel's last CTAs
  // run (it triggers at its start); this returns once that kernel has completed and its writes are visible
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned int n = L->count;
  if (n == 0) return;  // nothing was non-finite (the list is left as it is: all zero)
  const int tid = threadIdx.x;
  const float* in = (const float*)R.in;
  auto plane_in = [&](int64_t p) { return in + (p / R.lin.Cp) * R.lin.stride_n + (p % R.lin.Cp) * R.lin.stride_p; };
  auto plane_out = [&](int64_t p) { return (p / R.lout.Cp) * R.lout.stride_n + (p % R.lout.Cp) * R.lout.stride_p; };
  auto region = [&](int64_t p, int oy0, int oy1, int of0, int of1) {
    aa_exact_region<GEN, float>(plane_in(p), R.lin.stride_h, R.Ci, R.T, R.out, plane_out(p), R.lout.stride_h, R.epi, oy0, oy1, of0, of1,
                                tid, NT);
  };
  if (L->overflow) {
    // more dirty regions than entries: the whole output, in items of ROWS rows of one plane
    const int64_t chunks = (R.out_h + ROWS - 1) / ROWS;
    for (int64_t it = blockIdx.x; it < R.lin.planes * chunks; it += gridDim.x) {
      const int64_t p = it / chunks;
      const int oy0 = (int)(it - p * chunks) * ROWS;
      region(p, oy0, min(R.out_h, oy0 + ROWS), 0, R.out_wf);
    }
  } else {
    for (unsigned int i = blockIdx.x; i < n; i += gridDim.x) {
      const RedoEntry e = L->e[i];
      if (e.b < 0) {
        region(e.a, e.oy0, e.oy1, e.of0, e.of1);
      } else {
        // the streaming kernel's unit range: the same walk over (plane, strip, row range) segments as aa_stream.cu
        const int64_t oH = R.out_h;
        for (int64_t u = e.a; u < e.b;) {
          const int64_t col = u / oH;
          const int oyA = (int)(u - col * oH);
          const int64_t seg_end = min(e.b, (col + 1) * oH);
          const int64_t p = col / R.n_strips;
          const int s = (int)(col - p * R.n_strips);
          const int ox0 = s * R.strip_ox, ox1 = min(R.out_wf / R.Ci, ox0 + R.strip_ox);
          region(p, oyA, oyA + (int)(seg_end - u), ox0 * R.Ci, ox1 * R.Ci);
          u = seg_end;
        }
      }
    }
  }
  // the last CTA to finish empties the list for the next launch on this stream (every CTA has read it by then)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(&L->done, 1u) == gridDim.x - 1) {
      L->count = 0u;
      L->overflow = 0u;
      L->done = 0u;
    }
  }
}

thread_local bool t_redo_enabled = true;
std::mutex g_mu;
std::map<std::pair<int, cudaStream_t>, RedoList*> g_lists;  // the list of a (device, stream)
struct Chunk { int device; RedoList* base; int used; };
std::vector<Chunk> g_chunks;                                  // lists are carved out of chunks of kChunk
constexpr int kChunk = 32;

}  // namespace

// Lists come from per-device chunks that are zeroed when they are created and are left zeroed by every drain, so
// handing one to a new stream is pure host bookkeeping: legal inside a CUDA-graph capture as long as the device has a
// chunk with a free list (any earlier eager call or aa_warm_tables makes the first one).
int redo_list(int device, cudaStream_t stream, RedoList** out) {
  *out = nullptr;  // AA_FLAG_ASSUME_FINITE: no list -- the kernels then report nothing, and nothing stale is left for a later drain
  if (!t_redo_enabled) return AA_OK;
  // steady state: the calling thread asks for the same (device, stream) as last time -- no lock, no map lookup (lists are never freed)
  thread_local int last_dev = -1;
  thread_local cudaStream_t last_stream = nullptr;
  thread_local RedoList* last_list = nullptr;
  if (last_list && last_dev == device && last_stream == stream) { *out = last_list; return AA_OK; }
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_lists.find({device, stream});
  if (it != g_lists.end()) {
    *out = it->second;
    last_dev = device; last_stream = stream; last_list = it->second;
    return AA_OK;
  }
  Chunk* c = nullptr;
  for (auto& ch : g_chunks)
    if (ch.device == device && ch.used < kChunk) { c = &ch; break; }
  if (!c) {
    RedoList* base = nullptr;
    AA_CUDA_TRY(cudaMalloc(&base, sizeof(RedoList) * kChunk));
    const cudaError_t e = cudaMemset(base, 0, sizeof(RedoList) * kChunk);
    if (e != cudaSuccess) { cudaFree(base); return cuda_fail(e, "cudaMemset(redo lists)"); }
    g_chunks.push_back(Chunk{device, base, 0});
    c = &g_chunks.back();
  }
  RedoList* L = c->base + c->used++;
  g_lists[{device, stream}] = L;
  *out = L;
  last_dev = device; last_stream = stream; last_list = L;
  return AA_OK;
}

// aa_clear_table_cache: the lists are NOT freed.  A call on another thread may hold a list pointer it has not launched with
// yet (tables are reference-counted for exactly that case; lists are plain device memory), and there is nothing to gain:
// a list is 64 KB and their number is bounded by the number of distinct (device, stream) pairs the process ever used.
void redo_clear() {}

bool redo_set_enabled(bool on) {  // per calling thread: AA_FLAG_ASSUME_FINITE switches the drain launch off for one call
  const bool was = t_redo_enabled;
  t_redo_enabled = on;
  return was;
}

int launch_redo(const RedoParams& R, int device, cudaStream_t stream) {
  if (!R.list) return AA_OK;
  static std::atomic<int> sms[64];  // SM count per device (0 = not asked yet; racing first calls ask twice, harmlessly)
  int n = (device >= 0 && device < 64) ? sms[device].load(std::memory_order_relaxed) : 0;
  if (!n) {
    AA_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
    if (device >= 0 && device < 64) sms[device].store(n, std::memory_order_relaxed);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)n);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // overlap this launch with the fast kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AA_CUDA_TRY(R.epi.generic() ? cudaLaunchKernelEx(&cfg, aa_redo_kernel<true>, R) : cudaLaunchKernelEx(&cfg, aa_redo_kernel<false>, R));
  AA_LAUNCH_CHECK("aa_redo_kernel");
  return AA_OK;
}

}  // namespace aa
