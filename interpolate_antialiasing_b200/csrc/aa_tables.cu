// aa_tables.cu -- K1: the per-axis index/weight table builder as a tiny sm_100a kernel, plus the
// per-device table cache.
//
// Replaces HelperInterpBase::_compute_indices_weights_aa and the three filter helpers
// (/root/reference/step_two_dot_two/aa_interpolation_impl.h:194-281, :292-300, :367-372, :410-424)
// and torch's area_pixel_compute_scale (ATen/native/UpSample.h).  The reference evaluates the
// tables in mixed fp32/fp64 by C++ promotion rules on x86-64 without FMA; every operation below is
// an explicit round-to-nearest intrinsic (__fmul_rn, __dadd_rn, ...) so nvcc can neither contract
// nor reassociate, and the integer tables AND the weights come out bit-identical
// (tests/test_tables_gpu.py).  Compile this TU without --use_fast_math.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>

#include "aa_common.cuh"

namespace aa {

// ------------------------------------------------------------------------------------------------
// scalar recipe shared by host (launch planning) and device (the tables)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline double base_half(int filter) {  // interp_size * 0.5  (:208)
  return filter == AA_FILTER_BOX ? 0.5 : (filter == AA_FILTER_TRIANGLE ? 1.0 : 2.0);
}

// host versions: plain IEEE float/double operators (no a*b+c pattern exists in them)
static float h_scale_f32(int64_t in, int64_t out, int align) {
  if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
  return (float)in / (float)out;
}
static double h_scale_f64(int64_t in, int64_t out, int align) {
  if (align) return out > 1 ? (double)(in - 1) / (double)(out - 1) : 0.0;
  return (double)in / (double)out;
}
static float h_support_f32(float scale, int filter) {
  return (scale >= 1.0f) ? (float)(base_half(filter) * (double)scale) : (float)base_half(filter);
}
static double h_support_f64(double scale, int filter) {
  return (scale >= 1.0) ? base_half(filter) * scale : base_half(filter);
}

int host_interp_size(int64_t in, int64_t out, int filter, int align, int dtype) {
  if (dtype == AA_F64) {
    double s = h_support_f64(h_scale_f64(in, out, align), filter);
    return (int)ceilf((float)s) * 2 + 1;  // ceilf() of a double converts to float first (:210)
  }
  float s = h_support_f32(h_scale_f32(in, out, align), filter);
  return (int)ceilf(s) * 2 + 1;
}

// ------------------------------------------------------------------------------------------------
// device filters, explicit rounding
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float filt_f32(int filter, float x) {
  if (filter == AA_FILTER_BOX) return (x > -0.5f && x <= 0.5f) ? 1.0f : 0.0f;  // :367-372
  if (x < 0.0f) x = -x;
  if (filter == AA_FILTER_TRIANGLE) {  // :292-300, `1.0 - x` in double
    return x < 1.0f ? __double2float_rn(__dsub_rn(1.0, (double)x)) : 0.0f;
  }
  if (x < 1.0f) {  // :417 all double: ((a+2)x - (a+3)) x x + 1, a = -0.5
    double xd = (double)x;
    double t = __dsub_rn(__dmul_rn(1.5, xd), 2.5);
    t = __dmul_rn(t, xd);
    t = __dmul_rn(t, xd);
    return __double2float_rn(__dadd_rn(t, 1.0));
  }
  if (x < 2.0f) {  // :420 polynomial in float, final *a in double
    float p = __fsub_rn(x, 5.0f);
    p = __fmul_rn(p, x);
    p = __fadd_rn(p, 8.0f);
    p = __fmul_rn(p, x);
    p = __fsub_rn(p, 4.0f);
    return __double2float_rn(__dmul_rn((double)p, -0.5));
  }
  return 0.0f;
}

__device__ __forceinline__ double filt_f64(int filter, double x) {
  if (filter == AA_FILTER_BOX) return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0;
  if (x < 0.0) x = -x;
  if (filter == AA_FILTER_TRIANGLE) return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
  if (x < 1.0) {
    double t = __dsub_rn(__dmul_rn(1.5, x), 2.5);
    t = __dmul_rn(t, x);
    t = __dmul_rn(t, x);
    return __dadd_rn(t, 1.0);
  }
  if (x < 2.0) {
    double p = __dsub_rn(x, 5.0);
    p = __dmul_rn(p, x);
    p = __dadd_rn(p, 8.0);
    p = __dmul_rn(p, x);
    p = __dsub_rn(p, 4.0);
    return __dmul_rn(p, -0.5);
  }
  return 0.0;
}

struct TableMeta {
  int xsize_max;
  int kt_max;
  int monotone;
  int pad;
};

// One thread per output index.  fp32 instantiation of :194-281.
__global__ void aa_tables_fwd_f32(int64_t in, int64_t out, int filter, int align, int K,
                                  int32_t* __restrict__ xmin_o, int32_t* __restrict__ xsize_o,
                                  float* __restrict__ w, TableMeta* meta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out) return;
  float scale;
  if (align) scale = out > 1 ? __fdiv_rn(__ll2float_rn(in - 1), __ll2float_rn(out - 1)) : 0.0f;
  else scale = __fdiv_rn(__ll2float_rn(in), __ll2float_rn(out));
  const double bh = base_half(filter);
  const float support = (scale >= 1.0f) ? __double2float_rn(__dmul_rn(bh, (double)scale)) : (float)bh;  // :208-209
  const float invscale = (scale >= 1.0f) ? __double2float_rn(__ddiv_rn(1.0, (double)scale)) : 1.0f;     // :242
  const float center = __double2float_rn(__dmul_rn((double)scale, __dadd_rn((double)i, 0.5)));         // :253
  long long xmin = __double2ll_rz(__dadd_rn((double)__fsub_rn(center, support), 0.5));                  // :254
  if (xmin < 0) xmin = 0;
  long long xmax = __double2ll_rz(__dadd_rn((double)__fadd_rn(center, support), 0.5));                  // :255-257
  if (xmax > in) xmax = in;
  xmax -= xmin;
  xmin_o[i] = (int32_t)xmin;
  xsize_o[i] = (int32_t)xmax;
  float* wr = w + i * K;
  float total = 0.0f;
  long long j = 0;
  for (; j < xmax && j < K; j++) {
    float d = __fsub_rn(__ll2float_rn(j + xmin), center);                                               // :266
    float arg = __double2float_rn(__dmul_rn(__dadd_rn((double)d, 0.5), (double)invscale));
    float wj = filt_f32(filter, arg);
    wr[j] = wj;
    total = __fadd_rn(total, wj);                                                                      // :268
  }
  const long long nt = j;
  if (total != 0.0f)
    for (j = 0; j < nt; j++) wr[j] = __fdiv_rn(wr[j], total);                                           // :270-274
  for (j = nt; j < K; j++) wr[j] = 0.0f;                                                                // :276-278
  atomicMax(&meta->xsize_max, (int)xmax);
}

__global__ void aa_tables_fwd_f64(int64_t in, int64_t out, int filter, int align, int K,
                                  int32_t* __restrict__ xmin_o, int32_t* __restrict__ xsize_o,
                                  double* __restrict__ w, TableMeta* meta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out) return;
  double scale;
  if (align) scale = out > 1 ? __ddiv_rn((double)(in - 1), (double)(out - 1)) : 0.0;
  else scale = __ddiv_rn((double)in, (double)out);
  const double bh = base_half(filter);
  const double support = (scale >= 1.0) ? __dmul_rn(bh, scale) : bh;
  const double invscale = (scale >= 1.0) ? __ddiv_rn(1.0, scale) : 1.0;
  const double center = __dmul_rn(scale, __dadd_rn((double)i, 0.5));
  long long xmin = __double2ll_rz(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  long long xmax = __double2ll_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in) xmax = in;
  xmax -= xmin;
  xmin_o[i] = (int32_t)xmin;
  xsize_o[i] = (int32_t)xmax;
  double* wr = w + i * K;
  double total = 0.0;
  long long j = 0;
  for (; j < xmax && j < K; j++) {
    // (j + xmin - center + 0.5) * invscale, all double, left to right
    double arg = __dmul_rn(__dadd_rn(__dsub_rn((double)(j + xmin), center), 0.5), invscale);
    double wj = filt_f64(filter, arg);
    wr[j] = wj;
    total = __dadd_rn(total, wj);
  }
  const long long nt = j;
  if (total != 0.0)
    for (j = 0; j < nt; j++) wr[j] = __ddiv_rn(wr[j], total);
  for (j = nt; j < K; j++) wr[j] = 0.0;
  atomicMax(&meta->xsize_max, (int)xmax);
}

// One thread per input index x: the transposed (adjoint) tables.  Also verifies the monotonicity
// the contiguous-range argument relies on.
template <typename T>
__global__ void aa_tables_adj(int64_t in, int64_t out, int K, int KT, const int32_t* __restrict__ xmin,
                              const int32_t* __restrict__ xsize, const T* __restrict__ w,
                              int32_t* __restrict__ omin_o, int32_t* __restrict__ osize_o,
                              T* __restrict__ wT, TableMeta* meta) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x > 0 && x < out) {
    if (xmin[x] < xmin[x - 1] || xmin[x] + xsize[x] < xmin[x - 1] + xsize[x - 1]) meta->monotone = 0;
  }
  if (x >= in) return;
  // omin = first o with xmin[o]+xsize[o] > x ; oend = first o with xmin[o] > x
  int64_t lo = 0, hi = out;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)xmin[mid] + xsize[mid] > x) hi = mid; else lo = mid + 1;
  }
  const int64_t omin = lo;
  lo = omin; hi = out;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)xmin[mid] > x) hi = mid; else lo = mid + 1;
  }
  const int64_t osz = lo - omin;
  omin_o[x] = (int32_t)omin;
  osize_o[x] = (int32_t)osz;
  atomicMax(&meta->kt_max, (int)osz);
  T* row = wT + x * KT;
  for (int k = 0; k < KT; k++) {
    T v = (T)0;
    if (k < osz) {
      int64_t o = omin + k;
      int64_t j = x - xmin[o];
      if (j >= 0 && j < K) v = w[o * K + j];
    }
    row[k] = v;
  }
}

// One thread per input row y: the records used by the streaming kernel's vertical pass.
// Record = A weights wT[y][0..A) (k-th weight belongs to output omin[y]+k, the k-th OLDEST output row
// still open at y), followed by (first_flush_o | nflush << 24): the `nflush` oldest open outputs,
// starting at first_flush_o = omin[y], have y as the LAST row of their window.  Because the outputs
// that end at y are exactly the ones missing from row y+1's range, omin[y+1] = omin[y] + nflush.
__global__ void aa_tables_slots(int64_t in, int A, int RS, int KT, const int32_t* __restrict__ xmin,
                                const int32_t* __restrict__ xsize, const int32_t* __restrict__ omin,
                                const int32_t* __restrict__ osize, const float* __restrict__ wT,
                                float* __restrict__ slot) {
  int64_t y = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= in) return;
  float* rec = slot + y * RS;
  for (int a = 0; a < RS; a++) rec[a] = 0.0f;
  const int o0 = omin[y], n = osize[y];
  int nflush = 0;
  for (int k = 0; k < n && k < A; k++) {
    const int o = o0 + k;
    rec[k] = wT[y * KT + k];
    if ((int64_t)xmin[o] + xsize[o] - 1 == y) nflush++;
  }
  rec[A] = __int_as_float(o0 | (nflush << 24));
}

// Adjoint direction: one thread per forward-output index o (= a streamed grad_out row).  The open
// grad_in rows at o are [xmin[o], xmin[o]+xsize[o]) with the forward weights; the ones that no later o
// covers (y < xmin[o+1]) finish here.
__global__ void aa_tables_slots_adj(int64_t out, int A, int RS, int K, const int32_t* __restrict__ xmin,
                                    const int32_t* __restrict__ xsize, const float* __restrict__ w,
                                    float* __restrict__ slot) {
  int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out) return;
  float* rec = slot + o * RS;
  for (int a = 0; a < RS; a++) rec[a] = 0.0f;
  const int x0 = xmin[o], n = xsize[o];
  for (int k = 0; k < n && k < A; k++) rec[k] = w[o * K + k];
  const int end = x0 + n;
  const int nxt = (o + 1 < out) ? xmin[o + 1] : end;
  const int nflush = (nxt < end ? nxt : end) - x0;
  rec[A] = __int_as_float(x0 | (nflush << 24));
}

// ------------------------------------------------------------------------------------------------
// cache
// ------------------------------------------------------------------------------------------------
AxisTables::~AxisTables() {
  // Device memory is released with the owning context; freeing here is best effort.
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess) {
    cudaSetDevice(device);
    if (block) cudaFree(block);
    if (slot) cudaFree(slot);
    if (slot_adj) cudaFree(slot_adj);
    cudaSetDevice(cur);
  }
}

namespace {
using Key = std::tuple<int, int64_t, int64_t, int, int, int>;
std::mutex g_mu;
std::map<Key, std::shared_ptr<AxisTables>> g_cache;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int build_tables(AxisTables* t, cudaStream_t stream) {
  const int64_t in = t->in, out = t->out;
  const bool f64 = t->dtype == AA_F64;
  t->scale_f = h_scale_f32(in, out, t->align);
  t->support_f = h_support_f32(t->scale_f, t->filter);
  t->scale_d = h_scale_f64(in, out, t->align);
  t->support_d = h_support_f64(t->scale_d, t->filter);
  t->K = host_interp_size(in, out, t->filter, t->align, t->dtype);
  {
    // generous bound on how many outputs can cover one input index
    const double sc = f64 ? t->scale_d : (double)t->scale_f;
    const double sp = f64 ? t->support_d : (double)t->support_f;
    double b = sc > 0.0 ? ceil((2.0 * sp + 2.0) / sc) + 2.0 : (double)out;
    if (b > (double)out) b = (double)out;
    if (b < 1.0) b = 1.0;
    t->KT = (int)b;
  }
  const size_t es = f64 ? 8 : 4;
  size_t off = 0;
  const size_t o_xmin = off;  off = align_up(off + sizeof(int32_t) * out, 256);
  const size_t o_xsize = off; off = align_up(off + sizeof(int32_t) * out, 256);
  const size_t o_w = off;     off = align_up(off + es * out * t->K, 256);
  const size_t o_omin = off;  off = align_up(off + sizeof(int32_t) * in, 256);
  const size_t o_osize = off; off = align_up(off + sizeof(int32_t) * in, 256);
  const size_t o_wT = off;    off = align_up(off + es * in * t->KT, 256);
  const size_t o_meta = off;  off = align_up(off + sizeof(TableMeta), 256);
  AA_CUDA_TRY(cudaMalloc(&t->block, off));
  char* base = (char*)t->block;
  t->xmin = (int32_t*)(base + o_xmin);
  t->xsize = (int32_t*)(base + o_xsize);
  t->w = base + o_w;
  t->omin = (int32_t*)(base + o_omin);
  t->osize = (int32_t*)(base + o_osize);
  t->wT = base + o_wT;
  TableMeta* meta = (TableMeta*)(base + o_meta);
  TableMeta init = {0, 0, 1, 0};
  AA_CUDA_TRY(cudaMemcpyAsync(meta, &init, sizeof(init), cudaMemcpyHostToDevice, stream));
  const int NT = 128;
  const unsigned gb_out = (unsigned)((out + NT - 1) / NT);
  const int64_t nmax = in > out ? in : out;
  const unsigned gb_adj = (unsigned)((nmax + NT - 1) / NT);
  if (f64) {
    aa_tables_fwd_f64<<<gb_out, NT, 0, stream>>>(in, out, t->filter, t->align, t->K, t->xmin, t->xsize,
                                                  (double*)t->w, meta);
    AA_LAUNCH_CHECK("aa_tables_fwd_f64");
    aa_tables_adj<double><<<gb_adj, NT, 0, stream>>>(in, out, t->K, t->KT, t->xmin, t->xsize,
                                                      (const double*)t->w, t->omin, t->osize, (double*)t->wT, meta);
    AA_LAUNCH_CHECK("aa_tables_adj<double>");
  } else {
    aa_tables_fwd_f32<<<gb_out, NT, 0, stream>>>(in, out, t->filter, t->align, t->K, t->xmin, t->xsize,
                                                  (float*)t->w, meta);
    AA_LAUNCH_CHECK("aa_tables_fwd_f32");
    aa_tables_adj<float><<<gb_adj, NT, 0, stream>>>(in, out, t->K, t->KT, t->xmin, t->xsize,
                                                     (const float*)t->w, t->omin, t->osize, (float*)t->wT, meta);
    AA_LAUNCH_CHECK("aa_tables_adj<float>");
  }
  TableMeta h = {};
  t->h_xmin.resize(out);
  t->h_xsize.resize(out);
  t->h_omin.resize(in);
  t->h_osize.resize(in);
  AA_CUDA_TRY(cudaMemcpyAsync(&h, meta, sizeof(h), cudaMemcpyDeviceToHost, stream));
  AA_CUDA_TRY(cudaMemcpyAsync(t->h_xmin.data(), t->xmin, sizeof(int32_t) * out, cudaMemcpyDeviceToHost, stream));
  AA_CUDA_TRY(cudaMemcpyAsync(t->h_xsize.data(), t->xsize, sizeof(int32_t) * out, cudaMemcpyDeviceToHost, stream));
  AA_CUDA_TRY(cudaMemcpyAsync(t->h_omin.data(), t->omin, sizeof(int32_t) * in, cudaMemcpyDeviceToHost, stream));
  AA_CUDA_TRY(cudaMemcpyAsync(t->h_osize.data(), t->osize, sizeof(int32_t) * in, cudaMemcpyDeviceToHost, stream));
  AA_CUDA_TRY(cudaStreamSynchronize(stream));  // cache miss only
  t->xsize_max = h.xsize_max;
  t->kt_max = h.kt_max;
  t->monotone = h.monotone;
  if (h.kt_max > t->KT)
    return fail(AA_ERR_INVALID, "internal: adjoint pitch bound too small (kt_max " + std::to_string(h.kt_max) +
                                    " > KT " + std::to_string(t->KT) + ")");
  if (!h.monotone) return fail(AA_ERR_INVALID, "internal: window tables are not monotone");
  return AA_OK;
}
}  // namespace

int get_axis_tables(int device, int64_t in, int64_t out, int filter, int align, int dtype,
                    cudaStream_t stream, std::shared_ptr<AxisTables>* result) {
  if (in <= 0 || out <= 0) return fail(AA_ERR_INVALID, "table sizes must be positive");
  if (in >= (1ll << 31) || out >= (1ll << 31)) return fail(AA_ERR_UNSUPPORTED, "axis length >= 2^31");
  if (dtype == AA_U8) dtype = AA_F32;
  Key key(device, in, out, filter, align ? 1 : 0, dtype);
  std::lock_guard<std::mutex> lock(g_mu);
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    // a cached entry is complete (the builder synchronised its stream before publishing it), so a hit
    // needs no stream dependency at all -- which also keeps hits legal under CUDA-graph capture
    *result = it->second;
    return AA_OK;
  }
  auto t = std::make_shared<AxisTables>();
  t->device = device;
  t->in = in;
  t->out = out;
  t->filter = filter;
  t->align = align ? 1 : 0;
  t->dtype = dtype;
  int rc = build_tables(t.get(), stream);
  if (rc != AA_OK) return rc;
  g_cache[key] = t;
  *result = t;
  return AA_OK;
}

int ensure_slot_tables(AxisTables* t, int A, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (t->dtype != AA_F32) return fail(AA_ERR_UNSUPPORTED, "slot tables are float only");
  if (t->slot && t->slot_A == A) return AA_OK;  // complete: the builder synchronised before publishing
  if (t->slot) return fail(AA_ERR_INVALID, "internal: slot tables requested with two different A");
  if (t->out >= (1 << 24)) return fail(AA_ERR_UNSUPPORTED, "streaming path needs out < 2^24");
  const int RS = (A + 1 + 3) / 4 * 4;
  AA_CUDA_TRY(cudaMalloc(&t->slot, sizeof(float) * (size_t)t->in * RS));
  const int NT = 128;
  aa_tables_slots<<<(unsigned)((t->in + NT - 1) / NT), NT, 0, stream>>>(
      t->in, A, RS, t->KT, t->xmin, t->xsize, t->omin, t->osize, (const float*)t->wT, t->slot);
  AA_LAUNCH_CHECK("aa_tables_slots");
  AA_CUDA_TRY(cudaStreamSynchronize(stream));  // first use only
  t->slot_A = A;
  t->slot_RS = RS;
  return AA_OK;
}

namespace stream_detail { void plan_clear(); }

int ensure_slot_tables_adj(AxisTables* t, int A, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (t->dtype != AA_F32) return fail(AA_ERR_UNSUPPORTED, "slot tables are float only");
  if (t->slot_adj && t->slot_adj_A == A) return AA_OK;
  if (t->slot_adj) return fail(AA_ERR_INVALID, "internal: adjoint slot tables requested with two different A");
  if (t->in >= (1 << 24)) return fail(AA_ERR_UNSUPPORTED, "streaming path needs in < 2^24");
  const int RS = (A + 1 + 3) / 4 * 4;
  AA_CUDA_TRY(cudaMalloc(&t->slot_adj, sizeof(float) * (size_t)t->out * RS));
  const int NT = 128;
  aa_tables_slots_adj<<<(unsigned)((t->out + NT - 1) / NT), NT, 0, stream>>>(t->out, A, RS, t->K, t->xmin, t->xsize,
                                                                             (const float*)t->w, t->slot_adj);
  AA_LAUNCH_CHECK("aa_tables_slots_adj");
  AA_CUDA_TRY(cudaStreamSynchronize(stream));  // first use only
  t->slot_adj_A = A;
  t->slot_adj_RS = RS;
  return AA_OK;
}

int clear_table_cache() {
  stream_detail::plan_clear();
  std::lock_guard<std::mutex> lock(g_mu);
  g_cache.clear();
  return AA_OK;
}

}  // namespace aa
