// torch_binding.cpp -- the reference's pybind operator boundary, re-hosted on the C ABI.
//
// Exports the same names with the same positional signatures as
// /root/reference/step_two_dot_two/extension_interpolate.cpp:46-51
//   linear_forward(input, output_size, align_corners)  -> Tensor      (:7-14)
//   cubic_forward(input, output_size, align_corners)   -> Tensor      (:35-42)
//   nearest_forward(input, output_size, align_corners) -> Tensor      (:26-33, box filter)
//   linear_backward(grad_output, output_size, input_size, align_corners) -> Tensor   (:16-24)
// plus what the reference stubs out (test.py:111-116): cubic_backward, nearest_backward, and
// linear_backward_nonaa (the reference's literal, non-antialiased backward arithmetic).
//
// Each forward/backward also takes an optional trailing `scale_factors` ([sh, sw]): the argument the reference's
// ti_upsample_*2d_cpu accept (aa_interpolation_impl.h:735,740-742) but its shim never passes
// (extension_interpolate.cpp:12).  With it, output_size may be None (= floor(in * scale), compute_output_size) and the
// table scale becomes 1/scale_factor (area_pixel_compute_scale), i.e. F.interpolate(scale_factor=s,
// recompute_scale_factor=False).
//
// Semantics kept: output_size = (oH, oW); scale is in/out unless scale_factors is given;
// align_corners only changes the scale; 4-D input; empty batch allowed; memory format of the
// result = input.suggest_memory_format() (aa_interpolation_impl.h:752) / grad_output's
// (aa_interpolation_backward_impl.h:214); errors surface as RuntimeError.  Differences: tensors
// must be CUDA tensors (there is NO CPU fallback), and uint8 inputs are accepted by
// linear/cubic_forward with the caller's `.float()` (test.py:55,67) fused (result is float32).
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include "aa_resize.h"

namespace {

int to_aa_dtype(at::ScalarType t) {
  switch (t) {
    case at::kByte: return AA_U8;
    case at::kFloat: return AA_F32;
    case at::kDouble: return AA_F64;
    default: TORCH_CHECK(false, "\"upsample_generic_Nd\" not implemented for '", c10::toString(t), "'");
  }
}

aa_tensor_desc make_desc(const at::Tensor& t) {
  aa_tensor_desc d;
  d.data = t.numel() ? t.data_ptr() : nullptr;
  d.dtype = to_aa_dtype(t.scalar_type());
  d.device = t.device().index();
  d.n = t.size(0); d.c = t.size(1); d.h = t.size(2); d.w = t.size(3);
  d.stride_n = t.stride(0); d.stride_c = t.stride(1); d.stride_h = t.stride(2); d.stride_w = t.stride(3);
  return d;
}

// same checks and messages as at::native::upsample_2d_common_check (ATen/native/UpSample.h)
void common_check(at::IntArrayRef input_size, at::IntArrayRef output_size) {
  TORCH_CHECK(output_size.size() == 2, "It is expected output_size equals to 2, but got size ", output_size.size());
  TORCH_CHECK(input_size.size() == 4, "It is expected input_size equals to 4, but got size ", input_size.size());
  TORCH_CHECK(input_size[2] > 0 && input_size[3] > 0 && output_size[0] > 0 && output_size[1] > 0,
              "Input and output sizes should be greater than 0, but got input (H: ", input_size[2], ", W: ", input_size[3],
              ") output (H: ", output_size[0], ", W: ", output_size[1], ")");
}

using OptSize = c10::optional<std::vector<int64_t>>;
using OptScales = c10::optional<std::vector<double>>;

// at::native::upsample::compute_output_size (ATen/native/UpSample.h): exactly one of output_size / scale_factors
std::vector<int64_t> resolve_output_size(at::IntArrayRef input_size, const OptSize& output_size, const OptScales& scale_factors) {
  const int64_t spatial = (int64_t)input_size.size() - 2;
  if (output_size.has_value()) {
    TORCH_CHECK(!scale_factors.has_value(), "Must specify exactly one of output_size and scale_factors");
    TORCH_CHECK((int64_t)output_size->size() == spatial, "It is expected output_size equals to ", spatial, ", but got size ", output_size->size());
    return *output_size;
  }
  TORCH_CHECK(scale_factors.has_value(), "Must specify exactly one of output_size and scale_factors");
  TORCH_CHECK((int64_t)scale_factors->size() == spatial, "It is expected scale_factors equals to ", spatial, ", but got size ", scale_factors->size());
  std::vector<int64_t> r;
  for (int64_t i = 0; i < spatial; i++) r.push_back((int64_t)((double)input_size[i + 2] * (*scale_factors)[i]));
  return r;
}
aa_scales to_scales(const OptScales& sf) {
  aa_scales sc = {0.0, 0.0};
  if (sf.has_value() && sf->size() == 2) { sc.scale_h = (*sf)[0]; sc.scale_w = (*sf)[1]; }
  return sc;
}

at::Tensor forward_common(const at::Tensor& input, const OptSize& osize_opt, bool align_corners, int filter, int64_t flags,
                          bool out_u8 = false, const OptScales& scale_factors = c10::nullopt) {
  TORCH_CHECK(input.is_cuda(), "aa_interp_b200: input must be a CUDA tensor (this build has no CPU fallback)");
  TORCH_CHECK(input.dim() == 4, "It is expected input_size equals to 4, but got size ", input.dim());
  const std::vector<int64_t> osize_v = resolve_output_size(input.sizes(), osize_opt, scale_factors);
  const at::IntArrayRef output_size(osize_v);
  common_check(input.sizes(), output_size);
  // Allow for empty batch size but not other dimensions (aa_interpolation_impl.h:747-750)
  TORCH_CHECK(input.numel() != 0 || c10::multiply_integers(input.sizes().begin() + 1, input.sizes().end()),
              "Non-empty 4D data tensor expected but got a tensor with sizes ", input.sizes());
  const auto fmt = input.suggest_memory_format();
  const at::Tensor x = input.contiguous(fmt);
  TORCH_CHECK(!out_u8 || x.scalar_type() != at::kDouble, "uint8 output needs a uint8 or float32 input");
  const auto out_dtype = out_u8 ? at::kByte : (x.scalar_type() == at::kDouble ? at::kDouble : at::kFloat);
  to_aa_dtype(x.scalar_type());
  at::Tensor out = at::empty({x.size(0), x.size(1), output_size[0], output_size[1]},
                             x.options().dtype(out_dtype).memory_format(fmt));
  c10::cuda::CUDAGuard guard(x.device());
  auto stream = c10::cuda::getCurrentCUDAStream();
  aa_tensor_desc di = make_desc(x), dd = make_desc(out);
  const aa_scales sc = to_scales(scale_factors);
  const int rc = aa_resize_forward_sf(&di, &dd, filter, align_corners ? 1 : 0, &sc, (uint32_t)flags, stream.stream());
  TORCH_CHECK(rc == 0, "aa_resize_forward failed (", rc, "): ", aa_last_error());
  return out;
}

at::Tensor backward_common(const at::Tensor& grad_output, const OptSize& osize_opt, at::IntArrayRef input_size,
                           bool align_corners, int filter, bool nonaa, const OptScales& scale_factors = c10::nullopt) {
  TORCH_CHECK(grad_output.is_cuda(), "aa_interp_b200: grad_output must be a CUDA tensor (no CPU fallback)");
  TORCH_CHECK(input_size.size() == 4, "It is expected input_size equals to 4, but got size ", input_size.size());
  const std::vector<int64_t> osize_v = resolve_output_size(input_size, osize_opt, scale_factors);
  const at::IntArrayRef output_size(osize_v);
  common_check(input_size, output_size);
  // same checks as ti_upsample_bilinear2d_backward_cpu, aa_interpolation_backward_impl.h:201-212
  TORCH_CHECK(grad_output.dim() == 4, "Expected grad_output to be a tensor of dimension 4 but got: dimension ", grad_output.dim());
  const int64_t full[4] = {input_size[0], input_size[1], output_size[0], output_size[1]};
  for (int i = 0; i < 4; ++i)
    TORCH_CHECK(grad_output.size(i) == full[i], "Expected grad_output to have the same shape as output;", " output.size(", i,
                ") = ", full[i], " but got grad_output.size(", i, ") = ", grad_output.size(i));
  TORCH_CHECK(grad_output.scalar_type() == at::kFloat || grad_output.scalar_type() == at::kDouble,
              "\"ti_upsample_bilinear2d_backward\" not implemented for '", c10::toString(grad_output.scalar_type()), "'");
  const auto fmt = grad_output.suggest_memory_format();
  const at::Tensor g = grad_output.contiguous(fmt);
  at::Tensor gin = at::empty(input_size, g.options().memory_format(fmt));  // no zero_(): every element is written
  c10::cuda::CUDAGuard guard(g.device());
  auto stream = c10::cuda::getCurrentCUDAStream();
  aa_tensor_desc dg = make_desc(g), di = make_desc(gin);
  const aa_scales sc = to_scales(scale_factors);
  const int rc = nonaa ? aa_resize_backward_nonaa_bilinear(&dg, &di, align_corners ? 1 : 0, stream.stream())
                       : aa_resize_backward_sf(&dg, &di, filter, align_corners ? 1 : 0, &sc, 0u, stream.stream());
  TORCH_CHECK(rc == 0, "aa_resize_backward failed (", rc, "): ", aa_last_error());
  return gin;
}

at::Tensor linear_forward(const at::Tensor& i, const OptSize& o, bool a, const OptScales& sf) {
  return forward_common(i, o, a, AA_FILTER_TRIANGLE, 0, false, sf);
}
at::Tensor cubic_forward(const at::Tensor& i, const OptSize& o, bool a, const OptScales& sf) {
  return forward_common(i, o, a, AA_FILTER_CUBIC, 0, false, sf);
}
// The box filter keeps the input dtype for uint8, as the reference's code intends (output = at::empty({0},
// input.options()), aa_interpolation_impl.h:764; uint8 dispatch :566-570, :615-619): the float sum is truncated to
// uint8 like the C++ store `*(scalar_t*)dst = ...`.  (As built against torch 2.11 the reference itself never gets
// there: compute_indices_weights overwrites interp_size (:210) before the `interp_size > 1` test of :608, so its uint8
// call raises "not implemented for 'Byte'" -- pinned in tests/test_oracle.py.)
at::Tensor nearest_forward(const at::Tensor& i, const OptSize& o, bool a, const OptScales& sf) {
  return forward_common(i, o, a, AA_FILTER_BOX, 0, /*out_u8=*/i.scalar_type() == at::kByte, sf);
}
at::Tensor forward_with_flags(const at::Tensor& i, at::IntArrayRef o, bool a, int64_t filter, int64_t flags) {
  return forward_common(i, o.vec(), a, (int)filter, flags);
}
// fused epilogue: clamp to [0,255] + truncate (reference caller, test.py:71-75) or round to nearest -> uint8
at::Tensor forward_u8(const at::Tensor& i, at::IntArrayRef o, bool a, int64_t filter, bool round_nearest) {
  return forward_common(i, o.vec(), a, (int)filter, round_nearest ? AA_FLAG_ROUND_NEAREST : 0, /*out_u8=*/true);
}
at::Tensor linear_backward(const at::Tensor& g, const OptSize& o, at::IntArrayRef i, bool a, const OptScales& sf) {
  return backward_common(g, o, i, a, AA_FILTER_TRIANGLE, false, sf);
}
at::Tensor cubic_backward(const at::Tensor& g, const OptSize& o, at::IntArrayRef i, bool a, const OptScales& sf) {
  return backward_common(g, o, i, a, AA_FILTER_CUBIC, false, sf);
}
at::Tensor nearest_backward(const at::Tensor& g, const OptSize& o, at::IntArrayRef i, bool a, const OptScales& sf) {
  return backward_common(g, o, i, a, AA_FILTER_BOX, false, sf);
}
at::Tensor linear_backward_nonaa(const at::Tensor& g, at::IntArrayRef o, at::IntArrayRef i, bool a) {
  return backward_common(g, o.vec(), i, a, AA_FILTER_TRIANGLE, true);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  namespace py = pybind11;
  // positional (input, output_size, align_corners) exactly as the reference; scale_factors is an optional extra
  m.def("linear_forward", &linear_forward, "Anti-Aliased Linear Interpolation forward (sm_100a)", py::arg("input"),
        py::arg("output_size"), py::arg("align_corners") = false, py::arg("scale_factors") = py::none());
  m.def("nearest_forward", &nearest_forward, "Anti-Aliased box ('nearest') Interpolation forward (sm_100a)", py::arg("input"),
        py::arg("output_size"), py::arg("align_corners") = false, py::arg("scale_factors") = py::none());
  m.def("cubic_forward", &cubic_forward, "Anti-Aliased Cubic Interpolation forward (sm_100a)", py::arg("input"),
        py::arg("output_size"), py::arg("align_corners") = false, py::arg("scale_factors") = py::none());
  m.def("linear_backward", &linear_backward, "Anti-Aliased Linear Interpolation backward: true adjoint (sm_100a)",
        py::arg("grad_output"), py::arg("output_size"), py::arg("input_size"), py::arg("align_corners") = false,
        py::arg("scale_factors") = py::none());
  m.def("cubic_backward", &cubic_backward, "Anti-Aliased Cubic Interpolation backward: true adjoint (sm_100a)",
        py::arg("grad_output"), py::arg("output_size"), py::arg("input_size"), py::arg("align_corners") = false,
        py::arg("scale_factors") = py::none());
  m.def("nearest_backward", &nearest_backward, "Anti-Aliased box Interpolation backward: true adjoint (sm_100a)",
        py::arg("grad_output"), py::arg("output_size"), py::arg("input_size"), py::arg("align_corners") = false,
        py::arg("scale_factors") = py::none());
  m.def("linear_backward_nonaa", &linear_backward_nonaa, "The reference's literal (non-antialiased) linear backward");
  m.def("forward_with_flags", &forward_with_flags, "forward(input, output_size, align_corners, filter, AA_FLAG_*)");
  m.def("forward_u8", &forward_u8, "forward(input, output_size, align_corners, filter, round_nearest) -> uint8 (fused clamp + round)");
}
