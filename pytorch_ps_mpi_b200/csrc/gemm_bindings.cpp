// Bindings of the broadcast-gated tcgen05 GEMM (csrc/kernels/bcast_gemm.cu): builds the TMA
// tensor maps with cuTensorMapEncodeTiled (resolved through the runtime, no -lcuda) and launches.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda.h>
#include <torch/extension.h>

#include "kernels.h"

namespace py = pybind11;
int psb_bcast_gemm_smem_bytes();

namespace {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn f = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr)
      throw std::runtime_error("cuTensorMapEncodeTiled not available");
    return reinterpret_cast<EncodeFn>(p);
  }();
  return f;
}

// 2-D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols] (128 bytes), 128B swizzle
CUtensorMap make_map(uint64_t ptr, int64_t rows, int64_t cols, int64_t row_stride_elems, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, reinterpret_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return m;
}

// y[M,N] = act(x[M,K] @ w[N,K]^T + bias).  `w_ptr` may be any mapped address holding the [N,K] bf16
// weight (this rank's parameter arena, or the server's arena over NVLink).  `flag_ptr` (optional) is
// the SIG_PARAMS_READY slot the TMA producer acquires before its first weight load.
at::Tensor bcast_gemm(const at::Tensor& x, uint64_t w_ptr, int64_t N, int64_t K, c10::optional<at::Tensor> bias, bool relu,
                      uint64_t flag_ptr, uint64_t epoch, double timeout_s) {
  TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.dim() == 2 && x.is_contiguous(), "x must be [M,K] bf16 contiguous");
  TORCH_CHECK(x.size(1) == K && K % 8 == 0, "K mismatch / K must be a multiple of 8");
  TORCH_CHECK(reinterpret_cast<uintptr_t>(x.data_ptr()) % 16 == 0 && w_ptr % 16 == 0, "operands must be 16-byte aligned");
  const int64_t M = x.size(0);
  auto y = at::empty({M, N}, x.options());
  if (M == 0) return y;
  CUtensorMap ma = make_map(reinterpret_cast<uint64_t>(x.data_ptr()), M, K, K, 128);
  CUtensorMap mb = make_map(w_ptr, N, K, K, 128);
  BcastGemmArgs a{};
  a.tmap_a = &ma;
  a.tmap_b = &mb;
  a.tmap_c = y.data_ptr();
  const float* bp = nullptr;
  at::Tensor bias_f;
  if (bias.has_value() && bias->defined()) {
    bias_f = bias->to(at::kFloat).contiguous();
    bp = bias_f.data_ptr<float>();
  }
  a.bias = bp;
  a.ready_flag = reinterpret_cast<const uint64_t*>(flag_ptr);
  a.ready_epoch = epoch;
  a.M = (int)M, a.N = (int)N, a.K = (int)K;
  a.relu = relu ? 1 : 0;
  a.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  psb_launch_bcast_gemm(c10::cuda::getCurrentCUDAStream().stream(), a, sms);
  cudaError_t e = cudaGetLastError();
  TORCH_CHECK(e == cudaSuccess, "psb_bcast_gemm_kernel launch: ", cudaGetErrorString(e));
  return y;
}

}  // namespace

void bind_gemm(py::module_& m) {
  m.def("bcast_gemm", &bcast_gemm, py::arg("x"), py::arg("w_ptr"), py::arg("N"), py::arg("K"), py::arg("bias"),
        py::arg("relu"), py::arg("flag_ptr") = 0, py::arg("epoch") = 0, py::arg("timeout_s") = 30.0,
        "tcgen05/TMEM/TMA GEMM whose weight tiles are gated on the PS broadcast epoch flag");
  m.def("bcast_gemm_smem_bytes", &psb_bcast_gemm_smem_bytes);
}
