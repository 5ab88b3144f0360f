// Row softmax of the output layer as a stand-alone kernel (one CTA per row): the full-forward path of long batches and every
// lazy / masked call.  The arithmetic lives in softmax_row.cuh, shared with the fused layer kernel.
//   SoftMax::apply           src/cpp/dnn.cc:534-544
//   LazyOutputActivations    src/cpp/dnn.cc:355-392

#include <cuda_runtime.h>

#include "kernels.h"
#include "ptx.cuh"
#include "softmax_row.cuh"

namespace fdnn {

namespace {

constexpr int kThreads = kSoftmaxThreads;
constexpr int kMaxSmemFloats = 56 * 1024;  // 224 KB of exponentials; wider rows recompute instead

template <bool kCache>
__global__ void __launch_bounds__(kThreads) softmax_kernel(const SoftmaxArgs a) {
  extern __shared__ __align__(16) float s_e[];
  __shared__ float s_red[kThreads / 32];
  ptx::griddep_wait();
  ptx::griddep_launch_dependents();
  const int row = int(blockIdx.x);
  const int O = a.O;
  const float *x = a.logits + size_t(row) * size_t(a.ld);
  const int8_t *m = a.mask ? a.mask + size_t(row) * size_t(a.mask_ld) : nullptr;
  float *y = a.out + size_t(row) * size_t(a.out_ld);
  const bool vec = (O % 4 == 0) && (a.ld % 4 == 0) && (a.out_ld % 4 == 0) && (m == nullptr || a.mask_ld % 4 == 0);
  softmax_row<kCache>(x, m, y, O, vec, s_e, s_red, int(threadIdx.x), [] { __syncthreads(); });
}

}  // namespace

cudaError_t softmax_configure() {
  return cudaFuncSetAttribute(softmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemFloats * 4);
}

cudaError_t launch_softmax(const SoftmaxArgs &a, cudaStream_t stream) {
  if (a.rows <= 0) return cudaSuccess;
  // the uncached variant re-reads its input in the second pass, so it cannot run in place
  const bool cache = a.O <= kMaxSmemFloats;
  if (!cache && a.logits == a.out) return cudaErrorInvalidValue;
  if (cache) return launch_pdl(softmax_kernel<true>, dim3(a.rows), dim3(kThreads), size_t(a.O) * 4, stream, pdl_enabled(), a);
  return launch_pdl(softmax_kernel<false>, dim3(a.rows), dim3(kThreads), size_t(0), stream, pdl_enabled(), a);
}

}  // namespace fdnn
