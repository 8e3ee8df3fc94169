// oracle/ref_stub/ATen/ATen.h -- TEST INFRASTRUCTURE, not product code.
//
// The reference's chamfer/chamfer3D/chamfer3D.cu includes <ATen/ATen.h> only for the at::Tensor arguments of its two
// host launchers (chamfer_cuda_forward / chamfer_cuda_backward, chamfer3D.cu:142-165,199-230), of which it uses
// exactly .size(i) and .data<T>().  Compiling against the real torch headers costs minutes and ties the checker to
// libtorch; this stub provides those two members over raw device pointers so that the UNMODIFIED reference source
// (kernels + launch configuration) compiles with plain nvcc.  oracle/ref_chamfer_shim.cu builds the stub tensors.
#pragma once
#include <cstdint>
namespace at {
struct Tensor {
  void* ptr = nullptr;
  int64_t dims[4] = {0, 0, 0, 0};
  int64_t size(int i) const { return dims[i]; }
  template <typename T>
  T* data() const { return reinterpret_cast<T*>(ptr); }
};
}  // namespace at
