// oracle/ref_chamfer_shim.cu -- TEST INFRASTRUCTURE, not product code.
//
// C-ABI shim around the UNMODIFIED reference Chamfer kernels (/root/reference/chamfer/chamfer3D/chamfer3D.cu:
// NmDistanceKernel :9-138 launched <<<dim3(32,16,1),512>>> by chamfer_cuda_forward :142-165, NmDistanceGradKernel
// :167-195 launched <<<dim3(1,16,1),256>>> by chamfer_cuda_backward :199-230).  The reference source is compiled where it
// lies (oracle/build_ref.sh) against oracle/ref_stub/ATen/ATen.h; this file only wraps raw device pointers into the stub
// tensors.  Like the reference, everything runs on the legacy default stream, the caller zero-fills dist / idx / grads
// (dist_chamfer_3D.py:55-63,77-81 does it with torch.zeros).
#include <ATen/ATen.h>

int chamfer_cuda_forward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor dist1, at::Tensor dist2, at::Tensor idx1,
                         at::Tensor idx2);
int chamfer_cuda_backward(at::Tensor xyz1, at::Tensor xyz2, at::Tensor gradxyz1, at::Tensor gradxyz2, at::Tensor graddist1,
                          at::Tensor graddist2, at::Tensor idx1, at::Tensor idx2);

static at::Tensor wrap(const void* p, int64_t d0, int64_t d1, int64_t d2) {
  at::Tensor t;
  t.ptr = const_cast<void*>(p);
  t.dims[0] = d0; t.dims[1] = d1; t.dims[2] = d2;
  return t;
}

extern "C" {
int chamferref_forward(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1, float* dist2, int* idx1,
                       int* idx2) {
  return chamfer_cuda_forward(wrap(xyz1, b, n, 3), wrap(xyz2, b, m, 3), wrap(dist1, b, n, 0), wrap(dist2, b, m, 0),
                              wrap(idx1, b, n, 0), wrap(idx2, b, m, 0));
}
int chamferref_backward(int b, int n, const float* xyz1, int m, const float* xyz2, float* gxyz1, float* gxyz2,
                        const float* gdist1, const float* gdist2, const int* idx1, const int* idx2) {
  return chamfer_cuda_backward(wrap(xyz1, b, n, 3), wrap(xyz2, b, m, 3), wrap(gxyz1, b, n, 3), wrap(gxyz2, b, m, 3),
                               wrap(gdist1, b, n, 0), wrap(gdist2, b, m, 0), wrap(idx1, b, n, 0), wrap(idx2, b, m, 0));
}
}
