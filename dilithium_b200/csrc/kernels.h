// kernels.h — host-side launchers of the engine's CUDA kernels (internal to the library).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace dil {

enum class EwOp : int { MUL = 0, MULACC = 1, ADD = 2, SUB = 3 };

cudaError_t launch_ntt_fwd(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st);
cudaError_t launch_ntt_inv(int32_t* dst, const int32_t* src, size_t n_polys, int sm_count, cudaStream_t st);
cudaError_t launch_elementwise(EwOp op, int32_t* c, const int32_t* a, const int32_t* b, size_t n_polys, int sm_count,
                               cudaStream_t st);
cudaError_t launch_matvec(int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l, size_t batch, int sm_count,
                          cudaStream_t st);
cudaError_t launch_expand_a(int32_t* a_hat, const uint8_t* rho, size_t n_rho, int k, int l, int sm_count, cudaStream_t st);
cudaError_t launch_matvec_expand(int32_t* w, const uint8_t* rho, const int32_t* v, int k, int l, size_t batch,
                                 unsigned flags, int sm_count, cudaStream_t st);
cudaError_t launch_signcore(int32_t* w, const int32_t* a_hat, const int32_t* y, int k, int l, size_t batch, int sm_count,
                            cudaStream_t st);

}  // namespace dil
