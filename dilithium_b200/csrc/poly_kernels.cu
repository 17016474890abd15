// poly_kernels.cu — coefficient-wise kernels: pointwise multiply (pointwise_barrett,
// ref_ntt.cpp:49-57), multiply-accumulate / add / sub (butterfly.v:144-164 MULT/ADD/SUB
// modes) and the NTT-domain matrix-vector product with a shared, pre-expanded A
// (MULT_MODE loop nest, combined_top.v:921-958, :1347-1386, :1875-1913).
//
// All are streaming HBM-bound kernels: 16-byte vector accesses, a warp covers 512
// contiguous bytes per access, grid sized to a multiple of the SM count.
#include <cuda_runtime.h>

#include "dil_field.cuh"
#include "kernels.h"

namespace dil {

__device__ __forceinline__ uint4 canon4(int4 v) {
    return make_uint4(canon_signed(v.x), canon_signed(v.y), canon_signed(v.z), canon_signed(v.w));
}

template <EwOp OP>
__device__ __forceinline__ uint32_t ew1(uint32_t c, uint32_t a, uint32_t b) {
    if constexpr (OP == EwOp::MUL) return mul_full(a, b);
    if constexpr (OP == EwOp::MULACC) return reduce49((uint64_t)a * b + c);
    if constexpr (OP == EwOp::ADD) return csub(a + b);
    return csub(a + Q - b);  // SUB
}

template <EwOp OP>
__global__ void __launch_bounds__(256) elementwise_kernel(int4* __restrict__ c, const int4* __restrict__ a,
                                                          const int4* __restrict__ b, size_t n_vec) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        uint4 va = canon4(a[i]), vb = canon4(b[i]);
        uint4 vc = make_uint4(0, 0, 0, 0);
        if constexpr (OP == EwOp::MULACC) vc = canon4(c[i]);
        uint4 r;
        r.x = ew1<OP>(vc.x, va.x, vb.x);
        r.y = ew1<OP>(vc.y, va.y, vb.y);
        r.z = ew1<OP>(vc.z, va.z, vb.z);
        r.w = ew1<OP>(vc.w, va.w, vb.w);
        c[i] = make_int4((int)r.x, (int)r.y, (int)r.z, (int)r.w);
    }
}

cudaError_t launch_elementwise(EwOp op, int32_t* c, const int32_t* a, const int32_t* b, size_t n_polys, int sm_count,
                               cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    size_t n_vec = n_polys * (N / 4);
    size_t want = (n_vec + 255) / 256;
    size_t cap = (size_t)sm_count * 8;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    auto* C = reinterpret_cast<int4*>(c);
    auto* A = reinterpret_cast<const int4*>(a);
    auto* B = reinterpret_cast<const int4*>(b);
    switch (op) {
        case EwOp::MUL: elementwise_kernel<EwOp::MUL><<<grid, 256, 0, st>>>(C, A, B, n_vec); break;
        case EwOp::MULACC: elementwise_kernel<EwOp::MULACC><<<grid, 256, 0, st>>>(C, A, B, n_vec); break;
        case EwOp::ADD: elementwise_kernel<EwOp::ADD><<<grid, 256, 0, st>>>(C, A, B, n_vec); break;
        case EwOp::SUB: elementwise_kernel<EwOp::SUB><<<grid, 256, 0, st>>>(C, A, B, n_vec); break;
    }
    return cudaGetLastError();
}

// ---- mat-vec with shared pre-expanded A ----
// One thread owns a 4-coefficient column slice of one batch item: it keeps the l input
// slices in registers, streams the k*l slices of A through the read-only path (A is
// 16..56 KiB and shared by the whole batch, so it lives in L1/L2), accumulates each output
// row in 64-bit and reduces once per output coefficient.  HBM traffic is exactly
// (l + k) KiB per item.
template <int K, int L, int ITEMS>
__global__ void __launch_bounds__(256) matvec_kernel(int4* __restrict__ w, const int4* __restrict__ a_hat,
                                                     const int4* __restrict__ v, size_t batch) {
    // a thread owns one 4-coefficient column slice of ITEMS consecutive batch items, so every slice of A it
    // pulls through L1 is used ITEMS times (A re-reads, not HBM, bound the level-3/5 shapes otherwise)
    const size_t n_groups = (batch + ITEMS - 1) / ITEMS;
    const size_t n_slices = n_groups * 64;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_slices; t += stride) {
        const size_t item0 = (t >> 6) * ITEMS;
        const unsigned col = (unsigned)(t & 63);
        uint4 vr[ITEMS][L];
#pragma unroll
        for (int q = 0; q < ITEMS; q++) {
            const size_t item = item0 + q < batch ? item0 + q : batch - 1;   // tail: duplicate the last item
            const int4* vi = v + item * (size_t)(L * 64) + col;
#pragma unroll
            for (int j = 0; j < L; j++) vr[q][j] = canon4(vi[j * 64]);
        }
#pragma unroll
        for (int i = 0; i < K; i++) {
            uint64_t acc[ITEMS][4];
#pragma unroll
            for (int q = 0; q < ITEMS; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0;
#pragma unroll
            for (int j = 0; j < L; j++) {
                uint4 a = canon4(__ldg(a_hat + (i * L + j) * 64 + col));
#pragma unroll
                for (int q = 0; q < ITEMS; q++) {
                    acc[q][0] += (uint64_t)a.x * vr[q][j].x;
                    acc[q][1] += (uint64_t)a.y * vr[q][j].y;
                    acc[q][2] += (uint64_t)a.z * vr[q][j].z;
                    acc[q][3] += (uint64_t)a.w * vr[q][j].w;
                }
            }
#pragma unroll
            for (int q = 0; q < ITEMS; q++) {
                if (item0 + q < batch) {
                    int4* wi = w + (item0 + q) * (size_t)(K * 64) + col;
                    wi[i * 64] = make_int4((int)reduce49(acc[q][0]), (int)reduce49(acc[q][1]), (int)reduce49(acc[q][2]),
                                           (int)reduce49(acc[q][3]));
                }
            }
        }
    }
}

// generic dims (k, l <= 8): same algorithm, runtime loops, inputs re-read through L1
__global__ void __launch_bounds__(256) matvec_generic_kernel(int4* __restrict__ w, const int4* __restrict__ a_hat,
                                                             const int4* __restrict__ v, int K, int L, size_t n_slices) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_slices; t += stride) {
        size_t item = t >> 6;
        unsigned col = (unsigned)(t & 63);
        const int4* vi = v + item * (size_t)(L * 64) + col;
        int4* wi = w + item * (size_t)(K * 64) + col;
        for (int i = 0; i < K; i++) {
            uint64_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
            for (int j = 0; j < L; j++) {
                uint4 a = canon4(__ldg(a_hat + (i * L + j) * 64 + col));
                uint4 x = canon4(vi[j * 64]);
                acc0 += (uint64_t)a.x * x.x;
                acc1 += (uint64_t)a.y * x.y;
                acc2 += (uint64_t)a.z * x.z;
                acc3 += (uint64_t)a.w * x.w;
            }
            wi[i * 64] = make_int4((int)reduce49(acc0), (int)reduce49(acc1), (int)reduce49(acc2), (int)reduce49(acc3));
        }
    }
}

cudaError_t launch_matvec(int32_t* w, const int32_t* a_hat, const int32_t* v, int k, int l, size_t batch, int sm_count,
                          cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    size_t n_slices = batch * 64;
    size_t want = (n_slices + 255) / 256;
    size_t cap = (size_t)sm_count * 8;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    auto* W = reinterpret_cast<int4*>(w);
    auto* A = reinterpret_cast<const int4*>(a_hat);
    auto* V = reinterpret_cast<const int4*>(v);
    if (k == 4 && l == 4) matvec_kernel<4, 4, 1><<<grid, 256, 0, st>>>(W, A, V, batch);
    else if (k == 6 && l == 5) matvec_kernel<6, 5, 1><<<grid, 256, 0, st>>>(W, A, V, batch);   // 2 items/thread measured slower here
    else if (k == 8 && l == 7) matvec_kernel<8, 7, 2><<<grid, 256, 0, st>>>(W, A, V, batch);
    else matvec_generic_kernel<<<grid, 256, 0, st>>>(W, A, V, k, l, n_slices);
    return cudaGetLastError();
}

}  // namespace dil
