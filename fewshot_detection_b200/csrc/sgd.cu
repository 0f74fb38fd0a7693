// Fused multi-tensor SGD(momentum, dampening, weight decay): one launch per step.
//
// Replaces torch.optim.SGD.step as configured by the reference driver
// (train_meta.py:143-147: momentum 0.9, dampening 0, weight decay applied to
// every parameter): ~5 pointwise passes x 89 tensors there, one pass here.
// HBM-bound: reads p, g, m and writes p, m once (20 B per element).
#include "common.cuh"

namespace fsdet {

__global__ void __launch_bounds__(256) sgd_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                                                        float* const* __restrict__ moms, const long long* __restrict__ sizes,
                                                        const int32_t* __restrict__ chunk_tensor,
                                                        const long long* __restrict__ chunk_offset, int chunk_elems, float lr,
                                                        float momentum, float dampening, float wd, int first,
                                                        const float* __restrict__ hyper) {
    if (hyper) {  // CUDA-graph mode: lr / momentum / dampening / weight decay live in device memory
        lr = hyper[0]; momentum = hyper[1]; dampening = hyper[2]; wd = hyper[3];
    }
    const int t = chunk_tensor[blockIdx.x];
    const long long off = chunk_offset[blockIdx.x];
    float* __restrict__ p = params[t] + off;
    const float* __restrict__ g = grads[t] + off;
    float* __restrict__ m = moms[t] + off;
    long long n = sizes[t] - off;
    if (n > chunk_elems) n = chunk_elems;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m)) & 15u) == 0;
    if (vec) {
        long long n4 = n >> 2;
        for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
            float4 pv = reinterpret_cast<float4*>(p)[i];
            float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
            float4 mv = first ? make_float4(0, 0, 0, 0) : reinterpret_cast<float4*>(m)[i];
            float d;
            d = fmaf(wd, pv.x, gv.x); mv.x = first ? d : fmaf(momentum, mv.x, (1.f - dampening) * d); pv.x = fmaf(-lr, mv.x, pv.x);
            d = fmaf(wd, pv.y, gv.y); mv.y = first ? d : fmaf(momentum, mv.y, (1.f - dampening) * d); pv.y = fmaf(-lr, mv.y, pv.y);
            d = fmaf(wd, pv.z, gv.z); mv.z = first ? d : fmaf(momentum, mv.z, (1.f - dampening) * d); pv.z = fmaf(-lr, mv.z, pv.z);
            d = fmaf(wd, pv.w, gv.w); mv.w = first ? d : fmaf(momentum, mv.w, (1.f - dampening) * d); pv.w = fmaf(-lr, mv.w, pv.w);
            reinterpret_cast<float4*>(m)[i] = mv;
            reinterpret_cast<float4*>(p)[i] = pv;
        }
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float d = fmaf(wd, p[i], g[i]);
            float mv = first ? d : fmaf(momentum, m[i], (1.f - dampening) * d);
            m[i] = mv;
            p[i] = fmaf(-lr, mv, p[i]);
        }
    } else {
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            float d = fmaf(wd, p[i], g[i]);
            float mv = first ? d : fmaf(momentum, m[i], (1.f - dampening) * d);
            m[i] = mv;
            p[i] = fmaf(-lr, mv, p[i]);
        }
    }
}

}  // namespace fsdet

#ifndef FSDET_HOST_EMULATION  // tools/host_emul compiles the kernel above with g++ for CPU logic tests
using namespace fsdet;

extern "C" int fsdet_sgd_step(float* const* params, const float* const* grads, float* const* moms, const long long* sizes,
                              const int32_t* chunk_tensor, const long long* chunk_offset, int n_chunks, int chunk_elems,
                              float lr, float momentum, float dampening, float weight_decay, int first_step,
                              const float* hyper_dev, void* stream) {
    FSDET_CHECK_ARG(params && grads && moms && sizes && chunk_tensor && chunk_offset, "sgd_step: null table");
    FSDET_CHECK_ARG(chunk_elems > 0 && chunk_elems % 4 == 0, "sgd_step: chunk_elems must be a positive multiple of 4");
    if (n_chunks == 0) return 0;
    sgd_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(params, grads, moms, sizes, chunk_tensor, chunk_offset,
                                                                 chunk_elems, lr, momentum, dampening, weight_decay, first_step, hyper_dev);
    return launch_status("sgd_step");
}
#endif  // FSDET_HOST_EMULATION
