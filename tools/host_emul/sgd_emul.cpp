// Host-emulated build of csrc/sgd.cu (see cuda_host_emul.h).  Test tooling only.
#include "../../fewshot_detection_b200/csrc/sgd.cu"

namespace emul {
Block g_block;
unsigned char* g_dyn_smem = nullptr;
}  // namespace emul
namespace fsdet {
void set_error(const char*, ...) {}
}  // namespace fsdet

using namespace fsdet;

extern "C" int emul_sgd_step(float* const* params, const float* const* grads, float* const* moms, const long long* sizes,
                             const int32_t* chunk_tensor, const long long* chunk_offset, int n_chunks, int chunk_elems, float lr,
                             float momentum, float dampening, float weight_decay, int first_step, const float* hyper) {
    emul::launch_serial(dim3(n_chunks), dim3(256), [&]() {
        sgd_multi_kernel(params, grads, moms, sizes, chunk_tensor, chunk_offset, chunk_elems, lr, momentum, dampening, weight_decay,
                         first_step, hyper);
    });
    return 0;
}
