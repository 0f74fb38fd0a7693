// Host-emulated build of csrc/region.cu (see cuda_host_emul.h): host pointers instead of device pointers.
// Test tooling only; built by tests/test_region_host_emul.py with g++.
#include "../../fewshot_detection_b200/csrc/region.cu"

namespace emul {
Block g_block;
unsigned char* g_dyn_smem = nullptr;
}  // namespace emul
namespace fsdet {
void set_error(const char*, ...) {}
}  // namespace fsdet

using namespace fsdet;

extern "C" int emul_region_decode(const float* output, const int32_t* inds, int nB, const int32_t* nB_dev, int A, int nC, int H,
                                  int W, const float* anchors_f32, float* pred_boxes) {
    long long n = (long long)nB * A * H * W;
    emul::launch_serial(dim3(ceil_div(n, 256)), dim3(256),
                        [&]() { region_decode_kernel(output, inds, nB, nB_dev, A, nC, H, W, anchors_f32, pred_boxes); });
    return 0;
}

extern "C" int emul_build_targets(const float* pred_boxes, const double* target, const double* anchors_f64, int nB, int A, int H,
                                  int W, int max_boxes, float noobject_scale, float object_scale, float sil_thresh,
                                  long long seen, float* coord_mask, float* conf_mask, float* cls_mask, float* tx, float* ty,
                                  float* tw, float* th, float* tconf, float* tcls, int32_t* counters, const int32_t* inds,
                                  const int32_t* nB_dev) {
    for (int i = 0; i < 4; ++i) counters[i] = 0;
    BTArgs a;
    a.pb = pred_boxes; a.target = target; a.anchors = anchors_f64; a.inds = inds; a.nB_dev = nB_dev; a.nB = nB; a.A = A;
    a.H = H; a.W = W;
    a.max_boxes = max_boxes; a.noobj = noobject_scale; a.obj = object_scale; a.thresh = sil_thresh; a.seen = seen;
    a.coord_mask = coord_mask; a.conf_mask = conf_mask; a.cls_mask = cls_mask; a.tx = tx; a.ty = ty; a.tw = tw; a.th = th;
    a.tconf = tconf; a.tcls = tcls; a.counters = counters;
    emul::launch(dim3(nB), dim3(256), 0, [&]() { build_targets_kernel(a); });
    return 0;
}
