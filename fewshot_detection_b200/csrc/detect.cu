// Detection decode + non-maximum suppression + reweighting-vector ensembling on the device.
//
// Replaces, for evaluation (valid_ensemble.py:86-178):
//   utils.get_region_boxes     utils.py:112-193   tensor prologue, 7 device->host copies, triple Python loop
//   utils.get_region_boxes_v2  utils.py:195-290   same + softmax across the n_cls class rows of each image
//   utils.nms                  utils.py:85-104    O(n^2) Python loop per (image, class) row
//   the running mean of the support net's vectors per class, valid_ensemble.py:86-100
//
// Number formats follow the reference under torch 0.3.1 (requirements.txt:3): tensor math in float32; everything
// that the reference does on elements indexed out of a tensor (Python floats) in float64: the confidence test
// `det_conf * cls_conf > conf_thresh`, the normalisation x/w, the NMS key float32(1 - det_conf) and the NMS IoUs
// (utils.bbox_iou, utils.py:21-52, operation order kept, no FMA contraction).
#include "common.cuh"

// tools/host_emul compiles this file with g++ (threads = OS threads) to test the block-level logic without a GPU
#ifdef FSDET_HOST_EMULATION
#define FSDET_DYN_SMEM(name) unsigned char* name = emul::g_dyn_smem
#else
#define FSDET_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace fsdet {

constexpr int kDetThreads = 256;
constexpr int kCandFloats = 8;  // xs, ys, ws, hs, det_conf, cls_max_conf, (int) cls_max_id, (int) a*HW + cell

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }

// Exclusive prefix (thread order) of a per-thread flag over a 256-thread block, and the block total.
__device__ __forceinline__ int block_flag_scan(bool flag, int* s_warp, int& total) {
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pre = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[w] = __popc(m);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < kDetThreads / 32; ++i) {
        const int c = s_warp[i];
        if (i < w) base += c;
        tot += c;
    }
    __syncthreads();  // s_warp is reused by the caller's next chunk
    total = tot;
    return base + pre;
}

struct DetArgs {
    const float* out;      // [N][A*(5+nC)][H][W]
    const float* anchors;  // [2A]
    float* cand;           // [N][cap][8]
    int32_t* count;        // [N]
    float* cls_dense;      // optional [N*A*HW][nC]
    int N, A, nC, H, W, cs, v2, only_obj, cap;
    double thresh;
};

// One CTA per row n (an image of the plain detector, an (image, class) pair of the meta detector).  Candidates are
// written in the reference's loop order (cy, cx, anchor) by an ordered block compaction.
__global__ void __launch_bounds__(kDetThreads) region_detect_kernel(DetArgs p) {
    __shared__ int s_warp[kDetThreads / 32];
    const int n = blockIdx.x;
    const int HW = p.H * p.W, K = p.A * HW, C5 = 5 + p.nC;
    const int b0 = (n / p.cs) * p.cs;  // first row of this row's image (cs == 1 for the plain detector)
    const size_t row_stride = (size_t)p.A * C5 * HW;
    int base = 0;
    for (int k0 = 0; k0 < K; k0 += kDetThreads) {
        const int k = k0 + threadIdx.x;
        bool pass = false;
        float xs = 0.f, ys = 0.f, ws = 0.f, hs = 0.f, det = 0.f, cmax = 0.f;
        int cid = 0, cell = 0, a = 0;
        if (k < K) {
            cell = k / p.A;
            a = k - cell * p.A;
            const float* o = p.out + (size_t)n * row_stride + (size_t)a * C5 * HW + cell;
            det = sigmoid_acc(__ldg(o + 4 * HW));
            cmax = -1.f;
            if (p.v2) {
                // utils.py:213-220: softmax over the cs class rows of the image, per (anchor, class channel, cell)
                for (int cc = 0; cc < p.nC; ++cc) {
                    const float* q = p.out + (size_t)b0 * row_stride + (size_t)a * C5 * HW + (size_t)(5 + cc) * HW + cell;
                    float mx = -INFINITY;
                    for (int m = 0; m < p.cs; ++m) mx = fmaxf(mx, __ldg(q + m * row_stride));
                    float sum = 0.f;
                    for (int m = 0; m < p.cs; ++m) sum += expf(__ldg(q + m * row_stride) - mx);
                    const float v = __fdiv_rn(expf(__ldg(q + (size_t)(n - b0) * row_stride) - mx), sum);
                    if (p.cls_dense) p.cls_dense[((size_t)n * K + (size_t)a * HW + cell) * p.nC + cc] = v;
                    if (v > cmax) { cmax = v; cid = cc; }
                }
            } else {
                // utils.py:141: softmax over the nC class logits of the anchor-cell
                float mx = -INFINITY;
                for (int cc = 0; cc < p.nC; ++cc) mx = fmaxf(mx, __ldg(o + (size_t)(5 + cc) * HW));
                float sum = 0.f;
                for (int cc = 0; cc < p.nC; ++cc) sum += expf(__ldg(o + (size_t)(5 + cc) * HW) - mx);
                for (int cc = 0; cc < p.nC; ++cc) {
                    const float v = __fdiv_rn(expf(__ldg(o + (size_t)(5 + cc) * HW) - mx), sum);
                    if (p.cls_dense) p.cls_dense[((size_t)n * K + (size_t)a * HW + cell) * p.nC + cc] = v;
                    if (v > cmax) { cmax = v; cid = cc; }
                }
            }
            const double conf = p.only_obj ? (double)det : __dmul_rn((double)det, (double)cmax);
            pass = conf > p.thresh;
            if (pass) {
                xs = __fadd_rn(sigmoid_acc(__ldg(o)), (float)(cell % p.W));
                ys = __fadd_rn(sigmoid_acc(__ldg(o + HW)), (float)(cell / p.W));
                ws = __fmul_rn(expf(__ldg(o + 2 * HW)), __ldg(p.anchors + 2 * a));
                hs = __fmul_rn(expf(__ldg(o + 3 * HW)), __ldg(p.anchors + 2 * a + 1));
            }
        }
        int total;
        const int slot = base + block_flag_scan(pass, s_warp, total);
        if (pass && slot < p.cap) {
            float4* dst = reinterpret_cast<float4*>(p.cand + ((size_t)n * p.cap + slot) * kCandFloats);
            dst[0] = make_float4(xs, ys, ws, hs);
            dst[1] = make_float4(det, cmax, __int_as_float(cid), __int_as_float(a * HW + cell));
        }
        base += total;
    }
    if (threadIdx.x == 0) p.count[n] = base < p.cap ? base : p.cap;
}

// float64 IoU of (cx, cy, w, h) boxes, utils.py:21-52 (x1y1x2y2=False)
__device__ __forceinline__ double nms_iou(const double4 p, const double4 q) {
    const double mx = fmin(__dsub_rn(p.x, __ddiv_rn(p.z, 2.0)), __dsub_rn(q.x, __ddiv_rn(q.z, 2.0)));
    const double Mx = fmax(__dadd_rn(p.x, __ddiv_rn(p.z, 2.0)), __dadd_rn(q.x, __ddiv_rn(q.z, 2.0)));
    const double my = fmin(__dsub_rn(p.y, __ddiv_rn(p.w, 2.0)), __dsub_rn(q.y, __ddiv_rn(q.w, 2.0)));
    const double My = fmax(__dadd_rn(p.y, __ddiv_rn(p.w, 2.0)), __dadd_rn(q.y, __ddiv_rn(q.w, 2.0)));
    const double uw = __dsub_rn(Mx, mx);
    const double uh = __dsub_rn(My, my);
    const double cw = __dsub_rn(__dadd_rn(p.z, q.z), uw);
    const double ch = __dsub_rn(__dadd_rn(p.w, q.w), uh);
    if (cw <= 0.0 || ch <= 0.0) return 0.0;
    const double area1 = __dmul_rn(p.z, p.w);
    const double area2 = __dmul_rn(q.z, q.w);
    const double carea = __dmul_rn(cw, ch);
    const double uarea = __dsub_rn(__dadd_rn(area1, area2), carea);
    return __ddiv_rn(carea, uarea);
}

// One CTA per row: bitonic sort of (float32(1 - det_conf), slot) ascending = torch.sort of utils.py:89-93 with list
// order on ties, then the greedy suppression loop of utils.py:95-103 with the inner loop spread over the block.
// Dynamic shared memory: P * (8 + 32 + 1) bytes, P = power of two >= cap.
// boxes64 != nullptr: rows of already-normalised float64 boxes [N][cap][5] = {x, y, w, h, det_conf} (the list-of-lists
// form utils.nms receives) instead of `cand`.
__global__ void __launch_bounds__(kDetThreads) nms_kernel(const float* __restrict__ cand, const double* __restrict__ boxes64,
                                                          const int32_t* __restrict__ count, int cap, int P, int H, int W,
                                                          double thresh, int32_t* __restrict__ keep,
                                                          int32_t* __restrict__ keep_count) {
    FSDET_DYN_SMEM(smem_raw);
    double4* box = reinterpret_cast<double4*>(smem_raw);                                   // [P]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(box + P);             // [P]
    unsigned char* alive = reinterpret_cast<unsigned char*>(keys + P);                     // [P]
    __shared__ int s_warp[kDetThreads / 32];
    const int row = blockIdx.x;
    const int n = min(count[row], cap);
    if (n <= 0) {
        if (threadIdx.x == 0) keep_count[row] = 0;
        return;
    }
    int Pn = 1;
    while (Pn < n) Pn <<= 1;
    const float* c = cand ? cand + (size_t)row * cap * kCandFloats : nullptr;
    const double* c64 = boxes64 ? boxes64 + (size_t)row * cap * 5 : nullptr;
    for (int t = threadIdx.x; t < Pn; t += kDetThreads) {
        unsigned long long key = ~0ull;
        if (t < n) {
            const double det = c64 ? c64[(size_t)t * 5 + 4] : (double)c[(size_t)t * kCandFloats + 4];
            const float kf = (float)__dsub_rn(1.0, det);   // det_confs[i] = 1 - boxes[i][4] into a FloatTensor
            key = ((unsigned long long)__float_as_uint(kf) << 32) | (unsigned)t;
        }
        keys[t] = key;
    }
    __syncthreads();
    for (int k = 2; k <= Pn; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < Pn; t += kDetThreads) {
                const int u = t ^ j;
                if (u > t) {
                    const unsigned long long x = keys[t], y = keys[u];
                    const bool up = (t & k) == 0;
                    if ((x > y) == up) { keys[t] = y; keys[u] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < n; t += kDetThreads) {
        const int slot = (int)(keys[t] & 0xffffffffu);
        if (c64) {
            const double* q = c64 + (size_t)slot * 5;
            box[t] = make_double4(q[0], q[1], q[2], q[3]);
            alive[t] = q[4] > 0.0 ? 1 : 0;
        } else {
            const float4 v = *reinterpret_cast<const float4*>(c + (size_t)slot * kCandFloats);
            const float det = c[(size_t)slot * kCandFloats + 4];
            box[t] = make_double4(__ddiv_rn((double)v.x, (double)W), __ddiv_rn((double)v.y, (double)H),
                                  __ddiv_rn((double)v.z, (double)W), __ddiv_rn((double)v.w, (double)H));
            alive[t] = det > 0.f ? 1 : 0;
        }
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) {
        if (!alive[i]) continue;  // block-uniform: alive[i] was last written before an earlier barrier
        const double4 bi = box[i];
        for (int j = i + 1 + threadIdx.x; j < n; j += kDetThreads)
            if (alive[j] && nms_iou(bi, box[j]) > thresh) alive[j] = 0;
        __syncthreads();
    }
    int base = 0;
    for (int t0 = 0; t0 < n; t0 += kDetThreads) {
        const int t = t0 + threadIdx.x;
        const bool f = t < n && alive[t];
        int total;
        const int pos = base + block_flag_scan(f, s_warp, total);
        if (f) keep[(size_t)row * cap + pos] = (int)(keys[t] & 0xffffffffu);
        base += total;
    }
    if (threadIdx.x == 0) keep_count[row] = base;
}

// valid_ensemble.py:96-98, float32: e[c] = e[c]*cnt/(cnt+1) + dw[i]/(cnt+1) for the samples i of class c, in order.
__global__ void rw_running_mean_kernel(float* __restrict__ e, const int32_t* __restrict__ cnt_in, int32_t* __restrict__ cnt_out,
                                       const float* __restrict__ dw, const int32_t* __restrict__ ids, int n, int n_cls, int C) {
    const int c = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = cnt_in[c];
    float v = (k < C) ? e[(size_t)c * C + k] : 0.f;
    for (int i = 0; i < n; ++i) {
        if (ids[i] != c) continue;
        const float f0 = (float)cnt, f1 = (float)(cnt + 1);
        if (k < C) v = __fadd_rn(__fdiv_rn(__fmul_rn(v, f0), f1), __fdiv_rn(dw[(size_t)i * C + k], f1));
        ++cnt;
    }
    if (k < C) e[(size_t)c * C + k] = v;
    if (k == 0) cnt_out[c] = cnt;
}

static inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace fsdet

#ifndef FSDET_HOST_EMULATION
using namespace fsdet;

extern "C" int fsdet_region_detect(const float* output, const float* anchors_f32, int N, int A, int nC, int H, int W,
                                   int n_models, int v2, int only_objectness, double conf_thresh, float* cand,
                                   int32_t* count, float* cls_dense, void* stream) {
    FSDET_CHECK_ARG(output && anchors_f32 && cand && count, "region_detect: null pointer");
    FSDET_CHECK_ARG(aligned16(cand), "region_detect: cand must be 16-byte aligned");
    FSDET_CHECK_ARG(N >= 0 && A > 0 && nC > 0 && H > 0 && W > 0, "region_detect: bad shape");
    FSDET_CHECK_ARG(n_models >= 1 && (v2 ? N % n_models == 0 : n_models == 1),
                    "region_detect: %d rows are not a multiple of n_models=%d", N, n_models);
    if (N == 0) return 0;
    DetArgs p;
    p.out = output; p.anchors = anchors_f32; p.cand = cand; p.count = count; p.cls_dense = cls_dense;
    p.N = N; p.A = A; p.nC = nC; p.H = H; p.W = W; p.cs = n_models; p.v2 = v2; p.only_obj = only_objectness;
    p.cap = A * H * W; p.thresh = conf_thresh;
    region_detect_kernel<<<N, kDetThreads, 0, (cudaStream_t)stream>>>(p);
    return launch_status("region_detect");
}

static int launch_nms(const float* cand, const double* boxes64, const int32_t* count, int N, int cap, int H, int W,
                      double nms_thresh, int32_t* keep, int32_t* keep_count, void* stream) {
    FSDET_CHECK_ARG((cand || boxes64) && count && keep && keep_count, "nms: null pointer");
    FSDET_CHECK_ARG(!cand || aligned16(cand), "nms: cand must be 16-byte aligned");
    FSDET_CHECK_ARG(cap > 0 && cap <= 4096, "nms: %d candidates per row (1..4096)", cap);
    FSDET_CHECK_ARG(H > 0 && W > 0 && N >= 0, "nms: bad shape");
    if (N == 0) return 0;
    const int P = next_pow2(cap);
    const size_t smem = (size_t)P * (sizeof(double4) + sizeof(unsigned long long) + 1);
    cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("nms: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
    nms_kernel<<<N, kDetThreads, smem, (cudaStream_t)stream>>>(cand, boxes64, count, cap, P, H, W, nms_thresh, keep,
                                                               keep_count);
    return launch_status("nms");
}

extern "C" int fsdet_nms(const float* cand, const int32_t* count, int N, int cap, int H, int W, double nms_thresh,
                         int32_t* keep, int32_t* keep_count, void* stream) {
    return launch_nms(cand, nullptr, count, N, cap, H, W, nms_thresh, keep, keep_count, stream);
}

extern "C" int fsdet_nms_boxes64(const double* boxes, const int32_t* count, int N, int cap, double nms_thresh,
                                 int32_t* keep, int32_t* keep_count, void* stream) {
    return launch_nms(nullptr, boxes, count, N, cap, 1, 1, nms_thresh, keep, keep_count, stream);
}

extern "C" int fsdet_rw_running_mean(float* enews, const int32_t* cnt_in, int32_t* cnt_out, const float* dw,
                                     const int32_t* ids, int n, int n_cls, int C, void* stream) {
    FSDET_CHECK_ARG(enews && cnt_in && cnt_out && cnt_in != cnt_out, "rw_running_mean: null or aliased counters");
    FSDET_CHECK_ARG(n == 0 || (dw && ids), "rw_running_mean: null pointer");
    FSDET_CHECK_ARG(n >= 0 && n_cls > 0 && C > 0, "rw_running_mean: bad shape");
    dim3 grid(ceil_div(C, 128), n_cls);
    rw_running_mean_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(enews, cnt_in, cnt_out, dw, ids, n, n_cls, C);
    return launch_status("rw_running_mean");
}
#endif  // FSDET_HOST_EMULATION
