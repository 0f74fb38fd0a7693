// Region loss on the device: decode, build_targets, loss + output gradient.
//
// Replaces RegionLoss.forward / RegionLossV2.forward (region_loss.py:148-232,
// 252-366) including the pure-Python CPU `build_targets` (region_loss.py:37-132)
// and its device<->host round trip (convert2cpu + 9 x .cuda()).
//
// Exactness contract of build_targets: phase 1 evaluates
// utils.bbox_ious (utils.py:54-83) in float32 with the reference's operation
// order using round-to-nearest intrinsics (never contracted to FMA); phase 2
// evaluates utils.bbox_iou (utils.py:21-52) in float64, sequentially per target
// row so that "the later ground truth overwrites" holds.
#include "common.cuh"

namespace fsdet {

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.f / (1.f + expf(-v)); }

// ------------------------------------------------------------------- decode
// `nB` rows are launched; when `nB_dev` is given (CUDA-graph replay: the number of rows kept by neg_filter changes from
// step to step but the launch is frozen at its capacity) only slots < *nB_dev are live.
__global__ void region_decode_kernel(const float* __restrict__ out, const int32_t* __restrict__ inds, int nB,
                                     const int32_t* __restrict__ nB_dev, int A, int nC,
                                     int H, int W, const float* __restrict__ anchors, float* __restrict__ pb) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    long long n = (long long)nB * A * HW;
    if (i >= n) return;
    int cell = (int)(i % HW);
    long long t = i / HW;
    int a = (int)(t % A);
    int slot = (int)(t / A);
    if (nB_dev && slot >= *nB_dev) return;
    int r = inds ? inds[slot] : slot;
    const float* o = out + ((long long)r * A * (5 + nC) + (long long)a * (5 + nC)) * HW + cell;
    float x = sigmoidf_acc(o[0]);
    float y = sigmoidf_acc(o[HW]);
    float w = expf(o[2 * HW]);
    float h = expf(o[3 * HW]);
    float4 v;
    v.x = __fadd_rn(x, (float)(cell % W));
    v.y = __fadd_rn(y, (float)(cell / W));
    v.z = __fmul_rn(w, anchors[2 * a]);
    v.w = __fmul_rn(h, anchors[2 * a + 1]);
    reinterpret_cast<float4*>(pb)[i] = v;
}

// ------------------------------------------------------------ build_targets
// float32 IoU of (cx,cy,w,h) boxes, utils.py:65-83 operation order
__device__ __forceinline__ float iou_f32(float x1, float y1, float w1, float h1, float x2, float y2, float w2, float h2) {
    float mx = fminf(__fsub_rn(x1, __fmul_rn(w1, 0.5f)), __fsub_rn(x2, __fmul_rn(w2, 0.5f)));
    float Mx = fmaxf(__fadd_rn(x1, __fmul_rn(w1, 0.5f)), __fadd_rn(x2, __fmul_rn(w2, 0.5f)));
    float my = fminf(__fsub_rn(y1, __fmul_rn(h1, 0.5f)), __fsub_rn(y2, __fmul_rn(h2, 0.5f)));
    float My = fmaxf(__fadd_rn(y1, __fmul_rn(h1, 0.5f)), __fadd_rn(y2, __fmul_rn(h2, 0.5f)));
    float uw = __fsub_rn(Mx, mx);
    float uh = __fsub_rn(My, my);
    float cw = __fsub_rn(__fadd_rn(w1, w2), uw);
    float ch = __fsub_rn(__fadd_rn(h1, h2), uh);
    bool mask = (cw <= 0.f) || (ch <= 0.f);
    float area1 = __fmul_rn(w1, h1);
    float area2 = __fmul_rn(w2, h2);
    float carea = mask ? 0.f : __fmul_rn(cw, ch);
    float uarea = __fsub_rn(__fadd_rn(area1, area2), carea);
    return __fdiv_rn(carea, uarea);
}

// float64 scalar IoU, utils.py:31-52
__device__ __forceinline__ double iou_f64(double x1, double y1, double w1, double h1, double x2, double y2, double w2,
                                          double h2) {
    double mx = fmin(__dsub_rn(x1, __dmul_rn(w1, 0.5)), __dsub_rn(x2, __dmul_rn(w2, 0.5)));
    double Mx = fmax(__dadd_rn(x1, __dmul_rn(w1, 0.5)), __dadd_rn(x2, __dmul_rn(w2, 0.5)));
    double my = fmin(__dsub_rn(y1, __dmul_rn(h1, 0.5)), __dsub_rn(y2, __dmul_rn(h2, 0.5)));
    double My = fmax(__dadd_rn(y1, __dmul_rn(h1, 0.5)), __dadd_rn(y2, __dmul_rn(h2, 0.5)));
    double uw = __dsub_rn(Mx, mx);
    double uh = __dsub_rn(My, my);
    double cw = __dsub_rn(__dadd_rn(w1, w2), uw);
    double ch = __dsub_rn(__dadd_rn(h1, h2), uh);
    if (cw <= 0.0 || ch <= 0.0) return 0.0;
    double area1 = __dmul_rn(w1, h1);
    double area2 = __dmul_rn(w2, h2);
    double carea = __dmul_rn(cw, ch);
    double uarea = __dsub_rn(__dadd_rn(area1, area2), carea);
    return __ddiv_rn(carea, uarea);
}

constexpr int kMaxGT = 64;

struct BTArgs {
    const float* pb;
    const double* target;
    const double* anchors;
    const int32_t* inds;     // optional: `target` is the FULL label matrix and slot b reads row inds[b]
    const int32_t* nB_dev;   // optional: live slots (CUDA-graph replay at fixed capacity)
    int nB, A, H, W, max_boxes;
    float noobj, obj, thresh;
    long long seen;
    float *coord_mask, *conf_mask, *cls_mask, *tx, *ty, *tw, *th, *tconf, *tcls;
    int32_t* counters;
};

// one CTA per (kept) target row
__global__ void __launch_bounds__(256) build_targets_kernel(const BTArgs a) {
    __shared__ float gt[kMaxGT][4];
    __shared__ int s_nt;
    const int b = blockIdx.x;
    if (a.nB_dev && b >= *a.nB_dev) return;          // block-uniform
    const int HW = a.H * a.W;
    const int nAnch = a.A * HW;
    const double* trow = a.target + (long long)(a.inds ? a.inds[b] : b) * 250;

    if (threadIdx.x == 0) {
        int nt = 0;
        int lim = a.max_boxes < kMaxGT ? a.max_boxes : kMaxGT;
        for (int t = 0; t < lim; ++t) {
            if (trow[t * 5 + 1] == 0.0) break;
            ++nt;
        }
        s_nt = nt;
    }
    __syncthreads();
    const int nt = s_nt;
    for (int t = threadIdx.x; t < nt; t += blockDim.x) {
        // f64 multiply, THEN the float32 cast of torch.FloatTensor([gx,gy,gw,gh]) (region_loss.py:61-65)
        gt[t][0] = __double2float_rn(__dmul_rn(trow[t * 5 + 1], (double)a.W));
        gt[t][1] = __double2float_rn(__dmul_rn(trow[t * 5 + 2], (double)a.H));
        gt[t][2] = __double2float_rn(__dmul_rn(trow[t * 5 + 3], (double)a.W));
        gt[t][3] = __double2float_rn(__dmul_rn(trow[t * 5 + 4], (double)a.H));
    }
    __syncthreads();

    const bool warm = a.seen < 12800;
    const long long base = (long long)b * nAnch;
    for (int c = threadIdx.x; c < nAnch; c += blockDim.x) {
        float4 p = __ldg(reinterpret_cast<const float4*>(a.pb) + base + c);
        float cur = 0.f;
        for (int t = 0; t < nt; ++t) {
            float v = iou_f32(p.x, p.y, p.z, p.w, gt[t][0], gt[t][1], gt[t][2], gt[t][3]);
            // torch.max(a, b) propagates NaN
            cur = (isnan(v) || isnan(cur)) ? nanf("") : fmaxf(cur, v);
        }
        a.conf_mask[base + c] = (cur > a.thresh) ? 0.f : a.noobj;
        a.coord_mask[base + c] = warm ? 1.f : 0.f;
        a.cls_mask[base + c] = 0.f;
        a.tx[base + c] = warm ? 0.5f : 0.f;
        a.ty[base + c] = warm ? 0.5f : 0.f;
        a.tw[base + c] = 0.f;
        a.th[base + c] = 0.f;
        a.tconf[base + c] = 0.f;
        a.tcls[base + c] = 0.f;
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        int nGT = 0, nCorrect = 0, nBad = 0;
        for (int t = 0; t < 50; ++t) {
            double x = trow[t * 5 + 1];
            if (x == 0.0) break;
            ++nGT;
            double gx = __dmul_rn(x, (double)a.W);
            double gy = __dmul_rn(trow[t * 5 + 2], (double)a.H);
            int gi = (int)gx;
            int gj = (int)gy;
            double gw = __dmul_rn(trow[t * 5 + 3], (double)a.W);
            double gh = __dmul_rn(trow[t * 5 + 4], (double)a.H);
            double best_iou = 0.0;
            int best_n = -1;
            for (int n = 0; n < a.A; ++n) {
                double v = iou_f64(0.0, 0.0, a.anchors[2 * n], a.anchors[2 * n + 1], 0.0, 0.0, gw, gh);
                if (v > best_iou) { best_iou = v; best_n = n; }
            }
            if (best_n < 0 || gi < 0 || gi >= a.W || gj < 0 || gj >= a.H) {
                // the reference raises here (math.log(0) / index out of range)
                ++nBad;
                continue;
            }
            long long idx = base + (long long)best_n * HW + gj * a.W + gi;
            float4 p = __ldg(reinterpret_cast<const float4*>(a.pb) + idx);
            a.coord_mask[idx] = 1.f;
            a.cls_mask[idx] = 1.f;
            a.conf_mask[idx] = a.obj;
            a.tx[idx] = (float)__dsub_rn(gx, (double)gi);
            a.ty[idx] = (float)__dsub_rn(gy, (double)gj);
            a.tw[idx] = (float)log(__ddiv_rn(gw, a.anchors[2 * best_n]));
            a.th[idx] = (float)log(__ddiv_rn(gh, a.anchors[2 * best_n + 1]));
            double v = iou_f64(gx, gy, gw, gh, (double)p.x, (double)p.y, (double)p.z, (double)p.w);
            a.tconf[idx] = (float)v;
            a.tcls[idx] = (float)trow[t * 5];
            if (v > 0.5) ++nCorrect;
        }
        if (nGT) atomicAdd(a.counters + 0, nGT);
        if (nCorrect) atomicAdd(a.counters + 1, nCorrect);
        if (nBad) atomicAdd(a.counters + 2, nBad);
    }
}

// ------------------------------------------------------------ loss + gradient
struct LossArgs {
    const float* out;
    float* grad;
    const int32_t* inds;
    const int32_t* nB_dev;   // optional: live slots (CUDA-graph replay at fixed capacity)
    const int32_t* img_start;
    int rows_total, nB, bs, cs, A, nC, H, W;
    const float *coord_mask, *conf_mask, *cls_mask, *tx, *ty, *tw, *th, *tconf, *tcls;
    float coord_scale, class_scale;
    int mode, metayolo;
    double* losses;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-reduce `n` doubles per thread and add them to dst[0..n)
template <int N>
__device__ __forceinline__ void block_accumulate(double (&v)[N], double* dst) {
    __shared__ double red[N][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        if (s != 0.0) atomicAdd(dst + threadIdx.x, s);
    }
}

// box / objectness terms (+ per-row class term for mode 1); thread per (slot, a, cell)
__global__ void __launch_bounds__(256) region_box_loss_kernel(const LossArgs a) {
    const int HW = a.H * a.W;
    long long n = (long long)a.nB * a.A * HW;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};  // x y w h conf cls proposals
    if (i < n && (!a.nB_dev || (int)(i / ((long long)a.A * HW)) < *a.nB_dev)) {
        int cell = (int)(i % HW);
        long long t = i / HW;
        int an = (int)(t % a.A);
        int slot = (int)(t / a.A);
        int r = a.inds ? a.inds[slot] : slot;
        long long off = ((long long)r * a.A * (5 + a.nC) + (long long)an * (5 + a.nC)) * HW + cell;
        const float* o = a.out + off;
        float* g = a.grad + off;
        float x = sigmoidf_acc(o[0]), y = sigmoidf_acc(o[HW]), w = o[2 * HW], h = o[3 * HW];
        float conf = sigmoidf_acc(o[4 * HW]);
        float m = a.coord_mask[i];
        float M = a.conf_mask[i];
        float sM = sqrtf(M);
        float dx = x * m - a.tx[i] * m;
        float dy = y * m - a.ty[i] * m;
        float dw = w * m - a.tw[i] * m;
        float dh = h * m - a.th[i] * m;
        float dc = conf * sM - a.tconf[i] * sM;
        acc[0] = 0.5 * a.coord_scale * (double)dx * dx;
        acc[1] = 0.5 * a.coord_scale * (double)dy * dy;
        acc[2] = 0.5 * a.coord_scale * (double)dw * dw;
        acc[3] = 0.5 * a.coord_scale * (double)dh * dh;
        acc[4] = 0.5 * (double)dc * dc;
        acc[6] = conf > 0.25f ? 1.0 : 0.0;
        g[0] = a.coord_scale * dx * m * x * (1.f - x);
        g[HW] = a.coord_scale * dy * m * y * (1.f - y);
        g[2 * HW] = a.coord_scale * dw * m;
        g[3 * HW] = a.coord_scale * dh * m;
        g[4 * HW] = dc * sM * conf * (1.f - conf);
        if (a.mode == 1 && a.cls_mask[i] == 1.f) {
            int tc = a.metayolo ? 0 : (int)a.tcls[i];
            float mxl = -INFINITY;
            for (int k = 0; k < a.nC; ++k) mxl = fmaxf(mxl, o[(5 + k) * HW]);
            float se = 0.f;
            for (int k = 0; k < a.nC; ++k) se += expf(o[(5 + k) * HW] - mxl);
            float lse = mxl + logf(se);
            acc[5] = (double)a.class_scale * (double)(lse - o[(5 + tc) * HW]);
            for (int k = 0; k < a.nC; ++k) {
                float pk = expf(o[(5 + k) * HW] - lse);
                g[(5 + k) * HW] = a.class_scale * (pk - (k == tc ? 1.f : 0.f));
            }
        }
    }
    block_accumulate<7>(acc, a.losses);
}

// RegionLossV2 class term: softmax across the cs class rows of one image.
// thread per (image, anchor, cell)
__global__ void __launch_bounds__(128) region_cls_v2_kernel(const LossArgs a) {
    const int HW = a.H * a.W;
    long long n = (long long)a.bs * a.A * HW;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc[1] = {0};
    if (i < n) {
        int cell = (int)(i % HW);
        long long t = i / HW;
        int an = (int)(t % a.A);
        int b = (int)(t / a.A);
        float msum = 0.f, tsum = 0.f;
        for (int s = a.img_start[b]; s < a.img_start[b + 1]; ++s) {
            long long k = ((long long)s * a.A + an) * HW + cell;
            msum += a.cls_mask[k];
            tsum += a.tcls[k];
        }
        if (msum == 1.f) {
            int tc = (int)tsum;
            const long long rstride = (long long)a.A * (5 + a.nC) * HW;
            const float* o = a.out + ((long long)b * a.cs) * rstride + ((long long)an * (5 + a.nC) + 5) * HW + cell;
            float* g = a.grad + ((long long)b * a.cs) * rstride + ((long long)an * (5 + a.nC) + 5) * HW + cell;
            float mxl = -INFINITY;
            for (int c = 0; c < a.cs; ++c) mxl = fmaxf(mxl, o[c * rstride]);
            float se = 0.f;
            for (int c = 0; c < a.cs; ++c) se += expf(o[c * rstride] - mxl);
            float lse = mxl + logf(se);
            if (tc >= 0 && tc < a.cs) acc[0] = (double)a.class_scale * (double)(lse - o[tc * rstride]);
            for (int c = 0; c < a.cs; ++c) {
                float pc = expf(o[c * rstride] - lse);
                g[c * rstride] = a.class_scale * (pc - (c == tc ? 1.f : 0.f));
            }
        }
    }
    block_accumulate<1>(acc, a.losses + 5);
}

__global__ void loss_total_kernel(double* losses) {
    // losses: x y w h conf cls [6]=proposals (from the box kernel) -> [6]=total, [7]=proposals
    double prop = losses[6];
    losses[7] = prop;
    losses[6] = losses[0] + losses[1] + losses[2] + losses[3] + losses[4] + losses[5];
}

}  // namespace fsdet

#ifndef FSDET_HOST_EMULATION  // tools/host_emul compiles the kernels above with g++ for CPU logic tests
using namespace fsdet;

extern "C" int fsdet_region_decode(const float* output, const int32_t* inds, int nB, const int32_t* nB_dev, int A, int nC,
                                   int H, int W, const float* anchors_f32, float* pred_boxes, void* stream) {
    FSDET_CHECK_ARG(output && anchors_f32 && pred_boxes && aligned16(pred_boxes), "region_decode: bad args");
    long long n = (long long)nB * A * H * W;
    if (n == 0) return 0;
    region_decode_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(output, inds, nB, nB_dev, A, nC, H, W, anchors_f32,
                                                                             pred_boxes);
    return launch_status("region_decode");
}

extern "C" int fsdet_build_targets(const float* pred_boxes, const double* target, const double* anchors_f64, int nB, int A,
                                   int H, int W, int max_boxes, float noobject_scale, float object_scale, float sil_thresh,
                                   long long seen, float* coord_mask, float* conf_mask, float* cls_mask, float* tx,
                                   float* ty, float* tw, float* th, float* tconf, float* tcls, int32_t* counters,
                                   const int32_t* inds, const int32_t* nB_dev, void* stream) {
    FSDET_CHECK_ARG(pred_boxes && target && anchors_f64 && counters && coord_mask && conf_mask && cls_mask && tx && ty &&
                        tw && th && tconf && tcls,
                    "build_targets: null pointer");
    FSDET_CHECK_ARG(aligned16(pred_boxes), "build_targets: pred_boxes must be 16-byte aligned");
    FSDET_CHECK_ARG(max_boxes > 0 && max_boxes <= 50, "build_targets: max_boxes %d (1..50)", max_boxes);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counters, 0, 4 * sizeof(int32_t), s);
    if (e != cudaSuccess) { set_error("build_targets: memset: %s", cudaGetErrorString(e)); return (int)e; }
    if (nB == 0) return 0;
    BTArgs a;
    a.pb = pred_boxes; a.target = target; a.anchors = anchors_f64; a.inds = inds; a.nB_dev = nB_dev; a.nB = nB; a.A = A;
    a.H = H; a.W = W;
    a.max_boxes = max_boxes; a.noobj = noobject_scale; a.obj = object_scale; a.thresh = sil_thresh; a.seen = seen;
    a.coord_mask = coord_mask; a.conf_mask = conf_mask; a.cls_mask = cls_mask; a.tx = tx; a.ty = ty; a.tw = tw; a.th = th;
    a.tconf = tconf; a.tcls = tcls; a.counters = counters;
    build_targets_kernel<<<nB, 256, 0, s>>>(a);
    return launch_status("build_targets");
}

extern "C" int fsdet_region_loss_grad(const float* output, float* grad_output, const int32_t* inds, const int32_t* nB_dev,
                                      const int32_t* img_start, int rows_total, int nB, int bs, int cs, int A, int nC,
                                      int H, int W, const float* coord_mask, const float* conf_mask,
                                      const float* cls_mask, const float* tx, const float* ty, const float* tw,
                                      const float* th, const float* tconf, const float* tcls, float coord_scale,
                                      float class_scale, int mode, int metayolo, double* losses, void* stream) {
    FSDET_CHECK_ARG(output && grad_output && losses, "region_loss_grad: null pointer");
    FSDET_CHECK_ARG(mode == 1 || (nC == 1 && img_start && bs * cs == rows_total),
                    "region_loss_grad: RegionLossV2 needs classes=1 and rows = bs*cs");
    cudaStream_t s = (cudaStream_t)stream;
    size_t total = (size_t)rows_total * A * (5 + nC) * H * W;
    cudaError_t e = cudaMemsetAsync(grad_output, 0, total * sizeof(float), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(losses, 0, 8 * sizeof(double), s);
    if (e != cudaSuccess) { set_error("region_loss_grad: memset: %s", cudaGetErrorString(e)); return (int)e; }
    LossArgs a;
    a.out = output; a.grad = grad_output; a.inds = inds; a.nB_dev = nB_dev; a.img_start = img_start; a.rows_total = rows_total; a.nB = nB;
    a.bs = bs; a.cs = cs; a.A = A; a.nC = nC; a.H = H; a.W = W; a.coord_mask = coord_mask; a.conf_mask = conf_mask;
    a.cls_mask = cls_mask; a.tx = tx; a.ty = ty; a.tw = tw; a.th = th; a.tconf = tconf; a.tcls = tcls;
    a.coord_scale = coord_scale; a.class_scale = class_scale; a.mode = mode; a.metayolo = metayolo; a.losses = losses;
    long long n = (long long)nB * A * H * W;
    if (n > 0) {
        region_box_loss_kernel<<<ceil_div(n, 256), 256, 0, s>>>(a);
        int st = launch_status("region_box_loss");
        if (st) return st;
    }
    if (mode == 0) {
        long long n2 = (long long)bs * A * H * W;
        if (n2 > 0) {
            region_cls_v2_kernel<<<ceil_div(n2, 128), 128, 0, s>>>(a);
            int st = launch_status("region_cls_v2");
            if (st) return st;
        }
    }
    loss_total_kernel<<<1, 1, 0, s>>>(losses);
    return launch_status("region_loss_total");
}
#endif  // FSDET_HOST_EMULATION
