// Memory-bound layer kernels over NHWC fp32 activations: boundary layout
// conversion, stand-alone max-pool, reorg, global max-pool, channel-slice copy,
// and the small kernels of the fused reweighting head.
//
// Reference ops replaced: F.max_pool2d / MaxPoolStride1 (darknet_meta.py:47-53,
// 260-268), Reorg (darknet_meta.py:55-74), torch.cat route (:157-171),
// GlobalMaxPool2d (pooling.py:8-27), DynamicConv2d (dynamic_conv.py:125-164).
#include "common.cuh"

namespace fsdet {

// ---------------------------------------------------------------- NCHW <-> NHWC
// tile transpose over (channel, pixel): in[b][c][p] -> out[b*HW+p][c]
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in0, int C0, const float* __restrict__ in1, int C1,
                                    float* __restrict__ out, int ld, int Cpad, int HW) {
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int C = C0 + C1;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, p = p0 + threadIdx.x;
        float v = 0.f;
        if (p < HW) {
            if (c < C0) v = in0[((long long)b * C0 + c) * HW + p];
            else if (c < C) v = in1[((long long)b * C1 + (c - C0)) * HW + p];
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int p = p0 + r, c = c0 + threadIdx.x;
        if (p < HW && c < Cpad) out[((long long)b * HW + p) * ld + c] = tile[threadIdx.x][r];
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int ld, const float* __restrict__ bias,
                                    float* __restrict__ out, int C, int HW) {
    __shared__ float tile[32][33];
    int b = blockIdx.z;
    int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int p = p0 + r, c = c0 + threadIdx.x;
        tile[r][threadIdx.x] = (p < HW && c < C) ? in[((long long)b * HW + p) * ld + c] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, p = p0 + threadIdx.x;
        if (c < C && p < HW) out[((long long)b * C + c) * HW + p] = tile[threadIdx.x][r] + (bias ? bias[c] : 0.f);
    }
}

// ------------------------------------------------------------------- max-pool
// stride 2: floor mode; stride 1: replicate pad right/bottom (MaxPoolStride1)
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int B, int H,
                                   int W, int C4, int stride, int Ho, int Wo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)B * Ho * Wo * C4;
    if (i >= n) return;
    int c = (int)(i % C4) * 4;
    long long t = i / C4;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int b = (int)(t / Ho);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            int h = min(ho * stride + dy, H - 1), w = min(wo * stride + dx, W - 1);
            float4 v = ldg4(x + (((long long)b * H + h) * W + w) * ldx + c);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    *reinterpret_cast<float4*>(y + (((long long)b * Ho + ho) * Wo + wo) * ldy + c) = m;
}

// gather form (no atomics): each input pixel sums the output windows whose first
// maximum (scan order, strict >) it is.  For stride 1 the clamped (replicated)
// taps map back to the same source pixel, matching autograd through F.pad.
__global__ void maxpool_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy,
                                   float* __restrict__ dx, int lddx, int B, int H, int W, int C, int stride, int Ho,
                                   int Wo) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)B * H * W * C;
    if (i >= n) return;
    int c = (int)(i % C);
    long long t = i / C;
    int w = (int)(t % W);
    t /= W;
    int h = (int)(t % H);
    int b = (int)(t / H);
    float g = 0.f;
    int ho_lo = stride == 2 ? h / 2 : max(h - 1, 0), ho_hi = stride == 2 ? h / 2 : h;
    int wo_lo = stride == 2 ? w / 2 : max(w - 1, 0), wo_hi = stride == 2 ? w / 2 : w;
    for (int ho = ho_lo; ho <= ho_hi && ho < Ho; ++ho)
        for (int wo = wo_lo; wo <= wo_hi && wo < Wo; ++wo) {
            // first max of this window
            int bh = -1, bw = -1;
            float bv = -INFINITY;
            for (int d = 0; d < 4; ++d) {
                int hh = min(ho * stride + (d >> 1), H - 1), ww = min(wo * stride + (d & 1), W - 1);
                float v = x[(((long long)b * H + hh) * W + ww) * ldx + c];
                if (v > bv || bh < 0) { bv = v; bh = hh; bw = ww; }
            }
            if (bh == h && bw == w) g += dy[(((long long)b * Ho + ho) * Wo + wo) * lddy + c];
        }
    dx[(((long long)b * H + h) * W + w) * lddx + c] = g;
}

// ---------------------------------------------------------------------- reorg
// out[b,(i*2+j)*C+c,h,w] = x[b,c,2h+i,2w+j]  -> NHWC: y[(b,h,w)][(i*2+j)*C+c] = x[(b,2h+i,2w+j)][c]
__global__ void reorg_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int B, int H, int W,
                             int C4, bool backward) {
    // forward: x is the fine tensor [B,H,W,C], y the coarse one [B,H/2,W/2,4C]
    // backward: y is dy (coarse, read), x is dx (fine, written)
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)B * H * W * C4;
    if (i >= n) return;
    int c = (int)(i % C4) * 4;
    long long t = i / C4;
    int w = (int)(t % W);
    t /= W;
    int h = (int)(t % H);
    int b = (int)(t / H);
    int C = C4 * 4;
    long long fine = (((long long)b * H + h) * W + w) * ldx + c;
    long long coarse = (((long long)b * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1)) * ldy + ((h & 1) * 2 + (w & 1)) * C + c;
    if (!backward) *reinterpret_cast<float4*>(y + coarse) = ldg4(x + fine);
    else *reinterpret_cast<float4*>(const_cast<float*>(x) + fine) = ldg4(y + coarse);
}

// ------------------------------------------------------------- global max-pool
__global__ void globalmax_fwd_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int32_t* __restrict__ arg,
                                     int N, int HW, int C) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    int c = i % C, n = i / C;
    float bv = -INFINITY;
    int bi = 0;
    for (int p = 0; p < HW; ++p) {
        float v = x[((long long)n * HW + p) * ldx + c];
        if (v > bv || p == 0) { bv = v; bi = p; }
    }
    y[i] = bv;
    arg[i] = bi;
}

__global__ void globalmax_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ arg, float* __restrict__ dx,
                                     int lddx, int N, int HW, int C) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)N * HW * C;
    if (i >= n) return;
    int c = (int)(i % C);
    long long t = i / C;
    int p = (int)(t % HW);
    int b = (int)(t / HW);
    dx[((long long)b * HW + p) * lddx + c] = (arg[b * C + c] == p) ? dy[b * C + c] : 0.f;
}

// ------------------------------------------------------------ channel-slice copy
__global__ void copy_channels_kernel(const float* __restrict__ src, int ldsrc, float* __restrict__ dst, int lddst,
                                     long long npix, int C4, int accumulate) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix * C4) return;
    long long p = i / C4;
    int c = (int)(i - p * C4) * 4;
    float4 v = ldg4(src + p * ldsrc + c);
    float* d = dst + p * lddst + c;
    if (accumulate) {
        float4 o = *reinterpret_cast<const float4*>(d);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
    }
    *reinterpret_cast<float4*>(d) = v;
}

__global__ void fill_kernel(float* p, float v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------- head
// weff[(c*O+o)][k] = W[o][k]*rw[c][k]; rows >= n_cls*O zero; bias_eff[c*O+o] = bias[o]
__global__ void head_weff_kernel(const float* __restrict__ Wt, const float* __restrict__ bias, const float* __restrict__ rw,
                                 float* __restrict__ weff, float* __restrict__ bias_eff, int n_cls, int O, int K, int Npad) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)Npad * K;
    if (i < Npad) {
        int r = (int)i;
        bias_eff[r] = (r < n_cls * O && bias) ? bias[r % O] : 0.f;
    }
    if (i >= n) return;
    int k = (int)(i % K);
    int r = (int)(i / K);
    float v = 0.f;
    if (r < n_cls * O) {
        int c = r / O, o = r - c * O;
        v = Wt[(long long)o * K + k] * rw[(long long)c * K + k];
    }
    weff[i] = v;
}

__global__ void head_param_grads_kernel(const float* __restrict__ dweff, const float* __restrict__ Wt,
                                        const float* __restrict__ rw, float* __restrict__ dW, float* __restrict__ drw,
                                        int n_cls, int O, int K) {
    // thread per (row, k): rows 0..O-1 -> dW, rows O..O+n_cls-1 -> drw
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)(O + n_cls) * K;
    if (i >= n) return;
    int k = (int)(i % K);
    int r = (int)(i / K);
    float s = 0.f;
    if (r < O) {
        for (int c = 0; c < n_cls; ++c) s += dweff[((long long)c * O + r) * K + k] * rw[(long long)c * K + k];
        dW[(long long)r * K + k] = s;
    } else {
        int c = r - O;
        for (int o = 0; o < O; ++o) s += dweff[((long long)c * O + o) * K + k] * Wt[(long long)o * K + k];
        drw[(long long)c * K + k] = s;
    }
}

// column sums: stage 1 -> partial[row][col] over a strip of pixels; stage 2 folds rows and classes
__global__ void colsum_strip_kernel(const float* __restrict__ d, int ld, float* __restrict__ part, long long npix, int ncol,
                                    int strip) {
    int col = blockIdx.y * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    long long p0 = (long long)blockIdx.x * strip, p1 = min(p0 + strip, npix);
    float s = 0.f;
    for (long long p = p0; p < p1; ++p) s += d[p * ld + col];
    part[(long long)blockIdx.x * ncol + col] = s;
}

__global__ void head_bias_fold_kernel(const float* __restrict__ part, int nrows, int n_cls, int O, float* __restrict__ dbias) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    double s = 0.0;
    int ncol = n_cls * O;
    for (int r = 0; r < nrows; ++r)
        for (int c = 0; c < n_cls; ++c) s += (double)part[(long long)r * ncol + c * O + o];
    dbias[o] = (float)s;
}

}  // namespace fsdet

using namespace fsdet;

extern "C" int fsdet_nchw_to_nhwc(const float* in0, int C0, const float* in1, int C1, float* out, int ld, int Cpad, int B,
                                  int HW, void* stream) {
    FSDET_CHECK_ARG(in0 && out && C0 > 0 && C1 >= 0 && (C1 == 0 || in1) && Cpad >= C0 + C1 && ld >= Cpad,
                    "nchw_to_nhwc: bad args");
    if (B == 0 || HW == 0) return 0;
    dim3 grid(ceil_div(HW, 32), ceil_div(Cpad, 32), B), block(32, 8);
    nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in0, C0, in1, C1, out, ld, Cpad, HW);
    return launch_status("nchw_to_nhwc");
}

extern "C" int fsdet_nhwc_to_nchw(const float* in, int ld, const float* bias, float* out, int B, int C, int HW,
                                  void* stream) {
    FSDET_CHECK_ARG(in && out && C > 0 && ld >= C, "nhwc_to_nchw: bad args");
    if (B == 0 || HW == 0) return 0;
    dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), block(32, 8);
    nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, ld, bias, out, C, HW);
    return launch_status("nhwc_to_nchw");
}

extern "C" int fsdet_maxpool_fwd(const float* x, int ldx, float* y, int ldy, int B, int H, int W, int C, int stride,
                                 void* stream) {
    FSDET_CHECK_ARG(x && y && (stride == 1 || stride == 2), "maxpool_fwd: bad args");
    FSDET_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "maxpool_fwd: C/ld must be multiples of 4");
    int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? W / 2 : W;
    long long n = (long long)B * Ho * Wo * (C / 4);
    if (n == 0) return 0;
    maxpool_fwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, B, H, W, C / 4, stride, Ho, Wo);
    return launch_status("maxpool_fwd");
}

extern "C" int fsdet_maxpool_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx, int B, int H,
                                 int W, int C, int stride, void* stream) {
    FSDET_CHECK_ARG(x && dy && dx && (stride == 1 || stride == 2), "maxpool_bwd: bad args");
    int Ho = stride == 2 ? H / 2 : H, Wo = stride == 2 ? W / 2 : W;
    long long n = (long long)B * H * W * C;
    if (n == 0) return 0;
    maxpool_bwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, dy, lddy, dx, lddx, B, H, W, C, stride,
                                                                           Ho, Wo);
    return launch_status("maxpool_bwd");
}

extern "C" int fsdet_reorg_fwd(const float* x, int ldx, float* y, int ldy, int B, int H, int W, int C, void* stream) {
    FSDET_CHECK_ARG(x && y && H % 2 == 0 && W % 2 == 0, "reorg_fwd: H, W must be even");
    FSDET_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "reorg_fwd: C/ld must be multiples of 4");
    long long n = (long long)B * H * W * (C / 4);
    if (n == 0) return 0;
    reorg_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, B, H, W, C / 4, false);
    return launch_status("reorg_fwd");
}

extern "C" int fsdet_reorg_bwd(const float* dy, int lddy, float* dx, int lddx, int B, int H, int W, int C, void* stream) {
    FSDET_CHECK_ARG(dy && dx && H % 2 == 0 && W % 2 == 0, "reorg_bwd: H, W must be even");
    FSDET_CHECK_ARG(C % 4 == 0 && lddx % 4 == 0 && lddy % 4 == 0, "reorg_bwd: C/ld must be multiples of 4");
    long long n = (long long)B * H * W * (C / 4);
    if (n == 0) return 0;
    reorg_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dx, lddx, const_cast<float*>(dy), lddy, B, H, W, C / 4,
                                                                     true);
    return launch_status("reorg_bwd");
}

extern "C" int fsdet_globalmax_fwd(const float* x, int ldx, float* y, int32_t* argmax, int N, int HW, int C, void* stream) {
    FSDET_CHECK_ARG(x && y && argmax && HW > 0, "globalmax_fwd: bad args");
    if (N * C == 0) return 0;
    globalmax_fwd_kernel<<<ceil_div(N * C, 128), 128, 0, (cudaStream_t)stream>>>(x, ldx, y, argmax, N, HW, C);
    return launch_status("globalmax_fwd");
}

extern "C" int fsdet_globalmax_bwd(const float* dy, const int32_t* argmax, float* dx, int lddx, int N, int HW, int C,
                                   void* stream) {
    FSDET_CHECK_ARG(dy && dx && argmax, "globalmax_bwd: bad args");
    long long n = (long long)N * HW * C;
    if (n == 0) return 0;
    globalmax_bwd_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, argmax, dx, lddx, N, HW, C);
    return launch_status("globalmax_bwd");
}

extern "C" int fsdet_copy_channels(const float* src, int ldsrc, float* dst, int lddst, size_t npix, int C, int accumulate,
                                   void* stream) {
    FSDET_CHECK_ARG(src && dst && C % 4 == 0 && ldsrc % 4 == 0 && lddst % 4 == 0, "copy_channels: alignment");
    long long n = (long long)npix * (C / 4);
    if (n == 0) return 0;
    copy_channels_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, ldsrc, dst, lddst, (long long)npix, C / 4,
                                                                             accumulate);
    return launch_status("copy_channels");
}

extern "C" int fsdet_fill(float* p, float v, size_t n, void* stream) {
    if (n == 0) return 0;
    FSDET_CHECK_ARG(p, "fill: null");
    fill_kernel<<<ceil_div((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(p, v, n);
    return launch_status("fill");
}

extern "C" int fsdet_head_weff(const float* W, const float* bias, const float* rw, float* weff, float* bias_eff, int n_cls,
                               int O, int K, int Npad, void* stream) {
    FSDET_CHECK_ARG(W && rw && weff && bias_eff && Npad >= n_cls * O, "head_weff: bad args");
    long long n = (long long)Npad * K;
    head_weff_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(W, bias, rw, weff, bias_eff, n_cls, O, K, Npad);
    return launch_status("head_weff");
}

extern "C" int fsdet_head_param_grads(const float* dweff, const float* W, const float* rw, float* dW, float* drw, int n_cls,
                                      int O, int K, void* stream) {
    FSDET_CHECK_ARG(dweff && W && rw && dW && drw, "head_param_grads: bad args");
    long long n = (long long)(O + n_cls) * K;
    head_param_grads_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dweff, W, rw, dW, drw, n_cls, O, K);
    return launch_status("head_param_grads");
}

static const int kBiasStrip = 512;
extern "C" size_t fsdet_head_bias_grad_workspace_floats(size_t npix, int n_cls, int O) {
    return (size_t)ceil_div((long long)npix, kBiasStrip) * (size_t)(n_cls * O);
}

extern "C" int fsdet_head_bias_grad(const float* d, int ld, float* dbias, float* workspace, size_t npix, int n_cls, int O,
                                    void* stream) {
    FSDET_CHECK_ARG(d && dbias && workspace, "head_bias_grad: bad args");
    int ncol = n_cls * O;
    int rows = ceil_div((long long)npix, kBiasStrip);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(rows, ceil_div(ncol, 128));
    colsum_strip_kernel<<<grid, 128, 0, s>>>(d, ld, workspace, (long long)npix, ncol, kBiasStrip);
    int st = launch_status("head_bias_grad/strip");
    if (st) return st;
    head_bias_fold_kernel<<<ceil_div(O, 64), 64, 0, s>>>(workspace, rows, n_cls, O, dbias);
    return launch_status("head_bias_grad");
}
