// Host side / C ABI of the recomputing first-layer kernels (conv_first_tc.cuh): the first convolution block of a
// network (nn.Conv2d(3 or 4 -> <=32, 3x3) + BatchNorm2d + LeakyReLU + MaxPool2d(2,2), darknet_meta.py:219-268 with
// cfg/darknet_dynamic.cfg:27-40 / cfg/reweighting_net.cfg:7-20) in four passes that never store its pre-BN output.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace fsdet {

#include "tc_ptx.cuh"

__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const float* src, bool valid) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void tc_kahan_add(float& s, float& e, float x) {
    const float y = x - e;
    const float t = s + y;
    e = (t - s) - y;
    s = t;
}
#define FSDET_TC_DYN_SMEM(name) extern __shared__ uint8_t name[]

#include "conv_first_tc.cuh"

static int ft_grid(int tiles) {
    const int cap = 2 * kNumSMs;
    return tiles < cap ? tiles : cap;
}

template <int MODE>
static int ft_launch(FtArgs a, cudaStream_t s) {
    using Cfg = FtCfg<MODE>;
    auto kern = conv_first_tc_kernel<MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
        set_error("conv_first_tc: cudaFuncSetAttribute(%d bytes): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        return (int)e;
    }
    kern<<<ft_grid(a.tiles), FT_THREADS, Cfg::SMEM_BYTES, s>>>(a);
    return launch_status("conv_first_tc");
}

static int ft_common(FtArgs& a, const float* in0, int C0, const float* in1, int C1, const float* w_pad4, const float* amax_x, int B,
                     int H, int W, int Cout) {
    FSDET_CHECK_ARG(in0 && w_pad4 && amax_x && C0 > 0 && C1 >= 0 && (C1 == 0 || in1) && C0 + C1 <= 4, "conv_first_tc: bad inputs");
    FSDET_CHECK_ARG(fsdet_conv_first_tc_supported(H, W, Cout), "conv_first_tc: unsupported H=%d W=%d Cout=%d", H, W, Cout);
    FSDET_CHECK_ARG((long long)B * H * W < (1ll << 31), "conv_first_tc: too many pixels");
    memset(&a, 0, sizeof(a));
    a.in0 = in0; a.in1 = in1; a.w = w_pad4; a.amax_x = amax_x; a.C0 = C0; a.C1 = C1; a.B = B; a.H = H; a.W = W; a.Cout = Cout;
    a.tiles_h = H / FT_TH; a.tiles_w = W / FT_TW; a.tiles = B * a.tiles_h * a.tiles_w;
    return 0;
}

}  // namespace fsdet

using namespace fsdet;

extern "C" int fsdet_conv_first_tc_supported(int H, int W, int Cout) {
    return H > 0 && W > 0 && H % FT_TH == 0 && W % FT_TW == 0 && Cout >= 4 && Cout <= 32 && Cout % 4 == 0;
}

extern "C" int fsdet_conv_first_tc_rows(int B, int H, int W) {
    if (H % FT_TH || W % FT_TW) return 0;
    return ft_grid(B * (H / FT_TH) * (W / FT_TW));
}

extern "C" int fsdet_conv_first_tc_stats(const float* in0, int C0, const float* in1, int C1, const float* w_pad4,
                                         const float* amax_x, float* stat_partial, int B, int H, int W, int Cout, void* stream) {
    FtArgs a;
    int rc = ft_common(a, in0, C0, in1, C1, w_pad4, amax_x, B, H, W, Cout);
    if (rc) return rc;
    FSDET_CHECK_ARG(stat_partial, "conv_first_tc_stats: null output");
    if (a.tiles == 0) return 0;
    a.stats = stat_partial;
    return ft_launch<FT_STATS>(a, (cudaStream_t)stream);
}

extern "C" int fsdet_conv_first_tc_apply(const float* in0, int C0, const float* in1, int C1, const float* w_pad4,
                                         const float* amax_x, const float* scale, const float* shift, float slope, float* y_pool,
                                         int ld_pool, void* pool_hi, void* pool_lo, int cpad, const float* amax_y, int B, int H,
                                         int W, int Cout, void* stream) {
    FtArgs a;
    int rc = ft_common(a, in0, C0, in1, C1, w_pad4, amax_x, B, H, W, Cout);
    if (rc) return rc;
    FSDET_CHECK_ARG(scale && shift && (y_pool || pool_hi), "conv_first_tc_apply: null pointer");
    FSDET_CHECK_ARG(!pool_hi || (pool_lo && amax_y && cpad >= 32 && cpad % 32 == 0 && aligned16(pool_hi) && aligned16(pool_lo)),
                    "conv_first_tc_apply: planes need lo, amax and a pitch that is a multiple of 32 (got %d)", cpad);
    FSDET_CHECK_ARG(!y_pool || (ld_pool % 4 == 0 && ld_pool >= Cout && aligned16(y_pool)), "conv_first_tc_apply: fp32 output ld=%d", ld_pool);
    if (a.tiles == 0) return 0;
    a.scale = scale; a.shift = shift; a.slope = slope; a.yp = y_pool; a.ldp = ld_pool; a.ph = pool_hi; a.pl = pool_lo; a.cpad = cpad;
    a.amax_y = amax_y;
    return ft_launch<FT_APPLY>(a, (cudaStream_t)stream);
}

extern "C" int fsdet_conv_first_tc_bwd_reduce(const float* in0, int C0, const float* in1, int C1, const float* w_pad4,
                                              const float* amax_x, const float* scale, const float* shift, const float* mean,
                                              const float* invstd, float slope, const float* dy_pool, int ld_dyp, double* partial,
                                              int B, int H, int W, int Cout, void* stream) {
    FtArgs a;
    int rc = ft_common(a, in0, C0, in1, C1, w_pad4, amax_x, B, H, W, Cout);
    if (rc) return rc;
    FSDET_CHECK_ARG(scale && shift && mean && invstd && dy_pool && partial && ld_dyp % 4 == 0 && aligned16(dy_pool),
                    "conv_first_tc_bwd_reduce: bad args");
    if (a.tiles == 0) return 0;
    a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.slope = slope; a.dyp = dy_pool; a.ld_dyp = ld_dyp;
    a.partial = partial;
    return ft_launch<FT_BWD_REDUCE>(a, (cudaStream_t)stream);
}

extern "C" size_t fsdet_conv_first_tc_wgrad_workspace_floats(int B, int H, int W) {
    return (size_t)fsdet_conv_first_tc_rows(B, H, W) * 36 * 32;
}

extern "C" int fsdet_conv_first_tc_bwd_wgrad(const float* in0, int C0, const float* in1, int C1, const float* w_pad4,
                                             const float* amax_x, const float* scale, const float* shift, const float* mean,
                                             const float* invstd, const double* coef, float slope, const float* dy_pool, int ld_dyp,
                                             const float* amax_dz, float* dw, float* workspace, size_t workspace_floats, int B,
                                             int H, int W, int Cout, void* stream) {
    FtArgs a;
    int rc = ft_common(a, in0, C0, in1, C1, w_pad4, amax_x, B, H, W, Cout);
    if (rc) return rc;
    FSDET_CHECK_ARG(scale && shift && mean && invstd && coef && dy_pool && amax_dz && dw && workspace && ld_dyp % 4 == 0 &&
                        aligned16(dy_pool) && aligned16(workspace),
                    "conv_first_tc_bwd_wgrad: bad args");
    FSDET_CHECK_ARG(workspace_floats >= fsdet_conv_first_tc_wgrad_workspace_floats(B, H, W), "conv_first_tc_bwd_wgrad: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    if (a.tiles == 0) {
        cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)Cout * 36 * sizeof(float), s);
        return e == cudaSuccess ? 0 : (int)e;
    }
    a.scale = scale; a.shift = shift; a.mean = mean; a.invstd = invstd; a.coef = coef; a.slope = slope; a.dyp = dy_pool;
    a.ld_dyp = ld_dyp; a.amax_dz = amax_dz; a.dw_partial = workspace;
    rc = ft_launch<FT_BWD_WGRAD>(a, s);
    if (rc) return rc;
    conv_first_tc_wgrad_reduce_kernel<<<ceil_div(Cout * 36, 128), 128, 0, s>>>(workspace, ft_grid(a.tiles), amax_x, amax_dz, dw, Cout);
    return launch_status("conv_first_tc_wgrad_reduce");
}
