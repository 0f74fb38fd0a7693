// fp32 SIMT implicit-GEMM convolution (stride 1, "same" padding, k in {1,3}).
//
// Replaces the cuDNN calls behind nn.Conv2d in the reference
// (darknet_meta.py:236-252) for every layer / precision where the tensor-core
// path (conv_tc.cu) is not used, and is the exact-fp32 parity baseline.
//
//   forward / dgrad : z[p][n]  = sum_k A[p][k] * w[n][k],   A = im2col(x) gathered on the fly
//   wgrad           : dw[n][k] = sum_p dz[p][n] * A[p][k]
//
// Tiles: 128 pixels x BN channels x 16 k, 256 threads, 8 x TN register tile,
// double-buffered shared memory with register prefetch.
#include "common.cuh"

namespace fsdet {

struct ConvArgs {
    const float* x;
    const float* w;
    const float* bias;
    float* z;
    float* stat;
    int ldx, ldz;
    int B, H, W, Cin, Cout, ks, pad;
    int K;        // ks*ks*Cin
    long long M;  // B*H*W
    int accumulate;
};

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int APAD = 4;

template <int TN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvArgs p) {
    constexpr int BN = 16 * TN;
    __shared__ __align__(16) float As[2][BK][BM + APAD];
    __shared__ __align__(16) float Bs[2][BK][BN + APAD];

    const int tid = threadIdx.x;
    const int tx = tid & 15;   // n direction
    const int ty = tid >> 4;   // m direction
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int HW = p.H * p.W;

    // ---- A (im2col) load bookkeeping: 2 rows per thread, one float4 of k each
    const int kv = tid & 3;
    const int rowA = tid >> 2;  // 0..63, second row = +64
    int a_h[2], a_w[2];
    bool a_ok[2];
    const float* a_ptr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        long long m = m0 + rowA + r * 64;
        a_ok[r] = m < p.M;
        long long mm = a_ok[r] ? m : 0;
        int rem = (int)(mm % HW);
        a_h[r] = rem / p.W;
        a_w[r] = rem - a_h[r] * p.W;
        a_ptr[r] = p.x + mm * p.ldx;
    }
    // running decomposition of this thread's k index into (tap, ci)
    int a_ci = kv * 4, a_tap = 0;
    while (a_ci >= p.Cin) { a_ci -= p.Cin; ++a_tap; }

    // ---- B (weights) load bookkeeping
    constexpr int B_ROWS_PER_PASS = 64;
    constexpr int B_PASSES = (BN + B_ROWS_PER_PASS - 1) / B_ROWS_PER_PASS;
    const int rowB = tid >> 2;

    float4 ra[2], rb[B_PASSES];
    const int nk = (p.K + BK - 1) / BK;

    auto load_global = [&](int kc) {
        const int k = kc * BK + kv * 4;
        const bool kok = k < p.K;
        int dy = 0, dx = 0;
        if (p.ks == 3) { dy = a_tap / 3 - 1; dx = a_tap - (a_tap / 3) * 3 - 1; }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int hh = a_h[r] + dy, ww = a_w[r] + dx;
            if (kok && a_ok[r] && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
                v = ldg4(a_ptr[r] + (long long)(dy * p.W + dx) * p.ldx + a_ci);
            ra[r] = v;
        }
#pragma unroll
        for (int q = 0; q < B_PASSES; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            int rr = rowB + q * B_ROWS_PER_PASS;
            int n = n0 + rr;
            if (rr < BN && kok && n < p.Cout) v = ldg4(p.w + (long long)n * p.K + k);
            rb[q] = v;
        }
        // advance (tap, ci) by BK for the next chunk
        a_ci += BK;
        while (a_ci >= p.Cin) { a_ci -= p.Cin; ++a_tap; }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            int row = rowA + r * 64;
            As[buf][kv * 4 + 0][row] = ra[r].x;
            As[buf][kv * 4 + 1][row] = ra[r].y;
            As[buf][kv * 4 + 2][row] = ra[r].z;
            As[buf][kv * 4 + 3][row] = ra[r].w;
        }
#pragma unroll
        for (int q = 0; q < B_PASSES; ++q) {
            int rr = rowB + q * B_ROWS_PER_PASS;
            if (rr < BN) {
                Bs[buf][kv * 4 + 0][rr] = rb[q].x;
                Bs[buf][kv * 4 + 1][rr] = rb[q].y;
                Bs[buf][kv * 4 + 2][rr] = rb[q].z;
                Bs[buf][kv * 4 + 3][rr] = rb[q].w;
            }
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_global(0);
    store_smem(0);
    __syncthreads();
    int buf = 0;
    for (int kc = 0; kc < nk; ++kc) {
        if (kc + 1 < nk) load_global(kc + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else {
                float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
                b[0] = b0.x; b[1] = b0.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nk) store_smem(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue ------------------------------------------------------
    // column index of register column j
    auto col_of = [&](int j) -> int {
        if constexpr (TN == 8) return (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4));
        else if constexpr (TN == 4) return tx * 4 + j;
        else return tx * 2 + j;
    };
    const bool vec_ok = ((p.ldz & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.z) & 15u) == 0) && (TN >= 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
        if (m >= p.M) continue;
        float* zr = p.z + m * p.ldz;
#pragma unroll
        for (int j0 = 0; j0 < TN; j0 += (TN >= 4 ? 4 : 2)) {
            int n = n0 + col_of(j0);
            if constexpr (TN >= 4) {
                if (vec_ok && n + 3 < p.Cout) {
                    float4 v = make_float4(acc[i][j0], acc[i][j0 + 1], acc[i][j0 + 2], acc[i][j0 + 3]);
                    if (p.bias) { v.x += p.bias[n]; v.y += p.bias[n + 1]; v.z += p.bias[n + 2]; v.w += p.bias[n + 3]; }
                    if (p.accumulate) {
                        float4 o = *reinterpret_cast<const float4*>(zr + n);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    *reinterpret_cast<float4*>(zr + n) = v;
                    continue;
                }
            }
#pragma unroll
            for (int j = j0; j < j0 + (TN >= 4 ? 4 : 2); ++j) {
                int nn = n0 + col_of(j);
                if (nn < p.Cout) {
                    float v = acc[i][j];
                    if (p.bias) v += p.bias[nn];
                    if (p.accumulate) v += zr[nn];
                    zr[nn] = v;
                }
            }
        }
    }

    if (p.stat) {
        // per-CTA column statistics over the valid rows: sum, sum of squares, min, max (BatchNorm partials +
        // the activation range).  Reduce the 16 ty-threads through shared memory (re-using the operand buffers).
        float* red_s = &As[0][0][0];       // [16][BN]
        float* red_mn = red_s + 16 * BN;   // [16][BN]   (As holds 2*16*(128+4) floats)
        float* red_q = &Bs[0][0][0];       // [16][BN]
        float* red_mx = red_q + 16 * BN;   // [16][BN]   (Bs holds 2*16*(BN+4) floats)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = 0.f, q = 0.f, mn = INFINITY, mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
                if (m < p.M) {
                    const float v = acc[i][j];
                    s += v; q += v * v; mn = fminf(mn, v); mx = fmaxf(mx, v);
                }
            }
            red_s[ty * BN + col_of(j)] = s;
            red_q[ty * BN + col_of(j)] = q;
            red_mn[ty * BN + col_of(j)] = mn;
            red_mx[ty * BN + col_of(j)] = mx;
        }
        __syncthreads();
        if (tid < BN) {
            int n = n0 + tid;
            if (n < p.Cout) {
                float s = 0.f, q = 0.f, mn = INFINITY, mx = -INFINITY;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    s += red_s[t * BN + tid]; q += red_q[t * BN + tid];
                    mn = fminf(mn, red_mn[t * BN + tid]); mx = fmaxf(mx, red_mx[t * BN + tid]);
                }
                float* dst = p.stat + (long long)blockIdx.x * 4 * p.Cout;
                dst[n] = s;
                dst[p.Cout + n] = q;
                dst[2 * p.Cout + n] = mn;
                dst[3 * p.Cout + n] = mx;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// wgrad: dw[co][kidx] = sum_p dz[p][co] * A[p][kidx]
struct WgradArgs {
    const float* x;
    const float* dz;
    float* out;  // dw (splits == 1) or workspace [splits][Cout][K]
    int ldx, lddz;
    int B, H, W, Cin, Cout, ks, pad, K;
    long long M;
    long long pix_per_split;
};

template <int BMC>  // co tile: 64 or 128
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const WgradArgs p) {
    constexpr int BNK = 128;
    constexpr int TM = BMC / 16;  // 4 or 8
    __shared__ __align__(16) float As[2][BK][BMC];
    __shared__ __align__(16) float Bs[2][BK][BNK];

    const int tid = threadIdx.x;
    const int tx = tid & 15;  // kidx direction
    const int ty = tid >> 4;  // co direction
    const int k0 = blockIdx.x * BNK;
    const int co0 = blockIdx.y * BMC;
    const long long pbeg = (long long)blockIdx.z * p.pix_per_split;
    long long pend = pbeg + p.pix_per_split;
    if (pend > p.M) pend = p.M;
    const int HW = p.H * p.W;

    // dz loads: BK x BMC floats
    constexpr int A_VECS_PER_ROW = BMC / 4;             // 16 or 32
    constexpr int A_ROWS_PER_PASS = 256 / A_VECS_PER_ROW;  // 16 or 8
    constexpr int A_PASSES = BK / A_ROWS_PER_PASS;      // 1 or 2
    const int a_vec = tid % A_VECS_PER_ROW;
    const int a_row = tid / A_VECS_PER_ROW;
    const int a_co = co0 + a_vec * 4;
    const bool a_cok = a_co < p.Cout;  // Cout % 4 == 0 is required

    // im2col loads: BK x 128 floats, 2 passes of 8 rows
    const int b_vec = tid & 31;
    const int b_row = tid >> 5;
    const int b_k = k0 + b_vec * 4;
    const bool b_kok = b_k < p.K;
    int b_tap = 0, b_ci = 0, b_dy = 0, b_dx = 0;
    if (b_kok) {
        b_tap = b_k / p.Cin;
        b_ci = b_k - b_tap * p.Cin;
        if (p.ks == 3) { b_dy = b_tap / 3 - 1; b_dx = b_tap - (b_tap / 3) * 3 - 1; }
    }

    float4 ra[A_PASSES], rb[2];
    const long long npix = pend > pbeg ? (pend - pbeg) : 0;
    const int nk = (int)((npix + BK - 1) / BK);

    auto load_global = [&](int kc) {
        const long long pb = pbeg + (long long)kc * BK;
#pragma unroll
        for (int q = 0; q < A_PASSES; ++q) {
            long long pp = pb + a_row + q * A_ROWS_PER_PASS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_cok && pp < pend) v = ldg4(p.dz + pp * p.lddz + a_co);
            ra[q] = v;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            long long pp = pb + b_row + q * 8;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b_kok && pp < pend) {
                int rem = (int)(pp % HW);
                int h = rem / p.W;
                int w = rem - h * p.W;
                int hh = h + b_dy, ww = w + b_dx;
                if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W)
                    v = ldg4(p.x + (pp + b_dy * p.W + b_dx) * p.ldx + b_ci);
            }
            rb[q] = v;
        }
    };
    auto store_smem = [&](int buf) {
#pragma unroll
        for (int q = 0; q < A_PASSES; ++q)
            *reinterpret_cast<float4*>(&As[buf][a_row + q * A_ROWS_PER_PASS][a_vec * 4]) = ra[q];
#pragma unroll
        for (int q = 0; q < 2; ++q) *reinterpret_cast<float4*>(&Bs[buf][b_row + q * 8][b_vec * 4]) = rb[q];
    };

    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    if (nk > 0) {
        load_global(0);
        store_smem(0);
    }
    __syncthreads();
    int buf = 0;
    for (int kc = 0; kc < nk; ++kc) {
        if (kc + 1 < nk) load_global(kc + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[8];
            if constexpr (TM == 8) {
                float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            } else {
                float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
                a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
            }
            float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kc + 1 < nk) store_smem(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    float* out = p.out + (long long)blockIdx.z * p.Cout * p.K;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int co = co0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
        if (co >= p.Cout) continue;
        float* orow = out + (long long)co * p.K;
#pragma unroll
        for (int j0 = 0; j0 < 8; j0 += 4) {
            int k = k0 + ((j0 == 0) ? tx * 4 : 64 + tx * 4);
            if (k < p.K)  // K % 4 == 0
                *reinterpret_cast<float4*>(orow + k) = make_float4(acc[i][j0], acc[i][j0 + 1], acc[i][j0 + 2], acc[i][j0 + 3]);
        }
    }
}

__global__ void splitk_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ out, long long n4, int splits) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 s = ws[i];
    for (int k = 1; k < splits; ++k) {
        float4 v = ws[(long long)k * n4 + i];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    out[i] = s;
}

__global__ void weight_flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int kk, int Cin) {
    // wt[ci][kk-1-tap][co] = w[co][tap][ci]; 32x32 smem transpose per tap
    __shared__ float tile[32][33];
    int tap = blockIdx.z;
    int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int co = co0 + r, ci = ci0 + threadIdx.x;
        tile[r][threadIdx.x] = (co < Cout && ci < Cin) ? w[((long long)co * kk + tap) * Cin + ci] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int ci = ci0 + r, co = co0 + threadIdx.x;
        if (ci < Cin && co < Cout) wt[((long long)ci * kk + (kk - 1 - tap)) * Cout + co] = tile[threadIdx.x][r];
    }
}

__global__ void pad_channels_kernel(const float* __restrict__ in, int cin, float* __restrict__ out, int cout, size_t rows) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = rows * (size_t)cout;
    if (i >= n) return;
    size_t r = i / cout;
    int c = (int)(i - r * cout);
    out[i] = c < cin ? in[r * cin + c] : 0.f;
}

// ---------------------------------------------------------------------------
// First layer (3 or 3+1 input channels read straight from the reference-facing NCHW tensors, Cout <= 32, 3x3).
// HBM-bound: forward writes 128 B per pixel, the weight gradient reads them once.

// channel c of the (virtually concatenated) NCHW input pair, zero outside the image / beyond C0+C1
__device__ __forceinline__ float in_px(const float* __restrict__ in0, int C0, const float* __restrict__ in1, int C1, int b, int c,
                                       int h, int w, int H, int W) {
    if (h < 0 || h >= H || w < 0 || w >= W) return 0.f;
    if (c < C0) return __ldg(in0 + (((long long)b * C0 + c) * H + h) * W + w);
    if (c < C0 + C1) return __ldg(in1 + (((long long)b * C1 + (c - C0)) * H + h) * W + w);
    return 0.f;
}

// forward: lane = output channel (its 9x4 filter lives in registers), one warp walks along an image row keeping the
// 3x3 input window in registers (3 broadcast shared-memory loads + 36 FMAs per pixel, one coalesced 128-byte store).
// CTA tile = FT_H rows (one per warp) x FT_W columns.
constexpr int FT_H = 8, FT_W = 104;
__global__ void __launch_bounds__(256) conv_first_fwd_kernel(const float* __restrict__ in0, int C0, const float* __restrict__ in1,
                                                             int C1, const float* __restrict__ w /* [Cout][9][4] */,
                                                             float* __restrict__ z, int ldz, int B, int H, int W, int Cout) {
    __shared__ float4 xs[FT_H + 2][FT_W + 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_w = (W + FT_W - 1) / FT_W, tiles_h = (H + FT_H - 1) / FT_H;
    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h;
    const int b = t / tiles_h;
    const int h0 = th * FT_H, w0 = tw * FT_W;
    // stage the input tile (+halo) channel plane by channel plane: one 32-bit-indexed load per element, batches of
    // 4 independent loads per thread for memory-level parallelism
    constexpr int NPX = (FT_H + 2) * (FT_W + 2);
    const int HW = H * W;
    float* xsf = reinterpret_cast<float*>(&xs[0][0]);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        const float* plane = ch < C0 ? in0 + ((long long)b * C0 + ch) * HW
                                     : (ch < C0 + C1 ? in1 + ((long long)b * C1 + (ch - C0)) * HW : nullptr);
        for (int base = 0; base < NPX; base += 4 * 256) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * 256 + tid;
                const int r = i / (FT_W + 2), c = i - r * (FT_W + 2);
                const int h = h0 + r - 1, ww = w0 + c - 1;
                v[u] = (plane && i < NPX && h >= 0 && h < H && ww >= 0 && ww < W) ? __ldg(plane + h * W + ww) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * 256 + tid;
                if (i < NPX) xsf[i * 4 + ch] = v[u];
            }
        }
    }
    float4 wr[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[k] = lane < Cout ? ldg4(w + (lane * 9 + k) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int h = h0 + warp;
    if (h >= H) return;
    const int wn = min(FT_W, W - w0);
    float4 x00 = xs[warp][0], x01 = xs[warp][1], x10 = xs[warp + 1][0], x11 = xs[warp + 1][1], x20 = xs[warp + 2][0], x21 = xs[warp + 2][1];
    float* zr = z + (((long long)b * H + h) * W + w0) * ldz + lane;
#pragma unroll 2
    for (int c = 0; c < wn; ++c) {
        const float4 x02 = xs[warp][c + 2], x12 = xs[warp + 1][c + 2], x22 = xs[warp + 2][c + 2];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent FMA chains (one per input channel)
#define FSDET_TAP(X, Wv) a0 = fmaf(X.x, Wv.x, a0); a1 = fmaf(X.y, Wv.y, a1); a2 = fmaf(X.z, Wv.z, a2); a3 = fmaf(X.w, Wv.w, a3);
        FSDET_TAP(x00, wr[0]) FSDET_TAP(x01, wr[1]) FSDET_TAP(x02, wr[2])
        FSDET_TAP(x10, wr[3]) FSDET_TAP(x11, wr[4]) FSDET_TAP(x12, wr[5])
        FSDET_TAP(x20, wr[6]) FSDET_TAP(x21, wr[7]) FSDET_TAP(x22, wr[8])
#undef FSDET_TAP
        if (lane < Cout) zr[(long long)c * ldz] = (a0 + a1) + (a2 + a3);
        x00 = x01; x01 = x02; x10 = x11; x11 = x12; x20 = x21; x21 = x22;
    }
}

// weight gradient: dw[co][tap][ci] = sum_p dz[p][co] * x[p+tap][ci].  One CTA walks image rows: the dz row and the
// three x rows it needs are staged in shared memory.  Thread (co, filter row ty, column segment) keeps the 3 taps of
// its filter row x 4 input channels in registers and slides a 3-pixel window along its third of the image row
// (one new x pixel + one dz value per step feed 12 FMAs); partials are reduced in a fixed order afterwards.
constexpr int FW_SEG = 3;   // column segments per row (thread groups)
__global__ void __launch_bounds__(288) conv_first_wgrad_kernel(const float* __restrict__ in0, int C0, const float* __restrict__ in1,
                                                               int C1, const float* __restrict__ dz, int lddz,
                                                               float* __restrict__ part, int B, int H, int W, int Cout) {
    extern __shared__ __align__(16) float sm[];
    float4* xs = reinterpret_cast<float4*>(sm);              // [3][W + 2] pixels of 4 channels (zero halo)
    float* ds = sm + 3 * (W + 2) * 4;                        // [W][32]
    const int tid = threadIdx.x;
    const int co = tid & 31;
    const int ty = (tid >> 5) % 3;                           // filter row
    const int seg = (tid >> 5) / 3;                          // column segment 0..FW_SEG-1
    const int wseg = (W + FW_SEG - 1) / FW_SEG;
    const int wbeg = seg * wseg, wend = min(W, wbeg + wseg);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;   // taps (ty, 0), (ty, 1), (ty, 2)
    const long long rows = (long long)B * H;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int b = (int)(row / H), h = (int)(row - (long long)b * H);
        __syncthreads();
        // staging, channel plane by channel plane, in batches of 4 independent loads per thread
        {
            const int HW = H * W, NX = 3 * (W + 2);
            float* xsf = reinterpret_cast<float*>(xs);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const float* plane = ch < C0 ? in0 + ((long long)b * C0 + ch) * HW
                                             : (ch < C0 + C1 ? in1 + ((long long)b * C1 + (ch - C0)) * HW : nullptr);
                for (int base = 0; base < NX; base += 4 * 288) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = base + u * 288 + tid;
                        const int r = i / (W + 2), c = i - r * (W + 2);
                        const int hh = h + r - 1, ww = c - 1;
                        v[u] = (plane && i < NX && hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(plane + hh * W + ww) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = base + u * 288 + tid;
                        if (i < NX) xsf[i * 4 + ch] = v[u];
                    }
                }
            }
        }
        const float* drow = dz + (row * W) * lddz;
        for (int base = 0; base < W * 8; base += 6 * 288) {  // 8 float4 per pixel (32 channels, zero beyond Cout)
            float4 v[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int i = base + u * 288 + tid;
                const int pw = i >> 3, c4 = (i & 7) * 4;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < W * 8 && c4 < Cout) v[u] = ldg4(drow + (long long)pw * lddz + c4);
            }
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int i = base + u * 288 + tid;
                if (i < W * 8) *reinterpret_cast<float4*>(ds + (i >> 3) * 32 + (i & 7) * 4) = v[u];
            }
        }
        __syncthreads();
        const float4* xr = xs + ty * (W + 2);                // xr[w + tx] = x[h + ty - 1][w + tx - 1]
        if (wbeg < wend) {
            float4 x0 = xr[wbeg], x1 = xr[wbeg + 1];
#pragma unroll 4
            for (int w = wbeg; w < wend; ++w) {
                const float4 x2 = xr[w + 2];
                const float d = ds[w * 32 + co];
                a0.x = fmaf(d, x0.x, a0.x); a0.y = fmaf(d, x0.y, a0.y); a0.z = fmaf(d, x0.z, a0.z); a0.w = fmaf(d, x0.w, a0.w);
                a1.x = fmaf(d, x1.x, a1.x); a1.y = fmaf(d, x1.y, a1.y); a1.z = fmaf(d, x1.z, a1.z); a1.w = fmaf(d, x1.w, a1.w);
                a2.x = fmaf(d, x2.x, a2.x); a2.y = fmaf(d, x2.y, a2.y); a2.z = fmaf(d, x2.z, a2.z); a2.w = fmaf(d, x2.w, a2.w);
                x0 = x1; x1 = x2;
            }
        }
    }
    if (co < Cout) {
        float* dst = part + (((long long)blockIdx.x * FW_SEG + seg) * Cout + co) * 36 + ty * 12;
        *reinterpret_cast<float4*>(dst) = a0;
        *reinterpret_cast<float4*>(dst + 4) = a1;
        *reinterpret_cast<float4*>(dst + 8) = a2;
    }
}

static int wgrad_splits(long long M, int Cin, int Cout, int ks, int bmc) {
    int K = ks * ks * Cin;
    long long tiles = (long long)ceil_div(K, 128) * ceil_div(Cout, bmc);
    long long want = (2LL * kNumSMs + tiles - 1) / tiles;
    long long maxs = (M + 255) / 256;  // at least 256 pixels per split
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    return (int)want;
}

}  // namespace fsdet

using namespace fsdet;

extern "C" int fsdet_conv_stat_rows(int npix) { return ceil_div(npix, BM); }

extern "C" int fsdet_conv_fwd(const float* x, int ldx, const float* w, const float* bias, float* z, int ldz,
                              float* stat_partial, int B, int H, int W, int Cin, int Cout, int ksize, int accumulate,
                              void* stream) {
    FSDET_CHECK_ARG(x && w && z, "conv_fwd: null pointer");
    FSDET_CHECK_ARG(ksize == 1 || ksize == 3, "conv_fwd: ksize %d unsupported (1 or 3)", ksize);
    FSDET_CHECK_ARG(Cin > 0 && Cin % 4 == 0 && ldx % 4 == 0 && ldx >= Cin, "conv_fwd: Cin=%d ldx=%d must be multiples of 4", Cin, ldx);
    FSDET_CHECK_ARG(Cout > 0 && ldz >= Cout, "conv_fwd: Cout=%d ldz=%d", Cout, ldz);
    FSDET_CHECK_ARG(aligned16(x) && aligned16(w), "conv_fwd: x/w must be 16-byte aligned");
    FSDET_CHECK_ARG(!(stat_partial && (bias || accumulate)), "conv_fwd: stats are only defined for the plain product");
    ConvArgs a;
    a.x = x; a.w = w; a.bias = bias; a.z = z; a.stat = stat_partial;
    a.ldx = ldx; a.ldz = ldz; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ks = ksize; a.pad = (ksize - 1) / 2;
    a.K = ksize * ksize * Cin; a.M = (long long)B * H * W; a.accumulate = accumulate;
    if (a.M == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 block(256);
    if (Cout > 64) {
        dim3 grid(ceil_div(a.M, BM), ceil_div(Cout, 128));
        conv_igemm_kernel<8><<<grid, block, 0, s>>>(a);
    } else if (Cout > 32) {
        dim3 grid(ceil_div(a.M, BM), 1);
        conv_igemm_kernel<4><<<grid, block, 0, s>>>(a);
    } else {
        dim3 grid(ceil_div(a.M, BM), 1);
        conv_igemm_kernel<2><<<grid, block, 0, s>>>(a);
    }
    return launch_status("conv_fwd");
}

extern "C" size_t fsdet_conv_wgrad_workspace_floats(int B, int H, int W, int Cin, int Cout, int ksize) {
    long long M = (long long)B * H * W;
    int bmc = Cout > 64 ? 128 : 64;
    int splits = wgrad_splits(M, Cin, Cout, ksize, bmc);
    if (splits <= 1) return 0;
    return (size_t)splits * (size_t)Cout * (size_t)(ksize * ksize * Cin);
}

extern "C" int fsdet_conv_wgrad(const float* x, int ldx, const float* dz, int lddz, float* dw, float* workspace,
                                size_t workspace_floats, int B, int H, int W, int Cin, int Cout, int ksize, void* stream) {
    FSDET_CHECK_ARG(x && dz && dw, "conv_wgrad: null pointer");
    FSDET_CHECK_ARG(ksize == 1 || ksize == 3, "conv_wgrad: ksize %d unsupported", ksize);
    FSDET_CHECK_ARG(Cin % 4 == 0 && Cout % 4 == 0 && ldx % 4 == 0 && lddz % 4 == 0,
                    "conv_wgrad: Cin=%d Cout=%d ldx=%d lddz=%d must be multiples of 4", Cin, Cout, ldx, lddz);
    FSDET_CHECK_ARG(aligned16(x) && aligned16(dz) && aligned16(dw), "conv_wgrad: pointers must be 16-byte aligned");
    WgradArgs a;
    a.x = x; a.dz = dz; a.ldx = ldx; a.lddz = lddz; a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
    a.ks = ksize; a.pad = (ksize - 1) / 2; a.K = ksize * ksize * Cin; a.M = (long long)B * H * W;
    int bmc = Cout > 64 ? 128 : 64;
    int splits = wgrad_splits(a.M, Cin, Cout, ksize, bmc);
    size_t need = splits > 1 ? (size_t)splits * Cout * a.K : 0;
    FSDET_CHECK_ARG(workspace_floats >= need && (need == 0 || (workspace && aligned16(workspace))),
                    "conv_wgrad: workspace too small (%zu < %zu floats)", workspace_floats, need);
    long long pps = (a.M + splits - 1) / splits;
    pps = (pps + BK - 1) / BK * BK;
    a.pix_per_split = pps;
    a.out = splits > 1 ? workspace : dw;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(ceil_div(a.K, 128), ceil_div(Cout, bmc), splits);
    if (bmc == 128) conv_wgrad_kernel<128><<<grid, 256, 0, s>>>(a);
    else conv_wgrad_kernel<64><<<grid, 256, 0, s>>>(a);
    int st = launch_status("conv_wgrad");
    if (st) return st;
    if (splits > 1) {
        long long n4 = (long long)Cout * a.K / 4;
        splitk_reduce_kernel<<<ceil_div(n4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(workspace),
                                                              reinterpret_cast<float4*>(dw), n4, splits);
        st = launch_status("conv_wgrad_reduce");
    }
    return st;
}

extern "C" int fsdet_weight_flip_transpose(const float* w, float* wt, int Cout, int kk, int Cin, void* stream) {
    FSDET_CHECK_ARG(w && wt && Cout > 0 && Cin > 0 && kk > 0, "weight_flip_transpose: bad args");
    dim3 grid(ceil_div(Cin, 32), ceil_div(Cout, 32), kk), block(32, 8);
    weight_flip_transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(w, wt, Cout, kk, Cin);
    return launch_status("weight_flip_transpose");
}

extern "C" int fsdet_pad_channels(const float* in, int cin, float* out, int cout, size_t rows, void* stream) {
    FSDET_CHECK_ARG(in && out && cin > 0 && cout > 0, "pad_channels: bad args");
    size_t n = rows * (size_t)cout;
    if (n == 0) return 0;
    pad_channels_kernel<<<ceil_div((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(in, cin, out, cout, rows);
    return launch_status("pad_channels");
}

static int first_wgrad_ctas(int B, int H) {
    long long rows = (long long)B * H;
    long long n = 3LL * kNumSMs;
    return (int)(rows < n ? rows : n);
}

extern "C" size_t fsdet_conv_first_wgrad_workspace_floats(int B, int H, int W, int Cout) {
    (void)W;
    return (size_t)first_wgrad_ctas(B, H) * FW_SEG * Cout * 36;
}

extern "C" int fsdet_conv_first_fwd(const float* in0, int C0, const float* in1, int C1, const float* w_pad4, float* z, int ldz,
                                    int B, int H, int W, int Cout, void* stream) {
    FSDET_CHECK_ARG(in0 && w_pad4 && z && C0 > 0 && C1 >= 0 && (C1 == 0 || in1) && C0 + C1 <= 4, "conv_first_fwd: bad inputs");
    FSDET_CHECK_ARG(Cout > 0 && Cout <= 32 && Cout % 4 == 0 && ldz % 4 == 0 && aligned16(z) && aligned16(w_pad4),
                    "conv_first_fwd: Cout=%d ldz=%d", Cout, ldz);
    long long tiles = (long long)B * ceil_div(H, FT_H) * ceil_div(W, FT_W);
    if (tiles == 0) return 0;
    FSDET_CHECK_ARG(tiles < (1ll << 31), "conv_first_fwd: too many tiles");
    conv_first_fwd_kernel<<<(unsigned)tiles, 256, 0, (cudaStream_t)stream>>>(in0, C0, in1, C1, w_pad4, z, ldz, B, H, W, Cout);
    return launch_status("conv_first_fwd");
}

extern "C" int fsdet_conv_first_wgrad(const float* in0, int C0, const float* in1, int C1, const float* dz, int lddz, float* dw,
                                      float* workspace, size_t workspace_floats, int B, int H, int W, int Cout, void* stream) {
    FSDET_CHECK_ARG(in0 && dz && dw && workspace && C0 > 0 && C1 >= 0 && (C1 == 0 || in1) && C0 + C1 <= 4,
                    "conv_first_wgrad: bad inputs");
    FSDET_CHECK_ARG(Cout > 0 && Cout <= 32 && Cout % 4 == 0 && lddz % 4 == 0, "conv_first_wgrad: Cout=%d lddz=%d", Cout, lddz);
    FSDET_CHECK_ARG(aligned16(dz) && aligned16(dw) && aligned16(workspace), "conv_first_wgrad: alignment");
    const int ctas = first_wgrad_ctas(B, H);
    FSDET_CHECK_ARG(workspace_floats >= (size_t)ctas * FW_SEG * Cout * 36, "conv_first_wgrad: workspace too small");
    if (ctas == 0) return 0;
    const size_t smem = ((size_t)3 * (W + 2) * 4 + (size_t)W * 32) * sizeof(float);
    FSDET_CHECK_ARG(smem <= 200 * 1024, "conv_first_wgrad: image width %d too large", W);
    cudaStream_t s = (cudaStream_t)stream;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("conv_first_wgrad: %s", cudaGetErrorString(e)); return (int)e; }
        smem_set = smem;
    }
    conv_first_wgrad_kernel<<<ctas, 288, smem, s>>>(in0, C0, in1, C1, dz, lddz, workspace, B, H, W, Cout);
    int st = launch_status("conv_first_wgrad");
    if (st) return st;
    long long n4 = (long long)Cout * 36 / 4;
    splitk_reduce_kernel<<<ceil_div(n4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(dw),
                                                          n4, ctas * FW_SEG);
    return launch_status("conv_first_wgrad_reduce");
}
