// Persistent variant of the tcgen05 convolution (EXPERIMENTAL, off by default) - included by conv_tc.cu after the PTX
// wrappers, and by tools/host_emul/conv_persist_emul.cpp after FUNCTIONAL MODELS of the same wrappers (mbarrier,
// TMA, UMMA, TMEM as host code), which is how its control flow - barrier phases, tile sequencing, accumulator
// double buffering, epilogue staging - is tested on the CPU (tests/test_conv_persist_host_emul.py).
//
// Same tiles, operands and arithmetic as conv_tc_kernel's short-K flavour, but ONE CTA per SM walks a list of
// output tiles: the TMA producer and the MMA issuer run ahead across tile boundaries (no pipeline drain, no
// per-tile barrier / TMEM set-up), and the accumulators are double buffered in TMEM so that the epilogue of tile i
// (tcgen05.ld -> swizzled shared memory -> TMA store) overlaps the MMAs of tile i+1.  Motivation
// (profiles/conv_layers_r01.md): conv2 (32->64 channels at 208x208) launches 21,632 one-tile CTAs and runs at
// 199 / 104 TFLOP/s (forward / dgrad) against 480 for the long-K layers.
// Enabled with FSDET_TC_PERSIST=1; written in round 1 after the GPU budget was spent, NOT yet run on a GPU:
// tools/try_persist.sh is the first thing to run (under `timeout`) before trusting it.
#pragma once

#ifdef FSDET_HOST_EMULATION
#define FSDET_TC_DYN_SMEM(name) uint8_t* name = emul::g_dyn_smem
#else
#define FSDET_TC_DYN_SMEM(name) extern __shared__ uint8_t name[]
#endif

template <int BN, int BK, int STAGES_, int NH>
struct TcPersistCfg {
    static constexpr int ROW_BYTES = BK * 2;
    static constexpr int A_BYTES = TC_BM * ROW_BYTES;
    static constexpr int B_BYTES = BN * ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = STAGES_;
    static constexpr int ACC_COLS = (NH + 1) * BN;                 // one accumulator set
    static constexpr int TMEM_COLS = tmem_cols(2 * ACC_COLS);       // two sets
    static constexpr int EPI_BYTES = 4 * 2 * 4096;                  // 4 epilogue warps x two 32x32 fp32 staging tiles
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int BK, int STAGES_, int NH>
__global__ void __launch_bounds__(192, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                       const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo,
                       const __grid_constant__ CUtensorMap tmZ, const TcArgs p, const int tiles_n, const int tiles_total) {
    using Cfg = TcPersistCfg<BN, BK, STAGES_, NH>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int A_BYTES = Cfg::A_BYTES;
    static_assert(2 * Cfg::ACC_COLS <= 512, "two accumulator sets must fit in TMEM");
    FSDET_TC_DYN_SMEM(smem_raw);
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + Cfg::EPI_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;      // [2] MMA issuer -> epilogue
    uint64_t* acc_empty = acc_full + 2;           // [2] epilogue (4 warps) -> MMA issuer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kchunks = p.Cin / BK;
    const int nk = p.ks * p.ks * kchunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmAhi);
        tma_prefetch_desc(&tmAlo);
        tma_prefetch_desc(&tmBhi);
        tma_prefetch_desc(&tmBlo);
        tma_prefetch_desc(&tmZ);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, (uint32_t)Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int HW = p.H * p.W;
            unsigned it = 0;                                   // k-blocks issued so far (all tiles)
            for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x) {
                const int n_tile = tile % tiles_n;
                const long long m0 = (long long)(tile / tiles_n) * TC_BM;
                const int img = (int)(m0 / HW);
                const int rem = (int)(m0 - (long long)img * HW);
                const int ph = rem / p.W, pw = rem - ph * p.W;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    const int tap = kb / kchunks;
                    const int c0 = (kb - tap * kchunks) * BK;
                    const int r = tap / p.ks, sx = tap - r * p.ks;
                    tma_load_im2col_4d(st, &tmAhi, &full_bar[s], c0, pw - p.pad, ph - p.pad, img, (uint16_t)sx, (uint16_t)r);
                    tma_load_im2col_4d(st + A_BYTES, &tmAlo, &full_bar[s], c0, pw - p.pad, ph - p.pad, img, (uint16_t)sx, (uint16_t)r);
                    tma_load_2d(st + 2 * A_BYTES, &tmBhi, &full_bar[s], tap * p.cpitch + c0, n_tile * BN);
                    tma_load_2d(st + 2 * A_BYTES + Cfg::B_BYTES, &tmBlo, &full_bar[s], tap * p.cpitch + c0, n_tile * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            unsigned it = 0;
            unsigned t = 0;                                    // tiles done by this CTA
            for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++t) {
                const unsigned a = t & 1u;
                mbar_wait(&acc_empty[a], ((t >> 1) & 1u) ^ 1u);    // the epilogue has drained this accumulator set
                tc_fence_after();
                const uint32_t acc = tmem_base + a * (uint32_t)Cfg::ACC_COLS;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&full_bar[s], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    uint64_t ahi, alo, bhi, blo;
                    if constexpr (BK == 64) {
                        ahi = umma_desc_k_sw128(sa); alo = umma_desc_k_sw128(sa + A_BYTES);
                        bhi = umma_desc_k_sw128(sa + 2 * A_BYTES); blo = umma_desc_k_sw128(sa + 2 * A_BYTES + Cfg::B_BYTES);
                    } else {
                        ahi = umma_desc_k_sw64(sa); alo = umma_desc_k_sw64(sa + A_BYTES);
                        bhi = umma_desc_k_sw64(sa + 2 * A_BYTES); blo = umma_desc_k_sw64(sa + 2 * A_BYTES + Cfg::B_BYTES);
                    }
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        const uint32_t dhi = acc + (uint32_t)((kb % NH) * BN);
                        const uint32_t dlo = acc + (uint32_t)(NH * BN);
                        umma_f16(dhi, ahi + adv, bhi + adv, idesc, (kb >= NH || k > 0) ? 1u : 0u);
                        umma_f16(dlo, alo + adv, bhi + adv, idesc, (kb | k) ? 1u : 0u);
                        umma_f16(dlo, ahi + adv, blo + adv, idesc, 1u);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&acc_full[a]);
            }
        }
    } else {
        const int quarter = warp & 3;
        const float inv = 1.f / (scale_from_amax(p.amax_a ? ldg_f32(p.amax_a) : 0.f) * scale_from_amax(p.amax_b ? ldg_f32(p.amax_b) : 0.f));
        const int nhi = nk < NH ? nk : NH;
        uint8_t* stage_buf = epi + quarter * 8192;
        unsigned t = 0, stores = 0;                            // tiles done, TMA stores issued by this warp
        for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, ++t) {
            const unsigned a = t & 1u;
            const int n_tile = tile % tiles_n;
            const long long m0 = (long long)(tile / tiles_n) * TC_BM;
            const long long mrow = m0 + quarter * 32;
            mbar_wait(&acc_full[a], (t >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                uint32_t r[32];
                float acc[32];
                const uint32_t taddr = tmem_base + a * (uint32_t)Cfg::ACC_COLS + ((uint32_t)(quarter * 32) << 16) + ch * 32;
                tmem_ld32(taddr + NH * BN, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
                for (int h = nhi - 1; h >= 0; --h) {
                    tmem_ld32(taddr + h * BN, r);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
                }
                const int n0 = n_tile * BN + ch * 32;
                if (n0 < p.Cout && mrow < p.M) {               // warp-uniform
                    uint8_t* buf = stage_buf + (stores & 1u) * 4096;
                    if (stores >= 2) {                         // the store that last read this buffer must have drained
                        if (lane == 0) tma_store_wait_read<1>();
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 v = make_float4(acc[4 * j] * inv, acc[4 * j + 1] * inv, acc[4 * j + 2] * inv, acc[4 * j + 3] * inv);
                        *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.accumulate) tma_reduce_add_2d(&tmZ, buf, n0, (int)mrow);
                        else tma_store_2d(&tmZ, buf, n0, (int)mrow);
                        tma_store_commit();
                    }
                    ++stores;
                }
            }
            // this warp's TMEM reads of set `a` are complete (tcgen05.wait::ld in tmem_ld32): hand the set back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)Cfg::TMEM_COLS);
    }
}
