// conv_gemm_ts.cu -- the implicit-GEMM convolution of conv_gemm.cu with the operand roles swapped and the ACTIVATIONS IN TENSOR MEMORY:
//
//     D[pixel, co] = sum_{tap, ci} X[pixel + shift(tap), ci] * Wt[co, tap, ci]          A = activations (M = 128 pixels), B = weights (N = 128)
//
// conv_gemm.cu keeps both operands in shared memory.  For kind::tf32 a 128 x 128 x 8 MMA reads 4 KB of A and 4 KB of B and takes
// ~67 cycles, i.e. 122 of the 128 bytes per cycle an SM's shared memory delivers -- before the producers' own stores (32 KB per K
// block) and the weight records' TMA writes (32 KB) are counted.  ncu showed that kernel at 52 % tensor-pipe utilisation with the
// shared-memory pipe saturated.  Here the producer threads write the split (hi / lo) activations straight from registers into TENSOR
// MEMORY (tcgen05.st, thread = pixel = TMEM lane, columns = the K block's 32 channels) and the MMAs take their A operand from there
// (tcgen05.mma [d_tmem], [a_tmem], b_desc): shared memory carries the packed weight records only -- 48 KB of MMA reads + 32 KB of TMA
// writes per K block instead of 96 + 32 + 32 -- no generic-proxy stores, no proxy fence, and 5 weight stages instead of 3.
//
// The accumulator comes out pixel-major (lane = pixel, column = output channel), so the epilogue needs no transpose: for a fixed
// channel the 32 lanes of a warp store 32 consecutive pixels of the NCHW plane.  Packed weight records (hdn_conv_pack_weight_f32) are
// the canonical K-major tile either way and are used as they are.  Chunked accumulation (two TMEM accumulators, fp32 folds in
// registers), stride / padding / dilation handling and the programmatic dependent launch are those of conv_gemm.cu.
//
// TMEM budget (512 columns): 2 accumulators x 128 + 3 activation stages x (32 hi + 32 lo) = 448.
// Used for the large launches (>= 2 CTAs per SM worth of 128 x 128 tiles); split-K clusters and the projection epilogue stay in
// conv_gemm.cu.
#include "umma.cuh"

namespace hdn {

constexpr int TS_BN = 128;       // output channels per tile = rows of one packed weight record
constexpr int TS_ASTAGES = 3;    // activation operand stages in TMEM
constexpr int TS_WSTAGES = 5;    // weight records in shared memory
constexpr int TS_ACOL0 = 2 * TS_BN;                      // first activation column
constexpr int TS_W_TILE = CG_BM * CG_BK * 4;             // one operand tile (hi or lo) of a weight record, bytes
constexpr size_t TS_SMEM = (size_t)TS_WSTAGES * 2 * TS_W_TILE + 1024;

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
__device__ __forceinline__ void ts_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (320 threads are allocated as 12 warps' worth of registers, so 168 per thread is the ceiling: __maxnreg__(200) compiles without
// spills but fails to launch.  The projection epilogue of conv_gemm.cu was ported to this operand placement as well and measured no
// faster -- those launches are epilogue-bound -- so it stays where it is.)
__global__ void __launch_bounds__(CG_THREADS + 64, 1) conv_gemm_ts_kernel(const __grid_constant__ ConvGemmArgs a) {
    constexpr uint32_t W_SBO = 128, W_LBO = (CG_BM / 8) * 128;  // the packed record's K-major tile: 8-row groups 128 B apart, K chunks of 4 W_LBO apart
    // kind::tf32, fp32 accumulate, A and B K-major, M = 128 (pixels), N = 128 (channels)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TS_BN >> 3) << 17) | ((uint32_t)(CG_BM >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t w_full[TS_WSTAGES], w_free[TS_WSTAGES], a_full[TS_ASTAGES], a_free[TS_ASTAGES], acc_full[2], acc_free[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_scale[TS_BN], s_shift[TS_BN];  // the tile's folded BatchNorm (constants of the layer: read before the dependency wait)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HW = a.H * a.W, HWo = a.Ho * a.Wo;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // see conv_gemm.cu
    // SPLIT-K over a thread-block cluster (tracking batch sizes, see conv_gemm.cu): the `splitk` CTAs of a cluster own one output tile
    // and 1/splitk of the K blocks each; the leader adds the peers' fp32 partial tiles through distributed shared memory, in rank order.
    const int S = a.splitk > 1 ? a.splitk : 1;
    uint32_t crank = 0;
    if (S > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int pix0 = (S > 1 ? (int)blockIdx.x / S : (int)blockIdx.x) * CG_BM, co0 = blockIdx.y * TS_BN, prob = blockIdx.z / a.B,
              img = blockIdx.z - prob * a.B;
    const int nkb_all = a.taps * a.Cin / CG_BK;
    const int nkb = nkb_all / S;       // K blocks of THIS CTA ...
    const int kb0 = (int)crank * nkb;  // ... starting at global block kb0
    const int nchunks = (nkb + CG_KCB - 1) / CG_KCB;

    if (tid == 0) {
        for (int i = 0; i < TS_WSTAGES; ++i) {
            mbar_init(&w_full[i], 1);  // the loader's arrive.expect_tx
            mbar_init(&w_free[i], 1);  // tcgen05.commit
        }
        for (int i = 0; i < TS_ASTAGES; ++i) {
            mbar_init(&a_full[i], CG_THREADS / 32);  // one arrival per producer warp
            mbar_init(&a_free[i], 1);                // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], CG_THREADS / 32);
        }
        mbar_fence_init();
    }
    if (tid >= 64 && tid < 64 + TS_BN) {
        const int c = tid - 64, co = co0 + c;
        const float *scp = a.scale[prob], *shp = a.shift[prob];
        s_scale[c] = scp && co < a.Cout ? __ldg(scp + co) : 1.f;
        s_shift[c] = shp && co < a.Cout ? __ldg(shp + co) : 0.f;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == CG_THREADS / 32) {
        // ============================== MMA issuer (one elected lane) ==============================
        if (elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int ws = kb % TS_WSTAGES, as = kb % TS_ASTAGES, chunk = kb / CG_KCB;
                if (kb % CG_KCB == 0 && chunk >= 2) mbar_wait(&acc_free[chunk & 1], ((chunk >> 1) - 1) & 1);  // producers drained this accumulator
                mbar_wait(&w_full[ws], (kb / TS_WSTAGES) & 1);
                mbar_wait(&a_full[as], (kb / TS_ASTAGES) & 1);
                tc_fence_after();
                const uint32_t sw_hi = smem_u32(smem + ws * 2 * TS_W_TILE), sw_lo = sw_hi + TS_W_TILE;
                const uint32_t ta_hi = tmem_d + (uint32_t)(TS_ACOL0 + as * 2 * CG_BK), ta_lo = ta_hi + CG_BK;
                const uint32_t acc = tmem_d + (uint32_t)((chunk & 1) * TS_BN);
#pragma unroll
                for (int ks = 0; ks < CG_BK / 8; ++ks) {
                    const uint64_t dwh = umma_smem_desc(sw_hi + ks * 2 * W_LBO, W_LBO, W_SBO), dwl = umma_smem_desc(sw_lo + ks * 2 * W_LBO, W_LBO, W_SBO);
                    umma_tf32_ts(acc, ta_lo + ks * 8, dwh, IDESC, ((kb % CG_KCB) | ks) != 0);  // small terms first; a chunk's first MMA overwrites
                    umma_tf32_ts(acc, ta_hi + ks * 8, dwl, IDESC, 1);
                    umma_tf32_ts(acc, ta_hi + ks * 8, dwh, IDESC, 1);
                }
                umma_commit(&w_free[ws]);  // arrive when the MMAs above have finished reading the weight stage ...
                umma_commit(&a_free[as]);  // ... and the activation stage
                if (kb % CG_KCB == CG_KCB - 1 || kb == nkb - 1) umma_commit(&acc_full[chunk & 1]);
            }
        }
        __syncwarp();
        if (S > 1) { ts_cluster_sync(); ts_cluster_sync(); }  // the producers' two split-K barriers (every thread of the cluster takes part)
    } else if (warp == CG_THREADS / 32 + 1) {
        // ============================== weight loader: one 32 KB TMA bulk copy per K block (constants: no dependency wait) ==============================
        if (elect_one()) {
            const float *wsrc = a.wpk[prob] + ((size_t)blockIdx.y * nkb_all + kb0) * (2 * TS_W_TILE / 4);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TS_WSTAGES;
                if (kb >= TS_WSTAGES) mbar_wait(&w_free[s], ((kb / TS_WSTAGES) - 1) & 1);
                mbar_expect_tx(&w_full[s], 2 * TS_W_TILE);
                bulk_g2s(smem + s * 2 * TS_W_TILE, wsrc + (size_t)kb * (2 * TS_W_TILE / 4), 2 * TS_W_TILE, &w_full[s]);
            }
        }
        __syncwarp();
        if (S > 1) { ts_cluster_sync(); ts_cluster_sync(); }
    } else {
        // ============================== producers: thread = pixel (TMEM lane), 16 of the K block's 32 channels ==============================
        const float *xb = a.x[prob] + (size_t)img * a.Cin * HW;
        const int row = (warp & 3) * 32 + lane;  // TMEM lane = pixel of the tile; a warp may only touch its own lane quadrant
        const int kh = warp >> 2;                // channel half of the K block: channels kh*16 .. kh*16 + 15
        const int bp = pix0 + row;
        const int b_r = bp < HWo ? (bp / a.Wo) * a.stride + a.off : -(1 << 20);  // beyond the plane: never valid
        const int b_c = bp < HWo ? (bp - (bp / a.Wo) * a.Wo) * a.stride + a.off : 0;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;

        // fp32 register accumulator of this thread's share of the tile: its pixel row, channels [col_lo, col_lo + 64)
        constexpr int HALF = TS_BN / 2;
        const int col_lo = kh * HALF;
        float racc[HALF];
#pragma unroll
        for (int e = 0; e < HALF; ++e) racc[e] = 0.f;
        auto drain = [&](int chunk) {
            mbar_wait(&acc_full[chunk & 1], (chunk >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_d + (uint32_t)((chunk & 1) * TS_BN) + lane_base + (uint32_t)(col_lo + c0), v);
#pragma unroll
                for (int e = 0; e < 16; ++e) racc[c0 + e] += v[e];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[chunk & 1]);
        };

        // A block's 16 loads are  ublk[toff[e]]:  ublk = warp-uniform base of (channel block, tap shift), advanced incrementally;
        // toff = 16 per-thread element offsets fixed for the whole kernel.  The offsets are made opaque to the compiler: left to
        // itself it rematerialises them inside the loop as an add + LEA + LEA.HI.X chain (3 instructions per load, 19 % of the
        // kernel's issue slots in ncu's source view) instead of keeping 16 registers and issuing one IMAD.WIDE per load.
        unsigned toff[16];
        {
            const unsigned pix_off = bp < HWo ? (unsigned)(b_r * a.W + b_c) : 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                toff[e] = (unsigned)(kh * 16 + e) * (unsigned)HW + pix_off;
                asm volatile("" : "+r"(toff[e]));
            }
        }
        int ld_ci0, ld_ty, ld_tx;  // position of the NEXT block to load (blocks are loaded in order)
        {
            const int k0 = kb0 * CG_BK, tap = k0 / a.Cin;
            ld_ci0 = k0 - tap * a.Cin;
            ld_ty = tap / 3;
            ld_tx = tap - 3 * ld_ty;
        }
        auto load_block = [&](float (&v)[16]) {
            const int dy = a.taps == 1 ? 0 : (ld_ty - 1) * a.dil, dx = a.taps == 1 ? 0 : (ld_tx - 1) * a.dil;
            const bool ok = (unsigned)(b_r + dy) < (unsigned)a.H && (unsigned)(b_c + dx) < (unsigned)a.W;
            const float *ublk = xb + ((long long)ld_ci0 * HW + dy * a.W + dx);  // the same for every thread of the CTA
            asm volatile("" : "+l"(ublk));
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = ok ? __ldg(ublk + toff[e]) : 0.f;
            ld_ci0 += CG_BK;
            if (ld_ci0 == a.Cin) {
                ld_ci0 = 0;
                if (++ld_tx == 3) { ld_tx = 0; ++ld_ty; }
            }
        };
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the producing kernel(s) have completed and flushed: activations may be read
        int drained = 0;
        auto stage_block = [&](int kb, const float (&v)[16]) {
            const int s = kb % TS_ASTAGES;
            if (kb >= TS_ASTAGES) mbar_wait(&a_free[s], ((kb / TS_ASTAGES) - 1) & 1);  // the MMAs that read this stage have retired
            // fold a finished chunk only once a block past it is known to have retired (the wait above): never stalls on the tensor core
            if (drained < nchunks && kb >= (drained + 1) * CG_KCB - 1 + TS_ASTAGES) drain(drained++);
            float hi[16], lo[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) split_tf32(v[e], hi[e], lo[e]);
            tc_fence_after();
            const uint32_t ta = tmem_d + lane_base + (uint32_t)(TS_ACOL0 + s * 2 * CG_BK + kh * 16);
            tmem_st16(ta, hi);
            tmem_st16(ta + CG_BK, lo);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
        };
        float va[16], vb[16], vc[16];  // three register sets rotate: two blocks of loads are in flight behind the one in hand
        load_block(va);
        if (nkb > 1) load_block(vb);
        for (int kb = 0; kb < nkb; kb += 3) {
            if (kb + 2 < nkb) load_block(vc);
            stage_block(kb, va);
            if (kb + 1 < nkb) {
                if (kb + 3 < nkb) load_block(va);
                stage_block(kb + 1, vb);
            }
            if (kb + 2 < nkb) {
                if (kb + 4 < nkb) load_block(vb);
                stage_block(kb + 2, vc);
            }
        }
        while (drained < nchunks) drain(drained++);

        if (S > 1) {
            // ---- split-K: the peers park their partial tile in the (dead) weight stages as ys[channel][pixel]; the leader adds them to
            //      its registers through distributed shared memory, ranks in ascending order (deterministic) ----
            float *ys = reinterpret_cast<float *>(smem);
            if (crank != 0) {
#pragma unroll
                for (int e = 0; e < HALF; ++e) ys[(col_lo + e) * CG_BM + row] = racc[e];
            }
            ts_cluster_sync();
            if (crank == 0) {
                const uint32_t ys_s = smem_u32(ys) + (uint32_t)(col_lo * CG_BM + row) * 4u;
                for (int peer = 1; peer < S; ++peer) {
                    uint32_t remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(ys_s), "r"(peer));
#pragma unroll
                    for (int e = 0; e < HALF; ++e) {
                        float v;
                        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote + (uint32_t)(e * CG_BM * 4)) : "memory");
                        racc[e] += v;
                    }
                }
            }
            ts_cluster_sync();  // nobody leaves (and frees its shared memory) before the leader has read every partial tile
        }

        // ---- epilogue: lane = pixel, so a warp's store of one channel is 32 consecutive pixels of the NCHW plane ----
        if (crank == 0 && bp < HWo && co0 + col_lo < a.Cout) {  // (Cout = 64 layers run in a zero-padded 128-row record: their upper half stores nothing)
            const float *resp = a.residual[prob];
            const size_t base = ((size_t)img * a.Cout + co0 + col_lo) * HWo + bp;
            float *o = a.out[prob] + base;
            const float *ssc = s_scale + col_lo, *ssh = s_shift + col_lo;
            const bool relu = a.relu != 0;
            if (resp) {
                const float *r = resp + base;
#pragma unroll
                for (int e = 0; e < HALF; ++e) {  // fully unrolled: racc stays in registers
                    float y = fmaf(racc[e], ssc[e], ssh[e]) + __ldg(r);
                    *o = relu ? fmaxf(y, 0.f) : y;
                    o += HWo, r += HWo;
                }
            } else {
#pragma unroll
                for (int e = 0; e < HALF; ++e) {
                    const float y = fmaf(racc[e], ssc[e], ssh[e]);
                    *o = relu ? fmaxf(y, 0.f) : y;
                    o += HWo;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u));
}

extern int g_conv_pdl;  // conv_gemm.cu

int launch_conv_gemm_ts(const ConvGemmArgs &a, int nprob, cudaStream_t st) {
    static_assert(TS_SMEM <= 227 * 1024, "shared memory budget");
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(conv_gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM); })) return e;
    cudaLaunchConfig_t cfg{};
    const int S = a.splitk > 1 ? a.splitk : 1;
    cfg.gridDim = dim3((a.Ho * a.Wo + CG_BM - 1) / CG_BM * S, (a.Cout + TS_BN - 1) / TS_BN, a.B * nprob);
    cfg.blockDim = dim3(CG_THREADS + 64);
    cfg.dynamicSmemBytes = TS_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (S > 1) {  // a cluster of S CTAs along x per output tile
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = S;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (g_conv_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm_ts_kernel, a);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    return launch_status();
}

}  // namespace hdn
