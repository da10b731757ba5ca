// conv_shift.cu -- 3x3 'valid' convolutions (the BAN heads' conv_search / conv_kernel layers, hdn/models/head/ban.py:56-61) on
// tcgen05 with the ACTIVATIONS STAGED ONCE PER CHANNEL BLOCK: the nine taps are nine shifted windows of one shared-memory tile.
//
// conv_gemm.cu is an implicit GEMM that re-reads and re-splits the activation tile for every (tap, channel block): A/B runs with the
// MMAs removed showed its `conv_search` launches -- 75 % of the fused head chain -- to be bound by that producer path, not by the
// tensor core (15.8 ms with all MMAs, 13.8 with one of three, 12.6 with none).  Here the GEMM's N axis is the input-width grid
//     q = r * W + c        (c < W: the Wo = W - 2 valid columns of output row r plus 2 columns nobody stores)
// so that output pixel q under tap (ty, tx) reads input pixel q + ty * W + tx: a pure shift.  The canonical no-swizzle K-major UMMA
// layout with SBO = 128 keeps pixel rows 16 bytes apart LINEARLY (row n of a k-group at n * 16), so the B operand of a tap is the
// same shared-memory tile addressed (ty * W + tx) * 16 bytes further on -- a different descriptor start address, no data movement.
// Per 32-channel block the 256 producer threads load BN + 2W + 2 pixels x 32 channels once (coalesced 4-byte loads, one pixel per
// thread), split them hi / lo (3xTF32) and store them K-major; the tensor core then runs 9 taps x 4 k-steps x 3 MMAs off that tile
// while the next channel block is being staged into the other buffer.  Producer work per MMA drops ~4.5x; weights stream through a
// 3-stage ring of the same packed 32 KB records conv_gemm.cu uses (record index tap * Cin/32 + channel block, one TMA bulk copy).
//
// Accumulation chunks are channel blocks (288 of K): the MMAs of block cb accumulate in TMEM accumulator cb & 1; the producers fold
// it into fp32 registers (round-to-nearest adds, see conv_gemm.cu) right before they stage block cb + 2 -- i.e. once the barrier that
// frees B buffer cb & 1 says those MMAs have retired -- so the fold never waits on the tensor core and needs no barrier of its own.
#include "umma.cuh"

namespace hdn {

constexpr int CS_BN = 128;       // output pixels (on the input-width grid) per CTA
constexpr int CS_A_TILE = CG_BM * CG_BK * 4;           // one operand tile (hi or lo) of the weights
constexpr uint32_t CS_A_SBO = 128, CS_A_LBO = (CG_BM / 8) * 128;
constexpr uint32_t CS_B_SBO = 128;
// Two shapes: WIN = staged pixels per channel block (BN + halo, halo = 2 * W + 2), ASTAGES = depth of the weight ring that the
// rest of the 227 KB buys.  W <= 63 (61x61 conv_search at 256/512 crops): 256 / 3;  W <= 31 (the tracker's native crops, the
// log-polar branch, the template side): 192 / 4.
template <int WIN, int ASTAGES>
struct CSCfg {
    static constexpr int B_TILE = CG_BK * WIN * 4;  // one operand tile (hi or lo) of the staged window
    static constexpr uint32_t B_LBO = WIN * 16;     // k-groups of 4 are WIN pixel rows apart
    static constexpr size_t SMEM = 2 * 2 * (size_t)B_TILE + ASTAGES * 2 * (size_t)CS_A_TILE + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert((size_t)(CG_BM * (CS_BN + 1)) * 4 <= 2 * 2 * (size_t)B_TILE, "epilogue staging must fit the window buffers");
};

// ---- cluster helpers (MC = a cluster of two CTAs on neighbouring pixel tiles shares every weight record through TMA multicast) ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {  // arrive on the barrier at the same offset in CTA `cta` of the cluster
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {  // wait for arrivals that may come from the peer CTA
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAITC:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONEC;\n"
        "bra LAB_WAITC;\n"
        "DONEC:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
// global -> the SAME shared-memory offset in every CTA of `mask`; each destination's mbarrier (same offset) gets the complete_tx
__device__ __forceinline__ void bulk_g2s_multicast(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <bool MC, int CS_WIN, int CS_ASTAGES>
__global__ void __launch_bounds__(CG_THREADS + 64, 1) conv3x3_shift_kernel(const __grid_constant__ ConvGemmArgs a) {
    constexpr int BN = CS_BN;
    constexpr int CS_B_TILE = CSCfg<CS_WIN, CS_ASTAGES>::B_TILE;
    constexpr uint32_t CS_B_LBO = CSCfg<CS_WIN, CS_ASTAGES>::B_LBO;
#ifdef HDN_EXP_HALFN  // timing experiment only (wrong results): the same number of MMA instructions, half the math each
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((BN / 2) >> 3) << 17) | ((uint32_t)(CG_BM >> 4) << 24);
#else
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CG_BM >> 4) << 24);
#endif
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *smem_b = smem;                                  // [2 buffers][hi | lo][k/4 (8)][WIN pixels][4]
    unsigned char *smem_a = smem + 2 * 2 * CS_B_TILE;              // [ASTAGES][hi | lo] packed weight records
    __shared__ uint64_t bar_bfull[2], bar_bdone[2], bar_afull[CS_ASTAGES], bar_afree[CS_ASTAGES], bar_peer[CS_ASTAGES];
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = a.W, HW = a.H * a.W, Ho = a.Ho, Wo = a.Wo;
    const int q0 = blockIdx.x * BN, co0 = blockIdx.y * CG_BM, prob = blockIdx.z / a.B, img = blockIdx.z - prob * a.B;
    const int ncb = a.Cin / CG_BK;         // channel blocks
    const int win = BN + 2 * W + 2;        // pixels actually staged (<= CS_WIN)

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_bfull[i], CG_THREADS / 32);  // one arrival per producer warp
            mbar_init(&bar_bdone[i], 1);                // tcgen05.commit after the last tap of the channel block
        }
        for (int i = 0; i < CS_ASTAGES; ++i) {
            mbar_init(&bar_afull[i], 1);                // the loader's arrive.expect_tx
            mbar_init(&bar_afree[i], 1);
            mbar_init(&bar_peer[i], 1);                 // MC: the peer CTA's "my stage is free and armed"
        }
        mbar_fence_init();
    }
    uint32_t crank = 0;
    if (MC) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)(2 * BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (MC) cluster_sync_all();  // the peer's barriers are initialised before anybody arrives on them or multicasts into its shared memory
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == CG_THREADS / 32) {
        // ============================== MMA issuer ==============================
        if (elect_one()) {
            int blk = 0;
            for (int cb = 0; cb < ncb; ++cb) {
                const int buf = cb & 1;
                mbar_wait(&bar_bfull[buf], (cb >> 1) & 1);  // window of this channel block staged (and accumulator cb & 1 folded, see header)
                tc_fence_after();
                const uint32_t sb_hi = smem_u32(smem_b + buf * 2 * CS_B_TILE), sb_lo = sb_hi + CS_B_TILE;
                const uint32_t acc = tmem_d + (uint32_t)(buf * BN);
                for (int tap = 0; tap < 9; ++tap, ++blk) {
                    const int s = blk % CS_ASTAGES;
                    mbar_wait(&bar_afull[s], (blk / CS_ASTAGES) & 1);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(smem_a + s * 2 * CS_A_TILE), sa_lo = sa_hi + CS_A_TILE;
                    const uint32_t shift = (uint32_t)((tap / 3) * W + (tap % 3)) * 16u;  // the tap = a start address
#pragma unroll
                    for (int ks = 0; ks < CG_BK / 8; ++ks) {
                        const uint64_t dah = umma_smem_desc(sa_hi + ks * 2 * CS_A_LBO, CS_A_LBO, CS_A_SBO), dal = umma_smem_desc(sa_lo + ks * 2 * CS_A_LBO, CS_A_LBO, CS_A_SBO);
                        const uint64_t dbh = umma_smem_desc(sb_hi + shift + ks * 2 * CS_B_LBO, CS_B_LBO, CS_B_SBO);
                        const uint64_t dbl = umma_smem_desc(sb_lo + shift + ks * 2 * CS_B_LBO, CS_B_LBO, CS_B_SBO);
                        umma_tf32(acc, dal, dbh, IDESC, (tap | ks) != 0);  // small terms first; a channel block's first MMA overwrites
                        umma_tf32(acc, dah, dbl, IDESC, 1);
                        umma_tf32(acc, dah, dbh, IDESC, 1);
                    }
                    umma_commit(&bar_afree[s]);
                }
                umma_commit(&bar_bdone[buf]);  // every MMA of this channel block has retired: buffer reusable, accumulator complete
            }
        }
        __syncwarp();
    } else if (warp == CG_THREADS / 32 + 1) {
        // ============================== weight loader ==============================
        if (elect_one()) {
            const float *wsrc = a.wpk[prob] + (size_t)blockIdx.y * (9 * ncb) * (2 * CS_A_TILE / 4);
            int blk = 0;
            for (int cb = 0; cb < ncb; ++cb)
                for (int tap = 0; tap < 9; ++tap, ++blk) {
                    const int s = blk % CS_ASTAGES;
                    if (blk >= CS_ASTAGES) mbar_wait(&bar_afree[s], ((blk / CS_ASTAGES) - 1) & 1);
                    mbar_expect_tx(&bar_afull[s], 2 * CS_A_TILE);
                    const float *rec = wsrc + (size_t)(tap * ncb + cb) * (2 * CS_A_TILE / 4);
                    if (!MC) {
                        bulk_g2s(smem_a + s * 2 * CS_A_TILE, rec, 2 * CS_A_TILE, &bar_afull[s]);
                    } else {
                        // Both CTAs of the cluster consume the same record: each fetches ONE half (rank 0 the hi tile, rank 1 the lo
                        // tile) and multicasts it into both shared memories -- half the L2 reads per CTA.  A stage may only be written
                        // once it is free and armed in BOTH CTAs: tell the peer about mine, wait for the peer's.
                        mbar_arrive_remote(&bar_peer[s], crank ^ 1u);
                        mbar_wait_cluster(&bar_peer[s], (blk / CS_ASTAGES) & 1);
                        bulk_g2s_multicast(smem_a + s * 2 * CS_A_TILE + crank * CS_A_TILE, rec + crank * (CS_A_TILE / 4), CS_A_TILE, &bar_afull[s],
                                           (uint16_t)3);
                    }
                }
        }
        __syncwarp();
    } else {
        // ============================== producers: one window pixel per thread, 32 channels ==============================
        const float *xb = a.x[prob] + (size_t)img * a.Cin * HW;
        const int q = q0 + tid;                       // input pixel (linear) this thread stages
        const bool live = tid < win && q < HW;        // beyond the plane: zeros (only garbage columns / the last tile read them)
        const unsigned qoff = live ? (unsigned)q : 0u;
        constexpr int HALF = BN / 2;
        const int row = (warp & 3) * 32 + lane;       // TMEM lane = output channel
        const int col_lo = (warp >> 2) * HALF;
        float racc[HALF];
#pragma unroll
        for (int e = 0; e < HALF; ++e) racc[e] = 0.f;
        auto drain = [&](int buf) {  // fold TMEM accumulator `buf` into racc (the barrier that guards it has been waited on by the caller)
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_d + (uint32_t)(buf * BN) + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col_lo + c0), v);
#pragma unroll
                for (int e = 0; e < 16; ++e) racc[c0 + e] += v[e];
            }
            tc_fence_before();
        };
        float v[CG_BK];
        auto load = [&](int cb) {
            const float *ublk = xb + (size_t)cb * CG_BK * HW;  // uniform over the CTA
            asm volatile("" : "+l"(ublk));
            unsigned o = qoff;
#pragma unroll
            for (int k = 0; k < CG_BK; ++k) {
                v[k] = live ? __ldg(ublk + o) : 0.f;
                o += (unsigned)HW;
            }
        };
        load(0);
        for (int cb = 0; cb < ncb; ++cb) {
            const int buf = cb & 1;
            if (cb >= 2) {
                mbar_wait(&bar_bdone[buf], ((cb >> 1) - 1) & 1);  // MMAs of channel block cb - 2 retired: its buffer is free ...
                drain(buf);                                        // ... and its accumulator complete: fold it
            }
            float *b_hi = reinterpret_cast<float *>(smem_b + buf * 2 * CS_B_TILE), *b_lo = b_hi + CS_B_TILE / 4;
            if (tid < CS_WIN) {
#pragma unroll
                for (int kc = 0; kc < CG_BK / 4; ++kc) {
                    float4 h, l;
                    split_tf32(v[4 * kc + 0], h.x, l.x); split_tf32(v[4 * kc + 1], h.y, l.y);
                    split_tf32(v[4 * kc + 2], h.z, l.z); split_tf32(v[4 * kc + 3], h.w, l.w);
                    const int off = kc * (CS_WIN * 4) + tid * 4;  // floats: k-group kc, pixel row tid
                    *reinterpret_cast<float4 *>(b_hi + off) = h;
                    *reinterpret_cast<float4 *>(b_lo + off) = l;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_bfull[buf]);
            if (cb + 1 < ncb) load(cb + 1);  // in flight while the tensor core works through this block's 108 MMAs
        }
        // the last two channel blocks' accumulators
        for (int cb = ncb > 1 ? ncb - 2 : 0; cb < ncb; ++cb) {
            mbar_wait(&bar_bdone[cb & 1], (cb >> 1) & 1);
            drain(cb & 1);
        }
        // ---- epilogue: transpose through the (dead) staging buffers, then pixel-contiguous rows with BatchNorm (+ residual) (+ ReLU) ----
        constexpr int YP = BN + 1;
        float *ys = reinterpret_cast<float *>(smem);
        const int co = co0 + row;
        const float *scp = a.scale[prob], *shp = a.shift[prob], *resp = a.residual[prob];
        const float sc = scp ? __ldg(scp + co) : 1.f, sh = shp ? __ldg(shp + co) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(CG_THREADS) : "memory");  // see conv_gemm.cu: states the staging-store -> epilogue ordering for racecheck
#pragma unroll
        for (int e = 0; e < HALF; ++e) ys[row * YP + col_lo + e] = fmaf(racc[e], sc, sh);
        asm volatile("bar.sync 1, %0;" ::"n"(CG_THREADS) : "memory");
        const int HWo = Ho * Wo;
        const size_t tile_base = ((size_t)img * a.Cout + co0) * HWo;
        float *outp = a.out[prob];
        const int c = tid % BN, p = q0 + c;          // this thread's column of the tile -> (output row, column) on the input-width grid
        const int orow = p / W, ocol = p - orow * W;
        if (orow < Ho && ocol < Wo) {
            const int po = orow * Wo + ocol;
            for (int r = tid / BN; r < CG_BM; r += CG_THREADS / BN) {
                float y = ys[r * YP + c];
                if (resp) y += __ldg(resp + tile_base + (size_t)r * HWo + po);
                if (a.relu) y = fmaxf(y, 0.f);
                outp[tile_base + (size_t)r * HWo + po] = y;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MC) cluster_sync_all();  // nobody leaves while the peer could still address this CTA's shared memory
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)(2 * BN)));
}

bool conv_shift_applicable(const ConvGemmArgs &a, int ksize, int valid) {
    return ksize == 3 && valid && a.stride == 1 && a.dil == 1 && 2 * a.W + 2 <= 256 - CS_BN && a.H >= 3 && a.W >= 3 && a.Cin % CG_BK == 0 &&
           a.Cout % CG_BM == 0;
}

int g_conv_shift_multicast = 0;  // hdn_conv_gemm_set_shift(2): weight records multicast across a 2-CTA cluster -- built, correct, and measured
                                 // SLOWER than every CTA loading its own (16.1 vs 14.2 ms for the fused chain): the per-block handshake couples
                                 // the two CTAs' pipelines and L2 bandwidth was not the limit.  Kept as an A/B switch.

template <int WIN, int ASTAGES>
static int launch_conv_shift_cfg(const ConvGemmArgs &a, int nprob, cudaStream_t st) {
    constexpr size_t SMEM = CSCfg<WIN, ASTAGES>::SMEM;
    static DeviceOnce once;
    if (int e = once.run([] {
            cudaError_t e1 = cudaFuncSetAttribute(conv3x3_shift_kernel<false, WIN, ASTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
            return e1 != cudaSuccess ? e1 : cudaFuncSetAttribute(conv3x3_shift_kernel<true, WIN, ASTAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        }))
        return e;
    const int nq = (a.Ho - 1) * a.W + a.Wo;  // outputs on the input-width grid (the last row stops at its last valid column)
    const int tiles = (nq + CS_BN - 1) / CS_BN;
    if (!g_conv_shift_multicast || tiles < 2) {
        conv3x3_shift_kernel<false, WIN, ASTAGES><<<dim3(tiles, a.Cout / CG_BM, a.B * nprob), CG_THREADS + 64, SMEM, st>>>(a);
    } else {  // clusters of two neighbouring pixel tiles (an odd tile count gets one idle partner that stores nothing)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((tiles + 1) / 2 * 2, a.Cout / CG_BM, a.B * nprob);
        cfg.blockDim = dim3(CG_THREADS + 64);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_shift_kernel<true, WIN, ASTAGES>, a);
        if (e != cudaSuccess) return (int)e;
    }
    count_launch();
    return launch_status();
}

int launch_conv_shift(const ConvGemmArgs &a, int nprob, cudaStream_t st) {
    return 2 * a.W + 2 <= 192 - CS_BN ? launch_conv_shift_cfg<192, 4>(a, nprob, st) : launch_conv_shift_cfg<256, 3>(a, nprob, st);
}

}  // namespace hdn
