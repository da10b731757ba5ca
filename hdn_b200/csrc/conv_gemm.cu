// conv_gemm.cu -- stride-1 1x1 / (dilated) 3x3 convolution as an implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma, accumulator in TMEM), fp32-accurate through 3xTF32 operand splitting, with the eval-mode BatchNorm,
// the residual add and the ReLU of a ResNet block fused into the epilogue.
//
// Replaces, for the backbone's stride-1 layers (hdn/models/backbone/resnet_atrous.py:62-110, neck.py:11-29), the
// cuDNN fp32 convolutions of the reference.  With TF32 off cuDNN runs these at 2-7 TFLOP/s for tracking batch sizes
// (and falls back to a direct kernel for the dilated layers); plain TF32 would be fast but its 10-bit mantissa can
// flip the arg-max that must stay bit-exact.  3xTF32 keeps fp32 accuracy on the tensor cores:
//     x = hi + lo,  hi = x with the low 13 mantissa bits cleared (exactly a TF32 number),  lo = x - hi  (exact in fp32)
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (dropped term a_lo*b_lo <= 2^-20 |a b|)
// accumulated in fp32 in TMEM.  The tensor core adds into its accumulator with truncation, whose bias grows with the
// length of the sum (2.7e-5 of max|out| at K = 4608, probed), so the accumulation is CHUNKED: every 256 of K the MMAs
// switch to the other of two TMEM accumulators and the threads fold the finished one into fp32 registers with
// round-to-nearest adds (Ootomo & Yokota's remedy) while the tensor core keeps running.
//
// GEMM view per image:  D[co, p] = sum_{tap, ci} Wt[co, tap, ci] * X[ci, p + shift(tap)]      (zero outside the image)
//     A = Wt   [Cout x K]   K = taps*Cin contiguous  -> UMMA K-major
//     B = X    [K x HW]     pixels contiguous in HBM -> transposed to UMMA K-major ([pixel][k]) while staging (MN-major
//                                                        TF32 without the 32B-base swizzle reads as zeros; probed)
// One CTA owns a 128 (co) x BN (pixels) tile and walks K in 32-deep blocks through a ring of stages:
//   * WEIGHTS are constants, so their hi / lo split and the UMMA tiling are done once (hdn_conv_pack_weight_f32): a block's
//     A operand is one contiguous 32 KB record that a single thread fetches with a 1-D TMA bulk copy (UBLKCP);
//   * ACTIVATIONS are fetched 4 blocks ahead with cp.async into a landing ring (4-byte copies: NCHW rows of odd length are
//     only 4-byte aligned, zero-fill = padding), then 256 producer threads split them (hi / lo) and store them
//     transposed into the canonical no-swizzle K-major UMMA layout (8 x 16-byte core matrices);
//   * a dedicated warp's elected lane issues 4 k-steps x 3 tcgen05.mma (kind::tf32, M=128, N=BN, K=8) per block and commits
//     them to the mbarrier that frees the stage.
// Epilogue: tcgen05.ld (32 lanes x 32b x 16 columns) -> y = acc*scale[co] + shift[co] (+ residual) (ReLU) -> global NCHW.
#include <cstdlib>

#include "umma.cuh"

namespace hdn {


// Warp-specialised: warps 0..7 (256 threads) are PRODUCERS (global -> registers -> hi/lo -> shared, TMEM drains, epilogue),
// warp 8 is the MMA ISSUER.  Stages hand over through mbarriers (full: 256 producer arrivals; free: tcgen05.commit), so
// staging of block k+1.. never waits for the issue of block k; two TMEM accumulators alternate per 256-deep K chunk.
// PROJECT: instead of storing its 128 x BN tile y = ReLU(BN(conv)), the CTA multiplies it by the matching 128-column slice of a
// second 1x1 convolution w2 [L, Cout] (the `head[3]` layer of DepthwiseXCorr, ban.py:62-66) and stores the L x BN PARTIAL sums
// (one per 128-channel tile; fixed summation order -> deterministic) -- the 256-channel hidden map never goes to HBM.
template <int BN, int STAGES, int RAW, bool PROJECT>
__global__ void __launch_bounds__(CG_THREADS + 64, 1) conv_gemm_tf32x3_kernel(const __grid_constant__ ConvGemmArgs a) {
    constexpr int A_TILE = CG_BM * CG_BK * 4;  // bytes of one operand tile (hi or lo)
    constexpr int B_TILE = CG_BK * BN * 4;
    constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    constexpr uint32_t A_SBO = 128, A_LBO = (CG_BM / 8) * 128;  // K-major: 8-row groups 128 B apart, 4-wide K chunks A_LBO apart
    constexpr uint32_t B_SBO = 128, B_LBO = (BN / 8) * 128;     // K-major as well (pixel rows): the tile is transposed while staging
    // kind::tf32, fp32 accumulate, A and B K-major, N = BN, M = 128 (cute::UMMA::InstrDescriptor bit layout)
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(CG_BM >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar_full[STAGES], bar_free[STAGES], bar_acc_full[2], bar_acc_free[2];
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int HW = a.H * a.W, HWo = a.Ho * a.Wo;
    // PROGRAMMATIC DEPENDENT LAUNCH (launch_conv_gemm sets the stream-serialisation attribute): the next convolution of the stream may
    // start as soon as every CTA of this one is running -- on the SMs this grid leaves free at tracking batch sizes -- and does
    // everything that does not depend on this kernel's output (barrier init, TMEM allocation, the first weight records, which are
    // constants) before its producers block in griddepcontrol.wait.  All reads of activations / residuals and all global writes of a
    // kernel come after its own wait, so the chain N-1 -> N -> N+1 stays ordered.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // SPLIT-K over a thread-block cluster (small problems, tracking batch sizes): the `splitk` CTAs of a cluster own the same output
    // tile and 1/splitk of the K blocks each; their fp32 partial tiles meet in the leader through distributed shared memory.
    const int S = PROJECT ? 1 : a.splitk;
    uint32_t crank = 0;
    if (S > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int tile_x = S > 1 ? (int)blockIdx.x / S : (int)blockIdx.x;
    const int pix0 = tile_x * BN, co0 = blockIdx.y * CG_BM, prob = blockIdx.z / a.B, img = blockIdx.z - prob * a.B;
    const int Ktot = a.taps * a.Cin;
    const int nkb = Ktot / CG_BK / S;     // K blocks of THIS CTA ...
    const int kb0 = (int)crank * nkb;     // ... starting at global block kb0
    const int nchunks = (nkb + CG_KCB - 1) / CG_KCB;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&bar_full[i], CG_THREADS / 32 + 1);  // one arrival per producer warp + the TMA issuer's arrive.expect_tx
            mbar_init(&bar_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_acc_full[i], 1);
            mbar_init(&bar_acc_free[i], CG_THREADS / 32);  // one arrival per producer warp
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)(2 * BN)));
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
                const int s = kb % STAGES, chunk = kb / CG_KCB;
                if (kb % CG_KCB == 0 && chunk >= 2) mbar_wait(&bar_acc_free[chunk & 1], ((chunk >> 1) - 1) & 1);  // producers drained this accumulator
                mbar_wait(&bar_full[s], (kb / STAGES) & 1);
                tc_fence_after();
                const uint32_t sa_hi = smem_u32(smem + s * STAGE), sa_lo = sa_hi + A_TILE, sb_hi = sa_hi + 2 * A_TILE, sb_lo = sb_hi + B_TILE;
                const uint32_t acc = tmem_d + (uint32_t)((chunk & 1) * BN);
#pragma unroll
                for (int ks = 0; ks < CG_BK / 8; ++ks) {
                    const uint64_t dah = umma_smem_desc(sa_hi + ks * 2 * A_LBO, A_LBO, A_SBO), dal = umma_smem_desc(sa_lo + ks * 2 * A_LBO, A_LBO, A_SBO);
                    const uint64_t dbh = umma_smem_desc(sb_hi + ks * 2 * B_LBO, B_LBO, B_SBO), dbl = umma_smem_desc(sb_lo + ks * 2 * B_LBO, B_LBO, B_SBO);
#ifndef HDN_EXP_NOMMA
                    umma_tf32(acc, dal, dbh, IDESC, ((kb % CG_KCB) | ks) != 0);  // small terms first; a chunk's first MMA overwrites
#ifndef HDN_EXP_ONEMMA
                    umma_tf32(acc, dah, dbl, IDESC, 1);
                    umma_tf32(acc, dah, dbh, IDESC, 1);
#endif
#endif
                }
                umma_commit(&bar_free[s]);  // arrives when the MMAs above have finished reading the stage
                if (kb % CG_KCB == CG_KCB - 1 || kb == nkb - 1) umma_commit(&bar_acc_full[chunk & 1]);  // ... and the chunk is complete
            }
        }
        __syncwarp();
    } else if (warp == CG_THREADS / 32 + 1) {
        // ============================== weight loader: one 32 KB TMA bulk copy per K block ==============================
        if (elect_one()) {
            const float *wsrc = a.wpk[prob] + (size_t)blockIdx.y * (nkb * S) * (2 * A_TILE / 4);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&bar_free[s], ((kb / STAGES) - 1) & 1);
                mbar_expect_tx(&bar_full[s], 2 * A_TILE);
                bulk_g2s(smem + s * STAGE, wsrc + (size_t)(kb0 + kb) * (2 * A_TILE / 4), 2 * A_TILE, &bar_full[s]);
            }
        }
        __syncwarp();
    } else {
        // ============================== producers (activations) ==============================
        const float *xb = a.x[prob] + (size_t)img * a.Cin * HW;
        // A: float4 slots f = tid + 256*j, j < 4:      r0 = f&7, kc = (f>>3)&7, rg = f>>6        (row = rg*8 + r0, k = kc*4..+3)
        // B: slots f = tid + 256*j, j < BN/32:          n = f % BN (= tid % BN for every j), kc = f / BN   (k = kc*4..+3)
        //    consecutive lanes -> consecutive pixels: coalesced scalar loads, and one conflict-free 16-byte store per slot.
        constexpr int NBJ = BN / 32;
        const int bn = tid % BN;
        const int bp = pix0 + bn;
        const int b_r = bp < HWo ? (bp / a.Wo) * a.stride + a.off : -(1 << 20);  // beyond the tile: never valid
        const int b_c = bp < HWo ? (bp - (bp / a.Wo) * a.Wo) * a.stride + a.off : 0;

        // fp32 register accumulator of this thread's share of the tile: TMEM lane (= channel) row, columns [col_lo, col_lo + BN/2)
        constexpr int HALF = BN / 2;
        const int row = (warp & 3) * 32 + lane;  // a warp may only touch its own TMEM lane quadrant
        const int col_lo = (warp >> 2) * HALF;
        float racc[HALF];
#pragma unroll
        for (int e = 0; e < HALF; ++e) racc[e] = 0.f;
        auto drain = [&](int chunk) {  // fold finished TMEM accumulator `chunk & 1` into racc (round-to-nearest adds), then hand it back
            mbar_wait(&bar_acc_full[chunk & 1], (chunk >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_d + (uint32_t)((chunk & 1) * BN) + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(col_lo + c0), v);
#pragma unroll
                for (int e = 0; e < 16; ++e) racc[c0 + e] += v[e];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_free[chunk & 1]);
        };

        // Activations: block kb's [32 k][BN pixels] tile is read straight into registers with coalesced 4-byte loads (NCHW rows
        // of odd length are only 4-byte aligned; consecutive lanes = consecutive pixels -> one 128-byte line per warp load,
        // predicated off = convolution padding / tile tail), ONE block ahead of the block being converted, so the loads of
        // block kb+1 are in flight while block kb is split (hi / lo) and stored into the canonical no-swizzle K-major UMMA
        // layout.  (The first version staged them through a cp.async landing ring: 16 4-byte LDGSTS per thread and block kept
        // the load/store unit busier than the tensor core -- profiles/r01_ncu_conv_gemm.md.)
        // Everything that does not change from block to block is hoisted: the producers' issue slots, not the tensor core, bounded
        // the first versions (3,800 warp instructions per K block and CTA, 40 % of them 64-bit address arithmetic and integer
        // divisions, 22 % barrier polling).  A block's loads are  ublk[toff[j][e]]:  ublk = warp-uniform base of (channel block,
        // tap shift), advanced incrementally (no division), toff = 16 per-thread element offsets fixed for the whole kernel.
        unsigned toff[NBJ][4];
        {
            const unsigned pix_off = bp < HWo ? (unsigned)(b_r * a.W + b_c) : 0u;  // off >= 0 (pad <= (ksize/2)*dil): never negative
#pragma unroll
            for (int j = 0; j < NBJ; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) toff[j][e] = (unsigned)(((tid + CG_THREADS * j) / BN) * 4 + e) * (unsigned)HW + pix_off;
        }
        int ld_tap, ld_ci0, ld_ty, ld_tx;  // position of the NEXT block to load (blocks are loaded in order)
        {
            const int k0 = kb0 * CG_BK;
            ld_tap = k0 / a.Cin;
            ld_ci0 = k0 - ld_tap * a.Cin;
            ld_ty = ld_tap / 3;
            ld_tx = ld_tap - 3 * ld_ty;
        }
        auto load_block = [&](int, float (&v)[NBJ][4]) {
            const int dy = a.taps == 1 ? 0 : (ld_ty - 1) * a.dil, dx = a.taps == 1 ? 0 : (ld_tx - 1) * a.dil;
            const bool ok = (unsigned)(b_r + dy) < (unsigned)a.H && (unsigned)(b_c + dx) < (unsigned)a.W;
            const float *ublk = xb + ((long long)ld_ci0 * HW + dy * a.W + dx);  // the same for every thread of the CTA
            asm volatile("" : "+l"(ublk));  // keep it ONE pointer: each load is then base + toff * 4 (a single IMAD.WIDE), not a re-associated 64-bit sum
#pragma unroll
            for (int j = 0; j < NBJ; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[j][e] = ok ? __ldg(ublk + toff[j][e]) : 0.f;
            ld_ci0 += CG_BK;
            if (ld_ci0 == a.Cin) {
                ld_ci0 = 0;
                if (++ld_tx == 3) { ld_tx = 0; ++ld_ty; }
            }
        };
        unsigned soff[NBJ];  // this thread's 16-byte slots inside a stage's B tile (canonical K-major layout), in floats
#pragma unroll
        for (int j = 0; j < NBJ; ++j) soff[j] = (unsigned)((((tid + CG_THREADS * j) / BN) * B_LBO + (bn >> 3) * B_SBO + (bn & 7) * 16) >> 2);
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the producing kernel(s) have completed and flushed: activations may be read
        int drained = 0;
        auto stage_block = [&](int kb, const float (&v)[NBJ][4]) {
            const int s = kb % STAGES;
            // ---- the MMAs that read this stage STAGES blocks ago must have retired ----
            if (kb >= STAGES) mbar_wait(&bar_free[s], ((kb / STAGES) - 1) & 1);
            // ---- fold a finished accumulation chunk into the fp32 registers.  Only once block kb - STAGES -- the chunk's last or a
            //      later one -- is known to have retired (the wait above), so this never stalls on the tensor core: draining two
            //      blocks after a chunk was STAGED made all producers wait ~2 blocks of MMA time per chunk and then starved the
            //      tensor core in turn (tensor pipe 40 %, profiles/r02_ncu_conv_search.md) ----
            if (drained < nchunks && kb >= (drained + 1) * CG_KCB - 1 + STAGES) drain(drained++);
            float *b_hi = reinterpret_cast<float *>(smem + s * STAGE + 2 * A_TILE), *b_lo = b_hi + B_TILE / 4;
#pragma unroll
            for (int j = 0; j < NBJ; ++j) {
                float4 h, l;
                split_tf32(v[j][0], h.x, l.x); split_tf32(v[j][1], h.y, l.y); split_tf32(v[j][2], h.z, l.z); split_tf32(v[j][3], h.w, l.w);
                *reinterpret_cast<float4 *>(b_hi + soff[j]) = h;
                *reinterpret_cast<float4 *>(b_lo + soff[j]) = l;
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[s]);  // one arrival per producer warp (its 32 threads' stores are fenced and ordered by the warp barrier)
        };
#ifndef HDN_CG_PREFETCH
#define HDN_CG_PREFETCH 2  // K blocks of activation loads in flight ahead of the block being converted (build-time A/B switch)
#endif
#if HDN_CG_PREFETCH == 1
        float va[NBJ][4], vb[NBJ][4];
        load_block(0, va);
        for (int kb = 0; kb < nkb; kb += 2) {
            if (kb + 1 < nkb) load_block(kb + 1, vb);
            stage_block(kb, va);
            if (kb + 1 < nkb) {
                if (kb + 2 < nkb) load_block(kb + 2, va);
                stage_block(kb + 1, vb);
            }
        }
#else
        float va[NBJ][4], vb[NBJ][4], vc[NBJ][4];  // three register sets rotate: two blocks of loads are in flight behind the one in hand
        load_block(0, va);
        if (nkb > 1) load_block(1, vb);
        for (int kb = 0; kb < nkb; kb += 3) {
            if (kb + 2 < nkb) load_block(kb + 2, vc);
            stage_block(kb, va);
            if (kb + 1 < nkb) {
                if (kb + 3 < nkb) load_block(kb + 3, va);
                stage_block(kb + 1, vb);
            }
            if (kb + 2 < nkb) {
                if (kb + 4 < nkb) load_block(kb + 4, vb);
                stage_block(kb + 2, vc);
            }
        }
#endif

        // ---- epilogue: remaining chunks -> registers -> BN / residual / ReLU -> global NCHW ----
        while (drained < nchunks) drain(drained++);
        const int co = co0 + row;
        const bool co_ok = co < a.Cout;  // Cout = 64 layers run in a zero-padded 128-row tile
        const float *scp = a.scale[prob], *shp = a.shift[prob], *resp = a.residual[prob];
        const float sc = scp && co_ok ? __ldg(scp + co) : 1.f, sh = shp && co_ok ? __ldg(shp + co) : 0.f;
        const size_t obase = ((size_t)img * a.Cout + co) * HWo;
        if (!PROJECT) {
            // A thread owns one CHANNEL row of the tile, so direct stores would touch 32 different lines per warp instruction.
            // The pipeline stages are dead by now: transpose through them (pitch BN + 1) and write pixel-contiguous rows,
            // residual add and ReLU applied on the way out (coalesced residual reads as well).
            constexpr int YP = BN + 1;
            float *ys = reinterpret_cast<float *>(smem);
            // Every producer's last staging store is already ordered before this point through bar_full -> MMA -> commit -> the drain's
            // wait; the named barrier states that ordering directly (compute-sanitizer's racecheck does not follow the tensor-core hop).
            asm volatile("bar.sync 1, %0;" ::"n"(CG_THREADS) : "memory");
            if (S > 1) {  // split-K: leave the raw fp32 partial tile in shared memory; the cluster reduces it below
#pragma unroll
                for (int e = 0; e < HALF; ++e) ys[row * YP + col_lo + e] = racc[e];
            } else {
#pragma unroll
            for (int e = 0; e < HALF; ++e) ys[row * YP + col_lo + e] = fmaf(racc[e], sc, sh);
            asm volatile("bar.sync 1, %0;" ::"n"(CG_THREADS) : "memory");  // the 256 producer threads only
            const size_t tile_base = ((size_t)img * a.Cout + co0) * HWo;
            float *outp = a.out[prob];
#pragma unroll 4
            for (int i = tid; i < CG_BM * BN; i += CG_THREADS) {
                const int r = i / BN, c = i - r * BN, p = pix0 + c;
                if (p < HWo && co0 + r < a.Cout) {
                    float y = ys[r * YP + c];
                    if (resp) y += __ldg(resp + tile_base + (size_t)r * HWo + p);
                    if (a.relu) y = fmaxf(y, 0.f);
                    outp[tile_base + (size_t)r * HWo + p] = y;
                }
            }
            }
        } else {
            // The pipeline stages are dead (the last accumulator was committed after every MMA had read them): stage the tile as
            // ys[channel][pixel] (pitch BN + 1: lanes = consecutive channels hit consecutive banks) next to the w2 slice.
            constexpr int YP = BN + 1;
            float *ys = reinterpret_cast<float *>(smem);
            float *w2s = ys + CG_BM * YP;  // [L][128]
            asm volatile("bar.sync 1, %0;" ::"n"(CG_THREADS) : "memory");  // (ordering already holds; stated for racecheck, see above)
#pragma unroll
            for (int e = 0; e < HALF; ++e) {
                float y = fmaf(racc[e], sc, sh);
                if (a.relu) y = fmaxf(y, 0.f);
                ys[row * YP + col_lo + e] = y;
            }
            for (int i = tid; i < a.L * CG_BM; i += CG_THREADS) w2s[i] = __ldg(a.w2[prob] + (size_t)(i / CG_BM) * a.Cout + co0 + (i % CG_BM));
            asm volatile("bar.sync 1, %0;" ::"n"(CG_THREADS) : "memory");  // the 256 producer threads only
            const int c = tid % BN, p = pix0 + c;
            for (int l = tid / BN; l < a.L; l += CG_THREADS / BN) {
                const float *wr = w2s + l * CG_BM;
                float s0 = 0.f, s1 = 0.f;  // two chains: channels in fixed order r = 0, 2, 4, ... and 1, 3, 5, ...
#pragma unroll 8
                for (int r = 0; r < CG_BM; r += 2) {
                    s0 = fmaf(wr[r], ys[r * YP + c], s0);
                    s1 = fmaf(wr[r + 1], ys[(r + 1) * YP + c], s1);
                }
                // partial sums of channel tile blockIdx.y:  out[tile][img][l][pixel]
                if (p < HWo) a.out[prob][(((size_t)blockIdx.y * a.B + img) * a.L + l) * HWo + p] = s0 + s1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (!PROJECT && S > 1) {
        // ---- split-K reduction through distributed shared memory: every CTA of the cluster holds its partial tile in ys; the leader
        //      adds them in rank order (deterministic), applies BatchNorm / residual / ReLU and stores pixel-contiguous rows ----
        constexpr int YP = BN + 1;
        const float *ys = reinterpret_cast<const float *>(smem);
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (crank == 0) {
            const float *scp = a.scale[prob], *shp = a.shift[prob], *resp = a.residual[prob];
            const size_t tile_base = ((size_t)img * a.Cout + co0) * HWo;
            float *outp = a.out[prob];
            const uint32_t ys_s = smem_u32(ys);
            for (int i = tid; i < CG_BM * BN; i += CG_THREADS + 64) {
                const int r = i / BN, c = i - r * BN, p = pix0 + c;
                if (p < HWo && co0 + r < a.Cout) {
                    float acc = ys[r * YP + c];
                    for (int peer = 1; peer < S; ++peer) {
                        uint32_t remote;
                        float v;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(ys_s + (uint32_t)(r * YP + c) * 4u), "r"(peer));
                        asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
                        acc += v;
                    }
                    const int co = co0 + r;
                    float y = fmaf(acc, scp ? __ldg(scp + co) : 1.f, shp ? __ldg(shp + co) : 0.f);
                    if (resp) y += __ldg(resp + tile_base + (size_t)r * HWo + p);
                    if (a.relu) y = fmaxf(y, 0.f);
                    outp[tile_base + (size_t)r * HWo + p] = y;
                }
            }
        }
        // nobody leaves (and frees its shared memory) before the leader has read every partial tile
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)(2 * BN)));
}

int g_conv_pdl = 1;  // hdn_conv_gemm_set_pdl: programmatic dependent launch of consecutive convolutions (default on)

template <int BN, int STAGES, int RAW, bool PROJECT = false>
static int launch_conv_gemm(const ConvGemmArgs &a, int nprob, cudaStream_t st) {
    constexpr size_t SMEM = (size_t)STAGES * (2 * CG_BM * CG_BK * 4 + 2 * CG_BK * BN * 4) + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static_assert((size_t)(CG_BM * (BN + 1) + 8 * CG_BM) * 4 <= SMEM, "epilogue staging must fit the pipeline's shared memory");
    static DeviceOnce once;
    if (int e = once.run([] { return cudaFuncSetAttribute(conv_gemm_tf32x3_kernel<BN, STAGES, RAW, PROJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM); }))
        return e;
    const int S = a.splitk > 1 ? a.splitk : 1;
    dim3 grid((a.Ho * a.Wo + BN - 1) / BN * S, (a.Cout + CG_BM - 1) / CG_BM, a.B * nprob);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(CG_THREADS + 64);
    cfg.dynamicSmemBytes = SMEM;
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
    if (g_conv_pdl) {  // see the kernel: prologue + first weight records overlap the tail of the previous kernel of the stream
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_gemm_tf32x3_kernel<BN, STAGES, RAW, PROJECT>, a);
    if (e != cudaSuccess) return (int)e;
    count_launch();
    return launch_status();
}

// Weights -> per (128-channel tile, 32-deep K block) records of [hi | lo] x [k/4 (8)][row/8 (16)][row%8 (8)][k%4 (4)] floats:
// exactly the bytes a stage's A region holds, so the kernel moves a block's A operand with one bulk copy.
__global__ void conv_pack_weight_kernel(const float *__restrict__ wt, float *__restrict__ out, int Cout, int Ktot) {
    const int rows = (Cout + CG_BM - 1) / CG_BM * CG_BM;  // a 64-wide layer is packed into a zero-padded 128-row tile
    const long long total = (long long)rows * Ktot;
    const int nkb = Ktot / CG_BK;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i / Ktot), k = (int)(i - (long long)co * Ktot);
        const int mt = co / CG_BM, row = co % CG_BM, kb = k / CG_BK, kk = k % CG_BK;
        const long long rec = ((long long)mt * nkb + kb) * (2 * CG_BM * CG_BK);
        const int off = (kk >> 2) * (CG_BM * 4) + (row >> 3) * 32 + (row & 7) * 4 + (kk & 3);
        float hi, lo;
        split_tf32(co < Cout ? wt[i] : 0.f, hi, lo);
        out[rec + off] = hi;
        out[rec + CG_BM * CG_BK + off] = lo;
    }
}

__global__ void conv_pack_fence_kernel() {}

}  // namespace hdn

using namespace hdn;

extern "C" int hdn_conv_pack_weight_f32(const float *wt, float *packed, int Cout, int Ktot, hdn_stream_t stream) {
    if (!wt || !packed) return HDN_ERR_NULL;
    if (Cout < 64 || Cout % 64 || Ktot < CG_BK || Ktot % CG_BK) return HDN_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(packed) & 15u) return HDN_ERR_ALIGN;
    conv_pack_weight_kernel<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(wt, packed, Cout, Ktot);
    // A convolution launched right behind this one reads its first weight records BEFORE its griddepcontrol.wait (they are constants
    // to it).  The empty kernel makes the pack kernel's completion -- full stream-order completion, memory flushed -- the thing
    // that convolution's launch depends on, instead of a programmatic edge to the pack kernel itself.
    conv_pack_fence_kernel<<<1, 32, 0, (cudaStream_t)stream>>>();
    count_launch(2);
    return launch_status();
}

extern "C" int hdn_conv_gemm_supported(int Cin, int Cout, int ksize, int dilation) {
    return (Cin >= 32 && Cin % 32 == 0 && Cout >= 64 && Cout % 64 == 0 && (ksize == 1 || ksize == 3) && dilation >= 1) ? 1 : 0;
}

static int g_conv_splitk = 1;  // hdn_conv_gemm_set_splitk (A/B switch; the result is deterministic either way)
static int g_conv_shift = 1;   // hdn_conv_gemm_set_shift: conv_shift.cu for 3x3 'valid' layers (A/B switch)
static int g_conv_ts = 2;      // hdn_conv_gemm_set_ts: conv_gemm_ts.cu (activations in tensor memory); 0 = off, 1 = large launches only, 2 = every launch (default)

static int conv_gemm_multi(int n, const float *const *x, const float *const *wpk, const float *const *scale, const float *const *shift,
                           const float *const *residual, const float *const *w2, float *const *out, int B, int Cin, int Cout, int H, int W,
                           int ksize, int dilation, int valid, int relu, int L, cudaStream_t st, int stride = 1, int pad = -1) {
    if (!x || !wpk || !out) return HDN_ERR_NULL;
    if (n < 1 || n > HDN_MAX_PROBLEMS) return HDN_ERR_UNSUPPORTED;
    if (B < 1 || H < 1 || W < 1 || (long long)B * n > 65535) return HDN_ERR_SHAPE;
    if (!hdn_conv_gemm_supported(Cin, Cout, ksize, dilation)) return HDN_ERR_UNSUPPORTED;
    if (pad < 0) pad = valid ? 0 : dilation * (ksize / 2);  // the two geometries of hdn_conv_gemm_f32
    if ((stride != 1 && stride != 2) || pad > dilation * (ksize / 2)) return HDN_ERR_UNSUPPORTED;
    valid = (pad == 0 && stride == 1) ? 1 : 0;
    const int Ho_ = (H + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1, Wo_ = (W + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1;
    if (Ho_ < 1 || Wo_ < 1) return HDN_ERR_SHAPE;
    ConvGemmArgs a{};
    for (int i = 0; i < n; ++i) {
        if (!x[i] || !wpk[i] || !out[i] || (w2 && !w2[i])) return HDN_ERR_NULL;
        if (reinterpret_cast<uintptr_t>(wpk[i]) & 15u) return HDN_ERR_ALIGN;
        a.x[i] = x[i];
        a.wpk[i] = wpk[i];
        a.scale[i] = scale ? scale[i] : nullptr;
        a.shift[i] = shift ? shift[i] : nullptr;
        a.residual[i] = residual ? residual[i] : nullptr;
        a.w2[i] = w2 ? w2[i] : nullptr;
        a.out[i] = out[i];
    }
    a.B = B; a.L = L; a.splitk = 1;
    a.Cin = Cin; a.Cout = Cout; a.H = H; a.W = W; a.taps = ksize * ksize; a.dil = dilation; a.relu = relu;
    a.Ho = Ho_; a.Wo = Wo_; a.off = (ksize / 2) * dilation - pad; a.stride = stride;
    const int mtiles = (Cout + CG_BM - 1) / CG_BM;
    const long long tiles128 = (long long)((a.Ho * a.Wo + 127) / 128) * mtiles * B * n;
    // large launches (>= 2 CTAs per SM of 128 x 128 tiles): activation operand in tensor memory (conv_gemm_ts.cu).  It also beats the
    // shifted-window kernel below on the heads' batched conv_search launches (fused chain 14.05 -> 13.29 ms per 64 pairs).
    static const int ts_min_pct = getenv("HDN_B200_TS_MIN_TILES_PCT") ? atoi(getenv("HDN_B200_TS_MIN_TILES_PCT")) : 60;  // % of the SM count (sweep: DESIGN.md K7d)
    if (!w2 && g_conv_ts && tiles128 * 100 >= (long long)ts_min_pct * sm_count()) return launch_conv_gemm_ts(a, n, st);
    // 3x3 'valid' layers (the heads' conv_search / conv_kernel) at tracking batch sizes: activations staged once per channel block,
    // taps = shifted windows
    if (!w2 && g_conv_shift && conv_shift_applicable(a, ksize, valid)) return launch_conv_shift(a, n, st);
    if (w2) {  // fused second 1x1: narrow pixel tiles (the projection's staging pitch), L <= 8
        if (L < 1 || L > 8) return HDN_ERR_UNSUPPORTED;
        return launch_conv_gemm<64, 4, 0, true>(a, n, st);
    }
    const int nkb = a.taps * Cin / CG_BK;
    if (!w2 && g_conv_ts == 2) {  // (A/B) the tensor-memory-operand kernel for the small launches as well, K split over a cluster
        int split = 1;
        if (g_conv_splitk != 0)
            for (int s2 = 8; s2 >= 2; s2 /= 2)
                if (nkb % s2 == 0 && nkb / s2 >= 4 && tiles128 * s2 <= sm_count() + sm_count() / 4) { split = s2; break; }
        a.splitk = split;
        return launch_conv_gemm_ts(a, n, st);
    }
    // small problems (tracking batch sizes): narrower pixel tiles put more CTAs on the 148 SMs ...
    if (tiles128 >= 2 * sm_count()) return launch_conv_gemm<128, 3, 0>(a, n, st);
    // ... and when even those leave most SMs idle (a 15x15 or 31x31 map at batch 1), K is split over a cluster of 2 / 4 / 8 CTAs
    const long long ctas = (long long)((a.Ho * a.Wo + 63) / 64) * mtiles * B * n;
    int split = 1;
    if (g_conv_splitk != 0)
        for (int s2 = 8; s2 >= 2; s2 /= 2)
            if (nkb % s2 == 0 && nkb / s2 >= 4 && ctas * s2 <= sm_count() + sm_count() / 4) { split = s2; break; }
    a.splitk = split;
    return launch_conv_gemm<64, 4, 0>(a, n, st);
}

extern "C" int hdn_conv_gemm_f32(const float *x, const float *wpk, const float *scale, const float *shift, const float *residual, float *out,
                                 int B, int Cin, int Cout, int H, int W, int ksize, int dilation, int valid, int relu, hdn_stream_t stream) {
    if (!x || !wpk || !out) return HDN_ERR_NULL;
    return conv_gemm_multi(1, &x, &wpk, &scale, &shift, &residual, nullptr, &out, B, Cin, Cout, H, W, ksize, dilation, valid, relu, 0,
                           (cudaStream_t)stream);
}

extern "C" int hdn_conv_gemm_ex_f32(const float *x, const float *wpk, const float *scale, const float *shift, const float *residual, float *out,
                                    int B, int Cin, int Cout, int H, int W, int ksize, int stride, int pad, int dilation, int relu,
                                    hdn_stream_t stream) {
    if (!x || !wpk || !out) return HDN_ERR_NULL;
    if (pad < 0) return HDN_ERR_SHAPE;
    return conv_gemm_multi(1, &x, &wpk, &scale, &shift, &residual, nullptr, &out, B, Cin, Cout, H, W, ksize, dilation, 0, relu, 0,
                           (cudaStream_t)stream, stride, pad);
}

extern "C" int hdn_conv_gemm_multi_f32(int n, const float *const *x_host, const float *const *wpk_host, const float *const *scale_host,
                                       const float *const *shift_host, float *const *out_host, int B, int Cin, int Cout, int H, int W, int ksize,
                                       int dilation, int valid, int relu, hdn_stream_t stream) {
    return conv_gemm_multi(n, x_host, wpk_host, scale_host, shift_host, nullptr, nullptr, out_host, B, Cin, Cout, H, W, ksize, dilation, valid, relu,
                           0, (cudaStream_t)stream);
}

extern "C" int hdn_head_project_multi_f32(int n, const float *const *x_host, const float *const *wpk_host, const float *const *scale_host,
                                          const float *const *shift_host, const float *const *w2_host, float *const *part_host, int B, int C,
                                          int H, int W, int L, hdn_stream_t stream) {
    if (!w2_host) return HDN_ERR_NULL;
    return conv_gemm_multi(n, x_host, wpk_host, scale_host, shift_host, nullptr, w2_host, part_host, B, C, C, H, W, 1, 1, 0, 1, L,
                           (cudaStream_t)stream);
}

extern "C" int hdn_conv_gemm_set_ts(int enable) {
    g_conv_ts = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
    return HDN_OK;
}

extern "C" int hdn_conv_gemm_set_pdl(int enable) {
    hdn::g_conv_pdl = enable ? 1 : 0;
    return HDN_OK;
}

extern "C" int hdn_conv_gemm_set_splitk(int enable) {
    g_conv_splitk = enable ? 1 : 0;
    return HDN_OK;
}

namespace hdn {
extern int g_conv_shift_multicast;  // conv_shift.cu
}

extern "C" int hdn_conv_gemm_set_shift(int mode) {
    g_conv_shift = mode ? 1 : 0;
    hdn::g_conv_shift_multicast = mode == 2 ? 1 : 0;
    return HDN_OK;
}
