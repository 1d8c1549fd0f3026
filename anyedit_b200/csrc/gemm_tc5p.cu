// Persistent tcgen05 contraction (second generation of gemm_tc5.cu; same operand plumbing).
//
//   * one CTA per SM, static round-robin over output tiles (128 x <=256), n-tile fastest so that the CTAs
//     running concurrently share A rows and weight tiles in L2;
//   * TMA producer warp / single-thread MMA issuer / 4 epilogue warps, smem ring of 3 x (16 KB A + 32 KB B);
//   * TWO accumulator buffers in TMEM (2 x 256 columns): the MMA warp starts tile t+1 while the epilogue
//     drains tile t; the TMA ring keeps running across tile boundaries;
//   * tile width is a run-time quantity (UMMA N in the instruction descriptor): N = 320 is 256 + 64, no padding work;
//   * epilogue in 64-column sub-tiles through 4 swizzled 16 KB staging buffers: the residual sub-tile is
//     TMA-PREFETCHED into the buffer, combined in place (bias / time-embedding row add / SiLU / GEGLU /
//     residual), and written back with a TMA store (2-D map for token matrices, 4-D NHWC map for conv
//     patches -- edge clipping is the tensor map's job, no per-row predicates, no scattered 16-byte stores).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace anysd {

constexpr int P_BM = 128, P_BK = 64, P_BN = 256, P_NSTG = 4, P_SUB = 64;
constexpr int P_THREADS = 384;                       // warps 0..3: TMA, MMA, TMEM-alloc, spare; warps 4..11: epilogue
constexpr int P_A_BYTES = P_BM * P_BK * 2;           // 16 KB
constexpr int P_STG_BYTES = P_BM * P_SUB * 2;        // 16 KB
// CTAS = 1: one CTA owns a 128 x <=256 tile, stage = 16 KB A + 32 KB B, 3 stages.
// CTAS = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2, UMMA M = 256) owns a 256 x <=256 tile; each CTA
//           stages its 128 rows of A and HALF of the B rows (16 + 16 KB), 4 stages: L2->SM operand traffic per
//           FLOP is two thirds of the single-CTA tile's, which is what bounds the 128-row kernel.
template <int CTAS>
struct PCfg {
    static constexpr int B_BYTES = (P_BN / CTAS) * P_BK * 2;
    static constexpr int STAGE_BYTES = P_A_BYTES + B_BYTES;
    static constexpr int STAGES = CTAS == 2 ? 4 : 3;
    // + 1024 alignment slack + 512 (mbarriers, TMEM slot) + the epilogue parameters of a tile staged by warp 3, double-buffered:
    // 2 x 256 bias, 2 x 256 folded-LayerNorm column sums, 2 x 128 x {-mean, rstd} row coefficients
    static constexpr int SMEM = STAGES * STAGE_BYTES + P_NSTG * P_STG_BYTES + 1024 + 512 + 2048 + 2048 + 2048;
};

struct PArgs {
    const float* bias;
    const float* rowadd;
    int M, N, ld_rowadd, rows_per_batch;
    int act, has_res;
    int num_kb, tiles_m, tiles_n, num_tiles;
    int bn, stage_tx;        // tile width (64 | 128 | 256) and TMA bytes per stage per CTA (A + its share of B)
    // conv geometry
    int Nimg, Ho, Wo;
    int BW, BH, NB, tiles_w, tiles_h;
    int kb_per_tap, stride;
    int pad_lo;              // zero rows / columns before the image: 1 (pad 1 on every side) or 0 (right / bottom padding only)
    // per-(32-row slab, channel) {sum, sum of squares} of the OUTPUT for the consumer's GroupNorm (anysd_gemm_params::stats)
    float* stats;
    int stats_hw, stats_spi, stats_nimg;     // rows per image, slabs per image (= hw / 32), image slots in the buffer
    // split-K (few output tiles, long K: the 8x8 level): a work unit = (tile, k-range); every unit dumps its raw fp32
    // accumulators, the LAST unit of a tile to arrive (per epilogue warp, atomic counter) adds the partials in split order
    // -- a fixed order, so the result does not depend on the arrival order -- and runs the normal epilogue
    int splits, kb_per_split, num_units;
    float* sk_ws;                            // [tile][cta of the pair][split][128 rows][bn] fp32
    unsigned int* sk_cnt;                    // [tile][cta of the pair][8 epilogue warps], zero between launches
    // LayerNorm around the contraction (anysd_gemm_params::row_stats / ln_stats; template parameter L):
    //   producer side: per-row {sum, sum of squares} of the OUTPUT per 64-column slab -> row_stats[slab][M] (float2)
    //   consumer side: A is the un-normalised x, W carries gamma, out = rstd_m (acc - mean_m colsum_n) + bias_n
    float* row_stats;
    const float* ln_stats;                   // [ln_slabs][M] float2 = the producer's row_stats
    const float* ln_colsum;                  // [N]
    int ln_slabs;
    float ln_eps, ln_inv_k;
};

// ---- PTX wrappers (same forms as gemm_tc5.cu) -------------------------------------------------------
__device__ __forceinline__ void pm_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void pm_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pm_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void p_tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void p_tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// ---- 2-CTA (cta_group::2) forms ----
__device__ __forceinline__ uint32_t p_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void p_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (+ expect_tx) on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void pm_expect_tx_remote(uint32_t bar, uint32_t bytes, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], %2;\n\t"
        "}" ::"r"(bar), "r"(cta), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar), "r"(cta) : "memory");
}
// TMA loads of a CTA pair: the transaction bytes are credited to the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void p_tma2_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_tma2_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void p_umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// completion of all prior MMAs of the pair -> arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void p_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}

__device__ __forceinline__ void p_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void p_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void p_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void p_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void p_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void p_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void p_tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void p_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void p_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void p_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void p_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint64_t p_sdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// Exact (erf) GELU  x Phi(x)  in 11 instructions: Phi(x) = 1 / (1 + 2^(-x P(x^2))) with a cubic P fitted (minimax on the ABSOLUTE
// error of x Phi(x), x^2 clamped at 36 -- beyond it the logistic is saturated either way) to |err| < 1.2e-5 for all x: 1/40 of the
// fp16 spacing at unit magnitude, 1/7 of it at the minimum of GELU (-0.17).  The epilogue of the GEGLU contractions is
// issue-bound ([measured] 59 % of the issue slots at 36 % tensor-pipe activity with the 18-instruction Abramowitz-Stegun erf
// this replaces); tests/test_gpu_ops.py::test_gelu_epilogue_accuracy pins the bound against torch's erf GELU.
__device__ __forceinline__ float p_gelu(float v) {
    const float t = fminf(v * v, 36.0f);
    float q = fmaf(t, 2.483638929e-05f, 7.36060983e-04f);
    q = fmaf(q, t, -0.10598272654f);
    q = fmaf(q, t, -2.30164716054f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * q));          // 2^(-x P(x^2))
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
}

// Column sums of a 32 x 32 block held one ROW per lane (r[c] = this row's value in column c): after five exchange
// rounds (16 + 8 + 4 + 2 + 1 shuffles) lane l holds the sum of column l over the 32 rows.  Fixed tree: deterministic.
__device__ __forceinline__ float p_colsum32(float (&r)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? r[i] : r[i + s];
            const float keep = up ? r[i + s] : r[i];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return r[0];
}

struct TileCoord {
    int n0, nw;          // first column, width (multiple of 16, <= 256)
    int m0;              // dense: first row
    int tw, th, tn;      // conv: patch indices
};

template <bool CONV, int CTAS>
__device__ __forceinline__ TileCoord tile_coord(const PArgs& p, int tile, int rank) {
    TileCoord c;
    const int nt = tile % p.tiles_n, mt = (tile / p.tiles_n) * CTAS + rank;   // a pair owns m-tiles 2t, 2t+1
    c.n0 = nt * p.bn;
    int w = p.N - c.n0;
    if (w > p.bn) w = p.bn;
    c.nw = CTAS == 2 ? ((w + 31) & ~31) : ((w + 15) & ~15);
    c.m0 = mt * P_BM;
    c.tw = c.th = c.tn = 0;
    if (CONV) {
        c.tw = mt % p.tiles_w;
        c.th = (mt / p.tiles_w) % p.tiles_h;
        c.tn = mt / (p.tiles_w * p.tiles_h);
    }
    return c;
}

// X ("extended"): the epilogue statistics, split-K and the GELU / QuickGELU activations are compiled in only when a launch
// asks for one of them -- the plain variant keeps the epilogue of the HBM- / issue-bound K = 320 projections and GEGLU
// contractions free of their branches and registers ([measured] the merged kernel ran the 161 linear launches of a forward
// in 6.9 ms instead of 6.2 ms).
// L ("LayerNorm"): the row-statistics output and the folded-LayerNorm input, dense contractions only (same reasoning: the
// variants that do not use them do not pay for them).
template <bool CONV, int CTAS, bool X, bool L>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_tc5p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR, const PArgs p) {
    extern __shared__ unsigned char p_smem_raw[];
    const uint32_t raw = smem_u32(p_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = p_smem_raw + (base - raw);
    constexpr int P_STAGES = PCfg<CTAS>::STAGES, P_STAGE_BYTES = PCfg<CTAS>::STAGE_BYTES;
    const int rank = CTAS == 2 ? (int)p_cluster_rank() : 0;
    const bool leader = rank == 0;
    const int cta_stride = gridDim.x / CTAS, cta_first = blockIdx.x / CTAS;     // tile loop runs over CTA pairs
    const uint32_t stg_base = base + P_STAGES * P_STAGE_BYTES;
    const uint32_t bar_base = stg_base + P_NSTG * P_STG_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (P_STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * P_STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * P_STAGES + 2 + b); };
    auto res_bar = [&](int b) { return bar_base + 8u * (2 * P_STAGES + 4 + b); };
    volatile uint32_t* tmem_slot =
        reinterpret_cast<volatile uint32_t*>(smem + P_STAGES * P_STAGE_BYTES + P_NSTG * P_STG_BYTES + 8 * (2 * P_STAGES + 4 + 16));
    auto cpfull_bar = [&](int b) { return bar_base + 256u + 8u * b; };       // epilogue parameters of tile t staged / consumed
    auto cpempty_bar = [&](int b) { return bar_base + 272u + 8u * b; };
    unsigned char* cp_base = smem + P_STAGES * P_STAGE_BYTES + P_NSTG * P_STG_BYTES + 512;
    float* cp_bias = reinterpret_cast<float*>(cp_base);                      // [2][256]
    float* cp_cs = reinterpret_cast<float*>(cp_base + 2048);                 // [2][256]  (L)
    float2* ln_buf = reinterpret_cast<float2*>(cp_base + 4096);              // [2][128] {-mean, rstd}  (L)
    // bias / column sums / row coefficients reach the epilogue through shared memory: every tile touches new columns, so a
    // direct __ldg is an L2 round trip that 8 warps x every 32-column step would each sit out ([measured] the folded-LayerNorm
    // epilogue with direct loads: GEGLU 65536 x 2560 x 320 in 222 us instead of 142)
    const bool staged = p.bias != nullptr || (L && p.ln_stats != nullptr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
        if (p.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < P_STAGES; ++s) {
            pm_init(full_bar(s), CTAS);      // one arrive.expect_tx per CTA of the pair (on the leader's barrier)
            pm_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            pm_init(tfull_bar(b), 1);
            pm_init(tempty_bar(b), 8 * CTAS);   // one arrive per epilogue warp of every CTA of the pair
        }
        for (int b = 0; b < 16; ++b) pm_init(res_bar(b), 1);      // 8 epilogue warps x 2 residual slabs
        for (int b = 0; b < 2; ++b) {
            pm_init(cpfull_bar(b), 1);
            pm_init(cpempty_bar(b), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    p_fence_before();
    __syncthreads();
    if (CTAS == 2) p_cluster_sync();         // the peer's barriers exist before anything remote touches them
    p_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: runs ahead across tile boundaries =====
            uint32_t g = 0;
            for (int unit = cta_first; unit < p.num_units; unit += cta_stride) {
                const bool sp = X && p.splits > 1;
                const int tile = sp ? unit / p.splits : unit;
                const int kb0 = sp ? (unit - tile * p.splits) * p.kb_per_split : 0;
                const int kb1 = sp ? (kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb) : p.num_kb;
                const TileCoord c = tile_coord<CONV, CTAS>(p, tile, rank);
                for (int kb = kb0; kb < kb1; ++kb, ++g) {
                    const int s = g % P_STAGES;
                    const uint32_t ph = (g / P_STAGES) & 1;
                    pm_wait(empty_bar(s), ph ^ 1);
                    const uint32_t sa = base + s * P_STAGE_BYTES, sb = sa + P_A_BYTES;
                    int tap = 0, c0 = 0, ky = 0, kx = 0;
                    if (CONV) {
                        tap = kb / p.kb_per_tap;
                        c0 = (kb - tap * p.kb_per_tap) * P_BK;
                        ky = tap / 3;
                        kx = tap - ky * 3;
                    }
                    if (CTAS == 2) {
                        pm_expect_tx_remote(full_bar(s), p.stage_tx, 0);         // credited to the leader's barrier
                        if (CONV)
                            p_tma2_load_4d(sa, &tmA, full_bar(s), c0, c.tw * p.BW * p.stride + kx - p.pad_lo,
                                           c.th * p.BH * p.stride + ky - p.pad_lo, c.tn * p.NB);
                        else
                            p_tma2_load_2d(sa, &tmA, full_bar(s), kb * P_BK, c.m0);
                        p_tma2_load_2d(sb, &tmB, full_bar(s), kb * P_BK, c.n0 + rank * (c.nw >> 1));   // its half of B
                    } else {
                        pm_expect_tx(full_bar(s), p.stage_tx);
                        if (CONV)
                            p_tma_load_4d(sa, &tmA, full_bar(s), c0, c.tw * p.BW * p.stride + kx - p.pad_lo,
                                          c.th * p.BH * p.stride + ky - p.pad_lo, c.tn * p.NB);
                        else
                            p_tma_load_2d(sa, &tmA, full_bar(s), kb * P_BK, c.m0);
                        p_tma_load_2d(sb, &tmB, full_bar(s), kb * P_BK, c.n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===== MMA issuer (the leader CTA issues for the pair) =====
            uint32_t g = 0, t = 0;
            for (int unit = cta_first; unit < p.num_units; unit += cta_stride, ++t) {
                const bool sp = X && p.splits > 1;
                const int tile = sp ? unit / p.splits : unit;
                const int kb0 = sp ? (unit - tile * p.splits) * p.kb_per_split : 0;
                const int kb1 = sp ? (kb0 + p.kb_per_split < p.num_kb ? kb0 + p.kb_per_split : p.num_kb) : p.num_kb;
                const TileCoord c = tile_coord<CONV, CTAS>(p, tile, 0);
                const uint32_t buf = t & 1, bph = (t >> 1) & 1;
                pm_wait(tempty_bar(buf), bph ^ 1);           // epilogue(s) have drained this accumulator
                p_fence_after();
                const uint32_t idesc = (1u << 4) | ((uint32_t)(c.nw >> 3) << 17) | ((uint32_t)((P_BM * CTAS) >> 4) << 24);
                const uint32_t d = tmem_base + buf * P_BN;
                for (int kb = kb0; kb < kb1; ++kb, ++g) {
                    const int s = g % P_STAGES;
                    const uint32_t ph = (g / P_STAGES) & 1;
                    pm_wait(full_bar(s), ph);
                    p_fence_after();
                    const uint32_t sa = base + s * P_STAGE_BYTES, sb = sa + P_A_BYTES;
                    const uint64_t ad = p_sdesc(sa), bd = p_sdesc(sb);
#pragma unroll
                    for (int k = 0; k < P_BK / 16; ++k) {
                        if (CTAS == 2) p_umma2(d, ad + 2 * k, bd + 2 * k, idesc, ((kb - kb0) | k) ? 1u : 0u);
                        else p_umma(d, ad + 2 * k, bd + 2 * k, idesc, ((kb - kb0) | k) ? 1u : 0u);
                    }
                    if (CTAS == 2) p_commit2(empty_bar(s)); else p_commit(empty_bar(s));
                }
                if (CTAS == 2) p_commit2(tfull_bar(buf)); else p_commit(tfull_bar(buf));
            }
        }
    } else if (warp == 3) {
        // ===== parameter stager: bias (and, L, the folded LayerNorm's column sums and per-row {-mean, rstd}) of the tile's columns /
        // rows into shared memory, up to two tiles ahead of the epilogue.  The row moments are folded slab by slab in index
        // order, the final combine runs in double. =====
        if (staged) {
            const bool ln_in = L && p.ln_stats != nullptr;
            const bool sp = X && p.splits > 1;
            uint32_t t = 0;
            for (int unit = cta_first; unit < p.num_units; unit += cta_stride, ++t) {
                const TileCoord c = tile_coord<CONV, CTAS>(p, sp ? unit / p.splits : unit, rank);
                const uint32_t b = t & 1, ph = (t >> 1) & 1;
                pm_wait(cpempty_bar(b), ph ^ 1);
                for (int i = lane; i < p.bn; i += 32) {
                    const int n = c.n0 + i;
                    cp_bias[b * P_BN + i] = (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
                    if (ln_in) cp_cs[b * P_BN + i] = n < p.N ? __ldg(p.ln_colsum + n) : 0.f;
                }
                if (ln_in) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int r = rr * 32 + lane;
                        int m = c.m0 + r;
                        if (m >= p.M) m = p.M - 1;
                        const float2* sp2 = reinterpret_cast<const float2*>(p.ln_stats) + m;
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll 5
                        for (int sl = 0; sl < p.ln_slabs; ++sl) {
                            const float2 f = __ldg(sp2 + (size_t)sl * p.M);
                            s1 += f.x;
                            s2 += f.y;
                        }
                        const double mean = (double)s1 * (double)p.ln_inv_k;
                        double var = (double)s2 * (double)p.ln_inv_k - mean * mean;
                        if (var < 0.0) var = 0.0;
                        ln_buf[b * P_BM + r] = make_float2(-(float)mean, rsqrtf((float)var + p.ln_eps));
                    }
                }
                __syncwarp();
                if (lane == 0) pm_arrive(cpfull_bar(b));
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 INDEPENDENT warps (no CTA-wide barriers).  TMEM lane group = warp % 4 (hardware rule),
        // lane = accumulator row.  Warps 4..7 take the even 64-column sub-tiles, warps 8..11 the odd ones.  Each warp
        // owns two 4 KB staging slabs [32 rows x 64 cols, 128B swizzle]: lane 0 TMA-prefetches the residual slab of
        // the warp's NEXT sub-tile while the current one is combined in place and TMA-stored. =====
        const int ew = warp - 4, lg = warp & 3, set = ew >> 2;
        const int row = lg * 32 + lane;
        constexpr uint32_t SLAB = 32 * 128;                                  // bytes
        const uint32_t wstg = stg_base + ew * 2 * SLAB;
        unsigned char* wstg_ptr = smem + P_STAGES * P_STAGE_BYTES + ew * 2 * SLAB;
        auto rbar = [&](uint32_t b) { return res_bar(0) + 8u * (ew * 2 + b); };
        const int acc_per_sub = (p.act == 2) ? 2 * P_SUB : P_SUB;   // accumulator columns feeding one 64-column store
        int sdx = 0, sdy = 0, sdn = 0;                               // slab origin inside a conv patch
        if (CONV) {
            const int r0 = lg * 32;
            sdx = r0 % p.BW;
            sdy = (r0 / p.BW) % p.BH;
            sdn = r0 / (p.BW * p.BH);
        }
        auto nsub_of = [&](const TileCoord& c) { return (c.nw + acc_per_sub - 1) / acc_per_sub; };
        auto slab_load = [&](uint32_t b, const TileCoord& c, int j) {        // residual slab -> staging buffer b
            const int col = ((p.act == 2) ? (c.n0 >> 1) : c.n0) + j * P_SUB;
            pm_expect_tx(rbar(b), SLAB);
            if (CONV) p_tma_load_4d(wstg + b * SLAB, &tmR, rbar(b), col, c.tw * p.BW + sdx, c.th * p.BH + sdy, c.tn * p.NB + sdn);
            else p_tma_load_2d(wstg + b * SLAB, &tmR, rbar(b), col, c.m0 + lg * 32);
        };
        auto slab_store = [&](uint32_t b, const TileCoord& c, int j) {
            const int col = ((p.act == 2) ? (c.n0 >> 1) : c.n0) + j * P_SUB;
            if (CONV) p_tma_store_4d(&tmO, wstg + b * SLAB, col, c.tw * p.BW + sdx, c.th * p.BH + sdy, c.tn * p.NB + sdn);
            else p_tma_store_2d(&tmO, wstg + b * SLAB, col, c.m0 + lg * 32);
            p_store_commit();
        };
        // this warp's first slab: residual prefetch before anything else
        uint32_t t = 0, q = 0;                                // tile counter, per-warp slab counter
        const bool split = X && p.splits > 1;
        if (p.has_res && lane == 0) {
            for (int ul = cta_first; ul < p.num_units; ul += cta_stride) {
                const TileCoord c0 = tile_coord<CONV, CTAS>(p, split ? ul / p.splits : ul, rank);
                if (set < nsub_of(c0)) { slab_load(0, c0, set); break; }
            }
        }
        for (int unit = cta_first; unit < p.num_units; unit += cta_stride, ++t) {
            const int tile = split ? unit / p.splits : unit;
            const TileCoord c = tile_coord<CONV, CTAS>(p, tile, rank);
            const uint32_t buf = t & 1, bph = (t >> 1) & 1;
            const int nsub = nsub_of(c);
            int img = 0;
            if (p.rowadd) {
                if (CONV) {
                    img = c.tn * p.NB + row / (p.BW * p.BH);
                    if (img >= p.Nimg) img = p.Nimg - 1;
                } else {
                    int m = c.m0 + row;
                    if (m >= p.M) m = p.M - 1;
                    img = m / p.rows_per_batch;
                }
            }
            const float* radd = p.rowadd ? p.rowadd + (size_t)img * p.ld_rowadd : nullptr;
            // this tile's staged parameters (warp 3); folded LayerNorm: this thread's row coefficients
            float ln_nmean = 0.f, ln_rstd = 1.f;
            const bool ln_in = L && p.ln_stats != nullptr;
            if (staged) pm_wait(cpfull_bar(buf), bph);
            const float* tb = cp_bias + buf * P_BN;
            const float* tcs = cp_cs + buf * P_BN;
            if (ln_in) {
                const float2 cf = ln_buf[buf * P_BM + row];
                ln_nmean = cf.x;
                ln_rstd = cf.y;
            }
            pm_wait(tfull_bar(buf), bph);
            p_fence_after();
            // ---- split-K: dump this unit's raw accumulators, find out whether this warp is the last of the tile's units ----
            bool sk_last = true;
            const float* sk_tile = nullptr;
            if (split) {
                const int ks = unit - tile * p.splits;
                float* wst = p.sk_ws + ((size_t)(tile * CTAS + rank) * p.splits) * (size_t)(P_BM * p.bn);
                sk_tile = wst;
                float* mine = wst + (size_t)ks * (P_BM * p.bn) + (size_t)row * p.bn;
                for (int j = set; j < nsub; j += 2) {
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        uint32_t r[32];
                        const int ac = j * P_SUB + half * 32;
                        __syncwarp();
                        p_tmem_ld32(tmem_base + buf * P_BN + ((uint32_t)(lg * 32) << 16) + ac, r);
                        p_tmem_wait_ld();
                        if (ac < c.nw) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                __stcg(reinterpret_cast<float4*>(mine + ac) + i,
                                       make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                                   __uint_as_float(r[4 * i + 3])));
                        }
                    }
                }
                p_fence_before();                              // accumulator read: hand the TMEM buffer back now
                __syncwarp();
                if (lane == 0) {
                    if (CTAS == 2) pm_arrive_remote(tempty_bar(buf), 0); else pm_arrive(tempty_bar(buf));
                }
                __threadfence();                               // partials visible before the arrival is counted
                __syncwarp();
                unsigned int old = 0;
                unsigned int* cnt = p.sk_cnt + (size_t)(tile * CTAS + rank) * 8 + ew;
                if (lane == 0) {
                    old = atomicAdd(cnt, 1u);
                    if (old == (unsigned)p.splits - 1) *cnt = 0;   // re-arm for the next launch (stream-ordered)
                }
                old = __shfl_sync(0xffffffffu, old, 0);
                sk_last = old == (unsigned)p.splits - 1;
                __threadfence();                               // ... and the other units' partials visible to the adder
            }
            for (int j = set; j < nsub; j += 2, ++q) {
                const uint32_t b = q & 1;
                if (lane == 0) {
                    p_store_wait_read<0>();                    // the previous slab's store has released buffer b^1
                    if (p.has_res) {                           // prefetch the residual of this warp's next slab
                        if (j + 2 < nsub) {
                            slab_load(b ^ 1, c, j + 2);
                        } else {
                            for (int ul = unit + cta_stride; ul < p.num_units; ul += cta_stride) {
                                const TileCoord c2 = tile_coord<CONV, CTAS>(p, split ? ul / p.splits : ul, rank);
                                if (set < nsub_of(c2)) { slab_load(b ^ 1, c2, set); break; }
                            }
                        }
                    }
                }
                __syncwarp();
                unsigned char* stg = wstg_ptr + b * SLAB + lane * 128;
                if (p.has_res) pm_wait(rbar(b), (q >> 1) & 1);
                if (!sk_last) continue;                        // another unit finishes this tile (barrier phases stay in step)
                float rs1 = 0.f, rs2 = 0.f;                    // L: this row's {sum, sum of squares} over the 64-column slab
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {         // 2 x 32 output columns
                    float v[32];
                    const int oc = half * 32;                  // output column offset inside the sub-tile
                    if (p.act == 2) {
                        // GEGLU: 64 accumulator columns (a_j, gate_j interleaved) -> 32 outputs
                        uint32_t r0[32], r1[32];
                        const int ac = j * 2 * P_SUB + half * 64;
                        __syncwarp();
                        p_tmem_ld32(tmem_base + buf * P_BN + ((uint32_t)(lg * 32) << 16) + ac, r0);
                        p_tmem_ld32(tmem_base + buf * P_BN + ((uint32_t)(lg * 32) << 16) + ac + 32, r1);
                        p_tmem_wait_ld();
                        if (ln_in) {                            // folded LayerNorm (bias always present: beta W^T + b)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b0 = *(reinterpret_cast<const float4*>(tb + ac) + i);
                                const float4 b1 = *(reinterpret_cast<const float4*>(tb + ac + 32) + i);
                                const float4 c0 = *(reinterpret_cast<const float4*>(tcs + ac) + i);
                                const float4 c1 = *(reinterpret_cast<const float4*>(tcs + ac + 32) + i);
                                r0[4 * i] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c0.x, __uint_as_float(r0[4 * i])), b0.x));
                                r0[4 * i + 1] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c0.y, __uint_as_float(r0[4 * i + 1])), b0.y));
                                r0[4 * i + 2] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c0.z, __uint_as_float(r0[4 * i + 2])), b0.z));
                                r0[4 * i + 3] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c0.w, __uint_as_float(r0[4 * i + 3])), b0.w));
                                r1[4 * i] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c1.x, __uint_as_float(r1[4 * i])), b1.x));
                                r1[4 * i + 1] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c1.y, __uint_as_float(r1[4 * i + 1])), b1.y));
                                r1[4 * i + 2] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c1.z, __uint_as_float(r1[4 * i + 2])), b1.z));
                                r1[4 * i + 3] = __float_as_uint(fmaf(ln_rstd, fmaf(ln_nmean, c1.w, __uint_as_float(r1[4 * i + 3])), b1.w));
                            }
                        } else if (p.bias) {                    // N % 128 == 0 for GEGLU: the 64 columns are in range
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b0 = *(reinterpret_cast<const float4*>(tb + ac) + i);
                                const float4 b1 = *(reinterpret_cast<const float4*>(tb + ac + 32) + i);
                                r0[4 * i] = __float_as_uint(__uint_as_float(r0[4 * i]) + b0.x);
                                r0[4 * i + 1] = __float_as_uint(__uint_as_float(r0[4 * i + 1]) + b0.y);
                                r0[4 * i + 2] = __float_as_uint(__uint_as_float(r0[4 * i + 2]) + b0.z);
                                r0[4 * i + 3] = __float_as_uint(__uint_as_float(r0[4 * i + 3]) + b0.w);
                                r1[4 * i] = __float_as_uint(__uint_as_float(r1[4 * i]) + b1.x);
                                r1[4 * i + 1] = __float_as_uint(__uint_as_float(r1[4 * i + 1]) + b1.y);
                                r1[4 * i + 2] = __float_as_uint(__uint_as_float(r1[4 * i + 2]) + b1.z);
                                r1[4 * i + 3] = __float_as_uint(__uint_as_float(r1[4 * i + 3]) + b1.w);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            v[i] = __uint_as_float(r0[2 * i]) * p_gelu(__uint_as_float(r0[2 * i + 1]));
                            v[16 + i] = __uint_as_float(r1[2 * i]) * p_gelu(__uint_as_float(r1[2 * i + 1]));
                        }
                    } else {
                        const int ac = j * P_SUB + oc;
                        const int nb = c.n0 + ac;
                        if (split) {
                            // the tile's partials in split order (fixed: independent of which unit arrived last)
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0.f;
                            if (ac < c.nw) {
                                for (int ks = 0; ks < p.splits; ++ks) {
                                    const float4* src = reinterpret_cast<const float4*>(sk_tile + (size_t)ks * (P_BM * p.bn) + (size_t)row * p.bn + ac);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const float4 f = __ldcg(src + i);
                                        v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
                                    }
                                }
                            }
                        } else {
                            uint32_t r[32];
                            __syncwarp();
                            p_tmem_ld32(tmem_base + buf * P_BN + ((uint32_t)(lg * 32) << 16) + ac, r);
                            p_tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        }
                        // bias (and the folded LayerNorm's column sums) from the staged copy: zero beyond column N
                        if (ln_in) {                           // folded LayerNorm: rstd (acc - mean colsum) + (beta W^T + b)
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 bb = *(reinterpret_cast<const float4*>(tb + ac) + i);
                                const float4 cc = *(reinterpret_cast<const float4*>(tcs + ac) + i);
                                v[4 * i] = fmaf(ln_rstd, fmaf(ln_nmean, cc.x, v[4 * i]), bb.x);
                                v[4 * i + 1] = fmaf(ln_rstd, fmaf(ln_nmean, cc.y, v[4 * i + 1]), bb.y);
                                v[4 * i + 2] = fmaf(ln_rstd, fmaf(ln_nmean, cc.z, v[4 * i + 2]), bb.z);
                                v[4 * i + 3] = fmaf(ln_rstd, fmaf(ln_nmean, cc.w, v[4 * i + 3]), bb.w);
                            }
                        } else if (p.bias) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 bb = *(reinterpret_cast<const float4*>(tb + ac) + i);
                                v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
                            }
                        }
                        if (radd) {
                            if (nb + 32 <= p.N) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 bb = __ldg(reinterpret_cast<const float4*>(radd + nb) + i);
                                    v[4 * i] += bb.x; v[4 * i + 1] += bb.y; v[4 * i + 2] += bb.z; v[4 * i + 3] += bb.w;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 32; ++i)
                                    if (nb + i < p.N) v[i] += __ldg(radd + nb + i);
                            }
                        }
                        if (p.act == 1) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
                        } else if (X && p.act == 3) {         // nn.GELU() (CLIP-H MLP, Resampler FeedForward)
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = p_gelu(v[i]);
                        } else if (X && p.act == 4) {         // QuickGELU (CLIP-L MLP)
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = quick_gelu_f(v[i]);
                        }
                    }
                    // combine with the prefetched residual in place; 128-byte swizzled staging row
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const int chunk = (oc >> 3) + ch;       // 16-byte chunk index 0..7 in the 128-byte row
                        uint4* slot = reinterpret_cast<uint4*>(stg + ((chunk ^ (lane & 7)) << 4));
                        float* vv = v + ch * 8;
                        if (p.has_res) {
                            float rr[8];
                            unpack8(*slot, rr);
#pragma unroll
                            for (int i = 0; i < 8; ++i) vv[i] += rr[i];
                        }
                        *slot = pack8(vv);
                        if (L && p.row_stats != nullptr) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                rs1 += vv[i];
                                rs2 = fmaf(vv[i], vv[i], rs2);
                            }
                        }
                    }
                    if (X && p.stats != nullptr) {
                        // GroupNorm statistics of what was just produced (fp32, before the fp16 rounding): per channel over
                        // this warp's 32 rows.  Each (slab, channel) cell is written exactly once per launch, by one warp.
                        const int r0 = CONV ? 0 : c.m0 + lg * 32;
                        int img, sii;
                        if (CONV) {
                            const int ppi = p.BW * p.BH;                           // patch pixels per image (multiple of 32)
                            img = c.tn * p.NB + sdn;
                            sii = (c.th * p.tiles_w + c.tw) * (ppi >> 5) + (((lg * 32) % ppi) >> 5);
                        } else {
                            img = r0 / p.stats_hw;
                            sii = (r0 - img * p.stats_hw) >> 5;
                        }
                        float q2[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) q2[i] = v[i] * v[i];
                        const float s1 = p_colsum32(v, lane);
                        const float s2 = p_colsum32(q2, lane);
                        const int col = c.n0 + j * P_SUB + oc + lane;
                        if (img < p.stats_nimg && col < p.N)
                            reinterpret_cast<float2*>(p.stats)[((size_t)img * p.stats_spi + sii) * p.N + col] = make_float2(s1, s2);
                    }
                }
                if (L && p.row_stats != nullptr) {             // one cell per (64-column slab, row): written once, by one thread
                    const int m = c.m0 + row;
                    if (m < p.M)
                        reinterpret_cast<float2*>(p.row_stats)[(size_t)((c.n0 >> 6) + j) * p.M + m] = make_float2(rs1, rs2);
                }
                p_fence_async_smem();                          // generic-proxy writes -> visible to the TMA store
                __syncwarp();
                if (lane == 0) slab_store(b, c, j);
            }
            // accumulator fully read by this warp: hand the TMEM buffer back to the MMA warp (split-K did so above)
            if (!split) {
                p_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CTAS == 2) pm_arrive_remote(tempty_bar(buf), 0); else pm_arrive(tempty_bar(buf));
                }
            }
            if (staged) {                                      // ... and the staged parameters back to warp 3
                __syncwarp();
                if (lane == 0) pm_arrive(cpempty_bar(buf));
            }
        }
        if (lane == 0) p_store_wait<0>();                      // smem must outlive the last store
    }
    p_fence_before();
    __syncthreads();
    if (CTAS == 2) p_cluster_sync();         // neither CTA leaves while the pair's MMAs / remote arrives may still touch it
    if (warp == 2) {
        if (CTAS == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnP p_get_encode() {
    static EncodeTiledFnP fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFnP)f;
    }
    return fn;
}

// Encoded tensor maps are kept in a small direct-mapped table keyed by everything that goes into the encoding: a sampling
// loop launches the same few hundred (pointer, shape) combinations over and over (the caching allocator hands the same
// blocks back), and cuTensorMapEncodeTiled is a few microseconds of host time per map, four maps per launch.
struct PMapKey {
    const void* ptr;
    uint64_t d0, d1, d2, d3, ld;
    uint32_t b0, b1, b2, b3, es, rank;
    bool operator==(const PMapKey& o) const {
        return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && d3 == o.d3 && ld == o.ld && b0 == o.b0 && b1 == o.b1 &&
               b2 == o.b2 && b3 == o.b3 && es == o.es && rank == o.rank;
    }
};
struct PMapSlot {
    PMapKey key;
    CUtensorMap map;
    bool valid;
};
constexpr int P_MAP_SLOTS = 4096;
static thread_local PMapSlot* p_map_table = nullptr;
static PMapSlot* p_map_slot(const PMapKey& k) {
    if (p_map_table == nullptr) p_map_table = (PMapSlot*)calloc(P_MAP_SLOTS, sizeof(PMapSlot));
    uint64_t h = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
    h ^= (k.d0 * 31 + k.d1) * 0xC2B2AE3D27D4EB4Full + (k.d2 * 131 + k.d3 * 17 + k.ld) * 0x165667B19E3779F9ull;
    h ^= ((uint64_t)k.b0 << 40) ^ ((uint64_t)k.b1 << 28) ^ ((uint64_t)k.b2 << 16) ^ ((uint64_t)k.b3 << 4) ^ k.es ^ ((uint64_t)k.rank << 60);
    return p_map_table ? &p_map_table[(h >> 20) % P_MAP_SLOTS] : nullptr;
}

static bool p_map_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t bi, uint32_t bo) {
    const PMapKey key = {ptr, inner, outer, 0, 0, ld, bi, bo, 0, 0, 1, 2};
    PMapSlot* slot = p_map_slot(key);
    if (slot && slot->valid && slot->key == key) {
        *tm = slot->map;
        return true;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {bi, bo};
    cuuint32_t es[2] = {1, 1};
    const bool ok = p_get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    if (ok && slot) {
        slot->key = key;
        slot->map = *tm;
        slot->valid = true;
    }
    return ok;
}
// NHWC tensor [N, H, W, C] with row pitch ld (elements per pixel), box 64 x BW x BH x NB, element stride s in W/H
static bool p_map_nhwc(CUtensorMap* tm, const void* ptr, int N, int H, int W, int C, int ld, int BW, int BH, int NB, int s) {
    const PMapKey key = {ptr, (uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N, (uint64_t)ld, 64u, (uint32_t)BW, (uint32_t)BH, (uint32_t)NB,
                         (uint32_t)s, 4};
    PMapSlot* slot = p_map_slot(key);
    if (slot && slot->valid && slot->key == key) {
        *tm = slot->map;
        return true;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {64u, (cuuint32_t)(BW * s), (cuuint32_t)(BH * s), (cuuint32_t)NB};
    cuuint32_t es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    const bool ok = p_get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    if (ok && slot) {
        slot->key = key;
        slot->map = *tm;
        slot->valid = true;
    }
    return ok;
}

static int p_pow2_ceil(int v) {
    int r = 1;
    while (r < v) r <<= 1;
    return r;
}
static int p_pick_extent(int len, int cap) {
    int lim = p_pow2_ceil(len);
    if (lim > cap) lim = cap;
    for (int c = lim; c >= 4; c >>= 1)
        if (len % c == 0) return c;
    return lim;
}

int launch_upsample2x(const __half* src, __half* dst, int N, int H, int W, int C, cudaStream_t st);

// GroupNorm statistics from the epilogue: possible when every 32-row slab of the output lies inside one image and the
// conv patches tile the image exactly.  Returns the slabs per image (rows_per_batch / 32), 0 when not possible.
static void p_conv_patch(const anysd_gemm_params* q, int* Ho, int* Wo, int* BW, int* BH, int* NB);
int tc5p_stats_slabs(const anysd_gemm_params* q) {
    if (q->act == 2 || q->out_dtype != ANYSD_F16) return 0;
    const int hw = q->rows_per_batch;
    if (hw <= 0 || hw % 32 != 0 || q->M % hw != 0) return 0;
    if (q->conv) {
        int Ho, Wo, BW, BH, NB;
        p_conv_patch(q, &Ho, &Wo, &BW, &BH, &NB);
        if (Ho * Wo != hw || (BW * BH) % 32 != 0 || Wo % BW != 0 || Ho % BH != 0) return 0;
    }
    return hw / 32;
}

static void p_select(const anysd_gemm_params* q, int tiles_m, int num_kb, int* ctas_out, int* bn_out, int* splits_out);
static size_t p_splitk_bytes(int tiles_m, int N, int ctas, int bn, int splits);
// scratch the caller should provide for this contraction to run split-K (0: the schedule does not split it)
size_t tc5p_splitk_bytes(const anysd_gemm_params* q) {
    int tiles_m, num_kb;
    if (q->conv) {
        int Ho, Wo, BW, BH, NB;
        p_conv_patch(q, &Ho, &Wo, &BW, &BH, &NB);
        tiles_m = cdiv(Wo, BW) * cdiv(Ho, BH) * cdiv(q->Nimg, NB);
        num_kb = 9 * (q->Cin / P_BK);
    } else {
        tiles_m = cdiv(q->M, P_BM);
        num_kb = cdiv(q->K, P_BK);
    }
    int ctas, bn, splits;
    p_select(q, tiles_m, num_kb, &ctas, &bn, &splits);
    return p_splitk_bytes(tiles_m, q->N, ctas, bn, splits);
}

// LayerNorm around a contraction (row_stats output / ln_stats input): NULL when this kernel can do it, else the reason.
const char* tc5p_ln_unsupported(const anysd_gemm_params* q) {
    if (q->conv) return "dense contractions only";
    if (q->out_dtype != ANYSD_F16) return "fp16 output only";
    if (q->row_stats != nullptr) {
        if (q->act != 0) return "row statistics need act = 0";
        if (q->N % 64 != 0) return "row statistics need N % 64 == 0";
        if ((uintptr_t)q->row_stats % 8) return "row_stats must be 8-byte aligned";
    }
    if (q->ln_stats != nullptr) {
        if (q->K % 64 != 0) return "folded LayerNorm needs K % 64 == 0";
        if (q->bias == nullptr || q->ln_colsum == nullptr) return "folded LayerNorm needs bias (beta W^T + b) and ln_colsum";
        if (((uintptr_t)q->ln_stats % 8) || ((uintptr_t)q->ln_colsum % 16)) return "ln_stats / ln_colsum misaligned";
        if (q->N % 32 != 0) return "folded LayerNorm needs N % 32 == 0";
        if (q->act != 0 && q->act != 1 && q->act != 2) return "folded LayerNorm: act must be 0, 1 or 2";
    }
    return nullptr;
}

bool tc5p_supported(const anysd_gemm_params* q) {
    if (q->out_dtype != ANYSD_F16) return false;
    if (q->N % 8 != 0 || q->K % 8 != 0) return false;
    if (q->act == 2 && q->N % 128 != 0) return false;
    const int n_out = q->act == 2 ? q->N / 2 : q->N;
    if (((uintptr_t)q->out % 16) || q->ldo % 8 != 0 || n_out % 8 != 0) return false;
    if (q->residual && (((uintptr_t)q->residual % 16) || q->ldr % 8 != 0)) return false;
    if (q->bias && ((uintptr_t)q->bias % 16)) return false;
    if (q->rowadd && (((uintptr_t)q->rowadd % 16) || q->ld_rowadd % 4 != 0)) return false;
    if (q->conv) {
        if (q->Cin % P_BK != 0) return false;
        if (q->upsample) {
            const size_t need = (size_t)q->Nimg * (2 * q->H) * (2 * q->Wd) * q->Cin * sizeof(__half);
            if (q->workspace == nullptr || q->workspace_bytes < need || ((uintptr_t)q->workspace % 16)) return false;
        }
    }
    return p_get_encode() != nullptr;
}

template <bool CONV, int CTAS, bool X, bool L = false>
static int p_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
                    const PArgs& a, cudaStream_t st) {
    static bool done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc5p_kernel<CONV, CTAS, X, L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             PCfg<CTAS>::SMEM);
        if (e != cudaSuccess) {
            set_error("tcgen05 gemm: smem opt-in failed: %s", cudaGetErrorString(e));
            return ANYSD_ECUDA;
        }
        done[dev] = true;
    }
    int grid = sm_count();
    if (grid > a.num_units * CTAS) grid = a.num_units * CTAS;
    grid -= grid % CTAS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = PCfg<CTAS>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CTAS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CTAS == 2 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc5p_kernel<CONV, CTAS, X, L>, tmA, tmB, tmO, tmR, a);
    if (e != cudaSuccess) {
        set_error("tcgen05 gemm launch failed: %s", cudaGetErrorString(e));
        return ANYSD_ECUDA;
    }
    return check_launch(CONV ? "conv3x3 (tcgen05 persistent)" : "gemm (tcgen05 persistent)");
}

static void p_conv_patch(const anysd_gemm_params* q, int* Ho, int* Wo, int* BW, int* BH, int* NB) {
    const int Hin = q->H << q->upsample, Win = q->Wd << q->upsample;
    *Ho = q->conv_pad ? (Hin - 2) / 2 + 1 : (Hin - 1) / q->stride + 1;
    *Wo = q->conv_pad ? (Win - 2) / 2 + 1 : (Win - 1) / q->stride + 1;
    *BW = p_pick_extent(*Wo, 128);
    *BH = p_pick_extent(*Ho, 128 / *BW);
    *NB = 128 / (*BW * *BH);
}

// K split: a function of the PER-IMAGE geometry only -- never of the batch -- because splitting changes the fp32 summation
// order: a layer either always runs as `splits` ordered partial sums or never does, so every output bit stays independent
// of how many images share the batch (and of the grid).  Layers with at most one 128-row tile per image (8x8 maps and
// smaller) and a very long K are the ones that leave most SMs idle at sampling batch sizes.  [measured, B200, batch 16,
// tests/diag_splitk.py] conv 2560->1280 @8x8 (360 k-blocks): 110.7 us unsplit, 86.2 / 78.6 / 106.9 us with 2 / 3 / 4
// partials; conv 1280->1280 @8x8 (180 k-blocks): 49.5 us unsplit, 52.3 / 55.4 us with 2 / 3 (the dump + ordered add costs
// more than the idle SMs); dense M = 1024 contractions lose 30-50 %.  Hence: 3 partials from 256 k-blocks on, convs only.
// ANYSD_GEMM_SPLITK=0|n forces a count (experiments).
static int p_geometry_splits(const anysd_gemm_params* q, int num_kb) {
    static const char* force_sk = getenv("ANYSD_GEMM_SPLITK");
    if (q->act == 2 || q->out_dtype != ANYSD_F16 || q->row_stats != nullptr || q->ln_stats != nullptr) return 1;
    int rows_per_image = q->rows_per_batch;
    if (q->conv) {
        int Ho, Wo, BW, BH, NB;
        p_conv_patch(q, &Ho, &Wo, &BW, &BH, &NB);
        rows_per_image = Ho * Wo;
    }
    if (rows_per_image <= 0 || rows_per_image > P_BM) return 1;
    if (force_sk) {
        const int f = atoi(force_sk);
        return (f <= 1 || num_kb / f < 8) ? 1 : (f > 8 ? 8 : f);
    }
    return (q->conv && num_kb >= 256) ? 3 : 1;
}

// Tile shape selection.  Candidates: CTA pairs (cta_group::2, 256-row tiles, less L2 traffic per FLOP) or single CTAs, tile
// width 256 / 192 / 128 / 64 (GEGLU pairs need >= 128).  Cost model = scheduling rounds of the work units (tiles x splits) on
// the 148 SMs x per-unit time ~ (bn + 270) x (k-blocks of the unit + 8 when split: the partial dump and the ordered add):
// deep levels (8 row tiles x 5 column tiles, K = 11520 .. 23040) otherwise leave most of the machine idle while every busy
// SM streams a full K operand pair through its L2 port; mid levels suffer from wave quantisation.
// ANYSD_GEMM_CTAS=1|2 / ANYSD_GEMM_BN force a choice.
static void p_select(const anysd_gemm_params* q, int tiles_m, int num_kb, int* ctas_out, int* bn_out, int* splits_out) {
    static const char* force = getenv("ANYSD_GEMM_CTAS");
    static const char* force_bn = getenv("ANYSD_GEMM_BN");
    const int sp = splits_out ? p_geometry_splits(q, num_kb) : 1;
    int ctas = 1, bn = 256;
    double best = -1;
    for (int c = 2; c >= 1; --c) {
        if (c == 2 && (tiles_m < 2 || sm_count() < 2)) continue;
        if (force && (force[0] == '1' || force[0] == '2') && c != force[0] - '0') continue;
        static const int widths[4] = {256, 192, 128, 64};     // 192 balances N = 320 (192 + 128) and divides 960 / 1920
        for (int wi = 0; wi < 4; ++wi) {
            const int w = widths[wi];
            if (w < (q->act == 2 ? 128 : 64) || (q->act == 2 && w == 192)) continue;
            if (force_bn && atoi(force_bn) != w) continue;
            const long units = (long)cdiv(tiles_m, c) * cdiv(q->N, w) * sp;
            const long slots = sm_count() / c;
            const long rounds = (units + slots - 1) / slots;
            const double cost = (double)rounds * (w + 270) * ((double)cdiv(num_kb, sp) + (sp > 1 ? 8.0 : 0.0));
            if (best < 0 || cost < best) { best = cost; ctas = c; bn = w; }
        }
    }
    *ctas_out = ctas;
    *bn_out = bn;
    if (splits_out) *splits_out = sp;
}
static size_t p_splitk_bytes(int tiles_m, int N, int ctas, int bn, int splits) {
    if (splits <= 1) return 0;
    return (size_t)cdiv(tiles_m, ctas) * cdiv(N, bn) * ctas * splits * P_BM * bn * sizeof(float);
}

int launch_gemm_tc5p(const anysd_gemm_params* q, cudaStream_t st) {
    PArgs a;
    a.stats = nullptr;
    a.stats_hw = a.stats_spi = a.stats_nimg = 0;
    if (q->stats != nullptr) {
        a.stats_spi = tc5p_stats_slabs(q);
        if (a.stats_spi == 0 || q->stats_images <= 0) {
            set_error("gemm: output statistics requested for a shape that cannot produce them (M=%d rows_per_batch=%d conv=%d act=%d); "
                      "query anysd_gemm_stats_slabs first", q->M, q->rows_per_batch, q->conv, q->act);
            return ANYSD_EUNSUPPORTED;
        }
        a.stats = q->stats;
        a.stats_hw = q->rows_per_batch;
        a.stats_nimg = q->stats_images;
    }
    a.row_stats = q->row_stats;
    a.ln_stats = q->ln_stats;
    a.ln_colsum = q->ln_colsum;
    a.ln_slabs = q->K / 64;
    a.ln_eps = q->ln_eps;
    a.ln_inv_k = 1.0f / (float)q->K;
    const bool ln = q->row_stats != nullptr || q->ln_stats != nullptr;
    if (ln) {
        const char* why = tc5p_ln_unsupported(q);
        if (why) {
            set_error("gemm: LayerNorm fold / row statistics: %s (M=%d N=%d K=%d act=%d)", why, q->M, q->N, q->K, q->act);
            return ANYSD_EUNSUPPORTED;
        }
    }
    a.bias = q->bias;
    a.rowadd = q->rowadd;
    a.M = q->M;
    a.N = q->N;
    a.ld_rowadd = q->ld_rowadd;
    a.rows_per_batch = q->rows_per_batch > 0 ? q->rows_per_batch : 1;
    a.act = q->act;
    a.has_res = q->residual != nullptr;
    a.num_kb = cdiv(q->K, P_BK);
    a.Nimg = q->Nimg;
    a.BW = a.BH = a.NB = a.tiles_w = a.tiles_h = 1;
    a.kb_per_tap = 1;
    a.stride = q->conv ? q->stride : 1;
    a.pad_lo = (q->conv && q->conv_pad) ? 0 : 1;
    a.Ho = a.Wo = 0;
    const int n_out = q->act == 2 ? q->N / 2 : q->N;
    CUtensorMap tmA, tmB, tmO, tmR;
    bool ok = true;
    if (q->conv) {
        const void* img = q->A;
        int Hin = q->H, Win = q->Wd;
        if (q->upsample) {
            int rc = launch_upsample2x((const __half*)q->A, (__half*)q->workspace, q->Nimg, q->H, q->Wd, q->Cin, st);
            if (rc) return rc;
            img = q->workspace;
            Hin *= 2;
            Win *= 2;
        }
        a.Ho = q->conv_pad ? (Hin - 2) / 2 + 1 : (Hin - 1) / a.stride + 1;
        a.Wo = q->conv_pad ? (Win - 2) / 2 + 1 : (Win - 1) / a.stride + 1;
        a.BW = p_pick_extent(a.Wo, 128);
        a.BH = p_pick_extent(a.Ho, 128 / a.BW);
        a.NB = 128 / (a.BW * a.BH);
        a.tiles_w = cdiv(a.Wo, a.BW);
        a.tiles_h = cdiv(a.Ho, a.BH);
        a.tiles_m = a.tiles_w * a.tiles_h * cdiv(q->Nimg, a.NB);
        a.kb_per_tap = q->Cin / P_BK;
        a.num_kb = 9 * a.kb_per_tap;
        ok = ok && p_map_nhwc(&tmA, img, q->Nimg, Hin, Win, q->Cin, q->Cin, a.BW, a.BH, a.NB, a.stride);
        // output / residual move in 32-pixel slabs of the 128-pixel patch (one per epilogue warp)
        const int sw = a.BW < 32 ? a.BW : 32, sh = (32 / sw) < a.BH ? (32 / sw) : a.BH, sn = 32 / (sw * sh);
        ok = ok && p_map_nhwc(&tmO, q->out, q->Nimg, a.Ho, a.Wo, n_out, q->ldo, sw, sh, sn, 1);
        if (q->residual) ok = ok && p_map_nhwc(&tmR, q->residual, q->Nimg, a.Ho, a.Wo, n_out, q->ldr, sw, sh, sn, 1);
    } else {
        a.tiles_m = cdiv(q->M, P_BM);
        ok = ok && p_map_2d(&tmA, q->A, (uint64_t)q->K, (uint64_t)q->M, (uint64_t)q->lda, P_BK, P_BM);
        ok = ok && p_map_2d(&tmO, q->out, (uint64_t)n_out, (uint64_t)q->M, (uint64_t)q->ldo, P_SUB, 32);
        if (q->residual) ok = ok && p_map_2d(&tmR, q->residual, (uint64_t)n_out, (uint64_t)q->M, (uint64_t)q->ldr, P_SUB, 32);
    }
    int ctas = 1, bn = 256, splits = 1;
    p_select(q, a.tiles_m, a.num_kb, &ctas, &bn, &splits);
    const size_t sk_need = p_splitk_bytes(a.tiles_m, q->N, ctas, bn, splits);
    if (splits > 1 && (q->splitk_workspace == nullptr || q->splitk_workspace_bytes < sk_need || q->splitk_counters == nullptr ||
                       (size_t)cdiv(a.tiles_m, ctas) * cdiv(q->N, bn) * ctas * 8 * sizeof(unsigned int) > q->splitk_counters_bytes)) {
        splits = 1;                                   // no scratch from the caller: the plain schedule
        p_select(q, a.tiles_m, a.num_kb, &ctas, &bn, nullptr);
    }
    a.bn = bn;
    a.tiles_n = cdiv(q->N, bn);
    a.stage_tx = P_A_BYTES + (bn / ctas) * P_BK * 2;
    ok = ok && p_map_2d(&tmB, q->W, (uint64_t)q->K, (uint64_t)q->N, (uint64_t)q->ldw, P_BK, bn / ctas);
    if (!ok) {
        set_error("tcgen05 gemm: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d conv=%d)", q->M, q->N, q->K, q->conv);
        return ANYSD_ECUDA;
    }
    if (!q->residual) tmR = tmO;
    a.num_tiles = cdiv(a.tiles_m, ctas) * a.tiles_n;
    a.splits = splits;
    a.kb_per_split = cdiv(a.num_kb, splits);
    a.num_units = a.num_tiles * splits;
    a.sk_ws = (float*)q->splitk_workspace;
    a.sk_cnt = (unsigned int*)q->splitk_counters;
    const bool ext = a.stats != nullptr || a.splits > 1 || a.act >= 3;
    if (ln) {
        if (ext) {
            set_error("gemm: LayerNorm fold / row statistics cannot be combined with GroupNorm statistics, split-K or GELU epilogues");
            return ANYSD_EUNSUPPORTED;
        }
        return ctas == 2 ? p_launch<false, 2, false, true>(tmA, tmB, tmO, tmR, a, st) : p_launch<false, 1, false, true>(tmA, tmB, tmO, tmR, a, st);
    }
    if (ext) {
        if (q->conv) return ctas == 2 ? p_launch<true, 2, true>(tmA, tmB, tmO, tmR, a, st) : p_launch<true, 1, true>(tmA, tmB, tmO, tmR, a, st);
        return ctas == 2 ? p_launch<false, 2, true>(tmA, tmB, tmO, tmR, a, st) : p_launch<false, 1, true>(tmA, tmB, tmO, tmR, a, st);
    }
    if (q->conv) return ctas == 2 ? p_launch<true, 2, false>(tmA, tmB, tmO, tmR, a, st) : p_launch<true, 1, false>(tmA, tmB, tmO, tmR, a, st);
    return ctas == 2 ? p_launch<false, 2, false>(tmA, tmB, tmO, tmR, a, st) : p_launch<false, 1, false>(tmA, tmB, tmO, tmR, a, st);
}

}  // namespace anysd
