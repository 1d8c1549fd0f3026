// Blackwell-native contraction: tcgen05.mma (UMMA 128 x BN x 16, fp16 operands, fp32 accumulators in
// TMEM) fed by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier ring.
//
//   dense  : A [M, K] row-major (nn.Linear / 1x1 conv on NHWC tokens), 2-D tensor map, box 64(K) x 128(M)
//   conv3x3: implicit GEMM, stride 1, pad 1, Cin % 64 == 0.  A is the NHWC image itself: a 4-D tensor map
//            (C, W, H, N) with box 64 x BW x BH x NB (BW*BH*NB = 128 output pixels of one rectangular patch);
//            k-block kb = (tap, channel chunk) loads the patch shifted by (kx-1, ky-1) -- the zero halo is
//            TMA out-of-bounds fill, no im2col buffer, no predicates in the main loop.
//   B      : weights [N, K] K-major, box 64(K) x BN(N).
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers via tcgen05.ld 32x32b, fused bias / time-embedding row add /
// SiLU / GEGLU / residual, 16-byte row-contiguous stores).  One output tile per CTA; two CTAs per SM
// (3-stage ring each) so one CTA's epilogue overlaps the other's main loop.
#include <cuda.h>

#include "common.cuh"

namespace anysd {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 3, TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;   // 16 KB

struct Tc5Args {
    const float* bias;
    const float* rowadd;
    const __half* residual;
    void* out;
    int M, N, K;
    int ldo, ldr, ld_rowadd, rows_per_batch;
    int act, out_f16;
    int num_kb;
    // conv geometry
    int Nimg, Ho, Wo, Cin;
    int BW, BH, NB, tiles_w, tiles_h;
    int kb_per_tap;
    int stride;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte swizzle: 8-row groups 1024 B apart (SBO), LBO unused (1), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN>
struct TcCfg {
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = BN <= 128 ? 128 : 256;
    static constexpr int SMEM = TC_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
};

template <int BN, bool CONV>
__global__ void __launch_bounds__(TC_THREADS, 2) gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const Tc5Args p) {
    using Cfg = TcCfg<BN>;
    extern __shared__ unsigned char tc_smem_raw[];
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = tc_smem_raw + (base - raw);
    const uint32_t bar_base = base + TC_STAGES * Cfg::STAGE_BYTES;
    // barriers: full[s] at +8s, empty[s] at +8(STAGES+s), tmem_full at +8*2*STAGES, tmem ptr after
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * TC_STAGES);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + TC_STAGES * Cfg::STAGE_BYTES + 8 * (2 * TC_STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- tile coordinates ----
    int m0 = 0, tn = 0, th = 0, tw = 0;
    if (CONV) {
        tw = m_tile % p.tiles_w;
        th = (m_tile / p.tiles_w) % p.tiles_h;
        tn = m_tile / (p.tiles_w * p.tiles_h);
    } else {
        m0 = m_tile * TC_BM;
    }
    const int n0 = n_tile * BN;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
                const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + TC_A_BYTES;
                if (CONV) {
                    const int tap = kb / p.kb_per_tap;
                    const int c0 = (kb - tap * p.kb_per_tap) * TC_BK;
                    const int ky = tap / 3, kx = tap - ky * 3;
                    tma_load_4d(sa, &tmA, full_bar(s), c0, tw * p.BW * p.stride + kx - 1, th * p.BH * p.stride + ky - 1,
                                tn * p.NB);
                } else {
                    tma_load_2d(sa, &tmA, full_bar(s), kb * TC_BK, m0);
                }
                tma_load_2d(sb, &tmB, full_bar(s), kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t ph = (kb / TC_STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + TC_A_BYTES;
                const uint64_t ad = make_sdesc(sa), bd = make_sdesc(sb);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)   // +32 bytes along K inside the 128B swizzle atom = +2 (16B units)
                    umma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                umma_commit(empty_bar(s));             // frees the smem slot when these MMAs retire
            }
            umma_commit(tmem_full_bar);                // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane group = warp % 4 =====
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        long long m;          // global output row
        int img;              // image index for the row-add
        bool row_ok;
        if (CONV) {
            const int dx = row % p.BW, dy = (row / p.BW) % p.BH, nl = row / (p.BW * p.BH);
            const int ox = tw * p.BW + dx, oy = th * p.BH + dy;
            img = tn * p.NB + nl;
            row_ok = ox < p.Wo && oy < p.Ho && img < p.Nimg;
            m = ((long long)img * p.Ho + oy) * p.Wo + ox;
        } else {
            m = m0 + row;
            row_ok = m < p.M;
            img = row_ok ? (int)(m / p.rows_per_batch) : 0;
        }
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const float* radd = (p.rowadd && row_ok) ? p.rowadd + (size_t)img * p.ld_rowadd : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            __syncwarp();                               // tcgen05.ld is warp-collective (.sync.aligned)
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c * 32), r);
            const int nb = n0 + c * 32;
            if (row_ok && nb < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll
            for (int g = 0; g < 4; ++g) {              // groups of 8 columns (N % 8 == 0)
                const int n = nb + g * 8;
                if (n >= p.N) break;
                float* vv = v + g * 8;
                if (p.bias) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                    vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
                    vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
                }
                if (radd) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(radd + n));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(radd + n + 4));
                    vv[0] += b0.x; vv[1] += b0.y; vv[2] += b0.z; vv[3] += b0.w;
                    vv[4] += b1.x; vv[5] += b1.y; vv[6] += b1.z; vv[7] += b1.w;
                }
                if (p.act == 2) {                      // GEGLU: (a, gate) pairs -> 4 outputs at column n/2
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = vv[2 * j] * gelu_erf_f(vv[2 * j + 1]);
                    const int no = n >> 1;
                    if (p.residual) {
                        const uint2 u = *reinterpret_cast<const uint2*>(p.residual + (size_t)m * p.ldr + no);
                        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
                        const float2 a = __half22float2(h2[0]), b = __half22float2(h2[1]);
                        o[0] += a.x; o[1] += a.y; o[2] += b.x; o[3] += b.y;
                    }
                    if (p.out_f16) {
                        uint2 u;
                        __half2* h2 = reinterpret_cast<__half2*>(&u);
                        h2[0] = __floats2half2_rn(o[0], o[1]);
                        h2[1] = __floats2half2_rn(o[2], o[3]);
                        *reinterpret_cast<uint2*>((__half*)p.out + (size_t)m * p.ldo + no) = u;
                    } else {
                        *reinterpret_cast<float4*>((float*)p.out + (size_t)m * p.ldo + no) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    continue;
                }
                if (p.act != 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) vv[j] = act_f(vv[j], p.act);
                }
                if (p.residual) {
                    const uint4 u = *reinterpret_cast<const uint4*>(p.residual + (size_t)m * p.ldr + n);
                    float rr[8];
                    unpack8(u, rr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) vv[j] += rr[j];
                }
                if (p.out_f16) {
                    *reinterpret_cast<uint4*>((__half*)p.out + (size_t)m * p.ldo + n) = pack8(vv);
                } else {
                    float4* o4 = reinterpret_cast<float4*>((float*)p.out + (size_t)m * p.ldo + n);
                    o4[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    o4[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
                }
            }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ---- host side: tensor maps through the driver entry point (no libcuda link dependency) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
    }
    return fn;
}

static bool encode_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
                      uint32_t box_outer) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// stride s (1|2): the box traverses s*BW x s*BH input pixels with element stride s, i.e. loads BW x BH
// pixels (cuTensorMapEncodeTiled: "to load N elements along a dimension, boxDim = N * elementStrides").
static bool encode_nhwc(CUtensorMap* tm, const void* ptr, int N, int H, int W, int C, int BW, int BH, int NB, int s) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)(BW * s), (cuuint32_t)(BH * s), (cuuint32_t)NB};
    cuuint32_t es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int pick_bn(int N) {
    // exact divisors first (no wasted columns), widest first; then the least padded
    const int cand[3] = {256, 160, 128};
    for (int i = 0; i < 3; ++i)
        if (N % cand[i] == 0) return cand[i];
    int best = 128, waste = 1 << 30;
    for (int i = 0; i < 3; ++i) {
        const int w = cdiv(N, cand[i]) * cand[i] - N;
        if (w < waste) { waste = w; best = cand[i]; }
    }
    return best;
}

// rectangular patch of 128 output pixels (BW x BH x NB, all powers of two): prefer extents that divide
// the image exactly (64->64x2, 32->32x4, 16->16x8, 8->8x8x2, 96->32x4, 48->16x8, 24->8x8x2), otherwise the
// smallest power of two that covers it (the overshoot is TMA out-of-bounds fill + masked rows).
static int pow2_ceil(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
static int pick_extent(int len, int cap) {
    int lim = pow2_ceil(len);
    if (lim > cap) lim = cap;
    for (int c = lim; c >= 4; c >>= 1)
        if (len % c == 0) return c;
    return lim;
}
static void pick_patch(int Ho, int Wo, int* BW, int* BH, int* NB) {
    const int bw = pick_extent(Wo, 128);
    const int bh = pick_extent(Ho, 128 / bw);
    *BW = bw;
    *BH = bh;
    *NB = 128 / (bw * bh);
}

// nearest-neighbour x2 (F.interpolate(scale_factor=2, mode="nearest"), openaimodel.py:110-115) on NHWC fp16
__global__ void upsample2x_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W, int CV, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int ox = (int)(r % (2 * W));
        r /= (2 * W);
        const int oy = (int)(r % (2 * H));
        const long long n = r / (2 * H);
        dst[i] = __ldg(src + ((n * H + (oy >> 1)) * W + (ox >> 1)) * CV + cv);
    }
}

int launch_upsample2x(const __half* src, __half* dst, int N, int H, int W, int C, cudaStream_t st) {
    const long long total = (long long)N * 4 * H * W * (C / 8);
    int grid = (int)((total + 255) / 256);
    if (grid > sm_count() * 16) grid = sm_count() * 16;
    upsample2x_kernel<<<grid, 256, 0, st>>>((const uint4*)src, (uint4*)dst, H, W, C / 8, total);
    return check_launch("upsample2x");
}

bool tc5_supported(const anysd_gemm_params* q) {
    if (q->N % 8 != 0 || q->K % 8 != 0) return false;
    if (q->act == 2 && q->N % 16 != 0) return false;
    if (((uintptr_t)q->out % 16) || (q->residual && ((uintptr_t)q->residual % 16))) return false;
    const int n_out = q->act == 2 ? q->N / 2 : q->N;
    if (q->ldo % 8 != 0 || (q->residual && q->ldr % 8 != 0) || n_out % 4 != 0) return false;
    if (q->bias && ((uintptr_t)q->bias % 16)) return false;
    if (q->rowadd && (((uintptr_t)q->rowadd % 16) || q->ld_rowadd % 4 != 0)) return false;
    if (q->conv) {
        if (q->Cin % TC_BK != 0) return false;
        if (q->upsample) {   // nearest x2 is materialised into the caller's workspace first
            const size_t need = (size_t)q->Nimg * (2 * q->H) * (2 * q->Wd) * q->Cin * sizeof(__half);
            if (q->workspace == nullptr || q->workspace_bytes < need || ((uintptr_t)q->workspace % 16)) return false;
        }
    }
    return get_encode() != nullptr;
}

template <int BN, bool CONV>
static int launch_tc5_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const Tc5Args& a, dim3 grid, cudaStream_t st) {
    using Cfg = TcCfg<BN>;
    static bool done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc5_kernel<BN, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) {
            set_error("tcgen05 gemm: smem opt-in failed: %s", cudaGetErrorString(e));
            return ANYSD_ECUDA;
        }
        done[dev] = true;
    }
    gemm_tc5_kernel<BN, CONV><<<grid, TC_THREADS, Cfg::SMEM, st>>>(tmA, tmB, a);
    return check_launch(CONV ? "conv3x3 (tcgen05)" : "gemm (tcgen05)");
}

int launch_gemm_tc5(const anysd_gemm_params* q, cudaStream_t st) {
    Tc5Args a;
    a.bias = q->bias; a.rowadd = q->rowadd; a.residual = (const __half*)q->residual; a.out = q->out;
    a.M = q->M; a.N = q->N; a.K = q->K;
    a.ldo = q->ldo; a.ldr = q->ldr; a.ld_rowadd = q->ld_rowadd;
    a.rows_per_batch = q->rows_per_batch > 0 ? q->rows_per_batch : 1;
    a.act = q->act; a.out_f16 = q->out_dtype == ANYSD_F16;
    a.num_kb = cdiv(q->K, TC_BK);
    a.Nimg = q->Nimg; a.Cin = q->Cin;
    a.BW = a.BH = a.NB = a.tiles_w = a.tiles_h = 1;
    a.kb_per_tap = 1;
    a.stride = q->conv ? q->stride : 1;
    // conv input as the tensor map sees it (after the optional materialised nearest-x2 upsample)
    const void* img = q->A;
    int Hin = q->H, Win = q->Wd;
    if (q->conv && q->upsample) {
        int rc = launch_upsample2x((const __half*)q->A, (__half*)q->workspace, q->Nimg, q->H, q->Wd, q->Cin, st);
        if (rc) return rc;
        img = q->workspace;
        Hin = 2 * q->H;
        Win = 2 * q->Wd;
    }
    a.Ho = (Hin - 1) / a.stride + 1;
    a.Wo = (Win - 1) / a.stride + 1;
    const int BN = pick_bn(q->N);
    CUtensorMap tmA, tmB;
    if (!encode_2d(&tmB, q->W, (uint64_t)q->K, (uint64_t)q->N, (uint64_t)q->ldw, TC_BK, BN)) {
        set_error("tcgen05 gemm: cuTensorMapEncodeTiled failed for W (N=%d K=%d ldw=%d)", q->N, q->K, q->ldw);
        return ANYSD_ECUDA;
    }
    dim3 grid;
    if (q->conv) {
        pick_patch(a.Ho, a.Wo, &a.BW, &a.BH, &a.NB);
        a.tiles_w = cdiv(a.Wo, a.BW);
        a.tiles_h = cdiv(a.Ho, a.BH);
        const int tiles_n = cdiv(q->Nimg, a.NB);
        a.kb_per_tap = q->Cin / TC_BK;
        a.num_kb = 9 * a.kb_per_tap;
        if (!encode_nhwc(&tmA, img, q->Nimg, Hin, Win, q->Cin, a.BW, a.BH, a.NB, a.stride)) {
            set_error("tcgen05 conv: cuTensorMapEncodeTiled failed for x (N=%d H=%d W=%d C=%d box %dx%dx%d)", q->Nimg, q->H,
                      q->Wd, q->Cin, a.BW, a.BH, a.NB);
            return ANYSD_ECUDA;
        }
        grid = dim3(cdiv(q->N, BN), a.tiles_w * a.tiles_h * tiles_n);
    } else {
        if (!encode_2d(&tmA, q->A, (uint64_t)q->K, (uint64_t)q->M, (uint64_t)q->lda, TC_BK, TC_BM)) {
            set_error("tcgen05 gemm: cuTensorMapEncodeTiled failed for A (M=%d K=%d lda=%d)", q->M, q->K, q->lda);
            return ANYSD_ECUDA;
        }
        grid = dim3(cdiv(q->N, BN), cdiv(q->M, TC_BM));
    }
    if (grid.y > 65535) {
        set_error("tcgen05 gemm: too many M tiles (%u)", grid.y);
        return ANYSD_EINVAL;
    }
    if (q->conv) {
        switch (BN) {
            case 256: return launch_tc5_t<256, true>(tmA, tmB, a, grid, st);
            case 160: return launch_tc5_t<160, true>(tmA, tmB, a, grid, st);
            default: return launch_tc5_t<128, true>(tmA, tmB, a, grid, st);
        }
    }
    switch (BN) {
        case 256: return launch_tc5_t<256, false>(tmA, tmB, a, grid, st);
        case 160: return launch_tc5_t<160, false>(tmA, tmB, a, grid, st);
        default: return launch_tc5_t<128, false>(tmA, tmB, a, grid, st);
    }
}

}  // namespace anysd
