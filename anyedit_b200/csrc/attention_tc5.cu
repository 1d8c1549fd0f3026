// Blackwell-native attention: S = Q K^T and O += P V on tcgen05 (accumulators in TMEM), operands staged by
// TMA, softmax in fp32 registers with ONE thread per query row (no shuffles), lazy accumulator rescale.
// Replaces the n x n score tensor of the reference's eager CrossAttention (ldm/modules/attention.py:171-193).
//
//   CTA = 128 query rows of one (batch, head).  192 threads: warp 0 TMA producer, warp 1 MMA issuer,
//   warps 2..5 softmax/correction/epilogue (TMEM lane group = warp % 4, lane = row).
//   TMEM: S [128 x 128] fp32 at columns 0..127, O [128 x d_ext] fp32 at columns 128...
//   smem: Q (NA atoms of 128 x 64 fp16, 128B swizzle), 2-stage ring of (K, V) tiles, P [128 x 128] fp16.
//   Pipeline per kv tile j:   MMA: S(j+1) = Q K(j+1)^T is issued as soon as the softmax warps have pulled
//   S(j) into registers, so the tensor pipe works under the exp; then O += P(j) V(j).
//   Softmax: p = exp2(s*scale*log2e - m) with m only raised when the tile max exceeds it by > 8 (p <= 256,
//   safe in fp16; any reference max gives the same O / l), so the TMEM read-modify-write rescale of O is rare.
//   d = 40 needs the head stride padded to >= 48 with zero columns (anyedit_b200.unet packs q/k/v that way);
//   K-extent = ceil16(d), V is consumed as an MN-major B operand straight from its row-major tile.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "attention_tc5.cuh"

namespace anysd {

constexpr int AT_BQ = 128, AT_THREADS = 192, AT_ATOM = 128 * 128;   // 16 KB atom: 128 rows x 64 halves

struct AtArgs {
    __half* out;
    long long obs;
    int ldo;
    int n_q, n_kv, d, d_ext, hs;       // hs = head stride (elements) inside q/k/v rows
    int NA, stages, tmem_cols;
    float scale_log2;
    const float* gate;
    int gate_stride, accumulate;
    float* lse;                        // optional [B, heads, n_q]: base-2 log-sum-exp of every score row
};

// BKV = kv rows per tile.  128: 2 CTAs/SM (TMEM 256 columns each).  64: S + O fit 128 TMEM columns, ~100 registers
// and 56 KB smem per CTA -> 3 CTAs/SM = 12 softmax warps per SM, which is what the latency-bound exp/max/pack
// stream of the d = 40, 4096-token level needs (ncu: XU 64 %, issue 46 %, top stall "wait").
// AUX (needs head stride >= d + 2 and the operand contract of anysd_attn_params::aux_cols): the softmax reference
// and the denominator move INTO the tensor-core products.  Q carries (-m_hi, -m_lo) in padding columns d, d+1 and K
// carries 1.0 there, so S arrives as  q.k*scale*log2e - m  and the CUDA cores only take exp2 -- no FFMA; V carries
// 1.0 in padding column d, so column d of P.V is the softmax denominator, accumulated in fp32 from exactly the fp16 P
// the numerator uses -- no FADD.  Per pair of scores: 2 MUFU + 1 pack + max tracking instead of 8.5 instructions
// (the kernel is issue-bound in the softmax warps: every variant that ADDED instructions got slower).  The reference
// only moves when a tile exceeds it by > 8; the thread then rewrites its two Q columns in smem before it releases
// S(j), i.e. before the MMA warp may issue S(j+1).
template <int BKV, bool AUX>
__global__ void __launch_bounds__(AT_THREADS, BKV == 64 ? 3 : 2)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AtArgs p) {
    // 128B-swizzled tiles need a 1024-byte aligned base; the dynamic smem window of a kernel without static
    // smem starts 1024-aligned (declared so below; checked, not padded: padding would cost the 2nd CTA per SM).
    extern __shared__ __align__(1024) unsigned char at_smem_raw[];
    const uint32_t base = smem_u32(at_smem_raw);
    if (base & 1023u) __trap();
    unsigned char* smem = at_smem_raw;
    const int NA = p.NA, ST = p.stages;
    constexpr int KV_ATOM = BKV * 128;                          // bytes of one [BKV x 64] K/V atom
    constexpr int AT_BKV = BKV;
    const uint32_t q_off = 0;
    const uint32_t kv_off = NA * AT_ATOM;                       // stage s: K at kv_off + s*2*NA*KV_ATOM, V right after K
    const uint32_t p_off = kv_off + ST * 2 * NA * KV_ATOM;      // P: BKV/64 atoms of [128 x 64]
    const uint32_t bar_off = p_off + (BKV / 64) * AT_ATOM;
    const uint32_t bars = base + bar_off;
    // barriers: 0 q_full | 1,2 kv_full | 3,4 kv_empty | 5 s_full | 6 s_free | 7 p_full | 8 o_done
    auto BAR = [&](int i) { return bars + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AT_BQ;
    const int nt = (p.n_kv + AT_BKV - 1) / AT_BKV;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    }
    if (warp == 1 && lane == 0) {
        am_init(BAR(0), 1);
        am_init(BAR(1), 1); am_init(BAR(2), 1);
        am_init(BAR(3), 1); am_init(BAR(4), 1);
        am_init(BAR(5), 1);
        am_init(BAR(6), 4);      // one arrive per softmax warp
        am_init(BAR(7), 4);
        am_init(BAR(8), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    a_fence_before();
    __syncthreads();
    a_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_S = tmem, tmem_O = tmem + BKV;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const int col0 = h * p.hs;
            am_expect_tx(BAR(0), NA * AT_ATOM);
            for (int a = 0; a < NA; ++a) a_tma_2d(base + q_off + a * AT_ATOM, &tmQ, BAR(0), col0 + a * 64, b * p.n_q + q0);
            for (int j = 0; j < nt; ++j) {
                const int s = j % ST;
                am_wait_relaxed(BAR(3 + s), (((uint32_t)j / ST) & 1) ^ 1);
                am_expect_tx(BAR(1 + s), 2 * NA * KV_ATOM);
                const uint32_t kb = base + kv_off + s * 2 * NA * KV_ATOM, vb = kb + NA * KV_ATOM;
                for (int a = 0; a < NA; ++a) a_tma_2d(kb + a * KV_ATOM, &tmK, BAR(1 + s), col0 + a * 64, b * p.n_kv + j * AT_BKV);
                for (int a = 0; a < NA; ++a) a_tma_2d(vb + a * KV_ATOM, &tmV, BAR(1 + s), col0 + a * 64, b * p.n_kv + j * AT_BKV);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc_s = (1u << 4) | ((uint32_t)(AT_BKV >> 3) << 17) | ((uint32_t)(AT_BQ >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (1u << 16) /*B is MN-major*/ | ((uint32_t)(p.d_ext >> 3) << 17) |
                                     ((uint32_t)(AT_BQ >> 4) << 24);
            const int ksteps = p.d_ext / 16;
            auto issue_S = [&](int j) {
                const uint32_t kb = base + kv_off + (j % ST) * 2 * NA * KV_ATOM;
                for (int k = 0; k < ksteps; ++k) {
                    const uint32_t qo = (k >> 2) * AT_ATOM + (k & 3) * 32, ko = (k >> 2) * KV_ATOM + (k & 3) * 32;
                    a_umma(tmem_S, a_desc_k(base + q_off + qo), a_desc_k(kb + ko), idesc_s, k ? 1u : 0u);
                }
                a_commit(BAR(5));
            };
            am_wait(BAR(0), 0);
            am_wait(BAR(1), 0);
            a_fence_after();
            issue_S(0);
            for (int j = 0; j < nt; ++j) {
                if (j + 1 < nt && ST > 1) {                               // run S(j+1) under the softmax of tile j
                    const int s1 = (j + 1) % ST;
                    am_wait(BAR(1 + s1), ((uint32_t)(j + 1) / ST) & 1);   // K(j+1), V(j+1) landed
                    am_wait_relaxed(BAR(6), j & 1);                       // S(j) is in registers
                    a_fence_after();
                    issue_S(j + 1);
                }
                am_wait_relaxed(BAR(7), j & 1);                           // P(j) written (and O rescaled)
                a_fence_after();
                const uint32_t vb = base + kv_off + (j % ST) * 2 * NA * KV_ATOM + NA * KV_ATOM;
                for (int k = 0; k < AT_BKV / 16; ++k) {
                    const uint32_t poff = (k >> 2) * AT_ATOM + (k & 3) * 32;
                    a_umma(tmem_O, a_desc_k(base + p_off + poff), a_desc_mn(vb + k * 2048, KV_ATOM), idesc_o, (j | k) ? 1u : 0u);
                }
                a_commit(BAR(8));                 // O(j) accumulated, P free
                a_commit(BAR(3 + (j % ST)));      // K/V stage free
                if (j + 1 < nt && ST == 1) {      // single K/V stage (d = 160): the next tile can only land now
                    am_wait(BAR(1), (uint32_t)(j + 1) & 1);
                    am_wait(BAR(6), j & 1);
                    a_fence_after();
                    issue_S(j + 1);
                }
            }
        }
    } else {
        // ===== softmax / correction / epilogue: 4 warps, one thread per query row =====
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        float m_run = AUX ? 0.f : -INFINITY, l_run = 0.f;
        unsigned char* prow = smem + p_off + row * 128;
        for (int j = 0; j < nt; ++j) {
            am_wait(BAR(5), j & 1);
            a_fence_after();
            const int kv_left = p.n_kv - j * AT_BKV;
            bool upd = false;
            float corr = 1.0f;
            uint32_t pk[BKV / 2];
            bool done = false;
            if (AUX && kv_left >= AT_BKV) {
                // Streamed tile: with the reference inside the product the exponentials do not wait for the row
                // maximum, so chunk c+1 (32 columns) travels TMEM -> registers while chunk c goes through the XU.
                // The tile maximum is complete before the LAST chunk's exponentials; if no row of the warp moves its
                // reference (the common case) S(j) is released there and S(j+1) runs under the remaining work.
                // Otherwise nothing is released and the tile is redone from TMEM on the general path below.
                constexpr int NC = BKV / 32;
                uint32_t cb[2][32];
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                __syncwarp();
                a_ld32(tmem_S + lane_addr, cb[0]);
                a_wait_ld();
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (c + 1 < NC) a_ld32(tmem_S + lane_addr + (c + 1) * 32, cb[(c + 1) & 1]);
                    const uint32_t* cur = cb[c & 1];
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(cur[i + 4]), __uint_as_float(cur[i + 5])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(cur[i + 6]), __uint_as_float(cur[i + 7])));
                    }
                    if (c == NC - 1) {
                        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        if (!__any_sync(0xffffffffu, mx > 8.0f)) {
                            a_fence_before();
                            __syncwarp();
                            if (lane == 0) am_arrive(BAR(6));     // S(j) consumed, reference unchanged
                            done = true;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        __half2 hh = __floats2half2_rn(a_ex2(__uint_as_float(cur[2 * i])), a_ex2(__uint_as_float(cur[2 * i + 1])));
                        pk[c * 16 + i] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    if (c + 1 < NC) a_wait_ld();
                }
            }
            if (!done) {
                uint32_t sr[BKV];
                __syncwarp();
    #pragma unroll
                for (int c = 0; c < BKV / 32; ++c) a_ld32(tmem_S + lane_addr + c * 32, sr + c * 32);
                a_wait_ld();
                if (!AUX) {
                    a_fence_before();
                    __syncwarp();
                    if (lane == 0) am_arrive(BAR(6));             // S(j) consumed: the MMA warp may overwrite it
                }
                float mx = -INFINITY;
                if (kv_left >= AT_BKV) {
                    // four independent chains (a single running max is a 128-deep dependent chain)
                    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
    #pragma unroll
                    for (int i = 0; i < BKV; i += 8) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7])));
                    }
                    mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                } else {
    #pragma unroll
                    for (int i = 0; i < BKV; ++i) {
                        float v = (i < kv_left) ? __uint_as_float(sr[i]) : -INFINITY;
                        sr[i] = __float_as_uint(v);
                        mx = fmaxf(mx, v);
                    }
                }
                if (AUX) {
                    // S is already relative to the reference held in the Q columns: mx is the excess over it
                    upd = mx > 8.0f;
                    corr = upd ? a_ex2(-mx) : 1.0f;
                    if (upd) {
                        m_run += mx;
                        const __half mh = __float2half_rn(-m_run);
                        const __half ml = __float2half_rn(-m_run - __half2float(mh));
                        __half2 pair = __halves2half2(mh, ml);
                        unsigned char* qrow = smem + q_off + row * 128 + ((((p.d >> 3) & 7) ^ (row & 7)) << 4) + (p.d >> 6) * AT_ATOM;
                        *reinterpret_cast<__half2*>(qrow) = pair;        // columns d, d+1 of this row (d % 8 == 0)
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    a_fence_before();
                    __syncwarp();
                    if (lane == 0) am_arrive(BAR(6));             // S(j) consumed AND the Q reference is in place
                    if (__any_sync(0xffffffffu, upd)) {
                        const float sub = upd ? mx : 0.f;
    #pragma unroll
                        for (int i = 0; i < BKV / 2; ++i) {
                            __half2 hh = __floats2half2_rn(a_ex2(__uint_as_float(sr[2 * i]) - sub), a_ex2(__uint_as_float(sr[2 * i + 1]) - sub));
                            pk[i] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                    } else {
    #pragma unroll
                        for (int i = 0; i < BKV / 2; ++i) {
                            __half2 hh = __floats2half2_rn(a_ex2(__uint_as_float(sr[2 * i])), a_ex2(__uint_as_float(sr[2 * i + 1])));
                            pk[i] = *reinterpret_cast<uint32_t*>(&hh);
                        }
                    }
                } else {
                    mx *= p.scale_log2;
                    // lazy max: only move the reference when the tile max exceeds it by more than 8 (2^8 headroom)
                    upd = mx > m_run + 8.0f;
                    const float m_new = upd ? mx : m_run;
                    corr = upd ? a_ex2(m_run - m_new) : 1.0f;     // first tile: ex2(-inf) = 0
                    m_run = m_new;
                    float sum = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
    #pragma unroll
                    for (int i = 0; i < BKV / 2; ++i) {
                        const float p0 = a_ex2(fmaf(__uint_as_float(sr[2 * i]), p.scale_log2, -m_new));
                        const float p1 = a_ex2(fmaf(__uint_as_float(sr[2 * i + 1]), p.scale_log2, -m_new));
                        if ((i & 3) == 0) sum += p0 + p1;          // four independent accumulation chains
                        else if ((i & 3) == 1) sum1 += p0 + p1;
                        else if ((i & 3) == 2) sum2 += p0 + p1;
                        else sum3 += p0 + p1;
                        __half2 hh = __floats2half2_rn(p0, p1);
                        pk[i] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    l_run = l_run * corr + ((sum + sum1) + (sum2 + sum3));
                }
            }
            // P smem and the O accumulator are free once PV(j-1) has retired
            if (j > 0) {
                am_wait(BAR(8), (j - 1) & 1);
                a_fence_after();
            }
            const bool need = __any_sync(0xffffffffu, upd) && j > 0;
            if (need) {
                for (int c = 0; c < p.d_ext; c += 16) {
                    uint32_t o[16];
                    a_ld16(tmem_O + lane_addr + c, o);
                    a_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                    a_st16(tmem_O + lane_addr + c, o);
                }
                a_wait_st();
            }
#pragma unroll
            for (int c = 0; c < BKV / 8; ++c) {               // chunks of 8 halves; chunk c lives in atom c/8
                uint4 u = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                *reinterpret_cast<uint4*>(prow + (c >> 3) * AT_ATOM + (((c & 7) ^ (row & 7)) << 4)) = u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            a_fence_before();
            __syncwarp();
            if (lane == 0) am_arrive(BAR(7));
        }
        // ---- epilogue: O / l -> fp16 -> global ----
        am_wait(BAR(8), (nt - 1) & 1);
        a_fence_after();
        const float g = p.gate ? p.gate[(size_t)b * p.gate_stride] : 1.0f;
        if (AUX) {                                         // denominator = column d of the accumulator
            uint32_t o[16];
            __syncwarp();
            a_ld16(tmem_O + lane_addr + (p.d & ~15), o);
            a_wait_ld();
            l_run = (p.d & 8) ? __uint_as_float(o[8]) : __uint_as_float(o[0]);
        }
        const float inv = g / l_run;
        const int qr = q0 + row;
        if (p.lse != nullptr && qr < p.n_q) p.lse[((size_t)b * gridDim.y + h) * p.n_q + qr] = m_run + log2f(l_run);
        __half* orow = p.out + (size_t)b * p.obs + (size_t)qr * p.ldo + (size_t)h * p.d;
        for (int c = 0; c < p.d_ext; c += 16) {
            uint32_t o[16];
            __syncwarp();
            a_ld16(tmem_O + lane_addr + c, o);
            a_wait_ld();
            if (qr < p.n_q) {
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                    const int col = c + g8 * 8;
                    if (col >= p.d) break;
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g8 * 8 + i]) * inv;
                    uint4* dst = reinterpret_cast<uint4*>(orow + col);
                    if (p.accumulate) {
                        float prev[8];
                        unpack8(*dst, prev);
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] += prev[i];
                    }
                    *dst = pack8(f);
                }
            }
        }
    }
    a_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ===== paired variant: TWO query tiles per CTA, one CTA per SM =====================================================
// Same pipeline as above for each of two 128-row query tiles that share every K/V tile (half the K/V traffic and TMA work, a
// 3-stage K/V ring instead of 2), 8 softmax warps per CTA.  AUX operand contract, d_ext <= 64, n_q % 256 == 0, n_kv % 128 == 0
// (the d = 40 / 4096-token level this is for: [measured] 744-751 us against 775-795 us for two independent CTAs per SM);
// everything else runs the kernel above.
//   320 threads: warp 0 TMA, warp 1 MMA issuer (S0(j+1), P0 V(j), S1(j+1), P1 V(j)), warps 2..5 tile 0, warps 6..9 tile 1.
//   TMEM (512 columns): S0 0..127 | S1 128..255 | O0 256.. | O1 320..      smem: Q0 Q1 | 3 x (K, V) | P0 P1
// What was tried on top of this and did not pay (round 2; logs under profiles/r2_attn_*): a half-tile stagger of
// the two tiles enforced by sequence barriers; P handed to P.V through tensor memory (tcgen05.st + A operand from TMEM); 64-key
// score tiles with S and P double-buffered so that no softmax warp ever waits for the tensor core; one MMA issuer warp per tile.
// The per-phase clock counts those variants were instrumented with say where the time is: a tcgen05.mma of this size
// (128 x 128 x 16, 128 x 48 x 16) occupies its issuing thread for ~90-110 clk whatever its 24-64 clk of math, so the 11
// instructions + 3 commits per (tile, kv tile) are ~1100 clk of a ~3000 clk iteration, and the exponential phase of one tile
// slows down by as much as the other tile's products are moved under it.
constexpr int AW_THREADS = 320, AW_ST = 3, AW_BKV = 128;

__global__ void __launch_bounds__(AW_THREADS, 1)
attention_tc5x2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AtArgs p) {
    extern __shared__ __align__(1024) unsigned char aw_smem_raw[];
    const uint32_t base = smem_u32(aw_smem_raw);
    if (base & 1023u) __trap();
    unsigned char* smem = aw_smem_raw;
    constexpr int KV_ATOM = AW_BKV * 128;                       // 16 KB: [128 rows x 64 halves]
    constexpr uint32_t q_off = 0;                               // Q tile x at q_off + x * AT_ATOM
    constexpr uint32_t kv_off = 2 * AT_ATOM;                    // stage s: K at kv_off + s * 2 * KV_ATOM, V right after
    constexpr uint32_t p_off = kv_off + AW_ST * 2 * KV_ATOM;    // P tile x: two atoms at p_off + x * 2 * AT_ATOM
    constexpr uint32_t bar_off = p_off + 4 * AT_ATOM;
    const uint32_t bars = base + bar_off;
    // barriers: 0 q_full | 1..3 kv_full | 4..6 kv_empty | 7,8 s_full | 9,10 s_free | 11,12 p_full | 13,14 o_done
    auto BAR = [&](int i) { return bars + 8u * i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 8 * 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 2 * AT_BQ;
    const int nt = p.n_kv / AW_BKV;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i <= 8; ++i) am_init(BAR(i), 1);        // q_full, kv_full, kv_empty, s_full
        for (int i = 9; i <= 12; ++i) am_init(BAR(i), 4);       // s_free, p_full: one arrive per softmax warp of the tile
        am_init(BAR(13), 1); am_init(BAR(14), 1);               // o_done
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    a_fence_before();
    __syncthreads();
    a_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            const int col0 = h * p.hs;
            am_expect_tx(BAR(0), 2 * AT_ATOM);
            a_tma_2d(base + q_off, &tmQ, BAR(0), col0, b * p.n_q + q0);
            a_tma_2d(base + q_off + AT_ATOM, &tmQ, BAR(0), col0, b * p.n_q + q0 + AT_BQ);
            for (int j = 0; j < nt; ++j) {
                const int s = j % AW_ST;
                am_wait_relaxed(BAR(4 + s), (((uint32_t)j / AW_ST) & 1) ^ 1);
                am_expect_tx(BAR(1 + s), 2 * KV_ATOM);
                const uint32_t kb = base + kv_off + s * 2 * KV_ATOM;
                a_tma_2d(kb, &tmK, BAR(1 + s), col0, b * p.n_kv + j * AW_BKV);
                a_tma_2d(kb + KV_ATOM, &tmV, BAR(1 + s), col0, b * p.n_kv + j * AW_BKV);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: S0(j+1), P0 V(j), S1(j+1), P1 V(j).  Descriptors are built once; a k-step is + 32 bytes. =====
            const uint32_t idesc_s = (1u << 4) | ((uint32_t)(AW_BKV >> 3) << 17) | ((uint32_t)(AT_BQ >> 4) << 24);
            const uint32_t idesc_o = (1u << 4) | (1u << 16) /*B is MN-major*/ | ((uint32_t)(p.d_ext >> 3) << 17) |
                                     ((uint32_t)(AT_BQ >> 4) << 24);
            const int ksteps = p.d_ext / 16;
            const uint64_t qd = a_desc_k(base + q_off), pd = a_desc_k(base + p_off);
            auto issue_S = [&](int x, int j) {
                const uint64_t kd = a_desc_k(base + kv_off + (j % AW_ST) * 2 * KV_ATOM);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < ksteps) a_umma(tmem + x * AW_BKV, qd + x * (AT_ATOM >> 4) + 2 * k, kd + 2 * k, idesc_s, k ? 1u : 0u);
                a_commit(BAR(7 + x));
            };
            am_wait(BAR(0), 0);
            am_wait(BAR(1), 0);
            a_fence_after();
            issue_S(0, 0);
            issue_S(1, 0);
            for (int j = 0; j < nt; ++j) {
#pragma unroll
                for (int x = 0; x < 2; ++x) {
                    if (j + 1 < nt) {
                        if (x == 0) am_wait(BAR(1 + (j + 1) % AW_ST), ((uint32_t)(j + 1) / AW_ST) & 1);   // K(j+1), V(j+1) landed
                        am_wait_relaxed(BAR(9 + x), j & 1);                                               // S_x(j) is in registers
                        a_fence_after();
                        issue_S(x, j + 1);
                    }
                    am_wait_relaxed(BAR(11 + x), j & 1);                                                  // P_x(j) written
                    a_fence_after();
                    const uint64_t vd = a_desc_mn(base + kv_off + (j % AW_ST) * 2 * KV_ATOM + KV_ATOM, KV_ATOM);
#pragma unroll
                    for (int k = 0; k < AW_BKV / 16; ++k)                     // P atom k / 4, + 32 bytes per k-step; V + 2048 bytes per 16 keys
                        a_umma(tmem + 256 + x * 64, pd + (x * 2 + (k >> 2)) * (AT_ATOM >> 4) + (k & 3) * 2, vd + 128 * k, idesc_o, (j | k) ? 1u : 0u);
                    a_commit(BAR(13 + x));                       // O_x(j) accumulated, P_x free
                    if (x == 1) a_commit(BAR(4 + j % AW_ST));    // both tiles are through with this K/V stage
                }
            }
        }
    } else {
        // ===== softmax / correction / epilogue: 2 x 4 warps, one thread per query row =====
        const int x = (warp - 2) >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const uint32_t tmem_S = tmem + x * AW_BKV, tmem_O = tmem + 256 + x * 64;
        float m_run = 0.f, l_run = 0.f;
        unsigned char* prow = smem + p_off + x * 2 * AT_ATOM + row * 128;
        constexpr int NC = AW_BKV / 32;
        for (int j = 0; j < nt; ++j) {
            am_wait(BAR(7 + x), j & 1);
            a_fence_after();
            bool upd = false;
            float corr = 1.0f;
            uint32_t pk[AW_BKV / 2];
            bool done = false;
            {
                // streamed tile (see the kernel above): chunk c + 1 travels TMEM -> registers under the exponentials of chunk c
                uint32_t cb[2][32];
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                __syncwarp();
                a_ld32(tmem_S + lane_addr, cb[0]);
                a_wait_ld();
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (c + 1 < NC) a_ld32(tmem_S + lane_addr + (c + 1) * 32, cb[(c + 1) & 1]);
                    const uint32_t* cur = cb[c & 1];
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])));
                        m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[i + 2]), __uint_as_float(cur[i + 3])));
                        m2 = fmaxf(m2, fmaxf(__uint_as_float(cur[i + 4]), __uint_as_float(cur[i + 5])));
                        m3 = fmaxf(m3, fmaxf(__uint_as_float(cur[i + 6]), __uint_as_float(cur[i + 7])));
                    }
                    if (c == NC - 1) {
                        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        if (!__any_sync(0xffffffffu, mx > 8.0f)) {
                            a_fence_before();
                            __syncwarp();
                            if (lane == 0) am_arrive(BAR(9 + x));         // S(j) consumed, reference unchanged
                            done = true;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        __half2 hh = __floats2half2_rn(a_ex2(__uint_as_float(cur[2 * i])), a_ex2(__uint_as_float(cur[2 * i + 1])));
                        pk[c * 16 + i] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    if (c + 1 < NC) a_wait_ld();
                }
            }
            if (!done) {
                // some row of the warp moves its reference: redo the tile from TMEM (rare: the first tiles of a row)
                uint32_t sr[AW_BKV];
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; ++c) a_ld32(tmem_S + lane_addr + c * 32, sr + c * 32);
                a_wait_ld();
                float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < AW_BKV; i += 8) {
                    m0 = fmaxf(m0, fmaxf(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])));
                    m1 = fmaxf(m1, fmaxf(__uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3])));
                    m2 = fmaxf(m2, fmaxf(__uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5])));
                    m3 = fmaxf(m3, fmaxf(__uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7])));
                }
                const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                upd = mx > 8.0f;
                corr = upd ? a_ex2(-mx) : 1.0f;
                if (upd) {
                    m_run += mx;
                    const __half mh = __float2half_rn(-m_run);
                    const __half ml = __float2half_rn(-m_run - __half2float(mh));
                    __half2 pair = __halves2half2(mh, ml);
                    unsigned char* qrow = smem + q_off + x * AT_ATOM + row * 128 + ((((p.d >> 3) & 7) ^ (row & 7)) << 4);
                    *reinterpret_cast<__half2*>(qrow) = pair;            // columns d, d+1 of this row (d % 8 == 0)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                a_fence_before();
                __syncwarp();
                if (lane == 0) am_arrive(BAR(9 + x));                     // S(j) consumed AND the Q reference is in place
                const float sub = upd ? mx : 0.f;
#pragma unroll
                for (int i = 0; i < AW_BKV / 2; ++i) {
                    __half2 hh = __floats2half2_rn(a_ex2(__uint_as_float(sr[2 * i]) - sub), a_ex2(__uint_as_float(sr[2 * i + 1]) - sub));
                    pk[i] = *reinterpret_cast<uint32_t*>(&hh);
                }
            }
            // P smem and the O accumulator are free once PV(j-1) has retired
            if (j > 0) {
                am_wait(BAR(13 + x), (j - 1) & 1);
                a_fence_after();
            }
            const bool need = __any_sync(0xffffffffu, upd) && j > 0;
            if (need) {
                for (int c = 0; c < p.d_ext; c += 16) {
                    uint32_t o[16];
                    a_ld16(tmem_O + lane_addr + c, o);
                    a_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
                    a_st16(tmem_O + lane_addr + c, o);
                }
                a_wait_st();
            }
#pragma unroll
            for (int c = 0; c < AW_BKV / 8; ++c) {            // chunks of 8 halves; chunk c lives in atom c/8
                uint4 u = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                *reinterpret_cast<uint4*>(prow + (c >> 3) * AT_ATOM + (((c & 7) ^ (row & 7)) << 4)) = u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            a_fence_before();
            __syncwarp();
            if (lane == 0) am_arrive(BAR(11 + x));
        }
        // ---- epilogue: O / l -> fp16 -> global (the denominator is column d of the accumulator) ----
        am_wait(BAR(13 + x), (nt - 1) & 1);
        a_fence_after();
        const float g = p.gate ? p.gate[(size_t)b * p.gate_stride] : 1.0f;
        {
            uint32_t o[16];
            __syncwarp();
            a_ld16(tmem_O + lane_addr + (p.d & ~15), o);
            a_wait_ld();
            l_run = (p.d & 8) ? __uint_as_float(o[8]) : __uint_as_float(o[0]);
        }
        const float inv = g / l_run;
        const int qr = q0 + x * AT_BQ + row;
        if (p.lse != nullptr) p.lse[((size_t)b * gridDim.y + h) * p.n_q + qr] = m_run + log2f(l_run);
        __half* orow = p.out + (size_t)b * p.obs + (size_t)qr * p.ldo + (size_t)h * p.d;
        for (int c = 0; c < p.d_ext; c += 16) {
            uint32_t o[16];
            __syncwarp();
            a_ld16(tmem_O + lane_addr + c, o);
            a_wait_ld();
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
                const int col = c + g8 * 8;
                if (col >= p.d) break;
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[g8 * 8 + i]) * inv;
                uint4* dst = reinterpret_cast<uint4*>(orow + col);
                if (p.accumulate) {
                    float prev[8];
                    unpack8(*dst, prev);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] += prev[i];
                }
                *dst = pack8(f);
            }
        }
    }
    a_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- host ---------------------------------------------------------------------------------------------
bool attention_tc5_supported(const anysd_attn_params* q) {
    const int hs = q->head_stride > 0 ? q->head_stride : q->d;
    const int d_ext = (q->d + 15) / 16 * 16;
    if (d_ext > hs && q->heads > 1) return false;              // K-extent would reach into the next head's columns
    if (q->d % 8 != 0 || d_ext > 160) return false;
    if (q->aux_cols && (q->d % 16 != 8 || hs < q->d + 8)) return false;
    if (q->ld_q % 8 || q->ld_k % 8 || q->ld_v % 8 || q->ld_o % 8) return false;
    if (((uintptr_t)q->q % 16) || ((uintptr_t)q->k % 16) || ((uintptr_t)q->v % 16) || ((uintptr_t)q->out % 16)) return false;
    // batches must be stacked rows of one matrix (what the UNet produces): batch stride = n * ld
    if (q->q_batch_stride != (long long)q->n_q * q->ld_q || q->k_batch_stride != (long long)q->n_kv * q->ld_k ||
        q->v_batch_stride != (long long)q->n_kv * q->ld_v)
        return false;
    if ((q->o_batch_stride % 8) != 0) return false;
    return a_get_encode() != nullptr;
}

int launch_attention_tc5(const anysd_attn_params* q, cudaStream_t st) {
    AtArgs a;
    const int hs = q->head_stride > 0 ? q->head_stride : q->d;
    a.out = (__half*)q->out;
    a.obs = q->o_batch_stride;
    a.ldo = q->ld_o;
    a.n_q = q->n_q; a.n_kv = q->n_kv; a.d = q->d; a.hs = hs;
    a.d_ext = (q->d + 15) / 16 * 16;
    a.NA = (a.d_ext + 63) / 64;
    // kv tile [measured, B200]: 128 rows for the long small-head case (d = 40, 4096 tokens: 835 us vs 942 us),
    // 64 rows elsewhere (d = 80 @1024: 121 vs 133 us; 77-token cross attention: 70 vs 83 us)
    static const char* bkv_env = getenv("ANYSD_ATTN_BKV");
    int bkv = (a.d_ext <= 64 && q->n_kv >= 1024) ? 128 : 64;
    if (bkv_env) bkv = atoi(bkv_env) == 64 ? 64 : 128;
    a.stages = (bkv == 64 || a.NA <= 2) ? 2 : 1;
    int need_cols = bkv + a.d_ext;
    a.tmem_cols = 32;
    while (a.tmem_cols < need_cols) a.tmem_cols <<= 1;
    a.scale_log2 = q->scale * 1.4426950408889634f;
    a.gate = q->gate; a.gate_stride = q->gate_stride; a.accumulate = q->accumulate;
    a.lse = q->lse;
    CUtensorMap tmQ, tmK, tmV;
    // valid columns from the slice pointer (aux_cols: the last head's padding columns are operands too)
    const uint64_t width = q->aux_cols ? (uint64_t)q->heads * hs : (uint64_t)(q->heads - 1) * hs + q->d;
    bool ok = a_map(&tmQ, q->q, width, (uint64_t)q->B * q->n_q, q->ld_q, 128) &&
              a_map(&tmK, q->k, width, (uint64_t)q->B * q->n_kv, q->ld_k, bkv) &&
              a_map(&tmV, q->v, width, (uint64_t)q->B * q->n_kv, q->ld_v, bkv);
    if (!ok) {
        set_error("attention (tcgen05): cuTensorMapEncodeTiled failed (B=%d n_q=%d n_kv=%d d=%d)", q->B, q->n_q, q->n_kv, q->d);
        return ANYSD_ECUDA;
    }
    // paired variant (two query tiles per CTA sharing the K/V tiles): the long small-head level.  ANYSD_ATTN_PAIR=0 switches it off.
    static const char* pair_env = getenv("ANYSD_ATTN_PAIR");
    if (q->aux_cols && a.d_ext <= 64 && q->n_q % (2 * AT_BQ) == 0 && q->n_kv % AW_BKV == 0 && q->n_kv >= 1024 && !bkv_env &&
        !(pair_env && pair_env[0] == '0')) {
        const int smem2 = 2 * AT_ATOM + AW_ST * 2 * AW_BKV * 128 + 4 * AT_ATOM + 256;
        static int attr2[64];
        int dev2 = 0;
        cudaGetDevice(&dev2);
        dev2 &= 63;
        if (!attr2[dev2]) {
            cudaError_t e = cudaFuncSetAttribute(attention_tc5x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
            if (e != cudaSuccess) {
                set_error("attention (tcgen05, paired): smem opt-in failed: %s", cudaGetErrorString(e));
                return ANYSD_ECUDA;
            }
            attr2[dev2] = 1;
        }
        dim3 grid2(q->n_q / (2 * AT_BQ), q->heads, q->B);
        attention_tc5x2_kernel<<<grid2, AW_THREADS, smem2, st>>>(tmQ, tmK, tmV, a);
        return check_launch("attention (tcgen05, paired)");
    }
    const int smem = a.NA * AT_ATOM + a.stages * 2 * a.NA * bkv * 128 + (bkv / 64) * AT_ATOM + 128;
    static int attr_set[64][4];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    const int vi = (bkv == 64 ? 1 : 0) + (q->aux_cols ? 2 : 0);
    const void* fn = bkv == 64 ? (q->aux_cols ? (const void*)attention_tc5_kernel<64, true> : (const void*)attention_tc5_kernel<64, false>)
                               : (q->aux_cols ? (const void*)attention_tc5_kernel<128, true> : (const void*)attention_tc5_kernel<128, false>);
    if (attr_set[dev][vi] < smem) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("attention (tcgen05): smem opt-in failed: %s", cudaGetErrorString(e));
            return ANYSD_ECUDA;
        }
        attr_set[dev][vi] = smem;
    }
    dim3 grid(cdiv(q->n_q, AT_BQ), q->heads, q->B);
    if (bkv == 64) {
        if (q->aux_cols) attention_tc5_kernel<64, true><<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, a);
        else attention_tc5_kernel<64, false><<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, a);
    } else {
        if (q->aux_cols) attention_tc5_kernel<128, true><<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, a);
        else attention_tc5_kernel<128, false><<<grid, AT_THREADS, smem, st>>>(tmQ, tmK, tmV, a);
    }
    return check_launch("attention (tcgen05)");
}

}  // namespace anysd
