// Tensor-core conditioner (FS_PREC_TF32): the residual conditioner of one coupling layer
// (NF/normflows/nets/resnet.py:7-104 in eval mode) for a 128-row tile per CTA, hand-written for
// sm_100a: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM, activations fed back to the tensor
// core as the A operand *in TMEM* (the accumulator layout lane = row, column = feature is exactly the
// A-operand layout, so BatchNorm/ReLU/bias run register-side between tcgen05.ld and tcgen05.st and
// activations never touch shared memory), weights streamed from L2/HBM by TMA bulk copies
// (cp.async.bulk) of tiles that fs_flow_create laid out in the 128B-swizzled K-major image the UMMA
// shared-memory descriptor expects.
//
//   TMEM columns   R0 = [0, H)   : GEMM0 / linear-1 accumulator  -> after the epilogue: A operand `a`
//                  R1 = [H, 2H)  : features / linear-0 accumulator -> after the epilogue: A operand relu(t)
//   each region is split in two halves of NH = H/2 columns = one MMA (M=128, N=NH, K=8) per k-step,
//   so the epilogue of one half overlaps the MMAs of the other.
//   shared memory  h[H][128] FP32 (residual stream, column-major so a warp reads 32 consecutive rows),
//                  NSTAGE weight tiles of NH x 32 tf32 (NH*128 bytes), mbarriers.
//   warps          0: TMA producer   1: MMA issuer   2: TMEM allocator   4-11: epilogue
//                  (warp w touches TMEM lanes 32*(w%4).., warps 4-7 / 8-11 split the columns of a half)
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "flow.cuh"

namespace fs {

static constexpr int TC_KB = 32;          // K elements per weight tile (128 bytes of tf32)
static constexpr int TC_THREADS = 384;

struct TcLayer {
    float* wstream;    // all weight tiles of the layer in consumption order
    float* b_init;     // [H]
    float* bn0_s;      // [n_blocks, H]
    float* bn0_o;      // [n_blocks, H]
    float* b0;         // [n_blocks, H]
    float* b1;         // [n_blocks, H]
    float* b_final;    // [n_chunks * NH]
};

struct TcPack {
    int H, NH, Kp0, n_pieces, n_chunks, nstage;
    size_t tiles_per_layer;
    size_t smem_bytes;
    std::vector<TcLayer> layers;
};

struct TcArgs {
    const float* A0;       // [rows, K0]
    float* theta;          // [rows, NP]
    int rows, K0, NP;
    int Kp0, n_pieces, n_blocks, n_chunks;
    unsigned long long n_tiles;
    TcLayer L;
    int* err;
};

// ---------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > 40000000u) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T, M = 128, N from idesc, K = 8 (tf32)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: 8-row groups of 128-byte rows, 1024 B apart.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                                // leading byte off.  bits [16,30) (unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

// region ids: 0 = R0 half 0, 1 = R0 half 1, 2 = R1 half 0, 3 = R1 half 1
template <int H>
struct TcSmem {
    static constexpr int NH = H / 2;
    static constexpr int STAGE_BYTES = NH * 128;
    static constexpr int H_BYTES = H * 128 * 4;
    static constexpr int NSTAGE = (H == 256) ? 6 : 8;
    static constexpr int BAR_OFF = H_BYTES + NSTAGE * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256;
};

template <int H>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_conditioner_kernel(TcArgs g) {
    using S = TcSmem<H>;
    constexpr int NH = S::NH;
    constexpr int NSTAGE = S::NSTAGE;
    constexpr int KT = H / TC_KB;              // weight tiles along K for an H-wide GEMM
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* hs = (float*)smem;                                   // h[c][r]
    const uint32_t w_base = smem_u32(smem + S::H_BYTES);        // weight stages (1024-aligned)
    const uint32_t bar_base = smem_u32(smem + S::BAR_OFF);
    // barrier map (8 bytes each)
    const uint32_t bar_wfull = bar_base;                        // [NSTAGE]
    const uint32_t bar_wempty = bar_base + 8 * NSTAGE;          // [NSTAGE]
    const uint32_t bar_full = bar_base + 16 * NSTAGE;           // [4]  MMA -> epilogue (accumulator half ready)
    const uint32_t bar_ready = bar_full + 32;                   // [4]  epilogue -> MMA (operand half written / drained)
    const uint32_t bar_free1 = bar_ready + 32;                  // [1]  MMA -> epilogue (feature piece consumed)
    uint32_t* tmem_slot = (uint32_t*)(smem + S::BAR_OFF + 16 * NSTAGE + 72);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int row0 = blockIdx.x * 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(bar_wfull + 8 * i, 1);
            mbar_init(bar_wempty + 8 * i, 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_ready + 8 * i, 8);
        }
        mbar_init(bar_free1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * H));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: stream the layer's weight tiles in order =====================
        if (lane == 0) {
            const uint8_t* src = (const uint8_t*)g.L.wstream;
            uint32_t stage = 0, phase = 0;
            for (unsigned long long t = 0; t < g.n_tiles; ++t) {
                mbar_wait(bar_wempty + 8 * stage, phase ^ 1, g.err, 1);
                mbar_expect_tx(bar_wfull + 8 * stage, S::STAGE_BYTES);
                tma_bulk_g2s(w_base + stage * S::STAGE_BYTES, src + t * S::STAGE_BYTES, S::STAGE_BYTES,
                             bar_wfull + 8 * stage);
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D=F32, A=B=TF32, both K-major, N = NH, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NH >> 3) << 17) | (8u << 24);
            uint32_t stage = 0, wphase = 0;
            uint32_t ph_ready = 0;     // parity to wait for, per region
            auto wait_ready = [&](int region) {
                mbar_wait(bar_ready + 8 * region, (ph_ready >> region) & 1, g.err, 2);
                ph_ready ^= 1u << region;
                tc_fence_after();
            };
            // one weight tile = 4 k-steps of 8
            auto mma_tile = [&](uint32_t dcol, uint32_t acol, bool first) {
                mbar_wait(bar_wfull + 8 * stage, wphase, g.err, 3);
                tc_fence_after();
                const uint32_t sb = w_base + stage * S::STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tc_mma_ts(tmem + dcol, tmem + acol + 8 * j, make_b_desc(sb + 32 * j), idesc,
                              (first && j == 0) ? 0u : 1u);
                tc_commit(bar_wempty + 8 * stage);
                if (++stage == NSTAGE) { stage = 0; wphase ^= 1; }
            };
            // GEMM over K = ktiles*32 columns of operand region `abase` (TMEM column), K-ordered waits on
            // the two halves of that region (region ids ra0, ra0 + 1)
            auto gemm_half = [&](uint32_t dcol, uint32_t abase, int ra0, int ktiles, bool fresh, bool waitA) {
                for (int kt = 0; kt < ktiles; ++kt) {
                    if (waitA) {
                        if (kt == 0) wait_ready(ra0);
                        if (kt * TC_KB == NH) wait_ready(ra0 + 1);
                    }
                    mma_tile(dcol, abase + kt * TC_KB, fresh && kt == 0);
                }
            };
            // ---- GEMM0: features (R1) -> R0 ----
            for (int p = 0; p < g.n_pieces; ++p) {
                const int kcols = min(H, g.Kp0 - p * H);
                const int ktiles = kcols / TC_KB;
                for (int nh = 0; nh < 2; ++nh) {
                    // both feature halves are signalled for every piece; wait for them once per piece
                    if (nh == 0) { wait_ready(2); wait_ready(3); }
                    gemm_half(nh * NH, H, 2, ktiles, p == 0, false);
                    if (p == g.n_pieces - 1) tc_commit(bar_full + 8 * nh);
                }
                if (p < g.n_pieces - 1) tc_commit(bar_free1);
            }
            // ---- residual blocks ----
            for (int b = 0; b < g.n_blocks; ++b) {
                for (int nh = 0; nh < 2; ++nh) {           // linear 0: A = R0 (a), D = R1 half nh
                    gemm_half(H + nh * NH, 0, 0, KT, true, nh == 0);
                    tc_commit(bar_full + 8 * (2 + nh));
                }
                for (int nh = 0; nh < 2; ++nh) {           // linear 1: A = R1 (relu t), D = R0 half nh
                    gemm_half(nh * NH, H, 2, KT, true, nh == 0);
                    tc_commit(bar_full + 8 * nh);
                }
            }
            // ---- final layer: A = R0 (h), D = R1 half (c & 1), chunk after chunk ----
            for (int c = 0; c < g.n_chunks; ++c) {
                if (c >= 2) wait_ready(2 + (c & 1));       // epilogue drained chunk c-2
                gemm_half(H + (c & 1) * NH, 0, 0, KT, true, c == 0);
                tc_commit(bar_full + 8 * (2 + (c & 1)));
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int ew = warp - 4;
        const int q = ew & 3;                    // TMEM lane quadrant (== warp % 4)
        const int cg = ew >> 2;                  // which half of the columns of a region half
        const int r = 32 * q + lane;             // row inside the tile
        const int grow = row0 + r;
        const bool row_ok = grow < g.rows;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        constexpr int CW = NH / 2;               // columns per thread per region half
        uint32_t ph_full = 0, ph_free1 = 0;
        auto wait_full = [&](int region) {
            mbar_wait(bar_full + 8 * region, (ph_full >> region) & 1, g.err, 4);
            ph_full ^= 1u << region;
            tc_fence_after();
        };
        auto signal_ready = [&](int region) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready + 8 * region);
        };
        uint32_t v[32];

        // ---- features -> R1 (A operand of GEMM0), piece by piece ----
        for (int p = 0; p < g.n_pieces; ++p) {
            if (p > 0) {
                mbar_wait(bar_free1, ph_free1, g.err, 5);
                ph_free1 ^= 1;
                tc_fence_after();
            }
            const int kcols = min(H, g.Kp0 - p * H);
            for (int nh = 0; nh < 2; ++nh) {
#pragma unroll 1
                for (int c0 = 0; c0 < CW; c0 += 32) {
                    const int col = nh * NH + cg * CW + c0;          // column inside R1
                    if (col < kcols) {
                        const int k0 = p * H + col;
                        const float* src = g.A0 + (size_t)grow * g.K0 + k0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float x = 0.f;
                            if (row_ok && k0 + i < g.K0) x = __ldg(src + i);
                            v[i] = to_tf32(x);
                        }
                        tc_st32(lane_addr + H + col, v);
                    }
                }
                tc_wait_st();
                signal_ready(2 + nh);
            }
        }
        // ---- GEMM0 epilogue: h = D + b_init ; a = relu(bn0_0(h)) (or raw h when there are no blocks) ----
        auto epi_residual = [&](int nh, const float* bias, const float* s_next, const float* o_next, bool init) {
            wait_full(nh);
#pragma unroll 1
            for (int c0 = 0; c0 < CW; c0 += 32) {
                const int col = nh * NH + cg * CW + c0;
                tc_ld32(lane_addr + col, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float hval = __uint_as_float(v[i]) + __ldg(bias + col + i);
                    float* hp = hs + (size_t)(col + i) * 128 + r;
                    if (!init) hval += *hp;
                    *hp = hval;
                    float a = hval;
                    if (s_next) a = fmaxf(__fmaf_rn(hval, __ldg(s_next + col + i), __ldg(o_next + col + i)), 0.f);
                    v[i] = to_tf32(a);
                }
                tc_st32(lane_addr + col, v);
            }
            tc_wait_st();
            signal_ready(nh);
        };
        for (int nh = 0; nh < 2; ++nh)
            epi_residual(nh, g.L.b_init, g.n_blocks ? g.L.bn0_s : nullptr, g.L.bn0_o, true);
        // ---- residual blocks ----
        for (int b = 0; b < g.n_blocks; ++b) {
            const float* b0 = g.L.b0 + (size_t)b * H;
            for (int nh = 0; nh < 2; ++nh) {                 // relu(t + b0') in place in R1
                wait_full(2 + nh);
#pragma unroll 1
                for (int c0 = 0; c0 < CW; c0 += 32) {
                    const int col = nh * NH + cg * CW + c0;
                    tc_ld32(lane_addr + H + col, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        v[i] = to_tf32(fmaxf(__uint_as_float(v[i]) + __ldg(b0 + col + i), 0.f));
                    tc_st32(lane_addr + H + col, v);
                }
                tc_wait_st();
                signal_ready(2 + nh);
            }
            const bool last = (b == g.n_blocks - 1);
            for (int nh = 0; nh < 2; ++nh)
                epi_residual(nh, g.L.b1 + (size_t)b * H, last ? nullptr : g.L.bn0_s + (size_t)(b + 1) * H,
                             g.L.bn0_o + (size_t)(b + 1) * H, false);
        }
        // ---- final layer: theta chunk = D + b_final -> global ----
        for (int c = 0; c < g.n_chunks; ++c) {
            const int region = 2 + (c & 1);
            wait_full(region);
#pragma unroll 1
            for (int c0 = 0; c0 < CW; c0 += 32) {
                const int col = cg * CW + c0;                 // column inside the chunk
                tc_ld32(lane_addr + H + (c & 1) * NH + col, v);
                tc_wait_ld();
                const int ocol = c * NH + col;
                if (row_ok) {
                    float* dst = g.theta + (size_t)grow * g.NP + ocol;
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (ocol + i < g.NP) dst[i] = __uint_as_float(v[i]) + __ldg(g.L.b_final + ocol + i);
                }
            }
            signal_ready(region);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * H));
    }
}

// ---------------------------------------------------------------------------
// host: pack weights as a stream of pre-swizzled tiles
// ---------------------------------------------------------------------------
static inline float tf32_round(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return x;
    u += 0x00001000u;            // round to nearest, ties away (cvt.rna)
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

// appends the tile rows [n0, n0+NH) x cols [k0, k0+32) of W [n_rows, n_cols] (row-major) in the
// SWIZZLE_128B K-major shared-memory image: row i at i*128 bytes, its 16-byte chunk c at (c ^ (i & 7)).
static void append_tile(std::vector<float>& out, const float* W, int n_rows, int n_cols, int n0, int k0, int NH) {
    const size_t base = out.size();
    out.resize(base + (size_t)NH * 32, 0.f);
    for (int i = 0; i < NH; ++i) {
        const int n = n0 + i;
        if (n >= n_rows) continue;
        for (int c = 0; c < 8; ++c) {
            const int pc = c ^ (i & 7);
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + 4 * c + e;
                if (k < n_cols) out[base + (size_t)i * 32 + pc * 4 + e] = tf32_round(W[(size_t)n * n_cols + k]);
            }
        }
    }
}

template <typename T>
static int tc_upload(fs_flow* f, const std::vector<T>& h, T** out) {
    void* d = nullptr;
    FS_CUDA(cudaMalloc(&d, h.size() * sizeof(T) + 16));
    f->allocs.push_back(d);
    FS_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (T*)d;
    return FS_OK;
}

int tc_pack(fs_flow* f, const fs_flow_desc* d) {
    f->tc = nullptr;
    const int H = f->H;
    if ((H != 128 && H != 256) || f->n_blocks < 1) return FS_OK;     // shape not covered: FP32 path only
    int dev = 0, smem_max = 0, cc = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return FS_OK;
    cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (cc != 10) return FS_OK;
    TcPack* P = new TcPack();
    P->H = H;
    P->NH = H / 2;
    const int K0 = 2 * f->N;
    P->Kp0 = (K0 + TC_KB - 1) / TC_KB * TC_KB;
    P->n_pieces = (P->Kp0 + H - 1) / H;
    const int NP = f->N * f->P;
    P->n_chunks = (NP + P->NH - 1) / P->NH;
    P->smem_bytes = (H == 256 ? TcSmem<256>::TOTAL : TcSmem<128>::TOTAL) + 1024;
    if ((int)P->smem_bytes > smem_max) { delete P; return FS_OK; }
    const int NH = P->NH, KT = H / TC_KB;
    P->layers.resize(f->K);
    for (int li = 0; li < f->K; ++li) {
        const fs_layer_params* p = &d->layers[li];
        // folded parameters exactly as the FP32 path uses them (flow.cu: pack_layer)
        const int nB = f->n_blocks;
        std::vector<float> s0((size_t)nB * H), o0((size_t)nB * H), w0((size_t)nB * H * H), b0((size_t)nB * H);
        for (int b = 0; b < nB; ++b)
            for (int j = 0; j < 2; ++j) {
                const size_t o = ((size_t)b * 2 + j) * H;
                for (int c = 0; c < H; ++c) {
                    const double sc = (double)p->bn_w[o + c] / sqrt((double)p->bn_var[o + c] + (double)d->bn_eps);
                    const double of = (double)p->bn_b[o + c] - (double)p->bn_mean[o + c] * sc;
                    if (j == 0) {
                        s0[(size_t)b * H + c] = (float)sc;
                        o0[(size_t)b * H + c] = (float)of;
                    } else {
                        const float* wr = p->lin_w + (((size_t)b * 2 + 0) * H + c) * H;
                        for (int k = 0; k < H; ++k) w0[((size_t)b * H + c) * H + k] = (float)(sc * (double)wr[k]);
                        b0[(size_t)b * H + c] = (float)(sc * (double)p->lin_b[((size_t)b * 2 + 0) * H + c] + of);
                    }
                }
            }
        std::vector<float> stream;
        stream.reserve((size_t)NH * 32 * (2 * P->Kp0 / TC_KB + (size_t)nB * 4 * KT + (size_t)P->n_chunks * KT));
        for (int pc = 0; pc < P->n_pieces; ++pc) {                       // GEMM0
            const int kcols = std::min(H, P->Kp0 - pc * H);
            for (int nh = 0; nh < 2; ++nh)
                for (int kt = 0; kt < kcols / TC_KB; ++kt)
                    append_tile(stream, p->init_w, H, K0, nh * NH, pc * H + kt * TC_KB, NH);
        }
        for (int b = 0; b < nB; ++b) {
            for (int nh = 0; nh < 2; ++nh)                               // linear 0 (BN1 folded)
                for (int kt = 0; kt < KT; ++kt) append_tile(stream, &w0[(size_t)b * H * H], H, H, nh * NH, kt * TC_KB, NH);
            const float* w1 = p->lin_w + ((size_t)b * 2 + 1) * H * H;
            for (int nh = 0; nh < 2; ++nh)                               // linear 1
                for (int kt = 0; kt < KT; ++kt) append_tile(stream, w1, H, H, nh * NH, kt * TC_KB, NH);
        }
        for (int c = 0; c < P->n_chunks; ++c)                            // final layer
            for (int kt = 0; kt < KT; ++kt) append_tile(stream, p->final_w, NP, H, c * NH, kt * TC_KB, NH);
        P->tiles_per_layer = stream.size() / ((size_t)NH * 32);
        TcLayer& L = P->layers[li];
        std::vector<float> v;
        int r = tc_upload(f, stream, &L.wstream);
        v.assign(p->init_b, p->init_b + H);
        if (!r) r = tc_upload(f, v, &L.b_init);
        if (!r) r = tc_upload(f, s0, &L.bn0_s);
        if (!r) r = tc_upload(f, o0, &L.bn0_o);
        if (!r) r = tc_upload(f, b0, &L.b0);
        std::vector<float> b1((size_t)nB * H);
        for (int b = 0; b < nB; ++b) memcpy(&b1[(size_t)b * H], p->lin_b + ((size_t)b * 2 + 1) * H, sizeof(float) * H);
        if (!r) r = tc_upload(f, b1, &L.b1);
        v.assign((size_t)P->n_chunks * NH, 0.f);
        memcpy(v.data(), p->final_b, sizeof(float) * NP);
        if (!r) r = tc_upload(f, v, &L.b_final);
        if (r) { delete P; return r; }
    }
    int* err = nullptr;
    if (cudaMalloc(&err, sizeof(int)) != cudaSuccess) { delete P; return FS_ERR_CUDA; }
    cudaMemset(err, 0, sizeof(int));
    f->allocs.push_back(err);
    f->tc = P;
    f->tc_err = err;
    if (H == 256)
        FS_CUDA(cudaFuncSetAttribute(tc_conditioner_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_bytes));
    else
        FS_CUDA(cudaFuncSetAttribute(tc_conditioner_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem_bytes));
    return FS_OK;
}

void tc_free(fs_flow* f) {
    if (f->tc) delete (TcPack*)f->tc;
    f->tc = nullptr;
}

size_t tc_workspace_bytes(const fs_flow*, int) { return 0; }

int tc_conditioner(fs_flow* f, int layer, const float* A0, int rows, float* theta, void*, size_t, cudaStream_t s) {
    TcPack* P = (TcPack*)f->tc;
    if (!P) {
        set_error("tensor-core conditioner not available for this flow shape (H=%d, blocks=%d)", f->H, f->n_blocks);
        return FS_ERR_UNSUPPORTED;
    }
    TcArgs g;
    g.A0 = A0;
    g.theta = theta;
    g.rows = rows;
    g.K0 = 2 * f->N;
    g.NP = f->N * f->P;
    g.Kp0 = P->Kp0;
    g.n_pieces = P->n_pieces;
    g.n_blocks = f->n_blocks;
    g.n_chunks = P->n_chunks;
    g.n_tiles = P->tiles_per_layer;
    g.L = P->layers[layer];
    g.err = f->tc_err;
    const int grid = (rows + 127) / 128;
    if (P->H == 256)
        tc_conditioner_kernel<256><<<grid, TC_THREADS, P->smem_bytes, s>>>(g);
    else
        tc_conditioner_kernel<128><<<grid, TC_THREADS, P->smem_bytes, s>>>(g);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "tc_conditioner_kernel");
}

}  // namespace fs
