// Tensor-core conditioner (FS_PREC_TF32 selects this path): the residual conditioner of one coupling layer
// (NF/normflows/nets/resnet.py:7-104 in eval mode) and, on the fused path, the conditional spline of the layer
// (flows/neural_spline/coupling.py:86-135, utils/splines.py:84-222) for a 128-row tile per CTA, hand-written for
// sm_100a: tcgen05.mma with FP32 accumulators in TMEM, activations fed back to the tensor core as the A operand
// *in TMEM* (TS mode: lane = row, so BatchNorm / ReLU / bias run register-side between tcgen05.ld and tcgen05.st and
// activations never touch shared memory), weights streamed from L2/HBM by TMA bulk copies (cp.async.bulk) of tiles that
// fs_flow_create laid out in the 128B-swizzled K-major image the UMMA shared-memory descriptor expects.
//
//   operands       GEMM0 (periodic features -> H): kind::tf32.  Every later GEMM: kind::f16 with FP16 operands (same
//                  11-bit significand as TF32, half the bytes, K = 16 per instruction = twice the rate).  An epilogue
//                  thread packs the 32 features it owns into the first 16 of its own 32 accumulator columns
//                  (cvt.rn[.relu].f16x2.f32), so the in-place hand-over never touches a column another thread still has
//                  to read; K-block j of a 64-feature weight tile is addressed at column 32 (j / 2) + 8 (j % 2).
//   TMEM columns   R0 = [0, H)   : GEMM0 / linear-1 accumulator  -> after the epilogue: operand `a` (packed halves)
//                  R1 = [H, 2H)  : features / linear-0 accumulator -> after the epilogue: operand relu(t)
//                  final layer   : two accumulators ping-pong in R1 (H = 256: H and H + 128; H = 128: R1 and 2H)
//   schedule       H = 256: every H x H GEMM = (all columns, K lo: N = 256 MMAs) | (columns lo, K hi: N = 128) |
//                  (columns hi, K hi): the first part needs only the low half of the operand, the accumulator's low
//                  half completes a quarter-GEMM early; per-half FULL / RDY mbarriers hand over in both directions.
//   residual       the stream u (biases folded out at pack time) belongs to the epilogue threads: thread (quadrant q,
//                  column group c) owns u[row 32q+lane][half*NH + 32c .. +32): half 0 in registers, half 1 in smem.
//   final layer    fused path: one chunk of chn <= 112 columns per transformed coordinate
//                  [widths | pad | heights | pad | derivatives]; the epilogue warps of a quadrant form two pairs that
//                  take alternate chunks: softmax / prefix sums / bin search / rational-quadratic evaluation in registers,
//                  output coordinate and log-det written directly (theta is never materialised).  theta path
//                  (fs_flow_conditioner, FS_NO_FUSE=1): TF32, 128-column chunks, coalesced stores through a per-warp
//                  smem transpose.
//   shared memory  NSTAGE weight stages of H*128 bytes, the second half of u (reused by the final layer), two parameter
//                  sets (BatchNorm scale / offset, folded bias) prefetched by TMA, mbarriers, mailboxes.
//   warps          0: TMA producer   1: MMA issuer (converged warp, elect.sync; one smem descriptor per stage)
//                  2: TMEM allocator   4..19: epilogue (warp w touches TMEM lanes 32*(w%4)..).
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "flow.cuh"

namespace fs {

struct TcArgs {
    const float* A0;       // [rows, K0] row-major, or row-tiled (a0_tiled, flow.cuh) when a0_tiled != 0
    int a0_tiled;
    float* theta;          // [rows, NP]
    int rows, K0, NP;
    int Kp0, n_pieces, n_blocks, n_chunks;
    unsigned long long n_tiles, tiles_before_final;
    TcLayer L;
    // fused spline epilogue (fused != 0): the final layer runs one chunk per transformed coordinate and the
    // epilogue applies the spline in place of writing theta.  1 = density direction, 2 = sampling direction.
    int fused, N, D, nb, chn;
    float bound, inv_sqrt_h;
    const float* xin;      // [rows, D] layer input
    float* xout;           // [rows, D] layer output (transformed half written here)
    float* logdet;         // [rows] accumulated, or nullptr
    const int* xc_in;      // [N] column of the layer input read for transformed coordinate c (this direction)
    const int* xc_out;     // [N] column of the layer output it is written to
    int* nan_flag;
    int* err;
    long long* dbg;   // optional per-CTA wait-cycle counters (FS_TC_DEBUG=1), 16 per CTA
    // Layer-parallel launch (lp_layers > 0, fused path only): CTA b works on layer step b / tiles (layer
    // lp_rev ? lp_layers - 1 - step : step) of row tile b % tiles.  Every conditioner input is known up front (the
    // identity coordinates only ever go through the unconditional splines: SURVEY.md A.4-Q2), so the GEMM stacks of all
    // layers are independent; only the spline of the transformed coordinates is a chain.  It is ordered per (tile, lane
    // quadrant) by ONE counter in `flags`: the two B warps of the quadrant bump it (after a device-scope fence) when
    // they have written their last coordinate, and step s starts its first chunk only when step s - 1 shows 2.  (A
    // per-chunk hand-over was tried: the consumer then stalls in the middle of its chunk loop, which turned out not to
    // be repeatable run to run at H = 128; with one wait in front of the loop the loop itself is the single-layer one.)
    const TcLayer* Ls;     // device array [lp_layers]
    int lp_layers, lp_rev, tiles;
    unsigned long long a0_stride, ld_stride;   // floats between the feature matrices / log-det partials of two steps
    float* buf0;           // activations: step s reads buf[s & 1] and writes buf[(s + 1) & 1] (transformed columns only)
    float* buf1;
    int* flags;            // [lp_layers][tiles][4 lane quadrants]: B warps of that quadrant that have finished the step
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {   // non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code,
                                          long long* waited = nullptr) {
    const long long t0 = waited ? clock64() : 0;
    if (mbar_test(bar, parity)) {
        if (waited) *waited += clock64() - t0;
        return;
    }
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > 40000000u) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
    if (waited) *waited += clock64() - t0;
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
// One lane of a CONVERGED warp.  tcgen05.mma / commit / TMA must be issued from warp-uniform control
// flow: from a divergent `if (lane == 0)` branch the compiler wraps every UTCHMMA in an R2UR waterfall
// loop and the issue rate drops to ~92 clk per MMA (measured, scripts/mma_bench3.cu) - slower than the
// 64 clk the tensor pipe needs for a 128x128x8 tf32 MMA.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// two consecutive K elements in one 32-bit TMEM column: the lower half holds the even one
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// the same with ReLU folded into the conversion (negative inputs become +0; out-of-range inputs become +inf and are
// caught by the range guard of the epilogue, see hmax)
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// Epilogue arithmetic.  The epilogue warps are issue-bound, so they use Blackwell's packed FP32 pairs
// (FADD2 / FFMA2) and fold ReLU and the TF32 rounding into one integer instruction: for finite x,
// round-to-nearest (ties away) to 10 mantissa bits is bits(x) + 0x1000 with the low 13 bits ignored by the
// tensor core, and relu on the bit pattern is a signed max with 0, so
// tf32(relu(x)) = max(bits(x) + 0x1000, 0x1000)  (VIADDMNMX).  Same values as cvt.rna.tf32.f32 for finite
// inputs; an infinity turns into a NaN, which the flow's nan_flag reports either way.
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint32_t relu_tf32(float x) {
    return (uint32_t)__viaddmax_s32(__float_as_int(x), 0x1000, 0x1000);
}
__device__ __forceinline__ uint32_t round_tf32(float x) { return __float_as_uint(x) + 0x1000u; }

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

// barrier ids.  full: MMA -> epilogue (accumulator complete).  rdy: epilogue -> MMA (operand half written, or
// final-layer accumulator drained).
enum { FULL_R0H0 = 0, FULL_R0H1 = 1, FULL_R1H0 = 2, FULL_R1H1 = 3, FULL_F0 = 4, FULL_F1 = 5, FULL_F2 = 6, N_FULL = 7 };
enum { RDY_R0H0 = 0, RDY_R0H1 = 1, RDY_R1H0 = 2, RDY_R1H1 = 3, RDY_F0 = 4, RDY_F1 = 5, RDY_F2 = 6, N_RDY = 7 };

// In-kernel wait/phase timers (scripts/tc_debug.py): compiled in only with -DFS_TC_TIMERS=1; the clock reads sit in
// the issue loops and cost the single MMA-issuing warp real time.
#ifndef FS_TC_TIMERS
#define FS_TC_TIMERS 0
#endif

template <int H>
struct TcCfg {
    static constexpr int NH = H / 2;
    static constexpr int EPI_WARPS = 16;                     // 4 TMEM lane quadrants x 4 column groups
    static constexpr int BLK_WARPS = 4 * (NH / 32);          // of which take part in the H-wide stages (one 32-column
                                                             // chunk per region half per thread): 16 (H = 256) / 8 (H = 128);
                                                             // all 16 work on the final layer
    static constexpr int THREADS = 128 + 32 * EPI_WARPS;
    static constexpr int STAGE_BYTES = H * 128;              // H rows x 32 tf32
    static constexpr int NSTAGE = (H == 256) ? 4 : 8;
    static constexpr int GROUP = 2;                            // stages issued per barrier batch (>= 512 clk of MMAs)
    static constexpr int FCH = 128;                          // final-layer chunk width (MMA N)
    static constexpr bool SPLIT = (H == 256);                // block GEMMs as N = 128 quarter-GEMMs (see gemm_split)
    static constexpr int KPS = STAGE_BYTES / (FCH * 128);    // final-layer k-tiles per stage
    static constexpr int FIN0 = H;                           // TMEM column of final accumulator 0 (theta path: two, ping-pong)
    static constexpr int FIN1 = (H == 256) ? H + 128 : 2 * H;
    // Fused final layer: THREE accumulators in rotation (chunk c -> accumulator c % 3) while the two epilogue pairs
    // still alternate (chunk c -> pair c % 2), so the tensor pipe works one chunk ahead of the pairs instead of
    // idling until the pair that owns its only other accumulator has read it out.  H = 256 has no room for a third
    // accumulator next to the operand's scattered hand-over layout, so the LAST residual step writes the packed
    // operand compactly into the (dead) low half of R1, features 2j, 2j+1 in column H + j, and the accumulators
    // take R0 and the high half of R1.  H = 128: operand stays in R0, accumulators at 128 / 256 / 384.
    static constexpr bool COMPACT = (H == 256);
    static constexpr int FA = COMPACT ? H : 0;               // operand column of the fused final layer
    static constexpr int FINF0 = COMPACT ? 0 : 128;
    static constexpr int FINF1 = COMPACT ? 128 : 256;
    static constexpr int FINF2 = 384;
    static constexpr int TMEM_COLS = 512;
    static constexpr int PSET_FLOATS = 3 * H;                // per parameter set: b0' | s | o'
    static constexpr int U_OFF = NSTAGE * STAGE_BYTES;       // second column half of the residual stream: [NH/4][128] float4;
                                                             // in the final layer 4 KB per epilogue warp (transpose patch / parked values)
    static constexpr int U_BYTES = (NH * 128 * 4 > EPI_WARPS * 4096) ? NH * 128 * 4 : EPI_WARPS * 4096;
    static constexpr int P_OFF = U_OFF + U_BYTES;
    static constexpr int BAR_OFF = P_OFF + 2 * PSET_FLOATS * 4;
    static constexpr int MB_OFF = BAR_OFF + 512;             // fused epilogue mailboxes: 8 pair groups x 2 x 5 x 32 floats
    static constexpr int BIAS_SLOTS = 8;                     // fused final layer: ring of per-coordinate bias vectors
    static constexpr int BIAS_OFF = MB_OFF + 8 * 320 * 4;    // (<= 128 floats each), filled by the TMA producer
    static constexpr int TOTAL = BIAS_OFF + BIAS_SLOTS * 512;
};

template <int H>
__global__ void __launch_bounds__(TcCfg<H>::THREADS, 1) tc_conditioner_kernel(TcArgs g) {
    long long* const dbg = FS_TC_TIMERS ? g.dbg : nullptr;
    using S = TcCfg<H>;
    constexpr int NH = S::NH;
    constexpr int NSTAGE = S::NSTAGE;
    constexpr int EPI_WARPS = S::EPI_WARPS;
    constexpr int KT = H / TC_KB;              // weight stages along K for an H-wide GEMM
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t w_base = smem_u32(smem);                     // weight stages (1024-aligned)
    const float* pbuf = reinterpret_cast<const float*>(smem + S::P_OFF);
    const uint32_t p_base = smem_u32(smem + S::P_OFF);
    const uint32_t bar_base = smem_u32(smem + S::BAR_OFF);
    // barrier map (8 bytes each)
    const uint32_t bar_wfull = bar_base;                        // [NSTAGE] TMA -> MMA
    const uint32_t bar_wempty = bar_base + 8 * NSTAGE;          // [NSTAGE] MMA -> TMA
    const uint32_t bar_full = bar_base + 16 * NSTAGE;           // [N_FULL]
    const uint32_t bar_rdy = bar_full + 8 * N_FULL;             // [N_RDY]
    const uint32_t bar_free1 = bar_rdy + 8 * N_RDY;             // [1]  MMA -> epilogue (feature piece consumed)
    const uint32_t bar_pfull = bar_free1 + 8;                   // [2]  TMA -> epilogue (parameter set landed)
    const uint32_t bar_pempty = bar_pfull + 16;                 // [2]  epilogue -> TMA
    const uint32_t bar_bfull = bar_pempty + 16;                 // [BIAS_SLOTS] TMA -> epilogue pairs (bias of a chunk landed)
    const uint32_t bar_bempty = bar_bfull + 8 * S::BIAS_SLOTS;  // [BIAS_SLOTS] epilogue pairs -> TMA
    uint32_t* tmem_slot = (uint32_t*)(smem + S::BAR_OFF + 16 * NSTAGE + 8 * (N_FULL + N_RDY) + 56 + 16 * S::BIAS_SLOTS);
    static_assert(16 * NSTAGE + 8 * (N_FULL + N_RDY) + 60 + 16 * S::BIAS_SLOTS <= 512, "barrier area overflow");
    const uint32_t bias_base = smem_u32(smem + S::BIAS_OFF);
    float4* us4 = reinterpret_cast<float4*>(smem + S::U_OFF);   // u of column half 1: [col/4][row]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int lp = g.lp_layers;
    const int step = lp ? (int)blockIdx.x / g.tiles : 0;
    const int tile = lp ? (int)blockIdx.x - step * g.tiles : (int)blockIdx.x;
    const TcLayer L = lp ? g.Ls[g.lp_rev ? lp - 1 - step : step] : g.L;
    const int row0 = tile * 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(bar_wfull + 8 * i, 1);
            mbar_init(bar_wempty + 8 * i, 1);
        }
        for (int i = 0; i < N_FULL; ++i) mbar_init(bar_full + 8 * i, 1);
        for (int i = 0; i < N_RDY; ++i)
            mbar_init(bar_rdy + 8 * i, i < RDY_F0 ? S::BLK_WARPS : (g.fused ? EPI_WARPS / 2 : EPI_WARPS));
        mbar_init(bar_free1, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_pfull + 8 * i, 1);
            mbar_init(bar_pempty + 8 * i, S::BLK_WARPS);
        }
        for (int i = 0; i < S::BIAS_SLOTS; ++i) {
            mbar_init(bar_bfull + 8 * i, 1);
            mbar_init(bar_bempty + 8 * i, EPI_WARPS / 2);       // the pair that owns the chunk, in each of the four lane quadrants
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(S::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const bool f16mode = g.fused != 0;   // fused final layer: u is handed over as packed halves like every block operand

    if (warp < 4) {
        if (warp == 0) {
            // ===================== TMA producer: parameter sets + the layer's weight stages, in order =====================
            const uint8_t* src = (const uint8_t*)L.wstream;
            uint32_t stage = 0, phase = 0;
            long long w_empty = 0;
            unsigned long long t = 0;
            uint32_t pph = 0;                  // per-buffer parity of bar_pempty
            auto load_pset = [&](int j) {      // set j: 0 = {-, s_0, o'_0}; b+1 = {b0'_b, s_{b+1}, o'_{b+1}}
                const int bsel = j & 1;
                mbar_wait(bar_pempty + 8 * bsel, ((pph >> bsel) & 1) ^ 1, g.err, 6);
                pph ^= 1u << bsel;
                if (elect_one()) {
                    mbar_expect_tx(bar_pfull + 8 * bsel, S::PSET_FLOATS * 4);
                    tma_bulk_g2s(p_base + bsel * S::PSET_FLOATS * 4, L.psets + (size_t)j * S::PSET_FLOATS,
                                 S::PSET_FLOATS * 4, bar_pfull + 8 * bsel);
                }
                __syncwarp();
            };
            auto stream = [&](unsigned long long n, uint32_t nbytes) {   // n stages of nbytes each, contiguous at src
                for (unsigned long long e = t + n; t < e; ++t) {
                    mbar_wait(bar_wempty + 8 * stage, phase ^ 1, g.err, 1, dbg ? &w_empty : nullptr);
                    // Bulk copies issued by ONE thread complete one at a time (~800 clk each, measured:
                    // scripts/tma_bw2.cu); copies issued by different lanes overlap -> rotate the issuing lane.
                    if (lane == (int)(t & 7)) {
                        mbar_expect_tx(bar_wfull + 8 * stage, nbytes);
                        tma_bulk_g2s(w_base + stage * S::STAGE_BYTES, src, nbytes, bar_wfull + 8 * stage);
                    }
                    src += nbytes;
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            };
            load_pset(0);
            load_pset(1);
            stream((unsigned long long)(g.Kp0 / TC_KB), S::STAGE_BYTES);          // GEMM0
            for (int b = 0; b < g.n_blocks; ++b) {
                if (b >= 1) load_pset(b + 1);
                stream(2ull * (H / 64), S::STAGE_BYTES);           // two GEMMs of H/64 FP16 stages
            }
            if (g.fused) {                                                        // final layer, FP16 tiles of 64 k
                src = (const uint8_t*)L.wfused;
                // per coordinate chunk: its bias vector into the ring (the epilogue pair reads it from shared memory: no
                // global loads on its serial chain), then its weight stages
                for (int c = 0; c < g.N; ++c) {
                    const int slot = c % S::BIAS_SLOTS;
                    mbar_wait(bar_bempty + 8 * slot, ((c / S::BIAS_SLOTS) & 1) ^ 1, g.err, 8);
                    if (elect_one()) {
                        mbar_expect_tx(bar_bfull + 8 * slot, (uint32_t)g.chn * 4);
                        tma_bulk_g2s(bias_base + slot * 512, L.b_fused + (size_t)c * g.chn, (uint32_t)g.chn * 4,
                                     bar_bfull + 8 * slot);
                    }
                    __syncwarp();
                    stream((unsigned long long)((H / 64) / S::KPS), (uint32_t)(S::KPS * g.chn * 128));
                }
            } else {
                stream((unsigned long long)g.n_chunks * (KT / S::KPS), S::STAGE_BYTES);
            }
            if (dbg && lane == 0) dbg[16 * blockIdx.x + 0] = w_empty;
        } else if (warp == 1) {
            // ===================== MMA issuer (whole warp converged, one elected lane issues) =====================
            // instruction descriptors: D=F32, A=B=TF32, both K-major, M = 128, N = H (blocks) / 128 (final layer)
            const uint32_t idesc_blk = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(H >> 3) << 17) | (8u << 24);
            const uint32_t idesc_fin = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(S::FCH >> 3) << 17) | (8u << 24);
            // fused final layer: chunk of g.chn weight rows (one coordinate's spline parameters), FP16 operands:
            // D = F32, A = B = F16 (format 0), K = 16 per instruction
            const uint32_t idesc_f16 = (1u << 4) | ((uint32_t)(g.chn >> 3) << 17) | (8u << 24);
            const uint32_t fus_tile = (uint32_t)g.chn * 128;
            // residual blocks with FP16 operands: N = H and N = 128
            const uint32_t idesc_blk16 = (1u << 4) | ((uint32_t)(H >> 3) << 17) | (8u << 24);
            const uint32_t idesc_q16 = (1u << 4) | ((uint32_t)(S::FCH >> 3) << 17) | (8u << 24);
            uint32_t stage = 0, wphase = 0;
            uint32_t ph_rdy = 0;       // parity to wait for, per rdy barrier
            long long w_ready = 0, w_weights = 0, t_issue = 0;
            const long long t_start = dbg ? clock64() : 0;
            auto wait_rdy = [&](int id) {
                mbar_wait(bar_rdy + 8 * id, (ph_rdy >> id) & 1, g.err, 2, dbg ? &w_ready : nullptr);
                ph_rdy ^= 1u << id;
                tc_fence_after();
            };
            auto commit = [&](uint32_t bar) {
                if (elect_one()) tc_commit(bar);
                __syncwarp();
            };
            // Issues `n` weight stages (n <= GROUP): their full-barriers are waited for back to back, then all the
            // MMAs are queued at once.  Block GEMMs: one stage = one 32-wide k-tile of all H output columns
            // (4 MMAs of N = H).  Final layer: one stage = KPS consecutive k-tiles of a 128-column chunk.
            // FP16 operands: a weight tile row is 64 halves = 4 MMAs of K = 16.  The epilogue thread that owns 32 features
            // packs them into the first 16 of its own 32 TMEM columns (no thread ever writes a column another thread
            // still has to read), so K-block j of a 64-feature tile sits at column 32 (j / 2) + 8 (j % 2).
            auto issue = [&](uint32_t dcol, uint32_t acol, int n, bool first, bool fin, bool fus = false, bool f16 = false) {
                uint32_t st[S::GROUP];
#pragma unroll
                for (int i = 0; i < S::GROUP; ++i) {
                    if (i < n) {
                        mbar_wait(bar_wfull + 8 * stage, wphase, g.err, 3, dbg ? &w_weights : nullptr);
                        st[i] = stage;
                        if (++stage == NSTAGE) { stage = 0; wphase ^= 1; }
                    }
                }
                tc_fence_after();
                const long long ti = dbg ? clock64() : 0;
                if (elect_one()) {
#pragma unroll
                    for (int i = 0; i < S::GROUP; ++i) {
                        if (i < n) {
                            // one descriptor per stage; a tile / k-slice offset only moves the 14-bit start-address
                            // field (units of 16 bytes, no carry out of it: shared memory is < 256 KB)
                            const uint64_t d0 = make_b_desc(w_base + st[i] * S::STAGE_BYTES);
                            const uint32_t d0_lo = (uint32_t)d0, d0_hi = (uint32_t)(d0 >> 32);
                            auto desc = [&](uint32_t byte_off) {
                                return ((uint64_t)d0_hi << 32) | (uint64_t)(d0_lo + (byte_off >> 4));
                            };
                            if (!fin) {
                                if (f16) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        tc_mma_ts_f16(tmem + dcol, tmem + acol + 64 * i + 32 * (j >> 1) + 8 * (j & 1),
                                                      desc(32 * j), idesc_blk16, (first && i == 0 && j == 0) ? 0u : 1u);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        tc_mma_ts(tmem + dcol, tmem + acol + 32 * i + 8 * j, desc(32 * j),
                                                  idesc_blk, (first && i == 0 && j == 0) ? 0u : 1u);
                                }
                            } else {
                                const uint32_t tile = fus ? fus_tile : (uint32_t)(S::FCH * 128);
                                const uint32_t idesc = idesc_fin;
                                if (f16) {
#pragma unroll
                                    for (int kk = 0; kk < S::KPS; ++kk)
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            tc_mma_ts_f16(tmem + dcol,
                                                          (S::COMPACT && fus)
                                                              ? tmem + acol + 32 * (i * S::KPS + kk) + 8 * j
                                                              : tmem + acol + 64 * (i * S::KPS + kk) + 32 * (j >> 1) + 8 * (j & 1),
                                                          desc(kk * tile + 32 * j), fus ? idesc_f16 : idesc_q16,
                                                          (first && i == 0 && kk == 0 && j == 0) ? 0u : 1u);
                                } else {
#pragma unroll
                                    for (int kk = 0; kk < S::KPS; ++kk)
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            tc_mma_ts(tmem + dcol, tmem + acol + 32 * (i * S::KPS + kk) + 8 * j,
                                                      desc(kk * tile + 32 * j), idesc,
                                                      (first && i == 0 && kk == 0 && j == 0) ? 0u : 1u);
                                }
                            }
                            tc_commit(bar_wempty + 8 * st[i]);
                        }
                    }
                }
                __syncwarp();
                if (dbg) t_issue += clock64() - ti;
            };
            // block-type GEMM over `ktiles` k-tiles of operand region `abase`; K-ordered waits on its two halves
            auto gemm_blk = [&](uint32_t dcol, uint32_t abase, int rdy0, int ktiles, bool fresh, bool waitA) {
                for (int kt = 0; kt < ktiles; kt += S::GROUP) {
                    if (waitA) {
                        if (kt == 0) wait_rdy(rdy0);
                        if (kt * TC_KB == NH) wait_rdy(rdy0 + 1);
                    }
                    issue(dcol, abase + kt * TC_KB, min(S::GROUP, ktiles - kt), fresh && kt == 0, false);
                }
            };
            // the same with FP16 operands: k-tiles of 64 features (64 TMEM columns of the operand region each)
            auto gemm_blk16 = [&](uint32_t dcol, uint32_t abase, int rdy0) {
                constexpr int KT16 = H / 64;
                for (int kt = 0; kt < KT16; ++kt) {             // one stage per call: the half boundary may fall on any tile
                    if (kt == 0) wait_rdy(rdy0);
                    if (kt * 64 == NH) wait_rdy(rdy0 + 1);
                    issue(dcol, abase + kt * 64, 1, kt == 0, false, false, true);
                }
            };
            // ---- GEMM0: features (R1) -> R0 ----
            for (int p = 0; p < g.n_pieces; ++p) {
                const int kcols = min(H, g.Kp0 - p * H);
                wait_rdy(RDY_R1H0);
                wait_rdy(RDY_R1H1);
                gemm_blk(0, H, RDY_R1H0, kcols / TC_KB, p == 0, false);
                if (p < g.n_pieces - 1) commit(bar_free1);
            }
            commit(bar_full + 8 * FULL_R0H0);
            commit(bar_full + 8 * FULL_R0H1);
            // H = 256: split schedule.  Each H x H GEMM runs as
            //   (all columns, K lo: 16 MMAs of N = 256) | (columns lo, K hi: 16 MMAs of N = 128) (columns hi, K hi: 16 of N = 128)
            // The first part only needs the low half of the operand region, so it starts while the epilogue is still
            // writing the high half; the accumulator's low column half completes a quarter-GEMM before the high one,
            // so the epilogue starts on it while the tensor pipe finishes the other.  K-lo stages are k-tiles of all
            // H weight rows; K-hi stages hold KPS k-tiles of a 128-row weight chunk (the final layer's stage format).
            auto gemm_split = [&](uint32_t dcol, uint32_t abase, int rdy0, int full0) {
                constexpr int KT16 = H / 64;                    // k-tiles of 64 features (FP16 operands)
                constexpr int SPQ = (KT16 / 2) / S::KPS;        // stages per quarter
                wait_rdy(rdy0);
#pragma unroll 1
                for (int kt = 0; kt < KT16 / 2; ++kt) issue(dcol, abase + kt * 64, 1, kt == 0, false, false, true);
                wait_rdy(rdy0 + 1);
#pragma unroll 1
                for (int nh = 0; nh < 2; ++nh) {
                    for (int sg = 0; sg < SPQ; ++sg)
                        issue(dcol + nh * NH, abase + NH + sg * S::KPS * 64, 1, false, true, false, true);
                    commit(bar_full + 8 * (full0 + nh));
                }
            };
            // ---- residual blocks ----
            for (int b = 0; b < g.n_blocks; ++b) {
                if (S::SPLIT) {
                    gemm_split(H, 0, RDY_R0H0, FULL_R1H0);          // linear 0: A = R0 (a), D = R1
                    gemm_split(0, H, RDY_R1H0, FULL_R0H0);          // linear 1: A = R1 (relu t), D = R0
                } else {
                    gemm_blk16(H, 0, RDY_R0H0);
                    commit(bar_full + 8 * FULL_R1H0);
                    commit(bar_full + 8 * FULL_R1H1);
                    gemm_blk16(0, H, RDY_R1H0);
                    commit(bar_full + 8 * FULL_R0H0);
                    commit(bar_full + 8 * FULL_R0H1);
                }
            }
            // ---- final layer: A = R0 (u).  theta path: 128-column accumulators ping-pong, chunk after chunk.  Fused
            //      path: one chunk per transformed coordinate, three accumulators in rotation (see TcCfg). ----
            wait_rdy(RDY_R0H0);
            wait_rdy(RDY_R0H1);
            const int n_final = g.fused ? g.N : g.n_chunks;
            const long long t_final0 = dbg ? clock64() : 0;
            long long w_rdyf = 0;
            const bool fus = g.fused != 0;
            const int SPC = (fus ? H / 64 : KT) / S::KPS;       // stages per chunk
            const int n_acc = fus ? 3 : 2;
            const uint32_t astep = fus ? (S::COMPACT ? 32u : 64u) : (uint32_t)TC_KB;   // operand columns per k-tile
            int f = 0;                                          // accumulator of chunk c: c % n_acc
            for (int c = 0; c < n_final; ++c) {
                if (c >= n_acc) {                               // the epilogue drained chunk c - n_acc
                    const long long w0 = w_ready;
                    wait_rdy(RDY_F0 + f);
                    w_rdyf += w_ready - w0;
                }
                const uint32_t dcol = fus ? (f == 0 ? S::FINF0 : (f == 1 ? S::FINF1 : S::FINF2)) : (f ? S::FIN1 : S::FIN0);
                for (int sg = 0; sg < SPC; sg += S::GROUP)
                    issue(dcol, (fus ? S::FA : 0) + sg * S::KPS * astep, min(S::GROUP, SPC - sg), sg == 0, true, fus, fus);
                commit(bar_full + 8 * (FULL_F0 + f));
                if (++f == n_acc) f = 0;
            }
            if (dbg && lane == 0) {
                dbg[16 * blockIdx.x + 1] = w_ready;
                dbg[16 * blockIdx.x + 2] = w_weights;
                dbg[16 * blockIdx.x + 3] = clock64() - t_start;
                dbg[16 * blockIdx.x + 6] = t_issue;
                dbg[16 * blockIdx.x + 14] = clock64() - t_final0;     // final layer: whole phase of the MMA warp
                dbg[16 * blockIdx.x + 15] = w_rdyf;                   // ... of which waiting for a drained accumulator
            }
        }
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 4;
        const int q = ew & 3;                    // TMEM lane quadrant (== warp % 4)
        const int cgp = ew >> 2;                 // 32-column group inside a region half
        const int r = 32 * q + lane;             // row inside the tile
        const int grow = row0 + r;
        const bool row_ok = grow < g.rows;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
        uint32_t ph_full = 0, ph_free1 = 0, ph_pfull = 0;
        long long w_full = 0, t_res0 = 0, t_res1 = 0, t_relu = 0, t_fin = 0, t_mark = 0, t_f1 = 0, t_f2 = 0, t_f3 = 0, t_m2 = 0;
        const bool dbg_me = dbg && ew == 0 && lane == 0;
        const long long e_start = dbg ? clock64() : 0;
        auto wait_full = [&](int id) {
            mbar_wait(bar_full + 8 * id, (ph_full >> id) & 1, g.err, 4,
                      (dbg && ew == 0 && lane == 0) ? &w_full : nullptr);
            ph_full ^= 1u << id;
            tc_fence_after();
        };
        auto signal_rdy = [&](int id) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rdy + 8 * id);
        };
        auto wait_pset = [&](int j) {
            mbar_wait(bar_pfull + 8 * (j & 1), (ph_pfull >> (j & 1)) & 1, g.err, 7);
            ph_pfull ^= 1u << (j & 1);
        };
        auto release_pset = [&](int j) {
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_pempty + 8 * (j & 1));
        };
        uint32_t v[16];
        float u0[32];                            // residual stream, column half 0 (half 1 lives in shared memory)
        // FP16 range guard: running maximum (per 16-bit half) of every operand this thread packed.  An activation
        // beyond 65504 packs to +inf (0x7c00), a NaN to 0x7fff; either raises the caller's flag at the end of the
        // H-wide stages, so an overflow can never be laundered through a later ReLU into finite-looking numbers.
        uint32_t hmax = 0;

        if (cgp < NH / 32) {   // warps of the H-wide stages (all of them at H = 256)
        // ---- features -> R1 (A operand of GEMM0), piece by piece ----
        for (int p = 0; p < g.n_pieces; ++p) {
            if (p > 0) {
                mbar_wait(bar_free1, ph_free1, g.err, 5);
                ph_free1 ^= 1;
                tc_fence_after();
            }
            const int kcols = min(H, g.Kp0 - p * H);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int col = hf * NH + cgp * 32;              // column inside R1
                if (col < kcols) {
                    const int k0 = p * H + col;
                    if (g.a0_tiled) {
                        // row-tiled features (a0_tiled): quad k / 4 of the tile's 128 rows is 2 KB contiguous
                        const float4* src = reinterpret_cast<const float4*>(g.A0 + (size_t)step * g.a0_stride) +
                                            ((size_t)tile * (size_t)((g.K0 + 3) >> 2) + (size_t)(k0 >> 2)) * 128 + r;
                        float4 q4[8];
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            q4[i4] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (row_ok && k0 + 4 * i4 < g.K0) q4[i4] = __ldg(src + (size_t)i4 * 128);
                        }
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                            for (int i4 = 0; i4 < 4; ++i4) {
                                const float4 t = q4[4 * sub + i4];
                                const int k = k0 + 16 * sub + 4 * i4;     // K0 = 2N is even: a quad holds 4 or 2 features
                                v[4 * i4] = to_tf32(t.x);
                                v[4 * i4 + 1] = to_tf32(t.y);
                                v[4 * i4 + 2] = to_tf32(k + 2 < g.K0 ? t.z : 0.f);
                                v[4 * i4 + 3] = to_tf32(k + 3 < g.K0 ? t.w : 0.f);
                            }
                            tc_st16(lane_addr + H + col + 16 * sub, v);
                        }
                    } else {                                     // caller's row-major matrix (single-layer entry points)
                        const float* src = g.A0 + (size_t)grow * g.K0 + k0;
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float x = 0.f;
                                if (row_ok && k0 + 16 * sub + i < g.K0) x = __ldg(src + 16 * sub + i);
                                v[i] = to_tf32(x);
                            }
                            tc_st16(lane_addr + H + col + 16 * sub, v);
                        }
                    }
                }
                tc_wait_st();
                signal_rdy(RDY_R1H0 + hf);
            }
        }
        if (FS_TC_TIMERS == 2 && dbg_me) t_f1 = clock64() - e_start;   // timeline: features packed
        // ---- residual-stream step on R0 (u = h - c, c = biases so far, folded at pack time):
        //      u (+)= D ; operand = relu(s u + o')   (or u itself in front of the final layer) ----
#define FS_EPI_RESIDUAL(HF, PRM, HAS_NEXT, INIT)                                                               \
    {                                                                                                          \
        const int col = (HF) * NH + cgp * 32;                                                                  \
        _Pragma("unroll") for (int sub = 0; sub < 2; ++sub) {                                                  \
            tc_ld16(lane_addr + col + 16 * sub, v);                                                            \
            float4 hh[4];                                                                                      \
            if ((HF) == 1 && !(INIT)) {                                                                        \
                _Pragma("unroll") for (int i4 = 0; i4 < 4; ++i4)                                               \
                    hh[i4] = us4[(size_t)(cgp * 8 + sub * 4 + i4) * 128 + r];                                  \
            }                                                                                                  \
            tc_wait_ld();                                                                                      \
            _Pragma("unroll") for (int i4 = 0; i4 < 4; ++i4) {                                                 \
                unsigned long long ua = pk2(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1]));       \
                unsigned long long ub = pk2(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3]));   \
                float4 uu;                                                                                     \
                if ((HF) == 0) {                                                                               \
                    if (!(INIT)) {                                                                             \
                        ua = add2(ua, pk2(u0[16 * sub + 4 * i4], u0[16 * sub + 4 * i4 + 1]));                  \
                        ub = add2(ub, pk2(u0[16 * sub + 4 * i4 + 2], u0[16 * sub + 4 * i4 + 3]));              \
                    }                                                                                          \
                    upk2(ua, uu.x, uu.y); upk2(ub, uu.z, uu.w);                                                \
                    u0[16 * sub + 4 * i4] = uu.x; u0[16 * sub + 4 * i4 + 1] = uu.y;                            \
                    u0[16 * sub + 4 * i4 + 2] = uu.z; u0[16 * sub + 4 * i4 + 3] = uu.w;                        \
                } else {                                                                                       \
                    if (!(INIT)) {                                                                             \
                        ua = add2(ua, pk2(hh[i4].x, hh[i4].y)); ub = add2(ub, pk2(hh[i4].z, hh[i4].w));        \
                    }                                                                                          \
                    upk2(ua, uu.x, uu.y); upk2(ub, uu.z, uu.w);                                                \
                    hh[i4] = uu;   /* stored after the loop: a store here would fence the parameter loads below */ \
                }                                                                                              \
                if (HAS_NEXT) {                                                                                \
                    const float4 sc = *reinterpret_cast<const float4*>((PRM) + H + col + 16 * sub + 4 * i4);   \
                    const float4 of = *reinterpret_cast<const float4*>((PRM) + 2 * H + col + 16 * sub + 4 * i4); \
                    ua = fma2(ua, pk2(sc.x, sc.y), pk2(of.x, of.y));                                           \
                    ub = fma2(ub, pk2(sc.z, sc.w), pk2(of.z, of.w));                                           \
                    upk2(ua, uu.x, uu.y); upk2(ub, uu.z, uu.w);                                                \
                    v[2 * i4] = pack_relu_f16x2(uu.x, uu.y);                                                   \
                    v[2 * i4 + 1] = pack_relu_f16x2(uu.z, uu.w);                                               \
                } else if (f16mode) {                                                                          \
                    v[2 * i4] = pack_f16x2(uu.x, uu.y); v[2 * i4 + 1] = pack_f16x2(uu.z, uu.w);                \
                } else {                                                                                       \
                    v[4 * i4] = round_tf32(uu.x); v[4 * i4 + 1] = round_tf32(uu.y);                            \
                    v[4 * i4 + 2] = round_tf32(uu.z); v[4 * i4 + 3] = round_tf32(uu.w);                        \
                }                                                                                              \
            }                                                                                                  \
            if ((HF) == 1) {                                                                                   \
                _Pragma("unroll") for (int i4 = 0; i4 < 4; ++i4)                                               \
                    us4[(size_t)(cgp * 8 + sub * 4 + i4) * 128 + r] = hh[i4];                                  \
            }                                                                                                  \
            /* FP16 range guard: largest packed half so far (ReLU output: sign bit clear; u itself: masked) */ \
            if (HAS_NEXT) {                                                                                    \
                _Pragma("unroll") for (int i = 0; i < 8; i += 2) hmax = __vimax3_u16x2(hmax, v[i], v[i + 1]);  \
            } else if (f16mode) {                                                                              \
                _Pragma("unroll") for (int i = 0; i < 8; i += 2)                                               \
                    hmax = __vimax3_u16x2(hmax, v[i] & 0x7fff7fffu, v[i + 1] & 0x7fff7fffu);                   \
            }                                                                                                  \
            /* packed halves go to the first 16 of this thread's own 32 columns; the theta path keeps TF32 u */ \
            /* the fused final layer's operand (H = 256): compact, features 2j, 2j+1 in column H + j (see TcCfg) */ \
            if (HAS_NEXT) tc_st8(lane_addr + col + 8 * sub, v);                                                \
            else if (f16mode) tc_st8(lane_addr + (S::COMPACT ? H + (col >> 1) : col) + 8 * sub, v);            \
            else tc_st16(lane_addr + col + 16 * sub, v);                                                       \
        }                                                                                                      \
        tc_wait_st();                                                                                          \
        signal_rdy(RDY_R0H0 + (HF));                                                                           \
    }
        wait_pset(0);
        wait_full(FULL_R0H0);
        FS_EPI_RESIDUAL(0, pbuf, true, true)
        wait_full(FULL_R0H1);
        FS_EPI_RESIDUAL(1, pbuf, true, true)
        if (FS_TC_TIMERS == 2 && dbg_me) t_f2 = clock64() - e_start;   // timeline: first residual step done
        release_pset(0);
        // ---- residual blocks ----
        for (int b = 0; b < g.n_blocks; ++b) {
            const float* prm = pbuf + (size_t)((b + 1) & 1) * S::PSET_FLOATS;
            wait_pset(b + 1);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                  // relu(t + b0') in place in R1
                wait_full(FULL_R1H0 + hf);
                if (dbg_me) t_mark = clock64();
                const int col = hf * NH + cgp * 32;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    tc_ld16(lane_addr + H + col + 16 * sub, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 bb = *reinterpret_cast<const float4*>(prm + col + 16 * sub + 4 * i4);
                        const unsigned long long ta =
                            add2(pk2(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1])), pk2(bb.x, bb.y));
                        const unsigned long long tb =
                            add2(pk2(__uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3])), pk2(bb.z, bb.w));
                        float t0, t1, t2, t3;
                        upk2(ta, t0, t1);
                        upk2(tb, t2, t3);
                        v[2 * i4] = pack_relu_f16x2(t0, t1);
                        v[2 * i4 + 1] = pack_relu_f16x2(t2, t3);
                    }
#pragma unroll
                    for (int i = 0; i < 8; i += 2) hmax = __vimax3_u16x2(hmax, v[i], v[i + 1]);
                    tc_st8(lane_addr + H + col + 8 * sub, v);
                }
                tc_wait_st();
                signal_rdy(RDY_R1H0 + hf);
                if (dbg_me) t_relu += clock64() - t_mark;
            }
            const bool last = (b == g.n_blocks - 1);
            wait_full(FULL_R0H0);
            if (!last) {
                if (dbg_me) t_mark = clock64();
                FS_EPI_RESIDUAL(0, prm, true, false)
                if (dbg_me) t_res0 += clock64() - t_mark;
                wait_full(FULL_R0H1);
                if (dbg_me) t_mark = clock64();
                FS_EPI_RESIDUAL(1, prm, true, false)
                if (dbg_me) t_res1 += clock64() - t_mark;
            } else {
                FS_EPI_RESIDUAL(0, prm, false, false)
                wait_full(FULL_R0H1);
                FS_EPI_RESIDUAL(1, prm, false, false)
            }
            release_pset(b + 1);
        }
        if (FS_TC_TIMERS == 2 && dbg_me) t_f3 = clock64() - e_start;   // timeline: trunk done
        if (((hmax & 0xffffu) >= 0x7c00u || (hmax >> 16) >= 0x7c00u) && g.nan_flag) atomicOr(g.nan_flag, 2);
        }   // H-wide stages
#undef FS_EPI_RESIDUAL
        // ---- fused final layer: one chunk = the 3nb+1 spline parameters of ONE transformed coordinate for the 128
        //      rows of the tile, laid out [widths | pad to 32 | heights | pad to 64 | derivatives]; theta never leaves
        //      the SM.  The epilogue warps of a TMEM lane quadrant (they share an SM sub-partition) form two pairs:
        //      pair 0 (column groups 0,1) takes the even chunks / accumulator 0, pair 1 the odd chunks /
        //      accumulator 1, so every warp has two chunk periods of the tensor pipe for its serial
        //      softmax -> search -> evaluate chain.  In a pair, warp A owns the search axis (widths in the density
        //      direction, heights when sampling): softmax numerators and inclusive prefix sums in registers, bin
        //      search, the two derivatives of the bin; warp B owns the other axis, receives bin index, edge,
        //      width and derivatives through a per-row mailbox (one named barrier per chunk), evaluates the
        //      rational-quadratic formula, writes the output coordinate and accumulates the log-determinant.
        //      coupling.py:86-102 / 126-135, utils/splines.py:84-222. ----
        if (g.fused) {
            const int nb = g.nb, pair = cgp >> 1;
            const bool isA = (cgp & 1) == 0;
            const bool inv = (g.fused == 2);
            const int axis = isA ? (inv ? 1 : 0) : (inv ? 0 : 1);   // 0 = widths columns, 1 = heights columns
            const float c2 = g.inv_sqrt_h * 1.4426950408889634f, bound = g.bound, two_b = 2.0f * g.bound;
            const float* xrow_in = (lp ? ((step & 1) ? g.buf1 : g.buf0) : g.xin) + (size_t)grow * g.D;
            float* xrow_out = (lp ? ((step & 1) ? g.buf0 : g.buf1) : g.xout) + (size_t)grow * g.D;
            float* const ld_out = g.logdet ? g.logdet + (size_t)step * g.ld_stride : nullptr;
            // layer-parallel launch: the previous step must have written every coordinate of these 32 rows (both of its
            // pairs) before this warp reads its first one - and, the other way round, nothing of the previous step still
            // reads the buffer this step writes
            const bool publish = lp && step < lp - 1;
            if (lp && step > 0) {
                const int* fl = g.flags + (size_t)((step - 1) * g.tiles + tile) * 4 + q;
                int seen;
                unsigned spins = 0;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(fl) : "memory");
                    if (seen >= 2) break;
                    __nanosleep(400);
                    if (++spins > 8000000u) {
                        if (g.err) atomicExch(g.err, 11);
                        __trap();
                    }
                }
            }
            auto xload = [&](int col) {                        // another CTA of this launch may have written it: not .nc
                return lp ? __ldcg(xrow_in + col) : __ldg(xrow_in + col);
            };
            const float gnum = 1.0f - kMinW * (float)nb;       // kMinW == kMinH
            const float rgnum = 1.0f / gnum, r2b = 1.0f / two_b;
            // The u buffer is dead: thread-private 128-byte rows for values a later dynamic index picks from
            // (float4 slot i4 at i4 ^ (lane & 7): conflict-free stores).  Mailbox [pair group][parity][field][lane].
            float* myrow = reinterpret_cast<float*>(smem + S::U_OFF) + (size_t)(ew * 32 + lane) * 32;
            float* mbq = reinterpret_cast<float*>(smem + S::MB_OFF) + (size_t)(q * 2 + pair) * 320 + lane;
            float acc_ld = 0.f;
            bool bad = false;
            auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q + 4 * pair) : "memory"); };
            auto pick = [&](int i) { return myrow[(((i >> 2) ^ (lane & 7)) << 2) | (i & 3)]; };
            // The layer input of this row for the coordinate of chunk c is a dependent load (column table -> x) of two
            // L2 latencies; issued at the top of chunk c it would stall the in-order warp for ~1 k clk before it even
            // looks at the accumulator.  Both loads run one own-chunk ahead instead: x of chunk c + 2 is requested while
            // chunk c is processed, its column index one chunk earlier still.
            int col_n2 = (pair + 2 < g.N) ? __ldg(g.xc_in + pair + 2) : 0;
            float x_nxt = (row_ok && pair < g.N) ? xload(__ldg(g.xc_in + pair)) : 0.f;
            // chunk c = 3 k + a lives in accumulator a; the k-th completion of FULL_F[a] has parity k & 1.  The pair sees
            // only every other completion of a barrier, but the one before (chunk c - 3, the other pair's) is older
            // than chunk c - 2, which this pair has already consumed: the parity wait cannot alias.
            int acc_a = pair, acc_k = 0;
            for (int c = pair; c < g.N; c += 2) {
                const uint32_t fcol = acc_a == 0 ? S::FINF0 : (acc_a == 1 ? S::FINF1 : S::FINF2);
                float* mb = mbq + ((c >> 1) & 1) * 160;
                const float x = x_nxt;
                if (c + 2 < g.N) x_nxt = row_ok ? xload(col_n2) : 0.f;
                if (c + 4 < g.N) col_n2 = __ldg(g.xc_in + c + 4);
                const int col_out = isA ? 0 : __ldg(g.xc_out + c);   // requested now, needed at the end of the chunk
                const int bslot = c % S::BIAS_SLOTS;
                const float* bch = reinterpret_cast<const float*>(smem + S::BIAS_OFF) + bslot * 128;
                mbar_wait(bar_bfull + 8 * bslot, (c / S::BIAS_SLOTS) & 1, g.err, 9);
                mbar_wait(bar_full + 8 * (FULL_F0 + acc_a), acc_k & 1, g.err, 4, dbg_me ? &w_full : nullptr);
                tc_fence_after();
                if (dbg_me) t_mark = clock64();
                float e[32];
                float d32 = 0.f;
                if (isA) {                                          // derivatives first, parked right away
                    tc_ld16(lane_addr + fcol + 64, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4)
                        reinterpret_cast<float4*>(myrow)[i4 ^ (lane & 7)] =
                            make_float4(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1]),
                                        __uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3]));
                    tc_ld16(lane_addr + fcol + 80, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4)
                        reinterpret_cast<float4*>(myrow)[(4 + i4) ^ (lane & 7)] =
                            make_float4(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1]),
                                        __uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3]));
                    tc_ld16(lane_addr + fcol + 96, v);              // derivative 32 sits in column 96
                    tc_wait_ld();
                    d32 = __uint_as_float(v[0]);
                }
                tc_ld16(lane_addr + fcol + 32 * axis, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) e[i] = __uint_as_float(v[i]);
                tc_ld16(lane_addr + fcol + 32 * axis + 16, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) e[16 + i] = __uint_as_float(v[i]);
                signal_rdy(RDY_F0 + acc_a);                         // accumulator drained
                if (acc_a == 0) acc_a = 2; else { --acc_a; ++acc_k; }   // chunk c + 2
                if (dbg_me) { t_m2 = clock64(); if (FS_TC_TIMERS != 2) t_f1 += t_m2 - t_mark; }
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 bb = reinterpret_cast<const float4*>(bch + 32 * axis)[i4];   // broadcast LDS.128
                    upk2(add2(pk2(e[4 * i4], e[4 * i4 + 1]), pk2(bb.x, bb.y)), e[4 * i4], e[4 * i4 + 1]);
                    upk2(add2(pk2(e[4 * i4 + 2], e[4 * i4 + 3]), pk2(bb.z, bb.w)), e[4 * i4 + 2], e[4 * i4 + 3]);
                }
                if (!isA) {                                         // warp B is done with the bias; warp A after its picks
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_bempty + 8 * bslot);
                }
                // softmax numerators and inclusive prefix sums; pad columns (>= nb) carry a bias of -3e38 (pack time)
                // and contribute exp(-inf) = 0
                float m8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) m8[i] = fmaxf(fmaxf(e[i], e[i + 8]), fmaxf(e[i + 16], e[i + 24]));
                const float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                      fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                const float off = -m * c2;
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float ex;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(__fmaf_rn(e[i], c2, off)));
                    sum += ex;
                    e[i] = sum;
                }
                const float gs = gnum * __frcp_rn(sum);
                if (dbg_me) { const long long tt = clock64(); if (FS_TC_TIMERS != 2) t_f2 += tt - t_m2; t_m2 = tt; }
                if (isA) {                                          // last knot <= x (utils/splines.py:11-13)
                    // x >= knot_i = 2b (gs S[i-1] + min i) - b   <=>   S[i-1] <= (t - min i) / gs,  t = (x + b) / 2b
                    const float rg = sum * rgnum;
                    const float t0 = (x + bound) * r2b * rg, dt = -kMinW * rg;
                    // Two-level search in registers without divergence (the lanes of a warp hold different rows).  The knot
                    // conditions are monotone (true up to the bin): level 1 picks the block of 8 bins from the knots 8, 16,
                    // 24; level 2 selects that block's nine prefix sums S[8 blk - 1 .. 8 blk + 7] with a two-predicate
                    // select tree (static register indices only), counts the hits inside it and selects the bin's two sums.
                    auto hit = [&](int i, float s_prev) { return i < nb && s_prev <= __fmaf_rn((float)i, dt, t0); };
                    const bool h8 = hit(8, e[7]), h16 = hit(16, e[15]), h24 = hit(24, e[23]);
                    const int blk8 = (h8 ? 1 : 0) + (h16 ? 1 : 0) + (h24 ? 1 : 0);
                    const bool b0 = (blk8 & 1) != 0, b1 = (blk8 & 2) != 0;
                    float w9[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        const float lo0 = (i == 0) ? 0.f : e[i - 1];      // block 0: S[-1] = 0
                        const float lo = b0 ? e[8 + i - 1] : lo0;
                        const float hi = b0 ? e[24 + i - 1] : e[16 + i - 1];
                        w9[i] = b1 ? hi : lo;
                    }
                    const int base8 = 8 * blk8;
                    const float t8 = __fmaf_rn((float)base8, dt, t0);
                    int cnt = 0;
#pragma unroll
                    for (int i = 1; i < 8; ++i)
                        cnt += (base8 + i < nb && w9[i] <= __fmaf_rn((float)i, dt, t8)) ? 1 : 0;
                    const int sel = base8 + cnt;
                    // S[sel - 1], S[sel] = w9[cnt], w9[cnt + 1]: three-level select trees on the bits of cnt
                    const bool c0 = (cnt & 1) != 0, c1 = (cnt & 2) != 0, c2b = (cnt & 4) != 0;
                    const float p01 = c0 ? w9[1] : w9[0], p23 = c0 ? w9[3] : w9[2], p45 = c0 ? w9[5] : w9[4],
                                p67 = c0 ? w9[7] : w9[6];
                    const float q12 = c0 ? w9[2] : w9[1], q34 = c0 ? w9[4] : w9[3], q56 = c0 ? w9[6] : w9[5],
                                q78 = c0 ? w9[8] : w9[7];
                    const float s0 = c2b ? (c1 ? p67 : p45) : (c1 ? p23 : p01);
                    const float s1 = c2b ? (c1 ? q78 : q56) : (c1 ? q34 : q12);
                    const float bd0 = bch[64 + sel], bd1 = bch[65 + sel];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_bempty + 8 * bslot);
                    const float left = __fmaf_rn(two_b, __fmaf_rn(gs, s0, kMinW * (float)sel), -bound);
                    const float right =
                        (sel == nb - 1) ? bound : __fmaf_rn(two_b, __fmaf_rn(gs, s1, kMinW * (float)(sel + 1)), -bound);
                    const float dk = pick(sel) + bd0;
                    const float dk1 = (sel < 31 ? pick(sel + 1) : d32) + bd1;
                    mb[0] = __int_as_float(sel);
                    mb[32] = left;
                    mb[64] = right - left;
                    mb[96] = kMinD + softplus_fast(dk);
                    mb[128] = kMinD + softplus_fast(dk1);
                    if (dbg_me && FS_TC_TIMERS != 2) t_f3 += clock64() - t_m2;
                    pair_sync();
                } else {
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4)
                        reinterpret_cast<float4*>(myrow)[i4 ^ (lane & 7)] =
                            make_float4(e[4 * i4], e[4 * i4 + 1], e[4 * i4 + 2], e[4 * i4 + 3]);
                    pair_sync();
                    const int sel = __float_as_int(mb[0]);
                    const float s1 = pick(sel), s0 = sel ? pick(sel - 1) : 0.f;
                    const float left = __fmaf_rn(two_b, __fmaf_rn(gs, s0, kMinW * (float)sel), -bound);
                    const float right =
                        (sel == nb - 1) ? bound : __fmaf_rn(two_b, __fmaf_rn(gs, s1, kMinW * (float)(sel + 1)), -bound);
                    const float aL = mb[32], aW = mb[64];
                    float y = x, ld = 0.f;
                    if (x >= -bound && x <= bound) {
                        if (inv) rq_eval_fast(x, left, right - left, aL, aW, mb[96], mb[128], true, y, ld);
                        else rq_eval_fast(x, aL, aW, left, right - left, mb[96], mb[128], false, y, ld);
                    }
                    if (row_ok) {
                        xrow_out[col_out] = y;
                        acc_ld += ld;
                        bad = bad || (y != y) || (ld != ld);
                    }
                }
                if (dbg_me) t_fin += clock64() - t_mark;
            }
            if (!isA && publish) {                                  // every coordinate of this pair is written: release the
                __threadfence();                                    // next step's warps of this lane quadrant
                __syncwarp();
                if (lane == 0)
                    asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(g.flags + (size_t)(step * g.tiles + tile) * 4 + q)
                                 : "memory");
            }
            if (!isA) {                                             // the two B warps of a quadrant: fixed-order sum
                if (pair == 1) mbq[0] = acc_ld;                     // pair 1's mailbox is idle now
                asm volatile("bar.sync %0, 64;" ::"r"(9 + q) : "memory");
                if (pair == 0 && row_ok && ld_out) ld_out[grow] += acc_ld + mbq[320];
            }
            if (bad && g.nan_flag) atomicOr(g.nan_flag, 1);
        } else {
        // ---- final layer: theta chunk = D + b_final' -> global ----
        // A thread holds 32 columns of ONE row, so storing straight from registers would touch 32 different lines
        // per instruction.  Each warp transposes its 32 x 32 block through a private 4 KB patch of the (now dead)
        // u buffer -- float4 slot c of row rr at rr*8 + (c ^ (rr & 7)), conflict-free both ways -- and stores
        // 4 rows x 128 contiguous bytes per instruction.
        const bool vec_ok = (g.NP & 3) == 0;
        constexpr int CPT = (S::FCH / 32) * 4 / EPI_WARPS;     // 32-column chunks of a final accumulator per thread: 1 / 2
        float4* patch = reinterpret_cast<float4*>(smem + S::U_OFF) + ew * 256;
        const int slot = lane & 7, rsub = lane >> 3;
        for (int c = 0; c < g.n_chunks; ++c) {
            const int f = c & 1;
            wait_full(FULL_F0 + f);
            if (dbg_me) t_mark = clock64();
            const uint32_t fcol = f ? S::FIN1 : S::FIN0;
#pragma unroll
            for (int cc = 0; cc < CPT; ++cc) {
                const int col = (cgp + cc * (EPI_WARPS / 4)) * 32;     // column inside the chunk
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    tc_ld16(lane_addr + fcol + col + 16 * sub, v);
                    tc_wait_ld();
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4)
                        patch[lane * 8 + ((sub * 4 + i4) ^ (lane & 7))] =
                            make_float4(__uint_as_float(v[4 * i4]), __uint_as_float(v[4 * i4 + 1]),
                                        __uint_as_float(v[4 * i4 + 2]), __uint_as_float(v[4 * i4 + 3]));
                }
                if (cc == CPT - 1) signal_rdy(RDY_F0 + f);             // all of this thread's TMEM reads are done
                __syncwarp();
                const int ocol = c * S::FCH + col + 4 * slot;          // this lane's 4 columns
                if (vec_ok && c * S::FCH + col + 32 <= g.NP) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(L.b_final + ocol));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + rsub;
                        const int gr = row0 + 32 * q + rr;
                        const float4 t = patch[rr * 8 + (slot ^ (rr & 7))];
                        if (gr < g.rows)
                            *reinterpret_cast<float4*>(g.theta + (size_t)gr * g.NP + ocol) =
                                make_float4(t.x + bb.x, t.y + bb.y, t.z + bb.z, t.w + bb.w);
                    }
                } else {
#pragma unroll 1
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + rsub;
                        const int gr = row0 + 32 * q + rr;
                        const float4 t = patch[rr * 8 + (slot ^ (rr & 7))];
                        const float tv[4] = {t.x, t.y, t.z, t.w};
                        if (gr < g.rows)
                            for (int e = 0; e < 4; ++e)
                                if (ocol + e < g.NP) g.theta[(size_t)gr * g.NP + ocol + e] = tv[e] + __ldg(L.b_final + ocol + e);
                    }
                }
                __syncwarp();
            }
            if (dbg_me) t_fin += clock64() - t_mark;
        }
        }   // !fused
        if (dbg_me) {
            dbg[16 * blockIdx.x + 4] = w_full;
            dbg[16 * blockIdx.x + 5] = clock64() - e_start;
            dbg[16 * blockIdx.x + 7] = t_res0;
            dbg[16 * blockIdx.x + 8] = t_res1;
            dbg[16 * blockIdx.x + 9] = t_relu;
            dbg[16 * blockIdx.x + 10] = t_fin;
            dbg[16 * blockIdx.x + 11] = t_f1;
            dbg[16 * blockIdx.x + 12] = t_f2;
            dbg[16 * blockIdx.x + 13] = t_f3;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(S::TMEM_COLS));
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

// appends the tile rows [n0, n0+ROWS) x cols [k0, k0+32) of W [n_rows, n_cols] (row-major) in the
// SWIZZLE_128B K-major shared-memory image: row i at i*128 bytes, its 16-byte chunk c at (c ^ (i & 7)).
static float half_round(float x) { return __half2float(__float2half_rn(x)); }

// [ROWS x 64] FP16 tile, K-major, 128-byte swizzle (16-byte chunk c of row i at chunk c ^ (i & 7))
static void append_tile16(std::vector<uint16_t>& out, const float* W, int n_rows, int n_cols, int k0, int ROWS) {
    const size_t base = out.size();
    out.resize(base + (size_t)ROWS * 64, 0);
    for (int i = 0; i < ROWS && i < n_rows; ++i)
        for (int c = 0; c < 8; ++c) {
            const int pc = c ^ (i & 7);
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + 8 * c + e;
                if (k < n_cols) {
                    const __half h = __float2half_rn(W[(size_t)i * n_cols + k]);
                    uint16_t bits;
                    memcpy(&bits, &h, 2);
                    out[base + (size_t)i * 64 + pc * 8 + e] = bits;
                }
            }
        }
}

static void append_tile(std::vector<float>& out, const float* W, int n_rows, int n_cols, int n0, int k0, int ROWS) {
    const size_t base = out.size();
    out.resize(base + (size_t)ROWS * 32, 0.f);
    for (int i = 0; i < ROWS; ++i) {
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
    const int FCH = 128;                                   // final-layer chunk width (TcCfg::FCH)
    P->n_chunks = (NP + FCH - 1) / FCH;
    // fused spline epilogue: derivatives 0..nb must fit columns 64..127
    P->chn = (f->nb <= 32) ? (64 + f->nb + 1 + 15) / 16 * 16 : 0;
    P->smem_bytes = (H == 256 ? TcCfg<256>::TOTAL : TcCfg<128>::TOTAL) + 1024;
    if ((int)P->smem_bytes > smem_max) { delete P; return FS_OK; }
    const int KT = H / TC_KB;
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
        std::vector<float> fw, fbias;                       // final layer in parameter-major row order
        permute_final(p, f->N, f->P, H, fw, fbias);
        // weight stream in consumption order, one stage = H*128 bytes:
        //   block GEMMs: k-tile kt of all H output rows;  final layer: KPS consecutive k-tiles of a 128-row chunk
        const int KPS = (H * 128) / (FCH * 128);
        std::vector<float> stream;
        stream.reserve((size_t)H * 32 * (P->Kp0 / TC_KB + (size_t)nB * 2 * KT) + (size_t)P->n_chunks * FCH * H);
        for (int pc = 0; pc < P->n_pieces; ++pc) {                       // GEMM0
            const int kcols = std::min(H, P->Kp0 - pc * H);
            for (int kt = 0; kt < kcols / TC_KB; ++kt) append_tile(stream, p->init_w, H, K0, 0, pc * H + kt * TC_KB, H);
        }
        // Residual-block GEMMs: FP16 tiles of 64 k.  H = 256 (split schedule, see gemm_split): the two K-lo tiles of all
        // H rows, then (rows lo, K hi) and (rows hi, K hi) as stages of KPS tiles of a 128-row chunk.  H = 128: the two
        // tiles of all rows.
        auto append16 = [&](const float* W, int rows, int k0) {
            std::vector<uint16_t> t16;
            append_tile16(t16, W, rows, H, k0, rows);
            const size_t base = stream.size();
            stream.resize(base + t16.size() / 2);
            memcpy(&stream[base], t16.data(), t16.size() * 2);
        };
        auto append_gemm = [&](const float* W) {
            const int KT16 = H / 64;
            if (H == 256) {
                for (int kt = 0; kt < KT16 / 2; ++kt) append16(W, H, kt * 64);
                for (int nh = 0; nh < 2; ++nh)
                    for (int kt = KT16 / 2; kt < KT16; ++kt) append16(W + (size_t)nh * (H / 2) * H, H / 2, kt * 64);
            } else {
                for (int kt = 0; kt < KT16; ++kt) append16(W, H, kt * 64);
            }
        };
        for (int b = 0; b < nB; ++b) {
            append_gemm(&w0[(size_t)b * H * H]);                               // linear 0 (BN1 folded)
            append_gemm(p->lin_w + ((size_t)b * 2 + 1) * H * H);               // linear 1
        }
        for (int c = 0; c < P->n_chunks; ++c)                            // final layer
            for (int sg = 0; sg < KT / KPS; ++sg)
                for (int kk = 0; kk < KPS; ++kk)
                    append_tile(stream, fw.data(), NP, H, c * FCH, (sg * KPS + kk) * TC_KB, FCH);
        P->tiles_per_layer = stream.size() / ((size_t)H * 32);
        // Fused-spline final layer (H = 256, nb <= 32): chunk j = the parameters of transformed coordinate j in the
        // order [widths | pad to 32 | heights | pad to 64 | derivatives | pad to chn]; rows come from the
        // reference's coordinate-major final layer (row j*P + index, coupling.py:166).
        std::vector<float> bfused;
        std::vector<uint16_t> fstream16;
        std::vector<int> frow;                                  // chunk column -> parameter index (or -1)
        if (P->chn) {
            const int chn = P->chn, nb = f->nb, Pp = f->P;
            frow.assign(chn, -1);
            for (int k = 0; k < nb; ++k) { frow[k] = k; frow[32 + k] = nb + k; }
            for (int k = 0; k <= nb; ++k) frow[64 + k] = 2 * nb + k;
            std::vector<float> wc((size_t)chn * H);
            fstream16.reserve((size_t)f->N * chn * H);
            for (int j = 0; j < f->N; ++j) {
                std::fill(wc.begin(), wc.end(), 0.f);
                for (int cidx = 0; cidx < chn; ++cidx)
                    if (frow[cidx] >= 0)
                        memcpy(&wc[(size_t)cidx * H], p->final_w + ((size_t)j * Pp + frow[cidx]) * H, sizeof(float) * H);
                for (int kt = 0; kt < H / 64; ++kt) append_tile16(fstream16, wc.data(), chn, H, kt * 64, chn);
            }
        }
        TcLayer& L = P->layers[li];
        // Bias folding: the kernel carries u = h - c (c = b_init + sum of linear-1 biases so far):
        //   relu(s h + o) = relu(s u + (o + s c)),   W_f h + b_f = W_f u + (b_f + W_f c)
        std::vector<double> c(p->init_b, p->init_b + H);
        std::vector<float> o0f((size_t)nB * H);
        for (int b = 0; b < nB; ++b) {
            for (int k = 0; k < H; ++k) {
                o0f[(size_t)b * H + k] = (float)((double)o0[(size_t)b * H + k] + (double)s0[(size_t)b * H + k] * c[k]);
                c[k] += (double)p->lin_b[((size_t)b * 2 + 1) * H + k];
            }
        }
        std::vector<float> bfin((size_t)P->n_chunks * FCH, 0.f);
        for (int n = 0; n < NP; ++n) {
            double acc = (double)fbias[n];
            const float* wr = fw.data() + (size_t)n * H;
            for (int k = 0; k < H; ++k) acc += (double)tf32_round(wr[k]) * c[k];
            bfin[n] = (float)acc;
        }
        if (P->chn) {
            const int chn = P->chn, Pp = f->P;
            bfused.assign((size_t)f->N * chn + 32, 0.f);
            for (int j = 0; j < f->N; ++j)
                for (int cidx = 0; cidx < chn; ++cidx) {
                    if (frow[cidx] < 0) continue;
                    const size_t row = (size_t)j * Pp + frow[cidx];
                    double acc = (double)p->final_b[row];
                    const float* wr = p->final_w + row * H;
                    for (int k = 0; k < H; ++k) acc += (double)half_round(wr[k]) * c[k];
                    bfused[(size_t)j * chn + cidx] = (float)acc;
                }
            // pad columns of the two softmax groups: zero weights and a bias of -3e38, so their numerators are
            // exp(-inf) = 0 without any masking in the epilogue
            for (int j = 0; j < f->N; ++j)
                for (int k = f->nb; k < 32; ++k) {
                    bfused[(size_t)j * chn + k] = -3.0e38f;
                    bfused[(size_t)j * chn + 32 + k] = -3.0e38f;
                }
        }
        std::vector<float> psets((size_t)(nB + 1) * 3 * H, 0.f);
        for (int j = 0; j <= nB; ++j) {
            float* ps = &psets[(size_t)j * 3 * H];
            if (j >= 1) memcpy(ps, &b0[(size_t)(j - 1) * H], sizeof(float) * H);
            if (j < nB) {
                memcpy(ps + H, &s0[(size_t)j * H], sizeof(float) * H);
                memcpy(ps + 2 * H, &o0f[(size_t)j * H], sizeof(float) * H);
            }
        }
        int r = tc_upload(f, stream, &L.wstream);
        if (!r) r = tc_upload(f, psets, &L.psets);
        if (!r) r = tc_upload(f, s0, &L.bn0_s);
        if (!r) r = tc_upload(f, o0f, &L.bn0_o);
        if (!r) r = tc_upload(f, b0, &L.b0);
        if (!r) r = tc_upload(f, bfin, &L.b_final);
        L.wfused = nullptr;
        L.b_fused = nullptr;
        if (!r && P->chn) r = tc_upload(f, bfused, &L.b_fused);
        if (!r && P->chn) {
            void* d16 = nullptr;
            r = cuda_check(cudaMalloc(&d16, fstream16.size() * 2), "cudaMalloc");
            if (!r) {
                f->allocs.push_back(d16);
                r = cuda_check(cudaMemcpy(d16, fstream16.data(), fstream16.size() * 2, cudaMemcpyHostToDevice), "cudaMemcpy");
                L.wfused = d16;
            }
        }
        if (r) { delete P; return r; }
    }
    {   // columns read / written by the fused epilogue: density reads feature ft and writes (ft + D/2) % D (the roll,
        // coupling.py:100-101), sampling reads the rolled input (coupling.py:113-114) and writes ft
        std::vector<int> xc((size_t)4 * f->N, -1);
        for (int j = 0; j < f->N; ++j) {
            const int ft = d->transform_features[j], rolled = (ft + f->D / 2) % f->D;
            xc[j] = ft;
            xc[f->N + j] = rolled;
            xc[2 * f->N + j] = rolled;
            xc[3 * f->N + j] = ft;
        }
        if (int r = tc_upload(f, xc, &P->xcols)) { delete P; return r; }
        // Layer-parallel launches need the identity set to be closed under the roll by D/2 (then the conditioner inputs
        // of all layers follow from the unconditional splines alone) and disjoint from the transformed set.
        std::vector<char> is_id((size_t)f->D, 0);
        for (int j = 0; j < f->N; ++j) is_id[d->identity_features[j]] = 1;
        bool closed = P->chn > 0 && f->D == 2 * f->N;
        for (int j = 0; j < f->N && closed; ++j)
            closed = is_id[(d->identity_features[j] + f->D / 2) % f->D] && !is_id[d->transform_features[j]];
        P->lp_ok = closed;
    }
    {   // device copy of the per-layer pointer table (the buffers behind it are re-packed in place by fs_flow_update)
        void* dl = nullptr;
        if (cudaMalloc(&dl, sizeof(TcLayer) * P->layers.size()) != cudaSuccess) { delete P; return FS_ERR_CUDA; }
        f->allocs.push_back(dl);
        if (cudaMemcpy(dl, P->layers.data(), sizeof(TcLayer) * P->layers.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            delete P;
            return FS_ERR_CUDA;
        }
        P->layers_dev = (TcLayer*)dl;
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

// FS_TC_DEBUG=1: per-CTA wait-cycle counters of the last launch, readable through fs_tc_debug_read()
static long long* g_dbg = nullptr;
static int g_dbg_ctas = 0;
static long long* tc_debug_buffer(int ctas) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("FS_TC_DEBUG"); enabled = (e && e[0] == '1') ? 1 : 0; }
    if (!enabled) return nullptr;
    if (ctas > g_dbg_ctas) {
        if (g_dbg) cudaFree(g_dbg);
        cudaMalloc(&g_dbg, sizeof(long long) * 16 * ctas);
        g_dbg_ctas = ctas;
    }
    cudaMemset(g_dbg, 0, sizeof(long long) * 16 * ctas);
    return g_dbg;
}

static int tc_launch(fs_flow* f, int layer, const float* A0, bool tiled, int rows, float* theta, int fused,
                     const float* xin, float* xout, float* logdet, int* nan_flag, cudaStream_t s) {
    TcPack* P = (TcPack*)f->tc;
    if (!P) {
        set_error("tensor-core conditioner not available for this flow shape (H=%d, blocks=%d)", f->H, f->n_blocks);
        return FS_ERR_UNSUPPORTED;
    }
    if (fused && !P->chn) {
        set_error("fused spline epilogue not available for this flow shape (H=%d, bins=%d)", f->H, f->nb);
        return FS_ERR_UNSUPPORTED;
    }
    TcArgs g;
    g.A0 = A0;
    g.a0_tiled = tiled ? 1 : 0;
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
    g.fused = fused;
    g.N = f->N;
    g.D = f->D;
    g.nb = f->nb;
    g.chn = P->chn;
    g.bound = f->bound_f;
    g.inv_sqrt_h = f->inv_sqrt_h;
    g.xin = xin;
    g.xout = xout;
    g.logdet = logdet;
    g.xc_in = P->xcols + (size_t)(fused == 2 ? 2 : 0) * f->N;
    g.xc_out = g.xc_in + f->N;
    g.nan_flag = nan_flag;
    g.dbg = tc_debug_buffer((rows + 127) / 128);
    g.Ls = nullptr;
    g.lp_layers = g.lp_rev = 0;
    g.tiles = (rows + 127) / 128;
    g.a0_stride = g.ld_stride = 0;
    g.buf0 = g.buf1 = nullptr;
    g.flags = nullptr;
    const int grid = (rows + 127) / 128;
    if (P->H == 256)
        tc_conditioner_kernel<256><<<grid, TcCfg<256>::THREADS, P->smem_bytes, s>>>(g);
    else
        tc_conditioner_kernel<128><<<grid, TcCfg<128>::THREADS, P->smem_bytes, s>>>(g);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "tc_conditioner_kernel");
}

int tc_conditioner(fs_flow* f, int layer, const float* A0, bool tiled, int rows, float* theta, void*, size_t,
                   int* nan_flag, cudaStream_t s) {
    return tc_launch(f, layer, A0, tiled, rows, theta, 0, nullptr, nullptr, nullptr, nan_flag, s);
}

bool tc_has_fused(const fs_flow* f) { return f->tc && ((TcPack*)f->tc)->chn > 0; }

// conditioner + conditional spline of the transformed half in one kernel: direction 1 = density (coupling.py:86-102),
// 2 = sampling (coupling.py:126-135).  Reads the layer input xin, writes the transformed half of xout and
// accumulates the log-determinant.
int tc_conditioner_spline(fs_flow* f, int layer, const float* A0, bool tiled, int rows, int direction, const float* xin,
                          float* xout, float* logdet, int* nan_flag, cudaStream_t s) {
    return tc_launch(f, layer, A0, tiled, rows, nullptr, direction, xin, xout, logdet, nan_flag, s);
}

bool tc_layer_parallel_ok(const fs_flow* f) {
    return f->tc && ((TcPack*)f->tc)->lp_ok && ((TcPack*)f->tc)->chn > 0 && f->K >= 2 && !getenv("FS_NO_LP");
}

size_t tc_lp_flag_ints(const fs_flow* f, int rows) { return (size_t)f->K * (size_t)((rows + 127) / 128) * 4; }

// All K coupling layers of one pass in ONE launch of K x tiles CTAs (see TcArgs).  direction 1: density (layers
// K-1 .. 0), 2: sampling (0 .. K-1).  A0: K row-tiled feature matrices a0_stride floats apart, in step order; buf0 holds
// the input of step 0; ldp: K log-det partials ld_stride floats apart (zeroed by the caller, summed by finish_kernel);
// flags: tc_lp_flag_ints() ints, zeroed by the caller.
int tc_conditioner_spline_all(fs_flow* f, int direction, int rows, const float* A0, size_t a0_stride, float* buf0,
                              float* buf1, float* ldp, size_t ld_stride, int* flags, int* nan_flag, cudaStream_t s) {
    TcPack* P = (TcPack*)f->tc;
    if (!P || !P->chn || !P->lp_ok) {
        set_error("layer-parallel conditioner launch not available for this flow");
        return FS_ERR_UNSUPPORTED;
    }
    TcArgs g;
    g.A0 = A0;
    g.a0_tiled = 1;
    g.theta = nullptr;
    g.rows = rows;
    g.K0 = 2 * f->N;
    g.NP = f->N * f->P;
    g.Kp0 = P->Kp0;
    g.n_pieces = P->n_pieces;
    g.n_blocks = f->n_blocks;
    g.n_chunks = P->n_chunks;
    g.n_tiles = P->tiles_per_layer;
    g.L = P->layers[0];
    g.err = f->tc_err;
    g.fused = direction;
    g.N = f->N;
    g.D = f->D;
    g.nb = f->nb;
    g.chn = P->chn;
    g.bound = f->bound_f;
    g.inv_sqrt_h = f->inv_sqrt_h;
    g.xin = nullptr;
    g.xout = nullptr;
    g.logdet = ldp;
    g.xc_in = P->xcols + (size_t)(direction == 2 ? 2 : 0) * f->N;
    g.xc_out = g.xc_in + f->N;
    g.nan_flag = nan_flag;
    g.tiles = (rows + 127) / 128;
    g.dbg = tc_debug_buffer(g.tiles * f->K);
    g.Ls = P->layers_dev;
    g.lp_layers = f->K;
    g.lp_rev = direction == 1 ? 1 : 0;
    g.a0_stride = a0_stride;
    g.ld_stride = ld_stride;
    g.buf0 = buf0;
    g.buf1 = buf1;
    g.flags = flags;
    const int grid = g.tiles * f->K;
    if (P->H == 256)
        tc_conditioner_kernel<256><<<grid, TcCfg<256>::THREADS, P->smem_bytes, s>>>(g);
    else
        tc_conditioner_kernel<128><<<grid, TcCfg<128>::THREADS, P->smem_bytes, s>>>(g);
    fs::count_launch();
    return cuda_check(cudaGetLastError(), "tc_conditioner_kernel (layer-parallel)");
}

}  // namespace fs

// development aid: copies the wait-cycle counters of the last tensor-kernel launch (16 int64 per CTA:
// producer wait-empty, MMA wait-operand, MMA wait-weights, MMA total, epilogue wait-accumulator,
// epilogue total, 0, 0).  Returns the number of CTAs copied.
extern "C" int fs_tc_debug_read(long long* host, int max_ctas) {
    int n = fs::g_dbg_ctas < max_ctas ? fs::g_dbg_ctas : max_ctas;
    if (!fs::g_dbg || n <= 0) return 0;
    cudaDeviceSynchronize();
    cudaMemcpy(host, fs::g_dbg, sizeof(long long) * 16 * n, cudaMemcpyDeviceToHost);
    return n;
}
