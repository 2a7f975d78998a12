// sdn_conv_tc.cu -- 2-D convolution as an implicit GEMM on Blackwell's 5th-generation tensor cores, hand-written for
// sm_100a:  TMA tensor-map loads (cp.async.bulk.tensor, 128-byte swizzle) -> shared memory -> tcgen05.mma (bf16 x bf16
// -> fp32 in tensor memory) -> tcgen05.ld epilogue (per-channel scale / shift = folded BatchNorm or bias, optional
// residual add, optional ReLU) -> bf16 NHWC.
//
// It serves the convolution trunk of the image encoder (reference: geometric/derender3d/models/derenderer.py:25-27,
// torchvision resnet18: 3x3 / 1x1 convolutions, stride 1 and 2, 64..512 channels) and the textural branch's generator
// (reference: textural/models/networks.py:211-283: 7x7 stem, 3x3 stride-2 down-sampling, 3x3 ResnetBlocks at 1024
// channels, transposed convolutions).
//
// GEMM view:  D[M, N] = A[M, K] . B[N, K]^T,  M = output pixels (n, ho, wo), N = Cout, K = (kh, kw, cin).
//   * Activations are NHWC bf16.  An M tile is a PATCH of bn x bh x bw output pixels (bn*bh*bw = 128): for filter tap
//     (kh, kw) and 64 input channels its A operand is the input box of that patch shifted by the tap -- one 4-D TMA
//     box load {64 ch, bw, bh, bn} (traversal stride = the convolution stride) whose out-of-bounds elements TMA fills
//     with zeros: zero padding costs nothing and no im2col buffer exists.
//   * Weights are [Cout, KH*KW*Cin] bf16 (K contiguous): the K-major B operand, one 2-D TMA box {64, TN} per K block.
//   * Both land in the 128-byte-swizzled K-major layout (row r of the tile = 128 bytes at r*128, 16-byte chunks XOR-ed
//     with r mod 8) that the UMMA shared-memory descriptor names with layout_type = SWIZZLE_128B, SBO = 1024 B; a
//     K = 16 slice is selected by advancing the descriptor's start address by 32 bytes.
//   * Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane; also owns the TMEM
//     allocation), warps 2..5 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31 = tile rows).  A 6- or 8-stage ring of
//     (A | B) tiles is handed back and forth with mbarriers: `full` (TMA complete_tx) and `empty` (tcgen05.commit).
//   * PERSISTENT CTAs (one per SM) walk the (m tile, channel block) work list; the accumulator is double-buffered in
//     tensor memory (2 x TN columns) so the epilogue of one item overlaps the main loop of the next.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "sdn_b200.h"

namespace sdn_internal {     // sdn_raster.cu: launch counter + optional per-kernel event timing
void launch_begin(const char* name, cudaStream_t st);
int launch_end();
}  // namespace sdn_internal

namespace {

constexpr int TM = 128;                 // output pixels per tile (= TMEM lanes)
constexpr int TK = 64;                  // K block: 64 bf16 = 128 bytes = one swizzle row
// operand ring depth: 128-wide tiles run one persistent CTA per SM with ~190 KB of operands in flight; narrower tiles
// (N = 64 halves the work per tcgen05.mma, so one issuing thread cannot keep the tensor pipe busy) run two CTAs per SM
template <int TN> struct Ring { static constexpr int STAGES = TN == 128 ? 6 : 4; };      // 256: 4 x 48 KB      // TN <= 64: two CTAs per SM (two MMA issuers)
constexpr int A_BYTES = TM * TK * 2;    // 16 KB
constexpr int NTHREADS = 192;

struct ConvParams {
    int N, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
    int bw, bh, bn, tiles_w, tiles_h, tiles_n;
    __nv_bfloat16* out;                 // pixel (n, ho, wo) at out + n*osn + ho*osh + wo*osw (elements), Cout contiguous
    long osn, osh, osw;
    const float* scale;                 // [Cout] or null (= 1)
    const float* shift;                 // [Cout] or null (= 0)
    const __nv_bfloat16* residual;      // [N, Ho, Wo, Cout] or null
    int relu;                           // activation: 0 none, 1 ReLU, 2 tanh
    float* stats;                       // [N, Cout, 2] or null: per (image, channel) sum / sum of squares of the result
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp `SmemDescriptor`), K-major tile of 128-byte rows with the
// 128-byte swizzle: start address >> 4 in bits [0,14), LBO unused (0), SBO = 1024 B (one 8-row group) in bits [32,46),
// version 1 in bits [46,48), layout_type SWIZZLE_128B = 2 in bits [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr_bytes & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (`InstrDescriptor`): D = F32 (bit 4), A = BF16 (bits 7..9 = 1), B = BF16 (bits 10..12 = 1),
// both K-major, N >> 3 in bits [17,23), M >> 4 in bits [24,29)
template <int TN>
__device__ __forceinline__ constexpr uint32_t idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// CTA-pair (tcgen05 cta_group::2) forms: the load lands in THIS CTA's shared memory, its bytes are counted on the LEADER's
// barrier (same offset in the even CTA of the pair: the peer bit, bit 24 of a shared::cluster address, cleared)
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
// instruction descriptor of the pair form: M = 256 (128 rows per CTA), N = TN
template <int TN>
__device__ __forceinline__ constexpr uint32_t idesc_pair() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// PAIR = true: the CTA PAIR form (cluster of 2, tcgen05 cta_group::2).  The two CTAs of a cluster own two consecutive m tiles
// (M = 256 output pixels) and share one 256-wide block of output channels: each loads its own 128-pixel A tile and HALF of
// the weight tile (128 of the 256 output channels), the leader (cluster rank 0) issues ONE tcgen05.mma.cta_group::2 of
// M256 x N256 x K16 that reads both CTAs' shared memory and writes each CTA's 128 accumulator rows into its own tensor
// memory.  Per SM and K block that is 32 KB of operands for 128 x 256 x 64 MACs -- half the L2 -> SM bytes per MAC of the
// single-CTA 128 x 128 tile, which sits at the 64 B/clk port of the SM.  Barriers: the `full` barriers live in the leader
// (both CTAs' TMA loads complete_tx on it: .cta_group::2 with the peer bit of the barrier address cleared); `empty` and
// `tmem_full` are signalled in both CTAs by multicast commits; the follower's epilogue warps arrive on the leader's
// `tmem_empty` through mapa.
template <int TN, bool PAIR>
__global__ void __launch_bounds__(NTHREADS) k_conv_tc(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    extern __shared__ unsigned char smem_raw[];
    // the swizzle pattern is a function of the address bits [4,10): tiles must start on 1024-byte boundaries
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int STAGES = PAIR ? 6 : Ring<TN>::STAGES;
    constexpr int B_BYTES = (PAIR ? TN / 2 : TN) * TK * 2;      // PAIR: this CTA's half of the weight tile
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES], s_tmem_full[2], s_tmem_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_stat[2 * 256];                   // per-CTA (sums | sums of squares) of the current (image, channel block)
    constexpr uint32_t TNA = TN < 32 ? 32 : TN;         // accumulator columns per buffer (TMEM allocations are >= 32)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // PERSISTENT CTAs: work item w -> (m tile, output-channel block), m fastest, so the CTAs that run together share one
    // weight block and neighbouring input patches in L2.  The accumulator is double-buffered in tensor memory: the
    // epilogue of item i drains buffer i & 1 while the MMA warp already fills the other one for item i + 1.
    const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_n;
    uint32_t rank = 0;
    if constexpr (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tiles_mw = PAIR ? (tiles_m + 1) / 2 : tiles_m;             // m work units (pairs of tiles)
    const int total_work = tiles_mw * (p.Cout / TN);
    const int w_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, w_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int cblocks = p.Cin / TK;
    const int kblocks = p.KH * p.KW * cblocks;

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&s_tmem_full[s], 1); mbar_init(&s_tmem_empty[s], PAIR ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // the MMA warp owns the accumulator columns in tensor memory
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                         "r"(2u * TNA));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                         "r"(2u * TNA));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (PAIR) {       // both CTAs' barriers are initialised before either signals the other's
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
            int kbg = 0;                                     // K blocks issued by this CTA so far (ring position)
            for (int w = w_first; w < total_work; w += w_step) {
                int t = w % tiles_mw;
                const int co0 = (w / tiles_mw) * TN;
                if (PAIR) t = t * 2 + (int)rank;             // an odd tile count leaves the last follower a tile beyond N: zeros
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int wo0 = tw * p.bw, ho0 = th * p.bh, n0 = t * p.bn;
                for (int kb = 0; kb < kblocks; kb++, kbg++) {
                    const int st = kbg % STAGES;
                    if (kbg >= STAGES) mbar_wait(&s_empty[st], (uint32_t)(((kbg / STAGES) - 1) & 1));
                    const int tap = kb / cblocks, cb = kb - tap * cblocks;
                    const int kh = tap / p.KW, kw = tap - kh * p.KW;
                    unsigned char* sA = smem + st * STAGE_BYTES;
                    unsigned char* sB = sA + A_BYTES;
                    if constexpr (PAIR) {
                        if (rank == 0) mbar_expect_tx(&s_full[st], 2u * (uint32_t)STAGE_BYTES);       // both CTAs' bytes
                        tma_load_4d_pair(sA, &tmA, &s_full[st], cb * TK, wo0 * p.stride + kw - p.pad, ho0 * p.stride + kh - p.pad, n0);
                        tma_load_2d_pair(sB, &tmB, &s_full[st], kb * TK, co0 + (int)rank * (TN / 2));
                    } else {
                        mbar_expect_tx(&s_full[st], (uint32_t)STAGE_BYTES);
                        tma_load_4d(sA, &tmA, &s_full[st], cb * TK, wo0 * p.stride + kw - p.pad, ho0 * p.stride + kh - p.pad, n0);
                        tma_load_2d(sB, &tmB, &s_full[st], kb * TK, co0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && (!PAIR || rank == 0)) {
            int kbg = 0, it = 0;
            for (int w = w_first; w < total_work; w += w_step, it++) {
                const int ab = it & 1;
                if (it >= 2) {      // the epilogue must have drained this accumulator buffer (two items ago)
                    mbar_wait(&s_tmem_empty[ab], (uint32_t)(((it >> 1) - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t acc = tmem + (uint32_t)ab * TNA;
                for (int kb = 0; kb < kblocks; kb++, kbg++) {
                    const int st = kbg % STAGES;
                    mbar_wait(&s_full[st], (uint32_t)((kbg / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(smem + st * STAGE_BYTES), b0 = a0 + A_BYTES;
#pragma unroll
                    for (int j = 0; j < TK / 16; j++) {     // K = 16 per UMMA = 32 bytes along the swizzled row
                        const uint64_t da = make_desc_sw128(a0 + j * 32), db = make_desc_sw128(b0 + j * 32);
                        const uint32_t accumulate = (kb > 0 || j > 0) ? 1u : 0u;
                        if constexpr (PAIR) {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                ::"r"(acc), "l"(da), "l"(db), "r"(idesc_pair<TN>()), "r"(accumulate) : "memory");
                        } else {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                ::"r"(acc), "l"(da), "l"(db), "r"(idesc<TN>()), "r"(accumulate) : "memory");
                        }
                    }
                    // frees the stage when every UMMA issued so far has completed (implies fence::before_thread_sync);
                    // PAIR: in both CTAs
                    if constexpr (PAIR)
                        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(smem_u32(&s_empty[st])), "h"((uint16_t)3) : "memory");
                    else
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     ::"r"(smem_u32(&s_empty[st])) : "memory");
                }
                if constexpr (PAIR)
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(&s_tmem_full[ab])), "h"((uint16_t)3) : "memory");
                else
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(smem_u32(&s_tmem_full[ab])) : "memory");
            }
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4) .. +31; lane = tile row = one output pixel =====
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int per_img = p.bh * p.bw;
        const int nl = row / per_img, rem = row - nl * per_img;
        const int hl = rem / p.bw, wl = rem - hl * p.bw;
        // statistics scratch: 32 x 33 floats per epilogue warp, behind the operand ring (which is busy with the next item)
        float* sT = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES) + q * (32 * 17);
        // statistics are accumulated per CTA across its work items and leave as ONE atomic per channel when the
        // (image, channel block) changes or the CTA is done (tiles that span several images use direct atomics)
        const bool cta_stats = p.stats && p.bn == 1;
        const int et = tid - 64;                            // 0..127 within the four epilogue warps
        int key_n = -1, key_c = -1;
        if (cta_stats) {
            for (int c = et; c < 2 * 256; c += 128) s_stat[c] = 0.0f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        int it = 0;
        for (int w = w_first; w < total_work; w += w_step, it++) {
        const int ab = it & 1;
        int t = w % tiles_mw;
        const int co0 = (w / tiles_mw) * TN;
        if (PAIR) t = t * 2 + (int)rank;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int wo0 = tw * p.bw, ho0 = th * p.bh, n0 = t * p.bn;
        const int n = n0 + nl, ho = ho0 + hl, wo = wo0 + wl;
        const bool valid = n < p.N && ho < p.Ho && wo < p.Wo;
        const size_t pix = ((size_t)n * p.Ho + ho) * p.Wo + wo;                 // residual: dense [N, Ho, Wo, Cout]
        const size_t opix = (size_t)((long)n * p.osn + (long)ho * p.osh + (long)wo * p.osw);
        if (cta_stats && (n0 != key_n || co0 != key_c)) {
            asm volatile("bar.sync 1, 128;" ::: "memory");      // every warp has finished adding the previous item
            if (key_n >= 0 && key_n < p.N)
                for (int c = et; c < TN && key_c + c < p.Cout; c += 128) {
                    float* o = p.stats + ((size_t)key_n * p.Cout + key_c + c) * 2;
                    atomicAdd(o, s_stat[c]); atomicAdd(o + 1, s_stat[c + 256]);
                }
            for (int c = et; c < 2 * 256; c += 128) s_stat[c] = 0.0f;
            key_n = n0; key_c = co0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        mbar_wait(&s_tmem_full[ab], (uint32_t)((it >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem + (uint32_t)ab * TNA;
        constexpr int CW = TN >= 32 ? 32 : 16;          // accumulator columns per tcgen05.ld
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += CW) {
            uint32_t v[32];
            const uint32_t taddr = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            if (CW == 32) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
            } else {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int c = co0 + c0;
            // scale * acc + shift for this lane's pixel, CW channels (vector loads of the per-channel constants; the
            // null checks are hoisted: per-element `p.scale ? .. : ..` costs a branch per channel)
            float y[CW];
            if (c < p.Cout) {
                if (p.scale) {
                    const float4* s4 = reinterpret_cast<const float4*>(p.scale + c);
#pragma unroll
                    for (int g = 0; g < CW / 4; g++) {
                        const float4 t4 = __ldg(s4 + g);
                        y[4 * g] = __uint_as_float(v[4 * g]) * t4.x; y[4 * g + 1] = __uint_as_float(v[4 * g + 1]) * t4.y;
                        y[4 * g + 2] = __uint_as_float(v[4 * g + 2]) * t4.z; y[4 * g + 3] = __uint_as_float(v[4 * g + 3]) * t4.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CW; j++) y[j] = __uint_as_float(v[j]);
                }
                if (p.shift) {
                    const float4* s4 = reinterpret_cast<const float4*>(p.shift + c);
#pragma unroll
                    for (int g = 0; g < CW / 4; g++) {
                        const float4 t4 = __ldg(s4 + g);
                        y[4 * g] += t4.x; y[4 * g + 1] += t4.y; y[4 * g + 2] += t4.z; y[4 * g + 3] += t4.w;
                    }
                }
            }
            if (p.stats && c < p.Cout) {
                // column sums of this warp's 32 rows, 16 channels at a time, through a 32 x 17 shared-memory transpose
                // (lane j < 16 then owns channel c + h + j); all 32 rows of a warp lie in ONE image (128 / bn >= 32 rows);
                // rows outside the image add nothing
#pragma unroll
                for (int h = 0; h < CW; h += 16) {
#pragma unroll
                    for (int j = 0; j < 16; j++) sT[lane * 17 + j] = valid ? y[h + j] : 0.0f;
                    __syncwarp();
                    if (lane < 16) {
                        float sm = 0.0f, sq = 0.0f;
#pragma unroll
                        for (int r = 0; r < 32; r++) { const float v1 = sT[r * 17 + lane]; sm += v1; sq = __fmaf_rn(v1, v1, sq); }
                        if (cta_stats) {
                            atomicAdd(&s_stat[c0 + h + lane], sm); atomicAdd(&s_stat[c0 + h + lane + 256], sq);
                        } else {
                            const int nimg = n0 + (q * 32) / per_img;
                            if (nimg < p.N) {
                                float* o = p.stats + ((size_t)nimg * p.Cout + c + h + lane) * 2;
                                atomicAdd(o, sm); atomicAdd(o + 1, sq);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            if (valid && c < p.Cout) {
                if (p.residual) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + c);
#pragma unroll
                    for (int g = 0; g < CW / 8; g++) {
                        const uint4 r = __ldg(rp + g);
                        const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int h = 0; h < 4; h++) {
                            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[h]);
                            y[g * 8 + h * 2] += __low2float(b2);
                            y[g * 8 + h * 2 + 1] += __high2float(b2);
                        }
                    }
                }
                uint4* op = reinterpret_cast<uint4*>(p.out + opix + c);
#pragma unroll
                for (int g = 0; g < CW / 8; g++) {
                    uint32_t w4[4];
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        float a = y[g * 8 + h * 2], b = y[g * 8 + h * 2 + 1];
                        if (p.relu == 1) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                        else if (p.relu == 2) { a = tanhf(a); b = tanhf(b); }
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
                        w4[h] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    op[g] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                }
            }
        }
        // this accumulator buffer is drained: hand it back to the MMA warp (one arrival per epilogue warp)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            if (PAIR && rank != 0)      // the follower's epilogue reports to the leader's barrier (the MMA issuer waits there)
                asm volatile(
                    "{\n\t.reg .b32 ra;\n\t"
                    "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
                    "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
                    ::"r"(smem_u32(&s_tmem_empty[ab])) : "memory");
            else
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_tmem_empty[ab])) : "memory");
        }
        }
        if (cta_stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (key_n >= 0 && key_n < p.N)
                for (int c = et; c < TN && key_c + c < p.Cout; c += 128) {
                    float* o = p.stats + ((size_t)key_n * p.Cout + key_c + c) * 2;
                    atomicAdd(o, s_stat[c]); atomicAdd(o + 1, s_stat[c + 256]);
                }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (PAIR) {       // neither CTA may leave (or free tensor memory) while the other's MMAs / commits can still touch it
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (warp == 1)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2u * TNA));
    } else {
        if (warp == 1)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2u * TNA));
    }
}

// ------------------------------------------------------------------------------------------------------------
// Row-ring form for LARGE filters (the 7x7 layers of the generator; stride 1, input already padded, Cin = 64).  The tile
// form above fetches one 16 KB activation box per tap: 49 boxes per 128-pixel tile, 784 KB through the SM's L2 port for
// 128 x TN outputs -- the 7x7 layers ran exactly at that port's rate.  Here a persistent CTA walks DOWN a strip of 128
// output columns: every input row of the strip (128 + KW - 1 pixels x 64 channels, one 3-D TMA box) is loaded ONCE into a
// ring of row slots and serves all KH x KW taps of the output rows that touch it: the A operand of tap (kh, kw) for output
// row ho is the slot of input row ho + kh, its start address advanced by kw pixels (kw x 128 B; the 128-byte swizzle is a
// function of the shared-memory address bits, so a row-shifted start stays consistent with what TMA wrote into the
// 1024-byte-aligned slot).  Two output rows share each weight tap (two accumulators in tensor memory, double-buffered), so
// per output row the SM pulls one input row (17 KB) + half a filter (KH*KW*TN*64 B) instead of KH*KW x 16 KB + a filter.
constexpr uint32_t DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);   // high word of make_desc_sw128: SBO, version, layout
constexpr int RW_PX = 128;              // output pixels per strip row = TMEM lanes
constexpr int RW_RG = 2;                // output rows per weight pass
constexpr int RW_BST = 4;               // weight-tap ring depth
constexpr int RW_SEG = 16;              // output rows per work item (a vertical segment of a strip)
constexpr int RW_MAXSLOTS = 12;
constexpr int RW_THREADS = 224;         // producer, MMA issuer (row 0), 4 epilogue warps, MMA issuer (row 1)

template <int TN>
__global__ void __launch_bounds__(RW_THREADS) k_conv_rows(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB, const ConvParams p,
                                                        const int slot_bytes, const int nslots) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int B_BYTES = TN * TK * 2;
    constexpr uint32_t TNA = TN < 32 ? 32 : TN;
    __shared__ __align__(8) uint64_t s_rfull[RW_MAXSLOTS], s_rempty[RW_MAXSLOTS], s_bfull[RW_BST], s_bempty[RW_BST],
        s_tmem_full[2], s_tmem_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_stat[2 * 64];
    unsigned char* ring = smem;
    unsigned char* bring = smem + (size_t)nslots * slot_bytes;
    float* sT_all = reinterpret_cast<float*>(bring + RW_BST * B_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = p.KH * p.KW;
    const int strips = (p.Wo + RW_PX - 1) / RW_PX, segs = (p.Ho + RW_SEG - 1) / RW_SEG;
    const int total_work = p.N * strips * segs;

    if (tid == 0) {
        // two MMA issuers (one per output row of a group): what they release needs both of their commits
        for (int s = 0; s < nslots; s++) { mbar_init(&s_rfull[s], 1); mbar_init(&s_rempty[s], RW_RG); }
        for (int s = 0; s < RW_BST; s++) { mbar_init(&s_bfull[s], 1); mbar_init(&s_bempty[s], RW_RG); }
        for (int s = 0; s < 2; s++) { mbar_init(&s_tmem_full[s], RW_RG); mbar_init(&s_tmem_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"(2u * RW_RG * TNA));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // work item -> (image, strip, segment); segments of one strip are consecutive items (their halo rows are L2-hot)
    auto decode = [&](int w, int& n, int& wo0, int& ho0, int& nrows) {
        const int seg = w % segs; w /= segs;
        const int strip = w % strips; n = w / strips;
        wo0 = strip * RW_PX; ho0 = seg * RW_SEG; nrows = min(RW_SEG, p.Ho - ho0);
    };

    // The producer and the MMA issuer are single threads: their per-tap code must stay a handful of instructions (a 64-bit
    // modulo per tap made the first version of this kernel issue-latency-bound at ~1800 clk per tap).  Ring positions are
    // (slot, lap) pairs advanced by compare-and-wrap; laps give the mbarrier parities.
    struct RingPos { int slot; uint32_t lap; };
    auto advance = [&](RingPos& r) { if (++r.slot == nslots) { r.slot = 0; r.lap++; } };
    const uint32_t row_bytes = (uint32_t)((RW_PX + p.KW - 1) * 128);
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
            RingPos ld = {0, 0};            // next row slot to fill
            uint32_t tcount = 0;            // weight taps issued so far
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                int n, wo0, ho0, nrows;
                decode(w, n, wo0, ho0, nrows);
                const int nin = nrows + p.KH - 1;           // input rows of this item
                int loaded = 0;
                auto load_rows_upto = [&](int upto) {       // rows [loaded, upto) of the item
                    for (; loaded < upto && loaded < nin; loaded++) {
                        if (ld.lap > 0) mbar_wait(&s_rempty[ld.slot], (ld.lap - 1) & 1);
                        mbar_expect_tx(&s_rfull[ld.slot], row_bytes);
                        tma_load_4d(ring + (size_t)ld.slot * slot_bytes, &tmA, &s_rfull[ld.slot], 0, wo0, ho0 + loaded, n);
                        advance(ld);
                    }
                };
                const int groups = (nrows + RW_RG - 1) / RW_RG;
                load_rows_upto(RW_RG + p.KH - 1);           // the first group's rows
                for (int g = 0; g < groups; g++) {
                    for (int t = 0; t < taps; t++, tcount++) {
                        if (t == (taps >> 1)) load_rows_upto((g + 2) * RW_RG + p.KH - 1);      // next group's rows, early
                        const int st = (int)(tcount & (RW_BST - 1));
                        if (tcount >= RW_BST) mbar_wait(&s_bempty[st], ((tcount / RW_BST) - 1) & 1);
                        mbar_expect_tx(&s_bfull[st], (uint32_t)B_BYTES);
                        tma_load_2d(bring + st * B_BYTES, &tmB, &s_bfull[st], t * TK, 0);
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 6) {
        // ONE issuing thread per output row of the group: a tcgen05.mma costs its issuer ~100 clk (descriptor moves into
        // uniform registers), 8 of them per tap from one thread were the bound at ~850 clk per tap
        const int jrow = warp == 1 ? 0 : 1;
        if (lane == 0) {
            RingPos wt = {0, 0};            // next row to wait for
            RingPos gr = {0, 0};            // first row of the current group
            uint32_t tcount = 0, gcount = 0;
            const uint32_t ring_u32 = smem_u32(ring), bring_u32 = smem_u32(bring);
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                int n, wo0, ho0, nrows;
                decode(w, n, wo0, ho0, nrows);
                const int nin = nrows + p.KH - 1;
                const int groups = (nrows + RW_RG - 1) / RW_RG;
                int waited = 0;                             // input rows of this item known to have landed
                for (int g = 0; g < groups; g++, gcount++) {
                    const int ab = (int)(gcount & 1);
                    if (gcount >= 2) {
                        mbar_wait(&s_tmem_empty[ab], ((gcount >> 1) - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const int rows_here = min(RW_RG, nrows - g * RW_RG);
                    const int need = g * RW_RG + rows_here + p.KH - 1;
                    for (; waited < need; waited++) { mbar_wait(&s_rfull[wt.slot], wt.lap & 1); advance(wt); }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t acc0 = tmem + (uint32_t)(ab * RW_RG) * TNA;
                    int t = 0;
                    for (int kh = 0; kh < p.KH; kh++) {
                        // shared-memory addresses of the input rows that tap row kh pairs with output rows 0 / 1 of the group
                        int s0 = gr.slot + kh; if (s0 >= nslots) s0 -= nslots;
                        int s1 = s0 + 1; if (s1 >= nslots) s1 -= nslots;
                        // descriptor low words (start address >> 4; + 8 per pixel of kw shift, + 2 per K = 16 step): the
                        // issuing thread spends a handful of instructions per tcgen05.mma, not a descriptor build
                        const uint32_t r0 = ((ring_u32 + (uint32_t)s0 * (uint32_t)slot_bytes) & 0x3FFFFu) >> 4;
                        const uint32_t r1 = ((ring_u32 + (uint32_t)s1 * (uint32_t)slot_bytes) & 0x3FFFFu) >> 4;
                        for (int kw = 0; kw < p.KW; kw++, t++, tcount++) {
                            const int st = (int)(tcount & (RW_BST - 1));
                            mbar_wait(&s_bfull[st], (tcount / RW_BST) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t b_lo = ((bring_u32 + (uint32_t)st * B_BYTES) & 0x3FFFFu) >> 4;
                            {
                                const int j = jrow;
                                if (j < rows_here) {
                                    const uint32_t a_lo = (j ? r1 : r0) + (uint32_t)kw * 8u;
                                    const uint32_t acc = acc0 + (uint32_t)j * TNA;
#pragma unroll
                                    for (int q = 0; q < TK / 16; q++) {
                                        const uint32_t accumulate = (t > 0 || q > 0) ? 1u : 0u;
                                        asm volatile(
                                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                                            "mov.b64 da, {%1, %5};\n\t"
                                            "mov.b64 db, {%2, %5};\n\t"
                                            "setp.ne.b32 p, %4, 0;\n\t"
                                            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                                            ::"r"(acc), "r"(a_lo + 2u * q), "r"(b_lo + 2u * q), "r"(idesc<TN>()), "r"(accumulate),
                                              "r"(DESC_HI_SW128) : "memory");
                                    }
                                }
                            }
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                         ::"r"(smem_u32(&s_bempty[st])) : "memory");
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(smem_u32(&s_tmem_full[ab])) : "memory");
                    // input rows no later group needs: the group's own first rows; after the last group also the halo
                    const int retire = (g == groups - 1) ? nin - g * RW_RG : RW_RG;
                    for (int i = 0; i < retire; i++) {
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                     ::"r"(smem_u32(&s_rempty[gr.slot])) : "memory");
                        advance(gr);
                    }
                }
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        const int q = warp & 3;
        float* sT = sT_all + q * (32 * 17);
        const bool cta_stats = p.stats != nullptr;
        const int et = tid - 64;
        int key_n = -1;
        if (cta_stats) { if (et < 2 * 64) s_stat[et] = 0.0f; asm volatile("bar.sync 1, 128;" ::: "memory"); }
        long gcount = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            int n, wo0, ho0, nrows;
            decode(w, n, wo0, ho0, nrows);
            if (cta_stats && n != key_n) {
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (key_n >= 0 && et < TN && et < p.Cout) {
                    float* o = p.stats + ((size_t)key_n * p.Cout + et) * 2;
                    atomicAdd(o, s_stat[et]); atomicAdd(o + 1, s_stat[et + 64]);
                }
                if (et < 2 * 64) s_stat[et] = 0.0f;
                key_n = n;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            const int groups = (nrows + RW_RG - 1) / RW_RG;
            const int wo = wo0 + q * 32 + lane;
            for (int g = 0; g < groups; g++, gcount++) {
                const int ab = (int)(gcount & 1);
                mbar_wait(&s_tmem_full[ab], (uint32_t)((gcount >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int rows_here = min(RW_RG, nrows - g * RW_RG);
                for (int j = 0; j < rows_here; j++) {
                    const int ho = ho0 + g * RW_RG + j;
                    const bool valid = wo < p.Wo;
                    const size_t opix = (size_t)((long)n * p.osn + (long)ho * p.osh + (long)wo * p.osw);
                    constexpr int CW = TN >= 32 ? 32 : 16;
#pragma unroll 1
                    for (int c0 = 0; c0 < TN; c0 += CW) {
                        uint32_t v[32];
                        const uint32_t taddr = tmem + (uint32_t)(ab * RW_RG + j) * TNA + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                        if (CW == 32) {
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                : "r"(taddr));
                        } else {
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                : "r"(taddr));
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        float y[CW];
#pragma unroll
                        for (int k = 0; k < CW; k++) {
                            float t = __uint_as_float(v[k]);
                            if (p.scale) t *= __ldg(p.scale + c0 + k);
                            if (p.shift) t += __ldg(p.shift + c0 + k);
                            y[k] = t;
                        }
                        if (p.stats) {
#pragma unroll
                            for (int h = 0; h < CW; h += 16) {
#pragma unroll
                                for (int k = 0; k < 16; k++) sT[lane * 17 + k] = valid ? y[h + k] : 0.0f;
                                __syncwarp();
                                if (lane < 16) {
                                    float sm = 0.0f, sq = 0.0f;
#pragma unroll
                                    for (int r = 0; r < 32; r++) { const float v1 = sT[r * 17 + lane]; sm += v1; sq = __fmaf_rn(v1, v1, sq); }
                                    atomicAdd(&s_stat[c0 + h + lane], sm); atomicAdd(&s_stat[c0 + h + lane + 64], sq);
                                }
                                __syncwarp();
                            }
                        }
                        if (valid) {
                            uint4* op = reinterpret_cast<uint4*>(p.out + opix + c0);
#pragma unroll
                            for (int gq = 0; gq < CW / 8; gq++) {
                                uint32_t w4[4];
#pragma unroll
                                for (int h = 0; h < 4; h++) {
                                    float a = y[gq * 8 + h * 2], b = y[gq * 8 + h * 2 + 1];
                                    if (p.relu == 1) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
                                    else if (p.relu == 2) { a = tanhf(a); b = tanhf(b); }
                                    const __nv_bfloat162 b2 = __floats2bfloat162_rn(a, b);
                                    w4[h] = *reinterpret_cast<const uint32_t*>(&b2);
                                }
                                op[gq] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_tmem_empty[ab])) : "memory");
            }
        }
        if (cta_stats) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (key_n >= 0 && et < TN && et < p.Cout) {
                float* o = p.stats + ((size_t)key_n * p.Cout + et) * 2;
                atomicAdd(o, s_stat[et]); atomicAdd(o + 1, s_stat[et + 64]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2u * RW_RG * TNA));
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient of the convolution (training path of the encoder trunk):
//     dW[co, kh, kw, ci] = sum_{n, ho, wo} g[n, ho, wo, co] * x[n, ho*s + kh - p, wo*s + kw - p, ci]
// as a GEMM whose REDUCTION runs over pixels:  D[co, ci] (+)= G^T[co, pix] . X_tap[pix, ci], one (tap, channel blocks,
// pixel range) per work item.  Both operands are "MN-major" for the tensor core: in memory a pixel is a row of
// contiguous channels, i.e. the M (co) and N (ci) indices are the contiguous ones and K (pixels) strides by a row.  A
// K block = a patch of 64 pixels: TMA box {64 channels, bw, bh, bn} of g, and the box of x shifted by the tap (traversal
// stride = convolution stride, out-of-bounds = zero padding) -- 64 rows of 128 bytes, 128-byte swizzle, exactly the
// canonical MN-major SWIZZLE_128B operand ((8 x 16 B, n), (8, k)) : ((1, LBO), (8 x 16 B, SBO)): SBO = 1024 B between
// 8-pixel groups, LBO = 8192 B between 64-channel blocks (a second box), K = 16 pixels per tcgen05.mma = +2048 B.
// The pixel range is split over several work items (the layers have 9 taps x 1..16 channel-block pairs only), partial
// results are added into the fp32 gradient with atomics.  Same warp roles / mbarrier ring / persistent loop as above.
// ------------------------------------------------------------------------------------------------------------
struct WgradParams {
    int N, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad;
    int bw, bh, bn, tiles_w, tiles_h, tiles_n;      // 64-pixel patches of the OUTPUT grid
    int splits, patches_per_split, atomic;         // atomic: add into dw (several pixel ranges, or accumulate mode)
    float* dw;                                      // [Cout, KH, KW, Cin] fp32, accumulated into
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t addr_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr_bytes & 0x3FFFFu) >> 4);
    d |= (uint64_t)(8192u >> 4) << 16;              // LBO: next 64-channel block (a second 64 x 128 B box)
    d |= (uint64_t)(1024u >> 4) << 32;              // SBO: next group of 8 pixels
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int TNW>
__global__ void __launch_bounds__(NTHREADS) k_conv_wgrad(const __grid_constant__ CUtensorMap tmG,
                                                         const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int STAGES = 6;
    constexpr int A_B = 2 * 8192, B_B = (TNW / 64) * 8192, STAGE_BYTES = A_B + B_B;
    __shared__ __align__(8) uint64_t s_full[STAGES], s_empty[STAGES], s_tmem_full[2], s_tmem_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_tr[4][32][36];          // per epilogue warp: 32 x 32 transpose, rows padded to 144 B
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cobs = (p.Cout + 127) / 128, cibs = (p.Cin + TNW - 1) / TNW, taps = p.KH * p.KW;
    const int total_work = cobs * cibs * taps * p.splits;
    const int patches = p.tiles_w * p.tiles_h * p.tiles_n;
    // instruction descriptor: F32 accumulate, BF16 x BF16, A and B MN-major (bits 15, 16), N = TNW, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TNW >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&s_tmem_full[s], 1); mbar_init(&s_tmem_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(2u * TNW));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // work item -> (split, tap, ci block, co block); split fastest so that concurrent CTAs hit different pixels
    auto decode = [&](int w, int& split, int& tap, int& cib, int& cob) {
        split = w % p.splits; w /= p.splits;
        tap = w % taps; w /= taps;
        cib = w % cibs; cob = w / cibs;
    };
    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmG) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmX) : "memory");
            int kbg = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                int split, tap, cib, cob;
                decode(w, split, tap, cib, cob);
                const int kh = tap / p.KW, kw = tap - kh * p.KW;
                const int p0 = split * p.patches_per_split, p1 = min(p0 + p.patches_per_split, patches);
                for (int pt = p0; pt < p1; pt++, kbg++) {
                    int t = pt;
                    const int tw = t % p.tiles_w; t /= p.tiles_w;
                    const int th = t % p.tiles_h; t /= p.tiles_h;
                    const int wo0 = tw * p.bw, ho0 = th * p.bh, n0 = t * p.bn;
                    const int st = kbg % STAGES;
                    if (kbg >= STAGES) mbar_wait(&s_empty[st], (uint32_t)(((kbg / STAGES) - 1) & 1));
                    unsigned char* sA = smem + st * STAGE_BYTES;
                    unsigned char* sB = sA + A_B;
                    mbar_expect_tx(&s_full[st], (uint32_t)STAGE_BYTES);
                    tma_load_4d(sA, &tmG, &s_full[st], cob * 128, wo0, ho0, n0);
                    tma_load_4d(sA + 8192, &tmG, &s_full[st], cob * 128 + 64, wo0, ho0, n0);
                    const int xw = wo0 * p.stride + kw - p.pad, xh = ho0 * p.stride + kh - p.pad;
#pragma unroll
                    for (int h = 0; h < TNW / 64; h++) tma_load_4d(sB + h * 8192, &tmX, &s_full[st], cib * TNW + h * 64, xw, xh, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int kbg = 0, it = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
                int split, tap, cib, cob;
                decode(w, split, tap, cib, cob);
                const int p0 = split * p.patches_per_split, p1 = min(p0 + p.patches_per_split, patches);
                const int ab = it & 1;
                if (it >= 2) {
                    mbar_wait(&s_tmem_empty[ab], (uint32_t)(((it >> 1) - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t acc = tmem + (uint32_t)ab * TNW;
                for (int pt = p0; pt < p1; pt++, kbg++) {
                    const int st = kbg % STAGES;
                    mbar_wait(&s_full[st], (uint32_t)((kbg / STAGES) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(smem + st * STAGE_BYTES), b0 = a0 + A_B;
#pragma unroll
                    for (int j = 0; j < 4; j++) {           // 16 pixels per UMMA = two 8-pixel groups = 2048 bytes
                        const uint64_t da = make_desc_mn_sw128(a0 + j * 2048), db = make_desc_mn_sw128(b0 + j * 2048);
                        const uint32_t accumulate = (pt > p0 || j > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(acc), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(smem_u32(&s_empty[st])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"(smem_u32(&s_tmem_full[ab])) : "memory");
            }
        }
    } else {
        const int q = warp & 3;
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
            int split, tap, cib, cob;
            decode(w, split, tap, cib, cob);
            const int ab = it & 1;
            mbar_wait(&s_tmem_full[ab], (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem + (uint32_t)ab * TNW;
            const bool empty = split * p.patches_per_split >= patches;     // a trailing split without patches
#pragma unroll 1
            for (int c0 = 0; c0 < TNW; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // lane = output channel (row): transpose the 32 x 32 block through shared memory so that a warp instruction
                // covers four rows of 128 contiguous bytes (16-byte vector reductions / stores instead of 32 scattered words)
                float* tr = s_tr[q][lane];
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(tr + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                __syncwarp();
                const int ci = cib * TNW + c0 + 4 * (lane & 7);
                if (!empty && ci < p.Cin) {
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) {
                        const int r = rr * 4 + (lane >> 3), co = cob * 128 + q * 32 + r;
                        if (co >= p.Cout) continue;
                        const float4 val = *reinterpret_cast<const float4*>(&s_tr[q][r][4 * (lane & 7)]);
                        float* o = p.dw + ((size_t)co * taps + tap) * p.Cin + ci;
                        if (!p.atomic) *reinterpret_cast<float4*>(o) = val;
                        else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                                          ::"l"(o), "f"(val.x), "f"(val.y), "f"(val.z), "f"(val.w) : "memory");
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_tmem_empty[ab])) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2u * TNW));
}

// im2col for convolutions whose input has too few channels for a 64-wide K block (the 7x7 stem of ResNet-18, Cin = 3):
// out[n, ho, wo, (kh*KW + kw)*Cin + c] = in[n, ho*s + kh - p, wo*s + kw - p, c] (zero outside, zero beyond KH*KW*Cin up
// to Kpad); the stem then runs as a 1x1 convolution over Kpad channels.
__global__ void __launch_bounds__(256) k_im2col_nhwc(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                     int N, int H, int W, int Cin, int Ho, int Wo, int KH, int KW,
                                                     int stride, int pad, int Kpad) {
    // one thread per 8 consecutive output columns (one 16-byte store); Kpad is a multiple of 64
    const int groups = Kpad >> 3, K = KH * KW * Cin;
    const long total = (long)N * Ho * Wo * groups;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % groups);
        long r = i / groups;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int n = (int)(r / Ho);
        const __nv_bfloat16* img = in + (size_t)n * H * W * Cin;
        __nv_bfloat16 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = g * 8 + j;
            __nv_bfloat16 x = __float2bfloat16(0.0f);
            if (k < K) {
                const int c = k % Cin, tap = k / Cin, kw = tap % KW, kh = tap / KW;
                const int hi = ho * stride + kh - pad, wi = wo * stride + kw - pad;
                if (hi >= 0 && hi < H && wi >= 0 && wi < W) x = img[((size_t)hi * W + wi) * Cin + c];
            }
            v[j] = x;
        }
        *reinterpret_cast<uint4*>(out + (size_t)i * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

// The same, one block per output row (n, ho): the KH input rows it reads are staged in shared memory once (for the 7x7 /
// stride-2 stem every input element is used ~12 times), then 16-byte stores.  Used when the rows fit in 48 KB.
__global__ void __launch_bounds__(256) k_im2col_rows(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                     int H, int W, int Cin, int Ho, int Wo, int KH, int KW, int stride, int pad,
                                                     int Kpad) {
    extern __shared__ __align__(16) unsigned char s_rows_raw[];
    __nv_bfloat16* s_rows = reinterpret_cast<__nv_bfloat16*>(s_rows_raw);
    const int n = blockIdx.x / Ho, ho = blockIdx.x - n * Ho;
    const int rowlen = W * Cin;                                  // elements per input row
    for (int kh = 0; kh < KH; kh++) {
        const int hi = ho * stride + kh - pad;
        __nv_bfloat16* dst = s_rows + (size_t)kh * rowlen;
        if (hi < 0 || hi >= H) continue;                         // never read: the gather below tests hi again
        const __nv_bfloat16* src = in + ((size_t)n * H + hi) * rowlen;
        if ((rowlen & 7) == 0) {
            for (int i = threadIdx.x; i < (rowlen >> 3); i += 256)
                reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
        } else {
            for (int i = threadIdx.x; i < rowlen; i += 256) dst[i] = src[i];
        }
    }
    __syncthreads();
    const int groups = Kpad >> 3, KWC = KW * Cin, K = KH * KWC;
    __nv_bfloat16* orow = out + (size_t)blockIdx.x * Wo * Kpad;
    const __nv_bfloat16 zero = __float2bfloat16(0.0f);
    for (int i = threadIdx.x; i < Wo * groups; i += 256) {
        const int wo = i / groups, g = i - wo * groups;
        int k = g * 8, kh = k / KWC, j = k - kh * KWC;
        const int base = (wo * stride - pad) * Cin;              // element index in the row of (kw = 0, c = 0)
        __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++, k++) {
            __nv_bfloat16 x = zero;
            if (k < K) {
                const int hi = ho * stride + kh - pad, idx = base + j;
                if (hi >= 0 && hi < H && idx >= 0 && idx < rowlen) x = s_rows[kh * rowlen + idx];
            }
            v[e] = x;
            if (++j == KWC) { j = 0; kh++; }
        }
        *reinterpret_cast<uint4*>(orow + (size_t)i * 8) = *reinterpret_cast<const uint4*>(v);
    }
}

// ---- instance normalisation of an NHWC bf16 tensor (textural/models/networks.py:23-29: InstanceNorm2d, affine=False) ----
// stats: per (n, c) sum and sum of squares over H*W in fp32.  A block of 256 threads covers `rows` pixels of one image
// and a slab of up to 512 channels: thread t owns channel pair t % (slab/2) and walks rows t / (slab/2), += 256/(slab/2)
// (coalesced 4-byte reads along C, several rows in flight per block); partials are combined through shared memory and
// leave the block as one atomicAdd per channel.
__global__ void __launch_bounds__(256) k_inorm_stats(const __nv_bfloat16* __restrict__ x, float* __restrict__ stats,
                                                     int HW, int C, int rows, int pairs) {
    __shared__ float s_part[256 * 4];
    const int n = blockIdx.z;
    const int cp = threadIdx.x % pairs, rl = threadIdx.x / pairs, nrl = 256 / pairs;
    const int c = (blockIdx.y * pairs + cp) * 2;
    const int r0 = blockIdx.x * rows, r1 = min(r0 + rows, HW);
    float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
    if (c < C && rl < nrl) {
        const __nv_bfloat16* base = x + (size_t)n * HW * C + c;
#pragma unroll 4
        for (int r = r0 + rl; r < r1; r += nrl) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(base + (size_t)r * C);
            const float a = __low2float(v), b = __high2float(v);
            s0 += a; s1 += b; q0 = __fmaf_rn(a, a, q0); q1 = __fmaf_rn(b, b, q1);
        }
    }
    s_part[threadIdx.x * 4] = s0; s_part[threadIdx.x * 4 + 1] = q0; s_part[threadIdx.x * 4 + 2] = s1; s_part[threadIdx.x * 4 + 3] = q1;
    __syncthreads();
    if (rl == 0 && c < C) {
        for (int k = 1; k < nrl; k++) {
            const float* o = s_part + (k * pairs + cp) * 4;
            s0 += o[0]; q0 += o[1]; s1 += o[2]; q1 += o[3];
        }
        float* o = stats + ((size_t)n * C + c) * 2;
        atomicAdd(o, s0); atomicAdd(o + 1, q0); atomicAdd(o + 2, s1); atomicAdd(o + 3, q1);
    }
}

// apply: y = act((x - mean) * rsqrt(var + eps) (+ residual)); biased variance like torch.nn.InstanceNorm2d.
// grid.y = image; a thread handles 8 consecutive channels (16 bytes) of one pixel; 32-bit index arithmetic.
__global__ void __launch_bounds__(256) k_inorm_apply(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                                     const __nv_bfloat16* __restrict__ residual,
                                                     __nv_bfloat16* __restrict__ y, int HW, int C, float eps, int act) {
    const int n = blockIdx.y;
    const int cg = C >> 3;                                      // 16-byte groups per pixel
    const unsigned total = (unsigned)HW * (unsigned)cg;
    const float inv = 1.0f / (float)HW;
    const size_t img = (size_t)n * HW * C;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned g = i % (unsigned)cg;
        const float* st = stats + ((size_t)n * C + g * 8) * 2;
        const size_t e = img + (size_t)i * 8;
        const uint4 xv = *reinterpret_cast<const uint4*>(x + e);
        uint4 rv = make_uint4(0, 0, 0, 0);
        if (residual) rv = *reinterpret_cast<const uint4*>(residual + e);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t ow[4];
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const float4 s4 = *reinterpret_cast<const float4*>(st + h * 4);      // (sum, sumsq) of two channels
            const float m0 = s4.x * inv, m1 = s4.z * inv;
            const float r0 = rsqrtf(fmaxf(s4.y * inv - m0 * m0, 0.0f) + eps), r1 = rsqrtf(fmaxf(s4.w * inv - m1 * m1, 0.0f) + eps);
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&xw[h]);
            float a = (__low2float(v) - m0) * r0, b = (__high2float(v) - m1) * r1;
            if (residual) {
                const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(&rw[h]);
                a += __low2float(rr); b += __high2float(rr);
            }
            if (act == 1) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
            const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
            ow[h] = *reinterpret_cast<const uint32_t*>(&o2);
        }
        *reinterpret_cast<uint4*>(y + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// apply + nn.ReflectionPad2d(pad) in one pass: y [N, H + 2 pad, W + 2 pad, C] = reflect_pad(act(norm(x) (+ residual))).
// The next 3x3 / 7x7 convolution of the generator reads exactly this tensor, so the normalised activation is never
// written unpadded (one read + one write instead of two of each).
__global__ void __launch_bounds__(256) k_inorm_apply_pad(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                                         const __nv_bfloat16* __restrict__ residual,
                                                         __nv_bfloat16* __restrict__ y, int H, int W, int C, float eps,
                                                         int act, int pad) {
    const int n = blockIdx.y;
    const int cg = C >> 3, Hp = H + 2 * pad, Wp = W + 2 * pad;
    const unsigned total = (unsigned)Hp * (unsigned)Wp * (unsigned)cg;
    const float inv = 1.0f / (float)(H * W);
    const size_t img = (size_t)n * H * W * C, imgp = (size_t)n * Hp * Wp * C;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned g = i % (unsigned)cg;
        const unsigned pp = i / (unsigned)cg;
        int h = (int)(pp / (unsigned)Wp) - pad, w = (int)(pp % (unsigned)Wp) - pad;
        h = h < 0 ? -h : (h >= H ? 2 * (H - 1) - h : h);
        w = w < 0 ? -w : (w >= W ? 2 * (W - 1) - w : w);
        const float* st = stats + ((size_t)n * C + g * 8) * 2;
        const size_t e = img + ((size_t)h * W + w) * C + g * 8;
        const uint4 xv = *reinterpret_cast<const uint4*>(x + e);
        uint4 rv = make_uint4(0, 0, 0, 0);
        if (residual) rv = *reinterpret_cast<const uint4*>(residual + e);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 s4 = *reinterpret_cast<const float4*>(st + k * 4);
            const float m0 = s4.x * inv, m1 = s4.z * inv;
            const float r0 = rsqrtf(fmaxf(s4.y * inv - m0 * m0, 0.0f) + eps), r1 = rsqrtf(fmaxf(s4.w * inv - m1 * m1, 0.0f) + eps);
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&xw[k]);
            float a = (__low2float(v) - m0) * r0, b = (__high2float(v) - m1) * r1;
            if (residual) {
                const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(&rw[k]);
                a += __low2float(rr); b += __high2float(rr);
            }
            if (act == 1) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
            const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
            ow[k] = *reinterpret_cast<const uint32_t*>(&o2);
        }
        *reinterpret_cast<uint4*>(y + imgp + (size_t)i * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ---- BatchNorm2d in training mode around the convolutions of the encoder trunk (torchvision BasicBlock: conv -> bn -> relu,
// conv -> bn -> (+ identity) -> relu).  Forward: the convolution's epilogue already produced the per-channel sums; the
// normalisation is y = act(z * scale[c] + shift[c] (+ residual)) with scale = gamma * rstd, shift = beta - mean * scale.
__global__ void __launch_bounds__(256) k_affine_act(const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, const __nv_bfloat16* __restrict__ residual,
                                                    __nv_bfloat16* __restrict__ y, long groups, int C, int act) {
    const int cg = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cg) * 8;
        const uint4 zv = *reinterpret_cast<const uint4*>(z + i * 8);
        uint4 rv = make_uint4(0, 0, 0, 0);
        if (residual) rv = *reinterpret_cast<const uint4*>(residual + i * 8);
        const uint32_t zw[4] = {zv.x, zv.y, zv.z, zv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&zw[k]);
            float a = __fmaf_rn(__low2float(v), __ldg(scale + c + 2 * k), __ldg(shift + c + 2 * k));
            float b = __fmaf_rn(__high2float(v), __ldg(scale + c + 2 * k + 1), __ldg(shift + c + 2 * k + 1));
            if (residual) {
                const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(&rw[k]);
                a += __low2float(rr); b += __high2float(rr);
            }
            if (act == 1) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
            const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
            ow[k] = *reinterpret_cast<const uint32_t*>(&o2);
        }
        *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ---- MaxPool2d (ResNet stem: 3x3, stride 2, padding 1) on NHWC bf16; a thread owns 8 consecutive channels of one pixel.
// Forward: maximum over the window's in-bounds elements (torch pads with -inf) and, optionally, the position kh * K + kw of
// the window's FIRST maximum in scan order (one byte per element; the element torch's max_pool2d_with_indices records).
// Backward: GATHER form -- each INPUT pixel visits the (at most ceil(K/s)^2) windows that contain it and takes a window's
// gradient when the recorded position is its own; no atomics, every gradient element written exactly once.
__device__ __forceinline__ void bf8_to_float(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        f[2 * k] = __low2float(b); f[2 * k + 1] = __high2float(b);
    }
}
__device__ __forceinline__ uint4 float_to_bf8(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
        w[k] = *reinterpret_cast<const uint32_t*>(&b);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256) k_max_pool_nhwc(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                       unsigned char* __restrict__ idx, int N, int H, int W, int C, int Ho,
                                                       int Wo, int K, int stride, int pad) {
    const int cg = C >> 3;
    const long total = (long)N * Ho * Wo * cg;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        long r = i / cg;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const int n = (int)(r / Ho);
        float m[8];
        unsigned char am[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { m[k] = -INFINITY; am[k] = 255; }
        for (int kh = 0; kh < K; kh++) {
            const int hi = ho * stride + kh - pad;
            if (hi < 0 || hi >= H) continue;
            for (int kw = 0; kw < K; kw++) {
                const int wi = wo * stride + kw - pad;
                if (wi < 0 || wi >= W) continue;
                float v[8];
                bf8_to_float(*reinterpret_cast<const uint4*>(x + (((size_t)n * H + hi) * W + wi) * C + g * 8), v);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (v[k] > m[k]) { m[k] = v[k]; am[k] = (unsigned char)(kh * K + kw); }      // strict: the FIRST maximum
            }
        }
        *reinterpret_cast<uint4*>(y + (size_t)i * 8) = float_to_bf8(m);
        if (idx) *reinterpret_cast<uint2*>(idx + (size_t)i * 8) = *reinterpret_cast<const uint2*>(am);
    }
}

__global__ void __launch_bounds__(256) k_max_pool_bwd_nhwc(const unsigned char* __restrict__ idx, const __nv_bfloat16* __restrict__ gy,
                                                           __nv_bfloat16* __restrict__ gx, int N, int H, int W, int C, int Ho,
                                                           int Wo, int K, int stride, int pad) {
    const int cg = C >> 3;
    const long total = (long)N * H * W * cg;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        long r = i / cg;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int n = (int)(r / H);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0.0f;
        // windows (ho, wo) with ho * stride - pad <= h <= ho * stride - pad + K - 1
        int ho0 = h + pad - (K - 1); ho0 = ho0 <= 0 ? 0 : (ho0 + stride - 1) / stride;
        int wo0 = w + pad - (K - 1); wo0 = wo0 <= 0 ? 0 : (wo0 + stride - 1) / stride;
        const int ho1 = min(Ho - 1, (h + pad) / stride), wo1 = min(Wo - 1, (w + pad) / stride);
        for (int ho = ho0; ho <= ho1; ho++)
            for (int wo = wo0; wo <= wo1; wo++) {
                const unsigned char code = (unsigned char)((h - (ho * stride - pad)) * K + (w - (wo * stride - pad)));
                const size_t o = (((size_t)n * Ho + ho) * Wo + wo) * C + g * 8;
                const uint2 iv = *reinterpret_cast<const uint2*>(idx + o);
                const unsigned char* am = reinterpret_cast<const unsigned char*>(&iv);
                float gv[8];
                bf8_to_float(*reinterpret_cast<const uint4*>(gy + o), gv);
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] += am[k] == code ? gv[k] : 0.0f;
            }
        *reinterpret_cast<uint4*>(gx + (size_t)i * 8) = float_to_bf8(acc);
    }
}

// Batch statistics -> everything the normalising pass and the backward need, in one launch of C threads: the convolution's
// epilogue left per-(image, channel) sum / sum of squares in stats [N, C, 2]; out [4, C] = (mean, rstd, scale, shift) with
// scale = gamma * rstd, shift = beta - mean * scale; running statistics updated in place as torch.nn.BatchNorm2d does
// (momentum, unbiased variance).  The combination sum / count - mean^2 is taken in double.
__global__ void __launch_bounds__(128) k_bn_finalize(const float* __restrict__ stats, int N, int C, double count, float eps,
                                                     float momentum, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ running_mean,
                                                     float* __restrict__ running_var, long long* __restrict__ batches,
                                                     float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && batches) *batches += 1;
    if (c >= C) return;
    double s = 0.0, ss = 0.0;
    for (int n = 0; n < N; n++) {
        s += (double)stats[((size_t)n * C + c) * 2];
        ss += (double)stats[((size_t)n * C + c) * 2 + 1];
    }
    const double m = s / count;
    double v = ss / count - m * m;
    if (v < 0.0) v = 0.0;
    const float mean = (float)m, var = (float)v;
    const float rstd = rsqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    out[c] = mean; out[C + c] = rstd; out[2 * C + c] = sc; out[3 * C + c] = beta[c] - mean * sc;
    if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) {
        const float unbiased = (float)(v * (count / (count > 1.0 ? count - 1.0 : 1.0)));
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// Weights for the data gradient: src fp32 [Co, KH, KW, Ci] (the physical layout of a channels-last Conv2d weight) ->
// bf16 K-major operands of the transposed convolution, through a 32x32 shared-memory transpose (reads coalesced along Ci,
// writes along Co).  mode 0 (stride 1): one tensor [Ci, KH, KW, Co] with the taps flipped.  mode 1 (3x3, stride 2): the
// four output-parity phases (ph, pw) in the order (0,0) (0,1) (1,0) (1,1), each [Ci, Ta, Tb, Co] with taps {1} for parity
// 0 and {2, 0} for parity 1, concatenated (sizes Ci * Co * {1, 2, 2, 4}).
__global__ void __launch_bounds__(256) k_pack_weight_t(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Co, int Ci,
                                                       int KH, int KW, int mode) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z, kh = tap / KW, kw = tap % KW;
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int co = co0 + ty + 8 * k, ci = ci0 + tx;
        tile[ty + 8 * k][tx] = (co < Co && ci < Ci) ? src[(((size_t)co * KH + kh) * KW + kw) * Ci + ci] : 0.0f;
    }
    __syncthreads();
    size_t base = 0;
    int Ta = KH, Tb = KW, a = KH - 1 - kh, b = KW - 1 - kw;
    if (mode == 1) {
        const int ph = kh == 1 ? 0 : 1, pw = kw == 1 ? 0 : 1;
        Ta = ph ? 2 : 1; Tb = pw ? 2 : 1;
        a = kh == 0 ? 1 : 0; b = kw == 0 ? 1 : 0;
        const size_t unit = (size_t)Ci * Co;
        base = (ph * 2 + pw == 0) ? 0 : (ph * 2 + pw == 1) ? unit : (ph * 2 + pw == 2) ? 3 * unit : 5 * unit;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ci = ci0 + ty + 8 * k, co = co0 + tx;
        if (ci < Ci && co < Co)
            dst[base + (((size_t)ci * Ta + a) * Tb + b) * Co + co] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
    }
}

// Backward, pass 1: per channel  s1 = sum g^,  s2 = sum g^ * zhat,  g^ = gy * [y > 0] (when the unit ends in a ReLU),
// zhat = (z - mean) * rstd.  Same block shape as k_inorm_stats (rows of one slab of channels, smem combine, one atomic
// per channel per block).
__global__ void __launch_bounds__(256) k_bn_bwd_reduce(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y,
                                                       const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, float* __restrict__ sums, long M, int C,
                                                       int rows, int groups, int relu) {
    // a thread owns 8 consecutive channels (16-byte loads) of every nrl-th row of the block's row range; `groups` = 8-channel
    // groups per slab (blockIdx.y), a divisor of 256
    __shared__ float s_part[256 * 17];
    const int gl = threadIdx.x % groups, rl = threadIdx.x / groups, nrl = 256 / groups;
    const int c = (blockIdx.y * groups + gl) * 8;
    const long r0 = (long)blockIdx.x * rows, r1 = r0 + rows < M ? r0 + rows : M;
    float a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = 0.0f; b[k] = 0.0f; }
    if (c < C) {
        float m[8], iv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { m[k] = mean[c + k]; iv[k] = rstd[c + k]; }
#pragma unroll 2
        for (long r = r0 + rl; r < r1; r += nrl) {
            const size_t e = (size_t)r * C + c;
            const uint4 gv = *reinterpret_cast<const uint4*>(gy + e);
            const uint4 zv = *reinterpret_cast<const uint4*>(z + e);
            uint4 yv = make_uint4(0, 0, 0, 0);
            if (relu) yv = *reinterpret_cast<const uint4*>(y + e);
            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, zw[4] = {zv.x, zv.y, zv.z, zv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&gw[k]);
                const __nv_bfloat162 z2 = *reinterpret_cast<const __nv_bfloat162*>(&zw[k]);
                float g0 = __low2float(g2), g1 = __high2float(g2);
                if (relu) {
                    const __nv_bfloat162 y2 = *reinterpret_cast<const __nv_bfloat162*>(&yw[k]);
                    if (!(__low2float(y2) > 0.0f)) g0 = 0.0f;
                    if (!(__high2float(y2) > 0.0f)) g1 = 0.0f;
                }
                a[2 * k] += g0; a[2 * k + 1] += g1;
                b[2 * k] = __fmaf_rn(g0, (__low2float(z2) - m[2 * k]) * iv[2 * k], b[2 * k]);
                b[2 * k + 1] = __fmaf_rn(g1, (__high2float(z2) - m[2 * k + 1]) * iv[2 * k + 1], b[2 * k + 1]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) { s_part[threadIdx.x * 17 + 2 * k] = a[k]; s_part[threadIdx.x * 17 + 2 * k + 1] = b[k]; }
    __syncthreads();
    // 16 values per group of 8 channels: thread t < groups * 16 adds the nrl partials of one value
    if (threadIdx.x < groups * 16) {
        const int g = threadIdx.x >> 4, k = threadIdx.x & 15;
        const int cc = (blockIdx.y * groups + g) * 8;
        if (cc < C) {
            float t = 0.0f;
            for (int j = 0; j < nrl; j++) t += s_part[(j * groups + g) * 17 + k];
            atomicAdd(sums + (size_t)cc * 2 + k, t);            // (channel cc + k/2, {sum, sum * zhat}[k & 1])
        }
    }
}

// Backward, pass 2: gz = gamma * rstd * (g^ - s1 / M - zhat * s2 / M); optionally also writes g^ (the gradient that flows
// into the skip connection of a BasicBlock).
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const __nv_bfloat16* __restrict__ gy, const __nv_bfloat16* __restrict__ y,
                                                      const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean,
                                                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                      const float* __restrict__ sums, __nv_bfloat16* __restrict__ gz,
                                                      __nv_bfloat16* __restrict__ gres, long groups, int C, float inv_m, int relu,
                                                      float* __restrict__ gg_acc, float* __restrict__ gb_acc) {
    const int cg = C >> 3;
    if (blockIdx.x == 0 && gg_acc)                          // parameter gradients accumulated in place (one block: no race)
        for (int c = threadIdx.x; c < C; c += blockDim.x) { gg_acc[c] += sums[2 * c + 1]; gb_acc[c] += sums[2 * c]; }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cg) * 8;
        const uint4 gv = *reinterpret_cast<const uint4*>(gy + i * 8), zv = *reinterpret_cast<const uint4*>(z + i * 8);
        uint4 yv = make_uint4(0, 0, 0, 0);
        if (relu) yv = *reinterpret_cast<const uint4*>(y + i * 8);
        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, zw[4] = {zv.x, zv.y, zv.z, zv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
        uint32_t ow[4], rw[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(&gw[k]);
            const __nv_bfloat162 z2 = *reinterpret_cast<const __nv_bfloat162*>(&zw[k]);
            float g0 = __low2float(g2), g1 = __high2float(g2);
            if (relu) {
                const __nv_bfloat162 y2 = *reinterpret_cast<const __nv_bfloat162*>(&yw[k]);
                if (!(__low2float(y2) > 0.0f)) g0 = 0.0f;
                if (!(__high2float(y2) > 0.0f)) g1 = 0.0f;
            }
            const int c0 = c + 2 * k, c1 = c0 + 1;
            const float zh0 = (__low2float(z2) - mean[c0]) * rstd[c0], zh1 = (__high2float(z2) - mean[c1]) * rstd[c1];
            const float o0 = gamma[c0] * rstd[c0] * (g0 - sums[2 * c0] * inv_m - zh0 * sums[2 * c0 + 1] * inv_m);
            const float o1 = gamma[c1] * rstd[c1] * (g1 - sums[2 * c1] * inv_m - zh1 * sums[2 * c1 + 1] * inv_m);
            const __nv_bfloat162 o2 = __floats2bfloat162_rn(o0, o1), r2 = __floats2bfloat162_rn(g0, g1);
            ow[k] = *reinterpret_cast<const uint32_t*>(&o2);
            rw[k] = *reinterpret_cast<const uint32_t*>(&r2);
        }
        *reinterpret_cast<uint4*>(gz + i * 8) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        if (gres) *reinterpret_cast<uint4*>(gres + i * 8) = make_uint4(rw[0], rw[1], rw[2], rw[3]);
    }
}

// ---- C. reflection padding of an NHWC bf16 tensor (nn.ReflectionPad2d), with optional zero channels up to Cp ----
__global__ void __launch_bounds__(256) k_reflect_pad_nhwc(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                          int N, int H, int W, int C, int pad, int Cp) {
    const int Hp = H + 2 * pad, Wp = W + 2 * pad, cg = Cp >> 3;
    const long total = (long)N * Hp * Wp * cg;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % cg);
        long r = i / cg;
        const int wp = (int)(r % Wp); r /= Wp;
        const int hp = (int)(r % Hp);
        const int n = (int)(r / Hp);
        int h = hp - pad, w = wp - pad;
        h = h < 0 ? -h : (h >= H ? 2 * (H - 1) - h : h);
        w = w < 0 ? -w : (w >= W ? 2 * (W - 1) - w : w);
        const __nv_bfloat16* src = x + (((size_t)n * H + h) * W + w) * C + g * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (g * 8 + 8 <= C && (C & 7) == 0) v = *reinterpret_cast<const uint4*>(src);
        else if (g * 8 < C) {
            __nv_bfloat16 t[8];
#pragma unroll
            for (int k = 0; k < 8; k++) t[k] = g * 8 + k < C ? src[k] : __float2bfloat16(0.0f);
            v = *reinterpret_cast<const uint4*>(t);
        }
        *reinterpret_cast<uint4*>(y + (size_t)i * 8) = v;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;      // resolved once; the driver entry point never changes
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int pick_patch(int Ho, int Wo, int N, int stride, int* bw, int* bh, int* bn, int rows = 128, int min_hw = 1) {
    // min_hw: smallest bh * bw allowed (32 when the epilogue accumulates per-image statistics: a warp's rows in one image)
    // `rows` output pixels per tile as bn x bh x bw; choose the shape that wastes the fewest rows.  TMA boxes are at most
    // 256 elements per dimension before the traversal stride: bw * stride <= 256.
    long best = -1;
    for (int w = rows; w >= 1; w >>= 1) {
        if (w * stride > 256) continue;
        for (int h = rows / w; h >= 1; h >>= 1) {
            if (h * stride > 256) continue;
            if (w * h < min_hw) continue;
            const int n = rows / (w * h);
            const long tiles = (long)((Wo + w - 1) / w) * ((Ho + h - 1) / h) * ((N + n - 1) / n);
            if (best < 0 || tiles < best) { best = tiles; *bw = w; *bh = h; *bn = n; }
        }
    }
    return best > 0 ? 0 : SDN_E_ARG;
}

}  // namespace

extern "C" int sdn_conv2d_forward(const sdn_conv2d* c, void* stream) {
    if (!c || !c->input || !c->weight || !c->output) return SDN_E_ARG;
    if (c->N <= 0 || c->H <= 0 || c->W <= 0 || c->Cin <= 0 || c->Cout <= 0 || c->KH <= 0 || c->KW <= 0 || c->stride <= 0 ||
        c->pad < 0)
        return SDN_E_ARG;
    if (c->Cin % TK != 0 || (c->Cout % 64 != 0 && c->Cout != 16)) return SDN_E_ARG;    // pad channels on the caller's side
    if ((reinterpret_cast<uintptr_t>(c->input) & 15) || (reinterpret_cast<uintptr_t>(c->weight) & 15) ||
        (reinterpret_cast<uintptr_t>(c->output) & 15) || (reinterpret_cast<uintptr_t>(c->residual) & 15) ||
        (reinterpret_cast<uintptr_t>(c->scale) & 15) || (reinterpret_cast<uintptr_t>(c->shift) & 15))
        return SDN_E_ARG;
    // output extent: the convolution's own, or the caller's (reads beyond the input are zeros: asymmetric padding)
    const int Ho = c->out_h > 0 ? c->out_h : (c->H + 2 * c->pad - c->KH) / c->stride + 1;
    const int Wo = c->out_w > 0 ? c->out_w : (c->W + 2 * c->pad - c->KW) / c->stride + 1;
    if (Ho <= 0 || Wo <= 0) return SDN_E_ARG;
    if (c->residual && (c->out_stride_w > 0)) return SDN_E_ARG;     // residual is read densely
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return (int)cudaErrorNotSupported;

    ConvParams p;
    p.N = c->N; p.H = c->H; p.W = c->W; p.Cin = c->Cin; p.Ho = Ho; p.Wo = Wo; p.Cout = c->Cout;
    p.KH = c->KH; p.KW = c->KW; p.stride = c->stride; p.pad = c->pad;
    int rc = pick_patch(Ho, Wo, c->N, c->stride, &p.bw, &p.bh, &p.bn, 128, c->stats ? 32 : 1);
    if (rc) return rc;
    p.tiles_w = (Wo + p.bw - 1) / p.bw; p.tiles_h = (Ho + p.bh - 1) / p.bh; p.tiles_n = (c->N + p.bn - 1) / p.bn;
    p.out = (__nv_bfloat16*)c->output; p.scale = c->scale; p.shift = c->shift;
    if (c->out_stride_w > 0) { p.osn = c->out_stride_n; p.osh = c->out_stride_h; p.osw = c->out_stride_w; }
    else { p.osw = c->Cout; p.osh = (long)Wo * c->Cout; p.osn = (long)Ho * Wo * c->Cout; }
    if ((p.osn | p.osh | p.osw) & 7) return SDN_E_ARG;              // 16-byte stores
    p.residual = (const __nv_bfloat16*)c->residual; p.relu = c->relu;
    p.stats = c->stats;
    if (p.stats) {
        if (c->residual) return SDN_E_ARG;                          // statistics are of scale * conv + shift
        if ((p.bh * p.bw) % 32) return SDN_E_ARG;                   // a warp's 32 rows must lie in one image
        if (!c->stats_keep) {
            const cudaError_t em = cudaMemsetAsync(p.stats, 0, sizeof(float) * 2 * (size_t)c->N * c->Cout, (cudaStream_t)stream);
            if (em != cudaSuccess) return (int)em;
        }
    }

    // large filters on pre-padded 64-channel inputs (the generator's 7x7 layers): row-ring form, OPT-IN (SDN_CONV_ROWS=1).
    // Correct (tests/test_gpu_conv_tc.py::test_large_filters_row_ring_form) and, with two MMA-issuing threads, 348 us against
    // 380 us in the tile form for the 64-channel stem in isolation, but 332 against 293 us for the 16-column head and no
    // gain inside the network (statistics epilogue): the tile form stays the default.
    {
        const char* rows_env = getenv("SDN_CONV_ROWS");
        const int rows_mode = rows_env ? atoi(rows_env) : 0;
        const bool rows_ok = c->KH == c->KW && c->KH * c->KW >= 25 && c->KW <= 9 && c->stride == 1 && c->pad == 0 && c->Cin == 64 &&
                             !c->residual && c->out_stride_w <= 0 && c->out_h <= 0 && c->out_w <= 0 &&
                             (c->Cout == 64 || c->Cout == 16) && Wo >= RW_PX &&
                             rows_mode == 1;
        if (rows_ok) {
            const int rowpx = RW_PX + c->KW - 1;
            const int slot_bytes = (rowpx * 128 + 1023) / 1024 * 1024;
            const int nslots = c->KH + 2 * RW_RG - 1;
            const int TNr = c->Cout;
            const size_t smem_r = (size_t)nslots * slot_bytes + (size_t)RW_BST * TNr * TK * 2 + 4 * 32 * 17 * sizeof(float) + 1024;
            if (nslots <= RW_MAXSLOTS && smem_r <= 220 * 1024) {
                alignas(64) CUtensorMap tmAr, tmBr;
                const cuuint64_t dimsA[4] = {(cuuint64_t)c->Cin, (cuuint64_t)c->W, (cuuint64_t)c->H, (cuuint64_t)c->N};
                const cuuint64_t strA[3] = {(cuuint64_t)c->Cin * 2, (cuuint64_t)c->W * c->Cin * 2, (cuuint64_t)c->H * c->W * c->Cin * 2};
                const cuuint32_t boxA[4] = {(cuuint32_t)TK, (cuuint32_t)rowpx, 1, 1};
                const cuuint32_t one4[4] = {1, 1, 1, 1};
                if (enc(&tmAr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c->input, dimsA, strA, boxA, one4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return SDN_E_ARG;
                const cuuint64_t Kr = (cuuint64_t)c->KH * c->KW * c->Cin;
                const cuuint64_t dimsB[2] = {Kr, (cuuint64_t)c->Cout};
                const cuuint64_t strB[1] = {Kr * 2};
                const cuuint32_t boxB[2] = {(cuuint32_t)TK, (cuuint32_t)TNr};
                const cuuint32_t one2[2] = {1, 1};
                if (enc(&tmBr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)c->weight, dimsB, strB, boxB, one2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                    return SDN_E_ARG;
                int dev_r = 0, sms_r = 148;
                if (cudaGetDevice(&dev_r) == cudaSuccess) cudaDeviceGetAttribute(&sms_r, cudaDevAttrMultiProcessorCount, dev_r);
                const long work = (long)c->N * ((Wo + RW_PX - 1) / RW_PX) * ((Ho + RW_SEG - 1) / RW_SEG);
                const dim3 grid_r((unsigned)(work < sms_r ? work : sms_r), 1);
                cudaStream_t st_r = (cudaStream_t)stream;
                cudaError_t er;
                if (TNr == 64) {
                    er = cudaFuncSetAttribute(k_conv_rows<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
                    if (er) return (int)er;
                    sdn_internal::launch_begin("k_conv_rows", st_r);
                    k_conv_rows<64><<<grid_r, RW_THREADS, smem_r, st_r>>>(tmAr, tmBr, p, slot_bytes, nslots);
                } else {
                    er = cudaFuncSetAttribute(k_conv_rows<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
                    if (er) return (int)er;
                    sdn_internal::launch_begin("k_conv_rows", st_r);
                    k_conv_rows<16><<<grid_r, RW_THREADS, smem_r, st_r>>>(tmAr, tmBr, p, slot_bytes, nslots);
                }
                return sdn_internal::launch_end();
            }
        }
    }
    int TN = c->Cout == 16 ? 16 : ((c->Cout % 128 == 0) ? 128 : 64);
    {   // few output pixels (ResNet layer3 / layer4 at small batch): 128-wide N tiles would occupy less than half of the SMs;
        // 64-wide ones double the CTAs (the A tile is re-read from L2 by twice as many CTAs, which is cheaper than idle SMs)
        int dev_ = 0, sms_ = 148;
        if (cudaGetDevice(&dev_) == cudaSuccess) cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, dev_);
        static const bool no_narrow = getenv("SDN_CONV_NO_NARROW") != nullptr;
        if (TN == 128 && !no_narrow && (long)p.tiles_w * p.tiles_h * p.tiles_n * (c->Cout / 128) * 2 <= sms_) TN = 64;
        // many output channels: 256-wide tiles read the activation tile half as often; taken when they also save a wave
        // (e.g. 1024 channels on 24 m tiles: 96 items in one wave instead of 192 in two)
        static const char* force_tn = getenv("SDN_CONV_TN");
        const long items128 = (long)p.tiles_w * p.tiles_h * p.tiles_n * (c->Cout / 128);
        if (TN == 128 && c->Cout % 256 == 0 && (items128 + sms_ - 1) / sms_ > (items128 / 2 + sms_ - 1) / sms_ &&
            (items128 + sms_ - 1) / sms_ <= 4)
            TN = 256;
        if (force_tn && atoi(force_tn) == 128 && TN == 256) TN = 128;
        if (force_tn && atoi(force_tn) == 256 && TN == 128 && c->Cout % 256 == 0) TN = 256;
    }
    // CTA-pair form (cta_group::2): a 256 x 256 tile over two SMs, each loading its own pixels and half of the weight tile
    bool pair = false;
    {
        static const char* pair_env = getenv("SDN_CONV_PAIR");
        const long tiles_m_ = (long)p.tiles_w * p.tiles_h * p.tiles_n;
        const bool can = c->Cout % 256 == 0 && tiles_m_ >= 2;
        const bool want = pair_env ? atoi(pair_env) != 0 : (TN == 256);
        pair = can && want;
        if (pair) TN = 256;
    }
    alignas(64) CUtensorMap tmA, tmB;
    {   // activations: dims (fastest first) {C, W, H, N}
        const cuuint64_t dims[4] = {(cuuint64_t)c->Cin, (cuuint64_t)c->W, (cuuint64_t)c->H, (cuuint64_t)c->N};
        const cuuint64_t strides[3] = {(cuuint64_t)c->Cin * 2, (cuuint64_t)c->W * c->Cin * 2,
                                       (cuuint64_t)c->H * c->W * c->Cin * 2};
        const cuuint32_t box[4] = {(cuuint32_t)TK, (cuuint32_t)(p.bw * c->stride), (cuuint32_t)(p.bh * c->stride),
                                   (cuuint32_t)p.bn};
        const cuuint32_t estr[4] = {1, (cuuint32_t)c->stride, (cuuint32_t)c->stride, 1};
        const CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)c->input, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return SDN_E_ARG;
    }
    {   // weights: [Cout, KH*KW*Cin], dims {K, Cout}
        const cuuint64_t K = (cuuint64_t)c->KH * c->KW * c->Cin;
        const cuuint64_t dims[2] = {K, (cuuint64_t)c->Cout};
        const cuuint64_t strides[1] = {K * 2};
        const cuuint32_t box[2] = {(cuuint32_t)TK, (cuuint32_t)(pair ? TN / 2 : TN)};     // pair: each CTA loads half
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)c->weight, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return SDN_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long tiles_m_all = (long)p.tiles_w * p.tiles_h * p.tiles_n;
    const long total_work = (pair ? (tiles_m_all + 1) / 2 : tiles_m_all) * (c->Cout / TN);
    const int stages = pair ? 6 : (TN == 128 ? Ring<128>::STAGES : Ring<64>::STAGES);
    const size_t smem = (size_t)stages * (A_BYTES + (pair ? TN / 2 : TN) * TK * 2) + 4 * 32 * 17 * sizeof(float) + 1024;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = smem * 2 + 4096 <= 227 * 1024 ? 2 : 1;      // CTAs that fit one SM's shared memory
    const dim3 grid((unsigned)(total_work < (long)sms * per_sm ? total_work : (long)sms * per_sm), 1);
    cudaError_t e;
    if (pair) {
        e = cudaFuncSetAttribute(k_conv_tc<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        const long pairs = total_work < sms / 2 ? total_work : sms / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
        cfg.blockDim = dim3(NTHREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        sdn_internal::launch_begin("k_conv_tc_pair", st);
        e = cudaLaunchKernelEx(&cfg, k_conv_tc<256, true>, tmA, tmB, p);
        if (e != cudaSuccess) { sdn_internal::launch_end(); return (int)e; }
        return sdn_internal::launch_end();
    }
    if (TN == 256) {
        e = cudaFuncSetAttribute(k_conv_tc<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        sdn_internal::launch_begin("k_conv_tc", st);
        k_conv_tc<256, false><<<grid, NTHREADS, smem, st>>>(tmA, tmB, p);
    } else if (TN == 128) {
        e = cudaFuncSetAttribute(k_conv_tc<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        sdn_internal::launch_begin("k_conv_tc", st);
        k_conv_tc<128, false><<<grid, NTHREADS, smem, st>>>(tmA, tmB, p);
    } else if (TN == 64) {
        e = cudaFuncSetAttribute(k_conv_tc<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        sdn_internal::launch_begin("k_conv_tc", st);
        k_conv_tc<64, false><<<grid, NTHREADS, smem, st>>>(tmA, tmB, p);
    } else {
        e = cudaFuncSetAttribute(k_conv_tc<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        sdn_internal::launch_begin("k_conv_tc", st);
        k_conv_tc<16, false><<<grid, NTHREADS, smem, st>>>(tmA, tmB, p);
    }
    return sdn_internal::launch_end();
}

extern "C" int sdn_im2col_nhwc(const void* input, void* output, int N, int H, int W, int Cin, int KH, int KW, int stride,
                               int pad, int Kpad, void* stream) {
    if (!input || !output || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || pad < 0 ||
        Kpad < KH * KW * Cin)
        return SDN_E_ARG;
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return SDN_E_ARG;
    if (Kpad & 7) return SDN_E_ARG;
    const long total = (long)N * Ho * Wo * (Kpad / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    const size_t row_bytes = (size_t)KH * W * Cin * 2;
    if (row_bytes <= 48 * 1024 && Wo * (Kpad >> 3) >= 256) {
        sdn_internal::launch_begin("k_im2col_rows", (cudaStream_t)stream);
        k_im2col_rows<<<(unsigned)(N * Ho), 256, row_bytes, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)input, (__nv_bfloat16*)output, H, W, Cin, Ho, Wo, KH, KW, stride, pad, Kpad);
        return sdn_internal::launch_end();
    }
    sdn_internal::launch_begin("k_im2col_nhwc", (cudaStream_t)stream);
    k_im2col_nhwc<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)input, (__nv_bfloat16*)output, N, H, W, Cin,
                                                            Ho, Wo, KH, KW, stride, pad, Kpad);
    return sdn_internal::launch_end();
}

extern "C" int sdn_instance_norm_nhwc(const void* x, const void* residual, void* y, float* stats, int N, int HW, int C,
                                      float eps, int act, void* stream) {
    if (!x || !y || !stats || N <= 0 || HW <= 0 || C <= 0 || (C & 7)) return SDN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(float) * 2 * (size_t)N * C, st);
    if (e != cudaSuccess) return (int)e;
    int pairs = C / 2;                      // channel pairs per block: a power-of-two divisor of 256 where possible
    if (pairs > 256) pairs = 256;
    while (256 % pairs) pairs--;
    const int slabs = (C / 2 + pairs - 1) / pairs;
    // rows per block: enough blocks to fill the machine (~8 per SM), at least one trip of the block's row lanes
    const int nrl = 256 / pairs;
    long want = ((long)HW * slabs * N + 148L * 8 - 1) / (148L * 8);
    // (every block ends with one atomicAdd per channel: at least 32 rows per row lane keep that traffic negligible)
    int rows = (int)(want < 32L * nrl ? 32L * nrl : (want > 8192 ? 8192 : want));
    rows = (rows + nrl - 1) / nrl * nrl;
    const dim3 grid((unsigned)((HW + rows - 1) / rows), (unsigned)slabs, (unsigned)N);
    sdn_internal::launch_begin("k_inorm_stats", st);
    k_inorm_stats<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, stats, HW, C, rows, pairs);
    int rc = sdn_internal::launch_end();
    if (rc) return rc;
    const long total = (long)HW * (C / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148L * 16 ? 148L * 16 : (total + 255) / 256);
    sdn_internal::launch_begin("k_inorm_apply", st);
    k_inorm_apply<<<dim3(blocks, (unsigned)N), 256, 0, st>>>((const __nv_bfloat16*)x, stats, (const __nv_bfloat16*)residual,
                                                            (__nv_bfloat16*)y, HW, C, eps, act);
    return sdn_internal::launch_end();
}

extern "C" int sdn_instance_norm_apply_nhwc(const void* x, const void* residual, void* y, const float* stats, int N, int HW,
                                            int C, float eps, int act, void* stream) {
    if (!x || !y || !stats || N <= 0 || HW <= 0 || C <= 0 || (C & 7)) return SDN_E_ARG;
    const long total = (long)HW * (C / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148L * 16 ? 148L * 16 : (total + 255) / 256);
    sdn_internal::launch_begin("k_inorm_apply", (cudaStream_t)stream);
    k_inorm_apply<<<dim3(blocks, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, stats, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, HW, C, eps, act);
    return sdn_internal::launch_end();
}

extern "C" int sdn_instance_norm_apply_pad_nhwc(const void* x, const void* residual, void* y, const float* stats, int N, int H,
                                                int W, int C, float eps, int act, int pad, void* stream) {
    if (!x || !y || !stats || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || pad < 0 || pad >= H || pad >= W) return SDN_E_ARG;
    const long total = (long)(H + 2 * pad) * (W + 2 * pad) * (C / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148L * 16 ? 148L * 16 : (total + 255) / 256);
    sdn_internal::launch_begin("k_inorm_apply_pad", (cudaStream_t)stream);
    k_inorm_apply_pad<<<dim3(blocks, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, stats, (const __nv_bfloat16*)residual, (__nv_bfloat16*)y, H, W, C, eps, act, pad);
    return sdn_internal::launch_end();
}

extern "C" int sdn_reflect_pad_nhwc(const void* x, void* y, int N, int H, int W, int C, int pad, int Cp, void* stream) {
    if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || pad < 0 || pad >= H || pad >= W || Cp < C || (Cp & 7))
        return SDN_E_ARG;
    const long total = (long)N * (H + 2 * pad) * (W + 2 * pad) * (Cp / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148L * 32 ? 148L * 32 : (total + 255) / 256);
    sdn_internal::launch_begin("k_reflect_pad_nhwc", (cudaStream_t)stream);
    k_reflect_pad_nhwc<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C, pad, Cp);
    return sdn_internal::launch_end();
}

extern "C" int sdn_conv2d_wgrad(const void* grad_out, const void* input, float* grad_weight, int N, int H, int W, int Cin,
                                int Cout, int KH, int KW, int stride, int pad, int accumulate, void* stream) {
    if (!grad_out || !input || !grad_weight || N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 ||
        stride <= 0 || pad < 0)
        return SDN_E_ARG;
    if (Cin % 64 != 0 || Cout % 64 != 0) return SDN_E_ARG;
    if ((reinterpret_cast<uintptr_t>(grad_out) & 15) || (reinterpret_cast<uintptr_t>(input) & 15)) return SDN_E_ARG;
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return SDN_E_ARG;
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return (int)cudaErrorNotSupported;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSuccess;

    WgradParams p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
    int rc = pick_patch(Ho, Wo, N, stride, &p.bw, &p.bh, &p.bn, 64);
    if (rc) return rc;
    p.tiles_w = (Wo + p.bw - 1) / p.bw; p.tiles_h = (Ho + p.bh - 1) / p.bh; p.tiles_n = (N + p.bn - 1) / p.bn;
    p.dw = grad_weight;
    const int TNW = Cin >= 128 ? 128 : 64;
    const int patches = p.tiles_w * p.tiles_h * p.tiles_n;
    const int base_items = ((Cout + 127) / 128) * ((Cin + TNW - 1) / TNW) * KH * KW;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // how many pixel ranges: minimise  waves * (K blocks per item + epilogue), the epilogue (128 x TNW fp32 through L2, atomics
    // when there is more than one range) costing about as much as 4 K blocks; one range = plain stores, no clearing pass
    {
        long best_cost = -1; int best = 1;
        for (int sp = 1; sp <= patches && sp <= 256; sp++) {
            const int pps = (patches + sp - 1) / sp, real = (patches + pps - 1) / pps;
            if (real != sp) continue;
            const long waves = ((long)base_items * sp + sms - 1) / sms;
            const long cost = waves * (pps + (sp == 1 ? 3 : 4));
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
        }
        p.splits = best;
        p.patches_per_split = (patches + best - 1) / best;
    }
    p.atomic = (p.splits > 1 || accumulate) ? 1 : 0;
    if (p.splits > 1 && !accumulate) {
        e = cudaMemsetAsync(grad_weight, 0, sizeof(float) * (size_t)Cout * KH * KW * Cin, st);
        if (e != cudaSuccess) return (int)e;
    }

    alignas(64) CUtensorMap tmG, tmX;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)N};
        const cuuint64_t strides[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wo * Cout * 2, (cuuint64_t)Ho * Wo * Cout * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bn};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        if (enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)grad_out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return SDN_E_ARG;
    }
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)(p.bw * stride), (cuuint32_t)(p.bh * stride), (cuuint32_t)p.bn};
        const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        if (enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)input, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return SDN_E_ARG;
    }
    const long total_work = (long)base_items * p.splits;
    const size_t smem = (size_t)6 * (16384 + (TNW / 64) * 8192) + 1024;     // + 18.6 KB static (barriers, transpose)
    const dim3 grid((unsigned)(total_work < sms ? total_work : sms), 1);
    if (TNW == 128) {
        e = cudaFuncSetAttribute(k_conv_wgrad<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        sdn_internal::launch_begin("k_conv_wgrad", st);
        k_conv_wgrad<128><<<grid, NTHREADS, smem, st>>>(tmG, tmX, p);
    } else {
        e = cudaFuncSetAttribute(k_conv_wgrad<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        sdn_internal::launch_begin("k_conv_wgrad", st);
        k_conv_wgrad<64><<<grid, NTHREADS, smem, st>>>(tmG, tmX, p);
    }
    return sdn_internal::launch_end();
}

extern "C" int sdn_max_pool_nhwc(const void* x, void* y, void* argmax, int N, int H, int W, int C, int K, int stride, int pad,
                                 void* stream) {
    if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || K <= 0 || K > 15 || stride <= 0 || pad < 0 || 2 * pad > K)
        return SDN_E_ARG;
    const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return SDN_E_ARG;
    const long total = (long)N * Ho * Wo * (C / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148L * 32 ? 148L * 32 : (total + 255) / 256);
    sdn_internal::launch_begin("k_max_pool_nhwc", (cudaStream_t)stream);
    k_max_pool_nhwc<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, (unsigned char*)argmax, N, H,
                                                              W, C, Ho, Wo, K, stride, pad);
    return sdn_internal::launch_end();
}

extern "C" int sdn_max_pool_backward_nhwc(const void* argmax, const void* grad_y, void* grad_x, int N, int H, int W, int C, int K,
                                          int stride, int pad, void* stream) {
    if (!argmax || !grad_y || !grad_x || N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || K <= 0 || K > 15 || stride <= 0 ||
        pad < 0 || 2 * pad > K)
        return SDN_E_ARG;
    const int Ho = (H + 2 * pad - K) / stride + 1, Wo = (W + 2 * pad - K) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return SDN_E_ARG;
    const long total = (long)N * H * W * (C / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148L * 32 ? 148L * 32 : (total + 255) / 256);
    sdn_internal::launch_begin("k_max_pool_bwd_nhwc", (cudaStream_t)stream);
    k_max_pool_bwd_nhwc<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)argmax, (const __nv_bfloat16*)grad_y,
                                                                  (__nv_bfloat16*)grad_x, N, H, W, C, Ho, Wo, K, stride, pad);
    return sdn_internal::launch_end();
}

extern "C" int sdn_batch_norm_finalize(const float* stats, int N, int C, long long pixels, float eps, float momentum,
                                       const float* gamma, const float* beta, float* running_mean, float* running_var,
                                       long long* num_batches, float* out, void* stream) {
    if (!stats || !gamma || !beta || !out || N <= 0 || C <= 0 || pixels <= 0) return SDN_E_ARG;
    sdn_internal::launch_begin("k_bn_finalize", (cudaStream_t)stream);
    k_bn_finalize<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, N, C, (double)pixels, eps, momentum, gamma, beta,
                                                                    running_mean, running_var, num_batches, out);
    return sdn_internal::launch_end();
}

extern "C" int sdn_pack_conv_weight_transposed(const float* weight, void* packed, int Cout, int Cin, int KH, int KW, int stride,
                                               void* stream) {
    if (!weight || !packed || Cout <= 0 || Cin <= 0 || KH <= 0 || KW <= 0) return SDN_E_ARG;
    int mode = 0;
    if (stride == 2 && KH == 3 && KW == 3) mode = 1;
    else if (stride != 1 && !(KH == 1 && KW == 1)) return SDN_E_ARG;
    const dim3 grid((unsigned)((Cin + 31) / 32), (unsigned)((Cout + 31) / 32), (unsigned)(KH * KW));
    sdn_internal::launch_begin("k_pack_weight_t", (cudaStream_t)stream);
    k_pack_weight_t<<<grid, 256, 0, (cudaStream_t)stream>>>(weight, (__nv_bfloat16*)packed, Cout, Cin, KH, KW, mode);
    return sdn_internal::launch_end();
}

extern "C" int sdn_affine_act_nhwc(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                                   long long pixels, int C, int act, void* stream) {
    if (!z || !scale || !shift || !y || pixels <= 0 || C <= 0 || (C & 7)) return SDN_E_ARG;
    const long groups = (long)pixels * (C / 8);
    const unsigned blocks = (unsigned)((groups + 255) / 256 > 148L * 16 ? 148L * 16 : (groups + 255) / 256);
    sdn_internal::launch_begin("k_affine_act", (cudaStream_t)stream);
    k_affine_act<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)z, scale, shift, (const __nv_bfloat16*)residual,
                                                           (__nv_bfloat16*)y, groups, C, act);
    return sdn_internal::launch_end();
}

extern "C" int sdn_batch_norm_backward_nhwc(const void* grad_y, const void* y, const void* z, const float* mean, const float* rstd,
                                            const float* gamma, float* sums, void* grad_z, void* grad_skip,
                                            float* grad_gamma_acc, float* grad_beta_acc, long long pixels, int C, int relu,
                                            void* stream) {
    if (!grad_y || !z || !mean || !rstd || !gamma || !sums || !grad_z || pixels <= 0 || C <= 0 || (C & 7) || ((relu & 1) && !y) ||
        ((grad_gamma_acc == nullptr) != (grad_beta_acc == nullptr)))
        return SDN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const bool precleared = (relu & 2) != 0;             // bit 1: the caller cleared `sums` (one clearing pass for many layers)
    relu &= 1;
    if (!precleared) {
        cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * 2 * (size_t)C, st);
        if (e != cudaSuccess) return (int)e;
    }
    int grp = C / 8;                                    // 8-channel groups per block: a divisor of 256, at most 16
    if (grp > 16) grp = 16;
    while (256 % grp || (C / 8) % grp) grp--;
    const int slabs = (C / 8) / grp, nrl = 256 / grp;
    long want = ((long)pixels * slabs + 148L * 4 - 1) / (148L * 4);        // ~4 blocks per SM
    int rows = (int)(want < 4L * nrl ? 4L * nrl : (want > 16384 ? 16384 : want));
    rows = (rows + nrl - 1) / nrl * nrl;
    sdn_internal::launch_begin("k_bn_bwd_reduce", st);
    k_bn_bwd_reduce<<<dim3((unsigned)((pixels + rows - 1) / rows), (unsigned)slabs), 256, 0, st>>>(
        (const __nv_bfloat16*)grad_y, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, mean, rstd, sums, (long)pixels, C, rows,
        grp, relu);
    int rc = sdn_internal::launch_end();
    if (rc) return rc;
    const long groups = (long)pixels * (C / 8);
    const unsigned blocks = (unsigned)((groups + 255) / 256 > 148L * 16 ? 148L * 16 : (groups + 255) / 256);
    sdn_internal::launch_begin("k_bn_bwd_apply", st);
    k_bn_bwd_apply<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)grad_y, (const __nv_bfloat16*)y, (const __nv_bfloat16*)z, mean, rstd,
                                           gamma, sums, (__nv_bfloat16*)grad_z, (__nv_bfloat16*)grad_skip, groups, C,
                                           1.0f / (float)pixels, relu, grad_gamma_acc, grad_gamma_acc ? grad_beta_acc : nullptr);
    return sdn_internal::launch_end();
}
