// sdn_linear_tc.cu -- Y[M,N] = act(X[M,K] . W[N,K]^T + b) on Blackwell's 5th-generation tensor cores (tcgen05),
// hand-written for sm_100a: bf16 operands staged in shared memory in the canonical K-major "interleave" (no swizzle)
// UMMA layout, fp32 accumulators in tensor memory (TMEM), one elected thread issues tcgen05.mma, completion comes back
// through tcgen05.commit on an mbarrier, the epilogue reads TMEM with tcgen05.ld and fuses bias + ReLU.
//
// It serves the fully connected stack of the encoder (reference: derender3d/models/derenderer.py:26-32,35-43:
// net.fc 512->256, fc1 260->256, fc2 256->256, _fc3 256->1552).  PyTorch's Linear keeps W as [out, in] = [N, K] with
// K contiguous, which IS the K-major B operand UMMA wants, so no transpose is needed.
//
// Tile: 128 (M) x 128 (N) per CTA, K in blocks of 64 (four K=16 UMMAs per block).  Operand tile in shared memory
// (128 rows x 64 bf16): 16-byte chunk c (8 bf16 along K) of row r lives at  c * 2048 + r * 16  bytes, i.e. in the
// descriptor's units of 16 bytes: row stride 1, 8-row-group stride SBO = 8, K-chunk stride LBO = 128
// (cute/atom/mma_traits_sm100.hpp: "LayoutType::INTERLEAVE : ((8,n),2):((1,SBO),LBO)").
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "sdn_b200.h"

namespace sdn_internal {     // sdn_raster.cu: launch counter + optional per-kernel event timing
void launch_begin(const char* name, cudaStream_t st);
int launch_end();
}  // namespace sdn_internal

namespace {

constexpr int TM = 128, TN = 128, TK = 64;
constexpr int TILE_BYTES = TM * TK * 2;            // 16 KB per operand tile
constexpr uint32_t TMEM_COLS = 128;                // fp32 accumulator columns (= TN)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp, union SmemDescriptor)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr_bytes & 0x3FFFFu) >> 4);             // start address, bits [0,14)
    d |= (uint64_t)(2048u >> 4) << 16;                         // leading byte offset (K-chunk stride), bits [16,30)
    d |= (uint64_t)(128u >> 4) << 32;                          // stride byte offset (8-row group stride), bits [32,46)
    d |= (uint64_t)1 << 46;                                    // version = 1 (Blackwell), bits [46,48)
    return d;                                                  // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// instruction descriptor (union InstrDescriptor): D = F32, A = B = BF16, K-major both, N = 128, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITL_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEL_%=;\n\t"
        "bra WAITL_%=;\n\t"
        "DONEL_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// 8 consecutive K elements of row `row` (fp32 in global memory) -> 8 bf16 (16 bytes); out-of-range -> 0
__device__ __forceinline__ uint4 load_chunk_bf16(const float* __restrict__ base, int row, int nrows, int k, int K,
                                                 size_t ld) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0.0f;
    if (row < nrows) {
        const float* p = base + (size_t)row * ld + k;
        if (k + 8 <= K && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) if (k + i < K) v[i] = p[i];
        }
    }
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
    o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
    return o;
}

__global__ void __launch_bounds__(128) k_linear_tc(const float* __restrict__ X, const float* __restrict__ Wt,
                                                   const float* __restrict__ bias, float* __restrict__ Y, int M, int N,
                                                   int K, int relu) {
    extern __shared__ __align__(128) unsigned char smem[];      // 2 stages x (A tile | B tile)
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;

    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    if (warp == 0) {        // one warp allocates the accumulator columns in tensor memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    // Two-stage ring.  Staging map: per step a warp covers 8 rows x 4 chunks -- lane l takes row (l & 7), chunk
    // (l >> 3) -- so a global request touches 8 lines of 128 B (one per row) and a quarter warp's 16-byte shared
    // stores hit 8 different rows of one chunk (128 contiguous bytes: conflict free).  The global loads of block kb
    // are issued before waiting for the UMMAs of block kb-2 (same stage): they fly while kb-1 is in the tensor core.
    const int kblocks = (K + TK - 1) / TK;
    const int lr = lane & 7, lc = lane >> 3;
    for (int kb = 0; kb < kblocks; kb++) {
        const int k0 = kb * TK, st = kb & 1;
        unsigned char* sA = smem + st * 2 * TILE_BYTES;
        unsigned char* sB = sA + TILE_BYTES;
        uint4 ra[8], rb[8];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int blk = it * 4 + warp, r = (blk >> 1) * 8 + lr, c = (blk & 1) * 4 + lc;
            ra[it] = load_chunk_bf16(X, m0 + r, M, k0 + 8 * c, K, (size_t)K);
            rb[it] = load_chunk_bf16(Wt, n0 + r, N, k0 + 8 * c, K, (size_t)K);
        }
        if (kb >= 2) {      // the UMMAs that read this stage two blocks ago must have drained
            mbar_wait(&s_bar[st], (uint32_t)(((kb >> 1) - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int blk = it * 4 + warp, r = (blk >> 1) * 8 + lr, c = (blk & 1) * 4 + lc;
            *reinterpret_cast<uint4*>(sA + c * 2048 + r * 16) = ra[it];
            *reinterpret_cast<uint4*>(sB + c * 2048 + r * 16) = rb[it];
        }
        // generic-proxy writes must be visible to the tensor core's async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_addr(sA), b0 = smem_addr(sB);
#pragma unroll
            for (int j = 0; j < TK / 16; j++) {     // one UMMA consumes K = 16 = two 16-byte chunks per row
                const uint64_t da = make_smem_desc(a0 + j * 2 * 2048), db = make_smem_desc(b0 + j * 2 * 2048);
                const uint32_t accumulate = (kb > 0 || j > 0) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
            }
            // arrives on the mbarrier when every UMMA issued so far has completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_addr(&s_bar[st])) : "memory");
        }
    }
    {   // the last commit covers every UMMA of this CTA
        const int last = kblocks - 1;
        mbar_wait(&s_bar[last & 1], (uint32_t)((last >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- epilogue: warp w owns TMEM lanes (= tile rows) 32w .. 32w+31 and reads 32 columns per tcgen05.ld (one
    // row per lane).  The 32x32 block is transposed through shared memory (the operand ring is free now) so that the
    // global stores are row-contiguous: lane = column, 128 B per row.
    float* sT = reinterpret_cast<float*>(smem) + warp * (32 * 33);
#pragma unroll 1
    for (int c0 = 0; c0 < TN; c0 += 32) {
        if (n0 + c0 >= N) break;
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
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
#pragma unroll
        for (int j = 0; j < 32; j++) sT[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int col = n0 + c0 + lane;
        if (col < N) {
            const float bv = bias ? bias[col] : 0.0f;
            const int rows = min(32, M - (m0 + warp * 32));
            for (int rr = 0; rr < rows; rr++) {
                float y = sT[rr * 33 + lane] + bv;
                if (relu) y = fmaxf(y, 0.0f);
                Y[(size_t)(m0 + warp * 32 + rr) * N + col] = y;
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace

extern "C" int sdn_linear_forward(const float* x, const float* weight, const float* bias, float* y, int M, int N,
                                  int K, int relu, void* stream) {
    if (!x || !weight || !y || M <= 0 || N <= 0 || K <= 0) return SDN_E_ARG;
    const dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((M + TM - 1) / TM));
    const size_t smem = 4 * TILE_BYTES;     // two stages of (A tile | B tile)
    cudaError_t e = cudaFuncSetAttribute(k_linear_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    sdn_internal::launch_begin("k_linear_tc", (cudaStream_t)stream);
    k_linear_tc<<<grid, 128, smem, (cudaStream_t)stream>>>(x, weight, bias, y, M, N, K, relu);
    return sdn_internal::launch_end();
}

// ------------------------------------------------------------------------------------------------------------
// The whole encoder head in ONE launch (derender3d/models/derenderer.py:26-32,35-43: net.fc -> ReLU -> cat(mroi, droi) ->
// fc1 -> ReLU -> fc2 -> ReLU -> fc3).  At the encoder's batch sizes (M = 16 objects; M <= 64 here) a 128-row tensor-core
// tile is an eighth full and each of the four k_linear_tc launches is ~37 us of latency, so this path does not use tensor
// cores at all: it is fp32 FMA work spread over 32 CTAs, a warp per output feature, the 16-row activation chunk in
// shared memory, the four layers separated by grid-wide barriers (all CTAs are co-resident: 32 CTAs on 148 SMs).  fp32
// operands and accumulation -- the reference's arithmetic, not bf16.  The layer outputs are written to caller buffers (the
// backward needs them).
// ------------------------------------------------------------------------------------------------------------
namespace {

constexpr int HEAD_CTAS = 32;
constexpr int HEAD_ROWS = 16;           // activation rows staged per pass
constexpr int HEAD_MAXK = 1024;         // widest layer input (net.fc output 1000 + 4 extras)

__device__ __forceinline__ void head_grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned*>(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

// out[m, f] = act(sum_k in[m, k] * W[f, k] + b[f]) for all m < M, features strided over the grid's warps
__device__ void head_layer(const float* in, int ldi, int K, const float* __restrict__ W, const float* __restrict__ b,
                           float* __restrict__ out, int ldo, int N, int M, bool relu, float* s_in) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int m0 = 0; m0 < M; m0 += HEAD_ROWS) {
        const int rows = min(HEAD_ROWS, M - m0);
        __syncthreads();                                        // the previous chunk's readers are done
        for (int i = threadIdx.x; i < HEAD_ROWS * K; i += blockDim.x) {
            const int r = i / K, k = i - r * K;
            // L2 loads: for layers 1..3 `in` was written by other CTAs of this launch (before the grid barrier)
            s_in[i] = r < rows ? __ldcg(in + (size_t)(m0 + r) * ldi + k) : 0.0f;
        }
        __syncthreads();
        for (int f = blockIdx.x * nwarps + warp; f < N; f += gridDim.x * nwarps) {
            float acc[HEAD_ROWS];
#pragma unroll
            for (int r = 0; r < HEAD_ROWS; r++) acc[r] = 0.0f;
            const float* w = W + (size_t)f * K;
            for (int k = lane; k < K; k += 32) {
                const float wv = __ldg(w + k);
#pragma unroll
                for (int r = 0; r < HEAD_ROWS; r++) acc[r] = __fmaf_rn(wv, s_in[r * K + k], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < HEAD_ROWS; r++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
            }
            if (lane < rows) {
                float v = 0.0f;
#pragma unroll
                for (int r = 0; r < HEAD_ROWS; r++) if (r == lane) v = acc[r];
                v += b ? b[f] : 0.0f;
                if (relu) v = fmaxf(v, 0.0f);
                out[(size_t)(m0 + lane) * ldo + f] = v;
            }
        }
    }
}

struct HeadParams {
    const float* x; const float* extra;
    int M, K0, N0, E, N1, N2, N3;
    const float *W0, *b0, *W1, *b1, *W2, *b2, *W3, *b3;
    float *cat, *h1, *h2, *y;
    unsigned* barrier;
};

__global__ void __launch_bounds__(256) k_mlp_head(const HeadParams p) {
    extern __shared__ float s_head_in[];
    const unsigned G = gridDim.x;
    // layer 0 -> cat[:, :N0]; the extras go to cat[:, N0:]
    head_layer(p.x, p.K0, p.K0, p.W0, p.b0, p.cat, p.N0 + p.E, p.N0, p.M, true, s_head_in);
    if (p.E > 0)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.M * p.E; i += gridDim.x * blockDim.x) {
            const int m = i / p.E, e = i - m * p.E;
            p.cat[(size_t)m * (p.N0 + p.E) + p.N0 + e] = p.extra[i];
        }
    head_grid_barrier(p.barrier, G);
    head_layer(p.cat, p.N0 + p.E, p.N0 + p.E, p.W1, p.b1, p.h1, p.N1, p.N1, p.M, true, s_head_in);
    head_grid_barrier(p.barrier, 2 * G);
    head_layer(p.h1, p.N1, p.N1, p.W2, p.b2, p.h2, p.N2, p.N2, p.M, true, s_head_in);
    head_grid_barrier(p.barrier, 3 * G);
    head_layer(p.h2, p.N2, p.N2, p.W3, p.b3, p.y, p.N3, p.N3, p.M, false, s_head_in);
}

}  // namespace

extern "C" int sdn_mlp_head_forward(const float* x, const float* extra, int M, int K0, int N0, int E, int N1, int N2, int N3,
                                    const float* W0, const float* b0, const float* W1, const float* b1, const float* W2,
                                    const float* b2, const float* W3, const float* b3, float* cat, float* h1, float* h2, float* y,
                                    unsigned* barrier, void* stream) {
    if (!x || !W0 || !W1 || !W2 || !W3 || !cat || !h1 || !h2 || !y || !barrier || M <= 0 || K0 <= 0 || N0 <= 0 || E < 0 ||
        N1 <= 0 || N2 <= 0 || N3 <= 0 || (E > 0 && !extra))
        return SDN_E_ARG;
    int kmax = K0;
    if (N0 + E > kmax) kmax = N0 + E;
    if (N1 > kmax) kmax = N1;
    if (N2 > kmax) kmax = N2;
    if (kmax > HEAD_MAXK) return SDN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(barrier, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)HEAD_ROWS * kmax * sizeof(float);
    e = cudaFuncSetAttribute(k_mlp_head, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    HeadParams p = {x, extra, M, K0, N0, E, N1, N2, N3, W0, b0, W1, b1, W2, b2, W3, b3, cat, h1, h2, y, barrier};
    sdn_internal::launch_begin("k_mlp_head", st);
    k_mlp_head<<<HEAD_CTAS, 256, smem, st>>>(p);
    return sdn_internal::launch_end();
}
