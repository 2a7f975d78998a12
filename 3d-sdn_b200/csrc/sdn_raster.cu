// sdn_raster.cu -- sm_100a differentiable mesh rasteriser behind the C ABI of include/sdn_b200.h.
//
// Replaces the seven CuPy elementwise kernels of /root/reference/geometric/neural_renderer/rasterize.py and the
// Chainer glue around them.  Design (DESIGN.md has the long form):
//
//   k_faces_forward   fill_back + camera + perspective + gather      (renderer.py:41, look.py:38-43, perspective.py:13-18,
//                                                                     vertices_to_faces.py:16-21)
//   k_face_setup      cull, barycentric inverse, bbox, active list   (rasterize.py:240-277)
//   k_raster_tiles    persistent CTAs over 64x32-sample tiles: faces are streamed against the tile, coverage is
//                     resolved FACE-parallel into a shared-memory z-buffer of packed (depth_bits<<32 | face) keys with
//                     64-bit atomicMin -- which is exactly the safe path's "strict <, faces in index order" rule
//                     (rasterize.py:301-341) and is order-independent, hence deterministic; then texture sampling
//                     (rasterize.py:362-435), alpha / background (:437-462), row flip + 2x2 mean (:950-966) are
//                     applied in registers and written once, as float2 / int4 rows.
//   k_backward_items  one warp per (visible face, edge, axis) item of the NMR gradient (rasterize.py:514-745) plus one
//                     "interior" item per face that gathers the depth (:791-844) and texture (:747-789) gradients
//                     without atomics; out-sweeps run warp-wide and coalesced.
//   k_grad_assemble   deterministic sum of the per-item partials into grad_faces.
//   k_faces_backward* scatter onto vertices + chain rule through perspective / camera / flip.
//
// Arithmetic contract: everything that decides visibility is IEEE float32 in the reference's op order, compiled with
// -fmad=false (no contraction), IEEE division and sqrt; doubles only where the reference's literals force them.
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "sdn_b200.h"

namespace {

constexpr int TW = 64;        // tile width  in samples
constexpr int TH = 32;        // tile height in samples
constexpr int NT = 256;       // threads per CTA (raster)
constexpr int QCAP = 768;     // per-tile face queue
constexpr unsigned FULL = 0xffffffffu;
typedef unsigned long long u64;

std::atomic<long long> g_launches{0};

// Optional per-kernel timing (sdn_timing_enable): every launch is bracketed by CUDA events on the launching stream.
// Off by default; it is diagnostic state only and never influences results.
std::atomic<int> g_timing{0};
std::mutex g_timing_mu;
struct TimedLaunch { const char* name; cudaEvent_t e0, e1; };
std::vector<TimedLaunch> g_timed;
// the launch being bracketed belongs to the calling thread (DataParallel drives one thread per GPU): thread-local
thread_local const char* g_pending_name = nullptr;
thread_local cudaEvent_t g_pending_e0 = nullptr;
thread_local cudaStream_t g_pending_stream = nullptr;

// SDN_DEBUG_SYNC=1 in the environment: print every launch and synchronise after it (hang / fault localisation)
const bool g_debug_sync = getenv("SDN_DEBUG_SYNC") != nullptr;
thread_local const char* g_dbg_name = nullptr;
thread_local cudaStream_t g_dbg_stream = nullptr;

inline void timing_begin(const char* name, cudaStream_t st) {
    if (g_debug_sync) { g_dbg_name = name; g_dbg_stream = st; fprintf(stderr, "[sdn] launch %s\n", name); fflush(stderr); }
    if (!g_timing.load(std::memory_order_relaxed)) return;
    cudaEvent_t e0;
    if (cudaEventCreate(&e0) != cudaSuccess) return;
    cudaEventRecord(e0, st);
    g_pending_name = name; g_pending_e0 = e0; g_pending_stream = st;
}
inline void timing_end() {
    if (!g_pending_name) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) == cudaSuccess) {
        cudaEventRecord(e1, g_pending_stream);
        std::lock_guard<std::mutex> lk(g_timing_mu);      // the list of finished brackets is the only shared state
        g_timed.push_back({g_pending_name, g_pending_e0, e1});
    }
    g_pending_name = nullptr;
}

// ------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------
struct Query;
struct Workspace {
    int* active_count;      // [B]
    int* vis_count;         // [4]    [0] = number of visible faces in the batch
    int* img_bbox;          // [B,4]  {65535-xmin, 65535-ymin, xmax+1, ymax+1}, all grown with atomicMax from 0
    int* counters;          // [4]    {tile ticket, item ticket, -, -}
    unsigned char* visible; // [B*F]  face won at least one sample
    float* face_inv;        // [B*F,9]
    short* face_bbox;       // [B*F,4] x0,y0,x1,y1 inclusive, clamped to the raster
    int* active_ids;        // [B*Fp]  per image: ids of the faces that can draw (front-facing, on screen)
    short* active_bbox;     // [B*Fp,4] their bboxes, same order (streamed with 1-D bulk copies)
    int* tile_prefix;       // [B+1]   tiles inside each image's face bbox, exclusive prefix
    int* tile_off;          // [tiles+1] faces binned to every tile: counts, then (scan) offsets into bin_ids
    int* tile_cursor;       // [tiles]
    int* bin_ids;           // [bin_cap] face ids grouped by tile
    int* bin_sums;          // scan scratch
    int num_tiles, bin_cap;
    int Fp;                 // F rounded up so that per-image list starts stay 16-byte aligned
    int* vis_ids;           // [B*F]  global face ids b*F+fn of the visible faces
    float* gpart;           // [B*F,12]  K6 partials [edge][axis][2]
    float* gint;            // [B*F,6]   interior sums: S_k (3), colour grad (3)
    // line-sweep backward (fast path)
    int* vis_slot;          // [B*F]     position of the face in vis_ids
    int* scan_buf;          // [lines + 1 + B*F*7 + 1]  counts -> exclusive bases (lines first, then items)
    int* block_sums;        // [scan blocks]
    int* line_cursor;       // [lines]
    int* line_items;        // [qmax]    query indices grouped by line
    struct Query* queries;  // [qmax]
    float2* qres;           // [qmax]    (g0, g1) of each query
    float2* qres2;          // [qmax]    second addend (colour role of the split launch)
    int2* unit_tab;         // [umax]    (visible-face slot, block index) of every 4x32-sample unit of k_bwd_face
    float* upart;           // [umax,18] per-unit partial sums
    int lines, qmax, scan_n, umax;
    size_t zero_bytes;      // leading bytes cleared at the start of every forward
    size_t total;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

Workspace make_ws(void* base, int B, int F, int Hs, int Ws) {
    Workspace w;
    char* p = (char*)base;
    size_t off = 0;
    const size_t BF = (size_t)B * F;
    w.active_count = (int*)(p + off); off += sizeof(int) * B;
    w.img_bbox = (int*)(p + off); off += sizeof(int) * 4 * B;
    w.vis_count = (int*)(p + off); off += sizeof(int) * 4;      // [0] = visible faces of the whole batch
    w.counters = (int*)(p + off); off += sizeof(int) * 4;       // contiguous with vis_count (one memset)
    w.visible = (unsigned char*)(p + off); off += BF;
    off = align_up(off, 256);
    w.num_tiles = B * ((Ws + TW - 1) / TW) * ((Hs + TH - 1) / TH);
    w.tile_off = (int*)(p + off); off = align_up(off + (size_t)(w.num_tiles + 1) * 4, 256);
    w.tile_cursor = (int*)(p + off); off = align_up(off + (size_t)w.num_tiles * 4, 256);
    w.zero_bytes = off;
    w.face_inv = (float*)(p + off); off = align_up(off + BF * 36, 256);
    w.face_bbox = (short*)(p + off); off = align_up(off + BF * 8, 256);
    w.Fp = (F + 3) / 4 * 4 + 16;
    w.active_ids = (int*)(p + off); off = align_up(off + (size_t)B * w.Fp * 4, 256);
    w.active_bbox = (short*)(p + off); off = align_up(off + (size_t)B * w.Fp * 8, 256);
    w.tile_prefix = (int*)(p + off); off = align_up(off + (size_t)(B + 1) * 4, 256);
    const size_t bc = BF * 3 + 262144 + (size_t)w.num_tiles * 3;       // + padding of every tile's slice to 4 ids
    w.bin_cap = (int)(bc > 64000000 ? 64000000 : bc);
    w.bin_ids = (int*)(p + off); off = align_up(off + (size_t)w.bin_cap * 4, 256);
    w.bin_sums = (int*)(p + off); off = align_up(off + ((size_t)w.num_tiles / 2048 + 2) * 4, 256);
    w.vis_ids = (int*)(p + off); off = align_up(off + BF * 4, 256);
    w.gpart = (float*)(p + off); off = align_up(off + BF * 48, 256);
    w.gint = (float*)(p + off); off = align_up(off + BF * 24, 256);
    w.lines = B * (Hs + Ws);
    const size_t qm = BF * 24;
    w.qmax = (int)(qm < 65536 ? 65536 : (qm > 48000000 ? 48000000 : qm));
    w.scan_n = w.lines + 1 + (int)BF * 7 + 1 + (int)BF + 1;
    const size_t um = BF * 2 + 65536;
    w.umax = (int)(um > 8000000 ? 8000000 : um);
    w.vis_slot = (int*)(p + off); off = align_up(off + BF * 4, 256);
    w.scan_buf = (int*)(p + off); off = align_up(off + (size_t)w.scan_n * 4, 256);
    w.block_sums = (int*)(p + off); off = align_up(off + ((size_t)w.scan_n / 2048 + 2) * 4, 256);
    w.line_cursor = (int*)(p + off); off = align_up(off + (size_t)w.lines * 4, 256);
    w.line_items = (int*)(p + off); off = align_up(off + (size_t)w.qmax * 4, 256);
    w.queries = (struct Query*)(p + off); off = align_up(off + (size_t)w.qmax * 32, 256);
    w.qres = (float2*)(p + off); off = align_up(off + (size_t)w.qmax * 8, 256);
    w.qres2 = (float2*)(p + off); off = align_up(off + (size_t)w.qmax * 8, 256);
    w.unit_tab = (int2*)(p + off); off = align_up(off + (size_t)w.umax * 8, 256);
    w.upart = (float*)(p + off); off = align_up(off + (size_t)w.umax * 72, 256);
    w.total = off;
    return w;
}

// ------------------------------------------------------------------------------------------------------------
// shared device arithmetic (op order of rasterize.py; -fmad=false keeps it uncontracted)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_pixel(float v, int is) {     // rasterize.py:258
    float t = v * (float)is;
    t = t + (float)is;
    t = t - 1.0f;
    return 0.5f * t;
}
__device__ __forceinline__ bool is_backside(const float* f) {    // rasterize.py:252
    return (f[7] - f[1]) * (f[3] - f[0]) < (f[4] - f[1]) * (f[6] - f[0]);
}

struct Clip {
    double near_, far_;
    float far_f;
};

// One (face, sample) evaluation of the safe kernel, rasterize.py:310-332.  Returns false when rejected.
__device__ __forceinline__ bool eval_sample(const float* f, const float* inv, float xp, float yp, float xif,
                                            float yif, const Clip& c, float* w, float& zp) {
    if (((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) ||
        ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) ||
        ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7]))) return false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float t = inv[3 * k + 0] * xif;
        float u = inv[3 * k + 1] * yif;
        t = t + u;
        w[k] = fminf(fmaxf(t + inv[3 * k + 2], 0.0f), 1.0f);
    }
    float ws = w[0] + w[1];
    ws = ws + w[2];
    w[0] = __fdiv_rn(w[0], ws);
    w[1] = __fdiv_rn(w[1], ws);
    w[2] = __fdiv_rn(w[2], ws);
    float s = __fdiv_rn(w[0], f[2]) + __fdiv_rn(w[1], f[5]);
    s = s + __fdiv_rn(w[2], f[8]);
    zp = __fdiv_rn(1.0f, s);
    if ((double)zp <= c.near_ || c.far_ <= (double)zp) return false;
    return zp < c.far_f;        // depth_min starts at (float)far, rasterize.py:297; also rejects NaN
}


// Same arithmetic as eval_sample, split so that two samples can be evaluated with interleaved (independent) division
// chains: `inside_test` is the three half-plane tests, `depth_of` the weights / depth / clip tests.
__device__ __forceinline__ bool inside_test(const float* f, float xp, float yp) {
    return !(((yp - f[1]) * (f[3] - f[0]) < (xp - f[0]) * (f[4] - f[1])) ||
             ((yp - f[4]) * (f[6] - f[3]) < (xp - f[3]) * (f[7] - f[4])) ||
             ((yp - f[7]) * (f[0] - f[6]) < (xp - f[6]) * (f[1] - f[7])));
}
__device__ __forceinline__ bool depth_of(const float* f, const float* inv, float xif, float yif, const Clip& c,
                                         float& zp) {
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float t = inv[3 * k + 0] * xif;
        float u = inv[3 * k + 1] * yif;
        t = t + u;
        w[k] = fminf(fmaxf(t + inv[3 * k + 2], 0.0f), 1.0f);
    }
    float ws = w[0] + w[1];
    ws = ws + w[2];
    w[0] = __fdiv_rn(w[0], ws);
    w[1] = __fdiv_rn(w[1], ws);
    w[2] = __fdiv_rn(w[2], ws);
    float s = __fdiv_rn(w[0], f[2]) + __fdiv_rn(w[1], f[5]);
    s = s + __fdiv_rn(w[2], f[8]);
    zp = __fdiv_rn(1.0f, s);
    if ((double)zp <= c.near_ || c.far_ <= (double)zp) return false;
    return zp < c.far_f;
}

// rasterize.py:399-423: the 8 trilinear corners of one covered sample.  ti[k] float index along texture axis k.
__device__ __forceinline__ void texture_index(const float* w, float depth, const float* f, int ts, double eps,
                                              float* tif) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float q = __fdiv_rn(depth, f[3 * k + 2]);
        tif[k] = (float)(((double)w[k] * ((double)(ts - 1) - eps)) * (double)q);
    }
}
__device__ __forceinline__ float corner_weight(const float* tif, int pn, int* tii) {
    float w = 1.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int ti = __float2int_rz(tif[k]);
        const float frac = tif[k] - (float)ti;
        if (((pn >> k) & 1) == 0) { w = w * (1.0f - frac); tii[k] = ti; }
        else { w = w * frac; tii[k] = ti + 1; }
    }
    return w;
}

__device__ __forceinline__ float xp_of(int xi, int Ws) { return __fdiv_rn((float)(2 * xi + 1 - Ws), (float)Ws); }

// ------------------------------------------------------------------------------------------------------------
// faces forward: fill_back + flip + camera + perspective + gather (+ normals)
// ------------------------------------------------------------------------------------------------------------
struct CamDev {
    const float* rot; const float* eye; const float* wx; const float* wy; float flip;
};

__device__ __forceinline__ void cam_vertex(const float* v, int b, const CamDev& c, float* pre, float* out) {
    float x = v[0] * c.flip, y = v[1], z = v[2];
    pre[0] = x; pre[1] = y; pre[2] = z;
    if (c.eye) { x = x - c.eye[b * 3 + 0]; y = y - c.eye[b * 3 + 1]; z = z - c.eye[b * 3 + 2]; }
    if (c.rot) {
        const float* r = c.rot + b * 9;
        float o[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            float s = x * r[3 * i + 0] + y * r[3 * i + 1];
            o[i] = s + z * r[3 * i + 2];
        }
        x = o[0]; y = o[1]; z = o[2];
    }
    if (c.wx) {
        x = __fdiv_rn(__fdiv_rn(x, z), c.wx[b]);
        y = __fdiv_rn(__fdiv_rn(y, z), c.wy[b]);
    }
    out[0] = x; out[1] = y; out[2] = z;
}

__global__ void k_faces_forward(const float* __restrict__ vertices, const int32_t* __restrict__ faces, int B, int nv,
                                int nf, int faces_batched, int fill_back, CamDev cam, float* __restrict__ face_ndc,
                                float* __restrict__ face_normal) {
    const int F = fill_back ? 2 * nf : nf;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * F) return;
    const int b = (int)(i / F), fn = (int)(i % F);
    const int src = fn < nf ? fn : fn - nf;
    const int32_t* tri = faces + ((size_t)(faces_batched ? b : 0) * nf + src) * 3;
    int vid[3] = {tri[0], tri[1], tri[2]};
    if (fn >= nf) { int t = vid[0]; vid[0] = vid[2]; vid[2] = t; }      // faces[:, :, ::-1]
    float pre[3][3], ndc[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++) cam_vertex(vertices + ((size_t)b * nv + vid[k]) * 3, b, cam, pre[k], ndc[k]);
    float* o = face_ndc + (size_t)i * 9;
#pragma unroll
    for (int k = 0; k < 3; k++) { o[3 * k] = ndc[k][0]; o[3 * k + 1] = ndc[k][1]; o[3 * k + 2] = ndc[k][2]; }
    if (face_normal) {      // derender3d/models/renderer.py:66-73
        const float a0 = pre[0][0] - pre[1][0], a1 = pre[0][1] - pre[1][1], a2 = pre[0][2] - pre[1][2];
        const float b0 = pre[2][0] - pre[1][0], b1 = pre[2][1] - pre[1][1], b2 = pre[2][2] - pre[1][2];
        const float c0 = a1 * b2 - a2 * b1, c1 = a2 * b0 - a0 * b2, c2 = a0 * b1 - a1 * b0;
        float n2 = c0 * c0 + c1 * c1;
        n2 = n2 + c2 * c2;
        const float norm = __fsqrt_rn(n2) + 1e-5f;
        float* n = face_normal + (size_t)i * 3;
        // x is reported in the caller's (un-flipped) frame: derender3d/models/renderer.py:268-270 negates it again
        n[0] = __fdiv_rn(c0, norm) * cam.flip; n[1] = __fdiv_rn(c1, norm); n[2] = __fdiv_rn(c2, norm);
    }
}

// ------------------------------------------------------------------------------------------------------------
// face setup (K1a) + active list + image bbox
// ------------------------------------------------------------------------------------------------------------
__global__ void k_face_setup(const float* __restrict__ faces, int B, int F, int Hs, int Ws, Workspace ws) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i < (long)B * F;
    const int b = in_range ? (int)(i / F) : 0;
    const int fn = in_range ? (int)(i % F) : 0;
    bool active = false;
    int x0 = 0, y0 = 0, x1 = -1, y1 = -1;
    if (in_range) {
        float f[9];
#pragma unroll
        for (int k = 0; k < 9; k++) f[k] = faces[(size_t)i * 9 + k];
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 9; k++) finite = finite && (fabsf(f[k]) <= 3.0e38f);
        float inv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (finite && !is_backside(f)) {
            float p[3][2];
#pragma unroll
            for (int n = 0; n < 3; n++) { p[n][0] = to_pixel(f[3 * n], Ws); p[n][1] = to_pixel(f[3 * n + 1], Hs); }
            const float fi[9] = {
                p[1][1] - p[2][1], p[2][0] - p[1][0], p[1][0] * p[2][1] - p[2][0] * p[1][1],
                p[2][1] - p[0][1], p[0][0] - p[2][0], p[2][0] * p[0][1] - p[0][0] * p[2][1],
                p[0][1] - p[1][1], p[1][0] - p[0][0], p[0][0] * p[1][1] - p[1][0] * p[0][1]};
            const float den = (p[2][0] * (p[0][1] - p[1][1]) + p[0][0] * (p[1][1] - p[2][1])) +
                              p[1][0] * (p[2][1] - p[0][1]);
#pragma unroll
            for (int k = 0; k < 9; k++) inv[k] = __fdiv_rn(fi[k], den);
            // a zero / non-finite denominator can never produce a finite depth (w -> NaN): never drawn.
            if (den != 0.0f && fabsf(den) <= 3.0e38f) {
                const float mxa = fmaxf(fmaxf(fabsf(f[0]), fabsf(f[3])), fabsf(f[6]));
                const float mya = fmaxf(fmaxf(fabsf(f[1]), fabsf(f[4])), fabsf(f[7]));
                if (mxa > 64.0f || mya > 64.0f) {       // rounding error may exceed a sample: no bbox culling
                    x0 = 0; y0 = 0; x1 = Ws - 1; y1 = Hs - 1;
                } else {
                    // Samples sit at integer pixel coordinates.  One that passes the float half-plane tests lies inside
                    // the triangle up to the tests' rounding error (~1e-7 relative in NDC, i.e. < 1e-3 pixel here), so
                    // the integer range of [pmin - m, pmax + m] with m = 0.01 px + 8 ulp is conservative -- and 3-9x
                    // tighter than a whole-sample margin for the 1-2 pixel faces of ShapeNet-sized meshes.
                    const float pxmin = fminf(fminf(p[0][0], p[1][0]), p[2][0]);
                    const float pxmax = fmaxf(fmaxf(p[0][0], p[1][0]), p[2][0]);
                    const float pymin = fminf(fminf(p[0][1], p[1][1]), p[2][1]);
                    const float pymax = fmaxf(fmaxf(p[0][1], p[1][1]), p[2][1]);
                    const float mx = 0.01f + 1e-6f * fmaxf(fabsf(pxmin), fabsf(pxmax));
                    const float my = 0.01f + 1e-6f * fmaxf(fabsf(pymin), fabsf(pymax));
                    x0 = max(__float2int_ru(pxmin - mx), 0);
                    y0 = max(__float2int_ru(pymin - my), 0);
                    x1 = min(__float2int_rd(pxmax + mx), Ws - 1);
                    y1 = min(__float2int_rd(pymax + my), Hs - 1);
                }
                active = (x0 <= x1) && (y0 <= y1);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; k++) ws.face_inv[(size_t)i * 9 + k] = inv[k];
        short* bb = ws.face_bbox + (size_t)i * 4;
        if (active) { bb[0] = (short)x0; bb[1] = (short)y0; bb[2] = (short)x1; bb[3] = (short)y1; }
        else { bb[0] = 1; bb[1] = 1; bb[2] = 0; bb[3] = 0; }
    }
    // warp-aggregated append + bbox growth.  A warp may straddle two images: handle per distinct b.
    const unsigned lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(FULL, active);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lb = __shfl_sync(FULL, b, leader);
        const unsigned same = __ballot_sync(FULL, active && b == lb);
        if (active && b == lb) {
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(&ws.active_count[lb], __popc(same));
            base = __shfl_sync(same, base, leader);
            const size_t slot = (size_t)lb * ws.Fp + base + __popc(same & ((1u << lane) - 1));
            ws.active_ids[slot] = fn;
            *reinterpret_cast<short4*>(ws.active_bbox + slot * 4) = make_short4((short)x0, (short)y0, (short)x1, (short)y1);
            int a0 = 65535 - x0, a1 = 65535 - y0, a2 = x1 + 1, a3 = y1 + 1;
            a0 = __reduce_max_sync(same, a0); a1 = __reduce_max_sync(same, a1);
            a2 = __reduce_max_sync(same, a2); a3 = __reduce_max_sync(same, a3);
            if ((int)lane == leader) {
                atomicMax(&ws.img_bbox[lb * 4 + 0], a0); atomicMax(&ws.img_bbox[lb * 4 + 1], a1);
                atomicMax(&ws.img_bbox[lb * 4 + 2], a2); atomicMax(&ws.img_bbox[lb * 4 + 3], a3);
            }
        }
        todo &= ~same;
    }
}

// ------------------------------------------------------------------------------------------------------------
// tiled forward
// ------------------------------------------------------------------------------------------------------------
struct FwdParams {
    int B, F, H, W, ss, Hs, Ws, flags, ts, batch_bug;
    Clip clip;
    double eps;
    float bg[3];
    const float* bg_b;
    const float* faces;
    const float* textures;
    float* rgb; float* alpha; float* depth; int32_t* face_index; float* rgb_ss;
    int tiles_x, tiles_y, num_tiles;
};

__device__ __forceinline__ void load_bg(const FwdParams& P, int b, float* bg) {
    if (P.bg_b) { bg[0] = P.bg_b[b * 3]; bg[1] = P.bg_b[b * 3 + 1]; bg[2] = P.bg_b[b * 3 + 2]; }
    else { bg[0] = P.bg[0]; bg[1] = P.bg[1]; bg[2] = P.bg[2]; }
}

// colour of a covered sample (rasterize.py:378-423).  The barycentric weights are recomputed with the visibility
// arithmetic, so they equal the reference's weight_map entry bit for bit.
__device__ __forceinline__ void shade_sample(const FwdParams& P, const float* __restrict__ face_inv, int b, int fn,
                                             int xi, int yi, float depth, float* rgb) {
    const float* f = P.faces + ((size_t)(P.batch_bug ? 0 : b) * P.F + fn) * 9;   // rasterize.py:390
    const float* ff = P.faces + ((size_t)b * P.F + fn) * 9;
    const float* iv = face_inv + ((size_t)b * P.F + fn) * 9;
    float fr[9], inv[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { fr[k] = __ldg(ff + k); inv[k] = __ldg(iv + k); }
    float w[3];
    const float xif = (float)xi, yif = (float)yi;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float t = inv[3 * k + 0] * xif;
        float u = inv[3 * k + 1] * yif;
        t = t + u;
        w[k] = fminf(fmaxf(t + inv[3 * k + 2], 0.0f), 1.0f);
    }
    float wsum = w[0] + w[1];
    wsum = wsum + w[2];
    w[0] = __fdiv_rn(w[0], wsum); w[1] = __fdiv_rn(w[1], wsum); w[2] = __fdiv_rn(w[2], wsum);

    const int ts = P.ts == 1 ? 2 : P.ts;
    float zf[9];
    if (P.batch_bug) {
#pragma unroll
        for (int k = 0; k < 9; k++) zf[k] = __ldg(f + k);
    } else {
#pragma unroll
        for (int k = 0; k < 9; k++) zf[k] = fr[k];
    }
    float tif[3];
    texture_index(w, depth, zf, ts, P.eps, tif);
    float acc[3] = {0.0f, 0.0f, 0.0f};
    if (P.ts == 1) {
        const float* c = P.textures + ((size_t)b * P.F + fn) * 3;
        const float c0 = __ldg(c), c1 = __ldg(c + 1), c2 = __ldg(c + 2);
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            int tii[3];
            const float cw = corner_weight(tif, pn, tii);
            acc[0] = acc[0] + cw * c0; acc[1] = acc[1] + cw * c1; acc[2] = acc[2] + cw * c2;
        }
    } else {
        const float* tex = P.textures + ((size_t)b * P.F + fn) * ts * ts * ts * 3;
#pragma unroll
        for (int pn = 0; pn < 8; pn++) {
            int tii[3];
            const float cw = corner_weight(tif, pn, tii);
            const int isc = tii[0] * ts * ts + tii[1] * ts + tii[2];
            acc[0] = acc[0] + cw * __ldg(tex + isc * 3 + 0);
            acc[1] = acc[1] + cw * __ldg(tex + isc * 3 + 1);
            acc[2] = acc[2] + cw * __ldg(tex + isc * 3 + 2);
        }
    }
    rgb[0] = acc[0]; rgb[1] = acc[1]; rgb[2] = acc[2];
}

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* s_warp, int& total);

// ---- background: everything is first filled with the "no face" values by a pure streaming kernel (16-byte stores);
// the tile kernel below then only visits tiles that can contain a face.
__device__ __forceinline__ void fill_words(uint32_t* __restrict__ dst, size_t n, uint32_t v, size_t tid, size_t nthreads) {
    const size_t n4 = n / 4;
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const uint4 v4 = make_uint4(v, v, v, v);
    for (size_t i = tid; i < n4; i += nthreads) d4[i] = v4;
    for (size_t i = n4 * 4 + tid; i < n; i += nthreads) dst[i] = v;
}

__global__ void __launch_bounds__(256) k_fill_background(FwdParams P) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    const size_t P1 = (size_t)P.H * P.W;
    fill_words(reinterpret_cast<uint32_t*>(P.face_index), (size_t)P.B * P.Hs * P.Ws, 0xffffffffu, tid, nth);
    if (P.flags & SDN_ALPHA) fill_words(reinterpret_cast<uint32_t*>(P.alpha), P.B * P1, 0u, tid, nth);
    if (P.flags & SDN_DEPTH) fill_words(reinterpret_cast<uint32_t*>(P.depth), P.B * P1, __float_as_uint(P.clip.far_f), tid, nth);
    if (P.flags & SDN_RGB) {
        const bool uniform = !P.bg_b && P.bg[0] == P.bg[1] && P.bg[1] == P.bg[2];
        if (uniform) {
            fill_words(reinterpret_cast<uint32_t*>(P.rgb), P.B * 3 * P1, __float_as_uint(P.bg[0]), tid, nth);
            if (P.rgb_ss) fill_words(reinterpret_cast<uint32_t*>(P.rgb_ss), (size_t)P.B * P.Hs * P.Ws * 3,
                                     __float_as_uint(P.bg[0]), tid, nth);
        } else {
            for (size_t i = tid; i < P.B * 3 * P1; i += nth) {
                const int bimg = (int)(i / (3 * P1)), c = (int)((i / P1) % 3);
                P.rgb[i] = P.bg_b ? P.bg_b[bimg * 3 + c] : P.bg[c];
            }
            if (P.rgb_ss) {
                const size_t per = (size_t)P.Hs * P.Ws * 3;
                for (size_t i = tid; i < P.B * per; i += nth) {
                    const int bimg = (int)(i / per), c = (int)(i % 3);
                    P.rgb_ss[i] = P.bg_b ? P.bg_b[bimg * 3 + c] : P.bg[c];
                }
            }
        }
    }
}

// faces -> tiles: count (FILL = false), exclusive scan, fill (FILL = true).  One thread per active face.
template <bool FILL>
__global__ void __launch_bounds__(256) k_bin_faces(int B, int tiles_x, int tiles_y, Workspace ws) {
    const int b = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ws.active_count[b]) return;
    if (FILL && ws.tile_off[ws.num_tiles] > ws.bin_cap) return;        // does not fit: tiles scan the lists instead
    const size_t slot = (size_t)b * ws.Fp + e;
    const short4 bb = *reinterpret_cast<const short4*>(ws.active_bbox + slot * 4);
    const int id = FILL ? ws.active_ids[slot] : 0;
    for (int ty = bb.y / TH; ty <= bb.w / TH; ty++)
        for (int tx = bb.x / TW; tx <= bb.z / TW; tx++) {
            const int t = (b * tiles_y + ty) * tiles_x + tx;
            if (!FILL) atomicAdd(&ws.tile_off[t], 1);
            else ws.bin_ids[ws.tile_off[t] + atomicAdd(&ws.tile_cursor[t], 1)] = id;
        }
}

// Rounds every tile's face count up to a multiple of 4 before the scan: each tile's slice of bin_ids then starts on a
// 16-byte boundary and can be staged into shared memory with ONE 1-D TMA bulk copy (k_raster_tiles).  The real counts
// are what k_bin_faces<true> leaves in tile_cursor.
__global__ void __launch_bounds__(256) k_pad_tile_counts(Workspace ws) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ws.num_tiles) ws.tile_off[t] = (ws.tile_off[t] + 3) & ~3;
}

// tiles that intersect each image's face bounding box: prefix[b] .. prefix[b+1]
__device__ __forceinline__ void image_tile_range(const Workspace& ws, int b, int& tx0, int& ty0, int& nx, int& ny) {
    nx = 0; ny = 0; tx0 = 0; ty0 = 0;
    if (ws.active_count[b] == 0) return;
    const int ixmin = 65535 - ws.img_bbox[b * 4 + 0], iymin = 65535 - ws.img_bbox[b * 4 + 1];
    const int ixmax = ws.img_bbox[b * 4 + 2] - 1, iymax = ws.img_bbox[b * 4 + 3] - 1;
    tx0 = ixmin / TW; ty0 = iymin / TH;
    nx = ixmax / TW - tx0 + 1; ny = iymax / TH - ty0 + 1;
}

__global__ void __launch_bounds__(256) k_tile_prefix(int B, Workspace ws) {
    __shared__ int s_warp[8];
    int carry = 0;
    for (int base = 0; base < B; base += 256) {
        const int b = base + threadIdx.x;
        int n = 0;
        if (b < B) { int tx0, ty0, nx, ny; image_tile_range(ws, b, tx0, ty0, nx, ny); n = nx * ny; }
        int total;
        const int ex = block_exclusive_scan_256(n, s_warp, total);
        if (b < B) ws.tile_prefix[b] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) ws.tile_prefix[B] = carry;
}

// ---- 1-D bulk copies (TMA) + mbarrier: the per-image active-face lists are streamed into shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Early depth rejection in k_raster_tiles.  depth_of's result is 1 / sum_k (w_k / z_k) with w clamped to [0, 1] and
// normalised: a convex combination of the vertex depths in 1/z, so it cannot be smaller than the smallest vertex depth
// except by rounding (seven correctly rounded operations: a few ulp).  A face whose smallest vertex depth, less a 1e-5
// relative margin, is already greater than the depth stored for a sample cannot win that sample's atomicMin, and the
// seven IEEE divisions are skipped.  Depths that reach the z-buffer are positive, so their bit patterns order like the
// values; a face with a non-positive vertex depth gets bound 0 (never rejected).  The stored depth only decreases, so a
// stale read is conservative.  The winner of every sample, hence the face-index map, is unchanged.
__device__ __forceinline__ unsigned z_lower_bits(const float* f) {
    const float zmin = fminf(f[2], fminf(f[5], f[8]));
    return zmin > 0.0f ? __float_as_uint(zmin * 0.99999f) : 0u;
}
__device__ __forceinline__ unsigned key_depth_bits(const u64* key) {
    return reinterpret_cast<const volatile unsigned*>(key)[1];      // high word of (depth bits << 32 | face)
}

constexpr int BCH = 1024;       // active-list entries per pipeline stage

template <int SS>
__global__ void __launch_bounds__(NT, 4) k_raster_tiles(FwdParams P, Workspace ws) {
    __shared__ u64 s_key[TH * TW];
    __shared__ __align__(128) short4 s_bb[2][BCH];
    __shared__ __align__(128) int s_ids[2][BCH];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ __align__(16) int s_queue[QCAP];
    __shared__ int s_big[QCAP];
    __shared__ float s_xp[TW], s_yp[TH];
    __shared__ int s_qn, s_tile, s_nbig;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float far_f = P.clip.far_f;
    const int total_tiles = ws.tile_prefix[P.B];
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase[2] = {0u, 0u};

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(&ws.counters[0], 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= total_tiles) break;
        // image of this ticket: largest b with prefix[b] <= tile
        int lo = 0, hi = P.B - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ws.tile_prefix[mid] <= tile) lo = mid; else hi = mid - 1; }
        const int b = lo;
        int tx0, ty0, nx, ny;
        image_tile_range(ws, b, tx0, ty0, nx, ny);
        const int trem = tile - ws.tile_prefix[b];
        const int ty = ty0 + trem / nx, tx = tx0 + trem % nx;
        const int sx0 = tx * TW, sy0 = ty * TH;
        const int sx1 = min(sx0 + TW, P.Ws) - 1, sy1 = min(sy0 + TH, P.Hs) - 1;
        const int count = ws.active_count[b];

        // ---- face-parallel coverage of the queued faces: one warp per face, lanes over 8x4-sample blocks, two blocks
        //      per iteration so that their division chains are independent (ILP)
        auto rasterise_queue = [&](const int qn) {
            // (1) small faces (<= 16 samples of this tile): one LANE per face -- ShapeNet-sized meshes are almost
            //     entirely such faces and a whole warp per face would idle 3/4 of its lanes on per-face overhead
            if (tid == 0) s_nbig = 0;
            __syncthreads();
            for (int q = tid; q < qn; q += NT) {
                const int fn = s_queue[q];
                const size_t fo = ((size_t)b * P.F + fn);
                const short4 bb = *reinterpret_cast<const short4*>(ws.face_bbox + fo * 4);
                const int bx0 = max((int)bb.x, sx0), bx1 = min((int)bb.z, sx1);
                const int by0 = max((int)bb.y, sy0), by1 = min((int)bb.w, sy1);
                if ((bx1 - bx0 + 1) * (by1 - by0 + 1) > 16) { s_big[atomicAdd(&s_nbig, 1)] = fn; continue; }
                float f[9], inv[9];
#pragma unroll
                for (int k = 0; k < 9; k++) { f[k] = __ldg(P.faces + fo * 9 + k); inv[k] = __ldg(ws.face_inv + fo * 9 + k); }
                const unsigned zlo = z_lower_bits(f);
                for (int yy = by0; yy <= by1; yy++)
                    for (int xx = bx0; xx <= bx1; xx++) {
                        if (!inside_test(f, s_xp[xx - sx0], s_yp[yy - sy0])) continue;
                        if (zlo > key_depth_bits(&s_key[(yy - sy0) * TW + (xx - sx0)])) continue;      // cannot win
                        float z;
                        if (depth_of(f, inv, (float)xx, (float)yy, P.clip, z))
                            atomicMin(&s_key[(yy - sy0) * TW + (xx - sx0)], ((u64)__float_as_uint(z) << 32) | (unsigned)fn);
                    }
            }
            __syncthreads();
            // (2) larger faces: one WARP per face, lanes over 8x4-sample blocks, two blocks per iteration so that
            //     their division chains are independent (ILP)
            const int nbig = s_nbig;
            for (int q = warp; q < nbig; q += NT / 32) {
                const int fn = s_big[q];
                const size_t fo = ((size_t)b * P.F + fn);
                float f[9], inv[9];
#pragma unroll
                for (int k = 0; k < 9; k++) { f[k] = __ldg(P.faces + fo * 9 + k); inv[k] = __ldg(ws.face_inv + fo * 9 + k); }
                const short4 bb = *reinterpret_cast<const short4*>(ws.face_bbox + fo * 4);
                const int bx0 = max((int)bb.x, sx0), bx1 = min((int)bb.z, sx1);
                const int by0 = max((int)bb.y, sy0), by1 = min((int)bb.w, sy1);
                const int nbx = (bx1 - bx0) / 8 + 1, nby = (by1 - by0) / 4 + 1, nblk = nbx * nby;
                const unsigned zlo = z_lower_bits(f);
                int cxa = 0, cya = 0;                      // block (column, row) of n0; n0 + 1 is its successor in row order
                for (int n0 = 0; n0 < nblk; n0 += 2) {
                    int cxb = cxa + 1, cyb = cya;
                    if (cxb == nbx) { cxb = 0; cyb++; }
                    int xa = bx0 + cxa * 8 + (lane & 7), ya = by0 + cya * 4 + (lane >> 3);
                    int xb = bx0 + cxb * 8 + (lane & 7), yb = by0 + cyb * 4 + (lane >> 3);
                    const bool va = xa <= bx1 && ya <= by1, vb = n0 + 1 < nblk && xb <= bx1 && yb <= by1;
                    cxa = cxb + 1; cya = cyb;
                    if (cxa == nbx) { cxa = 0; cya++; }
                    xa = min(xa, sx1); ya = min(ya, sy1); xb = min(xb, sx1); yb = min(yb, sy1);
                    // inside the triangle AND not already behind the sample's current depth (the interpolated depth is a
                    // convex combination of the vertex depths: its minimum, less a rounding margin, bounds it from below)
                    const bool ia = va && inside_test(f, s_xp[xa - sx0], s_yp[ya - sy0]) &&
                                    zlo <= key_depth_bits(&s_key[(ya - sy0) * TW + (xa - sx0)]);
                    const bool ib = vb && inside_test(f, s_xp[xb - sx0], s_yp[yb - sy0]) &&
                                    zlo <= key_depth_bits(&s_key[(yb - sy0) * TW + (xb - sx0)]);
                    if (!__any_sync(FULL, ia || ib)) continue;
                    float za, zb;
                    const bool oa = depth_of(f, inv, (float)xa, (float)ya, P.clip, za);
                    const bool ob = depth_of(f, inv, (float)xb, (float)yb, P.clip, zb);
                    if (ia && oa) atomicMin(&s_key[(ya - sy0) * TW + (xa - sx0)], ((u64)__float_as_uint(za) << 32) | (unsigned)fn);
                    if (ib && ob) atomicMin(&s_key[(yb - sy0) * TW + (xb - sx0)], ((u64)__float_as_uint(zb) << 32) | (unsigned)fn);
                }
            }
        };

        const bool binned = ws.tile_off[ws.num_tiles] <= ws.bin_cap;
        if (binned) {
            // the faces of this tile were binned by k_bin_faces: every entry is a hit.  The tile's slice of bin_ids
            // (16-byte aligned, see k_pad_tile_counts) is staged into shared memory in chunks of QCAP ids, each ONE
            // 1-D TMA bulk copy (cp.async.bulk + mbarrier); the first one flies while the z-buffer is initialised.
            const int tg = (b * P.tiles_y + ty) * P.tiles_x + tx;
            const int off = ws.tile_off[tg], cnt = ws.tile_cursor[tg];
            if (tid == 0 && cnt > 0) {
                const uint32_t bytes = (uint32_t)(((min(QCAP, cnt) + 3) / 4) * 16);
                mbar_expect_tx(&s_bar[0], bytes);
                bulk_g2s(&s_queue[0], ws.bin_ids + off, bytes, &s_bar[0]);
            }
            for (int k = tid; k < TH * TW; k += NT) s_key[k] = ~0ull;
            if (tid < TW) s_xp[tid] = xp_of(sx0 + tid, P.Ws);
            if (tid < TH) s_yp[tid] = xp_of(sy0 + tid, P.Hs);
            for (int base = 0; base < cnt; base += QCAP) {
                const int n = min(QCAP, cnt - base);
                if (base > 0) {
                    __syncthreads();                    // the previous chunk has been consumed
                    if (tid == 0) {
                        const uint32_t bytes = (uint32_t)(((n + 3) / 4) * 16);
                        mbar_expect_tx(&s_bar[0], bytes);
                        bulk_g2s(&s_queue[0], ws.bin_ids + off + base, bytes, &s_bar[0]);
                    }
                }
                mbar_wait(&s_bar[0], phase[0]);
                phase[0] ^= 1u;
                __syncthreads();                        // + the shared-memory initialisation above
                rasterise_queue(n);
            }
        } else {
        // ---- overflow of the bins: stream the image's whole active list (ids + bboxes) through shared memory with
        //      1-D TMA bulk copies and test every bbox against the tile
        const int* ids_g = ws.active_ids + (size_t)b * ws.Fp;
        const short4* bb_g = reinterpret_cast<const short4*>(ws.active_bbox) + (size_t)b * ws.Fp;
        const int nchunks = (count + BCH - 1) / BCH;
        int qtotal = 0;                                 // faces queued so far (uniform across the CTA)
        // stage 0 of the pipeline goes out first; the shared-memory initialisation below hides its latency
        if (tid == 0 && nchunks > 0) {
            const int n = min(BCH, count);
            const uint32_t by_i = ((n + 3) / 4) * 16, by_b = ((n + 1) / 2) * 16;
            mbar_expect_tx(&s_bar[0], by_i + by_b);
            bulk_g2s(&s_ids[0][0], ids_g, by_i, &s_bar[0]);
            bulk_g2s(&s_bb[0][0], bb_g, by_b, &s_bar[0]);
        }
        for (int k = tid; k < TH * TW; k += NT) s_key[k] = ~0ull;
        if (tid < TW) s_xp[tid] = xp_of(sx0 + tid, P.Ws);
        if (tid < TH) s_yp[tid] = xp_of(sy0 + tid, P.Hs);
        if (tid == 0) s_qn = 0;
        __syncthreads();
        for (int c = 0; c < nchunks; c++) {
            const int st = c & 1;
            if (tid == 0 && c + 1 < nchunks) {      // prefetch the next chunk into the other stage (already drained)
                const int n = min(BCH, count - (c + 1) * BCH);
                const uint32_t by_i = ((n + 3) / 4) * 16, by_b = ((n + 1) / 2) * 16;
                mbar_expect_tx(&s_bar[st ^ 1], by_i + by_b);
                bulk_g2s(&s_ids[st ^ 1][0], ids_g + (size_t)(c + 1) * BCH, by_i, &s_bar[st ^ 1]);
                bulk_g2s(&s_bb[st ^ 1][0], bb_g + (size_t)(c + 1) * BCH, by_b, &s_bar[st ^ 1]);
            }
            mbar_wait(&s_bar[st], phase[st]);
            phase[st] ^= 1u;
            const int nin = min(BCH, count - c * BCH);
            for (int base = 0; base < nin; base += NT) {
                const int e = base + tid;
                bool hit = false;
                int id = -1;
                if (e < nin) {
                    const short4 bb = s_bb[st][e];
                    hit = !(bb.x > sx1 || bb.z < sx0 || bb.y > sy1 || bb.w < sy0);
                    id = s_ids[st][e];
                }
                const unsigned m = __ballot_sync(FULL, hit);
                int wb = 0;
                if (lane == 0 && m) wb = atomicAdd(&s_qn, __popc(m));
                wb = __shfl_sync(FULL, wb, 0);
                if (hit) s_queue[wb + __popc(m & ((1u << lane) - 1))] = id;
                // The flush decision must be identical in every thread (it guards barriers): use the barrier's own
                // CTA-wide count, never a racy read of s_qn (faster warps may already be appending the next batch).
                qtotal += __syncthreads_count(hit);
                const int qn = qtotal;
                const bool last = (c == nchunks - 1) && (base + NT >= nin);
                if (qn > QCAP - NT || last) {
                    rasterise_queue(qn);
                    __syncthreads();
                    if (tid == 0) s_qn = 0;
                    qtotal = 0;
                    __syncthreads();
                }
            }
        }
        }
        __syncthreads();

        // ---------------- resolve + write
        float bg[3];
        load_bg(P, b, bg);
        const bool want_rgb = (P.flags & SDN_RGB) != 0;
        // (1) face-index map, int4 rows; mark the faces that own a sample
        {
            int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws;
            const bool vec = (P.Ws & 3) == 0;
            for (int k = tid; k < TH * (TW / 4); k += NT) {
                const int r = k / (TW / 4), c4 = (k % (TW / 4)) * 4;
                const int yy = sy0 + r, xx = sx0 + c4;
                if (yy > sy1 || xx > sx1) continue;
                int v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = (int)(unsigned)(s_key[r * TW + c4 + j] & 0xffffffffu);
                int32_t* dst = fim + (size_t)yy * P.Ws + xx;
                if (xx + 3 <= sx1 && vec) *reinterpret_cast<int4*>(dst) = make_int4(v[0], v[1], v[2], v[3]);
                else for (int j = 0; j < 4 && xx + j <= sx1; j++) dst[j] = v[j];
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (v[j] >= 0 && (j == 0 || v[j] != v[j - 1])) ws.visible[(size_t)b * P.F + v[j]] = 1;
            }
        }
        // (2) colour / alpha / depth: one output pixel (SS x SS samples) per thread-iteration
        constexpr int OW = TW / SS, OH = TH / SS;
        const size_t plane = (size_t)P.H * P.W;
        for (int k = tid; k < OW * OH; k += NT) {
            const int orow = k / OW, ocol = k % OW;
            const int ox = sx0 / SS + ocol, oyr = sy0 / SS + orow;    // oyr: row in raster orientation
            if (ox >= P.W || oyr >= P.H) continue;
            float a_sum = 0.0f, d_sum = 0.0f, c_sum[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int j = 0; j < SS * SS; j++) {
                const int lx = ocol * SS + (j % SS), ly = orow * SS + (j / SS);
                const u64 key = s_key[ly * TW + lx];
                const int fn = (int)(unsigned)(key & 0xffffffffu);
                const float zp = fn >= 0 ? __uint_as_float((unsigned)(key >> 32)) : far_f;
                float col[3] = {bg[0], bg[1], bg[2]};
                if (fn >= 0) {
                    a_sum += 1.0f;
                    if (want_rgb) shade_sample(P, ws.face_inv, b, fn, sx0 + lx, sy0 + ly, zp, col);
                }
                d_sum += zp;
                c_sum[0] += col[0]; c_sum[1] += col[1]; c_sum[2] += col[2];
                if (want_rgb && P.rgb_ss) {
                    float* o = P.rgb_ss + (((size_t)b * P.Hs + (sy0 + ly)) * P.Ws + (sx0 + lx)) * 3;
                    o[0] = col[0]; o[1] = col[1]; o[2] = col[2];
                }
            }
            const float scale = (SS == 2) ? 0.25f : 1.0f;
            const int oy = P.H - 1 - oyr;                               // rasterize.py:953-957 row flip
            const size_t po = (size_t)oy * P.W + ox;
            if (P.flags & SDN_ALPHA) P.alpha[(size_t)b * plane + po] = a_sum * scale;
            if (P.flags & SDN_DEPTH) P.depth[(size_t)b * plane + po] = d_sum * scale;
            if (want_rgb) {
                P.rgb[((size_t)b * 3 + 0) * plane + po] = c_sum[0] * scale;
                P.rgb[((size_t)b * 3 + 1) * plane + po] = c_sum[1] * scale;
                P.rgb[((size_t)b * 3 + 2) * plane + po] = c_sum[2] * scale;
            }
        }
        __syncthreads();
    }
}

// visible-face compaction for the backward: one global list of b*F+fn
__global__ void k_compact_visible(int B, int F, Workspace ws) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vis = (i < (long)B * F) && ws.visible[i];
    const unsigned lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(FULL, vis);
    if (!m) return;
    int base = 0;
    const int leader = __ffs(m) - 1;
    if ((int)lane == leader) base = atomicAdd(&ws.vis_count[0], __popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (vis) {
        const int slot = base + __popc(m & ((1u << lane) - 1));
        ws.vis_ids[slot] = (int)i;
        ws.vis_slot[i] = slot;
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
struct BwdParams {
    int B, F, H, W, ss, Hs, Ws, flags, ts, batch_bug;
    Clip clip;
    double eps_tex;
    int use_alpha, use_rgb, use_depth;      // which outputs the forward had (K6 runs iff alpha|rgb)
    int split;                              // 1: alpha and rgb are separate reference passes
    double eps_a, eps_r;                    // K6 pads (equal unless split)
    int fast_math;
    int line_path;                          // out-sweeps were deferred to k_line_sweep
    float bg[3];
    const float* bg_b;
    const float* faces;
    const float* textures;
    const int32_t* face_index;
    const float* rgb_ss;
    const float* g_rgb; const float* g_alpha; const float* g_depth;
    float* grad_textures;
};

// gradient of the SS raster sample (xi, yi): avg-pool backward (x 1/ss^2) + row un-flip
__device__ __forceinline__ size_t out_index(const BwdParams& P, int xi, int yi) {
    return (size_t)(P.H - 1 - yi / P.ss) * P.W + xi / P.ss;
}

struct PixelRef { float alpha; float c[3]; };

__device__ __forceinline__ void colour_at(const BwdParams& P, int b, size_t mi, int idx, const float* bg, float* c) {
    if (P.ts >= 2) {
        const float* s = P.rgb_ss + mi * 3;
        c[0] = s[0]; c[1] = s[1]; c[2] = s[2];
    } else if (idx >= 0) {
        const float* s = P.textures + ((size_t)b * P.F + idx) * 3;
        c[0] = __ldg(s); c[1] = __ldg(s + 1); c[2] = __ldg(s + 2);
    } else { c[0] = bg[0]; c[1] = bg[1]; c[2] = bg[2]; }
}

// dist of rasterize.py:646-647 for numerator a, denominator b (a/b hoisted in q)
__device__ __forceinline__ float k6_dist(float q, float t, float isf, double eps) {
    float d = q * t;
    d = __fdiv_rn(d * 2.0f, isf);
    return (0.0f < d) ? (float)((double)d + eps) : (float)((double)d - eps);
}

struct SweepCtx {
    // per (item, d0) constants
    float d1_cross, q0, q1;     // q0 = (p10-p00)/(p10-d0), q1 = (p10-p00)/(d0-p00)
    float k0, k1;               // q * 2 / is1 (fast path)
    bool use0, use1;            // p10 != d0, p00 != d0
    float ref_alpha, ref_c[3];  // alpha_in / rgb_in for the out sweep, *_out for the in sweep
};

__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// one visited pixel of a sweep: accumulates into g0 (vertex pi[0]) and g1 (vertex pi[1]).
// FAST = false: the IEEE op order of rasterize.py:628-654.  FAST = true: dist = (q*2/is)*t with a hardware
// reciprocal (relative error ~1e-7 per term, same inclusion tests).
template <bool FAST>
__device__ __forceinline__ void visit(const BwdParams& P, int b, const float* bg, size_t mi, int idx, int xi, int yi,
                                      int d1, const SweepCtx& s, float is1f, float& g0, float& g1) {
    const size_t oi = out_index(P, xi, yi);
    const float sc = (P.ss == 2) ? 0.25f : 1.0f;
    float da = 0.0f, dr = 0.0f;
    if (P.use_alpha) {
        const float ga = P.g_alpha ? __ldg(P.g_alpha + (size_t)b * P.H * P.W + oi) * sc : 0.0f;
        const float alpha = idx >= 0 ? 1.0f : 0.0f;
        da = (alpha - s.ref_alpha) * ga;
    }
    if (P.use_rgb) {
        float c[3];
        colour_at(P, b, mi, idx, bg, c);
        float g[3] = {0.0f, 0.0f, 0.0f};
        if (P.g_rgb) {
#pragma unroll
            for (int k = 0; k < 3; k++) g[k] = __ldg(P.g_rgb + ((size_t)b * 3 + k) * P.H * P.W + oi) * sc;
        }
        if (P.split || !P.use_alpha) {
            dr = (c[0] - s.ref_c[0]) * g[0];
            dr = dr + (c[1] - s.ref_c[1]) * g[1];
            dr = dr + (c[2] - s.ref_c[2]) * g[2];
        } else {        // one reference pass: alpha term first, then the three colour terms (rasterize.py:628-635)
            da = da + (c[0] - s.ref_c[0]) * g[0];
            da = da + (c[1] - s.ref_c[1]) * g[1];
            da = da + (c[2] - s.ref_c[2]) * g[2];
        }
    }
    const float t = (float)d1 - s.d1_cross;
    if (FAST) {
        const float pa = fmaxf(da, 0.0f), pr = fmaxf(dr, 0.0f);
        if (pa > 0.0f || pr > 0.0f) {
            const float ea = (float)P.eps_a, er = (float)P.eps_r;
            const float d0 = s.k0 * t, d1v = s.k1 * t;
            const float a0 = (0.0f < d0) ? d0 + ea : d0 - ea, r0 = (0.0f < d0) ? d0 + er : d0 - er;
            const float a1 = (0.0f < d1v) ? d1v + ea : d1v - ea, r1 = (0.0f < d1v) ? d1v + er : d1v - er;
            if (s.use0) g0 -= __fmaf_rn(pa, rcp_approx(a0), pr * rcp_approx(r0));
            if (s.use1) g1 -= __fmaf_rn(pa, rcp_approx(a1), pr * rcp_approx(r1));
        }
    } else {
        if (da > 0.0f) {
            if (s.use0) g0 -= __fdiv_rn(da, k6_dist(s.q0, t, is1f, P.eps_a));
            if (s.use1) g1 -= __fdiv_rn(da, k6_dist(s.q1, t, is1f, P.eps_a));
        }
        if (dr > 0.0f) {
            if (s.use0) g0 -= __fdiv_rn(dr, k6_dist(s.q0, t, is1f, P.eps_r));
            if (s.use1) g1 -= __fdiv_rn(dr, k6_dist(s.q1, t, is1f, P.eps_r));
        }
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// One out-sweep of rasterize.py:601-656, deferred to the line kernel: pixels lo..hi of raster line `line` (a row for
// axis 1, a column for axis 0), crossing coordinate xc, distance slopes k = q*2/is of the edge's two vertices.
struct __align__(16) Query {
    int lo;
    int hi_flags;       // bits 0..15 hi, 16 use0, 17 use1, 18 pad0 positive, 19 pad1 positive
    float xc, k0, k1;
    float c0, c1, c2;   // rgb_in
};
constexpr int QF_USE0 = 1 << 16, QF_USE1 = 1 << 17, QF_POS0 = 1 << 18, QF_POS1 = 1 << 19;

// phases of the item kernel
constexpr int PH_DIRECT = 0;    // everything inline (exact-math parity path)
constexpr int PH_COUNT = 1;     // count the out-sweep queries of each item and of each raster line
constexpr int PH_EMIT = 2;      // write the queries, run in-sweeps and interior gathers; overflow sweeps inline

// items per visible face: 0..5 = (edge, axis) of K6, 6 = interior gather (K8 + texture gradient)
constexpr int ITEMS = 7;
constexpr int BWD_WARPS = 4;

// layout of ws.scan_buf: [lines + 1 | 7 per visible face + 1 | 1 per visible face + 1]; after the exclusive scan the
// first region holds the per-line query bases, the second the per-item query bases, the third the per-face unit bases
constexpr int ITEMS_ = 7;
__device__ __forceinline__ int* item_region(const Workspace& ws) { return ws.scan_buf + ws.lines + 1; }
__device__ __forceinline__ int* unit_region(const Workspace& ws) { return ws.scan_buf + ws.lines + 1 + ws.vis_count[0] * ITEMS_ + 1; }
// the deferred work does not fit the workspace: the direct kernel (k_backward_items<true>) does everything instead
__device__ __forceinline__ bool bwd_overflow(const Workspace& ws) {
    const int* u = unit_region(ws);
    return ws.scan_buf[ws.lines] > ws.qmax || (u[ws.vis_count[0]] - u[0]) > ws.umax;
}
constexpr int UNIT_W = 32, UNIT_H = 4;      // samples per unit of k_bwd_face

// Direct form: every sweep runs inline, one warp per (face, edge, axis) item.  FAST = false is the exact-math parity
// path; FAST = true is the fallback of the line-sweep path when the queries do not fit the workspace
// (`only_on_overflow`): it returns at once unless that happened.
template <bool FAST>
__global__ void __launch_bounds__(BWD_WARPS * 32) k_backward_items(BwdParams P, Workspace ws, int only_on_overflow) {
    const int lane = threadIdx.x & 31;
    const int nitems = ws.vis_count[0] * ITEMS;
    if (only_on_overflow && !bwd_overflow(ws)) return;
    constexpr int PHASE = PH_DIRECT;
    constexpr int ticket = 1;
    int* item_n = ws.scan_buf + ws.lines + 1;
    const bool defer_ok = false;

    for (;;) {
        // dynamic scheduling: work per item varies by orders of magnitude (edge length x sweep length)
        int it = 0;
        if (lane == 0) it = atomicAdd(&ws.counters[ticket], 1);
        it = __shfl_sync(FULL, it, 0);
        if (it >= nitems) break;
        const size_t fo = (size_t)ws.vis_ids[it / ITEMS];
        const int kind = it % ITEMS;
        const int b = (int)(fo / P.F), fn = (int)(fo % P.F);
        const int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws;
        const size_t base_px = (size_t)b * P.Hs * P.Ws;
        float bg[3];
        if (P.bg_b) { bg[0] = P.bg_b[b * 3]; bg[1] = P.bg_b[b * 3 + 1]; bg[2] = P.bg_b[b * 3 + 2]; }
        else { bg[0] = P.bg[0]; bg[1] = P.bg[1]; bg[2] = P.bg[2]; }
        float f[9];
#pragma unroll
        for (int k = 0; k < 9; k++) f[k] = __ldg(P.faces + fo * 9 + k);

        if (kind == 6) {
            if (PHASE == PH_COUNT) { if (lane == 0) item_n[it] = 0; continue; }
            // ---------------- interior gather: samples this face won
            float S[3] = {0.0f, 0.0f, 0.0f}, G[3] = {0.0f, 0.0f, 0.0f};
            const bool need_depth = P.use_depth && P.g_depth;
            const bool need_tex = P.use_rgb && P.g_rgb && P.grad_textures;
            if (need_depth || need_tex) {
                float inv[9];
#pragma unroll
                for (int k = 0; k < 9; k++) inv[k] = __ldg(ws.face_inv + fo * 9 + k);
                const short4 bb = *reinterpret_cast<const short4*>(ws.face_bbox + fo * 4);
                const float sc = (P.ss == 2) ? 0.25f : 1.0f;
                const int ts = P.ts == 1 ? 2 : P.ts;
                float zf[9];
                const float* zsrc = P.faces + ((size_t)(P.batch_bug ? 0 : b) * P.F + fn) * 9;
#pragma unroll
                for (int k = 0; k < 9; k++) zf[k] = __ldg(zsrc + k);
                for (int yy = bb.y + (lane >> 3); yy <= bb.w; yy += 4) {
                    const float yp = xp_of(yy, P.Hs);
                    for (int xx = bb.x + (lane & 7); xx <= bb.z; xx += 8) {
                        if (fim[(size_t)yy * P.Ws + xx] != fn) continue;
                        float w[3], zp;
                        if (!eval_sample(f, inv, xp_of(xx, P.Ws), yp, (float)xx, (float)yy, P.clip, w, zp)) continue;
                        const size_t oi = out_index(P, xx, yy);
                        if (need_depth) {      // rasterize.py:821-824: sum_s gd * w_k * depth^2
                            const float gd = P.g_depth[(size_t)b * P.H * P.W + oi] * sc;
                            const float d2 = zp * zp;
#pragma unroll
                            for (int k = 0; k < 3; k++) S[k] += gd * w[k] * d2;
                        }
                        if (need_tex) {        // rasterize.py:772-778
                            float g[3];
#pragma unroll
                            for (int k = 0; k < 3; k++) g[k] = P.g_rgb[((size_t)b * 3 + k) * P.H * P.W + oi] * sc;
                            float tif[3];
                            texture_index(w, zp, zf, ts, P.eps_tex, tif);
                            if (P.ts == 1) {
                                float wsum = 0.0f;
#pragma unroll
                                for (int pn = 0; pn < 8; pn++) { int tii[3]; wsum += corner_weight(tif, pn, tii); }
#pragma unroll
                                for (int k = 0; k < 3; k++) G[k] += wsum * g[k];
                            } else {
                                float* gt = P.grad_textures + fo * ts * ts * ts * 3;
#pragma unroll
                                for (int pn = 0; pn < 8; pn++) {
                                    int tii[3];
                                    const float cw = corner_weight(tif, pn, tii);
                                    const int isc = tii[0] * ts * ts + tii[1] * ts + tii[2];
#pragma unroll
                                    for (int k = 0; k < 3; k++) atomicAdd(gt + isc * 3 + k, cw * g[k]);
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) { S[k] = warp_sum(S[k]); G[k] = warp_sum(G[k]); }
            if (lane == 0) {
                float* o = ws.gint + fo * 6;
                o[0] = S[0]; o[1] = S[1]; o[2] = S[2]; o[3] = G[0]; o[4] = G[1]; o[5] = G[2];
            }
            continue;
        }

        // ---------------- K6 item (edge, axis)   rasterize.py:540-729
        const int edge = kind >> 1, axis = kind & 1;
        float g0 = 0.0f, g1 = 0.0f;
        int nq = 0;                                     // queries of this item so far (warp-uniform)
        const int qbase = PHASE == PH_EMIT ? item_n[it] - item_n[0] : 0;
        const int line0 = b * (P.Hs + P.Ws) + (axis == 0 ? P.Hs : 0);
        if (P.use_alpha || P.use_rgb) {
            int pi[3];
            float pp[3][2], p[3][2];
#pragma unroll
            for (int n = 0; n < 3; n++) pi[n] = (edge + n) % 3;
#pragma unroll
            for (int n = 0; n < 3; n++) { pp[n][0] = to_pixel(f[3 * pi[n]], P.Ws); pp[n][1] = to_pixel(f[3 * pi[n] + 1], P.Hs); }
#pragma unroll
            for (int n = 0; n < 3; n++) { p[n][0] = pp[n][axis]; p[n][1] = pp[n][1 - axis]; }
            const int is0 = axis == 0 ? P.Ws : P.Hs, is1 = axis == 0 ? P.Hs : P.Ws;
            const float is1f = (float)is1;
            int direction;
            if (axis == 0) direction = (p[0][0] < p[1][0]) ? -1 : 1;
            else direction = (p[0][0] < p[1][0]) ? 1 : -1;
            const int d0_from = __double2int_rz(fmax((double)ceilf(fminf(p[0][0], p[1][0])), 0.));
            const int d0_to = __double2int_rz(fmin((double)fmaxf(p[0][0], p[1][0]), is0 - 1.));
            const float slope = __fdiv_rn(p[1][1] - p[0][1], p[1][0] - p[0][0]);
            const float dp = p[1][0] - p[0][0];

            for (int sbase = d0_from; sbase <= d0_to; sbase += 32) {
                const int d0 = sbase + lane;
                const float d0f = (float)d0;
                bool valid = d0 <= d0_to;
                SweepCtx so;     // out-sweep context (refs = in pixel)
                SweepCtx si;     // in-sweep context (refs = out pixel)
                int d1_in = 0, d1_out = 0;
                bool is_in_fn = false;
                {
                    float dc = slope * (d0f - p[0][0]);
                    dc = dc + p[0][1];
                    so.d1_cross = dc;
                    d1_in = (0 < direction) ? __float2int_rz(floorf(dc)) : __float2int_rz(ceilf(dc));
                    if (d1_in < 0 || is1 <= d1_in) valid = false;
                    d1_out = d1_in + direction;
                    if (valid && (d1_out < 0 || is1 <= d1_out)) valid = false;
                    so.use0 = p[1][0] != d0f; so.use1 = p[0][0] != d0f;
                    so.q0 = __fdiv_rn(dp, p[1][0] - d0f); so.q1 = __fdiv_rn(dp, d0f - p[0][0]);
                    so.k0 = so.q0 * (2.0f / is1f); so.k1 = so.q1 * (2.0f / is1f);
                    si = so;
                }
                if (valid) {
                    const int xin = axis == 0 ? d0 : d1_in, yin = axis == 0 ? d1_in : d0;
                    const int xout = axis == 0 ? d0 : d1_out, yout = axis == 0 ? d1_out : d0;
                    const size_t mi_in = (size_t)yin * P.Ws + xin, mi_out = (size_t)yout * P.Ws + xout;
                    const int idx_in = fim[mi_in], idx_out = fim[mi_out];
                    is_in_fn = idx_in == fn;
                    so.ref_alpha = idx_in >= 0 ? 1.0f : 0.0f;
                    si.ref_alpha = idx_out >= 0 ? 1.0f : 0.0f;
                    if (PHASE != PH_COUNT && P.use_rgb) {
                        colour_at(P, b, base_px + mi_in, idx_in, bg, so.ref_c);
                        colour_at(P, b, base_px + mi_out, idx_out, bg, si.ref_c);
                    }
                    // ---- in sweep (short, lane-serial)   rasterize.py:659-727
                    float dc2;
                    if ((d0f - p[0][0]) * (d0f - p[2][0]) < 0) {
                        dc2 = __fdiv_rn(p[2][1] - p[0][1], p[2][0] - p[0][0]);
                        dc2 = dc2 * (d0f - p[0][0]);
                        dc2 = dc2 + p[0][1];
                    } else {
                        dc2 = __fdiv_rn(p[1][1] - p[2][1], p[1][0] - p[2][0]);
                        dc2 = dc2 * (d0f - p[2][0]);
                        dc2 = dc2 + p[2][1];
                    }
                    const int d1_limit = (0 < direction) ? __float2int_rz(ceilf(dc2)) : __float2int_rz(floorf(dc2));
                    const int d1_from = max(min(d1_in, d1_limit), 0);
                    const int d1_to = PHASE == PH_COUNT ? -1 : min(max(d1_in, d1_limit), is1 - 1);
                    for (int d1 = d1_from; d1 <= d1_to; d1++) {
                        const int xi = axis == 0 ? d0 : d1, yi = axis == 0 ? d1 : d0;
                        const size_t mi = (size_t)yi * P.Ws + xi;
                        if (fim[mi] != fn) continue;
                        visit<FAST>(P, b, bg, base_px + mi, fn, xi, yi, d1, si, is1f, g0, g1);
                    }
                }
                // ---- out sweeps   rasterize.py:601-656
                bool sweep = valid && is_in_fn;
                if (PHASE != PH_DIRECT) {
                    const unsigned m = __ballot_sync(FULL, sweep);
                    const int qi = qbase + nq + __popc(m & ((1u << lane) - 1));
                    nq += __popc(m);
                    if (PHASE == PH_COUNT) {
                        if (sweep) atomicAdd(&ws.scan_buf[line0 + d0], 1);
                        continue;
                    }
                    if (sweep && defer_ok) {            // defer to the line kernel
                        const int d1_limit = (0 < direction) ? is1 - 1 : 0;
                        Query q;
                        q.lo = max(min(d1_out, d1_limit), 0);
                        int hf = min(max(d1_out, d1_limit), is1 - 1);
                        if (so.use0) hf |= QF_USE0;
                        if (so.use1) hf |= QF_USE1;
                        if ((so.k0 > 0.0f) == (direction > 0) && so.k0 != 0.0f) hf |= QF_POS0;
                        if ((so.k1 > 0.0f) == (direction > 0) && so.k1 != 0.0f) hf |= QF_POS1;
                        q.hi_flags = hf;
                        q.xc = so.d1_cross; q.k0 = so.use0 ? so.k0 : 0.0f; q.k1 = so.use1 ? so.k1 : 0.0f;
                        q.c0 = so.ref_c[0]; q.c1 = so.ref_c[1]; q.c2 = so.ref_c[2];
                        ws.queries[qi] = q;
                        const int line = line0 + d0;
                        const int pos = atomicAdd(&ws.line_cursor[line], 1);
                        ws.line_items[ws.scan_buf[line] + pos] = qi;
                        sweep = false;
                    }
                }
                if (axis == 0) {
                    // vertical: every lane owns a column; iterate absolute rows so loads stay coalesced
                    const int d1_limit = (0 < direction) ? is1 - 1 : 0;
                    int lo = max(min(d1_out, d1_limit), 0), hi = min(max(d1_out, d1_limit), is1 - 1);
                    if (!sweep) { lo = 1; hi = 0; }
                    int wlo = sweep ? lo : 0x7fffffff, whi = sweep ? hi : -1;
                    wlo = __reduce_min_sync(FULL, wlo); whi = __reduce_max_sync(FULL, whi);
                    for (int d1 = wlo; d1 <= whi; d1++) {
                        if (d1 < lo || d1 > hi) continue;
                        const size_t mi = (size_t)d1 * P.Ws + d0;
                        visit<FAST>(P, b, bg, base_px + mi, fim[mi], d0, d1, d1, so, is1f, g0, g1);
                    }
                } else {
                    // horizontal: the warp takes the steps one at a time, lanes stride along the row
                    unsigned todo = __ballot_sync(FULL, sweep);
                    while (todo) {
                        const int src = __ffs(todo) - 1;
                        todo &= todo - 1;
                        SweepCtx s;
                        s.d1_cross = __shfl_sync(FULL, so.d1_cross, src);
                        s.q0 = __shfl_sync(FULL, so.q0, src); s.q1 = __shfl_sync(FULL, so.q1, src);
                        s.k0 = __shfl_sync(FULL, so.k0, src); s.k1 = __shfl_sync(FULL, so.k1, src);
                        s.use0 = __shfl_sync(FULL, (int)so.use0, src); s.use1 = __shfl_sync(FULL, (int)so.use1, src);
                        s.ref_alpha = __shfl_sync(FULL, so.ref_alpha, src);
#pragma unroll
                        for (int k = 0; k < 3; k++) s.ref_c[k] = __shfl_sync(FULL, so.ref_c[k], src);
                        const int row = sbase + src;
                        const int dout = __shfl_sync(FULL, d1_out, src);
                        const int d1_limit = (0 < direction) ? is1 - 1 : 0;
                        const int lo = max(min(dout, d1_limit), 0), hi = min(max(dout, d1_limit), is1 - 1);
                        for (int d1 = lo + lane; d1 <= hi; d1 += 32) {
                            const size_t mi = (size_t)row * P.Ws + d1;
                            visit<FAST>(P, b, bg, base_px + mi, fim[mi], d1, row, d1, s, is1f, g0, g1);
                        }
                    }
                }
            }
        }
        if (PHASE == PH_COUNT) { if (lane == 0) item_n[it] = nq; continue; }
        g0 = warp_sum(g0); g1 = warp_sum(g1);
        if (lane == 0) { ws.gpart[fo * 12 + kind * 2 + 0] = g0; ws.gpart[fo * 12 + kind * 2 + 1] = g1; }
    }
}



// ------------------------------------------------------------------------------------------------------------
// fast path, step 1: enumerate the edge steps of every visible face (rasterize.py:540-599) -- one warp per face,
// the six (edge, axis) items in turn, lanes over the steps d0.  EMIT = false counts the out-sweep queries per item and
// per raster line; EMIT = true (after the scan) writes them.
// ------------------------------------------------------------------------------------------------------------
struct EdgeSetup {
    float p[3][2];
    int direction, d0_from, d0_to, is0, is1;
    float slope, dp;
};

__device__ __forceinline__ void edge_setup(const float* f, int edge, int axis, int Hs, int Ws, EdgeSetup& E) {
    int pi[3];
    float pp[3][2];
#pragma unroll
    for (int n = 0; n < 3; n++) pi[n] = (edge + n) % 3;
#pragma unroll
    for (int n = 0; n < 3; n++) { pp[n][0] = to_pixel(f[3 * pi[n]], Ws); pp[n][1] = to_pixel(f[3 * pi[n] + 1], Hs); }
#pragma unroll
    for (int n = 0; n < 3; n++) { E.p[n][0] = pp[n][axis]; E.p[n][1] = pp[n][1 - axis]; }
    E.is0 = axis == 0 ? Ws : Hs; E.is1 = axis == 0 ? Hs : Ws;
    if (axis == 0) E.direction = (E.p[0][0] < E.p[1][0]) ? -1 : 1;
    else E.direction = (E.p[0][0] < E.p[1][0]) ? 1 : -1;
    E.d0_from = __double2int_rz(fmax((double)ceilf(fminf(E.p[0][0], E.p[1][0])), 0.));
    E.d0_to = __double2int_rz(fmin((double)fmaxf(E.p[0][0], E.p[1][0]), E.is0 - 1.));
    E.slope = __fdiv_rn(E.p[1][1] - E.p[0][1], E.p[1][0] - E.p[0][0]);
    E.dp = E.p[1][0] - E.p[0][0];
}

// crossing of step d0: returns false when the step is skipped by rasterize.py:575-576
__device__ __forceinline__ bool step_cross(const EdgeSetup& E, int d0, float& d1_cross, int& d1_in, int& d1_out) {
    const float d0f = (float)d0;
    float dc = E.slope * (d0f - E.p[0][0]);
    dc = dc + E.p[0][1];
    d1_cross = dc;
    d1_in = (0 < E.direction) ? __float2int_rz(floorf(dc)) : __float2int_rz(ceilf(dc));
    if (d1_in < 0 || E.is1 <= d1_in) return false;
    d1_out = d1_in + E.direction;
    return !(d1_out < 0 || E.is1 <= d1_out);
}

// A face is "small" when its bbox fits 8x8 samples: ShapeNet-sized meshes consist almost entirely of such faces, and
// warp-per-face stages would idle most lanes on them, so they get lane-per-item / lane-per-face paths.
__device__ __forceinline__ bool small_face(const short4 bb) { return (bb.z - bb.x) < 8 && (bb.w - bb.y) < 8; }
// A lane-per-face path for k_bwd_face was tried for tiny faces and measured slower on both workloads (the per-face
// state costs 164 registers and serialises the samples); every face therefore goes through the 4x32-sample units.
__device__ __forceinline__ bool tiny_face(const short4) { return false; }

// writes the out-sweep query of one edge step (rasterize.py:601-656, deferred to k_line_sweep) and files it under its line
__device__ __forceinline__ void emit_query(const BwdParams& P, const Workspace& ws, const EdgeSetup& E, int b, int fn,
                                           const float* bg, int axis, int d0, float d1_cross, int d1_out, size_t mi_in,
                                           int qi, int line) {
    const float d0f = (float)d0;
    const bool use0 = E.p[1][0] != d0f, use1 = E.p[0][0] != d0f;
    const float k0 = __fdiv_rn(E.dp, E.p[1][0] - d0f) * (2.0f / (float)E.is1);
    const float k1 = __fdiv_rn(E.dp, d0f - E.p[0][0]) * (2.0f / (float)E.is1);
    const int d1_limit = (0 < E.direction) ? E.is1 - 1 : 0;
    Query q;
    q.lo = max(min(d1_out, d1_limit), 0);
    int hf = min(max(d1_out, d1_limit), E.is1 - 1);
    if (use0) hf |= QF_USE0;
    if (use1) hf |= QF_USE1;
    if ((k0 > 0.0f) == (E.direction > 0) && k0 != 0.0f) hf |= QF_POS0;
    if ((k1 > 0.0f) == (E.direction > 0) && k1 != 0.0f) hf |= QF_POS1;
    q.hi_flags = hf;
    q.xc = d1_cross; q.k0 = use0 ? k0 : 0.0f; q.k1 = use1 ? k1 : 0.0f;
    float c[3] = {0.0f, 0.0f, 0.0f};
    if (P.use_rgb) colour_at(P, b, (size_t)b * P.Hs * P.Ws + mi_in, fn, bg, c);
    q.c0 = c[0]; q.c1 = c[1]; q.c2 = c[2];
    ws.queries[qi] = q;
    const int pos = atomicAdd(&ws.line_cursor[line], 1);
    ws.line_items[ws.scan_buf[line] + pos] = qi;
}

template <bool EMIT>
__global__ void __launch_bounds__(256) k_bwd_steps(BwdParams P, Workspace ws) {
    const int lane = threadIdx.x & 31;
    const int nvis = ws.vis_count[0];
    int* item_n = ws.scan_buf + ws.lines + 1;       // counts, then (after the scan) exclusive bases
    int* unit_n = unit_region(ws);
    if (!EMIT && blockIdx.x == 0 && threadIdx.x == 0) { item_n[nvis * ITEMS] = 0; unit_n[nvis] = 0; }   // scan sentinels
    if (EMIT && bwd_overflow(ws)) return;      // does not fit: k_backward_items<true> takes over
    const int q_origin = EMIT ? item_n[0] : 0;
    const int gwarp = blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = gridDim.x * 8;

    // ---- (A) larger faces: one warp per face, the six (edge, axis) items in turn, lanes over the steps
    for (int vi = gwarp; vi < nvis; vi += nwarps) {
        const size_t fo = (size_t)ws.vis_ids[vi];
        const short4 bb = *reinterpret_cast<const short4*>(ws.face_bbox + fo * 4);
        if (small_face(bb)) continue;
        const int b = (int)(fo / P.F), fn = (int)(fo % P.F);
        const int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws;
        float f[9];
#pragma unroll
        for (int k = 0; k < 9; k++) f[k] = __ldg(P.faces + fo * 9 + k);
        float bg[3];
        if (P.bg_b) { bg[0] = P.bg_b[b * 3]; bg[1] = P.bg_b[b * 3 + 1]; bg[2] = P.bg_b[b * 3 + 2]; }
        else { bg[0] = P.bg[0]; bg[1] = P.bg[1]; bg[2] = P.bg[2]; }
        {   // units of k_bwd_face: the face's bbox cut into 4-row x 32-column blocks
            const int nbx = (bb.z - bb.x) / UNIT_W + 1, nby = (bb.w - bb.y) / UNIT_H + 1;
            if (!EMIT) { if (lane == 0) { item_n[vi * ITEMS + 6] = 0; unit_n[vi] = nbx * nby; } }
            else {
                const int ub = unit_n[vi] - unit_n[0];
                for (int u = lane; u < nbx * nby; u += 32) ws.unit_tab[ub + u] = make_int2(vi, u);
            }
        }
        for (int kind = 0; kind < 6; kind++) {
            const int axis = kind & 1;
            EdgeSetup E;
            edge_setup(f, kind >> 1, axis, P.Hs, P.Ws, E);
            const int line0 = b * (P.Hs + P.Ws) + (axis == 0 ? P.Hs : 0);
            const int qbase = EMIT ? item_n[vi * ITEMS + kind] - q_origin : 0;
            int nq = 0;
            for (int sbase = E.d0_from; sbase <= E.d0_to; sbase += 32) {
                const int d0 = sbase + lane;
                float d1_cross;
                int d1_in = 0, d1_out = 0;
                bool sweep = (d0 <= E.d0_to) && step_cross(E, d0, d1_cross, d1_in, d1_out);
                size_t mi_in = 0;
                if (sweep) {
                    mi_in = axis == 0 ? (size_t)d1_in * P.Ws + d0 : (size_t)d0 * P.Ws + d1_in;
                    sweep = __ldg(fim + mi_in) == fn;              // rasterize.py:601
                }
                const unsigned m = __ballot_sync(FULL, sweep);
                const int qi = qbase + nq + __popc(m & ((1u << lane) - 1));
                nq += __popc(m);
                if (!sweep) continue;
                const int line = line0 + d0;
                if (!EMIT) { atomicAdd(&ws.scan_buf[line], 1); continue; }
                emit_query(P, ws, E, b, fn, bg, axis, d0, d1_cross, d1_out, mi_in, qi, line);
            }
            if (!EMIT && lane == 0) item_n[vi * ITEMS + kind] = nq;
        }
    }

    // ---- (B) small faces: one LANE per (face, edge, axis) item, five faces per warp, steps walked serially
    for (int base = gwarp * 5; base < nvis; base += nwarps * 5) {
        const int vi = base + lane / 6, kind = lane % 6;
        if (lane >= 30 || vi >= nvis) continue;
        const size_t fo = (size_t)ws.vis_ids[vi];
        const short4 bb = *reinterpret_cast<const short4*>(ws.face_bbox + fo * 4);
        if (!small_face(bb)) continue;
        const int b = (int)(fo / P.F), fn = (int)(fo % P.F);
        const int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws;
        float f[9];
#pragma unroll
        for (int k = 0; k < 9; k++) f[k] = __ldg(P.faces + fo * 9 + k);
        float bg[3];
        if (P.bg_b) { bg[0] = P.bg_b[b * 3]; bg[1] = P.bg_b[b * 3 + 1]; bg[2] = P.bg_b[b * 3 + 2]; }
        else { bg[0] = P.bg[0]; bg[1] = P.bg[1]; bg[2] = P.bg[2]; }
        if (kind == 0) {        // units of k_bwd_face: none for tiny faces (its lane path), else the 4x32 blocks of the bbox
            const int nun = tiny_face(bb) ? 0 : ((bb.z - bb.x) / UNIT_W + 1) * ((bb.w - bb.y) / UNIT_H + 1);
            if (!EMIT) { item_n[vi * ITEMS + 6] = 0; unit_n[vi] = nun; }
            else { const int ub = unit_n[vi] - unit_n[0]; for (int u = 0; u < nun; u++) ws.unit_tab[ub + u] = make_int2(vi, u); }
        }
        const int axis = kind & 1;
        EdgeSetup E;
        edge_setup(f, kind >> 1, axis, P.Hs, P.Ws, E);
        const int line0 = b * (P.Hs + P.Ws) + (axis == 0 ? P.Hs : 0);
        const int qbase = EMIT ? item_n[vi * ITEMS + kind] - q_origin : 0;
        int nq = 0;
        for (int d0 = E.d0_from; d0 <= E.d0_to; d0++) {
            float d1_cross;
            int d1_in = 0, d1_out = 0;
            if (!step_cross(E, d0, d1_cross, d1_in, d1_out)) continue;
            const size_t mi_in = axis == 0 ? (size_t)d1_in * P.Ws + d0 : (size_t)d0 * P.Ws + d1_in;
            if (__ldg(fim + mi_in) != fn) continue;                // rasterize.py:601
            const int line = line0 + d0;
            if (!EMIT) atomicAdd(&ws.scan_buf[line], 1);
            else emit_query(P, ws, E, b, fn, bg, axis, d0, d1_cross, d1_out, mi_in, qbase + nq, line);
            nq++;
        }
        if (!EMIT) item_n[vi * ITEMS + kind] = nq;
    }
}


// Backward-only: weights and depth of a sample the face is known to own (no inclusion tests, reciprocal arithmetic).
__device__ __forceinline__ void eval_owned_fast(const float* f, const float* inv, float xif, float yif, float* w,
                                                float& zp) {
#pragma unroll
    for (int k = 0; k < 3; k++)
        w[k] = fminf(fmaxf(__fmaf_rn(inv[3 * k], xif, __fmaf_rn(inv[3 * k + 1], yif, inv[3 * k + 2])), 0.0f), 1.0f);
    const float rs = rcp_approx(w[0] + w[1] + w[2]);
    w[0] *= rs; w[1] *= rs; w[2] *= rs;
    zp = rcp_approx(w[0] * rcp_approx(f[2]) + w[1] * rcp_approx(f[5]) + w[2] * rcp_approx(f[8]));
}

// ------------------------------------------------------------------------------------------------------------
// fast path, step 2: everything that happens ON the samples a face owns -- the depth / texture gathers
// (rasterize.py:747-844) and the in-sweeps (rasterize.py:659-727).  The reference walks each edge step across the
// triangle; here each owned sample asks, for the six (edge, axis) items, whether the step through its own row / column
// would have visited it (identical float tests), so the cost is proportional to the owned area.
// ------------------------------------------------------------------------------------------------------------
// per-face constants of k_bwd_face (registers)
struct FaceCtx {
    size_t fo, base_px;
    int b, fn, ts;
    const int32_t* fim;
    const float *ga_p, *gr_p, *gd_p;
    float f[9], inv[9], zf[9], bg[3];
    float pc[3][2];          // pixel coordinates [vertex][x|y]
    float slope[2][3];       // [axis][edge e = (e, e+1)]: exact IEEE quotient, orientation-independent
    int d0_from[6], d0_to[6];
    short4 bb;
};

__device__ __forceinline__ void face_ctx_load(const BwdParams& P, const Workspace& ws, size_t fo, bool need_depth,
                                              FaceCtx& C) {
    const size_t plane = (size_t)P.H * P.W;
    C.fo = fo;
    C.b = (int)(fo / P.F); C.fn = (int)(fo % P.F);
    C.fim = P.face_index + (size_t)C.b * P.Hs * P.Ws;
    C.base_px = (size_t)C.b * P.Hs * P.Ws;
#pragma unroll
    for (int k = 0; k < 9; k++) { C.f[k] = __ldg(P.faces + fo * 9 + k); C.inv[k] = __ldg(ws.face_inv + fo * 9 + k); }
    const float* zsrc = P.faces + ((size_t)(P.batch_bug ? 0 : C.b) * P.F + C.fn) * 9;
#pragma unroll
    for (int k = 0; k < 9; k++) C.zf[k] = __ldg(zsrc + k);
    C.bb = *reinterpret_cast<const short4*>(ws.face_bbox + fo * 4);
    if (P.bg_b) { C.bg[0] = P.bg_b[C.b * 3]; C.bg[1] = P.bg_b[C.b * 3 + 1]; C.bg[2] = P.bg_b[C.b * 3 + 2]; }
    else { C.bg[0] = P.bg[0]; C.bg[1] = P.bg[1]; C.bg[2] = P.bg[2]; }
    C.ga_p = (P.use_alpha && P.g_alpha) ? P.g_alpha + (size_t)C.b * plane : nullptr;
    C.gr_p = (P.use_rgb && P.g_rgb) ? P.g_rgb + (size_t)C.b * 3 * plane : nullptr;
    C.gd_p = need_depth ? P.g_depth + (size_t)C.b * plane : nullptr;
    C.ts = P.ts == 1 ? 2 : P.ts;
    // constants of the six (edge, axis) items (rasterize.py:540-566), hoisted
#pragma unroll
    for (int n = 0; n < 3; n++) { C.pc[n][0] = to_pixel(C.f[3 * n], P.Ws); C.pc[n][1] = to_pixel(C.f[3 * n + 1], P.Hs); }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int e = 0; e < 3; e++)
            C.slope[a][e] = __fdiv_rn(C.pc[(e + 1) % 3][1 - a] - C.pc[e][1 - a], C.pc[(e + 1) % 3][a] - C.pc[e][a]);
#pragma unroll
    for (int kind = 0; kind < 6; kind++) {
        const int e = kind >> 1, a = kind & 1;
        const float p00 = C.pc[e][a], p10 = C.pc[(e + 1) % 3][a];
        const int is0 = a == 0 ? P.Ws : P.Hs;
        C.d0_from[kind] = __double2int_rz(fmax((double)ceilf(fminf(p00, p10)), 0.));
        C.d0_to[kind] = __double2int_rz(fmin((double)fmaxf(p00, p10), is0 - 1.));
    }
}

// everything one owned sample (xx, yy) contributes: depth / texture gathers (rasterize.py:747-844) and the in-sweeps
// (rasterize.py:659-727: would the step through this sample's row / column have visited it?  same float tests)
__device__ __forceinline__ void face_sample(const BwdParams& P, const FaceCtx& C, int xx, int yy, bool need_depth,
                                            bool need_tex, bool need_k6, float* acc, float* S, float* G) {
    const float sc = (P.ss == 2) ? 0.25f : 1.0f;
    const size_t plane = (size_t)P.H * P.W;
    const float ea = (float)P.eps_a, er = (float)P.eps_r;
    const bool one_pass = !(P.split || !P.use_alpha);       // alpha and rgb summed before the > 0 test
    const int fn = C.fn, b = C.b;
    const size_t mi = (size_t)yy * P.Ws + xx;
    const size_t oi = out_index(P, xx, yy);
    // gradients at this sample (one set of loads for everything below)
    const float ga = C.ga_p ? __ldg(C.ga_p + oi) * sc : 0.0f;
    float g[3] = {0.0f, 0.0f, 0.0f};
    if (C.gr_p) { g[0] = __ldg(C.gr_p + oi) * sc; g[1] = __ldg(C.gr_p + plane + oi) * sc; g[2] = __ldg(C.gr_p + 2 * plane + oi) * sc; }

    if (need_depth || need_tex) {
        float w[3], zp;
        eval_owned_fast(C.f, C.inv, (float)xx, (float)yy, w, zp);
        if (need_depth) {      // rasterize.py:821-824: sum_s gd * w_k * depth^2
            const float gd = __ldg(C.gd_p + oi) * sc;
            const float d2 = zp * zp;
#pragma unroll
            for (int k = 0; k < 3; k++) S[k] += gd * w[k] * d2;
        }
        if (need_tex) {        // rasterize.py:772-778
            float tif[3];
            texture_index(w, zp, C.zf, C.ts, P.eps_tex, tif);
            if (P.ts == 1) {
                float wsum = 0.0f;
#pragma unroll
                for (int pn = 0; pn < 8; pn++) { int tii[3]; wsum += corner_weight(tif, pn, tii); }
#pragma unroll
                for (int k = 0; k < 3; k++) G[k] += wsum * g[k];
            } else {
                const int ts = C.ts;
                float* gt = P.grad_textures + C.fo * ts * ts * ts * 3;
#pragma unroll
                for (int pn = 0; pn < 8; pn++) {
                    int tii[3];
                    const float cw = corner_weight(tif, pn, tii);
                    const int isc = tii[0] * ts * ts + tii[1] * ts + tii[2];
#pragma unroll
                    for (int k = 0; k < 3; k++) atomicAdd(gt + isc * 3 + k, cw * g[k]);
                }
            }
        }
    }
    if (!need_k6) return;

    float csel[3] = {0.0f, 0.0f, 0.0f};
    if (P.use_rgb) colour_at(P, b, C.base_px + mi, fn, C.bg, csel);
    float d1c[6];
    int d1o[6];
    unsigned inc = 0;
#pragma unroll
    for (int kind = 0; kind < 6; kind++) {
        const int e = kind >> 1, a = kind & 1;
        const int i0 = e, i1 = (e + 1) % 3, i2 = (e + 2) % 3;
        const float p00 = C.pc[i0][a], p01 = C.pc[i0][1 - a], p10 = C.pc[i1][a];
        const float p20 = C.pc[i2][a], p21 = C.pc[i2][1 - a];
        const int is1 = a == 0 ? P.Hs : P.Ws;
        const int d0 = a == 0 ? xx : yy, d1 = a == 0 ? yy : xx;
        const float d0f = (float)d0;
        int direction;
        if (a == 0) direction = (p00 < p10) ? -1 : 1; else direction = (p00 < p10) ? 1 : -1;
        float dc = C.slope[a][e] * (d0f - p00);
        dc = dc + p01;
        const int d1_in = (0 < direction) ? __float2int_rz(floorf(dc)) : __float2int_rz(ceilf(dc));
        const int d1_out = d1_in + direction;
        float dc2;                                   // rasterize.py:662-667
        if ((d0f - p00) * (d0f - p20) < 0) { dc2 = C.slope[a][i2] * (d0f - p00); dc2 = dc2 + p01; }
        else { dc2 = C.slope[a][i1] * (d0f - p20); dc2 = dc2 + p21; }
        const int d1_limit = (0 < direction) ? __float2int_rz(ceilf(dc2)) : __float2int_rz(floorf(dc2));
        const int d1_from = max(min(d1_in, d1_limit), 0), d1_to = min(max(d1_in, d1_limit), is1 - 1);
        const bool ok = d0 >= C.d0_from[kind] && d0 <= C.d0_to[kind] && d1_in >= 0 && d1_in < is1 &&
                        d1_out >= 0 && d1_out < is1 && d1 >= d1_from && d1 <= d1_to;
        d1c[kind] = dc; d1o[kind] = d1_out;
        if (ok) inc |= 1u << kind;
    }
    if (!inc) return;
    int ido[6];
#pragma unroll
    for (int kind = 0; kind < 6; kind++) {      // the out-pixel loads go out together
        const int a = kind & 1;
        const size_t mo = a == 0 ? (size_t)d1o[kind] * P.Ws + xx : (size_t)yy * P.Ws + d1o[kind];
        ido[kind] = (inc >> kind) & 1 ? __ldg(C.fim + mo) : -1;
    }
#pragma unroll
    for (int kind = 0; kind < 6; kind++) {
        if (!((inc >> kind) & 1)) continue;
        const int e = kind >> 1, a = kind & 1;
        const float p00 = C.pc[e][a], p10 = C.pc[(e + 1) % 3][a];
        const int is1 = a == 0 ? P.Hs : P.Ws;
        const int d0 = a == 0 ? xx : yy, d1 = a == 0 ? yy : xx;
        const float d0f = (float)d0;
        const size_t mo = a == 0 ? (size_t)d1o[kind] * P.Ws + xx : (size_t)yy * P.Ws + d1o[kind];
        float da = ((1.0f - (ido[kind] >= 0 ? 1.0f : 0.0f)) * ga), dr = 0.0f;     // alpha here is 1
        if (P.use_rgb) {
            float co[3];
            colour_at(P, b, C.base_px + mo, ido[kind], C.bg, co);
            const float d = (csel[0] - co[0]) * g[0] + (csel[1] - co[1]) * g[1] + (csel[2] - co[2]) * g[2];
            if (one_pass) da += d; else dr = d;
        }
        const float pa = fmaxf(da, 0.0f), pr = fmaxf(dr, 0.0f);
        if (pa <= 0.0f && pr <= 0.0f) continue;
        const float t = (float)d1 - d1c[kind];
        const float dp = p10 - p00, k2 = 2.0f / (float)is1;
        const float x0 = dp * rcp_approx(p10 - d0f) * k2 * t, x1 = dp * rcp_approx(d0f - p00) * k2 * t;
        if (p10 != d0f) {
            const float a0 = (0.0f < x0) ? x0 + ea : x0 - ea, r0 = (0.0f < x0) ? x0 + er : x0 - er;
            acc[kind * 2] -= __fmaf_rn(pa, rcp_approx(a0), pr * rcp_approx(r0));
        }
        if (p00 != d0f) {
            const float a1 = (0.0f < x1) ? x1 + ea : x1 - ea, r1 = (0.0f < x1) ? x1 + er : x1 - er;
            acc[kind * 2 + 1] -= __fmaf_rn(pa, rcp_approx(a1), pr * rcp_approx(r1));
        }
    }
}

__device__ __forceinline__ void bwd_face_role(const BwdParams& P, const Workspace& ws) {
    const int lane = threadIdx.x & 31;
    if (bwd_overflow(ws)) return;                    // the direct kernel does all of it
    const int nvis = ws.vis_count[0];
    const int* unit_n = unit_region(ws);
    const int nunits = unit_n[nvis] - unit_n[0];
    const bool need_depth = P.use_depth && P.g_depth;
    const bool need_tex = P.use_rgb && P.g_rgb && P.grad_textures;
    const bool need_k6 = P.use_alpha || P.use_rgb;

    // ---- (A) larger faces: unit = one 4-row x 32-column block of the face's bbox, one warp per unit
    for (;;) {
        int un = 0;
        if (lane == 0) un = atomicAdd(&ws.counters[2], 1);
        un = __shfl_sync(FULL, un, 0);
        if (un >= nunits) break;
        const int2 ut = ws.unit_tab[un];
        FaceCtx C;
        face_ctx_load(P, ws, (size_t)ws.vis_ids[ut.x], need_depth, C);
        float S[3] = {0.0f, 0.0f, 0.0f}, G[3] = {0.0f, 0.0f, 0.0f};
        float acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0f;
        const int nbx = (C.bb.z - C.bb.x) / UNIT_W + 1;
        const int yy0 = C.bb.y + (ut.y / nbx) * UNIT_H, xx0 = C.bb.x + (ut.y % nbx) * UNIT_W;
        const int yy = yy0 + (lane >> 3);
        int idv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {           // four independent loads in flight
            const int xx = xx0 + 8 * u + (lane & 7);
            idv[u] = (yy <= C.bb.w && xx <= C.bb.z) ? __ldg(C.fim + (size_t)yy * P.Ws + xx) : -2;
        }
#pragma unroll 1
        for (int u = 0; u < 4; u++)
            if (idv[u] == C.fn) face_sample(P, C, xx0 + 8 * u + (lane & 7), yy, need_depth, need_tex, need_k6, acc, S, G);
#pragma unroll
        for (int k = 0; k < 3; k++) { S[k] = warp_sum(S[k]); G[k] = warp_sum(G[k]); }
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = warp_sum(acc[k]);
        if (lane == 0) {
            float* o = ws.upart + (size_t)un * 18;
#pragma unroll
            for (int k = 0; k < 12; k++) o[k] = acc[k];
            o[12] = S[0]; o[13] = S[1]; o[14] = S[2]; o[15] = G[0]; o[16] = G[1]; o[17] = G[2];
        }
    }

}

__global__ void __launch_bounds__(128) k_bwd_face(BwdParams P, Workspace ws) { bwd_face_role(P, ws); }

// sums the unit partials of every visible face, in unit order (deterministic)
__global__ void __launch_bounds__(256) k_face_sums(Workspace ws) {
    if (bwd_overflow(ws)) return;
    const int nvis = ws.vis_count[0];
    const int* unit_n = unit_region(ws);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int vi = t / 18, k = t % 18;
    if (vi >= nvis) return;
    float a = 0.0f;
    const int ue = unit_n[vi + 1] - unit_n[0];
    if (ue == unit_n[vi] - unit_n[0]) return;           // small face: k_bwd_face wrote its sums directly
    for (int u = unit_n[vi] - unit_n[0]; u < ue; u++) a += ws.upart[(size_t)u * 18 + k];
    const size_t fo = (size_t)ws.vis_ids[vi];
    if (k < 12) ws.gpart[fo * 12 + k] = a;
    else ws.gint[fo * 6 + (k - 12)] = a;
}

// ------------------------------------------------------------------------------------------------------------
// exclusive scan of ws.scan_buf[0 .. n) (n = lines + 1 + 7 * visible faces + 1), three small kernels
// ------------------------------------------------------------------------------------------------------------
constexpr int SCAN_ELEMS = 2048;        // per block of 256 threads

__device__ __forceinline__ int scan_len(const Workspace& ws) { return ws.lines + 1 + ws.vis_count[0] * ITEMS + 1 + ws.vis_count[0] + 1; }

__device__ __forceinline__ int block_exclusive_scan_256(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int w = lane < 8 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int y = __shfl_up_sync(FULL, w, o); if (lane >= o) w += y; }
        if (lane < 8) s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[7];
    const int before = warp > 0 ? s_warp[warp - 1] : 0;
    __syncthreads();
    return before + x - v;
}

// The scan runs either over ws.scan_buf (length known on the device only: n_fixed < 0) or over any int buffer of a
// length the host knows (the per-tile face counts of the forward).
__global__ void __launch_bounds__(256) k_scan_reduce(Workspace ws, int* buf, int* sums, int n_fixed) {
    __shared__ int s_warp[8];
    const int n = n_fixed >= 0 ? n_fixed : scan_len(ws);
    const int base = blockIdx.x * SCAN_ELEMS;
    if (base >= n) return;
    int v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { const int i = base + threadIdx.x * 8 + k; if (i < n) v += buf[i]; }
    int total;
    block_exclusive_scan_256(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_scan_sums(Workspace ws, int* sums, int n_fixed) {
    __shared__ int s_warp[8];
    const int n = n_fixed >= 0 ? n_fixed : scan_len(ws);
    const int nb = (n + SCAN_ELEMS - 1) / SCAN_ELEMS;
    int carry = 0;
    for (int base = 0; base < nb; base += 256) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan_256(v, s_warp, total);
        if (i < nb) sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(256) k_scan_apply(Workspace ws, int* buf, const int* sums, int n_fixed) {
    __shared__ int s_warp[8];
    const int n = n_fixed >= 0 ? n_fixed : scan_len(ws);
    const int base = blockIdx.x * SCAN_ELEMS;
    if (base >= n) return;
    int v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { const int i = base + threadIdx.x * 8 + k; v[k] = i < n ? buf[i] : 0; sum += v[k]; }
    int total;
    int run = sums[blockIdx.x] + block_exclusive_scan_256(sum, s_warp, total);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = base + threadIdx.x * 8 + k;
        if (i < n) buf[i] = run;
        run += v[k];
    }
}

// ------------------------------------------------------------------------------------------------------------
// line sweep: one warp per raster line.  The warp walks its line in passes of 256 pixels (8 per lane, kept in
// registers) and, inside a pass, loops over the line's queries; per (pixel, query) pair the work is ~24 issue slots and
// ONE reciprocal (the four distances of a pair share it).  No atomics; each query's result is written once.
// ------------------------------------------------------------------------------------------------------------
constexpr int LM_ALPHA = 0;     // silhouette only: the positive part does not depend on the query
constexpr int LM_SINGLE = 1;    // one reference pass (rgb, or alpha+rgb summed before the > 0 test)
constexpr int LM_SPLIT = 2;     // alpha pass (eps_a) + rgb pass (eps_r), separate > 0 tests
constexpr int LW = 8;           // warps (= adjacent lines) per CTA
// LP = pixels per lane per pass (a pass = 32*LP pixels of the line, held in registers);
// LQ = queries staged per warp at a time; MINB = CTAs per SM asked of ptxas

__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 relu2(float2 a) { return make_float2(fmaxf(a.x, 0.0f), fmaxf(a.y, 0.0f)); }
__device__ __forceinline__ float2 rcp2(float2 a) { return make_float2(rcp_approx(a.x), rcp_approx(a.y)); }

// Line sweep: one warp per raster line (8 adjacent lines per CTA so that column bundles share 32-byte sectors).
// The warp walks its line in passes of 256 pixels held in registers (two pixels per float2) and, inside a pass, loops
// over the line's queries with packed FP32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100): per pixel pair 3 FFMA2 for
// the colour term, 4 for the distances, 2 reciprocals, and fused accumulates.  Partial sums stay per lane in shared
// memory (conflict-free) and are reduced once per query chunk -- no shuffles or atomics in the hot loop, and every
// query's result is written exactly once, in a fixed order (deterministic).
// a query decoded once per chunk: signed pads and negated colour ready for the packed arithmetic
struct __align__(16) QueryX {
    int lo, hi, flags;
    float xc, k0, k1, nc0, nc1, nc2, e0a, e0r, e1a, e1r;
    float pad_[3];
};

template <int MODE, int LP, int LQ>
__device__ __forceinline__ void line_role(const BwdParams& P, const Workspace& ws, const int block_index) {
    extern __shared__ __align__(16) unsigned char s_line_dyn[];
    float2 (*s_acc)[LQ][33] = reinterpret_cast<float2 (*)[LQ][33]>(s_line_dyn);
    __shared__ QueryX s_q[LW][LQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_img = P.Hs + P.Ws;
    const int bundles_rows = (P.Hs + LW - 1) / LW, bundles_cols = (P.Ws + LW - 1) / LW;
    const int bundles = bundles_rows + bundles_cols;
    const int b = block_index / bundles, bi = block_index % bundles;
    int axis, d0;
    if (bi < bundles_rows) { axis = 1; d0 = bi * LW + warp; if (d0 >= P.Hs) return; }
    else { axis = 0; d0 = (bi - bundles_rows) * LW + warp; if (d0 >= P.Ws) return; }
    const int line = b * per_img + (axis == 0 ? P.Hs : 0) + d0;
    if (bwd_overflow(ws)) return;                       // nothing was deferred (see k_bwd_steps)
    const int qbeg = ws.scan_buf[line], qend = ws.scan_buf[line + 1];
    if (qbeg == qend) return;
    const int is1 = axis == 0 ? P.Hs : P.Ws;
    const int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws;
    const float sc = (P.ss == 2) ? 0.25f : 1.0f;
    float bg[3];
    if (P.bg_b) { bg[0] = P.bg_b[b * 3]; bg[1] = P.bg_b[b * 3 + 1]; bg[2] = P.bg_b[b * 3 + 2]; }
    else { bg[0] = P.bg[0]; bg[1] = P.bg[1]; bg[2] = P.bg[2]; }
    // LM_SINGLE: the rgb pass of a split backward uses eps_r, a genuine single pass has eps_a == eps_r
    const float ea = (MODE == LM_SINGLE && (P.split || !P.use_alpha)) ? (float)P.eps_r : (float)P.eps_a;
    const float er = (float)P.eps_r;
    const size_t plane = (size_t)P.H * P.W;
    const float* ga_p = P.g_alpha ? P.g_alpha + (size_t)b * plane : nullptr;
    const float* gr_p = P.g_rgb ? P.g_rgb + (size_t)b * 3 * plane : nullptr;
    constexpr int LP2 = LP / 2;

    for (int qchunk = qbeg; qchunk < qend; qchunk += LQ) {
        const int qn = min(LQ, qend - qchunk);
        // stage the chunk's queries (lane k <-> query k); union of their ranges tells which passes are needed
        int ulo = 0x7fffffff, uhi = -1;
        for (int k = lane; k < qn; k += 32) {
            const Query q = ws.queries[ws.line_items[qchunk + k]];
            QueryX x;
            x.lo = q.lo; x.hi = q.hi_flags & 0xffff; x.flags = q.hi_flags;
            x.xc = q.xc; x.k0 = q.k0; x.k1 = q.k1; x.nc0 = -q.c0; x.nc1 = -q.c1; x.nc2 = -q.c2;
            x.e0a = (q.hi_flags & QF_POS0) ? ea : -ea; x.e0r = (q.hi_flags & QF_POS0) ? er : -er;
            x.e1a = (q.hi_flags & QF_POS1) ? ea : -ea; x.e1r = (q.hi_flags & QF_POS1) ? er : -er;
            s_q[warp][k] = x;
            ulo = min(ulo, x.lo); uhi = max(uhi, x.hi);
        }
#pragma unroll 4
        for (int k = 0; k < LQ; k++) s_acc[warp][k][lane] = make_float2(0.0f, 0.0f);
        __syncwarp();
        ulo = __reduce_min_sync(FULL, ulo); uhi = __reduce_max_sync(FULL, uhi);

        for (int pbase = (ulo / (32 * LP)) * (32 * LP); pbase <= uhi; pbase += 32 * LP) {
            // ---- pixel records of this pass; position j = pbase + i*32 + lane, pixels (2m, 2m+1) packed in a float2.
            // All loads are issued before anything is consumed (memory-level parallelism: one latency per pass).
            float2 pa[LP2], D[LP2], g0v[LP2], g1v[LP2], g2v[LP2];
            {
                int idxv[LP];
                float gav[LP], r0[LP], r1[LP], r2[LP];
#pragma unroll
                for (int i = 0; i < LP; i++) {
                    const int j = min(pbase + i * 32 + lane, is1 - 1);
                    const int xi = axis == 0 ? d0 : j, yi = axis == 0 ? j : d0;
                    const size_t oi = (size_t)(P.H - 1 - yi / P.ss) * P.W + xi / P.ss;
                    idxv[i] = __ldg(fim + (size_t)yi * P.Ws + xi);
                    gav[i] = (P.use_alpha && ga_p && (MODE != LM_SINGLE || !P.split)) ? __ldg(ga_p + oi) : 0.0f;
                    if (MODE != LM_ALPHA && gr_p) {
                        r0[i] = __ldg(gr_p + oi); r1[i] = __ldg(gr_p + plane + oi); r2[i] = __ldg(gr_p + 2 * plane + oi);
                    } else { r0[i] = 0.0f; r1[i] = 0.0f; r2[i] = 0.0f; }
                }
#pragma unroll
                for (int i = 0; i < LP; i++) {
                    const int j = pbase + i * 32 + lane;
                    float v_pa = 0.0f, v_D = 0.0f, v_g0 = 0.0f, v_g1 = 0.0f, v_g2 = 0.0f;
                    if (j < is1) {
                        const int idx = idxv[i];
                        const float A = idx < 0 ? -(gav[i] * sc) : 0.0f;                   // (alpha - 1) * g
                        if (MODE == LM_ALPHA) { v_pa = fmaxf(A, 0.0f); }
                        else {
                            const int xi = axis == 0 ? d0 : j, yi = axis == 0 ? j : d0;
                            float c[3];
                            colour_at(P, b, (size_t)b * P.Hs * P.Ws + (size_t)yi * P.Ws + xi, idx, bg, c);
                            v_g0 = r0[i] * sc; v_g1 = r1[i] * sc; v_g2 = r2[i] * sc;
                            const float d = __fmaf_rn(c[0], v_g0, __fmaf_rn(c[1], v_g1, c[2] * v_g2));
                            if (MODE == LM_SPLIT) { v_pa = fmaxf(A, 0.0f); v_D = d; }
                            else v_D = A + d;
                        }
                    }
                    if (i & 1) { pa[i / 2].y = v_pa; D[i / 2].y = v_D; g0v[i / 2].y = v_g0; g1v[i / 2].y = v_g1; g2v[i / 2].y = v_g2; }
                    else { pa[i / 2].x = v_pa; D[i / 2].x = v_D; g0v[i / 2].x = v_g0; g1v[i / 2].x = v_g1; g2v[i / 2].x = v_g2; }
                }
            }
            const float xbase = (float)(pbase + lane);
            const int pass_hi = pbase + 32 * LP - 1;
            // ---- queries
            for (int k = 0; k < qn; k++) {
                const int lo = s_q[warp][k].lo, hi = s_q[warp][k].hi;       // broadcast reads
                if (hi < pbase || lo > pass_hi) continue;
                const QueryX q = s_q[warp][k];
                const bool full = (lo <= pbase) && (hi >= pass_hi);        // every pixel of the pass is in range
                const float2 e0a = f2(q.e0a), e1a = f2(q.e1a);
                const float2 k0 = f2(q.k0), k1 = f2(q.k1);
                float2 t = make_float2(xbase - q.xc, (xbase + 32.0f) - q.xc);
                float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
                if (MODE == LM_ALPHA) {
#pragma unroll
                    for (int m = 0; m < LP2; m++) {
                        const float2 da0 = __ffma2_rn(k0, t, e0a), da1 = __ffma2_rn(k1, t, e1a);
                        float2 r = rcp2(__fmul2_rn(da0, da1));
                        const float2 pm = pa[m];
                        if (!full) {        // the RECIPROCAL is zeroed: past the crossing a distance can be exactly 0
                            const int j0 = pbase + (2 * m) * 32 + lane, j1 = j0 + 32;
                            if (j0 < lo || j0 > hi) r.x = 0.0f;
                            if (j1 < lo || j1 > hi) r.y = 0.0f;
                        }
                        const float2 pr_ = __fmul2_rn(pm, r);
                        a0 = __ffma2_rn(pr_, da1, a0);
                        a1 = __ffma2_rn(pr_, da0, a1);
                        t = __fadd2_rn(t, f2(64.0f));
                    }
                } else {
                    const float2 nc0 = f2(q.nc0), nc1 = f2(q.nc1), nc2 = f2(q.nc2);
                    const float2 e0r = f2(q.e0r), e1r = f2(q.e1r);
#pragma unroll
                    for (int m = 0; m < LP2; m++) {
                        float2 dr = __ffma2_rn(nc0, g0v[m], D[m]);
                        dr = __ffma2_rn(nc1, g1v[m], dr);
                        dr = __ffma2_rn(nc2, g2v[m], dr);
                        const float2 pr = relu2(dr);
                        const float2 pm = pa[m];
                        // boundary pass: drop the pixels outside [lo, hi] by zeroing the RECIPROCAL -- past the crossing a
                        // distance can be exactly 0, and 0 * inf would be NaN
                        const int j0 = pbase + (2 * m) * 32 + lane, j1 = j0 + 32;
                        const bool out0 = !full && (j0 < lo || j0 > hi), out1 = !full && (j1 < lo || j1 > hi);
                        if (MODE == LM_SINGLE) {
                            const float2 da0 = __ffma2_rn(k0, t, e0a), da1 = __ffma2_rn(k1, t, e1a);
                            float2 rr = rcp2(__fmul2_rn(da0, da1));
                            if (out0) rr.x = 0.0f;
                            if (out1) rr.y = 0.0f;
                            const float2 r = __fmul2_rn(pr, rr);
                            a0 = __ffma2_rn(r, da1, a0);
                            a1 = __ffma2_rn(r, da0, a1);
                        } else {
                            const float2 da0 = __ffma2_rn(k0, t, e0a), dr0 = __ffma2_rn(k0, t, e0r);
                            const float2 da1 = __ffma2_rn(k1, t, e1a), dr1 = __ffma2_rn(k1, t, e1r);
                            float2 r0 = rcp2(__fmul2_rn(da0, dr0)), r1 = rcp2(__fmul2_rn(da1, dr1));
                            if (out0) { r0.x = 0.0f; r1.x = 0.0f; }
                            if (out1) { r0.y = 0.0f; r1.y = 0.0f; }
                            const float2 n0 = __ffma2_rn(pm, dr0, __fmul2_rn(pr, da0));
                            const float2 n1 = __ffma2_rn(pm, dr1, __fmul2_rn(pr, da1));
                            a0 = __ffma2_rn(n0, r0, a0);
                            a1 = __ffma2_rn(n1, r1, a1);
                        }
                        t = __fadd2_rn(t, f2(64.0f));
                    }
                }
                float2 v = s_acc[warp][k][lane];
                v.x += a0.x + a0.y; v.y += a1.x + a1.y;
                s_acc[warp][k][lane] = v;
            }
        }
        __syncwarp();
        for (int k = lane; k < qn; k += 32) {   // one lane reduces query k over the 32 per-lane partials, in lane order
            float r0 = 0.0f, r1 = 0.0f;
#pragma unroll 8
            for (int l = 0; l < 32; l++) { const float2 v = s_acc[warp][k][l]; r0 += v.x; r1 += v.y; }
            const int hf = s_q[warp][k].flags;
            // g -= diff / dist   (rasterize.py:648,653)
            ws.qres[ws.line_items[qchunk + k]] = make_float2((hf & QF_USE0) ? -r0 : 0.0f, (hf & QF_USE1) ? -r1 : 0.0f);
        }
        __syncwarp();
    }
}


template <int MODE, int LP, int LQ, int MINB>
__global__ void __launch_bounds__(LW * 32, MINB) k_line_sweep(BwdParams P, Workspace ws) {
    line_role<MODE, LP, LQ>(P, ws, blockIdx.x);
}

// The two heavy backward stages are independent of each other (the line sweep needs the queries, the face stage the
// unit table; both exist after k_bwd_steps<EMIT>), one is issue-bound and the other latency-bound: run them as ONE
// launch whose leading CTAs take the face role and the rest the line role, so they overlap on the SMs.
template <int MODE>
__global__ void __launch_bounds__(LW * 32, 2) k_bwd_main(BwdParams P, Workspace ws, int face_ctas) {
    if ((int)blockIdx.x < face_ctas) { bwd_face_role(P, ws); return; }
    line_role<MODE, 8, 32>(P, ws, (int)blockIdx.x - face_ctas);
}

// ------------------------------------------------------------------------------------------------------------
// Alpha (silhouette) out-sweeps by WELL-SEPARATED BLOCK MOMENTS.
//
// For an alpha pass the out-sweep only runs when the in-pixel belongs to the face (rasterize.py:601), so alpha_in == 1
// and the positive part of rasterize.py:628-631, max(0, (alpha_j - 1) * g_j), is a per-pixel weight
//        A_j = background_j ? max(0, -g_j) : 0
// that does NOT depend on the query.  A query (crossing xc, slopes k_v, pads e_v, range = from the crossing to the
// border of the line) needs  R_v = sum_j A_j / (k_v (j - xc) + e_v) = (dir / k_v) sum_j A_j / D_v(j),
// D_v(j) = |j - xc| + |e_v / k_v| > 0.  The line is cut into a binary hierarchy of blocks (16, 32, 64, ... pixels);
// every block keeps the first AM_P scaled moments  M_n = sum_j A_j ((j - c) / h)^n  about its centre c (h = half
// width).  A block whose nearest pixel is at least two block sizes away from the crossing contributes
//        sum_j A_j / (U + dir (j - c)) = (1 / U) sum_n M_n (-dir h / U)^n,      U = D_v(c),   h / U <= 0.2025,
// truncated after AM_P = 8 terms: relative error of that block's (positive) term <= 0.2025^8 / (1 - 0.2025) = 3.6e-6.
// Walking outwards from the crossing, a query therefore sums <= 47 pixels directly (the arithmetic of the direct
// sweep: fma(k, t, e), hardware reciprocal) and then <= ~3 blocks per octave of distance: O(log L) work instead of
// O(L).  One warp per raster line: (1) weights of the line -> shared memory, (2) level-0 moments, one lane per block,
// (3) parents from children (binomial shift), (4) one lane per query.  Every result is written exactly once by one
// lane in a fixed order: deterministic.
// ------------------------------------------------------------------------------------------------------------
constexpr int AM_P = 8;                 // moments per block
constexpr int AM_LOG0 = 4;              // level-0 block: 16 pixels
constexpr int AM_S0 = 1 << AM_LOG0;
constexpr int AM_MAXLEV = 10;           // 16 ... 8192
constexpr int AM_LMAX = 4096;           // longest line the shared-memory layout is sized for (else: direct sweep)

__host__ __device__ constexpr float am_binom(int n, int k) {
    float r = 1.0f;
    for (int i = 1; i <= k; i++) r = r * (float)(n - k + i) / (float)i;
    return r;
}
__host__ __device__ constexpr float am_pow(float x, int n) {
    float r = 1.0f;
    for (int i = 0; i < n; i++) r *= x;
    return r;
}
// number of levels for a line of L pixels: the largest block that can ever be used has 3 s <= L
inline int am_levels(int L) {
    int lv = 1;
    while (lv < AM_MAXLEV && 3 * (AM_S0 << lv) <= L) lv++;
    return lv;
}
inline size_t am_smem_floats(int lcap, int levels) {
    const int nb0 = lcap >> AM_LOG0;
    return (size_t)(2 * nb0 + levels + 1) * AM_P;
}

// weight of sample j of the line: background ? max(0, -g) : 0   (alpha_in == 1, rasterize.py:601,628-631)
__device__ __forceinline__ float am_weight(int idx, float g, float sc) { return idx < 0 ? fmaxf(-(g * sc), 0.0f) : 0.0f; }

__device__ __forceinline__ void line_alpha_role(const BwdParams& P, const Workspace& ws, int lcap, int levels,
                                                const int block_index, float* s_am) {
    __shared__ int s_off[AM_MAXLEV + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb0 = lcap >> AM_LOG0;
    const int mcap = (2 * nb0 + levels + 1) * AM_P;
    float* s_m = s_am + (size_t)warp * mcap;                // moments [level][block][n]
    if (threadIdx.x == 0) {
        int off = 0, nb = nb0;
        for (int l = 0; l < levels; l++) { s_off[l] = off; off += nb * AM_P; nb = (nb + 1) >> 1; }
        s_off[levels] = off;
    }
    __syncthreads();
    const int per_img = P.Hs + P.Ws;
    const int bundles_rows = (P.Hs + LW - 1) / LW, bundles_cols = (P.Ws + LW - 1) / LW;
    const int bundles = bundles_rows + bundles_cols;
    const int b = block_index / bundles, bi = block_index % bundles;
    int axis, d0;
    if (bi < bundles_rows) { axis = 1; d0 = bi * LW + warp; if (d0 >= P.Hs) return; }
    else { axis = 0; d0 = (bi - bundles_rows) * LW + warp; if (d0 >= P.Ws) return; }
    const int line = b * per_img + (axis == 0 ? P.Hs : 0) + d0;
    if (bwd_overflow(ws)) return;
    const int qbeg = ws.scan_buf[line], qend = ws.scan_buf[line + 1];
    if (qbeg == qend) return;
    const int L = axis == 0 ? P.Hs : P.Ws;
    const int nbl0 = (L + AM_S0 - 1) >> AM_LOG0;
    const float sc = (P.ss == 2) ? 0.25f : 1.0f;
    const float ea = (float)P.eps_a;
    // sample j of the line lives at fim[j * fstep] and reads the upstream gradient at ga[(j / ss) * gstep]
    const int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws + (axis == 0 ? (size_t)d0 : (size_t)d0 * P.Ws);
    const int fstep = axis == 0 ? P.Ws : 1;                 // (32-bit offsets: validate() bounds Hs, Ws <= 32767)
    const float* ga = P.g_alpha + (size_t)b * P.H * P.W +
                      (axis == 0 ? (size_t)(P.H - 1) * P.W + d0 / P.ss : (size_t)(P.H - 1 - d0 / P.ss) * P.W);
    const int gstep = axis == 0 ? -P.W : 1;
    const int ssh = P.ss == 2 ? 1 : 0;

    // (1) level-0 moments, one lane per block of 16 samples, straight from global memory (all loads first).
    // Even / odd moments use the mirrored sums and differences: x_i = (i - 7.5) / 8 = -x_(15 - i).
    for (int blk = lane; blk < nbl0; blk += 32) {
        const int j0 = blk << AM_LOG0;
        float w[AM_S0];
        if (axis == 1 && j0 + AM_S0 <= L && (((size_t)(fim + j0)) & 15) == 0) {
            const int4* ip = reinterpret_cast<const int4*>(fim + j0);
            const int4 i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2), i3 = __ldg(ip + 3);
            const int id[16] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w, i3.x, i3.y, i3.z, i3.w};
            if (ssh) {
                const float4* gp = reinterpret_cast<const float4*>(ga + (j0 >> 1));
                float g[8];
                if ((((size_t)(ga + (j0 >> 1))) & 15) == 0) {
                    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
                    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) g[i] = __ldg(ga + (j0 >> 1) + i);
                }
#pragma unroll
                for (int i = 0; i < AM_S0; i++) w[i] = am_weight(id[i], g[i >> 1], sc);
            } else {
                float g[16];
#pragma unroll
                for (int i = 0; i < AM_S0; i++) g[i] = __ldg(ga + j0 + i);
#pragma unroll
                for (int i = 0; i < AM_S0; i++) w[i] = am_weight(id[i], g[i], sc);
            }
        } else {
            int id[AM_S0];
            float g[AM_S0];
#pragma unroll
            for (int i = 0; i < AM_S0; i++) {
                const int j = min(j0 + i, L - 1);
                id[i] = __ldg(fim + j * fstep);
                g[i] = __ldg(ga + (j >> ssh) * gstep);
            }
#pragma unroll
            for (int i = 0; i < AM_S0; i++) w[i] = (j0 + i < L) ? am_weight(id[i], g[i], sc) : 0.0f;
        }
        float m[AM_P];
#pragma unroll
        for (int n = 0; n < AM_P; n++) m[n] = 0.0f;
#pragma unroll
        for (int i = 0; i < AM_S0 / 2; i++) {
            const float sp = w[AM_S0 - 1 - i] + w[i], sm = w[AM_S0 - 1 - i] - w[i];     // x of the upper sample is +|x|
#pragma unroll
            for (int n = 0; n < AM_P; n++) {
                const float xp = am_pow((0.5f * (AM_S0 - 1) - (float)i) / (0.5f * AM_S0), n);
                m[n] = __fmaf_rn((n & 1) ? sm : sp, xp, m[n]);
            }
        }
        float4* o = reinterpret_cast<float4*>(s_m + blk * AM_P);
        o[0] = make_float4(m[0], m[1], m[2], m[3]); o[1] = make_float4(m[4], m[5], m[6], m[7]);
    }
    __syncwarp();
    // (2) parents: x_parent = (y_child -+ 1) / 2  ->  M_n = 2^-n sum_k C(n,k) (-+1)^(n-k) m_k
    {
        int nb = nbl0;
        for (int l = 1; l < levels; l++) {
            const int nbc = nb;
            nb = (nb + 1) >> 1;
            const float* src = s_m + s_off[l - 1];
            float* dst = s_m + s_off[l];
            for (int blk = lane; blk < nb; blk += 32) {
                float cl[AM_P], cr[AM_P];
                const float4* pl = reinterpret_cast<const float4*>(src + (2 * blk) * AM_P);
                const float4 l0 = pl[0], l1 = pl[1];
                cl[0] = l0.x; cl[1] = l0.y; cl[2] = l0.z; cl[3] = l0.w; cl[4] = l1.x; cl[5] = l1.y; cl[6] = l1.z; cl[7] = l1.w;
                if (2 * blk + 1 < nbc) {
                    const float4 r0 = pl[2], r1 = pl[3];
                    cr[0] = r0.x; cr[1] = r0.y; cr[2] = r0.z; cr[3] = r0.w; cr[4] = r1.x; cr[5] = r1.y; cr[6] = r1.z; cr[7] = r1.w;
                } else {
#pragma unroll
                    for (int n = 0; n < AM_P; n++) cr[n] = 0.0f;
                }
                // sums and differences of the two children halve the work: even (n-k) terms take (cr + cl), odd (cr - cl)
                float se[AM_P], so[AM_P];
#pragma unroll
                for (int n = 0; n < AM_P; n++) { se[n] = cr[n] + cl[n]; so[n] = cr[n] - cl[n]; }
                float m[AM_P];
#pragma unroll
                for (int n = 0; n < AM_P; n++) {
                    float a = 0.0f;
#pragma unroll
                    for (int k = 0; k <= n; k++)
                        a = __fmaf_rn(am_binom(n, k) * am_pow(0.5f, n), ((n - k) & 1) ? so[k] : se[k], a);
                    m[n] = a;
                }
                float4* o = reinterpret_cast<float4*>(dst + blk * AM_P);
                o[0] = make_float4(m[0], m[1], m[2], m[3]); o[1] = make_float4(m[4], m[5], m[6], m[7]);
            }
            __syncwarp();
        }
    }
    // (3) one lane per query
    for (int qk = qbeg + lane; qk < qend; qk += 32) {
        const int qi = ws.line_items[qk];
        const Query q = ws.queries[qi];
        const int lo = q.lo, hi = q.hi_flags & 0xffff;
        const bool use0 = (q.hi_flags & QF_USE0) != 0, use1 = (q.hi_flags & QF_USE1) != 0;
        const float e0 = (q.hi_flags & QF_POS0) ? ea : -ea, e1 = (q.hi_flags & QF_POS1) ? ea : -ea;
        const int dir = (hi == L - 1) ? 1 : -1;                    // the sweep always ends at a border of the line
        const float fdir = (float)dir;
        const float p0 = use0 ? fabsf(__fdiv_rn(e0, q.k0)) : 0.0f, p1 = use1 ? fabsf(__fdiv_rn(e1, q.k1)) : 0.0f;
        float near0 = 0.0f, near1 = 0.0f;       // sum A_j / dist_v(j), samples next to the crossing
        float far0 = 0.0f, far1 = 0.0f;         // sum A_j / D_v(j), well-separated blocks
        int pos = dir > 0 ? lo : hi;
        int n = hi - lo + 1;
        // direct samples until the position is block-aligned and two level-0 blocks away: their count is known up front
        {
            const float tt0 = fabsf((float)pos - q.xc);
            int nd = 0;
            for (;;) {      // <= 3 iterations: advance to the first aligned position that is far enough
                const int al = dir > 0 ? pos + nd : pos - nd + 1;
                const int pad = (dir > 0 ? -al : al) & (AM_S0 - 1);
                nd += pad;
                if (tt0 + (float)nd >= (float)(2 * AM_S0) || nd >= n) break;
                nd += AM_S0;
            }
            nd = min(nd, n);
            for (int i0 = 0; i0 < nd; i0 += 8) {        // eight samples per trip: their loads fly together
                int id[8];
                float g[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int j = pos + dir * min(i0 + u, nd - 1);
                    id[u] = __ldg(fim + j * fstep);
                    g[u] = __ldg(ga + (j >> ssh) * gstep);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    // a trip's spare slots repeat the last sample with weight 0 (its distance is never zero)
                    const float w = (i0 + u < nd) ? am_weight(id[u], g[u], sc) : 0.0f;
                    const float tt = (float)(pos + dir * min(i0 + u, nd - 1)) - q.xc;
                    const float d0v = __fmaf_rn(q.k0, tt, e0), d1v = __fmaf_rn(q.k1, tt, e1);
                    const float r = rcp_approx(d0v * d1v) * w;
                    near0 = __fmaf_rn(r, d1v, near0); near1 = __fmaf_rn(r, d0v, near1);
                }
            }
            pos += dir * nd; n -= nd;
        }
        while (n > 0) {                          // blocks: the largest aligned one that is two sizes away
            const float att = fabsf((float)pos - q.xc);
            const int al = dir > 0 ? pos : pos + 1;
            int lv = min(__ffs(al) - 1 - AM_LOG0, ((__float_as_int(att) >> 23) - 127) - (AM_LOG0 + 1));
            lv = min(lv, levels - 1);
            const int s = AM_S0 << lv;
            const int blk = (dir > 0 ? pos : pos - s + 1) >> (AM_LOG0 + lv);
            const float4* mp = reinterpret_cast<const float4*>(s_m + s_off[lv] + blk * AM_P);
            const float4 m0 = mp[0], m1 = mp[1];
            const float half = 0.5f * (float)(s - 1);
            const float U0 = att + half + p0, U1 = att + half + p1;
            const float r = rcp_approx(U0 * U1);
            const float i0 = r * U1, i1 = r * U0;
            const float hz = -fdir * (float)(s >> 1);
            const float z0 = hz * i0, z1 = hz * i1;
            float a0 = m1.w, a1 = m1.w;
            a0 = __fmaf_rn(a0, z0, m1.z); a1 = __fmaf_rn(a1, z1, m1.z);
            a0 = __fmaf_rn(a0, z0, m1.y); a1 = __fmaf_rn(a1, z1, m1.y);
            a0 = __fmaf_rn(a0, z0, m1.x); a1 = __fmaf_rn(a1, z1, m1.x);
            a0 = __fmaf_rn(a0, z0, m0.w); a1 = __fmaf_rn(a1, z1, m0.w);
            a0 = __fmaf_rn(a0, z0, m0.z); a1 = __fmaf_rn(a1, z1, m0.z);
            a0 = __fmaf_rn(a0, z0, m0.y); a1 = __fmaf_rn(a1, z1, m0.y);
            a0 = __fmaf_rn(a0, z0, m0.x); a1 = __fmaf_rn(a1, z1, m0.x);
            far0 = __fmaf_rn(a0, i0, far0); far1 = __fmaf_rn(a1, i1, far1);
            pos += dir * s; n -= s;
        }
        // 1 / dist_v = dir / (k_v D_v);   g -= diff / dist   (rasterize.py:648,653)
        const float r0 = use0 ? -(near0 + fdir * __fdiv_rn(far0, q.k0)) : 0.0f;
        const float r1 = use1 ? -(near1 + fdir * __fdiv_rn(far1, q.k1)) : 0.0f;
        ws.qres[qi] = make_float2(r0, r1);
    }
}

__global__ void __launch_bounds__(LW * 32) k_line_alpha(BwdParams P, Workspace ws, int lcap, int levels) {
    extern __shared__ __align__(16) float s_am_dyn[];
    line_alpha_role(P, ws, lcap, levels, (int)blockIdx.x, s_am_dyn);
}

// ------------------------------------------------------------------------------------------------------------
// Colour (rgb / normal) out-sweeps at 2x SSAA, pairwise but lean.  The positive part of rasterize.py:632-635,
//        max(0, D_j - c_in . g_j),      D_j = c_j . g_j,
// depends on the query through c_in (the colour of the face that owns the in-pixel), so it cannot be tabulated per
// line; what CAN be shared is everything that does not depend on the sample inside an OUTPUT pixel: with 2x SSAA the
// two samples (2o, 2o+1) of an output pixel o on a line read the same upstream gradient g_o, so c_in . g_o is computed
// once per output pixel, and where both samples are background (D equal) the two distances fold into one reciprocal:
//        1/(u - 1/2) + 1/(u + 1/2) = 2u / (u^2 - 1/4),        u = (2o + 1/2 - xc) + e_v / k_v
// (for |u| >= 160 the 1/4 is dropped: relative error <= 1e-5).  Distances are kept in unit-slope form
// dist_v = k_v ((j - xc) + e_v / k_v); the 1 / k_v is applied once per query.  The two vertex distances of a sample
// share one hardware reciprocal, r = 1 / (d0 d1), 1/d0 = r d1, 1/d1 = r d0.  Packed FP32x2 arithmetic throughout
// (two output pixels per instruction).  One warp per raster line, passes of 32 x OP output pixels in registers, the
// line's queries looped inside a pass; per-lane partial sums in shared memory, reduced once per query chunk in lane
// order -- no atomics, every result written once: deterministic.
// ------------------------------------------------------------------------------------------------------------
struct __align__(16) QueryR {
    int lo, hi, flags;
    float xc, s0, s1, nc0, nc1, nc2, ik0, ik1;
    float pad_;
};
constexpr float RGB_FAR = 160.0f;       // samples: beyond this distance 2u/(u^2 - 1/4) -> 2/u

template <int OP, int LQ>
__device__ __forceinline__ void line_rgb_role(const BwdParams& P, const Workspace& ws, const int block_index,
                                              const int accumulate, float2* __restrict__ out, unsigned char* s_line_dyn) {
    float2 (*s_acc)[LQ][33] = reinterpret_cast<float2 (*)[LQ][33]>(s_line_dyn);
    __shared__ QueryR s_q[LW][LQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_img = P.Hs + P.Ws;
    const int bundles_rows = (P.Hs + LW - 1) / LW, bundles_cols = (P.Ws + LW - 1) / LW;
    const int bundles = bundles_rows + bundles_cols;
    const int b = block_index / bundles, bi = block_index % bundles;
    int axis, d0;
    if (bi < bundles_rows) { axis = 1; d0 = bi * LW + warp; if (d0 >= P.Hs) return; }
    else { axis = 0; d0 = (bi - bundles_rows) * LW + warp; if (d0 >= P.Ws) return; }
    const int line = b * per_img + (axis == 0 ? P.Hs : 0) + d0;
    if (bwd_overflow(ws)) return;
    const int qbeg = ws.scan_buf[line], qend = ws.scan_buf[line + 1];
    if (qbeg == qend) return;
    const int L = axis == 0 ? P.Hs : P.Ws;          // samples on the line (even: ss == 2)
    const int Lo = L >> 1;                           // output pixels on the line
    const int32_t* fim = P.face_index + (size_t)b * P.Hs * P.Ws;
    const float sc = 0.25f;
    float bg[3];
    if (P.bg_b) { bg[0] = P.bg_b[b * 3]; bg[1] = P.bg_b[b * 3 + 1]; bg[2] = P.bg_b[b * 3 + 2]; }
    else { bg[0] = P.bg[0]; bg[1] = P.bg[1]; bg[2] = P.bg[2]; }
    const float er = (float)P.eps_r;
    const size_t plane = (size_t)P.H * P.W;
    const float* gr_p = P.g_rgb + (size_t)b * 3 * plane;
    constexpr int OP2 = OP / 2;
    constexpr int PASS_O = 32 * OP;                  // output pixels per pass
    constexpr int PASS_S = 2 * PASS_O;               // samples per pass

    for (int qchunk = qbeg; qchunk < qend; qchunk += LQ) {
        const int qn = min(LQ, qend - qchunk);
        int ulo = 0x7fffffff, uhi = -1;
        for (int k = lane; k < qn; k += 32) {
            const Query q = ws.queries[ws.line_items[qchunk + k]];
            QueryR x;
            x.lo = q.lo; x.hi = q.hi_flags & 0xffff; x.flags = q.hi_flags;
            x.xc = q.xc; x.nc0 = -q.c0; x.nc1 = -q.c1; x.nc2 = -q.c2;
            const bool u0 = (q.hi_flags & QF_USE0) != 0, u1 = (q.hi_flags & QF_USE1) != 0;
            const float e0 = (q.hi_flags & QF_POS0) ? er : -er, e1 = (q.hi_flags & QF_POS1) ? er : -er;
            // unit-slope pads e_v / k_v; an unused vertex borrows the other one's (finite, same sign; result dropped)
            float s0 = u0 ? __fdiv_rn(e0, q.k0) : 0.0f, s1 = u1 ? __fdiv_rn(e1, q.k1) : 0.0f;
            if (!u0) s0 = u1 ? s1 : ((x.hi == L - 1) ? 1.0f : -1.0f);
            if (!u1) s1 = s0;
            x.s0 = s0; x.s1 = s1;
            x.ik0 = u0 ? __fdiv_rn(1.0f, q.k0) : 0.0f; x.ik1 = u1 ? __fdiv_rn(1.0f, q.k1) : 0.0f;
            s_q[warp][k] = x;
            ulo = min(ulo, x.lo); uhi = max(uhi, x.hi);
        }
#pragma unroll 4
        for (int k = 0; k < LQ; k++) s_acc[warp][k][lane] = make_float2(0.0f, 0.0f);
        __syncwarp();
        ulo = __reduce_min_sync(FULL, ulo); uhi = __reduce_max_sync(FULL, uhi);

        for (int sbase = (ulo / PASS_S) * PASS_S; sbase <= uhi; sbase += PASS_S) {
            // ---- records of this pass.  Output pixel o = sbase/2 + i*32 + lane owns samples 2o, 2o+1; pairs (i, i+1)
            // of output pixels are packed in one float2.  All loads are issued before anything is consumed.
            float2 G0[OP2], G1[OP2], G2[OP2], DA[OP2], DB[OP2];
            bool mybg = true;
            {
                int ia[OP], ib[OP];
                float r0[OP], r1[OP], r2[OP];
                const int obase = sbase >> 1;
#pragma unroll
                for (int i = 0; i < OP; i++) {
                    const int o = min(obase + i * 32 + lane, Lo - 1);
                    if (axis == 1) {
                        const int2 v = __ldg(reinterpret_cast<const int2*>(fim + (size_t)d0 * P.Ws + 2 * o));
                        ia[i] = v.x; ib[i] = v.y;
                    } else {
                        ia[i] = __ldg(fim + (size_t)(2 * o) * P.Ws + d0);
                        ib[i] = __ldg(fim + (size_t)(2 * o + 1) * P.Ws + d0);
                    }
                    const size_t oi = axis == 1 ? (size_t)(P.H - 1 - (d0 >> 1)) * P.W + o
                                                : (size_t)(P.H - 1 - o) * P.W + (d0 >> 1);
                    r0[i] = __ldg(gr_p + oi); r1[i] = __ldg(gr_p + plane + oi); r2[i] = __ldg(gr_p + 2 * plane + oi);
                }
#pragma unroll
                for (int i = 0; i < OP; i++) {
                    const int o = obase + i * 32 + lane;
                    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f, dA = 0.0f, dB = 0.0f;
                    if (o < Lo) {
                        g0 = r0[i] * sc; g1 = r1[i] * sc; g2 = r2[i] * sc;
                        const int xa = axis == 1 ? 2 * o : d0, ya = axis == 1 ? d0 : 2 * o;
                        const int xb = axis == 1 ? 2 * o + 1 : d0, yb = axis == 1 ? d0 : 2 * o + 1;
                        float c[3];
                        colour_at(P, b, (size_t)b * P.Hs * P.Ws + (size_t)ya * P.Ws + xa, ia[i], bg, c);
                        dA = __fmaf_rn(c[0], g0, __fmaf_rn(c[1], g1, c[2] * g2));
                        if (ib[i] == ia[i] && P.ts < 2) dB = dA;
                        else {
                            colour_at(P, b, (size_t)b * P.Hs * P.Ws + (size_t)yb * P.Ws + xb, ib[i], bg, c);
                            dB = __fmaf_rn(c[0], g0, __fmaf_rn(c[1], g1, c[2] * g2));
                        }
                        if (ia[i] >= 0 || ib[i] >= 0) mybg = false;
                    }       // beyond the line: g = D = 0, a harmless "background" pair
                    if (i & 1) { G0[i / 2].y = g0; G1[i / 2].y = g1; G2[i / 2].y = g2; DA[i / 2].y = dA; DB[i / 2].y = dB; }
                    else { G0[i / 2].x = g0; G1[i / 2].x = g1; G2[i / 2].x = g2; DA[i / 2].x = dA; DB[i / 2].x = dB; }
                }
            }
            const bool allbg = __all_sync(FULL, mybg);      // every sample of the pass is background (and on the line)
            const float jbase = (float)(sbase + 2 * lane);   // first sample of this lane's first output pixel
            const int pass_hi = sbase + PASS_S - 1;
            // queries of the chunk that reach into this pass: lane k looks at query k, the warp walks the set bits.
            // A range that ends at a border of the LINE is treated as unbounded on that side: samples beyond the line
            // carry g = D = 0 and contribute exactly 0, so the last pass needs no masking.
            static_assert(LQ <= 32, "one ballot bit per staged query");
            unsigned active;
            {
                int qlo = 0x7fffffff, qhi = -1;
                if (lane < qn) { qlo = s_q[warp][lane].lo; qhi = s_q[warp][lane].hi; }
                active = __ballot_sync(FULL, !(qhi < sbase || qlo > pass_hi));
            }
            while (active) {
                const int k = __ffs(active) - 1;
                active &= active - 1;
                const QueryR q = s_q[warp][k];
                const int lo = q.lo == 0 ? -0x40000000 : q.lo, hi = q.hi == L - 1 ? 0x40000000 : q.hi;
                const bool full = (lo <= sbase) && (hi >= pass_hi);
                const float2 nc0 = f2(q.nc0), nc1 = f2(q.nc1), nc2 = f2(q.nc2);
                // signed unit-slope offset of the pair's first sample: (j - xc), elements 64 samples apart
                float2 tt = make_float2(jbase - q.xc, (jbase + 64.0f) - q.xc);
                float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
                if (full && allbg) {
                    // nearest sample of the pass to the crossing
                    const float near = fminf(fabsf((float)sbase - q.xc), fabsf((float)pass_hi - q.xc));
                    float2 u0 = __fadd2_rn(tt, f2(0.5f + q.s0)), u1 = __fadd2_rn(tt, f2(0.5f + q.s1));
                    if (near >= RGB_FAR) {
#pragma unroll
                        for (int m = 0; m < OP2; m++) {
                            float2 x = __fmul2_rn(nc0, G0[m]);
                            x = __ffma2_rn(nc1, G1[m], x);
                            x = __ffma2_rn(nc2, G2[m], x);
                            const float2 v = relu2(__fadd2_rn(DA[m], x));
                            const float2 a = __fmul2_rn(v, rcp2(__fmul2_rn(u0, u1)));
                            a0 = __ffma2_rn(a, u1, a0);
                            a1 = __ffma2_rn(a, u0, a1);
                            u0 = __fadd2_rn(u0, f2(128.0f)); u1 = __fadd2_rn(u1, f2(128.0f));
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < OP2; m++) {
                            float2 x = __fmul2_rn(nc0, G0[m]);
                            x = __ffma2_rn(nc1, G1[m], x);
                            x = __ffma2_rn(nc2, G2[m], x);
                            const float2 v = relu2(__fadd2_rn(DA[m], x));
                            const float2 p0 = __ffma2_rn(u0, u0, f2(-0.25f)), p1 = __ffma2_rn(u1, u1, f2(-0.25f));
                            const float2 a = __fmul2_rn(v, rcp2(__fmul2_rn(p0, p1)));
                            a0 = __ffma2_rn(a, __fmul2_rn(u0, p1), a0);
                            a1 = __ffma2_rn(a, __fmul2_rn(u1, p0), a1);
                            u0 = __fadd2_rn(u0, f2(128.0f)); u1 = __fadd2_rn(u1, f2(128.0f));
                        }
                    }
                    a0 = __fadd2_rn(a0, a0); a1 = __fadd2_rn(a1, a1);       // two samples per output pixel
                } else {
                    const float2 s0 = f2(q.s0), s1 = f2(q.s1);
#pragma unroll
                    for (int m = 0; m < OP2; m++) {
                        // iteration m covers the 128 contiguous samples [sbase + 128 m, +127] across the warp: skip it,
                        // run it unmasked, or (the one span holding the end of the range) run it masked -- warp-uniform
                        const int span_lo = sbase + 128 * m, span_hi = span_lo + 127;
                        if (hi >= span_lo && lo <= span_hi) {
                            float2 x = __fmul2_rn(nc0, G0[m]);
                            x = __ffma2_rn(nc1, G1[m], x);
                            x = __ffma2_rn(nc2, G2[m], x);
                            const float2 pA = relu2(__fadd2_rn(DA[m], x)), pB = relu2(__fadd2_rn(DB[m], x));
                            const float2 dA0 = __fadd2_rn(tt, s0), dA1 = __fadd2_rn(tt, s1);
                            const float2 dB0 = __fadd2_rn(dA0, f2(1.0f)), dB1 = __fadd2_rn(dA1, f2(1.0f));
                            float2 aA = __fmul2_rn(pA, rcp2(__fmul2_rn(dA0, dA1)));
                            float2 aB = __fmul2_rn(pB, rcp2(__fmul2_rn(dB0, dB1)));
                            if (lo > span_lo || hi < span_hi) {
                                // drop the samples outside [lo, hi].  The QUOTIENT is zeroed, not the numerator: on the
                                // far side of the crossing t + s_v can be exactly 0 (0 * inf = NaN)
                                const int jA0 = span_lo + 2 * lane, jA1 = jA0 + 64;
                                if (jA0 < lo || jA0 > hi) aA.x = 0.0f;
                                if (jA0 + 1 < lo || jA0 + 1 > hi) aB.x = 0.0f;
                                if (jA1 < lo || jA1 > hi) aA.y = 0.0f;
                                if (jA1 + 1 < lo || jA1 + 1 > hi) aB.y = 0.0f;
                            }
                            a0 = __ffma2_rn(aA, dA1, a0); a0 = __ffma2_rn(aB, dB1, a0);
                            a1 = __ffma2_rn(aA, dA0, a1); a1 = __ffma2_rn(aB, dB0, a1);
                        }
                        tt = __fadd2_rn(tt, f2(128.0f));
                    }
                }
                float2 v = s_acc[warp][k][lane];
                v.x += a0.x + a0.y; v.y += a1.x + a1.y;
                s_acc[warp][k][lane] = v;
            }
        }
        __syncwarp();
        for (int k = lane; k < qn; k += 32) {   // one lane reduces query k over the 32 per-lane partials, in lane order
            float r0 = 0.0f, r1 = 0.0f;
#pragma unroll 8
            for (int l = 0; l < 32; l++) { const float2 v = s_acc[warp][k][l]; r0 += v.x; r1 += v.y; }
            const QueryR q = s_q[warp][k];
            // dist_v = k_v (t + s_v);   g -= diff / dist   (rasterize.py:648,653)
            float2 o = make_float2((q.flags & QF_USE0) ? -(r0 * q.ik0) : 0.0f, (q.flags & QF_USE1) ? -(r1 * q.ik1) : 0.0f);
            const int qi = ws.line_items[qchunk + k];
            if (accumulate) { const float2 p = out[qi]; o.x += p.x; o.y += p.y; }
            out[qi] = o;
        }
        __syncwarp();
    }
}

template <int OP, int LQ, int MINB>
__global__ void __launch_bounds__(LW * 32, MINB) k_line_rgb(BwdParams P, Workspace ws, int accumulate) {
    extern __shared__ __align__(16) unsigned char s_rgb_dyn[];
    line_rgb_role<OP, LQ>(P, ws, (int)blockIdx.x, accumulate, ws.qres, s_rgb_dyn);
}

// Split backward (alpha pass + colour pass): the two sweeps of a bundle of lines as ONE launch whose even CTAs take the
// colour role and odd CTAs the alpha role.  The colour sweep is issue-bound, the block-moment sweep latency-bound
// (one lane per query, dependent shared-memory reads): co-resident on an SM they fill each other's idle issue slots.
// The colour role writes ws.qres2, the alpha role ws.qres; k_item_sums adds the two.
template <int OP, int LQ>
__global__ void __launch_bounds__(LW * 32, 2) k_line_split(BwdParams P, Workspace ws, int lcap, int levels) {
    extern __shared__ __align__(16) unsigned char s_split_dyn[];
    if (blockIdx.x & 1) line_alpha_role(P, ws, lcap, levels, (int)(blockIdx.x >> 1), reinterpret_cast<float*>(s_split_dyn));
    else line_rgb_role<OP, LQ>(P, ws, (int)(blockIdx.x >> 1), 0, ws.qres2, s_split_dyn);
}

// adds the deferred out-sweep results of every (edge, axis) item onto its in-sweep partial, in step order
__global__ void __launch_bounds__(256) k_item_sums(Workspace ws, int two) {
    if (bwd_overflow(ws)) return;
    const int n = ws.vis_count[0] * ITEMS;
    const int* item_n = ws.scan_buf + ws.lines + 1;
    const int q0 = item_n[0];
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
        const int kind = it % ITEMS;
        if (kind == 6) continue;
        float a0 = 0.0f, a1 = 0.0f;
        const int qe = item_n[it + 1] - q0;
        for (int q = item_n[it] - q0; q < qe; q++) {
            float2 r = ws.qres[q];
            if (two) { const float2 r2 = ws.qres2[q]; r.x += r2.x; r.y += r2.y; }
            a0 += r.x; a1 += r.y;
        }
        float* part = ws.gpart + (size_t)ws.vis_ids[it / ITEMS] * 12 + kind * 2;
        part[0] += a0; part[1] += a1;
    }
}

// deterministic assembly of grad_faces [B,F,9] (+ per-face colour gradient when ts == 1)
__global__ void k_grad_assemble(BwdParams P, Workspace ws, float* __restrict__ grad_faces) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)P.B * P.F) return;
    float g[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    float gc[3] = {0, 0, 0};
    if (ws.visible[i]) {
        float part[12];
#pragma unroll
        for (int k = 0; k < 12; k++) part[k] = ws.gpart[(size_t)i * 12 + k];
        if (P.use_alpha || P.use_rgb) {
#pragma unroll
            for (int v = 0; v < 3; v++) {
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const int axis = 1 - c;
                    // vertex v is pi[0] of edge v and pi[1] of edge (v+2)%3   (rasterize.py:544,648,653)
                    g[3 * v + c] = part[(v * 2 + axis) * 2 + 0] + part[(((v + 2) % 3) * 2 + axis) * 2 + 1];
                }
            }
        }
        const float* in = ws.gint + (size_t)i * 6;
        if (P.use_depth && P.g_depth) {     // rasterize.py:820-834 with the per-face constants factored out
            const float* f = P.faces + (size_t)i * 9;
            const float* inv = ws.face_inv + (size_t)i * 9;
            float tmp[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int l = 0; l < 3; l++) tmp[k] += -inv[3 * l + k] / f[3 * l + 2];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float zk = f[3 * k + 2];
                g[3 * k + 2] += in[k] / (zk * zk);
                g[3 * k + 0] += -tmp[0] * in[k] * (float)P.Ws / 2.0f;
                g[3 * k + 1] += -tmp[1] * in[k] * (float)P.Hs / 2.0f;
            }
        }
        gc[0] = in[3]; gc[1] = in[4]; gc[2] = in[5];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) grad_faces[(size_t)i * 9 + k] = g[k];
    if (P.ts == 1 && P.grad_textures && P.use_rgb) {
        P.grad_textures[(size_t)i * 3 + 0] = gc[0];
        P.grad_textures[(size_t)i * 3 + 1] = gc[1];
        P.grad_textures[(size_t)i * 3 + 2] = gc[2];
    }
}

// ------------------------------------------------------------------------------------------------------------
// faces backward: scatter onto vertices, then chain rule
// ------------------------------------------------------------------------------------------------------------
__global__ void k_faces_backward_scatter(const float* __restrict__ g_ndc, const float* __restrict__ g_nrm,
                                         const float* __restrict__ vertices, const int32_t* __restrict__ faces,
                                         int B, int nv, int nf, int faces_batched, int fill_back, float flip,
                                         float* __restrict__ acc_ndc, float* __restrict__ acc_pre) {
    const int F = fill_back ? 2 * nf : nf;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * F) return;
    const int b = (int)(i / F), fn = (int)(i % F);
    const int src = fn < nf ? fn : fn - nf;
    const int32_t* tri = faces + ((size_t)(faces_batched ? b : 0) * nf + src) * 3;
    int vid[3] = {tri[0], tri[1], tri[2]};
    if (fn >= nf) { int t = vid[0]; vid[0] = vid[2]; vid[2] = t; }
    if (g_ndc) {
        const float* g = g_ndc + (size_t)i * 9;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float* a = acc_ndc + ((size_t)b * nv + vid[k]) * 3;
#pragma unroll
            for (int c = 0; c < 3; c++) { const float v = g[3 * k + c]; if (v != 0.0f) atomicAdd(a + c, v); }
        }
    }
    if (g_nrm) {
        const float gn[3] = {g_nrm[(size_t)i * 3] * flip, g_nrm[(size_t)i * 3 + 1], g_nrm[(size_t)i * 3 + 2]};
        if (gn[0] != 0.0f || gn[1] != 0.0f || gn[2] != 0.0f) {
            float pre[3][3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float* v = vertices + ((size_t)b * nv + vid[k]) * 3;
                pre[k][0] = v[0] * flip; pre[k][1] = v[1]; pre[k][2] = v[2];
            }
            const float a[3] = {pre[0][0] - pre[1][0], pre[0][1] - pre[1][1], pre[0][2] - pre[1][2]};
            const float bb[3] = {pre[2][0] - pre[1][0], pre[2][1] - pre[1][1], pre[2][2] - pre[1][2]};
            const float c[3] = {a[1] * bb[2] - a[2] * bb[1], a[2] * bb[0] - a[0] * bb[2], a[0] * bb[1] - a[1] * bb[0]};
            const float len = sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            const float norm = len + 1e-5f;
            // y = c / (|c| + eps):  gc = gn/norm - c * (gn.c) / (|c| * norm^2)
            const float dot = gn[0] * c[0] + gn[1] * c[1] + gn[2] * c[2];
            const float k2 = len > 0.0f ? dot / (len * norm * norm) : 0.0f;
            const float gc[3] = {gn[0] / norm - c[0] * k2, gn[1] / norm - c[1] * k2, gn[2] / norm - c[2] * k2};
            // c = a x b : ga = b x gc, gb = gc x a      (cross.py:53-57)
            const float ga[3] = {bb[1] * gc[2] - bb[2] * gc[1], bb[2] * gc[0] - bb[0] * gc[2], bb[0] * gc[1] - bb[1] * gc[0]};
            const float gb[3] = {gc[1] * a[2] - gc[2] * a[1], gc[2] * a[0] - gc[0] * a[2], gc[0] * a[1] - gc[1] * a[0]};
            float* p0 = acc_pre + ((size_t)b * nv + vid[0]) * 3;
            float* p1 = acc_pre + ((size_t)b * nv + vid[1]) * 3;
            float* p2 = acc_pre + ((size_t)b * nv + vid[2]) * 3;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicAdd(p0 + k, ga[k]);
                atomicAdd(p2 + k, gb[k]);
                atomicAdd(p1 + k, -(ga[k] + gb[k]));
            }
        }
    }
}

__global__ void k_faces_backward_vertex(const float* __restrict__ vertices, int B, int nv, CamDev cam,
                                        const float* __restrict__ acc_ndc, const float* __restrict__ acc_pre,
                                        float* __restrict__ grad_vertices) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * nv) return;
    const int b = (int)(i / nv);
    const float* v = vertices + (size_t)i * 3;
    // recompute the camera-space position (needed by the quotient rule)
    float x = v[0] * cam.flip, y = v[1], z = v[2];
    if (cam.eye) { x -= cam.eye[b * 3]; y -= cam.eye[b * 3 + 1]; z -= cam.eye[b * 3 + 2]; }
    if (cam.rot) {
        const float* r = cam.rot + b * 9;
        const float ox = x * r[0] + y * r[1] + z * r[2];
        const float oy = x * r[3] + y * r[4] + z * r[5];
        const float oz = x * r[6] + y * r[7] + z * r[8];
        x = ox; y = oy; z = oz;
    }
    float gx = acc_ndc[(size_t)i * 3], gy = acc_ndc[(size_t)i * 3 + 1], gz = acc_ndc[(size_t)i * 3 + 2];
    if (cam.wx) {       // x' = x / z / wx
        const float wx = cam.wx[b], wy = cam.wy[b];
        const float ix = 1.0f / (z * wx), iy = 1.0f / (z * wy);
        const float ngx = gx * ix, ngy = gy * iy;
        gz = gz - (gx * x * ix + gy * y * iy) / z;
        gx = ngx; gy = ngy;
    }
    if (cam.rot) {      // v' = R v  ->  g = R^T g'
        const float* r = cam.rot + b * 9;
        const float ox = r[0] * gx + r[3] * gy + r[6] * gz;
        const float oy = r[1] * gx + r[4] * gy + r[7] * gz;
        const float oz = r[2] * gx + r[5] * gy + r[8] * gz;
        gx = ox; gy = oy; gz = oz;
    }
    gx += acc_pre[(size_t)i * 3]; gy += acc_pre[(size_t)i * 3 + 1]; gz += acc_pre[(size_t)i * 3 + 2];
    grad_vertices[(size_t)i * 3] = gx * cam.flip;
    grad_vertices[(size_t)i * 3 + 1] = gy;
    grad_vertices[(size_t)i * 3 + 2] = gz;
}

// ------------------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------------------
inline int check_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    timing_end();
    if (g_debug_sync && g_dbg_name) {
        const cudaError_t e = cudaStreamSynchronize(g_dbg_stream);
        fprintf(stderr, "[sdn] done   %s -> %s\n", g_dbg_name, cudaGetErrorString(e));
        fflush(stderr);
        g_dbg_name = nullptr;
    }
    return (int)cudaGetLastError();     // also clears a non-sticky launch error, so one bad call does not poison the next
}

Clip make_clip(double near_, double far_) {
    Clip c;
    c.near_ = near_; c.far_ = far_; c.far_f = (float)far_;
    return c;
}

int sm_count() {
    // asked once per device (the answer never changes); -1 = not asked yet
    static std::atomic<int> cached[64];
    static std::once_flag init;
    std::call_once(init, [] { for (auto& c : cached) c.store(-1); });
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n < 0) {
        n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is needed once per (kernel, device), not on every call.  The table is
// keyed by the kernel's ADDRESS (distinct template instantiations share one function-pointer type).
template <typename K>
inline void allow_smem_once(K kernel, size_t bytes) {
    struct Entry { const void* fn; size_t granted[16]; };
    static Entry table[64];
    static int used = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) dev = 0;
    const void* key = (const void*)kernel;
    std::lock_guard<std::mutex> lk(mu);
    Entry* e = nullptr;
    for (int i = 0; i < used; i++) if (table[i].fn == key) { e = &table[i]; break; }
    if (!e && used < 64) { e = &table[used++]; e->fn = key; for (auto& g : e->granted) g = 0; }
    if (e && e->granted[dev] >= bytes && bytes > 0) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e) e->granted[dev] = bytes;
}

CamDev make_cam(const sdn_camera* cam) {
    CamDev c;
    c.rot = cam ? cam->rot : nullptr; c.eye = cam ? cam->eye : nullptr;
    c.wx = cam ? cam->width_x : nullptr; c.wy = cam ? cam->width_y : nullptr;
    c.flip = cam ? cam->flip_x : 1.0f;
    return c;
}

int validate(const sdn_raster* r, bool need_outputs = true) {
    if (!r || r->B <= 0 || r->F <= 0 || r->H <= 0 || r->W <= 0) return SDN_E_ARG;
    if (r->ss != 1 && r->ss != 2) return SDN_E_ARG;
    if ((r->flags & (SDN_RGB | SDN_ALPHA | SDN_DEPTH)) == 0) return SDN_E_NOTHING;
    if (!r->faces || !r->face_index || !r->workspace) return SDN_E_ARG;
    if ((r->flags & SDN_RGB) && (!r->textures || r->texture_size < 1)) return SDN_E_ARG;
    if (need_outputs) {
        if ((r->flags & SDN_RGB) && !r->rgb) return SDN_E_ARG;
        if ((r->flags & SDN_ALPHA) && !r->alpha) return SDN_E_ARG;
        if ((r->flags & SDN_DEPTH) && !r->depth) return SDN_E_ARG;
    }
    if ((long)r->H * r->ss > 32767 || (long)r->W * r->ss > 32767) return SDN_E_ARG;
    if (r->workspace_bytes < sdn_workspace_bytes(r->B, r->F, r->H * r->ss, r->W * r->ss)) return SDN_E_WORKSPACE;
    return 0;
}

int backward_impl(const sdn_raster* r, const sdn_raster_grad* g, int split, double eps_a, double eps_r, void* stream_) {
    int rc = validate(r, false);
    if (rc) return rc;
    if (!g || !g->grad_faces) return SDN_E_ARG;
    if ((r->flags & SDN_RGB) && r->texture_size >= 2 && !r->rgb_ss) return SDN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream_;
    Workspace ws = make_ws(r->workspace, r->B, r->F, r->H * r->ss, r->W * r->ss);
    BwdParams P;
    P.B = r->B; P.F = r->F; P.H = r->H; P.W = r->W; P.ss = r->ss; P.Hs = r->H * r->ss; P.Ws = r->W * r->ss;
    P.flags = r->flags; P.ts = r->texture_size; P.batch_bug = r->batch_bug;
    P.clip = make_clip(r->near, r->far);
    P.eps_tex = r->eps;
    P.use_alpha = (r->flags & SDN_ALPHA) ? 1 : 0; P.use_rgb = (r->flags & SDN_RGB) ? 1 : 0;
    P.use_depth = (r->flags & SDN_DEPTH) ? 1 : 0;
    P.split = split; P.eps_a = eps_a; P.eps_r = eps_r; P.fast_math = g->fast_math;
    P.bg[0] = r->background[0]; P.bg[1] = r->background[1]; P.bg[2] = r->background[2]; P.bg_b = r->background_b;
    P.faces = r->faces; P.textures = r->textures; P.face_index = r->face_index; P.rgb_ss = r->rgb_ss;
    P.g_rgb = P.use_rgb ? g->grad_rgb : nullptr;
    P.g_alpha = P.use_alpha ? g->grad_alpha : nullptr;
    P.g_depth = P.use_depth ? g->grad_depth : nullptr;
    P.grad_textures = P.use_rgb ? g->grad_textures : nullptr;

    cudaError_t e;
    if ((e = cudaMemsetAsync(ws.vis_count, 0, sizeof(int) * 8, st)) != cudaSuccess) return (int)e;
    P.line_path = (P.fast_math && (P.use_alpha || P.use_rgb)) ? 1 : 0;
    if (P.line_path) {
        if ((e = cudaMemsetAsync(ws.scan_buf, 0, sizeof(int) * (ws.lines + 1), st)) != cudaSuccess) return (int)e;
        if ((e = cudaMemsetAsync(ws.line_cursor, 0, sizeof(int) * ws.lines, st)) != cudaSuccess) return (int)e;
    }
    if (P.grad_textures && P.ts >= 2) {
        const size_t nb = (size_t)r->B * r->F * P.ts * P.ts * P.ts * 3 * sizeof(float);
        if ((e = cudaMemsetAsync(P.grad_textures, 0, nb, st)) != cudaSuccess) return (int)e;
    }
    const long BF = (long)r->B * r->F;
    int two_results = 0;            // the line stage left its result in two arrays (qres + qres2)
    timing_begin("k_compact_visible", st);
    k_compact_visible<<<(unsigned)((BF + 255) / 256), 256, 0, st>>>(r->B, r->F, ws);
    if ((rc = check_launch())) return rc;
    const int grid = sm_count() * 4;
    if (P.line_path) {
        const int fgrid = sm_count() * 8;
        timing_begin("k_bwd_steps", st);
        k_bwd_steps<false><<<fgrid, 256, 0, st>>>(P, ws);
        if ((rc = check_launch())) return rc;
        const int sblocks = (ws.scan_n + SCAN_ELEMS - 1) / SCAN_ELEMS;
        timing_begin("k_scan_reduce", st);
        k_scan_reduce<<<sblocks, 256, 0, st>>>(ws, ws.scan_buf, ws.block_sums, -1);
        if ((rc = check_launch())) return rc;
        timing_begin("k_scan_sums", st);
        k_scan_sums<<<1, 256, 0, st>>>(ws, ws.block_sums, -1);
        if ((rc = check_launch())) return rc;
        timing_begin("k_scan_apply", st);
        k_scan_apply<<<sblocks, 256, 0, st>>>(ws, ws.scan_buf, ws.block_sums, -1);
        if ((rc = check_launch())) return rc;
        timing_begin("k_bwd_steps", st);
        k_bwd_steps<true><<<fgrid, 256, 0, st>>>(P, ws);
        if ((rc = check_launch())) return rc;
        static const bool fuse = getenv("SDN_FUSE_BWD") != nullptr;     // measured: no gain (both stages are register-bound)
        if (!fuse) {
            timing_begin("k_bwd_face", st);
            k_bwd_face<<<sm_count() * 8, 128, 0, st>>>(P, ws);
            if ((rc = check_launch())) return rc;
        }
        timing_begin("k_backward_items", st);
        k_backward_items<true><<<grid, BWD_WARPS * 32, 0, st>>>(P, ws, 1);      // no-op unless the queries overflowed
        if ((rc = check_launch())) return rc;
        const int bundles = (P.Hs + LW - 1) / LW + (P.Ws + LW - 1) / LW;
        const unsigned lgrid = (unsigned)(r->B * bundles);
        // kernel variant (registers per pass / staged queries / CTAs per SM); SDN_LINE_VARIANT overrides for tuning
        static const int variant = getenv("SDN_LINE_VARIANT") ? atoi(getenv("SDN_LINE_VARIANT")) : 0;
        const int lmode = (P.use_alpha && !P.use_rgb) ? LM_ALPHA : ((P.use_alpha && P.use_rgb && P.split) ? LM_SPLIT : LM_SINGLE);
#define SDN_LAUNCH_LINE(MODE, LP_, LQ_, MINB_)                                                                          \
    do {                                                                                                                \
        const size_t lsmem = sizeof(float2) * LW * LQ_ * 33;                                                             \
        allow_smem_once(k_line_sweep<MODE, LP_, LQ_, MINB_>, lsmem);                                                     \
        timing_begin("k_line_sweep", st);                                                                                \
        k_line_sweep<MODE, LP_, LQ_, MINB_><<<lgrid, LW * 32, lsmem, st>>>(P, ws);                                       \
    } while (0)
#define SDN_LINE_VARIANTS(MODE)                                                                                          \
    do {                                                                                                                \
        if (variant == 1) SDN_LAUNCH_LINE(MODE, 16, 32, 1);                                                              \
        else if (variant == 2) SDN_LAUNCH_LINE(MODE, 12, 32, 1);                                                         \
        else SDN_LAUNCH_LINE(MODE, 8, 32, 2);                                                                            \
    } while (0)
        if (fuse) {
            const int face_ctas = sm_count() * 4;
            const size_t lsmem = sizeof(float2) * LW * 32 * 33;
#define SDN_LAUNCH_MAIN(MODE)                                                                                           \
    do {                                                                                                                \
        allow_smem_once(k_bwd_main<MODE>, lsmem);                \
        timing_begin("k_bwd_main", st);                                                                                 \
        k_bwd_main<MODE><<<lgrid + face_ctas, LW * 32, lsmem, st>>>(P, ws, face_ctas);                                  \
    } while (0)
            if (lmode == LM_ALPHA) SDN_LAUNCH_MAIN(LM_ALPHA);
            else if (lmode == LM_SPLIT) SDN_LAUNCH_MAIN(LM_SPLIT);
            else SDN_LAUNCH_MAIN(LM_SINGLE);
        } else {
            // SDN_LINE_OLD=1: the round-1 direct sweeps (A/B measurements); default: block moments for the alpha term,
            // the lean paired sweep for the colour term (2x SSAA), the direct sweep for what neither covers
            static const bool old_line = getenv("SDN_LINE_OLD") != nullptr;
            const int lmax = max(P.Hs, P.Ws);
            // Measured on config 2 (B200, ms per batch of 32): silhouette only 0.72 direct -> 0.30 block moments; colour only
            // 0.85 direct -> 0.65 paired sweep; split = the two new kernels back to back.
            const bool want_new = !old_line;
            const bool alpha_moments = want_new && lmax <= AM_LMAX && P.g_alpha && (lmode == LM_ALPHA || lmode == LM_SPLIT);
            const bool rgb_pairs = want_new && P.ss == 2 && P.use_rgb && P.g_rgb &&
                                   (lmode == LM_SPLIT || (lmode == LM_SINGLE && (P.split || !P.use_alpha)));
            if (lmode == LM_SPLIT && alpha_moments != rgb_pairs) {
                SDN_LINE_VARIANTS(LM_SPLIT);                    // one of the two halves is unavailable: direct split sweep
            } else {
                static const int op = getenv("SDN_RGB_OP") ? atoi(getenv("SDN_RGB_OP")) : 8;
                static const bool no_fuse = getenv("SDN_LINE_NO_FUSE") != nullptr;
                const int lcap = (lmax + AM_S0 - 1) / AM_S0 * AM_S0;
                const int levels = am_levels(lmax);
                const size_t asmem = sizeof(float) * LW * am_smem_floats(lcap, levels);
                const size_t lsmem = sizeof(float2) * LW * 32 * 33;
                if (alpha_moments && rgb_pairs && !no_fuse) {
                    const size_t ssmem = asmem > lsmem ? asmem : lsmem;
                    two_results = 1;
                    timing_begin("k_line_split", st);
                    if (op == 4) {
                        allow_smem_once(k_line_split<4, 32>, ssmem);
                        k_line_split<4, 32><<<2 * lgrid, LW * 32, ssmem, st>>>(P, ws, lcap, levels);
                    } else {
                        allow_smem_once(k_line_split<8, 32>, ssmem);
                        k_line_split<8, 32><<<2 * lgrid, LW * 32, ssmem, st>>>(P, ws, lcap, levels);
                    }
                } else {
                    if (alpha_moments) {
                        allow_smem_once(k_line_alpha, asmem);
                        timing_begin("k_line_alpha", st);
                        k_line_alpha<<<lgrid, LW * 32, asmem, st>>>(P, ws, lcap, levels);
                        if (rgb_pairs && (rc = check_launch())) return rc;
                    }
                    if (rgb_pairs) {
                        timing_begin("k_line_rgb", st);
                        if (op == 4) {
                            allow_smem_once(k_line_rgb<4, 32, 2>, lsmem);
                            k_line_rgb<4, 32, 2><<<lgrid, LW * 32, lsmem, st>>>(P, ws, alpha_moments ? 1 : 0);
                        } else if (op == 3) {       // three CTAs per SM: 24 staged queries, <= 85 registers
                            const size_t lsmem3 = sizeof(float2) * LW * 24 * 33;
                            allow_smem_once(k_line_rgb<4, 24, 3>, lsmem3);
                            k_line_rgb<4, 24, 3><<<lgrid, LW * 32, lsmem3, st>>>(P, ws, alpha_moments ? 1 : 0);
                        } else if (op == 5) {
                            const size_t lsmem3 = sizeof(float2) * LW * 24 * 33;
                            allow_smem_once(k_line_rgb<8, 24, 3>, lsmem3);
                            k_line_rgb<8, 24, 3><<<lgrid, LW * 32, lsmem3, st>>>(P, ws, alpha_moments ? 1 : 0);
                        } else {
                            allow_smem_once(k_line_rgb<8, 32, 2>, lsmem);
                            k_line_rgb<8, 32, 2><<<lgrid, LW * 32, lsmem, st>>>(P, ws, alpha_moments ? 1 : 0);
                        }
                    }
                }
                if (!alpha_moments && !rgb_pairs) {
                    if (lmode == LM_ALPHA) SDN_LINE_VARIANTS(LM_ALPHA);
                    else if (lmode == LM_SPLIT) SDN_LINE_VARIANTS(LM_SPLIT);
                    else SDN_LINE_VARIANTS(LM_SINGLE);
                }
            }
        }
        if ((rc = check_launch())) return rc;
        timing_begin("k_face_sums", st);
        k_face_sums<<<(unsigned)((BF * 18 + 255) / 256), 256, 0, st>>>(ws);
    } else if (P.fast_math) {
        timing_begin("k_backward_items", st);
        k_backward_items<true><<<grid, BWD_WARPS * 32, 0, st>>>(P, ws, 0);
    } else {
        timing_begin("k_backward_items", st);
        k_backward_items<false><<<grid, BWD_WARPS * 32, 0, st>>>(P, ws, 0);
    }
    if ((rc = check_launch())) return rc;
    if (P.line_path) {
        timing_begin("k_item_sums", st);
        k_item_sums<<<sm_count() * 8, 256, 0, st>>>(ws, two_results);
        if ((rc = check_launch())) return rc;
    }
    timing_begin("k_grad_assemble", st);
    k_grad_assemble<<<(unsigned)((BF + 255) / 256), 256, 0, st>>>(P, ws, g->grad_faces);
    return check_launch();
}

}  // namespace

// launch accounting shared with the other translation units of this library (sdn_linear_tc.cu); not part of the ABI
namespace sdn_internal {
void launch_begin(const char* name, cudaStream_t st) { timing_begin(name, st); }
int launch_end() { return check_launch(); }
}  // namespace sdn_internal

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int sdn_version(void) { return SDN_VERSION; }

long long sdn_launch_count(void) { return g_launches.load(); }

int sdn_timing_enable(int on) {
    g_timing.store(on ? 1 : 0);
    return 0;
}

// Diagnostics for hangs: names of timed launches that have started but not finished (non-blocking event queries).
int sdn_debug_pending(char* buf, int cap) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    int off = 0, idx = 0;
    for (auto& t : g_timed) {
        const bool started = cudaEventQuery(t.e0) == cudaSuccess, done = cudaEventQuery(t.e1) == cudaSuccess;
        if (!done) {
            const int w = snprintf(buf + off, (size_t)(cap - off > 0 ? cap - off : 0), "#%d %s %s\n", idx, t.name,
                                   started ? "RUNNING" : "queued");
            if (w < 0 || off + w >= cap) break;
            off += w;
        }
        idx++;
    }
    return off;
}

// Writes "kernel_name total_ms launches\n" lines (aggregated since the last read) into buf; returns bytes written.
int sdn_timing_read(char* buf, int cap) {
    std::lock_guard<std::mutex> lk(g_timing_mu);
    struct Agg { std::string name; double ms; int n; };
    std::vector<Agg> agg;
    for (auto& t : g_timed) {
        cudaEventSynchronize(t.e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, t.e0, t.e1);
        cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
        bool found = false;
        for (auto& a : agg) if (a.name == t.name) { a.ms += ms; a.n += 1; found = true; break; }
        if (!found) agg.push_back({t.name, ms, 1});
    }
    g_timed.clear();
    int off = 0;
    for (auto& a : agg) {
        const int w = snprintf(buf ? buf + off : nullptr, buf ? (size_t)(cap - off > 0 ? cap - off : 0) : 0, "%s %.6f %d\n",
                               a.name.c_str(), a.ms, a.n);
        if (w < 0 || off + w >= cap) break;
        off += w;
    }
    return off;
}

const char* sdn_error_string(int code) {
    if (code == 0) return "ok";
    if (code == SDN_E_ARG) return "sdn: invalid argument";
    if (code == SDN_E_WORKSPACE) return "sdn: workspace too small";
    if (code == SDN_E_NOTHING) return "sdn: nothing to draw (no SDN_RGB/SDN_ALPHA/SDN_DEPTH flag)";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "sdn: unknown error";
}

size_t sdn_workspace_bytes(int B, int F, int Hs, int Ws) {
    if (B <= 0 || F <= 0 || Hs <= 0 || Ws <= 0) return 0;
    return make_ws(nullptr, B, F, Hs, Ws).total;
}

int sdn_faces_forward(const float* vertices, const int32_t* faces, int B, int nv, int nf, int faces_batched,
                      int fill_back, const sdn_camera* cam, float* face_ndc, float* face_normal, void* stream) {
    if (!vertices || !faces || !face_ndc || B <= 0 || nv <= 0 || nf <= 0) return SDN_E_ARG;
    if (cam && cam->width_x && !cam->width_y) return SDN_E_ARG;
    const long n = (long)B * (fill_back ? 2 * nf : nf);
    timing_begin("k_faces_forward", (cudaStream_t)stream);
    k_faces_forward<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        vertices, faces, B, nv, nf, faces_batched, fill_back, make_cam(cam), face_ndc, face_normal);
    return check_launch();
}

int sdn_faces_backward(const float* grad_face_ndc, const float* grad_face_normal, const float* vertices,
                       const int32_t* faces, int B, int nv, int nf, int faces_batched, int fill_back,
                       const sdn_camera* cam, float* grad_vertices, void* scratch, size_t scratch_bytes,
                       void* stream) {
    if (!vertices || !faces || !grad_vertices || !scratch || B <= 0 || nv <= 0 || nf <= 0) return SDN_E_ARG;
    const size_t need = (size_t)2 * B * nv * 3 * sizeof(float);
    if (scratch_bytes < need) return SDN_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, need, st);
    if (e != cudaSuccess) return (int)e;
    float* acc_ndc = (float*)scratch;
    float* acc_pre = acc_ndc + (size_t)B * nv * 3;
    const CamDev c = make_cam(cam);
    const long n = (long)B * (fill_back ? 2 * nf : nf);
    int rc;
    if (grad_face_ndc || grad_face_normal) {
        timing_begin("k_faces_backward_scatter", st);
        k_faces_backward_scatter<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            grad_face_ndc, grad_face_normal, vertices, faces, B, nv, nf, faces_batched, fill_back, c.flip, acc_ndc,
            acc_pre);
        if ((rc = check_launch())) return rc;
    }
    const long nvt = (long)B * nv;
    timing_begin("k_faces_backward_vertex", st);
    k_faces_backward_vertex<<<(unsigned)((nvt + 255) / 256), 256, 0, st>>>(vertices, B, nv, c, acc_ndc, acc_pre,
                                                                           grad_vertices);
    return check_launch();
}

int sdn_rasterize_forward(const sdn_raster* r, void* stream_) {
    int rc = validate(r);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream_;
    Workspace ws = make_ws(r->workspace, r->B, r->F, r->H * r->ss, r->W * r->ss);
    cudaError_t e = cudaMemsetAsync(r->workspace, 0, ws.zero_bytes, st);
    if (e != cudaSuccess) return (int)e;
    const int Hs = r->H * r->ss, Wss = r->W * r->ss;
    const long BF = (long)r->B * r->F;
    timing_begin("k_face_setup", st);
    k_face_setup<<<(unsigned)((BF + 255) / 256), 256, 0, st>>>(r->faces, r->B, r->F, Hs, Wss, ws);
    if ((rc = check_launch())) return rc;

    FwdParams P;
    P.B = r->B; P.F = r->F; P.H = r->H; P.W = r->W; P.ss = r->ss; P.Hs = Hs; P.Ws = Wss;
    P.flags = r->flags; P.ts = r->texture_size; P.batch_bug = r->batch_bug;
    P.clip = make_clip(r->near, r->far);
    P.eps = r->eps;
    P.bg[0] = r->background[0]; P.bg[1] = r->background[1]; P.bg[2] = r->background[2]; P.bg_b = r->background_b;
    P.faces = r->faces; P.textures = r->textures;
    P.rgb = r->rgb; P.alpha = r->alpha; P.depth = r->depth; P.face_index = r->face_index; P.rgb_ss = r->rgb_ss;
    P.tiles_x = (Wss + TW - 1) / TW; P.tiles_y = (Hs + TH - 1) / TH;
    P.num_tiles = P.tiles_x * P.tiles_y * r->B;
    {   // faces -> tiles (count, scan, fill); skipped work when it does not fit is handled inside the kernels
        const dim3 bgrid((unsigned)((ws.Fp + 255) / 256), (unsigned)r->B);
        const int tn = ws.num_tiles + 1, tblocks = (tn + SCAN_ELEMS - 1) / SCAN_ELEMS;
        timing_begin("k_bin_faces", st);
        k_bin_faces<false><<<bgrid, 256, 0, st>>>(r->B, P.tiles_x, P.tiles_y, ws);
        if ((rc = check_launch())) return rc;
        timing_begin("k_pad_tile_counts", st);
        k_pad_tile_counts<<<(unsigned)((ws.num_tiles + 255) / 256), 256, 0, st>>>(ws);
        if ((rc = check_launch())) return rc;
        timing_begin("k_scan_reduce", st);
        k_scan_reduce<<<tblocks, 256, 0, st>>>(ws, ws.tile_off, ws.bin_sums, tn);
        if ((rc = check_launch())) return rc;
        timing_begin("k_scan_sums", st);
        k_scan_sums<<<1, 256, 0, st>>>(ws, ws.bin_sums, tn);
        if ((rc = check_launch())) return rc;
        timing_begin("k_scan_apply", st);
        k_scan_apply<<<tblocks, 256, 0, st>>>(ws, ws.tile_off, ws.bin_sums, tn);
        if ((rc = check_launch())) return rc;
        timing_begin("k_bin_faces", st);
        k_bin_faces<true><<<bgrid, 256, 0, st>>>(r->B, P.tiles_x, P.tiles_y, ws);
        if ((rc = check_launch())) return rc;
    }
    timing_begin("k_fill_background", st);
    k_fill_background<<<sm_count() * 8, 256, 0, st>>>(P);
    if ((rc = check_launch())) return rc;
    timing_begin("k_tile_prefix", st);
    k_tile_prefix<<<1, 256, 0, st>>>(r->B, ws);
    if ((rc = check_launch())) return rc;
    const int grid = min(P.num_tiles, sm_count() * 5);
    if (r->ss == 2) { timing_begin("k_raster_tiles", st); k_raster_tiles<2><<<grid, NT, 0, st>>>(P, ws); }
    else { timing_begin("k_raster_tiles", st); k_raster_tiles<1><<<grid, NT, 0, st>>>(P, ws); }
    return check_launch();
}

int sdn_rasterize_backward(const sdn_raster* r, const sdn_raster_grad* g, void* stream) {
    return backward_impl(r, g, 0, r ? r->eps : 0.0, r ? r->eps : 0.0, stream);
}

int sdn_rasterize_backward_split(const sdn_raster* r, const sdn_raster_grad* g, double eps_alpha, double eps_rgb,
                                 void* stream) {
    return backward_impl(r, g, 1, eps_alpha, eps_rgb, stream);
}

}  // extern "C"
