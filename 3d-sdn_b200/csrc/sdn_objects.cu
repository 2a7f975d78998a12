// sdn_objects.cu -- the pre-raster stage of the geometric branch for a RAGGED batch of objects in one launch:
// free-form deformation of each object's template mesh followed by the camera-space transform.
//
// Reference (one object at a time, ~40 PyTorch kernels and a 4.8 MB upload per object):
//   derender3d/models/__init__.py:161-200   the per-object loop of Derenderer3d.render
//   derender3d/models/transforms.py:68-99   FFD.forward:  V = sum_ijk (P0 + dP)[c,i,j,k] * B[v,i,j,k]
//   derender3d/models/transforms.py:102-158 PerspectiveTransform.forward: scale, quaternion rotation, translation,
//                                           shear along the principal ray, zoom (test mode: zoom derived from the
//                                           vertices, :148-149)
// Here: object i carries a template-class id; the templates' Bernstein tables live concatenated in one device array
// (the "mesh bank"); the output is [N, nv_max, 3], (0,0,1) beyond each object's own vertex count (unreferenced by its
// faces; z = 1 keeps the perspective divide of the face kernel's vertex pass finite), ready for one
// sdn_faces_forward / sdn_rasterize_forward call over the whole batch.  The constraint step of the FFD
// (transforms.py:71-95) is linear in the coefficients and stays on the caller's side; `control` here is P0 + dP.
//
// Arithmetic: float32, the reference's operation order per vertex, no contraction (-fmad=false); only the 64-term
// FFD sum is ordered differently from torch's reduction (the tests state 1e-5 rel-L2).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sdn_b200.h"

namespace sdn_internal {
void launch_begin(const char* name, cudaStream_t st);
int launch_end();
}  // namespace sdn_internal

namespace {

constexpr int TV = 128;             // vertices per block
constexpr int NB = 64;              // Bernstein terms per vertex (4 x 4 x 4 lattice)
constexpr int NACC = 24;            // per-object scalar accumulators of the backward (see acc layout below)
// acc layout: [0..2] g_scales, [3..11] g_T (row-major 3x3), [12..14] g_translations, [15] g_kx, [16] g_ky, [17] g_zoom

struct ObjParams {
    const float* basis;             // [sum nv, 64]
    const long long* basis_off;     // [K] first row of class k
    const int* nv;                  // [K]
    const long long* classes;       // [N]
    const float* control;           // [N,3,64]
    const float* scales;            // [N,3]
    const float* rotations;         // [N,4] unit quaternion (a,b,c,d)
    const float* translations;      // [N,3]
    const float* ptrans;            // [N,3] principal-ray point
    const float* zooms;             // [N] (given: training mode; derived: written by the forward in test mode)
    int N, K, nv_max;
};

struct ObjCtx {                     // per-object constants, same expressions as transforms.py:107-126
    float s[3], T[9], t[3], kx, ky;
};

__device__ __forceinline__ void load_ctx(const ObjParams& P, int i, ObjCtx& c) {
#pragma unroll
    for (int k = 0; k < 3; k++) { c.s[k] = P.scales[i * 3 + k]; c.t[k] = P.translations[i * 3 + k]; }
    const float a = P.rotations[i * 4], b = P.rotations[i * 4 + 1], cc = P.rotations[i * 4 + 2], d = P.rotations[i * 4 + 3];
    c.T[0] = a * a + b * b - cc * cc - d * d; c.T[1] = 2 * b * cc - 2 * a * d;          c.T[2] = 2 * b * d + 2 * a * cc;
    c.T[3] = 2 * b * cc + 2 * a * d;          c.T[4] = a * a - b * b + cc * cc - d * d; c.T[5] = 2 * cc * d - 2 * a * b;
    c.T[6] = 2 * b * d - 2 * a * cc;          c.T[7] = 2 * cc * d + 2 * a * b;          c.T[8] = a * a - b * b - cc * cc + d * d;
    const float x0 = P.ptrans[i * 3], y0 = P.ptrans[i * 3 + 1], z0 = P.ptrans[i * 3 + 2];
    c.kx = __fdiv_rn(x0, z0);
    c.ky = __fdiv_rn(y0, z0);
}

// stage TV rows of the class's Bernstein table (coalesced) and the object's control points in shared memory
__device__ __forceinline__ void stage(const ObjParams& P, int i, int k, int v0, int nv, float (*sB)[NB + 1],
                                      float* sP) {
    const float* src = P.basis + (size_t)(P.basis_off[k] + v0) * NB;
    const int rows = min(TV, nv - v0);
    for (int e = threadIdx.x; e < rows * NB; e += blockDim.x) sB[e / NB][e % NB] = src[e];
    for (int e = threadIdx.x; e < 3 * NB; e += blockDim.x) sP[e] = P.control[(size_t)i * 3 * NB + e];
    __syncthreads();
}

__device__ __forceinline__ void ffd_vertex(const float* row, const float* sP, float V[3]) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll 8
    for (int j = 0; j < NB; j++) {
        const float bj = row[j];
        a0 += sP[j] * bj; a1 += sP[NB + j] * bj; a2 += sP[2 * NB + j] * bj;
    }
    V[0] = a0; V[1] = a1; V[2] = a2;
}

// scale, rotate, translate, shear: everything of PerspectiveTransform.forward up to the zoom division
__device__ __forceinline__ void pre_zoom(const ObjCtx& c, const float V[3], float u[3], float out[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) u[k] = V[k] * c.s[k];
    float w[3];
#pragma unroll
    for (int r = 0; r < 3; r++) w[r] = (u[0] * c.T[3 * r] + u[1] * c.T[3 * r + 1]) + u[2] * c.T[3 * r + 2];
    const float tx = w[0] + c.t[0], ty = w[1] + c.t[1], tz = w[2] + c.t[2];
    out[0] = tx - c.kx * tz;
    out[1] = ty - c.ky * tz;
    out[2] = tz;
}

// ---- forward --------------------------------------------------------------------------------------------------
// derive_zoom: test mode -- vertices are written un-zoomed and the per-object minimum of |z| / max(|x|,|y|) is taken
// with a packed (ratio bits << 32 | vertex) atomicMin (ratio >= 0, so its bit pattern orders like the float).
template <bool DERIVE>
__global__ void __launch_bounds__(TV) k_object_vertices(ObjParams P, float* __restrict__ vertices,
                                                        unsigned long long* __restrict__ zoom_key) {
    __shared__ float sB[TV][NB + 1];
    __shared__ float sP[3 * NB];
    const int i = blockIdx.y, v0 = blockIdx.x * TV, v = v0 + threadIdx.x;
    const int k = (int)P.classes[i];
    const int nv = (k >= 0 && k < P.K) ? P.nv[k] : 0;
    float* o = vertices + ((size_t)i * P.nv_max + v) * 3;
    if (v0 >= nv) {                         // whole block is padding
        if (v < P.nv_max) { o[0] = 0.0f; o[1] = 0.0f; o[2] = 1.0f; }
        return;
    }
    stage(P, i, k, v0, nv, sB, sP);
    ObjCtx c;
    load_ctx(P, i, c);
    unsigned long long key = ~0ull;
    if (v < nv) {
        float V[3], u[3], q[3];
        ffd_vertex(sB[threadIdx.x], sP, V);
        pre_zoom(c, V, u, q);
        if (DERIVE) {
            const float r = __fdiv_rn(fabsf(q[2]), fmaxf(fabsf(q[0]), fabsf(q[1])));
            if (r == r) key = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned)v;     // NaN never wins
        } else {
            q[2] = __fdiv_rn(q[2], P.zooms[i]);
        }
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
    } else if (v < P.nv_max) {
        o[0] = 0.0f; o[1] = 0.0f; o[2] = 1.0f;
    }
    if (DERIVE) {
        for (int s = 16; s > 0; s >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, s);
            key = other < key ? other : key;
        }
        if ((threadIdx.x & 31) == 0 && key != ~0ull) atomicMin(zoom_key + i, key);
    }
}

// test mode, second pass: zoom_i = min ratio * zoom_to_i (transforms.py:148-149), z /= zoom
__global__ void k_object_apply_zoom(ObjParams P, const float* __restrict__ zoom_tos,
                                    const unsigned long long* __restrict__ zoom_key, float* __restrict__ zooms_out,
                                    float* __restrict__ vertices) {
    const int i = blockIdx.y, v = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = (int)P.classes[i];
    const int nv = (k >= 0 && k < P.K) ? P.nv[k] : 0;
    const float zoom = __uint_as_float((unsigned)(zoom_key[i] >> 32)) * zoom_tos[i];
    if (v == 0) zooms_out[i] = zoom;
    if (v < nv) {
        float* o = vertices + ((size_t)i * P.nv_max + v) * 3;
        o[2] = __fdiv_rn(o[2], zoom);
    }
}

// ---- backward -------------------------------------------------------------------------------------------------
// Test mode, pass 1: g_zoom_i = sum_v g_z * (-z'/zoom^2) is needed before the per-vertex pass, because the zoom is a
// function of one vertex (the argmin) whose own chain then receives it.
__global__ void __launch_bounds__(TV) k_object_zoom_grad(ObjParams P, const float* __restrict__ grad_vertices,
                                                         float* __restrict__ acc) {
    __shared__ float sB[TV][NB + 1];
    __shared__ float sP[3 * NB];
    const int i = blockIdx.y, v0 = blockIdx.x * TV, v = v0 + threadIdx.x;
    const int k = (int)P.classes[i];
    const int nv = (k >= 0 && k < P.K) ? P.nv[k] : 0;
    if (v0 >= nv) return;
    stage(P, i, k, v0, nv, sB, sP);
    ObjCtx c;
    load_ctx(P, i, c);
    float gz = 0.0f;
    if (v < nv) {
        float V[3], u[3], q[3];
        ffd_vertex(sB[threadIdx.x], sP, V);
        pre_zoom(c, V, u, q);
        const float zoom = P.zooms[i];
        gz = grad_vertices[((size_t)i * P.nv_max + v) * 3 + 2] * (-q[2] / (zoom * zoom));
    }
    for (int s = 16; s > 0; s >>= 1) gz += __shfl_xor_sync(0xffffffffu, gz, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc + (size_t)i * NACC + 17, gz);
}

// DERIVE: test mode (zoom derived from vertex argmin; acc[17] already holds g_zoom from pass 1)
template <bool DERIVE>
__global__ void __launch_bounds__(TV) k_object_vertices_bwd(ObjParams P, const float* __restrict__ grad_vertices,
                                                            const float* __restrict__ zoom_tos,
                                                            const unsigned long long* __restrict__ zoom_key,
                                                            float* __restrict__ acc, float* __restrict__ grad_control) {
    __shared__ float sB[TV][NB + 1];
    __shared__ float sP[3 * NB];
    __shared__ float sG[TV][3];
    __shared__ float sRed[TV / 32][NACC];
    const int i = blockIdx.y, v0 = blockIdx.x * TV, v = v0 + threadIdx.x;
    const int k = (int)P.classes[i];
    const int nv = (k >= 0 && k < P.K) ? P.nv[k] : 0;
    if (v0 >= nv) return;
    stage(P, i, k, v0, nv, sB, sP);
    ObjCtx c;
    load_ctx(P, i, c);
    const float zoom = P.zooms[i];
    float a[18];
#pragma unroll
    for (int n = 0; n < 18; n++) a[n] = 0.0f;
    float gV[3] = {0.0f, 0.0f, 0.0f};
    if (v < nv) {
        float V[3], u[3], q[3];
        ffd_vertex(sB[threadIdx.x], sP, V);
        pre_zoom(c, V, u, q);
        const float* g = grad_vertices + ((size_t)i * P.nv_max + v) * 3;
        float gx = g[0], gy = g[1], gzp = g[2] / zoom;                    // gradient on the un-zoomed z'
        if (DERIVE) {
            if ((unsigned)(zoom_key[i] & 0xffffffffu) == (unsigned)v) {    // the argmin vertex carries d zoom
                const float g_r = acc[(size_t)i * NACC + 17] * zoom_tos[i];
                const float ax = fabsf(q[0]), ay = fabsf(q[1]), az = fabsf(q[2]);
                const float m = fmaxf(ax, ay);
                const float sz = (q[2] > 0.0f) - (q[2] < 0.0f);
                gzp += g_r * sz / m;
                const float gm = -g_r * az / (m * m);
                if (ax >= ay) gx += gm * ((q[0] > 0.0f) - (q[0] < 0.0f));
                else gy += gm * ((q[1] > 0.0f) - (q[1] < 0.0f));
            }
        } else {
            a[17] = g[2] * (-q[2] / (zoom * zoom));
        }
        const float gw[3] = {gx, gy, gzp - c.kx * gx - c.ky * gy};        // gradient on (tx, ty, tz)
        a[15] = -q[2] * gx; a[16] = -q[2] * gy;                            // d / d kx, ky  (tz == q[2])
#pragma unroll
        for (int r = 0; r < 3; r++) {
            a[12 + r] = gw[r];
#pragma unroll
            for (int cc = 0; cc < 3; cc++) a[3 + 3 * r + cc] = gw[r] * u[cc];
        }
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
            const float gu = (c.T[cc] * gw[0] + c.T[3 + cc] * gw[1]) + c.T[6 + cc] * gw[2];
            a[cc] = gu * V[cc];
            gV[cc] = gu * c.s[cc];
        }
    }
    sG[threadIdx.x][0] = gV[0]; sG[threadIdx.x][1] = gV[1]; sG[threadIdx.x][2] = gV[2];
    // per-object scalars: warp shuffle, then one atomic per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < 18; n++) {
        float x = a[n];
        for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
        if (lane == 0) sRed[warp][n] = x;
    }
    __syncthreads();
    if (threadIdx.x < 18 && !(DERIVE && threadIdx.x == 17)) {
        float x = 0.0f;
        for (int w = 0; w < TV / 32; w++) x += sRed[w][threadIdx.x];
        atomicAdd(acc + (size_t)i * NACC + threadIdx.x, x);
    }
    // g_control[c][j] += sum_v gV[v][c] * B[v][j]
    const int rows = min(TV, nv - v0);
    for (int o = threadIdx.x; o < 3 * NB; o += TV) {
        const int cc = o / NB, j = o % NB;
        float x = 0.0f;
        for (int r = 0; r < rows; r++) x += sG[r][cc] * sB[r][j];
        atomicAdd(grad_control + (size_t)i * 3 * NB + o, x);
    }
}

// per-object finalisation: accumulators -> gradients of scales, quaternion, translations, principal-ray point, zoom
__global__ void k_object_finalize(ObjParams P, const float* __restrict__ acc, const float* __restrict__ zoom_tos,
                                  const unsigned long long* __restrict__ zoom_key, int derive,
                                  float* __restrict__ g_scales, float* __restrict__ g_rot, float* __restrict__ g_trans,
                                  float* __restrict__ g_ptrans, float* __restrict__ g_zoom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const float* A = acc + (size_t)i * NACC;
#pragma unroll
    for (int k = 0; k < 3; k++) { g_scales[i * 3 + k] = A[k]; g_trans[i * 3 + k] = A[12 + k]; }
    const float a = P.rotations[i * 4], b = P.rotations[i * 4 + 1], c = P.rotations[i * 4 + 2], d = P.rotations[i * 4 + 3];
    const float* G = A + 3;     // dL/dT, row major
    g_rot[i * 4 + 0] = 2 * (a * (G[0] + G[4] + G[8]) + d * (G[3] - G[1]) + c * (G[2] - G[6]) + b * (G[7] - G[5]));
    g_rot[i * 4 + 1] = 2 * (b * (G[0] - G[4] - G[8]) + c * (G[1] + G[3]) + d * (G[2] + G[6]) + a * (G[7] - G[5]));
    g_rot[i * 4 + 2] = 2 * (c * (G[4] - G[0] - G[8]) + b * (G[1] + G[3]) + a * (G[2] - G[6]) + d * (G[5] + G[7]));
    g_rot[i * 4 + 3] = 2 * (d * (G[8] - G[0] - G[4]) + a * (G[3] - G[1]) + b * (G[2] + G[6]) + c * (G[5] + G[7]));
    const float x0 = P.ptrans[i * 3], y0 = P.ptrans[i * 3 + 1], z0 = P.ptrans[i * 3 + 2];
    g_ptrans[i * 3 + 0] = A[15] / z0;
    g_ptrans[i * 3 + 1] = A[16] / z0;
    g_ptrans[i * 3 + 2] = -(A[15] * x0 + A[16] * y0) / (z0 * z0);
    // training: gradient of the given zoom; test: gradient of zoom_to (zoom = ratio * zoom_to)
    g_zoom[i] = derive ? A[17] * __uint_as_float((unsigned)(zoom_key[i] >> 32)) : A[17];
}

int check_params(const sdn_objects* o) {
    if (!o || o->N <= 0 || o->K <= 0 || o->nv_max <= 0) return SDN_E_ARG;
    if (!o->basis || !o->basis_off || !o->nv || !o->classes || !o->control || !o->scales || !o->rotations ||
        !o->translations || !o->ptrans || !o->vertices || !o->zooms)
        return SDN_E_ARG;
    if (o->derive_zoom && (!o->zoom_tos || !o->zoom_key)) return SDN_E_ARG;
    return 0;
}

ObjParams to_params(const sdn_objects* o) {
    ObjParams P;
    P.basis = o->basis; P.basis_off = o->basis_off; P.nv = o->nv; P.classes = o->classes; P.control = o->control;
    P.scales = o->scales; P.rotations = o->rotations; P.translations = o->translations; P.ptrans = o->ptrans;
    P.zooms = o->zooms; P.N = o->N; P.K = o->K; P.nv_max = o->nv_max;
    return P;
}

}  // namespace

extern "C" int sdn_objects_forward(const sdn_objects* o, void* stream) {
    int rc = check_params(o);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const ObjParams P = to_params(o);
    const dim3 grid((unsigned)((o->nv_max + TV - 1) / TV), (unsigned)o->N);
    if (o->derive_zoom) {
        cudaError_t e = cudaMemsetAsync(o->zoom_key, 0xff, sizeof(unsigned long long) * o->N, st);
        if (e != cudaSuccess) return (int)e;
        sdn_internal::launch_begin("k_object_vertices", st);
        k_object_vertices<true><<<grid, TV, 0, st>>>(P, o->vertices, o->zoom_key);
        if ((rc = sdn_internal::launch_end())) return rc;
        sdn_internal::launch_begin("k_object_apply_zoom", st);
        k_object_apply_zoom<<<grid, TV, 0, st>>>(P, o->zoom_tos, o->zoom_key, o->zooms, o->vertices);
        return sdn_internal::launch_end();
    }
    sdn_internal::launch_begin("k_object_vertices", st);
    k_object_vertices<false><<<grid, TV, 0, st>>>(P, o->vertices, nullptr);
    return sdn_internal::launch_end();
}

extern "C" int sdn_objects_backward(const sdn_objects* o, const float* grad_vertices, float* grad_control,
                                    float* grad_scales, float* grad_rotations, float* grad_translations,
                                    float* grad_ptrans, float* grad_zoom, float* scratch, size_t scratch_bytes,
                                    void* stream) {
    int rc = check_params(o);
    if (rc) return rc;
    if (!grad_vertices || !grad_control || !grad_scales || !grad_rotations || !grad_translations || !grad_ptrans ||
        !grad_zoom || !scratch)
        return SDN_E_ARG;
    if (scratch_bytes < sizeof(float) * NACC * (size_t)o->N) return SDN_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const ObjParams P = to_params(o);
    const dim3 grid((unsigned)((o->nv_max + TV - 1) / TV), (unsigned)o->N);
    cudaError_t e;
    if ((e = cudaMemsetAsync(scratch, 0, sizeof(float) * NACC * (size_t)o->N, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(grad_control, 0, sizeof(float) * 3 * NB * (size_t)o->N, st)) != cudaSuccess) return (int)e;
    if (o->derive_zoom) {
        sdn_internal::launch_begin("k_object_zoom_grad", st);
        k_object_zoom_grad<<<grid, TV, 0, st>>>(P, grad_vertices, scratch);
        if ((rc = sdn_internal::launch_end())) return rc;
        sdn_internal::launch_begin("k_object_vertices_bwd", st);
        k_object_vertices_bwd<true><<<grid, TV, 0, st>>>(P, grad_vertices, o->zoom_tos, o->zoom_key, scratch, grad_control);
    } else {
        sdn_internal::launch_begin("k_object_vertices_bwd", st);
        k_object_vertices_bwd<false><<<grid, TV, 0, st>>>(P, grad_vertices, nullptr, nullptr, scratch, grad_control);
    }
    if ((rc = sdn_internal::launch_end())) return rc;
    sdn_internal::launch_begin("k_object_finalize", st);
    k_object_finalize<<<(o->N + 127) / 128, 128, 0, st>>>(P, scratch, o->zoom_tos, o->zoom_key, o->derive_zoom,
                                                            grad_scales, grad_rotations, grad_translations,
                                                            grad_ptrans, grad_zoom);
    return sdn_internal::launch_end();
}
