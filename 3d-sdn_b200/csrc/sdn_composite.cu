// sdn_composite.cu -- the test-time scene compositor: every object's rendered silhouette / normal / depth crop
// (render_size^2) is resized to its on-screen size and pasted, far to near, into the full-frame instance, normal and
// depth maps.  Reference: geometric/scripts/main.py:547-606, one object at a time through
// torchvision.to_pil_image / resize / PIL.Image.paste / to_tensor on the CPU (GPU -> CPU -> GPU per object).
// Here: two launches for the whole scene.  The arithmetic is Pillow's bilinear `Image.resize` (libImaging/Resample.c),
// restated so that the result is bit-identical: 8-bit planes in 22-bit fixed point with a uint8 intermediate between
// the horizontal and the vertical pass, the float depth plane accumulated in double with a float32 intermediate,
// coefficient tables computed by the caller exactly as precompute_coeffs does (oracle/compositor.py pins them).
// Painter's order becomes "nearest covering object wins" per frame pixel (the blend weight is a rounded 0/1 mask).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sdn_b200.h"

namespace sdn_internal {
void launch_begin(const char* name, cudaStream_t st);
int launch_end();
}  // namespace sdn_internal

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ int clip8(int v) {
    v >>= PRECISION_BITS;                   // arithmetic shift, like Resample.c's lookup index
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// torchvision.to_pil_image of a float tensor: pic.mul(255).byte()
__device__ __forceinline__ int to_byte(float v) { return (int)(unsigned char)__float2int_rz(v * 255.0f); }

// ---- horizontal pass: [S rows, S cols] source -> [S rows, size cols] intermediate, all three planes ----------
__global__ void k_composite_rows(const sdn_composite_obj* __restrict__ objs, const float* __restrict__ masks,
                                 const float* __restrict__ normals, const float* __restrict__ depth_maps,
                                 const float* __restrict__ zooms, int S, const int* __restrict__ bounds,
                                 const int* __restrict__ kk8, const double* __restrict__ kkd,
                                 uchar4* __restrict__ tmp8, float* __restrict__ tmpf) {
    const sdn_composite_obj o = objs[blockIdx.y];
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)S * o.size) return;             // also skips mask-only entries (size == 0)
    const int y = (int)(t / o.size), xx = (int)(t % o.size);
    const int x0 = bounds[(o.row_off + xx) * 2], n = bounds[(o.row_off + xx) * 2 + 1];
    const int* k8 = kk8 + o.k_off + (long)xx * o.ksize;
    const double* kd = kkd + o.k_off + (long)xx * o.ksize;
    const size_t plane = (size_t)S * S;
    const float* m = masks + (size_t)o.index * plane + (size_t)y * S + x0;
    const float* d = depth_maps + (size_t)o.index * plane + (size_t)y * S + x0;
    const float* nr = normals ? normals + (size_t)o.index * 3 * plane + (size_t)y * S + x0 : nullptr;
    const float zoom = zooms[o.index];
    int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
    double ad = 0.0;
    for (int j = 0; j < n; j++) {
        const int k = k8[j];
        a0 += to_byte(m[j]) * k;
        if (nr) {       // _normals / 2 + 0.5  (main.py:565)
            a1 += to_byte(nr[j] / 2.0f + 0.5f) * k;
            a2 += to_byte(nr[plane + j] / 2.0f + 0.5f) * k;
            a3 += to_byte(nr[2 * plane + j] / 2.0f + 0.5f) * k;
        }
        const float dv = fminf(__fdiv_rn(d[j] * zoom, 100.0f), 1.0f);      // main.py:581
        ad += (double)dv * kd[j];
    }
    const size_t at = (size_t)o.tmp_off + (size_t)y * o.size + xx;
    tmp8[at] = make_uchar4((unsigned char)clip8(a0), (unsigned char)clip8(a1), (unsigned char)clip8(a2),
                           (unsigned char)clip8(a3));
    tmpf[at] = (float)ad;
}

// ---- vertical pass + paste + far-to-near composite, one thread per frame pixel ------------------------------
// Entries with size == 0 are MASK-ONLY: a frame-resolution mask (`frame_masks[index]`, [H,W]) blended into the instance
// map alone, `inst = (1 - m) * inst + m * id` -- what main.py:601-604 does with the ground-truth mask of an object that
// is not "interesting" when no edit operations are given.  Rendered entries blend with a rounded 0/1 weight, so for
// them the same expression is an overwrite; the loop therefore walks far -> near for the instance id and remembers
// the nearest covering RENDERED entry, whose normal / depth are evaluated once afterwards.
__global__ void k_composite_frame(const sdn_composite_obj* __restrict__ objs, int M, int S, int H, int W,
                                  const int* __restrict__ bounds, const int* __restrict__ kk8,
                                  const double* __restrict__ kkd, const uchar4* __restrict__ tmp8,
                                  const float* __restrict__ tmpf, const float* __restrict__ frame_masks,
                                  float* __restrict__ instance, float* __restrict__ normal,
                                  float* __restrict__ depth) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const size_t at = (size_t)y * W + x, plane = (size_t)H * W;
    float inst = 0.0f, n0 = 0.5f, n1 = 0.5f, n2 = 0.5f, dep = 1.0f;     // main.py:540-542
    int top = -1;
    for (int mi = 0; mi < M; mi++) {                                      // far -> near (main.py:538,543)
        const sdn_composite_obj o = objs[mi];
        if (o.size == 0) {                                                // mask-only entry (main.py:601-604)
            const float m = frame_masks[(size_t)o.index * plane + at];
            inst = (1.0f - m) * inst + m * (float)o.id;
            continue;
        }
        const int lx = x - o.box_x, ly = y - o.box_y;
        if (lx < 0 || ly < 0 || lx >= o.size || ly >= o.size) continue;
        const int y0 = bounds[(o.row_off + ly) * 2], n = bounds[(o.row_off + ly) * 2 + 1];
        const int* k8 = kk8 + o.k_off + (long)ly * o.ksize;
        const uchar4* col = tmp8 + (size_t)o.tmp_off + (size_t)y0 * o.size + lx;
        int a0 = 1 << (PRECISION_BITS - 1);
        for (int j = 0; j < n; j++) a0 += (int)col[(size_t)j * o.size].x * k8[j];
        if (clip8(a0) < 128) continue;                                    // torch.round(v / 255) == 0
        inst = (float)o.id;                                               // (1 - 1) * inst + 1 * id
        top = mi;
    }
    if (top >= 0) {
        const sdn_composite_obj o = objs[top];
        const int lx = x - o.box_x, ly = y - o.box_y;
        const int y0 = bounds[(o.row_off + ly) * 2], n = bounds[(o.row_off + ly) * 2 + 1];
        const int* k8 = kk8 + o.k_off + (long)ly * o.ksize;
        const uchar4* col = tmp8 + (size_t)o.tmp_off + (size_t)y0 * o.size + lx;
        int a1 = 1 << (PRECISION_BITS - 1), a2 = a1, a3 = a1;
        double ad = 0.0;
        const double* kd = kkd + o.k_off + (long)ly * o.ksize;
        const float* fcol = tmpf + (size_t)o.tmp_off + (size_t)y0 * o.size + lx;
        for (int j = 0; j < n; j++) {
            const uchar4 c = col[(size_t)j * o.size];
            a1 += (int)c.y * k8[j]; a2 += (int)c.z * k8[j]; a3 += (int)c.w * k8[j];
            ad += (double)fcol[(size_t)j * o.size] * kd[j];
        }
        n0 = __fdiv_rn((float)clip8(a1), 255.0f);                         // to_tensor: byte / 255
        n1 = __fdiv_rn((float)clip8(a2), 255.0f);
        n2 = __fdiv_rn((float)clip8(a3), 255.0f);
        dep = (float)ad;
    }
    instance[at] = inst;
    if (normal) { normal[at] = n0; normal[plane + at] = n1; normal[2 * plane + at] = n2; }
    depth[at] = dep;
}

}  // namespace

extern "C" int sdn_composite_scene(const sdn_composite_obj* objs, int M, const float* masks, const float* normals,
                                   const float* depth_maps, const float* zooms, int S, const int* bounds,
                                   const int* kk8, const double* kkd, void* tmp8, float* tmpf, int max_size, int H,
                                   int W, const float* frame_masks, int n_rendered, float* instance, float* normal,
                                   float* depth, void* stream) {
    if (M < 0 || n_rendered < 0 || n_rendered > M || S <= 0 || H <= 0 || W <= 0 || !instance || !depth ||
        (normals == nullptr) != (normal == nullptr))
        return SDN_E_ARG;
    if (M > 0 && !objs) return SDN_E_ARG;
    if (n_rendered > 0 && (!masks || !depth_maps || !zooms || !bounds || !kk8 || !kkd || !tmp8 || !tmpf ||
                           max_size <= 0))
        return SDN_E_ARG;
    if (n_rendered < M && !frame_masks) return SDN_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (n_rendered > 0) {
        const dim3 grid((unsigned)(((long)S * max_size + 255) / 256), (unsigned)M);
        sdn_internal::launch_begin("k_composite_rows", st);
        k_composite_rows<<<grid, 256, 0, st>>>(objs, masks, normals, depth_maps, zooms, S, bounds, kk8, kkd,
                                               (uchar4*)tmp8, tmpf);
        if ((rc = sdn_internal::launch_end())) return rc;
    }
    sdn_internal::launch_begin("k_composite_frame", st);
    k_composite_frame<<<dim3((unsigned)((W + 127) / 128), (unsigned)H), 128, 0, st>>>(
        objs, M, S, H, W, bounds, kk8, kkd, (const uchar4*)tmp8, tmpf, frame_masks, instance, normal, depth);
    return sdn_internal::launch_end();
}
