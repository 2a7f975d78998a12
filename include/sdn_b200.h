/*
 * sdn_b200.h -- C ABI of the B200-native (sm_100a) differentiable mesh rasteriser that replaces the
 * render hot path of ysymyth/3D-SDN (geometric/neural_renderer + the bridge in geometric/derender3d).
 *
 * Boundary contract:
 *   - extern "C", raw DEVICE pointers, row-major, float32 / int32, explicit cudaStream_t (passed as void*),
 *   - every function returns int: 0 = ok, >0 = cudaError_t, <0 = SDN_E_* argument error; nothing throws or exits,
 *   - the library never allocates or frees device memory and keeps no state that influences results; scratch comes
 *     in as (workspace, bytes), sized by sdn_workspace_bytes(); the SAME workspace that served a forward must be
 *     handed to its backward.  The only process-wide state is diagnostic: a launch counter, the opt-in per-launch
 *     event list of sdn_timing_enable() (mutex-protected; the bracket in flight is thread-local) and per-device
 *     caches of constant device attributes (SM count, shared-memory opt-in already granted),
 *   - re-entrant: works on the caller's current device and the stream given; no device-wide synchronisation.
 *
 * Each entry point names the reference interface it replaces (paths relative to /root/reference/geometric/).
 * Rasters are H rows x W columns (reference: image_size x image_size); "ss" is the super-sampling factor
 * (2 when anti_aliasing, rasterize.py:942-945, else 1).  F is the EFFECTIVE face count (2*nf with fill_back).
 */
#ifndef SDN_B200_H
#define SDN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDN_VERSION 201            /* 0.2.0: sdn_composite_scene takes frame masks */

/* what to render: rasterize.py:21-22 return_rgb / return_alpha / return_depth */
#define SDN_RGB    1
#define SDN_ALPHA  2
#define SDN_DEPTH  4

/* argument errors */
#define SDN_E_ARG        (-1)
#define SDN_E_WORKSPACE  (-2)
#define SDN_E_NOTHING    (-3)      /* nothing to draw: rasterize.py:25-27 raises */

/* Camera applied per image by sdn_faces_forward / sdn_faces_backward.  All pointers are device pointers and may
 * be NULL (= that step is skipped).  rot rows are x_axis, y_axis, z_axis (look.py:33-35, look_at.py:34-36);
 * width_x / width_y are tan(viewing half-angle) per image (perspective.py:13); flip_x multiplies x first
 * (derender3d/models/renderer.py:243 uses -1). */
typedef struct sdn_camera {
    const float* rot;       /* [B,3,3] or NULL */
    const float* eye;       /* [B,3]   or NULL */
    const float* width_x;   /* [B]     or NULL: no perspective division */
    const float* width_y;   /* [B]     (required when width_x is set) */
    float flip_x;           /* +1 or -1 */
} sdn_camera;

/* One rasterisation call == one reference `Rasterize` instance (rasterize.py:19-64) plus the SSAA / flip /
 * transpose wrapper rasterize_rgbad (rasterize.py:897-974). */
typedef struct sdn_raster {
    int B, F;               /* batch, effective faces */
    int H, W;               /* OUTPUT rows, cols */
    int ss;                 /* 1 or 2 */
    int flags;              /* SDN_RGB | SDN_ALPHA | SDN_DEPTH */
    int texture_size;       /* ts >= 2: textures [B,F,ts,ts,ts,3];  1: one colour per face [B,F,3]
                               (the constant 2^3 cube of derender3d/models/renderer.py:75-76) */
    int batch_bug;          /* 1 reproduces rasterize.py:390 (face pointer without bn*nf) for B > 1 */
    double near, far;       /* rasterize.py:353-354: substituted as C double literals */
    double eps;             /* rasterize.py:401 (texture index) and :647 (K6 distance pad) */
    float background[3];    /* rasterize.py:457-460, used when background_b == NULL */
    const float* background_b;  /* [B,3] or NULL (rasterize.py:461-462) */

    const float* faces;     /* [B,F,3,3] NDC faces (vertices_to_faces output) */
    const float* textures;  /* see texture_size; NULL unless SDN_RGB */

    /* outputs, caller-allocated; NULL when not requested */
    float* rgb;             /* [B,3,H,W] */
    float* alpha;           /* [B,H,W]   */
    float* depth;           /* [B,H,W]   */
    int32_t* face_index;    /* [B,H*ss,W*ss]  raster orientation (row 0 = bottom), -1 = background; REQUIRED */
    float* rgb_ss;          /* [B,H*ss,W*ss,3] pre-resolve colours; required for backward when texture_size >= 2 */

    void* workspace;        /* sdn_workspace_bytes(B, F, H*ss, W*ss) */
    size_t workspace_bytes;
} sdn_raster;

/* Gradients of one rasterisation.  Mirrors Rasterize.backward_gpu (rasterize.py:846-886) including the
 * avg-pool / flip / transpose backward of rasterize_rgbad.  A NULL grad_* means "no gradient from that output"
 * (rasterize.py:856-875 substitute zeros). */
typedef struct sdn_raster_grad {
    const float* grad_rgb;      /* [B,3,H,W] or NULL */
    const float* grad_alpha;    /* [B,H,W]   or NULL */
    const float* grad_depth;    /* [B,H,W]   or NULL */
    float* grad_faces;          /* [B,F,3,3] OUT (fully written) */
    float* grad_textures;       /* same shape as textures, OUT (fully written) or NULL */
    int fast_math;              /* 0: IEEE divisions in the op order of rasterize.py:646-653;
                                   1: reciprocal-multiply distance math (<= 1e-6 relative, not bit-faithful)
                                   2: as 1, plus the far-field (block-moment) sweep when only alpha is rendered */
} sdn_raster_grad;

int sdn_version(void);
const char* sdn_error_string(int code);

/* bytes of scratch for one (B, F, Hs = H*ss, Ws = W*ss) rasterisation */
size_t sdn_workspace_bytes(int B, int F, int Hs, int Ws);

/* Replaces fill_back (neural_renderer/renderer.py:41), look / look_at (look.py:38-43, look_at.py:39-44),
 * perspective (perspective.py:13-18), vertices_to_faces (vertices_to_faces.py:16-21) and, when face_normal is
 * given, the per-face normals of derender3d/models/renderer.py:66-73 (taken BEFORE the camera, AFTER flip_x; their x
 * component is multiplied by flip_x again, i.e. reported in the caller's frame -- the negation of renderer.py:268-270).
 *   vertices [B,nv,3]; faces [B,nf,3] (faces_batched=1) or [nf,3] shared by the batch (0)
 *   face_ndc [B,F,3,3], face_normal [B,F,3] or NULL;  F = fill_back ? 2*nf : nf */
int sdn_faces_forward(const float* vertices, const int32_t* faces, int B, int nv, int nf, int faces_batched,
                      int fill_back, const sdn_camera* cam, float* face_ndc, float* face_normal, void* stream);

/* Backward of the above: scatter-add of grad_face_ndc / grad_face_normal onto the vertices, then the chain rule
 * through perspective, the camera rotation and flip_x.  grad_vertices [B,nv,3] is fully written.
 * scratch: 2*B*nv*3 floats. */
int sdn_faces_backward(const float* grad_face_ndc, const float* grad_face_normal, const float* vertices,
                       const int32_t* faces, int B, int nv, int nf, int faces_batched, int fill_back,
                       const sdn_camera* cam, float* grad_vertices, void* scratch, size_t scratch_bytes,
                       void* stream);

/* Replaces Rasterize.forward_gpu (rasterize.py:464-510: K1a+K1b visibility, K2 texture sampling, alpha map,
 * background) and rasterize_rgbad's transpose / row flip / 2x2 mean (rasterize.py:950-966). */
int sdn_rasterize_forward(const sdn_raster* r, void* stream);

/* Replaces Rasterize.backward_gpu (rasterize.py:846-886: K6 pixel-map gradient, K7 texture gradient,
 * K8 depth gradient).  `r` must be the descriptor (same buffers, same workspace contents) of the forward. */
int sdn_rasterize_backward(const sdn_raster* r, const sdn_raster_grad* g, void* stream);

/* Fused multi-output backward used by the derender3d bridge, where silhouette, depth and normal images of the
 * same faces come from three reference Rasterize instances with different eps: the alpha
 * gradient uses eps_alpha with alpha only, the rgb gradient uses eps_rgb with rgb only (each with its own
 * `diff_grad > 0` test, rasterize.py:644,714), and the results add. */
int sdn_rasterize_backward_split(const sdn_raster* r, const sdn_raster_grad* g, double eps_alpha, double eps_rgb,
                                 void* stream);

/* ---- pre-raster stage: free-form deformation + camera-space transform of a ragged batch of objects --------------
 * Replaces, for all N objects of a batch at once, the body of the per-object loop of Derenderer3d.render
 * (derender3d/models/__init__.py:161-200): FFD.forward (derender3d/models/transforms.py:68-99, the sum
 * V = sum (P0+dP) * B) and PerspectiveTransform.forward (transforms.py:102-158).  Object i deforms template mesh
 * classes[i]; the templates' Bernstein tables B (transforms.py:57-64) are concatenated in `basis`.  The linear
 * constraint step (transforms.py:71-95) stays with the caller: `control` is P0 + constrained dP.
 * All pointers are device pointers; nothing is allocated. */
typedef struct sdn_objects {
    int N;                          /* objects in the batch */
    int K;                          /* template meshes in the bank */
    int nv_max;                     /* vertex stride of the output (>= every nv[k]) */
    int derive_zoom;                /* 0: training mode, zooms given (transforms.py:151); 1: test mode, zoom_i =
                                       min_v |z|/max(|x|,|y|) * zoom_tos[i] (transforms.py:148-149), written to zooms */
    const float* basis;             /* [sum_k nv[k], 64] Bernstein weights, row = vertex, 64 = 4x4x4 lattice (i,j,k) */
    const long long* basis_off;     /* [K] first row of template k in basis */
    const int* nv;                  /* [K] vertex count of template k */
    const long long* classes;       /* [N] template id of object i (int64, as torch.max / Categorical.sample give) */
    const float* control;           /* [N,3,64] control points P0 + dP */
    const float* scales;            /* [N,3] */
    const float* rotations;         /* [N,4] quaternion (a,b,c,d), transforms.py:110-124 */
    const float* translations;      /* [N,3] */
    const float* ptrans;            /* [N,3] perspective_translations (principal ray), transforms.py:133-143 */
    float* zooms;                   /* [N] read when derive_zoom == 0, written when 1 */
    const float* zoom_tos;          /* [N] derive_zoom == 1 only */
    unsigned long long* zoom_key;   /* [N] scratch kept between forward and backward when derive_zoom == 1 */
    float* vertices;                /* out [N,nv_max,3]; rows >= nv[classes[i]] are (0,0,1) */
} sdn_objects;

int sdn_objects_forward(const sdn_objects* o, void* stream);
/* gradients of everything differentiable; grad_zoom is d/d zooms (derive_zoom 0) or d/d zoom_tos (1).
 * scratch: >= 96 * N bytes.  `o` must describe the forward's buffers (zooms / zoom_key as the forward left them). */
int sdn_objects_backward(const sdn_objects* o, const float* grad_vertices, float* grad_control, float* grad_scales,
                         float* grad_rotations, float* grad_translations, float* grad_ptrans, float* grad_zoom,
                         float* scratch, size_t scratch_bytes, void* stream);

/* ---- convolution on the tcgen05 tensor cores -------------------------------------------------------------------
 * Replaces the cuDNN convolutions behind torch.nn.Conv2d in the image encoder's ResNet-18 trunk
 * (geometric/derender3d/models/derenderer.py:25-27,35-43 -> torchvision resnet18) and in the textural generator
 * (textural/models/networks.py:211-283).  Implicit GEMM, bf16 operands, fp32 accumulation in tensor memory:
 *   output[n, ho, wo, co] = act( scale[co] * sum_{kh,kw,ci} input[n, ho*stride + kh - pad, wo*stride + kw - pad, ci]
 *                                                     * weight[co, kh, kw, ci]  + shift[co]  (+ residual[n, ho, wo, co]) )
 * input  [N, H, W, Cin]   bf16 NHWC (channels-last), Cin a multiple of 64 (pad with zero channels);
 * weight [Cout, KH, KW, Cin] bf16 (K contiguous),   Cout a multiple of 64, or exactly 16 (few-channel heads);
 * output [N, Ho, Wo, Cout] bf16 NHWC, Ho = (H + 2 pad - KH) / stride + 1; residual (optional) has the output's shape;
 * scale / shift [Cout] fp32 (NULL = 1 / 0): a folded BatchNorm (eval mode) or a bias.  All pointers 16-byte aligned.
 * Zero padding is done by the TMA unit (out-of-bounds box elements read as 0); other paddings (reflection) are the
 * caller's (pad the input, pass pad = 0). */
typedef struct sdn_conv2d {
    int N, H, W, Cin, Cout, KH, KW, stride, pad;
    int relu;               /* activation after scale / shift / residual: 0 none, 1 ReLU, 2 tanh */
    const void* input;
    const void* weight;
    const float* scale;
    const float* shift;
    const void* residual;
    void* output;
    int out_h, out_w;       /* 0 = the convolution's own extent; else that many output rows / columns (input reads
                               beyond the tensor are zeros: one-sided padding, used by the transposed-convolution phases) */
    long long out_stride_n, out_stride_h, out_stride_w;   /* element strides of an output pixel; all 0 = dense NHWC.
                               Multiples of 8.  Lets four parity phases interleave into one [N, 2H, 2W, C] tensor. */
    float* stats;           /* NULL, or fp32 [N, Cout, 2]: the call clears it and adds, per (image, channel), the sum and
                               the sum of squares of scale * conv + shift (no residual allowed) -- the statistics the following
                               InstanceNorm2d needs, taken from the accumulators instead of a second pass */
    int stats_keep;         /* 1: add into `stats` without clearing it first (later phases of a transposed convolution) */
} sdn_conv2d;

int sdn_conv2d_forward(const sdn_conv2d* c, void* stream);

/* Weight gradient of the convolution above (training path of the encoder trunk; replaces cuDNN's backward-filter behind
 * torch.autograd for the convolutions of derenderer.py:25-27):
 *   grad_weight[co, kh, kw, ci] = sum_{n,ho,wo} grad_out[n,ho,wo,co] * input[n, ho*stride + kh - pad, wo*stride + kw - pad, ci]
 * grad_out [N,Ho,Wo,Cout] bf16, input [N,H,W,Cin] bf16 (Cin, Cout multiples of 64), grad_weight fp32 [Cout,KH,KW,Cin]
 * (overwritten; with accumulate = 1 the result is ADDED to what grad_weight holds -- e.g. straight into a parameter's .grad --
 * by vector reductions, no clearing pass).  The data gradient needs no entry point of its own: it is sdn_conv2d_forward of grad_out with the
 * spatially flipped, channel-transposed weights (stride 1) or, for stride 2, its four output-parity phases. */
int sdn_conv2d_wgrad(const void* grad_out, const void* input, float* grad_weight, int N, int H, int W, int Cin, int Cout,
                     int KH, int KW, int stride, int pad, int accumulate, void* stream);

/* BatchNorm2d in training mode around those convolutions (torchvision BasicBlock), NHWC bf16, `pixels` = N*H*W:
 *   forward : y = act(z * scale[c] + shift[c] (+ residual)), scale = gamma * rstd, shift = beta - mean * scale (the batch
 *             statistics come from sdn_conv2d.stats); act 0 none / 1 ReLU;
 *   backward: g^ = grad_y * [y > 0] (relu = 1), sums[c] = (sum g^, sum g^ * zhat) -- i.e. (grad_beta, grad_gamma) --,
 *             grad_z = gamma * rstd * (g^ - sums[c,0] / pixels - zhat * sums[c,1] / pixels); grad_skip (optional) = g^,
 *             the gradient entering the block's skip connection.  sums: fp32 [C, 2], cleared by the call unless bit 1 of
 *             `relu` is set (relu = 2 | 3: the caller cleared it, e.g. one clearing pass for all layers of a network).
 *             grad_gamma_acc / grad_beta_acc (optional, fp32 [C]): sums[c,1] / sums[c,0] are added to them (parameter
 *             gradients accumulated in place). */
int sdn_affine_act_nhwc(const void* z, const float* scale, const float* shift, const void* residual, void* y,
                        long long pixels, int C, int act, void* stream);
int sdn_batch_norm_backward_nhwc(const void* grad_y, const void* y, const void* z, const float* mean, const float* rstd,
                                 const float* gamma, float* sums, void* grad_z, void* grad_skip, float* grad_gamma_acc,
                                 float* grad_beta_acc, long long pixels, int C, int relu, void* stream);

/* nn.MaxPool2d(K, stride, pad) of the ResNet stem on NHWC bf16 (derenderer.py:25-27 -> torchvision resnet18.maxpool), C a
 * multiple of 8, floor mode, y [N, Ho, Wo, C]; argmax (optional, uint8 [N, Ho, Wo, C]) receives kh * K + kw of each window's
 * first maximum in scan order -- the element torch's max_pool2d_with_indices records.  Backward: grad_x [N,H,W,C] (every
 * element written) from grad_y and argmax; gather form, no atomics. */
int sdn_max_pool_nhwc(const void* x, void* y, void* argmax, int N, int H, int W, int C, int K, int stride, int pad,
                      void* stream);
int sdn_max_pool_backward_nhwc(const void* argmax, const void* grad_y, void* grad_x, int N, int H, int W, int C, int K,
                               int stride, int pad, void* stream);

/* Batch statistics of a convolution result (sdn_conv2d.stats, fp32 [N, C, 2] = per-image sum / sum of squares) -> out
 * fp32 [4, C] = (mean, rstd, scale = gamma * rstd, shift = beta - mean * scale); running_mean / running_var (optional)
 * are updated in place with `momentum` and the unbiased variance, *num_batches (optional) is incremented: what
 * torch.nn.BatchNorm2d does in training mode (reference: torchvision resnet18 inside derenderer.py:25-27). */
int sdn_batch_norm_finalize(const float* stats, int N, int C, long long pixels, float eps, float momentum, const float* gamma,
                            const float* beta, float* running_mean, float* running_var, long long* num_batches, float* out,
                            void* stream);

/* Operand of the data gradient: weight fp32 [Cout, KH, KW, Cin] (physical layout of a channels-last Conv2d weight) ->
 * bf16.  stride 1 (or 1x1): [Cin, KH, KW, Cout] with flipped taps.  stride 2, 3x3: the four output-parity phases
 * (0,0) (0,1) (1,0) (1,1) concatenated, each [Cin, Ta, Tb, Cout] (taps {1} for parity 0, {2, 0} for parity 1). */
int sdn_pack_conv_weight_transposed(const float* weight, void* packed, int Cout, int Cin, int KH, int KW, int stride,
                                    void* stream);

/* im2col for inputs with few channels (the 7x7 stem, Cin = 3): output [N, Ho, Wo, Kpad] bf16 with
 * output[.., (kh*KW + kw)*Cin + ci] = input[n, ho*stride + kh - pad, wo*stride + kw - pad, ci] (0 outside and beyond
 * KH*KW*Cin); the stem then runs through sdn_conv2d_forward as a 1x1 convolution over Kpad channels. */
int sdn_im2col_nhwc(const void* input, void* output, int N, int H, int W, int Cin, int KH, int KW, int stride, int pad,
                    int Kpad, void* stream);

/* Instance normalisation of an NHWC bf16 tensor [N, HW, C] (textural/models/networks.py:23-29: InstanceNorm2d with
 * affine=False): y = act((x - mean_nc) / sqrt(var_nc + eps) (+ residual)), biased variance over HW, act 0 none / 1 ReLU.
 * stats: fp32 scratch [N, C, 2] (sum, sum of squares), cleared by the call.  C a multiple of 8; y may alias x. */
int sdn_instance_norm_nhwc(const void* x, const void* residual, void* y, float* stats, int N, int HW, int C, float eps,
                           int act, void* stream);

/* The normalisation alone, with statistics [N, C, 2] produced elsewhere (sdn_conv2d.stats; several convolutions may have
 * added into one buffer, e.g. the four phases of a transposed convolution). */
int sdn_instance_norm_apply_nhwc(const void* x, const void* residual, void* y, const float* stats, int N, int HW, int C,
                                 float eps, int act, void* stream);

/* Normalisation + nn.ReflectionPad2d(pad) in one pass: y [N, H + 2 pad, W + 2 pad, C] = reflect_pad(act(norm(x) (+ residual)))
 * -- what the generator's next 3x3 / 7x7 convolution reads (networks.py:250-269). */
int sdn_instance_norm_apply_pad_nhwc(const void* x, const void* residual, void* y, const float* stats, int N, int H, int W,
                                     int C, float eps, int act, int pad, void* stream);

/* nn.ReflectionPad2d(pad) on an NHWC bf16 tensor, channels zero-padded from C to Cp (a multiple of 8):
 * y [N, H + 2 pad, W + 2 pad, Cp] (textural/models/networks.py:218,238,250-269). */
int sdn_reflect_pad_nhwc(const void* x, void* y, int N, int H, int W, int C, int pad, int Cp, void* stream);

/* ---- test-time scene compositor ------------------------------------------------------------------------------
 * Replaces the per-object loop of geometric/scripts/main.py:547-606 (to_pil_image -> resize -> PIL paste -> to_tensor
 * on the CPU, far to near) for the M drawn objects of one frame.  Bit-identical to Pillow's bilinear resize; the
 * caller supplies Pillow's coefficient tables (Resample.c precompute_coeffs; oracle/compositor.py `pil_coeffs`).
 * objs[] is sorted far to near.  All pointers are device pointers. */
typedef struct sdn_composite_obj {
    int index;              /* row of masks / normals / depth_maps / zooms this object reads */
    int size;               /* on-screen edge: int(render_size / zoom), main.py:556; 0 = MASK-ONLY entry: `index` is a
                               row of frame_masks, blended into the instance map alone (main.py:601-604) */
    int box_x, box_y;       /* paste position of the top-left corner (may be negative), main.py:562-565 */
    int ksize;              /* taps per output sample in the coefficient tables */
    int id;                 /* value written to the instance map: object index + 1, main.py:570 */
    long long row_off;      /* first row of this object in `bounds` */
    long long k_off;        /* first element of this object in kk8 / kkd ([size][ksize]) */
    long long tmp_off;      /* first pixel of this object's [S][size] intermediate in tmp8 / tmpf */
} sdn_composite_obj;

/* masks [N,S,S], normals [N,3,S,S] or NULL (then normal must be NULL), depth_maps [N,S,S], zooms [N];
 * bounds [sum size][2] (first tap, count), kk8 / kkd [sum size*ksize]; tmp8: 4 bytes per intermediate pixel, tmpf:
 * one float; max_size = max objs[].size.  frame_masks [*,H,W] (may be NULL when every entry is rendered): the
 * frame-resolution masks of the mask-only entries -- objects that are not "interesting" still occlude and get an
 * instance id when no edit operations are given (main.py:601-604); n_rendered = number of entries with size > 0.
 * Out: instance [H,W], normal [3,H,W], depth [H,W] (backgrounds 0, 0.5, 1). */
int sdn_composite_scene(const sdn_composite_obj* objs, int M, const float* masks, const float* normals,
                        const float* depth_maps, const float* zooms, int S, const int* bounds, const int* kk8,
                        const double* kkd, void* tmp8, float* tmpf, int max_size, int H, int W,
                        const float* frame_masks, int n_rendered, float* instance, float* normal, float* depth,
                        void* stream);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long sdn_launch_count(void);

/* ---- encoder fully-connected layer on the tcgen05 tensor cores (bf16 operands, fp32 accumulate in TMEM) ----------
 * y[M,N] = act(x[M,K] . weight[N,K]^T + bias[N]); all fp32 device pointers, row-major, rows 16-byte aligned when
 * K % 4 == 0 (any K is accepted).  bias may be NULL; relu != 0 applies max(.,0).  Replaces the torch.nn.Linear calls
 * of the encoder head: derender3d/models/derenderer.py:26-32 (net.fc, fc1, fc2, _fc3) and :35-43 (their use).
 * Operands are rounded to bf16 (8-bit mantissa): |err| <= ~2^-8 * sum|x_k w_k|; the test states the tolerance. */
int sdn_linear_forward(const float* x, const float* weight, const float* bias, float* y, int M, int N, int K,
                       int relu, void* stream);

/* The whole encoder head in one launch (derenderer.py:26-32,35-43), fp32 operands and accumulation, for small batches (any
 * M; meant for M <= 64):  cat = [relu(x . W0^T + b0) | extra],  h1 = relu(cat . W1^T + b1),  h2 = relu(h1 . W2^T + b2),
 * y = h2 . W3^T + b3.   x [M,K0], extra [M,E] (E may be 0), W_i row-major [N_i, K_i] as torch.nn.Linear stores them; cat
 * [M, N0+E], h1 [M,N1], h2 [M,N2], y [M,N3] are written (the backward needs them); K0, N0+E, N1, N2 <= 1024.
 * `barrier`: one device word used by the grid-wide barriers between the layers (cleared by the call). */
int sdn_mlp_head_forward(const float* x, const float* extra, int M, int K0, int N0, int E, int N1, int N2, int N3,
                         const float* W0, const float* b0, const float* W1, const float* b1, const float* W2, const float* b2,
                         const float* W3, const float* b3, float* cat, float* h1, float* h2, float* y, unsigned* barrier,
                         void* stream);

/* Diagnostics: when enabled, every kernel launch is bracketed by CUDA events on the launching stream;
 * sdn_timing_read() drains them as "kernel_name total_ms launches\n" lines (returns bytes written).  Off by default. */
int sdn_timing_enable(int on);
int sdn_timing_read(char* buf, int cap);
int sdn_debug_pending(char* buf, int cap);   /* timed launches not finished yet: "#i name RUNNING|queued\n" */

#ifdef __cplusplus
}
#endif
#endif /* SDN_B200_H */
