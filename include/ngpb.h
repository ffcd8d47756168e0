/*
 * ngpb.h -- C ABI of the B200-native NeRF train/render hot path (libngpb200.so).
 *
 * The reference (JamesPerlman/blender-ngp) has no C/FFI boundary for this path: kernels are
 * called from `Testbed` members and tiny-cuda-nn objects are C++ virtual classes.
 * This header is the seam a maintainer would bind instead. Two levels:
 *
 *   1. kernel level (`ngpb_*` taking raw device pointers + a cudaStream_t): one entry point per
 *      stage of the reference's training iteration, each citing the reference code it replaces;
 *   2. testbed level (`ngpb_testbed_*`, host buffers in/out): the operations `pyngp.Testbed`
 *      exposes for this path (src/python_api.cu:540-732).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a non-zero
 * status (a cudaError_t value, or NGPB_ERR_*), with the message available from
 * ngpb_last_error(); the caller supplies the stream (void* = cudaStream_t); kernel-level entry
 * points never allocate. Paths in comments are relative to the reference root.
 */
#ifndef NGPB_H
#define NGPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGPB_ERR_INVALID_ARGUMENT 10001
#define NGPB_ERR_RUNTIME          10002

typedef uint16_t ngpb_half; /* IEEE binary16 bit pattern (tcnn::network_precision_t == __half) */

const char* ngpb_last_error(void);
int ngpb_version(void);
/* Compute capability check: returns 0 only on an sm_100 device. */
int ngpb_check_device(int device);

/* ---- enums: integer values equal the reference's (include/neural-graphics-primitives/common.h:103-134) ---- */
enum { NGPB_LOSS_L2 = 0, NGPB_LOSS_L1 = 1, NGPB_LOSS_MAPE = 2, NGPB_LOSS_SMAPE = 3, NGPB_LOSS_HUBER = 4, NGPB_LOSS_LOGL1 = 5, NGPB_LOSS_RELATIVE_L2 = 6 };
enum { NGPB_ACT_NONE = 0, NGPB_ACT_RELU = 1, NGPB_ACT_LOGISTIC = 2, NGPB_ACT_EXPONENTIAL = 3 };
enum { NGPB_COLOR_LINEAR = 0, NGPB_COLOR_SRGB = 1 };

/* ---- hash grid (replaces tcnn GridEncodingTemplated<__half,3,2>, grid.h:944-1198) ---- */
#define NGPB_MAX_LEVELS 32
typedef struct {
	uint32_t n_levels;            /* 16 */
	uint32_t base_resolution;     /* 16 */
	float log2_per_level_scale;   /* std::log2(per_level_scale), as passed to the kernel at grid.h:1079 */
	uint32_t offsets[NGPB_MAX_LEVELS + 1]; /* entry offsets per level; offsets[n_levels] = total entries (grid.h:985-1018) */
	float scale[NGPB_MAX_LEVELS];          /* grid_scale(level) = exp2f(level*log2_pls)*base-1 (grid.h:194-199), host-evaluated by
	                                          ngpb_grid_init; the reference evaluates it per thread with the device exp2f */
	uint32_t resolution[NGPB_MAX_LEVELS];  /* grid_resolution(scale) = ceil(scale)+1 (grid.h:201-203) */
	uint32_t n_pos_dims;                   /* 3 (NeRF, SDF) or 2 (neural image); 0 reads as 3 */
} ngpb_grid;

/* Host-only: fills `g` like the GridEncodingTemplated constructor (grid.h:959-1025). Returns total entries. */
uint32_t ngpb_grid_init(ngpb_grid* g, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale);
/* Same for an N-dimensional input (N_POS_DIMS of GridEncodingTemplated): 2 = the neural-image model (configs/image/base.json), 3 = NeRF / SDF.
 * The testbed / render entry points are 3-D only; ngpb_hash_encode_forward / _backward and ngpb_model accept 2-D grids. */
uint32_t ngpb_grid_init_nd(ngpb_grid* g, uint32_t n_pos_dims, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale);

/* Replaces g->scale[] / g->resolution[] by the values the DEVICE's exp2f gives, which is what the reference's kernels use (they call
 * grid_scale per thread, grid.h:241); glibc's exp2f differs by one ulp on some levels. The level offsets stay host-derived, as in
 * the reference's constructor. Needs a GPU; synchronises the stream. */
int ngpb_grid_device_scales(void* stream, ngpb_grid* g);

/* kernel_grid (grid.h:220-349). positions: float, `pos_stride` floats per sample (3 used).
 * encoded: [n][2*n_levels] half, sample-major (the MLP kernels' input layout). */
int ngpb_hash_encode_forward(void* stream, const ngpb_grid* g, const ngpb_half* grid, const float* positions, uint32_t pos_stride,
                             uint32_t n, ngpb_half* encoded);
/* kernel_grid_backward (grid.h:395-518). dL_dencoded: [n][2*n_levels] half. grid_grad: float[2*total entries],
 * accumulated into with fp32 atomics (zero it first for EGradientMode::Overwrite, grid.h:1154). */
int ngpb_hash_encode_backward(void* stream, const ngpb_grid* g, const float* positions, uint32_t pos_stride, uint32_t n,
                              const ngpb_half* dL_dencoded, float* grid_grad);

/* The 32 -> 64 -> 64 -> 16 fully fused MLP alone (ReLU hidden, no output activation): the network of the neural-image and SDF models
 * (configs/image/base.json, configs/sdf/base.json; tcnn FullyFusedMLP<__half,64>::inference_mixed_precision, fully_fused_mlp.cu:500-557,:661-729).
 * weights: FullyFusedMLP's parameter order, row-major [out][in]: [64][32], [64][64], [16][64] (7168 halves). input [n][32], output [n][16]
 * (the padded output width; an image model uses columns 0..2, an SDF column 0). n: multiple of 128. */
int ngpb_mlp_forward(void* stream, const ngpb_half* weights, const ngpb_half* input, uint32_t n, ngpb_half* output);
/* Element-wise losses of the image / SDF models (tcnn losses/l2.h:40-80, losses/mape.h:40-80 through Trainer::training_step, trainer.h:121-159):
 * predictions [n][16] fp16 (padded network output), targets [n][dims] fp32 -> values [n][16] fp32 (nullable; per-element loss / (n * dims)) and
 * gradients [n][16] fp16 = loss_scale * dL/d(prediction) / (n * dims), padded columns zero. */
#define NGPB_ELEMENT_LOSS_L2 0
#define NGPB_ELEMENT_LOSS_MAPE 1
#define NGPB_ELEMENT_LOSS_RELATIVE_L2 2 /* losses/relative_l2.h:40-77 */
int ngpb_loss(void* stream, int kind, uint32_t n, uint32_t dims, float loss_scale, const ngpb_half* predictions, const float* targets, float* values,
              ngpb_half* gradients);
/* Forward + backward of the same network in one kernel (FullyFusedMLP::forward + backward, fully_fused_mlp.cu:151-314,:759-850): recomputes the activations of
 * `input`, then from dL_dout [n][16] (all padded columns, as the loss kernel wrote them) produces dL_dinput [n][32] (fp16, for the hash-grid backward) and the
 * weight gradients grad[7168] (fp32, overwritten, summed over the batch, loss scale left in). workspace: ngpb_nerf_mlp_workspace_bytes() bytes. */
int ngpb_mlp_forward_backward(void* stream, const ngpb_half* weights, const ngpb_half* input, const ngpb_half* dL_dout, uint32_t n, ngpb_half* dL_dinput,
                              float* grad, void* workspace);

/* ---- NeRF MLPs (replaces NerfNetwork glue nerf_network.h:103-266 + tcnn FullyFusedMLP, fully_fused_mlp.cu) ----
 * mlp: half[10240] in the reference's flat order: density W1[64][32], W2[16][64]; rgb W1[64][32], W2[64][64], W3[16][64].
 * coords: [n][7] float NerfCoordinate (nerf.h:81-107). rgbsigma: [n][4] half {r,g,b,sigma} network outputs (pre-activation).
 * n must be a multiple of 128 (tcnn batch_size_granularity, common.h:280). */
int ngpb_nerf_mlp_forward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, uint32_t n, ngpb_half* rgbsigma);
/* Forward + backward in one pass: dL_dout [n][4] half -> dL_dencoded [n][32] half; mlp_grad float[10240] is overwritten.
 * workspace: ngpb_nerf_mlp_workspace_bytes() bytes of device memory. */
uint64_t ngpb_nerf_mlp_workspace_bytes(void);
int ngpb_nerf_mlp_forward_backward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, const ngpb_half* dL_dout,
                                   uint32_t n, ngpb_half* dL_dencoded, float* mlp_grad, void* workspace);
/* As above, and dL_dsh [n][16] half: the gradient with respect to the rgb network's spherical-harmonics inputs (columns 16..31 of its input), which
 * camera-extrinsics optimisation carries on to dL/d(direction) (NerfNetwork::backward_impl with dL_dinput, nerf_network.h:286-400). */
int ngpb_nerf_mlp_forward_backward_sh(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, const ngpb_half* dL_dout,
                                      uint32_t n, ngpb_half* dL_dencoded, float* mlp_grad, void* workspace, ngpb_half* dL_dsh);
/* density network only (NerfNetwork::density, nerf_network.h:268-284): encoded [n][32] -> density logit half[n]. */
int ngpb_nerf_density_mlp_forward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, uint32_t n, ngpb_half* density);

/* ---- training images ---- */
typedef struct {
	const uint8_t* pixels; /* device pointer; layout by image_type below (16-byte aligned for the float type, 8 for half) */
	int32_t w, h;
	float fx, fy;          /* focal length in pixels (nerf_loader.h:42) */
	float cx, cy;          /* principal point as a fraction of the resolution (nerf_loader.h:41) */
	float xform[12];       /* 3x4 camera-to-world, column-major, after the per-ray quaternion round trip of
	                          get_xform_given_rolling_shutter (common_device.cuh:224-234); see ngpb_effective_xform */
	float raw_xform[12];   /* the unmodified 3x4 (used by mark_untrained_density_grid, testbed_nerf.cu:398) */
	int32_t lens_mode;     /* NGPB_LENS_* (ELensMode, common.h; nerf_loader.cu:197-269 read_lens) */
	float lens_params[7];  /* OpenCV: k1, k2, p1, p2; FTheta: p0..p4, w, h (common_device.cuh:141-160,:236-249) */
	int32_t image_type;    /* NGPB_IMAGE_* (EImageDataType; read_rgba, common_device.cuh:677-705) */
} ngpb_image;
/* Byte: RGBA8 sRGB with straight alpha (0x00FF00FF = masked-away pixel); Half / Float: 4 x fp16 / fp32 linear colours with premultiplied alpha, as the
 * reference keeps EXR files and the arrays handed to nerf.training.set_image (a negative red channel = masked away). */
enum { NGPB_IMAGE_BYTE = 0, NGPB_IMAGE_HALF = 1, NGPB_IMAGE_FLOAT = 2 };
enum { NGPB_LENS_PERSPECTIVE = 0, NGPB_LENS_OPENCV = 1, NGPB_LENS_FTHETA = 2, NGPB_LENS_LATLONG = 3 };

/* Host-only: the transform generate_training_samples_nerf effectively uses for a camera without rolling shutter. */
void ngpb_effective_xform(const float* xform12, float* out12);

typedef struct { uint64_t state, inc; } ngpb_rng; /* tcnn pcg32 */

/* ---- K1: generate_training_samples_nerf (src/testbed_nerf.cu:1085-1260) ----
 * Deterministic: samples are laid out in ray order (an exclusive scan replaces the reference's atomicAdd,
 * which makes it one valid serialisation of the reference's allocation order).
 * counters[0] = total requested samples (numsteps_counter), counters[1] = rays kept (ray_counter).
 * scratch: ngpb_generate_training_samples_scratch_bytes(n_rays) bytes (per-ray counts, prefixes and march records). */
uint64_t ngpb_generate_training_samples_scratch_bytes(uint32_t n_rays);
int ngpb_generate_training_samples(void* stream, uint32_t n_rays, const float* aabb6, uint32_t max_samples, ngpb_rng rng,
                                   uint32_t n_images, const ngpb_image* images_dev, const uint8_t* density_grid_bitfield,
                                   int snap_to_pixel_centers, float cone_angle_constant,
                                   uint32_t* counters, uint32_t* ray_indices, float* rays /*[n][6]*/, uint32_t* numsteps /*[n][2]*/,
                                   float* coords /*[max_samples][7]*/, void* scratch);

/* Ray-sharded form for data-parallel training: this call handles the global rays [ray_offset, ray_offset + n_rays) of a batch of
 * n_rays_global rays. Pixel, image and RNG stream of a ray derive from its global index (:1062-1083, :1118-1121), so the shards of a batch
 * together produce exactly the samples of the unsharded call; ray_indices holds global indices. */
int ngpb_generate_training_samples_sharded(void* stream, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, ngpb_rng rng,
                                           uint32_t n_images, const ngpb_image* images_dev, const uint8_t* density_grid_bitfield,
                                           int snap_to_pixel_centers, float cone_angle_constant,
                                           uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch);

/* ---- K6+K7: compute_loss_kernel_train_nerf (:1280-1597) + fill_rollover(_and_rescale) (tcnn common_device.h:517-537) ----
 * counters_in: the device counters written by K1. counters_out[0] = compacted sample count (unclipped).
 * coords_out [batch][7] and dloss_dout [batch][4] are padded to `batch` by rollover.
 * scratch: ngpb_compute_loss_scratch_bytes(n_rays) bytes. */
uint64_t ngpb_compute_loss_scratch_bytes(uint32_t n_rays);
typedef struct {
	float loss_scale;            /* LOSS_SCALE = 128, testbed.h:272 */
	float background_color[3];
	int32_t color_space, random_bg_color, linear_colors, loss_type, rgb_activation, density_activation, snap_to_pixel_centers;
	float near_distance;
} ngpb_loss_config;
int ngpb_compute_loss(void* stream, uint32_t n_rays, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                      uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                      const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                      const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch);

/* Ray-sharded form: n_rays rays of this shard out of n_rays_global; the loss and its gradient are normalised by n_rays_global (:1493), so the
 * sum of the shards' gradients is the gradient of the unsharded batch. */
int ngpb_compute_loss_sharded(void* stream, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                              uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                              const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                              const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch);
/* Same, and additionally compacts the samples' hash-grid feature rows ([n][32] fp16, as written by ngpb_hash_encode_forward on coords_in with the
 * weights of this step) into encoded_out[batch][32] with the same compaction and roll-over as coords_out. The training forward pass on coords_out
 * (NerfNetwork::forward, nerf_network.h:143-185, re-encodes them in the reference) can then start from encoded_out: same function of the same inputs. */
int ngpb_compute_loss_compact_features(void* stream, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                              uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                              const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                              const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                                       const ngpb_half* encoded_in, ngpb_half* encoded_out);

/* ---- K15: Ema(ExponentialDecay(Adam)) in one pass (tcnn adam.h:48-119, ema.h:63-76, exponential_decay.h:60-72) ---- */
typedef struct {
	float learning_rate, beta1, beta2, epsilon, l2_reg, ema_decay;
	uint32_t decay_start, decay_interval; float decay_base;
	uint32_t step; float lr_factor; /* state, advanced by the host wrapper */
} ngpb_optimizer;
void ngpb_optimizer_init(ngpb_optimizer* o); /* configs/nerf/base.json:5-22 */
/* grad is fp32 and is zeroed for the next iteration by the same pass. */
int ngpb_optimizer_step(void* stream, ngpb_optimizer* o, uint32_t n_params, uint32_t n_matrix_params, float loss_scale, float* grad,
                        float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema, float* m1, float* m2, uint32_t* param_steps);

/* ---- K16: occupancy grid (src/testbed_nerf.cu:369-610, :2761-2859) ---- */
int ngpb_mark_untrained_density_grid(void* stream, uint32_t n_elements, float* grid, uint32_t n_images, const ngpb_image* images_dev, int clear_visible);
int ngpb_generate_grid_samples(void* stream, uint32_t n_elements, ngpb_rng rng, uint32_t step, const float* aabb6, const float* grid_in,
                               float* positions3, uint32_t* indices, uint32_t n_cascades, float thresh);
int ngpb_splat_and_ema(void* stream, uint32_t n_samples, const uint32_t* indices, const ngpb_half* density, float* grid_tmp, uint32_t n_elements, float decay, float* grid);
/* mean of max(v,0) over the first cascade -> mean_dev[0], then bitfield + 7 max-pooled mips (2 MiB).
 * mean_dev: NGPB_MEAN_WORKSPACE_BYTES bytes of device memory, 8-byte aligned, owned by the caller (one per testbed / field / stream): the mean in
 * the first float, the block partials of its fixed-order reduction behind it. */
#define NGPB_MEAN_WORKSPACE_BYTES (8 + 1024 * 8)
int ngpb_update_bitfield(void* stream, uint32_t n_cascades_used, const float* grid, float* mean_dev, uint8_t* bitfield);

/* ---- K17: classic single-NeRF render (Testbed::render_nerf, src/testbed_nerf.cu:2354-2499; NerfTracer :2047-2267), ERenderMode::Shade,
 * perspective camera, followed by CudaRenderBuffer::accumulate + tonemap (src/render_buffer.cu:606-660). ---- */
#define NGPB_RENDER_FIRST_PASS_STEPS 4   /* march steps per live ray in the first pass; later passes take more as rays die */
#define NGPB_RENDER_MAX_PASS_STEPS 32
typedef struct {
	int32_t width, height;
	float fx, fy;               /* focal length in pixels (Testbed::calc_focal_length, src/testbed.cu:2589) */
	float screen_center[2];     /* Testbed::render_screen_center(), (0.5, 0.5) by default */
	float camera[12];           /* 3x4 camera-to-world, column-major, ngp convention */
	int32_t spp, snap_to_pixel_centers;
	float aabb[6], render_aabb[6];
	float cone_angle_constant, min_transmittance, near_distance;
	int32_t rgb_activation, density_activation, train_in_linear_colors;
	int32_t color_space;        /* m_color_space: the space the accumulation buffer averages in */
	int32_t output_srgb;        /* !linear argument of Testbed::render */
	float exposure, background_color[4];
	int32_t tonemap_curve;      /* NGPB_TONEMAP_* (m_tonemap_curve, testbed.h:847) */
} ngpb_render_config;
uint64_t ngpb_render_workspace_bytes(uint32_t n_pixels);
/* params: half[10240 + grid] (inference = EMA weights for Testbed::render); bitfield: occupancy bits incl. mips; workspace: device memory of
 * ngpb_render_workspace_bytes(width*height), zero-filled once by the caller (the network passes run over whole 128-sample tiles and read the slots past a
 * wave's last sample). out_rgba_host: float [height][width][4] (host). n_samples_out: samples that went through the
 * network for live rays; n_launches_out: kernels launched. Synchronises the stream (the live-ray count is read back once per pass). */
int ngpb_render_nerf(void* stream, const ngpb_render_config* cfg, const ngpb_grid* g, const ngpb_half* params, const uint8_t* bitfield,
                     void* workspace, float* out_rgba_host, uint64_t* n_samples_out, uint32_t* n_launches_out);

/* ---- K18: the Blender multi-NeRF renderer (NerfRenderer::render, src/nerf_renderer.cu:565-791; Testbed::bl_render_frame, src/testbed.cu:2675) ----
 * A field is one NeRF loaded from a snapshot (NeuralRadianceField, nerf/neural_radiance_field.cuh:153-298): fp16 inference parameters in the
 * reference's flat order, occupancy grid (float, 128^3 per cascade; n_cells may be 0 for an untrained model), aabb_scale. */
typedef struct ngpb_field ngpb_field;
int ngpb_field_create(ngpb_field** out, int device, uint32_t aabb_scale, const ngpb_half* params_host, uint32_t n_params, const float* density_grid_host, uint32_t n_cells);
void ngpb_field_destroy(ngpb_field* f);
typedef struct {            /* Mask3D (nerf/mask_3D.cuh:128-257): an SDF shape that adds or subtracts visibility, with a feathered border */
	int32_t shape;          /* EMaskShape: 0 Box (config = dims xyz), 1 Cylinder (radius, height), 2 Sphere (radius), 3 All */
	int32_t mode;           /* EMaskMode: 0 Add, 1 Subtract */
	float transform[16];    /* 4x4 shape -> NeRF-local (per-NeRF masks) or shape -> world (request masks), column-major */
	float config[6];
	float feather, opacity;
} ngpb_mask;
typedef struct {            /* NerfDescriptor (nerf/nerf_descriptor.cuh) */
	const ngpb_field* field;
	float aabb[6];          /* render box in the NeRF's local frame */
	float transform[16];    /* 4x4 local -> world, column-major */
	float opacity;
	uint32_t n_masks;       /* NerfDescriptor::modifiers.masks, in the NeRF's local frame (host array) */
	const ngpb_mask* masks;
} ngpb_nerf_instance;
enum { NGPB_CAMERA_PERSPECTIVE = 0, NGPB_CAMERA_QUADRILATERAL_HEXAHEDRON = 1, NGPB_CAMERA_SPHERICAL_QUADRILATERAL = 2 }; /* ECameraModel, camera_models.cuh:27-31 */
enum { NGPB_TONEMAP_IDENTITY = 0, NGPB_TONEMAP_ACES = 1, NGPB_TONEMAP_HABLE = 2, NGPB_TONEMAP_REINHARD = 3 };                 /* ETonemapCurve, common.h:136-141 */
typedef struct {            /* RenderRequest: RenderOutputProperties + RenderCameraProperties (nerf/render_request.cuh), perspective camera */
	int32_t width, height;  /* output resolution */
	int32_t mip;            /* DownsampleInfo::MakeFromMip(resolution, mip): every 2^mip-th pixel is traced and splatted */
	int32_t flip_y;
	float camera[12];       /* 3x4 camera-to-world, column-major */
	float focal_length;     /* pixels */
	float near_distance;
	int32_t color_space;    /* NGPB_COLOR_* : accumulation and output space */
	float exposure, background_color[4];
	int32_t camera_model;   /* NGPB_CAMERA_* (RenderCameraProperties::model) */
	float aperture_size, focus_z;            /* depth of field: thin-lens blur of the ray origin when aperture_size > 0 (camera_models.cuh:99-104,:196-201,:232-237) */
	float spherical_quadrilateral[3];        /* width, height, curvature (camera_models.cuh:119-135) */
	float quadrilateral_hexahedron[24];      /* front tl, tr, bl, br then back tl, tr, bl, br, xyz each (camera_models.cuh:33-80) */
	int32_t tonemap_curve;  /* NGPB_TONEMAP_* (RenderOutputProperties::tonemap_curve) */
	uint32_t n_masks;       /* RenderRequest::modifiers.masks, in world space (host array); applied to every NeRF */
	const ngpb_mask* masks;
} ngpb_blender_request;
/* out_rgba_host: float [height][width][4]. Synchronises the stream once per wave (the reference's schedule, :704-705). */
int ngpb_blender_render(void* stream, const ngpb_blender_request* rq, uint32_t n_nerfs, const ngpb_nerf_instance* nerfs, float* out_rgba_host,
                        uint64_t* n_samples_out, uint32_t* n_launches_out);

/* ---- development self-test of the tcgen05 building blocks (tests/test_umma_selftest.py) ---- */
int ngpb_selftest_umma(void* stream, int variant, const ngpb_half* a, const ngpb_half* b, float* d);

/* =====================================================================================================
 * Testbed level: host-facing mirror of pyngp.Testbed for the NeRF mode (src/python_api.cu:540-732).
 * ===================================================================================================== */
typedef struct ngpb_testbed ngpb_testbed;

typedef struct {
	const uint8_t* pixels; /* host pointer, w*h*{4, 8, 16} bytes by image_type */
	int32_t w, h;
	float fx, fy, cx, cy;
	float xform[12];       /* 3x4 camera-to-world, column-major, ngp convention (after nerf_matrix_to_ngp, nerf_loader.h:113) */
	int32_t lens_mode;     /* NGPB_LENS_* */
	float lens_params[7];
	int32_t image_type;    /* NGPB_IMAGE_* */
} ngpb_host_image;

int ngpb_testbed_create(ngpb_testbed** out, int device);
void ngpb_testbed_destroy(ngpb_testbed* t);
/* Testbed::load_training_data (src/testbed.cu:97) for already-decoded images; aabb_scale as in transforms.json. */
int ngpb_testbed_load_training_data(ngpb_testbed* t, uint32_t n_images, const ngpb_host_image* images, uint32_t aabb_scale);
/* Datasets filled from the caller's arrays while training runs (python_api.cu:545, :56-76; src/testbed_nerf.cu:2502-2516, :2635-2641):
 * Testbed::create_empty_nerf_dataset allocates n_images empty slots and sets n_images_for_training = 0 (train() then returns without a step, :2897);
 * nerf.training.set_image replaces one slot's pixels (any resolution / NGPB_IMAGE_* type; only pixels, w, h, image_type of `image` are read);
 * set_camera_intrinsics follows the reference's conventions (fx or fy <= 0: copied from the other; cx, cy >= 0 in pixels, negative = minus the fraction;
 * any of k1 k2 p1 p2 non-zero selects the OpenCV lens). Option "n_images_for_training" (ngpb_testbed_set_option): the first n slots take part in training;
 * the occupancy grid is re-marked when it changes (:2783-2799). Extrinsics: ngpb_testbed_set_camera_extrinsics. */
int ngpb_testbed_create_empty_dataset(ngpb_testbed* t, uint32_t n_images, uint32_t aabb_scale);
int ngpb_testbed_set_training_image(ngpb_testbed* t, uint32_t frame_idx, const ngpb_host_image* image);
int ngpb_testbed_set_camera_intrinsics(ngpb_testbed* t, uint32_t frame_idx, float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2);
/* Testbed::reset_network (src/testbed.cu:2249) with configs/nerf/base.json and the given seed (m_seed, testbed.h:567). */
int ngpb_testbed_reset_network(ngpb_testbed* t, uint32_t seed);
/* Testbed::train(batch_size) (src/testbed.cu:2527): exactly one optimizer step. */
int ngpb_testbed_train(ngpb_testbed* t, uint32_t batch_size);
/* Runs `n_steps` iterations without host round trips in between (device-side batch-size controller). */
int ngpb_testbed_train_n(ngpb_testbed* t, uint32_t batch_size, uint32_t n_steps);
float ngpb_testbed_loss(ngpb_testbed* t);
uint32_t ngpb_testbed_training_step(const ngpb_testbed* t);
/* stats: [0] rays_per_batch of the last step, [1] measured_batch_size_before_compaction, [2] measured_batch_size, [3] number of kernel launches so far */
int ngpb_testbed_stats(ngpb_testbed* t, uint64_t* stats4);
uint32_t ngpb_testbed_n_params(const ngpb_testbed* t);
int ngpb_testbed_get_params(ngpb_testbed* t, float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema);
int ngpb_testbed_set_params(ngpb_testbed* t, const float* w_fp32);
int ngpb_testbed_get_density_grid(ngpb_testbed* t, float* grid, uint8_t* bitfield);
/* ---- data-parallel training over the GPUs of one box (the reference is single-GPU, README.md:239-241) ----
 * One process per GPU. Every rank loads the same dataset and resets the network with the same seed (bit-identical replicas), then:
 *   rank 0: ngpb_nccl_unique_id(id); broadcast the 128 bytes to all ranks by any means (bench.py uses torch.distributed);
 *   all:    ngpb_testbed_init_data_parallel(t, rank, world, id).
 * Afterwards each ngpb_testbed_train step marches this rank's shard of a global batch of world x rays_per_batch rays to `batch_size`
 * compacted samples, sums the batch-size counters and exchanges the gradients over NCCL: by default the partial gradients are rounded to bf16 and
 * reduce-scattered, every rank runs Adam on its 1/world slice of the parameters and the fp16 weights are all-gathered (options "dp_half_gradients"
 * 0 = fp32 exchange / 1 = bf16 / 2 = fp16, "dp_sharded_optimizer" 0 = plain fp32 all-reduce + full Adam on every rank, "dp_exchange" 1 = the
 * library's own peer-memory kernels instead of NCCL collectives). Every rank ends a step with identical weights. With "optimize_extrinsics" the
 * per-camera gradients are all-reduced before the host-side Adam. NCCL is resolved with dlopen("libnccl.so.2") at the first call. */
/* Host-only arithmetic of the sharded batch (what train() uses): the next ray count from the SUMMED compacted-sample count (every rank derives the
 * same value; NerfCounters::update_after_training, src/testbed_nerf.cu:2890-2891), and a rank's slice of the global ray batch. */
uint32_t ngpb_next_rays_per_batch(uint32_t rays_per_batch, uint32_t batch, uint32_t global_compacted, uint32_t world);
void ngpb_ray_shard(uint32_t rank, uint32_t world, uint32_t rays_per_batch, uint32_t* ray_offset, uint32_t* n_rays_global);
int ngpb_nccl_unique_id(void* out128);
int ngpb_testbed_init_data_parallel(ngpb_testbed* t, int rank, int world, const void* unique_id128);

/* ---- camera-extrinsics optimisation (K13 / K14; nerf.training.optimize_extrinsics, python_api.cu:806) ----
 * Gradient of the loss with respect to the network inputs of n samples: coords_gradient [n][7] float = {dL/dpos[3], 0, dL/ddir[3]} in the
 * coordinates' own [0,1] units. Position part: kernel_grid's dy_dx + kernel_grid_backward_input (tcnn encodings/grid.h:351-392,:551-575) from
 * dL_dencoded [n][32] half and the grid; direction part: kernel_sh_backward (encodings/spherical_harmonics.h:154-390) from dL_dsh [n][16] half
 * (NULL: zero). grid: half2 table of all levels (the parameters after the 10240 MLP weights). */
int ngpb_nerf_input_gradient(void* stream, const ngpb_grid* g, const ngpb_half* grid, const float* coords, uint32_t n, const ngpb_half* dL_dencoded,
                             const ngpb_half* dL_dsh, float* coords_gradient);
/* compute_cam_gradient_train_nerf (src/testbed_nerf.cu:1600-1707), uniform pixel sampling: per kept ray, sums the sample gradients into a ray
 * origin / direction gradient and accumulates (atomicAdd) into cam_pos_gradient / cam_rot_gradient [n_images][3] (either may be NULL). numsteps
 * [max_rays][2] = {compacted sample count, compacted base} as ngpb_compute_loss leaves it; coords / coords_gradient: the compacted batch. */
int ngpb_compute_cam_gradient(void* stream, uint32_t max_rays, uint32_t n_rays_global, const float* aabb6, const uint32_t* rays_counter_dev, uint32_t n_images,
                              const uint32_t* ray_indices, const float* rays, const uint32_t* numsteps, const float* coords, const float* coords_gradient,
                              float* cam_pos_gradient, float* cam_rot_gradient);
/* Host-side: one step of AdamOptimizer<Vector3f> (rotation = 0) or RotationAdamOptimizer (rotation = 1; the variable is an angle-axis vector and the
 * update is composed as a rotation) (include/neural-graphics-primitives/adam_optimizer.h:20-159; beta1 .9, beta2 .99, eps 1e-8).
 * state10 = {iter, first_moment[3], second_moment[3], variable[3]}. */
void ngpb_camera_adam_step(float* state10, const float* gradient3, float learning_rate, int rotation);
/* Host-side: Training::update_transforms for one camera (src/testbed_nerf.cu:2597-2633): out = dataset transform (3x4 column-major) with the rotation
 * offset (angle-axis) applied on the left of its 3x3 block and the position offset added to its translation. */
void ngpb_apply_camera_offsets(const float* xform12, const float* pos_offset3, const float* rot_offset3, float* out12);
/* Host-side: the exposure block of Testbed::train_nerf (src/testbed_nerf.cu:3105-3131): one AdamOptimizer<Array3f> step per image on
 * gradient * per_camera_loss_scale + l2_reg * exposure, then the exposures are re-centred on a zero mean. states [n_images][10] as
 * ngpb_camera_adam_step, gradients [n_images][3] (what K6 accumulated over the last n_steps_between_cam_updates steps). */
void ngpb_exposure_update(uint32_t n_images, float* states, const float* gradients, float per_camera_loss_scale, float l2_reg, float learning_rate);
/* K6 with per-image exposures (src/testbed_nerf.cu:1326-1327,:1403,:1558-1571): exposure_dev [n_images][3] (in stops; the target colour of a ray
 * is its image's colour x 2^exposure) and, when exposure_gradient_dev [n_images][3] is given, the gradient of the loss w.r.t. the exposures is
 * accumulated into it (atomicAdd, scaled by loss_scale / n_rays_global like the network gradients). Otherwise as ngpb_compute_loss_sharded. */
int ngpb_compute_loss_exposure(void* stream, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                               uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                               const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                               const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                               const float* exposure_dev, float* exposure_gradient_dev);
/* ---- K19: importance sampling of training pixels / images by accumulated error (Testbed::Nerf::Training::ErrorMap, testbed.h:600-615;
 * nerf.training.sample_focal_plane_proportional_to_error / sample_image_proportional_to_error, python_api.cu:817-818) ----
 * The CDFs (device pointers): cdf_x_cond_y [n_images][res_y][res_x] + cdf_y [n_images][res_y] choose a texel of the image's error map for half of
 * the rays (sample_cdf_2d, src/testbed_nerf.cu:991-1022), cdf_img [n_images] chooses the image (image_idx, :1062-1083). Null members switch the
 * respective choice back to uniform pixels / round-robin images. */
typedef struct {
	const float* cdf_x_cond_y;
	const float* cdf_y;
	const float* cdf_img;
	int32_t res_x, res_y;
} ngpb_error_cdf;
/* K1 drawing its pixels / images from the CDFs; otherwise ngpb_generate_training_samples_sharded (error_cdf may be NULL). */
int ngpb_generate_training_samples_cdf(void* stream, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, ngpb_rng rng,
                                       uint32_t n_images, const ngpb_image* images_dev, const uint8_t* density_grid_bitfield,
                                       int snap_to_pixel_centers, float cone_angle_constant,
                                       uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch,
                                       const ngpb_error_cdf* error_cdf);
/* K6 recovering its pixels from the same CDFs (error_cdf, may be NULL): the per-ray loss -- not its gradient -- is divided by the sampling density
 * (:1448-1458) and the exposure gradient by the pixel density (:1560). error_map_dev (may be NULL) [n_images][error_map_res_y][error_map_res_x]: every
 * ray with a compacted sample deposits its loss bilinearly (atomicAdd, :1465-1491; no sharpness weighting). Otherwise ngpb_compute_loss_exposure. */
int ngpb_compute_loss_error_map(void* stream, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                                uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                                const float* exposure_dev, float* exposure_gradient_dev,
                                const ngpb_error_cdf* error_cdf, float* error_map_dev, int32_t error_map_res_x, int32_t error_map_res_y);
/* construct_cdf_2d + construct_cdf_1d (:1984-2037) and the image normalisation the reference does on the host (:3000-3015), all on the device:
 * per row the running sum of error + 1e-10, normalised and blended with 1 % uniform; the same over the row sums per image; then over the image sums
 * with 10 % uniform. cdf_x_cond_y [n_images][res_y][res_x], cdf_y [n_images][res_y], cdf_img [n_images], pmf_img [n_images] (may be NULL: the
 * per-image sampling probabilities, pmf_img_cpu). The sums run serially in the reference's order (bit-exact CDFs); one thread per row / image. */
int ngpb_construct_error_cdfs(void* stream, uint32_t n_images, uint32_t res_y, uint32_t res_x, const float* error_map_dev,
                              float* cdf_x_cond_y, float* cdf_y, float* cdf_img, float* pmf_img);
/* Options "sample_focal_plane_proportional_to_error" and "sample_image_proportional_to_error" (ngpb_testbed_set_option) switch the error map on inside
 * ngpb_testbed_train with the reference's cadence (:2933-2939, :2971-3023: map re-sized and cleared at the start of a window, CDFs rebuilt after
 * n_steps_between_error_map_updates steps, which then grows by 1.5x from 128); read-only options "error_map_res", "error_cdf_valid",
 * "n_steps_between_error_map_updates". pmf_img [n_images]: sampling probabilities of the images after the last CDF update. */
int ngpb_testbed_get_error_map_pmf(ngpb_testbed* t, float* pmf_img);
/* Options "optimize_exposure" and "exposure_l2_reg" (ngpb_testbed_set_option) switch the per-image exposure optimisation on inside ngpb_testbed_train
 * (same 16-step cadence as the extrinsics; Adam at the network optimizer's learning rate). The reference shows the exposures only in its GUI; these two
 * calls read / replace them: exposures3 [n_images][3]. Setting exposures resets their optimizer state. */
int ngpb_testbed_get_camera_exposures(ngpb_testbed* t, float* exposures3);
int ngpb_testbed_set_camera_exposures(ngpb_testbed* t, const float* exposures3);
/* Options "optimize_extrinsics", "extrinsic_learning_rate", "extrinsic_l2_reg", "n_steps_between_cam_updates" (ngpb_testbed_set_option) switch the
 * per-camera optimisation on inside ngpb_testbed_train. The current training transform of a frame (dataset transform + learned offsets, ngp coordinates),
 * and the offsets themselves (any pointer may be NULL): */
int ngpb_testbed_get_camera_extrinsics(ngpb_testbed* t, uint32_t frame_idx, float* xform12, float* pos_offset3, float* rot_offset3);
/* Training::set_camera_extrinsics (:2539): replaces the dataset transform of a frame (ngp coordinates); reset_camera_extrinsics (:2543) zeroes every
 * camera's offsets and Adam state. */
int ngpb_testbed_set_camera_extrinsics(ngpb_testbed* t, uint32_t frame_idx, const float* xform12);
int ngpb_testbed_reset_camera_extrinsics(ngpb_testbed* t);

/* ---- neural-image and SDF modes (ETestbedMode::Image / Sdf): NetworkWithInputEncoding + Trainer around the same kernels ----
 * Replaces Testbed::reset_network for these modes (src/testbed.cu:2244-2470), tcnn Trainer::training_step / optimizer_step (trainer.h:108-190),
 * Testbed::train_image / render_image / compute_image_mse (src/testbed_image.cu:220-523) and Testbed::train_sdf (src/testbed_sdf.cu:1229-1252) on
 * supplied (position, distance) pairs. Parameters in NetworkWithInputEncoding's flat order: network (7168), then the grid levels. */
typedef struct ngpb_model ngpb_model;
typedef struct {
	uint32_t n_pos_dims;        /* 2 neural image, 3 SDF */
	uint32_t n_output_dims;     /* 3 / 1 */
	uint32_t n_levels, log2_hashmap_size, base_resolution; /* encoding section */
	float per_level_scale;      /* > 0: taken as is; otherwise derived like reset_network does (src/testbed.cu:2318-2322) from ... */
	float desired_resolution;   /* ... the finest level's target resolution: 2048 (SDF), half the image's larger side (image) */
	int32_t loss;               /* NGPB_ELEMENT_LOSS_* */
	int32_t use_ema;            /* optimizer is Ema(...) (configs/sdf/base.json) or not (configs/image/base.json) */
	ngpb_optimizer optimizer;   /* hyper-parameters; step / lr_factor ignored */
	uint32_t seed;              /* m_seed (1337) */
} ngpb_model_config;
int ngpb_model_create(ngpb_model** out, int device, const ngpb_model_config* cfg);
void ngpb_model_destroy(ngpb_model* m);
int ngpb_model_reset(ngpb_model* m, uint32_t seed);                 /* reset_network: fresh parameters, optimizer and RNG */
uint32_t ngpb_model_n_params(const ngpb_model* m);
uint32_t ngpb_model_training_step(const ngpb_model* m);
float ngpb_model_loss(const ngpb_model* m);                         /* the last step trained with get_loss != 0: Trainer::loss = sum of the per-element values */
uint64_t ngpb_model_launches(const ngpb_model* m);
void* ngpb_model_stream(ngpb_model* m);
int ngpb_model_set_option(ngpb_model* m, const char* name, double value); /* "snap_to_pixel_centers", "linear_colors" (m_image.training.*), "learning_rate" */
int ngpb_model_get_params(ngpb_model* m, float* w_fp32, ngpb_half* w_half, ngpb_half* w_inference);
int ngpb_model_set_params_half(ngpb_model* m, const ngpb_half* params, uint32_t n);
int ngpb_model_set_training_step(ngpb_model* m, uint32_t step);
/* One Trainer::training_step on a caller-supplied device batch (positions [n][n_pos_dims], targets [n][n_output_dims], n a multiple of 128). */
int ngpb_model_train(ngpb_model* m, const float* positions_dev, const float* targets_dev, uint32_t n, int run_optimizer, int get_loss);
/* Network::inference: positions [n][n_pos_dims] -> out [n][n_output_dims] fp32 (device pointers, any n). */
int ngpb_model_inference(ngpb_model* m, const float* positions_dev, uint32_t n, float* out_dev, int use_inference_params);
/* Neural image. pixels: RGBA float (is_half 0) or RGBA half (1), linear, row-major, host. */
int ngpb_model_set_image(ngpb_model* m, const void* pixels_host, int width, int height, int is_half);
int ngpb_model_set_image_rgba8(ngpb_model* m, const uint8_t* pixels_host, int width, int height); /* an 8-bit file: sRGB -> linear, premultiplied, on the device (from_rgba32) */
int ngpb_model_train_image(ngpb_model* m, uint32_t batch, int get_loss);
int ngpb_model_image_mse(ngpb_model* m, int quantize_to_byte, float* mse_out);
/* view5 = {m_scale, m_image.pos x, y, m_screen_center x, y} (defaults 1, 0, 0, 0.5, 0.5); out: width x height RGBA float, host. */
int ngpb_model_render_image(ngpb_model* m, int width, int height, int spp, const float* view5, int render_snap_to_pixel_centers, int color_space, int output_srgb,
                            float exposure, const float* background4, int tonemap_curve, float* out_rgba_host);
/* SDF on supplied pairs (positions already in the unit cube, distances in its units): the pool, and one train_sdf step drawn from it. */
int ngpb_model_set_sdf_data(ngpb_model* m, const float* positions_host, const float* distances_host, uint32_t n);
int ngpb_model_train_sdf(ngpb_model* m, uint32_t batch, int get_loss);
/* The batch the last training step used (device -> host), for parity tests. */
int ngpb_model_get_training_batch(ngpb_model* m, uint32_t n, float* positions_host, float* targets_host);

/* Development aid (tools/umma_probe.py): cycle counts of tcgen05 issue / commit / wait sequences on this GPU; sections is a bit mask. */
int ngpb_probe_umma(void* stream, long long* cycles_host, uint32_t n, uint32_t sections);

/* The stream every testbed kernel is launched on (Testbed::m_stream, testbed.h:895), as a cudaStream_t. */
void* ngpb_testbed_stream(ngpb_testbed* t);
/* Per-stage device time of train(), measured with CUDA events on that stream while the option "profile_stages" is 1.
 * Stage indices: */
enum { NGPB_STAGE_SAMPLING = 0,      /* K1  generate_training_samples                        units: rays      */
       NGPB_STAGE_ENCODE_INFERENCE,  /* K2  hash encode of the uncompacted samples            units: samples   */
       NGPB_STAGE_MLP_INFERENCE,     /* K3  MLP forward of the uncompacted samples            units: samples   */
       NGPB_STAGE_LOSS,              /* K6+K7 compositing, loss, compaction, roll-over        units: rays      */
       NGPB_STAGE_ENCODE_TRAIN,      /* K2  hash encode of the compacted batch                units: samples   */
       NGPB_STAGE_MLP_TRAIN,         /* K8-K11 MLP forward + backward + weight gradients      units: samples   */
       NGPB_STAGE_ENCODE_BACKWARD,   /* K12 hash-grid scatter-add                             units: samples   */
       NGPB_STAGE_OPTIMIZER,         /* K15 Adam + EMA + gradient reset                       units: params    */
       NGPB_STAGE_DENSITY_GRID,      /* K16 occupancy-grid refresh                            units: samples   */
       NGPB_STAGE_ALLREDUCE,         /* gradient all-reduce (multi-GPU only)                  units: bytes     */
       NGPB_N_STAGES };
/* ms[NGPB_N_STAGES], calls[NGPB_N_STAGES], units[NGPB_N_STAGES]: accumulated since the last reset (reset != 0 clears). */
int ngpb_testbed_stage_times(ngpb_testbed* t, double* ms, uint64_t* calls, uint64_t* units, int reset);
/* ---- snapshot state (Testbed::save_snapshot / load_snapshot, src/testbed.cu:3008-3106; tcnn Trainer::serialize / deserialize, trainer.h:270-310).
 * The .msgpack container itself is written / read by the host binding (pyngp.save_snapshot / load_snapshot); these calls move the state. ---- */
/* Configures the scene box like load_nerf_post (src/testbed_nerf.cu:2714-2730) WITHOUT a dataset and resets the network: what load_snapshot does
 * for a render-only session. */
int ngpb_testbed_configure(ngpb_testbed* t, uint32_t aabb_scale, uint32_t seed);
/* Trainer::deserialize for params_type "__half": the fp16 snapshot parameters become the training, inference (EMA) and fp32 master copies. */
int ngpb_testbed_set_params_half(ngpb_testbed* t, const ngpb_half* params, uint32_t n_params);
/* density_grid: float[128^3 * (max_cascade + 1)] (host). Uploads it and recomputes mean + bitfield (update_density_grid_mean_and_bitfield, :2844). */
int ngpb_testbed_set_density_grid(ngpb_testbed* t, const float* density_grid, uint32_t n_cells);
typedef struct {
	uint32_t training_step, rays_per_batch, measured_batch_size, measured_batch_size_before_compaction;
	float loss;
	uint32_t optimizer_step;          /* Adam current_step (adam.h:284) */
	float learning_rate, learning_rate_factor; /* ExponentialDecay state (exponential_decay.h:139-140) */
} ngpb_training_state;
int ngpb_testbed_get_training_state(ngpb_testbed* t, ngpb_training_state* out);
int ngpb_testbed_set_training_state(ngpb_testbed* t, const ngpb_training_state* in);
/* Adam moments and per-parameter step counters (host buffers of n_params elements each; any may be NULL). */
int ngpb_testbed_get_optimizer_state(ngpb_testbed* t, float* first_moments, float* second_moments, uint32_t* param_steps);
int ngpb_testbed_set_optimizer_state(ngpb_testbed* t, const float* first_moments, const float* second_moments, const uint32_t* param_steps);

/* training options exposed as pyngp properties (python_api.cu:650-852) */
int ngpb_testbed_set_option(ngpb_testbed* t, const char* name, double value);
double ngpb_testbed_get_option(ngpb_testbed* t, const char* name);
/* Testbed::render (python_api.cu:132-190) for the classic single-NeRF path: camera12 = 3x4 column-major camera matrix.
 * out_rgba: host float [h][w][4]. n_samples_out (optional): network-evaluated samples. */
int ngpb_testbed_render(ngpb_testbed* t, const float* camera12, int w, int h, float fx, float fy, int spp, int linear, float* out_rgba, uint64_t* n_samples_out);
/* Device milliseconds (CUDA events on the testbed's stream) of the last ngpb_testbed_render call, excluding the final device-to-host copy. */
double ngpb_testbed_last_render_ms(const ngpb_testbed* t);

#ifdef __cplusplus
}
#endif
#endif /* NGPB_H */
