// TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("oracle") of the reference's NeRF train/render hot path. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library, and only as the checker / reported CPU baseline -- never on the product path.
//
// Parity pinning: the reference holds no golden vectors or tests for this path.
// The restatement is pinned instead against outputs of the reference's own kernels compiled from
// /root/reference (oracle/ref_harness/*.cu -> oracle/_ref/*.so) and run on a B200 by
// oracle/gen_golden.py; the resulting vectors are committed under tests/golden/.
//
// Every function cites the reference file:line it follows. Paths are relative to /root/reference;
// "tcnn/" = dependencies/tiny-cuda-nn/.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t orc_half; // IEEE binary16 bit pattern

// ---- RNG: tcnn/dependencies/pcg32/pcg32.h --------------------------------------------------
typedef struct { uint64_t state, inc; } orc_pcg32;
void     orc_pcg32_seed(orc_pcg32* rng, uint64_t initstate, uint64_t initseq);
uint32_t orc_pcg32_next_uint(orc_pcg32* rng);
float    orc_pcg32_next_float(orc_pcg32* rng);
void     orc_pcg32_advance(orc_pcg32* rng, int64_t delta);

// ---- hash grid: tcnn/include/tiny-cuda-nn/encodings/grid.h ----------------------------------
// offsets has n_levels+1 entries (grid.h:985-1018). Returns total number of grid entries.
// `scales` (nullable) overrides the per-level grid_scale (grid.h:194-199), which the reference evaluates with the
// device's exp2f; pass the device's values to compare index-exactly with reference output, NULL to use the host's.
uint32_t orc_grid_offsets(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, uint32_t* offsets);
// corner indices for one level, used by the integer-exact tests (grid.h:164-186 + common_device.h:402-445).
void orc_grid_indices(uint32_t n, uint32_t level, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                      const float* positions, uint32_t pos_stride, uint32_t* indices8, float* weights8);
// encoded: [n][2*n_levels] half (sample-major). grid.h:220-349.
void orc_grid_forward(uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                      const orc_half* grid, const float* positions, uint32_t pos_stride, orc_half* encoded);
// dL_dy: [n][2*n_levels] half. grad: float[2*offsets[n_levels]], overwritten. grid.h:395-518.
void orc_grid_backward(uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                       const float* positions, uint32_t pos_stride, const orc_half* dL_dy, float* grad);

// ---- SH degree 4: tcnn/include/tiny-cuda-nn/encodings/spherical_harmonics.h:46-150 ----------
void orc_sh4(uint32_t n, const float* dirs, uint32_t dir_stride, orc_half* out, uint32_t out_stride);

// ---- NeRF network (nerf_network.h:103-137; tcnn/src/fully_fused_mlp.cu:500-557) -------------
// mlp params (half) in the reference's flat order (nerf_network.h:361-394):
//   density W1[64][32], W2[16][64]; rgb W1[64][32], W2[64][64], W3[16][64]   (row-major [out][in])
#define ORC_MLP_PARAMS 10240
// encoded [n][32] half, coords [n][7] float -> rgbsigma [n][4] half. Optional saved activations
// (each [n][64] half, post-ReLU) and rgb_in [n][32] half, for the backward pass.
void orc_nerf_mlp_forward(uint32_t n, const orc_half* mlp, const orc_half* encoded, const float* coords,
                          orc_half* rgbsigma, orc_half* act_h1, orc_half* rgb_in, orc_half* act_g1, orc_half* act_g2);
// Backward (nerf_network.h:187-266; fully_fused_mlp.cu:151-314,:759-850): dL_dout [n][4] half ->
// dL_dencoded [n][32] half and mlp gradient (float[10240], overwritten).
void orc_nerf_mlp_backward(uint32_t n, const orc_half* mlp, const orc_half* encoded, const float* coords, const orc_half* dL_dout,
                           orc_half* dL_dencoded, float* mlp_grad);

// ---- model description shared by the composite entry points ---------------------------------
typedef struct {
	uint32_t n_levels, log2_hashmap_size, base_resolution;
	float per_level_scale;
	uint32_t offsets[33];
	float scales[32];
	uint32_t n_grid_params; // 2 * offsets[n_levels]
} orc_model;
void orc_model_init(orc_model* m, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale);
// params: half[10240 + n_grid_params] (mlp then grid). coords [n][7] -> rgbsigma [n][4] half.
void orc_nerf_inference(const orc_model* m, const orc_half* params, uint32_t n, const float* coords, orc_half* rgbsigma);
// density network only (nerf_network.h:268-284): positions [n][3] -> density logit half[n].
void orc_nerf_density(const orc_model* m, const orc_half* params, uint32_t n, const float* positions, uint32_t pos_stride, orc_half* density);
// forward+backward on a (padded) batch; grad float[10240 + n_grid_params] overwritten.
void orc_nerf_forward_backward(const orc_model* m, const orc_half* params, uint32_t n, const float* coords, const orc_half* dL_dout, float* grad);

// ---- dataset ---------------------------------------------------------------------------------
typedef struct {
	const uint8_t* pixels; // row-major, w*h*{4, 8, 16} bytes by image_type
	int32_t w, h;
	float fx, fy, cx, cy;  // focal length in pixels, principal point as a fraction (nerf_loader.h:41-42)
	float xform[12];       // 3x4 camera-to-world, column-major, ngp convention (nerf_loader.h:113-132)
	int32_t lens_mode;     // ELensMode {Perspective, OpenCV, FTheta, LatLong} (common.h)
	float lens_params[7];  // OpenCV: k1, k2, p1, p2; FTheta: p0..p4, w, h
	int32_t image_type;    // EImageDataType {0 Byte, 1 Half, 2 Float} (read_rgba, common_device.cuh:677-705); Half / Float: linear, premultiplied alpha
} orc_image;

// effective per-image transform the reference derives per ray through a quaternion round trip
// (common_device.cuh:224-234, no rolling shutter / motion blur).
void orc_effective_xform(const float* xform12, float* out12);

// ---- K1: src/testbed_nerf.cu:1085-1260 (sequential in ray order) ----------------------------
// Returns number of rays kept. counters_out[0] = total samples requested (numsteps_counter),
// counters_out[1] = rays kept. rays [n][6], numsteps [n][2], coords [max_samples][7].
uint32_t orc_generate_training_samples(
	uint32_t n_rays, const float* aabb6, uint32_t max_samples, uint32_t n_rays_total, orc_pcg32 rng,
	uint32_t n_images, const orc_image* images, const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t* counters_out);

// Rays [ray_offset, ray_offset + n_rays) of a batch of n_rays_global rays (ray_indices holds global indices): the data-parallel shard.
uint32_t orc_generate_training_samples_sharded(
	uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, orc_pcg32 rng,
	uint32_t n_images, const orc_image* images, const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t* counters_out);

// ---- K6: src/testbed_nerf.cu:1280-1597 (sequential in ray order) ----------------------------
// Returns the compacted sample count (unclipped numsteps_counter_compacted).
uint32_t orc_compute_loss(
	uint32_t n_rays_kept, uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, orc_pcg32 rng, uint32_t max_samples_compacted,
	float loss_scale, const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, const orc_image* images, const orc_half* rgbsigma /*[.][4]*/, const uint32_t* ray_indices, const float* rays,
	uint32_t* numsteps, const float* coords_in, float* coords_out, orc_half* dloss_dout /*[.][4]*/, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, float mean_density, float near_distance);
// The same with per-image exposures (3 floats per image, null = zero; :1403) and, when `exposure_gradient` is given, the accumulated
// d loss / d exposure per image (:1558-1571) that Testbed::train_nerf hands to the per-image Adam optimisers (:3105-3131).
uint32_t orc_compute_loss_exposure(
	uint32_t n_rays_kept, uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, orc_pcg32 rng, uint32_t max_samples_compacted,
	float loss_scale, const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, const orc_image* images, const orc_half* rgbsigma /*[.][4]*/, const uint32_t* ray_indices, const float* rays,
	uint32_t* numsteps, const float* coords_in, float* coords_out, orc_half* dloss_dout /*[.][4]*/, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, float mean_density, float near_distance,
	const float* exposure, float* exposure_gradient);
// ---- K19: error-map importance sampling (src/testbed_nerf.cu:991-1083, :1465-1491, :1982-2037, :2933-3023) ----
// The CDFs K1 / K6 sample pixels (x_cond_y [n_images][res_y][res_x], y [n_images][res_y]) and images (img [n_images]) from; null pointers switch
// the respective importance sampling off. Global state of the restatement: set, call orc_generate_training_samples* / orc_compute_loss*, clear.
void orc_set_error_cdf(const float* cdf_x_cond_y, const float* cdf_y, const float* cdf_img, int res_x, int res_y);
// The error map K6 deposits every ray's loss into ([n_images][res_y][res_x], added to); null = no accumulation.
void orc_set_error_map(float* error_map, int res_x, int res_y);
void orc_construct_cdfs(uint32_t n_images, uint32_t height, uint32_t width, const float* error_map, float* cdf_x_cond_y, float* cdf_y, float* cdf_img_sums);
void orc_normalize_image_cdf(uint32_t n_images, const float* image_sums, float* pmf_img, float* cdf_img);
// K7: tcnn common_device.h:517-537
void orc_fill_rollover(uint32_t n_target, uint32_t n_valid, float* coords /*[.][7]*/, orc_half* dloss_dout /*[.][4]*/);

// ---- K15: tcnn adam.h:48-119, ema.h:63-76, exponential_decay.h:60-72 ------------------------
typedef struct {
	float learning_rate, beta1, beta2, epsilon, l2_reg, ema_decay;
	uint32_t decay_start, decay_interval; float decay_base;
	uint32_t step; float lr_factor;
} orc_optimizer;
void orc_optimizer_init(orc_optimizer* o);
void orc_optimizer_step(orc_optimizer* o, uint32_t n_params, uint32_t n_matrix_params, float loss_scale, const float* grad,
                        float* w_fp32, orc_half* w_half, orc_half* w_ema, float* m1, float* m2, uint32_t* param_steps);

// ---- K16: src/testbed_nerf.cu:369-610,:2761-2859 --------------------------------------------
void orc_mark_untrained_density_grid(uint32_t n_elements, float* grid, uint32_t n_images, const orc_image* images, int clear_visible);
void orc_generate_grid_samples(uint32_t n_elements, orc_pcg32 rng, uint32_t step, const float* aabb6, const float* grid_in,
                               float* positions3, uint32_t* indices, uint32_t n_cascades, float thresh);
void orc_splat_and_ema(uint32_t n_samples, const uint32_t* indices, const orc_half* density, uint32_t n_elements, float decay, float* grid);
float orc_density_grid_mean(const float* grid);
void orc_bitfield(uint32_t n_cascades_used, const float* grid, float mean_density, uint8_t* bitfield);

// ---- whole training state + one iteration (Testbed::train, src/testbed.cu:2527; train_nerf :2896) ---
typedef struct orc_trainer orc_trainer;
orc_trainer* orc_trainer_create(uint32_t n_images, const orc_image* images, uint32_t aabb_scale, uint32_t seed);
void orc_trainer_destroy(orc_trainer* t);
uint32_t orc_trainer_n_params(const orc_trainer* t);
void orc_trainer_get_params(const orc_trainer* t, float* w_fp32, orc_half* w_half, orc_half* w_ema);
void orc_trainer_set_params(orc_trainer* t, const float* w_fp32);
const uint8_t* orc_trainer_bitfield(const orc_trainer* t);
const float* orc_trainer_density_grid(const orc_trainer* t);
// One Testbed::train(batch) call. stats_out: [0]=loss, [1]=rays_per_batch used, [2]=measured batch (uncompacted), [3]=measured compacted.
void orc_trainer_train(orc_trainer* t, uint32_t batch_size, float* stats_out);
uint32_t orc_trainer_step(const orc_trainer* t);
/* bench.py's CPU baseline only: puts the trainer into a given regime without running the preceding steps -- sets the step
 * counter (which fixes the occupancy-refresh cadence, src/testbed.cu:2538), the ray count, and replaces the occupancy grid
 * (mean + bitfield recomputed as update_density_grid_mean_and_bitfield does, src/testbed_nerf.cu:2844-2859). */
/* per-level grid scales as the device evaluates them (see orc_grid_forward's `scales`) */
void orc_trainer_set_level_scales(orc_trainer* t, const float* scales);
void orc_trainer_set_state(orc_trainer* t, uint32_t training_step, uint32_t rays_per_batch, const float* density_grid);

// ---- K17 classic render (src/testbed_nerf.cu:612-989,:1748-1978,:2047-2267): one pixel at a time ----
// camera12: 3x4 column-major camera matrix. out_rgba: [h][w][4] float (linear, premultiplied, before tonemap).
// Returns the number of network-evaluated samples.
uint64_t orc_render(const orc_model* m, const orc_half* params, const uint8_t* bitfield, const float* aabb6, uint32_t max_cascade,
                    const float* camera12, int w, int h, float fx, float fy, float cx, float cy, float cone_angle_constant,
                    float min_transmittance, int rgb_activation, int density_activation, float* out_rgba, float* out_depth);

// ---- classic render (Testbed::render_nerf, ERenderMode::Shade): see ngp_oracle.cpp for the reference sections ----
typedef struct {
	int32_t width, height;
	float fx, fy;               // focal length in pixels (Testbed::calc_focal_length, src/testbed.cu:2589)
	float screen_center[2];     // render_screen_center(), (0.5, 0.5) by default
	float camera[12];           // 3x4 camera-to-world, column-major, ngp convention
	int32_t spp, snap_to_pixel_centers;
	float aabb[6], render_aabb[6];
	float cone_angle_constant, min_transmittance, near_distance;
	int32_t rgb_activation, density_activation, train_in_linear_colors;
	int32_t color_space;        // m_color_space (0 linear, 1 sRGB): the space the accumulation buffer averages in
	int32_t output_srgb;        // !linear argument of Testbed::render
	float exposure, background_color[4];
	int32_t tonemap_curve;      // ETonemapCurve
} orc_render_config;
// out_rgba: float [height][width][4]. n_samples_out (nullable): network-evaluated samples.
void orc_render_nerf(const orc_model* m, const orc_half* params, const uint8_t* bitfield, const orc_render_config* c, float* out_rgba, uint64_t* n_samples_out);

int orc_set_num_threads(int n); // n <= 0: query only; returns omp_get_max_threads()

// ---- Blender multi-NeRF render (NerfRenderer::render, src/nerf_renderer.cu:565-791): see ngp_oracle.cpp ----
typedef struct {               // Mask3D (nerf/mask_3D.cuh:128-257)
	int32_t shape, mode;       // EMaskShape {Box, Cylinder, Sphere, All}, EMaskMode {Add, Subtract}
	float transform[16];       // 4x4 column-major: shape -> NeRF-local (instance masks) or shape -> world (request masks)
	float config[6];
	float feather, opacity;
} orc_mask;
typedef struct {
	const orc_model* model; const orc_half* params; const uint8_t* bitfield; // the NeRF's snapshot (inference parameters, occupancy bits)
	float train_aabb[6]; uint32_t aabb_scale;
	float render_aabb[6];      // NerfDescriptor::aabb, in the NeRF's local frame
	float transform[16];       // NerfDescriptor::transform, 4x4 column-major (local -> world)
	float opacity;
	int32_t rgb_activation, density_activation;
	float min_transmittance;
	uint32_t n_masks; const orc_mask* masks; // NerfDescriptor::modifiers.masks
} orc_nerf_instance;
typedef struct {
	int32_t width, height, mip, flip_y; // RenderOutputProperties: resolution, DownsampleInfo::MakeFromMip(resolution, mip), flip_y
	float camera[12];                   // RenderCameraProperties::transform, 3x4 column-major
	float focal_length, near_distance;  // pixels (same for x and y, :52)
	int32_t color_space; float exposure, background_color[4];
	int32_t camera_model;               // ECameraModel {Perspective, QuadrilateralHexahedron, SphericalQuadrilateral} (camera_models.cuh:27-31)
	float aperture_size, focus_z;
	float spherical_quadrilateral[3];   // width, height, curvature
	float quadrilateral_hexahedron[24]; // front tl, tr, bl, br; back tl, tr, bl, br
	int32_t tonemap_curve;              // ETonemapCurve {Identity, ACES, Hable, Reinhard}
	uint32_t n_masks; const orc_mask* masks; // RenderRequest::modifiers.masks (world space)
} orc_blender_request;
void orc_blender_render(const orc_blender_request* rq, uint32_t n_nerfs, const orc_nerf_instance* nerfs, float* out_rgba, uint64_t* n_samples_out);

// ---- neural-image / SDF model family: N-dimensional hash grid (N in {2,3}) + FullyFusedMLP 32 -> 64 x n_hidden -> 16 (see ngp_oracle.cpp) ----
uint32_t orc_grid_offsets_nd(uint32_t n_dims, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, uint32_t* offsets);
void orc_grid_forward_nd(uint32_t n_dims, uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale,
                         const float* scales, const orc_half* grid, const float* positions, uint32_t pos_stride, orc_half* encoded);
void orc_mlp_forward_backward(uint32_t n_hidden, uint32_t n, const orc_half* weights, const orc_half* input, orc_half* out,
                              const orc_half* dL_dout, orc_half* dL_dinput, float* grad);
void orc_loss(int kind, uint32_t n, uint32_t dims, float loss_scale, const orc_half* predictions, const float* targets, float* values, orc_half* gradients);

// ---- input gradients for camera-extrinsics optimisation (K13/K14); pinned on tests/golden/ref_camera.npz (see ngp_oracle.cpp) ----
void orc_grid_input_gradient(uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                             const orc_half* grid, const float* positions, uint32_t pos_stride, const orc_half* dL_dy, float* dL_dx);
void orc_sh4_input_gradient(uint32_t n, const float* dirs, uint32_t stride, const orc_half* dL_dy, float* dL_dx);
void orc_nerf_input_gradient(const orc_model* m, const orc_half* params, uint32_t n, const float* coords, const orc_half* dL_dout, float* dL_dcoords);
void orc_compute_cam_gradient(uint32_t n_kept, uint32_t n_rays_total, uint32_t n_images, const float* aabb6, const uint32_t* ray_indices,
                              const float* rays_unnormalized, const uint32_t* numsteps, const float* coords, const float* coords_gradient,
                              float* cam_pos_gradient, float* cam_rot_gradient);

#ifdef __cplusplus
}
#endif
