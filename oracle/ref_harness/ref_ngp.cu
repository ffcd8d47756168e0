// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" launcher around the reference's own NeRF training kernels, obtained by including
// src/testbed_nerf.cu from where it lies under /root/reference (nothing is copied). Only the
// __global__ kernels are called; Testbed member functions defined in other translation units stay
// unresolved (the library is linked with --unresolved-symbols=ignore-all and never calls them).
//
// Kernels launched (reference file:line, all in src/testbed_nerf.cu):
//   generate_training_samples_nerf :1085      compute_loss_kernel_train_nerf :1280
//   mark_untrained_density_grid :369          generate_grid_samples_nerf_nonuniform :465
//   splat_grid_samples_nerf_max_nearest_neighbor :496   ema_grid_samples_nerf :532
//   grid_to_bitfield :563                     bitfield_max_pool :589
#include <testbed_nerf.cu>

using namespace ngp;
using namespace tcnn;
using namespace Eigen;

namespace {

struct DeviceDataset {
	GPUMemory<TrainingImageMetadata> metadata;
	GPUMemory<TrainingXForm> xforms;
};

// Builds the per-image metadata array the reference kernels read (nerf_loader.h:30-46).
// pixels: one device buffer holding n_images RGBA8 images of w*h pixels back to back.
// lens applied to every image of the next make_dataset call (set by ref_set_lens; mode 0 = perspective)
static Lens g_lens{};

DeviceDataset make_dataset(uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_3x4_colmajor_host) {
	std::vector<TrainingImageMetadata> md(n_images);
	std::vector<TrainingXForm> xf(n_images);
	for (uint32_t i = 0; i < n_images; ++i) {
		md[i].pixels = pixels + (size_t)i * w * h * 4;
		md[i].image_data_type = EImageDataType::Byte;
		md[i].resolution = {w, h};
		md[i].focal_length = {fx, fy};
		md[i].principal_point = {cx, cy};
		md[i].lens = g_lens;
		Matrix<float, 3, 4> m;
		for (int k = 0; k < 12; ++k) m.data()[k] = xforms_3x4_colmajor_host[i * 12 + k];
		xf[i].start = m;
		xf[i].end = m;
	}
	DeviceDataset d;
	d.metadata.resize_and_copy_from_host(md);
	d.xforms.resize_and_copy_from_host(xf);
	return d;
}

BoundingBox make_aabb(const float* aabb6) {
	return BoundingBox{Vector3f{aabb6[0], aabb6[1], aabb6[2]}, Vector3f{aabb6[3], aabb6[4], aabb6[5]}};
}

}

extern "C" {

// ELensMode + the 7 lens parameters used by the datasets the following calls build (generate_training_samples, mark_untrained_density_grid)
void ref_set_lens(int mode, const float* params7) { g_lens.mode = (ELensMode)mode; for (int k = 0; k < 7; ++k) g_lens.params[k] = params7 ? params7[k] : 0.f; }

// iters > 0: the launch (with its two counter memsets, as in train_nerf_step) is repeated `iters` times after one warm-up and *ms receives the mean
// milliseconds per repetition (CUDA events on the NULL stream); the outputs are those of the last repetition.
int ref_generate_training_samples_timed(
	uint32_t n_rays, const float* aabb6, uint32_t max_samples, uint32_t n_rays_total,
	uint64_t rng_state, uint64_t rng_inc,
	uint32_t* ray_counter, uint32_t* numsteps_counter, uint32_t* ray_indices, float* rays /*6 floats*/, uint32_t* numsteps, float* coords /*7 floats*/,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant, int iters, float* ms
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int it = 0; it <= iters; ++it) {
		if (it == 1) cudaEventRecord(e0, nullptr);
		cudaMemsetAsync(ray_counter, 0, 4, nullptr);
		cudaMemsetAsync(numsteps_counter, 0, 4, nullptr);
		linear_kernel(generate_training_samples_nerf, 0, 0,
			n_rays, make_aabb(aabb6), max_samples, n_rays_total, rng,
			ray_counter, numsteps_counter, ray_indices, (Ray*)rays, numsteps,
			PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0),
			n_images, d.metadata.data(), d.xforms.data(), bitfield,
			false, (float*)nullptr, (bool)snap_to_pixel_centers, false, cone_angle_constant,
			(const float*)nullptr, Vector2i{0, 0}, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i{0, 0},
			(const float*)nullptr, 0u);
	}
	cudaEventRecord(e1, nullptr);
	cudaEventSynchronize(e1);
	if (iters > 0 && ms) { cudaEventElapsedTime(ms, e0, e1); *ms /= (float)iters; }
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return (int)cudaDeviceSynchronize();
}

int ref_generate_training_samples(
	uint32_t n_rays, const float* aabb6, uint32_t max_samples, uint32_t n_rays_total,
	uint64_t rng_state, uint64_t rng_inc,
	uint32_t* ray_counter, uint32_t* numsteps_counter, uint32_t* ray_indices, float* rays /*6 floats*/, uint32_t* numsteps, float* coords /*7 floats*/,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	cudaMemset(ray_counter, 0, 4);
	cudaMemset(numsteps_counter, 0, 4);
	linear_kernel(generate_training_samples_nerf, 0, 0,
		n_rays, make_aabb(aabb6), max_samples, n_rays_total, rng,
		ray_counter, numsteps_counter, ray_indices, (Ray*)rays, numsteps,
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0),
		n_images, d.metadata.data(), d.xforms.data(), bitfield,
		false, (float*)nullptr, (bool)snap_to_pixel_centers, false, cone_angle_constant,
		(const float*)nullptr, Vector2i{0, 0}, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i{0, 0},
		(const float*)nullptr, 0u);
	return (int)cudaDeviceSynchronize();
}

// loss_type / activations are the reference's enum integers (common.h:103-119).
int ref_compute_loss(
	uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, uint64_t rng_state, uint64_t rng_inc,
	uint32_t max_samples_compacted, const uint32_t* rays_counter, float loss_scale, int padded_output_width,
	const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const void* network_output_half, uint32_t* numsteps_counter_compacted, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps,
	const float* coords_in, float* coords_out, void* dloss_doutput_half, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, const float* mean_density_ptr, float near_distance
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	GPUMemory<Array3f> exposure(n_images);
	exposure.memset(0);
	cudaMemset(numsteps_counter_compacted, 0, 4);
	linear_kernel(compute_loss_kernel_train_nerf, 0, 0,
		n_rays, make_aabb(aabb6), n_rays_total, rng, max_samples_compacted, rays_counter, loss_scale, padded_output_width,
		(const float*)nullptr, (float*)nullptr, Vector2i{0, 0}, ELossType::L2,
		Array3f{background_color3[0], background_color3[1], background_color3[2]}, (EColorSpace)color_space, (bool)random_bg, (bool)linear_colors,
		n_images, d.metadata.data(), (const network_precision_t*)network_output_half, numsteps_counter_compacted,
		ray_indices, (const Ray*)rays, numsteps,
		PitchedPtr<const NerfCoordinate>((NerfCoordinate*)coords_in, 1, 0, 0),
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_out, 1, 0, 0),
		(network_precision_t*)dloss_doutput_half, (ELossType)loss_type, ELossType::L1, loss_output,
		false, (float*)nullptr, (ENerfActivation)rgb_activation, (ENerfActivation)density_activation, (bool)snap_to_pixel_centers,
		(float*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i{0, 0}, Vector2i{0, 0},
		(const float*)nullptr, Vector2i{0, 0}, (float*)nullptr, (float*)nullptr, mean_density_ptr,
		(const Array3f*)exposure.data(), (Array3f*)nullptr, 0.0f, near_distance);
	return (int)cudaDeviceSynchronize();
}

// The same kernel with per-image exposures (host float[n_images*3]) and the accumulated exposure gradient (host float[n_images*3], out):
// src/testbed_nerf.cu:1403,:1558-1571.
int ref_compute_loss_exposure(
	uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, uint64_t rng_state, uint64_t rng_inc,
	uint32_t max_samples_compacted, const uint32_t* rays_counter, float loss_scale, int padded_output_width,
	const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const void* network_output_half, uint32_t* numsteps_counter_compacted, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps,
	const float* coords_in, float* coords_out, void* dloss_doutput_half, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, const float* mean_density_ptr, float near_distance,
	const float* exposure_host, float* exposure_gradient_host
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	GPUMemory<Array3f> exposure(n_images), exposure_gradient(n_images);
	cudaMemcpy(exposure.data(), exposure_host, (size_t)n_images * 12, cudaMemcpyHostToDevice);
	exposure_gradient.memset(0);
	cudaMemset(numsteps_counter_compacted, 0, 4);
	linear_kernel(compute_loss_kernel_train_nerf, 0, 0,
		n_rays, make_aabb(aabb6), n_rays_total, rng, max_samples_compacted, rays_counter, loss_scale, padded_output_width,
		(const float*)nullptr, (float*)nullptr, Vector2i{0, 0}, ELossType::L2,
		Array3f{background_color3[0], background_color3[1], background_color3[2]}, (EColorSpace)color_space, (bool)random_bg, (bool)linear_colors,
		n_images, d.metadata.data(), (const network_precision_t*)network_output_half, numsteps_counter_compacted,
		ray_indices, (const Ray*)rays, numsteps,
		PitchedPtr<const NerfCoordinate>((NerfCoordinate*)coords_in, 1, 0, 0),
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_out, 1, 0, 0),
		(network_precision_t*)dloss_doutput_half, (ELossType)loss_type, ELossType::L1, loss_output,
		false, (float*)nullptr, (ENerfActivation)rgb_activation, (ENerfActivation)density_activation, (bool)snap_to_pixel_centers,
		(float*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i{0, 0}, Vector2i{0, 0},
		(const float*)nullptr, Vector2i{0, 0}, (float*)nullptr, (float*)nullptr, mean_density_ptr,
		(const Array3f*)exposure.data(), exposure_gradient.data(), 0.0f, near_distance);
	int st = (int)cudaDeviceSynchronize();
	cudaMemcpy(exposure_gradient_host, exposure_gradient.data(), (size_t)n_images * 12, cudaMemcpyDeviceToHost);
	return st;
}

// ---- K19: error-map importance sampling (src/testbed_nerf.cu:991-1083, :1465-1491, :1984-2039) ----
// K1 with the error-map CDFs (device pointers; cdf_img may be null): pixels and images are drawn proportionally to the accumulated error.
int ref_generate_training_samples_cdf(
	uint32_t n_rays, const float* aabb6, uint32_t max_samples, uint32_t n_rays_total,
	uint64_t rng_state, uint64_t rng_inc,
	uint32_t* ray_counter, uint32_t* numsteps_counter, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant,
	const float* cdf_x_cond_y, const float* cdf_y, const float* cdf_img, int cdf_res_x, int cdf_res_y
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	cudaMemset(ray_counter, 0, 4);
	cudaMemset(numsteps_counter, 0, 4);
	linear_kernel(generate_training_samples_nerf, 0, 0,
		n_rays, make_aabb(aabb6), max_samples, n_rays_total, rng,
		ray_counter, numsteps_counter, ray_indices, (Ray*)rays, numsteps,
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0),
		n_images, d.metadata.data(), d.xforms.data(), bitfield,
		false, (float*)nullptr, (bool)snap_to_pixel_centers, false, cone_angle_constant,
		(const float*)nullptr, Vector2i{0, 0}, cdf_x_cond_y, cdf_y, cdf_img, Vector2i{cdf_res_x, cdf_res_y},
		(const float*)nullptr, 0u);
	return (int)cudaDeviceSynchronize();
}

// K6 with the CDFs (the loss is divided by the sampling pdf) and the error-map accumulation (error_map: device float[n_images * res_y * res_x], added to).
int ref_compute_loss_error_map(
	uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, uint64_t rng_state, uint64_t rng_inc,
	uint32_t max_samples_compacted, const uint32_t* rays_counter, float loss_scale, int padded_output_width,
	const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const void* network_output_half, uint32_t* numsteps_counter_compacted, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps,
	const float* coords_in, float* coords_out, void* dloss_doutput_half, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, const float* mean_density_ptr, float near_distance,
	float* error_map, int error_map_res_x, int error_map_res_y,
	const float* cdf_x_cond_y, const float* cdf_y, const float* cdf_img, int cdf_res_x, int cdf_res_y
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	GPUMemory<Array3f> exposure(n_images);
	exposure.memset(0);
	cudaMemset(numsteps_counter_compacted, 0, 4);
	linear_kernel(compute_loss_kernel_train_nerf, 0, 0,
		n_rays, make_aabb(aabb6), n_rays_total, rng, max_samples_compacted, rays_counter, loss_scale, padded_output_width,
		(const float*)nullptr, (float*)nullptr, Vector2i{0, 0}, ELossType::L2,
		Array3f{background_color3[0], background_color3[1], background_color3[2]}, (EColorSpace)color_space, (bool)random_bg, (bool)linear_colors,
		n_images, d.metadata.data(), (const network_precision_t*)network_output_half, numsteps_counter_compacted,
		ray_indices, (const Ray*)rays, numsteps,
		PitchedPtr<const NerfCoordinate>((NerfCoordinate*)coords_in, 1, 0, 0),
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_out, 1, 0, 0),
		(network_precision_t*)dloss_doutput_half, (ELossType)loss_type, ELossType::L1, loss_output,
		false, (float*)nullptr, (ENerfActivation)rgb_activation, (ENerfActivation)density_activation, (bool)snap_to_pixel_centers,
		error_map, cdf_x_cond_y, cdf_y, cdf_img, Vector2i{error_map_res_x, error_map_res_y}, Vector2i{cdf_res_x, cdf_res_y},
		(const float*)nullptr, Vector2i{0, 0}, (float*)nullptr, (float*)nullptr, mean_density_ptr,
		(const Array3f*)exposure.data(), (Array3f*)nullptr, 0.0f, near_distance);
	return (int)cudaDeviceSynchronize();
}

// construct_cdf_2d + construct_cdf_1d with the launch shapes of Testbed::train_nerf (:2985-2998). All pointers on the device; cdf_img receives the
// un-normalised per-image sums (the reference normalises them on the host, :3000-3015).
int ref_construct_cdfs(uint32_t n_images, uint32_t height, uint32_t width, const float* error_map, float* cdf_x_cond_y, float* cdf_y, float* cdf_img) {
	const dim3 threads = { 16, 8, 1 };
	const dim3 blocks = { div_round_up(height, threads.x), div_round_up(n_images, threads.y), 1 };
	construct_cdf_2d<<<blocks, threads, 0, nullptr>>>(n_images, height, width, error_map, cdf_x_cond_y, cdf_y);
	linear_kernel(construct_cdf_1d, 0, nullptr, n_images, height, cdf_y, cdf_img);
	return (int)cudaDeviceSynchronize();
}

// Timing variant: see ref_generate_training_samples_timed.
int ref_compute_loss_timed(
	uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, uint64_t rng_state, uint64_t rng_inc,
	uint32_t max_samples_compacted, const uint32_t* rays_counter, float loss_scale, int padded_output_width,
	const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const void* network_output_half, uint32_t* numsteps_counter_compacted, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps,
	const float* coords_in, float* coords_out, void* dloss_doutput_half, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, const float* mean_density_ptr, float near_distance, int iters, float* ms
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	GPUMemory<Array3f> exposure(n_images);
	exposure.memset(0);
	GPUMemory<uint32_t> numsteps_backup((size_t)n_rays * 2);
	cudaMemcpy(numsteps_backup.data(), numsteps, (size_t)n_rays * 8, cudaMemcpyDeviceToDevice);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	float total = 0.f;
	for (int it = 0; it <= iters; ++it) {
	cudaMemcpy(numsteps, numsteps_backup.data(), (size_t)n_rays * 8, cudaMemcpyDeviceToDevice); // the kernel rewrites numsteps in place: restore (untimed)
	cudaEventRecord(e0, nullptr);
	cudaMemsetAsync(numsteps_counter_compacted, 0, 4, nullptr);
	linear_kernel(compute_loss_kernel_train_nerf, 0, 0,
		n_rays, make_aabb(aabb6), n_rays_total, rng, max_samples_compacted, rays_counter, loss_scale, padded_output_width,
		(const float*)nullptr, (float*)nullptr, Vector2i{0, 0}, ELossType::L2,
		Array3f{background_color3[0], background_color3[1], background_color3[2]}, (EColorSpace)color_space, (bool)random_bg, (bool)linear_colors,
		n_images, d.metadata.data(), (const network_precision_t*)network_output_half, numsteps_counter_compacted,
		ray_indices, (const Ray*)rays, numsteps,
		PitchedPtr<const NerfCoordinate>((NerfCoordinate*)coords_in, 1, 0, 0),
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_out, 1, 0, 0),
		(network_precision_t*)dloss_doutput_half, (ELossType)loss_type, ELossType::L1, loss_output,
		false, (float*)nullptr, (ENerfActivation)rgb_activation, (ENerfActivation)density_activation, (bool)snap_to_pixel_centers,
		(float*)nullptr, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i{0, 0}, Vector2i{0, 0},
		(const float*)nullptr, Vector2i{0, 0}, (float*)nullptr, (float*)nullptr, mean_density_ptr,
		(const Array3f*)exposure.data(), (Array3f*)nullptr, 0.0f, near_distance);
	cudaEventRecord(e1, nullptr);
	cudaEventSynchronize(e1);
	float t = 0.f; cudaEventElapsedTime(&t, e0, e1);
	if (it > 0) total += t;
	}
	if (iters > 0 && ms) *ms = total / (float)iters;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return (int)cudaDeviceSynchronize();
}

int ref_mark_untrained_density_grid(uint32_t n_elements, float* grid, uint32_t n_images, int w, int h, float fx, float fy, const float* xforms_host, int clear_visible) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, 0.5f, 0.5f, nullptr, xforms_host);
	linear_kernel(mark_untrained_density_grid, 0, 0, n_elements, grid, n_images, d.metadata.data(), d.xforms.data(), (bool)clear_visible);
	return (int)cudaDeviceSynchronize();
}

int ref_generate_grid_samples(uint32_t n_elements, uint64_t rng_state, uint64_t rng_inc, uint32_t step, const float* aabb6, const float* grid_in, float* positions3, uint32_t* indices, uint32_t n_cascades, float thresh) {
	default_rng_t rng; rng.state = rng_state; rng.inc = rng_inc;
	linear_kernel(generate_grid_samples_nerf_nonuniform, 0, 0, n_elements, rng, step, make_aabb(aabb6), grid_in, (NerfPosition*)positions3, indices, n_cascades, thresh);
	return (int)cudaDeviceSynchronize();
}

int ref_splat_and_ema(uint32_t n_samples, const uint32_t* indices, const void* density_half, float* grid_tmp, uint32_t n_elements, float decay, float* grid) {
	cudaMemset(grid_tmp, 0, sizeof(float) * n_elements);
	linear_kernel(splat_grid_samples_nerf_max_nearest_neighbor, 0, 0, n_samples, indices, (const network_precision_t*)density_half, grid_tmp, ENerfActivation::Logistic, ENerfActivation::Exponential);
	linear_kernel(ema_grid_samples_nerf, 0, 0, n_elements, decay, 0u, grid, grid_tmp);
	return (int)cudaDeviceSynchronize();
}

// mean_density_ptr is an input here (the reference computes it with tcnn::reduce_sum, testbed_nerf.cu:2852).
int ref_bitfield(uint32_t n_cascades_used, const float* grid, uint8_t* bitfield, const float* mean_density_ptr) {
	const uint32_t n_elements = 128 * 128 * 128;
	linear_kernel(grid_to_bitfield, 0, 0, n_elements / 8 * 8, n_elements / 8 * n_cascades_used, grid, bitfield, mean_density_ptr);
	for (uint32_t level = 1; level < 8; ++level) {
		linear_kernel(bitfield_max_pool, 0, 0, n_elements / 64, bitfield + grid_mip_offset(level - 1) / 8, bitfield + grid_mip_offset(level) / 8);
	}
	return (int)cudaDeviceSynchronize();
}


// compute_cam_gradient_train_nerf (src/testbed_nerf.cu:1598-1707) with uniform pixel sampling, no distortion map, no focal-length gradient.
// cam_pos_gradient / cam_rot_gradient: device float[n_images][3], zeroed here first. All other pointers are device buffers.
int ref_compute_cam_gradient(
	uint32_t n_rays, uint32_t n_rays_total, const float* aabb6, const uint32_t* rays_counter,
	uint32_t n_images, int w, int h, float fx, float fy, float cx, float cy, const uint8_t* pixels, const float* xforms_host,
	const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords, const float* coords_gradient,
	float* cam_pos_gradient, float* cam_rot_gradient
) {
	DeviceDataset d = make_dataset(n_images, w, h, fx, fy, cx, cy, pixels, xforms_host);
	default_rng_t rng;
	cudaMemset(cam_pos_gradient, 0, sizeof(float) * 3 * n_images);
	cudaMemset(cam_rot_gradient, 0, sizeof(float) * 3 * n_images);
	linear_kernel(compute_cam_gradient_train_nerf, 0, 0,
		n_rays, n_rays_total, rng, make_aabb(aabb6), rays_counter, d.xforms.data(), false,
		(Vector3f*)cam_pos_gradient, (Vector3f*)cam_rot_gradient, n_images, d.metadata.data(),
		ray_indices, (const Ray*)rays, numsteps,
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords, 1, 0, 0),
		PitchedPtr<NerfCoordinate>((NerfCoordinate*)coords_gradient, 1, 0, 0),
		(float*)nullptr, (float*)nullptr, Vector2i{0, 0}, (Vector2f*)nullptr,
		(const float*)nullptr, (const float*)nullptr, (const float*)nullptr, Vector2i{0, 0});
	return (int)cudaDeviceSynchronize();
}

}
