// K6 + K7: volumetric compositing, loss, dL/d(network output), compaction and roll-over padding.
// Replaces compute_loss_kernel_train_nerf (reference: src/testbed_nerf.cu:1280-1597) and
// fill_rollover / fill_rollover_and_rescale (tcnn common_device.h:517-537).
//
// Same three-stage shape as K1 so that compaction is deterministic: (A) per-ray forward compositing
// with early termination, (B) exclusive scan of the per-ray compacted step counts in ray-slot order
// (one valid serialisation of the atomicAdd at :1434), (C) per-ray gradient pass that writes the
// compacted coordinates and dL/dout, (D) roll-over padding to the fixed batch size.
// Not built (out of scope for the lego/fox configs): envmap, exposure gradients,
// depth supervision, error-map accumulation, max_level_rand_training.
#include "nerf_device.cuh"

namespace ngpb {

struct LossParams {
	uint32_t n_rays, batch, n_images;
	Aabb aabb;
	Pcg32 rng;
	ngpb_loss_config cfg;
};

struct LossAndGradient { float loss[3], gradient[3]; };

// src/testbed_nerf.cu:121-189,:1263-1278
__device__ inline LossAndGradient loss_and_gradient(const float* target, const float* prediction, int loss_type) {
	LossAndGradient r;
	#pragma unroll
	for (int c = 0; c < 3; ++c) {
		const float difference = prediction[c] - target[c];
		switch (loss_type) {
			case NGPB_LOSS_RELATIVE_L2: {
				const float factor = 1.0f / (prediction[c] * prediction[c] + 1e-2f);
				r.loss[c] = difference * difference * factor; r.gradient[c] = 2.0f * difference * factor; break;
			}
			case NGPB_LOSS_L1: r.loss[c] = fabsf(difference); r.gradient[c] = copysignf(1.0f, difference); break;
			case NGPB_LOSS_MAPE: {
				const float factor = 1.0f / (fabsf(prediction[c]) + 1e-2f);
				r.loss[c] = fabsf(difference) * factor; r.gradient[c] = copysignf(factor, difference); break;
			}
			case NGPB_LOSS_SMAPE: {
				const float factor = 1.0f / (0.5f * (fabsf(prediction[c]) + fabsf(target[c])) + 1e-2f);
				r.loss[c] = fabsf(difference) * factor; r.gradient[c] = copysignf(factor, difference); break;
			}
			case NGPB_LOSS_HUBER: { // huber_loss(alpha = 0.1) / 5 (:1274)
				const float alpha = 0.1f;
				const float abs_diff = fabsf(difference);
				const float square = 0.5f / alpha * difference * difference;
				const float l = abs_diff > alpha ? (abs_diff - 0.5f * alpha) : square;
				const float g = abs_diff > alpha ? (difference > 0 ? 1.0f : -1.0f) : (difference / alpha);
				r.loss[c] = l / 5.0f; r.gradient[c] = g / 5.0f; break;
			}
			case NGPB_LOSS_LOGL1: {
				const float divisor = fabsf(difference) + 1.0f;
				r.loss[c] = logf(divisor); r.gradient[c] = copysignf(1.0f / divisor, difference); break;
			}
			default: r.loss[c] = difference * difference; r.gradient[c] = 2.0f * difference; break;
		}
	}
	return r;
}

struct __align__(16) RayState { float rgb_ray[3]; float depth_ray; float rgbtarget[3]; uint32_t compacted; };

__device__ __forceinline__ void load_rgbsigma(const __half* p, float o[4]) {
	const uint2 raw = *reinterpret_cast<const uint2*>(p);
	const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
	o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

// (A) forward compositing, :1341-1428
__global__ void __launch_bounds__(128) loss_composite_kernel(
	const LossParams P, const ngpb_image* __restrict__ images, const uint32_t* __restrict__ counters_in,
	const __half* __restrict__ rgbsigma, const uint32_t* __restrict__ ray_indices, const float* __restrict__ rays, const uint32_t* __restrict__ numsteps_in,
	const float* __restrict__ coords_in, RayState* __restrict__ state, uint32_t* __restrict__ compacted_counts)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_rays) return;
	if (i >= counters_in[1]) { compacted_counts[i] = 0; return; }
	const uint32_t numsteps = numsteps_in[i * 2 + 0], base = numsteps_in[i * 2 + 1];
	const float* cin = coords_in + (size_t)base * COORD_FLOATS;
	const __half* no = rgbsigma + (size_t)base * 4;
	const float EPSILON = 1e-4f;
	float T = 1.f;
	float rgb_ray[3] = {0.f, 0.f, 0.f};
	float depth_ray = 0.f;
	uint32_t cn = 0;
	const float ro[3] = {rays[(size_t)i * 6 + 0], rays[(size_t)i * 6 + 1], rays[(size_t)i * 6 + 2]};
	for (; cn < numsteps; ++cn) {
		if (T < EPSILON) break;
		float o[4];
		load_rgbsigma(no, o);
		const float rgb[3] = {network_to_rgb(o[0], P.cfg.rgb_activation), network_to_rgb(o[1], P.cfg.rgb_activation), network_to_rgb(o[2], P.cfg.rgb_activation)};
		const V3 pos = unwarp_position(cin, P.aabb);
		const float dt = unwarp_dt(cin[3]);
		const float dx = pos.x - ro[0], dy = pos.y - ro[1], dz = pos.z - ro[2];
		const float cur_depth = sqrtf(sum3(dx * dx, dy * dy, dz * dz));
		const float density = network_to_density(o[3], P.cfg.density_activation);
		const float alpha = 1.f - __expf(-density * dt);
		const float weight = alpha * T;
		#pragma unroll
		for (int c = 0; c < 3; ++c) rgb_ray[c] += weight * rgb[c];
		depth_ray += weight * cur_depth;
		T *= (1.f - alpha);
		no += 4; cin += COORD_FLOATS;
	}
	// same RNG stream as K1 to recover the pixel, then the random background (:1378-1392)
	const uint32_t ray_idx = ray_indices[i];
	Pcg32 rng = P.rng;
	rng.advance((int64_t)ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY);
	const uint32_t img = image_idx(ray_idx, P.n_rays, P.n_images);
	const ngpb_image& im = images[img];
	float x, y;
	random_image_pos_training(rng, im.w, im.h, P.cfg.snap_to_pixel_centers != 0, &x, &y);
	float bg[3] = {P.cfg.background_color[0], P.cfg.background_color[1], P.cfg.background_color[2]};
	if (P.cfg.random_bg_color) { bg[0] = rng.next_float(); bg[1] = rng.next_float(); bg[2] = rng.next_float(); }
	#pragma unroll
	for (int c = 0; c < 3; ++c) bg[c] = srgb_to_linear(bg[c]);
	float texsamp[4];
	read_rgba(x, y, im, texsamp);
	float rgbtarget[3];
	if (P.cfg.linear_colors || P.cfg.color_space == NGPB_COLOR_LINEAR) {
		#pragma unroll
		for (int c = 0; c < 3; ++c) rgbtarget[c] = 1.0f * texsamp[c] + (1.0f - texsamp[3]) * bg[c];
		if (!P.cfg.linear_colors) {
			#pragma unroll
			for (int c = 0; c < 3; ++c) { rgbtarget[c] = linear_to_srgb(rgbtarget[c]); bg[c] = linear_to_srgb(bg[c]); }
		}
	} else {
		#pragma unroll
		for (int c = 0; c < 3; ++c) bg[c] = linear_to_srgb(bg[c]);
		if (texsamp[3] > 0) {
			#pragma unroll
			for (int c = 0; c < 3; ++c) rgbtarget[c] = linear_to_srgb(1.0f * texsamp[c] / texsamp[3]) * texsamp[3] + (1.0f - texsamp[3]) * bg[c];
		} else {
			#pragma unroll
			for (int c = 0; c < 3; ++c) rgbtarget[c] = bg[c];
		}
	}
	if (cn == numsteps) {
		#pragma unroll
		for (int c = 0; c < 3; ++c) rgb_ray[c] += T * bg[c];
	}
	RayState s;
	#pragma unroll
	for (int c = 0; c < 3; ++c) { s.rgb_ray[c] = rgb_ray[c]; s.rgbtarget[c] = rgbtarget[c]; }
	s.depth_ray = depth_ray;
	s.compacted = cn;
	state[i] = s;
	compacted_counts[i] = cn;
}

// (B) one block: exclusive scan of compacted counts + clipping to the batch (:1434-1437)
__global__ void __launch_bounds__(1024) loss_scan_kernel(const uint32_t n_rays, const uint32_t batch, uint32_t* __restrict__ compacted_counts,
                                                         uint32_t* __restrict__ compacted_bases, uint32_t* __restrict__ counters_out)
{
	__shared__ uint32_t smem[33];
	const uint32_t per_thread = (n_rays + 1023) / 1024;
	const uint32_t begin = min(threadIdx.x * per_thread, n_rays), end = min(begin + per_thread, n_rays);
	uint32_t sum = 0;
	for (uint32_t i = begin; i < end; ++i) sum += compacted_counts[i];
	uint32_t total;
	uint32_t base = block_exclusive_scan_1024(sum, smem, &total);
	for (uint32_t i = begin; i < end; ++i) {
		const uint32_t c = compacted_counts[i];
		compacted_bases[i] = base;
		compacted_counts[i] = min(batch - min(batch, base), c);
		base += c;
	}
	if (threadIdx.x == 0) counters_out[0] = total;
}

// (C) gradient pass, :1436-1556
__global__ void __launch_bounds__(128) loss_gradient_kernel(
	const LossParams P, const uint32_t* __restrict__ counters_in, const float* __restrict__ mean_density_ptr,
	const __half* __restrict__ rgbsigma, const float* __restrict__ rays, uint32_t* __restrict__ numsteps_io, const float* __restrict__ coords_in,
	const RayState* __restrict__ state, const uint32_t* __restrict__ compacted_counts, const uint32_t* __restrict__ compacted_bases,
	float* __restrict__ coords_out, __half* __restrict__ dloss_dout, float* __restrict__ loss_output)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_rays) return;
	if (i >= counters_in[1]) { if (loss_output) loss_output[i] = 0.f; return; }
	const uint32_t base = numsteps_io[i * 2 + 1];
	const uint32_t cn = compacted_counts[i], compacted_base = compacted_bases[i];
	numsteps_io[i * 2 + 0] = cn;
	numsteps_io[i * 2 + 1] = compacted_base;
	if (cn == 0) { if (loss_output) loss_output[i] = 0.f; return; }
	const RayState s = state[i];
	const float* cin = coords_in + (size_t)base * COORD_FLOATS;
	const __half* no = rgbsigma + (size_t)base * 4;
	float* cout = coords_out + (size_t)compacted_base * COORD_FLOATS;
	__half* dout = dloss_dout + (size_t)compacted_base * 4;
	const float ro[3] = {rays[(size_t)i * 6 + 0], rays[(size_t)i * 6 + 1], rays[(size_t)i * 6 + 2]};

	const LossAndGradient lg = loss_and_gradient(s.rgbtarget, s.rgb_ray, P.cfg.loss_type);
	const float mean_loss = sum3(lg.loss[0], lg.loss[1], lg.loss[2]) / 3.0f;
	if (loss_output) loss_output[i] = mean_loss / (float)P.n_rays;

	const float loss_scale = P.cfg.loss_scale / P.n_rays;
	const float output_l2_reg = P.cfg.rgb_activation == NGPB_ACT_EXPONENTIAL ? 1e-4f : 0.0f;
	const float output_l1_reg_density = *mean_density_ptr < NERF_MIN_OPTICAL_THICKNESS ? 1e-4f : 0.0f;

	float rgb_ray2[3] = {0.f, 0.f, 0.f};
	float depth_ray2 = 0.f;
	float T = 1.f;
	for (uint32_t j = 0; j < cn; ++j) {
		const float* ci = cin + (size_t)j * COORD_FLOATS;
		float cv[COORD_FLOATS];
		#pragma unroll
		for (int k = 0; k < (int)COORD_FLOATS; ++k) { cv[k] = ci[k]; cout[(size_t)j * COORD_FLOATS + k] = cv[k]; }
		const V3 pos = unwarp_position(cv, P.aabb);
		const float dx = pos.x - ro[0], dy = pos.y - ro[1], dz = pos.z - ro[2];
		const float depth = sqrtf(sum3(dx * dx, dy * dy, dz * dz));
		const float dt = unwarp_dt(cv[3]);
		float o[4];
		load_rgbsigma(no, o);
		const float rgb[3] = {network_to_rgb(o[0], P.cfg.rgb_activation), network_to_rgb(o[1], P.cfg.rgb_activation), network_to_rgb(o[2], P.cfg.rgb_activation)};
		const float density = network_to_density(o[3], P.cfg.density_activation);
		const float alpha = 1.f - __expf(-density * dt);
		const float weight = alpha * T;
		#pragma unroll
		for (int c = 0; c < 3; ++c) rgb_ray2[c] += weight * rgb[c];
		depth_ray2 += weight * depth;
		T *= (1.f - alpha);

		float suffix[3], tv[3], g[4];
		#pragma unroll
		for (int c = 0; c < 3; ++c) {
			suffix[c] = s.rgb_ray[c] - rgb_ray2[c];
			const float dloss_by_drgb = weight * lg.gradient[c];
			g[c] = loss_scale * (dloss_by_drgb * network_to_rgb_derivative(o[c], P.cfg.rgb_activation) + fmaxf(0.0f, output_l2_reg * o[c]));
			tv[c] = T * rgb[c] - suffix[c];
		}
		const float density_derivative = network_to_density_derivative(o[3], P.cfg.density_activation);
		const float depth_suffix = s.depth_ray - depth_ray2;
		const float depth_supervision = 0.0f * (T * depth - depth_suffix); // depth supervision off (:1450-1452)
		const float dloss_by_dmlp = density_derivative * (dt * (dot3(lg.gradient, tv) + depth_supervision));
		g[3] = loss_scale * dloss_by_dmlp +
			(o[3] < 0.0f ? -output_l1_reg_density : 0.0f) +
			(o[3] > -10.0f && depth < P.cfg.near_distance ? 1e-4f : 0.0f);
		const __half2 h01 = __halves2half2(__float2half_rn(g[0]), __float2half_rn(g[1]));
		const __half2 h23 = __halves2half2(__float2half_rn(g[2]), __float2half_rn(g[3]));
		uint2 packed;
		packed.x = *reinterpret_cast<const uint32_t*>(&h01);
		packed.y = *reinterpret_cast<const uint32_t*>(&h23);
		*reinterpret_cast<uint2*>(dout + (size_t)j * 4) = packed;
		no += 4;
	}
}

// (D) roll-over padding of the compacted batch (tcnn common_device.h:517-537): element e >= n_valid copies
// element e % n_valid; gradients of the padded copies are rescaled by n_valid / batch.
__global__ void __launch_bounds__(256) rollover_kernel(const uint32_t batch, const uint32_t* __restrict__ counters_out, float* __restrict__ coords, __half* __restrict__ dloss_dout)
{
	const uint32_t n_valid = min(counters_out[0], batch);
	if (n_valid == 0 || n_valid >= batch) return;
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= batch || e < n_valid) return;
	// the reference wraps the flat index: i % (n_valid * stride); with e = i / stride this is element e % n_valid, same member
	const uint32_t src = e % n_valid;
	#pragma unroll
	for (int k = 0; k < (int)COORD_FLOATS; ++k) coords[(size_t)e * COORD_FLOATS + k] = coords[(size_t)src * COORD_FLOATS + k];
	#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float v = __half2float(dloss_dout[(size_t)src * 4 + k]);
		dloss_dout[(size_t)e * 4 + k] = __float2half_rn(v * n_valid / batch);
	}
}

Aabb make_aabb(const float* a);

} // namespace ngpb

using namespace ngpb;

extern "C" int ngpb_compute_loss(void* stream_, uint32_t n_rays, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch) {
	try {
		if (!aabb6 || !cfg || !images_dev || !counters_in || !rgbsigma || !ray_indices || !rays || !numsteps || !coords_in || !mean_density_dev ||
		    !coords_out || !dloss_dout || !counters_out || !scratch || n_images == 0 || batch == 0) {
			set_last_error("ngpb_compute_loss: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		if (n_rays == 0) { NGPB_CUDA_CHECK(cudaMemsetAsync(counters_out, 0, sizeof(uint32_t), stream)); return 0; }
		LossParams P;
		P.n_rays = n_rays; P.batch = batch; P.n_images = n_images;
		P.aabb = make_aabb(aabb6);
		P.rng.state = rng_.state; P.rng.inc = rng_.inc;
		P.cfg = *cfg;
		// scratch layout: RayState[n_rays] (32 B each) | uint32 counts[n_rays] | uint32 bases[n_rays]
		RayState* state = reinterpret_cast<RayState*>(scratch);
		uint32_t* counts = reinterpret_cast<uint32_t*>(state + n_rays);
		uint32_t* bases = counts + n_rays;
		const uint32_t blocks = div_round_up(n_rays, 128);
		loss_composite_kernel<<<blocks, 128, 0, stream>>>(P, images_dev, counters_in, (const __half*)rgbsigma, ray_indices, rays, numsteps, coords_in, state, counts);
		NGPB_LAUNCH_CHECK();
		loss_scan_kernel<<<1, 1024, 0, stream>>>(n_rays, batch, counts, bases, counters_out);
		NGPB_LAUNCH_CHECK();
		loss_gradient_kernel<<<blocks, 128, 0, stream>>>(P, counters_in, mean_density_dev, (const __half*)rgbsigma, rays, numsteps, coords_in, state, counts, bases,
			coords_out, (__half*)dloss_dout, loss_per_ray);
		NGPB_LAUNCH_CHECK();
		rollover_kernel<<<div_round_up(batch, 256), 256, 0, stream>>>(batch, counters_out, coords_out, (__half*)dloss_dout);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
