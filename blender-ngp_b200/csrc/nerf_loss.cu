// K6 + K7: volumetric compositing, loss, dL/d(network output), compaction and roll-over padding.
// Replaces compute_loss_kernel_train_nerf (reference: src/testbed_nerf.cu:1280-1597) and
// fill_rollover / fill_rollover_and_rescale (tcnn common_device.h:517-537).
//
// Same three-stage shape as K1 so that compaction is deterministic: (A) forward compositing with early
// termination, one WARP per ray and one lane per sample (the reference walks each ray serially in one
// thread, so its run time is set by the longest ray), (B) exclusive scan of the compacted step counts
// in ray-slot order (one valid serialisation of the atomicAdd at :1434; block-local in (A), the
// per-block totals by one small block), (C) gradient pass, again a warp per ray, that writes the
// compacted coordinates and dL/dout, (D) roll-over padding to the fixed batch size.
// Not built (out of scope for the lego/fox configs): envmap, exposure gradients,
// depth supervision, error-map accumulation, max_level_rand_training.
#include "nerf_device.cuh"
#include <cstdio>
#include <cstdlib>
#include <utility>

namespace ngpb {

struct LossParams {
	uint32_t rows_tiled; // layout of the feature rows that are compacted along with the samples: 0 = [n][32] row-major (C ABI), 1 = UMMA tiles (hash_grid.cu)
	uint32_t n_rays, batch, n_images;
	uint32_t n_rays_global; // rays of the whole (all-shard) batch: pixel selection and loss normalisation use it (:1062-1083, :1493)
	Aabb aabb;
	Pcg32 rng;
	ngpb_loss_config cfg;
	ErrorCdf cdf; // K19: the CDFs K1 drew this batch's pixels / images from (null members: uniform)
};

// index (in 16-byte units) of chunk c of sample i's 64-byte feature row
__device__ __forceinline__ size_t feature_chunk(uint32_t tiled, size_t i, uint32_t c) {
	return tiled ? ((i >> 7) << 9) + (((i & 127u) >> 3) << 5) + (c << 3) + (i & 7u) : i * 4 + c;
}

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

// pdf = img_pdf * xy_pdf: density (relative to uniform) the ray's pixel was drawn with, xy_pdf its pixel factor; both exactly 1 without error-map sampling
struct __align__(16) RayState { float rgb_ray[3]; float pdf; float rgbtarget[3]; uint32_t compacted; float bg[3]; float xy_pdf; };

__device__ __forceinline__ void load_rgbsigma(const __half* p, float o[4]) {
	const uint2 raw = *reinterpret_cast<const uint2*>(p);
	const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
	o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

// ---- warp-cooperative ray walk --------------------------------------------------------------------
// One warp per ray, one lane per sample, 32 samples per round. The reference walks a ray's samples serially in one
// thread (:1341-1373); here the per-sample math (activations, exp) runs in parallel across the lanes and only the
// transmittance product is serial: lane l multiplies a_0 .. a_{l-1} in the reference's left-to-right order, so T
// before every sample -- and with it the early-termination index -- carries the same rounding as the serial loop.
constexpr uint32_t LOSS_RAYS_PER_BLOCK = 32; // 1024 threads
constexpr float T_EPSILON = 1e-4f;

struct SampleEval { float rgb[3]; float alpha, one_minus_alpha, dt; float o[4]; };

__device__ __forceinline__ SampleEval eval_sample(const __half* __restrict__ rgbsigma, const float* __restrict__ coords_in, size_t idx, const ngpb_loss_config& cfg, bool valid) {
	SampleEval e;
	if (valid) {
		load_rgbsigma(rgbsigma + idx * 4, e.o);
		e.dt = unwarp_dt(coords_in[idx * COORD_FLOATS + 3]);
		#pragma unroll
		for (int c = 0; c < 3; ++c) e.rgb[c] = network_to_rgb(e.o[c], cfg.rgb_activation);
		const float density = network_to_density(e.o[3], cfg.density_activation);
		e.alpha = 1.f - __expf(-density * e.dt);
		e.one_minus_alpha = 1.f - e.alpha;
	} else {
		e.o[0] = e.o[1] = e.o[2] = e.o[3] = 0.f; e.dt = 0.f;
		e.rgb[0] = e.rgb[1] = e.rgb[2] = 0.f; e.alpha = 0.f; e.one_minus_alpha = 1.f; // neutral: multiplying by 1.0f is exact
	}
	return e;
}

// A ray is walked by a GROUP of W adjacent lanes (W = 8, 16 or 32; 32 / W rays per warp): rays keep ~6 samples after compaction and ~15 before,
// so a whole warp per ray leaves most lanes idle. All shuffles run warp-wide (every lane takes part), groups only select their own lanes.
// T before this lane's sample, given T before the group's first sample; products taken in sample order.
template <uint32_t W>
__device__ __forceinline__ float sequential_prefix_product(float T_in, float one_minus_alpha, uint32_t lane) {
	const uint32_t glane = lane & (W - 1), gbase = lane & ~(W - 1);
	float T = T_in;
	#pragma unroll
	for (uint32_t k = 0; k < W - 1; ++k) {
		const float v = __shfl_sync(0xffffffffu, one_minus_alpha, gbase + k);
		if (k < glane) T *= v;
	}
	return T;
}
template <uint32_t W>
__device__ __forceinline__ float group_sum(float v) {
	#pragma unroll
	for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
template <uint32_t W>
__device__ __forceinline__ float group_inclusive_sum(float v, uint32_t lane) {
	const uint32_t glane = lane & (W - 1);
	#pragma unroll
	for (int o = 1; o < (int)W; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, v, o); if (glane >= (uint32_t)o) v += t; }
	return v;
}
template <uint32_t W>
__device__ __forceinline__ uint32_t group_ballot(bool pred, uint32_t lane) { // the group's W bits of a warp-wide ballot, group lane 0 = bit 0
	const uint32_t b = __ballot_sync(0xffffffffu, pred);
	return W == 32 ? b : (b >> (lane & ~(W - 1))) & ((1u << (W & 31)) - 1u);
}

// Block-wide exclusive scan of one value per warp (32 warps); returns the exclusive prefix for this warp and the block total.
__device__ __forceinline__ uint32_t block_scan_warps(uint32_t v_warp, uint32_t* smem /*[33]*/, uint32_t* total) {
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0) smem[warp] = v_warp;
	__syncthreads();
	if (warp == 0) {
		const uint32_t w = smem[lane];
		uint32_t wi = w;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += t; }
		smem[lane] = wi - w;
		if (lane == 31) smem[32] = wi;
	}
	__syncthreads();
	*total = smem[32];
	return smem[warp];
}

// (A0) target colour and background of every ray, one THREAD per ray (:1378-1420): the pixel is recovered from the same RNG stream as K1, the
// random background comes from the next three draws. Kept out of the warp-per-ray kernels, where all 32 lanes would repeat it.
__global__ void __launch_bounds__(128) loss_target_kernel(const LossParams P, const ngpb_image* __restrict__ images, const uint32_t* __restrict__ counters_in,
                                                          const uint32_t* __restrict__ ray_indices, RayState* __restrict__ state, const float* __restrict__ exposure)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_rays || i >= counters_in[1]) return;
	{
		// same RNG stream as K1 to recover the pixel, then the random background (:1378-1392)
		const uint32_t ray_idx = ray_indices[i];
		Pcg32 rng = P.rng;
		rng.advance((int64_t)ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float img_pdf = 1.0f, xy_pdf = 1.0f;
		const uint32_t img = image_idx(ray_idx, P.n_rays_global, P.n_images, P.cdf.img, &img_pdf);
		const ngpb_image& im = images[img];
		float x, y;
		random_image_pos_training(rng, im.w, im.h, P.cfg.snap_to_pixel_centers != 0, &x, &y, P.cdf, img, &xy_pdf);
		float bg[3] = {P.cfg.background_color[0], P.cfg.background_color[1], P.cfg.background_color[2]};
		if (P.cfg.random_bg_color) { bg[0] = rng.next_float(); bg[1] = rng.next_float(); bg[2] = rng.next_float(); }
		#pragma unroll
		for (int c = 0; c < 3; ++c) bg[c] = srgb_to_linear(bg[c]);
		float texsamp[4];
		read_rgba(x, y, im, texsamp);
		float rgbtarget[3];
		// exposure_scale = exp(ln 2 * exposure) per channel of the image (:1403); exactly 1 for a zero exposure
		float es[3] = {1.0f, 1.0f, 1.0f};
		if (exposure) {
			#pragma unroll
			for (int c = 0; c < 3; ++c) es[c] = expf(0.6931471805599453f * exposure[(size_t)img * 3 + c]);
		}
		if (P.cfg.linear_colors || P.cfg.color_space == NGPB_COLOR_LINEAR) {
			#pragma unroll
			for (int c = 0; c < 3; ++c) rgbtarget[c] = es[c] * texsamp[c] + (1.0f - texsamp[3]) * bg[c];
			if (!P.cfg.linear_colors) {
				#pragma unroll
				for (int c = 0; c < 3; ++c) { rgbtarget[c] = linear_to_srgb(rgbtarget[c]); bg[c] = linear_to_srgb(bg[c]); }
			}
		} else {
			#pragma unroll
			for (int c = 0; c < 3; ++c) bg[c] = linear_to_srgb(bg[c]);
			if (texsamp[3] > 0) {
				#pragma unroll
				for (int c = 0; c < 3; ++c) rgbtarget[c] = linear_to_srgb(es[c] * texsamp[c] / texsamp[3]) * texsamp[3] + (1.0f - texsamp[3]) * bg[c];
			} else {
				#pragma unroll
				for (int c = 0; c < 3; ++c) rgbtarget[c] = bg[c];
			}
		}
		RayState s{};
		#pragma unroll
		for (int c = 0; c < 3; ++c) { s.rgbtarget[c] = rgbtarget[c]; s.bg[c] = bg[c]; }
		s.pdf = img_pdf * xy_pdf; s.xy_pdf = xy_pdf;
		state[i] = s;
	}
}

// (A) forward compositing, :1341-1428. Writes the per-ray state, the ray's compacted step count, its exclusive prefix inside
// the block (local_bases) and the block's total (block_sums). 1024 / W rays per block.
template <uint32_t W>
__global__ void __launch_bounds__(1024) loss_composite_kernel(
	const LossParams P, const ngpb_image* __restrict__ images, const uint32_t* __restrict__ counters_in,
	const __half* __restrict__ rgbsigma, const uint32_t* __restrict__ ray_indices, const uint32_t* __restrict__ numsteps_in,
	const float* __restrict__ coords_in, RayState* __restrict__ state, uint32_t* __restrict__ compacted_counts, uint32_t* __restrict__ local_bases,
	uint32_t* __restrict__ block_sums)
{
	__shared__ uint32_t scan_smem[33];
	constexpr uint32_t RAYS_PER_WARP = 32 / W, RAYS_PER_BLOCK = 1024 / W;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, glane = lane & (W - 1);
	const uint32_t i = blockIdx.x * RAYS_PER_BLOCK + warp * RAYS_PER_WARP + lane / W;
	const bool active = i < P.n_rays && i < counters_in[1];
	uint32_t cn = 0;
	const uint32_t numsteps = active ? numsteps_in[i * 2 + 0] : 0u, base = active ? numsteps_in[i * 2 + 1] : 0u;
	float T_in = 1.f;
	float rgb_ray[3] = {0.f, 0.f, 0.f};
	bool stopped = false;
	for (uint32_t r0 = 0; __any_sync(0xffffffffu, r0 < numsteps && !stopped); r0 += W) {
		const bool walking = r0 < numsteps && !stopped; // (uniform within the group)
		const uint32_t j = r0 + glane;
		const bool valid = walking && j < numsteps;
		const SampleEval e = eval_sample(rgbsigma, coords_in, (size_t)base + j, P.cfg, valid);
		const float T = sequential_prefix_product<W>(T_in, e.one_minus_alpha, lane);
		// the serial loop tests `T < EPSILON` before it touches a sample (:1352): the walk ends at the first such sample
		const uint32_t stop_mask = group_ballot<W>(valid && T < T_EPSILON, lane);
		const uint32_t first_stop = stop_mask ? (uint32_t)__ffs(stop_mask) - 1 : 32u;
		const bool counted = valid && glane < first_stop;
		const float weight = counted ? e.alpha * T : 0.f;
		#pragma unroll
		for (int c = 0; c < 3; ++c) rgb_ray[c] += group_sum<W>(weight * e.rgb[c]);
		cn += __popc(group_ballot<W>(counted, lane));
		const float T_next = __shfl_sync(0xffffffffu, T * e.one_minus_alpha, (lane & ~(W - 1)) + W - 1);
		if (walking) { stopped = stop_mask != 0; T_in = T_next; }
	}
	if (active) {
		const RayState tgt = state[i]; // target colour and background, from loss_target_kernel
		const float bg[3] = {tgt.bg[0], tgt.bg[1], tgt.bg[2]};
		const float rgbtarget[3] = {tgt.rgbtarget[0], tgt.rgbtarget[1], tgt.rgbtarget[2]};
		if (cn == numsteps) { // the ray reached the end of its samples: composite the background behind it (:1424-1427)
			#pragma unroll
			for (int c = 0; c < 3; ++c) rgb_ray[c] += T_in * bg[c];
		}
		if (glane == 0) {
			RayState s;
			#pragma unroll
			for (int c = 0; c < 3; ++c) { s.rgb_ray[c] = rgb_ray[c]; s.rgbtarget[c] = rgbtarget[c]; }
			s.pdf = tgt.pdf; s.xy_pdf = tgt.xy_pdf;
			#pragma unroll
			for (int c = 0; c < 3; ++c) s.bg[c] = bg[c];
			s.compacted = cn;
			state[i] = s;
		}
	}
	// exclusive prefix of the rays' counts: across the groups of the warp (leaders carry the count), then across the warps of the block
	uint32_t incl = glane == 0 ? cn : 0u;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
	const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
	uint32_t total;
	const uint32_t warp_base = block_scan_warps(warp_total, scan_smem, &total);
	if (glane == 0 && i < P.n_rays) { compacted_counts[i] = cn; local_bases[i] = warp_base + incl - cn; }
	if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// (B) one block: exclusive scan of the per-block totals (one valid serialisation of the atomicAdd at :1434).
// Compaction order. Ray-slot order while the batch holds every sample. When it overflows, the rays served last are clipped (:1436-1437): in slot order
// those would always be the rays of the LAST training images (slots follow the ray index, the ray index selects the image), whereas the reference's
// atomics clip whichever rays its blocks happen to process last. So the order then starts at a ray drawn from the step's RNG (past the rays' own
// sub-streams) and wraps around: rot[0] = that ray's slot, rot[1] = its base in slot order; the gradient kernel rotates every base by it.
__global__ void __launch_bounds__(1024) loss_scan_kernel(const LossParams P, const uint32_t* __restrict__ counters_in, const uint32_t n_blocks, uint32_t* __restrict__ block_sums,
                                                         const uint32_t* __restrict__ local_bases, const uint32_t RB_C, uint32_t* __restrict__ counters_out, uint32_t* __restrict__ rot)
{
	__shared__ uint32_t smem[33];
	uint32_t carry = 0;
	for (uint32_t t0 = 0; t0 < n_blocks; t0 += 1024) {
		const uint32_t b = t0 + threadIdx.x;
		const uint32_t v = b < n_blocks ? block_sums[b] : 0u;
		uint32_t total;
		const uint32_t excl = block_exclusive_scan_1024(v, smem, &total);
		if (b < n_blocks) block_sums[b] = carry + excl;
		carry += total;
	}
	__syncthreads(); // (thread 0 reads prefixes written by other threads)
	if (threadIdx.x == 0) {
		counters_out[0] = carry;
		uint32_t first = 0, start = 0;
		const uint32_t kept = min(P.n_rays, counters_in[1]);
		if (carry > P.batch && kept > 0) {
			Pcg32 r = P.rng;
			r.advance((int64_t)P.n_rays_global * N_MAX_RANDOM_SAMPLES_PER_RAY);
			(void)r.next_uint(); // (the first draw of this stream rotates K1's allocation order)
			first = r.next_uint() % kept;
			start = block_sums[first / RB_C] + local_bases[first];
		}
		rot[0] = first; rot[1] = start;
	}
}

// (C) gradient pass, :1436-1556: same walk over the first `cn` samples of the ray, now with the ray's colour known. W lanes per ray;
// RB_C = rays per block of the composite kernel (the granularity of block_prefix).
// (256-thread blocks: the kernel needs 62 registers, so a 1024-thread block filled an SM alone and held it until its longest ray finished)
constexpr uint32_t GRAD_BLOCK = 256;
template <uint32_t W>
__global__ void __launch_bounds__(GRAD_BLOCK) loss_gradient_kernel(
	const LossParams P, const uint32_t* __restrict__ counters_in, const float* __restrict__ mean_density_ptr,
	const __half* __restrict__ rgbsigma, const float* __restrict__ rays, uint32_t* __restrict__ numsteps_io, const float* __restrict__ coords_in,
	const RayState* __restrict__ state, const uint32_t* __restrict__ compacted_counts, const uint32_t* __restrict__ local_bases, const uint32_t* __restrict__ block_prefix,
	const uint32_t* __restrict__ rot, const uint32_t* __restrict__ compacted_total, const uint32_t RB_C, float* __restrict__ coords_out, __half* __restrict__ dloss_dout, float* __restrict__ loss_output,
	const uint4* __restrict__ rows_in, uint4* __restrict__ rows_out)
{
	constexpr uint32_t RAYS_PER_WARP = 32 / W, RAYS_PER_BLOCK = GRAD_BLOCK / W;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, glane = lane & (W - 1);
	const uint32_t i = blockIdx.x * RAYS_PER_BLOCK + warp * RAYS_PER_WARP + lane / W;
	const bool in_range = i < P.n_rays;
	const bool kept = in_range && i < counters_in[1];
	uint32_t cn = 0, base = 0, compacted_base = 0;
	if (kept) {
		base = numsteps_io[i * 2 + 1];
		// clip to the batch (:1436-1437)
		compacted_base = block_prefix[i / RB_C] + local_bases[i];
		const uint32_t total = *compacted_total;
		if (total > P.batch) { // overflow: the order starts at slot rot[0] and wraps around (see loss_scan_kernel)
			const uint32_t first = rot[0], start = rot[1];
			compacted_base = i >= first ? compacted_base - start : compacted_base + (total - start);
		}
		cn = min(P.batch - min(P.batch, compacted_base), compacted_counts[i]);
	}
	__syncwarp();
	if (kept && glane == 0) { numsteps_io[i * 2 + 0] = cn; numsteps_io[i * 2 + 1] = compacted_base; }
	if (in_range && cn == 0 && loss_output && glane == 0) loss_output[i] = 0.f;
	RayState s{};
	float ro[3] = {0.f, 0.f, 0.f};
	LossAndGradient lg{};
	if (cn > 0) {
		s = state[i];
		ro[0] = rays[(size_t)i * 6 + 0]; ro[1] = rays[(size_t)i * 6 + 1]; ro[2] = rays[(size_t)i * 6 + 2];
		lg = loss_and_gradient(s.rgbtarget, s.rgb_ray, P.cfg.loss_type);
		// the loss -- not its gradient -- is divided by the density the pixel was sampled with (:1448, :1454-1458; a division by exactly 1 without K19)
		#pragma unroll
		for (int c = 0; c < 3; ++c) lg.loss[c] /= s.pdf;
		const float mean_loss = sum3(lg.loss[0], lg.loss[1], lg.loss[2]) / 3.0f;
		if (loss_output && glane == 0) loss_output[i] = mean_loss / (float)P.n_rays_global;
		// compacted coordinates: a contiguous copy of the ray's first cn records, done by the group
		const float* src = coords_in + (size_t)base * COORD_FLOATS;
		float* dst = coords_out + (size_t)compacted_base * COORD_FLOATS;
		for (uint32_t k = glane; k < cn * COORD_FLOATS; k += W) dst[k] = src[k];
		// ... and of their hash-grid features (64-byte rows), which the training pass would otherwise recompute from the same weights
		if (rows_in) {
			for (uint32_t k = glane; k < cn * 4; k += W) rows_out[feature_chunk(P.rows_tiled, (size_t)compacted_base + (k >> 2), k & 3u)] = rows_in[feature_chunk(P.rows_tiled, (size_t)base + (k >> 2), k & 3u)];
		}
	}
	const float loss_scale = P.cfg.loss_scale / P.n_rays_global;
	const float output_l2_reg = P.cfg.rgb_activation == NGPB_ACT_EXPONENTIAL ? 1e-4f : 0.0f;
	const float output_l1_reg_density = *mean_density_ptr < NERF_MIN_OPTICAL_THICKNESS ? 1e-4f : 0.0f;

	float T_in = 1.f;
	float prefix_rgb[3] = {0.f, 0.f, 0.f}; // rgb_ray2 after the previous round
	for (uint32_t r0 = 0; __any_sync(0xffffffffu, r0 < cn); r0 += W) {
		const uint32_t j = r0 + glane;
		const bool valid = j < cn;
		const size_t idx = (size_t)base + j;
		const SampleEval e = eval_sample(rgbsigma, coords_in, idx, P.cfg, valid);
		const float T_before = sequential_prefix_product<W>(T_in, e.one_minus_alpha, lane);
		const float weight = valid ? e.alpha * T_before : 0.f;
		const float T = T_before * e.one_minus_alpha; // transmittance after this sample (:1520)
		float rgb_ray2[3];
		#pragma unroll
		for (int c = 0; c < 3; ++c) rgb_ray2[c] = prefix_rgb[c] + group_inclusive_sum<W>(weight * e.rgb[c], lane);
		if (valid) {
			const float* ci = coords_in + idx * COORD_FLOATS;
			const float cpos[3] = {ci[0], ci[1], ci[2]};
			const V3 pos = unwarp_position(cpos, P.aabb);
			const float dx = pos.x - ro[0], dy = pos.y - ro[1], dz = pos.z - ro[2];
			const float depth = sqrtf(sum3(dx * dx, dy * dy, dz * dz));
			float tv[3], g[4];
			#pragma unroll
			for (int c = 0; c < 3; ++c) {
				const float suffix = s.rgb_ray[c] - rgb_ray2[c];
				const float dloss_by_drgb = weight * lg.gradient[c];
				g[c] = loss_scale * (dloss_by_drgb * network_to_rgb_derivative(e.o[c], P.cfg.rgb_activation) + fmaxf(0.0f, output_l2_reg * e.o[c]));
				tv[c] = T * e.rgb[c] - suffix;
			}
			const float density_derivative = network_to_density_derivative(e.o[3], P.cfg.density_activation);
			// depth supervision is off (:1450-1452): its term is identically zero
			const float dloss_by_dmlp = density_derivative * (e.dt * dot3(lg.gradient, tv));
			g[3] = loss_scale * dloss_by_dmlp +
				(e.o[3] < 0.0f ? -output_l1_reg_density : 0.0f) +
				(e.o[3] > -10.0f && depth < P.cfg.near_distance ? 1e-4f : 0.0f);
			const __half2 h01 = __halves2half2(__float2half_rn(g[0]), __float2half_rn(g[1]));
			const __half2 h23 = __halves2half2(__float2half_rn(g[2]), __float2half_rn(g[3]));
			uint2 packed;
			packed.x = *reinterpret_cast<const uint32_t*>(&h01);
			packed.y = *reinterpret_cast<const uint32_t*>(&h23);
			*reinterpret_cast<uint2*>(dloss_dout + ((size_t)compacted_base + j) * 4) = packed;
		}
		const uint32_t last = (lane & ~(W - 1)) + W - 1;
		#pragma unroll
		for (int c = 0; c < 3; ++c) prefix_rgb[c] = __shfl_sync(0xffffffffu, rgb_ray2[c], last);
		T_in = __shfl_sync(0xffffffffu, T, last);
	}
}

// (C') gradient of the loss w.r.t. each image's exposure (:1558-1571), one thread per ray, only launched while exposures are being optimised. The loss
// is symmetric in (prediction, target), so d loss / d target = -d loss / d prediction; training in sRGB adds the derivative of the transfer function.
__global__ void __launch_bounds__(128) exposure_gradient_kernel(const LossParams P, const uint32_t* __restrict__ counters_in, const uint32_t* __restrict__ ray_indices,
                                                                const uint32_t* __restrict__ numsteps, const RayState* __restrict__ state,
                                                                const float* __restrict__ exposure, float* __restrict__ exposure_gradient)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_rays || i >= counters_in[1]) return;
	if (numsteps[i * 2 + 0] == 0) return; // rays without a compacted sample leave the kernel before this point (:1438)
	const RayState s = state[i];
	const uint32_t img = image_idx(ray_indices[i], P.n_rays_global, P.n_images, P.cdf.img);
	const LossAndGradient lg = loss_and_gradient(s.rgbtarget, s.rgb_ray, P.cfg.loss_type);
	const float loss_scale = P.cfg.loss_scale / P.n_rays_global;
	#pragma unroll
	for (int c = 0; c < 3; ++c) {
		float dloss_by_dgt = -lg.gradient[c] / s.xy_pdf;
		if (!P.cfg.linear_colors) dloss_by_dgt /= srgb_to_linear_derivative(s.rgbtarget[c]);
		const float es = expf(0.6931471805599453f * exposure[(size_t)img * 3 + c]);
		atomicAdd(&exposure_gradient[(size_t)img * 3 + c], loss_scale * dloss_by_dgt * es * 0.6931471805599453f);
	}
}

// (C'') K19: every ray with a compacted sample deposits its loss (after the division by its sampling density) bilinearly into its image's error map
// (:1465-1491; no sharpness weighting). One thread per ray, only launched while an error map is accumulated; the pixel is recovered from the ray's RNG
// stream like in loss_target_kernel. The texel index is clamped against the IMAGE resolution, as the reference does.
__global__ void __launch_bounds__(128) error_map_deposit_kernel(const LossParams P, const ngpb_image* __restrict__ images, const uint32_t* __restrict__ counters_in,
                                                                const uint32_t* __restrict__ ray_indices, const uint32_t* __restrict__ numsteps, const RayState* __restrict__ state,
                                                                float* __restrict__ error_map, const int res_x, const int res_y)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n_rays || i >= counters_in[1]) return;
	if (numsteps[i * 2 + 0] == 0) return; // (:1438)
	const RayState s = state[i];
	const uint32_t ray_idx = ray_indices[i];
	Pcg32 rng = P.rng;
	rng.advance((int64_t)ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY);
	const uint32_t img = image_idx(ray_idx, P.n_rays_global, P.n_images, P.cdf.img);
	const ngpb_image& im = images[img];
	float x, y;
	random_image_pos_training(rng, im.w, im.h, P.cfg.snap_to_pixel_centers != 0, &x, &y, P.cdf, img);
	LossAndGradient lg = loss_and_gradient(s.rgbtarget, s.rgb_ray, P.cfg.loss_type);
	#pragma unroll
	for (int c = 0; c < 3; ++c) lg.loss[c] /= s.pdf;
	const float mean_loss = sum3(lg.loss[0], lg.loss[1], lg.loss[2]) / 3.0f;
	const float px = fminf(fmaxf(x * (float)res_x - 0.5f, 0.0f), (float)res_x - (1.0f + 1e-4f));
	const float py = fminf(fmaxf(y * (float)res_y - 0.5f, 0.0f), (float)res_y - (1.0f + 1e-4f));
	const int ix = (int)px, iy = (int)py;
	const float wx = px - (float)ix, wy = py - (float)iy;
	const int jx = max(min(ix, im.w - 2), 0), jy = max(min(iy, im.h - 2), 0);
	float* em = error_map + (size_t)img * res_x * res_y;
	atomicAdd(&em[(size_t)jy * res_x + jx], (1 - wx) * (1 - wy) * mean_loss);
	atomicAdd(&em[(size_t)jy * res_x + jx + 1], wx * (1 - wy) * mean_loss);
	atomicAdd(&em[(size_t)(jy + 1) * res_x + jx], (1 - wx) * wy * mean_loss);
	atomicAdd(&em[(size_t)(jy + 1) * res_x + jx + 1], wx * wy * mean_loss);
}

// construct_cdf_2d (:1984-2012): one thread per (row, image) runs the row's cumulative sum serially -- the rounding of the running sum is part of the
// result -- and normalises it with a correctly rounded reciprocal, blended with MIN_PDF = 1 % uniform.
__global__ void __launch_bounds__(128) construct_cdf_2d_kernel(const uint32_t n_images, const uint32_t height, const uint32_t width, const float* __restrict__ data,
                                                               float* __restrict__ cdf_x_cond_y, float* __restrict__ cdf_y)
{
	const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; // = img * height + y
	if (row >= n_images * height) return;
	const size_t off = (size_t)row * width;
	float cum = 0;
	for (uint32_t x = 0; x < width; ++x) { cum += data[off + x] + 1e-10f; cdf_x_cond_y[off + x] = cum; }
	cdf_y[row] = cum;
	const float norm = __frcp_rn(cum);
	for (uint32_t x = 0; x < width; ++x) cdf_x_cond_y[off + x] = (1.0f - 0.01f) * cdf_x_cond_y[off + x] * norm + 0.01f * (float)(x + 1) / (float)width;
}
// construct_cdf_1d (:2014-2037) per image over its row sums, then -- one thread, as on the reference's host (:3000-3015) -- the CDF over the image sums
// with MIN_PMF = 10 % uniform.
__global__ void __launch_bounds__(128) construct_cdf_1d_kernel(const uint32_t n_images, const uint32_t height, float* __restrict__ cdf_y, float* __restrict__ cdf_img)
{
	const uint32_t img = blockIdx.x * blockDim.x + threadIdx.x;
	if (img >= n_images) return;
	float* cy = cdf_y + (size_t)img * height;
	float cum = 0;
	for (uint32_t y = 0; y < height; ++y) { cum += cy[y]; cy[y] = cum; }
	cdf_img[img] = cum;
	const float norm = __frcp_rn(cum);
	for (uint32_t y = 0; y < height; ++y) cy[y] = (1.0f - 0.01f) * cy[y] * norm + 0.01f * (float)(y + 1) / (float)height;
}
__global__ void normalize_image_cdf_kernel(const uint32_t n_images, float* __restrict__ cdf_img, float* __restrict__ pmf_img)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	float total = 0;
	for (uint32_t i = 0; i < n_images; ++i) total += cdf_img[i];
	const float norm = 1.0f / total;
	float cum = 0;
	for (uint32_t i = 0; i < n_images; ++i) {
		const float sum = cdf_img[i];
		cum += sum;
		if (pmf_img) pmf_img[i] = (1.0f - 0.1f) * sum * norm + 0.1f / (float)n_images;
		cdf_img[i] = (1.0f - 0.1f) * cum * norm + 0.1f * (float)(i + 1) / (float)n_images;
	}
}

// (D) roll-over padding of the compacted batch (tcnn common_device.h:517-537): element e >= n_valid copies
// element e % n_valid; gradients of the padded copies are rescaled by n_valid / batch.
__global__ void __launch_bounds__(256) rollover_kernel(const uint32_t batch, const uint32_t* __restrict__ counters_out, float* __restrict__ coords, __half* __restrict__ dloss_dout,
                                                       uint4* __restrict__ rows, const uint32_t rows_tiled)
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
	if (rows) {
		#pragma unroll
		for (int k = 0; k < 4; ++k) rows[feature_chunk(rows_tiled, e, k)] = rows[feature_chunk(rows_tiled, src, k)];
	}
}

Aabb make_aabb(const float* a);
int compute_loss_launch(void* stream_, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                        uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                        const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                        const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                        const ngpb_half* encoded_in, ngpb_half* encoded_out, bool rows_tiled, const float* exposure, float* exposure_gradient,
                        const ErrorCdf& cdf, float* error_map, int error_map_res_x, int error_map_res_y);

} // namespace ngpb

using namespace ngpb;

extern "C" uint64_t ngpb_compute_loss_scratch_bytes(uint32_t n_rays) {
	return (uint64_t)n_rays * (sizeof(RayState) + 8) + (uint64_t)div_round_up(n_rays, LOSS_RAYS_PER_BLOCK) * 4 + 64;
}

extern "C" int ngpb_compute_loss(void* stream, uint32_t n_rays, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch) {
	return ngpb_compute_loss_sharded(stream, n_rays, n_rays, aabb6, rng, batch, cfg, n_images, images_dev, counters_in, rgbsigma, ray_indices, rays, numsteps, coords_in,
		mean_density_dev, coords_out, dloss_dout, loss_per_ray, counters_out, scratch);
}

extern "C" int ngpb_compute_loss_sharded(void* stream, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch) {
	return ngpb_compute_loss_compact_features(stream, n_rays, n_rays_global, aabb6, rng, batch, cfg, n_images, images_dev, counters_in, rgbsigma, ray_indices, rays, numsteps, coords_in,
		mean_density_dev, coords_out, dloss_dout, loss_per_ray, counters_out, scratch, nullptr, nullptr);
}

extern "C" int ngpb_compute_loss_compact_features(void* stream_, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                                 const ngpb_half* encoded_in, ngpb_half* encoded_out) {
	return ngpb::compute_loss_launch(stream_, n_rays, n_rays_global, aabb6, rng_, batch, cfg, n_images, images_dev, counters_in, rgbsigma, ray_indices, rays, numsteps, coords_in,
		mean_density_dev, coords_out, dloss_dout, loss_per_ray, counters_out, scratch, encoded_in, encoded_out, false, nullptr, nullptr, no_error_cdf(), nullptr, 0, 0);
}

extern "C" int ngpb_compute_loss_exposure(void* stream_, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                                 const float* exposure_dev, float* exposure_gradient_dev) {
	return ngpb::compute_loss_launch(stream_, n_rays, n_rays_global, aabb6, rng_, batch, cfg, n_images, images_dev, counters_in, rgbsigma, ray_indices, rays, numsteps, coords_in,
		mean_density_dev, coords_out, dloss_dout, loss_per_ray, counters_out, scratch, nullptr, nullptr, false, exposure_dev, exposure_gradient_dev, no_error_cdf(), nullptr, 0, 0);
}

extern "C" int ngpb_compute_loss_error_map(void* stream_, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                                 const float* exposure_dev, float* exposure_gradient_dev,
                                 const ngpb_error_cdf* error_cdf, float* error_map_dev, int32_t error_map_res_x, int32_t error_map_res_y) {
	return ngpb::compute_loss_launch(stream_, n_rays, n_rays_global, aabb6, rng_, batch, cfg, n_images, images_dev, counters_in, rgbsigma, ray_indices, rays, numsteps, coords_in,
		mean_density_dev, coords_out, dloss_dout, loss_per_ray, counters_out, scratch, nullptr, nullptr, false, exposure_dev, exposure_gradient_dev,
		make_error_cdf(error_cdf), error_map_dev, error_map_res_x, error_map_res_y);
}

extern "C" int ngpb_construct_error_cdfs(void* stream_, uint32_t n_images, uint32_t res_y, uint32_t res_x, const float* error_map_dev,
                                         float* cdf_x_cond_y, float* cdf_y, float* cdf_img, float* pmf_img) {
	try {
		if (!error_map_dev || !cdf_x_cond_y || !cdf_y || !cdf_img || n_images == 0 || res_x == 0 || res_y == 0) {
			set_last_error("ngpb_construct_error_cdfs: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		construct_cdf_2d_kernel<<<div_round_up(n_images * res_y, 128), 128, 0, stream>>>(n_images, res_y, res_x, error_map_dev, cdf_x_cond_y, cdf_y);
		NGPB_LAUNCH_CHECK();
		construct_cdf_1d_kernel<<<div_round_up(n_images, 128), 128, 0, stream>>>(n_images, res_y, cdf_y, cdf_img);
		NGPB_LAUNCH_CHECK();
		normalize_image_cdf_kernel<<<1, 32, 0, stream>>>(n_images, cdf_img, pmf_img);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

// The implementation behind the C entry points; `rows_tiled` selects the layout of encoded_in / encoded_out (the testbed hands tiles from the hash-grid kernel to the MLP kernel).
int ngpb::compute_loss_launch(void* stream_, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                                 uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                                 const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                                 const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                                 const ngpb_half* encoded_in, ngpb_half* encoded_out, bool rows_tiled, const float* exposure, float* exposure_gradient,
                                 const ErrorCdf& cdf, float* error_map, int error_map_res_x, int error_map_res_y) {
	try {
		if (cdf.x_cond_y && (!cdf.y || cdf.res_x <= 0 || cdf.res_y <= 0)) { set_last_error("ngpb_compute_loss: cdf_x_cond_y needs cdf_y and a resolution"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (error_map && (error_map_res_x < 2 || error_map_res_y < 2)) { set_last_error("ngpb_compute_loss: the error map needs at least 2 x 2 texels"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (exposure_gradient && !exposure) { set_last_error("ngpb_compute_loss: an exposure gradient needs the exposures"); return NGPB_ERR_INVALID_ARGUMENT; }
		if ((encoded_in == nullptr) != (encoded_out == nullptr) || (encoded_in && encoded_in == encoded_out)) {
			set_last_error("ngpb_compute_loss: encoded_in and encoded_out must both be given, and differ");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		if (!aabb6 || !cfg || !images_dev || !counters_in || !rgbsigma || !ray_indices || !rays || !numsteps || !coords_in || !mean_density_dev ||
		    !coords_out || !dloss_dout || !counters_out || !scratch || n_images == 0 || batch == 0) {
			set_last_error("ngpb_compute_loss: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		if (n_rays == 0) { NGPB_CUDA_CHECK(cudaMemsetAsync(counters_out, 0, sizeof(uint32_t), stream)); return 0; }
		LossParams P;
		P.rows_tiled = rows_tiled ? 1u : 0u;
		P.n_rays = n_rays; P.batch = batch; P.n_images = n_images; P.n_rays_global = n_rays_global;
		P.aabb = make_aabb(aabb6);
		P.rng.state = rng_.state; P.rng.inc = rng_.inc;
		P.cfg = *cfg;
		P.cdf = cdf;
		// scratch layout: RayState[n_rays] (32 B each) | uint32 counts[n_rays] | uint32 local_bases[n_rays] | uint32 block_sums[n_blocks]
		RayState* state = reinterpret_cast<RayState*>(scratch);
		uint32_t* counts = reinterpret_cast<uint32_t*>(state + n_rays);
		uint32_t* local_bases = counts + n_rays;
		uint32_t* block_sums = local_bases + n_rays;
		// lanes per ray for the composite / gradient pass. Measured on the Lego-shaped scene (loss stage, us): 32,32 -> 86; 16,16 -> 83; 16,8 -> 99; 8,8 -> 109:
		// long rays dominate (a warp runs as many rounds as its longest ray), so narrow groups lose. NGPB_LOSS_LANES="c,g" overrides.
		static const std::pair<int, int> lanes = [] {
			int c = 16, g = 16;
			if (const char* e = std::getenv("NGPB_LOSS_LANES")) std::sscanf(e, "%d,%d", &c, &g);
			return std::make_pair(c, g);
		}();
		const uint32_t wc = lanes.first == 8 ? 8u : lanes.first == 32 ? 32u : 16u, wg = lanes.second == 16 ? 16u : lanes.second == 32 ? 32u : 8u;
		const uint32_t rb_c = 1024 / wc, blocks_c = div_round_up(n_rays, rb_c), blocks_g = div_round_up(n_rays, GRAD_BLOCK / wg);
		loss_target_kernel<<<div_round_up(n_rays, 128), 128, 0, stream>>>(P, images_dev, counters_in, ray_indices, state, exposure);
		NGPB_LAUNCH_CHECK();
		#define NGPB_COMPOSITE(W) loss_composite_kernel<W><<<blocks_c, 1024, 0, stream>>>(P, images_dev, counters_in, (const __half*)rgbsigma, ray_indices, numsteps, coords_in, state, counts, local_bases, block_sums)
		if (wc == 8) NGPB_COMPOSITE(8); else if (wc == 16) NGPB_COMPOSITE(16); else NGPB_COMPOSITE(32);
		#undef NGPB_COMPOSITE
		NGPB_LAUNCH_CHECK();
		uint32_t* rot = block_sums + blocks_c; // (two words inside the scratch's slack, see ngpb_compute_loss_scratch_bytes)
		loss_scan_kernel<<<1, 1024, 0, stream>>>(P, counters_in, blocks_c, block_sums, local_bases, rb_c, counters_out, rot);
		NGPB_LAUNCH_CHECK();
		#define NGPB_GRADIENT(W) loss_gradient_kernel<W><<<blocks_g, GRAD_BLOCK, 0, stream>>>(P, counters_in, mean_density_dev, (const __half*)rgbsigma, rays, numsteps, coords_in, state, counts, local_bases, \
			block_sums, rot, counters_out, rb_c, coords_out, (__half*)dloss_dout, loss_per_ray, reinterpret_cast<const uint4*>(encoded_in), reinterpret_cast<uint4*>(encoded_out))
		if (wg == 8) NGPB_GRADIENT(8); else if (wg == 16) NGPB_GRADIENT(16); else NGPB_GRADIENT(32);
		#undef NGPB_GRADIENT
		NGPB_LAUNCH_CHECK();
		if (exposure_gradient) {
			exposure_gradient_kernel<<<div_round_up(n_rays, 128), 128, 0, stream>>>(P, counters_in, ray_indices, numsteps, state, exposure, exposure_gradient);
			NGPB_LAUNCH_CHECK();
		}
		if (error_map) {
			error_map_deposit_kernel<<<div_round_up(n_rays, 128), 128, 0, stream>>>(P, images_dev, counters_in, ray_indices, numsteps, state, error_map, error_map_res_x, error_map_res_y);
			NGPB_LAUNCH_CHECK();
		}
		rollover_kernel<<<div_round_up(batch, 256), 256, 0, stream>>>(batch, counters_out, coords_out, (__half*)dloss_dout, reinterpret_cast<uint4*>(encoded_out), P.rows_tiled);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
