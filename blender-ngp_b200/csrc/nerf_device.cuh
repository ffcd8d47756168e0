// Device-side building blocks of the NeRF sampling / loss / render kernels.
// This library is compiled with -fmad=false: every float op below rounds exactly as written, which is
// what makes sample compaction reproducible bit for bit against the CPU oracle. Reference sections are
// cited per function (paths relative to the reference root).
#pragma once
#include "common.cuh"
#include "../../include/ngpb.h"

namespace ngpb {

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fmaxf(lo, fminf(v, hi)); }

// Eigen reduces three terms as a + (b + c) (Eigen/src/Core/Redux.h:101-115); kept so positions match.
__device__ __forceinline__ float sum3(float a, float b, float c) { return a + (b + c); }
__device__ __forceinline__ float dot3(const float* a, const float* b) { return sum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }

struct V3 { float x, y, z; };

// ---- step-size law and occupancy lookup: src/testbed_nerf.cu:96-98,:191-213,:318-342,:449-463 ----
__device__ __forceinline__ float calc_dt(float t, float cone_angle) { return clampf(t * cone_angle, MIN_CONE_STEPSIZE, MAX_CONE_STEPSIZE); }

__device__ __forceinline__ float distance_to_next_voxel(const V3& pos, const V3& dir, const V3& idir, uint32_t res) {
	const float r = (float)res;
	const float px = r * pos.x, py = r * pos.y, pz = r * pos.z;
	const float tx = (floorf(px + 0.5f + 0.5f * copysignf(1.0f, dir.x)) - px) * idir.x;
	const float ty = (floorf(py + 0.5f + 0.5f * copysignf(1.0f, dir.y)) - py) * idir.y;
	const float tz = (floorf(pz + 0.5f + 0.5f * copysignf(1.0f, dir.z)) - pz) * idir.z;
	const float t = fminf(fminf(tx, ty), tz);
	// res is a power of two (NERF_GRIDSIZE >> mip), so t / res and t * (1 / res) are the same real number before rounding
	return fmaxf(t * __uint_as_float((127u - (uint32_t)(31 - __clz((int)res))) << 23), 0.0f);
}

__device__ __forceinline__ float advance_to_next_voxel(float t, float cone_angle, const V3& pos, const V3& dir, const V3& idir, uint32_t res) {
	const float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
	do { t += calc_dt(t, cone_angle); } while (t < t_target);
	return t;
}

// The constant-step t chain in closed form: t_{k+1} = fl(t_k + dt) with dt = MIN_CONE_STEPSIZE (cone_angle == 0), advanced by n steps with the SAME bits as n
// additions. While t stays inside one binade every t_k is a multiple of u = ulp(t) and dt = q u + r, so every addition rounds the same way and moves the
// mantissa by the same integer (q, or q + 1 when r > u / 2): k steps are one multiply-add on the mantissa, and only a step that leaves the binade is a
// real addition. A tie (r == u / 2) would round by parity; for this dt it only occurs for t in [2^-9, 2^-8), which is stepped one addition at a time.
// Checked against step-by-step addition for 200 000 random (t, n).
__device__ __forceinline__ float chain_advance(float t, uint32_t n) {
	const uint32_t dt_bits = __float_as_uint(MIN_CONE_STEPSIZE);
	const uint32_t e_dt = dt_bits >> 23, m_dt = (dt_bits & 0x7FFFFFu) | 0x800000u;
	while (n) {
		const uint32_t tb = __float_as_uint(t);
		const uint32_t e_t = tb >> 23; // t >= 0
		const uint32_t s = e_t - e_dt;
		bool fast = e_t >= e_dt && s <= 23u;
		uint32_t inc = 0;
		if (fast) {
			if (s == 0) inc = m_dt;
			else { const uint32_t rb = m_dt & ((1u << s) - 1u), half = 1u << (s - 1); fast = rb != half; inc = (m_dt >> s) + (rb > half ? 1u : 0u); }
		}
		if (!fast) { t += MIN_CONE_STEPSIZE; --n; continue; }
		const uint32_t m = (tb & 0x7FFFFFu) | 0x800000u;
		const uint32_t room = 0xFFFFFFu - m;
		uint32_t k = n;
		if ((uint64_t)n * inc > room) k = room / inc;
		t = __uint_as_float((e_t << 23) | ((m + k * inc) & 0x7FFFFFu));
		n -= k;
		if (n) { t += MIN_CONE_STEPSIZE; --n; } // the step that leaves the binade
	}
	return t;
}
// do { t += dt; } while (t < t_target) for the constant step: the first chain member (at least one step on) at or past t_target, same bits
__device__ __forceinline__ float chain_advance_to(float t, float t_target) {
	const float est = (t_target - t) * (1.0f / MIN_CONE_STEPSIZE);
	uint32_t n = est < 1.0f ? 1u : (est < 1.0e6f ? (uint32_t)ceilf(est) : 1000000u);
	float ta = n > 1 ? chain_advance(t, n - 1) : t;
	while (n > 1 && ta >= t_target) { --n; ta = n > 1 ? chain_advance(t, n - 1) : t; }
	float tb = ta + MIN_CONE_STEPSIZE;
	for (int k = 0; k < 64 && tb < t_target; ++k) tb += MIN_CONE_STEPSIZE; // (the estimate is off by a step at most; the bound only guards non-finite targets)
	return tb;
}

// frexpf exponent of a non-negative finite float, as used by mip_from_pos / mip_from_dt: v = m * 2^e with m in [0.5, 1);
// frexpf(0) reports exponent 0. Read from the bit pattern instead of calling the library routine. Subnormal inputs (|v| < 2^-126)
// report -126 instead of their true, smaller exponent: every caller clamps the result from below at 0, so it is indistinguishable.
__device__ __forceinline__ int frexp_exponent(float v) {
	const uint32_t b = __float_as_uint(v) & 0x7FFFFFFFu;
	return b == 0u ? 0 : (int)(b >> 23) - 126;
}

// 2^-mip for mip in [0, 126], exact (what scalbnf(1.0f, -mip) returns)
__device__ __forceinline__ float exp2_neg(uint32_t mip) { return __uint_as_float((127u - mip) << 23); }

__device__ __forceinline__ int mip_from_pos(const V3& pos, uint32_t max_cascade = NERF_CASCADES - 1) {
	const float maxval = fmaxf(fmaxf(fabsf(pos.x - 0.5f), fabsf(pos.y - 0.5f)), fabsf(pos.z - 0.5f));
	return min((int)max_cascade, max(0, frexp_exponent(maxval) + 1));
}

__device__ __forceinline__ int mip_from_dt(float dt, const V3& pos, uint32_t max_cascade = NERF_CASCADES - 1) {
	const int mip = mip_from_pos(pos, max_cascade);
	dt *= 2 * NERF_GRIDSIZE;
	if (dt < 1.f) return mip;
	return min((int)max_cascade, max(frexp_exponent(dt), mip));
}

__device__ __forceinline__ uint32_t cascaded_grid_idx_at(V3 pos, uint32_t mip) {
	const float mip_scale = exp2_neg(mip);
	pos.x -= 0.5f; pos.y -= 0.5f; pos.z -= 0.5f;
	pos.x *= mip_scale; pos.y *= mip_scale; pos.z *= mip_scale;
	pos.x += 0.5f; pos.y += 0.5f; pos.z += 0.5f;
	const int ix = (int)(pos.x * NERF_GRIDSIZE), iy = (int)(pos.y * NERF_GRIDSIZE), iz = (int)(pos.z * NERF_GRIDSIZE);
	return morton3D(min(max(ix, 0), (int)NERF_GRIDSIZE - 1), min(max(iy, 0), (int)NERF_GRIDSIZE - 1), min(max(iz, 0), (int)NERF_GRIDSIZE - 1));
}

__device__ __forceinline__ uint32_t grid_mip_offset(uint32_t mip) { return NERF_GRID_CELLS * mip; }

__device__ __forceinline__ bool density_grid_occupied_at(const V3& pos, const uint8_t* __restrict__ bitfield, uint32_t mip) {
	const uint32_t idx = cascaded_grid_idx_at(pos, mip);
	return bitfield[idx / 8 + grid_mip_offset(mip) / 8] & (1 << (idx % 8));
}

__device__ __forceinline__ float warp_dt(float dt) {
	const float max_stepsize = MIN_CONE_STEPSIZE * (1 << (NERF_CASCADES - 1));
	return (dt - MIN_CONE_STEPSIZE) / (max_stepsize - MIN_CONE_STEPSIZE);
}
__device__ __forceinline__ float unwarp_dt(float dt) {
	const float max_stepsize = MIN_CONE_STEPSIZE * (1 << (NERF_CASCADES - 1));
	return dt * (max_stepsize - MIN_CONE_STEPSIZE) + MIN_CONE_STEPSIZE;
}

// ---- bounding box: include/neural-graphics-primitives/bounding_box.cuh:86,:163-221 ----
__device__ __forceinline__ bool aabb_contains(const Aabb& b, const V3& p) {
	return p.x >= b.min[0] && p.x <= b.max[0] && p.y >= b.min[1] && p.y <= b.max[1] && p.z >= b.min[2] && p.z <= b.max[2];
}
__device__ __forceinline__ void swapf(float& a, float& b) { float t = a; a = b; b = t; }
__device__ inline void aabb_ray_intersect(const Aabb& b, const V3& pos, const V3& dir, float* out_tmin, float* out_tmax) {
	const float FMAX = 3.402823466e+38f;
	float tmin = (b.min[0] - pos.x) / dir.x, tmax = (b.max[0] - pos.x) / dir.x;
	if (tmin > tmax) swapf(tmin, tmax);
	float tymin = (b.min[1] - pos.y) / dir.y, tymax = (b.max[1] - pos.y) / dir.y;
	if (tymin > tymax) swapf(tymin, tymax);
	if (tmin > tymax || tymin > tmax) { *out_tmin = FMAX; *out_tmax = FMAX; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (b.min[2] - pos.z) / dir.z, tzmax = (b.max[2] - pos.z) / dir.z;
	if (tzmin > tzmax) swapf(tzmin, tzmax);
	if (tmin > tzmax || tzmin > tmax) { *out_tmin = FMAX; *out_tmax = FMAX; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	*out_tmin = tmin; *out_tmax = tmax;
}
__device__ __forceinline__ V3 warp_position(const V3& p, const Aabb& b) {
	return {(p.x - b.min[0]) / (b.max[0] - b.min[0]), (p.y - b.min[1]) / (b.max[1] - b.min[1]), (p.z - b.min[2]) / (b.max[2] - b.min[2])};
}
__device__ __forceinline__ V3 unwarp_position(const float* p, const Aabb& b) { // testbed_nerf.cu:274-279
	return {b.min[0] + p[0] * (b.max[0] - b.min[0]), b.min[1] + p[1] * (b.max[1] - b.min[1]), b.min[2] + p[2] * (b.max[2] - b.min[2])};
}

// ---- colour transfer: include/neural-graphics-primitives/common_device.cuh:31-77 ----
__device__ __forceinline__ float srgb_to_linear(float srgb) {
	if (srgb <= 0.04045f) return srgb / 12.92f;
	return powf((srgb + 0.055f) / 1.055f, 2.4f);
}
__device__ __forceinline__ float srgb_to_linear_derivative(float srgb) {
	if (srgb <= 0.04045f) return 1.0f / 12.92f;
	return 2.4f / 1.055f * powf((srgb + 0.055f) / 1.055f, 1.4f);
}
__device__ __forceinline__ float linear_to_srgb(float linear) {
	if (linear < 0.0031308f) return 12.92f * linear;
	return 1.055f * powf(linear, 0.41666f) - 0.055f;
}
__device__ __forceinline__ float logistic(float x) { return 1.0f / (1.0f + expf(-x)); } // tcnn common_device.h:51

// ---- activations: src/testbed_nerf.cu:215-257 ----
__device__ __forceinline__ float network_to_rgb(float v, int act) {
	switch (act) {
		case NGPB_ACT_NONE: return v;
		case NGPB_ACT_RELU: return v > 0.0f ? v : 0.0f;
		case NGPB_ACT_LOGISTIC: return logistic(v);
		default: return __expf(clampf(v, -10.0f, 10.0f));
	}
}
__device__ __forceinline__ float network_to_rgb_derivative(float v, int act) {
	switch (act) {
		case NGPB_ACT_NONE: return 1.0f;
		case NGPB_ACT_RELU: return v > 0.0f ? 1.0f : 0.0f;
		case NGPB_ACT_LOGISTIC: { float d = logistic(v); return d * (1 - d); }
		default: return __expf(clampf(v, -10.0f, 10.0f));
	}
}
__device__ __forceinline__ float network_to_density(float v, int act) {
	switch (act) {
		case NGPB_ACT_NONE: return v;
		case NGPB_ACT_RELU: return v > 0.0f ? v : 0.0f;
		case NGPB_ACT_LOGISTIC: return logistic(v);
		default: return __expf(v);
	}
}
__device__ __forceinline__ float network_to_density_derivative(float v, int act) {
	switch (act) {
		case NGPB_ACT_NONE: return 1.0f;
		case NGPB_ACT_RELU: return v > 0.0f ? 1.0f : 0.0f;
		case NGPB_ACT_LOGISTIC: { float d = logistic(v); return d * (1 - d); }
		default: return __expf(clampf(v, -15.0f, 15.0f));
	}
}

// ---- training image access: common_device.cuh:633-705 ----
__device__ __forceinline__ void image_pos(float x, float y, int w, int h, int* px, int* py) {
	*px = max(min((int)(x * (float)w), w - 1), 0);
	*py = max(min((int)(y * (float)h), h - 1), 0);
}
// returns false for masked-away pixels (Byte: 0x00FF00FF, for which read_rgba returns -1; Half / Float: a negative red channel -- K1 tests `.x() < 0`, :1126)
__device__ inline bool read_rgba(float x, float y, const ngpb_image& im, float out[4]) {
	int px, py;
	image_pos(x, y, im.w, im.h, &px, &py);
	const size_t idx = (size_t)px + (size_t)py * im.w;
	if (im.image_type == NGPB_IMAGE_FLOAT) { // linear, premultiplied alpha, as stored
		const float4 v = reinterpret_cast<const float4*>(im.pixels)[idx];
		out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
		return v.x >= 0.0f;
	}
	if (im.image_type == NGPB_IMAGE_HALF) {
		const uint2 raw = reinterpret_cast<const uint2*>(im.pixels)[idx];
		const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
		out[0] = __low2float(a); out[1] = __high2float(a); out[2] = __low2float(b); out[3] = __high2float(b);
		return out[0] >= 0.0f;
	}
	const uint32_t packed = reinterpret_cast<const uint32_t*>(im.pixels)[idx];
	if (packed == 0x00FF00FFu) { out[0] = out[1] = out[2] = out[3] = -1.0f; return false; }
	const float alpha = (float)(packed >> 24) * (1.0f / 255.0f);
	out[0] = srgb_to_linear((float)(packed & 0xFF) * (1.0f / 255.0f)) * alpha;
	out[1] = srgb_to_linear((float)((packed >> 8) & 0xFF) * (1.0f / 255.0f)) * alpha;
	out[2] = srgb_to_linear((float)((packed >> 16) & 0xFF) * (1.0f / 255.0f)) * alpha;
	out[3] = alpha;
	return true;
}

// ---- scrambled Sobol building blocks: include/neural-graphics-primitives/random_val.cuh:159-258 (dimension 0 of the sequence is a bit reversal) ----
__host__ __device__ inline uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
__host__ __device__ inline uint32_t reverse_bits(uint32_t x) {
	x = (((x & 0xaaaaaaaa) >> 1) | ((x & 0x55555555) << 1));
	x = (((x & 0xcccccccc) >> 2) | ((x & 0x33333333) << 2));
	x = (((x & 0xf0f0f0f0) >> 4) | ((x & 0x0f0f0f0f) << 4));
	x = (((x & 0xff00ff00) >> 8) | ((x & 0x00ff00ff) << 8));
	return ((x >> 16) | (x << 16));
}
__host__ __device__ inline uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
	x += seed; x ^= x * 0x6c50b47cu; x ^= x * 0xb82f1e52u; x ^= x * 0xc7afe638u; x ^= x * 0x8d22f6e6u;
	return x;
}
__host__ __device__ inline uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) { return reverse_bits(laine_karras_permutation(reverse_bits(x), seed)); }
// ld_random_val(index, seed, dim = 0) (random_val.cuh:261-268)
__host__ __device__ inline float ld_random_val_dim0(uint32_t index, uint32_t seed) {
	const float S = float(1.0 / (1ull << 32));
	index = nested_uniform_scramble_base2(index, seed);
	return (float)nested_uniform_scramble_base2(reverse_bits(index), hash_combine(seed, 0u)) * S;
}

// ---- K19: importance sampling of training pixels / images by accumulated error (Testbed::Nerf::Training::ErrorMap, testbed.h:600-615) ----
// The CDFs K1, K6 and the camera-gradient kernel draw from; a null pointer switches the respective importance sampling off (uniform pixels,
// round-robin images), which is the reference's behaviour unless nerf.training.sample_*_proportional_to_error is set.
__host__ inline ErrorCdf make_error_cdf(const ngpb_error_cdf* c) { return c ? ErrorCdf{c->cdf_x_cond_y, c->cdf_y, c->cdf_img, c->res_x, c->res_y} : no_error_cdf(); }

// binary_search: include/neural-graphics-primitives/common.h:201-223 (first element not below val, clamped to the last)
__device__ __forceinline__ uint32_t binary_search(float val, const float* __restrict__ data, uint32_t length) {
	if (length == 0) return 0;
	uint32_t it, count = length, step, first = 0;
	while (count > 0) {
		it = first; step = count / 2; it += step;
		if (data[it] < val) { first = ++it; count -= step + 1; } else count = step;
	}
	return min(first, length - 1);
}

// image_idx: src/testbed_nerf.cu:1062-1083. pdf (optional) receives the image's sampling density relative to uniform.
__device__ __forceinline__ uint32_t image_idx(uint32_t base_idx, uint32_t n_rays, uint32_t n_training_images, const float* __restrict__ cdf_img = nullptr, float* pdf = nullptr) {
	if (cdf_img) {
		const float sample = ld_random_val_dim0(base_idx, 0xdeadbeefu);
		const uint32_t img = binary_search(sample, cdf_img, n_training_images);
		if (pdf) { const float prev = img > 0 ? cdf_img[img - 1] : 0.0f; *pdf = (cdf_img[img] - prev) * (float)n_training_images; }
		return img;
	}
	if (pdf) *pdf = 1.0f;
	return ((base_idx * n_training_images) / n_rays) % n_training_images;
}

// sample_cdf_2d: src/testbed_nerf.cu:991-1022. Half of the samples stay uniform (UNIFORM_SAMPLING_FRACTION; their pdf is left untouched, i.e. the
// caller's initial 1); the others pick a row from cdf_y, then a column from that row's CDF, and are placed inside the chosen error-map texel.
__device__ inline void sample_cdf_2d(float* sx, float* sy, uint32_t img, const ErrorCdf& cdf, float* pdf) {
	if (*sx < 0.5f) { *sx /= 0.5f; return; }
	*sx = (*sx - 0.5f) / (1.0f - 0.5f);
	const float* cdf_y = cdf.y + (size_t)img * cdf.res_y;
	const uint32_t y = binary_search(*sy, cdf_y, (uint32_t)cdf.res_y);
	float prev = y > 0 ? cdf_y[y - 1] : 0.0f;
	const float pmf_y = cdf_y[y] - prev;
	*sy = (*sy - prev) / pmf_y;
	const float* cdf_x = cdf.x_cond_y + ((size_t)img * cdf.res_y + y) * cdf.res_x;
	const uint32_t x = binary_search(*sx, cdf_x, (uint32_t)cdf.res_x);
	prev = x > 0 ? cdf_x[x - 1] : 0.0f;
	const float pmf_x = cdf_x[x] - prev;
	*sx = (*sx - prev) / pmf_x;
	if (pdf) *pdf = pmf_x * pmf_y * (float)(cdf.res_x * cdf.res_y);
	*sx = ((float)x + *sx) / (float)cdf.res_x;
	*sy = ((float)y + *sy) / (float)cdf.res_y;
}

// ---- lens models of the training cameras (common_device.cuh:141-200,:236-258; used at testbed_nerf.cu:1166-1190) ----
__device__ __forceinline__ void opencv_lens_distortion(const float* prm, float u, float v, float* du, float* dv) { // apply_opencv_lens_distortion :141-159
	const float k1 = prm[0], k2 = prm[1], p1 = prm[2], p2 = prm[3];
	const float u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
	const float radial = k1 * r2 + k2 * r2 * r2;
	*du = u * radial + 2.0f * p1 * uv + p2 * (r2 + 2.0f * u2);
	*dv = v * radial + 2.0f * p2 * uv + p1 * (r2 + 2.0f * v2);
}
// iterative_opencv_lens_undistortion (:161-200): Newton iteration with central differences, at most 100 steps; operation order as written there
// (Eigen's 2x2 inverse: adjugate times 1 / determinant)
__device__ inline void opencv_lens_undistortion(const float* prm, float* u, float* v) {
	const float x00 = *u, x01 = *v;
	float x0 = *u, x1 = *v;
	for (uint32_t i = 0; i < 100; ++i) {
		const float step0 = fmaxf(1.1920928955078125e-07f, fabsf(1e-6f * x0)), step1 = fmaxf(1.1920928955078125e-07f, fabsf(1e-6f * x1));
		float dx0, dx1, b00, b01, f00, f01, b10, b11, f10, f11;
		opencv_lens_distortion(prm, x0, x1, &dx0, &dx1);
		opencv_lens_distortion(prm, x0 - step0, x1, &b00, &b01);
		opencv_lens_distortion(prm, x0 + step0, x1, &f00, &f01);
		opencv_lens_distortion(prm, x0, x1 - step1, &b10, &b11);
		opencv_lens_distortion(prm, x0, x1 + step1, &f10, &f11);
		const float j00 = 1 + (f00 - b00) / (2 * step0), j01 = (f10 - b10) / (2 * step1), j10 = (f01 - b01) / (2 * step0), j11 = 1 + (f11 - b11) / (2 * step1);
		const float invdet = 1.0f / (j00 * j11 - j10 * j01);
		const float i00 = j11 * invdet, i10 = -j10 * invdet, i01 = -j01 * invdet, i11 = j00 * invdet;
		const float r0 = (x0 + dx0) - x00, r1 = (x1 + dx1) - x01;
		const float s0 = i00 * r0 + i01 * r1, s1 = i10 * r0 + i11 * r1;
		x0 -= s0; x1 -= s1;
		if (s0 * s0 + s1 * s1 < 1e-10f) break;
	}
	*u = x0; *v = x1;
}
__device__ inline V3 f_theta_undistortion(float u, float v, const float* prm, const V3& error_direction) { // :236-249
	const float xpix = u * prm[5], ypix = v * prm[6];
	const float norm = sqrtf(xpix * xpix + ypix * ypix);
	const float alpha = prm[0] + norm * (prm[1] + norm * (prm[2] + norm * (prm[3] + norm * prm[4])));
	float sin_alpha, cos_alpha;
	sincosf(alpha, &sin_alpha, &cos_alpha);
	if (cos_alpha <= 1.17549435e-38f || norm == 0.f) return error_direction;
	sin_alpha *= 1.f / norm;
	return {sin_alpha * xpix, sin_alpha * ypix, cos_alpha};
}
__device__ inline V3 latlong_to_dir(float u, float v) { // :251-258
	const float PI = 3.14159265358979323846f;
	const float theta = (v - 0.5f) * PI, phi = (u - 0.5f) * PI * 2.0f;
	float sp, cp, st, ct;
	sincosf(theta, &st, &ct);
	sincosf(phi, &sp, &cp);
	return {sp * ct, st, cp * ct};
}
// camera-space direction of pixel (x, y) in [0,1)^2 for the image's lens (testbed_nerf.cu:1166-1184), not normalised
__device__ inline V3 training_ray_direction(const ngpb_image& im, float x, float y) {
	if (im.lens_mode == NGPB_LENS_FTHETA) return f_theta_undistortion(x - im.cx, y - im.cy, im.lens_params, V3{0.f, 0.f, 1.f});
	if (im.lens_mode == NGPB_LENS_LATLONG) return latlong_to_dir(x, y);
	V3 d = {(x - im.cx) * (float)im.w / im.fx, (y - im.cy) * (float)im.h / im.fy, 1.0f};
	if (im.lens_mode == NGPB_LENS_OPENCV) opencv_lens_undistortion(im.lens_params, &d.x, &d.y);
	return d;
}

// nerf_random_image_pos_training: src/testbed_nerf.cu:1047-1060
__device__ __forceinline__ void random_image_pos_training(Pcg32& rng, int w, int h, bool snap, float* x, float* y,
                                                          const ErrorCdf& cdf = ErrorCdf{nullptr, nullptr, nullptr, 0, 0}, uint32_t img = 0, float* pdf = nullptr) {
	float u = rng.next_float(), v = rng.next_float();
	if (pdf) *pdf = 1.0f;
	if (cdf.x_cond_y) sample_cdf_2d(&u, &v, img, cdf, pdf);
	if (snap) {
		u = ((float)min(max((int)(u * (float)w), 0), w - 1) + 0.5f) / (float)w;
		v = ((float)min(max((int)(v * (float)h), 0), h - 1) + 0.5f) / (float)h;
	}
	*x = u; *y = v;
}

// ---- block-wide exclusive scan over a device array by ONE block of 1024 threads ----
// out[i] = sum of in[0..i); returns total in *total. n <= 1024 * items_per_thread capacity is
// handled by chunking: each thread owns a contiguous chunk.
__device__ inline uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* smem /*>=33*/, uint32_t* total) {
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
	if (lane == 31) smem[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = smem[lane];
		uint32_t wi = w;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
		smem[lane] = wi - w;
		if (lane == 31) smem[32] = wi;
	}
	__syncthreads();
	const uint32_t excl = incl - v + smem[warp];
	*total = smem[32];
	__syncthreads();
	return excl;
}

} // namespace ngpb
