// TEST INFRASTRUCTURE ONLY -- see ngp_oracle.h. CPU restatement of the reference's NeRF hot path.
// Built with -ffp-contract=off so every float op rounds exactly as written (no FMA contraction).
#include "ngp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include <omp.h>

namespace {

// ---------------------------------------------------------------------------------------------
// binary16 helpers (round-to-nearest-even, like __float2half_rn)
// ---------------------------------------------------------------------------------------------
inline float h2f(orc_half h) { _Float16 v; std::memcpy(&v, &h, 2); return (float)v; }
inline orc_half f2h(float f) { _Float16 v = (_Float16)f; orc_half h; std::memcpy(&h, &v, 2); return h; }
// fp16 add as done by the reference's `result += (T)(weight * data)` (grid.h:341): both operands
// are halfs, the sum is rounded to half.
inline orc_half hadd(orc_half a, orc_half b) { return f2h(h2f(a) + h2f(b)); }

// ---------------------------------------------------------------------------------------------
// constants: src/testbed_nerf.cu:53-73, include/neural-graphics-primitives/nerf.h:24
// ---------------------------------------------------------------------------------------------
constexpr uint32_t NERF_GRIDSIZE = 128;
constexpr uint32_t NERF_STEPS = 1024;
constexpr uint32_t NERF_CASCADES = 8;
constexpr uint32_t N_MAX_RANDOM_SAMPLES_PER_RAY = 8;
constexpr float SQRT3 = 1.73205080757f;
constexpr float STEPSIZE = SQRT3 / NERF_STEPS;
constexpr float MIN_CONE_STEPSIZE = STEPSIZE;
constexpr float MAX_CONE_STEPSIZE = STEPSIZE * (1 << (NERF_CASCADES - 1)) * NERF_STEPS / NERF_GRIDSIZE;
constexpr float NERF_MIN_OPTICAL_THICKNESS = 0.01f;
constexpr uint32_t GRID_CELLS = NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE;

inline float clampf(float v, float lo, float hi) { return std::max(lo, std::min(v, hi)); } // tcnn common.h clamp
inline int clampi(int v, int lo, int hi) { return std::max(lo, std::min(v, hi)); }

// Eigen sums three terms as a + (b + c) (Eigen/src/Core/Redux.h:101-115, redux_novec_unroller).
inline float sum3(float a, float b, float c) { return a + (b + c); }
inline float dot3(const float* a, const float* b) { return sum3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
inline float sum4(float a, float b, float c, float d) { return (a + b) + (c + d); }

struct Vec3 { float x, y, z; };

// ---------------------------------------------------------------------------------------------
// morton: tcnn common_device.h:338-362
// ---------------------------------------------------------------------------------------------
inline uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
inline uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}

// ---------------------------------------------------------------------------------------------
// colour transfer: include/neural-graphics-primitives/common_device.cuh:31-77
// ---------------------------------------------------------------------------------------------
inline float srgb_to_linear(float srgb) {
	if (srgb <= 0.04045f) return srgb / 12.92f;
	return std::pow((srgb + 0.055f) / 1.055f, 2.4f);
}
inline float srgb_to_linear_derivative(float srgb) { // common_device.cuh:43-49
	if (srgb <= 0.04045f) return 1.0f / 12.92f;
	return 2.4f / 1.055f * std::pow((srgb + 0.055f) / 1.055f, 1.4f);
}
inline float linear_to_srgb(float linear) {
	if (linear < 0.0031308f) return 12.92f * linear;
	return 1.055f * std::pow(linear, 0.41666f) - 0.055f;
}
inline float logistic(float x) { return 1.0f / (1.0f + std::exp(-x)); } // tcnn common_device.h:51

// activations: src/testbed_nerf.cu:215-257 (the reference uses __expf; expf here, tolerance in the tests)
inline float network_to_rgb(float v, int act) {
	switch (act) {
		case 0: return v;
		case 1: return v > 0.0f ? v : 0.0f;
		case 2: return logistic(v);
		default: return std::exp(clampf(v, -10.0f, 10.0f));
	}
}
inline float network_to_rgb_derivative(float v, int act) {
	switch (act) {
		case 0: return 1.0f;
		case 1: return v > 0.0f ? 1.0f : 0.0f;
		case 2: { float d = logistic(v); return d * (1 - d); }
		default: return std::exp(clampf(v, -10.0f, 10.0f));
	}
}
inline float network_to_density(float v, int act) {
	switch (act) {
		case 0: return v;
		case 1: return v > 0.0f ? v : 0.0f;
		case 2: return logistic(v);
		default: return std::exp(v);
	}
}
inline float network_to_density_derivative(float v, int act) {
	switch (act) {
		case 0: return 1.0f;
		case 1: return v > 0.0f ? 1.0f : 0.0f;
		case 2: { float d = logistic(v); return d * (1 - d); }
		default: return std::exp(clampf(v, -15.0f, 15.0f));
	}
}

// ---------------------------------------------------------------------------------------------
// step-size law and occupancy lookup: src/testbed_nerf.cu:96-98,:191-213,:318-342,:449-463
// ---------------------------------------------------------------------------------------------
inline float calc_dt(float t, float cone_angle) { return clampf(t * cone_angle, MIN_CONE_STEPSIZE, MAX_CONE_STEPSIZE); }
inline float signf(float x) { return std::copysign(1.0f, x); } // common.h:197

inline float distance_to_next_voxel(const Vec3& pos, const Vec3& dir, const Vec3& idir, uint32_t res) {
	float r = (float)res;
	float px = r * pos.x, py = r * pos.y, pz = r * pos.z;
	float tx = (std::floor(px + 0.5f + 0.5f * signf(dir.x)) - px) * idir.x;
	float ty = (std::floor(py + 0.5f + 0.5f * signf(dir.y)) - py) * idir.y;
	float tz = (std::floor(pz + 0.5f + 0.5f * signf(dir.z)) - pz) * idir.z;
	float t = std::min(std::min(tx, ty), tz);
	return std::fmax(t / r, 0.0f);
}
inline float advance_to_next_voxel(float t, float cone_angle, const Vec3& pos, const Vec3& dir, const Vec3& idir, uint32_t res) {
	float t_target = t + distance_to_next_voxel(pos, dir, idir, res);
	do { t += calc_dt(t, cone_angle); } while (t < t_target);
	return t;
}
inline int mip_from_pos(const Vec3& pos, uint32_t max_cascade = NERF_CASCADES - 1) {
	int exponent;
	float maxval = std::max(std::max(std::fabs(pos.x - 0.5f), std::fabs(pos.y - 0.5f)), std::fabs(pos.z - 0.5f));
	std::frexp(maxval, &exponent);
	return std::min((int)max_cascade, std::max(0, exponent + 1));
}
inline int mip_from_dt(float dt, const Vec3& pos, uint32_t max_cascade = NERF_CASCADES - 1) {
	int mip = mip_from_pos(pos, max_cascade);
	dt *= 2 * NERF_GRIDSIZE;
	if (dt < 1.f) return mip;
	int exponent;
	std::frexp(dt, &exponent);
	return std::min((int)max_cascade, std::max(exponent, mip));
}
inline uint32_t cascaded_grid_idx_at(Vec3 pos, uint32_t mip) {
	float mip_scale = std::scalbn(1.0f, -(int)mip);
	pos.x -= 0.5f; pos.y -= 0.5f; pos.z -= 0.5f;
	pos.x *= mip_scale; pos.y *= mip_scale; pos.z *= mip_scale;
	pos.x += 0.5f; pos.y += 0.5f; pos.z += 0.5f;
	int ix = (int)(pos.x * NERF_GRIDSIZE), iy = (int)(pos.y * NERF_GRIDSIZE), iz = (int)(pos.z * NERF_GRIDSIZE);
	return morton3D(clampi(ix, 0, NERF_GRIDSIZE - 1), clampi(iy, 0, NERF_GRIDSIZE - 1), clampi(iz, 0, NERF_GRIDSIZE - 1));
}
inline uint32_t grid_mip_offset(uint32_t mip) { return GRID_CELLS * mip; }
inline bool density_grid_occupied_at(const Vec3& pos, const uint8_t* bitfield, uint32_t mip) {
	uint32_t idx = cascaded_grid_idx_at(pos, mip);
	return bitfield[idx / 8 + grid_mip_offset(mip) / 8] & (1 << (idx % 8));
}
inline float warp_dt(float dt) {
	float max_stepsize = MIN_CONE_STEPSIZE * (1 << (NERF_CASCADES - 1));
	return (dt - MIN_CONE_STEPSIZE) / (max_stepsize - MIN_CONE_STEPSIZE);
}
inline float unwarp_dt(float dt) {
	float max_stepsize = MIN_CONE_STEPSIZE * (1 << (NERF_CASCADES - 1));
	return dt * (max_stepsize - MIN_CONE_STEPSIZE) + MIN_CONE_STEPSIZE;
}

// bounding box: include/neural-graphics-primitives/bounding_box.cuh:163-221
struct AABB { Vec3 min, max; };
inline AABB make_aabb(const float* a) { return {{a[0], a[1], a[2]}, {a[3], a[4], a[5]}}; }
inline bool aabb_contains(const AABB& b, const Vec3& p) {
	return p.x >= b.min.x && p.x <= b.max.x && p.y >= b.min.y && p.y <= b.max.y && p.z >= b.min.z && p.z <= b.max.z;
}
inline void aabb_ray_intersect(const AABB& b, const Vec3& pos, const Vec3& dir, float* out_tmin, float* out_tmax) {
	const float FMAX = std::numeric_limits<float>::max();
	float tmin = (b.min.x - pos.x) / dir.x, tmax = (b.max.x - pos.x) / dir.x;
	if (tmin > tmax) std::swap(tmin, tmax);
	float tymin = (b.min.y - pos.y) / dir.y, tymax = (b.max.y - pos.y) / dir.y;
	if (tymin > tymax) std::swap(tymin, tymax);
	if (tmin > tymax || tymin > tmax) { *out_tmin = FMAX; *out_tmax = FMAX; return; }
	if (tymin > tmin) tmin = tymin;
	if (tymax < tmax) tmax = tymax;
	float tzmin = (b.min.z - pos.z) / dir.z, tzmax = (b.max.z - pos.z) / dir.z;
	if (tzmin > tzmax) std::swap(tzmin, tzmax);
	if (tmin > tzmax || tzmin > tmax) { *out_tmin = FMAX; *out_tmax = FMAX; return; }
	if (tzmin > tmin) tmin = tzmin;
	if (tzmax < tmax) tmax = tzmax;
	*out_tmin = tmin; *out_tmax = tmax;
}
inline Vec3 warp_position(const Vec3& p, const AABB& b) { // bounding_box.cuh:86 relative_pos
	return {(p.x - b.min.x) / (b.max.x - b.min.x), (p.y - b.min.y) / (b.max.y - b.min.y), (p.z - b.min.z) / (b.max.z - b.min.z)};
}
inline Vec3 unwarp_position(const float* p, const AABB& b) { // testbed_nerf.cu:274-279
	return {b.min.x + p[0] * (b.max.x - b.min.x), b.min.y + p[1] * (b.max.y - b.min.y), b.min.z + p[2] * (b.max.z - b.min.z)};
}

// ---------------------------------------------------------------------------------------------
// image access: common_device.cuh:633-705
// ---------------------------------------------------------------------------------------------
inline void image_pos(float x, float y, int w, int h, int* px, int* py) {
	*px = std::max(std::min((int)(x * (float)w), w - 1), 0);
	*py = std::max(std::min((int)(y * (float)h), h - 1), 0);
}
inline void read_rgba(float x, float y, const orc_image& img, float out[4]) {
	int px, py;
	image_pos(x, y, img.w, img.h, &px, &py);
	if (img.image_type == 2) { std::memcpy(out, img.pixels + 16 * ((size_t)px + (size_t)py * img.w), 16); return; } // Float: as stored (:702-703)
	if (img.image_type == 1) { // Half (:697-700)
		orc_half v[4];
		std::memcpy(v, img.pixels + 8 * ((size_t)px + (size_t)py * img.w), 8);
		for (int c = 0; c < 4; ++c) out[c] = h2f(v[c]);
		return;
	}
	uint8_t val[4];
	std::memcpy(val, img.pixels + 4 * ((size_t)px + (size_t)py * img.w), 4);
	uint32_t packed; std::memcpy(&packed, val, 4);
	if (packed == 0x00FF00FF) { out[0] = out[1] = out[2] = out[3] = -1.0f; return; }
	float alpha = (float)val[3] * (1.0f / 255.0f);
	out[0] = srgb_to_linear((float)val[0] * (1.0f / 255.0f)) * alpha;
	out[1] = srgb_to_linear((float)val[1] * (1.0f / 255.0f)) * alpha;
	out[2] = srgb_to_linear((float)val[2] * (1.0f / 255.0f)) * alpha;
	out[3] = alpha;
}

// ---- K19: importance sampling by accumulated error (Testbed::Nerf::Training::ErrorMap, testbed.h:600-615) ----
// The CDFs / error map the K1, K6 restatements below use; set through orc_set_error_cdf / orc_set_error_map (all null = uniform sampling, no accumulation,
// which is the reference's behaviour until its sample_*_proportional_to_error switches are on).
struct ErrorCdf { const float* x_cond_y; const float* y; const float* img; int res_x, res_y; };
static ErrorCdf g_cdf = {nullptr, nullptr, nullptr, 0, 0};
static float* g_error_map = nullptr;
static int g_error_map_res[2] = {0, 0};

inline float ld_random_val(uint32_t index, uint32_t seed, uint32_t dim); // (defined with the render restatement below)

// binary_search: include/neural-graphics-primitives/common.h:201-223
inline uint32_t binary_search(float val, const float* data, uint32_t length) {
	if (length == 0) return 0;
	uint32_t it, count = length, step, first = 0;
	while (count > 0) {
		it = first; step = count / 2; it += step;
		if (data[it] < val) { first = ++it; count -= step + 1; } else count = step;
	}
	return std::min(first, length - 1);
}

// image_idx: src/testbed_nerf.cu:1062-1083. uint32 wrap-around multiply as in the reference.
inline uint32_t image_idx(uint32_t base_idx, uint32_t n_rays, uint32_t n_training_images, float* pdf = nullptr) {
	if (g_cdf.img) {
		const float sample = ld_random_val(base_idx, 0xdeadbeefu, 0);
		const uint32_t img = binary_search(sample, g_cdf.img, n_training_images);
		if (pdf) { const float prev = img > 0 ? g_cdf.img[img - 1] : 0.0f; *pdf = (g_cdf.img[img] - prev) * n_training_images; }
		return img;
	}
	if (pdf) *pdf = 1.0f;
	return ((base_idx * n_training_images) / n_rays) % n_training_images;
}

// ---------------------------------------------------------------------------------------------
// loss functions: src/testbed_nerf.cu:121-189,:1263-1278
// ---------------------------------------------------------------------------------------------
struct LossAndGradient { float loss[3], gradient[3]; };
inline LossAndGradient loss_and_gradient(const float* target, const float* prediction, int loss_type) {
	LossAndGradient r;
	for (int c = 0; c < 3; ++c) {
		float difference = prediction[c] - target[c];
		switch (loss_type) {
			case 6: { // RelativeL2
				float factor = 1.0f / (prediction[c] * prediction[c] + 1e-2f);
				r.loss[c] = difference * difference * factor; r.gradient[c] = 2.0f * difference * factor; break;
			}
			case 1: r.loss[c] = std::fabs(difference); r.gradient[c] = std::copysign(1.0f, difference); break; // L1
			case 2: { // Mape
				float factor = 1.0f / (std::fabs(prediction[c]) + 1e-2f);
				r.loss[c] = std::fabs(difference) * factor; r.gradient[c] = std::copysign(factor, difference); break;
			}
			case 3: { // Smape
				float factor = 1.0f / (0.5f * (std::fabs(prediction[c]) + std::fabs(target[c])) + 1e-2f);
				r.loss[c] = std::fabs(difference) * factor; r.gradient[c] = std::copysign(factor, difference); break;
			}
			case 4: { // Huber(alpha=0.1)/5
				const float alpha = 0.1f;
				float abs_diff = std::fabs(difference);
				float square = 0.5f / alpha * difference * difference;
				float l = abs_diff > alpha ? (abs_diff - 0.5f * alpha) : square;
				float g = abs_diff > alpha ? (difference > 0 ? 1.0f : -1.0f) : (difference / alpha);
				r.loss[c] = l / 5.0f; r.gradient[c] = g / 5.0f; break;
			}
			case 5: { // LogL1
				float divisor = std::fabs(difference) + 1.0f;
				r.loss[c] = std::log(divisor); r.gradient[c] = std::copysign(1.0f / divisor, difference); break;
			}
			default: r.loss[c] = difference * difference; r.gradient[c] = 2.0f * difference; break; // L2
		}
	}
	return r;
}

inline orc_pcg32 rng_advanced(orc_pcg32 rng, int64_t delta) { orc_pcg32_advance(&rng, delta); return rng; }

// sample_cdf_2d: src/testbed_nerf.cu:991-1022 (UNIFORM_SAMPLING_FRACTION = 0.5: half of the samples stay uniform)
inline void sample_cdf_2d(float* sx, float* sy, uint32_t img, float* pdf) {
	const int rx = g_cdf.res_x, ry = g_cdf.res_y;
	if (*sx < 0.5f) { *sx /= 0.5f; return; }
	*sx = (*sx - 0.5f) / (1.0f - 0.5f);
	const float* cdf_y = g_cdf.y + (size_t)img * ry;
	const uint32_t y = binary_search(*sy, cdf_y, (uint32_t)ry);
	float prev = y > 0 ? cdf_y[y - 1] : 0.0f;
	const float pmf_y = cdf_y[y] - prev;
	*sy = (*sy - prev) / pmf_y;
	const float* cdf_x = g_cdf.x_cond_y + (size_t)img * ry * rx + (size_t)y * rx;
	const uint32_t x = binary_search(*sx, cdf_x, (uint32_t)rx);
	prev = x > 0 ? cdf_x[x - 1] : 0.0f;
	const float pmf_x = cdf_x[x] - prev;
	*sx = (*sx - prev) / pmf_x;
	if (pdf) *pdf = pmf_x * pmf_y * (float)(rx * ry);
	*sx = ((float)x + *sx) / (float)rx;
	*sy = ((float)y + *sy) / (float)ry;
}

// nerf_random_image_pos_training: src/testbed_nerf.cu:1047-1060
inline void random_image_pos_training(orc_pcg32& rng, int w, int h, bool snap, float* x, float* y, uint32_t img = 0, float* pdf = nullptr) {
	float u = orc_pcg32_next_float(&rng), v = orc_pcg32_next_float(&rng);
	if (pdf) *pdf = 1.0f; // (the uniform half of sample_cdf_2d leaves the caller's initial value, 1, :1378-1386)
	if (g_cdf.x_cond_y) sample_cdf_2d(&u, &v, img, pdf);
	if (snap) {
		u = ((float)std::min(std::max((int)(u * (float)w), 0), w - 1) + 0.5f) / (float)w;
		v = ((float)std::min(std::max((int)(v * (float)h), 0), h - 1) + 0.5f) / (float)h;
	}
	*x = u; *y = v;
}

} // namespace

// =============================================================================================
// PCG32
// =============================================================================================
extern "C" uint32_t orc_pcg32_next_uint(orc_pcg32* rng) {
	uint64_t oldstate = rng->state;
	rng->state = oldstate * 0x5851f42d4c957f2dULL + rng->inc;
	uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
	uint32_t rot = (uint32_t)(oldstate >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
extern "C" void orc_pcg32_seed(orc_pcg32* rng, uint64_t initstate, uint64_t initseq) {
	rng->state = 0U;
	rng->inc = (initseq << 1u) | 1u;
	orc_pcg32_next_uint(rng);
	rng->state += initstate;
	orc_pcg32_next_uint(rng);
}
extern "C" float orc_pcg32_next_float(orc_pcg32* rng) {
	union { uint32_t u; float f; } x;
	x.u = (orc_pcg32_next_uint(rng) >> 9) | 0x3f800000u;
	return x.f - 1.0f;
}
extern "C" void orc_pcg32_advance(orc_pcg32* rng, int64_t delta_) {
	uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = rng->inc, acc_mult = 1u, acc_plus = 0u;
	uint64_t delta = (uint64_t)delta_;
	while (delta > 0) {
		if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
		cur_plus = (cur_mult + 1) * cur_plus;
		cur_mult *= cur_mult;
		delta /= 2;
	}
	rng->state = acc_mult * rng->state + acc_plus;
}

// =============================================================================================
// Hash grid
// =============================================================================================
namespace {
inline float grid_scale(uint32_t level, float log2_per_level_scale, uint32_t base_resolution) { // grid.h:194-199
	return std::exp2(level * log2_per_level_scale) * base_resolution - 1.0f;
}
inline uint32_t grid_resolution(float scale) { return (uint32_t)std::ceil(scale) + 1; } // grid.h:201-203

// grid.h:164-186 with GridType::Hash, HashType::CoherentPrime, 3 dims; returns the entry index (feature 0 = index*2).
inline uint32_t grid_index(uint32_t hashmap_size, uint32_t resolution, const uint32_t p[3]) {
	uint32_t stride = 1, index = 0;
	for (uint32_t dim = 0; dim < 3 && stride <= hashmap_size; ++dim) {
		index += p[dim] * stride;
		stride *= resolution;
	}
	if (hashmap_size < stride) {
		index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u); // grid.h:111-128
	}
	return index % hashmap_size;
}
inline void pos_fract(float input, float* pos, uint32_t* pos_grid, float scale) { // tcnn common_device.h:434-445
	*pos = std::fma(input, scale, 0.5f); // nvcc contracts the reference's `input * scale + 0.5f` into one FFMA
	int tmp = (int)std::floor(*pos);
	*pos_grid = (uint32_t)tmp;
	*pos -= (float)tmp;
}
} // namespace

extern "C" uint32_t orc_grid_offsets(uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, uint32_t* offsets) {
	uint32_t offset = 0;
	for (uint32_t i = 0; i < n_levels; ++i) {
		const uint32_t resolution = grid_resolution(grid_scale(i, std::log2(per_level_scale), base_resolution));
		uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
		uint32_t params_in_level = std::pow((float)resolution, 3) > (float)max_params ? max_params : resolution * resolution * resolution;
		params_in_level = (params_in_level + 7u) / 8u * 8u;
		params_in_level = std::min(params_in_level, 1u << log2_hashmap_size);
		offsets[i] = offset;
		offset += params_in_level;
	}
	offsets[n_levels] = offset;
	return offset;
}

extern "C" void orc_grid_indices(uint32_t n, uint32_t level, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                                 const float* positions, uint32_t pos_stride, uint32_t* indices8, float* weights8) {
	const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
	const float scale = scales ? scales[level] : grid_scale(level, log2_per_level_scale, base_resolution);
	const uint32_t resolution = grid_resolution(scale);
	for (uint32_t i = 0; i < n; ++i) {
		float pos[3]; uint32_t pg[3];
		for (int d = 0; d < 3; ++d) pos_fract(positions[(size_t)i * pos_stride + d], &pos[d], &pg[d], scale);
		for (uint32_t idx = 0; idx < 8; ++idx) {
			float weight = 1; uint32_t pl[3];
			for (uint32_t d = 0; d < 3; ++d) {
				if ((idx & (1 << d)) == 0) { weight *= 1 - pos[d]; pl[d] = pg[d]; } else { weight *= pos[d]; pl[d] = pg[d] + 1; }
			}
			indices8[(size_t)i * 8 + idx] = grid_index(hashmap_size, resolution, pl);
			weights8[(size_t)i * 8 + idx] = weight;
		}
	}
}

extern "C" void orc_grid_forward(uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                                 const orc_half* grid, const float* positions, uint32_t pos_stride, orc_half* encoded) {
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		for (uint32_t level = 0; level < n_levels; ++level) {
			const orc_half* g = grid + (size_t)offsets[level] * 2;
			const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
			const float scale = scales ? scales[level] : grid_scale(level, log2_per_level_scale, base_resolution);
			const uint32_t resolution = grid_resolution(scale);
			float pos[3]; uint32_t pg[3];
			for (int d = 0; d < 3; ++d) pos_fract(positions[(size_t)i * pos_stride + d], &pos[d], &pg[d], scale);
			orc_half result[2] = {0, 0};
			for (uint32_t idx = 0; idx < 8; ++idx) {
				float weight = 1; uint32_t pl[3];
				for (uint32_t d = 0; d < 3; ++d) {
					if ((idx & (1 << d)) == 0) { weight *= 1 - pos[d]; pl[d] = pg[d]; } else { weight *= pos[d]; pl[d] = pg[d] + 1; }
				}
				uint32_t index = grid_index(hashmap_size, resolution, pl) * 2;
				for (int f = 0; f < 2; ++f) {
					float data = h2f(g[index + f]);
					result[f] = hadd(result[f], f2h(weight * data)); // grid.h:339-341 (fp16 accumulation)
				}
			}
			encoded[(size_t)i * (2 * n_levels) + level * 2 + 0] = result[0];
			encoded[(size_t)i * (2 * n_levels) + level * 2 + 1] = result[1];
		}
	}
}

// The reference accumulates with atomicAdd(__half2) in launch order (grid.h:436-441), which is
// order dependent; the oracle accumulates each contribution `(float)grad * weight` exactly in
// double and rounds once to float. Tests compare with a tolerance that covers fp16 atomics.
extern "C" void orc_grid_backward(uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                                  const float* positions, uint32_t pos_stride, const orc_half* dL_dy, float* grad) {
	#pragma omp parallel for schedule(dynamic, 1)
	for (int level = 0; level < (int)n_levels; ++level) {
		const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
		std::vector<double> acc((size_t)hashmap_size * 2, 0.0);
		const float scale = scales ? scales[level] : grid_scale(level, log2_per_level_scale, base_resolution);
		const uint32_t resolution = grid_resolution(scale);
		for (uint32_t i = 0; i < n; ++i) {
			float pos[3]; uint32_t pg[3];
			for (int d = 0; d < 3; ++d) pos_fract(positions[(size_t)i * pos_stride + d], &pos[d], &pg[d], scale);
			float g0 = h2f(dL_dy[(size_t)i * (2 * n_levels) + level * 2 + 0]);
			float g1 = h2f(dL_dy[(size_t)i * (2 * n_levels) + level * 2 + 1]);
			if (g0 == 0.0f && g1 == 0.0f) continue;
			for (uint32_t idx = 0; idx < 8; ++idx) {
				float weight = 1; uint32_t pl[3];
				for (uint32_t d = 0; d < 3; ++d) {
					if ((idx & (1 << d)) == 0) { weight *= 1 - pos[d]; pl[d] = pg[d]; } else { weight *= pos[d]; pl[d] = pg[d] + 1; }
				}
				uint32_t index = grid_index(hashmap_size, resolution, pl) * 2;
				acc[index + 0] += (double)(g0 * weight);
				acc[index + 1] += (double)(g1 * weight);
			}
		}
		float* out = grad + (size_t)offsets[level] * 2;
		for (size_t k = 0; k < acc.size(); ++k) out[k] = (float)acc[k];
	}
}

// =============================================================================================
// Spherical harmonics degree 4 (16 coefficients): spherical_harmonics.h:46-150
// =============================================================================================
namespace {
inline void sh4(const float* d, float* out) {
	float x = d[0] * 2.f - 1.f, y = d[1] * 2.f - 1.f, z = d[2] * 2.f - 1.f;
	float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
	out[0] = 0.28209479177387814f;
	out[1] = -0.48860251190291987f * y;
	out[2] = 0.48860251190291987f * z;
	out[3] = -0.48860251190291987f * x;
	out[4] = 1.0925484305920792f * xy;
	out[5] = -1.0925484305920792f * yz;
	out[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
	out[7] = -1.0925484305920792f * xz;
	out[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
	out[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
	out[10] = 2.8906114426405538f * xy * z;
	out[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
	out[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
	out[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
	out[14] = 1.4453057213202769f * z * (x2 - y2);
	out[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}
} // namespace

extern "C" void orc_sh4(uint32_t n, const float* dirs, uint32_t dir_stride, orc_half* out, uint32_t out_stride) {
	for (uint32_t i = 0; i < n; ++i) {
		float v[16];
		sh4(dirs + (size_t)i * dir_stride, v);
		for (int k = 0; k < 16; ++k) out[(size_t)i * out_stride + k] = f2h(v[k]);
	}
}

// =============================================================================================
// NeRF network: density MLP 32->64->16, rgb MLP 32->64->64->16(3), ReLU hidden, no output act.
// fp32 accumulation, fp16 storage between layers. (The reference's wmma path accumulates in fp16,
// tcnn/src/fully_fused_mlp.cu:66-68; the difference is covered by the stated tolerance.)
// =============================================================================================
namespace {
constexpr int W1D = 0, W2D = 2048, W1R = 3072, W2R = 5120, W3R = 9216;

struct MlpWeightsF {
	std::vector<float> w;
	explicit MlpWeightsF(const orc_half* mlp) : w(ORC_MLP_PARAMS) { for (int i = 0; i < ORC_MLP_PARAMS; ++i) w[i] = h2f(mlp[i]); }
};

// out[o] = sum_k W[o][k] * in[k]
inline void matvec(const float* W, int n_out, int n_in, const float* in, float* out) {
	for (int o = 0; o < n_out; ++o) {
		float acc = 0.f;
		const float* row = W + o * n_in;
		for (int k = 0; k < n_in; ++k) acc += row[k] * in[k];
		out[o] = acc;
	}
}
// out[k] = sum_o W[o][k] * in[o]
inline void matvec_t(const float* W, int n_out, int n_in, const float* in, float* out) {
	for (int k = 0; k < n_in; ++k) out[k] = 0.f;
	for (int o = 0; o < n_out; ++o) {
		const float* row = W + o * n_in;
		float v = in[o];
		for (int k = 0; k < n_in; ++k) out[k] += row[k] * v;
	}
}
inline float rh(float v) { return h2f(f2h(v)); } // round through fp16

struct SampleActs { float x[32], h1[64], od[16], rin[32], g1[64], g2[64], orgb[16]; };

inline void mlp_forward_one(const float* W, const orc_half* enc, const float* coord, SampleActs& a) {
	float tmp[64];
	for (int k = 0; k < 32; ++k) a.x[k] = h2f(enc[k]);
	matvec(W + W1D, 64, 32, a.x, tmp);
	for (int k = 0; k < 64; ++k) a.h1[k] = rh(tmp[k] > 0.f ? tmp[k] : 0.f);
	matvec(W + W2D, 16, 64, a.h1, tmp);
	for (int k = 0; k < 16; ++k) a.od[k] = rh(tmp[k]);
	float sh[16];
	sh4(coord + 4, sh);
	for (int k = 0; k < 16; ++k) { a.rin[k] = a.od[k]; a.rin[16 + k] = rh(sh[k]); }
	matvec(W + W1R, 64, 32, a.rin, tmp);
	for (int k = 0; k < 64; ++k) a.g1[k] = rh(tmp[k] > 0.f ? tmp[k] : 0.f);
	matvec(W + W2R, 64, 64, a.g1, tmp);
	for (int k = 0; k < 64; ++k) a.g2[k] = rh(tmp[k] > 0.f ? tmp[k] : 0.f);
	matvec(W + W3R, 16, 64, a.g2, tmp);
	for (int k = 0; k < 16; ++k) a.orgb[k] = rh(tmp[k]);
}
} // namespace

extern "C" void orc_nerf_mlp_forward(uint32_t n, const orc_half* mlp, const orc_half* encoded, const float* coords,
                                     orc_half* rgbsigma, orc_half* act_h1, orc_half* rgb_in, orc_half* act_g1, orc_half* act_g2) {
	MlpWeightsF Wf(mlp);
	const float* W = Wf.w.data();
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		SampleActs a;
		mlp_forward_one(W, encoded + (size_t)i * 32, coords + (size_t)i * 7, a);
		// nerf_network.h:130-136: rgb from the rgb net, density = density-net output 0
		rgbsigma[i * 4 + 0] = f2h(a.orgb[0]); rgbsigma[i * 4 + 1] = f2h(a.orgb[1]); rgbsigma[i * 4 + 2] = f2h(a.orgb[2]);
		rgbsigma[i * 4 + 3] = f2h(a.od[0]);
		if (act_h1) for (int k = 0; k < 64; ++k) act_h1[(size_t)i * 64 + k] = f2h(a.h1[k]);
		if (rgb_in) for (int k = 0; k < 32; ++k) rgb_in[(size_t)i * 32 + k] = f2h(a.rin[k]);
		if (act_g1) for (int k = 0; k < 64; ++k) act_g1[(size_t)i * 64 + k] = f2h(a.g1[k]);
		if (act_g2) for (int k = 0; k < 64; ++k) act_g2[(size_t)i * 64 + k] = f2h(a.g2[k]);
	}
}

extern "C" void orc_nerf_mlp_backward(uint32_t n, const orc_half* mlp, const orc_half* encoded, const float* coords, const orc_half* dL_dout,
                                      orc_half* dL_dencoded, float* mlp_grad) {
	MlpWeightsF Wf(mlp);
	const float* W = Wf.w.data();
	const int nt = omp_get_max_threads();
	std::vector<std::vector<double>> partial(nt, std::vector<double>(ORC_MLP_PARAMS, 0.0));
	#pragma omp parallel
	{
		double* G = partial[omp_get_thread_num()].data();
		#pragma omp for schedule(static)
		for (int64_t i = 0; i < (int64_t)n; ++i) {
			SampleActs a;
			mlp_forward_one(W, encoded + (size_t)i * 32, coords + (size_t)i * 7, a);
			// nerf_network.h:202-206: dL_drgb = first three outputs, rest zero
			float d_orgb[16] = {0};
			for (int c = 0; c < 3; ++c) d_orgb[c] = h2f(dL_dout[(size_t)i * 4 + c]);
			float tmp[64], d_g2[64], d_g1[64], d_rin[32], d_od[16], d_h1[64], d_x[32];
			// rgb net, last layer (fully_fused_mlp.cu:759-850)
			for (int o = 0; o < 16; ++o) for (int k = 0; k < 64; ++k) G[W3R + o * 64 + k] += (double)(d_orgb[o] * a.g2[k]);
			matvec_t(W + W3R, 16, 64, d_orgb, tmp);
			for (int k = 0; k < 64; ++k) d_g2[k] = rh(a.g2[k] > 0.f ? tmp[k] : 0.f);
			for (int o = 0; o < 64; ++o) for (int k = 0; k < 64; ++k) G[W2R + o * 64 + k] += (double)(d_g2[o] * a.g1[k]);
			matvec_t(W + W2R, 64, 64, d_g2, tmp);
			for (int k = 0; k < 64; ++k) d_g1[k] = rh(a.g1[k] > 0.f ? tmp[k] : 0.f);
			for (int o = 0; o < 64; ++o) for (int k = 0; k < 32; ++k) G[W1R + o * 32 + k] += (double)(d_g1[o] * a.rin[k]);
			matvec_t(W + W1R, 64, 32, d_g1, tmp);
			for (int k = 0; k < 32; ++k) d_rin[k] = rh(tmp[k]);
			// nerf_network.h:232-239: density-net output gradient = rgb-net input gradient rows 0..15, plus dL/dsigma on row 0
			for (int k = 0; k < 16; ++k) d_od[k] = d_rin[k];
			d_od[0] = rh(d_od[0] + h2f(dL_dout[(size_t)i * 4 + 3]));
			for (int o = 0; o < 16; ++o) for (int k = 0; k < 64; ++k) G[W2D + o * 64 + k] += (double)(d_od[o] * a.h1[k]);
			matvec_t(W + W2D, 16, 64, d_od, tmp);
			for (int k = 0; k < 64; ++k) d_h1[k] = rh(a.h1[k] > 0.f ? tmp[k] : 0.f);
			for (int o = 0; o < 64; ++o) for (int k = 0; k < 32; ++k) G[W1D + o * 32 + k] += (double)(d_h1[o] * a.x[k]);
			matvec_t(W + W1D, 64, 32, d_h1, d_x);
			for (int k = 0; k < 32; ++k) dL_dencoded[(size_t)i * 32 + k] = f2h(d_x[k]);
		}
	}
	for (int k = 0; k < ORC_MLP_PARAMS; ++k) {
		double s = 0.0;
		for (int t = 0; t < nt; ++t) s += partial[t][k];
		mlp_grad[k] = (float)s;
	}
}

// =============================================================================================
// composite model entry points
// =============================================================================================
extern "C" void orc_model_init(orc_model* m, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale) {
	m->n_levels = n_levels; m->log2_hashmap_size = log2_hashmap_size; m->base_resolution = base_resolution; m->per_level_scale = per_level_scale;
	m->n_grid_params = 2 * orc_grid_offsets(n_levels, log2_hashmap_size, base_resolution, per_level_scale, m->offsets);
	for (uint32_t l = 0; l < n_levels; ++l) m->scales[l] = grid_scale(l, std::log2(per_level_scale), base_resolution);
}

extern "C" void orc_nerf_inference(const orc_model* m, const orc_half* params, uint32_t n, const float* coords, orc_half* rgbsigma) {
	std::vector<orc_half> enc((size_t)n * 32);
	orc_grid_forward(n, m->n_levels, m->offsets, m->base_resolution, std::log2(m->per_level_scale), m->scales, params + ORC_MLP_PARAMS, coords, 7, enc.data());
	orc_nerf_mlp_forward(n, params, enc.data(), coords, rgbsigma, nullptr, nullptr, nullptr, nullptr);
}

extern "C" void orc_nerf_density(const orc_model* m, const orc_half* params, uint32_t n, const float* positions, uint32_t pos_stride, orc_half* density) {
	std::vector<orc_half> enc((size_t)n * 32);
	orc_grid_forward(n, m->n_levels, m->offsets, m->base_resolution, std::log2(m->per_level_scale), m->scales, params + ORC_MLP_PARAMS, positions, pos_stride, enc.data());
	MlpWeightsF Wf(params);
	const float* W = Wf.w.data();
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		float x[32], h1[64], tmp[64];
		for (int k = 0; k < 32; ++k) x[k] = h2f(enc[(size_t)i * 32 + k]);
		matvec(W + W1D, 64, 32, x, tmp);
		for (int k = 0; k < 64; ++k) h1[k] = rh(tmp[k] > 0.f ? tmp[k] : 0.f);
		matvec(W + W2D, 1, 64, h1, tmp);
		density[i] = f2h(tmp[0]);
	}
}

extern "C" void orc_nerf_forward_backward(const orc_model* m, const orc_half* params, uint32_t n, const float* coords, const orc_half* dL_dout, float* grad) {
	std::vector<orc_half> enc((size_t)n * 32), denc((size_t)n * 32);
	const float l2 = std::log2(m->per_level_scale);
	orc_grid_forward(n, m->n_levels, m->offsets, m->base_resolution, l2, m->scales, params + ORC_MLP_PARAMS, coords, 7, enc.data());
	orc_nerf_mlp_backward(n, params, enc.data(), coords, dL_dout, denc.data(), grad);
	orc_grid_backward(n, m->n_levels, m->offsets, m->base_resolution, l2, m->scales, coords, 7, denc.data(), grad + ORC_MLP_PARAMS);
}

// =============================================================================================
// camera transform the kernels actually use (common_device.cuh:224-234 with zero rolling shutter:
// matrix -> quaternion -> slerp(t=0) -> normalize -> matrix). Eigen/src/Geometry/Quaternion.h.
// =============================================================================================
extern "C" void orc_effective_xform(const float* m12, float* out12) {
	auto M = [&](int r, int c) { return m12[c * 3 + r]; };
	float q[4]; // x y z w
	float t = sum3(M(0, 0), M(1, 1), M(2, 2));
	if (t > 0.f) {
		t = std::sqrt(t + 1.0f);
		q[3] = 0.5f * t;
		t = 0.5f / t;
		q[0] = (M(2, 1) - M(1, 2)) * t;
		q[1] = (M(0, 2) - M(2, 0)) * t;
		q[2] = (M(1, 0) - M(0, 1)) * t;
	} else {
		int i = 0;
		if (M(1, 1) > M(0, 0)) i = 1;
		if (M(2, 2) > M(i, i)) i = 2;
		int j = (i + 1) % 3, k = (j + 1) % 3;
		t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0f);
		q[i] = 0.5f * t;
		t = 0.5f / t;
		q[3] = (M(k, j) - M(j, k)) * t;
		q[j] = (M(j, i) + M(i, j)) * t;
		q[k] = (M(k, i) + M(i, k)) * t;
	}
	// slerp(0, same quaternion): scale0 = 1, scale1 = 0 in both branches -> coefficients unchanged up to +0*q.
	for (int c = 0; c < 4; ++c) q[c] = 1.0f * q[c] + 0.0f * q[c];
	float z = sum4(q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]);
	if (z > 0.f) { float nrm = std::sqrt(z); for (int c = 0; c < 4; ++c) q[c] = q[c] / nrm; }
	const float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
	const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
	const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
	const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
	auto R = [&](int r, int c) -> float& { return out12[c * 3 + r]; };
	R(0, 0) = 1.f - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
	R(1, 0) = txy + twz; R(1, 1) = 1.f - (txx + tzz); R(1, 2) = tyz - twx;
	R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1.f - (txx + tyy);
	out12[9] = m12[9] + (m12[9] - m12[9]) * 0.f; out12[10] = m12[10] + (m12[10] - m12[10]) * 0.f; out12[11] = m12[11] + (m12[11] - m12[11]) * 0.f;
}

// =============================================================================================
// K1: generate_training_samples_nerf, src/testbed_nerf.cu:1085-1260
// =============================================================================================
namespace {
struct TrainRay { Vec3 o, d_unnorm, d; float startt; float cone_angle; bool valid; };

// ---- lens models (include/neural-graphics-primitives/common_device.cuh:141-200, :236-258) ----
inline void opencv_lens_distortion(const float* prm, float u, float v, float* du, float* dv) { // apply_opencv_lens_distortion
	const float k1 = prm[0], k2 = prm[1], p1 = prm[2], p2 = prm[3];
	const float u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
	const float radial = k1 * r2 + k2 * r2 * r2;
	*du = u * radial + 2.0f * p1 * uv + p2 * (r2 + 2.0f * u2);
	*dv = v * radial + 2.0f * p2 * uv + p1 * (r2 + 2.0f * v2);
}
inline void opencv_lens_undistortion(const float* prm, float* u, float* v) { // iterative_opencv_lens_undistortion: Newton, central differences, <= 100 steps
	const float x00 = *u, x01 = *v;
	float x0 = *u, x1 = *v;
	for (uint32_t i = 0; i < 100; ++i) {
		const float step0 = std::max(std::numeric_limits<float>::epsilon(), std::abs(1e-6f * x0)), step1 = std::max(std::numeric_limits<float>::epsilon(), std::abs(1e-6f * x1));
		float dx0, dx1, b00, b01, f00, f01, b10, b11, f10, f11;
		opencv_lens_distortion(prm, x0, x1, &dx0, &dx1);
		opencv_lens_distortion(prm, x0 - step0, x1, &b00, &b01);
		opencv_lens_distortion(prm, x0 + step0, x1, &f00, &f01);
		opencv_lens_distortion(prm, x0, x1 - step1, &b10, &b11);
		opencv_lens_distortion(prm, x0, x1 + step1, &f10, &f11);
		const float j00 = 1 + (f00 - b00) / (2 * step0), j01 = (f10 - b10) / (2 * step1), j10 = (f01 - b01) / (2 * step0), j11 = 1 + (f11 - b11) / (2 * step1);
		const float invdet = 1.0f / (j00 * j11 - j10 * j01); // Eigen's 2x2 inverse: adjugate / determinant
		const float i00 = j11 * invdet, i10 = -j10 * invdet, i01 = -j01 * invdet, i11 = j00 * invdet;
		const float r0 = (x0 + dx0) - x00, r1 = (x1 + dx1) - x01;
		const float s0 = i00 * r0 + i01 * r1, s1 = i10 * r0 + i11 * r1;
		x0 -= s0; x1 -= s1;
		if (s0 * s0 + s1 * s1 < 1e-10f) break;
	}
	*u = x0; *v = x1;
}
inline Vec3 f_theta_undistortion(float u, float v, const float* prm, const Vec3& error_direction) {
	const float xpix = u * prm[5], ypix = v * prm[6];
	const float norm = std::sqrt(xpix * xpix + ypix * ypix);
	const float alpha = prm[0] + norm * (prm[1] + norm * (prm[2] + norm * (prm[3] + norm * prm[4])));
	float sin_alpha = std::sin(alpha), cos_alpha = std::cos(alpha);
	if (cos_alpha <= std::numeric_limits<float>::min() || norm == 0.f) return error_direction;
	sin_alpha *= 1.f / norm;
	return {sin_alpha * xpix, sin_alpha * ypix, cos_alpha};
}
inline Vec3 latlong_to_dir(float u, float v) {
	const float PI = 3.14159265358979323846f;
	const float theta = (v - 0.5f) * PI, phi = (u - 0.5f) * PI * 2.0f;
	const float st = std::sin(theta), ct = std::cos(theta), sp = std::sin(phi), cp = std::cos(phi);
	return {sp * ct, st, cp * ct};
}
// camera-space direction of pixel (x, y) for the image's lens (testbed_nerf.cu:1166-1184), not normalised
inline Vec3 training_ray_direction(const orc_image& im, float x, float y) {
	if (im.lens_mode == 2) return f_theta_undistortion(x - im.cx, y - im.cy, im.lens_params, Vec3{0.f, 0.f, 1.f});
	if (im.lens_mode == 3) return latlong_to_dir(x, y);
	Vec3 d = {(x - im.cx) * (float)im.w / im.fx, (y - im.cy) * (float)im.h / im.fy, 1.0f};
	if (im.lens_mode == 1) opencv_lens_undistortion(im.lens_params, &d.x, &d.y);
	return d;
}

// ray set-up shared by K1 and nothing else (K6 only re-derives the pixel). testbed_nerf.cu:1118-1202
inline TrainRay setup_training_ray(uint32_t i, uint32_t n_rays, orc_pcg32 rng, uint32_t n_images, const orc_image* images,
                                   const float* eff_xforms, const AABB& aabb, bool snap, float cone_angle_constant) {
	TrainRay r{};
	uint32_t img = image_idx(i, n_rays, n_images);
	const orc_image& im = images[img];
	orc_pcg32_advance(&rng, (int64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
	float x, y;
	random_image_pos_training(rng, im.w, im.h, snap, &x, &y, img);
	float px[4];
	read_rgba(x, y, im, px);
	if (px[0] < 0.0f) { r.valid = false; return r; }
	/* max_level */ // max_level_rand_training is off: no draw (testbed_nerf.cu:1130)
	float motionblur_time = orc_pcg32_next_float(&rng); (void)motionblur_time;
	const float* xf = eff_xforms + (size_t)img * 12;
	r.o = {xf[9], xf[10], xf[11]};
	const Vec3 dc = training_ray_direction(im, x, y);
	float dcam[3] = {dc.x, dc.y, dc.z};
	// xform.block<3,3>(0,0) * d: row r = sum3 over columns (column-major storage)
	float row0[3] = {xf[0], xf[3], xf[6]}, row1[3] = {xf[1], xf[4], xf[7]}, row2[3] = {xf[2], xf[5], xf[8]};
	r.d_unnorm = {dot3(row0, dcam), dot3(row1, dcam), dot3(row2, dcam)};
	float z = sum3(r.d_unnorm.x * r.d_unnorm.x, r.d_unnorm.y * r.d_unnorm.y, r.d_unnorm.z * r.d_unnorm.z);
	if (z > 0.f) { float nrm = std::sqrt(z); r.d = {r.d_unnorm.x / nrm, r.d_unnorm.y / nrm, r.d_unnorm.z / nrm}; } else r.d = r.d_unnorm;
	float tmin, tmax;
	aabb_ray_intersect(aabb, r.o, r.d, &tmin, &tmax);
	r.cone_angle = cone_angle_constant;
	tmin = std::fmax(tmin, 0.0f);
	float startt = tmin;
	startt += calc_dt(startt, r.cone_angle) * orc_pcg32_next_float(&rng);
	r.startt = startt;
	r.valid = true;
	return r;
}

template <typename F>
inline uint32_t march_training_ray(const TrainRay& r, const AABB& aabb, const uint8_t* bitfield, uint32_t max_steps, F&& on_sample) {
	Vec3 idir = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
	uint32_t j = 0;
	float t = r.startt;
	Vec3 pos;
	while (aabb_contains(aabb, pos = Vec3{r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z}) && j < max_steps) {
		float dt = calc_dt(t, r.cone_angle);
		uint32_t mip = mip_from_dt(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) {
			on_sample(j, pos, dt);
			++j;
			t += dt;
		} else {
			uint32_t res = NERF_GRIDSIZE >> mip;
			t = advance_to_next_voxel(t, r.cone_angle, pos, r.d, idir, res);
		}
	}
	return j;
}
} // namespace

extern "C" uint32_t orc_generate_training_samples(
	uint32_t n_rays, const float* aabb6, uint32_t max_samples, uint32_t n_rays_total, orc_pcg32 rng,
	uint32_t n_images, const orc_image* images, const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t* counters_out) {
	(void)n_rays_total;
	return orc_generate_training_samples_sharded(n_rays, 0, n_rays, aabb6, max_samples, rng, n_images, images, bitfield, snap_to_pixel_centers, cone_angle_constant,
		ray_indices, rays, numsteps, coords, counters_out);
}

// The rays [ray_offset, ray_offset + n_rays) of a batch of n_rays_global rays: pixel, image and RNG stream of a ray derive from its global
// index (:1062-1083, :1118-1121), which is what makes ray-sharded data parallelism reproduce the unsharded batch.
extern "C" uint32_t orc_generate_training_samples_sharded(
	uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, orc_pcg32 rng,
	uint32_t n_images, const orc_image* images, const uint8_t* bitfield, int snap_to_pixel_centers, float cone_angle_constant,
	uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t* counters_out) {
	const AABB aabb = make_aabb(aabb6);
	std::vector<float> eff((size_t)n_images * 12);
	for (uint32_t k = 0; k < n_images; ++k) orc_effective_xform(images[k].xform, eff.data() + (size_t)k * 12);

	// pass 1 (parallel): per-ray set-up and step count (testbed_nerf.cu:1204-1219)
	std::vector<TrainRay> tr(n_rays);
	std::vector<uint32_t> counts(n_rays, 0);
	#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t i = 0; i < (int64_t)n_rays; ++i) {
		tr[i] = setup_training_ray(ray_offset + (uint32_t)i, n_rays_global, rng, n_images, images, eff.data(), aabb, snap_to_pixel_centers != 0, cone_angle_constant);
		if (tr[i].valid) counts[i] = march_training_ray(tr[i], aabb, bitfield, NERF_STEPS, [](uint32_t, const Vec3&, float) {});
	}
	// Allocation order: one valid serialisation of the reference's atomicAdd (:1225,:1232). Ray order while every sample fits max_samples. When the demand
	// exceeds it, the rays served last are dropped (:1226) -- in ray order always those of the LAST training images (the ray index selects the image),
	// whereas the reference's atomics drop whichever rays its blocks process last. So the order then starts at a ray drawn from the step's RNG (first draw
	// past the rays' own sub-streams; K6 takes the second) and wraps around.
	uint32_t first = 0;
	{
		uint64_t total = 0;
		for (uint32_t i = 0; i < n_rays; ++i) total += counts[i];
		if (total > max_samples && n_rays > 0) {
			orc_pcg32 r = rng_advanced(rng, (int64_t)n_rays_global * N_MAX_RANDOM_SAMPLES_PER_RAY);
			first = orc_pcg32_next_uint(&r) % n_rays;
		}
	}
	uint32_t numsteps_counter = 0, ray_counter = 0;
	std::vector<uint32_t> base_of(n_rays, 0xFFFFFFFFu), slot_of(n_rays, 0);
	for (uint32_t k = 0; k < n_rays; ++k) {
		const uint32_t i = first + k < n_rays ? first + k : first + k - n_rays;
		if (!tr[i].valid || counts[i] == 0) continue;
		uint32_t base = numsteps_counter;
		numsteps_counter += counts[i];
		if (base + counts[i] > max_samples) continue;
		base_of[i] = base; slot_of[i] = ray_counter++;
	}
	// pass 2 (parallel): write samples (testbed_nerf.cu:1239-1253)
	#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t i = 0; i < (int64_t)n_rays; ++i) {
		if (base_of[i] == 0xFFFFFFFFu) continue;
		const TrainRay& r = tr[i];
		uint32_t slot = slot_of[i], base = base_of[i];
		ray_indices[slot] = ray_offset + (uint32_t)i;
		float* ro = rays + (size_t)slot * 6;
		ro[0] = r.o.x; ro[1] = r.o.y; ro[2] = r.o.z; ro[3] = r.d_unnorm.x; ro[4] = r.d_unnorm.y; ro[5] = r.d_unnorm.z;
		numsteps[slot * 2 + 0] = counts[i];
		numsteps[slot * 2 + 1] = base;
		Vec3 wd = {(r.d.x + 1.0f) * 0.5f, (r.d.y + 1.0f) * 0.5f, (r.d.z + 1.0f) * 0.5f}; // warp_direction :292
		float* co = coords + (size_t)base * 7;
		march_training_ray(r, aabb, bitfield, counts[i], [&](uint32_t j, const Vec3& pos, float dt) {
			Vec3 wp = warp_position(pos, aabb);
			float* c = co + (size_t)j * 7;
			c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
		});
	}
	if (counters_out) { counters_out[0] = numsteps_counter; counters_out[1] = ray_counter; }
	return ray_counter;
}

// =============================================================================================
// K6: compute_loss_kernel_train_nerf, src/testbed_nerf.cu:1280-1597
// =============================================================================================
// `exposure` (3 floats per image, null = all zero) scales the target colour by 2^exposure (:1403); `exposure_gradient` (3 floats per image, null = off)
// accumulates d loss / d exposure per image (:1558-1571).
extern "C" uint32_t orc_compute_loss_exposure(
	uint32_t n_rays_kept, uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, orc_pcg32 rng, uint32_t max_samples_compacted,
	float loss_scale_in, const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, const orc_image* images, const orc_half* rgbsigma, const uint32_t* ray_indices, const float* rays,
	uint32_t* numsteps_io, const float* coords_in_all, float* coords_out_all, orc_half* dloss_dout_all, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, float mean_density, float near_distance,
	const float* exposure, float* exposure_gradient) {
	(void)n_rays_total;
	const AABB aabb = make_aabb(aabb6);
	const float EPSILON = 1e-4f;

	struct PerRay { float rgb_ray[3]; float depth_ray; uint32_t compacted; float rgbtarget[3]; float exposure_scale[3]; uint32_t img; float img_pdf, xy_pdf, x, y; int w, h; };
	std::vector<PerRay> pr(n_rays_kept);

	// phase 1 (parallel): forward compositing + target colour (:1341-1428)
	#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t i = 0; i < (int64_t)n_rays_kept; ++i) {
		uint32_t numsteps = numsteps_io[i * 2 + 0], base = numsteps_io[i * 2 + 1];
		const float* cin = coords_in_all + (size_t)base * 7;
		const orc_half* no = rgbsigma + (size_t)base * 4;
		float T = 1.f;
		float rgb_ray[3] = {0, 0, 0};
		float depth_ray = 0.f;
		uint32_t cn = 0;
		const float* ro = rays + (size_t)i * 6;
		for (; cn < numsteps; ++cn) {
			if (T < EPSILON) break;
			float rgb[3] = {network_to_rgb(h2f(no[0]), rgb_activation), network_to_rgb(h2f(no[1]), rgb_activation), network_to_rgb(h2f(no[2]), rgb_activation)};
			Vec3 pos = unwarp_position(cin, aabb);
			float dt = unwarp_dt(cin[3]);
			float dx = pos.x - ro[0], dy = pos.y - ro[1], dz = pos.z - ro[2];
			float cur_depth = std::sqrt(sum3(dx * dx, dy * dy, dz * dz));
			float density = network_to_density(h2f(no[3]), density_activation);
			const float alpha = 1.f - std::exp(-density * dt);
			const float weight = alpha * T;
			for (int c = 0; c < 3; ++c) rgb_ray[c] += weight * rgb[c];
			depth_ray += weight * cur_depth;
			T *= (1.f - alpha);
			no += 4; cin += 7;
		}
		uint32_t ray_idx = ray_indices[i];
		orc_pcg32 r = rng_advanced(rng, (int64_t)ray_idx * N_MAX_RANDOM_SAMPLES_PER_RAY);
		float img_pdf = 1.0f, xy_pdf = 1.0f;
		uint32_t img = image_idx(ray_idx, n_rays, n_images, &img_pdf);
		const orc_image& im = images[img];
		float x, y;
		random_image_pos_training(r, im.w, im.h, snap_to_pixel_centers != 0, &x, &y, img, &xy_pdf);
		float bg[3] = {background_color3[0], background_color3[1], background_color3[2]};
		if (random_bg) { bg[0] = orc_pcg32_next_float(&r); bg[1] = orc_pcg32_next_float(&r); bg[2] = orc_pcg32_next_float(&r); }
		for (int c = 0; c < 3; ++c) bg[c] = srgb_to_linear(bg[c]);
		float texsamp[4];
		read_rgba(x, y, im, texsamp);
		float rgbtarget[3];
		float es[3] = {1.0f, 1.0f, 1.0f}; // exposure_scale = exp(0.693.. * exposure) (:1403); exp(0) = 1
		if (exposure) { for (int c = 0; c < 3; ++c) es[c] = std::exp(0.6931471805599453f * exposure[(size_t)img * 3 + c]); }
		if (linear_colors || color_space == 0) {
			for (int c = 0; c < 3; ++c) rgbtarget[c] = es[c] * texsamp[c] + (1.0f - texsamp[3]) * bg[c];
			if (!linear_colors) { for (int c = 0; c < 3; ++c) { rgbtarget[c] = linear_to_srgb(rgbtarget[c]); bg[c] = linear_to_srgb(bg[c]); } }
		} else {
			for (int c = 0; c < 3; ++c) bg[c] = linear_to_srgb(bg[c]);
			if (texsamp[3] > 0) {
				for (int c = 0; c < 3; ++c) rgbtarget[c] = linear_to_srgb(es[c] * texsamp[c] / texsamp[3]) * texsamp[3] + (1.0f - texsamp[3]) * bg[c];
			} else {
				for (int c = 0; c < 3; ++c) rgbtarget[c] = bg[c];
			}
		}
		if (cn == numsteps) { for (int c = 0; c < 3; ++c) rgb_ray[c] += T * bg[c]; }
		PerRay& p = pr[i];
		for (int c = 0; c < 3; ++c) { p.rgb_ray[c] = rgb_ray[c]; p.rgbtarget[c] = rgbtarget[c]; p.exposure_scale[c] = es[c]; }
		p.depth_ray = depth_ray; p.compacted = cn; p.img = img; p.img_pdf = img_pdf; p.xy_pdf = xy_pdf; p.x = x; p.y = y; p.w = im.w; p.h = im.h;
	}

	// Compaction order: one valid serialisation of the atomicAdd at :1434. Ray-slot order while the batch holds every sample. When it overflows, the rays
	// served last are clipped (:1436-1437) -- in slot order those would always be the rays of the LAST training images (slots follow the ray index, the
	// ray index selects the image), whereas the reference's atomics clip whichever rays its blocks happen to process last. So the order then starts at a
	// ray drawn from the step's RNG (past the rays' own sub-streams) and wraps around.
	uint32_t first = 0;
	{
		uint64_t total = 0;
		for (uint32_t i = 0; i < n_rays_kept; ++i) total += pr[i].compacted;
		if (total > max_samples_compacted && n_rays_kept > 0) {
			orc_pcg32 r = rng_advanced(rng, (int64_t)n_rays * N_MAX_RANDOM_SAMPLES_PER_RAY);
			(void)orc_pcg32_next_uint(&r); // (the first draw of this stream rotates K1's allocation order)
			first = orc_pcg32_next_uint(&r) % n_rays_kept;
		}
	}
	uint32_t counter = 0;
	std::vector<uint32_t> cbase(n_rays_kept);
	for (uint32_t k = 0; k < n_rays_kept; ++k) {
		const uint32_t i = first + k < n_rays_kept ? first + k : first + k - n_rays_kept;
		uint32_t compacted_base = counter;
		counter += pr[i].compacted;
		uint32_t cn = std::min(max_samples_compacted - std::min(max_samples_compacted, compacted_base), pr[i].compacted);
		cbase[i] = compacted_base;
		pr[i].compacted = cn;
	}

	const float loss_scale = loss_scale_in / n_rays;
	const float output_l2_reg = rgb_activation == 3 ? 1e-4f : 0.0f;
	const float output_l1_reg_density = mean_density < NERF_MIN_OPTICAL_THICKNESS ? 1e-4f : 0.0f;

	// phase 2 (parallel): gradients + compaction (:1436-1556)
	#pragma omp parallel for schedule(dynamic, 64)
	for (int64_t i = 0; i < (int64_t)n_rays_kept; ++i) {
		uint32_t base = numsteps_io[i * 2 + 1];
		const PerRay& p = pr[i];
		uint32_t cn = p.compacted, compacted_base = cbase[i];
		numsteps_io[i * 2 + 0] = cn;
		numsteps_io[i * 2 + 1] = compacted_base;
		if (cn == 0) continue;
		const float* cin = coords_in_all + (size_t)base * 7;
		const orc_half* no = rgbsigma + (size_t)base * 4;
		float* cout = coords_out_all + (size_t)compacted_base * 7;
		orc_half* dout = dloss_dout_all + (size_t)compacted_base * 4;
		const float* ro = rays + (size_t)i * 6;

		LossAndGradient lg = loss_and_gradient(p.rgbtarget, p.rgb_ray, loss_type);
		{ const float pdf = p.img_pdf * p.xy_pdf; for (int c = 0; c < 3; ++c) lg.loss[c] /= pdf; } // the loss, not its gradient, is divided by the sampling pdf (:1448, :1454-1458)
		float mean_loss = sum3(lg.loss[0], lg.loss[1], lg.loss[2]) / 3.0f; // Eigen mean(): redux sum / size
		if (loss_output) loss_output[i] = mean_loss / (float)n_rays;
		if (g_error_map) { // bilinear deposit of the ray's loss into the image's error map (:1465-1491; no sharpness weighting)
			const int rx = g_error_map_res[0], ry = g_error_map_res[1];
			const float px = std::min(std::max(p.x * (float)rx - 0.5f, 0.0f), (float)rx - (1.0f + 1e-4f));
			const float py = std::min(std::max(p.y * (float)ry - 0.5f, 0.0f), (float)ry - (1.0f + 1e-4f));
			const int ix = (int)px, iy = (int)py;
			const float wx = px - (float)ix, wy = py - (float)iy;
			const int jx = std::max(std::min(ix, p.w - 2), 0), jy = std::max(std::min(iy, p.h - 2), 0); // (clamped against the IMAGE resolution, as the reference does)
			float* em = g_error_map + (size_t)p.img * rx * ry;
			const float vals[4] = {(1 - wx) * (1 - wy) * mean_loss, wx * (1 - wy) * mean_loss, (1 - wx) * wy * mean_loss, wx * wy * mean_loss};
			const size_t at[4] = {(size_t)jy * rx + jx, (size_t)jy * rx + jx + 1, (size_t)(jy + 1) * rx + jx, (size_t)(jy + 1) * rx + jx + 1};
			for (int k = 0; k < 4; ++k) {
				#pragma omp atomic
				em[at[k]] += vals[k];
			}
		}

		float rgb_ray2[3] = {0, 0, 0};
		float depth_ray2 = 0.f;
		float T = 1.f;
		for (uint32_t j = 0; j < cn; ++j) {
			std::memcpy(cout + (size_t)j * 7, cin + (size_t)j * 7, 7 * sizeof(float));
			const float* ci = cin + (size_t)j * 7;
			Vec3 pos = unwarp_position(ci, aabb);
			float dx = pos.x - ro[0], dy = pos.y - ro[1], dz = pos.z - ro[2];
			float depth = std::sqrt(sum3(dx * dx, dy * dy, dz * dz));
			float dt = unwarp_dt(ci[3]);
			float o0 = h2f(no[0]), o1 = h2f(no[1]), o2 = h2f(no[2]), o3 = h2f(no[3]);
			float rgb[3] = {network_to_rgb(o0, rgb_activation), network_to_rgb(o1, rgb_activation), network_to_rgb(o2, rgb_activation)};
			const float density = network_to_density(o3, density_activation);
			const float alpha = 1.f - std::exp(-density * dt);
			const float weight = alpha * T;
			for (int c = 0; c < 3; ++c) rgb_ray2[c] += weight * rgb[c];
			depth_ray2 += weight * depth;
			T *= (1.f - alpha);

			float suffix[3], dloss_by_drgb[3];
			for (int c = 0; c < 3; ++c) { suffix[c] = p.rgb_ray[c] - rgb_ray2[c]; dloss_by_drgb[c] = weight * lg.gradient[c]; }
			float ov[3] = {o0, o1, o2};
			for (int c = 0; c < 3; ++c) {
				dout[j * 4 + c] = f2h(loss_scale * (dloss_by_drgb[c] * network_to_rgb_derivative(ov[c], rgb_activation) + std::fmax(0.0f, output_l2_reg * ov[c])));
			}
			float density_derivative = network_to_density_derivative(o3, density_activation);
			// depth supervision is off: depth_loss_gradient = 0 -> depth_supervision = 0 * (...) (:1450-1452,:1537)
			const float depth_suffix = p.depth_ray - depth_ray2;
			const float depth_supervision = 0.0f * (T * depth - depth_suffix);
			float tv[3] = {T * rgb[0] - suffix[0], T * rgb[1] - suffix[1], T * rgb[2] - suffix[2]};
			float dloss_by_dmlp = density_derivative * (dt * (dot3(lg.gradient, tv) + depth_supervision));
			dout[j * 4 + 3] = f2h(
				loss_scale * dloss_by_dmlp +
				(o3 < 0.0f ? -output_l1_reg_density : 0.0f) +
				(o3 > -10.0f && depth < near_distance ? 1e-4f : 0.0f));
			no += 4;
		}
		if (exposure_gradient) { // :1558-1571 (xy_pdf = 1 without an error map)
			for (int c = 0; c < 3; ++c) {
				float dloss_by_dgt = -lg.gradient[c] / p.xy_pdf;
				if (!linear_colors) dloss_by_dgt /= srgb_to_linear_derivative(p.rgbtarget[c]);
				const float g = loss_scale * dloss_by_dgt * p.exposure_scale[c] * 0.6931471805599453f;
				#pragma omp atomic
				exposure_gradient[(size_t)p.img * 3 + c] += g;
			}
		}
	}
	return counter;
}

extern "C" uint32_t orc_compute_loss(
	uint32_t n_rays_kept, uint32_t n_rays, const float* aabb6, uint32_t n_rays_total, orc_pcg32 rng, uint32_t max_samples_compacted,
	float loss_scale_in, const float* background_color3, int color_space, int random_bg, int linear_colors,
	uint32_t n_images, const orc_image* images, const orc_half* rgbsigma, const uint32_t* ray_indices, const float* rays,
	uint32_t* numsteps_io, const float* coords_in_all, float* coords_out_all, orc_half* dloss_dout_all, int loss_type, float* loss_output,
	int rgb_activation, int density_activation, int snap_to_pixel_centers, float mean_density, float near_distance) {
	return orc_compute_loss_exposure(n_rays_kept, n_rays, aabb6, n_rays_total, rng, max_samples_compacted, loss_scale_in, background_color3, color_space, random_bg, linear_colors,
		n_images, images, rgbsigma, ray_indices, rays, numsteps_io, coords_in_all, coords_out_all, dloss_dout_all, loss_type, loss_output,
		rgb_activation, density_activation, snap_to_pixel_centers, mean_density, near_distance, nullptr, nullptr);
}

// ---- K19 state and CDF construction ----
extern "C" void orc_set_error_cdf(const float* cdf_x_cond_y, const float* cdf_y, const float* cdf_img, int res_x, int res_y) { g_cdf = {cdf_x_cond_y, cdf_y, cdf_img, res_x, res_y}; }
extern "C" void orc_set_error_map(float* error_map, int res_x, int res_y) { g_error_map = error_map; g_error_map_res[0] = res_x; g_error_map_res[1] = res_y; }

// construct_cdf_2d + construct_cdf_1d (src/testbed_nerf.cu:1982-2037): per row the running sum of error + 1e-10, normalised by a correctly rounded
// reciprocal and blended with 1 % uniform; then the same over the row sums per image. cdf_img receives the un-normalised image sums.
extern "C" void orc_construct_cdfs(uint32_t n_images, uint32_t height, uint32_t width, const float* error_map, float* cdf_x_cond_y, float* cdf_y, float* cdf_img) {
	const float MIN_PDF = 0.01f;
	for (uint32_t img = 0; img < n_images; ++img) {
		for (uint32_t y = 0; y < height; ++y) {
			const size_t off = ((size_t)img * height + y) * width;
			float cum = 0;
			for (uint32_t x = 0; x < width; ++x) { cum += error_map[off + x] + 1e-10f; cdf_x_cond_y[off + x] = cum; }
			cdf_y[(size_t)img * height + y] = cum;
			const float norm = 1.0f / cum; // __frcp_rn: IEEE round-to-nearest reciprocal
			for (uint32_t x = 0; x < width; ++x) cdf_x_cond_y[off + x] = (1.0f - MIN_PDF) * cdf_x_cond_y[off + x] * norm + MIN_PDF * (float)(x + 1) / (float)width;
		}
		float* cy = cdf_y + (size_t)img * height;
		float cum = 0;
		for (uint32_t y = 0; y < height; ++y) { cum += cy[y]; cy[y] = cum; }
		cdf_img[img] = cum;
		const float norm = 1.0f / cum;
		for (uint32_t y = 0; y < height; ++y) cy[y] = (1.0f - MIN_PDF) * cy[y] * norm + MIN_PDF * (float)(y + 1) / (float)height;
	}
}

// The host part of the CDF update (:3000-3015): image sums -> per-image sampling probabilities (10 % uniform) and their CDF.
extern "C" void orc_normalize_image_cdf(uint32_t n_images, const float* image_sums, float* pmf_img, float* cdf_img) {
	float cum = 0;
	for (uint32_t i = 0; i < n_images; ++i) { cum += image_sums[i]; cdf_img[i] = cum; }
	const float norm = 1.0f / cum;
	const float MIN_PMF = 0.1f;
	for (uint32_t i = 0; i < n_images; ++i) {
		pmf_img[i] = (1.0f - MIN_PMF) * image_sums[i] * norm + MIN_PMF / (float)n_images;
		cdf_img[i] = (1.0f - MIN_PMF) * cdf_img[i] * norm + MIN_PMF * (float)(i + 1) / (float)n_images;
	}
}

// K7: fill_rollover / fill_rollover_and_rescale, tcnn common_device.h:517-537
extern "C" void orc_fill_rollover(uint32_t n_target, uint32_t n_valid, float* coords, orc_half* dloss_dout) {
	if (n_valid == 0 || n_valid >= n_target) return;
	for (size_t i = (size_t)n_valid * 7; i < (size_t)n_target * 7; ++i) coords[i] = coords[i % ((size_t)n_valid * 7)];
	for (size_t i = (size_t)n_valid * 4; i < (size_t)n_target * 4; ++i) {
		float v = h2f(dloss_dout[i % ((size_t)n_valid * 4)]);
		dloss_dout[i] = f2h(v * n_valid / n_target);
	}
}

// =============================================================================================
// K15: optimizer. Ema(ExponentialDecay(Adam)) -- tcnn adam.h:48-119,:152-190; exponential_decay.h:60-72; ema.h:63-140
// =============================================================================================
extern "C" void orc_optimizer_init(orc_optimizer* o) { // configs/nerf/base.json:5-22
	o->learning_rate = 1e-2f; o->beta1 = 0.9f; o->beta2 = 0.99f; o->epsilon = 1e-15f; o->l2_reg = 1e-6f; o->ema_decay = 0.95f;
	o->decay_start = 20000; o->decay_interval = 10000; o->decay_base = 0.33f;
	o->step = 0; o->lr_factor = 1.0f;
}

extern "C" void orc_optimizer_step(orc_optimizer* o, uint32_t n_params, uint32_t n_matrix_params, float loss_scale, const float* grad,
                                   float* w_fp32, orc_half* w_half, orc_half* w_ema, float* m1, float* m2, uint32_t* param_steps) {
	// ExponentialDecay::step (uses the step count before Adam increments it)
	if (o->step == 0) o->lr_factor = 1.0f;
	if (o->step >= o->decay_start && (o->step - o->decay_start) % o->decay_interval == 0) o->lr_factor *= o->decay_base;
	const float base_lr = o->learning_rate * o->lr_factor;
	++o->step; // Adam::step
	const float beta1 = o->beta1, beta2 = o->beta2, epsilon = o->epsilon, l2_reg = o->l2_reg;
	// Ema::step
	const uint32_t current_step = o->step;
	const float ema_decay = o->ema_decay;
	const float ema_debias_old = 1 - (float)std::pow(ema_decay, current_step - 1);
	const float ema_debias_new = 1.0f / (1 - (float)std::pow(ema_decay, current_step));
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n_params; ++i) {
		float gradient = grad[i] / loss_scale;
		bool skip = ((uint32_t)i >= n_matrix_params && gradient == 0);
		if (!skip) {
			const float weight_fp = w_fp32[i];
			if ((uint32_t)i < n_matrix_params) gradient += l2_reg * weight_fp;
			const float gradient_sq = gradient * gradient;
			float first_moment = m1[i] = beta1 * m1[i] + (1 - beta1) * gradient;
			const float second_moment = m2[i] = beta2 * m2[i] + (1 - beta2) * gradient_sq;
			float learning_rate = base_lr;
			const uint32_t cs = ++param_steps[i];
			learning_rate *= std::sqrt(1 - std::pow(beta2, (float)cs)) / (1 - std::pow(beta1, (float)cs));
			const float effective_learning_rate = std::fmin(std::fmax(learning_rate / (std::sqrt(second_moment) + epsilon), 0.f), std::numeric_limits<float>::max());
			const float decayed_weight = (1 - 0.f * learning_rate) * weight_fp - std::copysign(0.f * learning_rate, weight_fp);
			float new_weight = decayed_weight - effective_learning_rate * first_moment;
			w_fp32[i] = new_weight;
			w_half[i] = f2h(new_weight);
		}
		float filtered_val = (h2f(w_ema[i]) * ema_decay * ema_debias_old + h2f(w_half[i]) * (1 - ema_decay)) * ema_debias_new;
		w_ema[i] = f2h(filtered_val);
	}
}

// =============================================================================================
// K16: density grid. src/testbed_nerf.cu:369-610
// =============================================================================================
extern "C" void orc_mark_untrained_density_grid(uint32_t n_elements, float* grid, uint32_t n_images, const orc_image* images, int clear_visible) {
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n_elements; ++i) {
		uint32_t level = (uint32_t)i / GRID_CELLS, pos_idx = (uint32_t)i % GRID_CELLS;
		uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
		float s = std::scalbn(1.0f, (int)level);
		float pos[3] = {
			(((float)x + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
			(((float)y + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
			(((float)z + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f};
		float voxel_radius = 0.5f * SQRT3 * s / NERF_GRIDSIZE;
		int count = 0;
		for (uint32_t j = 0; j < n_images; ++j) {
			const orc_image& im = images[j];
			if (im.lens_mode == 2 || im.lens_mode == 3) { count++; break; } // f-theta / lat-long: "not supported for now", every cell counts as visible (:391-395)
			float half_resx = im.w * 0.5f, half_resy = im.h * 0.5f;
			const float* xf = im.xform;
			float ploc[3] = {pos[0] - xf[9], pos[1] - xf[10], pos[2] - xf[11]};
			float cx = dot3(ploc, xf + 0), cy = dot3(ploc, xf + 3), cz = dot3(ploc, xf + 6);
			if (cz > 0.f) {
				if (std::fabs(cx) - voxel_radius < cz / im.fx * half_resx && std::fabs(cy) - voxel_radius < cz / im.fy * half_resy) {
					count++;
					if (count > 0) break;
				}
			}
		}
		if (clear_visible || (grid[i] < 0) != (count <= 0)) grid[i] = (count > 0) ? 0.f : -1.f;
	}
}

extern "C" void orc_generate_grid_samples(uint32_t n_elements, orc_pcg32 rng_in, uint32_t step, const float* aabb6, const float* grid_in,
                                          float* positions3, uint32_t* indices, uint32_t n_cascades, float thresh) {
	const AABB aabb = make_aabb(aabb6);
	#pragma omp parallel for schedule(static)
	for (int64_t ii = 0; ii < (int64_t)n_elements; ++ii) {
		uint32_t i = (uint32_t)ii;
		orc_pcg32 rng = rng_advanced(rng_in, (int64_t)i * 4);
		uint32_t level = (uint32_t)(orc_pcg32_next_float(&rng) * n_cascades) % n_cascades;
		uint32_t idx = 0;
		for (uint32_t j = 0; j < 10; ++j) {
			idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % GRID_CELLS;
			idx += level * GRID_CELLS;
			if (grid_in[idx] > thresh) break;
		}
		uint32_t pos_idx = idx % GRID_CELLS;
		uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
		float rx = orc_pcg32_next_float(&rng), ry = orc_pcg32_next_float(&rng), rz = orc_pcg32_next_float(&rng);
		float s = std::scalbn(1.0f, (int)level);
		Vec3 pos = {
			(((float)x + rx) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
			(((float)y + ry) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
			(((float)z + rz) / NERF_GRIDSIZE - 0.5f) * s + 0.5f};
		Vec3 wp = warp_position(pos, aabb);
		positions3[(size_t)i * 3 + 0] = wp.x; positions3[(size_t)i * 3 + 1] = wp.y; positions3[(size_t)i * 3 + 2] = wp.z;
		indices[i] = idx;
	}
}

extern "C" void orc_splat_and_ema(uint32_t n_samples, const uint32_t* indices, const orc_half* density, uint32_t n_elements, float decay, float* grid) {
	std::vector<float> tmp(n_elements, 0.f);
	for (uint32_t i = 0; i < n_samples; ++i) {
		float mlp = std::exp(h2f(density[i])); // density activation Exponential (:506)
		float optical_thickness = mlp * std::scalbn(MIN_CONE_STEPSIZE, 0);
		// atomicMax on the uint bit pattern (:509-511)
		uint32_t a, b; std::memcpy(&a, &tmp[indices[i]], 4); std::memcpy(&b, &optical_thickness, 4);
		if (b > a) tmp[indices[i]] = optical_thickness;
	}
	for (uint32_t i = 0; i < n_elements; ++i) {
		float prev_val = grid[i];
		grid[i] = (prev_val < 0.f) ? prev_val : std::fmax(prev_val * decay, tmp[i]);
	}
}

// mean over the first cascade of max(v,0)/n (:2851-2852). Summation order is unspecified in the
// reference (tcnn::reduce_sum); the oracle sums in double.
extern "C" float orc_density_grid_mean(const float* grid) {
	double s = 0.0;
	for (uint32_t i = 0; i < GRID_CELLS; ++i) s += (double)(std::fmax(grid[i], 0.f) / (float)GRID_CELLS);
	return (float)s;
}

extern "C" void orc_bitfield(uint32_t n_cascades_used, const float* grid, float mean_density, uint8_t* bitfield) {
	const uint32_t n_bytes = GRID_CELLS / 8 * NERF_CASCADES, n_nonzero = GRID_CELLS / 8 * n_cascades_used;
	float thresh = std::min(NERF_MIN_OPTICAL_THICKNESS, mean_density);
	for (uint32_t i = 0; i < n_bytes; ++i) {
		if (i >= n_nonzero) { bitfield[i] = 0; continue; }
		uint8_t bits = 0;
		for (uint8_t j = 0; j < 8; ++j) bits |= grid[(size_t)i * 8 + j] > thresh ? ((uint8_t)1 << j) : 0;
		bitfield[i] = bits;
	}
	for (uint32_t level = 1; level < NERF_CASCADES; ++level) {
		const uint8_t* prev = bitfield + grid_mip_offset(level - 1) / 8;
		uint8_t* next = bitfield + grid_mip_offset(level) / 8;
		for (uint32_t i = 0; i < GRID_CELLS / 64; ++i) {
			uint8_t bits = 0;
			for (uint8_t j = 0; j < 8; ++j) bits |= prev[(size_t)i * 8 + j] > 0 ? ((uint8_t)1 << j) : 0;
			uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8, y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8, z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
			next[morton3D(x, y, z)] |= bits;
		}
	}
}

// =============================================================================================
// Classic single-NeRF render (Testbed::render_nerf, src/testbed_nerf.cu:2354-2499), ERenderMode::Shade,
// perspective camera, no lens distortion / masks / envmap / depth of field:
//   init_rays_with_payload_kernel_nerf :1809-1978 + pixel_to_ray common_device.cuh:260-317
//   advance_pos_nerf :612-664, generate_next_nerf_network_inputs :705-766, composite_kernel_nerf :767-989,
//   compact_kernel_nerf :1784-1807 (a finished ray is shaded only if its alpha exceeds 0.001), shade_kernel_nerf :1748-1782
//   low-discrepancy jitter: random_val.cuh:159-322 (Burley's shuffled scrambled Sobol)
//   CudaRenderBuffer::accumulate / tonemap: src/render_buffer.cu:235-266, :268-349, :540-567, :606-660
// The reference's wavefront (compaction every 1-8 steps) only schedules the work: per ray the samples, their order and the
// termination test are those of the serial loop below.
// =============================================================================================
namespace {
inline uint32_t sobol(uint32_t index, uint32_t dim) {
	static const uint32_t directions[2][32] = {
		{0x80000000, 0x40000000, 0x20000000, 0x10000000, 0x08000000, 0x04000000, 0x02000000, 0x01000000,
		 0x00800000, 0x00400000, 0x00200000, 0x00100000, 0x00080000, 0x00040000, 0x00020000, 0x00010000,
		 0x00008000, 0x00004000, 0x00002000, 0x00001000, 0x00000800, 0x00000400, 0x00000200, 0x00000100,
		 0x00000080, 0x00000040, 0x00000020, 0x00000010, 0x00000008, 0x00000004, 0x00000002, 0x00000001},
		{0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000,
		 0x80800000, 0xc0c00000, 0xa0a00000, 0xf0f00000, 0x88880000, 0xcccc0000, 0xaaaa0000, 0xffff0000,
		 0x80008000, 0xc000c000, 0xa000a000, 0xf000f000, 0x88008800, 0xcc00cc00, 0xaa00aa00, 0xff00ff00,
		 0x80808080, 0xc0c0c0c0, 0xa0a0a0a0, 0xf0f0f0f0, 0x88888888, 0xcccccccc, 0xaaaaaaaa, 0xffffffff}};
	uint32_t X = 0;
	for (uint32_t bit = 0; bit < 32; bit++) X ^= ((index >> bit) & 1) * directions[dim][bit];
	return X;
}
inline uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
inline uint32_t reverse_bits(uint32_t x) {
	x = (((x & 0xaaaaaaaa) >> 1) | ((x & 0x55555555) << 1));
	x = (((x & 0xcccccccc) >> 2) | ((x & 0x33333333) << 2));
	x = (((x & 0xf0f0f0f0) >> 4) | ((x & 0x0f0f0f0f) << 4));
	x = (((x & 0xff00ff00) >> 8) | ((x & 0x00ff00ff) << 8));
	return ((x >> 16) | (x << 16));
}
inline uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
	x += seed; x ^= x * 0x6c50b47cu; x ^= x * 0xb82f1e52u; x ^= x * 0xc7afe638u; x ^= x * 0x8d22f6e6u;
	return x;
}
inline uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) { return reverse_bits(laine_karras_permutation(reverse_bits(x), seed)); }
inline float ld_random_val(uint32_t index, uint32_t seed, uint32_t dim = 0) {
	const float S = float(1.0 / (1ull << 32));
	index = nested_uniform_scramble_base2(index, seed);
	return (float)nested_uniform_scramble_base2(sobol(index, dim), hash_combine(seed, dim)) * S;
}
inline void ld_random_val_2d(uint32_t index, uint32_t seed, float* out) {
	const float S = float(1.0 / (1ull << 32));
	index = nested_uniform_scramble_base2(index, seed);
	for (uint32_t i = 0; i < 2; ++i) out[i] = (float)nested_uniform_scramble_base2(sobol(index, i), hash_combine(seed, i)) * S;
}
inline float fractf(float x) { return x - std::floor(x); }
// tonemap(x, curve), src/render_buffer.cu:272-329
inline void tonemap_curve(float c[3], int curve) {
	if (curve == 0) return; // Identity
	for (int k = 0; k < 3; ++k) c[k] = std::fmax(c[k], 0.f);
	float k0, k1, k2, k3, k4, k5;
	if (curve == 1) { // ACES
		k0 = 0.6f * 0.6f * 2.51f; k1 = 0.6f * 0.03f; k2 = 0.0f; k3 = 0.6f * 0.6f * 2.43f; k4 = 0.6f * 0.59f; k5 = 0.14f;
	} else if (curve == 2) { // Hable
		const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
		k0 = A * F - A * E; k1 = C * B * F - B * E; k2 = 0.0f; k3 = A * F; k4 = B * F; k5 = D * F * F;
		const float W = 11.2f;
		const float nom = k0 * (W * W) + k1 * W + k2, denom = k3 * (W * W) + k4 * W + k5;
		const float white_scale = denom / nom;
		k0 = 4.0f * k0 * white_scale; k1 = 2.0f * k1 * white_scale; k2 = k2 * white_scale; k3 = 4.0f * k3; k4 = 2.0f * k4;
	} else { // Reinhard
		const float Y = sum3(0.2126f * c[0], 0.7152f * c[1], 0.0722f * c[2]);
		for (int k = 0; k < 3; ++k) c[k] = c[k] * (1.f / (Y + 1.0f));
		return;
	}
	for (int k = 0; k < 3; ++k) { const float sq = c[k] * c[k]; c[k] = (sq * k0 + k1 * c[k] + k2) / (k3 * sq + k4 * c[k] + k5); }
}
inline void ld_random_pixel_offset(uint32_t spp, float* out) {
	float a[2], b[2];
	ld_random_val_2d(0, 0xdeadbeef, a);
	ld_random_val_2d(spp, 0xdeadbeef, b);
	for (int i = 0; i < 2; ++i) out[i] = fractf((0.5f - a[i]) + b[i]);
}
} // namespace

extern "C" void orc_render_nerf(const orc_model* m, const orc_half* params, const uint8_t* bitfield, const orc_render_config* c, float* out_rgba, uint64_t* n_samples_out) {
	const int W = c->width, H = c->height;
	const AABB train_aabb = make_aabb(c->aabb), render_aabb = make_aabb(c->render_aabb);
	std::vector<float> accumulate((size_t)W * H * 4, 0.f);
	uint64_t n_samples_total = 0;
	for (int s = 0; s < c->spp; ++s) {
		std::vector<float> frame((size_t)W * H * 4, 0.f);
		float offset[2];
		ld_random_pixel_offset(c->snap_to_pixel_centers ? 0 : (uint32_t)s, offset);
		uint64_t n_samples_spp = 0;
		#pragma omp parallel for schedule(dynamic, 64) reduction(+ : n_samples_spp)
		for (int64_t idx = 0; idx < (int64_t)W * H; ++idx) {
			const int x = (int)(idx % W), y = (int)(idx / W);
			// pixel_to_ray, perspective
			const float uvx = ((float)x + offset[0]) / (float)W, uvy = ((float)y + offset[1]) / (float)H;
			const float dcam[3] = {(uvx - c->screen_center[0]) * (float)W / c->fx, (uvy - c->screen_center[1]) * (float)H / c->fy, 1.0f};
			const float* cm = c->camera;
			const float row0[3] = {cm[0], cm[3], cm[6]}, row1[3] = {cm[1], cm[4], cm[7]}, row2[3] = {cm[2], cm[5], cm[8]};
			Vec3 d = {dot3(row0, dcam), dot3(row1, dcam), dot3(row2, dcam)};
			Vec3 o = {cm[9], cm[10], cm[11]};
			o = {o.x + d.x * c->near_distance, o.y + d.y * c->near_distance, o.z + d.z * c->near_distance};
			// init_rays_with_payload_kernel_nerf
			const float z = sum3(d.x * d.x, d.y * d.y, d.z * d.z);
			if (z > 0.f) { const float nrm = std::sqrt(z); d = {d.x / nrm, d.y / nrm, d.z / nrm}; }
			float tmin, tmax;
			aabb_ray_intersect(render_aabb, o, d, &tmin, &tmax);
			float t = std::fmax(tmin, 0.0f) + 1e-6f;
			if (!aabb_contains(render_aabb, Vec3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t})) continue;
			// advance_pos_nerf
			const Vec3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
			const float cone_angle = c->cone_angle_constant;
			t += ld_random_val((uint32_t)s, (uint32_t)idx * 786433u) * calc_dt(t, cone_angle);
			bool alive = true;
			auto march_to_occupied = [&](Vec3& pos, float& dt) -> bool { // false: left the render box
				while (true) {
					pos = {o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
					if (!aabb_contains(render_aabb, pos)) return false;
					dt = calc_dt(t, cone_angle);
					const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
					if (density_grid_occupied_at(pos, bitfield, mip)) return true;
					t = advance_to_next_voxel(t, cone_angle, pos, d, idir, NERF_GRIDSIZE >> mip);
				}
			};
			{ Vec3 p; float dt; alive = march_to_occupied(p, dt); }
			float rgba[4] = {0.f, 0.f, 0.f, 0.f};
			const Vec3 wd = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f};
			while (alive) {
				// generate_next_nerf_network_inputs: up to 8 samples, then the network, then composite_kernel_nerf
				float coords[8 * 7];
				orc_half out[8 * 4];
				uint32_t n = 0;
				for (; n < 8; ++n) {
					Vec3 pos; float dt;
					if (!march_to_occupied(pos, dt)) break;
					const Vec3 wp = warp_position(pos, train_aabb);
					float* cc = coords + n * 7;
					cc[0] = wp.x; cc[1] = wp.y; cc[2] = wp.z; cc[3] = warp_dt(dt); cc[4] = wd.x; cc[5] = wd.y; cc[6] = wd.z;
					t += dt;
				}
				if (n) orc_nerf_inference(m, params, n, coords, out);
				n_samples_spp += n;
				uint32_t j = 0;
				for (; j < n; ++j) {
					const float T = 1.f - rgba[3];
					const float dt = unwarp_dt(coords[j * 7 + 3]);
					const float alpha = 1.f - std::exp(-network_to_density(h2f(out[j * 4 + 3]), c->density_activation) * dt);
					const float weight = alpha * T;
					for (int k = 0; k < 3; ++k) rgba[k] += network_to_rgb(h2f(out[j * 4 + k]), c->rgb_activation) * weight;
					rgba[3] += weight;
					if (rgba[3] > (1.0f - c->min_transmittance)) {
						const float w = rgba[3];
						for (int k = 0; k < 4; ++k) rgba[k] /= w;
						break;
					}
				}
				if (j < 8) alive = false; // terminated, or left the box before the chunk was full
			}
			if (rgba[3] > 0.001f) { // compact_kernel_nerf keeps only these; shade_kernel_nerf
				float tmp[4] = {rgba[0], rgba[1], rgba[2], rgba[3]};
				if (!c->train_in_linear_colors) for (int k = 0; k < 3; ++k) tmp[k] = srgb_to_linear(tmp[k]);
				float* f = &frame[(size_t)idx * 4];
				const float one_minus = 1.0f - tmp[3];
				for (int k = 0; k < 4; ++k) f[k] = tmp[k] + f[k] * one_minus;
			}
		}
		n_samples_total += n_samples_spp;
		// accumulate_kernel
		const float sample_count = (float)s;
		for (size_t i = 0; i < (size_t)W * H; ++i) {
			float color[4] = {frame[i * 4], frame[i * 4 + 1], frame[i * 4 + 2], frame[i * 4 + 3]};
			float* tmp = &accumulate[i * 4];
			if (c->color_space == 1) for (int k = 0; k < 3; ++k) color[k] = linear_to_srgb(color[k]);
			for (int k = 0; k < 4; ++k) tmp[k] = (tmp[k] * sample_count + color[k]) / (sample_count + 1);
		}
	}
	// tonemap_kernel (identity curve, no DLSS clamp)
	float bg[4] = {c->background_color[0], c->background_color[1], c->background_color[2], c->background_color[3]};
	if (c->color_space != 1) for (int k = 0; k < 3; ++k) bg[k] = srgb_to_linear(bg[k]);
	const float exposure_scale = std::pow(2.0f, c->exposure);
	for (size_t i = 0; i < (size_t)W * H; ++i) {
		float color[4] = {accumulate[i * 4], accumulate[i * 4 + 1], accumulate[i * 4 + 2], accumulate[i * 4 + 3]};
		const float weight = (1 - color[3]) * bg[3];
		for (int k = 0; k < 3; ++k) color[k] += bg[k] * weight;
		color[3] += weight;
		for (int k = 0; k < 3; ++k) {
			float v = color[k];
			if (c->color_space == 1) v = srgb_to_linear(v);
			color[k] = v * exposure_scale;
		}
		tonemap_curve(color, c->tonemap_curve);
		if (c->output_srgb) for (int k = 0; k < 3; ++k) color[k] = linear_to_srgb(color[k]);
		for (int k = 0; k < 4; ++k) out_rgba[i * 4 + k] = color[k];
	}
	if (n_samples_out) *n_samples_out = n_samples_total;
}

// =============================================================================================
// Blender multi-NeRF renderer (the fork's addition): NerfRenderer::render, src/nerf_renderer.cu:565-791
//   init_global_rays_kernel :17-92 (+ perspective_pixel_to_ray, camera_models.cuh:205-241), init_proxy_rays_kernel :94-146,
//   compact_rays_kernel :240-270, hit_test_and_march :149-211 (helpers in src/nerf_utils.cu), march_active_rays :274-315,
//   cull_global_rays_and_set_proxy_rays_active_kernel :376-428, march_proxy_rays_and_generate_next_network_inputs :317-373,
//   composite_proxy_ray_colors_kernel :431-517, shade_buffer_with_rays_kernel :519-563, the wave loop :661-791,
//   then Testbed::bl_render_frame (src/testbed.cu:2675-2693): accumulate + tonemap with the request's colour space.
// Camera models and depth of field: camera_models.cuh:82-241; masks: nerf/mask_3D.cuh, nerf/render_modifiers.cuh:28-62 (mask list of a NeRF),
// :127-140 (ray/mask culling at init), :490-496 (mask weight while compositing). The result depends on the reference's wave schedule (the nearest live proxy ray is
// re-selected once per wave of clamp(n_initial / n_alive, 1, 8) steps), so the waves are followed literally, all rays in lockstep.
// =============================================================================================
namespace {
struct M4 { float m[16]; }; // column-major
inline Vec3 xform_point(const M4& M, const Vec3& p) { // (M * p.homogeneous()).head<3>()
	return {sum4(M.m[0] * p.x, M.m[4] * p.y, M.m[8] * p.z, M.m[12]), sum4(M.m[1] * p.x, M.m[5] * p.y, M.m[9] * p.z, M.m[13]), sum4(M.m[2] * p.x, M.m[6] * p.y, M.m[10] * p.z, M.m[14])};
}
inline Vec3 xform_dir(const M4& M, const Vec3& d) { // M.topLeftCorner<3,3>() * d
	return {sum3(M.m[0] * d.x, M.m[4] * d.y, M.m[8] * d.z), sum3(M.m[1] * d.x, M.m[5] * d.y, M.m[9] * d.z), sum3(M.m[2] * d.x, M.m[6] * d.y, M.m[10] * d.z)};
}
inline Vec3 normalized(const Vec3& v) {
	const float z = sum3(v.x * v.x, v.y * v.y, v.z * v.z);
	if (z > 0.f) { const float n = std::sqrt(z); return {v.x / n, v.y / n, v.z / n}; }
	return v;
}
inline bool invert4(const float* a, float* out) { // Gauss-Jordan with partial pivoting, column-major in and out
	double m[4][8];
	for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { m[r][c] = a[c * 4 + r]; m[r][4 + c] = r == c ? 1.0 : 0.0; }
	for (int c = 0; c < 4; ++c) {
		int piv = c;
		for (int r = c + 1; r < 4; ++r) if (std::fabs(m[r][c]) > std::fabs(m[piv][c])) piv = r;
		if (std::fabs(m[piv][c]) < 1e-30) return false;
		if (piv != c) for (int k = 0; k < 8; ++k) std::swap(m[piv][k], m[c][k]);
		const double d = m[c][c];
		for (int k = 0; k < 8; ++k) m[c][k] /= d;
		for (int r = 0; r < 4; ++r) if (r != c) { const double f = m[r][c]; for (int k = 0; k < 8; ++k) m[r][k] -= f * m[c][k]; }
	}
	for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[c * 4 + r] = (float)m[r][4 + c];
	return true;
}
struct ProxyRay { Vec3 o, d; float t; uint32_t n_steps; bool alive, active; };
struct GlobalRay { Vec3 o, d; float rgba[4]; uint32_t idx; bool alive; };

// hit_test_and_march (:149-211) with no masks: advances t to the next sample position in an occupied cell
inline bool hit_test_and_march(const Vec3& o, const Vec3& d, const Vec3& idir, float t_in, const AABB& box, const uint8_t* bitfield, float cone_angle, float* t_out, float* dt_out) {
	float t = t_in, dt = 0.0f, prev_t = t;
	while (true) {
		const Vec3 pos = {o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
		if (!aabb_contains(box, pos)) { *t_out = prev_t; if (dt_out) *dt_out = dt; return false; }
		dt = calc_dt(t, cone_angle);
		const uint32_t mip = (uint32_t)std::max(0, mip_from_dt(dt, pos));
		if (density_grid_occupied_at(pos, bitfield, mip)) break;
		prev_t = t;
		t = advance_to_next_voxel(t, cone_angle, pos, d, idir, NERF_GRIDSIZE >> mip);
	}
	*t_out = t; if (dt_out) *dt_out = dt;
	return true;
}
} // namespace

namespace {
struct MaskL { int32_t shape, mode; M4 itransform; float config[6]; float feather, opacity; };
inline float mask_signed_distance(const MaskL& m, const Vec3& p) { // Mask3D::signed_distance_to_point, mask_3D.cuh:162-183 (sdf_* :31-45)
	const Vec3 q = xform_point(m.itransform, p);
	float d = 0.0f;
	if (m.shape == 0) {
		const float dx = std::fabs(q.x) - 0.5f * m.config[0], dy = std::fabs(q.y) - 0.5f * m.config[1], dz = std::fabs(q.z) - 0.5f * m.config[2];
		const float ox = std::fmax(dx, 0.0f), oy = std::fmax(dy, 0.0f), oz = std::fmax(dz, 0.0f);
		d = std::sqrt(sum3(ox * ox, oy * oy, oz * oz)) + std::fmin(std::fmax(dx, std::fmax(dy, dz)), 0.0f);
	} else if (m.shape == 1) {
		const float dr = std::fabs(std::sqrt(q.y * q.y + q.x * q.x)) - m.config[0], dh = std::fabs(q.z) - 0.5f * m.config[1];
		const float orr = std::fmax(dr, 0.0f), oh = std::fmax(dh, 0.0f);
		d = std::sqrt(orr * orr + oh * oh) + std::fmin(std::fmax(dr, dh), 0.0f);
	} else if (m.shape == 2) {
		d = std::sqrt(sum3(q.x * q.x, q.y * q.y, q.z * q.z)) - m.config[0];
	} else {
		d = -1.0f;
	}
	return d * (m.mode == 0 ? 1.0f : -1.0f);
}
inline float mask_sample(const MaskL& m, const Vec3& p) { // Mask3D::sample, :194-213
	const float k = m.mode == 0 ? 1.0f : -1.0f;
	if (m.shape == 3) return k;
	const float d = mask_signed_distance(m, p);
	const float alpha = m.feather == 0.0f ? (d < 0.0f ? 1.0f : 0.0f) : std::fmin(std::fmax(0.5f - d / m.feather, 0.0f), 1.0f);
	return m.opacity * alpha * k;
}
inline bool plane_hit(const Vec3& o, const Vec3& d, float nz, float pz, float* t) { // intersect_plane_ray :74-81, n = (0,0,nz), p = (0,0,pz)
	const float denom = nz * d.z;
	if (denom > 1e-6f) { *t = ((pz - o.z) * nz) / denom; return *t >= 0.0f; }
	return false;
}
inline bool mask_intersects_ray(const MaskL& m, const Vec3& ro, const Vec3& rd) { // Mask3D::intersects_ray, :215-247
	if (m.mode == 1) return true;
	if (m.shape == 3) return m.mode == 0;
	const Vec3 o = xform_point(m.itransform, ro);
	const Vec3 d = normalized(xform_dir(m.itransform, rd));
	if (m.shape == 0) { // ray_intersects_box :46-56
		float lo = -INFINITY, hi = INFINITY;
		const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
		for (int a = 0; a < 3; ++a) {
			const float size = m.config[a] + 0.5f * m.feather, inv = 1.0f / dd[a];
			const float t0 = (-0.5f * size - oo[a]) * inv, t1 = (0.5f * size - oo[a]) * inv;
			lo = std::fmax(lo, std::fmin(t0, t1)); hi = std::fmin(hi, std::fmax(t0, t1));
		}
		return lo <= hi;
	}
	if (m.shape == 2) { // ray_intersects_sphere :59-63
		const float radius = m.config[0] + 0.5f * m.feather;
		const float dot = sum3(d.x * o.x, d.y * o.y, d.z * o.z);
		return !((dot * dot - (sum3(o.x * o.x, o.y * o.y, o.z * o.z) - radius * radius)) < 0.0f);
	}
	const float radius = m.config[0] + 0.5f * m.feather, height = m.config[1] + 0.5f * m.feather; // ray_intersects_cylinder :83-123
	const float a = d.x * d.x + d.y * d.y, b = 2.0f * (d.x * o.x + d.y * o.y), c = (o.x * o.x + o.y * o.y) - radius * radius;
	const float disc = b * b - 4.0f * a * c;
	if (disc < 0.0f) return false;
	const float sq = std::sqrt(disc), a2 = 2.0f * a, h2 = 0.5f * height;
	if (a2 > 1e-6f) {
		const float z0 = o.z + ((-b - sq) / a2) * d.z, z1 = o.z + ((-b + sq) / a2) * d.z;
		if ((z0 >= -h2 && z0 <= h2) || (z1 >= -h2 && z1 <= h2)) return true;
	}
	float t = 0.0f;
	if (plane_hit(o, d, 1.0f, h2, &t)) { const float px = o.x + t * d.x, py = o.y + t * d.y; if (px * px + py * py <= radius * radius) return true; }
	if (plane_hit(o, d, -1.0f, -h2, &t)) { const float px = o.x + t * d.x, py = o.y + t * d.y; if (px * px + py * py <= radius * radius) return true; }
	return false;
}
inline void mul4(const float* a, const float* b, float* out) {
	for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) { float v = 0.f; for (int k = 0; k < 4; ++k) v += a[k * 4 + r] * b[c * 4 + k]; out[c * 4 + r] = v; }
}
// RenderModifiers (render_modifiers.cuh:28-62): own masks, the request's masks moved into the NeRF's frame, and an `All` mask of the opposite mode in front
inline std::vector<MaskL> build_mask_list(const orc_nerf_instance& in, const M4& nerf_itransform, const orc_blender_request* rq) {
	std::vector<MaskL> out;
	auto push = [&](const orc_mask& m, const float* to_local) {
		MaskL d{};
		d.shape = m.shape; d.mode = m.mode; d.feather = m.feather; d.opacity = m.opacity;
		std::memcpy(d.config, m.config, sizeof(d.config));
		float t[16];
		if (to_local) mul4(to_local, m.transform, t); else std::memcpy(t, m.transform, 64);
		if (!invert4(t, d.itransform.m)) std::memset(d.itransform.m, 0, 64);
		out.push_back(d);
	};
	for (uint32_t k = 0; k < in.n_masks; ++k) push(in.masks[k], nullptr);
	for (uint32_t k = 0; k < rq->n_masks; ++k) push(rq->masks[k], nerf_itransform.m);
	if (!out.empty() && out[0].shape != 3) {
		MaskL all{};
		all.shape = 3; all.mode = out[0].mode == 0 ? 1 : 0; all.opacity = 1.0f;
		for (int k = 0; k < 4; ++k) all.itransform.m[k * 5] = 1.0f;
		out.insert(out.begin(), all);
	}
	return out;
}
inline void square2disk_shirley(float a, float b, float out[2]) { // random_val.cuh:109-125
	const float PI = 3.14159265358979323846f;
	float phi, r;
	if (a * a > b * b) { r = a; phi = (PI / 4.0f) * (b / a); } else { r = b; phi = (PI / 2.0f) - (PI / 4.0f) * (a / b); }
	out[0] = r * std::cos(phi); out[1] = r * std::sin(phi);
}
inline Vec3 cam_rotate(const float* cm, const Vec3& v) {
	const float row0[3] = {cm[0], cm[3], cm[6]}, row1[3] = {cm[1], cm[4], cm[7]}, row2[3] = {cm[2], cm[5], cm[8]}, vv[3] = {v.x, v.y, v.z};
	return {dot3(row0, vv), dot3(row1, vv), dot3(row2, vv)};
}
// init_global_rays_kernel's switch over the camera model (nerf_renderer.cu:44-82) with sample index 0
inline void blender_pixel_to_ray(const orc_blender_request* rq, const float offset[2], uint32_t x, uint32_t y, Vec3& origin, Vec3& dir) {
	const float W = (float)rq->width, H = (float)rq->height;
	const float* cm = rq->camera;
	if (rq->camera_model == 0) { // perspective_pixel_to_ray, camera_models.cuh:205-241
		const float uvx = ((float)x + offset[0]) / W, uvy = ((float)y + offset[1]) / H;
		dir = cam_rotate(cm, Vec3{(uvx - 0.5f) * W / rq->focal_length, (uvy - 0.5f) * H / rq->focal_length, 1.0f});
		origin = {cm[9], cm[10], cm[11]};
	} else if (rq->camera_model == 1) { // quadrilateral_hexahedron_pixel_to_ray, :82-110
		const float ux = ((float)x + 0.5f) / W, uy = ((float)y + 0.5f) / H;
		const float* q = rq->quadrilateral_hexahedron;
		float fp[3], bp[3];
		for (int k = 0; k < 3; ++k) {
			const float f_ab = q[0 + k] + ux * (q[3 + k] - q[0 + k]), f_dc = q[6 + k] + ux * (q[9 + k] - q[6 + k]);
			fp[k] = f_ab + uy * (f_dc - f_ab);
			const float b_ab = q[12 + k] + ux * (q[15 + k] - q[12 + k]), b_dc = q[18 + k] + ux * (q[21 + k] - q[18 + k]);
			bp[k] = b_ab + uy * (b_dc - b_ab);
		}
		Vec3 d = {fp[0] - bp[0], fp[1] - bp[1], fp[2] - bp[2]};
		d = {d.x / d.z, d.y / d.z, d.z / d.z};
		const Vec3 o = cam_rotate(cm, Vec3{bp[0], bp[1], bp[2]});
		origin = {o.x + cm[9], o.y + cm[10], o.z + cm[11]};
		dir = cam_rotate(cm, d);
	} else { // spherical_quadrilateral_pixel_to_ray, :159-203 (walk_along_sphere / walk_along_circle :137-157)
		const float PI = 3.14159265358979323846f;
		const float sw = rq->spherical_quadrilateral[0], sh = rq->spherical_quadrilateral[1], curvature = rq->spherical_quadrilateral[2];
		const float max_len = std::sqrt(sw * sw + sh * sh);
		const float ux = 2.0f * (((float)x + 0.5f) / W - 0.5f), uy = 2.0f * (((float)y + 0.5f) / H - 0.5f);
		const float px = sw * ux, py = sh * uy;
		const float az = std::atan2(py, px), r = std::sqrt(px * px + py * py);
		float rz0 = 0.0f, rz1 = 0.0f;
		const float arc_t = r / (2.0f * max_len);
		if (!(arc_t == 0.0f || max_len == 0.0f)) {
			if (curvature == 0.0f) rz0 = max_len * arc_t;
			else { const float tpc = 2.0f * PI * curvature, s_tpc = max_len / tpc; rz0 = s_tpc * std::sin(tpc * arc_t); rz1 = s_tpc * (1.0f - std::cos(tpc * arc_t)); }
		}
		const Vec3 o_local = {rz0 * std::cos(az), rz0 * std::sin(az), rz1};
		Vec3 d_local = {0.0f, 0.0f, 1.0f};
		if (curvature != 0.0f) {
			const Vec3 n = normalized(Vec3{0.0f - o_local.x, 0.0f - o_local.y, max_len / (2.0f * PI * curvature) - o_local.z});
			const float k = curvature > 0.0f ? 1.0f : -1.0f;
			d_local = {k * n.x, k * n.y, k * n.z};
		}
		const Vec3 o = cam_rotate(cm, o_local);
		origin = {o.x + cm[9], o.y + cm[10], o.z + cm[11]};
		dir = cam_rotate(cm, d_local);
	}
	if (rq->aperture_size > 0.0f) { // thin-lens depth of field (:99-104, :196-201, :232-237)
		const Vec3 lookat = {origin.x + dir.x * rq->focus_z, origin.y + dir.y * rq->focus_z, origin.z + dir.z * rq->focus_z};
		float u[2], disk[2];
		ld_random_val_2d(0u, x * 19349663u + y * 96925573u, u);
		square2disk_shirley(u[0] * 2.0f - 1.0f, u[1] * 2.0f - 1.0f, disk);
		const float bx = rq->aperture_size * disk[0], by = rq->aperture_size * disk[1];
		origin.x += cm[0] * bx + cm[3] * by; origin.y += cm[1] * bx + cm[4] * by; origin.z += cm[2] * bx + cm[5] * by;
		dir = {(lookat.x - origin.x) / rq->focus_z, (lookat.y - origin.y) / rq->focus_z, (lookat.z - origin.z) / rq->focus_z};
	}
	origin = {origin.x + dir.x * rq->near_distance, origin.y + dir.y * rq->near_distance, origin.z + dir.z * rq->near_distance};
}
} // namespace

extern "C" void orc_blender_render(const orc_blender_request* rq, uint32_t n_nerfs, const orc_nerf_instance* nerfs, float* out_rgba, uint64_t* n_samples_out) {
	const int W = rq->width, H = rq->height;
	const int skip = 1 << rq->mip;
	const int SW = (W + skip - 1) / skip, SH = (H + skip - 1) / skip; // DownsampleInfo::MakeFromMip (common.h:337-355)
	const uint32_t n_init = (uint32_t)SW * SH;
	std::vector<M4> T(n_nerfs), IT(n_nerfs);
	std::vector<AABB> render_box(n_nerfs), train_box(n_nerfs);
	std::vector<float> cone(n_nerfs);
	std::vector<std::vector<MaskL>> masks(n_nerfs);
	for (uint32_t n = 0; n < n_nerfs; ++n) {
		std::memcpy(T[n].m, nerfs[n].transform, 64);
		if (!invert4(nerfs[n].transform, IT[n].m)) std::memset(IT[n].m, 0, 64);
		render_box[n] = make_aabb(nerfs[n].render_aabb); train_box[n] = make_aabb(nerfs[n].train_aabb);
		cone[n] = nerfs[n].aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);
		masks[n] = build_mask_list(nerfs[n], IT[n], rq);
	}
	// init_global_rays_kernel, sample_index 0
	float offset[2];
	ld_random_pixel_offset(0, offset);
	std::vector<GlobalRay> rays(n_init);
	std::vector<std::vector<ProxyRay>> proxies(n_nerfs, std::vector<ProxyRay>(n_init));
	const float* cm = rq->camera;
	for (uint32_t idx = 0; idx < n_init; ++idx) {
		const uint32_t x = (idx % SW) * skip, y = (idx / SW) * skip;
		GlobalRay& g = rays[idx];
		Vec3 d;
		blender_pixel_to_ray(rq, offset, x, y, g.o, d);
		g.d = normalized(d);
		g.idx = idx; g.alive = true;
		g.rgba[0] = g.rgba[1] = g.rgba[2] = g.rgba[3] = 0.f;
		for (uint32_t n = 0; n < n_nerfs; ++n) { // init_proxy_rays_kernel
			ProxyRay& p = proxies[n][idx];
			p = ProxyRay{};
			const Vec3 o = xform_point(IT[n], g.o);
			p.d = normalized(xform_dir(IT[n], normalized(g.d)));
			float tmin, tmax;
			aabb_ray_intersect(render_box[n], o, p.d, &tmin, &tmax);
			const float t = std::fmax(tmin, 0.0f) + 1e-5f;
			if (!aabb_contains(render_box[n], Vec3{o.x + p.d.x * t, o.y + p.d.y * t, o.z + p.d.z * t})) { p.alive = false; continue; }
			bool hits_a_mask = masks[n].empty();
			for (size_t k = 0; k < masks[n].size() && !hits_a_mask; ++k) hits_a_mask = mask_intersects_ray(masks[n][k], o, p.d);
			p.active = true; p.alive = hits_a_mask; p.t = 0.0f; p.n_steps = 0;
			p.o = {o.x + t * p.d.x, o.y + t * p.d.y, o.z + t * p.d.z};
		}
	}
	std::vector<float> frame((size_t)W * H * 4, 0.f);
	const Vec3 cam_pos = {cm[9], cm[10], cm[11]};
	uint64_t n_samples = 0;
	std::vector<uint32_t> alive_list(n_init);
	for (uint32_t i = 0; i < n_init; ++i) alive_list[i] = i;
	uint32_t step = 1;
	while (step < 10000) {
		// compact_rays_kernel: live rays stay, finished rays with alpha > 0.001 are shaded at the end
		std::vector<uint32_t> next;
		next.reserve(alive_list.size());
		for (uint32_t r : alive_list) if (rays[r].alive) next.push_back(r);
		alive_list.swap(next);
		const uint32_t n_alive = (uint32_t)alive_list.size();
		if (n_alive == 0) break;
		// march_active_rays + cull_global_rays_and_set_proxy_rays_active_kernel
		#pragma omp parallel for schedule(dynamic, 64)
		for (int64_t a = 0; a < (int64_t)n_alive; ++a) {
			const uint32_t r = alive_list[a];
			for (uint32_t n = 0; n < n_nerfs; ++n) {
				ProxyRay& p = proxies[n][r];
				if (!p.alive || !p.active) continue;
				const Vec3 idir = {1.0f / p.d.x, 1.0f / p.d.y, 1.0f / p.d.z};
				float t;
				p.alive = hit_test_and_march(p.o, p.d, idir, p.t, render_box[n], nerfs[n].bitfield, cone[n], &t, nullptr);
				p.t = t;
			}
			float min_d2 = 0.0f; int active = -1; uint32_t n_proxy_alive = 0;
			for (uint32_t n = 0; n < n_nerfs; ++n) {
				ProxyRay& p = proxies[n][r];
				if (!p.alive) continue;
				++n_proxy_alive;
				const Vec3 q = xform_point(T[n], Vec3{p.o.x + p.d.x * p.t, p.o.y + p.d.y * p.t, p.o.z + p.d.z * p.t});
				const float dx = q.x - cam_pos.x, dy = q.y - cam_pos.y, dz = q.z - cam_pos.z;
				const float d2 = sum3(dx * dx, dy * dy, dz * dz);
				if (d2 < min_d2 || active == -1) { min_d2 = d2; active = (int)n; }
				p.active = false;
			}
			if (active >= 0) proxies[active][r].active = true;
			if (n_proxy_alive == 0) rays[r].alive = false;
		}
		const uint32_t n_steps = std::max(1u, std::min(8u, n_init / n_alive));
		for (uint32_t n = 0; n < n_nerfs; ++n) {
			// march_proxy_rays_and_generate_next_network_inputs
			std::vector<float> coords((size_t)n_alive * n_steps * 7, 0.f);
			std::vector<uint8_t> takes(n_alive, 0);
			#pragma omp parallel for schedule(dynamic, 64)
			for (int64_t a = 0; a < (int64_t)n_alive; ++a) {
				const uint32_t r = alive_list[a];
				ProxyRay& p = proxies[n][r];
				if (!rays[r].alive || !p.active) continue;
				takes[a] = 1;
				const Vec3 idir = {1.0f / p.d.x, 1.0f / p.d.y, 1.0f / p.d.z};
				float t = p.t, dt = calc_dt(t, cone[n]);
				bool done = false;
				for (uint32_t j = 0; j < n_steps; ++j) {
					const Vec3 pos = {p.o.x + p.d.x * t, p.o.y + p.d.y * t, p.o.z + p.d.z * t};
					const Vec3 wp = warp_position(pos, train_box[n]);
					float* c = &coords[((size_t)a * n_steps + j) * 7];
					c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = (p.d.x + 1.0f) * 0.5f; c[5] = (p.d.y + 1.0f) * 0.5f; c[6] = (p.d.z + 1.0f) * 0.5f;
					if (!hit_test_and_march(p.o, p.d, idir, t, render_box[n], nerfs[n].bitfield, cone[n], &t, &dt)) { p.n_steps = j; done = true; break; }
					t += dt;
				}
				if (!done) { p.t = t; p.n_steps = n_steps; }
			}
			std::vector<orc_half> out((size_t)n_alive * n_steps * 4);
			orc_nerf_inference(nerfs[n].model, nerfs[n].params, n_alive * n_steps, coords.data(), out.data());
			// composite_proxy_ray_colors_kernel
			for (uint32_t a = 0; a < n_alive; ++a) {
				const uint32_t r = alive_list[a];
				ProxyRay& p = proxies[n][r];
				if (!rays[r].alive || !p.alive || !p.active) continue;
				(void)takes;
				float* rgba = rays[r].rgba;
				uint32_t j = 0;
				for (; j < p.n_steps; ++j) {
					const size_t k = (size_t)a * n_steps + j;
					const float T_ = 1.f - rgba[3];
					const float dt = unwarp_dt(coords[k * 7 + 3]);
					const float alpha = 1.f - std::exp(-network_to_density(h2f(out[k * 4 + 3]), nerfs[n].density_activation) * dt);
					float weight = alpha * T_;
					if (!masks[n].empty()) {
						const Vec3 pos = unwarp_position(&coords[k * 7], train_box[n]);
						float mask_weight = 1.f;
						for (const MaskL& mk : masks[n]) mask_weight = std::fmin(std::fmax(mask_weight + mask_sample(mk, pos), 0.0f), 1.0f);
						weight *= mask_weight;
					}
					weight *= nerfs[n].opacity;
					for (int c = 0; c < 3; ++c) rgba[c] += network_to_rgb(h2f(out[k * 4 + c]), nerfs[n].rgb_activation) * weight;
					rgba[3] += weight;
					++n_samples;
					if (rgba[3] > (1.0f - nerfs[n].min_transmittance)) { const float w = rgba[3]; for (int c = 0; c < 4; ++c) rgba[c] /= w; break; }
				}
				if (j < n_steps) { p.alive = false; p.n_steps = j + step; }
			}
		}
		step += n_steps;
	}
	// shade_buffer_with_rays_kernel (train_in_linear_colors = false)
	for (uint32_t idx = 0; idx < n_init; ++idx) {
		const GlobalRay& g = rays[idx];
		if (g.alive || !(g.rgba[3] > 0.001f)) continue;
		const uint32_t x = skip * (g.idx % (uint32_t)SW);
		uint32_t y = skip * (g.idx / (uint32_t)SW);
		if (rq->flip_y) y = H - y - 1;
		float tmp[4] = {srgb_to_linear(g.rgba[0]), srgb_to_linear(g.rgba[1]), srgb_to_linear(g.rgba[2]), g.rgba[3]};
		for (int u = 0; u < skip; ++u) for (int v = 0; v < skip; ++v) {
			const uint64_t pix = (uint64_t)(x + u) + (uint64_t)(y + v) * W;
			if (pix >= (uint64_t)W * H) continue; // (the reference tests only the linear index, :552-554)
			float* f = &frame[pix * 4];
			const float k = 1.0f - tmp[3];
			for (int c = 0; c < 4; ++c) f[c] = tmp[c] + f[c] * k;
		}
	}
	// accumulate (sample count 0) + tonemap, both in the request's colour space
	float bg[4] = {rq->background_color[0], rq->background_color[1], rq->background_color[2], rq->background_color[3]};
	if (rq->color_space != 1) for (int k = 0; k < 3; ++k) bg[k] = srgb_to_linear(bg[k]);
	const float exposure_scale = std::pow(2.0f, rq->exposure);
	for (size_t i = 0; i < (size_t)W * H; ++i) {
		float color[4] = {frame[i * 4], frame[i * 4 + 1], frame[i * 4 + 2], frame[i * 4 + 3]};
		if (rq->color_space == 1) for (int k = 0; k < 3; ++k) color[k] = linear_to_srgb(color[k]);
		for (int k = 0; k < 4; ++k) color[k] = (0.f * 0.f + color[k]) / (0.f + 1);
		const float weight = (1 - color[3]) * bg[3];
		for (int k = 0; k < 3; ++k) color[k] += bg[k] * weight;
		color[3] += weight;
		for (int k = 0; k < 3; ++k) {
			float v = color[k];
			if (rq->color_space == 1) v = srgb_to_linear(v);
			color[k] = v * exposure_scale;
		}
		tonemap_curve(color, rq->tonemap_curve);
		if (rq->color_space == 1) for (int k = 0; k < 3; ++k) color[k] = linear_to_srgb(color[k]); // bl_render_frame passes the same colour space as the output space (:2691)
		for (int k = 0; k < 4; ++k) out_rgba[i * 4 + k] = color[k];
	}
	if (n_samples_out) *n_samples_out = n_samples;
}

// =============================================================================================
// The neural-image / SDF model family (BASELINE configs 1 and 5): N-dimensional hash grid + one FullyFusedMLP 32 -> 64 x n_hidden -> 16.
// =============================================================================================
// GridEncodingTemplated constructor for N_POS_DIMS in {2, 3} (grid.h:985-1018)
extern "C" uint32_t orc_grid_offsets_nd(uint32_t n_dims, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale, uint32_t* offsets) {
	uint32_t offset = 0;
	for (uint32_t i = 0; i < n_levels; ++i) {
		const uint32_t resolution = grid_resolution(grid_scale(i, std::log2(per_level_scale), base_resolution));
		uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
		uint32_t dense = 1;
		for (uint32_t d = 0; d < n_dims; ++d) dense *= resolution;
		uint32_t params_in_level = std::pow((float)resolution, (float)n_dims) > (float)max_params ? max_params : dense;
		params_in_level = (params_in_level + 7u) / 8u * 8u;
		params_in_level = std::min(params_in_level, 1u << log2_hashmap_size);
		offsets[i] = offset;
		offset += params_in_level;
	}
	offsets[n_levels] = offset;
	return offset;
}

// kernel_grid for N_POS_DIMS = n_dims (grid.h:220-349): grid_index with the stride loop's early stop (:164-186), prime_hash over n_dims (:111-128)
extern "C" void orc_grid_forward_nd(uint32_t n_dims, uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale,
                                    const float* scales, const orc_half* grid, const float* positions, uint32_t pos_stride, orc_half* encoded) {
	static const uint32_t primes[3] = {1u, 2654435761u, 805459861u};
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		for (uint32_t level = 0; level < n_levels; ++level) {
			const orc_half* g = grid + (size_t)offsets[level] * 2;
			const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
			const float scale = scales ? scales[level] : grid_scale(level, log2_per_level_scale, base_resolution);
			const uint32_t resolution = grid_resolution(scale);
			float pos[3]; uint32_t pg[3];
			for (uint32_t d = 0; d < n_dims; ++d) pos_fract(positions[(size_t)i * pos_stride + d], &pos[d], &pg[d], scale);
			orc_half result[2] = {0, 0};
			for (uint32_t idx = 0; idx < (1u << n_dims); ++idx) {
				float weight = 1; uint32_t pl[3];
				for (uint32_t d = 0; d < n_dims; ++d) {
					if ((idx & (1u << d)) == 0) { weight *= 1 - pos[d]; pl[d] = pg[d]; } else { weight *= pos[d]; pl[d] = pg[d] + 1; }
				}
				uint32_t stride = 1, index = 0;
				for (uint32_t d = 0; d < n_dims && stride <= hashmap_size; ++d) { index += pl[d] * stride; stride *= resolution; }
				if (hashmap_size < stride) { index = 0; for (uint32_t d = 0; d < n_dims; ++d) index ^= pl[d] * primes[d]; }
				index = (index % hashmap_size) * 2;
				for (int f = 0; f < 2; ++f) result[f] = hadd(result[f], f2h(weight * h2f(g[index + f])));
			}
			encoded[(size_t)i * (2 * n_levels) + level * 2 + 0] = result[0];
			encoded[(size_t)i * (2 * n_levels) + level * 2 + 1] = result[1];
		}
	}
}

// FullyFusedMLP<__half, 64> with in = 32, ReLU hidden layers, no output activation, 16 padded outputs (fully_fused_mlp.cu:500-557 forward,
// :151-314 backward, :805-847 weight gradients). weights: [64][32], (n_hidden - 1) x [64][64], [16][64], row-major [out][in].
// Forward: out[n][16]. With dL_dout[n][16]: also dL_dinput[n][32] (nullable) and grad[n_params] (fp32, unscaled sum over the batch).
// Activations round to fp16 at layer boundaries like the reference; accumulation is fp32 (the reference's wmma accumulates in fp16).
extern "C" void orc_mlp_forward_backward(uint32_t n_hidden, uint32_t n, const orc_half* weights, const orc_half* input, orc_half* out,
                                         const orc_half* dL_dout, orc_half* dL_dinput, float* grad) {
	const uint32_t n_params = 64 * 32 + (n_hidden - 1) * 64 * 64 + 16 * 64;
	std::vector<float> W(n_params);
	for (uint32_t k = 0; k < n_params; ++k) W[k] = h2f(weights[k]);
	const uint32_t w_last = 64 * 32 + (n_hidden - 1) * 64 * 64;
	const int nt = omp_get_max_threads();
	std::vector<std::vector<double>> partial(grad ? nt : 0, std::vector<double>(grad ? n_params : 0, 0.0));
	#pragma omp parallel
	{
		double* G = grad ? partial[omp_get_thread_num()].data() : nullptr;
		std::vector<float> acts((size_t)(n_hidden + 1) * 64), tmp(64), d(64), dprev(64);
		#pragma omp for schedule(static)
		for (int64_t i = 0; i < (int64_t)n; ++i) {
			float* x = acts.data();
			for (int k = 0; k < 32; ++k) x[k] = h2f(input[(size_t)i * 32 + k]);
			const float* w = W.data();
			int n_in = 32;
			for (uint32_t l = 0; l < n_hidden; ++l) {
				float* h = acts.data() + (size_t)(l + 1) * 64;
				matvec(w, 64, n_in, acts.data() + (size_t)l * 64, tmp.data());
				for (int k = 0; k < 64; ++k) h[k] = rh(tmp[k] > 0.f ? tmp[k] : 0.f);
				w += 64 * n_in; n_in = 64;
			}
			const float* last = acts.data() + (size_t)n_hidden * 64;
			matvec(W.data() + w_last, 16, 64, last, tmp.data());
			if (out) for (int k = 0; k < 16; ++k) out[(size_t)i * 16 + k] = f2h(tmp[k]);
			if (!dL_dout) continue;
			float dout[16];
			for (int k = 0; k < 16; ++k) dout[k] = h2f(dL_dout[(size_t)i * 16 + k]);
			if (G) for (int o = 0; o < 16; ++o) for (int k = 0; k < 64; ++k) G[w_last + o * 64 + k] += (double)(dout[o] * last[k]);
			matvec_t(W.data() + w_last, 16, 64, dout, d.data());
			for (int k = 0; k < 64; ++k) d[k] = rh(last[k] > 0.f ? d[k] : 0.f); // backward through the ReLU, fp16 like the reference's fragments
			for (int l = (int)n_hidden - 1; l >= 0; --l) {
				const int in_w = l == 0 ? 32 : 64;
				const uint32_t w_off = l == 0 ? 0u : 64u * 32u + (uint32_t)(l - 1) * 64u * 64u;
				const float* a_in = acts.data() + (size_t)l * 64;
				if (G) for (int o = 0; o < 64; ++o) for (int k = 0; k < in_w; ++k) G[w_off + o * in_w + k] += (double)(d[o] * a_in[k]);
				matvec_t(W.data() + w_off, 64, in_w, d.data(), dprev.data());
				if (l > 0) { for (int k = 0; k < 64; ++k) d[k] = rh(a_in[k] > 0.f ? dprev[k] : 0.f); }
				else if (dL_dinput) { for (int k = 0; k < 32; ++k) dL_dinput[(size_t)i * 32 + k] = f2h(dprev[k]); }
			}
		}
	}
	if (grad) for (uint32_t k = 0; k < n_params; ++k) { double s = 0.0; for (int t = 0; t < nt; ++t) s += partial[t][k]; grad[k] = (float)s; }
}

// tcnn l2_loss (losses/l2.h:40-80) / mape_loss (losses/mape.h:40-80), stride 16 (the network's padded output width), no data pdf. kind: 0 = L2, 1 = MAPE.
extern "C" void orc_loss(int kind, uint32_t n, uint32_t dims, float loss_scale, const orc_half* predictions, const float* targets, float* values, orc_half* gradients) {
	const uint32_t stride = 16, n_elements = n * stride, n_total = n_elements / stride * dims;
	for (uint32_t i = 0; i < n_elements; ++i) {
		const uint32_t intra = i % stride, inter = i / stride;
		if (intra >= dims) { if (values) values[i] = 0.f; gradients[i] = f2h(0.f); continue; }
		const float prediction = h2f(predictions[i]);
		const float target = targets[inter * dims + intra];
		const float difference = prediction - target;
		float value, gradient;
		if (kind == 1) {
			const float scale = 1.0f / (std::fabs(target) + 1e-2f);
			value = std::fabs(difference) * scale / n_total;
			gradient = std::copysign(scale, difference);
		} else {
			value = difference * difference / n_total;
			gradient = 2 * difference;
		}
		if (values) values[i] = value;
		gradients[i] = f2h(loss_scale * gradient / n_total);
	}
}

// =============================================================================================
// Input gradients (camera-extrinsics optimisation, K13/K14). Pinned on the reference's own kernels (kernel_grid dy_dx + kernel_grid_backward_input,
// kernel_sh_backward, compute_cam_gradient_train_nerf run on a B200: tests/golden/ref_camera.npz, tests/test_camera_optimizer.py); tests/test_oracle_cpu.py
// additionally checks them against finite differences of the oracle's own forward paths.
// =============================================================================================

// dL/dx of the hash-grid encoding: kernel_grid's dy_dx (grid.h:351-392: per level and feature, for each input dimension the difference of the two
// neighbours along that dimension, blended linearly over the other dimensions, times the level scale) contracted with dL/dy as in
// kernel_grid_backward_input (grid.h:546-575). positions [n][stride], dL_dy [n][2 * n_levels] fp16, dL_dx [n][3] fp32.
extern "C" void orc_grid_input_gradient(uint32_t n, uint32_t n_levels, const uint32_t* offsets, uint32_t base_resolution, float log2_per_level_scale, const float* scales,
                                        const orc_half* grid, const float* positions, uint32_t pos_stride, const orc_half* dL_dy, float* dL_dx) {
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) {
		float result[3] = {0.f, 0.f, 0.f};
		for (uint32_t level = 0; level < n_levels; ++level) {
			const orc_half* g = grid + (size_t)offsets[level] * 2;
			const uint32_t hashmap_size = offsets[level + 1] - offsets[level];
			const float scale = scales ? scales[level] : grid_scale(level, log2_per_level_scale, base_resolution);
			const uint32_t resolution = grid_resolution(scale);
			float pos[3]; uint32_t pg[3];
			for (int d = 0; d < 3; ++d) pos_fract(positions[(size_t)i * pos_stride + d], &pos[d], &pg[d], scale);
			float grads[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
			for (uint32_t grad_dim = 0; grad_dim < 3; ++grad_dim) {
				for (uint32_t idx = 0; idx < 4; ++idx) {
					float weight = scale;
					uint32_t pl[3];
					for (uint32_t ng = 0; ng < 2; ++ng) {
						const uint32_t dim = ng >= grad_dim ? ng + 1 : ng;
						if ((idx & (1u << ng)) == 0) { weight *= 1 - pos[dim]; pl[dim] = pg[dim]; } else { weight *= pos[dim]; pl[dim] = pg[dim] + 1; }
					}
					pl[grad_dim] = pg[grad_dim];
					const uint32_t left = grid_index(hashmap_size, resolution, pl) * 2;
					pl[grad_dim] = pg[grad_dim] + 1;
					const uint32_t right = grid_index(hashmap_size, resolution, pl) * 2;
					for (int f = 0; f < 2; ++f) grads[f][grad_dim] += weight * (h2f(g[right + f]) - h2f(g[left + f])); // pos_derivative = 1 (linear interpolation)
				}
			}
			for (int f = 0; f < 2; ++f) {
				const float dl = h2f(dL_dy[(size_t)i * (2 * n_levels) + level * 2 + f]);
				for (int d = 0; d < 3; ++d) result[d] += dl * grads[f][d];
			}
		}
		for (int d = 0; d < 3; ++d) dL_dx[(size_t)i * 3 + d] = result[d];
	}
}

// dL/d(direction input) of the degree-4 spherical harmonics (kernel_sh_backward, spherical_harmonics.h:154-390): the polynomials of sh4() above
// differentiated term by term in x = 2 d - 1, times 2 for the [0,1] -> [-1,1] mapping. dirs [n][stride], dL_dy [n][16] fp16, dL_dx [n][3] fp32.
extern "C" void orc_sh4_input_gradient(uint32_t n, const float* dirs, uint32_t stride, const orc_half* dL_dy, float* dL_dx) {
	for (uint32_t i = 0; i < n; ++i) {
		const float x = dirs[(size_t)i * stride + 0] * 2.f - 1.f, y = dirs[(size_t)i * stride + 1] * 2.f - 1.f, z = dirs[(size_t)i * stride + 2] * 2.f - 1.f;
		const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
		float g[16];
		for (int k = 0; k < 16; ++k) g[k] = h2f(dL_dy[(size_t)i * 16 + k]);
		// Y1 = -c1 y, Y2 = c1 z, Y3 = -c1 x
		const float c1 = 0.48860251190291987f, c4 = 1.0925484305920792f, c6 = 0.94617469575755997f, c8 = 0.54627421529603959f;
		const float c9 = 0.59004358992664352f, c10 = 2.8906114426405538f, c11 = 0.45704579946446572f, c12 = 0.3731763325901154f, c14 = 1.4453057213202769f;
		float dx = 0.f, dy = 0.f, dz = 0.f;
		dy += g[1] * -c1; dz += g[2] * c1; dx += g[3] * -c1;
		// Y4 = c4 xy, Y5 = -c4 yz, Y6 = c6 z2 - const, Y7 = -c4 xz, Y8 = c8 (x2 - y2)
		dx += g[4] * (c4 * y); dy += g[4] * (c4 * x);
		dy += g[5] * (-c4 * z); dz += g[5] * (-c4 * y);
		dz += g[6] * (2.f * c6 * z);
		dx += g[7] * (-c4 * z); dz += g[7] * (-c4 * x);
		dx += g[8] * (2.f * c8 * x); dy += g[8] * (-2.f * c8 * y);
		// Y9 = c9 y (y2 - 3 x2), Y10 = c10 xyz, Y11 = c11 y (1 - 5 z2), Y12 = c12 z (5 z2 - 3), Y13 = c11 x (1 - 5 z2), Y14 = c14 z (x2 - y2), Y15 = c9 x (3 y2 - x2)
		dx += g[9] * (-6.f * c9 * xy); dy += g[9] * (3.f * c9 * (y2 - x2));
		dx += g[10] * (c10 * yz); dy += g[10] * (c10 * xz); dz += g[10] * (c10 * xy);
		dy += g[11] * (c11 * (1.f - 5.f * z2)); dz += g[11] * (-10.f * c11 * yz);
		dz += g[12] * (c12 * (15.f * z2 - 3.f));
		dx += g[13] * (c11 * (1.f - 5.f * z2)); dz += g[13] * (-10.f * c11 * xz);
		dx += g[14] * (2.f * c14 * xz); dy += g[14] * (-2.f * c14 * yz); dz += g[14] * (c14 * (x2 - y2));
		dx += g[15] * (3.f * c9 * (y2 - x2)); dy += g[15] * (6.f * c9 * xy);
		dL_dx[(size_t)i * 3 + 0] = dx * 2.0f; dL_dx[(size_t)i * 3 + 1] = dy * 2.0f; dL_dx[(size_t)i * 3 + 2] = dz * 2.0f;
	}
}

// dL/d(network input) of the NeRF model for a batch (NerfNetwork::backward with dL_dinput, nerf_network.h:187-266): position gradient through the density
// network and the hash grid, direction gradient through the rgb network's SH inputs. dL_dcoords [n][7] fp32: {pos 3, dt (no gradient), dir 3}.
extern "C" void orc_nerf_input_gradient(const orc_model* m, const orc_half* params, uint32_t n, const float* coords, const orc_half* dL_dout, float* dL_dcoords) {
	std::vector<orc_half> enc((size_t)n * 32), denc((size_t)n * 32), dsh((size_t)n * 16);
	const float l2 = std::log2(m->per_level_scale);
	orc_grid_forward(n, m->n_levels, m->offsets, m->base_resolution, l2, m->scales, params + ORC_MLP_PARAMS, coords, 7, enc.data());
	MlpWeightsF Wf(params);
	const float* W = Wf.w.data();
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n; ++i) { // the data-gradient half of orc_nerf_mlp_backward, keeping the SH part of the rgb network's input gradient
		SampleActs a;
		mlp_forward_one(W, enc.data() + (size_t)i * 32, coords + (size_t)i * 7, a);
		float d_orgb[16] = {0}, tmp[64], d_g2[64], d_g1[64], d_rin[32], d_od[16], d_h1[64], d_x[32];
		for (int c = 0; c < 3; ++c) d_orgb[c] = h2f(dL_dout[(size_t)i * 4 + c]);
		matvec_t(W + W3R, 16, 64, d_orgb, tmp);
		for (int k = 0; k < 64; ++k) d_g2[k] = rh(a.g2[k] > 0.f ? tmp[k] : 0.f);
		matvec_t(W + W2R, 64, 64, d_g2, tmp);
		for (int k = 0; k < 64; ++k) d_g1[k] = rh(a.g1[k] > 0.f ? tmp[k] : 0.f);
		matvec_t(W + W1R, 64, 32, d_g1, tmp);
		for (int k = 0; k < 32; ++k) d_rin[k] = rh(tmp[k]);
		for (int k = 0; k < 16; ++k) { d_od[k] = d_rin[k]; dsh[(size_t)i * 16 + k] = f2h(d_rin[16 + k]); }
		d_od[0] = rh(d_od[0] + h2f(dL_dout[(size_t)i * 4 + 3]));
		matvec_t(W + W2D, 16, 64, d_od, tmp);
		for (int k = 0; k < 64; ++k) d_h1[k] = rh(a.h1[k] > 0.f ? tmp[k] : 0.f);
		matvec_t(W + W1D, 64, 32, d_h1, d_x);
		for (int k = 0; k < 32; ++k) denc[(size_t)i * 32 + k] = f2h(d_x[k]);
	}
	std::vector<float> dpos((size_t)n * 3), ddir((size_t)n * 3);
	orc_grid_input_gradient(n, m->n_levels, m->offsets, m->base_resolution, l2, m->scales, params + ORC_MLP_PARAMS, coords, 7, denc.data(), dpos.data());
	orc_sh4_input_gradient(n, coords + 4, 7, dsh.data(), ddir.data());
	for (uint32_t i = 0; i < n; ++i) {
		float* o = dL_dcoords + (size_t)i * 7;
		o[0] = dpos[(size_t)i * 3]; o[1] = dpos[(size_t)i * 3 + 1]; o[2] = dpos[(size_t)i * 3 + 2]; o[3] = 0.f;
		o[4] = ddir[(size_t)i * 3]; o[5] = ddir[(size_t)i * 3 + 1]; o[6] = ddir[(size_t)i * 3 + 2];
	}
}

// compute_cam_gradient_train_nerf (src/testbed_nerf.cu:1600-1707) without distortion / focal-length terms and with uniform pixel sampling (xy_pdf = 1):
// per kept ray, the sum of its samples' position gradients is the gradient of the camera origin; the direction gradient (position gradients scaled by
// the distance along the ray, plus the direction-input gradients) gives the rotation gradient as ray.d x ray_gradient.d. Sums per image (the kernel uses
// atomicAdd; here doubles). rays_unnormalized [n][6]; cam_pos_gradient, cam_rot_gradient [n_images][3].
extern "C" void orc_compute_cam_gradient(uint32_t n_kept, uint32_t n_rays_total, uint32_t n_images, const float* aabb6, const uint32_t* ray_indices,
                                         const float* rays_unnormalized, const uint32_t* numsteps, const float* coords, const float* coords_gradient,
                                         float* cam_pos_gradient, float* cam_rot_gradient) {
	const AABB aabb = make_aabb(aabb6);
	std::vector<double> pos_acc((size_t)n_images * 3, 0.0), rot_acc((size_t)n_images * 3, 0.0);
	const Vec3 diag = {aabb.max.x - aabb.min.x, aabb.max.y - aabb.min.y, aabb.max.z - aabb.min.z};
	for (uint32_t i = 0; i < n_kept; ++i) {
		const uint32_t ns = numsteps[i * 2 + 0], base = numsteps[i * 2 + 1];
		if (ns == 0) continue;
		const uint32_t img = image_idx(ray_indices[i], n_rays_total, n_images);
		const float* r = rays_unnormalized + (size_t)i * 6;
		const Vec3 ro = {r[0], r[1], r[2]};
		const Vec3 rd = normalized(Vec3{r[3], r[4], r[5]});
		Vec3 go = {0.f, 0.f, 0.f}, gd = {0.f, 0.f, 0.f};
		for (uint32_t j = 0; j < ns; ++j) {
			const float* c = coords + (size_t)(base + j) * 7;
			const float* gc = coords_gradient + (size_t)(base + j) * 7;
			const Vec3 pg = {gc[0] * (1.0f / diag.x), gc[1] * (1.0f / diag.y), gc[2] * (1.0f / diag.z)}; // warp_position_derivative = 1 / aabb.diag()
			go = {go.x + pg.x, go.y + pg.y, go.z + pg.z};
			const Vec3 pos = unwarp_position(c, aabb);
			const float dx = pos.x - ro.x, dy = pos.y - ro.y, dz = pos.z - ro.z;
			const float t = std::sqrt(dx * dx + dy * dy + dz * dz);
			gd = {gd.x + pg.x * t + gc[4] * 0.5f, gd.y + pg.y * t + gc[5] * 0.5f, gd.z + pg.z * t + gc[6] * 0.5f}; // warp_direction_derivative = 0.5
		}
		const Vec3 axis = {rd.y * gd.z - rd.z * gd.y, rd.z * gd.x - rd.x * gd.z, rd.x * gd.y - rd.y * gd.x};
		pos_acc[(size_t)img * 3 + 0] += go.x; pos_acc[(size_t)img * 3 + 1] += go.y; pos_acc[(size_t)img * 3 + 2] += go.z;
		rot_acc[(size_t)img * 3 + 0] += axis.x; rot_acc[(size_t)img * 3 + 1] += axis.y; rot_acc[(size_t)img * 3 + 2] += axis.z;
	}
	for (size_t k = 0; k < pos_acc.size(); ++k) { cam_pos_gradient[k] = (float)pos_acc[k]; cam_rot_gradient[k] = (float)rot_acc[k]; }
}

// Number of OpenMP threads the restatement uses (bench.py's CPU arm sets it to the host's core count: torchrun exports OMP_NUM_THREADS=1).
extern "C" int orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
