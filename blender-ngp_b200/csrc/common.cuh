// Shared device/host helpers for the B200 NeRF hot path. sm_100a only.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <stdexcept>
#include <string>

namespace ngpb {

// ---- error handling: int status at the C boundary, exceptions inside the C++ host -------------
void set_last_error(const std::string& msg);

#define NGPB_CUDA_CHECK(expr)                                                                       \
	do {                                                                                            \
		cudaError_t _e = (expr);                                                                    \
		if (_e != cudaSuccess) {                                                                    \
			throw std::runtime_error(std::string(#expr " failed: ") + cudaGetErrorString(_e));      \
		}                                                                                           \
	} while (0)

// Launch-error check that does not synchronise.
#define NGPB_LAUNCH_CHECK() NGPB_CUDA_CHECK(cudaGetLastError())

constexpr uint32_t kNumSMs = 148;

// Every kernel of a training step asks for the same L1 / shared-memory split. An SM can only change its split while it is idle,
// so a kernel with a different requirement (the MLP training kernel needs 143 KB of shared memory) would otherwise wait for the
// long-running ray-marching kernel that overlaps it on the second stream to drain first (measured: 130 us per step).
constexpr int kSmemCarveoutPercent = 72; // 164 KB shared + ~64 KB L1
inline int step_carveout() { static const int v = getenv("NGPB_CARVEOUT") ? atoi(getenv("NGPB_CARVEOUT")) : -1; return v; }
#define NGPB_STEP_KERNEL(kernel)                                                                                           \
	do {                                                                                                                   \
		static bool _configured = false;                                                                                   \
		if (!_configured) {                                                                                                \
			if (step_carveout() >= 0) NGPB_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, step_carveout())); \
			_configured = true;                                                                                            \
		}                                                                                                                  \
	} while (0)

inline uint32_t div_round_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline uint32_t next_multiple(uint32_t a, uint32_t b) { return div_round_up(a, b) * b; }

// ---- constants of the NeRF path (reference: src/testbed_nerf.cu:53-73, nerf.h:24) -------------
constexpr uint32_t NERF_GRIDSIZE = 128;
constexpr uint32_t NERF_GRID_CELLS = NERF_GRIDSIZE * NERF_GRIDSIZE * NERF_GRIDSIZE;
constexpr uint32_t NERF_STEPS = 1024;
constexpr uint32_t NERF_CASCADES = 8;
constexpr uint32_t N_MAX_RANDOM_SAMPLES_PER_RAY = 8;
constexpr float SQRT3 = 1.73205080757f;
constexpr float STEPSIZE = SQRT3 / NERF_STEPS;
constexpr float MIN_CONE_STEPSIZE = STEPSIZE;
constexpr float MAX_CONE_STEPSIZE = STEPSIZE * (1 << (NERF_CASCADES - 1)) * NERF_STEPS / NERF_GRIDSIZE;
constexpr float NERF_MIN_OPTICAL_THICKNESS = 0.01f;
constexpr float LOSS_SCALE = 128.0f; // testbed.h:272

// Network dimensions (configs/nerf/base.json): hash grid 16 levels x 2 features; density MLP
// 32->64->16, rgb MLP 32->64->64->16. Flat parameter order follows nerf_network.h:361-394.
constexpr uint32_t N_ENC = 32;
constexpr uint32_t MLP_W1D = 0, MLP_W2D = 2048, MLP_W1R = 3072, MLP_W2R = 5120, MLP_W3R = 9216, MLP_PARAMS = 10240;
constexpr uint32_t COORD_FLOATS = 7; // NerfCoordinate {pos[3], dt, dir[3]}, nerf.h:81-107

// ---- PCG32 (tcnn/dependencies/pcg32/pcg32.h), device + host ----------------------------------
// K19: the error-map CDFs the training kernels draw pixels / images from (see nerf_device.cuh); null members = uniform.
struct ErrorCdf {
	const float* x_cond_y; // [n_images][res_y][res_x]: per row, the CDF over the columns
	const float* y;        // [n_images][res_y]: CDF over the rows
	const float* img;      // [n_images]: CDF over the images
	int res_x, res_y;
};
__host__ __device__ inline ErrorCdf no_error_cdf() { return ErrorCdf{nullptr, nullptr, nullptr, 0, 0}; }

struct Pcg32 {
	uint64_t state, inc;
	__host__ __device__ uint32_t next_uint() {
		uint64_t oldstate = state;
		state = oldstate * 0x5851f42d4c957f2dULL + inc;
		uint32_t xorshifted = (uint32_t)(((oldstate >> 18u) ^ oldstate) >> 27u);
		uint32_t rot = (uint32_t)(oldstate >> 59u);
		return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
	}
	__host__ __device__ float next_float() {
		uint32_t u = (next_uint() >> 9) | 0x3f800000u;
#ifdef __CUDA_ARCH__
		return __uint_as_float(u) - 1.0f;
#else
		float f; memcpy(&f, &u, 4); return f - 1.0f;
#endif
	}
	__host__ __device__ void advance(int64_t delta_ = (1ll << 32)) {
		uint64_t cur_mult = 0x5851f42d4c957f2dULL, cur_plus = inc, acc_mult = 1u, acc_plus = 0u;
		uint64_t delta = (uint64_t)delta_;
		while (delta > 0) {
			if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
			cur_plus = (cur_mult + 1) * cur_plus;
			cur_mult *= cur_mult;
			delta /= 2;
		}
		state = acc_mult * state + acc_plus;
	}
	__host__ __device__ void seed(uint64_t initstate, uint64_t initseq = 1) {
		state = 0U; inc = (initseq << 1u) | 1u;
		next_uint(); state += initstate; next_uint();
	}
};

// ---- morton (tcnn common_device.h:338-362) ----------------------------------------------------
__host__ __device__ inline uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
__host__ __device__ inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
	return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ inline uint32_t morton3D_invert(uint32_t x) {
	x = x & 0x49249249;
	x = (x | (x >> 2)) & 0xc30c30c3;
	x = (x | (x >> 4)) & 0x0f00f00f;
	x = (x | (x >> 8)) & 0xff0000ff;
	x = (x | (x >> 16)) & 0x0000ffff;
	return x;
}

struct Aabb { float min[3], max[3]; };

} // namespace ngpb
