// The neural-image and SDF modes of the Testbed around the same kernels as the NeRF path: a hash-grid encoding (2-D or 3-D) feeding the 32 -> 64 -> 64 -> 16
// fully fused MLP, an element-wise loss, Adam (optionally wrapped in an EMA), and the mode's own training-batch generation.
// Replaces (reference): tcnn NetworkWithInputEncoding + Trainer::training_step / optimizer_step (dependencies/tiny-cuda-nn/include/tiny-cuda-nn/trainer.h:108-190),
// Testbed::reset_network for these modes (src/testbed.cu:2244-2470), Testbed::train_image / render_image / compute_image_mse (src/testbed_image.cu:220-523),
// Testbed::train_sdf (src/testbed_sdf.cu:1229-1252) on supplied (position, distance) pairs (override_sdf_training_data, src/python_api.cu:74-104).
// Mesh loading, BVH distance queries and SDF sphere tracing are outside the path.
#include "common.cuh"
#include "nerf_device.cuh"
#include "../../include/ngpb.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

namespace ngpb {

void hash_encode_forward_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* positions, uint32_t pos_stride, uint32_t n, const uint32_t* n_dev, __half* encoded, bool tiled);
void hash_encode_backward_launch(cudaStream_t stream, const ngpb_grid* g, const float* positions, uint32_t pos_stride, uint32_t n, const __half* dL_dencoded, float* grid_grad, uint32_t level_begin, uint32_t level_end);
void plain_mlp_launch(cudaStream_t stream, const __half* weights, const __half* input, uint32_t n, __half* output);
void plain_mlp_forward_backward_launch(cudaStream_t stream, const __half* weights, const __half* input, const __half* dL_dout16, uint32_t n, __half* dL_dinput, float* grad, float* partials);
void optimizer_prepare(ngpb_optimizer* o, float loss_scale, void* params_out);
void optimizer_disable_fused_ema(void* params);
void optimizer_launch(cudaStream_t stream, const void* params, uint32_t first, uint32_t count, uint32_t n_matrix_params, float* grad, float* w_fp32, __half* w_half, __half* w_ema, float* m1, float* m2, uint32_t* param_steps);
void render_accumulate_launch(cudaStream_t stream, uint32_t n_pixels, const float* frame_rgba, float* accumulate_rgba, float sample_count, int color_space);
void render_tonemap_launch(cudaStream_t stream, uint32_t n_pixels, float exposure, const float* background4, const float* accumulate_rgba, int color_space, int output_srgb, int curve, float* out_rgba);
void ld_random_pixel_offset_host(uint32_t spp, float* out2);

constexpr uint32_t NET_PARAMS = 64 * 32 + 64 * 64 + 16 * 64; // 7168
constexpr uint32_t OUT_STRIDE = 16;                          // padded output width of the fully fused MLP

// ---- generic kernels ----------------------------------------------------------------------------------------------------------------------------------
// tcnn generate_random_uniform (random.h:66-97): thread i advances the stream by 4 i and writes elements i + k * n_threads
__global__ void __launch_bounds__(128) random_uniform_kernel(const uint64_t n_elements, const uint64_t n_threads, Pcg32 rng, float* __restrict__ out, const float lower, const float upper)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_threads) return;
	rng.advance((int64_t)(i * 4));
	#pragma unroll
	for (uint64_t j = 0; j < 4; ++j) {
		const uint64_t idx = i + n_threads * j;
		if (idx >= n_elements) return;
		out[idx] = __fmaf_rn(rng.next_float(), upper - lower, lower); // the reference's lambda is compiled with FMA contraction
	}
}
static void random_uniform(cudaStream_t stream, Pcg32& rng, uint64_t n_elements, float* out, float lower, float upper) {
	const uint64_t n_threads = next_multiple((uint32_t)((n_elements + 3) / 4), 128u);
	random_uniform_kernel<<<(uint32_t)(n_threads / 128), 128, 0, stream>>>(n_elements, n_threads, rng, out, lower, upper);
	NGPB_LAUNCH_CHECK();
	rng.advance((int64_t)n_elements);
}

__global__ void __launch_bounds__(256) model_cast_params_kernel(const uint32_t n, const float* __restrict__ w_fp32, __half* __restrict__ w_half)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) w_half[i] = __float2half_rn(w_fp32[i]);
}
__global__ void __launch_bounds__(256) model_widen_params_kernel(const uint32_t n, const __half* __restrict__ w_half, float* __restrict__ w_fp32)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) w_fp32[i] = __half2float(w_half[i]);
}

// network output [n][16] fp16 -> [n][dims] fp32 (tcnn Network::inference into a float matrix, network.h:60-90)
__global__ void __launch_bounds__(256) output_to_float_kernel(const uint32_t n, const uint32_t dims, const __half* __restrict__ out16, float* __restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n * dims) return;
	out[i] = __half2float(out16[(size_t)(i / dims) * OUT_STRIDE + i % dims]);
}

// sum of n floats in double, one block, fixed order (the loss scalar: tcnn reduce_sum over Trainer::ForwardContext::L)
__global__ void __launch_bounds__(1024) model_sum_kernel(const float* __restrict__ v, const uint32_t n, const float scale, float* __restrict__ out)
{
	__shared__ double sm[1024];
	double s = 0.0;
	for (uint32_t i = threadIdx.x; i < n; i += 1024) s += (double)v[i];
	sm[threadIdx.x] = s;
	__syncthreads();
	for (uint32_t o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) *out = (float)(sm[0] * (double)scale);
}

// ---- neural image: training batch (src/testbed_image.cu:63-80,:174-218) --------------------------------------------------------------------------------
// stratify2_kernel: the batch is a sqrt(B) x sqrt(B) grid of cells, each holding one uniform sample
__global__ void __launch_bounds__(256) stratify2_kernel(const uint32_t n_elements, const uint32_t log2_batch_size, float2* __restrict__ inout)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t log2_size = log2_batch_size / 2, size = 1u << log2_size;
	const uint32_t in_batch_index = i & ((1u << log2_batch_size) - 1);
	const uint32_t x = in_batch_index & ((1u << log2_size) - 1), y = in_batch_index >> log2_size;
	const float2 v = inout[i];
	inout[i] = make_float2(v.x / size + ((float)x / size), v.y / size + ((float)y / size));
}

template <typename T> __device__ __forceinline__ float4 read_texel(const T* texture, int idx);
template <> __device__ __forceinline__ float4 read_texel<float>(const float* texture, int idx) { return reinterpret_cast<const float4*>(texture)[idx]; }
template <> __device__ __forceinline__ float4 read_texel<__half>(const __half* texture, int idx) {
	const uint2 raw = reinterpret_cast<const uint2*>(texture)[idx];
	const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
	return make_float4(a.x, a.y, b.x, b.y);
}

// eval_image_kernel_and_snap<T, 3>: the image's value at a position (nearest texel when snapping, which also moves the position onto the texel centre;
// bilinear otherwise), converted to sRGB unless the model trains in linear colours. result: [n][3]
template <typename T>
__global__ void __launch_bounds__(256) eval_image_and_snap_kernel(const uint32_t n_elements, const T* __restrict__ texture, float2* __restrict__ positions, const int res_x, const int res_y,
                                                                  float* __restrict__ result, const bool snap_to_pixel_centers, const bool linear_colors)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	float2 pos = positions[i];
	auto read_val = [&](int x, int y) {
		float4 v = read_texel<T>(texture, y * res_x + x);
		if (!linear_colors) { v.x = linear_to_srgb(v.x); v.y = linear_to_srgb(v.y); v.z = linear_to_srgb(v.z); }
		return v;
	};
	float4 val;
	if (snap_to_pixel_centers) {
		int px = (int)floorf(pos.x * (float)res_x), py = (int)floorf(pos.y * (float)res_y);
		positions[i] = make_float2(((float)px + 0.5f) / (float)res_x, ((float)py + 0.5f) / (float)res_y);
		px = min(max(px, 0), res_x - 1); py = min(max(py, 0), res_y - 1);
		val = read_val(px, py);
	} else {
		pos.x = fminf(fmaxf(pos.x * (float)res_x - 0.5f, 0.0f), (float)res_x - (1.0f + 1e-4f));
		pos.y = fminf(fmaxf(pos.y * (float)res_y - 0.5f, 0.0f), (float)res_y - (1.0f + 1e-4f));
		const int px = (int)pos.x, py = (int)pos.y;
		const float wx = pos.x - (float)px, wy = pos.y - (float)py;
		const int ix = max(min(px, res_x - 2), 0), iy = max(min(py, res_y - 2), 0);
		const float4 a = read_val(ix, iy), b = read_val(ix + 1, iy), c = read_val(ix, iy + 1), d = read_val(ix + 1, iy + 1);
		const float w00 = (1 - wx) * (1 - wy), w10 = wx * (1 - wy), w01 = (1 - wx) * wy, w11 = wx * wy;
		val.x = ((w00 * a.x + w10 * b.x) + w01 * c.x) + w11 * d.x;
		val.y = ((w00 * a.y + w10 * b.y) + w01 * c.y) + w11 * d.y;
		val.z = ((w00 * a.z + w10 * b.z) + w01 * c.z) + w11 * d.z;
	}
	result[(size_t)i * 3 + 0] = val.x; result[(size_t)i * 3 + 1] = val.y; result[(size_t)i * 3 + 2] = val.z;
}

// image_coords_from_idx (:436-446): texel centres in scan order
__global__ void __launch_bounds__(256) image_coords_from_idx_kernel(const uint32_t n_elements, const uint32_t offset, float2* __restrict__ pos, const int res_x, const int res_y)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t idx = i + offset;
	const int x = min(max((int)(idx % (uint32_t)res_x), 0), res_x - 1), y = min(max((int)(idx / (uint32_t)res_x), 0), res_y - 1);
	pos[i] = make_float2(((float)x + 0.5f) / (float)res_x, ((float)y + 0.5f) / (float)res_y);
}

// image_mse_kernel (:448-459)
__global__ void __launch_bounds__(256) image_mse_kernel(const uint32_t n_elements, const float* __restrict__ target, const float* __restrict__ prediction, float* __restrict__ result, const bool quantize_to_byte)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	float se[3];
	#pragma unroll
	for (int c = 0; c < 3; ++c) {
		float p = prediction[(size_t)i * 3 + c];
		if (quantize_to_byte) p = (float)min(max((int)(p * 255.0f + 0.5f), 0), 255) / 255.0f;
		const float d = target[(size_t)i * 3 + c] - p;
		se[c] = d * d;
	}
	result[i] = sum3(se[0], se[1], se[2]) / 3.0f;
}

// init_image_coords + pixel_to_image_uv (:82-110, common_device.cuh:397-417)
__global__ void __launch_bounds__(256) init_image_coords_kernel(float2* __restrict__ positions, const int res_x, const int res_y, const int img_x, const int img_y, const float view_dist,
                                                                const float image_pos_x, const float image_pos_y, const float center_x, const float center_y, const float jit_x, const float jit_y)
{
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (uint32_t)res_x * (uint32_t)res_y) return;
	const int x = (int)(idx % (uint32_t)res_x), y = (int)(idx / (uint32_t)res_x);
	const float off_x = center_x * (float)res_x + jit_x, off_y = center_y * (float)res_y + jit_y;
	const float y_scale = view_dist, x_scale = y_scale * (float)res_x / (float)res_y;
	positions[idx] = make_float2(((x_scale * ((float)x + off_x)) / (float)res_x - view_dist * image_pos_x) / (float)img_x * (float)img_y,
	                             (y_scale * ((float)y + off_y)) / (float)res_y - view_dist * image_pos_y);
}

// shade_kernel_image (:132-172): outside the image transparent black, inside the network's colour (to linear unless trained in linear colours), alpha 1
__global__ void __launch_bounds__(256) shade_image_kernel(const uint32_t n_pixels, const float2* __restrict__ positions, const float* __restrict__ colors, float4* __restrict__ frame_buffer, const bool linear_colors)
{
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n_pixels) return;
	const float2 uv = positions[idx];
	if (uv.x < 0.0f || uv.x > 1.0f || uv.y < 0.0f || uv.y > 1.0f) { frame_buffer[idx] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
	float c[3] = {colors[(size_t)idx * 3], colors[(size_t)idx * 3 + 1], colors[(size_t)idx * 3 + 2]};
	if (!linear_colors) { c[0] = srgb_to_linear(c[0]); c[1] = srgb_to_linear(c[1]); c[2] = srgb_to_linear(c[2]); }
	frame_buffer[idx] = make_float4(c[0], c[1], c[2], 1.0f);
}

// from_rgba32<float> (common_device.cuh:562-590) without the NSVF / mask options: RGBA8 -> linear, premultiplied float
__global__ void __launch_bounds__(256) from_rgba8_kernel(const uint32_t n_pixels, const uchar4* __restrict__ pixels, float4* __restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	const uchar4 p = pixels[i];
	const float alpha = p.w * (1.0f / 255.0f);
	out[i] = make_float4(srgb_to_linear(p.x * (1.0f / 255.0f)) * alpha, srgb_to_linear(p.y * (1.0f / 255.0f)) * alpha, srgb_to_linear(p.z * (1.0f / 255.0f)) * alpha, alpha);
}

// ---- SDF: tcnn shuffle (common_device.h:499-514): out[i] = in[permute(i + seed)] ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) shuffle_kernel(const uint32_t n_elements, const uint32_t stride, const uint32_t seed, const float* __restrict__ in, float* __restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements * stride) return;
	const uint32_t elem_id = i / stride, member_id = i % stride;
	const uint32_t src = (uint32_t)(((uint64_t)(elem_id + seed) * 1434869437ull + 2097192037ull) % n_elements);
	out[i] = in[(size_t)src * stride + member_id];
}

} // namespace ngpb

using namespace ngpb;

struct ngpb_model {
	int device = 0;
	cudaStream_t stream = nullptr;
	ngpb_model_config cfg{};
	ngpb_grid grid{};
	uint32_t n_params = 0, n_alloc = 0;
	float* w_fp32 = nullptr; __half* w_half = nullptr; __half* w_ema = nullptr;
	float* m1 = nullptr; float* m2 = nullptr; uint32_t* param_steps = nullptr; float* grad = nullptr;
	ngpb_optimizer opt{};
	Pcg32 rng{};
	uint32_t seed = 1337, training_step = 0;
	float loss_scalar = 0.f;
	// workspace, sized by the largest batch seen
	uint32_t ws_n = 0;
	__half* enc = nullptr; __half* out16 = nullptr; __half* dout = nullptr; __half* denc = nullptr; float* values = nullptr; float* partials = nullptr;
	float* loss_dev = nullptr; float* loss_host = nullptr;
	float* positions = nullptr; float* targets = nullptr; // the mode's training batch
	// image mode
	void* image = nullptr; int img_w = 0, img_h = 0, img_half = 0;
	int snap_to_pixel_centers = 1, linear_colors = 0;
	float* render_ws = nullptr; size_t render_ws_floats = 0;
	// sdf mode
	float* sdf_pos = nullptr; float* sdf_dist = nullptr; float* sdf_pos_shuffled = nullptr; float* sdf_dist_shuffled = nullptr; uint32_t sdf_size = 0;
	uint64_t n_launches = 0;

	std::vector<void*> allocations;
	void* dalloc(size_t bytes) { void* p = nullptr; NGPB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(bytes, 16))); allocations.push_back(p); return p; }
	void dfree(void* p) { if (!p) return; auto it = std::find(allocations.begin(), allocations.end(), p); if (it != allocations.end()) allocations.erase(it); cudaFree(p); }

	uint32_t n_in() const { return cfg.n_pos_dims; }
	bool use_ema() const { return cfg.use_ema != 0; }
	const __half* inference_params(bool use_inference_params) const { return use_inference_params && use_ema() ? w_ema : w_half; }

	~ngpb_model() {
		cudaSetDevice(device);
		if (stream) cudaStreamSynchronize(stream);
		for (void* p : allocations) cudaFree(p);
		if (loss_host) cudaFreeHost(loss_host);
		if (stream) cudaStreamDestroy(stream);
	}

	void ensure_workspace(uint32_t n) {
		if (n <= ws_n) return;
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		dfree(enc); dfree(out16); dfree(dout); dfree(denc); dfree(values); dfree(positions); dfree(targets);
		enc = (__half*)dalloc(sizeof(__half) * 32 * n); out16 = (__half*)dalloc(sizeof(__half) * OUT_STRIDE * n); dout = (__half*)dalloc(sizeof(__half) * OUT_STRIDE * n);
		denc = (__half*)dalloc(sizeof(__half) * 32 * n); values = (float*)dalloc(sizeof(float) * OUT_STRIDE * n);
		positions = (float*)dalloc(sizeof(float) * 3 * n); targets = (float*)dalloc(sizeof(float) * 3 * n);
		ws_n = n;
	}

	// Testbed::reset_network (src/testbed.cu:2244-2470) + Trainer ctor / initialize_params (trainer.h:53-99)
	void reset(uint32_t seed_) {
		NGPB_CUDA_CHECK(cudaSetDevice(device));
		seed = seed_;
		rng.seed(seed);
		rng.next_uint(); // density_grid_rng = default_rng_t{m_rng.next_uint()} (:2265) draws from m_rng in every mode
		// m_per_level_scale = exp(log(desired_resolution * aabb_scale / base_resolution) / (n_levels - 1)) in float, by the host's libm like the reference
		// (src/testbed.cu:2318-2322): the level resolutions are ceil()s of powers of it, so the last bit matters
		float per_level_scale = cfg.per_level_scale;
		if (per_level_scale <= 0.0f) per_level_scale = std::exp(std::log(cfg.desired_resolution * (float)1 / (float)cfg.base_resolution) / (float)(cfg.n_levels - 1));
		const uint32_t entries = ngpb_grid_init_nd(&grid, cfg.n_pos_dims, cfg.n_levels, cfg.log2_hashmap_size, cfg.base_resolution, per_level_scale);
		if (entries == 0) throw std::runtime_error("model: invalid grid configuration");
		if (ngpb_grid_device_scales(stream, &grid) != 0) throw std::runtime_error(ngpb_last_error());
		const uint32_t new_n = NET_PARAMS + 2 * entries;
		if (new_n != n_params) {
			dfree(w_fp32); dfree(w_half); dfree(w_ema); dfree(m1); dfree(m2); dfree(param_steps); dfree(grad);
			n_params = new_n; n_alloc = next_multiple(n_params, 8u) + 8;
			w_fp32 = (float*)dalloc(sizeof(float) * n_alloc); w_half = (__half*)dalloc(sizeof(__half) * n_alloc); w_ema = (__half*)dalloc(sizeof(__half) * n_alloc);
			m1 = (float*)dalloc(sizeof(float) * n_alloc); m2 = (float*)dalloc(sizeof(float) * n_alloc); param_steps = (uint32_t*)dalloc(sizeof(uint32_t) * n_alloc);
			grad = (float*)dalloc(sizeof(float) * n_alloc);
		}
		NGPB_CUDA_CHECK(cudaMemsetAsync(w_fp32, 0, sizeof(float) * n_alloc, stream)); NGPB_CUDA_CHECK(cudaMemsetAsync(w_half, 0, sizeof(__half) * n_alloc, stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(w_ema, 0, sizeof(__half) * n_alloc, stream)); NGPB_CUDA_CHECK(cudaMemsetAsync(m1, 0, sizeof(float) * n_alloc, stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(m2, 0, sizeof(float) * n_alloc, stream)); NGPB_CUDA_CHECK(cudaMemsetAsync(param_steps, 0, sizeof(uint32_t) * n_alloc, stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(grad, 0, sizeof(float) * n_alloc, stream));
		// NetworkWithInputEncoding::initialize_params: the network's matrices (xavier uniform, host pcg32 seeded from std::seed_seq{seed}), then the grid
		std::seed_seq seq{seed};
		std::vector<uint32_t> seeds(2);
		seq.generate(seeds.begin(), seeds.end());
		Pcg32 rnd;
		rnd.seed(seeds.front());
		std::vector<float> net(NET_PARAMS);
		const int shapes[3][2] = {{64, 32}, {64, 64}, {16, 64}};
		size_t pos = 0;
		for (auto& s : shapes) {
			const float scale = std::sqrt(6.0f / (float)(s[0] + s[1]));
			for (int i = 0; i < s[0] * s[1]; ++i) net[pos++] = rnd.next_float() * 2.0f * scale - scale;
		}
		NGPB_CUDA_CHECK(cudaMemcpyAsync(w_fp32, net.data(), sizeof(float) * NET_PARAMS, cudaMemcpyHostToDevice, stream));
		random_uniform(stream, rnd, 2ull * entries, w_fp32 + NET_PARAMS, -1e-4f, 1e-4f);
		model_cast_params_kernel<<<div_round_up(n_params, 256), 256, 0, stream>>>(n_params, w_fp32, w_half);
		NGPB_LAUNCH_CHECK();
		const ngpb_optimizer hyper = cfg.optimizer;
		opt = hyper; opt.step = 0; opt.lr_factor = 1.0f;
		training_step = 0; loss_scalar = 0.f;
		if (!partials) partials = (float*)dalloc((size_t)ngpb_nerf_mlp_workspace_bytes());
		if (!loss_dev) loss_dev = (float*)dalloc(sizeof(float));
		if (!loss_host) NGPB_CUDA_CHECK(cudaMallocHost(&loss_host, sizeof(float)));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	}

	// Trainer::training_step (trainer.h:160-183) with loss scale 128, then optimizer_step. positions [n][n_pos_dims], targets [n][n_output_dims], device.
	void training_step_on(const float* pos_dev, const float* tgt_dev, uint32_t n, bool run_optimizer, bool get_loss) {
		if (n == 0 || n % 128 != 0) throw std::runtime_error("training batch size must be a non-zero multiple of 128 (tcnn batch_size_granularity)");
		const float loss_scale = 128.0f;
		hash_encode_forward_launch(stream, &grid, w_half + NET_PARAMS, pos_dev, n_in(), n, nullptr, enc, false);
		plain_mlp_launch(stream, w_half, enc, n, out16);
		if (ngpb_loss(stream, cfg.loss, n, cfg.n_output_dims, loss_scale, (const ngpb_half*)out16, tgt_dev, get_loss ? values : nullptr, (ngpb_half*)dout) != 0) throw std::runtime_error(ngpb_last_error());
		plain_mlp_forward_backward_launch(stream, w_half, enc, dout, n, denc, grad, partials);
		hash_encode_backward_launch(stream, &grid, pos_dev, n_in(), n, denc, grad + NET_PARAMS, 0, grid.n_levels);
		n_launches += 6;
		if (get_loss) {
			model_sum_kernel<<<1, 1024, 0, stream>>>(values, n * OUT_STRIDE, 1.0f, loss_dev);
			NGPB_LAUNCH_CHECK();
			NGPB_CUDA_CHECK(cudaMemcpyAsync(loss_host, loss_dev, sizeof(float), cudaMemcpyDeviceToHost, stream));
			++n_launches;
		}
		if (run_optimizer) optimizer_step(loss_scale);
		if (get_loss) { NGPB_CUDA_CHECK(cudaStreamSynchronize(stream)); loss_scalar = *loss_host; }
	}
	void optimizer_step(float loss_scale) {
		uint8_t P[256];
		optimizer_prepare(&opt, loss_scale, P);
		if (!use_ema()) optimizer_disable_fused_ema(P);
		optimizer_launch(stream, P, 0, n_params, NET_PARAMS, grad, w_fp32, w_half, w_ema, m1, m2, param_steps);
		++n_launches;
	}

	// Network::inference into float (any n: padded to the MLP's granularity internally). positions [n][n_pos_dims] device -> out [n][n_output_dims] device
	void inference(const float* pos_dev, uint32_t n, float* out_dev, bool use_inference_params) {
		if (n == 0) return;
		const uint32_t n_pad = next_multiple(n, 128u);
		ensure_workspace(n_pad);
		const __half* w = inference_params(use_inference_params);
		if (n_pad != n) NGPB_CUDA_CHECK(cudaMemsetAsync(enc + (size_t)n * 32, 0, sizeof(__half) * 32 * (n_pad - n), stream));
		hash_encode_forward_launch(stream, &grid, w + NET_PARAMS, pos_dev, n_in(), n, nullptr, enc, false);
		plain_mlp_launch(stream, w, enc, n_pad, out16);
		output_to_float_kernel<<<div_round_up(n * cfg.n_output_dims, 256u), 256, 0, stream>>>(n, cfg.n_output_dims, out16, out_dev);
		NGPB_LAUNCH_CHECK();
		n_launches += 3;
	}

	void eval_image(uint32_t n, float* pos, float* result, bool snap) {
		if (img_half) eval_image_and_snap_kernel<__half><<<div_round_up(n, 256u), 256, 0, stream>>>(n, (const __half*)image, (float2*)pos, img_w, img_h, result, snap, linear_colors != 0);
		else eval_image_and_snap_kernel<float><<<div_round_up(n, 256u), 256, 0, stream>>>(n, (const float*)image, (float2*)pos, img_w, img_h, result, snap, linear_colors != 0);
		NGPB_LAUNCH_CHECK();
		++n_launches;
	}
};

#define NGPB_MODEL_BEGIN try { if (!m) { set_last_error("null model"); return NGPB_ERR_INVALID_ARGUMENT; } NGPB_CUDA_CHECK(cudaSetDevice(m->device));
#define NGPB_MODEL_END return 0; } catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }

extern "C" int ngpb_model_create(ngpb_model** out, int device, const ngpb_model_config* cfg) {
	try {
		if (!out || !cfg) { set_last_error("ngpb_model_create: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		if ((cfg->n_pos_dims != 2 && cfg->n_pos_dims != 3) || (cfg->per_level_scale <= 0.0f && cfg->desired_resolution <= 0.0f) || cfg->base_resolution == 0 || cfg->n_output_dims == 0 || cfg->n_output_dims > 16 || cfg->n_levels != 16 ||
		    (cfg->loss != NGPB_ELEMENT_LOSS_L2 && cfg->loss != NGPB_ELEMENT_LOSS_MAPE && cfg->loss != NGPB_ELEMENT_LOSS_RELATIVE_L2)) {
			set_last_error("ngpb_model_create: unsupported configuration (2-D / 3-D hash grid of 16 levels x 2 features, 64-wide network with two hidden layers, L2 / RelativeL2 / MAPE loss)");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		if (ngpb_check_device(device) != 0) return NGPB_ERR_RUNTIME;
		auto* m = new ngpb_model();
		m->device = device; m->cfg = *cfg;
		try {
			NGPB_CUDA_CHECK(cudaSetDevice(device));
			NGPB_CUDA_CHECK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
			m->reset(cfg->seed);
		} catch (...) { delete m; throw; }
		*out = m;
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
extern "C" void ngpb_model_destroy(ngpb_model* m) { delete m; }
extern "C" int ngpb_model_reset(ngpb_model* m, uint32_t seed) { NGPB_MODEL_BEGIN m->reset(seed); NGPB_MODEL_END }
extern "C" uint32_t ngpb_model_n_params(const ngpb_model* m) { return m ? m->n_params : 0; }
extern "C" uint32_t ngpb_model_training_step(const ngpb_model* m) { return m ? m->training_step : 0; }
extern "C" float ngpb_model_loss(const ngpb_model* m) { return m ? m->loss_scalar : 0.f; }
extern "C" uint64_t ngpb_model_launches(const ngpb_model* m) { return m ? m->n_launches : 0; }
extern "C" void* ngpb_model_stream(ngpb_model* m) { return m ? (void*)m->stream : nullptr; }
extern "C" int ngpb_model_set_option(ngpb_model* m, const char* name, double v) {
	NGPB_MODEL_BEGIN
	const std::string k = name ? name : "";
	if (k == "snap_to_pixel_centers") m->snap_to_pixel_centers = v != 0;
	else if (k == "linear_colors") m->linear_colors = v != 0;
	else if (k == "learning_rate") { m->opt.learning_rate = (float)v; m->cfg.optimizer.learning_rate = (float)v; }
	else throw std::runtime_error("unknown model option: " + k);
	NGPB_MODEL_END
}

extern "C" int ngpb_model_get_params(ngpb_model* m, float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema) {
	NGPB_MODEL_BEGIN
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	if (w_fp32) NGPB_CUDA_CHECK(cudaMemcpy(w_fp32, m->w_fp32, sizeof(float) * m->n_params, cudaMemcpyDeviceToHost));
	if (w_half) NGPB_CUDA_CHECK(cudaMemcpy(w_half, m->w_half, sizeof(__half) * m->n_params, cudaMemcpyDeviceToHost));
	if (w_ema) NGPB_CUDA_CHECK(cudaMemcpy(w_ema, m->use_ema() ? m->w_ema : m->w_half, sizeof(__half) * m->n_params, cudaMemcpyDeviceToHost));
	NGPB_MODEL_END
}
// Trainer::deserialize for params_type "__half" (trainer.h:288-310): the snapshot parameters become the training, inference and fp32 master copies
extern "C" int ngpb_model_set_params_half(ngpb_model* m, const ngpb_half* params, uint32_t n) {
	NGPB_MODEL_BEGIN
	if (!params || n != m->n_params) throw std::runtime_error("set_params_half: parameter count does not match the model");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	// (stream-ordered copies: a plain cudaMemcpy from pageable memory may return before its DMA has landed, and this stream does not wait for the default one)
	NGPB_CUDA_CHECK(cudaMemcpyAsync(m->w_half, params, sizeof(__half) * n, cudaMemcpyHostToDevice, m->stream));
	NGPB_CUDA_CHECK(cudaMemcpyAsync(m->w_ema, params, sizeof(__half) * n, cudaMemcpyHostToDevice, m->stream));
	model_widen_params_kernel<<<div_round_up(n, 256u), 256, 0, m->stream>>>(n, m->w_half, m->w_fp32);
	NGPB_LAUNCH_CHECK();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	NGPB_MODEL_END
}
extern "C" int ngpb_model_set_training_step(ngpb_model* m, uint32_t step) { NGPB_MODEL_BEGIN m->training_step = step; NGPB_MODEL_END }

extern "C" int ngpb_model_train(ngpb_model* m, const float* positions_dev, const float* targets_dev, uint32_t n, int run_optimizer, int get_loss) {
	NGPB_MODEL_BEGIN
	if (!positions_dev || !targets_dev) throw std::runtime_error("ngpb_model_train: invalid argument");
	m->ensure_workspace(n);
	m->training_step_on(positions_dev, targets_dev, n, run_optimizer != 0, get_loss != 0);
	++m->training_step;
	NGPB_MODEL_END
}
extern "C" int ngpb_model_inference(ngpb_model* m, const float* positions_dev, uint32_t n, float* out_dev, int use_inference_params) {
	NGPB_MODEL_BEGIN
	if ((!positions_dev || !out_dev) && n) throw std::runtime_error("ngpb_model_inference: invalid argument");
	m->inference(positions_dev, n, out_dev, use_inference_params != 0);
	NGPB_MODEL_END
}

// ---- neural image ---------------------------------------------------------------------------------------------------------------------------------------
// m_image.data: RGBA float (PNG / JPEG / EXR through load_stbi / load_exr, linear, src/common_device.cu:49-80) or RGBA half (.bin), row-major, host pointer
extern "C" int ngpb_model_set_image(ngpb_model* m, const void* pixels_host, int width, int height, int is_half) {
	NGPB_MODEL_BEGIN
	if (m->cfg.n_pos_dims != 2 || m->cfg.n_output_dims != 3) throw std::runtime_error("set_image: the model is not a neural image (2-D input, 3 outputs)");
	if (!pixels_host || width <= 0 || height <= 0) throw std::runtime_error("set_image: invalid argument");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	m->dfree(m->image);
	const size_t bytes = (size_t)width * height * 4 * (is_half ? 2 : 4);
	m->image = m->dalloc(bytes);
	NGPB_CUDA_CHECK(cudaMemcpyAsync(m->image, pixels_host, bytes, cudaMemcpyHostToDevice, m->stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	m->img_w = width; m->img_h = height; m->img_half = is_half != 0;
	NGPB_MODEL_END
}

// load_stbi for an 8-bit file (src/common_device.cu:49-80): RGBA8 host pixels, converted on the device like the reference does
extern "C" int ngpb_model_set_image_rgba8(ngpb_model* m, const uint8_t* pixels_host, int width, int height) {
	NGPB_MODEL_BEGIN
	if (m->cfg.n_pos_dims != 2 || m->cfg.n_output_dims != 3) throw std::runtime_error("set_image: the model is not a neural image (2-D input, 3 outputs)");
	if (!pixels_host || width <= 0 || height <= 0) throw std::runtime_error("set_image: invalid argument");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	m->dfree(m->image);
	const uint32_t n = (uint32_t)width * (uint32_t)height;
	m->image = m->dalloc(sizeof(float) * 4 * n);
	uchar4* bytes = (uchar4*)m->dalloc((size_t)n * 4);
	NGPB_CUDA_CHECK(cudaMemcpyAsync(bytes, pixels_host, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
	from_rgba8_kernel<<<div_round_up(n, 256u), 256, 0, m->stream>>>(n, bytes, (float4*)m->image);
	NGPB_LAUNCH_CHECK();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	m->dfree(bytes);
	m->img_w = width; m->img_h = height; m->img_half = 0;
	NGPB_MODEL_END
}

// Testbed::train_image (:220-283), ERandomMode::Stratified: uniform positions from m_rng, stratified when the batch is an even power of two, snapped /
// evaluated against the image, one training step without the optimizer, then optimizer_step(128).
extern "C" int ngpb_model_train_image(ngpb_model* m, uint32_t batch, int get_loss) {
	NGPB_MODEL_BEGIN
	if (!m->image) throw std::runtime_error("train_image: no image loaded");
	if (batch == 0 || batch % 128 != 0) throw std::runtime_error("training batch size must be a non-zero multiple of 128 (tcnn batch_size_granularity)");
	m->ensure_workspace(batch);
	random_uniform(m->stream, m->rng, (uint64_t)batch * 2, m->positions, 0.0f, 1.0f);
	uint32_t log2_batch = 0;
	while ((1u << log2_batch) < batch) ++log2_batch;
	if ((1u << log2_batch) == batch && log2_batch % 2 == 0) { // (otherwise the reference warns and trains on the unstratified positions)
		stratify2_kernel<<<div_round_up(batch, 256u), 256, 0, m->stream>>>(batch, log2_batch, (float2*)m->positions);
		NGPB_LAUNCH_CHECK();
		++m->n_launches;
	}
	m->eval_image(batch, m->positions, m->targets, m->snap_to_pixel_centers != 0);
	++m->n_launches;
	m->training_step_on(m->positions, m->targets, batch, false, get_loss != 0);
	m->optimizer_step(128.0f);
	++m->training_step;
	NGPB_MODEL_END
}

// The last training batch (device -> host), for the parity tests: positions [n][n_pos_dims], targets [n][n_output_dims]
extern "C" int ngpb_model_get_training_batch(ngpb_model* m, uint32_t n, float* positions_host, float* targets_host) {
	NGPB_MODEL_BEGIN
	if (n > m->ws_n) throw std::runtime_error("get_training_batch: larger than the last batch");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	const float* p = m->cfg.n_pos_dims == 3 && m->sdf_pos_shuffled ? m->sdf_pos_shuffled : m->positions;
	const float* t = m->cfg.n_pos_dims == 3 && m->sdf_dist_shuffled ? m->sdf_dist_shuffled : m->targets;
	if (positions_host) NGPB_CUDA_CHECK(cudaMemcpy(positions_host, p, sizeof(float) * m->cfg.n_pos_dims * n, cudaMemcpyDeviceToHost));
	if (targets_host) NGPB_CUDA_CHECK(cudaMemcpy(targets_host, t, sizeof(float) * m->cfg.n_output_dims * n, cudaMemcpyDeviceToHost));
	NGPB_MODEL_END
}

// Testbed::compute_image_mse (:461-523): all texel centres in batches of 2^20, snapped targets, network inference, mean of the per-pixel squared error
extern "C" int ngpb_model_image_mse(ngpb_model* m, int quantize_to_byte, float* mse_out) {
	NGPB_MODEL_BEGIN
	if (!m->image || !mse_out) throw std::runtime_error("compute_image_mse: no image loaded");
	const uint32_t n_elements = (uint32_t)m->img_w * (uint32_t)m->img_h, max_batch = 1u << 20;
	float* se = (float*)m->dalloc(sizeof(float) * ((size_t)n_elements + 256));
	float* pred = (float*)m->dalloc(sizeof(float) * 3 * max_batch);
	m->ensure_workspace(max_batch);
	for (uint32_t offset = 0; offset < n_elements; offset += max_batch) {
		const uint32_t batch = (std::min(max_batch, n_elements - offset) + 255u) & ~255u;
		const uint32_t live = std::min(batch, n_elements - offset);
		image_coords_from_idx_kernel<<<div_round_up(batch, 256u), 256, 0, m->stream>>>(batch, offset, (float2*)m->positions, m->img_w, m->img_h);
		NGPB_LAUNCH_CHECK();
		m->eval_image(batch, m->positions, m->targets, true);
		m->inference(m->positions, batch, pred, true);
		image_mse_kernel<<<div_round_up(live, 256u), 256, 0, m->stream>>>(live, m->targets, pred, se + offset, quantize_to_byte != 0);
		NGPB_LAUNCH_CHECK();
		m->n_launches += 2;
	}
	model_sum_kernel<<<1, 1024, 0, m->stream>>>(se, n_elements, 1.0f / (float)n_elements, m->loss_dev);
	NGPB_LAUNCH_CHECK();
	NGPB_CUDA_CHECK(cudaMemcpyAsync(m->loss_host, m->loss_dev, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	*mse_out = *m->loss_host;
	m->dfree(se); m->dfree(pred);
	NGPB_MODEL_END
}

// Testbed::render_image (:285-347) for spp samples, then accumulate + tonemap as the frame leaves render_frame (src/render_buffer.cu:235-266,:540-567).
// view: {scale (m_scale), image_pos x, y (m_image.pos), screen centre x, y (m_screen_center)}; defaults {1, 0, 0, 0.5, 0.5} (Testbed::reset_camera).
extern "C" int ngpb_model_render_image(ngpb_model* m, int width, int height, int spp, const float* view5, int render_snap_to_pixel_centers, int color_space, int output_srgb,
                                       float exposure, const float* background4, int tonemap_curve, float* out_rgba_host) {
	NGPB_MODEL_BEGIN
	if (!m->image || !out_rgba_host || width <= 0 || height <= 0 || spp <= 0 || !view5 || !background4) throw std::runtime_error("render_image: invalid argument");
	const uint32_t n_pixels = (uint32_t)width * (uint32_t)height, n_elements = next_multiple(n_pixels, 128u);
	const size_t need = (size_t)n_elements * (2 + 3 + 4 + 4 + 4);
	if (need > m->render_ws_floats) { NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream)); m->dfree(m->render_ws); m->render_ws = (float*)m->dalloc(sizeof(float) * need); m->render_ws_floats = need; }
	float* coords = m->render_ws; float* colors = coords + (size_t)n_elements * 2; float* frame = colors + (size_t)n_elements * 3;
	float* accum = frame + (size_t)n_elements * 4; float* out = accum + (size_t)n_elements * 4;
	NGPB_CUDA_CHECK(cudaMemsetAsync(accum, 0, sizeof(float) * 4 * n_pixels, m->stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(coords, 0, sizeof(float) * 2 * n_elements, m->stream));
	for (int s = 0; s < spp; ++s) {
		float jit[2];
		ld_random_pixel_offset_host(render_snap_to_pixel_centers ? 0u : (uint32_t)s, jit);
		init_image_coords_kernel<<<div_round_up(n_pixels, 256u), 256, 0, m->stream>>>((float2*)coords, width, height, m->img_w, m->img_h, view5[0], view5[1], view5[2],
			view5[3] - 0.5f, view5[4] - 0.5f, jit[0], jit[1]);
		NGPB_LAUNCH_CHECK();
		// (the reference evaluates the ground-truth image here, which snaps the query positions onto texel centres when the model trains that way)
		m->ensure_workspace(n_elements);
		m->eval_image(n_elements, coords, m->targets, m->snap_to_pixel_centers != 0);
		m->inference(coords, n_elements, colors, true);
		shade_image_kernel<<<div_round_up(n_pixels, 256u), 256, 0, m->stream>>>(n_pixels, (const float2*)coords, colors, (float4*)frame, m->linear_colors != 0);
		NGPB_LAUNCH_CHECK();
		render_accumulate_launch(m->stream, n_pixels, frame, accum, (float)s, color_space);
		m->n_launches += 3;
	}
	render_tonemap_launch(m->stream, n_pixels, exposure, background4, accum, color_space, output_srgb, tonemap_curve, out);
	++m->n_launches;
	NGPB_CUDA_CHECK(cudaMemcpyAsync(out_rgba_host, out, sizeof(float) * 4 * n_pixels, cudaMemcpyDeviceToHost, m->stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	NGPB_MODEL_END
}

// ---- SDF on supplied pairs ------------------------------------------------------------------------------------------------------------------------------
// override_sdf_training_data (src/python_api.cu:74-104) after the host has mapped points / distances into the unit cube: the pool the training batches come from
extern "C" int ngpb_model_set_sdf_data(ngpb_model* m, const float* positions_host, const float* distances_host, uint32_t n) {
	NGPB_MODEL_BEGIN
	if (m->cfg.n_pos_dims != 3 || m->cfg.n_output_dims != 1) throw std::runtime_error("set_sdf_data: the model is not an SDF (3-D input, 1 output)");
	if (!positions_host || !distances_host || n == 0) throw std::runtime_error("set_sdf_data: invalid argument");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	m->dfree(m->sdf_pos); m->dfree(m->sdf_dist); m->dfree(m->sdf_pos_shuffled); m->dfree(m->sdf_dist_shuffled);
	m->sdf_pos = (float*)m->dalloc(sizeof(float) * 3 * n); m->sdf_dist = (float*)m->dalloc(sizeof(float) * n);
	m->sdf_pos_shuffled = (float*)m->dalloc(sizeof(float) * 3 * n); m->sdf_dist_shuffled = (float*)m->dalloc(sizeof(float) * n);
	NGPB_CUDA_CHECK(cudaMemcpyAsync(m->sdf_pos, positions_host, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, m->stream));
	NGPB_CUDA_CHECK(cudaMemcpyAsync(m->sdf_dist, distances_host, sizeof(float) * n, cudaMemcpyHostToDevice, m->stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(m->stream));
	m->sdf_size = n;
	NGPB_MODEL_END
}

// Testbed::train_sdf (:1229-1252): nothing happens unless the pool holds a full batch; the pool is permuted with the step as seed and the first `batch`
// records are one training step including the optimizer.
extern "C" int ngpb_model_train_sdf(ngpb_model* m, uint32_t batch, int get_loss) {
	NGPB_MODEL_BEGIN
	if (batch == 0 || batch % 128 != 0) throw std::runtime_error("training batch size must be a non-zero multiple of 128 (tcnn batch_size_granularity)");
	if (m->sdf_size < batch) return 0;
	m->ensure_workspace(batch);
	shuffle_kernel<<<div_round_up(m->sdf_size * 3, 256u), 256, 0, m->stream>>>(m->sdf_size, 3, m->training_step, m->sdf_pos, m->sdf_pos_shuffled);
	NGPB_LAUNCH_CHECK();
	shuffle_kernel<<<div_round_up(m->sdf_size, 256u), 256, 0, m->stream>>>(m->sdf_size, 1, m->training_step, m->sdf_dist, m->sdf_dist_shuffled);
	NGPB_LAUNCH_CHECK();
	m->n_launches += 2;
	m->training_step_on(m->sdf_pos_shuffled, m->sdf_dist_shuffled, batch, true, get_loss != 0);
	++m->training_step;
	NGPB_MODEL_END
}
