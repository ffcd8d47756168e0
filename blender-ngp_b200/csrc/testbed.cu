// Host side of the NeRF mode: the `Testbed`-shaped object behind the ngpb_testbed_* C ABI.
// Mirrors, for this path only, ngp::Testbed (reference: include/neural-graphics-primitives/testbed.h,
// src/testbed.cu:2249-2470 reset_network, :2527-2588 train; src/testbed_nerf.cu:2643-2733 load_nerf_post,
// :2761-2859 density grid, :2861-2968 counters + train_nerf, :3138-3385 train_nerf_step, :3388-3401
// training_prep_nerf). All device memory is owned here and allocated once per (dataset, batch size):
// nothing is allocated in the steady state, like the reference's per-stream arena.
#include "testbed.h"
#include <cuda_bf16.h>
#include "nccl_dl.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>

namespace ngpb {

// launchers defined in the kernel translation units
void hash_encode_forward_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* positions, uint32_t pos_stride, uint32_t n, const uint32_t* n_dev, __half* encoded, bool tiled);
bool features_tiled();
int compute_loss_launch(void* stream_, uint32_t n_rays, uint32_t n_rays_global, const float* aabb6, ngpb_rng rng_, uint32_t batch, const ngpb_loss_config* cfg,
                        uint32_t n_images, const ngpb_image* images_dev, const uint32_t* counters_in,
                        const ngpb_half* rgbsigma, const uint32_t* ray_indices, const float* rays, uint32_t* numsteps, const float* coords_in,
                        const float* mean_density_dev, float* coords_out, ngpb_half* dloss_dout, float* loss_per_ray, uint32_t* counters_out, void* scratch,
                        const ngpb_half* encoded_in, ngpb_half* encoded_out, bool rows_tiled, const float* exposure, float* exposure_gradient,
                        const ErrorCdf& cdf, float* error_map, int error_map_res_x, int error_map_res_y);
void hash_encode_backward_launch(cudaStream_t stream, const ngpb_grid* g, const float* positions, uint32_t pos_stride, uint32_t n, const __half* dL_dencoded, float* grid_grad,
                                 uint32_t level_begin, uint32_t level_end);
void optimizer_prepare(ngpb_optimizer* o, float loss_scale, void* params_out);
size_t optimizer_params_bytes();
void optimizer_disable_fused_ema(void* params);
void ema_sweep_launch(cudaStream_t stream, const void* params, uint32_t n_padded, const __half* w_half, __half* w_ema);
void optimizer_launch(cudaStream_t stream, const void* params, uint32_t first, uint32_t count, uint32_t n_matrix_params, float* grad, float* w_fp32, __half* w_half,
                      __half* w_ema, float* m1, float* m2, uint32_t* param_steps);
void nerf_mlp_forward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma);
void nerf_density_mlp_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, uint32_t n, __half* density);
void nerf_mlp_forward_backward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, const __half* dL_dout, uint32_t n, __half* dL_dencoded, float* mlp_grad, float* partials, __half* dL_dsh);
void nerf_input_gradient_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* coords, uint32_t n, const __half* dL_dencoded, const __half* dL_dsh, float* coords_gradient);
void cam_gradient_launch(cudaStream_t stream, uint32_t max_rays, uint32_t n_rays_global, const float* aabb6, const uint32_t* rays_counter, uint32_t n_images, const uint32_t* ray_indices,
                         const float* rays, const uint32_t* numsteps, const float* coords, const float* coords_gradient, float* cam_pos_gradient, float* cam_rot_gradient,
                         const float* cdf_img);

// sum of n floats in double, one block, fixed order (tcnn reduce_sum.h:54-118 is the reference's loss reduction)
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ v, const uint32_t n, float* __restrict__ out)
{
	__shared__ double sm[1024];
	double s = 0.0;
	for (uint32_t i = threadIdx.x; i < n; i += 1024) s += (double)v[i];
	sm[threadIdx.x] = s;
	__syncthreads();
	for (uint32_t o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) *out = (float)sm[0];
}

__global__ void __launch_bounds__(256) widen_params_kernel(const uint32_t n, const __half* __restrict__ w_half, float* __restrict__ w_fp32)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) w_fp32[i] = __half2float(w_half[i]);
}

__global__ void __launch_bounds__(256) cast_params_kernel(const uint32_t n, const float* __restrict__ w_fp32, __half* __restrict__ w_half)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) w_half[i] = __float2half_rn(w_fp32[i]);
}

// Grid initialisation on the device: uniform in [-1e-4, 1e-4) with the reference's thread -> element
// mapping (tcnn random.h:66-97: thread i advances the stream by 4i and writes elements i + k*n_threads).
__global__ void __launch_bounds__(128) init_grid_kernel(const uint64_t n_elements, const uint64_t n_threads, Pcg32 rng, float* __restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_threads) return;
	rng.advance((int64_t)(i * 4));
	#pragma unroll
	for (uint64_t j = 0; j < 4; ++j) {
		const uint64_t idx = i + n_threads * j;
		if (idx >= n_elements) return;
		out[idx] = __fmaf_rn(rng.next_float(), 1e-4f - -1e-4f, -1e-4f); // the reference's lambda (val * (upper - lower) + lower) is compiled with FMA contraction
	}
}

Aabb make_aabb(const float* a);

} // namespace ngpb

using namespace ngpb;

// Camera transform used by the sampling kernel: matrix -> quaternion -> slerp(t = 0) -> normalize -> matrix,
// i.e. what get_xform_given_rolling_shutter (common_device.cuh:224-234) evaluates for every ray when there is
// no rolling shutter. Done once per image on the host. Operation order follows Eigen's Quaternion.h.
extern "C" void ngpb_effective_xform(const float* m12, float* out12) {
	auto M = [&](int r, int c) { return m12[c * 3 + r]; };
	volatile float q[4]; // x y z w; volatile keeps the host compiler from contracting/reassociating
	float t = M(0, 0) + (M(1, 1) + M(2, 2));
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
		const int j = (i + 1) % 3, k = (j + 1) % 3;
		t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0f);
		q[i] = 0.5f * t;
		t = 0.5f / t;
		q[3] = (M(k, j) - M(j, k)) * t;
		q[j] = (M(j, i) + M(i, j)) * t;
		q[k] = (M(k, i) + M(i, k)) * t;
	}
	volatile float sq[4];
	for (int c = 0; c < 4; ++c) sq[c] = q[c] * q[c];
	volatile float s01 = sq[0] + sq[1], s23 = sq[2] + sq[3];
	const float z = s01 + s23;
	if (z > 0.f) { const float nrm = std::sqrt(z); for (int c = 0; c < 4; ++c) q[c] = q[c] / nrm; }
	volatile float tx = 2.f * q[0], ty = 2.f * q[1], tz = 2.f * q[2];
	volatile float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
	volatile float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
	volatile float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
	auto R = [&](int r, int c) -> float& { return out12[c * 3 + r]; };
	volatile float a;
	a = tyy + tzz; R(0, 0) = 1.f - a; R(0, 1) = txy - twz; R(0, 2) = txz + twy;
	a = txx + tzz; R(1, 0) = txy + twz; R(1, 1) = 1.f - a; R(1, 2) = tyz - twx;
	a = txx + tyy; R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1.f - a;
	out12[9] = m12[9]; out12[10] = m12[10]; out12[11] = m12[11];
}

// Data-parallel gradient exchange in fp16: cast the fp32 accumulation buffer (4 values per thread) and reset it in the same pass; widen the reduced slice.
// ---- peer-memory gradient / weight exchange (dp_exchange = 1) ------------------------------------------------------------------------------
// One kernel per direction instead of cast + ncclReduceScatter + widen and ncclAllGather:
//   scatter: every rank casts its fp32 partial gradients to bf16 and stores slice s straight into rank s's receive buffer over NVLink (and resets its
//            accumulation buffer); the last block to finish publishes "step n delivered" in every peer's flag array (system-scope release).
//   reduce : the owner waits for all world flags (system-scope acquire), sums the world slices of its range in rank order in fp32.
//   gather : after Adam, every rank stores its updated fp16 weights into every peer's weight array and publishes a second flag; the next kernel that
//            reads the weights is preceded by a wait on those flags.
// Ordering across steps needs no double buffering: rank B only overwrites A's receive buffer in step n+1 after it has seen A's weights of step n, which A
// sent after its reduce of step n had consumed the buffer.
struct P2pTable { void* ptr[16][3]; }; // [rank] = {recv, flags, w_half}
constexpr uint32_t P2P_SPIN_LIMIT = 1u << 26; // ~ seconds; on expiry the error word is set and the host raises at the next read-back

__device__ __forceinline__ void p2p_publish(uint32_t* flag, uint32_t value) { asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(value) : "memory"); }
__device__ __forceinline__ uint32_t p2p_peek(const uint32_t* flag) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); return v; }
__device__ __forceinline__ void p2p_wait_all(const uint32_t* flags, uint32_t world, uint32_t step, uint32_t* error_word) {
	if (threadIdx.x < world) {
		uint32_t spins = 0;
		while ((int32_t)(p2p_peek(flags + threadIdx.x) - step) < 0) { if (++spins > P2P_SPIN_LIMIT) { *error_word = 0xDEADu; break; } __nanosleep(64); }
	}
	__syncthreads();
}

__global__ void __launch_bounds__(256) p2p_scatter_gradients_kernel(const uint32_t n8, const uint32_t count, const uint32_t rank, const uint32_t world, const uint32_t step,
                                                                    float4* __restrict__ grad, const P2pTable T, uint32_t* __restrict__ my_flags)
{
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; // group of 8 parameters (count is a multiple of 8, so a group never straddles two slices)
	if (q < n8) {
		const float4 a = grad[2 * q], b = grad[2 * q + 1];
		const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w), h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
		const uint32_t i = q * 8, s = i / count, off = i - s * count;
		uint4 v = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1), *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
		uint16_t* dst = reinterpret_cast<uint16_t*>(T.ptr[s][0]) + (size_t)rank * count + off;
		*reinterpret_cast<uint4*>(dst) = v;
		grad[2 * q] = make_float4(0.f, 0.f, 0.f, 0.f); grad[2 * q + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint32_t done = atomicAdd(my_flags + 2 * world, 1u);
		if (done == gridDim.x - 1) {
			my_flags[2 * world] = 0u;
			__threadfence_system();
			for (uint32_t p = 0; p < world; ++p) p2p_publish(reinterpret_cast<uint32_t*>(T.ptr[p][1]) + rank, step);
		}
	}
}

__global__ void __launch_bounds__(256) p2p_reduce_gradients_kernel(const uint32_t mine, const uint32_t count, const uint32_t world, const uint32_t step,
                                                                   const uint16_t* __restrict__ recv, uint32_t* __restrict__ my_flags, float* __restrict__ grad_shard)
{
	p2p_wait_all(my_flags, world, step, my_flags + 2 * world + 1);
	const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
	if (j >= mine) return;
	float s0 = 0.f, s1 = 0.f;
	for (uint32_t r = 0; r < world; ++r) { // fixed order: the same bits on every run
		const uint32_t v = *reinterpret_cast<const volatile uint32_t*>(recv + (size_t)r * count + j);
		s0 += __uint_as_float(v << 16); s1 += __uint_as_float(v & 0xFFFF0000u);
	}
	grad_shard[j] = s0;
	if (j + 1 < mine) grad_shard[j + 1] = s1;
}

__global__ void __launch_bounds__(256) p2p_gather_weights_kernel(const uint32_t n8, const uint32_t first, const uint32_t rank, const uint32_t world, const uint32_t step,
                                                                 const __half* __restrict__ w_half, const P2pTable T, uint32_t* __restrict__ my_flags)
{
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q < n8) {
		const uint4 v = *reinterpret_cast<const uint4*>(w_half + first + (size_t)q * 8);
		for (uint32_t p = 0; p < world; ++p) if (p != rank) *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(T.ptr[p][2]) + first + (size_t)q * 8) = v;
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint32_t done = atomicAdd(my_flags + 2 * world, 1u);
		if (done == gridDim.x - 1) {
			my_flags[2 * world] = 0u;
			__threadfence_system();
			for (uint32_t p = 0; p < world; ++p) p2p_publish(reinterpret_cast<uint32_t*>(T.ptr[p][1]) + world + rank, step);
		}
	}
}

__global__ void __launch_bounds__(32) p2p_wait_weights_kernel(const uint32_t world, const uint32_t step, uint32_t* __restrict__ my_flags)
{
	p2p_wait_all(my_flags + world, world, step, my_flags + 2 * world + 1);
}

template <bool BF16>
__global__ void __launch_bounds__(256) grads_to_16bit_and_reset_kernel(const uint32_t n4, float4* __restrict__ grad, uint2* __restrict__ out)
{
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n4) return;
	const float4 g = grad[q];
	uint2 o;
	if (BF16) {
		const __nv_bfloat162 a = __floats2bfloat162_rn(g.x, g.y), b = __floats2bfloat162_rn(g.z, g.w);
		o = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
	} else {
		const __half2 a = __floats2half2_rn(g.x, g.y), b = __floats2half2_rn(g.z, g.w);
		o = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
	}
	out[q] = o;
	grad[q] = make_float4(0.f, 0.f, 0.f, 0.f);
}
template <bool BF16>
__global__ void __launch_bounds__(256) widen_16bit_kernel(const uint32_t n, const uint16_t* __restrict__ in, float* __restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = BF16 ? __uint_as_float((uint32_t)in[i] << 16) : __half2float(__ushort_as_half(in[i]));
}

// ---------------------------------------------------------------------------------------------------
ngpb_testbed::ngpb_testbed(int device_) : device(device_) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	// Stream priorities (NGPB_STREAM_PRIORITY = "sampling" | "main" | "equal"): whichever stream has the higher priority gets its pending blocks placed
	// first as resident blocks retire; the other stream's kernels fill what is left.
	int prio_low = 0, prio_high = 0;
	NGPB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
	if (const char* ge = std::getenv("NGPB_DP_HALF_GRADIENTS")) dp_half_gradients = std::atoi(ge); // 0 fp32, 1 bf16 (default), 2 fp16
	if (const char* xe = std::getenv("NGPB_DP_EXCHANGE")) dp_exchange = std::atoi(xe);             // 0 NCCL (default), 1 peer memory
	const char* pe = std::getenv("NGPB_STREAM_PRIORITY");
	const std::string pmode = pe ? pe : "sampling";
	NGPB_CUDA_CHECK(cudaStreamCreateWithPriority(&stream, cudaStreamNonBlocking, pmode == "main" ? prio_high : prio_low));
	NGPB_CUDA_CHECK(cudaStreamCreateWithPriority(&sampling_stream, cudaStreamNonBlocking, pmode == "sampling" ? prio_high : prio_low));
	NGPB_CUDA_CHECK(cudaStreamCreateWithPriority(&ema_stream, cudaStreamNonBlocking, prio_low));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&weights_gathered, cudaEventDisableTiming));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&ema_done, cudaEventDisableTiming));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&prefetch_done, cudaEventDisableTiming));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&loss_ready, cudaEventDisableTiming));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&counters_ready, cudaEventDisableTiming));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&mlp_train_done, cudaEventDisableTiming));
	NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&error_cdf_built, cudaEventDisableTiming));
	ngpb_optimizer_init(&opt);
	ngpb_optimizer_init(&opt_hyper);
	loss_cfg.loss_scale = LOSS_SCALE;
	loss_cfg.background_color[0] = loss_cfg.background_color[1] = loss_cfg.background_color[2] = 0.f;
	loss_cfg.color_space = NGPB_COLOR_LINEAR;     // m_color_space default, testbed.h:846
	loss_cfg.random_bg_color = 1;                 // testbed.h:651
	loss_cfg.linear_colors = 0;
	loss_cfg.loss_type = NGPB_LOSS_HUBER;         // configs/nerf/base.json:2-4
	loss_cfg.rgb_activation = NGPB_ACT_LOGISTIC;  // LDR dataset (testbed_nerf.cu:2644)
	loss_cfg.density_activation = NGPB_ACT_EXPONENTIAL;
	loss_cfg.snap_to_pixel_centers = 1;           // testbed.h nerf.training.snap_to_pixel_centers default true
	loss_cfg.near_distance = 0.2f;                // testbed.h:675
}

ngpb_testbed::~ngpb_testbed() {
	cudaSetDevice(device);
	if (sampling_stream) cudaStreamSynchronize(sampling_stream);
	if (stream) cudaStreamSynchronize(stream);
	p2p_teardown();
	if (nccl_comm) { try { NcclApi::get().CommDestroy(nccl_comm); } catch (...) {} }
	if (ema_stream) { cudaStreamSynchronize(ema_stream); cudaStreamDestroy(ema_stream); }
	if (weights_gathered) cudaEventDestroy(weights_gathered);
	if (ema_done) cudaEventDestroy(ema_done);
	if (prefetch_done) cudaEventDestroy(prefetch_done);
	if (loss_ready) cudaEventDestroy(loss_ready);
	if (counters_ready) cudaEventDestroy(counters_ready);
	if (mlp_train_done) cudaEventDestroy(mlp_train_done);
	if (error_cdf_built) cudaEventDestroy(error_cdf_built);
	if (sampling_stream) cudaStreamDestroy(sampling_stream);
	for (void* p : allocations) cudaFree(p);
	for (auto& e : stage_ev) { if (e[0]) cudaEventDestroy(e[0]); if (e[1]) cudaEventDestroy(e[1]); }
	if (host_readback) cudaFreeHost(host_readback);
	if (stream) cudaStreamDestroy(stream);
}

void ngpb_testbed::stage_begin(int s, cudaStream_t st) {
	if (!profile_stages) return;
	if (!stage_ev[s][0]) { NGPB_CUDA_CHECK(cudaEventCreate(&stage_ev[s][0])); NGPB_CUDA_CHECK(cudaEventCreate(&stage_ev[s][1])); }
	if (stage_used[s]) stage_collect(); // a stage that runs twice before a collect: fold the first interval in now
	NGPB_CUDA_CHECK(cudaEventRecord(stage_ev[s][0], st));
}
void ngpb_testbed::stage_end(int s, uint64_t units, cudaStream_t st) {
	if (!profile_stages) return;
	NGPB_CUDA_CHECK(cudaEventRecord(stage_ev[s][1], st));
	stage_used[s] = true;
	stage_calls[s] += 1;
	stage_units[s] += units;
}
void ngpb_testbed::stage_collect() {
	for (int s = 0; s < NGPB_N_STAGES; ++s) {
		if (!stage_used[s]) continue;
		NGPB_CUDA_CHECK(cudaEventSynchronize(stage_ev[s][1]));
		float ms = 0.f;
		NGPB_CUDA_CHECK(cudaEventElapsedTime(&ms, stage_ev[s][0], stage_ev[s][1]));
		stage_ms[s] += ms;
		stage_used[s] = false;
	}
}

void* ngpb_testbed::dalloc(size_t bytes) {
	void* p = nullptr;
	NGPB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
	allocations.push_back(p);
	return p;
}
void ngpb_testbed::dfree(void* p) {
	if (!p) return;
	auto it = std::find(allocations.begin(), allocations.end(), p);
	if (it != allocations.end()) allocations.erase(it);
	cudaFree(p);
}

// Testbed::load_training_data -> load_nerf -> load_nerf_post (src/testbed_nerf.cu:2643-2733)
static size_t image_bytes_per_pixel(int32_t image_type) {
	if (image_type == NGPB_IMAGE_BYTE) return 4;
	if (image_type == NGPB_IMAGE_HALF) return 8;
	if (image_type == NGPB_IMAGE_FLOAT) return 16;
	throw std::runtime_error("training image: unknown image_type");
}

void ngpb_testbed::load_training_data(uint32_t n, const ngpb_host_image* host_images, uint32_t aabb_scale_, bool allow_empty) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	if (n == 0 || !host_images) throw std::runtime_error("load_training_data: no images");
	if (aabb_scale_ == 0 || (aabb_scale_ & (aabb_scale_ - 1)) != 0) throw std::runtime_error("NeRF dataset's `aabb_scale` must be a power of two");
	if (aabb_scale_ > (1u << (NERF_CASCADES - 1))) throw std::runtime_error("NeRF dataset must have `aabb_scale <= 128`");
	aabb_scale = aabb_scale_;
	size_t total = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const bool empty = !host_images[i].pixels && host_images[i].w == 0 && host_images[i].h == 0;
		if (!(allow_empty && empty) && (!host_images[i].pixels || host_images[i].w <= 0 || host_images[i].h <= 0)) throw std::runtime_error("load_training_data: invalid image");
		total += next_multiple((size_t)host_images[i].w * host_images[i].h * image_bytes_per_pixel(host_images[i].image_type), (size_t)16); // (float pixels are read as 16-byte vectors)
	}
	drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	// a reload of a same-sized dataset reuses the device buffers (no allocation in the steady state, like the reference's arena)
	for (void* p : own_pixels) dfree(p);
	own_pixels.assign(n, nullptr);
	if (total != pixels_bytes || n != images.size()) {
		dfree(pixels); dfree(images_dev);
		pixels = (uint8_t*)dalloc(total);
		images_dev = (ngpb_image*)dalloc(sizeof(ngpb_image) * n);
		pixels_bytes = total;
	}
	images.resize(n);
	dataset_xforms.assign((size_t)n * 12, 0.f);
	dfree(cam_gradients); dfree(cam_exposure);
	// error-map buffers are sized by the image count: start over
	dfree(error_map); dfree(error_cdf_x_cond_y); dfree(error_cdf_y); dfree(error_cdf_img); dfree(error_pmf_img);
	error_map = error_cdf_x_cond_y = error_cdf_y = error_cdf_img = error_pmf_img = nullptr;
	error_map_capacity = error_cdf_capacity = 0;
	error_cdf_valid = error_map_live = error_cdf_used = false;
	cam_gradients = (float*)dalloc(sizeof(float) * 9 * n);
	cam_exposure = (float*)dalloc(sizeof(float) * 3 * n);
	NGPB_CUDA_CHECK(cudaMemsetAsync(cam_gradients, 0, sizeof(float) * 9 * n, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(cam_exposure, 0, sizeof(float) * 3 * n, stream));
	cam_gradients_host.assign((size_t)n * 9, 0.f);
	cam_pos_state.assign((size_t)n * 10, 0.f); cam_rot_state.assign((size_t)n * 10, 0.f); cam_exposure_state.assign((size_t)n * 10, 0.f);
	exposure_active = false;
	n_steps_since_cam_update = 0;
	size_t off = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const ngpb_host_image& h = host_images[i];
		std::memcpy(&dataset_xforms[(size_t)i * 12], h.xform, sizeof(float) * 12);
		const size_t bytes = (size_t)h.w * h.h * image_bytes_per_pixel(h.image_type);
		if (bytes) NGPB_CUDA_CHECK(cudaMemcpyAsync(pixels + off, h.pixels, bytes, cudaMemcpyHostToDevice, stream));
		ngpb_image& im = images[i];
		im.pixels = pixels + off;
		im.image_type = h.image_type;
		im.w = h.w; im.h = h.h; im.fx = h.fx; im.fy = h.fy; im.cx = h.cx; im.cy = h.cy;
		if (h.lens_mode < NGPB_LENS_PERSPECTIVE || h.lens_mode > NGPB_LENS_LATLONG) throw std::runtime_error("load_training_data: unknown lens mode");
		im.lens_mode = h.lens_mode; std::memcpy(im.lens_params, h.lens_params, sizeof(im.lens_params));
		std::memcpy(im.raw_xform, h.xform, sizeof(float) * 12);
		ngpb_effective_xform(h.xform, im.xform);
		off += next_multiple(bytes, (size_t)16);
	}
	n_images_for_training = n_images_for_training_prev = n; // load_nerf_post (:2646)
	NGPB_CUDA_CHECK(cudaMemcpyAsync(images_dev, images.data(), sizeof(ngpb_image) * n, cudaMemcpyHostToDevice, stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	h2d_bytes += total + sizeof(ngpb_image) * n;

	configure_scene_box();
	training_data_available = true;
	reset_network(seed);
}

// load_nerf_post (src/testbed_nerf.cu:2714-2730): scene box, cascade count and cone angle follow from aabb_scale
void ngpb_testbed::configure_scene_box() {
	const float half = 0.5f * std::min(1u << (NERF_CASCADES - 1), aabb_scale);
	for (int c = 0; c < 3; ++c) { aabb[c] = 0.5f - half; aabb[3 + c] = 0.5f + half; }
	max_cascade = 0;
	while ((1u << max_cascade) < aabb_scale) ++max_cascade;
	cone_angle_constant = aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);
	scene_configured = true;
}

// Testbed::reset_network (src/testbed.cu:2249-2470) for configs/nerf/base.json
void ngpb_testbed::reset_network(uint32_t seed_) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	drop_prefetch();
	loss_pending = false;
	seed = seed_;
	// m_rng = default_rng_t{m_seed}; density_grid_rng = default_rng_t{m_rng.next_uint()} (:2252,:2265)
	rng.seed(seed);
	density_grid_rng.seed(rng.next_uint());
	// m_nerf.training.reset_camera_extrinsics() (src/testbed.cu:2267); the training transforms follow at once (the reference leaves them until the next update)
	if (!images.empty()) {
		bool moved = false;
		for (size_t i = 0; i < images.size() && !moved; ++i) for (int c = 7; c < 10; ++c) moved |= cam_pos_state[i * 10 + c] != 0.f || cam_rot_state[i * 10 + c] != 0.f;
		reset_camera_extrinsics();
		if (moved) update_transforms();
		if (exposure_active) upload_exposures();
	}
	n_steps_since_cam_update = 0;
	// (src/testbed.cu:2261-2264)
	n_steps_since_error_map_update = 0;
	n_steps_between_error_map_updates = 128;
	error_cdf_valid = false;
	error_map_live = false;
	const uint32_t n_levels = 16, base_resolution = 16;
	// per_level_scale = exp(ln(desired_resolution * aabb_scale / base) / (L-1)) (:2313-2325)
	const float per_level_scale = std::exp(std::log(2048.0f * (float)aabb_scale / (float)base_resolution) / (n_levels - 1));
	const uint32_t entries = ngpb_grid_init(&grid, n_levels, log2_hashmap_size, base_resolution, per_level_scale);
	if (ngpb_grid_device_scales(stream, &grid) != 0) throw std::runtime_error(ngpb_last_error());
	const uint32_t new_n_params = MLP_PARAMS + 2 * entries;
	if (new_n_params != n_params) {
		dfree(w_fp32); dfree(w_half); dfree(w_ema); dfree(m1); dfree(m2); dfree(param_steps); dfree(grad); dfree(grad_half); grad_half = nullptr;
		p2p_teardown(); // the peers hold mappings of the old weight array
		n_params = new_n_params;
		n_alloc = n_params + 4096; // slack: the sharded data-parallel optimizer works on world x ceil(n_params / world) padded ranges
		w_fp32 = (float*)dalloc(sizeof(float) * n_alloc);
		w_half = (__half*)dalloc(sizeof(__half) * n_alloc);
		w_ema = (__half*)dalloc(sizeof(__half) * n_alloc);
		m1 = (float*)dalloc(sizeof(float) * n_alloc);
		m2 = (float*)dalloc(sizeof(float) * n_alloc);
		param_steps = (uint32_t*)dalloc(sizeof(uint32_t) * n_alloc);
		grad = (float*)dalloc(sizeof(float) * n_alloc);
	}
	NGPB_CUDA_CHECK(cudaMemsetAsync(w_fp32, 0, sizeof(float) * n_alloc, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(w_half, 0, sizeof(__half) * n_alloc, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(w_ema, 0, sizeof(__half) * n_alloc, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(m1, 0, sizeof(float) * n_alloc, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(m2, 0, sizeof(float) * n_alloc, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(param_steps, 0, sizeof(uint32_t) * n_alloc, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(grad, 0, sizeof(float) * n_alloc, stream));
	master_weights_sharded = false;

	// Trainer ctor (tcnn trainer.h:53-99): pcg32 seeded from std::seed_seq{seed}; xavier-uniform MLP matrices drawn on the
	// host in parameter order (gpu_matrix.h:291-305), grid drawn on the device (grid.h:1364-1369).
	std::seed_seq seq{seed};
	std::vector<uint32_t> seeds(2);
	seq.generate(seeds.begin(), seeds.end());
	Pcg32 rnd;
	rnd.seed(seeds.front());
	std::vector<float> mlp(MLP_PARAMS);
	const int shapes[5][2] = {{64, 32}, {16, 64}, {64, 32}, {64, 64}, {16, 64}};
	size_t pos = 0;
	for (auto& s : shapes) {
		const float scale = std::sqrt(6.0f / (float)(s[0] + s[1]));
		for (int i = 0; i < s[0] * s[1]; ++i) mlp[pos++] = rnd.next_float() * 2.0f * scale - scale;
	}
	NGPB_CUDA_CHECK(cudaMemcpyAsync(w_fp32, mlp.data(), sizeof(float) * MLP_PARAMS, cudaMemcpyHostToDevice, stream));
	const uint64_t n_grid = 2ull * entries;
	const uint64_t n_threads = next_multiple((uint32_t)((n_grid + 3) / 4), 128);
	init_grid_kernel<<<(uint32_t)(n_threads / 128), 128, 0, stream>>>(n_grid, n_threads, rnd, w_fp32 + MLP_PARAMS);
	NGPB_LAUNCH_CHECK();
	cast_params_kernel<<<div_round_up(n_params, 256), 256, 0, stream>>>(n_params, w_fp32, w_half);
	NGPB_LAUNCH_CHECK();
	n_launches += 2;

	// density grid state (testbed.h:698-704)
	const size_t n_cells = (size_t)NERF_GRID_CELLS * (max_cascade + 1);
	const bool realloc_grid = n_cells != density_grid_cells;
	if (realloc_grid) {
		dfree(density_grid); dfree(density_grid_tmp); dfree(bitfield); dfree(mean_density);
		density_grid = (float*)dalloc(sizeof(float) * n_cells);
		density_grid_tmp = (float*)dalloc(sizeof(float) * n_cells);
		bitfield = (uint8_t*)dalloc((size_t)NERF_GRID_CELLS * NERF_CASCADES / 8);
		mean_density = (float*)dalloc(NGPB_MEAN_WORKSPACE_BYTES);
	}
	NGPB_CUDA_CHECK(cudaMemsetAsync(density_grid, 0, sizeof(float) * n_cells, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(bitfield, 0, (size_t)NERF_GRID_CELLS * NERF_CASCADES / 8, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(mean_density, 0, sizeof(float), stream));
	// density-grid sample buffers: at most GRID_CELLS * n_cascades samples per refresh (testbed_nerf.cu:3396-3400)
	const size_t n_dg = next_multiple((uint32_t)n_cells, 128);
	if (realloc_grid) {
		dfree(dg_positions); dfree(dg_indices); dfree(dg_density); dfree(dg_encoded);
		dg_positions = (float*)dalloc(sizeof(float) * 3 * n_dg);
		dg_indices = (uint32_t*)dalloc(sizeof(uint32_t) * n_dg);
		dg_density = (__half*)dalloc(sizeof(__half) * n_dg);
		dg_encoded = (__half*)dalloc(sizeof(__half) * N_ENC * n_dg);
		density_grid_cells = n_cells;
	}
	NGPB_CUDA_CHECK(cudaMemsetAsync(dg_positions, 0, sizeof(float) * 3 * n_dg, stream));

	opt = opt_hyper; // fresh optimizer state with the configured hyper-parameters (Trainer ctor, tcnn trainer.h:53-99)
	opt.step = 0; opt.lr_factor = 1.0f;
	training_step = 0;
	density_grid_ema_step = 0;
	rays_per_batch = 1u << 12; // testbed.h:374
	measured_batch_size = measured_batch_size_before_compaction = 0;
	n_rays_total = 0;
	loss_scalar = 0.f;
	if (!host_readback) { NGPB_CUDA_CHECK(cudaMallocHost(&host_readback, 64)); std::memset(host_readback, 0, 64); }
	host_readback[8] = 0;
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
}

void ngpb_testbed::ensure_workspace(uint32_t batch) {
	if (batch == ws_batch) return;
	drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	if (batch == 0 || batch % 128 != 0) throw std::runtime_error("training batch size must be a non-zero multiple of 128 (tcnn batch_size_granularity)");
	dfree(ray_indices); dfree(rays); dfree(numsteps); dfree(coords); dfree(rgbsigma); dfree(encoded); dfree(encoded_compacted); dfree(coords_compacted);
	dfree(dloss); dfree(denc); dfree(loss); dfree(scratch); dfree(counters); dfree(partials);
	dfree(dL_dsh); dfree(coords_gradient); dL_dsh = nullptr; coords_gradient = nullptr;
	const size_t max_rays = 1u << 18;            // rays_per_batch cap (testbed_nerf.cu:2891)
	const size_t max_samples = (size_t)batch * 16; // testbed_nerf.cu:3140
	ray_indices = (uint32_t*)dalloc(sizeof(uint32_t) * max_rays);
	rays = (float*)dalloc(sizeof(float) * 6 * max_rays);
	numsteps = (uint32_t*)dalloc(sizeof(uint32_t) * 2 * max_rays);
	coords = (float*)dalloc(sizeof(float) * COORD_FLOATS * max_samples);
	rgbsigma = (__half*)dalloc(sizeof(__half) * 4 * max_samples);
	encoded = (__half*)dalloc(sizeof(__half) * N_ENC * max_samples);
	coords_compacted = (float*)dalloc(sizeof(float) * COORD_FLOATS * batch);
	encoded_compacted = (__half*)dalloc(sizeof(__half) * N_ENC * batch);
	dloss = (__half*)dalloc(sizeof(__half) * 4 * batch);
	denc = (__half*)dalloc(sizeof(__half) * N_ENC * batch);
	loss = (float*)dalloc(sizeof(float) * max_rays);
	scratch = dalloc((size_t)std::max(ngpb_generate_training_samples_scratch_bytes((uint32_t)max_rays), ngpb_compute_loss_scratch_bytes((uint32_t)max_rays)));
	counters = (uint32_t*)dalloc(sizeof(uint32_t) * 16); // [0..1] K1, [2] K6, [4] loss sum, [8..10] all-reduced copies (data parallel)
	partials = (float*)dalloc((size_t)ngpb_nerf_mlp_workspace_bytes());
	// keep the padding rows of the sample buffers finite
	NGPB_CUDA_CHECK(cudaMemsetAsync(coords, 0, sizeof(float) * COORD_FLOATS * max_samples, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(coords_compacted, 0, sizeof(float) * COORD_FLOATS * batch, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(dloss, 0, sizeof(__half) * 4 * batch, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(encoded_compacted, 0, sizeof(__half) * N_ENC * batch, stream));
	NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, sizeof(uint32_t) * 16, stream));
	ws_batch = batch;
}

// update_density_grid_nerf + update_density_grid_mean_and_bitfield (src/testbed_nerf.cu:2761-2859)
void ngpb_testbed::update_density_grid(uint32_t n_uniform, uint32_t n_nonuniform) {
	const uint32_t n_elements = NERF_GRID_CELLS * (max_cascade + 1);
	const uint32_t n_samples = n_uniform + n_nonuniform;
	auto check = [](int st) { if (st != 0) throw std::runtime_error(ngpb_last_error()); };
	if (training_step == 0 || n_images_for_training != n_images_for_training_prev) { // (:2783-2799)
		n_images_for_training_prev = n_images_for_training;
		if (training_step == 0) density_grid_ema_step = 0;
		check(ngpb_mark_untrained_density_grid(stream, n_elements, density_grid, n_images_for_training, images_dev, training_step == 0 ? 1 : 0));
		n_launches += 1;
	}
	ngpb_rng r{density_grid_rng.state, density_grid_rng.inc};
	check(ngpb_generate_grid_samples(stream, n_uniform, r, density_grid_ema_step, aabb, density_grid, dg_positions, dg_indices, max_cascade + 1, -0.01f));
	density_grid_rng.advance();
	r = ngpb_rng{density_grid_rng.state, density_grid_rng.inc};
	check(ngpb_generate_grid_samples(stream, n_nonuniform, r, density_grid_ema_step, aabb, density_grid, dg_positions + (size_t)n_uniform * 3, dg_indices + n_uniform, max_cascade + 1, NERF_MIN_OPTICAL_THICKNESS));
	density_grid_rng.advance();
	// NerfNetwork::density with the training parameters (use_inference_params = false, :2833)
	const uint32_t n_padded = next_multiple(n_samples, 128);
	hash_encode_forward_launch(stream, &grid, w_half + MLP_PARAMS, dg_positions, 3, n_padded, nullptr, dg_encoded, features_tiled());
	nerf_density_mlp_launch(stream, w_half, dg_encoded, features_tiled(), n_padded, dg_density);
	check(ngpb_splat_and_ema(stream, n_samples, dg_indices, (const ngpb_half*)dg_density, density_grid_tmp, n_elements, density_grid_decay, density_grid));
	++density_grid_ema_step;
	check(ngpb_update_bitfield(stream, max_cascade + 1, density_grid, mean_density, bitfield));
	n_launches += (n_uniform ? 1 : 0) + (n_nonuniform ? 1 : 0) + 2 + 2 + 3 + (NERF_CASCADES - 1);
}

// Sample budget of the next step's K1 (the reference uses last step's uncompacted count, :3185-3190). With several ranks the count known
// here is the average over the ranks, and a rank's own shard can need more: a 25 % margin keeps a shard from dropping rays.
uint32_t ngpb_testbed::inference_budget(uint32_t measured_before_compaction) const {
	return dp_world > 1 ? measured_before_compaction + measured_before_compaction / 4 + 1024 : measured_before_compaction;
}

void ngpb_testbed::init_data_parallel(int rank, int world, const void* unique_id128) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("init_data_parallel: invalid rank / world size");
	drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	if (nccl_comm) { NcclApi::get().CommDestroy(nccl_comm); nccl_comm = nullptr; }
	dp_rank = rank; dp_world = world;
	if (world == 1) return;
	if (!unique_id128) throw std::runtime_error("init_data_parallel: the NCCL unique id is required for world > 1");
	NcclApi& nccl = NcclApi::get();
	NcclApi::UniqueId id;
	std::memcpy(&id, unique_id128, sizeof(id));
	nccl.check(nccl.CommInitRank(&nccl_comm, world, id, rank), "ncclCommInitRank");
}

// Parameters per rank of the sharded optimizer: ceil(n_params / world), rounded up to the vector width of the optimizer kernels.
uint32_t ngpb_testbed::dp_shard_count() const { return next_multiple(div_round_up(n_params, (uint32_t)dp_world), 8u); }

// Allocates the receive buffer and flags, exchanges the CUDA IPC handles of {receive buffer, flags, fp16 weights} over the NCCL communicator and maps
// every peer's buffers. Called by all ranks at the same point (first data-parallel step after the parameters were (re)allocated).
void ngpb_testbed::p2p_setup() {
	p2p_teardown();
	if (dp_world > 16) throw std::runtime_error("peer-memory exchange supports up to 16 ranks");
	const uint32_t count = dp_shard_count();
	NGPB_CUDA_CHECK(cudaMalloc(&p2p_recv, sizeof(uint16_t) * (size_t)count * dp_world));
	NGPB_CUDA_CHECK(cudaMalloc(&p2p_flags, sizeof(uint32_t) * (2 * dp_world + 2)));
	NGPB_CUDA_CHECK(cudaMemset(p2p_flags, 0, sizeof(uint32_t) * (2 * dp_world + 2)));
	cudaIpcMemHandle_t mine[3];
	NGPB_CUDA_CHECK(cudaIpcGetMemHandle(&mine[0], p2p_recv));
	NGPB_CUDA_CHECK(cudaIpcGetMemHandle(&mine[1], p2p_flags));
	NGPB_CUDA_CHECK(cudaIpcGetMemHandle(&mine[2], w_half));
	const size_t blob = sizeof(mine);
	uint8_t* dev_blobs = nullptr;
	NGPB_CUDA_CHECK(cudaMalloc(&dev_blobs, blob * dp_world));
	NGPB_CUDA_CHECK(cudaMemcpy(dev_blobs + blob * dp_rank, mine, blob, cudaMemcpyHostToDevice));
	NGPB_CUDA_CHECK(cudaDeviceSynchronize()); // (the copy's DMA may still be in flight when cudaMemcpy returns; `stream` does not wait for the default stream)
	NcclApi& nccl = NcclApi::get();
	nccl.check(nccl.AllGather(dev_blobs + blob * dp_rank, dev_blobs, blob, 1 /* ncclUint8 */, nccl_comm, stream), "ncclAllGather(ipc handles)");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	std::vector<cudaIpcMemHandle_t> all((size_t)3 * dp_world);
	NGPB_CUDA_CHECK(cudaMemcpy(all.data(), dev_blobs, blob * dp_world, cudaMemcpyDeviceToHost));
	cudaFree(dev_blobs);
	for (int r = 0; r < dp_world; ++r) {
		if (r == dp_rank) { p2p_table[r][0] = p2p_recv; p2p_table[r][1] = p2p_flags; p2p_table[r][2] = w_half; continue; }
		for (int k = 0; k < 3; ++k) {
			void* p = nullptr;
			NGPB_CUDA_CHECK(cudaIpcOpenMemHandle(&p, all[(size_t)3 * r + k], cudaIpcMemLazyEnablePeerAccess));
			p2p_opened.push_back(p);
			p2p_table[r][k] = p;
		}
	}
	p2p_step = 0;
	// nobody may publish into a peer's flags before that peer has zeroed them: a barrier through the communicator
	nccl.check(nccl.AllReduce(p2p_flags + 2 * dp_world, p2p_flags + 2 * dp_world, 1, NcclApi::Uint32, NcclApi::Sum, nccl_comm, stream), "ncclAllReduce(barrier)");
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	p2p_ready = true;
}

void ngpb_testbed::p2p_teardown() {
	for (void* p : p2p_opened) cudaIpcCloseMemHandle(p);
	p2p_opened.clear();
	if (p2p_recv) cudaFree(p2p_recv);
	if (p2p_flags) cudaFree(p2p_flags);
	p2p_recv = nullptr; p2p_flags = nullptr;
	p2p_ready = false;
}

// Launches K1 for the step described by `p` on stream `st`.
void ngpb_testbed::launch_sampling(cudaStream_t st, const SamplingRequest& p) {
	stage_begin(NGPB_STAGE_SAMPLING, st);
	// this rank's shard: local rays [rank * n_rays, (rank + 1) * n_rays) of a global batch of world * n_rays rays
	if (ngpb_generate_training_samples_cdf(st, p.n_rays, (uint32_t)dp_rank * p.n_rays, (uint32_t)dp_world * p.n_rays, aabb, p.max_inference, p.rng,
		n_images_for_training, images_dev, bitfield, p.snap, p.cone_angle, counters, ray_indices, rays, numsteps, coords, scratch, &p.cdf) != 0) throw std::runtime_error(ngpb_last_error());
	if (p.cdf.cdf_x_cond_y || p.cdf.cdf_img) error_cdf_used = true;
	stage_end(NGPB_STAGE_SAMPLING, p.n_rays, st);
	n_launches += 5;
}

void ngpb_testbed::drop_prefetch() {
	join_ema(); // (every caller is about to touch state the side streams may still use)
	if (!prefetch_valid) return;
	NGPB_CUDA_CHECK(cudaStreamSynchronize(sampling_stream)); // its outputs are about to be overwritten or freed
	prefetch_valid = false;
}

void ngpb_testbed::collect_loss_scalar() {
	if (!loss_pending) return;
	NGPB_CUDA_CHECK(cudaEventSynchronize(loss_ready));
	float sum; std::memcpy(&sum, &host_readback[8], 4);
	loss_scalar = sum * loss_pending_scale;
	loss_pending = false;
}

// Testbed::train (src/testbed.cu:2527-2588): exactly one optimizer step.
//
// Stream schedule. The reference runs every stage back to back on one stream and reads two counters back at the end of the step
// (NerfCounters::update_after_training, testbed_nerf.cu:2870-2894) to size the next step's ray batch. Here the read-back sits right
// after the loss kernels, which is where the counters are final, and ray generation + marching for the NEXT step (K1: it depends on
// the occupancy grid, the RNG and the ray count, not on the weights) is launched on a second stream at that point, so that this
// latency-bound kernel overlaps the bandwidth-bound second half of the step (forward/backward on the compacted batch, optimizer).
// Steps that refresh the occupancy grid first are not prefetched. Results are identical to the serial schedule.
void ngpb_testbed::train(uint32_t batch, bool sync_at_end) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	if (!training_data_available) throw std::runtime_error("train: no training data loaded (a snapshot alone can be rendered, not trained)");
	if (n_images_for_training == 0) return; // (train_nerf :2897, training_prep_nerf :3389: an empty dataset trains nothing)
	for (uint32_t i = 0; i < n_images_for_training; ++i)
		if (images[i].w <= 0 || images[i].h <= 0) throw std::runtime_error("train: a training image among the first n_images_for_training has not been set");
	if (batch != ws_batch) drop_prefetch();
	ensure_workspace(batch);
	auto check = [](int st) { if (st != 0) throw std::runtime_error(ngpb_last_error()); };

	// training_prep_nerf cadence (:2538-2554, testbed_nerf.cu:3388-3401)
	const uint32_t n_prep_to_skip = std::max(1u, std::min(training_step / 16u, 16u));
	if (training_step % n_prep_to_skip == 0) {
		drop_prefetch();
		const uint32_t n_cascades = max_cascade + 1;
		stage_begin(NGPB_STAGE_DENSITY_GRID, stream);
		if (training_step < 256) update_density_grid(NERF_GRID_CELLS * n_cascades, 0);
		else update_density_grid(NERF_GRID_CELLS / 4 * n_cascades, NERF_GRID_CELLS / 4 * n_cascades);
		stage_end(NGPB_STAGE_DENSITY_GRID, training_step < 256 ? NERF_GRID_CELLS * n_cascades : NERF_GRID_CELLS / 2 * n_cascades, stream);
	}
	const bool get_loss_scalar = training_step % 16 == 0;

	// ---- train_nerf_step (src/testbed_nerf.cu:3138-3385) ----
	const uint32_t max_samples = batch * 16;
	uint32_t max_inference;
	if (measured_batch_size_before_compaction == 0) {
		measured_batch_size_before_compaction = max_inference = max_samples;
	} else {
		max_inference = next_multiple(std::min(inference_budget(measured_batch_size_before_compaction), max_samples), 128);
	}
	if (training_step == 0) n_rays_total = 0;
	n_rays_total += rays_per_batch;
	// Development aid (NGPB_POISON = bit mask): overwrite per-step buffers with 0xFF (NaN patterns) before the step, so that any read of a value the
	// step itself did not write shows up as NaN in the loss / parameters instead of hiding behind stale data of an earlier step or process.
	static const uint32_t poison = getenv("NGPB_POISON") ? (uint32_t)strtoul(getenv("NGPB_POISON"), nullptr, 0) : 0u;
	if (poison) {
		drop_prefetch();
		const size_t max_rays = 1u << 18;
		if (poison & 1u) NGPB_CUDA_CHECK(cudaMemsetAsync(encoded, 0xFF, sizeof(__half) * N_ENC * max_samples, stream));
		if (poison & 2u) NGPB_CUDA_CHECK(cudaMemsetAsync(rgbsigma, 0xFF, sizeof(__half) * 4 * max_samples, stream));
		if (poison & 4u) NGPB_CUDA_CHECK(cudaMemsetAsync(coords, 0xFF, sizeof(float) * COORD_FLOATS * max_samples, stream));
		if (poison & 8u) NGPB_CUDA_CHECK(cudaMemsetAsync(encoded_compacted, 0xFF, sizeof(__half) * N_ENC * batch, stream));
		if (poison & 16u) NGPB_CUDA_CHECK(cudaMemsetAsync(coords_compacted, 0xFF, sizeof(float) * COORD_FLOATS * batch, stream));
		if (poison & 32u) NGPB_CUDA_CHECK(cudaMemsetAsync(dloss, 0xFF, sizeof(__half) * 4 * batch, stream));
		if (poison & 64u) NGPB_CUDA_CHECK(cudaMemsetAsync(denc, 0xFF, sizeof(__half) * N_ENC * batch, stream));
		if (poison & 128u) NGPB_CUDA_CHECK(cudaMemsetAsync(partials, 0xFF, (size_t)ngpb_nerf_mlp_workspace_bytes(), stream));
		if (poison & 256u) NGPB_CUDA_CHECK(cudaMemsetAsync(scratch, 0xFF, (size_t)std::max(ngpb_generate_training_samples_scratch_bytes((uint32_t)max_rays), ngpb_compute_loss_scratch_bytes((uint32_t)max_rays)), stream));
		if (poison & 512u) NGPB_CUDA_CHECK(cudaMemsetAsync(loss, 0xFF, sizeof(float) * max_rays, stream));
		if (poison & 1024u) { NGPB_CUDA_CHECK(cudaMemsetAsync(rays, 0xFF, sizeof(float) * 6 * max_rays, stream)); NGPB_CUDA_CHECK(cudaMemsetAsync(numsteps, 0xFF, sizeof(uint32_t) * 2 * max_rays, stream)); NGPB_CUDA_CHECK(cudaMemsetAsync(ray_indices, 0xFF, sizeof(uint32_t) * max_rays, stream)); }
	}
	if (n_steps_since_cam_update == 0) NGPB_CUDA_CHECK(cudaMemsetAsync(cam_gradients, 0, sizeof(float) * 9 * images.size(), stream)); // :2916-2919
	if (optimize_exposure) exposure_active = true;
	if (n_steps_since_error_map_update == 0) begin_error_map_window(); // (:2933-2939)
	const uint32_t R = rays_per_batch;
	const ngpb_error_cdf step_cdf = sampling_cdf();
	const SamplingRequest req{training_step, R, max_inference, ngpb_rng{rng.state, rng.inc}, loss_cfg.snap_to_pixel_centers, cone_angle_constant, step_cdf};
	const ngpb_rng r = req.rng;

	// the uncompacted sample count of this step stays on the device; for the per-stage accounting the previous step's is used
	const uint64_t n_uncompacted_est = std::min(measured_batch_size_before_compaction, max_inference);
	if (prefetch_valid && prefetch == req) {
		NGPB_CUDA_CHECK(cudaStreamWaitEvent(stream, prefetch_done, 0)); // K1 of this step already ran (or is running) on the sampling stream
		prefetch_valid = false;
	} else {
		drop_prefetch();
		launch_sampling(stream, req);
	}
	// network inference on the uncompacted samples; the sample count stays on the device
	stage_begin(NGPB_STAGE_ENCODE_INFERENCE, stream);
	const bool tiled = features_tiled();
	hash_encode_forward_launch(stream, &grid, w_half + MLP_PARAMS, coords, COORD_FLOATS, max_inference, counters, encoded, tiled);
	stage_end(NGPB_STAGE_ENCODE_INFERENCE, n_uncompacted_est, stream);
	stage_begin(NGPB_STAGE_MLP_INFERENCE, stream);
	nerf_mlp_forward_launch(stream, w_half, encoded, tiled, coords, max_inference, counters, rgbsigma);
	stage_end(NGPB_STAGE_MLP_INFERENCE, n_uncompacted_est, stream);
	stage_begin(NGPB_STAGE_LOSS, stream);
	check(compute_loss_launch(stream, R, (uint32_t)dp_world * R, aabb, r, batch, &loss_cfg, n_images_for_training, images_dev, counters, (const ngpb_half*)rgbsigma,
		ray_indices, rays, numsteps, coords, mean_density, coords_compacted, (ngpb_half*)dloss, loss, counters + 2, scratch,
		reuse_encoding ? (const ngpb_half*)encoded : nullptr, reuse_encoding ? (ngpb_half*)encoded_compacted : nullptr, tiled,
		exposure_active ? cam_exposure : nullptr, optimize_exposure ? cam_gradients + 6 * images.size() : nullptr,
		ErrorCdf{step_cdf.cdf_x_cond_y, step_cdf.cdf_y, step_cdf.cdf_img, step_cdf.res_x, step_cdf.res_y}, error_map_live ? error_map : nullptr, error_map_res[0], error_map_res[1]));
	stage_end(NGPB_STAGE_LOSS, R, stream);
	n_launches += 2 + 5 + (optimize_exposure ? 1 : 0) + (error_map_live ? 1 : 0); // encode, mlp; loss target / composite / scan / gradient / rollover (+ exposure gradient, + error-map deposit)
	if (dp_world > 1) {
		// the controller needs the GLOBAL sample counts so that every rank derives the same next ray count: sum {uncompacted, kept rays,
		// compacted} into counters[8..10] (the local values stay in [0..2] for the kernels of this step)
		NcclApi& nccl = NcclApi::get();
		nccl.check(nccl.AllReduce(counters, counters + 8, 3, NcclApi::Uint32, NcclApi::Sum, nccl_comm, stream), "ncclAllReduce(counters)");
	}

	// ---- the two counters of NerfCounters::update_after_training (:2870-2894) are final here: start their read-back ----
	collect_loss_scalar(); // (frees the read-back slot of an earlier step's loss)
	NGPB_CUDA_CHECK(cudaMemcpyAsync(host_readback, counters + (dp_world > 1 ? 8 : 0), sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, stream));
	NGPB_CUDA_CHECK(cudaEventRecord(counters_ready, stream));
	d2h_bytes += 16;

	// ---- second half of the step, enqueued without waiting for the read-back (nothing in it depends on the host) ----
	// forward + backward on the compacted, padded batch
	// The reference re-encodes the compacted samples (NerfNetwork::forward). Inference and training use the same fp16 weights within a step, so the
	// features the inference pass produced for these very samples are bit-identical: the loss stage compacted them along with the coordinates.
	if (!reuse_encoding) {
		stage_begin(NGPB_STAGE_ENCODE_TRAIN, stream);
		hash_encode_forward_launch(stream, &grid, w_half + MLP_PARAMS, coords_compacted, COORD_FLOATS, batch, nullptr, encoded_compacted, tiled);
		stage_end(NGPB_STAGE_ENCODE_TRAIN, batch, stream);
		++n_launches;
	}
	stage_begin(NGPB_STAGE_MLP_TRAIN, stream);
	if (optimize_extrinsics && !dL_dsh) {
		dL_dsh = (__half*)dalloc(sizeof(__half) * 16 * batch);
		coords_gradient = (float*)dalloc(sizeof(float) * COORD_FLOATS * batch);
	}
	nerf_mlp_forward_backward_launch(stream, w_half, encoded_compacted, tiled, coords_compacted, dloss, batch, denc, grad, partials, optimize_extrinsics ? dL_dsh : nullptr);
	stage_end(NGPB_STAGE_MLP_TRAIN, batch, stream);
	if (!optimize_extrinsics) NGPB_CUDA_CHECK(cudaEventRecord(mlp_train_done, stream));
	if (optimize_extrinsics) {
		// K13 / K14 (train_nerf_step :3324-3372): gradients w.r.t. the network inputs of the compacted batch, reduced per camera. The reference runs the
		// input gradient over the padded batch too; padding samples carry a zero loss gradient and belong to no ray.
		nerf_input_gradient_launch(stream, &grid, w_half + MLP_PARAMS, coords_compacted, batch, denc, dL_dsh, coords_gradient);
		cam_gradient_launch(stream, R, (uint32_t)dp_world * R, aabb, counters + 1, n_images_for_training, ray_indices, rays, numsteps, coords_compacted, coords_gradient,
			cam_gradients, cam_gradients + 3 * images.size(), step_cdf.cdf_img);
		n_launches += 2;
		// the camera-gradient kernel is the last reader of this step's rays / numsteps / ray counter: the prefetched sampling of the next step, which
		// rewrites them on the sampling stream, has to wait for it
		NGPB_CUDA_CHECK(cudaEventRecord(mlp_train_done, stream));
	}
	if (dp_world == 1) {
		stage_begin(NGPB_STAGE_ENCODE_BACKWARD, stream);
		hash_encode_backward_launch(stream, &grid, coords_compacted, COORD_FLOATS, batch, denc, grad + MLP_PARAMS, 0, grid.n_levels);
		stage_end(NGPB_STAGE_ENCODE_BACKWARD, batch, stream);
		// optimizer (train_nerf :2950)
		stage_begin(NGPB_STAGE_OPTIMIZER, stream);
		check(ngpb_optimizer_step(stream, &opt, n_params, MLP_PARAMS, LOSS_SCALE, grad, w_fp32, (ngpb_half*)w_half, (ngpb_half*)w_ema, m1, m2, param_steps));
		stage_end(NGPB_STAGE_OPTIMIZER, n_params, stream);
	} else {
		// Data parallel: the one exchange step of the path. The sum of the shards' gradients is the gradient of the global batch (the loss is
		// normalised by the global ray count). Default exchange: partial gradients rounded to bf16, reduce-scattered, Adam on this rank's slice, fp16
		// weights all-gathered (dp_half_gradients = 0 keeps the exchange in fp32; tools/dp_equivalence.py bounds the difference: bf16 vs fp32 within 3 %
		// of the loss, both within 10 % of single-GPU training of the same global batch). Every rank receives the same bits, so replicas stay identical.
		stage_begin(NGPB_STAGE_ENCODE_BACKWARD, stream);
		hash_encode_backward_launch(stream, &grid, coords_compacted, COORD_FLOATS, batch, denc, grad + MLP_PARAMS, 0, grid.n_levels);
		stage_end(NGPB_STAGE_ENCODE_BACKWARD, batch, stream);
		NcclApi& nccl = NcclApi::get();
		uint8_t opt_params[256];
		optimizer_prepare(&opt, LOSS_SCALE, opt_params);
		if (dp_sharded_optimizer && dp_exchange == 1) {
			if (!p2p_ready) p2p_setup();
			const uint32_t count = dp_shard_count(), first = (uint32_t)dp_rank * count;
			const uint32_t mine = first < n_params ? std::min(count, n_params - first) : 0u;
			const uint32_t n_cast = count * (uint32_t)dp_world;
			if (n_cast > n_alloc) throw std::runtime_error("data parallel: parameter padding too small for this world size");
			P2pTable T;
			static_assert(sizeof(T.ptr) == sizeof(p2p_table), "table layout");
			std::memcpy(T.ptr, p2p_table, sizeof(p2p_table));
			if (host_readback[12] == 0xDEADu) throw std::runtime_error("data parallel: a peer's gradients or weights did not arrive (peer-memory exchange timed out)");
			const uint32_t step = ++p2p_step;
			stage_begin(NGPB_STAGE_ALLREDUCE, stream);
			p2p_scatter_gradients_kernel<<<div_round_up(n_cast / 8, 256), 256, 0, stream>>>(n_cast / 8, count, (uint32_t)dp_rank, (uint32_t)dp_world, step, reinterpret_cast<float4*>(grad), T, p2p_flags);
			NGPB_LAUNCH_CHECK();
			if (mine) { p2p_reduce_gradients_kernel<<<div_round_up(div_round_up(mine, 2u), 256), 256, 0, stream>>>(mine, count, (uint32_t)dp_world, step, p2p_recv, p2p_flags, grad + first); NGPB_LAUNCH_CHECK(); }
			stage_end(NGPB_STAGE_ALLREDUCE, (uint64_t)n_params * 2, stream);
			stage_begin(NGPB_STAGE_OPTIMIZER, stream);
			optimizer_disable_fused_ema(opt_params);
			join_ema(); // the previous step's EMA sweep still reads the fp16 weights this launch overwrites
			optimizer_launch(stream, opt_params, first, mine, MLP_PARAMS, grad, w_fp32, w_half, w_ema, m1, m2, param_steps);
			p2p_gather_weights_kernel<<<std::max(1u, div_round_up(count / 8, 256)), 256, 0, stream>>>(count / 8, first, (uint32_t)dp_rank, (uint32_t)dp_world, step, w_half, T, p2p_flags);
			NGPB_LAUNCH_CHECK();
			p2p_wait_weights_kernel<<<1, 32, 0, stream>>>((uint32_t)dp_world, step, p2p_flags);
			NGPB_LAUNCH_CHECK();
			launch_ema_sweep_async(opt_params); // (on its own stream: only renders and snapshots read the EMA copy)
			stage_end(NGPB_STAGE_OPTIMIZER, n_params, stream);
			NGPB_CUDA_CHECK(cudaMemcpyAsync(host_readback + 12, p2p_flags + 2 * dp_world + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)); // time-out word
			master_weights_sharded = true;
			n_launches += 5;
		} else if (dp_sharded_optimizer) {
			// reduce-scatter the gradients, run Adam on this rank's 1/world of the parameters, all-gather the updated fp16 weights; the EMA
			// copy follows from the gathered weights on every rank. Moves 3/4 of the all-reduce's bytes and divides the optimizer sweep by world.
			const uint32_t count = dp_shard_count(), first = (uint32_t)dp_rank * count;
			const uint32_t mine = first < n_params ? std::min(count, n_params - first) : 0u;
			stage_begin(NGPB_STAGE_ALLREDUCE, stream);
			if (dp_half_gradients) {
				// Exchange the gradients in 16 bits, which halves the bytes on NVLink. The reference keeps its gradients in fp16 everywhere (tcnn
				// accumulates them with atomicAdd(__half2) and hands fp16 to Adam); a rank's PARTIAL gradient is world x smaller than that, deep in
				// fp16's subnormal range for sparsely hit grid entries, so the default is bf16 (fp32's range, 2^-9 relative rounding per partial sum;
				// dp_half_gradients = 2 selects fp16). One pass casts the fp32 accumulation buffer and resets it; the reduced slice is widened again for
				// the fp32 optimizer. Every rank receives the same bits, so replicas stay identical.
				if (!grad_half) grad_half = (__half*)dalloc(sizeof(__half) * ((size_t)count * dp_world + count));
				__half* shard_half = grad_half + (size_t)count * dp_world;
				const uint32_t n_cast = count * (uint32_t)dp_world; // n_alloc >= n_params + 4096 covers the padding of the last slice
				if (n_cast > n_alloc) throw std::runtime_error("data parallel: parameter padding too small for this world size");
				const bool bf16 = dp_half_gradients == 1;
				if (bf16) grads_to_16bit_and_reset_kernel<true><<<div_round_up(n_cast / 4, 256), 256, 0, stream>>>(n_cast / 4, reinterpret_cast<float4*>(grad), reinterpret_cast<uint2*>(grad_half));
				else grads_to_16bit_and_reset_kernel<false><<<div_round_up(n_cast / 4, 256), 256, 0, stream>>>(n_cast / 4, reinterpret_cast<float4*>(grad), reinterpret_cast<uint2*>(grad_half));
				NGPB_LAUNCH_CHECK();
				nccl.check(nccl.ReduceScatter(grad_half, shard_half, count, bf16 ? NcclApi::Bfloat16 : NcclApi::Float16, NcclApi::Sum, nccl_comm, stream), "ncclReduceScatter(gradients, 16 bit)");
				if (mine) {
					if (bf16) widen_16bit_kernel<true><<<div_round_up(mine, 256), 256, 0, stream>>>(mine, reinterpret_cast<const uint16_t*>(shard_half), grad + first);
					else widen_16bit_kernel<false><<<div_round_up(mine, 256), 256, 0, stream>>>(mine, reinterpret_cast<const uint16_t*>(shard_half), grad + first);
					NGPB_LAUNCH_CHECK();
				}
				n_launches += 2;
			} else {
				nccl.check(nccl.ReduceScatter(grad, grad + first, count, NcclApi::Float32, NcclApi::Sum, nccl_comm, stream), "ncclReduceScatter(gradients)");
			}
			stage_end(NGPB_STAGE_ALLREDUCE, (uint64_t)n_params * (dp_half_gradients ? 2 : 4), stream);
			stage_begin(NGPB_STAGE_OPTIMIZER, stream);
			optimizer_disable_fused_ema(opt_params);
			join_ema(); // the previous step's EMA sweep still reads the fp16 weights this launch overwrites
			optimizer_launch(stream, opt_params, first, mine, MLP_PARAMS, grad, w_fp32, w_half, w_ema, m1, m2, param_steps);
			// fp32 exchange: the other ranks' ranges still hold this rank's partial sums (the fp16 path reset the buffer while casting)
			if (!dp_half_gradients) NGPB_CUDA_CHECK(cudaMemsetAsync(grad, 0, sizeof(float) * n_alloc, stream));
			nccl.check(nccl.AllGather(w_half + first, w_half, count, NcclApi::Float16, nccl_comm, stream), "ncclAllGather(weights)");
			launch_ema_sweep_async(opt_params); // (on its own stream: only renders and snapshots read the EMA copy)
			stage_end(NGPB_STAGE_OPTIMIZER, n_params, stream);
			master_weights_sharded = true;
			n_launches += 3;
		} else {
			stage_begin(NGPB_STAGE_ALLREDUCE, stream);
			nccl.check(nccl.AllReduce(grad, grad, n_params, NcclApi::Float32, NcclApi::Sum, nccl_comm, stream), "ncclAllReduce(gradients)");
			stage_end(NGPB_STAGE_ALLREDUCE, (uint64_t)n_params * 4, stream);
			stage_begin(NGPB_STAGE_OPTIMIZER, stream);
			join_ema();
			optimizer_launch(stream, opt_params, 0, n_params, MLP_PARAMS, grad, w_fp32, w_half, w_ema, m1, m2, param_steps);
			stage_end(NGPB_STAGE_OPTIMIZER, n_params, stream);
			n_launches += 1;
		}
	}
	n_launches += 2 + 1 + 1; // mlp_train + reduce_partials, encode_backward, optimizer
	// loss scalar every 16th step (:2885-2888): reduced on the device, read back without stalling the stream
	if (get_loss_scalar) {
		NGPB_STEP_KERNEL(sum_kernel);
		sum_kernel<<<1, 1024, 0, stream>>>(loss, R, reinterpret_cast<float*>(counters + 4));
		NGPB_LAUNCH_CHECK();
		n_launches += 1;
		if (dp_world > 1) {
			NcclApi& nccl = NcclApi::get();
			nccl.check(nccl.AllReduce(counters + 4, counters + 4, 1, NcclApi::Float32, NcclApi::Sum, nccl_comm, stream), "ncclAllReduce(loss)");
		}
		NGPB_CUDA_CHECK(cudaMemcpyAsync(host_readback + 8, counters + 4, sizeof(float), cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaEventRecord(loss_ready, stream));
		d2h_bytes += 4;
	}
	++training_step;
	rng.advance(); // m_rng.advance() (:3380)
	++n_steps_since_cam_update; // (:3026)
	if ((optimize_extrinsics || optimize_exposure) && n_steps_since_cam_update >= n_steps_between_cam_updates) camera_update_step();
	// CDFs from the error map (:2970-3023); enqueued on the training stream, after this step's loss kernels and before the next step's sampling
	++n_steps_since_error_map_update;
	bool cdf_rebuilt = false;
	if (n_steps_since_error_map_update >= n_steps_between_error_map_updates) { cdf_rebuilt = error_map_live; finish_error_map_window(); }

	// ---- batch-size controller (:2870-2894) ----
	NGPB_CUDA_CHECK(cudaEventSynchronize(counters_ready));
	const uint32_t counter_cpu = host_readback[0], compacted_counter_cpu = host_readback[2];
	measured_batch_size = 0;
	measured_batch_size_before_compaction = 0;
	if (counter_cpu == 0 || compacted_counter_cpu == 0) {
		// "Nerf training generated 0 samples. Aborting training." (:2964-2968)
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		loss_scalar = 0.f;
		loss_pending = false;
		shall_train = false;
		if (profile_stages) stage_collect();
		return;
	}
	// per-rank averages of the global counts: with one rank these are the reference's values; with several, every rank sees the same numbers
	measured_batch_size_before_compaction = counter_cpu / (uint32_t)dp_world;
	measured_batch_size = std::max(1u, compacted_counter_cpu / (uint32_t)dp_world);
	if (get_loss_scalar) { loss_pending = true; loss_pending_scale = (float)measured_batch_size / (float)batch; }
	rays_per_batch = ngpb_next_rays_per_batch(rays_per_batch, batch, compacted_counter_cpu, (uint32_t)dp_world);

	// ---- K1 of the next step, on the sampling stream, unless that step starts with an occupancy-grid refresh ----
	{
		const uint32_t next_skip = std::max(1u, std::min(training_step / 16u, 16u));
		if (overlap_sampling && training_step % next_skip != 0) {
			prefetch = SamplingRequest{training_step, rays_per_batch, next_multiple(std::min(inference_budget(measured_batch_size_before_compaction), max_samples), 128),
				ngpb_rng{rng.state, rng.inc}, loss_cfg.snap_to_pixel_centers, cone_angle_constant, sampling_cdf()};
			static const bool after_mlp = getenv("NGPB_K1_AFTER_MLP") && atoi(getenv("NGPB_K1_AFTER_MLP")) != 0;
			if (after_mlp || optimize_extrinsics) NGPB_CUDA_CHECK(cudaStreamWaitEvent(sampling_stream, mlp_train_done, 0));
			if (cdf_rebuilt) NGPB_CUDA_CHECK(cudaStreamWaitEvent(sampling_stream, error_cdf_built, 0)); // the next step draws from the CDFs just enqueued
			launch_sampling(sampling_stream, prefetch);
			NGPB_CUDA_CHECK(cudaEventRecord(prefetch_done, sampling_stream));
			prefetch_valid = true;
		}
	}
	if (sync_at_end) {
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		collect_loss_scalar();
		if (profile_stages) stage_collect();
	}
}

// The batch-size controller's arithmetic and the ray shard of a rank, host only (used by train() above; tests/test_data_parallel.py runs them under gloo).
// NerfCounters::update_after_training (:2890-2891) on the per-rank average of the summed compacted count, in float like the reference.
extern "C" uint32_t ngpb_next_rays_per_batch(uint32_t rays_per_batch, uint32_t batch, uint32_t global_compacted, uint32_t world) {
	const uint32_t measured = std::max(1u, global_compacted / std::max(world, 1u));
	const uint32_t r = (uint32_t)((float)rays_per_batch * (float)batch / (float)measured);
	return std::min(next_multiple(r, 128u), 1u << 18);
}
// every rank marches rays [rank * rays_per_batch, (rank + 1) * rays_per_batch) of a global batch of world * rays_per_batch
extern "C" void ngpb_ray_shard(uint32_t rank, uint32_t world, uint32_t rays_per_batch, uint32_t* ray_offset, uint32_t* n_rays_global) {
	if (ray_offset) *ray_offset = rank * rays_per_batch;
	if (n_rays_global) *n_rays_global = world * rays_per_batch;
}

// ---- K19: error-map windows (train_nerf :2933-2939, :2970-3023) ----
ngpb_error_cdf ngpb_testbed::sampling_cdf() const {
	ngpb_error_cdf c{nullptr, nullptr, nullptr, 0, 0};
	if (!error_cdf_valid) return c; // (:3211-3212)
	if (sample_focal_plane_proportional_to_error) { c.cdf_x_cond_y = error_cdf_x_cond_y; c.cdf_y = error_cdf_y; c.res_x = error_cdf_res[0]; c.res_y = error_cdf_res[1]; }
	if (sample_image_proportional_to_error) c.cdf_img = error_cdf_img;
	return c;
}

// Start of a window: the map's resolution follows the number of rays an image will receive during the window, capped by the first image's resolution; the
// map is cleared. Without one of the two switches nothing would read the map, and nothing is accumulated.
void ngpb_testbed::begin_error_map_window() {
	error_map_live = false;
	if (!(sample_focal_plane_proportional_to_error || sample_image_proportional_to_error) || images.empty()) return;
	const uint32_t n_images = (uint32_t)images.size();
	const uint32_t n_samples_per_image = (n_steps_between_error_map_updates * rays_per_batch) / n_images; // (32-bit like the reference)
	const int side = (int)(std::sqrt(std::sqrt((float)n_samples_per_image)) * 3.5f);
	const int rx = std::min(side, images[0].w), ry = std::min(side, images[0].h);
	if (rx < 2 || ry < 2) return; // (the bilinear deposit touches texel + 1)
	const size_t n = (size_t)n_images * rx * ry;
	if (n > error_map_capacity) {
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream)); // the previous window's deposits
		dfree(error_map);
		error_map = (float*)dalloc(sizeof(float) * n);
		error_map_capacity = n;
	}
	error_map_res[0] = rx; error_map_res[1] = ry;
	NGPB_CUDA_CHECK(cudaMemsetAsync(error_map, 0, sizeof(float) * n, stream));
	error_map_live = true;
}

// End of a window: CDFs from the accumulated map (device kernels, incl. the image normalisation the reference runs on the host), counters reset, and the
// next window is 1.5x longer. Data parallel: every rank deposited its shard's rays; the maps are summed first, so all ranks build the same CDFs.
void ngpb_testbed::finish_error_map_window() {
	if (error_map_live) {
		const uint32_t n_images = (uint32_t)images.size();
		const size_t n = (size_t)n_images * error_map_res[0] * error_map_res[1];
		if (n > error_cdf_capacity) {
			if (error_cdf_used) { drop_prefetch(); NGPB_CUDA_CHECK(cudaStreamSynchronize(stream)); } // kernels that still read the old CDFs
			dfree(error_cdf_x_cond_y); dfree(error_cdf_y);
			error_cdf_x_cond_y = (float*)dalloc(sizeof(float) * n);
			error_cdf_y = (float*)dalloc(sizeof(float) * n); // (rows only; sized like the map so that any later aspect fits)
			error_cdf_capacity = n;
		}
		if (!error_cdf_img) { error_cdf_img = (float*)dalloc(sizeof(float) * n_images); error_pmf_img = (float*)dalloc(sizeof(float) * n_images); }
		if (dp_world > 1) {
			NcclApi& nccl = NcclApi::get();
			nccl.check(nccl.AllReduce(error_map, error_map, n, NcclApi::Float32, NcclApi::Sum, nccl_comm, stream), "ncclAllReduce(error map)");
		}
		// The sampling stream may still run the prefetched K1 of a step that reads the CDFs about to be overwritten: it is always consumed (waited for by
		// `stream`) before this point, because this runs after the step's own loss kernels.
		error_cdf_res[0] = error_map_res[0]; error_cdf_res[1] = error_map_res[1];
		if (ngpb_construct_error_cdfs(stream, n_images, (uint32_t)error_cdf_res[1], (uint32_t)error_cdf_res[0], error_map, error_cdf_x_cond_y, error_cdf_y, error_cdf_img, error_pmf_img) != 0)
			throw std::runtime_error(ngpb_last_error());
		NGPB_CUDA_CHECK(cudaEventRecord(error_cdf_built, stream));
		n_launches += 3;
		error_cdf_valid = true;
		error_cdf_used = false;
	}
	n_steps_since_error_map_update = 0;
	n_steps_between_error_map_updates = (uint32_t)((float)n_steps_between_error_map_updates * 1.5f);
	error_map_live = false;
}

// The sharded optimizer's EMA sweep (a function of the gathered fp16 weights only), off the training stream: it runs under the next step's first half.
void ngpb_testbed::launch_ema_sweep_async(const void* opt_params) {
	NGPB_CUDA_CHECK(cudaEventRecord(weights_gathered, stream));
	NGPB_CUDA_CHECK(cudaStreamWaitEvent(ema_stream, weights_gathered, 0));
	ema_sweep_launch(ema_stream, opt_params, next_multiple(n_params, 8), w_half, w_ema);
	NGPB_CUDA_CHECK(cudaEventRecord(ema_done, ema_stream));
	ema_pending = true;
}
void ngpb_testbed::join_ema() {
	if (!ema_pending) return;
	NGPB_CUDA_CHECK(cudaStreamWaitEvent(stream, ema_done, 0));
	ema_pending = false;
}

// Host half of the camera optimisation (train_nerf :3056-3083, :3134): read the accumulated gradients, one Adam step per camera and offset, new transforms.
// Synchronises the stream (as the reference does); nothing that reads the image table is in flight afterwards -- the prefetch of the next step's sampling is
// launched after this returns.
void ngpb_testbed::camera_update_step() {
	const size_t n = images.size();
	if (dp_world > 1) {
		NcclApi& nccl = NcclApi::get();
		nccl.check(nccl.AllReduce(cam_gradients, cam_gradients, 9 * n, NcclApi::Float32, NcclApi::Sum, nccl_comm, stream), "ncclAllReduce(camera gradients)");
	}
	NGPB_CUDA_CHECK(cudaMemcpyAsync(cam_gradients_host.data(), cam_gradients, sizeof(float) * 9 * n, cudaMemcpyDeviceToHost, stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	d2h_bytes += sizeof(float) * 9 * n;
	const size_t n_train = n_images_for_training; // (:3058, :3067, :3111: scale and loops run over the images in training)
	const float per_camera_loss_scale = (float)n_train / LOSS_SCALE / (float)n_steps_between_cam_updates;
	const float lr_floor = opt.learning_rate * opt.lr_factor / 1000.0f; // m_optimizer->learning_rate() / 1000
	if (optimize_exposure) { // (:3105-3131)
		ngpb_exposure_update((uint32_t)n_train, cam_exposure_state.data(), &cam_gradients_host[6 * n], per_camera_loss_scale, exposure_l2_reg, opt.learning_rate * opt.lr_factor);
		upload_exposures();
	}
	for (size_t i = 0; optimize_extrinsics && i < n_train; ++i) {
		float* ps = &cam_pos_state[i * 10]; float* rs = &cam_rot_state[i * 10];
		float pg[3], rg[3];
		for (int c = 0; c < 3; ++c) {
			pg[c] = cam_gradients_host[i * 3 + c] * per_camera_loss_scale + ps[7 + c] * extrinsic_l2_reg;
			rg[c] = cam_gradients_host[(n + i) * 3 + c] * per_camera_loss_scale + rs[7 + c] * extrinsic_l2_reg;
		}
		const float lr_p = std::max(extrinsic_learning_rate * std::pow(0.33f, (float)((uint32_t)ps[0] / 128u)), lr_floor);
		const float lr_r = std::max(extrinsic_learning_rate * std::pow(0.33f, (float)((uint32_t)rs[0] / 128u)), lr_floor);
		ngpb_camera_adam_step(ps, pg, lr_p, 0);
		ngpb_camera_adam_step(rs, rg, lr_r, 1);
	}
	if (optimize_extrinsics) update_transforms();
	n_steps_since_cam_update = 0;
}

void ngpb_testbed::upload_exposures() {
	const size_t n = images.size();
	std::vector<float> e(n * 3);
	for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) e[i * 3 + c] = cam_exposure_state[i * 10 + 7 + c];
	NGPB_CUDA_CHECK(cudaMemcpyAsync(cam_exposure, e.data(), sizeof(float) * 3 * n, cudaMemcpyHostToDevice, stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream)); // (the staging vector is about to go away)
	h2d_bytes += sizeof(float) * 3 * n;
}

// Training::update_transforms (:2597-2633): training transform = dataset transform with the camera's offsets applied; re-upload the image table.
void ngpb_testbed::update_transforms() {
	const size_t n = images.size();
	if (n == 0) return;
	drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	for (size_t i = 0; i < n; ++i) {
		ngpb_apply_camera_offsets(&dataset_xforms[i * 12], &cam_pos_state[i * 10 + 7], &cam_rot_state[i * 10 + 7], images[i].raw_xform);
		ngpb_effective_xform(images[i].raw_xform, images[i].xform);
	}
	upload_image_table();
}

void ngpb_testbed::upload_image_table() {
	const size_t n = images.size();
	NGPB_CUDA_CHECK(cudaMemcpyAsync(images_dev, images.data(), sizeof(ngpb_image) * n, cudaMemcpyHostToDevice, stream));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	h2d_bytes += sizeof(ngpb_image) * n;
}

// Testbed::create_empty_nerf_dataset (src/testbed_nerf.cu:2635-2641): n empty slots at identity poses, nothing in training yet
void ngpb_testbed::create_empty_dataset(uint32_t n, uint32_t aabb_scale_) {
	if (n == 0) throw std::runtime_error("create_empty_nerf_dataset: n_images must be positive");
	std::vector<ngpb_host_image> empty(n);
	for (auto& h : empty) {
		h = ngpb_host_image{};
		h.fx = h.fy = 1000.0f; h.cx = h.cy = 0.5f; // (nerf_loader.cu create_empty_nerf_dataset: metadata defaults)
		h.xform[0] = h.xform[4] = h.xform[8] = 1.0f;
	}
	load_training_data(n, empty.data(), aabb_scale_, true);
	n_images_for_training = n_images_for_training_prev = 0;
}

// nerf.training.set_image -> NerfDataset::set_training_image (python_api.cu:56-76): one slot's pixels, at any resolution and data type
void ngpb_testbed::set_training_image(uint32_t frame_idx, const ngpb_host_image& h) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	if (frame_idx >= images.size()) throw std::runtime_error("Invalid frame index");
	if (!h.pixels || h.w <= 0 || h.h <= 0) throw std::runtime_error("set_image: image should be (H,W,C) where C=4");
	const size_t bytes = (size_t)h.w * h.h * image_bytes_per_pixel(h.image_type);
	drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream)); // kernels of earlier steps may still read the slot
	ngpb_image& im = images[frame_idx];
	const size_t old_bytes = (size_t)im.w * im.h * image_bytes_per_pixel(im.image_type);
	if (bytes != old_bytes || !im.pixels) { // a slot of another size gets its own allocation (slots loaded in bulk live in `pixels`)
		dfree(own_pixels[frame_idx]);
		own_pixels[frame_idx] = dalloc(bytes);
		im.pixels = (const uint8_t*)own_pixels[frame_idx];
	}
	NGPB_CUDA_CHECK(cudaMemcpyAsync(const_cast<uint8_t*>(im.pixels), h.pixels, bytes, cudaMemcpyHostToDevice, stream));
	im.w = h.w; im.h = h.h; im.image_type = h.image_type;
	h2d_bytes += bytes;
	upload_image_table();
}

// Training::set_camera_intrinsics (src/testbed_nerf.cu:2502-2516)
void ngpb_testbed::set_camera_intrinsics(uint32_t frame_idx, float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	if (frame_idx >= images.size()) return; // (the reference ignores out-of-range frames)
	if (fx <= 0.f) fx = fy;
	if (fy <= 0.f) fy = fx;
	ngpb_image& im = images[frame_idx];
	if (cx < 0.f) cx = -cx; else cx = cx / (float)im.w;
	if (cy < 0.f) cy = -cy; else cy = cy / (float)im.h;
	im.lens_mode = (k1 != 0.f || k2 != 0.f || p1 != 0.f || p2 != 0.f) ? NGPB_LENS_OPENCV : NGPB_LENS_PERSPECTIVE;
	std::memset(im.lens_params, 0, sizeof(im.lens_params));
	im.lens_params[0] = k1; im.lens_params[1] = k2; im.lens_params[2] = p1; im.lens_params[3] = p2;
	im.cx = cx; im.cy = cy; im.fx = fx; im.fy = fy;
	drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	upload_image_table();
}

void ngpb_testbed::reset_camera_extrinsics() { // Training::reset_camera_extrinsics (:2543-2555)
	std::fill(cam_pos_state.begin(), cam_pos_state.end(), 0.f);
	std::fill(cam_rot_state.begin(), cam_rot_state.end(), 0.f);
	std::fill(cam_exposure_state.begin(), cam_exposure_state.end(), 0.f);
}

void ngpb_testbed::get_params(float* o_fp32, ngpb_half* o_half, ngpb_half* o_ema) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	join_ema();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
	if (o_fp32 && dp_world > 1 && master_weights_sharded) {
		// the fp32 master copy is only current in each rank's own range: gather it (a collective -- call get_params on every rank)
		NcclApi& nccl = NcclApi::get();
		const uint32_t count = dp_shard_count();
		nccl.check(nccl.AllGather(w_fp32 + (size_t)dp_rank * count, w_fp32, count, NcclApi::Float32, nccl_comm, stream), "ncclAllGather(master weights)");
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		master_weights_sharded = false;
	}
	if (o_fp32) NGPB_CUDA_CHECK(cudaMemcpy(o_fp32, w_fp32, sizeof(float) * n_params, cudaMemcpyDeviceToHost));
	if (o_half) NGPB_CUDA_CHECK(cudaMemcpy(o_half, w_half, sizeof(__half) * n_params, cudaMemcpyDeviceToHost));
	if (o_ema) NGPB_CUDA_CHECK(cudaMemcpy(o_ema, w_ema, sizeof(__half) * n_params, cudaMemcpyDeviceToHost));
}

void ngpb_testbed::set_params(const float* i_fp32) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	master_weights_sharded = false;
	NGPB_CUDA_CHECK(cudaMemcpyAsync(w_fp32, i_fp32, sizeof(float) * n_params, cudaMemcpyHostToDevice, stream));
	cast_params_kernel<<<div_round_up(n_params, 256), 256, 0, stream>>>(n_params, w_fp32, w_half);
	NGPB_LAUNCH_CHECK();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
#define NGPB_API_BEGIN try {
#define NGPB_API_END } catch (const std::exception& e) { ngpb::set_last_error(e.what()); return NGPB_ERR_RUNTIME; } return 0;

extern "C" int ngpb_testbed_create(ngpb_testbed** out, int device) {
	NGPB_API_BEGIN
	if (!out) throw std::runtime_error("ngpb_testbed_create: null output");
	if (ngpb_check_device(device) != 0) return NGPB_ERR_RUNTIME;
	*out = new ngpb_testbed(device);
	NGPB_API_END
}
extern "C" void ngpb_testbed_destroy(ngpb_testbed* t) { delete t; }
extern "C" int ngpb_testbed_load_training_data(ngpb_testbed* t, uint32_t n_images, const ngpb_host_image* images, uint32_t aabb_scale) {
	NGPB_API_BEGIN t->load_training_data(n_images, images, aabb_scale); NGPB_API_END
}
extern "C" int ngpb_testbed_reset_network(ngpb_testbed* t, uint32_t seed) {
	NGPB_API_BEGIN
	if (!t->scene_configured) throw std::runtime_error("reset_network: load training data or a snapshot first (the grid resolution depends on aabb_scale)");
	t->reset_network(seed);
	NGPB_API_END
}
extern "C" int ngpb_testbed_train(ngpb_testbed* t, uint32_t batch_size) { NGPB_API_BEGIN t->train(batch_size, true); NGPB_API_END }
extern "C" int ngpb_testbed_train_n(ngpb_testbed* t, uint32_t batch_size, uint32_t n_steps) {
	NGPB_API_BEGIN
	// steps run back to back: the only host wait inside a step is the counter read-back that sizes the next ray batch
	for (uint32_t i = 0; i < n_steps && t->shall_train; ++i) t->train(batch_size, i + 1 == n_steps);
	NGPB_API_END
}
extern "C" float ngpb_testbed_loss(ngpb_testbed* t) {
	try { cudaSetDevice(t->device); t->collect_loss_scalar(); } catch (...) {}
	return t->loss_scalar;
}
extern "C" uint32_t ngpb_testbed_training_step(const ngpb_testbed* t) { return t->training_step; }
extern "C" int ngpb_testbed_stats(ngpb_testbed* t, uint64_t* s) {
	NGPB_API_BEGIN
	s[0] = t->rays_per_batch; s[1] = t->measured_batch_size_before_compaction; s[2] = t->measured_batch_size; s[3] = t->n_launches;
	NGPB_API_END
}
extern "C" uint32_t ngpb_testbed_n_params(const ngpb_testbed* t) { return t->n_params; }
extern "C" int ngpb_testbed_get_params(ngpb_testbed* t, float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema) { NGPB_API_BEGIN t->get_params(w_fp32, w_half, w_ema); NGPB_API_END }
extern "C" int ngpb_testbed_set_params(ngpb_testbed* t, const float* w_fp32) { NGPB_API_BEGIN t->set_params(w_fp32); NGPB_API_END }
extern "C" int ngpb_testbed_get_density_grid(ngpb_testbed* t, float* grid_out, uint8_t* bitfield_out) {
	NGPB_API_BEGIN
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	if (grid_out) NGPB_CUDA_CHECK(cudaMemcpy(grid_out, t->density_grid, sizeof(float) * NERF_GRID_CELLS * (t->max_cascade + 1), cudaMemcpyDeviceToHost));
	if (bitfield_out) NGPB_CUDA_CHECK(cudaMemcpy(bitfield_out, t->bitfield, (size_t)NERF_GRID_CELLS * NERF_CASCADES / 8, cudaMemcpyDeviceToHost));
	NGPB_API_END
}

extern "C" int ngpb_nccl_unique_id(void* out128) {
	NGPB_API_BEGIN
	if (!out128) throw std::runtime_error("ngpb_nccl_unique_id: null output");
	NcclApi& nccl = NcclApi::get();
	NcclApi::UniqueId id;
	nccl.check(nccl.GetUniqueId(&id), "ncclGetUniqueId");
	std::memcpy(out128, &id, sizeof(id));
	NGPB_API_END
}
extern "C" int ngpb_testbed_init_data_parallel(ngpb_testbed* t, int rank, int world, const void* unique_id128) {
	NGPB_API_BEGIN t->init_data_parallel(rank, world, unique_id128); NGPB_API_END
}

extern "C" int ngpb_testbed_configure(ngpb_testbed* t, uint32_t aabb_scale, uint32_t seed) {
	NGPB_API_BEGIN
	if (aabb_scale == 0 || (aabb_scale & (aabb_scale - 1)) != 0 || aabb_scale > (1u << (NERF_CASCADES - 1))) throw std::runtime_error("aabb_scale must be a power of two <= 128");
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	t->drop_prefetch();
	t->aabb_scale = aabb_scale;
	t->configure_scene_box();
	t->reset_network(seed);
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_params_half(ngpb_testbed* t, const ngpb_half* params, uint32_t n) {
	NGPB_API_BEGIN
	if (!params || n != t->n_params) throw std::runtime_error("set_params_half: parameter count does not match the network (" + std::to_string(t->n_params) + ")");
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	t->drop_prefetch();
	NGPB_CUDA_CHECK(cudaMemcpyAsync(t->w_half, params, sizeof(__half) * n, cudaMemcpyHostToDevice, t->stream));
	NGPB_CUDA_CHECK(cudaMemcpyAsync(t->w_ema, t->w_half, sizeof(__half) * n, cudaMemcpyDeviceToDevice, t->stream));
	widen_params_kernel<<<div_round_up(n, 256), 256, 0, t->stream>>>(n, t->w_half, t->w_fp32);
	NGPB_LAUNCH_CHECK();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	t->master_weights_sharded = false;
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_density_grid(ngpb_testbed* t, const float* density_grid, uint32_t n_cells) {
	NGPB_API_BEGIN
	const uint32_t expect = NERF_GRID_CELLS * (t->max_cascade + 1);
	if (!density_grid || n_cells != expect) throw std::runtime_error("Incompatible number of grid cascades."); // src/testbed.cu:3093
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	t->drop_prefetch();
	NGPB_CUDA_CHECK(cudaMemcpyAsync(t->density_grid, density_grid, sizeof(float) * n_cells, cudaMemcpyHostToDevice, t->stream));
	if (ngpb_update_bitfield(t->stream, t->max_cascade + 1, t->density_grid, t->mean_density, t->bitfield) != 0) throw std::runtime_error(ngpb_last_error());
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	NGPB_API_END
}
extern "C" int ngpb_testbed_get_training_state(ngpb_testbed* t, ngpb_training_state* o) {
	NGPB_API_BEGIN
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	t->collect_loss_scalar();
	*o = ngpb_training_state{t->training_step, t->rays_per_batch, t->measured_batch_size, t->measured_batch_size_before_compaction, t->loss_scalar,
		t->opt.step, t->opt.learning_rate, t->opt.lr_factor};
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_training_state(ngpb_testbed* t, const ngpb_training_state* in) {
	NGPB_API_BEGIN
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	t->drop_prefetch();
	t->training_step = in->training_step; t->density_grid_ema_step = in->training_step;
	t->rays_per_batch = in->rays_per_batch ? in->rays_per_batch : (1u << 12);
	t->measured_batch_size = in->measured_batch_size; t->measured_batch_size_before_compaction = in->measured_batch_size_before_compaction;
	t->loss_scalar = in->loss; t->loss_pending = false;
	t->opt.step = in->optimizer_step;
	if (in->learning_rate > 0.f) t->opt.learning_rate = in->learning_rate;
	t->opt.lr_factor = in->learning_rate_factor > 0.f ? in->learning_rate_factor : 1.0f;
	NGPB_API_END
}
extern "C" int ngpb_testbed_get_optimizer_state(ngpb_testbed* t, float* fm, float* sm, uint32_t* ps) {
	NGPB_API_BEGIN
	if (t->dp_world > 1 && t->dp_sharded_optimizer) throw std::runtime_error("optimizer state is sharded across the data-parallel ranks; save without it");
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	if (fm) NGPB_CUDA_CHECK(cudaMemcpy(fm, t->m1, sizeof(float) * t->n_params, cudaMemcpyDeviceToHost));
	if (sm) NGPB_CUDA_CHECK(cudaMemcpy(sm, t->m2, sizeof(float) * t->n_params, cudaMemcpyDeviceToHost));
	if (ps) NGPB_CUDA_CHECK(cudaMemcpy(ps, t->param_steps, sizeof(uint32_t) * t->n_params, cudaMemcpyDeviceToHost));
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_optimizer_state(ngpb_testbed* t, const float* fm, const float* sm, const uint32_t* ps) {
	NGPB_API_BEGIN
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	t->drop_prefetch();
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	if (fm) NGPB_CUDA_CHECK(cudaMemcpy(t->m1, fm, sizeof(float) * t->n_params, cudaMemcpyHostToDevice));
	if (sm) NGPB_CUDA_CHECK(cudaMemcpy(t->m2, sm, sizeof(float) * t->n_params, cudaMemcpyHostToDevice));
	if (ps) NGPB_CUDA_CHECK(cudaMemcpy(t->param_steps, ps, sizeof(uint32_t) * t->n_params, cudaMemcpyHostToDevice));
	NGPB_CUDA_CHECK(cudaDeviceSynchronize()); // (see p2p_setup: pageable copies may outlive the call)
	NGPB_API_END
}

extern "C" int ngpb_testbed_create_empty_dataset(ngpb_testbed* t, uint32_t n_images, uint32_t aabb_scale) {
	NGPB_API_BEGIN
	t->create_empty_dataset(n_images, aabb_scale);
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_training_image(ngpb_testbed* t, uint32_t frame_idx, const ngpb_host_image* image) {
	NGPB_API_BEGIN
	if (!image) throw std::runtime_error("set_training_image: invalid argument");
	t->set_training_image(frame_idx, *image);
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_camera_intrinsics(ngpb_testbed* t, uint32_t frame_idx, float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2) {
	NGPB_API_BEGIN
	t->set_camera_intrinsics(frame_idx, fx, fy, cx, cy, k1, k2, p1, p2);
	NGPB_API_END
}
// nerf.training.get_camera_extrinsics / set_camera_extrinsics / reset_camera_extrinsics (:2518-2555, :2590-2595), in the library's (ngp) coordinates
extern "C" int ngpb_testbed_get_camera_extrinsics(ngpb_testbed* t, uint32_t frame_idx, float* xform12, float* pos_offset3, float* rot_offset3) {
	NGPB_API_BEGIN
	if (frame_idx >= t->images.size()) throw std::runtime_error("get_camera_extrinsics: frame index out of range");
	if (xform12) std::memcpy(xform12, t->images[frame_idx].raw_xform, sizeof(float) * 12);
	if (pos_offset3) std::memcpy(pos_offset3, &t->cam_pos_state[(size_t)frame_idx * 10 + 7], sizeof(float) * 3);
	if (rot_offset3) std::memcpy(rot_offset3, &t->cam_rot_state[(size_t)frame_idx * 10 + 7], sizeof(float) * 3);
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_camera_extrinsics(ngpb_testbed* t, uint32_t frame_idx, const float* xform12) {
	NGPB_API_BEGIN
	if (frame_idx >= t->images.size() || !xform12) throw std::runtime_error("set_camera_extrinsics: invalid argument");
	std::memcpy(&t->dataset_xforms[(size_t)frame_idx * 12], xform12, sizeof(float) * 12);
	t->update_transforms();
	NGPB_API_END
}
extern "C" int ngpb_testbed_reset_camera_extrinsics(ngpb_testbed* t) {
	NGPB_API_BEGIN
	t->reset_camera_extrinsics();
	t->update_transforms();
	if (t->exposure_active) t->upload_exposures();
	NGPB_API_END
}
// Per-image exposures (Training::cam_exposure, testbed.h:632): exposures3[n_images][3], in stops (the loss sees colours x 2^exposure).
extern "C" int ngpb_testbed_get_camera_exposures(ngpb_testbed* t, float* exposures3) {
	NGPB_API_BEGIN
	if (!exposures3) throw std::runtime_error("get_camera_exposures: invalid argument");
	for (size_t i = 0; i < t->images.size(); ++i) for (int c = 0; c < 3; ++c) exposures3[i * 3 + c] = t->cam_exposure_state[i * 10 + 7 + c];
	NGPB_API_END
}
// Sampling probabilities of the images after the last CDF update (ErrorMap::pmf_img_cpu); uniform before the first.
extern "C" int ngpb_testbed_get_error_map_pmf(ngpb_testbed* t, float* pmf_img) {
	NGPB_API_BEGIN
	if (!pmf_img) throw std::runtime_error("get_error_map_pmf: invalid argument");
	const size_t n = t->images.size();
	if (!t->error_cdf_valid) { for (size_t i = 0; i < n; ++i) pmf_img[i] = 1.0f / (float)n; }
	else {
		NGPB_CUDA_CHECK(cudaSetDevice(t->device));
		NGPB_CUDA_CHECK(cudaMemcpyAsync(pmf_img, t->error_pmf_img, sizeof(float) * n, cudaMemcpyDeviceToHost, t->stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	}
	NGPB_API_END
}
extern "C" int ngpb_testbed_set_camera_exposures(ngpb_testbed* t, const float* exposures3) {
	NGPB_API_BEGIN
	if (!exposures3 || t->images.empty()) throw std::runtime_error("set_camera_exposures: invalid argument");
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	std::fill(t->cam_exposure_state.begin(), t->cam_exposure_state.end(), 0.f); // (fresh optimizer state, as Training::set_camera_extrinsics does for a frame)
	for (size_t i = 0; i < t->images.size(); ++i) for (int c = 0; c < 3; ++c) t->cam_exposure_state[i * 10 + 7 + c] = exposures3[i * 3 + c];
	t->exposure_active = true;
	t->upload_exposures();
	NGPB_API_END
}

extern "C" void* ngpb_testbed_stream(ngpb_testbed* t) { return t ? (void*)t->stream : nullptr; }
extern "C" int ngpb_testbed_stage_times(ngpb_testbed* t, double* ms, uint64_t* calls, uint64_t* units, int reset) {
	NGPB_API_BEGIN
	NGPB_CUDA_CHECK(cudaSetDevice(t->device));
	NGPB_CUDA_CHECK(cudaStreamSynchronize(t->stream));
	t->stage_collect();
	for (int s = 0; s < NGPB_N_STAGES; ++s) {
		if (ms) ms[s] = t->stage_ms[s];
		if (calls) calls[s] = t->stage_calls[s];
		if (units) units[s] = t->stage_units[s];
		if (reset) { t->stage_ms[s] = 0.0; t->stage_calls[s] = 0; t->stage_units[s] = 0; }
	}
	NGPB_API_END
}

extern "C" int ngpb_testbed_set_option(ngpb_testbed* t, const char* name, double v) {
	NGPB_API_BEGIN
	const std::string k = name ? name : "";
	if (k == "random_bg_color") t->loss_cfg.random_bg_color = v != 0;
	else if (k == "linear_colors") t->loss_cfg.linear_colors = v != 0;
	else if (k == "snap_to_pixel_centers") t->loss_cfg.snap_to_pixel_centers = v != 0;
	else if (k == "loss_type") t->loss_cfg.loss_type = (int)v;
	else if (k == "color_space") t->loss_cfg.color_space = (int)v;
	else if (k == "rgb_activation") t->loss_cfg.rgb_activation = (int)v;
	else if (k == "density_activation") t->loss_cfg.density_activation = (int)v;
	else if (k == "near_distance") t->loss_cfg.near_distance = (float)v;
	else if (k == "background_color_r") t->loss_cfg.background_color[0] = (float)v;
	else if (k == "background_color_g") t->loss_cfg.background_color[1] = (float)v;
	else if (k == "background_color_b") t->loss_cfg.background_color[2] = (float)v;
	else if (k == "cone_angle_constant") t->cone_angle_constant = (float)v;
	else if (k == "density_grid_decay") t->density_grid_decay = (float)v;
	else if (k == "shall_train") t->shall_train = v != 0;
	else if (k == "render_min_transmittance") t->render_min_transmittance = (float)v;
	else if (k == "learning_rate") { t->opt.learning_rate = (float)v; t->opt_hyper.learning_rate = (float)v; }
	// optimizer section of the network config (adam.h:121-160 update_hyperparams, exponential_decay.h:100-120, ema.h:150-170)
	else if (k == "adam_beta1") { t->opt.beta1 = t->opt_hyper.beta1 = (float)v; }
	else if (k == "adam_beta2") { t->opt.beta2 = t->opt_hyper.beta2 = (float)v; }
	else if (k == "adam_epsilon") { t->opt.epsilon = t->opt_hyper.epsilon = (float)v; }
	else if (k == "adam_l2_reg") { t->opt.l2_reg = t->opt_hyper.l2_reg = (float)v; }
	else if (k == "ema_decay") { t->opt.ema_decay = t->opt_hyper.ema_decay = (float)v; }
	else if (k == "decay_start") { t->opt.decay_start = t->opt_hyper.decay_start = (uint32_t)v; }
	else if (k == "decay_interval") { t->opt.decay_interval = t->opt_hyper.decay_interval = (uint32_t)v; }
	else if (k == "decay_base") { t->opt.decay_base = t->opt_hyper.decay_base = (float)v; }
	else if (k == "log2_hashmap_size") {
		if (v < 14 || v > 24) throw std::runtime_error("encoding.log2_hashmap_size must be in [14, 24]");
		t->log2_hashmap_size = (uint32_t)v; // takes effect at the next reset_network
	}
	else if (k == "optimize_extrinsics") t->optimize_extrinsics = v != 0;
	else if (k == "optimize_exposure") t->optimize_exposure = v != 0;
	else if (k == "sample_focal_plane_proportional_to_error") t->sample_focal_plane_proportional_to_error = v != 0;
	else if (k == "sample_image_proportional_to_error") t->sample_image_proportional_to_error = v != 0;
	else if (k == "n_images_for_training") {
		if (v < 0 || v > (double)t->images.size()) throw std::runtime_error("n_images_for_training must be in [0, n_images]");
		t->drop_prefetch(); // (a prefetched K1 drew its images from the previous count)
		t->n_images_for_training = (uint32_t)v;
	}
	else if (k == "exposure_l2_reg") t->exposure_l2_reg = (float)v;
	else if (k == "extrinsic_learning_rate") t->extrinsic_learning_rate = (float)v;
	else if (k == "extrinsic_l2_reg") t->extrinsic_l2_reg = (float)v;
	else if (k == "n_steps_between_cam_updates") { if (v < 1) throw std::runtime_error("n_steps_between_cam_updates must be >= 1"); t->n_steps_between_cam_updates = (uint32_t)v; }
	else if (k == "profile_stages") t->profile_stages = v != 0;
	else if (k == "render_snap_to_pixel_centers") t->render_snap_to_pixel_centers = v != 0;
	else if (k == "render_near_distance") t->render_near_distance = (float)v;
	else if (k == "exposure") t->exposure = (float)v;
	else if (k == "tonemap_curve") { if (v < 0 || v > 3) throw std::runtime_error("tonemap_curve must be one of Identity, ACES, Hable, Reinhard"); t->tonemap_curve = (int)v; }
	else if (k == "background_color_a") t->background_alpha = (float)v;
	else if (k == "render_with_training_params") t->render_with_training_params = v != 0;
	else if (k == "dp_sharded_optimizer") t->dp_sharded_optimizer = v != 0;
	else if (k == "overlap_sampling") { t->drop_prefetch(); t->overlap_sampling = v != 0; }
	else if (k == "reuse_encoding") t->reuse_encoding = v != 0;
	else if (k == "dp_half_gradients") t->dp_half_gradients = (int)v; // 0 fp32, 1 bf16, 2 fp16
	else if (k == "dp_exchange") t->dp_exchange = (int)v;             // 0 NCCL collectives, 1 peer-memory kernels
	else throw std::runtime_error("unknown option: " + k);
	NGPB_API_END
}
extern "C" double ngpb_testbed_get_option(ngpb_testbed* t, const char* name) {
	const std::string k = name ? name : "";
	if (k == "random_bg_color") return t->loss_cfg.random_bg_color;
	if (k == "linear_colors") return t->loss_cfg.linear_colors;
	if (k == "snap_to_pixel_centers") return t->loss_cfg.snap_to_pixel_centers;
	if (k == "loss_type") return t->loss_cfg.loss_type;
	if (k == "color_space") return t->loss_cfg.color_space;
	if (k == "rgb_activation") return t->loss_cfg.rgb_activation;
	if (k == "density_activation") return t->loss_cfg.density_activation;
	if (k == "near_distance") return t->loss_cfg.near_distance;
	if (k == "cone_angle_constant") return t->cone_angle_constant;
	if (k == "density_grid_decay") return t->density_grid_decay;
	if (k == "shall_train") return t->shall_train;
	if (k == "render_min_transmittance") return t->render_min_transmittance;
	if (k == "learning_rate") return t->opt.learning_rate;
	if (k == "adam_beta1") return t->opt.beta1;
	if (k == "adam_beta2") return t->opt.beta2;
	if (k == "adam_epsilon") return t->opt.epsilon;
	if (k == "adam_l2_reg") return t->opt.l2_reg;
	if (k == "ema_decay") return t->opt.ema_decay;
	if (k == "decay_start") return t->opt.decay_start;
	if (k == "decay_interval") return t->opt.decay_interval;
	if (k == "decay_base") return t->opt.decay_base;
	if (k == "log2_hashmap_size") return t->log2_hashmap_size;
	if (k == "optimize_extrinsics") return t->optimize_extrinsics;
	if (k == "optimize_exposure") return t->optimize_exposure;
	if (k == "sample_focal_plane_proportional_to_error") return t->sample_focal_plane_proportional_to_error;
	if (k == "sample_image_proportional_to_error") return t->sample_image_proportional_to_error;
	if (k == "error_map_res") return t->error_cdf_valid ? t->error_cdf_res[0] : 0;
	if (k == "error_cdf_valid") return t->error_cdf_valid;
	if (k == "n_steps_between_error_map_updates") return t->n_steps_between_error_map_updates;
	if (k == "n_steps_since_error_map_update") return t->n_steps_since_error_map_update;
	if (k == "n_images") return (double)t->images.size();
	if (k == "n_images_for_training") return (double)t->n_images_for_training;
	if (k == "exposure_l2_reg") return t->exposure_l2_reg;
	if (k == "extrinsic_learning_rate") return t->extrinsic_learning_rate;
	if (k == "extrinsic_l2_reg") return t->extrinsic_l2_reg;
	if (k == "n_steps_between_cam_updates") return t->n_steps_between_cam_updates;
	if (k == "n_steps_since_cam_update") return t->n_steps_since_cam_update;
	if (k == "render_snap_to_pixel_centers") return t->render_snap_to_pixel_centers;
	if (k == "render_near_distance") return t->render_near_distance;
	if (k == "exposure") return t->exposure;
	if (k == "tonemap_curve") return t->tonemap_curve;
	if (k == "background_color_r") return t->loss_cfg.background_color[0];
	if (k == "background_color_g") return t->loss_cfg.background_color[1];
	if (k == "background_color_b") return t->loss_cfg.background_color[2];
	if (k == "background_color_a") return t->background_alpha;
	if (k == "aabb_scale") return t->aabb_scale;
	if (k == "max_cascade") return t->max_cascade;
	if (k == "h2d_bytes") return (double)t->h2d_bytes;
	if (k == "d2h_bytes") return (double)t->d2h_bytes;
	return std::nan("");
}

// Testbed::render_to_cpu (src/python_api.cu:132-190) -> render_frame (src/testbed.cu:2695) -> render_nerf (src/testbed_nerf.cu:2354) for a static camera.
void ngpb_testbed::render(const float* camera12, int w, int h, float fx, float fy, int spp, bool linear, float* out_rgba, uint64_t* n_samples_out) {
	NGPB_CUDA_CHECK(cudaSetDevice(device));
	join_ema();
	if (!camera12 || !out_rgba || w <= 0 || h <= 0 || spp <= 0) throw std::runtime_error("render: invalid argument");
	if (n_params == 0) throw std::runtime_error("render: no network (load training data or a snapshot first)");
	const uint32_t n_pixels = (uint32_t)w * (uint32_t)h;
	const size_t need = (size_t)ngpb_render_workspace_bytes(n_pixels);
	if (need > render_ws_bytes) {
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		dfree(render_ws);
		render_ws = dalloc(need);
		render_ws_bytes = need;
		NGPB_CUDA_CHECK(cudaMemsetAsync(render_ws, 0, need, stream)); // (the network passes run over whole 128-sample tiles: slots past a wave's last sample are read and ignored)
	}
	ngpb_render_config c{};
	c.width = w; c.height = h; c.fx = fx; c.fy = fy;
	c.screen_center[0] = c.screen_center[1] = 0.5f;
	std::memcpy(c.camera, camera12, sizeof(float) * 12);
	c.spp = spp; c.snap_to_pixel_centers = render_snap_to_pixel_centers;
	std::memcpy(c.aabb, aabb, sizeof(aabb));
	std::memcpy(c.render_aabb, aabb, sizeof(aabb)); // m_render_aabb = m_aabb (testbed_nerf.cu:2717)
	c.cone_angle_constant = cone_angle_constant; c.min_transmittance = render_min_transmittance; c.near_distance = render_near_distance;
	c.rgb_activation = loss_cfg.rgb_activation; c.density_activation = loss_cfg.density_activation; c.train_in_linear_colors = loss_cfg.linear_colors;
	c.color_space = loss_cfg.color_space; c.output_srgb = linear ? 0 : 1;
	c.exposure = exposure; c.tonemap_curve = tonemap_curve;
	for (int k = 0; k < 3; ++k) c.background_color[k] = loss_cfg.background_color[k];
	c.background_color[3] = background_alpha;
	// the prefetched sampling kernels of the next training step share no buffer with the render path; the main stream orders the rest
	cudaEvent_t e0, e1;
	NGPB_CUDA_CHECK(cudaEventCreate(&e0)); NGPB_CUDA_CHECK(cudaEventCreate(&e1));
	NGPB_CUDA_CHECK(cudaEventRecord(e0, stream));
	uint32_t launches = 0;
	// Network::inference_mixed_precision defaults to the inference (EMA) parameters (testbed_nerf.cu:2223)
	const int st = ngpb_render_nerf(stream, &c, &grid, (const ngpb_half*)(render_with_training_params ? w_half : w_ema), bitfield, render_ws, out_rgba, n_samples_out, &launches);
	NGPB_CUDA_CHECK(cudaEventRecord(e1, stream));
	NGPB_CUDA_CHECK(cudaEventSynchronize(e1));
	float ms = 0.f;
	cudaEventElapsedTime(&ms, e0, e1);
	last_render_ms = ms;
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if (st != 0) throw std::runtime_error(ngpb_last_error());
	n_launches += launches;
	d2h_bytes += (uint64_t)n_pixels * 16;
}

extern "C" double ngpb_testbed_last_render_ms(const ngpb_testbed* t) { return t ? t->last_render_ms : 0.0; }

extern "C" int ngpb_testbed_render(ngpb_testbed* t, const float* camera12, int w, int h, float fx, float fy, int spp, int linear, float* out_rgba, uint64_t* n_samples_out) {
	NGPB_API_BEGIN
	t->render(camera12, w, h, fx, fy, spp, linear != 0, out_rgba, n_samples_out);
	NGPB_API_END
}
