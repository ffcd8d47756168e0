// NeRF MLPs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM): launchers, the fixed-order reduction of the weight-gradient partial
// sums, the tcgen05 building-block self-test and the C ABI. The kernels are in nerf_mlp_pipe.cu, the pieces they share in nerf_mlp_shared.cuh / umma.cuh.
// Replaces tcnn FullyFusedMLP<__half,64> forward/backward (reference: dependencies/tiny-cuda-nn/src/fully_fused_mlp.cu:151-314 backward, :500-557 forward,
// :759-850 weight gradients through CUTLASS split-K GEMMs) together with the NerfNetwork glue kernels around it (include/neural-graphics-primitives/
// nerf_network.h:32-74 extract_density / extract_rgb / add_density_gradient, :103-266) and the SH direction encoding (spherical_harmonics.h:46-150).
//
// Shape of the computation. A tile is 128 samples (= UMMA M, one TMEM lane per sample, one thread per sample row in the epilogues). Every layer is one batch
// of tcgen05.mma K=16 steps: D[128 x N] (+)= A[128 x K] * W[N x K]^T with A = the previous layer's fp16 activations in shared memory and W = the fp16
// weights in shared memory. The epilogue reads the fp32 accumulator row with tcgen05.ld, applies ReLU, rounds to fp16 and writes the row back to shared
// memory as the next layer's A operand. Activations never leave the SM.
//
// Training (forward + backward in ONE kernel): the activation tiles stay in shared memory, the data-gradient GEMMs reuse the same weight tiles as MN-major B
// operands (no transposed copy), and the five weight-gradient GEMMs dW = dOut^T * Act (K = the 128 samples of the tile) read the activation / gradient tiles
// as MN-major operands and ACCUMULATE IN TMEM ACROSS ALL TILES of a slot -- the reference writes activations to HBM, re-reads them, and runs five split-K
// CUTLASS GEMMs on side streams instead. Each slot writes one fp32 partial per weight at the end; reduce_partials_kernel sums them in a fixed order.
//
// Numerics: fp16 operands, fp32 accumulation (the reference's wmma path accumulates in fp16, fully_fused_mlp.cu:66-68), fp16 rounding at the same layer
// boundaries as the reference.
#include "common.cuh"
#include "umma.cuh"
#include "nerf_mlp_shared.cuh"
#include "../../include/ngpb.h"

namespace ngpb {
using namespace umma;

// Fixed-order sum of the per-CTA partials (deterministic), overwrites mlp_grad.
// Sums the per-CTA weight-gradient partials in a fixed order: 32 parameters x 8 groups per block, group g adds partials g, g+8, g+16, ... in
// sequence (four independent chains in flight), then the 8 group sums are added in order 0..7. Same result on every run.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, const uint32_t n_parts, float* __restrict__ mlp_grad, const uint32_t n_weights)
{
	__shared__ float part[8][32];
	const uint32_t col = threadIdx.x & 31, grp = threadIdx.x >> 5;
	const uint32_t i = blockIdx.x * 32 + col;
	float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
	if (i < n_weights) {
		uint32_t p = grp;
		for (; p + 24 < n_parts; p += 32) {
			s0 += partials[(size_t)p * n_weights + i];
			s1 += partials[(size_t)(p + 8) * n_weights + i];
			s2 += partials[(size_t)(p + 16) * n_weights + i];
			s3 += partials[(size_t)(p + 24) * n_weights + i];
		}
		for (; p < n_parts; p += 8) s0 += partials[(size_t)p * n_weights + i];
	}
	part[grp][col] = (s0 + s1) + (s2 + s3);
	__syncthreads();
	if (grp == 0 && i < n_weights) {
		float s = part[0][col];
		#pragma unroll
		for (int g = 1; g < 8; ++g) s += part[g][col];
		mlp_grad[i] = s;
	}
}

constexpr uint32_t TRAIN_GRID = kNumSMs * 2; // partial-sum sets of the training kernel: one per tile slot, two slots per CTA, one CTA per SM

// The kernels themselves: nerf_mlp_pipe.cu (warp-specialised persistent pipelines). Round 1's one-tile-at-a-time kernels (47.6 / 77 us against 42.5 / 77 us) are in the history.
void pipe_nerf_forward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma);
void pipe_density_forward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, uint32_t n, __half* density);
void pipe_plain_forward(cudaStream_t stream, const __half* weights, const __half* input, uint32_t n, __half* output);
uint32_t pipe_nerf_forward_backward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, const __half* dL_dout, uint32_t n,
                                    __half* dL_dencoded, float* partials, __half* dL_dsh);
uint32_t pipe_plain_forward_backward(cudaStream_t stream, const __half* weights, const __half* input, const __half* dL_dout16, uint32_t n, __half* dL_dinput, float* partials);

// Layout of the [n][32] fp16 hash-grid features handed from the encoding kernels to the MLP kernels inside the library: per 128 samples one 8 KB block in
// the UMMA core-matrix layout (umma.cuh tile_offset), which the pipelined kernels fetch with one TMA bulk copy. The public C ABI stays row-major.
bool features_tiled() { return true; }

// Internal launchers shared with the testbed host (device-side sample count supported). `tiled`: layout of `encoded` (see features_tiled()).
void nerf_mlp_forward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma) {
	pipe_nerf_forward(stream, mlp, encoded, tiled, coords, n, n_dev, rgbsigma);
}
void plain_mlp_launch(cudaStream_t stream, const __half* weights, const __half* input, uint32_t n, __half* output) {
	pipe_plain_forward(stream, weights, input, n, output);
}
void plain_mlp_forward_backward_launch(cudaStream_t stream, const __half* weights, const __half* input, const __half* dL_dout16, uint32_t n, __half* dL_dinput, float* grad,
                                       float* partials) {
	const uint32_t n_sets = pipe_plain_forward_backward(stream, weights, input, dL_dout16, n, dL_dinput, partials);
	reduce_partials_kernel<<<div_round_up(PLAIN_PARAMS, 32), 256, 0, stream>>>(partials, n_sets, grad, PLAIN_PARAMS);
	NGPB_LAUNCH_CHECK();
}
void nerf_density_mlp_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, uint32_t n, __half* density) {
	pipe_density_forward(stream, mlp, encoded, tiled, n, density);
}
void nerf_mlp_forward_backward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, const __half* dL_dout, uint32_t n,
                                      __half* dL_dencoded, float* mlp_grad, float* partials, __half* dL_dsh) {
	const uint32_t n_sets = pipe_nerf_forward_backward(stream, mlp, encoded, tiled, coords, dL_dout, n, dL_dencoded, partials, dL_dsh);
	NGPB_STEP_KERNEL(reduce_partials_kernel);
	reduce_partials_kernel<<<div_round_up(MLP_PARAMS, 32), 256, 0, stream>>>(partials, n_sets, mlp_grad, MLP_PARAMS);
	NGPB_LAUNCH_CHECK();
}

// ---- tcgen05 building-block self-test (tests/test_umma_selftest.py) -------------------------------
// variant 0: D[128x64] = A[128x32] * B[64x32]^T           (K-major A, K-major B: forward layer)
// variant 1: D[128x32] = A[128x64] * B[64x32]             (K-major A, MN-major B: data gradient)
// variant 2: D[64x32]  = A[128x64]^T * B[128x32]          (MN-major A and B, M = 64: weight gradient)
__global__ void __launch_bounds__(128) umma_selftest_kernel(const int variant, const __half* __restrict__ a, const __half* __restrict__ b, float* __restrict__ d)
{
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr uint32_t OFF_A = 0, OFF_B = 16384, OFF_CTRL = 32768;
	uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_CTRL);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_CTRL + 16);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (warp == 0) tmem_alloc<64>(tmem_slot);
	if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
	const uint32_t a_rows = 128, a_cols = variant == 0 ? 32 : 64;
	const uint32_t b_rows = variant == 2 ? 128 : 64, b_cols = 32;
	load_matrix_to_tile(smem, OFF_A, a, a_rows, a_cols);
	load_matrix_to_tile(smem, OFF_B, b, b_rows, b_cols);
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t sbase = smem_u32(smem);
	if (tid == 0) {
		if (variant == 0) issue_forward(tmem_base, sbase + OFF_A, 32, sbase + OFF_B, 32, 64);
		else if (variant == 1) issue_dgrad(tmem_base, sbase + OFF_A, 64, sbase + OFF_B, 32, 64);
		else issue_wgrad(tmem_base, sbase + OFF_A, sbase + OFF_B, 32, 32, true);
		mma_commit(mbar);
	}
	__syncwarp();
	mbar_wait(mbar, 0);
	tc_fence_after_sync();
	const uint32_t t_row = tmem_base + ((warp * 32u) << 16);
	uint32_t r[32];
	if (variant == 0) {
		for (uint32_t h = 0; h < 2; ++h) {
			tmem_ld_x32(t_row + h * 32, r); tmem_ld_wait();
			for (uint32_t i = 0; i < 32; ++i) d[(size_t)tid * 64 + h * 32 + i] = __uint_as_float(r[i]);
		}
	} else if (variant == 1) {
		tmem_ld_x32(t_row, r); tmem_ld_wait();
		for (uint32_t i = 0; i < 32; ++i) d[(size_t)tid * 32 + i] = __uint_as_float(r[i]);
	} else {
		tmem_ld_x32(t_row, r); tmem_ld_wait();
		if (lane < 16) for (uint32_t i = 0; i < 32; ++i) d[(size_t)(warp * 16 + lane) * 32 + i] = __uint_as_float(r[i]);
	}
	tc_fence_before_sync();
	__syncthreads();
	if (warp == 0) tmem_dealloc<64>(tmem_base);
}

} // namespace ngpb

using namespace ngpb;

extern "C" int ngpb_nerf_mlp_forward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, uint32_t n, ngpb_half* rgbsigma) {
	try {
		if (!mlp || !encoded || !coords || !rgbsigma || n % TILE != 0) { set_last_error("ngpb_nerf_mlp_forward: invalid argument (n must be a multiple of 128)"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (n == 0) return 0;
		nerf_mlp_forward_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, coords, n, nullptr, (__half*)rgbsigma);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_mlp_forward(void* stream, const ngpb_half* weights, const ngpb_half* input, uint32_t n, ngpb_half* output) {
	try {
		if (!weights || !input || !output || n == 0 || n % TILE != 0) { set_last_error("ngpb_mlp_forward: null pointer or n not a non-zero multiple of 128"); return NGPB_ERR_INVALID_ARGUMENT; }
		plain_mlp_launch((cudaStream_t)stream, (const __half*)weights, (const __half*)input, n, (__half*)output);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_mlp_forward_backward(void* stream, const ngpb_half* weights, const ngpb_half* input, const ngpb_half* dL_dout, uint32_t n, ngpb_half* dL_dinput,
                                         float* grad, void* workspace) {
	try {
		if (!weights || !input || !dL_dout || !dL_dinput || !grad || !workspace || n == 0 || n % TILE != 0) {
			set_last_error("ngpb_mlp_forward_backward: null pointer or n not a non-zero multiple of 128");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		plain_mlp_forward_backward_launch((cudaStream_t)stream, (const __half*)weights, (const __half*)input, (const __half*)dL_dout, n, (__half*)dL_dinput, grad, (float*)workspace);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_nerf_density_mlp_forward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, uint32_t n, ngpb_half* density) {
	try {
		if (!mlp || !encoded || !density || n % TILE != 0) { set_last_error("ngpb_nerf_density_mlp_forward: invalid argument (n must be a multiple of 128)"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (n == 0) return 0;
		nerf_density_mlp_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, n, (__half*)density);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" uint64_t ngpb_nerf_mlp_workspace_bytes(void) { return (uint64_t)TRAIN_GRID * MLP_PARAMS * sizeof(float); }

extern "C" int ngpb_nerf_mlp_forward_backward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, const ngpb_half* dL_dout,
                                              uint32_t n, ngpb_half* dL_dencoded, float* mlp_grad, void* workspace) {
	try {
		if (!mlp || !encoded || !coords || !dL_dout || !dL_dencoded || !mlp_grad || !workspace || n == 0 || n % TILE != 0) {
			set_last_error("ngpb_nerf_mlp_forward_backward: invalid argument (n must be a non-zero multiple of 128)");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		nerf_mlp_forward_backward_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, coords, (const __half*)dL_dout, n,
			(__half*)dL_dencoded, mlp_grad, (float*)workspace, nullptr);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

// The same pass, also returning dL/d(SH inputs of the rgb network) [n][16] half: the direction half of the input gradient (camera_optimizer.cu)
extern "C" int ngpb_nerf_mlp_forward_backward_sh(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, const ngpb_half* dL_dout,
                                                 uint32_t n, ngpb_half* dL_dencoded, float* mlp_grad, void* workspace, ngpb_half* dL_dsh) {
	try {
		if (!mlp || !encoded || !coords || !dL_dout || !dL_dencoded || !mlp_grad || !workspace || !dL_dsh || n == 0 || n % TILE != 0) {
			set_last_error("ngpb_nerf_mlp_forward_backward_sh: invalid argument (n must be a non-zero multiple of 128)");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		nerf_mlp_forward_backward_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, coords, (const __half*)dL_dout, n,
			(__half*)dL_dencoded, mlp_grad, (float*)workspace, (__half*)dL_dsh);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_selftest_umma(void* stream, int variant, const ngpb_half* a, const ngpb_half* b, float* d) {
	try {
		if (variant < 0 || variant > 2 || !a || !b || !d) { set_last_error("ngpb_selftest_umma: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		umma_selftest_kernel<<<1, 128, 32768 + 64, (cudaStream_t)stream>>>(variant, (const __half*)a, (const __half*)b, d);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
