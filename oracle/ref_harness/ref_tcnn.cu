// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" launcher around the reference's *own* tiny-cuda-nn kernels, compiled from
// the sources where they lie under /root/reference (nothing is copied into this repo).
// Built by oracle/Makefile into oracle/_ref/libref_tcnn.so and run on the GPU box by
// oracle/gen_golden.py to produce the golden vectors committed under tests/golden/.
//
// Kernels launched (reference file:line):
//   kernel_grid<__half,3,2>, <__half,2,2>   dependencies/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:220
//   kernel_grid_backward<__half,__half,3,2,2>                                          grid.h:395
//   kernel_sh<__half>                  .../encodings/spherical_harmonics.h:46
//   kernel_grid (dy_dx) + kernel_grid_backward_input<__half,3>                          grid.h:351,:551
//   kernel_sh_backward<__half>         .../encodings/spherical_harmonics.h:154
//   adam_step<__half>                  .../optimizers/adam.h:48
//   ema_step_half_precision<__half>    .../optimizers/ema.h:63
#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/gpu_memory.h>
#include <tiny-cuda-nn/encodings/grid.h>
#include <tiny-cuda-nn/encodings/spherical_harmonics.h>
#include <tiny-cuda-nn/optimizers/adam.h>
#include <tiny-cuda-nn/optimizers/ema.h>

using namespace tcnn;

static GridOffsetTable make_table(const uint32_t* offsets, uint32_t n_levels) {
	GridOffsetTable t;
	for (uint32_t i = 0; i <= n_levels; ++i) t.data[i] = offsets[i];
	t.size = n_levels + 1;
	return t;
}

extern "C" {

// positions: AoS float with `pos_stride` floats per sample (reference passes the NerfCoordinate
// matrix view: stride_i=1, stride_j=floats_per_coord). out: SoA half [n_features][n].
int ref_grid_forward(uint32_t n, uint32_t n_levels, const uint32_t* offsets_host, uint32_t base_resolution,
                     float log2_per_level_scale, const void* grid_half, const float* positions, uint32_t pos_stride,
                     void* out_soa_half) {
	GridOffsetTable t = make_table(offsets_host, n_levels);
	const dim3 blocks = { div_round_up(n, 512u), n_levels, 1 };
	kernel_grid<__half, 3, 2><<<blocks, 512>>>(
		n, n_levels * 2, t, base_resolution, log2_per_level_scale, 0.f, 1000.f, nullptr,
		InterpolationType::Linear, GridType::Hash, HashType::CoherentPrime,
		(const __half*)grid_half, MatrixView<const float>(positions, 1, pos_stride), (__half*)out_soa_half, nullptr);
	return (int)cudaDeviceSynchronize();
}

// The same kernel instantiated for a two-dimensional input (the neural-image model, N_POS_DIMS = 2).
int ref_grid_forward_2d(uint32_t n, uint32_t n_levels, const uint32_t* offsets_host, uint32_t base_resolution,
                        float log2_per_level_scale, const void* grid_half, const float* positions, uint32_t pos_stride,
                        void* out_soa_half) {
	GridOffsetTable t = make_table(offsets_host, n_levels);
	const dim3 blocks = { div_round_up(n, 512u), n_levels, 1 };
	kernel_grid<__half, 2, 2><<<blocks, 512>>>(
		n, n_levels * 2, t, base_resolution, log2_per_level_scale, 0.f, 1000.f, nullptr,
		InterpolationType::Linear, GridType::Hash, HashType::CoherentPrime,
		(const __half*)grid_half, MatrixView<const float>(positions, 1, pos_stride), (__half*)out_soa_half, nullptr);
	return (int)cudaDeviceSynchronize();
}

// Timing: `iters` back-to-back launches of the reference's forward / backward grid kernels on the NULL stream, mean milliseconds per launch
// (CUDA events), after one warm-up launch. The backward figure includes the gradient memset the reference issues before it (grid.h:1154).
int ref_grid_time(uint32_t n, uint32_t n_levels, const uint32_t* offsets_host, uint32_t base_resolution, float log2_per_level_scale,
                  const void* grid_half, void* grid_gradient_half, const float* positions, uint32_t pos_stride, void* out_soa_half,
                  const void* dL_dy_soa_half, int iters, float* fwd_ms, float* bwd_ms) {
	GridOffsetTable t = make_table(offsets_host, n_levels);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	const dim3 fb = { div_round_up(n, 512u), n_levels, 1 }, bb = { div_round_up(n * 2 / 2, 256u), n_levels, 1 };
	for (int pass = 0; pass < 2; ++pass) {
		const int reps = pass == 0 ? 1 : iters;
		cudaEventRecord(e0, nullptr);
		for (int it = 0; it < reps; ++it)
			kernel_grid<__half, 3, 2><<<fb, 512>>>(n, n_levels * 2, t, base_resolution, log2_per_level_scale, 0.f, 1000.f, nullptr, InterpolationType::Linear, GridType::Hash,
				HashType::CoherentPrime, (const __half*)grid_half, MatrixView<const float>(positions, 1, pos_stride), (__half*)out_soa_half, nullptr);
		cudaEventRecord(e1, nullptr);
		cudaEventSynchronize(e1);
		if (pass) { cudaEventElapsedTime(fwd_ms, e0, e1); *fwd_ms /= (float)iters; }
	}
	for (int pass = 0; pass < 2; ++pass) {
		const int reps = pass == 0 ? 1 : iters;
		cudaEventRecord(e0, nullptr);
		for (int it = 0; it < reps; ++it) {
			cudaMemsetAsync(grid_gradient_half, 0, (size_t)offsets_host[n_levels] * 2 * sizeof(__half), nullptr);
			kernel_grid_backward<__half, __half, 3, 2, 2><<<bb, 256>>>(n, n_levels * 2, t, base_resolution, log2_per_level_scale, 1000.f, nullptr, false, InterpolationType::Linear,
				GridType::Hash, HashType::CoherentPrime, (__half*)grid_gradient_half, MatrixView<const float>(positions, 1, pos_stride), (const __half*)dL_dy_soa_half);
		}
		cudaEventRecord(e1, nullptr);
		cudaEventSynchronize(e1);
		if (pass) { cudaEventElapsedTime(bwd_ms, e0, e1); *bwd_ms /= (float)iters; }
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	return (int)cudaDeviceSynchronize();
}

// dL_dy: SoA half [n_features][n]; grad: half table, zeroed here first (EGradientMode::Overwrite, grid.h:1154).
int ref_grid_backward(uint32_t n, uint32_t n_levels, const uint32_t* offsets_host, uint32_t base_resolution,
                      float log2_per_level_scale, void* grid_gradient_half, const float* positions, uint32_t pos_stride,
                      const void* dL_dy_soa_half) {
	GridOffsetTable t = make_table(offsets_host, n_levels);
	cudaMemset(grid_gradient_half, 0, (size_t)offsets_host[n_levels] * 2 * sizeof(__half));
	const dim3 blocks = { div_round_up(n * 2 / 2, 256u), n_levels, 1 };
	kernel_grid_backward<__half, __half, 3, 2, 2><<<blocks, 256>>>(
		n, n_levels * 2, t, base_resolution, log2_per_level_scale, 1000.f, nullptr, false,
		InterpolationType::Linear, GridType::Hash, HashType::CoherentPrime,
		(__half*)grid_gradient_half, MatrixView<const float>(positions, 1, pos_stride), (const __half*)dL_dy_soa_half);
	return (int)cudaDeviceSynchronize();
}

// dirs: AoS float (stride floats per sample); out: AoS half with out_stride halfs per sample.
int ref_sh4(uint32_t n, const float* dirs, uint32_t dir_stride, void* out_half, uint32_t out_stride) {
	linear_kernel(kernel_sh<__half>, 0, 0, n, 4u, 0u, MatrixView<const float>(dirs, 1, dir_stride), MatrixView<__half>((__half*)out_half, 1, out_stride));
	return (int)cudaDeviceSynchronize();
}

int ref_adam_step(uint32_t n, uint32_t n_matrix, float loss_scale, float lr, float beta1, float beta2, float eps, float l2_reg,
                  float* w_fp, void* w_half, const void* g_half, float* m1, float* m2, uint32_t* steps) {
	linear_kernel(adam_step<__half>, 0, 0, n, n_matrix, 0.f, 0.f, 0.f, loss_scale, lr, 1.f, true, true, beta1, beta2, eps,
		0.f, std::numeric_limits<float>::max(), l2_reg, w_fp, (__half*)w_half, (const __half*)g_half, m1, m2, steps);
	return (int)cudaDeviceSynchronize();
}

int ref_ema_step(uint32_t n, float decay, float debias_old, float debias_new, const void* w_half, void* w_ema_half) {
	linear_kernel(ema_step_half_precision<__half>, 0, 0, n, decay, debias_old, debias_new, (const __half*)w_half, (__half*)w_ema_half);
	return (int)cudaDeviceSynchronize();
}


// Input gradient of the grid encoding: kernel_grid with dy_dx (grid.h:351-392) followed by kernel_grid_backward_input (grid.h:551-575), as
// GridEncoding::forward_impl(prepare_input_gradients) + backward_impl do. dL_dy: SoA half [n_features][n]. dL_dx: AoS float, dx_stride floats per sample.
int ref_grid_input_gradient(uint32_t n, uint32_t n_levels, const uint32_t* offsets_host, uint32_t base_resolution, float log2_per_level_scale,
                            const void* grid_half, const float* positions, uint32_t pos_stride, const void* dL_dy_soa_half, float* dL_dx, uint32_t dx_stride) {
	GridOffsetTable t = make_table(offsets_host, n_levels);
	GPUMemory<float> dy_dx((size_t)n * n_levels * 2 * 3);
	GPUMemory<__half> encoded((size_t)n * n_levels * 2);
	const dim3 blocks = { div_round_up(n, 512u), n_levels, 1 };
	kernel_grid<__half, 3, 2><<<blocks, 512>>>(
		n, n_levels * 2, t, base_resolution, log2_per_level_scale, 0.f, 1000.f, nullptr,
		InterpolationType::Linear, GridType::Hash, HashType::CoherentPrime,
		(const __half*)grid_half, MatrixView<const float>(positions, 1, pos_stride), encoded.data(), dy_dx.data());
	linear_kernel(kernel_grid_backward_input<__half, 3>, 0, 0, n, n_levels * 2, (const __half*)dL_dy_soa_half, dy_dx.data(), MatrixView<float>(dL_dx, 1, dx_stride));
	return (int)cudaDeviceSynchronize();
}

// kernel_sh_backward (spherical_harmonics.h:154-390), degree 4. dL_dy: AoS half, dy_stride halfs per sample; dirs / dL_dx: AoS float.
int ref_sh4_backward(uint32_t n, const void* dL_dy_half, uint32_t dy_stride, const float* dirs, uint32_t dir_stride, float* dL_dx, uint32_t dx_stride) {
	linear_kernel(kernel_sh_backward<__half>, 0, 0, n, 4u, 0u, MatrixView<const __half>((const __half*)dL_dy_half, 1, dy_stride),
		MatrixView<const float>(dirs, 1, dir_stride), MatrixView<float>(dL_dx, 1, dx_stride));
	return (int)cudaDeviceSynchronize();
}

}
