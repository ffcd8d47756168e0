// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" launcher around the reference's own FullyFusedMLP<__half,64>
// (dependencies/tiny-cuda-nn/src/fully_fused_mlp.cu: kernel_mlp_fused :500, kernel_mlp_fused_backward :151,
// CUTLASS split-K weight gradients :805-847), linked against the object file compiled from that source
// where it lies under /root/reference (nothing is copied). Built by oracle/Makefile into
// oracle/_ref/libref_mlp.so; run on the GPU box by oracle/gen_golden.py (golden vectors for the MLP parity
// tests) and by oracle/time_reference_kernels.py (the reference's kernels timed on the same B200).
#include <tiny-cuda-nn/common.h>
#include <tiny-cuda-nn/gpu_matrix.h>
#include <tiny-cuda-nn/networks/fully_fused_mlp.h>

using namespace tcnn;
using T = __half;

// FullyFusedMLP::hyperparams() names its activations through tcnn::to_string(Activation), defined in tcnn's src/network.cu together with the whole
// network factory (which would pull in every other network and CUTLASS instantiation). Nothing here calls hyperparams(): satisfy the linker.
namespace tcnn { std::string to_string(Activation) { return "unused"; } }

extern "C" {

// params: half, reference order (first layer [64][in], hidden [64][64]..., last [16][64]), row-major [out][in].
// input [n][in_w] half (column-major width x n), output [n][16] half.
// dL_doutput == nullptr: inference only. Otherwise forward + backward: dL_dinput [n][in_w], gradients half[n_params].
// iters > 1 repeats the call and returns the mean milliseconds per call in *ms (CUDA events on the NULL stream).
int ref_mlp_run(int in_w, int out_w, int n_hidden, const void* params, const void* input, uint32_t n, void* output,
                const void* dL_doutput, void* dL_dinput, void* gradients, int iters, float* ms) {
	try {
		FullyFusedMLP<T, 64> net((uint32_t)in_w, (uint32_t)out_w, (uint32_t)n_hidden, false, Activation::ReLU, Activation::None);
		net.set_params((T*)params, (T*)params, (T*)params, (T*)gradients);
		GPUMatrixDynamic<T> in((T*)input, (uint32_t)in_w, n, CM);
		GPUMatrixDynamic<T> out((T*)output, net.padded_output_width(), n, CM);
		cudaEvent_t e0, e1;
		cudaEventCreate(&e0); cudaEventCreate(&e1);
		if (iters < 1) iters = 1;
		for (int pass = 0; pass < 2; ++pass) { // pass 0 = warm-up (arena growth), pass 1 = timed
			const int reps = pass == 0 ? 1 : iters;
			cudaEventRecord(e0, nullptr);
			for (int it = 0; it < reps; ++it) {
				if (!dL_doutput) {
					net.inference_mixed_precision(nullptr, in, out, true);
				} else {
					GPUMatrixDynamic<T> dout((T*)dL_doutput, net.padded_output_width(), n, CM);
					GPUMatrixDynamic<T> din((T*)dL_dinput, (uint32_t)in_w, n, CM);
					auto ctx = net.forward(nullptr, in, &out, false, true);
					net.backward(nullptr, *ctx, in, out, dout, &din, false, EGradientMode::Overwrite);
				}
			}
			cudaEventRecord(e1, nullptr);
			cudaEventSynchronize(e1);
		}
		float t = 0.f;
		cudaEventElapsedTime(&t, e0, e1);
		if (ms) *ms = t / (float)iters;
		cudaEventDestroy(e0); cudaEventDestroy(e1);
		return (int)cudaDeviceSynchronize();
	} catch (const std::exception& e) {
		fprintf(stderr, "ref_mlp_run: %s\n", e.what());
		return -1;
	}
}

uint32_t ref_mlp_n_params(int in_w, int out_w, int n_hidden) {
	FullyFusedMLP<T, 64> net((uint32_t)in_w, (uint32_t)out_w, (uint32_t)n_hidden, false, Activation::ReLU, Activation::None);
	return (uint32_t)net.n_params();
}

}
