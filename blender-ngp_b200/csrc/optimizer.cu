// K15: Ema(ExponentialDecay(Adam)) as ONE pass over the parameters.
// Replaces adam_step (reference: dependencies/tiny-cuda-nn/include/tiny-cuda-nn/optimizers/adam.h:48-119),
// ema_step_half_precision (ema.h:63-76), the host-side schedule in exponential_decay.h:60-72 and
// ema.h:102-118, and the gradient memset that precedes the next backward pass (grid.h:1154).
//
// The reference makes three dense sweeps per iteration over all 12.2 M parameters (memset of the
// gradients, Adam, EMA). Here it is one sweep: read grad (fp32), update touched parameters, write
// the fp16 training copy and the fp16 EMA copy, and zero the gradient for the next iteration.
// Hash-grid entries whose gradient is exactly zero are skipped by Adam (adam.h:76-79) but still take
// part in the EMA, exactly as in the reference.
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "../../include/ngpb.h"

namespace ngpb {

struct AdamParams {
	uint32_t n, n_matrix;
	float loss_scale, base_lr, beta1, beta2, epsilon, l2_reg;
	float log2_beta1, log2_beta2;
	float ema_decay, ema_debias_old, ema_debias_new;
	uint32_t do_ema; // 0: the EMA is swept separately (sharded data-parallel optimizer)
	float inv_loss_scale; // 1 / loss_scale when loss_scale is a power of two (the division is then an exact scaling), else 0
};

// The debiased learning rate depends only on the parameter's step count; the two features of a grid entry -- usually all four parameters of a
// thread's group -- share it, so it is computed once per distinct count (two exp2f, a square root and a division per parameter otherwise).
struct LrCache { float step = -1.f, lr = 0.f; };

// One parameter. Returns the fp16 value the EMA filters (the updated weight, or the unchanged one when Adam skips the entry).
__device__ __forceinline__ void adam_ema_one(const AdamParams& P, bool is_matrix, float& grad, float& w, float& m1, float& m2, uint32_t& steps, __half& wh, __half& ema, LrCache& cache) {
	float gradient = P.inv_loss_scale != 0.f ? grad * P.inv_loss_scale : grad / P.loss_scale;
	if (is_matrix || gradient != 0.f) { // hash-grid entries with a zero gradient are skipped (adam.h:76-79)
		grad = 0.f;
		if (is_matrix) gradient += P.l2_reg * w; // L2 only on matrix params (adam.h:88-91)
		const float gradient_sq = gradient * gradient;
		m1 = P.beta1 * m1 + (1 - P.beta1) * gradient;
		m2 = P.beta2 * m2 + (1 - P.beta2) * gradient_sq;
		const float step = (float)(++steps); // per-parameter debiasing (adam.h:104)
		// beta^step as exp2(step * log2(beta)): the reference calls powf here; the two agree to ~1e-6 relative, far inside the fp32
		// resolution of the resulting weight change, and powf was two thirds of this kernel's instructions
		if (step != cache.step) { cache.step = step; cache.lr = P.base_lr * sqrtf(1 - exp2f(step * P.log2_beta2)) / (1 - exp2f(step * P.log2_beta1)); }
		const float learning_rate = cache.lr;
		const float effective_learning_rate = fminf(fmaxf(learning_rate / (sqrtf(m2) + P.epsilon), 0.f), 3.402823466e+38f);
		w = w - effective_learning_rate * m1; // weight decay terms are zero in nerf/base.json
		wh = __float2half_rn(w);
	}
	if (P.do_ema) ema = __float2half_rn((__half2float(ema) * P.ema_decay * P.ema_debias_old + __half2float(wh) * (1 - P.ema_decay)) * P.ema_debias_new);
}

// Four parameters per thread (128-bit accesses on the fp32 arrays, 64-bit on the fp16 ones); n4 = n / 4 full groups, the tail is scalar.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) adam_ema_kernel(const AdamParams P, float* __restrict__ grad, float* __restrict__ w_fp32, __half* __restrict__ w_half,
                                                       __half* __restrict__ w_ema, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ param_steps)
{
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n4 = P.n / 4;
	if (q < n4) {
		float4 g = reinterpret_cast<float4*>(grad)[q];
		const uint2 whr = reinterpret_cast<const uint2*>(w_half)[q];
		uint2 emr = reinterpret_cast<uint2*>(w_ema)[q];
		__half wh[4], em[4];
		*reinterpret_cast<uint2*>(wh) = whr; *reinterpret_cast<uint2*>(em) = emr;
		float gv[4] = {g.x, g.y, g.z, g.w};
		const uint32_t i0 = q * 4;
		bool any = i0 < P.n_matrix;
		#pragma unroll
		for (int k = 0; k < 4; ++k) any |= gv[k] != 0.f;
		if (any) { // at least one entry takes an Adam step: bring in the optimizer state of the group
			float4 w = reinterpret_cast<float4*>(w_fp32)[q], a = reinterpret_cast<float4*>(m1)[q], b = reinterpret_cast<float4*>(m2)[q];
			uint4 st = reinterpret_cast<uint4*>(param_steps)[q];
			float wv[4] = {w.x, w.y, w.z, w.w}, av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
			uint32_t sv[4] = {st.x, st.y, st.z, st.w};
			LrCache cache;
			#pragma unroll
			for (int k = 0; k < 4; ++k) adam_ema_one(P, i0 + k < P.n_matrix, gv[k], wv[k], av[k], bv[k], sv[k], wh[k], em[k], cache);
			reinterpret_cast<float4*>(grad)[q] = make_float4(gv[0], gv[1], gv[2], gv[3]);
			reinterpret_cast<float4*>(w_fp32)[q] = make_float4(wv[0], wv[1], wv[2], wv[3]);
			reinterpret_cast<float4*>(m1)[q] = make_float4(av[0], av[1], av[2], av[3]);
			reinterpret_cast<float4*>(m2)[q] = make_float4(bv[0], bv[1], bv[2], bv[3]);
			reinterpret_cast<uint4*>(param_steps)[q] = make_uint4(sv[0], sv[1], sv[2], sv[3]);
			reinterpret_cast<uint2*>(w_half)[q] = *reinterpret_cast<uint2*>(wh);
		} else if (P.do_ema) { // only the EMA moves
			#pragma unroll
			for (int k = 0; k < 4; ++k) em[k] = __float2half_rn((__half2float(em[k]) * P.ema_decay * P.ema_debias_old + __half2float(wh[k]) * (1 - P.ema_decay)) * P.ema_debias_new);
		}
		if (P.do_ema) reinterpret_cast<uint2*>(w_ema)[q] = *reinterpret_cast<uint2*>(em);
	} else {
		const uint32_t i = n4 * 4 + (q - n4);
		LrCache cache;
		if (i < P.n) adam_ema_one(P, i < P.n_matrix, grad[i], w_fp32[i], m1[i], m2[i], param_steps[i], w_half[i], w_ema[i], cache);
	}
}

// EMA of the fp16 weights alone (ema_step_half_precision, ema.h:63-76), 8 parameters per thread; n rounded up to a multiple of 8 by the caller's padding.
__global__ void __launch_bounds__(256) ema_sweep_kernel(const AdamParams P, const uint32_t n8, const __half* __restrict__ w_half, __half* __restrict__ w_ema)
{
	const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n8) return;
	const uint4 wr = reinterpret_cast<const uint4*>(w_half)[q];
	uint4 er = reinterpret_cast<uint4*>(w_ema)[q];
	const __half2* w2 = reinterpret_cast<const __half2*>(&wr);
	__half2* e2 = reinterpret_cast<__half2*>(&er);
	#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const float2 w = __half22float2(w2[k]), e = __half22float2(e2[k]);
		e2[k] = __floats2half2_rn((e.x * P.ema_decay * P.ema_debias_old + w.x * (1 - P.ema_decay)) * P.ema_debias_new,
		                          (e.y * P.ema_decay * P.ema_debias_old + w.y * (1 - P.ema_decay)) * P.ema_debias_new);
	}
	reinterpret_cast<uint4*>(w_ema)[q] = er;
}

} // namespace ngpb

using namespace ngpb;

extern "C" void ngpb_optimizer_init(ngpb_optimizer* o) { // configs/nerf/base.json:5-22
	o->learning_rate = 1e-2f; o->beta1 = 0.9f; o->beta2 = 0.99f; o->epsilon = 1e-15f; o->l2_reg = 1e-6f; o->ema_decay = 0.95f;
	o->decay_start = 20000; o->decay_interval = 10000; o->decay_base = 0.33f;
	o->step = 0; o->lr_factor = 1.0f;
}

namespace ngpb {
// Host part of one optimizer step: the learning-rate schedule and EMA debiasing factors (advances o->step). Returned as an opaque blob of
// sizeof(AdamParams) so that the testbed can launch the parameter sweep in chunks (data-parallel pipeline, testbed.cu).
void optimizer_prepare(ngpb_optimizer* o, float loss_scale, void* params_out) {
	// ExponentialDecayOptimizer::step reads the step count before Adam increments it (exponential_decay.h:60-72)
	if (o->step == 0) o->lr_factor = 1.0f;
	if (o->step >= o->decay_start && o->decay_interval > 0 && (o->step - o->decay_start) % o->decay_interval == 0) o->lr_factor *= o->decay_base;
	AdamParams P{};
	P.loss_scale = loss_scale;
	{ int e = 0; const float mant = std::frexp(loss_scale, &e); P.inv_loss_scale = (loss_scale > 0.f && mant == 0.5f) ? 1.0f / loss_scale : 0.f; }
	P.base_lr = o->learning_rate * o->lr_factor;
	P.beta1 = o->beta1; P.beta2 = o->beta2; P.epsilon = o->epsilon; P.l2_reg = o->l2_reg;
	P.log2_beta1 = (float)std::log2((double)o->beta1); P.log2_beta2 = (float)std::log2((double)o->beta2);
	++o->step; // AdamOptimizer::step (adam.h:152)
	// EmaOptimizer::step (ema.h:102-108)
	P.ema_decay = o->ema_decay;
	P.ema_debias_old = 1 - (float)std::pow(o->ema_decay, o->step - 1);
	P.ema_debias_new = 1.0f / (1 - (float)std::pow(o->ema_decay, o->step));
	P.do_ema = 1;
	std::memcpy(params_out, &P, sizeof(P));
}
void optimizer_disable_fused_ema(void* params) { AdamParams P; std::memcpy(&P, params, sizeof(P)); P.do_ema = 0; std::memcpy(params, &P, sizeof(P)); }
// EMA over parameters [0, n_padded), n_padded a multiple of 8 (the arrays are padded)
void ema_sweep_launch(cudaStream_t stream, const void* params, uint32_t n_padded, const __half* w_half, __half* w_ema) {
	AdamParams P;
	std::memcpy(&P, params, sizeof(P));
	NGPB_STEP_KERNEL(ema_sweep_kernel);
	ema_sweep_kernel<<<div_round_up(n_padded / 8, 256), 256, 0, stream>>>(P, n_padded / 8, w_half, w_ema);
	NGPB_LAUNCH_CHECK();
}
size_t optimizer_params_bytes() { return sizeof(AdamParams); }
// Sweeps parameters [first, first + count) (first and count multiples of 4 except for the last range).
void optimizer_launch(cudaStream_t stream, const void* params, uint32_t first, uint32_t count, uint32_t n_matrix_params, float* grad, float* w_fp32, __half* w_half,
                      __half* w_ema, float* m1, float* m2, uint32_t* param_steps) {
	if (count == 0) return;
	AdamParams P;
	std::memcpy(&P, params, sizeof(P));
	P.n = count;
	P.n_matrix = n_matrix_params > first ? n_matrix_params - first : 0u;
	// The sweep is a dependent pair of DRAM round trips per thread (gradient -> optimizer state), so bytes in flight = resident threads x 32 B:
	// NGPB_ADAM_BLOCKS picks the occupancy target (registers per thread follow from it).
	static const int min_blocks = [] { const char* e = std::getenv("NGPB_ADAM_BLOCKS"); return e ? std::atoi(e) : 6; }();
	const uint32_t blocks = div_round_up(count / 4 + count % 4, 256);
	if (min_blocks >= 8) adam_ema_kernel<8><<<blocks, 256, 0, stream>>>(P, grad + first, w_fp32 + first, w_half + first, w_ema + first, m1 + first, m2 + first, param_steps + first);
	else if (min_blocks >= 6) adam_ema_kernel<6><<<blocks, 256, 0, stream>>>(P, grad + first, w_fp32 + first, w_half + first, w_ema + first, m1 + first, m2 + first, param_steps + first);
	else adam_ema_kernel<4><<<blocks, 256, 0, stream>>>(P, grad + first, w_fp32 + first, w_half + first, w_ema + first, m1 + first, m2 + first, param_steps + first);
	NGPB_LAUNCH_CHECK();
}
} // namespace ngpb

extern "C" int ngpb_optimizer_step(void* stream, ngpb_optimizer* o, uint32_t n_params, uint32_t n_matrix_params, float loss_scale, float* grad,
                                   float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema, float* m1, float* m2, uint32_t* param_steps) {
	try {
		if (!o || !grad || !w_fp32 || !w_half || !w_ema || !m1 || !m2 || !param_steps) { set_last_error("ngpb_optimizer_step: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		AdamParams P;
		optimizer_prepare(o, loss_scale, &P);
		optimizer_launch((cudaStream_t)stream, &P, 0, n_params, n_matrix_params, grad, w_fp32, (__half*)w_half, (__half*)w_ema, m1, m2, param_steps);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
