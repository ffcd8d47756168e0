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
#include "../../include/ngpb.h"

namespace ngpb {

struct AdamParams {
	uint32_t n, n_matrix;
	float loss_scale, base_lr, beta1, beta2, epsilon, l2_reg;
	float ema_decay, ema_debias_old, ema_debias_new;
};

__global__ void __launch_bounds__(256) adam_ema_kernel(const AdamParams P, float* __restrict__ grad, float* __restrict__ w_fp32, __half* __restrict__ w_half,
                                                       __half* __restrict__ w_ema, float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ param_steps)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P.n) return;
	float gradient = grad[i] / P.loss_scale;
	__half wh = w_half[i];
	const bool is_matrix = i < P.n_matrix;
	if (is_matrix || gradient != 0.f) {
		grad[i] = 0.f; // (a zero hash gradient is already zero)
		const float weight_fp = w_fp32[i];
		if (is_matrix) gradient += P.l2_reg * weight_fp; // L2 only on matrix params (adam.h:88-91)
		const float gradient_sq = gradient * gradient;
		const float first_moment = P.beta1 * m1[i] + (1 - P.beta1) * gradient;
		const float second_moment = P.beta2 * m2[i] + (1 - P.beta2) * gradient_sq;
		m1[i] = first_moment;
		m2[i] = second_moment;
		float learning_rate = P.base_lr;
		const uint32_t current_step = ++param_steps[i]; // per-parameter debiasing (adam.h:104)
		learning_rate *= sqrtf(1 - powf(P.beta2, (float)current_step)) / (1 - powf(P.beta1, (float)current_step));
		const float effective_learning_rate = fminf(fmaxf(learning_rate / (sqrtf(second_moment) + P.epsilon), 0.f), 3.402823466e+38f);
		const float new_weight = weight_fp - effective_learning_rate * first_moment; // weight decay terms are zero in nerf/base.json
		w_fp32[i] = new_weight;
		wh = __float2half_rn(new_weight);
		w_half[i] = wh;
	}
	const float filtered_val = (__half2float(w_ema[i]) * P.ema_decay * P.ema_debias_old + __half2float(wh) * (1 - P.ema_decay)) * P.ema_debias_new;
	w_ema[i] = __float2half_rn(filtered_val);
}

} // namespace ngpb

using namespace ngpb;

extern "C" void ngpb_optimizer_init(ngpb_optimizer* o) { // configs/nerf/base.json:5-22
	o->learning_rate = 1e-2f; o->beta1 = 0.9f; o->beta2 = 0.99f; o->epsilon = 1e-15f; o->l2_reg = 1e-6f; o->ema_decay = 0.95f;
	o->decay_start = 20000; o->decay_interval = 10000; o->decay_base = 0.33f;
	o->step = 0; o->lr_factor = 1.0f;
}

extern "C" int ngpb_optimizer_step(void* stream, ngpb_optimizer* o, uint32_t n_params, uint32_t n_matrix_params, float loss_scale, float* grad,
                                   float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema, float* m1, float* m2, uint32_t* param_steps) {
	try {
		if (!o || !grad || !w_fp32 || !w_half || !w_ema || !m1 || !m2 || !param_steps) { set_last_error("ngpb_optimizer_step: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		// ExponentialDecayOptimizer::step reads the step count before Adam increments it (exponential_decay.h:60-72)
		if (o->step == 0) o->lr_factor = 1.0f;
		if (o->step >= o->decay_start && o->decay_interval > 0 && (o->step - o->decay_start) % o->decay_interval == 0) o->lr_factor *= o->decay_base;
		AdamParams P;
		P.n = n_params; P.n_matrix = n_matrix_params; P.loss_scale = loss_scale;
		P.base_lr = o->learning_rate * o->lr_factor;
		P.beta1 = o->beta1; P.beta2 = o->beta2; P.epsilon = o->epsilon; P.l2_reg = o->l2_reg;
		++o->step; // AdamOptimizer::step (adam.h:152)
		// EmaOptimizer::step (ema.h:102-108)
		P.ema_decay = o->ema_decay;
		P.ema_debias_old = 1 - (float)std::pow(o->ema_decay, o->step - 1);
		P.ema_debias_new = 1.0f / (1 - (float)std::pow(o->ema_decay, o->step));
		if (n_params == 0) return 0;
		NGPB_STEP_KERNEL(adam_ema_kernel);
		adam_ema_kernel<<<div_round_up(n_params, 256), 256, 0, (cudaStream_t)stream>>>(P, grad, w_fp32, (__half*)w_half, (__half*)w_ema, m1, m2, param_steps);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
