// Element-wise training losses of the neural-image and SDF models. Replaces tcnn's l2_loss (dependencies/tiny-cuda-nn/include/tiny-cuda-nn/losses/l2.h:40-80,
// configs/image/base.json), relative_l2_loss (losses/relative_l2.h:40-77) and mape_loss (losses/mape.h:40-80, configs/sdf/base.json) as called by Trainer::training_step (trainer.h:121-159): from the
// network's padded fp16 output [n][16] and the fp32 targets [n][dims] it writes the per-element loss values and dL/d(output) (fp16, loss scale applied,
// padded columns zero) that ngpb_mlp_forward_backward consumes.
#include "common.cuh"
#include "../../include/ngpb.h"

namespace ngpb {

template <int KIND>
__global__ void __launch_bounds__(256) loss_kernel(const uint32_t n_elements, const uint32_t dims, const float loss_scale, const __half* __restrict__ predictions,
                                                   const float* __restrict__ targets, float* __restrict__ values, __half* __restrict__ gradients)
{
	constexpr uint32_t stride = 16; // padded output width of the fully fused MLP
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t intra = i % stride, inter = i / stride;
	if (intra >= dims) {
		if (values) values[i] = 0.f;
		gradients[i] = __float2half_rn(0.f);
		return;
	}
	const uint32_t n_total = n_elements / stride * dims;
	const float prediction = __half2float(predictions[i]);
	const float target = targets[inter * dims + intra];
	const float difference = prediction - target;
	float value, gradient;
	if (KIND == NGPB_ELEMENT_LOSS_MAPE) {
		const float scale = 1.0f / (fabsf(target) + 1e-2f);
		value = fabsf(difference) * scale / n_total;
		gradient = copysignf(scale, difference);
	} else if (KIND == NGPB_ELEMENT_LOSS_RELATIVE_L2) { // losses/relative_l2.h:40-77
		const float prediction_sq_plus_epsilon = prediction * prediction + 0.01f;
		value = difference * difference / prediction_sq_plus_epsilon / 1.0f / n_total;
		gradient = 2 * difference / prediction_sq_plus_epsilon / 1.0f;
	} else {
		value = difference * difference / n_total;
		gradient = 2 * difference;
	}
	if (values) values[i] = value;
	gradients[i] = __float2half_rn(loss_scale * gradient / n_total);
}

} // namespace ngpb

using namespace ngpb;

extern "C" int ngpb_loss(void* stream_, int kind, uint32_t n, uint32_t dims, float loss_scale, const ngpb_half* predictions, const float* targets, float* values,
                         ngpb_half* gradients) {
	try {
		if (!predictions || !targets || !gradients || dims == 0 || dims > 16 || (kind != NGPB_ELEMENT_LOSS_L2 && kind != NGPB_ELEMENT_LOSS_MAPE && kind != NGPB_ELEMENT_LOSS_RELATIVE_L2)) {
			set_last_error("ngpb_loss: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		if (n == 0) return 0;
		if (n > (0xFFFFFFFFu >> 4)) { set_last_error("ngpb_loss: batch too large (n * 16 must fit 32 bits)"); return NGPB_ERR_INVALID_ARGUMENT; }
		cudaStream_t stream = (cudaStream_t)stream_;
		const uint32_t n_elements = n * 16;
		if (kind == NGPB_ELEMENT_LOSS_MAPE) loss_kernel<NGPB_ELEMENT_LOSS_MAPE><<<div_round_up(n_elements, 256), 256, 0, stream>>>(n_elements, dims, loss_scale, (const __half*)predictions, targets, values, (__half*)gradients);
		else if (kind == NGPB_ELEMENT_LOSS_RELATIVE_L2) loss_kernel<NGPB_ELEMENT_LOSS_RELATIVE_L2><<<div_round_up(n_elements, 256), 256, 0, stream>>>(n_elements, dims, loss_scale, (const __half*)predictions, targets, values, (__half*)gradients);
		else loss_kernel<NGPB_ELEMENT_LOSS_L2><<<div_round_up(n_elements, 256), 256, 0, stream>>>(n_elements, dims, loss_scale, (const __half*)predictions, targets, values, (__half*)gradients);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
