// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Host-only launcher around the reference's own per-camera optimizers, compiled from the header where it lies under /root/reference:
//   ngp::AdamOptimizer<Eigen::Vector3f / Eigen::Array3f>, ngp::RotationAdamOptimizer    include/neural-graphics-primitives/adam_optimizer.h:20-159
// and the Eigen expression Training::update_transforms applies to a camera (src/testbed_nerf.cu:2614-2627), evaluated with the reference's Eigen.
// Needs no GPU: oracle/gen_golden.py runs it in the build container to produce tests/golden/ref_camera_adam.npz.
#include <neural-graphics-primitives/adam_optimizer.h>
#include <vector>

using namespace ngp;
using namespace Eigen;

extern "C" {

// Runs n_steps of the optimizer on the given gradients (n_steps x 3) with the per-step learning rates; writes the variable after every step (n_steps x 3).
int ref_camera_adam(int rotation, int n_steps, const float* gradients, const float* learning_rates, float* variables_out) {
	if (rotation) {
		RotationAdamOptimizer opt(1e-3f);
		for (int i = 0; i < n_steps; ++i) {
			opt.set_learning_rate(learning_rates[i]);
			opt.step(Vector3f{gradients[i * 3], gradients[i * 3 + 1], gradients[i * 3 + 2]});
			for (int c = 0; c < 3; ++c) variables_out[i * 3 + c] = opt.variable()[c];
		}
	} else {
		AdamOptimizer<Vector3f> opt(1e-3f);
		for (int i = 0; i < n_steps; ++i) {
			opt.set_learning_rate(learning_rates[i]);
			opt.step(Vector3f{gradients[i * 3], gradients[i * 3 + 1], gradients[i * 3 + 2]});
			for (int c = 0; c < 3; ++c) variables_out[i * 3 + c] = opt.variable()[c];
		}
	}
	return 0;
}

// The exposure block of Testbed::train_nerf (src/testbed_nerf.cu:3105-3131) with the reference's AdamOptimizer<Array3f>: n_updates camera updates over
// n_images images; gradients [n_updates][n_images][3] (as read back from the device), learning_rates [n_updates]; writes the exposures after each
// update [n_updates][n_images][3].
int ref_exposure_updates(int n_images, int n_updates, const float* gradients, const float* learning_rates, float per_camera_loss_scale, float l2_reg, float* exposures_out) {
	std::vector<AdamOptimizer<Array3f>> cam_exposure(n_images, AdamOptimizer<Array3f>(1e-3f, Array3f::Zero()));
	for (int u = 0; u < n_updates; ++u) {
		Array3f mean_exposure = Array3f::Constant(0.0f);
		for (int i = 0; i < n_images; ++i) {
			const float* gp = gradients + ((size_t)u * n_images + i) * 3;
			Array3f gradient = Array3f{gp[0], gp[1], gp[2]} * per_camera_loss_scale;
			gradient += cam_exposure[i].variable() * l2_reg;
			cam_exposure[i].set_learning_rate(learning_rates[u]);
			cam_exposure[i].step(gradient);
			mean_exposure += cam_exposure[i].variable();
		}
		mean_exposure /= n_images;
		for (int i = 0; i < n_images; ++i) {
			Array3f e = cam_exposure[i].variable() -= mean_exposure;
			for (int c = 0; c < 3; ++c) exposures_out[((size_t)u * n_images + i) * 3 + c] = e[c];
		}
	}
	return 0;
}

// xform: 3x4 column-major in / out; the statements of update_transforms for one camera.
int ref_apply_camera_offsets(const float* xform12, const float* pos3, const float* rot3, float* out12) {
	Matrix<float, 3, 4> xform;
	for (int k = 0; k < 12; ++k) xform.data()[k] = xform12[k];
	Vector3f rot{rot3[0], rot3[1], rot3[2]};
	float angle = rot.norm();
	rot /= angle;
	if (angle > 0) xform.block<3, 3>(0, 0) = AngleAxisf(angle, rot) * xform.block<3, 3>(0, 0);
	xform.col(3) += Vector3f{pos3[0], pos3[1], pos3[2]};
	for (int k = 0; k < 12; ++k) out12[k] = xform.data()[k];
	return 0;
}

}
