// TEST INFRASTRUCTURE ONLY -- see ngp_oracle.h. Whole-iteration CPU restatement of
// Testbed::train (src/testbed.cu:2527-2588) -> training_prep_nerf (src/testbed_nerf.cu:3388-3401)
// -> train_nerf (:2896-2968) -> train_nerf_step (:3138-3385) for the NeRF mode with
// configs/nerf/base.json. This is the unit bench.py times as the CPU baseline.
#include "ngp_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

namespace {
constexpr uint32_t GRID_CELLS = 128u * 128u * 128u;
constexpr float NERF_MIN_OPTICAL_THICKNESS = 0.01f;
constexpr float LOSS_SCALE = 128.0f; // testbed.h:272
inline uint32_t next_multiple(uint32_t v, uint32_t d) { return (v + d - 1) / d * d; }
inline orc_half f2h(float f) { _Float16 v = (_Float16)f; orc_half h; std::memcpy(&h, &v, 2); return h; }
}

struct orc_trainer {
	std::vector<orc_image> images;
	uint32_t aabb_scale = 1, max_cascade = 0;
	float aabb[6];
	float cone_angle_constant = 0.f;
	orc_model model;
	uint32_t n_params = 0;
	std::vector<float> w_fp32, m1, m2, grad;
	std::vector<orc_half> w_half, w_ema;
	std::vector<uint32_t> param_steps;
	orc_optimizer opt;
	std::vector<float> density_grid;
	std::vector<uint8_t> bitfield;
	float mean_density = 0.f;
	orc_pcg32 rng, density_grid_rng;
	uint32_t training_step = 0, density_grid_ema_step = 0;
	uint32_t rays_per_batch = 1u << 12; // testbed.h:374
	uint32_t measured_batch_size_before_compaction = 0, measured_batch_size = 0, n_rays_total = 0;
	// nerf.training defaults (testbed.h:640-676)
	int snap_to_pixel_centers = 1, random_bg_color = 1, linear_colors = 0, loss_type = 4 /*Huber*/;
	int rgb_activation = 2 /*Logistic*/, density_activation = 3 /*Exponential*/, color_space = 0 /*Linear: m_color_space default, testbed.h:846*/;
	float near_distance = 0.2f, density_grid_decay = 0.95f;
	float background_color[3] = {0.f, 0.f, 0.f};
};

// Parameter initialisation: Trainer ctor + initialize_params (tcnn trainer.h:53-99), xavier uniform
// for each MLP matrix (gpu_matrix.h:291-305), uniform +-1e-4 for the grid (grid.h:1364-1369, random.h:66-97).
static void init_params(orc_trainer* t, uint32_t seed) {
	std::seed_seq seq{seed};
	std::vector<uint32_t> seeds(2);
	seq.generate(seeds.begin(), seeds.end());
	orc_pcg32 rnd;
	orc_pcg32_seed(&rnd, seeds.front(), 1);
	const int shapes[5][2] = {{64, 32}, {16, 64}, {64, 32}, {64, 64}, {16, 64}};
	size_t pos = 0;
	for (auto& s : shapes) {
		float scale = std::sqrt(6.0f / (float)(s[0] + s[1]));
		for (int i = 0; i < s[0] * s[1]; ++i) t->w_fp32[pos++] = orc_pcg32_next_float(&rnd) * 2.0f * scale - scale;
	}
	const size_t n = t->model.n_grid_params;
	const size_t n_threads_needed = (n + 3) / 4;
	const size_t n_threads = (n_threads_needed + 127) / 128 * 128; // linear_kernel: 128 threads per block
	float* out = t->w_fp32.data() + ORC_MLP_PARAMS;
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)n_threads; ++i) {
		orc_pcg32 r = rnd;
		orc_pcg32_advance(&r, i * 4);
		for (size_t j = 0; j < 4; ++j) {
			size_t idx = (size_t)i + n_threads * j;
			if (idx >= n) break;
			out[idx] = std::fmaf(orc_pcg32_next_float(&r), 1e-4f - -1e-4f, -1e-4f); // the reference's device lambda is compiled with FMA contraction (pinned by tests/golden/ref_modes.npz)
		}
	}
	for (size_t i = 0; i < t->n_params; ++i) t->w_half[i] = f2h(t->w_fp32[i]);
}

extern "C" orc_trainer* orc_trainer_create(uint32_t n_images, const orc_image* images, uint32_t aabb_scale, uint32_t seed) {
	orc_trainer* t = new orc_trainer();
	t->images.assign(images, images + n_images);
	t->aabb_scale = aabb_scale;
	// load_nerf_post: src/testbed_nerf.cu:2714-2730
	float half = 0.5f * std::min(1u << 7, aabb_scale);
	for (int c = 0; c < 3; ++c) { t->aabb[c] = 0.5f - half; t->aabb[3 + c] = 0.5f + half; }
	t->max_cascade = 0;
	while ((1u << t->max_cascade) < aabb_scale) ++t->max_cascade;
	t->cone_angle_constant = aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);
	// reset_network: src/testbed.cu:2313-2325 (n_levels 16, base 16, log2_hashmap 19, desired resolution 2048)
	float per_level_scale = std::exp(std::log(2048.0f * (float)aabb_scale / 16.0f) / 15);
	orc_model_init(&t->model, 16, 19, 16, per_level_scale);
	t->n_params = ORC_MLP_PARAMS + t->model.n_grid_params;
	t->w_fp32.assign(t->n_params, 0.f); t->m1.assign(t->n_params, 0.f); t->m2.assign(t->n_params, 0.f); t->grad.assign(t->n_params, 0.f);
	t->w_half.assign(t->n_params, 0); t->w_ema.assign(t->n_params, 0); t->param_steps.assign(t->n_params, 0);
	orc_optimizer_init(&t->opt);
	init_params(t, seed);
	t->density_grid.assign((size_t)GRID_CELLS * (t->max_cascade + 1), 0.f);
	t->bitfield.assign((size_t)GRID_CELLS * 8 / 8, 0);
	// src/testbed.cu:2252,:2265
	orc_pcg32_seed(&t->rng, seed, 1);
	orc_pcg32_seed(&t->density_grid_rng, orc_pcg32_next_uint(&t->rng), 1);
	return t;
}
extern "C" void orc_trainer_destroy(orc_trainer* t) { delete t; }
extern "C" uint32_t orc_trainer_n_params(const orc_trainer* t) { return t->n_params; }
extern "C" uint32_t orc_trainer_step(const orc_trainer* t) { return t->training_step; }
extern "C" const uint8_t* orc_trainer_bitfield(const orc_trainer* t) { return t->bitfield.data(); }
extern "C" const float* orc_trainer_density_grid(const orc_trainer* t) { return t->density_grid.data(); }
extern "C" void orc_trainer_get_params(const orc_trainer* t, float* w_fp32, orc_half* w_half, orc_half* w_ema) {
	if (w_fp32) std::memcpy(w_fp32, t->w_fp32.data(), t->n_params * 4);
	if (w_half) std::memcpy(w_half, t->w_half.data(), t->n_params * 2);
	if (w_ema) std::memcpy(w_ema, t->w_ema.data(), t->n_params * 2);
}
extern "C" void orc_trainer_set_params(orc_trainer* t, const float* w_fp32) {
	std::memcpy(t->w_fp32.data(), w_fp32, t->n_params * 4);
	for (size_t i = 0; i < t->n_params; ++i) t->w_half[i] = f2h(t->w_fp32[i]);
}

extern "C" void orc_trainer_set_level_scales(orc_trainer* t, const float* scales) {
	for (uint32_t l = 0; l < t->model.n_levels; ++l) t->model.scales[l] = scales[l];
}

extern "C" void orc_trainer_set_state(orc_trainer* t, uint32_t training_step, uint32_t rays_per_batch, const float* density_grid) {
	t->training_step = training_step;
	t->density_grid_ema_step = training_step;
	if (rays_per_batch) t->rays_per_batch = rays_per_batch;
	if (density_grid) {
		std::memcpy(t->density_grid.data(), density_grid, t->density_grid.size() * sizeof(float));
		t->mean_density = orc_density_grid_mean(t->density_grid.data());
		orc_bitfield(t->max_cascade + 1, t->density_grid.data(), t->mean_density, t->bitfield.data());
	}
}

// Teacher forcing for the whole-iteration parity test: the state another implementation reached (Adam moments and step counters, the decay state, the
// batch-size controller's memory) replaces this trainer's, so that the next step starts from identical inputs.
extern "C" void orc_trainer_set_optimizer_state(orc_trainer* t, const float* m1, const float* m2, const uint32_t* param_steps, uint32_t optimizer_step, float lr_factor,
                                                uint32_t measured_batch_size_before_compaction, uint32_t n_rays_total) {
	std::memcpy(t->m1.data(), m1, t->n_params * 4);
	std::memcpy(t->m2.data(), m2, t->n_params * 4);
	std::memcpy(t->param_steps.data(), param_steps, t->n_params * 4);
	t->opt.step = optimizer_step; t->opt.lr_factor = lr_factor;
	t->measured_batch_size_before_compaction = measured_batch_size_before_compaction;
	t->n_rays_total = n_rays_total;
}

// update_density_grid_nerf + update_density_grid_mean_and_bitfield: src/testbed_nerf.cu:2761-2859
static void update_density_grid(orc_trainer* t, uint32_t n_uniform, uint32_t n_nonuniform) {
	const uint32_t n_elements = GRID_CELLS * (t->max_cascade + 1);
	const uint32_t n_samples = n_uniform + n_nonuniform;
	if (t->training_step == 0) {
		t->density_grid_ema_step = 0;
		orc_mark_untrained_density_grid(n_elements, t->density_grid.data(), (uint32_t)t->images.size(), t->images.data(), 1);
	}
	std::vector<float> positions((size_t)n_samples * 3);
	std::vector<uint32_t> indices(n_samples);
	orc_generate_grid_samples(n_uniform, t->density_grid_rng, t->density_grid_ema_step, t->aabb, t->density_grid.data(), positions.data(), indices.data(), t->max_cascade + 1, -0.01f);
	orc_pcg32_advance(&t->density_grid_rng, 1ll << 32);
	orc_generate_grid_samples(n_nonuniform, t->density_grid_rng, t->density_grid_ema_step, t->aabb, t->density_grid.data(), positions.data() + (size_t)n_uniform * 3, indices.data() + n_uniform, t->max_cascade + 1, NERF_MIN_OPTICAL_THICKNESS);
	orc_pcg32_advance(&t->density_grid_rng, 1ll << 32);
	std::vector<orc_half> density(n_samples);
	orc_nerf_density(&t->model, t->w_half.data(), n_samples, positions.data(), 3, density.data());
	orc_splat_and_ema(n_samples, indices.data(), density.data(), n_elements, t->density_grid_decay, t->density_grid.data());
	++t->density_grid_ema_step;
	t->mean_density = orc_density_grid_mean(t->density_grid.data());
	orc_bitfield(t->max_cascade + 1, t->density_grid.data(), t->mean_density, t->bitfield.data());
}

extern "C" void orc_trainer_train(orc_trainer* t, uint32_t batch_size, float* stats_out) {
	// Testbed::train: src/testbed.cu:2538-2554
	uint32_t n_prep_to_skip = std::max(1u, std::min(t->training_step / 16u, 16u));
	if (t->training_step % n_prep_to_skip == 0) {
		uint32_t n_cascades = t->max_cascade + 1;
		if (t->training_step < 256) update_density_grid(t, GRID_CELLS * n_cascades, 0);
		else update_density_grid(t, GRID_CELLS / 4 * n_cascades, GRID_CELLS / 4 * n_cascades);
	}
	const bool get_loss_scalar = t->training_step % 16 == 0;

	// train_nerf_step: src/testbed_nerf.cu:3138-3385
	const uint32_t max_samples = batch_size * 16;
	uint32_t max_inference;
	if (t->measured_batch_size_before_compaction == 0) {
		t->measured_batch_size_before_compaction = max_inference = max_samples;
	} else {
		max_inference = next_multiple(std::min(t->measured_batch_size_before_compaction, max_samples), 128);
	}
	if (t->training_step == 0) t->n_rays_total = 0;
	uint32_t n_rays_total = t->n_rays_total;
	t->n_rays_total += t->rays_per_batch;
	const uint32_t R = t->rays_per_batch;

	std::vector<uint32_t> ray_indices(R), numsteps((size_t)R * 2);
	std::vector<float> rays((size_t)R * 6), coords((size_t)max_inference * 7);
	uint32_t counters[2];
	uint32_t n_kept = orc_generate_training_samples(R, t->aabb, max_inference, n_rays_total, t->rng, (uint32_t)t->images.size(), t->images.data(),
		t->bitfield.data(), t->snap_to_pixel_centers, t->cone_angle_constant, ray_indices.data(), rays.data(), numsteps.data(), coords.data(), counters);
	uint32_t n_used = 0;
	for (uint32_t i = 0; i < n_kept; ++i) n_used = std::max(n_used, numsteps[i * 2 + 1] + numsteps[i * 2]);

	std::vector<orc_half> rgbsigma((size_t)std::max(n_used, 1u) * 4);
	orc_nerf_inference(&t->model, t->w_half.data(), n_used, coords.data(), rgbsigma.data());

	std::vector<float> coords_compacted((size_t)batch_size * 7, 0.f), loss(R, 0.f);
	std::vector<orc_half> dloss((size_t)batch_size * 4, 0);
	uint32_t compacted = orc_compute_loss(n_kept, R, t->aabb, n_rays_total, t->rng, batch_size, LOSS_SCALE, t->background_color, t->color_space,
		t->random_bg_color, t->linear_colors, (uint32_t)t->images.size(), t->images.data(), rgbsigma.data(), ray_indices.data(), rays.data(), numsteps.data(),
		coords.data(), coords_compacted.data(), dloss.data(), t->loss_type, loss.data(), t->rgb_activation, t->density_activation,
		t->snap_to_pixel_centers, t->mean_density, t->near_distance);
	orc_fill_rollover(batch_size, std::min(compacted, batch_size), coords_compacted.data(), dloss.data());

	orc_nerf_forward_backward(&t->model, t->w_half.data(), batch_size, coords_compacted.data(), dloss.data(), t->grad.data());
	orc_pcg32_advance(&t->rng, 1ll << 32);

	// train_nerf: optimizer step (:2950), then counters (:2870-2894)
	orc_optimizer_step(&t->opt, t->n_params, ORC_MLP_PARAMS, LOSS_SCALE, t->grad.data(), t->w_fp32.data(), t->w_half.data(), t->w_ema.data(),
		t->m1.data(), t->m2.data(), t->param_steps.data());
	++t->training_step;

	float loss_scalar = 0.f;
	t->measured_batch_size = 0;
	t->measured_batch_size_before_compaction = 0;
	if (counters[0] != 0 && compacted != 0) {
		t->measured_batch_size_before_compaction = counters[0];
		t->measured_batch_size = compacted;
		if (get_loss_scalar) {
			double s = 0.0;
			for (uint32_t i = 0; i < R; ++i) s += loss[i];
			loss_scalar = (float)s * (float)t->measured_batch_size / (float)batch_size;
		}
		t->rays_per_batch = (uint32_t)((float)t->rays_per_batch * (float)batch_size / (float)t->measured_batch_size);
		t->rays_per_batch = std::min(next_multiple(t->rays_per_batch, 128), 1u << 18);
	}
	if (stats_out) {
		stats_out[0] = loss_scalar; stats_out[1] = (float)R; stats_out[2] = (float)counters[0]; stats_out[3] = (float)compacted;
	}
}
