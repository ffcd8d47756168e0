// `Testbed`-shaped host object for the NeRF mode (see testbed.cu). Opaque behind the C ABI.
#pragma once
#include "common.cuh"
#include "../../include/ngpb.h"

#include <vector>

struct ngpb_testbed {
	explicit ngpb_testbed(int device);
	~ngpb_testbed();
	ngpb_testbed(const ngpb_testbed&) = delete;
	ngpb_testbed& operator=(const ngpb_testbed&) = delete;

	void load_training_data(uint32_t n, const ngpb_host_image* host_images, uint32_t aabb_scale, bool allow_empty = false);
	// datasets filled while training runs (Testbed::create_empty_nerf_dataset, nerf.training.set_image / set_camera_intrinsics / n_images_for_training)
	void create_empty_dataset(uint32_t n, uint32_t aabb_scale);
	void set_training_image(uint32_t frame_idx, const ngpb_host_image& image);
	void set_camera_intrinsics(uint32_t frame_idx, float fx, float fy, float cx, float cy, float k1, float k2, float p1, float p2);
	void upload_image_table();
	uint32_t n_images_for_training = 0, n_images_for_training_prev = 0; // the first n images take part in training (testbed.h:660)
	std::vector<void*> own_pixels;                                        // per-image allocations made by set_training_image (null: the image lives in `pixels`)
	void reset_network(uint32_t seed);
	void train(uint32_t batch, bool sync_at_end = true);
	void update_density_grid(uint32_t n_uniform, uint32_t n_nonuniform);
	void ensure_workspace(uint32_t batch);
	void get_params(float* w_fp32, ngpb_half* w_half, ngpb_half* w_ema);
	void set_params(const float* w_fp32);
	void render(const float* camera12, int w, int h, float fx, float fy, int spp, bool linear, float* out_rgba, uint64_t* n_samples_out);

	// K1 of the next step, prefetched on a second stream (see train())
	struct SamplingRequest {
		uint32_t step, n_rays, max_inference; ngpb_rng rng; int snap; float cone_angle;
		ngpb_error_cdf cdf; // K19: the CDFs the pixels / images are drawn from (null members: uniform)
		bool operator==(const SamplingRequest& o) const {
			return cdf.cdf_x_cond_y == o.cdf.cdf_x_cond_y && cdf.cdf_y == o.cdf.cdf_y && cdf.cdf_img == o.cdf.cdf_img && cdf.res_x == o.cdf.res_x && cdf.res_y == o.cdf.res_y && step == o.step && n_rays == o.n_rays && max_inference == o.max_inference && rng.state == o.rng.state && rng.inc == o.rng.inc && snap == o.snap && cone_angle == o.cone_angle;
		}
	};
	void launch_sampling(cudaStream_t st, const SamplingRequest& p);
	void drop_prefetch();
	void collect_loss_scalar();
	cudaStream_t sampling_stream = nullptr;
	cudaEvent_t prefetch_done = nullptr, loss_ready = nullptr, counters_ready = nullptr, mlp_train_done = nullptr;
	SamplingRequest prefetch{};
	bool prefetch_valid = false, overlap_sampling = true;
	int dp_half_gradients = 1; // data parallel gradient exchange: 0 fp32, 1 bf16, 2 fp16
	__half* grad_half = nullptr;
	bool reuse_encoding = true; // the training pass starts from the inference pass's hash-grid features (compacted with the samples) instead of re-encoding
	bool loss_pending = false;
	float loss_pending_scale = 0.f;

	// data parallelism: every rank holds a replica and marches its shard of the global ray batch; gradients (and the two
	// batch-size counters) are summed over NCCL before the optimizer, so the replicas stay bit-identical.
	int dp_rank = 0, dp_world = 1;
	void* nccl_comm = nullptr;
	// Peer-memory exchange (dp_exchange = 1): gradients and weights move between the replicas' buffers with direct NVLink stores from this library's
	// own kernels instead of NCCL collectives. Buffers are shared through CUDA IPC handles (exchanged once over the NCCL communicator).
	int dp_exchange = 0;               // 0: NCCL reduce-scatter / all-gather, 1: peer-memory kernels
	bool p2p_ready = false;
	uint32_t p2p_step = 0;
	uint16_t* p2p_recv = nullptr;      // [world][count] bf16: slice r holds rank r's partial gradients of THIS rank's parameter range
	uint32_t* p2p_flags = nullptr;     // [2 * world + 2]: [r] = last step whose gradients rank r delivered, [world + r] = ... weights; [2w] block counter, [2w+1] error
	void* p2p_table[16][3] = {};       // {recv, flags, w_half} of every rank as mapped into this process (own entries = local pointers)
	std::vector<void*> p2p_opened;     // IPC mappings to close
	void p2p_setup();
	void p2p_teardown();
	// the EMA sweep of the sharded optimizer runs on its own stream, under the next step's first half (only renders / snapshots read the EMA copy)
	cudaStream_t ema_stream = nullptr;
	cudaEvent_t weights_gathered = nullptr, ema_done = nullptr;
	bool ema_pending = false;
	void launch_ema_sweep_async(const void* opt_params);
	void join_ema(); // makes `stream` wait for a pending EMA sweep (before anything that overwrites the fp16 weights or reads the EMA copy)
	bool dp_sharded_optimizer = true;    // reduce-scatter + Adam on 1/world of the parameters + all-gather (false: all-reduce + full Adam)
	bool master_weights_sharded = false; // the fp32 master copy is current only in this rank's range
	uint32_t dp_shard_count() const;
	void init_data_parallel(int rank, int world, const void* unique_id128);
	uint32_t inference_budget(uint32_t measured_before_compaction) const;

	void* dalloc(size_t bytes);
	void dfree(void* p);

	int device = 0;
	cudaStream_t stream = nullptr; // everything runs on this stream (Testbed::m_stream, testbed.h:895)
	std::vector<void*> allocations;

	// dataset (NerfDataset, nerf_loader.h:66-110)
	bool training_data_available = false, scene_configured = false;
	void configure_scene_box();
	std::vector<ngpb_image> images;
	ngpb_image* images_dev = nullptr;
	uint8_t* pixels = nullptr;
	size_t pixels_bytes = 0, density_grid_cells = 0;
	uint32_t aabb_scale = 1, max_cascade = 0;
	float aabb[6] = {0, 0, 0, 1, 1, 1};
	float cone_angle_constant = 0.f;

	// camera-extrinsics optimisation (Testbed::Nerf::Training, testbed.h:620-650; train_nerf :3056-3083): per-camera position / rotation offsets moved by a
	// host-side Adam every n_steps_between_cam_updates steps from gradients accumulated on the device (camera_optimizer.cu)
	bool optimize_extrinsics = false;
	float extrinsic_learning_rate = 1e-3f, extrinsic_l2_reg = 1e-4f;
	uint32_t n_steps_since_cam_update = 0, n_steps_between_cam_updates = 16;
	std::vector<float> dataset_xforms;                // [n_images][12]: the transforms as loaded (dataset.xforms); images[i].raw_xform = these + offsets
	std::vector<float> cam_pos_state, cam_rot_state;  // [n_images][10]: {iter, m1[3], m2[3], variable[3]} of ngpb_camera_adam_step
	std::vector<float> cam_gradients_host;            // [3][n_images][3]
	float* cam_gradients = nullptr;                   // device, [3][n_images][3]: position, rotation, exposure
	// per-image exposure optimisation (train_nerf :3105-3131): 2^exposure scales the image's colours in the loss (K6 :1403); the gradient accumulates in
	// cam_gradients[2] and a host Adam (learning rate of the network optimizer) moves the exposures, which are then re-centred on a zero mean
	bool optimize_exposure = false;
	float exposure_l2_reg = 0.0f;
	bool exposure_active = false;                     // the loss kernels read cam_exposure (set once exposures are optimised or set)
	std::vector<float> cam_exposure_state;            // [n_images][10], as cam_pos_state
	float* cam_exposure = nullptr;                    // device, [n_images][3]
	void upload_exposures();
	// K19, error-map importance sampling (Testbed::Nerf::Training::ErrorMap testbed.h:600-615, train_nerf :2933-2939 and :2971-3023): every ray's loss is
	// deposited into a low-resolution map per image; every n_steps_between_error_map_updates steps the map becomes a set of CDFs the next windows draw their
	// pixels / images from. The counters always follow the reference's cadence; the map is only accumulated while one of the two switches is on (the
	// reference always accumulates, for its GUI -- without the switches nothing reads the result).
	bool sample_focal_plane_proportional_to_error = false, sample_image_proportional_to_error = false;
	uint32_t n_steps_between_error_map_updates = 128, n_steps_since_error_map_update = 0;
	bool error_map_live = false;   // the map was cleared at the start of the current window and receives every step's losses
	bool error_cdf_valid = false;  // is_cdf_valid
	bool error_cdf_used = false;   // the CDFs have been handed to a kernel since they were built (their buffers may not be re-sized without a sync)
	int error_map_res[2] = {0, 0}, error_cdf_res[2] = {0, 0};
	float *error_map = nullptr, *error_cdf_x_cond_y = nullptr, *error_cdf_y = nullptr, *error_cdf_img = nullptr, *error_pmf_img = nullptr;
	size_t error_map_capacity = 0, error_cdf_capacity = 0; // floats
	cudaEvent_t error_cdf_built = nullptr;
	ngpb_error_cdf sampling_cdf() const; // what K1 / K6 / the camera gradient receive this step (null members unless valid and switched on)
	void begin_error_map_window();
	void finish_error_map_window();
	float* coords_gradient = nullptr; __half* dL_dsh = nullptr; // workspace, sized by the batch (allocated on first use)
	void reset_camera_extrinsics();
	void update_transforms();
	void camera_update_step();

	// model + optimizer state: fp32 master, fp16 training copy, fp16 EMA (inference) copy, Adam moments
	// (tcnn Trainer buffer trainer.h:80,:317-332). Flat order: density net, rgb net, grid levels.
	ngpb_grid grid{};
	uint32_t log2_hashmap_size = 19; // encoding.log2_hashmap_size of the network config (configs/nerf/base.json:27); takes effect at reset_network
	ngpb_optimizer opt_hyper{};      // hyper-parameters of the network config's optimizer section, re-applied by every reset_network
	uint32_t n_params = 0, n_alloc = 0;
	float* w_fp32 = nullptr; __half* w_half = nullptr; __half* w_ema = nullptr;
	float* m1 = nullptr; float* m2 = nullptr; uint32_t* param_steps = nullptr; float* grad = nullptr;
	ngpb_optimizer opt{};

	// occupancy grid (testbed.h:698-704)
	float* density_grid = nullptr; float* density_grid_tmp = nullptr; uint8_t* bitfield = nullptr; float* mean_density = nullptr;
	float* dg_positions = nullptr; uint32_t* dg_indices = nullptr; __half* dg_density = nullptr; __half* dg_encoded = nullptr;
	float density_grid_decay = 0.95f;
	uint32_t density_grid_ema_step = 0;

	// per-iteration workspace, sized by the batch (train_nerf_step scratch, testbed_nerf.cu:3145-3170)
	uint32_t ws_batch = 0;
	uint32_t* ray_indices = nullptr; float* rays = nullptr; uint32_t* numsteps = nullptr; float* coords = nullptr;
	__half* rgbsigma = nullptr; __half* encoded = nullptr; __half* encoded_compacted = nullptr; float* coords_compacted = nullptr; __half* dloss = nullptr; __half* denc = nullptr;
	float* loss = nullptr; void* scratch = nullptr; uint32_t* counters = nullptr; float* partials = nullptr;
	uint32_t* host_readback = nullptr;

	// training state (NerfCounters testbed.h:366-386; m_rng testbed.h:923)
	uint32_t seed = 1337; // m_seed, testbed.h:567
	ngpb::Pcg32 rng{}, density_grid_rng{};
	uint32_t training_step = 0;
	uint32_t rays_per_batch = 1u << 12;
	uint32_t measured_batch_size = 0, measured_batch_size_before_compaction = 0, n_rays_total = 0;
	float loss_scalar = 0.f;
	bool shall_train = true;
	ngpb_loss_config loss_cfg{};
	float render_min_transmittance = 0.01f; // testbed.h:725
	// render state (testbed.h:547,:853,:875,:889)
	int render_snap_to_pixel_centers = 0;
	float render_near_distance = 0.f, exposure = 0.f, background_alpha = 1.0f;
	int tonemap_curve = NGPB_TONEMAP_IDENTITY; // m_tonemap_curve, testbed.h:847
	bool render_with_training_params = false;
	void* render_ws = nullptr; size_t render_ws_bytes = 0;
	double last_render_ms = 0.0;

	uint64_t n_launches = 0, h2d_bytes = 0, d2h_bytes = 0;

	// Optional per-stage device timing (CUDA events on `stream`), read by bench.py for the roofline of each kernel.
	// Off by default: when on, two event records bracket every stage of train().
	bool profile_stages = false;
	cudaEvent_t stage_ev[NGPB_N_STAGES][2] = {};
	bool stage_used[NGPB_N_STAGES] = {};
	double stage_ms[NGPB_N_STAGES] = {};
	uint64_t stage_calls[NGPB_N_STAGES] = {};
	uint64_t stage_units[NGPB_N_STAGES] = {}; // samples (or rays / params) the stage processed, summed over calls
	void stage_begin(int s, cudaStream_t st);
	void stage_end(int s, uint64_t units, cudaStream_t st);
	void stage_collect(); // after a stream synchronize
};
