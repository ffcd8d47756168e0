// TEST INFRASTRUCTURE ONLY. C entry points around the UNMODIFIED reference ngp::Testbed (headless build, oracle/Makefile.full), driven the
// way src/python_api.cu drives it. Used on the GPU box to (a) write snapshots with the reference's own save_snapshot, (b) render frames
// with the reference's render_frame / bl_render_frame from those snapshots, (c) time the reference's training loop. The outputs are committed as
// fixtures under tests/golden/ (oracle/gen_golden_full.py); nothing in the product links or loads this file.
#include <neural-graphics-primitives/testbed.h>
#include <neural-graphics-primitives/nerf/render_request.cuh>
#include <neural-graphics-primitives/nerf/mask_3D.cuh>

#include <json/json.hpp>

#include <cstring>
#include <string>

using namespace ngp;
using namespace Eigen;

static thread_local std::string g_err;
#define REFF_BEGIN try {
#define REFF_END } catch (const std::exception& e) { g_err = e.what(); return 1; } return 0;

static Matrix<float, 3, 4> mat34(const float* m) { // column-major 3x4
	Matrix<float, 3, 4> r;
	for (int c = 0; c < 4; ++c) for (int k = 0; k < 3; ++k) r(k, c) = m[c * 3 + k];
	return r;
}

extern "C" {

const char* reff_last_error() { return g_err.c_str(); }

int reff_create(void** out, int mode) {
	REFF_BEGIN
	*out = new Testbed((ETestbedMode)mode);
	REFF_END
}
void reff_destroy(void* t) { delete (Testbed*)t; }

int reff_load_training_data(void* t, const char* path) { REFF_BEGIN ((Testbed*)t)->load_training_data(path); REFF_END }
int reff_reload_network_from_json(void* t, const char* json_text) {
	REFF_BEGIN
	((Testbed*)t)->reload_network_from_json(nlohmann::json::parse(json_text, nullptr, true, true), "");
	REFF_END
}
// Testbed::train(batch) n times (python_api.cu:594); returns the loss EMA value
int reff_train(void* tp, uint32_t batch, uint32_t n_steps, float* loss_out) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	for (uint32_t i = 0; i < n_steps; ++i) t->train(batch);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	if (loss_out) *loss_out = t->m_loss_scalar.val();
	REFF_END
}
uint32_t reff_training_step(void* t) { return ((Testbed*)t)->m_training_step; }
// stats: rays_per_batch, measured_batch_size, measured_batch_size_before_compaction, n_params
int reff_stats(void* tp, uint64_t* s) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	s[0] = t->m_nerf.training.counters_rgb.rays_per_batch;
	s[1] = t->m_nerf.training.counters_rgb.measured_batch_size;
	s[2] = t->m_nerf.training.counters_rgb.measured_batch_size_before_compaction;
	s[3] = t->m_network ? t->m_network->n_params() : 0;
	REFF_END
}
int reff_save_snapshot(void* t, const char* path, int with_optimizer) { REFF_BEGIN ((Testbed*)t)->save_snapshot(path, with_optimizer != 0); REFF_END }
int reff_load_snapshot(void* t, const char* path) { REFF_BEGIN ((Testbed*)t)->load_snapshot(path); REFF_END }

int reff_set_option(void* tp, const char* name, double v) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	const std::string k = name;
	if (k == "seed") t->m_seed = (uint32_t)v;
	else if (k == "random_bg_color") t->m_nerf.training.random_bg_color = v != 0;
	else if (k == "linear_colors") t->m_nerf.training.linear_colors = v != 0;
	else if (k == "snap_to_pixel_centers_training") t->m_nerf.training.snap_to_pixel_centers = v != 0;
	else if (k == "snap_to_pixel_centers") t->m_snap_to_pixel_centers = v != 0;
	else if (k == "optimize_extrinsics") t->m_nerf.training.optimize_extrinsics = v != 0;
	else if (k == "optimize_exposure") t->m_nerf.training.optimize_exposure = v != 0;
	else if (k == "sample_focal_plane_proportional_to_error") t->m_nerf.training.sample_focal_plane_proportional_to_error = v != 0;
	else if (k == "sample_image_proportional_to_error") t->m_nerf.training.sample_image_proportional_to_error = v != 0;
	else if (k == "optimize_distortion") t->m_nerf.training.optimize_distortion = v != 0;
	else if (k == "optimize_focal_length") t->m_nerf.training.optimize_focal_length = v != 0;
	else if (k == "near_distance") t->m_nerf.training.near_distance = (float)v;
	else if (k == "render_min_transmittance") t->m_nerf.render_min_transmittance = (float)v;
	else if (k == "cone_angle_constant") t->m_nerf.cone_angle_constant = (float)v;
	else if (k == "exposure") t->m_exposure = (float)v;
	else if (k == "fov") t->set_fov((float)v);
	else if (k == "fov_axis") t->m_fov_axis = (uint32_t)v;
	else if (k == "color_space") t->m_color_space = (EColorSpace)(int)v;
	else if (k == "tonemap_curve") t->m_tonemap_curve = (ETonemapCurve)(int)v;
	else if (k == "render_mode") t->m_render_mode = (ERenderMode)(int)v;
	else if (k == "shall_train") t->m_train = v != 0;
	else if (k == "dynamic_res") t->m_dynamic_res = v != 0;
	else if (k == "background_color_r") t->m_background_color[0] = (float)v;
	else if (k == "background_color_g") t->m_background_color[1] = (float)v;
	else if (k == "background_color_b") t->m_background_color[2] = (float)v;
	else if (k == "background_color_a") t->m_background_color[3] = (float)v;
	else if (k == "n_images_for_training") t->m_nerf.training.n_images_for_training = (int)v;
	else if (k == "render_with_lens_distortion") t->m_nerf.render_with_lens_distortion = v != 0;
	else if (k == "sharpen") t->m_nerf.sharpen = (float)v;
	else throw std::runtime_error("reff_set_option: unknown option " + k);
	REFF_END
}

// Testbed::render_to_cpu (python_api.cu:132-190) for a static camera: spp x render_frame into a windowless surface, then the read-back.
int reff_render(void* tp, const float* camera12, int w, int h, int spp, int linear, float* out_rgba) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	CudaRenderBuffer surface{std::make_shared<CudaSurface2D>()};
	surface.resize({w, h});
	surface.reset_accumulation();
	if (camera12) t->set_nerf_camera_matrix(mat34(camera12));
	t->m_smoothed_camera = t->m_camera;
	for (int i = 0; i < spp; ++i) t->render_frame(t->m_smoothed_camera, t->m_smoothed_camera, Vector4f::Zero(), surface, !linear);
	CUDA_CHECK_THROW(cudaMemcpy2DFromArray(out_rgba, w * sizeof(float) * 4, surface.surface_provider().array(), 0, 0, w * sizeof(float) * 4, h, cudaMemcpyDeviceToHost));
	REFF_END
}
// the camera matrix Testbed::render would use for m_camera (ngp convention) -- to hand the SAME matrix to the product
int reff_get_camera(void* tp, float* camera12) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	for (int c = 0; c < 4; ++c) for (int k = 0; k < 3; ++k) camera12[c * 3 + k] = t->m_camera(k, c);
	REFF_END
}
int reff_set_camera_to_training_view(void* tp, int view) { REFF_BEGIN ((Testbed*)tp)->set_camera_to_training_view(view); REFF_END }

// ---- Blender path: RenderRequest -> bl_render_frame (python_api.cu:233-262) ----
struct reff_mask { int shape; int mode; float transform[16]; float feather; float opacity; float dims[3]; }; // shape 0 box (dims), 1 cylinder (radius, height), 2 sphere (radius)
struct reff_nerf { const char* snapshot_path; float aabb[6]; float transform[16]; float opacity; int n_masks; const reff_mask* masks; };
struct reff_request {
	int width, height, mip, flip_y, spp, color_space, tonemap_curve;
	float exposure, background_color[4];
	float camera[12]; int camera_model; float focal_length, near_distance, aperture_size, focus_z;
	float spherical_quadrilateral[3];   // width, height, curvature
	float quadrilateral_hexahedron[24]; // front tl,tr,bl,br then back tl,tr,bl,br
	float aabb[6];
	int n_masks; const reff_mask* masks;
	int n_nerfs; const reff_nerf* nerfs;
};

static Matrix4f mat44(const float* m) { Matrix4f r; for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r(k, c) = m[c * 4 + k]; return r; }
static Mask3D make_mask(const reff_mask& m) {
	const Matrix4f tf = mat44(m.transform);
	if (m.shape == 0) return Mask3D::Box(Vector3f(m.dims[0], m.dims[1], m.dims[2]), tf, (EMaskMode)m.mode, m.feather, m.opacity);
	if (m.shape == 1) return Mask3D::Cylinder(m.dims[0], m.dims[1], tf, (EMaskMode)m.mode, m.feather, m.opacity);
	if (m.shape == 2) return Mask3D::Sphere(m.dims[0], tf, (EMaskMode)m.mode, m.feather, m.opacity);
	return Mask3D::All((EMaskMode)m.mode);
}

int reff_bl_render(void* tp, const reff_request* rq, float* out_rgba) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	RenderOutputProperties output(Vector2i(rq->width, rq->height), DownsampleInfo::MakeFromMip(Vector2i(rq->width, rq->height), rq->mip), (uint32_t)rq->spp,
		(EColorSpace)rq->color_space, (ETonemapCurve)rq->tonemap_curve, rq->exposure,
		Vector4f(rq->background_color[0], rq->background_color[1], rq->background_color[2], rq->background_color[3]), rq->flip_y != 0);
	const float* q = rq->quadrilateral_hexahedron;
	auto v3 = [&](int i) { return Vector3f(q[i * 3], q[i * 3 + 1], q[i * 3 + 2]); };
	QuadrilateralHexahedron qh(Quadrilateral3D(v3(0), v3(1), v3(2), v3(3)), Quadrilateral3D(v3(4), v3(5), v3(6), v3(7)));
	SphericalQuadrilateral sq(rq->spherical_quadrilateral[0], rq->spherical_quadrilateral[1], rq->spherical_quadrilateral[2]);
	RenderCameraProperties camera(mat34(rq->camera), (ECameraModel)rq->camera_model, rq->focal_length, rq->near_distance, rq->aperture_size, rq->focus_z, sq, qh);
	RenderModifiersDescriptor mods;
	for (int i = 0; i < rq->n_masks; ++i) mods.masks.push_back(make_mask(rq->masks[i]));
	std::vector<NerfDescriptor> nerfs;
	for (int i = 0; i < rq->n_nerfs; ++i) {
		const reff_nerf& n = rq->nerfs[i];
		RenderModifiersDescriptor nm;
		for (int k = 0; k < n.n_masks; ++k) nm.masks.push_back(make_mask(n.masks[k]));
		nerfs.emplace_back(std::string(n.snapshot_path), BoundingBox(Vector3f(n.aabb[0], n.aabb[1], n.aabb[2]), Vector3f(n.aabb[3], n.aabb[4], n.aabb[5])), mat44(n.transform), nm, n.opacity);
	}
	RenderRequest request(output, camera, mods, nerfs, BoundingBox(Vector3f(rq->aabb[0], rq->aabb[1], rq->aabb[2]), Vector3f(rq->aabb[3], rq->aabb[4], rq->aabb[5])));
	CudaRenderBuffer render_buffer{std::make_shared<CudaSurface2D>()};
	render_buffer.resize({rq->width, rq->height});
	render_buffer.reset_accumulation();
	t->bl_render_frame(render_buffer, request);
	CUDA_CHECK_THROW(cudaMemcpy2DFromArray(out_rgba, rq->width * sizeof(float) * 4, render_buffer.surface_provider().array(), 0, 0, rq->width * sizeof(float) * 4, rq->height, cudaMemcpyDeviceToHost));
	REFF_END
}

// ---- state read-back for parity checks ----
int reff_get_params(void* tp, float* fp32, uint16_t* inference_half, uint32_t n) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	if (n != t->m_network->n_params()) throw std::runtime_error("reff_get_params: size mismatch");
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	if (fp32) CUDA_CHECK_THROW(cudaMemcpy(fp32, t->m_trainer->params_full_precision(), sizeof(float) * n, cudaMemcpyDeviceToHost));
	if (inference_half) CUDA_CHECK_THROW(cudaMemcpy(inference_half, t->m_trainer->params_inference(), 2 * n, cudaMemcpyDeviceToHost));
	REFF_END
}
int reff_get_density_grid(void* tp, float* grid, uint32_t n_cells, uint8_t* bitfield, uint32_t n_bytes) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	if (grid) { if (n_cells != t->m_nerf.density_grid.size()) throw std::runtime_error("reff_get_density_grid: size mismatch"); t->m_nerf.density_grid.copy_to_host(grid, n_cells); }
	if (bitfield) { if (n_bytes != t->m_nerf.density_grid_bitfield.size()) throw std::runtime_error("reff_get_density_grid: bitfield size mismatch"); t->m_nerf.density_grid_bitfield.copy_to_host(bitfield, n_bytes); }
	REFF_END
}
// camera extrinsics after optimisation (Testbed::Nerf::Training::get_camera_extrinsics)
int reff_get_camera_extrinsics(void* tp, int frame, float* out12) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	auto m = t->m_nerf.training.get_camera_extrinsics(frame);
	for (int c = 0; c < 4; ++c) for (int k = 0; k < 3; ++k) out12[c * 3 + k] = m(k, c);
	REFF_END
}

// learned per-image exposures (m_nerf.training.cam_exposure[i].variable(), testbed.h:632): out[n_images][3]
int reff_get_exposures(void* tp, float* out) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	for (size_t i = 0; i < t->m_nerf.training.cam_exposure.size(); ++i) {
		auto v = t->m_nerf.training.cam_exposure[i].variable();
		for (int c = 0; c < 3; ++c) out[i * 3 + c] = v[c];
	}
	REFF_END
}
// error-map state (Testbed::Nerf::Training::ErrorMap, testbed.h:600-615): state5 = {error-map res x, y, n_steps_between_error_map_updates, is_cdf_valid,
// n_steps_since_error_map_update}; pmf_img (may be null) receives the per-image sampling probabilities of the last CDF update (pmf_img_cpu).
int reff_error_map_state(void* tp, int* state5, float* pmf_img) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	auto& tr = t->m_nerf.training;
	state5[0] = tr.error_map.resolution.x(); state5[1] = tr.error_map.resolution.y(); state5[2] = (int)tr.n_steps_between_error_map_updates;
	state5[3] = tr.error_map.is_cdf_valid ? 1 : 0; state5[4] = (int)tr.n_steps_since_error_map_update;
	if (pmf_img) for (size_t i = 0; i < tr.error_map.pmf_img_cpu.size(); ++i) pmf_img[i] = tr.error_map.pmf_img_cpu[i];
	REFF_END
}

// ---- neural-image and SDF modes (reff_create with mode 2 / 1): what the parity tests need beyond load / train / render / snapshots ----
// the batch Testbed::train_image just trained on (m_image.training.positions / targets, src/testbed_image.cu:228-272)
int reff_image_training_batch(void* tp, uint32_t n, float* positions2, float* targets3) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	if (n > t->m_image.training.positions.size()) throw std::runtime_error("reff_image_training_batch: larger than the last batch");
	CUDA_CHECK_THROW(cudaMemcpy(positions2, t->m_image.training.positions.data(), sizeof(float) * 2 * n, cudaMemcpyDeviceToHost));
	CUDA_CHECK_THROW(cudaMemcpy(targets3, t->m_image.training.targets.data(), sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
	REFF_END
}
int reff_image_resolution(void* tp, int* wh) { REFF_BEGIN Testbed* t = (Testbed*)tp; wh[0] = t->m_image.resolution.x(); wh[1] = t->m_image.resolution.y(); REFF_END }
// the loaded image as the training kernels see it (RGBA float, linear)
int reff_image_data(void* tp, float* rgba) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	if (t->m_image.type != Testbed::EDataType::Float) throw std::runtime_error("reff_image_data: half-precision image");
	CUDA_CHECK_THROW(cudaMemcpy(rgba, t->m_image.data.data(), sizeof(float) * 4 * (size_t)t->m_image.resolution.prod(), cudaMemcpyDeviceToHost));
	REFF_END
}
int reff_image_mse(void* tp, int quantize_to_byte, float* out) { REFF_BEGIN *out = ((Testbed*)tp)->compute_image_mse(quantize_to_byte != 0); REFF_END }
// m_network->inference on host positions [n][dims_in] -> [n][dims_out] (n a multiple of 128)
int reff_inference(void* tp, const float* positions, uint32_t n, uint32_t dims_in, uint32_t dims_out, float* out) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	tcnn::GPUMemory<float> in(n * dims_in), o(n * dims_out);
	in.copy_from_host(positions);
	tcnn::GPUMatrix<float> im(in.data(), dims_in, n), om(o.data(), dims_out, n);
	t->m_network->inference(t->m_stream.get(), im, om);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	o.copy_to_host(out);
	REFF_END
}
// the last assignments of Testbed::override_sdf_training_data (src/python_api.cu:96-103) for pairs that already are in the unit cube
int reff_sdf_override_training_data(void* tp, const float* points3, const float* distances, uint32_t n) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	auto& tr = t->m_sdf.training;
	tr.positions.enlarge(n); tr.positions_shuffled.enlarge(n); tr.distances.enlarge(n); tr.distances_shuffled.enlarge(n);
	CUDA_CHECK_THROW(cudaMemcpy(tr.positions.data(), points3, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
	CUDA_CHECK_THROW(cudaMemcpy(tr.distances.data(), distances, sizeof(float) * n, cudaMemcpyHostToDevice));
	tr.size = n; tr.idx = 0; tr.max_size = n; tr.generate_sdf_data_online = false;
	REFF_END
}
// the batch Testbed::train_sdf just trained on (positions_shuffled / distances_shuffled, src/testbed_sdf.cu:1237-1243)
int reff_sdf_training_batch(void* tp, uint32_t n, float* positions3, float* distances) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	CUDA_CHECK_THROW(cudaMemcpy(positions3, t->m_sdf.training.positions_shuffled.data(), sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
	CUDA_CHECK_THROW(cudaMemcpy(distances, t->m_sdf.training.distances_shuffled.data(), sizeof(float) * n, cudaMemcpyDeviceToHost));
	REFF_END
}
// Testbed::load_mesh's bounding box and scale (src/testbed_sdf.cu:1026-1036): {raw_aabb.min[3], raw_aabb.max[3], mesh_scale}
int reff_sdf_mesh_info(void* tp, float* out7) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	for (int k = 0; k < 3; ++k) { out7[k] = t->m_raw_aabb.min[k]; out7[3 + k] = t->m_raw_aabb.max[k]; }
	out7[6] = t->m_sdf.mesh_scale;
	REFF_END
}
// (position, distance) pairs as the reference samples them from the loaded mesh (generate_training_samples_sdf, src/testbed_sdf.cu:1108-1178): the SDF workload
int reff_sdf_generate(void* tp, uint32_t n, float* positions3, float* distances) {
	REFF_BEGIN
	Testbed* t = (Testbed*)tp;
	tcnn::GPUMemory<Eigen::Vector3f> p(n); tcnn::GPUMemory<float> d(n);
	t->generate_training_samples_sdf(p.data(), d.data(), n, t->m_stream.get(), false);
	CUDA_CHECK_THROW(cudaDeviceSynchronize());
	CUDA_CHECK_THROW(cudaMemcpy(positions3, p.data(), sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
	d.copy_to_host(distances);
	REFF_END
}

} // extern "C"
