// K13 / K14: gradients of the loss with respect to the network INPUTS (sample position through the hash grid, view direction through the SH encoding),
// reduced per training camera into position / rotation gradients, and the host-side per-camera Adam that moves the camera extrinsics.
// Replaces (reference): kernel_grid with dy_dx + kernel_grid_backward_input (dependencies/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:351-392, :551-575),
// kernel_sh_backward (encodings/spherical_harmonics.h:154-390), compute_cam_gradient_train_nerf (src/testbed_nerf.cu:1600-1707) and the host loop of
// Testbed::train_nerf (:3056-3083) with AdamOptimizer / RotationAdamOptimizer (include/neural-graphics-primitives/adam_optimizer.h) and
// Testbed::Nerf::Training::update_transforms (:2597-2633). Distortion-map, focal-length and exposure optimisation are not built.
#include "nerf_device.cuh"
#include "../../include/ngpb.h"

#include <cmath>
#include <cstring>

namespace ngpb {

struct GridLevelsIG { float scale[NGPB_MAX_LEVELS]; uint32_t resolution[NGPB_MAX_LEVELS], offset[NGPB_MAX_LEVELS], size[NGPB_MAX_LEVELS]; uint32_t n_levels; };

__device__ __forceinline__ uint32_t grid_entry_index(uint32_t size, uint32_t res, uint32_t x, uint32_t y, uint32_t z) { // grid_index + prime_hash, grid.h:111-128,:164-186
	uint32_t index;
	if ((uint64_t)size < (uint64_t)res * res * res) index = x ^ (y * 2654435761u) ^ (z * 805459861u);
	else index = x + y * res + z * res * res;
	return index % size;
}

// One thread per (sample, level), the 16 level-threads of a sample adjacent (as in the forward kernel): the level's contribution to dL/dposition
// is d/dx of the trilinear blend (grid.h:351-392, linear interpolation: pos_derivative = 1) times dL/dy of the level's two features; the levels are then
// summed over the 16 lanes and lane 0 writes {dL/dpos, 0, dL/ddir}. dL/ddir comes from the 16 SH-input gradients of the rgb network (kernel_sh_backward).
__global__ void __launch_bounds__(256) nerf_input_gradient_kernel(const uint32_t n, const GridLevelsIG L, const __half2* __restrict__ grid, const float* __restrict__ coords,
                                                                  const __half2* __restrict__ dL_dencoded, const __half* __restrict__ dL_dsh, float* __restrict__ coords_gradient)
{
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t level = tid & 15u, i = tid >> 4;
	float g[3] = {0.f, 0.f, 0.f};
	const bool valid = i < n;
	if (valid && level < L.n_levels) {
		const float scale = L.scale[level];
		const uint32_t res = L.resolution[level], size = L.size[level];
		const __half2* __restrict__ gl = grid + L.offset[level];
		float pos[3]; uint32_t pg[3];
		#pragma unroll
		for (int d = 0; d < 3; ++d) { // pos_fract, tcnn common_device.h:434-445
			const float p = __fmaf_rn(coords[(size_t)i * COORD_FLOATS + d], scale, 0.5f);
			const float fl = floorf(p);
			pg[d] = (uint32_t)(int)fl; pos[d] = p - fl;
		}
		const float2 dl = __half22float2(dL_dencoded[(size_t)i * 16 + level]);
		#pragma unroll
		for (uint32_t grad_dim = 0; grad_dim < 3; ++grad_dim) {
			float acc0 = 0.f, acc1 = 0.f;
			#pragma unroll
			for (uint32_t idx = 0; idx < 4; ++idx) {
				float weight = scale;
				uint32_t pl[3];
				#pragma unroll
				for (uint32_t ng = 0; ng < 2; ++ng) {
					const uint32_t dim = ng >= grad_dim ? ng + 1 : ng;
					if ((idx & (1u << ng)) == 0) { weight *= 1 - pos[dim]; pl[dim] = pg[dim]; } else { weight *= pos[dim]; pl[dim] = pg[dim] + 1; }
				}
				pl[grad_dim] = pg[grad_dim];
				const float2 left = __half22float2(__ldg(gl + grid_entry_index(size, res, pl[0], pl[1], pl[2])));
				pl[grad_dim] = pg[grad_dim] + 1;
				const float2 right = __half22float2(__ldg(gl + grid_entry_index(size, res, pl[0], pl[1], pl[2])));
				acc0 += weight * (right.x - left.x);
				acc1 += weight * (right.y - left.y);
			}
			g[grad_dim] = dl.x * acc0 + dl.y * acc1;
		}
	}
	#pragma unroll
	for (int o = 8; o > 0; o >>= 1) {
		#pragma unroll
		for (int d = 0; d < 3; ++d) g[d] += __shfl_xor_sync(0xffffffffu, g[d], o);
	}
	if (!valid || level != 0) return;
	float* out = coords_gradient + (size_t)i * COORD_FLOATS;
	out[0] = g[0]; out[1] = g[1]; out[2] = g[2]; out[3] = 0.f;
	float dx = 0.f, dy = 0.f, dz = 0.f;
	if (dL_dsh) { // kernel_sh_backward, degree 4: the polynomials of sh4() differentiated term by term in x = 2 d - 1, times 2 for the [0,1] -> [-1,1] mapping
		const float* c = coords + (size_t)i * COORD_FLOATS;
		const float x = c[4] * 2.f - 1.f, y = c[5] * 2.f - 1.f, z = c[6] * 2.f - 1.f;
		const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
		float s[16];
		#pragma unroll
		for (int k = 0; k < 16; ++k) s[k] = __half2float(dL_dsh[(size_t)i * 16 + k]);
		const float c1 = 0.48860251190291987f, c4 = 1.0925484305920792f, c6 = 0.94617469575755997f, c8 = 0.54627421529603959f;
		const float c9 = 0.59004358992664352f, c10 = 2.8906114426405538f, c11 = 0.45704579946446572f, c12 = 0.3731763325901154f, c14 = 1.4453057213202769f;
		dy += s[1] * -c1; dz += s[2] * c1; dx += s[3] * -c1;
		dx += s[4] * (c4 * y); dy += s[4] * (c4 * x);
		dy += s[5] * (-c4 * z); dz += s[5] * (-c4 * y);
		dz += s[6] * (2.f * c6 * z);
		dx += s[7] * (-c4 * z); dz += s[7] * (-c4 * x);
		dx += s[8] * (2.f * c8 * x); dy += s[8] * (-2.f * c8 * y);
		dx += s[9] * (-6.f * c9 * xy); dy += s[9] * (3.f * c9 * (y2 - x2));
		dx += s[10] * (c10 * yz); dy += s[10] * (c10 * xz); dz += s[10] * (c10 * xy);
		dy += s[11] * (c11 * (1.f - 5.f * z2)); dz += s[11] * (-10.f * c11 * yz);
		dz += s[12] * (c12 * (15.f * z2 - 3.f));
		dx += s[13] * (c11 * (1.f - 5.f * z2)); dz += s[13] * (-10.f * c11 * xz);
		dx += s[14] * (2.f * c14 * xz); dy += s[14] * (-2.f * c14 * yz); dz += s[14] * (c14 * (x2 - y2));
		dx += s[15] * (3.f * c9 * (y2 - x2)); dy += s[15] * (6.f * c9 * xy);
		dx *= 2.0f; dy *= 2.0f; dz *= 2.0f;
	}
	out[4] = dx; out[5] = dy; out[6] = dz;
}

// compute_cam_gradient_train_nerf (:1600-1707), no distortion map / focal length: one thread per kept ray. cdf_img (may be null): the image CDF the batch's
// rays were drawn from (K19), so that a ray finds its image again.
// numsteps holds, per kept ray, {compacted sample count, compacted base} as left by the loss stage; coords / coords_gradient are the compacted batch.
__global__ void __launch_bounds__(128) cam_gradient_kernel(const uint32_t n_rays_global, const Aabb aabb, const uint32_t* __restrict__ rays_counter, const uint32_t n_images,
                                                           const uint32_t* __restrict__ ray_indices, const float* __restrict__ rays_unnormalized, const uint32_t* __restrict__ numsteps,
                                                           const float* __restrict__ coords, const float* __restrict__ coords_gradient,
                                                           float* __restrict__ cam_pos_gradient, float* __restrict__ cam_rot_gradient, const float* __restrict__ cdf_img)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *rays_counter) return;
	const uint32_t ns = numsteps[i * 2 + 0];
	if (ns == 0) return;
	const uint32_t base = numsteps[i * 2 + 1];
	const uint32_t img = image_idx(ray_indices[i], n_rays_global, n_images, cdf_img);
	const float* r = rays_unnormalized + (size_t)i * 6;
	const V3 ro = {r[0], r[1], r[2]};
	V3 rd = {r[3], r[4], r[5]};
	{ const float z = sum3(rd.x * rd.x, rd.y * rd.y, rd.z * rd.z); if (z > 0.f) { const float nrm = sqrtf(z); rd = {rd.x / nrm, rd.y / nrm, rd.z / nrm}; } }
	const float inv_diag[3] = {1.0f / (aabb.max[0] - aabb.min[0]), 1.0f / (aabb.max[1] - aabb.min[1]), 1.0f / (aabb.max[2] - aabb.min[2])}; // warp_position_derivative
	V3 go = {0.f, 0.f, 0.f}, gd = {0.f, 0.f, 0.f};
	for (uint32_t j = 0; j < ns; ++j) {
		const float* c = coords + (size_t)(base + j) * COORD_FLOATS;
		const float* gc = coords_gradient + (size_t)(base + j) * COORD_FLOATS;
		const V3 pgrad = {gc[0] * inv_diag[0], gc[1] * inv_diag[1], gc[2] * inv_diag[2]};
		go = {go.x + pgrad.x, go.y + pgrad.y, go.z + pgrad.z};
		const V3 pos = unwarp_position(c, aabb);
		const float dx = pos.x - ro.x, dy = pos.y - ro.y, dz = pos.z - ro.z;
		const float t = sqrtf(sum3(dx * dx, dy * dy, dz * dz)); // further-away samples move more as the direction changes
		gd = {gd.x + (pgrad.x * t + gc[4] * 0.5f), gd.y + (pgrad.y * t + gc[5] * 0.5f), gd.z + (pgrad.z * t + gc[6] * 0.5f)}; // warp_direction_derivative = 0.5
	}
	if (cam_pos_gradient) {
		atomicAdd(&cam_pos_gradient[img * 3 + 0], go.x); atomicAdd(&cam_pos_gradient[img * 3 + 1], go.y); atomicAdd(&cam_pos_gradient[img * 3 + 2], go.z);
	}
	if (cam_rot_gradient) { // rotations are averaged in log space: angle-axis = ray.d x ray_gradient.d
		atomicAdd(&cam_rot_gradient[img * 3 + 0], rd.y * gd.z - rd.z * gd.y);
		atomicAdd(&cam_rot_gradient[img * 3 + 1], rd.z * gd.x - rd.x * gd.z);
		atomicAdd(&cam_rot_gradient[img * 3 + 2], rd.x * gd.y - rd.y * gd.x);
	}
}

Aabb make_aabb(const float* a);

void nerf_input_gradient_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* coords, uint32_t n, const __half* dL_dencoded, const __half* dL_dsh,
                                float* coords_gradient) {
	if (n == 0) return;
	if (g->n_levels > 16 || g->n_pos_dims == 2) throw std::runtime_error("nerf_input_gradient: 3-D grids of up to 16 levels");
	GridLevelsIG L{};
	L.n_levels = g->n_levels;
	for (uint32_t l = 0; l < g->n_levels; ++l) { L.scale[l] = g->scale[l]; L.resolution[l] = g->resolution[l]; L.offset[l] = g->offsets[l]; L.size[l] = g->offsets[l + 1] - g->offsets[l]; }
	const uint64_t threads = (uint64_t)n * 16;
	if (threads > 0xFFFFFFFFull) throw std::runtime_error("nerf_input_gradient: n * 16 must fit 32 bits");
	nerf_input_gradient_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, stream>>>(n, L, (const __half2*)grid, coords, (const __half2*)dL_dencoded, dL_dsh, coords_gradient);
	NGPB_LAUNCH_CHECK();
}

void cam_gradient_launch(cudaStream_t stream, uint32_t max_rays, uint32_t n_rays_global, const float* aabb6, const uint32_t* rays_counter, uint32_t n_images, const uint32_t* ray_indices,
                         const float* rays, const uint32_t* numsteps, const float* coords, const float* coords_gradient, float* cam_pos_gradient, float* cam_rot_gradient,
                         const float* cdf_img) {
	if (max_rays == 0) return;
	cam_gradient_kernel<<<div_round_up(max_rays, 128), 128, 0, stream>>>(n_rays_global, make_aabb(aabb6), rays_counter, n_images, ray_indices, rays, numsteps, coords, coords_gradient,
		cam_pos_gradient, cam_rot_gradient, cdf_img);
	NGPB_LAUNCH_CHECK();
}

// ---- host: AdamOptimizer<Vector3f> / RotationAdamOptimizer (adam_optimizer.h:20-159) --------------------------------------------------------------------
static void rodrigues(float angle, const float* axis, float R[9]) { // Eigen::AngleAxisf(angle, axis).toRotationMatrix(), row-major
	const float c = std::cos(angle), s = std::sin(angle), t = 1.0f - c;
	const float x = axis[0], y = axis[1], z = axis[2];
	R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
	R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
	R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}
// Eigen reduces a fixed-size sum of three as a0 + (a1 + a2) (redux unroller: halves); norms, traces and matrix products below follow that order
static float sum3e(float a, float b, float c) { volatile float t = b + c; return a + t; }
static float norm3(const float* v) { return std::sqrt(sum3e(v[0] * v[0], v[1] * v[1], v[2] * v[2])); }
static void mat3_mul(const float* a, const float* b, float* out) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[r * 3 + c] = sum3e(a[r * 3] * b[c], a[r * 3 + 1] * b[3 + c], a[r * 3 + 2] * b[6 + c]); }
static void angle_axis_from_matrix(const float* R, float* out3) { // AngleAxis::fromRotationMatrix: through the quaternion (Eigen Quaternion.h / AngleAxis.h)
	float q[4]; // x y z w
	float t = sum3e(R[0], R[4], R[8]);
	if (t > 0.f) {
		t = std::sqrt(t + 1.0f); q[3] = 0.5f * t; t = 0.5f / t;
		q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
	} else {
		int i = 0;
		if (R[4] > R[0]) i = 1;
		if (R[8] > R[i * 4]) i = 2;
		const int j = (i + 1) % 3, k = (j + 1) % 3;
		t = std::sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0f);
		q[i] = 0.5f * t; t = 0.5f / t;
		q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t; q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t; q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
	}
	float nrm = norm3(q);
	if (nrm != 0.f) {
		const float angle = 2.0f * std::atan2(nrm, std::fabs(q[3]));
		if (q[3] < 0.f) nrm = -nrm;
		for (int c = 0; c < 3; ++c) out3[c] = q[c] / nrm * angle;
	} else {
		out3[0] = out3[1] = out3[2] = 0.f;
	}
}

} // namespace ngpb

using namespace ngpb;

// One Adam step of a camera's position offset (rotation = 0) or rotation offset in angle-axis form (rotation = 1), host only (adam_optimizer.h:37-44, :104-121).
// state: {iter, first_moment[3], second_moment[3], variable[3]} as 10 floats (iter stored as a float-valued count).
extern "C" void ngpb_camera_adam_step(float* state10, const float* gradient3, float learning_rate, int rotation) {
	const float eps = 1e-8f, beta1 = 0.9f, beta2 = 0.99f;
	state10[0] += 1.0f;
	const float iter = state10[0];
	float* m1 = state10 + 1; float* m2 = state10 + 4; float* var = state10 + 7;
	// (the rotation optimizer passes its integer iteration count to std::pow, which promotes the bias correction to double)
	const float lr = rotation ? (float)(learning_rate * std::sqrt(1 - std::pow((double)beta2, (double)iter)) / (1 - std::pow((double)beta1, (double)iter)))
	                          : learning_rate * std::sqrt(1 - std::pow(beta2, iter)) / (1 - std::pow(beta1, iter));
	float upd[3];
	for (int c = 0; c < 3; ++c) {
		m1[c] = beta1 * m1[c] + (1 - beta1) * gradient3[c];
		m2[c] = beta2 * m2[c] + (1 - beta2) * (gradient3[c] * gradient3[c]);
		upd[c] = lr * (m1[c] / (std::sqrt(m2[c]) + eps));
	}
	if (!rotation) { for (int c = 0; c < 3; ++c) var[c] -= upd[c]; return; }
	// the update is applied as a rotation composed with the current one: R(-|rot|, rot / |rot|) * R(|var|, var / |var|), back to angle-axis
	const float rot_len = norm3(upd), var_len = norm3(var);
	const float Z[3] = {0.f, 0.f, 1.f};
	float a1[3], a2[3];
	for (int c = 0; c < 3; ++c) { a1[c] = rot_len > 0 ? upd[c] / rot_len : Z[c]; a2[c] = var_len > 0 ? var[c] / var_len : Z[c]; }
	float R1[9], R2[9], M[9];
	rodrigues(-rot_len, a1, R1); rodrigues(var_len, a2, R2);
	mat3_mul(R1, R2, M);
	angle_axis_from_matrix(M, var);
}

// The exposure block of Testbed::train_nerf (src/testbed_nerf.cu:3105-3131), host only: one AdamOptimizer<Array3f> step per image on
// gradient * per_camera_loss_scale + l2_reg * exposure, then all exposures are re-centred on a zero mean (the renormalisation writes through to the
// optimizers' variables). states: [n_images][10] as ngpb_camera_adam_step; gradients: [n_images][3].
extern "C" void ngpb_exposure_update(uint32_t n_images, float* states, const float* gradients, float per_camera_loss_scale, float l2_reg, float learning_rate) {
	float mean[3] = {0.f, 0.f, 0.f};
	for (uint32_t i = 0; i < n_images; ++i) {
		float* st = states + (size_t)i * 10;
		float g[3];
		for (int c = 0; c < 3; ++c) g[c] = gradients[(size_t)i * 3 + c] * per_camera_loss_scale + st[7 + c] * l2_reg;
		ngpb_camera_adam_step(st, g, learning_rate, 0);
		for (int c = 0; c < 3; ++c) mean[c] += st[7 + c];
	}
	for (int c = 0; c < 3; ++c) mean[c] /= (float)n_images;
	for (uint32_t i = 0; i < n_images; ++i) for (int c = 0; c < 3; ++c) states[(size_t)i * 10 + 7 + c] -= mean[c];
}

// Training::update_transforms (:2597-2633): the camera's transform = dataset transform with the rotation offset applied on the left of its 3x3 block and the
// position offset added to its translation. xform12 in / out: 3x4 column-major.
extern "C" void ngpb_apply_camera_offsets(const float* xform12, const float* pos_offset3, const float* rot_offset3, float* out12) {
	std::memcpy(out12, xform12, 12 * sizeof(float));
	const float angle = norm3(rot_offset3);
	if (angle > 0) {
		const float axis[3] = {rot_offset3[0] / angle, rot_offset3[1] / angle, rot_offset3[2] / angle};
		float R[9];
		rodrigues(angle, axis, R);
		for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) out12[c * 3 + r] = sum3e(R[r * 3] * xform12[c * 3], R[r * 3 + 1] * xform12[c * 3 + 1], R[r * 3 + 2] * xform12[c * 3 + 2]);
	}
	for (int r = 0; r < 3; ++r) out12[9 + r] = xform12[9 + r] + pos_offset3[r];
}

extern "C" int ngpb_nerf_input_gradient(void* stream, const ngpb_grid* g, const ngpb_half* grid, const float* coords, uint32_t n, const ngpb_half* dL_dencoded, const ngpb_half* dL_dsh,
                                        float* coords_gradient) {
	try {
		if (!g || !grid || !coords || !dL_dencoded || !coords_gradient) { set_last_error("ngpb_nerf_input_gradient: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		nerf_input_gradient_launch((cudaStream_t)stream, g, (const __half*)grid, coords, n, (const __half*)dL_dencoded, (const __half*)dL_dsh, coords_gradient);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_compute_cam_gradient(void* stream, uint32_t max_rays, uint32_t n_rays_global, const float* aabb6, const uint32_t* rays_counter_dev, uint32_t n_images,
                                         const uint32_t* ray_indices, const float* rays, const uint32_t* numsteps, const float* coords, const float* coords_gradient,
                                         float* cam_pos_gradient, float* cam_rot_gradient) {
	try {
		if (!aabb6 || !rays_counter_dev || !ray_indices || !rays || !numsteps || !coords || !coords_gradient || n_images == 0) { set_last_error("ngpb_compute_cam_gradient: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		cam_gradient_launch((cudaStream_t)stream, max_rays, n_rays_global, aabb6, rays_counter_dev, n_images, ray_indices, rays, numsteps, coords, coords_gradient, cam_pos_gradient, cam_rot_gradient, nullptr);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
