// K17: classic single-NeRF render, ERenderMode::Shade, perspective camera.
// Replaces init_rays_with_payload_kernel_nerf (reference: src/testbed_nerf.cu:1809-1978) + pixel_to_ray
// (include/neural-graphics-primitives/common_device.cuh:260-317), advance_pos_nerf (:612-664),
// generate_next_nerf_network_inputs (:705-766), composite_kernel_nerf (:767-989), compact_kernel_nerf (:1784-1807),
// shade_kernel_nerf (:1748-1782), the NerfTracer host loop (:2047-2267), and CudaRenderBuffer::accumulate / tonemap
// (src/render_buffer.cu:235-266, :268-349, :540-567, :606-660).
//
// Shape. The reference compacts the live rays every 1-8 march steps and reads the live count back to the host each time
// (:2184-2193), i.e. hundreds of stream synchronisations per frame. Here a pass marches every live ray S steps, runs the
// network once over all of them, composites, and the composite kernel itself appends the surviving rays to the next pass's
// list (warp-aggregated atomics) and shades the finished ones; S grows as rays die (4 .. 32), so a frame takes a handful of
// passes. Per ray, the samples, their order and the termination test are exactly the reference's serial semantics; only the
// chunking differs, which does not change the result.
// Not built (outside SURVEY.md s8 for this path): lens distortion, depth of field, render masks, envmap, glow / debug modes.
#include "nerf_device.cuh"

#include <algorithm>
#include <cstring>

namespace ngpb {

// ---- low-discrepancy jitter: include/neural-graphics-primitives/random_val.cuh:159-322 ----
__device__ __constant__ uint32_t c_sobol_dir1[32] = {
	0x80000000, 0xc0000000, 0xa0000000, 0xf0000000, 0x88000000, 0xcc000000, 0xaa000000, 0xff000000,
	0x80800000, 0xc0c00000, 0xa0a00000, 0xf0f00000, 0x88880000, 0xcccc0000, 0xaaaa0000, 0xffff0000,
	0x80008000, 0xc000c000, 0xa000a000, 0xf000f000, 0x88008800, 0xcc00cc00, 0xaa00aa00, 0xff00ff00,
	0x80808080, 0xc0c0c0c0, 0xa0a0a0a0, 0xf0f0f0f0, 0x88888888, 0xcccccccc, 0xaaaaaaaa, 0xffffffff};

__host__ __device__ inline uint32_t sobol_dim(uint32_t index, uint32_t dim) {
	if (dim == 0) { // direction numbers of dimension 0 are the single bits in reverse order: the result is the bit reversal
		uint32_t x = index;
		x = (((x & 0xaaaaaaaa) >> 1) | ((x & 0x55555555) << 1));
		x = (((x & 0xcccccccc) >> 2) | ((x & 0x33333333) << 2));
		x = (((x & 0xf0f0f0f0) >> 4) | ((x & 0x0f0f0f0f) << 4));
		x = (((x & 0xff00ff00) >> 8) | ((x & 0x00ff00ff) << 8));
		return ((x >> 16) | (x << 16));
	}
	uint32_t X = 0;
#ifdef __CUDA_ARCH__
	for (uint32_t bit = 0; bit < 32; bit++) X ^= ((index >> bit) & 1) * c_sobol_dir1[bit];
#else
	// host: same table, generated (dimension 1 direction numbers are the rows of Pascal's triangle mod 2)
	uint32_t v = 0x80000000u;
	for (uint32_t bit = 0; bit < 32; bit++) { X ^= ((index >> bit) & 1) * v; v ^= v >> 1; }
#endif
	return X;
}
__host__ __device__ inline uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }
__host__ __device__ inline uint32_t reverse_bits(uint32_t x) {
	x = (((x & 0xaaaaaaaa) >> 1) | ((x & 0x55555555) << 1));
	x = (((x & 0xcccccccc) >> 2) | ((x & 0x33333333) << 2));
	x = (((x & 0xf0f0f0f0) >> 4) | ((x & 0x0f0f0f0f) << 4));
	x = (((x & 0xff00ff00) >> 8) | ((x & 0x00ff00ff) << 8));
	return ((x >> 16) | (x << 16));
}
__host__ __device__ inline uint32_t laine_karras_permutation(uint32_t x, uint32_t seed) {
	x += seed; x ^= x * 0x6c50b47cu; x ^= x * 0xb82f1e52u; x ^= x * 0xc7afe638u; x ^= x * 0x8d22f6e6u;
	return x;
}
__host__ __device__ inline uint32_t nested_uniform_scramble_base2(uint32_t x, uint32_t seed) { return reverse_bits(laine_karras_permutation(reverse_bits(x), seed)); }
__host__ __device__ inline float ld_random_val(uint32_t index, uint32_t seed, uint32_t dim = 0) {
	const float S = float(1.0 / (1ull << 32));
	index = nested_uniform_scramble_base2(index, seed);
	return (float)nested_uniform_scramble_base2(sobol_dim(index, dim), hash_combine(seed, dim)) * S;
}
// ld_random_pixel_offset (random_val.cuh:313-322), evaluated on the host once per sample
static void ld_random_pixel_offset(uint32_t spp, float* out) {
	for (uint32_t i = 0; i < 2; ++i) {
		const float S = float(1.0 / (1ull << 32));
		const uint32_t i0 = nested_uniform_scramble_base2(0u, 0xdeadbeefu), i1 = nested_uniform_scramble_base2(spp, 0xdeadbeefu);
		const float a = (float)nested_uniform_scramble_base2(sobol_dim(i0, i), hash_combine(0xdeadbeefu, i)) * S;
		const float b = (float)nested_uniform_scramble_base2(sobol_dim(i1, i), hash_combine(0xdeadbeefu, i)) * S;
		volatile float d = 0.5f - a; // (Constant(0.5) - a) + b, rounded as written
		volatile float e = d + b;
		out[i] = e - floorf(e);
	}
}

struct RenderRay { float o[3]; float t; float d[3]; uint32_t idx; };

struct RenderParams {
	int32_t width, height;
	float fx, fy, cx, cy;       // focal length in pixels; screen centre
	float cam[12];              // 3x4 camera-to-world, column-major
	float offset[2];            // sub-pixel offset of this sample
	uint32_t sample_index;
	Aabb render_aabb, train_aabb;
	float cone_angle, near_distance, min_transmittance;
	int32_t rgb_activation, density_activation, train_in_linear_colors;
};

// Marches ray (o, d) from t to the next sample position inside an occupied cell. false: the ray left the render box.
__device__ __forceinline__ bool march_to_occupied(const V3& o, const V3& d, const V3& idir, float cone_angle, const Aabb& box, const uint8_t* __restrict__ bitfield,
                                                  float& t, V3& pos, float& dt) {
	while (true) {
		pos = V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
		if (!aabb_contains(box, pos)) return false;
		dt = calc_dt(t, cone_angle);
		const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) return true;
		t = advance_to_next_voxel(t, cone_angle, pos, d, idir, NERF_GRIDSIZE >> mip);
	}
}

// Appends one record per set predicate to a list with ONE atomic per warp. Returns the slot (valid where pred).
__device__ __forceinline__ uint32_t warp_append(bool pred, uint32_t* __restrict__ counter) {
	const uint32_t mask = __ballot_sync(0xffffffffu, pred);
	const uint32_t lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (mask && lane == (uint32_t)__ffs(mask) - 1) base = atomicAdd(counter, (uint32_t)__popc(mask));
	base = __shfl_sync(0xffffffffu, base, mask ? __ffs(mask) - 1 : 0);
	return base + __popc(mask & ((1u << lane) - 1u));
}

// init_rays_with_payload_kernel_nerf + advance_pos_nerf: one thread per pixel; live rays are appended to `rays`.
__global__ void __launch_bounds__(128) render_init_kernel(const RenderParams P, const uint8_t* __restrict__ bitfield, RenderRay* __restrict__ rays, float4* __restrict__ rgba,
                                                          uint32_t* __restrict__ counter)
{
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_pixels = (uint32_t)P.width * (uint32_t)P.height;
	bool alive = idx < n_pixels;
	RenderRay r{};
	if (alive) {
		const uint32_t x = idx % (uint32_t)P.width, y = idx / (uint32_t)P.width;
		const float uvx = ((float)x + P.offset[0]) / (float)P.width, uvy = ((float)y + P.offset[1]) / (float)P.height;
		const float dcam[3] = {(uvx - P.cx) * (float)P.width / P.fx, (uvy - P.cy) * (float)P.height / P.fy, 1.0f};
		const float row0[3] = {P.cam[0], P.cam[3], P.cam[6]}, row1[3] = {P.cam[1], P.cam[4], P.cam[7]}, row2[3] = {P.cam[2], P.cam[5], P.cam[8]};
		V3 d = {dot3(row0, dcam), dot3(row1, dcam), dot3(row2, dcam)};
		V3 o = {P.cam[9] + d.x * P.near_distance, P.cam[10] + d.y * P.near_distance, P.cam[11] + d.z * P.near_distance};
		const float z = sum3(d.x * d.x, d.y * d.y, d.z * d.z);
		if (z > 0.f) { const float nrm = sqrtf(z); d = V3{d.x / nrm, d.y / nrm, d.z / nrm}; }
		float tmin, tmax;
		aabb_ray_intersect(P.render_aabb, o, d, &tmin, &tmax);
		float t = fmaxf(tmin, 0.0f) + 1e-6f;
		alive = aabb_contains(P.render_aabb, V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t});
		if (alive) {
			const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
			t += ld_random_val(P.sample_index, idx * 786433u) * calc_dt(t, P.cone_angle);
			V3 pos; float dt;
			alive = march_to_occupied(o, d, idir, P.cone_angle, P.render_aabb, bitfield, t, pos, dt);
			r = RenderRay{{o.x, o.y, o.z}, t, {d.x, d.y, d.z}, idx};
		}
	}
	const uint32_t slot = warp_append(alive, counter);
	if (alive) { rays[slot] = r; rgba[slot] = make_float4(0.f, 0.f, 0.f, 0.f); }
}

// generate_next_nerf_network_inputs: one thread per live ray, up to n_steps samples, ray-major slots [i * n_steps + j].
__global__ void __launch_bounds__(128) render_march_kernel(const RenderParams P, const uint32_t* __restrict__ n_rays_dev, const uint32_t n_steps, const uint8_t* __restrict__ bitfield,
                                                           RenderRay* __restrict__ rays, float* __restrict__ coords, uint32_t* __restrict__ ray_steps)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *n_rays_dev) return;
	RenderRay r = rays[i];
	const V3 o = {r.o[0], r.o[1], r.o[2]}, d = {r.d[0], r.d[1], r.d[2]};
	const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
	const V3 wd = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f};
	float t = r.t;
	float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
	uint32_t j = 0;
	for (; j < n_steps; ++j) {
		V3 pos; float dt;
		if (!march_to_occupied(o, d, idir, P.cone_angle, P.render_aabb, bitfield, t, pos, dt)) break;
		const V3 wp = warp_position(pos, P.train_aabb);
		c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
		c += COORD_FLOATS;
		t += dt;
	}
	ray_steps[i] = j;
	for (uint32_t k = j; k < n_steps; ++k) { // unused slots still go through the network: keep them finite
		c[0] = c[1] = c[2] = 0.5f; c[3] = 0.f; c[4] = c[5] = c[6] = 0.5f;
		c += COORD_FLOATS;
	}
	rays[i].t = t;
}

// composite_kernel_nerf (Shade) + compact_kernel_nerf + shade_kernel_nerf: composites this pass's samples; a ray that is still live is
// appended to the next pass's list, a finished one with alpha > 0.001 is shaded into the frame buffer.
__global__ void __launch_bounds__(128) render_composite_kernel(const RenderParams P, const uint32_t* __restrict__ n_rays_dev, const uint32_t n_steps,
                                                               const RenderRay* __restrict__ rays, const float4* __restrict__ rgba_in, const float* __restrict__ coords,
                                                               const __half* __restrict__ rgbsigma, const uint32_t* __restrict__ ray_steps,
                                                               RenderRay* __restrict__ rays_next, float4* __restrict__ rgba_next, uint32_t* __restrict__ next_counter,
                                                               float4* __restrict__ frame_buffer, unsigned long long* __restrict__ n_samples)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool active = i < *n_rays_dev;
	bool alive = false;
	float4 local = make_float4(0.f, 0.f, 0.f, 0.f);
	RenderRay r{};
	uint32_t actual = 0;
	if (active) {
		r = rays[i];
		local = rgba_in[i];
		actual = ray_steps[i];
		const float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
		const __half* out = rgbsigma + (size_t)i * n_steps * 4;
		uint32_t j = 0;
		for (; j < actual; ++j) {
			const uint2 raw = *reinterpret_cast<const uint2*>(out + j * 4);
			const __half2 h01 = *reinterpret_cast<const __half2*>(&raw.x), h23 = *reinterpret_cast<const __half2*>(&raw.y);
			const float T = 1.f - local.w;
			const float dt = unwarp_dt(c[j * COORD_FLOATS + 3]);
			const float alpha = 1.f - __expf(-network_to_density(__high2float(h23), P.density_activation) * dt);
			const float weight = alpha * T;
			local.x += network_to_rgb(__low2float(h01), P.rgb_activation) * weight;
			local.y += network_to_rgb(__high2float(h01), P.rgb_activation) * weight;
			local.z += network_to_rgb(__low2float(h23), P.rgb_activation) * weight;
			local.w += weight;
			if (local.w > (1.0f - P.min_transmittance)) {
				const float w = local.w;
				local.x /= w; local.y /= w; local.z /= w; local.w /= w;
				break;
			}
		}
		alive = !(j < n_steps); // terminated early, or left the box before the chunk was full (:979-982)
		if (!alive && local.w > 0.001f) {
			float4 tmp = local;
			if (!P.train_in_linear_colors) { tmp.x = srgb_to_linear(tmp.x); tmp.y = srgb_to_linear(tmp.y); tmp.z = srgb_to_linear(tmp.z); }
			const float4 f = frame_buffer[r.idx];
			const float k = 1.0f - tmp.w;
			frame_buffer[r.idx] = make_float4(tmp.x + f.x * k, tmp.y + f.y * k, tmp.z + f.z * k, tmp.w + f.w * k);
		}
	}
	const uint32_t slot = warp_append(alive, next_counter);
	if (alive) { rays_next[slot] = r; rgba_next[slot] = local; }
	// samples that went through the network for live rays
	uint32_t s = active ? actual : 0u;
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if ((threadIdx.x & 31) == 0 && s) atomicAdd(n_samples, (unsigned long long)s);
}

// accumulate_kernel (render_buffer.cu:235-266)
__global__ void __launch_bounds__(256) render_accumulate_kernel(const uint32_t n_pixels, const float4* __restrict__ frame_buffer, float4* __restrict__ accumulate_buffer,
                                                                const float sample_count, const int color_space)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	float4 color = frame_buffer[i];
	float4 tmp = accumulate_buffer[i];
	if (color_space == NGPB_COLOR_SRGB) { color.x = linear_to_srgb(color.x); color.y = linear_to_srgb(color.y); color.z = linear_to_srgb(color.z); }
	tmp.x = (tmp.x * sample_count + color.x) / (sample_count + 1);
	tmp.y = (tmp.y * sample_count + color.y) / (sample_count + 1);
	tmp.z = (tmp.z * sample_count + color.z) / (sample_count + 1);
	tmp.w = (tmp.w * sample_count + color.w) / (sample_count + 1);
	accumulate_buffer[i] = tmp;
}

// tonemap_kernel (render_buffer.cu:540-567) with ETonemapCurve::Identity, writing linear memory instead of a CUDA surface
__global__ void __launch_bounds__(256) render_tonemap_kernel(const uint32_t n_pixels, const float exposure_scale, float4 background_color, const float4* __restrict__ accumulate_buffer,
                                                             const int color_space, const int output_srgb, float4* __restrict__ out)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	if (color_space != NGPB_COLOR_SRGB) {
		background_color.x = srgb_to_linear(background_color.x); background_color.y = srgb_to_linear(background_color.y); background_color.z = srgb_to_linear(background_color.z);
	}
	float4 color = accumulate_buffer[i];
	const float weight = (1 - color.w) * background_color.w;
	color.x += background_color.x * weight; color.y += background_color.y * weight; color.z += background_color.z * weight;
	color.w += weight;
	float c[3] = {color.x, color.y, color.z};
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		float v = c[k];
		if (color_space == NGPB_COLOR_SRGB) v = srgb_to_linear(v);
		v *= exposure_scale;
		if (output_srgb) v = linear_to_srgb(v);
		c[k] = v;
	}
	out[i] = make_float4(c[0], c[1], c[2], color.w);
}

Aabb make_aabb(const float* a);
void hash_encode_forward_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* positions, uint32_t pos_stride, uint32_t n, const uint32_t* n_dev, __half* encoded);
void nerf_mlp_forward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma);

} // namespace ngpb

using namespace ngpb;

namespace {
struct RenderWorkspace {
	float4 *frame, *accum, *out, *rgba[2];
	RenderRay* rays[2];
	uint32_t *ray_steps, *counters;
	float* coords;
	__half *encoded, *rgbsigma;
	size_t bytes;
};
// Carves (or, with base == nullptr, only sizes) the render workspace for a frame of n_pixels.
RenderWorkspace render_workspace(uint32_t n_pixels, void* base) {
	const size_t slots = (size_t)next_multiple(n_pixels, 128) * NGPB_RENDER_FIRST_PASS_STEPS + 128;
	uint8_t* p = reinterpret_cast<uint8_t*>(base);
	size_t off = 0;
	auto take = [&](size_t bytes) { uint8_t* q = p ? p + off : nullptr; off += (bytes + 255) / 256 * 256; return q; };
	RenderWorkspace w{};
	w.frame = (float4*)take((size_t)n_pixels * 16);
	w.accum = (float4*)take((size_t)n_pixels * 16);
	w.out = (float4*)take((size_t)n_pixels * 16);
	for (int k = 0; k < 2; ++k) { w.rays[k] = (RenderRay*)take((size_t)n_pixels * sizeof(RenderRay)); w.rgba[k] = (float4*)take((size_t)n_pixels * 16); }
	w.ray_steps = (uint32_t*)take((size_t)n_pixels * 4);
	w.coords = (float*)take(slots * COORD_FLOATS * 4);
	w.encoded = (__half*)take(slots * N_ENC * 2);
	w.rgbsigma = (__half*)take(slots * 4 * 2);
	w.counters = (uint32_t*)take(64); // [0],[1]: live-ray counts of the two lists; [2..3]: 64-bit sample counter
	w.bytes = off;
	return w;
}
} // namespace

extern "C" uint64_t ngpb_render_workspace_bytes(uint32_t n_pixels) { return render_workspace(n_pixels, nullptr).bytes; }

extern "C" int ngpb_render_nerf(void* stream_, const ngpb_render_config* cfg, const ngpb_grid* g, const ngpb_half* params, const uint8_t* bitfield,
                                void* workspace, float* out_rgba_host, uint64_t* n_samples_out, uint32_t* n_launches_out) {
	try {
		if (!cfg || !g || !params || !bitfield || !workspace || !out_rgba_host || cfg->width <= 0 || cfg->height <= 0 || cfg->spp <= 0) {
			set_last_error("ngpb_render_nerf: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		const uint32_t n_pixels = (uint32_t)cfg->width * (uint32_t)cfg->height;
		const RenderWorkspace ws = render_workspace(n_pixels, workspace);
		float4 *frame = ws.frame, *accum = ws.accum, *out = ws.out;
		RenderRay* const* rays = ws.rays;
		float4* const* rgba = ws.rgba;
		uint32_t *ray_steps = ws.ray_steps, *counters = ws.counters;
		float* coords = ws.coords;
		__half *encoded = ws.encoded, *rgbsigma = ws.rgbsigma;
		uint32_t* host_counter = nullptr;
		NGPB_CUDA_CHECK(cudaMallocHost(&host_counter, 16));
		uint32_t launches = 0;

		RenderParams P{};
		P.width = cfg->width; P.height = cfg->height; P.fx = cfg->fx; P.fy = cfg->fy; P.cx = cfg->screen_center[0]; P.cy = cfg->screen_center[1];
		for (int k = 0; k < 12; ++k) P.cam[k] = cfg->camera[k];
		P.render_aabb = make_aabb(cfg->render_aabb); P.train_aabb = make_aabb(cfg->aabb);
		P.cone_angle = cfg->cone_angle_constant; P.near_distance = cfg->near_distance; P.min_transmittance = cfg->min_transmittance;
		P.rgb_activation = cfg->rgb_activation; P.density_activation = cfg->density_activation; P.train_in_linear_colors = cfg->train_in_linear_colors;

		NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64, stream));
		for (int s = 0; s < cfg->spp; ++s) {
			P.sample_index = (uint32_t)s;
			ld_random_pixel_offset(cfg->snap_to_pixel_centers ? 0u : (uint32_t)s, P.offset);
			NGPB_CUDA_CHECK(cudaMemsetAsync(frame, 0, (size_t)n_pixels * 16, stream)); // CudaRenderBuffer::clear_frame
			NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 8, stream));
			render_init_kernel<<<div_round_up(n_pixels, 128), 128, 0, stream>>>(P, bitfield, rays[0], rgba[0], counters + 0);
			NGPB_LAUNCH_CHECK(); ++launches;
			uint32_t cur = 0, n_alive = 0, n_alive0 = 0;
			for (uint32_t pass = 0;; ++pass) {
				NGPB_CUDA_CHECK(cudaMemcpyAsync(host_counter, counters + cur, 4, cudaMemcpyDeviceToHost, stream));
				NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
				n_alive = host_counter[0];
				if (n_alive == 0) break;
				if (pass == 0) n_alive0 = n_alive;
				// steps this pass: grows as rays die, bounded so that the slot count never exceeds the first pass's
				uint32_t n_steps = std::min<uint64_t>(NGPB_RENDER_MAX_PASS_STEPS, std::max<uint64_t>(NGPB_RENDER_FIRST_PASS_STEPS, (uint64_t)n_alive0 * NGPB_RENDER_FIRST_PASS_STEPS / n_alive));
				const uint32_t n_slots = next_multiple(n_alive * n_steps, 128);
				const uint32_t blocks = div_round_up(n_alive, 128);
				NGPB_CUDA_CHECK(cudaMemsetAsync(counters + (cur ^ 1), 0, 4, stream));
				render_march_kernel<<<blocks, 128, 0, stream>>>(P, counters + cur, n_steps, bitfield, rays[cur], coords, ray_steps);
				NGPB_LAUNCH_CHECK();
				hash_encode_forward_launch(stream, g, (const __half*)params + MLP_PARAMS, coords, COORD_FLOATS, n_slots, nullptr, encoded);
				nerf_mlp_forward_launch(stream, (const __half*)params, encoded, coords, n_slots, nullptr, rgbsigma);
				render_composite_kernel<<<blocks, 128, 0, stream>>>(P, counters + cur, n_steps, rays[cur], rgba[cur], coords, rgbsigma, ray_steps,
					rays[cur ^ 1], rgba[cur ^ 1], counters + (cur ^ 1), frame, reinterpret_cast<unsigned long long*>(counters + 2));
				NGPB_LAUNCH_CHECK();
				launches += 4;
				cur ^= 1;
			}
			if (s == 0) NGPB_CUDA_CHECK(cudaMemsetAsync(accum, 0, (size_t)n_pixels * 16, stream));
			render_accumulate_kernel<<<div_round_up(n_pixels, 256), 256, 0, stream>>>(n_pixels, frame, accum, (float)s, cfg->color_space);
			NGPB_LAUNCH_CHECK(); ++launches;
		}
		render_tonemap_kernel<<<div_round_up(n_pixels, 256), 256, 0, stream>>>(n_pixels, powf(2.0f, cfg->exposure),
			make_float4(cfg->background_color[0], cfg->background_color[1], cfg->background_color[2], cfg->background_color[3]), accum, cfg->color_space, cfg->output_srgb, out);
		NGPB_LAUNCH_CHECK(); ++launches;
		NGPB_CUDA_CHECK(cudaMemcpyAsync(out_rgba_host, out, (size_t)n_pixels * 16, cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaMemcpyAsync(host_counter, counters + 2, 8, cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		if (n_samples_out) std::memcpy(n_samples_out, host_counter, 8);
		if (n_launches_out) *n_launches_out = launches;
		cudaFreeHost(host_counter);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
