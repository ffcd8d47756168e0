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
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <thread>
#include <vector>

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
// (hash_combine, reverse_bits, laine_karras_permutation, nested_uniform_scramble_base2: nerf_device.cuh, shared with the error-map image sampling)
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

// ---- empty-space skipping over 8 x 8 x 8 blocks of occupancy cells -----------------------------------------------------------
// The samples of a ray are the members of its t chain (t += dt) that fall into occupied cells; how the march gets across EMPTY cells does not
// change them, as long as it stays on the chain. The reference tests one 1/128 cell per hop (advance_to_next_voxel, testbed_nerf.cu:449-463), which
// makes a background ray cost ~128 occupancy tests and leaves most of a render's march time in empty space. Here a second bitfield holds one bit per
// block of 8^3 cells (Morton order makes a block 64 consecutive bytes of the fine bitfield); an empty block is crossed in one hop, by the same
// `t += dt` additions. Blocks are only used for cascades 0..3, where they never straddle a cascade boundary, and a hop never runs past the t at which
// the step size selects the next cascade (the coarser cascade's cell may be occupied on its own, cf. bitfield_max_pool's `|=`).
// A half-trained occupancy grid is fluffy (at step 530 of the benchmark scene 56 % of the 8^3 blocks hold at least one occupied cell, but only 13 % of the
// 4^3 blocks), so a second level of 4 x 4 x 4-cell blocks (8 consecutive bytes of the bitfield) sits between the two.
constexpr uint32_t COARSE_BLOCKS_PER_CASCADE = NERF_GRID_CELLS / 512; // 4096
constexpr uint32_t COARSE8_WORDS = NERF_CASCADES * COARSE_BLOCKS_PER_CASCADE / 32;
constexpr uint32_t MID_BLOCKS_PER_CASCADE = NERF_GRID_CELLS / 64;      // 32768
constexpr uint32_t COARSE_WORDS = COARSE8_WORDS + NERF_CASCADES * MID_BLOCKS_PER_CASCADE / 32; // [8^3 level | 4^3 level]

__global__ void __launch_bounds__(256) coarse_occupancy_kernel(const uint8_t* __restrict__ bitfield, uint32_t* __restrict__ coarse)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; // cascade * 4096 + Morton index of the block
	const uint4* p = reinterpret_cast<const uint4*>(bitfield + (size_t)b * 64);
	uint32_t any = 0;
	#pragma unroll
	for (int k = 0; k < 4; ++k) { const uint4 v = __ldg(p + k); any |= v.x | v.y | v.z | v.w; }
	const uint32_t word = __ballot_sync(0xffffffffu, any != 0);
	if ((threadIdx.x & 31) == 0) coarse[b >> 5] = word;
}
__global__ void __launch_bounds__(256) mid_occupancy_kernel(const uint8_t* __restrict__ bitfield, uint32_t* __restrict__ mid)
{
	const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; // cascade * 32768 + Morton index of the 4^3 block
	const uint2 v = __ldg(reinterpret_cast<const uint2*>(bitfield + (size_t)b * 8));
	const uint32_t word = __ballot_sync(0xffffffffu, (v.x | v.y) != 0);
	if ((threadIdx.x & 31) == 0) mid[b >> 5] = word;
}
// Page-locked staging for the per-pass live-ray count and the finished frame, kept per host thread and grown on demand: a device-to-host copy into
// the caller's pageable numpy buffer runs at ~5 GB/s (2.1 ms for an 800 x 800 float4 frame), into pinned memory at PCIe speed.
struct PinnedScratch {
	uint8_t* p = nullptr; size_t bytes = 0;
	uint8_t* get(size_t n) {
		if (n > bytes) { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; NGPB_CUDA_CHECK(cudaMallocHost(&p, n)); bytes = n; }
		return p;
	}
};
static thread_local PinnedScratch g_pinned;
// The finished frame's way to the caller: straight into the caller's buffer when that is page-locked memory (pyngp hands out such buffers), otherwise through the
// page-locked staging buffer and a host copy split over a few threads (a single-threaded copy into freshly allocated pageable memory runs at ~10 GB/s, longer
// than the PCIe transfer of an 800 x 800 frame).
static bool host_pointer_is_pinned(const void* p) {
	cudaPointerAttributes attr{};
	if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return attr.type == cudaMemoryTypeHost;
}
static void parallel_host_copy(void* dst, const void* src, size_t bytes) {
	constexpr size_t MIN_PER_THREAD = 1u << 20;
	const unsigned n_threads = (unsigned)std::min<size_t>(4, bytes / MIN_PER_THREAD);
	if (n_threads <= 1) { std::memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	const size_t chunk = (bytes / n_threads + 4095) & ~(size_t)4095;
	for (unsigned k = 1; k < n_threads; ++k) {
		const size_t off = k * chunk;
		if (off >= bytes) break;
		th.emplace_back([=] { std::memcpy((uint8_t*)dst + off, (const uint8_t*)src + off, std::min(chunk, bytes - off)); });
	}
	std::memcpy(dst, src, std::min(chunk, bytes));
	for (auto& t : th) t.join();
}
// Device workspace of the Blender renderer, kept per host thread and device and grown on demand (cudaMalloc + cudaFree per frame cost more than the
// frame's first waves).
struct DeviceScratch {
	uint8_t* p = nullptr; size_t bytes = 0; int device = -1;
	uint8_t* get(size_t n) {
		int dev = 0;
		NGPB_CUDA_CHECK(cudaGetDevice(&dev));
		if (n > bytes || dev != device) {
			if (p) cudaFree(p);
			p = nullptr; bytes = 0;
			NGPB_CUDA_CHECK(cudaMalloc(&p, n));
			NGPB_CUDA_CHECK(cudaMemset(p, 0, n)); // the network passes run over whole 128-sample tiles: the slots past a wave's last sample are read (and ignored), never left undefined
			NGPB_CUDA_CHECK(cudaDeviceSynchronize());
			bytes = n; device = dev;
		}
		return p;
	}
};
static thread_local DeviceScratch g_blender_ws, g_blender_masks;

static uint32_t render_max_hops() { static const uint32_t h = [] { const char* e = std::getenv("NGPB_RENDER_HOPS"); return e ? (uint32_t)std::atoi(e) : 24u; }(); return h; }
static bool empty_space_blocks() { static const bool on = [] { const char* e = std::getenv("NGPB_RENDER_BLOCK_SKIP"); return !e || std::atoi(e) != 0; }(); return on; }
static void coarse_occupancy_launch(cudaStream_t stream, const uint8_t* bitfield, uint32_t* coarse) {
	coarse_occupancy_kernel<<<NERF_CASCADES * COARSE_BLOCKS_PER_CASCADE / 256, 256, 0, stream>>>(bitfield, coarse);
	NGPB_LAUNCH_CHECK();
	mid_occupancy_kernel<<<NERF_CASCADES * MID_BLOCKS_PER_CASCADE / 256, 256, 0, stream>>>(bitfield, coarse + COARSE8_WORDS);
	NGPB_LAUNCH_CHECK();
}

// One hop of the march from an unoccupied position: the t of the first chain member at or past the exit of the empty cell (or empty block).
__device__ __forceinline__ float hop_over_empty(float t, float dt, float cone_angle, const V3& pos, const V3& d, const V3& idir, uint32_t mip, uint32_t cell_idx,
                                                const uint32_t* __restrict__ coarse) {
	uint32_t res = NERF_GRIDSIZE >> mip;
	float t_cap = 3.4e38f;
	if (coarse && mip <= 3) {
		const uint32_t b8 = mip * COARSE_BLOCKS_PER_CASCADE + (cell_idx >> 9), b4 = mip * MID_BLOCKS_PER_CASCADE + (cell_idx >> 6);
		const bool empty8 = !((coarse[b8 >> 5] >> (b8 & 31)) & 1u);
		const bool empty4 = empty8 || !((coarse[COARSE8_WORDS + (b4 >> 5)] >> (b4 & 31)) & 1u);
		if (empty4) {
			res >>= empty8 ? 3 : 2;
			if (cone_angle > 0.f) { // dt * 128 reaches the next power of two at t_cap: from there on mip_from_dt selects the next cascade
				const float v = dt * (float)NERF_GRIDSIZE;
				const float next_pow2 = __uint_as_float((__float_as_uint(v) & 0x7F800000u) + 0x00800000u);
				t_cap = next_pow2 / ((float)NERF_GRIDSIZE * cone_angle);
			}
		}
	}
	const float t_target = fminf(t + distance_to_next_voxel(pos, d, idir, res), t_cap);
	if (cone_angle == 0.f) return chain_advance_to(t, t_target); // constant step: the same chain member, without the additions in between
	do { t += calc_dt(t, cone_angle); } while (t < t_target);
	return t;
}

// Marches ray (o, d) from t to the next sample position inside an occupied cell. MARCH_LEFT_BOX: the ray left the render box; MARCH_OUT_OF_HOPS:
// `hops` empty cells were crossed without finding one (t stays on the chain; the caller resumes from it in the next pass).
enum MarchResult { MARCH_LEFT_BOX = 0, MARCH_SAMPLE = 1, MARCH_OUT_OF_HOPS = 2 };
__device__ __forceinline__ MarchResult march_to_occupied(const V3& o, const V3& d, const V3& idir, float cone_angle, const Aabb& box, const uint8_t* __restrict__ bitfield,
                                                         const uint32_t* __restrict__ coarse, float& t, V3& pos, float& dt, uint32_t& hops) {
	while (true) {
		pos = V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
		if (!aabb_contains(box, pos)) return MARCH_LEFT_BOX;
		dt = calc_dt(t, cone_angle);
		const uint32_t mip = (uint32_t)mip_from_dt(dt, pos);
		const uint32_t idx = cascaded_grid_idx_at(pos, mip);
		if (bitfield[idx / 8 + grid_mip_offset(mip) / 8] & (1 << (idx % 8))) return MARCH_SAMPLE;
		if (hops == 0) return MARCH_OUT_OF_HOPS;
		--hops;
		t = hop_over_empty(t, dt, cone_angle, pos, d, idir, mip, idx, coarse);
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

// live rays x steps of the pass, for the network kernels' device-side sample count
__global__ void render_slot_count_kernel(const uint32_t* __restrict__ n_rays_dev, const uint32_t n_steps, uint32_t* __restrict__ n_slots_dev) { *n_slots_dev = *n_rays_dev * n_steps; }

// init_rays_with_payload_kernel_nerf + advance_pos_nerf: one thread per pixel; live rays are appended to `rays`.
__global__ void __launch_bounds__(128) render_init_kernel(const RenderParams P, const uint8_t* __restrict__ bitfield, const uint32_t* __restrict__ coarse,
                                                          RenderRay* __restrict__ rays, float4* __restrict__ rgba, uint32_t* __restrict__ counter)
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
			uint32_t hops = 0xFFFFFFFFu;
			alive = march_to_occupied(o, d, idir, P.cone_angle, P.render_aabb, bitfield, coarse, t, pos, dt, hops) == MARCH_SAMPLE;
			r = RenderRay{{o.x, o.y, o.z}, t, {d.x, d.y, d.z}, idx};
		}
	}
	const uint32_t slot = warp_append(alive, counter);
	if (alive) { rays[slot] = r; rgba[slot] = make_float4(0.f, 0.f, 0.f, 0.f); }
}

// generate_next_nerf_network_inputs: one thread per live ray, up to n_steps samples, ray-major slots [i * n_steps + j].
// A ray crosses at most `max_hops` empty cells per pass: a warp runs as long as its slowest lane, and a lane that has to cross the whole volume
// (~128 cells x ~170 instructions) next to 31 lanes that find their samples in adjacent cells left 3.5 of 32 lanes active on average (ncu).
// A ray that runs out of hops keeps the samples it found, stays alive and resumes in the next pass (RAY_STEPS_RESUME).
constexpr uint32_t RAY_STEPS_RESUME = 0x80000000u;
__global__ void __launch_bounds__(128) render_march_kernel(const RenderParams P, const uint32_t* __restrict__ n_rays_dev, const uint32_t n_steps, const uint32_t max_hops,
                                                           const uint8_t* __restrict__ bitfield, const uint32_t* __restrict__ coarse, RenderRay* __restrict__ rays, float* __restrict__ coords, uint32_t* __restrict__ ray_steps)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *n_rays_dev) return;
	RenderRay r = rays[i];
	const V3 o = {r.o[0], r.o[1], r.o[2]}, d = {r.d[0], r.d[1], r.d[2]};
	const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
	const V3 wd = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f};
	float t = r.t;
	float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
	uint32_t j = 0, hops = max_hops, resume = 0;
	for (; j < n_steps; ++j) {
		V3 pos; float dt;
		const MarchResult m = march_to_occupied(o, d, idir, P.cone_angle, P.render_aabb, bitfield, coarse, t, pos, dt, hops);
		if (m == MARCH_OUT_OF_HOPS) { resume = RAY_STEPS_RESUME; break; }
		if (m == MARCH_LEFT_BOX) break;
		const V3 wp = warp_position(pos, P.train_aabb);
		c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
		c += COORD_FLOATS;
		t += dt;
	}
	ray_steps[i] = j | resume;
	for (uint32_t k = j; k < n_steps; ++k) { // unused slots still go through the network: keep them finite
		c[0] = c[1] = c[2] = 0.5f; c[3] = 0.f; c[4] = c[5] = c[6] = 0.5f;
		c += COORD_FLOATS;
	}
	rays[i].t = t;
}

// composite_kernel_nerf (Shade) + compact_kernel_nerf + shade_kernel_nerf: composites this pass's samples; a ray that is still live is
// The same pass with a WARP per ray: the 32 lanes test 32 consecutive members of the ray's t chain at once, the occupied ones (before the first one
// outside the box) are the samples, in order. A ray's samples are exactly the chain members that fall into occupied cells -- the cell-to-cell hops of the
// serial march only decide how empty space is crossed -- so this finds the same samples; positions agree to float rounding (lane l computes
// t + l * dt where the serial march adds dt l times). The serial march is a chain of dependent occupancy loads with 3.5 of 32 lanes active on average;
// here every load of a round is independent and rays inside the object finish a pass in one round. A ray scans at most `max_rounds` x 32 members per
// pass (RAY_STEPS_RESUME otherwise).
__global__ void __launch_bounds__(256) render_march_warp_kernel(const RenderParams P, const uint32_t* __restrict__ n_rays_dev, const uint32_t n_steps, const uint32_t max_rounds,
                                                                const uint8_t* __restrict__ bitfield, RenderRay* __restrict__ rays, float* __restrict__ coords,
                                                                uint32_t* __restrict__ ray_steps)
{
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (i >= *n_rays_dev) return;
	const RenderRay r = rays[i];
	const V3 o = {r.o[0], r.o[1], r.o[2]}, d = {r.d[0], r.d[1], r.d[2]};
	const V3 wd = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f};
	float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
	float t0 = r.t;
	uint32_t j = 0, resume = RAY_STEPS_RESUME;
	for (uint32_t round = 0; round < max_rounds; ++round) {
		float t = t0;
		if (P.cone_angle == 0.f) t = t0 + (float)lane * MIN_CONE_STEPSIZE;
		else for (uint32_t k = 0; k < lane; ++k) t += calc_dt(t, P.cone_angle);
		const float dt = calc_dt(t, P.cone_angle);
		const V3 pos = V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
		const bool inside = aabb_contains(P.render_aabb, pos);
		bool occ = false;
		if (inside) occ = density_grid_occupied_at(pos, bitfield, (uint32_t)mip_from_dt(dt, pos));
		const uint32_t out_mask = __ballot_sync(0xffffffffu, !inside);
		const uint32_t before_exit = out_mask ? ((1u << (__ffs(out_mask) - 1)) - 1u) : 0xffffffffu; // lanes before the first member outside the box
		const uint32_t occ_mask = __ballot_sync(0xffffffffu, occ) & before_exit;
		const uint32_t need = n_steps - j;
		const uint32_t rank = __popc(occ_mask & ((1u << lane) - 1u));
		if (((occ_mask >> lane) & 1u) && rank < need) {
			const V3 wp = warp_position(pos, P.train_aabb);
			float* s = c + (size_t)(j + rank) * COORD_FLOATS;
			s[0] = wp.x; s[1] = wp.y; s[2] = wp.z; s[3] = warp_dt(dt); s[4] = wd.x; s[5] = wd.y; s[6] = wd.z;
		}
		const uint32_t found = __popc(occ_mask);
		if (found >= need) { // the pass is full: resume after the last sample taken
			const uint32_t last = __fns(occ_mask, 0, (int)need);
			t0 = __shfl_sync(0xffffffffu, t + dt, last);
			j = n_steps;
			resume = 0;
			break;
		}
		j += found;
		if (out_mask) { resume = 0; break; } // the ray left the box
		t0 = __shfl_sync(0xffffffffu, t + dt, 31);
	}
	if (lane == 0) { ray_steps[i] = j | resume; rays[i].t = t0; }
	for (uint32_t k = j * COORD_FLOATS + lane; k < n_steps * COORD_FLOATS; k += 32) { // unused slots still go through the network: keep them finite
		const uint32_t f = k % COORD_FLOATS;
		c[k] = f == 3 ? 0.f : 0.5f;
	}
}

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
		const uint32_t steps_word = ray_steps[i];
		actual = steps_word & ~RAY_STEPS_RESUME;
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
		// dead: terminated early (j < actual), or left the box before the chunk was full (:979-982); a ray that only ran out of hops lives on
		alive = j == actual && (actual == n_steps || (steps_word & RAY_STEPS_RESUME));
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

// tonemap(x, curve) (render_buffer.cu:272-329): identity, or a rational polynomial (ACES, Hable) / luminance (Reinhard) curve on the clamped colour
__device__ __forceinline__ void tonemap_curve(float c[3], const int curve) {
	if (curve == NGPB_TONEMAP_IDENTITY) return;
	#pragma unroll
	for (int k = 0; k < 3; ++k) c[k] = fmaxf(c[k], 0.f);
	float k0, k1, k2, k3, k4, k5;
	if (curve == NGPB_TONEMAP_ACES) {
		k0 = 0.6f * 0.6f * 2.51f; k1 = 0.6f * 0.03f; k2 = 0.0f; k3 = 0.6f * 0.6f * 2.43f; k4 = 0.6f * 0.59f; k5 = 0.14f;
	} else if (curve == NGPB_TONEMAP_HABLE) {
		const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
		k0 = A * F - A * E; k1 = C * B * F - B * E; k2 = 0.0f; k3 = A * F; k4 = B * F; k5 = D * F * F;
		const float W = 11.2f;
		const float nom = k0 * (W * W) + k1 * W + k2, denom = k3 * (W * W) + k4 * W + k5;
		const float white_scale = denom / nom;
		k0 = 4.0f * k0 * white_scale; k1 = 2.0f * k1 * white_scale; k2 = k2 * white_scale; k3 = 4.0f * k3; k4 = 2.0f * k4;
	} else { // Reinhard
		const float Y = sum3(0.2126f * c[0], 0.7152f * c[1], 0.0722f * c[2]);
		const float s = 1.f / (Y + 1.0f);
		#pragma unroll
		for (int k = 0; k < 3; ++k) c[k] = c[k] * s;
		return;
	}
	#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const float sq = c[k] * c[k];
		c[k] = (sq * k0 + k1 * c[k] + k2) / (k3 * sq + k4 * c[k] + k5);
	}
}

// tonemap_kernel (render_buffer.cu:540-567), writing linear memory instead of a CUDA surface
__global__ void __launch_bounds__(256) render_tonemap_kernel(const uint32_t n_pixels, const float exposure_scale, float4 background_color, const float4* __restrict__ accumulate_buffer,
                                                             const int color_space, const int output_srgb, const int curve, float4* __restrict__ out)
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
	for (int k = 0; k < 3; ++k) { // 1. to linear, 2. exposure (render_buffer.cu:331-339)
		float v = c[k];
		if (color_space == NGPB_COLOR_SRGB) v = srgb_to_linear(v);
		c[k] = v * exposure_scale;
	}
	tonemap_curve(c, curve);      // 3. the curve, in linear space
	if (output_srgb) {            // 4. to the output colour space
		#pragma unroll
		for (int k = 0; k < 3; ++k) c[k] = linear_to_srgb(c[k]);
	}
	out[i] = make_float4(c[0], c[1], c[2], color.w);
}

Aabb make_aabb(const float* a);
// the frame's way out of render_frame, shared with the neural-image mode (model.cu)
void render_accumulate_launch(cudaStream_t stream, uint32_t n_pixels, const float* frame_rgba, float* accumulate_rgba, float sample_count, int color_space) {
	render_accumulate_kernel<<<div_round_up(n_pixels, 256u), 256, 0, stream>>>(n_pixels, (const float4*)frame_rgba, (float4*)accumulate_rgba, sample_count, color_space);
	NGPB_LAUNCH_CHECK();
}
void render_tonemap_launch(cudaStream_t stream, uint32_t n_pixels, float exposure, const float* background4, const float* accumulate_rgba, int color_space, int output_srgb, int curve, float* out_rgba) {
	render_tonemap_kernel<<<div_round_up(n_pixels, 256u), 256, 0, stream>>>(n_pixels, powf(2.0f, exposure), make_float4(background4[0], background4[1], background4[2], background4[3]),
		(const float4*)accumulate_rgba, color_space, output_srgb, curve, (float4*)out_rgba);
	NGPB_LAUNCH_CHECK();
}
void ld_random_pixel_offset_host(uint32_t spp, float* out2) { ld_random_pixel_offset(spp, out2); }

void hash_encode_forward_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* positions, uint32_t pos_stride, uint32_t n, const uint32_t* n_dev, __half* encoded, bool tiled);
void nerf_mlp_forward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma);
bool features_tiled();

} // namespace ngpb

using namespace ngpb;

namespace {
struct RenderWorkspace {
	float4 *frame, *accum, *out, *rgba[2];
	RenderRay* rays[2];
	uint32_t *ray_steps, *counters, *coarse;
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
	w.coarse = (uint32_t*)take(COARSE_WORDS * 4);
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
		uint8_t* pinned = g_pinned.get(256 + (size_t)n_pixels * 16);
		uint32_t* host_counter = reinterpret_cast<uint32_t*>(pinned);
		float* host_frame = reinterpret_cast<float*>(pinned + 256);
		uint32_t launches = 0;

		RenderParams P{};
		P.width = cfg->width; P.height = cfg->height; P.fx = cfg->fx; P.fy = cfg->fy; P.cx = cfg->screen_center[0]; P.cy = cfg->screen_center[1];
		for (int k = 0; k < 12; ++k) P.cam[k] = cfg->camera[k];
		P.render_aabb = make_aabb(cfg->render_aabb); P.train_aabb = make_aabb(cfg->aabb);
		P.cone_angle = cfg->cone_angle_constant; P.near_distance = cfg->near_distance; P.min_transmittance = cfg->min_transmittance;
		P.rgb_activation = cfg->rgb_activation; P.density_activation = cfg->density_activation; P.train_in_linear_colors = cfg->train_in_linear_colors;

		NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64, stream));
		const uint32_t* coarse = nullptr;
		if (empty_space_blocks()) { coarse_occupancy_launch(stream, bitfield, ws.coarse); coarse = ws.coarse; ++launches; }
		for (int s = 0; s < cfg->spp; ++s) {
			P.sample_index = (uint32_t)s;
			ld_random_pixel_offset(cfg->snap_to_pixel_centers ? 0u : (uint32_t)s, P.offset);
			NGPB_CUDA_CHECK(cudaMemsetAsync(frame, 0, (size_t)n_pixels * 16, stream)); // CudaRenderBuffer::clear_frame
			NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 8, stream));
			render_init_kernel<<<div_round_up(n_pixels, 128), 128, 0, stream>>>(P, bitfield, coarse, rays[0], rgba[0], counters + 0);
			NGPB_LAUNCH_CHECK(); ++launches;
			// Passes without a blocking read-back: a pass's launches are sized by (and its step count derived from) the PREVIOUS pass's live-ray count, whose
			// copy to the host completed while that pass's kernels were queued; the kernels read the current count from the device. The image does not depend
			// on the pass schedule (unlike the Blender path's), so a step count derived from the slightly larger previous count is as good as the exact one.
			uint32_t cur = 0, bound = n_pixels, n_alive0 = 0;
			static thread_local cudaEvent_t count_ready[2] = {nullptr, nullptr};
			for (auto& e : count_ready) if (!e) NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			for (uint32_t pass = 0; pass < 100000; ++pass) {
				NGPB_CUDA_CHECK(cudaMemcpyAsync(host_counter + 4 + (pass & 1), counters + cur, 4, cudaMemcpyDeviceToHost, stream));
				NGPB_CUDA_CHECK(cudaEventRecord(count_ready[pass & 1], stream));
				if (pass > 0) {
					NGPB_CUDA_CHECK(cudaEventSynchronize(count_ready[(pass - 1) & 1]));
					const uint32_t prev = host_counter[4 + ((pass - 1) & 1)];
					if (prev == 0) break; // the previous pass already ran on no rays
					bound = std::min(bound, prev);
					if (pass == 1) n_alive0 = prev;
				}
				// steps this pass: grows as rays die, bounded so that the slot count never exceeds the first pass's
				const uint32_t n_steps = pass == 0 ? NGPB_RENDER_FIRST_PASS_STEPS
					: (uint32_t)std::min<uint64_t>(NGPB_RENDER_MAX_PASS_STEPS, std::max<uint64_t>(NGPB_RENDER_FIRST_PASS_STEPS, (uint64_t)n_alive0 * NGPB_RENDER_FIRST_PASS_STEPS / bound));
				const uint32_t n_slots = next_multiple(bound * n_steps, 128);
				const uint32_t blocks = div_round_up(bound, 128);
				NGPB_CUDA_CHECK(cudaMemsetAsync(counters + (cur ^ 1), 0, 4, stream));
				render_slot_count_kernel<<<1, 1, 0, stream>>>(counters + cur, n_steps, counters + 8);
				static const bool warp_march = [] { const char* e = std::getenv("NGPB_RENDER_WARP_MARCH"); return !e || std::atoi(e) != 0; }();
				static const uint32_t max_rounds = [] { const char* e = std::getenv("NGPB_RENDER_ROUNDS"); return e ? (uint32_t)std::atoi(e) : 16u; }();
				if (warp_march) render_march_warp_kernel<<<div_round_up(bound, 8), 256, 0, stream>>>(P, counters + cur, n_steps, max_rounds, bitfield, rays[cur], coords, ray_steps);
				else render_march_kernel<<<blocks, 128, 0, stream>>>(P, counters + cur, n_steps, render_max_hops(), bitfield, coarse, rays[cur], coords, ray_steps);
				NGPB_LAUNCH_CHECK();
				hash_encode_forward_launch(stream, g, (const __half*)params + MLP_PARAMS, coords, COORD_FLOATS, n_slots, counters + 8, encoded, features_tiled());
				nerf_mlp_forward_launch(stream, (const __half*)params, encoded, features_tiled(), coords, n_slots, counters + 8, rgbsigma);
				render_composite_kernel<<<blocks, 128, 0, stream>>>(P, counters + cur, n_steps, rays[cur], rgba[cur], coords, rgbsigma, ray_steps,
					rays[cur ^ 1], rgba[cur ^ 1], counters + (cur ^ 1), frame, reinterpret_cast<unsigned long long*>(counters + 2));
				NGPB_LAUNCH_CHECK();
				launches += 5;
				cur ^= 1;
			}
			if (s == 0) NGPB_CUDA_CHECK(cudaMemsetAsync(accum, 0, (size_t)n_pixels * 16, stream));
			render_accumulate_kernel<<<div_round_up(n_pixels, 256), 256, 0, stream>>>(n_pixels, frame, accum, (float)s, cfg->color_space);
			NGPB_LAUNCH_CHECK(); ++launches;
		}
		render_tonemap_kernel<<<div_round_up(n_pixels, 256), 256, 0, stream>>>(n_pixels, powf(2.0f, cfg->exposure),
			make_float4(cfg->background_color[0], cfg->background_color[1], cfg->background_color[2], cfg->background_color[3]), accum, cfg->color_space, cfg->output_srgb,
			cfg->tonemap_curve, out);
		NGPB_LAUNCH_CHECK(); ++launches;
		const bool direct = host_pointer_is_pinned(out_rgba_host);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(direct ? out_rgba_host : host_frame, out, (size_t)n_pixels * 16, cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaMemcpyAsync(host_counter, counters + 2, 8, cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		if (!direct) parallel_host_copy(out_rgba_host, host_frame, (size_t)n_pixels * 16);
		if (n_samples_out) std::memcpy(n_samples_out, host_counter, 8);
		if (n_launches_out) *n_launches_out = launches;
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

// =====================================================================================================================
// K18: the Blender multi-NeRF renderer (the fork's addition). Replaces NerfRenderer::render (reference: src/nerf_renderer.cu:565-791):
// init_global_rays_kernel :17-92, init_proxy_rays_kernel :94-146, compact_rays_kernel :240-270, march_active_rays :274-315,
// cull_global_rays_and_set_proxy_rays_active_kernel :376-428, march_proxy_rays_and_generate_next_network_inputs :317-373,
// composite_proxy_ray_colors_kernel :431-517, shade_buffer_with_rays_kernel :519-563, and Testbed::bl_render_frame (src/testbed.cu:2675).
//
// One global ray per (down-sampled) pixel and one proxy ray per NeRF in that NeRF's local frame. Every wave the nearest live proxy of a
// ray becomes its active one, marches clamp(n_initial / n_alive, 1, 8) steps through its NeRF and is composited into the global ray's
// colour with the NeRF's opacity. The image depends on that wave schedule, so the schedule is the reference's, including one live-ray
// count read-back per wave; inside a wave the per-ray kernels are merged (init global+proxy, march+cull, compact+shade) and compaction
// uses one atomic per warp. Each NeRF brings its own parameters, hash-grid geometry and occupancy bits (an ngpb_field).
// Perspective camera; render masks and depth of field are not built.
// =====================================================================================================================
namespace ngpb {

struct BlGlobalRay { float o[3]; uint32_t idx; float d[3]; uint32_t alive; float rgba[4]; };
struct BlProxyRay { float o[3]; float t; float d[3]; uint32_t n_steps; uint32_t alive, active; uint32_t pad[2]; };
// Mask3D on the device (nerf/mask_3D.cuh:128-257): only the inverse transform is needed
struct BlMask { int32_t shape, mode; float itransform[16]; float config[6]; float feather, opacity; };
enum { MASK_BOX = 0, MASK_CYLINDER = 1, MASK_SPHERE = 2, MASK_ALL = 3, MASK_ADD = 0, MASK_SUBTRACT = 1 };
struct BlNerfProps {
	float transform[16], itransform[16]; // column-major
	const BlMask* masks;                 // [All(opposite of the first mask's mode)] + the NeRF's own masks + the request's masks in local space (render_modifiers.cuh:28-62)
	uint32_t n_masks, pad_;
	const uint8_t* bitfield;
	const uint32_t* coarse; // one bit per 8^3 block of occupancy cells (see hop_over_empty), or null
	Aabb render_aabb, train_aabb;
	float cone_angle, opacity, min_transmittance;
	int32_t rgb_activation, density_activation;
};
struct BlRequest {
	int32_t width, height, skip, scaled_w, scaled_h, flip_y;
	float cam[12], focal_length, near_distance, offset[2];
	int32_t camera_model;
	float aperture_size, focus_z, sq[3], qh[24];
};
constexpr uint32_t BL_MAX_NERFS = 16;

__device__ __forceinline__ float sum4(float a, float b, float c, float d) { return (a + b) + (c + d); }
__device__ __forceinline__ V3 xform_point(const float* M, const V3& p) {
	return {sum4(M[0] * p.x, M[4] * p.y, M[8] * p.z, M[12]), sum4(M[1] * p.x, M[5] * p.y, M[9] * p.z, M[13]), sum4(M[2] * p.x, M[6] * p.y, M[10] * p.z, M[14])};
}
__device__ __forceinline__ V3 xform_dir(const float* M, const V3& d) {
	return {sum3(M[0] * d.x, M[4] * d.y, M[8] * d.z), sum3(M[1] * d.x, M[5] * d.y, M[9] * d.z), sum3(M[2] * d.x, M[6] * d.y, M[10] * d.z)};
}
__device__ __forceinline__ V3 normalized3(const V3& v) {
	const float z = sum3(v.x * v.x, v.y * v.y, v.z * v.z);
	if (z > 0.f) { const float n = sqrtf(z); return {v.x / n, v.y / n, v.z / n}; }
	return v;
}

// ---- Mask3D (nerf/mask_3D.cuh) ----
__device__ __forceinline__ float mask_signed_distance(const BlMask& m, const V3& p) { // signed_distance_to_point, :162-183
	const V3 q = xform_point(m.itransform, p);
	float d = 0.0f;
	if (m.shape == MASK_BOX) { // sdf_box :31-34
		const float dx = fabsf(q.x) - 0.5f * m.config[0], dy = fabsf(q.y) - 0.5f * m.config[1], dz = fabsf(q.z) - 0.5f * m.config[2];
		const float ox = fmaxf(dx, 0.0f), oy = fmaxf(dy, 0.0f), oz = fmaxf(dz, 0.0f);
		d = sqrtf(sum3(ox * ox, oy * oy, oz * oz)) + fminf(fmaxf(dx, fmaxf(dy, dz)), 0.0f);
	} else if (m.shape == MASK_CYLINDER) { // sdf_cylinder :36-39: radial distance in xy, half height along z
		const float dr = fabsf(sqrtf(q.y * q.y + q.x * q.x)) - m.config[0], dh = fabsf(q.z) - 0.5f * m.config[1];
		const float orr = fmaxf(dr, 0.0f), oh = fmaxf(dh, 0.0f);
		d = sqrtf(orr * orr + oh * oh) + fminf(fmaxf(dr, dh), 0.0f);
	} else if (m.shape == MASK_SPHERE) {
		d = sqrtf(sum3(q.x * q.x, q.y * q.y, q.z * q.z)) - m.config[0];
	} else {
		d = -1.0f;
	}
	return d * (m.mode == MASK_ADD ? 1.0f : -1.0f);
}
__device__ __forceinline__ float mask_sample(const BlMask& m, const V3& p) { // Mask3D::sample :194-213
	const float k = m.mode == MASK_ADD ? 1.0f : -1.0f;
	if (m.shape == MASK_ALL) return k;
	const float d = mask_signed_distance(m, p);
	const float alpha = m.feather == 0.0f ? (d < 0.0f ? 1.0f : 0.0f) : fminf(fmaxf(0.5f - d / m.feather, 0.0f), 1.0f);
	return m.opacity * alpha * k;
}
__device__ __forceinline__ bool plane_hit(const V3& o, const V3& d, const float nz, const float pz, float* t) { // intersect_plane_ray :74-81 for n = (0, 0, nz), p = (0, 0, pz)
	const float denom = nz * d.z;
	if (denom > 1e-6f) { *t = ((pz - o.z) * nz) / denom; return *t >= 0.0f; }
	return false;
}
__device__ __forceinline__ bool mask_intersects_ray(const BlMask& m, const V3& ro, const V3& rd) { // Mask3D::intersects_ray :215-247
	if (m.mode == MASK_SUBTRACT) return true;
	if (m.shape == MASK_ALL) return m.mode == MASK_ADD;
	const V3 o = xform_point(m.itransform, ro);
	const V3 d = normalized3(xform_dir(m.itransform, rd));
	if (m.shape == MASK_BOX) { // ray_intersects_box :46-56 with box_dims + 0.5 * feather
		float tmin_max = -INFINITY, tmax_min = INFINITY;
		const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
		#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const float size = m.config[a] + 0.5f * m.feather, inv = 1.0f / dd[a];
			const float t0 = (-0.5f * size - oo[a]) * inv, t1 = (0.5f * size - oo[a]) * inv;
			tmin_max = fmaxf(tmin_max, fminf(t0, t1)); tmax_min = fminf(tmax_min, fmaxf(t0, t1));
		}
		return tmin_max <= tmax_min;
	}
	if (m.shape == MASK_SPHERE) { // ray_intersects_sphere :59-63
		const float radius = m.config[0] + 0.5f * m.feather;
		const float dot = sum3(d.x * o.x, d.y * o.y, d.z * o.z);
		const float a = dot * dot, b = sum3(o.x * o.x, o.y * o.y, o.z * o.z) - radius * radius;
		return !((a - b) < 0.0f);
	}
	// ray_intersects_cylinder :83-123
	const float radius = m.config[0] + 0.5f * m.feather, height = m.config[1] + 0.5f * m.feather;
	const float a = d.x * d.x + d.y * d.y, b = 2.0f * (d.x * o.x + d.y * o.y), c = (o.x * o.x + o.y * o.y) - radius * radius;
	const float disc = b * b - 4.0f * a * c;
	if (disc < 0.0f) return false;
	const float sq = sqrtf(disc), a2 = 2.0f * a, h2 = 0.5f * height;
	if (a2 > 1e-6f) {
		const float t0 = (-b - sq) / a2, t1 = (-b + sq) / a2;
		const float z0 = o.z + t0 * d.z, z1 = o.z + t1 * d.z;
		if ((z0 >= -h2 && z0 <= h2) || (z1 >= -h2 && z1 <= h2)) return true;
	}
	float t = 0.0f;
	if (plane_hit(o, d, 1.0f, h2, &t)) { const float px = o.x + t * d.x, py = o.y + t * d.y; if (px * px + py * py <= radius * radius) return true; }
	if (plane_hit(o, d, -1.0f, -h2, &t)) { const float px = o.x + t * d.x, py = o.y + t * d.y; if (px * px + py * py <= radius * radius) return true; }
	return false;
}

// ---- camera models (camera_models.cuh) ----
__device__ __forceinline__ void ld_random_val_2d_dev(uint32_t index, uint32_t seed, float out[2]) { // random_val.cuh:261-268,:277-281
	const float S = float(1.0 / (1ull << 32));
	index = nested_uniform_scramble_base2(index, seed);
	out[0] = (float)nested_uniform_scramble_base2(sobol_dim(index, 0), hash_combine(seed, 0u)) * S;
	out[1] = (float)nested_uniform_scramble_base2(sobol_dim(index, 1), hash_combine(seed, 1u)) * S;
}
__device__ __forceinline__ void square2disk_shirley(const float a, const float b, float out[2]) { // random_val.cuh:109-125
	const float PI = 3.14159265358979323846f;
	float phi, r;
	if (a * a > b * b) { r = a; phi = (PI / 4.0f) * (b / a); } else { r = b; phi = (PI / 2.0f) - (PI / 4.0f) * (a / b); }
	float s, c;
	sincosf(phi, &s, &c);
	out[0] = r * c; out[1] = r * s;
}
// thin-lens blur shared by the three models (:99-104, :196-201, :232-237): sample index 0 ("todo: sample index", nerf_renderer.cu:624)
__device__ __forceinline__ void apply_depth_of_field(const BlRequest& R, const uint32_t px, const uint32_t py, V3& origin, V3& dir) {
	if (!(R.aperture_size > 0.0f)) return;
	const V3 lookat = {origin.x + dir.x * R.focus_z, origin.y + dir.y * R.focus_z, origin.z + dir.z * R.focus_z};
	float u[2], disk[2];
	ld_random_val_2d_dev(0u, px * 19349663u + py * 96925573u, u);
	square2disk_shirley(u[0] * 2.0f - 1.0f, u[1] * 2.0f - 1.0f, disk);
	const float bx = R.aperture_size * disk[0], by = R.aperture_size * disk[1];
	origin.x += R.cam[0] * bx + R.cam[3] * by; origin.y += R.cam[1] * bx + R.cam[4] * by; origin.z += R.cam[2] * bx + R.cam[5] * by;
	dir = {(lookat.x - origin.x) / R.focus_z, (lookat.y - origin.y) / R.focus_z, (lookat.z - origin.z) / R.focus_z};
}
__device__ __forceinline__ V3 cam_rotate(const BlRequest& R, const V3& v) {
	const float row0[3] = {R.cam[0], R.cam[3], R.cam[6]}, row1[3] = {R.cam[1], R.cam[4], R.cam[7]}, row2[3] = {R.cam[2], R.cam[5], R.cam[8]};
	const float vv[3] = {v.x, v.y, v.z};
	return {dot3(row0, vv), dot3(row1, vv), dot3(row2, vv)};
}
// pixel -> (origin, un-normalised direction) for the request's camera model (init_global_rays_kernel, nerf_renderer.cu:44-82)
__device__ __forceinline__ void blender_pixel_to_ray(const BlRequest& R, const uint32_t x, const uint32_t y, V3& origin, V3& dir) {
	const float W = (float)R.width, H = (float)R.height;
	if (R.camera_model == NGPB_CAMERA_PERSPECTIVE) { // perspective_pixel_to_ray :205-241
		const float uvx = ((float)x + R.offset[0]) / W, uvy = ((float)y + R.offset[1]) / H;
		dir = cam_rotate(R, V3{(uvx - 0.5f) * W / R.focal_length, (uvy - 0.5f) * H / R.focal_length, 1.0f});
		origin = {R.cam[9], R.cam[10], R.cam[11]};
	} else if (R.camera_model == NGPB_CAMERA_QUADRILATERAL_HEXAHEDRON) { // quadrilateral_hexahedron_pixel_to_ray :82-110
		const float ux = ((float)x + 0.5f) / W, uy = ((float)y + 0.5f) / H;
		const float* q = R.qh;
		float fp[3], bp[3];
		#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const float f_ab = q[0 + k] + ux * (q[3 + k] - q[0 + k]), f_dc = q[6 + k] + ux * (q[9 + k] - q[6 + k]);
			fp[k] = f_ab + uy * (f_dc - f_ab);
			const float b_ab = q[12 + k] + ux * (q[15 + k] - q[12 + k]), b_dc = q[18 + k] + ux * (q[21 + k] - q[18 + k]);
			bp[k] = b_ab + uy * (b_dc - b_ab);
		}
		V3 d = {fp[0] - bp[0], fp[1] - bp[1], fp[2] - bp[2]};
		d = {d.x / d.z, d.y / d.z, d.z / d.z};
		const V3 o = cam_rotate(R, V3{bp[0], bp[1], bp[2]});
		origin = {o.x + R.cam[9], o.y + R.cam[10], o.z + R.cam[11]};
		dir = cam_rotate(R, d);
	} else { // spherical_quadrilateral_pixel_to_ray :159-203
		const float PI = 3.14159265358979323846f;
		const float sw = R.sq[0], sh = R.sq[1], curvature = R.sq[2];
		const float max_len = sqrtf(sw * sw + sh * sh);
		const float ux = 2.0f * (((float)x + 0.5f) / W - 0.5f), uy = 2.0f * (((float)y + 0.5f) / H - 0.5f);
		const float px = sw * ux, py = sh * uy;
		const float az = atan2f(py, px), r = sqrtf(px * px + py * py);
		// walk_along_sphere / walk_along_circle :137-157
		float rz0 = 0.0f, rz1 = 0.0f;
		const float arc_t = r / (2.0f * max_len);
		if (!(arc_t == 0.0f || max_len == 0.0f)) {
			if (curvature == 0.0f) { rz0 = max_len * arc_t; }
			else { const float tpc = 2.0f * PI * curvature, s_tpc = max_len / tpc; rz0 = s_tpc * sinf(tpc * arc_t); rz1 = s_tpc * (1.0f - cosf(tpc * arc_t)); }
		}
		const V3 o_local = {rz0 * cosf(az), rz0 * sinf(az), rz1};
		V3 d_local = {0.0f, 0.0f, 1.0f};
		if (curvature != 0.0f) {
			const V3 to_center = {0.0f - o_local.x, 0.0f - o_local.y, max_len / (2.0f * PI * curvature) - o_local.z};
			const V3 n = normalized3(to_center);
			const float k = curvature > 0.0f ? 1.0f : -1.0f;
			d_local = {k * n.x, k * n.y, k * n.z};
		}
		const V3 o = cam_rotate(R, o_local);
		origin = {o.x + R.cam[9], o.y + R.cam[10], o.z + R.cam[11]};
		dir = cam_rotate(R, d_local);
	}
	apply_depth_of_field(R, x, y, origin, dir);
	origin = {origin.x + dir.x * R.near_distance, origin.y + dir.y * R.near_distance, origin.z + dir.z * R.near_distance};
}

// hit_test_and_march (:149-211), no masks: advances t to the next sample position inside an occupied cell; false = left the NeRF's render box
__device__ __forceinline__ bool hit_test_and_march(const V3& o, const V3& d, const V3& idir, float t_in, const BlNerfProps& P, float* t_out, float* dt_out) {
	float t = t_in, dt = 0.0f, prev_t = t;
	while (true) {
		const V3 pos = V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
		if (!aabb_contains(P.render_aabb, pos)) { *t_out = prev_t; if (dt_out) *dt_out = dt; return false; }
		dt = calc_dt(t, P.cone_angle);
		const uint32_t mip = (uint32_t)max(0, mip_from_dt(dt, pos));
		const uint32_t idx = cascaded_grid_idx_at(pos, mip);
		if (P.bitfield[idx / 8 + grid_mip_offset(mip) / 8] & (1 << (idx % 8))) break;
		prev_t = t;
		t = hop_over_empty(t, dt, P.cone_angle, pos, d, idir, mip, idx, P.coarse);
	}
	*t_out = t; if (dt_out) *dt_out = dt;
	return true;
}

// init_global_rays_kernel + init_proxy_rays_kernel: one thread per down-sampled pixel
__global__ void __launch_bounds__(128) bl_init_rays_kernel(const BlRequest R, const uint32_t n_nerfs, const BlNerfProps* __restrict__ props, const uint32_t stride,
                                                           BlGlobalRay* __restrict__ rays, BlProxyRay* __restrict__ proxies)
{
	const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= (uint32_t)R.scaled_w * (uint32_t)R.scaled_h) return;
	const uint32_t x = (idx % (uint32_t)R.scaled_w) * R.skip, y = (idx / (uint32_t)R.scaled_w) * R.skip;
	V3 go, dw;
	blender_pixel_to_ray(R, x, y, go, dw);
	const V3 gd = normalized3(dw);
	BlGlobalRay g;
	g.o[0] = go.x; g.o[1] = go.y; g.o[2] = go.z; g.d[0] = gd.x; g.d[1] = gd.y; g.d[2] = gd.z;
	g.idx = idx; g.alive = 1u; g.rgba[0] = g.rgba[1] = g.rgba[2] = g.rgba[3] = 0.f;
	rays[idx] = g;
	for (uint32_t n = 0; n < n_nerfs; ++n) {
		const BlNerfProps& P = props[n];
		BlProxyRay p{};
		const V3 o = xform_point(P.itransform, go);
		const V3 d = normalized3(xform_dir(P.itransform, normalized3(gd)));
		float tmin, tmax;
		aabb_ray_intersect(P.render_aabb, o, d, &tmin, &tmax);
		const float t = fmaxf(tmin, 0.0f) + 1e-5f;
		p.d[0] = d.x; p.d[1] = d.y; p.d[2] = d.z;
		if (aabb_contains(P.render_aabb, V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t})) {
			bool hits_a_mask = P.n_masks == 0; // a NeRF with masks is only traced by rays that cross one of them (:127-140)
			for (uint32_t k = 0; k < P.n_masks && !hits_a_mask; ++k) hits_a_mask = mask_intersects_ray(P.masks[k], o, d);
			p.active = 1u; p.alive = hits_a_mask ? 1u : 0u; p.t = 0.0f; p.n_steps = 0;
			p.o[0] = o.x + t * d.x; p.o[1] = o.y + t * d.y; p.o[2] = o.z + t * d.z;
		}
		proxies[(size_t)n * stride + idx] = p;
	}
}

// compact_rays_kernel + shade_buffer_with_rays_kernel: live rays (with their proxies) move to the next list, finished rays are shaded
__global__ void __launch_bounds__(128) bl_compact_kernel(const BlRequest R, const uint32_t* __restrict__ n_in_dev, const uint32_t n_nerfs, const uint32_t stride,
                                                         const BlGlobalRay* __restrict__ rays_in, const BlProxyRay* __restrict__ proxies_in,
                                                         BlGlobalRay* __restrict__ rays_out, BlProxyRay* __restrict__ proxies_out, uint32_t* __restrict__ n_out_dev,
                                                         float4* __restrict__ frame_buffer)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const bool in_range = i < *n_in_dev;
	BlGlobalRay g{};
	if (in_range) g = rays_in[i];
	const bool alive = in_range && g.alive;
	const uint32_t slot = warp_append(alive, n_out_dev);
	if (alive) {
		rays_out[slot] = g;
		for (uint32_t n = 0; n < n_nerfs; ++n) proxies_out[(size_t)n * stride + slot] = proxies_in[(size_t)n * stride + i];
	} else if (in_range && g.rgba[3] > 0.001f) {
		const uint32_t x = R.skip * (g.idx % (uint32_t)R.scaled_w);
		uint32_t y = R.skip * (g.idx / (uint32_t)R.scaled_w);
		if (R.flip_y) y = R.height - y - 1;
		const float4 tmp = make_float4(srgb_to_linear(g.rgba[0]), srgb_to_linear(g.rgba[1]), srgb_to_linear(g.rgba[2]), g.rgba[3]); // train_in_linear_colors = false (:599)
		const uint32_t max_pixels = (uint32_t)R.width * (uint32_t)R.height;
		for (int u = 0; u < R.skip; ++u) for (int v = 0; v < R.skip; ++v) {
			const uint32_t pix = (x + u) + (y + v) * (uint32_t)R.width;
			if (pix >= max_pixels) continue;
			const float4 f = frame_buffer[pix];
			const float k = 1.0f - tmp.w;
			frame_buffer[pix] = make_float4(tmp.x + f.x * k, tmp.y + f.y * k, tmp.z + f.z * k, tmp.w + f.w * k);
		}
	}
}

// The wave schedule on the device. The reference reads the live-ray count back after every compaction and derives the wave's step count from it on the
// host (nerf_renderer.cu:650-700); here the count stays on the device: a one-thread kernel derives n_steps = clamp(n_initial / n_alive, 1, max_steps), the
// sample-slot count and the running step index from it, and the wave's kernels read them from there. The host only needs an UPPER BOUND of the live count
// to size its launches -- the previous wave's count, whose read-back has long completed by the time it is needed -- so it never waits for the GPU.
struct BlWave { uint32_t n_steps, n_slots, current_step, next_step; };
__global__ void bl_wave_setup_kernel(const uint32_t* __restrict__ n_alive_dev, const uint32_t n_initial, const uint32_t max_steps, BlWave* __restrict__ wave)
{
	const uint32_t n_alive = *n_alive_dev;
	const uint32_t n_steps = n_alive ? max(1u, min(max_steps, n_initial / n_alive)) : 1u;
	wave->n_steps = n_steps;
	wave->n_slots = n_alive * n_steps;
	wave->current_step = wave->next_step;
	wave->next_step += n_steps;
}

// march_active_rays + cull_global_rays_and_set_proxy_rays_active_kernel
__global__ void __launch_bounds__(128) bl_march_cull_kernel(const uint32_t* __restrict__ n_alive_dev, const uint32_t n_nerfs, const BlNerfProps* __restrict__ props, const uint32_t stride,
                                                            const float cam_x, const float cam_y, const float cam_z, BlGlobalRay* __restrict__ rays, BlProxyRay* __restrict__ proxies)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *n_alive_dev) return;
	if (!rays[i].alive) return;
	float min_d2 = 0.0f;
	int active = -1;
	uint32_t n_proxy_alive = 0;
	for (uint32_t n = 0; n < n_nerfs; ++n) {
		BlProxyRay& pr = proxies[(size_t)n * stride + i];
		BlProxyRay p = pr;
		if (p.alive && p.active) {
			const V3 o = {p.o[0], p.o[1], p.o[2]}, d = {p.d[0], p.d[1], p.d[2]};
			const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
			float t;
			p.alive = hit_test_and_march(o, d, idir, p.t, props[n], &t, nullptr) ? 1u : 0u;
			p.t = t;
		}
		if (p.alive) {
			++n_proxy_alive;
			const V3 q = xform_point(props[n].transform, V3{p.o[0] + p.d[0] * p.t, p.o[1] + p.d[1] * p.t, p.o[2] + p.d[2] * p.t});
			const float dx = q.x - cam_x, dy = q.y - cam_y, dz = q.z - cam_z;
			const float d2 = sum3(dx * dx, dy * dy, dz * dz);
			if (d2 < min_d2 || active == -1) { min_d2 = d2; active = (int)n; }
			p.active = 0u;
		}
		pr = p;
	}
	if (active >= 0) proxies[(size_t)active * stride + i].active = 1u;
	if (n_proxy_alive == 0) rays[i].alive = 0u;
}

// march_proxy_rays_and_generate_next_network_inputs for one NeRF; ray-major slots [i * n_steps + j]
__global__ void __launch_bounds__(128) bl_generate_inputs_kernel(const uint32_t* __restrict__ n_alive_dev, const BlWave* __restrict__ wave, const BlNerfProps* __restrict__ props_n,
                                                                 const BlGlobalRay* __restrict__ rays, BlProxyRay* __restrict__ proxies_n, float* __restrict__ coords)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= *n_alive_dev) return;
	const uint32_t n_steps = wave->n_steps;
	float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
	BlProxyRay p = proxies_n[i];
	uint32_t j = 0;
	if (rays[i].alive && p.active) { // (the reference tests `active` only, :344-346)
		const BlNerfProps& P = *props_n;
		const V3 o = {p.o[0], p.o[1], p.o[2]}, d = {p.d[0], p.d[1], p.d[2]};
		const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
		const V3 wd = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f};
		float t = p.t, dt = calc_dt(t, P.cone_angle);
		bool done = false;
		for (; j < n_steps; ++j) {
			const V3 wp = warp_position(V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t}, P.train_aabb);
			c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
			c += COORD_FLOATS;
			if (!hit_test_and_march(o, d, idir, t, P, &t, &dt)) { p.n_steps = j; done = true; ++j; break; }
			t += dt;
		}
		if (!done) { p.t = t; p.n_steps = n_steps; }
		proxies_n[i] = p;
	}
	for (; j < n_steps; ++j) { c[0] = c[1] = c[2] = 0.5f; c[3] = 0.f; c[4] = c[5] = c[6] = 0.5f; c += COORD_FLOATS; } // unused slots still go through the network
}

// ---- warp-per-ray forms of the two march kernels for NeRFs with a constant step (cone_angle == 0, every aabb_scale-1 snapshot) ----
// hit_test_and_march finds the first member of the ray's t chain, from t on, that lies in an occupied cell (or leaves the render box first). The serial walk
// gets there by dependent cell / block hops with a handful of the warp's lanes active; here the 32 lanes test 32 consecutive chain members at once -- their
// exact t values come from the closed form of the chain (chain_advance) -- and a ballot picks the first. Same members, same occupancy tests; the two can
// differ only where a hop of the serial walk lands within rounding of a cell boundary (the tolerance the classic renderer's warp march already works to).
__device__ __forceinline__ bool warp_first_occupied(const V3& o, const V3& d, const float t_start, const BlNerfProps& P, const uint32_t lane, float* t_found) {
	float base = t_start;
	for (uint32_t round = 0; round < 256; ++round) { // 8192 chain members: far beyond any box a constant step is used in
		const float t = chain_advance(base, lane);
		const V3 pos = V3{o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
		const bool inside = aabb_contains(P.render_aabb, pos);
		bool occ = false;
		if (inside) {
			const uint32_t mip = (uint32_t)max(0, mip_from_dt(MIN_CONE_STEPSIZE, pos));
			const uint32_t idx = cascaded_grid_idx_at(pos, mip);
			occ = (P.bitfield[idx / 8 + grid_mip_offset(mip) / 8] & (1 << (idx % 8))) != 0;
		}
		const uint32_t out_mask = __ballot_sync(0xffffffffu, !inside), occ_mask = __ballot_sync(0xffffffffu, occ);
		const int f_out = __ffs(out_mask), f_occ = __ffs(occ_mask);
		if (f_out && (!f_occ || f_out < f_occ)) { *t_found = __shfl_sync(0xffffffffu, t, max(f_out - 2, 0)); return false; }
		if (f_occ) { *t_found = __shfl_sync(0xffffffffu, t, f_occ - 1); return true; }
		base = __shfl_sync(0xffffffffu, t + MIN_CONE_STEPSIZE, 31);
	}
	*t_found = base;
	return false;
}

__global__ void __launch_bounds__(256) bl_generate_inputs_warp_kernel(const uint32_t* __restrict__ n_alive_dev, const BlWave* __restrict__ wave, const BlNerfProps* __restrict__ props_n,
                                                                      const BlGlobalRay* __restrict__ rays, BlProxyRay* __restrict__ proxies_n, float* __restrict__ coords)
{
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (i >= *n_alive_dev) return;
	const uint32_t n_steps = wave->n_steps;
	float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
	BlProxyRay p = proxies_n[i];
	uint32_t w = 0; // sample slots written
	if (rays[i].alive && p.active) {
		const BlNerfProps& P = *props_n;
		const V3 o = {p.o[0], p.o[1], p.o[2]}, d = {p.d[0], p.d[1], p.d[2]};
		const float wdv[3] = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f};
		float t = p.t; // position of the next sample (member 0 of the window)
		bool done = false;
		while (w < n_steps && !done) {
			// one window: 32 consecutive chain members from t
			const float tl = chain_advance(t, lane);
			const V3 pos = V3{o.x + d.x * tl, o.y + d.y * tl, o.z + d.z * tl};
			const bool inside = aabb_contains(P.render_aabb, pos);
			bool occ = false;
			if (inside) {
				const uint32_t mip = (uint32_t)max(0, mip_from_dt(MIN_CONE_STEPSIZE, pos));
				const uint32_t idx = cascaded_grid_idx_at(pos, mip);
				occ = (P.bitfield[idx / 8 + grid_mip_offset(mip) / 8] & (1 << (idx % 8))) != 0;
			}
			const uint32_t out_mask = __ballot_sync(0xffffffffu, !inside), occ_mask = __ballot_sync(0xffffffffu, occ);
			const V3 wp = warp_position(pos, P.train_aabb);
			uint32_t cur = 0;
			while (true) { // the serial loop of march_proxy_rays_and_generate_next_network_inputs on the window's masks: sample at `cur`, then the first occupied member from there
				// the sample at member cur: that lane holds its position; lanes cur .. cur + 6 (mod 32) would be awkward, so broadcast the three coordinates
				const float sx = __shfl_sync(0xffffffffu, wp.x, cur), sy = __shfl_sync(0xffffffffu, wp.y, cur), sz = __shfl_sync(0xffffffffu, wp.z, cur);
				if (lane < COORD_FLOATS) c[w * COORD_FLOATS + lane] = lane == 0 ? sx : lane == 1 ? sy : lane == 2 ? sz : lane == 3 ? warp_dt(MIN_CONE_STEPSIZE) : wdv[lane - 4];
				++w;
				const uint32_t rest_occ = occ_mask >> cur, rest_out = out_mask >> cur;
				const int f_occ = __ffs(rest_occ), f_out = __ffs(rest_out);
				if (f_out && (!f_occ || f_out < f_occ)) { p.n_steps = w - 1; done = true; break; } // left the box before another occupied member
				if (!f_occ) { // none in the rest of this window: keep searching beyond it
					float tf;
					const float t32 = __shfl_sync(0xffffffffu, tl + MIN_CONE_STEPSIZE, 31);
					if (!warp_first_occupied(o, d, t32, P, lane, &tf)) { p.n_steps = w - 1; done = true; break; }
					t = tf + MIN_CONE_STEPSIZE;
					break;
				}
				const uint32_t k = cur + (uint32_t)f_occ - 1; // the occupied member found; the next sample sits one step behind it
				if (w == n_steps || k + 1 >= 32) { t = __shfl_sync(0xffffffffu, tl, k) + MIN_CONE_STEPSIZE; break; }
				cur = k + 1;
			}
		}
		if (!done) { p.t = t; p.n_steps = n_steps; }
		__syncwarp();
		if (lane == 0) proxies_n[i] = p;
	}
	for (uint32_t k = w * COORD_FLOATS + lane; k < n_steps * COORD_FLOATS; k += 32) c[k] = (k % COORD_FLOATS) == 3 ? 0.f : 0.5f; // unused slots still go through the network
}

// composite_proxy_ray_colors_kernel for one NeRF
__global__ void __launch_bounds__(128) bl_composite_kernel(const uint32_t* __restrict__ n_alive_dev, const BlWave* __restrict__ wave, const BlNerfProps* __restrict__ props_n,
                                                           BlGlobalRay* __restrict__ rays, BlProxyRay* __restrict__ proxies_n, const float* __restrict__ coords,
                                                           const __half* __restrict__ rgbsigma, unsigned long long* __restrict__ n_samples)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_alive = *n_alive_dev, n_steps = wave->n_steps, current_step = wave->current_step;
	uint32_t done = 0;
	if (i < n_alive && rays[i].alive) {
		BlProxyRay p = proxies_n[i];
		if (p.alive && p.active) {
			const BlNerfProps& P = *props_n;
			float4 local = make_float4(rays[i].rgba[0], rays[i].rgba[1], rays[i].rgba[2], rays[i].rgba[3]);
			const float* c = coords + (size_t)i * n_steps * COORD_FLOATS;
			const __half* out = rgbsigma + (size_t)i * n_steps * 4;
			uint32_t j = 0;
			for (; j < p.n_steps; ++j) {
				const uint2 raw = *reinterpret_cast<const uint2*>(out + j * 4);
				const __half2 h01 = *reinterpret_cast<const __half2*>(&raw.x), h23 = *reinterpret_cast<const __half2*>(&raw.y);
				const float T = 1.f - local.w;
				const float dt = unwarp_dt(c[j * COORD_FLOATS + 3]);
				const float alpha = 1.f - __expf(-network_to_density(__high2float(h23), P.density_activation) * dt);
				float weight = alpha * T;
				if (P.n_masks) { // masks add / subtract visibility at the sample's position in the NeRF's frame (:490-496)
					const V3 pos = unwarp_position(c + j * COORD_FLOATS, P.train_aabb);
					float mask_weight = 1.f;
					for (uint32_t k = 0; k < P.n_masks; ++k) mask_weight = fminf(fmaxf(mask_weight + mask_sample(P.masks[k], pos), 0.0f), 1.0f);
					weight *= mask_weight;
				}
				weight *= P.opacity;
				local.x += network_to_rgb(__low2float(h01), P.rgb_activation) * weight;
				local.y += network_to_rgb(__high2float(h01), P.rgb_activation) * weight;
				local.z += network_to_rgb(__low2float(h23), P.rgb_activation) * weight;
				local.w += weight;
				++done;
				if (local.w > (1.0f - P.min_transmittance)) {
					const float w = local.w;
					local.x /= w; local.y /= w; local.z /= w; local.w /= w;
					break;
				}
			}
			if (j < n_steps) { p.alive = 0u; p.n_steps = j + current_step; proxies_n[i] = p; }
			rays[i].rgba[0] = local.x; rays[i].rgba[1] = local.y; rays[i].rgba[2] = local.z; rays[i].rgba[3] = local.w;
		}
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) done += __shfl_xor_sync(0xffffffffu, done, o);
	if ((threadIdx.x & 31) == 0 && done) atomicAdd(n_samples, (unsigned long long)done);
}

} // namespace ngpb

// ---- a NeRF loaded from a snapshot, independent of any Testbed (NeuralRadianceField, nerf/neural_radiance_field.cuh) ----
struct ngpb_field {
	int device = 0;
	uint32_t aabb_scale = 1, max_cascade = 0, n_params = 0;
	ngpb_grid grid{};
	__half* params = nullptr;
	uint8_t* bitfield = nullptr;
	float train_aabb[6] = {0, 0, 0, 1, 1, 1};
	float cone_angle = 0.f;
};

extern "C" int ngpb_field_create(ngpb_field** out, int device, uint32_t aabb_scale, const ngpb_half* params_host, uint32_t n_params, const float* density_grid_host, uint32_t n_cells) {
	ngpb_field* f = nullptr;
	float *grid_dev = nullptr, *mean_dev = nullptr;
	try {
		if (!out || !params_host || aabb_scale == 0 || (aabb_scale & (aabb_scale - 1)) != 0 || aabb_scale > 128) { set_last_error("ngpb_field_create: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (ngpb_check_device(device) != 0) return NGPB_ERR_RUNTIME;
		NGPB_CUDA_CHECK(cudaSetDevice(device));
		f = new ngpb_field();
		f->device = device; f->aabb_scale = aabb_scale;
		while ((1u << f->max_cascade) < aabb_scale) ++f->max_cascade;
		const float half = 0.5f * (float)aabb_scale;
		for (int c = 0; c < 3; ++c) { f->train_aabb[c] = 0.5f - half; f->train_aabb[3 + c] = 0.5f + half; }
		f->cone_angle = aabb_scale <= 1 ? 0.0f : (1.0f / 256.0f);
		const float per_level_scale = std::exp(std::log(2048.0f * (float)aabb_scale / 16.0f) / 15.0f); // reset_network, src/testbed.cu:2313-2325
		// encoding.log2_hashmap_size is the one architecture value a snapshot may vary (configs/nerf/*.json): it follows from the parameter count
		uint32_t entries = 0;
		for (uint32_t log2_t = 14; log2_t <= 24; ++log2_t) {
			entries = ngpb_grid_init(&f->grid, 16, log2_t, 16, per_level_scale);
			if (n_params == MLP_PARAMS + 2 * entries) break;
		}
		if (n_params != MLP_PARAMS + 2 * entries) throw std::runtime_error("snapshot parameter count does not match the base NeRF architecture for this aabb_scale");
		if (ngpb_grid_device_scales(nullptr, &f->grid) != 0) throw std::runtime_error(ngpb_last_error());
		f->n_params = n_params;
		NGPB_CUDA_CHECK(cudaMalloc(&f->params, sizeof(__half) * n_params));
		NGPB_CUDA_CHECK(cudaMemcpy(f->params, params_host, sizeof(__half) * n_params, cudaMemcpyHostToDevice));
		const size_t bits_bytes = (size_t)NERF_GRID_CELLS * NERF_CASCADES / 8;
		NGPB_CUDA_CHECK(cudaMalloc(&f->bitfield, bits_bytes));
		NGPB_CUDA_CHECK(cudaMemset(f->bitfield, 0, bits_bytes));
		if (n_cells) {
			if (!density_grid_host || n_cells != NERF_GRID_CELLS * (f->max_cascade + 1)) throw std::runtime_error("Incompatible number of grid cascades.");
			NGPB_CUDA_CHECK(cudaMalloc(&grid_dev, sizeof(float) * n_cells));
			NGPB_CUDA_CHECK(cudaMalloc(&mean_dev, NGPB_MEAN_WORKSPACE_BYTES));
			NGPB_CUDA_CHECK(cudaMemcpy(grid_dev, density_grid_host, sizeof(float) * n_cells, cudaMemcpyHostToDevice));
			if (ngpb_update_bitfield(nullptr, f->max_cascade + 1, grid_dev, mean_dev, f->bitfield) != 0) throw std::runtime_error(ngpb_last_error());
			NGPB_CUDA_CHECK(cudaDeviceSynchronize());
			cudaFree(grid_dev); cudaFree(mean_dev);
		}
		NGPB_CUDA_CHECK(cudaDeviceSynchronize()); // a cudaMemcpy from pageable memory may return before its DMA has landed; renders run on the caller's streams
		*out = f;
		return 0;
	} catch (const std::exception& e) {
		if (grid_dev) cudaFree(grid_dev);
		if (mean_dev) cudaFree(mean_dev);
		if (f) { if (f->params) cudaFree(f->params); if (f->bitfield) cudaFree(f->bitfield); delete f; }
		set_last_error(e.what());
		return NGPB_ERR_RUNTIME;
	}
}
extern "C" void ngpb_field_destroy(ngpb_field* f) {
	if (!f) return;
	cudaSetDevice(f->device);
	if (f->params) cudaFree(f->params);
	if (f->bitfield) cudaFree(f->bitfield);
	delete f;
}

static bool invert4(const float* a, float* out) { // Gauss-Jordan with partial pivoting, column-major in and out (Eigen's Matrix4f::inverse in the reference)
	double m[4][8];
	for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { m[r][c] = a[c * 4 + r]; m[r][4 + c] = r == c ? 1.0 : 0.0; }
	for (int c = 0; c < 4; ++c) {
		int piv = c;
		for (int r = c + 1; r < 4; ++r) if (std::fabs(m[r][c]) > std::fabs(m[piv][c])) piv = r;
		if (std::fabs(m[piv][c]) < 1e-30) return false;
		if (piv != c) for (int k = 0; k < 8; ++k) std::swap(m[piv][k], m[c][k]);
		const double d = m[c][c];
		for (int k = 0; k < 8; ++k) m[c][k] /= d;
		for (int r = 0; r < 4; ++r) if (r != c) { const double f = m[r][c]; for (int k = 0; k < 8; ++k) m[r][k] -= f * m[c][k]; }
	}
	for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[c * 4 + r] = (float)m[r][4 + c];
	return true;
}

static void mul4(const float* a, const float* b, float* out) { // column-major out = a * b
	for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {
		float v = 0.f;
		for (int k = 0; k < 4; ++k) v += a[k * 4 + r] * b[c * 4 + k];
		out[c * 4 + r] = v;
	}
}

// The device mask list of one NeRF (RenderModifiers, nerf/render_modifiers.cuh:28-62): the NeRF's own masks, then the request's masks brought into the
// NeRF's frame (Mask3D::transformed_by(itransform)), and in front of them an `All` mask of the opposite mode of the first one, so that a list starting
// with an Add mask begins from "nothing visible" and one starting with a Subtract mask from "everything visible".
static void build_mask_list(const ngpb_nerf_instance& in, const float* nerf_itransform, const ngpb_blender_request* rq, std::vector<BlMask>& out) {
	auto push = [&](const ngpb_mask& m, const float* to_local) {
		if (m.shape < MASK_BOX || m.shape > MASK_ALL || (m.mode != MASK_ADD && m.mode != MASK_SUBTRACT)) throw std::runtime_error("ngpb_blender_render: invalid mask shape or mode");
		BlMask d{};
		d.shape = m.shape; d.mode = m.mode; d.feather = m.feather; d.opacity = m.opacity;
		std::memcpy(d.config, m.config, sizeof(d.config));
		float t[16];
		if (to_local) mul4(to_local, m.transform, t); else std::memcpy(t, m.transform, 64);
		if (!invert4(t, d.itransform)) throw std::runtime_error("ngpb_blender_render: singular mask transform");
		out.push_back(d);
	};
	const size_t first = out.size();
	if (in.n_masks && !in.masks) throw std::runtime_error("ngpb_blender_render: mask count without masks");
	for (uint32_t k = 0; k < in.n_masks; ++k) push(in.masks[k], nullptr);
	for (uint32_t k = 0; k < rq->n_masks; ++k) push(rq->masks[k], nerf_itransform);
	if (out.size() > first && out[first].shape != MASK_ALL) {
		BlMask all{};
		all.shape = MASK_ALL; all.mode = out[first].mode == MASK_ADD ? MASK_SUBTRACT : MASK_ADD; all.opacity = 1.0f;
		for (int k = 0; k < 4; ++k) all.itransform[k * 5] = 1.0f;
		out.insert(out.begin() + first, all);
	}
}

extern "C" int ngpb_blender_render(void* stream_, const ngpb_blender_request* rq, uint32_t n_nerfs, const ngpb_nerf_instance* nerfs, float* out_rgba_host,
                                   uint64_t* n_samples_out, uint32_t* n_launches_out) {
	uint8_t* ws = nullptr;
	try {
		if (!rq || !out_rgba_host || rq->width <= 0 || rq->height <= 0 || rq->mip < 0 || rq->mip > 12 || n_nerfs > BL_MAX_NERFS || (n_nerfs && !nerfs) || (rq->n_masks && !rq->masks) ||
			rq->camera_model < NGPB_CAMERA_PERSPECTIVE || rq->camera_model > NGPB_CAMERA_SPHERICAL_QUADRILATERAL || rq->tonemap_curve < NGPB_TONEMAP_IDENTITY || rq->tonemap_curve > NGPB_TONEMAP_REINHARD) {
			set_last_error("ngpb_blender_render: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		BlRequest R{};
		R.width = rq->width; R.height = rq->height; R.skip = 1 << rq->mip; R.flip_y = rq->flip_y;
		R.scaled_w = (rq->width + R.skip - 1) / R.skip; R.scaled_h = (rq->height + R.skip - 1) / R.skip;
		for (int k = 0; k < 12; ++k) R.cam[k] = rq->camera[k];
		R.focal_length = rq->focal_length; R.near_distance = rq->near_distance;
		R.camera_model = rq->camera_model; R.aperture_size = rq->aperture_size; R.focus_z = rq->focus_z;
		std::memcpy(R.sq, rq->spherical_quadrilateral, sizeof(R.sq)); std::memcpy(R.qh, rq->quadrilateral_hexahedron, sizeof(R.qh));
		ld_random_pixel_offset(0u, R.offset); // sample index 0 ("todo: sample index", :624)
		const uint32_t n_pixels = (uint32_t)rq->width * (uint32_t)rq->height, n_init = (uint32_t)R.scaled_w * (uint32_t)R.scaled_h;
		const uint32_t max_steps = 8, nn = std::max(n_nerfs, 1u);
		const size_t slots = (size_t)next_multiple(n_init, 128) * max_steps + 128;
		// workspace (allocated per call: a Blender frame is milliseconds, the allocation microseconds)
		size_t off = 0;
		auto reserve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
		const size_t o_frame = reserve((size_t)n_pixels * 16), o_accum = reserve((size_t)n_pixels * 16), o_out = reserve((size_t)n_pixels * 16);
		const size_t o_rays0 = reserve((size_t)n_init * sizeof(BlGlobalRay)), o_rays1 = reserve((size_t)n_init * sizeof(BlGlobalRay));
		const size_t o_prox0 = reserve((size_t)n_init * nn * sizeof(BlProxyRay)), o_prox1 = reserve((size_t)n_init * nn * sizeof(BlProxyRay));
		const size_t o_coords = reserve(slots * COORD_FLOATS * 4), o_enc = reserve(slots * N_ENC * 2), o_rgbs = reserve(slots * 8);
		const size_t o_props = reserve(sizeof(BlNerfProps) * nn), o_cnt = reserve(64), o_coarse = reserve((size_t)COARSE_WORDS * 4 * nn);
		ws = g_blender_ws.get(off);
		uint8_t* pinned = g_pinned.get(256 + (size_t)rq->width * (size_t)rq->height * 16);
		uint32_t* host_counter = reinterpret_cast<uint32_t*>(pinned);
		float* host_frame = reinterpret_cast<float*>(pinned + 256);
		float4 *frame = (float4*)(ws + o_frame), *accum = (float4*)(ws + o_accum), *out = (float4*)(ws + o_out);
		BlGlobalRay* rays[2] = {(BlGlobalRay*)(ws + o_rays0), (BlGlobalRay*)(ws + o_rays1)};
		BlProxyRay* prox[2] = {(BlProxyRay*)(ws + o_prox0), (BlProxyRay*)(ws + o_prox1)};
		float* coords = (float*)(ws + o_coords);
		__half *encoded = (__half*)(ws + o_enc), *rgbsigma = (__half*)(ws + o_rgbs);
		BlNerfProps* props_dev = (BlNerfProps*)(ws + o_props);
		uint32_t* counters = (uint32_t*)(ws + o_cnt);
		std::vector<BlNerfProps> props(nn);
		std::vector<BlMask> masks_host;
		std::vector<uint32_t> mask_first(nn + 1, 0u);
		for (uint32_t n = 0; n < n_nerfs; ++n) {
			float it[16];
			if (!invert4(nerfs[n].transform, it)) throw std::runtime_error("ngpb_blender_render: singular NeRF transform");
			build_mask_list(nerfs[n], it, rq, masks_host);
			mask_first[n + 1] = (uint32_t)masks_host.size();
		}
		BlMask* masks_dev = reinterpret_cast<BlMask*>(g_blender_masks.get(std::max<size_t>(masks_host.size(), 1) * sizeof(BlMask)));
		if (!masks_host.empty()) NGPB_CUDA_CHECK(cudaMemcpyAsync(masks_dev, masks_host.data(), masks_host.size() * sizeof(BlMask), cudaMemcpyHostToDevice, stream));
		for (uint32_t n = 0; n < n_nerfs; ++n) {
			const ngpb_nerf_instance& in = nerfs[n];
			if (!in.field) throw std::runtime_error("ngpb_blender_render: NeRF without a field");
			BlNerfProps& P = props[n];
			std::memcpy(P.transform, in.transform, 64);
			if (!invert4(in.transform, P.itransform)) throw std::runtime_error("ngpb_blender_render: singular NeRF transform");
			P.masks = masks_dev + mask_first[n]; P.n_masks = mask_first[n + 1] - mask_first[n];
			P.bitfield = in.field->bitfield;
			P.coarse = nullptr;
			if (empty_space_blocks()) {
				uint32_t* c = reinterpret_cast<uint32_t*>(ws + o_coarse) + (size_t)n * COARSE_WORDS;
				coarse_occupancy_launch(stream, in.field->bitfield, c);
				P.coarse = c;
			}
			P.render_aabb = make_aabb(in.aabb); P.train_aabb = make_aabb(in.field->train_aabb);
			P.cone_angle = in.field->cone_angle; P.opacity = in.opacity; P.min_transmittance = 0.01f; // NeuralRadianceField::min_transmittance
			P.rgb_activation = NGPB_ACT_LOGISTIC; P.density_activation = NGPB_ACT_EXPONENTIAL;           // NeuralRadianceField defaults
		}
		uint32_t launches = 0;
		NGPB_CUDA_CHECK(cudaMemcpyAsync(props_dev, props.data(), sizeof(BlNerfProps) * nn, cudaMemcpyHostToDevice, stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(frame, 0, (size_t)n_pixels * 16, stream)); // CudaRenderBuffer::clear_frame
		NGPB_CUDA_CHECK(cudaMemsetAsync(accum, 0, (size_t)n_pixels * 16, stream));
		NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 64, stream));
		if (n_nerfs) { // "if (render_request.nerfs.size() == 0) return;" (:570): an empty request renders the background
			bl_init_rays_kernel<<<div_round_up(n_init, 128), 128, 0, stream>>>(R, n_nerfs, props_dev, n_init, rays[0], prox[0]);
			NGPB_LAUNCH_CHECK(); ++launches;
			NGPB_CUDA_CHECK(cudaMemcpyAsync(counters + 0, &n_init, 4, cudaMemcpyHostToDevice, stream));
			uint32_t cur = 0, bound = n_init; // bound >= the live-ray count of the wave being launched
			// warp-per-ray march kernels when every NeRF of the request steps with the constant dt (NGPB_BLENDER_WARP_MARCH=0: the serial kernels)
			static const bool warp_march_enabled = [] { const char* e = std::getenv("NGPB_BLENDER_WARP_MARCH"); return !e || std::atoi(e) != 0; }();
			bool warp_march = warp_march_enabled;
			for (uint32_t n = 0; n < n_nerfs; ++n) warp_march = warp_march && props[n].cone_angle == 0.f;
			BlWave* wave = reinterpret_cast<BlWave*>(counters + 8);
			const BlWave wave0{1u, 0u, 1u, 1u}; // the reference's `n_steps_total = 1` before the first wave
			NGPB_CUDA_CHECK(cudaMemcpyAsync(wave, &wave0, sizeof(BlWave), cudaMemcpyHostToDevice, stream));
			static thread_local cudaEvent_t count_ready[2] = {nullptr, nullptr};
			for (auto& e : count_ready) if (!e) NGPB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
			for (uint32_t k = 0; k < 10000; ++k) {
				NGPB_CUDA_CHECK(cudaMemsetAsync(counters + (cur ^ 1), 0, 4, stream));
				bl_compact_kernel<<<div_round_up(bound, 128), 128, 0, stream>>>(R, counters + cur, n_nerfs, n_init, rays[cur], prox[cur], rays[cur ^ 1], prox[cur ^ 1], counters + (cur ^ 1), frame);
				NGPB_LAUNCH_CHECK(); ++launches;
				cur ^= 1;
				// this wave's live count travels to the host without anybody waiting for it ...
				NGPB_CUDA_CHECK(cudaMemcpyAsync(host_counter + 4 + (k & 1), counters + cur, 4, cudaMemcpyDeviceToHost, stream));
				NGPB_CUDA_CHECK(cudaEventRecord(count_ready[k & 1], stream));
				// ... the launches below only need a bound: the PREVIOUS wave's count (rays only die), whose copy finished while that wave's kernels were queued
				if (k > 0) {
					NGPB_CUDA_CHECK(cudaEventSynchronize(count_ready[(k - 1) & 1]));
					bound = std::min(bound, host_counter[4 + ((k - 1) & 1)]);
					if (bound == 0) break; // the previous wave already ran on no rays: the frame is complete (this wave's compaction was a no-op)
				}
				const uint32_t blocks = div_round_up(bound, 128);
				bl_wave_setup_kernel<<<1, 1, 0, stream>>>(counters + cur, n_init, max_steps, wave);
				NGPB_LAUNCH_CHECK(); ++launches;
				bl_march_cull_kernel<<<blocks, 128, 0, stream>>>(counters + cur, n_nerfs, props_dev, n_init, R.cam[9], R.cam[10], R.cam[11], rays[cur], prox[cur]);
				NGPB_LAUNCH_CHECK(); ++launches;
				// n_alive * n_steps <= max(n_initial, ...) by construction of n_steps; with the bound: min(n_initial, bound * max_steps) slots at most
				const uint32_t slot_bound = next_multiple((uint32_t)std::min<uint64_t>((uint64_t)bound * max_steps, std::max<uint64_t>(n_init, bound)), 128);
				for (uint32_t n = 0; n < n_nerfs; ++n) {
					BlProxyRay* pn = prox[cur] + (size_t)n * n_init;
					if (warp_march) bl_generate_inputs_warp_kernel<<<div_round_up(bound, 8u), 256, 0, stream>>>(counters + cur, wave, props_dev + n, rays[cur], pn, coords);
					else bl_generate_inputs_kernel<<<blocks, 128, 0, stream>>>(counters + cur, wave, props_dev + n, rays[cur], pn, coords);
					NGPB_LAUNCH_CHECK();
					const ngpb_field* f = nerfs[n].field;
					hash_encode_forward_launch(stream, &f->grid, f->params + MLP_PARAMS, coords, COORD_FLOATS, slot_bound, &wave->n_slots, encoded, features_tiled());
					nerf_mlp_forward_launch(stream, f->params, encoded, features_tiled(), coords, slot_bound, &wave->n_slots, rgbsigma);
					bl_composite_kernel<<<blocks, 128, 0, stream>>>(counters + cur, wave, props_dev + n, rays[cur], pn, coords, rgbsigma, reinterpret_cast<unsigned long long*>(counters + 2));
					NGPB_LAUNCH_CHECK();
					launches += 4;
				}
			}
		}
		// Testbed::bl_render_frame: accumulate (first sample) + tonemap, buffer and output colour space = the request's
		render_accumulate_kernel<<<div_round_up(n_pixels, 256), 256, 0, stream>>>(n_pixels, frame, accum, 0.0f, rq->color_space);
		render_tonemap_kernel<<<div_round_up(n_pixels, 256), 256, 0, stream>>>(n_pixels, powf(2.0f, rq->exposure),
			make_float4(rq->background_color[0], rq->background_color[1], rq->background_color[2], rq->background_color[3]), accum, rq->color_space,
			rq->color_space == NGPB_COLOR_SRGB ? 1 : 0, rq->tonemap_curve, out);
		NGPB_LAUNCH_CHECK(); launches += 2;
		const bool direct = host_pointer_is_pinned(out_rgba_host);
		NGPB_CUDA_CHECK(cudaMemcpyAsync(direct ? out_rgba_host : host_frame, out, (size_t)n_pixels * 16, cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaMemcpyAsync(host_counter, counters + 2, 8, cudaMemcpyDeviceToHost, stream));
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		if (!direct) parallel_host_copy(out_rgba_host, host_frame, (size_t)n_pixels * 16);
		if (n_samples_out) std::memcpy(n_samples_out, host_counter, 8);
		if (n_launches_out) *n_launches_out = launches;
		return 0;
	} catch (const std::exception& e) {
		set_last_error(e.what());
		return NGPB_ERR_RUNTIME;
	}
}
