// K16: occupancy (density) grid maintenance.
// Replaces mark_untrained_density_grid, generate_grid_samples_nerf_nonuniform,
// splat_grid_samples_nerf_max_nearest_neighbor, ema_grid_samples_nerf, grid_to_bitfield,
// bitfield_max_pool and the reduce_sum mean (reference: src/testbed_nerf.cu:369-610, :2844-2859).
#include "nerf_device.cuh"

namespace ngpb {

__global__ void __launch_bounds__(128) mark_untrained_kernel(const uint32_t n_elements, float* __restrict__ grid_out, const uint32_t n_images,
                                                             const ngpb_image* __restrict__ images, const bool clear_visible_voxels)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint32_t level = i / NERF_GRID_CELLS, pos_idx = i % NERF_GRID_CELLS;
	const uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
	const float s = scalbnf(1.0f, (int)level);
	const float pos[3] = {
		(((float)x + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
		(((float)y + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
		(((float)z + 0.5f) / NERF_GRIDSIZE - 0.5f) * s + 0.5f};
	const float voxel_radius = 0.5f * SQRT3 * s / NERF_GRIDSIZE;
	int count = 0;
	for (uint32_t j = 0; j < n_images; ++j) {
		const ngpb_image& im = images[j];
		if (im.lens_mode == NGPB_LENS_FTHETA || im.lens_mode == NGPB_LENS_LATLONG) { count++; break; } // "not supported for now": such a camera sees every cell (:391-395)
		const float half_resx = im.w * 0.5f, half_resy = im.h * 0.5f;
		const float* xf = im.raw_xform;
		const float ploc[3] = {pos[0] - xf[9], pos[1] - xf[10], pos[2] - xf[11]};
		const float cx = dot3(ploc, xf + 0), cy = dot3(ploc, xf + 3), cz = dot3(ploc, xf + 6);
		if (cz > 0.f) {
			if (fabsf(cx) - voxel_radius < cz / im.fx * half_resx && fabsf(cy) - voxel_radius < cz / im.fy * half_resy) { count++; break; }
		}
	}
	if (clear_visible_voxels || (grid_out[i] < 0) != (count <= 0)) grid_out[i] = (count > 0) ? 0.f : -1.f;
}

__global__ void __launch_bounds__(128) generate_grid_samples_kernel(const uint32_t n_elements, Pcg32 rng, const uint32_t step, const Aabb aabb,
                                                                    const float* __restrict__ grid_in, float* __restrict__ out, uint32_t* __restrict__ indices,
                                                                    const uint32_t n_cascades, const float thresh)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	rng.advance((int64_t)i * 4); // 1 random number to select the level, 3 to select the position (:469-471)
	const uint32_t level = (uint32_t)(rng.next_float() * n_cascades) % n_cascades;
	uint32_t idx = 0;
	for (uint32_t j = 0; j < 10; ++j) {
		idx = ((i + step * n_elements) * 56924617u + j * 19349663u + 96925573u) % NERF_GRID_CELLS;
		idx += level * NERF_GRID_CELLS;
		if (grid_in[idx] > thresh) break;
	}
	const uint32_t pos_idx = idx % NERF_GRID_CELLS;
	const uint32_t x = morton3D_invert(pos_idx >> 0), y = morton3D_invert(pos_idx >> 1), z = morton3D_invert(pos_idx >> 2);
	const float rx = rng.next_float(), ry = rng.next_float(), rz = rng.next_float();
	const float s = scalbnf(1.0f, (int)level);
	const V3 pos = {
		(((float)x + rx) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
		(((float)y + ry) / NERF_GRIDSIZE - 0.5f) * s + 0.5f,
		(((float)z + rz) / NERF_GRIDSIZE - 0.5f) * s + 0.5f};
	const V3 wp = warp_position(pos, aabb);
	out[(size_t)i * 3 + 0] = wp.x; out[(size_t)i * 3 + 1] = wp.y; out[(size_t)i * 3 + 2] = wp.z;
	indices[i] = idx;
}

__global__ void __launch_bounds__(128) splat_kernel(const uint32_t n, const uint32_t* __restrict__ indices, const __half* __restrict__ density, float* __restrict__ grid_tmp)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float mlp = __expf(__half2float(density[i])); // network_to_density, Exponential (:506)
	const float optical_thickness = mlp * scalbnf(MIN_CONE_STEPSIZE, 0);
	atomicMax(reinterpret_cast<uint32_t*>(&grid_tmp[indices[i]]), __float_as_uint(optical_thickness)); // (:509-511)
}

__global__ void __launch_bounds__(256) ema_kernel(const uint32_t n, const float decay, float* __restrict__ grid, const float* __restrict__ grid_tmp)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float prev_val = grid[i];
	grid[i] = (prev_val < 0.f) ? prev_val : fmaxf(prev_val * decay, grid_tmp[i]); // (:552-554)
}

// mean of max(v, 0) / n over the first cascade (:2851-2852). Deterministic two-stage tree:
// 1024 block partials in double, then one block folds them in a fixed order.
__global__ void __launch_bounds__(256) mean_partial_kernel(const float* __restrict__ grid, double* __restrict__ partials)
{
	__shared__ double sm[256];
	double s = 0.0;
	const uint32_t per_block = NERF_GRID_CELLS / 1024;
	const uint32_t b0 = blockIdx.x * per_block;
	for (uint32_t k = threadIdx.x; k < per_block; k += 256) s += (double)(fmaxf(grid[b0 + k], 0.f) / (float)NERF_GRID_CELLS);
	sm[threadIdx.x] = s;
	__syncthreads();
	for (uint32_t o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(1024) mean_final_kernel(const double* __restrict__ partials, float* __restrict__ mean_out)
{
	__shared__ double sm[1024];
	sm[threadIdx.x] = partials[threadIdx.x];
	__syncthreads();
	for (uint32_t o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) *mean_out = (float)sm[0];
}

__global__ void __launch_bounds__(256) grid_to_bitfield_kernel(const uint32_t n_elements, const uint32_t n_nonzero_elements, const float* __restrict__ grid,
                                                               uint8_t* __restrict__ bitfield, const float* __restrict__ mean_density_ptr)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	if (i >= n_nonzero_elements) { bitfield[i] = 0; return; }
	const float thresh = fminf(NERF_MIN_OPTICAL_THICKNESS, *mean_density_ptr);
	const float4 a = reinterpret_cast<const float4*>(grid)[(size_t)i * 2], b = reinterpret_cast<const float4*>(grid)[(size_t)i * 2 + 1];
	uint8_t bits = 0;
	bits |= a.x > thresh ? 1 : 0; bits |= a.y > thresh ? 2 : 0; bits |= a.z > thresh ? 4 : 0; bits |= a.w > thresh ? 8 : 0;
	bits |= b.x > thresh ? 16 : 0; bits |= b.y > thresh ? 32 : 0; bits |= b.z > thresh ? 64 : 0; bits |= b.w > thresh ? 128 : 0;
	bitfield[i] = bits;
}

__global__ void __launch_bounds__(256) bitfield_max_pool_kernel(const uint32_t n_elements, const uint8_t* __restrict__ prev_level, uint8_t* __restrict__ next_level)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_elements) return;
	const uint2 raw = reinterpret_cast<const uint2*>(prev_level)[i];
	uint8_t bits = 0;
	#pragma unroll
	for (int j = 0; j < 4; ++j) {
		bits |= ((raw.x >> (8 * j)) & 0xFF) ? (uint8_t)(1 << j) : 0;
		bits |= ((raw.y >> (8 * j)) & 0xFF) ? (uint8_t)(16 << j) : 0;
	}
	const uint32_t x = morton3D_invert(i >> 0) + NERF_GRIDSIZE / 8, y = morton3D_invert(i >> 1) + NERF_GRIDSIZE / 8, z = morton3D_invert(i >> 2) + NERF_GRIDSIZE / 8;
	next_level[morton3D(x, y, z)] |= bits;
}

Aabb make_aabb(const float* a);

} // namespace ngpb

using namespace ngpb;

extern "C" int ngpb_mark_untrained_density_grid(void* stream, uint32_t n_elements, float* grid, uint32_t n_images, const ngpb_image* images_dev, int clear_visible) {
	try {
		if (!grid || !images_dev) { set_last_error("ngpb_mark_untrained_density_grid: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		mark_untrained_kernel<<<div_round_up(n_elements, 128), 128, 0, (cudaStream_t)stream>>>(n_elements, grid, n_images, images_dev, clear_visible != 0);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_generate_grid_samples(void* stream, uint32_t n_elements, ngpb_rng rng_, uint32_t step, const float* aabb6, const float* grid_in,
                                          float* positions3, uint32_t* indices, uint32_t n_cascades, float thresh) {
	try {
		if (!aabb6 || !grid_in || !positions3 || !indices || n_cascades == 0) { set_last_error("ngpb_generate_grid_samples: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (n_elements == 0) return 0;
		Pcg32 rng; rng.state = rng_.state; rng.inc = rng_.inc;
		generate_grid_samples_kernel<<<div_round_up(n_elements, 128), 128, 0, (cudaStream_t)stream>>>(n_elements, rng, step, make_aabb(aabb6), grid_in, positions3, indices, n_cascades, thresh);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_splat_and_ema(void* stream_, uint32_t n_samples, const uint32_t* indices, const ngpb_half* density, float* grid_tmp, uint32_t n_elements, float decay, float* grid) {
	try {
		if (!indices || !density || !grid_tmp || !grid) { set_last_error("ngpb_splat_and_ema: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		cudaStream_t stream = (cudaStream_t)stream_;
		NGPB_CUDA_CHECK(cudaMemsetAsync(grid_tmp, 0, sizeof(float) * n_elements, stream));
		if (n_samples) splat_kernel<<<div_round_up(n_samples, 128), 128, 0, stream>>>(n_samples, indices, (const __half*)density, grid_tmp);
		NGPB_LAUNCH_CHECK();
		ema_kernel<<<div_round_up(n_elements, 256), 256, 0, stream>>>(n_elements, decay, grid, grid_tmp);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_update_bitfield(void* stream_, uint32_t n_cascades_used, const float* grid, float* mean_dev, uint8_t* bitfield) {
	try {
		if (!grid || !mean_dev || !bitfield || n_cascades_used == 0 || n_cascades_used > NERF_CASCADES) { set_last_error("ngpb_update_bitfield: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		cudaStream_t stream = (cudaStream_t)stream_;
		// per-owner scratch (no process-wide buffer: testbeds and fields live on different devices and streams): the 1024 block partials follow the mean
		double* partials = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(mean_dev) + 8);
		mean_partial_kernel<<<1024, 256, 0, stream>>>(grid, partials);
		NGPB_LAUNCH_CHECK();
		mean_final_kernel<<<1, 1024, 0, stream>>>(partials, mean_dev);
		NGPB_LAUNCH_CHECK();
		const uint32_t n_bytes = NERF_GRID_CELLS / 8 * NERF_CASCADES;
		grid_to_bitfield_kernel<<<div_round_up(n_bytes, 256), 256, 0, stream>>>(n_bytes, NERF_GRID_CELLS / 8 * n_cascades_used, grid, bitfield, mean_dev);
		NGPB_LAUNCH_CHECK();
		for (uint32_t level = 1; level < NERF_CASCADES; ++level) {
			bitfield_max_pool_kernel<<<div_round_up(NERF_GRID_CELLS / 64, 256), 256, 0, stream>>>(NERF_GRID_CELLS / 64,
				bitfield + (size_t)NERF_GRID_CELLS * (level - 1) / 8, bitfield + (size_t)NERF_GRID_CELLS * level / 8);
			NGPB_LAUNCH_CHECK();
		}
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
