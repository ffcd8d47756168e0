// K1: training ray generation + occupancy-grid marching with deterministic compaction.
// Replaces generate_training_samples_nerf (reference: src/testbed_nerf.cu:1085-1260).
//
// The reference allocates sample ranges and ray slots with atomicAdd (:1225,:1232), so the slot
// ORDER differs from run to run while the SET of (ray, numsteps, samples) is fixed. Here the same
// allocation is done by an exclusive scan in ray-index order: count pass -> scan -> write pass.
// That is one valid serialisation of the reference's atomics, reproducible bit for bit, and it
// leaves the samples of consecutive rays adjacent in memory (rays of one image are consecutive,
// :1078-1082), which the hash-grid kernels downstream turn into cache hits.
#include "nerf_device.cuh"

namespace ngpb {

struct TrainRay { V3 o, d_unnorm, d; float startt, cone_angle; bool valid; };

// Ray set-up: testbed_nerf.cu:1118-1202 (perspective lens, no rolling shutter, no distortion map,
// uniform pixel sampling). RNG draws in the reference's order: xy(2), motion-blur time(1), start jitter(1).
__device__ inline TrainRay setup_training_ray(uint32_t i, uint32_t n_rays, Pcg32 rng, uint32_t n_images, const ngpb_image* __restrict__ images,
                                              const Aabb& aabb, bool snap, float cone_angle_constant) {
	TrainRay r;
	r.valid = false;
	const uint32_t img = image_idx(i, n_rays, n_images);
	const ngpb_image& im = images[img];
	rng.advance((int64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
	float x, y;
	random_image_pos_training(rng, im.w, im.h, snap, &x, &y);
	float px[4];
	if (!read_rgba(x, y, im, px)) return r; // masked-away pixel (:1126)
	(void)rng.next_float(); // motionblur_time (:1132); max_level_rand_training is off so no draw at :1130
	const float* xf = im.xform;
	r.o = {xf[9], xf[10], xf[11]};
	const float dcam[3] = {(x - im.cx) * (float)im.w / im.fx, (y - im.cy) * (float)im.h / im.fy, 1.0f};
	const float row0[3] = {xf[0], xf[3], xf[6]}, row1[3] = {xf[1], xf[4], xf[7]}, row2[3] = {xf[2], xf[5], xf[8]};
	r.d_unnorm = {dot3(row0, dcam), dot3(row1, dcam), dot3(row2, dcam)};
	const float z = sum3(r.d_unnorm.x * r.d_unnorm.x, r.d_unnorm.y * r.d_unnorm.y, r.d_unnorm.z * r.d_unnorm.z);
	if (z > 0.f) { const float nrm = sqrtf(z); r.d = {r.d_unnorm.x / nrm, r.d_unnorm.y / nrm, r.d_unnorm.z / nrm}; } else r.d = r.d_unnorm;
	float tmin, tmax;
	aabb_ray_intersect(aabb, r.o, r.d, &tmin, &tmax);
	r.cone_angle = cone_angle_constant; // calc_cone_angle (:87-94)
	tmin = fmaxf(tmin, 0.0f);
	float startt = tmin;
	startt += calc_dt(startt, r.cone_angle) * rng.next_float();
	r.startt = startt;
	r.valid = true;
	return r;
}

// March (testbed_nerf.cu:1204-1219 count pass, :1239-1253 write pass). WRITE selects the pass.
template <bool WRITE>
__device__ inline uint32_t march_training_ray(const TrainRay& r, const Aabb& aabb, const uint8_t* __restrict__ bitfield, uint32_t max_steps, float* __restrict__ coords_out) {
	const V3 idir = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
	V3 wd;
	if (WRITE) wd = {(r.d.x + 1.0f) * 0.5f, (r.d.y + 1.0f) * 0.5f, (r.d.z + 1.0f) * 0.5f}; // warp_direction (:292)
	uint32_t j = 0;
	float t = r.startt;
	V3 pos;
	while (aabb_contains(aabb, pos = V3{r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z}) && j < max_steps) {
		const float dt = calc_dt(t, r.cone_angle);
		const uint32_t mip = mip_from_dt(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) {
			if (WRITE) {
				const V3 wp = warp_position(pos, aabb);
				float* c = coords_out + (size_t)j * COORD_FLOATS;
				c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
			}
			++j;
			t += dt;
		} else {
			const uint32_t res = NERF_GRIDSIZE >> mip;
			t = advance_to_next_voxel(t, r.cone_angle, pos, r.d, idir, res);
		}
	}
	return j;
}

__global__ void __launch_bounds__(128) count_training_samples_kernel(
	const uint32_t n_rays, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const uint8_t* __restrict__ bitfield, const bool snap, const float cone_angle_constant, uint32_t* __restrict__ counts)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rays) return;
	const TrainRay r = setup_training_ray(i, n_rays, rng, n_images, images, aabb, snap, cone_angle_constant);
	counts[i] = r.valid ? march_training_ray<false>(r, aabb, bitfield, NERF_STEPS, nullptr) : 0u;
}

// One block of 1024 threads. bases[i] = exclusive prefix of counts (the reference's numsteps_counter
// atomicAdd, :1225); a ray is kept iff count > 0 and base + count <= max_samples (:1221-1228); kept rays
// get consecutive slots (the ray_counter atomicAdd, :1232). slots[i] = slot or 0xFFFFFFFF.
__global__ void __launch_bounds__(1024) scan_training_samples_kernel(
	const uint32_t n_rays, const uint32_t max_samples, const uint32_t* __restrict__ counts, uint32_t* __restrict__ bases, uint32_t* __restrict__ slots,
	uint32_t* __restrict__ counters)
{
	__shared__ uint32_t smem[33];
	const uint32_t per_thread = (n_rays + 1023) / 1024;
	const uint32_t begin = min(threadIdx.x * per_thread, n_rays), end = min(begin + per_thread, n_rays);
	uint32_t sum = 0;
	for (uint32_t i = begin; i < end; ++i) sum += counts[i];
	uint32_t total;
	uint32_t base = block_exclusive_scan_1024(sum, smem, &total);
	uint32_t kept = 0;
	for (uint32_t i = begin; i < end; ++i) {
		const uint32_t c = counts[i];
		bases[i] = base;
		kept += (c > 0 && base + c <= max_samples) ? 1u : 0u;
		base += c;
	}
	uint32_t total_kept;
	uint32_t slot = block_exclusive_scan_1024(kept, smem, &total_kept);
	for (uint32_t i = begin; i < end; ++i) {
		const uint32_t c = counts[i];
		const bool k = c > 0 && bases[i] + c <= max_samples;
		slots[i] = k ? slot : 0xFFFFFFFFu;
		slot += k ? 1u : 0u;
	}
	if (threadIdx.x == 0) { counters[0] = total; counters[1] = total_kept; }
}

__global__ void __launch_bounds__(128) write_training_samples_kernel(
	const uint32_t n_rays, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const uint8_t* __restrict__ bitfield, const bool snap, const float cone_angle_constant,
	const uint32_t* __restrict__ counts, const uint32_t* __restrict__ bases, const uint32_t* __restrict__ slots,
	uint32_t* __restrict__ ray_indices, float* __restrict__ rays, uint32_t* __restrict__ numsteps, float* __restrict__ coords)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_rays) return;
	const uint32_t slot = slots[i];
	if (slot == 0xFFFFFFFFu) return;
	const uint32_t base = bases[i], count = counts[i];
	const TrainRay r = setup_training_ray(i, n_rays, rng, n_images, images, aabb, snap, cone_angle_constant);
	ray_indices[slot] = i;
	float* ro = rays + (size_t)slot * 6;
	ro[0] = r.o.x; ro[1] = r.o.y; ro[2] = r.o.z; ro[3] = r.d_unnorm.x; ro[4] = r.d_unnorm.y; ro[5] = r.d_unnorm.z;
	numsteps[slot * 2 + 0] = count;
	numsteps[slot * 2 + 1] = base;
	march_training_ray<true>(r, aabb, bitfield, count, coords + (size_t)base * COORD_FLOATS);
}

Aabb make_aabb(const float* a) {
	Aabb b;
	for (int c = 0; c < 3; ++c) { b.min[c] = a[c]; b.max[c] = a[3 + c]; }
	return b;
}

} // namespace ngpb

using namespace ngpb;

extern "C" int ngpb_generate_training_samples(void* stream_, uint32_t n_rays, const float* aabb6, uint32_t max_samples, ngpb_rng rng_,
                                              uint32_t n_images, const ngpb_image* images_dev, const uint8_t* bitfield,
                                              int snap_to_pixel_centers, float cone_angle_constant,
                                              uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, uint32_t* scratch) {
	try {
		if (!aabb6 || !images_dev || !bitfield || !counters || !ray_indices || !rays || !numsteps || !coords || !scratch || n_images == 0) {
			set_last_error("ngpb_generate_training_samples: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		if (n_rays == 0) { NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), stream)); return 0; }
		const Aabb aabb = make_aabb(aabb6);
		Pcg32 rng; rng.state = rng_.state; rng.inc = rng_.inc;
		uint32_t* counts = scratch;
		uint32_t* bases = scratch + n_rays;
		uint32_t* slots = scratch + 2 * (size_t)n_rays;
		const uint32_t blocks = div_round_up(n_rays, 128);
		count_training_samples_kernel<<<blocks, 128, 0, stream>>>(n_rays, aabb, rng, n_images, images_dev, bitfield, snap_to_pixel_centers != 0, cone_angle_constant, counts);
		NGPB_LAUNCH_CHECK();
		scan_training_samples_kernel<<<1, 1024, 0, stream>>>(n_rays, max_samples, counts, bases, slots, counters);
		NGPB_LAUNCH_CHECK();
		write_training_samples_kernel<<<blocks, 128, 0, stream>>>(n_rays, aabb, rng, n_images, images_dev, bitfield, snap_to_pixel_centers != 0, cone_angle_constant,
			counts, bases, slots, ray_indices, rays, numsteps, coords);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
