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

// ---- marching ----------------------------------------------------------------------------------------
// Reference: testbed_nerf.cu:1204-1219 (count pass) and :1239-1253 (write pass): the ray is marched twice.
// Both branches of the march advance t the same way -- `t += calc_dt(t, cone_angle)`, once per sample in an occupied cell,
// repeatedly until the cell's exit in an empty one (advance_to_next_voxel, :449-463) -- so the ray's t values form ONE chain
// t_0, t_1, ... that does not depend on the occupancy grid; the grid only selects which t_k become samples. The march is
// therefore done once: it records, per 32 chain steps, a bit mask of the emitted samples and the t at the first step
// (MarchWord). The write pass is then parallel over the words of a ray: every lane replays 32 steps of the chain from its
// word's t (the same float additions) and writes the samples whose bits are set.
struct MarchWord { uint32_t mask; float t; };
constexpr uint32_t MARCH_MAX_WORDS = 72;         // 2304 chain steps; a unit-cube ray takes <= 1025, aabb_scale 128 about 2100
constexpr uint32_t MARCH_OVERFLOW = 0x80000000u; // flag in n_words: the chain outgrew the record, the write pass re-marches this ray

// Serial march of one ray. Returns the number of samples; fills words[0 .. *n_words) unless the chain overflows.
__device__ inline uint32_t march_and_record(const TrainRay& r, const Aabb& aabb, const uint8_t* __restrict__ bitfield, MarchWord* __restrict__ words, uint32_t* n_words_out) {
	const V3 idir = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
	uint32_t j = 0, k = 0, mask = 0, n_words = 0;
	bool overflow = false;
	float t = r.startt, word_t = r.startt;
	V3 pos;
	// one chain step: t_{k+1} = t_k + calc_dt(t_k); closes the current word every 32 steps
	#define NGPB_CHAIN_STEP()                                                             \
		do {                                                                              \
			t += calc_dt(t, r.cone_angle);                                                \
			if ((++k & 31u) == 0) {                                                       \
				if (n_words < MARCH_MAX_WORDS) { words[n_words].mask = mask; words[n_words].t = word_t; } else overflow = true; \
				++n_words; mask = 0; word_t = t;                                          \
			}                                                                             \
		} while (0)
	while (aabb_contains(aabb, pos = V3{r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z}) && j < NERF_STEPS) {
		const float dt = calc_dt(t, r.cone_angle);
		const uint32_t mip = mip_from_dt(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) {
			mask |= 1u << (k & 31u);
			++j;
			NGPB_CHAIN_STEP();
		} else {
			const uint32_t res = NERF_GRIDSIZE >> mip;
			const float t_target = t + distance_to_next_voxel(pos, r.d, idir, res);
			do { NGPB_CHAIN_STEP(); } while (t < t_target);
		}
	}
	#undef NGPB_CHAIN_STEP
	if (mask) {
		if (n_words < MARCH_MAX_WORDS) { words[n_words].mask = mask; words[n_words].t = word_t; } else overflow = true;
		++n_words;
	}
	*n_words_out = overflow ? MARCH_OVERFLOW : n_words;
	return j;
}

// Re-march that writes directly (the reference's second pass); only used for rays whose chain overflowed the record.
__device__ inline void march_and_write(const TrainRay& r, const Aabb& aabb, const uint8_t* __restrict__ bitfield, uint32_t max_steps, float* __restrict__ coords_out) {
	const V3 idir = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
	const V3 wd = {(r.d.x + 1.0f) * 0.5f, (r.d.y + 1.0f) * 0.5f, (r.d.z + 1.0f) * 0.5f}; // warp_direction (:292)
	uint32_t j = 0;
	float t = r.startt;
	V3 pos;
	while (aabb_contains(aabb, pos = V3{r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z}) && j < max_steps) {
		const float dt = calc_dt(t, r.cone_angle);
		const uint32_t mip = mip_from_dt(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) {
			const V3 wp = warp_position(pos, aabb);
			float* c = coords_out + (size_t)j * COORD_FLOATS;
			c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
			++j;
			t += dt;
		} else {
			const uint32_t res = NERF_GRIDSIZE >> mip;
			t = advance_to_next_voxel(t, r.cone_angle, pos, r.d, idir, res);
		}
	}
}

constexpr uint32_t K1_BLOCK = 128;

// Pass 1: one thread per ray. Sample count, march record, and the block-local exclusive prefixes of (count, count > 0).
__global__ void __launch_bounds__(K1_BLOCK) count_training_samples_kernel(
	const uint32_t n_rays, const uint32_t ray_offset, const uint32_t n_rays_global, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const uint8_t* __restrict__ bitfield, const bool snap, const float cone_angle_constant,
	uint32_t* __restrict__ counts, uint32_t* __restrict__ n_words, MarchWord* __restrict__ words, uint32_t* __restrict__ local_bases, uint32_t* __restrict__ local_slots,
	uint2* __restrict__ block_sums)
{
	__shared__ uint32_t warp_sums[2][K1_BLOCK / 32];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t c = 0;
	if (i < n_rays) {
		// everything about a ray derives from its GLOBAL index (:1118-1121, :1062-1083): a shard reproduces its slice of the full batch
		const TrainRay r = setup_training_ray(ray_offset + i, n_rays_global, rng, n_images, images, aabb, snap, cone_angle_constant);
		uint32_t nw = 0;
		if (r.valid) c = march_and_record(r, aabb, bitfield, words + (size_t)i * MARCH_MAX_WORDS, &nw);
		counts[i] = c;
		n_words[i] = nw;
	}
	// block-local exclusive scans in ray order
	uint32_t incl_c = c, incl_z = c > 0 ? 1u : 0u;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t tc = __shfl_up_sync(0xffffffffu, incl_c, o), tz = __shfl_up_sync(0xffffffffu, incl_z, o);
		if (lane >= (uint32_t)o) { incl_c += tc; incl_z += tz; }
	}
	if (lane == 31) { warp_sums[0][warp] = incl_c; warp_sums[1][warp] = incl_z; }
	__syncthreads();
	uint32_t off_c = 0, off_z = 0, tot_c = 0, tot_z = 0;
	#pragma unroll
	for (uint32_t w = 0; w < K1_BLOCK / 32; ++w) {
		if (w < warp) { off_c += warp_sums[0][w]; off_z += warp_sums[1][w]; }
		tot_c += warp_sums[0][w]; tot_z += warp_sums[1][w];
	}
	if (i < n_rays) { local_bases[i] = off_c + incl_c - c; local_slots[i] = off_z + incl_z - (c > 0 ? 1u : 0u); }
	if (threadIdx.x == 0) block_sums[blockIdx.x] = make_uint2(tot_c, tot_z);
}

// Pass 2: one block. Exclusive scan of the per-block totals: sample bases (the reference's numsteps_counter atomicAdd, :1225)
// and ray slots (the ray_counter atomicAdd, :1232). counters[0] = samples requested; counters[1] is zeroed and counted by pass 3.
__global__ void __launch_bounds__(1024) scan_training_samples_kernel(const uint32_t n_blocks, uint2* __restrict__ block_sums, uint32_t* __restrict__ counters)
{
	__shared__ uint32_t smem[33];
	uint32_t carry_c = 0, carry_z = 0;
	for (uint32_t t0 = 0; t0 < n_blocks; t0 += 1024) {
		const uint32_t b = t0 + threadIdx.x;
		const uint2 v = b < n_blocks ? block_sums[b] : make_uint2(0u, 0u);
		uint32_t tot_c, tot_z;
		const uint32_t ex_c = block_exclusive_scan_1024(v.x, smem, &tot_c);
		const uint32_t ex_z = block_exclusive_scan_1024(v.y, smem, &tot_z);
		if (b < n_blocks) block_sums[b] = make_uint2(carry_c + ex_c, carry_z + ex_z);
		carry_c += tot_c; carry_z += tot_z;
	}
	if (threadIdx.x == 0) { counters[0] = carry_c; counters[1] = 0; }
}

// Pass 3: one warp per ray, one lane per march word. A ray is kept iff it has samples and base + count <= max_samples
// (:1221-1228); since bases grow with the ray index the kept rays are exactly the sample-bearing rays before the cut, so a kept
// ray's slot is the number of sample-bearing rays before it.
constexpr uint32_t WRITE_RAYS_PER_BLOCK = 8;
__global__ void __launch_bounds__(WRITE_RAYS_PER_BLOCK * 32) write_training_samples_kernel(
	const uint32_t n_rays, const uint32_t ray_offset, const uint32_t n_rays_global, const uint32_t max_samples, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const uint8_t* __restrict__ bitfield, const bool snap, const float cone_angle_constant,
	const uint32_t* __restrict__ counts, const uint32_t* __restrict__ n_words, const MarchWord* __restrict__ words,
	const uint32_t* __restrict__ local_bases, const uint32_t* __restrict__ local_slots, const uint2* __restrict__ block_prefix,
	uint32_t* __restrict__ counters, uint32_t* __restrict__ ray_indices, float* __restrict__ rays, uint32_t* __restrict__ numsteps, float* __restrict__ coords)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t i = blockIdx.x * WRITE_RAYS_PER_BLOCK + warp;
	bool kept = false;
	uint32_t count = 0, base = 0, slot = 0;
	if (i < n_rays) {
		count = counts[i];
		const uint2 bp = block_prefix[i / K1_BLOCK];
		base = bp.x + local_bases[i];
		slot = bp.y + local_slots[i];
		kept = count > 0 && base + count <= max_samples;
	}
	const uint32_t n_kept_block = __syncthreads_count(kept && lane == 0);
	if (threadIdx.x == 0 && n_kept_block) atomicAdd(&counters[1], n_kept_block);
	if (!kept) return;

	const TrainRay r = setup_training_ray(ray_offset + i, n_rays_global, rng, n_images, images, aabb, snap, cone_angle_constant); // same on every lane
	if (lane == 0) {
		ray_indices[slot] = ray_offset + i; // global ray index: K6 re-derives pixel and background from it
		float* ro = rays + (size_t)slot * 6;
		ro[0] = r.o.x; ro[1] = r.o.y; ro[2] = r.o.z; ro[3] = r.d_unnorm.x; ro[4] = r.d_unnorm.y; ro[5] = r.d_unnorm.z;
		numsteps[slot * 2 + 0] = count;
		numsteps[slot * 2 + 1] = base;
	}
	float* out = coords + (size_t)base * COORD_FLOATS;
	const uint32_t nw = n_words[i];
	if (nw & MARCH_OVERFLOW) {
		if (lane == 0) march_and_write(r, aabb, bitfield, count, out);
		return;
	}
	const V3 wd = {(r.d.x + 1.0f) * 0.5f, (r.d.y + 1.0f) * 0.5f, (r.d.z + 1.0f) * 0.5f}; // warp_direction (:292)
	const MarchWord* rw = words + (size_t)i * MARCH_MAX_WORDS;
	uint32_t carry = 0;
	for (uint32_t w0 = 0; w0 < nw; w0 += 32) {
		const uint32_t w = w0 + lane;
		MarchWord mw = {0u, 0.f};
		if (w < nw) mw = rw[w];
		const uint32_t n_here = __popc(mw.mask);
		uint32_t incl = n_here;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint32_t tt = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += tt; }
		uint32_t j = carry + incl - n_here;
		carry += __shfl_sync(0xffffffffu, incl, 31);
		float t = mw.t;
		uint32_t m = mw.mask;
		// replay this word's 32 chain steps; stop after the last set bit
		while (m) {
			const float dt = calc_dt(t, r.cone_angle);
			if (m & 1u) {
				const V3 pos = V3{r.o.x + t * r.d.x, r.o.y + t * r.d.y, r.o.z + t * r.d.z};
				const V3 wp = warp_position(pos, aabb);
				float* c = out + (size_t)j * COORD_FLOATS;
				c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
				++j;
			}
			t += dt;
			m >>= 1;
		}
	}
}

Aabb make_aabb(const float* a) {
	Aabb b;
	for (int c = 0; c < 3; ++c) { b.min[c] = a[c]; b.max[c] = a[3 + c]; }
	return b;
}

} // namespace ngpb

using namespace ngpb;

extern "C" uint64_t ngpb_generate_training_samples_scratch_bytes(uint32_t n_rays) {
	return (uint64_t)n_rays * 16 + 8 + (uint64_t)div_round_up(n_rays, K1_BLOCK) * sizeof(uint2) + (uint64_t)n_rays * MARCH_MAX_WORDS * sizeof(MarchWord) + 64;
}

extern "C" int ngpb_generate_training_samples(void* stream, uint32_t n_rays, const float* aabb6, uint32_t max_samples, ngpb_rng rng,
                                              uint32_t n_images, const ngpb_image* images_dev, const uint8_t* bitfield,
                                              int snap_to_pixel_centers, float cone_angle_constant,
                                              uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch) {
	return ngpb_generate_training_samples_sharded(stream, n_rays, 0, n_rays, aabb6, max_samples, rng, n_images, images_dev, bitfield, snap_to_pixel_centers, cone_angle_constant,
		counters, ray_indices, rays, numsteps, coords, scratch);
}

extern "C" int ngpb_generate_training_samples_sharded(void* stream_, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, ngpb_rng rng_,
                                              uint32_t n_images, const ngpb_image* images_dev, const uint8_t* bitfield,
                                              int snap_to_pixel_centers, float cone_angle_constant,
                                              uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch_) {
	try {
		uint32_t* scratch = reinterpret_cast<uint32_t*>(scratch_);
		if (!aabb6 || !images_dev || !bitfield || !counters || !ray_indices || !rays || !numsteps || !coords || !scratch || n_images == 0 ||
		    (uint64_t)ray_offset + n_rays > n_rays_global) {
			set_last_error("ngpb_generate_training_samples: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		if (n_rays == 0) { NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), stream)); return 0; }
		const Aabb aabb = make_aabb(aabb6);
		Pcg32 rng; rng.state = rng_.state; rng.inc = rng_.inc;
		// scratch layout: counts | n_words | local_bases | local_slots (uint32[n_rays] each) | block_sums (uint2[n_blocks]) | MarchWord[n_rays][MARCH_MAX_WORDS]
		uint32_t* counts = scratch;
		uint32_t* n_words = counts + n_rays;
		uint32_t* local_bases = n_words + n_rays;
		uint32_t* local_slots = local_bases + n_rays;
		const uint32_t blocks = div_round_up(n_rays, K1_BLOCK);
		uint2* block_sums = reinterpret_cast<uint2*>(local_slots + next_multiple(n_rays, 2));
		MarchWord* words = reinterpret_cast<MarchWord*>(block_sums + blocks);
		NGPB_STEP_KERNEL(count_training_samples_kernel); NGPB_STEP_KERNEL(scan_training_samples_kernel); NGPB_STEP_KERNEL(write_training_samples_kernel);
		count_training_samples_kernel<<<blocks, K1_BLOCK, 0, stream>>>(n_rays, ray_offset, n_rays_global, aabb, rng, n_images, images_dev, bitfield, snap_to_pixel_centers != 0, cone_angle_constant,
			counts, n_words, words, local_bases, local_slots, block_sums);
		NGPB_LAUNCH_CHECK();
		scan_training_samples_kernel<<<1, 1024, 0, stream>>>(blocks, block_sums, counters);
		NGPB_LAUNCH_CHECK();
		write_training_samples_kernel<<<div_round_up(n_rays, WRITE_RAYS_PER_BLOCK), WRITE_RAYS_PER_BLOCK * 32, 0, stream>>>(n_rays, ray_offset, n_rays_global, max_samples, aabb, rng, n_images, images_dev, bitfield,
			snap_to_pixel_centers != 0, cone_angle_constant, counts, n_words, words, local_bases, local_slots, block_sums, counters, ray_indices, rays, numsteps, coords);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
