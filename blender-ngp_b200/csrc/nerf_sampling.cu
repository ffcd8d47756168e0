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

struct TrainRay { V3 o, d_unnorm, d; float startt, cone_angle, tmax; bool valid; };

// Ray set-up: testbed_nerf.cu:1118-1202 (perspective / OpenCV / f-theta / lat-long lens, no rolling shutter, no distortion map,
// pixels / images uniform or drawn from the error-map CDFs). RNG draws in the reference's order: xy(2), motion-blur time(1), start jitter(1).
__device__ inline TrainRay setup_training_ray(uint32_t i, uint32_t n_rays, Pcg32 rng, uint32_t n_images, const ngpb_image* __restrict__ images,
                                              const Aabb& aabb, bool snap, float cone_angle_constant, const ErrorCdf& cdf) {
	TrainRay r;
	r.valid = false;
	const uint32_t img = image_idx(i, n_rays, n_images, cdf.img);
	const ngpb_image& im = images[img];
	rng.advance((int64_t)i * N_MAX_RANDOM_SAMPLES_PER_RAY);
	float x, y;
	random_image_pos_training(rng, im.w, im.h, snap, &x, &y, cdf, img);
	float px[4];
	if (!read_rgba(x, y, im, px)) return r; // masked-away pixel (:1126)
	(void)rng.next_float(); // motionblur_time (:1132); max_level_rand_training is off so no draw at :1130
	const float* xf = im.xform;
	r.o = {xf[9], xf[10], xf[11]};
	const V3 dc = training_ray_direction(im, x, y);
	const float dcam[3] = {dc.x, dc.y, dc.z};
	const float row0[3] = {xf[0], xf[3], xf[6]}, row1[3] = {xf[1], xf[4], xf[7]}, row2[3] = {xf[2], xf[5], xf[8]};
	r.d_unnorm = {dot3(row0, dcam), dot3(row1, dcam), dot3(row2, dcam)};
	const float z = sum3(r.d_unnorm.x * r.d_unnorm.x, r.d_unnorm.y * r.d_unnorm.y, r.d_unnorm.z * r.d_unnorm.z);
	if (z > 0.f) { const float nrm = sqrtf(z); r.d = {r.d_unnorm.x / nrm, r.d_unnorm.y / nrm, r.d_unnorm.z / nrm}; } else r.d = r.d_unnorm;
	float tmin, tmax;
	aabb_ray_intersect(aabb, r.o, r.d, &tmin, &tmax);
	r.tmax = tmax;
	r.cone_angle = cone_angle_constant; // calc_cone_angle (:87-94)
	tmin = fmaxf(tmin, 0.0f);
	float startt = tmin;
	startt += calc_dt(startt, r.cone_angle) * rng.next_float();
	r.startt = startt;
	r.valid = true;
	return r;
}

// ---- marching ----------------------------------------------------------------------------------------
// Reference: testbed_nerf.cu:1204-1219 (count pass) and :1239-1253 (write pass): one thread marches a whole ray, twice.
//
// Both branches of that march advance t the same way -- `t += calc_dt(t, cone_angle)`, once per sample in an occupied cell, repeatedly
// until the cell's exit in an empty one (advance_to_next_voxel, :449-463). So the t values of a ray form ONE chain t_0, t_1, ... that does
// not depend on the occupancy grid ("candidates"); the march only decides which candidates it VISITS (an empty visited candidate jumps to
// the first later candidate with t >= its cell-exit time; an occupied one emits a sample and moves to the next candidate) and a visited
// occupied candidate is a sample. That turns the serial march into a parallel one without changing a single result:
//   1. chain kernel (thread per ray): set-up, then the chain alone (one FADD per step), storing t at every 32nd candidate ("word");
//   2. word kernel (thread per word, ~22 per ray): marches its 32 candidates ASSUMING candidate 0 is visited, records the visited set, the
//      emitted set and how the walk leaves the word (exit state);
//   3. resolve kernel (thread per ray): walks the ray's words in order with the true entry; if the true entry candidate is in the word's
//      visited set, the walk from there on is the recorded one (the march is deterministic from any visited candidate), otherwise -- only when
//      a candidate sits within rounding of a cell boundary -- that word is re-marched from the true entry. Also applies the 1024-sample cap.
//   4. scan + write kernels: as before, parallel over words, replaying each word's 32 chain steps with the same float additions.
// The ray's critical path drops from ~200 dependent cell hops to ~7, and the kernel becomes throughput- instead of latency-bound.
// (Round 2 tried the opposite trade on top of a closed form of the constant-step chain -- one thread per ray again, crossing empty space in 8^3-cell
// blocks with exact landings: bit-exact on every K1 test, but 511 us instead of 165 us for the stage. The occupancy grid of a half-trained scene is
// fluffy -- 56 % of the 8^3 blocks hold an occupied cell at step 530 -- and 45 k serial walks cannot hide their load latency. The same exact block hops
// (8^3 and 4^3) inside march_word, keeping the word-parallel structure: bit-exact again, stage 261 us -- deciding that a hop is unambiguous costs more
// instructions than the four cell-sized hops it replaces.)
struct __align__(16) MarchWord {
	float t;            // chain value at the word's first candidate
	uint32_t visited;   // candidates the walk visits (word kernel: assuming candidate 0 is visited)
	uint32_t emitted;   // samples; after the resolve kernel: the ray's true samples in this word
	float tau;          // exit state: -inf = the next word's candidate 0 is visited; finite = pending jump target (next visited: first later
	                    // candidate with t >= tau); NaN = the ray ended in this word (a visited candidate lay outside the box)
};
struct __align__(16) RayRec { float o[3]; float cone_angle; float d[3]; uint32_t n_words; float d_unnorm[3]; float pad; };
constexpr uint32_t MARCH_MAX_WORDS = 72;         // 2304 candidates; a unit-cube ray has <= 1025, aabb_scale 128 about 2100
constexpr uint32_t MARCH_OVERFLOW = 0x80000000u; // flag in n_words: the chain outgrew the record; this ray is marched serially instead
constexpr float TAU_NEXT = -INFINITY;

__device__ __forceinline__ bool tau_is_end(float tau) { return tau != tau; }

// With cone_angle == 0 (every aabb_scale-1 scene) calc_dt is the constant minimum step (t * 0 clamps to MIN_CONE_STEPSIZE) and mip_from_dt reduces to
// mip_from_pos (dt * 256 < 1): the CONST_DT instantiations drop those per-candidate computations. Same values, fewer instructions.
template <bool CONST_DT> __device__ __forceinline__ float chain_dt(float t, float cone_angle) { return CONST_DT ? MIN_CONE_STEPSIZE : calc_dt(t, cone_angle); }
template <bool CONST_DT> __device__ __forceinline__ uint32_t chain_mip(float dt, const V3& pos) { return (uint32_t)(CONST_DT ? mip_from_pos(pos) : mip_from_dt(dt, pos)); }

// Marches the 32 candidates of one word starting with candidate `entry` visited. t_word: chain value at candidate 0.
template <bool CONST_DT>
__device__ inline void march_word(const V3& o, const V3& d, const V3& idir, float cone_angle, const Aabb& aabb, const uint8_t* __restrict__ bitfield,
                                  float t_word, uint32_t entry, uint32_t* visited_out, uint32_t* emitted_out, float* tau_out) {
	float t = t_word;
	uint32_t c = 0, visited = 0, emitted = 0;
	for (; c < entry; ++c) t += chain_dt<CONST_DT>(t, cone_angle);
	float tau = TAU_NEXT;
	while (c < 32) {
		const V3 pos = V3{o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
		if (!aabb_contains(aabb, pos)) { tau = __uint_as_float(0x7FC00000u); break; }
		visited |= 1u << c;
		const float dt = chain_dt<CONST_DT>(t, cone_angle);
		const uint32_t mip = chain_mip<CONST_DT>(dt, pos);
		if (density_grid_occupied_at(pos, bitfield, mip)) {
			emitted |= 1u << c;
			t += dt;
			++c;
		} else {
			const float t_target = t + distance_to_next_voxel(pos, d, idir, NERF_GRIDSIZE >> mip);
			do { t += chain_dt<CONST_DT>(t, cone_angle); ++c; } while (t < t_target && c < 32); // advance_to_next_voxel, cut at the word's end
			if (t < t_target) { tau = t_target; break; } // the jump continues in a later word
		}
	}
	*visited_out = visited; *emitted_out = emitted; *tau_out = tau;
}

// Serial march of a whole ray (the reference's loop), used only for rays whose chain does not fit MARCH_MAX_WORDS words.
template <bool WRITE>
__device__ inline uint32_t march_serial(const TrainRay& r, const Aabb& aabb, const uint8_t* __restrict__ bitfield, uint32_t max_steps, float* __restrict__ coords_out) {
	const V3 idir = {1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z};
	const V3 wd = {(r.d.x + 1.0f) * 0.5f, (r.d.y + 1.0f) * 0.5f, (r.d.z + 1.0f) * 0.5f}; // warp_direction (:292)
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
			t = advance_to_next_voxel(t, r.cone_angle, pos, r.d, idir, NERF_GRIDSIZE >> mip);
		}
	}
	return j;
}

constexpr uint32_t K1_BLOCK = 128;

// Pass 1: one thread per ray. Set-up and the t chain; appends one work item per word to `items`.
template <bool CONST_DT>
__global__ void __launch_bounds__(K1_BLOCK) chain_training_rays_kernel(
	const uint32_t n_rays, const uint32_t ray_offset, const uint32_t n_rays_global, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const bool snap, const float cone_angle_constant, const ErrorCdf cdf, RayRec* __restrict__ recs, MarchWord* __restrict__ words, uint32_t* __restrict__ items, uint32_t* __restrict__ n_items)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31;
	uint32_t n_words = 0;
	bool overflow = false;
	if (i < n_rays) {
		// everything about a ray derives from its GLOBAL index (:1118-1121, :1062-1083): a shard reproduces its slice of the full batch
		const TrainRay r = setup_training_ray(ray_offset + i, n_rays_global, rng, n_images, images, aabb, snap, cone_angle_constant, cdf);
		if (r.valid && r.tmax < 3.0e38f) {
			MarchWord* w = words + (size_t)i * MARCH_MAX_WORDS;
			// every candidate beyond t_end is outside the box by a margin far above the rounding of o + t * d, so no visited candidate of the
			// reference's march lies past the last stored word
			const float t_end = r.tmax + 2.0f * calc_dt(fmaxf(r.tmax, 0.f), r.cone_angle) + 1e-4f;
			float t = r.startt;
			while (true) { // one word (32 chain steps) per iteration; the end test runs once per word, so up to 31 surplus candidates are harmless
				if (n_words == MARCH_MAX_WORDS) { overflow = true; break; }
				w[n_words++].t = t;
				if (!(t <= t_end)) break;
				if (CONST_DT) t = chain_advance(t, 32); // the same 32 additions in closed form (nerf_device.cuh): no dependent chain of FADDs
				else {
					#pragma unroll
					for (int k = 0; k < 32; ++k) t += chain_dt<CONST_DT>(t, r.cone_angle);
				}
			}
		}
		RayRec rec;
		rec.o[0] = r.o.x; rec.o[1] = r.o.y; rec.o[2] = r.o.z; rec.cone_angle = r.cone_angle;
		rec.d[0] = r.d.x; rec.d[1] = r.d.y; rec.d[2] = r.d.z;
		rec.d_unnorm[0] = r.d_unnorm.x; rec.d_unnorm[1] = r.d_unnorm.y; rec.d_unnorm[2] = r.d_unnorm.z; rec.pad = 0.f;
		rec.n_words = overflow ? MARCH_OVERFLOW : n_words;
		recs[i] = rec;
	}
	// one work item per word, appended in ray order within the warp (ONE atomic per warp)
	const uint32_t mine = overflow ? 0u : n_words;
	uint32_t incl = mine;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t tt = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += tt; }
	const uint32_t warp_total = __shfl_sync(0xffffffffu, incl, 31);
	uint32_t base = 0;
	if (lane == 31 && warp_total) base = atomicAdd(n_items, warp_total);
	base = __shfl_sync(0xffffffffu, base, 31) + incl - mine;
	for (uint32_t w = 0; w < mine; ++w) items[base + w] = i * 128u + w;
}

// Pass 2: one thread per word: march the word's 32 candidates assuming candidate 0 is visited.
template <bool CONST_DT>
__global__ void __launch_bounds__(128) march_words_kernel(const uint32_t* __restrict__ n_items, const uint32_t* __restrict__ items, const Aabb aabb, const uint8_t* __restrict__ bitfield,
                                                          const RayRec* __restrict__ recs, MarchWord* __restrict__ words)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= *n_items) return;
	const uint32_t item = items[k], i = item >> 7, w = item & 127u;
	const RayRec rec = recs[i];
	const V3 o = {rec.o[0], rec.o[1], rec.o[2]}, d = {rec.d[0], rec.d[1], rec.d[2]};
	const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
	MarchWord* mw = words + (size_t)i * MARCH_MAX_WORDS + w;
	uint32_t visited, emitted; float tau;
	march_word<CONST_DT>(o, d, idir, rec.cone_angle, aabb, bitfield, mw->t, 0, &visited, &emitted, &tau);
	mw->visited = visited; mw->emitted = emitted; mw->tau = tau;
}

// Pass 3: one thread per ray: resolve the words in order, then the block-local exclusive prefixes of (count, count > 0).
template <bool CONST_DT>
__global__ void __launch_bounds__(K1_BLOCK) resolve_training_rays_kernel(
	const uint32_t n_rays, const uint32_t ray_offset, const uint32_t n_rays_global, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const uint8_t* __restrict__ bitfield, const bool snap, const float cone_angle_constant, const ErrorCdf cdf, const RayRec* __restrict__ recs, MarchWord* __restrict__ words,
	uint32_t* __restrict__ counts, uint32_t* __restrict__ n_words_out, uint32_t* __restrict__ local_bases, uint32_t* __restrict__ local_slots, uint2* __restrict__ block_sums)
{
	__shared__ uint32_t warp_sums[2][K1_BLOCK / 32];
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t c = 0;
	if (i < n_rays) {
		const RayRec rec = recs[i];
		uint32_t nw = rec.n_words;
		if (nw == MARCH_OVERFLOW) {
			const TrainRay r = setup_training_ray(ray_offset + i, n_rays_global, rng, n_images, images, aabb, snap, cone_angle_constant, cdf);
			c = march_serial<false>(r, aabb, bitfield, NERF_STEPS, nullptr);
		} else if (nw) {
			const V3 o = {rec.o[0], rec.o[1], rec.o[2]}, d = {rec.d[0], rec.d[1], rec.d[2]};
			const V3 idir = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
			MarchWord* rw = words + (size_t)i * MARCH_MAX_WORDS;
			float tau_in = TAU_NEXT;
			uint32_t w = 0;
			for (; w < nw; ++w) {
				const MarchWord mw = rw[w];
				// true entry: the first candidate of this word with t >= tau_in
				uint32_t e = 0;
				if (tau_in != TAU_NEXT) {
					float t = mw.t;
					while (e < 32 && t < tau_in) { t += chain_dt<CONST_DT>(t, rec.cone_angle); ++e; }
					if (e == 32) { rw[w].emitted = 0; continue; } // the jump passes over the whole word
				}
				uint32_t emitted; float tau_out;
				if ((mw.visited >> e) & 1u) { emitted = mw.emitted & (0xFFFFFFFFu << e); tau_out = mw.tau; }
				else { uint32_t v; march_word<CONST_DT>(o, d, idir, rec.cone_angle, aabb, bitfield, mw.t, e, &v, &emitted, &tau_out); }
				// the march stops once it holds NERF_STEPS samples (:1209)
				const uint32_t n_here = __popc(emitted);
				if (c + n_here >= NERF_STEPS) {
					while (c + __popc(emitted) > NERF_STEPS) emitted &= ~(1u << (31 - __clz((int)emitted)));
					c = NERF_STEPS;
					rw[w].emitted = emitted;
					++w;
					break;
				}
				c += n_here;
				rw[w].emitted = emitted;
				if (tau_is_end(tau_out)) { ++w; break; }
				tau_in = tau_out;
			}
			nw = w; // later words hold no samples
		}
		counts[i] = c;
		n_words_out[i] = nw;
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

// Pass 4: one block. Exclusive scan of the per-block totals: sample bases (the reference's numsteps_counter atomicAdd, :1225)
// and ray slots (the ray_counter atomicAdd, :1232). counters[0] = samples requested; counters[1] is zeroed and counted by pass 5.
// Allocation order. Ray order while every sample fits max_samples. When the demand exceeds it, the rays served last are dropped (:1226): in ray order
// those would always be the rays of the LAST training images (the ray index selects the image, :1076-1082), whereas the reference's atomics drop
// whichever rays its blocks process last. So the order then starts at a ray drawn from the step's RNG (past the rays' own sub-streams) and wraps
// around: rot = {that ray, its sample base and its slot in ray order, number of sample-bearing rays}; pass 5 rotates bases and slots by it.
__global__ void __launch_bounds__(1024) scan_training_samples_kernel(const uint32_t n_blocks, uint2* __restrict__ block_sums, uint32_t* __restrict__ counters,
                                                                     const uint32_t n_rays, const uint32_t n_rays_global, const uint32_t max_samples, const Pcg32 rng,
                                                                     const uint32_t* __restrict__ local_bases, const uint32_t* __restrict__ local_slots, uint32_t* __restrict__ rot)
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
	__syncthreads(); // (thread 0 reads prefixes written by other threads)
	if (threadIdx.x == 0) {
		counters[0] = carry_c; counters[1] = 0;
		uint32_t first = 0, start_c = 0, start_z = 0;
		if (carry_c > max_samples && n_rays > 0) {
			Pcg32 r = rng;
			r.advance((int64_t)n_rays_global * N_MAX_RANDOM_SAMPLES_PER_RAY);
			first = r.next_uint() % n_rays;
			const uint2 bp = block_sums[first / K1_BLOCK];
			start_c = bp.x + local_bases[first]; start_z = bp.y + local_slots[first];
		}
		rot[0] = first; rot[1] = start_c; rot[2] = start_z; rot[3] = carry_z;
	}
}

// Pass 5: one warp per ray, one lane per march word. A ray is kept iff it has samples and base + count <= max_samples
// (:1221-1228); since bases grow with the ray index the kept rays are exactly the sample-bearing rays before the cut, so a kept
// ray's slot is the number of sample-bearing rays before it.
constexpr uint32_t WRITE_RAYS_PER_BLOCK = 8;
template <bool CONST_DT>
__global__ void __launch_bounds__(WRITE_RAYS_PER_BLOCK * 32) write_training_samples_kernel(
	const uint32_t n_rays, const uint32_t ray_offset, const uint32_t n_rays_global, const uint32_t max_samples, const Aabb aabb, const Pcg32 rng, const uint32_t n_images, const ngpb_image* __restrict__ images,
	const uint8_t* __restrict__ bitfield, const bool snap, const float cone_angle_constant, const ErrorCdf cdf,
	const uint32_t* __restrict__ counts, const uint32_t* __restrict__ n_words, const RayRec* __restrict__ recs, const MarchWord* __restrict__ words,
	const uint32_t* __restrict__ local_bases, const uint32_t* __restrict__ local_slots, const uint2* __restrict__ block_prefix, const uint32_t* __restrict__ rot,
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
		const uint32_t total = counters[0];
		if (total > max_samples) { // the demand exceeds the budget: allocation starts at ray rot[0] and wraps around (see pass 4)
			const uint32_t first = rot[0], start_c = rot[1], start_z = rot[2], total_z = rot[3];
			if (i >= first) { base -= start_c; slot -= start_z; } else { base += total - start_c; slot += total_z - start_z; }
		}
		kept = count > 0 && base + count <= max_samples;
	}
	const uint32_t n_kept_block = __syncthreads_count(kept && lane == 0);
	if (threadIdx.x == 0 && n_kept_block) atomicAdd(&counters[1], n_kept_block);
	if (!kept) return;

	const RayRec rec = recs[i];
	const V3 o = {rec.o[0], rec.o[1], rec.o[2]}, d = {rec.d[0], rec.d[1], rec.d[2]};
	if (lane == 0) {
		ray_indices[slot] = ray_offset + i; // global ray index: K6 re-derives pixel and background from it
		float* ro = rays + (size_t)slot * 6;
		ro[0] = o.x; ro[1] = o.y; ro[2] = o.z; ro[3] = rec.d_unnorm[0]; ro[4] = rec.d_unnorm[1]; ro[5] = rec.d_unnorm[2];
		numsteps[slot * 2 + 0] = count;
		numsteps[slot * 2 + 1] = base;
	}
	float* out = coords + (size_t)base * COORD_FLOATS;
	const uint32_t nw = n_words[i];
	if (nw & MARCH_OVERFLOW) {
		if (lane == 0) {
			const TrainRay r = setup_training_ray(ray_offset + i, n_rays_global, rng, n_images, images, aabb, snap, cone_angle_constant, cdf);
			march_serial<true>(r, aabb, bitfield, count, out);
		}
		return;
	}
	const V3 wd = {(d.x + 1.0f) * 0.5f, (d.y + 1.0f) * 0.5f, (d.z + 1.0f) * 0.5f}; // warp_direction (:292)
	const MarchWord* rw = words + (size_t)i * MARCH_MAX_WORDS;
	// One lane per SAMPLE (a lane per word leaves most lanes idle: a ray's ~15 samples sit in two or three of its ~22 words).
	// Words are taken 32 at a time (lane = word): prefix counts give every sample its word; the lane then replays that word's chain up to
	// the sample's candidate (the same float additions as the march).
	uint32_t carry = 0;
	for (uint32_t w0 = 0; w0 < nw; w0 += 32) {
		const uint32_t w = w0 + lane;
		MarchWord mw = {0.f, 0u, 0u, 0.f};
		if (w < nw) mw = rw[w];
		const uint32_t n_here = __popc(mw.emitted);
		uint32_t incl = n_here;
		#pragma unroll
		for (int k = 1; k < 32; k <<= 1) { const uint32_t tt = __shfl_up_sync(0xffffffffu, incl, k); if (lane >= (uint32_t)k) incl += tt; }
		const uint32_t excl = incl - n_here;
		const uint32_t n_chunk = __shfl_sync(0xffffffffu, incl, 31);
		for (uint32_t s0 = 0; s0 < n_chunk; s0 += 32) {
			const uint32_t s = s0 + lane; // sample index within this chunk of words
			// the last word whose exclusive prefix is <= s holds sample s
			uint32_t lo = 0;
			#pragma unroll
			for (uint32_t step = 16; step > 0; step >>= 1) {
				const uint32_t cand = lo + step;
				const uint32_t v = __shfl_sync(0xffffffffu, excl, cand & 31);
				if (v <= s) lo = cand;
			}
			const uint32_t mask = __shfl_sync(0xffffffffu, mw.emitted, lo);
			const float t_word = __shfl_sync(0xffffffffu, mw.t, lo);
			const uint32_t rank = s - __shfl_sync(0xffffffffu, excl, lo);
			if (s < n_chunk) {
				const uint32_t bit = __fns(mask, 0, (int)rank + 1); // position of the (rank+1)-th set bit
				float t = t_word;
				if (CONST_DT) t = chain_advance(t, bit);
				else for (uint32_t k = 0; k < bit; ++k) t += chain_dt<CONST_DT>(t, rec.cone_angle);
				const float dt = chain_dt<CONST_DT>(t, rec.cone_angle);
				const V3 pos = V3{o.x + t * d.x, o.y + t * d.y, o.z + t * d.z};
				const V3 wp = warp_position(pos, aabb);
				float* c = out + (size_t)(carry + s) * COORD_FLOATS;
				c[0] = wp.x; c[1] = wp.y; c[2] = wp.z; c[3] = warp_dt(dt); c[4] = wd.x; c[5] = wd.y; c[6] = wd.z;
			}
		}
		carry += n_chunk;
	}
}

Aabb make_aabb(const float* a) {
	Aabb b;
	for (int c = 0; c < 3; ++c) { b.min[c] = a[c]; b.max[c] = a[3 + c]; }
	return b;
}

} // namespace ngpb

using namespace ngpb;

namespace {
struct K1Scratch {
	uint32_t *counts, *n_words, *local_bases, *local_slots, *items, *n_items;
	uint2* block_sums;
	RayRec* recs;
	MarchWord* words;
	size_t bytes;
};
// Carves (base == nullptr: only sizes) the K1 scratch buffer for n_rays rays.
K1Scratch k1_scratch(uint32_t n_rays, void* base) {
	uint8_t* p = reinterpret_cast<uint8_t*>(base);
	size_t off = 0;
	auto take = [&](size_t bytes) { uint8_t* q = p ? p + off : nullptr; off += (bytes + 255) / 256 * 256; return q; };
	K1Scratch s{};
	s.counts = (uint32_t*)take((size_t)n_rays * 4);
	s.n_words = (uint32_t*)take((size_t)n_rays * 4);
	s.local_bases = (uint32_t*)take((size_t)n_rays * 4);
	s.local_slots = (uint32_t*)take((size_t)n_rays * 4);
	s.block_sums = (uint2*)take((size_t)div_round_up(n_rays, K1_BLOCK) * sizeof(uint2));
	s.n_items = (uint32_t*)take(256);
	s.recs = (RayRec*)take((size_t)n_rays * sizeof(RayRec));
	s.items = (uint32_t*)take((size_t)n_rays * MARCH_MAX_WORDS * 4);
	s.words = (MarchWord*)take((size_t)n_rays * MARCH_MAX_WORDS * sizeof(MarchWord));
	s.bytes = off;
	return s;
}
} // namespace

extern "C" uint64_t ngpb_generate_training_samples_scratch_bytes(uint32_t n_rays) { return k1_scratch(n_rays, nullptr).bytes; }

extern "C" int ngpb_generate_training_samples(void* stream, uint32_t n_rays, const float* aabb6, uint32_t max_samples, ngpb_rng rng,
                                              uint32_t n_images, const ngpb_image* images_dev, const uint8_t* bitfield,
                                              int snap_to_pixel_centers, float cone_angle_constant,
                                              uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch) {
	return ngpb_generate_training_samples_sharded(stream, n_rays, 0, n_rays, aabb6, max_samples, rng, n_images, images_dev, bitfield, snap_to_pixel_centers, cone_angle_constant,
		counters, ray_indices, rays, numsteps, coords, scratch);
}

extern "C" int ngpb_generate_training_samples_sharded(void* stream, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, ngpb_rng rng,
                                              uint32_t n_images, const ngpb_image* images_dev, const uint8_t* bitfield,
                                              int snap_to_pixel_centers, float cone_angle_constant,
                                              uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch) {
	return ngpb_generate_training_samples_cdf(stream, n_rays, ray_offset, n_rays_global, aabb6, max_samples, rng, n_images, images_dev, bitfield, snap_to_pixel_centers, cone_angle_constant,
		counters, ray_indices, rays, numsteps, coords, scratch, nullptr);
}

// K1 with the error-map CDFs (K19): pixels and / or images drawn proportionally to the accumulated training error (:991-1083). error_cdf == NULL or
// null members: the uniform choices above.
extern "C" int ngpb_generate_training_samples_cdf(void* stream_, uint32_t n_rays, uint32_t ray_offset, uint32_t n_rays_global, const float* aabb6, uint32_t max_samples, ngpb_rng rng_,
                                              uint32_t n_images, const ngpb_image* images_dev, const uint8_t* bitfield,
                                              int snap_to_pixel_centers, float cone_angle_constant,
                                              uint32_t* counters, uint32_t* ray_indices, float* rays, uint32_t* numsteps, float* coords, void* scratch_, const ngpb_error_cdf* error_cdf) {
	try {
		void* scratch = scratch_;
		const ErrorCdf cdf = make_error_cdf(error_cdf);
		if (cdf.x_cond_y && (!cdf.y || cdf.res_x <= 0 || cdf.res_y <= 0)) {
			set_last_error("ngpb_generate_training_samples: cdf_x_cond_y needs cdf_y and a resolution");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		if (!aabb6 || !images_dev || !bitfield || !counters || !ray_indices || !rays || !numsteps || !coords || !scratch || n_images == 0 ||
		    (uint64_t)ray_offset + n_rays > n_rays_global) {
			set_last_error("ngpb_generate_training_samples: invalid argument");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		cudaStream_t stream = (cudaStream_t)stream_;
		if (n_rays == 0) { NGPB_CUDA_CHECK(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), stream)); return 0; }
		const Aabb aabb = make_aabb(aabb6);
		Pcg32 rng; rng.state = rng_.state; rng.inc = rng_.inc;
		const K1Scratch sc = k1_scratch(n_rays, scratch);
		uint32_t *counts = sc.counts, *n_words = sc.n_words, *local_bases = sc.local_bases, *local_slots = sc.local_slots;
		uint2* block_sums = sc.block_sums;
		MarchWord* words = sc.words;
		const uint32_t blocks = div_round_up(n_rays, K1_BLOCK);
		const bool snap = snap_to_pixel_centers != 0;
		NGPB_CUDA_CHECK(cudaMemsetAsync(sc.n_items, 0, 4, stream));
		const bool const_dt = cone_angle_constant == 0.f;
		// Development aid: NGPB_K1_SMEM = bytes of (unused) dynamic shared memory per block, which caps the resident blocks per SM of the K1 kernels so that
		// kernels of the training stream can co-reside while K1 runs on the sampling stream.
		static const uint32_t k1_smem = [] { const char* e = std::getenv("NGPB_K1_SMEM"); return e ? (uint32_t)std::atoi(e) : 0u; }();
		#define NGPB_K1(kernel, grid, block, ...) do { if (const_dt) kernel<true><<<grid, block, k1_smem, stream>>>(__VA_ARGS__); else kernel<false><<<grid, block, k1_smem, stream>>>(__VA_ARGS__); } while (0)
		NGPB_K1(chain_training_rays_kernel, blocks, K1_BLOCK, n_rays, ray_offset, n_rays_global, aabb, rng, n_images, images_dev, snap, cone_angle_constant, cdf,
			sc.recs, words, sc.items, sc.n_items);
		NGPB_LAUNCH_CHECK();
		// one thread per word; the word count lives on the device, so the grid covers the worst case and surplus blocks exit at once.
		// A ray of the unit cube has at most 33 words (1025 candidates); larger scenes up to MARCH_MAX_WORDS.
		const uint64_t max_items = (uint64_t)n_rays * MARCH_MAX_WORDS;
		NGPB_K1(march_words_kernel, (uint32_t)((max_items + 127) / 128), 128, sc.n_items, sc.items, aabb, bitfield, sc.recs, words);
		NGPB_LAUNCH_CHECK();
		NGPB_K1(resolve_training_rays_kernel, blocks, K1_BLOCK, n_rays, ray_offset, n_rays_global, aabb, rng, n_images, images_dev, bitfield, snap, cone_angle_constant, cdf,
			sc.recs, words, counts, n_words, local_bases, local_slots, block_sums);
		NGPB_LAUNCH_CHECK();
		uint32_t* rot = sc.n_items + 4; // (four words of the 256-byte slot that holds the item counter)
		scan_training_samples_kernel<<<1, 1024, 0, stream>>>(blocks, block_sums, counters, n_rays, n_rays_global, max_samples, rng, local_bases, local_slots, rot);
		NGPB_LAUNCH_CHECK();
		NGPB_K1(write_training_samples_kernel, div_round_up(n_rays, WRITE_RAYS_PER_BLOCK), WRITE_RAYS_PER_BLOCK * 32, n_rays, ray_offset, n_rays_global, max_samples, aabb, rng, n_images, images_dev, bitfield,
			snap_to_pixel_centers != 0, cone_angle_constant, cdf, counts, n_words, sc.recs, words, local_bases, local_slots, block_sums, rot, counters, ray_indices, rays, numsteps, coords);
		NGPB_LAUNCH_CHECK();
		#undef NGPB_K1
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
