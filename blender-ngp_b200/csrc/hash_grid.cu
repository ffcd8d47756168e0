// Multiresolution hash-grid encoding, forward and backward, 3-D positions, 2 features per level.
// Replaces tcnn kernel_grid / kernel_grid_backward (reference:
// dependencies/tiny-cuda-nn/include/tiny-cuda-nn/encodings/grid.h:220-349, :395-518) for
// GridType::Hash, HashType::CoherentPrime, InterpolationType::Linear (grid.h:1472-1476 defaults).
//
// Layout: see the note above hash_encode_forward_kernel. The whole table (24.4 MB fp16) is L2-resident on B200.
#include "common.cuh"
#include <cstdlib>
#include "../../include/ngpb.h"

namespace ngpb {

struct GridLevels {
	float scale[NGPB_MAX_LEVELS];
	uint32_t resolution[NGPB_MAX_LEVELS];
	uint32_t offset[NGPB_MAX_LEVELS];
	uint32_t size[NGPB_MAX_LEVELS];
	uint32_t n_levels;
};

static GridLevels make_levels(const ngpb_grid* g) {
	GridLevels L{};
	L.n_levels = g->n_levels;
	for (uint32_t l = 0; l < g->n_levels; ++l) {
		L.scale[l] = g->scale[l];
		L.resolution[l] = g->resolution[l];
		L.offset[l] = g->offsets[l];
		L.size[l] = g->offsets[l + 1] - g->offsets[l];
	}
	return L;
}

// Entry index of a grid vertex. Reference: grid_index + prime_hash<3,true>, grid.h:111-128,:164-186.
// Dense levels: x + y*res + z*res^2; hashed levels (size < res^3): x ^ y*2654435761 ^ z*805459861;
// both reduced modulo the level size. `% size` is exact; the fast paths avoid the division.
struct LevelIndexer {
	uint32_t size, res, res2;
	bool hashed, pow2;
	__device__ LevelIndexer(uint32_t size_, uint32_t res_) : size(size_), res(res_) {
		// the reference's stride loop (grid.h:169-173) stops multiplying once stride > size, so
		// `size < stride` <=> size < res^3 without overflow for res <= 2^10.. use 64-bit to be safe.
		uint64_t dense = (uint64_t)res_ * res_ * res_;
		hashed = (uint64_t)size_ < dense;
		res2 = res_ * res_;
		pow2 = (size_ & (size_ - 1)) == 0;
	}
	__device__ uint32_t operator()(uint32_t x, uint32_t y, uint32_t z) const {
		uint32_t index;
		if (hashed) {
			index = x ^ (y * 2654435761u) ^ (z * 805459861u);
			return pow2 ? (index & (size - 1)) : (index % size);
		}
		index = x + y * res + z * res2;
		if (index >= size) { index -= size; if (index >= size) index %= size; }
		return index;
	}
};

// The reference's index loop for a level where stride overflows hashmap_size early adds only the
// dims visited before `stride > hashmap_size`; when the level is hashed that partial sum is
// discarded, and when it is dense all three dims are visited. So the two cases above are complete.

// Backward thread mapping: a block of 512 threads handles ENC_SAMPLES = 32 consecutive samples, warp w works on level w (and w+16, ...),
// lane l on sample s0 + l: the level is warp-uniform (uniform constant loads) and dL/dy enters through shared memory.
constexpr uint32_t ENC_SAMPLES = 32;
constexpr uint32_t ENC_WARPS = 16;
constexpr uint32_t COARSE_RES = 128; // levels up to this resolution aggregate their atomics per run of equal cells (cell coordinates fit 8 bits)

__device__ __forceinline__ void level_position(const float* __restrict__ positions, size_t i, uint32_t pos_stride, float scale, float pos[3], uint32_t pg[3]) {
	// pos_fract, tcnn common_device.h:434-445. The reference's `input * scale + 0.5f` is contracted to one FFMA by nvcc's default
	// -fmad=true; this library is built with -fmad=false, so the FMA is explicit.
	#pragma unroll
	for (int d = 0; d < 3; ++d) {
		const float p = __fmaf_rn(positions[i * pos_stride + d], scale, 0.5f);
		const float fl = floorf(p);
		pg[d] = (uint32_t)(int)fl;
		pos[d] = p - fl;
	}
}

// Forward. One thread per (sample, level); the 16 level-threads of a sample are adjacent, so a warp holds two consecutive samples, the
// position is a broadcast load and the 16 half2 results of a sample form one coalesced 64-byte row of `encoded[n][32]`. Consecutive warps
// work on consecutive samples of a ray at the same time, which keeps the coarse and middle levels' lines hot in L1 (a level-per-warp
// mapping was measured: 45 % L1 hit rate instead of 69 %, twice the L2 sectors, 1.6x slower).
// The per-level constants come from a shared-memory copy of the table: indexing the kernel-parameter arrays with a per-thread level
// compiles to divergent constant-bank loads, which serialise in the address-divergence unit (ncu: ADU pipe at 88 % of peak).
struct LevelConst { float scale; uint32_t size, resolution, offset; };

// LevelConst.offset carries two flags in its top bits (entry offsets stay far below 2^30): the level is hashed (smaller than its dense grid), and
// its size is a power of two. Both are per-level facts; evaluating them per corner cost a 64-bit product and a branch around every index.
constexpr uint32_t LEVEL_HASHED = 0x80000000u, LEVEL_POW2 = 0x40000000u, LEVEL_OFFSET_MASK = 0x3FFFFFFFu;
__device__ __forceinline__ LevelConst make_level_const(const GridLevels& L, uint32_t l) {
	const uint32_t size = L.size[l], res = L.resolution[l];
	const bool hashed = (uint64_t)size < (uint64_t)res * res * res, pow2 = (size & (size - 1)) == 0;
	return LevelConst{L.scale[l], size, res, L.offset[l] | (hashed ? LEVEL_HASHED : 0u) | (pow2 ? LEVEL_POW2 : 0u)};
}

// The 8 corner entry indices of a cell (grid_index + prime_hash, grid.h:111-128,:164-186), with the per-axis products shared between corners.
__device__ __forceinline__ void corner_indices(const LevelConst& c, const uint32_t pg[3], uint32_t idx[8]) {
	const uint32_t x[2] = {pg[0], pg[0] + 1};
	if (c.offset & LEVEL_HASHED) { // hashed level: x ^ y * 2654435761 ^ z * 805459861
		const uint32_t y0 = pg[1] * 2654435761u, z0 = pg[2] * 805459861u;
		const uint32_t y[2] = {y0, y0 + 2654435761u}, z[2] = {z0, z0 + 805459861u};
		#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) idx[k] = x[k & 1] ^ y[(k >> 1) & 1] ^ z[k >> 2];
		if (c.offset & LEVEL_POW2) {
			const uint32_t mask = c.size - 1;
			#pragma unroll
			for (uint32_t k = 0; k < 8; ++k) idx[k] &= mask;
		} else {
			#pragma unroll
			for (uint32_t k = 0; k < 8; ++k) idx[k] %= c.size;
		}
	} else { // dense level: x + y * res + z * res^2, wrapped into the level (the +0.5 offset can index `res`, common_device.h:404-408)
		const uint32_t y0 = pg[1] * c.resolution, z0 = pg[2] * c.resolution * c.resolution;
		const uint32_t y[2] = {y0, y0 + c.resolution}, z[2] = {z0, z0 + c.resolution * c.resolution};
		bool far = false; // one subtraction wraps every index a position inside the unit cube can produce; anything beyond takes the modulo
		#pragma unroll
		for (uint32_t k = 0; k < 8; ++k) {
			uint32_t h = x[k & 1] + y[(k >> 1) & 1] + z[k >> 2];
			h = h >= c.size ? h - c.size : h;
			far |= h >= c.size;
			idx[k] = h;
		}
		if (far) {
			#pragma unroll
			for (uint32_t k = 0; k < 8; ++k) idx[k] %= c.size;
		}
	}
}

// Per-(sample, level) work shared by the forward kernels: position -> 8 gathers -> blended feature pair.
__device__ __forceinline__ __half2 encode_one(const LevelConst& c, const __half2* __restrict__ grid, const float* __restrict__ positions, size_t i, uint32_t pos_stride) {
	const __half2* __restrict__ g = grid + (c.offset & LEVEL_OFFSET_MASK);
	float pos[3];
	uint32_t pg[3];
	level_position(positions, i, pos_stride, c.scale, pos, pg);
	uint32_t idx[8];
	corner_indices(c, pg, idx);

	// issue the gathers first, then blend: maximises loads in flight per thread. The two x-neighbours of a corner pair are adjacent entries of one aligned
	// 8-byte pair whenever their indices differ in bit 0 only -- always for an even x on a hashed level ((x ^ c) and ((x + 1) ^ c)), and for an even entry
	// index on a dense one -- and then ONE 8-byte load fetches both: a quarter fewer L1 wavefronts per warp, which is what bounds this kernel.
	__half2 v[8];
	#pragma unroll
	for (uint32_t p = 0; p < 4; ++p) {
		const uint32_t a = idx[2 * p], b = idx[2 * p + 1];
		if (b == (a ^ 1u)) {
			const uint2 raw = __ldg(reinterpret_cast<const uint2*>(g + (a & ~1u))); // (level offsets are multiples of 8 entries: 8-byte aligned)
			const uint32_t lo = (a & 1u) ? raw.y : raw.x, hi = (a & 1u) ? raw.x : raw.y;
			v[2 * p] = *reinterpret_cast<const __half2*>(&lo);
			v[2 * p + 1] = *reinterpret_cast<const __half2*>(&hi);
		} else {
			v[2 * p] = __ldg(g + a);
			v[2 * p + 1] = __ldg(g + b);
		}
	}

	// trilinear weights in the reference's multiplication order ((wx * wy) * wz, grid.h:326-337), products shared between corners
	const float wx[2] = {1.f - pos[0], pos[0]}, wy[2] = {1.f - pos[1], pos[1]}, wz[2] = {1.f - pos[2], pos[2]};
	float wxy[4];
	#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) wxy[k] = wx[k & 1] * wy[k >> 1];
	// fp16 accumulation in the reference's corner order and rounding: result += (half)(weight * data), grid.h:339-341 (both features at once)
	__half2 r = __floats2half2_rn(0.f, 0.f);
	#pragma unroll
	for (uint32_t k = 0; k < 8; ++k) {
		const float w = wxy[k & 3] * wz[k >> 2];
		const float2 f = __half22float2(v[k]);
		r = __hadd2(r, __floats2half2_rn(w * f.x, w * f.y));
	}
	return r;
}

// Generic level count: thread per (sample, level), level fastest.
__global__ void __launch_bounds__(256) hash_encode_forward_kernel(
	const uint32_t n, const uint32_t* __restrict__ n_dev, const GridLevels L, const __half2* __restrict__ grid, const float* __restrict__ positions, const uint32_t pos_stride,
	__half2* __restrict__ encoded)
{
	__shared__ LevelConst lc[NGPB_MAX_LEVELS];
	if (threadIdx.x < L.n_levels) lc[threadIdx.x] = make_level_const(L, threadIdx.x);
	__syncthreads();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t level = tid % L.n_levels, i = tid / L.n_levels;
	if (i >= n) return;
	if (n_dev && i >= *n_dev) return; // device-side sample count (no host round trip between K1 and the network)
	encoded[(size_t)i * L.n_levels + level] = encode_one(lc[level], grid, positions, i, pos_stride);
}

// 16 levels (every NeRF / SDF config of the reference): a block is 16 consecutive samples x 16 levels, as before, but a WARP is 16 samples x 2
// levels. Dense and hashed levels index differently; with all 16 levels of a sample in one warp every warp executed both index computations
// (ncu: 26 of 32 lanes active, 137 warp instructions per sample). Grouped by level, only the warp that holds the last dense and the first hashed
// level diverges, and the 16 samples of a ray that share a coarse cell hit the same lines within one load. The block's working set -- what decides
// the L1 hit rate -- is unchanged. Results go through a shared-memory tile so that the block still writes its 1 KB of `encoded` rows contiguously.
// TILED: the block's 1 KB of output is written in the MLP kernels' UMMA core-matrix layout instead of row-major (umma.cuh tile_offset: per 128 samples an
// 8 KB block, 8-row groups of 512 B, inside a group four 128-byte chunks of 4 levels x 8 rows). A block of 16 samples covers two whole row groups, i.e. the
// SAME contiguous 1 KB as in the row-major layout, only permuted inside: the stores stay fully coalesced and the MLP kernel fetches a tile with one bulk copy.
template <bool TILED>
__global__ void __launch_bounds__(256) hash_encode_forward16_kernel(
	const uint32_t n, const uint32_t* __restrict__ n_dev, const GridLevels L, const __half2* __restrict__ grid, const float* __restrict__ positions, const uint32_t pos_stride,
	__half2* __restrict__ encoded)
{
	__shared__ LevelConst lc[16];
	__shared__ __half2 tile[16][17];
	if (threadIdx.x < 16) lc[threadIdx.x] = make_level_const(L, threadIdx.x);
	__syncthreads();
	const uint32_t level = threadIdx.x >> 4, s = threadIdx.x & 15u;
	const uint32_t i0 = blockIdx.x * 16u, i = i0 + s;
	const uint32_t n_eff = n_dev ? min(n, *n_dev) : n; // device-side sample count (no host round trip between K1 and the network)
	if (i0 >= n_eff) return;
	if (i < n_eff) tile[s][level] = encode_one(lc[level], grid, positions, i, pos_stride);
	__syncthreads();
	// (sample, level) of the element this thread writes out: thread k owns bytes [4k, 4k + 4) of the block's 1 KB
	uint32_t row, col;
	if (TILED) { row = ((threadIdx.x >> 7) << 3) | ((threadIdx.x >> 2) & 7u); col = (((threadIdx.x >> 5) & 3u) << 2) | (threadIdx.x & 3u); }
	else { row = threadIdx.x >> 4; col = threadIdx.x & 15u; }
	if (i0 + row < n_eff) encoded[(size_t)i0 * 16u + threadIdx.x] = tile[row][col];
}

// Two-dimensional input (the neural-image model: N_POS_DIMS = 2 of kernel_grid, grid.h:220-349): four corners, x ^ y * 2654435761 when hashed.
__global__ void __launch_bounds__(256) hash_encode_forward_2d_kernel(
	const uint32_t n, const GridLevels L, const __half2* __restrict__ grid, const float* __restrict__ positions, const uint32_t pos_stride, __half2* __restrict__ encoded)
{
	__shared__ LevelConst lc[NGPB_MAX_LEVELS];
	if (threadIdx.x < L.n_levels) lc[threadIdx.x] = LevelConst{L.scale[threadIdx.x], L.size[threadIdx.x], L.resolution[threadIdx.x], L.offset[threadIdx.x]};
	__syncthreads();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t level = tid % L.n_levels, i = tid / L.n_levels;
	if (i >= n) return;
	const LevelConst c = lc[level];
	const __half2* __restrict__ g = grid + c.offset;
	float pos[2];
	uint32_t pg[2];
	#pragma unroll
	for (int d = 0; d < 2; ++d) { // pos_fract (tcnn common_device.h:434-445), see level_position
		const float p = __fmaf_rn(positions[(size_t)i * pos_stride + d], c.scale, 0.5f);
		const float fl = floorf(p);
		pg[d] = (uint32_t)(int)fl;
		pos[d] = p - fl;
	}
	// grid_index (grid.h:164-186): the stride loop stops once the stride exceeds the level size; hashed when the level is smaller than res^2
	const bool x_only = c.resolution > c.size;                                   // stride after dim 0 already exceeds the table
	const bool hashed = (uint64_t)c.size < (uint64_t)c.resolution * (x_only ? 1u : c.resolution);
	__half2 r = __floats2half2_rn(0.f, 0.f);
	#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		const uint32_t x = pg[0] + (k & 1), y = pg[1] + (k >> 1);
		const uint32_t h = hashed ? (x ^ (y * 2654435761u)) : (x_only ? x : x + y * c.resolution);
		const __half2 v = __ldg(g + h % c.size);
		const float w = ((k & 1) ? pos[0] : 1.f - pos[0]) * ((k >> 1) ? pos[1] : 1.f - pos[1]);
		const float2 f = __half22float2(v);
		r = __hadd2(r, __floats2half2_rn(w * f.x, w * f.y));
	}
	encoded[(size_t)i * L.n_levels + level] = r;
}

// Backward: scatter-add weight * dL/dy into the fp32 gradient table. The reference uses atomicAdd(__half2) into an fp16 table
// (grid.h:436-441); here the table is fp32 (one vectorised red.global.add.v2.f32 per corner), which removes the order-dependent fp16
// rounding. Same thread mapping as the forward kernel; dL/dy enters through shared memory (coalesced read of the block's 2 KB).
__global__ void __launch_bounds__(ENC_SAMPLES * ENC_WARPS) hash_encode_backward_kernel(
	const uint32_t n, const GridLevels L, const float* __restrict__ positions, const uint32_t pos_stride,
	const __half2* __restrict__ dL_dencoded, float2* __restrict__ grid_grad, const uint32_t level_begin, const uint32_t level_end)
{
	__shared__ __half2 tile[ENC_SAMPLES][NGPB_MAX_LEVELS + 1];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t s0 = blockIdx.x * ENC_SAMPLES;
	if (s0 >= n) return;
	const uint32_t n_here = min(ENC_SAMPLES, n - s0);
	for (uint32_t k = threadIdx.x; k < n_here * L.n_levels; k += blockDim.x) {
		tile[k / L.n_levels][k % L.n_levels] = dL_dencoded[(size_t)s0 * L.n_levels + k];
	}
	__syncthreads();
	const uint32_t i = s0 + lane;
	const bool valid = i < n; // (no early exit: the coarse levels use warp-wide shuffles)

	for (uint32_t level = level_begin + warp; level < level_end; level += ENC_WARPS) {
		const __half2 gh = valid ? tile[lane][level] : __floats2half2_rn(0.f, 0.f);
		const float g0 = __low2float(gh), g1 = __high2float(gh);
		const bool contributes = g0 != 0.f || g1 != 0.f; // a zero gradient adds nothing
		const uint32_t res = L.resolution[level];
		const LevelIndexer index_of(L.size[level], res);
		float2* __restrict__ gg = grid_grad + L.offset[level];
		float pos[3] = {0.f, 0.f, 0.f};
		uint32_t pg[3] = {0u, 0u, 0u};
		if (contributes) level_position(positions, i, pos_stride, L.scale[level], pos, pg);

		if (res <= COARSE_RES) {
			// Coarse level (warp-uniform branch): consecutive samples of a ray sit in the same cell, so most of the warp's 32 x 8 updates go
			// to a handful of entries -- and the whole batch hammers a few thousand addresses. Sum the contributions of each run of lanes
			// with equal cell over the run (segmented warp scan) and let the run's last lane issue the 8 atomics.
			// (a cell coordinate that does not fit a byte -- positions outside [0,1] -- keeps its lane as a run of its own)
			const uint32_t key = contributes && (pg[0] | pg[1] | pg[2]) < 256u ? (pg[0] | (pg[1] << 8) | (pg[2] << 16)) : (0xFF000000u | lane);
			const uint32_t key_prev = __shfl_up_sync(0xffffffffu, key, 1), key_next = __shfl_down_sync(0xffffffffu, key, 1);
			const bool head = lane == 0 || key != key_prev, last = lane == 31 || key != key_next;
			uint32_t head_lane = head ? lane : 0u; // inclusive max-scan: the lane at which this lane's run starts
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const uint32_t h = __shfl_up_sync(0xffffffffu, head_lane, o); if (lane >= (uint32_t)o) head_lane = max(head_lane, h); }
			float acc[8][2];
			#pragma unroll
			for (uint32_t idx = 0; idx < 8; ++idx) {
				float w = 1.f;
				#pragma unroll
				for (int d = 0; d < 3; ++d) w *= (idx & (1u << d)) ? pos[d] : (1.f - pos[d]);
				acc[idx][0] = contributes ? g0 * w : 0.f;
				acc[idx][1] = contributes ? g1 * w : 0.f;
			}
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const bool take = lane >= (uint32_t)o && lane - (uint32_t)o >= head_lane;
				#pragma unroll
				for (uint32_t idx = 0; idx < 8; ++idx) {
					const float a0 = __shfl_up_sync(0xffffffffu, acc[idx][0], o), a1 = __shfl_up_sync(0xffffffffu, acc[idx][1], o);
					if (take) { acc[idx][0] += a0; acc[idx][1] += a1; }
				}
			}
			if (last && contributes) {
				#pragma unroll
				for (uint32_t idx = 0; idx < 8; ++idx) {
					const uint32_t e = index_of(pg[0] + (idx & 1), pg[1] + ((idx >> 1) & 1), pg[2] + ((idx >> 2) & 1));
					float* addr = reinterpret_cast<float*>(gg + e);
					asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr), "f"(acc[idx][0]), "f"(acc[idx][1]) : "memory");
				}
			}
		} else if (contributes) {
			// (fusing the two x-neighbours of a corner pair into one red.global.add.v4.f32, as the forward kernel fuses its loads, was measured slower: 72 vs 64 us)
			#pragma unroll
			for (uint32_t idx = 0; idx < 8; ++idx) {
				float w = 1.f;
				#pragma unroll
				for (int d = 0; d < 3; ++d) w *= (idx & (1u << d)) ? pos[d] : (1.f - pos[d]);
				const uint32_t e = index_of(pg[0] + (idx & 1), pg[1] + ((idx >> 1) & 1), pg[2] + ((idx >> 2) & 1));
				float* addr = reinterpret_cast<float*>(gg + e);
				asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr), "f"(g0 * w), "f"(g1 * w) : "memory");
			}
		}
	}
}

// Two-dimensional input (kernel_grid_backward with N_POS_DIMS = 2, grid.h:395-518): one thread per (sample, level), four corners.
__global__ void __launch_bounds__(256) hash_encode_backward_2d_kernel(
	const uint32_t n, const GridLevels L, const float* __restrict__ positions, const uint32_t pos_stride, const __half2* __restrict__ dL_dencoded, float2* __restrict__ grid_grad)
{
	__shared__ LevelConst lc[NGPB_MAX_LEVELS];
	if (threadIdx.x < L.n_levels) lc[threadIdx.x] = LevelConst{L.scale[threadIdx.x], L.size[threadIdx.x], L.resolution[threadIdx.x], L.offset[threadIdx.x]};
	__syncthreads();
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t level = tid % L.n_levels, i = tid / L.n_levels;
	if (i >= n) return;
	const float2 g = __half22float2(dL_dencoded[(size_t)i * L.n_levels + level]);
	if (g.x == 0.f && g.y == 0.f) return;
	const LevelConst c = lc[level];
	float2* __restrict__ gg = grid_grad + c.offset;
	float pos[2];
	uint32_t pg[2];
	#pragma unroll
	for (int d = 0; d < 2; ++d) {
		const float p = __fmaf_rn(positions[(size_t)i * pos_stride + d], c.scale, 0.5f);
		const float fl = floorf(p);
		pg[d] = (uint32_t)(int)fl;
		pos[d] = p - fl;
	}
	const bool x_only = c.resolution > c.size;
	const bool hashed = (uint64_t)c.size < (uint64_t)c.resolution * (x_only ? 1u : c.resolution);
	#pragma unroll
	for (uint32_t k = 0; k < 4; ++k) {
		const uint32_t x = pg[0] + (k & 1), y = pg[1] + (k >> 1);
		const uint32_t h = hashed ? (x ^ (y * 2654435761u)) : (x_only ? x : x + y * c.resolution);
		const float w = ((k & 1) ? pos[0] : 1.f - pos[0]) * ((k >> 1) ? pos[1] : 1.f - pos[1]);
		float* addr = reinterpret_cast<float*>(gg + h % c.size);
		asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(addr), "f"(g.x * w), "f"(g.y * w) : "memory");
	}
}

// Internal launchers shared with the testbed host.
void hash_encode_forward_launch(cudaStream_t stream, const ngpb_grid* g, const __half* grid, const float* positions, uint32_t pos_stride,
                                uint32_t n, const uint32_t* n_dev, __half* encoded, bool tiled) {
	if (n == 0) return;
	if (tiled && (g->n_levels != 16 || g->n_pos_dims == 2)) throw std::runtime_error("hash_encode_forward: the tiled feature layout needs a 3-D grid of 16 levels");
	const GridLevels L = make_levels(g);
	const uint64_t threads = (uint64_t)n * L.n_levels;
	if (threads > 0xFFFFFFFFull) throw std::runtime_error("hash_encode_forward: n * n_levels must fit 32 bits");
	if (g->n_pos_dims == 2) {
		if (n_dev) throw std::runtime_error("hash_encode_forward: a device-side count is not supported for 2-D grids");
		hash_encode_forward_2d_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, stream>>>(n, L, (const __half2*)grid, positions, pos_stride, (__half2*)encoded);
		NGPB_LAUNCH_CHECK();
		return;
	}
	// no explicit shared-memory carve-out for this kernel: any non-default preference was measured 3x slower (L1 is what feeds the gathers)
	if (L.n_levels == 16 && tiled) hash_encode_forward16_kernel<true><<<(uint32_t)((threads + 255) / 256), 256, 0, stream>>>(n, n_dev, L, (const __half2*)grid, positions, pos_stride, (__half2*)encoded);
	else if (L.n_levels == 16) hash_encode_forward16_kernel<false><<<(uint32_t)((threads + 255) / 256), 256, 0, stream>>>(n, n_dev, L, (const __half2*)grid, positions, pos_stride, (__half2*)encoded);
	else hash_encode_forward_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, stream>>>(n, n_dev, L, (const __half2*)grid, positions, pos_stride, (__half2*)encoded);
	NGPB_LAUNCH_CHECK();
}
// [level_begin, level_end): the levels to scatter (all of them: 0, n_levels); the data-parallel pipeline launches level groups separately
void hash_encode_backward_launch(cudaStream_t stream, const ngpb_grid* g, const float* positions, uint32_t pos_stride, uint32_t n,
                                 const __half* dL_dencoded, float* grid_grad, uint32_t level_begin, uint32_t level_end) {
	if (n == 0 || level_begin >= level_end) return;
	const GridLevels L = make_levels(g);
	if (g->n_pos_dims == 2) {
		const uint64_t threads = (uint64_t)n * L.n_levels;
		if (threads > 0xFFFFFFFFull) throw std::runtime_error("hash_encode_backward: n * n_levels must fit 32 bits");
		if (level_begin != 0 || level_end != L.n_levels) throw std::runtime_error("hash_encode_backward: level groups are not supported for 2-D grids");
		hash_encode_backward_2d_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, stream>>>(n, L, positions, pos_stride, (const __half2*)dL_dencoded, (float2*)grid_grad);
		NGPB_LAUNCH_CHECK();
		return;
	}
	hash_encode_backward_kernel<<<div_round_up(n, ENC_SAMPLES), ENC_SAMPLES * ENC_WARPS, 0, stream>>>(n, L, positions, pos_stride, (const __half2*)dL_dencoded, (float2*)grid_grad, level_begin, level_end);
	NGPB_LAUNCH_CHECK();
}

} // namespace ngpb

using namespace ngpb;

extern "C" uint32_t ngpb_grid_init(ngpb_grid* g, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale) {
	return ngpb_grid_init_nd(g, 3, n_levels, log2_hashmap_size, base_resolution, per_level_scale);
}
extern "C" uint32_t ngpb_grid_init_nd(ngpb_grid* g, uint32_t n_pos_dims, uint32_t n_levels, uint32_t log2_hashmap_size, uint32_t base_resolution, float per_level_scale) {
	if (!g || n_levels == 0 || n_levels > NGPB_MAX_LEVELS || (n_pos_dims != 2 && n_pos_dims != 3) || log2_hashmap_size > 30) return 0;
	g->n_levels = n_levels;
	g->n_pos_dims = n_pos_dims;
	g->base_resolution = base_resolution;
	g->log2_per_level_scale = std::log2(per_level_scale);
	uint32_t offset = 0;
	for (uint32_t l = 0; l < n_levels; ++l) {
		// grid_scale / grid_resolution, grid.h:194-203; level table, grid.h:985-1018
		const float scale = exp2f((float)l * g->log2_per_level_scale) * (float)base_resolution - 1.0f;
		const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
		g->scale[l] = scale;
		g->resolution[l] = resolution;
		const uint32_t max_params = 0xFFFFFFFFu / 2;
		uint32_t params_in_level = powf((float)resolution, (float)n_pos_dims) > (float)max_params ? max_params : (n_pos_dims == 2 ? resolution * resolution : resolution * resolution * resolution);
		params_in_level = next_multiple(params_in_level, 8u);
		params_in_level = params_in_level < (1u << log2_hashmap_size) ? params_in_level : (1u << log2_hashmap_size);
		g->offsets[l] = offset;
		offset += params_in_level;
	}
	g->offsets[n_levels] = offset;
	return offset;
}

// grid_scale as the reference's kernels evaluate it: per thread, with the device's exp2f (grid.h:194-199, called at :241).
__global__ void grid_scales_kernel(const uint32_t n_levels, const float log2_per_level_scale, const uint32_t base_resolution, float* __restrict__ out)
{
	const uint32_t l = threadIdx.x;
	if (l < n_levels) out[l] = exp2f((float)l * log2_per_level_scale) * (float)base_resolution - 1.0f;
}

extern "C" int ngpb_grid_device_scales(void* stream_, ngpb_grid* g) {
	try {
		if (!g || g->n_levels == 0 || g->n_levels > NGPB_MAX_LEVELS) { set_last_error("ngpb_grid_device_scales: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		cudaStream_t stream = (cudaStream_t)stream_;
		float* d = nullptr;
		NGPB_CUDA_CHECK(cudaMalloc(&d, sizeof(float) * NGPB_MAX_LEVELS));
		grid_scales_kernel<<<1, NGPB_MAX_LEVELS, 0, stream>>>(g->n_levels, g->log2_per_level_scale, g->base_resolution, d);
		cudaError_t e = cudaGetLastError();
		if (e == cudaSuccess) e = cudaMemcpyAsync(g->scale, d, sizeof(float) * g->n_levels, cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
		cudaFree(d);
		NGPB_CUDA_CHECK(e);
		for (uint32_t l = 0; l < g->n_levels; ++l) g->resolution[l] = (uint32_t)ceilf(g->scale[l]) + 1;
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_hash_encode_forward(void* stream, const ngpb_grid* g, const ngpb_half* grid, const float* positions, uint32_t pos_stride,
                                        uint32_t n, ngpb_half* encoded) {
	try {
		if (!g || !grid || !positions || !encoded || pos_stride < (g->n_pos_dims == 2 ? 2u : 3u)) { set_last_error("ngpb_hash_encode_forward: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		hash_encode_forward_launch((cudaStream_t)stream, g, (const __half*)grid, positions, pos_stride, n, nullptr, (__half*)encoded, false);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_hash_encode_backward(void* stream, const ngpb_grid* g, const float* positions, uint32_t pos_stride, uint32_t n,
                                         const ngpb_half* dL_dencoded, float* grid_grad) {
	try {
		if (!g || !positions || !dL_dencoded || !grid_grad || pos_stride < (g->n_pos_dims == 2 ? 2u : 3u)) { set_last_error("ngpb_hash_encode_backward: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		hash_encode_backward_launch((cudaStream_t)stream, g, positions, pos_stride, n, (const __half*)dL_dencoded, grid_grad, 0, g->n_levels);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
