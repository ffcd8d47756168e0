// Warp-specialised, software-pipelined NeRF MLP kernels for sm_100a (tcgen05 + TMEM + TMA bulk copies).
// Same arithmetic as nerf_mlp.cu's one-tile-at-a-time kernels (reference: tcnn FullyFusedMLP<__half,64>, dependencies/tiny-cuda-nn/src/
// fully_fused_mlp.cu:151-314 backward, :500-557 forward, :759-850 weight gradients; NerfNetwork glue nerf_network.h:103-266; SH
// spherical_harmonics.h:46-150), different schedule.
//
// Why. A layer of the 64-wide network is a latency chain: MMA (32-128 tensor-pipe cycles) -> commit -> mbarrier -> tcgen05.ld -> ReLU / fp16
// -> shared memory -> fence -> next MMA. With one tile per CTA and every thread walking that chain in lockstep the tensor pipe sat idle
// 80 % of the time (ncu round 1: 15-23 % active, IPC 0.8-1.2). Here a persistent CTA keeps SLOTS tiles in flight at different points of
// the chain, each with its own activation tiles in shared memory and its own accumulator columns in TMEM:
//
//   warps 0 .. 4*SLOTS-1   SLOTS epilogue warpgroups; warpgroup s owns slot s (thread = one sample row = one TMEM lane)
//   issuer warps           one thread each walks (tile iteration, layer) of ITS slots, waits for the slot's operand tile, issues the layer's
//                          tcgen05.mma batch and commits it to the slot's `mma_done` mbarrier. Several issuers, because issuing is not free:
//                          measured on B200 (tools/umma_probe.py) one tcgen05.mma + commit costs the issuing thread ~340 cycles even when
//                          nothing waits on it, ~100 cycles per further MMA of the batch, independent of N -- for these small tiles (8-32
//                          tensor-pipe cycles per instruction) a single issuer caps the tensor pipe at ~10 % (first version of this file).
//   last warp              producer: fetches the 8 KB feature tile of the slot's NEXT sample tile with ONE TMA bulk copy
//                          (cp.async.bulk, completion on the slot's `x_full` mbarrier) as soon as the current tile's last reader is done
//
// so while warpgroup s runs the epilogue of layer l, the tensor pipe executes layer l of the other slots. Epilogue -> issuer and epilogue ->
// producer signals are shared-memory counters (ld.acquire polling, ~30 cycles) rather than mbarriers (try_wait on a completed barrier: ~170 cycles). The hash-grid kernels hand the
// features over in the UMMA core-matrix layout (tile_offset in umma.cuh), 8 KB contiguous per 128 samples, which is what makes the bulk
// copy a single instruction; the row-major layout of the public C ABI is converted by the producer warp with plain loads.
//
// Inference: 6 slots x (8 KB X/Rin + 16 KB hidden) + 20 KB weights = 164 KB, 6 x 64 TMEM columns, 24 epilogue + 3 issuer + 1 producer warps, one CTA per SM.
// Training:  2 slots x 92 KB (all activations of the tile stay resident for the backward pass and the weight-gradient GEMMs, the feature
//            tile is double-buffered) + 20 KB weights = 204 KB; per slot a 64-column accumulator and 160 columns of weight gradients accumulated
//            across all tiles of the slot; 8 epilogue + 2 issuer + 1 producer warps, one CTA per SM.
#include "common.cuh"
#include "umma.cuh"
#include "nerf_mlp_shared.cuh"
#include "../../include/ngpb.h"
#include <cstdio>
#include <vector>

namespace ngpb {
using namespace umma;

// ---- barriers of one slot ---------------------------------------------------------------------------------------------------------------
struct SlotBars {
	uint64_t x_full[2];   // mbarrier, producer -> MMA issuer: feature tile landed (TMA transaction bytes, or the producer warp's arrive)
	uint64_t mma_done;    // mbarrier, MMA issuer (tcgen05.commit) -> epilogue warpgroup: this layer's accumulator is complete
	uint32_t x_empty[2];  // counter, epilogue -> producer: +1 when the last MMA reading the feature buffer has completed
	uint32_t act_ready;   // counter, epilogue -> MMA issuer: +4 (one per warp) per finished epilogue: next operand tile written, accumulator columns free
	uint32_t pad_;
};

__device__ __forceinline__ void signal_act_ready(uint32_t* flag, uint32_t lane) {
	tc_fence_before_sync();     // our tcgen05.ld of the accumulator are ordered before the MMA that overwrites it
	fence_proxy_async_smem();   // our shared-memory writes are visible to the tensor core's (async proxy) reads
	__syncwarp();
	if (lane == 0) flag_signal(flag);
}

// ---- one layer's MMA batch, descriptors reduced to `address >> 4` + constants (see mma_f16_ss_fast) -------------------------------------
// D[128 x N] = A[128 x K] * W[N x K]^T, both K-major; A lives in a tile of A_COLS columns. a16 / w16: shared-memory byte address >> 4.
template <uint32_t A_COLS, uint32_t K, uint32_t N>
__device__ __forceinline__ void issue_fwd(uint32_t acc, uint32_t a16, uint32_t w16) {
	constexpr uint32_t ID = make_idesc_f16(128, N, false, false), LO = desc_lo_const(128), A_HI = desc_hi_const((A_COLS >> 3) * 128), W_HI = desc_hi_const((K >> 3) * 128);
	mma_f16_ss_fast<A_HI, W_HI, ID, 0>(acc, a16 + LO, w16 + LO);
	#pragma unroll
	for (uint32_t k = 1; k < K / 16; ++k) mma_f16_ss_fast<A_HI, W_HI, ID, 1>(acc, a16 + LO + 16 * k, w16 + LO + 16 * k);
}
// dIn[128 x N_IN] = dOut[128 x N_OUT] * W[N_OUT x N_IN]: A K-major in a tile of G_COLS columns, B = the forward weight tile read MN-major
template <uint32_t G_COLS, uint32_t N_IN, uint32_t N_OUT>
__device__ __forceinline__ void issue_dg(uint32_t acc, uint32_t g16, uint32_t w16) {
	constexpr uint32_t ID = make_idesc_f16(128, N_IN, false, true), A_LO = desc_lo_const(128), A_HI = desc_hi_const((G_COLS >> 3) * 128);
	constexpr uint32_t B_LO = desc_lo_const((N_IN >> 3) * 128), B_HI = desc_hi_const(128), B_STEP = 16 * (N_IN >> 3);
	mma_f16_ss_fast<A_HI, B_HI, ID, 0>(acc, g16 + A_LO, w16 + B_LO);
	#pragma unroll
	for (uint32_t k = 1; k < N_OUT / 16; ++k) mma_f16_ss_fast<A_HI, B_HI, ID, 1>(acc, g16 + A_LO + 16 * k, w16 + B_LO + B_STEP * k);
}
// D[64 x N] (+)= P[128 x 64]^T * Q[128 x N]: both tiles read MN-major, K = the 128 sample rows, M = 64
template <uint32_t N, bool FIRST>
__device__ __forceinline__ void issue_wg(uint32_t acc, uint32_t p16, uint32_t q16) {
	constexpr uint32_t ID = make_idesc_f16(64, N, true, true), A_LO = desc_lo_const(8 * 128), B_LO = desc_lo_const((N >> 3) * 128), HI = desc_hi_const(128), B_STEP = 16 * (N >> 3);
	mma_f16_ss_fast<HI, HI, ID, FIRST ? 0 : 1>(acc, p16 + A_LO, q16 + B_LO);
	#pragma unroll
	for (uint32_t k = 1; k < TILE / 16; ++k) mma_f16_ss_fast<HI, HI, ID, 1>(acc, p16 + A_LO + 128 * k, q16 + B_LO + B_STEP * k);
}

// accumulator row (64 fp32 columns) -> ReLU -> fp16 -> row `row` of a [128][64] tile
__device__ __forceinline__ void epi_relu64(uint32_t taddr, uint8_t* tile, uint32_t row) {
	#pragma unroll
	for (uint32_t h = 0; h < 2; ++h) {
		uint32_t r[32];
		tmem_ld_x32(taddr + h * 32, r);
		tmem_ld_wait();
		#pragma unroll
		for (uint32_t c = 0; c < 4; ++c) {
			uint4 v;
			v.x = pack_half2_relu(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1]));
			v.y = pack_half2_relu(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
			v.z = pack_half2_relu(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5]));
			v.w = pack_half2_relu(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
			*reinterpret_cast<uint4*>(tile + tile_offset(row, h * 4 + c, 64)) = v;
		}
	}
}

// data-gradient row (64 fp32) masked by the ReLU of the saved activation (same row of `act`, fp16) -> fp16 -> row of `out`
__device__ __forceinline__ void epi_dgrad64(uint32_t taddr, const uint8_t* act, uint8_t* out, uint32_t row) {
	const __half2 zero = __floats2half2_rn(0.f, 0.f);
	#pragma unroll
	for (uint32_t h = 0; h < 2; ++h) {
		uint32_t r[32];
		tmem_ld_x32(taddr + h * 32, r);
		tmem_ld_wait();
		#pragma unroll
		for (uint32_t c = 0; c < 4; ++c) {
			const uint4 a = *reinterpret_cast<const uint4*>(act + tile_offset(row, h * 4 + c, 64));
			uint4 v;
			v.x = pack_half2_rn(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1])) & __hgt2_mask(*reinterpret_cast<const __half2*>(&a.x), zero);
			v.y = pack_half2_rn(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3])) & __hgt2_mask(*reinterpret_cast<const __half2*>(&a.y), zero);
			v.z = pack_half2_rn(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5])) & __hgt2_mask(*reinterpret_cast<const __half2*>(&a.z), zero);
			v.w = pack_half2_rn(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7])) & __hgt2_mask(*reinterpret_cast<const __half2*>(&a.w), zero);
			*reinterpret_cast<uint4*>(out + tile_offset(row, h * 4 + c, 64)) = v;
		}
	}
}

__device__ __forceinline__ uint4 pack8(const uint32_t* r) {
	uint4 v;
	v.x = pack_half2_rn(__uint_as_float(r[0]), __uint_as_float(r[1])); v.y = pack_half2_rn(__uint_as_float(r[2]), __uint_as_float(r[3]));
	v.z = pack_half2_rn(__uint_as_float(r[4]), __uint_as_float(r[5])); v.w = pack_half2_rn(__uint_as_float(r[6]), __uint_as_float(r[7]));
	return v;
}

// SH degree 4 of the sample's direction -> columns 16..31 of the rgb network's input tile
__device__ __forceinline__ void write_sh(uint8_t* rin, uint32_t row, float dx, float dy, float dz) {
	float sh[16];
	sh4(dx, dy, dz, sh);
	#pragma unroll
	for (uint32_t h = 0; h < 2; ++h) {
		uint4 v;
		v.x = pack_half2(sh[h * 8 + 0], sh[h * 8 + 1]); v.y = pack_half2(sh[h * 8 + 2], sh[h * 8 + 3]);
		v.z = pack_half2(sh[h * 8 + 4], sh[h * 8 + 5]); v.w = pack_half2(sh[h * 8 + 6], sh[h * 8 + 7]);
		*reinterpret_cast<uint4*>(rin + tile_offset(row, 2 + h, 32)) = v;
	}
}

// producer: one 128 x 32 fp16 feature tile into `dst`
__device__ __forceinline__ void fetch_features(const __half* encoded, uint32_t tiled, uint32_t tile, uint8_t* dst, uint64_t* bar, uint32_t lane) {
	const uint8_t* src = reinterpret_cast<const uint8_t*>(encoded) + (size_t)tile * (TILE * N_ENC * 2);
	if (tiled) { // already in core-matrix layout: one TMA bulk copy
		if (lane == 0) {
			mbar_arrive_expect_tx(bar, TILE * N_ENC * 2);
			bulk_copy_g2s(dst, src, TILE * N_ENC * 2, bar);
		}
	} else {     // [n][32] rows of the C ABI: 512 chunks of 16 bytes, re-tiled by the warp
		#pragma unroll 4
		for (uint32_t q = lane; q < TILE * 4; q += 32) {
			*reinterpret_cast<uint4*>(dst + tile_offset(q >> 2, q & 3, 32)) = __ldg(reinterpret_cast<const uint4*>(src) + q);
		}
		fence_proxy_async_smem();
		__syncwarp();
		if (lane == 0) mbar_arrive(bar);
	}
}

// =========================================================================================================================================
// Inference: MODE_DENSITY (density network only), MODE_INFERENCE (density + SH + rgb network), MODE_PLAIN (32 -> 64 -> 64 -> 16 alone)
// =========================================================================================================================================
constexpr uint32_t PI_SLOTS = 6, PI_ISSUERS = 3, PI_SLOTS_PER_ISSUER = PI_SLOTS / PI_ISSUERS;
constexpr uint32_t PI_ISSUER_WARP0 = PI_SLOTS * 4, PI_PRODUCER_WARP = PI_ISSUER_WARP0 + PI_ISSUERS;
constexpr uint32_t PI_THREADS = (PI_PRODUCER_WARP + 1) * 32;
constexpr uint32_t PI_XR = 0, PI_H = 8192, PI_SLOT_BYTES = 24576; // XR: features, later the rgb network's input [density out 16 | SH 16]
constexpr uint32_t PI_SLOT0 = SW_END;
constexpr uint32_t PI_CTRL = PI_SLOT0 + PI_SLOTS * PI_SLOT_BYTES;
constexpr uint32_t PI_SMEM = PI_CTRL + PI_SLOTS * (uint32_t)sizeof(SlotBars) + 16;
constexpr uint32_t PI_TMEM_COLS = 512; // 6 x 64 used; allocations are powers of two

struct PipeInferArgs {
	const __half* mlp; const __half* encoded; const float* coords; __half* out;
	uint32_t n; const uint32_t* n_dev; uint32_t tiled;
	long long* trace; // development aid (NGPB_PIPE_TRACE=path): CTA 0 records (role, slot, iteration, step, clock) events
};
#define NGPB_TRACE(role, slot_, it_, step_) do { if (args.trace && blockIdx.x == 0 && (it_) < 3) { args.trace[((((role) * 6 + (slot_)) * 3 + (it_)) * 8 + (step_))] = clock64(); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(PI_THREADS, 1) nerf_mlp_pipe_infer_kernel(const PipeInferArgs args)
{
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr uint32_t NSTEPS = MODE == MODE_DENSITY ? 2 : (MODE == MODE_INFERENCE ? 5 : 3);
	constexpr uint32_t X_FREE_STEP = MODE == MODE_INFERENCE ? 2 : 0; // the last step whose MMAs read XR
	SlotBars* bars = reinterpret_cast<SlotBars*>(smem + PI_CTRL);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PI_CTRL + PI_SLOTS * sizeof(SlotBars));
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	uint32_t n = args.n;
	if (args.n_dev) n = min(n, (*args.n_dev + TILE - 1) / TILE * TILE);
	const uint32_t n_tiles = n / TILE;
	const uint32_t n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u; // tiles blockIdx.x, + gridDim.x, ...

	if (warp == PI_ISSUER_WARP0) tmem_alloc<PI_TMEM_COLS>(tmem_slot);
	if (tid == 0) {
		for (uint32_t s = 0; s < PI_SLOTS; ++s) {
			mbar_init(&bars[s].x_full[0], 1); mbar_init(&bars[s].x_full[1], 1); mbar_init(&bars[s].mma_done, 1);
			bars[s].x_empty[0] = bars[s].x_empty[1] = bars[s].act_ready = 0u;
		}
		fence_mbar_init();
	}
	if (MODE == MODE_PLAIN) { // FullyFusedMLP parameter order: [64][32], [64][64], [16][64]
		load_matrix_to_tile(smem, SW_W1R, args.mlp, 64, 32);
		load_matrix_to_tile(smem, SW_W2R, args.mlp + 2048, 64, 64);
		load_matrix_to_tile(smem, SW_W3R, args.mlp + 6144, 16, 64);
	} else {
		load_weights(smem, args.mlp);
	}
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t sbase = smem_u32(smem);

	if (warp < PI_SLOTS * 4) {
		// ------------------------------------------------ epilogue warpgroup of slot s ------------------------------------------------
		const uint32_t s = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
		uint8_t* slot = smem + PI_SLOT0 + s * PI_SLOT_BYTES;
		SlotBars& B = bars[s];
		const uint32_t t_acc = tmem_base + ((wq * 32u) << 16) + s * 64u;
		uint32_t n_done = 0; // completions of mma_done consumed so far
		for (uint32_t k = s; k < n_my; k += PI_SLOTS) {
			const uint32_t tile = blockIdx.x + k * gridDim.x;
			const size_t row_g = (size_t)tile * TILE + row;
			float dx = 0.f, dy = 0.f, dz = 0.f;
			if (MODE == MODE_INFERENCE) { const float* c = args.coords + row_g * COORD_FLOATS; dx = c[4]; dy = c[5]; dz = c[6]; }
			float sigma_logit = 0.f;
			#pragma unroll
			for (uint32_t step = 0; step < NSTEPS; ++step) {
				mbar_wait_bounded(&B.mma_done, n_done & 1u); ++n_done;
				tc_fence_after_sync();
				if (wq == 0 && lane == 0) NGPB_TRACE(0, s, k / PI_SLOTS, step);
				if (step == X_FREE_STEP && wq == 0 && lane == 0) flag_signal(&B.x_empty[0]);
				const bool hidden = MODE == MODE_PLAIN ? step < 2 : (step == 0 || step == 2 || step == 3);
				if (hidden) {
					epi_relu64(t_acc, slot + PI_H, row);
				} else if (MODE != MODE_PLAIN && step == 1) { // density network output: 16 columns, no activation; column 0 is the density logit
					uint32_t r[16];
					tmem_ld_x16(t_acc, r);
					tmem_ld_wait();
					sigma_logit = __uint_as_float(r[0]);
					if (MODE == MODE_DENSITY) {
						args.out[row_g] = __float2half_rn(sigma_logit);
					} else {
						*reinterpret_cast<uint4*>(slot + PI_XR + tile_offset(row, 0, 32)) = pack8(r);
						*reinterpret_cast<uint4*>(slot + PI_XR + tile_offset(row, 1, 32)) = pack8(r + 8);
						write_sh(slot + PI_XR, row, dx, dy, dz);
					}
				} else if (MODE == MODE_PLAIN) {               // all 16 padded outputs ([n][16] fp16, what FullyFusedMLP writes)
					uint32_t r[16];
					tmem_ld_x16(t_acc, r);
					tmem_ld_wait();
					*reinterpret_cast<uint4*>(args.out + row_g * 16) = pack8(r);
					*reinterpret_cast<uint4*>(args.out + row_g * 16 + 8) = pack8(r + 8);
				} else {                                       // rgb network output: {r, g, b} + the density logit (nerf_network.h:128-136)
					uint32_t r[4];
					tmem_ld_x4(t_acc, r);
					tmem_ld_wait();
					uint2 o;
					o.x = pack_half2(__uint_as_float(r[0]), __uint_as_float(r[1]));
					o.y = pack_half2(__uint_as_float(r[2]), __half2float(__float2half_rn(sigma_logit)));
					*reinterpret_cast<uint2*>(args.out + row_g * 4) = o;
				}
				signal_act_ready(&B.act_ready, lane);
				if (wq == 0 && lane == 0) NGPB_TRACE(1, s, k / PI_SLOTS, step);
			}
		}
	} else if (warp < PI_PRODUCER_WARP) {
		// ------------------------------------------------ MMA issuer of slots [s0, s0 + PI_SLOTS_PER_ISSUER) ------------------------------------------------
		// (the whole warp walks the loop with warp-uniform values; tcgen05.mma / commit are issued by one elected lane, see mma_f16_ss_fast)
		{
			const uint32_t s0 = (warp - PI_ISSUER_WARP0) * PI_SLOTS_PER_ISSUER;
			const uint32_t n_iters = (n_my + PI_SLOTS - 1) / PI_SLOTS;
			const uint32_t w16 = sbase >> 4;
			for (uint32_t it = 0; it < n_iters; ++it) {
				#pragma unroll
				for (uint32_t step = 0; step < NSTEPS; ++step) {
					#pragma unroll
					for (uint32_t q = 0; q < PI_SLOTS_PER_ISSUER; ++q) {
						const uint32_t s = s0 + q;
						if (it * PI_SLOTS + s >= n_my) continue;
						SlotBars& B = bars[s];
						if (step == 0) mbar_wait_bounded(&B.x_full[0], it & 1u);
						const uint32_t n_epilogues = it * NSTEPS + step; // epilogues of this slot that must have finished (the previous tile's last one for step 0)
						if (n_epilogues > 0) flag_wait_bounded(&B.act_ready, 4u * n_epilogues);
						__syncwarp();
						tc_fence_after_sync();
						if (lane == 0) NGPB_TRACE(2, s, it, step);
						const uint32_t acc = tmem_base + s * 64u;
						const uint32_t xr = (sbase + PI_SLOT0 + s * PI_SLOT_BYTES + PI_XR) >> 4, hh = (sbase + PI_SLOT0 + s * PI_SLOT_BYTES + PI_H) >> 4;
						if (MODE == MODE_PLAIN) {
							if (step == 0) issue_fwd<32, 32, 64>(acc, xr, w16 + (SW_W1R >> 4));
							else if (step == 1) issue_fwd<64, 64, 64>(acc, hh, w16 + (SW_W2R >> 4));
							else issue_fwd<64, 64, 16>(acc, hh, w16 + (SW_W3R >> 4));
						} else {
							if (step == 0) issue_fwd<32, 32, 64>(acc, xr, w16 + (SW_W1D >> 4));
							else if (step == 1) issue_fwd<64, 64, 16>(acc, hh, w16 + (SW_W2D >> 4));
							else if (step == 2) issue_fwd<32, 32, 64>(acc, xr, w16 + (SW_W1R >> 4));
							else if (step == 3) issue_fwd<64, 64, 64>(acc, hh, w16 + (SW_W2R >> 4));
							else issue_fwd<64, 64, 16>(acc, hh, w16 + (SW_W3R >> 4));
						}
						mma_commit_elect(&B.mma_done);
						if (lane == 0) NGPB_TRACE(3, s, it, step);
					}
				}
			}
		}
	} else {
		// ------------------------------------------------ producer ------------------------------------------------
		for (uint32_t k = 0; k < n_my; ++k) {
			const uint32_t s = k % PI_SLOTS, it = k / PI_SLOTS;
			if (it > 0) flag_wait_bounded(&bars[s].x_empty[0], it);
			fetch_features(args.encoded, args.tiled, blockIdx.x + k * gridDim.x, smem + PI_SLOT0 + s * PI_SLOT_BYTES + PI_XR, &bars[s].x_full[0], lane);
		}
	}

	tc_fence_before_sync();
	__syncthreads();
	if (warp == PI_ISSUER_WARP0) { __syncwarp(); tmem_dealloc<PI_TMEM_COLS>(tmem_base); }
}

// =========================================================================================================================================
// Training: forward + data gradients + weight gradients in one pass. MODE_TRAIN (NeRF networks), MODE_PLAIN_TRAIN (32 -> 64 -> 64 -> 16)
// =========================================================================================================================================
constexpr uint32_t PT_SLOTS = 2;
constexpr uint32_t PT_ISSUER_WARP0 = PT_SLOTS * 4, PT_PRODUCER_WARP = PT_ISSUER_WARP0 + PT_SLOTS; // one issuer per slot
constexpr uint32_t PT_THREADS = (PT_PRODUCER_WARP + 1) * 32;
// per-slot tiles (bytes). dG1 is written after G2's last use, dH1 after dG2's, dOd after dOr's (see the step list below).
constexpr uint32_t PT_X0 = 0, PT_X1 = 8192, PT_H1 = 16384, PT_RIN = 32768, PT_G1 = 40960, PT_G2 = 57344, PT_DG1 = PT_G2, PT_DO = 73728, PT_DG2 = 77824, PT_DH1 = PT_DG2;
constexpr uint32_t PT_SLOT_BYTES = 94208;
constexpr uint32_t PT_SLOT0 = SW_END;
constexpr uint32_t PT_CTRL = PT_SLOT0 + PT_SLOTS * PT_SLOT_BYTES;
constexpr uint32_t PT_SMEM = PT_CTRL + PT_SLOTS * (uint32_t)sizeof(SlotBars) + 32;
// TMEM columns: one 64-column accumulator per slot, then per slot 160 columns of weight gradients (M = 64 accumulators). Each slot has its own issuer
// thread, and two threads must not accumulate into the same columns; the two partial sums are added by the fixed-order reduction kernel.
constexpr uint32_t PT_ACC = 0, PT_DW = 128, PT_DW_COLS = 160, PT_DW1D = 0, PT_DW2D = 32, PT_DW1R = 48, PT_DW2R = 80, PT_DW3R = 144, PT_TMEM_COLS = 512;

struct PipeTrainArgs {
	const __half* mlp; const __half* encoded; const float* coords; const __half* dL_dout; __half* dL_dencoded; float* partials;
	uint32_t n; uint32_t tiled;
	__half* dL_dsh; // optional [n][16]: gradient of the loss with respect to the rgb network's 16 SH inputs (camera-extrinsics optimisation needs dL/d(direction))
};

template <int MODE>
__global__ void __launch_bounds__(PT_THREADS, 1) nerf_mlp_pipe_train_kernel(const PipeTrainArgs args)
{
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr bool PLAIN = MODE == MODE_PLAIN_TRAIN;
	constexpr uint32_t NSTEPS = PLAIN ? 5 : 9;
	SlotBars* bars = reinterpret_cast<SlotBars*>(smem + PT_CTRL);
	uint64_t* final_bar = reinterpret_cast<uint64_t*>(smem + PT_CTRL + PT_SLOTS * sizeof(SlotBars));
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + PT_CTRL + PT_SLOTS * sizeof(SlotBars) + 16);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const uint32_t n_tiles = args.n / TILE;
	const uint32_t n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

	if (warp == PT_ISSUER_WARP0) tmem_alloc<PT_TMEM_COLS>(tmem_slot);
	if (tid == 0) {
		for (uint32_t s = 0; s < PT_SLOTS; ++s) {
			mbar_init(&bars[s].x_full[0], 1); mbar_init(&bars[s].x_full[1], 1); mbar_init(&bars[s].mma_done, 1);
			bars[s].x_empty[0] = bars[s].x_empty[1] = bars[s].act_ready = 0u;
		}
		mbar_init(final_bar, PT_SLOTS);
		fence_mbar_init();
	}
	if (PLAIN) {
		load_matrix_to_tile(smem, SW_W1R, args.mlp, 64, 32);
		load_matrix_to_tile(smem, SW_W2R, args.mlp + 2048, 64, 64);
		load_matrix_to_tile(smem, SW_W3R, args.mlp + 6144, 16, 64);
	} else {
		load_weights(smem, args.mlp);
	}
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t sbase = smem_u32(smem);

	if (warp < PT_SLOTS * 4) {
		// ------------------------------------------------ epilogue warpgroup of slot s ------------------------------------------------
		const uint32_t s = warp >> 2, wq = warp & 3, row = wq * 32 + lane;
		uint8_t* slot = smem + PT_SLOT0 + s * PT_SLOT_BYTES;
		SlotBars& B = bars[s];
		const uint32_t t_acc = tmem_base + ((wq * 32u) << 16) + PT_ACC + s * 64u;
		uint32_t n_done = 0, it = 0;
		auto wait_mma = [&]() { mbar_wait_bounded(&B.mma_done, n_done & 1u); ++n_done; tc_fence_after_sync(); };
		// The previous tile's feature buffer is last read by that tile's final weight-gradient batch, which is issued AFTER its step's commit; the first
		// commit of this tile covers it, so the buffer goes back to the producer here (called right after wait_mma of step 0).
		auto release_prev_x = [&]() { if (it > 0 && wq == 0 && lane == 0) flag_signal(&B.x_empty[(it - 1u) & 1u]); };
		for (uint32_t k = s; k < n_my; k += PT_SLOTS, ++it) {
			const uint32_t tile = blockIdx.x + k * gridDim.x;
			const size_t row_g = (size_t)tile * TILE + row;
			if (PLAIN) {
				// dL/d(output), all 16 padded columns as the caller's loss kernel wrote them. DO's last reader (the previous tile's B1) is long done.
				const uint4* g = reinterpret_cast<const uint4*>(args.dL_dout + row_g * 16);
				*reinterpret_cast<uint4*>(slot + PT_DO + tile_offset(row, 0, 16)) = __ldg(g);
				*reinterpret_cast<uint4*>(slot + PT_DO + tile_offset(row, 1, 16)) = __ldg(g + 1);
				wait_mma(); release_prev_x(); epi_relu64(t_acc, slot + PT_G1, row); signal_act_ready(&B.act_ready, lane);   // 0: G1 = relu(X W1^T)
				wait_mma(); epi_relu64(t_acc, slot + PT_G2, row); signal_act_ready(&B.act_ready, lane);                      // 1: G2 = relu(G1 W2^T)
				wait_mma(); epi_dgrad64(t_acc, slot + PT_G2, slot + PT_DG2, row); signal_act_ready(&B.act_ready, lane);      // 2: dG2 = (dO W3) . relu'(G2)
				wait_mma(); epi_dgrad64(t_acc, slot + PT_G1, slot + PT_DG1, row); signal_act_ready(&B.act_ready, lane);      // 3: dG1 = (dG2 W2) . relu'(G1)
				wait_mma();                                                                                                   // 4: dX = dG1 W1 -> HBM
				{
					uint32_t r[32];
					tmem_ld_x32(t_acc, r);
					tmem_ld_wait();
					uint4* dst = reinterpret_cast<uint4*>(args.dL_dencoded + row_g * N_ENC);
					#pragma unroll
					for (uint32_t c = 0; c < 4; ++c) dst[c] = pack8(r + c * 8);
				}
				signal_act_ready(&B.act_ready, lane);
				continue;
			}
			const float* cd = args.coords + row_g * COORD_FLOATS;
			const float dx = cd[4], dy = cd[5], dz = cd[6];
			const uint2 g = __ldg(reinterpret_cast<const uint2*>(args.dL_dout + row_g * 4));
			const float dsigma = __high2float(*reinterpret_cast<const __half2*>(&g.y));
			wait_mma(); release_prev_x(); epi_relu64(t_acc, slot + PT_H1, row); signal_act_ready(&B.act_ready, lane);        // 0: H1 = relu(X W1d^T)
			wait_mma();                                                                                                       // 1: Od = H1 W2d^T -> Rin[:, :16]; SH -> Rin[:, 16:]; dOr
			{
				uint32_t r[16];
				tmem_ld_x16(t_acc, r);
				tmem_ld_wait();
				*reinterpret_cast<uint4*>(slot + PT_RIN + tile_offset(row, 0, 32)) = pack8(r);
				*reinterpret_cast<uint4*>(slot + PT_RIN + tile_offset(row, 1, 32)) = pack8(r + 8);
				write_sh(slot + PT_RIN, row, dx, dy, dz);
				// dL/d(rgb out) = first three components, the other 13 padded outputs get zero (nerf_network.h:202-206)
				*reinterpret_cast<uint4*>(slot + PT_DO + tile_offset(row, 0, 16)) = make_uint4(g.x, g.y & 0x0000FFFFu, 0u, 0u);
				*reinterpret_cast<uint4*>(slot + PT_DO + tile_offset(row, 1, 16)) = make_uint4(0u, 0u, 0u, 0u);
			}
			signal_act_ready(&B.act_ready, lane);
			wait_mma(); epi_relu64(t_acc, slot + PT_G1, row); signal_act_ready(&B.act_ready, lane);                           // 2: G1 = relu(Rin W1r^T)
			wait_mma(); epi_relu64(t_acc, slot + PT_G2, row); signal_act_ready(&B.act_ready, lane);                           // 3: G2 = relu(G1 W2r^T)
			wait_mma(); epi_dgrad64(t_acc, slot + PT_G2, slot + PT_DG2, row); signal_act_ready(&B.act_ready, lane);           // 4: dG2 = (dOr W3r) . relu'(G2)
			wait_mma(); epi_dgrad64(t_acc, slot + PT_G1, slot + PT_DG1, row); signal_act_ready(&B.act_ready, lane);           // 5: dG1 = (dG2 W2r) . relu'(G1)
			wait_mma();                                                                                                       // 6: dRin = dG1 W1r; dOd = dRin[:, :16] (+ dL/dsigma on column 0, nerf_network.h:232-239)
			if (args.dL_dsh) { // columns 16..31 of dRin: dL/d(SH coefficients), kept for kernel_sh_backward's counterpart (camera_optimizer.cu)
				uint32_t r[16];
				tmem_ld_x16(t_acc + 16, r);
				tmem_ld_wait();
				*reinterpret_cast<uint4*>(args.dL_dsh + row_g * 16) = pack8(r);
				*reinterpret_cast<uint4*>(args.dL_dsh + row_g * 16 + 8) = pack8(r + 8);
			}
			{
				uint32_t r[16];
				tmem_ld_x16(t_acc, r);
				tmem_ld_wait();
				r[0] = __float_as_uint(__half2float(__float2half_rn(__uint_as_float(r[0]))) + dsigma); // half + half, rounded again by the pack
				*reinterpret_cast<uint4*>(slot + PT_DO + tile_offset(row, 0, 16)) = pack8(r);
				*reinterpret_cast<uint4*>(slot + PT_DO + tile_offset(row, 1, 16)) = pack8(r + 8);
			}
			signal_act_ready(&B.act_ready, lane);
			wait_mma(); epi_dgrad64(t_acc, slot + PT_H1, slot + PT_DH1, row); signal_act_ready(&B.act_ready, lane);           // 7: dH1 = (dOd W2d) . relu'(H1)
			wait_mma();                                                                                                       // 8: dX = dH1 W1d -> dL/dencoded (HBM)
			{
				uint32_t r[32];
				tmem_ld_x32(t_acc, r);
				tmem_ld_wait();
				uint4* dst = reinterpret_cast<uint4*>(args.dL_dencoded + row_g * N_ENC);
				#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) dst[c] = pack8(r + c * 8);
			}
			signal_act_ready(&B.act_ready, lane);
		}
		// ---- this slot's weight-gradient partial (set 2 * blockIdx.x + s), written once every MMA of the CTA has completed. M = 64 accumulators
		// occupy lanes 0-15 of every 32-lane quadrant: warp w, lane l < 16 holds row 16 w + l. ----
		{
			mbar_wait_bounded(final_bar, 0u);
			tc_fence_after_sync();
			const uint32_t t_row = tmem_base + ((wq * 32u) << 16) + PT_DW + s * PT_DW_COLS;
			const uint32_t wrow = wq * 16 + lane;
			const bool have = n_my > s; // a slot that never got a tile contributes zeros
			uint32_t r[32];
			if (PLAIN) {
				float* part = args.partials + (size_t)(blockIdx.x * PT_SLOTS + s) * PLAIN_PARAMS;
				tmem_ld_x32(t_row + PT_DW1R, r); tmem_ld_wait();
				if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[PLAIN_W1 + wrow * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
				#pragma unroll
				for (uint32_t h = 0; h < 2; ++h) {
					tmem_ld_x32(t_row + PT_DW2R + h * 32, r); tmem_ld_wait();
					if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[PLAIN_W2 + wrow * 64 + h * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
				}
				tmem_ld_x16(t_row + PT_DW3R, r); tmem_ld_wait();
				if (lane < 16) { for (uint32_t o = 0; o < 16; ++o) part[PLAIN_W3 + o * 64 + wrow] = have ? __uint_as_float(r[o]) : 0.f; }
			} else {
				float* part = args.partials + (size_t)(blockIdx.x * PT_SLOTS + s) * MLP_PARAMS;
				tmem_ld_x32(t_row + PT_DW1D, r); tmem_ld_wait();   // dW1d[o][i]
				if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[MLP_W1D + wrow * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
				tmem_ld_x16(t_row + PT_DW2D, r); tmem_ld_wait();   // dW2d^T[i][o]
				if (lane < 16) { for (uint32_t o = 0; o < 16; ++o) part[MLP_W2D + o * 64 + wrow] = have ? __uint_as_float(r[o]) : 0.f; }
				tmem_ld_x32(t_row + PT_DW1R, r); tmem_ld_wait();   // dW1r[o][i]
				if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[MLP_W1R + wrow * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
				#pragma unroll
				for (uint32_t h = 0; h < 2; ++h) {                 // dW2r[o][i]
					tmem_ld_x32(t_row + PT_DW2R + h * 32, r); tmem_ld_wait();
					if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[MLP_W2R + wrow * 64 + h * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
				}
				tmem_ld_x16(t_row + PT_DW3R, r); tmem_ld_wait();   // dW3r^T[i][o]
				if (lane < 16) { for (uint32_t o = 0; o < 16; ++o) part[MLP_W3R + o * 64 + wrow] = have ? __uint_as_float(r[o]) : 0.f; }
			}
		}
	} else if (warp < PT_PRODUCER_WARP) {
		// ------------------------------------------------ MMA issuer of slot s (whole warp, one elected lane issues) ------------------------------------------------
		{
			const uint32_t s = warp - PT_ISSUER_WARP0;
			SlotBars& B = bars[s];
			const uint32_t acc = tmem_base + PT_ACC + s * 64u, dw = tmem_base + PT_DW + s * PT_DW_COLS;
			const uint32_t w16 = sbase >> 4, sl = (sbase + PT_SLOT0 + s * PT_SLOT_BYTES) >> 4;
			const uint32_t W1D = w16 + (SW_W1D >> 4), W2D = w16 + (SW_W2D >> 4), W1R = w16 + (SW_W1R >> 4), W2R = w16 + (SW_W2R >> 4), W3R = w16 + (SW_W3R >> 4);
			const uint32_t H1 = sl + (PT_H1 >> 4), RIN = sl + (PT_RIN >> 4), G1 = sl + (PT_G1 >> 4), G2 = sl + (PT_G2 >> 4), DG1 = sl + (PT_DG1 >> 4), DO = sl + (PT_DO >> 4),
			               DG2 = sl + (PT_DG2 >> 4), DH1 = sl + (PT_DH1 >> 4);
			uint32_t it = 0;
			for (uint32_t k = s; k < n_my; k += PT_SLOTS, ++it) {
				const uint32_t xb = it & 1u;
				const uint32_t X = sl + ((xb ? PT_X1 : PT_X0) >> 4);
				#pragma unroll
				for (uint32_t step = 0; step < NSTEPS; ++step) {
					if (step == 0) mbar_wait_bounded(&B.x_full[xb], (it >> 1) & 1u);
					const uint32_t n_epilogues = it * NSTEPS + step;
					if (n_epilogues > 0) flag_wait_bounded(&B.act_ready, 4u * n_epilogues);
					__syncwarp();
					tc_fence_after_sync();
					// The commit sits between the data-gradient batch and the weight-gradient batch: the epilogue only needs the former, and runs while the
					// tensor pipe works through the eight dependent K-steps of the latter. The NEXT step's commit also covers this step's weight-gradient
					// MMAs (a commit tracks everything its thread issued before), which is what the aliased tiles need before they are overwritten.
					// The first tile of the slot overwrites its weight-gradient accumulators, later tiles add to them.
					#define NGPB_WG(N, DST, P, Q) do { mma_commit_elect(&B.mma_done); if (it == 0) issue_wg<N, true>(dw + DST, P, Q); else issue_wg<N, false>(dw + DST, P, Q); } while (0)
					if (PLAIN) {
						if (step == 0) issue_fwd<32, 32, 64>(acc, X, W1R);                                  // G1 = relu(X W1^T)
						else if (step == 1) issue_fwd<64, 64, 64>(acc, G1, W2R);                            // G2 = relu(G1 W2^T)
						else if (step == 2) { issue_dg<16, 64, 16>(acc, DO, W3R); NGPB_WG(16, PT_DW3R, G2, DO); }    // dG2; dW3^T += G2^T dO
						else if (step == 3) { issue_dg<64, 64, 64>(acc, DG2, W2R); NGPB_WG(64, PT_DW2R, DG2, G1); }  // dG1; dW2 += dG2^T G1
						else { issue_dg<64, 32, 64>(acc, DG1, W1R); NGPB_WG(32, PT_DW1R, DG1, X); }                  // dX;  dW1 += dG1^T X
					} else {
						if (step == 0) issue_fwd<32, 32, 64>(acc, X, W1D);                                  // H1 = relu(X W1d^T)
						else if (step == 1) issue_fwd<64, 64, 16>(acc, H1, W2D);                            // Od = H1 W2d^T
						else if (step == 2) issue_fwd<32, 32, 64>(acc, RIN, W1R);                           // G1 = relu(Rin W1r^T)
						else if (step == 3) issue_fwd<64, 64, 64>(acc, G1, W2R);                            // G2 = relu(G1 W2r^T)
						else if (step == 4) { issue_dg<16, 64, 16>(acc, DO, W3R); NGPB_WG(16, PT_DW3R, G2, DO); }    // dG2 = (dOr W3r) . relu'(G2);  dW3r^T += G2^T dOr
						else if (step == 5) { issue_dg<64, 64, 64>(acc, DG2, W2R); NGPB_WG(64, PT_DW2R, DG2, G1); }  // dG1 = (dG2 W2r) . relu'(G1);  dW2r += dG2^T G1
						else if (step == 6) { issue_dg<64, 32, 64>(acc, DG1, W1R); NGPB_WG(32, PT_DW1R, DG1, RIN); } // dRin = dG1 W1r;               dW1r += dG1^T Rin
						else if (step == 7) { issue_dg<16, 64, 16>(acc, DO, W2D); NGPB_WG(16, PT_DW2D, H1, DO); }    // dH1 = (dOd W2d) . relu'(H1);  dW2d^T += H1^T dOd
						else { issue_dg<64, 32, 64>(acc, DH1, W1D); NGPB_WG(32, PT_DW1D, DH1, X); }                  // dX = dH1 W1d;                 dW1d += dH1^T X
					}
					#undef NGPB_WG
					if (step < (PLAIN ? 2u : 4u)) mma_commit_elect(&B.mma_done); // forward steps: no weight-gradient batch, commit here
				}
			}
			mma_commit_elect(final_bar); // arrives once every MMA this thread issued has completed: with both issuers' arrivals the weight gradients are final
		}
	} else {
		// ------------------------------------------------ producer (feature tiles, one tile ahead per slot) ------------------------------------------------
		for (uint32_t k = 0; k < n_my; ++k) {
			const uint32_t s = k % PT_SLOTS, it = k / PT_SLOTS, xb = it & 1u;
			if (it >= 2) flag_wait_bounded(&bars[s].x_empty[xb], it >> 1);
			fetch_features(args.encoded, args.tiled, blockIdx.x + k * gridDim.x, smem + PT_SLOT0 + s * PT_SLOT_BYTES + (xb ? PT_X1 : PT_X0), &bars[s].x_full[xb], lane);
		}
	}

	tc_fence_before_sync();
	__syncthreads();
	if (warp == PT_ISSUER_WARP0) { __syncwarp(); tmem_dealloc<PT_TMEM_COLS>(tmem_base); }
}

// ---- launchers --------------------------------------------------------------------------------------------------------------------------
template <int MODE>
static void launch_pipe_infer(cudaStream_t stream, const PipeInferArgs& a) {
	static bool configured = false;
	if (!configured) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(nerf_mlp_pipe_infer_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PI_SMEM));
		configured = true;
	}
	const uint32_t tiles = a.n / TILE;
	nerf_mlp_pipe_infer_kernel<MODE><<<std::min(tiles, kNumSMs), PI_THREADS, PI_SMEM, stream>>>(a);
	NGPB_LAUNCH_CHECK();
}
template <int MODE>
static uint32_t launch_pipe_train(cudaStream_t stream, const PipeTrainArgs& a) {
	static bool configured = false;
	if (!configured) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(nerf_mlp_pipe_train_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM));
		configured = true;
	}
	const uint32_t grid = std::min(a.n / TILE, kNumSMs);
	nerf_mlp_pipe_train_kernel<MODE><<<grid, PT_THREADS, PT_SMEM, stream>>>(a);
	NGPB_LAUNCH_CHECK();
	return grid * PT_SLOTS; // partial sets written (one per slot)
}

static long long* g_trace_dev = nullptr;
void pipe_nerf_forward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma) {
	static const char* trace_path = getenv("NGPB_PIPE_TRACE");
	static const bool force_tiled = getenv("NGPB_PIPE_FORCE_TILED") != nullptr; // timing experiments through the row-major C ABI: the layout does not change the arithmetic's cost
	if (force_tiled) tiled = true;
	if (trace_path && !g_trace_dev) { NGPB_CUDA_CHECK(cudaMalloc(&g_trace_dev, 4 * 6 * 3 * 8 * sizeof(long long))); }
	if (g_trace_dev) NGPB_CUDA_CHECK(cudaMemsetAsync(g_trace_dev, 0, 4 * 6 * 3 * 8 * sizeof(long long), stream));
	launch_pipe_infer<MODE_INFERENCE>(stream, PipeInferArgs{mlp, encoded, coords, rgbsigma, n, n_dev, tiled ? 1u : 0u, g_trace_dev});
	if (g_trace_dev) {
		std::vector<long long> h(4 * 6 * 3 * 8);
		NGPB_CUDA_CHECK(cudaStreamSynchronize(stream));
		NGPB_CUDA_CHECK(cudaMemcpy(h.data(), g_trace_dev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
		if (FILE* f = fopen(trace_path, "w")) { for (size_t i = 0; i < h.size(); ++i) fprintf(f, "%lld\n", h[i]); fclose(f); }
	}
}
void pipe_density_forward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, uint32_t n, __half* density) {
	launch_pipe_infer<MODE_DENSITY>(stream, PipeInferArgs{mlp, encoded, nullptr, density, n, nullptr, tiled ? 1u : 0u, nullptr});
}
void pipe_plain_forward(cudaStream_t stream, const __half* weights, const __half* input, uint32_t n, __half* output) {
	launch_pipe_infer<MODE_PLAIN>(stream, PipeInferArgs{weights, input, nullptr, output, n, nullptr, 0u, nullptr});
}
uint32_t pipe_nerf_forward_backward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, const __half* dL_dout, uint32_t n,
                                    __half* dL_dencoded, float* partials, __half* dL_dsh) {
	return launch_pipe_train<MODE_TRAIN>(stream, PipeTrainArgs{mlp, encoded, coords, dL_dout, dL_dencoded, partials, n, tiled ? 1u : 0u, dL_dsh});
}
uint32_t pipe_plain_forward_backward(cudaStream_t stream, const __half* weights, const __half* input, const __half* dL_dout16, uint32_t n, __half* dL_dinput, float* partials) {
	return launch_pipe_train<MODE_PLAIN_TRAIN>(stream, PipeTrainArgs{weights, input, nullptr, dL_dout16, dL_dinput, partials, n, 0u, nullptr});
}

} // namespace ngpb
