// NeRF MLPs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Replaces tcnn FullyFusedMLP<__half,64> forward/backward (reference:
// dependencies/tiny-cuda-nn/src/fully_fused_mlp.cu:151-314 backward, :500-557 forward, :759-850 weight
// gradients through CUTLASS split-K GEMMs) together with the NerfNetwork glue kernels around it
// (include/neural-graphics-primitives/nerf_network.h:32-74 extract_density / extract_rgb /
// add_density_gradient, :103-266) and the SH direction encoding (spherical_harmonics.h:46-150).
//
// Shape of the computation. A CTA owns tiles of 128 samples (= UMMA M, one TMEM lane per sample,
// one thread per sample row in the epilogues). Every layer is one batch of tcgen05.mma K=16 steps
// issued by a single thread: D[128 x N] (+)= A[128 x K] * W[N x K]^T with A = the previous layer's
// fp16 activations in shared memory and W = the fp16 weights in shared memory. The epilogue reads
// the fp32 accumulator row with tcgen05.ld, applies ReLU, rounds to fp16 and writes the row back to
// shared memory as the next layer's A operand. Activations never leave the SM.
//
// Training (forward + backward in ONE kernel): the activation tiles stay in shared memory, the
// data-gradient GEMMs reuse the same weight tiles as MN-major B operands (no transposed copy), and
// the five weight-gradient GEMMs dW = dOut^T * Act (K = the 128 samples of the tile) read the
// activation / gradient tiles as MN-major operands and ACCUMULATE IN TMEM ACROSS ALL TILES of the
// CTA -- the reference writes activations to HBM, re-reads them, and runs five split-K CUTLASS
// GEMMs on side streams instead. Each CTA writes one fp32 partial per weight at the end; a second
// tiny kernel sums the partials in a fixed order (deterministic).
//
// Numerics: fp16 operands, fp32 accumulation (the reference's wmma path accumulates in fp16,
// fully_fused_mlp.cu:66-68), fp16 rounding at the same layer boundaries as the reference.
#include "common.cuh"
#include "umma.cuh"
#include "nerf_mlp_shared.cuh"
#include "../../include/ngpb.h"

namespace ngpb {
using namespace umma;

// Epilogue: accumulator row (64 fp32 columns) -> optional ReLU -> fp16 -> row `row` of a [128][64] tile.
template <bool RELU>
__device__ __forceinline__ void epilogue_store64(uint32_t taddr, uint8_t* smem, uint32_t tile, uint32_t row) {
	#pragma unroll
	for (uint32_t h = 0; h < 2; ++h) {
		uint32_t r[32];
		tmem_ld_x32(taddr + h * 32, r);
		tmem_ld_wait();
		#pragma unroll
		for (uint32_t c = 0; c < 4; ++c) {
			uint4 v;
			uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
			#pragma unroll
			for (uint32_t e = 0; e < 4; ++e) {
				float a = __uint_as_float(r[c * 8 + e * 2]), b = __uint_as_float(r[c * 8 + e * 2 + 1]);
				if (RELU) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
				pv[e] = pack_half2(a, b);
			}
			*reinterpret_cast<uint4*>(smem + tile + tile_offset(row, h * 4 + c, 64)) = v;
		}
	}
}

// Data-gradient epilogue: accumulator row (64 fp32) masked by the ReLU of the saved forward activation
// (same row of `act_tile`, fp16) -> fp16 -> row of `out_tile`.
__device__ __forceinline__ void epilogue_dgrad64(uint32_t taddr, uint8_t* smem, uint32_t act_tile, uint32_t out_tile, uint32_t row) {
	#pragma unroll
	for (uint32_t h = 0; h < 2; ++h) {
		uint32_t r[32];
		tmem_ld_x32(taddr + h * 32, r);
		tmem_ld_wait();
		#pragma unroll
		for (uint32_t c = 0; c < 4; ++c) {
			const uint4 act = *reinterpret_cast<const uint4*>(smem + act_tile + tile_offset(row, h * 4 + c, 64));
			const uint32_t* pa = reinterpret_cast<const uint32_t*>(&act);
			uint4 v;
			uint32_t* pv = reinterpret_cast<uint32_t*>(&v);
			#pragma unroll
			for (uint32_t e = 0; e < 4; ++e) {
				const __half2 ah = *reinterpret_cast<const __half2*>(&pa[e]);
				const float a = __low2float(ah) > 0.f ? __uint_as_float(r[c * 8 + e * 2]) : 0.f;
				const float b = __high2float(ah) > 0.f ? __uint_as_float(r[c * 8 + e * 2 + 1]) : 0.f;
				pv[e] = pack_half2(a, b);
			}
			*reinterpret_cast<uint4*>(smem + out_tile + tile_offset(row, h * 4 + c, 64)) = v;
		}
	}
}

struct MlpArgs {
	const __half* mlp;
	const __half* encoded;   // [n][32]
	const float* coords;     // [n][7]
	const __half* dL_dout;   // [n][4]             (train)
	__half* out;             // rgbsigma [n][4] (inference) | density [n] (density) | dL_dencoded [n][32] (train)
	float* partials;         // [gridDim.x][MLP_PARAMS] (train)
	uint32_t n;              // multiple of 128
	const uint32_t* n_dev;   // optional device-side sample count (rounded up to 128 by the kernel)
};

template <int MODE>
__global__ void __launch_bounds__(128, (MODE == 2 || MODE == 4) ? 2 : 4) nerf_mlp_kernel(const MlpArgs args)
{
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr bool TRAINING = MODE == MODE_TRAIN || MODE == MODE_PLAIN_TRAIN, PLAIN_NET = MODE == MODE_PLAIN || MODE == MODE_PLAIN_TRAIN;
	constexpr uint32_t TILES_END = TRAINING ? S_TRAIN_END : S_INFER_END;
	uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + TILES_END);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TILES_END + 16);
	constexpr uint32_t TM_COLS = TRAINING ? TM_COLS_TRAIN : TM_COLS_INFER;

	const uint32_t tid = threadIdx.x, warp = tid >> 5;
	uint32_t n = args.n;
	if (args.n_dev) n = min(n, (*args.n_dev + TILE - 1) / TILE * TILE);
	const uint32_t n_tiles = n / TILE;

	if (warp == 0) tmem_alloc<TM_COLS>(tmem_slot);
	if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
	if (PLAIN_NET) { // FullyFusedMLP parameter order: first layer [64][32], hidden [64][64], last [16][64]
		load_matrix_to_tile(smem, SW_W1R, args.mlp, 64, 32);
		load_matrix_to_tile(smem, SW_W2R, args.mlp + 2048, 64, 64);
		load_matrix_to_tile(smem, SW_W3R, args.mlp + 6144, 16, 64);
	} else {
		load_weights(smem, args.mlp);
	}
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t t_row = tmem_base + ((warp * 32u) << 16); // this thread's TMEM lane = its sample row
	const uint32_t sbase = smem_u32(smem);
	uint32_t phase = 0;

	// one MMA batch: make this CTA's smem writes visible to the async proxy, let thread 0 issue, wait for completion
	#define NGPB_MMA_BATCH(ISSUE)                                 \
		do {                                                      \
			fence_proxy_async_smem();                             \
			tc_fence_before_sync();                               \
			__syncthreads();                                      \
			if (tid == 0) { tc_fence_after_sync(); ISSUE; mma_commit(mbar); } \
			__syncwarp();                                         \
			mbar_wait(mbar, phase);                               \
			phase ^= 1;                                           \
			tc_fence_after_sync();                                \
		} while (0)

	uint32_t it = 0;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
		const size_t s0 = (size_t)tile * TILE;   // first sample of the tile
		const size_t row_g = s0 + tid;            // this thread's sample

		// ---- stage inputs: X tile (coalesced 16-B chunks), SH -> Rin[:,16:32], dL/dout -> dOr ----
		{
			const uint4* src = reinterpret_cast<const uint4*>(args.encoded + s0 * N_ENC);
			#pragma unroll
			for (uint32_t k = 0; k < 4; ++k) {
				const uint32_t q = tid + 128 * k;
				*reinterpret_cast<uint4*>(smem + (PLAIN_NET ? S_RIN : S_X) + tile_offset(q >> 2, q & 3, 32)) = __ldg(src + q);
			}
		}
		float dsigma = 0.f;
		if (MODE == MODE_INFERENCE || MODE == MODE_TRAIN) {
			const float* c = args.coords + row_g * COORD_FLOATS;
			float sh[16];
			sh4(c[4], c[5], c[6], sh);
			#pragma unroll
			for (uint32_t h = 0; h < 2; ++h) {
				uint4 v;
				v.x = pack_half2(sh[h * 8 + 0], sh[h * 8 + 1]); v.y = pack_half2(sh[h * 8 + 2], sh[h * 8 + 3]);
				v.z = pack_half2(sh[h * 8 + 4], sh[h * 8 + 5]); v.w = pack_half2(sh[h * 8 + 6], sh[h * 8 + 7]);
				*reinterpret_cast<uint4*>(smem + S_RIN + tile_offset(tid, 2 + h, 32)) = v;
			}
		}
		if (MODE == MODE_PLAIN_TRAIN) { // dL/d(output), all 16 padded columns as the caller's loss kernel wrote them
			const uint4* gsrc = reinterpret_cast<const uint4*>(args.dL_dout + row_g * 16);
			*reinterpret_cast<uint4*>(smem + S_DOR + tile_offset(tid, 0, 16)) = __ldg(gsrc);
			*reinterpret_cast<uint4*>(smem + S_DOR + tile_offset(tid, 1, 16)) = __ldg(gsrc + 1);
		}
		if (MODE == MODE_TRAIN) {
			// dL/d(rgb out) = first three components, other 13 padded outputs get zero (nerf_network.h:202-206)
			const uint2 g = *reinterpret_cast<const uint2*>(args.dL_dout + row_g * 4);
			const __half2 g23 = *reinterpret_cast<const __half2*>(&g.y);
			dsigma = __high2float(g23);
			uint4 v0 = {g.x, g.y & 0x0000FFFFu, 0u, 0u}, v1 = {0u, 0u, 0u, 0u};
			*reinterpret_cast<uint4*>(smem + S_DOR + tile_offset(tid, 0, 16)) = v0;
			*reinterpret_cast<uint4*>(smem + S_DOR + tile_offset(tid, 1, 16)) = v1;
		}

		float sigma_logit = 0.f;
		if (!PLAIN_NET) {
		// ---- density net layer 1: H1 = relu(X W1d^T) ----
		NGPB_MMA_BATCH(issue_forward(tmem_base + TM_ACC, sbase + S_X, 32, sbase + SW_W1D, 32, 64));
		epilogue_store64<true>(t_row + TM_ACC, smem, S_H1, tid);

		// ---- density net layer 2: Od = H1 W2d^T (16 outputs, no activation) ----
		NGPB_MMA_BATCH(issue_forward(tmem_base + TM_ACC, sbase + S_H1, 64, sbase + SW_W2D, 64, 16));
		{
			uint32_t r[16];
			tmem_ld_x16(t_row + TM_ACC, r);
			tmem_ld_wait();
			sigma_logit = __uint_as_float(r[0]);
			if (MODE == MODE_DENSITY) {
				args.out[row_g] = __float2half_rn(sigma_logit);
			} else {
				#pragma unroll
				for (uint32_t h = 0; h < 2; ++h) {
					uint4 v;
					v.x = pack_half2(__uint_as_float(r[h * 8 + 0]), __uint_as_float(r[h * 8 + 1])); v.y = pack_half2(__uint_as_float(r[h * 8 + 2]), __uint_as_float(r[h * 8 + 3]));
					v.z = pack_half2(__uint_as_float(r[h * 8 + 4]), __uint_as_float(r[h * 8 + 5])); v.w = pack_half2(__uint_as_float(r[h * 8 + 6]), __uint_as_float(r[h * 8 + 7]));
					*reinterpret_cast<uint4*>(smem + S_RIN + tile_offset(tid, h, 32)) = v;
				}
			}
		}
		}
		if (MODE == MODE_DENSITY) { tc_fence_before_sync(); continue; }

		constexpr uint32_t T_G1 = TRAINING ? S_G1 : S_H1;
		constexpr uint32_t T_G2 = TRAINING ? S_G2 : S_H1;

		// ---- rgb net layer 1: G1 = relu(Rin W1r^T) ----
		NGPB_MMA_BATCH(issue_forward(tmem_base + TM_ACC, sbase + S_RIN, 32, sbase + SW_W1R, 32, 64));
		epilogue_store64<true>(t_row + TM_ACC, smem, T_G1, tid);

		// ---- rgb net layer 2: G2 = relu(G1 W2r^T) ----
		NGPB_MMA_BATCH(issue_forward(tmem_base + TM_ACC, sbase + T_G1, 64, sbase + SW_W2R, 64, 64));
		epilogue_store64<true>(t_row + TM_ACC, smem, T_G2, tid);

		if (MODE == MODE_PLAIN) {
			// ---- output layer: all 16 padded outputs, no activation ([n][16] fp16, what FullyFusedMLP writes) ----
			NGPB_MMA_BATCH(issue_forward(tmem_base + TM_ACC, sbase + T_G2, 64, sbase + SW_W3R, 64, 16));
			uint32_t r[16];
			tmem_ld_x16(t_row + TM_ACC, r);
			tmem_ld_wait();
			#pragma unroll
			for (uint32_t h = 0; h < 2; ++h) {
				uint4 v;
				v.x = pack_half2(__uint_as_float(r[h * 8 + 0]), __uint_as_float(r[h * 8 + 1])); v.y = pack_half2(__uint_as_float(r[h * 8 + 2]), __uint_as_float(r[h * 8 + 3]));
				v.z = pack_half2(__uint_as_float(r[h * 8 + 4]), __uint_as_float(r[h * 8 + 5])); v.w = pack_half2(__uint_as_float(r[h * 8 + 6]), __uint_as_float(r[h * 8 + 7]));
				*reinterpret_cast<uint4*>(args.out + row_g * 16 + h * 8) = v;
			}
			tc_fence_before_sync();
			continue;
		}

		if (MODE == MODE_INFERENCE) {
			// ---- rgb net layer 3: 16 padded outputs, 3 used; output {r,g,b,sigma} (nerf_network.h:128-136) ----
			NGPB_MMA_BATCH(issue_forward(tmem_base + TM_ACC, sbase + T_G2, 64, sbase + SW_W3R, 64, 16));
			uint32_t r[4];
			tmem_ld_x4(t_row + TM_ACC, r);
			tmem_ld_wait();
			uint2 o;
			o.x = pack_half2(__uint_as_float(r[0]), __uint_as_float(r[1]));
			o.y = pack_half2(__uint_as_float(r[2]), __half2float(__float2half_rn(sigma_logit)));
			*reinterpret_cast<uint2*>(args.out + row_g * 4) = o;
			tc_fence_before_sync();
			continue;
		}

		if (TRAINING) {
			const bool first = it == 0;
			// ---- dG2 = (dOr W3r) . relu'(G2);  dW3r^T += G2^T dOr ----
			NGPB_MMA_BATCH(
				issue_dgrad(tmem_base + TM_ACC, sbase + S_DOR, 16, sbase + SW_W3R, 64, 16);
				issue_wgrad(tmem_base + TM_DW3R, sbase + S_G2, sbase + S_DOR, 16, 16, first));
			epilogue_dgrad64(t_row + TM_ACC, smem, S_G2, S_DG2, tid);

			// ---- dG1 = (dG2 W2r) . relu'(G1);  dW2r += dG2^T G1 ----
			NGPB_MMA_BATCH(
				issue_dgrad(tmem_base + TM_ACC, sbase + S_DG2, 64, sbase + SW_W2R, 64, 64);
				issue_wgrad(tmem_base + TM_DW2R, sbase + S_DG2, sbase + S_G1, 64, 64, first));
			epilogue_dgrad64(t_row + TM_ACC, smem, S_G1, S_DG1, tid);

			// ---- dRin = dG1 W1r;  dW1r += dG1^T Rin.  dOd = dRin[:, :16] (+ dL/dsigma on column 0, nerf_network.h:232-239) ----
			NGPB_MMA_BATCH(
				issue_dgrad(tmem_base + TM_ACC, sbase + S_DG1, 64, sbase + SW_W1R, 32, 64);
				issue_wgrad(tmem_base + TM_DW1R, sbase + S_DG1, sbase + S_RIN, 32, 32, first));
			if (MODE == MODE_PLAIN_TRAIN) { // dL/d(input), 32 columns, is this network's last product
				uint32_t r[32];
				tmem_ld_x32(t_row + TM_ACC, r);
				tmem_ld_wait();
				uint4* dst = reinterpret_cast<uint4*>(args.out + row_g * N_ENC);
				#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) {
					uint4 v;
					v.x = pack_half2(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1])); v.y = pack_half2(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
					v.z = pack_half2(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5])); v.w = pack_half2(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
					dst[c] = v;
				}
				tc_fence_before_sync();
				continue;
			}
			{
				uint32_t r[16];
				tmem_ld_x16(t_row + TM_ACC, r);
				tmem_ld_wait();
				const float d0 = __half2float(__float2half_rn(__uint_as_float(r[0]))) + dsigma; // half + half, rounded again below
				r[0] = __float_as_uint(d0);
				#pragma unroll
				for (uint32_t h = 0; h < 2; ++h) {
					uint4 v;
					v.x = pack_half2(__uint_as_float(r[h * 8 + 0]), __uint_as_float(r[h * 8 + 1])); v.y = pack_half2(__uint_as_float(r[h * 8 + 2]), __uint_as_float(r[h * 8 + 3]));
					v.z = pack_half2(__uint_as_float(r[h * 8 + 4]), __uint_as_float(r[h * 8 + 5])); v.w = pack_half2(__uint_as_float(r[h * 8 + 6]), __uint_as_float(r[h * 8 + 7]));
					*reinterpret_cast<uint4*>(smem + S_DOD + tile_offset(tid, h, 16)) = v;
				}
			}

			// ---- dH1 = (dOd W2d) . relu'(H1);  dW2d^T += H1^T dOd ----
			NGPB_MMA_BATCH(
				issue_dgrad(tmem_base + TM_ACC, sbase + S_DOD, 16, sbase + SW_W2D, 64, 16);
				issue_wgrad(tmem_base + TM_DW2D, sbase + S_H1, sbase + S_DOD, 16, 16, first));
			epilogue_dgrad64(t_row + TM_ACC, smem, S_H1, S_DH1, tid);

			// ---- dX = dH1 W1d -> dL/dencoded (HBM);  dW1d += dH1^T X ----
			NGPB_MMA_BATCH(
				issue_dgrad(tmem_base + TM_ACC, sbase + S_DH1, 64, sbase + SW_W1D, 32, 64);
				issue_wgrad(tmem_base + TM_DW1D, sbase + S_DH1, sbase + S_X, 32, 32, first));
			{
				uint32_t r[32];
				tmem_ld_x32(t_row + TM_ACC, r);
				tmem_ld_wait();
				uint4* dst = reinterpret_cast<uint4*>(args.out + row_g * N_ENC);
				#pragma unroll
				for (uint32_t c = 0; c < 4; ++c) {
					uint4 v;
					v.x = pack_half2(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1])); v.y = pack_half2(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3]));
					v.z = pack_half2(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5])); v.w = pack_half2(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7]));
					dst[c] = v;
				}
			}
			tc_fence_before_sync();
		}
	}

	if (MODE == MODE_PLAIN_TRAIN) { // same accumulator layout as below, FullyFusedMLP's parameter order
		float* part = args.partials + (size_t)blockIdx.x * PLAIN_PARAMS;
		const uint32_t lane = tid & 31, row = warp * 16 + lane;
		const bool have = it > 0;
		tc_fence_after_sync();
		uint32_t r[32];
		tmem_ld_x32(t_row + TM_DW1R, r); tmem_ld_wait();
		if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[PLAIN_W1 + row * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
		#pragma unroll
		for (uint32_t h = 0; h < 2; ++h) {
			tmem_ld_x32(t_row + TM_DW2R + h * 32, r); tmem_ld_wait();
			if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[PLAIN_W2 + row * 64 + h * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
		}
		tmem_ld_x16(t_row + TM_DW3R, r); tmem_ld_wait();
		if (lane < 16) { for (uint32_t o = 0; o < 16; ++o) part[PLAIN_W3 + o * 64 + row] = have ? __uint_as_float(r[o]) : 0.f; }
	}
	if (MODE == MODE_TRAIN) {
		// ---- write this CTA's weight-gradient partial. M = 64 accumulators occupy lanes 0-15 of every
		// 32-lane quadrant: warp w, lane l < 16 holds row 16*w + l. ----
		float* part = args.partials + (size_t)blockIdx.x * MLP_PARAMS;
		const uint32_t lane = tid & 31, row = warp * 16 + lane;
		const bool have = it > 0; // a CTA without tiles contributes zeros
		tc_fence_after_sync();
		uint32_t r[32];
		// dW1d[o][i]
		tmem_ld_x32(t_row + TM_DW1D, r); tmem_ld_wait();
		if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[MLP_W1D + row * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
		// dW2d^T[i][o]
		tmem_ld_x16(t_row + TM_DW2D, r); tmem_ld_wait();
		if (lane < 16) { for (uint32_t o = 0; o < 16; ++o) part[MLP_W2D + o * 64 + row] = have ? __uint_as_float(r[o]) : 0.f; }
		// dW1r[o][i]
		tmem_ld_x32(t_row + TM_DW1R, r); tmem_ld_wait();
		if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[MLP_W1R + row * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
		// dW2r[o][i]
		#pragma unroll
		for (uint32_t h = 0; h < 2; ++h) {
			tmem_ld_x32(t_row + TM_DW2R + h * 32, r); tmem_ld_wait();
			if (lane < 16) { for (uint32_t i = 0; i < 32; ++i) part[MLP_W2R + row * 64 + h * 32 + i] = have ? __uint_as_float(r[i]) : 0.f; }
		}
		// dW3r^T[i][o]
		tmem_ld_x16(t_row + TM_DW3R, r); tmem_ld_wait();
		if (lane < 16) { for (uint32_t o = 0; o < 16; ++o) part[MLP_W3R + o * 64 + row] = have ? __uint_as_float(r[o]) : 0.f; }
	}

	tc_fence_before_sync();
	__syncthreads();
	if (warp == 0) tmem_dealloc<TM_COLS>(tmem_base);
	#undef NGPB_MMA_BATCH
}

// Fixed-order sum of the per-CTA partials (deterministic), overwrites mlp_grad.
// Sums the per-CTA weight-gradient partials in a fixed order: 32 parameters x 8 groups per block, group g adds partials g, g+8, g+16, ... in
// sequence (four independent chains in flight), then the 8 group sums are added in order 0..7. Same result on every run.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, const uint32_t n_parts, float* __restrict__ mlp_grad, const uint32_t n_weights)
{
	__shared__ float part[8][32];
	const uint32_t col = threadIdx.x & 31, grp = threadIdx.x >> 5;
	const uint32_t i = blockIdx.x * 32 + col;
	float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
	if (i < n_weights) {
		uint32_t p = grp;
		for (; p + 24 < n_parts; p += 32) {
			s0 += partials[(size_t)p * n_weights + i];
			s1 += partials[(size_t)(p + 8) * n_weights + i];
			s2 += partials[(size_t)(p + 16) * n_weights + i];
			s3 += partials[(size_t)(p + 24) * n_weights + i];
		}
		for (; p < n_parts; p += 8) s0 += partials[(size_t)p * n_weights + i];
	}
	part[grp][col] = (s0 + s1) + (s2 + s3);
	__syncthreads();
	if (grp == 0 && i < n_weights) {
		float s = part[0][col];
		#pragma unroll
		for (int g = 1; g < 8; ++g) s += part[g][col];
		mlp_grad[i] = s;
	}
}

constexpr uint32_t SMEM_INFER = S_INFER_END + S_CTRL;
constexpr uint32_t SMEM_TRAIN = S_TRAIN_END + S_CTRL;
constexpr uint32_t TRAIN_CTAS_PER_SM = 2;
constexpr uint32_t TRAIN_GRID = kNumSMs * TRAIN_CTAS_PER_SM; // two CTAs per SM (104 KB of shared memory and 256 TMEM columns each)
constexpr uint32_t INFER_CTAS_PER_SM = 4;

template <int MODE>
static void launch_mlp(cudaStream_t stream, const MlpArgs& a, uint32_t grid, uint32_t smem_bytes) {
	static bool configured = false;
	if (!configured) {
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(nerf_mlp_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
		// the training kernel overlaps the sampling stream (see common.cuh); the inference kernels run alone and want 4 CTAs x 53 KB
		if (MODE == MODE_TRAIN && step_carveout() >= 0) NGPB_CUDA_CHECK(cudaFuncSetAttribute(nerf_mlp_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, step_carveout()));
		configured = true;
	}
	nerf_mlp_kernel<MODE><<<grid, 128, smem_bytes, stream>>>(a);
	NGPB_LAUNCH_CHECK();
}

// Pipelined kernels (nerf_mlp_pipe.cu). The one-tile-at-a-time kernels above remain selectable with NGPB_MLP_LEGACY=1 for A/B timing; they read row-major features only.
void pipe_nerf_forward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma);
void pipe_density_forward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, uint32_t n, __half* density);
void pipe_plain_forward(cudaStream_t stream, const __half* weights, const __half* input, uint32_t n, __half* output);
uint32_t pipe_nerf_forward_backward(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, const __half* dL_dout, uint32_t n,
                                    __half* dL_dencoded, float* partials, __half* dL_dsh);
uint32_t pipe_plain_forward_backward(cudaStream_t stream, const __half* weights, const __half* input, const __half* dL_dout16, uint32_t n, __half* dL_dinput, float* partials);

bool mlp_legacy() { static const bool v = getenv("NGPB_MLP_LEGACY") && atoi(getenv("NGPB_MLP_LEGACY")) != 0; return v; }
// Layout of the [n][32] fp16 hash-grid features handed from the encoding kernels to the MLP kernels inside the library: per 128 samples one 8 KB block in
// the UMMA core-matrix layout (umma.cuh tile_offset), which the pipelined kernels fetch with one TMA bulk copy. The public C ABI stays row-major.
bool features_tiled() { return !mlp_legacy(); }

// Internal launchers shared with the testbed host (device-side sample count supported). `tiled`: layout of `encoded` (see features_tiled()).
void nerf_mlp_forward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, uint32_t n, const uint32_t* n_dev, __half* rgbsigma) {
	if (!mlp_legacy()) { pipe_nerf_forward(stream, mlp, encoded, tiled, coords, n, n_dev, rgbsigma); return; }
	if (tiled) throw std::runtime_error("NGPB_MLP_LEGACY kernels read row-major features");
	MlpArgs a{mlp, encoded, coords, nullptr, rgbsigma, nullptr, n, n_dev};
	const uint32_t tiles = n / TILE;
	launch_mlp<MODE_INFERENCE>(stream, a, std::min(tiles, kNumSMs * INFER_CTAS_PER_SM), SMEM_INFER);
}
void plain_mlp_launch(cudaStream_t stream, const __half* weights, const __half* input, uint32_t n, __half* output) {
	if (!mlp_legacy()) { pipe_plain_forward(stream, weights, input, n, output); return; }
	MlpArgs a{weights, input, nullptr, nullptr, output, nullptr, n, nullptr};
	const uint32_t tiles = n / TILE;
	launch_mlp<MODE_PLAIN>(stream, a, std::min(tiles, kNumSMs * INFER_CTAS_PER_SM), SMEM_INFER);
}
void plain_mlp_forward_backward_launch(cudaStream_t stream, const __half* weights, const __half* input, const __half* dL_dout16, uint32_t n, __half* dL_dinput, float* grad,
                                       float* partials) {
	uint32_t grid;
	if (!mlp_legacy()) {
		grid = pipe_plain_forward_backward(stream, weights, input, dL_dout16, n, dL_dinput, partials);
	} else {
		MlpArgs a{weights, input, nullptr, dL_dout16, dL_dinput, partials, n, nullptr};
		grid = std::min(n / TILE, TRAIN_GRID);
		launch_mlp<MODE_PLAIN_TRAIN>(stream, a, grid, SMEM_TRAIN);
	}
	reduce_partials_kernel<<<div_round_up(PLAIN_PARAMS, 32), 256, 0, stream>>>(partials, grid, grad, PLAIN_PARAMS);
	NGPB_LAUNCH_CHECK();
}
void nerf_density_mlp_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, uint32_t n, __half* density) {
	if (!mlp_legacy()) { pipe_density_forward(stream, mlp, encoded, tiled, n, density); return; }
	if (tiled) throw std::runtime_error("NGPB_MLP_LEGACY kernels read row-major features");
	MlpArgs a{mlp, encoded, nullptr, nullptr, density, nullptr, n, nullptr};
	const uint32_t tiles = n / TILE;
	launch_mlp<MODE_DENSITY>(stream, a, std::min(tiles, kNumSMs * INFER_CTAS_PER_SM), SMEM_INFER);
}
void nerf_mlp_forward_backward_launch(cudaStream_t stream, const __half* mlp, const __half* encoded, bool tiled, const float* coords, const __half* dL_dout, uint32_t n,
                                      __half* dL_dencoded, float* mlp_grad, float* partials, __half* dL_dsh) {
	uint32_t grid;
	if (!mlp_legacy()) {
		grid = pipe_nerf_forward_backward(stream, mlp, encoded, tiled, coords, dL_dout, n, dL_dencoded, partials, dL_dsh);
	} else {
		if (tiled || dL_dsh) throw std::runtime_error("NGPB_MLP_LEGACY kernels read row-major features and do not return the direction gradient");
		MlpArgs a{mlp, encoded, coords, dL_dout, dL_dencoded, partials, n, nullptr};
		grid = std::min(n / TILE, TRAIN_GRID);
		launch_mlp<MODE_TRAIN>(stream, a, grid, SMEM_TRAIN);
	}
	NGPB_STEP_KERNEL(reduce_partials_kernel);
	reduce_partials_kernel<<<div_round_up(MLP_PARAMS, 32), 256, 0, stream>>>(partials, grid, mlp_grad, MLP_PARAMS);
	NGPB_LAUNCH_CHECK();
}

// ---- tcgen05 building-block self-test (tests/test_umma_selftest.py) -------------------------------
// variant 0: D[128x64] = A[128x32] * B[64x32]^T           (K-major A, K-major B: forward layer)
// variant 1: D[128x32] = A[128x64] * B[64x32]             (K-major A, MN-major B: data gradient)
// variant 2: D[64x32]  = A[128x64]^T * B[128x32]          (MN-major A and B, M = 64: weight gradient)
__global__ void __launch_bounds__(128) umma_selftest_kernel(const int variant, const __half* __restrict__ a, const __half* __restrict__ b, float* __restrict__ d)
{
	extern __shared__ __align__(128) uint8_t smem[];
	constexpr uint32_t OFF_A = 0, OFF_B = 16384, OFF_CTRL = 32768;
	uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_CTRL);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_CTRL + 16);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (warp == 0) tmem_alloc<64>(tmem_slot);
	if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
	const uint32_t a_rows = 128, a_cols = variant == 0 ? 32 : 64;
	const uint32_t b_rows = variant == 2 ? 128 : 64, b_cols = 32;
	load_matrix_to_tile(smem, OFF_A, a, a_rows, a_cols);
	load_matrix_to_tile(smem, OFF_B, b, b_rows, b_cols);
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();
	const uint32_t tmem_base = *tmem_slot;
	const uint32_t sbase = smem_u32(smem);
	if (tid == 0) {
		if (variant == 0) issue_forward(tmem_base, sbase + OFF_A, 32, sbase + OFF_B, 32, 64);
		else if (variant == 1) issue_dgrad(tmem_base, sbase + OFF_A, 64, sbase + OFF_B, 32, 64);
		else issue_wgrad(tmem_base, sbase + OFF_A, sbase + OFF_B, 32, 32, true);
		mma_commit(mbar);
	}
	__syncwarp();
	mbar_wait(mbar, 0);
	tc_fence_after_sync();
	const uint32_t t_row = tmem_base + ((warp * 32u) << 16);
	uint32_t r[32];
	if (variant == 0) {
		for (uint32_t h = 0; h < 2; ++h) {
			tmem_ld_x32(t_row + h * 32, r); tmem_ld_wait();
			for (uint32_t i = 0; i < 32; ++i) d[(size_t)tid * 64 + h * 32 + i] = __uint_as_float(r[i]);
		}
	} else if (variant == 1) {
		tmem_ld_x32(t_row, r); tmem_ld_wait();
		for (uint32_t i = 0; i < 32; ++i) d[(size_t)tid * 32 + i] = __uint_as_float(r[i]);
	} else {
		tmem_ld_x32(t_row, r); tmem_ld_wait();
		if (lane < 16) for (uint32_t i = 0; i < 32; ++i) d[(size_t)(warp * 16 + lane) * 32 + i] = __uint_as_float(r[i]);
	}
	tc_fence_before_sync();
	__syncthreads();
	if (warp == 0) tmem_dealloc<64>(tmem_base);
}

} // namespace ngpb

using namespace ngpb;

extern "C" int ngpb_nerf_mlp_forward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, uint32_t n, ngpb_half* rgbsigma) {
	try {
		if (!mlp || !encoded || !coords || !rgbsigma || n % TILE != 0) { set_last_error("ngpb_nerf_mlp_forward: invalid argument (n must be a multiple of 128)"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (n == 0) return 0;
		nerf_mlp_forward_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, coords, n, nullptr, (__half*)rgbsigma);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_mlp_forward(void* stream, const ngpb_half* weights, const ngpb_half* input, uint32_t n, ngpb_half* output) {
	try {
		if (!weights || !input || !output || n == 0 || n % TILE != 0) { set_last_error("ngpb_mlp_forward: null pointer or n not a non-zero multiple of 128"); return NGPB_ERR_INVALID_ARGUMENT; }
		plain_mlp_launch((cudaStream_t)stream, (const __half*)weights, (const __half*)input, n, (__half*)output);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_mlp_forward_backward(void* stream, const ngpb_half* weights, const ngpb_half* input, const ngpb_half* dL_dout, uint32_t n, ngpb_half* dL_dinput,
                                         float* grad, void* workspace) {
	try {
		if (!weights || !input || !dL_dout || !dL_dinput || !grad || !workspace || n == 0 || n % TILE != 0) {
			set_last_error("ngpb_mlp_forward_backward: null pointer or n not a non-zero multiple of 128");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		plain_mlp_forward_backward_launch((cudaStream_t)stream, (const __half*)weights, (const __half*)input, (const __half*)dL_dout, n, (__half*)dL_dinput, grad, (float*)workspace);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_nerf_density_mlp_forward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, uint32_t n, ngpb_half* density) {
	try {
		if (!mlp || !encoded || !density || n % TILE != 0) { set_last_error("ngpb_nerf_density_mlp_forward: invalid argument (n must be a multiple of 128)"); return NGPB_ERR_INVALID_ARGUMENT; }
		if (n == 0) return 0;
		nerf_density_mlp_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, n, (__half*)density);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" uint64_t ngpb_nerf_mlp_workspace_bytes(void) { return (uint64_t)TRAIN_GRID * MLP_PARAMS * sizeof(float); }

extern "C" int ngpb_nerf_mlp_forward_backward(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, const ngpb_half* dL_dout,
                                              uint32_t n, ngpb_half* dL_dencoded, float* mlp_grad, void* workspace) {
	try {
		if (!mlp || !encoded || !coords || !dL_dout || !dL_dencoded || !mlp_grad || !workspace || n == 0 || n % TILE != 0) {
			set_last_error("ngpb_nerf_mlp_forward_backward: invalid argument (n must be a non-zero multiple of 128)");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		nerf_mlp_forward_backward_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, coords, (const __half*)dL_dout, n,
			(__half*)dL_dencoded, mlp_grad, (float*)workspace, nullptr);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

// The same pass, also returning dL/d(SH inputs of the rgb network) [n][16] half: the direction half of the input gradient (camera_optimizer.cu)
extern "C" int ngpb_nerf_mlp_forward_backward_sh(void* stream, const ngpb_half* mlp, const ngpb_half* encoded, const float* coords, const ngpb_half* dL_dout,
                                                 uint32_t n, ngpb_half* dL_dencoded, float* mlp_grad, void* workspace, ngpb_half* dL_dsh) {
	try {
		if (!mlp || !encoded || !coords || !dL_dout || !dL_dencoded || !mlp_grad || !workspace || !dL_dsh || n == 0 || n % TILE != 0) {
			set_last_error("ngpb_nerf_mlp_forward_backward_sh: invalid argument (n must be a non-zero multiple of 128)");
			return NGPB_ERR_INVALID_ARGUMENT;
		}
		nerf_mlp_forward_backward_launch((cudaStream_t)stream, (const __half*)mlp, (const __half*)encoded, false, coords, (const __half*)dL_dout, n,
			(__half*)dL_dencoded, mlp_grad, (float*)workspace, (__half*)dL_dsh);
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}

extern "C" int ngpb_selftest_umma(void* stream, int variant, const ngpb_half* a, const ngpb_half* b, float* d) {
	try {
		if (variant < 0 || variant > 2 || !a || !b || !d) { set_last_error("ngpb_selftest_umma: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		umma_selftest_kernel<<<1, 128, 32768 + 64, (cudaStream_t)stream>>>(variant, (const __half*)a, (const __half*)b, d);
		NGPB_LAUNCH_CHECK();
		return 0;
	} catch (const std::exception& e) { set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
