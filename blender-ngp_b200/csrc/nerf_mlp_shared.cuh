// Pieces shared by the NeRF MLP kernels (nerf_mlp.cu: one tile per CTA at a time; nerf_mlp_pipe.cu: warp-specialised pipeline):
// shared-memory map of the weights, tile loaders, tcgen05.mma batches of one layer, the SH direction encoding.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace ngpb {
using namespace umma;

constexpr uint32_t TILE = 128;

// ---- shared-memory map (bytes) --------------------------------------------------------------------
// weights (tile layout, see umma.cuh): rows = out, cols = in
constexpr uint32_t SW_W1D = 0;      // [64][32]
constexpr uint32_t SW_W2D = 4096;   // [16][64]
constexpr uint32_t SW_W1R = 6144;   // [64][32]
constexpr uint32_t SW_W2R = 10240;  // [64][64]
constexpr uint32_t SW_W3R = 18432;  // [16][64]
constexpr uint32_t SW_END = 20480;
// activation tiles, 128 rows each
constexpr uint32_t S_X   = SW_END;            // [128][32]  hash-grid features
constexpr uint32_t S_H1  = S_X + 8192;        // [128][64]  relu(density hidden)      (inference: reused for G1, G2)
constexpr uint32_t S_RIN = S_H1 + 16384;      // [128][32]  rgb-net input = [density out 16 | SH 16]
constexpr uint32_t S_INFER_END = S_RIN + 8192;
// training-only tiles. Buffers whose lifetimes do not overlap share storage, which brings a CTA to 104 KB so that TWO fit on an SM and
// their serial MMA -> epilogue chains overlap:  dG1 is written after G2's last use (the ReLU mask of dG2), dH1 after dG2's last use (the
// GEMMs of the step that produces dG1), dOd after dOr's last use (the first backward step).
constexpr uint32_t S_G1  = S_INFER_END;       // [128][64]
constexpr uint32_t S_G2  = S_G1 + 16384;      // [128][64]
constexpr uint32_t S_DG1 = S_G2;              //            aliases G2
constexpr uint32_t S_DOR = S_G2 + 16384;      // [128][16]  dL/d(rgb-net out), cols 3..15 zero
constexpr uint32_t S_DOD = S_DOR;             //            dL/d(density-net out), aliases dOr
constexpr uint32_t S_DG2 = S_DOR + 4096;      // [128][64]
constexpr uint32_t S_DH1 = S_DG2;             //            aliases dG2
constexpr uint32_t S_TRAIN_END = S_DG2 + 16384;
constexpr uint32_t S_CTRL = 64;               // mbarrier + tmem address, placed after the tiles

// ---- TMEM map (columns) ---------------------------------------------------------------------------
constexpr uint32_t TM_ACC = 0;      // 64 columns: layer output / data gradient (M = 128)
constexpr uint32_t TM_DW1D = 64;    // [64 x 32]   dW1d[o][i]          (M = 64)
constexpr uint32_t TM_DW2D = 96;    // [64 x 16]   dW2d^T[i][o]
constexpr uint32_t TM_DW1R = 112;   // [64 x 32]   dW1r[o][i]
constexpr uint32_t TM_DW2R = 144;   // [64 x 64]   dW2r[o][i]
constexpr uint32_t TM_DW3R = 208;   // [64 x 16]   dW3r^T[i][o]
constexpr uint32_t TM_COLS_TRAIN = 256, TM_COLS_INFER = 64;

// PLAIN / PLAIN_TRAIN: the 32 -> 64 -> 64 -> 16 network alone (neural image / SDF models), inference and forward + backward + weight gradients
enum MlpMode { MODE_DENSITY = 0, MODE_INFERENCE = 1, MODE_TRAIN = 2, MODE_PLAIN = 3, MODE_PLAIN_TRAIN = 4 };
constexpr uint32_t PLAIN_PARAMS = 64 * 32 + 64 * 64 + 16 * 64, PLAIN_W1 = 0, PLAIN_W2 = 2048, PLAIN_W3 = 6144;

// ---- helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ void load_matrix_to_tile(uint8_t* smem, uint32_t dst, const __half* __restrict__ src, uint32_t rows, uint32_t cols) {
	const uint32_t chunks_per_row = cols >> 3, n_chunks = rows * chunks_per_row;
	for (uint32_t q = threadIdx.x; q < n_chunks; q += blockDim.x) {
		const uint32_t r = q / chunks_per_row, c = q % chunks_per_row;
		*reinterpret_cast<uint4*>(smem + dst + tile_offset(r, c, cols)) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * cols + c * 8));
	}
}

__device__ __forceinline__ void load_weights(uint8_t* smem, const __half* __restrict__ mlp) {
	load_matrix_to_tile(smem, SW_W1D, mlp + MLP_W1D, 64, 32);
	load_matrix_to_tile(smem, SW_W2D, mlp + MLP_W2D, 16, 64);
	load_matrix_to_tile(smem, SW_W1R, mlp + MLP_W1R, 64, 32);
	load_matrix_to_tile(smem, SW_W2R, mlp + MLP_W2R, 64, 64);
	load_matrix_to_tile(smem, SW_W3R, mlp + MLP_W3R, 16, 64);
}

// D[128 x N] = A[128 x K] * W[N x K]^T : both operands K-major.
__device__ __forceinline__ void issue_forward(uint32_t d_tmem, uint32_t a_saddr, uint32_t a_cols, uint32_t w_saddr, uint32_t K, uint32_t N) {
	const uint32_t idesc = make_idesc_f16(128, N, false, false);
	for (uint32_t k = 0; k < K / 16; ++k) {
		mma_f16_ss(d_tmem, desc_kmajor(a_saddr, a_cols, 0, 2 * k), desc_kmajor(w_saddr, K, 0, 2 * k), idesc, k > 0);
	}
}
// dIn[128 x n_in] = dOut[128 x n_out] * W[n_out x n_in] : A K-major (K = n_out), B = the forward weight tile read MN-major.
__device__ __forceinline__ void issue_dgrad(uint32_t d_tmem, uint32_t g_saddr, uint32_t g_cols, uint32_t w_saddr, uint32_t n_in, uint32_t n_out) {
	const uint32_t idesc = make_idesc_f16(128, n_in, false, true);
	for (uint32_t k = 0; k < n_out / 16; ++k) {
		mma_f16_ss(d_tmem, desc_kmajor(g_saddr, g_cols, 0, 2 * k), desc_mnmajor(w_saddr, n_in, 0, 16 * k), idesc, k > 0);
	}
}
// D[64 x N] += P[128 x 64]^T * Q[128 x N] : both tiles read MN-major, K = the 128 sample rows. M = 64.
__device__ __forceinline__ void issue_wgrad(uint32_t d_tmem, uint32_t p_saddr, uint32_t q_saddr, uint32_t q_cols, uint32_t N, bool first) {
	const uint32_t idesc = make_idesc_f16(64, N, true, true);
	for (uint32_t k = 0; k < TILE / 16; ++k) {
		mma_f16_ss(d_tmem, desc_mnmajor(p_saddr, 64, 0, 16 * k), desc_mnmajor(q_saddr, q_cols, 0, 16 * k), idesc, !(first && k == 0));
	}
}

// Degree-4 spherical harmonics of dir*2-1 (tcnn spherical_harmonics.h:62-101), 16 coefficients.
__device__ __forceinline__ void sh4(float dx, float dy, float dz, float* out) {
	const float x = dx * 2.f - 1.f, y = dy * 2.f - 1.f, z = dz * 2.f - 1.f;
	const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
	out[0] = 0.28209479177387814f;
	out[1] = -0.48860251190291987f * y;
	out[2] = 0.48860251190291987f * z;
	out[3] = -0.48860251190291987f * x;
	out[4] = 1.0925484305920792f * xy;
	out[5] = -1.0925484305920792f * yz;
	out[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
	out[7] = -1.0925484305920792f * xz;
	out[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
	out[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
	out[10] = 2.8906114426405538f * xy * z;
	out[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
	out[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
	out[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
	out[14] = 1.4453057213202769f * z * (x2 - y2);
	out[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}


} // namespace ngpb
