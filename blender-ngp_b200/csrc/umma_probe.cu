// Timing probe of the tcgen05 building blocks used by the MLP kernels (development tool behind ngpb_probe_umma; tools/umma_probe.py prints the table).
// One CTA of 128 threads; every figure is SM clock cycles (clock64) seen by the issuing / reading thread.
#include "common.cuh"
#include "umma.cuh"
#include "nerf_mlp_shared.cuh"
#include "../../include/ngpb.h"

namespace ngpb {
using namespace umma;

constexpr uint32_t PROBE_BARS = 40;

// out[0..]: see the table in ngpb_probe_umma
__global__ void __launch_bounds__(128) umma_probe_kernel(long long* __restrict__ out, const uint32_t sections)
{
	extern __shared__ __align__(128) uint8_t smem[];
	// A tile [128][64], W tile [64][64], both K-major; contents irrelevant (zeros)
	constexpr uint32_t OFF_A = 0, OFF_W = 16384, OFF_H = 24576, OFF_BAR = 40960;
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
	uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + PROBE_BARS * 8);
	const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (uint32_t i = tid; i < OFF_BAR / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
	if (warp == 0) tmem_alloc<512>(tmem_slot);
	if (tid == 0) { for (uint32_t i = 0; i < PROBE_BARS; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
	fence_proxy_async_smem();
	tc_fence_before_sync();
	__syncthreads();
	tc_fence_after_sync();
	const uint32_t tmem_base = *tmem_slot, sbase = smem_u32(smem);
	uint32_t bar_i = 0; (void)bar_i;
	long long slot = 0;
	auto record = [&](long long v) { if (tid == 0) out[slot] = v; ++slot; };

	// ---- A: latency of one batch: k MMAs (M=128, N, K=16 each) + commit -> barrier complete, measured by the issuing thread ----
	if (sections & 1u) for (uint32_t ni = 0; ni < 2; ++ni) { const uint32_t N = ni == 0 ? 64u : 16u;
		for (uint32_t k = 1; k <= 8; k *= 2) {
			long long best = 1ll << 60;
			for (int rep = 0; rep < 4; ++rep) {
				__syncthreads();
				if (tid == 0) { bar_i = 0; mbar_init(&bars[0], 1); fence_mbar_init(); }
				__syncthreads();
				if (tid == 0) {
					const uint32_t idesc = make_idesc_f16(128, N, false, false);
					const long long t0 = clock64();
					for (uint32_t j = 0; j < k; ++j) mma_f16_ss(tmem_base, desc_kmajor(sbase + OFF_A, 64, 0, 2 * (j & 3)), desc_kmajor(sbase + OFF_W, 64, 0, 2 * (j & 3)), idesc, j > 0);
					mma_commit(&bars[bar_i]);
					mbar_wait_bounded(&bars[bar_i], 0);
					const long long t1 = clock64();
					best = min(best, t1 - t0);
				}
			}
			record(best);
		}
	}
	// ---- B: throughput of back-to-back batches into 4 different accumulators, no waiting in between: cycles per batch ----
	slot = 8;
	if (sections & 2u) for (uint32_t ni = 0; ni < 2; ++ni) { const uint32_t N = ni == 0 ? 64u : 16u;
		for (uint32_t k = 1; k <= 4; k *= 2) {
			__syncthreads();
			if (tid == 0) { for (uint32_t i = 0; i < 16; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
			__syncthreads();
			long long dt = 0, dt_issue = 0;
			if (tid == 0) {
				const uint32_t idesc = make_idesc_f16(128, N, false, false);
				const long long t0 = clock64();
				for (uint32_t b = 0; b < 16; ++b) { // 16 batches, each with its own single-use barrier, round-robin over 4 accumulators
					for (uint32_t j = 0; j < k; ++j) mma_f16_ss(tmem_base + (b & 3) * 64, desc_kmajor(sbase + OFF_A, 64, 0, 2 * (j & 3)), desc_kmajor(sbase + OFF_W, 64, 0, 2 * (j & 3)), idesc, j > 0);
					mma_commit(&bars[b]);
				}
				const long long t1 = clock64();
				for (uint32_t b = 0; b < 16; ++b) mbar_wait_bounded(&bars[b], 0);
				const long long t2 = clock64();
				dt = t2 - t0; dt_issue = t1 - t0;
			}
			record(dt / 16);
			record(dt_issue / 16);
		}
	}
	// ---- C: tcgen05.ld: latency of x32 + wait (one warp), and all four warps together ----
	slot = 20;
	if (sections & 4u) {
		__syncthreads();
		uint32_t r[32];
		const uint32_t t_row = tmem_base + ((warp * 32u) << 16);
		long long t0 = clock64();
		if (warp == 0) { tmem_ld_x32(t_row, r); tmem_ld_wait(); }
		long long t1 = clock64();
		uint32_t acc = 0;
		for (int i = 0; i < 32; ++i) acc ^= r[i];
		if (acc == 0x12345678u) out[200] = 1;
		record(t1 - t0);
		__syncthreads();
		t0 = clock64();
		tmem_ld_x32(t_row, r); tmem_ld_wait();
		tmem_ld_x32(t_row + 32, r); tmem_ld_wait();
		t1 = clock64();
		for (int i = 0; i < 32; ++i) acc ^= r[i];
		if (acc == 0x12345678u) out[200] = 1;
		record(t1 - t0); // 2 x (x32 + wait), 4 warps concurrently
		__syncthreads();
		t0 = clock64();
		for (int q = 0; q < 8; ++q) { tmem_ld_x32(t_row + (q & 1) * 32, r); tmem_ld_wait(); for (int i = 0; i < 32; ++i) acc ^= r[i]; }
		t1 = clock64();
		if (acc == 0x12345678u) out[200] = 1;
		record((t1 - t0) / 8); // per x32 load, sustained, 4 warps
	}
	// ---- D: fences and the epilogue's stores ----
	slot = 23;
	if (sections & 8u) {
		__syncthreads();
		long long t0 = clock64();
		fence_proxy_async_smem();
		long long t1 = clock64();
		record(t1 - t0);
		t0 = clock64();
		tc_fence_before_sync();
		t1 = clock64();
		record(t1 - t0);
		t0 = clock64();
		tc_fence_after_sync();
		t1 = clock64();
		record(t1 - t0);
		// 8 x STS.128 into the tile layout + proxy fence (what one epilogue thread does per 64-column layer)
		t0 = clock64();
		#pragma unroll
		for (uint32_t c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(smem + OFF_H + tile_offset(tid, c, 64)) = make_uint4(c, tid, 0, 0);
		fence_proxy_async_smem();
		t1 = clock64();
		record(t1 - t0);
		// mbarrier: arrive + try_wait on a barrier that completes by this arrive
		__syncthreads();
		if (tid == 0) { mbar_init(&bars[0], 1); fence_mbar_init(); }
		__syncthreads();
		if (tid == 0) {
			t0 = clock64();
			mbar_arrive(&bars[0]);
			mbar_wait_bounded(&bars[0], 0);
			t1 = clock64();
		}
		record(t1 - t0);
		// try_wait on an already completed barrier
		if (tid == 0) { t0 = clock64(); mbar_wait_bounded(&bars[0], 0); t1 = clock64(); }
		record(t1 - t0);
	}
	// ---- E: ping-pong round trip: thread 0 issues a batch (2 MMAs N=64) + commit; warps wait, tcgen05.ld x32 x2, fence, arrive; thread 0 waits: cycles per round ----
	slot = 29;
	if (sections & 16u) {
		__syncthreads();
		uint64_t* done = &bars[0]; uint64_t* ready = &bars[1];
		if (tid == 0) { mbar_init(done, 1); mbar_init(ready, 4); fence_mbar_init(); }
		__syncthreads();
		const uint32_t t_row = tmem_base + ((warp * 32u) << 16);
		const uint32_t idesc = make_idesc_f16(128, 64, false, false);
		long long t0 = clock64();
		uint32_t acc = 0;
		for (uint32_t round = 0; round < 8; ++round) {
			if (tid == 0) {
				if (round > 0) mbar_wait_bounded(ready, (round - 1) & 1);
				tc_fence_after_sync();
				mma_f16_ss(tmem_base, desc_kmajor(sbase + OFF_A, 64, 0, 0), desc_kmajor(sbase + OFF_W, 64, 0, 0), idesc, false);
				mma_f16_ss(tmem_base, desc_kmajor(sbase + OFF_A, 64, 0, 2), desc_kmajor(sbase + OFF_W, 64, 0, 2), idesc, true);
				mma_commit(done);
			}
			__syncwarp();
			mbar_wait_bounded(done, round & 1);
			tc_fence_after_sync();
			uint32_t r[32];
			tmem_ld_x32(t_row, r); tmem_ld_wait();
			for (int i = 0; i < 32; ++i) acc ^= r[i];
			tmem_ld_x32(t_row + 32, r); tmem_ld_wait();
			for (int i = 0; i < 32; ++i) acc ^= r[i];
			*reinterpret_cast<uint4*>(smem + OFF_H + tile_offset(tid, 0, 64)) = make_uint4(acc, 0, 0, 0);
			tc_fence_before_sync();
			fence_proxy_async_smem();
			__syncwarp();
			if (lane == 0) mbar_arrive(ready);
		}
		long long t1 = clock64();
		if (acc == 0x12345678u) out[200] = 1;
		record((t1 - t0) / 8);
	}
	// ---- F: tensor-pipe throughput with the warp-uniform issue path (mma_f16_ss_fast): W warps each issue 32 MMAs into their own accumulator, one commit each ----
	slot = 30;
	if (sections & 32u) {
		for (uint32_t cfg = 0; cfg < 6; ++cfg) {
			// cfg: 0: 1 warp N=64 K-major; 1: 4 warps N=64; 2: 1 warp N=16; 3: 4 warps N=16; 4: 1 warp wgrad-shaped (M=64, N=64, MN-major both); 5: 4 warps wgrad-shaped
			const uint32_t n_warps = (cfg & 1u) ? 4u : 1u;
			__syncthreads();
			if (tid == 0) { for (uint32_t i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
			__syncthreads();
			const long long t0 = clock64();
			if (warp < n_warps) {
				const uint32_t acc = tmem_base + warp * 64u, a16 = (sbase + OFF_A) >> 4, w16 = (sbase + OFF_W) >> 4;
				constexpr uint32_t LO = desc_lo_const(128), HI64 = desc_hi_const(8 * 128), MN_LO = desc_lo_const(8 * 128), MN_HI = desc_hi_const(128);
				#pragma unroll
				for (uint32_t j = 0; j < 32; ++j) {
					if (cfg < 2) mma_f16_ss_fast<HI64, HI64, make_idesc_f16(128, 64, false, false), 1>(acc, a16 + LO + 16 * (j & 3), w16 + LO + 16 * (j & 3));
					else if (cfg < 4) mma_f16_ss_fast<HI64, HI64, make_idesc_f16(128, 16, false, false), 1>(acc, a16 + LO + 16 * (j & 3), w16 + LO + 16 * (j & 3));
					else mma_f16_ss_fast<MN_HI, MN_HI, make_idesc_f16(64, 64, true, true), 1>(acc, a16 + MN_LO + 128 * (j & 7), a16 + MN_LO + 128 * (j & 7));
				}
				mma_commit_elect(&bars[warp]);
				__syncwarp();
				mbar_wait_bounded(&bars[warp], 0);
			}
			__syncthreads();
			const long long t1 = clock64();
			record((t1 - t0) / 32);
		}
	}
	tc_fence_before_sync();
	__syncthreads();
	if (warp == 0) tmem_dealloc<512>(tmem_base);
}

} // namespace ngpb

using namespace ngpb;

// Development probe: fills cycles[0..n) (n <= 64) with the measurements listed in tools/umma_probe.py. Synchronises the stream.
extern "C" int ngpb_probe_umma(void* stream, long long* cycles_host, uint32_t n, uint32_t sections) {
	long long* d = nullptr;
	try {
		if (!cycles_host || n == 0 || n > 64) { set_last_error("ngpb_probe_umma: invalid argument"); return NGPB_ERR_INVALID_ARGUMENT; }
		NGPB_CUDA_CHECK(cudaMalloc(&d, 256 * sizeof(long long)));
		NGPB_CUDA_CHECK(cudaMemset(d, 0, 256 * sizeof(long long)));
		const int smem = 40960 + PROBE_BARS * 8 + 16;
		NGPB_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		for (int rep = 0; rep < 2; ++rep) { // second run: warm instruction cache
			umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d, sections);
			NGPB_LAUNCH_CHECK();
		}
		NGPB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
		NGPB_CUDA_CHECK(cudaMemcpy(cycles_host, d, n * sizeof(long long), cudaMemcpyDeviceToHost));
		cudaFree(d);
		return 0;
	} catch (const std::exception& e) { if (d) cudaFree(d); set_last_error(e.what()); return NGPB_ERR_RUNTIME; }
}
