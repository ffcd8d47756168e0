// tcgen05 / TMEM / mbarrier primitives for sm_100a, written as inline PTX.
// Bit layouts follow the PTX ISA "tcgen05" chapter (shared-memory matrix descriptor, instruction
// descriptor for .kind::f16, TMEM addressing lane<<16 | column).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace ngpb {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Bounded wait for the warp-specialised pipelines: a protocol error must end in a trap (a sticky launch failure the host reports), never in a hung GPU.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
	const uint32_t addr = smem_u32(bar);
	uint32_t done = 0, spins = 0;
	long long t0 = 0;
	while (true) {
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
			"selp.u32 %0, 1, 0, p;\n"
			"}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
		if (done) return;
		if ((++spins & 1023u) == 0u) {
			const long long now = clock64();
			if (t0 == 0) t0 = now;
			else if (now - t0 > 4000000000ll) asm volatile("trap;"); // ~2 s at 2 GHz
		}
	}
}
// arrive (count 1) with release semantics at CTA scope
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" :: "r"(smem_u32(bar)) : "memory");
}
// arrive (count 1) and announce `bytes` of asynchronous (bulk copy) traffic that will complete on this barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, `bytes` a multiple of 16, both addresses 16-byte aligned); completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- shared-memory counters as lightweight signals (a try_wait on an already completed mbarrier costs ~170 cycles on B200, an ld.acquire ~30) ----
// Producers add 1 with release semantics after their own writes (and proxy fences); the consumer polls until the counter reaches its target.
__device__ __forceinline__ void flag_signal(uint32_t* flag) {
	asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" :: "r"(smem_u32(flag)) : "memory");
}
__device__ __forceinline__ uint32_t flag_peek(const uint32_t* flag) {
	uint32_t v;
	asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(flag)) : "memory");
	return v;
}
// Polls with relaxed loads (an acquire load per iteration carries a fence) and acquires once at the end.
__device__ __forceinline__ void flag_wait_bounded(const uint32_t* flag, uint32_t target) {
	const uint32_t addr = smem_u32(flag);
	uint32_t spins = 0;
	long long t0 = 0;
	while (true) {
		uint32_t v;
		asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
		if ((int32_t)(v - target) >= 0) break;
		if ((++spins & 4095u) == 0u) {
			const long long now = clock64();
			if (t0 == 0) t0 = now;
			else if (now - t0 > 4000000000ll) asm volatile("trap;");
		}
	}
	asm volatile("fence.acquire.cta;" ::: "memory");
}

// ---- proxies and tcgen05 fences -------------------------------------------------------------------
// Generic-proxy shared-memory writes must be fenced before the async proxy (tcgen05.mma) reads them.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) ----------------------------------------------
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
	static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns: power of two in [32,512]");
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "n"(NCOLS) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(NCOLS) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// Shared-memory matrix descriptor, no swizzle (layout_type 0), sm_100 version field = 1.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// Instruction descriptor for tcgen05.mma.kind::f16 with fp16 inputs and fp32 accumulation.
//   bits [4,6) D format (1 = f32), [7,10) A format (0 = f16), [10,13) B format (0 = f16),
//   bit 15 A major (0 = K, 1 = MN), bit 16 B major, bits [17,23) N >> 3, bits [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
	return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
		"}\n" :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

// Same with the descriptors given as (low word, compile-time high word): the low word is `address >> 4` plus a constant, so the issuing thread
// spends one add per operand instead of rebuilding the 64-bit descriptor (measured: ~100 cycles of the issuing thread per tcgen05.mma with make_smem_desc).
// Executed by ALL lanes of a converged warp with warp-uniform operands; one elected lane issues. Keeping the issuing code warp-uniform lets the
// compiler hold the operands in uniform registers (a divergent `if (lane == 0)` body costs an ELECT / R2UR.BROADCAST loop per instruction).
template <uint32_t A_HI, uint32_t B_HI, uint32_t IDESC, uint32_t ACCUMULATE>
__device__ __forceinline__ void mma_f16_ss_fast(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo) {
	asm volatile(
		"{\n"
		".reg .pred p, pe;\n"
		".reg .b64 da, db;\n"
		"mov.b64 da, {%1, %2};\n"
		"mov.b64 db, {%3, %4};\n"
		"setp.ne.b32 p, %6, 0;\n"
		"elect.sync _|pe, 0xffffffff;\n"
		"@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
		"}\n" :: "r"(d_tmem), "r"(a_lo), "n"(A_HI), "r"(b_lo), "n"(B_HI), "n"(IDESC), "n"(ACCUMULATE) : "memory");
}
// tcgen05.commit from a converged warp (one elected lane)
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
	asm volatile(
		"{\n"
		".reg .pred pe;\n"
		"elect.sync _|pe, 0xffffffff;\n"
		"@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
		"}\n" :: "r"(smem_u32(bar)) : "memory");
}
// low / high words of a no-swizzle descriptor (see make_smem_desc): low = address >> 4 | (LBO >> 4) << 16, high = SBO >> 4 | version 1 << 14
__host__ __device__ constexpr uint32_t desc_lo_const(uint32_t lbo_bytes) { return ((lbo_bytes >> 4) & 0x3FFFu) << 16; }
__host__ __device__ constexpr uint32_t desc_hi_const(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

// Makes the mbarrier track completion of all tcgen05 ops issued so far by this thread
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: shape 32x32b, each thread reads its own lane (row), N consecutive columns ----
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t* r) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
		: "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
	asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		: "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
		  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
		  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
		  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
		: "r"(taddr) : "memory");
}

// ---- operand tiles in shared memory ------------------------------------------------------------
// A tile holds R rows x C fp16 columns as 8x8 "core matrices" of 128 contiguous bytes (8 rows x 16 B),
// core matrices of one 8-row group adjacent along the column direction. The same bytes are a valid
//   * K-major operand  (rows = M or N, columns = K):  LBO = 128 B, SBO = (C/8)*128 B, +256 B per K=16 step
//   * MN-major operand (columns = M or N, rows = K):  SBO = 128 B, LBO = (C/8)*128 B, +2*(C/8)*128 B per K=16 step
// which is what lets one activation tile feed the forward GEMM (K = features) and the weight-gradient
// GEMM (K = samples) without a transpose.
__host__ __device__ constexpr uint32_t tile_bytes(uint32_t rows, uint32_t cols) { return rows * cols * 2; }
__device__ __forceinline__ uint32_t tile_offset(uint32_t row, uint32_t col_chunk, uint32_t cols) {
	return (row >> 3) * (cols >> 3) * 128 + col_chunk * 128 + (row & 7) * 16;
}

// K-major view of rows [row0, row0+rows) with K starting at column chunk k_chunk0.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, uint32_t cols, uint32_t row0, uint32_t k_chunk0) {
	return make_smem_desc(tile_saddr + (row0 >> 3) * (cols >> 3) * 128 + k_chunk0 * 128, /*lbo*/128, /*sbo*/(cols >> 3) * 128);
}
// MN-major view: MN starts at column chunk mn_chunk0, K (rows) starts at row k_row0 (multiple of 8).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, uint32_t cols, uint32_t mn_chunk0, uint32_t k_row0) {
	return make_smem_desc(tile_saddr + (k_row0 >> 3) * (cols >> 3) * 128 + mn_chunk0 * 128, /*lbo*/(cols >> 3) * 128, /*sbo*/128);
}

// {lo = relu(a), hi = relu(b)} as fp16x2 in one conversion (cvt's first source goes to the upper half)
__device__ __forceinline__ uint32_t pack_half2_relu(float a, float b) {
	uint32_t r;
	asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
	return r;
}
__device__ __forceinline__ uint32_t pack_half2_rn(float a, float b) {
	uint32_t r;
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
	return r;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
	__half2 h = __floats2half2_rn(a, b);
	return *reinterpret_cast<uint32_t*>(&h);
}

} // namespace umma
} // namespace ngpb
