/*
 * k_peak_i8.cu -- roofline denominator: what the tcgen05 tensor pipe of THIS GPU delivers for
 * kind::i8 under THIS box's power conditions, measured with a loads-free loop of the very MMA
 * shape the product kernel issues (cta_group::2, M = 256, N = 256, K = 32, int32 accumulate in
 * TMEM).  Operands are two static shared-memory tiles of +1 / -1 / 0 bytes (the value mix of
 * the real operand panel, so the switching power is comparable); nothing is loaded from HBM.
 * SURVEY.md section 8(d) asks for this next to the 2 x bf16 figure of MEASURED_PEAKS.json.
 */
#include "ccg_internal.h"

namespace {

constexpr int THREADS = 128;
constexpr int TILE_BYTES = 128 * 128;
constexpr uint32_t IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (256 >> 3) << 17) | ((uint32_t) (256 >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
	return (uint64_t) ((saddr & 0x3FFFFu) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) |
	       ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_peak_i8(int iters, unsigned *sink) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t raw = smem_u32(smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	uint8_t *tiles = smem_raw + (base - raw);
	const uint32_t bar_done = base + 2 * TILE_BYTES;
	const uint32_t bar_mid = bar_done + 8;
	const uint32_t tmem_slot = bar_mid + 8;
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));
	uint32_t cta_rank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
	const int warp = threadIdx.x >> 5;

	/* +1 / -1 / 0 bytes, 1 % zeros like the masked positions of the operand panel */
	uint32_t x = 0x9E3779B9u * (blockIdx.x * THREADS + threadIdx.x + 1);
	for(int i = threadIdx.x; i < 2 * TILE_BYTES; i += THREADS) {
		x = x * 1664525u + 1013904223u;
		const unsigned r = x >> 24;
		tiles[i] = r < 3 ? 0 : ((r & 1) ? 0x01 : 0xFF);
	}
	if(threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_done) : "memory");
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mid) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_slot_ptr;

	if(threadIdx.x == 32 && cta_rank == 0) {
		const uint64_t adesc = make_desc(base), bdesc = make_desc(base + TILE_BYTES);
		for(int it = 0; it < iters; ++it) {
			const uint32_t d = tmem + ((it & 3) == 3 ? 256u : 0u);          /* 3 : 1 split over the two accumulators */
#pragma unroll
			for(int k = 0; k < 4; ++k) {
				asm volatile(
				    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				    "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
				    ::"r"(d), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(IDESC2), "r"((uint32_t) (it > 3)) : "memory");
			}
			/* the product kernel commits once per stage as well */
			asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
			             ::"r"(bar_mid), "h"((uint16_t) 1) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
		             ::"r"(bar_done), "h"((uint16_t) 1) : "memory");
		uint32_t ok = 0;
		while(!ok) {
			asm volatile(
			    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			    : "=r"(ok) : "r"(bar_done) : "memory");
		}
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		if(sink && iters < 0) *sink = tmem;
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	if(warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

} // namespace

/* Runs the loop for about target_ms on every CTA pair of the device and returns the rate in
 * int8 TOP/s (2 ops per MAC); < 0 on failure. */
extern "C" double ccg_measure_i8_peak(ccg_ctx *ctx, double target_ms) {
	if(!ctx) return -1.0;
	constexpr int smem = 2 * TILE_BYTES + 64 + 1024;
	if(cudaSetDevice(ctx->device) != cudaSuccess) return -1.0;
	if(cudaFuncSetAttribute(k_peak_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1.0;
	const int pairs = ccg_umma_pair_slots(ctx);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double tops = -1.0;
	int iters = 2000;
	for(int pass = 0; pass < 3; ++pass) {
		cudaEventRecord(e0, ctx->stream);
		k_peak_i8<<<2 * pairs, THREADS, smem, ctx->stream>>>(iters, 0);
		cudaEventRecord(e1, ctx->stream);
		if(cudaEventSynchronize(e1) != cudaSuccess) { tops = -1.0; break; }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		const double ops = (double) pairs * iters * 4.0 * 256.0 * 256.0 * 32.0 * 2.0;
		tops = ops / (ms * 1e-3) / 1e12;
		ctx->launches++;
		if(pass == 2) break;
		/* size the next pass for the requested duration */
		double scale = target_ms / (ms > 0.01f ? ms : 0.01f);
		double next = iters * scale;
		if(next > 2.0e9) next = 2.0e9;
		if(next < 1000) next = 1000;
		iters = (int) next;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	return tops;
}
