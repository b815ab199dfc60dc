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


/* ---- kind::mxf4 (e2m1 operands, UE8M0 block scales = 1.0, f32 accumulate): pipe rate and exactness probe ----
 * Would FP4 operands serve this integer contraction?  The +1 / -1 / 0 channel values are exact in
 * e2m1 and kind::mxf4 runs at twice the kind::i8 rate, but the accumulator is f32: the sums stay
 * exact only if the pipe adds exact partial dot products into a true fp32 accumulator (|sum| < 2^24).
 * This kernel runs the MMA on operand tiles whose every 16-byte chunk is the same 32-nibble pattern
 * (so the 128-byte swizzle is invisible), which makes every D element iters * 8 * dot32(PA, PB), and
 * compares all of them with that integer. */
constexpr uint32_t IDESC_MXF4 = (1u << 7) | (1u << 10) | ((uint32_t) (256 >> 3) << 17) | (1u << 23) | ((uint32_t) (256 >> 4) << 24);

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_peak_fp4(int iters, int check, unsigned long long *mismatch, int *dot_out) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t raw = smem_u32(smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	uint8_t *tiles = smem_raw + (base - raw);
	const uint32_t bar_done = base + 2 * TILE_BYTES;
	const uint32_t bar_mid = bar_done + 8;
	const uint32_t tmem_slot = bar_mid + 8;
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));
	__shared__ int s_dot;
	uint32_t cta_rank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
	const int warp = threadIdx.x >> 5;

	/* the two 32-nibble patterns: +1 = 0x2, -1 = 0xA, 0 = 0x0 (e2m1); PB = PA with every 4th sign flipped */
	if(threadIdx.x == 0) {
		uint32_t x = 0x2545F491u;
		int dot = 0;
		uint8_t pa[16], pb[16];
		for(int b = 0; b < 16; ++b) {
			uint8_t va = 0, vb = 0;
			for(int h = 0; h < 2; ++h) {
				x = x * 1664525u + 1013904223u;
				const unsigned r = x >> 24;
				const int a = r < 8 ? 0 : ((r & 1) ? 1 : -1);
				const int bb = ((2 * b + h) & 3) == 0 ? -a : a;
				va |= (uint8_t) ((a == 0 ? 0x0 : (a > 0 ? 0x2 : 0xA)) << (4 * h));
				vb |= (uint8_t) ((bb == 0 ? 0x0 : (bb > 0 ? 0x2 : 0xA)) << (4 * h));
				dot += a * bb;
			}
			pa[b] = va;
			pb[b] = vb;
		}
		s_dot = dot;
		if(dot_out) *dot_out = dot;
		for(int i = 0; i < TILE_BYTES; ++i) { tiles[i] = pa[i & 15]; tiles[TILE_BYTES + i] = pb[i & 15]; }
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_done) : "memory");
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mid) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	if(warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_slot_ptr;
	/* scale factors: UE8M0 1.0 = 0x7F in every byte of columns [256, 288) of every lane */
	{
		const uint32_t one = 0x7F7F7F7Fu;
		const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16) + 256;
		for(int c = 0; c < 32; c += 8)
			asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(one) : "memory");
		asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	if(threadIdx.x == 32 && cta_rank == 0) {
		const uint64_t adesc = make_desc(base), bdesc = make_desc(base + TILE_BYTES);
		const uint32_t sfa = tmem + 256, sfb = tmem + 264;
		for(int it = 0; it < iters; ++it) {
#pragma unroll
			for(int k = 0; k < 4; ++k) {
				asm volatile(
				    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				    "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
				    ::"r"(tmem), "l"(adesc + 2 * k), "l"(bdesc + 2 * k), "r"(IDESC_MXF4), "r"((uint32_t) (it > 0 || k > 0)), "r"(sfa), "r"(sfb)
				    : "memory");
			}
			asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
			             ::"r"(bar_mid), "h"((uint16_t) 1) : "memory");
		}
		asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
		             ::"r"(bar_done), "h"((uint16_t) 3) : "memory");
	}
	/* both CTAs: wait for the accumulators, then compare every element with the exact integer */
	{
		uint32_t ok = 0;
		while(!ok) {
			asm volatile(
			    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			    : "=r"(ok) : "r"(bar_done) : "memory");
		}
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		if(check) {
			const float expect = (float) ((long long) iters * 8LL * s_dot);
			unsigned long long bad = 0;
			for(int cb = 0; cb < 256; cb += 32) {
				uint32_t r[32];
				asm volatile(
				    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
				    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
				    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
				      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
				      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
				      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
				    : "r"(tmem + ((uint32_t) (warp * 32) << 16) + cb) : "memory");
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				for(int e = 0; e < 32; ++e) bad += __uint_as_float(r[e]) != expect;
				if(cb == 0 && blockIdx.x == 0 && threadIdx.x == 0 && dot_out) dot_out[1] = (int) __uint_as_float(r[0]);
			}
			if(bad) atomicAdd(mismatch, bad);
		}
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

/* kind::mxf4 probe: rate in e2m1 TOP/s (2 ops per MAC) of the loads-free loop, and -- in *inexact -- how many
 * accumulator elements differed from the exact integer after accumulating up to about `check_sum` (< 2^24). */
extern "C" double ccg_measure_fp4_peak(ccg_ctx *ctx, double target_ms, double check_sum, long long *inexact, int *info) {
	if(!ctx) return -1.0;
	constexpr int smem = 2 * TILE_BYTES + 64 + 1024;
	if(cudaSetDevice(ctx->device) != cudaSuccess) return -1.0;
	if(cudaFuncSetAttribute(k_peak_fp4, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1.0;
	const int pairs = ccg_umma_pair_slots(ctx);
	unsigned long long *d_bad = 0;
	int *d_info = 0;
	if(cudaMalloc(&d_bad, 8) != cudaSuccess || cudaMalloc(&d_info, 8) != cudaSuccess) return -1.0;
	cudaMemsetAsync(d_bad, 0, 8, ctx->stream);
	cudaMemsetAsync(d_info, 0, 8, ctx->stream);
	/* exactness: a short run to learn dot32, then one sized for the requested sum */
	int h_info[2] = {0, 0};
	k_peak_fp4<<<2 * pairs, THREADS, smem, ctx->stream>>>(16, 1, d_bad, d_info);
	cudaMemcpyAsync(h_info, d_info, 8, cudaMemcpyDeviceToHost, ctx->stream);
	if(cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaFree(d_bad); cudaFree(d_info); cudaGetLastError(); return -1.0; }
	const int dot = h_info[0] ? h_info[0] : 1;
	long long it_check = (long long) (check_sum / (8.0 * (dot < 0 ? -dot : dot)));
	if(it_check < 1) it_check = 1;
	if(it_check > 2000000000LL) it_check = 2000000000LL;
	k_peak_fp4<<<2 * pairs, THREADS, smem, ctx->stream>>>((int) it_check, 1, d_bad, d_info);
	unsigned long long bad = 0;
	cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream);
	cudaMemcpyAsync(h_info, d_info, 8, cudaMemcpyDeviceToHost, ctx->stream);
	cudaStreamSynchronize(ctx->stream);
	if(inexact) *inexact = (long long) bad;
	if(info) { info[0] = dot; info[1] = h_info[1]; info[2] = (int) it_check; }
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double tops = -1.0;
	int iters = 2000;
	for(int pass = 0; pass < 3; ++pass) {
		cudaEventRecord(e0, ctx->stream);
		k_peak_fp4<<<2 * pairs, THREADS, smem, ctx->stream>>>(iters, 0, d_bad, 0);
		cudaEventRecord(e1, ctx->stream);
		if(cudaEventSynchronize(e1) != cudaSuccess) { tops = -1.0; break; }
		float ms = 0.f;
		cudaEventElapsedTime(&ms, e0, e1);
		tops = (double) pairs * iters * 4.0 * 256.0 * 256.0 * 64.0 * 2.0 / (ms * 1e-3) / 1e12;
		ctx->launches++;
		if(pass == 2) break;
		double next = iters * (target_ms / (ms > 0.01f ? ms : 0.01f));
		if(next > 2.0e9) next = 2.0e9;
		if(next < 1000) next = 1000;
		iters = (int) next;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d_bad);
	cudaFree(d_info);
	return tops;
}
