/*
 * k_pairdist_popc.cu -- K2b: bit-sliced LOP3+POPC all-vs-all compare with the
 * fused epilogue (K3).
 *
 * Replaces the reference's per-pair hot loop
 *     maskProxi (proxi == 0)  fsacmp.c:355-389   inc = inc_i & inc_j
 *     fsacmpair               fsacmp.c:587-633   n += popcount(inc), dist += differing lanes
 *     fsacmp                  fsacmp.c:552-585   (shared-mask mode)
 * and the work distribution + epilogue of cmpairFsaThrd / cmpFsaThrd
 * (fsacmpthrd.c:261-480 / :108-259).
 *
 * Work item = (64x64 sample tile of the lower triangle, K slice).  A CTA of
 * 256 threads (16x16, each thread a 4x4 block of pairs) streams the two
 * sample panels of its tile through a 3-stage TMA + mbarrier pipeline:
 * one 4-D TMA box [KC chunks][planes][64 slots][4 words] per panel and stage.
 * Per 32-base word pair the INT pipe does
 *     t = h_i ^ h_j;  u = (l_i ^ l_j) | t;  m = m_i & m_j;  d = u & m;
 *     mism += popc(d);  ninc += popc(m)
 * (4 LOP3 + 2 POPC + 2 IADD; shared-mask mode: 2 LOP3 + 1 POPC + 1 IADD).
 * K slices of one tile are combined with integer RED.ADD (exact, order
 * independent); the slice that arrives last (ticket counter) runs the epilogue.
 *
 * Roofline: INT-pipe bound (smem traffic is 0.375 LDS.128 per pair-word, HBM
 * traffic is one read of the planes per co-resident wave thanks to L2).
 */
#include "ccg_internal.h"
#include "epilogue.cuh"

namespace {

constexpr int KC = 4;        /* chunks (of 128 bases) per pipeline stage */
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr int T = CCG_TILE;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, unsigned parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return ok != 0;
}

/* bounded wait: a lost TMA completion traps instead of hanging the GPU */
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
	if(mbar_try_wait(bar, parity)) return;
	const long long t0 = clock64();
	unsigned spins = 0;
	while(!mbar_try_wait(bar, parity)) {
		/* watchdog: ~2 s at 2 GHz, far beyond any legitimate wait in these kernels */
		if((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();
	}
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4, %5, %6}], [%2];"
	    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
	    : "memory");
}

template <int NPL>
__global__ void __launch_bounds__(THREADS, 2)
k_pairdist_popc(const __grid_constant__ CUtensorMap tmap, const PopcParams p) {
	constexpr int PANEL_VEC = KC * NPL * T;            /* uint4 per panel */
	constexpr int PANEL_BYTES = PANEL_VEC * 16;
	constexpr int STAGE_BYTES = 2 * PANEL_BYTES;
	constexpr bool PAIR = NPL == 3;

	extern __shared__ uint8_t smem_raw[];
	const uint32_t raw = smem_u32(smem_raw);
	const uint32_t base = (raw + 127u) & ~127u;
	const uint8_t *sm = smem_raw + (base - raw);
	const uint32_t bars = base + STAGES * STAGE_BYTES;

	const int tid = threadIdx.x;
	const int tx = tid & 15, ty = tid >> 4;

	/* work item -> (local tile, K slice); K slices outermost so co-resident
	 * CTAs walk the same chunk range and share it through L2 */
	const int lt = blockIdx.x % p.ntiles;
	const int ks = blockIdx.x / p.ntiles;
	const int ti = p.tiles[lt].x, tj = p.tiles[lt].y;

	const int c_begin = ks * p.chunks_per_split;
	int span = p.chunks - c_begin;
	if(span > p.chunks_per_split) span = p.chunks_per_split;
	const int niter = span > 0 ? (span + KC - 1) / KC : 0;

	if(tid == 0) {
#pragma unroll
		for(int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * s, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();

	auto issue = [&](int it) {
		const int s = it % STAGES;
		const uint32_t dst = base + s * STAGE_BYTES;
		const uint32_t bar = bars + 8 * s;
		mbar_expect_tx(bar, STAGE_BYTES);
		tma_load_4d(dst, &tmap, bar, 0, ti * T, 0, c_begin + it * KC);
		tma_load_4d(dst + PANEL_BYTES, &tmap, bar, 0, tj * T, 0, c_begin + it * KC);
	};
	if(tid == 0) {
		for(int it = 0; it < STAGES && it < niter; ++it) issue(it);
	}

	unsigned accD[4][4], accN[4][4];
#pragma unroll
	for(int a = 0; a < 4; ++a)
#pragma unroll
		for(int b = 0; b < 4; ++b) accD[a][b] = accN[a][b] = 0;

	for(int it = 0; it < niter; ++it) {
		const int s = it % STAGES;
		mbar_wait(bars + 8 * s, (it / STAGES) & 1);
		const uint4 *pI = reinterpret_cast<const uint4 *>(sm + s * STAGE_BYTES);
		const uint4 *pJ = pI + PANEL_VEC;
#pragma unroll 1
		for(int c = 0; c < KC; ++c) {
			uint4 jh[4], jl[4], jm[4];
#pragma unroll
			for(int b = 0; b < 4; ++b) {
				jh[b] = pJ[(c * NPL + 0) * T + tx + 16 * b];
				jl[b] = pJ[(c * NPL + 1) * T + tx + 16 * b];
				if(PAIR) jm[b] = pJ[(c * NPL + 2) * T + tx + 16 * b];
			}
#pragma unroll
			for(int a = 0; a < 4; ++a) {
				const uint4 ih = pI[(c * NPL + 0) * T + ty + 16 * a];
				const uint4 il = pI[(c * NPL + 1) * T + ty + 16 * a];
				uint4 im = make_uint4(0, 0, 0, 0);
				if(PAIR) im = pI[(c * NPL + 2) * T + ty + 16 * a];
#pragma unroll
				for(int b = 0; b < 4; ++b) {
#define CCG_WORD(f)                                                          \
	{                                                                        \
		unsigned u = (il.f ^ jl[b].f) | (ih.f ^ jh[b].f);                    \
		if(PAIR) {                                                           \
			unsigned m = im.f & jm[b].f;                                     \
			accN[a][b] += __popc(m);                                         \
			u &= m;                                                          \
		}                                                                    \
		accD[a][b] += __popc(u);                                             \
	}
					CCG_WORD(x) CCG_WORD(y) CCG_WORD(z) CCG_WORD(w)
#undef CCG_WORD
				}
			}
		}
		__syncthreads();                       /* everyone is done reading stage s */
		if(tid == 0 && it + STAGES < niter) issue(it + STAGES);
	}

	/* ---- combine K slices, fused epilogue by the last arriver ---- */
	uint32_t *accT = p.acc + (size_t) lt * 2 * T * T;
	if(p.ksplit == 1) {
#pragma unroll
		for(int a = 0; a < 4; ++a)
#pragma unroll
			for(int b = 0; b < 4; ++b) {
				const int e = (ty + 16 * a) * T + tx + 16 * b;
				accT[e] = accD[a][b];
				if(PAIR) accT[T * T + e] = accN[a][b];
			}
	} else {
		__shared__ int s_last;
#pragma unroll
		for(int a = 0; a < 4; ++a)
#pragma unroll
			for(int b = 0; b < 4; ++b) {
				const int e = (ty + 16 * a) * T + tx + 16 * b;
				atomicAdd(accT + e, accD[a][b]);
				if(PAIR) atomicAdd(accT + T * T + e, accN[a][b]);
			}
		__threadfence();
		__syncthreads();
		if(tid == 0) {
			unsigned ticket = atomicAdd(p.tickets + lt, 1u);
			s_last = ticket == (unsigned) p.ksplit - 1;
		}
		__syncthreads();
		if(!s_last) return;
		__threadfence();
#pragma unroll
		for(int a = 0; a < 4; ++a)
#pragma unroll
			for(int b = 0; b < 4; ++b) {
				const int e = (ty + 16 * a) * T + tx + 16 * b;
				accD[a][b] = __ldcg(accT + e);
				if(PAIR) accN[a][b] = __ldcg(accT + T * T + e);
			}
	}
#pragma unroll
	for(int a = 0; a < 4; ++a)
#pragma unroll
		for(int b = 0; b < 4; ++b) {
			const int i = ti * T + ty + 16 * a;
			const int j = tj * T + tx + 16 * b;
			if(i > j) ccg_write_cell(p.ep, i, j, accD[a][b], accN[a][b]);
		}
}

template <int NPL>
cudaError_t launch(ccg_ctx *ctx, const PopcParams &p) {
	constexpr int smem = STAGES * 2 * KC * NPL * T * 16 + STAGES * 8 + 128;
	cudaError_t e = cudaFuncSetAttribute(k_pairdist_popc<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	if(e != cudaSuccess) return e;
	const long long items = (long long) p.ntiles * p.ksplit;
	if(items <= 0) return cudaSuccess;
	k_pairdist_popc<NPL><<<(unsigned) items, THREADS, smem, ctx->stream>>>(ctx->tmap, p);
	ctx->launches++;
	return cudaGetLastError();
}

} // namespace

int ccg_popc_kc(void) { return KC; }

cudaError_t ccg_launch_popc(ccg_ctx *ctx, const PopcParams &p) {
	return ctx->nplanes == 3 ? launch<3>(ctx, p) : launch<2>(ctx, p);
}
