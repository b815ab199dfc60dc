/*
 * k_pairdist_fused.cu -- K2a, fused form: the int8 tcgen05 contraction of
 * k_pairdist_umma.cu with the operand expansion done inside the CTA, so the
 * only HBM-resident sample data is the 3-bit/base plane store (no int8 panel,
 * no K slabs, nothing replicated per rank but the planes themselves).
 *
 * Replaces the same reference code: maskProxi (proxi == 0) fsacmp.c:355-389,
 * fsacmpair fsacmp.c:587-633, fsacmp fsacmp.c:552-585 and the pair loop of
 * cmpairFsaThrd / cmpFsaThrd (fsacmpthrd.c:261-480 / :108-259).
 *
 * Work item = (128 x 256 macro tile, K slice of chunks).  Per chunk (128 bases):
 *   warp 0   TMA: plane boxes [planes][128 slots][4 words] for the A rows and the two
 *            B half-panels (18 KiB) into a 2-deep ring
 *   warps 2-9  expanders: bit planes -> four K-major SWIZZLE_128B operand stages
 *            (mask channel, then the three tetrahedral code channels), 48 KiB each,
 *            3-deep ring; generic-proxy stores are published to the tensor core with
 *            fence.proxy.async + mbarrier
 *   warp 1   issues tcgen05.mma kind::i8 M128 N256 K32 (4 per stage) into the two TMEM
 *            accumulators (I: mask channel, S: code channels)
 *   warps 2-5  afterwards drain TMEM and RED.ADD the int32 partials into C_S / C_I
 * Diagonal macro tiles (A rows inside the B panel) skip the A expansion and point the
 * A descriptor into the B stage.
 *
 * Roofline: tensor pipe, paced by the expanders' ALU work (about 1.2 integer ops per
 * operand byte); algorithmic work 8 int8 ops per pairwise base comparison.
 */
#include "ccg_internal.h"

namespace {

constexpr int BM = 128;              /* this kernel works on 128 x 256 halves of the macro tiles */
constexpr int BN = CCG_UMMA_BN;
constexpr int BK = 128;
constexpr int OP_STAGES = 3;
constexpr int PL_STAGES = 2;
constexpr int A_BYTES = BM * BK;
constexpr int B_BYTES = BN * BK;
constexpr int OP_BYTES = A_BYTES + B_BYTES;                 /* 49152 */
constexpr int PL_ROWS = BM + BN;                            /* 384 slots per chunk */
constexpr int PL_BYTES = 3 * PL_ROWS * 16;                  /* 18432: [A|B0|B1][plane][128][4 words] */
constexpr int EXP_WARPS = 8;
constexpr int EXP_THREADS = EXP_WARPS * 32;
constexpr int THREADS = 64 + EXP_THREADS;
constexpr int ITEMS = PL_ROWS * 8 / EXP_THREADS;            /* (row, 16-base segment) items per thread: 12 */
constexpr int TMEM_COLS = 512;
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) (BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, unsigned parity) {
	uint32_t ok;
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
	    "selp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity) {
	if(mbar_try_wait(bar, parity)) return;
	const long long t0 = clock64();
	unsigned spins = 0;
	while(!mbar_try_wait(bar, parity)) {
		if((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) __trap();   /* ~2 s watchdog */
	}
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
	asm volatile(
	    "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
	    " [%0], [%1, {%3, %4, %5, %6}], [%2];"
	    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
	return (uint64_t) ((saddr & 0x3FFFFu) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) |
	       ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
	    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
	      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
	      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
	      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	    : "r"(taddr) : "memory");
	asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
/* 4 plane bits (bit j <-> base j of the group) -> 4 bytes of 0/1 */
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

template <int NPL>
__global__ void __launch_bounds__(THREADS, 1)
k_pairdist_fused(const __grid_constant__ CUtensorMap tmap_pl, const UmmaParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t raw = smem_u32(smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	const uint32_t op_base = base;                                         /* OP_STAGES x 48 KiB */
	const uint32_t pl_base = base + OP_STAGES * OP_BYTES;                  /* PL_STAGES x 18 KiB */
	const uint32_t bar_pl_full = pl_base + PL_STAGES * PL_BYTES;
	const uint32_t bar_pl_empty = bar_pl_full + 8 * PL_STAGES;
	const uint32_t bar_op_full = bar_pl_empty + 8 * PL_STAGES;
	const uint32_t bar_op_empty = bar_op_full + 8 * OP_STAGES;
	const uint32_t bar_accum = bar_op_empty + 8 * OP_STAGES;
	const uint32_t tmem_slot = bar_accum + 8;
	const uint8_t *sm = smem_raw + (base - raw);
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int tile = blockIdx.x % p.ntiles;
	const int ks = blockIdx.x / p.ntiles;
	const int tm = p.tiles[tile].x, tn = p.tiles[tile].y;
	const int c_begin = ks * p.chunks_per_slice;
	int nchunk = p.slab_chunks - c_begin;
	if(nchunk > p.chunks_per_slice) nchunk = p.chunks_per_slice;
	if(nchunk < 0) nchunk = 0;
	/* diagonal macro tile: the A rows are one half of the B panel */
	const bool a_in_b = (tm >> 1) == tn;
	const int a_half = tm & 1;

	if(threadIdx.x == 0) {
		for(int s = 0; s < PL_STAGES; ++s) {
			mbar_init(bar_pl_full + 8 * s, 1);
			mbar_init(bar_pl_empty + 8 * s, EXP_WARPS);
		}
		for(int s = 0; s < OP_STAGES; ++s) {
			mbar_init(bar_op_full + 8 * s, EXP_WARPS);
			mbar_init(bar_op_empty + 8 * s, 1);
		}
		mbar_init(bar_accum, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	if(warp == 2) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_slot_ptr;

	if(warp == 0) {
		/* ===== plane TMA producer ===== */
		if(lane == 0) {
			for(int c = 0; c < nchunk; ++c) {
				const int s = c % PL_STAGES;
				if(c >= PL_STAGES) mbar_wait(bar_pl_empty + 8 * s, ((c / PL_STAGES) - 1) & 1);
				const uint32_t dst = pl_base + s * PL_BYTES;
				const uint32_t bar = bar_pl_full + 8 * s;
				const int box = NPL * 128 * 16;
				mbar_expect_tx(bar, (a_in_b ? 2 : 3) * box);
				if(!a_in_b) tma_load_4d(dst, &tmap_pl, bar, 0, tm * BM, 0, c_begin + c);
				tma_load_4d(dst + 3 * 128 * 16, &tmap_pl, bar, 0, tn * BN, 0, c_begin + c);
				tma_load_4d(dst + 2 * 3 * 128 * 16, &tmap_pl, bar, 0, tn * BN + 128, 0, c_begin + c);
			}
		}
	} else if(warp == 1) {
		/* ===== MMA issuer ===== */
		if(lane == 0) {
			uint32_t usedS = 0, usedI = 0;
			const int nkb = nchunk * 4;
			for(int kb = 0; kb < nkb; ++kb) {
				const int s = kb % OP_STAGES;
				mbar_wait(bar_op_full + 8 * s, (kb / OP_STAGES) & 1);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t st = op_base + s * OP_BYTES;
				const uint64_t bdesc = make_desc(st + A_BYTES);
				const uint64_t adesc = a_in_b ? make_desc(st + A_BYTES + a_half * A_BYTES) : make_desc(st);
				const bool is_mask = (kb & 3) == 0;             /* stage order within a chunk: mask, c0, c1, c2 */
				const uint32_t d = tmem + (is_mask ? BN : 0);
#pragma unroll
				for(int k = 0; k < BK / 32; ++k) {
					umma_i8(d, adesc + 2 * k, bdesc + 2 * k, is_mask ? usedI : usedS);
					if(is_mask) usedI = 1; else usedS = 1;
				}
				umma_commit(bar_op_empty + 8 * s);
			}
			umma_commit(bar_accum);
		}
	} else {
		/* ===== expanders (then epilogue) ===== */
		const int et = threadIdx.x - 64;                 /* 0 .. EXP_THREADS-1 */
		const int seg = et & 7;                          /* 16-base segment of the chunk */
		const int q = seg >> 1, half = seg & 1;
		const int r0 = et >> 3;                          /* rows r0 + 32*i */
		int kb = 0;
		for(int c = 0; c < nchunk; ++c) {
			const int ps = c % PL_STAGES;
			mbar_wait(bar_pl_full + 8 * ps, (c / PL_STAGES) & 1);
			const uint32_t *pl = reinterpret_cast<const uint32_t *>(sm + OP_STAGES * OP_BYTES + ps * PL_BYTES);
			uint32_t hh[ITEMS], ll[ITEMS], mm[ITEMS];
#pragma unroll
			for(int i = 0; i < ITEMS; ++i) {
				const int row = r0 + 32 * i;                     /* 0..127 A, 128..383 B */
				const int part = row >> 7, slot = row & 127;     /* part 0 = A, 1 = B0, 2 = B1 */
				const uint32_t *b = pl + part * (3 * 128 * 4);
				if(part == 0 && a_in_b) { hh[i] = ll[i] = mm[i] = 0; continue; }
				uint32_t h = b[(0 * 128 + slot) * 4 + q];
				uint32_t l = b[(1 * 128 + slot) * 4 + q];
				uint32_t m = NPL == 3 ? b[(2 * 128 + slot) * 4 + q] : 0xFFFFFFFFu;
				/* base k <-> bit 31-k; reverse so base k <-> bit k, keep this thread's 16 bases */
				hh[i] = (__brev(h) >> (16 * half)) & 0xFFFFu;
				ll[i] = (__brev(l) >> (16 * half)) & 0xFFFFu;
				mm[i] = (__brev(m) >> (16 * half)) & 0xFFFFu;
			}
			/* the plane buffer is free once every expander has its words in registers */
			__syncwarp();
			if(lane == 0) mbar_arrive(bar_pl_empty + 8 * ps);

#pragma unroll 1
			for(int ch = 0; ch < 4; ++ch, ++kb) {
				const int s = kb % OP_STAGES;
				if(kb >= OP_STAGES) mbar_wait(bar_op_empty + 8 * s, ((kb / OP_STAGES) - 1) & 1);
				const uint32_t st = op_base + s * OP_BYTES;
#pragma unroll
				for(int i = 0; i < ITEMS; ++i) {
					const int row = r0 + 32 * i;
					if(row < BM && a_in_b) continue;
					/* sign plane of this stage: mask stage has none; c0 = h, c1 = l, c2 = h ^ l */
					const uint32_t neg = ch == 0 ? 0u : ch == 1 ? hh[i] : ch == 2 ? ll[i] : (hh[i] ^ ll[i]);
					uint32_t o[4];
#pragma unroll
					for(int g = 0; g < 4; ++g) {
						const uint32_t ones = spread4((mm[i] >> (4 * g)) & 0xFu);
						const uint32_t minus = spread4((neg >> (4 * g)) & 0xFu);
						o[g] = minus * 0xFEu + ones;             /* +1 = 0x01, -1 = 0xFF, unknown = 0 */
					}
					/* K-major SWIZZLE_128B: 16-byte chunk index XOR (row % 8); the B tile follows the A tile */
					const uint32_t addr = st + row * BK + ((seg ^ (row & 7)) << 4);
					st_shared_v4(addr, o[0], o[1], o[2], o[3]);
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* publish to the tensor core's proxy */
				__syncwarp();
				if(lane == 0) mbar_arrive(bar_op_full + 8 * s);
			}
		}
		/* ===== epilogue: TMEM -> registers -> RED.ADD into C (warps 2..5, one TMEM lane quarter each) ===== */
		if(warp < 6 && nchunk > 0) {
			const int quarter = warp & 3;
			const int row = tm * BM + quarter * 32 + lane;
			mbar_wait(bar_accum, 0);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			int *cS = p.C_S + (size_t) row * p.ldc + tn * BN;
			int *cI = p.C_I + (size_t) row * p.ldc + tn * BN;
			const int jlim = row - tn * BN;
#pragma unroll 1
			for(int cb = 0; cb < BN / 32; ++cb) {
				if(__all_sync(0xffffffffu, cb * 32 >= jlim)) break;
				uint32_t r[32];
				tmem_ld32(tmem + ((uint32_t) (quarter * 32) << 16) + cb * 32, r);
#pragma unroll
				for(int e = 0; e < 32; ++e)
					if(cb * 32 + e < jlim && r[e]) atomicAdd(cS + cb * 32 + e, (int) r[e]);
				tmem_ld32(tmem + ((uint32_t) (quarter * 32) << 16) + BN + cb * 32, r);
#pragma unroll
				for(int e = 0; e < 32; ++e)
					if(cb * 32 + e < jlim && r[e]) atomicAdd(cI + cb * 32 + e, (int) r[e]);
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if(warp == 2) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
	}
}

} // namespace

cudaError_t ccg_launch_fused(ccg_ctx *ctx, const UmmaParams &p) {
	constexpr int smem = OP_STAGES * OP_BYTES + PL_STAGES * PL_BYTES + 8 * (2 * PL_STAGES + 2 * OP_STAGES + 1) + 16 + 1024;
	const long long items = (long long) p.ntiles * p.kslices;
	if(items <= 0) return cudaSuccess;
	cudaError_t e;
	if(ctx->nplanes == 3) {
		e = cudaFuncSetAttribute(k_pairdist_fused<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		if(e != cudaSuccess) return e;
		k_pairdist_fused<3><<<(unsigned) items, THREADS, smem, ctx->stream>>>(ctx->tmap_pl, p);
	} else {
		e = cudaFuncSetAttribute(k_pairdist_fused<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		if(e != cudaSuccess) return e;
		k_pairdist_fused<2><<<(unsigned) items, THREADS, smem, ctx->stream>>>(ctx->tmap_pl, p);
	}
	ctx->launches++;
	return cudaGetLastError();
}
