/*
 * k_pairdist_umma.cu -- K2a: the all-vs-all compare as a dense contraction on the tcgen05 tensor cores, plus the
 * operand expansion and the finalising epilogue.
 *
 * Replaces the same reference code as K2b (k_pairdist_popc.cu): maskProxi (proxi == 0) fsacmp.c:355-389, fsacmpair
 * fsacmp.c:587-633, fsacmp fsacmp.c:552-585 and the pair loop + epilogue of cmpairFsaThrd / cmpFsaThrd
 * (fsacmpthrd.c:261-480 / :108-259).
 *
 * Algebra.  Per base every sample gets three "tetrahedral" channels and one mask
 * channel:
 *     A=(+1,+1,+1) C=(+1,-1,-1) G=(-1,+1,-1) T=(-1,-1,+1)  unknown=(0,0,0), m = known ? 1 : 0
 * Over the three code channels  S(i,j) = sum t_i.t_j = 3*match - mismatch  on jointly known positions, over the mask
 * channel  I(i,j) = sum m_i*m_j = the inclusion count.  Hence   mismatch = (3*I - S) / 4   exactly, in int32
 * (|S|, I <= 3L < 2^31).  K = 4L elements: 8 tensor ops per pairwise base comparison.
 *
 * What is in this file:
 *   k_expand_fp4 / k_expand_fp4_rows   operand panel X in e2m1 (two values per byte), from the bit planes or straight
 *                         from the reference's packed words (rows lent by the caller / streamed host rows):
 *                         X[slot/128][chunkpair*4+channel][slot%128][128 B], one 128-byte channel row = 256 bases,
 *                         tile-blocked so that a 128-row x 128-byte TMA box is 16 KiB contiguous
 *   k_expand              the same panel in int8 (one value per byte, 128 bases per channel row), CCG_I8=1
 *   k_pairdist_umma2<FP4> THE kernel: persistent CTA pairs (cluster of 2, tcgen05 cta_group::2), work item = (256 x 256
 *                         macro tile of the lower triangle, K slice); per CTA a 6-stage ring of 32 KiB (its 128 rows of
 *                         A + its half of B, TMA SWIZZLE_128B, complete_tx on the leader's mbarrier); warp 0 = TMA
 *                         producer, warp 1 of the leader = MMA issuer (kind::mxf4.block_scale M256 N256 K64 with all
 *                         block scales 1.0 and f32 accumulators, or kind::i8 K32 with two s32 accumulators), warps 2-5
 *                         of both CTAs = epilogue: tcgen05.ld, f32 -> s32, RED.ADD into the dense C_S / C_I (integer
 *                         split-K is exact).  The CTAs of a round run in lock-step (epoch counters in global memory) so
 *                         that the operand rows neighbouring tiles share are served by L2; "thin" items run the last,
 *                         mostly empty macro-tile row transposed with a small MMA N (see the kernel).
 *   k_pairdist_umma       the first, single-CTA int8 kernel (128 x 256 tiles, 4 stages), kept for CCG_UMMA1=1 experiments
 *   k_finalize_umma       mismatch = (3I - S) / 4 and the reference epilogue (epilogue.cuh) over the packed cells
 */
#include <string.h>

#include "ccg_internal.h"
#include "epilogue.cuh"

namespace {

constexpr int BMT = CCG_UMMA_BM;     /* 256 rows of a macro tile = one CTA pair (2 x 128 TMEM lanes) */
constexpr int BM = 128;              /* rows per CTA (A tile, TMEM lanes) */
constexpr int BN = CCG_UMMA_BN;      /* 256 cols  (B tile, TMEM columns per accumulator) */
constexpr int BK = 128;              /* K bytes per stage = one channel row of a chunk */
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK;
constexpr int B_BYTES = BN * BK;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
/* instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format S32 (2) @4,
 * a/b_format signed int8 (1) @7/@10, K-major A and B, n_dim = N>>3 @17, m_dim = M>>4 @24 */
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) (BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

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
	    : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok != 0;
}
/* bounded wait: a lost completion traps (reported as a CUDA error) instead of hanging the GPU.  limit = cycles
 * (UmmaParams::watchdog: ~2 s by default, far beyond any legitimate wait in these kernels; 0 = wait for ever, which
 * is what a profiler's kernel replay needs -- CCG_WATCHDOG_S=0) */
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity, long long limit) {
	if(mbar_try_wait(bar, parity)) return;
	const long long t0 = clock64();
	unsigned spins = 0;
	while(!mbar_try_wait(bar, parity)) {
		if((++spins & 1023u) == 0 && limit > 0 && clock64() - t0 > limit) __trap();
	}
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
/* K-major SWIZZLE_128B shared-memory matrix descriptor (SmemDescriptor in
 * cute/arch/mma_sm100_desc.hpp): start>>4 @0, LBO=1 @16 (ignored for swizzled K-major),
 * SBO = 1024 B (8 rows x 128 B) >> 4 @32, version 1 @46, layout SWIZZLE_128B (2) @61 */
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

/* ------------------------------------------------------------------ */
/* operand expansion: bit planes -> int8 panel                         */
/* ------------------------------------------------------------------ */
/* 4 plane bits (bit j <-> base j of the group) -> 4 bytes of 0/1 */
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

/* One thread per (slot, chunk, segment of 16 bases): it produces the 16 bytes of all four
 * channels, so the known-mask spread and the plane loads are shared.  A warp covers 4
 * consecutive slots x 8 segments and every store instruction writes 4 x 128 contiguous bytes.
 * The code planes are already ANDed with the mask (k_encode), so h, l and h^l are zero
 * wherever the base is unknown. */
__global__ void __launch_bounds__(256)
k_expand(const uint32_t *__restrict__ planes, int n_pad, int nplanes, int slot0, int slots, int chunk0, int nchunks,
         int8_t *__restrict__ X, size_t nkb) {
	/* bounded persistent grid (grid-stride): leaves thread slots on every SM for the GEMM CTAs of the
	 * previous slab, which run concurrently on the main stream */
	const long long total = (long long) slots * nchunks * 8;
	for(long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x; gid < total;
	    gid += (long long) gridDim.x * blockDim.x) {
		const long long item = gid >> 3;
		const int seg = (int) (gid & 7);
		const int slot = slot0 + (int) (item % slots);
		const int cl = (int) (item / slots);           /* chunk within the slab */
		const int q = seg >> 1, half = seg & 1;
		const size_t prow = (size_t) (chunk0 + cl) * nplanes;
		uint32_t h = planes[((prow + 0) * n_pad + slot) * 4 + q];
		uint32_t l = planes[((prow + 1) * n_pad + slot) * 4 + q];
		uint32_t m = nplanes == 3 ? planes[((prow + 2) * n_pad + slot) * 4 + q] : 0xFFFFFFFFu;
		/* base k of the word <-> bit 31-k; reverse so base k <-> bit k, then take this thread's 16 bases */
		h = __brev(h) >> (16 * half);
		l = __brev(l) >> (16 * half);
		m = __brev(m) >> (16 * half);
		uint32_t c0[4], c1[4], c2[4], c3[4];
#pragma unroll
		for(int g = 0; g < 4; ++g) {
			const uint32_t ones = spread4((m >> (4 * g)) & 0xFu);     /* 0x01 where known */
			const uint32_t sh = spread4((h >> (4 * g)) & 0xFu);       /* 0x01 where channel 0 is -1 */
			const uint32_t sl = spread4((l >> (4 * g)) & 0xFu);       /* 0x01 where channel 1 is -1 */
			const uint32_t sx = sh ^ sl;                              /* 0x01 where channel 2 is -1 */
			c0[g] = sh * 0xFEu + ones;                                /* +1 = 0x01, -1 = 0xFF, unknown = 0 */
			c1[g] = sl * 0xFEu + ones;
			c2[g] = sx * 0xFEu + ones;
			c3[g] = ones;
		}
		/* tile-blocked panel: X[slot/128][k-block][slot%128][128 B], k-block = chunk*4 + channel, so the
		 * 128 rows x 128 B box a TMA load fetches is 16 KiB contiguous */
		const size_t tile_row = ((size_t) (slot >> 7) * nkb + (size_t) cl * 4) * 128 + (slot & 127);
		uint4 *dst = reinterpret_cast<uint4 *>(X + tile_row * 128 + seg * 16);
		/* channel rows are 128 x 128 B = 1024 uint4 apart; streaming stores: the panel is far larger
		 * than L2 and must not evict the working set of a GEMM running beside this kernel */
		__stcs(dst + 0 * 1024, make_uint4(c0[0], c0[1], c0[2], c0[3]));
		__stcs(dst + 1 * 1024, make_uint4(c1[0], c1[1], c1[2], c1[3]));
		__stcs(dst + 2 * 1024, make_uint4(c2[0], c2[1], c2[2], c2[3]));
		__stcs(dst + 3 * 1024, make_uint4(c3[0], c3[1], c3[2], c3[3]));
	}
}

/* e2m1 variant of the operand panel (kind::mxf4): the same four channels as 4-bit floats, +1 = 0x2,
 * -1 = 0xA, 0 = 0x0 -- all exact in e2m1 -- two per byte, base k of a row in nibble k.  One 128-byte
 * channel row now holds a PAIR of chunks (256 bases), so the panel is
 *     X4[slot/128][chunk pair * 4 + channel][slot%128][128 B]
 * at 2 bytes per base and sample.  One thread per (slot, chunk, 32-base word): 16 bytes per channel. */
__device__ __forceinline__ uint32_t spread8(uint32_t x) {      /* bit k of the byte -> bit 4k */
	x = (x | (x << 12)) & 0x000F000Fu;
	x = (x | (x << 6)) & 0x03030303u;
	x = (x | (x << 3)) & 0x11111111u;
	return x;
}
/* keep the even bits of x (bit 2k -> bit k): the low / high code bit of every base of a packed reference word */
__device__ __forceinline__ uint32_t even_bits(uint64_t x) {
	x &= 0x5555555555555555ull;
	x = (x | (x >> 1)) & 0x3333333333333333ull;
	x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
	x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
	x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
	x = (x | (x >> 16));
	return (uint32_t) x;
}

/* where the expansion reads a sample from: the plane store, or -- for the slots [first, first + count) of rows the
 * caller lent us (ccg_put_samples_packed_dev_borrowed) -- straight from the reference's packed words, which saves
 * the 12 + 12 bytes per word of the plane store round trip (k_repack_packed) on the tensor path */
struct PackedSource {
	const uint64_t *seqs;          /* NULL: everything comes from the planes */
	const uint32_t *masks;         /* NULL: shared-mask mode, gmask below */
	const uint32_t *gmask;
	long wstride;
	int first, count, words;
};

__global__ void __launch_bounds__(256)
k_expand_fp4(const uint32_t *__restrict__ planes, int n_pad, int nplanes, int chunks_total, int slot0, int slots, int chunk0,
             int npairs, int8_t *__restrict__ X, size_t nkb, const PackedSource src) {
	const long long total = (long long) slots * npairs * 8;           /* 2 chunks x 4 words per pair */
	for(long long gid = (long long) blockIdx.x * blockDim.x + threadIdx.x; gid < total;
	    gid += (long long) gridDim.x * blockDim.x) {
		const long long item = gid >> 3;
		const int sub = (int) (gid & 7);                                /* chunk of the pair (bit 2), word of the chunk */
		const int slot = slot0 + (int) (item % slots);
		const int kp = (int) (item / slots);
		const int chunk = chunk0 + 2 * kp + (sub >> 2), q = sub & 3;
		uint32_t h = 0, l = 0, m = nplanes == 3 ? 0u : 0xFFFFFFFFu;
		if(src.seqs && slot >= src.first && slot < src.first + src.count) {
			const long w = (long) chunk * CCG_CHUNK_WORDS + q;
			if(w < src.words) {
				const size_t at = (size_t) (slot - src.first) * src.wstride + w;
				const uint64_t x = __ldg(src.seqs + at);
				const uint32_t mk = src.masks ? __ldg(src.masks + at) : __ldg(src.gmask + w);
				h = even_bits(x >> 1) & mk;
				l = even_bits(x) & mk;
				if(nplanes == 3) m = mk;
			}
		} else if(chunk < chunks_total) {
			const size_t prow = (size_t) chunk * nplanes;
			h = planes[((prow + 0) * n_pad + slot) * 4 + q];
			l = planes[((prow + 1) * n_pad + slot) * 4 + q];
			if(nplanes == 3) m = planes[((prow + 2) * n_pad + slot) * 4 + q];
		}
		h = __brev(h); l = __brev(l); m = __brev(m);                    /* base k of the word <-> bit k */
		uint32_t c0[4], c1[4], c2[4], c3[4];
#pragma unroll
		for(int g = 0; g < 4; ++g) {
			const uint32_t one = spread8((m >> (8 * g)) & 0xFFu) << 1;    /* 0x2 where known */
			const uint32_t sh = spread8((h >> (8 * g)) & 0xFFu) << 3;     /* sign bit where channel 0 is -1 */
			const uint32_t sl = spread8((l >> (8 * g)) & 0xFFu) << 3;
			c0[g] = one | sh;
			c1[g] = one | sl;
			c2[g] = one | (sh ^ sl);
			c3[g] = one;
		}
		const size_t tile_row = ((size_t) (slot >> 7) * nkb + (size_t) kp * 4) * 128 + (slot & 127);
		uint4 *dst = reinterpret_cast<uint4 *>(X + tile_row * 128 + sub * 16);
		__stcs(dst + 0 * 1024, make_uint4(c0[0], c0[1], c0[2], c0[3]));
		__stcs(dst + 1 * 1024, make_uint4(c1[0], c1[1], c1[2], c1[3]));
		__stcs(dst + 2 * 1024, make_uint4(c2[0], c2[1], c2[2], c2[3]));
		__stcs(dst + 3 * 1024, make_uint4(c3[0], c3[1], c3[2], c3[3]));
	}
}

/* The same expansion for host rows that are being STREAMED in K slabs (feed_slab in ccg_api.cu): the batch of rows a
 * copy stream has just staged on the device -- the words of one slab, reference packed format -- goes straight into
 * the slab's panel; the bit-plane store is not built at all on this path.  One thread per (row, word of a chunk
 * pair) walks the slab's chunk pairs: 8 neighbouring threads read 64 contiguous bytes of a row and write 128
 * contiguous bytes of every channel row, and a thread's popcounts add up to one atomic for the per-sample counts
 * (getNpos, fsacmp.c:487).  gmask (shared-mask mode, masks == NULL) points at the slab's first word. */
__global__ void __launch_bounds__(256)
k_expand_fp4_rows(const uint64_t *__restrict__ seqs, const uint32_t *__restrict__ masks, const uint32_t *__restrict__ gmask, long wstride,
                  int nvalid, int first, int count, int npairs, int pairs_per_block, int nplanes, int8_t *__restrict__ X, size_t nkb,
                  unsigned *__restrict__ inc) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	const int s = t >> 3, sub = t & 7;
	if(s >= count) return;
	const int slot = first + s;
	const int kp0 = blockIdx.y * pairs_per_block;
	int kp1 = kp0 + pairs_per_block;
	if(kp1 > npairs) kp1 = npairs;
	const uint64_t *srow = seqs + (size_t) s * wstride;
	const uint32_t *mrow = masks ? masks + (size_t) s * wstride : gmask;
	unsigned known = 0;
	for(int kp = kp0; kp < kp1; ++kp) {
		const int w = kp * 8 + sub;
		uint64_t x = 0;
		uint32_t mk = 0;
		if(w < nvalid) { x = __ldg(srow + w); mk = __ldg(mrow + w); }
		uint32_t h = __brev(even_bits(x >> 1) & mk), l = __brev(even_bits(x) & mk);
		const uint32_t m = __brev(nplanes == 3 ? mk : 0xFFFFFFFFu);
		known += (unsigned) __popc(mk);
		uint32_t c0[4], c1[4], c2[4], c3[4];
#pragma unroll
		for(int g = 0; g < 4; ++g) {
			const uint32_t one = spread8((m >> (8 * g)) & 0xFFu) << 1;
			const uint32_t sh = spread8((h >> (8 * g)) & 0xFFu) << 3;
			const uint32_t sl = spread8((l >> (8 * g)) & 0xFFu) << 3;
			c0[g] = one | sh;
			c1[g] = one | sl;
			c2[g] = one | (sh ^ sl);
			c3[g] = one;
		}
		const size_t tile_row = ((size_t) (slot >> 7) * nkb + (size_t) kp * 4) * 128 + (slot & 127);
		uint4 *dst = reinterpret_cast<uint4 *>(X + tile_row * 128 + sub * 16);
		__stcs(dst + 0 * 1024, make_uint4(c0[0], c0[1], c0[2], c0[3]));
		__stcs(dst + 1 * 1024, make_uint4(c1[0], c1[1], c1[2], c1[3]));
		__stcs(dst + 2 * 1024, make_uint4(c2[0], c2[1], c2[2], c2[3]));
		__stcs(dst + 3 * 1024, make_uint4(c3[0], c3[1], c3[2], c3[3]));
	}
	if(masks) {
		known += __shfl_xor_sync(0xffffffffu, known, 4);
		known += __shfl_xor_sync(0xffffffffu, known, 2);
		known += __shfl_xor_sync(0xffffffffu, known, 1);
		if(sub == 0 && known) atomicAdd(inc + slot, known);
	}
}

/* rows of the slab's panel that no streamed batch wrote (empty or excluded slots of a row block that is read): zeros */
__global__ void __launch_bounds__(256)
k_zero_panel_rows(const int *__restrict__ slots, int nslots, int npairs, int8_t *__restrict__ X, size_t nkb) {
	const long long total = (long long) nslots * npairs * 4 * 8;          /* 16-byte pieces */
	for(long long e = (long long) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x) {
		const int piece = (int) (e & 7), ch = (int) ((e >> 3) & 3);
		const long long rest = e >> 5;
		const int slot = slots[rest % nslots];
		const long long kp = rest / nslots;
		const size_t tile_row = ((size_t) (slot >> 7) * nkb + (size_t) kp * 4 + ch) * 128 + (slot & 127);
		reinterpret_cast<uint4 *>(X + tile_row * 128)[piece] = make_uint4(0, 0, 0, 0);
	}
}

/* ------------------------------------------------------------------ */
/* the GEMM                                                            */
/* ------------------------------------------------------------------ */
/* Lock-step throttle.  The CTAs of one round work on neighbouring tiles of the Z-order curve and
 * read the same operand row blocks; that reuse only reaches L2 (instead of HBM) while they are at
 * nearly the same K position.  Left alone they drift apart -- at 10,000 x 5 Mbp the panel was read
 * 33 times from HBM and the kernel ran at the HBM roof (profiles/r01_umma_10k_dram_bound.csv).
 * Every LOCK_E stages a producer announces the epoch it enters and waits (bounded) until every
 * CTA of the round has entered epoch - LOCK_LAG.  The wait is a throttle, not a correctness
 * barrier: on time-out the CTA simply proceeds. */
constexpr int LOCK_E = 8;
constexpr int LOCK_LAG = 2;

__device__ __forceinline__ void lockstep_arrive(unsigned *sync, long long g) { atomicAdd(sync + g, 1u); }
/* returns false on time-out (some CTA of the round is far behind or not resident): the caller then
 * runs LOCK_BACKOFF epochs unthrottled before it waits again */
constexpr int LOCK_BACKOFF = 64;
__device__ __forceinline__ bool lockstep_wait(const unsigned *sync, long long g, unsigned expected) {
	const volatile unsigned *c = sync + g;
	if(*c >= expected) return true;
	const long long t0 = clock64();
	while(*c < expected) {
		if(clock64() - t0 > 200000LL) return false;     /* ~0.1 ms */
	}
	return true;
}

/* Persistent: gridDim.x CTAs (one per SM); CTA b handles work items b, b + grid, b + 2 grid, ...
 * Work item w = (tile w % ntiles, K slice w / ntiles), so the items of one round are consecutive
 * tiles of the curve on the same K range. */
__global__ void __launch_bounds__(THREADS, 1)
k_pairdist_umma(const __grid_constant__ CUtensorMap tmap, const UmmaParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t raw = smem_u32(smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;                /* SWIZZLE_128B tiles need 1024-byte alignment */
	const uint32_t bar_full = base + STAGES * STAGE_BYTES;       /* STAGES x 8 B */
	const uint32_t bar_empty = bar_full + 8 * STAGES;            /* STAGES x 8 B */
	const uint32_t bar_accum = bar_empty + 8 * STAGES;           /* 8 B: accumulators of the item complete */
	const uint32_t bar_tfree = bar_accum + 8;                    /* 8 B: TMEM drained, next item may accumulate */
	const uint32_t tmem_slot = bar_tfree + 8;                    /* 4 B */
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int G = (int) gridDim.x;
	const int items = p.ntiles * p.kslices;
	const int nkb_slab = p.slab_chunks * 4;

	if(threadIdx.x == 0) {
		for(int s = 0; s < STAGES; ++s) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, 1);
		}
		mbar_init(bar_accum, 1);
		mbar_init(bar_tfree, 4);
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
	if(threadIdx.x == 0 && p.resident) atomicAdd(p.resident, 1u);    /* this CTA holds its SM: see run_umma */

	if(warp == 0) {
		/* ===== TMA producer ===== */
		if(lane == 0) {
			unsigned it = 0;                                      /* stages issued so far, over all items */
			int round = 0;
			int lock_skip = 0;
			for(int w = blockIdx.x; w < items; w += G, ++round) {
				const int tile = w % p.ntiles, ks = w / p.ntiles;
				const int tm = p.tiles[tile].x, tn = p.tiles[tile].y;
				const int c_begin = ks * p.chunks_per_slice;
				int nchunk = p.slab_chunks - c_begin;
				if(nchunk > p.chunks_per_slice) nchunk = p.chunks_per_slice;
				if(nchunk < 0) nchunk = 0;
				const int nkb = nchunk * 4;                       /* k-blocks: 4 channel rows per chunk */
				const long long g0 = (long long) round * p.epochs_per_item;
				for(int kb = 0; kb < nkb; ++kb, ++it) {
					if(p.sync && (kb % LOCK_E) == 0) {
						const long long g = g0 + kb / LOCK_E;
						lockstep_arrive(p.sync, g);
						if(lock_skip > 0) --lock_skip;
						else if(g >= LOCK_LAG) {
							const long long gw = g - LOCK_LAG;
							const int rw = (int) (gw / p.epochs_per_item);
							const int left = items - rw * G;
							if(!lockstep_wait(p.sync, gw, (unsigned) (left < G ? left : G))) lock_skip = LOCK_BACKOFF;
						}
					}
					const int s = it % STAGES;
					if(it >= STAGES) mbar_wait(bar_empty + 8 * s, ((it / STAGES) - 1) & 1, p.watchdog);
					const uint32_t dst = base + s * STAGE_BYTES;
					const uint32_t bar = bar_full + 8 * s;
					/* row coordinate of the 128-row box of row block rb, k-block kabs in the blocked panel */
					const int kabs = c_begin * 4 + kb;
					mbar_expect_tx(bar, STAGE_BYTES);
					tma_load_2d(dst, &tmap, bar, 0, p.row_base + (tm * nkb_slab + kabs) * 128);
					tma_load_2d(dst + A_BYTES, &tmap, bar, 0, p.row_base + ((2 * tn) * nkb_slab + kabs) * 128);
					tma_load_2d(dst + A_BYTES + 128 * BK, &tmap, bar, 0, p.row_base + ((2 * tn + 1) * nkb_slab + kabs) * 128);
				}
				/* a short (last) K slice still counts in every epoch of its round */
				if(p.sync)
					for(int e = (nkb + LOCK_E - 1) / LOCK_E; e < p.epochs_per_item; ++e) lockstep_arrive(p.sync, g0 + e);
			}
		}
	} else if(warp == 1) {
		/* ===== MMA issuer ===== */
		if(lane == 0) {
			unsigned it = 0;
			int round = 0;
			for(int w = blockIdx.x; w < items; w += G, ++round) {
				const int ks = w / p.ntiles;
				const int c_begin = ks * p.chunks_per_slice;
				int nchunk = p.slab_chunks - c_begin;
				if(nchunk > p.chunks_per_slice) nchunk = p.chunks_per_slice;
				if(nchunk < 0) nchunk = 0;
				const int nkb = nchunk * 4;
				if(round > 0) {
					mbar_wait(bar_tfree, (round - 1) & 1, p.watchdog);           /* the previous item's accumulators were read */
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				}
				uint32_t usedS = 0, usedI = 0;
				for(int kb = 0; kb < nkb; ++kb, ++it) {
					const int s = it % STAGES;
					mbar_wait(bar_full + 8 * s, (it / STAGES) & 1, p.watchdog);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t a0 = base + s * STAGE_BYTES;
					const uint64_t adesc = make_desc(a0);
					const uint64_t bdesc = make_desc(a0 + A_BYTES);
					const bool is_mask = (kb & 3) == 3;
					const uint32_t d = tmem + (is_mask ? BN : 0);
#pragma unroll
					for(int k = 0; k < BK / 32; ++k) {
						const uint32_t acc = is_mask ? usedI : usedS;
						/* advancing K by 32 bytes inside the 128-byte swizzle row: +2 in the (addr>>4) field */
						umma_i8(d, adesc + 2 * k, bdesc + 2 * k, acc);
						if(is_mask) usedI = 1; else usedS = 1;
					}
					umma_commit(bar_empty + 8 * s);        /* frees the stage once these MMAs have read it */
				}
				umma_commit(bar_accum);                     /* accumulators of this item complete */
			}
		}
	} else {
		/* ===== epilogue: TMEM -> registers -> RED.ADD into C ===== */
		const int quarter = warp & 3;                   /* a warp may only touch TMEM lanes 32*(warp%4).. */
		int round = 0;
		for(int w = blockIdx.x; w < items; w += G, ++round) {
			const int tile = w % p.ntiles, ks = w / p.ntiles;
			const int tm = p.tiles[tile].x, tn = p.tiles[tile].y;
			const int nleft = p.slab_chunks - ks * p.chunks_per_slice;
			const int row = tm * BM + quarter * 32 + lane;
			mbar_wait(bar_accum, round & 1, p.watchdog);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if(nleft > 0) {
				int *cS = p.C_S + (size_t) row * p.ldc + tn * BN;
				int *cI = p.C_I + (size_t) row * p.ldc + tn * BN;
				const int jlim = row - tn * BN;              /* only columns j < row are ever read back */
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
			/* all tcgen05.ld of this warp have completed (wait::ld inside tmem_ld32) */
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if(lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tfree) : "memory");
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if(warp == 2) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
	}
}

/* ------------------------------------------------------------------ */
/* the GEMM on CTA pairs (cta_group::2)                                */
/* ------------------------------------------------------------------ */
/* A cluster of two CTAs (one TPC) owns a 256 x 256 macro tile: tcgen05.mma.cta_group::2 with
 * M = 256, N = 256.  CTA r of the pair stages A rows [128 r, 128 r + 128) and B columns
 * [128 r, 128 r + 128) of the tile -- 32 KiB per stage instead of 48 KiB, so six stages fit and
 * each SM needs one third less L2 bandwidth for the same tensor work (the single-CTA kernel
 * above was limited by bytes in flight: 72-84 % tensor-pipe activity).  Both CTAs issue TMA,
 * all loads of a stage complete on the leader's full barrier; the leader (cluster rank 0)
 * issues the MMAs and its commits arrive on both CTAs' empty / accumulator barriers. */
constexpr int STAGES2 = 6;
constexpr int B2_BYTES = 128 * BK;
constexpr int STAGE2_BYTES = A_BYTES + B2_BYTES;              /* per CTA */
constexpr uint32_t IDESC2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (BN >> 3) << 17) | ((uint32_t) (256 >> 4) << 24);
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;                   /* shared::cluster address of the even CTA of the pair */

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1) {
	asm volatile(
	    "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
	    ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
	    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(IDESC2), "r"(accumulate) : "memory");
}
/* kind::mxf4: e2m1 operands, UE8M0 scale vectors (all 1.0, see k_pairdist_umma2<true>), f32 accumulators.
 * Block-scaled instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptorBlockScaled):
 * a/b_format E2M1 (1) @7/@10, K-major, n_dim @17, scale_format UE8M0 (1) @23, m_dim @24, K = 64. */
constexpr uint32_t IDESC_MXF4 = (1u << 7) | (1u << 10) | ((uint32_t) (BN >> 3) << 17) | (1u << 23) | ((uint32_t) (256 >> 4) << 24);
/* the same descriptor for an N of n columns (a multiple of 16): thin items, see k_pairdist_umma2 */
__device__ __forceinline__ uint32_t idesc_mxf4(int n) {
	return (1u << 7) | (1u << 10) | ((uint32_t) (n >> 3) << 17) | (1u << 23) | ((uint32_t) (256 >> 4) << 24);
}
__device__ __forceinline__ void umma2_mxf4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t sfa, uint32_t sfb,
                                           uint32_t idesc) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
	    ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
	             ::"r"(bar), "h"((uint16_t) 3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

/* FP4 = true: the same kernel on the e2m1 panel.  The f32 accumulator leaves no room for two 256-column
 * accumulators plus the scale vectors in the 512 TMEM columns, so S (channels 0-2) and I (mask channel) are
 * separate work items (first the S items of every K slice, then the I items); p.slab_chunks / p.chunks_per_slice count chunk
 * PAIRS (one 128-byte channel row = 256 bases).  Exactness: every product is -1, 0 or +1, the pipe adds exact
 * partial sums into f32 (measured: scripts/fp4_probe.py), so an item is exact while 3 * 256 * pairs < 2^24 --
 * the host keeps the slices below 20,000 pairs. */
/* THIN items (FP4 only).  The last macro-tile row of a sample count that is not a multiple of 256 holds only
 * v = n - 256 tm valid rows (16 of 256 at n = 10,000: 4.8 % of all MMA work would be padding).  When v is small
 * (p.thin_n = v rounded up to 16, p.thin_tm = that tile row) the off-diagonal tiles of that row are computed
 * TRANSPOSED: the 256 samples of the column block are the M side (A operand, one row block per CTA as usual), the v
 * valid samples of the row block are the N side -- CTA r of the pair stages rows [r N/2, r N/2 + N/2) of that row
 * block as its half of B through a tensor map whose box is N/2 rows (tmap_thin) -- and the MMA runs with N = thin_n.
 * TMEM lane = column sample j, TMEM column e = row sample 256 tm + e, so the epilogue adds into C[256 tm + e][j]:
 * the same cells, just reached from the other side (and with the 32 lanes of a warp on 32 consecutive ints). */
template <bool FP4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
k_pairdist_umma2(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_thin, const UmmaParams p) {
	extern __shared__ uint8_t smem_raw[];
	const uint32_t raw = smem_u32(smem_raw);
	const uint32_t base = (raw + 1023u) & ~1023u;
	const uint32_t bar_full = base + STAGES2 * STAGE2_BYTES;     /* used on the leader only */
	const uint32_t bar_empty = bar_full + 8 * STAGES2;
	const uint32_t bar_accum = bar_empty + 8 * STAGES2;
	const uint32_t bar_tfree = bar_accum + 8;                    /* used on the leader only: 8 epilogue warps */
	const uint32_t tmem_slot = bar_tfree + 8;
	volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - raw));

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	uint32_t cta_rank;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
	const int G = (int) (gridDim.x >> 1);                        /* clusters */
	const int cid = (int) (blockIdx.x >> 1);
	const int items = p.ntiles * p.kslices * ((FP4 && !p.no_mask_items) ? 2 : 1);   /* shared-mask mode: I is a constant */
	const int nkb_slab = p.slab_chunks * 4;

	if(threadIdx.x == 0) {
		for(int s = 0; s < STAGES2; ++s) {
			mbar_init(bar_full + 8 * s, 1);
			mbar_init(bar_empty + 8 * s, 1);
		}
		mbar_init(bar_accum, 1);
		mbar_init(bar_tfree, 8);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	if(warp == 2) {
		asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	cluster_sync_all();                                          /* both CTAs' barriers and TMEM are ready */
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_slot_ptr;
	if(threadIdx.x == 0 && p.resident) atomicAdd(p.resident, 1u);
	if(FP4) {
		/* scale vectors: UE8M0 1.0 (0x7F) in every byte of TMEM columns [256, 288) of every lane, whichever
		 * bytes the scale-factor ids select; written once, shared by all MMAs */
		if(warp >= 2) {
			const uint32_t one = 0x7F7F7F7Fu;
			const uint32_t taddr = tmem + ((uint32_t) ((warp & 3) * 32) << 16) + 256;
			for(int c = 0; c < 32; c += 8)
				asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(one) : "memory");
			asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
		}
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		cluster_sync_all();
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	}

	if(warp == 0) {
		/* ===== TMA producer (both CTAs) ===== */
		if(lane == 0) {
			unsigned it = 0;
			int round = 0;
			int lock_skip = 0;
			for(int w = cid; w < items; w += G, ++round) {
				const int tile = w % p.ntiles, q = w / p.ntiles;
				const int ks = FP4 ? q % p.kslices : q, which = FP4 ? q / p.kslices : 0;   /* all S items first, then the I items */
				const int tm = p.tiles[tile].x, tn = p.tiles[tile].y;
				const int c_begin = ks * p.chunks_per_slice;
				int nchunk = p.slab_chunks - c_begin;
				if(nchunk > p.chunks_per_slice) nchunk = p.chunks_per_slice;
				if(nchunk < 0) nchunk = 0;
				const int nkb = FP4 ? (which ? nchunk : 3 * nchunk) : nchunk * 4;   /* stages of this item */
				const long long g0 = (long long) round * p.epochs_per_item;
				const bool thin = FP4 && tm == p.thin_tm && tn < tm;
				const int rbA = thin ? 2 * tn + (int) cta_rank : 2 * tm + (int) cta_rank;
				const int rbB = thin ? 2 * tm : 2 * tn + (int) cta_rank;
				const int b_off = thin ? (int) cta_rank * (p.thin_n >> 1) : 0;      /* first row of this CTA's half of B */
				const unsigned stage_tx = thin ? (unsigned) (2 * (A_BYTES + (p.thin_n >> 1) * BK)) : (unsigned) (2 * STAGE2_BYTES);
				for(int kb = 0; kb < nkb; ++kb, ++it) {
					if(p.sync && cta_rank == 0 && (kb % LOCK_E) == 0) {
						const long long g = g0 + kb / LOCK_E;
						lockstep_arrive(p.sync, g);
						if(lock_skip > 0) --lock_skip;
						else if(g >= LOCK_LAG) {
							const long long gw = g - LOCK_LAG;
							const int rw = (int) (gw / p.epochs_per_item);
							const int left = items - rw * G;
							if(!lockstep_wait(p.sync, gw, (unsigned) (left < G ? left : G))) lock_skip = LOCK_BACKOFF;
						}
					}
					const int s = it % STAGES2;
					if(it >= STAGES2) mbar_wait(bar_empty + 8 * s, ((it / STAGES2) - 1) & 1, p.watchdog);
					const uint32_t dst = base + s * STAGE2_BYTES;
					const uint32_t bar = (bar_full + 8 * s) & PEER_MASK;      /* the leader's barrier */
					/* k-block of this stage inside the slab: 4 per chunk (pair), channel-minor */
					const int kabs = FP4 ? (which ? (c_begin + kb) * 4 + 3 : (c_begin + kb / 3) * 4 + kb % 3) : c_begin * 4 + kb;
					if(cta_rank == 0) mbar_expect_tx(bar_full + 8 * s, stage_tx);
					tma_load_2d_pair(dst, &tmap, bar, 0, p.row_base + (rbA * nkb_slab + kabs) * 128);
					tma_load_2d_pair(dst + A_BYTES, thin ? &tmap_thin : &tmap, bar, 0, p.row_base + (rbB * nkb_slab + kabs) * 128 + b_off);
				}
				if(p.sync && cta_rank == 0)
					for(int e = (nkb + LOCK_E - 1) / LOCK_E; e < p.epochs_per_item; ++e) lockstep_arrive(p.sync, g0 + e);
			}
		}
	} else if(warp == 1) {
		/* ===== MMA issuer (leader CTA only) ===== */
		if(lane == 0 && cta_rank == 0) {
			unsigned it = 0;
			int round = 0;
			for(int w = cid; w < items; w += G, ++round) {
				const int tile = w % p.ntiles, q = w / p.ntiles;
				const int ks = FP4 ? q % p.kslices : q, which = FP4 ? q / p.kslices : 0;   /* all S items first, then the I items */
				const int c_begin = ks * p.chunks_per_slice;
				int nchunk = p.slab_chunks - c_begin;
				if(nchunk > p.chunks_per_slice) nchunk = p.chunks_per_slice;
				if(nchunk < 0) nchunk = 0;
				const int nkb = FP4 ? (which ? nchunk : 3 * nchunk) : nchunk * 4;
				const bool thin = FP4 && p.tiles[tile].x == p.thin_tm && p.tiles[tile].y < p.tiles[tile].x;
				const uint32_t idesc = thin ? idesc_mxf4(p.thin_n) : IDESC_MXF4;
				if(round > 0) {
					mbar_wait(bar_tfree, (round - 1) & 1, p.watchdog);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				}
				uint32_t usedS = 0, usedI = 0;
				for(int kb = 0; kb < nkb; ++kb, ++it) {
					const int s = it % STAGES2;
					mbar_wait(bar_full + 8 * s, (it / STAGES2) & 1, p.watchdog);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t a0 = base + s * STAGE2_BYTES;
					const uint64_t adesc = make_desc(a0);
					const uint64_t bdesc = make_desc(a0 + A_BYTES);
					if(FP4) {
#pragma unroll
						for(int k = 0; k < BK / 32; ++k) {               /* 32 bytes = 64 e2m1 = one MMA K */
							umma2_mxf4(tmem, adesc + 2 * k, bdesc + 2 * k, usedS, tmem + 256, tmem + 264, idesc);
							usedS = 1;
						}
					} else {
						const bool is_mask = (kb & 3) == 3;
						const uint32_t d = tmem + (is_mask ? BN : 0);
#pragma unroll
						for(int k = 0; k < BK / 32; ++k) {
							const uint32_t acc = is_mask ? usedI : usedS;
							umma2_i8(d, adesc + 2 * k, bdesc + 2 * k, acc);
							if(is_mask) usedI = 1; else usedS = 1;
						}
					}
					umma2_commit(bar_empty + 8 * s);       /* frees the stage in both CTAs */
				}
				umma2_commit(bar_accum);                    /* both CTAs' epilogues may read their TMEM half */
			}
		}
	} else {
		/* ===== epilogue (both CTAs): TMEM -> registers -> RED.ADD into C ===== */
		const int quarter = warp & 3;
		uint32_t leader_tfree;
		asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(leader_tfree) : "r"(bar_tfree));
		int round = 0;
		for(int w = cid; w < items; w += G, ++round) {
			const int tile = w % p.ntiles, q = w / p.ntiles;
			const int ks = FP4 ? q % p.kslices : q, which = FP4 ? q / p.kslices : 0;   /* all S items first, then the I items */
			const int tm = p.tiles[tile].x, tn = p.tiles[tile].y;
			const int nleft = p.slab_chunks - ks * p.chunks_per_slice;
			const int row = tm * BMT + (int) cta_rank * 128 + quarter * 32 + lane;
			const bool thin = FP4 && tm == p.thin_tm && tn < tm;
			mbar_wait(bar_accum, round & 1, p.watchdog);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			if(nleft > 0 && FP4) {
				/* One loop for both item shapes (a second copy of it costs 140 registers, which the kernels that run
				 * beside this one -- operand expansion, slab uploads -- then lack).  Ordinary item: TMEM lane = matrix row,
				 * TMEM column e = matrix column tn * 256 + e, consecutive ints, only columns below the row are needed.
				 * Thin (transposed) item: TMEM lane = column sample j, TMEM column e = row sample 256 tm + e, so the
				 * 32 values of a thread are one matrix column: stride ldc, the first thin_valid of them exist. */
				const int j = tn * BN + (int) cta_rank * 128 + quarter * 32 + lane;
				int *cp = (which ? p.C_I : p.C_S) + (thin ? (size_t) (tm * BMT) * p.ldc + j : (size_t) row * p.ldc + tn * BN);
				const int stride = thin ? p.ldc : 1;
				const int lim = thin ? p.thin_valid : row - tn * BN;
				const int ncol = thin ? p.thin_n : BN;
#pragma unroll 1
				for(int cb = 0; cb * 32 < ncol; ++cb) {
					if(__all_sync(0xffffffffu, cb * 32 >= lim)) break;
					uint32_t r[32];
					tmem_ld32(tmem + ((uint32_t) (quarter * 32) << 16) + cb * 32, r);
					const int left = lim - cb * 32;
#pragma unroll
					for(int e = 0; e < 32; ++e) {
						const int v = __float2int_rn(__uint_as_float(r[e]));       /* exact integers in f32 */
						if(e < left && v) atomicAdd(cp, v);
						cp += stride;
					}
				}
			} else if(nleft > 0) {
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
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if(lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_tfree) : "memory");
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	cluster_sync_all();                                          /* the peer's smem / barriers stay alive until both are done */
	if(warp == 2) {
		asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
	}
}

/* mismatch = (3I - S)/4, then the reference epilogue.  Pair mode: I comes from the mask
 * channel.  Shared-mask mode: the planes are pre-masked and expanded with m = 1 everywhere,
 * so every masked / padded position is an (A, A) match and I is the constant i_const =
 * chunks * 128 (then 3*i_const - S = 4 * mismatch as well). */
__global__ void __launch_bounds__(256)
k_finalize_umma(const int *__restrict__ C_S, const int *__restrict__ C_I, int ldc, int n, int pair_mode, int i_const,
                const int2 *__restrict__ tiles, int ntiles, EpilogueParams ep) {
	const int tile = blockIdx.x;
	if(tile >= ntiles) return;
	const int tm = tiles[tile].x, tn = tiles[tile].y;
	for(int e = threadIdx.x; e < BMT * BN; e += blockDim.x) {
		const int i = tm * BMT + e / BN;
		const int j = tn * BN + e % BN;
		if(i >= n || j >= i) continue;
		const int S = C_S[(size_t) i * ldc + j];
		const int I = pair_mode ? C_I[(size_t) i * ldc + j] : i_const;
		const unsigned mism = (unsigned) ((3 * (long long) I - S) >> 2);
		ccg_write_cell(ep, i, j, mism, (unsigned) I);
	}
}

/* raw integer counts of the last run -> packed lower triangle over included samples */
__global__ void __launch_bounds__(256)
k_gather_raw_dense(const int *__restrict__ C_S, const int *__restrict__ C_I, int ldc, int n, int pair_mode, int i_const,
                   const int2 *__restrict__ tiles, int ntiles, const int *__restrict__ rank,
                   uint32_t *__restrict__ mism, uint32_t *__restrict__ ninc) {
	const int tile = blockIdx.x;
	if(tile >= ntiles) return;
	const int tm = tiles[tile].x, tn = tiles[tile].y;
	for(int e = threadIdx.x; e < BMT * BN; e += blockDim.x) {
		const int i = tm * BMT + e / BN;
		const int j = tn * BN + e % BN;
		if(i >= n || j >= i) continue;
		const int r = rank[i], c = rank[j];
		if(r < 0 || c < 0) continue;
		const long long cell = (long long) r * (r - 1) / 2 + c;
		const int S = C_S[(size_t) i * ldc + j];
		const int I = pair_mode ? C_I[(size_t) i * ldc + j] : i_const;
		if(mism) mism[cell] = (unsigned) ((3 * (long long) I - S) >> 2);
		if(ninc) ninc[cell] = (unsigned) I;
	}
}

} // namespace

/* expands the row blocks this rank needs (runs of consecutive needed 128-slot blocks) */
/* e2m1 panel: npairs chunk pairs from chunk0 on (chunks past the end of the alignment expand to zeros) */
cudaError_t ccg_launch_expand_fp4(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, int chunk0, int npairs, int bounded) {
	const int nblocks = ctx->n_pad / 128;
	int b = 0;
	while(b < nblocks) {
		if(!ctx->need[b]) { ++b; continue; }
		int e = b;
		while(e < nblocks && ctx->need[e]) ++e;
		const int slot0 = b * 128, slots = (e - b) * 128;
		const long long items = (long long) slots * npairs;
		if(items > 0) {
			long long blocks = (items * 8 + 255) / 256;
			if(bounded && blocks > 6LL * ctx->sm_count) blocks = 6LL * ctx->sm_count;
			PackedSource src;
			memset(&src, 0, sizeof(src));
			if(ctx->bor_pending) {
				src.seqs = ctx->bor_seqs;
				src.masks = ctx->bor_masks;
				src.gmask = ctx->d_gmask;
				src.wstride = ctx->bor_wstride;
				src.first = ctx->bor_first;
				src.count = ctx->bor_count;
				src.words = ctx->words;
			}
			k_expand_fp4<<<(unsigned) blocks, 256, 0, stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, ctx->chunks, slot0, slots, chunk0,
			                                                   npairs, X, (size_t) npairs * 4, src);
			ctx->launches++;
		}
		b = e;
	}
	return cudaGetLastError();
}

/* streamed batch of rows [first, first + count) (words of the slab that starts at chunk0) -> the slab's panel */
cudaError_t ccg_launch_expand_rows(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, int first, int count, const uint64_t *d_seqs,
                                   const uint32_t *d_masks, long wstride, int chunk0, int npairs) {
	if(count <= 0 || npairs <= 0) return cudaSuccess;
	int nvalid = ctx->words - chunk0 * CCG_CHUNK_WORDS;
	if(nvalid > npairs * 8) nvalid = npairs * 8;
	if(nvalid < 0) nvalid = 0;
	/* a few hundred pairs per thread: enough blocks to keep the copy stream's share of the SMs busy beside the GEMM */
	int per = 256;
	if(per > npairs) per = npairs;
	dim3 grid((unsigned) (((long long) count * 8 + 255) / 256), (unsigned) ((npairs + per - 1) / per));
	k_expand_fp4_rows<<<grid, 256, 0, stream>>>(d_seqs, d_masks, ctx->d_gmask ? ctx->d_gmask + (size_t) chunk0 * CCG_CHUNK_WORDS : 0, wstride, nvalid,
	                                            first, count, npairs, per, ctx->nplanes, X, (size_t) npairs * 4, ctx->d_inc);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_zero_panel_rows(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, const int *d_slots, int nslots, int npairs) {
	if(nslots <= 0 || npairs <= 0) return cudaSuccess;
	long long blocks = ((long long) nslots * npairs * 32 + 255) / 256;
	if(blocks > 2LL * ctx->sm_count) blocks = 2LL * ctx->sm_count;
	k_zero_panel_rows<<<(unsigned) blocks, 256, 0, stream>>>(d_slots, nslots, npairs, X, (size_t) npairs * 4);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_expand(ccg_ctx *ctx, cudaStream_t stream, int8_t *X, int chunk0, int nchunks, int bounded) {
	const int nblocks = ctx->n_pad / 128;
	int b = 0;
	while(b < nblocks) {
		if(!ctx->need[b]) { ++b; continue; }
		int e = b;
		while(e < nblocks && ctx->need[e]) ++e;
		const int slot0 = b * 128, slots = (e - b) * 128;
		const long long items = (long long) slots * nchunks;
		if(items > 0) {
			long long blocks = (items * 8 + 255) / 256;
			/* when a GEMM of the previous slab runs concurrently, leave it room on every SM */
			if(bounded && blocks > 6LL * ctx->sm_count) blocks = 6LL * ctx->sm_count;
			k_expand<<<(unsigned) blocks, 256, 0, stream>>>(ctx->d_planes, ctx->n_pad, ctx->nplanes, slot0, slots, chunk0, nchunks, X,
			                                      (size_t) nchunks * 4);
			ctx->launches++;
		}
		b = e;
	}
	return cudaGetLastError();
}

/* CTA pairs of k_pairdist_umma2 that can be resident at once (a pair needs both SMs of a TPC) */
int ccg_umma_pair_slots(ccg_ctx *ctx) {
	constexpr int smem2 = STAGES2 * STAGE2_BYTES + 8 * (2 * STAGES2 + 2) + 16 + 1024;
	if(!ctx->max_pairs) {
		cudaFuncSetAttribute(k_pairdist_umma2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3((unsigned) (2 * (ctx->sm_count / 2)));
		cfg.blockDim = dim3(THREADS);
		cfg.dynamicSmemBytes = smem2;
		cudaLaunchAttribute attr;
		attr.id = cudaLaunchAttributeClusterDimension;
		attr.val.clusterDim.x = 2;
		attr.val.clusterDim.y = 1;
		attr.val.clusterDim.z = 1;
		cfg.attrs = &attr;
		cfg.numAttrs = 1;
		int nclusters = 0;
		if(cudaOccupancyMaxActiveClusters(&nclusters, k_pairdist_umma2<false>, &cfg) != cudaSuccess || nclusters < 1) {
			cudaGetLastError();
			nclusters = ctx->sm_count / 2;
		}
		ctx->max_pairs = nclusters < ctx->sm_count / 2 ? nclusters : ctx->sm_count / 2;
	}
	return ctx->max_pairs;
}

/* p.tiles: 256 x 256 macro tiles for the CTA-pair kernel; 128 x 256 tiles (p.single != 0) for the
 * single-CTA kernel */
cudaError_t ccg_launch_umma(ccg_ctx *ctx, const UmmaParams &p_in) {
	constexpr int smem1 = STAGES * STAGE_BYTES + 8 * (2 * STAGES + 2) + 16 + 1024;
	constexpr int smem2 = STAGES2 * STAGE2_BYTES + 8 * (2 * STAGES2 + 2) + 16 + 1024;
	cudaError_t e = cudaFuncSetAttribute(k_pairdist_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
	if(e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(k_pairdist_umma2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
	if(e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(k_pairdist_umma2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
	if(e != cudaSuccess) return e;
	UmmaParams p = p_in;
	p.watchdog = ctx->watchdog_cycles;
	const long long items = (long long) p.ntiles * p.kslices * ((p.fp4 && !p.no_mask_items) ? 2 : 1);
	if(items <= 0) return cudaSuccess;
	const int slots = p.single ? ctx->sm_count : ccg_umma_pair_slots(ctx);   /* CTAs or CTA pairs */
	const int grid = items < slots ? (int) items : slots;
	/* lock-step counters: one per (round, epoch of LOCK_E stages) */
	const long long rounds = (items + grid - 1) / grid;
	p.epochs_per_item = (p.chunks_per_slice * (p.fp4 ? 3 : 4) + LOCK_E - 1) / LOCK_E;
	const size_t need = (size_t) rounds * p.epochs_per_item + 1;
	p.sync = 0;
	if(!ctx->dbg_nolock && grid > 1 && need <= ((size_t) 64 << 20)) {
		if(ctx->sync_cap < need) {
			if((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return e;
			cudaFree(ctx->d_sync);
			ctx->d_sync = 0;
			ctx->sync_cap = 0;
			if((e = cudaMalloc(&ctx->d_sync, need * sizeof(unsigned))) != cudaSuccess) return e;
			ctx->sync_cap = need;
		}
		if((e = cudaMemsetAsync(ctx->d_sync, 0, need * sizeof(unsigned), ctx->stream)) != cudaSuccess) return e;
		p.sync = ctx->d_sync;
	}
	ctx->last_gemm_ctas = p.single ? grid : 2 * grid;
	/* thin items: the last macro-tile row when it holds few valid samples (see k_pairdist_umma2) */
	p.thin_tm = -1;
	p.thin_n = p.thin_valid = 0;
	const int v = ctx->n % BMT;
	if(p.fp4 && !p.single && v > 0 && v <= 96 && ctx->n > BMT && !ctx->dbg_nothin) {
		const int tn_thin = (v + 15) / 16 * 16;
		if(ctx->tmap_thin_rows != tn_thin / 2 || !ctx->tmap_thin_valid) {
			cudaError_t et = ccg_make_thin_tmap(ctx, tn_thin / 2);
			if(et != cudaSuccess) return et;
		}
		p.thin_tm = ctx->n / BMT;
		p.thin_n = tn_thin;
		p.thin_valid = v;
	}
	if(p.single) k_pairdist_umma<<<(unsigned) grid, THREADS, smem1, ctx->stream>>>(ctx->tmap_x, p);
	else if(p.fp4) k_pairdist_umma2<true><<<(unsigned) (2 * grid), THREADS, smem2, ctx->stream>>>(ctx->tmap_x, p.thin_tm >= 0 ? ctx->tmap_thin : ctx->tmap_x, p);
	else k_pairdist_umma2<false><<<(unsigned) (2 * grid), THREADS, smem2, ctx->stream>>>(ctx->tmap_x, ctx->tmap_x, p);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_finalize_umma(ccg_ctx *ctx, const UmmaParams &p, const EpilogueParams &ep, int i_const) {
	if(p.ntiles <= 0) return cudaSuccess;
	k_finalize_umma<<<p.ntiles, 256, 0, ctx->stream>>>(p.C_S, p.C_I, p.ldc, ctx->n, ctx->pair_mode, i_const, p.tiles,
	                                                   p.ntiles, ep);
	ctx->launches++;
	return cudaGetLastError();
}

cudaError_t ccg_launch_gather_raw_dense(ccg_ctx *ctx, int i_const, uint32_t *d_mism, uint32_t *d_ninc) {
	if(ctx->last_ntiles <= 0) return cudaSuccess;
	int *C_S = ctx->d_C;
	int *C_I = ctx->d_C + (size_t) ctx->n_pad * ctx->n_pad;
	k_gather_raw_dense<<<ctx->last_ntiles, 256, 0, ctx->stream>>>(C_S, C_I, ctx->n_pad, ctx->n, ctx->pair_mode, i_const,
	                                                              ctx->d_tiles, ctx->last_ntiles, ctx->d_rank, d_mism,
	                                                              d_ninc);
	ctx->launches++;
	return cudaGetLastError();
}
