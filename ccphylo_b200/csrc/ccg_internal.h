/*
 * ccg_internal.h -- shared declarations of the CUDA side (not part of the ABI).
 *
 * Device-resident sample store ("planes"), one per context:
 *
 *     planes[chunk][plane][slot][4]   u32
 *
 *   chunk  = 128 consecutive bases (4 words of 32 bases)
 *   plane  = 0: high bit of the 2-bit code, 1: low bit, 2: inclusion mask
 *            (pair mode: 3 planes; shared-mask mode: 2 planes, the global mask
 *             is folded into the code planes at encode time)
 *   slot   = sample index, padded to a multiple of CCG_TILE
 *   bit 31-(p%32) of word (p/32)%4 <-> base p (the reference's mask bit order,
 *   fsacmp.c:164).  Code bits of masked positions are cleared, so a plane word
 *   of an absent / padded slot is all-zero and contributes nothing.
 *
 * One (chunk, plane) row of CCG_TILE slots is 1 KiB contiguous, every thread
 * access is a 128-bit vector, and a [KC chunks][planes][CCG_TILE slots][4]
 * box is a single TMA tile.
 */
#ifndef CCG_INTERNAL_H
#define CCG_INTERNAL_H

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ccphylo_gpu.h"

#define CCG_TILE 64          /* samples per tile edge (popc kernel) */
#define CCG_CHUNK_WORDS 4    /* u32 words per plane per chunk */
#define CCG_CHUNK_BASES 128

struct EpilogueParams {
	int mode;              /* 0 pair (fsacmpthrd.c:419-475), 1 global (fsacmpthrd.c:247-255) */
	int elem_size;         /* 8, 4, 2, 1 */
	unsigned norm;
	unsigned minLength;    /* final gate (already maxed with minCov*len) */
	double byteScale;
	double nFactor;        /* global mode: norm ? norm/inc : 1.0 */
	void *D;               /* device, packed over included samples */
	void *N;               /* device or NULL */
	const int *rank;       /* slot -> compact index, -1 = excluded */
};

struct PopcParams {
	int n_pad;             /* slots (multiple of CCG_TILE) */
	int chunks;            /* ceil(words/4) */
	int ksplit;            /* K slices per tile */
	int chunks_per_split;  /* multiple of KC */
	int ntiles_local;      /* tiles owned by this rank */
	int rank, world;       /* tile t is owned by rank t % world */
	uint32_t *acc;         /* [ntiles_local][2][TILE*TILE] raw mism / ninc */
	unsigned *tickets;     /* [ntiles_local] */
	EpilogueParams ep;
};

struct ccg_ctx {
	int device;
	int sm_count;
	cudaStream_t own_stream, stream;
	int kernel_choice;
	int rank, world;

	int n, len, pair_mode;
	int words, chunks, n_pad, nplanes;
	uint32_t *d_planes;
	size_t planes_bytes;
	uint32_t *d_gmask;         /* shared-mask mode: [words] */
	unsigned global_inc;
	unsigned *d_inc;           /* [n_pad] per-slot included counts */
	unsigned char *present;    /* host [n]: slot uploaded */

	void *d_stage;             /* staging for host rows */
	size_t stage_bytes;

	int *d_rank;               /* [n_pad] */
	int *h_rank;               /* host mirror of the last run */
	int last_Dn;
	uint32_t *d_acc;
	size_t acc_bytes;
	unsigned *d_tickets;
	size_t tickets_count;
	int last_ntiles_local;
	void *d_out_D, *d_out_N;   /* device result buffers for the host-output API */
	size_t out_bytes;

	CUtensorMap tmap;          /* planes as a 4-D tensor */
	int tmap_valid;

	cudaEvent_t ev0, ev1;
	int ev_valid;
	long long launches;
	char last_kernel[64];
	char err[512];
};

/* ---- launchers implemented in the kernel translation units ---- */

/* k_encode.cu */
cudaError_t ccg_launch_repack(ccg_ctx *ctx, int first, int count, const uint64_t *d_seqs,
                              const uint32_t *d_masks, long wstride);
cudaError_t ccg_launch_encode_codes(ccg_ctx *ctx, int first, int count, const unsigned char *d_codes,
                                    long stride);
cudaError_t ccg_launch_gather_raw(ccg_ctx *ctx, int Dn, uint32_t *d_mism, uint32_t *d_ninc);

/* k_pairdist_popc.cu */
cudaError_t ccg_launch_popc(ccg_ctx *ctx, const PopcParams &p);
int ccg_popc_kc(void);

/* tile bookkeeping shared by host and device */
static inline __host__ __device__ long long ccg_tiles_total(int ntile_rows) {
	return (long long) ntile_rows * (ntile_rows + 1) / 2;
}
static inline __host__ __device__ long long ccg_tiles_local(long long total, int rank, int world) {
	return total > rank ? (total - rank + world - 1) / world : 0;
}

#endif
