/*
 * ccg_group.cu -- the multi-GPU side of the C-ABI: the K-split group and the in-process multi-GPU context.
 *
 * The reference's fan-out is one C call that hands pairs to `tnum` threads sharing the heap
 * (fsaCmpThreadOut fsacmpthrd.c:76-106, pair loop of cmpairFsaThrd :261-480).  Across GPUs the unit that is
 * handed out here is not the pair but the ALIGNMENT AXIS: member g of a group of `world` GPUs owns the bases
 * [b_g, b_g+1) of EVERY sample, runs the whole lower triangle on that slice with the ordinary tensor-core pair
 * kernel, and the int32 partial sums S and I of the members are added up -- integer split-K is exact and
 * order-independent -- by the member that OWNS a matrix row, inside its epilogue kernel, reading the peers'
 * accumulators through peer-mapped pointers over NVLink (k_finalize_group: reduce-scatter fused with the
 * epilogue).  What that buys over the tile partition (ccg_set_partition): every member repacks / expands exactly
 * 1/world of the operands (a tile partition touches O(n / sqrt(world)) row blocks for 1/world of the tiles),
 * every member runs the same tile list (no tile imbalance), host rows reach the members as `world` parallel
 * uploads of disjoint word ranges, and there is no all-gather.
 *
 * Synchronisation is one device-side flag barrier per run (k_group_barrier: system-scope release / acquire on a
 * flag word per peer in every member's window) between "my accumulators are complete" and "I read everybody's".
 * The accumulators are double-buffered, so the barrier of run k also orders the zeroing of a buffer in run k+1
 * after every peer's reads of it in run k-1.
 *
 * Members may live in one process (ccg_init_multi: one host thread per device, peer access enabled) or in one
 * process per GPU (ccg_group_export / ccg_group_join: CUDA IPC handles exchanged by the caller, e.g. through
 * torch.distributed in bench.py).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "ccg_internal.h"
#include "epilogue.cuh"

#define CKG(ctx, call)                                                                                 \
	do {                                                                                               \
		cudaError_t e__ = (call);                                                                      \
		if(e__ != cudaSuccess) {                                                                       \
			ccg_set_err(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
			return CCG_ERR_CUDA;                                                                       \
		}                                                                                              \
	} while(0)

namespace {

struct GroupHandle {                /* what ccg_group_export writes into the caller's CCG_GROUP_HANDLE_BYTES */
	unsigned magic;
	int pid;
	int device;
	int npad_max;
	unsigned long long ptr;
	unsigned long long bytes;
	cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(GroupHandle) <= CCG_GROUP_HANDLE_BYTES, "handle does not fit");
constexpr unsigned GROUP_MAGIC = 0x43434747u;

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

/* Thread t talks to member t: everything this member's stream did before the barrier (its accumulator sums) is
 * released to t, and the kernel ends only once every member has announced the same epoch to this member.  A
 * member that never arrives (a failed peer) traps after q.timeout_ns (120 s, CCG_GROUP_TIMEOUT_S) instead of
 * hanging the GPU. */
__global__ void k_group_barrier(GroupBarrierParams q) {
	const int t = threadIdx.x;
	if(t >= q.world) return;
	__threadfence_system();
	q.hdr[t]->iconst_from[q.buf][q.rank] = q.i_const;
	st_release_sys(&q.hdr[t]->arrive[q.rank], q.epoch);
	if(!q.spin) return;             /* members sharing one device meet on the host instead (ccg_group_finalize) */
	const unsigned *mine = &q.hdr[q.rank]->arrive[t];
	const unsigned long long t0 = global_ns();
	while((int) (ld_acquire_sys(mine) - q.epoch) < 0) {
		__nanosleep(200);
		if(global_ns() - t0 > q.timeout_ns) __trap();
	}
}

__device__ __forceinline__ int4 ld_cg_int4(const int *p) { return __ldcg(reinterpret_cast<const int4 *>(p)); }

/* Reduce-scatter fused with the epilogue: matrix rows are owned in blocks of CCG_GROUP_ROW_BLOCK rows dealt
 * round-robin (row block b belongs to member b % world).  The owner adds the members' partial sums of its rows --
 * 128-bit loads, its own from HBM, the peers' over NVLink (L1 bypassed: the data belongs to another GPU's L2) --
 * and applies mismatch = (3I - S) / 4 and the reference epilogue (fsacmpthrd.c:419-475 / :247-255).
 * q.C[p] are biased pointers: element (i, j) of the matrix is at C[p][i * ldc + j] for the rows of this window. */
__global__ void __launch_bounds__(256)
k_finalize_group(const GroupFinalizeParams q, const EpilogueParams ep) {
	int i_const = 0;
	if(!q.pair_mode)
		for(int p = 0; p < q.world; ++p) i_const += q.own->iconst_from[q.buf][p];
	for(int i = q.row_lo + (int) blockIdx.x; i < q.row_hi; i += (int) gridDim.x) {
		if(i >= q.n) break;
		if((i / CCG_GROUP_ROW_BLOCK) % q.world != q.rank) continue;
		const long long off = (long long) i * q.ldc;
		for(int j4 = 4 * (int) threadIdx.x; j4 < i; j4 += 4 * (int) blockDim.x) {
			int4 S = make_int4(0, 0, 0, 0), I = make_int4(0, 0, 0, 0);
#pragma unroll 4
			for(int p = 0; p < q.world; ++p) {
				const int4 s = ld_cg_int4(q.C[p] + off + j4);
				S.x += s.x; S.y += s.y; S.z += s.z; S.w += s.w;
				if(q.pair_mode) {
					const int4 v = ld_cg_int4(q.C[p] + q.plane + off + j4);
					I.x += v.x; I.y += v.y; I.z += v.z; I.w += v.w;
				}
			}
			const int Sv[4] = {S.x, S.y, S.z, S.w};
			const int Iv[4] = {I.x, I.y, I.z, I.w};
#pragma unroll
			for(int e = 0; e < 4; ++e) {
				const int j = j4 + e;
				if(j >= i) break;
				const int inc = q.pair_mode ? Iv[e] : i_const;
				const unsigned mism = (unsigned) ((3 * (long long) inc - Sv[e]) >> 2);
				ccg_write_cell(ep, i, j, mism, (unsigned) inc);
			}
		}
	}
}

} // namespace

/* Members that share one process meet on the host before they launch the device barrier: everything a member
 * allocates or frees for the run (cudaFree waits for the whole device) has then happened, so -- also when several
 * members sit on ONE device, as in the single-GPU tests -- no member can block in the driver behind a peer's
 * spinning barrier kernel that is waiting for it. */
struct HostBarrier {
	std::mutex m;
	std::condition_variable cv;
	int count = 0, waiting = 0;
	unsigned gen = 0;
	bool broken = false;
};

static int host_barrier_wait(HostBarrier *b) {
	std::unique_lock<std::mutex> lk(b->m);
	if(b->broken) return 1;
	const unsigned g = b->gen;
	if(++b->waiting >= b->count) {
		b->waiting = 0;
		++b->gen;
		b->cv.notify_all();
		return 0;
	}
	const bool ok = b->cv.wait_for(lk, std::chrono::seconds(120), [&] { return b->gen != g || b->broken; });
	return (!ok || b->broken) ? 1 : 0;
}

static void host_barrier_break(HostBarrier *b) {
	std::lock_guard<std::mutex> lk(b->m);
	b->broken = true;
	b->cv.notify_all();
}

/* ---- row ownership (pure host arithmetic) ----
 * Matrix rows (sample slots) are owned in blocks of CCG_GROUP_ROW_BLOCK = 64 rows dealt round-robin: row i belongs
 * to member (i / 64) % world.  Every window of rows is thereby spread over all members (the NVLink reads of the
 * reduction run on every link at once) and a member's share of the cells is within a block row of 1 / world. */
extern "C" int ccg_group_row_block(void) { return CCG_GROUP_ROW_BLOCK; }

extern "C" int ccg_group_row_owner(int row, int world) {
	if(row < 0 || world < 1) return -1;
	return (row / CCG_GROUP_ROW_BLOCK) % world;
}

/* cells of the packed triangle over n samples (all included) that member `rank` owns */
extern "C" long long ccg_group_cells(int n, int rank, int world) {
	if(n < 0 || world < 1 || rank < 0 || rank >= world) return -1;
	long long cells = 0;
	for(int b = rank; (long long) b * CCG_GROUP_ROW_BLOCK < n; b += world) {
		const long long lo = (long long) b * CCG_GROUP_ROW_BLOCK;
		long long hi = lo + CCG_GROUP_ROW_BLOCK;
		if(hi > n) hi = n;
		cells += hi * (hi - 1) / 2 - lo * (lo - 1) / 2;
	}
	return cells;
}

/* ---- membership ---- */
void ccg_group_release(ccg_ctx *ctx) {
	if(!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	for(int p = 0; p < CCG_GROUP_MAX; ++p) {
		if(ctx->grp_opened[p] && ctx->grp_win[p]) cudaIpcCloseMemHandle(ctx->grp_win[p]);
		ctx->grp_opened[p] = 0;
		ctx->grp_win[p] = 0;
	}
	ctx->grp_world = 0;
	ctx->grp_rank = 0;
	cudaGetLastError();
}

extern "C" int ccg_group_leave(ccg_ctx *ctx) {
	if(!ctx) return CCG_ERR_ARG;
	ccg_group_release(ctx);
	return CCG_OK;
}

extern "C" int ccg_group_export(ccg_ctx *ctx, int max_samples, void *handle) {
	if(!ctx || !handle || max_samples < 1) return CCG_ERR_ARG;
	CKG(ctx, cudaSetDevice(ctx->device));
	ccg_group_release(ctx);
	const int npad = (max_samples + CCG_SLOT_PAD - 1) / CCG_SLOT_PAD * CCG_SLOT_PAD;
	/* two accumulator buffers of two int32 planes.  Up to 8 GiB they hold the whole n_pad x n_pad matrix; a larger
	 * problem is run in windows of whole macro-tile rows that fit (CCG_GROUP_WINDOW_BYTES: test hook) */
	size_t budget = (size_t) 8 << 30;
	if(getenv("CCG_GROUP_WINDOW_BYTES")) budget = (size_t) atoll(getenv("CCG_GROUP_WINDOW_BYTES"));
	const size_t one_tile_row = (size_t) 2 * 2 * CCG_UMMA_BM * npad * sizeof(int);
	size_t acc = (size_t) 2 * 2 * npad * npad * sizeof(int);
	if(acc > budget) acc = budget / one_tile_row * one_tile_row;
	if(acc < one_tile_row) acc = one_tile_row;
	const size_t bytes = CCG_GROUP_HDR_BYTES + acc;
	if(!ctx->grp_own_win || ctx->grp_win_bytes < bytes) {
		cudaFree(ctx->grp_own_win);
		ctx->grp_own_win = 0;
		ctx->grp_win_bytes = 0;
		if(cudaMalloc(&ctx->grp_own_win, bytes) != cudaSuccess) {
			ccg_set_err(ctx, "cudaMalloc of %zu bytes for the group window (accumulators of %d samples) failed: %s", bytes,
			            max_samples, cudaGetErrorString(cudaGetLastError()));
			return CCG_ERR_NOMEM;
		}
		ctx->grp_win_bytes = bytes;
		ctx->grp_npad_max = npad;
	}
	/* flags start at zero before any peer can know this window */
	CKG(ctx, cudaMemset(ctx->grp_own_win, 0, CCG_GROUP_HDR_BYTES));
	ctx->grp_epoch = 0;
	ctx->grp_buf = 0;
	GroupHandle h;
	memset(&h, 0, sizeof(h));
	h.magic = GROUP_MAGIC;
	h.pid = (int) getpid();
	h.device = ctx->device;
	h.npad_max = ctx->grp_npad_max;
	h.ptr = (unsigned long long) (uintptr_t) ctx->grp_own_win;
	h.bytes = ctx->grp_win_bytes;
	if(cudaIpcGetMemHandle(&h.ipc, ctx->grp_own_win) != cudaSuccess) {
		/* fine for members of one process (they use the pointer); a cross-process join will fail loudly */
		cudaGetLastError();
		memset(&h.ipc, 0, sizeof(h.ipc));
	}
	memset(handle, 0, CCG_GROUP_HANDLE_BYTES);
	memcpy(handle, &h, sizeof(h));
	return CCG_OK;
}

extern "C" int ccg_group_join(ccg_ctx *ctx, int rank, int world, const void *handles) {
	if(!ctx || !handles || world < 1 || world > CCG_GROUP_MAX || rank < 0 || rank >= world) return CCG_ERR_ARG;
	if(!ctx->grp_own_win) {
		ccg_set_err(ctx, "ccg_group_join before ccg_group_export");
		return CCG_ERR_ARG;
	}
	if(ctx->world > 1) {
		ccg_set_err(ctx, "a context is either a member of a K-split group or a rank of a tile partition, not both");
		return CCG_ERR_ARG;
	}
	CKG(ctx, cudaSetDevice(ctx->device));
	for(int p = 0; p < CCG_GROUP_MAX; ++p) {
		if(ctx->grp_opened[p] && ctx->grp_win[p]) cudaIpcCloseMemHandle(ctx->grp_win[p]);
		ctx->grp_opened[p] = 0;
		ctx->grp_win[p] = 0;
	}
	const int me = (int) getpid();
	ctx->grp_same_device = 0;
	int npad_min = ctx->grp_npad_max;
	size_t bytes_min = ctx->grp_win_bytes;
	for(int p = 0; p < world; ++p) {
		GroupHandle h;
		memcpy(&h, (const char *) handles + (size_t) p * CCG_GROUP_HANDLE_BYTES, sizeof(h));
		if(h.magic != GROUP_MAGIC) {
			ccg_set_err(ctx, "handle %d of the group is not a ccg_group_export handle", p);
			return CCG_ERR_ARG;
		}
		if(h.npad_max < npad_min) npad_min = h.npad_max;
		if(h.bytes < bytes_min) bytes_min = (size_t) h.bytes;
		if(p == rank) {
			if(h.pid != me || (void *) (uintptr_t) h.ptr != ctx->grp_own_win) {
				ccg_set_err(ctx, "handle %d is not this context's own export", p);
				return CCG_ERR_ARG;
			}
			ctx->grp_win[p] = ctx->grp_own_win;
		} else if(h.pid == me) {
			/* member of the same process: its pointer is valid here once peer access is on */
			if(h.device != ctx->device) {
				int can = 0;
				CKG(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
				if(!can) {
					ccg_set_err(ctx, "device %d cannot access device %d: no peer path for the K-split group", ctx->device, h.device);
					return CCG_ERR_UNSUPPORTED;
				}
				cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
				if(e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
					ccg_set_err(ctx, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(e));
					return CCG_ERR_CUDA;
				}
				cudaGetLastError();
			}
			ctx->grp_win[p] = (void *) (uintptr_t) h.ptr;
		} else {
			void *q = 0;
			cudaError_t e = cudaIpcOpenMemHandle(&q, h.ipc, cudaIpcMemLazyEnablePeerAccess);
			if(e != cudaSuccess) {
				ccg_set_err(ctx, "cudaIpcOpenMemHandle of member %d's window failed: %s", p, cudaGetErrorString(e));
				cudaGetLastError();
				return CCG_ERR_CUDA;
			}
			ctx->grp_win[p] = q;
			ctx->grp_opened[p] = 1;
		}
	}
	/* two members on one device (single-GPU tests): then EVERY member of the group synchronises on the host */
	for(int p = 0; p < world; ++p)
		for(int q = p + 1; q < world; ++q) {
			GroupHandle a, b;
			memcpy(&a, (const char *) handles + (size_t) p * CCG_GROUP_HANDLE_BYTES, sizeof(a));
			memcpy(&b, (const char *) handles + (size_t) q * CCG_GROUP_HANDLE_BYTES, sizeof(b));
			if(a.pid == b.pid && a.device == b.device) {
				if(a.pid != me) {
					ccg_set_err(ctx, "members %d and %d of the group share a device of another process", p, q);
					return CCG_ERR_UNSUPPORTED;
				}
				ctx->grp_same_device = 1;
			}
		}
	if(ctx->grp_same_device)
		for(int p = 0; p < world; ++p) {
			GroupHandle a;
			memcpy(&a, (const char *) handles + (size_t) p * CCG_GROUP_HANDLE_BYTES, sizeof(a));
			if(a.pid != me) {
				ccg_set_err(ctx, "a group with two members on one device must live in one process");
				return CCG_ERR_UNSUPPORTED;
			}
		}
	ctx->grp_npad_max = npad_min < ctx->grp_npad_max ? npad_min : ctx->grp_npad_max;
	ctx->grp_acc_bytes = bytes_min - CCG_GROUP_HDR_BYTES;      /* every member lays its buffers out the same way */
	ctx->grp_rank = rank;
	ctx->grp_world = world;
	ctx->grp_total_len = 0;
	ctx->grp_global_inc = 0;
	return CCG_OK;
}

extern "C" int ccg_group_set_alignment(ccg_ctx *ctx, long long total_len, unsigned global_inc) {
	if(!ctx || total_len < 0) return CCG_ERR_ARG;
	ctx->grp_total_len = total_len;
	ctx->grp_global_inc = global_inc;
	return CCG_OK;
}

extern "C" int ccg_group_set_output(ccg_ctx *ctx, int compact) {
	if(!ctx) return CCG_ERR_ARG;
	ctx->grp_compact = compact ? 1 : 0;
	return CCG_OK;
}

/* rows (a multiple of the macro-tile height) one accumulator buffer holds for the current problem */
int ccg_group_window_rows(ccg_ctx *ctx) {
	if(ctx->n_pad > ctx->grp_npad_max) return 0;
	const size_t per_row = (size_t) 2 * 2 * ctx->n_pad * sizeof(int);          /* 2 buffers x 2 planes */
	long long rows = (long long) (ctx->grp_acc_bytes / per_row) / CCG_UMMA_BM * CCG_UMMA_BM;
	if(rows > ctx->n_pad) rows = ctx->n_pad;
	return (int) rows;
}

/* The accumulator buffer of this run / window inside the member's own window, biased so that the kernels address
 * matrix row i as C[i * n_pad + j] for the rows [row0, row0 + window rows); *bytes = what to clear at *clear. */
int ccg_group_accumulators(ccg_ctx *ctx, int row0, int **C_S, int **C_I, void **clear, size_t *bytes) {
	const int rows = ccg_group_window_rows(ctx);
	if(rows < CCG_UMMA_BM) {
		ccg_set_err(ctx, "the group window (%zu accumulator bytes, exported for %d sample slots) cannot hold one macro-tile row of "
		            "%d slots", ctx->grp_acc_bytes, ctx->grp_npad_max, ctx->n_pad);
		return CCG_ERR_ARG;
	}
	const size_t plane = (size_t) rows * ctx->n_pad;
	int *base = (int *) ((char *) ctx->grp_own_win + CCG_GROUP_HDR_BYTES) + (size_t) ctx->grp_buf * 2 * plane;
	*C_S = base - (long long) row0 * ctx->n_pad;
	*C_I = *C_S + plane;
	*clear = base;
	*bytes = 2 * plane * sizeof(int);
	return CCG_OK;
}

/* after the member's GEMM of the rows [row0, row1): barrier with the peers, then reduce + epilogue of the rows of
 * that window this member owns */
int ccg_group_finalize(ccg_ctx *ctx, const EpilogueParams &ep, int i_const, int row0, int row1) {
	const int world = ctx->grp_world, rank = ctx->grp_rank;
	if(ctx->grp_same_device && !ctx->grp_host_barrier) {
		ccg_set_err(ctx, "members of a group that share a device run through ccg_init_multi_devices (they synchronise on the host)");
		return CCG_ERR_UNSUPPORTED;
	}
	if(ctx->grp_host_barrier && host_barrier_wait((HostBarrier *) ctx->grp_host_barrier)) {
		ccg_set_err(ctx, "another GPU of the group failed before the run reached the reduction");
		return CCG_ERR_CUDA;
	}
	GroupBarrierParams b;
	memset(&b, 0, sizeof(b));
	for(int p = 0; p < world; ++p) b.hdr[p] = (GroupHeader *) ctx->grp_win[p];
	b.rank = rank;
	b.world = world;
	b.buf = ctx->grp_buf;
	b.i_const = i_const;
	b.epoch = ++ctx->grp_epoch;
	b.spin = ctx->grp_same_device ? 0 : 1;
	b.timeout_ns = (unsigned long long) (getenv("CCG_GROUP_TIMEOUT_S") ? atof(getenv("CCG_GROUP_TIMEOUT_S")) : 120.0) * 1000000000ull;
	k_group_barrier<<<1, 32, 0, ctx->stream>>>(b);
	ctx->launches++;
	CKG(ctx, cudaGetLastError());
	/* Members that share ONE device (the single-GPU tests) must not wait for each other in spinning kernels -- nothing
	 * guarantees that two kernels of one device run at the same time: there the kernel above only publishes, every
	 * member drains its own stream, and the barrier is the host rendezvous below. */
	if(ctx->grp_same_device) CKG(ctx, cudaStreamSynchronize(ctx->stream));
	/* ... and once more after it: a host call that blocks until this stream has drained (a copy into pageable host
	 * memory, say) may hold the driver while it waits; by then every member's barrier kernel must be in its queue */
	if(ctx->grp_host_barrier && host_barrier_wait((HostBarrier *) ctx->grp_host_barrier)) {
		ccg_set_err(ctx, "another GPU of the group failed before the run reached the reduction");
		return CCG_ERR_CUDA;
	}
	if(getenv("CCG_DEBUG")) {
		cudaError_t e = cudaStreamSynchronize(ctx->stream);
		fprintf(stderr, "[ccg group] rank %d/%d epoch %u buf %d rows [%d,%d) n %d n_pad %d winrows %d acc_bytes %zu: barrier %s\n", rank, world,
		        b.epoch, b.buf, row0, row1, ctx->n, ctx->n_pad, ccg_group_window_rows(ctx), ctx->grp_acc_bytes, cudaGetErrorString(e));
	}

	const int rows_cap = ccg_group_window_rows(ctx);
	const size_t plane = (size_t) rows_cap * ctx->n_pad;
	GroupFinalizeParams q;
	memset(&q, 0, sizeof(q));
	for(int p = 0; p < world; ++p)
		q.C[p] = (const int *) ((const char *) ctx->grp_win[p] + CCG_GROUP_HDR_BYTES) + (size_t) ctx->grp_buf * 2 * plane -
		         (long long) row0 * ctx->n_pad;
	q.own = (const GroupHeader *) ctx->grp_win[rank];
	q.plane = plane;
	q.world = world;
	q.rank = rank;
	q.ldc = ctx->n_pad;
	q.n = ctx->n;
	q.pair_mode = ctx->pair_mode;            /* a three-plane store carries the inclusion counts in the I plane */
	q.buf = ctx->grp_buf;
	q.row_lo = row0;
	q.row_hi = row1 < ctx->n ? row1 : ctx->n;
	const int rows = q.row_hi - q.row_lo;
	if(rows > 0) {
		int grid = rows < 8 * ctx->sm_count ? rows : 8 * ctx->sm_count;
		k_finalize_group<<<grid, 256, 0, ctx->stream>>>(q, ep);
		ctx->launches++;
		CKG(ctx, cudaGetLastError());
		if(getenv("CCG_DEBUG")) {
			cudaError_t e = cudaStreamSynchronize(ctx->stream);
			fprintf(stderr, "[ccg group] rank %d finalize grid %d plane %zu row_base %p D %p: %s\n", rank, grid, q.plane, (const void *) ep.row_base,
			        ep.D, cudaGetErrorString(e));
		}
	}
	ctx->grp_buf ^= 1;
	return CCG_OK;
}

/* ====================================================================================================
 * In-process multi-GPU context (ccg_init_multi): one leader handle, one member context per device.
 * The leader takes the same calls as a single-device context; for a problem that is worth splitting
 * (tensor path, long alignment, no -P / -y) the members form a K-split group and every heavy call runs
 * on one host thread per member, otherwise member 0 works alone.  This is the fan-out the reference
 * does with pthreads inside fsaCmpThreadOut (fsacmpthrd.c:76-106), one level up.
 * ==================================================================================================== */
struct ccg_multi {
	int n;                                 /* members (devices) this handle may use */
	int created;                           /* member[0 .. created) exist: member 0 from the start, the others once a problem
	                                        * is split (a small job on an 8-GPU box starts one device context, not eight) */
	int device[CCG_GROUP_MAX];
	ccg_ctx *member[CCG_GROUP_MAX];
	int joined;                            /* members currently joined as a group of this many (0 = not joined) */
	int joined_samples;                    /* sample slots the members' windows were exported for */
	int active;                            /* members working on the current problem: 1 = member 0 alone */
	int samples, len, pair;
	int base0[CCG_GROUP_MAX + 1];          /* first base of every member's slice (multiples of 256) */
	int special;                           /* -P or -y set: those runs stay on one device */
	int kernel_choice;
	int force;                             /* CCG_MULTI_FORCE=1: split whatever the size (tests) */
	char last_kernel[160];
	HostBarrier *rendezvous;
	/* count matrices (.mat): the position axis is cut the same way */
	int mat_active, mat_n, mat_max_len;
	int mat_base0[CCG_GROUP_MAX + 1];
	int *mat_lens;                         /* whole length of every slot */
};

template <class F>
static int multi_parallel(ccg_multi *m, int count, F f) {
	std::vector<int> rcs((size_t) count, CCG_OK);
	std::vector<std::thread> th;
	auto run = [&](int g) {
		rcs[(size_t) g] = f(g);
		if(rcs[(size_t) g]) host_barrier_break(m->rendezvous);       /* nobody waits for a member that gave up */
	};
	for(int g = 1; g < count; ++g) th.emplace_back([&, g]() { run(g); });
	run(0);
	for(auto &t : th) t.join();
	m->rendezvous->broken = false;
	m->rendezvous->waiting = 0;
	for(int g = 0; g < count; ++g)
		if(rcs[(size_t) g]) return rcs[(size_t) g] | (g << 8);
	return CCG_OK;
}

static int multi_fail(ccg_ctx *lead, int packed) {
	const int g = packed >> 8, rc = packed & 0xFF;
	if(rc) ccg_set_err(lead, "gpu %d (device %d): %s", g, lead->multi->member[g]->device, lead->multi->member[g]->err);
	return rc;
}

static int multi_unsupported(ccg_ctx *lead, const char *what) {
	ccg_set_err(lead, "%s runs on one device: not available while the problem is split over %d GPUs", what, lead->multi->active);
	return CCG_ERR_UNSUPPORTED;
}

extern "C" int ccg_init_multi_devices(ccg_ctx **out, int ngpus, const int *devices) {
	if(!out || ngpus < 1 || ngpus > CCG_GROUP_MAX || !devices) return CCG_ERR_ARG;
	*out = 0;
	ccg_ctx *lead = 0;
	int rc = ccg_init(&lead, devices[0]);
	if(rc) return rc;
	ccg_multi *m = (ccg_multi *) calloc(1, sizeof(ccg_multi));
	if(!m) { ccg_destroy(lead); return CCG_ERR_NOMEM; }
	m->n = ngpus;
	m->active = 1;
	m->rendezvous = new HostBarrier();
	m->force = getenv("CCG_MULTI_FORCE") ? atoi(getenv("CCG_MULTI_FORCE")) : 0;
	for(int g = 0; g < ngpus; ++g) m->device[g] = devices[g];
	rc = ccg_init(&m->member[0], devices[0]);
	if(rc) {
		delete m->rendezvous;
		free(m);
		ccg_destroy(lead);
		return rc;
	}
	m->created = 1;
	lead->multi = m;
	*out = lead;
	return CCG_OK;
}

extern "C" int ccg_init_multi(ccg_ctx **out, int ngpus) {
	int count = 0;
	if(!out) return CCG_ERR_ARG;
	if(getenv("CCG_MULTI_DEVICES")) {
		/* test hook: an explicit member -> device list, ids may repeat ("0,0,0": three members on one GPU) */
		int devices[CCG_GROUP_MAX], k = 0;
		for(const char *p = getenv("CCG_MULTI_DEVICES"); *p && k < CCG_GROUP_MAX;) {
			devices[k++] = atoi(p);
			while(*p && *p != ',') ++p;
			if(*p == ',') ++p;
		}
		if(k > 0) return ccg_init_multi_devices(out, (ngpus > 0 && ngpus < k) ? ngpus : k, devices);
	}
	if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
		cudaGetLastError();
		ccg_set_err(0, "no CUDA device visible");
		return CCG_ERR_NO_DEVICE;
	}
	if(ngpus <= 0 || ngpus > count) ngpus = count;
	if(ngpus > CCG_GROUP_MAX) ngpus = CCG_GROUP_MAX;
	if(ngpus == 1) return ccg_init(out, 0);       /* one device: a plain context, nothing to fan out */
	int devices[CCG_GROUP_MAX];
	for(int g = 0; g < ngpus; ++g) devices[g] = g;
	return ccg_init_multi_devices(out, ngpus, devices);
}

extern "C" int ccg_multi_gpus(const ccg_ctx *ctx, int *active) {
	if(!ctx || !ctx->multi) { if(active) *active = 1; return 1; }
	if(active) *active = ctx->multi->active;
	return ctx->multi->n;
}

extern "C" int ccg_multi_contexts(const ccg_ctx *ctx) { return (ctx && ctx->multi) ? ctx->multi->created : 1; }

void ccg_multi_destroy(ccg_ctx *lead) {
	ccg_multi *m = lead->multi;
	if(!m) return;
	for(int g = 0; g < m->created; ++g) ccg_destroy(m->member[g]);
	delete m->rendezvous;
	free(m->mat_lens);
	free(m);
	lead->multi = 0;
}

/* creates the member contexts [created, count), one host thread per device (a device context takes a few hundred
 * milliseconds to start; eight in a row would be seconds) */
static int multi_ensure(ccg_ctx *lead, int count) {
	ccg_multi *m = lead->multi;
	if(count > m->n) count = m->n;
	if(count <= m->created) return CCG_OK;
	const int first = m->created;
	std::vector<int> rcs((size_t) count, CCG_OK);
	std::vector<std::thread> th;
	for(int g = first; g < count; ++g) th.emplace_back([&, g]() { rcs[(size_t) g] = ccg_init(&m->member[g], m->device[g]); });
	for(auto &t : th) t.join();
	int rc = CCG_OK;
	for(int g = first; g < count; ++g)
		if(rcs[(size_t) g] && !rc) {
			rc = rcs[(size_t) g];
			ccg_set_err(lead, "gpu %d (device %d): the device context could not be started (%s)", g, m->device[g], ccg_strerror(rc));
		}
	if(rc) {
		for(int g = first; g < count; ++g) {
			ccg_destroy(m->member[g]);
			m->member[g] = 0;
		}
		return rc;
	}
	for(int g = first; g < count; ++g) ccg_set_kernel(m->member[g], m->kernel_choice);
	m->created = count;
	return CCG_OK;
}

/* (re)joins the first `active` members as a group able to hold `samples` slots */
static int multi_join(ccg_ctx *lead, int active, int samples) {
	ccg_multi *m = lead->multi;
	int rc0 = multi_ensure(lead, active);
	if(rc0) return rc0;
	if(m->joined == active && m->joined_samples >= samples) return CCG_OK;
	for(int g = 0; g < m->created; ++g) {
		ccg_group_leave(m->member[g]);
		m->member[g]->grp_host_barrier = 0;
	}
	m->joined = 0;
	if(active < 2) return CCG_OK;
	std::vector<char> handles((size_t) active * CCG_GROUP_HANDLE_BYTES);
	for(int g = 0; g < active; ++g) {
		int rc = ccg_group_export(m->member[g], samples, handles.data() + (size_t) g * CCG_GROUP_HANDLE_BYTES);
		if(rc) return multi_fail(lead, rc | (g << 8));
	}
	for(int g = 0; g < active; ++g) {
		int rc = ccg_group_join(m->member[g], g, active, handles.data());
		if(rc) return multi_fail(lead, rc | (g << 8));
	}
	m->rendezvous->count = active;
	m->rendezvous->waiting = 0;
	for(int g = 0; g < active; ++g) m->member[g]->grp_host_barrier = m->rendezvous;
	m->joined = active;
	m->joined_samples = samples;
	return CCG_OK;
}

/* how many members a problem is split over: the tensor path on a long alignment, nothing that has to see a whole
 * sample at once (-P walks the alignment sequentially, a motif may straddle a slice boundary) */
static int multi_choose_active(const ccg_multi *m, int n, int len) {
	if(m->n < 2 || m->special || len < 512) return 1;
	if(m->kernel_choice == CCG_KERNEL_POPC || m->kernel_choice == CCG_KERNEL_FUSED) return 1;
	int a = m->n;
	if(m->force) {
		while(a > 1 && len / a < 256) --a;
		return a;
	}
	if(n < 512) return 1;
	while(a > 1 && len / a < 128 * 1024) --a;
	return a;
}

int ccg_multi_set_problem(ccg_ctx *lead, int n, int len, int pair_mode) {
	ccg_multi *m = lead->multi;
	const int active = multi_choose_active(m, n, len);
	int rc = multi_join(lead, active, n > 0 ? n : 1);
	if(rc) return rc;
	m->active = active;
	m->samples = n;
	m->len = len;
	m->pair = pair_mode ? 1 : 0;
	lead->n = n;
	lead->len = len;
	lead->pair_mode = m->pair;
	for(int g = 0; g <= active; ++g) m->base0[g] = g == active ? len : (int) ((long long) len * g / active / 256 * 256);
	for(int g = 0; g < active; ++g) {
		ccg_set_kernel(m->member[g], active > 1 ? CCG_KERNEL_UMMA : m->kernel_choice);
		rc = ccg_set_problem(m->member[g], n, m->base0[g + 1] - m->base0[g], pair_mode);
		if(rc) return multi_fail(lead, rc | (g << 8));
		if(active > 1) ccg_group_set_alignment(m->member[g], len, 0);
	}
	return CCG_OK;
}

int ccg_multi_set_kernel(ccg_ctx *lead, int kernel) {
	ccg_multi *m = lead->multi;
	m->kernel_choice = kernel;
	for(int g = 0; g < m->created; ++g) ccg_set_kernel(m->member[g], kernel);
	return CCG_OK;
}

int ccg_multi_sync(ccg_ctx *lead) {
	ccg_multi *m = lead->multi;
	for(int g = 0; g < m->created; ++g) {
		int rc = ccg_sync(m->member[g]);
		if(rc) return multi_fail(lead, rc | (g << 8));
	}
	return CCG_OK;
}

void ccg_multi_note_special(ccg_ctx *lead, int bit, int on) {
	if(on) lead->multi->special |= bit;
	else lead->multi->special &= ~bit;
}
ccg_ctx *ccg_multi_solo(ccg_ctx *lead, const char *what, int *rc) {
	ccg_multi *m = lead->multi;
	if(m->active > 1) { *rc = multi_unsupported(lead, what); return 0; }
	*rc = CCG_OK;
	return m->member[0];
}
/* copies member 0's message up after a forwarded call failed */
int ccg_multi_forwarded(ccg_ctx *lead, int rc) { return rc ? multi_fail(lead, rc) : CCG_OK; }

int ccg_multi_put_global_mask(ccg_ctx *lead, const uint32_t *mask, int apply) {
	ccg_multi *m = lead->multi;
	unsigned inc = 0;
	const int words = (m->len >> 5) + ((m->len & 31) ? 1 : 0);
	for(int w = 0; w < words; ++w) inc += (unsigned) __builtin_popcount(mask[w]);
	for(int g = 0; g < m->active; ++g) {
		const uint32_t *part = mask + m->base0[g] / 32;
		int rc = apply ? ccg_apply_global_mask(m->member[g], part) : ccg_put_global_mask(m->member[g], part);
		if(rc) return multi_fail(lead, rc | (g << 8));
		if(m->active > 1) ccg_group_set_alignment(m->member[g], m->len, inc);
	}
	lead->global_inc = inc;
	return CCG_OK;
}

int ccg_multi_build_global_mask(ccg_ctx *lead, const unsigned char *include, unsigned *global_inc) {
	ccg_multi *m = lead->multi;
	unsigned inc = 0;
	for(int g = 0; g < m->active; ++g) {
		unsigned part = 0;
		int rc = ccg_build_global_mask(m->member[g], include, &part);
		if(rc) return multi_fail(lead, rc | (g << 8));
		inc += part;
	}
	if(m->active > 1)
		for(int g = 0; g < m->active; ++g) ccg_group_set_alignment(m->member[g], m->len, inc);
	lead->global_inc = inc;
	if(global_inc) *global_inc = inc;
	return CCG_OK;
}

int ccg_multi_put_samples_packed(ccg_ctx *lead, int first, int count, const uint64_t *const *seqs, const uint32_t *const *includes) {
	ccg_multi *m = lead->multi;
	if(m->active == 1) return ccg_multi_forwarded(lead, ccg_put_samples_packed(m->member[0], first, count, seqs, includes));
	if(count <= 0) return CCG_OK;
	int rc = multi_parallel(m, m->active, [&](int g) -> int {
		const size_t w0 = (size_t) m->base0[g] / 32;
		std::vector<const uint64_t *> s((size_t) count);
		std::vector<const uint32_t *> k((size_t) count);
		for(int i = 0; i < count; ++i) {
			s[(size_t) i] = seqs[i] ? seqs[i] + w0 : 0;
			k[(size_t) i] = (includes && includes[i]) ? includes[i] + w0 : 0;
		}
		int r = ccg_put_samples_packed(m->member[g], first, count, s.data(), includes ? k.data() : 0);
		/* the row pointer arrays die here: the copies must have been issued AND staged */
		if(!r) r = ccg_sync(m->member[g]);
		return r;
	});
	return rc ? multi_fail(lead, rc) : CCG_OK;
}

int ccg_multi_put_sample_codes(ccg_ctx *lead, int idx, const unsigned char *codes) {
	ccg_multi *m = lead->multi;
	for(int g = 0; g < m->active; ++g) {
		int rc = ccg_put_sample_codes(m->member[g], idx, codes + m->base0[g]);
		if(rc) return multi_fail(lead, rc | (g << 8));
	}
	return CCG_OK;
}

int ccg_multi_get_inc_counts(ccg_ctx *lead, unsigned *out) {
	ccg_multi *m = lead->multi;
	const int n = m->samples;
	std::vector<unsigned> part((size_t) (n ? n : 1));
	for(int i = 0; i < n; ++i) out[i] = 0;
	for(int g = 0; g < m->active; ++g) {
		int rc = ccg_get_inc_counts(m->member[g], part.data());
		if(rc) return multi_fail(lead, rc | (g << 8));
		for(int i = 0; i < n; ++i) out[i] += part[(size_t) i];
	}
	return CCG_OK;
}

static void multi_note_kernel(ccg_ctx *lead) {
	ccg_multi *m = lead->multi;
	if(m->active > 1) snprintf(m->last_kernel, sizeof(m->last_kernel), "%s x %d gpus (K split)", ccg_last_kernel(m->member[0]), m->active);
	else snprintf(m->last_kernel, sizeof(m->last_kernel), "%s", ccg_last_kernel(m->member[0]));
}

int ccg_multi_run(ccg_ctx *lead, int pair, const unsigned char *include, unsigned norm, unsigned minLength, double minCov,
                  int elem_size, double byteScale, void *D, void *N, int *Dn, unsigned *global_inc) {
	ccg_multi *m = lead->multi;
	if(global_inc) *global_inc = lead->global_inc;
	std::vector<int> dn((size_t) m->active, 0);
	int rc = multi_parallel(m, m->active, [&](int g) -> int {
		unsigned gi = 0;
		return pair ? ccg_run_pair(m->member[g], include, norm, minLength, minCov, elem_size, byteScale, D, N, &dn[(size_t) g])
		            : ccg_run_global(m->member[g], include, norm, elem_size, byteScale, D, &dn[(size_t) g], &gi);
	});
	if(rc) return multi_fail(lead, rc);
	if(Dn) *Dn = dn[0];
	if(m->active == 1 && global_inc) *global_inc = m->member[0]->global_inc;
	multi_note_kernel(lead);
	return CCG_OK;
}

int ccg_multi_fsa_cmp_thread_out(ccg_ctx *lead, int pair, void *D, void *N, int elem_size, double byteScale, int n, int len,
                                 const uint64_t *const *seqs, const unsigned char *include, const uint32_t *const *includes,
                                 unsigned norm, unsigned minLength, double minCov, unsigned proxi, int *Dn, unsigned *global_inc) {
	ccg_multi *m = lead->multi;
	const int special = m->special;
	if(proxi && pair) m->special = 1;
	int rc = ccg_multi_set_problem(lead, n, len, pair);
	m->special = special;
	if(rc) return rc;
	if(m->active == 1) {
		rc = ccg_fsa_cmp_thread_out(m->member[0], pair, D, N, elem_size, byteScale, n, len, seqs, include, includes, norm, minLength,
		                            minCov, proxi, Dn, global_inc);
		multi_note_kernel(lead);
		return ccg_multi_forwarded(lead, rc);
	}
	unsigned ginc = 0;
	if(!pair) {
		const int words = (len >> 5) + ((len & 31) ? 1 : 0);
		for(int w = 0; w < words; ++w) ginc += (unsigned) __builtin_popcount(includes[0][w]);
	}
	std::vector<int> dn((size_t) m->active, 0);
	rc = multi_parallel(m, m->active, [&](int g) -> int {
		const size_t w0 = (size_t) m->base0[g] / 32;
		std::vector<const uint64_t *> s((size_t) (n ? n : 1));
		std::vector<const uint32_t *> k((size_t) (n ? n : 1));
		for(int i = 0; i < n; ++i) {
			s[(size_t) i] = seqs[i] ? seqs[i] + w0 : 0;
			k[(size_t) i] = pair ? (includes[i] ? includes[i] + w0 : 0) : includes[0] + w0;
		}
		if(!pair) k[0] = includes[0] + w0;
		ccg_group_set_alignment(m->member[g], len, ginc);
		unsigned gi = 0;
		return ccg_fsa_cmp_thread_out(m->member[g], pair, D, N, elem_size, byteScale, n, m->base0[g + 1] - m->base0[g], s.data(),
		                              include, k.data(), norm, minLength, minCov, 0, &dn[(size_t) g], &gi);
	});
	if(rc) return multi_fail(lead, rc);
	if(Dn) *Dn = dn[0];
	if(global_inc) *global_inc = ginc;
	lead->global_inc = ginc;
	multi_note_kernel(lead);
	return CCG_OK;
}

long long ccg_multi_launch_count(const ccg_ctx *lead) {
	long long k = 0;
	for(int g = 0; g < lead->multi->created; ++g) k += ccg_launch_count(lead->multi->member[g]);
	return k;
}
const char *ccg_multi_last_kernel(const ccg_ctx *lead) { return lead->multi->last_kernel; }
float ccg_multi_last_compare_ms(ccg_ctx *lead) {
	float ms = -1.0f;
	for(int g = 0; g < lead->multi->active; ++g) {
		const float v = ccg_last_compare_ms(lead->multi->member[g]);
		if(v > ms) ms = v;
	}
	return ms;
}
ccg_ctx *ccg_multi_member(ccg_ctx *lead, int g) {
	if(!lead->multi || g < 0 || g >= lead->multi->n || multi_ensure(lead, g + 1)) return 0;
	return lead->multi->member[g];
}

/* ---- count matrices on a multi-GPU context: member g holds the positions [mat_base0[g], mat_base0[g + 1]) of every
 * sample and returns its raw per-pair sums; the leader adds them in member order (a fixed order: the result does not
 * depend on timing) and finishes on the host with the reference's own arithmetic (ccg_mat_finalize_host) ---- */
extern "C" int ccg_mat_run_partial(ccg_ctx *ctx, const unsigned char *include, int method, unsigned order, double alpha,
                                   unsigned minDepth, double *dist, uint32_t *rows, int *Dn_out);
extern "C" int ccg_mat_finalize_host(int n, const unsigned char *include, const int *lens, const double *dist, const uint32_t *rows,
                                     unsigned norm, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N,
                                     uint32_t *rows_inc, int *Dn_out);

ccg_ctx *ccg_multi_mat_solo(ccg_ctx *lead) { return lead->multi->member[0]; }

int ccg_multi_mat_set_problem(ccg_ctx *lead, int n, int max_len) {
	ccg_multi *m = lead->multi;
	int a = m->n;
	if(m->force) { while(a > 1 && max_len / a < 64) --a; }
	else { if(n < 64) a = 1; while(a > 1 && max_len / a < 65536) --a; }
	m->mat_active = a;
	m->mat_n = n;
	m->mat_max_len = max_len;
	free(m->mat_lens);
	m->mat_lens = (int *) calloc((size_t) (n ? n : 1), sizeof(int));
	if(!m->mat_lens) return CCG_ERR_NOMEM;
	for(int g = 0; g <= a; ++g) m->mat_base0[g] = g == a ? max_len : (int) ((long long) max_len * g / a / 32 * 32);
	int rce = multi_ensure(lead, a);
	if(rce) return rce;
	for(int g = 0; g < m->created; ++g) {
		/* members outside the split give their store back */
		int rc = g < a ? ccg_mat_set_problem(m->member[g], n, m->mat_base0[g + 1] - m->mat_base0[g]) : ccg_mat_set_problem(m->member[g], 0, 0);
		if(rc) return multi_fail(lead, rc | (g << 8));
	}
	return CCG_OK;
}

int ccg_multi_mat_put_sample(ccg_ctx *lead, int idx, const uint16_t *counts6, const uint32_t *totals, int len) {
	ccg_multi *m = lead->multi;
	if(idx < 0 || idx >= m->mat_n || len < 0 || len > m->mat_max_len) return CCG_ERR_ARG;
	m->mat_lens[idx] = len;
	for(int g = 0; g < m->mat_active; ++g) {
		const int p0 = m->mat_base0[g], p1 = m->mat_base0[g + 1];
		const int part = len <= p0 ? 0 : (len < p1 ? len : p1) - p0;
		int rc = ccg_mat_put_sample(m->member[g], idx, counts6 + (size_t) p0 * 6, totals ? totals + p0 : 0, part);
		/* the member's pinned staging buffer is filled by this thread: the next put waits for the upload itself */
		if(rc) return multi_fail(lead, rc | (g << 8));
	}
	return CCG_OK;
}

int ccg_multi_mat_run(ccg_ctx *lead, const unsigned char *include, int method, unsigned order, double alpha, unsigned norm,
                      unsigned minDepth, unsigned minLength, double minCov, int elem_size, double byteScale, void *D, void *N, int *Dn_out,
                      uint32_t *rows_inc) {
	ccg_multi *m = lead->multi;
	if(!m->mat_lens) return CCG_ERR_ARG;
	if(m->mat_active == 1) {
		int rc = ccg_mat_run(m->member[0], include, method, order, alpha, norm, minDepth, minLength, minCov, elem_size, byteScale, D, N, Dn_out,
		                     rows_inc);
		multi_note_kernel(lead);
		return ccg_multi_forwarded(lead, rc);
	}
	int Dn = 0;
	for(int i = 0; i < m->mat_n; ++i) Dn += (!include || include[i]) ? 1 : 0;
	if(Dn_out) *Dn_out = Dn;
	if(Dn < 2) return CCG_OK;
	const size_t cells = (size_t) Dn * (Dn - 1) / 2;
	const int a = m->mat_active;
	std::vector<std::vector<double>> dist((size_t) a);
	std::vector<std::vector<uint32_t>> rows((size_t) a);
	int rc = multi_parallel(m, a, [&](int g) -> int {
		dist[(size_t) g].assign(cells, 0.0);
		rows[(size_t) g].assign(cells, 0u);
		int dn = 0;
		return ccg_mat_run_partial(m->member[g], include, method, order, alpha, minDepth, dist[(size_t) g].data(), rows[(size_t) g].data(), &dn);
	});
	if(rc) return multi_fail(lead, rc);
	for(int g = 1; g < a; ++g)
		for(size_t c = 0; c < cells; ++c) {
			dist[0][c] += dist[(size_t) g][c];
			rows[0][c] += rows[(size_t) g][c];
		}
	snprintf(m->last_kernel, sizeof(m->last_kernel), "%s x %d gpus (position split)", ccg_last_kernel(m->member[0]), a);
	return ccg_mat_finalize_host(m->mat_n, include, m->mat_lens, dist[0].data(), rows[0].data(), norm, minLength, minCov, elem_size, byteScale,
	                             D, N, rows_inc, 0);
}
