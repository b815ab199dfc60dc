/*
 * ordered_pool.h -- run jobs 0..njobs-1 on worker threads, hand the results to the caller in
 * job order, with at most `window` jobs parsed ahead of the consumer (bounds host memory to
 * O(window) samples however many input files there are).  This is what replaces the
 * reference's serial per-sample load loops (cdist.c:55-168, ltdmatrixthrd.c:469-538).
 */
#ifndef CCB_ORDERED_POOL_H
#define CCB_ORDERED_POOL_H

typedef struct OrderedPool OrderedPool;
/* work(job, slot_state, user): fills the slot's state (slot_state = states + slot_size * (job % window)) */
typedef void (*PoolWork)(int job, void *slot_state, void *user);

OrderedPool *pool_start(int njobs, int nthreads, int window, void *states, unsigned long slot_size, PoolWork work, void *user);
/* blocks until job (taken in increasing order) is done; returns its slot state */
void *pool_take(OrderedPool *p, int job);
/* the consumer is finished with the job's slot */
void pool_release(OrderedPool *p, int job);
void pool_finish(OrderedPool *p);

#endif
