/* ordered_pool_test.c -- the driver's parse pool (ccphylo_b200/host/ordered_pool.c): results arrive in job order
 * whatever the workers' timing, no slot is reused before it is released, the window bounds the look-ahead.
 *
 *   ordered_pool_test <njobs> <nthreads> <window> */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "ordered_pool.h"

typedef struct {
	int job, busy;
	unsigned long value;
} Slot;

static volatile int consumed = 0;       /* jobs the consumer has released */
static int window_g = 0;
static volatile int violation = 0;

static void work(int job, void *state, void *user) {
	Slot *s = (Slot *) state;
	if(s->busy) violation = 1;                              /* slot handed out while still in use */
	if(job >= consumed + window_g) violation = 2;           /* parsed further ahead than the window allows */
	s->busy = 1;
	s->job = job;
	/* uneven work so that later jobs finish before earlier ones */
	struct timespec ts = {0, (long) ((job * 7919u) % 13u) * 100000L};
	nanosleep(&ts, 0);
	s->value = (unsigned long) job * 2654435761ul;
}

int main(int argc, char **argv) {
	const int njobs = argc > 1 ? atoi(argv[1]) : 200, nthreads = argc > 2 ? atoi(argv[2]) : 8, window = argc > 3 ? atoi(argv[3]) : 10;
	window_g = window;
	Slot *slots = calloc((size_t) window, sizeof(Slot));
	OrderedPool *p = pool_start(njobs, nthreads, window, slots, sizeof(Slot), work, 0);
	if(!p) return 1;
	for(int j = 0; j < njobs; ++j) {
		Slot *s = (Slot *) pool_take(p, j);
		if(s != &slots[j % window] || s->job != j || s->value != (unsigned long) j * 2654435761ul) {
			printf("job %d: wrong slot or content (job %d)\n", j, s->job);
			return 1;
		}
		s->busy = 0;
		consumed = j + 1;
		pool_release(p, j);
	}
	pool_finish(p);
	if(violation) {
		printf("violation %d\n", violation);
		return 1;
	}
	printf("OK\n");
	free(slots);
	return 0;
}
