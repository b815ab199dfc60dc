/* ordered_pool.c -- see ordered_pool.h */
#include "ordered_pool.h"

#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct OrderedPool {
	int njobs, nthreads, window;
	char *states;
	unsigned long slot_size;
	PoolWork work;
	void *user;
	unsigned char *ready;      /* per window slot */
	int next_job, consumed;
	pthread_mutex_t mu;
	pthread_cond_t cv_ready, cv_room;
	pthread_t *th;
};

static void *worker(void *arg) {
	OrderedPool *p = (OrderedPool *) arg;
	for(;;) {
		pthread_mutex_lock(&p->mu);
		while(p->next_job < p->njobs && p->next_job >= p->consumed + p->window) pthread_cond_wait(&p->cv_room, &p->mu);
		if(p->next_job >= p->njobs) {
			pthread_mutex_unlock(&p->mu);
			return 0;
		}
		const int job = p->next_job++;
		pthread_mutex_unlock(&p->mu);
		p->work(job, p->states + p->slot_size * (unsigned long) (job % p->window), p->user);
		pthread_mutex_lock(&p->mu);
		p->ready[job % p->window] = 1;
		pthread_cond_broadcast(&p->cv_ready);
		pthread_mutex_unlock(&p->mu);
	}
}

OrderedPool *pool_start(int njobs, int nthreads, int window, void *states, unsigned long slot_size, PoolWork work, void *user) {
	OrderedPool *p = calloc(1, sizeof(*p));
	if(!p) return 0;
	if(nthreads < 1) nthreads = 1;
	if(nthreads > njobs) nthreads = njobs > 0 ? njobs : 1;
	if(window < 1) window = 1;
	if(nthreads > window) nthreads = window;      /* the caller's state array has `window` slots: never index past it */
	p->njobs = njobs;
	p->nthreads = nthreads;
	p->window = window;
	p->states = states;
	p->slot_size = slot_size;
	p->work = work;
	p->user = user;
	p->ready = calloc((size_t) window, 1);
	p->th = calloc((size_t) nthreads, sizeof(pthread_t));
	if(!p->ready || !p->th) {
		fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
		exit(errno ? errno : 1);
	}
	pthread_mutex_init(&p->mu, 0);
	pthread_cond_init(&p->cv_ready, 0);
	pthread_cond_init(&p->cv_room, 0);
	for(int k = 0; k < nthreads; ++k) {
		if((errno = pthread_create(&p->th[k], 0, worker, p))) {
			fprintf(stderr, "Error: %d (%s)\n", errno, strerror(errno));
			if(k == 0) exit(errno);
			fprintf(stderr, "Will continue with %d threads.\n", k);
			p->nthreads = k;
			break;
		}
	}
	return p;
}

void *pool_take(OrderedPool *p, int job) {
	pthread_mutex_lock(&p->mu);
	while(!p->ready[job % p->window]) pthread_cond_wait(&p->cv_ready, &p->mu);
	pthread_mutex_unlock(&p->mu);
	return p->states + p->slot_size * (unsigned long) (job % p->window);
}

void pool_release(OrderedPool *p, int job) {
	pthread_mutex_lock(&p->mu);
	p->ready[job % p->window] = 0;
	p->consumed = job + 1;
	pthread_cond_broadcast(&p->cv_room);
	pthread_mutex_unlock(&p->mu);
}

void pool_finish(OrderedPool *p) {
	for(int k = 0; k < p->nthreads; ++k) pthread_join(p->th[k], 0);
	pthread_mutex_destroy(&p->mu);
	pthread_cond_destroy(&p->cv_ready);
	pthread_cond_destroy(&p->cv_room);
	free(p->ready);
	free(p->th);
	free(p);
}
