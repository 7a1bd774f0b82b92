#include "par.h"

#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <unistd.h>

int par_threads(void)
{
    const char *s = getenv("ARBPLF_HOST_THREADS");
    if (s && *s) {
        int v = atoi(s);
        if (v >= 1) return v > 256 ? 256 : v;
    }
    int n = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    if (n < 1) n = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    return n > 32 ? 32 : n;
}

typedef struct { par_fn fn; void *ctx; int tid, n; } par_job;

static void *par_entry(void *p)
{
    par_job *j = p;
    j->fn(j->tid, j->n, j->ctx);
    return NULL;
}

void par_run(int nthreads, par_fn fn, void *ctx)
{
    if (nthreads <= 1) { fn(0, 1, ctx); return; }
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)nthreads);
    par_job *jobs = malloc(sizeof(par_job) * (size_t)nthreads);
    char *started = calloc((size_t)nthreads, 1);
    if (!th || !jobs || !started) {
        free(th); free(jobs); free(started);
        for (int t = 0; t < nthreads; t++) fn(t, nthreads, ctx);
        return;
    }
    for (int t = 1; t < nthreads; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].tid = t; jobs[t].n = nthreads;
        started[t] = pthread_create(&th[t], NULL, par_entry, &jobs[t]) == 0;
    }
    fn(0, nthreads, ctx);
    for (int t = 1; t < nthreads; t++) {
        if (started[t]) pthread_join(th[t], NULL);
        else fn(t, nthreads, ctx);
    }
    free(th); free(jobs); free(started);
}
