/*
 * Host threads for the two passes of the front end that scale with the alignment: reading the 'character_data'
 * matrix of a document (json.c) and turning it into code bytes (model.c).  The reference does both on one thread
 * inside jansson / parsemodel.c:514-628; at 10^6 site patterns that is most of the wall time of an arbplf-* call
 * once the likelihood itself takes milliseconds.
 */
#ifndef PLF_PAR_H
#define PLF_PAR_H

typedef void (*par_fn)(int tid, int nthreads, void *ctx);

/* threads worth starting: the CPUs this process may run on, at most 32; ARBPLF_HOST_THREADS overrides */
int par_threads(void);

/* runs fn(tid, nthreads, ctx) for tid = 0 .. nthreads-1 and waits; tid 0 runs on the calling thread.
 * Threads that cannot be started have their share run by the caller, so every tid is always served. */
void par_run(int nthreads, par_fn fn, void *ctx);

#endif
