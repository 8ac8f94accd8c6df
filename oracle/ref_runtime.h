/*
 * ref_runtime.h -- run-time support for the machine-translated reference (oracle/f90toc.py).
 *
 * TEST INFRASTRUCTURE ONLY (see d3q19_oracle.h).  This file is ours; the code that includes it
 * (oracle/_ref/ref_translated.c) is generated from the reference's Fortran sources at build
 * time and is never committed.  Provided here:
 *   - Fortran-style array descriptors with arbitrary lower bounds, column-major (ref_arr);
 *   - by-name overrides of the values the reference hard-codes (grid size, flow type, ...);
 *   - an in-process mini-MPI: each MPI rank of the reference is one thread, point-to-point
 *     messages go through a mailbox matched on (source, destination, tag) exactly like
 *     MPI_ISEND / MPI_IRECV / MPI_WAITALL (collision.f90:309-314,351-356), collectives
 *     (MPI_ALLGATHER para.f90:251-252, MPI_ALLREDUCE collision.f90:500-501, MPI_BARRIER)
 *     reduce in rank order;
 *   - a world object that runs one translated subroutine on every rank concurrently.
 */
#ifndef REF_RUNTIME_H
#define REF_RUNTIME_H

#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define REF_KIND_R 8
#define REF_KIND_I 4

enum { REF_MPI_REAL8 = 1, REF_MPI_INTEGER = 2, REF_MPI_BYTE = 3 };
enum { REF_MPI_SUM = 1, REF_MPI_MAX = 2, REF_MPI_MIN = 3 };

/* ---- arrays ------------------------------------------------------------------------------- */
typedef struct ref_arr {
    void *p;
    int kind, rank, owned;
    int lo[4], n[4];
} ref_arr;

#define REF_IDX1(A, i) ((size_t)((i) - (A).lo[0]))
#define REF_IDX2(A, i, j) (REF_IDX1(A, i) + (size_t)(A).n[0] * (size_t)((j) - (A).lo[1]))
#define REF_IDX3(A, i, j, k) (REF_IDX2(A, i, j) + (size_t)(A).n[0] * (size_t)(A).n[1] * (size_t)((k) - (A).lo[2]))
#define REF_IDX4(A, i, j, k, l) \
    (REF_IDX3(A, i, j, k) + (size_t)(A).n[0] * (size_t)(A).n[1] * (size_t)(A).n[2] * (size_t)((l) - (A).lo[3]))
#define REF_R1(A, i) (((double *)(A).p)[REF_IDX1(A, i)])
#define REF_R2(A, i, j) (((double *)(A).p)[REF_IDX2(A, i, j)])
#define REF_R3(A, i, j, k) (((double *)(A).p)[REF_IDX3(A, i, j, k)])
#define REF_R4(A, i, j, k, l) (((double *)(A).p)[REF_IDX4(A, i, j, k, l)])
#define REF_I1(A, i) (((int *)(A).p)[REF_IDX1(A, i)])
#define REF_I2(A, i, j) (((int *)(A).p)[REF_IDX2(A, i, j)])
#define REF_I3(A, i, j, k) (((int *)(A).p)[REF_IDX3(A, i, j, k)])
#define REF_I4(A, i, j, k, l) (((int *)(A).p)[REF_IDX4(A, i, j, k, l)])

static inline ref_arr ref_view(void *p, int kind, int rank, const int *lo, const int *hi)
{
    ref_arr a;
    memset(&a, 0, sizeof a);
    a.p = p; a.kind = kind; a.rank = rank; a.owned = 0;
    for (int d = 0; d < rank; ++d) { a.lo[d] = lo[d]; a.n[d] = hi[d] - lo[d] + 1; if (a.n[d] < 0) a.n[d] = 0; }
    return a;
}

static inline size_t ref_count(const ref_arr *a)
{
    size_t c = 1;
    for (int d = 0; d < a->rank; ++d) c *= (size_t)a->n[d];
    return c;
}

static inline ref_arr ref_alloc(int kind, int rank, const int *lo, const int *hi)
{
    ref_arr a = ref_view(0, kind, rank, lo, hi);
    size_t c = ref_count(&a);
    a.p = calloc(c ? c : 1, (size_t)kind);
    a.owned = 1;
    return a;
}

/* automatic (stack) arrays of a subroutine: Fortran leaves them uninitialised; the parity build
 * zeroes them for determinism, the timing build (-DREF_AUTO_MALLOC) does not pay for that */
static inline ref_arr ref_alloc_auto(int kind, int rank, const int *lo, const int *hi)
{
#ifdef REF_AUTO_MALLOC
    ref_arr a = ref_view(0, kind, rank, lo, hi);
    size_t c = ref_count(&a);
    a.p = malloc((c ? c : 1) * (size_t)kind);
    a.owned = 1;
    return a;
#else
    return ref_alloc(kind, rank, lo, hi);
#endif
}

static inline void ref_free(ref_arr *a)
{
    if (a->owned && a->p) free(a->p);
    a->p = 0; a->owned = 0;
}

static inline double ref_powi_d(double x, int n)
{
    /* x**n by repeated multiplication, left to right (what ifort/gfortran emit for small n) */
    if (n == 0) return 1.0;
    int m = n < 0 ? -n : n;
    double r = x;
    for (int k = 1; k < m; ++k) r = r * x;
    return n < 0 ? 1.0 / r : r;
}
static inline int ref_powi_i(int x, int n)
{
    int r = 1;
    for (int k = 0; k < n; ++k) r *= x;
    return r;
}
static inline double ref_max_d(double a, double b) { return a > b ? a : b; }
static inline double ref_min_d(double a, double b) { return a < b ? a : b; }
static inline int ref_max_i(int a, int b) { return a > b ? a : b; }
static inline int ref_min_i(int a, int b) { return a < b ? a : b; }

/* ---- mini-MPI ------------------------------------------------------------------------------- */
typedef struct ref_msg {
    int src, dst, tag;
    size_t bytes;
    void *data;
    struct ref_msg *next;
} ref_msg;

typedef struct ref_comm {
    int nproc;
    pthread_mutex_t mu;
    pthread_cond_t cv;
    ref_msg *head, *tail;
    int bar_count, bar_gen;
    void **slot;              /* one pointer per rank for collectives */
} ref_comm;

typedef struct ref_pending {
    void *buf;
    size_t bytes;
    int src, tag;
} ref_pending;

#define REF_MAX_OVERRIDES 64
typedef struct ref_common {
    ref_comm *comm;
    int rank;
    int nov;
    char ov_name[REF_MAX_OVERRIDES][32];
    double ov_val[REF_MAX_OVERRIDES];
    int npend;
    ref_pending pend[16];
    /* values the translated code handed to write(unit, fmt) statements (file output of statistc, diag, ...):
     * unit numbers and values in the order written; read back by the tests through ref_capture_get */
    int ncap, capcap;
    int *cap_unit;
    double *cap_val;
    /* values the translated code will be handed by read(unit) statements (loadcntdflow), in the order read */
    long nplay, iplay;
    double *play;
    long wtime_calls;
} ref_common;

static inline double ref_play(void *S, int unit)
{
    ref_common *c = (ref_common *)S;
    if (c->iplay >= c->nplay) {
        fprintf(stderr, "ref: rank %d reads past the end of the playback queue (unit %d)\n", c->rank, unit);
        abort();
    }
    return c->play[c->iplay++];
}

static inline void ref_capture(void *S, int unit, double v)
{
    ref_common *c = (ref_common *)S;
    if (c->ncap == c->capcap) {
        c->capcap = c->capcap ? 2 * c->capcap : 256;
        c->cap_unit = (int *)realloc(c->cap_unit, (size_t)c->capcap * sizeof(int));
        c->cap_val = (double *)realloc(c->cap_val, (size_t)c->capcap * sizeof(double));
    }
    c->cap_unit[c->ncap] = unit;
    c->cap_val[c->ncap++] = v;
}

struct ref_state;
typedef struct ref_state ref_state;
size_t ref_state_size(void);
int ref_dispatch(ref_state *S, const char *name);
void ref_module_free(ref_state *S);

static inline size_t ref_type_size(int t) { return t == REF_MPI_REAL8 ? 8 : (t == REF_MPI_INTEGER ? 4 : 1); }

static inline double ref_override_d(void *S, const char *name, double v)
{
    ref_common *c = (ref_common *)S;
    for (int i = 0; i < c->nov; ++i)
        if (!strcmp(c->ov_name[i], name)) return c->ov_val[i];
    return v;
}
static inline int ref_override_i(void *S, const char *name, int v)
{
    ref_common *c = (ref_common *)S;
    for (int i = 0; i < c->nov; ++i)
        if (!strcmp(c->ov_name[i], name)) return (int)c->ov_val[i];
    return v;
}

static inline void ref_stop(void *S)
{
    fprintf(stderr, "ref: the reference executed STOP on rank %d\n", ((ref_common *)S)->rank);
    abort();
}

static inline void ref_barrier_(ref_comm *cm)
{
    pthread_mutex_lock(&cm->mu);
    int gen = cm->bar_gen;
    if (++cm->bar_count == cm->nproc) {
        cm->bar_count = 0;
        cm->bar_gen++;
        pthread_cond_broadcast(&cm->cv);
    } else {
        while (gen == cm->bar_gen) pthread_cond_wait(&cm->cv, &cm->mu);
    }
    pthread_mutex_unlock(&cm->mu);
}

static inline void ref_mpi_barrier(void *S, int comm, int *ierr)
{
    (void)comm;
    ref_barrier_(((ref_common *)S)->comm);
    *ierr = 0;
}

static inline void ref_mpi_isend(void *S, void *buf, int *count, int type, int *dst, int *tag, int comm, int *req, int *ierr)
{
    (void)comm; (void)req;
    ref_common *c = (ref_common *)S;
    ref_msg *m = (ref_msg *)malloc(sizeof *m);
    m->src = c->rank; m->dst = *dst; m->tag = *tag;
    m->bytes = (size_t)*count * ref_type_size(type);
    m->data = malloc(m->bytes ? m->bytes : 1);
    memcpy(m->data, buf, m->bytes);
    m->next = 0;
    pthread_mutex_lock(&c->comm->mu);
    if (c->comm->tail) c->comm->tail->next = m; else c->comm->head = m;
    c->comm->tail = m;
    pthread_cond_broadcast(&c->comm->cv);
    pthread_mutex_unlock(&c->comm->mu);
    *ierr = 0;
}

static inline void ref_mpi_irecv(void *S, void *buf, int *count, int type, int *src, int *tag, int comm, int *req, int *ierr)
{
    (void)comm; (void)req;
    ref_common *c = (ref_common *)S;
    if (c->npend >= 16) { fprintf(stderr, "ref: too many pending receives\n"); abort(); }
    ref_pending *p = &c->pend[c->npend++];
    p->buf = buf; p->bytes = (size_t)*count * ref_type_size(type); p->src = *src; p->tag = *tag;
    *ierr = 0;
}

static inline void ref_mpi_waitall(void *S, int *n, void *req, void *status, int *ierr)
{
    (void)n; (void)req; (void)status;
    ref_common *c = (ref_common *)S;
    ref_comm *cm = c->comm;
    pthread_mutex_lock(&cm->mu);
    for (int i = 0; i < c->npend; ++i) {
        ref_pending *p = &c->pend[i];
        for (;;) {
            ref_msg *prev = 0, *m = cm->head;
            while (m && !(m->dst == c->rank && m->src == p->src && m->tag == p->tag)) { prev = m; m = m->next; }
            if (m) {
                if (m->bytes != p->bytes) { fprintf(stderr, "ref: message size mismatch\n"); abort(); }
                memcpy(p->buf, m->data, m->bytes);
                if (prev) prev->next = m->next; else cm->head = m->next;
                if (cm->tail == m) cm->tail = prev;
                free(m->data); free(m);
                break;
            }
            pthread_cond_wait(&cm->cv, &cm->mu);
        }
    }
    c->npend = 0;
    pthread_mutex_unlock(&cm->mu);
    *ierr = 0;
}

/* what PROGRAM main (main.f90:26-28,233) and the drop-in's Fortran shim use on top of the hot path's calls */
static inline void ref_mpi_init(void *S, int *ierr) { (void)S; *ierr = 0; }
static inline void ref_mpi_finalize(void *S, int *ierr) { (void)S; *ierr = 0; }
static inline void ref_mpi_comm_rank(void *S, int comm, int *rank, int *ierr) { (void)comm; *rank = ((ref_common *)S)->rank; *ierr = 0; }
static inline void ref_mpi_comm_size(void *S, int comm, int *n, int *ierr) { (void)comm; *n = ((ref_common *)S)->comm->nproc; *ierr = 0; }
static inline void ref_mpi_abort(void *S, int comm, int *code, int *ierr)
{
    (void)comm; (void)ierr;
    fprintf(stderr, "ref: MPI_ABORT(%d) on rank %d\n", *code, ((ref_common *)S)->rank);
    abort();
}
/* MPI_WTIME: the number of calls so far times the override "wtime_tick" (default 0: the wall-clock exit of
 * main.f90:197-207 never fires; a test sets a tick and `time_bond` to make it fire at a chosen step) */
static inline double ref_mpi_wtime(void *S)
{
    ref_common *c = (ref_common *)S;
    c->wtime_calls += 1;
    return (double)c->wtime_calls * ref_override_d(S, "wtime_tick", 0.0);
}
static inline void ref_untranslated(void *S, const char *name)
{
    fprintf(stderr, "ref: rank %d reached `call %s`, which is outside the translated set\n", ((ref_common *)S)->rank, name);
    abort();
}
static inline void ref_mpi_bcast(void *S, void *buf, int *count, int type, int *root, int comm, int *ierr)
{
    (void)comm;
    ref_common *c = (ref_common *)S;
    ref_comm *cm = c->comm;
    cm->slot[c->rank] = buf;
    ref_barrier_(cm);
    if (c->rank != *root) memcpy(buf, cm->slot[*root], (size_t)*count * ref_type_size(type));
    ref_barrier_(cm);
    *ierr = 0;
}

/* blocking point-to-point (outputuy / outputpress, saveload.f90:871,889) on top of the non-blocking pair */
static inline void ref_mpi_send(void *S, void *buf, int *count, int type, int *dst, int *tag, int comm, int *ierr)
{
    int req = 0;
    ref_mpi_isend(S, buf, count, type, dst, tag, comm, &req, ierr);
}
static inline void ref_mpi_recv(void *S, void *buf, int *count, int type, int *src, int *tag, int comm, void *status, int *ierr)
{
    int req = 0, one = 1;
    ref_mpi_irecv(S, buf, count, type, src, tag, comm, &req, ierr);
    ref_mpi_waitall(S, &one, &req, status, ierr);
}
static inline void ref_mpi_gather(void *S, void *sbuf, int *scount, int stype, void *rbuf, int *rcount, int rtype,
                                  int *root, int comm, int *ierr)
{
    (void)comm; (void)rcount; (void)rtype;
    ref_common *c = (ref_common *)S;
    ref_comm *cm = c->comm;
    size_t b = (size_t)*scount * ref_type_size(stype);
    cm->slot[c->rank] = sbuf;
    ref_barrier_(cm);
    if (c->rank == *root)
        for (int r = 0; r < cm->nproc; ++r) memcpy((char *)rbuf + (size_t)r * b, cm->slot[r], b);
    ref_barrier_(cm);
    *ierr = 0;
}

static inline void ref_mpi_allgather(void *S, void *sbuf, int *scount, int stype, void *rbuf, int *rcount, int rtype,
                                     int comm, int *ierr)
{
    (void)comm; (void)rcount; (void)rtype;
    ref_common *c = (ref_common *)S;
    ref_comm *cm = c->comm;
    size_t b = (size_t)*scount * ref_type_size(stype);
    cm->slot[c->rank] = sbuf;
    ref_barrier_(cm);
    for (int r = 0; r < cm->nproc; ++r) memcpy((char *)rbuf + (size_t)r * b, cm->slot[r], b);
    ref_barrier_(cm);
    *ierr = 0;
}

static inline void ref_mpi_allreduce(void *S, void *sbuf, void *rbuf, int *count, int type, int op, int comm, int *ierr)
{
    (void)comm;
    ref_common *c = (ref_common *)S;
    ref_comm *cm = c->comm;
    cm->slot[c->rank] = sbuf;
    ref_barrier_(cm);
    for (int k = 0; k < *count; ++k) {
        if (type == REF_MPI_REAL8) {
            double acc = ((double *)cm->slot[0])[k];
            for (int r = 1; r < cm->nproc; ++r) {
                double v = ((double *)cm->slot[r])[k];
                acc = op == REF_MPI_SUM ? acc + v : (op == REF_MPI_MAX ? (v > acc ? v : acc) : (v < acc ? v : acc));
            }
            ((double *)rbuf)[k] = acc;
        } else {
            int acc = ((int *)cm->slot[0])[k];
            for (int r = 1; r < cm->nproc; ++r) {
                int v = ((int *)cm->slot[r])[k];
                acc = op == REF_MPI_SUM ? acc + v : (op == REF_MPI_MAX ? (v > acc ? v : acc) : (v < acc ? v : acc));
            }
            ((int *)rbuf)[k] = acc;
        }
    }
    ref_barrier_(cm);
    *ierr = 0;
}

/* ---- the world: all ranks of one run ---------------------------------------------------------- */
typedef struct ref_world {
    int nproc;
    ref_comm comm;
    ref_state **st;
} ref_world;

static ref_world *ref_world_create_(int nproc)
{
    ref_world *w = (ref_world *)calloc(1, sizeof *w);
    w->nproc = nproc;
    w->comm.nproc = nproc;
    pthread_mutex_init(&w->comm.mu, 0);
    pthread_cond_init(&w->comm.cv, 0);
    w->comm.slot = (void **)calloc((size_t)nproc, sizeof(void *));
    w->st = (ref_state **)calloc((size_t)nproc, sizeof(ref_state *));
    for (int r = 0; r < nproc; ++r) {
        w->st[r] = (ref_state *)calloc(1, ref_state_size());
        ref_common *c = (ref_common *)w->st[r];
        c->comm = &w->comm;
        c->rank = r;
    }
    return w;
}

typedef struct ref_job {
    ref_state *S;
    const char *name;
    int rc;
} ref_job;

static void *ref_job_main(void *arg)
{
    ref_job *j = (ref_job *)arg;
    j->rc = ref_dispatch(j->S, j->name);
    return 0;
}

/* exported entry points (the Python wrapper oracle/ref.py binds these) */
ref_world *ref_world_create(int nproc) { return ref_world_create_(nproc); }

void ref_world_destroy(ref_world *w)
{
    if (!w) return;
    for (int r = 0; r < w->nproc; ++r) {
        ref_common *c = (ref_common *)w->st[r];
        free(c->cap_unit); free(c->cap_val); free(c->play);
        ref_module_free(w->st[r]); free(w->st[r]);
    }
    while (w->comm.head) { ref_msg *m = w->comm.head; w->comm.head = m->next; free(m->data); free(m); }
    free(w->st); free(w->comm.slot);
    pthread_mutex_destroy(&w->comm.mu);
    pthread_cond_destroy(&w->comm.cv);
    free(w);
}

ref_state *ref_world_state(ref_world *w, int rank) { return w->st[rank]; }

/* captured write() values of one rank: returns the count, copies up to `max` (unit, value) pairs */
int ref_capture_get(ref_world *w, int rank, int max, int *units, double *vals)
{
    ref_common *c = (ref_common *)w->st[rank];
    for (int i = 0; i < c->ncap && i < max; ++i) { units[i] = c->cap_unit[i]; vals[i] = c->cap_val[i]; }
    return c->ncap;
}
void ref_capture_clear(ref_world *w)
{
    for (int r = 0; r < w->nproc; ++r) ((ref_common *)w->st[r])->ncap = 0;
}

/* what the next read(unit) statements of `rank` will be handed */
void ref_world_set_playback(ref_world *w, int rank, long n, const double *vals)
{
    ref_common *c = (ref_common *)w->st[rank];
    free(c->play);
    c->play = (double *)malloc((size_t)(n ? n : 1) * sizeof(double));
    memcpy(c->play, vals, (size_t)n * sizeof(double));
    c->nplay = n; c->iplay = 0;
}
long ref_world_playback_left(ref_world *w, int rank)
{
    ref_common *c = (ref_common *)w->st[rank];
    return c->nplay - c->iplay;
}

int ref_world_set_override(ref_world *w, const char *name, double v)
{
    for (int r = 0; r < w->nproc; ++r) {
        ref_common *c = (ref_common *)w->st[r];
        int i = 0;
        while (i < c->nov && strcmp(c->ov_name[i], name)) ++i;
        if (i == c->nov) {
            if (c->nov >= REF_MAX_OVERRIDES || strlen(name) > 31) return 1;
            strcpy(c->ov_name[c->nov++], name);
        }
        c->ov_val[i] = v;
    }
    return 0;
}

/* run one translated subroutine on every rank, one thread per rank (like mpirun -np nproc) */
int ref_world_run(ref_world *w, const char *name)
{
    ref_job *jobs = (ref_job *)calloc((size_t)w->nproc, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)w->nproc, sizeof *th);
    int rc = 0;
    for (int r = 0; r < w->nproc; ++r) { jobs[r].S = w->st[r]; jobs[r].name = name; }
    for (int r = 1; r < w->nproc; ++r) pthread_create(&th[r], 0, ref_job_main, &jobs[r]);
    ref_job_main(&jobs[0]);
    for (int r = 1; r < w->nproc; ++r) pthread_join(th[r], 0);
    for (int r = 0; r < w->nproc; ++r) rc |= jobs[r].rc;
    free(jobs); free(th);
    return rc;
}

/* run `name_a; name_b` nsteps times on every rank without returning to the caller (timing) */
typedef struct ref_loop_job {
    ref_state *S;
    const char *a, *b;
    int nsteps;
} ref_loop_job;

static void *ref_loop_main(void *arg)
{
    ref_loop_job *j = (ref_loop_job *)arg;
    for (int i = 0; i < j->nsteps; ++i) {
        ref_dispatch(j->S, j->a);
        if (j->b && j->b[0]) ref_dispatch(j->S, j->b);
    }
    return 0;
}

int ref_world_loop(ref_world *w, const char *a, const char *b, int nsteps)
{
    ref_loop_job *jobs = (ref_loop_job *)calloc((size_t)w->nproc, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)w->nproc, sizeof *th);
    for (int r = 0; r < w->nproc; ++r) { jobs[r].S = w->st[r]; jobs[r].a = a; jobs[r].b = b; jobs[r].nsteps = nsteps; }
    for (int r = 1; r < w->nproc; ++r) pthread_create(&th[r], 0, ref_loop_main, &jobs[r]);
    ref_loop_main(&jobs[0]);
    for (int r = 1; r < w->nproc; ++r) pthread_join(th[r], 0);
    free(jobs); free(th);
    return 0;
}

#endif
