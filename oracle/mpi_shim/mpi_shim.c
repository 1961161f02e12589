/* TEST INFRASTRUCTURE ONLY (oracle/): the 17 MPI calls the reference uses (src/util/mp/DMPPolicy.h:103-343), for
 * ranks that live on ONE node.  See mpi.h in this directory for why it exists.
 *
 *   single rank (default)          rank 0 of 1; a periodic run still "sends to itself" (src/grid/grid_comm.cc:46-52), so
 *                                  Issend/Irecv pairs are matched by tag in either posting order.
 *   N ranks (oracle/mpi_shim/shimrun -n N prog ...)   VPIC_SHIM_SIZE / VPIC_SHIM_RANK / VPIC_SHIM_JOB in the environment:
 *                                  every ordered pair of ranks owns a mailbox in one POSIX shared-memory segment
 *                                  (descriptor ring + payload ring).  Sends are eager and buffered (Issend completes at
 *                                  once — a legal relaxation for this code, which never relies on the rendezvous),
 *                                  receives match by (source, tag) in posting order; collectives go through a slot per
 *                                  rank and a sense-reversing barrier.
 * That is enough to run the reference's own multi-rank decks (pcomm.deck, slab-decomposed samples) as the parity oracle
 * for the multi-GPU path, and to host the drop-in library under LD_PRELOAD with one process per GPU. */
#define _GNU_SOURCE
#include "mpi.h"
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

struct shim_mpi_request {
  int   is_send;
  int   tag;
  int   peer_rank;   /* source of a receive / destination of a send */
  int   bytes;       /* posted size; for a completed recv: delivered size */
  void *buf;
  int   done;
  struct shim_mpi_request *peer;
  int   live;
};

/* Up to 27 ports x (send+recv) can be in flight per grid; be generous. */
#define SHIM_MAX_REQ 512
static struct shim_mpi_request pool[SHIM_MAX_REQ];

static int g_rank = 0, g_size = 1;

static size_t type_bytes(MPI_Datatype t) {
  switch (t) {
    case MPI_BYTE: case MPI_CHAR: return 1;
    case MPI_INT: return sizeof(int);
    case MPI_LONG_LONG: return sizeof(long long);
    case MPI_DOUBLE: return sizeof(double);
  }
  fprintf(stderr, "mpi_shim: unknown datatype %d\n", t);
  abort();
}

static struct shim_mpi_request *grab(void) {
  for (int i = 0; i < SHIM_MAX_REQ; i++)
    if (!pool[i].live) { memset(&pool[i], 0, sizeof pool[i]); pool[i].live = 1; return &pool[i]; }
  fprintf(stderr, "mpi_shim: request table exhausted\n");
  abort();
}

/* ---- single-rank matching: deliver send -> recv if an unmatched partner with the same tag exists ------------ */
static void try_match(struct shim_mpi_request *r) {
  for (int i = 0; i < SHIM_MAX_REQ; i++) {
    struct shim_mpi_request *o = &pool[i];
    if (!o->live || o == r || o->done || o->peer) continue;
    if (o->is_send == r->is_send || o->tag != r->tag) continue;
    struct shim_mpi_request *s = r->is_send ? r : o, *d = r->is_send ? o : r;
    if (s->bytes > d->bytes) { fprintf(stderr, "mpi_shim: truncated message tag %d\n", r->tag); abort(); }
    memcpy(d->buf, s->buf, (size_t)s->bytes);
    d->bytes = s->bytes;
    s->done = d->done = 1;
    s->peer = d; d->peer = s;
    return;
  }
}

/* ---- N ranks: shared-memory mailboxes ---------------------------------------------------------------------------- */
#define NDESC 512
enum { D_FREE = 0, D_POSTED = 1, D_CONSUMED = 2 };
struct desc { _Atomic int state; int tag; int bytes; size_t off; };
struct mailbox {                      /* one per ordered pair (src -> dst); written by src, drained by dst */
  _Atomic unsigned long head;         /* descriptors posted so far (producer) */
  unsigned long tail;                 /* producer's view of the oldest descriptor not yet reclaimed */
  size_t heap_head, heap_tail;        /* producer's payload ring cursors (bytes, monotonically increasing) */
  unsigned long rd;                   /* consumer: first descriptor not yet consumed */
  struct desc d[NDESC];
};
struct shared_hdr {
  _Atomic int attached;
  _Atomic int barrier_count;
  _Atomic int barrier_gen;
  _Atomic int aborted;
};
#define COLL_SLOT (1u << 20)          /* bytes per rank for collectives */
static struct shared_hdr *g_hdr;
static struct mailbox *g_box;         /* [src * size + dst] */
static unsigned char *g_heap;         /* [src * size + dst] payload rings of g_heap_bytes each */
static unsigned char *g_coll;         /* [rank] slots of COLL_SLOT bytes */
static size_t g_heap_bytes;
static char g_shm_name[128];

static void die(const char *what) {
  fprintf(stderr, "mpi_shim[%d]: %s\n", g_rank, what);
  if (g_hdr) atomic_store(&g_hdr->aborted, 1);
  abort();
}
static void relax(void) {
  if (g_hdr && atomic_load(&g_hdr->aborted)) { fprintf(stderr, "mpi_shim[%d]: another rank aborted\n", g_rank); _exit(1); }
  sched_yield();
}

static void shm_attach(void) {
  const char *job = getenv("VPIC_SHIM_JOB");
  const char *heap_mb = getenv("VPIC_SHIM_HEAP_MB");
  g_heap_bytes = (size_t)(heap_mb ? atol(heap_mb) : 64) << 20;
  snprintf(g_shm_name, sizeof g_shm_name, "/vpicshim_%s", job ? job : "default");
  const size_t npair = (size_t)g_size * g_size;
  const size_t off_box = 4096, off_coll = off_box + npair * sizeof(struct mailbox);
  const size_t off_heap = (off_coll + (size_t)g_size * COLL_SLOT + 4095) & ~(size_t)4095;
  const size_t total = off_heap + npair * g_heap_bytes;
  int fd = -1;
  for (int tries = 0; tries < 20000 && fd < 0; tries++) {       /* rank 0 creates, the others wait for it */
    fd = g_rank == 0 ? shm_open(g_shm_name, O_CREAT | O_RDWR, 0600) : shm_open(g_shm_name, O_RDWR, 0600);
    if (fd < 0) { struct timespec ts = {0, 1000000}; nanosleep(&ts, 0); }
  }
  if (fd < 0) die("cannot open the shared segment");
  if (g_rank == 0 && ftruncate(fd, (off_t)total) != 0) die("cannot size the shared segment");
  if (g_rank != 0) {                                            /* wait until rank 0 has sized it */
    struct stat st;
    for (int tries = 0; tries < 20000; tries++) { if (fstat(fd, &st) == 0 && (size_t)st.st_size >= total) break; struct timespec ts = {0, 1000000}; nanosleep(&ts, 0); }
  }
  unsigned char *base = (unsigned char *)mmap(0, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (base == MAP_FAILED) die("cannot map the shared segment");
  g_hdr = (struct shared_hdr *)base;
  g_box = (struct mailbox *)(base + off_box);
  g_coll = base + off_coll;
  g_heap = base + off_heap;
  atomic_fetch_add(&g_hdr->attached, 1);
  while (atomic_load(&g_hdr->attached) < g_size) relax();
  if (g_rank == 0) shm_unlink(g_shm_name);                      /* everybody is mapped: the name can go */
}

static void shm_barrier(void) {
  const int gen = atomic_load(&g_hdr->barrier_gen);
  if (atomic_fetch_add(&g_hdr->barrier_count, 1) == g_size - 1) {
    atomic_store(&g_hdr->barrier_count, 0);
    atomic_fetch_add(&g_hdr->barrier_gen, 1);
  } else {
    while (atomic_load(&g_hdr->barrier_gen) == gen) relax();
  }
}

static void shm_send(const void *buf, int bytes, int dst, int tag) {
  if (dst < 0 || dst >= g_size) die("send to a rank outside the job");
  struct mailbox *m = &g_box[(size_t)g_rank * g_size + dst];
  unsigned char *heap = g_heap + ((size_t)g_rank * g_size + dst) * g_heap_bytes;
  if ((size_t)bytes > g_heap_bytes / 2) die("message larger than half the payload ring (raise VPIC_SHIM_HEAP_MB)");
  for (;;) {
    /* reclaim what the receiver has consumed, oldest first */
    unsigned long head = atomic_load(&m->head);
    while (m->tail < head && atomic_load(&m->d[m->tail % NDESC].state) == D_CONSUMED) {
      struct desc *o = &m->d[m->tail % NDESC];
      m->heap_tail = o->off + (size_t)o->bytes;
      atomic_store(&o->state, D_FREE);
      m->tail++;
    }
    if (m->tail == head) m->heap_tail = m->heap_head;
    /* the payload must be contiguous: skip the end of the ring when it does not fit there */
    size_t start = m->heap_head;
    if (start % g_heap_bytes + (size_t)bytes > g_heap_bytes) start += g_heap_bytes - start % g_heap_bytes;
    if (head - m->tail < NDESC && start + (size_t)bytes - m->heap_tail <= g_heap_bytes) {
      struct desc *d = &m->d[head % NDESC];
      memcpy(heap + start % g_heap_bytes, buf, (size_t)bytes);
      d->tag = tag; d->bytes = bytes; d->off = start;
      m->heap_head = start + (size_t)bytes;
      atomic_store(&d->state, D_POSTED);
      atomic_store(&m->head, head + 1);
      return;
    }
    relax();                                                     /* ring full: wait for the receiver */
  }
}

/* returns 1 and fills the request when a message (src -> me, tag) is available */
static int shm_try_recv(struct shim_mpi_request *r) {
  struct mailbox *m = &g_box[(size_t)r->peer_rank * g_size + g_rank];
  unsigned char *heap = g_heap + ((size_t)r->peer_rank * g_size + g_rank) * g_heap_bytes;
  const unsigned long head = atomic_load(&m->head);
  for (unsigned long i = m->rd; i < head; i++) {
    struct desc *d = &m->d[i % NDESC];
    if (atomic_load(&d->state) != D_POSTED || d->tag != r->tag) continue;
    if (d->bytes > r->bytes) die("truncated message");
    memcpy(r->buf, heap + d->off % g_heap_bytes, (size_t)d->bytes);
    r->bytes = d->bytes;
    atomic_store(&d->state, D_CONSUMED);
    while (m->rd < head && atomic_load(&m->d[m->rd % NDESC].state) != D_POSTED) m->rd++;
    r->done = 1;
    return 1;
  }
  return 0;
}

/* ---- the MPI surface ------------------------------------------------------------------------------------------- */
int MPI_Init(int *argc, char ***argv) {
  (void)argc; (void)argv;
  const char *s = getenv("VPIC_SHIM_SIZE"), *r = getenv("VPIC_SHIM_RANK");
  g_size = s ? atoi(s) : 1;
  g_rank = r ? atoi(r) : 0;
  if (g_size < 1 || g_rank < 0 || g_rank >= g_size) { fprintf(stderr, "mpi_shim: bad VPIC_SHIM_RANK/SIZE\n"); abort(); }
  if (g_size > 1) shm_attach();
  return MPI_SUCCESS;
}
int MPI_Finalize(void) { if (g_size > 1) shm_barrier(); return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; if (g_hdr) atomic_store(&g_hdr->aborted, 1); exit(code ? code : 1); }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) { *out = c; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm *c) { *c = MPI_COMM_SELF; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { *rank = c == MPI_COMM_SELF ? 0 : g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { *size = c == MPI_COMM_SELF ? 1 : g_size; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm c) { if (g_size > 1 && c != MPI_COMM_SELF) shm_barrier(); return MPI_SUCCESS; }

/* every rank writes its contribution into its slot, then everybody reads what it needs */
static void coll_publish(const void *src, size_t bytes) {
  if (bytes > COLL_SLOT) die("collective payload exceeds 1 MB per rank");
  shm_barrier();                                                 /* the previous collective has been read by all */
  memcpy(g_coll + (size_t)g_rank * COLL_SLOT, src, bytes);
  shm_barrier();
}

int MPI_Allreduce(const void *src, void *dst, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op;
  const size_t bytes = (size_t)n * type_bytes(t);
  if (g_size == 1 || c == MPI_COMM_SELF) { memmove(dst, src, bytes); return MPI_SUCCESS; }
  coll_publish(src, bytes);
  /* sum in rank order on every rank: all ranks get bit-identical results */
  for (int k = 0; k < n; k++) {
    if (t == MPI_DOUBLE) { double s = 0; for (int r = 0; r < g_size; r++) s += ((double *)(g_coll + (size_t)r * COLL_SLOT))[k]; ((double *)dst)[k] = s; }
    else if (t == MPI_INT) { int s = 0; for (int r = 0; r < g_size; r++) s += ((int *)(g_coll + (size_t)r * COLL_SLOT))[k]; ((int *)dst)[k] = s; }
    else if (t == MPI_LONG_LONG) { long long s = 0; for (int r = 0; r < g_size; r++) s += ((long long *)(g_coll + (size_t)r * COLL_SLOT))[k]; ((long long *)dst)[k] = s; }
    else die("Allreduce: unsupported datatype");
  }
  return MPI_SUCCESS;
}
int MPI_Allgather(const void *src, int ns, MPI_Datatype ts, void *dst, int nd, MPI_Datatype td, MPI_Comm c) {
  (void)nd; (void)td;
  const size_t bytes = (size_t)ns * type_bytes(ts);
  if (g_size == 1 || c == MPI_COMM_SELF) { memmove(dst, src, bytes); return MPI_SUCCESS; }
  coll_publish(src, bytes);
  for (int r = 0; r < g_size; r++) memcpy((unsigned char *)dst + (size_t)r * bytes, g_coll + (size_t)r * COLL_SLOT, bytes);
  return MPI_SUCCESS;
}
int MPI_Gather(const void *src, int ns, MPI_Datatype ts, void *dst, int nd, MPI_Datatype td, int root, MPI_Comm c) {
  (void)nd; (void)td;
  const size_t bytes = (size_t)ns * type_bytes(ts);
  if (g_size == 1 || c == MPI_COMM_SELF) { memmove(dst, src, bytes); return MPI_SUCCESS; }
  coll_publish(src, bytes);
  if (g_rank == root) for (int r = 0; r < g_size; r++) memcpy((unsigned char *)dst + (size_t)r * bytes, g_coll + (size_t)r * COLL_SLOT, bytes);
  return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req) {
  (void)c;
  struct shim_mpi_request *r = grab();
  r->is_send = 0; r->tag = tag; r->peer_rank = src; r->bytes = n * (int)type_bytes(t); r->buf = buf;
  if (g_size == 1) try_match(r);          /* N ranks: matched in MPI_Wait (self-messages use the rank's own mailbox) */
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Issend(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *req) {
  (void)c;
  struct shim_mpi_request *r = grab();
  r->is_send = 1; r->tag = tag; r->peer_rank = dst; r->bytes = n * (int)type_bytes(t); r->buf = (void *)buf;
  if (g_size == 1) try_match(r);
  else { shm_send(buf, r->bytes, dst, tag); r->done = 1; }
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *req, MPI_Status *st) {
  struct shim_mpi_request *r = *req;
  if (!r || !r->live) { fprintf(stderr, "mpi_shim: wait on dead request\n"); abort(); }
  if (g_size == 1) {
    if (!r->done) try_match(r);
    if (!r->done) { fprintf(stderr, "mpi_shim: wait would deadlock (tag %d, %s)\n", r->tag, r->is_send ? "send" : "recv"); abort(); }
  } else {
    while (!r->done) { if (!shm_try_recv(r)) relax(); }
  }
  if (st) st->byte_count = r->bytes;
  r->live = 0;
  *req = 0;
  return MPI_SUCCESS;
}
int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
  if (g_size == 1) { fprintf(stderr, "mpi_shim: blocking MPI_Send is unreachable at world_size 1\n"); abort(); }
  MPI_Request rq; MPI_Issend(b, n, t, dst, tag, c, &rq); return MPI_Wait(&rq, MPI_STATUS_IGNORE);
}
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) {
  if (g_size == 1) { fprintf(stderr, "mpi_shim: blocking MPI_Recv is unreachable at world_size 1\n"); abort(); }
  MPI_Request rq; MPI_Irecv(b, n, t, src, tag, c, &rq); return MPI_Wait(&rq, st);
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) {
  *count = st->byte_count / (int)type_bytes(t);
  return MPI_SUCCESS;
}
