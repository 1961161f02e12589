/* TEST INFRASTRUCTURE ONLY (oracle/): rank-0-of-1 implementation of mpi.h.
 * See mpi.h in this directory for why it exists. */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct shim_mpi_request {
  int   is_send;
  int   tag;
  int   bytes;       /* posted size; for a completed recv: delivered size */
  void *buf;
  int   done;
  struct shim_mpi_request *peer;
  int   live;
};

/* Up to 27 ports x (send+recv) can be in flight per grid; be generous. */
#define SHIM_MAX_REQ 512
static struct shim_mpi_request pool[SHIM_MAX_REQ];

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

/* Deliver send -> recv if an unmatched partner with the same tag exists. */
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

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; exit(code ? code : 1); }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) { *out = c; return MPI_SUCCESS; }
int MPI_Comm_free(MPI_Comm *c) { *c = MPI_COMM_SELF; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size) { (void)c; *size = 1; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }

int MPI_Allreduce(const void *src, void *dst, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  (void)op; (void)c; memmove(dst, src, (size_t)n * type_bytes(t)); return MPI_SUCCESS;
}
int MPI_Allgather(const void *src, int ns, MPI_Datatype ts, void *dst, int nd, MPI_Datatype td, MPI_Comm c) {
  (void)nd; (void)td; (void)c; memmove(dst, src, (size_t)ns * type_bytes(ts)); return MPI_SUCCESS;
}
int MPI_Gather(const void *src, int ns, MPI_Datatype ts, void *dst, int nd, MPI_Datatype td, int root, MPI_Comm c) {
  (void)nd; (void)td; (void)root; (void)c; memmove(dst, src, (size_t)ns * type_bytes(ts)); return MPI_SUCCESS;
}
int MPI_Send(const void *b, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
  (void)b; (void)n; (void)t; (void)dst; (void)tag; (void)c;
  fprintf(stderr, "mpi_shim: blocking MPI_Send is unreachable at world_size 1\n"); abort();
}
int MPI_Recv(void *b, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) {
  (void)b; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)st;
  fprintf(stderr, "mpi_shim: blocking MPI_Recv is unreachable at world_size 1\n"); abort();
}

int MPI_Irecv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req) {
  (void)src; (void)c;
  struct shim_mpi_request *r = grab();
  r->is_send = 0; r->tag = tag; r->bytes = n * (int)type_bytes(t); r->buf = buf;
  try_match(r);
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Issend(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *req) {
  (void)dst; (void)c;
  struct shim_mpi_request *r = grab();
  r->is_send = 1; r->tag = tag; r->bytes = n * (int)type_bytes(t); r->buf = (void *)buf;
  try_match(r);
  *req = r;
  return MPI_SUCCESS;
}
int MPI_Wait(MPI_Request *req, MPI_Status *st) {
  struct shim_mpi_request *r = *req;
  if (!r || !r->live) { fprintf(stderr, "mpi_shim: wait on dead request\n"); abort(); }
  if (!r->done) try_match(r);
  if (!r->done) { fprintf(stderr, "mpi_shim: wait would deadlock (tag %d, %s)\n", r->tag, r->is_send ? "send" : "recv"); abort(); }
  if (st) st->byte_count = r->bytes;
  r->live = 0;
  *req = 0;
  return MPI_SUCCESS;
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) {
  *count = st->byte_count / (int)type_bytes(t);
  return MPI_SUCCESS;
}
