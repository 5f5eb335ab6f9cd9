/* oracle/stubs/mpi_stub.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Stand-in for the MPI-1 subset the CALCLENS SHTONLY path touches, for a container that has no
 * MPI.  Two modes, chosen at the first MPI call from the environment:
 *   - single rank (default): collectives are copies; point-to-point with a peer aborts loudly.
 *   - N ranks = N processes on this host (CLB_MPI_NTASKS, CLB_MPI_RANK, CLB_MPI_SHM = a file under /dev/shm created by
 *     the launcher, oracle/mpirun.py): messages travel through a shared-memory segment.  Point-to-point uses one
 *     mailbox per ordered rank pair with a full/empty handshake, so the reference's hypercube MPI_Sendrecv exchanges
 *     (map2alm_transpose_mpi.c:356-381, alm2allmaps_transpose_mpi.c:699-724, map_shuffle.c) really transpose;
 *     collectives go through one slot per rank between two barriers and reduce in rank order (deterministic).
 * This is what lets the reference's own MPI code run on all host cores as the CPU baseline ("mpirun raytrace"
 * substitute) and lets the drop-in shim be tested with several ranks. */
#define _GNU_SOURCE
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include "mpi.h"

#define CHAN_BYTES (256L * 1024)
#define COLL_BYTES (4L * 1024 * 1024)

typedef struct {
  volatile long full;            /* 0 = empty, otherwise payload bytes + 1 */
  char pad[56];
  char data[CHAN_BYTES];
} Chan;

typedef struct {
  volatile int count, sense;
  volatile int aborted;
  char pad[52];
} Hdr;

static int g_init = 0, g_n = 1, g_rank = 0, g_sense = 0;
static Hdr *g_hdr = NULL;
static Chan *g_chan = NULL;      /* [src * n + dst] */
static char *g_coll = NULL;      /* [rank][COLL_BYTES] */

static size_t tsize(MPI_Datatype t)
{
  switch (t) { case MPI_BYTE: case MPI_CHAR: return 1; case MPI_INT: case MPI_FLOAT: return 4;
               case MPI_LONG: case MPI_DOUBLE: return 8; default: fprintf(stderr, "mpi_stub: bad type %d\n", t); abort(); }
}

size_t clb_mpi_shm_bytes(int n) { return sizeof(Hdr) + sizeof(Chan) * (size_t)n * n + (size_t)COLL_BYTES * n; }

static void init_once(void)
{
  if (g_init) return;
  g_init = 1;
  const char *sn = getenv("CLB_MPI_NTASKS"), *sr = getenv("CLB_MPI_RANK"), *sp = getenv("CLB_MPI_SHM");
  if (!sn || !sr || !sp || atoi(sn) <= 1) return;
  g_n = atoi(sn); g_rank = atoi(sr);
  int fd = open(sp, O_RDWR);
  if (fd < 0) { perror("mpi_stub: open CLB_MPI_SHM"); abort(); }
  size_t bytes = clb_mpi_shm_bytes(g_n);
  void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (p == MAP_FAILED) { perror("mpi_stub: mmap"); abort(); }
  close(fd);
  g_hdr = (Hdr *)p;
  g_chan = (Chan *)((char *)p + sizeof(Hdr));
  g_coll = (char *)p + sizeof(Hdr) + sizeof(Chan) * (size_t)g_n * g_n;
}

static void check_abort(void)
{
  if (g_hdr && g_hdr->aborted) { fprintf(stderr, "mpi_stub: rank %d leaves after a peer aborted\n", g_rank); _exit(g_hdr->aborted & 0xff ? g_hdr->aborted & 0xff : 1); }
}
static void relax(int *spins)
{
  if (++*spins > 64) { sched_yield(); if ((*spins & 1023) == 0) check_abort(); }
}

static void barrier(void)
{
  if (g_n == 1) return;
  g_sense = !g_sense;
  if (__atomic_add_fetch(&g_hdr->count, 1, __ATOMIC_ACQ_REL) == g_n) {
    __atomic_store_n(&g_hdr->count, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&g_hdr->sense, g_sense, __ATOMIC_RELEASE);
  } else {
    int spins = 0;
    while (__atomic_load_n(&g_hdr->sense, __ATOMIC_ACQUIRE) != g_sense) relax(&spins);
  }
}

/* progress a send to `dest` and a receive from `src` together (either may be absent: NULL buffer and peer < 0) */
static void exchange(const char *sbuf, size_t sbytes, int dest, char *rbuf, size_t rbytes, int src, size_t *got)
{
  size_t sent = 0, recvd = 0;
  int send_done = dest < 0, recv_done = src < 0, first_send = 1, spins = 0;
  Chan *out = dest >= 0 ? &g_chan[(size_t)g_rank * g_n + dest] : NULL;
  Chan *in = src >= 0 ? &g_chan[(size_t)src * g_n + g_rank] : NULL;
  size_t expect = 0; int have_hdr = 0;
  while (!send_done || !recv_done) {
    int progressed = 0;
    if (!send_done && __atomic_load_n(&out->full, __ATOMIC_ACQUIRE) == 0) {
      /* first chunk carries the total length in its first 8 bytes */
      size_t room = CHAN_BYTES, off = 0;
      if (first_send) { memcpy(out->data, &sbytes, sizeof(size_t)); off = sizeof(size_t); room -= off; first_send = 0; }
      size_t n = sbytes - sent < room ? sbytes - sent : room;
      memcpy(out->data + off, sbuf + sent, n);
      sent += n;
      __atomic_store_n(&out->full, (long)(off + n) + 1, __ATOMIC_RELEASE);
      if (sent == sbytes) send_done = 1;
      progressed = 1;
    }
    if (!recv_done) {
      long f = __atomic_load_n(&in->full, __ATOMIC_ACQUIRE);
      if (f != 0) {
        size_t n = (size_t)(f - 1), off = 0;
        if (!have_hdr) { memcpy(&expect, in->data, sizeof(size_t)); off = sizeof(size_t); have_hdr = 1;
          if (expect > rbytes) { fprintf(stderr, "mpi_stub: rank %d receives %zu bytes from %d into a %zu-byte buffer\n", g_rank, expect, src, rbytes); abort(); } }
        memcpy(rbuf + recvd, in->data + off, n - off);
        recvd += n - off;
        __atomic_store_n(&in->full, 0, __ATOMIC_RELEASE);
        if (recvd == expect) recv_done = 1;
        progressed = 1;
      }
    }
    if (progressed) spins = 0; else relax(&spins);
  }
  if (got) *got = recvd;
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; init_once(); return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int code)
{
  (void)c; init_once();
  fprintf(stderr, "MPI_Abort(%d) on rank %d\n", code, g_rank);
  if (g_hdr) g_hdr->aborted = code ? code : 1;
  exit(code ? (code & 0xff ? code & 0xff : 1) : 0);
}
int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; init_once(); *n = g_n; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; init_once(); *r = g_rank; return 0; }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Barrier(MPI_Comm c) { (void)c; init_once(); barrier(); return 0; }

/* every rank publishes `bytes` bytes (in chunks of COLL_BYTES); fn(ctx, rank q, chunk offset, chunk pointer, chunk bytes)
 * is called for every rank's chunk between the two barriers */
typedef void (*visit_fn)(void *ctx, int q, size_t off, const char *chunk, size_t n);
static void all_publish(const char *mine, size_t bytes, visit_fn fn, void *ctx)
{
  size_t off = 0;
  do {
    size_t n = bytes - off < (size_t)COLL_BYTES ? bytes - off : (size_t)COLL_BYTES;
    if (mine) memcpy(g_coll + (size_t)g_rank * COLL_BYTES, mine + off, n);
    barrier();
    for (int q = 0; q < g_n; ++q) fn(ctx, q, off, g_coll + (size_t)q * COLL_BYTES, n);
    barrier();
    off += n;
  } while (off < bytes);
}

typedef struct { char *r; size_t each; int root; int me; } GatherCtx;
static void visit_allgather(void *c, int q, size_t off, const char *chunk, size_t n)
{ GatherCtx *g = (GatherCtx *)c; memcpy(g->r + (size_t)q * g->each + off, chunk, n); }
static void visit_bcast(void *c, int q, size_t off, const char *chunk, size_t n)
{ GatherCtx *g = (GatherCtx *)c; if (q == g->root && g->me != g->root) memcpy(g->r + off, chunk, n); }

int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c)
{
  (void)c; init_once();
  if (g_n == 1) return 0;
  GatherCtx g = {(char *)b, 0, root, g_rank};
  all_publish(g_rank == root ? (const char *)b : NULL, tsize(t) * (size_t)n, visit_bcast, &g);
  return 0;
}
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{
  (void)nr; (void)tr; (void)c; init_once();
  size_t each = tsize(ts) * (size_t)ns;
  if (g_n == 1) { if (s != r) memmove(r, s, each); return 0; }
  GatherCtx g = {(char *)r, each, 0, g_rank};
  all_publish((const char *)s, each, visit_allgather, &g);
  return 0;
}

static void reduce_into(void *acc, const void *x, size_t n, MPI_Datatype t, MPI_Op op)
{
#define RED(T) do { T *a = (T *)acc; const T *b = (const T *)x; for (size_t i = 0; i < n; ++i) { \
    if (op == MPI_SUM) a[i] = a[i] + b[i]; else if (op == MPI_MAX) a[i] = a[i] > b[i] ? a[i] : b[i]; \
    else if (op == MPI_MIN) a[i] = a[i] < b[i] ? a[i] : b[i]; else if (op == MPI_LOR) a[i] = (T)((a[i] != 0) || (b[i] != 0)); \
    else { fprintf(stderr, "mpi_stub: bad op %d\n", op); abort(); } } } while (0)
  switch (t) { case MPI_INT: RED(int); break; case MPI_LONG: RED(long); break; case MPI_DOUBLE: RED(double); break;
               case MPI_FLOAT: RED(float); break; case MPI_CHAR: case MPI_BYTE: RED(char); break; default: abort(); }
#undef RED
}
static int allreduce_impl(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root)
{
  size_t each = tsize(t) * (size_t)n;
  if (g_n == 1) { if (s != r) memmove(r, s, each); return 0; }
  char *all = (char *)malloc(each * (size_t)g_n);
  GatherCtx g = {all, each, 0, g_rank};
  all_publish((const char *)s, each, visit_allgather, &g);
  if (root < 0 || root == g_rank) {
    memcpy(r, all, each);
    for (int q = 1; q < g_n; ++q) reduce_into(r, all + (size_t)q * each, (size_t)n, t, op);
  }
  free(all);
  return 0;
}
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)c; init_once(); return allreduce_impl(s, r, n, t, op, root); }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)c; init_once(); return allreduce_impl(s, r, n, t, op, -1); }

int MPI_Alltoall(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{
  (void)nr; (void)tr; (void)c; init_once();
  size_t each = tsize(ts) * (size_t)ns;
  if (g_n == 1) { memmove(r, s, each); return 0; }
  char *all = (char *)malloc(each * (size_t)g_n * g_n);
  GatherCtx g = {all, each * (size_t)g_n, 0, g_rank};
  all_publish((const char *)s, each * (size_t)g_n, visit_allgather, &g);
  for (int q = 0; q < g_n; ++q) memcpy((char *)r + (size_t)q * each, all + ((size_t)q * g_n + g_rank) * each, each);
  free(all);
  return 0;
}
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype ts, void *r, const int *rc,
                  const int *rd, MPI_Datatype tr, MPI_Comm c)
{
  (void)tr; (void)c; init_once();
  size_t e = tsize(ts);
  if (g_n == 1) { memmove((char *)r + e * (size_t)rd[0], (const char *)s + e * (size_t)sd[0], e * (size_t)sc[0]); return 0; }
  memmove((char *)r + e * (size_t)rd[g_rank], (const char *)s + e * (size_t)sd[g_rank], e * (size_t)sc[g_rank]);
  for (int k = 1; k < g_n; ++k) {   /* ring schedule: send to rank+k, receive from rank-k */
    int dest = (g_rank + k) % g_n, src = (g_rank - k + g_n) % g_n;
    exchange((const char *)s + e * (size_t)sd[dest], e * (size_t)sc[dest], dest, (char *)r + e * (size_t)rd[src], e * (size_t)rc[src], src, NULL);
  }
  return 0;
}
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr,
                 int src, int rtag, MPI_Comm c, MPI_Status *st)
{
  (void)stag; (void)rtag; (void)c; init_once();
  size_t sb = tsize(ts) * (size_t)ns, rb = tsize(tr) * (size_t)nr, got = 0;
  if (dest == g_rank && src == g_rank) { memmove(r, s, sb); got = sb; }
  else if (g_n == 1) { fprintf(stderr, "mpi_stub: MPI_Sendrecv with a peer reached on a single rank\n"); abort(); }
  else exchange((const char *)s, sb, dest, (char *)r, rb, src, &got);
  if (st) { st->MPI_SOURCE = src; st->count_bytes = (int)got; }
  return 0;
}
int MPI_Send(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c)
{
  (void)tag; (void)c; init_once();
  if (g_n == 1) { fprintf(stderr, "mpi_stub: MPI_Send reached on a single rank\n"); abort(); }
  exchange((const char *)s, tsize(t) * (size_t)n, d, NULL, 0, -1, NULL);
  return 0;
}
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { return MPI_Send(s, n, t, d, tag, c); }
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st)
{
  (void)tag; (void)c; init_once();
  if (g_n == 1) { fprintf(stderr, "mpi_stub: MPI_Recv reached on a single rank\n"); abort(); }
  size_t got = 0;
  exchange(NULL, 0, -1, (char *)r, tsize(t) * (size_t)n, src, &got);
  if (st) { st->MPI_SOURCE = src; st->count_bytes = (int)got; }
  return 0;
}
static void unsupported(const char *what) { fprintf(stderr, "mpi_stub: %s is not implemented (not on the SHTONLY map-input path)\n", what); abort(); }
int MPI_Issend(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *rq) { (void)s; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)rq; unsupported("MPI_Issend"); return 0; }
int MPI_Irecv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq) { (void)r; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)rq; unsupported("MPI_Irecv"); return 0; }
int MPI_Wait(MPI_Request *rq, MPI_Status *st) { (void)rq; (void)st; return 0; }
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) { *count = (int)(st->count_bytes / tsize(t)); return 0; }
int MPI_Comm_group(MPI_Comm c, MPI_Group *g) { (void)c; *g = 0; return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *ng) { (void)g; (void)n; (void)ranks; *ng = 0; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *nc) { (void)c; (void)g; *nc = 0; return 0; }
int MPI_Group_free(MPI_Group *g) { (void)g; return 0; }
int MPI_Comm_free(MPI_Comm *c) { (void)c; return 0; }
