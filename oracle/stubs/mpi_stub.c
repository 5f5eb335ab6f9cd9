/* oracle/stubs/mpi_stub.c -- TEST INFRASTRUCTURE ONLY.  One-rank implementation of the MPI-1 subset
 * the CALCLENS SHTONLY path touches.  Collectives over one rank are copies;
 * point-to-point with a peer is unreachable on one rank and aborts loudly. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mpi.h"
static size_t tsize(MPI_Datatype t)
{
  switch (t) { case MPI_BYTE: case MPI_CHAR: return 1; case MPI_INT: case MPI_FLOAT: return 4;
               case MPI_LONG: case MPI_DOUBLE: return 8; default: fprintf(stderr, "mpi_stub: bad type %d\n", t); abort(); }
}
static void unreachable(const char *what) { fprintf(stderr, "mpi_stub: %s reached on a single rank\n", what); abort(); }
int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int code) { (void)c; fprintf(stderr, "MPI_Abort(%d)\n", code); exit(code ? code & 0xff ? code & 0xff : 1 : 0); }
int MPI_Comm_size(MPI_Comm c, int *n) { (void)c; *n = 1; return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return 0; }
double MPI_Wtime(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b; (void)n; (void)t; (void)root; (void)c; return 0; }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void)op; (void)root; (void)c; if (s != r) memmove(r, s, tsize(t) * (size_t)n); return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void)op; (void)c; if (s != r) memmove(r, s, tsize(t) * (size_t)n); return 0; }
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{ (void)nr; (void)tr; (void)c; if (s != r) memmove(r, s, tsize(ts) * (size_t)ns); return 0; }
int MPI_Alltoall(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{ (void)nr; (void)tr; (void)c; memmove(r, s, tsize(ts) * (size_t)ns); return 0; }
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype ts, void *r, const int *rc,
                  const int *rd, MPI_Datatype tr, MPI_Comm c)
{ (void)rc; (void)tr; (void)c; memmove((char*)r + tsize(ts) * (size_t)rd[0], (const char*)s + tsize(ts) * (size_t)sd[0], tsize(ts) * (size_t)sc[0]); return 0; }
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr,
                 int src, int rtag, MPI_Comm c, MPI_Status *st)
{ (void)stag; (void)rtag; (void)c; (void)nr; (void)tr;
  if (dest != 0 || src != 0) unreachable("MPI_Sendrecv with a peer");
  memmove(r, s, tsize(ts) * (size_t)ns); if (st) st->count_bytes = (int)(tsize(ts) * (size_t)ns); return 0; }
int MPI_Send(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)s; (void)n; (void)t; (void)d; (void)tag; (void)c; unreachable("MPI_Send"); return 0; }
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)s; (void)n; (void)t; (void)d; (void)tag; (void)c; unreachable("MPI_Ssend"); return 0; }
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) { (void)r; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)st; unreachable("MPI_Recv"); return 0; }
int MPI_Issend(const void *s, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request *rq) { (void)s; (void)n; (void)t; (void)d; (void)tag; (void)c; (void)rq; unreachable("MPI_Issend"); return 0; }
int MPI_Irecv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq) { (void)r; (void)n; (void)t; (void)src; (void)tag; (void)c; (void)rq; unreachable("MPI_Irecv"); return 0; }
int MPI_Wait(MPI_Request *rq, MPI_Status *st) { (void)rq; (void)st; return 0; }
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count) { *count = (int)(st->count_bytes / tsize(t)); return 0; }
int MPI_Comm_group(MPI_Comm c, MPI_Group *g) { (void)c; *g = 0; return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *ng) { (void)g; (void)n; (void)ranks; *ng = 0; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *nc) { (void)c; (void)g; *nc = 0; return 0; }
int MPI_Group_free(MPI_Group *g) { (void)g; return 0; }
int MPI_Comm_free(MPI_Comm *c) { (void)c; return 0; }
