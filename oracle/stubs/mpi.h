/* oracle/stubs/mpi.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 * Declaration-level single-rank stand-in for <mpi.h> so that the UNMODIFIED CALCLENS sources under
 * /root/reference compile in a container that has no MPI.  Semantics: one rank, rank 0. */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
#include <stddef.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Group;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_bytes; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_LONG 8
#define MPI_DOUBLE 9
#define MPI_FLOAT 5
#define MPI_CHAR 2
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_LOR 4
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm c, int code);
int MPI_Comm_size(MPI_Comm c, int *n);
int MPI_Comm_rank(MPI_Comm c, int *r);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm c);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c);
int MPI_Alltoall(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c);
int MPI_Alltoallv(const void *s, const int *sc, const int *sd, MPI_Datatype ts, void *r, const int *rc,
                  const int *rd, MPI_Datatype tr, MPI_Comm c);
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr,
                 int src, int rtag, MPI_Comm c, MPI_Status *st);
int MPI_Send(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c);
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st);
int MPI_Issend(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Irecv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq);
int MPI_Wait(MPI_Request *rq, MPI_Status *st);
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count);
int MPI_Comm_group(MPI_Comm c, MPI_Group *g);
int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *ng);
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *nc);
int MPI_Group_free(MPI_Group *g);
int MPI_Comm_free(MPI_Comm *c);
#endif
