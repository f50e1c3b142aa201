/* mpi.h -- single-node mini-MPI used when no real MPI installation exists.
 *
 * The P3DFFT++ API carries MPI_Comm arguments (reference include/Cwrap.h:94,
 * build/init.C:1631-1672) and its samples call ~20 MPI entry points.  This image has
 * no MPI, so the B200 build ships this header plus p3dfft.3_b200/host/minimpi.cpp:
 * one process per rank (= per GPU) on ONE host, rendezvous through POSIX shared
 * memory.  Ranks are taken from the environment, in this order:
 *   P3DFFT_RANK / P3DFFT_NRANKS / P3DFFT_SESSION   (tools/mpirun.py)
 *   RANK / WORLD_SIZE / MASTER_PORT                (torchrun)
 * With neither set the process is a 1-rank world.
 * Build against a real MPI by putting its mpi.h ahead of include/compat on the
 * include path; nothing else in the library depends on this file.
 */
#ifndef P3DFFT_B200_COMPAT_MPI_H
#define P3DFFT_B200_COMPAT_MPI_H

#define P3DFFT_B200_MINIMPI 1

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Fint;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL  (-1)
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF  1

#define MPI_IDENT     0
#define MPI_CONGRUENT 1
#define MPI_SIMILAR   2
#define MPI_UNEQUAL   3

#define MPI_ANY_TAG    (-1)
#define MPI_ANY_SOURCE (-2)
#define MPI_STATUS_IGNORE   ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)

/* datatypes: value encodes nothing; sizes come from a table in minimpi.cpp */
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_LONG 4
#define MPI_LONG_LONG 5
#define MPI_UNSIGNED 6
#define MPI_UNSIGNED_LONG 7
#define MPI_FLOAT 8
#define MPI_REAL 8
#define MPI_DOUBLE 9
#define MPI_DOUBLE_PRECISION 9
#define MPI_COMPLEX 10
#define MPI_DOUBLE_COMPLEX 11
#define MPI_INTEGER 3

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_PROD 4

int MPI_Init(int *argc, char ***argv);
int MPI_Initialized(int *flag);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
double MPI_Wtime(void);

int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *result);
int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *out);
MPI_Comm MPI_Comm_f2c(MPI_Fint f);
MPI_Fint MPI_Comm_c2f(MPI_Comm c);

int MPI_Dims_create(int nnodes, int ndims, int *dims);
int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int reorder, MPI_Comm *out);
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords);
int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank);
int MPI_Cart_sub(MPI_Comm comm, const int *remain_dims, MPI_Comm *out);

int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype dt, int root, MPI_Comm comm);
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm comm);
int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Gather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, int root, MPI_Comm comm);
int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype sdt, void *rbuf, int rcount, MPI_Datatype rdt, MPI_Comm comm);
int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype sdt,
                  void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rdt, MPI_Comm comm);

/* point-to-point: declared so that third-party code links; not implemented (abort) */
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Isend(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts);

#ifdef __cplusplus
}
#endif
#endif
