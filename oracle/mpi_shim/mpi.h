/* mpi.h -- TEST INFRASTRUCTURE (oracle/): a minimal stand-in for MPI.
 *
 * The image has no MPI (no mpicc / mpiexec / mpi.h), but the reference's
 * diffusion_2D driver and SUNDIALS' nvector_parallel.c are MPI programs.  This
 * header plus mpi_shim.c implement exactly the subset those sources call
 * (listed in SURVEY.md section 8c) so that the UNMODIFIED reference sources
 * compile and run here:
 *
 *   - 1 rank by default;
 *   - MPISHIM_NP=<P> in the environment makes MPI_Init fork P-1 more ranks that
 *     talk through an anonymous shared-memory segment (eager single-slot
 *     mailboxes keyed by (dst, src, tag), a sense-reversing barrier, and an
 *     all-reduce that sums contributions in rank order on every rank).
 *
 * Nothing in the product links this; it exists only so the CPU reference can
 * be built into oracle/_ref/ and timed / compared against.
 */
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm; /* matches SUNComm == int when SUNDIALS_MPI_ENABLED is 0 */
typedef int MPI_Datatype;
typedef int MPI_Op;

typedef struct
{
  int kind; /* 0 = inactive, 1 = send (already delivered), 2 = pending receive */
  void* buf;
  size_t bytes;
  int peer;
  int tag;
} MPI_Request;

typedef struct
{
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
} MPI_Status;

#define MPI_SUCCESS    0
#define MPI_COMM_NULL  0
#define MPI_COMM_WORLD 1
#define MPI_CART       1
#define MPI_UNDEFINED  (-32766)

#define MPI_INT         1
#define MPI_FLOAT       2
#define MPI_DOUBLE      3
#define MPI_LONG_DOUBLE 4
#define MPI_INT32_T     5
#define MPI_INT64_T     6

#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

#define MPI_IN_PLACE ((void*)(intptr_t)(-1))

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
double MPI_Wtime(void);

int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_free(MPI_Comm* comm);

int MPI_Dims_create(int nnodes, int ndims, int dims[]);
int MPI_Cart_create(MPI_Comm comm, int ndims, const int dims[],
                    const int periods[], int reorder, MPI_Comm* comm_cart);
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[],
                 int coords[]);
int MPI_Cart_rank(MPI_Comm comm, const int coords[], int* rank);
int MPI_Cartdim_get(MPI_Comm comm, int* ndims);
int MPI_Topo_test(MPI_Comm comm, int* status);

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count,
                  MPI_Datatype datatype, MPI_Op op, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);

int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag,
              MPI_Comm comm, MPI_Request* request);
int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest,
              int tag, MPI_Comm comm, MPI_Request* request);
int MPI_Wait(MPI_Request* request, MPI_Status* status);

#ifdef __cplusplus
}
#endif

#endif /* ORACLE_MPI_SHIM_H */
