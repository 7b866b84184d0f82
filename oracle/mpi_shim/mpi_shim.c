/* mpi_shim.c -- TEST INFRASTRUCTURE (oracle/): see mpi.h.
 *
 * Multi-rank mode: MPI_Init reads MPISHIM_NP; rank 0 maps a shared segment and
 * forks the other ranks, all of which return from MPI_Init and run main() as
 * independent "MPI processes".  Point-to-point messages are eager: Isend copies
 * into the single mailbox slot (dst, src, tag) and returns; Irecv records the
 * destination buffer; MPI_Wait on a receive spins until the slot is full and
 * copies out.  That is sufficient (and deadlock-free) for the reference's halo
 * exchange, which posts 4 Irecv, 4 Isend and then waits on all 8
 * (/root/reference/diffusion_2D/diffusion_2D.cpp:400-584).
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define SHIM_MAX_RANKS 64
#define SHIM_NTAGS     4
#define SHIM_SLOT_BYTES (1u << 18) /* 32768 doubles per message */
#define SHIM_RED_MAX   64          /* doubles per all-reduce */

typedef struct
{
  volatile int full;
  int pad[15];
  unsigned char data[SHIM_SLOT_BYTES];
} shim_slot;

typedef struct
{
  volatile int bar_count;
  volatile int bar_sense;
  int pad[14];
  double red[SHIM_MAX_RANKS][SHIM_RED_MAX];
  /* slots follow: [dst][src][tag] */
} shim_shared;

static int g_np            = 1;
static int g_rank          = 0;
static int g_local_sense   = 0;
static shim_shared* g_sh   = NULL;
static shim_slot* g_slots  = NULL;
static pid_t g_children[SHIM_MAX_RANKS];

/* Cartesian topology of communicator 2 (the only one the reference creates) */
static int g_cart_dims[2]    = {1, 1};
static int g_cart_periods[2] = {1, 1};

static shim_slot* slot_of(int dst, int src, int tag)
{
  return &g_slots[((size_t)dst * g_np + src) * SHIM_NTAGS + tag];
}

static size_t dtype_size(MPI_Datatype t)
{
  switch (t)
  {
  case MPI_INT: return sizeof(int);
  case MPI_FLOAT: return sizeof(float);
  case MPI_DOUBLE: return sizeof(double);
  case MPI_LONG_DOUBLE: return sizeof(long double);
  case MPI_INT32_T: return 4;
  case MPI_INT64_T: return 8;
  }
  return 0;
}

static void shim_barrier(void)
{
  if (g_np == 1) return;
  g_local_sense = !g_local_sense;
  if (__atomic_add_fetch(&g_sh->bar_count, 1, __ATOMIC_ACQ_REL) == g_np)
  {
    g_sh->bar_count = 0;
    __atomic_store_n(&g_sh->bar_sense, g_local_sense, __ATOMIC_RELEASE);
  }
  else
  {
    while (__atomic_load_n(&g_sh->bar_sense, __ATOMIC_ACQUIRE) != g_local_sense)
    {
      sched_yield();
    }
  }
}

int MPI_Init(int* argc, char*** argv)
{
  (void)argc;
  (void)argv;
  const char* s = getenv("MPISHIM_NP");
  g_np          = s ? atoi(s) : 1;
  if (g_np < 1 || g_np > SHIM_MAX_RANKS)
  {
    fprintf(stderr, "mpi_shim: MPISHIM_NP must be in [1,%d]\n", SHIM_MAX_RANKS);
    exit(2);
  }
  size_t bytes = sizeof(shim_shared) +
                 (size_t)g_np * g_np * SHIM_NTAGS * sizeof(shim_slot);
  void* p = mmap(NULL, bytes, PROT_READ | PROT_WRITE,
                 MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED)
  {
    perror("mpi_shim: mmap");
    exit(2);
  }
  g_sh    = (shim_shared*)p;
  g_slots = (shim_slot*)((unsigned char*)p + sizeof(shim_shared));
  g_rank  = 0;
  fflush(NULL);
  for (int r = 1; r < g_np; r++)
  {
    pid_t pid = fork();
    if (pid < 0)
    {
      perror("mpi_shim: fork");
      exit(2);
    }
    if (pid == 0)
    {
      g_rank = r;
      break;
    }
    g_children[r] = pid;
  }
  return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
  fflush(NULL);
  shim_barrier();
  if (g_np > 1)
  {
    if (g_rank != 0) { _exit(0); }
    for (int r = 1; r < g_np; r++)
    {
      int st;
      waitpid(g_children[r], &st, 0);
    }
  }
  return MPI_SUCCESS;
}

double MPI_Wtime(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_size(MPI_Comm comm, int* size)
{
  (void)comm;
  *size = g_np;
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm comm, int* rank)
{
  (void)comm;
  *rank = g_rank;
  return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm* comm)
{
  *comm = MPI_COMM_NULL;
  return MPI_SUCCESS;
}

/* Balanced factorisation, factors in non-increasing order, entries that are
   already positive are kept (the MPI standard's MPI_Dims_create contract). */
int MPI_Dims_create(int nnodes, int ndims, int dims[])
{
  if (ndims != 2) return 1;
  if (dims[0] > 0 && dims[1] > 0) return (dims[0] * dims[1] == nnodes) ? 0 : 1;
  if (dims[0] > 0)
  {
    if (nnodes % dims[0]) return 1;
    dims[1] = nnodes / dims[0];
    return MPI_SUCCESS;
  }
  if (dims[1] > 0)
  {
    if (nnodes % dims[1]) return 1;
    dims[0] = nnodes / dims[1];
    return MPI_SUCCESS;
  }
  int b = 1;
  for (int f = 1; f * f <= nnodes; f++)
  {
    if (nnodes % f == 0) b = f;
  }
  dims[0] = nnodes / b;
  dims[1] = b;
  return MPI_SUCCESS;
}

int MPI_Cart_create(MPI_Comm comm, int ndims, const int dims[],
                    const int periods[], int reorder, MPI_Comm* comm_cart)
{
  (void)comm;
  (void)reorder;
  if (ndims != 2 || dims[0] * dims[1] != g_np) return 1;
  g_cart_dims[0]    = dims[0];
  g_cart_dims[1]    = dims[1];
  g_cart_periods[0] = periods[0];
  g_cart_periods[1] = periods[1];
  *comm_cart        = 2;
  return MPI_SUCCESS;
}

/* row-major rank order: rank = coords[0] * dims[1] + coords[1] */
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[],
                 int coords[])
{
  (void)comm;
  if (maxdims < 2) return 1;
  dims[0]    = g_cart_dims[0];
  dims[1]    = g_cart_dims[1];
  periods[0] = g_cart_periods[0];
  periods[1] = g_cart_periods[1];
  coords[0]  = g_rank / g_cart_dims[1];
  coords[1]  = g_rank % g_cart_dims[1];
  return MPI_SUCCESS;
}

int MPI_Cart_rank(MPI_Comm comm, const int coords[], int* rank)
{
  (void)comm;
  int c0 = coords[0], c1 = coords[1];
  c0     = ((c0 % g_cart_dims[0]) + g_cart_dims[0]) % g_cart_dims[0];
  c1     = ((c1 % g_cart_dims[1]) + g_cart_dims[1]) % g_cart_dims[1];
  *rank  = c0 * g_cart_dims[1] + c1;
  return MPI_SUCCESS;
}

int MPI_Cartdim_get(MPI_Comm comm, int* ndims)
{
  (void)comm;
  *ndims = 2;
  return MPI_SUCCESS;
}

int MPI_Topo_test(MPI_Comm comm, int* status)
{
  *status = (comm == 2) ? MPI_CART : MPI_UNDEFINED;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm)
{
  (void)comm;
  shim_barrier();
  return MPI_SUCCESS;
}

#define REDUCE_LOOP(T)                                               \
  do {                                                               \
    T* out = (T*)recvbuf;                                            \
    for (int i = 0; i < count; i++)                                  \
    {                                                                \
      T acc = ((T*)g_sh->red[0])[i];                                 \
      for (int r = 1; r < g_np; r++)                                 \
      {                                                              \
        T v = ((T*)g_sh->red[r])[i];                                 \
        if (op == MPI_SUM) acc += v;                                 \
        else if (op == MPI_MAX) acc = (v > acc) ? v : acc;           \
        else acc = (v < acc) ? v : acc;                              \
      }                                                              \
      out[i] = acc;                                                  \
    }                                                                \
  }                                                                  \
  while (0)

int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count,
                  MPI_Datatype datatype, MPI_Op op, MPI_Comm comm)
{
  (void)comm;
  size_t bytes = dtype_size(datatype) * (size_t)count;
  if (sendbuf == MPI_IN_PLACE) sendbuf = recvbuf;
  if (g_np == 1)
  {
    if (sendbuf != recvbuf) memcpy(recvbuf, sendbuf, bytes);
    return MPI_SUCCESS;
  }
  if (bytes > sizeof(double) * SHIM_RED_MAX) return 1;
  memcpy(g_sh->red[g_rank], sendbuf, bytes);
  shim_barrier();
  switch (datatype)
  {
  case MPI_DOUBLE: REDUCE_LOOP(double); break;
  case MPI_FLOAT: REDUCE_LOOP(float); break;
  case MPI_INT:
  case MPI_INT32_T: REDUCE_LOOP(int32_t); break;
  case MPI_INT64_T: REDUCE_LOOP(int64_t); break;
  default: return 1;
  }
  shim_barrier();
  return MPI_SUCCESS;
}

int MPI_Irecv(void* buf, int count, MPI_Datatype datatype, int source, int tag,
              MPI_Comm comm, MPI_Request* request)
{
  (void)comm;
  if (tag < 0 || tag >= SHIM_NTAGS) return 1;
  request->kind  = 2;
  request->buf   = buf;
  request->bytes = dtype_size(datatype) * (size_t)count;
  request->peer  = source;
  request->tag   = tag;
  if (request->bytes > SHIM_SLOT_BYTES) return 1;
  return MPI_SUCCESS;
}

int MPI_Isend(const void* buf, int count, MPI_Datatype datatype, int dest,
              int tag, MPI_Comm comm, MPI_Request* request)
{
  (void)comm;
  size_t bytes = dtype_size(datatype) * (size_t)count;
  if (tag < 0 || tag >= SHIM_NTAGS || bytes > SHIM_SLOT_BYTES) return 1;
  shim_slot* s = slot_of(dest, g_rank, tag);
  while (__atomic_load_n(&s->full, __ATOMIC_ACQUIRE)) { sched_yield(); }
  memcpy(s->data, buf, bytes);
  __atomic_store_n(&s->full, 1, __ATOMIC_RELEASE);
  request->kind = 1;
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request* request, MPI_Status* status)
{
  if (request->kind == 2)
  {
    shim_slot* s = slot_of(g_rank, request->peer, request->tag);
    while (!__atomic_load_n(&s->full, __ATOMIC_ACQUIRE)) { sched_yield(); }
    memcpy(request->buf, s->data, request->bytes);
    __atomic_store_n(&s->full, 0, __ATOMIC_RELEASE);
    if (status)
    {
      status->MPI_SOURCE = request->peer;
      status->MPI_TAG    = request->tag;
      status->MPI_ERROR  = MPI_SUCCESS;
    }
  }
  request->kind = 0;
  return MPI_SUCCESS;
}
