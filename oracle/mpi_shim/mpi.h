/*
 * mpi.h -- threads-as-ranks MPI substitute.  TEST INFRASTRUCTURE ONLY.
 *
 * This image has no MPI.  The reference CPU solver (solverPoissonMPI_CPU) is
 * compiled UNMODIFIED against this header so that it can serve as the parity
 * oracle and as the CPU baseline: every "rank" is a std::thread of one process,
 * the launcher (mpi_shim.cpp) starts px*py*pz of them and calls the reference's
 * main() (renamed with -Dmain=ref_main) on each.
 *
 * Only the MPI surface the reference actually uses is provided
 * (grep -oh "MPI_[A-Za-z_]*" over solverPoissonMPI_CPU/):
 *   Init Finalize Comm_size Comm_rank Barrier Allreduce Reduce Bcast
 *   Isend Irecv Waitall Type_indexed Type_commit
 * plus the four MPI-IO calls of the alpaka tree's main.cpp (solverPoissonMPI_alpaka/src/main.cpp:137-145:
 * File_open / File_seek / File_write / File_close with an individual file pointer per rank).
 * Semantics: sends are eager and buffered (the reference never waits on its
 * send requests, communicationMPI.hpp:306-316), matching is FIFO per (src,dst),
 * reductions sum in rank order 0..n-1 on every rank (deterministic).
 */
#ifndef PPS_ORACLE_MPI_SHIM_H
#define PPS_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

typedef struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
} MPI_Status;

typedef struct MPI_Request {
    int kind;            /* 0 = null/complete, 1 = pending receive */
    void* buf;
    int count;
    MPI_Datatype type;
    int peer;
} MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_DATATYPE_NULL 0
#define MPI_DOUBLE 1
#define MPI_FLOAT 2
#define MPI_INT 3
#define MPI_SUM 1

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Irecv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Waitall(int count, MPI_Request* reqs, MPI_Status* statuses);
int MPI_Type_indexed(int count, const int* blocklens, const int* displs, MPI_Datatype oldtype, MPI_Datatype* newtype);
int MPI_Type_commit(MPI_Datatype* type);
int MPI_Type_free(MPI_Datatype* type);

/* MPI-IO subset: every rank owns a descriptor and an individual file pointer (positional writes) */
typedef struct pps_shim_file* MPI_File;
typedef long long MPI_Offset;
typedef int MPI_Info;
#define MPI_INFO_NULL 0
#define MPI_MODE_CREATE 1
#define MPI_MODE_WRONLY 4
#define MPI_MODE_RDONLY 2
#define MPI_SEEK_SET 600
int MPI_File_open(MPI_Comm comm, const char* filename, int amode, MPI_Info info, MPI_File* fh);
int MPI_File_seek(MPI_File fh, MPI_Offset offset, int whence);
int MPI_File_write(MPI_File fh, const void* buf, int count, MPI_Datatype type, MPI_Status* status);
int MPI_File_close(MPI_File* fh);

/* launcher side (not part of MPI): run `fn(argc, argv)` on `world` rank-threads */
int pps_shim_run(int world, int (*fn)(int, char**), int argc, char** argv);
int pps_shim_rank(void);
int pps_shim_world(void);

#ifdef __cplusplus
}
#endif
#endif
