/*
 * nccl_dl.h -- NCCL bound at run time (dlopen) so that liblighter_b200.so has no link-time
 * dependency on a particular libnccl: inside a torch process it picks up the libnccl.so.2 torch
 * already loaded (2.28.x, the one that owns the NVLink/NVSwitch transport), in a plain C++ caller
 * it falls back to the system library.  Only the entry points the radiance exchange needs.
 */
#pragma once
#include <stddef.h>

struct NcclId { char internal[128]; };      /* ncclUniqueId: 128 opaque bytes, passed by value */

struct NcclApi {
    void *lib;
    int (*GetUniqueId)(NcclId *id);
    int (*CommInitRank)(void **comm, int nranks, NcclId id, int rank);
    int (*CommDestroy)(void *comm);
    int (*AllGather)(const void *send, void *recv, size_t count, int dtype, void *comm, void *stream);
    int (*Broadcast)(const void *send, void *recv, size_t count, int dtype, int root, void *comm, void *stream);
    int (*Send)(const void *send, size_t count, int dtype, int peer, void *comm, void *stream);
    int (*Recv)(void *recv, size_t count, int dtype, int peer, void *comm, void *stream);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char *(*GetErrorString)(int);
};

/* returns nullptr (and fills err) when no usable libnccl is found */
const NcclApi *nccl_api(char *err, size_t errlen);
