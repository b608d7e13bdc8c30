// pb_nccl.cpp - the library's NCCL communicator (one per process = one per GPU), resolved at run time.
//
// Image-sharded runs (patolette_b200_sharded) exchange on the DEVICE: the colour planes are
// all-gathered over NVLink after every rank transformed its slice, and every batch of the split loop all-gathers
// the children's segment descriptors and statistics - on the library's own stream, no host round trip besides
// the one synchronisation per batch the single-GPU path already has.  NCCL is dlopen'ed (libnccl.so.2: the copy
// torch already loaded when the caller is a torch process, else the system one), so the library has no link-time
// dependency on it and single-GPU users never touch it.
#include "pb_nccl.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "pb_error.h"

namespace {

struct Api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
} g_api;
std::mutex g_mu;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

bool load_locked() {
    if (g_api.handle) return true;
    const char *env = getenv("PATOLETTE_B200_NCCL");
    const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        fprintf(stderr, "patolette_b200: libnccl.so.2 not found (set PATOLETTE_B200_NCCL): %s\n", dlerror());
        return false;
    }
    Api a;
    a.handle = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.GetErrorString) {
        fprintf(stderr, "patolette_b200: the NCCL library lacks a required symbol\n");
        dlclose(h);
        return false;
    }
    g_api = a;
    return true;
}

} // namespace

int pb_nccl_unique_id(char out[128]) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!load_locked()) return -1;
    ncclUniqueId id;
    const ncclResult_t r = g_api.GetUniqueId(&id);
    if (r != ncclSuccess) {
        fprintf(stderr, "patolette_b200: ncclGetUniqueId: %s\n", g_api.GetErrorString(r));
        return -2;
    }
    static_assert(sizeof id == 128, "ncclUniqueId is 128 bytes");
    memcpy(out, &id, sizeof id);
    return 0;
}

int pb_nccl_init(int rank, int world, const char id_bytes[128]) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (world < 1 || rank < 0 || rank >= world) return -1;
    if (g_comm) {
        g_api.CommDestroy(g_comm);
        g_comm = nullptr;
        g_rank = 0;
        g_world = 1;
    }
    if (world == 1) return 0;
    if (!id_bytes || !load_locked()) return -1;
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    const ncclResult_t r = g_api.CommInitRank(&g_comm, world, id, rank);
    if (r != ncclSuccess) {
        fprintf(stderr, "patolette_b200: ncclCommInitRank(rank %d of %d): %s\n", rank, world, g_api.GetErrorString(r));
        g_comm = nullptr;
        return -2;
    }
    g_rank = rank;
    g_world = world;
    return 0;
}

void pb_nccl_destroy() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_comm) g_api.CommDestroy(g_comm);
    g_comm = nullptr;
    g_rank = 0;
    g_world = 1;
}

bool pb_nccl_active() { return g_comm != nullptr && g_world > 1; }
int pb_nccl_rank() { return g_rank; }
int pb_nccl_world() { return g_world; }

int pb_nccl_version() {
    std::lock_guard<std::mutex> lk(g_mu);
    int v = 0;
    if (!load_locked() || !g_api.GetVersion || g_api.GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

void pb_nccl_allgather(const void *send, void *recv, size_t bytes, cudaStream_t st) {
    if (!g_comm) throw pb_cuda_error(cudaErrorNotReady);
    const ncclResult_t r = g_api.AllGather(send, recv, bytes, ncclChar, g_comm, st);
    if (r != ncclSuccess) {
        fprintf(stderr, "patolette_b200: ncclAllGather(%zu bytes): %s\n", bytes, g_api.GetErrorString(r));
        throw pb_cuda_error(cudaErrorUnknown);
    }
}
