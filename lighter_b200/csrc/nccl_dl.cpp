/* nccl_dl.cpp -- see nccl_dl.h */
#include "nccl_dl.h"

#include <dlfcn.h>
#include <stdio.h>
#include <mutex>

static NcclApi g_api;
static bool g_ok = false;
static std::once_flag g_once;
static char g_err[256];

static void load_once()
{
    const char *names[] = { "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2" };
    void *lib = nullptr;
    /* RTLD_NOLOAD first: reuse the copy an embedding process (torch) has already mapped */
    for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (lib) break; }
    if (!lib) for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { snprintf(g_err, sizeof(g_err), "libnccl.so.2 not found: %s", dlerror()); return; }
    g_api.lib = lib;
    g_api.GetUniqueId = (int (*)(NcclId *))dlsym(lib, "ncclGetUniqueId");
    g_api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    g_api.CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
    g_api.AllGather = (int (*)(const void *, void *, size_t, int, void *, void *))dlsym(lib, "ncclAllGather");
    g_api.Broadcast = (int (*)(const void *, void *, size_t, int, int, void *, void *))dlsym(lib, "ncclBroadcast");
    g_api.Send = (int (*)(const void *, size_t, int, int, void *, void *))dlsym(lib, "ncclSend");
    g_api.Recv = (int (*)(void *, size_t, int, int, void *, void *))dlsym(lib, "ncclRecv");
    g_api.GroupStart = (int (*)(void))dlsym(lib, "ncclGroupStart");
    g_api.GroupEnd = (int (*)(void))dlsym(lib, "ncclGroupEnd");
    g_api.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_api.GetUniqueId || !g_api.CommInitRank || !g_api.CommDestroy || !g_api.AllGather || !g_api.Broadcast || !g_api.Send || !g_api.Recv || !g_api.GroupStart || !g_api.GroupEnd || !g_api.GetErrorString) {
        snprintf(g_err, sizeof(g_err), "libnccl is missing a required symbol");
        return;
    }
    g_ok = true;
}

const NcclApi *nccl_api(char *err, size_t errlen)
{
    std::call_once(g_once, load_once);
    if (!g_ok) { if (err && errlen) snprintf(err, errlen, "%s", g_err); return nullptr; }
    return &g_api;
}
