// The one exchange step of the render path: the sum of the per-GPU accumulation images (include/cornelis_cuda.h,
// "multi-GPU").  The reference's estimator is a plain mean over samples (reference src/Render.cpp:245-250), so shards of
// the sample range combine by ONE fp32 sum of width*height*4 floats per GPU: ncclAllReduce / ncclReduce over NVLink.
//
// NCCL is bound at run time (dlopen of libnccl.so.2), not at link time: a process that already carries a copy — torch
// ships its own — must not get a second one, and a single-GPU user of the library (the CLI, the stage entry points)
// needs none.  Only types and enums come from <nccl.h>.  A missing library is an error (CORNELIS_ERR_NCCL).
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "scene_access.h"

using namespace cornelis_b200;

namespace {

struct NcclApi {
    void *handle = nullptr;
    std::string origin;
    ncclResult_t (*getVersion)(int *) = nullptr;
    ncclResult_t (*getUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*commInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*commInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*allReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*groupStart)() = nullptr;
    ncclResult_t (*groupEnd)() = nullptr;
    const char *(*getErrorString)(ncclResult_t) = nullptr;
};

// Loads NCCL once per process.  Order: CORNELIS_NCCL_LIB (explicit path), the libnccl.so.2 already mapped into the
// process (RTLD_NOLOAD), then the loader's search path.
const NcclApi *nccl(std::string &why) {
    static std::mutex mutex;
    static NcclApi api;
    static bool tried = false;
    static std::string error;
    std::lock_guard<std::mutex> lock(mutex);
    if (!tried) {
        tried = true;
        void *h = nullptr;
        if (const char *env = std::getenv("CORNELIS_NCCL_LIB")) {
            h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
            api.origin = env;
        }
        if (!h) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
            api.origin = "libnccl.so.2 (already loaded in this process)";
        }
        if (!h) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
            api.origin = "libnccl.so.2";
        }
        if (!h) {
            const char *e = dlerror();
            error = std::string("cannot load libnccl.so.2: ") + (e ? e : "unknown error") +
                    " (multi-GPU rendering needs NCCL; there is no other exchange path)";
        } else {
            api.handle = h;
            bool ok = true;
            auto bind = [&](auto &fn, const char *name) {
                fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(h, name));
                if (!fn) {
                    ok = false;
                    error = std::string("libnccl.so.2 lacks ") + name;
                }
            };
            bind(api.getVersion, "ncclGetVersion");
            bind(api.getUniqueId, "ncclGetUniqueId");
            bind(api.commInitRank, "ncclCommInitRank");
            bind(api.commInitAll, "ncclCommInitAll");
            bind(api.commDestroy, "ncclCommDestroy");
            bind(api.allReduce, "ncclAllReduce");
            bind(api.reduce, "ncclReduce");
            bind(api.groupStart, "ncclGroupStart");
            bind(api.groupEnd, "ncclGroupEnd");
            bind(api.getErrorString, "ncclGetErrorString");
            if (!ok)
                api.handle = nullptr;
        }
    }
    if (!api.handle) {
        why = error;
        return nullptr;
    }
    return &api;
}

#define CB_NCCL(api, expr)                                                                                             \
    do {                                                                                                               \
        ncclResult_t r_ = (expr);                                                                                      \
        if (r_ != ncclSuccess)                                                                                         \
            return failWith(CORNELIS_ERR_NCCL, std::string(#expr) + ": " + (api)->getErrorString(r_));                \
    } while (0)

#define CB_CUDA(expr)                                                                                                  \
    do {                                                                                                               \
        cudaError_t e_ = (expr);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return failWith(CORNELIS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                   \
    } while (0)

static_assert(sizeof(ncclUniqueId) == CORNELIS_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");

} // namespace

struct cornelis_cuda_comm {
    const NcclApi *api = nullptr;
    std::vector<ncclComm_t> comms; // one per local rank
    std::vector<int> devices;      // device of each local rank
    int nRanks = 0;
    ~cornelis_cuda_comm() {
        int current = 0;
        cudaGetDevice(&current);
        for (size_t k = 0; k < comms.size(); k++)
            if (comms[k]) {
                cudaSetDevice(devices[k]);
                api->commDestroy(comms[k]);
            }
        cudaSetDevice(current);
    }
};

namespace {

// The frames of `n` scenes, checked to be renders of one frame size with the same moments.
int gatherFrames(cornelis_cuda_scene *const *scenes, int n, std::vector<FrameView> &frames) {
    if (!scenes || n <= 0)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "no scenes to reduce");
    frames.resize(static_cast<size_t>(n));
    for (int k = 0; k < n; k++) {
        if (!frameView(scenes[k], frames[k]))
            return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "nothing has been rendered yet on one of the scenes");
        if (frames[k].npixels != frames[0].npixels || frames[k].haveVariance != frames[0].haveVariance)
            return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "scenes must hold renders of the same frame");
    }
    return CORNELIS_OK;
}

// Process-wide communicators of cornelis_cuda_reduce_framebuffers, one per device set, created on first use.
std::mutex g_sharedMutex;
std::map<std::vector<int>, cornelis_cuda_comm *> g_shared;

} // namespace

extern "C" {

int cornelis_cuda_comm_unique_id(uint8_t id[CORNELIS_COMM_ID_BYTES]) {
    if (!id)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "id is null");
    std::string why;
    const NcclApi *api = nccl(why);
    if (!api)
        return failWith(CORNELIS_ERR_NCCL, why);
    ncclUniqueId u;
    CB_NCCL(api, api->getUniqueId(&u));
    std::memcpy(id, &u, sizeof u);
    return CORNELIS_OK;
}

int cornelis_cuda_comm_init_rank(const uint8_t id[CORNELIS_COMM_ID_BYTES], int rank, int n_ranks, int device,
                                 cornelis_cuda_comm **out) {
    if (!out)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "out_comm is null");
    *out = nullptr;
    if (!id || n_ranks <= 0 || rank < 0 || rank >= n_ranks || device < 0)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "bad rank / n_ranks / device");
    std::string why;
    const NcclApi *api = nccl(why);
    if (!api)
        return failWith(CORNELIS_ERR_NCCL, why);
    CB_CUDA(cudaSetDevice(device));
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclComm_t c = nullptr;
    CB_NCCL(api, api->commInitRank(&c, n_ranks, u, rank));
    auto *comm = new cornelis_cuda_comm;
    comm->api = api;
    comm->comms.push_back(c);
    comm->devices.push_back(device);
    comm->nRanks = n_ranks;
    *out = comm;
    return CORNELIS_OK;
}

int cornelis_cuda_comm_init_all(const int *devices, int n, cornelis_cuda_comm **out) {
    if (!out)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "out_comm is null");
    *out = nullptr;
    if (!devices || n <= 0)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "no devices");
    int count = 0;
    if (int rc = cornelis_cuda_device_count(&count))
        return rc;
    for (int a = 0; a < n; a++) {
        if (devices[a] < 0 || devices[a] >= count)
            return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "device index out of range");
        for (int b = 0; b < a; b++)
            if (devices[a] == devices[b])
                return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "a communicator takes every device once");
    }
    std::string why;
    const NcclApi *api = nccl(why);
    if (!api)
        return failWith(CORNELIS_ERR_NCCL, why);
    std::vector<ncclComm_t> comms(static_cast<size_t>(n), nullptr);
    CB_NCCL(api, api->commInitAll(comms.data(), n, devices));
    auto *comm = new cornelis_cuda_comm;
    comm->api = api;
    comm->comms = comms;
    comm->devices.assign(devices, devices + n);
    comm->nRanks = n;
    *out = comm;
    return CORNELIS_OK;
}

int cornelis_cuda_comm_destroy(cornelis_cuda_comm *comm) {
    delete comm;
    return CORNELIS_OK;
}

int cornelis_cuda_comm_info(cornelis_cuda_comm *comm, int *n_ranks, int *n_local, int *nccl_version) {
    if (!comm)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "comm is null");
    if (n_ranks)
        *n_ranks = comm->nRanks;
    if (n_local)
        *n_local = static_cast<int>(comm->comms.size());
    if (nccl_version)
        CB_NCCL(comm->api, comm->api->getVersion(nccl_version));
    return CORNELIS_OK;
}

int cornelis_cuda_allreduce_framebuffers(cornelis_cuda_comm *comm, cornelis_cuda_scene *const *scenes, int n_local) {
    if (!comm)
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "comm is null");
    if (n_local != static_cast<int>(comm->comms.size()))
        return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "one scene per local rank of the communicator is needed");
    std::vector<FrameView> frames;
    if (int rc = gatherFrames(scenes, n_local, frames))
        return rc;
    for (int k = 0; k < n_local; k++)
        if (frames[k].device != comm->devices[k])
            return failWith(CORNELIS_ERR_INVALID_ARGUMENT, "scenes must follow the communicator's device order");
    const NcclApi *api = comm->api;
    size_t const count = frames[0].npixels * 4;
    int current = 0;
    cudaGetDevice(&current);
    CB_NCCL(api, api->groupStart());
    for (int k = 0; k < n_local; k++) {
        FrameView const &f = frames[k];
        CB_NCCL(api, api->allReduce(f.accum, f.accum, count, ncclFloat32, ncclSum, comm->comms[k], f.stream));
        if (f.accum2)
            CB_NCCL(api, api->allReduce(f.accum2, f.accum2, count, ncclFloat32, ncclSum, comm->comms[k], f.stream));
    }
    CB_NCCL(api, api->groupEnd());
    cudaSetDevice(current);
    return CORNELIS_OK;
}

int cornelis_cuda_reduce_framebuffers(cornelis_cuda_scene *const *scenes, int n) {
    std::vector<FrameView> frames;
    if (int rc = gatherFrames(scenes, n, frames))
        return rc;
    int current = 0;
    cudaGetDevice(&current);
    // Scenes that share a GPU with an earlier one are added to it on that GPU: nothing to exchange.
    std::vector<int> leaders; // index of the first scene on each distinct device, scenes[0] first
    for (int k = 0; k < n; k++) {
        int leader = -1;
        for (int l : leaders)
            if (frames[l].device == frames[k].device)
                leader = l;
        if (leader < 0) {
            leaders.push_back(k);
            continue;
        }
        FrameView const &dst = frames[leader], &src = frames[k];
        CB_CUDA(cudaSetDevice(dst.device));
        CB_CUDA(cudaStreamSynchronize(src.stream));
        launchAddImages(dst.stream, *dst.shape, dst.npixels, dst.accum, src.accum);
        if (dst.accum2)
            launchAddImages(dst.stream, *dst.shape, dst.npixels, dst.accum2, src.accum2);
        CB_CUDA(cudaGetLastError());
    }
    if (leaders.size() > 1) {
        std::vector<int> devices;
        for (int l : leaders)
            devices.push_back(frames[l].device);
        cornelis_cuda_comm *comm = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_sharedMutex);
            auto it = g_shared.find(devices);
            if (it == g_shared.end()) {
                if (int rc = cornelis_cuda_comm_init_all(devices.data(), static_cast<int>(devices.size()), &comm))
                    return rc;
                g_shared[devices] = comm; // kept for the life of the process
            } else {
                comm = it->second;
            }
        }
        const NcclApi *api = comm->api;
        size_t const count = frames[0].npixels * 4;
        CB_NCCL(api, api->groupStart());
        for (size_t r = 0; r < leaders.size(); r++) {
            FrameView const &f = frames[leaders[r]];
            CB_NCCL(api, api->reduce(f.accum, f.accum, count, ncclFloat32, ncclSum, 0, comm->comms[r], f.stream));
            if (f.accum2)
                CB_NCCL(api, api->reduce(f.accum2, f.accum2, count, ncclFloat32, ncclSum, 0, comm->comms[r], f.stream));
        }
        CB_NCCL(api, api->groupEnd());
    }
    for (int l : leaders) { // the call is synchronous: every participant's part of the reduction has completed
        CB_CUDA(cudaSetDevice(frames[l].device));
        CB_CUDA(cudaStreamSynchronize(frames[l].stream));
    }
    CB_CUDA(cudaSetDevice(current));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

} // extern "C"
