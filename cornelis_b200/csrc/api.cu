// C-ABI of the render path (include/cornelis_cuda.h): scene upload, wavefront driver, stage entry points.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cornelis_cuda.h"
#include "device_types.h"
#include "materials.cuh"
#include "math.cuh"
#include "scene_access.h"
#include "scene_tables.h"
#include "wavefront.h"

using namespace cornelis_b200;

namespace {

thread_local std::string g_lastError;

int fail(cornelis_status code, const std::string &message) {
    g_lastError = message;
    return code;
}

#define CB_CUDA(expr)                                                                                                  \
    do {                                                                                                               \
        cudaError_t e_ = (expr);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(e_ == cudaErrorMemoryAllocation ? CORNELIS_ERR_OUT_OF_MEMORY : CORNELIS_ERR_CUDA,               \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                                          \
    } while (0)

// Device-memory cache.  A render service creates and destroys scene handles all the time (bench.py's end-to-end leg
// does so every step), and cudaFree / cudaMalloc of the framebuffer-sized blocks cost up to 400 ms per call on the
// B200 boxes (measured: profiles/r1_queue/e2e_breakdown.txt) — a third of a 1080p, 4096-spp render.  Blocks released
// by a handle are therefore kept per device, keyed by size, and handed to the next request of a similar size;
// cornelis_cuda_trim_memory() returns them to the driver.  At most kCacheLimitBytes stay cached per process.
class DeviceCache {
  public:
    cudaError_t take(void **ptr, size_t bytes) {
        int device = 0;
        cudaGetDevice(&device);
        size_t const rounded = roundUp(bytes);
        {
            std::lock_guard<std::mutex> lock(mutex_);
            auto &blocks = free_[device];
            auto it = blocks.lower_bound(rounded);
            if (it != blocks.end() && it->first <= rounded + rounded / 4) { // at most 25 % larger than asked for
                *ptr = it->second;
                cached_ -= it->first;
                sizes_[*ptr] = it->first;
                blocks.erase(it);
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(ptr, rounded);
        if (e == cudaErrorMemoryAllocation) { // give the cached blocks back and try once more
            cudaGetLastError();
            trim();
            e = cudaMalloc(ptr, rounded);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> lock(mutex_);
            sizes_[*ptr] = rounded;
        }
        return e;
    }
    // The caller has made sure no work that uses the block is still in flight.
    void give(void *ptr) {
        int device = 0;
        cudaGetDevice(&device);
        size_t bytes = 0;
        {
            std::lock_guard<std::mutex> lock(mutex_);
            auto it = sizes_.find(ptr);
            if (it != sizes_.end()) {
                bytes = it->second;
                sizes_.erase(it);
                if (cached_ + bytes <= kCacheLimitBytes) {
                    free_[device].emplace(bytes, ptr);
                    cached_ += bytes;
                    return;
                }
            }
        }
        cudaFree(ptr);
    }
    void trim() {
        std::map<int, std::multimap<size_t, void *>> blocks;
        {
            std::lock_guard<std::mutex> lock(mutex_);
            blocks.swap(free_);
            cached_ = 0;
        }
        int current = 0;
        cudaGetDevice(&current);
        for (auto &perDevice : blocks) {
            cudaSetDevice(perDevice.first);
            for (auto &b : perDevice.second)
                cudaFree(b.second);
        }
        cudaSetDevice(current);
    }

  private:
    static size_t roundUp(size_t bytes) { // 512-byte granules below 1 MiB, 1 MiB granules above
        size_t const g = bytes < (1u << 20) ? 512u : (1u << 20);
        return (std::max<size_t>(bytes, 1) + g - 1) / g * g;
    }
    // (16 GiB: the wavefront pipeline's 2^26-path pools and queues, 11 GB, are recycled between handles too)
    static constexpr size_t kCacheLimitBytes = size_t(16) << 30;
    std::mutex mutex_;
    std::map<int, std::multimap<size_t, void *>> free_; // device -> size -> block
    std::map<void *, size_t> sizes_;                    // blocks handed out
    size_t cached_ = 0;
};

DeviceCache &deviceCache() {
    static DeviceCache *cache = new DeviceCache; // never destroyed: handles may outlive static destruction order
    return *cache;
}

template <typename T>
struct DeviceBuffer {
    T *ptr = nullptr;
    size_t count = 0;
    cudaError_t reserve(size_t n) {
        if (n <= count)
            return cudaSuccess;
        release();
        cudaError_t e = deviceCache().take(reinterpret_cast<void **>(&ptr), n * sizeof(T));
        if (e == cudaSuccess)
            count = n;
        else
            ptr = nullptr;
        return e;
    }
    void release() {
        if (ptr) {
            cudaDeviceSynchronize(); // what cudaFree did implicitly: nothing in flight may still use the block
            deviceCache().give(ptr);
        }
        ptr = nullptr;
        count = 0;
    }
};

// PerspectiveCamera::lookAt, reference src/Camera.cpp:15-34, over nanovdb::Vec3<float> semantics
// (NanoVDB.h:911-953): cross = (a1 b2 - a2 b1, ...), normalize = multiply by the rounded reciprocal length,
// `0.5 * u` evaluated in double (exact halving).
DevCamera makeCamera(const cornelis_camera_desc &c) {
    V3 const from{c.origin[0], c.origin[1], c.origin[2]}, at{c.look_at[0], c.look_at[1], c.look_at[2]};
    V3 const up{0.0f, 1.0f, 0.0f};
    V3 dir = at - from;
    {
        float s = 1.0f / sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
        dir = V3{dir.x * s, dir.y * s, dir.z * s};
    }
    V3 u = cross(up, dir);
    V3 v = cross(u, dir);
    float const fovScale = static_cast<float>(2.0 * std::sin(c.horizontal_fov * 0.5));
    u = V3{u.x * fovScale, u.y * fovScale, u.z * fovScale};
    float const vs = c.aspect * fovScale;
    v = V3{v.x * vs, v.y * vs, v.z * vs};
    V3 const hu{static_cast<float>(0.5 * u.x), static_cast<float>(0.5 * u.y), static_cast<float>(0.5 * u.z)};
    V3 const hv{static_cast<float>(0.5 * v.x), static_cast<float>(0.5 * v.y), static_cast<float>(0.5 * v.z)};
    V3 const corner = (dir - hu) - hv;
    DevCamera cam{};
    cam.ex = from.x, cam.ey = from.y, cam.ez = from.z;
    cam.cx = corner.x, cam.cy = corner.y, cam.cz = corner.z;
    cam.ux = u.x, cam.uy = u.y, cam.uz = u.z;
    cam.vx = v.x, cam.vy = v.y, cam.vz = v.z;
    return cam;
}

constexpr size_t kAutoGridSpheres = 128; // CORNELIS_ACCEL_AUTO: scenes with at least this many spheres use the grid

} // namespace

struct cornelis_cuda_scene {
    int device = 0;
    cudaStream_t stream = nullptr;    // the stream in use
    cudaStream_t ownStream = nullptr; // created with the scene
    LaunchShape shape;
    // scene tables
    DeviceBuffer<DevSphere> spheres;
    DeviceBuffer<uint32_t> sphereMaterial;
    DeviceBuffer<DevPlane> planes;
    DeviceBuffer<DevMaterial> materials;
    DeviceBuffer<uint32_t> planeOrder;
    DeviceBuffer<DevAxisPlane> axisPlanes;
    SceneView view{};
    // acceleration structure (built on first use)
    std::vector<cornelis_sphere_desc> hostSpheres;
    double boxMin[3] = {0, 0, 0}, boxMax[3] = {0, 0, 0}; // spheres, plane rectangles and the camera eye
    int accelMode = CORNELIS_ACCEL_AUTO;
    bool gridBuilt = false, gridUsable = false;
    DevGrid grid{};
    uint64_t gridItems = 0;
    DeviceBuffer<uint2> gridCellRange;
    DeviceBuffer<float4> gridCellSpheres;
    DeviceBuffer<uint32_t> gridCellIds;
    int smemOptin = 0;
    // wavefront state
    uint32_t width = 0, height = 0;
    DeviceBuffer<float4> pool[2][4];
    DeviceBuffer<HitRecord> hits;
    DeviceBuffer<uint32_t> hitQueue;
    DeviceBuffer<FinishedPath> finished;
    DeviceBuffer<Control> control;
    DeviceBuffer<unsigned long long> claimCursor; // persistent pipeline: next camera path to claim
    int gridPersistent = 0;
    Control hostControlStorage{};
    Control *hostControl = &hostControlStorage; // 104 bytes: pageable is fine
    DeviceBuffer<float4> accum, accum2;
    bool haveImage = false;    // accum holds the sum of a render of this frame size (what CORNELIS_RENDER_KEEP adds to)
    bool haveVariance = false; // ... and accum2 the second moments of the same samples
    DeviceBuffer<float> outRgb, outVar;
    DeviceBuffer<uint8_t> outRgb8;
    // staging for the stage entry points
    DeviceBuffer<float> stageF[12];
    DeviceBuffer<int32_t> stageI[3];
    DeviceBuffer<float4> stage4[2];
    DeviceBuffer<HitRecord> stageHits;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    cudaEvent_t evStage[6] = {};

    ~cornelis_cuda_scene() {
        cudaSetDevice(device);
        if (stream)
            cudaStreamSynchronize(stream);
        spheres.release(), sphereMaterial.release(), planes.release(), materials.release();
        gridCellRange.release(), gridCellSpheres.release(), gridCellIds.release(), planeOrder.release(), axisPlanes.release();
        for (auto &half : pool)
            for (auto &b : half)
                b.release();
        hits.release(), hitQueue.release(), finished.release(), control.release(), claimCursor.release();
        accum.release(), accum2.release(), outRgb.release(), outVar.release(), outRgb8.release();
        for (auto &b : stageF)
            b.release();
        for (auto &b : stageI)
            b.release();
        for (auto &b : stage4)
            b.release();
        stageHits.release();
        if (evStart)
            cudaEventDestroy(evStart);
        if (evStop)
            cudaEventDestroy(evStop);
        for (auto &e : evStage)
            if (e)
                cudaEventDestroy(e);
        if (ownStream)
            cudaStreamDestroy(ownStream);
    }
};

namespace {

PathPool poolView(cornelis_cuda_scene *s, int which) {
    return PathPool{s->pool[which][0].ptr, s->pool[which][1].ptr, s->pool[which][2].ptr, s->pool[which][3].ptr};
}

// device memory per pooled path: two pools of four float4, a hit record, a hit-queue slot, two finished-path records
constexpr size_t kPoolBytesPerPath = 2 * 4 * sizeof(float4) + sizeof(HitRecord) + sizeof(uint32_t) + 2 * sizeof(FinishedPath);

int ensureFrame(cornelis_cuda_scene *s, uint32_t width, uint32_t height, uint32_t poolPaths, bool variance) {
    size_t const npix = static_cast<size_t>(width) * height;
    if (poolPaths) { // the wavefront pipeline's path pool and queues
        for (int half = 0; half < 2; half++)
            for (int a = 0; a < 4; a++)
                CB_CUDA(s->pool[half][a].reserve(poolPaths));
        CB_CUDA(s->hits.reserve(poolPaths));
        CB_CUDA(s->hitQueue.reserve(poolPaths));
        CB_CUDA(s->finished.reserve(2 * static_cast<size_t>(poolPaths)));
    }
    CB_CUDA(s->control.reserve(1));
    CB_CUDA(s->claimCursor.reserve(1));
    bool const resized = s->width != width || s->height != height;
    if (resized) {
        s->accum.release();
        s->accum2.release();
        s->haveImage = false;
        s->haveVariance = false;
    }
    // a block handed out by the device cache holds whatever its previous owner left in it: whenever the accumulators
    // are (re)allocated there is no image to keep
    float4 const *const before = s->accum.ptr;
    CB_CUDA(s->accum.reserve(npix));
    if (s->accum.ptr != before)
        s->haveImage = false;
    if (variance) {
        float4 const *const before2 = s->accum2.ptr;
        CB_CUDA(s->accum2.reserve(npix));
        if (s->accum2.ptr != before2)
            s->haveVariance = false;
    }
    s->width = width;
    s->height = height;
    return CORNELIS_OK;
}

int checkScene(cornelis_cuda_scene *s) {
    if (!s)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "scene handle is null");
    CB_CUDA(cudaSetDevice(s->device));
    return CORNELIS_OK;
}

// Chooses between the exhaustive scan (all tables staged in shared memory) and the uniform grid (spheres stay in
// global memory), builds and uploads the grid on first use, and re-derives the launch shapes.
int applyAcceleration(cornelis_cuda_scene *s, int mode) {
    size_t const nS = s->view.nSpheres, nP = s->view.nPlanes, nM = s->view.nMaterials;
    auto tableBytes = [&](bool spheresInShared) {
        return sharedSceneBytes(static_cast<uint32_t>(nS), static_cast<uint32_t>(nP), s->view.planeEnd[2],
                                static_cast<uint32_t>(nM), spheresInShared);
    };
    size_t const limit = static_cast<size_t>(s->smemOptin) - 1024;
    bool wantGrid = nS > 0 && (mode == CORNELIS_ACCEL_GRID || (mode == CORNELIS_ACCEL_AUTO && nS >= kAutoGridSpheres));
    if (mode == CORNELIS_ACCEL_AUTO && !wantGrid && nS > 0 && tableBytes(true) > limit)
        wantGrid = true;
    if (wantGrid && !s->gridBuilt) {
        HostGrid h;
        s->gridUsable = buildGrid(s->hostSpheres.data(), s->hostSpheres.size(), s->boxMin, s->boxMax, h);
        s->gridBuilt = true;
        if (s->gridUsable) {
            size_t const refs = h.cellIds.size();
            CB_CUDA(s->gridCellRange.reserve(h.cellRange.size()));
            CB_CUDA(s->gridCellSpheres.reserve(refs ? refs : 1));
            CB_CUDA(s->gridCellIds.reserve(refs ? refs : 1));
            CB_CUDA(cudaMemcpyAsync(s->gridCellRange.ptr, h.cellRange.data(), h.cellRange.size() * sizeof(uint2),
                                    cudaMemcpyHostToDevice, s->stream));
            if (refs) {
                CB_CUDA(cudaMemcpyAsync(s->gridCellSpheres.ptr, h.cellSpheres.data(), refs * sizeof(float4),
                                        cudaMemcpyHostToDevice, s->stream));
                CB_CUDA(cudaMemcpyAsync(s->gridCellIds.ptr, h.cellIds.data(), refs * sizeof(uint32_t),
                                        cudaMemcpyHostToDevice, s->stream));
            }
            CB_CUDA(cudaStreamSynchronize(s->stream));
            s->grid = h.g;
            s->grid.cellRange = s->gridCellRange.ptr;
            s->grid.cellSpheres = s->gridCellSpheres.ptr;
            s->grid.cellIds = s->gridCellIds.ptr;
            s->gridItems = refs;
        }
    }
    if (wantGrid && !s->gridUsable) {
        if (mode == CORNELIS_ACCEL_GRID)
            return fail(CORNELIS_ERR_INVALID_ARGUMENT, "the grid cannot be built for this scene (degenerate or non-finite bounds)");
        wantGrid = false;
    }
    size_t const smem = tableBytes(!wantGrid);
    if (smem > limit)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT,
                    "scene tables exceed the shared-memory staging limit of this build (" + std::to_string(smem) +
                        " bytes)");
    DevGrid off{};
    s->view.grid = wantGrid ? s->grid : off;
    s->accelMode = mode;
    s->shape.sceneSmemBytes = smem;
    CB_CUDA(configureKernels(s->shape));
    CB_CUDA(configurePersistent(s->shape, wantGrid, s->view.nSpheres, s->gridPersistent));
    return CORNELIS_OK;
}

} // namespace

namespace cornelis_b200 {

bool frameView(cornelis_cuda_scene *s, FrameView &out) {
    if (!s || !s->accum.ptr || !s->width || !s->haveImage)
        return false;
    out.device = s->device;
    out.stream = s->stream;
    out.accum = s->accum.ptr;
    out.accum2 = s->haveVariance ? s->accum2.ptr : nullptr;
    out.npixels = static_cast<size_t>(s->width) * s->height;
    out.haveVariance = s->haveVariance;
    out.shape = &s->shape;
    return true;
}

int failWith(cornelis_status code, const std::string &message) { return fail(code, message); }

} // namespace cornelis_b200

extern "C" {

int cornelis_cuda_abi_version(void) { return CORNELIS_CUDA_ABI_VERSION; }

int cornelis_cuda_trim_memory(void) {
    deviceCache().trim();
    return CORNELIS_OK;
}

const char *cornelis_cuda_last_error(void) { return g_lastError.c_str(); }

int cornelis_cuda_device_count(int *count) {
    if (!count)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        *count = 0;
        return fail(CORNELIS_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                                " (this library has no CPU path)");
    }
    *count = n;
    return CORNELIS_OK;
}

int cornelis_cuda_scene_create(int device, const cornelis_camera_desc *camera, const cornelis_sphere_desc *spheres,
                               size_t n_spheres, const cornelis_plane_desc *planes, size_t n_planes,
                               const cornelis_material_desc *materials, size_t n_materials,
                               cornelis_cuda_scene **out_scene) {
    if (!out_scene)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "out_scene is null");
    *out_scene = nullptr;
    if (!camera || (n_spheres && !spheres) || (n_planes && !planes) || !materials || n_materials == 0)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "scene arrays missing (at least the default material is required)");
    int count = 0;
    if (int rc = cornelis_cuda_device_count(&count))
        return rc;
    if (device < 0 || device >= count)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "device index out of range");
    for (size_t i = 0; i < n_spheres; i++)
        if (spheres[i].material >= static_cast<int32_t>(n_materials))
            return fail(CORNELIS_ERR_INVALID_ARGUMENT, "sphere material index out of range");
    for (size_t i = 0; i < n_planes; i++)
        if (planes[i].material >= static_cast<int32_t>(n_materials))
            return fail(CORNELIS_ERR_INVALID_ARGUMENT, "plane material index out of range");

    CB_CUDA(cudaSetDevice(device));
    cornelis_cuda_scene *s = new (std::nothrow) cornelis_cuda_scene();
    if (!s)
        return fail(CORNELIS_ERR_OUT_OF_MEMORY, "host allocation failed");
    s->device = device;
    struct Guard {
        cornelis_cuda_scene *p;
        ~Guard() { delete p; }
    } guard{s};

    CB_CUDA(cudaStreamCreateWithFlags(&s->ownStream, cudaStreamNonBlocking));
    s->stream = s->ownStream;
    CB_CUDA(cudaEventCreate(&s->evStart));
    CB_CUDA(cudaEventCreate(&s->evStop));
    for (auto &e : s->evStage)
        CB_CUDA(cudaEventCreate(&e));
    int numSMs = 0, smemOptin = 0;
    CB_CUDA(cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, device));
    CB_CUDA(cudaDeviceGetAttribute(&smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    s->shape.numSMs = numSMs;

    // SphereData / PlaneData / materials (Scene.cpp:5-53) flattened to the device tables.
    std::vector<DevSphere> hs(n_spheres);
    std::vector<uint32_t> hsm(n_spheres);
    bool radiiSafe = true; // see geometry.cuh scanSpheres
    for (size_t i = 0; i < n_spheres; i++) {
        hs[i] = DevSphere{spheres[i].center[0], spheres[i].center[1], spheres[i].center[2],
                          spheres[i].radius * spheres[i].radius};
        radiiSafe = radiiSafe && hs[i].r2 >= 0x1.0p-50f;
        hsm[i] = spheres[i].material >= 0 ? static_cast<uint32_t>(spheres[i].material) : 0u; // value_or(0)
    }
    std::vector<DevPlane> hp(n_planes);
    for (size_t i = 0; i < n_planes; i++) {
        DevPlane const p = makeDevPlane(planes[i]);
        hp[i] = p;
    }
    // plane indices by axis class, then index (SceneView::planeOrder)
    std::vector<uint32_t> order(n_planes ? n_planes : 1, 0u);
    uint32_t classEnd[4] = {0, 0, 0, 0};
    {
        size_t k = 0;
        for (uint32_t c = 0; c < 4; c++) {
            for (size_t i = 0; i < n_planes; i++)
                if (hp[i].pad == c)
                    order[k++] = static_cast<uint32_t>(i);
            classEnd[c] = static_cast<uint32_t>(k);
        }
    }
    std::vector<DevAxisPlane> axis(classEnd[2] ? classEnd[2] : 1);
    for (uint32_t k = 0; k < classEnd[2]; k++)
        axis[k] = makeDevAxisPlane(hp[order[k]], static_cast<uint32_t>(n_spheres), order[k]);
    std::vector<DevMaterial> hm(n_materials);
    for (size_t i = 0; i < n_materials; i++)
        hm[i] = makeDevMaterial(materials[i].albedo, materials[i].emissive, materials[i].roughness,
                                materials[i].reflection_tint, materials[i].ior);

    CB_CUDA(s->spheres.reserve(n_spheres ? n_spheres : 1));
    CB_CUDA(s->sphereMaterial.reserve(n_spheres ? n_spheres : 1));
    CB_CUDA(s->planes.reserve(n_planes ? n_planes : 1));
    CB_CUDA(s->materials.reserve(n_materials));
    CB_CUDA(s->planeOrder.reserve(order.size()));
    CB_CUDA(s->axisPlanes.reserve(axis.size()));
    CB_CUDA(cudaMemcpyAsync(s->axisPlanes.ptr, axis.data(), axis.size() * sizeof(DevAxisPlane), cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(cudaMemcpyAsync(s->planeOrder.ptr, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    if (n_spheres) {
        CB_CUDA(cudaMemcpyAsync(s->spheres.ptr, hs.data(), n_spheres * sizeof(DevSphere), cudaMemcpyHostToDevice, s->stream));
        CB_CUDA(cudaMemcpyAsync(s->sphereMaterial.ptr, hsm.data(), n_spheres * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
    }
    if (n_planes)
        CB_CUDA(cudaMemcpyAsync(s->planes.ptr, hp.data(), n_planes * sizeof(DevPlane), cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(cudaMemcpyAsync(s->materials.ptr, hm.data(), n_materials * sizeof(DevMaterial), cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));

    s->view.spheres = s->spheres.ptr;
    s->view.sphereMaterial = s->sphereMaterial.ptr;
    s->view.planes = s->planes.ptr;
    s->view.materials = s->materials.ptr;
    s->view.nSpheres = static_cast<uint32_t>(n_spheres);
    s->view.nPlanes = static_cast<uint32_t>(n_planes);
    s->view.nMaterials = static_cast<uint32_t>(n_materials);
    s->view.radiiSafe = radiiSafe ? 1u : 0u;
    s->view.planeOrder = s->planeOrder.ptr;
    s->view.axisPlanes = s->axisPlanes.ptr;
    s->view.planeEnd[0] = classEnd[0], s->view.planeEnd[1] = classEnd[1], s->view.planeEnd[2] = classEnd[2];
    s->view.camera = makeCamera(*camera);

    // bounding box of every possible ray origin (sizes the grid's error bounds)
    s->hostSpheres.assign(spheres, spheres + n_spheres);
    sceneOriginBox(*camera, spheres, n_spheres, hp.data(), n_planes, s->boxMin, s->boxMax);
    s->smemOptin = smemOptin;
    s->shape.smemOptin = static_cast<size_t>(smemOptin);
    if (int rc = applyAcceleration(s, CORNELIS_ACCEL_AUTO))
        return rc;

    guard.p = nullptr;
    *out_scene = s;
    return CORNELIS_OK;
}

int cornelis_cuda_scene_destroy(cornelis_cuda_scene *scene) {
    delete scene;
    return CORNELIS_OK;
}

int cornelis_cuda_scene_set_stream(cornelis_cuda_scene *s, void *cudaStream) {
    if (int rc = checkScene(s))
        return rc;
    CB_CUDA(cudaStreamSynchronize(s->stream));
    s->stream = cudaStream ? static_cast<cudaStream_t>(cudaStream) : s->ownStream;
    return CORNELIS_OK;
}

int cornelis_cuda_scene_set_acceleration(cornelis_cuda_scene *s, int mode) {
    if (int rc = checkScene(s))
        return rc;
    if (mode != CORNELIS_ACCEL_AUTO && mode != CORNELIS_ACCEL_NONE && mode != CORNELIS_ACCEL_GRID)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "unknown acceleration mode");
    CB_CUDA(cudaStreamSynchronize(s->stream));
    return applyAcceleration(s, mode);
}

int cornelis_cuda_scene_acceleration(cornelis_cuda_scene *s, int *grid_enabled, uint32_t dims[3], uint64_t *references) {
    if (int rc = checkScene(s))
        return rc;
    bool const on = s->view.grid.enabled != 0;
    if (grid_enabled)
        *grid_enabled = on ? 1 : 0;
    if (dims)
        dims[0] = on ? s->grid.nx : 0, dims[1] = on ? s->grid.ny : 0, dims[2] = on ? s->grid.nz : 0;
    if (references)
        *references = on ? s->gridItems : 0;
    return CORNELIS_OK;
}

int cornelis_cuda_render_accumulate(cornelis_cuda_scene *s, const cornelis_render_params *p, cornelis_progress_fn progress,
                                    void *progressUser, cornelis_render_stats *stats) {
    if (int rc = checkScene(s))
        return rc;
    if (!p)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "params is null");
    if (p->width <= 0 || p->height <= 0)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "frame cannot be a line or the empty rectangle"); // Math.hpp:236
    if (p->samples <= 0)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "AA samples must be > 0"); // Render.cpp:310-313
    int32_t const sampleCount = p->sample_count > 0 ? p->sample_count : p->samples;
    if (p->first_sample < 0 || static_cast<int64_t>(p->first_sample) + sampleCount > (1 << 24))
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "sample range must lie in [0, 2^24)");
    if (p->max_depth > 255)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "max_depth must be <= 255 (every path ends after 255 bounces)");
    uint64_t const npix64 = static_cast<uint64_t>(p->width) * static_cast<uint64_t>(p->height);
    if (npix64 > (1ull << 31))
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "frame too large");

    int pipeline = p->pipeline;
    if (pipeline == CORNELIS_PIPELINE_DEFAULT) {
        // Few primitives in shared memory: paths stay in registers (persistent).  Grid scenes: walks differ so much in
        // length that warps pulling rays from the pool (wavefront, k_walk) keep their lanes busier.
        pipeline = s->view.grid.enabled ? CORNELIS_PIPELINE_WAVEFRONT : CORNELIS_PIPELINE_PERSISTENT;
        if (const char *env = std::getenv("CORNELIS_PIPELINE"))
            pipeline = std::strcmp(env, "wavefront") == 0    ? CORNELIS_PIPELINE_WAVEFRONT
                       : std::strcmp(env, "persistent") == 0 ? CORNELIS_PIPELINE_PERSISTENT
                                                             : pipeline;
    }
    if (pipeline != CORNELIS_PIPELINE_WAVEFRONT && pipeline != CORNELIS_PIPELINE_PERSISTENT)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "unknown pipeline");
    bool const persistent = pipeline == CORNELIS_PIPELINE_PERSISTENT;

    // Paths in flight: 2^26 (10.5 GB of pools and queues out of 180 GB) when a quarter of the free device memory
    // holds them, else 2^25, else 2^24.  A pass is five or six dependent launches, each of which has to drain before the
    // next starts, and a pass of k_walk ends with a few warps still on their longest walks; with 2^22 paths those per-pass
    // costs were 8 % of the Cornell wavefront render (4867 -> 5251 Msamples/s at 2^24) and 17 % of config 4 (1262 ->
    // 1473); config 4 at the end of round 2: 1874 / 1908 / 1933 Msamples/s with 2^24 / 2^25 / 2^26.
    uint32_t pool = p->pool_paths > 0 ? static_cast<uint32_t>(p->pool_paths) : (1u << 24);
    if (p->pool_paths <= 0 && !persistent) {
        size_t freeBytes = 0, totalBytes = 0;
        CB_CUDA(cudaMemGetInfo(&freeBytes, &totalBytes));
        for (uint32_t shift = 26; shift > 24; shift--) {
            size_t const paths = static_cast<size_t>(1) << shift;
            if (s->pool[0][0].count >= paths || paths * kPoolBytesPerPath * 4u <= freeBytes) {
                pool = static_cast<uint32_t>(paths);
                break;
            }
        }
    }
    if (const char *env = std::getenv("CORNELIS_POOL_PATHS"))
        if (p->pool_paths <= 0 && std::atoll(env) > 0)
            pool = static_cast<uint32_t>(std::atoll(env));
    uint64_t const total = npix64 * static_cast<uint64_t>(sampleCount);
    if (pool > total)
        pool = static_cast<uint32_t>(total);
    pool = (pool + 255u) & ~255u;
    bool const variance = (p->flags & CORNELIS_RENDER_VARIANCE) != 0;
    if (int rc = ensureFrame(s, static_cast<uint32_t>(p->width), static_cast<uint32_t>(p->height), persistent ? 0u : pool,
                             variance))
        return rc;

    cudaStream_t st = s->stream;
    bool const keep = (p->flags & CORNELIS_RENDER_KEEP) != 0;
    if (keep && !s->haveImage)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT,
                    "KEEP needs a previous render of this frame size on this handle (the accumulators hold no image)");
    if (!keep) {
        CB_CUDA(cudaMemsetAsync(s->accum.ptr, 0, npix64 * sizeof(float4), st));
        if (variance)
            CB_CUDA(cudaMemsetAsync(s->accum2.ptr, 0, npix64 * sizeof(float4), st));
        s->haveVariance = variance;
        s->haveImage = true;
    } else if (variance && !s->haveVariance) {
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "KEEP with VARIANCE needs a previous VARIANCE render");
    }

    RenderConfig cfg{};
    cfg.width = static_cast<uint32_t>(p->width);
    cfg.height = static_cast<uint32_t>(p->height);
    cfg.npixels = static_cast<uint32_t>(npix64);
    cfg.firstSample = static_cast<uint32_t>(p->first_sample);
    cfg.maxDepth = p->max_depth > 0 ? static_cast<uint32_t>(p->max_depth) : 0u;
    cfg.poolPaths = pool;
    cfg.variance = variance;
    cfg.key0 = static_cast<uint32_t>(p->seed);
    cfg.key1 = static_cast<uint32_t>(p->seed >> 32);
    cfg.dx = 1.0f / static_cast<float>(p->width);  // Render.cpp:31
    cfg.dy = 1.0f / static_cast<float>(p->height);
    cfg.byWidth = makeFastDiv(cfg.width);
    cfg.tilesPerRow = (CORNELIS_RAYGEN_TILES && cfg.width % kTileWidth == 0u && cfg.height % kTileHeight == 0u)
                          ? cfg.width / kTileWidth
                          : 0u;
    cfg.byTilesPerRow = makeFastDiv(cfg.tilesPerRow ? cfg.tilesPerRow : 1u);
    cfg.keys = makePhiloxKeys(cfg.key0, cfg.key1);

    Control init{};
    init.total = total;
    *s->hostControl = init;
    CB_CUDA(cudaMemcpyAsync(s->control.ptr, s->hostControl, sizeof(Control), cudaMemcpyHostToDevice, st));

    bool const profileStages = (p->flags & CORNELIS_RENDER_STAGE_TIMING) != 0;
    int const profileEvery = 32;
    float stageMs[4] = {0, 0, 0, 0};
    uint64_t profiled = 0;
    uint64_t launches = 0;

    CB_CUDA(cudaEventRecord(s->evStart, st));
    int cur = 0;
    uint64_t pass = 0;
    int const passesPerBatch = 16;
    bool aborted = false;
    if (persistent) {
        // One launch renders a slice of the camera-path range; between slices the host reports progress and may abort.
        bool const drop = (p->flags & CORNELIS_RENDER_DROP_NONFINITE) != 0;
        // With a callback the render is cut into at least two slices so that it is consulted at least once.
        // Without a callback there is nothing to report or abort: one launch, one drain at the end.
        uint64_t slice = total;
        if (progress) {
            slice = std::max<uint64_t>(total / 32 + 1, 1ull << 22);
            if (total > 1)
                slice = std::min<uint64_t>(slice, (total + 1) / 2);
        }
        uint64_t done = 0;
        while (done < total) {
            uint64_t const limit = std::min(total, done + slice);
            unsigned long long const start = done;
            CB_CUDA(cudaMemcpyAsync(s->claimCursor.ptr, &start, sizeof start, cudaMemcpyHostToDevice, st));
            cfg.claim = persistentClaim(limit - done, s->gridPersistent);
            launchPersistent(st, s->shape, s->gridPersistent, cfg, s->view, s->claimCursor.ptr, limit, s->accum.ptr,
                             variance ? s->accum2.ptr : nullptr, drop, s->control.ptr);
            launches += 1;
            CB_CUDA(cudaStreamSynchronize(st));
            done = limit;
            if (done < total && progress && progress(progressUser, done, total) != 0) {
                aborted = true;
                break;
            }
        }
        CB_CUDA(cudaMemcpyAsync(s->hostControl, s->control.ptr, sizeof(Control), cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaStreamSynchronize(st));
        s->hostControl->iterations = launches;
    }
    for (; !persistent;) {
        for (int k = 0; k < passesPerBatch; k++, pass++) {
            PathPool const in = poolView(s, cur), out = poolView(s, cur ^ 1);
            bool const timed = profileStages && (pass % profileEvery == profileEvery - 1);
            launchPlan(st, s->control.ptr, cfg);
            if (timed)
                cudaEventRecord(s->evStage[0], st);
            launchRaygen(st, s->shape, s->control.ptr, cfg, s->view.camera, in);
            if (timed)
                cudaEventRecord(s->evStage[1], st);
            launchIntersect(st, s->shape, s->control.ptr, s->view, in, s->hits.ptr, s->hitQueue.ptr, s->finished.ptr);
            if (timed)
                cudaEventRecord(s->evStage[2], st);
            launchShade(st, s->shape, s->control.ptr, cfg, s->view, in, out, s->hits.ptr, s->hitQueue.ptr,
                        s->finished.ptr);
            if (timed)
                cudaEventRecord(s->evStage[3], st);
            launchAccumulate(st, s->shape, s->control.ptr, s->finished.ptr, s->accum.ptr,
                             variance ? s->accum2.ptr : nullptr, (p->flags & CORNELIS_RENDER_DROP_NONFINITE) != 0);
            if (timed) {
                cudaEventRecord(s->evStage[4], st);
                CB_CUDA(cudaEventSynchronize(s->evStage[4]));
                for (int g = 0; g < 4; g++) {
                    float ms = 0;
                    cudaEventElapsedTime(&ms, s->evStage[g], s->evStage[g + 1]);
                    stageMs[g] += ms;
                }
                profiled++;
            }
            launches += (s->view.grid.enabled && s->shape.walkPull) ? 6 : 5;
            cur ^= 1;
        }
        CB_CUDA(cudaMemcpyAsync(s->hostControl, s->control.ptr, sizeof(Control), cudaMemcpyDeviceToHost, st));
        CB_CUDA(cudaStreamSynchronize(st));
        Control const &c = *s->hostControl;
        if (c.cursor >= c.total && c.nSurvive == 0)
            break;
        if (progress && progress(progressUser, c.cursor, c.total) != 0) {
            aborted = true;
            break;
        }
    }
    CB_CUDA(cudaEventRecord(s->evStop, st));
    CB_CUDA(cudaEventSynchronize(s->evStop));
    CB_CUDA(cudaGetLastError());

    if (stats) {
        Control const &c = *s->hostControl;
        std::memset(stats, 0, sizeof *stats);
        stats->pixel_samples = c.cursor;
        stats->rays = c.rays;
        stats->shaded_hits = c.shaded;
        stats->iterations = c.iterations;
        stats->kernel_launches = launches;
        stats->contributions = c.contributions;
        stats->max_depth = c.maxDepth;
        cudaEventElapsedTime(&stats->gpu_ms, s->evStart, s->evStop);
        if (profiled) {
            // mean milliseconds per launch over the sampled passes
            stats->raygen_ms = stageMs[0] / profiled;
            stats->intersect_ms = stageMs[1] / profiled;
            stats->shade_ms = stageMs[2] / profiled;
            stats->accumulate_ms = stageMs[3] / profiled;
        }
    }
    if (aborted)
        return fail(CORNELIS_ERR_ABORTED, "render aborted by the progress callback");
    return CORNELIS_OK;
}

int cornelis_cuda_framebuffer_device(cornelis_cuda_scene *s, void **devicePtr, size_t *nFloats) {
    if (int rc = checkScene(s))
        return rc;
    if (!devicePtr || !nFloats)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "null output pointer");
    if (!s->accum.ptr || !s->width || !s->haveImage)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "nothing has been rendered yet");
    *devicePtr = s->accum.ptr;
    *nFloats = static_cast<size_t>(s->width) * s->height * 4;
    return CORNELIS_OK;
}

int cornelis_cuda_resolve(cornelis_cuda_scene *s, int32_t samples, float *hostRgb, float *hostVariance) {
    if (int rc = checkScene(s))
        return rc;
    if (samples <= 0 || !hostRgb)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "samples must be > 0 and host_rgb non-null");
    if (!s->accum.ptr || !s->width || !s->haveImage)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "nothing has been rendered yet");
    if (hostVariance && !s->haveVariance)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "variance was not accumulated (CORNELIS_RENDER_VARIANCE)");
    size_t const npix = static_cast<size_t>(s->width) * s->height;
    CB_CUDA(s->outRgb.reserve(3 * npix));
    if (hostVariance)
        CB_CUDA(s->outVar.reserve(3 * npix));
    launchResolve(s->stream, s->shape, static_cast<uint32_t>(npix), static_cast<uint32_t>(samples), s->accum.ptr,
                  hostVariance ? s->accum2.ptr : nullptr, s->outRgb.ptr, hostVariance ? s->outVar.ptr : nullptr);
    CB_CUDA(cudaMemcpyAsync(hostRgb, s->outRgb.ptr, 3 * npix * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    if (hostVariance)
        CB_CUDA(cudaMemcpyAsync(hostVariance, s->outVar.ptr, 3 * npix * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_resolve_device(cornelis_cuda_scene *s, int32_t samples, void **deviceRgb) {
    if (int rc = checkScene(s))
        return rc;
    if (samples <= 0 || !deviceRgb)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "samples must be > 0 and device_rgb non-null");
    if (!s->accum.ptr || !s->width || !s->haveImage)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "nothing has been rendered yet");
    size_t const npix = static_cast<size_t>(s->width) * s->height;
    CB_CUDA(s->outRgb.reserve(3 * npix));
    launchResolve(s->stream, s->shape, static_cast<uint32_t>(npix), static_cast<uint32_t>(samples), s->accum.ptr, nullptr,
                  s->outRgb.ptr, nullptr);
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    *deviceRgb = s->outRgb.ptr;
    return CORNELIS_OK;
}

int cornelis_cuda_resolve_srgb8(cornelis_cuda_scene *s, int32_t samples, uint8_t *hostRgb8) {
    if (int rc = checkScene(s))
        return rc;
    if (samples <= 0 || !hostRgb8)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "samples must be > 0 and host_rgb8 non-null");
    if (!s->accum.ptr || !s->width || !s->haveImage)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "nothing has been rendered yet");
    size_t const npix = static_cast<size_t>(s->width) * s->height;
    CB_CUDA(s->outRgb8.reserve(3 * npix));
    launchResolveSrgb8(s->stream, s->shape, static_cast<uint32_t>(npix), static_cast<uint32_t>(samples), s->accum.ptr,
                       s->outRgb8.ptr);
    CB_CUDA(cudaMemcpyAsync(hostRgb8, s->outRgb8.ptr, 3 * npix, cudaMemcpyDeviceToHost, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_render(cornelis_cuda_scene *s, const cornelis_render_params *p, float *hostRgb,
                         cornelis_render_stats *stats) {
    if (!hostRgb)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "host_rgb is null");
    if (int rc = cornelis_cuda_render_accumulate(s, p, nullptr, nullptr, stats))
        return rc;
    return cornelis_cuda_resolve(s, p->samples, hostRgb, nullptr);
}

// ------------------------------------------------------------------------------------------ stage entry points --

#define UPLOAD(buf, src, count)                                                                                        \
    do {                                                                                                               \
        CB_CUDA((buf).reserve((count) ? (count) : 1));                                                                 \
        CB_CUDA(cudaMemcpyAsync((buf).ptr, (src), (count) * sizeof(*(buf).ptr), cudaMemcpyHostToDevice, s->stream));   \
    } while (0)
#define DOWNLOAD(dst, buf, count)                                                                                      \
    CB_CUDA(cudaMemcpyAsync((dst), (buf).ptr, (count) * sizeof(*(buf).ptr), cudaMemcpyDeviceToHost, s->stream))

int cornelis_cuda_pixel_rays(cornelis_cuda_scene *s, int32_t width, int32_t height, size_t n, const int32_t *pi,
                             const int32_t *pj, const float *phi1, const float *phi2, float *org, float *dir) {
    if (int rc = checkScene(s))
        return rc;
    if (width <= 0 || height <= 0 || !pi || !pj || !phi1 || !phi2 || !org || !dir)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "bad argument");
    if (n == 0)
        return CORNELIS_OK;
    UPLOAD(s->stageI[0], pi, n);
    UPLOAD(s->stageI[1], pj, n);
    UPLOAD(s->stageF[0], phi1, n);
    UPLOAD(s->stageF[1], phi2, n);
    CB_CUDA(s->stageF[2].reserve(3 * n));
    CB_CUDA(s->stageF[3].reserve(3 * n));
    launchPixelRays(s->stream, s->shape, s->view.camera, static_cast<uint32_t>(n), 1.0f / static_cast<float>(width),
                    1.0f / static_cast<float>(height), s->stageI[0].ptr, s->stageI[1].ptr, s->stageF[0].ptr,
                    s->stageF[1].ptr, s->stageF[2].ptr, s->stageF[3].ptr);
    DOWNLOAD(org, s->stageF[2], 3 * n);
    DOWNLOAD(dir, s->stageF[3], 3 * n);
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_intersect(cornelis_cuda_scene *s, size_t n, const float *org, const float *dir, float *t,
                            int32_t *prim, float *P, float *N, int32_t *mat) {
    if (int rc = checkScene(s))
        return rc;
    if (!org || !dir || !t || !prim)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "org, dir, t and prim are required");
    if (n == 0)
        return CORNELIS_OK;
    UPLOAD(s->stageF[0], org, 3 * n);
    UPLOAD(s->stageF[1], dir, 3 * n);
    CB_CUDA(s->stage4[0].reserve(n));
    CB_CUDA(s->stage4[1].reserve(n));
    CB_CUDA(s->stageHits.reserve(n));
    launchPack4(s->stream, s->shape, n, s->stageF[0].ptr, s->stage4[0].ptr);
    launchPack4(s->stream, s->shape, n, s->stageF[1].ptr, s->stage4[1].ptr);
    launchIntersectBatch(s->stream, s->shape, s->view, n, s->stage4[0].ptr, s->stage4[1].ptr, s->stageHits.ptr);
    CB_CUDA(s->stageF[2].reserve(n));
    CB_CUDA(s->stageI[0].reserve(n));
    launchUnpackHits(s->stream, s->shape, n, s->stageHits.ptr, s->stageF[2].ptr, s->stageI[0].ptr);
    DOWNLOAD(t, s->stageF[2], n);
    DOWNLOAD(prim, s->stageI[0], n);
    if (P || N || mat) {
        if (P)
            CB_CUDA(s->stageF[3].reserve(3 * n));
        if (N)
            CB_CUDA(s->stageF[4].reserve(3 * n));
        if (mat)
            CB_CUDA(s->stageI[1].reserve(n));
        launchHitSurface(s->stream, s->shape, s->view, n, s->stage4[0].ptr, s->stage4[1].ptr, s->stageHits.ptr,
                         P ? s->stageF[3].ptr : nullptr, N ? s->stageF[4].ptr : nullptr, mat ? s->stageI[1].ptr : nullptr);
        if (P)
            DOWNLOAD(P, s->stageF[3], 3 * n);
        if (N)
            DOWNLOAD(N, s->stageF[4], 3 * n);
        if (mat)
            DOWNLOAD(mat, s->stageI[1], n);
    }
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_intersect_device(cornelis_cuda_scene *s, size_t n, const void *dOrg4, const void *dDir4, void *dHit2,
                                   int repeats, float *msPerLaunch) {
    if (int rc = checkScene(s))
        return rc;
    if (!dOrg4 || !dDir4 || !dHit2 || repeats <= 0)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "bad argument");
    CB_CUDA(cudaEventRecord(s->evStart, s->stream));
    for (int r = 0; r < repeats; r++)
        launchIntersectBatch(s->stream, s->shape, s->view, n, static_cast<const float4 *>(dOrg4),
                             static_cast<const float4 *>(dDir4), static_cast<HitRecord *>(dHit2));
    CB_CUDA(cudaEventRecord(s->evStop, s->stream));
    CB_CUDA(cudaEventSynchronize(s->evStop));
    CB_CUDA(cudaGetLastError());
    if (msPerLaunch) {
        float ms = 0;
        cudaEventElapsedTime(&ms, s->evStart, s->evStop);
        *msPerLaunch = ms / repeats;
    }
    return CORNELIS_OK;
}

int cornelis_cuda_intersect_compact(cornelis_cuda_scene *s, size_t n, const float *org, const float *dir,
                                    uint32_t *hitQueue, uint32_t *nHits, uint32_t *missQueue, uint32_t *nMisses) {
    if (int rc = checkScene(s))
        return rc;
    if (!org || !dir || !hitQueue || !nHits || !missQueue || !nMisses)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "null argument");
    *nHits = *nMisses = 0;
    if (n == 0)
        return CORNELIS_OK;
    if (n > (1u << 26)) // 80 B of pool and queues per ray on the device, 32 B of staging on the host
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "batch too large (at most 2^26 rays)");
    // the pool as a render pass would hold these rays: origin, direction, throughput 1, radiance (1, 0, 0) | ray index
    for (int a = 0; a < 4; a++)
        CB_CUDA(s->pool[0][a].reserve(n));
    CB_CUDA(s->hits.reserve(n));
    CB_CUDA(s->hitQueue.reserve(n));
    CB_CUDA(s->finished.reserve(2 * n));
    CB_CUDA(s->control.reserve(1));
    UPLOAD(s->stageF[0], org, 3 * n);
    UPLOAD(s->stageF[1], dir, 3 * n);
    launchPack4(s->stream, s->shape, n, s->stageF[0].ptr, s->pool[0][0].ptr);
    launchPack4(s->stream, s->shape, n, s->stageF[1].ptr, s->pool[0][1].ptr);
    std::vector<float4> state(n);
    for (size_t k = 0; k < n; k++)
        state[k] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    CB_CUDA(cudaMemcpyAsync(s->pool[0][2].ptr, state.data(), n * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    for (size_t k = 0; k < n; k++) {
        uint32_t const index = static_cast<uint32_t>(k);
        float w;
        std::memcpy(&w, &index, sizeof w);
        state[k] = make_float4(1.0f, 0.0f, 0.0f, w);
    }
    CB_CUDA(cudaMemcpyAsync(s->pool[0][3].ptr, state.data(), n * sizeof(float4), cudaMemcpyHostToDevice, s->stream));
    Control init{};
    init.nIn = static_cast<uint32_t>(n);
    // grid scenes: the first half of the batch plays the survivors of a previous pass (k_walk's pull model), the second
    // half this pass's new camera rays (its packet phase), as k_plan would have laid them out
    init.genBase = static_cast<uint32_t>(n / 2);
    init.walkCursorCamera = n / 2;
    *s->hostControl = init;
    CB_CUDA(cudaMemcpyAsync(s->control.ptr, s->hostControl, sizeof(Control), cudaMemcpyHostToDevice, s->stream));
    launchIntersect(s->stream, s->shape, s->control.ptr, s->view, poolView(s, 0), s->hits.ptr, s->hitQueue.ptr,
                    s->finished.ptr);
    CB_CUDA(cudaMemcpyAsync(s->hostControl, s->control.ptr, sizeof(Control), cudaMemcpyDeviceToHost, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    uint32_t const hitCount = s->hostControl->nHit, missCount = s->hostControl->nFinished;
    if (hitCount > n || missCount > n)
        return fail(CORNELIS_ERR_CUDA, "queue tails exceed the batch");
    if (hitCount)
        CB_CUDA(cudaMemcpyAsync(hitQueue, s->hitQueue.ptr, hitCount * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    std::vector<FinishedPath> done(missCount);
    if (missCount)
        CB_CUDA(cudaMemcpyAsync(done.data(), s->finished.ptr, missCount * sizeof(FinishedPath), cudaMemcpyDeviceToHost,
                                s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t k = 0; k < missCount; k++)
        missQueue[k] = done[k].pixel;
    *nHits = hitCount;
    *nMisses = missCount;
    return CORNELIS_OK;
}

static int checkMaterialIds(cornelis_cuda_scene *s, size_t n, const int32_t *mat) {
    for (size_t k = 0; k < n; k++)
        if (mat[k] < 0 || static_cast<uint32_t>(mat[k]) >= s->view.nMaterials)
            return fail(CORNELIS_ERR_INVALID_ARGUMENT, "material index out of range");
    return CORNELIS_OK;
}

int cornelis_cuda_bsdf_sample(cornelis_cuda_scene *s, size_t n, const int32_t *mat, const float *wo, const float *N,
                              const float *x, float *wi, float *pdf, float *f) {
    if (int rc = checkScene(s))
        return rc;
    if (!mat || !wo || !N || !x || !wi || !pdf || !f)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0)
        return CORNELIS_OK;
    if (int rc = checkMaterialIds(s, n, mat))
        return rc;
    UPLOAD(s->stageI[0], mat, n);
    UPLOAD(s->stageF[0], wo, 3 * n);
    UPLOAD(s->stageF[1], N, 3 * n);
    UPLOAD(s->stageF[2], x, 3 * n);
    CB_CUDA(s->stageF[3].reserve(3 * n));
    CB_CUDA(s->stageF[4].reserve(n));
    CB_CUDA(s->stageF[5].reserve(3 * n));
    launchBsdfSample(s->stream, s->shape, s->materials.ptr, static_cast<uint32_t>(n), s->stageI[0].ptr, s->stageF[0].ptr,
                     s->stageF[1].ptr, s->stageF[2].ptr, s->stageF[3].ptr, s->stageF[4].ptr, s->stageF[5].ptr);
    DOWNLOAD(wi, s->stageF[3], 3 * n);
    DOWNLOAD(pdf, s->stageF[4], n);
    DOWNLOAD(f, s->stageF[5], 3 * n);
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_bsdf_eval(cornelis_cuda_scene *s, size_t n, const int32_t *mat, const float *wi, const float *wo,
                            const float *N, float *f, float *pdf) {
    if (int rc = checkScene(s))
        return rc;
    if (!mat || !wi || !wo || !N || !f || !pdf)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0)
        return CORNELIS_OK;
    if (int rc = checkMaterialIds(s, n, mat))
        return rc;
    UPLOAD(s->stageI[0], mat, n);
    UPLOAD(s->stageF[0], wi, 3 * n);
    UPLOAD(s->stageF[1], wo, 3 * n);
    UPLOAD(s->stageF[2], N, 3 * n);
    CB_CUDA(s->stageF[3].reserve(3 * n));
    CB_CUDA(s->stageF[4].reserve(n));
    launchBsdfEval(s->stream, s->shape, s->materials.ptr, static_cast<uint32_t>(n), s->stageI[0].ptr, s->stageF[0].ptr,
                   s->stageF[1].ptr, s->stageF[2].ptr, s->stageF[3].ptr, s->stageF[4].ptr);
    DOWNLOAD(f, s->stageF[3], 3 * n);
    DOWNLOAD(pdf, s->stageF[4], n);
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_shade(cornelis_cuda_scene *s, size_t n, int32_t depth, const float *u, const float *P, const float *N,
                        const int32_t *mat, float *org, float *dir, float *thr, float *rad, uint8_t *alive) {
    if (int rc = checkScene(s))
        return rc;
    if (!u || !P || !N || !mat || !org || !dir || !thr || !rad || !alive || depth < 0)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0)
        return CORNELIS_OK;
    if (int rc = checkMaterialIds(s, n, mat))
        return rc;
    UPLOAD(s->stageI[0], mat, n);
    UPLOAD(s->stageF[0], u, 4 * n);
    UPLOAD(s->stageF[1], P, 3 * n);
    UPLOAD(s->stageF[2], N, 3 * n);
    UPLOAD(s->stageF[3], org, 3 * n);
    UPLOAD(s->stageF[4], dir, 3 * n);
    UPLOAD(s->stageF[5], thr, 3 * n);
    UPLOAD(s->stageF[6], rad, 3 * n);
    CB_CUDA(s->outRgb8.reserve(n));
    launchShadeExplicit(s->stream, s->shape, s->materials.ptr, static_cast<uint32_t>(n), static_cast<uint32_t>(depth),
                        s->stageF[0].ptr, s->stageF[1].ptr, s->stageF[2].ptr, s->stageI[0].ptr, s->stageF[3].ptr,
                        s->stageF[4].ptr, s->stageF[5].ptr, s->stageF[6].ptr, s->outRgb8.ptr);
    DOWNLOAD(org, s->stageF[3], 3 * n);
    DOWNLOAD(dir, s->stageF[4], 3 * n);
    DOWNLOAD(thr, s->stageF[5], 3 * n);
    DOWNLOAD(rad, s->stageF[6], 3 * n);
    DOWNLOAD(alive, s->outRgb8, n);
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

static int rngCommon(cornelis_cuda_scene *s, int rounds, uint64_t seed, size_t n, const uint32_t *pixel,
                     const uint32_t *sample, const uint32_t *block, float *uniforms, uint32_t *bits) {
    if (int rc = checkScene(s))
        return rc;
    if (!pixel || !sample || !block || (!uniforms && !bits))
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0)
        return CORNELIS_OK;
    DeviceBuffer<int32_t> &a = s->stageI[0], &b = s->stageI[1], &c = s->stageI[2];
    CB_CUDA(a.reserve(n));
    CB_CUDA(b.reserve(n));
    CB_CUDA(c.reserve(n));
    CB_CUDA(cudaMemcpyAsync(a.ptr, pixel, n * 4, cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(cudaMemcpyAsync(b.ptr, sample, n * 4, cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(cudaMemcpyAsync(c.ptr, block, n * 4, cudaMemcpyHostToDevice, s->stream));
    CB_CUDA(s->stageF[0].reserve(4 * n));
    launchRng(s->stream, s->shape, static_cast<uint32_t>(n), rounds, static_cast<uint32_t>(seed),
              static_cast<uint32_t>(seed >> 32), reinterpret_cast<const uint32_t *>(a.ptr),
              reinterpret_cast<const uint32_t *>(b.ptr), reinterpret_cast<const uint32_t *>(c.ptr),
              uniforms ? s->stageF[0].ptr : nullptr, bits ? reinterpret_cast<uint32_t *>(s->stageF[0].ptr) : nullptr);
    CB_CUDA(cudaMemcpyAsync(uniforms ? static_cast<void *>(uniforms) : static_cast<void *>(bits), s->stageF[0].ptr,
                            4 * n * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_rng_uniforms(cornelis_cuda_scene *s, uint64_t seed, size_t n, const uint32_t *pixel,
                               const uint32_t *sample, const uint32_t *block, float *out) {
    return rngCommon(s, 0, seed, n, pixel, sample, block, out, nullptr);
}

int cornelis_cuda_rng_rounds(void) { return kPhiloxRounds; }

int cornelis_cuda_rng_bits(cornelis_cuda_scene *s, int rounds, uint64_t seed, size_t n, const uint32_t *pixel,
                           const uint32_t *sample, const uint32_t *block, uint32_t *out) {
    if (rounds < 1 || rounds > 16)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "rounds must be in [1, 16]");
    return rngCommon(s, rounds, seed, n, pixel, sample, block, nullptr, out);
}

int cornelis_cuda_selftest_srgb8(cornelis_cuda_scene *s, uint32_t firstBits, size_t n, uint8_t *hostOut) {
    if (int rc = checkScene(s))
        return rc;
    if (!hostOut)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "host_out is null");
    if (n == 0)
        return CORNELIS_OK;
    CB_CUDA(s->outRgb8.reserve(n));
    launchSrgb8Sweep(s->stream, s->shape, firstBits, n, s->outRgb8.ptr);
    CB_CUDA(cudaMemcpyAsync(hostOut, s->outRgb8.ptr, n, cudaMemcpyDeviceToHost, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    return CORNELIS_OK;
}

int cornelis_cuda_selftest_arith(cornelis_cuda_scene *s, int mode, uint64_t n, uint32_t seed, uint64_t *mismatches) {
    if (int rc = checkScene(s))
        return rc;
    if (!mismatches || mode < 0 || mode > 2)
        return fail(CORNELIS_ERR_INVALID_ARGUMENT, "bad argument");
    DeviceBuffer<unsigned long long> counter;
    CB_CUDA(counter.reserve(1));
    CB_CUDA(cudaMemsetAsync(counter.ptr, 0, sizeof(unsigned long long), s->stream));
    launchSelftestArith(s->stream, s->shape, mode, n, seed, counter.ptr);
    unsigned long long host = 0;
    CB_CUDA(cudaMemcpyAsync(&host, counter.ptr, sizeof host, cudaMemcpyDeviceToHost, s->stream));
    CB_CUDA(cudaStreamSynchronize(s->stream));
    CB_CUDA(cudaGetLastError());
    counter.release();
    *mismatches = host;
    return CORNELIS_OK;
}

} // extern "C"
