// Library-wide state of the C ABI: error string, launch counter, version.
#include <string.h>

#include "dm_common.cuh"

namespace dm {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
int g_tuning[DM_TUNE_COUNT] = {1, 1};
}  // namespace dm

extern "C" int dm_set_tuning(int knob, int value) {
    DM_REQUIRE(knob >= 0 && knob < DM_TUNE_COUNT);
    dm::g_tuning[knob] = value;
    return DM_OK;
}
extern "C" int dm_get_tuning(int knob) { return (knob >= 0 && knob < DM_TUNE_COUNT) ? dm::g_tuning[knob] : DM_ERR_INVALID; }

extern "C" int dm_version(void) { return 100; }
extern "C" const char* dm_last_error(void) { return dm::g_err; }
extern "C" unsigned long long dm_launch_count(void) { return dm::g_launches.load(); }
extern "C" void dm_reset_launch_count(void) { dm::g_launches.store(0); }

// Let kernels launched on `device` dereference memory of `peer_device` (buffers of another process mapped through CUDA
// IPC are opened under the exporting device, which leaves the importing device without peer access).  The device is set
// explicitly: this library carries its own (static) CUDA runtime, whose notion of the current device is not the host
// framework's.
extern "C" int dm_enable_peer_access(int device, int peer_device) {
    if (device == peer_device) return DM_OK;
    int prev = -1, can = 0;
    DM_CUDA(cudaGetDevice(&prev));
    DM_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
    if (!can) return dm::fail(DM_ERR_UNSUPPORTED, "%s: device %d cannot access device %d", __func__, device, peer_device);
    DM_CUDA(cudaSetDevice(device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        (void)cudaGetLastError();
        e = cudaSuccess;
    }
    if (prev >= 0) (void)cudaSetDevice(prev);
    if (e != cudaSuccess)
        return dm::fail(DM_ERR_CUDA, "%s: cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", __func__, device,
                        peer_device, cudaGetErrorString(e));
    return DM_OK;
}

// ---- peer-visible buffers: cudaMalloc'ed here, exported / imported as CUDA IPC handles with the IMPORTING device current,
// so the returned pointers can be dereferenced by kernels running on the importing device (lazy peer access).
namespace {
struct DeviceScope {
    int prev = -1;
    bool ok = false;
    explicit DeviceScope(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceScope() {
        if (prev >= 0) (void)cudaSetDevice(prev);
    }
};
}  // namespace

extern "C" int dm_peer_alloc(int device, long long bytes, void** ptr) {
    DM_REQUIRE(ptr != nullptr && bytes > 0);
    DeviceScope scope(device);
    if (!scope.ok) return dm::fail(DM_ERR_CUDA, "%s: cannot select device %d", __func__, device);
    DM_CUDA(cudaMalloc(ptr, (size_t)bytes));
    DM_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    DM_CUDA(cudaDeviceSynchronize());
    return DM_OK;
}
extern "C" int dm_peer_free(int device, void* ptr) {
    DeviceScope scope(device);
    if (ptr) DM_CUDA(cudaFree(ptr));
    return DM_OK;
}
extern "C" int dm_ipc_export(int device, const void* ptr, unsigned char* handle64) {
    DM_REQUIRE(ptr != nullptr && handle64 != nullptr);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    DeviceScope scope(device);
    cudaIpcMemHandle_t h;
    DM_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle64, &h, 64);
    return DM_OK;
}
extern "C" int dm_ipc_open(int device, const unsigned char* handle64, void** ptr) {
    DM_REQUIRE(ptr != nullptr && handle64 != nullptr);
    DeviceScope scope(device);
    if (!scope.ok) return dm::fail(DM_ERR_CUDA, "%s: cannot select device %d", __func__, device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    DM_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return DM_OK;
}
extern "C" int dm_ipc_close(int device, void* ptr) {
    DeviceScope scope(device);
    if (ptr) DM_CUDA(cudaIpcCloseMemHandle(ptr));
    return DM_OK;
}
