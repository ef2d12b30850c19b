// Set-up helpers for the multi-GPU ragged gather (one process per GPU).
//
// The reference has no multi-GPU path at all; images are independent, so each
// rank runs the whole hot path on its own slice and nms_finalize_kernel stores the kept rows of
// its images straight into the root rank's result buffer through an NVLink peer mapping.  These
// calls only create / share / map that buffer (CUDA IPC); they run once at set-up, never per batch.
#include <cuda_runtime.h>
#include <string.h>

#include "../../include/yolo_b200.h"

extern "C" int yolo_b200_device_alloc(size_t bytes, void** out) {
    if (!out) return YOLO_B200_E_NULL;
    if (bytes == 0) return YOLO_B200_E_RANGE;
    return (int)cudaMalloc(out, bytes);      // plain cudaMalloc: exportable with cudaIpcGetMemHandle
}

extern "C" int yolo_b200_device_free(void* ptr) { return (int)cudaFree(ptr); }

extern "C" int yolo_b200_peer_export(void* dev_ptr, void* handle_out_host) {
    if (!dev_ptr || !handle_out_host) return YOLO_B200_E_NULL;
    static_assert(sizeof(cudaIpcMemHandle_t) == YOLO_B200_PEER_HANDLE_BYTES, "handle size");
    return (int)cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out_host), dev_ptr);
}

extern "C" int yolo_b200_peer_open(const void* handle_host, void** out) {
    if (!handle_host || !out) return YOLO_B200_E_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}

extern "C" int yolo_b200_peer_close(void* mapped) { return (int)cudaIpcCloseMemHandle(mapped); }
