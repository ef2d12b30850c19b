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

// ------------------------------------------------------------------------------------------------
// Step flags of the gather protocol (pytorch_yolo_b200/sharded.py): 32-bit sequence numbers in device memory -- the
// root's, possibly reached through the NVLink peer mapping -- written with release and polled with acquire semantics at
// system scope.  Every kernel keeps its own use counter in device memory (`seq`), so the launches are identical from
// step to step and can be replayed from a CUDA graph.  Waits are bounded by the global timer: on expiry the kernel
// records the failure in *err and returns (nothing traps, nothing hangs the GPU).
namespace {

__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// *seq += 1; wait until flags[i] >= *seq + bias for every i < n.
__global__ void flag_wait_kernel(const int32_t* flags, int n, int32_t* seq, int bias, int32_t* err,
                                 unsigned long long timeout_ns) {
    __shared__ int s_want;
    if (threadIdx.x == 0) { const int v = *seq + 1; *seq = v; s_want = v + bias; }
    __syncthreads();
    const int want = s_want;
    const unsigned long long t0 = global_ns();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned spins = 0;
        while (ld_acquire_sys(flags + i) - want < 0) {
            if ((++spins & 63u) == 0 && global_ns() - t0 > timeout_ns) {
                if (err) atomicMax(err, 1 + i);
                break;
            }
            __nanosleep(64);
        }
    }
}

// *seq += 1; *flag = *seq + bias  (release: everything this stream did before is visible to whoever acquires the flag)
__global__ void flag_post_kernel(int32_t* flag, int32_t* seq, int bias) {
    const int v = *seq + 1;
    *seq = v;
    __threadfence_system();
    st_release_sys(flag, v + bias);
}

}  // namespace

extern "C" int yolo_b200_flag_wait(const int32_t* flags, int n_flags, int32_t* seq, int bias, int32_t* err,
                                   double timeout_s, yolo_b200_stream_t stream) {
    if (!flags || !seq) return YOLO_B200_E_NULL;
    if (n_flags < 1 || n_flags > 1024 || !(timeout_s > 0.0)) return YOLO_B200_E_RANGE;
    const int threads = n_flags < 32 ? 32 : (n_flags + 31) / 32 * 32;
    flag_wait_kernel<<<1, threads, 0, stream>>>(flags, n_flags, seq, bias, err, (unsigned long long)(timeout_s * 1e9));
    return (int)cudaGetLastError();
}

extern "C" int yolo_b200_flag_post(int32_t* flag, int32_t* seq, int bias, yolo_b200_stream_t stream) {
    if (!flag || !seq) return YOLO_B200_E_NULL;
    flag_post_kernel<<<1, 1, 0, stream>>>(flag, seq, bias);
    return (int)cudaGetLastError();
}
