// Post-NMS epilogue: the reference's scale_coords + the rounding step of _dict_from_results
// (utils/utils.py:296-303 and :313): undo the letterbox padding, divide by the resize gain, clamp at 0, round.
//
//   x1,x2 -= pad_x ; y1,y2 -= pad_y ; all four /= gain ; max(.,0) ; rint        -- each step rounds to fp32 on its
// own, exactly like the reference's sequence of in-place torch ops (subtract, IEEE divide, clamp, round-half-even).
#include "common.cuh"

namespace yb {

__device__ __forceinline__ float unletterbox(float v, float pad, float gain, int do_round) {
    const float t = __fdiv_rn(__fsub_rn(v, pad), gain);             // utils.py:299-301
    const float r = (t != t) ? t : fmaxf(t, 0.0f);                   // utils.py:302 clamp(min=0); NaN stays NaN like torch.clamp
    return do_round ? rintf(r) : r;                                  // utils.py:313 .round(): half to even
}

// one thread per (row, coordinate); rows of one tensor, `row_stride` floats apart (a (n,4) view of (n,7) rows)
__global__ void __launch_bounds__(256)
scale_coords_kernel(float* coords, int n, int row_stride, float pad_x, float pad_y, float gain, int do_round) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * 4) return;
    const int r = e >> 2, c = e & 3;
    float* p = coords + (size_t)r * row_stride + c;
    *p = unletterbox(*p, (c & 1) ? pad_y : pad_x, gain, do_round);
}

// batched over a Detector result: image b owns rows [b*out_cap, b*out_cap + out_count[b]) of 7 floats;
// params[b] = (pad_x, pad_y, gain)
__global__ void __launch_bounds__(256)
scale_detections_kernel(float* out, const int32_t* out_count, int out_cap, const float* params, int do_round) {
    const int b = blockIdx.y;
    const int n = out_count[b];
    const float pad_x = params[b * 3 + 0], pad_y = params[b * 3 + 1], gain = params[b * 3 + 2];
    float* base = out + (size_t)b * out_cap * YOLO_B200_DET_COLS;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * 4; e += gridDim.x * blockDim.x) {
        const int r = e >> 2, c = e & 3;
        float* p = base + (size_t)r * YOLO_B200_DET_COLS + c;
        *p = unletterbox(*p, (c & 1) ? pad_y : pad_x, gain, do_round);
    }
}

}  // namespace yb

extern "C" int yolo_b200_scale_coords(float* coords, int n, int row_stride, float pad_x, float pad_y, float gain,
                                      int do_round, yolo_b200_stream_t stream) {
    if (!coords && n > 0) return YOLO_B200_E_NULL;
    if (n < 0 || row_stride < 4 || !(gain == gain) || gain == 0.0f) return YOLO_B200_E_RANGE;
    if ((uintptr_t)coords & 3u) return YOLO_B200_E_ALIGN;
    if (n == 0) return 0;
    const int threads = 256, blocks = (n * 4 + threads - 1) / threads;
    yb::scale_coords_kernel<<<blocks, threads, 0, stream>>>(coords, n, row_stride, pad_x, pad_y, gain, do_round);
    return (int)cudaGetLastError();
}

extern "C" int yolo_b200_scale_detections(float* out, const int32_t* out_count, int batch, int out_cap,
                                          const float* params, int do_round, yolo_b200_stream_t stream) {
    if (!out || !out_count || !params) return YOLO_B200_E_NULL;
    if (batch < 0 || out_cap < 1 || batch > 65535) return YOLO_B200_E_RANGE;
    if (batch == 0) return 0;
    const int threads = 256;
    int bx = (out_cap * 4 + threads - 1) / threads;
    if (bx > 8) bx = 8;
    yb::scale_detections_kernel<<<dim3(bx, batch), threads, 0, stream>>>(out, out_count, out_cap, params, do_round);
    return (int)cudaGetLastError();
}
