// Training-side consumer of the YOLOLayer constants: the reference's build_targets
// (utils/utils.py:160-197 with wh_iou, utils.py:99-121) for every YOLO layer of a model in ONE launch.
//
// The reference runs, per layer, about thirty tiny torch kernels (stack / max / mask / index / floor / log ...) over the
// (nt, 6) target table [image, class, x, y, w, h]; here one CTA per layer walks the table once:
//   gwh = wh * (nx, ny)                                   utils.py:170
//   iou_k = wh_iou(anchor_vec[k], gwh), best = first max  utils.py:172-173 (torch.max over the stacked anchors)
//   keep  = best > iou_thres, order preserved             utils.py:176-179 (boolean-mask indexing keeps the order)
//   b, c = long(image), long(class); gxy = xy * (nx, ny); gi, gj = long(gxy)       utils.py:182-184
//   txy = gxy - floor(gxy); twh = log(gwh / anchor_vec[a]); tcls = c               utils.py:188-194
// Every step rounds to fp32 on its own like the torch op it replaces (no FMA contraction); the only inexact function is
// logf (CUDA: 1 ulp; torch CPU: its vectorised log), which the parity test compares at 1e-6 relative.
#include "common.cuh"

namespace yb {

constexpr int kTgThreads = 256;

struct TargetLayerDev {
    int nx, ny, na;
    float av[YOLO_B200_MAX_ANCHORS][2];
    long long *b, *a, *gj, *gi, *tcls;
    float *txy, *twh;
};
struct TargetParams {
    TargetLayerDev layer[YOLO_B200_MAX_SCALES];
    const float* targets;
    int nt;
    float iou_thres;
    int32_t* count;
};

__global__ void __launch_bounds__(kTgThreads)
build_targets_kernel(const __grid_constant__ TargetParams P) {
    const TargetLayerDev& L = P.layer[blockIdx.x];
    __shared__ int warp_tot[kTgThreads / 32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const float fnx = (float)L.nx, fny = (float)L.ny;
    for (int i0 = 0; i0 < P.nt; i0 += kTgThreads) {
        const int i = i0 + tid;
        bool keep = false;
        int best_a = 0;
        float t[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gw = 0.f, gh = 0.f;
        if (i < P.nt) {
#pragma unroll
            for (int c = 0; c < 6; ++c) t[c] = P.targets[(size_t)i * 6 + c];
            gw = __fmul_rn(t[4], fnx);
            gh = __fmul_rn(t[5], fny);
            float best = 0.f;
            for (int k = 0; k < L.na; ++k) {                      // utils.py:109-121
                const float w1 = L.av[k][0], h1 = L.av[k][1];
                const float inter = __fmul_rn(fminf(w1, gw), fminf(h1, gh));
                const float uni = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, h1), 1e-16f), __fmul_rn(gw, gh)), inter);
                const float iou = __fdiv_rn(inter, uni);
                if (k == 0 || iou > best || (iou != iou && best == best)) { best = iou; best_a = k; }   // first max; NaN wins like torch.max
            }
            keep = best > P.iou_thres;
        }
        // order-preserving compaction of this chunk behind the rows kept so far
        const unsigned m = __ballot_sync(kFull, keep);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += warp_tot[w];
        if (keep) {
            const int o = before + __popc(m & ((1u << lane) - 1u));
            const float gx = __fmul_rn(t[2], fnx), gy = __fmul_rn(t[3], fny);
            L.b[o] = (long long)t[0];                              // .long(): truncation towards zero
            L.tcls[o] = (long long)t[1];
            L.a[o] = best_a;
            L.gi[o] = (long long)gx;
            L.gj[o] = (long long)gy;
            L.txy[2 * o] = __fsub_rn(gx, floorf(gx));
            L.txy[2 * o + 1] = __fsub_rn(gy, floorf(gy));
            L.twh[2 * o] = logf(__fdiv_rn(gw, L.av[best_a][0]));
            L.twh[2 * o + 1] = logf(__fdiv_rn(gh, L.av[best_a][1]));
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < kTgThreads / 32; ++w) tot += warp_tot[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) P.count[blockIdx.x] = s_base;
}

}  // namespace yb

extern "C" int yolo_b200_build_targets(const float* targets, int nt, const yolo_b200_target_layer* layers, int n_layers,
                                       float iou_thres, int32_t* count, yolo_b200_stream_t stream) {
    if (!layers || !count || (!targets && nt > 0)) return YOLO_B200_E_NULL;
    if (nt < 0 || n_layers < 1 || n_layers > YOLO_B200_MAX_SCALES) return YOLO_B200_E_RANGE;
    if ((uintptr_t)targets & 3u) return YOLO_B200_E_ALIGN;
    yb::TargetParams P{};
    for (int l = 0; l < n_layers; ++l) {
        const yolo_b200_target_layer& s = layers[l];
        if (s.nx < 1 || s.ny < 1 || s.na < 1 || s.na > YOLO_B200_MAX_ANCHORS) return YOLO_B200_E_RANGE;
        if (nt > 0 && (!s.b || !s.a || !s.gj || !s.gi || !s.tcls || !s.txy || !s.twh)) return YOLO_B200_E_NULL;
        yb::TargetLayerDev& d = P.layer[l];
        d.nx = s.nx; d.ny = s.ny; d.na = s.na;
        for (int k = 0; k < YOLO_B200_MAX_ANCHORS; ++k) { d.av[k][0] = s.anchor_vec[k][0]; d.av[k][1] = s.anchor_vec[k][1]; }
        d.b = (long long*)s.b; d.a = (long long*)s.a; d.gj = (long long*)s.gj; d.gi = (long long*)s.gi; d.tcls = (long long*)s.tcls;
        d.txy = s.txy; d.twh = s.twh;
    }
    P.targets = targets; P.nt = nt; P.iou_thres = iou_thres; P.count = count;
    yb::build_targets_kernel<<<n_layers, yb::kTgThreads, 0, stream>>>(P);
    return (int)cudaGetLastError();
}
