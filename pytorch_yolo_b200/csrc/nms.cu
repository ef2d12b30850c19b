// Segmented NMS kernels of the yolo_b200 hot path (sm_100a).
//
// Replaces the per-image / per-class Python loops of the reference's non_max_suppression
// (utils/utils.py:237-291) with three launches that work on all images at once:
//
//   bucket_by_class_kernel   one CTA per image: class histogram -> segment offsets, scatter of
//                            64-bit sort keys into class buckets (utils.py:241-242); single-box
//                            classes are emitted here, untouched (utils.py:244-246)
//   nms_segment_kernel       one CTA per (image, class) segment, persistent over a work list:
//                            order by (score desc, row asc) keeping the first max_per_class
//                            (utils.py:237, 247-250), IoU suppression bitmask in 64-box tiles,
//                            warp-sequential greedy sweep and the score-weighted MERGE box
//                            (utils.py:266-275 with bbox_iou utils.py:63-96)
//   nms_finalize_kernel      one CTA per image: order kept rows by (score desc, class asc, in-class
//                            order) and write the (n, 7) result (utils.py:289-291)
//
// Tie rule (documented): equal scores keep ascending anchor-row order.  Every key
// carries the anchor row, so keys are unique and the (unstable) bitonic networks used here give a
// deterministic result that does not depend on the order compaction produced.
#include "common.cuh"

namespace yb {

constexpr int kMaxPerClassLimit = 128;  // suppression mask = 2 x 64-bit tiles per box
constexpr int kSegThreads = 128;
constexpr int kSegSort = 1024;          // elements sorted per pass in a segment CTA
constexpr int kBucketThreads = 256;
constexpr int kFinalThreads = 512;
constexpr int kFinalSmemKeys = 8192;    // 64 KB of keys in shared memory, else the global fallback

struct NmsParams {
    const yolo_b200_box* cand_box;
    const yolo_b200_meta* cand_meta;
    const int32_t* count;
    int batch, cap, nc, mpc, stage_cap, out_cap;
    float nms_thres;
    // workspace
    unsigned long long* bucket_key;   // [batch*cap]  (~score_bits << 32) | row
    uint32_t* bucket_slot;            // [batch*cap]  candidate slot of the key
    int32_t* seg_off;                 // [batch*(nc+1)] start of every class bucket
    int32_t* stage_off;               // [batch*(nc+1)] start of every class in the staging rows (lengths capped at mpc)
    int32_t* kept_count;              // [batch*nc]
    int32_t* work_list;               // [batch*nc] segments with >= 2 boxes
    int32_t* work_count;              // [1]
    float4* stage;                    // [batch*stage_cap*2] kept rows: (x1,y1,x2,y2) (score,cls_conf,row,-)
    unsigned long long* final_keys;   // [batch*stage_cap] only used when an image keeps > kFinalSmemKeys rows
    // outputs
    float* out;
    int32_t* out_row;
    int32_t* out_count;
};

// Sort key for "score descending": monotone map of the float bits (handles negative scores a caller may
// feed through compact_from_dense), inverted.  -0.0 is folded into +0.0 (they compare equal in the reference).
__device__ __forceinline__ uint32_t score_key_desc(float s) {
    if (s == 0.0f) s = 0.0f;
    const uint32_t u = __float_as_uint(s);
    const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}

// In-place ascending bitonic sort of n (any n) elements; positions >= n act as +inf and are never
// touched (normalised network: every comparator moves the minimum to the lower index).
template <bool HAS_PAYLOAD, int THREADS>
__device__ __forceinline__ void bitonic_sort(unsigned long long* key, uint32_t* payload, int n) {
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        const int half = k >> 1;
        for (int t = threadIdx.x; t * 2 < n + half; t += THREADS) {       // mirror step
            const int blk = t / half, off = t - blk * half;
            const int i = blk * k + off, l = blk * k + k - 1 - off;
            if (l < n) {
                const unsigned long long a = key[i], b = key[l];
                if (a > b) {
                    key[i] = b; key[l] = a;
                    if (HAS_PAYLOAD) { const uint32_t pa = payload[i]; payload[i] = payload[l]; payload[l] = pa; }
                }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t * 2 < n + j; t += THREADS) {
                const int i = 2 * j * (t / j) + (t % j), l = i + j;
                if (l < n) {
                    const unsigned long long a = key[i], b = key[l];
                    if (a > b) {
                        key[i] = b; key[l] = a;
                        if (HAS_PAYLOAD) { const uint32_t pa = payload[i]; payload[i] = payload[l]; payload[l] = pa; }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Exclusive scan of n ints in shared memory by warp 0 (n is a few hundred at most in practice).
// in[] -> out[0..n], out[n] = total.  Must be called by all threads of warp 0 only.
__device__ __forceinline__ void warp_exclusive_scan(const int* in, int* out, int n) {
    const int lane = threadIdx.x & 31;
    const int per = (n + 31) / 32;
    const int lo = min(n, lane * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += in[i];
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += o;
    }
    int run = incl - sum;
    for (int i = lo; i < hi; ++i) { const int v = in[i]; out[i] = run; run += v; }
    if (lane == 31) out[n] = incl;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBucketThreads)
bucket_by_class_kernel(const __grid_constant__ NmsParams P) {
    extern __shared__ int sm_i[];
    const int nc = P.nc;
    int* hist = sm_i;                 // [nc]
    int* off = hist + nc;             // [nc+1] bucket offsets, then reused as scatter cursors
    int* capped = off + nc + 1;       // [nc]
    int* soff = capped + nc;          // [nc+1] staging offsets
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(P.count[b], P.cap);
    const yolo_b200_meta* meta = P.cand_meta + (size_t)b * P.cap;

    for (int c = tid; c < nc; c += kBucketThreads) hist[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kBucketThreads) atomicAdd(&hist[meta[i].cls], 1);
    __syncthreads();
    for (int c = tid; c < nc; c += kBucketThreads) capped[c] = min(hist[c], P.mpc);
    __syncthreads();
    if (tid < 32) {
        warp_exclusive_scan(hist, off, nc);
        warp_exclusive_scan(capped, soff, nc);
    }
    __syncthreads();
    int32_t* g_seg = P.seg_off + (size_t)b * (nc + 1);
    int32_t* g_stage = P.stage_off + (size_t)b * (nc + 1);
    for (int c = tid; c <= nc; c += kBucketThreads) { g_seg[c] = off[c]; g_stage[c] = soff[c]; }
    // work list: classes with >= 2 boxes (one atomic per warp); empty / single classes are final here
    for (int c0 = 0; c0 < nc; c0 += kBucketThreads) {
        const int c = c0 + tid;
        const int len = c < nc ? hist[c] : 0;
        if (c < nc && len < 2) P.kept_count[(size_t)b * nc + c] = len;
        const unsigned need = __ballot_sync(kFull, len >= 2);
        if (need) {
            const int lane = tid & 31, leader = __ffs(need) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(P.work_count, __popc(need));
            base = __shfl_sync(kFull, base, leader);
            if (len >= 2) P.work_list[base + __popc(need & ((1u << lane) - 1u))] = b * nc + c;
        }
    }
    __syncthreads();      // hist/off fully consumed above; off[] now becomes the scatter cursor
    unsigned long long* bkey = P.bucket_key + (size_t)b * P.cap;
    uint32_t* bslot = P.bucket_slot + (size_t)b * P.cap;
    for (int i = tid; i < n; i += kBucketThreads) {
        const int4 mt = reinterpret_cast<const int4*>(meta)[i];   // score, cls_conf, cls, row
        const int c = mt.z;
        if (hist[c] == 1) {
            // single box of its class: emitted as is (utils.py:244-246)
            const float4 bx = reinterpret_cast<const float4*>(P.cand_box)[(size_t)b * P.cap + i];
            float4* st = P.stage + ((size_t)b * P.stage_cap + soff[c]) * 2;
            st[0] = bx;
            st[1] = make_float4(__int_as_float(mt.x), __int_as_float(mt.y), __int_as_float(mt.w), 0.f);
        } else {
            const int pos = atomicAdd(&off[c], 1);
            bkey[pos] = ((unsigned long long)score_key_desc(__int_as_float(mt.x)) << 32) | (uint32_t)mt.w;
            bslot[pos] = (uint32_t)i;
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSegThreads)
nms_segment_kernel(const __grid_constant__ NmsParams P) {
    __shared__ unsigned long long skey[kSegSort];
    __shared__ uint32_t sslot[kSegSort];
    __shared__ float4 sbox[kMaxPerClassLimit];
    __shared__ float sscore[kMaxPerClassLimit];
    __shared__ unsigned long long smask[kMaxPerClassLimit][2];
    __shared__ unsigned long long sclu[kMaxPerClassLimit][2];
    __shared__ int skept[kMaxPerClassLimit];
    __shared__ int s_nkept;
    const int tid = threadIdx.x;
    const int n_work = *P.work_count;
    const int mpc = P.mpc;

    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int item = P.work_list[wi];
        const int b = item / P.nc, c = item - b * P.nc;
        const int s0 = P.seg_off[(size_t)b * (P.nc + 1) + c];
        const int n = P.seg_off[(size_t)b * (P.nc + 1) + c + 1] - s0;
        const unsigned long long* gkey = P.bucket_key + (size_t)b * P.cap + s0;
        const uint32_t* gslot = P.bucket_slot + (size_t)b * P.cap + s0;

        // ---- order by (score desc, row asc), keep the first mpc (utils.py:237, 247-250).
        // Buckets larger than the shared array are streamed: sort(carry + chunk), keep the first mpc.
        int carry = 0;
        for (int pos = 0; pos < n;) {
            const int take = min(n - pos, kSegSort - carry);
            for (int i = tid; i < take; i += kSegThreads) { skey[carry + i] = gkey[pos + i]; sslot[carry + i] = gslot[pos + i]; }
            __syncthreads();
            bitonic_sort<true, kSegThreads>(skey, sslot, carry + take);
            carry = min(carry + take, mpc);
            pos += take;
        }
        const int m = carry;

        // ---- gather the surviving boxes
        if (tid < m) {
            sbox[tid] = reinterpret_cast<const float4*>(P.cand_box)[(size_t)b * P.cap + sslot[tid]];
            sscore[tid] = P.cand_meta[(size_t)b * P.cap + sslot[tid]].score;
        }
        __syncthreads();

        // ---- suppression bitmask, 64-box tiles: bit j of smask[i][t] <=> IoU(box i, box 64t+j) > thr, j >= i
        const int ntile = (m + 63) >> 6;
        for (int w = tid; w < m * 2; w += kSegThreads) {
            const int i = w >> 1, tl = w & 1;
            unsigned long long bits = 0;
            if (tl < ntile) {
                const float4 bi4 = sbox[i];
                const yolo_b200_box bi = {bi4.x, bi4.y, bi4.z, bi4.w};
                const int j0 = max(tl << 6, i), j1 = min(m, (tl << 6) + 64);
                for (int j = j0; j < j1; ++j) {
                    const float4 bj4 = sbox[j];
                    const yolo_b200_box bj = {bj4.x, bj4.y, bj4.z, bj4.w};
                    if (iou_ref(bi, bj) > P.nms_thres) bits |= 1ull << (j & 63);   // utils.py:271 strict >
                }
            }
            smask[i][tl] = bits;
        }
        __syncthreads();

        // ---- greedy sweep, one warp, lane-uniform (utils.py:266-275)
        if (tid < 32) {
            unsigned long long a0 = m >= 64 ? ~0ull : ((1ull << m) - 1ull);
            unsigned long long a1 = m > 64 ? ((m >= 128) ? ~0ull : ((1ull << (m - 64)) - 1ull)) : 0ull;
            int nk = 0;
            while (a0 | a1) {
                const int i = a0 ? (__ffsll((long long)a0) - 1) : (64 + __ffsll((long long)a1) - 1);
                if (__popcll(a0) + __popcll(a1) == 1) {            // last survivor: emitted unmerged (utils.py:268-270)
                    if (tid == 0) { skept[nk] = i; sclu[nk][0] = 0; sclu[nk][1] = 0; }
                    ++nk;
                    break;
                }
                const unsigned long long c0 = smask[i][0] & a0, c1 = smask[i][1] & a1;
                if (tid == 0) { skept[nk] = i; sclu[nk][0] = c0; sclu[nk][1] = c1; }
                ++nk;
                a0 &= ~c0; a1 &= ~c1;
                if (i < 64) a0 &= ~(1ull << i); else a1 &= ~(1ull << (i - 64));
            }
            if (tid == 0) s_nkept = nk;
        }
        __syncthreads();

        // ---- MERGE box of every kept detection: sum_j s_j*box_j / sum_j s_j over its cluster, in order
        const int nk = s_nkept;
        if (tid < nk) {
            const int i = skept[tid];
            unsigned long long c0 = sclu[tid][0], c1 = sclu[tid][1];
            float4 o = sbox[i];
            if (c0 | c1) {
                float sw = 0.f, sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f;
                for (int half = 0; half < 2; ++half) {
                    unsigned long long bits = half ? c1 : c0;
                    while (bits) {
                        const int j = (half << 6) + __ffsll((long long)bits) - 1;
                        bits &= bits - 1;
                        const float s = sscore[j];
                        const float4 bj = sbox[j];
                        sw = __fadd_rn(sw, s);
                        sx1 = __fadd_rn(sx1, __fmul_rn(s, bj.x));
                        sy1 = __fadd_rn(sy1, __fmul_rn(s, bj.y));
                        sx2 = __fadd_rn(sx2, __fmul_rn(s, bj.z));
                        sy2 = __fadd_rn(sy2, __fmul_rn(s, bj.w));
                    }
                }
                o = make_float4(__fdiv_rn(sx1, sw), __fdiv_rn(sy1, sw), __fdiv_rn(sx2, sw), __fdiv_rn(sy2, sw));
            }
            const size_t cslot = (size_t)b * P.cap + sslot[i];
            const float cls_conf = P.cand_meta[cslot].cls_conf;
            const int row = (int)(uint32_t)skey[i];
            float4* st = P.stage + ((size_t)b * P.stage_cap + P.stage_off[(size_t)b * (P.nc + 1) + c] + tid) * 2;
            st[0] = o;
            st[1] = make_float4(sscore[i], cls_conf, __int_as_float(row), 0.f);
        }
        if (tid == 0) P.kept_count[(size_t)b * P.nc + c] = nk;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinalThreads)
nms_finalize_kernel(const __grid_constant__ NmsParams P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm_raw);   // [kFinalSmemKeys]
    int* kc = reinterpret_cast<int*>(skeys + kFinalSmemKeys);                       // [nc]
    int* pre = kc + P.nc;                                                           // [nc+1]
    const int b = blockIdx.x, tid = threadIdx.x, nc = P.nc;

    for (int c = tid; c < nc; c += kFinalThreads) kc[c] = P.kept_count[(size_t)b * nc + c];
    __syncthreads();
    if (tid < 32) warp_exclusive_scan(kc, pre, nc);
    __syncthreads();
    const int n_out = pre[nc];
    if (tid == 0) P.out_count[b] = n_out;
    if (n_out == 0) return;

    unsigned long long* keys = (n_out <= kFinalSmemKeys) ? skeys : (P.final_keys + (size_t)b * P.stage_cap);
    const int32_t* soff = P.stage_off + (size_t)b * (nc + 1);
    const float4* stage = P.stage + (size_t)b * P.stage_cap * 2;
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = warp; c < nc; c += kFinalThreads / 32) {
        const int k = kc[c];
        for (int r = lane; r < k; r += 32) {
            const float score = stage[(size_t)(soff[c] + r) * 2 + 1].x;
            // score desc, then class asc, then in-class order (utils.py:291 with the stable tie rule)
            keys[pre[c] + r] = ((unsigned long long)score_key_desc(score) << 32) | ((unsigned)c << 16) | (unsigned)r;
        }
    }
    __syncthreads();
    if (n_out <= kFinalSmemKeys) bitonic_sort<false, kFinalThreads>(skeys, nullptr, n_out);
    else                         bitonic_sort<false, kFinalThreads>(keys, nullptr, n_out);

    float* out = P.out + (size_t)b * P.out_cap * YOLO_B200_DET_COLS;
    int32_t* out_row = P.out_row + (size_t)b * P.out_cap;
    for (int e = tid; e < n_out * 8; e += kFinalThreads) {
        const int i = e >> 3, col = e & 7;
        const unsigned long long key = keys[i];
        const int c = (int)((key >> 16) & 0xffffu), r = (int)(key & 0xffffu);
        const float* src = reinterpret_cast<const float*>(stage + (size_t)(soff[c] + r) * 2);
        if (col < 6)       out[(size_t)i * YOLO_B200_DET_COLS + col] = src[col];
        else if (col == 6) out[(size_t)i * YOLO_B200_DET_COLS + 6] = (float)c;
        else               out_row[i] = __float_as_int(src[6]);
    }
}

}  // namespace yb

// ================================================================================================
using namespace yb;

namespace {
struct WsLayout {
    size_t bucket_key, bucket_slot, seg_off, stage_off, kept_count, work_list, work_count, stage, final_keys, total;
};
inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
WsLayout ws_layout(int batch, int cap, int nc, int mpc) {
    WsLayout L{};
    const size_t stage_cap = (size_t)((long long)cap < (long long)nc * mpc ? cap : nc * mpc);
    size_t o = 0;
    L.bucket_key = o;  o = align_up(o + (size_t)batch * cap * 8);
    L.bucket_slot = o; o = align_up(o + (size_t)batch * cap * 4);
    L.seg_off = o;     o = align_up(o + (size_t)batch * (nc + 1) * 4);
    L.stage_off = o;   o = align_up(o + (size_t)batch * (nc + 1) * 4);
    L.kept_count = o;  o = align_up(o + (size_t)batch * nc * 4);
    L.work_list = o;   o = align_up(o + (size_t)batch * nc * 4);
    L.work_count = o;  o = align_up(o + 4);
    L.stage = o;       o = align_up(o + (size_t)batch * stage_cap * 32);
    L.final_keys = o;  o = align_up(o + (size_t)batch * stage_cap * 8);
    L.total = o;
    return L;
}
}  // namespace

extern "C" size_t yolo_b200_nms_workspace_bytes(int batch, int cap_per_img, int nc, int max_per_class) {
    if (batch < 0 || cap_per_img < 1 || nc < 1 || max_per_class < 1) return 0;
    return ws_layout(batch, cap_per_img, nc, max_per_class).total;
}

extern "C" int yolo_b200_nms(const yolo_b200_box* cand_box, const yolo_b200_meta* cand_meta, const int32_t* count,
                             int batch, int cap_per_img, int nc, float nms_thres, int max_per_class,
                             float* out, int32_t* out_row, int out_cap, int32_t* out_count,
                             void* workspace, size_t workspace_bytes, yolo_b200_stream_t stream) {
    if (!cand_box || !cand_meta || !count || !out || !out_row || !out_count || !workspace) return YOLO_B200_E_NULL;
    if (batch < 0 || cap_per_img < 1 || nc < 1 || nc > YOLO_B200_MAX_CLASSES || max_per_class < 1 ||
        max_per_class > kMaxPerClassLimit)
        return YOLO_B200_E_RANGE;
    // nms_thres >= 1 never terminates in the reference (self-IoU 1.0 is not > 1); NaN likewise
    if (!(nms_thres < 1.0f)) return YOLO_B200_E_RANGE;
    if ((((uintptr_t)cand_box) | ((uintptr_t)cand_meta)) & 15u) return YOLO_B200_E_ALIGN;
    if ((uintptr_t)workspace & 255u) return YOLO_B200_E_ALIGN;
    const WsLayout L = ws_layout(batch, cap_per_img, nc, max_per_class);
    if (workspace_bytes < L.total) return YOLO_B200_E_WORKSPACE;
    const int stage_cap = (long long)cap_per_img < (long long)nc * max_per_class ? cap_per_img : nc * max_per_class;
    if (out_cap < stage_cap) return YOLO_B200_E_RANGE;
    if (batch == 0) return 0;

    unsigned char* ws = static_cast<unsigned char*>(workspace);
    NmsParams P{};
    P.cand_box = cand_box; P.cand_meta = cand_meta; P.count = count;
    P.batch = batch; P.cap = cap_per_img; P.nc = nc; P.mpc = max_per_class; P.stage_cap = stage_cap; P.out_cap = out_cap;
    P.nms_thres = nms_thres;
    P.bucket_key = reinterpret_cast<unsigned long long*>(ws + L.bucket_key);
    P.bucket_slot = reinterpret_cast<uint32_t*>(ws + L.bucket_slot);
    P.seg_off = reinterpret_cast<int32_t*>(ws + L.seg_off);
    P.stage_off = reinterpret_cast<int32_t*>(ws + L.stage_off);
    P.kept_count = reinterpret_cast<int32_t*>(ws + L.kept_count);
    P.work_list = reinterpret_cast<int32_t*>(ws + L.work_list);
    P.work_count = reinterpret_cast<int32_t*>(ws + L.work_count);
    P.stage = reinterpret_cast<float4*>(ws + L.stage);
    P.final_keys = reinterpret_cast<unsigned long long*>(ws + L.final_keys);
    P.out = out; P.out_row = out_row; P.out_count = out_count;

    cudaError_t e;
    if ((e = cudaMemsetAsync(P.work_count, 0, sizeof(int32_t), stream)) != cudaSuccess) return (int)e;
    const size_t bucket_smem = (size_t)(4 * nc + 2) * sizeof(int);
    if (bucket_smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(bucket_by_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bucket_smem)) != cudaSuccess)
        return (int)e;
    bucket_by_class_kernel<<<batch, kBucketThreads, bucket_smem, stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long segs = (long long)batch * nc;
    const int seg_grid = (int)(segs < (long long)sms * 8 ? segs : (long long)sms * 8);
    nms_segment_kernel<<<seg_grid, kSegThreads, 0, stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    const size_t final_smem = (size_t)kFinalSmemKeys * 8 + (size_t)(2 * nc + 1) * sizeof(int);
    if ((e = cudaFuncSetAttribute(nms_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)final_smem)) != cudaSuccess)
        return (int)e;
    nms_finalize_kernel<<<batch, kFinalThreads, final_smem, stream>>>(P);
    return (int)cudaGetLastError();
}

extern "C" int yolo_b200_abi_version(void) { return YOLO_B200_ABI_VERSION; }

extern "C" const char* yolo_b200_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case YOLO_B200_E_NULL: return "null pointer argument";
        case YOLO_B200_E_RANGE: return "argument out of range";
        case YOLO_B200_E_ALIGN: return "pointer not aligned as documented";
        case YOLO_B200_E_WORKSPACE: return "workspace too small";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
