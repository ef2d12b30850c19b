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
constexpr int kSegThreads = 128;        // CTA size of the segment kernel (4 warps)
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegSort = 1024;          // elements sorted per pass by a CTA working on a big segment
constexpr int kSmallSeg = 32;           // segments up to this size are handled by one warp, one box per lane
constexpr int kPairSeg = 64;            // ... up to this size by one warp with two boxes per lane; bigger ones by a CTA
constexpr int kBucketThreads = 256;
constexpr int kBucketRegs = 8;          // candidate records a bucket thread keeps in registers between its two passes
constexpr int kFinalThreadsBig = 1024;   // finalize CTA size when an image can stage many rows
constexpr int kFinalThreadsSmall = 256;  // ... and when it cannot (small per-image capacity, usually large batches)
constexpr int kFinalSmemKeys = 8192;    // 64 KB of keys in shared memory, else the global fallback

struct NmsParams {
    const yolo_b200_box* cand_box;
    const yolo_b200_meta* cand_meta;
    const int32_t* count;
    int batch, cap, nc, mpc, stage_cap, out_cap;
    int big_ctas;                     // the LAST big_ctas CTAs of the segment kernel serve the big-segment list
    int final_smem_keys;              // keys the finalize kernel can hold in shared memory (generic path)
    float nms_thres;
    // workspace
    unsigned long long* bucket_key;   // [batch*cap]  (score-descending key << 32) | row
    uint32_t* bucket_slot;            // [batch*cap]  candidate slot of the key
    int32_t* seg_off;                 // [batch*(nc+1)] start of every class bucket
    int32_t* stage_off;               // [batch*(nc+1)] start of every class in the staging rows (lengths capped at mpc)
    int32_t* work_big;                // [batch*nc] segments with more than kSmallSeg boxes (small ones need no list)
    int32_t* work_count;              // [2] number of big segments | finalize CTAs that have finished (self-resetting)
    float4* stage;                    // [batch*stage_cap*2] staged rows: (x1,y1,x2,y2) (score,cls_conf,row,cls); score NaN = not kept
    unsigned long long* final_keys;   // [batch*stage_cap] only used when an image keeps > kFinalSmemKeys rows
    // outputs
    float* out;
    int32_t* out_row;
    int32_t* out_count;
    // optional completion stamp (multi-GPU gather): once every result row of this call has been stored, the last
    // finalize CTA writes ++*step_seq to *step_stamp (release, system scope; step_stamp may be peer memory)
    int32_t* step_seq;
    int32_t* step_stamp;
};

// Sort key for "score descending": monotone map of the float bits (handles negative scores a caller may
// feed through compact_from_dense), inverted.  -0.0 is folded into +0.0 (they compare equal in the reference).
__device__ __forceinline__ uint32_t score_key_desc(float s) {
    if (s == 0.0f) s = 0.0f;
    const uint32_t u = __float_as_uint(s);
    const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ float box_area(const float4& b) {      // utils.py:94 (x2-x1)*(y2-y1)
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// Exactly "bbox_iou(a, b) > thr" (utils.py:89-96, 271) without paying for the IEEE division on every pair.
// Branch-free main path: outside a relative band of 2^-20 around thr * union the rounded quotient is provably on
// the same side of thr as the real one, so the comparison inter <> thr*union decides; only pairs inside the band
// (or with a degenerate union / threshold) execute the division.  area_a_eps = area_a + 1e-16f (utils.py:93).
__device__ __forceinline__ bool iou_gt(const float4& a, float area_a_eps, const float4& b, float area_b, float thr) {
    const float ix = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float iy = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float inter = __fmul_rn(fmaxf(ix, 0.0f), fmaxf(iy, 0.0f));
    const float uni = __fsub_rn(__fadd_rn(area_a_eps, area_b), inter);
    const float p = __fmul_rn(thr, uni);
    const bool yes = inter > __fmul_rn(p, 1.000001f);
    const bool no = inter < __fmul_rn(p, 0.999999f);
    const bool sure = (yes || no) && uni > 0.0f && uni < 3.0e38f && thr > 1e-30f;
    if (sure) return yes;
    return __fdiv_rn(inter, uni) > thr;
}

// In-place ascending bitonic sort of n (any n) elements in shared or global memory; positions >= n act as
// +inf and are never touched (normalised network: every comparator moves the minimum to the lower index).
template <bool HAS_PAYLOAD, int THREADS>
__device__ __forceinline__ void bitonic_sort(unsigned long long* key, uint32_t* payload, int n) {
    auto cmpswap = [&](int i, int l) {
        const unsigned long long a = key[i], b = key[l];
        if (a > b) {
            key[i] = b; key[l] = a;
            if (HAS_PAYLOAD) { const uint32_t pa = payload[i]; payload[i] = payload[l]; payload[l] = pa; }
        }
    };
    for (int lk = 1; (1 << (lk - 1)) < n; ++lk) {
        const int k = 1 << lk, half = k >> 1;
        for (int t = threadIdx.x; t * 2 < n + half; t += THREADS) {       // mirror step
            const int blk = t >> (lk - 1), off = t & (half - 1);
            const int i = (blk << lk) + off, l = (blk << lk) + k - 1 - off;
            if (l < n) cmpswap(i, l);
        }
        __syncthreads();
        for (int lj = lk - 2; lj >= 0; --lj) {
            const int j = 1 << lj;
            for (int t = threadIdx.x; t * 2 < n + j; t += THREADS) {
                const int i = ((t >> lj) << (lj + 1)) + (t & (j - 1)), l = i + j;
                if (l < n) cmpswap(i, l);
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// One CTA per image.  Pass 1 builds the class histogram (records stay in registers), warp 0 turns it into
// bucket / staging offsets and into two work lists, pass 2 scatters 64-bit keys into the class buckets.
__global__ void __launch_bounds__(kBucketThreads)
bucket_by_class_kernel(const __grid_constant__ NmsParams P) {
    extern __shared__ int sm_i[];
    const int nc = P.nc;
    int* hist = sm_i;                 // [nc]
    int* cur = hist + nc;             // [nc]   scatter cursors (start at the bucket offset)
    int* soff = cur + nc;             // [nc+1] staging offsets
    int* rank = soff + nc + 1;        // [nc]   position of the class inside the big-segment work list
    __shared__ int s_base;
    __shared__ int s_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = min(P.count[b], P.cap);
    const int4* meta4 = reinterpret_cast<const int4*>(P.cand_meta) + (size_t)b * P.cap;

    for (int c = tid; c < nc; c += kBucketThreads) hist[c] = 0;
    __syncthreads();
    int4 mt[kBucketRegs];
#pragma unroll
    for (int k = 0; k < kBucketRegs; ++k) {
        const int i = tid + k * kBucketThreads;
        if (i < n) { mt[k] = meta4[i]; if ((unsigned)mt[k].z < (unsigned)nc) atomicAdd(&hist[mt[k].z], 1); }
    }
    for (int i = tid + kBucketRegs * kBucketThreads; i < n; i += kBucketThreads) {
        const int c = meta4[i].z;
        if ((unsigned)c < (unsigned)nc) atomicAdd(&hist[c], 1);       // records with a class id outside [0, nc) are ignored
    }
    __syncthreads();

    if (tid < 32) {
        int run_off = 0, run_soff = 0, n_big = 0;
        for (int c0 = 0; c0 < nc; c0 += 32) {
            const int c = c0 + lane;
            const int len = c < nc ? hist[c] : 0;
            const int cp = min(len, P.mpc);
            const int il = warp_incl_scan(len, lane), ic = warp_incl_scan(cp, lane);
            const bool big = len > kPairSeg;
            const unsigned bb = __ballot_sync(kFull, big);
            if (c < nc) {
                cur[c] = run_off + il - len;
                soff[c] = run_soff + ic - cp;
                rank[c] = n_big + __popc(bb & ((1u << lane) - 1u));
            }
            n_big += __popc(bb);
            run_off += __shfl_sync(kFull, il, 31);
            run_soff += __shfl_sync(kFull, ic, 31);
        }
        if (lane == 0) {
            soff[nc] = run_soff;
            s_total = run_off;
            s_base = n_big ? atomicAdd(P.work_count, n_big) : 0;      // no global atomic unless the image has big segments
        }
    }
    __syncthreads();

    int32_t* g_seg = P.seg_off + (size_t)b * (nc + 1);
    int32_t* g_stage = P.stage_off + (size_t)b * (nc + 1);
    for (int c = tid; c < nc; c += kBucketThreads) {
        const int len = hist[c];
        g_seg[c] = cur[c];
        g_stage[c] = soff[c];
        if (len > kPairSeg) P.work_big[s_base + rank[c]] = b * nc + c;
    }
    if (tid == 0) { g_seg[nc] = s_total; g_stage[nc] = soff[nc]; }
    __syncthreads();          // cur[] is read above and bumped below

    unsigned long long* bkey = P.bucket_key + (size_t)b * P.cap;
    uint32_t* bslot = P.bucket_slot + (size_t)b * P.cap;
    auto place = [&](const int4& m, int i) {
        const int c = m.z;
        if ((unsigned)c >= (unsigned)nc) return;
        if (hist[c] == 1) {
            // single box of its class: emitted as is (utils.py:244-246)
            const float4 bx = reinterpret_cast<const float4*>(P.cand_box)[(size_t)b * P.cap + i];
            float4* st = P.stage + ((size_t)b * P.stage_cap + soff[c]) * 2;
            st[0] = bx;
            st[1] = make_float4(__int_as_float(m.x), __int_as_float(m.y), __int_as_float(m.w), (float)c);
        } else {
            const int pos = atomicAdd(&cur[c], 1);
            bkey[pos] = ((unsigned long long)score_key_desc(__int_as_float(m.x)) << 32) | (uint32_t)m.w;
            bslot[pos] = (uint32_t)i;
        }
    };
#pragma unroll
    for (int k = 0; k < kBucketRegs; ++k) {
        const int i = tid + k * kBucketThreads;
        if (i < n) place(mt[k], i);
    }
    for (int i = tid + kBucketRegs * kBucketThreads; i < n; i += kBucketThreads) place(meta4[i], i);
}

// ------------------------------------------------------------------------------------------------
// Small segments (2..32 boxes): one warp, everything in registers, no barriers.
__device__ __forceinline__ void nms_small_segment(const NmsParams& P, int b, int c, int s0, int n, int st_off, int lane) {
    unsigned long long key = ~0ull;
    uint32_t slot = 0;
    if (lane < n) {
        key = P.bucket_key[(size_t)b * P.cap + s0 + lane];
        slot = P.bucket_slot[(size_t)b * P.cap + s0 + lane];
    }
    // warp bitonic sort, ascending key = (score desc, row asc)  (utils.py:237); lanes >= n hold +inf keys, so the
    // network only has to span the next power of two >= n (warp-uniform bound)
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        if ((k >> 1) >= n) break;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(kFull, key, j);
            const uint32_t os = __shfl_xor_sync(kFull, slot, j);
            const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
            const bool swap = take_min ? (ok < key) : (ok > key);
            if (swap) { key = ok; slot = os; }
        }
    }
    n = min(n, P.mpc);                               // only the first max_per_class are considered (utils.py:247-250)
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    float score = 0.f, cls_conf = 0.f;
    if (lane < n) {
        const size_t cs = (size_t)b * P.cap + slot;
        box = reinterpret_cast<const float4*>(P.cand_box)[cs];
        const int4 m = reinterpret_cast<const int4*>(P.cand_meta)[cs];
        score = __int_as_float(m.x);
        cls_conf = __int_as_float(m.y);
    }
    const float area = box_area(box);
    const float thr = P.nms_thres;

    unsigned alive = n >= 32 ? ~0u : ((1u << n) - 1u);
    int nk = 0, kept_i = 0;
    float4 obox = box;
    while (alive) {                                              // lane-uniform greedy sweep (utils.py:266-275)
        const int i = __ffs(alive) - 1;
        float4 m4;
        if (__popc(alive) == 1) {                                // last survivor: emitted unmerged (utils.py:268-270)
            m4.x = __shfl_sync(kFull, box.x, i); m4.y = __shfl_sync(kFull, box.y, i);
            m4.z = __shfl_sync(kFull, box.z, i); m4.w = __shfl_sync(kFull, box.w, i);
            alive = 0;
        } else {
            float4 bi;
            bi.x = __shfl_sync(kFull, box.x, i); bi.y = __shfl_sync(kFull, box.y, i);
            bi.z = __shfl_sync(kFull, box.z, i); bi.w = __shfl_sync(kFull, box.w, i);
            const float ai = __fadd_rn(__shfl_sync(kFull, area, i), 1e-16f);
            const bool hit = lane >= i && lane < n && iou_gt(bi, ai, box, area, thr);     // utils.py:271
            unsigned cl = __ballot_sync(kFull, hit) & alive;
            alive &= ~cl;
            alive &= ~(1u << i);
            m4 = bi;
            if (cl) {                                            // score-weighted mean of the cluster, in order (utils.py:272-273)
                float sw = 0.f, sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f;
                while (cl) {
                    const int j = __ffs(cl) - 1;
                    cl &= cl - 1;
                    const float s = __shfl_sync(kFull, score, j);
                    sw = __fadd_rn(sw, s);
                    sx1 = __fadd_rn(sx1, __fmul_rn(s, __shfl_sync(kFull, box.x, j)));
                    sy1 = __fadd_rn(sy1, __fmul_rn(s, __shfl_sync(kFull, box.y, j)));
                    sx2 = __fadd_rn(sx2, __fmul_rn(s, __shfl_sync(kFull, box.z, j)));
                    sy2 = __fadd_rn(sy2, __fmul_rn(s, __shfl_sync(kFull, box.w, j)));
                }
                // the four IEEE divisions run as ONE warp instruction: lane c divides coordinate c
                const float num = (lane & 3) == 0 ? sx1 : (lane & 3) == 1 ? sy1 : (lane & 3) == 2 ? sx2 : sy2;
                const float q = __fdiv_rn(num, sw);
                m4 = make_float4(__shfl_sync(kFull, q, 0), __shfl_sync(kFull, q, 1), __shfl_sync(kFull, q, 2),
                                 __shfl_sync(kFull, q, 3));
            }
        }
        if (lane == nk) { obox = m4; kept_i = i; }
        ++nk;
    }
    // lane k holds the k-th kept detection: fetch its score / cls_conf / row from the lane that owns box kept_i
    const float o_score = __shfl_sync(kFull, score, kept_i);
    const float o_conf = __shfl_sync(kFull, cls_conf, kept_i);
    const int o_row = (int)__shfl_sync(kFull, (uint32_t)key, kept_i);
    if (lane < n) {                                  // staged slots beyond the kept ones are marked with a NaN score
        float4* st = P.stage + ((size_t)b * P.stage_cap + st_off + lane) * 2;
        if (lane < nk) st[0] = obox;
        st[1] = make_float4(lane < nk ? o_score : __int_as_float(0x7fc00000), o_conf, __int_as_float(o_row), (float)c);
    }
}

// Segments of 33..64 boxes: still one warp, two boxes per lane (elements lane and lane + 32).  The bitonic network
// uses shuffles for strides < 32 and a register-against-register exchange for stride 32; the greedy sweep keeps a
// 64-bit alive mask.  Kept rows are written as they are found.
__device__ __forceinline__ void nms_pair_segment(const NmsParams& P, int b, int c, int s0, int n, int st_off, int lane) {
    unsigned long long k0 = ~0ull, k1 = ~0ull;
    uint32_t sl0 = 0, sl1 = 0;
    const size_t gb = (size_t)b * P.cap + s0;
    if (lane < n) { k0 = P.bucket_key[gb + lane]; sl0 = P.bucket_slot[gb + lane]; }
    if (lane + 32 < n) { k1 = P.bucket_key[gb + lane + 32]; sl1 = P.bucket_slot[gb + lane + 32]; }
    auto xchg = [&](unsigned long long& key, uint32_t& slot, int j, bool take_min) {
        const unsigned long long ok = __shfl_xor_sync(kFull, key, j);
        const uint32_t os = __shfl_xor_sync(kFull, slot, j);
        const bool swap = take_min ? (ok < key) : (ok > key);
        if (swap) { key = ok; slot = os; }
    };
    // k = 2..32: both halves sort independently, the upper half (elements 32..63) of a 64-network runs descending
    // for k = 32 (bit 5 of the element index is set) -- standard bitonic directions with e = lane (+32)
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool low = (lane & j) == 0;
            xchg(k0, sl0, j, (((lane & k) == 0) == low));
            xchg(k1, sl1, j, ((((lane + 32) & k) == 0) == low));
        }
    }
    // k = 64: stride 32 inside the lane (ascending), then strides 16..1 ascending in both halves
    if (k0 > k1) { const unsigned long long t = k0; k0 = k1; k1 = t; const uint32_t u = sl0; sl0 = sl1; sl1 = u; }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const bool low = (lane & j) == 0;
        xchg(k0, sl0, j, low);
        xchg(k1, sl1, j, low);
    }
    const int m = min(n, P.mpc);                     // utils.py:247-250
    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
    float s_0 = 0.f, s_1 = 0.f, c_0 = 0.f, c_1 = 0.f;
    if (lane < m) {
        const size_t cs = (size_t)b * P.cap + sl0;
        b0 = reinterpret_cast<const float4*>(P.cand_box)[cs];
        const int4 mt = reinterpret_cast<const int4*>(P.cand_meta)[cs];
        s_0 = __int_as_float(mt.x); c_0 = __int_as_float(mt.y);
    }
    if (lane + 32 < m) {
        const size_t cs = (size_t)b * P.cap + sl1;
        b1 = reinterpret_cast<const float4*>(P.cand_box)[cs];
        const int4 mt = reinterpret_cast<const int4*>(P.cand_meta)[cs];
        s_1 = __int_as_float(mt.x); c_1 = __int_as_float(mt.y);
    }
    const float a0 = box_area(b0), a1 = box_area(b1);
    const int r0 = (int)(uint32_t)k0, r1 = (int)(uint32_t)k1;
    const float thr = P.nms_thres;
    unsigned al0 = m >= 32 ? ~0u : ((1u << m) - 1u);
    unsigned al1 = m >= 64 ? ~0u : (m > 32 ? ((1u << (m - 32)) - 1u) : 0u);
    float4* st = P.stage + ((size_t)b * P.stage_cap + st_off) * 2;
    int nk = 0;
    while (al0 | al1) {                                          // lane-uniform greedy sweep (utils.py:266-275)
        const bool hi = al0 == 0;
        const int l = __ffs(hi ? al1 : al0) - 1;
        const int i = l + (hi ? 32 : 0);
        const float4 bs = hi ? b1 : b0;
        const float4 bi = make_float4(__shfl_sync(kFull, bs.x, l), __shfl_sync(kFull, bs.y, l),
                                      __shfl_sync(kFull, bs.z, l), __shfl_sync(kFull, bs.w, l));
        float4 m4 = bi;
        if (__popc(al0) + __popc(al1) == 1) {                    // last survivor: emitted unmerged (utils.py:268-270)
            al0 = al1 = 0;
        } else {
            const float ai = __fadd_rn(__shfl_sync(kFull, hi ? a1 : a0, l), 1e-16f);
            unsigned c0 = 0, c1 = 0;
            if (!hi) c0 = __ballot_sync(kFull, lane >= i && lane < m && iou_gt(bi, ai, b0, a0, thr)) & al0;
            c1 = __ballot_sync(kFull, lane + 32 >= i && lane + 32 < m && iou_gt(bi, ai, b1, a1, thr)) & al1;   // utils.py:271
            al0 &= ~c0; al1 &= ~c1;
            if (hi) al1 &= ~(1u << l); else al0 &= ~(1u << l);
            if (c0 | c1) {                                       // score-weighted mean of the cluster, in order
                float sw = 0.f, sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f;
                while (c0) {
                    const int j = __ffs(c0) - 1; c0 &= c0 - 1;
                    const float sj = __shfl_sync(kFull, s_0, j);
                    sw = __fadd_rn(sw, sj);
                    sx1 = __fadd_rn(sx1, __fmul_rn(sj, __shfl_sync(kFull, b0.x, j)));
                    sy1 = __fadd_rn(sy1, __fmul_rn(sj, __shfl_sync(kFull, b0.y, j)));
                    sx2 = __fadd_rn(sx2, __fmul_rn(sj, __shfl_sync(kFull, b0.z, j)));
                    sy2 = __fadd_rn(sy2, __fmul_rn(sj, __shfl_sync(kFull, b0.w, j)));
                }
                while (c1) {
                    const int j = __ffs(c1) - 1; c1 &= c1 - 1;
                    const float sj = __shfl_sync(kFull, s_1, j);
                    sw = __fadd_rn(sw, sj);
                    sx1 = __fadd_rn(sx1, __fmul_rn(sj, __shfl_sync(kFull, b1.x, j)));
                    sy1 = __fadd_rn(sy1, __fmul_rn(sj, __shfl_sync(kFull, b1.y, j)));
                    sx2 = __fadd_rn(sx2, __fmul_rn(sj, __shfl_sync(kFull, b1.z, j)));
                    sy2 = __fadd_rn(sy2, __fmul_rn(sj, __shfl_sync(kFull, b1.w, j)));
                }
                const float num = (lane & 3) == 0 ? sx1 : (lane & 3) == 1 ? sy1 : (lane & 3) == 2 ? sx2 : sy2;
                const float q = __fdiv_rn(num, sw);
                m4 = make_float4(__shfl_sync(kFull, q, 0), __shfl_sync(kFull, q, 1), __shfl_sync(kFull, q, 2),
                                 __shfl_sync(kFull, q, 3));
            }
        }
        const float si = __shfl_sync(kFull, hi ? s_1 : s_0, l), ci = __shfl_sync(kFull, hi ? c_1 : c_0, l);
        const int rowi = __shfl_sync(kFull, hi ? r1 : r0, l);
        if (lane == 0) st[2 * nk] = m4;
        if (lane == 1) st[2 * nk + 1] = make_float4(si, ci, __int_as_float(rowi), (float)c);
        ++nk;
    }
    for (int e = nk + lane; e < m; e += 32)          // staged slots beyond the kept ones are marked with a NaN score
        st[2 * e + 1] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, (float)c);
}

// Big segments: one CTA.  Streams the bucket through a shared-memory bitonic sort keeping the best mpc,
// builds the suppression bitmask in 64-box tiles, sweeps it with one warp, merges per kept box.
struct BigSegSmem {
    unsigned long long key[kSegSort];
    uint32_t slot[kSegSort];
    float4 box[kMaxPerClassLimit];
    float area[kMaxPerClassLimit];
    float score[kMaxPerClassLimit];
    alignas(16) unsigned mask32[kMaxPerClassLimit][4];            // suppression bits of box i against boxes 32r .. 32r+31
    unsigned long long clu[kMaxPerClassLimit][2];
    int kept[kMaxPerClassLimit];
    int nkept;
};

__device__ __forceinline__ void nms_big_segment(const NmsParams& P, int item, BigSegSmem& S) {
    const int tid = threadIdx.x;
    const int mpc = P.mpc;
    const int b = item / P.nc, c = item - b * P.nc;
    const int s0 = P.seg_off[(size_t)b * (P.nc + 1) + c];
    const int n = P.seg_off[(size_t)b * (P.nc + 1) + c + 1] - s0;
    const unsigned long long* gkey = P.bucket_key + (size_t)b * P.cap + s0;
    const uint32_t* gslot = P.bucket_slot + (size_t)b * P.cap + s0;

    // ---- order by (score desc, row asc), keep the first mpc (utils.py:237, 247-250).
    int carry = 0;
    for (int pos = 0; pos < n;) {
        const int take = min(n - pos, kSegSort - carry);
        for (int i = tid; i < take; i += kSegThreads) { S.key[carry + i] = gkey[pos + i]; S.slot[carry + i] = gslot[pos + i]; }
        __syncthreads();
        bitonic_sort<true, kSegThreads>(S.key, S.slot, carry + take);
        carry = min(carry + take, mpc);
        pos += take;
    }
    const int m = carry;

    if (tid < m) {
        const size_t cs = (size_t)b * P.cap + S.slot[tid];
        const float4 bx = reinterpret_cast<const float4*>(P.cand_box)[cs];
        S.box[tid] = bx;
        S.area[tid] = box_area(bx);
        S.score[tid] = P.cand_meta[cs].score;
    }
    __syncthreads();

    // ---- suppression bitmask, two 64-box tiles (4 x 32 bits) per box: bit j <=> IoU(box i, box j) > thr, j >= i.
    // Warp-cooperative: every lane keeps boxes lane, lane+32, lane+64, lane+96 in registers, a warp takes every
    // 4th row i, broadcasts box i from shared memory and turns 32 comparisons into one ballot word.
    {
        const float thr = P.nms_thres;
        const int warp = tid >> 5, lane = tid & 31;
        float4 bj[4];
        float aj[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = (r << 5) + lane;
            bj[r] = j < m ? S.box[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            aj[r] = j < m ? S.area[j] : 0.f;
        }
        for (int i = warp; i < m; i += kSegWarps) {
            const float4 bi = S.box[i];
            const float ai = __fadd_rn(S.area[i], 1e-16f);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                unsigned word = 0;
                if ((r << 5) + 31 >= i && (r << 5) < m) {               // warp-uniform: tile intersects the triangle
                    const int j = (r << 5) + lane;
                    const bool hit = j >= i && j < m && iou_gt(bi, ai, bj[r], aj[r], thr);   // utils.py:271 strict >
                    word = __ballot_sync(kFull, hit);
                }
                if (lane == 0) S.mask32[i][r] = word;
            }
        }
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
                if (tid == 0) { S.kept[nk] = i; S.clu[nk][0] = 0; S.clu[nk][1] = 0; }
                ++nk;
                break;
            }
            const unsigned long long* mrow = reinterpret_cast<const unsigned long long*>(S.mask32[i]);
            const unsigned long long c0 = mrow[0] & a0, c1 = mrow[1] & a1;
            if (tid == 0) { S.kept[nk] = i; S.clu[nk][0] = c0; S.clu[nk][1] = c1; }
            ++nk;
            a0 &= ~c0; a1 &= ~c1;
            if (i < 64) a0 &= ~(1ull << i); else a1 &= ~(1ull << (i - 64));
        }
        if (tid == 0) S.nkept = nk;
    }
    __syncthreads();

    // ---- MERGE box of every kept detection: sum_j s_j*box_j / sum_j s_j over its cluster, in order
    const int nk = S.nkept;
    if (tid < nk) {
        const int i = S.kept[tid];
        const unsigned long long c0 = S.clu[tid][0], c1 = S.clu[tid][1];
        float4 o = S.box[i];
        if (c0 | c1) {
            float sw = 0.f, sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f;
            for (int half = 0; half < 2; ++half) {
                unsigned long long bits = half ? c1 : c0;
                while (bits) {
                    const int j = (half << 6) + __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    const float s = S.score[j];
                    const float4 bj = S.box[j];
                    sw = __fadd_rn(sw, s);
                    sx1 = __fadd_rn(sx1, __fmul_rn(s, bj.x));
                    sy1 = __fadd_rn(sy1, __fmul_rn(s, bj.y));
                    sx2 = __fadd_rn(sx2, __fmul_rn(s, bj.z));
                    sy2 = __fadd_rn(sy2, __fmul_rn(s, bj.w));
                }
            }
            o = make_float4(__fdiv_rn(sx1, sw), __fdiv_rn(sy1, sw), __fdiv_rn(sx2, sw), __fdiv_rn(sy2, sw));
        }
        const size_t cslot = (size_t)b * P.cap + S.slot[i];
        const float cls_conf = P.cand_meta[cslot].cls_conf;
        const int row = (int)(uint32_t)S.key[i];
        float4* st = P.stage + ((size_t)b * P.stage_cap + P.stage_off[(size_t)b * (P.nc + 1) + c] + tid) * 2;
        st[0] = o;
        st[1] = make_float4(S.score[i], cls_conf, __int_as_float(row), (float)c);
    } else if (tid < m) {                            // staged slots beyond the kept ones are marked with a NaN score
        float4* st = P.stage + ((size_t)b * P.stage_cap + P.stage_off[(size_t)b * (P.nc + 1) + c] + tid) * 2;
        st[1] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, (float)c);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSegThreads)
nms_segment_kernel(const __grid_constant__ NmsParams P) {
    __shared__ BigSegSmem S;
    const int small_ctas = (int)gridDim.x - P.big_ctas;
    if ((int)blockIdx.x >= small_ctas) {
        // big-segment role (last in launch order: with no big segment these CTAs leave at once and must not delay
        // the warp-per-segment CTAs)
        const int n_big = P.work_count[0];
        for (int wi = (int)blockIdx.x - small_ctas; wi < n_big; wi += P.big_ctas) nms_big_segment(P, P.work_big[wi], S);
    } else {
        // one warp per (image, class) pair; pairs that do not hold 2..32 boxes cost one offset read
        const int lane = threadIdx.x & 31;
        const int w0 = (int)blockIdx.x * kSegWarps + ((int)threadIdx.x >> 5);
        const int stride = small_ctas * kSegWarps;
        const int n_items = P.batch * P.nc;
        for (int item = w0; item < n_items; item += stride) {
            const int b = item / P.nc, c = item - b * P.nc;
            const size_t o = (size_t)b * (P.nc + 1) + c;
            const int s0 = P.seg_off[o];
            const int n = P.seg_off[o + 1] - s0;
            if (n >= 2 && n <= kSmallSeg) nms_small_segment(P, b, c, s0, n, P.stage_off[o], lane);
            else if (n > kSmallSeg && n <= kPairSeg) nms_pair_segment(P, b, c, s0, n, P.stage_off[o], lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One CTA per image: pull every staged row of the image into shared memory in one coalesced pass, sort the
// kept ones by (score desc, class asc, in-class order) -- a staged row's position already encodes (class, order)
// -- and write the (n, 7) result.  Up to 1024 staged rows: one key per thread, bitonic network in registers
// (shuffles inside a warp, shared memory across warps).  More: the generic shared/global-memory network.

__device__ __forceinline__ unsigned long long final_key(const float4& r1, int q) {
    const float score = r1.x;
    // NaN score = slot not kept.  score desc, then staging position = (class asc, in-class order): utils.py:291
    return (score != score) ? ~0ull : (((unsigned long long)score_key_desc(score) << 32) | (unsigned)q);
}

// Contiguous copy of n floats from shared to global memory with 16-byte vector stores; src and dst must have the
// same address phase modulo 16 bytes (the caller shifts the shared-memory block accordingly).
template <int THREADS>
__device__ __forceinline__ void flat_store(const float* src, float* dst, int n) {
    const int head = min(n, (int)((4 - ((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) & 3));
    if ((int)threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const int n4 = (n - head) >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src + head);
    float4* d4 = reinterpret_cast<float4*>(dst + head);
    for (int i = threadIdx.x; i < n4; i += THREADS) d4[i] = s4[i];
    const int done = head + (n4 << 2);
    if ((int)threadIdx.x < n - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

// Completion stamp of one yolo_b200_nms call: every thread fences its (possibly peer) result stores at system scope,
// the CTA's leader counts itself done, and the last CTA of the grid publishes the lane's next sequence number.
// fence -> device-scope atomic -> fence -> release store: the stamp's observer sees every row of every CTA.
__device__ __forceinline__ void signal_step(const NmsParams& P) {
    if (!P.step_stamp) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t* done = P.work_count + 1;
        if (atomicAdd(done, 1) == (int)gridDim.x - 1) {
            *done = 0;
            const int v = *P.step_seq + 1;
            *P.step_seq = v;
            __threadfence_system();
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(P.step_stamp), "r"(v) : "memory");
        }
    }
}

template <int kFinalThreads>
__device__ __forceinline__ void nms_finalize_body(const NmsParams& P) {
    constexpr int kFinalRows = kFinalThreads;          // fast path: one staged row per thread
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm_raw);   // [kFinalSmemKeys]
    const int b = blockIdx.x, tid = threadIdx.x, nc = P.nc;
    const float4* stage = P.stage + (size_t)b * P.stage_cap * 2;
    const int n_staged = P.stage_off[(size_t)b * (nc + 1) + nc];
    float* out = P.out + (size_t)b * P.out_cap * YOLO_B200_DET_COLS;
    int32_t* out_row = P.out_row + (size_t)b * P.out_cap;

    if (n_staged <= kFinalRows) {
        float4* srow = reinterpret_cast<float4*>(skeys + 2 * kFinalRows);          // [kFinalRows*2] behind two key arrays
        unsigned long long key = ~0ull;
        if (tid < n_staged) {
            const float4 r0 = stage[2 * tid], r1 = stage[2 * tid + 1];
            srow[2 * tid] = r0; srow[2 * tid + 1] = r1;
            key = final_key(r1, tid);
        }
        const int n_out = __syncthreads_count(key != ~0ull);
        if (tid == 0) P.out_count[b] = n_out;
        if (n_out == 0) return;
        int span = 32;
        while (span < n_staged) span <<= 1;                                        // block-uniform power of two
        int flip = 0;
        for (int k = 2; k <= span; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                unsigned long long other;
                if (j >= 32) {                                                     // partner in another warp
                    unsigned long long* sx = skeys + flip * kFinalRows;
                    flip ^= 1;
                    sx[tid] = key;
                    __syncthreads();
                    other = sx[tid ^ j];
                } else {
                    other = __shfl_xor_sync(kFull, key, j);
                }
                const bool take_min = ((tid & k) == 0) == ((tid & j) == 0);
                key = take_min ? (other < key ? other : key) : (other > key ? other : key);
            }
        }
        __syncthreads();
        skeys[tid] = key;
        __syncthreads();
        // assemble the (n_out, 7) block and the row ids in shared memory, then store them as one contiguous run of
        // 16-byte vectors each (the destination may be a peer GPU: wide, fully coalesced stores are what NVLink likes).
        // Both blocks are shifted inside shared memory so that shared and global addresses share the 16-byte phase.
        float* sflat = reinterpret_cast<float*>(srow + 2 * kFinalRows);            // [7*kFinalRows + 4]
        int32_t* sids = reinterpret_cast<int32_t*>(sflat + 7 * kFinalRows + 4);     // [kFinalRows + 4]
        const int mis_o = (int)((reinterpret_cast<uintptr_t>(out) >> 2) & 3), mis_r = (int)((reinterpret_cast<uintptr_t>(out_row) >> 2) & 3);
        for (int e = tid; e < n_out * 8; e += kFinalThreads) {
            const int i = e >> 3, col = e & 7;
            const int q = (int)(uint32_t)skeys[i];
            const float v = reinterpret_cast<const float*>(srow + 2 * q)[col];
            if (col < 6)       sflat[mis_o + i * YOLO_B200_DET_COLS + col] = v;
            else if (col == 7) sflat[mis_o + i * YOLO_B200_DET_COLS + 6] = v;      // class id as float (utils.py:228)
            else               sids[mis_r + i] = __float_as_int(v);
        }
        __syncthreads();
        flat_store<kFinalThreads>(sflat + mis_o, out, n_out * YOLO_B200_DET_COLS);
        flat_store<kFinalThreads>(reinterpret_cast<const float*>(sids + mis_r), reinterpret_cast<float*>(out_row), n_out);
        return;
    }

    // generic path: keys of all staged rows in shared (<= kFinalSmemKeys) or global memory, rows stay in global
    unsigned long long* keys = (n_staged <= P.final_smem_keys) ? skeys : (P.final_keys + (size_t)b * P.stage_cap);
    int mine = 0;
    for (int q = tid; q < n_staged; q += kFinalThreads) {
        const unsigned long long k = final_key(stage[2 * q + 1], q);
        keys[q] = k;
        mine += (k != ~0ull);
    }
    __shared__ int s_count;
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (mine) atomicAdd(&s_count, mine);
    __syncthreads();
    const int n_out = s_count;
    if (tid == 0) P.out_count[b] = n_out;
    if (n_out == 0) return;
    if (n_staged <= P.final_smem_keys) bitonic_sort<false, kFinalThreads>(skeys, nullptr, n_staged);
    else                            bitonic_sort<false, kFinalThreads>(keys, nullptr, n_staged);
    // result rows in chunks of kFinalThreads: gathered into shared memory behind the key array, then stored as
    // contiguous 16-byte vectors (same peer-friendly store pattern as the fast path)
    float* cflat = reinterpret_cast<float*>(skeys + ((P.final_smem_keys + 1) & ~1));   // 16-byte aligned, [7*kFinalThreads + 4]
    int32_t* cids = reinterpret_cast<int32_t*>(cflat + 7 * kFinalThreads + 4);       // [kFinalThreads + 4]
    for (int i0 = 0; i0 < n_out; i0 += kFinalThreads) {
        const int rows = min(kFinalThreads, n_out - i0);
        float* dst = out + (size_t)i0 * YOLO_B200_DET_COLS;
        int32_t* dst_id = out_row + i0;
        const int mis_o = (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3), mis_r = (int)((reinterpret_cast<uintptr_t>(dst_id) >> 2) & 3);
        if (tid < rows) {
            const int q = (int)(uint32_t)keys[i0 + tid];
            const float4 r0 = stage[2 * (size_t)q], r1 = stage[2 * (size_t)q + 1];
            float* o = cflat + mis_o + tid * YOLO_B200_DET_COLS;
            o[0] = r0.x; o[1] = r0.y; o[2] = r0.z; o[3] = r0.w; o[4] = r1.x; o[5] = r1.y; o[6] = r1.w;
            cids[mis_r + tid] = __float_as_int(r1.z);
        }
        __syncthreads();
        flat_store<kFinalThreads>(cflat + mis_o, dst, rows * YOLO_B200_DET_COLS);
        flat_store<kFinalThreads>(reinterpret_cast<const float*>(cids + mis_r), reinterpret_cast<float*>(dst_id), rows);
        __syncthreads();
    }
}

template <int kFinalThreads>
__global__ void __launch_bounds__(kFinalThreads)
nms_finalize_kernel(const __grid_constant__ NmsParams P) {
    nms_finalize_body<kFinalThreads>(P);
    signal_step(P);
}

}  // namespace yb

// ================================================================================================
using namespace yb;

namespace {
struct WsLayout {
    size_t bucket_key, bucket_slot, seg_off, stage_off, work_big, work_count, stage, final_keys, total;
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
    L.work_big = o;    o = align_up(o + (size_t)batch * nc * 4);
    L.work_count = o;  o = align_up(o + 8);
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

extern "C" int yolo_b200_nms_ex(const yolo_b200_box* cand_box, const yolo_b200_meta* cand_meta, const int32_t* count,
                                int batch, int cap_per_img, int nc, float nms_thres, int max_per_class,
                                float* out, int32_t* out_row, int out_cap, int32_t* out_count,
                                void* workspace, size_t workspace_bytes, const yolo_b200_nms_opts* opts,
                                yolo_b200_stream_t stream) {
    if (opts && ((opts->step_seq == nullptr) != (opts->step_stamp == nullptr))) return YOLO_B200_E_NULL;
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
    P.work_big = reinterpret_cast<int32_t*>(ws + L.work_big);
    P.work_count = reinterpret_cast<int32_t*>(ws + L.work_count);
    P.stage = reinterpret_cast<float4*>(ws + L.stage);
    P.final_keys = reinterpret_cast<unsigned long long*>(ws + L.final_keys);
    P.out = out; P.out_row = out_row; P.out_count = out_count;
    P.step_seq = opts ? opts->step_seq : nullptr;
    P.step_stamp = opts ? opts->step_stamp : nullptr;

    cudaError_t e;
    if ((e = cudaMemsetAsync(P.work_count, 0, 2 * sizeof(int32_t), stream)) != cudaSuccess) return (int)e;
    const size_t bucket_smem = (size_t)(4 * nc + 1) * sizeof(int);
    if (bucket_smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(bucket_by_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bucket_smem)) != cudaSuccess)
        return (int)e;
    bucket_by_class_kernel<<<batch, kBucketThreads, bucket_smem, stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long segs = (long long)batch * nc;
    // the LAST `big` CTAs take big segments one at a time, the others run one small segment per warp
    const int big = (int)(segs < (long long)sms * 11 ? segs : (long long)sms * 11);   // 20 KB shared each: 11 per SM
    const long long small_ctas = (segs + kSegWarps - 1) / kSegWarps;
    const int small = (int)(small_ctas < (long long)sms * 16 ? small_ctas : (long long)sms * 16);
    P.big_ctas = big;
    nms_segment_kernel<<<big + small, kSegThreads, 0, stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    // finalize: small CTAs when an image cannot stage many rows (more images resident per SM), big ones otherwise
    const bool small_final = stage_cap <= 2560;
    const int ft = small_final ? kFinalThreadsSmall : kFinalThreadsBig;
    P.final_smem_keys = stage_cap < kFinalSmemKeys ? stage_cap : kFinalSmemKeys;
    const size_t out_block = (size_t)(8 * ft + 8) * sizeof(float);               // (rows x 7) block + row ids, with phase slack
    size_t final_smem = (size_t)48 * ft + out_block;                             // fast path: 2 key arrays, staged rows, output block
    const size_t key_bytes = (size_t)((P.final_smem_keys + 1) & ~1) * 8;
    if (key_bytes + out_block > final_smem) final_smem = key_bytes + out_block;
    if (small_final) {
        if ((e = cudaFuncSetAttribute(nms_finalize_kernel<kFinalThreadsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)final_smem)) != cudaSuccess)
            return (int)e;
        nms_finalize_kernel<kFinalThreadsSmall><<<batch, kFinalThreadsSmall, final_smem, stream>>>(P);
    } else {
        if ((e = cudaFuncSetAttribute(nms_finalize_kernel<kFinalThreadsBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)final_smem)) != cudaSuccess)
            return (int)e;
        nms_finalize_kernel<kFinalThreadsBig><<<batch, kFinalThreadsBig, final_smem, stream>>>(P);
    }
    return (int)cudaGetLastError();
}

extern "C" int yolo_b200_nms(const yolo_b200_box* cand_box, const yolo_b200_meta* cand_meta, const int32_t* count,
                             int batch, int cap_per_img, int nc, float nms_thres, int max_per_class,
                             float* out, int32_t* out_row, int out_cap, int32_t* out_count,
                             void* workspace, size_t workspace_bytes, yolo_b200_stream_t stream) {
    return yolo_b200_nms_ex(cand_box, cand_meta, count, batch, cap_per_img, nc, nms_thres, max_per_class, out, out_row,
                            out_cap, out_count, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int yolo_b200_abi_version(void) { return YOLO_B200_ABI_VERSION; }

extern "C" const char* yolo_b200_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case YOLO_B200_E_NULL: return "null pointer argument";
        case YOLO_B200_E_RANGE: return "argument out of range";
        case YOLO_B200_E_ALIGN: return "pointer not aligned as documented";
        case YOLO_B200_E_WORKSPACE: return "workspace too small";
        case YOLO_B200_E_UNSUPPORTED: return "geometry not covered by the fused head kernel";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}
