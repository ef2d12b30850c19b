// Segmented NMS kernels of the yolo_b200 hot path (sm_100a).
//
// Replaces the per-image / per-class Python loops of the reference's non_max_suppression
// (utils/utils.py:237-291) with three launches that work on all images at once:
//
//   bucket_by_class_kernel   one CTA per image: class histogram -> segment offsets, scatter of
//                            64-bit sort keys into class buckets (utils.py:241-242); single-box
//                            classes are emitted here, untouched (utils.py:244-246); draws up the two
//                            work lists of the segment stage (groups of small segments | bigger segments)
//   nms_segment_kernel       single-warp CTAs over those lists, everything in registers: several small
//                            segments side by side on one warp's lanes, or one bigger segment per warp;
//                            order by (score desc, row asc) keeping the first max_per_class
//                            (utils.py:237, 247-250), IoU suppression bits by ballot / per-lane words,
//                            lane-uniform greedy sweep and the score-weighted MERGE box
//                            (utils.py:266-275 with bbox_iou utils.py:63-96)
//   nms_finalize_kernel      one CTA per image: order kept rows by (score desc, class asc, in-class
//                            order) and write the (n, 7) result (utils.py:289-291)
//
// Tie rule (documented): equal scores keep ascending anchor-row order.  Every key
// carries the anchor row, so keys are unique and the (unstable) bitonic networks used here give a
// deterministic result that does not depend on the order compaction produced.
#include "common.cuh"

namespace yb {

constexpr int kMaxPerClassLimit = 128;  // boxes a segment keeps for the sweep: four per lane
constexpr int kSegThreads = 32;         // CTA size of the segment kernel: ONE warp.  With a single-warp CTA every value derived
                                        // from blockIdx or loaded from a CTA-uniform address is provably warp-uniform, so ptxas
                                        // emits the shuffles and votes bare (no convergence guards: half the code size)
constexpr int kSmallSeg = 32;           // segments up to this size: one box per lane
constexpr int kPairSeg = 64;            // ... up to this size: two boxes per lane
constexpr int kQuadSeg = 128;           // ... any bigger: four boxes per lane -- the first kQuadSeg keys are sorted in registers,
                                        //     the rest is streamed through in groups of 32 keys
constexpr int kBucketThreadsBig = 1024;  // bucket CTA size when an image can hold many candidates (the kernel is latency-bound:
constexpr int kBucketThreadsSmall = 256; // threads = loads in flight), and when it cannot (many images, few candidates each)
constexpr int kPackSlotBits = 27;       // packed groups: candidate slot bits of the payload word (the host checks cap_per_img)
constexpr int kGroupRun = 8;            // a group of small segments is a run of classes inside one block of this many classes
constexpr int kBucketRegs = 1;          // candidate records a bucket thread keeps in registers between its two passes
constexpr int kFinalThreadsBig = 512;    // finalize CTA size when an image can stage many rows
constexpr int kFinalThreadsSmall = 128;  // ... and when it cannot (small per-image capacity, usually large batches: the CTA's registers x time
                                         //     is what the next batch's decode kernel loses -- 256 threads: cfg 3 step 181 us, 128: 177 us)
constexpr int kFinalSmallCap = kSmallImageRows;     // staging capacity per image up to which the small finalize CTA is used
constexpr int kFinalKpt = 16;            // keys a finalize thread sorts in registers at most

struct NmsParams {
    const yolo_b200_box* cand_box;
    const yolo_b200_meta* cand_meta;
    const int32_t* count;
    int batch, cap, nc, mpc, stage_cap, out_cap;
    int pack_ok;                      // segments of 2..32 boxes are packed several to a warp (max_per_class >= 32, slots fit kPackSlotBits)
    int packed_ctas;                  // CTAs of the segment kernel that work on the group list
    int final_smem_keys;              // staged rows per image up to which the finalize kernel sorts in registers / shared memory
    int final_key_slots;              // 64-bit slots of its shared-memory key area (>= threads x keys per thread of that sort)
    float nms_thres;
    // workspace
    unsigned long long* bucket_key;   // [batch*cap]  (score-descending key << 32) | row
    uint32_t* bucket_slot;            // [batch*cap]  candidate slot of the key
    int32_t* seg_off;                 // [batch*(nc+1)] start of every class bucket
    int32_t* stage_off;               // [batch*(nc+1)] start of every class in the staging rows (lengths capped at mpc)
    int32_t* work_count;              // [4] segment tickets handed out | finalize CTAs that have finished (self-resetting) |
                                      //     entries of group_list | entries of seg_list (both counted by the bucket kernel)
    int2* group_list;                 // [batch*nc] work of nms_packed_groups: (image, first class | classes << 16)
    int4* seg_list;                   // [batch*nc] one-segment work of nms_segment_kernel: (image, class, bucket offset, boxes)
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
// The band edges are hoisted out of the pair loop: hi = thr * (1 + 1e-6), lo = thr * (1 - 1e-6) (+inf / -inf when thr is
// too small for the argument above, which sends every pair to the division); the range check on the union is one
// unsigned comparison (normal, positive, below 2^127).
struct IouThr { float thr, hi, lo; };
__device__ __forceinline__ IouThr make_iou_thr(float thr) {
    IouThr t;
    t.thr = thr;
    const bool ok = thr > 1e-30f;
    t.hi = ok ? __fmul_rn(thr, 1.000001f) : __int_as_float(0x7f800000);
    t.lo = ok ? __fmul_rn(thr, 0.999999f) : __int_as_float(0xff800000);
    return t;
}
__device__ __forceinline__ bool iou_gt(const float4& a, float area_a_eps, const float4& b, float area_b, const IouThr& t) {
    const float ix = __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float iy = __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float inter = __fmul_rn(fmaxf(ix, 0.0f), fmaxf(iy, 0.0f));
    const float uni = __fsub_rn(__fadd_rn(area_a_eps, area_b), inter);
    const bool yes = inter > __fmul_rn(t.hi, uni);
    const bool no = inter < __fmul_rn(t.lo, uni);
    const bool in_range = (__float_as_uint(uni) - 0x00800000u) < 0x7e800000u;
    if ((yes || no) && in_range) return yes;
    return __fdiv_rn(inter, uni) > t.thr;
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

__device__ __forceinline__ float4 shfl4(const float4& v, int src) {
    return make_float4(__shfl_sync(kFull, v.x, src), __shfl_sync(kFull, v.y, src), __shfl_sync(kFull, v.z, src),
                       __shfl_sync(kFull, v.w, src));
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
// bucket / staging offsets, pass 2 scatters 64-bit keys into the class buckets.  Between the passes the two work lists of the
// segment kernel are drawn up: classes of 2..32 boxes are laid out greedily, in class order, into GROUPS of at most 32
// boxes (one warp of nms_packed_groups each; a group is a run of consecutive classes inside one block of kGroupRun classes:
// the blocks are laid out by different threads, and the warp fetches the run's offsets with one load per lane), every bigger class
// is one entry of the segment list.
template <int kBucketThreads>
__global__ void __launch_bounds__(kBucketThreads, 2048 / kBucketThreads)
bucket_by_class_kernel(const __grid_constant__ NmsParams P) {
    extern __shared__ int sm_i[];
    const int nc = P.nc;
    int* hist = sm_i;                 // [nc]
    int* cur = hist + nc;             // [nc]   scatter cursors (start at the bucket offset)
    int* soff = cur + nc;             // [nc+1] staging offsets
    int* item_g = soff + nc + 1;      // [nc]   groups of this image: first class | classes << 16
    int* item_s = item_g + nc;        // [nc]   classes of this image that get a warp of their own
    __shared__ int s_total, s_ng, s_ns, s_base_g, s_base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = min(P.count[b], P.cap);
    const int4* meta4 = reinterpret_cast<const int4*>(P.cand_meta) + (size_t)b * P.cap;

    for (int c = tid; c < nc; c += kBucketThreads) hist[c] = 0;
    if (tid == 0) { s_ng = 0; s_ns = 0; }
    __syncthreads();
    int4 mt[kBucketRegs];
#pragma unroll
    for (int k = 0; k < kBucketRegs; ++k) {
        const int i = tid + k * kBucketThreads;
        if (i < n) { mt[k] = meta4[i]; if ((unsigned)mt[k].z < (unsigned)nc) atomicAdd(&hist[mt[k].z], 1); }
    }
#pragma unroll 4
    for (int i = tid + kBucketRegs * kBucketThreads; i < n; i += kBucketThreads) {      // (several loads in flight per thread)
        const int c = reinterpret_cast<const int*>(meta4)[4 * (size_t)i + 2];
        if ((unsigned)c < (unsigned)nc) atomicAdd(&hist[c], 1);       // records with a class id outside [0, nc) are ignored
    }
    __syncthreads();

    if (tid < 32) {
        int run_off = 0, run_soff = 0;
        for (int c0 = 0; c0 < nc; c0 += 32) {
            const int c = c0 + lane;
            const int len = c < nc ? hist[c] : 0;
            const int cp = min(len, P.mpc);
            const int il = warp_incl_scan(len, lane), ic = warp_incl_scan(cp, lane);
            if (c < nc) {
                cur[c] = run_off + il - len;
                soff[c] = run_soff + ic - cp;
            }
            run_off += __shfl_sync(kFull, il, 31);
            run_soff += __shfl_sync(kFull, ic, 31);
        }
        if (lane == 0) { soff[nc] = run_soff; s_total = run_off; }
    }
    __syncthreads();

    int32_t* g_seg = P.seg_off + (size_t)b * (nc + 1);
    int32_t* g_stage = P.stage_off + (size_t)b * (nc + 1);
    for (int c = tid; c < nc; c += kBucketThreads) {
        g_seg[c] = cur[c];
        g_stage[c] = soff[c];
    }
    if (tid == 0) { g_seg[nc] = s_total; g_stage[nc] = soff[nc]; }
    for (int blk = tid; blk * kGroupRun < nc; blk += kBucketThreads) {       // classes [blk kGroupRun, (blk + 1) kGroupRun)
        const int c_end = min(nc, (blk + 1) * kGroupRun);
        int fill = 0, g0 = 0;
        for (int c = blk * kGroupRun; c < c_end; ++c) {
            const int len = hist[c];
            if (len < 2) continue;                                // nothing to suppress (utils.py:244-246)
            if (!P.pack_ok || len > kSmallSeg) { item_s[atomicAdd(&s_ns, 1)] = c; continue; }
            if (fill + len > 32) { item_g[atomicAdd(&s_ng, 1)] = g0 | ((c - g0) << 16); fill = 0; }
            if (fill == 0) g0 = c;
            fill += len;
        }
        if (fill) item_g[atomicAdd(&s_ng, 1)] = g0 | ((c_end - g0) << 16);
    }
    __syncthreads();          // cur[] is read above and bumped below
    if (tid == 0) {           // the image's slice of the global lists (the round trip hides behind the scatter)
        s_base_g = s_ng ? atomicAdd(P.work_count + 2, s_ng) : 0;
        s_base_s = s_ns ? atomicAdd(P.work_count + 3, s_ns) : 0;
    }

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
#pragma unroll 4
    for (int i = tid + kBucketRegs * kBucketThreads; i < n; i += kBucketThreads) place(meta4[i], i);
    __syncthreads();          // s_base_*; cur[c] is now the END of bucket c
    for (int i = tid; i < s_ng; i += kBucketThreads) P.group_list[s_base_g + i] = make_int2(b, item_g[i]);
    for (int i = tid; i < s_ns; i += kBucketThreads) {
        const int c = item_s[i], len = hist[c];
        P.seg_list[s_base_s + i] = make_int4(b, c, cur[c] - len, len);
    }
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
            if ((ok < key) == take_min) { key = ok; slot = os; }     // keys are unique (equal pads may swap freely)
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
    const IouThr thr = make_iou_thr(P.nms_thres);

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

// Several small segments in ONE warp (the single-warp CTAs [0, packed_ctas) of nms_segment_kernel stride over the bucket
// kernel's group list).  Most (image,
// class) pairs of a detection workload hold a handful of boxes (spp-608 at conf 0.3: ~10, tiny-416: ~7), so a warp per
// segment leaves most lanes idle and -- the work being a latency chain offsets -> keys -> boxes -> sweep -- holds its
// registers for a whole chain per segment, registers the next batch's decode CTAs are waiting for.  A group is a run of
// consecutive classes of one image whose segments of 2..32 boxes total at most 32: segment g lies on lanes
// [start_g, start_g + n_g).
//   * order: one bitonic network over composite keys (segment, score desc, row asc) -- the segment id (= its first lane, 5
//     bits) rides in the top bits of the payload word next to the candidate slot (< 2^27, checked by the host), so every
//     segment ends up sorted in place (utils.py:237)
//   * sweep: all segments advance in lockstep -- every lane reads ITS segment's pivot with an indexed shuffle, one ballot
//     collects the clusters of all segments, the MERGE sums run member by member in segment order with the lanes of a segment
//     computing the same sums (utils.py:266-275); the number of rounds is that of the segment keeping the most boxes
// Arithmetic and order of operations are those of nms_small_segment: results are bit-identical.
__device__ __forceinline__ void nms_packed_groups(const NmsParams& P, int* s_cls, int first, int stride, int lane) {
    const int total = P.work_count[2];
    const IouThr thr = make_iou_thr(P.nms_thres);
    for (int it = first; it < total; it += stride) {
        const int2 d = P.group_list[it];
        const int b = d.x, c0 = d.y & 0xffff, ncls = d.y >> 16;
        // lane j < ncls looks at class c0 + j; classes of the run that are empty, single or big take no lanes
        int len = 0, s0 = 0, sto = 0;
        if (lane < ncls) {
            const size_t o = (size_t)b * (P.nc + 1) + c0 + lane;
            s0 = P.seg_off[o];
            len = P.seg_off[o + 1] - s0;
            sto = P.stage_off[o];
            if (len < 2 || len > kSmallSeg) len = 0;
        }
        const int incl = warp_incl_scan(len, lane);
        const int fill = __shfl_sync(kFull, incl, 31);            // lanes in use
        const unsigned heads = __reduce_or_sync(kFull, len ? (1u << (incl - len)) : 0u);   // first lane of every segment
        if (len) s_cls[incl - len] = lane;
        __syncwarp();
        const bool act = lane < fill;
        const int my_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)) | 1u);
        const int my_j = s_cls[my_start] & 31;
        __syncwarp();                                             // s_cls is rewritten by the next group
        const int my_n = __shfl_sync(kFull, len, my_j);
        const int my_s0 = __shfl_sync(kFull, s0, my_j);
        const int st_off = __shfl_sync(kFull, sto, my_j);
        const int c = c0 + my_j;

        unsigned long long key = ~0ull;
        uint32_t pay = 0xffffffffu;                   // unused lanes: segment 31 (no segment starts there), key +inf
        if (act) {
            const size_t g = (size_t)b * P.cap + my_s0 + (lane - my_start);
            key = P.bucket_key[g];
            pay = ((uint32_t)my_start << kPackSlotBits) | P.bucket_slot[g];
        }
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
            if ((k >> 1) >= fill) break;              // the lanes in use already fit in one sorted block
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(kFull, key, j);
                const uint32_t op = __shfl_xor_sync(kFull, pay, j);
                const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
                const uint32_t so = op >> kPackSlotBits, sm = pay >> kPackSlotBits;
                const bool less = so < sm || (so == sm && ok < key);
                if (less == take_min) { key = ok; pay = op; }        // (segment, key) pairs are unique; equal pads may swap freely
            }
        }
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f, cls_conf = 0.f;
        if (act) {
            const size_t cs = (size_t)b * P.cap + (pay & ((1u << kPackSlotBits) - 1u));
            box = reinterpret_cast<const float4*>(P.cand_box)[cs];
            const int4 m = reinterpret_cast<const int4*>(P.cand_meta)[cs];
            score = __int_as_float(m.x);
            cls_conf = __int_as_float(m.y);
        }
        const float area = box_area(box);
        const unsigned segmask = act ? ((my_n >= 32 ? ~0u : ((1u << my_n) - 1u)) << my_start) : 0u;

        unsigned alive = __ballot_sync(kFull, act);
        int nk = 0, kept_i = lane;
        float4 obox = box;
        while (alive) {                                              // greedy sweep of every segment at once (utils.py:266-275)
            const unsigned mine = alive & segmask;
            const bool has = mine != 0u;
            const int piv = has ? __ffs(mine) - 1 : lane;
            const bool single = (mine & (mine - 1u)) == 0u;          // last survivor of its segment: emitted unmerged (utils.py:268-270)
            const float4 bi = shfl4(box, piv);
            const float ai = __fadd_rn(__shfl_sync(kFull, area, piv), 1e-16f);
            const bool hit = has && !single && lane >= piv && iou_gt(bi, ai, box, area, thr);     // utils.py:271
            const unsigned cl_all = __ballot_sync(kFull, hit) & alive;
            const unsigned pivots = __ballot_sync(kFull, has && lane == piv);
            alive &= ~(cl_all | pivots);
            const unsigned my_cl = cl_all & segmask;
            float4 m4 = bi;
            if (__any_sync(kFull, my_cl != 0u)) {                    // score-weighted mean of each cluster, in segment order (utils.py:272-273)
                float sw = 0.f, sx1 = 0.f, sy1 = 0.f, sx2 = 0.f, sy2 = 0.f;
                unsigned rem = my_cl;
                while (__any_sync(kFull, rem != 0u)) {
                    const int j = rem ? __ffs(rem) - 1 : lane;
                    const float sj = __shfl_sync(kFull, score, j);
                    const float4 bj = shfl4(box, j);
                    if (rem) {
                        sw = __fadd_rn(sw, sj);
                        sx1 = __fadd_rn(sx1, __fmul_rn(sj, bj.x));
                        sy1 = __fadd_rn(sy1, __fmul_rn(sj, bj.y));
                        sx2 = __fadd_rn(sx2, __fmul_rn(sj, bj.z));
                        sy2 = __fadd_rn(sy2, __fmul_rn(sj, bj.w));
                        rem &= rem - 1u;
                    }
                }
                if (my_cl) m4 = make_float4(__fdiv_rn(sx1, sw), __fdiv_rn(sy1, sw), __fdiv_rn(sx2, sw), __fdiv_rn(sy2, sw));
            }
            if (has && lane == my_start + nk) { obox = m4; kept_i = piv; }
            nk += has;
        }
        // lane start + k holds the k-th kept detection of its segment: fetch its score / cls_conf / row from the lane of box kept_i
        const float o_score = __shfl_sync(kFull, score, kept_i);
        const float o_conf = __shfl_sync(kFull, cls_conf, kept_i);
        const int o_row = (int)__shfl_sync(kFull, (uint32_t)key, kept_i);
        if (act) {                                       // staged slots beyond the kept ones are marked with a NaN score
            const int pos = lane - my_start;
            float4* st = P.stage + ((size_t)b * P.stage_cap + st_off + pos) * 2;
            if (pos < nk) st[0] = obox;
            st[1] = make_float4(pos < nk ? o_score : __int_as_float(0x7fc00000), o_conf, __int_as_float(o_row), (float)c);
        }
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
        if ((ok < key) == take_min) { key = ok; slot = os; }
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
    const IouThr thr = make_iou_thr(P.nms_thres);
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

// Segments of more than 64 boxes: still one warp, four boxes per lane (element e = 32 r + lane lives in register r of
// `lane`), no shared memory and no barriers.
//   1. the first 128 keys go through a register bitonic network (strides < 32: shuffles, strides 32 / 64: inside the lane);
//      every further group of 32 keys is sorted across the lanes in descending order, min-ed against the top quarter of
//      the kept 128 (ascending against descending: the 128 smallest of the union, as a bitonic sequence) and re-merged;
//      groups that hold nothing below the current max_per_class-th key are skipped after one vote (utils.py:237, 247-250)
//   2. suppression bits, dense: lane l owns the rows of its four boxes, the column boxes are broadcast one at a time;
//      bit j of mk[r][q] <=> IoU(box 32r+lane, box 32q+j) > thr.  Only q >= r is ever looked at (the sweep only asks about
//      boxes behind the kept one), and a column none of whose 32 x (q+1) pairs overlaps at all costs one vote
//   3. greedy sweep over the alive masks, lane-uniform, with the MERGE sums taken in segment order (utils.py:266-275)
template <int R>
__device__ __forceinline__ void warp_bitonic(unsigned long long (&key)[R], uint32_t (&slot)[R], int lane, int k_first, bool desc) {
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
        if (k < k_first) continue;                                // compile-time after unrolling: merge-only callers
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r & jr) continue;
                    const bool asc = (((r << 5) & k) == 0) != desc;
                    if ((key[r] > key[r | jr]) == asc) {
                        const unsigned long long t = key[r]; key[r] = key[r | jr]; key[r | jr] = t;
                        const uint32_t u = slot[r]; slot[r] = slot[r | jr]; slot[r | jr] = u;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int e = (r << 5) | lane;
                    const bool take_min = (((e & k) == 0) == ((lane & j) == 0)) != desc;
                    const unsigned long long ok = __shfl_xor_sync(kFull, key[r], j);
                    const uint32_t os = __shfl_xor_sync(kFull, slot[r], j);
                    if ((ok < key[r]) == take_min) { key[r] = ok; slot[r] = os; }
                }
            }
        }
    }
}

struct QuadSmem {
    float4 box[kMaxPerClassLimit];     // x1 y1 x2 y2
    float4 meta[kMaxPerClassLimit];    // score, cls_conf, row (bits), area
};

__device__ __forceinline__ void nms_quad_segment(const NmsParams& P, QuadSmem& S, int b, int c, int s0, int n, int st_off, int lane) {
    const IouThr thr = make_iou_thr(P.nms_thres);
    const unsigned long long* gkey = P.bucket_key + (size_t)b * P.cap + s0;
    const uint32_t* gslot = P.bucket_slot + (size_t)b * P.cap + s0;

    // ---- 1. order by (score desc, row asc), keep the first max_per_class
    unsigned long long key[4];
    uint32_t slot[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = (r << 5) + lane;
        key[r] = ~0ull; slot[r] = 0;
        if (e < n) { key[r] = gkey[e]; slot[r] = gslot[e]; }
    }
    warp_bitonic<4>(key, slot, lane, 2, false);
    const int m = min(n, P.mpc);                                  // utils.py:247-250
    const int ql = (m - 1) & 31, qr = (m - 1) >> 5;               // where the m-th key lives
    for (int pos = kQuadSeg; pos < n; pos += 128) {               // four groups of 32 keys are fetched at a time
        unsigned long long nk4[4];
        uint32_t ns4[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int e = pos + (g << 5) + lane;
            nk4[g] = ~0ull; ns4[g] = 0;
            if (e < n) { nk4[g] = gkey[e]; ns4[g] = gslot[e]; }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            // a key that is not below the current m-th key cannot be among the first max_per_class
            const unsigned long long kth = __shfl_sync(kFull, qr == 0 ? key[0] : qr == 1 ? key[1] : qr == 2 ? key[2] : key[3], ql);
            if (!__any_sync(kFull, nk4[g] < kth)) continue;
            unsigned long long gk[1] = {nk4[g]};
            uint32_t gs[1] = {ns4[g]};
            warp_bitonic<1>(gk, gs, lane, 2, true);
            if (gk[0] < key[3]) { key[3] = gk[0]; slot[3] = gs[0]; }
            warp_bitonic<4>(key, slot, lane, 128, false);
        }
    }

    // the kept boxes: four per lane in registers (the row side of the IoU tests) and all of them in shared memory, from
    // where the column side, the sweep and the MERGE sums read them with one broadcast load instead of five shuffles
    float4 box[4];
    float area[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int e = (r << 5) + lane;
        box[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 mt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < m) {
            const size_t cs = (size_t)b * P.cap + slot[r];
            box[r] = reinterpret_cast<const float4*>(P.cand_box)[cs];
            mt = reinterpret_cast<const float4*>(P.cand_meta)[cs];          // score, cls_conf, cls, row (bit patterns)
        }
        area[r] = box_area(box[r]);
        S.box[e] = box[r];
        S.meta[e] = make_float4(mt.x, mt.y, mt.w, area[r]);                 // score, cls_conf, row, area
    }
    __syncwarp();

    // ---- 2. suppression bits (utils.py:271: strict >, box1 = the kept box = the row)
    const bool may_skip = thr.thr >= 0.0f;                        // no overlap => IoU 0 => not above a non-negative threshold
    unsigned mk[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) mk[r][q] = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int nj = min(32, m - (q << 5));
        for (int l = 0; l < nj; ++l) {
            const float4 bj = S.box[(q << 5) + l];
            const float aj = S.meta[(q << 5) + l].w;
            bool ov = false;
#pragma unroll
            for (int r = 0; r <= q; ++r) {
                const float ix = __fsub_rn(fminf(box[r].z, bj.z), fmaxf(box[r].x, bj.x));
                const float iy = __fsub_rn(fminf(box[r].w, bj.w), fmaxf(box[r].y, bj.y));
                ov |= ix > 0.0f && iy > 0.0f;
            }
            if (may_skip && !__any_sync(kFull, ov)) continue;
#pragma unroll
            for (int r = 0; r <= q; ++r)
                mk[r][q] |= (unsigned)iou_gt(box[r], __fadd_rn(area[r], 1e-16f), bj, aj, thr) << l;
        }
    }

    // ---- 3. greedy sweep (utils.py:266-275); kept rows are written as they are found: lane 0 the box, lane 1 the rest
    unsigned alive[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int left = m - (r << 5);
        alive[r] = left >= 32 ? ~0u : (left > 0 ? ((1u << left) - 1u) : 0u);
    }
    float4* stp = P.stage + ((size_t)b * P.stage_cap + st_off) * 2 + (lane & 1);
    const float fc = (float)c;
    int nk = 0;
#pragma unroll
    for (int rs = 0; rs < 4; ++rs) {
        while (alive[rs]) {
            const int l = __ffs(alive[rs]) - 1;
            const int ei = (rs << 5) + l;
            bool last = (alive[rs] & (alive[rs] - 1u)) == 0u;
#pragma unroll
            for (int q = rs + 1; q < 4; ++q) last = last && alive[q] == 0u;
            float4 m4 = S.box[ei];
            const float4 mi = S.meta[ei];
            if (last) {                                          // last survivor: emitted unmerged (utils.py:268-270)
                alive[rs] = 0u;
            } else {
                unsigned cl[4] = {0u, 0u, 0u, 0u};
                unsigned others = 0u;
#pragma unroll
                for (int q = rs; q < 4; ++q) {
                    const unsigned rowbits = __shfl_sync(kFull, mk[rs][q], l);
                    cl[q] = rowbits & alive[q];
                    alive[q] &= ~rowbits;
                    if (q > rs) others |= cl[q];
                }
                alive[rs] &= ~(1u << l);
                if (cl[rs] | others) {                           // score-weighted mean of the cluster, in order (utils.py:272-273)
                    float sw, sx1, sy1, sx2, sy2;
                    if (others == 0u && cl[rs] == (1u << l)) {   // the usual case: the box is alone in its cluster
                        sw = mi.x;
                        sx1 = __fmul_rn(sw, m4.x); sy1 = __fmul_rn(sw, m4.y); sx2 = __fmul_rn(sw, m4.z); sy2 = __fmul_rn(sw, m4.w);
                    } else {
                        sw = 0.f; sx1 = 0.f; sy1 = 0.f; sx2 = 0.f; sy2 = 0.f;
#pragma unroll
                        for (int q = rs; q < 4; ++q) {
                            for (unsigned bits = cl[q]; bits; bits &= bits - 1u) {
                                const int j = (q << 5) + __ffs(bits) - 1;
                                const float sj = S.meta[j].x;
                                const float4 bj = S.box[j];
                                sw = __fadd_rn(sw, sj);
                                sx1 = __fadd_rn(sx1, __fmul_rn(sj, bj.x));
                                sy1 = __fadd_rn(sy1, __fmul_rn(sj, bj.y));
                                sx2 = __fadd_rn(sx2, __fmul_rn(sj, bj.z));
                                sy2 = __fadd_rn(sy2, __fmul_rn(sj, bj.w));
                            }
                        }
                    }
                    // the four IEEE divisions run as ONE warp instruction: lane c divides coordinate c
                    const float num = (lane & 3) == 0 ? sx1 : (lane & 3) == 1 ? sy1 : (lane & 3) == 2 ? sx2 : sy2;
                    const float qd = __fdiv_rn(num, sw);
                    m4 = make_float4(__shfl_sync(kFull, qd, 0), __shfl_sync(kFull, qd, 1), __shfl_sync(kFull, qd, 2),
                                     __shfl_sync(kFull, qd, 3));
                }
            }
            if (lane < 2) *stp = lane == 0 ? m4 : make_float4(mi.x, mi.y, mi.z, fc);
            stp += 2;
            ++nk;
        }
    }
    float4* st = P.stage + ((size_t)b * P.stage_cap + st_off) * 2;
    for (int e = nk + lane; e < m; e += 32)          // staged slots beyond the kept ones are marked with a NaN score
        st[2 * e + 1] = make_float4(__int_as_float(0x7fc00000), 0.f, 0.f, fc);
    __syncwarp();                                     // the next segment of this warp reuses the shared-memory boxes
}

// The segment stage, single-warp CTAs over the bucket kernel's two work lists (their lengths are only known on the device).
// CTAs [0, packed_ctas): groups of small segments (nms_packed_groups), strided.  The rest: one entry of the segment list
// each -- the (image, class) pairs with more than 32 boxes (every pair with at least two when packing is off) -- the first
// entry is the CTA's own index, further ones come by ticket (work_count[0], zeroed by the call): those segments differ in
// cost by an order of magnitude.
__global__ void __launch_bounds__(kSegThreads, 32)
nms_segment_kernel(const __grid_constant__ NmsParams P) {
    __shared__ QuadSmem S;
    __shared__ int s_item;
    __shared__ int s_cls[32];
    const int lane = threadIdx.x;
    if ((int)blockIdx.x < P.packed_ctas) {
        nms_packed_groups(P, s_cls, (int)blockIdx.x, P.packed_ctas, lane);
        return;
    }
    const int total = P.work_count[3];
    const int n_ctas = (int)gridDim.x - P.packed_ctas;
    for (int it = (int)blockIdx.x - P.packed_ctas; it < total;) {
        const int4 d = P.seg_list[it];                            // CTA-uniform address: provably warp-uniform
        const int b = d.x, c = d.y, s0 = d.z, n = d.w;
        const int st = P.stage_off[(size_t)b * (P.nc + 1) + c];
        if (n <= kSmallSeg)     nms_small_segment(P, b, c, s0, n, st, lane);
        else if (n <= kPairSeg) nms_pair_segment(P, b, c, s0, n, st, lane);
        else                    nms_quad_segment(P, S, b, c, s0, n, st, lane);
        if (n_ctas >= total) break;                               // every entry has its own CTA: no ticket needed
        if (lane == 0) s_item = n_ctas + atomicAdd(P.work_count, 1);
        __syncwarp();
        it = s_item;
        __syncwarp();
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

// Ascending sort of THREADS x KPT keys held KPT per thread (element e = tid * KPT + r), positions without a key carrying
// ~0.  Bitonic network: the KPT lowest strides are register-against-register, the next five go through shuffles and only
// the strides that cross warps through shared memory (`sx`, THREADS x KPT keys, register-major so that the exchange is
// bank-conflict free).  `span`: block-uniform power of two >= the number of real keys.
template <int KPT, int THREADS>
__device__ __forceinline__ void block_sort_keys(unsigned long long (&key)[KPT], unsigned long long* sx, int span) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= span; k <<= 1) {
        const bool up = ((tid * KPT) & k) == 0;                   // direction of this thread's keys once k >= KPT
        int j = k >> 1;
        for (; j >= 32 * KPT; j >>= 1) {
            const int tj = j / KPT;
            __syncthreads();
#pragma unroll
            for (int r = 0; r < KPT; ++r) sx[r * THREADS + tid] = key[r];
            __syncthreads();
            const bool take_min = up == ((tid & tj) == 0);
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const unsigned long long o = sx[r * THREADS + (tid ^ tj)];
                if ((o < key[r]) == take_min) key[r] = o;         // equal keys (pads) may go either way
            }
        }
        for (; j >= KPT; j >>= 1) {
            const int tj = j / KPT;
            const bool take_min = up == ((tid & tj) == 0);
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const unsigned long long o = __shfl_xor_sync(kFull, key[r], tj);
                if ((o < key[r]) == take_min) key[r] = o;
            }
        }
#pragma unroll
        for (int jj = KPT / 2; jj >= 1; jj >>= 1) {
            if (jj > (k >> 1)) continue;
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                if (r & jj) continue;
                const bool asc = ((tid * KPT + r) & k) == 0;
                const unsigned long long a = key[r], b = key[r | jj];
                if ((a > b) == asc) { key[r] = b; key[r | jj] = a; }
            }
        }
    }
}

// Keys of the staged rows of one image, KPT per thread, sorted; leaves them in skeys[0 .. THREADS*KPT) in final order and
// returns how many rows were kept.
template <int KPT, int THREADS>
__device__ __forceinline__ int sort_staged_keys(const float4* stage, int n_staged, unsigned long long* skeys, int* s_count) {
    const int tid = threadIdx.x;
    unsigned long long key[KPT];
    int mine = 0;
#pragma unroll
    for (int r = 0; r < KPT; ++r) {
        const int q = tid * KPT + r;
        key[r] = ~0ull;
        if (q < n_staged) key[r] = final_key(stage[2 * q + 1], q);
        mine += key[r] != ~0ull;
    }
    if (tid == 0) *s_count = 0;
    __syncthreads();
    if (mine) atomicAdd(s_count, mine);
    int span = 2;
    while (span < n_staged) span <<= 1;
    block_sort_keys<KPT, THREADS>(key, skeys, span);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < KPT; ++r) skeys[tid * KPT + r] = key[r];
    __syncthreads();
    return *s_count;
}

template <int kFinalThreads>
__device__ __forceinline__ void nms_finalize_body(const NmsParams& P) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    unsigned long long* skeys = reinterpret_cast<unsigned long long*>(sm_raw);   // [final_key_slots]
    const int b = blockIdx.x, tid = threadIdx.x, nc = P.nc;
    const float4* stage = P.stage + (size_t)b * P.stage_cap * 2;
    const int n_staged = P.stage_off[(size_t)b * (nc + 1) + nc];
    float* out = P.out + (size_t)b * P.out_cap * YOLO_B200_DET_COLS;
    int32_t* out_row = P.out_row + (size_t)b * P.out_cap;

    // rows stay in global memory (L2: the segment kernel has just written them); their keys are sorted in registers,
    // 2 .. 16 per thread depending on how many rows the image staged (block-uniform), or, beyond final_smem_keys rows,
    // by a network over a global key array
    __shared__ int s_count;
    unsigned long long* keys = skeys;
    int n_out;
    if (n_staged <= P.final_smem_keys) {
        if (n_staged <= 2 * kFinalThreads)      n_out = sort_staged_keys<2, kFinalThreads>(stage, n_staged, skeys, &s_count);
        else if (n_staged <= 4 * kFinalThreads) n_out = sort_staged_keys<4, kFinalThreads>(stage, n_staged, skeys, &s_count);
        else if (n_staged <= 8 * kFinalThreads) n_out = sort_staged_keys<8, kFinalThreads>(stage, n_staged, skeys, &s_count);
        else                                    n_out = sort_staged_keys<16, kFinalThreads>(stage, n_staged, skeys, &s_count);
        if (tid == 0) P.out_count[b] = n_out;
        if (n_out == 0) return;
    } else {
        // more rows than the register sort takes: the generic network, over the shared-memory key area when the rows fit it
        // (small CTAs: up to kFinalSmallCap keys), over the global key array otherwise
        if (n_staged > P.final_key_slots) keys = P.final_keys + (size_t)b * P.stage_cap;
        int mine = 0;
        for (int q = tid; q < n_staged; q += kFinalThreads) {
            const unsigned long long k = final_key(stage[2 * q + 1], q);
            keys[q] = k;
            mine += (k != ~0ull);
        }
        if (tid == 0) s_count = 0;
        __syncthreads();
        if (mine) atomicAdd(&s_count, mine);
        __syncthreads();
        n_out = s_count;
        if (tid == 0) P.out_count[b] = n_out;
        if (n_out == 0) return;
        bitonic_sort<false, kFinalThreads>(keys, nullptr, n_staged);
    }
    // result rows in chunks of kFinalThreads: gathered into shared memory behind the key array, then stored as
    // contiguous 16-byte vectors (same peer-friendly store pattern as the fast path)
    float* cflat = reinterpret_cast<float*>(skeys + P.final_key_slots);                // 16-byte aligned, [7*kFinalThreads + 4]
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
__global__ void __launch_bounds__(kFinalThreads, 65536 / 64 / kFinalThreads)
nms_finalize_kernel(const __grid_constant__ NmsParams P) {
    nms_finalize_body<kFinalThreads>(P);
    signal_step(P);
}

}  // namespace yb

// ================================================================================================
using namespace yb;

namespace {
struct WsLayout {
    size_t bucket_key, bucket_slot, seg_off, stage_off, work_count, group_list, seg_list, stage, final_keys, total;
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
    L.work_count = o;  o = align_up(o + 16);
    L.group_list = o;  o = align_up(o + (size_t)batch * nc * sizeof(int2));
    L.seg_list = o;    o = align_up(o + (size_t)batch * nc * sizeof(int4));
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
    // a rank without images (global batch smaller than the world size) still owes the gather its completion stamp
    if (batch == 0) return opts && opts->step_stamp ? yolo_b200_flag_post(opts->step_stamp, opts->step_seq, 0, stream) : 0;

    unsigned char* ws = static_cast<unsigned char*>(workspace);
    NmsParams P{};
    P.cand_box = cand_box; P.cand_meta = cand_meta; P.count = count;
    P.batch = batch; P.cap = cap_per_img; P.nc = nc; P.mpc = max_per_class; P.stage_cap = stage_cap; P.out_cap = out_cap;
    P.nms_thres = nms_thres;
    P.bucket_key = reinterpret_cast<unsigned long long*>(ws + L.bucket_key);
    P.bucket_slot = reinterpret_cast<uint32_t*>(ws + L.bucket_slot);
    P.seg_off = reinterpret_cast<int32_t*>(ws + L.seg_off);
    P.stage_off = reinterpret_cast<int32_t*>(ws + L.stage_off);
    P.work_count = reinterpret_cast<int32_t*>(ws + L.work_count);
    P.group_list = reinterpret_cast<int2*>(ws + L.group_list);
    P.seg_list = reinterpret_cast<int4*>(ws + L.seg_list);
    P.stage = reinterpret_cast<float4*>(ws + L.stage);
    P.final_keys = reinterpret_cast<unsigned long long*>(ws + L.final_keys);
    P.out = out; P.out_row = out_row; P.out_count = out_count;
    P.step_seq = opts ? opts->step_seq : nullptr;
    P.step_stamp = opts ? opts->step_stamp : nullptr;

    cudaError_t e;
    // small segments share warps when the reference's per-class cap cannot cut them (max_per_class >= 32) and a candidate
    // slot fits next to the segment id in one word
    P.pack_ok = max_per_class >= kSmallSeg && cap_per_img <= (1 << kPackSlotBits) && nc <= 0xffff;
    if ((e = cudaMemsetAsync(P.work_count, 0, 4 * sizeof(int32_t), stream)) != cudaSuccess) return (int)e;
    const size_t bucket_smem = (size_t)(5 * nc + 1) * sizeof(int);
    const bool small_bucket = cap_per_img <= 4096;
    const void* bucket_fn = small_bucket ? (const void*)bucket_by_class_kernel<kBucketThreadsSmall>
                                         : (const void*)bucket_by_class_kernel<kBucketThreadsBig>;
    if (bucket_smem > 48 * 1024 &&
        (e = cudaFuncSetAttribute(bucket_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bucket_smem)) != cudaSuccess)
        return (int)e;
    const int carve = yb_carveout_for(cap_per_img);               // one preference for every kernel of the call (common.cuh)
    if (small_bucket) yb_prefer_carveout(bucket_by_class_kernel<kBucketThreadsSmall>, carve);
    else              yb_prefer_carveout(bucket_by_class_kernel<kBucketThreadsBig>, carve);
    yb_prefer_carveout(nms_segment_kernel, carve);
    if (small_bucket) bucket_by_class_kernel<kBucketThreadsSmall><<<batch, kBucketThreadsSmall, bucket_smem, stream>>>(P);
    else              bucket_by_class_kernel<kBucketThreadsBig><<<batch, kBucketThreadsBig, bucket_smem, stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // Segment stage: one launch of single-warp CTAs over the bucket kernel's two work lists (their lengths are only known on
    // the device: CTAs beyond a list's end leave at once) -- half of the grid for the groups of small segments, half for the
    // bigger segments.  At most seg_warps_per_sm resident CTAs per SM and list (default 16: with several batches in flight the
    // register file is shared with the next batch's decode CTAs, and half of it for this kernel gave the best step time; 32 is
    // fastest when it runs alone -- profiles/r02_c_segment_residency.txt)
    const long long segs = (long long)batch * nc;
    const int seg_residency = opts && opts->seg_warps_per_sm >= 1 && opts->seg_warps_per_sm <= 32 ? opts->seg_warps_per_sm : 16;
    const long long seg_max = (long long)sms * seg_residency;
    const int seg_ctas = (int)(segs < seg_max ? segs : seg_max);
    P.packed_ctas = P.pack_ok ? seg_ctas : 0;
    nms_segment_kernel<<<seg_ctas + P.packed_ctas, kSegThreads, 0, stream>>>(P);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;

    // finalize: small CTAs when an image cannot stage many rows (more images resident per SM), big ones otherwise
    const bool small_final = stage_cap <= kFinalSmallCap;
    const int ft = small_final ? kFinalThreadsSmall : kFinalThreadsBig;
    P.final_smem_keys = stage_cap < ft * kFinalKpt ? stage_cap : ft * kFinalKpt;
    int kpt = 2;
    while (kpt * ft < P.final_smem_keys) kpt <<= 1;
    P.final_key_slots = kpt * ft;                                 // the register sort exchanges threads x keys-per-thread keys
    if (small_final && P.final_key_slots < stage_cap) P.final_key_slots = (stage_cap + 1) & ~1;   // every image's keys fit shared memory
    const size_t out_block = (size_t)(8 * ft + 8) * sizeof(float);               // (rows x 7) chunk + row ids, with phase slack
    const size_t final_smem = (size_t)P.final_key_slots * 8 + out_block;
    if (small_final) {
        yb_prefer_carveout(nms_finalize_kernel<kFinalThreadsSmall>, carve);
        if ((e = cudaFuncSetAttribute(nms_finalize_kernel<kFinalThreadsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)final_smem)) != cudaSuccess)
            return (int)e;
        nms_finalize_kernel<kFinalThreadsSmall><<<batch, kFinalThreadsSmall, final_smem, stream>>>(P);
    } else {
        yb_prefer_carveout(nms_finalize_kernel<kFinalThreadsBig>, carve);
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
