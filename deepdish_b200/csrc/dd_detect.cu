// dd_detect.cu -- sm_100a kernels + C ABI of the detector post-processing (YOLOv5 head decode +
// box filter, NMS).
//
//   k_yolo_decode  grid (tiles, frames): a CTA stages a tile of DD_YOLO_ROWS anchor rows in shared
//                  memory with coalesced 16-byte loads (the head is read exactly once: 8.568 MB per
//                  640x640 frame), one thread decodes one row (row stride 85 words is odd -> no bank
//                  conflicts) and survivors are appended through one atomic per frame.
//   k_yolo_order   one CTA per frame: sort the appended candidates back into anchor order (the order
//                  the reference's Python loop emits them, tools/yolov5.py:137-146) and apply the
//                  whole-frame NaN rule of deepdish.py:947-949.
//   k_nms          one CTA per frame: sort by score, N x N/64 suppression bitmask, serial scan.
#include <cuda_runtime.h>
#include "dd_detect_bodies.cuh"
#include "dd_tma.cuh"

#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

#define DD_YOLO_ROWS 128        // anchor rows per CTA, f32 head (43.5 KB tile)
#define DD_YOLO_ROWS_U8 512     // u8 head: same tile bytes, 4x the rows

struct RowF32 {
    const float* p;
    __device__ float operator()(int k) const { return p[k]; }
};
struct RowU8 {
    const unsigned char* p;
    float scale, zp;
    __device__ float operator()(int k) const { return dd_mulf(dd_subf((float)p[k], zp), scale); }
};

template <bool U8, int ROWS>
__global__ void __launch_bounds__(ROWS)
k_yolo_decode(const void* __restrict__ head, float scale, int zero_point, int na, DDYoloParams P,
              const unsigned char* __restrict__ wanted, int ncap, double* __restrict__ out_tlwh,
              float* __restrict__ out_score, int* __restrict__ out_class, int* __restrict__ out_anchor,
              int* __restrict__ out_count, int* __restrict__ out_flags) {
    extern __shared__ __align__(16) char smem[];
    const int frame = blockIdx.y;
    const int row0 = blockIdx.x * ROWS;
    const int rows = min(ROWS, na - row0);
    const int rw = 5 + P.nc;
    const size_t esz = U8 ? 1 : 4;
    const size_t row_bytes = (size_t)rw * esz;
    const char* src = (const char*)head + ((size_t)frame * na + row0) * row_bytes;
    const size_t bytes = (size_t)rows * row_bytes;
    __shared__ unsigned long long bar;
    if ((((uintptr_t)src) & 15) == 0 && (bytes & 15) == 0) {
        // one bulk asynchronous copy (TMA engine) brings the whole tile; no registers are tied up by
        // loads in flight and every resident CTA keeps its full tile outstanding
        if (threadIdx.x == 0) {
            dd_mbar_init(&bar, 1);
            dd_mbar_fence_init();
            dd_mbar_expect_tx(&bar, (unsigned)bytes);
            dd_bulk_g2s(smem, src, (unsigned)bytes, &bar);
        }
        __syncthreads();                 // barrier initialised before anyone polls it
        dd_mbar_wait(&bar, 0);
    } else if (U8) {
        for (int i = threadIdx.x; i < (int)bytes; i += ROWS) smem[i] = src[i];
    } else {
        const float* sf = (const float*)src;
        float* df = (float*)smem;
        for (int i = threadIdx.x; i < rows * rw; i += ROWS) df[i] = sf[i];
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r >= rows) return;
    double tlwh[4];
    float score;
    int cls;
    bool isnan = false, ok;
    if (U8) {
        RowU8 row{(const unsigned char*)smem + (size_t)r * rw, scale, (float)zero_point};
        // conf = dequant(cls) * dequant(obj) is linear in the class byte: its maximum over all bytes is at
        // byte 0 or 255, so a row whose objectness cannot reach the threshold even then is dropped after
        // reading one byte (the products are evaluated with the very same f32 operations as the full scan)
        const float obj = row(4);
        const float hi0 = dd_mulf(dd_mulf(dd_subf(0.f, (float)zero_point), scale), obj);
        const float hi1 = dd_mulf(dd_mulf(dd_subf(255.f, (float)zero_point), scale), obj);
        ok = false;
        if (hi0 >= P.thr || hi1 >= P.thr || hi0 != hi0 || hi1 != hi1)
            ok = dd_yolo_row(row, P, wanted, tlwh, &score, &cls, &isnan);
    } else {
        RowF32 row{(const float*)smem + (size_t)r * rw};
        ok = dd_yolo_row(row, P, wanted, tlwh, &score, &cls, &isnan);
    }
    if (isnan) atomicOr(out_flags + frame, 0x100);
    if (!ok) return;
    const int pos = atomicAdd(out_count + frame, 1);
    if (pos >= ncap) {
        atomicOr(out_flags + frame, DD_FLAG_DET_OVERFLOW);
        return;
    }
    const size_t o = (size_t)frame * ncap + pos;
    out_tlwh[o * 4 + 0] = tlwh[0]; out_tlwh[o * 4 + 1] = tlwh[1];
    out_tlwh[o * 4 + 2] = tlwh[2]; out_tlwh[o * 4 + 3] = tlwh[3];
    out_score[o] = score;
    out_class[o] = cls;
    out_anchor[o] = row0 + r;
}

// Re-order the unordered appends of one frame by anchor index.  Shared: keys[P] + payload copy.
__global__ void __launch_bounds__(512)
k_yolo_order(int ncap, double* __restrict__ out_tlwh, float* __restrict__ out_score,
             int* __restrict__ out_class, int* __restrict__ out_anchor, int* __restrict__ out_count,
             int* __restrict__ out_flags) {
    extern __shared__ __align__(16) char smem[];
    const int frame = blockIdx.x;
    BlockG g;
    int n = out_count[frame];
    if (n > ncap) n = ncap;
    const int flags = out_flags[frame];
    __syncthreads();
    if (flags & 0x100) {                       // NaN anywhere in the frame's boxes: drop all
        if (threadIdx.x == 0) { out_count[frame] = 0; out_flags[frame] = flags & ~0x100; }
        return;
    }
    if (threadIdx.x == 0) out_count[frame] = n;
    if (n <= 1) return;
    // rank of a candidate = number of candidates with a smaller anchor index (anchors are distinct)
    int* an = (int*)smem;
    double* tl = (double*)(an + ((n + 1) & ~1));
    float* sc = (float*)(tl + (size_t)n * 4);
    int* cl = (int*)(sc + n);
    const size_t base = (size_t)frame * ncap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        an[i] = out_anchor[base + i];
        tl[i * 4 + 0] = out_tlwh[(base + i) * 4 + 0]; tl[i * 4 + 1] = out_tlwh[(base + i) * 4 + 1];
        tl[i * 4 + 2] = out_tlwh[(base + i) * 4 + 2]; tl[i * 4 + 3] = out_tlwh[(base + i) * 4 + 3];
        sc[i] = out_score[base + i];
        cl[i] = out_class[base + i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int a = an[i];
        int r = 0;
        for (int j = 0; j < n; ++j) r += an[j] < a;
        out_tlwh[(base + r) * 4 + 0] = tl[i * 4 + 0]; out_tlwh[(base + r) * 4 + 1] = tl[i * 4 + 1];
        out_tlwh[(base + r) * 4 + 2] = tl[i * 4 + 2]; out_tlwh[(base + r) * 4 + 3] = tl[i * 4 + 3];
        out_score[base + r] = sc[i];
        out_class[base + r] = cl[i];
        out_anchor[base + r] = a;
    }
}

__global__ void __launch_bounds__(1024)
k_nms(const double* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ counts,
      int nmax, double max_overlap, int* __restrict__ out_keep, int* __restrict__ out_nkeep) {
    extern __shared__ __align__(16) char smem[];
    const int f = blockIdx.x;
    BlockG g;
    dd_nms_frame(g, boxes + (size_t)f * nmax * 4, scores + (size_t)f * nmax, counts[f], nmax,
                 max_overlap, out_keep + (size_t)f * nmax, out_nkeep + f, smem);
}

// ---- SSD-MobileNet: one CTA per frame -------------------------------------------------------------
//   phase A  the [na, ncls] score rows stream through a 3-stage shared-memory ring of 64-row tiles filled by
//            TMA bulk copies (cp.async.bulk + one mbarrier per stage; rows are only 4-byte aligned, so each
//            copy starts at the 16-byte boundary below its tile and the row addressing absorbs the offset);
//            4 threads per anchor find the best non-background class.  Only anchors whose best score reaches
//            the reference's confidence threshold are kept (compacted into shared memory and decoded, expf):
//            the op processes candidates in descending score and a candidate can only be suppressed by a
//            higher-scored pick, so every op selection below the confidence threshold comes after all the
//            ones above it and is dropped by tools/ssd_mobilenet.py:119 anyway -- the result is identical;
//   phase B  greedy NMS as <= max_det rounds of {arg-max over the live candidates, eager suppression of
//            every live candidate with IoU > 0.6 against the pick} -- the same selection as the op's
//            sort + scan, without a sort;
//   phase C  thread 0: the reference's own post-processing on the <= 10 selected boxes (dd_ssd_post).
#define DD_SSD_THREADS 256
#define DD_SSD_TILE 64           // anchor rows per stage (= DD_SSD_THREADS / 4)
#define DD_SSD_STAGES 3
#define DD_SSD_CAND 1024         // candidates above the confidence threshold kept per frame

__device__ __forceinline__ size_t dd_ssd_stage_bytes(int ncls) { return ((size_t)DD_SSD_TILE * ncls * 4 + 16 + 15) & ~(size_t)15; }

__global__ void __launch_bounds__(DD_SSD_THREADS)
k_ssd_decode(const float* __restrict__ raw_boxes, const float* __restrict__ raw_scores, size_t scores_bytes,
             const float* __restrict__ anchors, DDSsdParams P, const int* __restrict__ class_to_label,
             int ncap, double* __restrict__ out_tlwh, float* __restrict__ out_score,
             int* __restrict__ out_label, int* __restrict__ out_count, int* __restrict__ out_flags) {
    extern __shared__ __align__(128) char smem[];
    const int frame = blockIdx.x;
    const int na = P.na, ncls = P.ncls;
    const size_t stage_bytes = dd_ssd_stage_bytes(ncls);
    unsigned long long* keys = (unsigned long long*)(smem + DD_SSD_STAGES * stage_bytes);   // [DD_SSD_CAND]
    float* dec = (float*)(keys + DD_SSD_CAND);                                              // [DD_SSD_CAND][4]
    int* bcls = (int*)(dec + (size_t)DD_SSD_CAND * 4);                                      // [DD_SSD_CAND]
    __shared__ float sel_box[DD_SSD_MAXDET * 4];
    __shared__ int sel_cls[DD_SSD_MAXDET];
    __shared__ float sel_score[DD_SSD_MAXDET];
    __shared__ unsigned long long wmin[DD_SSD_THREADS / 32];
    __shared__ unsigned long long bars[DD_SSD_STAGES];
    __shared__ int n_cand, pick_slot;
    const char* all_scores = (const char*)raw_scores;
    const size_t frame_off = (size_t)frame * na * ncls * 4;
    const float* fb = raw_boxes + (size_t)frame * na * 4;
    const int ntiles = (na + DD_SSD_TILE - 1) / DD_SSD_TILE;

    auto issue = [&](int t) {            // thread 0: bulk copy of tile t into stage t % STAGES
        const size_t off = frame_off + (size_t)t * DD_SSD_TILE * ncls * 4;
        const size_t lo = off & ~(size_t)15;                             // 16-byte boundary below the tile
        const int rows = min(DD_SSD_TILE, na - t * DD_SSD_TILE);
        size_t hi = (off + (size_t)rows * ncls * 4 + 15) & ~(size_t)15;
        if (hi > scores_bytes) hi = scores_bytes & ~(size_t)15;          // never read past the buffer
        const int st = t % DD_SSD_STAGES;
        dd_mbar_expect_tx(&bars[st], (unsigned)(hi - lo));
        dd_bulk_g2s(smem + st * stage_bytes, all_scores + lo, (unsigned)(hi - lo), &bars[st]);
    };
    if (threadIdx.x == 0) {
        n_cand = 0;
        for (int i = 0; i < DD_SSD_STAGES; ++i) dd_mbar_init(&bars[i], 1);
        dd_mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int t = 0; t < DD_SSD_STAGES - 1 && t < ntiles; ++t) issue(t);
    const float keep_thr = fmaxf(P.score_thr, P.conf_thr);
    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;                // 4 threads per anchor row
    for (int t = 0; t < ntiles; ++t) {
        if (threadIdx.x == 0 && t + DD_SSD_STAGES - 1 < ntiles) issue(t + DD_SSD_STAGES - 1);
        const int st = t % DD_SSD_STAGES;
        dd_mbar_wait(&bars[st], (unsigned)((t / DD_SSD_STAGES) & 1));
        const size_t off = frame_off + (size_t)t * DD_SSD_TILE * ncls * 4;
        const float* tile = (const float*)(smem + st * stage_bytes + (off & 15));
        const int a = t * DD_SSD_TILE + r;
        float best = -3.0e38f;
        int bi = 0x7fffffff;
        if (a < na) {
            const float* row = tile + r * ncls;
            // the last bytes of the very last tile may lie beyond a clamped copy: they are read from global
            const bool tail = frame_off + ((size_t)(a + 1) * ncls) * 4 > (scores_bytes & ~(size_t)15);
            if (tail) row = (const float*)(all_scores + frame_off) + (size_t)a * ncls;
#pragma unroll 4
            for (int c = 1 + q; c < ncls; c += 4) {                     // skip background column 0
                const float v = row[c];
                if (v > best) { best = v; bi = c - 1; }
            }
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {                              // first maximum wins across the 4 lanes
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (a < na && q == 0 && best >= keep_thr) {
            const int k = atomicAdd(&n_cand, 1);
            if (k < DD_SSD_CAND) {
                dd_ssd_decode_box(fb + (size_t)a * 4, anchors + (size_t)a * 4, P, dec + (size_t)k * 4);
                bcls[k] = bi;
                keys[k] = (((unsigned long long)(~dd_f32_key(best))) << 32) | (unsigned)a;
            }
        }
        __syncthreads();                                                // stage st may be refilled next iteration
    }
    int nc = n_cand;
    if (nc > DD_SSD_CAND) {                                             // never silently truncated
        nc = DD_SSD_CAND;
        if (threadIdx.x == 0) atomicOr(out_flags + frame, DD_FLAG_DET_OVERFLOW);
    }
    int ns = 0;
    for (; ns < P.max_det; ++ns) {
        unsigned long long mk = ~0ull;                                  // smallest key = best live candidate
        int mi = -1;
        for (int k = threadIdx.x; k < nc; k += DD_SSD_THREADS)
            if (keys[k] < mk) { mk = keys[k]; mi = k; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(0xffffffffu, mk, o);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (ok < mk) { mk = ok; mi = oi; }
        }
        if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = mk;
        __syncthreads();
        unsigned long long bk = wmin[0];
#pragma unroll
        for (int w = 1; w < DD_SSD_THREADS / 32; ++w) bk = min(bk, wmin[w]);
        if (bk == ~0ull) break;                                         // uniform: no live candidate left
        if (mk == bk && (threadIdx.x & 31) == 0) pick_slot = mi;        // keys are distinct: exactly one warp
        __syncthreads();
        const int pick = pick_slot;
        const float pb[4] = {dec[pick * 4], dec[pick * 4 + 1], dec[pick * 4 + 2], dec[pick * 4 + 3]};
        if (threadIdx.x == 0) {
            for (int k = 0; k < 4; ++k) sel_box[ns * 4 + k] = pb[k];
            sel_cls[ns] = bcls[pick];
            union { unsigned u; float f; } cv;
            const unsigned kk = ~(unsigned)(bk >> 32);
            cv.u = (kk & 0x80000000u) ? (kk & 0x7fffffffu) : ~kk;
            sel_score[ns] = cv.f;
        }
        for (int k = threadIdx.x; k < nc; k += DD_SSD_THREADS) {
            if (keys[k] == ~0ull) continue;
            if (k == pick || dd_ssd_iou(pb, dec + (size_t)k * 4) > P.iou_thr) keys[k] = ~0ull;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const size_t o = (size_t)frame * ncap;
        double tl[DD_SSD_MAXDET * 4];
        float sc[DD_SSD_MAXDET];
        int lb[DD_SSD_MAXDET];
        int n = dd_ssd_post(sel_box, sel_cls, sel_score, ns, P, class_to_label, tl, sc, lb);
        if (n > ncap) n = ncap;
        for (int i = 0; i < n; ++i) {
            for (int k = 0; k < 4; ++k) out_tlwh[(o + i) * 4 + k] = tl[i * 4 + k];
            out_score[o + i] = sc[i];
            out_label[o + i] = lb[i];
        }
        out_count[frame] = n;
    }
}

// Kept candidates (NMS pick order) -> the tracker's padded detection batch for a range of streams:
// deepdish.py:996-998 (boxesA1 = boxesA0[indices], ...) + the Detection list of :1014 as SoA.
__global__ void __launch_bounds__(128)
k_gather_kept(const double* __restrict__ cand_tlwh, const float* __restrict__ cand_score,
              const int* __restrict__ cand_label, const int* __restrict__ label_map, int n_map, int ncap,
              const int* __restrict__ keep, const int* __restrict__ nkeep, int nmax, int dmax,
              double* __restrict__ det_tlwh, float* __restrict__ det_conf, int* __restrict__ det_label,
              int* __restrict__ det_count, int* __restrict__ out_flags) {
    const int f = blockIdx.x;
    int n = nkeep[f];
    if (n > dmax) {
        n = dmax;
        if (threadIdx.x == 0) atomicOr(out_flags + f, DD_FLAG_DET_OVERFLOW);
    }
    for (int k = threadIdx.x; k < dmax; k += blockDim.x) {
        const size_t o = (size_t)f * dmax + k;
        if (k < n) {
            const size_t c = (size_t)f * ncap + keep[(size_t)f * nmax + k];
            det_tlwh[o * 4 + 0] = cand_tlwh[c * 4 + 0]; det_tlwh[o * 4 + 1] = cand_tlwh[c * 4 + 1];
            det_tlwh[o * 4 + 2] = cand_tlwh[c * 4 + 2]; det_tlwh[o * 4 + 3] = cand_tlwh[c * 4 + 3];
            det_conf[o] = cand_score[c];
            const int l = cand_label[c];
            det_label[o] = (label_map && l >= 0 && l < n_map) ? label_map[l] : l;
        } else {
            det_tlwh[o * 4 + 0] = 0.0; det_tlwh[o * 4 + 1] = 0.0; det_tlwh[o * 4 + 2] = 0.0; det_tlwh[o * 4 + 3] = 0.0;
            det_conf[o] = 0.f;
            det_label[o] = 0;
        }
    }
    if (threadIdx.x == 0) det_count[f] = n;
}


// ---- TFLite object-detector adapter (tools/tflite_object_detector.py:234-295 + tools/tflite.py:26-41) -------------
// One warp per frame.  The detections that pass the score threshold (:254) and the deny / allow lists (:276-288)
// are ranked by a STABLE descending sort on the score (:270-273: ties keep the op's order) -- rank = #(higher score)
// + #(equal score, lower index) --, cut at max_results (:291-293), filtered by wanted_labels (tflite.py:31-35) and
// emitted in rank order as [left, top, right-left, bottom-top] with int() truncation of the float32 products
// (:256-260).
__global__ void __launch_bounds__(128)
k_tflite_post(const float* __restrict__ boxes, const float* __restrict__ classes, const float* __restrict__ scores,
              const int* __restrict__ count, int b, int n, float img_w, float img_h, float score_thr,
              const unsigned char* __restrict__ list_ok, const unsigned char* __restrict__ wanted, int n_labels,
              int max_results, int ncap, double* __restrict__ out_tlwh, float* __restrict__ out_score,
              int* __restrict__ out_label, int* __restrict__ out_count, int* __restrict__ out_flags) {
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= b) return;
    const float* sc = scores + (size_t)f * n;
    const float* cl = classes + (size_t)f * n;
    int cnt = count[f];
    if (cnt > n) cnt = n;
    int flags = 0, emitted = 0;
    // pass 1 is implicit: "alive" = passes threshold and lists; rank among alive entries
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        bool alive = false;
        float si = 0.f;
        int li = -1;
        if (i < cnt) {
            si = sc[i];
            li = (int)cl[i];
            if (si >= score_thr) {
                if (li < 0 || li >= n_labels) flags |= DD_FLAG_DET_OVERFLOW;       // IndexError in the reference
                else alive = list_ok[li] != 0;
            }
        }
        int rank = 0;
        if (alive) {
            for (int j = 0; j < cnt; ++j) {
                const float sj = sc[j];
                if (!(sj >= score_thr)) continue;
                const int lj = (int)cl[j];
                if (lj < 0 || lj >= n_labels || !list_ok[lj]) continue;
                rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
            }
            if (max_results > 0 && rank >= max_results) alive = false;
        }
        if (alive && wanted[li]) {
            // output position = number of emitted entries with a smaller rank: count them the same way
            int pos = 0;
            for (int j = 0; j < cnt; ++j) {
                const float sj = sc[j];
                if (!(sj >= score_thr)) continue;
                const int lj = (int)cl[j];
                if (lj < 0 || lj >= n_labels || !list_ok[lj] || !wanted[lj]) continue;
                pos += (sj > si || (sj == si && j < i)) ? 1 : 0;
            }
            if (pos < ncap) {
                const float* bx = boxes + ((size_t)f * n + i) * 4;
                const int top = (int)dd_mulf(bx[0], img_h), left = (int)dd_mulf(bx[1], img_w);
                const int bottom = (int)dd_mulf(bx[2], img_h), right = (int)dd_mulf(bx[3], img_w);
                double* o = out_tlwh + ((size_t)f * ncap + pos) * 4;
                o[0] = left; o[1] = top; o[2] = right - left; o[3] = bottom - top;
                out_score[(size_t)f * ncap + pos] = si;
                out_label[(size_t)f * ncap + pos] = li;
            } else {
                flags |= DD_FLAG_DET_OVERFLOW;
            }
            emitted += 1;
        }
    }
    emitted = __reduce_add_sync(0xffffffffu, emitted);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) {
        out_count[f] = emitted < ncap ? emitted : ncap;
        out_flags[f] = flags;
    }
}

// ---- Keras YOLOv3 adapter: one CTA per frame ------------------------------------------------------
//   phase A  all threads sweep the frame's boxes reading only the objectness logit: a class score is obj * cls <= obj,
//            so a box whose objectness sigmoid is not above the threshold has no non-zero class score -- it can neither
//            suppress nor be returned -- and its other 84 logits are never read;
//   phase B  the surviving ("active") boxes, in box order: one thread per (box, class) evaluates the thresholded class
//            scores, one thread per box decodes the integer box;
//   phase C  do_nms: one thread per class (the classes are independent);  phase D  thread 0 emits.
struct DDYolo3Maps { const float* map[3]; };

__global__ void __launch_bounds__(256)
k_yolo3_post(const DDYolo3Maps M, const DDYolo3Params P, const unsigned char* __restrict__ wanted, int ncap,
             double* __restrict__ out_box, float* __restrict__ out_score, int* __restrict__ out_label,
             int* __restrict__ out_count, int* __restrict__ out_flags) {
    extern __shared__ __align__(16) char smem[];
    const int f = blockIdx.x, tid = threadIdx.x;
    const int rw = 5 + P.nc;
    int* act = (int*)smem;                                   // [MAX_ACTIVE] global box index (map << 24 | cell * 3 + b)
    int* box = act + DD_Y3_MAX_ACTIVE;                       // [MAX_ACTIVE][4]
    int* order = box + DD_Y3_MAX_ACTIVE * 4;                 // [nc][MAX_ACTIVE] per-class sort scratch
    float* cls = (float*)(order + P.nc * DD_Y3_MAX_ACTIVE);  // [MAX_ACTIVE][nc]
    __shared__ int s_n, s_bad, s_over;
    if (tid == 0) { s_n = 0; s_bad = 0; s_over = 0; }
    __syncthreads();
    for (int k = 0; k < 3; ++k) {
        const int cells = P.g[k] * P.g[k];
        const float* base = M.map[k] + (size_t)f * cells * 3 * rw;
        for (int e = tid; e < cells * 3; e += blockDim.x) {
            const float obj = dd_sigmoid_f32(base[(size_t)e * rw + 4]);
            if (obj > P.thr) {
                const int pos = atomicAdd(&s_n, 1);
                if (pos < DD_Y3_MAX_ACTIVE) act[pos] = (k << 24) | e;
            }
        }
    }
    __syncthreads();
    int n = s_n;
    if (n > DD_Y3_MAX_ACTIVE) { n = DD_Y3_MAX_ACTIVE; if (tid == 0) s_over = 1; }
    if (tid == 0) {                                          // box order = the reference's append order (map, cell, anchor)
        for (int i = 1; i < n; ++i) {
            const int v = act[i];
            int j = i;
            while (j > 0 && act[j - 1] > v) { act[j] = act[j - 1]; --j; }
            act[j] = v;
        }
    }
    __syncthreads();
    for (int e = tid; e < n * P.nc; e += blockDim.x) {
        const int i = e / P.nc, c = e - i * P.nc;
        const int k = act[i] >> 24, cell = act[i] & 0xffffff;
        const float* raw = M.map[k] + ((size_t)f * P.g[k] * P.g[k] * 3 + cell) * rw;
        const float v = dd_mulf(dd_sigmoid_f32(raw[4]), dd_sigmoid_f32(raw[5 + c]));      // yolo.py:54-55
        cls[e] = v > P.thr ? v : 0.f;                                                     // :56
    }
    for (int i = tid; i < n; i += blockDim.x) {
        const int k = act[i] >> 24, cell = act[i] & 0xffffff;
        const float* raw = M.map[k] + ((size_t)f * P.g[k] * P.g[k] * 3 + cell) * rw;
        dd_yolo3_box(raw, k, cell / 3, cell % 3, P, box + i * 4);
    }
    __syncthreads();
    for (int c = tid; c < P.nc; c += blockDim.x) {
        int bad = 0;
        dd_y3_nms_class(c, n, P.nc, cls, box, P.nms_thresh, order + c * DD_Y3_MAX_ACTIVE, &bad);
        if (bad) s_bad = 1;
    }
    __syncthreads();
    if (tid == 0) {
        int over = s_over;
        const int k = dd_y3_emit(n, P.nc, cls, box, P.thr, wanted, ncap, out_box + (size_t)f * ncap * 4,
                                 out_score + (size_t)f * ncap, out_label + (size_t)f * ncap, &over);
        out_count[f] = k;
        out_flags[f] = (over ? DD_FLAG_DET_OVERFLOW : 0) | (s_bad ? DD_Y3_BAD_BOX : 0);
    }
}

extern "C" {

int dd_yolo3_decode(const float* map0, const float* map1, const float* map2, const int32_t* host_grids3,
                    const int32_t* host_anchors18, int32_t b, int32_t nc, const uint8_t* wanted, float score_thr,
                    double nms_thresh, int32_t image_w, int32_t image_h, int32_t net_w, int32_t net_h, int32_t ncap,
                    double* out_box, float* out_score, int32_t* out_label, int32_t* out_count, int32_t* out_flags,
                    void* stream) {
    if (!map0 || !map1 || !map2 || !host_grids3 || !host_anchors18 || !wanted || !out_box || !out_score || !out_label ||
        !out_count || !out_flags)
        return DD_ERR_INVALID;
    if (b < 0 || nc <= 0 || nc > 256 || ncap <= 0 || net_w <= 0 || net_h <= 0) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    DDYolo3Maps M;
    M.map[0] = map0; M.map[1] = map1; M.map[2] = map2;
    DDYolo3Params P;
    P.nc = nc; P.thr = score_thr; P.nms_thresh = nms_thresh;
    P.image_w = image_w; P.image_h = image_h; P.net_w = net_w; P.net_h = net_h;
    for (int k = 0; k < 3; ++k) {
        P.g[k] = host_grids3[k];
        if (P.g[k] <= 0 || (long long)P.g[k] * P.g[k] * 3 >= (1 << 24)) return DD_ERR_INVALID;
        for (int a = 0; a < 6; ++a) P.anchors[k][a] = host_anchors18[k * 6 + a];
    }
    const size_t smem = (size_t)DD_Y3_MAX_ACTIVE * 4 * (1 + 4 + nc) + (size_t)DD_Y3_MAX_ACTIVE * nc * 4;
    if (smem > 200 * 1024) return DD_ERR_INVALID;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_yolo3_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DD_ERR_CUDA;
    k_yolo3_post<<<b, 256, smem, (cudaStream_t)stream>>>(M, P, wanted, ncap, out_box, out_score, out_label, out_count,
                                                         out_flags);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_nms(const double* boxes, const float* scores, const int32_t* counts, int32_t b, int32_t nmax,
           double max_overlap, int32_t* out_keep, int32_t* out_nkeep, void* stream) {
    if (!boxes || !scores || !counts || !out_keep || !out_nkeep || b < 0 || nmax <= 0) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    const size_t smem = dd_nms_smem_bytes(nmax);
    if (smem > 227 * 1024) return DD_ERR_CAPACITY;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DD_ERR_CUDA;
    k_nms<<<b, 1024, smem, (cudaStream_t)stream>>>(boxes, scores, counts, nmax, max_overlap, out_keep, out_nkeep);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_yolo_decode(const void* head, int32_t head_is_u8, float scale, int32_t zero_point, int32_t b,
                   int32_t na, int32_t nc, const uint8_t* wanted, float score_thr, int32_t img_w,
                   int32_t img_h, int32_t frame_w, int32_t frame_h, int32_t ncap, double* out_tlwh,
                   float* out_score, int32_t* out_class, int32_t* out_anchor, int32_t* out_count,
                   int32_t* out_flags, void* stream) {
    if (!head || !wanted || !out_tlwh || !out_score || !out_class || !out_anchor || !out_count || !out_flags)
        return DD_ERR_INVALID;
    if (b < 0 || na <= 0 || nc <= 0 || ncap <= 0 || ncap > 4096) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(out_count, 0, sizeof(int) * b, st) != cudaSuccess) return DD_ERR_CUDA;
    if (cudaMemsetAsync(out_flags, 0, sizeof(int) * b, st) != cudaSuccess) return DD_ERR_CUDA;
    DDYoloParams P;
    P.nc = nc; P.thr = score_thr; P.img_w = (float)img_w; P.img_h = (float)img_h;
    P.frame_w = frame_w; P.frame_h = frame_h;
    P.max_area = 0.9 * frame_w * frame_h;                  // deepdish.py:953, left to right
    const int rows_per_cta = head_is_u8 ? DD_YOLO_ROWS_U8 : DD_YOLO_ROWS;
    const int tiles = (na + rows_per_cta - 1) / rows_per_cta;
    const size_t smem = (size_t)rows_per_cta * (5 + nc) * (head_is_u8 ? 1 : 4);
    if (smem > 227 * 1024) return DD_ERR_INVALID;
    dim3 grid(tiles, b);
    if (head_is_u8) {
        if (smem > 48 * 1024 && cudaFuncSetAttribute(k_yolo_decode<true, DD_YOLO_ROWS_U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DD_ERR_CUDA;
        k_yolo_decode<true, DD_YOLO_ROWS_U8><<<grid, DD_YOLO_ROWS_U8, smem, st>>>(head, scale, zero_point, na, P, wanted, ncap, out_tlwh, out_score, out_class, out_anchor, out_count, out_flags);
    } else {
        if (smem > 48 * 1024 && cudaFuncSetAttribute(k_yolo_decode<false, DD_YOLO_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DD_ERR_CUDA;
        k_yolo_decode<false, DD_YOLO_ROWS><<<grid, DD_YOLO_ROWS, smem, st>>>(head, scale, zero_point, na, P, wanted, ncap, out_tlwh, out_score, out_class, out_anchor, out_count, out_flags);
    }
    DD_CHECK_LAUNCH();
    const size_t osm = (size_t)(ncap + 2) * 4 + (size_t)ncap * (32 + 4 + 4);
    if (osm > 48 * 1024 && cudaFuncSetAttribute(k_yolo_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osm) != cudaSuccess) return DD_ERR_CUDA;
    k_yolo_order<<<b, 512, osm, st>>>(ncap, out_tlwh, out_score, out_class, out_anchor, out_count, out_flags);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_ssd_decode(const float* raw_boxes, const float* raw_scores, const float* anchors, int32_t b,
                  int32_t na, int32_t ncls, const int32_t* class_to_label, float conf_thr, double nms_iou,
                  int32_t img_w, int32_t img_h, int32_t frame_w, int32_t frame_h, int32_t ncap,
                  double* out_tlwh, float* out_score, int32_t* out_label, int32_t* out_count,
                  int32_t* out_flags, void* stream) {
    if (!raw_boxes || !raw_scores || !anchors || !class_to_label || !out_tlwh || !out_score || !out_label ||
        !out_count || !out_flags)
        return DD_ERR_INVALID;
    if (((uintptr_t)raw_scores & 15) != 0) return DD_ERR_INVALID;      /* bulk copies need a 16-byte aligned base */
    if (b < 0 || na <= 0 || na > 8192 || ncls < 2 || ncls > 1024 || ncap < 10) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    DDSsdParams P;
    P.na = na; P.ncls = ncls; P.max_det = 10; P.score_thr = 1e-8f; P.iou_thr = 0.6f;
    P.sy = 10.f; P.sx = 10.f; P.sh = 5.f; P.sw = 5.f;
    P.conf_thr = conf_thr; P.nms_iou = nms_iou;
    P.img_w = img_w; P.img_h = img_h; P.frame_w = frame_w; P.frame_h = frame_h;
    P.max_area = 0.9 * frame_w * frame_h;
    const size_t stage = ((size_t)DD_SSD_TILE * ncls * 4 + 16 + 15) & ~(size_t)15;
    const size_t smem = DD_SSD_STAGES * stage + (size_t)DD_SSD_CAND * (8 + 16 + 4);
    if (smem > 200 * 1024) return DD_ERR_CAPACITY;
    if (cudaMemsetAsync(out_flags, 0, sizeof(int) * b, (cudaStream_t)stream) != cudaSuccess) return DD_ERR_CUDA;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_ssd_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DD_ERR_CUDA;
    k_ssd_decode<<<b, DD_SSD_THREADS, smem, (cudaStream_t)stream>>>(raw_boxes, raw_scores,
                                                                   (size_t)b * na * ncls * 4, anchors, P,
                                                                   class_to_label, ncap, out_tlwh, out_score,
                                                                   out_label, out_count, out_flags);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

// The pre-NMS box filter as its own step (deepdish.py:941-960): one warp per frame; a NaN anywhere in the frame's
// boxes drops the whole frame, survivors keep their order.
__global__ void __launch_bounds__(128)
k_box_filter(const double* __restrict__ boxes, const int* __restrict__ counts, int b, int nmax, int frame_w, int frame_h,
             double max_area, double* __restrict__ out_tlwh, int* __restrict__ out_index, int* __restrict__ out_count) {
    const int f = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= b) return;
    int n = counts ? counts[f] : nmax;
    if (n > nmax) n = nmax;
    const double* bx = boxes + (size_t)f * nmax * 4;
    bool nan = false;
    for (int e = lane; e < n * 4; e += 32) nan = nan || (bx[e] != bx[e]);
    int nk = 0;
    if (!__any_sync(0xffffffffu, nan)) {
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            int ib[4];
            const bool ok = i < n && dd_box_clip(bx[i * 4], bx[i * 4 + 1], bx[i * 4 + 2], bx[i * 4 + 3], frame_w, frame_h,
                                                 max_area, ib);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const size_t o = (size_t)f * nmax + nk + __popc(m & ((1u << lane) - 1u));
                out_tlwh[o * 4] = ib[0]; out_tlwh[o * 4 + 1] = ib[1]; out_tlwh[o * 4 + 2] = ib[2]; out_tlwh[o * 4 + 3] = ib[3];
                out_index[o] = i;
            }
            nk += __popc(m);
        }
    }
    if (lane == 0) out_count[f] = nk;
}

int dd_box_filter(const double* boxes, const int32_t* counts, int32_t b, int32_t nmax, int32_t frame_w, int32_t frame_h,
                  double* out_tlwh, int32_t* out_index, int32_t* out_count, void* stream) {
    if (!boxes || !out_tlwh || !out_index || !out_count || b < 0 || nmax <= 0 || frame_w <= 0 || frame_h <= 0)
        return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    k_box_filter<<<(b + 3) / 4, 128, 0, (cudaStream_t)stream>>>(boxes, counts, b, nmax, frame_w, frame_h,
                                                               0.9 * frame_w * frame_h, out_tlwh, out_index, out_count);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_gather_detections(const double* cand_tlwh, const float* cand_score, const int32_t* cand_label,
                         const int32_t* label_map, int32_t n_map, int32_t ncap, const int32_t* keep,
                         const int32_t* nkeep, int32_t nmax, int32_t b, int32_t dmax, double* det_tlwh,
                         float* det_conf, int32_t* det_label, int32_t* det_count, int32_t* out_flags,
                         void* stream) {
    if (!cand_tlwh || !cand_score || !cand_label || !keep || !nkeep || !det_tlwh || !det_conf || !det_label ||
        !det_count || !out_flags || b < 0 || dmax <= 0 || ncap <= 0 || nmax <= 0)
        return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    k_gather_kept<<<b, 128, 0, (cudaStream_t)stream>>>(cand_tlwh, cand_score, cand_label, label_map, n_map, ncap, keep,
                                                      nkeep, nmax, dmax, det_tlwh, det_conf, det_label, det_count,
                                                      out_flags);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_tflite_postprocess(const float* op_boxes, const float* op_classes, const float* op_scores,
                          const int32_t* op_count, int32_t b, int32_t n, int32_t img_w, int32_t img_h,
                          float score_thr, const uint8_t* list_ok, const uint8_t* wanted, int32_t n_labels,
                          int32_t max_results, int32_t ncap, double* out_tlwh, float* out_score, int32_t* out_label,
                          int32_t* out_count, int32_t* out_flags, void* stream) {
    if (!op_boxes || !op_classes || !op_scores || !op_count || !list_ok || !wanted || !out_tlwh || !out_score ||
        !out_label || !out_count || !out_flags)
        return DD_ERR_INVALID;
    if (b < 0 || n <= 0 || n_labels <= 0 || ncap <= 0) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    k_tflite_post<<<(b + 3) / 4, 128, 0, (cudaStream_t)stream>>>(op_boxes, op_classes, op_scores, op_count, b, n,
                                                               (float)img_w, (float)img_h, score_thr, list_ok, wanted,
                                                               n_labels, max_results, ncap, out_tlwh, out_score,
                                                               out_label, out_count, out_flags);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

}  // extern "C"
