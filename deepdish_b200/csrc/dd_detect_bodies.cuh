// dd_detect_bodies.cuh -- detector post-processing bodies (group-generic, see dd_common.cuh).
//
//   dd_nms_frame        one CTA per frame: preprocessing.non_max_suppression as sort + bitmask + scan
//   dd_yolo_rows        one thread per anchor row of a staged tile: YOLOv5 head decode + box filter
//   dd_order_frame      one CTA per frame: put the appended candidates back into anchor order
#pragma once
#include "dd_common.cuh"
#include "../../include/deepdish_b200.h"

// ---- order-preserving f32 <-> u32 key (ascending key == ascending float) ---------------------
DD_HD unsigned dd_f32_key(float f) {
    union { float f; unsigned u; } c;
    c.f = f;
    return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}

// In-place ascending bitonic sort of P (power of two) 64-bit keys by the whole group.
template <class G>
DD_HD void dd_bitonic_sort(const G& g, unsigned long long* keys, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = g.lane; i < P; i += g.nl) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            g.sync();
        }
    }
}

DD_HD int dd_next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// Shared-memory plan of dd_nms_frame for up to n candidates.
struct DDNmsSmem {
    unsigned long long* keys;   // [P]
    unsigned long long* mask;   // [n * nw]
    unsigned long long* remv;   // [nw]
    double *x1, *y1, *x2, *y2, *area;   // [n] in sorted order
};
DD_HD size_t dd_nms_smem_bytes(int nmax) {
    const int P = dd_next_pow2(nmax), nw = (nmax + 63) / 64;
    return (size_t)P * 8 + (size_t)nmax * nw * 8 + (size_t)nw * 8 + (size_t)nmax * 5 * 8;
}
DD_HD void dd_nms_carve(char* mem, int nmax, DDNmsSmem& m) {
    const int P = dd_next_pow2(nmax), nw = (nmax + 63) / 64;
    m.keys = (unsigned long long*)mem;
    m.mask = m.keys + P;
    m.remv = m.mask + (size_t)nmax * nw;
    m.x1 = (double*)(m.remv + nw);
    m.y1 = m.x1 + nmax; m.x2 = m.y1 + nmax; m.y2 = m.x2 + nmax; m.area = m.y2 + nmax;
}

// deep_sort/preprocessing.py:6-73.  boxes f64 [n,4] tlwh, scores f32 [n].  keep[] receives the
// original indices in pick order (descending score); returns their number through *nkeep.
//   sort     : descending score (ties: lower index first -- the reference's np.argsort is unstable, so
//              callers keep scores unique);
//   bitmask  : bit j of row i = candidate j (ranked after i) has inter(i,j) / area(j) > max_overlap,
//              with the +1 pixel convention, all in f64 like boxes.astype(float);
//   scan     : serial over ranks, a candidate survives unless an earlier survivor set its bit.
template <class G>
DD_HD void dd_nms_frame(const G& g, const double* boxes, const float* scores, int n, int nmax,
                        double max_overlap, int* keep, int* nkeep, char* smem) {
    if (n > nmax) n = nmax;
    if (n <= 0) {
        if (g.lane == 0) *nkeep = 0;
        return;
    }
    DDNmsSmem m;
    dd_nms_carve(smem, nmax, m);
    const int P = dd_next_pow2(n), nw = (n + 63) / 64;
    for (int i = g.lane; i < P; i += g.nl) {
        unsigned long long k = ~0ull;
        if (i < n) k = ((unsigned long long)(~dd_f32_key(scores[i])) << 32) | (unsigned)i;
        m.keys[i] = k;
    }
    g.sync();
    dd_bitonic_sort(g, m.keys, P);
    for (int r = g.lane; r < n; r += g.nl) {
        const int i = (int)(m.keys[r] & 0xffffffffu);
        const double x = boxes[i * 4 + 0], y = boxes[i * 4 + 1], w = boxes[i * 4 + 2], h = boxes[i * 4 + 3];
        const double xx2 = dd_add(w, x), yy2 = dd_add(h, y);
        m.x1[r] = x; m.y1[r] = y; m.x2[r] = xx2; m.y2[r] = yy2;
        m.area[r] = dd_mul(dd_add(dd_sub(xx2, x), 1.0), dd_add(dd_sub(yy2, y), 1.0));
    }
    for (int w = g.lane; w < nw; w += g.nl) m.remv[w] = 0;
    g.sync();
    for (int e = g.lane; e < n * nw; e += g.nl) {
        const int i = e / nw, w = e - i * nw;
        unsigned long long bits = 0;
        if (w * 64 + 63 > i) {
            const double ax1 = m.x1[i], ay1 = m.y1[i], ax2 = m.x2[i], ay2 = m.y2[i];
            const int j1 = dd_imin(n, w * 64 + 64);
            for (int j = dd_imax(i + 1, w * 64); j < j1; ++j) {
                const double iw = dd_max(0.0, dd_add(dd_sub(dd_min(ax2, m.x2[j]), dd_max(ax1, m.x1[j])), 1.0));
                const double ih = dd_max(0.0, dd_add(dd_sub(dd_min(ay2, m.y2[j]), dd_max(ay1, m.y1[j])), 1.0));
                if (dd_div(dd_mul(iw, ih), m.area[j]) > max_overlap) bits |= 1ull << (j - w * 64);
            }
        }
        m.mask[e] = bits;
    }
    g.sync();
    // scan: jump from survivor to survivor (first zero bit of remv at or after the cursor)
    if (g.lane == 0) {
        int nk = 0;
        int i = 0;
        while (i < n) {
            const unsigned long long word = ~m.remv[i >> 6] & (~0ull << (i & 63));
            if (!word) { i = ((i >> 6) + 1) << 6; continue; }
#if defined(__CUDA_ARCH__)
            i = (i & ~63) + (__ffsll((long long)word) - 1);
#else
            i = (i & ~63) + __builtin_ctzll(word);
#endif
            if (i >= n) break;
            keep[nk++] = (int)(m.keys[i] & 0xffffffffu);
            for (int w = i >> 6; w < nw; ++w) m.remv[w] |= m.mask[(size_t)i * nw + w];
            ++i;
        }
        *nkeep = nk;
    }
}

// ------------------------------------------------------------------------------------------------
// YOLOv5 head row decode (tools/yolov5.py:120-146) + box filter (deepdish.py:946-955).
// row: 5+nc f32 values (x, y, w, h normalised, obj, cls...).  Returns true when the row survives
// every filter; the emitted box is the integer-valued tlwh after clipping.
// ------------------------------------------------------------------------------------------------
struct DDYoloParams {
    int nc;
    float thr;
    float img_w, img_h;       // PIL image size the head is scaled to (yolov5.py:131)
    int frame_w, frame_h;     // camera viewport of the box filter (deepdish.py:945)
    double max_area;          // 0.9 * frame_w * frame_h
};

template <class Row>
DD_HD bool dd_yolo_row(const Row& row, const DDYoloParams& p, const unsigned char* wanted,
                       double* out_tlwh, float* out_score, int* out_class, bool* out_nan) {
    const float obj = row(4);
    float best = dd_mulf(row(5), obj);
    int bi = 0;
    for (int c = 1; c < p.nc; ++c) {
        const float v = dd_mulf(row(5 + c), obj);
        if (v > best) { best = v; bi = c; }              // np.argmax: first maximum wins
    }
    if (!(best >= p.thr)) return false;                  // yolov5.py:130
    if (!wanted[bi]) return false;                       // yolov5.py:139
    const float cx = row(0), cy = row(1), w = row(2), h = row(3);
    const float hx = dd_divf(w, 2.0f), hy = dd_divf(h, 2.0f);
    const float x1 = dd_mulf(dd_subf(cx, hx), p.img_w), y1 = dd_mulf(dd_subf(cy, hy), p.img_h);
    const float x2 = dd_mulf(dd_addf(cx, hx), p.img_w), y2 = dd_mulf(dd_addf(cy, hy), p.img_h);
    const float bw = dd_subf(x2, x1), bh = dd_subf(y2, y1);       // yolov5.py:141-142
    *out_nan = (x1 != x1) || (y1 != y1) || (bw != bw) || (bh != bh);
    // deepdish.py:950-951: int(np.clip(...)) truncates toward zero
    const float fx = x1 < 0.f ? 0.f : (x1 > (float)p.frame_w ? (float)p.frame_w : x1);
    const float fy = y1 < 0.f ? 0.f : (y1 > (float)p.frame_h ? (float)p.frame_h : y1);
    const int ix = (int)fx, iy = (int)fy;
    const float mw = (float)(p.frame_w - ix), mh = (float)(p.frame_h - iy);
    const float fw = bw < 0.f ? 0.f : (bw > mw ? mw : bw);
    const float fh = bh < 0.f ? 0.f : (bh > mh ? mh : bh);
    const int iw = (int)fw, ih = (int)fh;
    if ((double)((long long)iw * ih) > p.max_area) return false;  // deepdish.py:953
    out_tlwh[0] = ix; out_tlwh[1] = iy; out_tlwh[2] = iw; out_tlwh[3] = ih;
    *out_score = best;
    *out_class = bi;
    return true;
}
