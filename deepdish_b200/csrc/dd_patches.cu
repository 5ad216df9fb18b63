// dd_patches.cu -- the step between NMS and the tracker's feature input:
// aspect-corrected crop + 8-bit bilinear resize of every detection box into the re-ID encoder's input
// tensor (tools/generate_detections.py:40-84, 198-205), and the reference's own arithmetic encoder
// (DummyImageEncoder, generate_detections.py:86-105).
//
// The resize reproduces OpenCV's 8-bit INTER_LINEAR fixed-point scheme bit for bit (cv2.resize is the
// third-party routine under extract_image_patch): 11-bit coefficients from float32 fractions, horizontal
// pass in int32, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
//
// Work split: a patch is covered by (patch_w / 4) x-groups x NB row bands; one thread owns 4 adjacent output
// pixels (12 bytes = three aligned 32-bit stores) and walks down its band.  The horizontally interpolated
// values of the two source rows it currently needs live in registers and are recomputed only when the row
// index changes, so an up-scaled box (the normal case: 30x80 -> 64x128) loads each source row once per
// thread instead of once per output row.  Frames are read through the read-only path; a frame's boxes
// overlap heavily, so the loads hit L1/L2 and HBM traffic is the frame once plus the patch writes.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/deepdish_b200.h"
#include "dd_common.cuh"

#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

#define DD_PATCH_THREADS 128

// generate_detections.py:64-80 -> crop rectangle [sx,ex) x [sy,ey) or false (the function returns None).
__device__ __forceinline__ bool dd_patch_box(const double* bx, int boxes_are_int, int ph, int pw, int img_h,
                                             int img_w, int& sx, int& sy, int& ex, int& ey) {
    const double aspect = dd_div((double)pw, (double)ph);                      // :66
    long long x0, y0, x1, y1;
    if (boxes_are_int) {        // int64 box array (deepdish.py:993): every item assignment truncates toward zero
        const long long x = (long long)bx[0], y = (long long)bx[1], w = (long long)bx[2], h = (long long)bx[3];
        const double nw = dd_mul(aspect, (double)h);                           // :67
        x0 = (long long)dd_sub((double)x, dd_div(dd_sub(nw, (double)w), 2.0)); // :68
        y0 = y;
        x1 = x0 + (long long)nw;                                               // :69, :72
        y1 = y + h;
    } else {                    // float box array: one truncation at astype(np.int) (:73)
        const double nw = dd_mul(aspect, bx[3]);
        const double x = dd_sub(bx[0], dd_div(dd_sub(nw, bx[2]), 2.0));
        x0 = (long long)x;
        y0 = (long long)bx[1];
        x1 = (long long)dd_add(nw, x);
        y1 = (long long)dd_add(bx[3], bx[1]);
    }
    x0 = x0 < 0 ? 0 : x0;                                                      // :76
    y0 = y0 < 0 ? 0 : y0;
    x1 = x1 > img_w - 1 ? img_w - 1 : x1;                                      // :77
    y1 = y1 > img_h - 1 ? img_h - 1 : y1;
    if (x0 >= x1 || y0 >= y1) return false;                                    // :78-79
    sx = (int)x0; sy = (int)y0; ex = (int)x1; ey = (int)y1;
    return true;
}

// cv2 resize coefficient for destination index d: source index pair and the two 11-bit weights.
// clamp_frac = true for x (cv2 zeroes the fraction at the borders), false for y (indices are clipped,
// the weights are kept).
__device__ __forceinline__ void dd_resize_coeff(int d, double scale, int sn, bool clamp_frac, int& i0, int& i1,
                                                int& a0, int& a1) {
    float f = (float)dd_sub(dd_mul(dd_add((double)d, 0.5), scale), 0.5);
    int s = (int)floorf(f);
    f = dd_subf(f, (float)s);
    if (clamp_frac) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        i0 = s;
        i1 = dd_imin(s + 1, sn - 1);
    } else {
        i0 = dd_imin(dd_imax(s, 0), sn - 1);
        i1 = dd_imin(dd_imax(s + 1, 0), sn - 1);
    }
    a0 = __float2int_rn(dd_mulf(dd_subf(1.f, f), 2048.f));
    a1 = __float2int_rn(dd_mulf(f, 2048.f));
}

struct DDHRow { int v[12]; };   // 4 pixels x 3 channels of (horizontal pass >> 4)

__device__ __forceinline__ void dd_hrow(const uint8_t* __restrict__ row, const int (&o0)[4], const int (&o1)[4],
                                        const int (&a0)[4], const int (&a1)[4], DDHRow& out) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out.v[k * 3 + c] = ((int)__ldg(row + o0[k] + c) * a0[k] + (int)__ldg(row + o1[k] + c) * a1[k]) >> 4;
}

__global__ void __launch_bounds__(DD_PATCH_THREADS)
k_extract_patches(const uint8_t* __restrict__ frames, int n_patches, int img_h, int img_w,
                  const double* __restrict__ boxes, const int* __restrict__ counts, int dmax, int boxes_are_int,
                  int ph, int pw, int xg, int nb, int rows_per_band, uint8_t* __restrict__ out_patches,
                  int* __restrict__ out_valid) {
    const int tpp = xg * nb;
    const int gt = blockIdx.x * (DD_PATCH_THREADS / tpp) * tpp + threadIdx.x;
    if (threadIdx.x >= (DD_PATCH_THREADS / tpp) * tpp) return;
    const int p = gt / tpp, t = gt - p * tpp;
    if (p >= n_patches) return;
    const int f = p / dmax, d = p - f * dmax;
    if (counts && d >= counts[f]) {
        if (t == 0) out_valid[p] = 0;
        return;
    }
    const int band = t / xg, gx = t - band * xg;
    const int y_begin = band * rows_per_band, y_end = dd_imin(ph, y_begin + rows_per_band);
    uint32_t* outw = (uint32_t*)(out_patches + (size_t)p * ph * pw * 3);
    int sx, sy, ex, ey;
    if (!dd_patch_box(boxes + (size_t)p * 4, boxes_are_int, ph, pw, img_h, img_w, sx, sy, ex, ey)) {
        if (t == 0) out_valid[p] = 0;
        for (int y = y_begin; y < y_end; ++y)
#pragma unroll
            for (int k = 0; k < 3; ++k) outw[((size_t)y * pw + gx * 4) * 3 / 4 + k] = 0u;
        return;
    }
    if (t == 0) out_valid[p] = 1;
    const int sw = ex - sx, sh = ey - sy;
    const double scale_x = dd_div(1.0, dd_div((double)pw, (double)sw));
    const double scale_y = dd_div(1.0, dd_div((double)ph, (double)sh));
    int o0[4], o1[4], a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int i0, i1;
        dd_resize_coeff(gx * 4 + k, scale_x, sw, true, i0, i1, a0[k], a1[k]);
        o0[k] = i0 * 3;
        o1[k] = i1 * 3;
    }
    const uint8_t* crop = frames + ((size_t)f * img_h * img_w + (size_t)sy * img_w + sx) * 3;
    const size_t pitch = (size_t)img_w * 3;
    DDHRow S0, S1;
    int c0 = -1, c1 = -1;
    for (int y = y_begin; y < y_end; ++y) {
        int r0, r1, b0, b1;
        dd_resize_coeff(y, scale_y, sh, false, r0, r1, b0, b1);
        if (r0 != c0) {
            if (r0 == c1) S0 = S1;
            else dd_hrow(crop + r0 * pitch, o0, o1, a0, a1, S0);
            c0 = r0;
        }
        if (r1 != c1) {
            if (r1 == c0) S1 = S0;
            else dd_hrow(crop + r1 * pitch, o0, o1, a0, a1, S1);
            c1 = r1;
        }
        uint32_t w[3] = {0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            int v = (((b0 * S0.v[j]) >> 16) + ((b1 * S1.v[j]) >> 16) + 2) >> 2;
            v = dd_imin(dd_imax(v, 0), 255);
            w[j >> 2] |= (uint32_t)v << (8 * (j & 3));
        }
        uint32_t* o = outw + ((size_t)y * pw + gx * 4) * 3 / 4;
        o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
    }
}

// DummyImageEncoder.__call__ (generate_detections.py:92-105): one warp per 16x8x3 patch.
// mean over the channels (exact integer sum, one rounded division), minus 128, L2 normalisation with numpy's
// pairwise float32 sum of the 128 squares (eight strided accumulators, then ((0+1)+(2+3))+((4+5)+(6+7))).
__global__ void __launch_bounds__(128)
k_dummy_encode(const uint8_t* __restrict__ patches, int n, float* __restrict__ out) {
    __shared__ float sq[4][128];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * 4 + wi;
    if (p >= n) return;
    const uint32_t* src = (const uint32_t*)(patches + (size_t)p * 384) + lane * 3;
    const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
    uint8_t by[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) { by[k] = (w0 >> (8 * k)) & 255; by[4 + k] = (w1 >> (8 * k)) & 255; by[8 + k] = (w2 >> (8 * k)) & 255; }
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float s = (float)((int)by[3 * k] + (int)by[3 * k + 1] + (int)by[3 * k + 2]);
        m[k] = dd_subf(dd_divf(s, 3.f), 128.f);
        sq[wi][lane * 4 + k] = dd_mulf(m[k], m[k]);
    }
    __syncwarp();
    float acc = 0.f;
    if (lane < 8) {
        acc = sq[wi][lane];
        for (int k = 1; k < 16; ++k) acc = dd_addf(acc, sq[wi][8 * k + lane]);
    }
    float r = dd_addf(acc, __shfl_down_sync(0xffffffffu, acc, 1));      // lanes 0,2,4,6: r0+r1, r2+r3, ...
    r = dd_addf(r, __shfl_down_sync(0xffffffffu, r, 2));                 // lanes 0,4
    r = dd_addf(r, __shfl_down_sync(0xffffffffu, r, 4));                 // lane 0
    const float l = dd_sqrtf(__shfl_sync(0xffffffffu, r, 0));
    float4 o;
    if (l == 0.f) {
        o = make_float4(m[0], m[1], m[2], m[3]);
        if (lane == 0) o.x = 1.f;
    } else {
        o = make_float4(dd_divf(m[0], l), dd_divf(m[1], l), dd_divf(m[2], l), dd_divf(m[3], l));
    }
    ((float4*)(out + (size_t)p * 128))[lane] = o;
}

extern "C" {

int dd_extract_patches(const uint8_t* frames, int32_t b, int32_t img_h, int32_t img_w, const double* boxes,
                       const int32_t* counts, int32_t dmax, int32_t boxes_are_int, int32_t patch_h,
                       int32_t patch_w, uint8_t* out_patches, int32_t* out_valid, void* stream) {
    if (!frames || !boxes || !out_patches || !out_valid) return DD_ERR_INVALID;
    if (b < 0 || dmax <= 0 || img_h <= 0 || img_w <= 0 || patch_h <= 0 || patch_w <= 0) return DD_ERR_INVALID;
    if ((patch_w & 3) != 0 || patch_w > 4 * DD_PATCH_THREADS) return DD_ERR_INVALID;
    if (((uintptr_t)out_patches & 3) != 0) return DD_ERR_INVALID;
    if ((long long)b * dmax > 0x7fffffffLL / DD_PATCH_THREADS) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    const int xg = patch_w / 4;
    int nb = DD_PATCH_THREADS / xg;
    if (nb > patch_h) nb = patch_h;
    const int rows_per_band = (patch_h + nb - 1) / nb;
    nb = (patch_h + rows_per_band - 1) / rows_per_band;
    const int per_block = DD_PATCH_THREADS / (xg * nb);
    const int n_patches = b * dmax;
    const int blocks = (n_patches + per_block - 1) / per_block;
    k_extract_patches<<<blocks, DD_PATCH_THREADS, 0, (cudaStream_t)stream>>>(
        frames, n_patches, img_h, img_w, boxes, counts, dmax, boxes_are_int, patch_h, patch_w, xg, nb,
        rows_per_band, out_patches, out_valid);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_dummy_encode(const uint8_t* patches, int32_t n, float* out_feat, void* stream) {
    if (!patches || !out_feat || n < 0) return DD_ERR_INVALID;
    if (((uintptr_t)patches & 3) != 0 || ((uintptr_t)out_feat & 15) != 0) return DD_ERR_INVALID;
    if (n == 0) return DD_OK;
    k_dummy_encode<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(patches, n, out_feat);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

}  // extern "C"
