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

#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

#define DD_YOLO_ROWS 128

struct RowF32 {
    const float* p;
    __device__ float operator()(int k) const { return p[k]; }
};
struct RowU8 {
    const unsigned char* p;
    float scale, zp;
    __device__ float operator()(int k) const { return dd_mulf(dd_subf((float)p[k], zp), scale); }
};

template <bool U8>
__global__ void __launch_bounds__(DD_YOLO_ROWS)
k_yolo_decode(const void* __restrict__ head, float scale, int zero_point, int na, DDYoloParams P,
              const unsigned char* __restrict__ wanted, int ncap, double* __restrict__ out_tlwh,
              float* __restrict__ out_score, int* __restrict__ out_class, int* __restrict__ out_anchor,
              int* __restrict__ out_count, int* __restrict__ out_flags) {
    extern __shared__ __align__(16) char smem[];
    const int frame = blockIdx.y;
    const int row0 = blockIdx.x * DD_YOLO_ROWS;
    const int rows = min(DD_YOLO_ROWS, na - row0);
    const int rw = 5 + P.nc;
    const size_t esz = U8 ? 1 : 4;
    const size_t row_bytes = (size_t)rw * esz;
    const char* src = (const char*)head + ((size_t)frame * na + row0) * row_bytes;
    const size_t bytes = (size_t)rows * row_bytes;
    if ((((uintptr_t)src) & 15) == 0 && (bytes & 15) == 0) {
        const int4* s4 = (const int4*)src;
        int4* d4 = (int4*)smem;
        const int n4 = (int)(bytes >> 4);
#pragma unroll 4
        for (int i = threadIdx.x; i < n4; i += DD_YOLO_ROWS) d4[i] = __ldg(s4 + i);
    } else if (U8) {
        for (int i = threadIdx.x; i < (int)bytes; i += DD_YOLO_ROWS) smem[i] = src[i];
    } else {
        const float* sf = (const float*)src;
        float* df = (float*)smem;
        for (int i = threadIdx.x; i < rows * rw; i += DD_YOLO_ROWS) df[i] = sf[i];
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
__global__ void __launch_bounds__(256)
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
    const int P = dd_next_pow2(n);
    unsigned long long* keys = (unsigned long long*)smem;
    double* tl = (double*)(keys + P);
    float* sc = (float*)(tl + (size_t)n * 4);
    int* cl = (int*)(sc + n);
    const size_t base = (size_t)frame * ncap;
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        keys[i] = i < n ? (((unsigned long long)(unsigned)out_anchor[base + i]) << 32) | (unsigned)i : ~0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        tl[i * 4 + 0] = out_tlwh[(base + i) * 4 + 0]; tl[i * 4 + 1] = out_tlwh[(base + i) * 4 + 1];
        tl[i * 4 + 2] = out_tlwh[(base + i) * 4 + 2]; tl[i * 4 + 3] = out_tlwh[(base + i) * 4 + 3];
        sc[i] = out_score[base + i];
        cl[i] = out_class[base + i];
    }
    __syncthreads();
    dd_bitonic_sort(g, keys, P);
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int i = (int)(keys[r] & 0xffffffffu);
        out_tlwh[(base + r) * 4 + 0] = tl[i * 4 + 0]; out_tlwh[(base + r) * 4 + 1] = tl[i * 4 + 1];
        out_tlwh[(base + r) * 4 + 2] = tl[i * 4 + 2]; out_tlwh[(base + r) * 4 + 3] = tl[i * 4 + 3];
        out_score[base + r] = sc[i];
        out_class[base + r] = cl[i];
        out_anchor[base + r] = (int)(keys[r] >> 32);
    }
}

__global__ void __launch_bounds__(256)
k_nms(const double* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ counts,
      int nmax, double max_overlap, int* __restrict__ out_keep, int* __restrict__ out_nkeep) {
    extern __shared__ __align__(16) char smem[];
    const int f = blockIdx.x;
    BlockG g;
    dd_nms_frame(g, boxes + (size_t)f * nmax * 4, scores + (size_t)f * nmax, counts[f], nmax,
                 max_overlap, out_keep + (size_t)f * nmax, out_nkeep + f, smem);
}

extern "C" {

int dd_nms(const double* boxes, const float* scores, const int32_t* counts, int32_t b, int32_t nmax,
           double max_overlap, int32_t* out_keep, int32_t* out_nkeep, void* stream) {
    if (!boxes || !scores || !counts || !out_keep || !out_nkeep || b < 0 || nmax <= 0) return DD_ERR_INVALID;
    if (b == 0) return DD_OK;
    const size_t smem = dd_nms_smem_bytes(nmax);
    if (smem > 227 * 1024) return DD_ERR_CAPACITY;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_nms, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DD_ERR_CUDA;
    k_nms<<<b, 256, smem, (cudaStream_t)stream>>>(boxes, scores, counts, nmax, max_overlap, out_keep, out_nkeep);
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
    const int tiles = (na + DD_YOLO_ROWS - 1) / DD_YOLO_ROWS;
    const size_t smem = (size_t)DD_YOLO_ROWS * (5 + nc) * (head_is_u8 ? 1 : 4);
    if (smem > 227 * 1024) return DD_ERR_INVALID;
    dim3 grid(tiles, b);
    if (head_is_u8) {
        if (smem > 48 * 1024 && cudaFuncSetAttribute(k_yolo_decode<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DD_ERR_CUDA;
        k_yolo_decode<true><<<grid, DD_YOLO_ROWS, smem, st>>>(head, scale, zero_point, na, P, wanted, ncap, out_tlwh, out_score, out_class, out_anchor, out_count, out_flags);
    } else {
        if (smem > 48 * 1024 && cudaFuncSetAttribute(k_yolo_decode<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return DD_ERR_CUDA;
        k_yolo_decode<false><<<grid, DD_YOLO_ROWS, smem, st>>>(head, scale, zero_point, na, P, wanted, ncap, out_tlwh, out_score, out_class, out_anchor, out_count, out_flags);
    }
    DD_CHECK_LAUNCH();
    const size_t osm = (size_t)dd_next_pow2(ncap) * 8 + (size_t)ncap * (32 + 4 + 4);
    if (osm > 48 * 1024 && cudaFuncSetAttribute(k_yolo_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osm) != cudaSuccess) return DD_ERR_CUDA;
    k_yolo_order<<<b, 256, osm, st>>>(ncap, out_tlwh, out_score, out_class, out_anchor, out_count, out_flags);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

}  // extern "C"
