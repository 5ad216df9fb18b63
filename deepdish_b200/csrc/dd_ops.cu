// dd_ops.cu -- stand-alone batched operators behind the per-function deep_sort API
// (KalmanFilter.*, NearestNeighborDistanceMetric.distance, iou_cost, linear_sum_assignment,
// CPython set-difference order, tools.intersection).  Same device functions as the fused tick.
#include <cuda_runtime.h>
#include "dd_tracker_bodies.cuh"

#define DD_WARPS 4
#define DD_CHECK_LAUNCH()                                         \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return DD_ERR_CUDA;               \
    } while (0)

static inline int wblocks(long long n_warps) { return (int)((n_warps + DD_WARPS - 1) / DD_WARPS); }

__global__ void __launch_bounds__(DD_WARPS * 32)
k_kf_initiate(const double* __restrict__ xyah, double* mean, double* cov, int n) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= n) return;
    WarpG g;
    dd_kf_initiate(g, xyah + (size_t)w * 4, mean + (size_t)w * 8, cov + (size_t)w * 64);
}

__global__ void __launch_bounds__(DD_WARPS * 32) k_kf_predict(double* mean, double* cov, int n) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= n) return;
    WarpG g;
    dd_kf_predict(g, mean + (size_t)w * 8, cov + (size_t)w * 64);
}

__global__ void k_kf_project(const double* __restrict__ mean, const double* __restrict__ cov,
                             double* pmean, double* pcov, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double S[16];
    dd_kf_project_cov(mean + (size_t)i * 8, cov + (size_t)i * 64, S);
    for (int k = 0; k < 4; ++k) pmean[(size_t)i * 4 + k] = mean[(size_t)i * 8 + k];
    for (int k = 0; k < 16; ++k) pcov[(size_t)i * 16 + k] = S[k];
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_kf_update(double* mean, double* cov, const double* __restrict__ xyah, int n) {
    __shared__ double scratch[DD_WARPS][64];
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= n) return;
    WarpG g;
    dd_kf_update(g, mean + (size_t)w * 8, cov + (size_t)w * 64, xyah + (size_t)w * 4, scratch[threadIdx.x >> 5]);
}

__global__ void __launch_bounds__(DD_WARPS * 32)
k_kf_gating(const double* __restrict__ mean, const double* __restrict__ cov, const double* __restrict__ xyah,
            int n, int m, int only_position, double* __restrict__ out) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= n) return;
    const int lane = threadIdx.x & 31;
    const double* mu = mean + (size_t)w * 8;
    double S[16], L[16], rinv[4];
    dd_kf_project_cov(mu, cov + (size_t)w * 64, S);
    const double pm[4] = {mu[0], mu[1], mu[2], mu[3]};
    if (only_position) {
        dd_chol<2>(S, L, rinv);
        for (int j = lane; j < m; j += 32) out[(size_t)w * m + j] = dd_maha_sq<2>(L, rinv, pm, xyah + (size_t)j * 4);
    } else {
        dd_chol<4>(S, L, rinv);
        for (int j = lane; j < m; j += 32) out[(size_t)w * m + j] = dd_maha_sq<4>(L, rinv, pm, xyah + (size_t)j * 4);
    }
}

// one warp per (target, query): nn_matching.py:31-54 (normalise rows, 1 - a.b) / :5-28 (pdist).
__global__ void __launch_bounds__(DD_WARPS * 32)
k_nn_distance(const float* __restrict__ gallery, const int* __restrict__ off, const float* __restrict__ feats,
              int n, int m, int metric, double* __restrict__ out) {
    const int w = blockIdx.x * DD_WARPS + (threadIdx.x >> 5);
    if (w >= n * m) return;
    WarpG g;
    const int i = w / m, j = w % m;
    const float4 b = ((const float4*)(feats + (size_t)j * DD_FEAT_DIM))[g.lane];
    float bb = dd_fmaf(b.x, b.x, 0.f); bb = dd_fmaf(b.y, b.y, bb); bb = dd_fmaf(b.z, b.z, bb); bb = dd_fmaf(b.w, b.w, bb);
    bb = g.sum(bb);
    const float bn = dd_sqrtf(bb);
    float4 bh = b;
    if (metric == 0) { bh.x = dd_divf(b.x, bn); bh.y = dd_divf(b.y, bn); bh.z = dd_divf(b.z, bn); bh.w = dd_divf(b.w, bn); }
    float best = 3.0e38f;
    for (int r = off[i]; r < off[i + 1]; ++r) {
        float4 a = ((const float4*)(gallery + (size_t)r * DD_FEAT_DIM))[g.lane];
        float aa = dd_fmaf(a.x, a.x, 0.f); aa = dd_fmaf(a.y, a.y, aa); aa = dd_fmaf(a.z, a.z, aa); aa = dd_fmaf(a.w, a.w, aa);
        aa = g.sum(aa);
        float d;
        if (metric == 0) {
            const float an = dd_sqrtf(aa);
            a.x = dd_divf(a.x, an); a.y = dd_divf(a.y, an); a.z = dd_divf(a.z, an); a.w = dd_divf(a.w, an);
            float p = dd_fmaf(a.x, bh.x, 0.f); p = dd_fmaf(a.y, bh.y, p); p = dd_fmaf(a.z, bh.z, p); p = dd_fmaf(a.w, bh.w, p);
            d = dd_subf(1.0f, g.sum(p));
        } else {
            float p = dd_fmaf(a.x, b.x, 0.f); p = dd_fmaf(a.y, b.y, p); p = dd_fmaf(a.z, b.z, p); p = dd_fmaf(a.w, b.w, p);
            p = g.sum(p);
            d = dd_addf(dd_addf(dd_mulf(-2.0f, p), aa), bb);
            d = d < 0.f ? 0.f : d;
        }
        best = d < best ? d : best;
    }
    if (metric != 0 && best < 0.f) best = 0.f;
    if (g.lane == 0) out[(size_t)i * m + j] = (double)best;
}

__global__ void k_iou_cost(const double* __restrict__ trk, const int* __restrict__ tsu,
                           const double* __restrict__ det, int n, int m, double* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * m) return;
    const int i = e / m, j = e % m;
    if (tsu[i] > 1) { out[e] = DD_INFTY_COST; return; }
    const double x = trk[i * 4], y = trk[i * 4 + 1], w = trk[i * 4 + 2], h = trk[i * 4 + 3];
    const double* b = det + (size_t)j * 4;
    const double tlx = dd_max(x, b[0]), tly = dd_max(y, b[1]);
    const double brx = dd_min(dd_add(x, w), dd_add(b[0], b[2])), bry = dd_min(dd_add(y, h), dd_add(b[1], b[3]));
    const double iw = dd_max(0.0, dd_sub(brx, tlx)), ih = dd_max(0.0, dd_sub(bry, tly));
    const double inter = dd_mul(iw, ih);
    out[e] = dd_sub(1.0, dd_div(inter, dd_sub(dd_add(dd_mul(w, h), dd_mul(b[2], b[3])), inter)));
}

struct DenseCost {
    const double* c;
    int nc;
    __device__ double operator()(int i, int j) const { return c[(size_t)i * nc + j]; }
};
struct DenseCostT {
    const double* c;
    int nc;
    __device__ double operator()(int i, int j) const { return c[(size_t)j * nc + i]; }
};

__global__ void __launch_bounds__(32)
k_lsap(const double* __restrict__ cost, int nr, int nc, int* __restrict__ out_col4row, int* __restrict__ out_status) {
    extern __shared__ __align__(16) char smem[];
    WarpG g;
    const int p = blockIdx.x;
    const int n = nr > nc ? nr : nc;
    DDLsapScratch s;
    dd_lsap_carve(smem, n, s);
    const double* c = cost + (size_t)p * nr * nc;
    int rc;
    if (nc < nr) {
        DenseCostT f{c, nc};
        rc = dd_lsap_solve(g, nc, nr, f, s);
        for (int r = g.lane; r < nr; r += 32) out_col4row[(size_t)p * nr + r] = rc ? -1 : s.row4col[r];
    } else {
        DenseCost f{c, nc};
        rc = dd_lsap_solve(g, nr, nc, f, s);
        for (int r = g.lane; r < nr; r += 32) out_col4row[(size_t)p * nr + r] = rc ? -1 : s.col4row[r];
    }
    if (g.lane == 0) out_status[p] = rc ? 1 : 0;
}

__global__ void __launch_bounds__(32)
k_set_diff(const int* __restrict__ a, const int* __restrict__ na, int na_max, const int* __restrict__ mm,
           const int* __restrict__ nm, int nm_max, int* __restrict__ out, int* __restrict__ out_n, int cap) {
    extern __shared__ __align__(16) char smem[];
    const int p = blockIdx.x;
    short* av = (short*)smem;
    short* ov = av + na_max;
    short* tA = ov + na_max;
    short* tB = tA + cap;
    short* tC = tB + cap;
    unsigned char* flag = (unsigned char*)(tC + cap);
    const int lane = threadIdx.x;
    const int n = na[p], k = nm[p];
    for (int i = lane; i < 1024; i += 32) flag[i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) av[i] = (short)a[(size_t)p * na_max + i];
    for (int i = lane; i < k; i += 32) flag[mm[(size_t)p * nm_max + i] & 1023] = 1;
    __syncwarp();
    int cnt = 0;
    if (lane == 0) {
        bool contig = true;
        for (int i = 0; i < n; ++i) contig = contig && av[i] == i;
        if (contig) {                       // the path the tracker takes (dd_match_stream)
            short* surv = tC;
            int ns = 0;
            for (int i = 0; i < n; ++i) if (!flag[i]) surv[ns++] = (short)i;
            if ((n >> 2) > k) { for (int i = 0; i < ns; ++i) ov[i] = surv[i]; cnt = ns; }
            else cnt = dd_set_order_from_survivors(surv, ns, ov, tA, tB);
        } else {
            cnt = dd_set_difference_order_serial(av, n, flag, k, ov, tA, tB, tC, cap);
        }
    }
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) out[(size_t)p * na_max + i] = ov[i];
    if (lane == 0) out_n[p] = cnt;
}

__global__ void k_intersection(const double* __restrict__ seg, int n, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = seg + (size_t)i * 8;
    out[i] = dd_segments_intersect(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]) ? 1 : 0;
}

extern "C" {

int dd_kalman_initiate(const double* xyah, double* mean, double* cov, int32_t n, void* stream) {
    if (!xyah || !mean || !cov || n < 0) return DD_ERR_INVALID;
    if (n == 0) return DD_OK;
    k_kf_initiate<<<wblocks(n), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(xyah, mean, cov, n);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_kalman_predict(double* mean, double* cov, int32_t n, void* stream) {
    if (!mean || !cov || n < 0) return DD_ERR_INVALID;
    if (n == 0) return DD_OK;
    k_kf_predict<<<wblocks(n), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(mean, cov, n);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_kalman_project(const double* mean, const double* cov, double* pmean, double* pcov, int32_t n, void* stream) {
    if (!mean || !cov || !pmean || !pcov || n < 0) return DD_ERR_INVALID;
    if (n == 0) return DD_OK;
    k_kf_project<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mean, cov, pmean, pcov, n);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_kalman_update(double* mean, double* cov, const double* xyah, int32_t n, void* stream) {
    if (!mean || !cov || !xyah || n < 0) return DD_ERR_INVALID;
    if (n == 0) return DD_OK;
    k_kf_update<<<wblocks(n), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(mean, cov, xyah, n);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_kalman_gating_distance(const double* mean, const double* cov, const double* xyah, int32_t n,
                              int32_t m, int32_t only_position, double* out, void* stream) {
    if (!mean || !cov || !xyah || !out || n < 0 || m < 0) return DD_ERR_INVALID;
    if (n == 0 || m == 0) return DD_OK;
    k_kf_gating<<<wblocks(n), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(mean, cov, xyah, n, m, only_position, out);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_nn_distance(const float* gallery, const int32_t* gal_offsets, const float* feats, int32_t n,
                   int32_t m, int32_t metric, double* out, void* stream) {
    if (!gal_offsets || !feats || !out || n < 0 || m < 0 || (metric != 0 && metric != 1)) return DD_ERR_INVALID;
    if (n == 0 || m == 0) return DD_OK;
    if (!gallery) return DD_ERR_INVALID;
    k_nn_distance<<<wblocks((long long)n * m), DD_WARPS * 32, 0, (cudaStream_t)stream>>>(gallery, gal_offsets, feats, n, m, metric, out);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_iou_cost(const double* track_tlwh, const int32_t* tsu, const double* det_tlwh, int32_t n,
                int32_t m, double* out, void* stream) {
    if (!track_tlwh || !tsu || !det_tlwh || !out || n < 0 || m < 0) return DD_ERR_INVALID;
    if (n == 0 || m == 0) return DD_OK;
    k_iou_cost<<<(n * m + 127) / 128, 128, 0, (cudaStream_t)stream>>>(track_tlwh, tsu, det_tlwh, n, m, out);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_lsap(const double* cost, int32_t b, int32_t nr, int32_t nc, int32_t* out_col4row,
            int32_t* out_status, void* stream) {
    if (!cost || !out_col4row || !out_status || b < 0 || nr <= 0 || nc <= 0) return DD_ERR_INVALID;
    if (nr > 4096 || nc > 4096) return DD_ERR_CAPACITY;
    if (b == 0) return DD_OK;
    const size_t smem = dd_lsap_scratch_bytes(nr > nc ? nr : nc);
    if (smem > 227 * 1024) return DD_ERR_CAPACITY;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_lsap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DD_ERR_CUDA;
    k_lsap<<<b, 32, smem, (cudaStream_t)stream>>>(cost, nr, nc, out_col4row, out_status);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_set_difference_order(const int32_t* a, const int32_t* na, int32_t na_max, const int32_t* m,
                            const int32_t* nm, int32_t nm_max, int32_t b, int32_t* out,
                            int32_t* out_n, void* stream) {
    if (!a || !na || !m || !nm || !out || !out_n || b < 0 || na_max <= 0 || nm_max <= 0) return DD_ERR_INVALID;
    if (na_max > 1024) return DD_ERR_CAPACITY;
    if (b == 0) return DD_OK;
    const int cap = dd_set_table_slots(na_max);
    const size_t smem = (size_t)na_max * 4 + (size_t)cap * 6 + 1024;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_set_diff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DD_ERR_CUDA;
    k_set_diff<<<b, 32, smem, (cudaStream_t)stream>>>(a, na, na_max, m, nm, nm_max, out, out_n, cap);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

int dd_intersection(const double* seg, int32_t n, int32_t* out, void* stream) {
    if (!seg || !out || n < 0) return DD_ERR_INVALID;
    if (n == 0) return DD_OK;
    k_intersection<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(seg, n, out);
    DD_CHECK_LAUNCH();
    return DD_OK;
}

}  // extern "C"
